// Host-side fp32 geometry primitives for the B200 SLR path.
//
// The arithmetic (operation order, IEEE division via reciprocal-then-multiply, fmin/fmax bounds)
// deliberately follows the reference's BasicTypes so that vertices baked on the host, bounding
// boxes and therefore the SBVH/QBVH trees come out bit-identical to libSLR's:
//   Vector3 ops        libSLR/BasicTypes/Vector3.h:24-124
//   Point3 ops         libSLR/BasicTypes/Point3.h:30-113
//   Matrix4x4 ops      libSLR/BasicTypes/Matrix4x4.h:61-81, Matrix4x4.cpp:36-139
//   BoundingBox3D      libSLR/Core/geometry.h:36-138
// Build without FMA contraction (no -march=native / -mfma): the reference is x86-64 baseline.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>

namespace slr {

struct Vec3 {
    float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    explicit Vec3(float v) : x(v), y(v), z(v) {}
    Vec3(float xx, float yy, float zz) : x(xx), y(yy), z(zz) {}
    float operator[](unsigned i) const { return (&x)[i]; }
    float& operator[](unsigned i) { return (&x)[i]; }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    Vec3 operator+(const Vec3& v) const { return Vec3(x + v.x, y + v.y, z + v.z); }
    Vec3 operator-(const Vec3& v) const { return Vec3(x - v.x, y - v.y, z - v.z); }
    Vec3 operator*(float s) const { return Vec3(x * s, y * s, z * s); }
    // division is "multiply by the rounded reciprocal", as in the reference (Vector3.h:32)
    Vec3 operator/(float s) const { float r = 1.0f / s; return Vec3(x * r, y * r, z * r); }
    bool operator==(const Vec3& v) const { return x == v.x && y == v.y && z == v.z; }
    float sqLength() const { return x * x + y * y + z * z; }
    float length() const { return std::sqrt(x * x + y * y + z * z); }
};
inline Vec3 operator*(float s, const Vec3& v) { return Vec3(s * v.x, s * v.y, s * v.z); }
inline float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vec3 normalize(const Vec3& v) { float l = v.length(); return v / l; }
inline Vec3 vmin(const Vec3& a, const Vec3& b) { return Vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline Vec3 vmax(const Vec3& a, const Vec3& b) { return Vec3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }

struct Vec4 {
    float x, y, z, w;
    Vec4() : x(0), y(0), z(0), w(0) {}
    Vec4(float xx, float yy, float zz, float ww) : x(xx), y(yy), z(zz), w(ww) {}
    float operator[](unsigned i) const { return (&x)[i]; }
    float& operator[](unsigned i) { return (&x)[i]; }
    bool operator==(const Vec4& v) const { return x == v.x && y == v.y && z == v.z && w == v.w; }
};
inline float dot(const Vec4& a, const Vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

struct Vec2 {
    float u, v;
    Vec2() : u(0), v(0) {}
    Vec2(float uu, float vv) : u(uu), v(vv) {}
};

// Column-major 4x4, columns c[0..3]; element (row r, col c) is c[c][r] (Matrix4x4.h:23-38).
struct Mat4 {
    Vec4 c[4];
    Mat4() {}
    Mat4(const Vec4& c0, const Vec4& c1, const Vec4& c2, const Vec4& c3) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
    static Mat4 identity() { return Mat4(Vec4(1, 0, 0, 0), Vec4(0, 1, 0, 0), Vec4(0, 0, 1, 0), Vec4(0, 0, 0, 1)); }
    float at(unsigned r, unsigned col) const { return c[col][r]; }
    float& at(unsigned r, unsigned col) { return c[col][r]; }
    Vec4 row(unsigned r) const { return Vec4(c[0][r], c[1][r], c[2][r], c[3][r]); }
    bool isIdentity() const { Mat4 i = identity(); return c[0] == i.c[0] && c[1] == i.c[1] && c[2] == i.c[2] && c[3] == i.c[3]; }
    bool operator==(const Mat4& m) const { return c[0] == m.c[0] && c[1] == m.c[1] && c[2] == m.c[2] && c[3] == m.c[3]; }

    Mat4 operator*(const Mat4& m) const {
        Vec4 r[4] = {row(0), row(1), row(2), row(3)};
        Mat4 o;
        for (int j = 0; j < 4; ++j) o.c[j] = Vec4(dot(r[0], m.c[j]), dot(r[1], m.c[j]), dot(r[2], m.c[j]), dot(r[3], m.c[j]));
        return o;
    }
    // direction: upper 3x3 only (Matrix4x4.h:71-73)
    Vec3 mulVector(const Vec3& v) const {
        return Vec3(c[0].x * v.x + c[1].x * v.y + c[2].x * v.z,
                    c[0].y * v.x + c[1].y * v.y + c[2].y * v.z,
                    c[0].z * v.x + c[1].z * v.y + c[2].z * v.z);
    }
    // point: homogeneous with w-divide only when w != 1 (Matrix4x4.h:75-81)
    Vec3 mulPoint(const Vec3& p) const {
        Vec4 ph(p.x, p.y, p.z, 1.0f);
        Vec4 t(dot(row(0), ph), dot(row(1), ph), dot(row(2), ph), dot(row(3), ph));
        if (t.w != 1.0f) { float r = 1.0f / t.w; t.x *= r; t.y *= r; t.z *= r; }
        return Vec3(t.x, t.y, t.z);
    }
};

Mat4 invert(const Mat4& m);                    // Gauss-Jordan with full pivoting, Matrix4x4.cpp:36-88
Mat4 translate(float x, float y, float z);     // Matrix4x4.h:224-236
Mat4 scale(float x, float y, float z);         // Matrix4x4.h:204-222
Mat4 rotate(float angle, const Vec3& axis);    // Matrix4x4.cpp:111-137
Mat4 lookAt(const Vec3& eye, const Vec3& tgt, const Vec3& up);  // Matrix4x4.cpp:92-107

struct AnimatedTransform;

// mat + cached inverse; normals go through the transposed inverse (Transform.h:38-51).
// `anim` (normally null) makes the value stand for an AnimatedTransform whose key frame at the begin time is mat / matInv:
// this is how the scene language's Transform values and InternalNode::setTransform carry motion (Transform.h:89-144)
// through code that is otherwise static. operator* and inverse() are defined for static values only.
struct StaticTransform {
    Mat4 mat, matInv;
    std::shared_ptr<const AnimatedTransform> anim;
    StaticTransform() : mat(Mat4::identity()), matInv(Mat4::identity()) {}
    explicit StaticTransform(const Mat4& m) : mat(m), matInv(invert(m)) {}
    StaticTransform(const Mat4& m, const Mat4& mi) : mat(m), matInv(mi) {}
    Vec3 point(const Vec3& p) const { return mat.mulPoint(p); }
    Vec3 vector(const Vec3& v) const { return mat.mulVector(v); }
    Vec3 normal(const Vec3& n) const {
        return Vec3(matInv.at(0, 0) * n.x + matInv.at(1, 0) * n.y + matInv.at(2, 0) * n.z,
                    matInv.at(0, 1) * n.x + matInv.at(1, 1) * n.y + matInv.at(2, 1) * n.z,
                    matInv.at(0, 2) * n.x + matInv.at(1, 2) * n.y + matInv.at(2, 2) * n.z);
    }
    StaticTransform operator*(const StaticTransform& t) const { return StaticTransform(mat * t.mat); }
    StaticTransform inverse() const { return StaticTransform(matInv, mat); }
    bool isIdentity() const { return mat.isIdentity() && !anim; }
};

// Quaternion (libSLR/BasicTypes/Quaternion.h:17-127): built from a rotation matrix, toMatrix, Slerp.
struct Quat {
    float x = 0, y = 0, z = 0, w = 1;
    Quat() {}
    Quat(float xx, float yy, float zz, float ww) : x(xx), y(yy), z(zz), w(ww) {}
    explicit Quat(const Mat4& m);
    Mat4 toMatrix() const;
};
Quat slerp(float t, const Quat& q0, const Quat& q1);
// polar decomposition M = translate(T) * R * S by the iteration R <- (R + (R^T)^-1) / 2 (Quaternion.cpp:15-43)
void decompose(const Mat4& m, Vec3* T, Quat* R, Mat4* S);

enum Axis : uint8_t { Axis_X = 0, Axis_Y = 1, Axis_Z = 2 };

struct BBox {
    Vec3 lo, hi;
    BBox() : lo(INFINITY), hi(-INFINITY) {}
    explicit BBox(const Vec3& p) : lo(p), hi(p) {}
    BBox(const Vec3& l, const Vec3& h) : lo(l), hi(h) {}
    BBox& grow(const Vec3& p) { lo = vmin(lo, p); hi = vmax(hi, p); return *this; }
    BBox& grow(const BBox& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); return *this; }
    Vec3 centroid() const { return (lo + hi) * 0.5f; }
    float centerOf(Axis a) const { return (lo[a] + hi[a]) * 0.5f; }
    float width(Axis a) const { return hi[a] - lo[a]; }
    float surfaceArea() const { Vec3 d = hi - lo; return 2 * (d.x * d.y + d.y * d.z + d.z * d.x); }
    Axis widestAxis() const {
        Vec3 d = hi - lo;
        if (d.x > d.y && d.x > d.z) return Axis_X;
        return d.y > d.z ? Axis_Y : Axis_Z;
    }
    bool isValid() const { Vec3 d = hi - lo; return d.x >= 0 && d.y >= 0 && d.z >= 0; }
};
inline BBox intersection(const BBox& a, const BBox& b) { return BBox(vmax(a.lo, b.lo), vmin(a.hi, b.hi)); }
BBox transformBounds(const Mat4& m, const BBox& b);  // 8-corner transform, Transform.h:54-65

// AnimatedTransform (libSLR/Core/Transform.h:89-144): two key frames, decomposed into translation, rotation and
// scale / shear; between the key times the three parts are interpolated (lerp, Slerp, lerp) and recomposed.
struct AnimatedTransform {
    StaticTransform begin, end;
    float tBegin, tEnd;
    Vec3 T[2];
    Quat R[2];
    Mat4 S[2];
    AnimatedTransform(const StaticTransform& b, const StaticTransform& e, float tb, float te);
    bool isStatic() const { return begin.mat == end.mat; }
    StaticTransform sample(float time) const;
    // union of the transformed box at 128 times across [tBegin, tEnd] (Transform.h:131-143: sampling, not a guaranteed bound)
    BBox motionBounds(const BBox& b) const;
    // static * this and this * static (Transform.cpp:67-72)
    std::shared_ptr<const AnimatedTransform> mulLeft(const StaticTransform& s) const;
    std::shared_ptr<const AnimatedTransform> mulRight(const StaticTransform& s) const;
};

// One mesh vertex as the reference stores it (geometry.h:148-156): 44 bytes.
struct Vertex {
    Vec3 position, normal, tangent;
    Vec2 texCoord;
};

}  // namespace slr
