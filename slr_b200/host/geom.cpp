// See geom.h for the parity notes and reference citations.
#include "geom.h"
#include <utility>

namespace slr {

// Full-pivot Gauss-Jordan elimination done in place on a copy, recording the row swaps and undoing
// them as column swaps at the end. Same pivot search order (columns outer, rows inner, strict >)
// and the same update arithmetic as Matrix4x4.cpp:36-88 so inverse matrices are bit-identical.
Mat4 invert(const Mat4& m) {
    Mat4 a = m;
    bool done[4] = {false, false, false, false};
    int swapR[4] = {0, 0, 0, 0}, swapC[4] = {0, 0, 0, 0};
    auto swapRows = [&a](int r0, int r1) {
        if (r0 == r1) return;
        for (int col = 0; col < 4; ++col) std::swap(a.c[col][r0], a.c[col][r1]);
    };
    for (int pass = 0; pass < 4; ++pass) {
        int pc = 0, pr = 0;
        float best = -1.0f;
        for (int col = 0; col < 4; ++col) {
            if (done[col]) continue;
            for (int r = 0; r < 4; ++r) {
                if (done[r]) continue;
                float v = std::fabs(a.c[col][r]);
                if (v > best) { pc = col; pr = r; best = v; }
            }
        }
        swapRows(pr, pc);
        swapR[pass] = pr; swapC[pass] = pc;
        float pivot = a.c[pc][pc];
        if (pivot == 0.0f) {
            float nan = std::numeric_limits<float>::quiet_NaN();
            Vec4 n(nan, nan, nan, nan);
            return Mat4(n, n, n, n);
        }
        a.c[pc][pc] = 1.0f;
        float s = 1.0f / pivot;
        for (int col = 0; col < 4; ++col) a.c[col][pc] *= s;
        Vec4 pivotRow = a.row(pc);
        for (int r = 0; r < 4; ++r) {
            if (r == pc) continue;
            float f = a.c[pc][r];
            a.c[pc][r] = 0.0f;
            float nf = -f;
            for (int col = 0; col < 4; ++col) a.c[col][r] += nf * pivotRow[col];
        }
        done[pc] = true;
    }
    for (int pass = 3; pass >= 0; --pass)
        if (swapR[pass] != swapC[pass]) std::swap(a.c[swapR[pass]], a.c[swapC[pass]]);
    return a;
}

Mat4 translate(float x, float y, float z) {
    return Mat4(Vec4(1, 0, 0, 0), Vec4(0, 1, 0, 0), Vec4(0, 0, 1, 0), Vec4(x, y, z, 1.0f));
}

Mat4 scale(float x, float y, float z) {
    // scalar * unit column, so a negative scale yields -0.0 in the off-diagonal slots like the reference
    return Mat4(Vec4(x * 1.0f, x * 0.0f, x * 0.0f, x * 0.0f),
                Vec4(y * 0.0f, y * 1.0f, y * 0.0f, y * 0.0f),
                Vec4(z * 0.0f, z * 0.0f, z * 1.0f, z * 0.0f),
                Vec4(0, 0, 0, 1));
}

Mat4 rotate(float angle, const Vec3& axis) {
    Vec3 n = normalize(axis);
    float c = std::cos(angle), s = std::sin(angle);
    float omc = 1 - c;
    Mat4 r;
    r.at(0, 0) = n.x * n.x * omc + c;
    r.at(1, 0) = n.x * n.y * omc + n.z * s;
    r.at(2, 0) = n.z * n.x * omc - n.y * s;
    r.at(0, 1) = n.x * n.y * omc - n.z * s;
    r.at(1, 1) = n.y * n.y * omc + c;
    r.at(2, 1) = n.y * n.z * omc + n.x * s;
    r.at(0, 2) = n.z * n.x * omc + n.y * s;
    r.at(1, 2) = n.y * n.z * omc - n.x * s;
    r.at(2, 2) = n.z * n.z * omc + c;
    r.at(3, 3) = 1.0f;
    return r;
}

Mat4 lookAt(const Vec3& eye, const Vec3& tgt, const Vec3& up) {
    Vec3 z = normalize(eye - tgt);
    Vec3 x = normalize(cross(up, z));
    Vec3 y = cross(z, x);
    return Mat4(Vec4(x.x, y.x, z.x, 0), Vec4(x.y, y.y, z.y, 0), Vec4(x.z, y.z, z.z, 0),
                Vec4(-dot(eye, x), -dot(eye, y), -dot(eye, z), 1.0f));
}

BBox transformBounds(const Mat4& m, const BBox& b) {
    BBox r;
    for (int i = 0; i < 8; ++i)  // x outermost, z innermost
        r.grow(m.mulPoint(Vec3((i & 4) ? b.hi.x : b.lo.x, (i & 2) ? b.hi.y : b.lo.y, (i & 1) ? b.hi.z : b.lo.z)));
    return r;
}

}  // namespace slr
