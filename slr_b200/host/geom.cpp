// See geom.h for the parity notes and reference citations.
#include "geom.h"
#include <algorithm>
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


// ---------------------------------------------------------------------------------------------
// motion: quaternions, polar decomposition, AnimatedTransform (Quaternion.h/.cpp, Transform.h:89-144)
// ---------------------------------------------------------------------------------------------
static Mat4 transpose(const Mat4& m) { return Mat4(m.row(0), m.row(1), m.row(2), m.row(3)); }

Quat::Quat(const Mat4& m) {
    auto e = [&m](int c, int r) { return m.c[c][r]; };      // the reference indexes m[column][row]
    const float trace = e(0, 0) + e(1, 1) + e(2, 2);
    if (trace > 0.0f) {
        const float s = std::sqrt(trace + 1.0f);
        const float k = 0.5f / s;
        x = k * (e(1, 2) - e(2, 1)); y = k * (e(2, 0) - e(0, 2)); z = k * (e(0, 1) - e(1, 0));
        w = s / 2.0f;
    } else {
        const int nxt[3] = {1, 2, 0};
        float q[3];
        int i = 0;
        if (e(1, 1) > e(0, 0)) i = 1;
        if (e(2, 2) > e(i, i)) i = 2;
        const int j = nxt[i], k = nxt[j];
        float s = std::sqrt((e(i, i) - (e(j, j) + e(k, k))) + 1.0f);
        q[i] = s * 0.5f;
        if (s != 0.0f) s = 0.5f / s;
        w = (e(j, k) - e(k, j)) * s;
        q[j] = (e(i, j) + e(j, i)) * s;
        q[k] = (e(i, k) + e(k, i)) * s;
        x = q[0]; y = q[1]; z = q[2];
    }
}

Mat4 Quat::toMatrix() const {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, zx = z * x, xw = x * w, yw = y * w, zw = z * w;
    return Mat4(Vec4(1 - 2 * (yy + zz), 2 * (xy + zw), 2 * (zx - yw), 0.0f), Vec4(2 * (xy - zw), 1 - 2 * (xx + zz), 2 * (yz + xw), 0.0f),
                Vec4(2 * (zx + yw), 2 * (yz - xw), 1 - 2 * (xx + yy), 0.0f), Vec4(0, 0, 0, 1));
}

Quat slerp(float t, const Quat& q0, const Quat& q1) {
    auto qdot = [](const Quat& a, const Quat& b) { return (a.x * b.x + a.y * b.y + a.z * b.z) + a.w * b.w; };
    auto qnorm = [&qdot](const Quat& q) { const float r = 1.0f / std::sqrt(qdot(q, q)); return Quat(q.x * r, q.y * r, q.z * r, q.w * r); };
    const float cosTheta = qdot(q0, q1);
    if (cosTheta > 0.9995f) {
        const float a = 1 - t;
        return qnorm(Quat(a * q0.x + t * q1.x, a * q0.y + t * q1.y, a * q0.z + t * q1.z, a * q0.w + t * q1.w));
    }
    const float theta = std::acos(std::fmin(std::fmax(cosTheta, -1.0f), 1.0f));
    const float thetap = theta * t;
    const Quat qPerp = qnorm(Quat(q1.x - q0.x * cosTheta, q1.y - q0.y * cosTheta, q1.z - q0.z * cosTheta, q1.w - q0.w * cosTheta));
    const float c = std::cos(thetap), sn = std::sin(thetap);
    return Quat(q0.x * c + qPerp.x * sn, q0.y * c + qPerp.y * sn, q0.z * c + qPerp.z * sn, q0.w * c + qPerp.w * sn);
}

void decompose(const Mat4& mat, Vec3* T, Quat* R, Mat4* S) {
    T->x = mat.c[3][0]; T->y = mat.c[3][1]; T->z = mat.c[3][2];
    Mat4 matRS = mat;
    for (int i = 0; i < 3; ++i) { matRS.c[3][i] = 0.0f; matRS.c[i][3] = 0.0f; }
    matRS.c[3][3] = 1.0f;
    float norm;
    int count = 0;
    Mat4 curR = matRS;
    do {
        const Mat4 itR = invert(transpose(curR));
        Mat4 nextR;
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) nextR.c[c][r] = 0.5f * (curR.c[c][r] + itR.c[c][r]);
        norm = 0;
        for (int i = 0; i < 3; ++i) {
            const float n = std::fabs(curR.c[0][i] - nextR.c[0][i]) + std::fabs(curR.c[1][i] - nextR.c[1][i]) + std::fabs(curR.c[2][i] - nextR.c[2][i]);
            norm = std::max(norm, n);
        }
        curR = nextR;
    } while (++count < 100 && norm > 0.0001);
    *R = Quat(curR);
    *S = invert(curR) * matRS;
}

AnimatedTransform::AnimatedTransform(const StaticTransform& b, const StaticTransform& e, float tb, float te)
    : begin(b.mat, b.matInv), end(e.mat, e.matInv), tBegin(tb), tEnd(te) {
    decompose(begin.mat, &T[0], &R[0], &S[0]);
    decompose(end.mat, &T[1], &R[1], &S[1]);
}

StaticTransform AnimatedTransform::sample(float time) const {
    if (time <= tBegin) return begin;
    if (time >= tEnd) return end;
    const float t = (time - tBegin) / (tEnd - tBegin);
    const Vec3 trans = (1 - t) * T[0] + t * T[1];
    const Quat rot = slerp(t, R[0], R[1]);
    Mat4 sc;
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) sc.c[c][r] = (1 - t) * S[0].c[c][r] + t * S[1].c[c][r];
    return StaticTransform(translate(trans.x, trans.y, trans.z) * rot.toMatrix() * sc);
}

BBox AnimatedTransform::motionBounds(const BBox& b) const {
    BBox ret;
    const uint32_t numIte = 128;
    for (uint32_t i = 0; i < numIte; ++i) {
        const float t = (float)i / (numIte - 1);
        const float sTime = (1 - t) * tBegin + t * tEnd;
        ret.grow(transformBounds(sample(sTime).mat, b));
    }
    return ret;
}

std::shared_ptr<const AnimatedTransform> AnimatedTransform::mulLeft(const StaticTransform& s) const {
    return std::make_shared<AnimatedTransform>(s * begin, s * end, tBegin, tEnd);
}
std::shared_ptr<const AnimatedTransform> AnimatedTransform::mulRight(const StaticTransform& s) const {
    return std::make_shared<AnimatedTransform>(begin * s, end * s, tBegin, tEnd);
}

}  // namespace slr
