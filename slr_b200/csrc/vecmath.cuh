// Small fp32 vector helpers for the shading kernels (Vector3D / Normal3D / ReferenceFrame of
// libSLR/BasicTypes/Vector3.h, libSLR/Core/geometry.h:224-236).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace slrgpu {

constexpr float kPi = 3.14159265358979323846f;

struct V3 {
    float x, y, z;
    __device__ __forceinline__ V3() {}
    __device__ __forceinline__ V3(float a, float b, float c) : x(a), y(b), z(c) {}
    __device__ __forceinline__ V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    __device__ __forceinline__ V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    __device__ __forceinline__ V3 operator-() const { return V3(-x, -y, -z); }
    __device__ __forceinline__ V3 operator*(float s) const { return V3(x * s, y * s, z * s); }
    __device__ __forceinline__ V3 operator/(float s) const { const float r = 1.0f / s; return V3(x * r, y * r, z * r); }
};
__device__ __forceinline__ V3 operator*(float s, const V3& v) { return V3(v.x * s, v.y * s, v.z * s); }
__device__ __forceinline__ float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float absDot(const V3& a, const V3& b) { return fabsf(dot(a, b)); }
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float sqLength(const V3& a) { return dot(a, a); }
__device__ __forceinline__ float length(const V3& a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(const V3& a) { return a / length(a); }
// halfVector(a, b) = normalize(a + b)
__device__ __forceinline__ V3 halfVector(const V3& a, const V3& b) { return normalize(a + b); }

struct Frame {
    V3 x, y, z;
    __device__ __forceinline__ V3 toLocal(const V3& v) const { return V3(dot(x, v), dot(y, v), dot(z, v)); }
    __device__ __forceinline__ V3 fromLocal(const V3& v) const {
        return V3(x.x * v.x + y.x * v.y + z.x * v.z, x.y * v.x + y.y * v.y + z.y * v.z, x.z * v.x + y.z * v.y + z.z * v.z);
    }
};

// column-major 4x4 (Matrix4x4.h): element (r, c) at m[4 * c + r]
__device__ __forceinline__ V3 xfmPoint(const float* __restrict__ m, const V3& p) {
    float tx = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    float ty = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    float tz = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    const float tw = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
    if (tw != 1.0f) { const float r = 1.0f / tw; tx *= r; ty *= r; tz *= r; }
    return V3(tx, ty, tz);
}
__device__ __forceinline__ V3 xfmVector(const float* __restrict__ m, const V3& v) {
    return V3(m[0] * v.x + m[4] * v.y + m[8] * v.z, m[1] * v.x + m[5] * v.y + m[9] * v.z, m[2] * v.x + m[6] * v.y + m[10] * v.z);
}
// StaticTransform * Normal3D: transposed inverse (Transform.h:47-52); takes the INVERSE matrix
__device__ __forceinline__ V3 xfmNormal(const float* __restrict__ mi, const V3& n) {
    return V3(mi[0] * n.x + mi[1] * n.y + mi[2] * n.z, mi[4] * n.x + mi[5] * n.y + mi[6] * n.z, mi[8] * n.x + mi[9] * n.y + mi[10] * n.z);
}

// "A Low Distortion Map Between Disk and Square" (distributions.cpp:35-70)
__device__ __forceinline__ void concentricSampleDisk(float u0, float u1, float* dx, float* dy) {
    const float sx = 2 * u0 - 1, sy = 2 * u1 - 1;
    if (sx == 0 && sy == 0) { *dx = 0; *dy = 0; return; }
    float r, theta;
    if (sx >= -sy) {
        if (sx > sy) { r = sx; theta = sy / sx; }
        else { r = sy; theta = 2 - sx / sy; }
    } else {
        if (sx > sy) { r = -sy; theta = 6 + sx / sy; }
        else { r = -sx; theta = 4 + sy / sx; }
    }
    theta *= 0.25f * kPi;
    float s, c;
    sincosf(theta, &s, &c);
    *dx = r * c; *dy = r * s;
}
__device__ __forceinline__ V3 cosineSampleHemisphere(float u0, float u1) {
    float x, y;
    concentricSampleDisk(u0, u1, &x, &y);
    return V3(x, y, sqrtf(fmaxf(0.0f, 1.0f - x * x - y * y)));
}
__device__ __forceinline__ V3 uniformSampleCone(float u0, float u1, float cosThetaMax) {
    const float phi = 2 * kPi * u1;
    const float theta = acosf(1 - (1 - cosThetaMax) * u0);
    float sp, cp, st, ct;
    sincosf(phi, &sp, &cp);
    sincosf(theta, &st, &ct);
    return V3(cp * st, sp * st, ct);
}

}  // namespace slrgpu
