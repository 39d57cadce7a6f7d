// Motion blur on the device: AnimatedTransform::sample (libSLR/Core/Transform.h:105-122) for an instance or the camera
// at a ray's time -- the key frame outside the key times, else translate(lerp T) * Slerp(R).toMatrix() * lerp(S) of the
// host's decomposition (Quaternion.h:86-127) -- and the inverse the reference takes with a general 4x4 Gauss-Jordan
// (Transform.cpp:31-35 invert(tf)); the composed matrix is affine, so the inverse here is the 3x3 adjugate form (agrees
// with the reference's to fp32 rounding; image parity for moving geometry is statistical, hit parity is defined on
// static scenes).
#pragma once
#include "device_scene.h"

namespace slrgpu {

// mat / matInv (column-major, 16 floats each) of the owner at `time`; begin = the owner's own mat / mat_inv
__device__ inline void sampleMotion(const SlrGpuMotion& m, const float* __restrict__ beginMat, const float* __restrict__ beginInv,
                                    float time, float* __restrict__ mat, float* __restrict__ matInv) {
    if (time <= m.t_begin) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { mat[i] = beginMat[i]; matInv[i] = beginInv[i]; }
        return;
    }
    if (time >= m.t_end) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { mat[i] = m.mat_end[i]; matInv[i] = m.mat_end_inv[i]; }
        return;
    }
    const float t = (time - m.t_begin) / (m.t_end - m.t_begin);
    const float a = 1.0f - t;
    const float tx = a * m.T0[0] + t * m.T1[0], ty = a * m.T0[1] + t * m.T1[1], tz = a * m.T0[2] + t * m.T1[2];
    // Slerp(t, R0, R1)
    float qx, qy, qz, qw;
    {
        const float cosTheta = (m.R0[0] * m.R1[0] + m.R0[1] * m.R1[1] + m.R0[2] * m.R1[2]) + m.R0[3] * m.R1[3];
        if (cosTheta > 0.9995f) {
            qx = a * m.R0[0] + t * m.R1[0]; qy = a * m.R0[1] + t * m.R1[1]; qz = a * m.R0[2] + t * m.R1[2]; qw = a * m.R0[3] + t * m.R1[3];
            const float r = 1.0f / sqrtf((qx * qx + qy * qy + qz * qz) + qw * qw);
            qx *= r; qy *= r; qz *= r; qw *= r;
        } else {
            const float theta = acosf(fminf(fmaxf(cosTheta, -1.0f), 1.0f));
            float px = m.R1[0] - m.R0[0] * cosTheta, py = m.R1[1] - m.R0[1] * cosTheta, pz = m.R1[2] - m.R0[2] * cosTheta, pw = m.R1[3] - m.R0[3] * cosTheta;
            const float r = 1.0f / sqrtf((px * px + py * py + pz * pz) + pw * pw);
            px *= r; py *= r; pz *= r; pw *= r;
            float sn, cs;
            sincosf(theta * t, &sn, &cs);
            qx = m.R0[0] * cs + px * sn; qy = m.R0[1] * cs + py * sn; qz = m.R0[2] * cs + pz * sn; qw = m.R0[3] * cs + pw * sn;
        }
    }
    // R = q.toMatrix() (columns), S = lerp(S0, S1): upper 3x3 of both (their last row / column is 0 0 0 1)
    const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, yz = qy * qz, zx = qz * qx, xw = qx * qw, yw = qy * qw, zw = qz * qw;
    const float R[9] = {1 - 2 * (yy + zz), 2 * (xy + zw), 2 * (zx - yw),          // column 0
                        2 * (xy - zw), 1 - 2 * (xx + zz), 2 * (yz + xw),          // column 1
                        2 * (zx + yw), 2 * (yz - xw), 1 - 2 * (xx + yy)};         // column 2
    float A[9];            // A = R * S, column-major 3x3
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float s0 = a * m.S0[4 * c] + t * m.S1[4 * c], s1 = a * m.S0[4 * c + 1] + t * m.S1[4 * c + 1], s2 = a * m.S0[4 * c + 2] + t * m.S1[4 * c + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) A[3 * c + r] = R[r] * s0 + R[3 + r] * s1 + R[6 + r] * s2;
    }
    mat[0] = A[0]; mat[1] = A[1]; mat[2] = A[2]; mat[3] = 0.0f;
    mat[4] = A[3]; mat[5] = A[4]; mat[6] = A[5]; mat[7] = 0.0f;
    mat[8] = A[6]; mat[9] = A[7]; mat[10] = A[8]; mat[11] = 0.0f;
    mat[12] = tx; mat[13] = ty; mat[14] = tz; mat[15] = 1.0f;
    // inverse of the affine matrix: B = A^-1 (adjugate / determinant), translation -B t
    const float c00 = A[4] * A[8] - A[7] * A[5], c01 = A[7] * A[2] - A[1] * A[8], c02 = A[1] * A[5] - A[4] * A[2];
    const float det = A[0] * c00 + A[3] * c01 + A[6] * c02;
    const float id = 1.0f / det;
    float B[9];
    B[0] = c00 * id; B[1] = c01 * id; B[2] = c02 * id;
    B[3] = (A[6] * A[5] - A[3] * A[8]) * id; B[4] = (A[0] * A[8] - A[6] * A[2]) * id; B[5] = (A[3] * A[2] - A[0] * A[5]) * id;
    B[6] = (A[3] * A[7] - A[6] * A[4]) * id; B[7] = (A[6] * A[1] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[3] * A[1]) * id;
    matInv[0] = B[0]; matInv[1] = B[1]; matInv[2] = B[2]; matInv[3] = 0.0f;
    matInv[4] = B[3]; matInv[5] = B[4]; matInv[6] = B[5]; matInv[7] = 0.0f;
    matInv[8] = B[6]; matInv[9] = B[7]; matInv[10] = B[8]; matInv[11] = 0.0f;
    matInv[12] = -(B[0] * tx + B[3] * ty + B[6] * tz);
    matInv[13] = -(B[1] * tx + B[4] * ty + B[7] * tz);
    matInv[14] = -(B[2] * tx + B[5] * ty + B[8] * tz);
    matInv[15] = 1.0f;
}

// An instance's transform at `time`: its own matrices when it does not move, else sampled into `scratch` (32 floats)
struct InstanceXfm { const float* mat; const float* matInv; };
__device__ __forceinline__ InstanceXfm instanceTransformAt(const DeviceScene& s, const SlrGpuInstance& in, float time, float* scratch) {
    if (in.motion == 0u || s.motions == nullptr) return InstanceXfm{in.mat, in.mat_inv};
    sampleMotion(s.motions[in.motion - 1u], in.mat, in.mat_inv, time, scratch, scratch + 16);
    return InstanceXfm{scratch, scratch + 16};
}

}  // namespace slrgpu
