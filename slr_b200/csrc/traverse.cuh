// Software ray traversal of the 4-wide QBVH and Moller-Trumbore triangle test for sm_100a.
//
// B200 has no RT cores; this is the hand-written replacement for
//   QBVH::Node::intersect   libSLR/Accelerator/QBVH.h:55-76   (SSE 4-box slab test)
//   QBVH::intersect         libSLR/Accelerator/QBVH.h:295-337 (ordered stack traversal)
//   Triangle::intersect     libSLR/Surface/TriangleMesh.cpp:131-178
//   TransformedSurfaceObject::intersect  libSLR/Core/SurfaceObject.cpp:307-318
//
// PARITY CONTRACT: hit primitive/instance ids must equal the reference's bit for bit, including
// which primitive wins when two report the same distance. That pins (1) the arithmetic: plain IEEE
// fp32, no FMA contraction (this header must be compiled with -fmad=false), division as 1.0f/x then
// multiply; (2) the visiting order: children ordered by the sign of the ray direction along the
// node's three split axes, inner children pushed in reverse, leaf children tested immediately in
// order, `t > tmax` rejects so a later primitive at an equal distance replaces an earlier one.
//
// Data movement: a node is 128 B = eight 16-byte loads through the read-only path (LDG.E.128);
// a leaf record is 48 B = three. One thread owns one ray; the 64-entry stack lives in local memory.
#pragma once
#include "device_scene.h"

namespace slrgpu {

constexpr int kStackSize = 64;            // QBVH.h:299
constexpr uint32_t kEmptyChild = 0xFFFFFFFFu;

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float tmin, tmax;
};

struct Hit {
    uint32_t prim;      // SLRGPU_INVALID_ID = miss
    uint32_t inst;
    float t, u, v;
};

struct TraversalCounters {
    uint32_t nodes, tris;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 4-lane slab test; returns the 4-bit mask of lanes whose [tNear, tFar] is non-empty.
// fmaxf/fminf return the non-NaN operand, which coincides with _mm_max_ps/_mm_min_ps here because a
// NaN can only appear in the freshly computed first operand (0 * inf), never in the running bound.
__device__ __forceinline__ uint32_t slab4(const float4 lox, const float4 loy, const float4 loz,
                                          const float4 hix, const float4 hiy, const float4 hiz,
                                          const Ray& r, float ix, float iy, float iz) {
    const bool px = ix > 0.0f, py = iy > 0.0f, pz = iz > 0.0f;
    const float4 nx = px ? lox : hix, fx = px ? hix : lox;
    const float4 ny = py ? loy : hiy, fy = py ? hiy : loy;
    const float4 nz = pz ? loz : hiz, fz = pz ? hiz : loz;
    uint32_t mask = 0;
#define SLAB_LANE(L, bit)                                                   \
    {                                                                       \
        float tn = r.tmin, tf = r.tmax;                                     \
        tn = fmaxf((nx.L - r.ox) * ix, tn);                                 \
        tn = fmaxf((ny.L - r.oy) * iy, tn);                                 \
        tn = fmaxf((nz.L - r.oz) * iz, tn);                                 \
        tf = fminf((fx.L - r.ox) * ix, tf);                                 \
        tf = fminf((fy.L - r.oy) * iy, tf);                                 \
        tf = fminf((fz.L - r.oz) * iz, tf);                                 \
        if (tn <= tf) mask |= bit;                                          \
    }
    SLAB_LANE(x, 1u) SLAB_LANE(y, 2u) SLAB_LANE(z, 4u) SLAB_LANE(w, 8u)
#undef SLAB_LANE
    return mask;
}

// Moller-Trumbore, two-sided, exactly the reference's sequence of operations and comparisons
// (NaNs fall through the range checks the same way because the comparisons are not negated).
__device__ __forceinline__ bool triangleTest(const float4 a, const float4 b, const float4 c, const Ray& r,
                                             float* tOut, float* b0Out, float* b1Out) {
    const float e1x = b.x, e1y = b.y, e1z = b.z;
    const float e2x = c.x, e2y = c.y, e2z = c.z;
    const float px = r.dy * e2z - r.dz * e2y;
    const float py = r.dz * e2x - r.dx * e2z;
    const float pz = r.dx * e2y - r.dy * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.0f) return false;
    const float invDet = 1.0f / det;
    const float dx = r.ox - a.x, dy = r.oy - a.y, dz = r.oz - a.z;
    const float b1 = (dx * px + dy * py + dz * pz) * invDet;
    if (b1 < 0.0f || b1 > 1.0f) return false;
    const float qx = dy * e1z - dz * e1y;
    const float qy = dz * e1x - dx * e1z;
    const float qz = dx * e1y - dy * e1x;
    const float b2 = (r.dx * qx + r.dy * qy + r.dz * qz) * invDet;
    if (b2 < 0.0f || b1 + b2 > 1.0f) return false;
    const float tt = (e2x * qx + e2y * qy + e2z * qz) * invDet;
    if (tt < r.tmin || tt > r.tmax) return false;
    *tOut = tt;
    *b0Out = 1.0f - b1 - b2;
    *b1Out = b1;
    return true;
}

// Matrix4x4 * Point3 with the reference's homogeneous divide rule (Matrix4x4.h:75-81); column-major m.
__device__ __forceinline__ void mulPoint(const float* __restrict__ m, float x, float y, float z,
                                         float* ox, float* oy, float* oz) {
    float tx = m[0] * x + m[4] * y + m[8] * z + m[12] * 1.0f;
    float ty = m[1] * x + m[5] * y + m[9] * z + m[13] * 1.0f;
    float tz = m[2] * x + m[6] * y + m[10] * z + m[14] * 1.0f;
    float tw = m[3] * x + m[7] * y + m[11] * z + m[15] * 1.0f;
    if (tw != 1.0f) { float rcp = 1.0f / tw; tx *= rcp; ty *= rcp; tz *= rcp; }
    *ox = tx; *oy = ty; *oz = tz;
}
__device__ __forceinline__ void mulVector(const float* __restrict__ m, float x, float y, float z,
                                          float* ox, float* oy, float* oz) {
    *ox = m[0] * x + m[4] * y + m[8] * z;
    *oy = m[1] * x + m[5] * y + m[9] * z;
    *oz = m[2] * x + m[6] * y + m[10] * z;
}

// Traverses the BVH rooted at `root` for ray `r` (r.tmax shrinks on accepted hits). `sp` is the
// first free stack slot: a nested (instance) traversal runs on the same stack above its caller's
// entries and returns when it has popped back down to its own base, which reproduces the
// reference's recursion (the nested traversal completes before the caller's next leaf record).
// LEVEL bounds the instancing depth at compile time (0 = top level).
template <int LEVEL, bool ANY_HIT, bool COUNT>
__device__ bool traverse(const DeviceScene& s, uint32_t root, Ray& r, Hit& hit, uint32_t* stack, int sp,
                         TraversalCounters& cnt, bool& overflow) {
    const float ix = 1.0f / r.dx, iy = 1.0f / r.dy, iz = 1.0f / r.dz;
    const uint32_t pos = (r.dx >= 0.0f ? 1u : 0u) | (r.dy >= 0.0f ? 2u : 0u) | (r.dz >= 0.0f ? 4u : 0u);
    bool found = false;
    const int base = sp;
    stack[sp++] = root;
    while (sp > base) {
        const uint32_t nodeIdx = stack[--sp];
        const float4* n = s.nodes + (size_t)nodeIdx * 8;
        const float4 lox = ldg4(n + 0), loy = ldg4(n + 1), loz = ldg4(n + 2);
        const float4 hix = ldg4(n + 3), hiy = ldg4(n + 4), hiz = ldg4(n + 5);
        if (COUNT) ++cnt.nodes;
        const uint32_t mask = slab4(lox, loy, loz, hix, hiy, hiz, r, ix, iy, iz);
        if (mask == 0) continue;
        const uint4 kids = __ldg(reinterpret_cast<const uint4*>(n + 6));
        const uint32_t axes = __ldg(reinterpret_cast<const uint32_t*>(n + 7));
        const uint32_t T = (pos >> (axes & 0xFF)) & 1u;
        const uint32_t L = (pos >> ((axes >> 8) & 0xFF)) & 1u;
        const uint32_t R = (pos >> ((axes >> 16) & 0xFF)) & 1u;
        // visiting order (OrderTable, QBVH.h:309-312): near side pair first, near child first inside a pair
        const uint32_t l0 = L ? 0u : 1u, r0 = R ? 2u : 3u;
        uint32_t order[4];
        order[0] = T ? l0 : r0;        order[1] = T ? (l0 ^ 1u) : (r0 ^ 1u);
        order[2] = T ? r0 : l0;        order[3] = T ? (r0 ^ 1u) : (l0 ^ 1u);
        uint32_t ch[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t lane = order[i];
            uint32_t c = lane == 0 ? kids.x : lane == 1 ? kids.y : lane == 2 ? kids.z : kids.w;
            ch[i] = ((mask >> lane) & 1u) ? c : kEmptyChild;
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            const uint32_t c = ch[i];
            if (c == kEmptyChild || (c >> 31)) continue;
            if (sp >= kStackSize) { overflow = true; continue; }
            stack[sp++] = c & 0x07FFFFFFu;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t c = ch[i];
            if (c == kEmptyChild || !(c >> 31)) continue;
            const uint32_t first = c & 0x07FFFFFFu;
            const uint32_t count = (c >> 27) & 0xFu;
            for (uint32_t j = 0; j < count; ++j) {
                const float4* rec = s.leaves + (size_t)(first + j) * 3;
                const float4 a = ldg4(rec);
                const uint32_t id = __float_as_uint(a.w);
                if (COUNT) ++cnt.tris;
                if (id & 0x80000000u) {
                    if constexpr (LEVEL < 1) {
                        const uint32_t instId = id & 0x7FFFFFFFu;
                        const SlrGpuInstance* inst = s.instances + instId;
                        Ray lr;
                        mulPoint(inst->mat_inv, r.ox, r.oy, r.oz, &lr.ox, &lr.oy, &lr.oz);
                        mulVector(inst->mat_inv, r.dx, r.dy, r.dz, &lr.dx, &lr.dy, &lr.dz);
                        lr.tmin = r.tmin; lr.tmax = r.tmax;
                        if (traverse<LEVEL + 1, ANY_HIT, COUNT>(s, inst->root_node, lr, hit, stack, sp, cnt, overflow)) {
                            r.tmax = lr.tmax;
                            hit.inst = instId;
                            found = true;
                            if (ANY_HIT) return true;
                        }
                    }
                    continue;
                }
                const float4 b = ldg4(rec + 1), cc = ldg4(rec + 2);
                float t, b0, b1;
                if (triangleTest(a, b, cc, r, &t, &b0, &b1)) {
                    r.tmax = t;
                    hit.prim = id; hit.inst = SLRGPU_INVALID_ID;
                    hit.t = t; hit.u = b0; hit.v = b1;
                    found = true;
                    if (ANY_HIT) return true;
                }
            }
        }
    }
    return found;
}

}  // namespace slrgpu
