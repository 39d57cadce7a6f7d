// Where the ray stages of the wavefront read their rays and leave their results: the path queue -> hit
// buffer for `extend`, the shadow queue -> sensor splat for `shadow`. Shared by the wave kernels of
// trace.cu and the persistent tail kernel of tail.cu (both run the walk of traverse.cuh).
#pragma once
#include "traverse.cuh"
#include "wavefront.cuh"

namespace slrgpu {

// per-warp totals of the traversal counters -> one atomic pair per warp
__device__ __forceinline__ void addTraversalCounts(const TraversalCounters& cnt, unsigned long long* nodes, unsigned long long* leafRecords) {
    uint32_t n = cnt.nodes, t = cnt.tris;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xFFFFFFFFu, n, o); t += __shfl_xor_sync(0xFFFFFFFFu, t, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(nodes, (unsigned long long)n); atomicAdd(leafRecords, (unsigned long long)t); }
}

struct PathRaySource {
    PathQueue q;
    __device__ __forceinline__ void load(uint32_t i, Ray& r) const {
        const float4 o = q.org[i], d = q.dir[i];
        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.tmin = o.w;
        r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = INFINITY;
    }
    __device__ __forceinline__ float time(uint32_t i) const { return q.time ? q.time[i] : 0.0f; }
};
struct HitSink {
    HitBuffer hits;
    __device__ __forceinline__ void done(uint32_t i, const WalkState& w, const TraversalCounters&) const {
        hits.id[i] = make_uint2(w.hit.prim, w.hit.inst);
        hits.tuv[i] = make_float4(w.hit.t, w.hit.u, w.hit.v, __uint_as_float(w.hit.info));
    }
};

struct ShadowRaySource {
    ShadowQueue q;
    __device__ __forceinline__ void load(uint32_t i, Ray& r) const {
        const float4 o = q.org[i], d = q.dir[i];
        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.tmin = o.w;
        r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = d.w;
    }
    __device__ __forceinline__ float time(uint32_t i) const { return q.time ? q.time[i] : 0.0f; }
};
template <int NC> struct SplatSink {
    ShadowQueue q;
    float* accum;
    __device__ __forceinline__ void done(uint32_t i, const WalkState& w, const TraversalCounters&) const {
        if (w.found) return;         // occluded
        const uint2 pw = q.pixelWl[i];
        float v[NC == 3 ? 4 : NC];
        constexpr int Q = (NC + 3) / 4;
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const float4 c = q.contrib[(size_t)k * q.capacity + i];
            v[4 * k] = c.x;
            if (4 * k + 1 < (NC == 3 ? 4 : NC)) v[4 * k + 1] = c.y;
            if (4 * k + 2 < (NC == 3 ? 4 : NC)) v[4 * k + 2] = c.z;
            if (4 * k + 3 < (NC == 3 ? 4 : NC)) v[4 * k + 3] = c.w;
        }
        splat<NC>(accum, pw.x & 0x7FFFFFFFu, __uint_as_float(pw.y), (pw.x >> 31) != 0, v);
    }
};

}  // namespace slrgpu
