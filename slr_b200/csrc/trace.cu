// Wavefront ray stages: `extend` (closest hit for every path in the queue) and `shadow` (occlusion
// of the next-event-estimation rays, splatting the unoccluded contributions to the sensor).
//   Scene::intersect        libSLR/Core/SurfaceObject.cpp:408-416  -> QBVH traversal of traverse.cuh
//   Scene::testVisibility   libSLR/Core/SurfaceObject.cpp:418-430  (the reference runs a full closest-hit
//                           query and keeps only the boolean; any-hit with early exit gives the same boolean)
// MUST be compiled with -fmad=false like intersect.cu: the renderer's rays use the bit-exact traversal.
#include "traverse.cuh"
#include "wavefront.cuh"

#ifndef SLR_TRACE_MIN_BLOCKS
#define SLR_TRACE_MIN_BLOCKS 1
#endif

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
};
struct HitSink {
    HitBuffer hits;
    __device__ __forceinline__ void done(uint32_t i, const WalkState& w, const TraversalCounters&) const {
        hits.id[i] = make_uint2(w.hit.prim, w.hit.inst);
        hits.tuv[i] = make_float4(w.hit.t, w.hit.u, w.hit.v, 0.0f);
    }
};

template <bool INSTANCES, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, SLR_TRACE_MIN_BLOCKS)
extendKernel(const DeviceScene s, PathQueue q, HitBuffer hits, WavefrontCounters* counters) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, false, COUNT>(s, counters->numPaths, &counters->extendCursor, PathRaySource{q}, HitSink{hits}, cnt, overflow);
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
    if (COUNT) addTraversalCounts(cnt, &counters->extendNodes, &counters->extendLeafRecords);
}

struct ShadowRaySource {
    ShadowQueue q;
    __device__ __forceinline__ void load(uint32_t i, Ray& r) const {
        const float4 o = q.org[i], d = q.dir[i];
        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.tmin = o.w;
        r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = d.w;
    }
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

template <bool INSTANCES, int NC, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, SLR_TRACE_MIN_BLOCKS)
shadowKernel(const DeviceScene s, ShadowQueue q, float* __restrict__ accum, WavefrontCounters* counters) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, true, COUNT>(s, counters->numShadow, &counters->shadowCursor, ShadowRaySource{q}, SplatSink<NC>{q, accum}, cnt, overflow);
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
    if (COUNT) addTraversalCounts(cnt, &counters->shadowNodes, &counters->shadowLeafRecords);
}

template <bool INSTANCES, bool COUNT>
static void launchExtendT(const SlrGpuScene* sc, const PathQueue& q, const HitBuffer& hits, WavefrontCounters* counters, uint32_t grid, cudaStream_t stream) {
    extendKernel<INSTANCES, COUNT><<<grid, kTraceBlock, 0, stream>>>(sc->dev, q, hits, counters);
}
template <bool INSTANCES, bool COUNT>
static void launchShadowT(const SlrGpuScene* sc, const ShadowQueue& q, float* accum, WavefrontCounters* counters, uint32_t grid, cudaStream_t stream) {
    if (sc->channels == 3) shadowKernel<INSTANCES, 3, COUNT><<<grid, kTraceBlock, 0, stream>>>(sc->dev, q, accum, counters);
    else shadowKernel<INSTANCES, 16, COUNT><<<grid, kTraceBlock, 0, stream>>>(sc->dev, q, accum, counters);
}

int launchExtend(const SlrGpuScene* sc, const PathQueue& q, const HitBuffer& hits, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream) {
    if (sc->hasInstances) { if (count) launchExtendT<true, true>(sc, q, hits, counters, grid, stream); else launchExtendT<true, false>(sc, q, hits, counters, grid, stream); }
    else { if (count) launchExtendT<false, true>(sc, q, hits, counters, grid, stream); else launchExtendT<false, false>(sc, q, hits, counters, grid, stream); }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

int launchShadow(const SlrGpuScene* sc, const ShadowQueue& q, float* accum, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream) {
    if (sc->hasInstances) { if (count) launchShadowT<true, true>(sc, q, accum, counters, grid, stream); else launchShadowT<true, false>(sc, q, accum, counters, grid, stream); }
    else { if (count) launchShadowT<false, true>(sc, q, accum, counters, grid, stream); else launchShadowT<false, false>(sc, q, accum, counters, grid, stream); }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

}  // namespace slrgpu
