// Wavefront ray stages: `extend` (closest hit for every path in the queue) and `shadow` (occlusion
// of the next-event-estimation rays, splatting the unoccluded contributions to the sensor).
//   Scene::intersect        libSLR/Core/SurfaceObject.cpp:408-416  -> QBVH traversal of traverse.cuh
//   Scene::testVisibility   libSLR/Core/SurfaceObject.cpp:418-430  (the reference runs a full closest-hit
//                           query and keeps only the boolean; any-hit with early exit gives the same boolean)
// MUST be compiled with -fmad=false like intersect.cu: the renderer's rays use the bit-exact traversal.
#ifndef SLR_WALK_DEFER_SINK
#define SLR_WALK_DEFER_SINK 2      // results handed to the sink at the refill point: instanced instantiations only (traverse.cuh)
#endif
#include "ray_io.cuh"

#ifndef SLR_TRACE_MIN_BLOCKS
#define SLR_TRACE_MIN_BLOCKS 1
#endif
// the instanced walk sits at the 80-register boundary between 6 and 5 resident blocks per SM (registers are allocated
// 8 at a time per thread): bounded for 6
#ifndef SLR_TRACE_MIN_BLOCKS_INSTANCED
#define SLR_TRACE_MIN_BLOCKS_INSTANCED 6
#endif

namespace slrgpu {


// ALPHA: the scene has alpha-mapped (cut-out) triangles (TriangleMesh.cpp:160-168); such scenes run the general
// instantiation <INSTANCES = true, ALPHA = true>, which handles flat scenes too.
template <bool INSTANCES, bool COUNT, bool ALPHA>
__global__ void __launch_bounds__(kTraceBlock, (INSTANCES && !ALPHA) ? SLR_TRACE_MIN_BLOCKS_INSTANCED : SLR_TRACE_MIN_BLOCKS)
extendKernel(const DeviceScene s, PathQueue q, HitBuffer hits, WavefrontCounters* counters) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, false, COUNT, ALPHA>(s, counters->numPaths, &counters->extendCursor, PathRaySource{q}, HitSink{hits}, cnt, overflow);
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
    if (COUNT) addTraversalCounts(cnt, &counters->extendNodes, &counters->extendLeafRecords);
}

template <bool INSTANCES, int NC, bool COUNT, bool ALPHA>
__global__ void __launch_bounds__(kTraceBlock, (INSTANCES && !ALPHA) ? SLR_TRACE_MIN_BLOCKS_INSTANCED : SLR_TRACE_MIN_BLOCKS)
shadowKernel(const DeviceScene s, ShadowQueue q, float* __restrict__ accum, WavefrontCounters* counters) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, true, COUNT, ALPHA>(s, counters->numShadow, &counters->shadowCursor, ShadowRaySource{q}, SplatSink<NC>{q, accum}, cnt, overflow);
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
    if (COUNT) addTraversalCounts(cnt, &counters->shadowNodes, &counters->shadowLeafRecords);
}

template <bool INSTANCES, bool COUNT, bool ALPHA>
static void launchExtendT(const SlrGpuScene* sc, const PathQueue& q, const HitBuffer& hits, WavefrontCounters* counters, uint32_t grid, cudaStream_t stream) {
    extendKernel<INSTANCES, COUNT, ALPHA><<<residentGrid(extendKernel<INSTANCES, COUNT, ALPHA>, kTraceBlock, sc->numSMs, grid), kTraceBlock, 0, stream>>>(sc->dev, q, hits, counters);
}
template <bool INSTANCES, bool COUNT, bool ALPHA>
static void launchShadowT(const SlrGpuScene* sc, const ShadowQueue& q, float* accum, WavefrontCounters* counters, uint32_t grid, cudaStream_t stream) {
    if (sc->channels == 3) shadowKernel<INSTANCES, 3, COUNT, ALPHA><<<residentGrid(shadowKernel<INSTANCES, 3, COUNT, ALPHA>, kTraceBlock, sc->numSMs, grid), kTraceBlock, 0, stream>>>(sc->dev, q, accum, counters);
    else shadowKernel<INSTANCES, 16, COUNT, ALPHA><<<residentGrid(shadowKernel<INSTANCES, 16, COUNT, ALPHA>, kTraceBlock, sc->numSMs, grid), kTraceBlock, 0, stream>>>(sc->dev, q, accum, counters);
}

// variant of a scene: 0 flat, 1 instanced, 2 general (instances and / or alpha-mapped triangles)
static int walkVariant(const SlrGpuScene* sc) { return sc->hasAlpha ? 2 : sc->hasInstances ? 1 : 0; }

int launchExtend(const SlrGpuScene* sc, const PathQueue& q, const HitBuffer& hits, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream) {
    switch (walkVariant(sc) * 2 + (count ? 1 : 0)) {
    case 0: launchExtendT<false, false, false>(sc, q, hits, counters, grid, stream); break;
    case 1: launchExtendT<false, true, false>(sc, q, hits, counters, grid, stream); break;
    case 2: launchExtendT<true, false, false>(sc, q, hits, counters, grid, stream); break;
    case 3: launchExtendT<true, true, false>(sc, q, hits, counters, grid, stream); break;
    case 4: launchExtendT<true, false, true>(sc, q, hits, counters, grid, stream); break;
    default: launchExtendT<true, true, true>(sc, q, hits, counters, grid, stream); break;
    }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

int launchShadow(const SlrGpuScene* sc, const ShadowQueue& q, float* accum, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream) {
    switch (walkVariant(sc) * 2 + (count ? 1 : 0)) {
    case 0: launchShadowT<false, false, false>(sc, q, accum, counters, grid, stream); break;
    case 1: launchShadowT<false, true, false>(sc, q, accum, counters, grid, stream); break;
    case 2: launchShadowT<true, false, false>(sc, q, accum, counters, grid, stream); break;
    case 3: launchShadowT<true, true, false>(sc, q, accum, counters, grid, stream); break;
    case 4: launchShadowT<true, false, true>(sc, q, accum, counters, grid, stream); break;
    default: launchShadowT<true, true, true>(sc, q, accum, counters, grid, stream); break;
    }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

}  // namespace slrgpu
