// The tail of a frame: one persistent kernel that carries the last long paths to their end.
//
// Why: in the wavefront loop every bounce of every path costs one wave = ~15 kernel launches. Once the camera
// samples of a render call are used up, the number of paths in flight falls geometrically (Russian roulette), but
// the few paths that keep bouncing between the mirror and the glass sphere live until the length cap of 100
// (PathTracingRenderer.cpp:162): ncu launch list of Cornell_Box_Spheres, profiles/r01_ncu_launches_v5.csv -- 104
// waves per frame, the last ~90 of them with a few thousand paths each and 0.25-0.3 ms of launch floors and
// cold-cache latency per wave, together about 40 % of the frame for well under 1 % of its rays.
//
// How: when `generated == total` and at most one path per thread of this kernel is left, every thread takes the
// path in its queue slot and runs bounce after bounce -- walk (the same walkStep as the extend kernel), the
// `surface` item, the `material` item of the hit's class, the shadow walk and splat -- reading and writing the
// path state at its own slot of the two path queues. No queue compaction, no atomics, no grid-wide barrier: a
// bounce costs the latency of its own loads, the BVH nodes of the spheres stay in L1. The arithmetic is the
// stages' own (stages.cuh, traverse.cuh: explicit round-to-nearest intrinsics), and a path's random numbers are
// keyed by (pixel, sample, bounce), so which kernel runs a bounce does not change the path.
//
// The kernel sits in the wave graph after every wave pair and returns at once while its condition is false
// (~3 us per pair); tailEndKernel then closes the loop state like endWaveKernel does.
#include "ray_io.cuh"
#include "stages.cuh"

namespace slrgpu {

constexpr int kTailBlock = 256;

__device__ __forceinline__ bool tailCondition(const WavefrontCounters* c, uint32_t cap) {
    const uint32_t n = c->numPaths;
    return n != 0 && n <= cap && c->generated == c->total;
}

// one ray, start to end, by its own lane (INSTANCES variant of the step: it also handles flat scenes)
template <bool ANY_HIT>
__device__ __noinline__ void tailWalk(const DeviceScene& s, WalkState& w, uint32_t* stack, bool* overflow) {
    InstanceWalkState iw;
    iw.leaves.clear(); iw.saved.clear(); iw.curInst = SLRGPU_INVALID_ID;
    TraversalCounters cnt = {0, 0};
    bool ovf = false;
    walkBegin(w, stack);
    while (!walkStep<true, ANY_HIT, false>(s, w, iw, stack, cnt, ovf)) { }
    if (ovf) *overflow = true;
}

template <int NC, int CLASS>
__device__ __noinline__ void tailMaterial(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                          const PathQueue& out, const ShadowQueue& sq, uint32_t slot, uint32_t leaf, bool* alive, bool* shadow) {
    MaterialResult<NC> o;
    o.clear();
    materialItem<NC, CLASS>(s, rc, in, hits, slot, leaf, o);
    materialWrite<NC>(out, sq, slot, slot, o);
    *alive = o.alive; *shadow = o.shadow;
}

__device__ __forceinline__ uint32_t warpSum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

template <int NC>
__global__ void __launch_bounds__(kTailBlock, 1)
tailKernel(const DeviceScene s, const RenderConstants rc, PathQueue q0, PathQueue q1, HitBuffer hits, ShadowQueue sq,
           float* __restrict__ accum, WavefrontCounters* counters, uint32_t cap, uint32_t classMask) {
    if (!tailCondition(counters, cap)) return;
    const uint32_t n = counters->numPaths;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = slot < n;
    if (!__any_sync(0xFFFFFFFFu, alive)) return;

    uint32_t stack[kStackSize];
    uint32_t nExtend = 0, nShadow = 0, waves = 0;
    uint32_t classHits[SC_COUNT];
#pragma unroll
    for (int c = 0; c < (int)SC_COUNT; ++c) classHits[c] = 0;
    bool overflow = false;
    const TraversalCounters noCount = {0, 0};
    uint32_t cur = 0;

    while (__any_sync(0xFFFFFFFFu, alive)) {
        const PathQueue in = cur ? q1 : q0;
        const PathQueue out = cur ? q0 : q1;
        uint32_t cls = SC_NONE, leaf = SLRGPU_INVALID_ID;
        if (alive) {
            ++nExtend;
            WalkState w;
            PathRaySource{in}.load(slot, w.r);
            tailWalk<false>(s, w, stack, &overflow);
            HitSink{hits}.done(slot, w, noCount);
            surfaceItem<NC>(s, rc, in, hits, accum, slot, &cls, &leaf);
        }
        bool next = false, shadow = false;
#define SLR_TAIL_CLASS(C)                                                                                   \
        if ((classMask >> (C)) & 1u) {                                                                      \
            if (cls == (C)) { ++classHits[C]; tailMaterial<NC, C>(s, rc, in, hits, out, sq, slot, leaf, &next, &shadow); } \
        }
        SLR_TAIL_CLASS(SC_LAMBERT) SLR_TAIL_CLASS(SC_OREN_NAYAR) SLR_TAIL_CLASS(SC_SPECULAR_BRDF) SLR_TAIL_CLASS(SC_SPECULAR_BSDF)
        SLR_TAIL_CLASS(SC_WARD) SLR_TAIL_CLASS(SC_ASHIKHMIN) SLR_TAIL_CLASS(SC_MF_BRDF) SLR_TAIL_CLASS(SC_MF_BSDF) SLR_TAIL_CLASS(SC_GENERIC)
#undef SLR_TAIL_CLASS
        if (shadow) {
            ++nShadow;
            WalkState w;
            ShadowRaySource{sq}.load(slot, w.r);
            tailWalk<true>(s, w, stack, &overflow);
            SplatSink<NC>{sq, accum}.done(slot, w, noCount);
        }
        alive = next;
        cur ^= 1u;
        ++waves;
    }

    // the loop state's totals: one atomic per warp and counter
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t e = warpSum(nExtend), sh = warpSum(nShadow), started = warpSum(slot < n ? 1u : 0u);
    if (lane == 0) {
        atomicAdd(&counters->extendRays, (unsigned long long)e);
        atomicAdd(&counters->shadowRays, (unsigned long long)sh);
        atomicAdd(&counters->tailPaths, started);
        atomicMax(&counters->tailWaves, waves);
    }
#pragma unroll
    for (int c = 0; c < (int)SC_COUNT; ++c) {
        const uint32_t h = warpSum(classHits[c]);
        if (lane == 0 && h) atomicAdd(&counters->classTotal[c], (unsigned long long)h);
    }
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
}

// after the tail: nothing is in flight any more (single thread); rewrites the snapshot endWaveKernel left for the host
__global__ void tailEndKernel(WavefrontCounters* counters, WavefrontCounters* ring, uint32_t ringSize, uint32_t cap) {
    if (!tailCondition(counters, cap)) return;
    counters->numPaths = 0;
    counters->done = 1u;
    ring[(counters->waves - 1u) % ringSize] = *counters;
    __threadfence_system();
}

uint32_t tailCapacity(int numSMs) { return (uint32_t)numSMs * (uint32_t)kTailBlock; }

int launchTail(const SlrGpuScene* sc, const RenderConstants& rc, const PathQueue& q0, const PathQueue& q1, const HitBuffer& hits,
               const ShadowQueue& sq, float* accum, WavefrontCounters* counters, WavefrontCounters* ring, uint32_t ringSize,
               uint32_t cap, cudaStream_t stream) {
    if (cap == 0) return SLRGPU_OK;
    const uint32_t grid = (cap + kTailBlock - 1) / kTailBlock;
    if (sc->channels == 3) tailKernel<3><<<grid, kTailBlock, 0, stream>>>(sc->dev, rc, q0, q1, hits, sq, accum, counters, cap, sc->classMask);
    else tailKernel<16><<<grid, kTailBlock, 0, stream>>>(sc->dev, rc, q0, q1, hits, sq, accum, counters, cap, sc->classMask);
    tailEndKernel<<<1, 1, 0, stream>>>(counters, ring, ringSize, cap);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

}  // namespace slrgpu
