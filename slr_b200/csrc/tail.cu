// The tail of a frame: one persistent kernel that carries the last long paths to their end.
//
// Why: in the wavefront loop every bounce of every path costs one wave = ~15 kernel launches. Once the camera
// samples of a render call are used up, the number of paths in flight falls geometrically (Russian roulette), but
// the few paths that keep bouncing between the mirror and the glass sphere live until the length cap of 100
// (PathTracingRenderer.cpp:162): ncu launch list of Cornell_Box_Spheres, profiles/r01_ncu_launches_v5.csv -- 104
// waves per frame, the last ~90 of them with a few thousand paths each and 0.25-0.3 ms of launch floors and
// cold-cache latency per wave, together about 40 % of the frame for well under 1 % of its rays.
//
// How: when `generated == total` and only a few paths per thread of this kernel are left, every lane takes a path
// from the queue the last wave left and runs it bounce after bounce -- walk (the same walkStep as the extend
// kernel), the `surface` item, the `material` item of the hit's class, the shadow walk and splat -- reading and
// writing the path state at the path's own entry of the two path queues; a lane whose path ended takes the next
// entry (one atomic per warp and round). No queue compaction, no grid-wide barrier: a bounce costs the latency of
// its own loads, the BVH nodes of the spheres stay in L1. The arithmetic is the
// stages' own (stages.cuh, traverse.cuh: explicit round-to-nearest intrinsics), and a path's random numbers are
// keyed by (pixel, sample, bounce), so which kernel runs a bounce does not change the path.
//
// The kernel sits in the wave graph after every wave pair and returns at once while its condition is false
// (~3 us per pair); tailEndKernel then closes the loop state like endWaveKernel does.
#include "ray_io.cuh"
#include "single_ray.cuh"
#include "stages.cuh"
#include <algorithm>

namespace slrgpu {

constexpr int kTailBlock = 256;
// L1 prefetch of pushed children in the tail's walk (traverse.cuh walkNode PREFETCH): measured in round 2 and left off --
// tail kernel 2.95 -> 3.23 ms per C1 frame (profiles/r02_rejected_experiments.md)
#ifndef SLR_TAIL_PREFETCH
#define SLR_TAIL_PREFETCH 0
#endif
// the handed-over paths dealt out as equal shares per warp instead of 32 at a time: measured in round 2 and left off -- C1
// 622.0 vs 620.0 Mpaths/s (nothing), C2 728 vs 744 (worse: its warps then hold paths of more material classes each) --
// profiles/r02_rejected_experiments.md
#ifndef SLR_TAIL_SPREAD
#define SLR_TAIL_SPREAD 0
#endif
#ifndef SLR_TAIL_PATHS_PER_THREAD
#define SLR_TAIL_PATHS_PER_THREAD 1u
#endif

__device__ __forceinline__ bool tailCondition(const WavefrontCounters* c, uint32_t cap) {
    const uint32_t n = c->numPaths;
    return n != 0 && n <= cap && c->generated == c->total;
}

// one ray, start to end, by its own lane (60 % of the tail kernel's instructions: the long paths live in and between the
// two sphere meshes, whose interior rays visit many nodes): single_ray.cuh
template <bool ANY_HIT>
__device__ __forceinline__ void tailWalk(const DeviceScene& s, WalkState& w, uint32_t* stack, bool* overflow) {   // w.time set by the caller
    singleRayWalk<ANY_HIT, SLR_TAIL_PREFETCH != 0>(s, w, stack, overflow);
}

template <int NC, int CLASS>
__device__ __noinline__ void tailMaterial(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                          const PathQueue& out, const ShadowQueue& sq, uint32_t slot, uint32_t leaf, bool* alive, bool* shadow) {
    MaterialResult<NC> o;
    o.clear();
    materialItem<NC, CLASS>(s, rc, in, hits, slot, leaf, o);
    if (o.shadow) materialWriteShadow<NC>(sq, slot, o);
    if (o.alive) materialWriteNext<NC>(out, slot, o);
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
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;

    uint32_t stack[kStackSize];
    uint32_t nExtend = 0, nShadow = 0, rounds = 0;
    uint32_t classHits[SC_COUNT];
#pragma unroll
    for (int c = 0; c < (int)SC_COUNT; ++c) classHits[c] = 0;
    bool overflow = false;
    const TraversalCounters noCount = {0, 0};
    // the lane's path: its entry of the path queue left by the last wave; the state of bounce b sits at that entry of
    // q0 (b even) or q1 (b odd), its hit and its shadow ray at the same entry of the hit buffer and the shadow queue
    uint32_t slot = 0, cur = 0;
    bool alive = false;
    bool exhausted = false;         // warp-uniform: the queue has no entries left to hand out
    // the warp's fair share of the handed-over paths (SLR_TAIL_SPREAD): with one cursor and 32 entries per grab the first
    // n / 32 warps would each run 32 paths in lockstep -- every round as long as its slowest lane's walk, the material
    // classes of all its lanes one after the other -- while the other warps of the grid idle; a round of a warp that holds
    // a third as many paths is that much shorter, and the kernel's time is the sum of one warp's rounds
    const uint32_t numWarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t quota = SLR_TAIL_SPREAD ? max(1u, (n + numWarps - 1u) / numWarps) : 0xFFFFFFFFu;    // no share: a warp refills while entries are left
    uint32_t taken = 0;             // warp-uniform: entries this warp has taken so far

    while (true) {
        // lanes without a path take the next entries of the queue (one atomic per warp and round)
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !alive);
        if (idle != 0 && !exhausted && taken < quota) {
            const uint32_t want = min((uint32_t)__popc(idle), quota - taken);
            taken += want;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&counters->tailCursor, want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (!alive) {
                const uint32_t rank = (uint32_t)__popc(idle & lt);
                const uint32_t i = base + rank;
                if (rank < want && i < n) { slot = i; cur = 0; alive = true; }
            }
            if (base + want >= n) exhausted = true;
        }
        if (!__any_sync(0xFFFFFFFFu, alive)) break;

        const PathQueue in = cur ? q1 : q0;
        const PathQueue out = cur ? q0 : q1;
        uint32_t cls = SC_NONE, leaf = SLRGPU_INVALID_ID;
        if (alive) {
            ++nExtend;
            WalkState w;
            PathRaySource{in}.load(slot, w.r);
            w.time = PathRaySource{in}.time(slot);
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
            w.time = ShadowRaySource{sq}.time(slot);
            tailWalk<true>(s, w, stack, &overflow);
            SplatSink<NC>{sq, accum}.done(slot, w, noCount);
        }
        alive = next;
        cur ^= 1u;
        ++rounds;
    }

    // the loop state's totals: one atomic per warp and counter
    const uint32_t e = warpSum(nExtend), sh = warpSum(nShadow);
    if (lane == 0) {
        atomicAdd(&counters->extendRays, (unsigned long long)e);
        atomicAdd(&counters->shadowRays, (unsigned long long)sh);
        atomicMax(&counters->tailWaves, rounds);
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
    counters->tailPaths = counters->numPaths;
    counters->numPaths = 0;
    counters->done = 1u;
    ring[(counters->waves - 1u) % ringSize] = *counters;
    __threadfence_system();
}

// default limit: one path per thread of the kernel. Sweep (profiles/r01_tail_kernel.md): C1 480 / 489 / 478 / 441 Mpaths/s at
// 1 / 2 / 4 / 16 paths per thread, C2 (32 spp) 386 at 1 against 366 at 8: a round of the tail costs ~45 us whatever the load
uint32_t tailCapacity(int numSMs) { return (uint32_t)numSMs * (uint32_t)kTailBlock * SLR_TAIL_PATHS_PER_THREAD; }

int launchTail(const SlrGpuScene* sc, const RenderConstants& rc, const PathQueue& q0, const PathQueue& q1, const HitBuffer& hits,
               const ShadowQueue& sq, float* accum, WavefrontCounters* counters, WavefrontCounters* ring, uint32_t ringSize,
               uint32_t cap, cudaStream_t stream) {
    if (cap == 0) return SLRGPU_OK;
    int numSMs = 148;
    cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, sc->device);
    const uint32_t grid = std::min((uint32_t)numSMs, (cap + kTailBlock - 1) / kTailBlock);      // one resident block per SM
    if (sc->channels == 3) tailKernel<3><<<grid, kTailBlock, 0, stream>>>(sc->dev, rc, q0, q1, hits, sq, accum, counters, cap, sc->classMask);
    else tailKernel<16><<<grid, kTailBlock, 0, stream>>>(sc->dev, rc, q0, q1, hits, sq, accum, counters, cap, sc->classMask);
    tailEndKernel<<<1, 1, 0, stream>>>(counters, ring, ringSize, cap);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

}  // namespace slrgpu
