// Wavefront path-tracing state: the buffers that link the ray-generation, extend, shade and shadow
// kernels. Every stage reads a compacted queue front to back (fully coalesced) and appends its
// survivors to the next queue with one warp-aggregated atomic per warp, so the path state itself is
// what moves through the queues (no slot indirection).
//
// HBM layout, all SoA over the queue position i (capacity P):
//   PathQueue   org[i]   = (ray origin xyz, distMin)                         16 B
//               dir[i]   = (ray direction xyz, pdf the direction was sampled with)   16 B
//               meta[i]  = (pixel, global sample index, hero | flags<<8 | pathLength<<16, bits(wavelength offset))  16 B
//               weight[i]= camera-sample weight (Job::kernel, PathTracingRenderer.cpp:126)  4 B
//               aux[i]   = importance(alpha), then the Russian-roulette scale                4 B
//               alpha[q*P + i], q < NC/4 = path throughput, four components per 16 B load   64 B (16 B in RGB mode)
//   HitBuffer   id[i] = (prim, inst), tuv[i] = (t, b0, b1, bits(surface info of the hit triangle | miss))   24 B
//   ShadowQueue org/dir (distMin / distMax in .w), pixel+wavelength offset, contribution[q*P + i]   108 B
// S_state = 120 B per path per stage transition in spectral mode.
#pragma once
#include "device_scene.h"

namespace slrgpu {

constexpr uint32_t kFlagLambdaSelected = 1u, kFlagPrevDelta = 2u, kFlagCameraRay = 4u, kFlagStrataInPlace = 8u;

struct PathQueue {
    float4* org;
    float4* dir;
    uint4* meta;
    float* weight;
    float* aux;            // written by the stage that produced the entry: importance(alpha) (the Russian-roulette
                           // probability of the next hit); overwritten by `surface` with the roulette scale 1/q
    float4* alpha;
    float* time;           // ray time (Ray::time), only allocated for scenes with animated transforms (else null)
    uint32_t capacity;     // stride of the alpha quarters
};

struct HitBuffer {
    uint2* id;
    float4* tuv;
};

struct ShadowQueue {
    float4* org;
    float4* dir;
    uint2* pixelWl;        // pixel | strataInPlace << 31, bits(wavelength offset)
    float4* contrib;
    float* time;           // as PathQueue::time
    uint32_t capacity;     // stride of the contribution quarters
};

// Material class of a surviving hit: the lobe type of a single-lobe BSDF (numbered like LobeType in
// bsdf.cuh) or SC_GENERIC for anything wrapped in a MultiBSDF / InverseBSDF. The `surface` kernel
// sorts survivors into one queue per class so that every `material` kernel launch runs one BSDF
// model with full warps (the reference dispatches per hit through virtual calls).
enum ShadeClass : uint32_t {
    SC_LAMBERT = 0, SC_OREN_NAYAR = 1, SC_SPECULAR_BRDF = 2, SC_SPECULAR_BSDF = 3, SC_WARD = 4, SC_ASHIKHMIN = 5,
    SC_MF_BRDF = 6, SC_MF_BSDF = 7, SC_GENERIC = 8, SC_COUNT = 9, SC_NONE = 0xFFu
};
// Row SC_COUNT of the class queues is not a material class: the `surface` stage of a wave queues there the entries whose
// arriving ray SEES EMISSION (it hit an emitter, or left the scene into the environment), as (queue position, isEnv);
// emissionKernel adds them to the sensor with full warps. Inline, that path ran for 30 % of the warps of a later wave at
// 1.2 of 32 lanes and was half of the stage's instructions (ncu, wave 2 of C1, profiles/r02_ncu_c1_final.md).
constexpr uint32_t kEmissionRow = SC_COUNT;
constexpr uint32_t kClassQueueRows = SC_COUNT + 1;

// Device-resident loop state: every kernel of a wave reads its work size from here, so the host
// never has to wait for a count (it only polls, two waves behind, for termination).
struct WavefrontCounters {
    uint32_t numNext;          // entries appended to the next path queue   } one aligned 64-bit word: the material kernels
    uint32_t numShadow;        // entries appended to the shadow queue      } reserve both positions with ONE atomic per warp
    uint32_t numPaths;         // entries of the current path queue
    uint32_t stackOverflow;
    uint32_t classCount[16];   // entries of each material-class queue (SC_COUNT used)
    unsigned long long generated, total;          // camera samples started / to render in this call
    unsigned long long extendRays, shadowRays;
    unsigned long long extendNodes, extendLeafRecords, shadowNodes, shadowLeafRecords;   // only counted by the profiling variants
    uint32_t waves, done;
    uint32_t genPass, genOffset;                  // pass / pixel-order position of the next camera sample
    uint32_t extendCursor, shadowCursor;          // chunk cursors of the warp-cooperative ray kernels (zero at launch)
    unsigned long long classTotal[16];            // hits shaded per material class over the whole call
    uint32_t tailPaths, tailWaves;                // paths the tail kernel (tail.cu) finished / bounce rounds of its longest-running warp
    uint32_t tailCursor, tailPad;                 // next queue entry the tail kernel hands out
};

// One entry of a material-class queue: position in the current path queue + the leaf material id.
struct ClassQueue {
    uint2* entries;        // [kClassQueueRows][capacity]
    uint32_t capacity;
};

struct RenderConstants {
    uint32_t width, height, numPixels;
    uint32_t sppBegin;
    uint32_t capacity;          // P
    uint32_t maxPathLength;     // 100 (PathTracingRenderer.cpp:162)
    uint32_t seed;
    float timeStart, timeEnd;
    // camera constants derived once on the host (PerspectiveCamera ctor, PerspectiveCamera.cpp:15-24)
    float opWidth, opHeight, imgPlaneArea, lensAreaPDF;
    float selectWLPDF;          // 16/470 spectral, 1 RGB
    float recBinWidth;          // 16/470 spectral (SpectrumStorage::add), 1 RGB
};

// launch helpers implemented in trace.cu (compiled with -fmad=false: same arithmetic as the batch API)
// Grid-stride launches over counters->numPaths / counters->numShadow entries (`grid` blocks).
// `count` selects the variants that also total QBVH nodes popped / leaf records tested (the algorithmic-bytes model)
int launchExtend(const SlrGpuScene* sc, const PathQueue& q, const HitBuffer& hits, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream);
int launchShadow(const SlrGpuScene* sc, const ShadowQueue& q, float* accum, WavefrontCounters* counters, bool count, uint32_t grid, cudaStream_t stream);
constexpr int kTraceBlock = 128;
// tail.cu: the persistent kernel that finishes the last <= cap paths of a call once all camera samples are started
// (a no-op before), followed by the single-thread kernel that closes the loop state; both go after a wave PAIR
// (the current path queue is q0 again). cap = 0: nothing is launched.
uint32_t tailCapacity(int numSMs);
int launchTail(const SlrGpuScene* sc, const RenderConstants& rc, const PathQueue& q0, const PathQueue& q1, const HitBuffer& hits,
               const ShadowQueue& sq, float* accum, WavefrontCounters* counters, WavefrontCounters* ring, uint32_t ringSize,
               uint32_t cap, cudaStream_t stream);

// bpt.cu: the bidirectional path tracer behind SLRGPU_RENDER_BPT (one kernel launch per call, synchronised before return)
int renderBpt(SlrGpuScene* sc, const SlrGpuRenderParams* p, const RenderConstants& rc, float* accumDev, cudaStream_t stream, SlrGpuRenderStats* stats);
void releaseBptWorkspaces();

// Stratum of wavelength i for a path with stratification offset `wlOffset`: min(uint((lambda_i - 360)
// / 470 * 16), 15) (SpectrumTypes.h:826-835) in uncontracted fp32 as the x86-64 reference computes it
// (the index is a rounding decision). It equals i except at rounding boundaries.
__device__ __forceinline__ uint32_t stratumOf(int i, float wlOffset) {
    const float lambda = __fadd_rn(360.0f, __fdiv_rn(__fmul_rn(470.0f, __fadd_rn((float)i, wlOffset)), 16.0f));
    return min((uint32_t)__fmul_rn(__fdiv_rn(__fsub_rn(lambda, 360.0f), 470.0f), 16.0f), 15u);
}
__device__ __forceinline__ bool strataInPlace(float wlOffset) {
    bool inPlace = true;
#pragma unroll
    for (int i = 0; i < 16; ++i) inPlace = inPlace && (stratumOf(i, wlOffset) == (uint32_t)i);
    return inPlace;
}

// accum[pixel * NC + stratum(i)] += v[i]: the sensor splat of ImageSensor::add / SpectrumStorage::add
// (ImageSensor.cpp:124-129). `inPlace` (decided once per path at ray generation) says every
// wavelength falls into its own stratum: the splat is then four 16-byte vector reductions
// (RED.E.ADD.F32x4.FTZ.RN), else sixteen scalar ones at the computed strata.
template <int NC>
__device__ __forceinline__ void splat(float* __restrict__ accum, uint32_t pixel, float wlOffset, bool inPlace, const float* v) {
    float* px = accum + (size_t)pixel * NC;
    if (NC == 16) {
        if (inPlace) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(px + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                             : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(px + stratumOf(i, wlOffset), v[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NC; ++i) atomicAdd(px + i, v[i]);
    }
}

}  // namespace slrgpu
