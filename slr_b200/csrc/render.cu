// The wavefront path tracer behind slrgpu_render: ray generation and shading kernels plus the host
// loop that drives raygen -> extend -> shade -> shadow over compacted queues.
//   PathTracingRenderer::render      libSLR/Renderers/PathTracingRenderer.cpp:27-98   (pass loop -> queue loop)
//   Job::kernel                      PathTracingRenderer.cpp:100-135                  (raygenKernel + splat)
//   Job::contribution                PathTracingRenderer.cpp:137-261                  (shadeKernel, one bounce per launch)
//   PerspectiveCamera / IDF sample   libSLR/Cameras/PerspectiveCamera.cpp:33-74
//   WavelengthSamples                libSLR/BasicTypes/SpectrumTypes.h:54-64
// The reference walks one path at a time per CPU thread; here up to `pool_size` paths are in flight
// and every kernel launch advances all of them by one stage. Finished paths are replaced by new
// camera samples (path regeneration) until the sample range is exhausted.
#include "camera.cuh"
#include "stages.cuh"
#include "traverse.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace slrgpu {

#ifndef SLR_SURFACE_BLOCK
#define SLR_SURFACE_BLOCK 128
#endif
constexpr int kSurfaceBlock = SLR_SURFACE_BLOCK;
constexpr int kMaterialBlock = 128;
constexpr int kRaygenBlock = 128;
// resident blocks per SM the register allocation of the shade kernels is bounded for (tuning knobs, see profiles/)
#ifndef SLR_MATERIAL_MIN_BLOCKS
#define SLR_MATERIAL_MIN_BLOCKS 3
#endif
#ifndef SLR_SURFACE_MIN_BLOCKS
#define SLR_SURFACE_MIN_BLOCKS 6
#endif
#ifndef SLR_MATERIAL_MIN_BLOCKS_HEAVY
#define SLR_MATERIAL_MIN_BLOCKS_HEAVY 4
#endif
// Lambert and the two specular classes run best at 3 resident blocks (168 registers); the heavier BSDFs (Oren-Nayar, Ward,
// Ashikhmin, the microfacet pair, multi-lobe) gain from a fourth block at 128 registers -- profiles/r01_variant_sweep.md
constexpr int materialMinBlocks(int cls) {
    return (cls == SC_LAMBERT || cls == SC_SPECULAR_BRDF || cls == SC_SPECULAR_BSDF) ? SLR_MATERIAL_MIN_BLOCKS : SLR_MATERIAL_MIN_BLOCKS_HEAVY;
}


// ---------------------------------------------------------------------------------------------
// ray generation: fills the free tail of the current queue with fresh camera samples
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kRaygenBlock)
raygenKernel(const DeviceScene s, const RenderConstants rc, PathQueue out, const WavefrontCounters* __restrict__ counters) {
    const uint32_t outBase = counters->numPaths;
    const unsigned long long remaining = counters->total - counters->generated;
    const uint32_t room = rc.capacity - outBase;
    const uint32_t count = (uint32_t)(remaining < (unsigned long long)room ? remaining : (unsigned long long)room);
    // next sample = pixel-order position genOffset of pass genPass (kept by beginWaveKernel: no 64-bit division here)
    const uint32_t genPass = counters->genPass, genOffset = counters->genOffset;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        const uint32_t lin = genOffset + j;                       // < numPixels + capacity < 2^32
        const uint32_t pass = genPass + lin / rc.numPixels;
        const uint32_t r = lin % rc.numPixels;
        // pixel order: bands of 8 rows, column-major inside a band, so a warp covers a 4x8 pixel block
        const uint32_t band = r / (8u * rc.width);
        const uint32_t local = r - band * 8u * rc.width;
        const uint32_t rows = min(8u, rc.height - band * 8u);
        const uint32_t x = local / rows, y = band * 8u + local % rows;
        const uint32_t pixel = y * rc.width + x;
        const uint32_t sample = rc.sppBegin + pass;

        CameraSample cs;
        sampleCamera<NC>(s, rc, x, y, pixel, sample, &cs);
        const V3 org = cs.org, dir = cs.dir;
        const float weight = cs.weight, wlOffset = cs.wlOffset;
        const uint32_t ipx = cs.ipx, ipy = cs.ipy, hero = cs.hero, flags = cs.flags;

        const uint32_t pos = outBase + j;
        out.org[pos] = make_float4(org.x, org.y, org.z, 0.0f);
        out.dir[pos] = make_float4(dir.x, dir.y, dir.z, 0.0f);
        out.meta[pos] = make_uint4(ipy * rc.width + ipx, sample, hero | (flags << 8), __float_as_uint(wlOffset));
        out.weight[pos] = weight * rc.recBinWidth;
        if (out.time) out.time[pos] = cs.time;
        // no throughput (it is 1) and no roulette slot for a camera ray: the stages know from the flag (stages.cuh)
    }
}

// after raygen: account for the fresh samples (single thread)
__global__ void beginWaveKernel(const RenderConstants rc, WavefrontCounters* counters) {
    const uint32_t n = counters->numPaths;
    const unsigned long long remaining = counters->total - counters->generated;
    const uint32_t room = rc.capacity - n;
    const uint32_t fresh = (uint32_t)(remaining < (unsigned long long)room ? remaining : (unsigned long long)room);
    counters->numPaths = n + fresh;
    counters->generated += fresh;
    counters->extendRays += n + fresh;
    const uint32_t lin = counters->genOffset + fresh;
    counters->genPass += lin / rc.numPixels;
    counters->genOffset = lin % rc.numPixels;
}

// after shadow: the next queue becomes the current one (single thread)
// `ring` is pinned host memory (device-visible through UVA): the snapshot the host polls
// `log` (SLRGPU_WAVE_LOG runs only, else null): per wave the device clock and the sizes of the queues it leaves
__global__ void endWaveKernel(WavefrontCounters* counters, WavefrontCounters* ring, uint32_t ringSize, ulonglong2* log, uint32_t logSize) {
    if (log && counters->waves < logSize) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        log[counters->waves] = make_ulonglong2(t, (unsigned long long)counters->numNext | ((unsigned long long)counters->numShadow << 32));
    }
    counters->shadowRays += counters->numShadow;
    counters->numPaths = counters->numNext;
    counters->numNext = 0;
    counters->numShadow = 0;
    for (int c = 0; c < 16; ++c) { counters->classTotal[c] += counters->classCount[c]; counters->classCount[c] = 0; }
    counters->extendCursor = 0;
    counters->shadowCursor = 0;
    counters->done = (counters->numPaths == 0 && counters->generated == counters->total) ? 1u : 0u;
    const uint32_t slot = counters->waves % ringSize;
    counters->waves += 1;
    ring[slot] = *counters;
    __threadfence_system();
}

// last node of the device-driven loop's body: the WHILE node (renderImpl) runs the body again while paths are in flight
__global__ void loopConditionKernel(const WavefrontCounters* __restrict__ counters, cudaGraphConditionalHandle handle) {
    cudaGraphSetConditional(handle, counters->done ? 0u : 1u);
}

// the shade stages (stages.cuh) as grid-stride kernels over the device-resident counts
template <int NC>
__global__ void __launch_bounds__(kSurfaceBlock, SLR_SURFACE_MIN_BLOCKS)
surfaceKernel(const DeviceScene s, const RenderConstants rc, PathQueue in, HitBuffer hits, ClassQueue cq,
              float* __restrict__ accum, WavefrontCounters* counters) {
    surfaceStage<NC, kSurfaceBlock>(s, rc, in, hits, cq, accum, counters, counters->numPaths);
}

template <int NC>
__global__ void __launch_bounds__(128)
emissionKernel(const DeviceScene s, PathQueue in, HitBuffer hits, ClassQueue cq, float* __restrict__ accum, const WavefrontCounters* counters) {
    emissionStage<NC>(s, in, hits, cq, accum, counters->classCount[kEmissionRow]);
}

template <int NC, int CLASS>
__global__ void __launch_bounds__(kMaterialBlock, materialMinBlocks(CLASS))
materialKernel(const DeviceScene s, const RenderConstants rc, PathQueue in, HitBuffer hits, ClassQueue cq, PathQueue out, ShadowQueue sq,
               WavefrontCounters* counters) {
    materialStage<NC, CLASS>(s, rc, in, hits, cq, out, sq, counters, counters->classCount[CLASS]);
}

// ---------------------------------------------------------------------------------------------
// debug (AOV) renderer: DebugRenderer::Job::kernel / contribution (libSLR/Renderers/DebugRenderer.cpp:132-217) -- one
// camera sample per pixel, closest hit, Intersection::getSurfacePoint; what the reference quantises into its
// geometric_normal / shading_normal / shading_tangent images is written as floats:
//   out[pixel * SLRGPU_DEBUG_FLOATS + 0] = 1 hit / 0 miss, 1-3 geometric normal, 4-6 shading normal, 7-9 shading tangent
// (a miss leaves the zero vectors of the reference's default-constructed DebugInfo)
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(128)
debugKernel(const DeviceScene s, const RenderConstants rc, float* __restrict__ out, uint32_t* stackOverflow) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= rc.numPixels) return;
    const uint32_t x = pixel % rc.width, y = pixel / rc.width;
    CameraSample cs;
    sampleCamera<NC>(s, rc, x, y, pixel, rc.sppBegin, &cs);
    uint32_t stack[kStackSize];
    WalkState w;
    w.r.ox = cs.org.x; w.r.oy = cs.org.y; w.r.oz = cs.org.z; w.r.tmin = 0.0f;
    w.r.dx = cs.dir.x; w.r.dy = cs.dir.y; w.r.dz = cs.dir.z; w.r.tmax = INFINITY;
    w.time = cs.time;
    InstanceWalkState iw;
    iw.leaves.clear(); iw.saved.clear(); iw.curInst = SLRGPU_INVALID_ID;
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkBegin(w, stack);
    while (!walkStep<true, false, false, true>(s, w, iw, stack, cnt, overflow)) { }
    if (overflow) atomicExch(stackOverflow, 1u);
    float* o = out + (size_t)(cs.ipy * rc.width + cs.ipx) * SLRGPU_DEBUG_FLOATS;
#pragma unroll
    for (int k = 0; k < SLRGPU_DEBUG_FLOATS; ++k) o[k] = 0.0f;
    if (w.hit.prim == SLRGPU_INVALID_ID) return;
    SurfPt sp;
    float localArea;
    hitSurfacePoint(s, w.hit.prim, w.hit.inst, w.hit.t, w.hit.u, w.hit.v, cs.org, cs.dir, cs.time, &sp, &localArea);
    o[0] = 1.0f;
    o[1] = sp.gn.x; o[2] = sp.gn.y; o[3] = sp.gn.z;
    o[4] = sp.sf.z.x; o[5] = sp.sf.z.y; o[6] = sp.sf.z.z;
    o[7] = sp.sf.x.x; o[8] = sp.sf.x.y; o[9] = sp.sf.x.z;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct RenderWorkspace {
    void* ptrs[48] = {};
    int n = 0;
    uint32_t capacity = 0, channels = 0;
    bool motion = false;                         // the queues carry ray times (scene with animated transforms)
    PathQueue q[2];
    HitBuffer hits;
    ShadowQueue sq;
    ClassQueue cq;
    ulonglong2* dWaveLog = nullptr;              // kWaveLogSize entries, written only in SLRGPU_WAVE_LOG runs
    WavefrontCounters* dCounters = nullptr;
    WavefrontCounters* hCounters = nullptr;      // pinned ring of kRing snapshots, written by endWaveKernel
    cudaEvent_t ringEvents[8] = {};
    // frame buffer of the host-buffer entry point (slrgpu_render) and its pinned staging copy
    float* frame = nullptr;
    float* frameHost = nullptr;
    size_t frameBytes = 0;
    cudaStream_t stream = nullptr;               // non-blocking stream of the host-buffer entry point (graph capture needs one)
    // the material kernels of a wave are independent of each other: they run on side streams, forked from and
    // joined to the render stream with events (parallel branches of the captured graph)
    cudaStream_t side[kClassQueueRows] = {};
    cudaEvent_t forkEvent = nullptr, joinEvent[kClassQueueRows] = {};
    // the instantiated device-driven loop of the last render call and everything that is baked into its kernel nodes: a
    // call with the same scene view, constants and buffers (a bench / progressive loop) relaunches it without re-capture
    cudaGraph_t loopGraph = nullptr;
    cudaGraphExec_t loopExec = nullptr;
    std::vector<unsigned char> loopKey;
    void dropLoopGraph() {
        if (loopExec) cudaGraphExecDestroy(loopExec);
        if (loopGraph) cudaGraphDestroy(loopGraph);
        loopExec = nullptr; loopGraph = nullptr; loopKey.clear();
    }
    int ensureFrame(size_t bytes) {
        if (frame && frameBytes >= bytes) return SLRGPU_OK;
        if (frame) cudaFree(frame);
        if (frameHost) cudaFreeHost(frameHost);
        frame = nullptr; frameHost = nullptr; frameBytes = 0;
        cudaError_t e = cudaMalloc(&frame, bytes);
        if (e != cudaSuccess) return cudaFail(e, "cudaMalloc(frame buffer)");
        e = cudaMallocHost(&frameHost, bytes);
        if (e != cudaSuccess) { cudaFree(frame); frame = nullptr; return cudaFail(e, "cudaMallocHost(frame staging)"); }
        frameBytes = bytes;
        return SLRGPU_OK;
    }
    ~RenderWorkspace() {
        dropLoopGraph();
        for (int i = 0; i < n; ++i) cudaFree(ptrs[i]);
        if (frame) cudaFree(frame);
        if (frameHost) cudaFreeHost(frameHost);
        if (stream) cudaStreamDestroy(stream);
        for (cudaStream_t st : side) if (st) cudaStreamDestroy(st);
        if (forkEvent) cudaEventDestroy(forkEvent);
        for (cudaEvent_t e : joinEvent) if (e) cudaEventDestroy(e);
        if (hCounters) cudaFreeHost(hCounters);
        for (cudaEvent_t e : ringEvents) if (e) cudaEventDestroy(e);
    }
    template <typename T> int alloc(T** p, uint64_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
        if (e != cudaSuccess) return cudaFail(e, "cudaMalloc(render workspace)");
        ptrs[n++] = q;
        *p = reinterpret_cast<T*>(q);
        return SLRGPU_OK;
    }
};
constexpr int kRing = 8;
constexpr uint32_t kWaveLogSize = 4096;

static int allocPathQueue(RenderWorkspace& b, PathQueue* q, uint32_t P, int quarters, bool motion) {
    int rc;
    q->time = nullptr;
    if (motion && (rc = b.alloc(&q->time, P))) return rc;
    if ((rc = b.alloc(&q->org, P))) return rc;
    if ((rc = b.alloc(&q->dir, P))) return rc;
    if ((rc = b.alloc(&q->meta, P))) return rc;
    if ((rc = b.alloc(&q->weight, P))) return rc;
    if ((rc = b.alloc(&q->aux, P))) return rc;
    if ((rc = b.alloc(&q->alpha, (uint64_t)P * quarters))) return rc;
    q->capacity = P;
    return SLRGPU_OK;
}

// The queues of a render call (about 440 B per path in flight: ~0.9 GB at the default pool) are kept
// in a process-wide pool, one per device, and reused by later calls of the same shape -- a renderer
// front end that uploads the scene for every render() call does not pay for them again.
// slrgpu_release_workspaces() frees the pool.
static std::mutex g_poolMutex;
static RenderWorkspace* g_pool[64] = {};

static void releaseWorkspace(int device, RenderWorkspace* w) {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    if (device >= 0 && device < 64 && !g_pool[device]) g_pool[device] = w;
    else delete w;
}

static int acquireWorkspace(SlrGpuScene* sc, uint32_t P, RenderWorkspace** out) {
    RenderWorkspace* w = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_poolMutex);
        if (sc->device >= 0 && sc->device < 64) { w = g_pool[sc->device]; g_pool[sc->device] = nullptr; }
    }
    if (w && w->capacity == P && w->channels == sc->channels && w->motion == sc->hasMotion) { *out = w; return SLRGPU_OK; }
    delete w;
    w = new (std::nothrow) RenderWorkspace();
    if (!w) { setError("host allocation failed"); return SLRGPU_ERR_OUT_OF_MEMORY; }
    const int quarters = sc->channels == 3 ? 1 : 4;
    int rc = SLRGPU_OK;
    for (int k = 0; k < 2 && !rc; ++k) rc = allocPathQueue(*w, &w->q[k], P, quarters, sc->hasMotion);
    if (!rc) rc = w->alloc(&w->hits.id, P);
    if (!rc) rc = w->alloc(&w->hits.tuv, P);
    if (!rc) rc = w->alloc(&w->sq.org, P);
    if (!rc) rc = w->alloc(&w->sq.dir, P);
    if (!rc) rc = w->alloc(&w->sq.pixelWl, P);
    if (!rc) rc = w->alloc(&w->sq.contrib, (uint64_t)P * quarters);
    w->sq.time = nullptr;
    if (!rc && sc->hasMotion) rc = w->alloc(&w->sq.time, P);
    if (!rc) rc = w->alloc(&w->cq.entries, (uint64_t)P * kClassQueueRows);
    if (!rc) rc = w->alloc(&w->dCounters, 1);
    if (!rc) rc = w->alloc(&w->dWaveLog, kWaveLogSize);
    if (!rc) { cudaError_t e = cudaMallocHost(&w->hCounters, sizeof(WavefrontCounters) * kRing); if (e != cudaSuccess) rc = cudaFail(e, "cudaMallocHost"); }
    for (int k = 0; k < kRing && !rc; ++k) { cudaError_t e = cudaEventCreateWithFlags(&w->ringEvents[k], cudaEventDisableTiming); if (e != cudaSuccess) rc = cudaFail(e, "cudaEventCreate"); }
    if (!rc) { cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking); if (e != cudaSuccess) rc = cudaFail(e, "cudaStreamCreate"); }
    for (int k = 0; k < (int)kClassQueueRows && !rc; ++k) {
        cudaError_t e = cudaStreamCreateWithFlags(&w->side[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->joinEvent[k], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = cudaFail(e, "cudaStreamCreate(side)");
    }
    if (!rc) { cudaError_t e = cudaEventCreateWithFlags(&w->forkEvent, cudaEventDisableTiming); if (e != cudaSuccess) rc = cudaFail(e, "cudaEventCreate"); }
    if (rc) { delete w; return rc; }
    w->sq.capacity = P; w->cq.capacity = P;
    w->capacity = P; w->channels = sc->channels; w->motion = sc->hasMotion;
    *out = w;
    return SLRGPU_OK;
}

// one class kernel on its side stream, between the fork event (surface done) and its join event
template <int NC, int CLASS>
static void launchMaterial(const SlrGpuScene* sc, const RenderConstants& rc, const RenderWorkspace& w, int cur, uint32_t grid, cudaStream_t stream) {
    cudaStream_t st = w.side[CLASS];
    cudaStreamWaitEvent(st, w.forkEvent, 0);
    static const int perSM = [] { int n = 0; if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, materialKernel<NC, CLASS>, kMaterialBlock, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; } return n; }();
    grid = std::min(grid, (uint32_t)(sc->numSMs * perSM));
    materialKernel<NC, CLASS><<<grid, kMaterialBlock, 0, st>>>(sc->dev, rc, w.q[cur], w.hits, w.cq, w.q[cur ^ 1], w.sq, w.dCounters);
    cudaEventRecord(w.joinEvent[CLASS], st);
    cudaStreamWaitEvent(stream, w.joinEvent[CLASS], 0);
}

// the shade stage of one wave: surface + one material launch per class the scene contains
template <int NC, typename Mark>
static void launchShadeStage(const SlrGpuScene* sc, const RenderConstants& rc, const RenderWorkspace& w, int cur, float* accum, uint32_t grid,
                             cudaStream_t stream, const Mark& mark) {
    mark(2);
    surfaceKernel<NC><<<residentGrid(surfaceKernel<NC>, kSurfaceBlock, sc->numSMs, grid), kSurfaceBlock, 0, stream>>>(sc->dev, rc, w.q[cur], w.hits, w.cq, accum, w.dCounters);
    mark(2);
    mark(3);
    cudaEventRecord(w.forkEvent, stream);
    const uint32_t m = sc->classMask;
    if (m & (1u << SC_LAMBERT)) launchMaterial<NC, SC_LAMBERT>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_OREN_NAYAR)) launchMaterial<NC, SC_OREN_NAYAR>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_SPECULAR_BRDF)) launchMaterial<NC, SC_SPECULAR_BRDF>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_SPECULAR_BSDF)) launchMaterial<NC, SC_SPECULAR_BSDF>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_WARD)) launchMaterial<NC, SC_WARD>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_ASHIKHMIN)) launchMaterial<NC, SC_ASHIKHMIN>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_MF_BRDF)) launchMaterial<NC, SC_MF_BRDF>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_MF_BSDF)) launchMaterial<NC, SC_MF_BSDF>(sc, rc, w, cur, grid, stream);
    if (m & (1u << SC_GENERIC)) launchMaterial<NC, SC_GENERIC>(sc, rc, w, cur, grid, stream);
    {   // the entries that see emission, beside the material kernels (both only read the current queue)
        cudaStream_t st = w.side[kEmissionRow];
        cudaStreamWaitEvent(st, w.forkEvent, 0);
        emissionKernel<NC><<<std::min(grid, (uint32_t)(sc->numSMs * 4)), 128, 0, st>>>(sc->dev, w.q[cur], w.hits, w.cq, accum, w.dCounters);
        cudaEventRecord(w.joinEvent[kEmissionRow], st);
        cudaStreamWaitEvent(stream, w.joinEvent[kEmissionRow], 0);
    }
    mark(3);
}

static uint32_t poolCapacity(const SlrGpuRenderParams* p) {
    const unsigned long long totalSamples = (unsigned long long)p->width * p->height * (p->spp_end - p->spp_begin);
    if (p->flags & SLRGPU_RENDER_BPT) return 128u;          // the bidirectional path tracer keeps its own vertex storage (bpt.cu)
    uint32_t P = p->pool_size ? p->pool_size : (1u << 24);     // 16 Mi paths in flight (7.4 GB of queues): C1 371 / 419 / 454 / 463 / 480 Mpaths/s at 1 / 2 / 4 / 8 / 16 Mi
    if ((unsigned long long)P > totalSamples) P = (uint32_t)totalSamples;
    P = (P + 127u) & ~127u;
    return P == 0 ? 128u : P;
}

// SLRGPU_WAVE_LOG=<file>: appends one line per wave of the call -- microseconds since the first wave ended, paths
// and shadow rays that wave left -- the timeline the ncu launch list cannot give (its kernels run serialised)
static void dumpWaveLog(const char* path, const ulonglong2* dLog, uint32_t n, const WavefrontCounters& last) {
    std::vector<ulonglong2> log(n);
    if (n && cudaMemcpy(log.data(), dLog, n * sizeof(ulonglong2), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
    FILE* f = fopen(path, "a");
    if (!f) return;
    fprintf(f, "# render call: %u waves, tail kernel: %u paths, %u bounces\n", last.waves, last.tailPaths, last.tailWaves);
    for (uint32_t i = 0; i < n; ++i)
        fprintf(f, "%u,%.1f,%u,%u\n", i, (double)(log[i].x - log[0].x) * 1e-3, (uint32_t)(log[i].y & 0xFFFFFFFFull), (uint32_t)(log[i].y >> 32));
    fclose(f);
}

static RenderConstants makeRenderConstants(const SlrGpuScene* sc, const SlrGpuRenderParams* p, uint32_t capacity) {
    const bool rgb = sc->channels == 3;
    RenderConstants rc;
    memset(&rc, 0, sizeof(rc));
    rc.width = p->width; rc.height = p->height; rc.numPixels = p->width * p->height;
    rc.sppBegin = p->spp_begin;
    rc.capacity = capacity;
    rc.maxPathLength = p->max_path_length ? p->max_path_length : 100;
    rc.seed = (uint32_t)p->rng_seed;
    rc.timeStart = p->time_start; rc.timeEnd = p->time_end;
    const SlrGpuCamera& cam = sc->dev.camera;
    rc.opHeight = 2.0f * cam.obj_plane_dist * std::tan(cam.fov_y * 0.5f);
    rc.opWidth = rc.opHeight * cam.aspect;
    rc.imgPlaneArea = rc.opWidth * rc.opHeight * (float)std::pow(cam.img_plane_dist / cam.obj_plane_dist, 2);
    rc.lensAreaPDF = cam.lens_radius > 0.0f ? (float)(1.0f / (M_PI * cam.lens_radius * cam.lens_radius)) : 1.0f;
    rc.selectWLPDF = rgb ? 1.0f : 16.0f / (830.0f - 360.0f);
    rc.recBinWidth = rgb ? 1.0f : 16.0f / (830.0f - 360.0f);
    return rc;
}

static int renderImpl(SlrGpuScene* sc, const SlrGpuRenderParams* p, RenderWorkspace& w, float* accumDev, cudaStream_t stream, SlrGpuRenderStats* stats) {
    const bool rgb = sc->channels == 3;
    const uint32_t W = p->width, H = p->height;
    const unsigned long long numPixels = (unsigned long long)W * H;
    const unsigned long long totalSamples = numPixels * (p->spp_end - p->spp_begin);
    const uint32_t P = w.capacity;

    const RenderConstants rc = makeRenderConstants(sc, p, P);
    // setRenderer("BPT"): the bidirectional path tracer (bpt.cu) instead of the wavefront loop
    if (p->flags & SLRGPU_RENDER_BPT) return renderBpt(sc, p, rc, accumDev, stream, stats);

    int rcode = SLRGPU_OK;

    // grid-stride launches: never more blocks than the queue needs; every launch site caps this at the blocks its kernel
    // keeps resident (residentGrid)
    const int numSMs = sc->numSMs;
    const uint32_t grid = (P + 127u) / 128u;

    cudaEvent_t ev0, ev1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{ev0, ev1};
    SLRGPU_CUDA_TRY(cudaEventRecord(ev0, stream));

    // per-stage device time (SLRGPU_RENDER_PROFILE_STAGES): one event pair per launch group, summed at the end
    const bool profile = (p->flags & SLRGPU_RENDER_PROFILE_STAGES) != 0;
    struct StageTimer {
        std::vector<cudaEvent_t> ev[6];      // 0 raygen, 1 extend, 2 surface, 3 material kernels, 4 shadow, 5 tail kernel: begin/end pairs
        bool on;
        cudaStream_t st;
        void mark(int stage) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev[stage].push_back(e); }
        float total(int stage) {
            float sum = 0.0f;
            for (size_t i = 0; i + 1 < ev[stage].size(); i += 2) { float ms = 0.0f; cudaEventElapsedTime(&ms, ev[stage][i], ev[stage][i + 1]); sum += ms; }
            return sum;
        }
        ~StageTimer() { for (auto& v : ev) for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } timer;
    timer.on = profile; timer.st = stream;

    WavefrontCounters init;
    memset(&init, 0, sizeof(init));
    init.total = totalSamples;
    // the pinned slot 0 doubles as the staging buffer of the initial state (copied before any wave is enqueued)
    w.hCounters[0] = init;
    SLRGPU_CUDA_TRY(cudaMemcpyAsync(w.dCounters, &w.hCounters[0], sizeof(WavefrontCounters), cudaMemcpyHostToDevice, stream));
    SLRGPU_CUDA_TRY(cudaStreamSynchronize(stream));

    // Waves are enqueued without waiting for their counts; endWaveKernel leaves a snapshot of the loop
    // state in a pinned ring and the host reads the snapshot of the wave pair enqueued kLag rounds ago
    // to learn when the work ran out (waves enqueued in between find empty queues and cost a few
    // microseconds). Outside profiling two waves (one ping + one pong of the path queues) are captured
    // into a CUDA graph once per call and replayed: one driver call per ~20 kernel launches.
    constexpr int kLag = 2;
    // the tail kernel (tail.cu) takes over when only a few paths per thread of it are left; SLRGPU_TAIL_PATHS overrides
    // the limit (0 = no tail kernel: every bounce is a wave) -- a tuning / test knob, the image does not depend on it
    // Two paths per thread when the scene has at most three material classes: a round of a tail warp runs the classes of its
    // lanes one after the other, so a warp that holds more paths pays per class -- C1 (3 classes) 748 -> 757 Mpaths/s and C4 (2)
    // 400 -> 404 at 75 776 paths, C2 (8 classes) 795 -> 768 (profiles/r02_variant_sweep.md)
    uint32_t tailCap = tailCapacity(numSMs) * (__builtin_popcount(sc->classMask) <= 3 ? 2u : 1u);
    if (const char* e = getenv("SLRGPU_TAIL_PATHS")) tailCap = (uint32_t)strtoul(e, nullptr, 10);
    const char* waveLogPath = getenv("SLRGPU_WAVE_LOG");
    ulonglong2* waveLog = waveLogPath ? w.dWaveLog : nullptr;
    auto enqueueTail = [&]() -> int {
        timer.mark(5);
        int r = launchTail(sc, rc, w.q[0], w.q[1], w.hits, w.sq, accumDev, w.dCounters, w.hCounters, (uint32_t)kRing, tailCap, stream);
        timer.mark(5);
        return r;
    };
    const uint32_t launchesPerWave = 7u + (uint32_t)__builtin_popcount(sc->classMask);      // raygen, begin, extend, surface, emission, shadow, end + one per class
    unsigned long long wave = 0, launches = 0, round = 0;
    auto enqueueWave = [&](int cur) -> int {
        timer.mark(0);
        if (rgb) raygenKernel<3><<<residentGrid(raygenKernel<3>, kRaygenBlock, sc->numSMs, grid), kRaygenBlock, 0, stream>>>(sc->dev, rc, w.q[cur], w.dCounters);
        else raygenKernel<16><<<residentGrid(raygenKernel<16>, kRaygenBlock, sc->numSMs, grid), kRaygenBlock, 0, stream>>>(sc->dev, rc, w.q[cur], w.dCounters);
        beginWaveKernel<<<1, 1, 0, stream>>>(rc, w.dCounters);
        timer.mark(0);
        timer.mark(1);
        int r = launchExtend(sc, w.q[cur], w.hits, w.dCounters, profile, grid, stream);
        if (r) return r;
        timer.mark(1);
        auto mark = [&timer](int stage) { timer.mark(stage); };
        if (rgb) launchShadeStage<3>(sc, rc, w, cur, accumDev, grid, stream, mark);
        else launchShadeStage<16>(sc, rc, w, cur, accumDev, grid, stream, mark);
        timer.mark(4);
        if ((r = launchShadow(sc, w.sq, accumDev, w.dCounters, profile, grid, stream))) return r;
        endWaveKernel<<<1, 1, 0, stream>>>(w.dCounters, w.hCounters, (uint32_t)kRing, waveLog, kWaveLogSize);
        timer.mark(4);
        return SLRGPU_OK;
    };
    // ---- the loop, device driven: ONE graph launch per render call. The graph is a WHILE conditional node (CUDA 12.4+)
    // whose body is the wave pair + tail kernel above plus loopConditionKernel, which re-arms the node while paths are in
    // flight: no host thread in the loop, no empty waves enqueued past the end, no stall when the host is descheduled
    // (with one process per GPU a stalled host used to show up as a straggler rank in the frame's reduce).
    // SLRGPU_HOST_LOOP=1 (and the per-stage profiling mode, which needs events between the kernels) takes the host-driven
    // loop below instead: two waves per graph launch, termination read from the pinned snapshot ring two launches behind.
    WavefrontCounters last = init;
    bool deviceLoop = false;
    if (!profile && !getenv("SLRGPU_HOST_LOOP")) {
        std::vector<unsigned char> key(sizeof(DeviceScene) + sizeof(RenderConstants) + 4 * sizeof(void*) + 4 * sizeof(uint32_t));
        {
            unsigned char* k = key.data();
            memcpy(k, &sc->dev, sizeof(DeviceScene)); k += sizeof(DeviceScene);
            memcpy(k, &rc, sizeof(RenderConstants)); k += sizeof(RenderConstants);
            const void* ptrs[4] = {accumDev, waveLog, w.dCounters, nullptr};
            memcpy(k, ptrs, sizeof(ptrs)); k += sizeof(ptrs);
            const uint32_t words[4] = {tailCap, sc->classMask, (uint32_t)sc->hasInstances | ((uint32_t)sc->hasAlpha << 1), sc->channels};
            memcpy(k, words, sizeof(words));
        }
        if (!w.loopExec || w.loopKey != key) {
            w.dropLoopGraph();
            bool ok = cudaGraphCreate(&w.loopGraph, 0) == cudaSuccess;
            cudaGraphConditionalHandle handle = 0;
            ok = ok && cudaGraphConditionalHandleCreate(&handle, w.loopGraph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
            cudaGraph_t body = nullptr;
            if (ok) {
                cudaGraphNodeParams np = {};
                np.type = cudaGraphNodeTypeConditional;
                np.conditional.handle = handle;
                np.conditional.type = cudaGraphCondTypeWhile;
                np.conditional.size = 1;
                cudaGraphNode_t node;
                ok = cudaGraphAddNode(&node, w.loopGraph, nullptr, 0, &np) == cudaSuccess;
                if (ok) body = np.conditional.phGraph_out[0];
            }
            if (ok && cudaStreamBeginCaptureToGraph(stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                int r = enqueueWave(0);
                if (!r) r = enqueueWave(1);
                if (!r) r = enqueueTail();
                if (!r) loopConditionKernel<<<1, 1, 0, stream>>>(w.dCounters, handle);
                cudaGraph_t captured = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(stream, &captured);
                if (r) { w.dropLoopGraph(); return r; }
                ok = ce == cudaSuccess && cudaGraphInstantiate(&w.loopExec, w.loopGraph, 0) == cudaSuccess;
            } else ok = false;
            if (ok) w.loopKey = key;
            else { w.dropLoopGraph(); cudaGetLastError(); }      // e.g. an older driver: fall back to the host-driven loop
        }
        if (w.loopExec) {
            SLRGPU_CUDA_TRY(cudaGraphLaunch(w.loopExec, stream));
            SLRGPU_CUDA_TRY(cudaEventRecord(ev1, stream));
            SLRGPU_CUDA_TRY(cudaEventSynchronize(ev1));
            SLRGPU_CUDA_TRY(cudaMemcpy(&last, w.dCounters, sizeof(last), cudaMemcpyDeviceToHost));
            wave = last.waves;
            launches = 1ull + (wave / 2ull) * (2ull * launchesPerWave + (tailCap ? 2u : 0u) + 1u);
            deviceLoop = true;
        }
    }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graphExec = nullptr;
    struct GraphFree { cudaGraph_t* g; cudaGraphExec_t* e; ~GraphFree() { if (*e) cudaGraphExecDestroy(*e); if (*g) cudaGraphDestroy(*g); } } graphFree{&graph, &graphExec};
    if (!profile && !deviceLoop) {
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int r0 = enqueueWave(0);
            int r1 = r0 ? r0 : enqueueWave(1);
            if (!r1) r1 = enqueueTail();
            cudaError_t ce = cudaStreamEndCapture(stream, &graph);
            if (r1) return r1;
            if (ce == cudaSuccess && graph) { if (cudaGraphInstantiate(&graphExec, graph, 0) != cudaSuccess) graphExec = nullptr; }
        }
        cudaGetLastError();     // a failed capture falls back to plain launches
    }
    while (!deviceLoop) {
        if (graphExec) { SLRGPU_CUDA_TRY(cudaGraphLaunch(graphExec, stream)); }
        else {
            if ((rcode = enqueueWave(0))) return rcode;
            if ((rcode = enqueueWave(1))) return rcode;
            if ((rcode = enqueueTail())) return rcode;
            SLRGPU_CUDA_TRY(cudaGetLastError());
        }
        wave += 2;
        launches += 2 * launchesPerWave + (tailCap ? 2u : 0u);
        SLRGPU_CUDA_TRY(cudaEventRecord(w.ringEvents[round % kRing], stream));
        ++round;
        if (round >= (unsigned long long)kLag) {
            const unsigned long long back = round - kLag;             // round whose two waves are now complete
            SLRGPU_CUDA_TRY(cudaEventSynchronize(w.ringEvents[back % kRing]));
            last = w.hCounters[(2 * back + 1) % kRing];              // snapshot of its second wave
            if (last.done) break;
        }
    }
    if (!deviceLoop) {
        SLRGPU_CUDA_TRY(cudaEventRecord(ev1, stream));
        SLRGPU_CUDA_TRY(cudaEventSynchronize(ev1));
        last = w.hCounters[(wave - 1) % kRing];
        wave = last.waves;
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->paths = totalSamples;
        stats->extend_rays = last.extendRays; stats->shadow_rays = last.shadowRays;
        stats->rays = last.extendRays + last.shadowRays;
        stats->kernel_launches = launches;
        stats->waves = wave;
        for (int c = 0; c < 9; ++c) stats->class_hits[c] = last.classTotal[c];
        stats->tail_paths = last.tailPaths; stats->tail_waves = last.tailWaves;
        cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
        if (profile) {
            stats->raygen_ms = timer.total(0); stats->extend_ms = timer.total(1);
            stats->surface_ms = timer.total(2); stats->material_ms = timer.total(3); stats->shadow_ms = timer.total(4);
            stats->tail_ms = timer.total(5);
            stats->other_ms = stats->device_ms - stats->raygen_ms - stats->extend_ms - stats->surface_ms - stats->material_ms - stats->shadow_ms - stats->tail_ms;
            stats->extend_nodes = last.extendNodes; stats->extend_leaf_records = last.extendLeafRecords;
            stats->shadow_nodes = last.shadowNodes; stats->shadow_leaf_records = last.shadowLeafRecords;
        }
    }
    if (waveLogPath) dumpWaveLog(waveLogPath, w.dWaveLog, std::min((uint32_t)last.waves, kWaveLogSize), last);
    if (last.stackOverflow) { setError("traversal stack overflow (more than %d entries)", 64); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// shading probe (slrgpu_probe_shading): the material kernels' device functions on caller-given inputs
// ---------------------------------------------------------------------------------------------
__global__ void probeRaysKernel(const float* __restrict__ probes, uint32_t n, SlrGpuRayBatch rays) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = probes + (size_t)i * SLRGPU_PROBE_IN_FLOATS;
    const_cast<float*>(rays.org_x)[i] = p[0]; const_cast<float*>(rays.org_y)[i] = p[1]; const_cast<float*>(rays.org_z)[i] = p[2];
    const_cast<float*>(rays.dir_x)[i] = p[3]; const_cast<float*>(rays.dir_y)[i] = p[4]; const_cast<float*>(rays.dir_z)[i] = p[5];
    const_cast<float*>(rays.tmin)[i] = 0.0f; const_cast<float*>(rays.tmax)[i] = INFINITY;
}

template <int CLASS>
__device__ __noinline__ void probeBsdf(const DeviceScene& s, uint32_t leaf, const SurfPt& sp, const V3& dirOut, const V3& gNorm, float wlOffset,
                                       uint32_t hero, float uComp, float u0, float u1, const V3& evalDir, float* o) {
    HitBsdf<16, CLASS> bsdf;
    bsdf.build(s, leaf, sp, wlOffset, false);
    BsdfQuery q;
    q.dir = dirOut; q.gn = gNorm; q.hero = hero; q.flags = DT_All;
    o[11] = bsdf.hasNonDelta() ? 1.0f : 0.0f;
    BsdfSampleResult res;
    const Spec<16> fs = bsdf.sample(q, uComp, u0, u1, &res);
    for (int k = 0; k < 16; ++k) o[12 + k] = fs.v[k];
    o[28] = res.dir.x; o[29] = res.dir.y; o[30] = res.dir.z;
    o[31] = res.pdf;
    o[32] = (float)res.type;
    const Spec<16> fe = bsdf.evaluate(q, evalDir);
    for (int k = 0; k < 16; ++k) o[33 + k] = fe.v[k];
    o[49] = bsdf.pdf(q, evalDir);
}

__global__ void __launch_bounds__(64)
probeShadeKernel(const DeviceScene s, const float* __restrict__ probes, uint32_t n, SlrGpuHitBatch hits, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = probes + (size_t)i * SLRGPU_PROBE_IN_FLOATS;
    float* o = out + (size_t)i * SLRGPU_PROBE_OUT_FLOATS;
    for (int k = 0; k < SLRGPU_PROBE_OUT_FLOATS; ++k) o[k] = 0.0f;
    const uint32_t prim = hits.prim[i];
    if (prim == SLRGPU_INVALID_ID) { o[0] = s.envPresent ? 2.0f : 0.0f; return; }
    const V3 org(p[0], p[1], p[2]), dir(p[3], p[4], p[5]);
    const float wlOffset = p[6];
    const uint32_t hero = min((uint32_t)(16 * p[7]), 15u);
    SurfPt sp;
    float localArea;
    const SlrGpuTriangle tri = hitSurfacePoint(s, prim, hits.inst[i], hits.t[i], hits.u[i], hits.v[i], org, dir, 0.0f, &sp, &localArea);
    o[0] = 1.0f; o[1] = hits.t[i];
    o[2] = sp.p.x; o[3] = sp.p.y; o[4] = sp.p.z;
    o[5] = sp.sf.z.x; o[6] = sp.sf.z.y; o[7] = sp.sf.z.z;
    o[8] = sp.sf.x.x; o[9] = sp.sf.x.y; o[10] = sp.sf.x.z;
    const V3 dirOut = sp.sf.toLocal(-dir);
    const V3 gNorm = sp.sf.toLocal(sp.gn);
    const V3 evalDir = sp.sf.toLocal(V3(p[11], p[12], p[13]));
    uint32_t leaf = SLRGPU_INVALID_ID;
    const uint32_t cls = classifyMaterial(s, tri.material, &leaf);
    switch (cls) {
    case SC_LAMBERT: probeBsdf<SC_LAMBERT>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_OREN_NAYAR: probeBsdf<SC_OREN_NAYAR>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_SPECULAR_BRDF: probeBsdf<SC_SPECULAR_BRDF>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_SPECULAR_BSDF: probeBsdf<SC_SPECULAR_BSDF>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_WARD: probeBsdf<SC_WARD>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_ASHIKHMIN: probeBsdf<SC_ASHIKHMIN>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_MF_BRDF: probeBsdf<SC_MF_BRDF>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_MF_BSDF: probeBsdf<SC_MF_BSDF>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    case SC_GENERIC: probeBsdf<SC_GENERIC>(s, leaf, sp, dirOut, gNorm, wlOffset, hero, p[8], p[9], p[10], evalDir, o); break;
    default: break;
    }
    if (materialIsEmitting(s, tri.material)) {
        o[50] = 1.0f;
        const Spec<16> Le = materialEmittance<16>(s, tri.material, sp, wlOffset);
        for (int k = 0; k < 13; ++k) o[51 + k] = Le.v[k];
    }
}

// device -> caller's host buffer, on the workspace's stream, synchronised before return
static int downloadFrame(RenderWorkspace& w, float* accum, size_t bytes) {
    // a caller buffer that is page-locked (cudaHostAlloc / cudaHostRegister, e.g. a pinned framework tensor) takes the DMA directly
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, accum) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        SLRGPU_CUDA_TRY(cudaMemcpyAsync(accum, w.frame, bytes, cudaMemcpyDeviceToHost, w.stream));
        SLRGPU_CUDA_TRY(cudaStreamSynchronize(w.stream));
        return SLRGPU_OK;
    }
    cudaGetLastError();
    // device -> pinned staging -> the caller's (pageable) buffer, in pieces: the host copy of piece k runs while piece
    // k + 1 is still on the bus
    constexpr int kPieces = 8;
    const size_t piece = ((bytes / kPieces) + 4095) & ~(size_t)4095;
    int pieces = 0;
    for (size_t off = 0; off < bytes && pieces < kPieces; off += piece, ++pieces) {
        const size_t len = std::min(piece, bytes - off);
        SLRGPU_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(w.frameHost) + off, reinterpret_cast<const char*>(w.frame) + off, len,
                                        cudaMemcpyDeviceToHost, w.stream));
        SLRGPU_CUDA_TRY(cudaEventRecord(w.ringEvents[pieces], w.stream));
    }
    for (int k = 0; k < pieces; ++k) {
        const size_t off = (size_t)k * piece;
        SLRGPU_CUDA_TRY(cudaEventSynchronize(w.ringEvents[k]));
        memcpy(reinterpret_cast<char*>(accum) + off, reinterpret_cast<const char*>(w.frameHost) + off, std::min(piece, bytes - off));
    }
    return SLRGPU_OK;
}

// the multi-GPU exchange step: dst += sum of the other replicas' frames, read where they lie (peer memory over NVLink)
constexpr uint32_t kMaxReplicas = 16;
struct PeerFrames { const float* frame[kMaxReplicas]; uint32_t count; };
__global__ void __launch_bounds__(256)
sumFramesKernel(float* __restrict__ dst, const PeerFrames peers, uint32_t count) {
    const uint32_t n4 = count / 4;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 a = d4[i];
        for (uint32_t g = 0; g < peers.count; ++g) {
            const float4 b = reinterpret_cast<const float4*>(peers.frame[g])[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        d4[i] = a;
    }
    for (uint32_t i = n4 * 4 + blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        float a = dst[i];
        for (uint32_t g = 0; g < peers.count; ++g) a += peers.frame[g][i];
        dst[i] = a;
    }
}

static int checkRenderArgs(SlrGpuScene* sc, const SlrGpuRenderParams* p) {
    if (!sc || !p) { setError("slrgpu_render: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (p->struct_size != sizeof(SlrGpuRenderParams)) { setError("slrgpu_render: params struct_size mismatch"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (!sc->hasShading) { setError("slrgpu_render: the scene has no materials / camera (geometry-only scene)"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (p->width == 0 || p->height == 0 || p->spp_end <= p->spp_begin) { setError("slrgpu_render: empty image or sample range"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if ((unsigned long long)p->width * p->height >= 0x80000000ull) { setError("slrgpu_render: image too large"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    return SLRGPU_OK;
}

}  // namespace slrgpu

using namespace slrgpu;

extern "C" {

SLRGPU_API int slrgpu_probe_shading(SlrGpuScene* sc, const float* probes, uint64_t n, float* out) {
    if (!sc || !probes || !out) { setError("slrgpu_probe_shading: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (!sc->hasShading || sc->channels != 16) { setError("slrgpu_probe_shading: needs a spectral scene with materials"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (n == 0) return SLRGPU_OK;
    if (n > (1u << 26)) { setError("slrgpu_probe_shading: too many probes"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    struct Buffers { void* p[16] = {}; int k = 0; ~Buffers() { for (int i = 0; i < k; ++i) cudaFree(p[i]); } } bufs;
    auto alloc = [&bufs](size_t bytes) -> void* { void* q = nullptr; if (cudaMalloc(&q, bytes) != cudaSuccess) return nullptr; bufs.p[bufs.k++] = q; return q; };
    float* dProbes = (float*)alloc(n * SLRGPU_PROBE_IN_FLOATS * sizeof(float));
    float* dOut = (float*)alloc(n * SLRGPU_PROBE_OUT_FLOATS * sizeof(float));
    float* comp[8];
    for (int k = 0; k < 8; ++k) comp[k] = (float*)alloc(n * sizeof(float));
    SlrGpuHitBatch hits = {};
    hits.prim = (uint32_t*)alloc(n * 4); hits.inst = (uint32_t*)alloc(n * 4);
    hits.t = (float*)alloc(n * 4); hits.u = (float*)alloc(n * 4); hits.v = (float*)alloc(n * 4);
    int* dStatus = (int*)alloc(2 * sizeof(int));
    if (!dProbes || !dOut || !comp[7] || !hits.v || !dStatus) { setError("slrgpu_probe_shading: out of device memory"); return SLRGPU_ERR_OUT_OF_MEMORY; }
    SLRGPU_CUDA_TRY(cudaMemcpy(dProbes, probes, n * SLRGPU_PROBE_IN_FLOATS * sizeof(float), cudaMemcpyHostToDevice));
    SLRGPU_CUDA_TRY(cudaMemset(dStatus, 0, 2 * sizeof(int)));
    SlrGpuRayBatch rays = {comp[0], comp[1], comp[2], comp[3], comp[4], comp[5], comp[6], comp[7]};
    const uint32_t n32 = (uint32_t)n;
    probeRaysKernel<<<(n32 + 127) / 128, 128>>>(dProbes, n32, rays);
    int rc = launchIntersect(sc, rays, n, hits, dStatus, 0);
    if (rc) return rc;
    probeShadeKernel<<<(n32 + 63) / 64, 64>>>(sc->dev, dProbes, n32, hits, dOut);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaDeviceSynchronize());
    SLRGPU_CUDA_TRY(cudaMemcpy(out, dOut, n * SLRGPU_PROBE_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost));
    return SLRGPU_OK;
}

SLRGPU_API void slrgpu_release_workspaces(void) {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    for (int d = 0; d < 64; ++d)
        if (g_pool[d]) { cudaSetDevice(d); delete g_pool[d]; g_pool[d] = nullptr; }
    releaseBptWorkspaces();
    releaseSceneArenas();
}

SLRGPU_API int slrgpu_render_device(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* accumDev, void* stream, SlrGpuRenderStats* stats) {
    int rc = checkRenderArgs(sc, p);
    if (rc) return rc;
    if (!accumDev) { setError("slrgpu_render_device: null accumulation buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    RenderWorkspace* w = nullptr;
    if ((rc = acquireWorkspace(sc, poolCapacity(p), &w))) return rc;
    struct Release { int device; RenderWorkspace* w; ~Release() { releaseWorkspace(device, w); } } release{sc->device, w};
    return renderImpl(sc, p, *w, accumDev, (cudaStream_t)stream, stats);
}

SLRGPU_API int slrgpu_render_debug(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* out, SlrGpuRenderStats* stats) {
    int rc = checkRenderArgs(sc, p);
    if (rc) return rc;
    if (!out) { setError("slrgpu_render_debug: null output buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    const uint32_t numPixels = p->width * p->height;
    const size_t bytes = (size_t)numPixels * SLRGPU_DEBUG_FLOATS * sizeof(float);
    struct Buffers { void* p[2] = {}; ~Buffers() { for (void* q : p) if (q) cudaFree(q); } } bufs;
    SLRGPU_CUDA_TRY(cudaMalloc(&bufs.p[0], bytes));
    SLRGPU_CUDA_TRY(cudaMalloc(&bufs.p[1], sizeof(uint32_t)));
    SLRGPU_CUDA_TRY(cudaMemset(bufs.p[0], 0, bytes));
    SLRGPU_CUDA_TRY(cudaMemset(bufs.p[1], 0, sizeof(uint32_t)));
    const RenderConstants rcs = makeRenderConstants(sc, p, 0);
    cudaEvent_t ev0, ev1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{ev0, ev1};
    SLRGPU_CUDA_TRY(cudaEventRecord(ev0, 0));
    const uint32_t grid = (numPixels + 127u) / 128u;
    if (sc->channels == 3) debugKernel<3><<<grid, 128>>>(sc->dev, rcs, (float*)bufs.p[0], (uint32_t*)bufs.p[1]);
    else debugKernel<16><<<grid, 128>>>(sc->dev, rcs, (float*)bufs.p[0], (uint32_t*)bufs.p[1]);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaEventRecord(ev1, 0));
    SLRGPU_CUDA_TRY(cudaMemcpy(out, bufs.p[0], bytes, cudaMemcpyDeviceToHost));
    uint32_t overflow = 0;
    SLRGPU_CUDA_TRY(cudaMemcpy(&overflow, bufs.p[1], sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->paths = numPixels; stats->rays = numPixels; stats->extend_rays = numPixels; stats->kernel_launches = 1;
        cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
    }
    if (overflow) { setError("traversal stack overflow (more than %d entries)", 64); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

SLRGPU_API int slrgpu_render(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* accum, SlrGpuRenderStats* stats) {
    int rc = checkRenderArgs(sc, p);
    if (rc) return rc;
    if (!accum) { setError("slrgpu_render: null accumulation buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    RenderWorkspace* w = nullptr;
    if ((rc = acquireWorkspace(sc, poolCapacity(p), &w))) return rc;
    struct Release { int device; RenderWorkspace* w; ~Release() { releaseWorkspace(device, w); } } release{sc->device, w};
    // the frame buffer and its pinned staging copy live in the pooled workspace: no allocation per call
    const size_t bytes = (size_t)p->width * p->height * sc->channels * sizeof(float);
    if ((rc = w->ensureFrame(bytes))) return rc;
    SLRGPU_CUDA_TRY(cudaMemsetAsync(w->frame, 0, bytes, w->stream));
    rc = renderImpl(sc, p, *w, w->frame, w->stream, stats);
    if (rc) return rc;
    return downloadFrame(*w, accum, bytes);
}

// ---------------------------------------------------------------------------------------------
// One frame on several GPUs of one box: the frame's sample range is partitioned over the scene
// replicas -- replica g renders [begin + g n / N, begin + (g + 1) n / N) of every pixel on its own device, one host thread
// per device -- and the float accumulation buffers are then summed onto the first replica's device by ONE kernel that
// reads the other devices' buffers directly over NVLink (peer access; a device pair without peer access goes through a
// staged peer copy), followed by one download. A path's random numbers are keyed by (pixel, global sample index), so the
// set of paths is the one a single GPU renders; only the fp32 summation order differs.
// ---------------------------------------------------------------------------------------------
SLRGPU_API int slrgpu_render_multi(SlrGpuScene* const* scenes, uint32_t numReplicas, const SlrGpuRenderParams* p, float* accum,
                                   SlrGpuRenderStats* stats) {
    if (!scenes || numReplicas == 0 || numReplicas > kMaxReplicas) { setError("slrgpu_render_multi: need 1..%u scene replicas", kMaxReplicas); return SLRGPU_ERR_INVALID_ARGUMENT; }
    for (uint32_t g = 0; g < numReplicas; ++g) {
        int rc = checkRenderArgs(scenes[g], p);
        if (rc) return rc;
        if (scenes[g]->channels != scenes[0]->channels) { setError("slrgpu_render_multi: replicas differ in channel count"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    }
    if (!accum) { setError("slrgpu_render_multi: null accumulation buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (numReplicas == 1) return slrgpu_render(scenes[0], p, accum, stats);

    const uint32_t n = p->spp_end - p->spp_begin;
    const size_t bytes = (size_t)p->width * p->height * scenes[0]->channels * sizeof(float);
    struct Replica {
        RenderWorkspace* w = nullptr;
        SlrGpuRenderParams params;
        SlrGpuRenderStats st;
        int rc = SLRGPU_OK;
        bool rendered = false;
        char error[256] = "";
    };
    std::vector<Replica> reps(numReplicas);
    struct ReleaseAll {
        std::vector<Replica>& r; SlrGpuScene* const* sc;
        ~ReleaseAll() { for (size_t g = 0; g < r.size(); ++g) if (r[g].w) { cudaSetDevice(sc[g]->device); releaseWorkspace(sc[g]->device, r[g].w); } }
    } releaseAll{reps, scenes};

    auto work = [&](uint32_t g) {
        Replica& r = reps[g];
        SlrGpuScene* sc = scenes[g];
        r.params = *p;
        r.params.spp_begin = p->spp_begin + (uint32_t)((unsigned long long)n * g / numReplicas);
        r.params.spp_end = p->spp_begin + (uint32_t)((unsigned long long)n * (g + 1) / numReplicas);
        memset(&r.st, 0, sizeof(r.st));
        cudaError_t e = cudaSetDevice(sc->device);
        if (e != cudaSuccess) { r.rc = cudaFail(e, "cudaSetDevice"); }
        // replica 0 always owns a (cleared) frame: it is the destination of the sum
        const bool empty = r.params.spp_end <= r.params.spp_begin;
        if (!r.rc && (!empty || g == 0)) {
            r.rc = acquireWorkspace(sc, poolCapacity(empty ? p : &r.params), &r.w);
            if (!r.rc) r.rc = r.w->ensureFrame(bytes);
            if (!r.rc && (e = cudaMemsetAsync(r.w->frame, 0, bytes, r.w->stream)) != cudaSuccess) r.rc = cudaFail(e, "cudaMemsetAsync(frame)");
            if (!r.rc && !empty) { r.rc = renderImpl(sc, &r.params, *r.w, r.w->frame, r.w->stream, &r.st); r.rendered = !r.rc; }
            if (!r.rc && (e = cudaStreamSynchronize(r.w->stream)) != cudaSuccess) r.rc = cudaFail(e, "cudaStreamSynchronize");
        }
        if (r.rc) snprintf(r.error, sizeof(r.error), "%s", slrgpu_last_error());       // the message is thread local
    };
    {
        std::vector<std::thread> pool;
        for (uint32_t g = 1; g < numReplicas; ++g) pool.emplace_back(work, g);
        work(0);
        for (std::thread& t : pool) t.join();
    }
    for (uint32_t g = 0; g < numReplicas; ++g)
        if (reps[g].rc) { setError("slrgpu_render_multi: replica %u (device %d): %s", g, scenes[g]->device, reps[g].error); return reps[g].rc; }

    // ---- the exchange step: frames of replicas 1.. summed into replica 0's frame, on replica 0's device
    const int dev0 = scenes[0]->device;
    SLRGPU_CUDA_TRY(cudaSetDevice(dev0));
    RenderWorkspace& w0 = *reps[0].w;
    PeerFrames peers;
    peers.count = 0;
    float* staging = nullptr;
    struct StagingFree { float** p; ~StagingFree() { if (*p) cudaFree(*p); } } stagingFree{&staging};
    const uint32_t count = (uint32_t)(bytes / sizeof(float));
    cudaEvent_t r0, r1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&r0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&r1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{r0, r1};
    SLRGPU_CUDA_TRY(cudaEventRecord(r0, w0.stream));
    for (uint32_t g = 1; g < numReplicas; ++g) {
        if (!reps[g].rendered) continue;
        const int dev = scenes[g]->device;
        bool direct = dev == dev0;
        if (!direct) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev0, dev) == cudaSuccess && can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(dev, 0);
                direct = pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled;
            }
            cudaGetLastError();
        }
        if (direct) peers.frame[peers.count++] = reps[g].w->frame;
        else {
            // no peer access between the pair: copy through the driver and add from local memory
            if (!staging) SLRGPU_CUDA_TRY(cudaMalloc(&staging, bytes));
            SLRGPU_CUDA_TRY(cudaMemcpyPeerAsync(staging, dev0, reps[g].w->frame, dev, bytes, w0.stream));
            PeerFrames one; one.count = 1; one.frame[0] = staging;
            sumFramesKernel<<<148 * 4, 256, 0, w0.stream>>>(w0.frame, one, count);
        }
    }
    if (peers.count) sumFramesKernel<<<148 * 4, 256, 0, w0.stream>>>(w0.frame, peers, count);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaEventRecord(r1, w0.stream));
    SLRGPU_CUDA_TRY(cudaStreamSynchronize(w0.stream));
    float reduceMs = 0.0f;
    cudaEventElapsedTime(&reduceMs, r0, r1);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (uint32_t g = 0; g < numReplicas; ++g) {
            const SlrGpuRenderStats& s = reps[g].st;
            stats->paths += s.paths; stats->rays += s.rays; stats->extend_rays += s.extend_rays; stats->shadow_rays += s.shadow_rays;
            stats->kernel_launches += s.kernel_launches; stats->waves = std::max(stats->waves, s.waves);
            stats->device_ms = std::max(stats->device_ms, s.device_ms);
            for (int c = 0; c < 9; ++c) stats->class_hits[c] += s.class_hits[c];
            stats->tail_paths += s.tail_paths; stats->tail_waves = std::max(stats->tail_waves, s.tail_waves);
        }
        stats->kernel_launches += peers.count ? 1 : 0;
        stats->device_ms += reduceMs;          // slowest replica + the exchange step
        stats->other_ms = reduceMs;
    }
    return downloadFrame(w0, accum, bytes);
}

}  // extern "C"
