// Ray-batch entry points: closest hit (Accelerator::intersect, libSLR/Core/Accelerator.h:17-34 /
// QBVH::intersect, QBVH.h:295-337) and occlusion (Scene::testVisibility, SurfaceObject.cpp:418-430).
// MUST be compiled with -fmad=false (see traverse.cuh).
#define SLR_WALK_ONE_RECORD_PER_STEP 1      // measured faster for ray batches (traverse.cuh walkStep)
#ifndef SLR_WALK_STEPS_PER_ROUND
#define SLR_WALK_STEPS_PER_ROUND 2         // two steps between refill checks: +0.8 % on the batch bench (a step here is a node OR a record)
#endif
#include "traverse.cuh"
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace slrgpu {

constexpr int kIntersectBlock = 128;

// Both kernels run the warp-cooperative walk of traverse.cuh (walkQueue): `status` points at two
// words, [0] = stack-overflow flag, [1] = the chunk cursor (must be 0 at launch).
struct BatchRaySource {
    SlrGpuRayBatch rays;
    __device__ __forceinline__ void load(uint32_t i, Ray& r) const {
        r.ox = rays.org_x[i]; r.oy = rays.org_y[i]; r.oz = rays.org_z[i];
        r.dx = rays.dir_x[i]; r.dy = rays.dir_y[i]; r.dz = rays.dir_z[i];
        r.tmin = rays.tmin[i]; r.tmax = rays.tmax[i];
    }
    __device__ __forceinline__ float time(uint32_t) const { return 0.0f; }      // ray batches carry no time: moving instances stand at their begin key frame
};
template <bool COUNT> struct BatchHitSink {
    SlrGpuHitBatch hits;
    __device__ __forceinline__ void done(uint32_t i, const WalkState& w, const TraversalCounters& cnt) const {
        hits.prim[i] = w.hit.prim;
        hits.inst[i] = w.hit.inst;
        hits.t[i] = w.hit.t;
        if (hits.u) hits.u[i] = w.hit.u;
        if (hits.v) hits.v[i] = w.hit.v;
        if (COUNT) {
            if (hits.nodes_visited) hits.nodes_visited[i] = cnt.nodes - w.cnt0.nodes;
            if (hits.tris_tested) hits.tris_tested[i] = cnt.tris - w.cnt0.tris;
        }
    }
};
struct OccludedSink {
    uint8_t* occluded;
    __device__ __forceinline__ void done(uint32_t i, const WalkState& w, const TraversalCounters&) const { occluded[i] = w.found ? 1 : 0; }
};

// the flat closest-hit kernel is bounded for 9 resident blocks (56 registers): it waits on dependent node fetches from a
// scene larger than L2, so a resident block more is worth more than the registers (C5 3838 Mrays/s at 62 registers / 8 blocks)
#ifndef SLR_INTERSECT_MIN_BLOCKS
#define SLR_INTERSECT_MIN_BLOCKS 9
#endif
template <bool INSTANCES, bool COUNT, bool ALPHA>
__global__ void __launch_bounds__(kIntersectBlock, (!INSTANCES && !COUNT && !ALPHA) ? SLR_INTERSECT_MIN_BLOCKS : 1)
intersectBatchKernel(const DeviceScene s, SlrGpuRayBatch rays, uint32_t n, SlrGpuHitBatch hits, int* status) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, false, COUNT, ALPHA>(s, n, reinterpret_cast<uint32_t*>(status + 1), BatchRaySource{rays}, BatchHitSink<COUNT>{hits}, cnt, overflow);
    if (overflow) atomicExch(status, 1);
}

template <bool INSTANCES, bool ALPHA>
__global__ void __launch_bounds__(kIntersectBlock)
occludedBatchKernel(const DeviceScene s, SlrGpuRayBatch rays, uint32_t n, uint8_t* occluded, int* status) {
    TraversalCounters cnt = {0, 0};
    bool overflow = false;
    walkQueue<INSTANCES, true, false, ALPHA>(s, n, reinterpret_cast<uint32_t*>(status + 1), BatchRaySource{rays}, OccludedSink{occluded}, cnt, overflow);
    if (overflow) atomicExch(status, 1);
}


// ---------------------------------------------------------------------------------------------
// The binary SBVH, traversed as the reference's shipped build does:
//   SBVH::intersect              libSLR/Accelerator/SBVH.h:417-442   (stack of node indices, near child on top, a node's box is
//                                                                     tested when it is POPPED, leaves tested in list order)
//   BoundingBox3D::intersect     libSLR/Core/geometry.h:112-126      (per axis: swap so that tNear <= tFar, tighten, early out)
// One thread per ray -- this entry point exists for parity with the default accelerator (tie rays resolve differently than in
// the QBVH), not for speed; the renderer and the benchmarks run the QBVH walk of traverse.cuh.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sbvhBoxTest(const SlrGpuSbvhNode& n, const Ray& r, float ix, float iy, float iz) {
    float dist0 = r.tmin, dist1 = r.tmax;
    float tn, tf;
#define SBVH_AXIS(LO, HI, O, I)                                              \
    tn = __fmul_rn(__fsub_rn(LO, O), I); tf = __fmul_rn(__fsub_rn(HI, O), I); \
    if (tn > tf) { const float sw = tn; tn = tf; tf = sw; }                   \
    dist0 = tn > dist0 ? tn : dist0;                                          \
    dist1 = tf < dist1 ? tf : dist1;                                          \
    if (dist0 > dist1) return false;
    SBVH_AXIS(n.lo[0], n.hi[0], r.ox, ix)
    SBVH_AXIS(n.lo[1], n.hi[1], r.oy, iy)
    SBVH_AXIS(n.lo[2], n.hi[2], r.oz, iz)
#undef SBVH_AXIS
    return true;
}

// LEVEL 0: the top-level aggregate (instance records descend into LEVEL 1); LEVEL 1: a nested aggregate
template <int LEVEL>
__device__ bool sbvhTraverse(const DeviceScene& s, uint32_t root, Ray& r, Hit& hit, bool& overflow) {
    const float ix = __frcp_rn(r.dx), iy = __frcp_rn(r.dy), iz = __frcp_rn(r.dz);       // Vector3D::reciprocal
    const bool dirPos[3] = {r.dx >= 0.0f, r.dy >= 0.0f, r.dz >= 0.0f};
    uint32_t stack[kStackSize];
    int depth = 0;
    stack[depth++] = root;
    bool found = false;
    while (depth > 0) {
        const SlrGpuSbvhNode n = s.sbvhNodes[stack[--depth]];
        if (!sbvhBoxTest(n, r, ix, iy, iz)) continue;
        if (!(n.b & 0x80000000u)) {
            if (depth + 2 > kStackSize) { overflow = true; continue; }
            const uint32_t c0 = n.a, c1 = n.b & 0x0FFFFFFFu;
            const bool positive = dirPos[(n.b >> 28) & 3u];
            stack[depth++] = positive ? c1 : c0;
            stack[depth++] = positive ? c0 : c1;
            continue;
        }
        const uint32_t count = n.b & 0x7FFFFFFFu;
        for (uint32_t i = 0; i < count; ++i) {
            const float4* rec = s.sbvhLeaves + (size_t)(n.a + i) * 3;
            const float4 a = ldg4(rec), b = ldg4(rec + 1), c = ldg4(rec + 2);
            const uint32_t id = __float_as_uint(a.w);
            if (id & 0x80000000u) {
                if constexpr (LEVEL == 0) {
                    // TransformedSurfaceObject::intersect (SurfaceObject.cpp:307-318): the ray in the instance's space, distMax carried
                    const SlrGpuInstance* inst = s.instances + (id & 0x7FFFFFFFu);
                    Ray lr = r;
                    mulPoint(inst->mat_inv, r.ox, r.oy, r.oz, &lr.ox, &lr.oy, &lr.oz);
                    mulVector(inst->mat_inv, r.dx, r.dy, r.dz, &lr.dx, &lr.dy, &lr.dz);
                    if (sbvhTraverse<1>(s, inst->sbvh_root_node, lr, hit, overflow)) { r.tmax = lr.tmax; hit.inst = id & 0x7FFFFFFFu; found = true; }
                }
                continue;
            }
            float t, b0, b1, b2;
            bool accept = triangleTest(a, b, c, r, &t, &b0, &b1, &b2);
            if (accept && (__float_as_uint(b.w) & SLRGPU_LEAF_FLAG_ALPHA_TEST)) accept = alphaTestPasses(s, id, b0, b1, b2);
            if (accept) { r.tmax = t; hit.prim = id; hit.inst = SLRGPU_INVALID_ID; hit.t = t; hit.u = b0; hit.v = b1; found = true; }
        }
    }
    return found;
}

__global__ void __launch_bounds__(kIntersectBlock)
intersectSbvhKernel(const DeviceScene s, SlrGpuRayBatch rays, uint32_t n, SlrGpuHitBatch hits, int* status) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Ray r;
        BatchRaySource{rays}.load(i, r);
        Hit hit;
        hit.prim = SLRGPU_INVALID_ID; hit.inst = SLRGPU_INVALID_ID; hit.t = INFINITY; hit.u = 0.0f; hit.v = 0.0f; hit.info = kSurfaceInfoMiss;
        bool overflow = false;
        sbvhTraverse<0>(s, 0u, r, hit, overflow);
        hits.prim[i] = hit.prim; hits.inst[i] = hit.inst; hits.t[i] = hit.t;
        if (hits.u) hits.u[i] = hit.u;
        if (hits.v) hits.v[i] = hit.v;
        if (overflow) atomicExch(status, 1);
    }
}

// grid of the warp-cooperative kernels: enough resident warps to fill the machine, no more than the batch needs
static uint32_t batchGrid(const SlrGpuScene* sc, uint64_t n) {
    int numSMs = 148;
    cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, sc->device);
    const uint64_t need = (n + kIntersectBlock - 1) / kIntersectBlock;
    const uint64_t full = (uint64_t)numSMs * 16u;
    return (uint32_t)(need < full ? need : full);
}

int launchIntersect(SlrGpuScene* sc, const SlrGpuRayBatch& rays, uint64_t n, const SlrGpuHitBatch& hits,
                           int* dStatus, cudaStream_t stream) {
    if (n == 0) return SLRGPU_OK;
    if (n >= 0xFFFF0000ull) { setError("ray batch too large (at most 2^32 - 65536 rays per call)"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaMemsetAsync(dStatus + 1, 0, sizeof(int), stream));
    const bool count = hits.nodes_visited || hits.tris_tested;
    const dim3 grid(batchGrid(sc, n)), block(kIntersectBlock);
    const uint32_t n32 = (uint32_t)n;
    if (sc->hasAlpha) {        // the general instantiation: instances and / or alpha-mapped triangles
        if (count) intersectBatchKernel<true, true, true><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
        else       intersectBatchKernel<true, false, true><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
    } else if (sc->hasInstances) {
        if (count) intersectBatchKernel<true, true, false><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
        else       intersectBatchKernel<true, false, false><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
    } else {
        if (count) intersectBatchKernel<false, true, false><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
        else       intersectBatchKernel<false, false, false><<<grid, block, 0, stream>>>(sc->dev, rays, n32, hits, dStatus);
    }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    return SLRGPU_OK;
}

// lazily allocated ring of status slots for slrgpu_intersect_batch_device (one slot per launch in flight)
static int sceneStatusRing(SlrGpuScene* sc, int** ring) {
    std::lock_guard<std::mutex> lock(sc->statusMutex);
    if (!sc->statusRing) {
        SLRGPU_CUDA_TRY(cudaMalloc(&sc->statusRing, 2 * kStatusSlots * sizeof(int)));
        SLRGPU_CUDA_TRY(cudaMemset(sc->statusRing, 0, 2 * kStatusSlots * sizeof(int)));
    }
    *ring = sc->statusRing;
    return SLRGPU_OK;
}

struct DeviceBuffers {
    void* ptrs[24] = {};
    int n = 0;
    ~DeviceBuffers() { for (int i = 0; i < n; ++i) cudaFree(ptrs[i]); }
    template <typename T> int alloc(T** p, uint64_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 4);
        if (e != cudaSuccess) return cudaFail(e, "cudaMalloc(batch buffer)");
        ptrs[n++] = q; *p = reinterpret_cast<T*>(q);
        return SLRGPU_OK;
    }
};

static int uploadRays(DeviceBuffers& bufs, const SlrGpuRayBatch* rays, uint64_t n, SlrGpuRayBatch* d) {
    const float* src[8] = {rays->org_x, rays->org_y, rays->org_z, rays->dir_x, rays->dir_y, rays->dir_z, rays->tmin, rays->tmax};
    float* dst[8];
    for (int k = 0; k < 8; ++k) {
        if (!src[k]) { setError("ray batch: null component array"); return SLRGPU_ERR_INVALID_ARGUMENT; }
        int rc = bufs.alloc(&dst[k], n);
        if (rc != SLRGPU_OK) return rc;
        SLRGPU_CUDA_TRY(cudaMemcpy(dst[k], src[k], n * sizeof(float), cudaMemcpyHostToDevice));
    }
    d->org_x = dst[0]; d->org_y = dst[1]; d->org_z = dst[2];
    d->dir_x = dst[3]; d->dir_y = dst[4]; d->dir_z = dst[5];
    d->tmin = dst[6]; d->tmax = dst[7];
    return SLRGPU_OK;
}


// ---------------------------------------------------------------------------------------------
// Host-buffer batches in pieces: pageable caller arrays -> pinned staging (several host threads) -> device -> kernel ->
// pinned staging -> caller arrays, three pieces in flight, so the host copies of piece k + 1 run while piece k is on
// the bus / in the kernel. The plain path (one cudaMemcpy per array, then the kernel, then the copies back) moved the
// 3 GB of the 64 Mi-ray batch at ~11 GB/s: 279 ms around a 30 ms kernel. Staging and device buffers are pooled per
// thread (slrgpu_release_workspaces frees nothing here; they live until the thread ends / the process exits).
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kPieceRays = 1ull << 21;
constexpr int kPiecesInFlight = 3;
constexpr int kMaxOutputs = 7;
constexpr int kMaxPipelineDevices = 16;

struct BatchPipeline {
    int device = -1;
    float* hIn[kPiecesInFlight] = {};        // pinned [8][kPieceRays]
    float* dIn[kPiecesInFlight] = {};
    uint32_t* hOut[kPiecesInFlight] = {};    // pinned [kMaxOutputs][kPieceRays]
    uint32_t* dOut[kPiecesInFlight] = {};
    int* dStatus[kPiecesInFlight] = {};
    cudaStream_t stream[kPiecesInFlight] = {};
    cudaEvent_t done[kPiecesInFlight] = {}, k0[kPiecesInFlight] = {}, k1[kPiecesInFlight] = {};
    int ensure(int dev) {
        if (device == dev) return SLRGPU_OK;
        for (int i = 0; i < kPiecesInFlight; ++i) {
            SLRGPU_CUDA_TRY(cudaMallocHost(&hIn[i], 8 * kPieceRays * sizeof(float)));
            SLRGPU_CUDA_TRY(cudaMallocHost(&hOut[i], kMaxOutputs * kPieceRays * sizeof(uint32_t)));
            SLRGPU_CUDA_TRY(cudaMalloc(&dIn[i], 8 * kPieceRays * sizeof(float)));
            SLRGPU_CUDA_TRY(cudaMalloc(&dOut[i], kMaxOutputs * kPieceRays * sizeof(uint32_t)));
            SLRGPU_CUDA_TRY(cudaMalloc(&dStatus[i], 2 * sizeof(int)));
            SLRGPU_CUDA_TRY(cudaMemset(dStatus[i], 0, 2 * sizeof(int)));
            SLRGPU_CUDA_TRY(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
            SLRGPU_CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            SLRGPU_CUDA_TRY(cudaEventCreate(&k0[i]));
            SLRGPU_CUDA_TRY(cudaEventCreate(&k1[i]));
        }
        device = dev;
        return SLRGPU_OK;
    }
};

// copies `count` arrays of `len` 4-byte elements each, array a from src[a] to dst[a], split over a few host threads
static void parallelCopy(void* const* dst, const void* const* src, int count, uint64_t len) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int threads = (int)std::min<unsigned>(8u, std::max(1u, hw / 2));
    const uint64_t total = (uint64_t)count * len;
    if (threads == 1 || total < (1u << 18)) { for (int a = 0; a < count; ++a) memcpy(dst[a], src[a], len * 4); return; }
    auto work = [&](int t) {
        // thread t copies the t-th slice of every array
        const uint64_t b = len * t / threads, e = len * (t + 1) / threads;
        for (int a = 0; a < count; ++a)
            memcpy(static_cast<char*>(dst[a]) + b * 4, static_cast<const char*>(src[a]) + b * 4, (e - b) * 4);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (std::thread& th : pool) th.join();
}

static int intersectBatchPipelined(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t n, const SlrGpuHitBatch* hits, float* kernel_ms) {
    static thread_local BatchPipeline pipes[kMaxPipelineDevices];      // one per device this thread has used
    BatchPipeline& pipe = pipes[scene->device];
    int rc = pipe.ensure(scene->device);
    if (rc != SLRGPU_OK) return rc;
    const void* src[8] = {rays->org_x, rays->org_y, rays->org_z, rays->dir_x, rays->dir_y, rays->dir_z, rays->tmin, rays->tmax};
    for (const void* p : src) if (!p) { setError("ray batch: null component array"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    // outputs in staging order: prim, inst, t, then the optional ones
    void* out[kMaxOutputs] = {hits->prim, hits->inst, hits->t, hits->u, hits->v, hits->nodes_visited, hits->tris_tested};
    const uint64_t pieces = (n + kPieceRays - 1) / kPieceRays;
    float msTotal = 0.0f;
    int overflow = 0;
    auto pieceLen = [&](uint64_t k) { return std::min(kPieceRays, n - k * kPieceRays); };
    // takes the results of piece k (slot k % kPiecesInFlight) back to the caller's arrays
    auto retire = [&](uint64_t k) -> int {
        const int sl = (int)(k % kPiecesInFlight);
        SLRGPU_CUDA_TRY(cudaEventSynchronize(pipe.done[sl]));
        const uint64_t len = pieceLen(k), off = k * kPieceRays;
        void* dst[kMaxOutputs]; const void* from[kMaxOutputs];
        int cnt = 0;
        for (int a = 0; a < kMaxOutputs; ++a)
            if (out[a]) { dst[cnt] = static_cast<uint32_t*>(out[a]) + off; from[cnt] = pipe.hOut[sl] + (uint64_t)a * kPieceRays; ++cnt; }
        parallelCopy(dst, from, cnt, len);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, pipe.k0[sl], pipe.k1[sl]);
        msTotal += ms;
        return SLRGPU_OK;
    };
    for (uint64_t k = 0; k < pieces; ++k) {
        const int sl = (int)(k % kPiecesInFlight);
        if (k >= (uint64_t)kPiecesInFlight && (rc = retire(k - kPiecesInFlight))) return rc;
        const uint64_t len = pieceLen(k), off = k * kPieceRays;
        void* dst[8]; const void* from[8];
        for (int a = 0; a < 8; ++a) { dst[a] = pipe.hIn[sl] + (uint64_t)a * kPieceRays; from[a] = static_cast<const float*>(src[a]) + off; }
        parallelCopy(dst, from, 8, len);
        cudaStream_t st = pipe.stream[sl];
        if (len == kPieceRays) SLRGPU_CUDA_TRY(cudaMemcpyAsync(pipe.dIn[sl], pipe.hIn[sl], 8 * kPieceRays * sizeof(float), cudaMemcpyHostToDevice, st));
        else for (int a = 0; a < 8; ++a)
            SLRGPU_CUDA_TRY(cudaMemcpyAsync(pipe.dIn[sl] + (uint64_t)a * kPieceRays, pipe.hIn[sl] + (uint64_t)a * kPieceRays, len * sizeof(float), cudaMemcpyHostToDevice, st));
        float* d = pipe.dIn[sl];
        const SlrGpuRayBatch dr = {d, d + kPieceRays, d + 2 * kPieceRays, d + 3 * kPieceRays, d + 4 * kPieceRays, d + 5 * kPieceRays, d + 6 * kPieceRays, d + 7 * kPieceRays};
        uint32_t* o = pipe.dOut[sl];
        SlrGpuHitBatch dh = {};
        dh.prim = o; dh.inst = o + kPieceRays; dh.t = reinterpret_cast<float*>(o + 2 * kPieceRays);
        if (hits->u) dh.u = reinterpret_cast<float*>(o + 3 * kPieceRays);
        if (hits->v) dh.v = reinterpret_cast<float*>(o + 4 * kPieceRays);
        if (hits->nodes_visited) dh.nodes_visited = o + 5 * kPieceRays;
        if (hits->tris_tested) dh.tris_tested = o + 6 * kPieceRays;
        SLRGPU_CUDA_TRY(cudaEventRecord(pipe.k0[sl], st));
        if ((rc = launchIntersect(scene, dr, len, dh, pipe.dStatus[sl], st))) return rc;
        SLRGPU_CUDA_TRY(cudaEventRecord(pipe.k1[sl], st));
        for (int a = 0; a < kMaxOutputs; ++a)
            if (out[a]) SLRGPU_CUDA_TRY(cudaMemcpyAsync(pipe.hOut[sl] + (uint64_t)a * kPieceRays, o + (uint64_t)a * kPieceRays, len * 4, cudaMemcpyDeviceToHost, st));
        SLRGPU_CUDA_TRY(cudaEventRecord(pipe.done[sl], st));
    }
    for (uint64_t k = pieces > (uint64_t)kPiecesInFlight ? pieces - kPiecesInFlight : 0; k < pieces; ++k)
        if ((rc = retire(k))) return rc;
    for (int sl = 0; sl < kPiecesInFlight; ++sl) {
        int status = 0;
        SLRGPU_CUDA_TRY(cudaMemcpy(&status, pipe.dStatus[sl], sizeof(int), cudaMemcpyDeviceToHost));
        if (status) { overflow = 1; cudaMemset(pipe.dStatus[sl], 0, sizeof(int)); }
    }
    if (kernel_ms) *kernel_ms = msTotal;
    if (overflow) { setError("traversal stack overflow (more than %d pending nodes)", kStackSize); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

}  // namespace slrgpu

using namespace slrgpu;

extern "C" {

SLRGPU_API int slrgpu_intersect_launch_config(SlrGpuScene* scene, uint64_t num_rays, uint32_t* grid, uint32_t* block) {
    if (!scene) { setError("null scene"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (grid) *grid = batchGrid(scene, num_rays);
    if (block) *block = kIntersectBlock;
    return SLRGPU_OK;
}

SLRGPU_API int slrgpu_intersect_batch_device(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t num_rays,
                                             const SlrGpuHitBatch* hits, void* stream) {
    if (!scene || !rays || !hits || !hits->prim || !hits->inst || !hits->t) {
        setError("slrgpu_intersect_batch_device: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT;
    }
    SLRGPU_CUDA_TRY(cudaSetDevice(scene->device));
    // status words (overflow flag, chunk cursor) of this launch: the next slot of the scene's ring, on the scene's device.
    // The cursor is zeroed on `stream` before the kernel; the overflow flag is sticky until slrgpu_scene_poll_overflow.
    int* ring = nullptr;
    int rc = sceneStatusRing(scene, &ring);
    if (rc != SLRGPU_OK) return rc;
    const uint32_t slot = scene->statusNext.fetch_add(1u) % kStatusSlots;
    return launchIntersect(scene, *rays, num_rays, *hits, ring + 2 * slot, (cudaStream_t)stream);
}

SLRGPU_API int slrgpu_scene_poll_overflow(SlrGpuScene* scene, int* overflow) {
    if (!scene || !overflow) { setError("slrgpu_scene_poll_overflow: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    *overflow = 0;
    if (!scene->statusRing) return SLRGPU_OK;
    SLRGPU_CUDA_TRY(cudaSetDevice(scene->device));
    SLRGPU_CUDA_TRY(cudaDeviceSynchronize());
    int host[2 * kStatusSlots];
    SLRGPU_CUDA_TRY(cudaMemcpy(host, scene->statusRing, sizeof(host), cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < kStatusSlots; ++k) if (host[2 * k]) *overflow = 1;
    if (*overflow) SLRGPU_CUDA_TRY(cudaMemset(scene->statusRing, 0, sizeof(host)));
    return SLRGPU_OK;
}

SLRGPU_API int slrgpu_intersect_batch(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t n,
                                      const SlrGpuHitBatch* hits, float* kernel_ms) {
    if (!scene || !rays || !hits || !hits->prim || !hits->inst || !hits->t) {
        setError("slrgpu_intersect_batch: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT;
    }
    if (kernel_ms) *kernel_ms = 0.0f;
    if (n == 0) return SLRGPU_OK;
    SLRGPU_CUDA_TRY(cudaSetDevice(scene->device));
    if (n >= 2 * kPieceRays && scene->device >= 0 && scene->device < kMaxPipelineDevices) return intersectBatchPipelined(scene, rays, n, hits, kernel_ms);
    DeviceBuffers bufs;
    SlrGpuRayBatch dr;
    int rc = uploadRays(bufs, rays, n, &dr);
    if (rc != SLRGPU_OK) return rc;
    SlrGpuHitBatch dh = {};
    int* dStatus = nullptr;
    if ((rc = bufs.alloc(&dh.prim, n)) || (rc = bufs.alloc(&dh.inst, n)) || (rc = bufs.alloc(&dh.t, n)) ||
        (rc = bufs.alloc(&dStatus, 2))) return rc;
    if (hits->u && (rc = bufs.alloc(&dh.u, n))) return rc;
    if (hits->v && (rc = bufs.alloc(&dh.v, n))) return rc;
    if (hits->nodes_visited && (rc = bufs.alloc(&dh.nodes_visited, n))) return rc;
    if (hits->tris_tested && (rc = bufs.alloc(&dh.tris_tested, n))) return rc;
    SLRGPU_CUDA_TRY(cudaMemset(dStatus, 0, 2 * sizeof(int)));
    cudaEvent_t e0, e1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&e0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&e1));
    SLRGPU_CUDA_TRY(cudaEventRecord(e0, 0));
    rc = launchIntersect(scene, dr, n, dh, dStatus, 0);
    cudaEventRecord(e1, 0);
    cudaError_t se = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != SLRGPU_OK) return rc;
    SLRGPU_CUDA_TRY(se);
    if (kernel_ms) *kernel_ms = ms;
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->prim, dh.prim, n * 4, cudaMemcpyDeviceToHost));
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->inst, dh.inst, n * 4, cudaMemcpyDeviceToHost));
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->t, dh.t, n * 4, cudaMemcpyDeviceToHost));
    if (hits->u) SLRGPU_CUDA_TRY(cudaMemcpy(hits->u, dh.u, n * 4, cudaMemcpyDeviceToHost));
    if (hits->v) SLRGPU_CUDA_TRY(cudaMemcpy(hits->v, dh.v, n * 4, cudaMemcpyDeviceToHost));
    if (hits->nodes_visited) SLRGPU_CUDA_TRY(cudaMemcpy(hits->nodes_visited, dh.nodes_visited, n * 4, cudaMemcpyDeviceToHost));
    if (hits->tris_tested) SLRGPU_CUDA_TRY(cudaMemcpy(hits->tris_tested, dh.tris_tested, n * 4, cudaMemcpyDeviceToHost));
    int status = 0;
    SLRGPU_CUDA_TRY(cudaMemcpy(&status, dStatus, sizeof(int), cudaMemcpyDeviceToHost));
    if (status) { setError("traversal stack overflow (more than %d pending nodes)", kStackSize); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

SLRGPU_API int slrgpu_intersect_batch_sbvh(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t n, const SlrGpuHitBatch* hits, float* kernel_ms) {
    if (!scene || !rays || !hits || !hits->prim || !hits->inst || !hits->t) { setError("slrgpu_intersect_batch_sbvh: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (!scene->dev.sbvhNodes) { setError("slrgpu_intersect_batch_sbvh: the scene was created without sbvh_nodes"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (kernel_ms) *kernel_ms = 0.0f;
    if (n == 0) return SLRGPU_OK;
    if (n >= 0xFFFF0000ull) { setError("ray batch too large (at most 2^32 - 65536 rays per call)"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(scene->device));
    DeviceBuffers bufs;
    SlrGpuRayBatch dr;
    int rc = uploadRays(bufs, rays, n, &dr);
    if (rc != SLRGPU_OK) return rc;
    SlrGpuHitBatch dh = {};
    int* dStatus = nullptr;
    if ((rc = bufs.alloc(&dh.prim, n)) || (rc = bufs.alloc(&dh.inst, n)) || (rc = bufs.alloc(&dh.t, n)) || (rc = bufs.alloc(&dStatus, 2))) return rc;
    if (hits->u && (rc = bufs.alloc(&dh.u, n))) return rc;
    if (hits->v && (rc = bufs.alloc(&dh.v, n))) return rc;
    SLRGPU_CUDA_TRY(cudaMemset(dStatus, 0, 2 * sizeof(int)));
    cudaEvent_t e0, e1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&e0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&e1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{e0, e1};
    SLRGPU_CUDA_TRY(cudaEventRecord(e0, 0));
    intersectSbvhKernel<<<batchGrid(scene, n), kIntersectBlock>>>(scene->dev, dr, (uint32_t)n, dh, dStatus);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaEventRecord(e1, 0));
    SLRGPU_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->prim, dh.prim, n * 4, cudaMemcpyDeviceToHost));
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->inst, dh.inst, n * 4, cudaMemcpyDeviceToHost));
    SLRGPU_CUDA_TRY(cudaMemcpy(hits->t, dh.t, n * 4, cudaMemcpyDeviceToHost));
    if (hits->u) SLRGPU_CUDA_TRY(cudaMemcpy(hits->u, dh.u, n * 4, cudaMemcpyDeviceToHost));
    if (hits->v) SLRGPU_CUDA_TRY(cudaMemcpy(hits->v, dh.v, n * 4, cudaMemcpyDeviceToHost));
    int status = 0;
    SLRGPU_CUDA_TRY(cudaMemcpy(&status, dStatus, sizeof(int), cudaMemcpyDeviceToHost));
    if (status) { setError("traversal stack overflow (more than %d pending nodes)", kStackSize); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

SLRGPU_API int slrgpu_occluded_batch(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t n,
                                     uint8_t* occluded, float* kernel_ms) {
    if (!scene || !rays || !occluded) { setError("slrgpu_occluded_batch: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (kernel_ms) *kernel_ms = 0.0f;
    if (n == 0) return SLRGPU_OK;
    SLRGPU_CUDA_TRY(cudaSetDevice(scene->device));
    DeviceBuffers bufs;
    SlrGpuRayBatch dr;
    int rc = uploadRays(bufs, rays, n, &dr);
    if (rc != SLRGPU_OK) return rc;
    uint8_t* dOcc = nullptr; int* dStatus = nullptr;
    if ((rc = bufs.alloc(&dOcc, n)) || (rc = bufs.alloc(&dStatus, 2))) return rc;
    SLRGPU_CUDA_TRY(cudaMemset(dStatus, 0, 2 * sizeof(int)));
    if (n >= 0xFFFF0000ull) { setError("ray batch too large (at most 2^32 - 65536 rays per call)"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    const uint32_t blocks = batchGrid(scene, n);
    cudaEvent_t e0, e1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&e0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&e1));
    cudaEventRecord(e0, 0);
    if (scene->hasAlpha)          occludedBatchKernel<true, true><<<blocks, kIntersectBlock>>>(scene->dev, dr, (uint32_t)n, dOcc, dStatus);
    else if (scene->hasInstances) occludedBatchKernel<true, false><<<blocks, kIntersectBlock>>>(scene->dev, dr, (uint32_t)n, dOcc, dStatus);
    else                          occludedBatchKernel<false, false><<<blocks, kIntersectBlock>>>(scene->dev, dr, (uint32_t)n, dOcc, dStatus);
    cudaError_t le = cudaGetLastError();
    cudaEventRecord(e1, 0);
    cudaError_t se = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    SLRGPU_CUDA_TRY(le);
    SLRGPU_CUDA_TRY(se);
    if (kernel_ms) *kernel_ms = ms;
    SLRGPU_CUDA_TRY(cudaMemcpy(occluded, dOcc, n, cudaMemcpyDeviceToHost));
    int status = 0;
    SLRGPU_CUDA_TRY(cudaMemcpy(&status, dStatus, sizeof(int), cudaMemcpyDeviceToHost));
    if (status) { setError("traversal stack overflow"); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

}  // extern "C"
