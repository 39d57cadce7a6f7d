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
#include "rng.cuh"
#include "shade.cuh"
#include "wavefront.cuh"
#include <chrono>
#include <cstring>
#include <new>
#include <vector>

namespace slrgpu {

constexpr int kShadeBlock = 128;
constexpr int kRaygenBlock = 256;

// position of `alive` lanes in an output queue: one atomic per warp
__device__ __forceinline__ uint32_t warpAppend(bool alive, uint32_t* counter) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, alive);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

template <int NC> __device__ __forceinline__ void storeAlpha(const PathQueue& q, uint32_t pos, const Spec<NC>& a) {
    if (NC == 3) { q.alpha[pos] = make_float4(a.v[0], a.v[1], a.v[2], 0.0f); return; }
#pragma unroll
    for (int k = 0; k < NC / 4; ++k)
        q.alpha[(size_t)k * q.capacity + pos] = make_float4(a.v[4 * k], a.v[(4 * k + 1) % NC], a.v[(4 * k + 2) % NC], a.v[(4 * k + 3) % NC]);
}
template <int NC> __device__ __forceinline__ Spec<NC> loadAlpha(const PathQueue& q, uint32_t pos) {
    Spec<NC> a;
    if (NC == 3) { const float4 v = q.alpha[pos]; a.v[0] = v.x; a.v[1] = v.y; a.v[2] = v.z; return a; }
#pragma unroll
    for (int k = 0; k < NC / 4; ++k) {
        const float4 v = q.alpha[(size_t)k * q.capacity + pos];
        a.v[4 * k] = v.x; a.v[(4 * k + 1) % NC] = v.y; a.v[(4 * k + 2) % NC] = v.z; a.v[(4 * k + 3) % NC] = v.w;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------
// ray generation: camera samples [first, first + count) of this render call, appended to `out`
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kRaygenBlock)
raygenKernel(const DeviceScene s, const RenderConstants rc, unsigned long long first, uint32_t count, PathQueue out, uint32_t outBase) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const unsigned long long g = first + j;
    const uint32_t pass = (uint32_t)(g / rc.numPixels);
    const uint32_t r = (uint32_t)(g % rc.numPixels);
    // pixel order: bands of 8 rows, column-major inside a band, so a warp covers a 4x8 pixel block
    const uint32_t band = r / (8u * rc.width);
    const uint32_t local = r - band * 8u * rc.width;
    const uint32_t rows = min(8u, rc.height - band * 8u);
    const uint32_t x = local / rows, y = band * 8u + local % rows;
    const uint32_t pixel = y * rc.width + x;
    const uint32_t sample = rc.sppBegin + pass;

    const Rand4 r0 = pathRandom(rc.seed, pixel, sample, 0);    // time, pixel x, pixel y, wavelength offset
    const Rand4 r1 = pathRandom(rc.seed, pixel, sample, 1);    // wavelength selection, lens u0, lens u1
    const float px = x + r0.y, py = y + r0.z;
    const float wlOffset = r0.w;
    const uint32_t hero = min((uint32_t)(NC * r1.x), (uint32_t)(NC - 1));

    // PerspectiveCamera::sample
    float lx, ly;
    concentricSampleDisk(r1.y, r1.z, &lx, &ly);
    const SlrGpuCamera& cam = s.camera;
    const V3 orgLocal(cam.lens_radius * lx, cam.lens_radius * ly, 0.0f);
    const V3 org = xfmPoint(cam.mat, orgLocal);
    const V3 lensN = xfmNormal(cam.mat_inv, V3(0, 0, 1));
    Frame f;
    f.z = lensN;
    f.x = xfmVector(cam.mat, V3(1, 0, 0));
    f.y = cross(f.z, f.x);
    // PerspectiveIDF::sample with (p.x / W, p.y / H)
    const V3 pFocus(rc.opWidth * (0.5f - px / rc.width), rc.opHeight * (0.5f - py / rc.height), cam.obj_plane_dist);
    const V3 dirLocal = normalize(pFocus - orgLocal);
    const float dirPDF = cam.img_plane_dist * cam.img_plane_dist / ((dirLocal.z * dirLocal.z * dirLocal.z) * rc.imgPlaneArea);
    const V3 dir = f.fromLocal(dirLocal);
    const float weight = absDot(dir, lensN) / (rc.lensAreaPDF * dirPDF * rc.selectWLPDF);

    // ImageSensor::add bins by the float pixel position
    const uint32_t ipx = min((uint32_t)px, rc.width - 1), ipy = min((uint32_t)py, rc.height - 1);
    uint32_t flags = kFlagCameraRay;
    if (NC == 16 && strataInPlace(wlOffset)) flags |= kFlagStrataInPlace;

    const uint32_t pos = outBase + j;
    out.org[pos] = make_float4(org.x, org.y, org.z, 0.0f);
    out.dir[pos] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    out.meta[pos] = make_uint4(ipy * rc.width + ipx, sample, hero | (flags << 8), __float_as_uint(wlOffset));
    out.weight[pos] = weight * rc.recBinWidth;
    storeAlpha<NC>(out, pos, specConst<NC>(1.0f));
}

// ---------------------------------------------------------------------------------------------
// shade: one bounce of Job::contribution for every path of the queue
// ---------------------------------------------------------------------------------------------
template <int NC, int ML>
__global__ void __launch_bounds__(kShadeBlock)
shadeKernel(const DeviceScene s, const RenderConstants rc, PathQueue in, uint32_t n, HitBuffer hits, PathQueue out, ShadowQueue sq,
            float* __restrict__ accum, WavefrontCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false, shadow = false;
    // outputs of the bounce
    V3 nOrg(0, 0, 0), nDir(0, 0, 1);
    float nPdf = 0.0f;
    uint4 meta = make_uint4(0, 0, 0, 0);
    float weight = 0.0f;
    Spec<NC> alpha = specConst<NC>(0.0f);
    V3 sOrg(0, 0, 0), sDir(0, 0, 1);
    float sTmax = 0.0f;
    Spec<NC> sContrib = specConst<NC>(0.0f);

    if (i < n) {
        const float4 o4 = in.org[i], d4 = in.dir[i];
        meta = in.meta[i];
        weight = in.weight[i];
        alpha = loadAlpha<NC>(in, i);
        const uint2 hid = hits.id[i];
        const float4 htuv = hits.tuv[i];
        const V3 org(o4.x, o4.y, o4.z), dir(d4.x, d4.y, d4.z);
        const float prevPdf = d4.w;
        const uint32_t hero = meta.z & 0xFFu;
        uint32_t flags = (meta.z >> 8) & 0xFFu;
        uint32_t pathLength = meta.z >> 16;
        const float wlOffset = __uint_as_float(meta.w);
        const bool cameraRay = flags & kFlagCameraRay;
        const bool inPlace = flags & kFlagStrataInPlace;

        SurfPt sp;
        bool hit = true, emitting = false;
        uint32_t material = SLRGPU_INVALID_ID;
        SlrGpuTriangle tri = {};
        float localArea = 1.0f;
        if (hid.x != SLRGPU_INVALID_ID) {
            tri = hitSurfacePoint(s, hid.x, hid.y, htuv.x, htuv.y, htuv.z, org, dir, &sp, &localArea);
            material = tri.material;
            emitting = materialIsEmitting(s, material);
        } else if (s.envPresent) {
            envSurfacePoint(dir, &sp);
            material = s.envMaterial;
            emitting = true;
        } else {
            hit = false;
        }

        if (hit) {
            V3 dirOut = sp.sf.toLocal(-dir);
            bool cont = true;
            if (emitting) {
                // DiffuseEDF: 1/pi on the front side; IBLEDF: 1/pi
                const float edf = (sp.atInfinity || dirOut.z > 0.0f) ? 1.0f / kPi : 0.0f;
                float mis = 1.0f;
                if (!cameraRay && !(flags & kFlagPrevDelta)) {
                    const float lightProb = lightSelectionProb(s, tri, hid.y, sp.atInfinity);
                    float areaPDF, dist2;
                    if (sp.atInfinity) { areaPDF = envEvaluateUVPDF(s, sp.u / (2 * kPi), sp.v / kPi) / (2 * kPi * kPi * sinf(sp.v)); dist2 = 1.0f; }
                    else { areaPDF = 1.0f / localArea; dist2 = sqLength(sp.p - org); }
                    const float lightPDF = lightProb * areaPDF * dist2 / absDot(dir, sp.gn);
                    mis = (prevPdf * prevPdf) / (lightPDF * lightPDF + prevPdf * prevPdf);
                }
                if (edf > 0.0f) {
                    const Spec<NC> Le = materialEmittance<NC>(s, material, sp, wlOffset);
                    float v[NC == 3 ? 4 : NC];
                    const float k = edf * mis * weight;
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[c] = alpha.v[c] * Le.v[c] * k;
                    splat<NC>(accum, meta.x, wlOffset, inPlace, v);
                }
            }
            if (sp.atInfinity) cont = false;
            if (cont && !cameraRay) {
                // Russian roulette; initY = importance of a unit spectrum = 1
                const float continueProb = fminf(specImportance(alpha, hero) / 1.0f, 1.0f);
                const Rand4 rr = pathRandom(rc.seed, meta.x, meta.y, 2 * pathLength + 1);   // .z of the previous bounce's second block
                if (rr.z < continueProb) alpha = alpha * (1.0f / continueProb);
                else cont = false;
            }
            if (cont) {
                ++pathLength;
                if (pathLength >= rc.maxPathLength) cont = false;
            }
            if (cont) {
                const V3 gNorm = sp.sf.toLocal(sp.gn);
                Bsdf<NC, ML> bsdf;
                buildBsdf<NC, ML>(s, material, sp, wlOffset, (flags & kFlagLambdaSelected) != 0, &bsdf);
                BsdfQuery q;
                q.dir = dirOut; q.gn = gNorm; q.hero = hero; q.flags = DT_All;
                const Rand4 ra = pathRandom(rc.seed, meta.x, meta.y, 2 * pathLength);       // light select, light u0, u1, bsdf component
                const Rand4 rb = pathRandom(rc.seed, meta.x, meta.y, 2 * pathLength + 1);   // bsdf u0, u1

                // next event estimation
                if (bsdfHasNonDelta(bsdf) && (s.numTopLights > 0 || s.envPresent)) {
                    LightSample ls;
                    sampleLight(s, ra.x, ra.y, ra.z, &ls);
                    float dist2;
                    V3 shadowDir;
                    if (ls.sp.atInfinity) { dist2 = 1.0f; shadowDir = normalize(ls.sp.p); }
                    else { const V3 d = ls.sp.p - sp.p; dist2 = sqLength(d); shadowDir = d / sqrtf(dist2); }
                    const V3 shadowDir_l = ls.sp.sf.toLocal(-shadowDir);
                    const V3 shadowDir_sn = sp.sf.toLocal(shadowDir);
                    const float edf = (ls.isEnv || shadowDir_l.z > 0.0f) ? 1.0f / kPi : 0.0f;
                    if (edf > 0.0f && ls.areaPDF > 0.0f) {
                        const Spec<NC> fs = bsdfEvaluate(bsdf, q, shadowDir_sn);
                        if (!specIsZero(fs)) {
                            const Spec<NC> M = materialEmittance<NC>(s, ls.material, ls.sp, wlOffset);
                            const float cosLight = absDot(-shadowDir, ls.sp.gn);
                            const float bsdfPDF = bsdfPdf(bsdf, q, shadowDir_sn) * cosLight / dist2;
                            float mis = 1.0f;
                            if (!isinf(ls.areaPDF)) mis = (ls.lightPDF * ls.lightPDF) / (ls.lightPDF * ls.lightPDF + bsdfPDF * bsdfPDF);
                            const float G = absDot(shadowDir_sn, gNorm) * cosLight / dist2;
                            const float k = edf * (G * mis / ls.lightPDF) * weight;
#pragma unroll
                            for (int c = 0; c < NC; ++c) sContrib.v[c] = alpha.v[c] * M.v[c] * fs.v[c] * k;
                            // Scene::testVisibility
                            sOrg = sp.p;
                            if (ls.sp.atInfinity) { sDir = shadowDir; sTmax = 3.402823466e+38f; }
                            else { const float dist = length(ls.sp.p - sp.p); sDir = (ls.sp.p - sp.p) / dist; sTmax = dist * (1.0f - 0.0001f); }
                            shadow = true;
                        }
                    }
                }

                // sample the BSDF for the next direction
                BsdfSampleResult res;
                const Spec<NC> fs = bsdfSample(bsdf, q, ra.w, rb.x, rb.y, &res);
                if (!specIsZero(fs) && res.pdf != 0.0f) {
                    float dirPDF = res.pdf;
                    if (res.type & DT_Dispersive) { dirPDF /= NC; flags |= kFlagLambdaSelected; }
                    const float k = absDot(res.dir, gNorm) / dirPDF;
                    alpha = alpha * (fs * k);
                    nOrg = sp.p;
                    nDir = sp.sf.fromLocal(res.dir);
                    nPdf = dirPDF;
                    flags &= ~(kFlagCameraRay | kFlagPrevDelta);
                    if (dtIsDelta(res.type)) flags |= kFlagPrevDelta;
                    meta.z = hero | (flags << 8) | (pathLength << 16);
                    alive = true;
                }
            }
        }
    }

    const uint32_t spos = warpAppend(shadow, &counters->numShadow);
    if (shadow) {
        sq.org[spos] = make_float4(sOrg.x, sOrg.y, sOrg.z, 0.0001f);
        sq.dir[spos] = make_float4(sDir.x, sDir.y, sDir.z, sTmax);
        const bool inPlace = ((meta.z >> 8) & kFlagStrataInPlace) != 0;
        sq.pixelWl[spos] = make_uint2(meta.x | (inPlace ? 0x80000000u : 0u), meta.w);
        if (NC == 3) sq.contrib[spos] = make_float4(sContrib.v[0], sContrib.v[1], sContrib.v[2], 0.0f);
        else {
#pragma unroll
            for (int k = 0; k < NC / 4; ++k)
                sq.contrib[(size_t)k * sq.capacity + spos] =
                    make_float4(sContrib.v[4 * k], sContrib.v[(4 * k + 1) % NC], sContrib.v[(4 * k + 2) % NC], sContrib.v[(4 * k + 3) % NC]);
        }
    }
    const uint32_t npos = warpAppend(alive, &counters->numNext);
    if (alive) {
        out.org[npos] = make_float4(nOrg.x, nOrg.y, nOrg.z, 0.0001f);      // Ray::Epsilon
        out.dir[npos] = make_float4(nDir.x, nDir.y, nDir.z, nPdf);
        out.meta[npos] = meta;
        out.weight[npos] = weight;
        storeAlpha<NC>(out, npos, alpha);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct RenderBuffers {
    void* ptrs[40] = {};
    int n = 0;
    ~RenderBuffers() { for (int i = 0; i < n; ++i) cudaFree(ptrs[i]); }
    template <typename T> int alloc(T** p, uint64_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
        if (e != cudaSuccess) return cudaFail(e, "cudaMalloc(render buffers)");
        ptrs[n++] = q;
        *p = reinterpret_cast<T*>(q);
        return SLRGPU_OK;
    }
};

static int allocPathQueue(RenderBuffers& b, PathQueue* q, uint32_t P, int quarters) {
    int rc;
    if ((rc = b.alloc(&q->org, P))) return rc;
    if ((rc = b.alloc(&q->dir, P))) return rc;
    if ((rc = b.alloc(&q->meta, P))) return rc;
    if ((rc = b.alloc(&q->weight, P))) return rc;
    if ((rc = b.alloc(&q->alpha, (uint64_t)P * quarters))) return rc;
    q->capacity = P;
    return SLRGPU_OK;
}

template <int NC>
static void launchRaygen(const SlrGpuScene* sc, const RenderConstants& rc, unsigned long long first, uint32_t count, const PathQueue& out,
                         uint32_t outBase, cudaStream_t stream) {
    if (count == 0) return;
    raygenKernel<NC><<<(count + kRaygenBlock - 1) / kRaygenBlock, kRaygenBlock, 0, stream>>>(sc->dev, rc, first, count, out, outBase);
}

template <int NC>
static void launchShade(const SlrGpuScene* sc, const RenderConstants& rc, const PathQueue& in, uint32_t n, const HitBuffer& hits,
                        const PathQueue& out, const ShadowQueue& sq, float* accum, WavefrontCounters* counters, cudaStream_t stream) {
    if (n == 0) return;
    const dim3 grid((n + kShadeBlock - 1) / kShadeBlock), block(kShadeBlock);
    if (sc->maxLobes <= 1) shadeKernel<NC, 1><<<grid, block, 0, stream>>>(sc->dev, rc, in, n, hits, out, sq, accum, counters);
    else shadeKernel<NC, 4><<<grid, block, 0, stream>>>(sc->dev, rc, in, n, hits, out, sq, accum, counters);
}

static int renderImpl(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* accumDev, cudaStream_t stream, SlrGpuRenderStats* stats) {
    const bool rgb = sc->channels == 3;
    const uint32_t W = p->width, H = p->height;
    const unsigned long long numPixels = (unsigned long long)W * H;
    const unsigned long long totalSamples = numPixels * (p->spp_end - p->spp_begin);
    uint32_t P = p->pool_size ? p->pool_size : (1u << 21);
    if ((unsigned long long)P > totalSamples) P = (uint32_t)totalSamples;
    P = (P + 127u) & ~127u;
    if (P == 0) P = 128;

    RenderConstants rc;
    memset(&rc, 0, sizeof(rc));
    rc.width = W; rc.height = H; rc.numPixels = (uint32_t)numPixels;
    rc.sppBegin = p->spp_begin;
    rc.capacity = P;
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

    RenderBuffers bufs;
    PathQueue q[2];
    HitBuffer hits;
    ShadowQueue sq;
    WavefrontCounters* dCounters = nullptr;
    const int quarters = rgb ? 1 : 4;
    int rcode;
    for (int k = 0; k < 2; ++k) if ((rcode = allocPathQueue(bufs, &q[k], P, quarters))) return rcode;
    if ((rcode = bufs.alloc(&hits.id, P))) return rcode;
    if ((rcode = bufs.alloc(&hits.tuv, P))) return rcode;
    if ((rcode = bufs.alloc(&sq.org, P))) return rcode;
    if ((rcode = bufs.alloc(&sq.dir, P))) return rcode;
    if ((rcode = bufs.alloc(&sq.pixelWl, P))) return rcode;
    if ((rcode = bufs.alloc(&sq.contrib, (uint64_t)P * quarters))) return rcode;
    sq.capacity = P;
    if ((rcode = bufs.alloc(&dCounters, 1))) return rcode;
    SLRGPU_CUDA_TRY(cudaMemsetAsync(dCounters, 0, sizeof(WavefrontCounters), stream));
    WavefrontCounters* hCounters = nullptr;
    SLRGPU_CUDA_TRY(cudaMallocHost(&hCounters, sizeof(WavefrontCounters)));
    struct PinnedFree { void* p; ~PinnedFree() { cudaFreeHost(p); } } pinnedFree{hCounters};

    cudaEvent_t ev0, ev1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{ev0, ev1};
    SLRGPU_CUDA_TRY(cudaEventRecord(ev0, stream));

    // per-stage device time (SLRGPU_RENDER_PROFILE_STAGES): one event pair per launch, summed at the end
    const bool profile = (p->flags & SLRGPU_RENDER_PROFILE_STAGES) != 0;
    struct StageTimer {
        std::vector<cudaEvent_t> ev[4];      // 0 raygen, 1 extend, 2 shade, 3 shadow: begin/end pairs
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

    unsigned long long generated = 0, extendRays = 0, shadowRays = 0, launches = 0, waves = 0;
    int cur = 0;
    uint32_t nCur = 0;
    bool overflow = false;
    while (true) {
        // refill the current queue with fresh camera samples
        const unsigned long long remaining = totalSamples - generated;
        const uint32_t room = P - nCur;
        const uint32_t fresh = (uint32_t)(remaining < room ? remaining : room);
        if (fresh) {
            timer.mark(0);
            if (rgb) launchRaygen<3>(sc, rc, generated, fresh, q[cur], nCur, stream);
            else launchRaygen<16>(sc, rc, generated, fresh, q[cur], nCur, stream);
            timer.mark(0);
            ++launches;
            generated += fresh;
            nCur += fresh;
        }
        if (nCur == 0) break;
        SLRGPU_CUDA_TRY(cudaMemsetAsync(dCounters, 0, 16, stream));
        ++waves;
        timer.mark(1);
        if ((rcode = launchExtend(sc, q[cur], nCur, hits, dCounters, profile, stream))) return rcode;
        timer.mark(1);
        timer.mark(2);
        if (rgb) launchShade<3>(sc, rc, q[cur], nCur, hits, q[cur ^ 1], sq, accumDev, dCounters, stream);
        else launchShade<16>(sc, rc, q[cur], nCur, hits, q[cur ^ 1], sq, accumDev, dCounters, stream);
        timer.mark(2);
        SLRGPU_CUDA_TRY(cudaGetLastError());
        launches += 2;
        extendRays += nCur;
        SLRGPU_CUDA_TRY(cudaMemcpyAsync(hCounters, dCounters, 16, cudaMemcpyDeviceToHost, stream));
        SLRGPU_CUDA_TRY(cudaStreamSynchronize(stream));
        const uint32_t nShadow = hCounters->numShadow;
        if (hCounters->stackOverflow) overflow = true;
        if (nShadow) {
            timer.mark(3);
            if ((rcode = launchShadow(sc, sq, nShadow, accumDev, dCounters, profile, stream))) return rcode;
            timer.mark(3);
            ++launches;
            shadowRays += nShadow;
        }
        nCur = hCounters->numNext;
        cur ^= 1;
    }
    SLRGPU_CUDA_TRY(cudaEventRecord(ev1, stream));
    SLRGPU_CUDA_TRY(cudaEventSynchronize(ev1));
    SLRGPU_CUDA_TRY(cudaMemcpy(hCounters, dCounters, sizeof(WavefrontCounters), cudaMemcpyDeviceToHost));
    if (hCounters->stackOverflow) overflow = true;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->paths = totalSamples;
        stats->extend_rays = extendRays; stats->shadow_rays = shadowRays;
        stats->rays = extendRays + shadowRays;
        stats->kernel_launches = launches;
        stats->waves = waves;
        cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
        if (profile) {
            stats->raygen_ms = timer.total(0); stats->extend_ms = timer.total(1);
            stats->shade_ms = timer.total(2); stats->shadow_ms = timer.total(3);
            stats->other_ms = stats->device_ms - stats->raygen_ms - stats->extend_ms - stats->shade_ms - stats->shadow_ms;
            stats->extend_nodes = hCounters->extendNodes; stats->extend_leaf_records = hCounters->extendLeafRecords;
            stats->shadow_nodes = hCounters->shadowNodes; stats->shadow_leaf_records = hCounters->shadowLeafRecords;
        }
    }
    if (overflow) { setError("traversal stack overflow (more than %d entries)", 64); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
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

SLRGPU_API int slrgpu_render_device(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* accumDev, void* stream, SlrGpuRenderStats* stats) {
    int rc = checkRenderArgs(sc, p);
    if (rc) return rc;
    if (!accumDev) { setError("slrgpu_render_device: null accumulation buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    return renderImpl(sc, p, accumDev, (cudaStream_t)stream, stats);
}

SLRGPU_API int slrgpu_render(SlrGpuScene* sc, const SlrGpuRenderParams* p, float* accum, SlrGpuRenderStats* stats) {
    int rc = checkRenderArgs(sc, p);
    if (rc) return rc;
    if (!accum) { setError("slrgpu_render: null accumulation buffer"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    const size_t bytes = (size_t)p->width * p->height * sc->channels * sizeof(float);
    float* dAccum = nullptr;
    SLRGPU_CUDA_TRY(cudaMalloc(&dAccum, bytes));
    struct Free { float* p; ~Free() { cudaFree(p); } } guard{dAccum};
    SLRGPU_CUDA_TRY(cudaMemset(dAccum, 0, bytes));
    rc = renderImpl(sc, p, dAccum, 0, stats);
    if (rc) return rc;
    SLRGPU_CUDA_TRY(cudaMemcpy(accum, dAccum, bytes, cudaMemcpyDeviceToHost));
    return SLRGPU_OK;
}

}  // extern "C"
