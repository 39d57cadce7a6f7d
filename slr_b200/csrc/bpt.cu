// Bidirectional path tracing on the GPU: what setRenderer("BPT") selects in the reference.
//   BidirectionalPathTracingRenderer::render / Job::kernel   libSLR/Renderers/BidirectionalPathTracingRenderer.cpp:25-252
//   Job::generateSubPath                                      BidirectionalPathTracingRenderer.cpp:254-340
//   Job::calculateMISWeight (power heuristic)                 BidirectionalPathTracingRenderer.cpp:342-413
//   BPTVertex / DDF proxies                                   BidirectionalPathTracingRenderer.h:20-91
//   Light::sampleRay (area / instanced / infinite sphere)     libSLR/Core/SurfaceObject.cpp:93-106, 187-207, 374-391
//   Camera::sampleRay, PerspectiveIDF                         libSLR/Core/cameras.h:46-57, Cameras/PerspectiveCamera.cpp:63-99
//   DiffuseEDF / IBLEDF                                       libSLR/EDFs/basic_EDFs.cpp:12-29, IBLEDF.cpp:11-29
//
// Samples are processed in batches of up to 262 144: one thread per SUBPATH traces the light / eye subpaths of the batch
// (the eye threads also add the implicit s = 0 paths), then one thread per CONNECTION (slot, s, t) runs its shadow ray and MIS
// weight. The reference keeps the two vertex lists in std::vectors of ~300-byte objects per CPU thread; here they live in
// HBM as [subpath][vertex index][slot] records of 192 B (208 with the MIS record kept apart): the subpath threads of a warp
// (adjacent slots, same vertex index) write one contiguous 6 KB block, a connection thread reads its two vertices as two
// contiguous records. The rays are the traversal of traverse.cuh (single_ray.cuh), surface points / lights / materials / BSDFs the
// path tracer's device functions; a vertex stores its surface point and material, and the BSDF is rebuilt from them when a
// connection needs it (the reference keeps the arena-allocated object instead).
//
// Differences to the reference, all stated in DESIGN.md: a subpath is cut at kBptMaxVerts vertices (the reference has no cap;
// survival beyond 64 vertices needs 63 Russian-roulette wins); the per-thread "separated" sensor buffers of the light-tracing
// splats (ImageSensor::add(idx, ...)) are the one accumulation buffer, written with atomics; random numbers are the counter-
// based Philox stream of rng.cuh keyed by (pixel, sample, dimension) instead of one xorshift stream per CPU thread.
#include "bsdf_rev.cuh"
#include "camera.cuh"
#include "single_ray.cuh"
#include <algorithm>
#include <cstring>
#include <mutex>

namespace slrgpu {

constexpr int kBptMaxVerts = 64;
constexpr int kBptBlock = 128;
// resident blocks per SM the kernel's register allocation is bounded for (tuning knob, profiles/r02_bpt.md)
#ifndef SLR_BPT_MIN_BLOCKS
#define SLR_BPT_MIN_BLOCKS 4
#endif
constexpr uint32_t kWlLambdaIsSelected = 1u;       // WavelengthSamples::LambdaIsSelected

enum BptVertexKind : uint32_t { BV_EDF_DIFFUSE = 0, BV_EDF_IBL = 1, BV_IDF = 2, BV_BSDF = 3 };

// BPTVertex without its MIS record (kept apart: calculateMISWeight walks only those)
template <int NC> struct alignas(16) BptVertex {
    SurfPt sp;
    V3 dirIn;                // dirIn_sn
    V3 gn;                   // gNormal_sn
    uint32_t material;       // whose BSDF (BV_BSDF) / whose emittance (light origin)
    uint32_t kindFlags;      // BptVertexKind | wlFlags << 8
    Spec<NC> alpha;
    __device__ __forceinline__ uint32_t kind() const { return kindFlags & 0xFFu; }
    __device__ __forceinline__ uint32_t wlFlags() const { return kindFlags >> 8; }
};
template <int NC> __host__ __device__ constexpr int bptVertexWords() { return (int)((sizeof(BptVertex<NC>) + 15) / 16); }

struct BptStore {
    float4* verts;           // [2 subpaths][kBptMaxVerts][lanes][W]: a vertex is W contiguous 16-byte words
    float4* mis;             // [2][kBptMaxVerts][lanes]: areaPDF, RRProb, revAreaPDF, revRRProb
    uint32_t lanes;
};
struct BptCounters {
    unsigned long long extendRays, shadowRays, connections, truncated;
    uint32_t stackOverflow, pad;
};
struct BptLocalCounts { uint32_t extendRays, shadowRays, connections, truncated; };      // per lane, added up once at the end

template <int NC>
__device__ __forceinline__ void storeVertex(const BptStore& st, uint32_t lane, int sub, int v, const BptVertex<NC>& vtx) {
    constexpr int W = bptVertexWords<NC>();
    const float4* src = reinterpret_cast<const float4*>(&vtx);
    float4* dst = st.verts + ((size_t)(sub * kBptMaxVerts + v) * st.lanes + lane) * W;
#pragma unroll
    for (int w = 0; w < W; ++w) dst[w] = src[w];
}
template <int NC>
__device__ __forceinline__ void loadVertex(const BptStore& st, uint32_t lane, int sub, int v, BptVertex<NC>* vtx) {
    constexpr int W = bptVertexWords<NC>();
    float4* dst = reinterpret_cast<float4*>(vtx);
    const float4* src = st.verts + ((size_t)(sub * kBptMaxVerts + v) * st.lanes + lane) * W;
#pragma unroll
    for (int w = 0; w < W; ++w) dst[w] = src[w];
}
// position, geometric normal and atInfinity of a stored vertex (SurfPt starts with p, gn; atInfinity is its last member)
template <int NC>
__device__ __forceinline__ void loadVertexPoint(const BptStore& st, uint32_t lane, int sub, int v, V3* p, V3* gn, bool* atInfinity) {
    BptVertex<NC> tmp;
    loadVertex<NC>(st, lane, sub, v, &tmp);     // only the words of sp survive dead-code elimination
    *p = tmp.sp.p; *gn = tmp.sp.gn; *atInfinity = tmp.sp.atInfinity;
}
__device__ __forceinline__ float4& misRecord(const BptStore& st, uint32_t lane, int sub, int v) {
    return st.mis[(size_t)(sub * kBptMaxVerts + v) * st.lanes + lane];
}

// CompensatedSum<float> (FloatSum) as calculateMISWeight uses it
struct KahanSum {
    float sum, c;
    __device__ __forceinline__ void add(float v) { const float y = v - c; const float t = sum + y; c = (t - sum) - y; sum = t; }
};

// Job::calculateMISWeight: 1 / (1 + sum of squared pdf ratios of every other way to build the same path)
static __device__ __noinline__ float calculateMISWeight(const BptStore& st, uint32_t lane, unsigned long long eyeDelta, unsigned long long lightDelta,
                                                        float lExtend1stAreaPDF, float lExtend1stRRProb, float lExtend2ndAreaPDF, float lExtend2ndRRProb,
                                                        float eExtend1stAreaPDF, float eExtend1stRRProb, float eExtend2ndAreaPDF, float eExtend2ndRRProb,
                                                        uint32_t numLVtx, uint32_t numEVtx) {
    constexpr uint32_t minEyeVertices = 1, minLightVertices = 0;
    KahanSum rec = {1.0f, 0.0f};
    // extend the light subpath into the eye subpath (no implicit light subpath reaching the lens)
    if (numEVtx > minEyeVertices) {
        const float4 end = misRecord(st, lane, 0, (int)numEVtx - 1);
        float ratio = lExtend1stAreaPDF * lExtend1stRRProb / (end.x * end.y);
        bool shortenDelta = (eyeDelta >> (numEVtx - 1)) & 1ull;
        if (!shortenDelta) rec.add(ratio * ratio);
        bool prevDelta = shortenDelta;
        if (numEVtx - 1 > minEyeVertices) {
            const float4 v2 = misRecord(st, lane, 0, (int)numEVtx - 2);
            ratio *= lExtend2ndAreaPDF * lExtend2ndRRProb / (v2.x * v2.y);
            shortenDelta = (eyeDelta >> (numEVtx - 2)) & 1ull;
            if (!shortenDelta && !prevDelta) rec.add(ratio * ratio);
            prevDelta = shortenDelta;
            for (int t = (int)numEVtx - 2; t > (int)minEyeVertices; --t) {
                const float4 v = misRecord(st, lane, 0, t - 1);
                ratio *= v.z * v.w / (v.x * v.y);
                shortenDelta = (eyeDelta >> (t - 1)) & 1ull;
                if (!shortenDelta && !prevDelta) rec.add(ratio * ratio);
                prevDelta = shortenDelta;
            }
        }
    }
    // extend the eye subpath into the light subpath (down to the implicit path that hits the light)
    if (numLVtx > minLightVertices) {
        const float4 end = misRecord(st, lane, 1, (int)numLVtx - 1);
        float ratio = eExtend1stAreaPDF * eExtend1stRRProb / (end.x * end.y);
        bool shortenDelta = (lightDelta >> (numLVtx - 1)) & 1ull;
        if (!shortenDelta) rec.add(ratio * ratio);
        bool prevDelta = shortenDelta;
        if (numLVtx - 1 > minLightVertices) {
            const float4 v2 = misRecord(st, lane, 1, (int)numLVtx - 2);
            ratio *= eExtend2ndAreaPDF * eExtend2ndRRProb / (v2.x * v2.y);
            shortenDelta = (lightDelta >> (numLVtx - 2)) & 1ull;
            if (!shortenDelta && !prevDelta) rec.add(ratio * ratio);
            prevDelta = shortenDelta;
            for (int s = (int)numLVtx - 2; s > (int)minLightVertices; --s) {
                const float4 v = misRecord(st, lane, 1, s - 1);
                ratio *= v.z * v.w / (v.x * v.y);
                shortenDelta = (lightDelta >> (s - 1)) & 1ull;
                if (!shortenDelta && !prevDelta) rec.add(ratio * ratio);
                prevDelta = shortenDelta;
            }
        }
    }
    return 1.0f / rec.sum;
}

// what a sample shares between its subpaths and connections
struct BptSample {
    uint32_t pixelKey, sample;      // random-number key
    uint32_t sensorPixel;           // ImageSensor::add's pixel of the eye path
    float wlOffset, time;
    uint32_t hero;
    bool inPlace;
    unsigned long long eyeDelta, lightDelta;
    uint32_t numE, numL;
};

// sensor->add(px, py, wls, contribution): a non-finite sample is dropped (the reference asserts, compiled out)
template <int NC>
__device__ __forceinline__ void bptSplat(const RenderConstants& rc, float* __restrict__ accum, uint32_t pixel, const BptSample& smp, const Spec<NC>& c, float scale) {
    float v[NC == 3 ? 4 : NC];
    float sum = 0.0f;
    const float k = scale * rc.recBinWidth;
#pragma unroll
    for (int i = 0; i < NC; ++i) { v[i] = c.v[i] * k; sum += v[i]; }
    if (!isfinite(sum) || sum == 0.0f) return;
    splat<NC>(accum, pixel, smp.wlOffset, smp.inPlace, v);
}

// random block `block` of the sample: camera sample 0, 1 (camera.cuh); light origin 0x10000, 0x10001; eye bounce k 0x100 + k;
// light bounce k 0x10100 + k (BSDF component, BSDF u0, u1, Russian roulette)
__device__ __forceinline__ Rand4 bptRandom(const RenderConstants& rc, const BptSample& smp, uint32_t block) {
    return pathRandom(rc.seed, smp.pixelKey, smp.sample, block);
}

// DiffuseEDF / IBLEDF / PerspectiveIDF evaluate + evaluatePDF for a local direction (EDFProxy, IDFProxy)
__device__ __forceinline__ float worldDiscArea(const DeviceScene& s) { return kPi * s.worldRadius * s.worldRadius; }
__device__ __forceinline__ bool idfFocus(const DeviceScene& s, const RenderConstants& rc, float lensU, float lensV, const V3& dirIn, float* fx, float* fy) {
    const SlrGpuCamera& cam = s.camera;
    const float k = cam.obj_plane_dist / dirIn.z;
    *fx = dirIn.x * k + cam.lens_radius * lensU;
    *fy = dirIn.y * k + cam.lens_radius * lensV;
    return *fx >= -rc.opWidth * 0.5f && *fx <= rc.opWidth * 0.5f && *fy >= -rc.opHeight * 0.5f && *fy <= rc.opHeight * 0.5f && dirIn.z >= 0;
}

// ddf->evaluate + ddf->evaluatePDF of a stored vertex for the local direction `dir`; `bsdf` must be the vertex's BSDF when
// its kind is BV_BSDF. *revPdf is only defined for BSDF vertices (the EDF / IDF proxies leave it untouched: 0 here).
template <int NC>
__device__ __forceinline__ Spec<NC> vertexEvaluate(const DeviceScene& s, const RenderConstants& rc, const BptVertex<NC>& vtx, const Bsdf<NC, 4>& bsdf,
                                                   bool adjoint, uint32_t hero, const V3& dir, float* dirPdf, float* revPdf) {
    *revPdf = 0.0f;
    switch (vtx.kind()) {
    case BV_EDF_DIFFUSE:
        *dirPdf = dir.z > 0.0f ? dir.z / kPi : 0.0f;
        return specConst<NC>(dir.z > 0.0f ? 1.0f / kPi : 0.0f);
    case BV_EDF_IBL:
        *dirPdf = 1.0f / worldDiscArea(s);
        return specConst<NC>(1.0f / kPi);
    case BV_IDF: {
        float fx, fy;
        const bool valid = idfFocus(s, rc, vtx.sp.u, vtx.sp.v, dir, &fx, &fy);
        const SlrGpuCamera& cam = s.camera;
        *dirPdf = valid ? cam.img_plane_dist * cam.img_plane_dist / ((dir.z * dir.z * dir.z) * rc.imgPlaneArea) : 0.0f;
        return specConst<NC>(valid ? 1.0f : 0.0f);
    }
    default: {
        BsdfQuery q;
        q.dir = vtx.dirIn; q.gn = vtx.gn; q.hero = hero; q.flags = DT_All; q.adjoint = adjoint;
        const Spec<NC> f = bsdfEvaluate(bsdf, q, dir);
        *dirPdf = bsdfPdfRev(bsdf, q, dir, revPdf);
        return f;
    }
    }
}

// Job::generateSubPath: traces the subpath that starts with the ray (org, dir) leaving vertex 0 (already stored, its MIS record
// in `prevMis`), appends its vertices to list `sub` (0 eye, 1 light = adjoint) and, for the eye subpath, adds the implicit
// (s = 0) contributions of emitters it hits.
template <int NC>
static __device__ __noinline__ void generateSubPath(const DeviceScene& s, const RenderConstants& rc, const BptStore& st, uint32_t lane, int sub, BptSample& smp,
                                                    V3 org, V3 dir, float tmin, Spec<NC> alpha, float dirPDF, uint32_t sampledType, float cosLast,
                                                    V3 prevP, bool prevInf, float4 prevMis, float* __restrict__ accum, uint32_t* stack, BptLocalCounts* counts,
                                                    bool* overflow) {
    const bool adjoint = sub == 1;
    if (dirPDF == 0.0f) return;
    uint32_t wlFlags = 0;
    float RRProb = 1.0f;
    uint32_t n = 1;
    unsigned long long delta = adjoint ? smp.lightDelta : smp.eyeDelta;
    const uint32_t rngBase = adjoint ? 0x10100u : 0x100u;
    for (uint32_t bounce = 0;; ++bounce) {
        // scene->intersect(ray, &isect) + isect.getSurfacePoint
        WalkState w;
        w.r.ox = org.x; w.r.oy = org.y; w.r.oz = org.z; w.r.tmin = tmin;
        w.r.dx = dir.x; w.r.dy = dir.y; w.r.dz = dir.z; w.r.tmax = INFINITY;
        w.time = smp.time;
        singleRayWalk<false>(s, w, stack, overflow);
        ++counts->extendRays;
        BptVertex<NC> vtx;
        float localArea = 1.0f;
        SlrGpuTriangle tri = {};
        uint32_t inst = SLRGPU_INVALID_ID;
        if (w.hit.prim == SLRGPU_INVALID_ID) {
            if (!s.envPresent) break;
            envSurfacePoint(dir, &vtx.sp);
            vtx.material = s.envMaterial;
        } else {
            inst = w.hit.inst;
            tri = hitSurfacePoint(s, w.hit.prim, w.hit.inst, w.hit.t, w.hit.u, w.hit.v, org, dir, smp.time, &vtx.sp, &localArea);
            vtx.material = tri.material;
        }
        const SurfPt& sp = vtx.sp;
        const float dist2 = (prevInf || sp.atInfinity) ? 1.0f : sqLength(prevP - sp.p);
        const V3 dirOut = sp.sf.toLocal(-dir);
        const V3 gNorm = sp.sf.toLocal(sp.gn);
        const float areaPDF = dirPDF * absDot(dirOut, gNorm) / dist2;
        if (n >= (uint32_t)kBptMaxVerts) { ++counts->truncated; break; }
        vtx.dirIn = dirOut; vtx.gn = gNorm;
        vtx.kindFlags = BV_BSDF | (wlFlags << 8);
        vtx.alpha = alpha;
        storeVertex<NC>(st, lane, sub, (int)n, vtx);
        const float storedRRProb = RRProb;
        misRecord(st, lane, sub, (int)n) = make_float4(areaPDF, storedRRProb, CUDART_NAN_F, CUDART_NAN_F);
        if (dtIsDelta(sampledType)) delta |= 1ull << n; else delta &= ~(1ull << n);
        ++n;

        // implicit path (zero light subpath vertices, s = 0)
        const bool emitting = sp.atInfinity || materialIsEmitting(s, vtx.material);
        if (!adjoint && emitting) {
            const Spec<NC> Le0 = materialEmittance<NC>(s, vtx.material, sp, smp.wlOffset);
            const float Le1 = (sp.atInfinity || dirOut.z > 0.0f) ? 1.0f / kPi : 0.0f;
            const float lightProb = lightSelectionProb(s, tri, inst, sp.atInfinity);
            const float lightAreaPDF = sp.atInfinity ? envEvaluateUVPDF(s, sp.u / (2 * kPi), sp.v / kPi) / (2 * kPi * kPi * sinf(sp.v)) : 1.0f / localArea;
            const float edfPDF = sp.atInfinity ? 1.0f / worldDiscArea(s) : (dirOut.z > 0.0f ? dirOut.z / kPi : 0.0f);
            const float extend1stAreaPDF = lightProb * lightAreaPDF;
            const float extend2ndAreaPDF = edfPDF * cosLast / dist2;
            const float mis = calculateMISWeight(st, lane, delta, 0ull, extend1stAreaPDF, 1.0f, extend2ndAreaPDF, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0u, n);
            if (!isinf(mis) && !isnan(mis)) {
                float scale = mis * Le1;
                if (wlFlags & kWlLambdaIsSelected) scale *= NC;
                bptSplat<NC>(rc, accum, smp.sensorPixel, smp, alpha * Le0, scale);
            }
        }
        if (sp.atInfinity) { --n; break; }

        Bsdf<NC, 4> bsdf;
        buildBsdf<NC, 4>(s, vtx.material, sp, smp.wlOffset, (wlFlags & kWlLambdaIsSelected) != 0, &bsdf);
        BsdfQuery q;
        q.dir = dirOut; q.gn = gNorm; q.hero = smp.hero; q.flags = DT_All; q.adjoint = adjoint;
        const Rand4 u = bptRandom(rc, smp, rngBase + bounce);
        BsdfSampleResult res;
        BsdfRev<NC> rev;
        const Spec<NC> fs = bsdfSampleRev(bsdf, q, u.x, u.y, u.z, &res, &rev);
        if (specIsZero(fs) || res.pdf == 0.0f) break;
        if (res.type & DT_Dispersive) wlFlags |= kWlLambdaIsSelected;
        const float cosIn = absDot(res.dir, gNorm);
        Spec<NC> weight = fs * (cosIn / res.pdf);

        // Russian roulette
        RRProb = fminf(specImportance(weight, smp.hero), 1.0f);
        if (u.w < RRProb) weight = weight * (1.0f / RRProb);
        else break;
        if (!isfinite(specImportance(weight, smp.hero))) break;      // a NaN / inf throughput ends the path (never splatted)

        alpha = alpha * weight;
        org = sp.p; dir = sp.sf.fromLocal(res.dir); tmin = 0.0001f;      // Ray::Epsilon

        // the vertex before this one learns what sampling it from here would have cost
        const float revAreaPDF = rev.pdf * cosLast / dist2;
        const float revRRProb = fminf(specImportance(rev.fs * (absDot(dirOut, gNorm) / rev.pdf), smp.hero), 1.0f);
        misRecord(st, lane, sub, (int)n - 2) = make_float4(prevMis.x, prevMis.y, revAreaPDF, revRRProb);
        prevMis = make_float4(areaPDF, storedRRProb, 0.0f, 0.0f);

        cosLast = cosIn;
        dirPDF = res.pdf;
        sampledType = res.type;
        prevP = sp.p; prevInf = false;
    }
    if (adjoint) { smp.lightDelta = delta; smp.numL = n; }
    else { smp.eyeDelta = delta; smp.numE = n; }
}

// Vector3::makeCoordinateSystem (BasicTypes/Vector3.h:58-68)
__device__ __forceinline__ void makeCoordinateSystem(const V3& v, V3* vx, V3* vy) {
    if (fabsf(v.x) > fabsf(v.y)) { const float invLen = 1.0f / sqrtf(v.x * v.x + v.z * v.z); *vx = V3(-v.z * invLen, 0.0f, v.x * invLen); }
    else { const float invLen = 1.0f / sqrtf(v.y * v.y + v.z * v.z); *vx = V3(0.0f, v.z * invLen, -v.y * invLen); }
    *vy = cross(v, *vx);
}

// the connection of eye vertex t - 1 with light vertex s - 1 (BidirectionalPathTracingRenderer.cpp:166-249)
template <int NC>
static __device__ __noinline__ void connectVertices(const DeviceScene& s, const RenderConstants& rc, const BptStore& st, uint32_t lane, const BptSample& smp,
                                                    const BptVertex<NC>& eVtx, const Bsdf<NC, 4>& eBsdf, uint32_t t, uint32_t sIdx,
                                                    float* __restrict__ accum, uint32_t* stack, BptLocalCounts* counts, bool* overflow) {
    BptVertex<NC> lVtx;
    loadVertex<NC>(st, lane, 1, (int)sIdx - 1, &lVtx);
    // the remaining factors of the full path that are not in the precomputed weights
    float connectDist2;
    V3 connectionVector;
    if (lVtx.sp.atInfinity) { connectDist2 = 1.0f; connectionVector = normalize(lVtx.sp.p); }
    else { const V3 d = lVtx.sp.p - eVtx.sp.p; connectDist2 = sqLength(d); connectionVector = d / sqrtf(connectDist2); }
    const float cosLightEnd = absDot(connectionVector, lVtx.sp.gn);
    const float cosEyeEnd = absDot(connectionVector, eVtx.sp.gn);
    const float G = cosEyeEnd * cosLightEnd / connectDist2;

    const V3 lConnectVector = lVtx.sp.sf.toLocal(-connectionVector);
    Bsdf<NC, 4> lBsdf;
    if (lVtx.kind() == BV_BSDF) buildBsdf<NC, 4>(s, lVtx.material, lVtx.sp, smp.wlOffset, (lVtx.wlFlags() & kWlLambdaIsSelected) != 0, &lBsdf);
    float lExtend1stDirPDF, eExtend2ndDirPDF;
    const Spec<NC> lDDF = vertexEvaluate<NC>(s, rc, lVtx, lBsdf, true, smp.hero, lConnectVector, &lExtend1stDirPDF, &eExtend2ndDirPDF);

    const V3 eConnectVector = eVtx.sp.sf.toLocal(connectionVector);
    float eExtend1stDirPDF, lExtend2ndDirPDF;
    const Spec<NC> eDDF = vertexEvaluate<NC>(s, rc, eVtx, eBsdf, false, smp.hero, eConnectVector, &eExtend1stDirPDF, &lExtend2ndDirPDF);

    float wlProb = 1.0f;
    if ((lVtx.wlFlags() | eVtx.wlFlags()) & kWlLambdaIsSelected) wlProb = 1.0f / NC;
    const Spec<NC> connectionTerm = lDDF * (G / wlProb) * eDDF;
    if (specIsZero(connectionTerm)) return;

    // scene->testVisibility(eVtx.surfPt, lVtx.surfPt, time)
    {
        WalkState w;
        w.r.ox = eVtx.sp.p.x; w.r.oy = eVtx.sp.p.y; w.r.oz = eVtx.sp.p.z; w.r.tmin = 0.0001f;
        if (lVtx.sp.atInfinity) {
            const V3 d = normalize(lVtx.sp.p);
            w.r.dx = d.x; w.r.dy = d.y; w.r.dz = d.z; w.r.tmax = 3.402823466e+38f;
        } else {
            const float dist = length(lVtx.sp.p - eVtx.sp.p);
            const V3 d = (lVtx.sp.p - eVtx.sp.p) / dist;
            w.r.dx = d.x; w.r.dy = d.y; w.r.dz = d.z; w.r.tmax = dist * (1.0f - 0.0001f);
        }
        w.time = smp.time;
        singleRayWalk<true>(s, w, stack, overflow);
        ++counts->shadowRays;
        if (w.found) return;
    }

    // the 1st and 2nd subpath-extending pdfs and roulette probabilities: they depend on the connection
    const float lExtend1stAreaPDF = lExtend1stDirPDF * cosEyeEnd / connectDist2;
    const float lExtend1stRRProb = sIdx > 1 ? fminf(specImportance(lDDF * (cosLightEnd / lExtend1stDirPDF), smp.hero), 1.0f) : 1.0f;
    float lExtend2ndAreaPDF = 0.0f, lExtend2ndRRProb = 0.0f;
    if (t > 1) {
        V3 p2, gn2; bool inf2;
        loadVertexPoint<NC>(st, lane, 0, (int)t - 2, &p2, &gn2, &inf2);
        const V3 d = eVtx.sp.p - p2;                 // eye vertices are never at infinity
        const float dist2 = sqLength(d);
        const V3 dir2nd = d / sqrtf(dist2);
        lExtend2ndAreaPDF = lExtend2ndDirPDF * absDot(gn2, dir2nd) / dist2;
        lExtend2ndRRProb = fminf(specImportance(eDDF * (absDot(eVtx.gn, eVtx.dirIn) / lExtend2ndDirPDF), smp.hero), 1.0f);
    }
    const float eExtend1stAreaPDF = eExtend1stDirPDF * cosLightEnd / connectDist2;
    const float eExtend1stRRProb = t > 1 ? fminf(specImportance(eDDF * (cosEyeEnd / eExtend1stDirPDF), smp.hero), 1.0f) : 1.0f;
    float eExtend2ndAreaPDF = 0.0f, eExtend2ndRRProb = 0.0f;
    if (sIdx > 1) {
        V3 p2, gn2; bool inf2;
        loadVertexPoint<NC>(st, lane, 1, (int)sIdx - 2, &p2, &gn2, &inf2);
        float dist2;
        V3 dir2nd;
        if (inf2) { dist2 = 1.0f; dir2nd = normalize(p2); }
        else { const V3 d = p2 - lVtx.sp.p; dist2 = sqLength(d); dir2nd = d / sqrtf(dist2); }
        eExtend2ndAreaPDF = eExtend2ndDirPDF * absDot(gn2, dir2nd) / dist2;
        eExtend2ndRRProb = fminf(specImportance(lDDF * (absDot(lVtx.gn, lVtx.dirIn) / eExtend2ndDirPDF), smp.hero), 1.0f);
    }

    const float mis = calculateMISWeight(st, lane, smp.eyeDelta, smp.lightDelta, lExtend1stAreaPDF, lExtend1stRRProb, lExtend2ndAreaPDF, lExtend2ndRRProb,
                                         eExtend1stAreaPDF, eExtend1stRRProb, eExtend2ndAreaPDF, eExtend2ndRRProb, sIdx, t);
    if (isinf(mis) || isnan(mis)) return;
    const Spec<NC> contribution = lVtx.alpha * connectionTerm * eVtx.alpha;
    if (t > 1) {
        bptSplat<NC>(rc, accum, smp.sensorPixel, smp, contribution, mis);
    } else {
        // light tracing: the pixel the connection lands on (PerspectiveIDF::calculatePixel)
        float fx, fy;
        idfFocus(s, rc, eVtx.sp.u, eVtx.sp.v, eConnectVector, &fx, &fy);
        const float hitPx = rc.width * (0.5f - fx / rc.opWidth), hitPy = rc.height * (0.5f - fy / rc.opHeight);
        const uint32_t ipx = min((uint32_t)hitPx, rc.width - 1), ipy = min((uint32_t)hitPy, rc.height - 1);
        bptSplat<NC>(rc, accum, ipy * rc.width + ipx, smp, contribution, mis);
    }
}

// ---------------------------------------------------------------------------------------------
// The stages of a batch of samples (slot i of the batch = sample firstSample + i):
//   subpathKernel   2 x batch threads: thread i < batch traces the eye subpath of slot i (and keeps what the sample's other
//                   stages need), thread batch + i its light subpath -- a warp is all eye or all light
//   countKernel     connections per slot = numE x numL (then an exclusive prefix sum over the slots)
//   connectKernel   one thread per (slot, s, t): the thread's slot by binary search in the prefix sums
// A first version ran a whole sample per lane in one kernel: ncu showed 4 of 32 lanes active per instruction and 24 cycles
// of instruction-fetch stall per issue (the lanes of a warp sat in different phases of a 700 KB kernel), profiles/r02_bpt.md.
// ---------------------------------------------------------------------------------------------
struct BptSlotCommon {          // written by the slot's eye thread
    uint32_t pixelKey, sample, sensorPixel, heroInPlace;     // hero | inPlace << 8
    float wlOffset, time;
};
struct BptSlotEnd { unsigned long long delta; uint32_t num, pad; };
struct BptBatch {
    BptSlotCommon* common;
    BptSlotEnd* eyeEnd;
    BptSlotEnd* lightEnd;
    uint32_t* connCount;        // [batch] connections of a slot; connBase[batch + 1] = their exclusive prefix sums
    uint32_t* connBase;
};

__device__ __forceinline__ void samplePixel(const RenderConstants& rc, unsigned long long lin, uint32_t* x, uint32_t* y, uint32_t* pass) {
    // pixel order of the path tracer's ray generation: bands of 8 rows, column-major inside a band
    *pass = (uint32_t)(lin / rc.numPixels);
    const uint32_t r = (uint32_t)(lin % rc.numPixels);
    const uint32_t band = r / (8u * rc.width);
    const uint32_t local = r - band * 8u * rc.width;
    const uint32_t rows = min(8u, rc.height - band * 8u);
    *x = local / rows; *y = band * 8u + local % rows;
}

__device__ __forceinline__ void flushCounts(const BptLocalCounts& counts, bool overflow, BptCounters* counters) {
    // one atomic per warp and counter
    uint32_t e = counts.extendRays, sh = counts.shadowRays, cn = counts.connections, tr = counts.truncated;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xFFFFFFFFu, e, o); sh += __shfl_xor_sync(0xFFFFFFFFu, sh, o);
        cn += __shfl_xor_sync(0xFFFFFFFFu, cn, o); tr += __shfl_xor_sync(0xFFFFFFFFu, tr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (e) atomicAdd(&counters->extendRays, (unsigned long long)e);
        if (sh) atomicAdd(&counters->shadowRays, (unsigned long long)sh);
        if (cn) atomicAdd(&counters->connections, (unsigned long long)cn);
        if (tr) atomicAdd(&counters->truncated, (unsigned long long)tr);
    }
    if (overflow) atomicExch(&counters->stackOverflow, 1u);
}

template <int NC>
__global__ void __launch_bounds__(kBptBlock, SLR_BPT_MIN_BLOCKS)
subpathKernel(const DeviceScene s, const RenderConstants rc, BptStore st, BptBatch batch, float* __restrict__ accum, BptCounters* counters,
              unsigned long long firstSample, uint32_t count) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;       // count is a multiple of the warp size or the last batch
    const bool lightThread = tid >= st.lanes;
    const uint32_t slot = lightThread ? tid - st.lanes : tid;
    uint32_t stack[kStackSize];
    bool overflow = false;
    BptLocalCounts counts = {0, 0, 0, 0};
    if (slot < count) {
        uint32_t x, y, pass;
        samplePixel(rc, firstSample + slot, &x, &y, &pass);
        const uint32_t pixel = y * rc.width + x;
        BptSample smp;
        smp.pixelKey = pixel; smp.sample = rc.sppBegin + pass;
        smp.eyeDelta = 0; smp.lightDelta = 0; smp.numE = 0; smp.numL = 0;
        if (lightThread) {
            // ---- light subpath; of the camera sample it needs the time, the wavelengths and the hero wavelength
            const Rand4 r0 = pathRandom(rc.seed, pixel, smp.sample, 0);
            const Rand4 r1 = pathRandom(rc.seed, pixel, smp.sample, 1);
            smp.wlOffset = r0.w;
            smp.time = rc.timeStart * (1 - r0.x) + rc.timeEnd * r0.x;
            smp.hero = min((uint32_t)(NC * r1.x), (uint32_t)(NC - 1));
            smp.sensorPixel = 0; smp.inPlace = false;
            if (s.numTopLights > 0 || s.envPresent) {
                const Rand4 l0 = bptRandom(rc, smp, 0x10000u);      // light selection, light position u0, u1
                const Rand4 l1 = bptRandom(rc, smp, 0x10001u);      // EDF direction u0, u1
                LightSample ls;
                sampleLight(s, l0.x, l0.y, l0.z, smp.time, &ls);
                const Spec<NC> Le0 = materialEmittance<NC>(s, ls.material, ls.sp, smp.wlOffset);
                V3 org, dir;
                float dirPDF, dirZ, tmin;
                if (ls.isEnv) {
                    // InfiniteSphereSurfaceObject::sampleRay: parallel rays towards the scene through a disc of the world's radius
                    dirZ = 1.0f; dirPDF = 1.0f / worldDiscArea(s);
                    const V3 vz = ls.sp.sf.z;
                    V3 vx, vy;
                    makeCoordinateSystem(vz, &vx, &vy);
                    float dx, dy;
                    concentricSampleDisk(l1.x, l1.y, &dx, &dy);
                    const float R = s.worldRadius;
                    org = V3(s.worldCenter[0], s.worldCenter[1], s.worldCenter[2]) + (1.1f * R) * ls.sp.p + R * (dx * vx + dy * vy);
                    dir = vz; tmin = 0.0f;
                } else {
                    const V3 d = cosineSampleHemisphere(l1.x, l1.y);      // DiffuseEDF::sample
                    dirZ = d.z; dirPDF = d.z / kPi;
                    org = ls.sp.p; dir = ls.sp.sf.fromLocal(d); tmin = 0.0001f;
                    if (ls.sp.inst != SLRGPU_INVALID_ID) {
                        // TransformedSurfaceObject::sampleRay (SurfaceObject.cpp:374-391): the ray is built in the object's space
                        // and then transformed -- its direction is NOT re-normalised, so under a scaled instance the first
                        // segment of the light subpath carries the scale in |dir| (into the cosine of its throughput and into
                        // dirOut_sn at its first hit), exactly as the reference's does
                        const SlrGpuTriangle tri = s.triangles[ls.sp.prim];
                        SurfPt obj;
                        triangleFrame(loadTriangle(s, tri), ls.sp.u, ls.sp.v, 1.0f - ls.sp.u - ls.sp.v, false, &obj);
                        float scratch[32];
                        const InstanceXfm xf = instanceTransformAt(s, s.instances[ls.sp.inst], smp.time, scratch);
                        dir = xfmVector(xf.mat, obj.sf.fromLocal(d));
                    }
                }
                const float lightAreaPDF = ls.lightPDF;
                BptVertex<NC> v0;
                v0.sp = ls.sp;
                v0.dirIn = V3(0, 0, 0); v0.gn = V3(0, 0, 1);
                v0.material = ls.material;
                v0.kindFlags = ls.isEnv ? BV_EDF_IBL : BV_EDF_DIFFUSE;
                v0.alpha = Le0 * (1.0f / lightAreaPDF);
                storeVertex<NC>(st, slot, 1, 0, v0);
                const float4 mis0 = make_float4(lightAreaPDF, 1.0f, CUDART_NAN_F, CUDART_NAN_F);
                misRecord(st, slot, 1, 0) = mis0;
                smp.numL = 1;
                const Spec<NC> alpha = v0.alpha * ((1.0f / kPi) * (absDot(dir, ls.sp.gn) / dirPDF));
                generateSubPath<NC>(s, rc, st, slot, 1, smp, org, dir, tmin, alpha, dirPDF, DT_Reflection | DT_LowFreq, dirZ,
                                    ls.sp.p, ls.sp.atInfinity, mis0, accum, stack, &counts, &overflow);
            }
            BptSlotEnd end;
            end.delta = smp.lightDelta; end.num = smp.numL; end.pad = 0;
            batch.lightEnd[slot] = end;
        } else {
            // ---- eye subpath: time, pixel position, wavelengths (Job::kernel, BidirectionalPathTracingRenderer.cpp:104-110), lens sample
            CameraSample cs;
            sampleCamera<NC>(s, rc, x, y, pixel, smp.sample, &cs);
            smp.sensorPixel = cs.ipy * rc.width + cs.ipx;
            smp.wlOffset = cs.wlOffset; smp.time = cs.time; smp.hero = cs.hero;
            smp.inPlace = (cs.flags & kFlagStrataInPlace) != 0;
            BptSlotCommon c;
            c.pixelKey = smp.pixelKey; c.sample = smp.sample; c.sensorPixel = smp.sensorPixel;
            c.heroInPlace = smp.hero | (smp.inPlace ? 0x100u : 0u);
            c.wlOffset = smp.wlOffset; c.time = smp.time;
            batch.common[slot] = c;
            BptVertex<NC> v0;
            v0.sp.p = cs.org; v0.sp.gn = cs.lensFrame.z; v0.sp.sf = cs.lensFrame;
            v0.sp.u = cs.lensU; v0.sp.v = cs.lensV; v0.sp.tu = 0.0f; v0.sp.tv = 0.0f;
            v0.sp.prim = SLRGPU_INVALID_ID; v0.sp.inst = SLRGPU_INVALID_ID; v0.sp.atInfinity = false;
            v0.dirIn = V3(0, 0, 0); v0.gn = V3(0, 0, 1);
            v0.material = SLRGPU_INVALID_ID;
            v0.kindFlags = BV_IDF;
            v0.alpha = specConst<NC>(1.0f / (rc.lensAreaPDF * rc.selectWLPDF));
            storeVertex<NC>(st, slot, 0, 0, v0);
            const float4 mis0 = make_float4(rc.lensAreaPDF, 1.0f, CUDART_NAN_F, CUDART_NAN_F);
            misRecord(st, slot, 0, 0) = mis0;
            if (!(s.camera.lens_radius > 0.0f)) smp.eyeDelta |= 1ull;       // posType Delta0D for a pinhole
            smp.numE = 1;
            const Spec<NC> alpha = v0.alpha * (absDot(cs.dir, cs.lensFrame.z) / cs.dirPDF);
            generateSubPath<NC>(s, rc, st, slot, 0, smp, cs.org, cs.dir, 0.0f, alpha, cs.dirPDF, DT_Reflection | DT_LowFreq, cs.dirLocalZ,
                                cs.org, false, mis0, accum, stack, &counts, &overflow);
            BptSlotEnd end;
            end.delta = smp.eyeDelta; end.num = smp.numE; end.pad = 0;
            batch.eyeEnd[slot] = end;
        }
    }
    flushCounts(counts, overflow, counters);
}

__global__ void __launch_bounds__(256)
countKernel(BptBatch batch, uint32_t count) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < count) batch.connCount[slot] = batch.eyeEnd[slot].num * batch.lightEnd[slot].num;
}

// exclusive prefix sums of connCount[0, count) into connBase[0, count], one block (a batch is a few hundred thousand slots)
__global__ void __launch_bounds__(1024)
scanKernel(BptBatch batch, uint32_t count) {
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < count; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < count ? batch.connCount[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += n; }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warpSums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, w, o); if ((int)lane >= o) w += n; }
            warpSums[lane] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (warp ? warpSums[warp - 1] : 0u) + incl - v;
        if (i < count) batch.connBase[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) batch.connBase[count] = carry;
}

template <int NC>
__global__ void __launch_bounds__(kBptBlock, SLR_BPT_MIN_BLOCKS)
connectKernel(const DeviceScene s, const RenderConstants rc, BptStore st, BptBatch batch, float* __restrict__ accum, BptCounters* counters, uint32_t count) {
    uint32_t stack[kStackSize];
    bool overflow = false;
    BptLocalCounts counts = {0, 0, 0, 0};
    const uint32_t total = batch.connBase[count];
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
        // the slot whose range [connBase[slot], connBase[slot + 1]) holds c
        uint32_t lo = 0, hi = count;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (batch.connBase[mid] <= c) lo = mid; else hi = mid; }
        const uint32_t slot = lo;
        const BptSlotCommon cm = batch.common[slot];
        const BptSlotEnd ee = batch.eyeEnd[slot], le = batch.lightEnd[slot];
        BptSample smp;
        smp.pixelKey = cm.pixelKey; smp.sample = cm.sample; smp.sensorPixel = cm.sensorPixel;
        smp.hero = cm.heroInPlace & 0xFFu; smp.inPlace = (cm.heroInPlace & 0x100u) != 0;
        smp.wlOffset = cm.wlOffset; smp.time = cm.time;
        smp.eyeDelta = ee.delta; smp.lightDelta = le.delta; smp.numE = ee.num; smp.numL = le.num;
        const uint32_t k = c - batch.connBase[slot];
        const uint32_t t = k / le.num + 1u, sIdx = k % le.num + 1u;
        BptVertex<NC> eVtx;
        loadVertex<NC>(st, slot, 0, (int)t - 1, &eVtx);
        Bsdf<NC, 4> eBsdf;
        if (eVtx.kind() == BV_BSDF) buildBsdf<NC, 4>(s, eVtx.material, eVtx.sp, smp.wlOffset, (eVtx.wlFlags() & kWlLambdaIsSelected) != 0, &eBsdf);
        connectVertices<NC>(s, rc, st, slot, smp, eVtx, eBsdf, t, sIdx, accum, stack, &counts, &overflow);
        ++counts.connections;
    }
    flushCounts(counts, overflow, counters);
}

// ---------------------------------------------------------------------------------------------
// host side: vertex storage per device (kept for later calls), the batch loop, statistics
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kBptBatchSlots = 1u << 18;        // 262 144 samples per batch: 6.7 GB of vertex storage in spectral mode

struct BptWorkspace {
    BptStore store = {};
    BptBatch batch = {};
    BptCounters* dCounters = nullptr;
    size_t vertWords = 0;
    void release() {
        void* ptrs[] = {store.verts, store.mis, batch.common, batch.eyeEnd, batch.lightEnd, batch.connCount, batch.connBase, dCounters};
        for (void* q : ptrs) if (q) cudaFree(q);
        store = {}; batch = {}; dCounters = nullptr; vertWords = 0;
    }
};
static std::mutex g_bptMutex;
static BptWorkspace* g_bptPool[64] = {};

void releaseBptWorkspaces() {
    std::lock_guard<std::mutex> lock(g_bptMutex);
    for (int d = 0; d < 64; ++d)
        if (g_bptPool[d]) { cudaSetDevice(d); g_bptPool[d]->release(); delete g_bptPool[d]; g_bptPool[d] = nullptr; }
}

template <int NC>
static int renderBptT(SlrGpuScene* sc, const RenderConstants& rc, unsigned long long totalSamples, float* accumDev, cudaStream_t stream, SlrGpuRenderStats* stats) {
    const uint32_t slots = (uint32_t)std::min<unsigned long long>(kBptBatchSlots, (totalSamples + 127ull) & ~127ull);

    BptWorkspace* w = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_bptMutex);
        if (sc->device >= 0 && sc->device < 64) { w = g_bptPool[sc->device]; g_bptPool[sc->device] = nullptr; }
    }
    if (!w) w = new BptWorkspace();
    struct Release {
        int device; BptWorkspace* w;
        ~Release() {
            std::lock_guard<std::mutex> lock(g_bptMutex);
            if (device >= 0 && device < 64 && !g_bptPool[device]) g_bptPool[device] = w;
            else { w->release(); delete w; }
        }
    } release{sc->device, w};
    const size_t vertWords = (size_t)2 * kBptMaxVerts * bptVertexWords<NC>() * slots;
    if (w->store.lanes != slots || w->vertWords != vertWords) {
        w->release();
        SLRGPU_CUDA_TRY(cudaMalloc(&w->store.verts, vertWords * sizeof(float4)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->store.mis, (size_t)2 * kBptMaxVerts * slots * sizeof(float4)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->batch.common, (size_t)slots * sizeof(BptSlotCommon)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->batch.eyeEnd, (size_t)slots * sizeof(BptSlotEnd)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->batch.lightEnd, (size_t)slots * sizeof(BptSlotEnd)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->batch.connCount, (size_t)slots * sizeof(uint32_t)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->batch.connBase, ((size_t)slots + 1) * sizeof(uint32_t)));
        SLRGPU_CUDA_TRY(cudaMalloc(&w->dCounters, sizeof(BptCounters)));
        w->store.lanes = slots; w->vertWords = vertWords;
    }
    SLRGPU_CUDA_TRY(cudaMemsetAsync(w->dCounters, 0, sizeof(BptCounters), stream));
    cudaEvent_t ev0, ev1;
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev0));
    SLRGPU_CUDA_TRY(cudaEventCreate(&ev1));
    struct EventFree { cudaEvent_t a, b; ~EventFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } eventFree{ev0, ev1};
    SLRGPU_CUDA_TRY(cudaEventRecord(ev0, stream));
    const uint32_t connectGrid = residentGrid(connectKernel<NC>, kBptBlock, sc->numSMs, 1u << 20);
    unsigned long long launches = 0;
    for (unsigned long long first = 0; first < totalSamples; first += slots) {
        const uint32_t count = (uint32_t)std::min<unsigned long long>(slots, totalSamples - first);
        // thread i < slots: eye subpath of slot i, thread slots + i: its light subpath
        subpathKernel<NC><<<2u * slots / kBptBlock, kBptBlock, 0, stream>>>(sc->dev, rc, w->store, w->batch, accumDev, w->dCounters, first, count);
        countKernel<<<(count + 255u) / 256u, 256, 0, stream>>>(w->batch, count);
        scanKernel<<<1, 1024, 0, stream>>>(w->batch, count);
        connectKernel<NC><<<connectGrid, kBptBlock, 0, stream>>>(sc->dev, rc, w->store, w->batch, accumDev, w->dCounters, count);
        launches += 4;
    }
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaEventRecord(ev1, stream));
    SLRGPU_CUDA_TRY(cudaEventSynchronize(ev1));
    BptCounters c;
    SLRGPU_CUDA_TRY(cudaMemcpy(&c, w->dCounters, sizeof(c), cudaMemcpyDeviceToHost));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->paths = totalSamples;
        stats->extend_rays = c.extendRays; stats->shadow_rays = c.shadowRays;
        stats->rays = c.extendRays + c.shadowRays;
        stats->kernel_launches = launches;
        stats->waves = launches / 4;            // batches
        stats->tail_paths = c.truncated;        // subpaths cut at kBptMaxVerts vertices
        stats->class_hits[8] = c.connections;   // (s, t) pairs examined
        cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
    }
    if (c.stackOverflow) { setError("traversal stack overflow (more than %d entries)", 64); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}

// slrgpu_render* with SLRGPU_RENDER_BPT in params->flags (render.cu renderImpl)
int renderBpt(SlrGpuScene* sc, const SlrGpuRenderParams* p, const RenderConstants& rc, float* accumDev, cudaStream_t stream, SlrGpuRenderStats* stats) {
    const unsigned long long totalSamples = (unsigned long long)p->width * p->height * (p->spp_end - p->spp_begin);
    if (sc->channels == 3) return renderBptT<3>(sc, rc, totalSamples, accumDev, stream, stats);
    return renderBptT<16>(sc, rc, totalSamples, accumDev, stream, stats);
}

// ---------------------------------------------------------------------------------------------
// shading probe for the bidirectional path tracer (slrgpu_probe_shading_bpt): the queries of generateSubPath and of a
// connection on caller-given inputs -- BSDF::sample with result->reverse, BSDF::evaluate, BSDF::evaluatePDF with revPDF,
// for radiance (even probes) and importance (odd probes: adjoint = true) transport
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
probeBptKernel(const DeviceScene s, const float* __restrict__ probes, uint32_t n, float* __restrict__ out, uint32_t* stackOverflow) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = probes + (size_t)i * SLRGPU_PROBE_IN_FLOATS;
    float* o = out + (size_t)i * SLRGPU_PROBE_OUT_FLOATS;
    for (int k = 0; k < SLRGPU_PROBE_OUT_FLOATS; ++k) o[k] = 0.0f;
    const V3 org(p[0], p[1], p[2]), dir(p[3], p[4], p[5]);
    uint32_t stack[kStackSize];
    bool overflow = false;
    WalkState w;
    w.r.ox = org.x; w.r.oy = org.y; w.r.oz = org.z; w.r.tmin = 0.0f;
    w.r.dx = dir.x; w.r.dy = dir.y; w.r.dz = dir.z; w.r.tmax = INFINITY;
    w.time = 0.0f;
    singleRayWalk<false>(s, w, stack, &overflow);
    if (overflow) atomicExch(stackOverflow, 1u);
    if (w.hit.prim == SLRGPU_INVALID_ID) { o[0] = s.envPresent ? 2.0f : 0.0f; return; }
    const float wlOffset = p[6];
    const uint32_t hero = min((uint32_t)(16 * p[7]), 15u);
    SurfPt sp;
    float localArea;
    const SlrGpuTriangle tri = hitSurfacePoint(s, w.hit.prim, w.hit.inst, w.hit.t, w.hit.u, w.hit.v, org, dir, 0.0f, &sp, &localArea);
    o[0] = 1.0f; o[1] = w.hit.t;
    Bsdf<16, 4> bsdf;
    buildBsdf<16, 4>(s, tri.material, sp, wlOffset, false, &bsdf);
    BsdfQuery q;
    q.dir = sp.sf.toLocal(-dir); q.gn = sp.sf.toLocal(sp.gn); q.hero = hero; q.flags = DT_All; q.adjoint = (i & 1u) != 0;
    BsdfSampleResult res;
    BsdfRev<16> rev;
    const Spec<16> fs = bsdfSampleRev(bsdf, q, p[8], p[9], p[10], &res, &rev);
    for (int k = 0; k < 16; ++k) o[2 + k] = fs.v[k];
    o[18] = res.dir.x; o[19] = res.dir.y; o[20] = res.dir.z;
    o[21] = res.pdf; o[22] = (float)res.type;
    const bool sampled = !specIsZero(fs) && res.pdf != 0.0f;       // the reference leaves reverse untouched otherwise
    for (int k = 0; k < 16; ++k) o[23 + k] = sampled ? rev.fs.v[k] : 0.0f;
    o[39] = sampled ? rev.pdf : 0.0f;
    const V3 evalDir = sp.sf.toLocal(V3(p[11], p[12], p[13]));
    float revPdf;
    o[40] = bsdfPdfRev(bsdf, q, evalDir, &revPdf);
    o[41] = revPdf;
    const Spec<16> fe = bsdfEvaluate(bsdf, q, evalDir);
    for (int k = 0; k < 16; ++k) o[42 + k] = fe.v[k];
}

}  // namespace slrgpu

extern "C" SLRGPU_API int slrgpu_probe_shading_bpt(SlrGpuScene* sc, const float* probes, uint64_t n, float* out) {
    using namespace slrgpu;
    if (!sc || !probes || !out) { setError("slrgpu_probe_shading_bpt: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (!sc->hasShading || sc->channels != 16) { setError("slrgpu_probe_shading_bpt: needs a spectral scene with materials"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    if (n == 0) return SLRGPU_OK;
    if (n > (1u << 26)) { setError("slrgpu_probe_shading_bpt: too many probes"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(sc->device));
    struct Buffers { void* p[3] = {}; ~Buffers() { for (void* q : p) if (q) cudaFree(q); } } bufs;
    SLRGPU_CUDA_TRY(cudaMalloc(&bufs.p[0], n * SLRGPU_PROBE_IN_FLOATS * sizeof(float)));
    SLRGPU_CUDA_TRY(cudaMalloc(&bufs.p[1], n * SLRGPU_PROBE_OUT_FLOATS * sizeof(float)));
    SLRGPU_CUDA_TRY(cudaMalloc(&bufs.p[2], sizeof(uint32_t)));
    SLRGPU_CUDA_TRY(cudaMemcpy(bufs.p[0], probes, n * SLRGPU_PROBE_IN_FLOATS * sizeof(float), cudaMemcpyHostToDevice));
    SLRGPU_CUDA_TRY(cudaMemset(bufs.p[2], 0, sizeof(uint32_t)));
    const uint32_t n32 = (uint32_t)n;
    probeBptKernel<<<(n32 + 63) / 64, 64>>>(sc->dev, (const float*)bufs.p[0], n32, (float*)bufs.p[1], (uint32_t*)bufs.p[2]);
    SLRGPU_CUDA_TRY(cudaGetLastError());
    SLRGPU_CUDA_TRY(cudaDeviceSynchronize());
    SLRGPU_CUDA_TRY(cudaMemcpy(out, bufs.p[1], n * SLRGPU_PROBE_OUT_FLOATS * sizeof(float), cudaMemcpyDeviceToHost));
    uint32_t overflow = 0;
    SLRGPU_CUDA_TRY(cudaMemcpy(&overflow, bufs.p[2], sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (overflow) { setError("traversal stack overflow (more than %d entries)", 64); return SLRGPU_ERR_STACK_OVERFLOW; }
    return SLRGPU_OK;
}
