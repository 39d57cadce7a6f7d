// The shade stages of a wave as device functions: `surface` (emission seen by the arriving ray, Russian
// roulette, sort into material classes) and `material` (BSDF at the hit, next event estimation, BSDF
// sampling). render.cu wraps them into the grid-stride kernels of the wavefront loop; tail.cu calls the
// same functions from the persistent kernel that finishes the last long paths of a frame.
//   Job::contribution                PathTracingRenderer.cpp:137-261
#pragma once
#include "rng.cuh"
#include "shade.cuh"
#include "wavefront.cuh"

// material stage: next iteration's class-queue entry loaded ahead and its path state hinted into L2 (materialStage).
// Measured (profiles/r02_variant_sweep.md): material 7.19 -> 6.72 ms per C1 frame, C1 617 -> 625-632 Mpaths/s.
#ifndef SLR_STAGE_PREFETCH
#define SLR_STAGE_PREFETCH 1
#endif

namespace slrgpu {

// positions of the `alive` lanes in the next path queue and of the `shadow` lanes in the shadow queue: both counters
// sit in one 64-bit word (WavefrontCounters::numNext / numShadow), so one atomic per warp reserves both ranges
__device__ __forceinline__ void warpAppendPair(bool alive, bool shadow, WavefrontCounters* counters, uint32_t* npos, uint32_t* spos) {
    const unsigned am = __ballot_sync(0xFFFFFFFFu, alive), sm = __ballot_sync(0xFFFFFFFFu, shadow);
    *npos = 0; *spos = 0;
    if ((am | sm) == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0)
        base = atomicAdd(reinterpret_cast<unsigned long long*>(&counters->numNext),
                         (unsigned long long)__popc(am) | ((unsigned long long)__popc(sm) << 32));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    const unsigned lt = (1u << lane) - 1u;
    *npos = (uint32_t)(base & 0xFFFFFFFFull) + __popc(am & lt);
    *spos = (uint32_t)(base >> 32) + __popc(sm & lt);
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

// Emission seen by the ray that arrived at an emitter (or left the scene into the environment), with the
// MIS weight of implicit light sampling (PathTracingRenderer.cpp:152-156, 232-249). Kept out of line:
// only a few per cent of the hits take it, and inlined it doubles the register count of the surface kernel.
template <int NC>
static __device__ __noinline__ void surfaceEmission(const DeviceScene& s, const PathQueue& in, const HitBuffer& hits, uint32_t i, uint4 meta,
                                                    uint32_t flags, bool isEnv, float* __restrict__ accum) {
    const bool cameraRay = flags & kFlagCameraRay;
    // a camera ray's throughput is 1: ray generation does not store it
    const Spec<NC> alpha = cameraRay ? specConst<NC>(1.0f) : loadAlpha<NC>(in, i);
    const float4 o4 = in.org[i], d4 = in.dir[i];
    const V3 org(o4.x, o4.y, o4.z), dir(d4.x, d4.y, d4.z);
    const float prevPdf = d4.w;
    const float wlOffset = __uint_as_float(meta.w);
    SurfPt sp;
    float localArea = 1.0f;
    uint32_t material = s.envMaterial;
    uint2 hid = make_uint2(SLRGPU_INVALID_ID, SLRGPU_INVALID_ID);
    SlrGpuTriangle tri = {};
    if (isEnv) envSurfacePoint(dir, &sp);
    else {
        hid = hits.id[i];
        const float4 htuv = hits.tuv[i];
        tri = hitSurfacePoint(s, hid.x, hid.y, htuv.x, htuv.y, htuv.z, org, dir, in.time ? in.time[i] : 0.0f, &sp, &localArea);
        material = tri.material;
    }
    const V3 dirOut = sp.sf.toLocal(-dir);
    // DiffuseEDF: 1/pi on the front side; IBLEDF: 1/pi
    const float edf = (sp.atInfinity || dirOut.z > 0.0f) ? 1.0f / kPi : 0.0f;
    if (edf <= 0.0f) return;
    float mis = 1.0f;
    if (!cameraRay && !(flags & kFlagPrevDelta)) {
        const float lightProb = lightSelectionProb(s, tri, hid.y, sp.atInfinity);
        float areaPDF, dist2;
        if (sp.atInfinity) { areaPDF = envEvaluateUVPDF(s, sp.u / (2 * kPi), sp.v / kPi) / (2 * kPi * kPi * sinf(sp.v)); dist2 = 1.0f; }
        else { areaPDF = 1.0f / localArea; dist2 = sqLength(sp.p - org); }
        const float lightPDF = lightProb * areaPDF * dist2 / absDot(dir, sp.gn);
        mis = (prevPdf * prevPdf) / (lightPDF * lightPDF + prevPdf * prevPdf);
    }
    const Spec<NC> Le = materialEmittance<NC>(s, material, sp, wlOffset);
    float v[NC == 3 ? 4 : NC];
    const float k = edf * mis * in.weight[i];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { v[c] = alpha.v[c] * Le.v[c] * k; sum += v[c]; }
    if (!isfinite(sum)) return;          // see materialItem: a non-finite sample is dropped, never splatted
    splat<NC>(accum, meta.x, wlOffset, (flags & kFlagStrataInPlace) != 0, v);
}

// ---------------------------------------------------------------------------------------------
// surface: what Job::contribution does between a hit and the BSDF of that hit -- emission seen by the
// ray that arrived (implicit light sampling with MIS), the environment for rays that left the scene,
// Russian roulette and the path-length cap (PathTracingRenderer.cpp:147-163, 225-258). Survivors are
// sorted into one queue per material class.
// What the stage needs to know about the hit triangle (class | emitting | leaf material) arrives with the hit record:
// the traversal copies the spare word of the leaf record it accepted (device_scene.h packSurfaceInfo) into tuv.w, so an
// entry costs three streaming loads (hit, meta, roulette slot) and no dependent fetch of the triangle / material tables.
// ---------------------------------------------------------------------------------------------
// One path-queue entry: *cls = the material class of the surviving hit (SC_NONE: the path ended here), *leaf = its
// leaf material.
// The three streaming loads of an entry (hit record, meta, roulette slot), issued together before any of them is used.
struct SurfaceInput { uint32_t info; uint4 meta; float aux; };
__device__ __forceinline__ SurfaceInput surfaceLoad(const PathQueue& in, const HitBuffer& hits, uint32_t i) {
    SurfaceInput x;
    x.info = __float_as_uint(hits.tuv[i].w);
    x.meta = in.meta[i];
    x.aux = in.aux[i];          // importance(alpha) left by the material kernel; unused (and unwritten) for a camera ray
    return x;
}

// emitOut == nullptr: the emission seen by the arriving ray is added here (the tail kernel); else *emitOut says whether the
// entry sees emission (1 = an emitter was hit, 2 = the environment) and the caller queues it for emissionKernel.
template <int NC>
__device__ __forceinline__ void surfaceItem(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                            float* __restrict__ accum, uint32_t i, const SurfaceInput& x, uint32_t* clsOut, uint32_t* leafOut,
                                            uint32_t* emitOut = nullptr) {
    uint32_t cls = SC_NONE, leaf = SLRGPU_INVALID_ID;
    uint32_t info = x.info;
    const uint4 meta = x.meta;
    const uint32_t hero = meta.z & 0xFFu;
    const uint32_t flags = (meta.z >> 8) & 0xFFu;
    uint32_t pathLength = meta.z >> 16;
    const bool cameraRay = flags & kFlagCameraRay;
    const bool isEnv = info == kSurfaceInfoMiss;
    if (!isEnv || s.envPresent) {
        bool emitting = true;
        if (!isEnv) {
            if (info == kSurfaceInfoDynamic) {          // not precomputed (more than 2^23 materials): work it out from the tables
                const uint32_t material = s.triangles[hits.id[i].x].material;
                uint32_t lf = 0;
                const uint32_t c = classifyMaterial(s, material, &lf);
                info = c | ((materialIsEmitting(s, material) ? 1u : 0u) << 8) | ((c == 0xFFu ? 0u : lf) << 9);
            }
            emitting = ((info >> 8) & 1u) != 0;
        }
        if (emitting) {
            if (emitOut) *emitOut = isEnv ? 2u : 1u;
            else surfaceEmission<NC>(s, in, hits, i, meta, flags, isEnv, accum);
        }
        bool cont = !isEnv;
        if (cont && !cameraRay) {
            // Russian roulette; initY = importance of a unit spectrum = 1. importance(alpha) was left in
            // aux by the material kernel that produced this entry; the surviving path's 1/q goes back
            // into aux and is applied to alpha by the material kernel of this bounce.
            const float continueProb = fminf(x.aux, 1.0f);
            const Rand4 rr = pathRandom(rc.seed, meta.x, meta.y, 2 * pathLength + 1);   // .z of the previous bounce's second block
            if (rr.z < continueProb) in.aux[i] = 1.0f / continueProb;
            else cont = false;
        }
        if (cont) {
            ++pathLength;
            if (pathLength >= rc.maxPathLength) cont = false;
        }
        if (cont) {
            cls = info & 0xFFu; leaf = info >> 9;
            if (cls != SC_NONE) in.meta[i].z = hero | (flags << 8) | (pathLength << 16);
        }
    }
    *clsOut = cls; *leafOut = leaf;
}
template <int NC>
__device__ __forceinline__ void surfaceItem(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                            float* __restrict__ accum, uint32_t i, uint32_t* clsOut, uint32_t* leafOut, uint32_t* emitOut = nullptr) {
    surfaceItem<NC>(s, rc, in, hits, accum, i, surfaceLoad(in, hits, i), clsOut, leafOut, emitOut);
}

// append to the class queues: one atomic per (warp, class)
__device__ __forceinline__ void classAppend(const ClassQueue& cq, WavefrontCounters* counters, uint32_t lane, uint32_t i, uint32_t cls, uint32_t leaf) {
    const unsigned active = __ballot_sync(0xFFFFFFFFu, cls != SC_NONE);
    if (cls != SC_NONE) {
        const unsigned grp = __match_any_sync(active, cls);
        const int leader = __ffs(grp) - 1;
        uint32_t pos = 0;
        if ((int)lane == leader) pos = atomicAdd(&counters->classCount[cls], (uint32_t)__popc(grp));
        pos = __shfl_sync(grp, pos, leader) + __popc(grp & ((1u << lane) - 1u));
        cq.entries[(size_t)cls * cq.capacity + pos] = make_uint2(i, leaf);
    }
}

// Work items [0, n) are spread over the whole grid.
// (Two groups per iteration with both groups' loads issued up front were measured in round 2 and dropped: surface
// 2.54 -> 2.85 ms per C1 frame -- the extra registers cost more than the second set of loads in flight buys; an L2 hint
// for the next iteration's three loads changed nothing, 2.78 vs 2.80 ms -- profiles/r02_rejected_experiments.md.)
// SLR_SURFACE_APPEND: 0 = one atomic per (warp, class) and 32 entries, emission added inline (round 1); 2 = one per (warp,
// class) and kSurfaceGroups x 32 entries, entries that see emission queued for emissionKernel. (1 = one per block and
// class with the counts meeting in shared memory behind two barriers was measured and dropped: profiles/r02_variant_sweep.md.)
#ifndef SLR_SURFACE_APPEND
#define SLR_SURFACE_APPEND 2
#endif
constexpr int kSurfaceGroups = 4;

template <int NC, int BLOCK>
__device__ __forceinline__ void surfaceStage(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                             const ClassQueue& cq, float* __restrict__ accum, WavefrontCounters* counters, uint32_t n) {
    const uint32_t stride = gridDim.x * blockDim.x;
#if SLR_SURFACE_APPEND == 2
    // A warp takes kSurfaceGroups x 32 consecutive entries per iteration and reserves their places in the class queues with
    // ONE atomic per class: the groups' per-class counts are summed in a 16-word table of shared memory owned by the warp
    // (shared-memory atomics hand every group its offset inside the warp's range), lanes 0-15 then reserve one class each.
    __shared__ uint32_t table[BLOCK / 32][16];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mine = table[warp];
    for (uint32_t base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * kSurfaceGroups; base < n; base += stride * kSurfaceGroups) {
        if (lane < 16) mine[lane] = 0;
        __syncwarp();
        uint32_t cls[kSurfaceGroups], leaf[kSurfaceGroups], off[kSurfaceGroups], emit[kSurfaceGroups];      // emit: kind << 28 | offset
#pragma unroll 1
        for (int g = 0; g < kSurfaceGroups; ++g) {
            const uint32_t i = base + g * 32 + lane;
            uint32_t c = SC_NONE, l = SLRGPU_INVALID_ID, o = 0, em = 0;
            if (i < n) surfaceItem<NC>(s, rc, in, hits, accum, i, &c, &l, &em);
            const unsigned active = __ballot_sync(0xFFFFFFFFu, c != SC_NONE);
            if (c != SC_NONE) {
                const unsigned grp = __match_any_sync(active, c);
                const int leader = __ffs(grp) - 1;
                if ((int)lane == leader) o = atomicAdd(&mine[c], (uint32_t)__popc(grp));       // this group's offset in the warp's range
                o = __shfl_sync(grp, o, leader) + __popc(grp & ((1u << lane) - 1u));
            }
            // the entries that see emission: one more row of the class queues (kEmissionRow)
            const unsigned seen = __ballot_sync(0xFFFFFFFFu, em != 0);
            if (seen) {
                const int leader = __ffs(seen) - 1;
                uint32_t eo = 0;
                if ((int)lane == leader) eo = atomicAdd(&mine[kEmissionRow], (uint32_t)__popc(seen));
                eo = __shfl_sync(0xFFFFFFFFu, eo, leader) + __popc(seen & ((1u << lane) - 1u));
                if (em) em = (em << 28) | eo;
            }
            cls[g] = c; leaf[g] = l; off[g] = o; emit[g] = em;
        }
        __syncwarp();
        if (lane < 16) { const uint32_t total = mine[lane]; if (total) mine[lane] = atomicAdd(&counters->classCount[lane], total); }
        __syncwarp();
#pragma unroll 1
        for (int g = 0; g < kSurfaceGroups; ++g) {
            if (cls[g] != SC_NONE) cq.entries[(size_t)cls[g] * cq.capacity + mine[cls[g]] + off[g]] = make_uint2(base + g * 32 + lane, leaf[g]);
            if (emit[g]) cq.entries[(size_t)kEmissionRow * cq.capacity + mine[kEmissionRow] + (emit[g] & 0x0FFFFFFFu)] = make_uint2(base + g * 32 + lane, emit[g] >> 28);
        }
        __syncwarp();
    }
#else
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const uint32_t i = base + lane;
        uint32_t cls = SC_NONE, leaf = SLRGPU_INVALID_ID;
        if (i < n) surfaceItem<NC>(s, rc, in, hits, accum, i, &cls, &leaf);
        classAppend(cq, counters, lane, i, cls, leaf);
    }
#endif
}

// the entries of a wave that see emission (queued by the surface stage), one per thread
template <int NC>
__device__ __forceinline__ void emissionStage(const DeviceScene& s, const PathQueue& in, const HitBuffer& hits, const ClassQueue& cq,
                                              float* __restrict__ accum, uint32_t n) {
    const uint2* __restrict__ entries = cq.entries + (size_t)kEmissionRow * cq.capacity;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint2 e = entries[k];
        const uint4 meta = in.meta[e.x];
        surfaceEmission<NC>(s, in, hits, e.x, meta, (meta.z >> 8) & 0xFFu, e.y == 2u, accum);
    }
}

// ---------------------------------------------------------------------------------------------
// material: for every survivor of one class -- BSDF at the hit, next event estimation (the shadow ray
// goes to the shadow queue with its MIS-weighted contribution), BSDF sampling of the next direction
// (PathTracingRenderer.cpp:164-222). CLASS < SC_GENERIC: one lobe of compile-time type; SC_GENERIC:
// the tagged multi-lobe BSDF.
// ---------------------------------------------------------------------------------------------
template <int NC, int CLASS> struct HitBsdf {
    Lobe<NC> lobe;
    __device__ __forceinline__ void build(const DeviceScene& s, uint32_t leaf, const SurfPt& sp, float wlOffset, bool lambdaSelected) {
        fillLobeT<NC, classMaterialKind(CLASS)>(s, s.materials[leaf], sp, wlOffset, lambdaSelected, 1.0f, 0u, &lobe);
    }
    __device__ __forceinline__ bool hasNonDelta() const { return dtMatches(lobe.baseDirType, DT_WholeSphere | DT_NonDelta); }
    __device__ __forceinline__ Spec<NC> evaluate(const BsdfQuery& q, const V3& d) const { return lobeEvaluate<NC, CLASS>(lobe, q, d); }
    __device__ __forceinline__ float pdf(const BsdfQuery& q, const V3& d) const { return lobePdf<NC, CLASS>(lobe, q, d); }
    __device__ __forceinline__ Spec<NC> sample(const BsdfQuery& q, float uc, float u0, float u1, BsdfSampleResult* r) const { return lobeSample<NC, CLASS>(lobe, q, uc, u0, u1, r); }
};
template <int NC> struct HitBsdf<NC, SC_GENERIC> {
    Bsdf<NC, 4> bsdf;
    __device__ __forceinline__ void build(const DeviceScene& s, uint32_t leaf, const SurfPt& sp, float wlOffset, bool lambdaSelected) {
        buildBsdf<NC, 4>(s, leaf, sp, wlOffset, lambdaSelected, &bsdf);
    }
    __device__ __forceinline__ bool hasNonDelta() const { return bsdfHasNonDelta(bsdf); }
    __device__ __forceinline__ Spec<NC> evaluate(const BsdfQuery& q, const V3& d) const { return bsdfEvaluate(bsdf, q, d); }
    __device__ __forceinline__ float pdf(const BsdfQuery& q, const V3& d) const { return bsdfPdf(bsdf, q, d); }
    __device__ __forceinline__ Spec<NC> sample(const BsdfQuery& q, float uc, float u0, float u1, BsdfSampleResult* r) const { return bsdfSample(bsdf, q, uc, u0, u1, r); }
};

// What one hit leaves behind: the continued path (if `alive`) and the shadow ray of its light sample (if `shadow`).
template <int NC> struct MaterialResult {
    bool alive, shadow;
    V3 nOrg, nDir;
    float nPdf, nImp;
    uint4 meta;
    float weight;
    Spec<NC> alpha;
    V3 sOrg, sDir;
    float sTmax;
    float time;                 // the path's ray time, handed on to the continued path and to the shadow ray
    Spec<NC> sContrib;
    __device__ __forceinline__ void clear() {
        alive = false; shadow = false;
        nOrg = V3(0, 0, 0); nDir = V3(0, 0, 1); nPdf = 0.0f; nImp = 0.0f;
        meta = make_uint4(0, 0, 0, 0); weight = 0.0f; alpha = specConst<NC>(0.0f);
        sOrg = V3(0, 0, 0); sDir = V3(0, 0, 1); sTmax = 0.0f; time = 0.0f; sContrib = specConst<NC>(0.0f);
    }
};

// path-queue entry i, whose hit has the leaf material `leaf` of class CLASS
template <int NC, int CLASS>
__device__ __forceinline__ void materialItem(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                             uint32_t i, uint32_t leaf, MaterialResult<NC>& o) {
    const float4 o4 = in.org[i], d4 = in.dir[i];
    o.meta = in.meta[i];
    o.weight = in.weight[i];
    // Russian-roulette scale decided by `surface`; a camera ray carries throughput 1 and ray generation stores neither
    o.alpha = ((o.meta.z >> 8) & kFlagCameraRay) ? specConst<NC>(1.0f) : loadAlpha<NC>(in, i) * in.aux[i];
    const uint2 hid = hits.id[i];
    const float4 htuv = hits.tuv[i];
    const V3 org(o4.x, o4.y, o4.z), dir(d4.x, d4.y, d4.z);
    const uint32_t hero = o.meta.z & 0xFFu;
    uint32_t flags = (o.meta.z >> 8) & 0xFFu;
    const uint32_t pathLength = o.meta.z >> 16;
    const float wlOffset = __uint_as_float(o.meta.w);

    SurfPt sp;
    float localArea;
    o.time = in.time ? in.time[i] : 0.0f;
    hitSurfacePoint(s, hid.x, hid.y, htuv.x, htuv.y, htuv.z, org, dir, o.time, &sp, &localArea);
    const V3 dirOut = sp.sf.toLocal(-dir);
    const V3 gNorm = sp.sf.toLocal(sp.gn);
    HitBsdf<NC, CLASS> bsdf;
    bsdf.build(s, leaf, sp, wlOffset, (flags & kFlagLambdaSelected) != 0);
    BsdfQuery q;
    q.dir = dirOut; q.gn = gNorm; q.hero = hero; q.flags = DT_All;
    const Rand4 ra = pathRandom(rc.seed, o.meta.x, o.meta.y, 2 * pathLength);       // light select, light u0, u1, bsdf component
    const Rand4 rb = pathRandom(rc.seed, o.meta.x, o.meta.y, 2 * pathLength + 1);   // bsdf u0, u1

    // next event estimation
    if (bsdf.hasNonDelta() && (s.numTopLights > 0 || s.envPresent)) {
        LightSample ls;
        sampleLight(s, ra.x, ra.y, ra.z, o.time, &ls);
        float dist2;
        V3 shadowDir;
        if (ls.sp.atInfinity) { dist2 = 1.0f; shadowDir = normalize(ls.sp.p); }
        else { const V3 d = ls.sp.p - sp.p; dist2 = sqLength(d); shadowDir = d / sqrtf(dist2); }
        const V3 shadowDir_l = ls.sp.sf.toLocal(-shadowDir);
        const V3 shadowDir_sn = sp.sf.toLocal(shadowDir);
        const float edf = (ls.isEnv || shadowDir_l.z > 0.0f) ? 1.0f / kPi : 0.0f;
        if (edf > 0.0f && ls.areaPDF > 0.0f) {
            const Spec<NC> fs = bsdf.evaluate(q, shadowDir_sn);
            if (!specIsZero(fs)) {
                const Spec<NC> M = materialEmittance<NC>(s, ls.material, ls.sp, wlOffset);
                const float cosLight = absDot(-shadowDir, ls.sp.gn);
                const float bsdfPDF = bsdf.pdf(q, shadowDir_sn) * cosLight / dist2;
                float mis = 1.0f;
                if (!isinf(ls.areaPDF)) mis = (ls.lightPDF * ls.lightPDF) / (ls.lightPDF * ls.lightPDF + bsdfPDF * bsdfPDF);
                const float G = absDot(shadowDir_sn, gNorm) * cosLight / dist2;
                const float kk = edf * (G * mis / ls.lightPDF) * o.weight;
                float sum = 0.0f;
#pragma unroll
                for (int c = 0; c < NC; ++c) { o.sContrib.v[c] = o.alpha.v[c] * M.v[c] * fs.v[c] * kk; sum += o.sContrib.v[c]; }
                // Scene::testVisibility
                o.sOrg = sp.p;
                if (ls.sp.atInfinity) { o.sDir = shadowDir; o.sTmax = 3.402823466e+38f; }
                else { const float dist = length(ls.sp.p - sp.p); o.sDir = (ls.sp.p - sp.p) / dist; o.sTmax = dist * (1.0f - 0.0001f); }
                // The shading TUs divide with the 2-ulp hardware reciprocal (Makefile SHADE_MATH): a denormal denominator
                // gives inf where IEEE division gives a huge finite value (seen once per ~1e8 samples on the microfacet
                // scene). A sample that is not finite is dropped here / the path ended below, so a NaN never reaches the
                // sensor; the reference would have splatted a firefly.
                o.shadow = isfinite(sum);
            }
        }
    }

    // sample the BSDF for the next direction
    BsdfSampleResult res;
    const Spec<NC> fs = bsdf.sample(q, ra.w, rb.x, rb.y, &res);
    if (!specIsZero(fs) && res.pdf != 0.0f) {
        float dirPDF = res.pdf;
        if (res.type & DT_Dispersive) { dirPDF /= NC; flags |= kFlagLambdaSelected; }
        const float kk = absDot(res.dir, gNorm) / dirPDF;
        o.alpha = o.alpha * (fs * kk);
        o.nOrg = sp.p;
        o.nDir = sp.sf.fromLocal(res.dir);
        o.nPdf = dirPDF;
        flags &= ~(kFlagCameraRay | kFlagPrevDelta);
        if (dtIsDelta(res.type)) flags |= kFlagPrevDelta;
        o.meta.z = hero | (flags << 8) | (pathLength << 16);
        o.nImp = specImportance(o.alpha, hero);
        o.alive = isfinite(o.nImp);
    }
}

// the shadow ray of a light sample goes to position spos of the shadow queue ...
template <int NC>
__device__ __forceinline__ void materialWriteShadow(const ShadowQueue& sq, uint32_t spos, const MaterialResult<NC>& o) {
    sq.org[spos] = make_float4(o.sOrg.x, o.sOrg.y, o.sOrg.z, 0.0001f);
    sq.dir[spos] = make_float4(o.sDir.x, o.sDir.y, o.sDir.z, o.sTmax);
    const bool inPlace = ((o.meta.z >> 8) & kFlagStrataInPlace) != 0;
    sq.pixelWl[spos] = make_uint2(o.meta.x | (inPlace ? 0x80000000u : 0u), o.meta.w);
    if (sq.time) sq.time[spos] = o.time;
    if (NC == 3) sq.contrib[spos] = make_float4(o.sContrib.v[0], o.sContrib.v[1], o.sContrib.v[2], 0.0f);
    else {
#pragma unroll
        for (int c = 0; c < NC / 4; ++c)
            sq.contrib[(size_t)c * sq.capacity + spos] =
                make_float4(o.sContrib.v[4 * c], o.sContrib.v[(4 * c + 1) % NC], o.sContrib.v[(4 * c + 2) % NC], o.sContrib.v[(4 * c + 3) % NC]);
    }
}
// ... the continued path to position npos of the next path queue
template <int NC>
__device__ __forceinline__ void materialWriteNext(const PathQueue& out, uint32_t npos, const MaterialResult<NC>& o) {
    out.org[npos] = make_float4(o.nOrg.x, o.nOrg.y, o.nOrg.z, 0.0001f);      // Ray::Epsilon
    out.dir[npos] = make_float4(o.nDir.x, o.nDir.y, o.nDir.z, o.nPdf);
    out.meta[npos] = o.meta;
    out.weight[npos] = o.weight;
    out.aux[npos] = o.nImp;
    if (out.time) out.time[npos] = o.time;
    storeAlpha<NC>(out, npos, o.alpha);
}

// (Measured and dropped, round 2: storing the shadow entry where the light sample is complete -- its 23 values are then
// dead while the BSDF is sampled -- through an atomic of the lanes executing together: material 6.69 vs 6.72 ms per C1
// frame at 3 resident blocks, 7.2 ms at 4 (128 registers, 40 bytes of spills), C2 1.5 % slower.)
template <int NC, int CLASS>
__device__ __forceinline__ void materialStage(const DeviceScene& s, const RenderConstants& rc, const PathQueue& in, const HitBuffer& hits,
                                              const ClassQueue& cq, const PathQueue& out, const ShadowQueue& sq, WavefrontCounters* counters, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint2* __restrict__ entries = cq.entries + (size_t)CLASS * cq.capacity;
#if SLR_STAGE_PREFETCH
    // The class-queue entry of the NEXT iteration is loaded one iteration ahead (two registers), and the path state / hit
    // record it points at is hinted into L2 while this iteration's hit is shaded: the entry -> state indirection otherwise
    // costs two serial DRAM round trips at the top of every iteration (ncu: 15 % of the kernel's stall samples).
    const uint32_t first = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    uint2 eNext = make_uint2(0, 0);
    if (first + lane < n) eNext = entries[first + lane];
    for (uint32_t base = first; base < n; base += stride) {
        const uint32_t k = base + lane;
        const uint2 e = eNext;
        const bool more = k + stride < n;
        if (more) eNext = entries[k + stride];
        MaterialResult<NC> o;
        o.clear();
        if (k < n) materialItem<NC, CLASS>(s, rc, in, hits, e.x, e.y, o);
        if (more) {
            const uint32_t j = eNext.x;
            prefetchL2(in.org + j); prefetchL2(in.dir + j); prefetchL2(in.meta + j); prefetchL2(hits.tuv + j);
            prefetchL2(hits.id + j); prefetchL2(in.weight + j); prefetchL2(in.aux + j);
            if (!((o.meta.z >> 8) & kFlagCameraRay)) {      // this entry carried a throughput: the next one most likely does too
#pragma unroll
                for (int c = 0; c < (NC == 3 ? 1 : NC / 4); ++c) prefetchL2(in.alpha + (size_t)c * in.capacity + j);
            }
        }
        uint32_t npos, spos;
        warpAppendPair(o.alive, o.shadow, counters, &npos, &spos);
        if (o.shadow) materialWriteShadow<NC>(sq, spos, o);
        if (o.alive) materialWriteNext<NC>(out, npos, o);
    }
#else
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const uint32_t k = base + lane;
        MaterialResult<NC> o;
        o.clear();
        if (k < n) {
            const uint2 e = entries[k];
            materialItem<NC, CLASS>(s, rc, in, hits, e.x, e.y, o);
        }
        uint32_t npos, spos;
        warpAppendPair(o.alive, o.shadow, counters, &npos, &spos);
        if (o.shadow) materialWriteShadow<NC>(sq, spos, o);
        if (o.alive) materialWriteNext<NC>(out, npos, o);
    }
#endif
}

}  // namespace slrgpu
