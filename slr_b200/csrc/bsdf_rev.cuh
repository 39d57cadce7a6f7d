// Reverse information of BSDF queries for bidirectional path tracing: next to the value / pdf of a query (dir_in -> dir_out)
// the reference returns the value and pdf of the opposite transport direction (dir_out -> dir_in, adjoint flipped), which
// the MIS weights of a connection need (BidirectionalPathTracingRenderer.cpp:184-196, 320-325).
//   BSDFReverseInfo / result->reverse / rev_fs / revPDF   libSLR/Core/directional_distribution_functions.h:129-138, 231-275
//   per-model reverse values                               libSLR/BSDFs/*.cpp (every sampleInternal / evaluateInternal /
//                                                          evaluatePDFInternal), MultiBSDF.cpp:66-121, 167-201
// What the reference's models do, read off those files:
//   * evaluate: rev_fs is the forward value for every model (symmetric BRDFs; the rough-glass value is computed once and
//     returned for both), so evaluate needs no reverse twin here -- bsdfEvaluate's result is both.
//   * evaluatePDF: revPDF is the model's pdf with the two directions exchanged. Lambert / Oren-Nayar / Ward write it out
//     (|cos| of the query direction / the same value), the two microfacet models too (with the forward query's reflection
//     probability) and Ashikhmin-Shirley (the forward specular pdf with the lobe weights of the other direction).
//   * sample: as evaluatePDF for the non-delta models; the two specular models return their selection probability and, for
//     refraction, a value without the radiance scaling.
// Only bpt.cu includes this header; the path tracer's kernels never ask for reverse values.
#pragma once
#include "bsdf.cuh"

namespace slrgpu {

template <int NC> struct BsdfRev {
    Spec<NC> fs;
    float pdf;
};

// MicrofacetBRDF / MicrofacetBSDF::evaluatePDFInternal with revPDF (MicrofacetBSDF.cpp:72-99, 254-304). Written out rather
// than taken from the exchanged pair: the reverse pdf keeps the FORWARD query's reflection probability F(dotHV).importance,
// which for the non-hero wavelengths of a refraction is not the one the exchanged query would compute.
template <int NC>
__device__ __forceinline__ float microfacetPdfRev(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir, float* revPdf) {
    *revPdf = 0.0f;
    const float alpha = L.f0;
    const bool entering = q.dir.z >= 0.0f;
    const float sign = entering ? 1.0f : -1.0f;
    const float dotNVdotNL = dir.z * q.dir.z;
    if (L.type == LOBE_MF_BRDF) {
        if (dotNVdotNL <= 0) return 0.0f;
        const V3 m = sign * halfVector(q.dir, dir);
        const float dotHV = dot(q.dir, m);
        if (dotHV * sign <= 0) return 0.0f;
        const float commonPDFTerm = 1.0f / (4 * dotHV * sign);
        *revPdf = commonPDFTerm * ggxPdfVisible(alpha, sign * dir, m);
        return commonPDFTerm * ggxPdfVisible(alpha, sign * q.dir, m);
    }
    if (dotNVdotNL == 0) return 0.0f;
    const float etaExtH = specAt(L.s0, q.hero), etaIntH = specAt(L.s1, q.hero);
    const float eEnter = entering ? etaExtH : etaIntH;
    const float eExit = entering ? etaIntH : etaExtH;
    V3 m;
    if (dotNVdotNL > 0) m = sign * halfVector(q.dir, dir);
    else m = normalize(-(eEnter * q.dir + eExit * dir));
    const float dotHV = dot(q.dir, m);
    if (dotHV * sign <= 0) return 0.0f;
    const float mPDF = ggxPdfVisible(alpha, sign * q.dir, m);
    const Spec<NC> F = fresnelDielectric(L.s0, L.s1, dotHV);
    float reflectProb = specImportance(F, q.hero);
    if (dtIsReflection(q.flags)) reflectProb = 1.0f;
    if (dtIsTransmission(q.flags)) reflectProb = 0.0f;
    if (dotNVdotNL > 0) {
        const float commonPDFTerm = reflectProb / (4 * dotHV * sign);
        *revPdf = commonPDFTerm * ggxPdfVisible(alpha, sign * dir, m);
        return commonPDFTerm * mPDF;
    }
    const float dotHL = dot(dir, m);
    const float commonPDFTerm = (1 - reflectProb) / pow2(eEnter * dotHV + eExit * dotHL);
    *revPdf = commonPDFTerm * ggxPdfVisible(alpha, -sign * dir, m) * eEnter * eEnter * fabsf(dotHV);
    return commonPDFTerm * mPDF * eExit * eExit * fabsf(dotHL);
}

// evaluatePDFInternal(query, dir, &revPDF) of a base lobe
template <int NC>
static __device__ __noinline__ float basePdfRev(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir, float* revPdf) {
    if (L.type == LOBE_MF_BRDF || L.type == LOBE_MF_BSDF) return microfacetPdfRev(L, q, dir, revPdf);
    const float pdf = basePdf(L, q, dir);
    switch (L.type) {
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR:
        *revPdf = (q.dir.z * dir.z <= 0.0f) ? 0.0f : fabsf(q.dir.z) / kPi;
        break;
    case LOBE_SPECULAR_BRDF:
    case LOBE_SPECULAR_BSDF:
        *revPdf = 0.0f;
        break;
    case LOBE_WARD:
        *revPdf = pdf;
        break;
    default: {      // Ashikhmin-Shirley (AshikhminShirleyBRDF.cpp:120-152): the same specular pdf, the diffuse pdf of the query
                    // direction, the lobe weights at `dir`
        if (dir.z * q.dir.z <= 0) { *revPdf = 0.0f; break; }
        const float nu = L.f0, nv = L.f1;
        const V3 halfv = halfVector(q.dir, dir);
        const float dotHV = dot(halfv, q.dir);
        const float ex = (nu * halfv.x * halfv.x + nv * halfv.y * halfv.y) / (1 - halfv.z * halfv.z);
        const float specularDirPDF = sqrtf((nu + 1) * (nv + 1)) / (8 * kPi * dotHV) * powf(fabsf(halfv.z), ex);
        BsdfQuery rq = q;
        rq.dir = dir;
        float revSpecularWeight, revDiffuseWeight;
        ashikhminWeights(L, rq, &revSpecularWeight, &revDiffuseWeight);
        *revPdf = (specularDirPDF * revSpecularWeight + fabsf(q.dir.z) / kPi * revDiffuseWeight) / (revSpecularWeight + revDiffuseWeight);
        break;
    }
    }
    return pdf;
}

// result->reverse of a base lobe's sampleInternal; called after baseSample returned `fs` with a non-zero pdf
template <int NC>
static __device__ __noinline__ void baseSampleReverse(const Lobe<NC>& L, const BsdfQuery& q, const BsdfSampleResult& res, const Spec<NC>& fs, BsdfRev<NC>* rev) {
    rev->fs = fs;
    switch (L.type) {
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR:
        rev->pdf = fabsf(q.dir.z) / kPi;
        break;
    case LOBE_SPECULAR_BRDF:
        rev->pdf = 1.0f;
        break;
    case LOBE_SPECULAR_BSDF:
        rev->pdf = res.pdf;          // reflectProb resp. 1 - reflectProb
        if (res.type & DT_Transmission) {
            // basic_BSDFs.cpp:139-147: the hero wavelength's coeff (1 - F) over |cos| of the QUERY direction, scaled by
            // (eExit / eEnter)^2 only for an adjoint query
            const bool entering = q.dir.z > 0.0f;
            const float etaExtH = specAt(L.s1, q.hero), etaIntH = specAt(L.s2, q.hero);
            const float eEnter = entering ? etaExtH : etaIntH;
            const float eExit = entering ? etaIntH : etaExtH;
            float v = specAt(L.s0, q.hero) * (1.0f - fresnelDielectric1(etaExtH, etaIntH, q.dir.z));
            v /= fabsf(q.dir.z);
            if (q.adjoint) v *= (eExit * eExit) / (eEnter * eEnter);
#pragma unroll
            for (int i = 0; i < NC; ++i) rev->fs.v[i] = (i == (int)q.hero) ? v : 0.0f;
        }
        break;
    case LOBE_WARD:
        rev->pdf = res.pdf;
        break;
    case LOBE_ASHIKHMIN: {
        // AshikhminShirleyBRDF.cpp:85-94: the same specular pdf, the diffuse pdf of the query direction, the lobe weights
        // at the sampled direction
        float specularDirPDF;
        ashikhminEval(L, q, res.dir, halfVector(q.dir, res.dir), &specularDirPDF);
        BsdfQuery rq = q;
        rq.dir = res.dir;
        float revSpecularWeight, revDiffuseWeight;
        ashikhminWeights(L, rq, &revSpecularWeight, &revDiffuseWeight);
        rev->pdf = (specularDirPDF * revSpecularWeight + fabsf(q.dir.z) / kPi * revDiffuseWeight) / (revSpecularWeight + revDiffuseWeight);
        break;
    }
    default:        // microfacet BRDF / BSDF (MicrofacetBSDF.cpp:35-38, 150-153, 192-195): the closed forms of their evaluatePDFInternal
        microfacetPdfRev(L, q, res.dir, &rev->pdf);
        break;
    }
}

// BSDF::evaluatePDF(query, dir, &revPDF) of one component (the base lobe, or the InverseBSDF around it, whose
// evaluatePDFInternal leaves the query's flags as they are -- basic_BSDFs.cpp:192-198)
template <int NC>
__device__ __forceinline__ float lobePdfInternalRev(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir, float* revPdf) {
    if (!L.inverse) return basePdfRev(L, q, dir, revPdf);
    if (!dtMatches(L.baseDirType, q.flags)) { *revPdf = 0.0f; return 0.0f; }
    return basePdfRev(L, q, V3(dir.x, dir.y, -dir.z), revPdf);
}

// sampleInternal with result->reverse of one component
template <int NC>
__device__ __forceinline__ Spec<NC> lobeSampleInternalRev(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res,
                                                          BsdfRev<NC>* rev) {
    if (!L.inverse) {
        const Spec<NC> fs = baseSample(L, q, uComp, u0, u1, res);
        if (res->pdf != 0.0f) baseSampleReverse(L, q, *res, fs, rev);
        return fs;
    }
    // InverseBSDF::sampleInternal: the base's PUBLIC sample with flipped flags (its own shading-normal correction included)
    BsdfQuery mq = q;
    mq.flags = dtFlip(q.flags);
    if (!dtMatches(L.baseDirType, mq.flags)) { res->pdf = 0.0f; res->type = 0; return specZero<NC>(); }
    const Spec<NC> fs = baseSample(L, mq, uComp, u0, u1, res);
    if (res->pdf != 0.0f) baseSampleReverse(L, mq, *res, fs, rev);
    const float snCorrection = snCorrectionOf(mq, res->dir);
    rev->fs = rev->fs * snCorrection;
    res->type = dtFlip(res->type);
    res->dir.z *= -1;
    return fs * snCorrection;
}

// BSDF::sample with result->reverse for the BSDF of a hit (MultiBSDF::sampleInternalWithRev for sum / mix materials)
template <int NC, int ML>
__device__ inline Spec<NC> bsdfSampleRev(const Bsdf<NC, ML>& b, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res, BsdfRev<NC>* rev) {
    res->pdf = 0.0f; res->type = 0; res->dir = V3(0, 0, 1);
    rev->fs = specZero<NC>(); rev->pdf = 0.0f;
    if (!dtMatches(b.type, q.flags)) return specZero<NC>();
    Spec<NC> value;
    if (!b.multi) {
        value = lobeSampleInternalRev(b.lobes[0], q, uComp, u0, u1, res, rev);
    } else {
        float weights[ML];
        float sum = 0.0f, comp = 0.0f;           // CompensatedSum as in sampleDiscrete (distributions.cpp:13-30)
#pragma unroll
        for (int i = 0; i < ML; ++i) {
            weights[i] = i < b.numLobes ? lobeWeight(b.lobes[i], q) : 0.0f;
            if (i < b.numLobes) { const float y = weights[i] - comp; const float t = sum + y; comp = (t - sum) - y; sum = t; }
        }
        const float sumWeights = sum;
        const float su = uComp * sumWeights;
        int idx = 0;
        float base = 0.0f;
        {
            float cum = 0.0f, ccomp = 0.0f;
            bool found = false;
#pragma unroll
            for (int i = 0; i < ML; ++i) {
                if (i < b.numLobes && !found) {
                    base = cum;
                    const float y = weights[i] - ccomp; const float t = cum + y; ccomp = (t - cum) - y; cum = t;
                    if (su < cum) { idx = i; found = true; }
                }
            }
            if (!found) idx = 0;
        }
        if (sumWeights == 0.0f) { res->pdf = 0.0f; return specZero<NC>(); }
        float wSel = 0.0f;
#pragma unroll
        for (int i = 0; i < ML; ++i) if (i == idx) wSel = weights[i];
        const float uc = (uComp * sumWeights - base) / wSel;
        value = specZero<NC>();
#pragma unroll
        for (int i = 0; i < ML; ++i) if (i == idx) value = lobeSampleInternalRev(b.lobes[i], q, uc, u0, u1, res, rev);
        if (res->pdf == 0.0f) return specZero<NC>();
        // the component weights seen from the sampled direction, for the opposite transport direction
        BsdfQuery rq = q;
        rq.dir = res->dir;
        rq.adjoint = !q.adjoint;
        float revWeights[ML];
        float sumRevWeights = 0.0f;
        float revSel = 0.0f;
#pragma unroll
        for (int i = 0; i < ML; ++i) {
            revWeights[i] = i < b.numLobes ? lobeWeight(b.lobes[i], rq) : 0.0f;
            sumRevWeights += revWeights[i];
            if (i == idx) revSel = revWeights[i];
        }
        res->pdf *= wSel;
        rev->pdf *= revSel;
        if (!dtIsDelta(res->type)) {
#pragma unroll
            for (int i = 0; i < ML; ++i)
                if (i < b.numLobes && i != idx && dtMatches(lobeDirType(b.lobes[i]), q.flags)) {
                    float revPdf;
                    res->pdf += lobePdfInternalRev(b.lobes[i], q, res->dir, &revPdf) * weights[i];
                    rev->pdf += revPdf * revWeights[i];
                }
            BsdfQuery mq = q;
            mq.flags &= sideTest(q.gn, q.dir, res->dir);
            value = specZero<NC>();
#pragma unroll
            for (int i = 0; i < ML; ++i)
                if (i < b.numLobes && dtMatches(lobeDirType(b.lobes[i]), mq.flags))
                    value = value + lobeEvaluateInternal(b.lobes[i], mq, res->dir);
            rev->fs = value;
        }
        res->pdf /= sumWeights;
        rev->pdf /= sumRevWeights;
    }
    const float snCorrection = snCorrectionOf(q, res->dir);
    rev->fs = rev->fs * snCorrection;
    return value * snCorrection;
}

// BSDF::evaluatePDF(query, dir, &revPDF) (MultiBSDF::evaluatePDFInternalWithRev for sum / mix materials)
template <int NC, int ML>
__device__ inline float bsdfPdfRev(const Bsdf<NC, ML>& b, const BsdfQuery& q, const V3& dir, float* revPdf) {
    *revPdf = 0.0f;
    if (!dtMatches(b.type, q.flags)) return 0.0f;
    if (!b.multi) return lobePdfInternalRev(b.lobes[0], q, dir, revPdf);
    BsdfQuery rq = q;
    rq.dir = dir;
    rq.adjoint = !q.adjoint;
    float weights[ML], revWeights[ML];
    float sumWeights = 0.0f, sumRevWeights = 0.0f;
#pragma unroll
    for (int i = 0; i < ML; ++i) {
        weights[i] = i < b.numLobes ? lobeWeight(b.lobes[i], q) : 0.0f;
        revWeights[i] = i < b.numLobes ? lobeWeight(b.lobes[i], rq) : 0.0f;
        sumWeights += weights[i]; sumRevWeights += revWeights[i];
    }
    if (sumWeights == 0.0f) return 0.0f;
    float ret = 0.0f, rev = 0.0f;
#pragma unroll
    for (int i = 0; i < ML; ++i)
        if (i < b.numLobes && weights[i] > 0) {
            float r;
            ret += lobePdfInternalRev(b.lobes[i], q, dir, &r) * weights[i];
            rev += r * revWeights[i];
        }
    *revPdf = rev / sumRevWeights;
    return ret / sumWeights;
}

}  // namespace slrgpu
