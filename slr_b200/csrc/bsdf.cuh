// BSDF lobes as device functions: sample / evaluate / evaluatePDF / weight in the local shading frame
// (z = shading normal), for radiance transport (adjoint = false, as PathTracingRenderer queries them).
//
//   public wrappers + shading-normal correction   libSLR/Core/directional_distribution_functions.h:231-289
//   DirectionType flags                            directional_distribution_functions.h:18-91
//   Fresnel conductor / dielectric, GGX (VNDF)     libSLR/Core/directional_distribution_functions.cpp:60-268
//   Lambert, SpecularBRDF, SpecularBSDF, Inverse   libSLR/BSDFs/basic_BSDFs.cpp:12-207
//   Oren-Nayar                                     libSLR/BSDFs/OrenNayerBRDF.cpp:12-73 (sin^2 used as sin: kept)
//   modified Ward-Duer                             libSLR/BSDFs/ModifiedWardDurBRDF.cpp:11-86
//   Ashikhmin-Shirley                              libSLR/BSDFs/AshikhminShirleyBRDF.cpp:12-169
//   microfacet BRDF / BSDF                         libSLR/BSDFs/MicrofacetBSDF.cpp:11-315
//   MultiBSDF (sum / mix of up to 4 lobes)         libSLR/BSDFs/MultiBSDF.cpp:20-218
//
// The reference allocates one polymorphic object per hit in an arena; here a hit's BSDF is a small
// array of tagged lobes living in registers / local memory, and every virtual call is a switch.
#pragma once
#include "spectral.cuh"
#include "vecmath.cuh"

namespace slrgpu {

constexpr uint32_t DT_LowFreq = 1u << 0, DT_HighFreq = 1u << 1, DT_Delta0D = 1u << 2, DT_Delta1D = 1u << 3;
constexpr uint32_t DT_NonDelta = DT_LowFreq | DT_HighFreq, DT_Delta = DT_Delta0D | DT_Delta1D, DT_AllFreq = DT_NonDelta | DT_Delta;
constexpr uint32_t DT_Reflection = 1u << 4, DT_Transmission = 1u << 5, DT_WholeSphere = DT_Reflection | DT_Transmission;
constexpr uint32_t DT_All = DT_AllFreq | DT_WholeSphere, DT_Dispersive = 1u << 6;

__device__ __forceinline__ bool dtMatches(uint32_t type, uint32_t flags) {
    const uint32_t r = type & flags;
    return (r & DT_WholeSphere) && (r & DT_AllFreq);
}
__device__ __forceinline__ bool dtIsDelta(uint32_t t) { return (t & DT_Delta) && !(t & DT_NonDelta); }
__device__ __forceinline__ bool dtIsReflection(uint32_t t) { return (t & DT_Reflection) && !(t & DT_Transmission); }
__device__ __forceinline__ bool dtIsTransmission(uint32_t t) { return !(t & DT_Reflection) && (t & DT_Transmission); }
__device__ __forceinline__ uint32_t dtFlip(uint32_t t) { return t ^ DT_WholeSphere; }

enum LobeType : uint32_t {
    LOBE_LAMBERT = 0,        // s0 = R
    LOBE_OREN_NAYAR = 1,     // s0 = R, f0 = A, f1 = B
    LOBE_SPECULAR_BRDF = 2,  // s0 = coeffR, s1 = eta, s2 = k
    LOBE_SPECULAR_BSDF = 3,  // s0 = coeff, s1 = etaExt, s2 = etaInt
    LOBE_WARD = 4,           // s0 = R, f0 = anisoX, f1 = anisoY
    LOBE_ASHIKHMIN = 5,      // s0 = Rs, s1 = Rd, f0 = nu, f1 = nv
    LOBE_MF_BRDF = 6,        // s0 = eta, s1 = k, f0 = alpha_g
    LOBE_MF_BSDF = 7         // s0 = etaExt, s1 = etaInt, f0 = alpha_g
};

template <int NC> struct Lobe {
    uint32_t type;
    uint32_t baseDirType;    // m_type of the base BSDF
    uint32_t inverse;        // wrapped in an InverseBSDF
    float f0, f1;
    Spec<NC> s0, s1, s2;
};

struct BsdfQuery {
    V3 dir;          // dir_sn
    V3 gn;           // gNormal_sn
    uint32_t hero;   // wlHint
    uint32_t flags;
    // BSDFQuery::adjoint: false for radiance transport (every query of the path tracer), true for the queries of a light
    // subpath (bidirectional path tracing, bpt.cu). It moves the shading-normal correction to the query direction and drops
    // the (eta_enter / eta_exit)^2 radiance scaling of refraction.
    bool adjoint = false;
};
// BSDF::sample / evaluate (directional_distribution_functions.h:231-275): |cos| ratio between shading and geometric normal,
// taken at the sampled / evaluated direction for radiance and at the query direction for importance
__device__ __forceinline__ float snCorrectionOf(const BsdfQuery& q, const V3& dir) {
    return q.adjoint ? fabsf(q.dir.z / dot(q.dir, q.gn)) : fabsf(dir.z / dot(dir, q.gn));
}
// BSDF::weight (directional_distribution_functions.h:276-285)
__device__ __forceinline__ float snWeightCorrectionOf(const BsdfQuery& q) { return q.adjoint ? fabsf(q.dir.z / dot(q.dir, q.gn)) : 1.0f; }

struct BsdfSampleResult {
    V3 dir;
    float pdf;
    uint32_t type;
};

__device__ __forceinline__ uint32_t sideTest(const V3& ng, const V3& d0, const V3& d1) {
    const bool reflect = dot(ng, d0) * dot(ng, d1) > 0;
    return DT_AllFreq | (reflect ? DT_Reflection : DT_Transmission);
}

// ---------------------------------------------------------------------------------------------
// Fresnel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fresnelConductor1(float eta, float k, float cosEnter) {
    cosEnter = fabsf(cosEnter);
    const float c2 = cosEnter * cosEnter;
    const float twoEtaCos = 2.0f * eta * cosEnter;
    const float tmp_f = eta * eta + k * k;
    const float tmp = tmp_f * c2;
    const float Rparl2 = (tmp - twoEtaCos + 1) / (tmp + twoEtaCos + 1);
    const float Rperp2 = (tmp_f - twoEtaCos + c2) / (tmp_f + twoEtaCos + c2);
    return (Rparl2 + Rperp2) / 2.0f;
}
template <int NC> __device__ __forceinline__ Spec<NC> fresnelConductor(const Spec<NC>& eta, const Spec<NC>& k, float cosEnter) {
    Spec<NC> r;
#pragma unroll
    for (int i = 0; i < NC; ++i) r.v[i] = fresnelConductor1(eta.v[i], k.v[i], cosEnter);
    return r;
}
__device__ __forceinline__ float dielectricEvalF(float etaEnter, float etaExit, float cosEnter, float cosExit) {
    const float Rparl = ((etaExit * cosEnter) - (etaEnter * cosExit)) / ((etaExit * cosEnter) + (etaEnter * cosExit));
    const float Rperp = ((etaEnter * cosEnter) - (etaExit * cosExit)) / ((etaEnter * cosEnter) + (etaExit * cosExit));
    return (Rparl * Rparl + Rperp * Rperp) / 2.0f;
}
__device__ __forceinline__ float fresnelDielectric1(float etaExt, float etaInt, float cosEnter) {
    cosEnter = fminf(fmaxf(cosEnter, -1.0f), 1.0f);
    const bool entering = cosEnter > 0.0f;
    const float eEnter = entering ? etaExt : etaInt;
    const float eExit = entering ? etaInt : etaExt;
    const float sinExit = eEnter / eExit * sqrtf(fmaxf(0.0f, 1.0f - cosEnter * cosEnter));
    cosEnter = fabsf(cosEnter);
    if (sinExit >= 1.0f) return 1.0f;
    const float cosExit = sqrtf(fmaxf(0.0f, 1.0f - sinExit * sinExit));
    return dielectricEvalF(eEnter, eExit, cosEnter, cosExit);
}
template <int NC> __device__ __forceinline__ Spec<NC> fresnelDielectric(const Spec<NC>& etaExt, const Spec<NC>& etaInt, float cosEnter) {
    Spec<NC> r;
#pragma unroll
    for (int i = 0; i < NC; ++i) r.v[i] = fresnelDielectric1(etaExt.v[i], etaInt.v[i], cosEnter);
    return r;
}

// ---------------------------------------------------------------------------------------------
// GGX
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2(float x) { return x * x; }
__device__ __forceinline__ float pow4(float x) { const float y = x * x; return y * y; }
__device__ __forceinline__ float pow5(float x) { const float y = x * x; return y * y * x; }

// The reference writes tan(acos(z)) (MicrofacetBSDF.cpp GGX::evaluate / evaluateSmithG1); tan^2 = (1 - z^2) / z^2 is the
// same quantity without the two libm calls (ncu: they were 20 % of the rough-glass kernel's stall samples). The two forms
// differ by rounding only where z -> 1, and there tan^2 is negligible against alpha^2 / against 1.
__device__ __forceinline__ float tan2FromCos(float z) { return fmaxf(0.0f, 1.0f - z * z) / (z * z); }
__device__ __forceinline__ float ggxD(float alpha, const V3& m) {
    if (m.z <= 0) return 0.0f;
    return alpha * alpha / (kPi * pow4(m.z) * pow2(alpha * alpha + tan2FromCos(m.z)));
}
__device__ __forceinline__ float ggxSmithG1(float alpha, const V3& v, const V3& m) {
    const float chi = (dot(v, m) / v.z) > 0 ? 1.0f : 0.0f;
    const float z = fminf(fmaxf(v.z, -1.0f), 1.0f);
    return chi * 2 / (1 + sqrtf(1 + alpha * alpha * tan2FromCos(z)));
}
__device__ __forceinline__ float ggxPdfVisible(float alpha, const V3& v, const V3& m) {
    return ggxSmithG1(alpha, v, m) * absDot(v, m) * ggxD(alpha, m) / fabsf(v.z);
}
// Heitz's visible-normal sampling as the reference implements it (doubles where it uses double literals)
static __device__ __noinline__ float ggxSampleVisible(float alpha, const V3& v, float u0, float u1, V3* m, float* normalPDF) {
    V3 sv = normalize(V3(alpha * v.x, alpha * v.y, v.z));
    // theta_sv = acos(sv.z), phi_sv = atan2(sv.y, sv.x) in the reference; only tan(theta), cos(phi), sin(phi) and the test
    // theta < 1e-4 are used, all of which come from sv's components without the libm calls
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - sv.z * sv.z));
    const float lenXY = sqrtf(sv.x * sv.x + sv.y * sv.y);
    float cp = lenXY > 0.0f ? sv.x / lenXY : 1.0f, sp = lenXY > 0.0f ? sv.y / lenXY : 0.0f;
    if (sv.z > 0.99999f) { sinTheta = 0.0f; cp = 1.0f; sp = 0.0f; }
    float slope_x, slope_y;
    if (sinTheta < 0.0001f && sv.z > 0.0f) {
        const float r = sqrtf(u0 / (1 - u0));
        const float phi = 2 * kPi * u1;
        float s, c;
        sincosf(phi, &s, &c);
        slope_x = r * c; slope_y = r * s;
    } else {
        const float tan_theta_i = sinTheta / sv.z;
        const float a = 1 / tan_theta_i;
        const float G1 = 2 / (1 + sqrtf(1.0f + 1.0f / (a * a)));
        const float A = 2.0f * u0 / G1 - 1.0f;
        const float tmp = 1.0f / (A * A - 1.0f);
        const float B = tan_theta_i;
        const float D = sqrtf(B * B * tmp * tmp - (A * A - B * B) * tmp);
        const float slope_x_1 = B * tmp - D;
        const float slope_x_2 = B * tmp + D;
        slope_x = (A < 0 || slope_x_2 > 1.0f / tan_theta_i) ? slope_x_1 : slope_x_2;
        if (u0 == 0) slope_x = 0;
        float S;
        if (u1 > 0.5f) { S = 1.0f; u1 = 2.0f * (u1 - 0.5f); }
        else { S = -1.0f; u1 = 2.0f * (0.5f - u1); }
        const float z = (u1 * (u1 * (u1 * 0.27385f - 0.73369f) + 0.46341f)) / (u1 * (u1 * (u1 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
        slope_y = S * z * sqrtf(1.0f + slope_x * slope_x);
    }
    const float tmp = cp * slope_x - sp * slope_y;
    slope_y = sp * slope_x + cp * slope_y;
    slope_x = tmp;
    slope_x *= alpha; slope_y *= alpha;
    *m = normalize(V3(-slope_x, -slope_y, 1));
    const float D = ggxD(alpha, *m);
    *normalPDF = ggxSmithG1(alpha, v, *m) * absDot(v, *m) * D / fabsf(v.z);
    return D;
}

// ---------------------------------------------------------------------------------------------
// base lobes: sampleInternal / evaluateInternal / evaluatePDFInternal / weightInternal
// ---------------------------------------------------------------------------------------------
template <int NC> __device__ __forceinline__ Spec<NC> specZero() { return specConst<NC>(0.0f); }

template <int NC>
static __device__ __noinline__ Spec<NC> ashikhminEval(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir, const V3& halfv, float* specPDF) {
    const float nu = L.f0, nv = L.f1;
    const float dotHV = dot(halfv, q.dir);
    const float ex = (nu * halfv.x * halfv.x + nv * halfv.y * halfv.y) / (1 - halfv.z * halfv.z);
    const float commonTerm = sqrtf((nu + 1) * (nv + 1)) / (8 * kPi * dotHV) * powf(fabsf(halfv.z), ex);
    *specPDF = commonTerm;
    const float schlick = pow5(1.0f - dotHV);
    const float sScale = commonTerm / fmaxf(fabsf(q.dir.z), fabsf(dir.z));
    const float dScale = (1.0f - pow5(1.0f - fabsf(q.dir.z) / 2)) * (1.0f - pow5(1.0f - fabsf(dir.z) / 2));
    Spec<NC> fs;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const float Rs = L.s0.v[i], Rd = L.s1.v[i];
        const float F = Rs + (1.0f - Rs) * schlick;
        fs.v[i] = sScale * F + 28 * Rd / (23 * kPi) * (1.0f - Rs) * dScale;
    }
    return fs;
}
template <int NC>
__device__ __forceinline__ void ashikhminWeights(const Lobe<NC>& L, const BsdfQuery& q, float* specularWeight, float* diffuseWeight) {
    const float iRs = specImportance(L.s0, q.hero);
    const float iRd = specImportance(L.s1, q.hero);
    const float vDotHV = fabsf(q.dir.z);
    *specularWeight = iRs + (1 - iRs) * pow5(1.0f - vDotHV);
    const float transmissionTerm = 1 - pow5(1 - vDotHV * 0.5f);
    *diffuseWeight = 28 * iRd / 23 * (1 - iRs) * transmissionTerm * transmissionTerm;
}

// rough-refraction value for all wavelengths with per-wavelength half vectors (MicrofacetBSDF.cpp:174-188)
template <int NC>
static __device__ __noinline__ Spec<NC> mfTransmissionEval(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir, bool entering) {
    const float alpha = L.f0;
    Spec<NC> ret;
#pragma unroll 1
    for (int i = 0; i < NC; ++i) {
        const float eEnter = entering ? L.s0.v[i] : L.s1.v[i];
        const float eExit = entering ? L.s1.v[i] : L.s0.v[i];
        const V3 m = normalize(-(eEnter * q.dir + eExit * dir));
        const float dotHV = dot(q.dir, m), dotHL = dot(dir, m);
        const float F = fresnelDielectric1(L.s0.v[i], L.s1.v[i], dotHV);
        const float G = ggxSmithG1(alpha, q.dir, m) * ggxSmithG1(alpha, dir, m);
        const float D = ggxD(alpha, m);
        ret.v[i] = fabsf(dotHV * dotHL) * (1 - F) * G * D / pow2(eEnter * dotHV + eExit * dotHL);
    }
    const float inv = 1.0f / fabsf(q.dir.z * dir.z);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        // adjoint: the eEnter^2 / eExit^2 of radiance transport is cancelled (MicrofacetBSDF.cpp:187,239)
        const float eScale = (entering != q.adjoint) ? L.s0.v[i] : L.s1.v[i];
        ret.v[i] = ret.v[i] * inv * (eScale * eScale);
    }
    return ret;
}

// The four base functions are templates over LT: LT >= 0 fixes the lobe type at compile time (the
// per-material-class shade kernels; everything inlines and the other cases are pruned), LT = -1
// dispatches on L.type at run time (the generic kernel; kept out of line).
template <int NC, int LT>
__device__ __forceinline__ Spec<NC> baseSampleT(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    switch (LT >= 0 ? (uint32_t)LT : L.type) {
    case LOBE_LAMBERT: {
        res->dir = cosineSampleHemisphere(u0, u1);
        res->pdf = res->dir.z / kPi;
        res->type = L.baseDirType;
        res->dir.z *= dot(q.dir, q.gn) > 0 ? 1 : -1;
        return L.s0 * (1.0f / kPi);
    }
    case LOBE_OREN_NAYAR: {
        const bool frontSide = dot(q.dir, q.gn) > 0;
        res->dir = cosineSampleHemisphere(u0, u1);
        res->pdf = res->dir.z / kPi;
        res->type = L.baseDirType;
        res->dir.z *= frontSide ? 1 : -1;
        const float sinThetaI = 1.0f - res->dir.z * res->dir.z;
        const float sinThetaO = 1.0f - q.dir.z * q.dir.z;
        const float absTanThetaI = sinThetaI / fabsf(res->dir.z);
        const float absTanThetaO = sinThetaO / fabsf(q.dir.z);
        const float sinAlpha = fmaxf(sinThetaI, sinThetaO);
        const float tanBeta = fminf(absTanThetaI, absTanThetaO);
        float cos_dAzimuth = (res->dir.x * q.dir.x + res->dir.y * q.dir.y) / (sinThetaI * sinThetaO);
        if (!isfinite(cos_dAzimuth)) cos_dAzimuth = 0.0f;
        return L.s0 * ((L.f0 + L.f1 * fmaxf(0.0f, cos_dAzimuth) * sinAlpha * tanBeta) / kPi);
    }
    case LOBE_SPECULAR_BRDF: {
        res->dir = V3(-q.dir.x, -q.dir.y, q.dir.z);
        res->pdf = 1.0f;
        res->type = L.baseDirType;
        return L.s0 * fresnelConductor(L.s1, L.s2, q.dir.z) * (1.0f / fabsf(q.dir.z));
    }
    case LOBE_SPECULAR_BSDF: {
        const Spec<NC> F = fresnelDielectric(L.s1, L.s2, q.dir.z);
        float reflectProb = specImportance(F, q.hero);
        if (dtIsReflection(q.flags)) reflectProb = 1.0f;
        if (dtIsTransmission(q.flags)) reflectProb = 0.0f;
        if (uComp < reflectProb) {
            if (q.dir.z == 0.0f) { res->pdf = 0.0f; return specZero<NC>(); }
            res->dir = V3(-q.dir.x, -q.dir.y, q.dir.z);
            res->pdf = reflectProb;
            res->type = DT_Reflection | DT_Delta0D;
            return L.s0 * F * (1.0f / fabsf(q.dir.z));
        }
        const bool entering = q.dir.z > 0.0f;
        const float etaExtH = specAt(L.s1, q.hero), etaIntH = specAt(L.s2, q.hero);
        const float eEnter = entering ? etaExtH : etaIntH;
        const float eExit = entering ? etaIntH : etaExtH;
        const float sinEnter2 = 1.0f - q.dir.z * q.dir.z;
        const float rrEta = eEnter / eExit;
        const float sinExit2 = rrEta * rrEta * sinEnter2;
        if (sinExit2 >= 1.0f) { res->pdf = 0.0f; return specZero<NC>(); }
        float cosExit = sqrtf(fmaxf(0.0f, 1.0f - sinExit2));
        if (entering) cosExit = -cosExit;
        res->dir = V3(rrEta * -q.dir.x, rrEta * -q.dir.y, cosExit);
        res->pdf = 1.0f - reflectProb;
        res->type = DT_Transmission | DT_Delta0D | (L.baseDirType & DT_Dispersive);
        float v = specAt(L.s0, q.hero) * (1.0f - specAt(F, q.hero));
        if (!q.adjoint) v *= (eEnter * eEnter) / (eExit * eExit);
        v /= fabsf(cosExit);
        Spec<NC> ret;
#pragma unroll
        for (int i = 0; i < NC; ++i) ret.v[i] = (i == (int)q.hero) ? v : 0.0f;
        return ret;
    }
    case LOBE_WARD: {
        const float ax = L.f0, ay = L.f1;
        const float quad = 2 * kPi * u1;
        float sq, cq;
        sincosf(quad, &sq, &cq);
        const float phi_h = atan2f(ay * sq, ax * cq);
        float sph, cph;
        sincosf(phi_h, &sph, &cph);
        const float cosphi_ax = cph / ax, sinphi_ay = sph / ay;
        const float theta_h = atanf(sqrtf(-logf(1 - u0) / (cosphi_ax * cosphi_ax + sinphi_ay * sinphi_ay)));
        float sth, cth;
        sincosf(theta_h, &sth, &cth);
        V3 halfv(sth * cph, sth * sph, cth);
        halfv.z *= q.dir.z > 0 ? 1 : -1;
        res->dir = 2 * dot(q.dir, halfv) * halfv - q.dir;
        if (res->dir.z * q.dir.z <= 0) { res->pdf = 0.0f; return specZero<NC>(); }
        const float hx_ax = halfv.x / ax, hy_ay = halfv.y / ay;
        const float dotHN = fabsf(halfv.z);
        const float dotHI = dot(halfv, res->dir);
        const float numerator = expf(-(hx_ax * hx_ax + hy_ay * hy_ay) / (dotHN * dotHN));
        const float commonDenom = 4 * kPi * ax * ay * dotHI * dotHN * dotHN * dotHN;
        res->pdf = numerator / commonDenom;
        res->type = L.baseDirType;
        return L.s0 * (numerator / (commonDenom * dotHI * dotHN));
    }
    case LOBE_ASHIKHMIN: {
        const float nu = L.f0, nv = L.f1;
        float specularWeight, diffuseWeight;
        ashikhminWeights(L, q, &specularWeight, &diffuseWeight);
        const float sumWeights = specularWeight + diffuseWeight;
        float specularDirPDF, diffuseDirPDF;
        Spec<NC> fs;
        if (uComp * sumWeights < specularWeight) {
            res->type = DT_Reflection | DT_HighFreq;
            const float quad = 2 * kPi * u1;
            float sq, cq;
            sincosf(quad, &sq, &cq);
            const float phi_h = atan2f(sqrtf(nu + 1) * sq, sqrtf(nv + 1) * cq);
            float sinphi, cosphi;
            sincosf(phi_h, &sinphi, &cosphi);
            float theta_h = acosf(powf(1 - u0, 1.0f / (nu * cosphi * cosphi + nv * sinphi * sinphi + 1)));
            if (q.dir.z < 0) theta_h = kPi - theta_h;
            float sth, cth;
            sincosf(theta_h, &sth, &cth);
            const V3 halfv(sth * cosphi, sth * sinphi, cth);
            res->dir = 2 * dot(q.dir, halfv) * halfv - q.dir;
            if (res->dir.z * q.dir.z <= 0) { res->pdf = 0.0f; return specZero<NC>(); }
            fs = ashikhminEval(L, q, res->dir, halfv, &specularDirPDF);
            diffuseDirPDF = fabsf(res->dir.z) / kPi;
        } else {
            res->type = DT_Reflection | DT_LowFreq;
            res->dir = cosineSampleHemisphere(u0, u1);
            diffuseDirPDF = res->dir.z / kPi;
            res->dir.z *= dot(q.dir, q.gn) > 0 ? 1 : -1;
            const V3 halfv = halfVector(q.dir, res->dir);
            fs = ashikhminEval(L, q, res->dir, halfv, &specularDirPDF);
        }
        res->pdf = (specularDirPDF * specularWeight + diffuseDirPDF * diffuseWeight) / sumWeights;
        return fs;
    }
    case LOBE_MF_BRDF: {
        const float alpha = L.f0;
        const bool entering = q.dir.z >= 0.0f;
        const float sign = entering ? 1.0f : -1.0f;
        V3 m;
        float mPDF;
        const float D = ggxSampleVisible(alpha, sign * q.dir, u0, u1, &m, &mPDF);
        const float dotHV = dot(q.dir, m);
        if (dotHV * sign <= 0) { res->pdf = 0.0f; return specZero<NC>(); }
        res->dir = 2 * dotHV * m - q.dir;
        if (res->dir.z * q.dir.z <= 0) { res->pdf = 0.0f; return specZero<NC>(); }
        const float commonPDFTerm = 1.0f / (4 * dotHV * sign);
        res->pdf = commonPDFTerm * mPDF;
        res->type = L.baseDirType;
        const Spec<NC> F = fresnelConductor(L.s0, L.s1, dotHV);
        const float G = ggxSmithG1(alpha, q.dir, m) * ggxSmithG1(alpha, res->dir, m);
        return F * (D * G / (4 * q.dir.z * res->dir.z));
    }
    case LOBE_MF_BSDF: {
        const float alpha = L.f0;
        const bool entering = q.dir.z >= 0.0f;
        const float sign = entering ? 1.0f : -1.0f;
        V3 m;
        float mPDF;
        const float D = ggxSampleVisible(alpha, sign * q.dir, u0, u1, &m, &mPDF);
        const float dotHV = dot(q.dir, m);
        if (dotHV * sign <= 0 || isnan(D)) { res->pdf = 0.0f; return specZero<NC>(); }
        const Spec<NC> F = fresnelDielectric(L.s0, L.s1, dotHV);
        float reflectProb = specImportance(F, q.hero);
        if (dtIsReflection(q.flags)) reflectProb = 1.0f;
        if (dtIsTransmission(q.flags)) reflectProb = 0.0f;
        if (uComp < reflectProb) {
            res->dir = 2 * dotHV * m - q.dir;
            if (res->dir.z * q.dir.z <= 0) { res->pdf = 0.0f; return specZero<NC>(); }
            const float commonPDFTerm = reflectProb / (4 * dotHV * sign);
            res->pdf = commonPDFTerm * mPDF;
            res->type = DT_Reflection | DT_HighFreq;
            const float G = ggxSmithG1(alpha, q.dir, m) * ggxSmithG1(alpha, res->dir, m);
            return F * (D * G / (4 * q.dir.z * res->dir.z));
        }
        const float etaExtH = specAt(L.s0, q.hero), etaIntH = specAt(L.s1, q.hero);
        const float eEnterH = entering ? etaExtH : etaIntH;
        const float eExitH = entering ? etaIntH : etaExtH;
        const float recRelIOR = eEnterH / eExitH;
        const float innerRoot = 1 + recRelIOR * recRelIOR * (dotHV * dotHV - 1);
        if (innerRoot < 0) { res->pdf = 0.0f; return specZero<NC>(); }
        res->dir = (recRelIOR * dotHV - sign * sqrtf(innerRoot)) * m - recRelIOR * q.dir;
        if (res->dir.z * q.dir.z >= 0) { res->pdf = 0.0f; return specZero<NC>(); }
        const float dotHL = dot(res->dir, m);
        const float commonPDFTerm = (1 - reflectProb) / pow2(eEnterH * dotHV + eExitH * dotHL);
        res->pdf = commonPDFTerm * mPDF * eExitH * eExitH * fabsf(dotHL);
        res->type = DT_Transmission | DT_HighFreq;
        return mfTransmissionEval(L, q, res->dir, entering);
    }
    }
    res->pdf = 0.0f;
    return specZero<NC>();
}

template <int NC>
static __device__ __noinline__ Spec<NC> baseSample(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    return baseSampleT<NC, -1>(L, q, uComp, u0, u1, res);
}

template <int NC, int LT>
__device__ __forceinline__ Spec<NC> baseEvaluateT(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    switch (LT >= 0 ? (uint32_t)LT : L.type) {
    case LOBE_LAMBERT:
        if (q.dir.z * dir.z <= 0.0f) return specZero<NC>();
        return L.s0 * (1.0f / kPi);
    case LOBE_OREN_NAYAR: {
        if (q.dir.z * dir.z <= 0.0f) return specZero<NC>();
        const float sinThetaI = 1.0f - dir.z * dir.z;
        const float sinThetaO = 1.0f - q.dir.z * q.dir.z;
        const float absTanThetaI = sinThetaI / fabsf(dir.z);
        const float absTanThetaO = sinThetaO / fabsf(q.dir.z);
        const float sinAlpha = fmaxf(sinThetaI, sinThetaO);
        const float tanBeta = fminf(absTanThetaI, absTanThetaO);
        const float cos_dAzimuth = (dir.x * q.dir.x + dir.y * q.dir.y) / (sinThetaI * sinThetaO);
        return L.s0 * ((L.f0 + L.f1 * fmaxf(0.0f, cos_dAzimuth) * sinAlpha * tanBeta) / kPi);
    }
    case LOBE_SPECULAR_BRDF:
    case LOBE_SPECULAR_BSDF:
        return specZero<NC>();
    case LOBE_WARD: {
        if (dir.z * q.dir.z <= 0) return specZero<NC>();
        const float ax = L.f0, ay = L.f1;
        const V3 halfv = normalize(q.dir + dir);
        const float hx_ax = halfv.x / ax, hy_ay = halfv.y / ay;
        const float dotHN = fabsf(halfv.z);
        const float dotHI = dot(halfv, dir);
        const float numerator = expf(-(hx_ax * hx_ax + hy_ay * hy_ay) / (dotHN * dotHN));
        const float denominator = 4 * kPi * ax * ay * dotHI * dotHI * dotHN * dotHN * dotHN * dotHN;
        return L.s0 * (numerator / denominator);
    }
    case LOBE_ASHIKHMIN: {
        if (dir.z * q.dir.z <= 0) return specZero<NC>();
        float specPDF;
        return ashikhminEval(L, q, dir, halfVector(q.dir, dir), &specPDF);
    }
    case LOBE_MF_BRDF: {
        if (dir.z * q.dir.z <= 0) return specZero<NC>();
        const float alpha = L.f0;
        const float sign = q.dir.z >= 0.0f ? 1.0f : -1.0f;
        const V3 m = sign * halfVector(q.dir, dir);
        const float dotHV = dot(q.dir, m);
        const float D = ggxD(alpha, m);
        const Spec<NC> F = fresnelConductor(L.s0, L.s1, dotHV);
        const float G = ggxSmithG1(alpha, q.dir, m) * ggxSmithG1(alpha, dir, m);
        return F * (D * G / (4 * q.dir.z * dir.z));
    }
    case LOBE_MF_BSDF: {
        const float alpha = L.f0;
        const bool entering = q.dir.z >= 0.0f;
        const float sign = entering ? 1.0f : -1.0f;
        const float dotNVdotNL = dir.z * q.dir.z;
        if (dotNVdotNL > 0 && dtMatches(q.flags, DT_Reflection | DT_AllFreq)) {
            const V3 m = sign * halfVector(q.dir, dir);
            const float dotHV = dot(q.dir, m);
            const float D = ggxD(alpha, m);
            const Spec<NC> F = fresnelDielectric(L.s0, L.s1, dotHV);
            const float G = ggxSmithG1(alpha, q.dir, m) * ggxSmithG1(alpha, dir, m);
            return F * (D * G / (4 * dotNVdotNL));
        } else if (dotNVdotNL < 0 && dtMatches(q.flags, DT_Transmission | DT_AllFreq)) {
            return mfTransmissionEval(L, q, dir, entering);
        }
        return specZero<NC>();
    }
    }
    return specZero<NC>();
}

template <int NC>
static __device__ __noinline__ Spec<NC> baseEvaluate(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) { return baseEvaluateT<NC, -1>(L, q, dir); }

template <int NC, int LT>
__device__ __forceinline__ float basePdfT(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    switch (LT >= 0 ? (uint32_t)LT : L.type) {
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR:
        if (q.dir.z * dir.z <= 0.0f) return 0.0f;
        return fabsf(dir.z) / kPi;
    case LOBE_SPECULAR_BRDF:
    case LOBE_SPECULAR_BSDF:
        return 0.0f;
    case LOBE_WARD: {
        if (dir.z * q.dir.z <= 0) return 0.0f;
        const float ax = L.f0, ay = L.f1;
        const V3 halfv = normalize(q.dir + dir);
        const float hx_ax = halfv.x / ax, hy_ay = halfv.y / ay;
        const float dotHN = fabsf(halfv.z);
        const float dotHI = dot(halfv, dir);
        const float numerator = expf(-(hx_ax * hx_ax + hy_ay * hy_ay) / (dotHN * dotHN));
        const float denominator = 4 * kPi * ax * ay * dotHI * dotHN * dotHN * dotHN;
        return numerator / denominator;
    }
    case LOBE_ASHIKHMIN: {
        if (dir.z * q.dir.z <= 0) return 0.0f;
        const float nu = L.f0, nv = L.f1;
        const V3 halfv = halfVector(q.dir, dir);
        const float dotHV = dot(halfv, q.dir);
        const float ex = (nu * halfv.x * halfv.x + nv * halfv.y * halfv.y) / (1 - halfv.z * halfv.z);
        const float specularDirPDF = sqrtf((nu + 1) * (nv + 1)) / (8 * kPi * dotHV) * powf(fabsf(halfv.z), ex);
        const float diffuseDirPDF = fabsf(dir.z) / kPi;
        float specularWeight, diffuseWeight;
        ashikhminWeights(L, q, &specularWeight, &diffuseWeight);
        return (specularDirPDF * specularWeight + diffuseDirPDF * diffuseWeight) / (specularWeight + diffuseWeight);
    }
    case LOBE_MF_BRDF: {
        if (dir.z * q.dir.z <= 0) return 0.0f;
        const float alpha = L.f0;
        const float sign = q.dir.z >= 0.0f ? 1.0f : -1.0f;
        const V3 m = sign * halfVector(q.dir, dir);
        const float dotHV = dot(q.dir, m);
        if (dotHV * sign <= 0) return 0.0f;
        const float mPDF = ggxPdfVisible(alpha, sign * q.dir, m);
        return 1.0f / (4 * dotHV * sign) * mPDF;
    }
    case LOBE_MF_BSDF: {
        const float alpha = L.f0;
        const bool entering = q.dir.z >= 0.0f;
        const float sign = entering ? 1.0f : -1.0f;
        const float dotNVdotNL = dir.z * q.dir.z;
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
        if (dotNVdotNL > 0) return reflectProb / (4 * dotHV * sign) * mPDF;
        const float dotHL = dot(dir, m);
        const float commonPDFTerm = (1 - reflectProb) / pow2(eEnter * dotHV + eExit * dotHL);
        return commonPDFTerm * mPDF * eExit * eExit * fabsf(dotHL);
    }
    }
    return 0.0f;
}

template <int NC>
static __device__ __noinline__ float basePdf(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) { return basePdfT<NC, -1>(L, q, dir); }

template <int NC>
static __device__ __noinline__ float baseWeight(const Lobe<NC>& L, const BsdfQuery& q) {
    switch (L.type) {
    case LOBE_LAMBERT: return specImportance(L.s0, q.hero);
    case LOBE_OREN_NAYAR: {   // luminance() = plain mean in spectral mode (SpectrumTypes.h:504-509)
        return specLuminance(L.s0);
    }
    case LOBE_SPECULAR_BRDF: return specImportance(L.s0, q.hero) * specImportance(fresnelConductor(L.s1, L.s2, q.dir.z), q.hero);
    case LOBE_SPECULAR_BSDF: return specImportance(L.s0, q.hero);
    case LOBE_WARD: return specImportance(L.s0, q.hero);
    case LOBE_ASHIKHMIN: {
        float sw, dw;
        ashikhminWeights(L, q, &sw, &dw);
        return sw + dw;
    }
    case LOBE_MF_BRDF:
    case LOBE_MF_BSDF: {
        const float sign = q.dir.z >= 0.0f ? 1.0f : -1.0f;
        return ggxSmithG1(L.f0, q.dir * sign, V3(0, 0, 1));
    }
    }
    return 0.0f;
}

// ---------------------------------------------------------------------------------------------
// public wrappers on a base lobe (BSDF::sample / evaluate / evaluatePDF / weight, adjoint = false)
// ---------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ Spec<NC> basePublicSample(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    if (!dtMatches(L.baseDirType, q.flags)) { res->pdf = 0.0f; res->type = 0; return specZero<NC>(); }
    const Spec<NC> fs = baseSample(L, q, uComp, u0, u1, res);
    const float snCorrection = snCorrectionOf(q, res->dir);
    return fs * snCorrection;
}
template <int NC>
__device__ __forceinline__ Spec<NC> basePublicEvaluate(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    BsdfQuery mq = q;
    mq.flags &= sideTest(q.gn, q.dir, dir);
    if (!dtMatches(L.baseDirType, mq.flags)) return specZero<NC>();
    const Spec<NC> fs = baseEvaluate(L, mq, dir);
    const float snCorrection = snCorrectionOf(q, dir);
    return fs * snCorrection;
}
template <int NC>
__device__ __forceinline__ float basePublicPdf(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    if (!dtMatches(L.baseDirType, q.flags)) return 0.0f;
    return basePdf(L, q, dir);
}
template <int NC>
__device__ __forceinline__ float basePublicWeight(const Lobe<NC>& L, const BsdfQuery& q) {
    if (!dtMatches(L.baseDirType, q.flags)) return 0.0f;
    return baseWeight(L, q) * snWeightCorrectionOf(q);
}

// ---------------------------------------------------------------------------------------------
// one lobe as a component: the base BSDF, or an InverseBSDF around it (basic_BSDFs.cpp:173-207;
// its evaluatePDFInternal discards the flipped flags -- kept)
// ---------------------------------------------------------------------------------------------
template <int NC> __device__ __forceinline__ uint32_t lobeDirType(const Lobe<NC>& L) { return L.inverse ? dtFlip(L.baseDirType) : L.baseDirType; }

template <int NC>
__device__ __forceinline__ Spec<NC> lobeSampleInternal(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    if (!L.inverse) return baseSample(L, q, uComp, u0, u1, res);
    BsdfQuery mq = q;
    mq.flags = dtFlip(q.flags);
    const Spec<NC> ret = basePublicSample(L, mq, uComp, u0, u1, res);
    res->type = dtFlip(res->type);
    res->dir.z *= -1;
    return ret;
}
template <int NC>
__device__ __forceinline__ Spec<NC> lobeEvaluateInternal(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    if (!L.inverse) return baseEvaluate(L, q, dir);
    BsdfQuery mq = q;
    mq.flags = dtFlip(q.flags);
    return basePublicEvaluate(L, mq, V3(dir.x, dir.y, -dir.z));
}
template <int NC>
__device__ __forceinline__ float lobePdfInternal(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    if (!L.inverse) return basePdf(L, q, dir);
    return basePublicPdf(L, q, V3(dir.x, dir.y, -dir.z));
}
template <int NC>
__device__ __forceinline__ float lobeWeight(const Lobe<NC>& L, const BsdfQuery& q) {    // BSDF::weight (public)
    if (!dtMatches(lobeDirType(L), q.flags)) return 0.0f;
    if (!L.inverse) return baseWeight(L, q) * snWeightCorrectionOf(q);
    BsdfQuery mq = q;
    mq.flags = dtFlip(q.flags);
    return basePublicWeight(L, mq) * snWeightCorrectionOf(q);       // InverseBSDF::weightInternal inside BSDF::weight: corrected twice (kept)
}

// ---------------------------------------------------------------------------------------------
// the BSDF of a hit: one lobe, or a MultiBSDF over up to ML lobes
// ---------------------------------------------------------------------------------------------
template <int NC, int ML> struct Bsdf {
    Lobe<NC> lobes[ML];
    int numLobes;
    bool multi;        // wrapped in a MultiBSDF (sum / mix materials), even with one lobe
    uint32_t type;     // m_type: the lobe's, or the union over the lobes
};

template <int NC, int ML>
__device__ __forceinline__ bool bsdfHasNonDelta(const Bsdf<NC, ML>& b) { return dtMatches(b.type, DT_WholeSphere | DT_NonDelta); }

template <int NC, int ML>
__device__ inline Spec<NC> bsdfSample(const Bsdf<NC, ML>& b, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    res->pdf = 0.0f; res->type = 0; res->dir = V3(0, 0, 1);
    if (!dtMatches(b.type, q.flags)) return specZero<NC>();
    Spec<NC> value;
    if (!b.multi) {
        value = lobeSampleInternal(b.lobes[0], q, uComp, u0, u1, res);
    } else {
        // MultiBSDF::sampleInternalNoRev
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
            if (!found) idx = 0;         // base keeps the last prefix, as the reference's loop leaves it
        }
        if (sumWeights == 0.0f) { res->pdf = 0.0f; return specZero<NC>(); }
        float wSel = 0.0f;
#pragma unroll
        for (int i = 0; i < ML; ++i) if (i == idx) wSel = weights[i];
        const float uc = (uComp * sumWeights - base) / wSel;
        value = specZero<NC>();
#pragma unroll
        for (int i = 0; i < ML; ++i) if (i == idx) value = lobeSampleInternal(b.lobes[i], q, uc, u0, u1, res);
        res->pdf *= wSel;
        if (res->pdf == 0.0f) return specZero<NC>();
        if (!dtIsDelta(res->type)) {
#pragma unroll
            for (int i = 0; i < ML; ++i)
                if (i < b.numLobes && i != idx && dtMatches(lobeDirType(b.lobes[i]), q.flags))
                    res->pdf += lobePdfInternal(b.lobes[i], q, res->dir) * weights[i];
            BsdfQuery mq = q;
            mq.flags &= sideTest(q.gn, q.dir, res->dir);
            value = specZero<NC>();
#pragma unroll
            for (int i = 0; i < ML; ++i)
                if (i < b.numLobes && dtMatches(lobeDirType(b.lobes[i]), mq.flags))
                    value = value + lobeEvaluateInternal(b.lobes[i], mq, res->dir);
        }
        res->pdf /= sumWeights;
    }
    const float snCorrection = snCorrectionOf(q, res->dir);
    return value * snCorrection;
}

// ---- a hit whose BSDF is ONE base lobe of compile-time type LT (no InverseBSDF, no MultiBSDF) ----
template <int NC, int LT>
__device__ __forceinline__ Spec<NC> lobeSample(const Lobe<NC>& L, const BsdfQuery& q, float uComp, float u0, float u1, BsdfSampleResult* res) {
    res->pdf = 0.0f; res->type = 0; res->dir = V3(0, 0, 1);
    if (!dtMatches(L.baseDirType, q.flags)) return specZero<NC>();
    const Spec<NC> value = baseSampleT<NC, LT>(L, q, uComp, u0, u1, res);
    const float snCorrection = snCorrectionOf(q, res->dir);
    return value * snCorrection;
}
template <int NC, int LT>
__device__ __forceinline__ Spec<NC> lobeEvaluate(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    BsdfQuery mq = q;
    mq.flags &= sideTest(q.gn, q.dir, dir);
    if (!dtMatches(L.baseDirType, mq.flags)) return specZero<NC>();
    const Spec<NC> fs = baseEvaluateT<NC, LT>(L, mq, dir);
    const float snCorrection = snCorrectionOf(q, dir);
    return fs * snCorrection;
}
template <int NC, int LT>
__device__ __forceinline__ float lobePdf(const Lobe<NC>& L, const BsdfQuery& q, const V3& dir) {
    if (!dtMatches(L.baseDirType, q.flags)) return 0.0f;
    return basePdfT<NC, LT>(L, q, dir);
}

template <int NC, int ML>
__device__ inline Spec<NC> bsdfEvaluate(const Bsdf<NC, ML>& b, const BsdfQuery& q, const V3& dir) {
    BsdfQuery mq = q;
    mq.flags &= sideTest(q.gn, q.dir, dir);
    if (!dtMatches(b.type, mq.flags)) return specZero<NC>();
    Spec<NC> fs;
    if (!b.multi) {
        fs = lobeEvaluateInternal(b.lobes[0], mq, dir);
    } else {
        fs = specZero<NC>();
#pragma unroll
        for (int i = 0; i < ML; ++i)
            if (i < b.numLobes && dtMatches(lobeDirType(b.lobes[i]), mq.flags))
                fs = fs + lobeEvaluateInternal(b.lobes[i], mq, dir);
    }
    const float snCorrection = snCorrectionOf(q, dir);
    return fs * snCorrection;
}

template <int NC, int ML>
__device__ inline float bsdfPdf(const Bsdf<NC, ML>& b, const BsdfQuery& q, const V3& dir) {
    if (!dtMatches(b.type, q.flags)) return 0.0f;
    if (!b.multi) return lobePdfInternal(b.lobes[0], q, dir);
    // MultiBSDF::evaluatePDFInternalNoRev
    float weights[ML];
    float sum = 0.0f, comp = 0.0f;
#pragma unroll
    for (int i = 0; i < ML; ++i) {
        weights[i] = i < b.numLobes ? lobeWeight(b.lobes[i], q) : 0.0f;
        if (i < b.numLobes) { const float y = weights[i] - comp; const float t = sum + y; comp = (t - sum) - y; sum = t; }
    }
    if (sum == 0.0f) return 0.0f;
    float ret = 0.0f;
#pragma unroll
    for (int i = 0; i < ML; ++i)
        if (i < b.numLobes && weights[i] > 0) ret += lobePdfInternal(b.lobes[i], q, dir) * weights[i];
    return ret / sum;
}

}  // namespace slrgpu
