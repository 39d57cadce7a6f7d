// Spectral arithmetic on the device: 16 stratified wavelengths per path and evaluation of the three
// input-spectrum kinds at those wavelengths.
//   WavelengthSamples::createWithEqualOffsets   libSLR/BasicTypes/SpectrumTypes.h:54-64
//   Regular / Irregular / Upsampled evaluate    SpectrumTypes.h:92-110, 141-160, 239-339
//   importance()                                SpectrumTypes.h:512-526
//   RGB mode twins                              libSLR/BasicTypes/RGBTypes.h
// NC = number of components carried per path: 16 (spectral) or 3 (RGB mode).
#pragma once
#include "device_scene.h"

namespace slrgpu {

constexpr float kWlLow = 360.0f, kWlHigh = 830.0f;

template <int NC> struct Spec {
    float v[NC];
    __device__ __forceinline__ float& operator[](int i) { return v[i]; }
    __device__ __forceinline__ float operator[](int i) const { return v[i]; }
};

template <int NC> __device__ __forceinline__ Spec<NC> specConst(float c) {
    Spec<NC> s;
#pragma unroll
    for (int i = 0; i < NC; ++i) s.v[i] = c;
    return s;
}
template <int NC> __device__ __forceinline__ bool specIsZero(const Spec<NC>& s) {
    bool z = true;
#pragma unroll
    for (int i = 0; i < NC; ++i) z = z && (s.v[i] == 0.0f);
    return z;
}
template <int NC> __device__ __forceinline__ float specLuminance(const Spec<NC>& s) {
    if (NC == 3) return 0.222485f * s.v[0] + 0.716905f * s.v[1] + 0.060610f * s.v[2];     // RGBTypes.h luminance
    float sum = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) sum += s.v[i];
    return sum / NC;                                                                       // SpectrumTypes.h:504-509
}
// importance(): 0.9 on the selected (hero) component, the rest spread uniformly
// (SpectrumTypes.h:512-526, RGBTypes.h:103-108)
template <int NC> __device__ __forceinline__ float specImportance(const Spec<NC>& s, uint32_t hero) {
    float sum = 0, sel = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) { sum += s.v[i]; sel = (i == (int)hero) ? s.v[i] : sel; }
    const float primary = 0.9f;
    const float marginal = (1 - primary) / (NC - 1);
    return sum * marginal + sel * (primary - marginal);
}
template <int NC> __device__ __forceinline__ float specAt(const Spec<NC>& s, uint32_t idx) {
    float sel = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) sel = (i == (int)idx) ? s.v[i] : sel;
    return sel;
}
template <int NC> __device__ __forceinline__ Spec<NC> operator*(const Spec<NC>& a, const Spec<NC>& b) {
    Spec<NC> r;
#pragma unroll
    for (int i = 0; i < NC; ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}
template <int NC> __device__ __forceinline__ Spec<NC> operator+(const Spec<NC>& a, const Spec<NC>& b) {
    Spec<NC> r;
#pragma unroll
    for (int i = 0; i < NC; ++i) r.v[i] = a.v[i] + b.v[i];
    return r;
}
template <int NC> __device__ __forceinline__ Spec<NC> operator*(const Spec<NC>& a, float s) {
    Spec<NC> r;
#pragma unroll
    for (int i = 0; i < NC; ++i) r.v[i] = a.v[i] * s;
    return r;
}
template <int NC> __device__ __forceinline__ Spec<NC> operator*(float s, const Spec<NC>& a) { return a * s; }

// wavelength i of a path with stratification offset `off`
__device__ __forceinline__ float wavelengthOf(int i, float off) { return kWlLow + (kWlHigh - kWlLow) * (i + off) / 16; }

// lower_bound over [base, n), then the reference's clamping; *base carries the running search start
__device__ __forceinline__ float evalIrregular(const float* __restrict__ lambdas, const float* __restrict__ values, uint32_t n,
                                               float wl, uint32_t* base) {
    uint32_t lo = *base, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(lambdas + mid) < wl) lo = mid + 1; else hi = mid;
    }
    const int lowIdx = max((int)lo - 1, 0);
    *base = (uint32_t)lowIdx;
    if (lowIdx >= (int)n - 1) return __ldg(values + n - 1);
    const float l0 = __ldg(lambdas + lowIdx), l1 = __ldg(lambdas + lowIdx + 1);
    const float t = (wl - l0) / (l1 - l0);
    if (t <= 0.0f) return __ldg(values);
    return (1 - t) * __ldg(values + lowIdx) + t * __ldg(values + lowIdx + 1);
}

// Meng-Simon up-sampling: locate the grid cell of (u, v), pick the 4 bilinear (inside cell) or 3
// barycentric (fan triangulation on the locus boundary) data points and their weights.
struct UpsampleWeights {
    int n;                  // 0 (outside the grid), 3 or 4
    uint32_t idx[4];
    float w[4];
};

constexpr int kUpGridW = 12, kUpGridH = 14, kUpNumWl = 95, kUpPointStride = 99;   // xystar[2] uv[2] spectrum[95]

static __device__ __noinline__ UpsampleWeights upsampleWeights(const DeviceScene& s, float u, float v) {
    UpsampleWeights r;
    r.n = 0;
    if (u < 0.0f || u >= kUpGridW || v < 0.0f || v >= kUpGridH) return r;
    const int ui = (int)u, vi = (int)v;
    const float* cell = s.upsampleGrid + (ui + kUpGridW * vi) * 8;      // inside, numPoints, idx[6]
    const bool inside = __ldg(cell) != 0.0f;
    const int numPoints = (int)__ldg(cell + 1);
    if (inside) {
        const float sx = u - ui, ty = v - vi;
        r.w[0] = (1 - sx) * (1 - ty); r.w[1] = sx * (1 - ty); r.w[2] = (1 - sx) * ty; r.w[3] = sx * ty;
#pragma unroll
        for (int i = 0; i < 4; ++i) r.idx[i] = (uint32_t)__ldg(cell + 2 + i);
        r.n = 4;
        return r;
    }
    const uint32_t i0 = (uint32_t)__ldg(cell + 2);
    const float* p0 = s.upsamplePoints + i0 * kUpPointStride;
    const float p0u = __ldg(p0 + 2), p0v = __ldg(p0 + 3);
    const float ex = u - p0u, ey = v - p0v;
    const uint32_t i1 = (uint32_t)__ldg(cell + 3);
    float e0x = __ldg(s.upsamplePoints + i1 * kUpPointStride + 2) - p0u;
    float e0y = __ldg(s.upsamplePoints + i1 * kUpPointStride + 3) - p0v;
    float uu = e0x * ey - ex * e0y;
    for (int i = 1; i < numPoints; ++i) {
        const uint32_t idx = (uint32_t)__ldg(cell + 2 + (i % (numPoints - 1) + 1));
        const float e1x = __ldg(s.upsamplePoints + idx * kUpPointStride + 2) - p0u;
        const float e1y = __ldg(s.upsamplePoints + idx * kUpPointStride + 3) - p0v;
        const float vv = ex * e1y - e1x * ey;
        const float area = e0x * e1y - e1x * e0y;
        const float bu = uu / area, bv = vv / area;
        const float bw = 1.0f - bu - bv;
        if (bu < -1e-6 || bv < -1e-6 || bw < -1e-6) { uu = -vv; e0x = e1x; e0y = e1y; continue; }
        r.w[0] = bu; r.w[1] = bv; r.w[2] = bw;
        r.idx[0] = idx; r.idx[1] = (uint32_t)__ldg(cell + 2 + i); r.idx[2] = i0;
        r.n = 3;
        break;
    }
    return r;
}

// UpsampledContinuousSpectrum::evaluate at the path's 16 wavelengths. The fractional bin (lambda_i - 360) / 470 * 94
// is (i + offset) * 94 / 16: one multiply per wavelength instead of the subtraction / division / multiplication chain.
template <int NC>
__device__ __forceinline__ void evalUpsampledAll(const DeviceScene& s, const UpsampleWeights& w, float scale, float wlOffset, Spec<NC>* out) {
#pragma unroll
    for (int i = 0; i < NC; ++i) out->v[i] = 0.0f;
    if (w.n == 0) return;
    // always four points (the fourth with weight 0 on the three-point boundary cells): straight-line code, no inner trip count
    const float* sp[4];
    float wj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sp[j] = s.upsamplePoints + w.idx[j < w.n ? j : 0] * kUpPointStride + 4;
        wj[j] = (j < w.n ? w.w[j] : 0.0f) * scale;
    }
#pragma unroll 4
    for (int i = 0; i < NC; ++i) {
        const float sBinF = ((float)i + wlOffset) * ((float)(kUpNumWl - 1) / 16.0f);
        const uint32_t sBin = min((uint32_t)sBinF, (uint32_t)(kUpNumWl - 1));
        const uint32_t sNext = min(sBin + 1u, (uint32_t)(kUpNumWl - 1));
        const float t = sBinF - (float)sBin;
        float ret = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = __ldg(sp[j] + sBin), b = __ldg(sp[j] + sNext);
            ret = fmaf(wj[j], fmaf(t, b - a, a), ret);
        }
        out->v[i] = ret;
    }
}

// compiled irregular spectrum (scene.cu compileSpectra): the bin table gives the interval the reference's
// lower_bound would find at the start of the wavelength's 2 nm bin; walk forward to the wavelength itself
__device__ __forceinline__ float evalIrregularLut(const float* __restrict__ lambdas, const float* __restrict__ values, uint32_t n,
                                                  const uint32_t* __restrict__ lut, float wl) {
    int b = (int)((wl - kWlLow) * 0.5f);
    b = min(max(b, 0), SLRGPU_SPECTRUM_LUT_BINS - 1);
    uint32_t lowIdx = (__ldg(lut + (b >> 2)) >> (8 * (b & 3))) & 0xFFu;
    // lower_bound(wl) - 1 = last knot strictly below wl
    while (lowIdx + 1 < n && __ldg(lambdas + lowIdx + 1) < wl) ++lowIdx;
    if (lowIdx >= n - 1) return __ldg(values + n - 1);
    const float l0 = __ldg(lambdas + lowIdx), l1 = __ldg(lambdas + lowIdx + 1);
    const float t = (wl - l0) / (l1 - l0);
    if (t <= 0.0f) return __ldg(values);
    return (1 - t) * __ldg(values + lowIdx) + t * __ldg(values + lowIdx + 1);
}

// Evaluates input spectrum `id` at the path's wavelengths (or returns the RGB triple in RGB mode).
// The table was compiled at scene creation: up-sampled spectra arrive as REGULAR, irregular ones with
// their look-up table; the other kinds are evaluated as the reference does.
template <int NC>
static __device__ __noinline__ Spec<NC> evalInputSpectrum(const DeviceScene& s, uint32_t id, float wlOffset) {
    const SlrGpuSpectrum sp = s.spectra[id];
    Spec<NC> out;
    if (NC == 3) {
        out.v[0] = sp.p0; out.v[1] = sp.p1; out.v[2] = sp.p2;
        return out;
    }
    if (sp.kind == SLRGPU_SPECTRUM_REGULAR) {
        // RegularContinuousSpectrum::evaluate at the path's 16 equally spaced wavelengths: the fractional bin
        // (lambda_i - lo) / (hi - lo) * (n - 1) is linear in i, so one FMA per wavelength replaces its subtraction,
        // IEEE division and multiplication (ncu on the Lambert kernel: 21 % of its stall samples sat on those three
        // lines). The two end clamps become index / weight clamps of the same piecewise-linear function.
        const float* vals = s.spectrumData + sp.data_offset;
        const int last = (int)sp.num_samples - 1;
        if (last < 1) {
            const float v = __ldg(vals);
#pragma unroll
            for (int i = 0; i < NC; ++i) out.v[i] = v;
        } else {
            const float perNm = (float)last / (sp.p1 - sp.p0);
            const float step = (kWlHigh - kWlLow) / 16 * perNm;
            const float first = (kWlLow - sp.p0 + (kWlHigh - kWlLow) / 16 * wlOffset) * perNm;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const float binF = fmaf(step, (float)i, first);
                const int bin = min(max((int)binF, 0), last - 1);
                const float t = __saturatef(binF - (float)bin);
                const float v0 = __ldg(vals + bin), v1 = __ldg(vals + bin + 1);
                out.v[i] = fmaf(t, v1 - v0, v0);
            }
        }
    } else if (sp.kind == SLRGPU_SPECTRUM_IRREGULAR_LUT) {
        const float* lam = s.spectrumData + sp.data_offset;
        const float* vals = lam + sp.num_samples;
        const uint32_t* lut = reinterpret_cast<const uint32_t*>(vals + sp.num_samples);
#pragma unroll 4
        for (int i = 0; i < NC; ++i) out.v[i] = evalIrregularLut(lam, vals, sp.num_samples, lut, wavelengthOf(i, wlOffset));
    } else if (sp.kind == SLRGPU_SPECTRUM_IRREGULAR) {
        const float* lam = s.spectrumData + sp.data_offset;
        const float* vals = lam + sp.num_samples;
        uint32_t base = 0;
#pragma unroll 1
        for (int i = 0; i < NC; ++i) out.v[i] = evalIrregular(lam, vals, sp.num_samples, wavelengthOf(i, wlOffset), &base);
    } else {
        const UpsampleWeights w = upsampleWeights(s, sp.p0, sp.p1);
        evalUpsampledAll<NC>(s, w, sp.p2, wlOffset, &out);
    }
    return out;
}

// UpsampledContinuousSpectrum built per texel / per Voronoi cell from (u, v, scale)
template <int NC>
static __device__ __noinline__ Spec<NC> evalUVS(const DeviceScene& s, float u, float v, float scale, float wlOffset) {
    Spec<NC> out;
    const UpsampleWeights w = upsampleWeights(s, u, v);
    evalUpsampledAll<NC>(s, w, scale, wlOffset, &out);
    return out;
}

}  // namespace slrgpu
