#include "spectrum.h"
#include "../../include/slrgpu.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <stdexcept>

namespace slr {

// ---------------------------------------------------------------------------------------------
// table blob
// ---------------------------------------------------------------------------------------------

static std::string defaultTablePath() {
    if (const char* env = std::getenv("SLR_B200_DATA")) return std::string(env) + "/spectral_tables.bin";
    Dl_info info;
    if (dladdr((const void*)&defaultTablePath, &info) && info.dli_fname) {
        std::string so = info.dli_fname;                        // .../slr_b200/lib/libslrhost.so
        size_t p = so.find_last_of('/');
        std::string dir = p == std::string::npos ? "." : so.substr(0, p);
        return dir + "/../data/spectral_tables.bin";
    }
    return "slr_b200/data/spectral_tables.bin";
}

const SpectralTables& SpectralTables::instance() {
    static SpectralTables* t = nullptr;
    if (!t) {
        SpectralTables* n = new SpectralTables();
        n->load(defaultTablePath());
        n->integrateCMFs();
        t = n;
    }
    return *t;
}

void SpectralTables::load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open spectral table blob: " + path);
    char magic[4];
    uint32_t count = 0;
    bool ok = std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "SLRT", 4) == 0 && std::fread(&count, 4, 1, f) == 1;
    for (uint32_t i = 0; ok && i < count; ++i) {
        uint32_t len = 0, dtype = 0;
        uint64_t n = 0;
        ok = std::fread(&len, 4, 1, f) == 1 && len < 256;
        std::string name(len, '\0');
        ok = ok && std::fread(&name[0], 1, len, f) == len && std::fread(&dtype, 4, 1, f) == 1 && std::fread(&n, 8, 1, f) == 1;
        if (!ok) break;
        Entry& e = m_entries[name];
        e.dtype = dtype;
        if (dtype == 0) { e.f.resize(n); ok = std::fread(e.f.data(), 4, n, f) == n; }
        else            { e.b.resize(n); ok = std::fread(e.b.data(), 1, n, f) == n; }
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("corrupt spectral table blob: " + path);
    // repack the grid cells (8 bytes each: inside, numPoints, idx[6]) into one float-held integer per byte
    const std::vector<uint8_t>& g = bytes("upsampling/grid");
    upsampleGridWords.resize(g.size());
    for (size_t i = 0; i < g.size(); ++i) upsampleGridWords[i] = (float)g[i];
}

const std::vector<float>& SpectralTables::floats(const std::string& name) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end() || it->second.dtype != 0) throw std::runtime_error("spectral table missing: " + name);
    return it->second.f;
}
const std::vector<uint8_t>& SpectralTables::bytes(const std::string& name) const {
    auto it = m_entries.find(name);
    if (it == m_entries.end() || it->second.dtype != 1) throw std::runtime_error("spectral table missing: " + name);
    return it->second.b;
}
std::vector<std::string> SpectralTables::iorNames() const {
    std::vector<std::string> out;
    for (const auto& kv : m_entries)
        if (kv.first.rfind("ior/", 0) == 0 && kv.first.size() > 9 && kv.first.compare(kv.first.size() - 5, 5, "/meta") == 0)
            out.push_back(kv.first.substr(4, kv.first.size() - 9));
    return out;
}

// Integrates the 1 nm CMFs over the 16 storage strata with the trapezoid rule, splitting the
// trapezoid that straddles a stratum boundary (SpectrumTypes.h:746-795), fp32 throughout.
void SpectralTables::integrateCMFs() {
    const std::vector<float>& xb = floats("cmf/xbar_2deg");
    const std::vector<float>& yb = floats("cmf/ybar_2deg");
    const std::vector<float>& zb = floats("cmf/zbar_2deg");
    const uint32_t numStrata = 16;
    uint32_t bin = 0;
    float nextP = float(bin + 1) / numStrata;
    float xSum = 0, xPrev = xb[0], ySum = 0, yPrev = yb[0], zSum = 0, zPrev = zb[0];
    const float interval = 1;
    for (uint32_t i = 1; i < kNumCMFSamples; ++i) {
        float curP = float(i) / (kNumCMFSamples - 1);
        float width = interval;
        float xCur = xb[i], yCur = yb[i], zCur = zb[i];
        if (curP >= nextP) {
            width = (curP - nextP) * (kWavelengthHighBound - kWavelengthLowBound);
            float t = 1 - width / interval;
            float xIn = xPrev * (1 - t) + xCur * t;
            float yIn = yPrev * (1 - t) + yCur * t;
            float zIn = zPrev * (1 - t) + zCur * t;
            xSum += (xPrev + xIn) * (interval - width) * 0.5f;
            ySum += (yPrev + yIn) * (interval - width) * 0.5f;
            zSum += (zPrev + zIn) * (interval - width) * 0.5f;
            xbar16[bin] = xSum; ybar16[bin] = ySum; zbar16[bin] = zSum;
            xSum = ySum = zSum = 0;
            xPrev = xIn; yPrev = yIn; zPrev = zIn;
            ++bin;
            nextP = float(bin + 1) / numStrata;
        }
        xSum += (xPrev + xCur) * width * 0.5f;
        ySum += (yPrev + yCur) * width * 0.5f;
        zSum += (zPrev + zCur) * width * 0.5f;
        xPrev = xCur; yPrev = yCur; zPrev = zCur;
    }
    integralCMF = 0.0f;
    for (uint32_t i = 0; i < numStrata; ++i) integralCMF += ybar16[i];
}

// ---------------------------------------------------------------------------------------------
// colour helpers
// ---------------------------------------------------------------------------------------------

float sRGB_gamma(float v) {
    if (v <= 0.0031308) return (float)(12.92 * v);
    return (float)(1.055 * std::pow(v, 1.0 / 2.4) - 0.055);
}
float sRGB_degamma(float v) {
    if (v <= 0.04045) return (float)(v / 12.92);
    return (float)std::pow((v + 0.055) / 1.055, 2.4);
}
static void sRGB_to_XYZ(const float rgb[3], float xyz[3]) {
    xyz[0] = (float)(0.4124564 * rgb[0] + 0.3575761 * rgb[1] + 0.1804375 * rgb[2]);
    xyz[1] = (float)(0.2126729 * rgb[0] + 0.7151522 * rgb[1] + 0.0721750 * rgb[2]);
    xyz[2] = (float)(0.0193339 * rgb[0] + 0.1191920 * rgb[1] + 0.9503041 * rgb[2]);
}
static void sRGB_E_to_XYZ(const float rgb[3], float xyz[3]) {
    xyz[0] = (float)(0.4969 * rgb[0] + 0.3391 * rgb[1] + 0.1640 * rgb[2]);
    xyz[1] = (float)(0.2562 * rgb[0] + 0.6782 * rgb[1] + 0.0656 * rgb[2]);
    xyz[2] = (float)(0.0233 * rgb[0] + 0.1130 * rgb[1] + 0.8637 * rgb[2]);
}
static void XYZ_to_sRGB(const float xyz[3], float rgb[3]) {
    rgb[0] = (float)(3.2404542 * xyz[0] - 1.5371385 * xyz[1] - 0.4985314 * xyz[2]);
    rgb[1] = (float)(-0.9692660 * xyz[0] + 1.8760108 * xyz[1] + 0.0415560 * xyz[2]);
    rgb[2] = (float)(0.0556434 * xyz[0] - 0.2040259 * xyz[1] + 1.0572252 * xyz[2]);
}
static void XYZ_to_sRGB_E(const float xyz[3], float rgb[3]) {
    rgb[0] = (float)(2.6897 * xyz[0] - 1.2759 * xyz[1] - 0.4138 * xyz[2]);
    rgb[1] = (float)(-1.0221 * xyz[0] + 1.9783 * xyz[1] + 0.0438 * xyz[2]);
    rgb[2] = (float)(0.0612 * xyz[0] - 0.2245 * xyz[1] + 1.1633 * xyz[2]);
}

// ---------------------------------------------------------------------------------------------
// input spectra
// ---------------------------------------------------------------------------------------------

static uint32_t pushSpectrum(std::vector<SlrGpuSpectrum>& spectra, const SlrGpuSpectrum& s) {
    spectra.push_back(s);
    return (uint32_t)spectra.size() - 1;
}

InputSpectrumRef RegularContinuousSpectrum::createScaled(float scale) const {
    std::vector<float> v(values.size());
    for (size_t i = 0; i < v.size(); ++i) v[i] = scale * values[i];
    return std::make_shared<RegularContinuousSpectrum>(minLambda, maxLambda, v.data(), (uint32_t)v.size());
}
uint32_t RegularContinuousSpectrum::exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const {
    SlrGpuSpectrum s = {};
    s.kind = SLRGPU_SPECTRUM_REGULAR;
    s.data_offset = (uint32_t)data.size();
    s.num_samples = (uint32_t)values.size();
    s.p0 = minLambda; s.p1 = maxLambda;
    data.insert(data.end(), values.begin(), values.end());
    return pushSpectrum(spectra, s);
}

InputSpectrumRef IrregularContinuousSpectrum::createScaled(float scale) const {
    std::vector<float> v(values.size());
    for (size_t i = 0; i < v.size(); ++i) v[i] = scale * values[i];
    return std::make_shared<IrregularContinuousSpectrum>(lambdas.data(), v.data(), (uint32_t)v.size());
}
uint32_t IrregularContinuousSpectrum::exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const {
    SlrGpuSpectrum s = {};
    s.kind = SLRGPU_SPECTRUM_IRREGULAR;
    s.data_offset = (uint32_t)data.size();
    s.num_samples = (uint32_t)values.size();
    data.insert(data.end(), lambdas.begin(), lambdas.end());
    data.insert(data.end(), values.begin(), values.end());
    return pushSpectrum(spectra, s);
}

// Tristimulus -> (u, v, scale) of the Meng et al. grid, following the ctor's fall-through chain:
// non-linear sRGB -> linear sRGB -> XYZ -> xy + brightness -> uv (SpectrumTypes.h:180-237).
UpsampledContinuousSpectrum::UpsampledContinuousSpectrum(SpectrumType type, ColorSpace space, float e0, float e1, float e2) {
    float x = 0, y = 0, brightness = 0;
    if (space == ColorSpace::sRGB_NonLinear) {
        e0 = sRGB_degamma(e0); e1 = sRGB_degamma(e1); e2 = sRGB_degamma(e2);
        space = ColorSpace::sRGB;
    }
    if (space == ColorSpace::sRGB) {
        float rgb[3] = {e0, e1, e2}, xyz[3] = {0, 0, 0};
        if (type == SpectrumType::Reflectance) sRGB_E_to_XYZ(rgb, xyz);
        else if (type == SpectrumType::Illuminant) sRGB_to_XYZ(rgb, xyz);
        else throw std::runtime_error("UpsampledContinuousSpectrum: spectrum type must be Reflectance or Illuminant");
        e0 = xyz[0]; e1 = xyz[1]; e2 = xyz[2];
        space = ColorSpace::XYZ;
    }
    if (space == ColorSpace::XYZ) {
        brightness = e0 + e1 + e2;
        if (brightness == 0) { u = 6; v = 4; scale = 0; return; }
        x = e0 / brightness;
        y = e1 / brightness;
    } else {  // xyY
        x = e0; y = e1; brightness = e2 / e1;
    }
    scale = brightness / kEqualEnergyReflectance;
    u = (float)(16.730260708356887 * x + 7.7801960340706 * y - 2.170152247475828);
    v = (float)(-7.530081094743006 * x + 16.192422314095225 * y + 1.1125529268825947);
}
InputSpectrumRef UpsampledContinuousSpectrum::createScaled(float s) const {
    return std::make_shared<UpsampledContinuousSpectrum>(u, v, scale * s);
}
uint32_t UpsampledContinuousSpectrum::exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>&) const {
    SlrGpuSpectrum s = {};
    s.kind = SLRGPU_SPECTRUM_UPSAMPLED;
    s.p0 = u; s.p1 = v; s.p2 = scale;
    return pushSpectrum(spectra, s);
}

InputSpectrumRef RGBInputSpectrum::createScaled(float s) const { return std::make_shared<RGBInputSpectrum>(r * s, g * s, b * s); }
uint32_t RGBInputSpectrum::exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>&) const {
    SlrGpuSpectrum s = {};
    s.kind = SLRGPU_SPECTRUM_RGB;
    s.p0 = r; s.p1 = g; s.p2 = b;
    return pushSpectrum(spectra, s);
}

// ---------------------------------------------------------------------------------------------
// Spectrum::create
// ---------------------------------------------------------------------------------------------

namespace {

struct KahanF {
    float result = 0, comp = 0;
    void add(float v) { float y = v - comp; float t = result + y; comp = (t - result) - y; result = t; }
};

InputSpectrumRef rgbFromXYZ(SpectrumType type, const float XYZ[3]) {
    float RGB[3];
    if (type == SpectrumType::Illuminant) XYZ_to_sRGB(XYZ, RGB); else XYZ_to_sRGB_E(XYZ, RGB);
    for (int i = 0; i < 3; ++i) RGB[i] = RGB[i] < 0.0f ? 0.0f : RGB[i];
    return std::make_shared<RGBInputSpectrum>(RGB[0], RGB[1], RGB[2]);
}

// Merged-grid trapezoid integration of value(lambda) * CMF(lambda) over [360, 830] nm
// (API.cpp:1148-1275); `sample(curWL, &hitsSampleNode)` returns the spectrum value.
template <typename SampleFn, typename NextFn>
void spectrumToXYZ(SampleFn sample, NextFn nextSampleWL, float XYZ[3]) {
    const SpectralTables& T = SpectralTables::instance();
    const std::vector<float>& xb = T.floats("cmf/xbar_2deg");
    const std::vector<float>& yb = T.floats("cmf/ybar_2deg");
    const std::vector<float>& zb = T.floats("cmf/zbar_2deg");
    KahanF cum;
    for (uint32_t i = 1; i < kNumCMFSamples; ++i) cum.add((float)((yb[i - 1] + yb[i]) * 1 * 0.5));
    const float integralCMF = cum.result;
    const float cmfBin = (kWavelengthHighBound - kWavelengthLowBound) / (kNumCMFSamples - 1);
    uint32_t cmfIdx = 0;
    float curWL = kWavelengthLowBound;
    float px = 0, py = 0, pz = 0, prevValue = 0, halfWidth = 0;
    KahanF X, Y, Z;
    while (true) {
        float xv, yv, zv;
        if (curWL == kWavelengthLowBound + cmfIdx * cmfBin) {
            xv = xb[cmfIdx]; yv = yb[cmfIdx]; zv = zb[cmfIdx];
            ++cmfIdx;
        } else {
            // The reference computes uint32_t((curWL - 360) / bin) also when curWL < 360 (a spectrum whose first
            // sample lies below 360 nm, e.g. D65 from 300 nm, pulls curWL back: API.cpp:1203-1205), which is
            // undefined behaviour and in practice indexes one past the CMF tables. That read is replaced by a
            // clamp to the table: deterministic, and the RGB build is not part of the oracle anyway.
            float rel = (curWL - kWavelengthLowBound) / cmfBin;
            uint32_t idx = rel <= 0.0f ? 0u : std::min(uint32_t(rel), kNumCMFSamples - 1);
            uint32_t idx1 = std::min(idx + 1, kNumCMFSamples - 1);
            float base = kWavelengthLowBound + idx * cmfBin;
            float t = (curWL - base) / cmfBin;
            t = std::min(std::max(t, 0.0f), 1.0f);
            xv = (1 - t) * xb[idx] + t * xb[idx1];
            yv = (1 - t) * yb[idx] + t * yb[idx1];
            zv = (1 - t) * zb[idx] + t * zb[idx1];
        }
        float value = sample(curWL);
        float avg = (prevValue + value) * 0.5f;
        X.add(avg * (px + xv) * halfWidth);
        Y.add(avg * (py + yv) * halfWidth);
        Z.add(avg * (pz + zv) * halfWidth);
        px = xv; py = yv; pz = zv; prevValue = value;
        float prevWL = curWL;
        curWL = std::min(kWavelengthLowBound + cmfIdx * cmfBin, nextSampleWL());
        halfWidth = (curWL - prevWL) * 0.5f;
        if (cmfIdx == kNumCMFSamples) break;
    }
    XYZ[0] = X.result / integralCMF; XYZ[1] = Y.result / integralCMF; XYZ[2] = Z.result / integralCMF;
}

}  // namespace

namespace Spectrum {

InputSpectrumRef create(bool rgbMode, SpectrumType type, ColorSpace space, float e0, float e1, float e2) {
    if (!rgbMode) return std::make_shared<UpsampledContinuousSpectrum>(type, space, e0, e1, e2);
    // RGB build (API.cpp:1277-1325)
    if (space == ColorSpace::sRGB_NonLinear) {
        e0 = sRGB_degamma(e0); e1 = sRGB_degamma(e1); e2 = sRGB_degamma(e2);
        space = ColorSpace::sRGB;
    }
    if (space == ColorSpace::sRGB) return std::make_shared<RGBInputSpectrum>(e0, e1, e2);
    if (space == ColorSpace::xyY) {
        float b = e2 / e1;
        float X = e0 * b, Y = e2, Z = (1.0f - e0 - e1) * b;
        e0 = X; e1 = Y; e2 = Z;
    }
    float XYZ[3] = {e0, e1, e2};
    return rgbFromXYZ(type, XYZ);
}

InputSpectrumRef create(bool rgbMode, SpectrumType type, float minLambda, float maxLambda, const float* values, uint32_t n) {
    if (!rgbMode) return std::make_shared<RegularContinuousSpectrum>(minLambda, maxLambda, values, n);
    const float binWidth = (maxLambda - minLambda) / (n - 1);
    uint32_t baseIdx = 0;
    float XYZ[3];
    spectrumToXYZ(
        [&](float wl) {
            if (wl < minLambda) return values[0];
            if (wl > maxLambda) return values[n - 1];
            if (wl == minLambda + baseIdx * binWidth) return values[baseIdx++];
            uint32_t idx = std::min(uint32_t((wl - minLambda) / binWidth), n - 1);
            float t = (wl - (minLambda + idx * binWidth)) / binWidth;
            return (1 - t) * values[idx] + t * values[std::min(idx + 1, n - 1)];
        },
        [&]() { return baseIdx < n ? (minLambda + baseIdx * binWidth) : INFINITY; }, XYZ);
    return rgbFromXYZ(type, XYZ);
}

InputSpectrumRef create(bool rgbMode, SpectrumType type, const float* lambdas, const float* values, uint32_t n) {
    if (!rgbMode) return std::make_shared<IrregularContinuousSpectrum>(lambdas, values, n);
    uint32_t baseIdx = 0;
    float XYZ[3];
    spectrumToXYZ(
        [&](float wl) {
            if (wl < lambdas[0]) return values[0];
            if (wl > lambdas[n - 1]) return values[n - 1];
            if (wl == lambdas[baseIdx]) return values[baseIdx++];
            const float* lb = std::lower_bound(lambdas + std::max((int32_t)baseIdx - 1, 0), lambdas + n, wl);
            uint32_t idx = (uint32_t)std::max(int32_t(lb - lambdas) - 1, 0);
            uint32_t idx1 = std::min(idx + 1, n - 1);
            if (idx1 == idx) return values[idx];
            float t = (wl - lambdas[idx]) / (lambdas[idx1] - lambdas[idx]);
            return (1 - t) * values[idx] + t * values[idx1];
        },
        [&]() { return baseIdx < n ? lambdas[baseIdx] : INFINITY; }, XYZ);
    return rgbFromXYZ(type, XYZ);
}

}  // namespace Spectrum
}  // namespace slr
