// Host-side spectral data: the published tables (CIE CMFs, D65, ColorChecker, IOR library, the
// Meng-Simon up-sampling grid) loaded from slr_b200/data/spectral_tables.bin, the input-spectrum
// classes of the scene language, and the 16-strata CMF integration done at start-up.
//   InputSpectrum hierarchy   libSLR/BasicTypes/SpectrumTypes.h:66-346 (Regular / Irregular / Upsampled)
//   initSpectrum / strata     libSLR/BasicTypes/Spectrum.cpp:222, SpectrumTypes.h:746-795
//   colour-space conversion   libSLR/BasicTypes/Spectrum.h:52-178, SpectrumTypes.h:180-237
// Evaluation of spectra at wavelengths happens on the GPU; the host only builds the descriptors.
#pragma once
#include "../../include/slrgpu.h"
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace slr {

enum class SpectrumType : uint32_t { Reflectance = 0, Illuminant = 1, IndexOfRefraction = 2 };
enum class ColorSpace { sRGB, sRGB_NonLinear, XYZ, xyY };

constexpr float kWavelengthLowBound = 360.0f;
constexpr float kWavelengthHighBound = 830.0f;
constexpr uint32_t kNumCMFSamples = 471;
constexpr float kEqualEnergyReflectance = 0.009355121400914532f;

// Tagged table blob (see oracle/drivers/table_io.h for the format).
class SpectralTables {
public:
    struct Entry { uint32_t dtype; std::vector<float> f; std::vector<uint8_t> b; };
    static const SpectralTables& instance();          // loads on first use; throws if the blob is missing
    const std::vector<float>& floats(const std::string& name) const;
    const std::vector<uint8_t>& bytes(const std::string& name) const;
    bool has(const std::string& name) const { return m_entries.count(name) > 0; }
    std::vector<std::string> iorNames() const;
    // derived at load time (initSpectrum)
    float xbar16[16], ybar16[16], zbar16[16];
    float integralCMF;
    // Meng-Simon grid repacked for the GPU: per cell 8 floats-worth of bytes -> 8 uint32 words
    // {inside, numPoints, idx[6]}; points: 186 x (xystar[2], uv[2], spectrum[95]) floats as in the table.
    std::vector<float> upsampleGridWords;     // 168 cells x 8 values stored as floats holding small integers
private:
    std::map<std::string, Entry> m_entries;
    void load(const std::string& path);
    void integrateCMFs();
};

class InputSpectrum {
public:
    virtual ~InputSpectrum() {}
    virtual std::shared_ptr<InputSpectrum> createScaled(float scale) const = 0;
    // Appends this spectrum's descriptor (+ sample data) to the GPU tables; returns the spectrum id.
    virtual uint32_t exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const = 0;
};
typedef std::shared_ptr<InputSpectrum> InputSpectrumRef;

class RegularContinuousSpectrum : public InputSpectrum {
public:
    float minLambda, maxLambda;
    std::vector<float> values;
    RegularContinuousSpectrum(float lo, float hi, const float* v, uint32_t n) : minLambda(lo), maxLambda(hi), values(v, v + n) {}
    InputSpectrumRef createScaled(float scale) const override;
    uint32_t exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const override;
};

class IrregularContinuousSpectrum : public InputSpectrum {
public:
    std::vector<float> lambdas, values;
    IrregularContinuousSpectrum(const float* l, const float* v, uint32_t n) : lambdas(l, l + n), values(v, v + n) {}
    InputSpectrumRef createScaled(float scale) const override;
    uint32_t exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const override;
};

class UpsampledContinuousSpectrum : public InputSpectrum {
public:
    float u, v, scale;
    UpsampledContinuousSpectrum(float uu, float vv, float ss) : u(uu), v(vv), scale(ss) {}
    UpsampledContinuousSpectrum(SpectrumType type, ColorSpace space, float e0, float e1, float e2);
    InputSpectrumRef createScaled(float s) const override;
    uint32_t exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const override;
};

// RGB rendering mode (the reference's #undef Use_Spectral_Representation build): a plain triple.
class RGBInputSpectrum : public InputSpectrum {
public:
    float r, g, b;
    RGBInputSpectrum(float rr, float gg, float bb) : r(rr), g(gg), b(bb) {}
    InputSpectrumRef createScaled(float s) const override;
    uint32_t exportTo(std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) const override;
};

float sRGB_gamma(float v);      // Spectrum.cpp:15-20
float sRGB_degamma(float v);    // Spectrum.cpp:22-28

// Spectrum::create of libSLRSceneGraph/API.cpp:1148-1370; `rgbMode` picks the RGB-build branch.
namespace Spectrum {
InputSpectrumRef create(bool rgbMode, SpectrumType type, ColorSpace space, float e0, float e1, float e2);
InputSpectrumRef create(bool rgbMode, SpectrumType type, float minLambda, float maxLambda, const float* values, uint32_t n);
InputSpectrumRef create(bool rgbMode, SpectrumType type, const float* lambdas, const float* values, uint32_t n);
}

}  // namespace slr
