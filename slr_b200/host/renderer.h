// Renderer front end: the classes a libSLR user programs against, re-implemented over the GPU path.
//   Renderer / render(scene, settings)   libSLR/Core/Renderer.h:15-19
//   RenderSettings                       libSLR/Core/RenderSettings.h:14-41
//   ImageSensor                          libSLR/Core/ImageSensor.h:18-56, ImageSensor.cpp
//   RenderingContext                     libSLRSceneGraph/Scene.h:31-44
//   GPUPathTracingRenderer               drop-in for PathTracingRenderer (Renderers/PathTracingRenderer.cpp:27-98)
#pragma once
#include "scene.h"
#include <map>
#include <memory>
#include <string>

namespace slr {

enum class RenderSettingItem { ImageWidth, ImageHeight, TimeStart, TimeEnd, Brightness, RNGSeed };

class RenderSettings {
    std::map<RenderSettingItem, bool> m_bools;
    std::map<RenderSettingItem, int32_t> m_ints;
    std::map<RenderSettingItem, float> m_floats;
    std::map<RenderSettingItem, std::string> m_strings;
public:
    void addItem(RenderSettingItem k, bool v) { m_bools[k] = v; }
    void addItem(RenderSettingItem k, int32_t v) { m_ints[k] = v; }
    void addItem(RenderSettingItem k, float v) { m_floats[k] = v; }
    void addItem(RenderSettingItem k, const std::string& v) { m_strings[k] = v; }
    // a missing key throws std::out_of_range, like the reference's map::at
    bool getBool(RenderSettingItem k) const { return m_bools.at(k); }
    int32_t getInt(RenderSettingItem k) const { return m_ints.at(k); }
    float getFloat(RenderSettingItem k) const { return m_floats.at(k); }
    std::string getString(RenderSettingItem k) const { return m_strings.at(k); }
};

// Frame buffer of un-normalised per-pixel sums: 16 spectral strata (or 3 RGB channels) per pixel.
// The GPU accumulates in plain fp32; this host copy is what callers read through pixel() and what
// saveImage() tone-maps (strata -> XYZ -> sRGB, 1-exp(-Y), gamma, 24-bit BMP bottom-up).
class ImageSensor {
    uint32_t m_width = 0, m_height = 0, m_channels = 16;
    float m_sensitivity = 1.0f;
    std::vector<float> m_data;
    float* m_external = nullptr;      // caller-owned storage the sensor renders into instead of m_data (bindExternal)
public:
    explicit ImageSensor(float sensitivity = 1.0f) : m_sensitivity(sensitivity) {}
    void init(uint32_t width, uint32_t height, uint32_t channels = 16);
    // Like init() but WITHOUT clearing: for a renderer that overwrites every value (the GPU path downloads a whole frame).
    void initForOverwrite(uint32_t width, uint32_t height, uint32_t channels = 16);
    // The next init*/render uses `storage` (width*height*channels floats, caller-owned, must outlive the use of the
    // sensor) instead of the sensor's own vector: saves a frame-sized copy when the caller wants the data anyway.
    void bindExternal(float* storage) { m_external = storage; }
    void clear();
    uint32_t width() const { return m_width; }
    uint32_t height() const { return m_height; }
    uint32_t channels() const { return m_channels; }
    uint32_t tileWidth() const { return 8; }
    uint32_t tileHeight() const { return 8; }
    uint32_t numTileX() const { return (m_width + 7) >> 3; }
    uint32_t numTileY() const { return (m_height + 7) >> 3; }
    const float* pixel(uint32_t x, uint32_t y) const { return data() + ((size_t)y * m_width + x) * m_channels; }
    float* data() { return m_external ? m_external : m_data.data(); }
    const float* data() const { return m_external ? m_external : m_data.data(); }
    // linear sRGB of one pixel scaled by `scale` (before tone mapping)
    void pixelRGB(uint32_t x, uint32_t y, float scale, float rgb[3]) const;
    void saveImage(const std::string& path, float scale) const;
    float sensitivity() const { return m_sensitivity; }
    void setSensitivity(float s) { m_sensitivity = s; }
};

// The flattened scene as handed to a renderer (what SLR::Scene is in the reference).
struct RenderScene {
    FlatScene flat;
    std::shared_ptr<ImageSensor> sensor;
    ImageSensor* getSensor() const { return sensor.get(); }
};

struct RenderStatistics {
    uint64_t paths = 0, rays = 0;
    double deviceSeconds = 0.0, wallSeconds = 0.0, uploadSeconds = 0.0;
    uint32_t devices = 1;
};

class Renderer {
public:
    virtual ~Renderer() {}
    virtual void render(const RenderScene& scene, const RenderSettings& settings) const = 0;
};

// Unidirectional path tracing on the GPU through include/slrgpu.h. Throws std::runtime_error when
// no CUDA device is available or a GPU call fails -- there is no CPU fallback.
class GPUPathTracingRenderer : public Renderer {
    uint32_t m_samplesPerPixel;
public:
    // Global index of the first sample this renderer instance renders: a multi-process front end gives
    // process g of N the range [sampleBegin, sampleBegin + spp) and sums the sensors.
    uint32_t sampleBegin = 0;
    mutable RenderStatistics lastStatistics;
    bool exportProgressiveImages = true;     // NNN.bmp at 1, 2, 4, ... samples like the reference
    std::string outputDirectory = ".";
    // First device and number of devices the frame's samples are partitioned over (slrgpu_render_multi: scene replicated,
    // sample ranges split, accumulation buffers summed over NVLink onto `device`). deviceCount 0 = every visible device
    // from `device` on -- what a scene file's setRenderer("PT") gets; 1 = the single device `device`.
    int device = 0;
    int deviceCount = 0;
    // true: every sample is a bidirectional one (SLRGPU_RENDER_BPT) -- what GPUBidirectionalPathTracingRenderer sets
    bool bidirectional = false;
    explicit GPUPathTracingRenderer(uint32_t spp) : m_samplesPerPixel(spp) {}
    uint32_t samplesPerPixel() const { return m_samplesPerPixel; }
    void render(const RenderScene& scene, const RenderSettings& settings) const override;
};

// Drop-in for BidirectionalPathTracingRenderer (Renderers/BidirectionalPathTracingRenderer.cpp:25-100): the same pass loop,
// export cadence, sample partition over devices and sensor as the path tracer above, with bidirectional samples
// (csrc/bpt.cu). What setRenderer("BPT") creates.
class GPUBidirectionalPathTracingRenderer : public GPUPathTracingRenderer {
public:
    explicit GPUBidirectionalPathTracingRenderer(uint32_t spp) : GPUPathTracingRenderer(spp) { bidirectional = true; }
};

// The debug (AOV) renderer on the GPU: DebugRenderer (libSLR/Renderers/DebugRenderer.h/.cpp) -- one camera sample per
// pixel, the hit's geometric normal / shading normal / shading tangent quantised into <outputDirectory>/
// geometric_normal.bmp, shading_normal.bmp, shading_tangent.bmp like the reference's chImages. The "distance"
// channel is SLRAssert_NotImplemented in the reference (DebugRenderer.cpp:187) and is ignored here too.
class GPUDebugRenderer : public Renderer {
public:
    enum Channel { GeometricNormal = 0, ShadingNormal, ShadingTangent, Distance, NumChannels };
    bool channels[NumChannels] = {false, false, false, false};
    std::string outputDirectory = ".";
    int device = 0;
    mutable RenderStatistics lastStatistics;
    // caller-owned buffer of width*height*SLRGPU_DEBUG_FLOATS floats that receives the raw vectors (optional)
    float* rawOutput = nullptr;
    explicit GPUDebugRenderer(const bool flags[NumChannels]) { for (int i = 0; i < NumChannels; ++i) channels[i] = flags[i]; }
    void render(const RenderScene& scene, const RenderSettings& settings) const override;
};

struct RenderingContext {
    std::unique_ptr<Renderer> renderer;
    std::string rendererMethod;          // "PT", "BPT", "debug" as written in the scene file
    uint32_t samples = 8;
    int32_t width = 1024, height = 1024;
    float timeStart = 0.0f, timeEnd = 0.0f, brightness = 1.0f;
    int32_t rngSeed = 1509761209;
};

void saveBMP(const std::string& path, const uint8_t* bottomUpBGR, uint32_t width, uint32_t height);

}  // namespace slr
