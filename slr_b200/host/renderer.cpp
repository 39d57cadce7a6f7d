#include "renderer.h"
#include "spectrum.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace slr {

void ImageSensor::init(uint32_t width, uint32_t height, uint32_t channels) {
    initForOverwrite(width, height, channels);
    clear();
}
void ImageSensor::initForOverwrite(uint32_t width, uint32_t height, uint32_t channels) {
    m_width = width; m_height = height; m_channels = channels;
    if (!m_external) m_data.resize((size_t)width * height * channels);
}
void ImageSensor::clear() { std::fill(data(), data() + (size_t)m_width * m_height * m_channels, 0.0f); }

void ImageSensor::pixelRGB(uint32_t x, uint32_t y, float scale, float rgb[3]) const {
    const float* p = pixel(x, y);
    if (m_channels == 3) { for (int i = 0; i < 3; ++i) rgb[i] = p[i] * scale; return; }
    const SpectralTables& T = SpectralTables::instance();
    float XYZ[3] = {0, 0, 0};
    for (int i = 0; i < 16; ++i) {
        float v = p[i] * scale;
        XYZ[0] += T.xbar16[i] * v; XYZ[1] += T.ybar16[i] * v; XYZ[2] += T.zbar16[i] * v;
    }
    for (int i = 0; i < 3; ++i) XYZ[i] /= T.integralCMF;
    rgb[0] = (float)(3.2404542 * XYZ[0] - 1.5371385 * XYZ[1] - 0.4985314 * XYZ[2]);
    rgb[1] = (float)(-0.9692660 * XYZ[0] + 1.8760108 * XYZ[1] + 0.0415560 * XYZ[2]);
    rgb[2] = (float)(0.0556434 * XYZ[0] - 0.2040259 * XYZ[1] + 1.0572252 * XYZ[2]);
}

void saveBMP(const std::string& path, const uint8_t* pixels, uint32_t width, uint32_t height) {
    const uint32_t rowBytes = 3 * width + width % 4;       // the reference's (non-standard) row padding
    const uint32_t headerSize = 54, dataSize = rowBytes * height, fileSize = dataSize + headerSize;
    uint8_t h[54];
    std::memset(h, 0, sizeof(h));
    h[0] = 'B'; h[1] = 'M';
    auto put32 = [&h](int at, uint32_t v) { std::memcpy(h + at, &v, 4); };
    auto put16 = [&h](int at, uint16_t v) { std::memcpy(h + at, &v, 2); };
    put32(2, fileSize); put32(10, headerSize); put32(14, 40); put32(18, width); put32(22, height);
    put16(26, 1); put16(28, 24); put32(30, 0); put32(34, dataSize); put32(38, 1); put32(42, 1);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return;
    std::fwrite(h, 1, headerSize, f);
    std::fwrite(pixels, 1, dataSize, f);
    std::fclose(f);
}

void ImageSensor::saveImage(const std::string& path, float scale) const {
    float sens = std::isinf(m_sensitivity) ? 1.0f : m_sensitivity;
    scale *= sens;
    const uint32_t rowBytes = 3 * m_width + m_width % 4;
    std::vector<uint8_t> bmp((size_t)rowBytes * m_height, 0);
    for (uint32_t y = 0; y < m_height; ++y) {
        for (uint32_t x = 0; x < m_width; ++x) {
            float rgb[3];
            pixelRGB(x, y, scale, rgb);
            for (int c = 0; c < 3; ++c) rgb[c] = rgb[c] < 0.0f ? 0.0f : rgb[c];
            float Y = (float)(0.222485 * rgb[0] + 0.716905 * rgb[1] + 0.060610 * rgb[2]);
            float sy = Y != 0 ? (1.0f - std::exp(-Y)) / Y : 0.0f;
            uint8_t* dst = &bmp[(size_t)(m_height - y - 1) * rowBytes + 3 * x];
            for (int c = 0; c < 3; ++c) {
                float v = std::min(sy * rgb[c], 1.0f);
                dst[2 - c] = (uint8_t)(256 * std::min(sRGB_gamma(v), 0.999f));
            }
        }
    }
    saveBMP(path, bmp.data(), m_width, m_height);
}

void GPUPathTracingRenderer::render(const RenderScene& scene, const RenderSettings& settings) const {
    auto wall0 = std::chrono::steady_clock::now();
    ImageSensor* sensor = scene.getSensor();
    if (!sensor) throw std::runtime_error("GPUPathTracingRenderer: the scene has no camera/sensor");
    const uint32_t W = (uint32_t)settings.getInt(RenderSettingItem::ImageWidth);
    const uint32_t H = (uint32_t)settings.getInt(RenderSettingItem::ImageHeight);

    lastStatistics = RenderStatistics();
    SlrGpuSceneDesc desc;
    scene.flat.describe(&desc);
    // the scene goes to every device the frame is split over (replicated)
    const int visible = slrgpu_device_count();
    int count = deviceCount > 0 ? deviceCount : std::max(1, visible - device);
    if (visible > 0 && device + count > visible) count = std::max(1, visible - device);
    // more devices than samples make no sense: a device would render nothing
    count = (int)std::min<uint32_t>((uint32_t)count, std::max(1u, m_samplesPerPixel));
    std::vector<SlrGpuScene*> replicas;
    struct DestroyAll { std::vector<SlrGpuScene*>& r; ~DestroyAll() { for (SlrGpuScene* g : r) slrgpu_scene_destroy(g); } } destroyAll{replicas};
    auto up0 = std::chrono::steady_clock::now();
    for (int g = 0; g < count; ++g) {
        SlrGpuScene* gpu = nullptr;
        if (slrgpu_scene_create(&desc, device + g, &gpu) != SLRGPU_OK)
            throw std::runtime_error(std::string("slrgpu_scene_create failed: ") + slrgpu_last_error());
        replicas.push_back(gpu);
    }
    lastStatistics.uploadSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - up0).count();
    lastStatistics.devices = (uint32_t)count;
    const uint32_t channels = slrgpu_scene_channels(replicas[0]);
    if (exportProgressiveImages) sensor->init(W, H, channels);
    else sensor->initForOverwrite(W, H, channels);      // the render call overwrites the whole frame

    SlrGpuRenderParams p;
    std::memset(&p, 0, sizeof(p));
    p.struct_size = sizeof(p);
    p.width = W; p.height = H;
    p.time_start = settings.getFloat(RenderSettingItem::TimeStart);
    p.time_end = settings.getFloat(RenderSettingItem::TimeEnd);
    p.rng_seed = settings.getInt(RenderSettingItem::RNGSeed);
    if (bidirectional) p.flags |= SLRGPU_RENDER_BPT;
    const float brightness = settings.getFloat(RenderSettingItem::Brightness);
    // one frame segment [begin, end) of the sample range on all devices, into `dst` (overwritten)
    auto renderSegment = [&](uint32_t begin, uint32_t end, float* dst) {
        p.spp_begin = sampleBegin + begin; p.spp_end = sampleBegin + end;
        SlrGpuRenderStats st;
        const int rc = replicas.size() > 1 ? slrgpu_render_multi(replicas.data(), (uint32_t)replicas.size(), &p, dst, &st)
                                           : slrgpu_render(replicas[0], &p, dst, &st);
        if (rc != SLRGPU_OK) throw std::runtime_error(std::string("slrgpu_render failed: ") + slrgpu_last_error());
        lastStatistics.paths += st.paths; lastStatistics.rays += st.rays;
        lastStatistics.deviceSeconds += st.device_ms * 1e-3;
    };

    if (!exportProgressiveImages) {
        // one call for the whole sample range, straight into the sensor
        renderSegment(0, m_samplesPerPixel, sensor->data());
        lastStatistics.wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        return;
    }
    std::vector<float> pass((size_t)W * H * channels);
    // Progressive export cadence of the reference: an image after 1, 2, 4, ... samples (at most 16
    // images, PathTracingRenderer.cpp:63-65,83-94). Each segment [begin, end) is one GPU render call.
    uint32_t begin = 0, exportAt = 1, imgIdx = 0;
    while (begin < m_samplesPerPixel) {
        const uint32_t end = std::min(exportAt, m_samplesPerPixel);
        renderSegment(begin, end, pass.data());
        float* dst = sensor->data();
        for (size_t i = 0; i < pass.size(); ++i) dst[i] += pass[i];
        if (end == exportAt) {
            char name[64];
            std::snprintf(name, sizeof(name), "%03u.bmp", imgIdx);
            sensor->saveImage(outputDirectory + "/" + name, brightness / end);
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
            std::printf("%u samples: %s, %g[s]\n", end, name, el);
            if (++imgIdx == 16) { begin = end; break; }
            exportAt += exportAt;
        }
        begin = end;
    }
    lastStatistics.wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
}

void GPUDebugRenderer::render(const RenderScene& scene, const RenderSettings& settings) const {
    auto wall0 = std::chrono::steady_clock::now();
    const uint32_t W = (uint32_t)settings.getInt(RenderSettingItem::ImageWidth);
    const uint32_t H = (uint32_t)settings.getInt(RenderSettingItem::ImageHeight);
    lastStatistics = RenderStatistics();
    SlrGpuSceneDesc desc;
    scene.flat.describe(&desc);
    SlrGpuScene* gpu = nullptr;
    if (slrgpu_scene_create(&desc, device, &gpu) != SLRGPU_OK)
        throw std::runtime_error(std::string("slrgpu_scene_create failed: ") + slrgpu_last_error());
    SlrGpuRenderParams p;
    std::memset(&p, 0, sizeof(p));
    p.struct_size = sizeof(p);
    p.width = W; p.height = H;
    p.spp_begin = 0; p.spp_end = 1;
    p.time_start = settings.getFloat(RenderSettingItem::TimeStart);
    p.time_end = settings.getFloat(RenderSettingItem::TimeEnd);
    p.rng_seed = settings.getInt(RenderSettingItem::RNGSeed);
    std::vector<float> own;
    float* raw = rawOutput;
    if (!raw) { own.resize((size_t)W * H * SLRGPU_DEBUG_FLOATS); raw = own.data(); }
    SlrGpuRenderStats st;
    const int rc = slrgpu_render_debug(gpu, &p, raw, &st);
    std::string msg = rc == SLRGPU_OK ? std::string() : std::string("slrgpu_render_debug failed: ") + slrgpu_last_error();
    slrgpu_scene_destroy(gpu);
    if (rc != SLRGPU_OK) throw std::runtime_error(msg);
    lastStatistics.paths = st.paths; lastStatistics.rays = st.rays;
    lastStatistics.deviceSeconds = st.device_ms * 1e-3;

    // DebugRenderer.cpp:156-183 + Image2D::saveImage (Image.cpp:310-341): (uint8)clamp((0.5 v + 0.5) * 255, 0, 255), bottom-up BGR rows
    static const char* names[NumChannels] = {"geometric_normal.bmp", "shading_normal.bmp", "shading_tangent.bmp", "distance.bmp"};
    const uint32_t rowBytes = 3 * W + W % 4;
    for (int ch = 0; ch < 3; ++ch) {
        if (!channels[ch]) continue;
        std::vector<uint8_t> bmp((size_t)rowBytes * H, 0);
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                const float* v = raw + ((size_t)y * W + x) * SLRGPU_DEBUG_FLOATS + 1 + 3 * ch;
                uint8_t* dst = &bmp[(size_t)(H - y - 1) * rowBytes + 3 * x];
                for (int c = 0; c < 3; ++c) {
                    const float q = std::min(std::max((0.5f * v[c] + 0.5f) * 255, 0.0f), 255.0f);
                    dst[2 - c] = (uint8_t)q;
                }
            }
        saveBMP(outputDirectory + "/" + names[ch], bmp.data(), W, H);
    }
    lastStatistics.wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
}

}  // namespace slr
