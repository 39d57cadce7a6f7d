// GPUPathTracingRenderer -- the class a libSLR maintainer adds next to PathTracingRenderer
// (libSLR/Renderers/PathTracingRenderer.h) to drop the B200 path in behind the reference's own seam
//     class Renderer { virtual void render(const Scene &scene, const RenderSettings &settings) const = 0; }
// (libSLR/Core/Renderer.h:15-19). It flattens the SLR::Scene it is handed into the SoA buffers of
// include/slrgpu.h, calls the C ABI (slrgpu_scene_create / slrgpu_render[_multi]) and leaves the camera's
// ImageSensor filled, so ImageSensor::saveImage and every caller that reads pixel(x, y) keep working
// (seam 1 and seam 3).
//
// This file and GPUPathTracingRenderer.cpp are compiled INSIDE the reference build by oracle/Makefile (target
// `dropin`, output oracle/_ref/slr_gpu) against the untouched headers under /root/reference -- they are the compiled,
// tested form of the binding INTEGRATION.md describes. They contain no rendering code: everything numerical happens
// behind slrgpu.h.
#pragma once
#include <libSLR/Core/Renderer.h>
#include <cstdint>

namespace SLR {
    class GPUPathTracingRenderer : public Renderer {
        uint32_t m_samplesPerPixel;
    public:
        // device < 0: every visible GPU, the frame's samples partitioned over them (slrgpu_render_multi)
        int device = -1;
        bool exportProgressiveImages = true;      // NNN.bmp after 1, 2, 4, ... samples (PathTracingRenderer.cpp:63-65,83-94)
        bool bidirectional = false;               // SLRGPU_RENDER_BPT: what GPUBidirectionalPathTracingRenderer sets
        explicit GPUPathTracingRenderer(uint32_t spp) : m_samplesPerPixel(spp) { }
        void render(const Scene &scene, const RenderSettings &settings) const override;
    };
    // next to BidirectionalPathTracingRenderer (libSLR/Renderers/BidirectionalPathTracingRenderer.h): same seam, same pass
    // loop and export cadence, bidirectional samples on the GPU (the light-tracing splats land in the sensor's main buffer)
    class GPUBidirectionalPathTracingRenderer : public GPUPathTracingRenderer {
    public:
        explicit GPUBidirectionalPathTracingRenderer(uint32_t spp) : GPUPathTracingRenderer(spp) { bidirectional = true; }
    };
}
