// The reference's host program with the GPU renderer dropped in: HostProgram/main.cpp:20-62 line for line --
// initSpectrum, readScene (the reference's own flex/bison interpreter), Scene::build (the reference's own SBVH builder),
// RenderSettings from the RenderingContext -- except that the Renderer the scene file selected ("PT" / "BPT", created at
// libSLRSceneGraph/API.cpp:1016-1036) is replaced by SLR::GPUPathTracingRenderer with the same sample count. In a merge
// this is a two-line change at API.cpp:1025 (`new GPUPathTracingRenderer(samples)`); it is done here so that the
// reference's sources stay untouched.
//   slr_gpu scene.txt [sensor_out.bin] [bpt]
// Without the third argument every scene renders with the unidirectional GPU path tracer (the configuration the headline
// benchmark and the PT parity tests use); with "bpt" a scene that selected "BPT" gets SLR::GPUBidirectionalPathTracingRenderer
// -- in a merge the second line of the change, at API.cpp:1033.
// With a second argument the camera's ImageSensor is dumped after rendering (u32 width, height, 16, then
// width*height*16 f32 = ImageSensor::pixel(x, y)), the same format oracle/drivers/ref_render.cpp writes: the parity test
// compares the two.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

#include <libSLR/defines.h>
#include <libSLRSceneGraph/references.h>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include <libSLR/Core/ImageSensor.h>
#include <libSLR/Core/RenderSettings.h>
#include <libSLR/Core/Renderer.h>
#include <libSLR/Core/SurfaceObject.h>
#include <libSLR/Core/cameras.h>
#include <libSLR/Memory/ArenaAllocator.h>
#include <libSLR/Renderers/BidirectionalPathTracingRenderer.h>
#include <libSLR/Renderers/PathTracingRenderer.h>
#include <libSLRSceneGraph/API.hpp>
#include <libSLRSceneGraph/Scene.h>

#include "GPUPathTracingRenderer.h"

int main(int argc, const char* argv[]) {
    if (argc < 2) {
        fprintf(stderr, "Too few command line arguments.\n");
        return -1;
    }
    SLR::initSpectrum();

    SLRSceneGraph::SceneRef scene = createShared<SLRSceneGraph::Scene>();
    SLRSceneGraph::RenderingContext context;
    // the defaults API.cpp:1075-1081 would leave when the file does not call setRenderSettings
    context.width = 1024; context.height = 1024; context.timeStart = 0; context.timeEnd = 0; context.brightness = 1.0f; context.rngSeed = 1509761209;
    if (!SLRSceneGraph::readScene(argv[1], scene, &context)) {
        printf("Failed to read a scene file.\n");
        exit(-1);
    }
    const SLR::Scene* rawScene;
    SLR::ArenaAllocator mem;
    scene->build(&rawScene, mem);

    SLR::RenderSettings settings;
    settings.addItem(SLR::RenderSettingItem::ImageWidth, context.width);
    settings.addItem(SLR::RenderSettingItem::ImageHeight, context.height);
    settings.addItem(SLR::RenderSettingItem::TimeStart, context.timeStart);
    settings.addItem(SLR::RenderSettingItem::TimeEnd, context.timeEnd);
    settings.addItem(SLR::RenderSettingItem::Brightness, context.brightness);
    settings.addItem(SLR::RenderSettingItem::RNGSeed, context.rngSeed);

    // ---- the drop-in: same seam (Renderer::render), GPU implementation
    uint32_t spp = 8;
    bool bidirectional = false;
    if (auto pt = dynamic_cast<SLR::PathTracingRenderer*>(context.renderer.get())) spp = pt->m_samplesPerPixel;
    else if (auto bpt = dynamic_cast<SLR::BidirectionalPathTracingRenderer*>(context.renderer.get())) {
        spp = bpt->m_samplesPerPixel;
        bidirectional = argc > 3 && std::string(argv[3]) == "bpt";
    }
    if (bidirectional) context.renderer.reset(new SLR::GPUBidirectionalPathTracingRenderer(spp));
    else context.renderer.reset(new SLR::GPUPathTracingRenderer(spp));

    // the renderer writes NNN.bmp into the working directory, like the reference's
    std::string out = argc > 2 ? argv[2] : "";
    if (!out.empty() && out[0] != '/') { char cwd[4096]; out = std::string(getcwd(cwd, sizeof(cwd))) + "/" + out; }
    if (!out.empty() && chdir(out.substr(0, out.find_last_of('/')).c_str()) != 0) perror("chdir");

    try {
        context.renderer->render(*rawScene, settings);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }

    if (!out.empty()) {
        const SLR::ImageSensor* sensor = rawScene->getCamera()->getSensor();
        uint32_t W = sensor->width(), H = sensor->height(), C = 16;
        FILE* f = fopen(out.c_str(), "wb");
        if (!f) { perror(out.c_str()); return 1; }
        fwrite(&W, 4, 1, f); fwrite(&H, 4, 1, f); fwrite(&C, 4, 1, f);
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                SLR::DiscretizedSpectrum px = sensor->pixel(x, y);
                fwrite(px.values, 4, 16, f);
            }
        fclose(f);
        fprintf(stderr, "{\"width\": %u, \"height\": %u, \"spp\": %u, \"sensitivity\": %.9g}\n", W, H, spp, (double)sensor->m_sensitivity);
    }
    return 0;
}
