// Reference renderer driver (test infrastructure + CPU baseline): the reference's own HostProgram
// flow (HostProgram/main.cpp:20-62) -- initSpectrum, readScene, Scene::build, Renderer::render --
// with the renderer/size/seed overridable from the command line, the sensor dumped as raw floats and
// a timing JSON on stderr.
//   ref_render scene.txt out.bin [spp] [width] [height] [seed] [qbvh=0|1]
// out.bin: u32 width, height, channels (16; 3 in the RGB-mode build ref_render_rgb), then width*height*channels f32 (row-major, un-normalised sums,
// i.e. ImageSensor::pixel(x, y) after PathTracingRenderer::render). spp/width/height/seed <= 0 keep the
// scene file's values. The renderer is the unidirectional PathTracingRenderer unless the 8th argument says "bpt" (the
// reference's BidirectionalPathTracingRenderer; its light-tracing splats live in per-thread "separated" buffers, which the
// dump adds to the pixel exactly like ImageSensor::saveImage does, ImageSensor.cpp:153-156). With qbvh=1 every
// aggregate's accelerator is swapped for QBVH(SBVH) after construction.
#include <libSLR/Core/SurfaceObject.h>
#include <libSLR/Core/RenderSettings.h>
#include <libSLR/Core/ImageSensor.h>
#include <libSLR/Core/cameras.h>
#include <libSLR/Accelerator/SBVH.h>
#include <libSLR/Accelerator/QBVH.h>
#include <libSLR/Memory/ArenaAllocator.h>
#include <cstring>
#include <libSLR/Renderers/PathTracingRenderer.h>
#include <libSLR/Renderers/BidirectionalPathTracingRenderer.h>
#include <libSLR/Renderers/DebugRenderer.h>
#include <libSLR/Core/distributions.h>
#include <libSLR/Cameras/PerspectiveCamera.h>
#include <libSLR/Core/Transform.h>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include <libSLRSceneGraph/Scene.h>
#include <libSLRSceneGraph/API.hpp>
#include <chrono>
#include <set>
#include <thread>
#include <unistd.h>

using namespace SLR;

static void swapToQBVH(const SurfaceObjectAggregate* aggr, std::set<const SurfaceObjectAggregate*>& seen) {
    if (!aggr || seen.count(aggr)) return;
    seen.insert(aggr);
    SBVH* sbvh = dynamic_cast<SBVH*>(aggr->m_accelerator);
    if (!sbvh) return;
    for (const SurfaceObject* o : sbvh->m_objLists)
        if (const TransformedSurfaceObject* t = dynamic_cast<const TransformedSurfaceObject*>(o))
            swapToQBVH(dynamic_cast<const SurfaceObjectAggregate*>(t->m_surfObj), seen);
    const_cast<SurfaceObjectAggregate*>(aggr)->m_accelerator = new QBVH(*sbvh);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_render scene.txt out.bin [spp] [width] [height] [seed] [qbvh] [debug]\n"); return 2; }
    const int spp = argc > 3 ? atoi(argv[3]) : 0, w = argc > 4 ? atoi(argv[4]) : 0, h = argc > 5 ? atoi(argv[5]) : 0;
    const int seed = argc > 6 ? atoi(argv[6]) : 0, qbvh = argc > 7 ? atoi(argv[7]) : 0;
    const bool debugAOV = argc > 8 && std::string(argv[8]) == "debug";     // the reference's DebugRenderer instead of the path tracer
    initSpectrum();
    auto t0 = std::chrono::steady_clock::now();
    SLRSceneGraph::SceneRef scene = createShared<SLRSceneGraph::Scene>();
    SLRSceneGraph::RenderingContext context;
    context.width = 1024; context.height = 1024; context.timeStart = 0; context.timeEnd = 0; context.brightness = 1.0f; context.rngSeed = 1509761209;
    if (!SLRSceneGraph::readScene(argv[1], scene, &context)) { fprintf(stderr, "Failed to read a scene file.\n"); return 1; }
    auto t1 = std::chrono::steady_clock::now();
    const Scene* rawScene;
    ArenaAllocator mem;
    scene->build(&rawScene, mem);
    if (qbvh) { std::set<const SurfaceObjectAggregate*> seen; swapToQBVH(rawScene->m_aggregate, seen); }
    auto t2 = std::chrono::steady_clock::now();

    uint32_t useSpp = spp > 0 ? (uint32_t)spp : 8;
    if (spp <= 0) if (PathTracingRenderer* pt = dynamic_cast<PathTracingRenderer*>(context.renderer.get())) useSpp = pt->m_samplesPerPixel;
    RenderSettings settings;
    settings.addItem(RenderSettingItem::ImageWidth, (int32_t)(w > 0 ? w : context.width));
    settings.addItem(RenderSettingItem::ImageHeight, (int32_t)(h > 0 ? h : context.height));
    settings.addItem(RenderSettingItem::TimeStart, context.timeStart);
    settings.addItem(RenderSettingItem::TimeEnd, context.timeEnd);
    settings.addItem(RenderSettingItem::Brightness, context.brightness);
    settings.addItem(RenderSettingItem::RNGSeed, (int32_t)(seed != 0 ? seed : context.rngSeed));

    // the renderer writes NNN.bmp into the working directory: run inside the output file's directory
    std::string out = argv[2];
    if (out[0] != '/') { char cwd[4096]; out = std::string(getcwd(cwd, sizeof(cwd))) + "/" + out; }
    std::string outDir = out.substr(0, out.find_last_of('/'));
    if (chdir(outDir.c_str()) != 0) perror("chdir");

    if (argc > 8 && std::string(argv[8]) == "lights") {
        // the light-selection distribution the path tracer samples (SurfaceObject.cpp:232-243, 432-466, distributions.cpp:97-119),
        // as raw float bits: world centre / radius, the top-level aggregate's PMF, CDF and integral
        const SurfaceObjectAggregate* ag = rawScene->m_aggregate;
        const RegularConstantDiscrete1D* d = ag->m_lightDist1D;
        auto bits = [](float v) { uint32_t b; memcpy(&b, &v, 4); return b; };
        printf("world %08x %08x %08x %08x\n", bits(rawScene->m_worldCenter.x), bits(rawScene->m_worldCenter.y), bits(rawScene->m_worldCenter.z), bits(rawScene->m_worldRadius));
        if (const PerspectiveCamera* cam = dynamic_cast<const PerspectiveCamera*>(rawScene->getCamera())) {
            // the camera the ray generation uses: local-to-world matrix sampled at time 0 and its inverse, lens / frustum parameters
            StaticTransform tf;
            cam->m_transform->sample(0.0f, &tf);
            printf("camera");
            for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) printf(" %08x", bits(tf.mat[c][r]));
            for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) printf(" %08x", bits(tf.matInv[c][r]));
            printf(" %08x %08x %08x %08x %08x %08x\n", bits(cam->m_aspect), bits(cam->m_fovY), bits(cam->m_lensRadius), bits(cam->m_imgPlaneDistance),
                   bits(cam->m_objPlaneDistance), bits(cam->getSensor()->m_sensitivity));
        }
        printf("lights %u integral %08x\n", d->m_numValues, bits(d->m_integral));
        for (uint32_t i = 0; i < d->m_numValues; ++i) printf("pmf %u %08x cdf %08x %08x\n", i, bits(d->m_PMF[i]), bits(d->m_CDF[i]), bits(d->m_CDF[i + 1]));
        // the environment's importance map (InfiniteSphereSurfaceObject::m_dist, built by IBLEmission::createIBLImportanceMap):
        // size, integrals, an FNV-1a hash over all pdf / cdf bits and a few entries in the clear
        if (const InfiniteSphereSurfaceObject* env = rawScene->m_envSphere) {
            const RegularConstantContinuous2D* m = env->m_dist;
            const uint32_t H = m->m_num1DDists, W = m->m_1DDists[0].m_numValues;
            uint64_t h = 1469598103934665603ull;
            auto mix = [&h, &bits](float v) { uint32_t b = bits(v); for (int k = 0; k < 4; ++k) { h ^= (b >> (8 * k)) & 0xFFu; h *= 1099511628211ull; } };
            for (uint32_t y = 0; y < H; ++y) {
                const RegularConstantContinuous1D& r = m->m_1DDists[y];
                for (uint32_t x = 0; x < W; ++x) mix(r.m_PDF[x]);
                for (uint32_t x = 0; x <= W; ++x) mix(r.m_CDF[x]);
                mix(r.m_integral);
            }
            for (uint32_t y = 0; y < H; ++y) mix(m->m_top1DDist->m_PDF[y]);
            for (uint32_t y = 0; y <= H; ++y) mix(m->m_top1DDist->m_CDF[y]);
            printf("env %u %u integral %08x top %08x hash %016llx\n", W, H, bits(m->m_integral), bits(m->m_top1DDist->m_integral), (unsigned long long)h);
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t y = (k * 2654435761u) % H, x = (k * 40503u + 17u) % W;
                printf("envpdf %u %u %08x %08x %08x\n", x, y, bits(m->m_1DDists[y].m_PDF[x]), bits(m->m_1DDists[y].m_CDF[x]), bits(m->m_top1DDist->m_PDF[y]));
            }
        }
        return 0;
    }
    if (debugAOV) {
        // DebugRenderer writes geometric_normal.bmp / shading_normal.bmp / shading_tangent.bmp into the working directory
        bool flags[(int)ExtraChannel::NumChannels];
        for (int i = 0; i < (int)ExtraChannel::NumChannels; ++i) flags[i] = false;
        flags[(int)ExtraChannel::GeometricNormal] = flags[(int)ExtraChannel::ShadingNormal] = flags[(int)ExtraChannel::ShadingTangent] = true;
        DebugRenderer debugRenderer(flags);
        debugRenderer.render(*rawScene, settings);
        return 0;
    }
    const bool bpt = argc > 8 && std::string(argv[8]) == "bpt";
    if (bpt && spp <= 0) if (BidirectionalPathTracingRenderer* b = dynamic_cast<BidirectionalPathTracingRenderer*>(context.renderer.get())) useSpp = b->m_samplesPerPixel;
    auto t3 = std::chrono::steady_clock::now();
    if (bpt) { BidirectionalPathTracingRenderer renderer(useSpp); renderer.render(*rawScene, settings); }
    else { PathTracingRenderer renderer(useSpp); renderer.render(*rawScene, settings); }
    auto t4 = std::chrono::steady_clock::now();

    ImageSensor* sensor = rawScene->getCamera()->getSensor();
#ifdef Use_Spectral_Representation
    uint32_t W = sensor->width(), H = sensor->height(), C = 16;
#else
    uint32_t W = sensor->width(), H = sensor->height(), C = 3;      // the RGB-mode build (oracle/Makefile, -DSLR_ORACLE_RGB)
#endif
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror(out.c_str()); return 1; }
    fwrite(&W, 4, 1, f); fwrite(&H, 4, 1, f); fwrite(&C, 4, 1, f);
    for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            DiscretizedSpectrum px = ((const ImageSensor*)sensor)->pixel(x, y);
            for (uint32_t b = 0; b < sensor->m_numSeparated; ++b) px += ((const ImageSensor*)sensor)->pixel(b, x, y);
#ifdef Use_Spectral_Representation
            fwrite(px.values, 4, 16, f);
#else
            const float rgb[3] = {px.r, px.g, px.b};
            fwrite(rgb, 4, 3, f);
#endif
        }
    fclose(f);
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    double paths = (double)W * H * useSpp;
    fprintf(stderr, "{\"width\": %u, \"height\": %u, \"spp\": %u, \"threads\": %u, \"read_s\": %.3f, \"build_s\": %.3f, \"render_s\": %.4f, "
                    "\"mpaths_per_s\": %.5f, \"accelerator\": \"%s\", \"sensitivity\": %.9g, \"renderer\": \"%s\"}\n",
            W, H, useSpp, std::thread::hardware_concurrency(), sec(t0, t1), sec(t1, t2), sec(t3, t4), paths / sec(t3, t4) / 1e6,
            qbvh ? "QBVH" : "SBVH", (double)sensor->m_sensitivity, bpt ? "BPT" : "PT");
    return 0;
}
