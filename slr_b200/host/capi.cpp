// C entry points of the host library (include/slrhost.h).
#include "../../include/slrhost.h"
#include "scene.h"
#include "renderer.h"
#include "assets/png.h"
#include "assets/assbin.h"
#include "assets/exr.h"
#include "parser/scene_parser.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>

using namespace slr;

static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}

struct SlrHostBuilder {
    Scene scene;
    std::vector<TriangleMeshNodeRef> meshes;
    std::vector<bool> placed;
    std::vector<NodeRef> references;     // one shared ReferenceNode per instanced mesh
};
struct SlrHostScene {
    RenderScene render;
    FlatScene& flat = render.flat;
    RenderingContext context;
    bool hasRenderer = false;
};

static StaticTransform toTransform(const float* m) {
    if (!m) return StaticTransform();
    Mat4 mm;
    std::memcpy(static_cast<void*>(&mm), m, sizeof(float) * 16);
    return StaticTransform(mm);
}

extern "C" {

SLRGPU_API const char* slrhost_last_error(void) { return g_err; }

SLRGPU_API int slrhost_set_option(const char* name, int value) {
    if (!name) return fail("slrhost_set_option: null name");
    if (!std::strcmp(name, "export_sbvh")) { FlatScene::exportSbvh = value != 0; return 0; }
    return fail("slrhost_set_option: unknown option %s", name);
}

SLRGPU_API SlrHostBuilder* slrhost_builder_create(void) {
    try { return new SlrHostBuilder(); } catch (...) { fail("allocation failed"); return nullptr; }
}
SLRGPU_API void slrhost_builder_destroy(SlrHostBuilder* b) { delete b; }

SLRGPU_API int slrhost_builder_add_mesh(SlrHostBuilder* b, const float* positions, const float* normals,
                                        const float* tangents, const float* uvs, uint32_t nv,
                                        const uint32_t* indices, uint32_t nt) {
    if (!b || !positions || !indices || nv == 0 || nt == 0) return fail("slrhost_builder_add_mesh: invalid argument");
    try {
        TriangleMeshNodeRef mesh = std::make_shared<TriangleMeshNode>();
        for (uint32_t i = 0; i < nv; ++i) {
            Vertex v;
            v.position = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
            v.normal = normals ? Vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]) : Vec3(0, 1, 0);
            v.tangent = tangents ? Vec3(tangents[3 * i], tangents[3 * i + 1], tangents[3 * i + 2]) : Vec3(1, 0, 0);
            v.texCoord = uvs ? Vec2(uvs[2 * i], uvs[2 * i + 1]) : Vec2(0, 0);
            mesh->addVertex(v);
        }
        std::vector<uint32_t> idx(indices, indices + (size_t)nt * 3);
        mesh->addTriangles(nullptr, nullptr, nullptr, std::move(idx));
        b->meshes.push_back(mesh);
        b->placed.push_back(false);
        b->references.push_back(nullptr);
        return (int)b->meshes.size() - 1;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_builder_place_mesh(SlrHostBuilder* b, int mesh, const float* mat) {
    if (!b || mesh < 0 || mesh >= (int)b->meshes.size()) return fail("slrhost_builder_place_mesh: bad mesh id");
    if (b->placed[mesh]) return fail("mesh %d is already part of the scene; instance it instead", mesh);
    InternalNodeRef holder = std::make_shared<InternalNode>();
    holder->setTransform(toTransform(mat));
    holder->addChildNode(b->meshes[mesh]);
    b->scene.rootNode()->addChildNode(holder);
    b->placed[mesh] = true;
    return 0;
}

SLRGPU_API int slrhost_builder_instance_mesh(SlrHostBuilder* b, int mesh, const float* mat) {
    if (!b || mesh < 0 || mesh >= (int)b->meshes.size()) return fail("slrhost_builder_instance_mesh: bad mesh id");
    InternalNodeRef holder = std::make_shared<InternalNode>();
    holder->setTransform(toTransform(mat));
    // one ReferenceNode per mesh, shared by all its placements, so they share one nested aggregate
    if (!b->references[mesh]) b->references[mesh] = std::make_shared<ReferenceNode>(b->meshes[mesh]);
    holder->addChildNode(b->references[mesh]);
    b->scene.rootNode()->addChildNode(holder);
    b->placed[mesh] = true;
    return 0;
}

SLRGPU_API int slrhost_builder_finish(SlrHostBuilder* b, int rgb_mode, SlrHostScene** out) {
    if (!b || !out) return fail("slrhost_builder_finish: null argument");
    *out = nullptr;
    try {
        SlrHostScene* s = new SlrHostScene();
        try { b->scene.build(&s->render.flat, rgb_mode != 0); } catch (...) { delete s; throw; }
        *out = s;
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_read_scene(const char* path, int rgb_mode, SlrHostScene** out) {
    if (!path || !out) return fail("slrhost_read_scene: null argument");
    *out = nullptr;
    try {
        std::unique_ptr<SlrHostScene> s(new SlrHostScene());
        Scene graph;
        std::string err;
        if (!readScene(path, &graph, &s->context, &err, rgb_mode != 0)) return fail("%s", err.c_str());
        s->hasRenderer = s->context.renderer != nullptr;
        graph.build(&s->flat, rgb_mode != 0);
        *out = s.release();
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_scene_context(const SlrHostScene* s, double* c) {
    if (!s || !c) return fail("slrhost_scene_context: null argument");
    c[0] = s->context.width; c[1] = s->context.height; c[2] = s->context.samples; c[3] = s->context.rngSeed;
    c[4] = s->context.timeStart; c[5] = s->context.timeEnd; c[6] = s->context.brightness; c[7] = s->hasRenderer ? 1 : 0;
    return 0;
}

SLRGPU_API int slrhost_render(SlrHostScene* s, int device, int width, int height, int spp, int seed,
                              const char* bmp_dir, float* accum, double* stats) {
    return slrhost_render_range(s, device, width, height, 0, spp, seed, bmp_dir, accum, stats);
}

static int renderRange(SlrHostScene* s, bool bidirectional, int device, int width, int height, int spp_begin, int spp, int seed,
                       const char* bmp_dir, float* accum, double* stats);

SLRGPU_API int slrhost_render_range(SlrHostScene* s, int device, int width, int height, int spp_begin, int spp, int seed,
                                    const char* bmp_dir, float* accum, double* stats) {
    return renderRange(s, false, device, width, height, spp_begin, spp, seed, bmp_dir, accum, stats);
}

SLRGPU_API int slrhost_render_bpt(SlrHostScene* s, int device, int width, int height, int spp_begin, int spp, int seed,
                                  const char* bmp_dir, float* accum, double* stats) {
    return renderRange(s, true, device, width, height, spp_begin, spp, seed, bmp_dir, accum, stats);
}

SLRGPU_API int slrhost_scene_renderer_method(const SlrHostScene* s, char* method, uint32_t capacity) {
    if (!s || !method || capacity == 0) return fail("slrhost_scene_renderer_method: invalid argument");
    std::snprintf(method, capacity, "%s", s->context.rendererMethod.c_str());
    return 0;
}

static int renderRange(SlrHostScene* s, bool bidirectional, int device, int width, int height, int spp_begin, int spp, int seed,
                       const char* bmp_dir, float* accum, double* stats) {
    if (!s) return fail("slrhost_render: null scene");
    if (spp_begin < 0) return fail("slrhost_render_range: negative first sample index");
    try {
        if (!s->flat.hasCamera) return fail("the scene has no camera");
        if (!s->render.sensor) {
            float sens = s->flat.camera.sensitivity;
            float r = s->flat.camera.lens_radius;
            s->render.sensor = std::make_shared<ImageSensor>(sens > 0 ? sens : (float)(1.0f / (M_PI * r * r)));
        }
        RenderSettings settings;
        settings.addItem(RenderSettingItem::ImageWidth, (int32_t)(width > 0 ? width : s->context.width));
        settings.addItem(RenderSettingItem::ImageHeight, (int32_t)(height > 0 ? height : s->context.height));
        settings.addItem(RenderSettingItem::TimeStart, s->context.timeStart);
        settings.addItem(RenderSettingItem::TimeEnd, s->context.timeEnd);
        settings.addItem(RenderSettingItem::Brightness, s->context.brightness);
        settings.addItem(RenderSettingItem::RNGSeed, (int32_t)(seed != 0 ? seed : s->context.rngSeed));
        GPUPathTracingRenderer renderer(spp > 0 ? (uint32_t)spp : s->context.samples);
        // device >= 0: that one device; device < 0: every visible device, the frame's samples partitioned over them
        renderer.device = device >= 0 ? device : 0;
        renderer.deviceCount = device >= 0 ? 1 : 0;
        renderer.sampleBegin = (uint32_t)spp_begin;
        renderer.bidirectional = bidirectional;
        renderer.exportProgressiveImages = bmp_dir != nullptr;
        if (bmp_dir) renderer.outputDirectory = bmp_dir;
        // with a caller buffer the sensor renders straight into it (no frame-sized copies on the way out)
        s->render.sensor->bindExternal(accum);
        try { renderer.render(s->render, settings); } catch (...) { s->render.sensor->bindExternal(nullptr); throw; }
        const ImageSensor& sensor = *s->render.sensor;
        if (stats) {
            const RenderStatistics& st = renderer.lastStatistics;
            stats[0] = (double)st.paths; stats[1] = (double)st.rays; stats[2] = st.deviceSeconds; stats[3] = st.wallSeconds;
            stats[4] = st.uploadSeconds; stats[5] = sensor.channels() + 1000.0 * st.devices;     // channels + 1000 x devices used
        }
        s->render.sensor->bindExternal(nullptr);
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_render_debug(SlrHostScene* s, int device, int width, int height, int seed,
                                    const char* bmp_dir, float* out, double* stats) {
    if (!s) return fail("slrhost_render_debug: null scene");
    try {
        if (!s->flat.hasCamera) return fail("the scene has no camera");
        RenderSettings settings;
        settings.addItem(RenderSettingItem::ImageWidth, (int32_t)(width > 0 ? width : s->context.width));
        settings.addItem(RenderSettingItem::ImageHeight, (int32_t)(height > 0 ? height : s->context.height));
        settings.addItem(RenderSettingItem::TimeStart, s->context.timeStart);
        settings.addItem(RenderSettingItem::TimeEnd, s->context.timeEnd);
        settings.addItem(RenderSettingItem::Brightness, s->context.brightness);
        settings.addItem(RenderSettingItem::RNGSeed, (int32_t)(seed != 0 ? seed : s->context.rngSeed));
        // the channels the scene file asked for, or all three when it selected another renderer
        bool flags[GPUDebugRenderer::NumChannels] = {true, true, true, false};
        if (const GPUDebugRenderer* chosen = dynamic_cast<const GPUDebugRenderer*>(s->context.renderer.get()))
            for (int i = 0; i < GPUDebugRenderer::NumChannels; ++i) flags[i] = chosen->channels[i];
        if (!bmp_dir) for (bool& f : flags) f = false;
        GPUDebugRenderer renderer(flags);
        renderer.device = device;
        renderer.rawOutput = out;
        if (bmp_dir) renderer.outputDirectory = bmp_dir;
        renderer.render(s->render, settings);
        if (stats) {
            const RenderStatistics& st = renderer.lastStatistics;
            stats[0] = (double)st.paths; stats[1] = (double)st.rays; stats[2] = st.deviceSeconds; stats[3] = st.wallSeconds;
            stats[4] = st.uploadSeconds; stats[5] = SLRGPU_DEBUG_FLOATS;
        }
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_save_bmp(const char* path, const float* accum, int width, int height, int channels, float scale, float sensitivity) {
    if (!path || !accum || width <= 0 || height <= 0) return fail("slrhost_save_bmp: invalid argument");
    try {
        ImageSensor sensor(sensitivity);
        sensor.init(width, height, channels);
        std::memcpy(sensor.data(), accum, sizeof(float) * (size_t)width * height * channels);
        sensor.saveImage(path, scale);
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_decode_png(const char* path, int gamma_correction, uint32_t* width, uint32_t* height, uint32_t* channels,
                                  uint8_t* pixels, uint64_t capacity) {
    if (!path || !width || !height || !channels) return fail("slrhost_decode_png: null argument");
    slr::png::Image img;
    std::string err;
    if (!slr::png::load(path, gamma_correction != 0, &img, &err)) return fail("%s", err.c_str());
    *width = img.width; *height = img.height; *channels = img.channels | (img.hasAlpha ? 0x100u : 0u);
    if (pixels) {
        if (capacity < img.pixels.size()) return fail("slrhost_decode_png: output buffer too small");
        std::memcpy(pixels, img.pixels.data(), img.pixels.size());
    }
    return 0;
}

SLRGPU_API int slrhost_accum_to_rgb(const float* accum, int width, int height, int channels, float scale, float* rgb) {
    if (!accum || !rgb || width <= 0 || height <= 0) return fail("slrhost_accum_to_rgb: invalid argument");
    try {
        ImageSensor sensor(1.0f);
        sensor.init(width, height, channels);
        std::memcpy(sensor.data(), accum, sizeof(float) * (size_t)width * height * channels);
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) sensor.pixelRGB(x, y, scale, rgb + 3 * ((size_t)y * width + x));
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_write_assbin(const char* path, const float* positions, const float* normals, const float* tangents,
                                    const float* uvs, uint32_t nv, const uint32_t* indices, uint32_t nt,
                                    const char* material_name, const float* diffuse_rgb) {
    if (!path || !positions || !indices || nv == 0 || nt == 0) return fail("slrhost_write_assbin: invalid argument");
    assbin::Scene sc;
    assbin::Mesh m;
    m.name = "mesh";
    m.positions.assign(positions, positions + 3ull * nv);
    if (normals) m.normals.assign(normals, normals + 3ull * nv);
    if (tangents) {
        m.tangents.assign(tangents, tangents + 3ull * nv);
        m.bitangents.resize(3ull * nv, 0.0f);
        if (normals)
            for (uint32_t i = 0; i < nv; ++i) {
                Vec3 b = cross(Vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]), Vec3(tangents[3 * i], tangents[3 * i + 1], tangents[3 * i + 2]));
                m.bitangents[3 * i] = b.x; m.bitangents[3 * i + 1] = b.y; m.bitangents[3 * i + 2] = b.z;
            }
    }
    if (uvs) {
        m.numUVComponents = 2;
        m.texCoords.resize(3ull * nv, 0.0f);
        for (uint32_t i = 0; i < nv; ++i) { m.texCoords[3 * i] = uvs[2 * i]; m.texCoords[3 * i + 1] = uvs[2 * i + 1]; }
    }
    m.indices.assign(indices, indices + 3ull * nt);
    sc.meshes.push_back(m);
    assbin::Material mat;
    mat.setString("?mat.name", material_name ? material_name : "material");
    if (diffuse_rgb) mat.setColor("$clr.diffuse", diffuse_rgb[0], diffuse_rgb[1], diffuse_rgb[2]);
    sc.materials.push_back(mat);
    sc.root.name = "root";
    sc.root.meshes.push_back(0);
    return assbin::save(path, sc) ? 0 : fail("cannot write %s", path);
}

SLRGPU_API int slrhost_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgba) {
    if (!path || !rgba || !width || !height) return fail("slrhost_write_exr: invalid argument");
    return exr::save(path, width, height, rgba) ? 0 : fail("cannot write %s", path);
}

SLRGPU_API int slrhost_sample_animated(const float* mat_begin, const float* mat_end, float t_begin, float t_end, const float* box6,
                                       const float* times, uint32_t n, float* decomposition46, float* bounds6, float* out32n) {
    if (!mat_begin || !mat_end) return fail("slrhost_sample_animated: null argument");
    try {
        const AnimatedTransform a(toTransform(mat_begin), toTransform(mat_end), t_begin, t_end);
        if (decomposition46) {
            float* o = decomposition46;
            for (int k = 0; k < 2; ++k) {
                *o++ = a.T[k].x; *o++ = a.T[k].y; *o++ = a.T[k].z;
                *o++ = a.R[k].x; *o++ = a.R[k].y; *o++ = a.R[k].z; *o++ = a.R[k].w;
                std::memcpy(o, &a.S[k], 64); o += 16;
            }
        }
        if (bounds6 && box6) {
            const BBox b = a.motionBounds(BBox(Vec3(box6[0], box6[1], box6[2]), Vec3(box6[3], box6[4], box6[5])));
            bounds6[0] = b.lo.x; bounds6[1] = b.lo.y; bounds6[2] = b.lo.z; bounds6[3] = b.hi.x; bounds6[4] = b.hi.y; bounds6[5] = b.hi.z;
        }
        for (uint32_t i = 0; i < n && times && out32n; ++i) {
            const StaticTransform tf = a.sample(times[i]);
            std::memcpy(out32n + 32 * (size_t)i, &tf.mat, 64);
            std::memcpy(out32n + 32 * (size_t)i + 16, &tf.matInv, 64);
        }
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
}

SLRGPU_API int slrhost_read_exr(const char* path, uint32_t* width, uint32_t* height, float* rgba, uint64_t capacity_floats) {
    if (!path || !width || !height) return fail("slrhost_read_exr: null argument");
    exr::Image img;
    std::string err;
    if (!exr::load(path, &img, &err)) return fail("%s", err.c_str());
    *width = img.width; *height = img.height;
    if (rgba) {
        if (capacity_floats < img.rgba.size()) return fail("slrhost_read_exr: output buffer too small");
        for (size_t i = 0; i < img.rgba.size(); ++i) rgba[i] = exr::halfToFloat(img.rgba[i]);
    }
    return 0;
}

SLRGPU_API void slrhost_scene_destroy(SlrHostScene* s) { delete s; }

SLRGPU_API int slrhost_scene_describe(const SlrHostScene* s, SlrGpuSceneDesc* desc) {
    if (!s || !desc) return fail("slrhost_scene_describe: null argument");
    s->flat.describe(desc);
    return 0;
}

SLRGPU_API int slrhost_scene_stats(const SlrHostScene* s, uint32_t a, double* st) {
    if (!s) return fail("null scene");
    if (st && a < s->flat.stats.size()) {
        const FlatScene::AggregateStats& x = s->flat.stats[a];
        st[0] = x.numObjects; st[1] = x.sbvhNodes; st[2] = x.sbvhRefs; st[3] = x.sbvhDepth; st[4] = x.qbvhNodes;
        st[5] = x.qbvhDepth; st[6] = x.nodeBase; st[7] = x.leafBase; st[8] = x.sbvhCost; st[9] = x.qbvhCost;
    }
    return (int)s->flat.stats.size();
}

SLRGPU_API double slrhost_scene_build_seconds(const SlrHostScene* s) { return s ? s->flat.buildSeconds : 0.0; }

}  // extern "C"
