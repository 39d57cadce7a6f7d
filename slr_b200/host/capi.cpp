// C entry points of the host library (include/slrhost.h).
#include "../../include/slrhost.h"
#include "scene.h"
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
    FlatScene flat;
};

static StaticTransform toTransform(const float* m) {
    if (!m) return StaticTransform();
    Mat4 mm;
    std::memcpy(static_cast<void*>(&mm), m, sizeof(float) * 16);
    return StaticTransform(mm);
}

extern "C" {

SLRGPU_API const char* slrhost_last_error(void) { return g_err; }

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
        try { b->scene.build(&s->flat, rgb_mode != 0); } catch (...) { delete s; throw; }
        *out = s;
        return 0;
    } catch (const std::exception& e) { return fail("%s", e.what()); }
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
