// Scene::build: flatten -> aggregates -> SoA, plus camera / environment export.
#include "scene.h"
#include "shading.h"
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace slr {

void Scene::build(FlatScene* out, bool rgbMode) {
    auto t0 = std::chrono::steady_clock::now();
    *out = FlatScene();
    out->rgbMode = rgbMode;
    GpuSceneBuilder b(*out);
    RenderingData data;
    m_root->getRenderingData(b, nullptr, &data);
    if (data.objects.empty()) throw std::runtime_error("Scene::build: the scene has no surfaces");
    const uint32_t top = b.createAggregate(std::move(data.objects));
    b.finalize(top);

    if (data.camera) {
        const PerspectiveCamera& c = *data.camera;
        SlrGpuCamera& g = out->camera;
        std::memcpy(g.mat, &data.cameraTransform.mat, sizeof(float) * 16);
        std::memcpy(g.mat_inv, &data.cameraTransform.matInv, sizeof(float) * 16);
        g.sensitivity = c.sensitivity; g.aspect = c.aspect; g.fov_y = c.fovY;
        g.lens_radius = c.lensRadius; g.img_plane_dist = c.imgPlaneDistance; g.obj_plane_dist = c.objPlaneDistance;
        if (data.cameraTransform.anim) out->cameraMotion = b.addMotion(*data.cameraTransform.anim);       // the camera moves
        out->hasCamera = true;
    }
    if (m_env) exportEnvironment(b, *m_env);
    finishShadingTables(b);
    out->buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace slr
