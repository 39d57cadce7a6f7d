#include "shading.h"
#include <stdexcept>

namespace slr {

// Phase-1 placeholders: geometry-only scenes carry no materials yet.
uint32_t GpuSceneBuilder::exportMaterial(const SurfaceMaterial*) { throw std::runtime_error("materials not implemented yet"); }
uint32_t GpuSceneBuilder::exportNormalTexture(const Normal3DTexture*) { throw std::runtime_error("textures not implemented yet"); }
uint32_t GpuSceneBuilder::exportFloatTexture(const FloatTexture*) { throw std::runtime_error("textures not implemented yet"); }
bool GpuSceneBuilder::materialEmits(const SurfaceMaterial*) const { return false; }
void exportEnvironment(GpuSceneBuilder&, const InfiniteSphereNode&) { throw std::runtime_error("environment not implemented yet"); }
void finishShadingTables(GpuSceneBuilder&) {}

}  // namespace slr
