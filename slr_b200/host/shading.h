// Host-side shading model: spectra, textures, surface materials, and their export into the tagged
// GPU tables of include/slrgpu.h. (Declarations; see shading.cpp.)
#pragma once
#include "scene.h"

namespace slr {
void exportEnvironment(GpuSceneBuilder& b, const InfiniteSphereNode& env);
void finishShadingTables(GpuSceneBuilder& b);
}
