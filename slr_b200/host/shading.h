// Host-side shading model: textures, texture mappings, surface materials, emitter properties and
// images, plus their export into the tagged GPU tables of include/slrgpu.h.
//
// The reference has one C++ class per material / texture, instantiated by libSLRSceneGraph's
// factories (surface_materials.cpp:123-171, textures.cpp) and evaluated through virtual calls per
// hit. Here every object is a small tagged descriptor: the per-hit evaluation is device code
// (slr_b200/csrc/shade*.cuh), the host only records parameters. The factory names follow the
// reference so the scene-language builtins read the same.
#pragma once
#include "scene.h"
#include "spectrum.h"

namespace slr {

struct TextureMapping {
    SlrGpuMappingKind kind = SLRGPU_MAP_TEXCOORD;
    float offset[2] = {0, 0}, scale[2] = {1, 1};
};
typedef std::shared_ptr<TextureMapping> TextureMappingRef;

// Linear (row-major) image in one of the reference's texel formats (libSLR/Core/Image.h:18-45).
class Image2D {
public:
    uint32_t width = 0, height = 0;
    SlrGpuImageFormat format = SLRGPU_IMG_RGBA8x4;
    SpectrumType spectrumType = SpectrumType::Reflectance;
    std::vector<uint8_t> data;
    static size_t texelSize(SlrGpuImageFormat f);
};
typedef std::shared_ptr<Image2D> Image2DRef;

class SpectrumTexture {
public:
    SlrGpuTextureKind kind = SLRGPU_TEX_CONSTANT_SPECTRUM;
    InputSpectrumRef spectrum[2];
    Image2DRef image;
    TextureMappingRef mapping;
    float f0 = 0, f1 = 0;
    static SpectrumTextureRef constant(const InputSpectrumRef& s);
    static SpectrumTextureRef checkerBoard(const TextureMappingRef& m, const InputSpectrumRef& v0, const InputSpectrumRef& v1);
    static SpectrumTextureRef voronoi(const TextureMappingRef& m, float scale, float brightness);
    static SpectrumTextureRef imageTexture(const TextureMappingRef& m, const Image2DRef& img);
};

class Normal3DTexture {
public:
    SlrGpuTextureKind kind = SLRGPU_TEX_CHECKER_NORMAL;
    Image2DRef image;
    TextureMappingRef mapping;
    float f0 = 0, f1 = 0;
    uint32_t i0 = 0;
    static Normal3DTextureRef checkerBoard(const TextureMappingRef& m, float stepWidth, bool reverse);
    static Normal3DTextureRef voronoi(const TextureMappingRef& m, float scale, float thetaMax);
    static Normal3DTextureRef imageTexture(const TextureMappingRef& m, const Image2DRef& img);
};

class FloatTexture {
public:
    SlrGpuTextureKind kind = SLRGPU_TEX_CONSTANT_FLOAT;
    Image2DRef image;
    TextureMappingRef mapping;
    float f0 = 0, f1 = 0;
    uint32_t i0 = 0;
    static FloatTextureRef constant(float v);
    static FloatTextureRef checkerBoard(const TextureMappingRef& m, float v0, float v1);
    static FloatTextureRef voronoi(const TextureMappingRef& m, float scale, float valueScale, bool flat);
    static FloatTextureRef imageTexture(const TextureMappingRef& m, const Image2DRef& img);
};

class EmitterSurfaceProperty {
public:
    SlrGpuMaterialKind kind = SLRGPU_MAT_DIFFUSE_EMISSION;
    SpectrumTextureRef emittance;
    float scale = 1.0f;
};
typedef std::shared_ptr<EmitterSurfaceProperty> EmitterSurfacePropertyRef;

class IBLEmission : public EmitterSurfaceProperty {
public:
    IBLEmission(const SpectrumTextureRef& coeffM, float s) { kind = SLRGPU_MAT_IBL_EMISSION; emittance = coeffM; scale = s; }
};

class SurfaceMaterial {
public:
    SlrGpuMaterialKind kind = SLRGPU_MAT_DIFFUSE;
    SpectrumTextureRef stex[4];
    FloatTextureRef ftex[4];
    SurfaceMaterialRef sub[2];
    EmitterSurfacePropertyRef emitter;
    bool isEmitting() const { return kind == SLRGPU_MAT_EMITTER; }
    // libSLRSceneGraph/surface_materials.cpp:123-171
    static SurfaceMaterialRef createMatte(const SpectrumTextureRef& reflectance, const FloatTextureRef& sigma);
    static SurfaceMaterialRef createMetal(const SpectrumTextureRef& coeffR, const SpectrumTextureRef& eta, const SpectrumTextureRef& k);
    static SurfaceMaterialRef createGlass(const SpectrumTextureRef& coeff, const SpectrumTextureRef& etaExt, const SpectrumTextureRef& etaInt);
    static SurfaceMaterialRef createModifiedWardDur(const SpectrumTextureRef& R, const FloatTextureRef& anisoX, const FloatTextureRef& anisoY);
    static SurfaceMaterialRef createAshikhminShirley(const SpectrumTextureRef& Rd, const SpectrumTextureRef& Rs, const FloatTextureRef& nu, const FloatTextureRef& nv);
    static SurfaceMaterialRef createMicrofacetMetal(const SpectrumTextureRef& eta, const SpectrumTextureRef& k, const FloatTextureRef& alpha_g);
    static SurfaceMaterialRef createMicrofacetGlass(const SpectrumTextureRef& etaExt, const SpectrumTextureRef& etaInt, const FloatTextureRef& alpha_g);
    static SurfaceMaterialRef createInverseMaterial(const SurfaceMaterialRef& base);
    static SurfaceMaterialRef createSummedMaterial(const SurfaceMaterialRef& m0, const SurfaceMaterialRef& m1);
    static SurfaceMaterialRef createMixedMaterial(const SurfaceMaterialRef& m0, const SurfaceMaterialRef& m1, const FloatTextureRef& factor);
    static EmitterSurfacePropertyRef createDiffuseEmitter(const SpectrumTextureRef& emittance);
    static SurfaceMaterialRef createEmitterSurfaceMaterial(const SurfaceMaterialRef& mat, const EmitterSurfacePropertyRef& emit);
};

void exportEnvironment(GpuSceneBuilder& b, const InfiniteSphereNode& env);
void finishShadingTables(GpuSceneBuilder& b);

}  // namespace slr
