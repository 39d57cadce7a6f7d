// Per-hit material evaluation: walks the tagged material tree of include/slrgpu.h, evaluates the
// textures at the surface point and fills the lobes of bsdf.cuh.
//   SurfaceMaterial::getBSDF family   libSLR/SurfaceMaterials/*.cpp (basic_SurfaceMaterials.cpp:15-50,
//                                     ModifiedWardDurReflection.cpp, AshikhminShirleyReflection.cpp,
//                                     MicrofacetSurfaceMaterial.cpp, Summed/MixedSurfaceMaterial.cpp)
//   EmitterSurfaceMaterial            libSLR/Core/surface_material.h (getBSDF of the scattering part, emittance of the emitter)
//   DiffuseEmission / IBLEmission     libSLR/SurfaceMaterials/DiffuseEmission.cpp:14-16, IBLEmission.cpp:15-17
// Sum / mix trees are flattened into one MultiBSDF level: selecting a nested MultiBSDF in proportion
// to its summed weight and then one of its lobes is the same distribution as selecting among all
// leaves, and pdf / value are the same weighted sums (MultiBSDF.cpp:20-59).
#pragma once
#include "bsdf.cuh"
#include "textures.cuh"

namespace slrgpu {

// MicrofacetReflection / MicrofacetScattering ignore `scale` (MicrofacetSurfaceMaterial.cpp:14-29): kept.
// MK >= 0 fixes the material kind at compile time (per-class material kernels), MK = -1 dispatches at run time.
template <int NC, int MK>
__device__ __forceinline__ void fillLobeT(const DeviceScene& s, const SlrGpuMaterial& m, const SurfPt& sp, float wlOffset, bool lambdaSelected,
                                          float scale, uint32_t inverse, Lobe<NC>* L) {
    L->inverse = inverse;
    L->f0 = 0.0f; L->f1 = 0.0f;
    switch (MK >= 0 ? (uint32_t)MK : m.kind) {
    case SLRGPU_MAT_DIFFUSE:
        L->s0 = scale * evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        if (m.tex[1] != SLRGPU_INVALID_ID) {
            const float sigma = evalFloatTexture(s, m.tex[1], sp);
            L->type = LOBE_OREN_NAYAR;
            L->f0 = (float)(1.0f - 0.5f * sigma * sigma / (sigma * sigma + 0.33));
            L->f1 = (float)(0.45 * sigma * sigma / (sigma * sigma + 0.09));
        } else {
            L->type = LOBE_LAMBERT;
        }
        L->baseDirType = DT_Reflection | DT_LowFreq;
        break;
    case SLRGPU_MAT_SPECULAR_REFLECTION:
        L->type = LOBE_SPECULAR_BRDF;
        L->s0 = scale * evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        L->s1 = evalSpectrumTexture<NC>(s, m.tex[1], sp, wlOffset);
        L->s2 = evalSpectrumTexture<NC>(s, m.tex[2], sp, wlOffset);
        L->baseDirType = DT_Reflection | DT_Delta0D;
        break;
    case SLRGPU_MAT_SPECULAR_SCATTERING:
        L->type = LOBE_SPECULAR_BSDF;
        L->s0 = scale * evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        L->s1 = evalSpectrumTexture<NC>(s, m.tex[1], sp, wlOffset);
        L->s2 = evalSpectrumTexture<NC>(s, m.tex[2], sp, wlOffset);
        L->baseDirType = DT_Reflection | DT_Transmission | DT_Delta0D | (lambdaSelected ? 0u : DT_Dispersive);
        break;
    case SLRGPU_MAT_WARD_DUR:
        L->type = LOBE_WARD;
        L->s0 = scale * evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        L->f0 = evalFloatTexture(s, m.tex[1], sp);
        L->f1 = evalFloatTexture(s, m.tex[2], sp);
        L->baseDirType = DT_Reflection | DT_HighFreq;
        break;
    case SLRGPU_MAT_ASHIKHMIN_SHIRLEY:
        L->type = LOBE_ASHIKHMIN;
        L->s0 = scale * evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);    // Rs
        L->s1 = scale * evalSpectrumTexture<NC>(s, m.tex[1], sp, wlOffset);    // Rd
        L->f0 = evalFloatTexture(s, m.tex[2], sp);
        L->f1 = evalFloatTexture(s, m.tex[3], sp);
        L->baseDirType = DT_Reflection | DT_HighFreq | DT_LowFreq;
        break;
    case SLRGPU_MAT_MICROFACET_REFLECTION:
        L->type = LOBE_MF_BRDF;
        L->s0 = evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        L->s1 = evalSpectrumTexture<NC>(s, m.tex[1], sp, wlOffset);
        L->f0 = evalFloatTexture(s, m.tex[2], sp);
        L->baseDirType = DT_Reflection | DT_HighFreq;
        break;
    case SLRGPU_MAT_MICROFACET_SCATTERING:
        L->type = LOBE_MF_BSDF;
        L->s0 = evalSpectrumTexture<NC>(s, m.tex[0], sp, wlOffset);
        L->s1 = evalSpectrumTexture<NC>(s, m.tex[1], sp, wlOffset);
        L->f0 = evalFloatTexture(s, m.tex[2], sp);
        L->baseDirType = DT_Reflection | DT_Transmission | DT_HighFreq;
        break;
    default:
        L->type = LOBE_LAMBERT;
        L->s0 = specConst<NC>(0.0f);
        L->baseDirType = 0;
        break;
    }
}

template <int NC>
static __device__ __noinline__ void fillLobe(const DeviceScene& s, const SlrGpuMaterial& m, const SurfPt& sp, float wlOffset, bool lambdaSelected,
                                float scale, uint32_t inverse, Lobe<NC>* L) {
    fillLobeT<NC, -1>(s, m, sp, wlOffset, lambdaSelected, scale, inverse, L);
}

// SVGGX evaluates alpha_g with a surface point that only carries the texture coordinate
// (surface_material.cpp:27-29 -> FloatTexture::evaluate(TexCoord2D), textures.h:84-88): a world-position
// mapped alpha texture therefore sees an indeterminate position in the reference; here it sees p.

template <int NC, int ML>
static __device__ __noinline__ void buildBsdf(const DeviceScene& s, uint32_t materialId, const SurfPt& sp, float wlOffset, bool lambdaSelected,
                                 Bsdf<NC, ML>* out) {
    out->numLobes = 0; out->multi = false; out->type = 0;
    // explicit DFS over the material tree
    struct Item { uint32_t mat; float scale; uint32_t inverse; };
    Item stack[8];
    int spn = 0;
    stack[spn++] = Item{materialId, 1.0f, 0u};
    while (spn > 0) {
        const Item it = stack[--spn];
        if (it.mat == SLRGPU_INVALID_ID) continue;
        const SlrGpuMaterial m = s.materials[it.mat];
        switch (m.kind) {
        case SLRGPU_MAT_EMITTER:
            stack[spn++] = Item{m.sub[0], it.scale, it.inverse};
            break;
        case SLRGPU_MAT_INVERSE:
            stack[spn++] = Item{m.sub[0], it.scale, it.inverse ^ 1u};
            break;
        case SLRGPU_MAT_SUMMED:
            out->multi = true;
            if (spn + 2 <= 8) { stack[spn++] = Item{m.sub[1], it.scale, it.inverse}; stack[spn++] = Item{m.sub[0], it.scale, it.inverse}; }
            break;
        case SLRGPU_MAT_MIXED: {
            out->multi = true;
            const float factor = evalFloatTexture(s, m.tex[0], sp);
            if (spn + 2 <= 8) {
                stack[spn++] = Item{m.sub[1], it.scale * factor, it.inverse};
                stack[spn++] = Item{m.sub[0], it.scale * (1.0f - factor), it.inverse};
            }
            break;
        }
        default:
            if (out->numLobes < ML) {
#pragma unroll
                for (int i = 0; i < ML; ++i)
                    if (i == out->numLobes) {
                        fillLobe<NC>(s, m, sp, wlOffset, lambdaSelected, it.scale, it.inverse, &out->lobes[i]);
                        out->type |= lobeDirType(out->lobes[i]);
                    }
                ++out->numLobes;
            }
            break;
        }
    }
}

// classifyMaterialIn (device_scene.h) on the device's material table
__device__ __forceinline__ uint32_t classifyMaterial(const DeviceScene& s, uint32_t materialId, uint32_t* leaf) {
    return classifyMaterialIn(s.materials, materialId, leaf);
}
// material kind behind a single-lobe class
__host__ __device__ constexpr int classMaterialKind(int cls) {
    return cls == 0 || cls == 1 ? (int)SLRGPU_MAT_DIFFUSE : cls == 2 ? (int)SLRGPU_MAT_SPECULAR_REFLECTION
         : cls == 3 ? (int)SLRGPU_MAT_SPECULAR_SCATTERING : cls == 4 ? (int)SLRGPU_MAT_WARD_DUR
         : cls == 5 ? (int)SLRGPU_MAT_ASHIKHMIN_SHIRLEY : cls == 6 ? (int)SLRGPU_MAT_MICROFACET_REFLECTION
         : cls == 7 ? (int)SLRGPU_MAT_MICROFACET_SCATTERING : -1;
}

__device__ __forceinline__ bool materialIsEmitting(const DeviceScene& s, uint32_t materialId) {
    return materialId != SLRGPU_INVALID_ID && s.materials[materialId].kind == SLRGPU_MAT_EMITTER;
}

// SurfaceMaterial::emittance for an emitter material (DiffuseEmission or IBLEmission behind sub[1])
template <int NC>
__device__ inline Spec<NC> materialEmittance(const DeviceScene& s, uint32_t materialId, const SurfPt& sp, float wlOffset) {
    const SlrGpuMaterial m = s.materials[materialId];
    const SlrGpuMaterial e = (m.kind == SLRGPU_MAT_EMITTER) ? s.materials[m.sub[1]] : m;
    const Spec<NC> v = evalSpectrumTexture<NC>(s, e.tex[0], sp, wlOffset);
    if (e.kind == SLRGPU_MAT_IBL_EMISSION) return (kPi * v) * e.f0;
    return v;
}

}  // namespace slrgpu
