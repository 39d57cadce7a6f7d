// Procedural and image textures as device functions.
//   mappings                  libSLR/Core/textures.h:16-50
//   checker board             libSLR/Textures/checker_board_textures.h:22-25,48-51, .cpp:16-44
//   Voronoi / Worley          libSLR/Textures/voronoi_textures.cpp:36-170 (FNV-1 hash :20-26, LCG RNGs/LinearCongruentialRNG.cpp:12-14)
//   image textures            libSLR/Textures/image_textures.cpp:13-79,136-209
// The Voronoi cell pattern is defined by integer arithmetic (hash + LCG) and must match the reference
// exactly; the reference builds feature points as Point3D(ix + rng(), iy + rng(), iz + rng()) whose
// argument evaluation order is the compiler's -- kVoronoiArgsRightToLeft records what the oracle
// build (g++, x86-64 SysV) does, and tests/test_textures pins it.
#pragma once
#include <cuda_fp16.h>
#include "spectral.cuh"
#include "vecmath.cuh"

namespace slrgpu {

struct SurfPt {
    V3 p;
    V3 gn;            // geometric normal
    Frame sf;         // shading frame
    float u, v;       // surface parameters (b0, b1) or (phi, theta) on the infinite sphere
    float tu, tv;     // texture coordinate
    uint32_t prim, inst;
    bool atInfinity;
};

__device__ __forceinline__ V3 mapTexture(const SlrGpuTexture& t, const SurfPt& sp) {
    if (t.mapping == SLRGPU_MAP_WORLD_POS) return sp.p;
    if (t.mapping == SLRGPU_MAP_OFFSET_SCALE_2D)
        return V3((sp.tu + t.map_offset[0]) * t.map_scale[0], (sp.tv + t.map_offset[1]) * t.map_scale[1], 0.0f);
    return V3(sp.tu, sp.tv, 0.0f);
}

// ---- checker board -----------------------------------------------------------------------------
__device__ __forceinline__ int checkerIndex(const V3& tc) { return ((int)(tc.x * 2) + (int)(tc.y * 2)) % 2; }

__device__ inline V3 checkerNormal(const SlrGpuTexture& t, const SurfPt& sp) {
    const V3 tc = mapTexture(t, sp);
    const float halfWidth = t.f0 * 0.5f;
    float uComp = 0.0f;
    const float absWrapU = fmodf(fabsf(tc.x), 1.0f);
    if (absWrapU < halfWidth * 0.5f || absWrapU > 1.0f - halfWidth * 0.5f) uComp = 1.0f;
    else if (absWrapU > 0.5f - halfWidth * 0.5f && absWrapU < 0.5f + halfWidth * 0.5f) uComp = -1.0f;
    float vComp = 0.0f;
    const float absWrapV = fmodf(fabsf(tc.y), 1.0f);
    if (absWrapV < halfWidth * 0.5f || absWrapV > 1.0f - halfWidth * 0.5f) vComp = 1.0f;
    else if (absWrapV > 0.5f - halfWidth * 0.5f && absWrapV < 0.5f + halfWidth * 0.5f) vComp = -1.0f;
    if (absWrapV > 0.5f) uComp *= -1;
    if (absWrapU > 0.5f) vComp *= -1;
    if (t.i0) { uComp *= -1; vComp *= -1; }
    return normalize(V3(uComp, vComp, 1.0f));
}

// ---- Voronoi -----------------------------------------------------------------------------------
constexpr bool kVoronoiArgsRightToLeft = true;

struct Lcg32 {
    uint32_t s;
    __device__ __forceinline__ uint32_t next() { return (s = s * 1103515245u + 12345u); }
    __device__ __forceinline__ float nextFloat() { return __uint_as_float((next() >> 9) | 0x3f800000u) - 1.0f; }
};

__device__ __forceinline__ uint32_t fnv1Hash32(const int32_t c[3]) {
    uint32_t hash = 2166136261u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t w = (uint32_t)c[k];
#pragma unroll
        for (int b = 0; b < 4; ++b) hash = (16777619u * hash) ^ ((w >> (8 * b)) & 0xFFu);
    }
    return hash;
}

struct VoronoiCell {
    float closestDistance;
    uint32_t hashOfClosest, closestFPIdx;
};

static __device__ __noinline__ VoronoiCell voronoiClosest(const V3& evalp) {
    int32_t ie[3] = {(int32_t)floorf(evalp.x), (int32_t)floorf(evalp.y), (int32_t)floorf(evalp.z)};
    const int32_t rbx = -1 + (int32_t)roundf(evalp.x - ie[0]);
    const int32_t rby = -1 + (int32_t)roundf(evalp.y - ie[1]);
    const int32_t rbz = -1 + (int32_t)roundf(evalp.z - ie[2]);
    VoronoiCell r;
    r.closestDistance = CUDART_INF_F; r.hashOfClosest = 0; r.closestFPIdx = 0;
    for (int iz = rbz; iz < rbz + 2; ++iz)
        for (int iy = rby; iy < rby + 2; ++iy)
            for (int ix = rbx; ix < rbx + 2; ++ix) {
                const int32_t ic[3] = {ie[0] + ix, ie[1] + iy, ie[2] + iz};
                const uint32_t hash = fnv1Hash32(ic);
                Lcg32 rng{hash};
                const int32_t nf = (int32_t)(8 * rng.nextFloat());
                const uint32_t numFeaturePoints = 1 + (uint32_t)(nf < 8 ? nf : 8);
                for (uint32_t i = 0; i < numFeaturePoints; ++i) {
                    float a = rng.nextFloat(), b = rng.nextFloat(), c = rng.nextFloat();
                    V3 fp = kVoronoiArgsRightToLeft ? V3(ic[0] + c, ic[1] + b, ic[2] + a) : V3(ic[0] + a, ic[1] + b, ic[2] + c);
                    const float dist = length(evalp - fp);
                    if (dist < r.closestDistance) { r.closestDistance = dist; r.hashOfClosest = hash; r.closestFPIdx = i; }
                }
            }
    return r;
}

// ---- colour conversion used by per-texel / per-cell up-sampling (SpectrumTypes.h:180-237) ------
__device__ __forceinline__ float sRGBDegamma(float v) {
    if (v <= 0.04045f) return v / 12.92f;
    return powf((v + 0.055f) / 1.055f, 2.4f);
}
// (Reflectance, sRGB) -> (u, v, scale) of the Meng-Simon parameterisation
__device__ inline void sRGBReflectanceToUVS(float r, float g, float b, float* u, float* v, float* scale) {
    const float X = 0.4969f * r + 0.3391f * g + 0.1640f * b;
    const float Y = 0.2562f * r + 0.6782f * g + 0.0656f * b;
    const float Z = 0.0233f * r + 0.1130f * g + 0.8637f * b;
    const float brightness = X + Y + Z;
    if (brightness == 0) { *u = 6; *v = 4; *scale = 0; return; }
    const float x = X / brightness, y = Y / brightness;
    *scale = brightness / 0.009355121400914532f;
    *u = 16.730260708356887f * x + 7.7801960340706f * y - 2.170152247475828f;
    *v = -7.530081094743006f * x + 16.192422314095225f * y + 1.1125529268825947f;
}

// ---- images ------------------------------------------------------------------------------------
__device__ __forceinline__ void imageTexel(const SlrGpuImage& img, const V3& tc, uint32_t* px, uint32_t* py) {
    float u = fmodf(tc.x, 1.0f), v = fmodf(tc.y, 1.0f);
    u += u < 0 ? 1.0f : 0.0f;
    v += v < 0 ? 1.0f : 0.0f;
    *px = min((uint32_t)(img.width * u), img.width - 1);
    *py = min((uint32_t)(img.height * v), img.height - 1);
}
__device__ __forceinline__ float halfBitsToFloat(uint16_t h) { return __half2float(__ushort_as_half(h)); }

template <int NC>
static __device__ __noinline__ Spec<NC> imageSpectrum(const DeviceScene& s, const SlrGpuTexture& t, const SurfPt& sp, float wlOffset) {
    const SlrGpuImage img = s.images[t.i0];
    uint32_t px, py;
    imageTexel(img, mapTexture(t, sp), &px, &py);
    const uint8_t* base = s.imageData + img.data_offset;
    const size_t idx = (size_t)py * img.width + px;
    Spec<NC> ret = specConst<NC>(0.0f);
    switch (img.format) {
    case SLRGPU_IMG_UVS16Fx3: {
        const uint16_t* d = reinterpret_cast<const uint16_t*>(base) + idx * 3;
        if (NC != 3) ret = evalUVS<NC>(s, halfBitsToFloat(d[0]), halfBitsToFloat(d[1]), halfBitsToFloat(d[2]) / 0.009355121400914532f, wlOffset);
        break;
    }
    case SLRGPU_IMG_UVSA16Fx4: {
        const uint16_t* d = reinterpret_cast<const uint16_t*>(base) + idx * 4;
        if (NC != 3) ret = evalUVS<NC>(s, halfBitsToFloat(d[0]), halfBitsToFloat(d[1]), halfBitsToFloat(d[2]) / 0.009355121400914532f, wlOffset);
        break;
    }
    case SLRGPU_IMG_GRAY8: ret = specConst<NC>(base[idx] / 255.0f); break;
    case SLRGPU_IMG_RGB8x3:
        if (NC == 3) { const uint8_t* d = base + idx * 3; ret.v[0] = d[0] / 255.0f; ret.v[1] = d[1] / 255.0f; ret.v[2] = d[2] / 255.0f; }
        break;
    case SLRGPU_IMG_RGB_8x4:
    case SLRGPU_IMG_RGBA8x4:
        if (NC == 3) { const uint8_t* d = base + idx * 4; ret.v[0] = d[0] / 255.0f; ret.v[1] = d[1] / 255.0f; ret.v[2] = d[2] / 255.0f; }
        break;
    case SLRGPU_IMG_RGBA16Fx4:
        if (NC == 3) {
            const uint16_t* d = reinterpret_cast<const uint16_t*>(base) + idx * 4;
            ret.v[0] = halfBitsToFloat(d[0]); ret.v[1] = halfBitsToFloat(d[1]); ret.v[2] = halfBitsToFloat(d[2]);
        }
        break;
    default: break;
    }
    return ret;
}

// ---- dispatch ----------------------------------------------------------------------------------
template <int NC>
static __device__ __noinline__ Spec<NC> evalSpectrumTexture(const DeviceScene& s, uint32_t texId, const SurfPt& sp, float wlOffset) {
    const SlrGpuTexture t = s.textures[texId];
    switch (t.kind) {
    case SLRGPU_TEX_CONSTANT_SPECTRUM: return evalInputSpectrum<NC>(s, t.i0, wlOffset);
    case SLRGPU_TEX_CHECKER_SPECTRUM: return evalInputSpectrum<NC>(s, checkerIndex(mapTexture(t, sp)) ? t.i1 : t.i0, wlOffset);
    case SLRGPU_TEX_VORONOI_SPECTRUM: {
        const VoronoiCell c = voronoiClosest(mapTexture(t, sp) / t.f0);
        Lcg32 rng{c.hashOfClosest + c.closestFPIdx};
        const float r = rng.nextFloat() * t.f1, g = rng.nextFloat() * t.f1, b = rng.nextFloat() * t.f1;
        if (NC == 3) { Spec<NC> o = specConst<NC>(0.0f); o.v[0] = r; o.v[1] = g; o.v[2] = b; return o; }
        float u, v, scale;
        sRGBReflectanceToUVS(sRGBDegamma(r), sRGBDegamma(g), sRGBDegamma(b), &u, &v, &scale);
        return evalUVS<NC>(s, u, v, scale, wlOffset);
    }
    case SLRGPU_TEX_IMAGE_SPECTRUM: return imageSpectrum<NC>(s, t, sp, wlOffset);
    default: return specConst<NC>(0.0f);
    }
}

static __device__ __noinline__ float evalFloatTexture(const DeviceScene& s, uint32_t texId, const SurfPt& sp) {
    const SlrGpuTexture t = s.textures[texId];
    switch (t.kind) {
    case SLRGPU_TEX_CONSTANT_FLOAT: return t.f0;
    case SLRGPU_TEX_CHECKER_FLOAT: return checkerIndex(mapTexture(t, sp)) ? t.f1 : t.f0;
    case SLRGPU_TEX_VORONOI_FLOAT: {
        const VoronoiCell c = voronoiClosest(mapTexture(t, sp) / t.f0);
        if (t.i0) { Lcg32 rng{c.hashOfClosest + c.closestFPIdx}; return t.f1 * rng.nextFloat(); }
        return (float)((c.closestDistance / (1.414213562 * t.f0)) * t.f1);
    }
    case SLRGPU_TEX_IMAGE_FLOAT: {
        const SlrGpuImage img = s.images[t.i0];
        uint32_t px, py;
        imageTexel(img, mapTexture(t, sp), &px, &py);
        const uint8_t* base = s.imageData + img.data_offset;
        const size_t idx = (size_t)py * img.width + px;
        if (img.format == SLRGPU_IMG_GRAY8) return base[idx] / 255.0f;
        if (img.format == SLRGPU_IMG_UVSA16Fx4) return halfBitsToFloat(reinterpret_cast<const uint16_t*>(base)[idx * 4 + 3]);
        if (img.format == SLRGPU_IMG_RGBA16Fx4) return halfBitsToFloat(reinterpret_cast<const uint16_t*>(base)[idx * 4 + 3]);
        if (img.format == SLRGPU_IMG_FLOAT32) return reinterpret_cast<const float*>(base)[idx];
        return 0.0f;
    }
    default: return 0.0f;
    }
}

static __device__ __noinline__ V3 evalNormalTexture(const DeviceScene& s, uint32_t texId, const SurfPt& sp) {
    const SlrGpuTexture t = s.textures[texId];
    switch (t.kind) {
    case SLRGPU_TEX_CHECKER_NORMAL: return checkerNormal(t, sp);
    case SLRGPU_TEX_VORONOI_NORMAL: {
        const VoronoiCell c = voronoiClosest(mapTexture(t, sp) / t.f0);
        Lcg32 rng{c.hashOfClosest + c.closestFPIdx};
        float a = rng.nextFloat(), b = rng.nextFloat();
        return kVoronoiArgsRightToLeft ? uniformSampleCone(b, a, t.f1) : uniformSampleCone(a, b, t.f1);
    }
    case SLRGPU_TEX_IMAGE_NORMAL: {
        const SlrGpuImage img = s.images[t.i0];
        uint32_t px, py;
        imageTexel(img, mapTexture(t, sp), &px, &py);
        const uint8_t* base = s.imageData + img.data_offset;
        const size_t idx = (size_t)py * img.width + px;
        const int bpp = img.format == SLRGPU_IMG_RGB8x3 ? 3 : 4;
        if (img.format == SLRGPU_IMG_RGB8x3 || img.format == SLRGPU_IMG_RGB_8x4 || img.format == SLRGPU_IMG_RGBA8x4) {
            const uint8_t* d = base + idx * bpp;
            return normalize(V3(d[0] / 255.0f - 0.5f, d[1] / 255.0f - 0.5f, d[2] / 255.0f - 0.5f));
        }
        return V3(0, 0, 1);
    }
    default: return V3(0, 0, 1);
    }
}

}  // namespace slrgpu
