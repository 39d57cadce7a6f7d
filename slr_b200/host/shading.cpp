#include "shading.h"
#include "assets/exr.h"
#include "assets/images.h"
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>

namespace slr {

size_t Image2D::texelSize(SlrGpuImageFormat f) {
    switch (f) {
        case SLRGPU_IMG_RGB8x3: return 3;
        case SLRGPU_IMG_RGB_8x4: case SLRGPU_IMG_RGBA8x4: return 4;
        case SLRGPU_IMG_RGBA16Fx4: case SLRGPU_IMG_UVSA16Fx4: return 8;
        case SLRGPU_IMG_GRAY8: return 1;
        case SLRGPU_IMG_UVS16Fx3: return 6;
        case SLRGPU_IMG_FLOAT32: return 4;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// factories
// ---------------------------------------------------------------------------------------------

static TextureMappingRef orDefault(const TextureMappingRef& m) { return m ? m : std::make_shared<TextureMapping>(); }

SpectrumTextureRef SpectrumTexture::constant(const InputSpectrumRef& s) {
    auto t = std::make_shared<SpectrumTexture>(); t->kind = SLRGPU_TEX_CONSTANT_SPECTRUM; t->spectrum[0] = s; return t;
}
SpectrumTextureRef SpectrumTexture::checkerBoard(const TextureMappingRef& m, const InputSpectrumRef& v0, const InputSpectrumRef& v1) {
    auto t = std::make_shared<SpectrumTexture>(); t->kind = SLRGPU_TEX_CHECKER_SPECTRUM; t->mapping = orDefault(m);
    t->spectrum[0] = v0; t->spectrum[1] = v1; return t;
}
SpectrumTextureRef SpectrumTexture::voronoi(const TextureMappingRef& m, float scale, float brightness) {
    auto t = std::make_shared<SpectrumTexture>(); t->kind = SLRGPU_TEX_VORONOI_SPECTRUM; t->mapping = orDefault(m);
    t->f0 = scale; t->f1 = brightness; return t;
}
SpectrumTextureRef SpectrumTexture::imageTexture(const TextureMappingRef& m, const Image2DRef& img) {
    auto t = std::make_shared<SpectrumTexture>(); t->kind = SLRGPU_TEX_IMAGE_SPECTRUM; t->mapping = orDefault(m); t->image = img; return t;
}
Normal3DTextureRef Normal3DTexture::checkerBoard(const TextureMappingRef& m, float stepWidth, bool reverse) {
    auto t = std::make_shared<Normal3DTexture>(); t->kind = SLRGPU_TEX_CHECKER_NORMAL; t->mapping = orDefault(m);
    t->f0 = stepWidth; t->i0 = reverse ? 1 : 0; return t;
}
Normal3DTextureRef Normal3DTexture::voronoi(const TextureMappingRef& m, float scale, float thetaMax) {
    auto t = std::make_shared<Normal3DTexture>(); t->kind = SLRGPU_TEX_VORONOI_NORMAL; t->mapping = orDefault(m);
    t->f0 = scale; t->f1 = std::cos(thetaMax); return t;     // stores cos(thetaMax), voronoi_textures.h:37
}
Normal3DTextureRef Normal3DTexture::imageTexture(const TextureMappingRef& m, const Image2DRef& img) {
    auto t = std::make_shared<Normal3DTexture>(); t->kind = SLRGPU_TEX_IMAGE_NORMAL; t->mapping = orDefault(m); t->image = img; return t;
}
FloatTextureRef FloatTexture::constant(float v) {
    auto t = std::make_shared<FloatTexture>(); t->kind = SLRGPU_TEX_CONSTANT_FLOAT; t->f0 = v; return t;
}
FloatTextureRef FloatTexture::checkerBoard(const TextureMappingRef& m, float v0, float v1) {
    auto t = std::make_shared<FloatTexture>(); t->kind = SLRGPU_TEX_CHECKER_FLOAT; t->mapping = orDefault(m); t->f0 = v0; t->f1 = v1; return t;
}
FloatTextureRef FloatTexture::voronoi(const TextureMappingRef& m, float scale, float valueScale, bool flat) {
    auto t = std::make_shared<FloatTexture>(); t->kind = SLRGPU_TEX_VORONOI_FLOAT; t->mapping = orDefault(m);
    t->f0 = scale; t->f1 = valueScale; t->i0 = flat ? 1 : 0; return t;
}
FloatTextureRef FloatTexture::imageTexture(const TextureMappingRef& m, const Image2DRef& img) {
    auto t = std::make_shared<FloatTexture>(); t->kind = SLRGPU_TEX_IMAGE_FLOAT; t->mapping = orDefault(m); t->image = img; return t;
}

static SurfaceMaterialRef mk(SlrGpuMaterialKind k) { auto m = std::make_shared<SurfaceMaterial>(); m->kind = k; return m; }

SurfaceMaterialRef SurfaceMaterial::createMatte(const SpectrumTextureRef& R, const FloatTextureRef& sigma) {
    auto m = mk(SLRGPU_MAT_DIFFUSE); m->stex[0] = R; m->ftex[1] = sigma; return m;
}
SurfaceMaterialRef SurfaceMaterial::createMetal(const SpectrumTextureRef& c, const SpectrumTextureRef& eta, const SpectrumTextureRef& k) {
    auto m = mk(SLRGPU_MAT_SPECULAR_REFLECTION); m->stex[0] = c; m->stex[1] = eta; m->stex[2] = k; return m;
}
SurfaceMaterialRef SurfaceMaterial::createGlass(const SpectrumTextureRef& c, const SpectrumTextureRef& e, const SpectrumTextureRef& i) {
    auto m = mk(SLRGPU_MAT_SPECULAR_SCATTERING); m->stex[0] = c; m->stex[1] = e; m->stex[2] = i; return m;
}
SurfaceMaterialRef SurfaceMaterial::createModifiedWardDur(const SpectrumTextureRef& R, const FloatTextureRef& ax, const FloatTextureRef& ay) {
    auto m = mk(SLRGPU_MAT_WARD_DUR); m->stex[0] = R; m->ftex[1] = ax; m->ftex[2] = ay; return m;
}
SurfaceMaterialRef SurfaceMaterial::createAshikhminShirley(const SpectrumTextureRef& Rd, const SpectrumTextureRef& Rs,
                                                           const FloatTextureRef& nu, const FloatTextureRef& nv) {
    auto m = mk(SLRGPU_MAT_ASHIKHMIN_SHIRLEY); m->stex[0] = Rs; m->stex[1] = Rd; m->ftex[2] = nu; m->ftex[3] = nv; return m;
}
SurfaceMaterialRef SurfaceMaterial::createMicrofacetMetal(const SpectrumTextureRef& eta, const SpectrumTextureRef& k, const FloatTextureRef& a) {
    auto m = mk(SLRGPU_MAT_MICROFACET_REFLECTION); m->stex[0] = eta; m->stex[1] = k; m->ftex[2] = a; return m;
}
SurfaceMaterialRef SurfaceMaterial::createMicrofacetGlass(const SpectrumTextureRef& e, const SpectrumTextureRef& i, const FloatTextureRef& a) {
    auto m = mk(SLRGPU_MAT_MICROFACET_SCATTERING); m->stex[0] = e; m->stex[1] = i; m->ftex[2] = a; return m;
}
SurfaceMaterialRef SurfaceMaterial::createInverseMaterial(const SurfaceMaterialRef& base) {
    auto m = mk(SLRGPU_MAT_INVERSE); m->sub[0] = base; return m;
}
SurfaceMaterialRef SurfaceMaterial::createSummedMaterial(const SurfaceMaterialRef& m0, const SurfaceMaterialRef& m1) {
    auto m = mk(SLRGPU_MAT_SUMMED); m->sub[0] = m0; m->sub[1] = m1; return m;
}
SurfaceMaterialRef SurfaceMaterial::createMixedMaterial(const SurfaceMaterialRef& m0, const SurfaceMaterialRef& m1, const FloatTextureRef& f) {
    auto m = mk(SLRGPU_MAT_MIXED); m->sub[0] = m0; m->sub[1] = m1; m->ftex[0] = f; return m;
}
EmitterSurfacePropertyRef SurfaceMaterial::createDiffuseEmitter(const SpectrumTextureRef& emittance) {
    auto e = std::make_shared<EmitterSurfaceProperty>(); e->kind = SLRGPU_MAT_DIFFUSE_EMISSION; e->emittance = emittance; return e;
}
SurfaceMaterialRef SurfaceMaterial::createEmitterSurfaceMaterial(const SurfaceMaterialRef& mat, const EmitterSurfacePropertyRef& emit) {
    if (mat && mat->isEmitting()) throw std::runtime_error("EmitterSurfaceMaterial cannot wrap an emitting material");
    auto m = mk(SLRGPU_MAT_EMITTER); m->sub[0] = mat; m->emitter = emit; return m;
}

// ---------------------------------------------------------------------------------------------
// export (deduplicated by object identity)
// ---------------------------------------------------------------------------------------------

namespace {
struct ExportCache {
    std::map<const void*, uint32_t> spectra, textures, materials, images, emitters;
};
std::map<const GpuSceneBuilder*, ExportCache>& caches() { static std::map<const GpuSceneBuilder*, ExportCache> c; return c; }

uint32_t exportSpectrum(GpuSceneBuilder& b, const InputSpectrum* s) {
    if (!s) throw std::runtime_error("texture refers to a null spectrum");
    ExportCache& c = caches()[&b];
    auto it = c.spectra.find(s);
    if (it != c.spectra.end()) return it->second;
    uint32_t id = s->exportTo(b.flat.spectra, b.flat.spectrumData);
    c.spectra[s] = id;
    return id;
}

uint32_t exportImage(GpuSceneBuilder& b, const Image2D* img) {
    if (!img) throw std::runtime_error("image texture without an image");
    ExportCache& c = caches()[&b];
    auto it = c.images.find(img);
    if (it != c.images.end()) return it->second;
    SlrGpuImage g = {};
    g.format = img->format; g.width = img->width; g.height = img->height;
    while (b.flat.imageData.size() % 16) b.flat.imageData.push_back(0);
    g.data_offset = b.flat.imageData.size();
    g.spectrum_type = (uint32_t)img->spectrumType;
    b.flat.imageData.insert(b.flat.imageData.end(), img->data.begin(), img->data.end());
    b.flat.images.push_back(g);
    return c.images[img] = (uint32_t)b.flat.images.size() - 1;
}

void fillMapping(SlrGpuTexture& t, const TextureMappingRef& m) {
    t.mapping = m ? m->kind : SLRGPU_MAP_TEXCOORD;
    t.map_offset[0] = m ? m->offset[0] : 0; t.map_offset[1] = m ? m->offset[1] : 0;
    t.map_scale[0] = m ? m->scale[0] : 1;   t.map_scale[1] = m ? m->scale[1] : 1;
}

uint32_t exportSpectrumTexture(GpuSceneBuilder& b, const SpectrumTexture* t) {
    ExportCache& c = caches()[&b];
    auto it = c.textures.find(t);
    if (it != c.textures.end()) return it->second;
    SlrGpuTexture g = {};
    g.kind = t->kind;
    fillMapping(g, t->mapping);
    g.i0 = g.i1 = SLRGPU_INVALID_ID;
    g.f0 = t->f0; g.f1 = t->f1;
    switch (t->kind) {
        case SLRGPU_TEX_CONSTANT_SPECTRUM: g.i0 = exportSpectrum(b, t->spectrum[0].get()); break;
        case SLRGPU_TEX_CHECKER_SPECTRUM:
            g.i0 = exportSpectrum(b, t->spectrum[0].get()); g.i1 = exportSpectrum(b, t->spectrum[1].get()); break;
        case SLRGPU_TEX_VORONOI_SPECTRUM: break;
        case SLRGPU_TEX_IMAGE_SPECTRUM: g.i0 = exportImage(b, t->image.get()); break;
        default: throw std::runtime_error("not a spectrum texture kind");
    }
    b.flat.textures.push_back(g);
    return c.textures[t] = (uint32_t)b.flat.textures.size() - 1;
}

uint32_t exportEmitter(GpuSceneBuilder& b, const EmitterSurfaceProperty* e) {
    ExportCache& c = caches()[&b];
    auto it = c.emitters.find(e);
    if (it != c.emitters.end()) return it->second;
    SlrGpuMaterial g = {};
    g.kind = e->kind;
    for (int i = 0; i < 4; ++i) g.tex[i] = SLRGPU_INVALID_ID;
    g.sub[0] = g.sub[1] = SLRGPU_INVALID_ID;
    g.tex[0] = exportSpectrumTexture(b, e->emittance.get());
    g.f0 = e->scale;
    b.flat.materials.push_back(g);
    return c.emitters[e] = (uint32_t)b.flat.materials.size() - 1;
}
}  // namespace

uint32_t GpuSceneBuilder::exportFloatTexture(const FloatTexture* t) {
    ExportCache& c = caches()[this];
    auto it = c.textures.find(t);
    if (it != c.textures.end()) return it->second;
    SlrGpuTexture g = {};
    g.kind = t->kind;
    fillMapping(g, t->mapping);
    g.i0 = t->i0; g.i1 = SLRGPU_INVALID_ID;
    g.f0 = t->f0; g.f1 = t->f1;
    if (t->kind == SLRGPU_TEX_IMAGE_FLOAT) g.i0 = exportImage(*this, t->image.get());
    flat.textures.push_back(g);
    return c.textures[t] = (uint32_t)flat.textures.size() - 1;
}

uint32_t GpuSceneBuilder::exportNormalTexture(const Normal3DTexture* t) {
    ExportCache& c = caches()[this];
    auto it = c.textures.find(t);
    if (it != c.textures.end()) return it->second;
    SlrGpuTexture g = {};
    g.kind = t->kind;
    fillMapping(g, t->mapping);
    g.i0 = t->i0; g.i1 = SLRGPU_INVALID_ID;
    g.f0 = t->f0; g.f1 = t->f1;
    if (t->kind == SLRGPU_TEX_IMAGE_NORMAL) g.i0 = exportImage(*this, t->image.get());
    flat.textures.push_back(g);
    return c.textures[t] = (uint32_t)flat.textures.size() - 1;
}

uint32_t GpuSceneBuilder::exportMaterial(const SurfaceMaterial* m) {
    ExportCache& c = caches()[this];
    auto it = c.materials.find(m);
    if (it != c.materials.end()) return it->second;
    SlrGpuMaterial g = {};
    g.kind = m->kind;
    for (int i = 0; i < 4; ++i) {
        g.tex[i] = SLRGPU_INVALID_ID;
        if (m->stex[i]) g.tex[i] = exportSpectrumTexture(*this, m->stex[i].get());
        else if (m->ftex[i]) g.tex[i] = exportFloatTexture(m->ftex[i].get());
    }
    for (int i = 0; i < 2; ++i) g.sub[i] = m->sub[i] ? exportMaterial(m->sub[i].get()) : SLRGPU_INVALID_ID;
    if (m->kind == SLRGPU_MAT_EMITTER) {
        if (!m->emitter) throw std::runtime_error("emitter material without an emitter property");
        g.sub[1] = exportEmitter(*this, m->emitter.get());
    }
    flat.materials.push_back(g);
    return c.materials[m] = (uint32_t)flat.materials.size() - 1;
}

bool GpuSceneBuilder::materialEmits(const SurfaceMaterial* m) const { return m && m->isEmitting(); }

// InfiniteSphereSurfaceObject + IBLEmission: the environment's emitter material and the importance map
// of ImageSpectrumTexture::createIBLImportanceMap (image_textures.cpp:81-134) -- one cell per 4x4 texel
// block, weight = sin(theta) * luminance of the block's area average (Image2D::areaAverage,
// Image.cpp:19-330, whose summation order and half-float round trips are kept) -- as the
// RegularConstantContinuous2D the reference builds from it (distributions.cpp:126-206).
namespace {
struct KahanSum {
    float result = 0.0f, comp = 0.0f;
    void add(float v) { float c = v - comp; float t = result + c; comp = (t - result) - c; result = t; }
};
float halfRound(float f) { return exr::halfToFloat(exr::floatToHalf(f)); }

// luminance of the area average of the 4x4 block (bx, by)
float blockLuminance(const Image2D& img, uint32_t bx, uint32_t by, uint32_t dx, uint32_t dy) {
    const uint16_t* tex = reinterpret_cast<const uint16_t*>(img.data.data());
    const bool uvs = img.format == SLRGPU_IMG_UVSA16Fx4;
    KahanSum sr, sg, sb;
    auto addTexel = [&](uint32_t x, uint32_t y) {
        const uint16_t* t = tex + 4 * ((size_t)y * img.width + x);
        float v[3] = {exr::halfToFloat(t[0]), exr::halfToFloat(t[1]), exr::halfToFloat(t[2])};
        if (uvs) { float rgb[3]; uvs_to_sRGB(img.spectrumType, v, rgb); sr.add(rgb[0]); sg.add(rgb[1]); sb.add(rgb[2]); }
        else { sr.add(v[0]); sg.add(v[1]); sb.add(v[2]); }
    };
    const uint32_t x0 = bx * dx, x1 = x0 + dx - 1, y0 = by * dy, y1 = y0 + dy - 1;
    addTexel(x0, y0); addTexel(x1, y0); addTexel(x0, y1); addTexel(x1, y1);                    // corners
    for (uint32_t x = x0 + 1; x < x1; ++x) { addTexel(x, y0); addTexel(x, y1); }              // top / bottom edges
    for (uint32_t y = y0 + 1; y < y1; ++y) { addTexel(x0, y); addTexel(x1, y); }              // left / right edges
    for (uint32_t y = y0 + 1; y < y1; ++y) for (uint32_t x = x0 + 1; x < x1; ++x) addTexel(x, y);
    const float area = (float)dy * (float)dx;
    float rgb[3] = {sr.result / area, sg.result / area, sb.result / area};
    if (uvs) {
        float q[3];
        sRGB_to_uvs(img.spectrumType, rgb, q);
        for (int i = 0; i < 3; ++i) q[i] = halfRound(q[i]);
        uvs_to_sRGB(img.spectrumType, q, rgb);
    } else {
        for (int i = 0; i < 3; ++i) rgb[i] = halfRound(rgb[i]);
    }
    return 0.222485f * rgb[0] + 0.716905f * rgb[1] + 0.060610f * rgb[2];
}

// RegularConstantContinuous1D ctor on pdf[0..n) in place; cdf gets n + 1 entries; returns the integral
float buildContinuous1D(float* pdf, float* cdf, uint32_t n) {
    KahanSum sum;
    cdf[0] = 0.0f;
    for (uint32_t i = 0; i < n; ++i) { sum.add(pdf[i] / n); cdf[i + 1] = sum.result; }
    const float integral = sum.result;
    for (uint32_t i = 0; i < n; ++i) { pdf[i] /= integral; cdf[i + 1] /= integral; }
    return integral;
}
}  // namespace

void exportEnvironment(GpuSceneBuilder& b, const InfiniteSphereNode& env) {
    if (!env.emission || !env.emission->emittance || !env.emission->emittance->image)
        throw std::runtime_error("the environment needs an image texture");
    const Image2D& img = *env.emission->emittance->image;
    if (img.format != SLRGPU_IMG_UVSA16Fx4 && img.format != SLRGPU_IMG_RGBA16Fx4)
        throw std::runtime_error("the environment image must be a half-float RGBA image");
    FlatScene& f = b.flat;
    const uint32_t W = img.width / 4, H = img.height / 4;
    if (W == 0 || H == 0) throw std::runtime_error("the environment image is too small");
    const uint32_t dx = img.width / W, dy = img.height / H;
    f.envPresent = true;
    f.envMaterial = exportEmitter(b, env.emission.get());
    f.envMapWidth = W; f.envMapHeight = H;
    f.envRowPdf.assign((size_t)W * H, 0.0f);
    f.envRowCdf.assign((size_t)(W + 1) * H, 0.0f);
    f.envRowIntegral.assign(H, 0.0f);
    f.envMarginalPdf.assign(H, 0.0f);
    f.envMarginalCdf.assign(H + 1, 0.0f);
    for (uint32_t y = 0; y < H; ++y) {
        const double sinTheta = std::sin(M_PI * (y + 0.5f) / H);
        float* pdf = &f.envRowPdf[(size_t)y * W];
        for (uint32_t x = 0; x < W; ++x) pdf[x] = (float)(sinTheta * blockLuminance(img, x, y, dx, dy));
        f.envRowIntegral[y] = buildContinuous1D(pdf, &f.envRowCdf[(size_t)y * (W + 1)], W);
        f.envMarginalPdf[y] = f.envRowIntegral[y];
    }
    f.envMarginalIntegral = buildContinuous1D(f.envMarginalPdf.data(), f.envMarginalCdf.data(), H);
}

void finishShadingTables(GpuSceneBuilder& b) { caches().erase(&b); }

}  // namespace slr
