// See GPUPathTracingRenderer.h. The exporter below reads the reference's private members (QBVH::m_nodes,
// SurfaceObjectAggregate::m_lightList, the materials' texture pointers, ...): the oracle build compiles this file with
// -fno-access-control; in a real merge these would be `friend class GpuSceneExporter;` lines or accessors.
// Instancing nested in instancing is expanded to the one level the device walks (expandNesting); animated transforms are
// exported with the reference's own decomposition. Unsupported content (a chain of two animated transforms, surfaces
// other than triangles) throws std::runtime_error -- never a silently different image.
#include "GPUPathTracingRenderer.h"

#include <libSLR/Accelerator/QBVH.h>
#include <libSLR/Accelerator/SBVH.h>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include <libSLR/Cameras/PerspectiveCamera.h>
#include <libSLR/Core/Image.h>
#include <libSLR/Core/ImageSensor.h>
#include <libSLR/Core/RenderSettings.h>
#include <libSLR/Core/SurfaceObject.h>
#include <libSLR/Core/Transform.h>
#include <libSLR/Core/distributions.h>
#include <libSLR/Core/surface_material.h>
#include <libSLR/Core/textures.h>
#include <libSLR/Memory/ArenaAllocator.h>
#include <libSLR/Surface/TriangleMesh.h>
#include <libSLR/SurfaceMaterials/AshikhminShirleyReflection.h>
#include <libSLR/SurfaceMaterials/DiffuseEmission.h>
#include <libSLR/SurfaceMaterials/IBLEmission.h>
#include <libSLR/SurfaceMaterials/MicrofacetSurfaceMaterial.h>
#include <libSLR/SurfaceMaterials/MixedSurfaceMaterial.h>
#include <libSLR/SurfaceMaterials/ModifiedWardDurReflection.h>
#include <libSLR/SurfaceMaterials/SummedSurfaceMaterial.h>
#include <libSLR/SurfaceMaterials/basic_SurfaceMaterials.h>
#include <libSLR/Textures/checker_board_textures.h>
#include <libSLR/Textures/constant_textures.h>
#include <libSLR/Textures/image_textures.h>
#include <libSLR/Textures/voronoi_textures.h>

#include <slrgpu.h>

#include <chrono>
#include <memory>
#include <set>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace SLR {

namespace {

[[noreturn]] void unsupported(const std::string &what) { throw std::runtime_error("GPUPathTracingRenderer: " + what); }

// Flattens SLR::Scene (libSLR/Core/SurfaceObject.h:239-260) into the tables of SlrGpuSceneDesc.
class GpuSceneExporter {
    std::vector<SlrGpuBvhNode> nodes;
    std::vector<SlrGpuLeafRecord> leaves;
    std::vector<SlrGpuInstance> instances;
    std::vector<SlrGpuMotion> motions;
    std::vector<SlrGpuTriangle> triangles;
    std::vector<SlrGpuVertex> vertices;
    std::vector<SlrGpuMaterial> materials;
    std::vector<SlrGpuTexture> textures;
    std::vector<SlrGpuSpectrum> spectra;
    std::vector<float> spectrumData;
    std::vector<SlrGpuLight> lights;
    std::vector<float> gridFloats;
    std::vector<SlrGpuImage> images;
    std::vector<uint8_t> imageData;
    std::map<const void*, uint32_t> imageIds;
    // environment importance map (InfiniteSphereSurfaceObject::m_dist), row-major copies of the reference's own arrays
    std::vector<float> envRowPdf, envRowCdf, envRowIntegral, envMarginalPdf, envMarginalCdf;
    uint32_t numTopLights = 0;
    float topLightImportance = 0.0f;

    std::map<const Vertex*, uint32_t> vertexIds;
    std::map<const SurfaceObject*, uint32_t> triangleIds, instanceIds;
    std::map<const void*, uint32_t> materialIds, textureIds, spectrumIds;
    struct AggregateInfo { uint32_t nodeBase, lightBase, numLights; float importance; bool hasInstances; };
    std::map<const SurfaceObjectAggregate*, AggregateInfo> aggregates;

    // ---- instancing nested in instancing --------------------------------------------------------------------------
    // The device walks ONE level of instances (include/slrgpu.h). A TransformedSurfaceObject whose aggregate holds further
    // TransformedSurfaceObjects (the reference recurses to any depth, SurfaceObject.cpp:307-336) is expanded here before
    // anything is exported: its own triangles become one aggregate placed under its transform, every instance inside it is
    // placed again under the composed transform (recursively), and the parent aggregate is rebuilt over the expanded list.
    // All of it with the reference's own classes -- SurfaceObjectAggregate's constructor builds the SBVH and the light
    // distribution, ChainedTransform::reduce composes static and animated transforms as the scene graph does
    // (Transform.cpp:74-152) -- so the exporter below sees an ordinary one-level scene. Light selection is unchanged in
    // distribution: an aggregate's importance is the sum of its entries' importances (SurfaceObject.cpp:232-252), so the
    // product of pmfs along a chain equals the pmf of the expanded entry.
    ArenaAllocator expandMem;
    std::vector<std::unique_ptr<SurfaceObject>> expandOwned;
    std::map<const SurfaceObjectAggregate*, SurfaceObjectAggregate*> ownTriangles;     // aggregate -> aggregate of its triangles only
    std::map<const SurfaceObject*, SurfaceObjectAggregate*> wrapped;                    // lone object -> one-object aggregate

    static std::vector<const SurfaceObject*> objectsOf(const SurfaceObjectAggregate* ag) {
        auto sbvh = dynamic_cast<const SBVH*>(ag->m_accelerator);
        if (!sbvh) unsupported("the aggregate's accelerator is not the SBVH the reference builds by default");
        std::vector<const SurfaceObject*> objs;            // a spatial split references an object from several leaves
        std::set<const SurfaceObject*> seen;
        for (const SurfaceObject* o : sbvh->m_objLists) if (seen.insert(o).second) objs.push_back(o);
        return objs;
    }
    static const SurfaceObjectAggregate* nestedOf(const SurfaceObject* o) {
        auto tso = dynamic_cast<const TransformedSurfaceObject*>(o);
        return tso ? dynamic_cast<const SurfaceObjectAggregate*>(tso->m_surfObj) : nullptr;
    }
    static bool holdsInstances(const SurfaceObjectAggregate* ag) {
        for (const SurfaceObject* o : objectsOf(ag)) if (dynamic_cast<const TransformedSurfaceObject*>(o)) return true;
        return false;
    }
    template <typename T, typename... Args> T* own(Args&&... args) {
        T* p = new T(std::forward<Args>(args)...);
        expandOwned.emplace_back(p);
        return p;
    }
    // what the parent of `tso` sees once tso's aggregate holds no instances any more
    bool expanded = false;
    void place(const TransformedSurfaceObject* tso, std::vector<SurfaceObject*>* out) {
        const SurfaceObjectAggregate* ag = nestedOf(tso);
        if (!ag) {
            // an animated node with a single object below it is a TransformedSurfaceObject over that object itself
            // (nodes.cpp:131-134): give it the one-object aggregate the device's instance record needs
            auto it = wrapped.find(tso->m_surfObj);
            if (it == wrapped.end()) {
                std::vector<SurfaceObject*> one{const_cast<SurfaceObject*>(tso->m_surfObj)};
                it = wrapped.emplace(tso->m_surfObj, own<SurfaceObjectAggregate>(one)).first;
            }
            ag = it->second;
            tso = own<TransformedSurfaceObject>(ag, tso->m_transform);
            expanded = true;
        }
        if (!holdsInstances(ag)) { out->push_back(const_cast<TransformedSurfaceObject*>(tso)); return; }
        expanded = true;
        auto it = ownTriangles.find(ag);
        if (it == ownTriangles.end()) {
            std::vector<SurfaceObject*> singles;
            for (const SurfaceObject* o : objectsOf(ag))
                if (!dynamic_cast<const TransformedSurfaceObject*>(o)) singles.push_back(const_cast<SurfaceObject*>(o));
            it = ownTriangles.emplace(ag, singles.empty() ? nullptr : own<SurfaceObjectAggregate>(singles)).first;
        }
        if (it->second) out->push_back(own<TransformedSurfaceObject>(it->second, tso->m_transform));
        for (const SurfaceObject* o : objectsOf(ag)) {
            auto inner = dynamic_cast<const TransformedSurfaceObject*>(o);
            if (!inner) continue;
            const Transform* composed = ChainedTransform(tso->m_transform, inner->m_transform).reduce(expandMem);
            place(own<TransformedSurfaceObject>(inner->m_surfObj, composed), out);
        }
    }
    const SurfaceObjectAggregate* expandNesting(const SurfaceObjectAggregate* top) {
        std::vector<SurfaceObject*> objs;
        for (const SurfaceObject* o : objectsOf(top)) {
            if (auto tso = dynamic_cast<const TransformedSurfaceObject*>(o)) place(tso, &objs);
            else objs.push_back(const_cast<SurfaceObject*>(o));
        }
        return expanded ? own<SurfaceObjectAggregate>(objs) : top;       // untouched scenes keep the reference's own top-level tree
    }

    // ---- spectra (BasicTypes/SpectrumTypes.h:70-346) ----
    uint32_t spectrum(const InputSpectrum* s) {
        auto it = spectrumIds.find(s);
        if (it != spectrumIds.end()) return it->second;
        SlrGpuSpectrum g;
        memset(&g, 0, sizeof(g));
        if (auto r = dynamic_cast<const RegularContinuousSpectrum*>(s)) {
            g.kind = SLRGPU_SPECTRUM_REGULAR; g.data_offset = (uint32_t)spectrumData.size(); g.num_samples = r->numSamples;
            g.p0 = r->minLambda; g.p1 = r->maxLambda;
            spectrumData.insert(spectrumData.end(), r->values, r->values + r->numSamples);
        } else if (auto ir = dynamic_cast<const IrregularContinuousSpectrum*>(s)) {
            g.kind = SLRGPU_SPECTRUM_IRREGULAR; g.data_offset = (uint32_t)spectrumData.size(); g.num_samples = ir->numSamples;
            spectrumData.insert(spectrumData.end(), ir->lambdas, ir->lambdas + ir->numSamples);
            spectrumData.insert(spectrumData.end(), ir->values, ir->values + ir->numSamples);
        } else if (auto u = dynamic_cast<const UpsampledContinuousSpectrum*>(s)) {
            g.kind = SLRGPU_SPECTRUM_UPSAMPLED; g.p0 = u->u; g.p1 = u->v; g.p2 = u->scale;
        } else unsupported("unknown InputSpectrum class");
        spectra.push_back(g);
        return spectrumIds[s] = (uint32_t)spectra.size() - 1;
    }

    // ---- textures (Textures/*.h, Core/textures.h:16-50) ----
    void mapping2D(const Texture2DMapping* m, SlrGpuTexture* t) {
        t->mapping = SLRGPU_MAP_TEXCOORD; t->map_scale[0] = t->map_scale[1] = 1.0f;
        if (auto os = dynamic_cast<const OffsetAndScale2DMapping*>(m)) {
            t->mapping = SLRGPU_MAP_OFFSET_SCALE_2D;
            t->map_offset[0] = os->m_offsetX; t->map_offset[1] = os->m_offsetY; t->map_scale[0] = os->m_scaleX; t->map_scale[1] = os->m_scaleY;
        }
    }
    void mapping3D(const Texture3DMapping* m, SlrGpuTexture* t) {
        t->mapping = dynamic_cast<const WorldPosition3DMapping*>(m) ? SLRGPU_MAP_WORLD_POS : SLRGPU_MAP_TEXCOORD;
        t->map_scale[0] = t->map_scale[1] = 1.0f;
    }
    uint32_t addTexture(const void* key, const SlrGpuTexture &t) { textures.push_back(t); return textureIds[key] = (uint32_t)textures.size() - 1; }
    // TiledImage2D (Core/Image.h:77-353): texels un-tiled into rows; ColorFormat's values are SlrGpuImageFormat's
    uint32_t image(const TiledImage2D* img) {
        auto it = imageIds.find(img);
        if (it != imageIds.end()) return it->second;
        SlrGpuImage g;
        memset(&g, 0, sizeof(g));
        g.format = (uint32_t)img->format(); g.width = img->width(); g.height = img->height();
        g.spectrum_type = (uint32_t)img->spectrumType();
        const size_t texel = sizesOfColorFormats[g.format];
        while (imageData.size() % 16) imageData.push_back(0);
        g.data_offset = imageData.size();
        imageData.resize(imageData.size() + texel * g.width * g.height);
        uint8_t* dst = imageData.data() + g.data_offset;
        for (uint32_t y = 0; y < g.height; ++y)
            for (uint32_t x = 0; x < g.width; ++x, dst += texel) memcpy(dst, img->getInternal(x, y), texel);
        images.push_back(g);
        return imageIds[img] = (uint32_t)images.size() - 1;
    }
    uint32_t spectrumTexture(const SpectrumTexture* tex) {
        auto it = textureIds.find(tex);
        if (it != textureIds.end()) return it->second;
        SlrGpuTexture t;
        memset(&t, 0, sizeof(t));
        t.i0 = t.i1 = SLRGPU_INVALID_ID;
        if (auto c = dynamic_cast<const ConstantSpectrumTexture*>(tex)) { t.kind = SLRGPU_TEX_CONSTANT_SPECTRUM; t.i0 = spectrum(c->m_value); t.map_scale[0] = t.map_scale[1] = 1.0f; }
        else if (auto cb = dynamic_cast<const CheckerBoardSpectrumTexture*>(tex)) {
            t.kind = SLRGPU_TEX_CHECKER_SPECTRUM; mapping2D(cb->m_mapping, &t); t.i0 = spectrum(cb->m_values[0]); t.i1 = spectrum(cb->m_values[1]);
        } else if (auto v = dynamic_cast<const VoronoiSpectrumTexture*>(tex)) {
            t.kind = SLRGPU_TEX_VORONOI_SPECTRUM; mapping3D(v->m_mapping, &t); t.f0 = v->m_scale; t.f1 = v->m_brightness;
        } else if (auto im = dynamic_cast<const ImageSpectrumTexture*>(tex)) {
            t.kind = SLRGPU_TEX_IMAGE_SPECTRUM; mapping2D(im->m_mapping, &t); t.i0 = image(im->m_data);
        } else unsupported("unknown SpectrumTexture class");
        return addTexture(tex, t);
    }
    uint32_t floatTexture(const FloatTexture* tex) {
        auto it = textureIds.find(tex);
        if (it != textureIds.end()) return it->second;
        SlrGpuTexture t;
        memset(&t, 0, sizeof(t));
        t.i0 = 0; t.i1 = SLRGPU_INVALID_ID; t.map_scale[0] = t.map_scale[1] = 1.0f;
        if (auto c = dynamic_cast<const ConstantFloatTexture*>(tex)) { t.kind = SLRGPU_TEX_CONSTANT_FLOAT; t.f0 = c->m_value; }
        else if (auto cb = dynamic_cast<const CheckerBoardFloatTexture*>(tex)) {
            t.kind = SLRGPU_TEX_CHECKER_FLOAT; mapping2D(cb->m_mapping, &t); t.f0 = cb->m_values[0]; t.f1 = cb->m_values[1];
        } else if (auto v = dynamic_cast<const VoronoiFloatTexture*>(tex)) {
            t.kind = SLRGPU_TEX_VORONOI_FLOAT; mapping3D(v->m_mapping, &t); t.f0 = v->m_scale; t.f1 = v->m_valueScale; t.i0 = v->m_flat ? 1 : 0;
        } else if (auto im = dynamic_cast<const ImageFloatTexture*>(tex)) {
            t.kind = SLRGPU_TEX_IMAGE_FLOAT; mapping2D(im->m_mapping, &t); t.i0 = image(im->m_data);
        } else unsupported("unknown FloatTexture class");
        return addTexture(tex, t);
    }
    uint32_t normalTexture(const Normal3DTexture* tex) {
        auto it = textureIds.find(tex);
        if (it != textureIds.end()) return it->second;
        SlrGpuTexture t;
        memset(&t, 0, sizeof(t));
        t.i0 = 0; t.i1 = SLRGPU_INVALID_ID; t.map_scale[0] = t.map_scale[1] = 1.0f;
        if (auto cb = dynamic_cast<const CheckerBoardNormal3DTexture*>(tex)) {
            t.kind = SLRGPU_TEX_CHECKER_NORMAL; mapping2D(cb->m_mapping, &t); t.f0 = cb->m_stepWidth; t.i0 = cb->m_reverse ? 1 : 0;
        } else if (auto v = dynamic_cast<const VoronoiNormal3DTexture*>(tex)) {
            t.kind = SLRGPU_TEX_VORONOI_NORMAL; mapping3D(v->m_mapping, &t); t.f0 = v->m_scale; t.f1 = v->m_cosThetaMax;
        } else if (auto im = dynamic_cast<const ImageNormal3DTexture*>(tex)) {
            t.kind = SLRGPU_TEX_IMAGE_NORMAL; mapping2D(im->m_mapping, &t); t.i0 = image(im->m_data);
        } else unsupported("unknown Normal3DTexture class");
        return addTexture(tex, t);
    }
    uint32_t alphaG(const SVMicrofacetDistribution* D) {
        auto ggx = dynamic_cast<const SVGGX*>(D);
        if (!ggx) unsupported("unknown microfacet distribution");
        return floatTexture(ggx->m_alpha_g);
    }

    // ---- materials (SurfaceMaterials/*.h, Core/surface_material.h:55-70) ----
    uint32_t addMaterial(const void* key, const SlrGpuMaterial &m) { materials.push_back(m); return materialIds[key] = (uint32_t)materials.size() - 1; }
    static SlrGpuMaterial blank(uint32_t kind) {
        SlrGpuMaterial m;
        memset(&m, 0, sizeof(m));
        m.kind = kind;
        for (uint32_t &t : m.tex) t = SLRGPU_INVALID_ID;
        m.sub[0] = m.sub[1] = SLRGPU_INVALID_ID;
        return m;
    }
    uint32_t emitter(const EmitterSurfaceProperty* e) {
        auto it = materialIds.find(e);
        if (it != materialIds.end()) return it->second;
        if (auto ibl = dynamic_cast<const IBLEmission*>(e)) {        // SurfaceMaterials/IBLEmission.cpp:15-17: pi * coeffM * scale
            SlrGpuMaterial m = blank(SLRGPU_MAT_IBL_EMISSION);
            m.tex[0] = spectrumTexture(ibl->m_coeffM);
            m.f0 = ibl->m_scale;
            return addMaterial(e, m);
        }
        auto d = dynamic_cast<const DiffuseEmission*>(e);
        if (!d) unsupported("unknown EmitterSurfaceProperty class");
        SlrGpuMaterial m = blank(SLRGPU_MAT_DIFFUSE_EMISSION);
        m.tex[0] = spectrumTexture(d->m_emittance);
        return addMaterial(e, m);
    }
    uint32_t material(const SurfaceMaterial* mat) {
        auto it = materialIds.find(mat);
        if (it != materialIds.end()) return it->second;
        SlrGpuMaterial m = blank(0);
        if (auto d = dynamic_cast<const DiffuseReflection*>(mat)) {
            m.kind = SLRGPU_MAT_DIFFUSE; m.tex[0] = spectrumTexture(d->m_reflectance); if (d->m_sigma) m.tex[1] = floatTexture(d->m_sigma);
        } else if (auto sr = dynamic_cast<const SpecularReflection*>(mat)) {
            m.kind = SLRGPU_MAT_SPECULAR_REFLECTION; m.tex[0] = spectrumTexture(sr->m_coeffR); m.tex[1] = spectrumTexture(sr->m_eta); m.tex[2] = spectrumTexture(sr->m_k);
        } else if (auto ss = dynamic_cast<const SpecularScattering*>(mat)) {
            m.kind = SLRGPU_MAT_SPECULAR_SCATTERING; m.tex[0] = spectrumTexture(ss->m_coeff); m.tex[1] = spectrumTexture(ss->m_etaExt); m.tex[2] = spectrumTexture(ss->m_etaInt);
        } else if (auto w = dynamic_cast<const ModifiedWardDurReflection*>(mat)) {
            m.kind = SLRGPU_MAT_WARD_DUR; m.tex[0] = spectrumTexture(w->m_reflectance); m.tex[1] = floatTexture(w->m_anisoX); m.tex[2] = floatTexture(w->m_anisoY);
        } else if (auto as = dynamic_cast<const AshikhminShirleyReflection*>(mat)) {
            m.kind = SLRGPU_MAT_ASHIKHMIN_SHIRLEY; m.tex[0] = spectrumTexture(as->m_Rs); m.tex[1] = spectrumTexture(as->m_Rd);
            m.tex[2] = floatTexture(as->m_nu); m.tex[3] = floatTexture(as->m_nv);
        } else if (auto mr = dynamic_cast<const MicrofacetReflection*>(mat)) {
            m.kind = SLRGPU_MAT_MICROFACET_REFLECTION; m.tex[0] = spectrumTexture(mr->m_eta); m.tex[1] = spectrumTexture(mr->m_k); m.tex[2] = alphaG(mr->m_D);
        } else if (auto ms = dynamic_cast<const MicrofacetScattering*>(mat)) {
            m.kind = SLRGPU_MAT_MICROFACET_SCATTERING; m.tex[0] = spectrumTexture(ms->m_etaExt); m.tex[1] = spectrumTexture(ms->m_etaInt); m.tex[2] = alphaG(ms->m_D);
        } else if (auto inv = dynamic_cast<const InverseSurfaceMaterial*>(mat)) {
            m.kind = SLRGPU_MAT_INVERSE; m.sub[0] = material(inv->m_baseMat);
        } else if (auto sum = dynamic_cast<const SummedSurfaceMaterial*>(mat)) {
            m.kind = SLRGPU_MAT_SUMMED; m.sub[0] = material(sum->m_mat0); m.sub[1] = material(sum->m_mat1);
        } else if (auto mix = dynamic_cast<const MixedSurfaceMaterial*>(mat)) {
            m.kind = SLRGPU_MAT_MIXED; m.sub[0] = material(mix->m_mat0); m.sub[1] = material(mix->m_mat1); m.tex[0] = floatTexture(mix->m_factor);
        } else if (auto em = dynamic_cast<const EmitterSurfaceMaterial*>(mat)) {
            m.kind = SLRGPU_MAT_EMITTER; if (em->m_mat) m.sub[0] = material(em->m_mat); m.sub[1] = emitter(em->m_emit);
        } else unsupported("unknown SurfaceMaterial class");
        return addMaterial(mat, m);
    }

    // ---- geometry ----
    uint32_t vertex(const Vertex* v) {
        auto it = vertexIds.find(v);
        if (it != vertexIds.end()) return it->second;
        SlrGpuVertex g;
        g.position[0] = v->position.x; g.position[1] = v->position.y; g.position[2] = v->position.z; g.u = v->texCoord.u;
        g.normal[0] = v->normal.x; g.normal[1] = v->normal.y; g.normal[2] = v->normal.z; g.v = v->texCoord.v;
        g.tangent[0] = v->tangent.x; g.tangent[1] = v->tangent.y; g.tangent[2] = v->tangent.z; g.pad = 0.0f;
        vertices.push_back(g);
        return vertexIds[v] = (uint32_t)vertices.size() - 1;
    }
    uint32_t triangle(const SingleSurfaceObject* obj) {
        auto it = triangleIds.find(obj);
        if (it != triangleIds.end()) return it->second;
        auto tri = dynamic_cast<const Triangle*>(obj->m_surface);
        if (!tri) unsupported("only triangle surfaces are exported");
        SlrGpuTriangle t;
        memset(&t, 0, sizeof(t));
        for (int k = 0; k < 3; ++k) t.v[k] = vertex(tri->m_v[k]);
        t.material = material(obj->m_material);
        auto bump = dynamic_cast<const BumpSingleSurfaceObject*>(obj);
        t.normal_map = bump && bump->m_normalMap ? normalTexture(bump->m_normalMap) : SLRGPU_INVALID_ID;
        t.alpha_map = tri->m_alphaTex ? floatTexture(tri->m_alphaTex) : SLRGPU_INVALID_ID;
        t.light_index = SLRGPU_INVALID_ID;
        triangles.push_back(t);
        return triangleIds[obj] = (uint32_t)triangles.size() - 1;
    }

    // One SurfaceObjectAggregate (Core/SurfaceObject.cpp:226-253): its SBVH collapsed by the reference's own
    // QBVH(const SBVH&) (Accelerator/QBVH.h:253-285), nodes and leaf references appended to the scene-wide arrays with
    // global indices, its light list and selection distribution copied from m_lightList / m_lightDist1D.
    const AggregateInfo &aggregate(const SurfaceObjectAggregate* ag, int depth) {
        auto it = aggregates.find(ag);
        if (it != aggregates.end()) return it->second;
        auto sbvh = dynamic_cast<const SBVH*>(ag->m_accelerator);
        if (!sbvh) unsupported("the aggregate's accelerator is not the SBVH the reference builds by default");
        QBVH qbvh(*sbvh);
        AggregateInfo info = {(uint32_t)nodes.size(), SLRGPU_INVALID_ID, 0u, 0.0f, false};
        const uint32_t leafBase = (uint32_t)leaves.size();
        // reserve this aggregate's ranges first: nested aggregates append behind them
        nodes.resize(nodes.size() + qbvh.m_nodes.size());
        leaves.resize(leaves.size() + qbvh.m_objLists.size());
        static_assert(sizeof(QBVH::Node) == sizeof(SlrGpuBvhNode), "QBVH::Node is the 128-byte node of slrgpu.h");
        for (size_t i = 0; i < qbvh.m_nodes.size(); ++i) {
            SlrGpuBvhNode n;
            memcpy(&n, &qbvh.m_nodes[i], sizeof(n));
            n.top_axis = (uint8_t)qbvh.m_nodes[i].topAxis; n.left_axis = (uint8_t)qbvh.m_nodes[i].leftAxis; n.right_axis = (uint8_t)qbvh.m_nodes[i].rightAxis;
            n.pad0 = 0; n.pad[0] = n.pad[1] = n.pad[2] = 0;
            for (int l = 0; l < 4; ++l) {
                const uint32_t c = qbvh.m_nodes[i].children[l].asUInt;
                if (c == 0xFFFFFFFFu) continue;
                const uint32_t idx = (c & 0x07FFFFFFu) + ((c >> 31) ? leafBase : info.nodeBase);
                n.child[l] = (c & 0xF8000000u) | idx;
            }
            nodes[info.nodeBase + i] = n;
        }
        for (size_t i = 0; i < qbvh.m_objLists.size(); ++i) {
            const SurfaceObject* o = qbvh.m_objLists[i];
            SlrGpuLeafRecord r;
            memset(&r, 0, sizeof(r));
            uint32_t id;
            if (auto tso = dynamic_cast<const TransformedSurfaceObject*>(o)) {
                if (depth > 0) unsupported("instancing nested deeper than one level");
                id = 0x80000000u | instance(tso, depth);
                info.hasInstances = true;
            } else if (auto single = dynamic_cast<const SingleSurfaceObject*>(o)) {
                if (dynamic_cast<const InfiniteSphereSurfaceObject*>(o)) unsupported("environment sphere inside an aggregate");
                id = triangle(single);
                const Triangle* tri = static_cast<const Triangle*>(single->m_surface);
                const Point3D &p0 = tri->m_v[0]->position;
                const Vector3D e1 = tri->m_v[1]->position - p0, e2 = tri->m_v[2]->position - p0;      // TriangleMesh.cpp:136-137
                r.a[0] = p0.x; r.a[1] = p0.y; r.a[2] = p0.z;
                r.b[0] = e1.x; r.b[1] = e1.y; r.b[2] = e1.z;
                r.c[0] = e2.x; r.c[1] = e2.y; r.c[2] = e2.z;
                const uint32_t flags = tri->m_alphaTex ? SLRGPU_LEAF_FLAG_ALPHA_TEST : 0u;
                memcpy(&r.b[3], &flags, 4);
            } else unsupported("unknown SurfaceObject class in an aggregate");
            memcpy(&r.a[3], &id, 4);
            leaves[leafBase + i] = r;
        }
        // light list + RegularConstantDiscrete1D (SurfaceObject.cpp:232-252, distributions.cpp:81-119)
        const RegularConstantDiscrete1D* dist = ag->m_lightDist1D;
        const uint32_t numLights = dist ? dist->m_numValues : 0;
        std::vector<SlrGpuLight> mine(numLights);
        for (uint32_t i = 0; i < numLights; ++i) {
            const SurfaceObject* o = ag->m_lightList[i];
            SlrGpuLight l;
            memset(&l, 0, sizeof(l));
            l.importance = o->importance();
            l.pmf = dist->m_PMF[i]; l.cdf_lo = dist->m_CDF[i]; l.cdf_hi = dist->m_CDF[i + 1];
            if (auto tso = dynamic_cast<const TransformedSurfaceObject*>(o)) {
                const uint32_t inst = instance(tso, depth);
                l.object = 0x80000000u | inst;
                instances[inst].light_index = i;
                instances[inst].light_importance = l.importance;
            } else {
                const uint32_t tri = triangle(static_cast<const SingleSurfaceObject*>(o));
                l.object = tri;
                triangles[tri].light_index = i;
            }
            mine[i] = l;
        }
        info.lightBase = numLights ? (uint32_t)lights.size() : SLRGPU_INVALID_ID;
        info.numLights = numLights;
        info.importance = dist ? dist->m_integral : 0.0f;
        lights.insert(lights.end(), mine.begin(), mine.end());
        return aggregates[ag] = info;
    }
    // A node's transform after the scene graph's ChainedTransform::reduce (Transform.cpp:74-152): a StaticTransform, or ONE
    // AnimatedTransform with the static transforms around it folded into its key frames (createByMulLeft / Right).
    // mat / matInv receive the (begin) key frame; returns SlrGpuInstance::motion / SlrGpuSceneDesc::camera_motion:
    // 0 = static, else 1 + index of the SlrGpuMotion that carries the end key frame and the reference's own decomposition
    // (AnimatedTransform's T / R / S members, Transform.h:89-123), so the device interpolates what sample() would.
    uint32_t keyFrames(const Transform* tf, float* mat, float* matInv) {
        if (auto st = dynamic_cast<const StaticTransform*>(tf)) {
            memcpy(mat, &st->mat, 64);          // Matrix4x4 is four column vectors: column-major like SlrGpuInstance
            memcpy(matInv, &st->matInv, 64);
            return 0u;
        }
        auto at = dynamic_cast<const AnimatedTransform*>(tf);
        if (!at) unsupported("a chain of two animated transforms (motion blur inside motion blur)");
        memcpy(mat, &at->m_tfBegin.mat, 64);
        memcpy(matInv, &at->m_tfBegin.matInv, 64);
        if (at->isStatic()) return 0u;
        SlrGpuMotion m;
        memset(&m, 0, sizeof(m));
        memcpy(m.mat_end, &at->m_tfEnd.mat, 64);
        memcpy(m.mat_end_inv, &at->m_tfEnd.matInv, 64);
        for (int k = 0; k < 3; ++k) { m.T0[k] = at->m_T[0][k]; m.T1[k] = at->m_T[1][k]; }
        m.t_begin = at->m_tBegin; m.t_end = at->m_tEnd;
        m.R0[0] = at->m_R[0].x; m.R0[1] = at->m_R[0].y; m.R0[2] = at->m_R[0].z; m.R0[3] = at->m_R[0].w;
        m.R1[0] = at->m_R[1].x; m.R1[1] = at->m_R[1].y; m.R1[2] = at->m_R[1].z; m.R1[3] = at->m_R[1].w;
        memcpy(m.S0, &at->m_S[0], 64);
        memcpy(m.S1, &at->m_S[1], 64);
        motions.push_back(m);
        return (uint32_t)motions.size();
    }
    uint32_t instance(const TransformedSurfaceObject* tso, int depth) {
        auto it = instanceIds.find(tso);
        if (it != instanceIds.end()) return it->second;
        auto nested = dynamic_cast<const SurfaceObjectAggregate*>(tso->m_surfObj);
        if (!nested) unsupported("a TransformedSurfaceObject over something other than an aggregate");
        SlrGpuInstance inst;
        memset(&inst, 0, sizeof(inst));
        inst.motion = keyFrames(tso->m_transform, inst.mat, inst.mat_inv);
        inst.light_index = SLRGPU_INVALID_ID;
        const uint32_t id = (uint32_t)instances.size();
        instances.push_back(inst);
        instanceIds[tso] = id;
        const AggregateInfo info = aggregate(nested, depth + 1);       // may append to `instances`: index, not reference
        instances[id].root_node = info.nodeBase;
        instances[id].light_base = info.lightBase;
        instances[id].num_lights = info.numLights;
        return id;
    }

public:
    SlrGpuSceneDesc desc;

    explicit GpuSceneExporter(const Scene &scene) {
        // the top-level aggregate must own node 0: export it first
        const AggregateInfo top = aggregate(expandNesting(scene.m_aggregate), 0);
        numTopLights = top.numLights;
        topLightImportance = top.importance;
        // top-level lights must come first in `lights` (slrgpu.h): aggregate() appends an aggregate's own list after its
        // nested ones, so rotate the top-level slice to the front and patch the nested bases
        if (top.numLights && top.lightBase != 0) {
            std::vector<SlrGpuLight> reordered(lights.begin() + top.lightBase, lights.begin() + top.lightBase + top.numLights);
            reordered.insert(reordered.end(), lights.begin(), lights.begin() + top.lightBase);
            reordered.insert(reordered.end(), lights.begin() + top.lightBase + top.numLights, lights.end());
            for (SlrGpuInstance &in : instances)
                if (in.light_base != SLRGPU_INVALID_ID) in.light_base += in.light_base < top.lightBase ? top.numLights : 0;
            lights.swap(reordered);
        }

        memset(&desc, 0, sizeof(desc));
        desc.struct_size = sizeof(desc);
#ifdef Use_Spectral_Representation
        desc.rgb_mode = 0;
#else
        desc.rgb_mode = 1;
#endif
        desc.bvh_nodes = nodes.data(); desc.num_bvh_nodes = (uint32_t)nodes.size();
        desc.leaf_records = leaves.data(); desc.num_leaf_records = (uint32_t)leaves.size();
        desc.instances = instances.data(); desc.num_instances = (uint32_t)instances.size();
        desc.triangles = triangles.data(); desc.num_triangles = (uint32_t)triangles.size();
        desc.vertices = vertices.data(); desc.num_vertices = (uint32_t)vertices.size();
        desc.materials = materials.data(); desc.num_materials = (uint32_t)materials.size();
        desc.textures = textures.data(); desc.num_textures = (uint32_t)textures.size();
        desc.spectra = spectra.data(); desc.num_spectra = (uint32_t)spectra.size();
        desc.spectrum_data = spectrumData.data(); desc.num_spectrum_floats = (uint32_t)spectrumData.size();
        desc.lights = lights.data(); desc.num_lights = (uint32_t)lights.size();
        desc.num_top_lights = numTopLights;
        desc.top_light_importance = topLightImportance;
        desc.world_center[0] = scene.m_worldCenter.x; desc.world_center[1] = scene.m_worldCenter.y; desc.world_center[2] = scene.m_worldCenter.z;
        desc.world_radius = scene.m_worldRadius;

        desc.images = images.data(); desc.num_images = (uint32_t)images.size();
        desc.image_data = imageData.data(); desc.image_data_bytes = imageData.size();

        // the environment (InfiniteSphereSurfaceObject, SurfaceObject.cpp:140-222): its emitter and the importance map the
        // reference built for it (IBLEmission::createIBLImportanceMap), copied array by array
        if (const InfiniteSphereSurfaceObject* env = scene.m_envSphere) {
            auto em = dynamic_cast<const EmitterSurfaceMaterial*>(env->m_material);
            if (!em) unsupported("the environment sphere's material is not an EmitterSurfaceMaterial");
            const RegularConstantContinuous2D* dist = env->m_dist;
            const uint32_t H = dist->m_num1DDists, W = dist->m_1DDists[0].m_numValues;
            envRowPdf.resize((size_t)W * H); envRowCdf.resize((size_t)(W + 1) * H); envRowIntegral.resize(H);
            envMarginalPdf.resize(H); envMarginalCdf.resize(H + 1);
            for (uint32_t y = 0; y < H; ++y) {
                const RegularConstantContinuous1D &row = dist->m_1DDists[y];
                memcpy(&envRowPdf[(size_t)y * W], row.m_PDF, sizeof(float) * W);
                memcpy(&envRowCdf[(size_t)y * (W + 1)], row.m_CDF, sizeof(float) * (W + 1));
                envRowIntegral[y] = row.m_integral;
            }
            memcpy(envMarginalPdf.data(), dist->m_top1DDist->m_PDF, sizeof(float) * H);
            memcpy(envMarginalCdf.data(), dist->m_top1DDist->m_CDF, sizeof(float) * (H + 1));
            desc.environment.present = 1;
            desc.environment.material = emitter(em->m_emit);
            desc.environment.map_width = W; desc.environment.map_height = H;
            desc.environment.row_pdf = envRowPdf.data(); desc.environment.row_cdf = envRowCdf.data();
            desc.environment.row_integral = envRowIntegral.data();
            desc.environment.marginal_pdf = envMarginalPdf.data(); desc.environment.marginal_cdf = envMarginalCdf.data();
            desc.environment.marginal_integral = dist->m_top1DDist->m_integral;
            // the tables may have grown: the pointers taken above are refreshed below
        }
        desc.materials = materials.data(); desc.num_materials = (uint32_t)materials.size();
        desc.textures = textures.data(); desc.num_textures = (uint32_t)textures.size();
        desc.spectra = spectra.data(); desc.num_spectra = (uint32_t)spectra.size();
        desc.spectrum_data = spectrumData.data(); desc.num_spectrum_floats = (uint32_t)spectrumData.size();
        desc.images = images.data(); desc.num_images = (uint32_t)images.size();
        desc.image_data = imageData.data(); desc.image_data_bytes = imageData.size();

        auto cam = dynamic_cast<const PerspectiveCamera*>(scene.getCamera());
        if (!cam) unsupported("only PerspectiveCamera is exported");
        desc.camera_motion = keyFrames(cam->m_transform, desc.camera.mat, desc.camera.mat_inv);
        desc.motions = motions.empty() ? nullptr : motions.data(); desc.num_motions = (uint32_t)motions.size();
        desc.camera.sensitivity = cam->getSensor()->m_sensitivity;
        desc.camera.aspect = cam->m_aspect; desc.camera.fov_y = cam->m_fovY; desc.camera.lens_radius = cam->m_lensRadius;
        desc.camera.img_plane_dist = cam->m_imgPlaneDistance; desc.camera.obj_plane_dist = cam->m_objPlaneDistance;

        // spectral constant tables straight from the reference's statics (Spectrum.h:205-571, SpectrumTypes.h:746-795)
        const uint8_t* grid = reinterpret_cast<const uint8_t*>(Upsampling::spectrum_grid);
        gridFloats.resize(sizeof(Upsampling::spectrum_grid));
        for (size_t i = 0; i < gridFloats.size(); ++i) gridFloats[i] = (float)grid[i];
        desc.spectral.upsample_grid = gridFloats.data(); desc.spectral.upsample_grid_floats = (uint32_t)gridFloats.size();
        desc.spectral.upsample_points = reinterpret_cast<const float*>(Upsampling::spectrum_data_points);
        desc.spectral.upsample_points_floats = (uint32_t)(sizeof(Upsampling::spectrum_data_points) / 4);
#ifdef Use_Spectral_Representation
        desc.spectral.xbar_16 = DiscretizedSpectrum::xbar.get(); desc.spectral.ybar_16 = DiscretizedSpectrum::ybar.get();
        desc.spectral.zbar_16 = DiscretizedSpectrum::zbar.get(); desc.spectral.integral_cmf = DiscretizedSpectrum::integralCMF;
#endif
    }
};

}  // namespace

void GPUPathTracingRenderer::render(const Scene &scene, const RenderSettings &settings) const {
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t W = (uint32_t)settings.getInt(RenderSettingItem::ImageWidth);
    const uint32_t H = (uint32_t)settings.getInt(RenderSettingItem::ImageHeight);
    ImageSensor* sensor = scene.getCamera()->getSensor();
    sensor->init(W, H);

    GpuSceneExporter exporter(scene);
    // SLRGPU_DROPIN_DUMP=<file>: the geometry and light tables as they go to slrgpu_scene_create (a debugging aid; tests/test_dropin.py
    // walks them on the CPU and compares the hits with the reference's own Scene::intersect on the same file)
    if (const char* dumpPath = getenv("SLRGPU_DROPIN_DUMP")) {
        if (FILE* f = fopen(dumpPath, "wb")) {
            const SlrGpuSceneDesc &d = exporter.desc;
            const uint32_t head[6] = {0x44524F50u, d.num_bvh_nodes, d.num_leaf_records, d.num_instances, d.num_lights, d.num_top_lights};
            fwrite(head, 4, 6, f);
            fwrite(d.bvh_nodes, sizeof(SlrGpuBvhNode), d.num_bvh_nodes, f);
            fwrite(d.leaf_records, sizeof(SlrGpuLeafRecord), d.num_leaf_records, f);
            fwrite(d.instances, sizeof(SlrGpuInstance), d.num_instances, f);
            fwrite(d.lights, sizeof(SlrGpuLight), d.num_lights, f);
            fclose(f);
        }
    }
    // without a CUDA device slrgpu_scene_create fails below with SLRGPU_ERR_NO_DEVICE (after validating the tables): there is
    // no CPU fallback behind this renderer -- a host without a GPU keeps PathTracingRenderer
    const int visible = slrgpu_device_count();
    const int first = device >= 0 ? device : 0;
    const int count = device >= 0 ? 1 : (int)std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)visible, m_samplesPerPixel));
    std::vector<SlrGpuScene*> replicas;
    struct DestroyAll { std::vector<SlrGpuScene*> &r; ~DestroyAll() { for (SlrGpuScene* g : r) slrgpu_scene_destroy(g); } } destroyAll{replicas};
    for (int g = 0; g < count; ++g) {
        SlrGpuScene* gpu = nullptr;
        if (slrgpu_scene_create(&exporter.desc, first + g, &gpu) != SLRGPU_OK)
            throw std::runtime_error(std::string("slrgpu_scene_create: ") + slrgpu_last_error());
        replicas.push_back(gpu);
    }
    const uint32_t channels = slrgpu_scene_channels(replicas[0]);

    SlrGpuRenderParams p;
    memset(&p, 0, sizeof(p));
    p.struct_size = sizeof(p);
    p.width = W; p.height = H;
    p.time_start = settings.getFloat(RenderSettingItem::TimeStart);
    p.time_end = settings.getFloat(RenderSettingItem::TimeEnd);
    p.rng_seed = settings.getInt(RenderSettingItem::RNGSeed);
    if (bidirectional) p.flags |= SLRGPU_RENDER_BPT;
    const float brightness = settings.getFloat(RenderSettingItem::Brightness);
    std::vector<float> pass((size_t)W * H * channels);

    // the reference's export cadence: an image after 1, 2, 4, ... samples (PathTracingRenderer.cpp:63-65,83-94); every
    // segment is one GPU call whose sums are ADDED into the sensor's storage, so pixel(x, y) / saveImage see exactly what
    // PathTracingRenderer would have left there
    uint32_t begin = 0, exportAt = exportProgressiveImages ? 1 : m_samplesPerPixel, imgIdx = 0;
    while (begin < m_samplesPerPixel) {
        const uint32_t end = std::min(exportAt, m_samplesPerPixel);
        p.spp_begin = begin; p.spp_end = end;
        SlrGpuRenderStats st;
        const int rc = replicas.size() > 1 ? slrgpu_render_multi(replicas.data(), (uint32_t)replicas.size(), &p, pass.data(), &st)
                                           : slrgpu_render(replicas[0], &p, pass.data(), &st);
        if (rc != SLRGPU_OK) throw std::runtime_error(std::string("slrgpu_render: ") + slrgpu_last_error());
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                SpectrumStorage &px = sensor->pixel(x, y);              // tiled storage, ImageSensor.cpp:97-104
                const float* src = &pass[((size_t)y * W + x) * channels];
#ifdef Use_Spectral_Representation
                for (uint32_t c = 0; c < channels; ++c) px.value.result.values[c] += src[c];
#else
                px.value.result.r += src[0]; px.value.result.g += src[1]; px.value.result.b += src[2];
#endif
            }
        if (exportProgressiveImages && end == exportAt) {
            char filename[256];
            sprintf(filename, "%03u.bmp", imgIdx);
            sensor->saveImage(filename, brightness / end);
            const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%u samples: %s, %g[s]\n", end, filename, elapsed);
            if (++imgIdx == 16) break;
            exportAt += exportAt;
        }
        begin = end;
    }
}

}  // namespace SLR
