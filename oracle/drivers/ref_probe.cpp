// Shading probe through the REFERENCE (test infrastructure only): for every probe ray the reference's own
// Scene::intersect, Intersection::getSurfacePoint, SurfacePoint::createBSDF and BSDF::sample / evaluate /
// evaluatePDF are run with caller-given random numbers, and the results are dumped for comparison with the
// GPU's slrgpu_probe_shading. Pins surface points (incl. normal maps, instances), every material / texture /
// spectrum evaluation and every BSDF model at function level (PathTracingRenderer.cpp:147-210 is exactly
// this sequence of calls).
//   ref_probe scene.txt probes.bin out.bin [bpt]
// With "bpt" the queries are the bidirectional path tracer's (BidirectionalPathTracingRenderer.cpp:184-196, 302-325):
// BSDF::sample with result->reverse, BSDF::evaluatePDF with revPDF, BSDF::evaluate with rev_fs, as a radiance query for even
// probes and an importance query (adjoint = true) for odd probes; output layout: include/slrgpu.h slrgpu_probe_shading_bpt,
// plus [58] = the largest |rev_fs - fs| of the evaluate call (the GPU side relies on it being zero).
// probes.bin: u32 n, then n x 14 f32: org[3] dir[3] wlOffset uLambda uComponent uDir0 uDir1 evalDirWorld[3]
// out.bin:    u32 n, u32 64, then n x 64 f32 (layout: include/slrgpu.h SLRGPU_PROBE_*)
#include <libSLR/Core/SurfaceObject.h>
#include <libSLR/Core/directional_distribution_functions.h>
#include <libSLR/Core/cameras.h>
#include <libSLR/Memory/ArenaAllocator.h>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include <libSLRSceneGraph/Scene.h>
#include <libSLRSceneGraph/API.hpp>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

using namespace SLR;

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_probe scene.txt probes.bin out.bin [bpt]\n"); return 2; }
    const bool bptMode = argc > 4 && std::string(argv[4]) == "bpt";
    initSpectrum();
    SLRSceneGraph::SceneRef scene = createShared<SLRSceneGraph::Scene>();
    SLRSceneGraph::RenderingContext context;
    context.width = 64; context.height = 64; context.timeStart = 0; context.timeEnd = 0; context.brightness = 1.0f; context.rngSeed = 1;
    if (!SLRSceneGraph::readScene(argv[1], scene, &context)) { fprintf(stderr, "Failed to read a scene file.\n"); return 1; }
    const Scene* raw;
    ArenaAllocator sceneMem;
    scene->build(&raw, sceneMem);

    FILE* f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 1; }
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return 1;
    std::vector<float> probes((size_t)n * 14);
    if (fread(probes.data(), 4, probes.size(), f) != probes.size()) return 1;
    fclose(f);

    std::vector<float> out((size_t)n * 64, 0.0f);
    ArenaAllocator mem;
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = &probes[(size_t)i * 14];
        float* o = &out[(size_t)i * 64];
        float selectPDF;
        WavelengthSamples wls = WavelengthSamples::createWithEqualOffsets(p[6], p[7], &selectPDF);
        Ray ray(Point3D(p[0], p[1], p[2]), Vector3D(p[3], p[4], p[5]), 0.0f);
        Intersection isect;
        if (!raw->intersect(ray, &isect)) { mem.reset(); continue; }
        SurfacePoint sp;
        isect.getSurfacePoint(&sp);
        if (sp.atInfinity) { o[0] = 2.0f; mem.reset(); continue; }
        o[0] = 1.0f;
        o[1] = isect.dist;
        if (bptMode) {
            Vector3D dirOut_sn = sp.shadingFrame.toLocal(-ray.dir);
            Normal3D gNorm_sn = sp.shadingFrame.toLocal(sp.gNormal);
            BSDF* bsdf = sp.createBSDF(wls, mem);
            BSDFQuery query(dirOut_sn, gNorm_sn, wls.selectedLambda, DirectionType::All, (i & 1u) != 0);
            BSDFQueryResult res;
            BSDFReverseInfo rev;
            rev.fs = SampledSpectrum::Zero; rev.dirPDF = 0.0f;
            res.reverse = &rev;
            SampledSpectrum fs = bsdf->sample(query, BSDFSample(p[8], p[9], p[10]), &res);
            for (int k = 0; k < 16; ++k) o[2 + k] = fs[k];
            o[18] = res.dir_sn.x; o[19] = res.dir_sn.y; o[20] = res.dir_sn.z;
            o[21] = res.dirPDF;
            o[22] = (float)res.dirType.value;
            const bool sampled = !(fs == SampledSpectrum::Zero) && res.dirPDF != 0.0f;
            for (int k = 0; k < 16; ++k) o[23 + k] = sampled ? rev.fs[k] : 0.0f;
            o[39] = sampled ? rev.dirPDF : 0.0f;
            Vector3D evalDir_sn = sp.shadingFrame.toLocal(Vector3D(p[11], p[12], p[13]));
            float revPDF = 0.0f;
            o[40] = bsdf->evaluatePDF(query, evalDir_sn, &revPDF);
            o[41] = revPDF;
            SampledSpectrum revFs;
            SampledSpectrum fe = bsdf->evaluate(query, evalDir_sn, &revFs);
            float worst = 0.0f;
            for (int k = 0; k < 16; ++k) { o[42 + k] = fe[k]; worst = std::fmax(worst, std::fabs(revFs[k] - fe[k])); }
            o[58] = worst;
            mem.reset();
            continue;
        }
        o[2] = sp.p.x; o[3] = sp.p.y; o[4] = sp.p.z;
        o[5] = sp.shadingFrame.z.x; o[6] = sp.shadingFrame.z.y; o[7] = sp.shadingFrame.z.z;
        o[8] = sp.shadingFrame.x.x; o[9] = sp.shadingFrame.x.y; o[10] = sp.shadingFrame.x.z;
        Vector3D dirOut_sn = sp.shadingFrame.toLocal(-ray.dir);
        Normal3D gNorm_sn = sp.shadingFrame.toLocal(sp.gNormal);
        BSDF* bsdf = sp.createBSDF(wls, mem);
        BSDFQuery query(dirOut_sn, gNorm_sn, wls.selectedLambda);
        o[11] = bsdf->hasNonDelta() ? 1.0f : 0.0f;
        BSDFQueryResult res;
        SampledSpectrum fs = bsdf->sample(query, BSDFSample(p[8], p[9], p[10]), &res);
        for (int k = 0; k < 16; ++k) o[12 + k] = fs[k];
        o[28] = res.dir_sn.x; o[29] = res.dir_sn.y; o[30] = res.dir_sn.z;
        o[31] = res.dirPDF;
        o[32] = (float)res.dirType.value;
        Vector3D evalDir_sn = sp.shadingFrame.toLocal(Vector3D(p[11], p[12], p[13]));
        SampledSpectrum fe = bsdf->evaluate(query, evalDir_sn);
        for (int k = 0; k < 16; ++k) o[33 + k] = fe[k];
        o[49] = bsdf->evaluatePDF(query, evalDir_sn);
        o[50] = sp.isEmitting() ? 1.0f : 0.0f;
        if (sp.isEmitting()) { SampledSpectrum Le = sp.emittance(wls); for (int k = 0; k < 13; ++k) o[51 + k] = Le[k]; }
        mem.reset();
    }
    f = fopen(argv[3], "wb");
    if (!f) { perror(argv[3]); return 1; }
    uint32_t stride = 64;
    fwrite(&n, 4, 1, f); fwrite(&stride, 4, 1, f);
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
