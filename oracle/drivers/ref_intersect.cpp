// Reference driver for hit parity and the CPU Mrays/s baseline (test infrastructure).
// Builds the geometry spec with the reference's own scene-graph classes, flattens it with the
// reference's getRenderingData, builds SurfaceObjectAggregate (SBVH) and QBVH(SBVH), traces the ray
// batch through QBVH::intersect (and SBVH::intersect for cross-checking) and writes hit records and
// the trees.
//   ref_intersect spec.bin rays.bin hits_out.bin [trees_out.bin] [threads]
// Exercised reference code: libSLRSceneGraph/nodes.cpp:110-184, TriangleMeshNode.cpp:68-112,
// libSLR/Core/SurfaceObject.cpp:226-318, Accelerator/SBVH.h, Accelerator/QBVH.h, Surface/TriangleMesh.cpp:131-178.
#include <libSLR/Core/SurfaceObject.h>
#include <libSLR/Core/Transform.h>
#include <libSLR/Accelerator/SBVH.h>
#include <libSLR/Accelerator/QBVH.h>
#include <libSLR/Memory/ArenaAllocator.h>
#include <libSLR/Surface/TriangleMesh.h>
#include <libSLR/BasicTypes/Spectrum.h>
#include <libSLR/BasicTypes/SpectrumTypes.h>
#include <libSLRSceneGraph/nodes.h>
#include <libSLRSceneGraph/TriangleMeshNode.h>
#include <libSLRSceneGraph/surface_materials.hpp>
#include <libSLRSceneGraph/textures.hpp>
#include <functional>
#include <map>
#include <thread>
#include "geom_spec.h"

using namespace SLR;

struct AggInfo { const SurfaceObjectAggregate* aggr; const SBVH* sbvh; QBVH* qbvh; };

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_intersect spec rays hits_out [trees_out] [threads]\n"); return 2; }
    initSpectrum();
    GeomSpec spec = readGeomSpec(argv[1]);
    RayFile rays = readRays(argv[2]);
    const char* treesOut = (argc > 4 && argv[4][0] && strcmp(argv[4], "-") != 0) ? argv[4] : nullptr;
    unsigned threads = argc > 5 ? (unsigned)atoi(argv[5]) : std::thread::hardware_concurrency();
    if (threads == 0) threads = 1;

    ArenaAllocator mem;
    // a dummy matte material so SingleSurfaceObject has something to point at (never evaluated here)
    SLRSceneGraph::InputSpectrumRef grey(new UpsampledContinuousSpectrum(SpectrumType::Reflectance, ColorSpace::sRGB, 0.5f, 0.5f, 0.5f));
    SLRSceneGraph::SpectrumTextureRef tex = createShared<SLRSceneGraph::ConstantSpectrumTexture>(grey);
    SLRSceneGraph::SurfaceMaterialRef mat = SLRSceneGraph::SurfaceMaterial::createMatte(tex, nullptr);

    std::vector<SLRSceneGraph::TriangleMeshNodeRef> meshNodes;
    for (const GeomMesh& m : spec.meshes) {
        SLRSceneGraph::TriangleMeshNodeRef node = createShared<SLRSceneGraph::TriangleMeshNode>();
        for (size_t v = 0; v < m.pos.size() / 3; ++v)
            node->addVertex(Vertex(Point3D(m.pos[3 * v], m.pos[3 * v + 1], m.pos[3 * v + 2]), Normal3D(0, 1, 0), Tangent3D(1, 0, 0), TexCoord2D(0, 0)));
        std::vector<SLRSceneGraph::Triangle> tris;
        for (size_t t = 0; t < m.idx.size() / 3; ++t) tris.emplace_back(m.idx[3 * t], m.idx[3 * t + 1], m.idx[3 * t + 2]);
        // the constructor prints a warning per triangle whose normal faces away; silence stdout while adding
        FILE* saved = stdout; stdout = fopen("/dev/null", "w");
        node->addTriangles(mat, nullptr, nullptr, std::move(tris));
        fclose(stdout); stdout = saved;
        meshNodes.push_back(node);
    }
    SLRSceneGraph::InternalNodeRef root = createShared<SLRSceneGraph::InternalNode>();
    root->setTransform(createShared<StaticTransform>());
    std::vector<SLRSceneGraph::NodeRef> refNodes(spec.meshes.size());
    std::vector<int64_t> primBase(spec.meshes.size(), -1);
    int64_t nextPrim = 0;
    for (const GeomPlacement& p : spec.placements) {
        float m[16]; memcpy(m, p.mat, 64);
        SLRSceneGraph::InternalNodeRef holder = createShared<SLRSceneGraph::InternalNode>();
        holder->setTransform(createShared<StaticTransform>(Matrix4x4(m)));   // array ctor is column-major
        if (p.mode == 0) {
            holder->addChildNode(meshNodes[p.mesh]);
        } else {
            if (!refNodes[p.mesh]) refNodes[p.mesh] = createShared<SLRSceneGraph::ReferenceNode>(meshNodes[p.mesh]);
            holder->addChildNode(refNodes[p.mesh]);
        }
        root->addChildNode(holder);
        if (primBase[p.mesh] < 0) { primBase[p.mesh] = nextPrim; nextPrim += (int64_t)spec.meshes[p.mesh].idx.size() / 3; }
    }

    SLRSceneGraph::RenderingData data;
    root->getRenderingData(mem, nullptr, &data);
    auto tb0 = std::chrono::steady_clock::now();
    SurfaceObjectAggregate aggr(data.surfObjs);
    auto tb1 = std::chrono::steady_clock::now();

    // object -> id maps
    std::map<const SurfaceObject*, uint32_t> objId;
    for (size_t mi = 0; mi < meshNodes.size(); ++mi) {
        if (primBase[mi] < 0) continue;
        for (size_t k = 0; k < meshNodes[mi]->m_numRefinedObjs; ++k)
            objId[meshNodes[mi]->m_refinedObjs[k]] = (uint32_t)(primBase[mi] + k);
    }
    uint32_t nextInst = 0;
    std::vector<AggInfo> aggs;
    aggs.push_back(AggInfo{&aggr, (const SBVH*)aggr.m_accelerator, nullptr});
    for (SurfaceObject* o : data.surfObjs) {
        if (TransformedSurfaceObject* t = dynamic_cast<TransformedSurfaceObject*>(o)) {
            objId[t] = 0x80000000u | nextInst++;
            if (const SurfaceObjectAggregate* na = dynamic_cast<const SurfaceObjectAggregate*>(t->m_surfObj)) {
                bool seen = false;
                for (auto& a : aggs) seen |= a.aggr == na;
                if (!seen) aggs.push_back(AggInfo{na, (const SBVH*)na->m_accelerator, nullptr});
            }
        }
    }
    // swap every aggregate's accelerator for QBVH(SBVH) AFTER all SBVHs (and their costs) exist
    auto tq0 = std::chrono::steady_clock::now();
    for (auto& a : aggs) a.qbvh = new QBVH(*a.sbvh);
    auto tq1 = std::chrono::steady_clock::now();

    const uint64_t n = rays.n;
    std::vector<uint32_t> prim(n), inst(n), primS(n), instS(n);
    std::vector<float> t(n), u(n), v(n), tS(n), uS(n), vS(n);
    auto trace = [&](bool useQ, unsigned nth, std::vector<uint32_t>& outPrim, bool writeAll) {
        uint32_t* const instOut = useQ ? inst.data() : instS.data();
        float* const tOut = useQ ? t.data() : tS.data();
        float* const uOut = useQ ? u.data() : uS.data();
        float* const vOut = useQ ? v.data() : vS.data();
        for (auto& a : aggs) const_cast<SurfaceObjectAggregate*>(a.aggr)->m_accelerator = useQ ? (Accelerator*)a.qbvh : (Accelerator*)a.sbvh;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (unsigned th = 0; th < nth; ++th) {
            pool.emplace_back([&, th]() {
                const uint64_t lo = n * th / nth, hi = n * (th + 1) / nth;
                for (uint64_t i = lo; i < hi; ++i) {
                    Ray ray(Point3D(rays.c[0][i], rays.c[1][i], rays.c[2][i]), Vector3D(rays.c[3][i], rays.c[4][i], rays.c[5][i]), 0.0f, rays.c[6][i], rays.c[7][i]);
                    Intersection isect;
                    uint32_t p = 0xFFFFFFFFu, in = 0xFFFFFFFFu;
                    if (aggr.intersect(ray, &isect)) {
                        const auto& stk = isect.obj.c;
                        uint32_t top = objId.at(stk.back());
                        if (top & 0x80000000u) { in = top & 0x7FFFFFFFu; p = objId.at(stk[stk.size() - 2]); }
                        else p = top;
                    }
                    outPrim[i] = p;
                    if (writeAll) { instOut[i] = in; tOut[i] = isect.dist; uOut[i] = p == 0xFFFFFFFFu ? 0.0f : isect.u; vOut[i] = p == 0xFFFFFFFFu ? 0.0f : isect.v; }
                }
            });
        }
        for (auto& th : pool) th.join();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    double secQ1 = trace(true, 1, prim, true);
    double secQN = threads > 1 ? trace(true, threads, prim, true) : secQ1;
    double secS1 = trace(false, 1, primS, true);       // the shipped default accelerator: its answers are dumped too
    uint64_t mism = 0, nhit = 0;
    for (uint64_t i = 0; i < n; ++i) { mism += prim[i] != primS[i]; nhit += prim[i] != 0xFFFFFFFFu; }

    FILE* f = fopen(argv[3], "wb");
    fwrite(&n, 8, 1, f);
    fwrite(prim.data(), 4, n, f); fwrite(inst.data(), 4, n, f);
    fwrite(t.data(), 4, n, f); fwrite(u.data(), 4, n, f); fwrite(v.data(), 4, n, f);
    // then the same five arrays from SBVH::intersect
    fwrite(primS.data(), 4, n, f); fwrite(instS.data(), 4, n, f);
    fwrite(tS.data(), 4, n, f); fwrite(uS.data(), 4, n, f); fwrite(vS.data(), 4, n, f);
    fclose(f);

    if (treesOut) {
        // per aggregate: u32 numNodes, nodes (128 B each, child indices local), u32 numRefs, refs as object ids
        FILE* g = fopen(treesOut, "wb");
        uint32_t na = (uint32_t)aggs.size();
        fwrite(&na, 4, 1, g);
        for (auto& a : aggs) {
            uint32_t nn = (uint32_t)a.qbvh->m_nodes.size(), nr = (uint32_t)a.qbvh->m_objLists.size();
            fwrite(&nn, 4, 1, g);
            fwrite(a.qbvh->m_nodes.data(), 128, nn, g);
            fwrite(&nr, 4, 1, g);
            for (const SurfaceObject* o : a.qbvh->m_objLists) { uint32_t id = objId.at(o); fwrite(&id, 4, 1, g); }
            float costs[2] = {a.sbvh->m_cost, a.qbvh->m_cost};
            fwrite(costs, 4, 2, g);
        }
        fclose(g);
    }
    fprintf(stderr,
            "{\"rays\": %llu, \"hits\": %llu, \"threads\": %u, \"qbvh_1t_s\": %.6f, \"qbvh_nt_s\": %.6f, \"sbvh_1t_s\": %.6f, "
            "\"qbvh_vs_sbvh_mismatches\": %llu, \"sbvh_build_s\": %.3f, \"qbvh_build_s\": %.3f, \"triangles\": %lld}\n",
            (unsigned long long)n, (unsigned long long)nhit, threads, secQ1, secQN, secS1, (unsigned long long)mism,
            std::chrono::duration<double>(tb1 - tb0).count(), std::chrono::duration<double>(tq1 - tq0).count(), (long long)nextPrim);
    return 0;
}
