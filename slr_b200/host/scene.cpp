#include "scene.h"
#include "spectrum.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace slr {

bool FlatScene::exportSbvh = false;

// ---------------------------------------------------------------------------------------------
// nodes
// ---------------------------------------------------------------------------------------------

static bool containsNode(const InternalNode* self, const Node* target) {
    for (const NodeRef& c : self->children()) {
        if (c.get() == target) return true;
        if (const InternalNode* in = dynamic_cast<const InternalNode*>(c.get()))
            if (containsNode(in, target)) return true;
    }
    return false;
}

bool InternalNode::addChildNode(const NodeRef& n) {
    if (n.get() == this) return false;
    if (const InternalNode* in = dynamic_cast<const InternalNode*>(n.get()))
        if (containsNode(in, this)) return false;                 // would create a cycle
    if (n->isInstanced()) {
        if (std::find(m_children.begin(), m_children.end(), n) != m_children.end()) return false;
    } else if (containsNode(this, n.get())) {
        return false;
    }
    m_children.push_back(n);
    return true;
}

NodeRef InternalNode::copy() const {
    auto ret = std::make_shared<InternalNode>();
    ret->m_localToWorld = m_localToWorld;
    for (const NodeRef& c : m_children) {
        NodeRef cc = c->copy();
        if (!cc) throw std::runtime_error("copyNode: the subtree holds a node that cannot be copied (a reference or camera node)");
        ret->m_children.push_back(cc);
    }
    return ret;
}

NodeRef TriangleMeshNode::copy() const {
    auto ret = std::make_shared<TriangleMeshNode>();
    ret->m_vertices = m_vertices;
    ret->m_groups = m_groups;
    return ret;
}

void InternalNode::getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) {
    if (subTF && subTF->anim) throw std::runtime_error("internal: an animated transform was handed down the scene graph");
    if (!m_localToWorld.anim) {
        // parent * local, with the inverse recomputed from the product as StaticTransform's ctor does
        StaticTransform reduced = subTF ? (*subTF * m_localToWorld) : m_localToWorld;
        for (const NodeRef& c : m_children) c->getRenderingData(b, &reduced, data);
        return;
    }
    // An animated node (nodes.cpp:117-141): its subtree is flattened in its own space into a nested aggregate, which the
    // parent sees through a TransformedSurfaceObject carrying the animated transform -- an instance that moves. A static
    // parent transform is folded into both key frames (ChainedTransform::reduce -> createByMulLeft, Transform.cpp:67-69).
    const std::shared_ptr<const AnimatedTransform> reduced = subTF ? m_localToWorld.anim->mulLeft(*subTF) : m_localToWorld.anim;
    RenderingData sub;
    for (const NodeRef& c : m_children) c->getRenderingData(b, nullptr, &sub);
    if (!sub.objects.empty()) {
        PlacedSubtree subtree;
        b.prepare(subtree, std::move(sub.objects));
        StaticTransform tf(reduced->begin.mat, reduced->begin.matInv);
        tf.anim = reduced;
        b.place(subtree, tf, &data->objects);
    }
    if (sub.camera) {
        // the camera rides on the animated node: animated * (static camera transform) (createByMulRight, Transform.cpp:70-72)
        if (sub.cameraTransform.anim) throw std::runtime_error("a camera under two animated nodes is not supported");
        data->camera = sub.camera;
        const std::shared_ptr<const AnimatedTransform> camAnim = sub.hasCameraTransform ? reduced->mulRight(sub.cameraTransform) : reduced;
        data->cameraTransform = StaticTransform(camAnim->begin.mat, camAnim->begin.matInv);
        data->cameraTransform.anim = camAnim;
        data->hasCameraTransform = true;
    }
}

void TriangleMeshNode::addTriangles(const SurfaceMaterialRef& mat, const Normal3DTextureRef& normalMap,
                                    const FloatTextureRef& alphaMap, std::vector<uint32_t>&& indices) {
    if (indices.size() % 3 != 0) throw std::runtime_error("TriangleMeshNode::addTriangles: index count is not a multiple of 3");
    for (uint32_t i : indices)
        if (i >= m_vertices.size()) throw std::runtime_error("TriangleMeshNode::addTriangles: vertex index out of range");
    m_groups.emplace_back();
    MaterialGroup& g = m_groups.back();
    g.material = mat; g.normalMap = normalMap; g.alphaMap = alphaMap;
    g.indices = std::move(indices);
}

void TriangleMeshNode::getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) {
    if (m_flattened)
        throw std::runtime_error("a TriangleMeshNode is reachable twice in one flattening pass; wrap it in a ReferenceNode to instance it");
    m_flattened = true;
    StaticTransform tf = subTF ? *subTF : StaticTransform();
    FlatScene& f = b.flat;
    const uint32_t vBase = (uint32_t)f.vertices.size();
    f.vertices.reserve(f.vertices.size() + m_vertices.size());
    for (const Vertex& v : m_vertices) {
        // baked exactly as applyTransformForRendering does (TriangleMeshNode.cpp:68-78)
        Vec3 p = tf.point(v.position);
        Vec3 n = normalize(tf.normal(v.normal));
        Vec3 t = normalize(tf.vector(v.tangent));
        SlrGpuVertex o;
        o.position[0] = p.x; o.position[1] = p.y; o.position[2] = p.z; o.u = v.texCoord.u;
        o.normal[0] = n.x; o.normal[1] = n.y; o.normal[2] = n.z; o.v = v.texCoord.v;
        o.tangent[0] = t.x; o.tangent[1] = t.y; o.tangent[2] = t.z; o.pad = 0.0f;
        f.vertices.push_back(o);
    }
    for (const MaterialGroup& g : m_groups) {
        const uint32_t mat = g.material ? b.exportMaterial(g.material.get()) : SLRGPU_INVALID_ID;
        const uint32_t nmap = g.normalMap ? b.exportNormalTexture(g.normalMap.get()) : SLRGPU_INVALID_ID;
        const uint32_t amap = g.alphaMap ? b.exportFloatTexture(g.alphaMap.get()) : SLRGPU_INVALID_ID;
        const bool emits = g.material && b.materialEmits(g.material.get());
        for (size_t i = 0; i + 2 < g.indices.size(); i += 3) {
            SlrGpuTriangle t;
            t.v[0] = vBase + g.indices[i]; t.v[1] = vBase + g.indices[i + 1]; t.v[2] = vBase + g.indices[i + 2];
            t.material = mat; t.normal_map = nmap; t.alpha_map = amap;
            t.light_index = SLRGPU_INVALID_ID; t.pad = 0;
            const uint32_t id = (uint32_t)f.triangles.size();
            if (id >= 0x80000000u) throw std::runtime_error("more than 2^31 triangles");
            f.triangles.push_back(t);
            b.triangleEmits.push_back(emits ? 1 : 0);
            data->objects.push_back(ObjectRef{false, id});
        }
    }
}

void ReferenceNode::getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) {
    if (!m_subtree.ready) {
        RenderingData sub;
        m_node->getRenderingData(b, nullptr, &sub);
        if (sub.objects.empty()) throw std::runtime_error("ReferenceNode refers to a subtree without surfaces");
        b.prepare(m_subtree, std::move(sub.objects));
    }
    b.place(m_subtree, subTF ? *subTF : StaticTransform(), &data->objects);
}

void CameraNode::getRenderingData(GpuSceneBuilder&, const StaticTransform* subTF, RenderingData* data) {
    data->camera = m_camera;
    data->cameraTransform = subTF ? *subTF : StaticTransform();
    data->hasCameraTransform = subTF != nullptr;
}

uint32_t GpuSceneBuilder::addMotion(const AnimatedTransform& a) {
    SlrGpuMotion m;
    std::memset(&m, 0, sizeof(m));
    std::memcpy(m.mat_end, &a.end.mat, 64);
    std::memcpy(m.mat_end_inv, &a.end.matInv, 64);
    for (int k = 0; k < 3; ++k) { m.T0[k] = a.T[0][k]; m.T1[k] = a.T[1][k]; }
    m.t_begin = a.tBegin; m.t_end = a.tEnd;
    m.R0[0] = a.R[0].x; m.R0[1] = a.R[0].y; m.R0[2] = a.R[0].z; m.R0[3] = a.R[0].w;
    m.R1[0] = a.R[1].x; m.R1[1] = a.R[1].y; m.R1[2] = a.R[1].z; m.R1[3] = a.R[1].w;
    std::memcpy(m.S0, &a.S[0], 64);
    std::memcpy(m.S1, &a.S[1], 64);
    flat.motions.push_back(m);
    return (uint32_t)flat.motions.size();
}

// ---------------------------------------------------------------------------------------------
// aggregates
// ---------------------------------------------------------------------------------------------

static Vec3 vpos(const SlrGpuVertex& v) { return Vec3(v.position[0], v.position[1], v.position[2]); }

uint32_t GpuSceneBuilder::createAggregate(std::vector<ObjectRef>&& objects) {
    aggregates.emplace_back();
    const uint32_t id = (uint32_t)aggregates.size() - 1;
    Aggregate& ag = aggregates.back();
    ag.objects = std::move(objects);

    PrimitiveSet ps;
    ps.prims.reserve(ag.objects.size());
    for (const ObjectRef& o : ag.objects) {
        if (!o.isInstance) {
            const SlrGpuTriangle& t = flat.triangles[o.id];
            ps.addTriangle(vpos(flat.vertices[t.v[0]]), vpos(flat.vertices[t.v[1]]), vpos(flat.vertices[t.v[2]]));
        } else {
            ag.containsInstances = true;
            const Aggregate& nested = aggregates[instanceAggregate[o.id]];
            if (nested.containsInstances)
                throw std::runtime_error("instancing nested deeper than one level is not supported by the GPU traversal yet");
            Mat4 m;
            std::memcpy(static_cast<void*>(&m), flat.instances[o.id].mat, sizeof(float) * 16);
            // TransformedSurfaceObject::bounds / costForIntersect (SurfaceObject.cpp:303-305, SurfaceObject.h:206): the
            // transform's motionBounds of the nested bounds -- for a static transform that is the transformed box
            float cost = nested.objects.size() == 1 ? 1.0f : nested.sbvh.cost;
            const std::shared_ptr<const AnimatedTransform>& anim = instanceAnim[o.id];
            ps.addBox(anim ? anim->motionBounds(nested.sbvh.bounds) : transformBounds(m, nested.sbvh.bounds), cost);
        }
    }
    ag.sbvh.build(ps);
    ag.qbvh.build(ag.sbvh, ps);

    // light list: emitting objects in object order (SurfaceObject.cpp:232-252)
    for (const ObjectRef& o : ag.objects) {
        SlrGpuLight l;
        std::memset(&l, 0, sizeof(l));
        if (!o.isInstance) {
            if (!triangleEmits[o.id]) continue;
            flat.triangles[o.id].light_index = (uint32_t)ag.lights.size();
            l.object = o.id; l.importance = 1.0f;                       // SingleSurfaceObject::importance
        } else {
            const Aggregate& nested = aggregates[instanceAggregate[o.id]];
            if (nested.lights.empty()) continue;
            flat.instances[o.id].light_index = (uint32_t)ag.lights.size();
            flat.instances[o.id].light_importance = nested.lightImportance;
            l.object = 0x80000000u | o.id; l.importance = nested.lightImportance;
        }
        ag.lights.push_back(l);
    }
    // RegularConstantDiscrete1D over the importances: compensated running sum, then normalise
    if (!ag.lights.empty()) {
        const size_t n = ag.lights.size();
        std::vector<float> cdf(n + 1, 0.0f);
        float sum = 0.0f, comp = 0.0f;
        for (size_t i = 0; i < n; ++i) {
            float y = ag.lights[i].importance - comp;
            float t = sum + y;
            comp = (t - sum) - y;
            sum = t;
            cdf[i + 1] = sum;
        }
        ag.lightImportance = sum;
        for (size_t i = 0; i < n; ++i) {
            ag.lights[i].pmf = ag.lights[i].importance / sum;
            cdf[i + 1] /= sum;
        }
        for (size_t i = 0; i < n; ++i) { ag.lights[i].cdf_lo = cdf[i]; ag.lights[i].cdf_hi = cdf[i + 1]; }
    }
    return id;
}

void GpuSceneBuilder::prepare(PlacedSubtree& ps, std::vector<ObjectRef>&& objects) {
    std::vector<ObjectRef> triangles;
    for (const ObjectRef& o : objects) {
        if (o.isInstance) ps.nested.push_back(o.id);
        else triangles.push_back(o);
    }
    ps.hasTriangles = !triangles.empty();
    if (ps.hasTriangles) ps.triangleAggregate = createAggregate(std::move(triangles));
    ps.ready = true;
}

void GpuSceneBuilder::place(const PlacedSubtree& ps, const StaticTransform& tf, std::vector<ObjectRef>* out) {
    if (ps.hasTriangles) out->push_back(ObjectRef{true, addInstance(ps.triangleAggregate, tf)});
    for (uint32_t t : ps.nested) {
        // the template's own transform (inside the subtree's space), then the subtree's placement
        Mat4 m, mi;
        std::memcpy(static_cast<void*>(&m), flat.instances[t].mat, sizeof(float) * 16);
        std::memcpy(static_cast<void*>(&mi), flat.instances[t].mat_inv, sizeof(float) * 16);
        const StaticTransform inner(m, mi);
        const std::shared_ptr<const AnimatedTransform> innerAnim = instanceAnim[t];
        StaticTransform composed;
        if (tf.anim && innerAnim)
            throw std::runtime_error("an animated node inside an animated node (a chain of two moving transforms) is not supported");
        if (tf.anim) {
            const std::shared_ptr<const AnimatedTransform> a = tf.anim->mulRight(inner);       // animated * static (Transform.cpp:70-72)
            composed = StaticTransform(a->begin.mat, a->begin.matInv);
            composed.anim = a;
        } else if (innerAnim) {
            const std::shared_ptr<const AnimatedTransform> a = innerAnim->mulLeft(tf);         // static * animated (Transform.cpp:67-69)
            composed = StaticTransform(a->begin.mat, a->begin.matInv);
            composed.anim = a;
        } else {
            composed = tf * inner;        // the inverse recomputed from the product, as StaticTransform's operator* does
        }
        out->push_back(ObjectRef{true, addInstance(instanceAggregate[t], composed)});
    }
}

uint32_t GpuSceneBuilder::addInstance(uint32_t aggregate, const StaticTransform& tf) {
    SlrGpuInstance inst;
    std::memset(&inst, 0, sizeof(inst));
    std::memcpy(inst.mat, &tf.mat, sizeof(float) * 16);
    std::memcpy(inst.mat_inv, &tf.matInv, sizeof(float) * 16);
    inst.root_node = 0;           // patched in finalize()
    inst.light_base = SLRGPU_INVALID_ID;
    inst.num_lights = 0;
    inst.light_index = SLRGPU_INVALID_ID;
    inst.motion = tf.anim ? addMotion(*tf.anim) : 0u;
    flat.instances.push_back(inst);
    instanceAggregate.push_back(aggregate);
    instanceAnim.push_back(tf.anim);
    return (uint32_t)flat.instances.size() - 1;
}

void GpuSceneBuilder::finalize(uint32_t top) {
    std::vector<uint32_t> order;
    order.push_back(top);
    for (uint32_t i = 0; i < aggregates.size(); ++i) if (i != top) order.push_back(i);

    std::vector<uint32_t> nodeBase(aggregates.size()), leafBase(aggregates.size()), lightBase(aggregates.size());
    uint64_t nNodes = 0, nLeaves = 0, nLights = 0;
    for (uint32_t a : order) {
        nodeBase[a] = (uint32_t)nNodes; leafBase[a] = (uint32_t)nLeaves; lightBase[a] = (uint32_t)nLights;
        nNodes += aggregates[a].qbvh.nodes.size();
        nLeaves += aggregates[a].qbvh.refs.size();
        nLights += aggregates[a].lights.size();
    }
    if (nNodes > 0x07FFFFFFull || nLeaves > 0x07FFFFFFull)
        throw std::runtime_error("scene exceeds the 27-bit node / leaf index of the QBVH child word");
    flat.nodes.resize(nNodes);
    flat.leaves.resize(nLeaves);
    flat.lights.resize(nLights);
    flat.numTopLights = (uint32_t)aggregates[top].lights.size();
    flat.topLightImportance = aggregates[top].lightImportance;
    flat.stats.clear();

    auto fillLeafRecord = [this](const ObjectRef& o, SlrGpuLeafRecord& r) {
        std::memset(&r, 0, sizeof(r));
        uint32_t idBits;
        if (!o.isInstance) {
            const SlrGpuTriangle& t = flat.triangles[o.id];
            Vec3 p0 = vpos(flat.vertices[t.v[0]]), p1 = vpos(flat.vertices[t.v[1]]), p2 = vpos(flat.vertices[t.v[2]]);
            Vec3 e1 = p1 - p0, e2 = p2 - p0;
            r.a[0] = p0.x; r.a[1] = p0.y; r.a[2] = p0.z;
            r.b[0] = e1.x; r.b[1] = e1.y; r.b[2] = e1.z;
            r.c[0] = e2.x; r.c[1] = e2.y; r.c[2] = e2.z;
            uint32_t flags = t.alpha_map != SLRGPU_INVALID_ID ? SLRGPU_LEAF_FLAG_ALPHA_TEST : 0u;
            std::memcpy(&r.b[3], &flags, 4);
            idBits = o.id;
        } else {
            idBits = 0x80000000u | o.id;
        }
        std::memcpy(&r.a[3], &idBits, 4);
    };
    // optional: the binary SBVHs as they are (SBVH.h:21-45), global indices, leaf records in the SBVH's own leaf order
    std::vector<uint32_t> sbvhNodeBase(aggregates.size(), 0);
    flat.sbvhNodes.clear(); flat.sbvhLeaves.clear();
    if (FlatScene::exportSbvh) {
        uint64_t nS = 0, nL = 0;
        for (uint32_t a : order) { nS += aggregates[a].sbvh.nodes.size(); nL += aggregates[a].sbvh.refs.size(); }
        if (nS >= 0x10000000ull || nL >= 0x80000000ull) throw std::runtime_error("scene too large for the SBVH export (28-bit node index)");
        flat.sbvhNodes.reserve(nS); flat.sbvhLeaves.reserve(nL);
        for (uint32_t a : order) {
            const Aggregate& ag = aggregates[a];
            const uint32_t nb = (uint32_t)flat.sbvhNodes.size(), lb = (uint32_t)flat.sbvhLeaves.size();
            sbvhNodeBase[a] = nb;
            for (const SBVHNode& s : ag.sbvh.nodes) {
                SlrGpuSbvhNode d;
                d.lo[0] = s.bbox.lo.x; d.lo[1] = s.bbox.lo.y; d.lo[2] = s.bbox.lo.z;
                d.hi[0] = s.bbox.hi.x; d.hi[1] = s.bbox.hi.y; d.hi[2] = s.bbox.hi.z;
                if (s.numRefs > 0) { d.a = lb + s.firstRef; d.b = 0x80000000u | s.numRefs; }
                else { d.a = nb + s.c0; d.b = (nb + s.c1) | ((uint32_t)s.axis << 28); }
                flat.sbvhNodes.push_back(d);
            }
            for (uint32_t ref : ag.sbvh.refs) {
                SlrGpuLeafRecord r;
                fillLeafRecord(ag.objects[ref], r);
                flat.sbvhLeaves.push_back(r);
            }
        }
    }

    for (uint32_t a : order) {
        const Aggregate& ag = aggregates[a];
        for (size_t i = 0; i < ag.qbvh.nodes.size(); ++i) {
            const QBVHNode& s = ag.qbvh.nodes[i];
            SlrGpuBvhNode& d = flat.nodes[nodeBase[a] + i];
            static_assert(sizeof(QBVHNode) == sizeof(SlrGpuBvhNode), "node layouts must match");
            std::memcpy(&d, &s, sizeof(d));
            for (int l = 0; l < 4; ++l) {
                uint32_t c = s.child[l];
                if (c == kQBVHEmptyChild) continue;
                uint32_t idx = qbvhChildIdx(c) + (qbvhChildIsLeaf(c) ? leafBase[a] : nodeBase[a]);
                d.child[l] = (c & 0xF8000000u) | idx;
            }
        }
        for (size_t i = 0; i < ag.qbvh.refs.size(); ++i) {
            fillLeafRecord(ag.objects[ag.qbvh.refs[i]], flat.leaves[leafBase[a] + i]);
        }
        for (size_t i = 0; i < ag.lights.size(); ++i) flat.lights[lightBase[a] + i] = ag.lights[i];
        FlatScene::AggregateStats st;
        st.numObjects = (uint32_t)ag.objects.size();
        st.sbvhNodes = (uint32_t)ag.sbvh.nodes.size(); st.sbvhRefs = (uint32_t)ag.sbvh.refs.size(); st.sbvhDepth = ag.sbvh.depth;
        st.qbvhNodes = (uint32_t)ag.qbvh.nodes.size(); st.qbvhDepth = ag.qbvh.depth;
        st.nodeBase = nodeBase[a]; st.leafBase = leafBase[a];
        st.sbvhCost = ag.sbvh.cost; st.qbvhCost = ag.qbvh.cost;
        flat.stats.push_back(st);
    }
    for (size_t i = 0; i < flat.instances.size(); ++i) {
        const uint32_t a = instanceAggregate[i];
        flat.instances[i].root_node = nodeBase[a];
        flat.instances[i].sbvh_root_node = sbvhNodeBase[a];
        flat.instances[i].light_base = aggregates[a].lights.empty() ? SLRGPU_INVALID_ID : lightBase[a];
        flat.instances[i].num_lights = (uint32_t)aggregates[a].lights.size();
    }
    const BBox& wb = aggregates[top].sbvh.bounds;
    Vec3 c = wb.centroid();
    flat.worldCenter[0] = c.x; flat.worldCenter[1] = c.y; flat.worldCenter[2] = c.z;
    flat.worldRadius = (wb.hi - c).length();
}

// ---------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------

Scene::Scene() : m_root(std::make_shared<InternalNode>()) { m_root->name = "root"; }

void FlatScene::describe(SlrGpuSceneDesc* d) const {
    std::memset(d, 0, sizeof(*d));
    d->struct_size = sizeof(SlrGpuSceneDesc);
    d->rgb_mode = rgbMode ? 1 : 0;
    d->bvh_nodes = nodes.data();       d->num_bvh_nodes = (uint32_t)nodes.size();
    d->leaf_records = leaves.data();   d->num_leaf_records = (uint32_t)leaves.size();
    d->instances = instances.data();   d->num_instances = (uint32_t)instances.size();
    d->triangles = triangles.data();   d->num_triangles = (uint32_t)triangles.size();
    d->vertices = vertices.data();     d->num_vertices = (uint32_t)vertices.size();
    d->materials = materials.data();   d->num_materials = (uint32_t)materials.size();
    d->textures = textures.data();     d->num_textures = (uint32_t)textures.size();
    d->spectra = spectra.data();       d->num_spectra = (uint32_t)spectra.size();
    d->spectrum_data = spectrumData.data(); d->num_spectrum_floats = (uint32_t)spectrumData.size();
    d->images = images.data();         d->num_images = (uint32_t)images.size();
    d->image_data = imageData.data();  d->image_data_bytes = imageData.size();
    d->lights = lights.data();         d->num_lights = (uint32_t)lights.size();
    d->num_top_lights = numTopLights;
    d->top_light_importance = topLightImportance;
    for (int i = 0; i < 3; ++i) d->world_center[i] = worldCenter[i];
    d->world_radius = worldRadius;
    d->camera = camera;
    if (!motions.empty()) { d->motions = motions.data(); d->num_motions = (uint32_t)motions.size(); }
    d->camera_motion = cameraMotion;
    if (!sbvhNodes.empty()) {
        d->sbvh_nodes = sbvhNodes.data(); d->num_sbvh_nodes = (uint32_t)sbvhNodes.size();
        d->sbvh_leaf_records = sbvhLeaves.data(); d->num_sbvh_leaf_records = (uint32_t)sbvhLeaves.size();
    }
    d->environment.present = envPresent ? 1 : 0;
    d->environment.material = envMaterial;
    d->environment.map_width = envMapWidth; d->environment.map_height = envMapHeight;
    d->environment.row_pdf = envRowPdf.data(); d->environment.row_cdf = envRowCdf.data();
    d->environment.row_integral = envRowIntegral.data();
    d->environment.marginal_pdf = envMarginalPdf.data(); d->environment.marginal_cdf = envMarginalCdf.data();
    d->environment.marginal_integral = envMarginalIntegral;
    if (!materials.empty()) {
        // shading needs the spectral constant tables (loaded once per process)
        const SpectralTables& T = SpectralTables::instance();
        const std::vector<float>& pts = T.floats("upsampling/points");
        d->spectral.upsample_grid = T.upsampleGridWords.data();
        d->spectral.upsample_grid_floats = (uint32_t)T.upsampleGridWords.size();
        d->spectral.upsample_points = pts.data();
        d->spectral.upsample_points_floats = (uint32_t)pts.size();
        d->spectral.xbar_16 = T.xbar16; d->spectral.ybar_16 = T.ybar16; d->spectral.zbar_16 = T.zbar16;
        d->spectral.integral_cmf = T.integralCMF;
    }
}

}  // namespace slr
