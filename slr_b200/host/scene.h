// Host-side scene graph and its flattening into the SoA buffers the GPU consumes.
//
// Mirrors the part of libSLRSceneGraph / libSLR that feeds the renderer, with the same names where a
// user of the reference would look for them:
//   InternalNode / TriangleMeshNode / ReferenceNode / CameraNode / InfiniteSphereNode
//       libSLRSceneGraph/nodes.h, nodes.cpp:110-212, TriangleMeshNode.cpp:58-112, InfiniteSphereNode.cpp
//   Scene::build            libSLRSceneGraph/Scene.cpp:28-44  +  SLR::Scene::build SurfaceObject.cpp:396-406
//   SurfaceObjectAggregate  libSLR/Core/SurfaceObject.cpp:226-253 (accelerator + light list)
// Instead of allocating one heap object per triangle (SingleSurfaceObject + Triangle + 3 Vertex*),
// flattening appends to flat arrays: SlrGpuVertex / SlrGpuTriangle records, per-aggregate object
// lists, and per-aggregate SBVH -> QBVH trees that are concatenated into one node array and one
// leaf-record array for upload (include/slrgpu.h).
#pragma once
#include "../../include/slrgpu.h"
#include "bvh.h"
#include "geom.h"
#include <memory>
#include <string>
#include <vector>

namespace slr {

class SurfaceMaterial;
class Normal3DTexture;
class FloatTexture;
class IBLEmission;
class SpectrumTexture;
class GpuSceneBuilder;
typedef std::shared_ptr<SurfaceMaterial> SurfaceMaterialRef;
typedef std::shared_ptr<Normal3DTexture> Normal3DTextureRef;
typedef std::shared_ptr<FloatTexture> FloatTextureRef;
typedef std::shared_ptr<SpectrumTexture> SpectrumTextureRef;

class PerspectiveCamera {
public:
    float sensitivity, aspect, fovY, lensRadius, imgPlaneDistance, objPlaneDistance;
    PerspectiveCamera(float sens, float asp, float fov, float lensR, float imgDist, float objDist)
        : sensitivity(sens), aspect(asp), fovY(fov), lensRadius(lensR), imgPlaneDistance(imgDist), objPlaneDistance(objDist) {}
};

// One object of an aggregate, in the order the reference's surfObjs vector would hold them.
struct ObjectRef {
    bool isInstance;
    uint32_t id;            // triangle prim_id, or instance id
};

// A subtree flattened in its own space, ready to be placed any number of times (ReferenceNode, animated nodes): its own
// triangles as one aggregate plus the instances found INSIDE it (instancing nested in instancing), kept as templates.
// place() appends what the parent sees: an instance of the triangle aggregate under `tf` and, for every nested template, an
// instance of that template's aggregate under tf * template -- the reference walks the same chain of
// TransformedSurfaceObjects recursively (SurfaceObject.cpp:307-336); composing the transforms on the host keeps the
// device traversal at one level for any nesting depth. Light-selection probabilities are unchanged by the expansion:
// an aggregate's importance is the sum of its lights' (SurfaceObject.cpp:232-252, 283-285), so the product of the
// pmfs along a chain equals the pmf of the expanded entry.
struct PlacedSubtree {
    bool ready = false;
    bool hasTriangles = false;
    uint32_t triangleAggregate = 0;
    std::vector<uint32_t> nested;            // template instance ids (never referenced by a leaf record themselves)
    bool empty() const { return !hasTriangles && nested.empty(); }
};

struct RenderingData {
    std::vector<ObjectRef> objects;
    std::shared_ptr<PerspectiveCamera> camera;
    StaticTransform cameraTransform;       // with .anim set when the camera sits under an animated node
    bool hasCameraTransform = false;
};

class Node {
public:
    std::string name;
    virtual ~Node() {}
    virtual bool isInstanced() const { return false; }
    // `subTF` is the accumulated parent transform or null (nodes.cpp:110-141).
    virtual void getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) = 0;
    // A flattening pass marks the nodes it has visited (a mesh reachable twice is an error, an instanced subtree is built once);
    // this clears the marks so that the subtree can be flattened again by another pass (scanXZFromYPlus flattens a subtree while
    // the scene is still being written; Scene::build flattens everything later).
    virtual void resetFlattening() {}
    // deep copy (nodes.cpp:69-77, TriangleMeshNode.cpp:52-57); node kinds the reference cannot copy return null
    virtual std::shared_ptr<Node> copy() const { return nullptr; }
};
typedef std::shared_ptr<Node> NodeRef;

class InternalNode : public Node {
    std::vector<NodeRef> m_children;
    StaticTransform m_localToWorld;
public:
    bool addChildNode(const NodeRef& n);
    void setTransform(const StaticTransform& t) { m_localToWorld = t; }
    const StaticTransform& getTransform() const { return m_localToWorld; }
    const std::vector<NodeRef>& children() const { return m_children; }
    void getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) override;
    void resetFlattening() override { for (const NodeRef& c : m_children) c->resetFlattening(); }
    NodeRef copy() const override;
};
typedef std::shared_ptr<InternalNode> InternalNodeRef;

class TriangleMeshNode : public Node {
public:
    struct MaterialGroup {
        SurfaceMaterialRef material;
        Normal3DTextureRef normalMap;
        FloatTextureRef alphaMap;
        std::vector<uint32_t> indices;     // 3 per triangle
    };
private:
    std::vector<Vertex> m_vertices;
    std::vector<MaterialGroup> m_groups;
    bool m_flattened = false;
public:
    uint64_t addVertex(const Vertex& v) { m_vertices.push_back(v); return m_vertices.size() - 1; }
    void addTriangles(const SurfaceMaterialRef& mat, const Normal3DTextureRef& normalMap, const FloatTextureRef& alphaMap,
                      std::vector<uint32_t>&& indices);
    const std::vector<Vertex>& vertices() const { return m_vertices; }
    const std::vector<MaterialGroup>& groups() const { return m_groups; }
    void getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) override;
    void resetFlattening() override { m_flattened = false; }
    NodeRef copy() const override;
};
typedef std::shared_ptr<TriangleMeshNode> TriangleMeshNodeRef;

class ReferenceNode : public Node {
    NodeRef m_node;
    PlacedSubtree m_subtree;
public:
    explicit ReferenceNode(const NodeRef& n) : m_node(n) {}
    bool isInstanced() const override { return true; }
    void getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) override;
    void resetFlattening() override { m_subtree = PlacedSubtree(); m_node->resetFlattening(); }
};

class CameraNode : public Node {
    std::shared_ptr<PerspectiveCamera> m_camera;
public:
    explicit CameraNode(const std::shared_ptr<PerspectiveCamera>& c) : m_camera(c) {}
    void getRenderingData(GpuSceneBuilder& b, const StaticTransform* subTF, RenderingData* data) override;
};

// Environment light: InfiniteSphereNode (libSLRSceneGraph/InfiniteSphereNode.cpp) holding an
// IBLEmission over an image spectrum texture.
class InfiniteSphereNode {
public:
    std::shared_ptr<IBLEmission> emission;
    explicit InfiniteSphereNode(const std::shared_ptr<IBLEmission>& e) : emission(e) {}
};

// The flattened scene: owns every buffer SlrGpuSceneDesc points into.
struct FlatScene {
    std::vector<SlrGpuBvhNode> nodes;
    std::vector<SlrGpuLeafRecord> leaves;
    std::vector<SlrGpuInstance> instances;
    // the binary SBVHs themselves (optional export, exportSbvh: the accelerator the reference's shipped build traverses)
    std::vector<SlrGpuMotion> motions;     // animated transforms of instances / the camera (motion blur)
    uint32_t cameraMotion = 0;             // 0 = static, else 1 + index into motions
    std::vector<SlrGpuSbvhNode> sbvhNodes;
    std::vector<SlrGpuLeafRecord> sbvhLeaves;
    std::vector<SlrGpuTriangle> triangles;
    std::vector<SlrGpuVertex> vertices;
    std::vector<SlrGpuMaterial> materials;
    std::vector<SlrGpuTexture> textures;
    std::vector<SlrGpuSpectrum> spectra;
    std::vector<float> spectrumData;
    std::vector<SlrGpuImage> images;
    std::vector<uint8_t> imageData;
    std::vector<SlrGpuLight> lights;
    uint32_t numTopLights = 0;
    float topLightImportance = 0.0f;
    // environment
    bool envPresent = false;
    uint32_t envMaterial = SLRGPU_INVALID_ID, envMapWidth = 0, envMapHeight = 0;
    std::vector<float> envRowPdf, envRowCdf, envRowIntegral, envMarginalPdf, envMarginalCdf;
    float envMarginalIntegral = 0.0f;
    SlrGpuCamera camera = {};
    bool hasCamera = false;
    float worldCenter[3] = {0, 0, 0};
    float worldRadius = 0.0f;
    bool rgbMode = false;
    static bool exportSbvh;          // process-wide option (slrhost_set_option "export_sbvh"): also flatten the SBVHs
    // build statistics (per aggregate: 0 = top level)
    struct AggregateStats { uint32_t numObjects, sbvhNodes, sbvhRefs, sbvhDepth, qbvhNodes, qbvhDepth, nodeBase, leafBase; float sbvhCost, qbvhCost; };
    std::vector<AggregateStats> stats;
    double buildSeconds = 0.0;

    // Fills a descriptor whose pointers alias this object's vectors (valid while *this is alive and unchanged).
    void describe(SlrGpuSceneDesc* desc) const;
};

class Scene {
    InternalNodeRef m_root;
    std::shared_ptr<InfiniteSphereNode> m_env;
public:
    Scene();
    const InternalNodeRef& rootNode() const { return m_root; }
    void setEnvNode(const std::shared_ptr<InfiniteSphereNode>& e) { m_env = e; }
    const std::shared_ptr<InfiniteSphereNode>& envNode() const { return m_env; }
    // Flattens the graph, builds every aggregate's SBVH -> QBVH and the shading tables.
    // Throws std::runtime_error on unsupported input (e.g. instancing nested deeper than one level).
    void build(FlatScene* out, bool rgbMode = false);
};

// Working state of one flattening pass.
class GpuSceneBuilder {
public:
    struct Aggregate {
        std::vector<ObjectRef> objects;
        SBVH sbvh;
        QBVH qbvh;
        std::vector<SlrGpuLight> lights;     // emitting objects in object order
        bool containsInstances = false;
        float lightImportance = 0.0f;        // integral of the light distribution
    };
    FlatScene& flat;
    std::vector<Aggregate> aggregates;       // nested ones first as they are discovered; index = aggregate id
    std::vector<uint32_t> instanceAggregate; // instance id -> aggregate id
    std::vector<std::shared_ptr<const AnimatedTransform>> instanceAnim;     // instance id -> its motion (null: static)
    bool instanceIsAnimated(uint32_t id) const { return id < instanceAnim.size() && instanceAnim[id] != nullptr; }
    std::vector<uint8_t> triangleEmits;      // prim_id -> is emitting

    explicit GpuSceneBuilder(FlatScene& f) : flat(f) {}
    // Builds trees + light list for `objects`; returns the aggregate id.
    uint32_t createAggregate(std::vector<ObjectRef>&& objects);
    void prepare(PlacedSubtree& ps, std::vector<ObjectRef>&& objects);
    void place(const PlacedSubtree& ps, const StaticTransform& tf, std::vector<ObjectRef>* out);
    // tf.anim set: the instance moves (its record refers to a new entry of flat.motions)
    uint32_t addInstance(uint32_t aggregate, const StaticTransform& tf);
    uint32_t addMotion(const AnimatedTransform& a);       // returns 1 + index
    // Concatenates all aggregates (top level = `top` first) into flat.nodes / flat.leaves.
    void finalize(uint32_t top);
    // shading tables (materials.cpp)
    uint32_t exportMaterial(const SurfaceMaterial* m);
    uint32_t exportNormalTexture(const Normal3DTexture* t);
    uint32_t exportFloatTexture(const FloatTexture* t);
    bool materialEmits(const SurfaceMaterial* m) const;
};

}  // namespace slr
