// Host-side closest hit + surface point over a flattened scene: what the scene language needs when a builtin casts
// rays DURING scene construction (scanXZFromYPlus, libSLRSceneGraph/API.cpp:926-983 -- the reference builds a
// SurfaceObjectAggregate over the node and calls aggregate->intersect / Intersection::getSurfacePoint per ray).
// A handful of rays per scene: this is not a render path and has no GPU counterpart to fall back from.
//   QBVH::Node::intersect / QBVH::intersect   libSLR/Accelerator/QBVH.h:55-76, 295-337
//   Triangle::intersect / getSurfacePoint     libSLR/Surface/TriangleMesh.cpp:131-215
//   TransformedSurfaceObject                  libSLR/Core/SurfaceObject.cpp:307-336, geometry.cpp:63-78
#pragma once
#include "scene.h"

namespace slr {

struct HostHit {
    uint32_t prim = SLRGPU_INVALID_ID, inst = SLRGPU_INVALID_ID;
    float t = INFINITY, b0 = 0.0f, b1 = 0.0f;
};

struct HostSurfacePoint {
    Vec3 p, gNormal;
    Vec3 sx, sy, sz;        // shading frame: tangent, bitangent, normal
};

class HostRayCaster {
    const FlatScene& m_scene;
    bool walk(uint32_t root, Vec3 org, Vec3 dir, float tmin, float* tmax, int level, HostHit* hit) const;
public:
    explicit HostRayCaster(const FlatScene& scene) : m_scene(scene) {}
    BBox bounds() const;
    // closest hit in [tmin, tmax]; same visiting order and fp32 arithmetic as the reference's QBVH
    bool intersect(const Vec3& org, const Vec3& dir, float tmin, float tmax, HostHit* hit) const;
    // Intersection::getSurfacePoint of a triangle hit. Throws on a normal-mapped triangle (texture evaluation lives on the GPU).
    void surfacePoint(const HostHit& hit, const Vec3& org, const Vec3& dir, HostSurfacePoint* sp) const;
};

}  // namespace slr
