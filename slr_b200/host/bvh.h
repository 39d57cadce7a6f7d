// Host-side acceleration-structure build: spatial-split BVH (SBVH) and its collapse to the 4-wide
// QBVH that the GPU traverses. Both stay on the host by design (north_star: "the SBVH/QBVH build and
// scene-graph flattening remain on the host and are uploaded as SoA buffers").
//
// The builders are written to reproduce the reference's trees exactly (same fp32 arithmetic, same
// bin counts, same tie-breaking, same partition order), because bit-exact hit ids on tied distances
// depend on visiting leaves in the reference's order:
//   SBVH build      libSLR/Accelerator/SBVH.h:57-348, 379-407
//   QBVH collapse   libSLR/Accelerator/QBVH.h:85-202, 253-285
//   triangle chop / split bounds   libSLR/Surface/TriangleMesh.cpp:19-125
//   generic chop / split bounds    libSLR/Core/SurfaceObject.h:58-86
#pragma once
#include "geom.h"
#include <vector>

namespace slr {

// The objects an aggregate is built over. A primitive is either a triangle (exact clipping when the
// SBVH chops it) or an opaque box (an instance: its bounds are clipped as a box).
struct PrimitiveSet {
    struct Prim {
        BBox bounds;
        float cost;          // costForIntersect(): 1 for a triangle, SAH cost of the nested BVH for an instance
        bool isTriangle;
        Vec3 p[3];           // triangle corners (world/aggregate space)
    };
    std::vector<Prim> prims;
    void addTriangle(const Vec3& a, const Vec3& b, const Vec3& c);
    void addBox(const BBox& b, float cost);
};

struct SBVHNode {
    BBox bbox;
    uint32_t c0 = 0, c1 = 0;
    uint32_t firstRef = 0, numRefs = 0;   // leaf iff numRefs > 0
    Axis axis = Axis_X;
};

struct SBVH {
    std::vector<SBVHNode> nodes;
    std::vector<uint32_t> refs;           // primitive indices, leaf by leaf (duplicates possible)
    BBox bounds;
    uint32_t depth = 0;
    uint32_t numFragmentsAdded = 0;
    float cost = 0.0f;                    // SAH cost, SBVH.h:350-376
    void build(const PrimitiveSet& prims);
};

// 128-byte node, same field order as the reference's SSE node (QBVH.h:42-53) so one node is eight
// aligned 16-byte loads on the GPU: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] axes+pad.
struct alignas(16) QBVHNode {
    float lo_x[4], lo_y[4], lo_z[4];
    float hi_x[4], hi_y[4], hi_z[4];
    uint32_t child[4];                    // idx:27 | numLeaves:4 << 27 | isLeaf << 31 ; 0xFFFFFFFF = empty
    uint8_t topAxis, leftAxis, rightAxis, pad0;
    uint32_t pad[3];
};
static_assert(sizeof(QBVHNode) == 128, "QBVH node must be 128 bytes");

constexpr uint32_t kQBVHEmptyChild = 0xFFFFFFFFu;
inline uint32_t qbvhChildIdx(uint32_t c) { return c & 0x07FFFFFFu; }
inline uint32_t qbvhChildNumLeaves(uint32_t c) { return (c >> 27) & 0xFu; }
inline bool qbvhChildIsLeaf(uint32_t c) { return (c >> 31) != 0; }

struct QBVH {
    std::vector<QBVHNode> nodes;
    std::vector<uint32_t> refs;
    BBox bounds;
    uint32_t depth = 0;
    float cost = 0.0f;                    // QBVH.h:204-250
    // Throws std::runtime_error when a leaf has >= 16 references (the reference's 4-bit field would
    // silently wrap, QBVH.h:27-35 / :93) or when an index exceeds 27 bits.
    void build(const SBVH& sbvh, const PrimitiveSet& prims);
};

}  // namespace slr
