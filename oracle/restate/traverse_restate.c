/*
 * CPU restatement of the reference's intersection path -- TEST INFRASTRUCTURE ONLY.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (slr_b200/) never does. It is the checker, never the thing measured or shipped.
 *
 * It restates, in plain scalar C on the flattened SoA scene (include/slrgpu.h), exactly the
 * algorithm the reference runs:
 *   slab test        libSLR/Accelerator/QBVH.h:55-76   (4 lanes, near/far picked by invDir > 0,
 *                                                        max/min with the SSE operand order)
 *   traversal        libSLR/Accelerator/QBVH.h:295-337 (OrderTable, reverse push of inner children,
 *                                                        in-order immediate leaf tests, distMax shrink)
 *   triangle         libSLR/Surface/TriangleMesh.cpp:131-178 (Moller-Trumbore, ties at t == distMax accepted)
 *   instance         libSLR/Core/SurfaceObject.cpp:307-318, Matrix4x4.h:71-81 (ray to local space,
 *                                                        direction not re-normalised)
 * Pinned against the compiled reference itself (oracle/_ref/ref_intersect): tests/test_oracle.py
 * requires 0 id mismatches and bit-equal t on the committed golden batches. It also counts the
 * nodes popped and leaf records tested per ray -- the algorithmic-bytes model of DESIGN.md.
 *
 * Build without FMA contraction (x86-64 baseline, -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/slrgpu.h"

#define RESTATE_API __attribute__((visibility("default")))

typedef struct { float ox, oy, oz, dx, dy, dz, tmin, tmax; } RRay;
typedef struct { uint32_t prim, inst; float t, u, v; } RHit;

/* _mm_max_ps(a, b) = a > b ? a : b (returns b when either is NaN); _mm_min_ps likewise with < */
static inline float sse_max(float a, float b) { return a > b ? a : b; }
static inline float sse_min(float a, float b) { return a < b ? a : b; }

static uint32_t slab4(const SlrGpuBvhNode* n, const RRay* r, float ix, float iy, float iz) {
    uint32_t mask = 0;
    for (int l = 0; l < 4; ++l) {
        float tn = r->tmin, tf = r->tmax;
        tn = sse_max(((ix > 0.0f ? n->lo_x[l] : n->hi_x[l]) - r->ox) * ix, tn);
        tn = sse_max(((iy > 0.0f ? n->lo_y[l] : n->hi_y[l]) - r->oy) * iy, tn);
        tn = sse_max(((iz > 0.0f ? n->lo_z[l] : n->hi_z[l]) - r->oz) * iz, tn);
        tf = sse_min(((ix > 0.0f ? n->hi_x[l] : n->lo_x[l]) - r->ox) * ix, tf);
        tf = sse_min(((iy > 0.0f ? n->hi_y[l] : n->lo_y[l]) - r->oy) * iy, tf);
        tf = sse_min(((iz > 0.0f ? n->hi_z[l] : n->lo_z[l]) - r->oz) * iz, tf);
        if (tn <= tf) mask |= 1u << l;
    }
    return mask;
}

static int triangle(const SlrGpuLeafRecord* rec, const RRay* r, float* t, float* b0, float* b1o) {
    const float* v0 = rec->a; const float* e1 = rec->b; const float* e2 = rec->c;
    float px = r->dy * e2[2] - r->dz * e2[1];
    float py = r->dz * e2[0] - r->dx * e2[2];
    float pz = r->dx * e2[1] - r->dy * e2[0];
    float det = e1[0] * px + e1[1] * py + e1[2] * pz;
    if (det == 0.0f) return 0;
    float invDet = 1.0f / det;
    float dx = r->ox - v0[0], dy = r->oy - v0[1], dz = r->oz - v0[2];
    float b1 = (dx * px + dy * py + dz * pz) * invDet;
    if (b1 < 0.0f || b1 > 1.0f) return 0;
    float qx = dy * e1[2] - dz * e1[1];
    float qy = dz * e1[0] - dx * e1[2];
    float qz = dx * e1[1] - dy * e1[0];
    float b2 = (r->dx * qx + r->dy * qy + r->dz * qz) * invDet;
    if (b2 < 0.0f || b1 + b2 > 1.0f) return 0;
    float tt = (e2[0] * qx + e2[1] * qy + e2[2] * qz) * invDet;
    if (tt < r->tmin || tt > r->tmax) return 0;
    *t = tt; *b0 = 1.0f - b1 - b2; *b1o = b1;
    return 1;
}

static void mul_point(const float* m, float x, float y, float z, float* o) {
    float tx = m[0] * x + m[4] * y + m[8] * z + m[12] * 1.0f;
    float ty = m[1] * x + m[5] * y + m[9] * z + m[13] * 1.0f;
    float tz = m[2] * x + m[6] * y + m[10] * z + m[14] * 1.0f;
    float tw = m[3] * x + m[7] * y + m[11] * z + m[15] * 1.0f;
    if (tw != 1.0f) { float rc = 1.0f / tw; tx *= rc; ty *= rc; tz *= rc; }
    o[0] = tx; o[1] = ty; o[2] = tz;
}

static const uint32_t kOrderTable[8] = {0x0123, 0x0132, 0x1023, 0x1032, 0x2301, 0x3201, 0x2310, 0x3210};

typedef struct {
    const SlrGpuBvhNode* nodes; const SlrGpuLeafRecord* leaves; const SlrGpuInstance* instances;
    uint64_t nodesVisited, leavesTested;
    int overflow;
} RScene;

static int traverse(RScene* s, uint32_t root, RRay* r, RHit* h, int level) {
    const float ix = 1.0f / r->dx, iy = 1.0f / r->dy, iz = 1.0f / r->dz;
    const int pos[3] = {r->dx >= 0, r->dy >= 0, r->dz >= 0};
    uint32_t stack[64];
    int sp = 0, found = 0;
    stack[sp++] = root;
    while (sp > 0) {
        const SlrGpuBvhNode* n = &s->nodes[stack[--sp]];
        ++s->nodesVisited;
        uint32_t mask = slab4(n, r, ix, iy, iz);
        if (!mask) continue;
        uint32_t enc = kOrderTable[4 * pos[n->top_axis] + 2 * pos[n->left_axis] + pos[n->right_axis]];
        uint32_t ch[4];
        for (int i = 0; i < 4; ++i) {
            uint32_t lane = (enc >> (4 * i)) & 0xF;
            ch[i] = ((mask >> lane) & 1u) ? n->child[lane] : 0xFFFFFFFFu;
        }
        for (int i = 3; i >= 0; --i) {
            if (ch[i] == 0xFFFFFFFFu || (ch[i] >> 31)) continue;
            if (sp >= 64) { s->overflow = 1; continue; }
            stack[sp++] = ch[i] & 0x07FFFFFFu;
        }
        for (int i = 0; i < 4; ++i) {
            if (ch[i] == 0xFFFFFFFFu || !(ch[i] >> 31)) continue;
            uint32_t first = ch[i] & 0x07FFFFFFu, count = (ch[i] >> 27) & 0xFu;
            for (uint32_t j = 0; j < count; ++j) {
                const SlrGpuLeafRecord* rec = &s->leaves[first + j];
                uint32_t id; memcpy(&id, &rec->a[3], 4);
                ++s->leavesTested;
                if (id & 0x80000000u) {
                    if (level >= 1) continue;
                    const SlrGpuInstance* inst = &s->instances[id & 0x7FFFFFFFu];
                    RRay lr; float o[3];
                    mul_point(inst->mat_inv, r->ox, r->oy, r->oz, o);
                    lr.ox = o[0]; lr.oy = o[1]; lr.oz = o[2];
                    const float* m = inst->mat_inv;
                    lr.dx = m[0] * r->dx + m[4] * r->dy + m[8] * r->dz;
                    lr.dy = m[1] * r->dx + m[5] * r->dy + m[9] * r->dz;
                    lr.dz = m[2] * r->dx + m[6] * r->dy + m[10] * r->dz;
                    lr.tmin = r->tmin; lr.tmax = r->tmax;
                    if (traverse(s, inst->root_node, &lr, h, level + 1)) {
                        r->tmax = lr.tmax; h->inst = id & 0x7FFFFFFFu; found = 1;
                    }
                    continue;
                }
                float t, b0, b1;
                if (triangle(rec, r, &t, &b0, &b1)) {
                    r->tmax = t; h->prim = id; h->inst = 0xFFFFFFFFu; h->t = t; h->u = b0; h->v = b1; found = 1;
                }
            }
        }
    }
    return found;
}

/* Closest hit for a ray batch over the flattened scene. Any output pointer may be NULL except prim/inst/t.
 * Returns 0, or 1 if a traversal stack overflowed. */
RESTATE_API int slr_restate_intersect(const SlrGpuSceneDesc* d, const SlrGpuRayBatch* rays, uint64_t n,
                                      const SlrGpuHitBatch* out, uint64_t* totalNodes, uint64_t* totalLeaves) {
    RScene s;
    s.nodes = d->bvh_nodes; s.leaves = d->leaf_records; s.instances = d->instances;
    s.nodesVisited = s.leavesTested = 0; s.overflow = 0;
    for (uint64_t i = 0; i < n; ++i) {
        RRay r = {rays->org_x[i], rays->org_y[i], rays->org_z[i], rays->dir_x[i], rays->dir_y[i], rays->dir_z[i],
                  rays->tmin[i], rays->tmax[i]};
        RHit h = {0xFFFFFFFFu, 0xFFFFFFFFu, INFINITY, 0.0f, 0.0f};
        uint64_t n0 = s.nodesVisited, l0 = s.leavesTested;
        traverse(&s, 0, &r, &h, 0);
        out->prim[i] = h.prim; out->inst[i] = h.inst; out->t[i] = h.t;
        if (out->u) out->u[i] = h.u;
        if (out->v) out->v[i] = h.v;
        if (out->nodes_visited) out->nodes_visited[i] = (uint32_t)(s.nodesVisited - n0);
        if (out->tris_tested) out->tris_tested[i] = (uint32_t)(s.leavesTested - l0);
    }
    if (totalNodes) *totalNodes = s.nodesVisited;
    if (totalLeaves) *totalLeaves = s.leavesTested;
    return s.overflow;
}

/* ---------------------------------------------------------------------------------------------
 * The binary SBVH as the reference's SHIPPED build traverses it (SurfaceObject.cpp:226-230 picks SBVH):
 *   SBVH::intersect            libSLR/Accelerator/SBVH.h:417-442
 *   BoundingBox3D::intersect   libSLR/Core/geometry.h:112-126
 * on the optional sbvh_nodes / sbvh_leaf_records tables of the flattened scene. Pinned against ref_intersect's SBVH pass
 * (tests/golden/intersect_*.npz, fields *_sbvh).
 * ------------------------------------------------------------------------------------------- */
typedef struct { const SlrGpuSbvhNode* nodes; const SlrGpuLeafRecord* leaves; const SlrGpuInstance* instances; int overflow; } SScene;

static int sbvh_box(const SlrGpuSbvhNode* n, const RRay* r, float ix, float iy, float iz) {
    float dist0 = r->tmin, dist1 = r->tmax;
    const float org[3] = {r->ox, r->oy, r->oz}, inv[3] = {ix, iy, iz};
    for (int i = 0; i < 3; ++i) {
        float tn = (n->lo[i] - org[i]) * inv[i], tf = (n->hi[i] - org[i]) * inv[i];
        if (tn > tf) { float sw = tn; tn = tf; tf = sw; }
        dist0 = tn > dist0 ? tn : dist0;
        dist1 = tf < dist1 ? tf : dist1;
        if (dist0 > dist1) return 0;
    }
    return 1;
}

static int sbvh_traverse(SScene* s, uint32_t root, RRay* r, RHit* h, int level) {
    const float ix = 1.0f / r->dx, iy = 1.0f / r->dy, iz = 1.0f / r->dz;
    const int dirPos[3] = {r->dx >= 0.0f, r->dy >= 0.0f, r->dz >= 0.0f};
    uint32_t stack[64];
    int depth = 0, found = 0;
    stack[depth++] = root;
    while (depth > 0) {
        const SlrGpuSbvhNode* n = &s->nodes[stack[--depth]];
        if (!sbvh_box(n, r, ix, iy, iz)) continue;
        if (!(n->b & 0x80000000u)) {
            if (depth + 2 > 64) { s->overflow = 1; continue; }
            const uint32_t c0 = n->a, c1 = n->b & 0x0FFFFFFFu;
            const int positive = dirPos[(n->b >> 28) & 3u];
            stack[depth++] = positive ? c1 : c0;
            stack[depth++] = positive ? c0 : c1;
            continue;
        }
        const uint32_t count = n->b & 0x7FFFFFFFu;
        for (uint32_t j = 0; j < count; ++j) {
            const SlrGpuLeafRecord* rec = &s->leaves[n->a + j];
            uint32_t id; memcpy(&id, &rec->a[3], 4);
            if (id & 0x80000000u) {
                if (level >= 1) continue;
                const SlrGpuInstance* inst = &s->instances[id & 0x7FFFFFFFu];
                RRay lr; float o[3];
                mul_point(inst->mat_inv, r->ox, r->oy, r->oz, o);
                lr.ox = o[0]; lr.oy = o[1]; lr.oz = o[2];
                const float* m = inst->mat_inv;
                lr.dx = m[0] * r->dx + m[4] * r->dy + m[8] * r->dz;
                lr.dy = m[1] * r->dx + m[5] * r->dy + m[9] * r->dz;
                lr.dz = m[2] * r->dx + m[6] * r->dy + m[10] * r->dz;
                lr.tmin = r->tmin; lr.tmax = r->tmax;
                if (sbvh_traverse(s, inst->sbvh_root_node, &lr, h, level + 1)) { r->tmax = lr.tmax; h->inst = id & 0x7FFFFFFFu; found = 1; }
                continue;
            }
            float t, b0, b1;
            if (triangle(rec, r, &t, &b0, &b1)) { r->tmax = t; h->prim = id; h->inst = 0xFFFFFFFFu; h->t = t; h->u = b0; h->v = b1; found = 1; }
        }
    }
    return found;
}

RESTATE_API int slr_restate_intersect_sbvh(const SlrGpuSceneDesc* d, const SlrGpuRayBatch* rays, uint64_t n, const SlrGpuHitBatch* out) {
    if (!d->sbvh_nodes || !d->sbvh_leaf_records) return -1;
    SScene s = {d->sbvh_nodes, d->sbvh_leaf_records, d->instances, 0};
    for (uint64_t i = 0; i < n; ++i) {
        RRay r = {rays->org_x[i], rays->org_y[i], rays->org_z[i], rays->dir_x[i], rays->dir_y[i], rays->dir_z[i], rays->tmin[i], rays->tmax[i]};
        RHit h = {0xFFFFFFFFu, 0xFFFFFFFFu, INFINITY, 0.0f, 0.0f};
        sbvh_traverse(&s, 0, &r, &h, 0);
        out->prim[i] = h.prim; out->inst[i] = h.inst; out->t[i] = h.t;
        if (out->u) out->u[i] = h.u;
        if (out->v) out->v[i] = h.v;
    }
    return s.overflow;
}
