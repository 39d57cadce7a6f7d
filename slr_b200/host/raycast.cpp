#include "raycast.h"
#include <cstring>
#include <stdexcept>

namespace slr {

namespace {

// _mm_max_ps(a, b) = a > b ? a : b (b when either is NaN); _mm_min_ps likewise with <
inline float maxps(float a, float b) { return a > b ? a : b; }
inline float minps(float a, float b) { return a < b ? a : b; }

Vec3 mulPoint(const float* m, const Vec3& p) {
    float tx = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12] * 1.0f;
    float ty = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13] * 1.0f;
    float tz = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14] * 1.0f;
    const float tw = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15] * 1.0f;
    if (tw != 1.0f) { const float rc = 1.0f / tw; tx *= rc; ty *= rc; tz *= rc; }
    return Vec3(tx, ty, tz);
}
Vec3 mulVector(const float* m, const Vec3& v) {
    return Vec3(m[0] * v.x + m[4] * v.y + m[8] * v.z, m[1] * v.x + m[5] * v.y + m[9] * v.z, m[2] * v.x + m[6] * v.y + m[10] * v.z);
}
// StaticTransform * Normal3D: the transposed inverse (Transform.h:47-52)
Vec3 mulNormal(const float* mi, const Vec3& n) {
    return Vec3(mi[0] * n.x + mi[1] * n.y + mi[2] * n.z, mi[4] * n.x + mi[5] * n.y + mi[6] * n.z, mi[8] * n.x + mi[9] * n.y + mi[10] * n.z);
}

// visiting order of the four lanes by the ray's direction signs along the node's three split axes (QBVH.h:309-312)
const uint32_t kOrder[8] = {0x0123, 0x0132, 0x1023, 0x1032, 0x2301, 0x3201, 0x2310, 0x3210};

}  // namespace

BBox HostRayCaster::bounds() const {
    BBox b;
    if (m_scene.nodes.empty()) return b;
    const SlrGpuBvhNode& n = m_scene.nodes[0];
    for (int l = 0; l < 4; ++l) {
        if (n.child[l] == 0xFFFFFFFFu) continue;
        b.grow(Vec3(n.lo_x[l], n.lo_y[l], n.lo_z[l]));
        b.grow(Vec3(n.hi_x[l], n.hi_y[l], n.hi_z[l]));
    }
    return b;
}

bool HostRayCaster::walk(uint32_t root, Vec3 org, Vec3 dir, float tmin, float* tmax, int level, HostHit* hit) const {
    const float ix = 1.0f / dir.x, iy = 1.0f / dir.y, iz = 1.0f / dir.z;
    const uint32_t pos[3] = {dir.x >= 0 ? 1u : 0u, dir.y >= 0 ? 1u : 0u, dir.z >= 0 ? 1u : 0u};
    uint32_t stack[64];
    int sp = 0;
    bool found = false;
    stack[sp++] = root;
    while (sp > 0) {
        const SlrGpuBvhNode& n = m_scene.nodes[stack[--sp]];
        uint32_t mask = 0;
        for (int l = 0; l < 4; ++l) {
            float tn = tmin, tf = *tmax;
            tn = maxps(((ix > 0.0f ? n.lo_x[l] : n.hi_x[l]) - org.x) * ix, tn);
            tn = maxps(((iy > 0.0f ? n.lo_y[l] : n.hi_y[l]) - org.y) * iy, tn);
            tn = maxps(((iz > 0.0f ? n.lo_z[l] : n.hi_z[l]) - org.z) * iz, tn);
            tf = minps(((ix > 0.0f ? n.hi_x[l] : n.lo_x[l]) - org.x) * ix, tf);
            tf = minps(((iy > 0.0f ? n.hi_y[l] : n.lo_y[l]) - org.y) * iy, tf);
            tf = minps(((iz > 0.0f ? n.hi_z[l] : n.lo_z[l]) - org.z) * iz, tf);
            if (tn <= tf) mask |= 1u << l;
        }
        if (!mask) continue;
        const uint32_t order = kOrder[4 * pos[n.top_axis] + 2 * pos[n.left_axis] + pos[n.right_axis]];
        uint32_t kids[4];
        for (int i = 0; i < 4; ++i) {
            const uint32_t lane = (order >> (4 * i)) & 0xFu;
            kids[i] = ((mask >> lane) & 1u) ? n.child[lane] : 0xFFFFFFFFu;
        }
        for (int i = 3; i >= 0; --i) {          // inner children: far to near onto the stack
            if (kids[i] == 0xFFFFFFFFu || (kids[i] >> 31)) continue;
            if (sp >= 64) throw std::runtime_error("host ray cast: traversal stack overflow");
            stack[sp++] = kids[i] & 0x07FFFFFFu;
        }
        for (int i = 0; i < 4; ++i) {           // leaf children: tested at once, near to far
            if (kids[i] == 0xFFFFFFFFu || !(kids[i] >> 31)) continue;
            const uint32_t first = kids[i] & 0x07FFFFFFu, count = (kids[i] >> 27) & 0xFu;
            for (uint32_t j = 0; j < count; ++j) {
                const SlrGpuLeafRecord& rec = m_scene.leaves[first + j];
                uint32_t id;
                std::memcpy(&id, &rec.a[3], 4);
                if (id & 0x80000000u) {
                    if (level >= 1) continue;
                    const uint32_t instId = id & 0x7FFFFFFFu;
                    const SlrGpuInstance& in = m_scene.instances[instId];
                    float localMax = *tmax;
                    if (walk(in.root_node, mulPoint(in.mat_inv, org), mulVector(in.mat_inv, dir), tmin, &localMax, level + 1, hit)) {
                        *tmax = localMax; hit->inst = instId; found = true;
                    }
                    continue;
                }
                // Moller-Trumbore with the reference's operation order; the record holds v0, v1 - v0, v2 - v0
                const float* v0 = rec.a; const float* e1 = rec.b; const float* e2 = rec.c;
                const float px = dir.y * e2[2] - dir.z * e2[1], py = dir.z * e2[0] - dir.x * e2[2], pz = dir.x * e2[1] - dir.y * e2[0];
                const float det = e1[0] * px + e1[1] * py + e1[2] * pz;
                if (det == 0.0f) continue;
                const float invDet = 1.0f / det;
                const float dx = org.x - v0[0], dy = org.y - v0[1], dz = org.z - v0[2];
                const float b1 = (dx * px + dy * py + dz * pz) * invDet;
                if (b1 < 0.0f || b1 > 1.0f) continue;
                const float qx = dy * e1[2] - dz * e1[1], qy = dz * e1[0] - dx * e1[2], qz = dx * e1[1] - dy * e1[0];
                const float b2 = (dir.x * qx + dir.y * qy + dir.z * qz) * invDet;
                if (b2 < 0.0f || b1 + b2 > 1.0f) continue;
                const float tt = (e2[0] * qx + e2[1] * qy + e2[2] * qz) * invDet;
                if (tt < tmin || tt > *tmax) continue;
                *tmax = tt;
                hit->prim = id; hit->inst = SLRGPU_INVALID_ID; hit->t = tt; hit->b0 = 1.0f - b1 - b2; hit->b1 = b1;
                found = true;
            }
        }
    }
    return found;
}

bool HostRayCaster::intersect(const Vec3& org, const Vec3& dir, float tmin, float tmax, HostHit* hit) const {
    *hit = HostHit();
    if (m_scene.nodes.empty()) return false;
    float limit = tmax;
    return walk(0, org, dir, tmin, &limit, 0, hit);
}

void HostRayCaster::surfacePoint(const HostHit& hit, const Vec3& org, const Vec3& dir, HostSurfacePoint* sp) const {
    const SlrGpuTriangle& tri = m_scene.triangles[hit.prim];
    if (tri.normal_map != SLRGPU_INVALID_ID)
        throw std::runtime_error("host ray cast: the hit surface has a normal map (its textures are evaluated on the GPU only)");
    const SlrGpuVertex* v[3] = {&m_scene.vertices[tri.v[0]], &m_scene.vertices[tri.v[1]], &m_scene.vertices[tri.v[2]]};
    auto P = [&](int i) { return Vec3(v[i]->position[0], v[i]->position[1], v[i]->position[2]); };
    auto N = [&](int i) { return Vec3(v[i]->normal[0], v[i]->normal[1], v[i]->normal[2]); };
    auto T = [&](int i) { return Vec3(v[i]->tangent[0], v[i]->tangent[1], v[i]->tangent[2]); };
    const float b0 = hit.b0, b1 = hit.b1, b2 = 1.0f - b0 - b1;
    sp->gNormal = normalize(cross(P(1) - P(0), P(2) - P(0)));
    sp->sz = normalize(b0 * N(0) + b1 * N(1) + b2 * N(2));
    sp->sx = normalize(b0 * T(0) + b1 * T(1) + b2 * T(2));
    const float dotNT = dot(sp->sz, sp->sx);
    if (std::fabs(dotNT) >= 0.01f) sp->sx = normalize(sp->sx - dotNT * sp->sz);
    sp->sy = cross(sp->sz, sp->sx);
    if (hit.inst == SLRGPU_INVALID_ID) {
        sp->p = org + dir * hit.t;
        return;
    }
    // operator*(StaticTransform, SurfacePoint) (geometry.cpp:63-78)
    const SlrGpuInstance& in = m_scene.instances[hit.inst];
    const Vec3 lo = mulPoint(in.mat_inv, org), ld = mulVector(in.mat_inv, dir);
    sp->p = mulPoint(in.mat, lo + ld * hit.t);
    sp->gNormal = normalize(mulNormal(in.mat_inv, sp->gNormal));
    sp->sx = normalize(mulVector(in.mat, sp->sx));
    sp->sy = normalize(mulVector(in.mat, sp->sy));
    sp->sz = normalize(mulVector(in.mat, sp->sz));
}

}  // namespace slr
