// Software ray traversal of the 4-wide QBVH and Moller-Trumbore triangle test for sm_100a.
//
// B200 has no RT cores; this is the hand-written replacement for
//   QBVH::Node::intersect   libSLR/Accelerator/QBVH.h:55-76   (SSE 4-box slab test)
//   QBVH::intersect         libSLR/Accelerator/QBVH.h:295-337 (ordered stack traversal)
//   Triangle::intersect     libSLR/Surface/TriangleMesh.cpp:131-178
//   TransformedSurfaceObject::intersect  libSLR/Core/SurfaceObject.cpp:307-318
//
// PARITY CONTRACT: hit primitive/instance ids must equal the reference's bit for bit, including
// which primitive wins when two report the same distance. That pins (1) the arithmetic: plain IEEE
// fp32, no FMA contraction -- every operation below is an explicit round-to-nearest intrinsic (__fmul_rn,
// __fadd_rn, __fsub_rn, __fdiv_rn are never fused), so the header gives the same bits in any translation
// unit whatever its -fmad setting; division as 1.0f/x then multiply; (2) the visiting order: children ordered by the sign of the ray direction along the
// node's three split axes, inner children pushed in reverse, leaf children tested immediately in
// order, `t > tmax` rejects so a later primitive at an equal distance replaces an earlier one.
//
// Data movement: a node is 128 B = eight 16-byte loads through the read-only path (LDG.E.128);
// a leaf record is 48 B = three. One thread owns one ray; the 64-entry stack lives in local memory.
#pragma once
#include <cstddef>
#include "device_scene.h"
#include "motion.cuh"
#include "textures.cuh"

namespace slrgpu {

constexpr int kStackSize = 64;            // QBVH.h:299
static_assert(sizeof(SlrGpuInstance) % 16 == 0 && offsetof(SlrGpuInstance, mat_inv) % 16 == 0, "walkStep loads mat_inv as float4");
constexpr uint32_t kEmptyChild = 0xFFFFFFFFu;

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float tmin, tmax;
};

struct Hit {
    uint32_t prim;      // SLRGPU_INVALID_ID = miss
    uint32_t inst;
    float t, u, v;
    uint32_t info;      // the hit leaf record's spare word: what the `surface` stage needs to know about the triangle
                        // (device_scene.h packSurfaceInfo, written into the device copy of the records at scene creation)
};

struct TraversalCounters {
    uint32_t nodes, tris;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
// 4-lane slab test with the outcome applied to the child words: lane l's child if its [tNear, tFar] is non-empty, else
// kEmptyChild. The near / far planes are already selected by the caller (the selection depends only on the sign of the
// ray direction, so the walk fetches lo or hi per axis by ADDRESS instead of fetching both and selecting per lane).
// fmaxf/fminf return the non-NaN operand, which coincides with _mm_max_ps/_mm_min_ps here because a
// NaN can only appear in the freshly computed first operand (0 * inf), never in the running bound.
// (The predicates feed four selects directly; building a 4-bit mask first and testing its bits again cost ~12 instructions.)
__device__ __forceinline__ uint4 slab4Children(const float4 nx, const float4 ny, const float4 nz,
                                               const float4 fx, const float4 fy, const float4 fz,
                                               const Ray& r, float ix, float iy, float iz, const uint4 kids) {
    uint4 c;
#define SLAB_LANE(L)                                                        \
    {                                                                       \
        float tn = r.tmin, tf = r.tmax;                                     \
        tn = fmaxf(__fmul_rn(__fsub_rn(nx.L, r.ox), ix), tn);                                 \
        tn = fmaxf(__fmul_rn(__fsub_rn(ny.L, r.oy), iy), tn);                                 \
        tn = fmaxf(__fmul_rn(__fsub_rn(nz.L, r.oz), iz), tn);                                 \
        tf = fminf(__fmul_rn(__fsub_rn(fx.L, r.ox), ix), tf);                                 \
        tf = fminf(__fmul_rn(__fsub_rn(fy.L, r.oy), iy), tf);                                 \
        tf = fminf(__fmul_rn(__fsub_rn(fz.L, r.oz), iz), tf);                                 \
        c.L = (tn <= tf) ? kids.L : 0xFFFFFFFFu;                            \
    }
    SLAB_LANE(x) SLAB_LANE(y) SLAB_LANE(z) SLAB_LANE(w)
#undef SLAB_LANE
    return c;
}

// a*b - c*d and a*b + c*d + e*f, left to right, every product and sum rounded on its own
__device__ __forceinline__ float cross2(float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d)), __fmul_rn(e, f));
}

// Moller-Trumbore, two-sided, exactly the reference's sequence of operations and comparisons
// (NaNs fall through the range checks the same way because the comparisons are not negated).
__device__ __forceinline__ bool triangleTest(const float4 a, const float4 b, const float4 c, const Ray& r,
                                             float* tOut, float* b0Out, float* b1Out, float* b2Out) {
    const float e1x = b.x, e1y = b.y, e1z = b.z;
    const float e2x = c.x, e2y = c.y, e2z = c.z;
    const float px = cross2(r.dy, e2z, r.dz, e2y);
    const float py = cross2(r.dz, e2x, r.dx, e2z);
    const float pz = cross2(r.dx, e2y, r.dy, e2x);
    const float det = dot3(e1x, px, e1y, py, e1z, pz);
    if (det == 0.0f) return false;
    const float invDet = __frcp_rn(det);
    const float dx = __fsub_rn(r.ox, a.x), dy = __fsub_rn(r.oy, a.y), dz = __fsub_rn(r.oz, a.z);
    const float b1 = __fmul_rn(dot3(dx, px, dy, py, dz, pz), invDet);
    if (b1 < 0.0f || b1 > 1.0f) return false;
    const float qx = cross2(dy, e1z, dz, e1y);
    const float qy = cross2(dz, e1x, dx, e1z);
    const float qz = cross2(dx, e1y, dy, e1x);
    const float b2 = __fmul_rn(dot3(r.dx, qx, r.dy, qy, r.dz, qz), invDet);
    if (b2 < 0.0f || __fadd_rn(b1, b2) > 1.0f) return false;
    const float tt = __fmul_rn(dot3(e2x, qx, e2y, qy, e2z, qz), invDet);
    if (tt < r.tmin || tt > r.tmax) return false;
    *tOut = tt;
    *b0Out = __fsub_rn(__fsub_rn(1.0f, b1), b2);
    *b1Out = b1;
    *b2Out = b2;
    return true;
}

// The alpha-texture reject of Triangle::intersect (TriangleMesh.cpp:160-168): the candidate hit's texture coordinate
// b0 tc0 + b1 tc1 + b2 tc2 (left to right, every product and sum rounded on its own) looked up in the triangle's
// alpha map; a value of exactly 0 means the ray passes through. The reference evaluates the texture with a surface
// point that carries only the texture coordinate (FloatTexture::evaluate(TexCoord2D), Core/textures.h:84-88), so a
// world-position mapping sees the origin here (it sees an indeterminate point there). Only leaf records flagged
// SLRGPU_LEAF_FLAG_ALPHA_TEST come here: kept out of line, cut-out geometry is the exception.
static __device__ __noinline__ bool alphaTestPasses(const DeviceScene& s, uint32_t prim, float b0, float b1, float b2) {
    const SlrGpuTriangle tri = s.triangles[prim];
    if (tri.alpha_map == SLRGPU_INVALID_ID) return true;
    const float4* va = s.vertices + (size_t)tri.v[0] * 3;
    const float4* vb = s.vertices + (size_t)tri.v[1] * 3;
    const float4* vc = s.vertices + (size_t)tri.v[2] * 3;
    const float u0 = __ldg(va).w, v0 = __ldg(va + 1).w, u1 = __ldg(vb).w, v1 = __ldg(vb + 1).w, u2 = __ldg(vc).w, v2 = __ldg(vc + 1).w;
    SurfPt sp;
    sp.p = V3(0, 0, 0); sp.gn = V3(0, 0, 1);
    sp.sf.x = V3(1, 0, 0); sp.sf.y = V3(0, 1, 0); sp.sf.z = V3(0, 0, 1);
    sp.u = b0; sp.v = b1; sp.prim = prim; sp.inst = SLRGPU_INVALID_ID; sp.atInfinity = false;
    sp.tu = dot3(b0, u0, b1, u1, b2, u2);
    sp.tv = dot3(b0, v0, b1, v1, b2, v2);
    return evalFloatTexture(s, tri.alpha_map, sp) != 0.0f;
}

// Matrix4x4 * Point3 with the reference's homogeneous divide rule (Matrix4x4.h:75-81); column-major m
// (the reference's `m[12] * 1.0f` is m[12] bit for bit).
__device__ __forceinline__ void mulPoint(const float* __restrict__ m, float x, float y, float z,
                                         float* ox, float* oy, float* oz) {
    float tx = __fadd_rn(dot3(m[0], x, m[4], y, m[8], z), m[12]);
    float ty = __fadd_rn(dot3(m[1], x, m[5], y, m[9], z), m[13]);
    float tz = __fadd_rn(dot3(m[2], x, m[6], y, m[10], z), m[14]);
    float tw = __fadd_rn(dot3(m[3], x, m[7], y, m[11], z), m[15]);
    if (tw != 1.0f) { float rcp = __frcp_rn(tw); tx = __fmul_rn(tx, rcp); ty = __fmul_rn(ty, rcp); tz = __fmul_rn(tz, rcp); }
    *ox = tx; *oy = ty; *oz = tz;
}
__device__ __forceinline__ void mulVector(const float* __restrict__ m, float x, float y, float z,
                                          float* ox, float* oy, float* oz) {
    *ox = dot3(m[0], x, m[4], y, m[8], z);
    *oy = dot3(m[1], x, m[5], y, m[9], z);
    *oz = dot3(m[2], x, m[6], y, m[10], z);
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative scheduling of many rays over the traversal above ("persistent threads with
// dynamic fetch"): a ray's traversal length varies from a handful to dozens of node visits, so with
// one ray per lane per loop iteration most lanes of a warp wait for its slowest ray (ncu on the
// first version: 7.5 of 32 threads active per instruction). Here a lane that finishes its ray takes
// the next one from the warp's chunk while the other lanes keep walking; chunks of 32..256 rays are
// handed to warps through one global cursor. Per ray the sequence of operations is exactly the
// reference's (same pops, same box tests, same leaf order): only the interleaving ACROSS rays changes.
//
// The top-level walk is a per-lane state machine (one node visit per step); an instance leaf record
// is entered by pushing a return marker (see walkStep), as the reference's recursion does.
// (A "while-while" split into separate node and leaf phases was measured too -- round 1, profiles/ --
// and was slower on both benchmarks: rays here are short, the extra state and ballots cost more than
// the leaf-phase coherence returns. So was a warp-cooperative leaf phase -- the records waiting in all
// lanes numbered by a prefix sum, one record per lane, owners' rays fetched by shuffles, hits taken in
// record order from a shared-memory window: bit-exact, but 2695 -> 2289 Mrays/s on the intersect bench and
// extend 10.9 -> 12.6 ms per C1 frame, profiles/r01_rejected_experiments.md. Round 2 tried it again on the leaner walk,
// adaptively -- only in steps where some lane holds >= 3 (or >= 6) records, the winner per owner through a 64-bit
// shared-memory minimum of (t, ~record number), because ncu showed the per-lane leaf loop of an incoherent wave at 3.8
// iterations per node visit and 4.3 of 32 lanes: bit-exact again, and slower again -- C1 707 / 686 vs 730 Mpaths/s, C5 3189 vs
// 3874 Mrays/s: 72-80 registers instead of 64, and the step's dependent chain grows by the scan, the shuffles and the
// shared-memory round trip, while the leaf loop's time is load latency, not issue slots. profiles/r02_rejected_experiments.md.)
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kMaxChunk = 256;
#ifndef SLR_WALK_STEPS_PER_ROUND
#define SLR_WALK_STEPS_PER_ROUND 1
#endif
// Idle lanes that trigger a refill. A refill stalls the whole warp for the DRAM round trip of the new rays' loads, so it
// pays to refill rarely and many lanes at once: re-swept on the leaner node step (profiles/r02_variant_sweep.md) -- C1 699 /
// 729 / 744 / 750 / 750 / 748 Mpaths/s at 4 / 8 / 12 / 16 / 20 / 24, C4 371 / 386 / 396 / 393 / 395 / 401 (round 1 chose 8
// when a node step cost 420 instructions).
#ifndef SLR_WALK_REFILL_IDLE
#define SLR_WALK_REFILL_IDLE 16
#endif
#ifndef SLR_WALK_REFILL_IDLE_INSTANCED
#define SLR_WALK_REFILL_IDLE_INSTANCED 24
#endif
// (Measured and dropped, round 2: an L2 prefetch -- prefetch.global.L2 -- of every pushed inner child and of the first record
// of every queued leaf child at push time: C5 (10 M triangles, scene larger than L2) 2581 -> 1797 Mrays/s, C4 299 -> 288
// Mpaths/s. The walk already keeps the memory system busy; the extra requests for nodes that are never popped cost more
// than the latency they hide. profiles/r02_rejected_experiments.md)
// 0 = never, 1 = always, 2 = only in the instanced instantiations. Measured (profiles/r02_variant_sweep.md): the batch
// kernels gain 6 % (C5 2557 -> 2708 Mrays/s), the instanced renderer kernels 1-4 % (C4 extend 59.3 -> 58.8, shadow 23.5 ->
// 22.6 ms), the flat renderer kernels lose 3 % on extend (C1 9.43 -> 9.77 ms): intersect.cu sets 1, trace.cu 2.
#ifndef SLR_WALK_DEFER_SINK
#define SLR_WALK_DEFER_SINK 1
#endif
// Long walks amortise the refill check over several steps: the instanced kernels (C4: 23 node visits per ray through two BVH
// levels) 401 -> 413 Mpaths/s at two steps, 418 at three; the batch kernels +0.8 % at two (intersect.cu); the flat renderer kernels' short rays lose
// 4 % (C1 762 -> 729) and keep one step per check.
#ifndef SLR_WALK_STEPS_PER_ROUND_INSTANCED
#define SLR_WALK_STEPS_PER_ROUND_INSTANCED 3
#endif
constexpr int kStepsPerRoundInstanced = SLR_WALK_STEPS_PER_ROUND_INSTANCED;
constexpr int kStepsPerRound = SLR_WALK_STEPS_PER_ROUND;     // node visits between two refill checks (sweep: profiles/r01_variant_sweep.md)
constexpr int kRefillIdleFlat = SLR_WALK_REFILL_IDLE, kRefillIdleInstanced = SLR_WALK_REFILL_IDLE_INSTANCED;

struct WalkState {
    Ray r;
    Hit hit;
    float ix, iy, iz;
    uint32_t pos;
    int sp;
    uint32_t top;                // == stack[sp - 1] while sp > 0: the next pop comes from a register, its successor is
                                 // re-loaded one step ahead (the local-memory load is off the node fetch's critical path)
    bool found;
    float time;                  // the ray's time (motion blur): read only by the general instantiation, at a moving instance
    TraversalCounters cnt0;      // the lane's running counters when this ray started (per-ray counts = difference)
};

__device__ __forceinline__ void walkSetRay(WalkState& w);
__device__ __forceinline__ void walkBegin(WalkState& w, uint32_t* stack) {
    walkSetRay(w);
    w.hit.prim = SLRGPU_INVALID_ID; w.hit.inst = SLRGPU_INVALID_ID; w.hit.t = INFINITY; w.hit.u = 0.0f; w.hit.v = 0.0f; w.hit.info = kSurfaceInfoMiss;
    w.found = false;
    w.sp = 0;
    stack[w.sp++] = 0;      // the top-level root is node 0
    w.top = 0;
}

// ---- the step of the walk; instancing is part of the same state machine ---------------------------
// Entering an instance leaf record transforms the ray into the instance's space, pushes a RETURN marker
// and the nested root; the nested nodes are then ordinary steps of the walk (so they take part in the
// dynamic fetch like top-level nodes). Popping the marker restores the world-space ray and the leaf
// records of the top-level node that were still waiting -- exactly where the reference's recursion
// (TransformedSurfaceObject::intersect, SurfaceObject.cpp:307-318) returns to. One level of instancing.
constexpr uint32_t kReturnMarker = 0xFFFFFFFEu;

struct LeafQueue {
    uint32_t first, count;               // current range of leaf records
    uint32_t pending0, pending1, pending2;   // packed leaf child words still to come, kEmptyChild = none
    __device__ __forceinline__ void clear() { first = 0; count = 0; pending0 = pending1 = pending2 = kEmptyChild; }
    // the next waiting leaf child becomes the current range (a leaf child without records never enters the queue)
    __device__ __forceinline__ void next() {
        const uint32_t c = pending0;
        pending0 = pending1; pending1 = pending2; pending2 = kEmptyChild;
        first = c & 0x07FFFFFFu;
        count = c == kEmptyChild ? 0u : (c >> 27) & 0xFu;
    }
};

#ifndef SLR_WALK_KEEP_WORLD_RCP
#define SLR_WALK_KEEP_WORLD_RCP 1
#endif
struct InstanceWalkState {
    LeafQueue leaves;        // of the level the lane is walking
    LeafQueue saved;         // top-level queue while inside an instance
    float wox, woy, woz, wdx, wdy, wdz;      // world-space ray while inside an instance ...
#if SLR_WALK_KEEP_WORLD_RCP
    float wix, wiy, wiz;                     // ... with its reciprocal direction and sign word (three IEEE reciprocals are
    uint32_t wpos;                           // ~30 instructions: kept, not recomputed, when the walk leaves the instance)
#endif
    uint32_t curInst;        // SLRGPU_INVALID_ID at the top level
};

// pos: direction component >= 0 per axis (child ordering, QBVH.h:309-312) in bits 0-2, repeated in bits 8-10 and 16-18
// so that one AND with a node's three axis masks (kNodeAxisMasks below) answers all three ordering questions; bits
// 24-26: invDir > 0 per axis (which plane is the near one, QBVH.h:68-73) -- the two differ for a component of -0.0
constexpr uint32_t kPosNearX = 0x1000000u, kPosNearY = 0x2000000u, kPosNearZ = 0x4000000u;
__device__ __forceinline__ void walkSetRay(WalkState& w) {
    w.ix = __frcp_rn(w.r.dx); w.iy = __frcp_rn(w.r.dy); w.iz = __frcp_rn(w.r.dz);
    w.pos = ((w.r.dx >= 0.0f ? 1u : 0u) | (w.r.dy >= 0.0f ? 2u : 0u) | (w.r.dz >= 0.0f ? 4u : 0u)) * 0x010101u |
            (w.ix > 0.0f ? kPosNearX : 0u) | (w.iy > 0.0f ? kPosNearY : 0u) | (w.iz > 0.0f ? kPosNearZ : 0u);
}
// The node half of a step: pops one entry -- a node: 4-box test, push the inner children that were hit (far to near),
// queue the leaf children that were hit (near to far) in `leaves`; or the return marker of an instance.
// PREFETCH (the persistent tail kernel only): every inner child a node visit pushes and every leaf child it queues is
// prefetched into L1 at once. The tail runs a handful of paths per SM, each a chain of dependent node fetches at L2
// latency: a pushed child is popped only after its nearer siblings' subtrees, so its fetch overlaps that work. (In the
// throughput kernels the same hint costs more than it hides -- measured, see the note above walkQueue.)
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <bool INSTANCES, bool COUNT, bool PREFETCH = false>
__device__ __forceinline__ void walkNode(const DeviceScene& s, WalkState& w, InstanceWalkState& iw, LeafQueue& leaves, uint32_t* stack,
                                         TraversalCounters& cnt, bool& overflow) {
    Ray& r = w.r;
    const uint32_t entry = w.top;
    if (--w.sp > 0) w.top = stack[w.sp - 1];
    if (INSTANCES && entry == kReturnMarker) {
        r.ox = iw.wox; r.oy = iw.woy; r.oz = iw.woz; r.dx = iw.wdx; r.dy = iw.wdy; r.dz = iw.wdz;   // tmax carries over (SurfaceObject.cpp:314)
#if SLR_WALK_KEEP_WORLD_RCP
        w.ix = iw.wix; w.iy = iw.wiy; w.iz = iw.wiz; w.pos = iw.wpos;
#else
        walkSetRay(w);
#endif
        iw.leaves = iw.saved;
        iw.curInst = SLRGPU_INVALID_ID;
    } else {
        const float4* n = s.nodes + (size_t)entry * 8;
        // lo planes at n+0..2, hi planes at n+3..5: near = lo where invDir > 0, else hi
        const uint32_t ox = (w.pos & kPosNearX) ? 0u : 3u, oy = (w.pos & kPosNearY) ? 0u : 3u, oz = (w.pos & kPosNearZ) ? 0u : 3u;
        const float4 nx = ldg4(n + ox), ny = ldg4(n + 1 + oy), nz = ldg4(n + 2 + oz);
        const float4 fx = ldg4(n + 3 - ox), fy = ldg4(n + 4 - oy), fz = ldg4(n + 5 - oz);
        // the child words are fetched with the planes, not after the box test: one memory latency per step instead of two
        const uint4 kids = __ldg(reinterpret_cast<const uint4*>(n + 6));
        const uint32_t axes = __ldg(reinterpret_cast<const uint32_t*>(n + 7));      // (ptxas sinks this one behind the box test: an L1 hit by then, the line has just arrived)
        // (Measured and dropped, round 2: the node as four 32-byte loads -- LDG.E.256, new on sm_100 -- with the near / far
        // planes picked in registers: C1 extend 8.67 vs 8.73 ms, C4 unchanged, C5 3685 vs 3870 Mrays/s.)
        if (COUNT) ++cnt.nodes;
        // children in storage order (left pair 0 1, right pair 2 3), the lanes whose box was missed emptied
        const uint4 c = slab4Children(nx, ny, nz, fx, fy, fz, r, w.ix, w.iy, w.iz, kids);
        if ((c.x & c.y & c.z & c.w) != kEmptyChild) {
            // Visiting order (OrderTable, QBVH.h:309-312): the pair on the near side of the top split first, inside a pair
            // the child on the near side of the pair's split first -- as three conditional swaps, branch free. (The first
            // version picked the four children through an index table; ptxas compiled the indexed picks into ~220
            // instructions of branches per node visit, more than the box tests themselves.)
            const uint32_t side = w.pos & axes;
            const bool T = (side & 0x0000FFu) != 0, L = (side & 0x00FF00u) != 0, R = (side & 0xFF0000u) != 0;
            const uint32_t a0 = L ? c.x : c.y, a1 = L ? c.y : c.x;
            const uint32_t b0 = R ? c.z : c.w, b1 = R ? c.w : c.z;
            const uint32_t o0 = T ? a0 : b0, o1 = T ? a1 : b1, o2 = T ? b0 : a0, o3 = T ? b1 : a1;
            // inner children (bit 31 clear; the empty word has it set) are pushed far to near
            if (w.sp + 4 <= kStackSize) {
#define SLR_PUSH(o) if ((int32_t)(o) >= 0) { const uint32_t k = (o) & 0x07FFFFFFu; stack[w.sp++] = k; w.top = k; if (PREFETCH) prefetchL1(s.nodes + (size_t)k * 8); }
                SLR_PUSH(o3) SLR_PUSH(o2) SLR_PUSH(o1) SLR_PUSH(o0)
#undef SLR_PUSH
            } else {
                const uint32_t ord[4] = {o0, o1, o2, o3};
                for (int i = 3; i >= 0; --i) {
                    if ((int32_t)ord[i] < 0) continue;
                    if (w.sp >= kStackSize) { overflow = true; continue; }
                    stack[w.sp++] = ord[i] & 0x07FFFFFFu;
                    w.top = ord[i] & 0x07FFFFFFu;
                }
            }
            // leaf children (bit 31 set, not the empty word: as a signed number below -1) that hold records, in visiting
            // order: the first becomes the current range, up to three wait
#define SLR_IS_LEAF(o) ((int32_t)(o) < -1 && ((o) & 0x78000000u) != 0)
            uint32_t q0 = kEmptyChild, q1 = kEmptyChild, q2 = kEmptyChild, q3 = kEmptyChild;
            if (SLR_IS_LEAF(o3)) { q0 = o3; }
            if (SLR_IS_LEAF(o2)) { q1 = q0; q0 = o2; }
            if (SLR_IS_LEAF(o1)) { q2 = q1; q1 = q0; q0 = o1; }
            if (SLR_IS_LEAF(o0)) { q3 = q2; q2 = q1; q1 = q0; q0 = o0; }
#undef SLR_IS_LEAF
            if (q0 != kEmptyChild) {
                if (PREFETCH) prefetchL1(s.leaves + (size_t)(q0 & 0x07FFFFFFu) * 3);
                leaves.first = q0 & 0x07FFFFFFu;
                leaves.count = (q0 >> 27) & 0xFu;
                leaves.pending0 = q1; leaves.pending1 = q2; leaves.pending2 = q3;
            }
        }
    }
}

// One step of a lane: a node visit if no leaf record is waiting, then the leaf records that visit queued -- all of
// them (default), or AT MOST ONE per step with the rest left for the lane's next steps, before it pops another node
// (SLR_WALK_ONE_RECORD_PER_STEP). Per ray the sequence of box and triangle tests is the reference's either way; what
// changes is the interleaving across lanes: in the loop form the lanes that found no leaf wait for the lane with the
// longest list, in the one-record form every lane advances by a node or a triangle per iteration of the warp's loop
// but carries its leaf queue across iterations (5 more registers). Measured (profiles/r01_variant_sweep.md): the
// batch kernels on the 500 k-triangle heightfield gain 7 % from one record per step (3604 -> 3860 Mrays/s), the
// renderer's extend / shadow kernels lose 10 % (C1 extend 10.3 -> 11.4 ms) -- so intersect.cu sets it, trace.cu does not.
// Returns true when the ray is finished.
#ifndef SLR_WALK_ONE_RECORD_PER_STEP
#define SLR_WALK_ONE_RECORD_PER_STEP 0
#endif
template <bool INSTANCES, bool ANY_HIT, bool COUNT, bool ALPHA = false, bool PREFETCH = false>
__device__ __forceinline__ bool walkStep(const DeviceScene& s, WalkState& w, InstanceWalkState& iw, uint32_t* stack,
                                         TraversalCounters& cnt, bool& overflow) {
    Ray& r = w.r;
#if SLR_WALK_ONE_RECORD_PER_STEP
    LeafQueue& leaves = iw.leaves;
    if (leaves.count == 0) walkNode<INSTANCES, COUNT, PREFETCH>(s, w, iw, leaves, stack, cnt, overflow);
    if (leaves.count != 0) {
#else
    LeafQueue local;                     // flat scenes: the queue lives for one step only
    LeafQueue& leaves = INSTANCES ? iw.leaves : local;
    if (!INSTANCES) local.clear();
    walkNode<INSTANCES, COUNT, PREFETCH>(s, w, iw, leaves, stack, cnt, overflow);
    while (leaves.count != 0) {
#endif
        const float4* rec = s.leaves + (size_t)leaves.first * 3;
        const float4 a = ldg4(rec), b = ldg4(rec + 1), cc = ldg4(rec + 2);      // all 48 B at once (an instance record has them too)
        ++leaves.first;
        if (--leaves.count == 0) leaves.next();
        const uint32_t id = __float_as_uint(a.w);
        if (COUNT) ++cnt.tris;
        // (a flat scene holds no instance records -- slrgpu_scene_create refuses one without an instance table -- so the flat
        // instantiations do not test for them: the three loads of a record issue together instead of a.w first)
        // (Measured and dropped for the instanced kernels: running the triangle test before the record is told apart, so
        // that its three loads issue together there too -- C4 extend 50.1 -> 52.1 ms.)
        if (INSTANCES && (id & 0x80000000u)) {
            if constexpr (INSTANCES) {
                // nested instancing is rejected at scene build; no room for marker + root is reported as overflow
                if (iw.curInst == SLRGPU_INVALID_ID) {
                    if (w.sp + 2 > kStackSize) overflow = true;
                    else {
                        const uint32_t instId = id & 0x7FFFFFFFu;
                        const SlrGpuInstance* inst = s.instances + instId;
                        iw.wox = r.ox; iw.woy = r.oy; iw.woz = r.oz; iw.wdx = r.dx; iw.wdy = r.dy; iw.wdz = r.dz;
#if SLR_WALK_KEEP_WORLD_RCP
                        iw.wix = w.ix; iw.wiy = w.iy; iw.wiz = w.iz; iw.wpos = w.pos;
#endif
                        iw.saved = iw.leaves;
                        iw.leaves.clear();
                        float lx, ly, lz, mx, my, mz;
                        if (ALPHA && inst->motion != 0u && s.motions != nullptr) {
                            // a moving instance (general instantiation only): its transform at the ray's time (Transform.cpp:31-35)
                            float xf[32];
                            sampleMotion(s.motions[inst->motion - 1u], inst->mat, inst->mat_inv, w.time, xf, xf + 16);
                            mulPoint(xf + 16, r.ox, r.oy, r.oz, &lx, &ly, &lz);
                            mulVector(xf + 16, r.dx, r.dy, r.dz, &mx, &my, &mz);
                        } else {
                            // the inverse matrix as four 16-byte loads (the instance table is 256-byte aligned on the
                            // device and a record is 160 bytes: mat_inv sits on a 16-byte boundary)
                            float mi[16];
                            const float4* q = reinterpret_cast<const float4*>(inst->mat_inv);
#pragma unroll
                            for (int c = 0; c < 4; ++c) { const float4 col = __ldg(q + c); mi[4 * c] = col.x; mi[4 * c + 1] = col.y; mi[4 * c + 2] = col.z; mi[4 * c + 3] = col.w; }
                            mulPoint(mi, r.ox, r.oy, r.oz, &lx, &ly, &lz);
                            mulVector(mi, r.dx, r.dy, r.dz, &mx, &my, &mz);
                        }
                        r.ox = lx; r.oy = ly; r.oz = lz; r.dx = mx; r.dy = my; r.dz = mz;
                        walkSetRay(w);
                        iw.curInst = instId;
                        stack[w.sp++] = kReturnMarker;
                        stack[w.sp++] = inst->root_node;
                        w.top = inst->root_node;
                        return false;
                    }
                }
            }
        } else {
            float t, b0, b1, b2;
            bool accept = triangleTest(a, b, cc, r, &t, &b0, &b1, &b2);
            if (ALPHA && accept && (__float_as_uint(b.w) & SLRGPU_LEAF_FLAG_ALPHA_TEST)) accept = alphaTestPasses(s, id, b0, b1, b2);
            if (accept) {
                r.tmax = t;
                w.hit.prim = id; w.hit.inst = INSTANCES ? iw.curInst : SLRGPU_INVALID_ID;
                w.hit.t = t; w.hit.u = b0; w.hit.v = b1; w.hit.info = __float_as_uint(cc.w);
                w.found = true;
                if (ANY_HIT) return true;
            }
        }
    }
    return w.sp == 0 && leaves.count == 0;
}

// Runs `n` rays through the scene with one warp-cooperative loop. Source::load(i, Ray&) fetches ray i,
// Sink::done(i, state) consumes its result. *cursor must be 0 at launch.
template <bool INSTANCES, bool ANY_HIT, bool COUNT, bool ALPHA, typename Source, typename Sink>
__device__ __forceinline__ void walkQueue(const DeviceScene& s, uint32_t n, uint32_t* cursor, const Source& source, const Sink& sink,
                                          TraversalCounters& cnt, bool& overflow) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t stack[kStackSize];
    WalkState w;
    InstanceWalkState iw;                       // the leaf queue; the rest is only live in the INSTANCES instantiations
    bool active = false;
    uint32_t idx = 0;
    uint32_t chunkNext = 0, chunkEnd = 0;       // warp-uniform
    bool exhausted = false;                     // warp-uniform
    // chunk size: about half a warp's fair share of the batch, so that every warp of the grid gets work
    // (small batches: one ray per lane, like a plain launch) and the tail of the batch stays short
    const uint32_t totalWarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t chunk = 32;
    while (chunk < kMaxChunk && (unsigned long long)chunk * 2ull * totalWarps <= (unsigned long long)n / 2ull) chunk <<= 1;
    // SLR_WALK_DEFER_SINK: a lane that finishes its ray keeps the result in its registers and hands it to the sink when the
    // warp next services its idle lanes (the refill point), so the sink's code -- the hit record stores of `extend`, the
    // contribution loads + four 16-byte reductions of `shadow` -- runs once for all lanes that finished since, instead of
    // once per finishing lane at 1-3 active lanes.
    bool finished = false;
    // (the any-hit walk of `shadow` is short and likes rare refills too: C1 shadow 3.98 -> 3.90 ms at the instanced threshold)
    constexpr int kRefillIdle = (INSTANCES || ANY_HIT) ? kRefillIdleInstanced : kRefillIdleFlat;
    while (true) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        const int numIdle = __popc(idle);
        constexpr bool kDefer = SLR_WALK_DEFER_SINK == 1 || (SLR_WALK_DEFER_SINK == 2 && INSTANCES);
        if (kDefer && (numIdle >= kRefillIdle || exhausted) && finished) { sink.done(idx, w, cnt); finished = false; }
        // refill when a quarter of the warp is idle (or nothing is running)
        if (!exhausted && (numIdle >= kRefillIdle)) {
            if (chunkNext >= chunkEnd) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, chunk);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                chunkNext = base;
                chunkEnd = base < n ? min(base + chunk, n) : base;
                if (base >= n) exhausted = true;
            }
            if (!exhausted) {
                const uint32_t avail = chunkEnd - chunkNext;
                const uint32_t rank = __popc(idle & lt);
                if (!active && rank < avail) {
                    idx = chunkNext + rank;
                    source.load(idx, w.r);
                    w.time = ALPHA ? source.time(idx) : 0.0f;
                    walkBegin(w, stack);
                    iw.leaves.clear();
                    if (INSTANCES) { iw.saved.clear(); iw.curInst = SLRGPU_INVALID_ID; }
                    w.cnt0 = cnt;
                    active = true;
                }
                chunkNext += min((uint32_t)numIdle, avail);
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (exhausted) break;
            continue;
        }
#pragma unroll 1
        for (int it = 0; it < (INSTANCES ? kStepsPerRoundInstanced : kStepsPerRound); ++it) {
            if (active) {
                if (walkStep<INSTANCES, ANY_HIT, COUNT, ALPHA>(s, w, iw, stack, cnt, overflow)) {
                    if (kDefer) finished = true;
                    else sink.done(idx, w, cnt);
                    active = false;
                }
            }
        }
    }
}

}  // namespace slrgpu
