#include "bvh.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <future>
#include <stdexcept>
#include <thread>

namespace slr {

void PrimitiveSet::addTriangle(const Vec3& a, const Vec3& b, const Vec3& c) {
    Prim pr;
    pr.bounds = BBox(a).grow(b).grow(c);
    pr.cost = 1.0f;
    pr.isTriangle = true;
    pr.p[0] = a; pr.p[1] = b; pr.p[2] = c;
    prims.push_back(pr);
}

void PrimitiveSet::addBox(const BBox& b, float cost) {
    Prim pr;
    pr.bounds = b;
    pr.cost = cost;
    pr.isTriangle = false;
    prims.push_back(pr);
}

namespace {

// ---- clipping a primitive's bounds against axis-aligned planes (what the spatial splits of the SBVH need) ----------------
// The reference does this in Triangle::choppedBounds / splitBounds (libSLR/Surface/TriangleMesh.cpp:19-125) and, for
// opaque boxes, SurfaceObject.h:58-86. Trees must come out bit-identical, which pins the ARITHMETIC of a cut point -- the
// edge is taken from its lower to its upper end along the axis, t = (c - a) / (b - a), point = (1 - t) a + t b -- and the
// boundary conventions (an end point exactly on the plane counts as above it). The organisation here is this file's own:
// the triangle's corners ordered along the axis, a table of its three edges, one `pierce` primitive, and the bounds
// assembled as "corners inside the region + points where edges pierce its planes" (min / max are order independent).

struct OrderedCorners {
    Vec3 v[3];            // the triangle's corners, ascending along `axis` (std::sort on the same keys as the reference: ties alike)
    OrderedCorners(const PrimitiveSet::Prim& pr, Axis axis) {
        v[0] = pr.p[0]; v[1] = pr.p[1]; v[2] = pr.p[2];
        std::sort(v, v + 3, [axis](const Vec3& a, const Vec3& b) { return a[axis] < b[axis]; });
    }
};
constexpr int kEdge[3][2] = {{0, 1}, {0, 2}, {1, 2}};     // (lower end, upper end) in the ordered corners

// Where the edge a -> b (a not above b along the axis) pierces the plane `axis = c`: only if a lies strictly below the plane
// and b on or above it. (c - a > 0 and c - b <= 0 in the reference's form; the same predicate for IEEE floats.)
inline bool pierce(const Vec3& a, const Vec3& b, Axis axis, float c, Vec3* at) {
    const float toPlane = c - a[axis], pastPlane = c - b[axis];
    if (!(toPlane > 0 && pastPlane <= 0)) return false;
    const float t = toPlane / (b[axis] - a[axis]);
    *at = (1 - t) * a + t * b;
    return true;
}

// an opaque box (an instance's bounds) restricted to [lo, hi] along the axis; empty when it misses the region
inline BBox restrictBox(const BBox& base, Axis axis, float lo, float hi) {
    BBox r = base;
    r.lo[axis] = std::max(lo, r.lo[axis]);
    r.hi[axis] = std::min(hi, r.hi[axis]);
    return r;
}

// Bounds of the part of a primitive that lies in the slab [lo, hi) along `axis`.
BBox choppedBounds(const PrimitiveSet::Prim& pr, Axis axis, float lo, float hi) {
    if (!pr.isTriangle) {
        const BBox& base = pr.bounds;
        if (hi < base.lo[axis] || lo > base.hi[axis]) return BBox();
        if (lo < base.lo[axis] && hi > base.hi[axis]) return base;
        return restrictBox(base, axis, lo, hi);
    }
    const OrderedCorners tri(pr, axis);
    const float lowest = tri.v[0][axis], middle = tri.v[1][axis], highest = tri.v[2][axis];
    if (lowest >= hi || highest <= lo) return BBox();             // misses the slab
    if (lowest >= lo && highest <= hi) return pr.bounds;          // lies inside it

    BBox r;
    int pierced = 0;
    for (const auto& e : kEdge)
        for (const float plane : {lo, hi}) {
            Vec3 at;
            if (pierce(tri.v[e[0]], tri.v[e[1]], axis, plane, &at)) { r.grow(at); ++pierced; }
        }
    // corners inside the slab: the middle one by its own coordinate; when only ONE of the two planes cuts the triangle
    // (two edges pierced), the extreme corner on the side of the other plane is inside as well
    if (middle >= lo && middle < hi) r.grow(tri.v[1]);
    if (pierced == 2) r.grow(highest < hi ? tri.v[2] : tri.v[0]);
    return r;
}

// Bounds of the two halves of a primitive cut by the plane `axis = pos`.
void splitBounds(const PrimitiveSet::Prim& pr, Axis axis, float pos, BBox* left, BBox* right) {
    if (!pr.isTriangle) {
        const BBox& base = pr.bounds;
        if (pos < base.lo[axis]) { *left = BBox(); *right = base; return; }
        if (pos > base.hi[axis]) { *left = base; *right = BBox(); return; }
        *left = restrictBox(base, axis, -INFINITY, pos);
        *right = restrictBox(base, axis, pos, INFINITY);
        return;
    }
    const OrderedCorners tri(pr, axis);
    if (pos <= tri.v[0][axis]) { *left = BBox(); *right = pr.bounds; return; }
    if (pos >= tri.v[2][axis]) { *left = pr.bounds; *right = BBox(); return; }
    // the plane passes strictly between the lowest and the highest corner: those two seed the halves, the middle corner
    // joins the side it lies on, and the (two) points where edges pierce the plane belong to both
    *left = BBox(tri.v[0]);
    *right = BBox(tri.v[2]);
    (tri.v[1][axis] < pos ? left : right)->grow(tri.v[1]);
    int pierced = 0;
    for (const auto& e : kEdge) {
        Vec3 at;
        if (pierced < 2 && pierce(tri.v[e[0]], tri.v[e[1]], axis, pos, &at)) { left->grow(at); right->grow(at); ++pierced; }
    }
}

struct Fragment {
    uint32_t prim;
    BBox bbox;
    float cost;
};

constexpr uint32_t kObjectBins = 32;
constexpr uint32_t kSpatialBins = 16;
constexpr float kTraversalCost = 1.2f;
constexpr uint32_t kFragmentBudget = 5;

// What the builder produces for one subtree: its nodes in depth-first pre-order (node 0 = the subtree's root, child
// indices local to the subtree) and its leaf references, leaf by leaf in the same order.
struct SubTree {
    std::vector<SBVHNode> nodes;
    std::vector<uint32_t> refs;
    uint32_t depth = 0;          // deepest level reached, counted from the whole tree's root
    uint64_t added = 0;          // fragments the subtree's spatial splits added
};

// The reference grows ONE fragment array in place (SBVH.h:57-348): an object split partitions its range, a spatial
// split rewrites the range and shifts everything behind it. What a node decides depends only on its own fragments
// and their order, never on where the range sits in the array -- so here every node owns its fragments as a vector
// (same std::partition, same left / right order), which (1) drops the O(N) tail shift per spatial split and (2) makes
// sibling subtrees independent: above kParallelMin fragments the left child is built by another thread and the two
// results are spliced in pre-order. The tree is the reference's, node for node (tests/test_oracle_intersect.py).
constexpr uint32_t kParallelMin = 1u << 15;
std::atomic<int> g_buildThreads{0};

struct Builder {
    const PrimitiveSet& ps;
    const BBox sceneBounds;
    const int maxThreads;

    // SLRHOST_BUILD_THREADS=n caps the builder's threads (1 = the serial recursion; tests compare the two)
    static int threadLimit() {
        if (const char* e = std::getenv("SLRHOST_BUILD_THREADS")) { const int n = std::atoi(e); if (n >= 1) return n; }
        return (int)std::max(1u, std::thread::hardware_concurrency());
    }
    Builder(const PrimitiveSet& p, const BBox& b) : ps(p), sceneBounds(b), maxThreads(threadLimit()) {}

    static uint32_t binOf(uint32_t numBins, float v, float lo, float hi) {
        uint32_t b = (uint32_t)(numBins * ((v - lo) / (hi - lo)));
        return std::min(b, numBins - 1);
    }

    // children of the node `nodeIdx` of `out`: sequentially into `out` itself, or -- big ranges -- side by side
    void buildChildren(std::vector<Fragment>&& lefts, std::vector<Fragment>&& rights, uint32_t depth, Axis axis, const BBox& box,
                       uint32_t nodeIdx, SubTree& out) {
        uint32_t c0, c1;
        const bool parallel = maxThreads > 1 && lefts.size() + rights.size() >= kParallelMin && g_buildThreads.load(std::memory_order_relaxed) < maxThreads;
        if (!parallel) {
            c0 = recurse(std::move(lefts), depth, out);
            c1 = recurse(std::move(rights), depth, out);
        } else {
            SubTree l, r;
            g_buildThreads.fetch_add(1, std::memory_order_relaxed);
            auto job = std::async(std::launch::async, [this, &lefts, depth, &l]() {
                struct Done { ~Done() { g_buildThreads.fetch_sub(1, std::memory_order_relaxed); } } done;
                recurse(std::move(lefts), depth, l);
            });
            try { recurse(std::move(rights), depth, r); } catch (...) { job.wait(); throw; }
            job.get();
            c0 = splice(out, l);
            c1 = splice(out, r);
        }
        SBVHNode& nd = out.nodes[nodeIdx];
        nd.bbox = box; nd.c0 = c0; nd.c1 = c1; nd.axis = axis; nd.firstRef = nd.numRefs = 0;
    }

    // appends a finished subtree to `out`; returns the index its root got
    static uint32_t splice(SubTree& out, SubTree& sub) {
        const uint32_t nodeBase = (uint32_t)out.nodes.size(), refBase = (uint32_t)out.refs.size();
        out.nodes.reserve(out.nodes.size() + sub.nodes.size());
        for (SBVHNode nd : sub.nodes) {
            if (nd.numRefs > 0) nd.firstRef += refBase;
            else { nd.c0 += nodeBase; nd.c1 += nodeBase; }
            out.nodes.push_back(nd);
        }
        out.refs.insert(out.refs.end(), sub.refs.begin(), sub.refs.end());
        out.depth = std::max(out.depth, sub.depth);
        out.added += sub.added;
        sub = SubTree();
        return nodeBase;
    }

    // builds the subtree over `frags` (consumed) at the end of `out`; returns its root's index in `out`
    uint32_t recurse(std::vector<Fragment>&& fragsIn, uint32_t depth, SubTree& out) {
        std::vector<Fragment> frags = std::move(fragsIn);
        const uint32_t nodeIdx = (uint32_t)out.nodes.size();
        out.nodes.emplace_back();
        if (++depth > out.depth) out.depth = depth;
        const uint32_t count = (uint32_t)frags.size();

        BBox box, centroidBox;
        float leafCost = 0.0f;
        for (uint32_t i = 0; i < count; ++i) {
            box.grow(frags[i].bbox);
            centroidBox.grow(frags[i].bbox.centroid());
            leafCost += frags[i].cost;
        }
        const Axis axisO = centroidBox.widestAxis();
        const Axis axisS = box.widestAxis();
        const float areaParent = box.surfaceArea();
        const float cLo = centroidBox.lo[axisO], cHi = centroidBox.hi[axisO];
        const float bLo = box.lo[axisS], bHi = box.hi[axisS];

        auto makeLeaf = [&](uint32_t n) {
            SBVHNode& nd = out.nodes[nodeIdx];
            nd.bbox = box;
            nd.c0 = nd.c1 = 0;
            nd.firstRef = (uint32_t)out.refs.size();
            nd.numRefs = n;
            for (uint32_t i = 0; i < n; ++i) out.refs.push_back(frags[i].prim);
        };
        if (count == 1) { makeLeaf(1); return nodeIdx; }

        // ---- object binning over centroids
        struct ObjBin { BBox box; uint32_t n = 0; float cost = 0.0f; } objBins[kObjectBins];
        uint32_t planeO = 0;
        float bestO = INFINITY;
        if ((cHi - cLo) > 0) {
            for (uint32_t i = 0; i < count; ++i) {
                const Fragment& f = frags[i];
                uint32_t b = binOf(kObjectBins, f.bbox.centerOf(axisO), cLo, cHi);
                ++objBins[b].n;
                objBins[b].cost += f.cost;
                objBins[b].box.grow(f.bbox);
            }
            for (uint32_t i = 0; i < kObjectBins - 1; ++i) {
                BBox l, r;
                float cl = 0.0f, cr = 0.0f;
                for (uint32_t j = 0; j <= i; ++j) { l.grow(objBins[j].box); cl += objBins[j].cost; }
                for (uint32_t j = i + 1; j < kObjectBins; ++j) { r.grow(objBins[j].box); cr += objBins[j].cost; }
                float c = kTraversalCost + (l.surfaceArea() * cl + r.surfaceArea() * cr) / areaParent;
                if (c < bestO) { bestO = c; planeO = i; }
            }
        }

        // ---- how much do the two object-split children overlap? decides whether to try spatial splits
        BBox lO, rO;
        for (uint32_t j = 0; j <= planeO; ++j) lO.grow(objBins[j].box);
        for (uint32_t j = planeO + 1; j < kObjectBins; ++j) rO.grow(objBins[j].box);
        BBox overlap = intersection(lO, rO);
        float overlapArea = 0;
        if (overlap.isValid()) overlapArea = overlap.surfaceArea();

        struct SpBin { BBox box; uint32_t in = 0, outN = 0; float costIn = 0.0f, costOut = 0.0f; } spBins[kSpatialBins];
        const float binWidth = box.width(axisS) / kSpatialBins;
        uint32_t planeS = 0;
        float bestS = INFINITY;
        const float alpha = 1e-5;
        if (overlapArea / sceneBounds.surfaceArea() > alpha) {
            for (uint32_t i = 0; i < count; ++i) {
                const Fragment& f = frags[i];
                uint32_t bIn = binOf(kSpatialBins, f.bbox.lo[axisS], bLo, bHi);
                uint32_t bOut = binOf(kSpatialBins, f.bbox.hi[axisS], bLo, bHi);
                ++spBins[bIn].in;
                ++spBins[bOut].outN;
                spBins[bIn].costIn += f.cost;
                spBins[bOut].costOut += f.cost;
                for (int b = (int)bIn; b <= (int)bOut; ++b) {
                    float lo = b * binWidth + bLo;
                    BBox chopped = choppedBounds(ps.prims[f.prim], axisS, lo, lo + binWidth);
                    spBins[b].box.grow(intersection(chopped, f.bbox));
                }
            }
            for (uint32_t i = 0; i < kSpatialBins - 1; ++i) {
                BBox l, r;
                float cl = 0.0f, cr = 0.0f;
                for (uint32_t j = 0; j <= i; ++j) { l.grow(spBins[j].box); cl += spBins[j].costIn; }
                for (uint32_t j = i + 1; j < kSpatialBins; ++j) { r.grow(spBins[j].box); cr += spBins[j].costOut; }
                float c = kTraversalCost + (l.surfaceArea() * cl + r.surfaceArea() * cr) / areaParent;
                if (c < bestS) { bestS = c; planeS = i; }
            }
        }

        if (leafCost < bestO && leafCost < bestS) { makeLeaf(count); return nodeIdx; }

        if (bestO < bestS) {
            // ---- object partition about the chosen centroid plane
            float pivot = cLo + (cHi - cLo) / kObjectBins * (planeO + 1);
            Fragment* first = frags.data();
            Fragment* mid = std::partition(first, first + count,
                                           [axisO, pivot](const Fragment& f) { return f.bbox.centerOf(axisO) < pivot; });
            const uint32_t split = std::max((uint32_t)(mid - first), 1u);
            std::vector<Fragment> lefts(frags.begin(), frags.begin() + split), rights(frags.begin() + split, frags.end());
            std::vector<Fragment>().swap(frags);
            buildChildren(std::move(lefts), std::move(rights), depth, axisO, box, nodeIdx, out);
            return nodeIdx;
        }

        // ---- spatial partition: references straddling the plane go to both sides with clipped bounds
        uint32_t maxL = 0, maxR = 0;
        for (uint32_t j = 0; j <= planeS; ++j) maxL += spBins[j].in;
        for (uint32_t j = planeS + 1; j < kSpatialBins; ++j) maxR += spBins[j].outN;
        std::vector<Fragment> lefts(maxL), rights(maxR);
        uint32_t nL = 0, nR = 0;
        const float splitPos = (planeS + 1) * binWidth + bLo;
        for (uint32_t i = 0; i < count; ++i) {
            const Fragment& f = frags[i];
            uint32_t bIn = binOf(kSpatialBins, f.bbox.lo[axisS], bLo, bHi);
            uint32_t bOut = binOf(kSpatialBins, f.bbox.hi[axisS], bLo, bHi);
            if (bOut <= planeS) {
                lefts[nL++] = f;
            } else if (bIn > planeS) {
                rights[nR++] = f;
            } else {
                BBox sl, sr;
                splitBounds(ps.prims[f.prim], axisS, splitPos, &sl, &sr);
                Fragment& dl = lefts[nL++];
                Fragment& dr = rights[nR++];
                dl.prim = dr.prim = f.prim;
                dl.cost = dr.cost = f.cost;
                dl.bbox = intersection(sl, f.bbox);
                dr.bbox = intersection(sr, f.bbox);
                if (!sl.isValid()) --nL;
                if (!sr.isValid()) --nR;
            }
        }
        out.added += (uint64_t)(nL + nR) - count;
        lefts.resize(nL);
        rights.resize(nR);
        std::vector<Fragment>().swap(frags);
        buildChildren(std::move(lefts), std::move(rights), depth, axisS, box, nodeIdx, out);
        return nodeIdx;
    }
};

}  // namespace

void SBVH::build(const PrimitiveSet& ps) {
    nodes.clear(); refs.clear(); bounds = BBox(); depth = 0;
    if (ps.prims.empty()) throw std::runtime_error("SBVH::build: empty primitive set");
    const uint32_t n = (uint32_t)ps.prims.size();
    std::vector<Fragment> frags(n);
    for (uint32_t i = 0; i < n; ++i) {
        bounds.grow(ps.prims[i].bounds);
        frags[i].prim = i;
        frags[i].bbox = ps.prims[i].bounds;
        frags[i].cost = ps.prims[i].cost;
    }
    Builder b(ps, bounds);
    SubTree tree;
    b.recurse(std::move(frags), 0, tree);
    if ((uint64_t)n + tree.added > (uint64_t)kFragmentBudget * n)
        throw std::runtime_error("SBVH::build: spatial splits exceeded the 5x fragment budget of the reference (SBVH.h:385)");
    nodes = std::move(tree.nodes);
    refs = std::move(tree.refs);
    depth = tree.depth;
    numFragmentsAdded = (uint32_t)tree.added;

    // SAH cost (traversal 1.2, leaf overhead 0)
    float cInt = 0.0f, cLeaf = 0.0f, cObj = 0.0f;
    for (const SBVHNode& nd : nodes) {
        float sa = nd.bbox.surfaceArea();
        if (nd.numRefs == 0) {
            cInt += sa;
        } else {
            cLeaf += sa;
            float cp = 0.0f;
            for (uint32_t j = 0; j < nd.numRefs; ++j) cp += ps.prims[refs[nd.firstRef + j]].cost;
            cObj += sa * cp;
        }
    }
    float rootSA = nodes[0].bbox.surfaceArea();
    cInt *= 1.2f / rootSA;
    cLeaf *= 0.0f / rootSA;
    cObj /= rootSA;
    cost = cInt + cLeaf + cObj;
}

namespace {

struct Collapser {
    const SBVH& src;
    QBVH& out;
    Collapser(const SBVH& s, QBVH& o) : src(s), out(o) {}

    static uint32_t pack(uint32_t idx, uint32_t numLeaves, bool leaf) {
        if (idx > 0x07FFFFFFu) throw std::runtime_error("QBVH: index exceeds 27 bits");
        return idx | (numLeaves << 27) | (leaf ? 0x80000000u : 0u);
    }

    void setLane(QBVHNode& n, int lane, const BBox* b) {
        n.lo_x[lane] = b ? b->lo.x : INFINITY;  n.lo_y[lane] = b ? b->lo.y : INFINITY;  n.lo_z[lane] = b ? b->lo.z : INFINITY;
        n.hi_x[lane] = b ? b->hi.x : -INFINITY; n.hi_y[lane] = b ? b->hi.y : -INFINITY; n.hi_z[lane] = b ? b->hi.z : -INFINITY;
    }

    uint32_t collapse(uint32_t subtreeRoot, uint32_t depth) {
        const SBVHNode& root = src.nodes[subtreeRoot];
        if (root.numRefs > 0) {
            if (root.numRefs >= 16)
                throw std::runtime_error("QBVH: a leaf holds >= 16 references; the reference's 4-bit numLeaves field cannot represent it");
            uint32_t base = (uint32_t)out.refs.size();
            for (uint32_t i = 0; i < root.numRefs; ++i) out.refs.push_back(src.refs[root.firstRef + i]);
            return pack(base, root.numRefs, true);
        }
        if (++depth > out.depth) out.depth = depth;

        // lanes 0,1 come from the left child (its two children, or itself if it is a leaf -> lane 0),
        // lanes 2,3 from the right child likewise
        const uint32_t kid[2] = {root.c0, root.c1};
        uint32_t laneSrc[4] = {UINT32_MAX, UINT32_MAX, UINT32_MAX, UINT32_MAX};
        uint8_t sideAxis[2] = {Axis_X, Axis_X};
        for (int s = 0; s < 2; ++s) {
            const SBVHNode& k = src.nodes[kid[s]];
            if (k.numRefs == 0) { laneSrc[2 * s] = k.c0; laneSrc[2 * s + 1] = k.c1; sideAxis[s] = k.axis; }
            else                { laneSrc[2 * s] = kid[s]; }
        }
        const uint32_t nodeIdx = (uint32_t)out.nodes.size();
        out.nodes.emplace_back();
        {
            QBVHNode& n = out.nodes.back();
            for (int l = 0; l < 4; ++l) setLane(n, l, laneSrc[l] == UINT32_MAX ? nullptr : &src.nodes[laneSrc[l]].bbox);
            n.topAxis = root.axis; n.leftAxis = sideAxis[0]; n.rightAxis = sideAxis[1]; n.pad0 = 0;
            n.pad[0] = n.pad[1] = n.pad[2] = 0;
            for (int l = 0; l < 4; ++l) n.child[l] = kQBVHEmptyChild;
        }
        uint32_t kids[4] = {kQBVHEmptyChild, kQBVHEmptyChild, kQBVHEmptyChild, kQBVHEmptyChild};
        for (int l = 0; l < 4; ++l)
            if (laneSrc[l] != UINT32_MAX) kids[l] = collapse(laneSrc[l], depth);
        for (int l = 0; l < 4; ++l) out.nodes[nodeIdx].child[l] = kids[l];   // vector may have grown: index, not reference
        return pack(nodeIdx, 0, false);
    }
};

}  // namespace

void QBVH::build(const SBVH& sbvh, const PrimitiveSet& ps) {
    nodes.clear(); refs.clear(); depth = 0;
    bounds = sbvh.bounds;
    Collapser c(sbvh, *this);
    uint32_t rootChild = c.collapse(0, 0);
    if (qbvhChildIsLeaf(rootChild)) {
        // single-leaf tree: wrap it in one node whose lane 0 is the whole scene
        nodes.emplace_back();
        QBVHNode& n = nodes.back();
        c.setLane(n, 0, &sbvh.nodes[0].bbox);
        for (int l = 1; l < 4; ++l) c.setLane(n, l, nullptr);
        n.child[0] = rootChild;
        n.child[1] = n.child[2] = n.child[3] = kQBVHEmptyChild;
        n.topAxis = n.leftAxis = n.rightAxis = Axis_X; n.pad0 = 0;
        n.pad[0] = n.pad[1] = n.pad[2] = 0;
    }

    // SAH cost of the collapsed tree
    float cInt = 0.0f, cObj = 0.0f;
    for (const QBVHNode& n : nodes) {
        BBox nodeBox, kid[4];
        for (int l = 0; l < 4; ++l) {
            if (n.child[l] == kQBVHEmptyChild) continue;
            kid[l] = BBox(Vec3(n.lo_x[l], n.lo_y[l], n.lo_z[l]), Vec3(n.hi_x[l], n.hi_y[l], n.hi_z[l]));
            nodeBox.grow(kid[l]);
        }
        cInt += nodeBox.surfaceArea();
        for (int l = 0; l < 4; ++l) {
            uint32_t ch = n.child[l];
            if (ch == kQBVHEmptyChild) continue;
            float sa = kid[l].surfaceArea();
            if (qbvhChildIsLeaf(ch)) {
                float cp = 0.0f;
                for (uint32_t j = 0; j < qbvhChildNumLeaves(ch); ++j) cp += ps.prims[refs[qbvhChildIdx(ch) + j]].cost;
                cObj += sa * cp;
            }
        }
    }
    float rootSA = bounds.surfaceArea();
    cInt *= 1.2f / rootSA;
    cObj /= rootSA;
    cost = cInt + cObj;
}

}  // namespace slr
