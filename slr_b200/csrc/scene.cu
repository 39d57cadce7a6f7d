// Scene upload: copies the caller's SoA host buffers into HBM once and keeps a device view.
// Replaces, on the GPU side, the object graph a reference renderer receives from
// SLRSceneGraph::Scene::build (libSLRSceneGraph/Scene.cpp:28-44).
#include "device_scene.h"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace slrgpu {

static thread_local char g_error[512] = "";

void setError(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cudaFail(cudaError_t e, const char* what) {
    setError("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    if (e == cudaErrorMemoryAllocation) return SLRGPU_ERR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SLRGPU_ERR_NO_DEVICE;
    return SLRGPU_ERR_CUDA;
}

// A destroyed scene's arena (up to kArenaCacheLimit bytes) is kept, one per device, for the next scene
// that fits: a renderer front end that uploads the scene on every render() call then runs without any
// cudaMalloc / cudaFree (both are synchronising driver calls with occasional 100 ms outliers on a busy
// host). slrgpu_release_workspaces() drops the cache.
static std::mutex g_arenaMutex;
static void* g_arena[64] = {};
static uint64_t g_arenaBytes[64] = {};
constexpr uint64_t kArenaCacheLimit = 256ull << 20;

static void* takeCachedArena(int device, uint64_t need, uint64_t* got) {
    std::lock_guard<std::mutex> lock(g_arenaMutex);
    if (device < 0 || device >= 64 || !g_arena[device] || g_arenaBytes[device] < need) return nullptr;
    void* p = g_arena[device];
    *got = g_arenaBytes[device];
    g_arena[device] = nullptr; g_arenaBytes[device] = 0;
    return p;
}
static bool cacheArena(int device, void* p, uint64_t bytes) {
    std::lock_guard<std::mutex> lock(g_arenaMutex);
    if (device < 0 || device >= 64 || bytes > kArenaCacheLimit) return false;
    if (g_arena[device]) { if (g_arenaBytes[device] >= bytes) return false; cudaFree(g_arena[device]); }
    g_arena[device] = p; g_arenaBytes[device] = bytes;
    return true;
}
void releaseSceneArenas() {
    std::lock_guard<std::mutex> lock(g_arenaMutex);
    for (int d = 0; d < 64; ++d) if (g_arena[d]) { cudaSetDevice(d); cudaFree(g_arena[d]); g_arena[d] = nullptr; g_arenaBytes[d] = 0; }
}

// All scene buffers live in ONE device allocation (256-byte aligned sub-ranges): a scene is created and
// destroyed with one cudaMalloc / cudaFree, and a small scene goes up in one staged copy.
struct UploadPlan {
    struct Item { const void* src; uint64_t bytes; uint64_t offset; const void** dst; };
    std::vector<Item> items;
    uint64_t total = 0;
    template <typename T> void add(const T* src, uint64_t count, const T** dst) {
        *dst = nullptr;
        if (count == 0 || src == nullptr) return;
        const uint64_t bytes = count * sizeof(T);
        items.push_back(Item{src, bytes, total, reinterpret_cast<const void**>(dst)});
        total += (bytes + 255u) & ~uint64_t(255);
    }
    int commit(SlrGpuScene* sc) {
        if (total == 0) return SLRGPU_OK;
        void* base = takeCachedArena(sc->device, total, &sc->arenaBytes);
        if (!base) {
            SLRGPU_CUDA_TRY(cudaMalloc(&base, total));
            sc->arenaBytes = total;
        }
        sc->allocations[sc->numAllocations++] = base;
        sc->deviceBytes += total;
        if (total <= (64u << 20)) {
            std::vector<uint8_t> stage(total, 0);
            for (const Item& it : items) memcpy(stage.data() + it.offset, it.src, it.bytes);
            SLRGPU_CUDA_TRY(cudaMemcpy(base, stage.data(), total, cudaMemcpyHostToDevice));
        } else {
            for (const Item& it : items)
                SLRGPU_CUDA_TRY(cudaMemcpy(static_cast<uint8_t*>(base) + it.offset, it.src, it.bytes, cudaMemcpyHostToDevice));
        }
        for (const Item& it : items) *it.dst = static_cast<uint8_t*>(base) + it.offset;
        return SLRGPU_OK;
    }
};

// ---------------------------------------------------------------------------------------------
// Spectrum compilation (host, once per scene): rewrites the caller's spectrum table into the two
// forms the device evaluates with a couple of independent loads per wavelength.
//   UPSAMPLED (u, v, scale)  -> REGULAR, 95 samples on [360, 830] nm: the Meng-Simon evaluation
//       (SpectrumTypes.h:239-339) is linear in the 3-4 data-point spectra it blends, so blending them
//       once per spectrum and interpolating the blend is the same piecewise-linear function.
//   IRREGULAR (n knots)      -> the knots that can bracket a wavelength in [360, 830] (the tables of
//       spectrum_library.cpp run from 200 nm to 12 um) plus a 2 nm bin -> first-knot look-up table, so
//       the lower_bound of SpectrumTypes.h:141-160 becomes a table read and at most a step or two.
// Layout of a compiled IRREGULAR spectrum in spectrum_data: lambdas[n], values[n], lut (kLutBins bytes
// packed in (kLutBins + 3) / 4 words); num_samples = n.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kUpGridW = 12, kUpGridH = 14, kUpNumWl = 95, kUpPointStride = 99;
constexpr float kWlLow = 360.0f, kWlHigh = 830.0f;

// host twin of spectral.cuh upsampleWeights
int upsampleWeightsHost(const float* grid, const float* points, float u, float v, uint32_t idx[4], float w[4]) {
    if (u < 0.0f || u >= kUpGridW || v < 0.0f || v >= kUpGridH) return 0;
    const int ui = (int)u, vi = (int)v;
    const float* cell = grid + (ui + kUpGridW * vi) * 8;
    const bool inside = cell[0] != 0.0f;
    const int numPoints = (int)cell[1];
    if (inside) {
        const float sx = u - ui, ty = v - vi;
        w[0] = (1 - sx) * (1 - ty); w[1] = sx * (1 - ty); w[2] = (1 - sx) * ty; w[3] = sx * ty;
        for (int i = 0; i < 4; ++i) idx[i] = (uint32_t)cell[2 + i];
        return 4;
    }
    const uint32_t i0 = (uint32_t)cell[2];
    const float p0u = points[i0 * kUpPointStride + 2], p0v = points[i0 * kUpPointStride + 3];
    const float ex = u - p0u, ey = v - p0v;
    const uint32_t i1 = (uint32_t)cell[3];
    float e0x = points[i1 * kUpPointStride + 2] - p0u, e0y = points[i1 * kUpPointStride + 3] - p0v;
    float uu = e0x * ey - ex * e0y;
    for (int i = 1; i < numPoints; ++i) {
        const uint32_t id = (uint32_t)cell[2 + (i % (numPoints - 1) + 1)];
        const float e1x = points[id * kUpPointStride + 2] - p0u, e1y = points[id * kUpPointStride + 3] - p0v;
        const float vv = ex * e1y - e1x * ey;
        const float area = e0x * e1y - e1x * e0y;
        const float bu = uu / area, bv = vv / area, bw = 1.0f - bu - bv;
        if (bu < -1e-6 || bv < -1e-6 || bw < -1e-6) { uu = -vv; e0x = e1x; e0y = e1y; continue; }
        w[0] = bu; w[1] = bv; w[2] = bw;
        idx[0] = id; idx[1] = (uint32_t)cell[2 + i]; idx[2] = i0;
        return 3;
    }
    return 0;
}
}  // namespace

static void compileSpectra(const SlrGpuSceneDesc* d, std::vector<SlrGpuSpectrum>& spectra, std::vector<float>& data) {
    spectra.assign(d->spectra, d->spectra + d->num_spectra);
    data.assign(d->spectrum_data, d->spectrum_data + d->num_spectrum_floats);
    for (SlrGpuSpectrum& sp : spectra) {
        if (sp.kind == SLRGPU_SPECTRUM_UPSAMPLED && d->spectral.upsample_grid && d->spectral.upsample_points) {
            uint32_t idx[4];
            float w[4];
            const int n = upsampleWeightsHost(d->spectral.upsample_grid, d->spectral.upsample_points, sp.p0, sp.p1, idx, w);
            const uint32_t off = (uint32_t)data.size();
            for (int b = 0; b < kUpNumWl; ++b) {
                float ret = 0.0f;
                for (int j = 0; j < n; ++j) ret += w[j] * d->spectral.upsample_points[idx[j] * kUpPointStride + 4 + b];
                data.push_back(ret * sp.p2);
            }
            sp.kind = SLRGPU_SPECTRUM_REGULAR; sp.data_offset = off; sp.num_samples = kUpNumWl; sp.p0 = kWlLow; sp.p1 = kWlHigh; sp.p2 = 0.0f;
        } else if (sp.kind == SLRGPU_SPECTRUM_IRREGULAR && sp.num_samples >= 2) {
            const uint32_t n = sp.num_samples;
            const std::vector<float> lam(data.begin() + sp.data_offset, data.begin() + sp.data_offset + n);
            const std::vector<float> val(data.begin() + sp.data_offset + n, data.begin() + sp.data_offset + 2 * n);
            // keep knots [first, last]: first = last knot <= 360 (or 0), last = first knot >= 830 (or n - 1)
            uint32_t first = 0, last = n - 1;
            for (uint32_t i = 0; i < n; ++i) if (lam[i] <= kWlLow) first = i;
            for (uint32_t i = n; i-- > 0;) if (lam[i] >= kWlHigh) last = i;
            // a clipped table must clamp like the full one outside its range: an end is only clipped when the
            // visible range is bracketed on that side
            if (!(lam[first] <= kWlLow)) first = 0;
            if (!(lam[last] >= kWlHigh)) last = n - 1;
            const uint32_t m = last - first + 1;
            const uint32_t off = (uint32_t)data.size();
            for (uint32_t i = first; i <= last; ++i) data.push_back(lam[i]);
            for (uint32_t i = first; i <= last; ++i) data.push_back(val[i]);
            // lut[b] = lower-interval index for wavelengths in bin b = [360 + 2 b, 360 + 2 (b + 1)): the last knot
            // index i with lambda[i] < binStart (as max(lower_bound - 1, 0) would give at the bin start)
            std::vector<uint32_t> words((SLRGPU_SPECTRUM_LUT_BINS + 3) / 4, 0u);
            for (int b = 0; b < SLRGPU_SPECTRUM_LUT_BINS; ++b) {
                const float binStart = kWlLow + 2.0f * b;
                uint32_t lo = 0;
                while (lo < m && data[off + lo] < binStart) ++lo;            // lower_bound
                const uint32_t lowIdx = lo > 0 ? lo - 1 : 0;
                words[b >> 2] |= (lowIdx > 255u ? 255u : lowIdx) << (8 * (b & 3));
            }
            for (uint32_t wv : words) { float f; memcpy(&f, &wv, 4); data.push_back(f); }
            sp.data_offset = off; sp.num_samples = m; sp.kind = SLRGPU_SPECTRUM_IRREGULAR_LUT;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// One host pass over the caller's tables before anything is uploaded: every index the device (or compileSpectra)
// dereferences is range-checked here, so a malformed description is SLRGPU_ERR_INVALID_ARGUMENT instead of a GPU
// fault, and what the device code cannot represent (material trees with more than 4 leaf lobes or deeper than its
// 8-entry stack, emitter wrappers nested deeper than classifyMaterialIn peels) is SLRGPU_ERR_UNSUPPORTED instead of
// a silently different image. Also derives which material-class kernels the scene needs from the classification the
// device itself uses (classifyMaterialIn over the triangles' materials), and whether any leaf record asks for the
// alpha test.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxLobes = 4, kMaterialStack = 8, kMaxEmitterNesting = 4;

struct Validator {
    const SlrGpuSceneDesc* d;
    int fail(int code, const char* fmt, ...) {
        char buf[400];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        setError("slrgpu_scene_create: %s", buf);
        return code;
    }
    bool texOk(uint32_t t, bool optional) const { return t < d->num_textures || (optional && t == SLRGPU_INVALID_ID); }

    // leaf lobes and DFS stack depth of a material tree as buildBsdf walks it; -1: too deep / cyclic
    int countLobes(uint32_t id, int depth, int* maxStack, int stackNow) const {
        if (id == SLRGPU_INVALID_ID) return 0;
        if (id >= d->num_materials || depth > 16) return -1;
        const SlrGpuMaterial& m = d->materials[id];
        if (stackNow > *maxStack) *maxStack = stackNow;
        switch (m.kind) {
        case SLRGPU_MAT_EMITTER: case SLRGPU_MAT_INVERSE: return countLobes(m.sub[0], depth + 1, maxStack, stackNow);
        case SLRGPU_MAT_SUMMED: case SLRGPU_MAT_MIXED: {
            // sub[0] is walked with sub[1] still waiting on the stack
            const int a = countLobes(m.sub[0], depth + 1, maxStack, stackNow + 1);
            const int b = countLobes(m.sub[1], depth + 1, maxStack, stackNow);
            return (a < 0 || b < 0) ? -1 : a + b;
        }
        default: return 1;
        }
    }

    int materials() {
        for (uint32_t i = 0; i < d->num_materials; ++i) {
            const SlrGpuMaterial& m = d->materials[i];
            auto needTex = [&](int k, bool optional = false) { return texOk(m.tex[k], optional); };
            auto subOk = [&](int k, bool optional) { return m.sub[k] < d->num_materials || (optional && m.sub[k] == SLRGPU_INVALID_ID); };
            bool ok = true;
            switch (m.kind) {
            case SLRGPU_MAT_DIFFUSE: ok = needTex(0) && needTex(1, true); break;
            case SLRGPU_MAT_SPECULAR_REFLECTION: case SLRGPU_MAT_SPECULAR_SCATTERING: case SLRGPU_MAT_WARD_DUR:
            case SLRGPU_MAT_MICROFACET_REFLECTION: case SLRGPU_MAT_MICROFACET_SCATTERING: ok = needTex(0) && needTex(1) && needTex(2); break;
            case SLRGPU_MAT_ASHIKHMIN_SHIRLEY: ok = needTex(0) && needTex(1) && needTex(2) && needTex(3); break;
            case SLRGPU_MAT_INVERSE: ok = subOk(0, false); break;
            case SLRGPU_MAT_SUMMED: ok = subOk(0, false) && subOk(1, false); break;
            case SLRGPU_MAT_MIXED: ok = subOk(0, false) && subOk(1, false) && needTex(0); break;
            case SLRGPU_MAT_EMITTER: ok = subOk(0, true) && subOk(1, false); break;
            case SLRGPU_MAT_DIFFUSE_EMISSION: case SLRGPU_MAT_IBL_EMISSION: ok = needTex(0); break;
            default: return fail(SLRGPU_ERR_INVALID_ARGUMENT, "material %u has unknown kind %u", i, m.kind);
            }
            if (!ok) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "material %u (kind %u) refers to a texture / sub-material out of range", i, m.kind);
            if (m.kind == SLRGPU_MAT_EMITTER) {
                const uint32_t ek = d->materials[m.sub[1]].kind;
                if (ek != SLRGPU_MAT_DIFFUSE_EMISSION && ek != SLRGPU_MAT_IBL_EMISSION)
                    return fail(SLRGPU_ERR_INVALID_ARGUMENT, "material %u: emitter property %u is not an emission material", i, m.sub[1]);
                int nest = 0;
                for (uint32_t id = i; id != SLRGPU_INVALID_ID && id < d->num_materials && d->materials[id].kind == SLRGPU_MAT_EMITTER; id = d->materials[id].sub[0])
                    if (++nest > kMaxEmitterNesting) return fail(SLRGPU_ERR_UNSUPPORTED, "material %u: emitter wrappers nested deeper than %d", i, kMaxEmitterNesting);
            }
        }
        return SLRGPU_OK;
    }

    int textures() {
        for (uint32_t i = 0; i < d->num_textures; ++i) {
            const SlrGpuTexture& t = d->textures[i];
            bool ok = true;
            switch (t.kind) {
            case SLRGPU_TEX_CONSTANT_SPECTRUM: ok = t.i0 < d->num_spectra; break;
            case SLRGPU_TEX_CHECKER_SPECTRUM: ok = t.i0 < d->num_spectra && t.i1 < d->num_spectra; break;
            case SLRGPU_TEX_IMAGE_SPECTRUM: case SLRGPU_TEX_IMAGE_NORMAL: case SLRGPU_TEX_IMAGE_FLOAT: ok = t.i0 < d->num_images; break;
            case SLRGPU_TEX_CONSTANT_FLOAT: case SLRGPU_TEX_CHECKER_NORMAL: case SLRGPU_TEX_CHECKER_FLOAT:
            case SLRGPU_TEX_VORONOI_SPECTRUM: case SLRGPU_TEX_VORONOI_NORMAL: case SLRGPU_TEX_VORONOI_FLOAT: break;
            default: return fail(SLRGPU_ERR_INVALID_ARGUMENT, "texture %u has unknown kind %u", i, t.kind);
            }
            if (!ok) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "texture %u (kind %u) refers to a spectrum / image out of range", i, t.kind);
            if (t.mapping > SLRGPU_MAP_WORLD_POS) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "texture %u has unknown mapping %u", i, t.mapping);
        }
        for (uint32_t i = 0; i < d->num_spectra; ++i) {
            const SlrGpuSpectrum& sp = d->spectra[i];
            uint64_t need = 0;
            if (sp.kind == SLRGPU_SPECTRUM_REGULAR) need = sp.num_samples;
            else if (sp.kind == SLRGPU_SPECTRUM_IRREGULAR) need = 2ull * sp.num_samples;
            else if (sp.kind != SLRGPU_SPECTRUM_UPSAMPLED && sp.kind != SLRGPU_SPECTRUM_RGB)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "spectrum %u has unknown kind %u", i, sp.kind);
            if (need && ((uint64_t)sp.data_offset + need > d->num_spectrum_floats || !d->spectrum_data))
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "spectrum %u: samples [%u, +%llu) exceed spectrum_data (%u floats)", i, sp.data_offset,
                            (unsigned long long)need, d->num_spectrum_floats);
            if ((sp.kind == SLRGPU_SPECTRUM_REGULAR || sp.kind == SLRGPU_SPECTRUM_IRREGULAR) && sp.num_samples < 2)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "spectrum %u: a sampled spectrum needs at least 2 samples", i);
        }
        static const uint32_t texelBytes[] = {3, 4, 4, 8, 1, 6, 8, 4};
        for (uint32_t i = 0; i < d->num_images; ++i) {
            const SlrGpuImage& im = d->images[i];
            if (im.format > SLRGPU_IMG_FLOAT32 || im.width == 0 || im.height == 0)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "image %u: unknown format %u or empty size", i, im.format);
            const uint64_t bytes = (uint64_t)im.width * im.height * texelBytes[im.format];
            if (!d->image_data || im.data_offset + bytes > d->image_data_bytes)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "image %u: texels exceed image_data", i);
        }
        return SLRGPU_OK;
    }

    int geometry(bool* hasAlpha, uint32_t* classMask) {
        // a scene without materials is geometry only (intersect entry points): its triangle table, if any, is not read
        const bool shaded = d->triangles && d->num_triangles && d->materials && d->num_materials;
        if (shaded && (!d->vertices || !d->num_vertices)) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "triangles without vertices");
        for (uint32_t i = 0; i < d->num_bvh_nodes; ++i) {
            const SlrGpuBvhNode& n = d->bvh_nodes[i];
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = n.child[k];
                if (c == 0xFFFFFFFFu) continue;
                const uint32_t idx = c & 0x07FFFFFFu, cnt = (c >> 27) & 0xFu;
                if (c >> 31) { if ((uint64_t)idx + cnt > d->num_leaf_records) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "node %u child %d: leaf records [%u, +%u) out of range", i, k, idx, cnt); }
                else if (idx >= d->num_bvh_nodes) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "node %u child %d: node index %u out of range", i, k, idx);
            }
            if (n.top_axis > 2 || n.left_axis > 2 || n.right_axis > 2) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "node %u: split axis out of range", i);
        }
        if (d->sbvh_nodes && d->num_sbvh_nodes) {
            if (!d->sbvh_leaf_records || !d->num_sbvh_leaf_records) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "sbvh_nodes without sbvh_leaf_records");
            for (uint32_t i = 0; i < d->num_sbvh_nodes; ++i) {
                const SlrGpuSbvhNode& n = d->sbvh_nodes[i];
                if (n.b & 0x80000000u) { if ((uint64_t)n.a + (n.b & 0x7FFFFFFFu) > d->num_sbvh_leaf_records) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "SBVH node %u: leaf records out of range", i); }
                else if (n.a >= d->num_sbvh_nodes || (n.b & 0x0FFFFFFFu) >= d->num_sbvh_nodes || ((n.b >> 28) & 3u) > 2u) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "SBVH node %u: child / axis out of range", i);
            }
            for (uint32_t i = 0; i < d->num_sbvh_leaf_records; ++i) {
                uint32_t id;
                memcpy(&id, &d->sbvh_leaf_records[i].a[3], 4);
                if ((id & 0x80000000u) ? (id & 0x7FFFFFFFu) >= d->num_instances : (shaded && id >= d->num_triangles))
                    return fail(SLRGPU_ERR_INVALID_ARGUMENT, "SBVH leaf record %u: object out of range", i);
            }
            for (uint32_t i = 0; i < d->num_instances; ++i)
                if (d->instances[i].sbvh_root_node >= d->num_sbvh_nodes) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "instance %u: SBVH root out of range", i);
        }
        *hasAlpha = false;
        for (uint32_t i = 0; i < d->num_leaf_records; ++i) {
            uint32_t id, flags;
            memcpy(&id, &d->leaf_records[i].a[3], 4);
            memcpy(&flags, &d->leaf_records[i].b[3], 4);
            if (id & 0x80000000u) {
                if ((id & 0x7FFFFFFFu) >= d->num_instances) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "leaf record %u: instance %u out of range", i, id & 0x7FFFFFFFu);
                continue;
            }
            if (shaded && id >= d->num_triangles) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "leaf record %u: triangle %u out of range", i, id);
            if (flags & SLRGPU_LEAF_FLAG_ALPHA_TEST) {
                if (!shaded || !d->textures) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "leaf record %u asks for the alpha test but the scene has no triangle / texture tables", i);
                *hasAlpha = true;
            }
        }
        if (d->camera_motion > d->num_motions || (d->num_motions && !d->motions)) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "camera_motion / motions out of range");
        for (uint32_t i = 0; i < d->num_motions; ++i)
            if (!(d->motions[i].t_end > d->motions[i].t_begin)) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "motion %u: t_end must be greater than t_begin", i);
        for (uint32_t i = 0; i < d->num_instances; ++i) {
            const SlrGpuInstance& in = d->instances[i];
            if (in.motion > d->num_motions) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "instance %u: motion index out of range", i);
            if (in.root_node >= d->num_bvh_nodes) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "instance %u: root node out of range", i);
            if (in.num_lights && (in.light_base == SLRGPU_INVALID_ID || (uint64_t)in.light_base + in.num_lights > d->num_lights))
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "instance %u: light list out of range", i);
            if (in.light_index != SLRGPU_INVALID_ID && in.light_index >= d->num_lights) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "instance %u: light index out of range", i);
        }
        // one level of instancing (slrgpu.h, SlrGpuInstance): no instance record inside a nested BVH
        {
            std::vector<uint32_t> roots;
            for (uint32_t i = 0; i < d->num_instances; ++i) roots.push_back(d->instances[i].root_node);
            std::sort(roots.begin(), roots.end());
            roots.erase(std::unique(roots.begin(), roots.end()), roots.end());
            std::vector<uint32_t> stack;
            for (uint32_t root : roots) {
                stack.assign(1, root);
                uint64_t visited = 0;
                while (!stack.empty()) {
                    const SlrGpuBvhNode& n = d->bvh_nodes[stack.back()];
                    stack.pop_back();
                    if (++visited > d->num_bvh_nodes) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "the BVH under node %u is not a tree", root);
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t c = n.child[k];
                        if (c == 0xFFFFFFFFu) continue;
                        const uint32_t idx = c & 0x07FFFFFFu, cnt = (c >> 27) & 0xFu;
                        if (!(c >> 31)) { stack.push_back(idx); continue; }
                        for (uint32_t j = 0; j < cnt; ++j) {
                            uint32_t id;
                            memcpy(&id, &d->leaf_records[idx + j].a[3], 4);
                            if (id & 0x80000000u)
                                return fail(SLRGPU_ERR_UNSUPPORTED, "instancing nested deeper than one level (instance %u inside the BVH of another instance)", id & 0x7FFFFFFFu);
                        }
                    }
                }
            }
        }
        if (d->num_top_lights > d->num_lights) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "num_top_lights exceeds num_lights");
        if (d->environment.present) {
            const SlrGpuEnvironment& e = d->environment;
            if (e.material >= d->num_materials) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "environment material out of range");
            if (!e.map_width || !e.map_height || !e.row_pdf || !e.row_cdf || !e.row_integral || !e.marginal_pdf || !e.marginal_cdf)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "environment importance map missing");
        }
        for (uint32_t i = 0; i < d->num_lights; ++i) {
            const uint32_t o = d->lights[i].object;
            if (o & 0x80000000u) { if ((o & 0x7FFFFFFFu) >= d->num_instances) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "light %u: instance out of range", i); }
            else if (!shaded || o >= d->num_triangles) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "light %u: triangle out of range", i);
        }
        *classMask = 0;
        if (!shaded) return SLRGPU_OK;
        // per material: class bit and tree limits, worked out once
        std::vector<uint32_t> matClass(d->num_materials, 0xFFFFFFFFu);
        for (uint32_t i = 0; i < d->num_triangles; ++i) {
            const SlrGpuTriangle& t = d->triangles[i];
            if (t.v[0] >= d->num_vertices || t.v[1] >= d->num_vertices || t.v[2] >= d->num_vertices)
                return fail(SLRGPU_ERR_INVALID_ARGUMENT, "triangle %u: vertex index out of range", i);
            if (t.material >= d->num_materials) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "triangle %u: material %u out of range", i, t.material);
            if (!texOk(t.normal_map, true) || !texOk(t.alpha_map, true)) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "triangle %u: normal / alpha map out of range", i);
            if (t.light_index != SLRGPU_INVALID_ID && t.light_index >= d->num_lights) return fail(SLRGPU_ERR_INVALID_ARGUMENT, "triangle %u: light index out of range", i);
            if (matClass[t.material] == 0xFFFFFFFFu) {
                int maxStack = 1;
                const int lobes = countLobes(t.material, 0, &maxStack, 1);
                if (lobes < 0) return fail(SLRGPU_ERR_UNSUPPORTED, "material %u: tree deeper than 16 levels (or cyclic)", t.material);
                if (lobes > kMaxLobes) return fail(SLRGPU_ERR_UNSUPPORTED, "material %u has %d leaf lobes; at most %d are supported (MultiBSDF.h:17)", t.material, lobes, kMaxLobes);
                if (maxStack > kMaterialStack) return fail(SLRGPU_ERR_UNSUPPORTED, "material %u: tree needs a traversal stack of %d (limit %d)", t.material, maxStack, kMaterialStack);
                uint32_t leaf = SLRGPU_INVALID_ID;
                const uint32_t cls = classifyMaterialIn(d->materials, t.material, &leaf);
                matClass[t.material] = cls == 0xFFu ? 0x100u : cls;
            }
            if (matClass[t.material] < 16u) *classMask |= 1u << matClass[t.material];
        }
        return SLRGPU_OK;
    }
};
}  // namespace

// Copies every triangle's surface word (class | emitting | leaf material, packSurfaceInfo) into the spare word of the
// leaf records that refer to it, on the device copy: the traversal hands it on with the hit (traverse.cuh Hit::info), so
// the `surface` stage never has to fetch the triangle record of a hit.
__global__ void patchLeafSurfaceInfoKernel(float4* leaves, uint32_t numLeaves, const SlrGpuTriangle* __restrict__ triangles, uint32_t numTriangles) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numLeaves; i += gridDim.x * blockDim.x) {
        const uint32_t id = __float_as_uint(leaves[(size_t)i * 3].w);
        uint32_t info = kSurfaceInfoDynamic;
        if (!(id & 0x80000000u) && id < numTriangles) info = triangles[id].pad;
        leaves[(size_t)i * 3 + 2].w = __uint_as_float(info);
    }
}

// QBVH::Node's three axis numbers -> the three axis masks the walk tests (device_scene.h nodeAxisMasks), on the device copy
__global__ void patchNodeAxesKernel(float4* nodes, uint32_t numNodes) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numNodes; i += gridDim.x * blockDim.x) {
        uint32_t* w = reinterpret_cast<uint32_t*>(nodes + (size_t)i * 8 + 7);
        const uint32_t a = *w;
        *w = nodeAxisMasks(a & 0xFFu, (a >> 8) & 0xFFu, (a >> 16) & 0xFFu);
    }
}

}  // namespace slrgpu

using namespace slrgpu;

extern "C" {

SLRGPU_API int slrgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

SLRGPU_API uint32_t slrgpu_abi_version(void) { return (1u << 16) | 2u; }      // 1.1: SBVH tables, slrgpu_render_multi; 1.2: SLRGPU_RENDER_BPT, slrgpu_probe_shading_bpt

SLRGPU_API const char* slrgpu_last_error(void) { return g_error; }

SLRGPU_API uint32_t slrgpu_struct_size(int which) {
    static const uint32_t sizes[] = {
        sizeof(SlrGpuSceneDesc), sizeof(SlrGpuBvhNode), sizeof(SlrGpuLeafRecord), sizeof(SlrGpuInstance),
        sizeof(SlrGpuTriangle), sizeof(SlrGpuVertex), sizeof(SlrGpuSpectrum), sizeof(SlrGpuTexture),
        sizeof(SlrGpuImage), sizeof(SlrGpuMaterial), sizeof(SlrGpuLight), sizeof(SlrGpuCamera),
        sizeof(SlrGpuEnvironment), sizeof(SlrGpuSpectralTables), sizeof(SlrGpuRayBatch), sizeof(SlrGpuHitBatch),
        sizeof(SlrGpuRenderParams), sizeof(SlrGpuRenderStats), sizeof(SlrGpuSbvhNode), sizeof(SlrGpuMotion)};
    if (which < 0 || which >= (int)(sizeof(sizes) / sizeof(sizes[0]))) return 0;
    return sizes[which];
}

SLRGPU_API int slrgpu_scene_create(const SlrGpuSceneDesc* d, int device, SlrGpuScene** out) {
    if (!d || !out) { setError("slrgpu_scene_create: null argument"); return SLRGPU_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    if (d->struct_size != sizeof(SlrGpuSceneDesc)) {
        setError("slrgpu_scene_create: struct_size %u != %zu (ABI mismatch)", d->struct_size, sizeof(SlrGpuSceneDesc));
        return SLRGPU_ERR_INVALID_ARGUMENT;
    }
    if (!d->bvh_nodes || d->num_bvh_nodes == 0 || !d->leaf_records || d->num_leaf_records == 0) {
        setError("slrgpu_scene_create: a scene needs at least one BVH node and one leaf record");
        return SLRGPU_ERR_INVALID_ARGUMENT;
    }
    if (d->num_bvh_nodes > 0x07FFFFFFu || d->num_leaf_records > 0x07FFFFFFu) {
        setError("slrgpu_scene_create: node / leaf-record count exceeds the 27-bit child index");
        return SLRGPU_ERR_UNSUPPORTED;
    }
    bool hasAlpha = false;
    uint32_t classMask = 0;
    {
        Validator v{d};
        int vr;
        if ((vr = v.materials()) || (vr = v.textures()) || (vr = v.geometry(&hasAlpha, &classMask))) return vr;
    }
    int n = slrgpu_device_count();
    if (n == 0) { setError("no CUDA device available (this library has no CPU fallback)"); return SLRGPU_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { setError("device %d out of range [0,%d)", device, n); return SLRGPU_ERR_INVALID_ARGUMENT; }
    SLRGPU_CUDA_TRY(cudaSetDevice(device));

    SlrGpuScene* sc = new (std::nothrow) SlrGpuScene();
    if (!sc) { setError("host allocation failed"); return SLRGPU_ERR_OUT_OF_MEMORY; }
    sc->device = device;
    if (cudaDeviceGetAttribute(&sc->numSMs, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sc->numSMs < 1) sc->numSMs = 148;
    DeviceScene& v = sc->dev;
    memset(&v, 0, sizeof(v));
    int rc = SLRGPU_OK;
    UploadPlan plan;
#define UP(src, count, dst) plan.add(src, (uint64_t)(count), dst)
    UP(reinterpret_cast<const float4*>(d->bvh_nodes), (uint64_t)d->num_bvh_nodes * 8, &v.nodes);
    UP(reinterpret_cast<const float4*>(d->leaf_records), (uint64_t)d->num_leaf_records * 3, &v.leaves);
    UP(d->instances, d->num_instances, &v.instances);
    UP(d->motions, d->num_motions, &v.motions);
    if (d->sbvh_nodes && d->num_sbvh_nodes && d->sbvh_leaf_records && d->num_sbvh_leaf_records) {
        UP(d->sbvh_nodes, d->num_sbvh_nodes, &v.sbvhNodes);
        UP(reinterpret_cast<const float4*>(d->sbvh_leaf_records), (uint64_t)d->num_sbvh_leaf_records * 3, &v.sbvhLeaves);
    }
    // the device copy of the triangle records carries the surface stage's per-triangle facts in its spare word
    std::vector<SlrGpuTriangle> triangles;
    if (d->triangles && d->num_triangles) {
        triangles.assign(d->triangles, d->triangles + d->num_triangles);
        for (SlrGpuTriangle& t : triangles)
            t.pad = d->materials ? packSurfaceInfo(d->materials, d->num_materials, t.material) : kSurfaceInfoDynamic;
    }
    UP(triangles.data(), triangles.size(), &v.triangles);
    UP(reinterpret_cast<const float4*>(d->vertices), (uint64_t)d->num_vertices * 3, &v.vertices);
    UP(d->materials, d->num_materials, &v.materials);
    UP(d->textures, d->num_textures, &v.textures);
    std::vector<SlrGpuSpectrum> spectra;
    std::vector<float> spectrumData;
    if (d->spectra && d->num_spectra) compileSpectra(d, spectra, spectrumData);
    UP(spectra.data(), spectra.size(), &v.spectra);
    UP(spectrumData.data(), spectrumData.size(), &v.spectrumData);
    UP(d->images, d->num_images, &v.images);
    UP(d->image_data, d->image_data_bytes, &v.imageData);
    UP(d->lights, d->num_lights, &v.lights);
    const SlrGpuEnvironment& e = d->environment;
    if (e.present) {
        const uint64_t w = e.map_width, h = e.map_height;
        UP(e.row_pdf, w * h, &v.envRowPdf);
        UP(e.row_cdf, (w + 1) * h, &v.envRowCdf);
        UP(e.row_integral, h, &v.envRowIntegral);
        UP(e.marginal_pdf, h, &v.envMarginalPdf);
        UP(e.marginal_cdf, h + 1, &v.envMarginalCdf);
    }
    UP(d->spectral.upsample_grid, d->spectral.upsample_grid_floats, &v.upsampleGrid);
    UP(d->spectral.upsample_points, d->spectral.upsample_points_floats, &v.upsamplePoints);
#undef UP
    rc = plan.commit(sc);
    if (rc != SLRGPU_OK) { slrgpu_scene_destroy(sc); return rc; }
    if (v.triangles && v.leaves) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)d->num_leaf_records + 255) / 256, 148 * 8);
        patchLeafSurfaceInfoKernel<<<blocks, 256>>>(const_cast<float4*>(v.leaves), d->num_leaf_records, v.triangles, d->num_triangles);
        cudaError_t pe = cudaGetLastError();
        if (pe == cudaSuccess) pe = cudaDeviceSynchronize();
        if (pe != cudaSuccess) { slrgpu_scene_destroy(sc); return cudaFail(pe, "patchLeafSurfaceInfoKernel"); }
    }

    {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)d->num_bvh_nodes + 255) / 256, 148 * 8);
        patchNodeAxesKernel<<<blocks, 256>>>(const_cast<float4*>(v.nodes), d->num_bvh_nodes);
        cudaError_t pe = cudaGetLastError();
        if (pe == cudaSuccess) pe = cudaDeviceSynchronize();
        if (pe != cudaSuccess) { slrgpu_scene_destroy(sc); return cudaFail(pe, "patchNodeAxesKernel"); }
    }

    v.numNodes = d->num_bvh_nodes; v.numLeaves = d->num_leaf_records; v.numInstances = d->num_instances;
    v.numTriangles = d->num_triangles; v.numVertices = d->num_vertices; v.numMaterials = d->num_materials;
    v.numTextures = d->num_textures; v.numSpectra = d->num_spectra; v.numImages = d->num_images;
    v.numLights = d->num_lights; v.numTopLights = d->num_top_lights;
    v.envPresent = e.present; v.envMaterial = e.material; v.envMapWidth = e.map_width; v.envMapHeight = e.map_height;
    v.envMarginalIntegral = e.marginal_integral;
    v.topLightImportance = d->top_light_importance;
    v.rgbMode = d->rgb_mode;
    for (int i = 0; i < 3; ++i) v.worldCenter[i] = d->world_center[i];
    v.worldRadius = d->world_radius;
    v.camera = d->camera;
    if (d->spectral.xbar_16 && d->spectral.ybar_16 && d->spectral.zbar_16) {
        for (int i = 0; i < 16; ++i) { v.xbar16[i] = d->spectral.xbar_16[i]; v.ybar16[i] = d->spectral.ybar_16[i]; v.zbar16[i] = d->spectral.zbar_16[i]; }
    }
    v.integralCMF = d->spectral.integral_cmf;
    sc->hasInstances = d->num_instances > 0;
    sc->hasShading = d->num_materials > 0 && d->num_triangles > 0;
    sc->channels = d->rgb_mode ? 3 : 16;
    // which material-class kernels a wave has to launch: the classes the device's own classification (classifyMaterialIn)
    // gives the triangles' materials, worked out by the validation pass
    sc->classMask = classMask;
    sc->hasMotion = d->camera_motion != 0;
    bool movingInstance = false;
    for (uint32_t i = 0; i < d->num_instances; ++i) movingInstance = movingInstance || d->instances[i].motion != 0;
    sc->hasMotion = sc->hasMotion || movingInstance;
    sc->hasAlpha = hasAlpha || movingInstance;         // both take the general instantiation of the walk kernels
    v.hasAlpha = sc->hasAlpha ? 1u : 0u;
    v.cameraMotion = d->camera_motion;
    *out = sc;
    return SLRGPU_OK;
}

SLRGPU_API void slrgpu_scene_destroy(SlrGpuScene* sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    for (int i = 0; i < sc->numAllocations; ++i) {
        if (i == 0 && sc->arenaBytes && cacheArena(sc->device, sc->allocations[0], sc->arenaBytes)) continue;
        cudaFree(sc->allocations[i]);
    }
    if (sc->statusRing) cudaFree(sc->statusRing);
    delete sc;
}

SLRGPU_API uint64_t slrgpu_scene_device_bytes(const SlrGpuScene* sc) { return sc ? sc->deviceBytes : 0; }

SLRGPU_API uint32_t slrgpu_scene_channels(const SlrGpuScene* sc) { return sc ? sc->channels : 0; }

}  // extern "C"
