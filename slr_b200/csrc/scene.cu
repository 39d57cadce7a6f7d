// Scene upload: copies the caller's SoA host buffers into HBM once and keeps a device view.
// Replaces, on the GPU side, the object graph a reference renderer receives from
// SLRSceneGraph::Scene::build (libSLRSceneGraph/Scene.cpp:28-44).
#include "device_scene.h"
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

}  // namespace slrgpu

using namespace slrgpu;

extern "C" {

SLRGPU_API int slrgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

SLRGPU_API uint32_t slrgpu_abi_version(void) { return (1u << 16) | 0u; }

SLRGPU_API const char* slrgpu_last_error(void) { return g_error; }

SLRGPU_API uint32_t slrgpu_struct_size(int which) {
    static const uint32_t sizes[] = {
        sizeof(SlrGpuSceneDesc), sizeof(SlrGpuBvhNode), sizeof(SlrGpuLeafRecord), sizeof(SlrGpuInstance),
        sizeof(SlrGpuTriangle), sizeof(SlrGpuVertex), sizeof(SlrGpuSpectrum), sizeof(SlrGpuTexture),
        sizeof(SlrGpuImage), sizeof(SlrGpuMaterial), sizeof(SlrGpuLight), sizeof(SlrGpuCamera),
        sizeof(SlrGpuEnvironment), sizeof(SlrGpuSpectralTables), sizeof(SlrGpuRayBatch), sizeof(SlrGpuHitBatch),
        sizeof(SlrGpuRenderParams), sizeof(SlrGpuRenderStats)};
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
    // which material-class kernels a wave has to launch (same numbering as ShadeClass / LobeType)
    sc->classMask = 0;
    for (uint32_t i = 0; i < d->num_materials; ++i) {
        const SlrGpuMaterial& m = d->materials[i];
        switch (m.kind) {
            case SLRGPU_MAT_DIFFUSE: sc->classMask |= m.tex[1] == SLRGPU_INVALID_ID ? 1u << 0 : 1u << 1; break;
            case SLRGPU_MAT_SPECULAR_REFLECTION: sc->classMask |= 1u << 2; break;
            case SLRGPU_MAT_SPECULAR_SCATTERING: sc->classMask |= 1u << 3; break;
            case SLRGPU_MAT_WARD_DUR: sc->classMask |= 1u << 4; break;
            case SLRGPU_MAT_ASHIKHMIN_SHIRLEY: sc->classMask |= 1u << 5; break;
            case SLRGPU_MAT_MICROFACET_REFLECTION: sc->classMask |= 1u << 6; break;
            case SLRGPU_MAT_MICROFACET_SCATTERING: sc->classMask |= 1u << 7; break;
            case SLRGPU_MAT_INVERSE: case SLRGPU_MAT_SUMMED: case SLRGPU_MAT_MIXED: sc->classMask |= 1u << 8; break;
            default: break;
        }
    }
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
    delete sc;
}

SLRGPU_API uint64_t slrgpu_scene_device_bytes(const SlrGpuScene* sc) { return sc ? sc->deviceBytes : 0; }

SLRGPU_API uint32_t slrgpu_scene_channels(const SlrGpuScene* sc) { return sc ? sc->channels : 0; }

}  // extern "C"
