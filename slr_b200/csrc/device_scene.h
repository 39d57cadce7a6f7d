// Device-side view of an uploaded scene (pointers into HBM) shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include "../../include/slrgpu.h"

namespace slrgpu {

// internal spectrum kind produced by scene.cu compileSpectra (never part of the ABI)
constexpr uint32_t SLRGPU_SPECTRUM_IRREGULAR_LUT = 16;
constexpr int SLRGPU_SPECTRUM_LUT_BINS = 236;      // 2 nm bins over [360, 830] + 1

// Layout in HBM (all arrays 256-byte aligned by cudaMalloc):
//   nodes    : float4[8 * num_nodes]   one 128 B QBVH node = 8 consecutive float4 (see SlrGpuBvhNode)
//   leaves   : float4[3 * num_leaves]  one 48 B leaf record = 3 consecutive float4 (see SlrGpuLeafRecord)
//   instances: SlrGpuInstance[num_instances]
struct DeviceScene {
    const float4* nodes;
    const float4* leaves;
    const SlrGpuInstance* instances;
    const SlrGpuMotion* motions;         // animated transforms (motion blur), SlrGpuInstance::motion / cameraMotion index + 1
    const SlrGpuSbvhNode* sbvhNodes;     // optional second accelerator (slrgpu_intersect_batch_sbvh)
    const float4* sbvhLeaves;
    const SlrGpuTriangle* triangles;
    const float4* vertices;          // 3 float4 per vertex (SlrGpuVertex)
    const SlrGpuMaterial* materials;
    const SlrGpuTexture* textures;
    const SlrGpuSpectrum* spectra;
    const float* spectrumData;
    const SlrGpuImage* images;
    const uint8_t* imageData;
    const SlrGpuLight* lights;
    // environment importance map
    const float* envRowPdf;
    const float* envRowCdf;
    const float* envRowIntegral;
    const float* envMarginalPdf;
    const float* envMarginalCdf;
    // spectral tables
    const float* upsampleGrid;
    const float* upsamplePoints;
    uint32_t numNodes, numLeaves, numInstances, numTriangles, numVertices;
    uint32_t numMaterials, numTextures, numSpectra, numImages, numLights, numTopLights;
    uint32_t envPresent, envMaterial, envMapWidth, envMapHeight;
    float envMarginalIntegral;
    float topLightImportance;     // SurfaceObjectAggregate::importance() of the top-level aggregate
    uint32_t rgbMode;
    uint32_t hasAlpha;            // the scene needs the GENERAL walk: some leaf record carries SLRGPU_LEAF_FLAG_ALPHA_TEST, or an instance moves
    uint32_t cameraMotion;        // 0 = static camera, else 1 + index into motions
    float worldCenter[3];
    float worldRadius;
    SlrGpuCamera camera;
    float xbar16[16], ybar16[16], zbar16[16];
    float integralCMF;
};

// The axis word of a node ON THE DEVICE (scene.cu patchNodeAxesKernel rewrites it at scene creation): byte 0 = 1 << topAxis,
// byte 1 = 1 << leftAxis, byte 2 = 1 << rightAxis, byte 3 = 0 -- one AND with the ray's replicated direction signs
// (traverse.cuh walkSetRay) answers the three ordering questions of a node visit. The table handed to
// slrgpu_scene_create keeps QBVH::Node's plain axis numbers.
__host__ __device__ inline uint32_t nodeAxisMasks(uint32_t top, uint32_t left, uint32_t right) {
    return (1u << top) | ((1u << left) << 8) | ((1u << right) << 16);
}

#ifdef __CUDACC__
// cache hint for a queue entry a later iteration / refill will stream (no register is held for it)
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// Material class of a hit (wavefront.cuh: ShadeClass values, numbered like LobeType) and the leaf
// material the class kernel builds its lobe from. Emitter wrappers are peeled (their BSDF is the
// scattering material's, surface_material.h); sum / mix / inverse trees go to the generic kernel with
// the ORIGINAL material id. Returns 0xFF when the hit has no BSDF (an emitter without a scattering part).
__host__ __device__ inline uint32_t classifyMaterialIn(const SlrGpuMaterial* materials, uint32_t materialId, uint32_t* leaf) {
    uint32_t id = materialId;
    SlrGpuMaterial m = materials[id];
    for (int depth = 0; depth < 4 && m.kind == SLRGPU_MAT_EMITTER; ++depth) {
        if (m.sub[0] == SLRGPU_INVALID_ID) return 0xFFu;
        id = m.sub[0];
        m = materials[id];
    }
    *leaf = id;
    switch (m.kind) {
    case SLRGPU_MAT_DIFFUSE: return m.tex[1] == SLRGPU_INVALID_ID ? 0u : 1u;
    case SLRGPU_MAT_SPECULAR_REFLECTION: return 2u;
    case SLRGPU_MAT_SPECULAR_SCATTERING: return 3u;
    case SLRGPU_MAT_WARD_DUR: return 4u;
    case SLRGPU_MAT_ASHIKHMIN_SHIRLEY: return 5u;
    case SLRGPU_MAT_MICROFACET_REFLECTION: return 6u;
    case SLRGPU_MAT_MICROFACET_SCATTERING: return 7u;
    default: *leaf = materialId; return 8u;
    }
}

// What the `surface` stage needs to know about a hit triangle, worked out once at scene creation and kept in the
// triangle record's spare word (SlrGpuTriangle::pad on the device copy): class | emitting << 8 | leaf material << 9.
// It replaces two to three dependent loads of the material table per hit. kSurfaceInfoDynamic: not precomputed.
constexpr uint32_t kSurfaceInfoDynamic = 0xFFFFFFFFu;
constexpr uint32_t kSurfaceInfoMiss = 0xFFFFFFFEu;       // hit record of a ray that left the scene
inline uint32_t packSurfaceInfo(const SlrGpuMaterial* materials, uint32_t numMaterials, uint32_t materialId) {
    if (materialId >= numMaterials || numMaterials >= (1u << 23)) return kSurfaceInfoDynamic;
    uint32_t leaf = SLRGPU_INVALID_ID;
    const uint32_t cls = classifyMaterialIn(materials, materialId, &leaf);
    const uint32_t emitting = materials[materialId].kind == SLRGPU_MAT_EMITTER ? 1u : 0u;
    if (cls == 0xFFu) leaf = 0;
    if (leaf >= (1u << 23) - 2u) return kSurfaceInfoDynamic;      // keeps the two reserved words out of the packed range
    return cls | (emitting << 8) | (leaf << 9);
}

}  // namespace slrgpu

struct SlrGpuScene {
    int device = 0;
    slrgpu::DeviceScene dev;          // host copy of the device view (passed to kernels by value / constant)
    void* allocations[32] = {};
    int numAllocations = 0;
    uint64_t deviceBytes = 0;
    uint64_t arenaBytes = 0;          // size of allocations[0], the single arena all scene buffers live in
    bool hasInstances = false;
    bool hasAlpha = false;            // alpha-mapped (cut-out) triangles or moving instances: the walk kernels run their general instantiation
    bool hasMotion = false;           // animated transforms present: rays carry their time through the queues
    bool hasShading = false;
    uint32_t channels = 16;
    uint32_t classMask = 0;           // material classes (wavefront.cuh: ShadeClass) the scene's materials can produce
    int numSMs = 148;                 // of `device`
    // slrgpu_intersect_batch_device: kStatusSlots x (overflow flag, chunk cursor) on `device`, one slot per launch in flight
    int* statusRing = nullptr;
    std::atomic<uint32_t> statusNext{0};
    std::mutex statusMutex;
};

namespace slrgpu {
constexpr uint32_t kStatusSlots = 64;
void setError(const char* fmt, ...);
void releaseSceneArenas();
// intersect.cu: closest hits of a device-resident SoA ray batch; dStatus = two zeroed ints (overflow flag, chunk cursor)
int launchIntersect(SlrGpuScene* sc, const SlrGpuRayBatch& rays, uint64_t n, const SlrGpuHitBatch& hits, int* dStatus, cudaStream_t stream);        // scene.cu: drops the per-device arena cache
int cudaFail(cudaError_t e, const char* what);
// Grid of a grid-stride / persistent kernel: as many blocks as the device keeps resident (SMs x blocks per SM at the
// kernel's register count), never more than the work needs. Launching 16 blocks per SM for kernels that hold 3-6 costs
// ~10 us of block scheduling per launch -- with ~10 launches per wave that was the floor of the small waves.
template <typename Kernel>
inline uint32_t residentGrid(Kernel kernel, int blockSize, int numSMs, uint64_t maxBlocks) {
    int perSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, blockSize, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
    const uint64_t g = (uint64_t)numSMs * (uint64_t)perSM;
    return (uint32_t)(g < maxBlocks ? g : (maxBlocks ? maxBlocks : 1));
}
#define SLRGPU_CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return slrgpu::cudaFail(_e, #expr); } while (0)
}
