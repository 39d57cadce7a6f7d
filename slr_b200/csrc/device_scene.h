// Device-side view of an uploaded scene (pointers into HBM) shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
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
    float worldCenter[3];
    float worldRadius;
    SlrGpuCamera camera;
    float xbar16[16], ybar16[16], zbar16[16];
    float integralCMF;
};

}  // namespace slrgpu

struct SlrGpuScene {
    int device = 0;
    slrgpu::DeviceScene dev;          // host copy of the device view (passed to kernels by value / constant)
    void* allocations[32] = {};
    int numAllocations = 0;
    uint64_t deviceBytes = 0;
    uint64_t arenaBytes = 0;          // size of allocations[0], the single arena all scene buffers live in
    bool hasInstances = false;
    bool hasShading = false;
    uint32_t channels = 16;
    uint32_t classMask = 0;           // material classes (wavefront.cuh: ShadeClass) the scene's materials can produce
    int numSMs = 148;                 // of `device`
};

namespace slrgpu {
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
