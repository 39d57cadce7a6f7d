/*
 * slrgpu.h -- C ABI of the B200 (sm_100a) path-tracing hot path for SLR.
 *
 * This is the drop-in boundary. The reference (goofoo/SLR) has no plugin/FFI layer; its seams are
 * C++ virtual interfaces inside libSLR. Each entry point below names the reference interface it
 * stands in for:
 *
 *   slrgpu_scene_create     the scene a renderer receives: SLR::Scene built by
 *                           SLRSceneGraph::Scene::build (libSLRSceneGraph/Scene.cpp:28-44) over
 *                           SurfaceObjectAggregate + accelerator (libSLR/Core/SurfaceObject.cpp:226-253)
 *   slrgpu_intersect_batch  Accelerator::intersect(Ray&, Intersection*) (libSLR/Core/Accelerator.h:17-34)
 *                           as implemented by QBVH::intersect (libSLR/Accelerator/QBVH.h:295-337),
 *                           one call per ray there, one call per ray batch here
 *   slrgpu_occluded_batch   Scene::testVisibility (libSLR/Core/SurfaceObject.cpp:418-430)
 *   slrgpu_render           Renderer::render(const Scene&, const RenderSettings&)
 *                           (libSLR/Core/Renderer.h:15-19) as implemented by
 *                           PathTracingRenderer::render (libSLR/Renderers/PathTracingRenderer.cpp:27-98)
 *                           with the sensor accumulation of ImageSensor::add (libSLR/Core/ImageSensor.cpp:124-129)
 *
 * Conventions: plain C, plain pointers and sizes, no C++/torch types. Every function returns
 * SLRGPU_OK (0) or a negative SlrGpuStatus and never throws; slrgpu_last_error() gives a
 * thread-local human-readable message. Calls block until the result is in the caller's buffers.
 * Host-pointer entry points copy host<->device inside the call; *_device entry points take CUDA
 * device pointers on the scene's device and enqueue on `stream` (a cudaStream_t passed as void*).
 * A scene is immutable after creation; one call at a time per scene handle.
 *
 * There is no CPU fallback anywhere behind this interface: without a CUDA device every call that
 * needs one fails with SLRGPU_ERR_NO_DEVICE.
 */
#ifndef SLRGPU_H
#define SLRGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SLRGPU_API __declspec(dllexport)
#else
#define SLRGPU_API __attribute__((visibility("default")))
#endif

typedef enum SlrGpuStatus {
    SLRGPU_OK = 0,
    SLRGPU_ERR_INVALID_ARGUMENT = -1,
    SLRGPU_ERR_NO_DEVICE = -2,
    SLRGPU_ERR_CUDA = -3,
    SLRGPU_ERR_OUT_OF_MEMORY = -4,
    SLRGPU_ERR_UNSUPPORTED = -5,
    SLRGPU_ERR_STACK_OVERFLOW = -6   /* traversal stack (64 entries, as QBVH.h:299) overflowed */
} SlrGpuStatus;

#define SLRGPU_INVALID_ID 0xFFFFFFFFu
#define SLRGPU_NUM_WAVELENGTHS 16        /* libSLR/references.h:39 */

/* ---------------------------------------------------------------------------------------------
 * Geometry
 * ------------------------------------------------------------------------------------------- */

/* 128-byte 4-wide BVH node, field-for-field the layout of QBVH::Node (QBVH.h:42-53):
 * lanes' min x/y/z, max x/y/z, four packed children, three split axes. Empty lanes hold
 * (+inf, -inf) boxes and child 0xFFFFFFFF. Child word: idx:27 | numLeaves:4 | isLeaf:1 (QBVH.h:27-35).
 * For an inner child idx is a node index, for a leaf child idx is the first leaf record; both are
 * GLOBAL indices into the scene-wide arrays below (all BVH levels are concatenated).
 * The caller's table is read once: slrgpu_scene_create rewrites the three axis bytes of ITS device copy into bit masks
 * (1 << axis) for the traversal's ordering test, and copies each triangle's shading class into the spare word of the
 * device copy of the leaf records -- neither is visible through this interface. */
typedef struct SlrGpuBvhNode {
    float lo_x[4], lo_y[4], lo_z[4];
    float hi_x[4], hi_y[4], hi_z[4];
    uint32_t child[4];
    uint8_t top_axis, left_axis, right_axis, pad0;
    uint32_t pad[3];
} SlrGpuBvhNode;

/* 48-byte leaf record, one per BVH leaf reference, stored in leaf order so a leaf's records are
 * contiguous (SBVH duplicates a primitive referenced from several leaves).
 *   triangle: a = (v0.xyz, bits(prim_id)), b = (v1-v0, bits(flags)), c = (v2-v0, 0)
 *             prim_id < 2^31 indexes `triangles`; the edges are the fp32 differences
 *             Triangle::intersect computes (TriangleMesh.cpp:136-137), stored pre-subtracted.
 *   instance: a.w = bits(0x80000000 | instance_id); the rest is unused. */
typedef struct SlrGpuLeafRecord {
    float a[4], b[4], c[4];
} SlrGpuLeafRecord;

#define SLRGPU_LEAF_FLAG_ALPHA_TEST 0x1u   /* triangle has an alpha texture (TriangleMesh.cpp:163-168) */

/* 32-byte node of the binary spatial-split BVH the reference builds and -- in the shipped build -- traverses
 * (SBVH::Node, Accelerator/SBVH.h:21-45; SBVH::intersect :417-442). Optional second accelerator of a scene, used by
 * slrgpu_intersect_batch_sbvh. Inner node: a = child 0, b = child 1 | split axis << 28 (global node indices); leaf:
 * a = first record in sbvh_leaf_records, b = 0x80000000 | number of records. */
typedef struct SlrGpuSbvhNode {
    float lo[3], hi[3];
    uint32_t a, b;
} SlrGpuSbvhNode;

/* One TransformedSurfaceObject (SurfaceObject.cpp:303-392) over a nested aggregate.
 * Matrices are column-major (element (r,c) at [4*c + r]) like Matrix4x4 (Matrix4x4.h:23-38).
 * ONE LEVEL ON THE DEVICE: an instance record may only appear in the top-level BVH; a nested aggregate's leaf records must
 * all be triangles (slrgpu_scene_create walks every nested BVH and returns SLRGPU_ERR_UNSUPPORTED for an instance record
 * inside one; hit records, surface points and light sampling carry one instance id per hit). The reference recurses to
 * any depth (SurfaceObject.cpp:307-336): the producer of the tables expands a chain instead -- an instance found inside a
 * referenced subtree is placed again under (outer transform x its own), the subtree's own triangles become an aggregate
 * of their own. The host library does that when it flattens a scene (host/scene.h PlacedSubtree: same closest hits up to
 * the rounding of the composed matrix, same light-selection probabilities), and so does the reference-side exporter
 * (integration/GPUPathTracingRenderer.cpp expandNesting, with the reference's own aggregate and transform classes). */
typedef struct SlrGpuInstance {
    float mat[16];          /* local -> parent */
    float mat_inv[16];      /* parent -> local, as computed by the host's invert() */
    uint32_t root_node;     /* global node index of the nested BVH's root */
    uint32_t light_base;    /* first entry of the nested aggregate's light list in `lights`, or INVALID */
    uint32_t num_lights;
    uint32_t light_index;   /* position of this instance in ITS parent's light list, or INVALID */
    float light_importance; /* integral of the nested light distribution (importance of the instance) */
    uint32_t sbvh_root_node; /* global index of the nested SBVH's root in sbvh_nodes (when the scene carries the SBVH) */
    uint32_t motion;         /* 0 = static; else 1 + index into `motions`: mat / mat_inv are then the key frame at t_begin */
    uint32_t pad;
} SlrGpuInstance;

/* AnimatedTransform (libSLR/Core/Transform.h:89-144) of an instance or of the camera -- motion blur. The transform at ray
 * time `time` is: the begin key frame (the owner's mat / mat_inv) for time <= t_begin, the end key frame for
 * time >= t_end, else translate(lerp(T0, T1, t)) * Slerp(t, R0, R1).toMatrix() * lerp(S0, S1, t) with
 * t = (time - t_begin) / (t_end - t_begin); T, R (quaternion x y z w), S are the host's decomposition of the two key
 * frames (Quaternion.cpp:15-43). Matrices column-major. */
typedef struct SlrGpuMotion {
    float mat_end[16], mat_end_inv[16];
    float T0[3], t_begin;
    float T1[3], t_end;
    float R0[4], R1[4];
    float S0[16], S1[16];
} SlrGpuMotion;

/* 32-byte per-triangle shading record (indexed by prim_id). */
typedef struct SlrGpuTriangle {
    uint32_t v[3];          /* indices into `vertices` */
    uint32_t material;      /* index into `materials` */
    uint32_t normal_map;    /* texture id or INVALID (BumpSingleSurfaceObject, SurfaceObject.cpp:123-134) */
    uint32_t alpha_map;     /* texture id or INVALID */
    uint32_t light_index;   /* position in its aggregate's light list, or INVALID if not emitting */
    uint32_t pad;
} SlrGpuTriangle;

/* 48-byte vertex: (position, u) (normal, v) (tangent, 0) -- Vertex of geometry.h:148-156 padded to
 * three 16-byte loads. */
typedef struct SlrGpuVertex {
    float position[3], u;
    float normal[3], v;
    float tangent[3], pad;
} SlrGpuVertex;

/* ---------------------------------------------------------------------------------------------
 * Shading tables (materials, textures, spectra, lights, environment, camera)
 * ------------------------------------------------------------------------------------------- */

typedef enum SlrGpuSpectrumKind {
    SLRGPU_SPECTRUM_REGULAR = 0,    /* RegularContinuousSpectrum   (SpectrumTypes.h:70-117)  */
    SLRGPU_SPECTRUM_IRREGULAR = 1,  /* IrregularContinuousSpectrum (SpectrumTypes.h:119-167) */
    SLRGPU_SPECTRUM_UPSAMPLED = 2,  /* UpsampledContinuousSpectrum (SpectrumTypes.h:169-346), Meng-Simon */
    SLRGPU_SPECTRUM_RGB = 3         /* RGB mode: plain (r,g,b) triple (RGBTypes.h) */
} SlrGpuSpectrumKind;

/* 32 bytes. REGULAR: data_offset/num_samples index `spectrum_data` (values), p0=minLambda,
 * p1=maxLambda. IRREGULAR: data_offset -> num_samples lambdas followed by num_samples values.
 * UPSAMPLED: p0=u, p1=v, p2=scale (already divided by the equal-energy constant as the ctor does).
 * RGB: p0,p1,p2 = r,g,b. */
typedef struct SlrGpuSpectrum {
    uint32_t kind;
    uint32_t data_offset;
    uint32_t num_samples;
    float p0, p1, p2;
    uint32_t pad[2];
} SlrGpuSpectrum;

typedef enum SlrGpuTextureKind {
    SLRGPU_TEX_CONSTANT_SPECTRUM = 0,  /* p: spectrum id */
    SLRGPU_TEX_CONSTANT_FLOAT = 1,     /* f0: value */
    SLRGPU_TEX_CHECKER_SPECTRUM = 2,   /* ids: spectrum0, spectrum1; mapping */
    SLRGPU_TEX_CHECKER_NORMAL = 3,     /* f0: stepWidth, i0: reverse; mapping */
    SLRGPU_TEX_CHECKER_FLOAT = 4,      /* f0,f1: values; mapping */
    SLRGPU_TEX_VORONOI_SPECTRUM = 5,   /* f0: scale, f1: brightness; mapping (3D) */
    SLRGPU_TEX_VORONOI_NORMAL = 6,     /* f0: scale, f1: cos(thetaMax) (voronoi_textures.h:37) */
    SLRGPU_TEX_VORONOI_FLOAT = 7,      /* f0: scale, f1: valueScale, i0: flat */
    SLRGPU_TEX_IMAGE_SPECTRUM = 8,     /* i0: image id; mapping */
    SLRGPU_TEX_IMAGE_NORMAL = 9,
    SLRGPU_TEX_IMAGE_FLOAT = 10
} SlrGpuTextureKind;

typedef enum SlrGpuMappingKind {
    SLRGPU_MAP_TEXCOORD = 0,           /* Texture2DMapping: surfPt.texCoord (textures.h:16-24) */
    SLRGPU_MAP_OFFSET_SCALE_2D = 1,    /* OffsetAndScale2DMapping (textures.h:26-36) */
    SLRGPU_MAP_WORLD_POS = 2           /* WorldPosition3DMapping (textures.h:44-50) */
} SlrGpuMappingKind;

/* 48 bytes */
typedef struct SlrGpuTexture {
    uint32_t kind;
    uint32_t mapping;
    uint32_t i0, i1;
    float f0, f1, f2, f3;
    float map_offset[2], map_scale[2];
} SlrGpuTexture;

typedef enum SlrGpuImageFormat {
    SLRGPU_IMG_RGB8x3 = 0, SLRGPU_IMG_RGB_8x4 = 1, SLRGPU_IMG_RGBA8x4 = 2, SLRGPU_IMG_RGBA16Fx4 = 3,
    SLRGPU_IMG_GRAY8 = 4, SLRGPU_IMG_UVS16Fx3 = 5, SLRGPU_IMG_UVSA16Fx4 = 6, SLRGPU_IMG_FLOAT32 = 7
} SlrGpuImageFormat;

/* Row-major (not tiled) image; texel bytes start at image_data + data_offset. */
typedef struct SlrGpuImage {
    uint32_t format, width, height, pad;
    uint64_t data_offset;
    uint32_t spectrum_type;   /* 0 reflectance, 1 illuminant, 2 IOR (SpectrumType, Spectrum.h) */
    uint32_t pad1;
} SlrGpuImage;

typedef enum SlrGpuMaterialKind {
    SLRGPU_MAT_DIFFUSE = 0,             /* "matte": tex0 reflectance, tex1 sigma or INVALID      */
    SLRGPU_MAT_SPECULAR_REFLECTION = 1, /* "metal": tex0 coeffR, tex1 eta, tex2 k                */
    SLRGPU_MAT_SPECULAR_SCATTERING = 2, /* "glass": tex0 coeff, tex1 etaExt, tex2 etaInt         */
    SLRGPU_MAT_WARD_DUR = 3,            /* "Ward": tex0 R, tex1 anisoX, tex2 anisoY              */
    SLRGPU_MAT_ASHIKHMIN_SHIRLEY = 4,   /* "Ashikhmin": tex0 Rs, tex1 Rd, tex2 nu, tex3 nv       */
    SLRGPU_MAT_MICROFACET_REFLECTION = 5, /* "microfacet metal": tex0 eta, tex1 k, tex2 alpha_g  */
    SLRGPU_MAT_MICROFACET_SCATTERING = 6, /* "microfacet glass": tex0 etaExt, tex1 etaInt, tex2 alpha_g */
    SLRGPU_MAT_INVERSE = 7,             /* "inverse": sub0                                       */
    SLRGPU_MAT_SUMMED = 8,              /* "sum": sub0, sub1                                     */
    SLRGPU_MAT_MIXED = 9,               /* "mix": sub0, sub1, tex0 factor (float texture)        */
    SLRGPU_MAT_EMITTER = 10,            /* "emitter": sub0 scatter material (or INVALID), sub1 emitter property */
    SLRGPU_MAT_DIFFUSE_EMISSION = 11,   /* emitter property "diffuse": tex0 emittance            */
    SLRGPU_MAT_IBL_EMISSION = 12        /* environment: tex0 coeffM, f0 scale                    */
} SlrGpuMaterialKind;

/* 32 bytes */
typedef struct SlrGpuMaterial {
    uint32_t kind;
    uint32_t tex[4];
    uint32_t sub[2];
    float f0;
} SlrGpuMaterial;

/* One entry of an aggregate's light list (SurfaceObjectAggregate ctor, SurfaceObject.cpp:232-252):
 * object = prim_id of an emitting triangle, or 0x80000000|instance_id of an emitting instance.
 * pmf / cdf_lo / cdf_hi are this entry's slice of the aggregate's RegularConstantDiscrete1D
 * (distributions.cpp:81-119), built on the host with the reference's compensated sums:
 * pmf = importance / integral, cdf_lo = CDF[i], cdf_hi = CDF[i+1]. */
typedef struct SlrGpuLight {
    uint32_t object;
    float importance;
    float pmf, cdf_lo, cdf_hi;
    uint32_t pad[3];
} SlrGpuLight;

/* PerspectiveCamera (Cameras/PerspectiveCamera.cpp:15-74) with its static transform. */
typedef struct SlrGpuCamera {
    float mat[16];            /* camera -> world, column-major */
    float mat_inv[16];
    float sensitivity;
    float aspect, fov_y, lens_radius, img_plane_dist, obj_plane_dist;
    float pad[2];
} SlrGpuCamera;

/* InfiniteSphereSurfaceObject + IBLEmission (SurfaceObject.cpp:137-222, IBLEmission.cpp:15-17).
 * The importance map is the RegularConstantContinuous2D built by createIBLImportanceMap
 * (image_textures.cpp:81-134): `map_height` rows, each a 1D piecewise-constant distribution of
 * `map_width` cells; PDF and CDF arrays are the host-built ones (CDF has width+1 entries per row). */
typedef struct SlrGpuEnvironment {
    uint32_t present;
    uint32_t material;            /* IBL_EMISSION material id */
    uint32_t map_width, map_height;
    const float* row_pdf;         /* [map_height][map_width]   */
    const float* row_cdf;         /* [map_height][map_width+1] */
    const float* row_integral;    /* [map_height] */
    const float* marginal_pdf;    /* [map_height] */
    const float* marginal_cdf;    /* [map_height+1] */
    float marginal_integral;
    float pad;
} SlrGpuEnvironment;

/* Spectral constant tables the device code needs (built on the host by the same routines the
 * reference runs at start-up: initSpectrum, Spectrum.cpp:222 / SpectrumTypes.h:746-795). */
typedef struct SlrGpuSpectralTables {
    const float* upsample_grid;      /* Meng-Simon grid cells, raw table  (Spectrum.h:205-375)   */
    uint32_t upsample_grid_floats;
    const float* upsample_points;    /* Meng-Simon data points, raw table (Spectrum.h:384-571)   */
    uint32_t upsample_points_floats;
    const float* xbar_16;            /* 16-strata integrated CMFs (SpectrumTypes.h:746-795)      */
    const float* ybar_16;
    const float* zbar_16;
    float integral_cmf;
    float pad;
} SlrGpuSpectralTables;

typedef struct SlrGpuSceneDesc {
    uint32_t struct_size;            /* = sizeof(SlrGpuSceneDesc); rejects a mismatched ABI */
    uint32_t rgb_mode;               /* 0 = 16-wavelength spectral, 1 = RGB (references.h:45-60) */

    const SlrGpuBvhNode* bvh_nodes;        uint32_t num_bvh_nodes;     /* top-level root is node 0 */
    const SlrGpuLeafRecord* leaf_records;  uint32_t num_leaf_records;
    const SlrGpuInstance* instances;       uint32_t num_instances;
    const SlrGpuTriangle* triangles;       uint32_t num_triangles;
    const SlrGpuVertex* vertices;          uint32_t num_vertices;

    const SlrGpuMaterial* materials;       uint32_t num_materials;
    const SlrGpuTexture* textures;         uint32_t num_textures;
    const SlrGpuSpectrum* spectra;         uint32_t num_spectra;
    const float* spectrum_data;            uint32_t num_spectrum_floats;
    const SlrGpuImage* images;             uint32_t num_images;
    const uint8_t* image_data;             uint64_t image_data_bytes;

    /* light lists: entries [0, num_top_lights) belong to the top-level aggregate, nested
     * aggregates' lists follow (SlrGpuInstance::light_base). */
    const SlrGpuLight* lights;             uint32_t num_lights;
    uint32_t num_top_lights;
    float top_light_importance;      /* SurfaceObjectAggregate::importance() of the top-level aggregate */

    float world_center[3];           /* Scene::build, SurfaceObject.cpp:396-406 */
    float world_radius;

    SlrGpuCamera camera;
    SlrGpuEnvironment environment;
    SlrGpuSpectralTables spectral;

    /* optional: the binary SBVH of every aggregate (top level first: its root is node 0) with leaf records in the SBVH's own
     * leaf order, for slrgpu_intersect_batch_sbvh; NULL / 0 when not exported */
    const SlrGpuSbvhNode* sbvh_nodes;              uint32_t num_sbvh_nodes;
    const SlrGpuLeafRecord* sbvh_leaf_records;     uint32_t num_sbvh_leaf_records;

    /* optional: animated transforms (SlrGpuInstance::motion, camera_motion); camera_motion: 0 = static camera, else 1 +
     * index (camera.mat / mat_inv are then the begin key frame) */
    const SlrGpuMotion* motions;                   uint32_t num_motions;
    uint32_t camera_motion;                        uint32_t pad_motion;
} SlrGpuSceneDesc;

typedef struct SlrGpuScene SlrGpuScene;

/* ---------------------------------------------------------------------------------------------
 * Entry points
 * ------------------------------------------------------------------------------------------- */

/* Number of visible CUDA devices (0 when none / no driver). Never fails. */
SLRGPU_API int slrgpu_device_count(void);

/* Library/ABI version: (major << 16) | minor. */
SLRGPU_API uint32_t slrgpu_abi_version(void);

SLRGPU_API const char* slrgpu_last_error(void);

/* sizeof() of the ABI structs, so FFI bindings (ctypes, cgo, JNI) can verify their mirrors:
 * 0 SceneDesc, 1 BvhNode, 2 LeafRecord, 3 Instance, 4 Triangle, 5 Vertex, 6 Spectrum, 7 Texture,
 * 8 Image, 9 Material, 10 Light, 11 Camera, 12 Environment, 13 SpectralTables, 14 RayBatch,
 * 15 HitBatch, 16 RenderParams, 17 RenderStats, 18 SbvhNode, 19 Motion. Unknown index -> 0. */
SLRGPU_API uint32_t slrgpu_struct_size(int which);

/* Copies every buffer of `desc` to `device` (the caller keeps ownership of the host buffers, which
 * may be freed after the call returns). Geometry-only scenes (no materials/camera) are valid for
 * the intersect entry points. */
SLRGPU_API int slrgpu_scene_create(const SlrGpuSceneDesc* desc, int device, SlrGpuScene** out_scene);
SLRGPU_API void slrgpu_scene_destroy(SlrGpuScene* scene);
/* Device-memory footprint of the scene in bytes. */
SLRGPU_API uint64_t slrgpu_scene_device_bytes(const SlrGpuScene* scene);

/* Ray batches are SoA: eight float arrays. tmax may be +inf. */
typedef struct SlrGpuRayBatch {
    const float* org_x; const float* org_y; const float* org_z;
    const float* dir_x; const float* dir_y; const float* dir_z;
    const float* tmin;  const float* tmax;
} SlrGpuRayBatch;

/* Closest-hit results, SoA. prim = hit triangle's prim_id or SLRGPU_INVALID_ID on a miss;
 * inst = instance id of the outermost TransformedSurfaceObject or SLRGPU_INVALID_ID;
 * t = Intersection::dist; u, v = Intersection::u, ::v (b0 and b1, TriangleMesh.cpp:173-174).
 * Any of u, v, nodes_visited, tris_tested may be NULL. nodes_visited / tris_tested count QBVH nodes
 * popped and leaf records tested per ray (the algorithmic-bytes model of DESIGN.md). */
typedef struct SlrGpuHitBatch {
    uint32_t* prim; uint32_t* inst;
    float* t; float* u; float* v;
    uint32_t* nodes_visited; uint32_t* tris_tested;
} SlrGpuHitBatch;

/* Host buffers in, host buffers out. If kernel_ms is non-NULL it receives the device time of the
 * traversal kernel alone (CUDA events on the launch stream). */
SLRGPU_API int slrgpu_intersect_batch(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t num_rays,
                                      const SlrGpuHitBatch* hits, float* kernel_ms);
/* Device buffers in/out; enqueues on `stream` and returns without synchronising. Launches on different streams
 * (and from different threads) may be in flight at once -- each takes its own status slot on the scene's device, at
 * most 64 per scene. A traversal-stack overflow in such a launch is reported by slrgpu_scene_poll_overflow. */
SLRGPU_API int slrgpu_intersect_batch_device(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t num_rays,
                                             const SlrGpuHitBatch* hits, void* stream);
/* Waits for the scene's device and sets *overflow to 1 if any slrgpu_intersect_batch_device launch since the last
 * poll overflowed its traversal stack (QBVH.h:299 holds 64 entries; the reference would write out of bounds). */
SLRGPU_API int slrgpu_scene_poll_overflow(SlrGpuScene* scene, int* overflow);
/* Closest hits through the scene's binary SBVH instead of the QBVH: SBVH::intersect (Accelerator/SBVH.h:417-442) with
 * BoundingBox3D::intersect (Core/geometry.h:112-126) -- the accelerator the reference's shipped build traverses
 * (SurfaceObject.cpp:226-230). It visits leaves in a different order than the QBVH, so on rays that meet two primitives
 * at bit-equal distances (shared edges / vertices) it can report the other primitive: this entry point reproduces the
 * SBVH's answer, slrgpu_intersect_batch the QBVH's. Needs a scene created with sbvh_nodes (SLRGPU_ERR_INVALID_ARGUMENT
 * otherwise). Host buffers; u, v optional; the counters are not filled. */
SLRGPU_API int slrgpu_intersect_batch_sbvh(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t num_rays,
                                           const SlrGpuHitBatch* hits, float* kernel_ms);
/* Launch geometry the traversal kernel uses for n rays, for reporting (grid, block). */
SLRGPU_API int slrgpu_intersect_launch_config(SlrGpuScene* scene, uint64_t num_rays, uint32_t* grid, uint32_t* block);

/* occluded[i] = 1 if anything is hit in [tmin, tmax], else 0 -- the boolean of Scene::testVisibility
 * negated. Same traversal as the closest-hit query in the reference; here it may exit early. */
SLRGPU_API int slrgpu_occluded_batch(SlrGpuScene* scene, const SlrGpuRayBatch* rays, uint64_t num_rays,
                                     uint8_t* occluded, float* kernel_ms);

/* RenderSettings (RenderSettings.h:15-22) + PathTracingRenderer's spp, plus the sample range this
 * GPU renders (multi-GPU: spp is partitioned). */
typedef struct SlrGpuRenderParams {
    uint32_t struct_size;
    uint32_t width, height;
    uint32_t spp_begin, spp_end;     /* renders global sample indices [spp_begin, spp_end) for all pixels */
    float time_start, time_end;
    int32_t rng_seed;
    uint32_t max_path_length;        /* 0 = reference default (100, PathTracingRenderer.cpp:162) */
    uint32_t pool_size;              /* paths in flight (about 440 B of device memory each); 0 = default (16 Mi) */
    uint32_t flags;
} SlrGpuRenderParams;

typedef struct SlrGpuRenderStats {
    uint64_t paths;                  /* camera samples started */
    uint64_t rays;                   /* extend + shadow rays traced */
    uint64_t extend_rays, shadow_rays;
    uint64_t kernel_launches;
    float device_ms;                 /* first launch -> accumulators final, CUDA events */
    /* the rest only with SLRGPU_RENDER_PROFILE_STAGES (per-launch CUDA events; slows the run a little) */
    float raygen_ms, extend_ms, surface_ms, material_ms, shadow_ms, other_ms;   /* summed device time per kernel family */
    uint64_t waves;                  /* extend/shade iterations of the wavefront loop (always filled) */
    uint64_t extend_nodes, extend_leaf_records;   /* QBVH nodes popped / leaf records tested by extend rays */
    uint64_t shadow_nodes, shadow_leaf_records;   /* ... and by shadow rays (any-hit: stops at the first hit) */
    /* hits shaded per material class (always filled): 0 Lambert, 1 Oren-Nayar, 2 specular reflection, 3 specular
     * scattering, 4 Ward-Duer, 5 Ashikhmin-Shirley, 6 microfacet reflection, 7 microfacet scattering, 8 sum / mix / inverse */
    uint64_t class_hits[9];
    /* the tail kernel (the persistent kernel that finishes the last long paths of a call): paths it took over and
     * the longest run of bounces it made (always filled; 0 = it never ran); its device time only with PROFILE_STAGES */
    uint64_t tail_paths, tail_waves;
    float tail_ms, reserved0;
} SlrGpuRenderStats;

#define SLRGPU_RENDER_PROFILE_STAGES 0x1u
/* Bidirectional path tracing instead of unidirectional -- replaces BidirectionalPathTracingRenderer::render
 * (libSLR/Renderers/BidirectionalPathTracingRenderer.cpp:25-414; chosen by setRenderer("method": "BPT", "samples": n),
 * libSLRSceneGraph/API.cpp): per camera sample one light subpath, one eye subpath, every (s, t) connection with its
 * visibility ray and power-heuristic MIS weight; light-tracing connections (t = 1) land on the pixel the lens sees them in.
 * Same accumulation buffer, sample-range and multi-GPU semantics as the path tracer; max_path_length and pool_size are
 * ignored (the reference's BPT has no length cap; subpaths are cut at 64 vertices here). Stats: paths, rays (subpath +
 * connection rays), device_ms, class_hits[8] = connections examined, tail_paths = subpaths cut at the vertex limit. */
#define SLRGPU_RENDER_BPT 0x2u

/* Number of accumulation channels per pixel: 16 (spectral strata) or 3 (RGB). */
SLRGPU_API uint32_t slrgpu_scene_channels(const SlrGpuScene* scene);

/* Renders into accum[(y*width + x)*channels + c] (un-normalised sum over the rendered samples,
 * exactly what ImageSensor::pixel(x,y) holds after PathTracingRenderer::render). Host buffer. */
SLRGPU_API int slrgpu_render(SlrGpuScene* scene, const SlrGpuRenderParams* params, float* accum,
                             SlrGpuRenderStats* stats);
/* One frame on several GPUs of one box (the frame's samples-per-pixel are partitioned, the scene is
 * replicated, the float accumulation buffers are combined once per frame): `scenes` are `num_replicas` (1..16) replicas of
 * the same scene, normally one per device; replica g renders global sample indices
 * [spp_begin + g n / N, spp_begin + (g + 1) n / N), n = spp_end - spp_begin, on its own device from its own host thread;
 * the accumulation buffers are then summed onto scenes[0]'s device by one kernel reading the other devices' buffers over
 * NVLink peer access (staged peer copies where a pair has none) and downloaded into `accum` (host). stats: sums over the
 * replicas; device_ms = slowest replica + the exchange step, other_ms = the exchange step alone. Replaces nothing in the
 * reference (it renders on the host's threads, PathTracingRenderer.cpp:31-56); it is what Renderer::render calls when
 * more than one GPU is visible. */
SLRGPU_API int slrgpu_render_multi(SlrGpuScene* const* scenes, uint32_t num_replicas, const SlrGpuRenderParams* params, float* accum,
                                   SlrGpuRenderStats* stats);
/* Same, accumulating INTO a device buffer (not cleared), on `stream`, synchronised before return. */
SLRGPU_API int slrgpu_render_device(SlrGpuScene* scene, const SlrGpuRenderParams* params, float* accum_device,
                                    void* stream, SlrGpuRenderStats* stats);

/* Debug (AOV) renderer -- replaces DebugRenderer::render (libSLR/Renderers/DebugRenderer.cpp:29-217; chosen by
 * setRenderer("method": "debug", ("outputs": (...),)), libSLRSceneGraph/API.cpp:1037-1062): ONE camera sample per pixel
 * (global sample index params->spp_begin; spp_end must be greater), closest hit, Intersection::getSurfacePoint.
 * out[(y*width + x) * SLRGPU_DEBUG_FLOATS + k], host buffer: k = 0 hit (1) / miss (0), 1-3 geometric normal, 4-6 shading
 * normal, 7-9 shading tangent in world space -- the vectors the reference quantises as (uint8)clamp((0.5 v + 0.5) * 255)
 * into geometric_normal.bmp / shading_normal.bmp / shading_tangent.bmp. A miss leaves zeros (DebugInfo()). */
#define SLRGPU_DEBUG_FLOATS 10
SLRGPU_API int slrgpu_render_debug(SlrGpuScene* scene, const SlrGpuRenderParams* params, float* out, SlrGpuRenderStats* stats);

/* ---------------------------------------------------------------------------------------------
 * Shading probe (for parity tests): runs, for every probe ray, the same device functions the material
 * kernels run -- closest hit, surface point (normal map, instance transform), material -> BSDF with its
 * textures and spectra, BSDF::sample with the given random numbers, BSDF::evaluate / evaluatePDF for a
 * given world direction, emittance -- and returns the values the reference produces with
 * Intersection::getSurfacePoint, SurfacePoint::createBSDF, BSDF::sample/evaluate/evaluatePDF
 * (PathTracingRenderer.cpp:147-210). Spectral scenes only.
 * probes: n x 14 floats: org[3] dir[3] wlOffset uLambda uComponent uDir0 uDir1 evalDirWorld[3]
 * out:    n x 64 floats:
 *   [0] 0 miss / 1 surface hit / 2 environment   [1] t   [2..4] p   [5..7] shading normal   [8..10] shading tangent
 *   [11] hasNonDelta   [12..27] sampled fs   [28..30] sampled dir_sn   [31] dirPDF   [32] dirType
 *   [33..48] evaluated fs   [49] evaluated pdf   [50] isEmitting   [51..63] first 13 emittance values
 * ------------------------------------------------------------------------------------------- */
#define SLRGPU_PROBE_IN_FLOATS 14
#define SLRGPU_PROBE_OUT_FLOATS 64
SLRGPU_API int slrgpu_probe_shading(SlrGpuScene* scene, const float* probes, uint64_t num_probes, float* out);

/* The same for the queries of the bidirectional path tracer (BidirectionalPathTracingRenderer.cpp:184-196, 302-325): probe i
 * is a radiance query for even i and an importance query (BSDFQuery::adjoint = true) for odd i; sample() is asked for its
 * reverse information, evaluatePDF() for the reverse pdf. Same probe input layout.
 * out: n x 64 floats:
 *   [0] 0 miss / 1 surface hit / 2 environment   [1] t   [2..17] sampled fs   [18..20] sampled dir_sn   [21] dirPDF
 *   [22] dirType   [23..38] reverse->fs   [39] reverse->dirPDF (both 0 when nothing was sampled)
 *   [40] evaluatePDF   [41] its revPDF   [42..57] evaluate (the reference's rev_fs equals it for every model) */
SLRGPU_API int slrgpu_probe_shading_bpt(SlrGpuScene* scene, const float* probes, uint64_t num_probes, float* out);

/* Render calls keep their wavefront queues (about 440 bytes per path in flight) in a per-device pool
 * for reuse by later calls; this frees the pool. */
SLRGPU_API void slrgpu_release_workspaces(void);

#ifdef __cplusplus
}
#endif
#endif /* SLRGPU_H */
