/*
 * slrhost.h -- C entry points of the host library (libslrhost.so): the C++ mirror of libSLR /
 * libSLRSceneGraph's scene side (scene graph, flattening, SBVH -> QBVH build, renderer front end).
 * These exist so that tests, bench.py and other FFI users can drive the host C++ code; the GPU work
 * itself always goes through include/slrgpu.h.
 *
 * Return convention: 0 on success, negative on error; slrhost_last_error() has the message.
 */
#ifndef SLRHOST_H
#define SLRHOST_H
#include "slrgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct SlrHostBuilder SlrHostBuilder;   /* a scene graph under construction */
typedef struct SlrHostScene SlrHostScene;       /* a flattened scene (owns the SoA buffers) */

SLRGPU_API const char* slrhost_last_error(void);

/* Process-wide options for scenes built / read afterwards. "export_sbvh" (0 / 1, default 0): the flattened scene also
 * carries every aggregate's binary SBVH (SlrGpuSceneDesc::sbvh_nodes) for slrgpu_intersect_batch_sbvh. */
SLRGPU_API int slrhost_set_option(const char* name, int value);

/* --- programmatic scene graph (what the scene language's builtins do, API.cpp:663-800) --- */
SLRGPU_API SlrHostBuilder* slrhost_builder_create(void);
SLRGPU_API void slrhost_builder_destroy(SlrHostBuilder* b);
/* Adds a triangle mesh (TriangleMeshNode). positions: 3 floats per vertex; normals, tangents (3 per
 * vertex) and uvs (2 per vertex) may be NULL (then normal = (0,1,0), tangent = (1,0,0), uv = 0).
 * Returns the mesh id (>= 0) or a negative error. The mesh is not part of the scene until placed. */
SLRGPU_API int slrhost_builder_add_mesh(SlrHostBuilder* b, const float* positions, const float* normals,
                                        const float* tangents, const float* uvs, uint32_t num_vertices,
                                        const uint32_t* indices, uint32_t num_triangles);
/* Places mesh under the root inside an InternalNode with the column-major transform `mat` (NULL =
 * identity); the transform is baked into the vertices (nodes.cpp:110-123, TriangleMeshNode.cpp:68-78).
 * A mesh can be placed once. */
SLRGPU_API int slrhost_builder_place_mesh(SlrHostBuilder* b, int mesh, const float* mat);
/* Instances mesh through a ReferenceNode (nodes.cpp:174-184) with transform `mat`; may be called
 * many times per mesh -- all instances share one nested aggregate. */
SLRGPU_API int slrhost_builder_instance_mesh(SlrHostBuilder* b, int mesh, const float* mat);
/* Flattens and builds (SBVH -> QBVH). rgb_mode: 0 spectral, 1 RGB. */
SLRGPU_API int slrhost_builder_finish(SlrHostBuilder* b, int rgb_mode, SlrHostScene** out);

/* --- scene description language (libSLRSceneGraph/API.hpp:20 readScene + Scene::build) --- */
/* Parses and executes a scene file, flattens it and builds the acceleration structures.
 * rgb_mode: 0 = 16-wavelength spectral, 1 = RGB. */
SLRGPU_API int slrhost_read_scene(const char* path, int rgb_mode, SlrHostScene** out);
/* Rendering context read from the file (setRenderer / setRenderSettings):
 * ctx8 = {width, height, samples, rngSeed, timeStart, timeEnd, brightness, 1 if a renderer was set}. */
SLRGPU_API int slrhost_scene_context(const SlrHostScene* s, double* ctx8);
/* Renderer::render through GPUPathTracingRenderer on `device` (>= 0), or on EVERY visible device with the frame's
 * samples partitioned over them (device < 0; slrgpu_render_multi): width/height/spp <= 0 take the
 * scene file's values. On return accum (if non-NULL, width*height*channels floats) holds the
 * sensor's un-normalised sums; stats6 = {paths, rays, deviceSeconds, wallSeconds, uploadSeconds, channels + 1000 * devices used}.
 * bmp_dir: directory for the progressive NNN.bmp files, or NULL to skip image export. */
SLRGPU_API int slrhost_render(SlrHostScene* s, int device, int width, int height, int spp, int seed,
                              const char* bmp_dir, float* accum, double* stats6);
/* Same, for the global sample indices [spp_begin, spp_begin + spp): what process g of an N-process
 * front end calls before the sensors are summed (one process per GPU). */
SLRGPU_API int slrhost_render_range(SlrHostScene* s, int device, int width, int height, int spp_begin, int spp, int seed,
                                    const char* bmp_dir, float* accum, double* stats6);
/* Renderer::render through GPUBidirectionalPathTracingRenderer (the reference's BidirectionalPathTracingRenderer,
 * setRenderer("method": "BPT"); include/slrgpu.h SLRGPU_RENDER_BPT): arguments and results as slrhost_render_range. */
SLRGPU_API int slrhost_render_bpt(SlrHostScene* s, int device, int width, int height, int spp_begin, int spp, int seed,
                                  const char* bmp_dir, float* accum, double* stats6);
/* The method the scene file's setRenderer named ("PT", "BPT", "debug"; empty when it set none), NUL-terminated. */
SLRGPU_API int slrhost_scene_renderer_method(const SlrHostScene* s, char* method, uint32_t capacity);
/* Renderer::render through GPUDebugRenderer (the reference's DebugRenderer, setRenderer("method": "debug")): one camera
 * sample per pixel; out (if non-NULL) receives width*height*SLRGPU_DEBUG_FLOATS floats (include/slrgpu.h
 * slrgpu_render_debug), bmp_dir (if non-NULL) geometric_normal.bmp, shading_normal.bmp and shading_tangent.bmp.
 * stats6 as slrhost_render. */
SLRGPU_API int slrhost_render_debug(SlrHostScene* s, int device, int width, int height, int seed,
                                    const char* bmp_dir, float* out, double* stats6);
/* Tone-maps a frame buffer exactly like ImageSensor::saveImage (ImageSensor.cpp:138-186). */
SLRGPU_API int slrhost_save_bmp(const char* path, const float* accum, int width, int height, int channels,
                                float scale, float sensitivity);
/* linear sRGB (3 floats per pixel) of a frame buffer, before tone mapping */
SLRGPU_API int slrhost_accum_to_rgb(const float* accum, int width, int height, int channels, float scale, float* rgb);

/* --- synthetic asset writers (tests / bench build their own models and environment maps) --- */
SLRGPU_API int slrhost_write_assbin(const char* path, const float* positions, const float* normals,
                                    const float* tangents, const float* uvs, uint32_t num_vertices,
                                    const uint32_t* indices, uint32_t num_triangles, const char* material_name,
                                    const float* diffuse_rgb);
SLRGPU_API int slrhost_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgba);
/* The EXR reader behind Image2D / setEnvironment (scanline, uncompressed, half or float channels): width*height*4 floats
 * (R, G, B, A; a missing A reads 1), top row first. rgba may be NULL to query the size. */
SLRGPU_API int slrhost_read_exr(const char* path, uint32_t* width, uint32_t* height, float* rgba, uint64_t capacity_floats);

/* Decodes a PNG the way the reference's loadPNG asks libpng to (image_loader.cpp:186-280: strip 16 -> 8 bits, unpack
 * 1/2/4-bit samples, palette -> RGB, 0xFF filler after RGB, libpng's 8-bit gamma table for screen gamma 1.0 -- or 2.2 with
 * gamma_correction -- against the file's gAMA or 0.45455). *channels = 1 (grey) or 4, | 0x100 when the 4th byte is the
 * file's alpha. pixels (may be NULL to query the size) receives width*height*(channels & 0xFF) bytes, top row first. */
SLRGPU_API int slrhost_decode_png(const char* path, int gamma_correction, uint32_t* width, uint32_t* height, uint32_t* channels,
                                  uint8_t* pixels, uint64_t capacity);

/* The host's AnimatedTransform (motion blur; libSLR/Core/Transform.h:89-144) for tests: decomposition of the two key
 * matrices (T0[3] R0[4] S0[16] T1[3] R1[4] S1[16] = 46 floats), motionBounds of a box, and the transform sampled at n
 * times (mat[16], matInv[16] each, column-major). Any output may be NULL. */
SLRGPU_API int slrhost_sample_animated(const float* mat_begin, const float* mat_end, float t_begin, float t_end, const float* box6,
                                       const float* times, uint32_t n, float* decomposition46, float* bounds6, float* out32n);

SLRGPU_API void slrhost_scene_destroy(SlrHostScene* s);
/* Fills `desc` with pointers into the scene's buffers (valid until slrhost_scene_destroy). */
SLRGPU_API int slrhost_scene_describe(const SlrHostScene* s, SlrGpuSceneDesc* desc);
/* stats[0..9] of aggregate i (0 = top level): numObjects, sbvhNodes, sbvhRefs, sbvhDepth, qbvhNodes,
 * qbvhDepth, nodeBase, leafBase as doubles, then sbvhCost, qbvhCost. Returns number of aggregates. */
SLRGPU_API int slrhost_scene_stats(const SlrHostScene* s, uint32_t aggregate, double* stats10);
SLRGPU_API double slrhost_scene_build_seconds(const SlrHostScene* s);

#ifdef __cplusplus
}
#endif
#endif
