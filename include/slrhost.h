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
