/* plain_assets.h - C-ABI of the asset side of the frame path (SURVEY.md 8f N1 + N2).
 *
 *   N1  .plain scene files and the R16F 3-D .dds bricks the reference's asset pipeline writes
 *       (Plain/src/Common/ModelLoadSaveBinary.cpp:40-231, Scene.h:6-19, MeshProcessing.cpp:20-120, ImageIO.cpp:342-431):
 *       loaded into host arrays (and from there into the frontend's SDF instance table)
 *   N2  the SDF bake of one mesh (Plain/src/AssetPipeline/SceneSDF.cpp:296-514 computeSDF: 15x15 rays per texel through a
 *       16^3 uniform grid, closest hit, back-face vote for the sign, distance to the closest triangle when no ray hits,
 *       stored as half floats) as a CUDA kernel. Results are bit-identical with the bricks the reference binary writes
 *       (tests/golden/sdf, produced by oracle/_ref/PlainAssetPipeline).
 *
 * Implemented twice like plain_b200.h: CUDA (libplain_b200.so, plain_asset_*) and the CPU oracle (oracle_asset_*,
 * -DPLAIN_ASSET_PREFIX=oracle_asset_). The loaders are host code shared by both.
 * All functions return 0 on success; plain_asset_last_error() returns the message of the last failure (thread-local).
 */
#ifndef PLAIN_ASSETS_H
#define PLAIN_ASSETS_H

#include "plain_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PLAIN_ASSET_PREFIX
#define PLAIN_ASSET_PREFIX plain_asset_
#endif
#define PLAIN_ASSET(name) PLAIN_CAT(PLAIN_ASSET_PREFIX, name)

PLAIN_EXPORT const char* PLAIN_ASSET(last_error)(void);

/* ---- N1: .plain scenes ---- */
typedef struct plain_scene plain_scene;
typedef struct {
    uint32_t index_count, vertex_count;
    float bb_min[3], bb_max[3];   /* AxisAlignedBoundingBox of the mesh (object space) */
    float mean_albedo[3];
    char albedo_path[256], normal_path[256], specular_path[256], sdf_path[256];
} plain_mesh_info;
typedef struct {
    float model_matrix[16];       /* column-major, glm::mat4 */
    uint64_t mesh_index;
} plain_scene_object;
PLAIN_EXPORT int PLAIN_ASSET(scene_load)(const char* path, plain_scene** out);
PLAIN_EXPORT void PLAIN_ASSET(scene_destroy)(plain_scene* scene);
PLAIN_EXPORT int PLAIN_ASSET(scene_counts)(const plain_scene* scene, uint64_t* out_objects, uint64_t* out_meshes);
PLAIN_EXPORT int PLAIN_ASSET(scene_object)(const plain_scene* scene, uint64_t index, plain_scene_object* out);
PLAIN_EXPORT int PLAIN_ASSET(scene_mesh_info)(const plain_scene* scene, uint64_t mesh, plain_mesh_info* out);
/* positions: vertex_count * 3 floats; indices: index_count u32 (16-bit indices of the file are widened) */
PLAIN_EXPORT int PLAIN_ASSET(scene_mesh_geometry)(const plain_scene* scene, uint64_t mesh, float* out_positions, uint32_t* out_indices);
/* the vertex buffer as stored: vertex_count * 28 bytes (position 3 x f32, uv 2 x f16, normal / tangent / bitangent A2R10G10B10_SNORM),
 * what plain_create_meshes / plain_frontend_set_mesh_geometry consume */
PLAIN_EXPORT int PLAIN_ASSET(scene_mesh_vertices)(const plain_scene* scene, uint64_t mesh, void* out_vertices);

/* ---- N1: R16F 3-D .dds bricks (148-byte DDS + DX10 header, DXGI_FORMAT_R16_FLOAT) ---- */
PLAIN_EXPORT int PLAIN_ASSET(dds_r16f_info)(const char* path, uint32_t out_extent[3]);
PLAIN_EXPORT int PLAIN_ASSET(dds_r16f_load)(const char* path, uint16_t* out_texels, size_t capacity_texels);
PLAIN_EXPORT int PLAIN_ASSET(dds_r16f_save)(const char* path, const uint32_t extent[3], const uint16_t* texels);

/* ---- N2: SDF bake ---- */
/* brick resolution of a mesh bounding box: 4 texels per metre, next power of two, clamped to [16, 64] (SceneSDF.cpp:117-131) */
PLAIN_EXPORT void PLAIN_ASSET(sdf_resolution)(const float bb_min[3], const float bb_max[3], uint32_t out_extent[3]);
/* bakes extent[0] * extent[1] * extent[2] half-float texels (x fastest) of the mesh inside its padded bounding box.
 * device: CUDA device ordinal (ignored by the oracle). out_kernel_ms (may be NULL): device time of the bake kernel. */
PLAIN_EXPORT int PLAIN_ASSET(sdf_bake)(int device, const float* positions, uint32_t vertex_count, const uint32_t* indices, uint32_t index_count,
                                       const float bb_min[3], const float bb_max[3], const uint32_t extent[3], uint16_t* out_texels, float* out_kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* PLAIN_ASSETS_H */
