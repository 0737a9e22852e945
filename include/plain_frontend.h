/* plain_frontend.h - C entry points of the host-side frame driver (mirror of the reference's RenderFrontend +
 * technique classes, plainrenderer_b200/host/), for callers that are not C++ (tests, bench.py). One frame =
 * main.cpp:79-90 of the reference: markNewFrame -> prepareNewFrame -> setCameraExtrinsic -> prepareForDrawcalls ->
 * renderScene -> renderFrame, with the raster passes' outputs supplied by the caller.
 *
 * Built twice from the same sources: against the CUDA backend (symbols plain_frontend_*) and, for tests/bench
 * cpu_baseline only, against the CPU oracle backend (symbols oracle_frontend_*; -DPLAIN_FRONTEND_PREFIX=oracle_frontend_).
 */
#ifndef PLAIN_FRONTEND_H
#define PLAIN_FRONTEND_H
#include "plain_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PLAIN_FRONTEND_PREFIX
#define PLAIN_FRONTEND_PREFIX plain_frontend_
#endif
#define PLAIN_FE(name) PLAIN_CAT(PLAIN_FRONTEND_PREFIX, name)

typedef struct plain_frontend plain_frontend;

/* defaults = the reference's default member initialisers (SURVEY.md section 5 "Config / flags") */
typedef struct {
    uint32_t width, height;
    int32_t diffuse_brdf;             /* 0 lambert, 1 disney, 2 CoD WWII (default), 3 titanfall 2 */
    int32_t direct_multiscatter;      /* 0 McAuley (default) */
    int32_t indirect_lighting_tech;   /* 0 SDF trace (default), 1 constant ambient */
    int32_t use_geometry_aa;          /* default 1 */
    int32_t sun_shadow_cascade_count; /* default 3 */
    int32_t half_res_trace;           /* default 1 */
    int32_t strict_influence_radius_cutoff; /* default 1 */
    float trace_influence_radius;     /* default 5 */
    int32_t taa_enabled;              /* default 1 */
    int32_t taa_use_clipping, taa_use_motion_vector_dilation, taa_history_sampling_tech, taa_filter_use_tonemapping; /* 1,1,4,1 */
    int32_t bloom_enabled;            /* default 1 */
    float bloom_strength, bloom_radius; /* 0.05, 1.5 */
    float sun_direction_deg[2];       /* (phi, theta) as RenderFrontend::m_sunDirection */
    float camera_fov_deg, camera_near, camera_far; /* 35, 0.1, 300 */
    uint32_t noise_seed;
    /* screen-space row sharding across GPUs: this frontend renders the rows of rank shard_rank out of
     * shard_count bands (bands are multiples of 64 rows); 0 / 0 or count 1 = the whole frame. The caller drives the frame
     * with begin_frame / run_segment and performs the exchanges between segments. */
    uint32_t shard_rank, shard_count;
    /* the non-default passes beside the frame path (single GPU only) */
    int32_t taa_use_separate_supersampling;  /* default 0: colorToLuminance.comp + temporalSupersampling.comp before the resolve (TAA.cpp:85-137) */
    int32_t taa_supersample_use_tonemapping; /* default 1 */
    int32_t sdf_debug_mode;                  /* SDFVisualisationMode: 0 none (default), 1 SDF, 2 camera tile usage, 3 normals, 4 raymarching count;
                                                != 0 replaces the frame by RenderFrontend.cpp:321-340 (sdfDebugVisualisation.comp + tonemap) */
    int32_t sdf_debug_show_tile_usage_with_hiz; /* default 1 */
    int32_t sdf_debug_use_influence_radius;     /* default 0 */
    /* 1 = depth / motion / normal, the shadow cascades and the packed G-buffer are rasterised from the meshes'
     * geometry (set_mesh_geometry) by the backend's graphic passes; render_frame is then called with inputs = NULL. Default 0:
     * the caller uploads them (plain_frame_inputs). Single GPU only. */
    int32_t raster_inputs;
} plain_frontend_settings;

/* host pointers to the outputs of the out-of-scope raster passes for one frame */
typedef struct {
    const void* depth;        /* D32F, w*h*4 (depthPrepass) */
    const void* motion;       /* RG16_SNORM, w*h*4 */
    const void* normal;       /* RGBA8 world-space geometric normal, w*h*4 */
    const void* gbuffer;      /* RGBA32_UINT packed G-buffer, w*h*16 (include/plain_frame_types.h) */
    const void* shadow_maps[4]; /* D16 2048x2048 each, may be NULL to keep the previous contents */
    int32_t async_upload;     /* 1: pointers are pinned host memory, copies go through the backend stream */
    /* rows of depth / normal / gbuffer to upload (the pointers always address row 0 of full-frame buffers); 0, 0 = all rows.
     * A sharded rank uploads its band plus the halo its stencils read. motion and shadow maps are always uploaded whole. */
    uint32_t row_begin, row_end;
} plain_frame_inputs;

typedef struct {
    float position[3], forward[3], right[3], up[3];
} plain_camera_extrinsic;

PLAIN_EXPORT void PLAIN_FE(default_settings)(plain_frontend_settings* out, uint32_t width, uint32_t height);
PLAIN_EXPORT int PLAIN_FE(create)(int device, const plain_frontend_settings* settings, plain_frontend** out);
PLAIN_EXPORT void PLAIN_FE(destroy)(plain_frontend* fe);
PLAIN_EXPORT const char* PLAIN_FE(last_error)(plain_frontend* fe);
PLAIN_EXPORT plain_ctx* PLAIN_FE(backend)(plain_frontend* fe);

/* scene: SDF meshes (3-D R16F bricks) and objects instancing them (RenderFrontend::registerMeshes / App scene) */
PLAIN_EXPORT int PLAIN_FE(register_sdf_mesh)(plain_frontend* fe, const uint16_t* r16f_texels, uint32_t rx, uint32_t ry, uint32_t rz,
                                             const float local_bb_min[3], const float local_bb_max[3], const float mean_albedo[3], uint32_t* out_mesh);
/* geometry + material of a registered mesh (RenderFrontend::registerMeshes, RenderFrontend.cpp:456-531): MeshBinary index / vertex
 * buffers (include/plain_b200.h plain_mesh_binary) and the bindless slots (plain_get_image_global_texture_array_index) of its RGBA8
 * albedo / normal / specular textures; PLAIN_INVALID_INDEX = the reference's 1x1 default texture (RenderFrontend.cpp:86-150) */
PLAIN_EXPORT int PLAIN_FE(set_mesh_geometry)(plain_frontend* fe, uint32_t mesh, const plain_mesh_binary* geometry, uint32_t albedo_texture, uint32_t normal_texture, uint32_t specular_texture);
PLAIN_EXPORT int PLAIN_FE(set_scene)(plain_frontend* fe, uint32_t n_objects, const uint32_t* mesh_indices, const float* model_matrices /* 16 each, column-major */,
                                     const float* bb_world_min /* 3 each */, const float* bb_world_max /* 3 each */);

/* ---- the host-side functions of the frame driver, exposed so that tests can hold them against the reference's own host code
 * (oracle/ref/ref_host_shim.cpp over Camera.cpp, ViewFrustum.cpp, Culling.cpp, MathUtils.cpp, sdfUtilities.cpp). Frustum layout: points
 * l_l_n, l_l_f, l_u_n, l_u_f, r_l_n, r_l_f, r_u_n, r_u_f; normals top, bot, right, left, near, far (ViewFrustum.h:6-30). ---- */
PLAIN_EXPORT void PLAIN_FE(host_hammersley2d)(uint32_t index, float out[2]);                                  /* MathUtils.cpp:25-70 */
PLAIN_EXPORT void PLAIN_FE(host_direction_to_vector)(const float angles_deg[2], float out[3]);                /* MathUtils.cpp:4-15 */
PLAIN_EXPORT uint32_t PLAIN_FE(host_mip_count_from_resolution)(uint32_t width, uint32_t height, uint32_t depth); /* MathUtils.cpp:17-19 */
PLAIN_EXPORT void PLAIN_FE(host_camera_matrices)(const plain_camera_extrinsic* camera, float fov_deg, float aspect, float near_plane, float far_plane,
                                                 float out_view[16], float out_projection[16]);               /* Camera.cpp:4-27 */
PLAIN_EXPORT void PLAIN_FE(host_view_frustum)(const plain_camera_extrinsic* camera, float fov_deg, float aspect, float near_plane, float far_plane,
                                              float out_points[24], float out_normals[18]);                   /* ViewFrustum.cpp:4-60 */
PLAIN_EXPORT void PLAIN_FE(host_orthogonal_frustum_fitted_to_camera)(const float points[24], const float normals[18], const float light_direction[3],
                                                                     float out_points[24], float out_normals[18]);  /* ViewFrustum.cpp:231-271 */
PLAIN_EXPORT int PLAIN_FE(host_aabb_intersects_frustum)(const float points[24], const float normals[18], const float bb_min[3], const float bb_max[3]); /* Culling.cpp:5-42 */
PLAIN_EXPORT void PLAIN_FE(host_pad_sdf_bounding_box)(const float bb_min[3], const float bb_max[3], float out_min[3], float out_max[3]); /* sdfUtilities.cpp:5-19 */

PLAIN_EXPORT void PLAIN_FE(host_sdf_world_to_local)(const float model_matrix[16], const float bb_offset[3], float out[16]);  /* SDFGI.cpp:288-292 */

/* draw calls recorded for the last frame: out[0] main pass / prepass (after camera-frustum culling), out[1] per shadow cascade (after
 * culling against the sun shadow frustum): the reference's m_currentMainPassDrawcallCount / m_currentShadowPassDrawcallCount */
PLAIN_EXPORT int PLAIN_FE(get_drawcall_counts)(plain_frontend* fe, uint32_t out[2]);

/* one frame */
PLAIN_EXPORT int PLAIN_FE(render_frame)(plain_frontend* fe, const plain_camera_extrinsic* camera, float time, float delta_time, const plain_frame_inputs* inputs);
/* ---- row-sharded frames: one frame = begin_frame, then run_segment until it reports no pending exchange. Between two
 * segments the caller performs the exchange described by *out over its communicator (NCCL / gloo / local copies): every
 * rank calls the same sequence. render_frame == begin_frame + run_segment (a frontend with shard_count <= 1 never
 * reports an exchange). ---- */
typedef enum {
    PLAIN_EXCHANGE_NONE = 0,
    PLAIN_EXCHANGE_ALLREDUCE_SUM_U32 = 1, /* storage buffer of u32 counters: element-wise sum over ranks (luminance histogram) */
    PLAIN_EXCHANGE_ALLGATHER_ROWS = 2,    /* every rank contributes the rows of its band, afterwards all ranks hold all rows */
    PLAIN_EXCHANGE_HALO_ROWS = 3          /* halo_rows rows on each side of the band boundary are copied from the neighbouring ranks */
} plain_exchange_kind;
typedef struct {
    uint32_t kind;          /* plain_exchange_kind */
    uint32_t n_images;      /* images exchanged together (e.g. Y_SH + CoCg) */
    void* device_ptr[4];    /* base of the image level (row 0); for ALLREDUCE device_ptr[0] = the buffer */
    uint32_t row_pitch_bytes[4];
    uint32_t rows[4];       /* height of the level */
    uint32_t row_divisor[4];/* the band of rank r in this level is [Y0(r) / divisor, ceil(Y1(r) / divisor)) clipped to rows, see shard_band */
    uint32_t halo_rows;
    uint32_t element_count; /* ALLREDUCE: number of u32 */
    char name[32];
    plain_image_handle image[4]; /* the exchanged images (for peer_get_image_handle / peer_open_image) */
    uint32_t mip_level[4];
    plain_handle buffer;         /* ALLREDUCE: the storage buffer */
    uint32_t depth[4];           /* slices of the level (1 for 2-D images): a row range means those rows in EVERY slice; slices are
                                    rows * row_pitch_bytes apart */
    uint32_t deferred;           /* 1: only the NEXT frame reads the gathered rows - the exchange may complete any time before the next
                                    begin_frame (over peer exchange it runs behind the frame; a caller may also perform it right away) */
} plain_exchange;
PLAIN_EXPORT int PLAIN_FE(begin_frame)(plain_frontend* fe, const plain_camera_extrinsic* camera, float time, float delta_time, const plain_frame_inputs* inputs);
PLAIN_EXPORT int PLAIN_FE(run_segment)(plain_frontend* fe, plain_exchange* out);
/* With peer exchange enabled (after plain_peer_init and the sync blocks are mapped, include/plain_b200.h) run_segment performs
 * every exchange whose images are mapped on all ranks itself - row pushes into the peers' images + a flag barrier, enqueued
 * on the pass stream - and only returns the ones it cannot do yet, so that the caller can map their images (and perform that
 * one exchange itself). Once everything is mapped a whole frame is one run_segment call without a host round trip. */
PLAIN_EXPORT int PLAIN_FE(set_peer_exchange)(plain_frontend* fe, int32_t enabled);
/* rows [*out_begin, *out_end) of an image level with `rows` rows and the given divisor that belong to `rank` */
PLAIN_EXPORT void PLAIN_FE(shard_band)(uint32_t full_height, uint32_t shard_count, uint32_t rank, uint32_t divisor, uint32_t rows, uint32_t* out_begin, uint32_t* out_end);
PLAIN_EXPORT int PLAIN_FE(read_output_rows)(plain_frontend* fe, void* out_full_frame, uint32_t row_begin, uint32_t row_end, int32_t async_pinned);

/* read back the tonemapped B8G8R8A8 frame (w*h*4) */
PLAIN_EXPORT int PLAIN_FE(read_output)(plain_frontend* fe, void* out, size_t size, int32_t async_pinned);

/* named resources, for parity tests: images "color0|color1|depth0|depth1|motion0|motion1|post0|post1|normal|gbuffer|depthHalf|hiz|
 * brdfLut|skyTransmission|skyMultiscatter|skyLut|shadow0..3|giY0|giY1|giC0|giC1|giHistY0|giHistY1|giHistC0|giHistC1|giFullY|giFullC|
 * froxelMaterial|froxelScatter|froxelHist0|froxelHist1|froxelIntegration|taaHist0|taaHist1|taaLum0|taaLum1|bloomDown|bloomUp|output";
 * buffers "histogram|histogramPerTile|light|sunShadowInfo|sdfInstances|sdfCulled|sdfTiles" */
PLAIN_EXPORT int PLAIN_FE(get_image)(plain_frontend* fe, const char* name, plain_image_handle* out);
PLAIN_EXPORT int PLAIN_FE(get_storage_buffer)(plain_frontend* fe, const char* name, plain_handle* out);
PLAIN_EXPORT int PLAIN_FE(get_global_shader_info)(plain_frontend* fe, void* out_340_bytes);
PLAIN_EXPORT int PLAIN_FE(get_resolve_weights)(plain_frontend* fe, float out[9]);
/* preset the exposure state (LightBuffer.previousFrameExposure) so short test sequences start adapted */
PLAIN_EXPORT int PLAIN_FE(set_exposure)(plain_frontend* fe, float previous_frame_exposure);

/* ---- synthetic scene (stand-in for the .plain loader + raster passes): boxes with analytic SDF bricks,
 * and a CPU ray caster producing depth / motion / normal / G-buffer / shadow maps for a camera. Host-only. ---- */
typedef struct plain_synthetic_scene plain_synthetic_scene;
PLAIN_EXPORT int PLAIN_FE(synthetic_scene_create)(uint32_t seed, uint32_t n_instances, plain_synthetic_scene** out);
PLAIN_EXPORT void PLAIN_FE(synthetic_scene_destroy)(plain_synthetic_scene* s);
/* registers meshes + objects of the scene with the frontend */
PLAIN_EXPORT int PLAIN_FE(synthetic_scene_attach)(plain_synthetic_scene* s, plain_frontend* fe);
/* ray-casts the raster-pass outputs for one frame into caller buffers (sizes as in plain_frame_inputs). frame_index selects the
 * TAA jitter (Halton(2,3)[frame_index % 8]); previous_camera may be NULL (static camera). shadow maps are rendered when the
 * pointers are non-NULL. threads = 0 -> all cores */
PLAIN_EXPORT int PLAIN_FE(synthetic_scene_render_inputs)(plain_synthetic_scene* s, const plain_frontend_settings* settings, const plain_camera_extrinsic* camera,
                                                         const plain_camera_extrinsic* previous_camera, uint32_t frame_index, void* depth, void* motion, void* normal,
                                                         void* gbuffer, void* const shadow_maps[4], int32_t threads);

#ifdef __cplusplus
}
#endif
#endif
