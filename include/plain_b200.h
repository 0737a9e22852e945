/* plain_b200.h - C-ABI of the B200 frame-path backend.
 *
 * Drop-in boundary for the per-pixel frame path of Gaukler/PlainRenderer: every entry point replaces one
 * public method of `class RenderBackend` (reference Plain/src/Runtime/Rendering/Backend/RenderBackend.h:33-110)
 * for the compute-pass subset the frame path uses. Passes are identified by the reference's shader file name
 * (ResourceDescriptions.h:112-149), resources by integer handle + per-shader binding number
 * (ResourceDescriptions.h:9-78), executions are replayed in submission order (RenderBackend.cpp:259-265,
 * 769-786). Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * All functions return 0 on success, non-zero on failure; plain_last_error() returns the message
 * (the reference prints + throws instead, RenderBackend.cpp:442-445). Nothing throws across the ABI.
 *
 * The same header is implemented twice: by the CUDA backend (libplain_b200.so, symbols plain_*) and by the
 * CPU oracle used only as the checker in tests/bench (oracle/, symbols oracle_*; select with
 * -DPLAIN_FN_PREFIX=oracle_).
 */
#ifndef PLAIN_B200_H
#define PLAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PLAIN_FN_PREFIX
#define PLAIN_FN_PREFIX plain_
#endif
#define PLAIN_CAT2(a, b) a##b
#define PLAIN_CAT(a, b) PLAIN_CAT2(a, b)
#define PLAIN_FN(name) PLAIN_CAT(PLAIN_FN_PREFIX, name)

#if defined(__GNUC__)
#define PLAIN_EXPORT __attribute__((visibility("default")))
#else
#define PLAIN_EXPORT
#endif

typedef struct plain_ctx plain_ctx;

#define PLAIN_INVALID_INDEX 0xFFFFFFFFu /* RenderHandles.h:4 */

/* ImageDescription.h:4-5 */
typedef enum { PLAIN_IMAGE_TYPE_1D = 0, PLAIN_IMAGE_TYPE_2D = 1, PLAIN_IMAGE_TYPE_3D = 2, PLAIN_IMAGE_TYPE_CUBE = 3 } plain_image_type;
typedef enum { PLAIN_MIPS_ONE = 0, PLAIN_MIPS_FULL_CHAIN = 1, PLAIN_MIPS_MANUAL = 2, PLAIN_MIPS_FULL_CHAIN_ALREADY_IN_DATA = 3 } plain_mip_count;
/* ImageDescription.h:7-11 */
enum { PLAIN_USAGE_STORAGE = 1, PLAIN_USAGE_SAMPLED = 2, PLAIN_USAGE_ATTACHMENT = 4 };

/* ImageDescription.h:16, same order. PLAIN_FORMAT_RGBA32_UINT is new: the packed G-buffer texel. */
typedef enum {
    PLAIN_FORMAT_R8 = 0,
    PLAIN_FORMAT_RG8,
    PLAIN_FORMAT_RGBA8,
    PLAIN_FORMAT_R16_SFLOAT,
    PLAIN_FORMAT_RG16_SFLOAT,
    PLAIN_FORMAT_RG32_SFLOAT,
    PLAIN_FORMAT_RG16_SNORM,
    PLAIN_FORMAT_RGBA16_SFLOAT,
    PLAIN_FORMAT_RGBA16_SNORM,
    PLAIN_FORMAT_RGBA32_SFLOAT,
    PLAIN_FORMAT_R11G11B10_UFLOAT,
    PLAIN_FORMAT_DEPTH16,
    PLAIN_FORMAT_DEPTH32,
    PLAIN_FORMAT_BC1,
    PLAIN_FORMAT_BC3,
    PLAIN_FORMAT_BC5,
    PLAIN_FORMAT_BGRA8_UNORM,
    PLAIN_FORMAT_RGBA32_UINT,
    PLAIN_FORMAT_COUNT
} plain_image_format;

/* ImageDescription.h:18-30 */
typedef struct {
    uint32_t width, height, depth;
    uint32_t type;             /* plain_image_type */
    uint32_t format;           /* plain_image_format */
    uint32_t usage_flags;      /* PLAIN_USAGE_* */
    uint32_t mip_count;        /* plain_mip_count */
    uint32_t manual_mip_count; /* only if mip_count == PLAIN_MIPS_MANUAL */
    uint32_t auto_create_mips;
} plain_image_desc;

/* RenderHandles.h:16-21 */
typedef enum { PLAIN_IMAGE_HANDLE_DEFAULT = 0, PLAIN_IMAGE_HANDLE_TRANSIENT = 1, PLAIN_IMAGE_HANDLE_SWAPCHAIN = 2 } plain_image_handle_type;
typedef struct {
    uint32_t type; /* plain_image_handle_type */
    uint32_t index;
} plain_image_handle;
typedef uint32_t plain_handle; /* {uint32_t index} handles: pass, sampler, uniform buffer, storage buffer */

/* ResourceDescriptions.h:163-172 */
typedef enum { PLAIN_SAMPLER_NEAREST = 0, PLAIN_SAMPLER_LINEAR = 1 } plain_sampler_interpolation;
typedef enum { PLAIN_WRAP_CLAMP = 0, PLAIN_WRAP_COLOR = 1, PLAIN_WRAP_REPEAT = 2 } plain_sampler_wrapping;
typedef enum { PLAIN_BORDER_WHITE = 0, PLAIN_BORDER_BLACK = 1 } plain_sampler_border;
typedef struct {
    uint32_t interpolation, wrapping, use_anisotropy;
    float max_anisotropy;
    uint32_t border_color, max_mip;
} plain_sampler_desc;

/* ResourceDescriptions.h:112-120 */
typedef struct {
    uint32_t location;
    const void* data;
    uint32_t size;
} plain_spec_const;

/* ResourceDescriptions.h:9-53 */
typedef struct { plain_handle buffer; uint32_t read_only; uint32_t binding; } plain_storage_buffer_resource;
typedef struct { plain_handle buffer; uint32_t binding; } plain_uniform_buffer_resource;
typedef struct { plain_image_handle image; uint32_t mip_level; uint32_t binding; } plain_image_resource;
typedef struct { plain_handle sampler; uint32_t binding; } plain_sampler_resource;

typedef struct {
    const plain_sampler_resource* samplers; uint32_t n_samplers;
    const plain_storage_buffer_resource* storage_buffers; uint32_t n_storage_buffers;
    const plain_uniform_buffer_resource* uniform_buffers; uint32_t n_uniform_buffers;
    const plain_image_resource* sampled_images; uint32_t n_sampled_images;
    const plain_image_resource* storage_images; uint32_t n_storage_images;
} plain_pass_resources;

/* ComputePassExecution, ResourceDescriptions.h:73-78 */
typedef struct {
    plain_handle pass;
    plain_pass_resources resources;
    const void* push_constants; uint32_t push_constant_size;
    uint32_t dispatch_count[3];
    /* additions for screen-space row sharding across GPUs (not in the reference; 0/0/0 = the whole pass):
     * rows [row_begin, row_end) of the pass's output image are produced (units: rows of that image; for tile-based
     * passes rows of tiles / workgroups as documented per kernel); shard_phase splits passes that mix a banded and a
     * replicated part (depthHiZPyramid.comp: 1 = the levels reduced from the rank's own depth rows, 2 = the rest) */
    uint32_t row_begin, row_end, shard_phase;
} plain_compute_pass_execution;

/* ---- graphic passes (the rasterisation passes that feed the frame path) ----
 * MeshBinary, MeshData.h:27-35: index buffer u16 when index_count < 65535 else u32 (RenderBackend.cpp:483-488); vertex buffer
 * 28 bytes per vertex: position 3 x f32, uv 2 x f16, normal / tangent / bitangent A2R10G10B10_SNORM (VertexInput.h:27-31,
 * VulkanVertexInput.cpp:4-10, MeshProcessing.cpp:52-106). Both are copied before the call returns. */
typedef struct {
    uint32_t index_count, vertex_count;
    const void* index_buffer;
    const void* vertex_buffer;
} plain_mesh_binary;

/* ResourceDescriptions.h:80-103 */
typedef enum { PLAIN_CULL_NONE = 0, PLAIN_CULL_FRONT = 1, PLAIN_CULL_BACK = 2 } plain_cull_mode;
typedef enum { PLAIN_DEPTH_NEVER = 0, PLAIN_DEPTH_ALWAYS, PLAIN_DEPTH_LESS, PLAIN_DEPTH_GREATER, PLAIN_DEPTH_LESS_EQUAL, PLAIN_DEPTH_GREATER_EQUAL, PLAIN_DEPTH_EQUAL } plain_depth_function;
typedef enum { PLAIN_LOAD_OP_LOAD = 0, PLAIN_LOAD_OP_CLEAR = 1, PLAIN_LOAD_OP_DONT_CARE = 2 } plain_attachment_load_op;
typedef struct { uint32_t format; uint32_t load_op; } plain_attachment;
/* GraphicPassDescription, ResourceDescriptions.h:129-143 (vertex + fragment stage; fill mode, no blending, VertexFormat::Full).
 * Passes that exist as CUDA rasteriser programs (csrc/passes_raster.cu), by shader pair:
 *   depthPrepass.vert + depthPrepass.frag   attachments {RG16_SNORM motion, RGBA8 normal, DEPTH32}, cull back, GREATER_EQUAL
 *   sunShadow.vert    + sunShadow.frag      attachment  {DEPTH16}, cull front, depth clamp, GREATER_EQUAL; spec constant 0 = cascade
 *   triangle.vert     + gbufferFill.frag    attachments {RGBA32_UINT packed G-buffer, DEPTH32 (load)}, depth EQUAL against the
 *                                           prepass: the raster half of the reference's main pass (triangle.frag:178-193: the
 *                                           interpolated inputs and material texels), recast for the deferred shading kernel */
typedef struct {
    const char* vertex_shader; const plain_spec_const* vertex_consts; uint32_t n_vertex_consts;
    const char* fragment_shader; const plain_spec_const* fragment_consts; uint32_t n_fragment_consts;
    const plain_attachment* attachments; uint32_t n_attachments;
    uint32_t cull_mode;      /* plain_cull_mode; front face = counter clockwise (VulkanPipeline.cpp:61) */
    uint32_t clamp_depth;
    uint32_t depth_function; /* plain_depth_function */
    uint32_t depth_write;
    const char* debug_name;
} plain_graphic_pass_desc;
/* GraphicPassExecution, ResourceDescriptions.h:55-71: targets in attachment order */
typedef struct { plain_image_handle image; uint32_t mip_level; } plain_render_target;
typedef struct {
    plain_handle pass;
    plain_pass_resources resources;
    const plain_render_target* targets; uint32_t n_targets;
    /* addition for screen-space row sharding (0 / 0 = all rows): only rows [row_begin, row_end) of the targets are rasterised and
     * resolved, the other rows keep their contents (a rank renders its band + the halo its stencils read) */
    uint32_t row_begin, row_end;
} plain_graphic_pass_execution;

/* VulkanTimestampQueries.h:16-20 */
typedef struct {
    char name[64];
    float time_ms;
} plain_pass_time;

/* ---- lifetime: RenderBackend::setup / shutdown (RenderBackend.cpp:47-92); output size replaces the swapchain ---- */
PLAIN_EXPORT int PLAIN_FN(backend_create)(int device, uint32_t width, uint32_t height, plain_ctx** out_ctx);
PLAIN_EXPORT void PLAIN_FN(backend_destroy)(plain_ctx* ctx);
PLAIN_EXPORT const char* PLAIN_FN(last_error)(plain_ctx* ctx);
/* recreateSwapchain (RenderBackend.h:37) */
PLAIN_EXPORT int PLAIN_FN(recreate_swapchain)(plain_ctx* ctx, uint32_t width, uint32_t height);

/* ---- resources (RenderBackend.h:84-110) ---- */
PLAIN_EXPORT int PLAIN_FN(create_image)(plain_ctx* ctx, const plain_image_desc* desc, const void* initial_data, size_t initial_data_size, plain_image_handle* out);
PLAIN_EXPORT int PLAIN_FN(create_temporary_image)(plain_ctx* ctx, const plain_image_desc* desc, plain_image_handle* out);
PLAIN_EXPORT int PLAIN_FN(resize_images)(plain_ctx* ctx, const plain_image_handle* images, uint32_t n, uint32_t width, uint32_t height);
PLAIN_EXPORT int PLAIN_FN(get_image_description)(plain_ctx* ctx, plain_image_handle image, plain_image_desc* out);
PLAIN_EXPORT int PLAIN_FN(get_image_global_texture_array_index)(plain_ctx* ctx, plain_image_handle image, uint32_t* out);
PLAIN_EXPORT int PLAIN_FN(create_uniform_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out);
PLAIN_EXPORT int PLAIN_FN(create_storage_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out);
PLAIN_EXPORT int PLAIN_FN(create_sampler)(plain_ctx* ctx, const plain_sampler_desc* desc, plain_handle* out);
PLAIN_EXPORT int PLAIN_FN(get_swapchain_input_image)(plain_ctx* ctx, plain_image_handle* out);

/* ---- passes: createComputePass / updateComputePassShaderDescription (RenderBackend.h:77-96) ---- */
PLAIN_EXPORT int PLAIN_FN(create_compute_pass)(plain_ctx* ctx, const char* shader_src_path_relative, const plain_spec_const* consts, uint32_t n_consts, const char* debug_name, plain_handle* out);
PLAIN_EXPORT int PLAIN_FN(update_compute_pass_shader_description)(plain_ctx* ctx, plain_handle pass, const char* shader_src_path_relative, const plain_spec_const* consts, uint32_t n_consts);
/* setGlobalDescriptorSetResources (RenderBackend.h:73): set 0 = global UBO + the 8 immutable samplers (global.inc:4-42) */
PLAIN_EXPORT int PLAIN_FN(set_global_descriptor_set_resources)(plain_ctx* ctx, const plain_pass_resources* resources);

/* ---- frame protocol (RenderBackend.h:49-82) ---- */
PLAIN_EXPORT int PLAIN_FN(new_frame)(plain_ctx* ctx);
PLAIN_EXPORT int PLAIN_FN(set_compute_pass_execution)(plain_ctx* ctx, const plain_compute_pass_execution* execution);
PLAIN_EXPORT int PLAIN_FN(prepare_for_drawcall_recording)(plain_ctx* ctx);
/* data is copied before the call returns; the device copy lands before any pass of the frame (RenderBackend.cpp:315-321, 896-911) */
PLAIN_EXPORT int PLAIN_FN(set_uniform_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size);
PLAIN_EXPORT int PLAIN_FN(set_storage_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size);
PLAIN_EXPORT int PLAIN_FN(render_frame)(plain_ctx* ctx, int present_to_screen);
/* executes the executions recorded so far (after applying pending buffer fills) and clears the list without ending the
 * frame: transient images stay valid. Lets a sharded caller interleave collectives (halo / all-gather / all-reduce over
 * NCCL) between groups of passes of one frame. render_frame == submit of what is left. */
PLAIN_EXPORT int PLAIN_FN(submit_recorded_passes)(plain_ctx* ctx);
PLAIN_EXPORT int PLAIN_FN(wait_for_gpu_idle)(plain_ctx* ctx);
PLAIN_EXPORT int PLAIN_FN(get_renderpass_timings)(plain_ctx* ctx, plain_pass_time* out, uint32_t capacity, uint32_t* out_count);
PLAIN_EXPORT int PLAIN_FN(set_timing_enabled)(plain_ctx* ctx, int enabled);

/* ---- meshes and graphic passes (RenderBackend.h:57-96): createMeshes, createGraphicPass, setGraphicPassExecution (after
 * new_frame, before prepare_for_drawcall_recording), drawMeshes (after it; one push-constant block per mesh, its size fixed by
 * the pass: 16 bytes for depthPrepass / triangle, 8 for sunShadow; worker_index is accepted and ignored - draws are ordered by
 * call). A graphic pass runs where its execution was set in the pass order, with the draws recorded for it this frame. ---- */
PLAIN_EXPORT int PLAIN_FN(create_meshes)(plain_ctx* ctx, const plain_mesh_binary* meshes, uint32_t n, plain_handle* out_handles);
PLAIN_EXPORT int PLAIN_FN(create_graphic_pass)(plain_ctx* ctx, const plain_graphic_pass_desc* desc, plain_handle* out);
PLAIN_EXPORT int PLAIN_FN(set_graphic_pass_execution)(plain_ctx* ctx, const plain_graphic_pass_execution* execution);
PLAIN_EXPORT int PLAIN_FN(draw_meshes)(plain_ctx* ctx, const plain_handle* meshes, uint32_t n, const void* push_constants, plain_handle pass, int32_t worker_index);

/* ---- additions (not in the reference): the raster passes that produce depth/motion/normal/shadow maps/G-buffer are
 * out of scope, so their outputs are uploaded; read-back exists for parity tests and the e2e bench leg. Synchronous. ---- */
PLAIN_EXPORT int PLAIN_FN(write_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, const void* data, size_t size);
PLAIN_EXPORT int PLAIN_FN(read_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, void* out, size_t size);
PLAIN_EXPORT int PLAIN_FN(read_storage_buffer)(plain_ctx* ctx, plain_handle buffer, void* out, size_t size);
/* async variants with caller-pinned host memory (e2e leg: upload of the frame inputs, read-back of the frame). In the CUDA
 * backend they run on dedicated upload / download streams so the copy engines overlap the passes: an upload waits for the
 * last submission that referenced the image, a submission waits for the uploads issued before it, a read-back sees the
 * passes submitted before it. The host buffer of a read-back is valid after wait_for_gpu_idle (or once an event recorded
 * on the pass stream after join_transfers has completed). */
PLAIN_EXPORT int PLAIN_FN(write_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, const void* pinned_data, size_t size);
PLAIN_EXPORT int PLAIN_FN(read_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, void* pinned_out, size_t size);
/* rows [row_begin, row_end) of a 2-D image level (images are row-major and tightly packed): a rank of a row-sharded frame
 * uploads / reads back only its band (+ halo). `data` points at the first transferred row. */
PLAIN_EXPORT int PLAIN_FN(write_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, uint32_t row_begin, uint32_t row_end, const void* pinned_data, size_t size);
PLAIN_EXPORT int PLAIN_FN(read_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, uint32_t row_begin, uint32_t row_end, void* pinned_out, size_t size);
/* device pointer of mip 0 (CUDA backend only; oracle returns the host pointer). Lets a caller that owns device memory
 * (e.g. a torch tensor holding a shard received over NCCL) fill an image without a host round trip. */
PLAIN_EXPORT int PLAIN_FN(get_image_device_pointer)(plain_ctx* ctx, plain_image_handle image, uint32_t mip_level, void** out_ptr, size_t* out_size);
PLAIN_EXPORT int PLAIN_FN(get_storage_buffer_device_pointer)(plain_ctx* ctx, plain_handle buffer, void** out_ptr, size_t* out_size);
/* number of kernels launched by the last render_frame (bench "gpu_launches") */
PLAIN_EXPORT int PLAIN_FN(get_last_frame_launch_count)(plain_ctx* ctx, uint32_t* out);
/* enable replay of an unchanged pass list through a captured CUDA graph (CUDA backend; no-op in the oracle) */
PLAIN_EXPORT int PLAIN_FN(set_graph_replay_enabled)(plain_ctx* ctx, int enabled);
/* Pass fusion (default on): a producer pass whose only consumer in the submission can compute its texels inline is not launched.
 * Today: indirectLightUpscale.comp folded into gbufferShading.comp (the two full-resolution GI images are then NOT written); the small
 * levels of the bloom chain as one launch; the froxel chain froxelVolumeMaterial -> froxelLightScattering -> volumeLightingReprojection ->
 * volumetricLightingIntegration as one launch over froxel columns (the material and scattering volumes are then NOT written). Switch
 * fusion off when those images are to be read back. Results are bit-identical either way. No-op in the oracle. */
PLAIN_EXPORT int PLAIN_FN(set_pass_fusion_enabled)(plain_ctx* ctx, int enabled);
/* schedule the passes of a submission onto several streams from the hazards between their declared resources (default on):
 * a pass waits only for the passes it conflicts with, like the barriers the reference derives (RenderBackend.cpp:632-767).
 * Off: strictly in submission order on one stream. No-op in the oracle. */
PLAIN_EXPORT int PLAIN_FN(set_concurrent_passes_enabled)(plain_ctx* ctx, int enabled);
/* makes the pass stream wait for every asynchronous upload and read-back issued so far (no host synchronisation): an event
 * recorded on the pass stream afterwards covers them. No-op in the oracle. */
PLAIN_EXPORT int PLAIN_FN(join_transfers)(plain_ctx* ctx);

/* ---- peer exchange over NVLink (row-sharded frames, one process per GPU; no counterpart in the reference) ----
 * The ranks map each other's images (CUDA IPC) once; an exchange is then a kernel that stores this rank's rows straight into
 * the peers' copies of the image (peer_push_rows) followed by a flag barrier in peer memory (peer_barrier): no host round
 * trip, no library collective on the frame path. Everything is enqueued on the pass stream.
 *   set-up (collective, the caller ships the 64-byte handles between the processes, e.g. with torch.distributed):
 *     peer_init -> peer_get_sync_handle -> [exchange] -> peer_open_sync for every other rank
 *     per exchanged image: peer_get_image_handle -> [exchange] -> peer_open_image for every other rank
 *   per exchange: peer_push_rows (rows of image levels to given peers) + peer_barrier, or peer_allreduce_sum_u32.
 * The oracle backend returns an error from all of them. */
#define PLAIN_IPC_HANDLE_BYTES 64
#define PLAIN_MAX_PEERS 16
typedef struct {
    plain_image_handle image;
    uint32_t mip_level;
    uint32_t row_begin, row_end; /* rows of the level (of every slice of a 3-D level), copied to the same rows of the peer's image */
    uint32_t peer;               /* destination rank */
} plain_peer_push;
PLAIN_EXPORT int PLAIN_FN(peer_init)(plain_ctx* ctx, uint32_t rank, uint32_t count);
PLAIN_EXPORT int PLAIN_FN(peer_get_sync_handle)(plain_ctx* ctx, void* out_handle);
PLAIN_EXPORT int PLAIN_FN(peer_open_sync)(plain_ctx* ctx, uint32_t peer, const void* handle);
PLAIN_EXPORT int PLAIN_FN(peer_get_image_handle)(plain_ctx* ctx, plain_image_handle image, void* out_handle);
PLAIN_EXPORT int PLAIN_FN(peer_open_image)(plain_ctx* ctx, plain_image_handle image, uint32_t peer, const void* handle);
/* 1 when every other rank's copy of the image is mapped */
PLAIN_EXPORT int PLAIN_FN(peer_image_ready)(plain_ctx* ctx, plain_image_handle image);
PLAIN_EXPORT int PLAIN_FN(peer_push_rows)(plain_ctx* ctx, uint32_t n, const plain_peer_push* pushes);
/* every rank signals every other rank and waits for all of them: pushes enqueued before it on any rank are visible to the
 * passes enqueued after it on every rank */
PLAIN_EXPORT int PLAIN_FN(peer_barrier)(plain_ctx* ctx);
/* Deferred exchange, for rows that only the NEXT frame reads (GI / froxel / TAA history): the push is enqueued on a separate stream
 * behind the passes submitted so far and overlaps the rest of the frame; peer_flush_deferred (once per frame, after the last pass)
 * closes the frame's deferred pushes with one barrier on a second flag set. The first submission after the next new_frame - and any
 * read-back or wait_for_gpu_idle - waits for it. Rule for the caller: at least one peer_barrier / peer_allreduce_sum_u32 of the same
 * frame precedes the first deferred push (it orders the peers' readers of the previous contents before the new rows arrive). */
PLAIN_EXPORT int PLAIN_FN(peer_push_rows_deferred)(plain_ctx* ctx, uint32_t n, const plain_peer_push* pushes);
PLAIN_EXPORT int PLAIN_FN(peer_flush_deferred)(plain_ctx* ctx);
/* element-wise sum over the ranks of `count` (<= 256) u32 at the start of a storage buffer; includes its own barrier */
PLAIN_EXPORT int PLAIN_FN(peer_allreduce_sum_u32)(plain_ctx* ctx, plain_handle storage_buffer, uint32_t count);
/* non-zero once a barrier gave up waiting for a peer (about 2 s): the frame is invalid, the caller must abort */
PLAIN_EXPORT int PLAIN_FN(peer_error)(plain_ctx* ctx, uint32_t* out_error);
/* the same word without blocking: returns what the previous poll fetched and enqueues the next read-back (call it once per frame; the
 * word is sticky, so a time-out is seen one frame later at the latest). peer_error (blocking) also clears the word once it reported it */
PLAIN_EXPORT int PLAIN_FN(peer_error_poll)(plain_ctx* ctx, uint32_t* out_error);
/* the stream all passes run on (cudaStream_t as void*), for callers that time with CUDA events */
PLAIN_EXPORT int PLAIN_FN(get_stream)(plain_ctx* ctx, void** out_stream);

/* ---- device self test (new; no reference counterpart) ----
 * Exhaustive on-device comparison of the instruction-lean sequences the kernels use with the functions of the numeric contract they
 * stand for, over every binary32 value of their domain: [0] rcpf_nz vs __frcp_rn (every value that is not zero or denormal),
 * [1] rcpf_normal vs __frcp_rn (2^-126 <= |x| < 2^126), [2] lean R11G11B10 channel encoder vs the contract's (all 2^32 values x both
 * mantissa widths), [3] lean decoder vs the contract's (all codes), [4] floor2i + int->float vs floor_ + f2i (|x| < 2^24),
 * [5] FMNMX vs the pinned min / max on non-negative-zero operands (2^26 random pairs). out_mismatches receives 8 counters (unused: 0).
 * The CPU oracle returns zeros without running anything. */
PLAIN_EXPORT int PLAIN_FN(device_selftest)(plain_ctx* ctx, uint64_t* out_mismatches);

#ifdef __cplusplus
}
#endif
#endif /* PLAIN_B200_H */
