"""ctypes bindings for the C-ABI in include/plain_b200.h and include/plain_frontend.h.

`Api(path, fn_prefix, fe_prefix)` binds one shared library. The product library is bound by
`plainrenderer_b200.load()` (prefixes ``plain_`` / ``plain_frontend_``); tests bind the CPU oracle with the
same class and the ``oracle_`` prefixes - the structures are identical by construction (one header).
"""
import ctypes as C
import numpy as np

u32, i32, f32 = C.c_uint32, C.c_int32, C.c_float

# plain_image_format (ImageDescription.h:16 order + RGBA32_UINT)
FORMATS = ["R8", "RG8", "RGBA8", "R16_SFLOAT", "RG16_SFLOAT", "RG32_SFLOAT", "RG16_SNORM", "RGBA16_SFLOAT", "RGBA16_SNORM",
           "RGBA32_SFLOAT", "R11G11B10_UFLOAT", "DEPTH16", "DEPTH32", "BC1", "BC3", "BC5", "BGRA8_UNORM", "RGBA32_UINT"]
FORMAT = {n: i for i, n in enumerate(FORMATS)}
BYTES_PER_TEXEL = {"R8": 1, "RG8": 2, "RGBA8": 4, "R16_SFLOAT": 2, "RG16_SFLOAT": 4, "RG32_SFLOAT": 8, "RG16_SNORM": 4, "RGBA16_SFLOAT": 8,
                   "RGBA16_SNORM": 8, "RGBA32_SFLOAT": 16, "R11G11B10_UFLOAT": 4, "DEPTH16": 2, "DEPTH32": 4, "BGRA8_UNORM": 4, "RGBA32_UINT": 16}
IMAGE_1D, IMAGE_2D, IMAGE_3D, IMAGE_CUBE = 0, 1, 2, 3
MIPS_ONE, MIPS_FULL_CHAIN, MIPS_MANUAL, MIPS_IN_DATA = 0, 1, 2, 3
USAGE_STORAGE, USAGE_SAMPLED, USAGE_ATTACHMENT = 1, 2, 4


class ImageDesc(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("depth", u32), ("type", u32), ("format", u32), ("usage_flags", u32),
                ("mip_count", u32), ("manual_mip_count", u32), ("auto_create_mips", u32)]


class ImageHandle(C.Structure):
    _fields_ = [("type", u32), ("index", u32)]


class SamplerDesc(C.Structure):
    _fields_ = [("interpolation", u32), ("wrapping", u32), ("use_anisotropy", u32), ("max_anisotropy", f32), ("border_color", u32), ("max_mip", u32)]


class SpecConst(C.Structure):
    _fields_ = [("location", u32), ("data", C.c_void_p), ("size", u32)]


class StorageBufferResource(C.Structure):
    _fields_ = [("buffer", u32), ("read_only", u32), ("binding", u32)]


class UniformBufferResource(C.Structure):
    _fields_ = [("buffer", u32), ("binding", u32)]


class ImageResource(C.Structure):
    _fields_ = [("image", ImageHandle), ("mip_level", u32), ("binding", u32)]


class SamplerResource(C.Structure):
    _fields_ = [("sampler", u32), ("binding", u32)]


class PassResources(C.Structure):
    _fields_ = [("samplers", C.POINTER(SamplerResource)), ("n_samplers", u32),
                ("storage_buffers", C.POINTER(StorageBufferResource)), ("n_storage_buffers", u32),
                ("uniform_buffers", C.POINTER(UniformBufferResource)), ("n_uniform_buffers", u32),
                ("sampled_images", C.POINTER(ImageResource)), ("n_sampled_images", u32),
                ("storage_images", C.POINTER(ImageResource)), ("n_storage_images", u32)]


class ComputePassExecution(C.Structure):
    _fields_ = [("pass_", u32), ("resources", PassResources), ("push_constants", C.c_void_p), ("push_constant_size", u32), ("dispatch_count", u32 * 3),
                ("row_begin", u32), ("row_end", u32), ("shard_phase", u32)]


class ShadowCascadeInfo(C.Structure):  # plain_shadow_cascade_info (sunShadowCascades.inc:7-11)
    _fields_ = [("splits", f32 * 4), ("lightMatrices", (f32 * 16) * 4), ("lightSpaceScale", (f32 * 2) * 4)]


class MeshBinary(C.Structure):  # plain_mesh_binary
    _fields_ = [("index_count", u32), ("vertex_count", u32), ("index_buffer", C.c_void_p), ("vertex_buffer", C.c_void_p)]


class Attachment(C.Structure):
    _fields_ = [("format", u32), ("load_op", u32)]


class GraphicPassDesc(C.Structure):  # plain_graphic_pass_desc
    _fields_ = [("vertex_shader", C.c_char_p), ("vertex_consts", C.POINTER(SpecConst)), ("n_vertex_consts", u32),
                ("fragment_shader", C.c_char_p), ("fragment_consts", C.POINTER(SpecConst)), ("n_fragment_consts", u32),
                ("attachments", C.POINTER(Attachment)), ("n_attachments", u32), ("cull_mode", u32), ("clamp_depth", u32),
                ("depth_function", u32), ("depth_write", u32), ("debug_name", C.c_char_p)]


class RenderTarget(C.Structure):
    _fields_ = [("image", ImageHandle), ("mip_level", u32)]


class GraphicPassExecution(C.Structure):
    _fields_ = [("pass_", u32), ("resources", PassResources), ("targets", C.POINTER(RenderTarget)), ("n_targets", u32), ("row_begin", u32), ("row_end", u32)]


CULL_NONE, CULL_FRONT, CULL_BACK = 0, 1, 2
DEPTH_GREATER_EQUAL, DEPTH_EQUAL = 5, 6
LOAD_OP_LOAD, LOAD_OP_CLEAR = 0, 1


def pack_vertices(positions, uvs=None, normals=None, tangents=None, bitangents=None):
    """28-byte vertices of MeshBinary (MeshProcessing.cpp:52-106): f32 x 3, f16 x 2, three A2R10G10B10_SNORM (CompressedTypes.cpp:24-46)."""
    positions = np.asarray(positions, np.float32).reshape(-1, 3)
    n = len(positions)

    def pack_snorm(v):
        v = np.zeros((n, 3), np.float32) if v is None else np.asarray(v, np.float32).reshape(n, 3)
        bits = (np.clip(v, -1, 1) * 0.5 + 0.5) * np.float32(511 + 510) + np.float32(-510)
        bits = bits.astype(np.int32) & 1023
        return (bits[:, 0].astype(np.uint32) << 20) | (bits[:, 1].astype(np.uint32) << 10) | bits[:, 2].astype(np.uint32)
    out = np.zeros((n, 28), np.uint8)
    out[:, 0:12] = positions.view(np.uint8).reshape(n, 12)
    uv = np.zeros((n, 2), np.float16) if uvs is None else np.asarray(uvs, np.float32).reshape(n, 2).astype(np.float16)
    out[:, 12:16] = uv.view(np.uint8).reshape(n, 4)
    for k, v in enumerate((normals, tangents, bitangents)):
        out[:, 16 + 4 * k:20 + 4 * k] = pack_snorm(v).astype("<u4").view(np.uint8).reshape(n, 4)
    return out


class PassTime(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("time_ms", f32)]


class GlobalShaderInfo(C.Structure):  # include/plain_frame_types.h plain_global_shader_info
    _fields_ = [("viewProjection", f32 * 16), ("viewProjectionPrevious", f32 * 16), ("sunDirection", f32 * 4), ("cameraPosition", f32 * 4),
                ("cameraPositionPrevious", f32 * 4), ("cameraRight", f32 * 4), ("cameraUp", f32 * 4), ("cameraForward", f32 * 4),
                ("cameraForwardPrevious", f32 * 4), ("noiseTextureIndices", i32 * 4), ("currentFrameCameraJitter", f32 * 2),
                ("previousFrameCameraJitter", f32 * 2), ("screenResolution", i32 * 2), ("cameraTanFovHalf", f32), ("cameraAspectRatio", f32),
                ("nearPlane", f32), ("farPlane", f32), ("sunStrength", f32), ("exposureOffset", f32), ("exposureAdaptionSpeedEvPerSec", f32),
                ("deltaTime", f32), ("time", f32), ("mipBias", f32), ("cameraCut", u32), ("frameIndex", u32), ("frameIndexMod2", u32),
                ("frameIndexMod3", u32), ("frameIndexMod4", u32)]


class FrontendSettings(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("diffuse_brdf", i32), ("direct_multiscatter", i32), ("indirect_lighting_tech", i32),
                ("use_geometry_aa", i32), ("sun_shadow_cascade_count", i32), ("half_res_trace", i32), ("strict_influence_radius_cutoff", i32),
                ("trace_influence_radius", f32), ("taa_enabled", i32), ("taa_use_clipping", i32), ("taa_use_motion_vector_dilation", i32),
                ("taa_history_sampling_tech", i32), ("taa_filter_use_tonemapping", i32), ("bloom_enabled", i32), ("bloom_strength", f32),
                ("bloom_radius", f32), ("sun_direction_deg", f32 * 2), ("camera_fov_deg", f32), ("camera_near", f32), ("camera_far", f32),
                ("noise_seed", u32), ("shard_rank", u32), ("shard_count", u32), ("taa_use_separate_supersampling", i32), ("taa_supersample_use_tonemapping", i32),
                ("sdf_debug_mode", i32), ("sdf_debug_show_tile_usage_with_hiz", i32), ("sdf_debug_use_influence_radius", i32), ("raster_inputs", i32)]


class FrameInputs(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("motion", C.c_void_p), ("normal", C.c_void_p), ("gbuffer", C.c_void_p), ("shadow_maps", C.c_void_p * 4), ("async_upload", i32), ("row_begin", u32), ("row_end", u32)]


EXCHANGE_NONE, EXCHANGE_ALLREDUCE_SUM_U32, EXCHANGE_ALLGATHER_ROWS, EXCHANGE_HALO_ROWS = 0, 1, 2, 3


class Exchange(C.Structure):  # include/plain_frontend.h plain_exchange
    _fields_ = [("kind", u32), ("n_images", u32), ("device_ptr", C.c_void_p * 4), ("row_pitch_bytes", u32 * 4), ("rows", u32 * 4), ("row_divisor", u32 * 4),
                ("halo_rows", u32), ("element_count", u32), ("name", C.c_char * 32), ("image", ImageHandle * 4), ("mip_level", u32 * 4), ("buffer", u32), ("depth", u32 * 4), ("deferred", u32)]


class CameraExtrinsic(C.Structure):
    _fields_ = [("position", f32 * 3), ("forward", f32 * 3), ("right", f32 * 3), ("up", f32 * 3)]


BACKEND_SYMBOLS = [
    "backend_create", "backend_destroy", "last_error", "recreate_swapchain", "create_image", "create_temporary_image", "resize_images",
    "get_image_description", "get_image_global_texture_array_index", "create_uniform_buffer", "create_storage_buffer", "create_sampler",
    "get_swapchain_input_image", "create_compute_pass", "update_compute_pass_shader_description", "set_global_descriptor_set_resources",
    "new_frame", "set_compute_pass_execution", "prepare_for_drawcall_recording", "set_uniform_buffer_data", "set_storage_buffer_data",
    "render_frame", "submit_recorded_passes", "wait_for_gpu_idle", "get_renderpass_timings", "set_timing_enabled", "write_image", "read_image", "read_storage_buffer",
    "create_meshes", "create_graphic_pass", "set_graphic_pass_execution", "draw_meshes",
    "write_image_async", "read_image_async", "write_image_rows_async", "read_image_rows_async", "get_image_device_pointer", "get_storage_buffer_device_pointer", "get_last_frame_launch_count",
    "set_graph_replay_enabled", "set_pass_fusion_enabled", "set_concurrent_passes_enabled", "join_transfers", "get_stream",
    "peer_init", "peer_get_sync_handle", "peer_open_sync", "peer_get_image_handle", "peer_open_image", "peer_image_ready", "peer_push_rows", "peer_barrier", "peer_push_rows_deferred", "peer_flush_deferred",
    "peer_allreduce_sum_u32", "peer_error", "peer_error_poll", "device_selftest"]
FRONTEND_SYMBOLS = [
    "default_settings", "create", "destroy", "last_error", "backend", "register_sdf_mesh", "set_mesh_geometry", "set_scene", "render_frame", "begin_frame", "run_segment", "set_peer_exchange", "shard_band",
    "read_output_rows", "read_output", "get_image",
    "get_storage_buffer", "get_global_shader_info", "get_resolve_weights", "set_exposure", "synthetic_scene_create", "synthetic_scene_destroy",
    "synthetic_scene_attach", "synthetic_scene_render_inputs",
    "host_hammersley2d", "host_direction_to_vector", "host_mip_count_from_resolution", "host_camera_matrices", "host_view_frustum", "host_aabb_intersects_frustum",
    "host_pad_sdf_bounding_box", "host_sdf_world_to_local", "host_orthogonal_frustum_fitted_to_camera", "get_drawcall_counts"]


class ApiError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Api:
    """One loaded library exposing the backend C-ABI (`b_*`) and the frontend C API (`f_*`)."""

    def __init__(self, path, fn_prefix="plain_", fe_prefix="plain_frontend_"):
        self.path = str(path)
        self.lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        self.fn_prefix, self.fe_prefix = fn_prefix, fe_prefix
        self.b, self.f = {}, {}
        for s in BACKEND_SYMBOLS:
            self.b[s] = getattr(self.lib, fn_prefix + s)  # AttributeError if the library does not export it
        for s in FRONTEND_SYMBOLS:
            fn = getattr(self.lib, fe_prefix + s, None)
            if fn is not None:
                self.f[s] = fn
        self.b["last_error"].restype = C.c_char_p
        self.b["backend_destroy"].restype = None
        if "last_error" in self.f:
            self.f["last_error"].restype = C.c_char_p
            self.f["backend"].restype = C.c_void_p
            self.f["destroy"].restype = None
            self.f["default_settings"].restype = None
            self.f["synthetic_scene_destroy"].restype = None
            self.f["shard_band"].restype = None


class Backend:
    """Thin object wrapper over a `plain_ctx*` (reference `RenderBackend` method names, snake_case)."""

    def __init__(self, api, device=0, width=64, height=64, ctx=None):
        self.api = api
        self.owned = ctx is None
        if ctx is None:
            p = C.c_void_p()
            if api.b["backend_create"](C.c_int(device), u32(width), u32(height), C.byref(p)):
                raise ApiError("backend_create failed")
            ctx = p
        self.ctx = C.c_void_p(ctx) if not isinstance(ctx, C.c_void_p) else ctx
        self._keep = []

    def close(self):
        if self.owned and self.ctx:
            self.api.b["backend_destroy"](self.ctx)
        self.ctx = None

    def _check(self, rc, what):
        if rc:
            raise ApiError("%s: %s" % (what, self.api.b["last_error"](self.ctx).decode()))

    def create_image(self, width, height, fmt, depth=1, type_=IMAGE_2D, usage=USAGE_STORAGE | USAGE_SAMPLED, mips=MIPS_ONE, manual_mips=1, data=None):
        d = ImageDesc(width, height, depth, type_, FORMAT[fmt], usage, mips, manual_mips, 0)
        h = ImageHandle()
        if data is not None:
            data = np.ascontiguousarray(data)
        self._check(self.api.b["create_image"](self.ctx, C.byref(d), _ptr(data), C.c_size_t(0 if data is None else data.nbytes), C.byref(h)), "create_image")
        return h

    def create_temporary_image(self, width, height, fmt, mips=MIPS_ONE, manual_mips=1):
        d = ImageDesc(width, height, 1, IMAGE_2D, FORMAT[fmt], USAGE_STORAGE | USAGE_SAMPLED, mips, manual_mips, 0)
        h = ImageHandle()
        self._check(self.api.b["create_temporary_image"](self.ctx, C.byref(d), C.byref(h)), "create_temporary_image")
        return h

    def image_description(self, h):
        d = ImageDesc()
        self._check(self.api.b["get_image_description"](self.ctx, h, C.byref(d)), "get_image_description")
        return d

    def mip_shape(self, h, mip=0):
        d = self.image_description(h)
        return max(d.width >> mip, 1), max(d.height >> mip, 1), max(max(d.depth, 1) >> mip, 1), FORMATS[d.format]

    def global_texture_index(self, h):
        i = u32()
        self._check(self.api.b["get_image_global_texture_array_index"](self.ctx, h, C.byref(i)), "global texture index")
        return i.value

    def create_storage_buffer(self, size, data=None):
        h = u32()
        if data is not None:
            data = np.ascontiguousarray(data)
        self._check(self.api.b["create_storage_buffer"](self.ctx, C.c_size_t(size), _ptr(data), C.byref(h)), "create_storage_buffer")
        return h.value

    def create_uniform_buffer(self, size, data=None):
        h = u32()
        if data is not None:
            data = np.ascontiguousarray(data)
        self._check(self.api.b["create_uniform_buffer"](self.ctx, C.c_size_t(size), _ptr(data), C.byref(h)), "create_uniform_buffer")
        return h.value

    def swapchain_image(self):
        h = ImageHandle()
        self._check(self.api.b["get_swapchain_input_image"](self.ctx, C.byref(h)), "get_swapchain_input_image")
        return h

    def create_compute_pass(self, shader, spec=None, name=None):
        """spec: {location: bytes-like or numpy scalar}"""
        spec = spec or {}
        keep = [np.frombuffer(bytes(v) if isinstance(v, (bytes, bytearray)) else np.asarray(v).tobytes(), np.uint8).copy() for v in spec.values()]
        arr = (SpecConst * max(len(spec), 1))()
        for i, (loc, k) in enumerate(zip(spec.keys(), keep)):
            arr[i] = SpecConst(loc, k.ctypes.data, k.nbytes)
        h = u32()
        self._check(self.api.b["create_compute_pass"](self.ctx, shader.encode(), arr, u32(len(spec)), (name or shader).encode(), C.byref(h)), "create_compute_pass(%s)" % shader)
        return h.value

    def set_global_uniform_buffer(self, buffer):
        ub = (UniformBufferResource * 1)(UniformBufferResource(buffer, 0))
        r = PassResources()
        r.uniform_buffers, r.n_uniform_buffers = ub, 1
        self._check(self.api.b["set_global_descriptor_set_resources"](self.ctx, C.byref(r)), "set_global_descriptor_set_resources")

    def new_frame(self):
        self._check(self.api.b["new_frame"](self.ctx), "new_frame")

    def set_compute_pass_execution(self, pass_, dispatch, sampled=(), storage=(), storage_buffers=(), uniform_buffers=(), push=None, rows=None, shard_phase=0):
        """sampled/storage: [(handle, mip, binding)]; storage_buffers: [(handle, read_only, binding)]; uniform_buffers: [(handle, binding)]"""
        e = ComputePassExecution()
        e.pass_ = pass_
        si = (ImageResource * max(len(sampled), 1))(*[ImageResource(h, m, b) for h, m, b in sampled])
        st = (ImageResource * max(len(storage), 1))(*[ImageResource(h, m, b) for h, m, b in storage])
        sb = (StorageBufferResource * max(len(storage_buffers), 1))(*[StorageBufferResource(h, int(ro), b) for h, ro, b in storage_buffers])
        ub = (UniformBufferResource * max(len(uniform_buffers), 1))(*[UniformBufferResource(h, b) for h, b in uniform_buffers])
        e.resources.sampled_images, e.resources.n_sampled_images = si, len(sampled)
        e.resources.storage_images, e.resources.n_storage_images = st, len(storage)
        e.resources.storage_buffers, e.resources.n_storage_buffers = sb, len(storage_buffers)
        e.resources.uniform_buffers, e.resources.n_uniform_buffers = ub, len(uniform_buffers)
        pc = None
        if push is not None:
            pc = np.frombuffer(bytes(push), np.uint8).copy()
            e.push_constants, e.push_constant_size = pc.ctypes.data, pc.nbytes
        for i in range(3):
            e.dispatch_count[i] = dispatch[i] if i < len(dispatch) else 1
        if rows is not None:
            e.row_begin, e.row_end = rows
        e.shard_phase = shard_phase
        self._check(self.api.b["set_compute_pass_execution"](self.ctx, C.byref(e)), "set_compute_pass_execution")

    def create_meshes(self, meshes):
        """meshes: [(indices uint array, 28-byte vertex array from pack_vertices)]; indices travel as u16 when there are fewer than 65535"""
        arr = (MeshBinary * max(len(meshes), 1))()
        keep = []
        for i, (idx, vtx) in enumerate(meshes):
            idx = np.ascontiguousarray(np.asarray(idx).ravel().astype(np.uint16 if len(np.asarray(idx).ravel()) < 65535 else np.uint32))
            vtx = np.ascontiguousarray(np.asarray(vtx, np.uint8).reshape(-1, 28))
            keep += [idx, vtx]
            arr[i] = MeshBinary(len(idx), len(vtx), idx.ctypes.data, vtx.ctypes.data)
        out = (u32 * max(len(meshes), 1))()
        self._check(self.api.b["create_meshes"](self.ctx, arr, u32(len(meshes)), out), "create_meshes")
        return [out[i] for i in range(len(meshes))]

    def create_graphic_pass(self, vertex_shader, fragment_shader, attachments, cull_mode, clamp_depth=False, depth_function=DEPTH_GREATER_EQUAL, vertex_spec=None, name=None):
        """attachments: [(format name, load op)]"""
        spec = vertex_spec or {}
        keep = [np.frombuffer(np.asarray(v).tobytes(), np.uint8).copy() for v in spec.values()]
        sc = (SpecConst * max(len(spec), 1))()
        for i, (loc, k) in enumerate(zip(spec.keys(), keep)):
            sc[i] = SpecConst(loc, k.ctypes.data, k.nbytes)
        at = (Attachment * max(len(attachments), 1))(*[Attachment(FORMAT[f], op) for f, op in attachments])
        d = GraphicPassDesc(vertex_shader.encode(), sc, len(spec), fragment_shader.encode(), None, 0, at, len(attachments), cull_mode, int(clamp_depth), depth_function, 1,
                            (name or vertex_shader).encode())
        h = u32()
        self._check(self.api.b["create_graphic_pass"](self.ctx, C.byref(d), C.byref(h)), "create_graphic_pass(%s + %s)" % (vertex_shader, fragment_shader))
        return h.value

    def set_graphic_pass_execution(self, pass_, targets, storage_buffers=(), sampled=(), rows=None):
        """targets: [(image handle, mip)] in attachment order; rows: (begin, end) of the targets to render (row sharding)"""
        e = GraphicPassExecution()
        e.pass_ = pass_
        if rows is not None:
            e.row_begin, e.row_end = rows
        sb = (StorageBufferResource * max(len(storage_buffers), 1))(*[StorageBufferResource(h, int(ro), b) for h, ro, b in storage_buffers])
        si = (ImageResource * max(len(sampled), 1))(*[ImageResource(h, m, b) for h, m, b in sampled])
        tg = (RenderTarget * max(len(targets), 1))(*[RenderTarget(h, m) for h, m in targets])
        e.resources.storage_buffers, e.resources.n_storage_buffers = sb, len(storage_buffers)
        e.resources.sampled_images, e.resources.n_sampled_images = si, len(sampled)
        e.targets, e.n_targets = tg, len(targets)
        self._check(self.api.b["set_graphic_pass_execution"](self.ctx, C.byref(e)), "set_graphic_pass_execution")

    def draw_meshes(self, meshes, push_constants, pass_):
        """push_constants: uint32 array, one block per mesh (4 words for depthPrepass / triangle, 2 for sunShadow)"""
        m = (u32 * max(len(meshes), 1))(*meshes)
        pc = np.ascontiguousarray(np.asarray(push_constants, np.uint32))
        self._check(self.api.b["draw_meshes"](self.ctx, m, u32(len(meshes)), _ptr(pc), u32(pass_), i32(0)), "draw_meshes")

    def set_uniform_buffer_data(self, h, data):
        data = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        self._check(self.api.b["set_uniform_buffer_data"](self.ctx, u32(h), _ptr(data), C.c_size_t(data.nbytes)), "set_uniform_buffer_data")

    def set_storage_buffer_data(self, h, data):
        data = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        self._check(self.api.b["set_storage_buffer_data"](self.ctx, u32(h), _ptr(data), C.c_size_t(data.nbytes)), "set_storage_buffer_data")

    def render_frame(self):
        self._check(self.api.b["prepare_for_drawcall_recording"](self.ctx), "prepare_for_drawcall_recording")
        self._check(self.api.b["render_frame"](self.ctx, C.c_int(1)), "render_frame")
        self._check(self.api.b["wait_for_gpu_idle"](self.ctx), "wait_for_gpu_idle")

    def write_image(self, h, data, mip=0):
        data = np.ascontiguousarray(data)
        self._check(self.api.b["write_image"](self.ctx, h, u32(mip), _ptr(data), C.c_size_t(data.nbytes)), "write_image")

    def read_image(self, h, mip=0, dtype=np.uint8):
        w, hh, d, fmt = self.mip_shape(h, mip)
        out = np.empty(w * hh * d * BYTES_PER_TEXEL[fmt], np.uint8)
        self._check(self.api.b["read_image"](self.ctx, h, u32(mip), _ptr(out), C.c_size_t(out.nbytes)), "read_image")
        return out.view(dtype)

    def read_storage_buffer(self, h, size, dtype=np.uint8):
        out = np.empty(size, np.uint8)
        self._check(self.api.b["read_storage_buffer"](self.ctx, u32(h), _ptr(out), C.c_size_t(size)), "read_storage_buffer")
        return out.view(dtype)

    def set_timing_enabled(self, on):
        self._check(self.api.b["set_timing_enabled"](self.ctx, C.c_int(int(on))), "set_timing_enabled")

    def pass_timings(self):
        arr = (PassTime * 128)()
        n = u32()
        self._check(self.api.b["get_renderpass_timings"](self.ctx, arr, u32(128), C.byref(n)), "get_renderpass_timings")
        return [(arr[i].name.decode(), arr[i].time_ms) for i in range(min(n.value, 128))]

    def last_frame_launch_count(self):
        n = u32()
        self._check(self.api.b["get_last_frame_launch_count"](self.ctx, C.byref(n)), "get_last_frame_launch_count")
        return n.value

    def set_graph_replay_enabled(self, on):
        self._check(self.api.b["set_graph_replay_enabled"](self.ctx, C.c_int(int(on))), "set_graph_replay_enabled")


def default_settings(api, width, height, **overrides):
    s = FrontendSettings()
    api.f["default_settings"](C.byref(s), u32(width), u32(height))
    for k, v in overrides.items():
        if k == "sun_direction_deg":
            s.sun_direction_deg[0], s.sun_direction_deg[1] = v
        else:
            setattr(s, k, v)
    return s


def camera(position, forward, right, up):
    c = CameraExtrinsic()
    for i in range(3):
        c.position[i], c.forward[i], c.right[i], c.up[i] = position[i], forward[i], right[i], up[i]
    return c


class Frontend:
    """Frame driver (RenderFrontend mirror) of one library."""

    def __init__(self, api, settings, device=0):
        self.api, self.settings = api, settings
        p = C.c_void_p()
        if api.f["create"](C.c_int(device), C.byref(settings), C.byref(p)):
            raise ApiError("frontend create failed (see stderr)")
        self.fe = p
        self.backend = Backend(api, ctx=api.f["backend"](self.fe))
        self.width, self.height = settings.width, settings.height

    def close(self):
        if self.fe:
            self.api.f["destroy"](self.fe)
        self.fe = None

    def _check(self, rc, what):
        if rc:
            raise ApiError("%s: %s" % (what, self.api.f["last_error"](self.fe).decode()))

    def _inputs(self, depth, motion, normal, gbuffer, shadow_maps, async_upload, rows):
        fi = FrameInputs()
        fi.depth, fi.motion, fi.normal, fi.gbuffer = _ptr(depth), _ptr(motion), _ptr(normal), _ptr(gbuffer)
        for i in range(4):
            fi.shadow_maps[i] = _ptr(shadow_maps[i]) if shadow_maps is not None and i < len(shadow_maps) and shadow_maps[i] is not None else None
        fi.async_upload = int(async_upload)
        if rows is not None:
            fi.row_begin, fi.row_end = rows
        return fi

    def begin_frame(self, cam, time, delta_time, depth=None, motion=None, normal=None, gbuffer=None, shadow_maps=None, async_upload=False, rows=None):
        """Row-sharded frames: records the frame; then call run_segment() until it returns None, performing each returned Exchange."""
        fi = self._inputs(depth, motion, normal, gbuffer, shadow_maps, async_upload, rows)
        self._check(self.api.f["begin_frame"](self.fe, C.byref(cam), f32(time), f32(delta_time), C.byref(fi)), "begin_frame")

    def run_segment(self):
        x = Exchange()
        rc = self.api.f["run_segment"](self.fe, C.byref(x))
        if rc < 0:
            self._check(1, "run_segment")
        return x if rc == 1 else None

    def read_output_rows(self, out, rows, async_pinned=False):
        self._check(self.api.f["read_output_rows"](self.fe, _ptr(out), u32(rows[0]), u32(rows[1]), i32(int(async_pinned))), "read_output_rows")
        return out

    def register_sdf_mesh(self, texels, bb_min, bb_max, mean_albedo):
        """texels: uint16 half floats (d, h, w); returns the frontend mesh number"""
        texels = np.ascontiguousarray(texels, np.uint16)
        d, h, w = texels.shape
        out = u32()
        self._check(self.api.f["register_sdf_mesh"](self.fe, _ptr(texels), u32(w), u32(h), u32(d), (f32 * 3)(*bb_min), (f32 * 3)(*bb_max), (f32 * 3)(*mean_albedo), C.byref(out)), "register_sdf_mesh")
        return out.value

    def set_mesh_geometry(self, mesh, indices, vertices, textures=(None, None, None)):
        """indices: uint array; vertices: (n, 28) uint8 (pack_vertices / assets.Mesh.vertices); textures: bindless slots of RGBA8 albedo / normal / specular images or None"""
        idx = np.ascontiguousarray(np.asarray(indices).ravel().astype(np.uint16 if np.asarray(indices).size < 65535 else np.uint32))
        vtx = np.ascontiguousarray(np.asarray(vertices, np.uint8).reshape(-1, 28))
        mb = MeshBinary(len(idx), len(vtx), idx.ctypes.data, vtx.ctypes.data)
        t = [0xFFFFFFFF if x is None else int(x) for x in textures]
        self._check(self.api.f["set_mesh_geometry"](self.fe, u32(mesh), C.byref(mb), u32(t[0]), u32(t[1]), u32(t[2])), "set_mesh_geometry")

    def set_scene(self, objects):
        """objects: [(mesh number, model matrix float32[16] column-major, world bb min, world bb max)]"""
        n = len(objects)
        meshes = np.array([o[0] for o in objects], np.uint32)
        mats = np.ascontiguousarray(np.array([np.asarray(o[1], np.float32).ravel() for o in objects], np.float32).reshape(n, 16))
        bmin = np.ascontiguousarray(np.array([o[2] for o in objects], np.float32).reshape(n, 3))
        bmax = np.ascontiguousarray(np.array([o[3] for o in objects], np.float32).reshape(n, 3))
        self._check(self.api.f["set_scene"](self.fe, u32(n), _ptr(meshes), _ptr(mats), _ptr(bmin), _ptr(bmax)), "set_scene")

    def render_frame(self, cam, time, delta_time, depth=None, motion=None, normal=None, gbuffer=None, shadow_maps=None, async_upload=False):
        fi = FrameInputs()
        fi.depth, fi.motion, fi.normal, fi.gbuffer = _ptr(depth), _ptr(motion), _ptr(normal), _ptr(gbuffer)
        for i in range(4):
            fi.shadow_maps[i] = _ptr(shadow_maps[i]) if shadow_maps is not None and i < len(shadow_maps) and shadow_maps[i] is not None else None
        fi.async_upload = int(async_upload)
        self._check(self.api.f["render_frame"](self.fe, C.byref(cam), f32(time), f32(delta_time), C.byref(fi)), "render_frame")

    def drawcall_counts(self):
        """(main pass / prepass draws, draws per shadow cascade) of the last frame, after the host-side culling"""
        out = (u32 * 2)()
        self._check(self.api.f["get_drawcall_counts"](self.fe, out), "get_drawcall_counts")
        return out[0], out[1]

    def read_output(self, out=None, async_pinned=False):
        if out is None:
            out = np.empty(self.width * self.height * 4, np.uint8)
        self._check(self.api.f["read_output"](self.fe, _ptr(out), C.c_size_t(out.nbytes), i32(int(async_pinned))), "read_output")
        return out

    def image(self, name):
        h = ImageHandle()
        self._check(self.api.f["get_image"](self.fe, name.encode(), C.byref(h)), "get_image(%s)" % name)
        return h

    def storage_buffer(self, name):
        h = u32()
        self._check(self.api.f["get_storage_buffer"](self.fe, name.encode(), C.byref(h)), "get_storage_buffer(%s)" % name)
        return h.value

    def global_shader_info(self):
        g = GlobalShaderInfo()
        self.api.f["get_global_shader_info"](self.fe, C.byref(g))
        return g

    def resolve_weights(self):
        w = (f32 * 9)()
        self.api.f["get_resolve_weights"](self.fe, w)
        return np.array(list(w), np.float32)

    def set_exposure(self, e):
        self._check(self.api.f["set_exposure"](self.fe, f32(e)), "set_exposure")


class SyntheticScene:
    def __init__(self, api, seed=0x504C4149, n_instances=100):
        self.api = api
        p = C.c_void_p()
        if api.f["synthetic_scene_create"](u32(seed), u32(n_instances), C.byref(p)):
            raise ApiError("synthetic_scene_create failed")
        self.s = p

    def close(self):
        if self.s:
            self.api.f["synthetic_scene_destroy"](self.s)
        self.s = None

    def attach(self, frontend):
        if self.api.f["synthetic_scene_attach"](self.s, frontend.fe):
            raise ApiError("synthetic_scene_attach: " + self.api.f["last_error"](frontend.fe).decode())

    def render_inputs(self, settings, cam, frame_index, prev_cam=None, shadows=True, threads=0, out=None):
        """Returns dict(depth, motion, normal, gbuffer, shadow_maps[list]) of numpy arrays (reused from `out` if given)."""
        w, h = settings.width, settings.height
        o = out or {}
        o.setdefault("depth", np.empty(w * h, np.float32))
        o.setdefault("motion", np.empty(w * h * 2, np.int16))
        o.setdefault("normal", np.empty(w * h * 4, np.uint8))
        o.setdefault("gbuffer", np.empty(w * h * 4, np.uint32))
        n_casc = settings.sun_shadow_cascade_count
        if shadows:
            o.setdefault("shadow_maps", [np.zeros(2048 * 2048, np.uint16) for _ in range(n_casc)])
        sm = (C.c_void_p * 4)()
        for i in range(4):
            sm[i] = _ptr(o["shadow_maps"][i]) if shadows and i < n_casc else None
        rc = self.api.f["synthetic_scene_render_inputs"](self.s, C.byref(settings), C.byref(cam), C.byref(prev_cam) if prev_cam is not None else None,
                                                         u32(frame_index), _ptr(o["depth"]), _ptr(o["motion"]), _ptr(o["normal"]), _ptr(o["gbuffer"]), sm, i32(threads))
        if rc:
            raise ApiError("synthetic_scene_render_inputs failed")
        return o
