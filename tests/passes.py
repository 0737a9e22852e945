"""Single-pass drivers over the C-ABI (used with both the CUDA product and the CPU oracle): each builds the resources one
reference shader binds, records one execution with the reference's binding numbers and returns the outputs."""
import ctypes as C

import numpy as np


class PassRig:
    def __init__(self, ffi, api, w=64, h=64, time=0.5, screen=None):
        self.ffi, self.api = ffi, api
        self.be = ffi.Backend(api, width=w, height=h)
        g = ffi.GlobalShaderInfo()
        g.time = time
        g.screenResolution[0], g.screenResolution[1] = screen or (w, h)
        g.nearPlane, g.farPlane = 0.1, 300.0
        g.cameraTanFovHalf, g.cameraAspectRatio = 0.3153, (screen or (w, h))[0] / (screen or (w, h))[1]
        g.cameraForward[2] = -1.0
        g.cameraUp[1] = -1.0
        g.cameraRight[0] = 1.0
        g.sunDirection[1] = -1.0
        g.deltaTime = 1 / 60.0
        g.exposureAdaptionSpeedEvPerSec, g.exposureOffset, g.sunStrength = 2.0, 1.0, 128000.0
        self.g = g
        self.gbuf = self.be.create_uniform_buffer(C.sizeof(g), np.frombuffer(bytes(g), np.uint8))
        self.be.set_global_uniform_buffer(self.gbuf)

    def close(self):
        self.be.close()

    def run(self):
        self.be.render_frame()


def tonemap(ffi, api, packed, time=0.37):
    h, w = packed.shape
    rig = PassRig(ffi, api, w, h, time=time)
    be = rig.be
    src = be.create_image(w, h, "R11G11B10_UFLOAT", data=packed.astype(np.uint32))
    p = be.create_compute_pass("tonemapping.comp")
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(src, 0, 1)], storage=[(be.swapchain_image(), 0, 0)])
    rig.run()
    out = be.read_image(be.swapchain_image()).reshape(h, w, 4).copy()
    rig.close()
    return out


def histogram(ffi, api, packed, exposure, n_bins=128, lum_min=0.001, lum_max=200000.0):
    h, w = packed.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    tx, ty = (w + 31) // 32, (h + 31) // 32
    src = be.create_image(w, h, "R11G11B10_UFLOAT", data=packed.astype(np.uint32))
    per_tile = be.create_storage_buffer(tx * ty * n_bins * 4)
    hist = be.create_storage_buffer(n_bins * 4, np.full(n_bins, 77, np.uint32))
    light = be.create_storage_buffer(20, np.array([1, 1, 1, exposure, 1], np.float32))
    spec = {0: np.uint32(n_bins), 1: np.float32(lum_min), 2: np.float32(lum_max), 3: np.int32(tx * ty)}
    p0 = be.create_compute_pass("histogramPerTile.comp", spec)
    p1 = be.create_compute_pass("histogramReset.comp", {0: np.uint32(n_bins)})
    p2 = be.create_compute_pass("histogramCombineTiles.comp", {0: np.uint32(n_bins), 1: np.int32(tx * ty)})
    be.new_frame()
    be.set_compute_pass_execution(p0, (tx, ty, 1), sampled=[(src, 0, 2)], storage_buffers=[(per_tile, False, 0), (light, True, 3)])
    be.set_compute_pass_execution(p1, ((n_bins + 63) // 64, 1, 1), storage_buffers=[(hist, False, 1)])
    be.set_compute_pass_execution(p2, (tx * ty, (n_bins + 63) // 64, 1), storage_buffers=[(per_tile, False, 0), (hist, False, 1)])
    rig.run()
    out = (be.read_storage_buffer(per_tile, tx * ty * n_bins * 4, np.uint32).reshape(ty, tx, n_bins).copy(), be.read_storage_buffer(hist, n_bins * 4, np.uint32).copy())
    rig.close()
    return out


def mip_count(w, h):
    return 1 + int(np.floor(np.log2(max(w, h))))


def hiz(ffi, api, depth):
    """depthHiZPyramid.comp with the bindings of RenderFrontend::computeDepthPyramid (RenderFrontend.cpp:804-838)."""
    h, w = depth.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    pw, ph = w // 2, h // 2
    n = mip_count(pw, ph)
    src = be.create_image(w, h, "DEPTH32", data=depth.astype(np.float32))
    pyr = be.create_image(pw, ph, "RG32_SFLOAT", mips=ffi.MIPS_FULL_CHAIN)
    sync = be.create_storage_buffer(4, np.zeros(1, np.uint32))
    p = be.create_compute_pass("depthHiZPyramid.comp", {0: np.uint32(n), 1: np.uint32(w), 2: np.uint32(h), 3: np.uint32(1)})
    bindings = max(11, n)  # 11 in the reference; a 12th level (binding 11) is the extension for 7680x4320
    unused = bindings - n
    storage = [(pyr, max(i - unused, 0), i) for i in range(bindings)]
    be.new_frame()
    be.set_compute_pass_execution(p, ((pw + 31) // 32, (ph + 31) // 32, 1), sampled=[(src, 0, 13), (pyr, 0, 15)], storage=storage, storage_buffers=[(sync, False, 16)])
    rig.run()
    out = [be.read_image(pyr, m, np.float32).reshape(max(ph >> m, 1), max(pw >> m, 1), 2).copy() for m in range(n)]
    rig.close()
    return out


def depth_downscale(ffi, api, depth):
    h, w = depth.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    src = be.create_image(w, h, "DEPTH32", data=depth.astype(np.float32))
    dst = be.create_image(w // 2, h // 2, "R16_SFLOAT")
    p = be.create_compute_pass("depthDownscale.comp")
    be.new_frame()
    be.set_compute_pass_execution(p, ((w // 2 + 7) // 8, (h // 2 + 7) // 8, 1), sampled=[(src, 0, 1)], storage=[(dst, 0, 0)])
    rig.run()
    out = be.read_image(dst, 0, np.float16).reshape(h // 2, w // 2).copy()
    rig.close()
    return out


def bloom(ffi, api, packed, strength=0.05, radius=1.5, mips=6):
    """Bloom::computeBloom (Bloom.cpp:56-144): 5 down, 5 up, apply. Returns (downscale mips, upscale mips, result)."""
    h, w = packed.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    target = be.create_image(w, h, "R11G11B10_UFLOAT", data=packed.astype(np.uint32))
    down = be.create_temporary_image(w, h, "R11G11B10_UFLOAT", mips=ffi.MIPS_MANUAL, manual_mips=mips)
    up = be.create_temporary_image(w, h, "R11G11B10_UFLOAT", mips=ffi.MIPS_MANUAL, manual_mips=mips)
    pd = [be.create_compute_pass("bloomDownsample.comp") for _ in range(mips - 1)]
    pu = [be.create_compute_pass("bloomUpsample.comp", {0: np.uint32(1 if i == 0 else 0)}) for i in range(mips - 1)]
    pa = be.create_compute_pass("applyBloom.comp")
    res = lambda m, v: max(v // (1 << m), 1)
    be.new_frame()
    for i in range(mips - 1):
        be.set_compute_pass_execution(pd[i], ((res(i + 1, w) + 7) // 8, (res(i + 1, h) + 7) // 8, 1), sampled=[(target if i == 0 else down, i, 1)], storage=[(down, i + 1, 0)])
    for i in range(mips - 1):
        t = mips - 2 - i
        be.set_compute_pass_execution(pu[i], ((res(t, w) + 7) // 8, (res(t, h) + 7) // 8, 1), sampled=[(up, t + 1, 1), (down, t + 1, 2)], storage=[(up, t, 0)],
                                      push=np.float32(radius).tobytes())
    be.set_compute_pass_execution(pa, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(up, 0, 1)], storage=[(target, 0, 0)], push=np.float32(strength).tobytes())
    rig.run()
    out = ([be.read_image(down, m, np.uint32).copy() for m in range(1, mips)], [be.read_image(up, m, np.uint32).copy() for m in range(mips - 1)],
           be.read_image(target, 0, np.uint32).reshape(h, w).copy())
    rig.close()
    return out


def color_to_luminance(ffi, api, packed):
    """colorToLuminance.comp with the bindings of TAA::computeTemporalSuperSampling (TAA.cpp:94-107): R11G11B10 -> R8."""
    h, w = packed.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    src = be.create_image(w, h, "R11G11B10_UFLOAT", data=packed.astype(np.uint32))
    dst = be.create_image(w, h, "R8")
    p = be.create_compute_pass("colorToLuminance.comp")
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(src, 0, 0)], storage=[(dst, 0, 1)])
    rig.run()
    out = be.read_image(dst, 0, np.uint8).reshape(h, w).copy()
    rig.close()
    return out


def temporal_supersampling(ffi, api, current, last, motion, depth_current, depth_last, lum_current, lum_last, use_tonemap=True):
    """temporalSupersampling.comp with the bindings of TAA.cpp:109-136. current/last: packed R11G11B10 (h, w) uint32; motion: (h, w, 2)
    int16 SNORM; depths: float32 D32F; luminance: uint8 R8. Returns the packed target."""
    h, w = current.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    img = lambda fmt, data: be.create_image(w, h, fmt, data=np.ascontiguousarray(data))
    cur, las = img("R11G11B10_UFLOAT", current.astype(np.uint32)), img("R11G11B10_UFLOAT", last.astype(np.uint32))
    vel = img("RG16_SNORM", motion.astype(np.int16))
    dc, dl = img("DEPTH32", depth_current.astype(np.float32)), img("DEPTH32", depth_last.astype(np.float32))
    lc, ll = img("R8", lum_current.astype(np.uint8)), img("R8", lum_last.astype(np.uint8))
    target = be.create_image(w, h, "R11G11B10_UFLOAT")
    p = be.create_compute_pass("temporalSupersampling.comp", {0: np.uint32(1 if use_tonemap else 0)})
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(cur, 0, 1), (las, 0, 2), (vel, 0, 4), (dc, 0, 5), (dl, 0, 6), (lc, 0, 7), (ll, 0, 8)], storage=[(target, 0, 3)])
    rig.run()
    out = be.read_image(target, 0, np.uint32).reshape(h, w).copy()
    rig.close()
    return out


# ---------------- the rasterisation passes (graphic passes of the C-ABI) ----------------
def _raster_rig(ffi, api, w, h, jitter=((0.0, 0.0), (0.0, 0.0))):
    rig = PassRig(ffi, api, w, h)
    rig.g.currentFrameCameraJitter[0], rig.g.currentFrameCameraJitter[1] = jitter[0]
    rig.g.previousFrameCameraJitter[0], rig.g.previousFrameCameraJitter[1] = jitter[1]
    rig.be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(rig.g), np.uint8))
    return rig


def raster_prepass(ffi, api, w, h, meshes, draws, matrices, jitter=((0.0, 0.0), (0.0, 0.0)), textures=None, gbuffer=False):
    """depthPrepass.vert/.frag (RenderFrontend.cpp:792-802, 1716-1735) and, with gbuffer=True, the G-buffer fill of the main pass.
    meshes: [(indices, vertices)]; draws: [(mesh number, transform index)] or [(mesh, transform, albedo, normal, specular texture numbers)];
    matrices: float32 (n, 3, 16) MainPassMatrices {model, mvp, mvpPrevious} column-major; textures: [(w, h, RGBA8 bytes)].
    Returns depth (h, w) f32, motion (h, w, 2) i16, normal (h, w, 4) u8 [, gbuffer (h, w, 4) u32]."""
    rig = _raster_rig(ffi, api, w, h, jitter)
    be = rig.be
    handles = be.create_meshes(meshes)
    tex = [be.create_image(tw, th, "RGBA8", data=np.asarray(d, np.uint8)) for tw, th, d in (textures or [(1, 1, [255, 255, 255, 255])])]
    tex_index = [be.global_texture_index(t) for t in tex]
    depth, motion, normal = be.create_image(w, h, "DEPTH32"), be.create_image(w, h, "RG16_SNORM"), be.create_image(w, h, "RGBA8")
    transforms = be.create_storage_buffer(max(len(matrices), 1) * 192, np.asarray(matrices, np.float32))
    pre = be.create_graphic_pass("depthPrepass.vert", "depthPrepass.frag", [("RG16_SNORM", ffi.LOAD_OP_CLEAR), ("RGBA8", ffi.LOAD_OP_CLEAR), ("DEPTH32", ffi.LOAD_OP_CLEAR)], ffi.CULL_BACK, name="Depth prepass")
    push = np.array([[tex_index[d[2]] if len(d) > 2 else tex_index[0], tex_index[d[3]] if len(d) > 2 else tex_index[0], tex_index[d[4]] if len(d) > 2 else tex_index[0], d[1]] for d in draws], np.uint32)
    be.new_frame()
    be.set_graphic_pass_execution(pre, [(motion, 0), (normal, 0), (depth, 0)], storage_buffers=[(transforms, True, 0)])
    if gbuffer:
        gb = be.create_image(w, h, "RGBA32_UINT")
        fill = be.create_graphic_pass("triangle.vert", "gbufferFill.frag", [("RGBA32_UINT", ffi.LOAD_OP_CLEAR), ("DEPTH32", ffi.LOAD_OP_LOAD)], ffi.CULL_BACK, depth_function=ffi.DEPTH_EQUAL, name="G-buffer fill")
        be.set_graphic_pass_execution(fill, [(gb, 0), (depth, 0)], storage_buffers=[(transforms, True, 17)])
    be._check(api.b["prepare_for_drawcall_recording"](be.ctx), "prepare_for_drawcall_recording")
    be.draw_meshes([handles[d[0]] for d in draws], push, pre)
    if gbuffer:
        be.draw_meshes([handles[d[0]] for d in draws], push, fill)
    rig.run()
    out = [be.read_image(depth, 0, np.float32).reshape(h, w).copy(), be.read_image(motion, 0, np.int16).reshape(h, w, 2).copy(), be.read_image(normal).reshape(h, w, 4).copy()]
    if gbuffer:
        out.append(be.read_image(gb, 0, np.uint32).reshape(h, w, 4).copy())
    rig.close()
    return out


def raster_shadow(ffi, api, size, meshes, draws, model_matrices, light_matrices, cascade=0, albedo=None):
    """sunShadow.vert/.frag (RenderFrontend.cpp:760-775, 1563-1590): one D16 cascade. draws: [(mesh number, transform index)];
    model_matrices float32 (n, 16); light_matrices float32 (4, 16)."""
    rig = _raster_rig(ffi, api, size, size)
    be = rig.be
    handles = be.create_meshes(meshes)
    shadow_map = be.create_image(size, size, "DEPTH16")
    cascades = be.create_storage_buffer(304, _shadow_cascade_info(ffi, light_matrices))
    transforms = be.create_storage_buffer(max(len(model_matrices), 1) * 64, np.asarray(model_matrices, np.float32))
    p = be.create_graphic_pass("sunShadow.vert", "sunShadow.frag", [("DEPTH16", ffi.LOAD_OP_CLEAR)], ffi.CULL_FRONT, clamp_depth=True, vertex_spec={0: np.uint32(cascade)}, name="Shadow map cascade %d" % cascade)
    be.new_frame()
    be.set_graphic_pass_execution(p, [(shadow_map, 0)], storage_buffers=[(cascades, True, 0), (transforms, True, 1)])
    be._check(api.b["prepare_for_drawcall_recording"](be.ctx), "prepare_for_drawcall_recording")
    tex = 0
    if albedo is not None:  # (w, h, RGBA8 bytes): the albedo texture of every draw (alpha test)
        tex = be.global_texture_index(be.create_image(albedo[0], albedo[1], "RGBA8", data=np.asarray(albedo[2], np.uint8)))
    be.draw_meshes([handles[d[0]] for d in draws], np.array([[tex, d[1]] for d in draws], np.uint32), p)
    rig.run()
    out = be.read_image(shadow_map, 0, np.uint16).reshape(size, size).copy()
    rig.close()
    return out


def _shadow_cascade_info(ffi, light_matrices):
    info = ffi.ShadowCascadeInfo()
    lm = np.asarray(light_matrices, np.float32).reshape(4, 16)
    for c in range(4):
        for k in range(16):
            info.lightMatrices[c][k] = float(lm[c, k])
    return np.frombuffer(bytes(info), np.uint8).copy()


# ---------------- gbufferShading.comp (triangle.frag recast over the packed G-buffer) as a single pass ----------------
def shade(ffi, api, gbuffer, y_sh, co_cg, noise_rg8, sun_direction, camera_position, diffuse_brdf=2, direct_multiscatter=0, geometry_aa=0,
          indirect_tech=0, sun_color=(1.0, 0.9, 0.8), sun_strength_exposed=2.5, brdf_res=512, shadow_maps=None, cascade_info=None, froxel_volume=None,
          froxel_max_distance=30.0, cascades=3):
    """Bindings of RenderFrontend::shadeGBuffer: G-buffer (h, w, 4) uint32, full-res Y_SH (h, w, 4) / CoCg (h, w, 2) float16, the BRDF LUT
    computed by brdfLut.comp in the same rig, shadow maps without casters (everything lit), an identity froxel volume (no fog).
    Optional: shadow_maps = 4 square uint16 D16 images + cascade_info = ffi.ShadowCascadeInfo (splits, light matrices, light-space scales),
    froxel_volume = (d, h, w, 4) float16 integrated in-scattering / transmittance.
    Returns (R11G11B10 colour (h, w) uint32, BRDF LUT (res, res, 4) float16, the global shader info used)."""
    h, w = gbuffer.shape[:2]
    rig = PassRig(ffi, api, w, h)
    be, g = rig.be, rig.g
    noise = be.create_image(noise_rg8.shape[1], noise_rg8.shape[0], "RG8", data=np.ascontiguousarray(noise_rg8, np.uint8))
    for i in range(4):
        g.noiseTextureIndices[i] = be.global_texture_index(noise)
    g.frameIndexMod4 = 1
    for i in range(3):
        g.sunDirection[i] = float(sun_direction[i])
        g.cameraPosition[i] = float(camera_position[i])
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    gb = be.create_image(w, h, "RGBA32_UINT", data=np.ascontiguousarray(gbuffer, np.uint32))
    ysh = be.create_image(w, h, "RGBA16_SFLOAT", data=np.ascontiguousarray(y_sh, np.float16))
    cocg = be.create_image(w, h, "RG16_SFLOAT", data=np.ascontiguousarray(co_cg, np.float16))
    lut = be.create_image(brdf_res, brdf_res, "RGBA16_SFLOAT")
    if shadow_maps is None:
        shadow_maps = [np.zeros((64, 64), np.uint16)] * 4
    shadow = [be.create_image(m.shape[1], m.shape[0], "DEPTH16", data=np.ascontiguousarray(m, np.uint16)) for m in shadow_maps]
    if froxel_volume is None:
        froxel_volume = np.zeros((4, 4, 8, 4), np.float16)
        froxel_volume[..., 3] = 1.0  # in-scattering 0, transmittance 1
    fd, fh, fw = froxel_volume.shape[:3]
    froxels = be.create_image(fw, fh, "RGBA16_SFLOAT", depth=fd, type_=ffi.IMAGE_3D, data=np.ascontiguousarray(froxel_volume, np.float16))
    sky = be.create_image(8, 4, "R11G11B10_UFLOAT", data=np.zeros((4, 8), np.uint32))
    transmission = be.create_image(8, 8, "R11G11B10_UFLOAT", data=np.zeros((8, 8), np.uint32))
    color = be.create_image(w, h, "R11G11B10_UFLOAT")
    light = be.create_storage_buffer(20, np.array([sun_color[0], sun_color[1], sun_color[2], 1.0, sun_strength_exposed], np.float32))
    info = cascade_info
    if info is None:
        info = ffi.ShadowCascadeInfo()
        for c in range(4):
            info.splits[c] = 1e9  # every pixel in cascade 0
            for k in range(16):
                info.lightMatrices[c][k] = 1.0 if k % 5 == 0 else 0.0
            info.lightSpaceScale[c][0] = info.lightSpaceScale[c][1] = 1.0
    cascade_buf = be.create_storage_buffer(304, np.frombuffer(bytes(info), np.uint8).copy())
    vol = be.create_uniform_buffer(52, np.array([0, 0, 0, 0, 1, 1, 1, froxel_max_distance, 1.0, 0.003, 0.008, 0.5, 0.2], np.float32))
    p_lut = be.create_compute_pass("brdfLut.comp", {0: np.int32(diffuse_brdf)})
    p = be.create_compute_pass("gbufferShading.comp", {0: np.int32(diffuse_brdf), 1: np.int32(direct_multiscatter), 2: np.uint32(geometry_aa), 3: np.int32(indirect_tech), 4: np.uint32(cascades)})
    be.new_frame()
    be.set_compute_pass_execution(p_lut, (brdf_res // 8, brdf_res // 8, 1), storage=[(lut, 0, 0)])
    sampled = [(gb, 0, 0), (lut, 0, 3), (ysh, 0, 15), (cocg, 0, 16), (froxels, 0, 18), (sky, 0, 21), (transmission, 0, 22)] + [(shadow[i], 0, 9 + i) for i in range(4)]
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=sampled, storage=[(color, 0, 20)], storage_buffers=[(light, True, 7), (cascade_buf, True, 8)],
                                  uniform_buffers=[(vol, 19)])
    rig.run()
    out = (be.read_image(color, 0, np.uint32).reshape(h, w).copy(), be.read_image(lut, 0, np.float16).reshape(brdf_res, brdf_res, 4).copy(), g)
    rig.close()
    return out


def taa_resolve(ffi, api, current, history, motion, depth, weights, use_clipping=True, use_dilation=True, history_tech=0, use_tonemap=True, camera_cut=False):
    """temporalFilter.comp with the bindings of TAA::computeTemporalFilter (TAA.cpp:139-166). current / history: packed R11G11B10 (h, w);
    motion (h, w, 2) int16 SNORM; depth (h, w) float32; weights: 9 floats. Returns (output, new history), packed."""
    h, w = current.shape
    rig = PassRig(ffi, api, w, h)
    be = rig.be
    if camera_cut:
        rig.g.cameraCut = 1
        be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(rig.g), np.uint8))
    img = lambda fmt, data: be.create_image(w, h, fmt, data=np.ascontiguousarray(data))
    cur, his = img("R11G11B10_UFLOAT", current.astype(np.uint32)), img("R11G11B10_UFLOAT", history.astype(np.uint32))
    vel, dep = img("RG16_SNORM", motion.astype(np.int16)), img("DEPTH32", depth.astype(np.float32))
    out, his_dst = be.create_image(w, h, "R11G11B10_UFLOAT"), be.create_image(w, h, "R11G11B10_UFLOAT")
    wbuf = be.create_uniform_buffer(36, np.asarray(weights, np.float32))
    p = be.create_compute_pass("temporalFilter.comp", {0: np.uint32(int(use_clipping)), 1: np.uint32(int(use_dilation)), 2: np.int32(history_tech), 3: np.uint32(int(use_tonemap))})
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(cur, 0, 0), (his, 0, 3), (vel, 0, 4), (dep, 0, 5)], storage=[(out, 0, 1), (his_dst, 0, 2)],
                                  uniform_buffers=[(wbuf, 6)])
    rig.run()
    res = (be.read_image(out, 0, np.uint32).reshape(h, w).copy(), be.read_image(his_dst, 0, np.uint32).reshape(h, w).copy())
    rig.close()
    return res


def gi_spatial_filter(ffi, api, y_sh, co_cg, depth_half, normal_rgba8, filter_index, camera_position, view_projection, frame_index_mod4=2):
    """filterIndirectDiffuseSpatial.comp with the bindings of SDFGI::filterIndirectDiffuse (SDFGI.cpp:430-447): half-res Y_SH (h, w, 4) /
    CoCg (h, w, 2) float16, R16F half-res depth (h, w) float16, RGBA8 normals (h, w, 4). Returns (Y_SH, CoCg) float16 and the globals used."""
    h, w = depth_half.shape
    rig = PassRig(ffi, api, w, h, screen=(2 * w, 2 * h))
    be, g = rig.be, rig.g
    g.frameIndexMod4 = frame_index_mod4
    for i in range(3):
        g.cameraPosition[i] = float(camera_position[i])
    for i, v in enumerate(np.asarray(view_projection, np.float32).T.ravel()):
        g.viewProjection[i] = float(v)
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    src_y = be.create_image(w, h, "RGBA16_SFLOAT", data=np.ascontiguousarray(y_sh, np.float16))
    src_c = be.create_image(w, h, "RG16_SFLOAT", data=np.ascontiguousarray(co_cg, np.float16))
    dep = be.create_image(w, h, "R16_SFLOAT", data=np.ascontiguousarray(depth_half, np.float16))
    nrm = be.create_image(w, h, "RGBA8", data=np.ascontiguousarray(normal_rgba8, np.uint8))
    out_y, out_c = be.create_image(w, h, "RGBA16_SFLOAT"), be.create_image(w, h, "RG16_SFLOAT")
    p = be.create_compute_pass("filterIndirectDiffuseSpatial.comp", {0: np.int32(filter_index)})
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(src_y, 0, 2), (src_c, 0, 3), (dep, 0, 4), (nrm, 0, 5)], storage=[(out_y, 0, 0), (out_c, 0, 1)])
    rig.run()
    res = (be.read_image(out_y, 0, np.float16).reshape(h, w, 4).copy(), be.read_image(out_c, 0, np.float16).reshape(h, w, 2).copy(), g)
    rig.close()
    return res


def gi_temporal_filter(ffi, api, y_sh, co_cg, hist_y, hist_c, motion_current, motion_last, camera_cut=False):
    """filterIndirectDiffuseTemporal.comp with the bindings of SDFGI::filterIndirectDiffuse (SDFGI.cpp:449-474): half-res Y_SH (h, w, 4) /
    CoCg (h, w, 2) float16 inputs and histories, full-res RG16_SNORM motion (2h, 2w, 2) int16 of this and of the last frame.
    Returns (Y_SH, CoCg, historyOut_Y_SH, historyOut_CoCg) float16."""
    h, w = y_sh.shape[:2]
    rig = PassRig(ffi, api, w, h, screen=(2 * w, 2 * h))
    be, g = rig.be, rig.g
    g.cameraCut = int(camera_cut)
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    img = lambda a, fmt, dt, ww=w, hh=h: be.create_image(ww, hh, fmt, data=np.ascontiguousarray(a, dt))
    srcs = [img(y_sh, "RGBA16_SFLOAT", np.float16), img(co_cg, "RG16_SFLOAT", np.float16), img(hist_y, "RGBA16_SFLOAT", np.float16), img(hist_c, "RG16_SFLOAT", np.float16),
            img(motion_current, "RG16_SNORM", np.int16, 2 * w, 2 * h), img(motion_last, "RG16_SNORM", np.int16, 2 * w, 2 * h)]
    outs = [be.create_image(w, h, f) for f in ("RGBA16_SFLOAT", "RG16_SFLOAT", "RGBA16_SFLOAT", "RG16_SFLOAT")]
    p = be.create_compute_pass("filterIndirectDiffuseTemporal.comp")
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), sampled=[(s, 0, 4 + i) for i, s in enumerate(srcs)], storage=[(o, 0, i) for i, o in enumerate(outs)])
    rig.run()
    res = tuple(be.read_image(o, 0, np.float16).reshape(h, w, c).copy() for o, c in zip(outs, (4, 2, 4, 2)))
    rig.close()
    return res


def gi_upscale(ffi, api, y_sh, co_cg, depth_full, depth_half):
    """indirectLightUpscale.comp with the bindings of SDFGI::filterIndirectDiffuse (SDFGI.cpp:476-497): half-res Y_SH / CoCg float16 and
    R16F depth (h, w), full-res D32F depth (H, W). Returns full-res (Y_SH, CoCg) float16 and the globals used."""
    H, W = depth_full.shape
    h, w = depth_half.shape
    rig = PassRig(ffi, api, W, H)
    be = rig.be
    src_y = be.create_image(w, h, "RGBA16_SFLOAT", data=np.ascontiguousarray(y_sh, np.float16))
    src_c = be.create_image(w, h, "RG16_SFLOAT", data=np.ascontiguousarray(co_cg, np.float16))
    d_full = be.create_image(W, H, "DEPTH32", data=np.ascontiguousarray(depth_full, np.float32))
    d_half = be.create_image(w, h, "R16_SFLOAT", data=np.ascontiguousarray(depth_half, np.float16))
    out_y, out_c = be.create_image(W, H, "RGBA16_SFLOAT"), be.create_image(W, H, "RG16_SFLOAT")
    p = be.create_compute_pass("indirectLightUpscale.comp")
    be.new_frame()
    be.set_compute_pass_execution(p, ((W + 7) // 8, (H + 7) // 8, 1), sampled=[(src_y, 0, 2), (src_c, 0, 3), (d_full, 0, 4), (d_half, 0, 5)], storage=[(out_y, 0, 0), (out_c, 0, 1)])
    rig.run()
    res = (be.read_image(out_y, 0, np.float16).reshape(H, W, 4).copy(), be.read_image(out_c, 0, np.float16).reshape(H, W, 2).copy(), rig.g)
    rig.close()
    return res


def froxels(ffi, api, res, noise_r8, shadow_d16, light_matrix2, light, settings13, history, camera, prev, sun_direction, camera_cut=False, pass_fusion=False):
    """The four froxel passes with the bindings of Volumetrics::computeVolumetricLighting (Volumetrics.cpp:136-247) in one frame:
    froxelVolumeMaterial -> froxelLightScattering -> volumeLightingReprojection -> volumetricLightingIntegration.
    res = (w, h, d) of the volumes; noise_r8 (n, n, n) uint8; shadow_d16 (s, s) uint16; light_matrix2 = column-major 16 floats of cascade 2;
    light = (sunColor rgb, previousFrameExposure, sunStrengthExposed) (lightBuffer.inc:4-8); settings13 = the 13 floats of VolumetricLightingSettings;
    history (d, h, w, 4) float16; camera = dict(position, forward, up, right, tan_fov_half, aspect); prev = dict(view_projection (4x4, row-major
    numpy), position, forward). Returns (material, scattering, reprojected, integrated) float16 volumes (d, h, w, 4) and the globals.
    pass_fusion: the product runs the chain as one launch over froxel columns (backend.cu planFusions) - material and scattering are then
    not written and come back as zeros."""
    w, h, d = res
    rig = PassRig(ffi, api, w * 8, h * 8)
    be, g = rig.be, rig.g
    be._check(api.b["set_pass_fusion_enabled"](be.ctx, 1 if pass_fusion else 0), "set_pass_fusion_enabled")
    for i in range(3):
        g.cameraPosition[i], g.cameraForward[i], g.cameraUp[i], g.cameraRight[i] = (float(camera[k][i]) for k in ("position", "forward", "up", "right"))
        g.cameraPositionPrevious[i], g.cameraForwardPrevious[i] = float(prev["position"][i]), float(prev["forward"][i])
        g.sunDirection[i] = float(sun_direction[i])
    g.cameraTanFovHalf, g.cameraAspectRatio = float(camera["tan_fov_half"]), float(camera["aspect"])
    for i, v in enumerate(np.asarray(prev["view_projection"], np.float32).T.ravel()):
        g.viewProjectionPrevious[i] = float(v)
    g.cameraCut = int(camera_cut)
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    n = noise_r8.shape[0]
    noise = be.create_image(n, n, "R8", depth=n, type_=ffi.IMAGE_3D, data=np.ascontiguousarray(noise_r8, np.uint8))
    s = shadow_d16.shape[0]
    shadow = be.create_image(s, s, "DEPTH16", data=np.ascontiguousarray(shadow_d16, np.uint16))
    vol = lambda data=None: be.create_image(w, h, "RGBA16_SFLOAT", depth=d, type_=ffi.IMAGE_3D, data=data)
    material, scatter, target, integrated = vol(), vol(), vol(), vol()
    hist = vol(np.ascontiguousarray(history, np.float16))
    lm = np.zeros((4, 16), np.float32)
    lm[2] = np.asarray(light_matrix2, np.float32)
    info = be.create_storage_buffer(304, _shadow_cascade_info(ffi, lm))
    lightbuf = be.create_storage_buffer(20, np.asarray(light, np.float32))
    ubo = be.create_uniform_buffer(52, np.asarray(settings13, np.float32).view(np.uint8))
    p = [be.create_compute_pass(name) for name in ("froxelVolumeMaterial.comp", "froxelLightScattering.comp", "volumeLightingReprojection.comp", "volumetricLightingIntegration.comp")]
    g4 = ((w + 3) // 4, (h + 3) // 4, (d + 3) // 4)
    be.new_frame()
    be.set_compute_pass_execution(p[0], g4, storage=[(material, 0, 0)], sampled=[(noise, 0, 1)], uniform_buffers=[(ubo, 2)])
    be.set_compute_pass_execution(p[1], g4, storage=[(scatter, 0, 0)], sampled=[(shadow, 0, 1), (material, 0, 2)], storage_buffers=[(info, True, 3), (lightbuf, True, 4)], uniform_buffers=[(ubo, 5)])
    be.set_compute_pass_execution(p[2], g4, storage=[(target, 0, 0)], sampled=[(scatter, 0, 1), (hist, 0, 2)], uniform_buffers=[(ubo, 3)])
    be.set_compute_pass_execution(p[3], ((w + 7) // 8, (h + 7) // 8, 1), storage=[(integrated, 0, 0)], sampled=[(target, 0, 1)], uniform_buffers=[(ubo, 2)])
    rig.run()
    out = tuple(be.read_image(v, 0, np.float16).reshape(d, h, w, 4).copy() for v in (material, scatter, target, integrated)) + (g,)
    rig.close()
    return out


def pre_expose_lights(ffi, api, histogram_u32, light, transmission_packed, sun_direction_y, screen, lum_min=0.001, lum_max=200000.0,
                      exposure_offset=1.0, adaption_speed=2.0, delta_time=1 / 60.0, sun_strength=128000.0):
    """preExposeLights.comp with the bindings of RenderFrontend::computeColorBufferHistogram (RenderFrontend.cpp:741-752).
    Returns the light buffer after the pass (5 floats: sunColor, previousFrameExposure, sunStrengthExposed)."""
    rig = PassRig(ffi, api, 64, 64, screen=screen)
    be, g = rig.be, rig.g
    g.sunDirection[0], g.sunDirection[1], g.sunDirection[2] = 0.0, float(sun_direction_y), 0.0
    g.exposureOffset, g.exposureAdaptionSpeedEvPerSec, g.deltaTime, g.sunStrength = exposure_offset, adaption_speed, delta_time, sun_strength
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    n_bins = len(histogram_u32)
    lh, lw = transmission_packed.shape
    lut = be.create_image(lw, lh, "R11G11B10_UFLOAT", data=np.ascontiguousarray(transmission_packed, np.uint32))
    hist = be.create_storage_buffer(n_bins * 4, np.ascontiguousarray(histogram_u32, np.uint32))
    lightbuf = be.create_storage_buffer(20, np.asarray(light, np.float32))
    p = be.create_compute_pass("preExposeLights.comp", {0: np.int32(n_bins), 1: np.float32(lum_min), 2: np.float32(lum_max)})
    be.new_frame()
    be.set_compute_pass_execution(p, (1, 1, 1), sampled=[(lut, 0, 2)], storage_buffers=[(lightbuf, False, 0), (hist, False, 1)])
    rig.run()
    out = be.read_storage_buffer(lightbuf, 20, np.float32).copy()
    rig.close()
    return out


def light_matrix(ffi, api, depth_min_max, camera, sun_direction, cascades=4, extra_padding=5.0, min_far_plane=30.0, near=0.1, far=300.0):
    """lightMatrix.comp with the bindings of RenderFrontend::computeSunLightMatrices (RenderFrontend.cpp:840-861): depth_min_max = the single
    RG32F texel of the lowest HiZ mip (min, max of the reverse-z depth). Returns the 304-byte ShadowCascadeInfo as ffi.ShadowCascadeInfo."""
    rig = PassRig(ffi, api, 64, 64)
    be, g = rig.be, rig.g
    for i in range(3):
        g.cameraPosition[i], g.cameraForward[i], g.cameraUp[i], g.cameraRight[i] = (float(camera[k][i]) for k in ("position", "forward", "up", "right"))
        g.sunDirection[i] = float(sun_direction[i])
    g.cameraTanFovHalf, g.cameraAspectRatio, g.nearPlane, g.farPlane = float(camera["tan_fov_half"]), float(camera["aspect"]), near, far
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    top = be.create_image(1, 1, "RG32_SFLOAT", data=np.asarray(depth_min_max, np.float32))
    info = be.create_storage_buffer(304, np.zeros(304, np.uint8))
    p = be.create_compute_pass("lightMatrix.comp", {0: np.uint32(cascades)})
    be.new_frame()
    be.set_compute_pass_execution(p, (1, 1, 1), storage=[(top, 0, 1)], storage_buffers=[(info, False, 0)], push=np.array([extra_padding, min_far_plane], np.float32).tobytes())
    rig.run()
    out = ffi.ShadowCascadeInfo.from_buffer_copy(be.read_storage_buffer(info, 304).tobytes())
    rig.close()
    return out


ATMOSPHERE_DEFAULT = [0.0058, 0.0135, 0.0331, 6371.0, 0.0058, 0.0135, 0.0331, 100.0, 0.000650, 0.001881, 0.000085, 0.006, 1.11 * 0.006, 0.75]  # Sky.h:6-15


def sky_luts(ffi, api, sun_direction, sun_strength_exposed, atmosphere=ATMOSPHERE_DEFAULT):
    """skyTransmissionLut.comp -> skyMultiscatterLut.comp -> skyLut.comp in one frame with the bindings and dispatch sizes of
    Sky::updateTransmissionLut / updateSkyLut (Sky.cpp:260-316). Returns the three LUTs as packed R11G11B10 (128x128, 32x32, 100x200)."""
    rig = PassRig(ffi, api, 64, 64)
    be, g = rig.be, rig.g
    for i in range(3):
        g.sunDirection[i] = float(sun_direction[i])
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    trans, multi = be.create_image(128, 128, "R11G11B10_UFLOAT"), be.create_image(32, 32, "R11G11B10_UFLOAT")
    sky = be.create_image(200, 100, "R11G11B10_UFLOAT", data=np.zeros(200 * 100, np.uint32))
    ubo = be.create_uniform_buffer(56, np.asarray(atmosphere, np.float32).view(np.uint8))
    light = be.create_storage_buffer(20, np.array([1, 1, 1, 1, sun_strength_exposed], np.float32))
    p = [be.create_compute_pass(n) for n in ("skyTransmissionLut.comp", "skyMultiscatterLut.comp", "skyLut.comp")]
    be.new_frame()
    be.set_compute_pass_execution(p[0], (16, 16, 1), storage=[(trans, 0, 0)], uniform_buffers=[(ubo, 1)])
    be.set_compute_pass_execution(p[1], (4, 4, 1), storage=[(multi, 0, 0)], sampled=[(trans, 0, 1)], uniform_buffers=[(ubo, 3)])
    be.set_compute_pass_execution(p[2], (25, 12, 1), storage=[(sky, 0, 0)], sampled=[(trans, 0, 1), (multi, 0, 2)], uniform_buffers=[(ubo, 4)], storage_buffers=[(light, True, 5)])
    rig.run()
    out = (be.read_image(trans, 0, np.uint32).reshape(128, 128).copy(), be.read_image(multi, 0, np.uint32).reshape(32, 32).copy(), be.read_image(sky, 0, np.uint32).reshape(100, 200).copy())
    rig.close()
    return out


def sdf_diffuse_trace(ffi, api, depth, normal_rgba8, noise_rg8, sky_lut_packed, instances, bricks, shadow_d16, light_matrix, light, influence_range=5.0,
                      strict_cutoff=False, camera_position=(0.0, 0.0, 0.0), frame_index_mod4=1, cascade=3):
    """sdfDiffuseTrace.comp with the bindings of SDFGI::diffuseSDFTrace (SDFGI.cpp:380-419) at half resolution: depth (h, w) float32 and
    RGBA8 normals (h, w, 4) sampled at uv = iUV / size, instances = [(localExtends, brick index into `bricks`, meanAlbedo, worldToLocal 4x4 numpy
    row-major)], bricks = [(res, uint16 R16F volume (res, res, res))], every culling tile lists every instance. Returns (Y_SH, CoCg) float16, globals."""
    h, w = depth.shape
    rig = PassRig(ffi, api, w, h, screen=(2 * w, 2 * h))
    be, g = rig.be, rig.g
    noise = be.create_image(noise_rg8.shape[1], noise_rg8.shape[0], "RG8", data=np.ascontiguousarray(noise_rg8, np.uint8))
    for i in range(4):
        g.noiseTextureIndices[i] = be.global_texture_index(noise)
    g.frameIndexMod4 = frame_index_mod4
    for i in range(3):
        g.cameraPosition[i] = float(camera_position[i])
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    brick_index = []
    for vol in bricks:
        r = vol.shape[0]
        brick_index.append(be.global_texture_index(be.create_image(r, r, "R16_SFLOAT", depth=r, type_=ffi.IMAGE_3D, data=np.ascontiguousarray(vol, np.uint16))))
    inst = np.zeros(4 + 24 * max(len(instances), 1), np.float32)      # sdfDiffuseTrace.comp:35-41: uint instanceCount + 3 padding words, then SDFInstance[]
    inst[0:1].view(np.uint32)[0] = len(instances)
    for k, (ext, b, albedo, world_to_local) in enumerate(instances):
        rec = inst[4 + 24 * k: 4 + 24 * (k + 1)]
        rec[0:3], rec[4:7] = ext, albedo
        rec[3:4].view(np.uint32)[0] = brick_index[b]
        rec[8:24] = np.asarray(world_to_local, np.float32).T.ravel()
    tiles_x, tiles_y = (2 * w + 31) // 32, (2 * h + 31) // 32          # tileIndexFromTileUV strides by the FULL-resolution width (sdfCulling.inc:17-20)
    tiles = np.zeros((tiles_x * tiles_y, 101), np.uint32)
    tiles[:, 0] = len(instances)
    tiles[:, 1:1 + len(instances)] = np.arange(len(instances), dtype=np.uint32)
    d_img = be.create_image(w, h, "DEPTH32", data=np.ascontiguousarray(depth, np.float32))
    n_img = be.create_image(w, h, "RGBA8", data=np.ascontiguousarray(normal_rgba8, np.uint8))
    sky = be.create_image(sky_lut_packed.shape[1], sky_lut_packed.shape[0], "R11G11B10_UFLOAT", data=np.ascontiguousarray(sky_lut_packed, np.uint32))
    shadow = be.create_image(shadow_d16.shape[1], shadow_d16.shape[0], "DEPTH16", data=np.ascontiguousarray(shadow_d16, np.uint16))
    out_y, out_c = be.create_image(w, h, "RGBA16_SFLOAT"), be.create_image(w, h, "RG16_SFLOAT")
    lm = np.zeros((4, 16), np.float32)
    lm[cascade] = np.asarray(light_matrix, np.float32)
    bufs = [(be.create_storage_buffer(20, np.asarray(light, np.float32)), True, 5), (be.create_storage_buffer(inst.nbytes, inst.view(np.uint8)), True, 6),
            (be.create_storage_buffer(tiles.nbytes, tiles.view(np.uint8)), True, 7), (be.create_storage_buffer(304, _shadow_cascade_info(ffi, lm)), True, 9)]
    ubo = be.create_uniform_buffer(4, np.array([influence_range], np.float32).view(np.uint8))
    p = be.create_compute_pass("sdfDiffuseTrace.comp", {0: np.uint32(int(strict_cutoff)), 1: np.int32(cascade)})
    be.new_frame()
    be.set_compute_pass_execution(p, ((w + 7) // 8, (h + 7) // 8, 1), storage=[(out_y, 0, 0), (out_c, 0, 1)], sampled=[(d_img, 0, 2), (n_img, 0, 3), (sky, 0, 4), (shadow, 0, 10)],
                                  storage_buffers=bufs, uniform_buffers=[(ubo, 8)])
    rig.run()
    res = (be.read_image(out_y, 0, np.float16).reshape(h, w, 4).copy(), be.read_image(out_c, 0, np.float16).reshape(h, w, 2).copy(), g)
    rig.close()
    return res


def sdf_culling(ffi, api, bbs, frustum_points, frustum_normals, influence_range, target_size, screen, camera, hiz_min_max=None, near=0.1, far=300.0):
    """sdfCameraFrustumCulling.comp + sdfCameraTileCulling.comp with the bindings of SDFGI::sdfInstanceCulling (SDFGI.cpp:538-630).
    bbs (n, 2, 3) world-space boxes; frustum_points / normals (6, 3); target_size = size of the traced image (tiles of 32 px);
    hiz_min_max (ty, tx, 2) float32 = the depth pyramid level bound at binding 4 (None: useHiZ = false).
    Returns (culled instance list, per-tile lists [(count, indices)] as an array (tiles, 101), tile stride)."""
    n = len(bbs)
    rig = PassRig(ffi, api, 64, 64, screen=screen)
    be, g = rig.be, rig.g
    for i in range(3):
        g.cameraPosition[i], g.cameraForward[i], g.cameraUp[i], g.cameraRight[i] = (float(camera[k][i]) for k in ("position", "forward", "up", "right"))
    g.cameraTanFovHalf, g.cameraAspectRatio, g.nearPlane, g.farPlane = float(camera["tan_fov_half"]), float(camera["aspect"]), near, far
    be.set_uniform_buffer_data(rig.gbuf, np.frombuffer(bytes(g), np.uint8))
    inst = np.zeros(4 + 24 * n, np.float32)
    inst[0:1].view(np.uint32)[0] = n
    bb = np.zeros((n, 8), np.float32)
    bb[:, 0:3], bb[:, 4:7] = np.asarray(bbs)[:, 0], np.asarray(bbs)[:, 1]
    fr = np.zeros((12, 4), np.float32)
    fr[0:6, :3], fr[6:12, :3] = frustum_points, frustum_normals
    tx, ty = (target_size[0] + 31) // 32, (target_size[1] + 31) // 32
    stride = (screen[0] + 31) // 32                                      # tileIndexFromTileUV: full-resolution width
    tiles_total = stride * max(ty, (screen[1] + 31) // 32)
    b_inst = be.create_storage_buffer(inst.nbytes, inst.view(np.uint8))
    b_list = be.create_storage_buffer(4 + 4 * max(n, 1), np.full(1 + max(n, 1), 0xFFFFFFFF, np.uint32).view(np.uint8))
    be.set_storage_buffer_data(b_list, np.zeros(1, np.uint32).view(np.uint8))      # the host zeroes the counter every frame (SDFGI.cpp:557-558)
    b_bb = be.create_storage_buffer(bb.nbytes, bb.view(np.uint8))
    b_tiles = be.create_storage_buffer(tiles_total * 404, np.full(tiles_total * 101, 0xFFFFFFFF, np.uint32).view(np.uint8))
    u_fr = be.create_uniform_buffer(192, fr.view(np.uint8))
    u_inf = be.create_uniform_buffer(4, np.array([influence_range], np.float32).view(np.uint8))
    use_hiz = hiz_min_max is not None
    if not use_hiz:
        hiz_min_max = np.zeros((1, 1, 2), np.float32)
    hz = be.create_image(hiz_min_max.shape[1], hiz_min_max.shape[0], "RG32_SFLOAT", data=np.ascontiguousarray(hiz_min_max, np.float32))
    p0 = be.create_compute_pass("sdfCameraFrustumCulling.comp")
    p1 = be.create_compute_pass("sdfCameraTileCulling.comp", {0: np.uint32(int(use_hiz))})
    be.new_frame()
    be.set_compute_pass_execution(p0, ((n + 63) // 64, 1, 1), storage_buffers=[(b_inst, True, 0), (b_list, False, 2), (b_bb, True, 3)], uniform_buffers=[(u_fr, 1), (u_inf, 4)])
    be.set_compute_pass_execution(p1, ((tx + 7) // 8, (ty + 7) // 8, 1), storage_buffers=[(b_list, True, 0), (b_bb, True, 1), (b_tiles, False, 2)], uniform_buffers=[(u_inf, 3)],
                                  sampled=[(hz, 0, 4)], push=np.array([tx, ty], np.uint32).tobytes())
    rig.run()
    lst = be.read_storage_buffer(b_list, 4 + 4 * max(n, 1), np.uint32).copy()
    tiles = be.read_storage_buffer(b_tiles, tiles_total * 404, np.uint32).reshape(tiles_total, 101).copy()
    rig.close()
    return lst[1:1 + lst[0]], tiles, stride
