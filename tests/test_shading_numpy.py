"""gbufferShading (the recast of triangle.frag:177-341 + brdf.inc) of the oracle against an independent float64 numpy restatement
written from the GLSL: Cook-Torrance GGX with the height-correlated Smith visibility and Schlick Fresnel, the McAuley / simplified /
scaled-GGX multiscatter lobes, the four diffuse BRDFs with the in / out Fresnel factors, SH L1 irradiance + dominant-direction
specular or constant ambient. Shadow maps without casters and an identity froxel volume keep the sun fully lit and the fog out.
The result is an R11G11B10 texel (6 / 5 mantissa bits), so agreement is asked to 2^-6 relative (2^-5 for blue)."""
import numpy as np
import pytest

import passes
from conftest import decode_r11g11b10

PI = 3.1415926535  # global.inc:44


def srgb_to_linear(c):
    return np.where(c <= 0.004045, c / 12.92, (np.abs(c + 0.055) / 1.055) ** 2.4)


def f_schlick(f0, f90, x):
    return f0 + (f90 - f0) * (1 - x)[..., None] ** 5


def bilinear_clamp(img, u, v):
    h, w = img.shape[:2]
    fx, fy = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(fx).astype(int), np.floor(fy).astype(int)
    ax, ay = (fx - x0)[..., None], (fy - y0)[..., None]
    cx = lambda i: np.clip(i, 0, w - 1)
    cy = lambda i: np.clip(i, 0, h - 1)
    t = lambda xi, yi: img[cy(yi), cx(xi)].astype(np.float64)
    return t(x0, y0) * (1 - ax) * (1 - ay) + t(x0 + 1, y0) * ax * (1 - ay) + t(x0, y0 + 1) * (1 - ax) * ay + t(x0 + 1, y0 + 1) * ax * ay


def ggx_single(r, f0, NoH, NoV, VoH, NoL):
    a = NoH * r
    k = r / (1 - NoH * NoH + a * a)
    D = k * k / PI
    r2 = r * r
    vis = 0.5 / (NoL * np.sqrt(NoV * NoV * (1 - r2) + r2) + NoV * np.sqrt(NoL * NoL * (1 - r2) + r2))
    return (D * vis)[..., None] * f_schlick(f0, 1.0, VoH)


def multiscatter(mode, r, NoL, f0, single, lut_rgb, lut_img):
    e_out = lut_rgb[..., 1]
    f_avg = f0 + (1 - f0) / 21
    if mode == 0:
        s = 1 - np.sqrt(r)
        e_avg = np.minimum(0.999, 0.409255 + s * (1.04997 + s * (-0.0761947 - 0.383026 * s)))
        e_in = bilinear_clamp(lut_img, r, NoL)[..., 1]
        unscaled = (1 - e_in) * (1 - e_out) / (3.1415 * (1 - e_avg))
        scaling = (f_avg * f_avg * e_avg[..., None]) / (1 - f_avg * (1 - e_avg)[..., None])
        return unscaled[..., None] * scaling
    if mode == 1:
        scaling = (f_avg * f_avg * e_out[..., None]) / (1 - f_avg * (1 - e_out)[..., None])
        return ((1 - e_out) / PI)[..., None] * scaling
    if mode == 2:
        return f0 * (1 / e_out - 1)[..., None] * single
    return np.zeros_like(single)


def np_shade(gbuffer, y_sh, co_cg, lut_img, g, sun_color, sun_strength_exposed, diffuse_brdf, direct_multiscatter, indirect_tech, sun_shadow=None, geometry_aa=False, fog=None):
    """sun_shadow: callable(pos, pixel_depth) -> (h, w) shadow factor (calcShadow); fog: callable(colour, pixel_depth) -> colour"""
    h, w = gbuffer.shape[:2]
    depth = gbuffer[..., 0].view(np.float32).astype(np.float64)
    sn = np.stack([(gbuffer[..., 1] & 0xFFFF).astype(np.uint16).view(np.int16), (gbuffer[..., 1] >> 16).astype(np.uint16).view(np.int16)], -1).astype(np.float64) / 32767
    sn = np.maximum(sn, -1)
    n = np.concatenate([sn, (1 - np.abs(sn[..., 0]) - np.abs(sn[..., 1]))[..., None]], -1)
    t = np.maximum(-n[..., 2], 0)
    n[..., 0] += np.where(n[..., 0] >= 0, -t, t)
    n[..., 1] += np.where(n[..., 1] >= 0, -t, t)
    N = n / np.linalg.norm(n, axis=-1, keepdims=True)
    albedo = srgb_to_linear(np.stack([gbuffer[..., 2] & 0xFF, (gbuffer[..., 2] >> 8) & 0xFF, (gbuffer[..., 2] >> 16) & 0xFF], -1) / 255.0)
    rough, metal = ((gbuffer[..., 2] >> 24) & 0xFF) / 255.0, (gbuffer[..., 3] & 0xFF) / 255.0
    fwd, up, right = (np.array(list(v)[:3], np.float64) for v in (g.cameraForward, g.cameraUp, g.cameraRight))
    cam, sun = np.array(list(g.cameraPosition)[:3], np.float64), np.array(list(g.sunDirection)[:3], np.float64)
    ys, xs = np.mgrid[0:h, 0:w]
    ndc = np.stack([(xs + 0.5) / g.screenResolution[0], (ys + 0.5) / g.screenResolution[1]], -1) * 2 - 1
    Vd = -fwd + g.cameraTanFovHalf * ndc[..., 1:2] * up - g.cameraTanFovHalf * g.cameraAspectRatio * ndc[..., 0:1] * right  # screenToWorld.inc
    to_pixel = -Vd / np.linalg.norm(Vd, axis=-1, keepdims=True)
    depth_linear = g.nearPlane * g.farPlane / (g.farPlane + (1 - depth) * (g.nearPlane - g.farPlane))
    pos = cam + to_pixel / (to_pixel @ fwd)[..., None] * depth_linear[..., None]
    r = np.maximum(rough * rough, 0.0045)
    if geometry_aa:  # GeometricAA.inc:4-20; dFdxFine / dFdyFine = differences inside the pixel's 2x2 quad
        xe, ye = (xs & ~1), (ys & ~1)
        N_U, N_V = N[ys, xe + 1] - N[ys, xe], N[ye + 1, xs] - N[ye, xs]
        variance = 0.25 * ((N_U * N_U).sum(-1) + (N_V * N_V).sum(-1))
        r = np.clip(np.sqrt(r * r + np.minimum(2 * variance, 0.18)), 0, 1)
    diffuse_color = (1 - metal)[..., None] * albedo
    L = sun / np.linalg.norm(sun)
    V = cam - pos
    V /= np.linalg.norm(V, axis=-1, keepdims=True)
    H = V + L
    H /= np.linalg.norm(H, axis=-1, keepdims=True)
    dot = lambda a, b: (a * b).sum(-1)
    NoH, NoL, VoH, LoV = np.maximum(dot(N, H), 0), np.clip(dot(N, L), 0, 1), np.abs(dot(V, H)), np.maximum(dot(V, L), 0)
    NoV = np.maximum(np.abs(dot(N, V)), 0.0001)
    f0 = 0.04 * (1 - metal)[..., None] + albedo * metal[..., None]
    pixel_depth = dot(cam - pos, -fwd)  # triangle.frag:200
    direct = np.maximum(dot(N, L), 0)[..., None] * np.array(sun_color)  # sunShadow = 1 without casters
    if sun_shadow is not None:
        direct = direct * sun_shadow(pos, pixel_depth)[..., None]
    lut_rgb = bilinear_clamp(lut_img, r, NoV)[..., :3]
    integral = lut_rgb[..., 2:3] * np.ones(3)
    if diffuse_brdf == 0:
        fr = diffuse_color / PI
    elif diffuse_brdf == 1:
        f90 = (0.5 * r + 2 * VoH * VoH * r)[..., None]
        fr = diffuse_color / PI * f_schlick(1.0, f90, NoL) * f_schlick(1.0, f90, NoV) * (1 * (1 - r) + r / 1.51)[..., None]
    elif diffuse_brdf == 2:
        f0d = VoH + (1 - VoH) ** 5
        f1 = (1 - 0.75 * (1 - NoL) ** 5) * (1 - 0.75 * (1 - NoV) ** 5)
        gg = np.log2(2 / (r * r) - 1) / 18
        tt = np.clip(2.2 * gg - 0.5, 0, 1)
        fd = f0d + (f1 - f0d) * tt
        fb = (34.5 * gg * gg - 59 * gg + 24.5) * VoH * 2.0 ** (-np.maximum(73.2 * gg - 21.2, 8.9) * np.sqrt(NoH))
        fr = diffuse_color / PI * (fd + fb)[..., None]
    else:
        facing = 0.5 + 0.5 * LoV
        rough_term = facing * (0.9 - 0.4 * facing) * (0.5 + NoH) / np.maximum(NoH, 0.03)
        smooth = 1.05 * (1 - (1 - NoL) ** 5) * (1 - (1 - NoV) ** 5)
        single_c = (smooth * (1 - r) + rough_term * r) / PI
        fr = diffuse_color * (single_c[..., None] + diffuse_color * (0.1159 * r)[..., None])
        multi_integral = 0.1159 * r * PI * 2 * (1 - (0.04 + 0.96 * (1 - NoV) ** 5)) * 0.94291
        integral = np.minimum(lut_rgb[..., 2:3] + diffuse_color * multi_integral[..., None], 1.0)
    diffuse_direct = fr * direct * (1 - f_schlick(f0, 1.0, NoV)) * (1 - f_schlick(f0, 1.0, NoL))
    single = ggx_single(r, f0, NoH, NoV, VoH, NoL)
    specular_direct = direct * (single + multiscatter(direct_multiscatter, r, NoL, f0, single, lut_rgb, lut_img))
    if indirect_tech == 0:
        ysh, cc = y_sh.astype(np.float64), co_cg.astype(np.float64)
        sh = np.stack([np.full(N.shape[:-1], 1 / (2 * np.sqrt(PI))), -np.sqrt(3) * N[..., 1] / (2 * np.sqrt(PI)), np.sqrt(3) * N[..., 2] / (2 * np.sqrt(PI)), -np.sqrt(3) * N[..., 0] / (2 * np.sqrt(PI))], -1)
        sh /= np.linalg.norm(sh, axis=-1, keepdims=True)
        ycc = lambda y: np.stack([y + cc[..., 0] - cc[..., 1], y + cc[..., 1], y - cc[..., 0] - cc[..., 1]], -1)  # YCoCgToLinear
        diffuse_indirect = ycc(dot(ysh, sh)) * diffuse_color * integral
        dom = np.stack([-ysh[..., 3], -ysh[..., 1], ysh[..., 2]], -1)
        dl = np.clip(np.linalg.norm(dom, axis=-1), 0.01, 1)
        r_i = 1 * (1 - np.sqrt(dl)) + r * np.sqrt(dl)
        L_i = dom / dl[..., None]
        H_i = L_i + V
        H_i /= np.linalg.norm(H_i, axis=-1, keepdims=True)
        NoH_i, NoL_i, VoH_i = np.maximum(dot(N, H_i), 0), np.maximum(dot(N, L_i), 0), np.maximum(dot(V, H_i), 0)
        single_i = ggx_single(r_i, f0, NoH_i, NoV, VoH_i, NoL_i)
        multi_i = multiscatter(direct_multiscatter, r_i, NoL_i, f0, single_i, lut_rgb, lut_img)
        indirect = diffuse_indirect + (single_i + multi_i) * ycc(ysh[..., 0])
    else:
        amb = 0.003 * sun_strength_exposed
        indirect = amb * diffuse_color * integral + (lut_rgb[..., 0:1] * (1 - f0) + lut_rgb[..., 1:2] * f0) * amb
    colour = (diffuse_direct + specular_direct) * sun_strength_exposed + indirect
    return fog(colour, pixel_depth) if fog is not None else colour


def oct_encode(n):
    n = n / np.abs(n).sum(-1, keepdims=True)
    xy = n[..., :2].copy()
    neg = n[..., 2] < 0
    xy[neg] = ((1 - np.abs(n[..., 1::-1])) * np.where(n[..., :2] >= 0, 1.0, -1.0))[neg]
    q = np.round(np.clip(xy, -1, 1) * 32767).astype(np.int32)
    return (q[..., 0] & 0xFFFF).astype(np.uint32) | ((q[..., 1] & 0xFFFF).astype(np.uint32) << 16)


@pytest.mark.parametrize("diffuse_brdf,direct_multiscatter,indirect_tech", [(2, 0, 0), (2, 0, 1), (0, 1, 0), (1, 2, 0), (3, 3, 0), (3, 0, 1)])
def test_shading_matches_float64_restatement(ffi, oracle, diffuse_brdf, direct_multiscatter, indirect_tech):
    rng = np.random.default_rng(100 + diffuse_brdf * 10 + direct_multiscatter)
    w, h = 48, 32
    cam = np.array([0.5, -1.0, 2.0])
    # G-buffer: surfaces 2 .. 40 m away, normals in the hemisphere facing the camera (which looks down -z in the single-pass rig)
    depth_linear = rng.uniform(2.0, 40.0, (h, w))
    near, far = 0.1, 300.0
    depth = (1 - (near * far / depth_linear - far) / (near - far)).astype(np.float32)
    nrm = rng.normal(0, 1, (h, w, 3))
    nrm[..., 2] = np.abs(nrm[..., 2]) + 0.3
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    gb = np.zeros((h, w, 4), np.uint32)
    gb[..., 0] = depth.view(np.uint32)
    gb[..., 1] = oct_encode(nrm)
    alb = rng.integers(20, 250, (h, w, 3)).astype(np.uint32)
    gb[..., 2] = alb[..., 0] | (alb[..., 1] << 8) | (alb[..., 2] << 16) | (rng.integers(30, 250, (h, w)).astype(np.uint32) << 24)
    gb[..., 3] = np.where(rng.uniform(size=(h, w)) < 0.3, 255, rng.integers(0, 60, (h, w))).astype(np.uint32)
    # irradiance SH: positive luminance, a dominant direction of moderate length, small chroma
    ysh = np.concatenate([rng.uniform(0.2, 1.5, (h, w, 1)), rng.uniform(-0.25, 0.25, (h, w, 3))], -1).astype(np.float16)
    cocg = rng.uniform(-0.05, 0.05, (h, w, 2)).astype(np.float16)
    noise = rng.integers(0, 256, (32, 32, 2)).astype(np.uint8)
    sun = np.array([0.3, -0.6, 0.74])
    sun_color, sse = (1.0, 0.9, 0.8), 2.5
    packed, lut, g = passes.shade(ffi, oracle, gb, ysh, cocg, noise, sun, cam, diffuse_brdf, direct_multiscatter, 0, indirect_tech, sun_color, sse, brdf_res=128)
    got = decode_r11g11b10(packed)
    want = np_shade(gb, ysh, cocg, lut, g, sun_color, sse, diffuse_brdf, direct_multiscatter, indirect_tech)
    assert np.isfinite(want).all()
    want = np.maximum(want, 0)  # the packed format has no sign: negative irradiance (a random SH can produce it) stores as 0
    bright = want > 1e-3  # below that the 5-bit exponent of the packed format runs into denormals
    rel = np.abs(got - want) / np.maximum(want, 1e-9)
    assert bright.mean() > 0.9
    assert rel[..., :2][bright[..., :2]].max() < 2.0 ** -6 * 1.25, "red / green: 6 mantissa bits"
    assert rel[..., 2][bright[..., 2]].max() < 2.0 ** -5 * 1.25, "blue: 5 mantissa bits"
    assert np.median(rel[bright]) < 2.0 ** -7


def test_shadow_cascades_pcf_fog_and_geometric_aa(ffi, oracle):
    """The rest of triangle.frag's main(): cascade selection by view depth (:224-239), the 12-tap spiral PCF with the blue-noise rotation
    (calcShadow :92-120), geometric specular anti-aliasing from the quad's normal derivatives (GeometricAA.inc) and the froxel in-scattering /
    transmittance applied at the jittered screen position (applyVolumetricLighting :133-144, volumetricFroxelLighting.inc:33-53)."""
    from test_froxels_numpy import depth_to_uvz, trilinear
    rng = np.random.default_rng(77)
    w, h = 64, 40
    cam = np.array([0.5, -1.0, 2.0])
    near, far = 0.1, 300.0
    ys, xs = np.mgrid[0:h, 0:w]
    depth_linear = 3.0 + 36.0 * xs / (w - 1) + 0.3 * rng.random((h, w))           # 3 .. 39 m from left to right: all four cascades
    depth = (1 - (near * far / depth_linear - far) / (near - far)).astype(np.float32)
    # smooth normal field with a few creases (geometric AA widens the lobe only there)
    nrm = np.stack([0.3 * np.sin(xs / 6.0), 0.3 * np.cos(ys / 5.0), np.ones((h, w))], -1)
    nrm[:, 20:22, 0] += 0.8
    nrm[12:14, :, 1] -= 0.7
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    gb = np.zeros((h, w, 4), np.uint32)
    gb[..., 0] = depth.view(np.uint32)
    gb[..., 1] = oct_encode(nrm)
    alb = rng.integers(90, 250, (h, w, 3)).astype(np.uint32)
    gb[..., 2] = alb[..., 0] | (alb[..., 1] << 8) | (alb[..., 2] << 16) | (rng.integers(20, 120, (h, w)).astype(np.uint32) << 24)   # rather smooth: a visible lobe
    gb[..., 3] = np.where(rng.uniform(size=(h, w)) < 0.3, 255, 0).astype(np.uint32)
    ysh = np.concatenate([rng.uniform(0.2, 1.0, (h, w, 1)), rng.uniform(-0.2, 0.2, (h, w, 3))], -1).astype(np.float16)
    cocg = rng.uniform(-0.05, 0.05, (h, w, 2)).astype(np.float16)
    noise = rng.integers(0, 256, (32, 32, 2)).astype(np.uint8)
    sun = np.array([0.3, -0.6, 0.74])
    sun_color, sse = (1.0, 0.9, 0.8), 2.5
    # four cascades: top-down orthographic light matrices of different extent, maps of 8-texel blocks at two occluder depths
    splits = [10.0, 20.0, 30.0]
    info = ffi.ShadowCascadeInfo()
    mats, scales, maps = [], [], []
    for c in range(4):
        ext = 25.0 + 10.0 * c
        L = np.array([[1 / ext, 0, 0, 0.1 * c], [0, 0, 1 / ext, -0.05 * c], [0, 1 / 40.0, 0, 0.5], [0, 0, 0, 1]], np.float64)
        mats.append(L)
        scales.append(np.array([1.5 + 0.5 * c, 1.0 + 0.3 * c]))
        block = ((np.add.outer(np.arange(64) // 8, np.arange(64) // 8) + c) % 3 == 0)
        maps.append(np.where(block, 60000, 3000).astype(np.uint16))              # occluder depth 0.92 (shadows what is below) / 0.05
        for k, v in enumerate(L.T.ravel()):
            info.lightMatrices[c][k] = float(v)
        info.lightSpaceScale[c][0], info.lightSpaceScale[c][1] = float(scales[c][0]), float(scales[c][1])
    for c in range(3):
        info.splits[c] = splits[c]
    nz = noise[ys % 32, xs % 32].astype(np.float64) / 255.0                       # g_sampler_nearestRepeat at gl_FragCoord / textureSize

    def sun_shadow(pos, pixel_depth):
        cascade = sum((pixel_depth >= s).astype(int) for s in splits)
        out = np.zeros((h, w))
        for c in range(4):
            pl = np.concatenate([pos, np.ones((h, w, 1))], -1) @ mats[c].T
            pl = pl / pl[..., 3:4]
            uv = pl[..., :2] * 0.5 + 0.5
            actual = np.clip(pl[..., 2], 0, 1)
            lit = np.zeros((h, w))
            for i in range(12):
                d = np.sqrt((i + 0.5 * nz[..., 0]) / 12)
                angle = nz[..., 0] * 2 * PI + 2 * PI * i / 12
                sp = uv + np.stack([np.cos(angle), np.sin(angle)], -1) * (0.03 * scales[c]) * d[..., None]
                tx, ty = np.floor(sp[..., 0] * 64).astype(int), np.floor(sp[..., 1] * 64).astype(int)
                inside = (tx >= 0) & (tx < 64) & (ty >= 0) & (ty < 64)
                texel = np.where(inside, maps[c][np.clip(ty, 0, 63), np.clip(tx, 0, 63)] / 65535.0, 0.0)   # black border
                lit += actual >= texel
            out = np.where(cascade == c, lit / 12, out)
        return out
    volume = np.zeros((8, 5, 8, 4), np.float16)
    zz, yy, xx = np.mgrid[0:8, 0:5, 0:8]
    volume[..., :3] = (0.02 * (zz + 1) * (1 + 0.3 * np.sin(xx)))[..., None] * np.array([1.0, 0.8, 0.6])
    volume[..., 3] = np.exp(-0.12 * (zz + 1) * (1 + 0.2 * np.cos(yy)))

    def fog(colour, pixel_depth):
        jitter = (nz - 0.5) * 0.013
        suv = np.stack([(xs + 0.5) / w + jitter[..., 0], (ys + 0.5) / h + jitter[..., 1], depth_to_uvz(pixel_depth, 30.0)], -1)
        it = trilinear(volume, suv, repeat=False)
        return colour * it[..., 3:4] + it[..., :3]
    kw = dict(shadow_maps=maps, cascade_info=info, froxel_volume=volume, cascades=4)
    for aa in (0, 1):
        packed, lut, g = passes.shade(ffi, oracle, gb, ysh, cocg, noise, sun, cam, 2, 0, aa, 0, sun_color, sse, brdf_res=128, **kw)
        got = decode_r11g11b10(packed)
        seen = {}
        want = np_shade(gb, ysh, cocg, lut, g, sun_color, sse, 2, 0, 0, sun_shadow=lambda p, d: seen.setdefault("s", sun_shadow(p, d)), geometry_aa=bool(aa), fog=fog)
        s = seen["s"]
        assert 0.15 < s.mean() < 0.85 and ((s > 0.05) & (s < 0.95)).mean() > 0.05      # lit, shadowed and penumbra pixels
        assert all(((s > 0) & (np.floor((depth_linear - 0.001) / 10).clip(0, 3) == c)).any() for c in range(4))
        rel = np.abs(got - want) / np.maximum(want, 1e-6)
        ok = (rel[..., :2].max(-1) < 2.0 ** -6 * 1.25) & (rel[..., 2] < 2.0 ** -5 * 1.25)
        # a PCF tap within rounding distance of a texel border or a pixel on a cascade split may fall on the other side: a few pixels
        print("geometric AA %d: %.4f of the pixels within the packed format, median rel %.5f" % (aa, ok.mean(), float(np.median(rel))))
        assert ok.mean() > 0.995, "geometric AA %d: %d of %d pixels differ (median %.4f)" % (aa, int((~ok).sum()), ok.size, float(np.median(rel)))
        if aa:
            assert not np.array_equal(packed, first)                                   # the creases did widen some lobes
        first = packed
