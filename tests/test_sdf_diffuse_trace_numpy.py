"""sdfDiffuseTrace.comp main() (S2) of the oracle against a float64 numpy restatement written from the GLSL - everything around the
sphere tracer, which tests/test_sdf_trace_analytic.py pins by itself: world position of a half-res pixel at uv = iUV / size, the blue-noise
cosine-weighted ray (sampling.inc:25-45), the sky-view LUT lookup of a miss (sky.inc:85-116), hit shading (meanAlbedo^2.2 x shadow x sun,
the strict influence-radius cut), the 3x3 resolve inside the 8x8 group with its `greaterThan(rayIndex, 0)` quirk and normal / depth
rejection (:70-116), YCoCg and the L1 spherical-harmonics encoding. Whether a ray hits the one box of the scene is decided analytically
(ray / box intersection); the colour of a hit does not depend on where exactly it hit."""
import numpy as np
import pytest

import passes
from conftest import decode_r11g11b10
from test_gi_temporal_upscale_numpy import bilinear
from test_sdf_trace_analytic import box_brick

PI = 3.1415926535


def normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def np_trace(depth, normal_rgba8, noise, sky, g, box, hit_colour_fn):
    h, w = depth.shape
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = xs / w, ys / h                                                    # :120 no half-texel offset
    tex = lambda img: img[np.clip(np.floor(v * h).astype(int), 0, h - 1), np.clip(np.floor(u * w).astype(int), 0, w - 1)]
    d = tex(depth).astype(np.float64)
    lin = g.nearPlane * g.farPlane / (g.farPlane + (1 - d) * (g.nearPlane - g.farPlane))
    fwd, up, right = (np.array(list(x)[:3], np.float64) for x in (g.cameraForward, g.cameraUp, g.cameraRight))
    cam = np.array(list(g.cameraPosition)[:3], np.float64)
    ndc = np.stack([u, v], -1) * 2 - 1
    V = -normalize(-fwd + g.cameraTanFovHalf * ndc[..., 1:2] * up - g.cameraTanFovHalf * g.cameraAspectRatio * ndc[..., 0:1] * right)
    p_world = cam + V / (V @ fwd)[..., None] * lin[..., None]
    nh, nw = noise.shape[:2]
    xi = noise[ys % nh, xs % nw].astype(np.float64) / 255.0                  # nearest / repeat at iUV / textureSize
    N = tex(normal_rgba8)[..., :3].astype(np.float64) / 255.0 * 2 - 1
    origin = p_world + N * 0.2
    phi = 2 * PI * xi[..., 1]                                                # importanceSampleCosine
    cos_t, sin_t = np.sqrt(xi[..., 0]), np.sqrt(1 - xi[..., 0])
    up_v = np.where((np.abs(N[..., 2]) < 0.999)[..., None], np.array([0, 0, 1.0]), np.array([1.0, 0, 0]))
    tangent = normalize(np.cross(up_v, N))
    bitangent = np.cross(N, tangent)
    L = (np.cos(phi) * sin_t)[..., None] * tangent + (np.sin(phi) * sin_t)[..., None] * bitangent + cos_t[..., None] * N
    # hit or miss: ray (direction normalised as SDF.inc:106-107 does) against the box
    centre, half = box
    hit, t_hit = np.zeros((h, w), bool), np.zeros((h, w))
    if centre is not None:
        dirn = normalize(L)
        o = origin - centre
        with np.errstate(divide="ignore", invalid="ignore"):
            t1, t2 = (-half - o) / dirn, (half - o) / dirn
        t_enter, t_exit = np.minimum(t1, t2).max(-1), np.maximum(t1, t2).min(-1)
        hit, t_hit = (t_exit > t_enter) & (t_enter > 0), t_enter
    # miss: the sky-view LUT in direction L (toSkyLut, sampleSkyLut)
    theta = np.arccos(np.clip(-L[..., 1], -1, 1))
    y_low = theta / PI * 2 - 1
    sv = np.clip(np.sign(y_low) * np.sqrt(np.abs(y_low)) * 0.5 + 0.5, 0.005, 0.995)
    su = -np.arctan2(L[..., 2], L[..., 0]) / (2 * 3.1415) + 0.5
    colour = bilinear(sky, su, sv, repeat=True)
    if hit.any():
        colour = np.where(hit[..., None], hit_colour_fn(origin + normalize(L) * t_hit[..., None], t_hit), colour)
    # 3x3 resolve inside the 8x8 group (:70-116): neighbours with local index 0 are skipped (`greaterThan(rayIndex, 0)`)
    lx, ly = xs % 8, ys % 8
    acc, total = colour.copy(), np.ones((h, w))
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            nx, ny = lx + dx, ly + dy
            valid = (nx > 0) & (ny > 0) & (nx < 8) & (ny < 8)
            gx, gy = np.clip(xs + dx, 0, w - 1), np.clip(ys + dy, 0, h - 1)
            valid &= (xs + dx < w) & (ys + dy < h)                            # (image sizes here are multiples of 8: always true)
            non = np.clip((N * N[gy, gx]).sum(-1), 0, 1)
            ok = valid & (non > 0.9) & (np.abs(lin - lin[gy, gx]) < 0.5)
            wgt = (1.0 if dx == 0 else 0.5) * (1.0 if dy == 0 else 0.5)
            acc += np.where(ok[..., None], wgt * colour[gy, gx], 0.0)
            total += np.where(ok, wgt, 0.0)
    colour = acc / total[..., None]
    ycocg = np.stack([colour[..., 0] * 0.25 + 0.5 * colour[..., 1] + 0.25 * colour[..., 2], colour[..., 0] * 0.5 - 0.5 * colour[..., 2],
                      -colour[..., 0] * 0.25 + 0.5 * colour[..., 1] - 0.25 * colour[..., 2]], -1)
    sh = np.stack([np.full((h, w), 1 / (2 * np.sqrt(PI))), -np.sqrt(3) * L[..., 1] / (2 * np.sqrt(PI)), np.sqrt(3) * L[..., 2] / (2 * np.sqrt(PI)), -np.sqrt(3) * L[..., 0] / (2 * np.sqrt(PI))], -1)
    return ycocg[..., 0:1] * normalize(sh), ycocg[..., 1:], hit, t_hit


def wall_inputs(rng, w, h):
    near, far = 0.1, 300.0
    ys, xs = np.mgrid[0:h, 0:w]
    lin = np.where(xs >= 40, 10.8, 10.0)                                      # a step of 0.8 m: the resolve must not blend across it
    depth = (1 - (near * far / lin - far) / (near - far)).astype(np.float32)
    n = np.zeros((h, w, 3))
    n[..., 2] = 1.0
    n[20:28, :, :] = np.array([0.6, 0.0, 0.8])                                # a band of tilted normals: N.N = 0.8 < 0.9
    normal = np.zeros((h, w, 4), np.uint8)
    normal[..., :3] = np.round((n * 0.5 + 0.5) * 255)
    noise = rng.integers(0, 256, (32, 32, 2)).astype(np.uint8)
    ly, lx = np.mgrid[0:100, 0:200]
    smooth = np.stack([1.0 + 0.5 * np.sin(lx / 17.0 + k) * np.cos(ly / 9.0 - k) for k in range(3)], -1) * np.array([0.4, 0.7, 1.2])
    sky_p = np.zeros((100, 200), np.uint32)
    from conftest import decode_small_float
    for c, (mbits, shift) in enumerate(((6, 0), (6, 11), (5, 22))):
        vals = decode_small_float(np.arange(1 << (5 + mbits)), mbits)
        vals[~np.isfinite(vals)] = np.inf
        sky_p |= np.abs(smooth[..., c][..., None] - vals[None, None, :]).argmin(-1).astype(np.uint32) << shift
    return depth, normal, noise, sky_p


def compare(got_y, got_c, want_y, want_c, allowed):
    err = np.maximum(np.abs(got_y.astype(np.float64) - want_y).max(-1), np.abs(got_c.astype(np.float64) - want_c).max(-1))
    scale = np.maximum(np.abs(want_y).max(-1), 1e-3)
    bad = err > 2.0 ** -10 * 2.5 * scale + 1e-6
    assert bad.mean() <= allowed, "%d of %d texels differ, worst %.3e" % (int(bad.sum()), bad.size, float((err / scale).max()))
    return bad


def test_misses_sample_the_sky_lut_and_resolve(ffi, oracle):
    rng = np.random.default_rng(5)
    w, h = 64, 40
    depth, normal, noise, sky_p = wall_inputs(rng, w, h)
    L = np.eye(4)
    got_y, got_c, g = passes.sdf_diffuse_trace(ffi, oracle, depth, normal, noise, sky_p, [], [], np.zeros((8, 8), np.uint16), L.T.ravel(), [1, 1, 1, 1, 1])
    want_y, want_c, hit, _ = np_trace(depth, normal, noise, decode_r11g11b10(sky_p), g, (None, None), None)
    assert not hit.any()
    compare(got_y, got_c, want_y, want_c, 0.002)                              # (a noise texel of xi.x = 0 / N.z at the 0.999 switch: none here)
    assert np.abs(want_c).max() > 0.05 and want_y[..., 0].min() > 0.05


@pytest.mark.parametrize("strict", [False, True])
def test_hits_are_shaded_by_the_sun_and_cut_by_the_influence_radius(ffi, oracle, strict):
    rng = np.random.default_rng(6)
    w, h = 64, 40
    depth, normal, noise, sky_p = wall_inputs(rng, w, h)
    half, centre = np.array([1.6, 1.2, 0.9]), np.array([0.3, 0.2, -6.5])        # between the wall (z = -10 / -10.8) and the camera
    pad = np.maximum(2 * half * 0.075, 0.5)
    W2L = np.eye(4)
    W2L[:3, 3] = -centre
    albedo = np.array([0.7, 0.5, 0.3])
    # sun straight down onto a 40 m box; the shadow map shadows everything with light-space u >= 0.5 (world x >= 0)
    Lm = np.array([[1 / 20.0, 0, 0, 0], [0, 0, 1 / 20.0, 0], [0, 1 / 40.0, 0, 0.5], [0, 0, 0, 1]], np.float64)
    shadow = np.zeros((16, 16), np.uint16)
    shadow[:, 8:] = 65535
    light = np.array([1.0, 0.9, 0.8, 1.0, 2.0], np.float32)
    got_y, got_c, g = passes.sdf_diffuse_trace(ffi, oracle, depth, normal, noise, sky_p, [(2 * (half + pad), 0, albedo, W2L)], [box_brick(half, 32).reshape(32, 32, 32)],
                                               shadow, Lm.T.ravel(), light, influence_range=3.4, strict_cutoff=strict)

    def hit_colour(pos, t):
        lit = (pos[..., 0] < 0).astype(np.float64)                            # simpleShadow: lit where the map holds 0
        c = albedo ** 2.2 * lit[..., None] * light[4] * light[:3].astype(np.float64)
        return np.where((t >= 3.4)[..., None], 0.0, c) if strict else c
    want_y, want_c, hit, t_hit = np_trace(depth, normal, noise, decode_r11g11b10(sky_p), g, (centre, half), hit_colour)
    assert 0.1 < hit.mean() < 0.7 and (t_hit[hit] > 3.4).any() and (t_hit[hit] < 3.4).any()
    # rays grazing the box's edges (the tri-linear brick rounds them), hits within the threshold of the 3.4 m cut or of the shadow edge
    # differ - and each such ray also enters the resolve of up to 8 neighbours
    bad = compare(got_y, got_c, want_y, want_c, 0.06)
    print("strict %s: %.4f of the texels differ, %.4f of the interior hits" % (strict, bad.mean(), bad[hit & (np.abs(t_hit - 3.4) > 0.15)].mean()))
    interior = hit & (np.abs(t_hit - 3.4) > 0.15)
    assert bad[interior].mean() < 0.05
