"""The four froxel passes (S9: froxelVolumeMaterial.comp, froxelLightScattering.comp, volumeLightingReprojection.comp,
volumetricLightingIntegration.comp + volumetricFroxelLighting.inc, volumeShading.inc) of the oracle against independent float64 numpy
restatements written from the GLSL: exponential slice distribution (k = 3), world position of a froxel with the sub-froxel sample
offset, tri-linear wrapped density noise, sun shadow through cascade 2 with the black-border sampler, Henyey-Greenstein phase, the
luminance-of-extinction "transmittance" (0.21 / 0.72 / 0.07), reprojection into the previous frustum with the 0.95 moving average and
its off-frustum / camera-cut paths, and the front-to-back integration over res.z + 1 slices. Each pass is recomputed from the ORACLE's
output of the pass before it, so a comparison isolates one dispatch. Volumes are RGBA16F: agreement to half precision."""
import numpy as np
import pytest

import passes
from test_raster_oracle import perspective

K = 3.0  # volumetricFroxelLighting.inc:20


def uvz_to_depth(uvz, max_distance):  # :23-31
    return (np.exp(K * uvz) - 1) / (np.exp(K) - 1) * max_distance


def depth_to_uvz(depth, max_distance):  # :33-41
    with np.errstate(invalid="ignore"):
        return np.log(depth / max_distance * (np.exp(K) - 1) + 1) / K


def normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def froxel_world(res, offset, cam):  # froxelVolumeMaterial.comp:25-29 == froxelLightScattering.comp:39-43 == volumeLightingReprojection.comp:31-38
    w, h, d = res
    zs, ys, xs = np.mgrid[0:d, 0:h, 0:w]
    uv = np.stack([(xs + 0.5 + offset) / w, (ys + 0.5 + offset) / h, (zs + 0.5 + offset) / d], -1)
    ndc = 2 * (uv - 0.5)
    V = normalize(-cam["forward"] + cam["tan_fov_half"] * ndc[..., 1:2] * cam["up"] - cam["tan_fov_half"] * cam["aspect"] * ndc[..., 0:1] * cam["right"])
    return cam["position"] - V / (-V @ cam["forward"])[..., None] * uvz_to_depth(uv[..., 2], MAX_DISTANCE)[..., None], V


def trilinear(vol, uvw, repeat):
    """VK linear filter on a 3-D image (d, h, w, c); clamp-to-edge or repeat addressing"""
    d, h, w = vol.shape[:3]
    vol = vol.astype(np.float64).reshape(d, h, w, -1)
    p = uvw * np.array([w, h, d]) - 0.5
    p0 = np.floor(p)
    f = p - p0
    def idx(i, n):
        i = i.astype(np.int64)
        return np.mod(i, n) if repeat else np.clip(i, 0, n - 1)
    out = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                wgt = (f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1]) * (f[..., 2] if dz else 1 - f[..., 2])
                out = out + vol[idx(p0[..., 2] + dz, d), idx(p0[..., 1] + dy, h), idx(p0[..., 0] + dx, w)] * wgt[..., None]
    return out


MAX_DISTANCE = 30.0


def scene(seed, res, moving):
    rng = np.random.default_rng(seed)
    yaw = 0.3
    fwd = np.array([np.sin(yaw), 0.0, -np.cos(yaw)])
    up = np.array([0.0, -1.0, 0.0])                      # the renderer's camera up points down (y flip)
    right = np.cross(fwd, up)
    aspect = res[0] / res[1]
    cam = dict(position=np.array([1.0, -1.5, 2.0]), forward=fwd, up=up, right=right, tan_fov_half=np.tan(np.radians(35.0)), aspect=aspect)
    pyaw = yaw + (0.06 if moving else 0.0)
    pf = np.array([np.sin(pyaw), 0.0, -np.cos(pyaw)])
    pr = np.cross(pf, up)
    ppos = cam["position"] + (np.array([0.4, 0.05, -0.3]) if moving else 0.0)
    # view matrix of the previous camera (rows: right, -up (the projection flips y back), -forward), then the reference's projection
    R = np.stack([pr, -up, -pf])
    view = np.eye(4)
    view[:3, :3], view[:3, 3] = R, -R @ ppos
    prev = dict(view_projection=perspective(70.0, aspect, 0.1, 300.0) @ view, position=ppos, forward=pf)
    n = 8
    noise = rng.integers(0, 256, (n, n, n), dtype=np.uint8)
    s = 32
    shadow = np.where(np.add.outer(np.arange(s), np.arange(s)) % 11 < 5, 40000, 9000).astype(np.uint16)   # plateaus: occluder depth 0.61 / 0.14
    # an orthographic light looking down -y over a 40 m box: x, z -> [-1, 1], y -> depth [0, 1] (column-major)
    L = np.array([[1 / 20.0, 0, 0, 0], [0, 0, 1 / 20.0, 0], [0, 1 / 20.0, 0, 0.5], [0, 0, 0, 1]], np.float64)
    settings = np.array([0.3, -0.2, 0.7, rng.uniform(-0.5, 0.5), 0.7, 1.0, 1.3, MAX_DISTANCE, 0.3, 0.6, 0.9, 1.0, 0.2], np.float32)
    light = np.array([1.0, 0.9, 0.7, 0.123, 0.8], np.float32)   # lightBuffer.inc:4-8: sunColor, previousFrameExposure, sunStrengthExposed
    history = (rng.random((res[2], res[1], res[0], 4)) * np.array([0.5, 0.5, 0.5, 1.0])).astype(np.float16)
    sun = normalize(np.array([0.3, -0.8, 0.5]))
    return cam, prev, noise, shadow, L, settings, light, history, sun


def close(got, want, frac=0.0, what=""):
    got = got.astype(np.float64)
    tol = 2.0 ** -10 * 1.5 * np.maximum(np.abs(want), 2.0 ** -14) + 1e-7
    bad = np.abs(got - want) > tol
    assert bad.mean() <= frac, "%s: %d of %d values differ, worst %.3e" % (what, int(bad.sum()), bad.size, float((np.abs(got - want) / tol).max()))


@pytest.mark.parametrize("res,moving,cut", [((12, 7, 16), False, False), ((10, 6, 8), True, False), ((9, 5, 8), True, True)])
def test_froxel_passes_match_numpy(ffi, oracle, res, moving, cut):
    cam, prev, noise, shadow, L, settings, light, history, sun = scene(res[0] * 10 + res[2], res, moving)
    material, scatter, reprojected, integrated, g = passes.froxels(ffi, oracle, res, noise, shadow, L.T.ravel(), light, settings, history, cam, prev, sun, camera_cut=cut)
    wind, offset, scat, absorb, base, rng_, g_phase = settings[0:3].astype(np.float64), float(settings[3]), settings[4:7].astype(np.float64), float(settings[8]), float(settings[9]), float(settings[10]), float(settings[12])
    w, h, d = res

    # ---- froxelVolumeMaterial.comp:19-43 ----
    pos, V = froxel_world(res, offset, cam)
    n = trilinear(noise.astype(np.float64)[..., None] / 255.0, pos * 0.5 + wind, repeat=True)[..., 0]
    density = np.maximum(base + rng_ * (n - 0.5), 0)
    close(material, np.concatenate([scat * density[..., None], (absorb * density)[..., None]], -1), 0.0, "material")
    assert density.min() == 0 or density.min() > 0  # (both branches of the max are legal)

    # ---- froxelLightScattering.comp:31-64 from the oracle's material volume ----
    pl = np.concatenate([pos, np.ones(pos.shape[:-1] + (1,))], -1) @ L.T
    pl = pl / pl[..., 3:4]
    su, sv = pl[..., 0] * 0.5 + 0.5, pl[..., 1] * 0.5 + 0.5
    actual = np.clip(pl[..., 2], 0, 1)
    s = shadow.shape[0]
    tx, ty = np.floor(su * s).astype(np.int64), np.floor(sv * s).astype(np.int64)
    inside = (tx >= 0) & (tx < s) & (ty >= 0) & (ty < s)
    occluder = np.where(inside, shadow[np.clip(ty, 0, s - 1), np.clip(tx, 0, s - 1)] / 65535.0, 0.0)  # g_sampler_nearestBlackBorder
    lit = (actual > occluder).astype(np.float64)
    assert 0.1 < lit.mean() < 0.95
    VoL = -V @ sun
    phase = (1 - g_phase ** 2) / (4 * np.pi * (1 + g_phase ** 2 - 2 * g_phase * VoL) ** 1.5)
    m = material.astype(np.float64)
    inscatter = ((lit * light[4] * phase)[..., None] * light[0:3].astype(np.float64) + 0.02) * m[..., :3]
    extinction = (m[..., :3] + m[..., 3:4]) @ np.array([0.21, 0.72, 0.07])
    # a froxel whose light-space depth or texel coordinate is within rounding of a shadow-map step may fall on the other side
    close(scatter, np.concatenate([inscatter, extinction[..., None]], -1), 0.01, "scattering")

    # ---- volumeLightingReprojection.comp:19-62 from the oracle's scattering volume ----
    pos0, _ = froxel_world(res, 0.0, cam)
    ndc_prev = np.concatenate([pos0, np.ones(pos0.shape[:-1] + (1,))], -1) @ prev["view_projection"].T
    ndc_prev = ndc_prev[..., :3] / ndc_prev[..., 3:4]
    V_hist = normalize(prev["position"] - pos0)
    hist_depth = np.linalg.norm(pos0 - prev["position"], axis=-1) * (-V_hist @ prev["forward"])
    huv = np.stack([ndc_prev[..., 0] * 0.5 + 0.5, ndc_prev[..., 1] * 0.5 + 0.5, depth_to_uvz(hist_depth, MAX_DISTANCE)], -1)
    outside = (huv > 1).any(-1) | (huv < 0).any(-1)
    alpha = np.where(outside, 0.0, 0.95)[..., None]
    current = scatter.astype(np.float64)
    hist = current if cut else trilinear(history, np.nan_to_num(huv), repeat=False)
    want = current * (1 - alpha) + hist * alpha
    edge = (np.abs(huv - np.round(huv)) < 1e-5).any(-1)  # the off-frustum test within rounding of 0 or 1
    close(reprojected[~edge], want[~edge], 0.0, "reprojection")
    if moving and not cut:
        assert outside.any() and not outside.all()
    if not moving:
        assert not outside.any()

    # ---- volumetricLightingIntegration.comp:18-43 from the oracle's reprojected volume ----
    r = reprojected.astype(np.float64)
    total, transmittance = np.zeros((h, w, 3)), np.ones((h, w))
    want = np.zeros((d, h, w, 4))
    for z in range(d):  # the shader's loop also runs z = res.z: its fetch and its store are both outside the volume
        seg = uvz_to_depth((z + 1) / d, MAX_DISTANCE) - uvz_to_depth(z / d, MAX_DISTANCE)
        ins, ext = r[z, ..., :3], r[z, ..., 3]
        total = total + (ins - ins * np.exp(-ext * seg)[..., None]) / np.maximum(ext, 0.00001)[..., None]   # volumeShading.inc:24-26 (no transmittance so far: as the reference)
        transmittance = transmittance * np.exp(-ext * seg)
        want[z] = np.concatenate([total, transmittance[..., None]], -1)
    # inscattering - inscattering * exp(-x) cancels for thin media: the oracle's binary32 difference carries ~1e-7 / x relative error
    got = integrated.astype(np.float64)
    assert np.abs(got[..., 3] - want[..., 3]).max() <= 2.0 ** -10 * 1.5
    rel = np.abs(got[..., :3] - want[..., :3]) / np.maximum(np.abs(want[..., :3]), 1e-6)
    assert rel.max() < 5e-3 and np.median(rel) < 2.0 ** -10
