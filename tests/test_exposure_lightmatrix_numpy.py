"""preExposeLights.comp (S11: trimmed-mean auto exposure from the luminance histogram, Call-of-Duty EV offset curve with the reference's
`lightExp - darkOffset` divisor, EV100 clamp >= 10, adaption speed, sun colour from the transmittance LUT) and lightMatrix.comp (S14:
linear cascade splits between the HiZ min / max, orthographic fit of each sub-frustum in light space, last cascade from the near plane
to max(depthMax, 30 m) padded by the trace radius, FLOAT_MIN as the initial maximum) of the oracle against float64 numpy restatements."""
import numpy as np
import pytest

import passes
from conftest import decode_r11g11b10, random_r11g11b10


def np_pre_expose(hist, light, lut, sun_y, screen, lum_min, lum_max, exposure_offset, speed, dt, sun_strength):  # preExposeLights.comp:28-88
    n = len(hist)
    lo, hi = np.log(lum_min), np.log(lum_max)
    pixels = screen[0] * screen[1]
    mean, counted, running = 0.0, 0, 0
    for i in range(n):
        running += int(hist[i])
        pct = np.float32(running) / np.float32(pixels)
        if pct < np.float32(0.95) and pct >= np.float32(0.5):
            mean += int(hist[i]) * np.exp(lo + (hi - lo) * i / (n - 1.0))
            counted += int(hist[i])
    with np.errstate(all="ignore"):
        mean = np.float64(mean) / np.float64(counted)
        scene_ev = np.log2(mean * 100 / 12.5)
        t = np.clip((scene_ev - 2.84) / (12.81 - (-3.17)), 0, 1)        # :35 divides by lightExp - darkOffset
        offset = -3.17 * (1 - t) + 1.47 * t + exposure_offset
        target = max(scene_ev - offset, 10) if not np.isnan(scene_ev) else 10.0   # GLSL max(NaN, 10): the oracle's max drops the NaN
        prev_ev = np.log2(1 / (max(light[3], 0.000001) * 1.2))
        delta = target - prev_ev
        change = np.sign(delta) * min(abs(delta), abs(speed * dt))
        exposure = 1 / (2.0 ** (prev_ev + change) * 1.2)
    lh, lw = lut.shape[:2]
    # bilinear, clamp: u = 0 -> the first column; v = -sun_y / 2 + 1 / 2
    y = np.clip((-sun_y * 0.5 + 0.5) * lh - 0.5, None, None)
    y0 = int(np.floor(y))
    fy = y - y0
    c = lambda j: lut[int(np.clip(j, 0, lh - 1)), 0]
    sun_colour = c(y0) * (1 - fy) + c(y0 + 1) * fy
    return np.concatenate([sun_colour, [exposure, sun_strength * exposure]])


@pytest.mark.parametrize("case", ["bright", "dark", "adapting_up", "adapting_down", "single_bin", "empty_window"])
def test_pre_expose_lights_matches_numpy(ffi, oracle, case):
    rng = np.random.default_rng(len(case))
    screen = (320, 180)
    n = 128
    hist = np.zeros(n, np.uint32)
    centre = {"bright": 100, "dark": 30, "adapting_up": 90, "adapting_down": 60, "single_bin": 80, "empty_window": 70}[case]
    if case == "single_bin":
        hist[centre] = screen[0] * screen[1]       # percentage jumps from 0 to 1: nothing is counted -> mean = 0 / 0
    elif case == "empty_window":
        hist[10], hist[centre] = screen[0] * screen[1] * 0.45, screen[0] * screen[1] * 0.55   # 0.45 -> 1.0: the window [0.5, 0.95) is skipped
    else:
        samples = np.clip(rng.normal(centre, 6, screen[0] * screen[1]).round(), 0, n - 1).astype(int)
        hist = np.bincount(samples, minlength=n).astype(np.uint32)
    prev_exposure = {"adapting_up": 1e-6, "adapting_down": 5e-3}.get(case, 2e-5)
    light = np.array([1, 1, 1, prev_exposure, 1], np.float32)
    lut_p = random_r11g11b10(rng, 16 * 24, finite=True).reshape(24, 16)
    lut = decode_r11g11b10(lut_p)
    sun_y = -0.61
    got = passes.pre_expose_lights(ffi, oracle, hist, light, lut_p, sun_y, screen)
    want = np_pre_expose(hist, light.astype(np.float64), lut, sun_y, screen, 0.001, 200000.0, 1.0, 2.0, 1 / 60.0, 128000.0)
    assert np.allclose(got[:3], want[:3], rtol=2e-6), (got[:3], want[:3])
    if np.isnan(want[3]):
        assert np.isnan(got[3]) and np.isnan(got[4])
    else:
        assert np.allclose(got[3:], want[3:], rtol=2e-5), (case, got[3:], want[3:])   # 2^EV amplifies the log2 rounding of binary32
        prev_ev, ev = np.log2(1 / (prev_exposure * 1.2)), np.log2(1 / (float(got[3]) * 1.2))
        assert abs(ev - prev_ev) <= 2.0 / 60 + 1e-4                                     # at most 2 EV per second
        if case in ("adapting_up", "adapting_down"):
            assert abs(abs(ev - prev_ev) - 2.0 / 60) < 1e-4


def np_light_matrix(depth_min_max, cam, sun, cascades, extra_padding, min_far_plane, near, far):  # lightMatrix.comp:57-138
    lin = lambda d: near * far / (far + (1 - d) * (near - far))
    depth_max, depth_min = lin(depth_min_max[0]), lin(depth_min_max[1])     # reverse z: the smallest depth value is the farthest
    fwd = -np.asarray(sun, np.float64)
    up = np.array([0, -1.0, 0]) if abs(fwd[1]) < 0.9999 else np.array([0, 0, -1.0])
    right = np.cross(fwd, up)
    up = np.cross(right, fwd)
    V = np.eye(4)
    V[0, :3], V[1, :3], V[2, :3] = right / np.linalg.norm(right), up / np.linalg.norm(up), fwd     # rows after the transpose
    correction = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -0.5, 0.5], [0, 0, 0, 1.0]])
    splits = [depth_min + (depth_max - depth_min) * (i + 1) / cascades for i in range(cascades - 1)]
    mats, scales = [], []
    for i in range(cascades):
        lo = depth_min if i == 0 else splits[i - 1]
        hi = splits[i] if i < cascades - 1 else None
        if i == cascades - 1:
            lo, hi = near, max(depth_max, min_far_plane)
        pts = []
        for dist in (hi, lo):
            centre = cam["position"] + cam["forward"] * dist
            hh = cam["tan_fov_half"] * dist
            ww = hh * cam["aspect"]
            pts += [centre + sy * cam["up"] * hh + sx * cam["right"] * ww for sy in (1, -1) for sx in (1, -1)]
        pt = np.array(pts) @ V[:3, :3].T
        min_p, max_p = np.minimum(pt.min(0), 3.402823466e+38), np.maximum(pt.max(0), 1.175494351e-38)   # maxP starts at FLOAT_MIN (> 0)
        if i == cascades - 1:
            min_p, max_p = min_p - extra_padding, max_p + extra_padding
        min_p, max_p = min_p - 0.06, max_p + 0.06                                                      # shadowSampleRadius * 2
        scale = 2 / (max_p - min_p)
        offset = -0.5 * (max_p + min_p) * scale
        P = np.diag([scale[0], scale[1], scale[2], 1.0])
        P[:3, 3] = offset
        mats.append(correction @ P @ V)
        scales.append(scale[:2])
    return splits, mats, scales


@pytest.mark.parametrize("sun,cascades", [((0.35, -0.8, 0.48), 4), ((0.0, -1.0, 0.0), 4), ((-0.6, -0.3, -0.74), 3), ((0.7, 0.1, 0.7), 4)])
def test_light_matrix_matches_numpy(ffi, oracle, sun, cascades):
    sun = np.array(sun) / np.linalg.norm(sun)
    yaw = 0.7
    fwd = np.array([np.sin(yaw), 0.1, -np.cos(yaw)])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0, -1.0, 0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    cam = dict(position=np.array([3.0, -2.0, 1.0]), forward=fwd, up=up, right=right, tan_fov_half=0.41, aspect=16 / 9)
    near, far = 0.1, 300.0
    dmm = np.array([0.0021, 0.083], np.float32)   # reverse z: farthest 0.0021 (~41 m), nearest 0.083 (~1.2 m)
    info = passes.light_matrix(ffi, oracle, dmm, cam, sun, cascades=cascades, near=near, far=far)
    splits, mats, scales = np_light_matrix(dmm.astype(np.float64), cam, sun, cascades, 5.0, 30.0, near, far)
    assert np.allclose(list(info.splits)[:cascades - 1], splits, rtol=1e-4)   # far + (1 - d) * (near - far) cancels: binary32 loses ~1e-5 at 40 m
    for i in range(cascades):
        got = np.array(list(info.lightMatrices[i]), np.float64).reshape(4, 4).T
        assert np.allclose(got, mats[i], rtol=1e-3, atol=1e-4), (i, got, mats[i])
        assert np.allclose(list(info.lightSpaceScale[i]), scales[i], rtol=1e-3)
    # every corner of the camera frustum slice of a cascade lands inside its light-space clip volume
    last = np.array(list(info.lightMatrices[cascades - 1]), np.float64).reshape(4, 4).T
    p = last @ np.append(cam["position"] + cam["forward"] * 20.0, 1.0)
    assert (np.abs(p[:2]) <= 1).all() and 0 <= p[2] <= 1
