"""filterIndirectDiffuseSpatial.comp (the world-space disc filter of the GI denoiser, S4 - the most expensive pass of the frame) of the
oracle against an independent float64 numpy restatement written from the GLSL: the 32-sample xorshift disc shared by every pixel, the
tangent frame from neighbouring depth texels, reprojection with the view-projection matrix, mirrored / shrunk samples at the screen
border, tangent-plane distance weights. The sample sequence is integer arithmetic (exact); texel selection is nearest, so a sample that
lands within rounding distance of a texel border may pick the neighbour - the comparison allows a small fraction of such pixels."""
import numpy as np
import pytest

import passes
from test_raster_oracle import perspective


def wang_hash(seed):
    seed = np.uint32(seed)
    with np.errstate(over="ignore"):
        seed = (seed ^ np.uint32(61)) ^ (seed >> np.uint32(16))
        seed = seed * np.uint32(9)
        seed = seed ^ (seed >> np.uint32(4))
        seed = seed * np.uint32(0x27d4eb2d)
        seed = seed ^ (seed >> np.uint32(15))
    return seed


def rand01(state):
    with np.errstate(over="ignore"):
        state = state ^ (state << np.uint32(13))
        state = state ^ (state >> np.uint32(17))
        state = state ^ (state << np.uint32(5))
    scale = np.frombuffer(np.uint32(0x2f800004).tobytes(), np.float32)[0]
    return state, float(np.clip(np.float32(state) * scale, 0, 1))


def np_spatial(y_sh, co_cg, depth, normal_rgba8, g, filter_index):
    h, w = depth.shape
    near, far = g.nearPlane, g.farPlane
    fwd, up, right = (np.array(list(v)[:3], np.float64) for v in (g.cameraForward, g.cameraUp, g.cameraRight))
    cam = np.array(list(g.cameraPosition)[:3], np.float64)
    VP = np.array(list(g.viewProjection), np.float64).reshape(4, 4).T

    def nearest(img, u, v):
        x = np.clip(np.floor(u * w).astype(int), 0, w - 1)
        y = np.clip(np.floor(v * h).astype(int), 0, h - 1)
        return img[y, x]

    def pixel_to_world(u, v):  # :21-28
        d = nearest(depth, u, v).astype(np.float64)
        lin = near * far / (far + (1 - d) * (near - far))
        ndc_x, ndc_y = u * 2 - 1, v * 2 - 1
        Vd = -fwd + g.cameraTanFovHalf * ndc_y[..., None] * up - g.cameraTanFovHalf * g.cameraAspectRatio * ndc_x[..., None] * right
        to_pixel = -Vd / np.linalg.norm(Vd, axis=-1, keepdims=True)
        return cam + to_pixel / (to_pixel @ fwd)[..., None] * lin[..., None]
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = (xs + 0.5) / w, (ys + 0.5) / h
    p_c, p_r, p_u = pixel_to_world(u, v), pixel_to_world(u + 1 / w, v), pixel_to_world(u, v + 1 / h)
    tangent = (p_c - p_r) / np.linalg.norm(p_c - p_r, axis=-1, keepdims=True)
    bitangent = (p_c - p_u) / np.linalg.norm(p_c - p_u, axis=-1, keepdims=True)
    N = 2 * (nearest(normal_rgba8, u, v)[..., :3].astype(np.float64) / 255.0) - 1
    radius = 1.0 if filter_index == 1 else 1.5
    state = wang_hash(np.uint32(g.frameIndexMod4 + filter_index))
    acc_y, acc_c, total = np.zeros((h, w, 4)), np.zeros((h, w, 2)), np.zeros((h, w))
    length_mod = np.ones((h, w))
    for _ in range(32):
        state, r0 = rand01(state)
        state, r1 = rand01(state)
        d = np.sqrt(r0) * length_mod
        angle = 2 * 3.1415926535 * r1
        ox, oy = np.cos(angle) * d, np.sin(angle) * d
        world = p_c + radius * (ox[..., None] * tangent + oy[..., None] * bitangent)
        clip = np.einsum("ij,hwj->hwi", VP, np.concatenate([world, np.ones((h, w, 1))], -1))
        su, sv = clip[..., 0] / clip[..., 3] * 0.5 + 0.5, clip[..., 1] / clip[..., 3] * 0.5 + 0.5
        su = np.where(su < 0, u - ox, su); sv = np.where(sv < 0, v - oy, sv)
        su = np.where(su > 1, u - ox, su); sv = np.where(sv > 1, v - oy, sv)
        pw = pixel_to_world(su, sv)
        dist = np.abs(((pw - p_c) * N).sum(-1))
        weight = np.clip(0.25 / np.maximum(dist, 0.0001), 0, 1) ** 2
        outside = (su < 0) | (sv < 0) | (su > 1) | (sv > 1)
        weight = np.where(outside, 0.0, weight)
        length_mod = np.where(outside, length_mod * 0.98, length_mod)
        acc_y += weight[..., None] * nearest(y_sh, su, sv).astype(np.float64)
        acc_c += weight[..., None] * nearest(co_cg, su, sv).astype(np.float64)
        total += weight
    total = np.maximum(total, 0.00001)
    return acc_y / total[..., None], acc_c / total[..., None]


@pytest.mark.parametrize("filter_index", [0, 1])
def test_spatial_filter_matches_float64_restatement(ffi, oracle, filter_index):
    rng = np.random.default_rng(60 + filter_index)
    w, h = 64, 40
    # a floor seen by the single-pass rig's camera (forward -z, up (0, -1, 0)): depth grows towards the top of the image, plus two steps
    near, far = 0.1, 300.0
    ys, xs = np.mgrid[0:h, 0:w]
    lin = 4.0 + (h - ys) * 0.35 + np.where(xs > 40, 6.0, 0.0) + np.where((xs > 10) & (xs < 20) & (ys > 20), -2.0, 0.0)
    depth = (1 - (near * far / lin - far) / (near - far)).astype(np.float16)
    nrm = np.zeros((h, w, 4), np.uint8)
    n = np.array([0.1, -0.9, 0.42])
    nrm[..., :3] = np.round((n / np.linalg.norm(n) * 0.5 + 0.5) * 255)
    nrm[:, 41:, :3] = np.round((np.array([-0.7, -0.1, 0.7]) / np.linalg.norm([-0.7, -0.1, 0.7]) * 0.5 + 0.5) * 255)
    y_sh = np.concatenate([rng.uniform(0.1, 2.0, (h, w, 1)), rng.uniform(-0.5, 0.5, (h, w, 3))], -1).astype(np.float16)
    co_cg = rng.uniform(-0.2, 0.2, (h, w, 2)).astype(np.float16)
    cam = np.array([1.0, -2.0, 3.0])
    view = np.eye(4)
    view[:3, 3] = -cam                      # right (1,0,0), up (0,-1,0) -> row signs below, forward (0,0,-1)
    view[1, :] *= -1
    P = perspective(2 * np.degrees(np.arctan(0.3153)), w / h, near, far)
    got_y, got_c, g = passes.gi_spatial_filter(ffi, oracle, y_sh, co_cg, depth, nrm, filter_index, cam, P @ view)
    want_y, want_c = np_spatial(y_sh, co_cg, depth, nrm, g, filter_index)
    err_y = np.abs(got_y.astype(np.float64) - want_y).max(-1) / (np.abs(want_y).max(-1) + 1e-3)
    err_c = np.abs(got_c.astype(np.float64) - want_c).max(-1) / (np.abs(want_c).max(-1) + 1e-3)
    ok = (err_y < 2e-3) & (err_c < 4e-3)     # half-float outputs: 2^-11 relative
    assert ok.mean() > 0.97, "%d of %d pixels differ" % (int((~ok).sum()), ok.size)
    assert np.median(err_y) < 6e-4
    # the filter smooths: the luminance varies less than its input
    assert got_y[..., 0].astype(np.float64).std() < 0.6 * y_sh[..., 0].astype(np.float64).std()
