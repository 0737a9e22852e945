"""sdfCameraFrustumCulling.comp and sdfCameraTileCulling.comp (S3) of the oracle against float64 numpy restatements written from the GLSL:
bounding sphere (largest half extent + influence radius) against the six frustum planes; per 32x32 tile the cone around the centre view
vector (radius per metre from the two extreme view vectors), the projection clamped to the tile's HiZ depth range scaled by
dot(cameraToPixel, forward), the reference's quirks (view vectors of half-res tile pixels normalised by the FULL resolution, tile index
strided by the full-resolution width) and the 100-instance cap. The frustum-culled list is in ascending instance order (this build's
definition of the reference's atomic append order)."""
import numpy as np
import pytest

import passes


def normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def camera_and_frustum(near=0.1, far=300.0):
    yaw = 0.4
    fwd = normalize(np.array([np.sin(yaw), 0.08, -np.cos(yaw)]))
    right = normalize(np.cross(fwd, [0, -1.0, 0]))
    up = np.cross(right, fwd)
    cam = dict(position=np.array([2.0, -1.5, 3.0]), forward=fwd, up=up, right=right, tan_fov_half=0.45, aspect=16 / 9)
    # the six planes as (point, outward normal): top, bottom, near, far, left, right (SDFGI.cpp:543-553)
    t, a = cam["tan_fov_half"], cam["aspect"]
    corner = lambda sx, sy, d: cam["position"] + (fwd + sy * t * up + sx * t * a * right) * d
    planes = []
    for sy in (1, -1):                                                   # top / bottom contain the apex and two far corners
        n = np.cross(corner(-1, sy, far) - cam["position"], corner(1, sy, far) - cam["position"]) * sy
        planes.append((corner(-1, sy, far), normalize(n)))
    planes.append((cam["position"] + fwd * near, -fwd))
    planes.append((cam["position"] + fwd * far, fwd))
    for sx in (-1, 1):
        n = np.cross(corner(sx, 1, far) - cam["position"], corner(sx, -1, far) - cam["position"]) * sx
        planes.append((corner(sx, -1, far), normalize(n)))
    pts, nrm = np.array([p for p, _ in planes]), np.array([n for _, n in planes])
    # outward: the point 10 m ahead of the camera is inside every plane
    inside = ((cam["position"] + fwd * 10 - pts) * nrm).sum(-1)
    nrm = np.where((inside > 0)[:, None], -nrm, nrm)
    return cam, pts, nrm


def np_frustum_cull(bbs, pts, nrm, influence):   # sdfCameraFrustumCulling.comp:36-62
    centre = (bbs[:, 1] + bbs[:, 0]) * 0.5
    radius = (bbs[:, 1] - bbs[:, 0]).max(-1) * 0.5 + influence
    dist = ((centre[:, None, :] - pts[None]) * nrm[None]).sum(-1)
    return np.nonzero(~(dist > radius[:, None]).any(-1))[0], dist - radius[:, None]


def np_tile_cull(bbs, culled, cam, influence, target, screen, hiz, near, far):   # sdfCameraTileCulling.comp:42-99
    tx, ty = (target[0] + 31) // 32, (target[1] + 31) // 32
    fwd = cam["forward"]

    def view(px, py):   # VFromiUV: normalised by g_screenResolution, the FULL resolution
        pc = (np.array([px / screen[0], py / screen[1]]) - 0.5) * 2
        return normalize(-fwd + cam["tan_fov_half"] * pc[1] * cam["up"] - cam["tan_fov_half"] * cam["aspect"] * pc[0] * cam["right"])
    lin = lambda d: near * far / (far + (1 - d) * (near - far))
    out, margins = {}, {}
    for j in range(ty):
        for i in range(tx):
            c2p = -view(i * 32 + 16, j * 32 + 16)
            v_ll, v_ur = -view(i * 32, j * 32), -view(i * 32 + 32, j * 32 + 32)
            v_ll, v_ur = v_ll / (c2p @ v_ll), v_ur / (c2p @ v_ur)
            cone = np.linalg.norm(v_ll - v_ur) * 0.5
            d_min, d_max = near, far
            if hiz is not None:
                texel = hiz[min(int(np.floor(j / ty * hiz.shape[0])), hiz.shape[0] - 1), min(int(np.floor(i / tx * hiz.shape[1])), hiz.shape[1] - 1)]
                d_min, d_max = lin(float(texel[1])), lin(float(texel[0]))
            d_min, d_max = d_min * (c2p @ fwd), d_max * (c2p @ fwd)
            lst, mar = [], []
            for k in culled:
                if len(lst) >= 100:
                    break
                centre = (bbs[k, 1] + bbs[k, 0]) * 0.5
                radius = ((bbs[k, 1] - bbs[k, 0]) * 0.5).max() + influence
                proj = np.clip((centre - cam["position"]) @ c2p, d_min, d_max)
                d = np.linalg.norm(centre - (proj * c2p + cam["position"]))
                mar.append(d - (radius + cone * proj))
                if d < radius + cone * proj:
                    lst.append(k)
            out[(i, j)], margins[(i, j)] = lst, mar
    return out, margins


@pytest.mark.parametrize("n,use_hiz,influence", [(40, False, 5.0), (40, True, 5.0), (300, True, 2.0), (260, False, 40.0)])
def test_culling_matches_numpy(ffi, oracle, n, use_hiz, influence):
    rng = np.random.default_rng(n + use_hiz)
    near, far = 0.1, 300.0
    cam, pts, nrm = camera_and_frustum(near, far)
    centres = cam["position"] + rng.uniform(-60, 60, (n, 3)) * np.array([1.0, 0.3, 1.0])
    halfs = rng.uniform(0.3, 4.0, (n, 3))
    bbs = np.stack([centres - halfs, centres + halfs], 1).astype(np.float32).astype(np.float64)
    target, screen = (160, 96), (320, 192)                               # half-resolution trace: 5 x 3 tiles, indexed with a stride of 10
    hiz = None
    if use_hiz:
        lin_near, lin_far = rng.uniform(1.0, 20.0, (3, 5)), rng.uniform(25.0, 120.0, (3, 5))
        to_depth = lambda l: 1 - (near * far / l - far) / (near - far)
        hiz = np.stack([to_depth(lin_far), to_depth(lin_near)], -1).astype(np.float32)   # (min, max) of the reverse-z depth
    culled, tiles, stride = passes.sdf_culling(ffi, oracle, bbs, pts, nrm, influence, target, screen, cam, hiz, near, far)
    want, margin = np_frustum_cull(bbs, pts, nrm, influence)
    decided = np.abs(margin).min(-1) > 1e-3                              # boxes within a millimetre of a plane may fall on either side in binary32
    assert np.array_equal(np.sort(culled), culled)                       # ascending instance order
    assert set(culled[decided[culled]]) == set(want[decided[want]]) and 0 < len(culled) < n
    assert stride == 10
    lists, margins = np_tile_cull(bbs, list(culled), cam, influence, target, screen, hiz, near, far)
    capped = 0
    for (i, j), lst in lists.items():
        rec = tiles[i + j * stride]
        got = list(rec[1:1 + rec[0]])
        if len(lst) >= 100 or rec[0] >= 100:
            capped += 1
            assert rec[0] == 100
        close = {k for k, m in zip(culled, margins[(i, j)]) if abs(m) < 1e-3}
        assert set(got) - close == set(lst) - close, (i, j)
        assert got == sorted(got)                                        # appended in list order
    assert (tiles[[k for k in range(len(tiles)) if (k % stride >= 5 or k // stride >= 3)], 0] == 0xFFFFFFFF).all()   # tiles outside the 5 x 3 grid untouched
    if n == 260:
        assert capped > 0                                                # the 100-instance cap (sdfCulling.inc:5) is reached
    assert len({len(l) for l in lists.values()}) > 1 or n == 260        # tiles see different subsets
