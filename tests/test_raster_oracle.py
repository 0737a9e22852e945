"""The oracle's rasterisation passes against independent numpy statements of what they must produce:
coverage with the top-left rule, face culling, homogeneous (near-plane crossing) triangles against per-pixel ray casting,
the depth test, motion vectors, the geometric normal and the packed G-buffer. The reference leaves all of this to the Vulkan
rasteriser; oracle/passes_raster.cpp states the rules, these tests pin them."""
import ctypes as C

import numpy as np
import pytest

import passes

IDENTITY = np.eye(4, dtype=np.float32).T.ravel()


def mats(model=None, mvp=None, mvp_prev=None):
    m = np.zeros((1, 3, 16), np.float32)
    m[0, 0] = IDENTITY if model is None else np.asarray(model, np.float32).T.ravel()
    m[0, 1] = IDENTITY if mvp is None else np.asarray(mvp, np.float32).T.ravel()
    m[0, 2] = m[0, 1] if mvp_prev is None else np.asarray(mvp_prev, np.float32).T.ravel()
    return m


def ndc(px, size):
    return 2.0 * px / size - 1.0


def quad(ffi, x0, x1, y0, y1, w, h, z=0.25, front=True):
    """two triangles sharing the diagonal; positions are NDC (mvp = identity), front = counter clockwise on the y-down screen"""
    p = np.array([[ndc(x0, w), ndc(y0, h), z], [ndc(x1, w), ndc(y0, h), z], [ndc(x1, w), ndc(y1, h), z], [ndc(x0, w), ndc(y1, h), z]], np.float32)
    idx = [0, 3, 1, 1, 3, 2] if front else [0, 1, 3, 1, 2, 3]
    n = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    return idx, ffi.pack_vertices(p, uvs=[[0, 0], [1, 0], [1, 1], [0, 1]], normals=n, tangents=np.tile([[1, 0, 0]], (4, 1)), bitangents=np.tile([[0, 1, 0]], (4, 1)))


def test_quad_coverage_culling_and_top_left_rule(ffi, oracle):
    w, h = 32, 16
    depth, motion, normal = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 8, 24, 4, 12, w, h)], [(0, 0)], mats())
    want = np.zeros((h, w), np.float32)
    want[4:12, 8:24] = 0.25
    assert np.array_equal(depth, want)           # every pixel of the quad once, none on the shared diagonal missed
    assert (normal[4:12, 8:24] == [128, 128, 255, 0]).all() and (normal[depth == 0] == 0).all()
    assert (motion == 0).all()
    # clockwise on screen = back face: culled by the prepass (cull back, front face counter clockwise)
    depth_b, _, _ = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 8, 24, 4, 12, w, h, front=False)], [(0, 0)], mats())
    assert (depth_b == 0).all()
    # edges through pixel centres: left / top edges own the pixel, right / bottom edges do not
    depth_h, _, _ = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 8.5, 24.5, 4.5, 12.5, w, h)], [(0, 0)], mats())
    want_h = np.zeros((h, w), np.float32)
    want_h[4:12, 8:24] = 0.25
    assert np.array_equal(depth_h, want_h)
    # two quads tiling the screen edge to edge: no pixel left out along the shared edge, whatever its sub-pixel position
    for split in (11.0, 11.5, 11.25, 10.99609375):
        d2, _, _ = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 0, split, 0, h, w, h, z=0.5), quad(ffi, split, w, 0, h, w, h, z=0.25)], [(0, 0), (1, 0)], mats())
        assert (d2 > 0).all()
        assert (d2[:, : int(np.ceil(split - 0.5))] == 0.5).all() and (d2[:, int(np.ceil(split - 0.5)):] == 0.25).all()


def test_depth_test_greater_equal_in_draw_order(ffi, oracle):
    w, h = 16, 8
    near, far, same = quad(ffi, 0, 12, 0, 8, w, h, z=0.75), quad(ffi, 4, 16, 0, 8, w, h, z=0.5), quad(ffi, 8, 16, 0, 4, w, h, z=0.5)
    for order in ([0, 1, 2], [2, 1, 0], [1, 0, 2]):
        meshes = [near, far, same]
        depth, _, normal = passes.raster_prepass(ffi, oracle, w, h, meshes, [(i, 0) for i in order], mats())
        assert (depth[:, :12] == 0.75).all() and (depth[:, 12:] == 0.5).all()  # reverse z: the larger depth is closer, independent of draw order


def perspective(fov_deg, aspect, near, far):
    """projectionMatrixFromCameraIntrinsic (Camera.cpp:4-27): glm::perspective, y flipped, depth remapped to reverse z in [0, 1]"""
    f = 1.0 / np.tan(np.radians(fov_deg) / 2)
    p = np.array([[f / aspect, 0, 0, 0], [0, f, 0, 0], [0, 0, -(far + near) / (far - near), -2 * far * near / (far - near)], [0, 0, -1, 0]], np.float64)
    c = np.eye(4)
    c[1, 1], c[2, 2], c[2, 3] = -1.0, -0.5, 0.5
    return c @ p


def test_triangle_crossing_the_near_plane_matches_ray_casting(ffi, oracle):
    """a floor triangle that starts behind the camera: coverage and depth must equal a per-pixel ray / plane intersection"""
    w, h = 64, 48
    P = perspective(50.0, w / h, 0.1, 300.0)
    tri = np.array([[-3.0, 1.0, 2.0], [0.5, 1.2, -40.0], [4.0, 1.0, 1.0]], np.float32)  # view space: camera looks down -z, y flipped by the projection
    for order in ([0, 1, 2], [0, 2, 1]):
        idx = order
        nrm = np.tile([[0, 1, 0]], (3, 1))
        depth, motion, normal = passes.raster_prepass(ffi, oracle, w, h, [(idx, ffi.pack_vertices(tri, normals=nrm))], [(0, 0)], mats(mvp=P))
        if not (depth > 0).any():
            continue  # this winding faces away
        a, b, c = tri.astype(np.float64)
        n = np.cross(b - a, c - a)
        ys, xs = np.mgrid[0:h, 0:w]
        nx, ny = (xs + 0.5) / w * 2 - 1, (ys + 0.5) / h * 2 - 1
        Pinv = np.linalg.inv(P)
        far_pt = np.einsum("ij,hwj->hwi", Pinv, np.stack([nx, ny, np.full_like(nx, 0.5), np.ones_like(nx)], -1))
        d = far_pt[..., :3] / far_pt[..., 3:4]  # a point on the pixel's ray; the ray starts at the origin
        t = (a @ n) / (d @ n)
        hit = d * t[..., None]
        bary = np.stack([np.einsum("hwj,j->hw", np.cross(b - hit, c - hit), n), np.einsum("hwj,j->hw", np.cross(c - hit, a - hit), n), np.einsum("hwj,j->hw", np.cross(a - hit, b - hit), n)], -1) / (n @ n)
        clip = np.einsum("ij,hwj->hwi", P, np.concatenate([hit, np.ones((h, w, 1))], -1))
        z = clip[..., 2] / clip[..., 3]
        inside = (bary > 1e-3).all(-1) & (t > 0) & (z > 0) & (z < 1)
        outside = (bary < -1e-3).any(-1) | (t <= 0) | (z > 1 + 1e-6)
        assert inside.sum() > 200
        assert (depth[inside] > 0).all() and (depth[outside] == 0).all()
        assert np.allclose(depth[inside], z[inside], rtol=2e-6, atol=0)
        assert (motion[inside] == 0).all() and (normal[inside][:, :3] == [128, 255, 128]).all()
        return
    raise AssertionError("neither winding rendered")


def test_motion_vectors_and_jitter(ffi, oracle):
    """depthPrepass.frag:33-41: motion = ((ndcPrevious + jitterPrevious) - (ndcCurrent + jitterCurrent)) / 2, RG16_SNORM"""
    w, h = 32, 16
    shift = np.eye(4)
    shift[0, 3], shift[1, 3] = 0.25, -0.125  # the previous frame's matrix moves everything by (0.25, -0.125) NDC
    jitter = ((0.03125, -0.0625), (-0.015625, 0.0078125))
    depth, motion, _ = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 4, 28, 2, 14, w, h)], [(0, 0)], mats(mvp_prev=shift), jitter=jitter)
    want = (np.array([0.25, -0.125]) + np.array(jitter[1]) - np.array(jitter[0])) * 0.5
    got = motion[depth > 0].astype(np.float64) / 32767.0
    assert np.abs(got - want).max() <= 0.5 / 32767.0 + 1e-7


def test_gbuffer_fill_packs_the_inputs_of_the_shading_kernel(ffi, oracle):
    """the packed texel of include/plain_frame_types.h: depth bits, octahedral shading normal, albedo sRGB8 + roughness, metalness"""
    w, h = 32, 16
    albedo = (2, 1, [200, 100, 50, 255, 40, 80, 120, 255])     # two texels: bilinear + repeat across u
    flat_normal = (1, 1, [128, 128, 255, 255])
    specular = (1, 1, [255, 77, 200, 255])
    depth, _, _, gb = passes.raster_prepass(ffi, oracle, w, h, [quad(ffi, 0, w, 0, h, w, h)], [(0, 0, 0, 1, 2)], mats(), textures=[albedo, flat_normal, specular], gbuffer=True)
    assert (gb[..., 0] == depth.view(np.uint32)).all()
    assert ((gb[..., 2] >> 24) == 77).all() and (gb[..., 3] == 200).all()
    r = (gb[..., 2] & 0xFF).astype(np.float64)
    u = (np.arange(w) + 0.5) / w
    fx = u * 2 - 0.5
    wgt = fx - np.floor(fx)
    t0, t1 = np.array([200.0, 40.0])[np.floor(fx).astype(int) % 2], np.array([200.0, 40.0])[(np.floor(fx).astype(int) + 1) % 2]
    assert np.abs(r - (t0 * (1 - wgt) + t1 * wgt)[None, :]).max() <= 0.51
    # a flat normal map through triangle.frag:181's reconstruction leaves (almost) the geometric normal (0, 0, 1): octahedral (0, 0)
    oct_xy = np.stack([(gb[..., 1] & 0xFFFF).astype(np.uint16).view(np.int16), (gb[..., 1] >> 16).astype(np.uint16).view(np.int16)], -1).astype(np.float64) / 32767
    assert np.abs(oct_xy).max() < 0.02


def test_shadow_cascade_keeps_the_back_faces_and_clamps_depth(ffi, oracle):
    """sunShadow: cull front (the faces away from the light are kept), depth clamp on, D16"""
    size = 64
    lm = np.tile(IDENTITY, (4, 1))
    front, back = quad(ffi, 8, 40, 8, 40, size, size, z=0.5, front=True), quad(ffi, 16, 56, 16, 56, size, size, z=0.25, front=False)
    beyond = quad(ffi, 40, 60, 2, 10, size, size, z=1.5, front=False)  # in front of the near plane: clamped to 1, not clipped
    sm = passes.raster_shadow(ffi, oracle, size, [front, back, beyond], [(0, 0), (1, 0), (2, 0)], [IDENTITY], lm)
    want = np.zeros((size, size), np.uint16)
    want[16:56, 16:56] = int(0.25 * 65535 + 0.5)
    want[2:10, 40:60] = 65535
    assert np.array_equal(sm, want)


def test_frame_from_plain_meshes(ffi, oracle):
    """a frame rendered end to end from `.plain` geometry (raster_inputs = 1): what the rasterised inputs must look like"""
    from conftest import PlainSceneSequence, ROOT
    from plainrenderer_b200 import assets
    lib = assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_")
    w, h = 96, 54
    s = PlainSceneSequence(ffi, oracle, lib, w, h)
    for _ in range(3):
        s.step()
    snap = s.snapshot(["depth0", "depth1", "motion0", "motion1", "motion2", "normal", "gbuffer", "shadow2", "output", "giFullY"], [("sunShadowInfo", 304)])
    main_draws, shadow_draws = s.fe.drawcall_counts()
    assert 3 <= main_draws <= shadow_draws <= len(s.PLACEMENT), "camera-frustum culling keeps what is in view, the fitted shadow frustum a superset"
    # an object 500 m towards the sun is not in view but casts into it: kept by the shadow pass only (the near plane of the fitted frustum
    # is pushed 10 km towards the sun, RenderFrontend.cpp:616-623); an object 2 km to the side is culled from both
    fm, m = s.meshes["cube"]
    sun = (C.c_float * 3)()
    oracle.f["host_direction_to_vector"]((C.c_float * 2)(40.0, 35.0), sun)
    up_sun = np.array(list(sun)) * 500.0 + np.array([0.0, 0.0, 0.0])
    def placed(t):
        M = np.eye(4, dtype=np.float32)
        M[:3, 3] = t
        return (fm, M.T.ravel(), np.array(t) + m.bb_min, np.array(t) + m.bb_max)
    extra = [placed(up_sun), placed(np.array([0.0, 0.0, 2000.0]))]
    base = []
    for name, t, sc, rot in s.PLACEMENT:
        f2, m2 = s.meshes[name]
        c, s_ = np.cos(rot), np.sin(rot)
        M = np.array([[c * sc[0], 0, s_ * sc[2], t[0]], [0, sc[1], 0, t[1]], [-s_ * sc[0], 0, c * sc[2], t[2]], [0, 0, 0, 1]], np.float32)
        corners = np.array([[x, y, z, 1] for x in (m2.bb_min[0], m2.bb_max[0]) for y in (m2.bb_min[1], m2.bb_max[1]) for z in (m2.bb_min[2], m2.bb_max[2])], np.float32) @ M.T
        base.append((f2, M.T.ravel(), corners[:, :3].min(0), corners[:, :3].max(0)))
    s.fe.set_scene(base + extra)
    s.step()
    main2, shadow2 = s.fe.drawcall_counts()
    assert main2 == main_draws, "neither extra object is in view"
    assert shadow2 == shadow_draws + 1, "the object towards the sun still casts into the fitted frustum, the one 2 km to the side does not"
    s.close()
    depth = snap["depth1/0"].view(np.float32).reshape(h, w)  # frame 3 renders into target 1
    covered = depth > 0
    assert 0.2 < covered.mean() < 0.99, "the slab floor and the boxes cover part of the view, the sky the rest"
    assert depth.max() < 1.0
    gb = snap["gbuffer/0"].view(np.uint32).reshape(h, w, 4)
    assert np.array_equal(gb[..., 0], depth.view(np.uint32)) and (gb[~covered] == 0).all()
    normal = snap["normal/0"].reshape(h, w, 4)
    n = normal[covered][:, :3].astype(np.float64) / 255 * 2 - 1
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 0.02 and (normal[~covered] == 0).all()
    assert (n[:, 1] < -0.98).mean() > 0.05   # y points down in this world: the top faces' normal is (0, -1, 0)
    assert (n[:, 0] < -0.5).mean() > 0.2     # the camera looks down +x: many faces look back at it
    motions = [snap["motion%d/0" % i].view(np.int16) for i in range(3)]
    assert max(np.abs(m).max() for m in motions) <= 1, "static camera: the jitter cancels (depthPrepass.frag:36-39)"
    shadow = snap["shadow2/0"].view(np.uint16)
    assert 0.001 < (shadow > 0).mean() < 0.9
    out = snap["output/0"].reshape(h, w, 4)
    assert out[..., :3].std() > 2 and (out[..., 3] == 255).all()
    assert np.isfinite(snap["giFullY/0"].view(np.float16).astype(np.float32)).all()


def test_alpha_test_discards_before_the_depth_test(ffi, oracle):
    """depthPrepass.frag:28-31: fragments whose albedo alpha (bilinear, repeat) is below 0.5 are discarded - what lies behind shows"""
    w, h = 32, 16
    cutout = (2, 1, [255, 255, 255, 0, 255, 255, 255, 255])   # left texel transparent, right texel opaque
    opaque = (1, 1, [255, 255, 255, 255])
    front, back = quad(ffi, 0, w, 0, h, w, h, z=0.75), quad(ffi, 0, w, 0, h, w, h, z=0.25)
    depth, _, _, gb = passes.raster_prepass(ffi, oracle, w, h, [front, back], [(0, 0, 0, 1, 1), (1, 0, 1, 1, 1)], mats(), textures=[cutout, opaque], gbuffer=True)
    u = (np.arange(w) + 0.5) / w
    fx = u * 2 - 0.5
    wgt = fx - np.floor(fx)
    a0, a1 = np.array([0.0, 1.0])[np.floor(fx).astype(int) % 2], np.array([0.0, 1.0])[(np.floor(fx).astype(int) + 1) % 2]
    alpha = a0 * (1 - wgt) + a1 * wgt
    sure = np.abs(alpha - 0.5) > 1e-3
    want = np.where(alpha >= 0.5, 0.75, 0.25).astype(np.float32)
    assert (depth[:, sure] == want[None, sure]).all()
    assert 0.3 < (depth == 0.25).mean() < 0.7
    assert np.array_equal(gb[..., 0], depth.view(np.uint32))  # the G-buffer fill resolves the surviving fragments
    # the shadow pass applies the same test (sunShadow.frag:19-22)
    sm = passes.raster_shadow(ffi, oracle, 32, [quad(ffi, 0, 32, 0, 32, 32, 32, z=0.5, front=False)], [(0, 0)], [IDENTITY], np.tile(IDENTITY, (4, 1)), albedo=cutout)
    assert ((sm > 0).mean(axis=0)[sure] == (alpha >= 0.5)[sure]).all()


def test_fan_of_triangles_is_watertight(ffi, oracle):
    """a fan of triangles tiling a convex polygon at arbitrary sub-pixel positions: every pixel centre strictly inside the polygon
    is covered (no cracks along the shared edges), every centre strictly outside is not"""
    w, h = 64, 48
    rng = np.random.default_rng(7)
    for trial in range(4):
        n = 9
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(14, 22, n)
        centre = np.array([32.3, 24.6]) + rng.uniform(-0.5, 0.5, 2)
        ring = centre + np.stack([np.cos(ang), np.sin(ang)], -1) * rad[:, None]      # pixel coordinates, counter clockwise for y up
        pts = np.concatenate([[centre], ring])
        pos = np.stack([ndc(pts[:, 0], w), ndc(pts[:, 1], h), np.full(len(pts), 0.5)], -1).astype(np.float32)
        idx = []
        for k in range(n):
            a, b = 1 + k, 1 + (k + 1) % n
            idx += [0, b, a]  # front faces: counter clockwise on the y-down screen
        depth, _, _ = passes.raster_prepass(ffi, oracle, w, h, [(idx, ffi.pack_vertices(pos, normals=np.tile([[0, 0, 1]], (len(pts), 1))))], [(0, 0)], mats())
        if not (depth > 0).any():
            idx = [i for t in np.array(idx).reshape(-1, 3)[:, ::-1] for i in t]
            depth, _, _ = passes.raster_prepass(ffi, oracle, w, h, [(idx, ffi.pack_vertices(pos, normals=np.tile([[0, 0, 1]], (len(pts), 1))))], [(0, 0)], mats())
        # the snapped polygon (1/256 pixel) is what is rasterised: test against it with a margin of one snap step
        snapped = np.floor(((pos[:, :2].astype(np.float64) * 0.5 + 0.5) * [w, h]) * 256 + 0.5) / 256
        ys, xs = np.mgrid[0:h, 0:w]
        c = np.stack([xs + 0.5, ys + 0.5], -1)
        inside_all = np.zeros((h, w), bool)
        near_edge = np.zeros((h, w), bool)
        for t in np.array(idx).reshape(-1, 3):
            a, b, cc = snapped[t]
            e = lambda p, q: (q[0] - p[0]) * (c[..., 1] - p[1]) - (q[1] - p[1]) * (c[..., 0] - p[0])
            e0, e1, e2 = e(a, b), e(b, cc), e(cc, a)
            s = np.sign((b[0] - a[0]) * (cc[1] - a[1]) - (cc[0] - a[0]) * (b[1] - a[1]))
            inside_all |= (s * e0 >= 0) & (s * e1 >= 0) & (s * e2 >= 0)
        # polygon boundary = ring edges only
        ringp = snapped[1:]
        for k in range(n):
            p, q = ringp[k], ringp[(k + 1) % n]
            d = np.abs((q[0] - p[0]) * (c[..., 1] - p[1]) - (q[1] - p[1]) * (c[..., 0] - p[0])) / np.hypot(*(q - p))
            near_edge |= d < 0.02
        assert ((depth > 0) == inside_all)[~near_edge].all(), "trial %d: %d pixels differ" % (trial, int((((depth > 0) != inside_all) & ~near_edge).sum()))


def test_guard_band_far_plane_and_depth_range(ffi, oracle):
    """a triangle reaching 10^6 pixels off screen is clipped to the guard band without disturbing what is visible; fragments
    behind the far plane (depth < 0) are dropped by the prepass"""
    w, h = 32, 16
    # x from -0.5 NDC to 60000 NDC: the visible part must be exactly the half plane right of x = -0.5, between the two slanted edges
    pos = np.array([[-0.5, -0.75, 0.5], [60000.0, -30000.0, 0.5], [60000.0, 30000.0, 0.5]], np.float32)
    for idx in ([0, 1, 2], [0, 2, 1]):
        depth, _, _ = passes.raster_prepass(ffi, oracle, w, h, [(idx, ffi.pack_vertices(pos, normals=np.tile([[0, 0, 1]], (3, 1))))], [(0, 0)], mats())
        if (depth > 0).any():
            break
    ys, xs = np.mgrid[0:h, 0:w]
    nx, ny = (xs + 0.5) / w * 2 - 1, (ys + 0.5) / h * 2 - 1
    t = (nx + 0.5) / 60000.5
    lo, hi = -0.75 + t * (-30000 + 0.75), -0.75 + t * (30000 + 0.75)
    inside = (nx > -0.5 + 1e-3) & (ny > lo + 1e-3) & (ny < hi - 1e-3)
    outside = (nx < -0.5 - 1e-3) | (ny < lo - 1e-3) | (ny > hi + 1e-3)
    assert (depth[inside] == 0.5).all() and (depth[outside] == 0).all() and inside.sum() > 100
    # depth from 0.5 at the left edge to -0.5 at the right edge: the part behind the far plane (depth < 0) is dropped
    q = quad(ffi, 0, w, 0, h, w, h)
    vtx = np.array(q[1]).copy()
    zs = np.array([0.5, -0.5, -0.5, 0.5], np.float32)  # per corner (x0, x1, x1, x0)
    vtx[:, 8:12] = zs.view(np.uint8).reshape(4, 4)
    depth2, _, _ = passes.raster_prepass(ffi, oracle, w, h, [(q[0], vtx)], [(0, 0)], mats())
    want = 0.5 - (np.arange(w) + 0.5) / w
    assert ((depth2 > 0) == (want > 0)[None, :]).all()
    assert np.allclose(depth2[:, want > 0], want[None, want > 0], atol=1e-6)
