"""The CUDA rasteriser (csrc/passes_raster.cu) against the oracle's scalar rasteriser, bit for bit, through the
graphic-pass entry points of the C-ABI: depth prepass (depth / motion / normal), G-buffer fill, shadow cascades."""
import numpy as np
import pytest

import passes
from test_raster_oracle import IDENTITY, mats, perspective, quad

pytestmark = pytest.mark.gpu


def soup(ffi, rng, n_tris, big=0, z_range=(-40.0, 3.0), spread=12.0, size=1.5):
    """random triangles in view space (camera at the origin looking down -z): many small ones, `big` wall-sized ones; some cross
    the near plane, some lie behind the camera, some are degenerate"""
    centre = np.stack([rng.uniform(-spread, spread, n_tris), rng.uniform(-spread * 0.6, spread * 0.6, n_tris), rng.uniform(*z_range, n_tris)], -1)
    ext = np.full(n_tris, size)
    ext[:big] = 25.0
    pos = centre[:, None, :] + rng.normal(0, 1, (n_tris, 3, 3)) * ext[:, None, None]
    pos[-1, 2] = pos[-1, 1]  # a degenerate triangle
    pos = pos.reshape(-1, 3).astype(np.float32)
    nrm = rng.normal(0, 1, (n_tris * 3, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tan = np.cross(nrm, rng.normal(0, 1, (n_tris * 3, 3)))
    tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    bit = np.cross(nrm, tan)
    rng.permutation(n_tris * 3)  # (keeps the random stream of the committed GPU runs)
    idx = np.arange(n_tris * 3).reshape(n_tris, 3)[rng.permutation(n_tris)].reshape(-1)  # triangles share nothing; their order is shuffled
    return idx, ffi.pack_vertices(pos, uvs=rng.uniform(-2, 3, (n_tris * 3, 2)), normals=nrm, tangents=tan, bitangents=bit)


def random_textures(rng):
    return [(8, 4, rng.integers(0, 256, 8 * 4 * 4).astype(np.uint8)), (4, 4, rng.integers(0, 256, 4 * 4 * 4).astype(np.uint8)), (2, 8, rng.integers(0, 256, 2 * 8 * 4).astype(np.uint8)),
            (1, 1, np.array([128, 128, 255, 255], np.uint8))]


@pytest.mark.parametrize("w,h,n_tris,big", [(64, 48, 40, 2), (320, 200, 600, 6), (333, 130, 300, 3), (1920, 1080, 3000, 12)])
def test_prepass_and_gbuffer_bit_exact(ffi, cuda, oracle, w, h, n_tris, big):
    rng = np.random.default_rng(w * 31 + h)
    P = perspective(50.0, w / h, 0.1, 300.0)
    shift = np.eye(4)
    shift[0, 3], shift[2, 3] = 0.3, 0.2
    model = np.eye(4)
    model[:3, :3] = np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1.0]]) * 1.25
    meshes = [soup(ffi, rng, n_tris, big), soup(ffi, rng, max(n_tris // 4, 4), 1)]
    m = np.concatenate([mats(model=np.eye(4), mvp=P, mvp_prev=P @ shift), mats(model=model, mvp=P @ model, mvp_prev=P @ shift @ model)])
    draws = [(0, 0, 0, 1, 2), (1, 1, 2, 3, 0), (0, 1, 1, 3, 2)]
    kw = dict(jitter=((0.5 / w, -0.25 / h), (-0.125 / w, 0.375 / h)), textures=random_textures(rng), gbuffer=True)
    got = passes.raster_prepass(ffi, cuda, w, h, meshes, draws, m, **kw)
    want = passes.raster_prepass(ffi, oracle, w, h, meshes, draws, m, **kw)
    assert (want[0] > 0).mean() > 0.02, "the test scene should cover part of the frame (the random-alpha albedo textures cut holes)"
    for name, a, b in zip(("depth", "motion", "normal", "gbuffer"), got, want):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "%s: %d texels differ" % (name, int((a != b).reshape(h * w, -1).any(-1).sum()))


def test_quads_and_rules_bit_exact(ffi, cuda, oracle):
    w, h = 32, 16
    for mesh_list, draws in (([quad(ffi, 8.5, 24.5, 4.5, 12.5, w, h)], [(0, 0)]),
                             ([quad(ffi, 0, 11.25, 0, h, w, h, z=0.5), quad(ffi, 11.25, w, 0, h, w, h, z=0.25)], [(0, 0), (1, 0)]),
                             ([quad(ffi, 0, 12, 0, 8, w, h, z=0.5), quad(ffi, 4, 16, 0, 8, w, h, z=0.5)], [(1, 0), (0, 0)]),   # equal depth: the later draw wins
                             ([quad(ffi, 8, 24, 4, 12, w, h, front=False)], [(0, 0)]), ([], [])):
        got = passes.raster_prepass(ffi, cuda, w, h, mesh_list, draws, mats(), gbuffer=True)
        want = passes.raster_prepass(ffi, oracle, w, h, mesh_list, draws, mats(), gbuffer=True)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("size,n_tris", [(64, 30), (512, 400), (2048, 1500)])
def test_shadow_cascade_bit_exact(ffi, cuda, oracle, size, n_tris):
    rng = np.random.default_rng(size)
    meshes = [soup(ffi, rng, n_tris, big=3, z_range=(-1.2, 1.2), spread=1.0, size=0.12)]
    lm = np.tile(IDENTITY, (4, 1)).astype(np.float32)
    ortho = np.diag([0.9, 1.1, 0.45, 1.0])
    ortho[2, 3] = 0.5
    lm[2] = ortho.astype(np.float32).T.ravel()
    model = np.eye(4, dtype=np.float32)
    model[0, 3] = 0.1
    cutout = (4, 2, rng.integers(0, 256, 4 * 2 * 4).astype(np.uint8))  # random alpha: the alpha test of sunShadow.frag
    for cascade in (0, 2):
        albedo = cutout if cascade == 2 else None
        got = passes.raster_shadow(ffi, cuda, size, meshes, [(0, 0), (0, 1)], [IDENTITY, model.T.ravel()], lm, cascade=cascade, albedo=albedo)
        want = passes.raster_shadow(ffi, oracle, size, meshes, [(0, 0), (0, 1)], [IDENTITY, model.T.ravel()], lm, cascade=cascade, albedo=albedo)
        assert (want > 0).mean() > 0.01
        assert np.array_equal(got, want), "cascade %d: %d texels differ" % (cascade, int((got != want).sum()))


@pytest.mark.parametrize("w,h,moving,settings", [(192, 108, True, {}), (250, 142, False, {}), (160, 90, True, dict(sdf_debug_mode=1)), (128, 72, True, dict(sun_shadow_cascade_count=4, taa_enabled=0))])
def test_frames_from_plain_meshes_bit_exact(ffi, cuda, oracle, w, h, moving, settings):
    """frames rendered end to end from `.plain` geometry (raster_inputs = 1): every rasterised input and every resource of the
    frame path after it, CUDA against the oracle"""
    from conftest import PlainSceneSequence, ROOT, assert_snapshots_equal
    from plainrenderer_b200 import assets
    a = PlainSceneSequence(ffi, cuda, assets.Assets(), w, h, **settings)
    b = PlainSceneSequence(ffi, oracle, assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_"), w, h, **settings)
    try:
        for f in range(3):
            a.step(moving=moving)
            b.step(moving=moving)
            assert_snapshots_equal(a.snapshot(), b.snapshot(), "frame %d of %dx%d from .plain meshes %s" % (f, w, h, settings))
    finally:
        a.close()
        b.close()


def test_raster_graph_replay_is_identical(ffi, cuda):
    """the rasteriser's launches inside the captured pass list (CUDA graph) produce the same bytes as direct launches"""
    from conftest import PlainSceneSequence
    from plainrenderer_b200 import assets
    outs = []
    for graph in (False, True):
        s = PlainSceneSequence(ffi, cuda, assets.Assets(), 192, 108)
        s.fe.backend.set_graph_replay_enabled(graph)
        for _ in range(4):
            s.step(moving=True)
        outs.append(s.snapshot())
        s.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k
