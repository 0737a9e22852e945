"""Row-sharded frames on the GPU: R frontends (one per rank) on ONE device with the exchanges done by local copies
(sharding.LocalComm) must reproduce the unsharded frame bit for bit - validates the row windows of every kernel, the
overlapped computation and the exchange schedule. The same schedule runs over NCCL in bench.py --gpus N."""
import numpy as np
import pytest

from conftest import CAMERA, image_mips

pytestmark = pytest.mark.gpu

GATHERED_IMAGES = ["hiz", "depthHalf", "giHistY1", "giHistC1", "giHistY0", "giHistC0", "skyLut", "brdfLut"]
BANDED_IMAGES = [("hiz", 2, 0), ("hiz", 2, 1), ("hiz", 2, 2), ("output", 1, 0), ("post1", 1, 0), ("giFullY", 1, 0), ("giFullC", 1, 0), ("giY1", 2, 0), ("giC1", 2, 0),
                 ("giY0", 2, 0), ("giC0", 2, 0), ("color0", 1, 0), ("color1", 1, 0)]  # giY0/giC0: the temporal filter overwrites the gathered trace result with its banded target  # (name, divisor, mip): compared on the rank's own rows


def make(ffi, api, W, H, instances, rank=0, count=1, pass_fusion=False, **settings):
    s = ffi.default_settings(api, W, H, sun_direction_deg=(40.0, 35.0), shard_rank=rank, shard_count=count, **settings)
    fe = ffi.Frontend(api, s)
    fe.backend._check(api.b["set_pass_fusion_enabled"](fe.backend.ctx, 1 if pass_fusion else 0), "set_pass_fusion_enabled")  # off: giFullY / giFullC are compared
    scene = ffi.SyntheticScene(api, n_instances=instances)
    scene.attach(fe)
    fe.set_exposure(2e-5)
    return s, fe, scene


def camera(ffi, f, moving):
    p, fw, r, u = CAMERA
    if moving:
        p = (p[0] + 0.02 * f, p[1], p[2] + 0.01 * f)  # about a pixel per frame at this resolution: inside the TAA history halo
    return ffi.camera(p, fw, r, u)


@pytest.mark.parametrize("W,H,R,moving,fused", [(256, 192, 2, False, False), (256, 192, 3, True, False), (320, 256, 4, True, False), (320, 256, 4, True, True)])
def test_sharded_frame_equals_unsharded(ffi, cuda, W, H, R, moving, fused):
    import torch
    from plainrenderer_b200 import sharding
    s0, ref, scene0 = make(ffi, cuda, W, H, 14)  # the unsharded reference always runs every pass
    ranks = [make(ffi, cuda, W, H, 14, rank=r, count=R, pass_fusion=fused) for r in range(R)]
    fes = [x[1] for x in ranks]
    comm = sharding.LocalComm(cuda, H, R, torch.device("cuda", 0))
    prev = None
    for f in range(4):
        cam = camera(ffi, f, moving)
        inputs = scene0.render_inputs(s0, cam, f + 1, prev_cam=prev, shadows=True)
        prev = cam
        ref.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, inputs["depth"], inputs["motion"], inputs["normal"], inputs["gbuffer"], inputs["shadow_maps"])
        # every rank uploads only its band + 16 halo rows of depth / normal / G-buffer
        upload = []
        for r in range(R):
            a, b = sharding.full_res_band(cuda, H, R, r)
            upload.append((max(a - 16, 0), min(b + 16, H)))
        n_exchanges = sharding.run_frame_local(fes, comm, cam, (f + 1) / 60.0, 1 / 60.0, inputs, upload_rows=upload)
        assert n_exchanges == 8  # 5 on the critical path + 3 of next-frame data (deferred over peer exchange; ordinary all-gathers here)
        torch.cuda.synchronize()
        # images every rank holds completely
        for name in GATHERED_IMAGES:
            h = ref.image(name)
            for mip in range(image_mips(ref, h)):
                if name == "hiz" and mip < 3:
                    continue  # pyramid levels 0-2 stay banded (checked below); level 3 is all-gathered, the rest is replicated
                want = ref.backend.read_image(h, mip)
                for r, fe in enumerate(fes):
                    got = fe.backend.read_image(fe.image(name), mip)
                    assert np.array_equal(got, want), "frame %d rank %d: %s mip %d" % (f, r, name, mip)
        for name, size in (("histogram", 512), ("light", 20), ("sunShadowInfo", 304)):
            want = ref.backend.read_storage_buffer(ref.storage_buffer(name), size)
            for r, fe in enumerate(fes):
                assert np.array_equal(fe.backend.read_storage_buffer(fe.storage_buffer(name), size), want), "frame %d rank %d: buffer %s" % (f, r, name)
        # images a rank holds for its own rows
        for name, div, mip in BANDED_IMAGES:
            if fused and name in ("giFullY", "giFullC"):
                continue  # not written when the upscale is folded into the shading kernel; color0 / color1 below are its consumer's output
            h = ref.image(name)
            d = ref.backend.image_description(h)
            rows = d.height >> mip
            want = ref.backend.read_image(h, mip).reshape(rows, -1)
            for r, fe in enumerate(fes):
                a, b = sharding.shard_band(cuda, H, R, r, div << mip, rows)
                got = fe.backend.read_image(fe.image(name), mip).reshape(rows, -1)
                assert np.array_equal(got[a:b], want[a:b]), "frame %d rank %d: %s rows [%d, %d)" % (f, r, name, a, b)
        # froxel volumes: the integrated volume on the rank's band of froxel rows (8 screen rows each), this frame's reprojection
        # result (next frame's history) all-gathered
        want = ref.backend.read_image(ref.image("froxelIntegration"))
        d = ref.backend.image_description(ref.image("froxelIntegration"))
        want = want.reshape(d.depth, d.height, -1)
        for r, fe in enumerate(fes):
            a, b = sharding.shard_band(cuda, H, R, r, 8, d.height)
            got = fe.backend.read_image(fe.image("froxelIntegration")).reshape(d.depth, d.height, -1)
            assert np.array_equal(got[:, a:b], want[:, a:b]), "frame %d rank %d: froxelIntegration rows [%d, %d)" % (f, r, a, b)
        froxel_hist = "froxelHist%d" % ((f + 1) % 2)  # m_volumetricLightingHistory[frameIndex % 2], frameIndex = f + 1
        want = ref.backend.read_image(ref.image(froxel_hist))
        for r, fe in enumerate(fes):
            assert np.array_equal(fe.backend.read_image(fe.image(froxel_hist)), want), "frame %d rank %d: %s" % (f, r, froxel_hist)
        # TAA history of this frame (all-gathered for the next one)
        hist_name = "taaHist%d" % (f % 2)  # written this frame: m_historyBuffers[(frameIndex % 2 + 1) % 2], frameIndex = f + 1
        want = ref.backend.read_image(ref.image(hist_name))
        for r, fe in enumerate(fes):
            assert np.array_equal(fe.backend.read_image(fe.image(hist_name)), want), "frame %d rank %d: %s" % (f, r, hist_name)
    # assemble the frame from the ranks' bands (what a gather to rank 0 would deliver)
    frame = np.zeros((H, W * 4), np.uint8)
    for r, fe in enumerate(fes):
        a, b = sharding.full_res_band(cuda, H, R, r)
        fe.read_output_rows(frame, (a, b))
    assert np.array_equal(frame.ravel(), ref.read_output())
    for _, fe, sc in ranks:
        sc.close()
        fe.close()
    scene0.close()
    ref.close()


def test_config4_8k_256_instances(ffi, cuda):
    """BASELINE configs[4] on one device: 7680x4320, 256 SDF instances (the 100-instances-per-tile cap is active), a half-resolution
    depth pyramid of 12 levels (one more than the reference binds). Size-independent properties of the unsharded frame, and the
    frame split into 2 row bands (LocalComm exchanges) reproduces it bit for bit."""
    import torch
    from plainrenderer_b200 import sharding
    W, H, R = 7680, 4320, 2
    s0, ref, scene0 = make(ffi, cuda, W, H, 256)
    ranks = [make(ffi, cuda, W, H, 256, rank=r, count=R) for r in range(R)]
    fes = [x[1] for x in ranks]
    comm = sharding.LocalComm(cuda, H, R, torch.device("cuda", 0))
    cam = camera(ffi, 0, False)
    inputs = scene0.render_inputs(s0, cam, 1, prev_cam=None, shadows=True)
    for f in range(2):
        ref.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, inputs["depth"], inputs["motion"], inputs["normal"], inputs["gbuffer"], inputs["shadow_maps"])
        upload = []
        for r in range(R):
            a, b = sharding.full_res_band(cuda, H, R, r)
            upload.append((max(a - 16, 0), min(b + 16, H)))
        sharding.run_frame_local(fes, comm, cam, (f + 1) / 60.0, 1 / 60.0, inputs, upload_rows=upload)
    torch.cuda.synchronize()
    be = ref.backend
    hist = be.read_storage_buffer(ref.storage_buffer("histogram"), 512, np.uint32)
    assert int(hist.sum()) == W * H  # 240 x 135 whole tiles: every pixel of the previous frame's colour lands in exactly one bin
    depth = inputs["depth"].reshape(H, W)
    assert image_mips(ref, ref.image("hiz")) == 12
    hiz0 = be.read_image(ref.image("hiz"), 0, np.float32).reshape(H // 2, W // 2, 2)
    d4 = depth.reshape(H // 2, 2, W // 2, 2)
    assert np.array_equal(hiz0[..., 0], np.minimum(np.where(d4 == 0, np.float32(1.0), d4).min(axis=(1, 3)), np.float32(1.0)))
    assert np.array_equal(hiz0[..., 1], d4.max(axis=(1, 3)))
    top = be.read_image(ref.image("hiz"), 11, np.float32)
    assert top.size == 2 and top[1] == depth.max() and top[0] == depth[depth > 0].min()
    # half-res trace: 120 x 68 tiles, indexed with the full-resolution stride of 240 tiles per row (sdfCulling.inc:17-20); unused entries stay 0
    tiles = be.read_storage_buffer(ref.storage_buffer("sdfTiles"), 404 * 240 * 135, np.uint32).reshape(-1, 101)
    assert tiles[:, 0].max() == 100, "the per-tile instance cap (sdfCulling.inc:5) should be reached with 256 instances"
    out = ref.read_output().reshape(H, W, 4)
    assert out[..., :3].mean() > 1 and out[..., :3].std() > 1 and (out[..., 3] == 255).all()  # second frame: the exposure is still adapting (dark, not black)
    frame = np.zeros((H, W * 4), np.uint8)
    for r, fe in enumerate(fes):
        a, b = sharding.full_res_band(cuda, H, R, r)
        fe.read_output_rows(frame, (a, b))
    assert np.array_equal(frame.reshape(H, W, 4), out)
    for r, fe in enumerate(fes):
        assert np.array_equal(fe.backend.read_storage_buffer(fe.storage_buffer("histogram"), 512, np.uint32), hist), "rank %d histogram" % r
        assert np.array_equal(fe.backend.read_storage_buffer(fe.storage_buffer("sunShadowInfo"), 304), be.read_storage_buffer(ref.storage_buffer("sunShadowInfo"), 304))
    for _, fe, sc in ranks:
        sc.close()
        fe.close()
    scene0.close()
    ref.close()


@pytest.mark.parametrize("W,H,R,moving", [(256, 192, 2, False), (320, 256, 3, True)])
def test_sharded_frame_from_meshes_equals_unsharded(ffi, cuda, W, H, R, moving):
    """SURVEY.md 8f N3 + 8e: with rasterised inputs (raster_inputs = 1) a rank rasterises its band + 16 rows of depth / normal / motion /
    G-buffer, the motion vectors are all-gathered (an 11th exchange), the shadow cascades are rendered by every rank: bit for bit the
    unsharded frame"""
    import torch
    from conftest import PlainSceneSequence
    from plainrenderer_b200 import assets, sharding
    lib = assets.Assets()
    ref = PlainSceneSequence(ffi, cuda, lib, W, H)
    ranks = [PlainSceneSequence(ffi, cuda, lib, W, H, shard_rank=r, shard_count=R) for r in range(R)]
    fes = [x.fe for x in ranks]
    comm = sharding.LocalComm(cuda, H, R, torch.device("cuda", 0))
    for f in range(4):
        cam = ref.camera_at(f, moving, speed=0.05)  # about a pixel per frame: inside the TAA history halo, as in the uploaded-input tests
        ref.fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0)
        assert sharding.run_frame_local(fes, comm, cam, (f + 1) / 60.0, 1 / 60.0) == 9
        torch.cuda.synchronize()
        cur = (f + 1) % 2  # m_sceneRenderTargetIndex after frame f + 1
        for name in ["motion%d" % ((f + 1) % 3), "shadow0", "shadow1", "shadow2", "hiz"]:  # held completely by every rank
            h = ref.fe.image(name)
            for mip in range(image_mips(ref.fe, h)):
                if name == "hiz" and mip < 3:
                    continue
                want = ref.fe.backend.read_image(h, mip)
                for r, fe in enumerate(fes):
                    assert np.array_equal(fe.backend.read_image(fe.image(name), mip), want), "frame %d rank %d: %s mip %d" % (f, r, name, mip)
        for name, size in (("histogram", 512), ("light", 20), ("sunShadowInfo", 304)):
            want = ref.fe.backend.read_storage_buffer(ref.fe.storage_buffer(name), size)
            for r, fe in enumerate(fes):
                assert np.array_equal(fe.backend.read_storage_buffer(fe.storage_buffer(name), size), want), "frame %d rank %d: buffer %s" % (f, r, name)
        for name in ["depth%d" % cur, "normal", "gbuffer", "color%d" % cur, "post1", "giFullY", "output"]:  # the rank's own rows
            h = ref.fe.image(name)
            want = ref.fe.backend.read_image(h, 0).reshape(H, -1)
            for r, fe in enumerate(fes):
                a, b = sharding.full_res_band(cuda, H, R, r)
                got = fe.backend.read_image(fe.image(name), 0).reshape(H, -1)
                assert np.array_equal(got[a:b], want[a:b]), "frame %d rank %d: %s rows [%d, %d)" % (f, r, name, a, b)
    frame = np.zeros((H, W * 4), np.uint8)
    for r, fe in enumerate(fes):
        a, b = sharding.full_res_band(cuda, H, R, r)
        fe.read_output_rows(frame, (a, b))
    assert np.array_equal(frame.ravel(), ref.fe.read_output())
    for x in ranks:
        x.close()
    ref.close()


def test_overlapped_rows_equal_the_owners_rows(ffi, cuda):
    """The deferred exchanges (round 2) let a peer's rows land while this rank's own producer may still be writing its overlapped rows
    (TAA history +-4, second spatial GI filter +-2, froxel reprojection +-5 froxel rows): whoever writes last must write the same bits.
    Checked at every such exchange BEFORE the copy: a rank's rows just outside its band equal the owning rank's rows."""
    import torch
    from plainrenderer_b200 import sharding
    W, H, R = 320, 256, 4
    overlap = {"taaHistory": 4, "giSpatial1": 2, "froxelHistory": 5}
    checked = []

    class CheckingComm(sharding.LocalComm):
        def exchange_all(self, xs):
            name = xs[0].name.decode()
            if name in overlap:
                torch.cuda.synchronize()
                views = [sharding.exchange_views(x, self.device) for x in xs]
                for i in range(xs[0].n_images):
                    rows, div = views[0][i][1], views[0][i][2]
                    bands = [sharding.shard_band(self.api, self.H, self.world, r, div, rows) for r in range(self.world)]
                    for me in range(self.world):
                        a, b = bands[me]
                        for lo, hi in ((max(a - overlap[name], 0), a), (b, min(b + overlap[name], rows))):
                            for y in range(lo, hi):
                                owner = [r for r in range(self.world) if bands[r][0] <= y < bands[r][1]][0]
                                assert torch.equal(views[me][i][0][y], views[owner][i][0][y]), "%s image %d: row %d computed by rank %d differs from its owner %d" % (name, i, y, me, owner)
                                checked.append(name)
            super().exchange_all(xs)

    s0, ref, scene0 = make(ffi, cuda, W, H, 14)
    ranks = [make(ffi, cuda, W, H, 14, rank=r, count=R) for r in range(R)]
    fes = [x[1] for x in ranks]
    comm = CheckingComm(cuda, H, R, torch.device("cuda", 0))
    prev = None
    for f in range(4):
        cam = camera(ffi, f, True)
        inputs = scene0.render_inputs(s0, cam, f + 1, prev_cam=prev, shadows=True)
        prev = cam
        upload = []
        for r in range(R):
            a, b = sharding.full_res_band(cuda, H, R, r)
            upload.append((max(a - 16, 0), min(b + 16, H)))
        sharding.run_frame_local(fes, comm, cam, (f + 1) / 60.0, 1 / 60.0, inputs, upload_rows=upload)
    assert set(checked) == set(overlap)
    for _, fe, scene in [(s0, ref, scene0)] + ranks:
        scene.close()
        fe.close()
