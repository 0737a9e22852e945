"""Parity of the CUDA frame path against the CPU oracle, through the C-ABI, bit-exact (integer passes by nature; the
floating-point passes by the numeric contract of DESIGN.md: same binary32 operation order, -fmad=false, pinned libm).
The north-star tolerance (1e-3 relative L-inf on the tonemapped frame) is implied: the 8-bit frames are identical."""
import numpy as np
import pytest

import passes
from conftest import ALL_BUFFERS, ALL_IMAGES, FUSED_AWAY_IMAGES, N4_VARIANTS, Sequence, assert_snapshots_equal, decode_r11g11b10, image_mips, random_r11g11b10

pytestmark = pytest.mark.gpu


# ---------------- single passes on adversarial inputs ----------------
@pytest.mark.parametrize("w,h", [(64, 36), (50, 30), (37, 23), (130, 66), (2, 2), (18, 5), (257, 129), (512, 256), (4096, 4), (4098, 6), (7680, 64)])  # the last three: 12 levels
def test_hiz_bit_exact(ffi, cuda, oracle, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    depth = rng.uniform(0.0005, 0.9, (h, w)).astype(np.float32)
    depth[rng.uniform(size=(h, w)) < 0.2] = 0.0
    got, want = passes.hiz(ffi, cuda, depth), passes.hiz(ffi, oracle, depth)
    for lvl, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "level %d" % lvl


@pytest.mark.parametrize("w,h", [(96, 70), (100, 64), (64, 34), (33, 33), (320, 180)])
def test_histogram_bit_exact(ffi, cuda, oracle, w, h):
    rng = np.random.default_rng(w + h)
    packed = random_r11g11b10(rng, w * h, finite=False).reshape(h, w)  # includes inf / NaN texels
    packed[:2] = 0
    for exposure in (0.37, 2e-5):
        a, b = passes.histogram(ffi, cuda, packed, exposure), passes.histogram(ffi, oracle, packed, exposure)
        assert np.array_equal(a[1], b[1]), "histogram bins"
        assert np.array_equal(a[0], b[0]), "per-tile histograms"


@pytest.mark.parametrize("w,h", [(96, 40), (130, 67), (5, 3), (1024, 64)])
def test_tonemap_bit_exact(ffi, cuda, oracle, w, h):
    rng = np.random.default_rng(w * h)
    packed = random_r11g11b10(rng, w * h, finite=False).reshape(h, w)
    for time in (0.37, 12.75, 0.0):
        assert np.array_equal(passes.tonemap(ffi, cuda, packed, time), passes.tonemap(ffi, oracle, packed, time))


@pytest.mark.parametrize("w,h", [(64, 48), (100, 75), (33, 17), (256, 130)])
def test_bloom_chain_bit_exact(ffi, cuda, oracle, w, h):
    rng = np.random.default_rng(w ^ h)
    packed = random_r11g11b10(rng, w * h).reshape(h, w)
    a, b = passes.bloom(ffi, cuda, packed), passes.bloom(ffi, oracle, packed)
    for m, (x, y) in enumerate(zip(a[0], b[0])):
        assert np.array_equal(x, y), "downsample mip %d" % (m + 1)
    for m, (x, y) in enumerate(zip(a[1], b[1])):
        assert np.array_equal(x, y), "upsample mip %d" % m
    assert np.array_equal(a[2], b[2])


def test_depth_downscale_bit_exact(ffi, cuda, oracle):
    rng = np.random.default_rng(2)
    depth = rng.uniform(0, 1, (31, 45)).astype(np.float32)
    assert np.array_equal(passes.depth_downscale(ffi, cuda, depth).view(np.uint16), passes.depth_downscale(ffi, oracle, depth).view(np.uint16))


@pytest.mark.parametrize("w,h", [(96, 40), (70, 37), (5, 3), (513, 65)])
def test_color_to_luminance_bit_exact(ffi, cuda, oracle, w, h):
    rng = np.random.default_rng(w * 7 + h)
    packed = random_r11g11b10(rng, w * h, finite=False).reshape(h, w)  # includes inf / NaN texels
    assert np.array_equal(passes.color_to_luminance(ffi, cuda, packed), passes.color_to_luminance(ffi, oracle, packed))


@pytest.mark.parametrize("w,h,use_tonemap", [(96, 40, True), (70, 37, False), (9, 5, True), (260, 130, True)])
def test_temporal_supersampling_bit_exact(ffi, cuda, oracle, w, h, use_tonemap):
    """random colours (inf / NaN included), random motion up to +-1/8 screen with some vectors leaving the screen, depth steps"""
    rng = np.random.default_rng(w * 13 + h)
    kw = dict(current=random_r11g11b10(rng, w * h, finite=False).reshape(h, w), last=random_r11g11b10(rng, w * h, finite=False).reshape(h, w),
              motion=rng.integers(-4096, 4096, (h, w, 2)).astype(np.int16), depth_current=rng.uniform(0.0003, 0.05, (h, w)).astype(np.float32),
              depth_last=rng.uniform(0.0003, 0.05, (h, w)).astype(np.float32), lum_current=rng.integers(0, 256, (h, w)).astype(np.uint8),
              lum_last=rng.integers(0, 256, (h, w)).astype(np.uint8), use_tonemap=use_tonemap)
    kw["depth_current"][rng.uniform(size=(h, w)) < 0.2] = 0.0  # sky
    kw["motion"][rng.uniform(size=(h, w)) < 0.05] = 32767
    kw["lum_last"][: h // 2] = kw["lum_current"][: h // 2]      # upper half passes the contrast test
    kw["depth_last"][:, : w // 2] = kw["depth_current"][:, : w // 2]
    assert np.array_equal(passes.temporal_supersampling(ffi, cuda, **kw), passes.temporal_supersampling(ffi, oracle, **kw))


# ---------------- whole frames: every resource of every pass ----------------
def run_both(ffi, cuda, oracle, w, h, frames, moving, instances=12, **settings):
    """the CUDA product twice - every pass launched (all images compared) and with pass fusion, the product's default (the images the
    fused-away pass would have written are skipped) - against one run of the oracle"""
    a, b = Sequence(ffi, cuda, w, h, instances, **settings), Sequence(ffi, oracle, w, h, instances, **settings)
    fused = Sequence(ffi, cuda, w, h, instances, pass_fusion=True, **settings)
    try:
        for f in range(frames):
            inputs = a.step(moving=moving)
            b.step(moving=moving, inputs=inputs)
            fused.step(moving=moving, inputs=inputs)
            want = b.snapshot()
            assert_snapshots_equal(a.snapshot(), want, "frame %d of %dx%d %s" % (f, w, h, settings))
            assert_snapshots_equal(fused.snapshot(skip=FUSED_AWAY_IMAGES), want, "frame %d of %dx%d %s (pass fusion)" % (f, w, h, settings))
    finally:
        a.close()
        b.close()
        fused.close()


@pytest.mark.parametrize("w,h", [(256, 144), (200, 120), (250, 142)])
def test_frame_sequence_static_camera(ffi, cuda, oracle, w, h):
    run_both(ffi, cuda, oracle, w, h, frames=3, moving=False)


def test_frame_sequence_moving_camera(ffi, cuda, oracle):
    run_both(ffi, cuda, oracle, 224, 126, frames=4, moving=True, instances=30)


def test_frame_many_instances_hits_tile_cap(ffi, cuda, oracle):
    run_both(ffi, cuda, oracle, 160, 90, frames=2, moving=False, instances=160)


@pytest.mark.parametrize("settings", [
    dict(indirect_lighting_tech=1),                                   # BASELINE configs[1]: shade + tonemap with constant ambient
    dict(taa_history_sampling_tech=0), dict(taa_history_sampling_tech=1), dict(taa_history_sampling_tech=2), dict(taa_history_sampling_tech=3),
    dict(taa_use_clipping=0, taa_use_motion_vector_dilation=0, taa_filter_use_tonemapping=0),
    dict(diffuse_brdf=0), dict(diffuse_brdf=1), dict(diffuse_brdf=3), dict(direct_multiscatter=1), dict(direct_multiscatter=2), dict(direct_multiscatter=3),
    dict(use_geometry_aa=0), dict(sun_shadow_cascade_count=4), dict(half_res_trace=0), dict(strict_influence_radius_cutoff=0),
    dict(taa_enabled=0), dict(bloom_enabled=0), dict(sun_direction_deg=(200.0, 80.0)),
])
def test_frame_setting_variants(ffi, cuda, oracle, settings):
    run_both(ffi, cuda, oracle, 128, 72, frames=2, moving=True, instances=8, **settings)


@pytest.mark.parametrize("settings", N4_VARIANTS)
def test_frame_n4_variants(ffi, cuda, oracle, settings):
    """separate temporal supersampling (colorToLuminance.comp + temporalSupersampling.comp) and the SDF debug
    visualiser (sdfDebugVisualisation.comp, primary rays through the SDF scene), every resource of a 3-frame moving sequence."""
    run_both(ffi, cuda, oracle, 160, 90, frames=3, moving=True, instances=14, **settings)


def test_debug_visualisation_many_instances(ffi, cuda, oracle):
    """tiles at the 100-instance cap (camera tile usage shows red) and ties between overlapping bricks"""
    run_both(ffi, cuda, oracle, 128, 72, frames=1, moving=False, instances=160, sdf_debug_mode=2)
    run_both(ffi, cuda, oracle, 128, 72, frames=1, moving=False, instances=160, sdf_debug_mode=4, sdf_debug_use_influence_radius=1)


def test_config1_shade_and_tonemap_1080p(ffi, cuda, oracle):
    """BASELINE configs[1]: 1920x1080 synthetic G-buffer, Cook-Torrance shade + tonemap only (constant ambient, no TAA/bloom)."""
    a = Sequence(ffi, cuda, 1920, 1080, 24, indirect_lighting_tech=1, taa_enabled=0, bloom_enabled=0)
    b = Sequence(ffi, oracle, 1920, 1080, 24, indirect_lighting_tech=1, taa_enabled=0, bloom_enabled=0)
    try:
        inputs = a.step()
        b.step(inputs=inputs)
        assert_snapshots_equal(a.snapshot(["color0", "color1", "output"], [("histogram", 512)]), b.snapshot(["color0", "color1", "output"], [("histogram", 512)]), "1080p shade+tonemap")
    finally:
        a.close()
        b.close()


def test_graph_replay_is_identical(ffi, cuda):
    """Replaying the pass list as a CUDA graph produces the same bytes as launching the passes one by one."""
    outs = []
    for graph in (False, True):
        s = Sequence(ffi, cuda, 192, 108, 10)
        s.fe.backend.set_graph_replay_enabled(graph)
        for _ in range(5):
            s.step(moving=True)
        outs.append(s.snapshot())
        assert s.fe.backend.last_frame_launch_count() > 30
        s.close()
    assert_snapshots_equal(outs[0], outs[1], "graph replay")


def test_concurrent_pass_schedule_is_identical(ffi, cuda):
    """Scheduling the passes onto several streams from their resource hazards produces the same bytes as strict submission order."""
    outs = []
    for concurrent in (False, True):
        s = Sequence(ffi, cuda, 192, 108, 10)
        s.fe.backend._check(cuda.b["set_concurrent_passes_enabled"](s.fe.backend.ctx, 1 if concurrent else 0), "set_concurrent_passes_enabled")
        for _ in range(4):
            s.step(moving=True)
        outs.append(s.snapshot())
        s.close()
    assert_snapshots_equal(outs[0], outs[1], "concurrent pass schedule")


# ---------------- full size (BASELINE 3840x2160) against the oracle ----------------
def test_frame_4k_equals_oracle(ffi, cuda, oracle):
    """BASELINE configs[2] itself: 3840x2160, 100 SDF instances, two frames (the second consumes every history of the first), every image and
    buffer of the frame compared with the oracle bit for bit. The reference's resolution-capped buffers (SDFGI.cpp:146-151,
    RenderFrontend.cpp:1069-1070, sdfCulling.inc:17-20) only matter at this size. ~12 s per oracle frame on 16 host threads."""
    W, H = 3840, 2160
    a, b = Sequence(ffi, cuda, W, H, 100, pass_fusion=True), Sequence(ffi, oracle, W, H, 100)  # as benchmarked: the upscale folded into the shading kernel
    try:
        inputs = None
        for f in range(2):
            inputs = a.step(inputs=inputs) if inputs is not None else a.step()  # static camera: the inputs of frame 0 serve both frames
            b.step(inputs=inputs)
        # one image at a time: a full snapshot of both sides would hold ~4 GB
        bad = []
        for name in ALL_IMAGES:
            if name in FUSED_AWAY_IMAGES:
                continue  # compared unfused at smaller sizes (run_both); their consumer's output (color0 / color1) is compared here
            ha, hb = a.fe.image(name), b.fe.image(name)
            for mip in range(image_mips(a.fe, ha)):
                x, y = a.fe.backend.read_image(ha, mip), b.fe.backend.read_image(hb, mip)
                n = int((x != y).sum())
                if n:
                    bad.append("%s/%d: %d of %d bytes" % (name, mip, n, x.size))
        for name, size in ALL_BUFFERS:
            size = a.buffer_bytes(name) if size is None else size
            x, y = a.fe.backend.read_storage_buffer(a.fe.storage_buffer(name), size), b.fe.backend.read_storage_buffer(b.fe.storage_buffer(name), size)
            if not np.array_equal(x, y):
                bad.append("buf:" + name)
        assert not bad, "3840x2160 frame differs from the oracle: " + "; ".join(bad)
    finally:
        a.close()
        b.close()


# ---------------- full size (BASELINE 3840x2160): size-independent properties ----------------
def test_full_size_properties(ffi, cuda):
    W, H = 3840, 2160
    s = Sequence(ffi, cuda, W, H, 100)
    inputs = None
    for _ in range(3):
        inputs = s.step(inputs=inputs) if inputs is not None else s.step()
    be, fe = s.fe.backend, s.fe
    # histogram of the previous frame's colour: every pixel lands in exactly one bin (width is a multiple of 32, last tile row has 16 rows)
    hist = be.read_storage_buffer(fe.storage_buffer("histogram"), 512, np.uint32)
    assert int(hist.sum()) == W * H
    # numpy histogram of the same image (float64 log): cumulative counts agree up to bin-edge rounding
    prev = be.read_image(fe.image("color0" if s.frame % 2 == 0 else "color1"), 0, np.uint32)  # the frame before the last one
    # HiZ: top level = (min over non-sky, max) of the depth buffer; level 0 = 2x2 reduction
    depth = inputs["depth"].reshape(H, W)
    hiz0 = be.read_image(fe.image("hiz"), 0, np.float32).reshape(H // 2, W // 2, 2)
    d4 = depth.reshape(H // 2, 2, W // 2, 2)
    sky_excluded = np.where(d4 == 0, np.float32(1.0), d4)
    assert np.array_equal(hiz0[..., 0], np.minimum(sky_excluded.min(axis=(1, 3)), np.float32(1.0)))
    assert np.array_equal(hiz0[..., 1], d4.max(axis=(1, 3)))
    top = be.read_image(fe.image("hiz"), 10, np.float32)
    assert top[1] == depth.max()
    # depth downscale = strided copy in half precision
    half = be.read_image(fe.image("depthHalf"), 0, np.float16).reshape(H // 2, W // 2)
    assert np.array_equal(half, depth[::2, ::2].astype(np.float16))
    # the frame is finite and not black; determinism: the same sequence again gives the same bytes
    out = fe.read_output().reshape(H, W, 4)
    assert out[..., :3].mean() > 5 and (out[..., 3] == 255).all()
    first = out.copy()
    s.close()
    s2 = Sequence(ffi, cuda, W, H, 100)
    s2.fe.backend.set_graph_replay_enabled(True)
    for _ in range(3):
        s2.step(inputs=inputs)
    assert np.array_equal(s2.fe.read_output().reshape(H, W, 4), first)
    assert prev.size == W * H
    s2.close()
