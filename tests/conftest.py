import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


import os


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ffi():
    from plainrenderer_b200 import ffi as m
    return m


@pytest.fixture(scope="session")
def oracle(ffi):
    """CPU oracle (checker). Built on demand from oracle/*.cpp."""
    lib = ROOT / "oracle" / "_build" / "liboracle.so"
    if not lib.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return ffi.Api(str(lib), "oracle_", "oracle_frontend_")


@pytest.fixture(scope="session")
def product_lib():
    import plainrenderer_b200 as pr
    if not pr.LIB_PATH.exists() or not pr.LIB_FAST_PATH.exists():
        pr.build()
    return pr.LIB_PATH


@pytest.fixture(scope="session")
def cuda(product_lib):
    """The product library on a CUDA device. GPU tests fail (not skip) if the library cannot create a backend."""
    import plainrenderer_b200 as pr
    return pr.load()


@pytest.fixture(scope="session")
def cuda_fast(product_lib):
    """libplain_b200_fast.so: the same C-ABI with the floating-point passes under the "fast" contract."""
    import plainrenderer_b200 as pr
    return pr.load("fast")


# ---- independent numpy restatements of exactly specified pieces (texel formats) ----
def decode_small_float(v, mbits):
    v = np.asarray(v, np.uint32)
    e = (v >> mbits).astype(np.int64)
    m = (v & ((1 << mbits) - 1)).astype(np.float64)
    out = np.where(e == 0, m * 2.0 ** (-14 - mbits), (1.0 + m / (1 << mbits)) * 2.0 ** (e - 15))
    out = np.where(e == 31, np.where(m == 0, np.inf, np.nan), out)
    return out


def decode_r11g11b10(packed):
    packed = np.asarray(packed, np.uint32)
    return np.stack([decode_small_float(packed & 0x7FF, 6), decode_small_float((packed >> 11) & 0x7FF, 6), decode_small_float(packed >> 22, 5)], axis=-1)


def random_r11g11b10(rng, n, finite=True):
    """Random packed texels; with finite=True exponents stay below 31 (no inf/NaN)."""
    def chan(mbits):
        e = rng.integers(0, 31 if finite else 32, n, dtype=np.uint32)
        m = rng.integers(0, 1 << mbits, n, dtype=np.uint32)
        return (e << mbits) | m
    return chan(6) | (chan(6) << 11) | (chan(5) << 22)


CAMERA = ((-13.0, -1.7, 0.5), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))
ALL_IMAGES = ["skyTransmission", "skyMultiscatter", "skyLut", "hiz", "depthHalf", "giY0", "giC0", "giY1", "giC1", "giHistY0", "giHistC0", "giHistY1", "giHistC1", "giFullY", "giFullC",
              "froxelMaterial", "froxelScatter", "froxelHist0", "froxelHist1", "froxelIntegration", "color0", "color1", "taaHist0", "taaHist1", "taaLum0", "taaLum1", "post0", "post1", "brdfLut", "output"]
# not written under pass fusion: the outputs of indirectLightUpscale.comp (folded into the shading kernel) and the material / scattering volumes of
# the froxel chain (one launch over froxel columns keeps them in registers)
FUSED_AWAY_IMAGES = ("giFullY", "giFullC", "froxelMaterial", "froxelScatter")
ALL_BUFFERS = [("histogram", 512), ("light", 20), ("sunShadowInfo", 304), ("sdfCulled", None), ("sdfTiles", None)]  # None: the whole buffer (S3: instance culling lists)
# the non-default passes beside the frame path (temporalSupersampling.comp + colorToLuminance.comp, sdfDebugVisualisation.comp)
N4_VARIANTS = [dict(taa_use_separate_supersampling=1), dict(taa_use_separate_supersampling=1, taa_supersample_use_tonemapping=0),
               dict(sdf_debug_mode=1), dict(sdf_debug_mode=2), dict(sdf_debug_mode=3), dict(sdf_debug_mode=4),
               dict(sdf_debug_mode=2, sdf_debug_show_tile_usage_with_hiz=0), dict(sdf_debug_mode=1, sdf_debug_use_influence_radius=1)]


def image_mips(fe, h):
    d = fe.backend.image_description(h)
    if d.mip_count == 1:  # full chain
        return 1 + int(np.floor(np.log2(max(d.width, d.height, d.depth))))
    if d.mip_count == 2:
        return d.manual_mip_count
    return 1


class Sequence:
    """Drives the same frame sequence through one library (CUDA product or CPU oracle)."""

    def __init__(self, ffi, api, w, h, instances=12, frames=None, pass_fusion=False, **settings):
        self.ffi, self.api, self.w, self.h = ffi, api, w, h
        settings.setdefault("sun_direction_deg", (40.0, 35.0))
        self.s = ffi.default_settings(api, w, h, **settings)
        self.fe = ffi.Frontend(api, self.s)
        # pass fusion is the product's default; with it the images in FUSED_AWAY_IMAGES are not written, so the sequences that compare
        # EVERY image run unfused and the fused path is compared on everything else (run_both does both)
        self.pass_fusion = pass_fusion
        self.fe.backend._check(api.b["set_pass_fusion_enabled"](self.fe.backend.ctx, 1 if pass_fusion else 0), "set_pass_fusion_enabled")
        self.scene = ffi.SyntheticScene(api, n_instances=instances)
        self.scene.attach(self.fe)
        self.fe.set_exposure(2e-5)
        self.frame = 0
        self.prev_cam = None

    def camera(self, f, moving):
        p, fw, r, u = CAMERA
        if moving:
            p = (p[0] + 0.35 * f, p[1] - 0.02 * f, p[2] + 0.11 * f)
        return self.ffi.camera(p, fw, r, u)

    def step(self, moving=False, inputs=None):
        f = self.frame
        cam = self.camera(f, moving)
        if inputs is None:
            inputs = self.scene.render_inputs(self.s, cam, f + 1, prev_cam=self.prev_cam, shadows=True, threads=0)
        self.fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, inputs["depth"], inputs["motion"], inputs["normal"], inputs["gbuffer"], inputs.get("shadow_maps"))
        self.prev_cam = cam
        self.frame += 1
        return inputs

    def snapshot(self, images=ALL_IMAGES, buffers=ALL_BUFFERS, skip=()):
        out = {}
        for name in images:
            if name in skip:
                continue
            h = self.fe.image(name)
            for mip in range(image_mips(self.fe, h)):
                out["%s/%d" % (name, mip)] = self.fe.backend.read_image(h, mip).copy()
        for name, size in buffers:
            if size is None:
                size = self.buffer_bytes(name)
            out["buf:" + name] = self.fe.backend.read_storage_buffer(self.fe.storage_buffer(name), size).copy()
        return out

    def buffer_bytes(self, name):
        """sizes of the SDF culling buffers as SDFGI::init allocates them (Techniques.cpp: SDFGI.cpp:139-151)"""
        if name == "sdfCulled":
            return 4 + 4 * 1200  # maxObjectCountMainScene u32 + the count
        if name == "sdfTiles":
            tw, th = (self.w + 31) // 32, (self.h + 31) // 32
            return tw * th * 404
        raise KeyError(name)

    def close(self):
        self.scene.close()
        self.fe.close()


def assert_snapshots_equal(a, b, context=""):
    bad = []
    for k in a:
        n = int((a[k] != b[k]).sum())
        if n:
            bad.append("%s: %d/%d bytes differ" % (k, n, a[k].size))
    assert not bad, "%s differs from the oracle: %s" % (context, "; ".join(bad))


class PlainSceneSequence:
    """frames rendered end to end from `.plain` meshes - depth / motion / normal, the shadow cascades and the packed
    G-buffer come from the backend's rasteriser (settings.raster_inputs = 1), nothing is uploaded. Scene: the three golden assets
    of tests/golden/sdf (written by the reference's asset pipeline), instanced with translations / scales / a rotation."""
    PLACEMENT = [("slab", (0.0, 2.6, 0.0), (4.0, 0.4, 4.0), 0.0), ("cube", (-6.0, 0.0, 1.5), (1.0, 1.0, 1.0), 0.5), ("tall", (-4.0, -1.0, -3.0), (1.0, 0.5, 1.0), 0.0),
                 ("cube", (-9.5, 0.8, -0.5), (0.6, 0.6, 0.6), -0.8), ("cube", (-3.0, 1.2, 3.5), (1.5, 1.0, 0.7), 1.1), ("tall", (-14.0, 0.0, 4.0), (1.0, 0.6, 1.5), 0.3)]

    def __init__(self, ffi, api, asset_lib, w, h, device=0, **settings):
        self.ffi, self.w, self.h = ffi, w, h
        settings.setdefault("sun_direction_deg", (40.0, 35.0))
        self.s = ffi.default_settings(api, w, h, raster_inputs=1, **settings)
        self.fe = ffi.Frontend(api, self.s, device=device)
        self.fe.backend._check(api.b["set_pass_fusion_enabled"](self.fe.backend.ctx, 0), "set_pass_fusion_enabled")  # every image is compared
        be = self.fe.backend
        rng = np.random.default_rng(5)
        checker = np.zeros((4, 4, 4), np.uint8)
        checker[..., :3] = rng.integers(60, 255, (4, 4, 3))
        checker[..., 3] = 255
        checker[0, 0, 3], checker[2, 1, 3], checker[3, 3, 3] = 0, 90, 160  # cut-out texels: the alpha test of the prepass / shadow passes
        rough = np.zeros((2, 2, 4), np.uint8)
        rough[..., 1], rough[..., 2] = rng.integers(60, 230, (2, 2)), 0
        bumps = np.zeros((4, 4, 4), np.uint8)
        bumps[..., :2] = rng.integers(108, 148, (4, 4, 2))
        bumps[..., 2:] = 255
        tex = [be.global_texture_index(be.create_image(t.shape[1], t.shape[0], "RGBA8", data=t)) for t in (checker, bumps, rough)]
        golden = ROOT / "tests" / "golden" / "sdf"
        self.meshes = {}
        for k, name in enumerate(("cube", "slab", "tall")):
            m = asset_lib.load_scene(golden / (name + ".plain")).meshes[0]
            brick = asset_lib.load_brick(golden / (name + ".dds"))
            fm = self.fe.register_sdf_mesh(brick, m.bb_min, m.bb_max, m.mean_albedo)
            self.fe.set_mesh_geometry(fm, m.indices, m.vertices, textures=(tex[0], tex[1], tex[2]) if k != 1 else (None, None, tex[2]))
            self.meshes[name] = (fm, m)
        objects = []
        for name, t, sc, rot in self.PLACEMENT:
            fm, m = self.meshes[name]
            c, s_ = np.cos(rot), np.sin(rot)
            M = np.array([[c * sc[0], 0, s_ * sc[2], t[0]], [0, sc[1], 0, t[1]], [-s_ * sc[0], 0, c * sc[2], t[2]], [0, 0, 0, 1]], np.float32)
            corners = np.array([[x, y, z, 1] for x in (m.bb_min[0], m.bb_max[0]) for y in (m.bb_min[1], m.bb_max[1]) for z in (m.bb_min[2], m.bb_max[2])], np.float32) @ M.T
            objects.append((fm, M.T.ravel(), corners[:, :3].min(0), corners[:, :3].max(0)))
        self.fe.set_scene(objects)
        self.fe.set_exposure(2e-5)
        self.frame = 0

    def camera_at(self, f, moving, speed=1.0):
        p, fw, r, u = CAMERA
        if moving:
            p = (p[0] + 0.3 * speed * f, p[1] - 0.02 * speed * f, p[2] + 0.1 * speed * f)
        return self.ffi.camera(p, fw, r, u)

    def step(self, moving=False):
        f = self.frame
        self.fe.render_frame(self.camera_at(f, moving), (f + 1) / 60.0, 1 / 60.0)
        self.frame += 1

    RASTER_IMAGES = ["depth0", "depth1", "motion0", "motion1", "motion2", "normal", "gbuffer", "shadow0", "shadow1", "shadow2"]
    buffer_bytes = Sequence.buffer_bytes

    def snapshot(self, images=None, buffers=ALL_BUFFERS):
        return Sequence.snapshot(self, images or (self.RASTER_IMAGES + ALL_IMAGES), buffers)

    def close(self):
        self.fe.close()
