"""SDF bake on the GPU: the CUDA kernel against the bricks written by the reference binary
(tests/golden/sdf, bit-exact) and against the pinned CPU oracle on meshes / resolutions the fixtures do not cover."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden" / "sdf"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(oracle, product_lib):
    from plainrenderer_b200 import assets
    return assets.Assets(), assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_")


@pytest.mark.parametrize("name", ["cube", "slab", "tall"])
def test_cuda_bake_reproduces_the_reference_bricks(libs, name):
    cuda, _ = libs
    mesh = cuda.load_scene(GOLD / (name + ".plain")).meshes[0]
    want = cuda.load_brick(GOLD / (name + ".dds"))
    got, ms = cuda.bake(mesh)
    assert np.array_equal(got, want), "%d of %d texels differ from the reference's brick" % ((got != want).sum(), want.size)
    assert ms > 0.0  # the kernel ran on the device


def random_mesh(seed, triangles):
    """Closed-ish triangle soup: an icosphere-like blob with perturbed vertices (inside/outside votes, grid cells with many triangles)."""
    from plainrenderer_b200 import assets
    rng = np.random.default_rng(seed)
    n_lat, n_lon = max(3, int(np.sqrt(triangles / 2))), max(4, int(np.sqrt(triangles / 2)) * 2)
    th = np.linspace(0.0, np.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    radius = 1.5 + 0.3 * rng.standard_normal((n_lat + 1, n_lon)).clip(-1, 1)
    pts = np.stack([radius * np.sin(th)[:, None] * np.cos(ph)[None, :] * 1.8, radius * np.cos(th)[:, None] * np.ones_like(ph)[None, :], radius * np.sin(th)[:, None] * np.sin(ph)[None, :] * 0.9], -1)
    pos, idx = pts.reshape(-1, 3).astype(np.float32), []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b, c, d = i * n_lon + j, i * n_lon + (j + 1) % n_lon, (i + 1) * n_lon + j, (i + 1) * n_lon + (j + 1) % n_lon
            idx += [a, c, b, b, c, d]
    return assets.Mesh(pos, np.array(idx, np.uint32), pos.min(0), pos.max(0), np.full(3, 0.5, np.float32))


@pytest.mark.parametrize("seed,triangles,extent", [(1, 200, None), (2, 1200, (32, 32, 32)), (3, 60, (64, 16, 16))])
def test_cuda_bake_equals_the_pinned_oracle(libs, seed, triangles, extent):
    cuda, oracle = libs
    mesh = random_mesh(seed, triangles)
    want, _ = oracle.bake(mesh, extent)
    got, _ = cuda.bake(mesh, extent)
    assert np.array_equal(got, want), "%d of %d texels differ from the oracle" % ((got != want).sum(), want.size)


def test_full_size_brick_equals_the_reference_binary(libs):
    """BASELINE configs[0]: the 64^3 brick of the x8 cube from the CUDA bake, bit for bit the one the reference binary wrote (sha256)."""
    import hashlib
    cuda, _ = libs
    mesh = cuda.load_scene(GOLD / "cube64.plain").meshes[0]
    got, ms = cuda.bake(mesh)
    assert got.shape == (64, 64, 64)
    want = (GOLD / "cube64.dds.sha256").read_text().splitlines()[1].split()[0]
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want
    print("64^3 bake: %.2f ms on the device" % ms)


def test_full_size_brick_properties(libs):
    """64^3 (the reference's maximum, BASELINE configs[0]): the scaled cube of the first fixture. Size-independent checks: sign
    inside/outside, distances bounded by the padded box diagonal, symmetric about the cube's mirror planes up to half precision."""
    cuda, _ = libs
    mesh = cuda.load_scene(GOLD / "cube.plain").meshes[0]
    mesh.positions = (mesh.positions * 8.0).astype(np.float32)
    mesh.bb_min, mesh.bb_max = mesh.positions.min(0), mesh.positions.max(0)
    assert cuda.resolution(mesh.bb_min, mesh.bb_max) == (64, 64, 64)
    got, ms = cuda.bake(mesh)
    d = got.view(np.float16).astype(np.float32)
    assert np.isfinite(d).all() and np.abs(d).max() < 40.0
    assert d[32, 32, 32] < 0 and d[0, 0, 0] > 0   # centre inside, corner of the padded volume outside
    print("64^3 bake: %.2f ms on the device" % ms)
