"""Asset side of the frame path (SURVEY.md 8f N1 + N2), CPU part.

The fixtures under tests/golden/sdf were written by the REFERENCE's own asset pipeline (oracle/_ref/PlainAssetPipeline built by
oracle/build_ref.sh from /root/reference; tests/golden/make_sdf_golden.py): <case>.plain is the mesh exactly as its SDF bake saw
it, <case>.dds the brick it produced. So this is the one place where the oracle is pinned by real reference output:
the CPU restatement of the bake (oracle/sdf_bake.cpp) must reproduce the reference's bricks bit for bit."""
import ctypes as C
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden" / "sdf"
CASES = {"cube": (16, 16, 16), "slab": (32, 16, 16), "tall": (16, 64, 16)}


@pytest.fixture(scope="module")
def oracle_assets(oracle):
    from plainrenderer_b200 import assets
    return assets.Assets(ROOT / "oracle" / "_build" / "liboracle.so", "oracle_asset_")


@pytest.fixture(scope="module")
def product_assets(product_lib):
    from plainrenderer_b200 import assets
    return assets.Assets()


def test_product_library_exports_the_asset_abi(product_lib):
    text = (ROOT / "include" / "plain_assets.h").read_text()
    names = sorted(set(re.findall(r"PLAIN_ASSET\((\w+)\)\(", text)))
    from plainrenderer_b200 import assets
    assert names == sorted(assets.SYMBOLS)
    lib = C.CDLL(str(product_lib))
    for n in names:
        assert hasattr(lib, "plain_asset_" + n)


@pytest.mark.parametrize("name", sorted(CASES))
def test_plain_scene_loader(product_assets, name):
    """.plain layout: ModelLoadSaveBinary.cpp:40-231 (header 24 B, ObjectBinary 72 B, per mesh counts + AABB + 4 paths + mean albedo + indices + 28-byte vertices)."""
    sc = product_assets.load_scene(GOLD / (name + ".plain"))
    assert len(sc.objects) == 1 and len(sc.meshes) == 1
    m = sc.meshes[0]
    assert m.positions.shape == (36, 3) and m.indices.shape == (36,)   # the tinygltf cube: 12 triangles, unshared vertices
    assert m.indices.max() < len(m.positions)
    assert np.allclose(m.positions.min(0), m.bb_min) and np.allclose(m.positions.max(0), m.bb_max)  # AABB stored with the mesh
    assert m.paths["sdf"].endswith("sdfTextures/Cube.dds")
    assert sc.objects[0][1] == 0 and np.isfinite(sc.objects[0][0]).all()
    assert product_assets.resolution(m.bb_min, m.bb_max) == CASES[name]  # 4 texels/m, next power of two, [16, 64] (SceneSDF.cpp:117-131)


def test_loader_rejects_garbage(product_assets, tmp_path):
    from plainrenderer_b200 import assets
    bad = tmp_path / "bad.plain"
    bad.write_bytes(b"nope" + bytes(64))
    with pytest.raises(assets.AssetError):
        product_assets.load_scene(bad)
    good = (GOLD / "cube.plain").read_bytes()
    (tmp_path / "cut.plain").write_bytes(good[:-5])
    with pytest.raises(assets.AssetError):
        product_assets.load_scene(tmp_path / "cut.plain")
    with pytest.raises(assets.AssetError):
        product_assets.load_brick(GOLD / "cube.plain")


def test_dds_round_trip(product_assets, tmp_path):
    want = product_assets.load_brick(GOLD / "slab.dds")
    assert want.shape == (16, 16, 32)
    product_assets.save_brick(tmp_path / "copy.dds", want)
    assert np.array_equal(product_assets.load_brick(tmp_path / "copy.dds"), want)
    assert (tmp_path / "copy.dds").stat().st_size == (GOLD / "slab.dds").stat().st_size


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_bake_reproduces_the_reference_bricks(oracle_assets, name):
    """PIN: the CPU restatement against bricks written by the reference binary (bit-exact, inside and outside texels)."""
    mesh = oracle_assets.load_scene(GOLD / (name + ".plain")).meshes[0]
    want = oracle_assets.load_brick(GOLD / (name + ".dds"))
    got, _ = oracle_assets.bake(mesh)
    assert got.shape == want.shape
    assert np.array_equal(got, want), "%d of %d texels differ from the reference's brick" % ((got != want).sum(), want.size)
    assert (want >> 15).any() and not (want >> 15).all()  # the fixture has texels inside (negative) and outside the mesh


def test_oracle_bake_reproduces_the_reference_brick_at_64_cubed(oracle_assets):
    """BASELINE configs[0] itself: the cube scaled x8 -> 64^3, the reference's maximum resolution (SceneSDF.cpp:120-131). The reference
    binary's brick is pinned by its sha256 (tests/golden/sdf/cube64.dds.sha256, written by make_sdf_golden.py; the brick is 512 KB)."""
    import hashlib
    mesh = oracle_assets.load_scene(GOLD / "cube64.plain").meshes[0]
    assert oracle_assets.resolution(mesh.bb_min, mesh.bb_max) == (64, 64, 64)
    got, _ = oracle_assets.bake(mesh)
    want = (GOLD / "cube64.dds.sha256").read_text().splitlines()[1].split()[0]
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want


def test_bake_has_no_cpu_fallback(product_assets):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from plainrenderer_b200 import assets
    mesh = product_assets.load_scene(GOLD / "cube.plain").meshes[0]
    with pytest.raises(assets.AssetError):
        product_assets.bake(mesh)
