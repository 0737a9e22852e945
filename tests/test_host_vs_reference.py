"""The host-side mirror (plainrenderer_b200/host/: camera matrices, view frustum, frustum culling, Hammersley jitter, sun direction, mip
counts, SDF bounding-box padding) and the vertex packing of ffi.pack_vertices against the REFERENCE'S OWN host code: oracle/_ref/
libref_host.so is compiled by oracle/build_ref.sh from the reference's sources where they lie (Camera.cpp, ViewFrustum.cpp, Culling.cpp,
MathUtils.cpp, sdfUtilities.cpp, CompressedTypes.cpp) behind the C entry points of oracle/ref/ref_host_shim.cpp. Transcendentals differ
by an ulp (the mirror uses the pinned libm of csrc/detmath.h, the reference glm / libm), everything else is exact."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import ROOT

f32, u32 = C.c_float, C.c_uint32


@pytest.fixture(scope="module")
def ref():
    lib = ROOT / "oracle" / "_ref" / "libref_host.so"
    if not lib.exists() and (ROOT / "oracle" / "build_ref.sh").exists():
        subprocess.run(["bash", str(ROOT / "oracle" / "build_ref.sh")], check=False, capture_output=True)
    if not lib.exists():
        pytest.skip("oracle/_ref/libref_host.so not built (the reference sources are not present)")
    r = C.CDLL(str(lib))
    r.ref_mipCountFromResolution.restype = u32
    r.ref_vec3ToNormalizedR10B10G10A2.restype = u32
    return r


def arr(n):
    return (f32 * n)()


def vec(v):
    return (f32 * len(v))(*[float(x) for x in v])


CAMERAS = [((-13.0, -1.7, 0.5), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0), 35.0, 16 / 9, 0.1, 300.0),
           ((3.0, -2.0, 7.5), (0.0, 0.6, -0.8), (1.0, 0.0, 0.0), (0.0, -0.8, -0.6), 70.0, 1.0, 0.05, 80.0),
           ((0.0, -1.0, -5.0), (0.0, 0.0, -1.0), (1.0, 0.0, 0.0), (0.0, -1.0, 0.0), 35.0, 2.39, 0.1, 300.0)]


def test_hammersley_sun_direction_mip_count(ffi, oracle, ref):
    for i in range(64):
        a, b = arr(2), arr(2)
        oracle.f["host_hammersley2d"](u32(i), a)
        ref.ref_hammersley2D(u32(i), b)
        assert list(a) == list(b), "hammersley2D(%d)" % i  # bit operations and exact binary fractions / the same float operations
    for deg in [(0, 0), (40, 35), (200, 80), (-30, 10), (123.4, 56.7), (359, 179)]:
        a, b = arr(3), arr(3)
        oracle.f["host_direction_to_vector"](vec(deg), a)
        ref.ref_directionToVector(vec(deg), b)
        assert np.allclose(list(a), list(b), rtol=0, atol=3e-7), deg
    oracle.f["host_mip_count_from_resolution"].restype = u32
    for w, h, d in [(1, 1, 1), (1920, 1080, 1), (960, 540, 1), (3840, 2160, 1), (64, 64, 64), (5, 3, 17), (4096, 4096, 1), (3, 2, 1)]:
        assert oracle.f["host_mip_count_from_resolution"](u32(w), u32(h), u32(d)) == ref.ref_mipCountFromResolution(u32(w), u32(h), u32(d))


def test_camera_matrices_and_view_frustum(ffi, oracle, ref):
    for pos, fwd, right, up, fov, aspect, near, far in CAMERAS:
        cam = ffi.camera(pos, fwd, right, up)
        v0, p0, v1, p1 = arr(16), arr(16), arr(16), arr(16)
        oracle.f["host_camera_matrices"](C.byref(cam), f32(fov), f32(aspect), f32(near), f32(far), v0, p0)
        ref.ref_cameraMatrices(vec(pos), vec(fwd), vec(right), vec(up), f32(fov), f32(aspect), f32(near), f32(far), v1, p1)
        assert np.allclose(list(v0), list(v1), rtol=0, atol=1e-6)        # view: products and sums only
        assert np.allclose(list(p0), list(p1), rtol=3e-7, atol=1e-7)     # projection: tan(fov / 2) through two different libms
        pts0, nrm0, pts1, nrm1 = arr(24), arr(18), arr(24), arr(18)
        oracle.f["host_view_frustum"](C.byref(cam), f32(fov), f32(aspect), f32(near), f32(far), pts0, nrm0)
        ref.ref_viewFrustum(vec(pos), vec(fwd), vec(right), vec(up), f32(fov), f32(aspect), f32(near), f32(far), pts1, nrm1)
        assert np.allclose(list(pts0), list(pts1), rtol=2e-6, atol=2e-5)
        assert np.allclose(list(nrm0), list(nrm1), rtol=0, atol=2e-6)


def test_frustum_culling_decisions(ffi, oracle, ref):
    rng = np.random.default_rng(11)
    pos, fwd, right, up, fov, aspect, near, far = CAMERAS[0]
    pts, nrm = arr(24), arr(18)
    ref.ref_viewFrustum(vec(pos), vec(fwd), vec(right), vec(up), f32(fov), f32(aspect), f32(near), f32(far), pts, nrm)
    inside = 0
    for _ in range(2000):
        c = np.array(pos) + rng.uniform(-40, 60, 3) * [1.0, 0.4, 0.8]
        half = rng.uniform(0.05, 6.0, 3)
        a = oracle.f["host_aabb_intersects_frustum"](pts, nrm, vec(c - half), vec(c + half))
        b = ref.ref_aabbIntersectsFrustum(pts, nrm, vec(c - half), vec(c + half))
        assert a == b
        inside += a
    assert 200 < inside < 1800


def test_sdf_bounding_box_padding(ffi, oracle, ref):
    rng = np.random.default_rng(3)
    for _ in range(50):
        lo = rng.uniform(-20, 20, 3)
        hi = lo + rng.uniform(0.01, 30, 3)
        a0, a1, b0, b1 = arr(3), arr(3), arr(3), arr(3)
        oracle.f["host_pad_sdf_bounding_box"](vec(lo), vec(hi), a0, a1)
        ref.ref_padSDFBoundingBox(vec(lo), vec(hi), b0, b1)
        assert list(a0) == list(b0) and list(a1) == list(b1)


def test_vertex_normal_packing(ffi, ref):
    """ffi.pack_vertices (what the raster tests feed create_meshes) packs normals / tangents / bitangents as the reference's
    vec3ToNormalizedR10B10G10A2 (CompressedTypes.cpp:24-46) does"""
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.uniform(-1.2, 1.2, (500, 3)), [[0, 0, 0], [1, 1, 1], [-1, -1, -1], [0.5, -0.5, 0.25], [1e-4, -1e-4, 0.999]]]).astype(np.float32)
    packed = ffi.pack_vertices(np.zeros((len(v), 3), np.float32), normals=v)
    got = np.frombuffer(packed[:, 16:20].tobytes(), np.uint32)
    want = np.array([ref.ref_vec3ToNormalizedR10B10G10A2(vec(x)) for x in v], np.uint32)
    assert np.array_equal(got, want)


def test_sdf_instance_world_to_local(ffi, oracle, ref):
    """worldToLocal of an SDF instance = inverse(model * translate(bbOffset)) (SDFGI.cpp:288-292): the mirror's 4x4 inverse against glm's"""
    rng = np.random.default_rng(9)
    for _ in range(40):
        ang = rng.uniform(0, 2 * np.pi, 3)
        rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
        M = np.eye(4)
        M[:3, :3] = (rx @ ry) * rng.uniform(0.3, 4.0, 3)[None, :]
        M[:3, 3] = rng.uniform(-30, 30, 3)
        model = vec(M.T.ravel())
        off = vec(rng.uniform(-3, 3, 3))
        a, b = arr(16), arr(16)
        oracle.f["host_sdf_world_to_local"](model, off, a)
        ref.ref_sdfWorldToLocal(model, off, b)
        assert np.allclose(list(a), list(b), rtol=2e-5, atol=2e-5)


def test_sun_shadow_frustum(ffi, oracle, ref):
    """the orthographic frustum fitted to the camera frustum in the sun's view space (ViewFrustum.cpp:231-271), against which the
    shadow draws are culled (RenderFrontend.cpp:613-645)"""
    for pos, fwd, right, up, fov, aspect, near, far in CAMERAS:
        pts, nrm = arr(24), arr(18)
        ref.ref_viewFrustum(vec(pos), vec(fwd), vec(right), vec(up), f32(fov), f32(aspect), f32(near), f32(far), pts, nrm)
        for deg in [(40.0, 35.0), (0.0, 0.0), (200.0, 80.0), (90.0, 179.99)]:
            light = arr(3)
            ref.ref_directionToVector(vec(deg), light)
            p0, n0, p1, n1 = arr(24), arr(18), arr(24), arr(18)
            oracle.f["host_orthogonal_frustum_fitted_to_camera"](pts, nrm, light, p0, n0)
            ref.ref_orthogonalFrustumFittedToCamera(pts, nrm, light, p1, n1)
            scale = np.abs(np.array(list(p1))).max()
            assert np.allclose(list(p0), list(p1), rtol=0, atol=2e-4 * scale), (pos, deg)  # a float 4x4 inverse on both sides
            assert np.allclose(list(n0), list(n1), rtol=0, atol=2e-4), (pos, deg)
