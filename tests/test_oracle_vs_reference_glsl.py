"""The oracle's restatement of the reference's GLSL include files against the reference's OWN text.

oracle/build_ref.sh compiles resources/shaders/{brdf,tonemapping,colorConversion,sky,SDF,bicubicSampling,temporalReprojection,
SphericalHarmonics,sampling,noise,dither,luminance,linearDepth,screenToWorld,volumeShading}.inc from where they lie under
/root/reference as C++ (oracle/ref/glsl_to_cpp.py adapts spelling only: literal suffixes, swizzles, inout, array types) against the
GLSL built-ins of oracle/glsl.h - the numeric contract - and the oracle's image sampler, into oracle/_ref/libref_glsl.so. Every function
is then evaluated on the same random inputs through both libraries and compared BIT FOR BIT: what is pinned is the oracle's reading of
the reference's operations, their order, branches, constants and quirks; what stays the contract's choice is how
a built-in rounds. _ref travels to the GPU box; without it (no /root/reference at build time) the tests skip."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import ROOT

N = 100_000
REF_LIB = ROOT / "oracle" / "_ref" / "libref_glsl.so"


@pytest.fixture(scope="module")
def libs(oracle):
    if not REF_LIB.exists() and (ROOT.parent / "reference").exists():
        subprocess.run(["bash", str(ROOT / "oracle" / "build_ref.sh")], check=False, capture_output=True)
    if not REF_LIB.exists():
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs /root/reference at build time)")
    return C.CDLL(str(REF_LIB)), oracle.lib


def evaluate(lib, prefix, part, name, x, n_out):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((x.shape[0], n_out), np.float32)
    fn = getattr(lib, "%seval_%s" % (prefix, part))
    rc = fn(name.encode(), x.ctypes.data_as(C.c_void_p), C.c_int(x.shape[1]), out.ctypes.data_as(C.c_void_p), C.c_int(n_out), C.c_int(x.shape[0]))
    assert rc == 0, "%s%s(%s) rejected the call" % (prefix, part, name)
    return out


def same_bits(a, b):
    a, b = a.view(np.uint32), b.view(np.uint32)
    nan = lambda u: (u & 0x7FFFFFFF) > 0x7F800000
    return (a == b) | (nan(a) & nan(b))


def check(libs, part, name, x, n_out):
    ref, orc = libs
    a, b = evaluate(ref, "refglsl_", part, name, x, n_out), evaluate(orc, "oracle_inc_", part, name, x, n_out)
    ok = same_bits(a, b)
    assert ok.all(), "%s: %d of %d outputs differ, first at item %d: reference %r oracle %r" % (name, int((~ok).sum()), ok.size, int(np.argwhere(~ok)[0][0]),
                                                                                              a[np.argwhere(~ok)[0][0]], b[np.argwhere(~ok)[0][0]])
    assert np.isfinite(a).mean() > 0.5, name + ": inputs produce mostly non-finite values - not a meaningful comparison"


def unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)


def u32_as_f32(rng, n, hi=2**32):
    return rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32).view(np.float32)


def mix_edges(rng, x, values=(0.0, 1.0, -0.0)):
    """a few percent of the entries replaced by edge values"""
    x = x.copy()
    m = rng.random(x.shape) < 0.03
    x[m] = rng.choice(np.asarray(values, np.float32), int(m.sum()))
    return x


PURE = {
    "D_GGX": (2, 1), "Visibility": (3, 1), "F_Schlick": (7, 3), "DisneyDiffuse": (7, 3), "CoDWWIIDiffuse": (8, 3), "Titanfall2DiffuseSingleComponent": (5, 1),
    "Titanfall2Diffuse": (8, 3), "GGXSingleScattering": (8, 3), "RRTAndODTFit": (3, 3), "ACESFitted": (3, 3), "linearTosRGB": (3, 3), "sRGBToLinear": (3, 3),
    "linearToYCoCg": (3, 3), "YCoCgToLinear": (3, 3), "computeLuminance": (3, 1), "phaseGreenstein": (2, 1), "phaseRayleigh": (1, 1), "cornetteShanksPhase": (2, 1),
}


@pytest.mark.parametrize("name", sorted(PURE))
def test_unit_interval_functions(libs, name):
    """BRDF lobes, tonemapping, colour conversion, phase functions on [0, 1] inputs (with exact 0 / 1 / -0 mixed in)"""
    n_in, n_out = PURE[name]
    rng = np.random.default_rng(hash(name) & 0xFFFF)
    x = mix_edges(rng, rng.random((N, n_in)).astype(np.float32))
    if name in ("ACESFitted", "RRTAndODTFit", "linearTosRGB", "computeLuminance"):
        x = (x * np.exp(rng.normal(0, 3, (N, 1)))).astype(np.float32)  # HDR range
    if name in ("phaseGreenstein", "cornetteShanksPhase", "phaseRayleigh"):
        x[:, 0] = x[:, 0] * 2 - 1
    check(libs, "pure", name, x, n_out)


def test_directions_and_sampling(libs):
    rng = np.random.default_rng(1)
    v = unit(rng, N)
    check(libs, "pure", "directionToSH_L1", v, 4)
    check(libs, "pure", "dominantDirectionFromSH_L1", rng.normal(size=(N, 4)).astype(np.float32), 3)
    check(libs, "pure", "toSkyLut", v, 2)
    check(libs, "pure", "fromSkyLut", rng.random((N, 2)).astype(np.float32), 3)
    xi = mix_edges(rng, rng.random((N, 2)).astype(np.float32))
    nrm = unit(rng, N)
    nrm[:100] = np.array([0, 0, 1], np.float32)  # the |N.z| >= 0.999 branch of the tangent frame
    check(libs, "pure", "importanceSampleCosine", np.concatenate([xi, nrm], 1), 3)
    check(libs, "pure", "importanceSampleGGX", np.concatenate([xi, rng.random((N, 1)).astype(np.float32), nrm], 1), 3)
    ndc = (rng.random((N, 2)) * 2 - 1).astype(np.float32)
    check(libs, "pure", "calculateViewDirectionFromPixel", np.concatenate([ndc, unit(rng, N), unit(rng, N), unit(rng, N), rng.random((N, 1)).astype(np.float32) + 0.1,
                                                                        rng.random((N, 1)).astype(np.float32) + 1.0], 1), 3)
    check(libs, "pure", "computeLutUV", np.concatenate([rng.random((N, 1)).astype(np.float32) * 100, np.full((N, 1), 100, np.float32), unit(rng, N), unit(rng, N)], 1), 2)


def test_integer_hashes_and_dither(libs):
    rng = np.random.default_rng(2)
    seeds = u32_as_f32(rng, N).reshape(N, 1)
    for name, n_out in (("wang_hash", 1), ("xorshift32", 2), ("rand", 2), ("radicalInverse_VdC", 1)):
        check_bits_only(libs, name, seeds, n_out)
    check_bits_only(libs, "hammersley2d", np.stack([u32_as_f32(rng, N, 4096), np.full(N, 4096, np.uint32).view(np.float32)], 1), 2)
    q = (rng.random((N, 2)) * 5000).astype(np.float32)
    q[:50] *= -1  # negative and huge coordinates: the float -> int conversion rule
    q[50:100] = 3e9
    check(libs, "pure", "hash32", q, 3)
    c = rng.random((N, 3)).astype(np.float32)
    uv = rng.integers(0, 4000, (N, 2)).astype(np.float32)
    t = (rng.random((N, 1)) * 100).astype(np.float32)
    check(libs, "pure", "ditherRGB8", np.concatenate([c, uv, t], 1), 3)


def check_bits_only(libs, name, x, n_out):
    ref, orc = libs
    a, b = evaluate(ref, "refglsl_", "pure", name, x, n_out), evaluate(orc, "oracle_inc_", "pure", name, x, n_out)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


def test_depth_sky_and_volume_functions(libs):
    rng = np.random.default_rng(3)
    d = mix_edges(rng, rng.random((N, 1)).astype(np.float32))
    check(libs, "pure", "linearizeDepth", np.concatenate([d, np.full((N, 1), 0.1, np.float32), np.full((N, 1), 300, np.float32)], 1), 1)
    check(libs, "pure", "integrateInscattering", np.concatenate([rng.random((N, 3)), mix_edges(rng, rng.random((N, 3)) * 0.1, (0.0,)), rng.random((N, 1)) * 30], 1).astype(np.float32), 3)
    atmosphere = np.array([5.8e-3, 13.5e-3, 33.1e-3, 6371, 5.8e-3, 13.5e-3, 33.1e-3, 100, 0.65e-3, 1.88e-3, 0.085e-3, 3.996e-3, 4.4e-3, 0.8], np.float32)
    h = (rng.random((N, 1)) * 110 - 5).astype(np.float32)
    check(libs, "pure", "calculateCoefficients", np.concatenate([h, np.tile(atmosphere, (N, 1))], 1), 9)
    P = np.zeros((N, 3), np.float32)
    P[:, 1] = -(6371 + rng.random(N) * 100)
    check(libs, "pure", "rayEarthIntersection", np.concatenate([P, unit(rng, N), np.zeros((N, 3), np.float32), np.full((N, 1), 6371, np.float32), np.full((N, 1), 100, np.float32)], 1), 5)


def set_image(lib, prefix, part, slot, fmt, data, w, h, d):
    data = np.ascontiguousarray(data)
    rc = getattr(lib, "%sset_image_%s" % (prefix, part))(C.c_int(slot), C.c_uint32(fmt), C.c_int(w), C.c_int(h), C.c_int(d), data.ctypes.data_as(C.c_void_p), C.c_size_t(data.nbytes))
    assert rc == 0


def box_brick(res, extents):
    """exact distance field of the box [-e/2 * 0.8, e/2 * 0.8] sampled at texel centres, as halves (a brick the tracer can hit)"""
    ax = [(np.arange(r) + 0.5) / r * e - e / 2 for r, e in zip(res, extents)]
    z, y, x = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    q = np.stack([np.abs(x) - extents[0] * 0.4, np.abs(y) - extents[1] * 0.4, np.abs(z) - extents[2] * 0.4], -1)
    dist = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)
    return dist.astype(np.float16)


def test_sdf_box_intersection_and_sphere_tracer(libs, ffi):
    """rayAABBIntersection / isPointInAABB and the whole of traceRayTroughSDFInstance (SDF.inc:101-184: transform, slab clip, the early-out
    against the closest hit, 128-step march with the Claybook last step, normal, albedo) on rays around a rotated, scaled brick"""
    ref, orc = libs
    rng = np.random.default_rng(4)
    mn = -(rng.random((N, 3)) + 0.1).astype(np.float32)
    o = (rng.normal(size=(N, 3)) * 2).astype(np.float32)
    check(libs, "sdf", "isPointInAABB", np.concatenate([o * 0.3, mn, -mn], 1), 1)
    dirs = unit(rng, N)
    dirs[:200, 0] = 0  # axis-parallel rays: division by zero in the slab test
    check(libs, "sdf", "rayAABBIntersection", np.concatenate([o, dirs, mn, -mn], 1), 2)
    res, ext = (24, 16, 20), (3.0, 2.0, 2.5)
    brick = box_brick(res, ext)
    fmt = ffi.FORMAT["R16_SFLOAT"]
    for lib, prefix in ((ref, "refglsl_"), (orc, "oracle_inc_")):
        set_image(lib, prefix, "sdf", 0, fmt, brick, res[0], res[1], res[2])
    check(libs, "sdf", "normalFromSDF", np.concatenate([rng.random((N, 3)) * 1.2 - 0.1, np.tile(np.asarray(ext, np.float32), (N, 1))], 1).astype(np.float32), 3)
    # world-to-local: rotation * uniform scale + translation (column-major)
    ang = 0.7
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32) * 0.8
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = R
    M[:3, 3] = [0.3, -0.2, 0.1]
    inst = np.concatenate([np.asarray(ext, np.float32), np.array([0.5, 0.6, 0.7], np.float32), M.T.reshape(-1)])
    n = 20_000
    start = (unit(rng, n) * (rng.random((n, 1)) * 6 + 0.2)).astype(np.float32)   # inside and outside the box
    target = (rng.normal(size=(n, 3)) * 0.8).astype(np.float32)
    d = target - start
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    closest = np.where(rng.random((n, 1)) < 0.3, rng.random((n, 1)) * 4, 10000.0).astype(np.float32)  # some rays already have a closer hit (SDF.inc:141)
    x = np.concatenate([np.tile(inst, (n, 1)), start, d.astype(np.float32), closest], 1).astype(np.float32)
    a, b = evaluate(ref, "refglsl_", "sdf", "traceRayTroughSDFInstance", x, 12), evaluate(orc, "oracle_inc_", "sdf", "traceRayTroughSDFInstance", x, 12)
    ok = same_bits(a, b)
    assert ok.all(), "traceRayTroughSDFInstance: %d outputs differ" % int((~ok).sum())
    assert 0.2 < a[:, 0].mean() < 0.95, "the test rays should both hit and miss the brick (hit rate %.2f)" % a[:, 0].mean()


def test_taa_helpers_and_history_sampling(libs, ffi):
    """clipAABB, tonemap pair, Catmull-Rom weight (with the signed-d quirk of its second branch), sampleNeighbourhood + min / max, and the
    five history samplers of temporalFilter.comp:104-127 (bilinear, bicubic 16 / 9 / 5 / 1 tap) on random R11G11B10 images"""
    from conftest import random_r11g11b10
    ref, orc = libs
    rng = np.random.default_rng(5)
    check(libs, "taa", "catmullRomWeight1D", (rng.random((N, 1)) * 6 - 3).astype(np.float32), 1)
    t = rng.random((N, 3)).astype(np.float32) * 2
    lo = rng.random((N, 3)).astype(np.float32)
    hi = lo + rng.random((N, 3)).astype(np.float32) * mix_edges(rng, np.ones((N, 3), np.float32), (0.0,))
    check(libs, "taa", "clipAABB", np.concatenate([t, lo, hi], 1), 3)
    c = (rng.random((N, 3)) * np.exp(rng.normal(0, 2, (N, 1)))).astype(np.float32)
    check(libs, "taa", "tonemap", c, 3)
    check(libs, "taa", "tonemapReverse", rng.random((N, 3)).astype(np.float32), 3)
    w, h = 64, 40
    fmt = ffi.FORMAT["R11G11B10_UFLOAT"]
    hist = (random_r11g11b10(rng, w * h) & np.uint32(0xBBFEFBFF)).reshape(h, w)
    cur = (random_r11g11b10(rng, w * h) & np.uint32(0xBBFEFBFF)).reshape(h, w)
    for lib, prefix in ((ref, "refglsl_"), (orc, "oracle_inc_")):
        set_image(lib, prefix, "taa", 1, fmt, hist, w, h, 1)
        set_image(lib, prefix, "taa", 2, fmt, cur, w, h, 1)
    n = 20_000
    uv = (rng.random((n, 2)) * 1.2 - 0.1).astype(np.float32)  # clamp-to-edge on all sides
    ts = np.tile(np.array([1 / w, 1 / h], np.float32), (n, 1))
    check(libs, "taa", "sampleNeighbourhood", np.concatenate([uv, ts, (rng.random((n, 1)) < 0.5).astype(np.float32)], 1), 33)
    px = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.float32)
    motion = (rng.normal(0, 2.0, (n, 2)) / np.array([w, h])).astype(np.float32)
    res = np.tile(np.array([w, h], np.float32), (n, 1))
    for tech in range(5):
        check(libs, "taa", "historySample", np.concatenate([np.full((n, 1), tech, np.float32), px, motion, res, ts], 1), 3)
