"""brdfLut.comp (the split-sum / diffuse-integral table the shading pass samples, built once at start-up) of the oracle against a float64
numpy restatement written from the GLSL: 1024 Hammersley points, GGX importance sampling with the height-correlated visibility for the
scale / bias of the specular split sum, cosine sampling for the integral of the selected diffuse BRDF (Lambert, Disney, CoD WWII,
Titanfall 2 single component) with the in / out Fresnel factors. The table is RGBA16F: agreement to half precision."""
import numpy as np
import pytest

from passes import PassRig

PI = 3.1415926535  # global.inc:44


def radical_inverse(i):
    bits = i.astype(np.uint64)
    out = np.zeros(i.shape, np.uint64)
    for b in range(32):
        out |= ((bits >> np.uint64(b)) & np.uint64(1)) << np.uint64(31 - b)
    return out.astype(np.float64) * 2.3283064365386963e-10


def np_brdf_lut(res, diffuse_brdf, samples=1024):
    ys, xs = np.mgrid[0:res, 0:res]
    r = np.maximum(xs / res, 0.0001)[..., None]
    NoV = (np.maximum(ys.astype(np.float64), 0.1) / res)[..., None]
    Vx, Vz = np.sqrt(1 - NoV * NoV), NoV
    i = np.arange(samples)
    xi_x, xi_y = (i / samples)[None, None, :], radical_inverse(i)[None, None, :]
    schlick = lambda f0, x: f0 + (1 - f0) * (1 - x) ** 5
    # specular: importanceSampleGGX (sampling.inc:4-22) around N = +z: tangent = normalize(cross((1,0,0), N)) = (0,-1,0), bitangent = (1,0,0)
    r4 = r ** 4
    cos_t = np.sqrt((1 - xi_y) / (1 + (r4 - 1) * xi_y))
    sin_t = np.sqrt(1 - cos_t * cos_t)
    phi = 2 * PI * xi_x
    hx, hy, hz = np.sin(phi) * sin_t, -np.cos(phi) * sin_t, cos_t
    VoH_raw = Vx * hx + Vz * hz
    Lz = 2 * VoH_raw * hz - Vz
    VoH, NoH, NoL = np.maximum(VoH_raw, 0), np.maximum(hz, 0), np.maximum(Lz, 0)
    r2 = r * r
    with np.errstate(divide="ignore", invalid="ignore"):
        vis = 0.5 / (NoL * np.sqrt(NoV * NoV * (1 - r2) + r2) + NoV * np.sqrt(NoL * NoL * (1 - r2) + r2))
        k = np.where(NoL > 0, vis * VoH * NoL / NoH, 0.0)
    scale = ((1 - VoH) ** 5 * k).sum(-1) / samples * 4
    bias = k.sum(-1) / samples * 4
    # diffuse: importanceSampleCosine (sampling.inc:25-45)
    phi = 2 * PI * xi_y
    cos_t, sin_t = np.sqrt(xi_x), np.sqrt(1 - xi_x)
    lx, ly, lz = np.sin(phi) * sin_t, -np.cos(phi) * sin_t, cos_t * np.ones_like(phi)
    hx, hy, hz = Vx + lx, ly + 0 * Vx, Vz + lz
    hn = np.sqrt(hx * hx + hy * hy + hz * hz)
    VoH = np.clip((Vx * hx + Vz * hz) / hn, 0, 1)
    NoL, NoH = np.maximum(lz, 0), np.maximum(hz / hn, 0)
    fresnel = (1 - schlick(0.04, NoV)) * (1 - schlick(0.04, NoL))
    if diffuse_brdf == 0:
        f = 1 / PI
    elif diffuse_brdf == 1:
        f90 = 0.5 * r + 2 * VoH * VoH * r
        f = 1 / PI * (1 + (f90 - 1) * (1 - NoL) ** 5) * (1 + (f90 - 1) * (1 - NoV) ** 5) * (1 * (1 - r) + r / 1.51)
    elif diffuse_brdf == 2:
        f0 = VoH + (1 - VoH) ** 5
        f1 = (1 - 0.75 * (1 - NoL) ** 5) * (1 - 0.75 * (1 - NoV) ** 5)
        g = np.log2(2 / (r * r) - 1) / 18
        t = np.clip(2.2 * g - 0.5, 0, 1)
        f = 1 / PI * (f0 + (f1 - f0) * t + (34.5 * g * g - 59 * g + 24.5) * VoH * 2.0 ** (-np.maximum(73.2 * g - 21.2, 8.9) * np.sqrt(NoH)))
    else:
        LoV = np.clip(lx * Vx + lz * Vz, 0, 1)
        facing = 0.5 + 0.5 * LoV
        rough = facing * (0.9 - 0.4 * facing) * (0.5 + NoH) / np.maximum(NoH, 0.03)
        smooth = 1.05 * (1 - (1 - NoL) ** 5) * (1 - (1 - NoV) ** 5)
        f = 1 / PI * (smooth * (1 - r) + rough * r)
    diffuse = (f * fresnel).sum(-1) / samples
    return np.stack([scale, bias, diffuse, np.zeros_like(scale)], -1)


@pytest.mark.parametrize("diffuse_brdf", [0, 1, 2, 3])
def test_brdf_lut_matches_numpy(ffi, oracle, diffuse_brdf):
    res = 32
    rig = PassRig(ffi, oracle, 64, 64)
    be = rig.be
    lut = be.create_image(res, res, "RGBA16_SFLOAT")
    p = be.create_compute_pass("brdfLut.comp", {0: np.int32(diffuse_brdf)})
    be.new_frame()
    be.set_compute_pass_execution(p, (res // 8, res // 8, 1), storage=[(lut, 0, 0)])
    rig.run()
    got = be.read_image(lut, 0, np.float16).reshape(res, res, 4).astype(np.float64)
    rig.close()
    want = np_brdf_lut(res, diffuse_brdf)
    assert np.isfinite(got).all()
    # a sum of 1024 binary32 terms + the half-float store; at r -> 0 the GGX lobe is a delta (NoH -> 1, k is a ratio of tiny numbers)
    err = np.abs(got - want) / np.maximum(np.abs(want), 0.02)
    assert err[:, 1:].max() < 4e-3, "worst %.2e at %s" % (err[:, 1:].max(), np.unravel_index(err[:, 1:].argmax(), err[:, 1:].shape))
    assert err[:, 0].max() < 3e-2
    # physics: y = directional albedo of a GGX lobe with F = 1: <= 1, -> 1 for a smooth surface seen head on; x = its (1 - VoH)^5 share
    assert (got[..., 1] < 1.02).all() and got[res - 1, 1, 1] > 0.9 and (got[..., 0] <= got[..., 1] + 1e-3).all()
    if diffuse_brdf == 0:
        assert abs(got[res - 1, 4, 2] - (1 / PI) * 0.96 * 0.913) < 0.01              # the reference accumulates f itself: (1 / pi) x in / out Fresnel
