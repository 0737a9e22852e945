"""skyTransmissionLut.comp (Hillaire transmittance LUT, S8) of the oracle against a float64 numpy restatement written from the GLSL
(skyTransmissionLut.comp:16-48, sky.inc:12-86: exponential Rayleigh / Mie profiles, the ozone tent, 40 steps from the atmosphere's
edge back to the sample point), and against the closed form for the vertical path - to the precision of the R11G11B10 LUT."""
import numpy as np

from conftest import CAMERA, decode_r11g11b10

# AtmosphereSettings defaults, Sky.h:6-15
RAYLEIGH = np.array([0.0058, 0.0135, 0.0331])
OZONE = np.array([0.000650, 0.001881, 0.000085])
MIE_SCATTER, EARTH, TOP = 0.006, 6371.0, 100.0
MIE_EXT = 1.11 * MIE_SCATTER


def extinction(height):
    return (np.exp(-height / 8)[..., None] * RAYLEIGH + np.exp(-height / 1.2)[..., None] * MIE_EXT + np.maximum(0, 1 - np.abs(height - 25.0) / 15.0)[..., None] * OZONE)


def np_transmission(res=128):
    ys, xs = np.mgrid[0:res, 0:res]
    height = TOP * xs / (res - 1)
    up_dot = np.maximum(ys / (res - 1) * 2 - 1, -0.999)
    V = np.stack([np.zeros_like(up_dot), -up_dot, np.sqrt(1 - up_dot * up_dot)], -1)
    P = np.stack([np.zeros_like(height), -height - EARTH, np.zeros_like(height)], -1)
    o = P - 0.01
    L = -o
    t_ca = (L * V).sum(-1)
    d2 = (L * L).sum(-1) - t_ca * t_ca
    with np.errstate(invalid="ignore"):
        t_earth = t_ca - np.sqrt(EARTH * EARTH - d2)
    hit_earth = t_earth >= 0  # NaN (the ray misses the earth) compares false
    t_atm = t_ca + np.abs(np.sqrt((EARTH + TOP) ** 2 - d2))
    t = np.where(hit_earth, t_earth, t_atm)
    end = o + t[..., None] * V
    path = np.maximum(np.linalg.norm(end - P, axis=-1), 0.01)
    step = path / 40
    pos, absorption = end.copy(), np.ones(V.shape)
    for _ in range(40):
        pos = pos - V * step[..., None]
        h = np.maximum(np.linalg.norm(pos, axis=-1) - EARTH, 0)
        absorption *= np.exp(-extinction(h) * step[..., None])
    return np.where(hit_earth[..., None], 0.0, absorption)


def test_transmission_lut_matches_float64_restatement(ffi, oracle):
    s = ffi.default_settings(oracle, 64, 36)
    fe = ffi.Frontend(oracle, s)
    z4, z16 = np.zeros(64 * 36 * 4, np.uint8), np.zeros(64 * 36 * 16, np.uint8)
    fe.render_frame(ffi.camera(*CAMERA), 1 / 60.0, 1 / 60.0, z4, z4, z4, z16, None)
    h_img = fe.image("skyTransmission")
    d = fe.backend.image_description(h_img)
    lut = decode_r11g11b10(fe.backend.read_image(h_img, 0, np.uint32).reshape(d.height, d.width))
    fe.close()
    assert (d.width, d.height) == (128, 128)
    want = np_transmission(128)
    tol = np.array([2.0 ** -6, 2.0 ** -6, 2.0 ** -5]) * 1.3
    visible = want > 1e-3
    rel = np.abs(lut - want) / np.maximum(want, 1e-9)
    assert (rel <= tol)[visible].mean() > 0.999 and (lut[~visible] < 2e-3).all()
    # physics check, independent of the marching scheme: straight up from the ground, T = exp(-sum of the column integrals)
    column = RAYLEIGH * 8 * (1 - np.exp(-TOP / 8)) + MIE_EXT * 1.2 * (1 - np.exp(-TOP / 1.2)) + OZONE * 15.0
    # (the shader's 40 right-end samples of 2.5 km over-weight the dense air at the bottom - the Mie scale height is 1.2 km: a few per cent)
    assert np.allclose(lut[127, 0], np.exp(-column), rtol=0.08) and (lut[127, 0] <= np.exp(-column) + 1e-3).all()
    # monotone: a longer path through denser air transmits less
    assert (np.diff(lut[127, :, 0]) >= -1e-3).all() and lut[127, -1, 0] > 0.98
