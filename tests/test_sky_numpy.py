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


# ---------------- skyMultiscatterLut.comp:19-124 and skyLut.comp:36-97 ----------------
import passes  # noqa: E402
from test_gi_temporal_upscale_numpy import bilinear  # noqa: E402


def coefficients(height):  # sky.inc:30-44
    r, m, oz = np.exp(-height / 8), np.exp(-height / 1.2), np.maximum(0, 1 - np.abs(height - 25.0) / 15.0)
    return r[..., None] * RAYLEIGH, m[..., None] * np.full(3, MIE_SCATTER), r[..., None] * RAYLEIGH + m[..., None] * MIE_EXT + oz[..., None] * OZONE


def ray_earth_intersection(P, D):
    """sky.inc:62-83 for a point P on the planet's y axis. The distance to the ground, t_ca - sqrt(R^2 - d^2), cancels catastrophically
    (R^2 = 4e7 km^2 against path lengths of metres): it is evaluated in binary32 like the shader does - for P on the axis each dot()
    of the contract has a single non-zero product, so plain binary32 products are the same operations."""
    f = np.float32
    P, D = P.astype(f), D.astype(f)
    L = -P
    t_ca = (L[..., 1] * D[..., 1]).astype(f)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(((L[..., 1] * L[..., 1]).astype(f) - (t_ca * t_ca).astype(f)).astype(f))
        dd = (d * d).astype(f)
        t_earth = (t_ca - np.sqrt((f(EARTH) * f(EARTH) - dd).astype(f))).astype(f)
        r = f(EARTH) + f(TOP)
        t_atm = (t_ca + np.abs(np.sqrt((r * r - dd).astype(f)))).astype(f)
    hit = t_earth >= 0
    t = np.where(hit, t_earth, t_atm)
    return (P + t[..., None] * D).astype(np.float64), t.astype(np.float64), hit


def integrate_inscattering(inscattering, extinction, length):  # volumeShading.inc:24-26
    return (inscattering - inscattering * np.exp(-extinction * length)) / np.maximum(extinction, 0.00001)


def np_multiscatter(T, res=32):
    ys, xs = np.mgrid[0:res, 0:res]
    height = TOP * xs / res
    P = np.stack([0 * height, -height - EARTH, 0 * height], -1)
    up_dot = ys / res * 2 - 1
    L = np.stack([0 * up_dot, -up_dot, np.sqrt(1 - up_dot * up_dot)], -1)
    L_2nd, f_ms = np.zeros((res, res, 3)), np.zeros((res, res, 3))
    for i in range(8):
        for _ in range(8):                                     # phi is computed but never used (:52): eight identical samples per theta
            theta = np.pi * i / 8
            s, c = np.sin(theta), np.cos(theta)
            V = np.broadcast_to(np.array([s * c, -c, s * s]), P.shape)   # :56 not a unit vector for 0 < theta < pi - as the reference
            pos, dist, hit = ray_earth_intersection(P, V)
            step = (dist / 20)[..., None]
            normal = pos / np.linalg.norm(pos, axis=-1, keepdims=True)
            earth_nol = np.clip((normal * L).sum(-1), 0, 1)
            up = P / np.linalg.norm(P, axis=-1, keepdims=True)
            t_hit = bilinear(T, np.zeros_like(height), (up * L).sum(-1) * 0.5 + 0.5)
            direct = np.where(hit[..., None], 0.3 / np.pi * t_hit * earth_nol[..., None], 0.0)
            scatter_r, scatter_m, ext = coefficients(height)   # "approximation": height and up stay those of the start point
            t_sun = bilinear(T, (-P[..., 1] - EARTH) / TOP, -L[..., 1] * 0.5 + 0.5)
            transmission, L_f, inscattered = np.ones((res, res, 3)), np.zeros((res, res, 3)), np.zeros((res, res, 3))
            for _ in range(20):
                integral = integrate_inscattering(scatter_r + scatter_m, ext, step)
                L_f += integral * transmission
                inscattered += integral * t_sun / (4 * np.pi) * transmission
                transmission *= np.exp(-ext * step)
            direct = direct * transmission
            f_ms += L_f * s
            L_2nd += (direct * transmission + inscattered) * s   # :112 + :115 the ground term carries the transmission twice
    return (L_2nd / 64) / (1 - f_ms / 64)


def np_sky_view(T, M, sun, sun_strength_exposed, w=200, h=100, g=0.75):
    ys, xs = np.mgrid[0:h, 0:w]
    theta = (1 - ys / h) - 0.5                                  # fromSkyLut, sky.inc:97-104
    theta = np.sign(theta) * theta * theta * 2 * np.pi + np.pi * 0.5
    phi = (-(xs / w) + 0.5) * 2 * np.pi
    V = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    P = np.broadcast_to(np.array([0, -EARTH - 0.002, 0]), V.shape)
    _, dist, _ = ray_earth_intersection(P, V)
    step = (dist / 30)[..., None]
    VoL = V @ sun
    phase_r = 3 / (16 * np.pi) * (1 + VoL * VoL)
    phase_m = 3 / (8 * np.pi) * (1 - g * g) * (1 + VoL * VoL) / ((2 + g * g) * (1 + g * g - 2 * g * VoL) ** 1.5)   # Cornette-Shanks
    cur, absorption, colour = P.copy(), np.ones(V.shape), np.zeros(V.shape)
    for _ in range(30):
        cur = cur + V * step
        up_len = np.linalg.norm(cur, axis=-1)
        height, up = up_len - EARTH, cur / up_len[..., None]
        u, v = height / TOP, (up @ sun) * 0.5 + 0.5
        transmission = bilinear(T, u, v)
        t_ca = -cur @ sun                                      # shadowRay, skyLut.comp:23-33
        with np.errstate(invalid="ignore"):
            d2 = (cur * cur).sum(-1) - t_ca * t_ca
            t_earth = t_ca - np.sqrt(EARTH * EARTH - d2)
        incoming = sun_strength_exposed * transmission * np.where(t_earth > 0, 0.0, 1.0)[..., None]
        scatter_r, scatter_m, ext = coefficients(height)
        colour = colour + integrate_inscattering(scatter_r * incoming * phase_r[..., None] + scatter_m * incoming * phase_m[..., None], ext, step) * absorption
        absorption = absorption * np.exp(-ext * step)
        colour = colour + bilinear(M, u, v) * incoming * (scatter_r + scatter_m) * step * transmission
    return colour


def code_steps(got, want):
    """distance in units of HALF a step of the R11G11B10 code (6 / 6 / 5 mantissa bits; absolute below the smallest normal 2^-14):
    a correctly rounded result is within 1"""
    return np.abs(got - want) / np.maximum(np.abs(want) * np.array([2.0 ** -7, 2.0 ** -7, 2.0 ** -6]), np.array([2.0 ** -21, 2.0 ** -21, 2.0 ** -20]))


def test_multiscatter_and_sky_view_luts_match_float64_restatements(ffi, oracle):
    sun = np.array([0.3, -0.5, 0.81])
    sun /= np.linalg.norm(sun)
    strength = 3.0
    trans_p, multi_p, sky_p = passes.sky_luts(ffi, oracle, sun, strength)
    T, M, S = decode_r11g11b10(trans_p), decode_r11g11b10(multi_p), decode_r11g11b10(sky_p)
    # the transmittance LUT of this driver is the one the frame renders (pinned above)
    assert (np.abs(T - np_transmission(128)) <= np.maximum(T, 1e-3) * np.array([2.0 ** -6, 2.0 ** -6, 2.0 ** -5]) * 1.3).mean() > 0.999
    # multiscatter from the oracle's transmittance LUT: the nearest code of the restatement. Column 0 is height 0: the start point lies ON
    # the ground sphere, t_earth = t_ca - sqrt(R^2 - d^2) is 0 up to binary32 noise and `hitEarth = t_earth >= 0` is decided by that noise
    want_m = np_multiscatter(T)
    assert code_steps(M[:, 1:], want_m[:, 1:]).max() <= 1.05
    assert np.isfinite(M[:, 0]).all() and (np.abs(M[:, 0] - want_m[:, 0]) <= 0.45 * want_m[:, 0] + 1e-5).all()   # (which downward samples count as ground hits at distance 0 differs)
    assert M.max() < 0.2 and (M[20:, 1:] > 0).all()            # a few per cent of the sun's illuminance wherever the sun is up
    # sky view from the oracle's two LUTs. Rows 0..51: the sky above the horizon - the nearest code of the restatement
    want_s = np_sky_view(T, M, sun, strength)
    assert code_steps(S[:52], want_s[:52]).max() <= 1.05
    # rows 52..95 look at the ground from 2 m above it: the 30 march steps are shorter than one binary32 ulp of the planet-centred
    # coordinate (0.5 m at 6371 km), so the shader's positions - and with them the result - carry that rounding noise
    rel = np.abs(S[52:96] - want_s[52:96]) / np.maximum(want_s[52:96], 1e-6)
    assert np.median(rel) < 0.08 and rel.max() < 0.6 and np.isfinite(S[52:96]).all()
    assert (sky_p[96:] == 0).all()                              # 12 x 8 = 96 rows dispatched for a 100-row image (Sky.cpp:311-312)
    # the sky is brightest in the sun's azimuth, and blue dominates high above the horizon
    by, bx = np.unravel_index(S[:52].sum(-1).argmax(), (52, 200))
    ph = (-(bx / 200) + 0.5) * 2 * np.pi
    assert np.array([np.cos(ph), np.sin(ph)]) @ (sun[[0, 2]] / np.linalg.norm(sun[[0, 2]])) > 0.99   # the brightest texel has the sun's azimuth
    assert S[5, 100, 2] > S[5, 100, 0]
