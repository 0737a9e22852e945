// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// sfu_emulation.h - host stand-ins for the SFU approximations the "fast" contract uses on the device (detmath.h, DM_FAST;
// DESIGN.md section 12). Pre-included (`g++ -include`) in front of the oracle translation units of the second checker library,
// liboracle_sfu.so, where it defines DM_FAST so that detmath.h / glsl.h take the fast contract's branches on the host: each
// function returns the correctly rounded result displaced by a deterministic pseudo-random error of the size the PTX ISA documents for the instruction -
//   ex2.approx / rsqrt.approx: 2 ulp;  rcp.approx / sqrt.approx: 1 ulp;  lg2.approx: 2^-22 absolute (2 ulp away from 1);
//   sin.approx / cos.approx: 2^-20.9 absolute;  .ftz: denormal inputs and results flush to zero
// so that the CPU suite can measure how far errors of that size, fed back through the TAA / GI / froxel / exposure histories,
// move the frame (tests/test_fast_contract_emulation.py). It is a model of the error magnitudes, not of the hardware's bits.
#pragma once
#include <cmath>
#include <stdint.h>
#include <string.h>

namespace dm {

inline uint32_t emu_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float emu_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t emu_hash(float x, uint32_t salt) {  // lowbias32 of the argument's bits
    uint32_t h = emu_bits(x) ^ (salt * 0x9e3779b9u);
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
inline float emu_ftz(float x) { return (std::fabs(x) < 1.17549435e-38f) ? std::copysign(0.f, x) : x; }
// displace a finite non-zero result by -n .. +n ulp
inline float emu_ulp(float r, uint32_t h, int n) {
    if (!(std::fabs(r) > 1.17549435e-38f) || std::isinf(r)) return emu_ftz(r);
    const int32_t d = (int32_t)(h % (uint32_t)(2 * n + 1)) - n;
    return emu_ftz(emu_float((uint32_t)((int32_t)emu_bits(r) + d)));
}
// uniform in [-1, 1]
inline double emu_unit(uint32_t h) { return (double)h / 2147483647.5 - 1.0; }

inline float hw_ex2(float x) { x = emu_ftz(x); return emu_ulp((float)std::exp2((double)x), emu_hash(x, 1), 2); }
inline float hw_lg2(float x) {
    x = emu_ftz(x);
    if (x != x || x < 0.f) return emu_float(0x7fc00000u);
    if (x == 0.f) return -INFINITY;
    if (std::isinf(x)) return x;
    const double exact = std::log2((double)x);
    const float a = (float)(exact + emu_unit(emu_hash(x, 2)) * 2.384185791015625e-07);  // 2^-22 absolute
    const float b = emu_ulp((float)exact, emu_hash(x, 3), 2);
    return (std::fabs((double)a - exact) > std::fabs((double)b - exact)) ? a : b;           // the larger of the two documented bounds
}
inline float hw_rcp(float x) { x = emu_ftz(x); return emu_ulp((float)(1.0 / (double)x), emu_hash(x, 4), 1); }
inline float hw_rsqrt(float x) { x = emu_ftz(x); return emu_ulp((float)(1.0 / std::sqrt((double)x)), emu_hash(x, 5), 2); }
inline float hw_sqrt(float x) { x = emu_ftz(x); return emu_ulp((float)std::sqrt((double)x), emu_hash(x, 6), 1); }
inline float hw_sin(float x) { x = emu_ftz(x); return emu_ftz((float)(std::sin((double)x) + emu_unit(emu_hash(x, 7)) * 5.1e-7)); }
inline float hw_cos(float x) { x = emu_ftz(x); return emu_ftz((float)(std::cos((double)x) + emu_unit(emu_hash(x, 8)) * 5.1e-7)); }

}  // namespace dm
#define DM_FAST 1
