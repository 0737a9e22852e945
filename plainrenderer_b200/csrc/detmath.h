// detmath.h - deterministic single-precision transcendental functions.
//
// The reference's per-pixel arithmetic is GLSL run by a Vulkan driver, whose log/exp/pow/sin/cos/
// acos/atan precision is implementation defined. To make "bit-exact histogram bins"
// (histogramPerTile.comp:49-56 depends on log()) and frame parity testable, every transcendental on
// the frame path is pinned to the functions in this header. Each is a fixed sequence of IEEE-754
// binary32 add/mul/fma/div/sqrt plus integer bit operations (fma only where written: dm::fma_, one
// correctly rounded operation on both sides), so the same bits come out of nvcc (-fmad=false) and
// gcc (-ffp-contract=off). It plays the role of "the libm both sides link":
// the CUDA kernels use it on the device and the CPU oracle includes it instead of glibc's libm.
// tests/test_detmath.py checks every function against float64 numpy (max error in ulp).
//
// Algorithms: classic Cody-Waite range reduction + minimax polynomials in the style of the public
// domain Cephes single precision library (S. Moshier); accuracy is 1-2 ulp on the ranges used.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD inline
#endif

namespace dm {

DM_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

DM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// fused multiply-add, written explicitly where the contract contracts (Horner steps, Cody-Waite reduction)
DM_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// ---- contract "fast"  ----
// -DPLAIN_FAST_CONTRACT (device code of the floating-point passes only: GI, shading, TAA/bloom/tonemap, froxels) replaces the
// pinned sequences below by the SFU approximations (ex2 / lg2 / sin / cos / rcp / rsqrt / sqrt .approx.ftz, 1-2 ulp or 2^-21
// absolute) and lets ptxas contract; the results then match the oracle within a tolerance instead of bit for bit. The exact
// build, the oracle and the integer / LUT / rasterisation passes never see this branch. (A host translation unit that defines DM_FAST and
// the hw_* functions itself BEFORE this header takes the same branches: the test suite's error model of this contract does that.)
#if defined(PLAIN_FAST_CONTRACT) && defined(__CUDA_ARCH__)
#define DM_FAST 1
__device__ __forceinline__ float hw_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hw_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

DM_HD float nanf_() { return u2f(0x7fc00000u); }
DM_HD float inff_() { return u2f(0x7f800000u); }
DM_HD bool isnan_(float x) { return x != x; }
DM_HD float abs_(float x) { return u2f(f2u(x) & 0x7fffffffu); }

// floor for |x| < 2^31 via truncation (exact)
DM_HD float floor_(float x) {
#if defined(DM_FAST)
    return floorf(x);
#elif defined(__CUDA_ARCH__)
    return floorf(x) + 0.f;  // one FRND; "+ 0" turns floor(-0) = -0 into the +0 the portable sequence below yields
#else
    if (!(abs_(x) < 8388608.f)) return x;  // already integral (or nan/inf)
    float t = (float)(int)x;
    return (t > x) ? t - 1.f : t;
#endif
}

// x * 2^n, n roughly in [-300, 300]; handles results in the denormal range
DM_HD float ldexp_(float x, int n) {
    if (n > 127) {
        x *= u2f(0x7f000000u);  // 2^127
        n -= 127;
        if (n > 127) {
            x *= u2f(0x7f000000u);
            n -= 127;
            if (n > 127) n = 127;
        }
    } else if (n < -126) {
        x *= u2f(0x0c800000u);  // 2^-126 * 2^24 = 2^-102
        n += 102;
        if (n < -126) {
            x *= u2f(0x0c800000u);
            n += 102;
            if (n < -126) n = -126;
        }
    }
    return x * u2f((uint32_t)(0x7f + n) << 23);
}

// x = m * 2^e with m in [0.5, 1); x must be finite and > 0
DM_HD float frexp_pos(float x, int* e) {
    uint32_t u = f2u(x);
    int ex = (int)(u >> 23);
    int bias = 0;
    if (ex == 0) {  // denormal: scale up
        x *= 33554432.f;  // 2^25
        u = f2u(x);
        ex = (int)(u >> 23);
        bias = 25;
    }
    *e = ex - 126 - bias;
    return u2f((u & 0x007fffffu) | 0x3f000000u);
}

// e^r for |r| <= ~0.35
DM_HD float exp_core(float r) {
    float z = r * r;
    float p = 1.9875691500E-4f;
    p = fma_(p, r, 1.3981999507E-3f);
    p = fma_(p, r, 8.3334519073E-3f);
    p = fma_(p, r, 4.1665795894E-2f);
    p = fma_(p, r, 1.6666665459E-1f);
    p = fma_(p, r, 5.0000001201E-1f);
    p = fma_(p, z, r);
    return p + 1.f;
}

DM_HD float exp(float x) {
#if defined(DM_FAST)
    return hw_ex2(x * 1.44269504088896341f);
#endif
    if (isnan_(x)) return x;
    if (x > 88.72283905206835f) return inff_();
    if (x < -103.972084f) return 0.f;
    float z = floor_(fma_(1.44269504088896341f, x, 0.5f));
    float r = fma_(z, -0.693359375f, x);
    r = fma_(z, 2.12194440e-4f, r);
    return ldexp_(exp_core(r), (int)z);
}

DM_HD float exp2(float x) {
#if defined(DM_FAST)
    return hw_ex2(x);
#endif
    if (isnan_(x)) return x;
    if (x >= 128.f) return inff_();
    if (x < -150.f) return 0.f;
    float z = floor_(x + 0.5f);
    float f = x - z;  // exact, in [-0.5, 0.5]
    return ldexp_(exp_core(f * 0.693147180559945309f), (int)z);
}

// ln(m) for the reduced argument; returns polynomial pieces. m in [0.5,1) -> reduced to
// [sqrt(1/2), sqrt(2)) - 1
DM_HD float log_reduced(float m, int* e, float* xr) {
    float x;
    if (m < 0.707106781186547524f) {
        *e -= 1;
        x = (m + m) - 1.f;
    } else {
        x = m - 1.f;
    }
    float z = x * x;
    float y = 7.0376836292E-2f;
    y = fma_(y, x, -1.1514610310E-1f);
    y = fma_(y, x, 1.1676998740E-1f);
    y = fma_(y, x, -1.2420140846E-1f);
    y = fma_(y, x, 1.4249322787E-1f);
    y = fma_(y, x, -1.6668057665E-1f);
    y = fma_(y, x, 2.0000714765E-1f);
    y = fma_(y, x, -2.4999993993E-1f);
    y = fma_(y, x, 3.3333331174E-1f);
    y = y * x * z;
    *xr = x;
    return y;  // ln(1+x) = x - z/2 + y
}

DM_HD float log(float x) {
#if defined(DM_FAST)
    return hw_lg2(x) * 0.693147180559945309f;
#endif
    if (isnan_(x)) return x;
    if (x < 0.f) return nanf_();
    if (x == 0.f) return -inff_();
    if (f2u(x) == 0x7f800000u) return x;
    int e;
    float m = frexp_pos(x, &e);
    float xr;
    float z;
    float y = log_reduced(m, &e, &xr);
    z = xr * xr;
    float fe = (float)e;
    y = fma_(-2.12194440e-4f, fe, y);
    y = fma_(-0.5f, z, y);
    z = xr + y;
    z = fma_(0.693359375f, fe, z);
    return z;
}

DM_HD float log2(float x) {
#if defined(DM_FAST)
    return hw_lg2(x);
#endif
    if (isnan_(x)) return x;
    if (x < 0.f) return nanf_();
    if (x == 0.f) return -inff_();
    if (f2u(x) == 0x7f800000u) return x;
    int e;
    float m = frexp_pos(x, &e);
    float xr;
    float y = log_reduced(m, &e, &xr);
    float z = xr * xr;
    y = fma_(-0.5f, z, y);
    // ln(1+x) = xr + y ; log2 = e + (xr + y) * log2(e), split for accuracy
    float r = y * 1.44269504088896341f;
    r = fma_(xr, 0.44269504088896341f, r);
    r = r + xr;
    return r + (float)e;
}

// GLSL pow(x, y): undefined for x < 0. Pinned: a negative base is clamped to 0 (the frame path raises values like
// 1 - |dot(N, V)| to the 5th power, brdf.inc:35,50-52, and the dot of two normalised vectors overshoots 1 by an ulp;
// a NaN there would be spread over the whole frame by the bloom chain). pow(0, y>0) = 0.
DM_HD float pow(float x, float y) {
#if defined(DM_FAST)
    if (y == 0.f) return 1.f;
    return hw_ex2(y * hw_lg2(fmaxf(x, 0.f)));  // base clamped to 0 as below; lg2(0) = -inf gives 0 (y > 0) or inf (y < 0)
#endif
    if (isnan_(x) || isnan_(y)) return nanf_();
    if (y == 0.f) return 1.f;
    if (x < 0.f) x = 0.f;
    if (x == 0.f) return (y > 0.f) ? 0.f : inff_();
    return exp2(y * log2(x));
}

// ---- trigonometry (|x| < 8192 rad) ----
DM_HD float sincos_reduce(float ax, int* jout) {
    int j = (int)(1.27323954473516f * ax);  // 4/pi
    float y = (float)j;
    if (j & 1) {
        j += 1;
        y += 1.f;
    }
    *jout = j;
    float r = fma_(y, -0.78515625f, ax);
    r = fma_(y, -2.4187564849853515625e-4f, r);
    r = fma_(y, -3.77489497744594108e-8f, r);
    return r;
}

DM_HD float sin_poly(float x, float z) {
    float y = -1.9515295891E-4f;
    y = fma_(y, z, 8.3321608736E-3f);
    y = fma_(y, z, -1.6666654611E-1f);
    y = y * z;
    return fma_(y, x, x);
}

DM_HD float cos_poly(float z) {
    float y = 2.443315711809948E-005f;
    y = fma_(y, z, -1.388731625493765E-003f);
    y = fma_(y, z, 4.166664568298827E-002f);
    y = y * z * z;
    y = fma_(-0.5f, z, y);
    return y + 1.f;
}

DM_HD float sin(float x) {
#if defined(DM_FAST)
    return hw_sin(x);
#endif
    if (isnan_(x) || abs_(x) > 8192.f) return nanf_();
    bool neg = x < 0.f;
    float ax = abs_(x);
    int j;
    float r = sincos_reduce(ax, &j);
    j &= 7;
    if (j > 3) {
        neg = !neg;
        j -= 4;
    }
    float z = r * r;
    float y = (j == 1 || j == 2) ? cos_poly(z) : sin_poly(r, z);
    return neg ? -y : y;
}

DM_HD float cos(float x) {
#if defined(DM_FAST)
    return hw_cos(x);
#endif
    if (isnan_(x) || abs_(x) > 8192.f) return nanf_();
    float ax = abs_(x);
    int j;
    float r = sincos_reduce(ax, &j);
    j &= 7;
    bool neg = false;
    if (j > 3) {
        j -= 4;
        neg = !neg;
    }
    if (j > 1) neg = !neg;
    float z = r * r;
    float y = (j == 1 || j == 2) ? sin_poly(r, z) : cos_poly(z);
    return neg ? -y : y;
}

DM_HD float tan(float x) { return sin(x) / cos(x); }

#define DM_PIF 3.141592653589793238f
#define DM_PIO2F 1.5707963267948966192f
#define DM_PIO4F 0.7853981633974483096f

DM_HD float sqrt_(float x) {
#if defined(DM_FAST)
    return hw_sqrt(x);
#elif defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

DM_HD float asin(float x) {
    if (isnan_(x)) return x;
    bool neg = x < 0.f;
    float a = abs_(x);
    if (a > 1.f) return nanf_();
    if (a < 1.0e-4f) return x;
    float z, s;
    bool flag = a > 0.5f;
    if (flag) {
        z = 0.5f * (1.f - a);
        s = sqrt_(z);
    } else {
        s = a;
        z = s * s;
    }
    float p = 4.2163199048E-2f;
    p = fma_(p, z, 2.4181311049E-2f);
    p = fma_(p, z, 4.5470025998E-2f);
    p = fma_(p, z, 7.4953002686E-2f);
    p = fma_(p, z, 1.6666752422E-1f);
    p = p * z;
    p = fma_(p, s, s);
    if (flag) {
        p = p + p;
        p = DM_PIO2F - p;
    }
    return neg ? -p : p;
}

// input is clamped to [-1, 1] (GLSL leaves |x| > 1 undefined; normalised vectors overshoot by an ulp)
DM_HD float acos(float x) {
    if (isnan_(x)) return x;
    if (x > 1.f) x = 1.f;
    if (x < -1.f) x = -1.f;
    if (x < -0.5f) return DM_PIF - 2.f * asin(sqrt_(0.5f * (1.f + x)));
    if (x > 0.5f) return 2.f * asin(sqrt_(0.5f * (1.f - x)));
    return DM_PIO2F - asin(x);
}

DM_HD float atan(float x) {
    if (isnan_(x)) return x;
    bool neg = x < 0.f;
    float a = abs_(x);
    float y;
    if (a > 2.414213562373095f) {
        y = DM_PIO2F;
        a = -(1.f / a);
    } else if (a > 0.4142135623730950f) {
        y = DM_PIO4F;
        a = (a - 1.f) / (a + 1.f);
    } else {
        y = 0.f;
    }
    float z = a * a;
    float p = 8.05374449538e-2f;
    p = fma_(p, z, -1.38776856032E-1f);
    p = fma_(p, z, 1.99777106478E-1f);
    p = fma_(p, z, -3.33329491539E-1f);
    p = p * z;
    p = fma_(p, a, a);
    y = y + p;
    return neg ? -y : y;
}

// GLSL atan(y, x)
DM_HD float atan2(float y, float x) {
    if (isnan_(x) || isnan_(y)) return nanf_();
    if (x == 0.f) {
        if (y > 0.f) return DM_PIO2F;
        if (y < 0.f) return -DM_PIO2F;
        return 0.f;
    }
    float z = atan(y / x);
    if (x < 0.f) z = (y >= 0.f) ? z + DM_PIF : z - DM_PIF;
    return z;
}

}  // namespace dm
