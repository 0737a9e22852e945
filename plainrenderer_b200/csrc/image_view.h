// image_view.h - device-side image views, texel codecs and the explicit sampler rules (host+device).
// Layout in HBM: every image is one allocation, mip levels concatenated (256-byte aligned), each level row-major and
// tightly packed (x fastest, then y, then z). A view is one mip level: base pointer + extent.
// Codec and sampler rules are the ones of DESIGN.md "Numeric contract" (RNE to half / R11G11B10, UNORM8 =
// uint(clamp(x)*255+0.5), nearest = floor(u*size), linear with full fp32 weights blended as
// fma(t11,w11, fma(t01,w01, fma(t10,w10, t00*w00))), trilinear fma(slice1, fz, slice0*(1-fz)), NaN coordinate -> 0).
#pragma once
#include "pvec.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace pv {

struct ImgView {
    unsigned char* ptr;
    int w, h, d;
};

// ---------------- scalar codecs ----------------
PV_HD uint32_t encodeSmallFloat(float f, int mbits) {  // unsigned 5-bit-exponent float (R11G11B10 channels), RNE
    uint32_t u = dm::f2u(f);
    uint32_t au = u & 0x7fffffffu;
    const uint32_t expAll = 31u << mbits;
    const uint32_t maxFinite = (30u << mbits) | ((1u << mbits) - 1u);
    if (au > 0x7f800000u) return expAll | 1u;
    if (u >> 31) return 0u;
    if (au == 0x7f800000u) return expAll;
    int e = (int)(au >> 23) - 127;
    if (e > 15) return maxFinite;
    uint32_t m = au & 0x7fffffu;
    uint32_t value;
    if (e >= -14) {
        const int shift = 23 - mbits;
        const uint32_t mant = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        value = ((uint32_t)(e + 15) << mbits) + mant + ((rem > half || (rem == half && (mant & 1u))) ? 1u : 0u);
    } else {
        const int shift = (23 - mbits) + (-14 - e);
        if ((au >> 23) == 0 || shift > 25) return 0u;
        const uint32_t full = m | 0x800000u;
        const uint32_t mant = full >> shift, rem = full & ((1u << shift) - 1u), half = 1u << (shift - 1);
        value = mant + ((rem > half || (rem == half && (mant & 1u))) ? 1u : 0u);
    }
    return value >= expAll ? maxFinite : value;
}
PV_HD float decodeSmallFloat(uint32_t v, int mbits) {  // exact; branch-free: exponent and mantissa shifted into the binary32 fields
    const uint32_t e = v >> mbits;
    const uint32_t shifted = v << (23 - mbits);
    const float normal = dm::u2f(shifted + (112u << 23));                                    // rebias 15 -> 127
    const float denormal = dm::u2f(shifted + (113u << 23)) - dm::u2f(113u << 23);            // (1 + m/2^mbits) * 2^-14 - 2^-14, exact
    const float special = dm::u2f(shifted | 0x7f800000u);                                    // inf / NaN keep their mantissa
    return e == 0 ? denormal : (e == 31 ? special : normal);
}

PV_HD uint16_t floatToHalf(float f) {
#if defined(__CUDA_ARCH__)
    return __half_as_ushort(__float2half_rn(f));  // RNE, overflow -> inf, NaN -> 0x7fff
#else
    uint32_t u = dm::f2u(f), sign = (u >> 16) & 0x8000u, au = u & 0x7fffffffu;
    if (au > 0x7f800000u) return 0x7fffu;
    if (au >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);
    if (au < 0x38800000u) {
        if (au < 0x33000000u) return (uint16_t)sign;
        const uint32_t m = (au & 0x7fffffu) | 0x800000u;
        const int shift = 126 - (int)(au >> 23);
        uint32_t v = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (v & 1u))) v++;
        return (uint16_t)(sign | v);
    }
    uint32_t v = (au - 0x38000000u) >> 13;
    const uint32_t rem = au & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (v & 1u))) v++;
    return (uint16_t)(sign | v);
#endif
}
PV_HD float halfToFloat(uint16_t h) {
#if defined(__CUDA_ARCH__)
    return __half2float(__ushort_as_half(h));
#else
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3ffu;
    if (e == 0) {
        float v = (float)m * 5.9604644775390625e-08f;  // m * 2^-24
        return dm::u2f(dm::f2u(v) | sign);
    }
    if (e == 31) return dm::u2f(sign | 0x7f800000u | (m << 13));
    return dm::u2f(sign | ((e + 112u) << 23) | (m << 13));
#endif
}
// The same codes with fewer instructions (the device code of BOTH contracts uses them since round 2; tests/test_codecs.py holds both
// against the functions above on the host, exhaustively for the decoder and for the encoder over all 2^32 binary32 values; on the
// device only the payload of a NaN can differ, and every store canonicalises NaN):
// encode = round to nearest even by integer arithmetic on the binary32 bits (codes below the smallest normal through one float add),
// decode = the channel IS a half float without sign and with a short mantissa: shift it into place, one conversion.
PV_HD uint32_t encodeSmallFloatFast(float f, int mbits) {
    const uint32_t maxFinite = (30u << mbits) | ((1u << mbits) - 1u);
    const uint32_t u = dm::f2u(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (31u << mbits) | 1u;
    if (u >> 31) return 0u;
    if (u == 0x7f800000u) return 31u << mbits;
    const int shift = 23 - mbits;
    if (u < (113u << 23)) {  // below 2^-14: in [2^(9-mbits), 2^(10-mbits)) one binary32 ulp is one denormal code step
        const float magic = dm::u2f((uint32_t)(127 + 9 - mbits) << 23);
        return dm::f2u(f + magic) - dm::f2u(magic);
    }
    const uint32_t r = u - (112u << 23);  // rebias 127 -> 15
    const uint32_t v = (r + ((1u << (shift - 1)) - 1u) + ((r >> shift) & 1u)) >> shift;
    return v > maxFinite ? maxFinite : v;
}
PV_HD float decodeSmallFloatFast(uint32_t v, int mbits) { return halfToFloat((uint16_t)(v << (10 - mbits))); }
#if defined(__CUDA_ARCH__)
PV_HD uint32_t packR11G11B10(vec3 c) { return encodeSmallFloatFast(c.x, 6) | (encodeSmallFloatFast(c.y, 6) << 11) | (encodeSmallFloatFast(c.z, 5) << 22); }
PV_HD vec3 unpackR11G11B10(uint32_t v) { return v3(decodeSmallFloatFast(v & 0x7ffu, 6), decodeSmallFloatFast((v >> 11) & 0x7ffu, 6), decodeSmallFloatFast(v >> 22, 5)); }
#else
PV_HD uint32_t packR11G11B10(vec3 c) { return encodeSmallFloat(c.x, 6) | (encodeSmallFloat(c.y, 6) << 11) | (encodeSmallFloat(c.z, 5) << 22); }
PV_HD vec3 unpackR11G11B10(uint32_t v) { return v3(decodeSmallFloat(v & 0x7ffu, 6), decodeSmallFloat((v >> 11) & 0x7ffu, 6), decodeSmallFloat(v >> 22, 5)); }
#endif
PV_HD uint32_t floatToUnorm8(float x) { return isnanf_(x) ? 0u : (uint32_t)(clampf(x, 0.f, 1.f) * 255.f + 0.5f); }
PV_HD float unorm8(uint32_t v) { return (float)v / 255.f; }
PV_HD float snorm16(int16_t v) { return fmaxp((float)v / 32767.f, -1.f); }

// ---------------- typed texel access (in range) ----------------
template <typename T> PV_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PV_HD size_t texelIndex(const ImgView& v, int x, int y, int z = 0) { return ((size_t)z * v.h + y) * v.w + x; }
PV_HD bool inRange(const ImgView& v, int x, int y, int z = 0) { return x >= 0 && y >= 0 && z >= 0 && x < v.w && y < v.h && z < v.d; }

PV_HD vec3 loadR11(const ImgView& v, int x, int y) { return unpackR11G11B10(ldg((const uint32_t*)v.ptr + texelIndex(v, x, y))); }
PV_HD void storeR11(const ImgView& v, int x, int y, vec3 c) { ((uint32_t*)v.ptr)[texelIndex(v, x, y)] = packR11G11B10(c); }
PV_HD float loadD32(const ImgView& v, int x, int y) { return ldg((const float*)v.ptr + texelIndex(v, x, y)); }
PV_HD float loadD16(const ImgView& v, int x, int y) { return (float)ldg((const uint16_t*)v.ptr + texelIndex(v, x, y)) / 65535.f; }
PV_HD float loadR16F(const ImgView& v, int x, int y, int z = 0) { return halfToFloat(ldg((const uint16_t*)v.ptr + texelIndex(v, x, y, z))); }
PV_HD void storeR16F(const ImgView& v, int x, int y, float f) { ((uint16_t*)v.ptr)[texelIndex(v, x, y)] = floatToHalf(f); }
PV_HD float loadR8(const ImgView& v, int x, int y, int z = 0) { return unorm8(ldg((const uint8_t*)v.ptr + texelIndex(v, x, y, z))); }
PV_HD vec2 loadRG8(const ImgView& v, int x, int y) {
    const uint16_t t = ldg((const uint16_t*)v.ptr + texelIndex(v, x, y));
    return v2(unorm8(t & 0xffu), unorm8(t >> 8));
}
PV_HD vec3 loadRGBA8rgb(const ImgView& v, int x, int y) {
    const uint32_t t = ldg((const uint32_t*)v.ptr + texelIndex(v, x, y));
    return v3(unorm8(t & 0xffu), unorm8((t >> 8) & 0xffu), unorm8((t >> 16) & 0xffu));
}
PV_HD vec2 loadRG16F(const ImgView& v, int x, int y) {
    const uint32_t t = ldg((const uint32_t*)v.ptr + texelIndex(v, x, y));
    return v2(halfToFloat((uint16_t)(t & 0xffffu)), halfToFloat((uint16_t)(t >> 16)));
}
PV_HD void storeRG16F(const ImgView& v, int x, int y, vec2 c) { ((uint32_t*)v.ptr)[texelIndex(v, x, y)] = (uint32_t)floatToHalf(c.x) | ((uint32_t)floatToHalf(c.y) << 16); }
PV_HD vec4 loadRGBA16F(const ImgView& v, int x, int y, int z = 0) {
    const uint2 t = ldg((const uint2*)v.ptr + texelIndex(v, x, y, z));
    return v4(halfToFloat((uint16_t)(t.x & 0xffffu)), halfToFloat((uint16_t)(t.x >> 16)), halfToFloat((uint16_t)(t.y & 0xffffu)), halfToFloat((uint16_t)(t.y >> 16)));
}
PV_HD void storeRGBA16F(const ImgView& v, int x, int y, int z, vec4 c) {
    uint2 t;
    t.x = (uint32_t)floatToHalf(c.x) | ((uint32_t)floatToHalf(c.y) << 16);
    t.y = (uint32_t)floatToHalf(c.z) | ((uint32_t)floatToHalf(c.w) << 16);
    ((uint2*)v.ptr)[texelIndex(v, x, y, z)] = t;
}
PV_HD vec2 loadRG32F(const ImgView& v, int x, int y) {
    const float2 t = ldg((const float2*)v.ptr + texelIndex(v, x, y));
    return v2(t.x, t.y);
}
PV_HD void storeRG32F(const ImgView& v, int x, int y, vec2 c) {
    float2 t;
    t.x = c.x; t.y = c.y;
    ((float2*)v.ptr)[texelIndex(v, x, y)] = t;
}
PV_HD vec2 loadRG16SNORM(const ImgView& v, int x, int y) {
    const uint32_t t = ldg((const uint32_t*)v.ptr + texelIndex(v, x, y));
    return v2(snorm16((int16_t)(t & 0xffffu)), snorm16((int16_t)(t >> 16)));
}
PV_HD uint4 loadU4(const ImgView& v, int x, int y) { return ldg((const uint4*)v.ptr + texelIndex(v, x, y)); }

// ---------------- sampler ----------------
enum Wrap { WRAP_CLAMP = 0, WRAP_BORDER = 1, WRAP_REPEAT = 2 };
PV_HD float sanitizeCoord(float u) {
#if defined(__CUDA_ARCH__)
    return isnanf_(u) ? 0.f : fminf(fmaxf(u, -65536.f), 65536.f);  // two FMNMX + select; the bounds are non-zero, so the +-0 rule of fminp/fmaxp cannot matter
#else
    return isnanf_(u) ? 0.f : clampf(u, -65536.f, 65536.f);
#endif
}
// f2i(floor(x)) in one conversion on the device: cvt.rmi.s32.f32 rounds down, saturates, NaN -> 0 (and -0 -> 0)
PV_HD int floor2i(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rd(x);
#else
    return f2i(floorf_(x));
#endif
}
template <int WRAP> PV_HD bool wrapIndex(int& i, int size) {
    if (WRAP == WRAP_CLAMP) { i = iclamp(i, 0, size - 1); return true; }
    if (WRAP == WRAP_REPEAT) { i %= size; if (i < 0) i += size; return true; }
    return i >= 0 && i < size;
}
struct Bilerp { int x0, y0; float w00, w10, w01, w11; };
PV_HD Bilerp bilerpSetup(vec2 uv, int w, int h) {
    Bilerp b;
    const float fx = fmaf_(sanitizeCoord(uv.x), (float)w, -0.5f), fy = fmaf_(sanitizeCoord(uv.y), (float)h, -0.5f);
    const float x0f = floorf_(fx), y0f = floorf_(fy);
    const float ax = fx - x0f, ay = fy - y0f, bx = 1.f - ax, by = 1.f - ay;
    b.x0 = f2i(x0f); b.y0 = f2i(y0f);
    b.w00 = bx * by; b.w10 = ax * by; b.w01 = bx * ay; b.w11 = ax * ay;
    return b;
}
PV_HD ivec2 nearestTexel(vec2 uv, int w, int h) {
    ivec2 r;
    r.x = floor2i(sanitizeCoord(uv.x) * (float)w);
    r.y = floor2i(sanitizeCoord(uv.y) * (float)h);
    return r;
}

// generic 2-D sampling over a fetch functor F(x, y) -> T (texel already decoded); border value given by the caller
template <int WRAP, typename T, typename F> PV_HD T sampleLinear2D(F fetch, int w, int h, vec2 uv, T border) {
    const Bilerp b = bilerpSetup(uv, w, h);
    int x0 = b.x0, x1 = b.x0 + 1, y0 = b.y0, y1 = b.y0 + 1;
    const bool okx0 = wrapIndex<WRAP>(x0, w), okx1 = wrapIndex<WRAP>(x1, w), oky0 = wrapIndex<WRAP>(y0, h), oky1 = wrapIndex<WRAP>(y1, h);
    const T t00 = (okx0 && oky0) ? fetch(x0, y0) : border, t10 = (okx1 && oky0) ? fetch(x1, y0) : border;
    const T t01 = (okx0 && oky1) ? fetch(x0, y1) : border, t11 = (okx1 && oky1) ? fetch(x1, y1) : border;
    return vfma(t11, b.w11, vfma(t01, b.w01, vfma(t10, b.w10, t00 * b.w00)));
}
// nearest texel of an UNSANITISED coordinate, for clamp-to-edge and border addressing only: the float -> int conversion saturates
// and maps NaN to 0, so floor2i(u * size) lands on the same side of [0, size) as the sanitised coordinate does for |u| > 65536
// (both far outside: the edge texel / the border) and on texel 0 for a NaN (sanitised: 0 * size = 0). Sizes are below 65536.
// Repeat addressing takes the index modulo the size and keeps the sanitised form.
PV_HD ivec2 nearestTexelUnsanitized(vec2 uv, int w, int h) {
    ivec2 r;
    r.x = floor2i(uv.x * (float)w);
    r.y = floor2i(uv.y * (float)h);
    return r;
}
template <int WRAP, typename T, typename F> PV_HD T sampleNearest2D(F fetch, int w, int h, vec2 uv, T border) {
    ivec2 t = (WRAP == WRAP_REPEAT) ? nearestTexel(uv, w, h) : nearestTexelUnsanitized(uv, w, h);
    const bool ok = wrapIndex<WRAP>(t.x, w) & wrapIndex<WRAP>(t.y, h);
    return ok ? fetch(t.x, t.y) : border;
}
// trilinear over a fetch functor F(x, y, z) -> T: fma(slice1, fz, slice0*(1-fz))
template <int WRAP, typename T, typename F> PV_HD T sampleLinear3D(F fetch, int w, int h, int d, vec3 uvw, T border) {
    const Bilerp b = bilerpSetup(v2(uvw.x, uvw.y), w, h);
    const float fz = fmaf_(sanitizeCoord(uvw.z), (float)d, -0.5f);
    const float z0f = floorf_(fz);
    const float az = fz - z0f, bz = 1.f - az;
    int x0 = b.x0, x1 = b.x0 + 1, y0 = b.y0, y1 = b.y0 + 1, z0 = f2i(z0f), z1 = z0 + 1;
    const bool okx0 = wrapIndex<WRAP>(x0, w), okx1 = wrapIndex<WRAP>(x1, w), oky0 = wrapIndex<WRAP>(y0, h), oky1 = wrapIndex<WRAP>(y1, h);
    const bool okz0 = wrapIndex<WRAP>(z0, d), okz1 = wrapIndex<WRAP>(z1, d);
    const T a00 = (okx0 && oky0 && okz0) ? fetch(x0, y0, z0) : border, a10 = (okx1 && oky0 && okz0) ? fetch(x1, y0, z0) : border;
    const T a01 = (okx0 && oky1 && okz0) ? fetch(x0, y1, z0) : border, a11 = (okx1 && oky1 && okz0) ? fetch(x1, y1, z0) : border;
    const T b00 = (okx0 && oky0 && okz1) ? fetch(x0, y0, z1) : border, b10 = (okx1 && oky0 && okz1) ? fetch(x1, y0, z1) : border;
    const T b01 = (okx0 && oky1 && okz1) ? fetch(x0, y1, z1) : border, b11 = (okx1 && oky1 && okz1) ? fetch(x1, y1, z1) : border;
    const T s0 = vfma(a11, b.w11, vfma(a01, b.w01, vfma(a10, b.w10, a00 * b.w00)));
    const T s1 = vfma(b11, b.w11, vfma(b01, b.w01, vfma(b10, b.w10, b00 * b.w00)));
    return vfma(s1, az, s0 * bz);
}

// common instantiations
PV_HD vec3 sampleR11LinearClamp(const ImgView& v, vec2 uv) {
    return sampleLinear2D<WRAP_CLAMP, vec3>([&](int x, int y) { return loadR11(v, x, y); }, v.w, v.h, uv, v3(0.f));
}
PV_HD vec3 sampleR11LinearRepeat(const ImgView& v, vec2 uv) {
    return sampleLinear2D<WRAP_REPEAT, vec3>([&](int x, int y) { return loadR11(v, x, y); }, v.w, v.h, uv, v3(0.f));
}
PV_HD vec4 sampleRGBA16FLinearClamp(const ImgView& v, vec2 uv) {
    return sampleLinear2D<WRAP_CLAMP, vec4>([&](int x, int y) { return loadRGBA16F(v, x, y); }, v.w, v.h, uv, v4(0.f));
}
PV_HD vec2 sampleRG16FLinearClamp(const ImgView& v, vec2 uv) {
    return sampleLinear2D<WRAP_CLAMP, vec2>([&](int x, int y) { return loadRG16F(v, x, y); }, v.w, v.h, uv, v2(0.f));
}
PV_HD vec4 sampleRGBA16FLinearClamp3D(const ImgView& v, vec3 uvw) {
    return sampleLinear3D<WRAP_CLAMP, vec4>([&](int x, int y, int z) { return loadRGBA16F(v, x, y, z); }, v.w, v.h, v.d, uvw, v4(0.f));
}
PV_HD float sampleR16FLinearClamp3D(const ImgView& v, vec3 uvw) {
    return sampleLinear3D<WRAP_CLAMP, float>([&](int x, int y, int z) { return loadR16F(v, x, y, z); }, v.w, v.h, v.d, uvw, 0.f);
}

}  // namespace pv
