// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// image.h - texel formats, images with mip chains and a software sampler that states Vulkan's texel
// addressing rules explicitly (the driver-defined parts are pinned here, SURVEY.md 8c / DESIGN.md):
//   * float -> UNORM8: uint(clamp(x,0,1)*255 + 0.5); float -> half / R11G11B10: round to nearest even,
//     negative -> 0, > max finite -> max finite (R11G11B10) / inf (half), NaN kept
//   * nearest: texel = floor(u*size); linear: x = u*size - 0.5, i0 = floor(x), f = x - i0 in full fp32,
//     (x = fma(u, size, -0.5)), result = fma(t11, fx*fy, fma(t01, (1-fx)*fy, fma(t10, fx*(1-fy), t00*((1-fx)*(1-fy)))))
//   * clamp-to-edge clamps the integer texel index; repeat wraps it; border returns opaque black/white
//   * texelFetch/imageLoad out of range -> 0; imageStore out of range is dropped
// Formats follow VulkanImageFormats.cpp:4-25 (R11G11B10 = B10G11R11_UFLOAT_PACK32: R bits 0-10, G 11-21, B 22-31).
#pragma once
#include <vector>
#include <stdint.h>
#include <string.h>
#include "glsl.h"
#include "../include/plain_b200.h"

namespace orc {
using namespace gl;

// ---------------- scalar format conversions ----------------
inline uint32_t floatToSmallFloat(float f, int mbits, bool hasSign, bool overflowToInf) {
    uint32_t u = dm::f2u(f);
    uint32_t sign = u >> 31;
    uint32_t au = u & 0x7fffffffu;
    uint32_t signBits = hasSign ? (sign << (5 + mbits)) : 0u;
    const uint32_t expAll = 31u << mbits;
    const uint32_t maxFinite = (30u << mbits) | ((1u << mbits) - 1u);
    if (au > 0x7f800000u) return signBits | expAll | 1u;  // NaN
    if (!hasSign && sign) return 0u;                       // negative -> 0
    if (au == 0x7f800000u) return signBits | expAll;       // inf
    int e = (int)(au >> 23) - 127;
    uint32_t m = au & 0x7fffffu;
    if (e > 15) return signBits | (overflowToInf ? expAll : maxFinite);
    uint32_t value;
    if (e >= -14) {
        int shift = 23 - mbits;
        uint32_t mant = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        value = ((uint32_t)(e + 15) << mbits) + mant;
        if (rem > half || (rem == half && (mant & 1u))) value += 1u;
    } else {
        if ((au >> 23) == 0) return signBits;  // float denormal -> 0
        uint32_t full = m | 0x800000u;
        int shift = (23 - mbits) + (-14 - e);
        if (shift > 25) return signBits;
        uint32_t mant = full >> shift;
        uint32_t rem = full & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        value = mant;
        if (rem > half || (rem == half && (mant & 1u))) value += 1u;
    }
    if (value >= expAll) value = overflowToInf ? expAll : maxFinite;
    return signBits | value;
}

inline float smallFloatToFloat(uint32_t v, int mbits, bool hasSign) {
    uint32_t sign = hasSign ? ((v >> (5 + mbits)) & 1u) : 0u;
    uint32_t e = (v >> mbits) & 31u;
    uint32_t m = v & ((1u << mbits) - 1u);
    uint32_t out;
    if (e == 0) {
        if (m == 0) {
            out = 0;
        } else {  // denormal: m * 2^-14 / 2^mbits ; normalise
            int sh = 0;
            while (!(m & (1u << mbits))) { m <<= 1; sh++; }
            m &= (1u << mbits) - 1u;
            out = ((uint32_t)(127 - 14 - sh) << 23) | (m << (23 - mbits));
        }
    } else if (e == 31) {
        out = 0x7f800000u | (m << (23 - mbits));
    } else {
        out = ((e + 127 - 15) << 23) | (m << (23 - mbits));
    }
    return dm::u2f(out | (sign << 31));
}

inline uint16_t floatToHalf(float f) { return isnan(f) ? (uint16_t)0x7fffu : (uint16_t)floatToSmallFloat(f, 10, true, true); }  // NaN -> 0x7fff (what cvt.rn.f16.f32 gives)
inline float halfToFloat(uint16_t h) { return smallFloatToFloat(h, 10, true); }
inline uint32_t packR11G11B10(vec3 c) {
    return floatToSmallFloat(c.x, 6, false, false) | (floatToSmallFloat(c.y, 6, false, false) << 11) | (floatToSmallFloat(c.z, 5, false, false) << 22);
}
inline vec3 unpackR11G11B10(uint32_t v) {
    return vec3(smallFloatToFloat(v & 0x7ffu, 6, false), smallFloatToFloat((v >> 11) & 0x7ffu, 6, false), smallFloatToFloat((v >> 22) & 0x3ffu, 5, false));
}
inline uint8_t floatToUnorm8(float x) {
    if (isnan(x)) return 0;
    return (uint8_t)(uint32_t)(clamp(x, 0.f, 1.f) * 255.f + 0.5f);
}
inline float unorm8ToFloat(uint8_t v) { return (float)v / 255.f; }
inline float unorm16ToFloat(uint16_t v) { return (float)v / 65535.f; }
inline float snorm16ToFloat(int16_t v) { return max((float)v / 32767.f, -1.f); }

inline int formatBytesPerTexel(uint32_t fmt) {
    switch (fmt) {
        case PLAIN_FORMAT_R8: return 1;
        case PLAIN_FORMAT_RG8: return 2;
        case PLAIN_FORMAT_RGBA8: return 4;
        case PLAIN_FORMAT_R16_SFLOAT: return 2;
        case PLAIN_FORMAT_RG16_SFLOAT: return 4;
        case PLAIN_FORMAT_RG32_SFLOAT: return 8;
        case PLAIN_FORMAT_RG16_SNORM: return 4;
        case PLAIN_FORMAT_RGBA16_SFLOAT: return 8;
        case PLAIN_FORMAT_RGBA16_SNORM: return 8;
        case PLAIN_FORMAT_RGBA32_SFLOAT: return 16;
        case PLAIN_FORMAT_R11G11B10_UFLOAT: return 4;
        case PLAIN_FORMAT_DEPTH16: return 2;
        case PLAIN_FORMAT_DEPTH32: return 4;
        case PLAIN_FORMAT_BGRA8_UNORM: return 4;
        case PLAIN_FORMAT_RGBA32_UINT: return 16;
        default: return 0;  // BCn: not supported on the frame path
    }
}

// ---------------- images ----------------
struct MipLevel {
    int w = 0, h = 0, d = 0;
    std::vector<uint8_t> data;
};

struct Image {
    plain_image_desc desc{};
    std::vector<MipLevel> mips;
    bool inUse = false;  // transient pool
    bool transparentTexels = false;  // RGBA8 created with initial data containing a texel of alpha < 255 (alpha test of the raster passes)

    static int computeMipCount(const plain_image_desc& d) {
        if (d.mip_count == PLAIN_MIPS_ONE) return 1;
        if (d.mip_count == PLAIN_MIPS_MANUAL) return (int)d.manual_mip_count;
        uint32_t m = d.width > d.height ? d.width : d.height;
        if (d.depth > m) m = d.depth;
        int n = 1;
        while (m > 1) { m >>= 1; n++; }  // 1 + floor(log2(max)), MathUtils.cpp:17-19
        return n;
    }
    void allocate(const plain_image_desc& d) {
        desc = d;
        int n = computeMipCount(d);
        mips.assign(n, MipLevel());
        int bpt = formatBytesPerTexel(d.format);
        for (int i = 0; i < n; i++) {
            mips[i].w = max((int)d.width >> i, 1);
            mips[i].h = max((int)d.height >> i, 1);
            mips[i].d = max((int)(d.depth ? d.depth : 1) >> i, 1);
            mips[i].data.assign((size_t)mips[i].w * mips[i].h * mips[i].d * bpt, 0);
        }
    }
};

// a view of one mip level ("ImageResource{image, mipLevel, binding}", ResourceDescriptions.h:29-37)
struct View {
    Image* img = nullptr;
    int mip = 0;
    bool valid() const { return img != nullptr; }
    int w() const { return img->mips[mip].w; }
    int h() const { return img->mips[mip].h; }
    int d() const { return img->mips[mip].d; }
    uint32_t format() const { return img->desc.format; }
    uint8_t* texelPtr(int x, int y, int z) const {
        MipLevel& m = img->mips[mip];
        return m.data.data() + ((size_t)(z * m.h + y) * m.w + x) * formatBytesPerTexel(img->desc.format);
    }
    bool inRange(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < w() && y < h() && z < d(); }

    vec4 load(int x, int y, int z = 0) const {  // in range
        const uint8_t* p = texelPtr(x, y, z);
        switch (format()) {
            case PLAIN_FORMAT_R8: return vec4(unorm8ToFloat(p[0]), 0, 0, 1);
            case PLAIN_FORMAT_RG8: return vec4(unorm8ToFloat(p[0]), unorm8ToFloat(p[1]), 0, 1);
            case PLAIN_FORMAT_RGBA8: return vec4(unorm8ToFloat(p[0]), unorm8ToFloat(p[1]), unorm8ToFloat(p[2]), unorm8ToFloat(p[3]));
            case PLAIN_FORMAT_BGRA8_UNORM: return vec4(unorm8ToFloat(p[2]), unorm8ToFloat(p[1]), unorm8ToFloat(p[0]), unorm8ToFloat(p[3]));
            case PLAIN_FORMAT_R16_SFLOAT: { uint16_t v; memcpy(&v, p, 2); return vec4(halfToFloat(v), 0, 0, 1); }
            case PLAIN_FORMAT_RG16_SFLOAT: { uint16_t v[2]; memcpy(v, p, 4); return vec4(halfToFloat(v[0]), halfToFloat(v[1]), 0, 1); }
            case PLAIN_FORMAT_RGBA16_SFLOAT: { uint16_t v[4]; memcpy(v, p, 8); return vec4(halfToFloat(v[0]), halfToFloat(v[1]), halfToFloat(v[2]), halfToFloat(v[3])); }
            case PLAIN_FORMAT_RG32_SFLOAT: { float v[2]; memcpy(v, p, 8); return vec4(v[0], v[1], 0, 1); }
            case PLAIN_FORMAT_RGBA32_SFLOAT: { float v[4]; memcpy(v, p, 16); return vec4(v[0], v[1], v[2], v[3]); }
            case PLAIN_FORMAT_RG16_SNORM: { int16_t v[2]; memcpy(v, p, 4); return vec4(snorm16ToFloat(v[0]), snorm16ToFloat(v[1]), 0, 1); }
            case PLAIN_FORMAT_R11G11B10_UFLOAT: { uint32_t v; memcpy(&v, p, 4); return vec4(unpackR11G11B10(v), 1); }
            case PLAIN_FORMAT_DEPTH16: { uint16_t v; memcpy(&v, p, 2); return vec4(unorm16ToFloat(v), 0, 0, 1); }
            case PLAIN_FORMAT_DEPTH32: { float v; memcpy(&v, p, 4); return vec4(v, 0, 0, 1); }
            default: return vec4(0);
        }
    }
    void loadUint4(int x, int y, uint32_t out[4]) const { memcpy(out, texelPtr(x, y, 0), 16); }

    // imageStore: dropped when out of range
    void store(int x, int y, int z, vec4 c) const {
        if (!inRange(x, y, z)) return;
        uint8_t* p = texelPtr(x, y, z);
        switch (format()) {
            case PLAIN_FORMAT_R8: p[0] = floatToUnorm8(c.x); break;
            case PLAIN_FORMAT_RG8: p[0] = floatToUnorm8(c.x); p[1] = floatToUnorm8(c.y); break;
            case PLAIN_FORMAT_RGBA8: p[0] = floatToUnorm8(c.x); p[1] = floatToUnorm8(c.y); p[2] = floatToUnorm8(c.z); p[3] = floatToUnorm8(c.w); break;
            case PLAIN_FORMAT_BGRA8_UNORM: p[2] = floatToUnorm8(c.x); p[1] = floatToUnorm8(c.y); p[0] = floatToUnorm8(c.z); p[3] = floatToUnorm8(c.w); break;
            case PLAIN_FORMAT_R16_SFLOAT: { uint16_t v = floatToHalf(c.x); memcpy(p, &v, 2); break; }
            case PLAIN_FORMAT_RG16_SFLOAT: { uint16_t v[2] = {floatToHalf(c.x), floatToHalf(c.y)}; memcpy(p, v, 4); break; }
            case PLAIN_FORMAT_RGBA16_SFLOAT: { uint16_t v[4] = {floatToHalf(c.x), floatToHalf(c.y), floatToHalf(c.z), floatToHalf(c.w)}; memcpy(p, v, 8); break; }
            case PLAIN_FORMAT_RG32_SFLOAT: { float v[2] = {c.x, c.y}; memcpy(p, v, 8); break; }
            case PLAIN_FORMAT_RGBA32_SFLOAT: { float v[4] = {c.x, c.y, c.z, c.w}; memcpy(p, v, 16); break; }
            case PLAIN_FORMAT_R11G11B10_UFLOAT: { uint32_t v = packR11G11B10(c.xyz()); memcpy(p, &v, 4); break; }
            default: break;
        }
    }
    void store(ivec2 uv, vec4 c) const { store(uv.x, uv.y, 0, c); }
    // texelFetch / imageLoad: 0 when out of range
    vec4 fetch(int x, int y, int z = 0) const { return inRange(x, y, z) ? load(x, y, z) : vec4(0); }
    vec4 fetch(ivec2 uv) const { return fetch(uv.x, uv.y, 0); }
};

// ---------------- samplers (RenderFrontend.cpp:1300-1397, VulkanSampler.cpp:4-36) ----------------
struct Sampler {
    bool linear;
    int wrap;  // PLAIN_WRAP_*
    bool borderWhite;
};
static const Sampler s_anisotropicRepeat = {true, PLAIN_WRAP_REPEAT, true};
static const Sampler s_nearestBlackBorder = {false, PLAIN_WRAP_COLOR, false};
static const Sampler s_linearRepeat = {true, PLAIN_WRAP_REPEAT, true};
static const Sampler s_linearClamp = {true, PLAIN_WRAP_CLAMP, true};
static const Sampler s_nearestClamp = {false, PLAIN_WRAP_CLAMP, false};
static const Sampler s_linearWhiteBorder = {true, PLAIN_WRAP_COLOR, true};
static const Sampler s_nearestRepeat = {false, PLAIN_WRAP_REPEAT, false};
static const Sampler s_nearestWhiteBorder = {false, PLAIN_WRAP_COLOR, true};

inline bool wrapIndex(int& i, int size, int wrap) {  // false -> border
    if (wrap == PLAIN_WRAP_CLAMP) { i = clamp(i, 0, size - 1); return true; }
    if (wrap == PLAIN_WRAP_REPEAT) { i %= size; if (i < 0) i += size; return true; }
    return i >= 0 && i < size;
}
inline vec4 sampleTexel(const View& v, const Sampler& s, int x, int y, int z) {
    bool ok = wrapIndex(x, v.w(), s.wrap);
    ok = wrapIndex(y, v.h(), s.wrap) && ok;
    ok = wrapIndex(z, v.d(), s.wrap) && ok;
    if (!ok) return s.borderWhite ? vec4(1, 1, 1, 1) : vec4(0, 0, 0, 1);
    return v.load(x, y, z);
}

inline ivec2 textureSize(const View& v) { return ivec2(v.w(), v.h()); }

// texture coordinates entering the sampler: NaN -> 0 and the magnitude is limited to 65536 (GPUs convert to fixed point;
// the frame path needs this on its first frame: volumeLightingReprojection.comp:39-46 divides by w = 0 while
// g_viewProjectionPrevious is still the zero matrix)
inline float sanitizeCoord(float u) { return isnan(u) ? 0.f : clamp(u, -65536.f, 65536.f); }

// texture(sampler2D, uv) / textureLod(..., 0) on the bound mip view
inline vec4 texture(const View& v, const Sampler& s, vec2 uv) {
    uv = vec2(sanitizeCoord(uv.x), sanitizeCoord(uv.y));
    if (!s.linear) {
        int x = f2int(floor(uv.x * (float)v.w()));
        int y = f2int(floor(uv.y * (float)v.h()));
        return sampleTexel(v, s, x, y, 0);
    }
    float fx = fma_(uv.x, (float)v.w(), -0.5f);
    float fy = fma_(uv.y, (float)v.h(), -0.5f);
    float x0f = floor(fx), y0f = floor(fy);
    float ax = fx - x0f, ay = fy - y0f;
    int x0 = f2int(x0f), y0 = f2int(y0f);
    vec4 t00 = sampleTexel(v, s, x0, y0, 0), t10 = sampleTexel(v, s, x0 + 1, y0, 0);
    vec4 t01 = sampleTexel(v, s, x0, y0 + 1, 0), t11 = sampleTexel(v, s, x0 + 1, y0 + 1, 0);
    float bx = 1.f - ax, by = 1.f - ay;
    return vfma(t11, ax * ay, vfma(t01, bx * ay, vfma(t10, ax * by, t00 * (bx * by))));
}

// texture(sampler3D, uvw): trilinear = lerp of the two bilinear slices in z:
// r = fma(slice1, fz, slice0*(1-fz)) with each slice as in the 2D rule
inline vec4 texture3D(const View& v, const Sampler& s, vec3 uvw) {
    uvw = vec3(sanitizeCoord(uvw.x), sanitizeCoord(uvw.y), sanitizeCoord(uvw.z));
    if (!s.linear) {
        int x = f2int(floor(uvw.x * (float)v.w()));
        int y = f2int(floor(uvw.y * (float)v.h()));
        int z = f2int(floor(uvw.z * (float)v.d()));
        return sampleTexel(v, s, x, y, z);
    }
    float fx = fma_(uvw.x, (float)v.w(), -0.5f);
    float fy = fma_(uvw.y, (float)v.h(), -0.5f);
    float fz = fma_(uvw.z, (float)v.d(), -0.5f);
    float x0f = floor(fx), y0f = floor(fy), z0f = floor(fz);
    float ax = fx - x0f, ay = fy - y0f, az = fz - z0f;
    int x0 = f2int(x0f), y0 = f2int(y0f), z0 = f2int(z0f);
    float bx = 1.f - ax, by = 1.f - ay, bz = 1.f - az;
    vec4 sl[2];
    for (int k = 0; k < 2; k++) {
        vec4 t00 = sampleTexel(v, s, x0, y0, z0 + k), t10 = sampleTexel(v, s, x0 + 1, y0, z0 + k);
        vec4 t01 = sampleTexel(v, s, x0, y0 + 1, z0 + k), t11 = sampleTexel(v, s, x0 + 1, y0 + 1, z0 + k);
        sl[k] = vfma(t11, ax * ay, vfma(t01, bx * ay, vfma(t10, ax * by, t00 * (bx * by))));
    }
    return vfma(sl[1], az, sl[0] * bz);
}

// textureGather(sampler2D, uv) component 0: (i0,j1), (i1,j1), (i1,j0), (i0,j0) with i0 = floor(u*w - 0.5)
inline vec4 textureGather(const View& v, const Sampler& s, vec2 uv) {
    uv = vec2(sanitizeCoord(uv.x), sanitizeCoord(uv.y));
    int x0 = f2int(floor(uv.x * (float)v.w() - 0.5f));
    int y0 = f2int(floor(uv.y * (float)v.h() - 0.5f));
    return vec4(sampleTexel(v, s, x0, y0 + 1, 0).x, sampleTexel(v, s, x0 + 1, y0 + 1, 0).x, sampleTexel(v, s, x0 + 1, y0, 0).x, sampleTexel(v, s, x0, y0, 0).x);
}

}  // namespace orc
