// tile.cuh - shared-memory staging of 2-D tiles of packed images for the stencil-shaped passes (TAA, bloom).
// A block decodes every texel of its footprint ONCE into shared memory (float4 per texel, so a tap corner is one
// LDS.128) instead of once per bilinear tap corner; taps whose corners fall outside the staged rectangle (large motion
// vectors) read global memory through the same decode, so the arithmetic - and the result - is identical either way.
#pragma once
#include "pass_common.cuh"

namespace pb {

template <int TW, int TH>
struct TileR11 {
    const float4* s;
    int x0, y0;  // absolute texel rectangle [x0, x0 + TW) x [y0, y0 + TH); entries outside the image are never read
};

// all threads of the block; caller synchronises
template <int TW, int TH>
__device__ __forceinline__ void tileLoadR11(float4* smem, const ImgView& img, int x0, int y0) {
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) {
        const int tx = i % TW, ty = i / TW;
        const int ax = x0 + tx, ay = y0 + ty;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ax >= 0 && ay >= 0 && ax < img.w && ay < img.h) {
            const vec3 c = unpackR11G11B10(__ldg((const uint32_t*)img.ptr + (size_t)ay * img.w + ax));
            v = make_float4(c.x, c.y, c.z, 0.f);
        }
        smem[i] = v;
    }
}

// the four corners of a tap from global memory (taps that leave the staged rectangle); kept out of line so the many
// tap sites of a kernel share one copy of the decode
__device__ __noinline__ void tapCornersGlobalR11(const ImgView& img, int x0, int x1, int y0, int y1, vec3& t00, vec3& t10, vec3& t01, vec3& t11) {
    t00 = loadR11(img, x0, y0); t10 = loadR11(img, x1, y0);
    t01 = loadR11(img, x0, y1); t11 = loadR11(img, x1, y1);
}

// texture(sampler2D, uv) with the linear + clamp-to-edge sampler (image_view.h sampleLinear2D<WRAP_CLAMP>: same setup,
// same blend order); the 2x2 footprint is served from the tile when it lies inside it
template <int TW, int TH>
__device__ __forceinline__ vec3 sampleR11LinearClampTile(const TileR11<TW, TH>& t, const ImgView& img, vec2 uv) {
    const Bilerp b = bilerpSetup(uv, img.w, img.h);
    const int x0 = iclamp(b.x0, 0, img.w - 1), x1 = iclamp(b.x0 + 1, 0, img.w - 1);
    const int y0 = iclamp(b.y0, 0, img.h - 1), y1 = iclamp(b.y0 + 1, 0, img.h - 1);
    const unsigned lx0 = (unsigned)(x0 - t.x0), lx1 = (unsigned)(x1 - t.x0), ly0 = (unsigned)(y0 - t.y0), ly1 = (unsigned)(y1 - t.y0);
    vec3 t00, t10, t01, t11;
    if (lx0 < (unsigned)TW && lx1 < (unsigned)TW && ly0 < (unsigned)TH && ly1 < (unsigned)TH) {
        const float4 a = t.s[ly0 * TW + lx0], bq = t.s[ly0 * TW + lx1], c = t.s[ly1 * TW + lx0], d = t.s[ly1 * TW + lx1];
        t00 = v3(a.x, a.y, a.z); t10 = v3(bq.x, bq.y, bq.z); t01 = v3(c.x, c.y, c.z); t11 = v3(d.x, d.y, d.z);
    } else {
        tapCornersGlobalR11(img, x0, x1, y0, y1, t00, t10, t01, t11);
    }
    return vfma(t11, b.w11, vfma(t01, b.w01, vfma(t10, b.w10, t00 * b.w00)));
}

// ---- separable set-up for a 3x3 block of taps at uv + texelSize * (i - 1, j - 1) ----
// The bilinear set-up of image_view.h bilerpSetup works per axis: the coordinate, its floor, the two weights and the two
// clamped texel indices of the x axis depend only on i, those of the y axis only on j. Computing the three x axes and the
// three y axes once (6 set-ups instead of 18) and combining them per tap evaluates the same expressions on the same
// operands as nine calls of sampleR11LinearClampTile - the results are identical.
struct AxisTap {
    float a, b;      // weight of texel i1 / i0 (ax, 1 - ax)
    unsigned l0, l1; // clamped texel indices relative to the tile origin (unsigned: one compare tests the range)
    int i0, i1;      // clamped absolute texel indices (global fallback)
};
__device__ __forceinline__ AxisTap axisTap(float u, int size, int tileOrigin) {
    AxisTap t;
    const float f = fmaf_(sanitizeCoord(u), (float)size, -0.5f);
    const float f0 = floorf_(f);
    t.a = f - f0;
    t.b = 1.f - t.a;
    const int i = f2i(f0);
    t.i0 = iclamp(i, 0, size - 1);
    t.i1 = iclamp(i + 1, 0, size - 1);
    t.l0 = (unsigned)(t.i0 - tileOrigin);
    t.l1 = (unsigned)(t.i1 - tileOrigin);
    return t;
}
template <int TW, int TH>
__device__ __forceinline__ vec3 tapR11Tile(const TileR11<TW, TH>& t, const ImgView& img, const AxisTap& X, const AxisTap& Y) {
    const float w00 = X.b * Y.b, w10 = X.a * Y.b, w01 = X.b * Y.a, w11 = X.a * Y.a;
    vec3 t00, t10, t01, t11;
    if (X.l0 < (unsigned)TW && X.l1 < (unsigned)TW && Y.l0 < (unsigned)TH && Y.l1 < (unsigned)TH) {
        const float4 a = t.s[Y.l0 * TW + X.l0], bq = t.s[Y.l0 * TW + X.l1], c = t.s[Y.l1 * TW + X.l0], d = t.s[Y.l1 * TW + X.l1];
        t00 = v3(a.x, a.y, a.z); t10 = v3(bq.x, bq.y, bq.z); t01 = v3(c.x, c.y, c.z); t11 = v3(d.x, d.y, d.z);
    } else {
        tapCornersGlobalR11(img, X.i0, X.i1, Y.i0, Y.i1, t00, t10, t01, t11);
    }
    return vfma(t11, w11, vfma(t01, w01, vfma(t10, w10, t00 * w00)));
}

}  // namespace pb
