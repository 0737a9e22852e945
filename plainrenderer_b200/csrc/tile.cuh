// tile.cuh - shared-memory staging of 2-D tiles of packed images for the stencil-shaped passes (TAA, bloom).
// A block decodes every texel of its footprint ONCE into shared memory (float4 per texel, so a tap corner is one
// LDS.128) instead of once per bilinear tap corner; taps whose corners fall outside the staged rectangle (large motion
// vectors) read global memory through the same decode, so the arithmetic - and the result - is identical either way.
//
// Round 2 (instruction diet inside the bit-exact contract):
//   * interior blocks (the staged rectangle lies inside the image) skip every clamp and bounds test: clamp-to-edge is the
//     identity on an index that is inside the image, so the operands - and the bits - are the same
//   * coordinates that are finite and bounded by construction (pixel centres, + SNORM16 motion, + fixed stencil offsets)
//     are not sanitised: sanitizeCoord is the identity on them
//   * floor + float->int of a coordinate is ONE conversion (cvt.rmi) and the floor as a float is the int converted back
//     (exact below 2^24; floor_(-0) = +0 and (float)0 = +0 agree)
//   * a 3x3 block of taps one texel apart touches a 4x4 texel window: it is read once, row by row, instead of 36 corner
//     fetches (window3x3); the per-tap blend is the sampler's own fma chain on the same operands
#pragma once
#include "pass_common.cuh"

namespace pb {

template <int TW, int TH>
struct TileR11 {
    const float4* s;
    int x0, y0;  // absolute texel rectangle [x0, x0 + TW) x [y0, y0 + TH); entries outside the image are zero and never used
};

__device__ __forceinline__ float4 decodeR11Texel(uint32_t v) {
    const vec3 c = unpackR11G11B10(v);
    return make_float4(c.x, c.y, c.z, 0.f);
}

// a channel with all exponent bits set is inf or NaN (the fast paths of the stencil kernels exclude tiles that hold one)
__device__ __forceinline__ bool r11TexelIsSpecial(uint32_t v) {
    const uint32_t u = ~v;
    return (u & 0x7c0u) == 0u || (u & 0x3e0000u) == 0u || (u & 0xf8000000u) == 0u;
}
// All threads of a 256-thread block stage the rectangle [x0, x0 + TW) x [y0, y0 + TH) of img; the caller synchronises.
// Thread t owns column t % TW of rows t / TW, t / TW + 256 / TW, ...: one division per tile instead of one per texel, constant
// strides from row to row. INTERIOR: the rectangle lies inside the image (no bounds tests). Returns whether one of the texels
// this thread staged holds an inf / NaN channel.
// SWIZZLE: a row stores its even columns first, then its odd columns (tileSwizzle). A 2:1 downsampling stencil reads columns two
// apart in neighbouring lanes: with the plain layout a quarter warp's 128-bit reads hit every bank twice, with this one they are
// consecutive. The taps must then address columns through tileSwizzle too.
template <int TW> __device__ __forceinline__ unsigned tileSwizzle(unsigned column) { return (column >> 1) + (column & 1u) * (unsigned)(TW / 2); }
template <int TW, int TH, bool INTERIOR = false, bool SWIZZLE = false>
__device__ __forceinline__ bool tileLoadR11(float4* smem, const ImgView& img, int x0, int y0) {
    static_assert(!SWIZZLE || (TW % 2) == 0, "swizzled tiles have an even width");
    constexpr int RPP = 256 / TW;  // rows per pass
    static_assert(RPP >= 1, "tile wider than the block");
    const int tx = (int)threadIdx.x % TW, tyBase = (int)threadIdx.x / TW;
    bool special = false;
    if (tyBase < RPP) {
        const int ax = x0 + tx;
        const bool colOk = INTERIOR || (ax >= 0 && ax < img.w);
        const uint32_t* src = (const uint32_t*)img.ptr + (y0 + tyBase) * img.w + ax;
        float4* dst = smem + tyBase * TW + (SWIZZLE ? (int)tileSwizzle<TW>((unsigned)tx) : tx);
#pragma unroll
        for (int r = 0; r < (TH + RPP - 1) / RPP; r++) {
            const int ty = tyBase + r * RPP;
            if (ty < TH) {
                const int ay = y0 + ty;
                uint32_t v = 0u;  // decodes to (0, 0, 0)
                if (INTERIOR || (colOk && ay >= 0 && ay < img.h)) v = __ldg(src + r * RPP * img.w);
                special = special || r11TexelIsSpecial(v);
                dst[r * RPP * TW] = decodeR11Texel(v);
            }
        }
    }
    return special;
}
template <int TW, int TH>
__device__ __forceinline__ bool tileIsInterior(const ImgView& img, int x0, int y0) { return x0 >= 0 && y0 >= 0 && x0 + TW <= img.w && y0 + TH <= img.h; }

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

// ---- separable set-up: one axis of a bilinear tap (image_view.h bilerpSetup works per axis) ----
// The coordinate, its floor, the two weights and the two clamped texel indices of the x axis of a tap depend only on the
// tap's x coordinate, those of the y axis only on y: a stencil with m distinct x and n distinct y coordinates needs m + n
// set-ups instead of 2 per tap, evaluating the same expressions on the same operands.
struct AxisTap {
    float a, b;      // weight of texel i1 / i0 (ax, 1 - ax)
    unsigned l0, l1; // clamped texel indices relative to the tile origin (unsigned: one compare tests the range)
    int i0, i1;      // clamped absolute texel indices (global fallback)
};
// SANITIZE = false: the caller guarantees a finite coordinate with |u| <= 65536 (sanitizeCoord is then the identity)
// INTERIOR = true: the caller's tile lies inside the image, so an index inside the tile needs no clamp; the clamped absolute
//                  indices are still produced for the fallback of taps that leave the tile
template <bool SANITIZE, bool INTERIOR>
__device__ __forceinline__ AxisTap axisTapT(float u, int size, int tileOrigin) {
    AxisTap t;
    const float f = fmaf_(SANITIZE ? sanitizeCoord(u) : u, (float)size, -0.5f);
    const int i = floor2i(f);          // |f| < 2^31: the same integer as f2i(floorf_(f))
    const float f0 = (float)i;         // == floorf_(f) (+0 for -0, exact below 2^24)
    t.a = f - f0;
    t.b = 1.f - t.a;
    if (INTERIOR) {
        // unclamped tile-relative indices: whenever both lie inside the tile they are inside the image and equal the clamped ones
        t.l0 = (unsigned)(i - tileOrigin);
        t.l1 = t.l0 + 1u;
        t.i0 = i; t.i1 = i + 1;        // clamped lazily by the fallback (axisClampForFallback)
    } else {
        t.i0 = iclamp(i, 0, size - 1);
        t.i1 = iclamp(i + 1, 0, size - 1);
        t.l0 = (unsigned)(t.i0 - tileOrigin);
        t.l1 = (unsigned)(t.i1 - tileOrigin);
    }
    return t;
}
__device__ __forceinline__ AxisTap axisTap(float u, int size, int tileOrigin) { return axisTapT<true, false>(u, size, tileOrigin); }
template <int TW, int TH>
__device__ __forceinline__ vec3 tapR11Tile(const TileR11<TW, TH>& t, const ImgView& img, const AxisTap& X, const AxisTap& Y) {
    const float w00 = X.b * Y.b, w10 = X.a * Y.b, w01 = X.b * Y.a, w11 = X.a * Y.a;
    vec3 t00, t10, t01, t11;
    if (X.l0 < (unsigned)TW && X.l1 < (unsigned)TW && Y.l0 < (unsigned)TH && Y.l1 < (unsigned)TH) {
        const float4 a = t.s[Y.l0 * TW + X.l0], bq = t.s[Y.l0 * TW + X.l1], c = t.s[Y.l1 * TW + X.l0], d = t.s[Y.l1 * TW + X.l1];
        t00 = v3(a.x, a.y, a.z); t10 = v3(bq.x, bq.y, bq.z); t01 = v3(c.x, c.y, c.z); t11 = v3(d.x, d.y, d.z);
    } else {
        tapCornersGlobalR11(img, iclamp(X.i0, 0, img.w - 1), iclamp(X.i1, 0, img.w - 1), iclamp(Y.i0, 0, img.h - 1), iclamp(Y.i1, 0, img.h - 1), t00, t10, t01, t11);
    }
    return vfma(t11, w11, vfma(t01, w01, vfma(t10, w10, t00 * w00)));
}
// the same tap when the caller has already established that both axes lie inside the tile
template <int TW, int TH>
__device__ __forceinline__ vec3 tapR11TileInside(const TileR11<TW, TH>& t, const AxisTap& X, const AxisTap& Y) {
    const float w00 = X.b * Y.b, w10 = X.a * Y.b, w01 = X.b * Y.a, w11 = X.a * Y.a;
    const float4 a = t.s[Y.l0 * TW + X.l0], bq = t.s[Y.l0 * TW + X.l1], c = t.s[Y.l1 * TW + X.l0], d = t.s[Y.l1 * TW + X.l1];
    return vfma(v3(d.x, d.y, d.z), w11, vfma(v3(c.x, c.y, c.z), w01, vfma(v3(bq.x, bq.y, bq.z), w10, v3(a.x, a.y, a.z) * w00)));
}
template <int N> __device__ __forceinline__ bool axesInside(const AxisTap (&A)[N], unsigned extent) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; i++) ok = ok && A[i].l0 < extent && A[i].l1 < extent;
    return ok;
}

// ---- 3x3 block of taps one texel apart (sampleNeighbourhood, temporalReprojection.inc:42-52) ----
// X[i] / Y[j] are the axis set-ups of the taps at uv + texelSize * (i - 1, j - 1). When the three x axes address consecutive
// texel pairs (X[i].l0 == X[0].l0 + i, l1 == l0 + 1; the same in y) and everything lies inside the tile, the nine taps touch a
// 4x4 window that is read once, row by row (16 LDS.128 instead of 36); otherwise every tap is fetched on its own (tapR11Tile,
// out of line: taps that leave the tile are rare). Either way tap (i, j) is fma(t11, w11, fma(t01, w01, fma(t10, w10, t00 * w00)))
// of the same four texels and weights.
template <int TW, int TH>
__device__ __noinline__ void taps3x3Generic(const float4* tileData, int tileX0, int tileY0, const ImgView& img, const AxisTap* X, const AxisTap* Y, vec3* out) {
    const TileR11<TW, TH> t{tileData, tileX0, tileY0};
#pragma unroll 1
    for (int j = 0; j < 3; j++)
#pragma unroll 1
        for (int i = 0; i < 3; i++) out[i * 3 + j] = tapR11Tile(t, img, X[i], Y[j]);
}
template <int TW, int TH>
__device__ __forceinline__ bool window3x3Applies(const AxisTap (&X)[3], const AxisTap (&Y)[3]) {
    return X[0].l1 == X[0].l0 + 1u && X[1].l0 == X[0].l0 + 1u && X[1].l1 == X[0].l0 + 2u && X[2].l0 == X[0].l0 + 2u && X[2].l1 == X[0].l0 + 3u &&
           Y[0].l1 == Y[0].l0 + 1u && Y[1].l0 == Y[0].l0 + 1u && Y[1].l1 == Y[0].l0 + 2u && Y[2].l0 == Y[0].l0 + 2u && Y[2].l1 == Y[0].l0 + 3u &&
           X[0].l0 < (unsigned)(TW - 3) && Y[0].l0 < (unsigned)(TH - 3);
}
// out[i][j] = tap (i, j); caller checked window3x3Applies
template <int TW, int TH, typename F>
__device__ __forceinline__ void window3x3(const TileR11<TW, TH>& t, const AxisTap (&X)[3], const AxisTap (&Y)[3], F&& emit) {
    const float4* row = t.s + Y[0].l0 * TW + X[0].l0;
    float4 r0[4], r1[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { r0[c] = row[c]; r1[c] = row[TW + c]; }
#pragma unroll
    for (int j = 0; j < 3; j++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float w00 = X[i].b * Y[j].b, w10 = X[i].a * Y[j].b, w01 = X[i].b * Y[j].a, w11 = X[i].a * Y[j].a;
            const vec3 t00 = v3(r0[i].x, r0[i].y, r0[i].z), t10 = v3(r0[i + 1].x, r0[i + 1].y, r0[i + 1].z);
            const vec3 t01 = v3(r1[i].x, r1[i].y, r1[i].z), t11 = v3(r1[i + 1].x, r1[i + 1].y, r1[i + 1].z);
            emit(i, j, vfma(t11, w11, vfma(t01, w01, vfma(t10, w10, t00 * w00))));
        }
        if (j < 2) {
#pragma unroll
            for (int c = 0; c < 4; c++) { r0[c] = r1[c]; r1[c] = row[(j + 2) * TW + c]; }
        }
    }
}

}  // namespace pb
