// passes_post.cu - TAA resolve, bloom chain, tonemapping.
//   temporalFilter.comp:84-179 + temporalReprojection.inc:8-87 + bicubicSampling.inc:4-181;
//   bloomDownsample.comp:12-50, bloomUpsample.comp:19-58, applyBloom.comp:16-31; tonemapping.comp:17-27
#include <cmath>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "shader_inc.cuh"
#include "tile.cuh"

namespace pb {

// ---------------- tonemapping.comp ----------------
// four pixels per thread: one 128-bit load of packed R11G11B10, one 128-bit store of B8G8R8A8
__device__ __forceinline__ uint32_t tonemapTexel(uint32_t packed, int x, int y, float g_time) {
    const vec3 linearColor = unpackR11G11B10(packed);
    const vec3 tonemapped = ACESFitted(linearColor);
    vec3 sRGB = linearTosRGB(tonemapped);
    sRGB = ditherRGB8(sRGB, x, y, g_time);
    return floatToUnorm8(sRGB.z) | (floatToUnorm8(sRGB.y) << 8) | (floatToUnorm8(sRGB.x) << 16) | (255u << 24);  // B8G8R8A8, alpha 1
}
__global__ void __launch_bounds__(256) tonemappingKernel(ImgView imageOut, ImgView imageIn, const plain_global_shader_info* __restrict__ g, int limitX, int limitY, int yBegin) {
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y = yBegin + blockIdx.y * 8 + (threadIdx.x >> 5);
    const int w = imin(imin(imageOut.w, imageIn.w), limitX), h = imin(imin(imageOut.h, imageIn.h), limitY);
    if (y >= h || x0 >= w) return;
    const float g_time = g->time;
    const uint32_t* in = (const uint32_t*)imageIn.ptr + (size_t)y * imageIn.w + x0;
    uint32_t* out = (uint32_t*)imageOut.ptr + (size_t)y * imageOut.w + x0;
    if (x0 + 3 < w && ((imageIn.w | imageOut.w) & 3) == 0) {
        const uint4 t = __ldg((const uint4*)in);
        uint4 o;
        o.x = tonemapTexel(t.x, x0, y, g_time); o.y = tonemapTexel(t.y, x0 + 1, y, g_time);
        o.z = tonemapTexel(t.z, x0 + 2, y, g_time); o.w = tonemapTexel(t.w, x0 + 3, y, g_time);
        *(uint4*)out = o;
    } else {
        for (int i = 0; i < 4 && x0 + i < w; i++) out[i] = tonemapTexel(__ldg(in + i), x0 + i, y, g_time);
    }
}
PLAIN_PASS(launch_tonemapping, "tonemapping.comp") {
    const ImgView out = c.storage(0, PLAIN_FORMAT_BGRA8_UNORM);
    const ImgView in = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    int y0, y1;
    c.window(std::min(out.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(std::min(out.w, limX), 128), ceilDiv((unsigned)(y1 - y0), 8));
    PLAIN_LAUNCH(c, tonemappingKernel, grid, 256, 0, out, in, c.g, limX, y1, y0);
}

// ---------------- bloom ----------------
// bloomDownsample.comp: 13 bilinear taps of the finer mip. The bounds test is '>' in the reference (:16): the extra
// row/column of invocations only produces stores outside the image, which are dropped.
// Block = 32x8 target texels; their 13 taps touch a (64 + 6) x (16 + 6) rectangle of the finer mip, staged once.
// The 13 taps use 5 distinct x and 5 distinct y coordinates (uv + texelSize * {0, +-0.5, +-1.5}): 10 axis set-ups instead of 26
// (tile.cuh). FAST (block-uniform): the staged rectangle lies inside the source - no clamps, no bounds tests, no sanitising
// (the coordinates are pixel centres plus fixed offsets); a pixel whose taps all lie inside the tile reads them with no test at all.
#define BLOOM_DOWN_TW 72
#define BLOOM_DOWN_TH 24
__device__ __noinline__ void bloomDownsampleGeneric(ImgView target, ImgView source, const float4* sSrc, int sx0, int sy0, int ix, int iy) {
    const TileR11<BLOOM_DOWN_TW, BLOOM_DOWN_TH> tile{sSrc, sx0, sy0};
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    const vec2 texelSize = 1.f / v2((float)source.w, (float)source.h);
    vec3 color = v3(0.f);
    auto T = [&](float ox, float oy) { return sampleR11LinearClampTile(tile, source, uv + texelSize * v2(ox, oy)); };
    color = color + sampleR11LinearClampTile(tile, source, uv) * 0.125f;
    color = color + T(0.5f, 0.5f) * 0.125f; color = color + T(0.5f, -0.5f) * 0.125f; color = color + T(-0.5f, 0.5f) * 0.125f; color = color + T(-0.5f, -0.5f) * 0.125f;
    color = color + T(1.5f, 0.f) * 0.0625f; color = color + T(-1.5f, 0.f) * 0.0625f; color = color + T(0.f, 1.5f) * 0.0625f; color = color + T(0.f, -1.5f) * 0.0625f;
    color = color + T(1.5f, 1.5f) * 0.03125f; color = color + T(1.5f, -1.5f) * 0.03125f; color = color + T(-1.5f, 1.5f) * 0.03125f; color = color + T(-1.5f, -1.5f) * 0.03125f;
    storeR11(target, ix, iy, color);
}
// one 32x8 block of target texels at (bx, by); sSrc: BLOOM_DOWN_TW x BLOOM_DOWN_TH float4. Called by the per-level kernel and by the
// persistent kernel of the small levels (bloomTailKernel): every thread of the block calls it, returning early is fine (no exit)
__device__ __forceinline__ void bloomDownsampleTile(const ImgView& target, const ImgView& source, int bx, int by, int yEnd, float4* sSrc) {
    constexpr int TW = BLOOM_DOWN_TW, TH = BLOOM_DOWN_TH;
    const int sx0 = (int)(((long long)bx * source.w) / target.w) - 3, sy0 = (int)(((long long)by * source.h) / target.h) - 3;
    const bool interior = tileIsInterior<TW, TH>(source, sx0, sy0);
    if (interior) tileLoadR11<TW, TH, true, true>(sSrc, source, sx0, sy0);  // swizzled columns: neighbouring lanes read columns two apart
    else tileLoadR11<TW, TH, false>(sSrc, source, sx0, sy0);
    __syncthreads();
    const int ix = bx + (threadIdx.x & 31), iy = by + (threadIdx.x >> 5);
    if (ix >= target.w || iy >= yEnd) return;
    if (!interior) { bloomDownsampleGeneric(target, source, sSrc, sx0, sy0, ix, iy); return; }
    const TileR11<TW, TH> tile{sSrc, sx0, sy0};
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    const vec2 texelSize = 1.f / v2((float)source.w, (float)source.h);
    // axis k: offset {0, +0.5, -0.5, +1.5, -1.5}[k]; uv + texelSize * 0 == uv (uv > 0), so axis 0 is the centre tap's own set-up
    AxisTap X[5], Y[5];
    const float off[5] = {0.f, 0.5f, -0.5f, 1.5f, -1.5f};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        X[k] = axisTapT<false, true>(k == 0 ? uv.x : uv.x + texelSize.x * off[k], source.w, sx0);
        Y[k] = axisTapT<false, true>(k == 0 ? uv.y : uv.y + texelSize.y * off[k], source.h, sy0);
    }
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 5; k++) inside = inside && X[k].l0 < (unsigned)(TW - 1) && Y[k].l0 < (unsigned)(TH - 1);
    if (!inside) {  // cannot happen for 2:1 mip chains; the swizzled tile is not what the generic taps expect: they read the image itself
        bloomDownsampleGeneric(target, source, sSrc, source.w + TW, source.h + TH, ix, iy);  // a tile origin no texel can lie in
        return;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) { X[k].l0 = tileSwizzle<TW>(X[k].l0); X[k].l1 = tileSwizzle<TW>(X[k].l1); }
    vec3 color = v3(0.f);
    auto T = [&](int kx, int ky) { return tapR11TileInside(tile, X[kx], Y[ky]); };
    color = color + T(0, 0) * 0.125f;
    color = color + T(1, 1) * 0.125f; color = color + T(1, 2) * 0.125f; color = color + T(2, 1) * 0.125f; color = color + T(2, 2) * 0.125f;
    color = color + T(3, 0) * 0.0625f; color = color + T(4, 0) * 0.0625f; color = color + T(0, 3) * 0.0625f; color = color + T(0, 4) * 0.0625f;
    color = color + T(3, 3) * 0.03125f; color = color + T(3, 4) * 0.03125f; color = color + T(4, 3) * 0.03125f; color = color + T(4, 4) * 0.03125f;
    storeR11(target, ix, iy, color);
}
__global__ void __launch_bounds__(256) bloomDownsampleKernel(ImgView target, ImgView source, int yBegin, int yEnd) {
    __shared__ float4 sSrc[BLOOM_DOWN_TW * BLOOM_DOWN_TH];
    bloomDownsampleTile(target, source, blockIdx.x * 32, yBegin + blockIdx.y * 8, yEnd, sSrc);
}
// ---- the same pass with the source tile staged by TMA (the default when the source rows are 16-byte multiples; PLAIN_BLOOM_TMA=0
//      selects the plain loader; A / B in profiles/r2_tma_ab.md) ----
// One elected thread issues cp.async.bulk.tensor.2d for the RAW packed tile (72 x 24 texels x 4 bytes; texels outside the image are
// zero-filled by the descriptor, which decode to (0, 0, 0) like the loader's) and the block waits on an mbarrier; a second pass decodes
// shared -> shared into the float4 tile the taps read. Against the plain loader it trades the per-texel address arithmetic, bounds
// tests and LDG for an LDS + one more block-wide hand-over.
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define BLOOM_DOWN_TMA_TW 76  // the TMA box starts at a multiple of 4 texels (16-byte aligned innermost coordinate): up to 3 more columns on the left
__global__ void __launch_bounds__(256) bloomDownsampleTmaKernel(const __grid_constant__ CUtensorMap sourceMap, ImgView target, ImgView source, int yBegin, int yEnd) {
    constexpr int TW = BLOOM_DOWN_TMA_TW, TH = BLOOM_DOWN_TH;
    __shared__ __align__(128) uint32_t sRaw[TW * TH];
    __shared__ float4 sSrc[TW * TH];
    __shared__ __align__(8) unsigned long long sBar;
    const int bx = blockIdx.x * 32, by = yBegin + blockIdx.y * 8;
    // the innermost box coordinate must be a multiple of 16 bytes (a box at x = 125 or x = -3 faults with "illegal instruction",
    // tools/microbench/tma_probe.cu); negative and overflowing coordinates are fine: those texels arrive as zeros
    const int sx0 = ((int)(((long long)bx * source.w) / target.w) - 3) & ~3, sy0 = (int)(((long long)by * source.h) / target.h) - 3;
    const bool interior = tileIsInterior<TW, TH>(source, sx0, sy0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(&sBar)), "r"((uint32_t)(TW * TH * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smemAddr(sRaw)), "l"(&sourceMap), "r"(sx0), "r"(sy0), "r"(smemAddr(&sBar)) : "memory");
    }
    {   // every thread waits for the bytes to land (phase 0 of the barrier)
        uint32_t done = 0, spins = 0;
        while (!done) {
            if (++spins > (1u << 22)) __trap();  // a descriptor fault must end the kernel, not hang the GPU
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smemAddr(&sBar)) : "memory");
        }
    }
    // decode shared -> shared (swizzled columns in interior blocks, like the plain loader)
    for (int i = threadIdx.x; i < TW * TH; i += 256) {
        const int tx = i % TW, ty = i / TW;
        sSrc[ty * TW + (interior ? (int)tileSwizzle<TW>((unsigned)tx) : tx)] = decodeR11Texel(sRaw[i]);
    }
    __syncthreads();
    const int ix = bx + (threadIdx.x & 31), iy = by + (threadIdx.x >> 5);
    if (ix >= target.w || iy >= yEnd) return;
    const TileR11<TW, TH> tile{sSrc, sx0, sy0};
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    const vec2 texelSize = 1.f / v2((float)source.w, (float)source.h);
    if (!interior) {  // the spelled-out taps on the (unswizzled) tile, global loads for anything outside it
        vec3 color = v3(0.f);
        auto T = [&](float ox, float oy) { return sampleR11LinearClampTile(tile, source, uv + texelSize * v2(ox, oy)); };
        color = color + sampleR11LinearClampTile(tile, source, uv) * 0.125f;
        color = color + T(0.5f, 0.5f) * 0.125f; color = color + T(0.5f, -0.5f) * 0.125f; color = color + T(-0.5f, 0.5f) * 0.125f; color = color + T(-0.5f, -0.5f) * 0.125f;
        color = color + T(1.5f, 0.f) * 0.0625f; color = color + T(-1.5f, 0.f) * 0.0625f; color = color + T(0.f, 1.5f) * 0.0625f; color = color + T(0.f, -1.5f) * 0.0625f;
        color = color + T(1.5f, 1.5f) * 0.03125f; color = color + T(1.5f, -1.5f) * 0.03125f; color = color + T(-1.5f, 1.5f) * 0.03125f; color = color + T(-1.5f, -1.5f) * 0.03125f;
        storeR11(target, ix, iy, color);
        return;
    }
    AxisTap X[5], Y[5];
    const float off[5] = {0.f, 0.5f, -0.5f, 1.5f, -1.5f};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        X[k] = axisTapT<false, true>(k == 0 ? uv.x : uv.x + texelSize.x * off[k], source.w, sx0);
        Y[k] = axisTapT<false, true>(k == 0 ? uv.y : uv.y + texelSize.y * off[k], source.h, sy0);
    }
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 5; k++) inside = inside && X[k].l0 < (unsigned)(TW - 1) && Y[k].l0 < (unsigned)(TH - 1);
    if (!inside) { bloomDownsampleGeneric(target, source, nullptr, source.w + TW, source.h + TH, ix, iy); return; }  // taps from the image itself
#pragma unroll
    for (int k = 0; k < 5; k++) { X[k].l0 = tileSwizzle<TW>(X[k].l0); X[k].l1 = tileSwizzle<TW>(X[k].l1); }
    vec3 color = v3(0.f);
    auto T = [&](int kx, int ky) { return tapR11TileInside(tile, X[kx], Y[ky]); };
    color = color + T(0, 0) * 0.125f;
    color = color + T(1, 1) * 0.125f; color = color + T(1, 2) * 0.125f; color = color + T(2, 1) * 0.125f; color = color + T(2, 2) * 0.125f;
    color = color + T(3, 0) * 0.0625f; color = color + T(4, 0) * 0.0625f; color = color + T(0, 3) * 0.0625f; color = color + T(0, 4) * 0.0625f;
    color = color + T(3, 3) * 0.03125f; color = color + T(3, 4) * 0.03125f; color = color + T(4, 3) * 0.03125f; color = color + T(4, 4) * 0.03125f;
    storeR11(target, ix, iy, color);
}
// 2-D tensor map of one R11G11B10 mip level (32-bit texels, tightly packed rows), box = the staged tile, out-of-image texels zero
static bool makeTileTensorMap(CUtensorMap* map, const ImgView& img, int boxW, int boxH) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return false;
        encode = (PFN_cuTensorMapEncodeTiled)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)img.w, (cuuint64_t)img.h};
    const cuuint64_t strides[1] = {(cuuint64_t)img.w * 4};  // bytes, dimension 1 (must be a multiple of 16)
    const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH}, elemStrides[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, img.ptr, dims, strides, box, elemStrides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool launchBloomTail(LaunchCtx& c);  // after bloomTailKernel, below
PLAIN_PASS(launch_bloomDownsample, "bloomDownsample.comp") {
    if (!c.exec->fusedRun.empty() && launchBloomTail(c)) return;
    const ImgView target = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT), source = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < target.w || (int)c.exec->dispatch[1] * 8 < target.h) { c.fail("bloomDownsample.comp: dispatch does not cover the target"); return; }
    int y0, y1;
    c.window(target.h, y0, y1);
    if (y1 <= y0) return;
    const dim3 grid(ceilDiv(target.w, 32), ceilDiv((unsigned)(y1 - y0), 8));
    static const bool useTma = !getenv("PLAIN_BLOOM_TMA") || atoi(getenv("PLAIN_BLOOM_TMA")) != 0;  // on by default: 3.7 % faster at 4K (profiles/r2_tma_ab.md)
    if (useTma && (source.w % 4) == 0 && ((uintptr_t)source.ptr % 16) == 0 && source.w >= BLOOM_DOWN_TW / 4) {
        CUtensorMap map;
        if (makeTileTensorMap(&map, source, BLOOM_DOWN_TMA_TW, BLOOM_DOWN_TH)) {
            PLAIN_LAUNCH(c, bloomDownsampleTmaKernel, grid, 256, 0, map, target, source, y0, y1);
            return;
        }
    }
    PLAIN_LAUNCH(c, bloomDownsampleKernel, grid, 256, 0, target, source, y0, y1);
}

// bloomUpsample.comp: 9-tap tent of the coarser downsample mip (+ 4-tap box of the coarser upsample mip)
// Block = 32x8 target texels; the tent and box taps touch about (16 + 8) x (4 + 8) texels of the two coarser mips.
// Tent: 3 x + 3 y axis set-ups (uv + sampleStepSize * {0, +1, -1}); box: 2 + 2 (uv + texelSize * {+0.5, -0.5}).
#define BLOOM_UP_TW 24
#define BLOOM_UP_TH 12
__device__ __noinline__ void bloomUpsampleGeneric(ImgView target, ImgView targetPreviousMip, ImgView source, const float4* sSrc, const float4* sPrev, int sx0, int sy0, int isLowestMip, float blurRadius, int ix, int iy) {
    const TileR11<BLOOM_UP_TW, BLOOM_UP_TH> srcTile{sSrc, sx0, sy0}, prevTile{sPrev, sx0, sy0};
    const vec2 texelSize = 1.f / v2((float)source.w, (float)source.h);
    const vec2 sampleStepSize = blurRadius * texelSize;
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    vec3 color = v3(0.f);
    auto S = [&](float ox, float oy) { return sampleR11LinearClampTile(srcTile, source, uv + sampleStepSize * v2(ox, oy)); };
    color = color + sampleR11LinearClampTile(srcTile, source, uv) * 0.25f;
    color = color + S(1.f, 0.f) * 0.125f; color = color + S(-1.f, 0.f) * 0.125f; color = color + S(0.f, 1.f) * 0.125f; color = color + S(0.f, -1.f) * 0.125f;
    color = color + S(1.f, 1.f) * 0.0625f; color = color + S(1.f, -1.f) * 0.0625f; color = color + S(-1.f, 1.f) * 0.0625f; color = color + S(-1.f, -1.f) * 0.0625f;
    if (!isLowestMip) {
        auto P = [&](float ox, float oy) { return sampleR11LinearClampTile(prevTile, targetPreviousMip, uv + texelSize * v2(ox, oy)); };
        color = color + P(0.5f, 0.5f) * 0.25f; color = color + P(0.5f, -0.5f) * 0.25f; color = color + P(-0.5f, 0.5f) * 0.25f; color = color + P(-0.5f, -0.5f) * 0.25f;
    }
    storeR11(target, ix, iy, color);
}
__device__ __forceinline__ void bloomUpsampleTile(const ImgView& target, const ImgView& targetPreviousMip, const ImgView& source, int isLowestMip, float blurRadius, int fastAllowed,
                                                  int bx, int by, int yEnd, float4* sSrc, float4* sPrev) {
    constexpr int TW = BLOOM_UP_TW, TH = BLOOM_UP_TH;
    const int sx0 = (int)(((long long)bx * source.w) / target.w) - 3, sy0 = (int)(((long long)by * source.h) / target.h) - 3;
    // fastAllowed (launcher): the previous mip has the source's extent and blurRadius is a small finite number (bounded coordinates)
    const bool interior = fastAllowed && tileIsInterior<TW, TH>(source, sx0, sy0);
    if (interior) {
        tileLoadR11<TW, TH, true>(sSrc, source, sx0, sy0);
        if (!isLowestMip) tileLoadR11<TW, TH, true>(sPrev, targetPreviousMip, sx0, sy0);
    } else {
        tileLoadR11<TW, TH, false>(sSrc, source, sx0, sy0);
        if (!isLowestMip) tileLoadR11<TW, TH, false>(sPrev, targetPreviousMip, sx0, sy0);
    }
    __syncthreads();
    const int ix = bx + (threadIdx.x & 31), iy = by + (threadIdx.x >> 5);
    if (ix >= target.w || iy >= yEnd) return;
    if (!interior) { bloomUpsampleGeneric(target, targetPreviousMip, source, sSrc, sPrev, sx0, sy0, isLowestMip, blurRadius, ix, iy); return; }
    const TileR11<TW, TH> srcTile{sSrc, sx0, sy0}, prevTile{sPrev, sx0, sy0};
    const vec2 texelSize = 1.f / v2((float)source.w, (float)source.h);
    const vec2 sampleStepSize = blurRadius * texelSize;
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    // tent axes k: offset {0, +1, -1}[k]; uv + sampleStepSize * 0 == uv (a finite step times 0 is +-0, uv > 0)
    AxisTap X[3], Y[3], BX[2], BY[2];
    const float off[3] = {0.f, 1.f, -1.f};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        X[k] = axisTapT<false, true>(k == 0 ? uv.x : uv.x + sampleStepSize.x * off[k], source.w, sx0);
        Y[k] = axisTapT<false, true>(k == 0 ? uv.y : uv.y + sampleStepSize.y * off[k], source.h, sy0);
    }
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; k++) inside = inside && X[k].l0 < (unsigned)(TW - 1) && Y[k].l0 < (unsigned)(TH - 1);
    if (!isLowestMip) {
#pragma unroll
        for (int k = 0; k < 2; k++) {  // box axes: offset {+0.5, -0.5}[k]
            BX[k] = axisTapT<false, true>(uv.x + texelSize.x * (k == 0 ? 0.5f : -0.5f), targetPreviousMip.w, sx0);
            BY[k] = axisTapT<false, true>(uv.y + texelSize.y * (k == 0 ? 0.5f : -0.5f), targetPreviousMip.h, sy0);
            inside = inside && BX[k].l0 < (unsigned)(TW - 1) && BY[k].l0 < (unsigned)(TH - 1);
        }
    }
    if (!inside) { bloomUpsampleGeneric(target, targetPreviousMip, source, sSrc, sPrev, sx0, sy0, isLowestMip, blurRadius, ix, iy); return; }
    vec3 color = v3(0.f);
    auto S = [&](int kx, int ky) { return tapR11TileInside(srcTile, X[kx], Y[ky]); };
    color = color + S(0, 0) * 0.25f;
    color = color + S(1, 0) * 0.125f; color = color + S(2, 0) * 0.125f; color = color + S(0, 1) * 0.125f; color = color + S(0, 2) * 0.125f;
    color = color + S(1, 1) * 0.0625f; color = color + S(1, 2) * 0.0625f; color = color + S(2, 1) * 0.0625f; color = color + S(2, 2) * 0.0625f;
    if (!isLowestMip) {
        auto P = [&](int kx, int ky) { return tapR11TileInside(prevTile, BX[kx], BY[ky]); };
        color = color + P(0, 0) * 0.25f; color = color + P(0, 1) * 0.25f; color = color + P(1, 0) * 0.25f; color = color + P(1, 1) * 0.25f;
    }
    storeR11(target, ix, iy, color);
}
__global__ void __launch_bounds__(256, 4) bloomUpsampleKernel(ImgView target, ImgView targetPreviousMip, ImgView source, int isLowestMip, float blurRadius, int fastAllowed, int yBegin, int yEnd) {
    __shared__ float4 sSrc[BLOOM_UP_TW * BLOOM_UP_TH], sPrev[BLOOM_UP_TW * BLOOM_UP_TH];
    bloomUpsampleTile(target, targetPreviousMip, source, isLowestMip, blurRadius, fastAllowed, blockIdx.x * 32, yBegin + blockIdx.y * 8, yEnd, sSrc, sPrev);
}

// ---- the small levels of the chain in ONE launch (Bloom.cpp:65-122: downsample mips >= 2, then upsample back to mip 2) ----
// At 3840x2160 those are seven dependent launches of 36 .. 2025 blocks, 8 - 29 us each and mostly launch latency and tail. The backend
// (planFusions) hands the whole run to a persistent kernel: two blocks per SM walk the 32x8 tiles of a level, a grid-wide barrier (one
// counter per level in global memory, zeroed by the launcher) separates the levels. Same tile functions, same bits. The barrier gives up
// after ~50 ms into an error word instead of hanging the device if the grid is ever not co-resident.
#define BLOOM_TAIL_MAX_LEVELS 10
struct BloomTailLevel { ImgView target, source, prev; int isUpsample, isLowestMip, fastAllowed; float blurRadius; };
struct BloomTailParams { BloomTailLevel lv[BLOOM_TAIL_MAX_LEVELS]; int nLevels; unsigned int* counters; };  // counters[0 .. nLevels): barrier arrivals; counters[63]: sticky error word (wait_for_gpu_idle reports it)
__global__ void __launch_bounds__(256, 4) bloomTailKernel(const __grid_constant__ BloomTailParams p) {
    __shared__ float4 sTile[BLOOM_DOWN_TW * BLOOM_DOWN_TH];  // the upsample's two 24x12 tiles fit in it
    static_assert(2 * BLOOM_UP_TW * BLOOM_UP_TH <= BLOOM_DOWN_TW * BLOOM_DOWN_TH, "tile sizes");
    for (int l = 0; l < p.nLevels; l++) {
        const BloomTailLevel& L = p.lv[l];
        const int tilesX = (L.target.w + 31) / 32, tilesY = (L.target.h + 7) / 8;
        for (int t = blockIdx.x; t < tilesX * tilesY; t += gridDim.x) {
            const int bx = (t % tilesX) * 32, by = (t / tilesX) * 8;
            if (L.isUpsample) bloomUpsampleTile(L.target, L.prev, L.source, L.isLowestMip, L.blurRadius, L.fastAllowed, bx, by, L.target.h, sTile, sTile + BLOOM_UP_TW * BLOOM_UP_TH);
            else bloomDownsampleTile(L.target, L.source, bx, by, L.target.h, sTile);
            __syncthreads();  // the tile is reused
        }
        if (l + 1 < p.nLevels) {  // grid barrier: this level's stores are visible to every block before the next level reads them
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(p.counters + l, 1u);
                const long long start = clock64();
                while (*(volatile unsigned int*)(p.counters + l) < gridDim.x) {
                    if (clock64() - start > 100000000ll) { atomicExch(p.counters + 63, 1u); break; }
                    __nanosleep(32);
                }
                __threadfence();
            }
            __syncthreads();
        }
    }
}
// a run of small levels handed over by the backend (ExecRecord::fusedRun): one persistent launch, two blocks per SM
static bool launchBloomTail(LaunchCtx& c) {
    BloomTailParams p;
    p.nLevels = 0;
    for (int index : c.exec->fusedRun) {
        LaunchCtx m = c;
        m.exec = c.be_exec(index);
        m.pass = c.be_pass(m.exec->pass);
        BloomTailLevel& L = p.lv[p.nLevels++];
        L.isUpsample = m.pass->shader == "bloomUpsample.comp" ? 1 : 0;
        L.target = m.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
        if (L.isUpsample) {
            L.prev = m.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
            L.source = m.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
            L.isLowestMip = m.specBool(0, false) ? 1 : 0;
            L.blurRadius = m.push<float>(0);
            L.fastAllowed = (L.isLowestMip || (L.prev.w == L.source.w && L.prev.h == L.source.h)) && L.blurRadius == L.blurRadius && std::fabs(L.blurRadius) <= 64.f;
        } else {
            L.source = m.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
            L.prev = L.source;
            L.isLowestMip = 0; L.fastAllowed = 0; L.blurRadius = 0.f;
        }
        if (m.failed) { c.fail(m.error); return true; }
        if ((int)m.exec->dispatch[0] * 8 < L.target.w || (int)m.exec->dispatch[1] * 8 < L.target.h) { c.fail("bloom chain: dispatch does not cover the target"); return true; }
    }
    p.counters = c.be_fusionCounters();
    if (!p.counters) return false;
    if (cudaMemsetAsync(p.counters, 0, (size_t)p.nLevels * sizeof(unsigned int), c.stream) != cudaSuccess) return false;
    PLAIN_LAUNCH(c, bloomTailKernel, dim3((unsigned)(4 * c.smCount)), 256, 0, p);
    return true;
}
PLAIN_PASS(launch_bloomUpsample, "bloomUpsample.comp") {
    const ImgView target = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const ImgView prev = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT), source = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < target.w || (int)c.exec->dispatch[1] * 8 < target.h) { c.fail("bloomUpsample.comp: dispatch does not cover the target"); return; }
    int y0, y1;
    c.window(target.h, y0, y1);
    if (y1 <= y0) return;
    const int isLowestMip = c.specBool(0, false) ? 1 : 0;
    const float blurRadius = c.push<float>(0);
    const int fastAllowed = (isLowestMip || (prev.w == source.w && prev.h == source.h)) && blurRadius == blurRadius && std::fabs(blurRadius) <= 64.f;
    PLAIN_LAUNCH(c, bloomUpsampleKernel, dim3(ceilDiv(target.w, 32), ceilDiv((unsigned)(y1 - y0), 8)), 256, 0, target, prev, source, isLowestMip, blurRadius, fastAllowed, y0, y1);
}

// applyBloom.comp: mix(scene, bloom, strength) in place
__global__ void __launch_bounds__(256) applyBloomKernel(ImgView target, ImgView bloomTexture, float bloomStrength, int yBegin, int yEnd) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = yBegin + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= target.w || iy >= yEnd) return;
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)target.w, (float)target.h);
    const vec3 bloom = sampleR11LinearClamp(bloomTexture, uv);
    const vec3 scene = loadR11(target, ix, iy);
    storeR11(target, ix, iy, vmix(scene, bloom, bloomStrength));
}
PLAIN_PASS(launch_applyBloom, "applyBloom.comp") {
    const ImgView target = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT), bloom = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < target.w || (int)c.exec->dispatch[1] * 8 < target.h) { c.fail("applyBloom.comp: dispatch does not cover the target"); return; }
    int y0, y1;
    c.window(target.h, y0, y1);
    if (y1 <= y0) return;
    PLAIN_LAUNCH(c, applyBloomKernel, dim3(ceilDiv(target.w, 32), ceilDiv((unsigned)(y1 - y0), 8)), 256, 0, target, bloom, c.push<float>(0), y0, y1);
}

// ---------------- temporalFilter.comp ----------------
struct TaaParams {
    ImgView currentFrame, historySrc, motionBuffer, depthBuffer, outputImage, historyDst;
    const float* resolveWeights;  // 9 floats, uniform buffer binding 6 (TAA.cpp:181-202)
    const plain_global_shader_info* g;
    int useClipping, useMotionVectorDilation, historySampleTech, useTonemap;
    int y0, y1;  // rows to produce (row sharding)
    int sameExtents;  // every bound image has the target's extent (set by the launcher)
};
__device__ __forceinline__ vec3 taaTonemap(vec3 color) { return color / (1.f + computeLuminance(color)); }         // temporalReprojection.inc:34-36
__device__ __forceinline__ vec3 taaTonemapReverse(vec3 color) { return color / (1.f - computeLuminance(color)); }  // :38-40
// taaTonemap of a FINITE colour without negative components (a texel of an unsigned format or a bilinear blend of such texels):
// 1 + luminance lies in [1, 2^17), so the reciprocal needs no range test (pvec.h rcpf_normal: the same correctly rounded value)
__device__ __forceinline__ vec3 taaTonemapFiniteNonNegative(vec3 color) { return color * rcpf_normal(1.f + computeLuminance(color)); }
struct Nb { vec3 v[3][3]; };  // v[x+1][y+1]
// FAST (block-uniform, decided after staging): the tile lies inside the image and holds no inf / NaN texel.
template <bool TONEMAP, bool FAST, int TW, int TH, typename Store>
__device__ __forceinline__ void sampleNeighbourhoodTo(const TileR11<TW, TH>& tile, const ImgView& tex, vec2 uv, vec2 texelSize, Store store) {  // :42-52; store(x, y, tap)
    AxisTap X[3], Y[3];  // separable set-up of the nine taps (tile.cuh)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // FAST: the coordinates are finite and bounded (pixel centre + SNORM16 motion + one texel): no sanitising, no clamps
        X[i] = FAST ? axisTapT<false, true>(uv.x + texelSize.x * (float)(i - 1), tex.w, tile.x0) : axisTap(uv.x + texelSize.x * (float)(i - 1), tex.w, tile.x0);
        Y[i] = FAST ? axisTapT<false, true>(uv.y + texelSize.y * (float)(i - 1), tex.h, tile.y0) : axisTap(uv.y + texelSize.y * (float)(i - 1), tex.h, tile.y0);
    }
    if (FAST) {
        if (window3x3Applies<TW, TH>(X, Y)) {
            window3x3(tile, X, Y, [&](int x, int y, vec3 color) { store(x, y, TONEMAP ? taaTonemapFiniteNonNegative(color) : color); });
        } else {  // a tap leaves the tile (fast motion): every tap on its own, out of line
            AxisTap Xc[3] = {X[0], X[1], X[2]}, Yc[3] = {Y[0], Y[1], Y[2]};
            vec3 taps[9];
            taps3x3Generic<TW, TH>(tile.s, tile.x0, tile.y0, tex, Xc, Yc, taps);
#pragma unroll
            for (int x = 0; x < 3; x++)
#pragma unroll
                for (int y = 0; y < 3; y++) store(x, y, TONEMAP ? taaTonemapFiniteNonNegative(taps[x * 3 + y]) : taps[x * 3 + y]);
        }
    } else {
#pragma unroll
        for (int x = 0; x < 3; x++)
#pragma unroll
            for (int y = 0; y < 3; y++) {
                const vec3 color = tapR11Tile(tile, tex, X[x], Y[y]);
                store(x, y, TONEMAP ? taaTonemap(color) : color);
            }
    }
}
template <bool TONEMAP, bool FAST, int TW, int TH>
__device__ __forceinline__ void sampleNeighbourhood(const TileR11<TW, TH>& tile, const ImgView& tex, vec2 uv, vec2 texelSize, Nb& n) {
    sampleNeighbourhoodTo<TONEMAP, FAST>(tile, tex, uv, texelSize, [&](int x, int y, vec3 tap) { n.v[x][y] = tap; });
}
// the history neighbourhood only feeds its luminance contrast (temporalFilter.comp:150, the reference's own TODO): nine luminances are kept
// instead of nine colours - computeLuminance of the same tap, summed in neighbourhoodContrast's order
struct NbLum { float l[3][3]; };
template <bool TONEMAP, bool FAST, int TW, int TH>
__device__ __forceinline__ float neighbourhoodContrastAt(const TileR11<TW, TH>& tile, const ImgView& tex, vec2 uv, vec2 texelSize) {
    NbLum n;
    sampleNeighbourhoodTo<TONEMAP, FAST>(tile, tex, uv, texelSize, [&](int x, int y, vec3 tap) { n.l[x][y] = computeLuminance(tap); });
    const float c11 = n.l[1][1];
    return absf(n.l[0][0] - c11) + absf(n.l[1][0] - c11) + absf(n.l[2][0] - c11) + absf(n.l[0][2] - c11) + absf(n.l[1][2] - c11) + absf(n.l[2][2] - c11) +
           absf(n.l[0][1] - c11) + absf(n.l[2][1] - c11);
}
template <bool FAST>
__device__ __forceinline__ vec3 clipAABB(vec3 target, vec3 bbMin, vec3 bbMax) {  // :8-30
    const vec3 epsilon = v3(0.0001f);
    const vec3 center = 0.5f * (bbMax + bbMin);
    const vec3 extend = 0.5f * (bbMax - bbMin) + epsilon;
    const vec3 toTarget = target - center;
    // FAST: bbMin <= bbMax are finite, so extend lies in [0.0001, 2^17): the reciprocals of toTarget / extend need no range test
    const vec3 a = FAST ? vabs(v3(toTarget.x * rcpf_normal(extend.x), toTarget.y * rcpf_normal(extend.y), toTarget.z * rcpf_normal(extend.z))) : vabs(toTarget / extend);
    const float maxComponent = fmaxp(a.x, fmaxp(a.y, a.z));
    if (maxComponent < 1.f) return target;
    return center + toTarget / maxComponent;
}
__device__ __forceinline__ float catmullRomWeight1D(float d) {  // bicubicSampling.inc:4-17
    const float d1 = absf(d), d2 = d1 * d1, d3 = d2 * d1;
    if (d1 <= 1.f) return (1.f / 6.f) * (9.f * d3 - 15.f * d2 + 6.f);
    else if (d1 <= 2.f) return (1.f / 6.f) * (-3.f * d3 + 15.f * d2 - 24.f * d + 12.f);
    return 0.f;
}
__device__ __forceinline__ float neighbourhoodContrast(const Nb& n) {  // temporalFilter.comp:59-69
    const float c11 = computeLuminance(n.v[1][1]);
    return absf(computeLuminance(n.v[0][0]) - c11) + absf(computeLuminance(n.v[1][0]) - c11) + absf(computeLuminance(n.v[2][0]) - c11) +
           absf(computeLuminance(n.v[0][2]) - c11) + absf(computeLuminance(n.v[1][2]) - c11) + absf(computeLuminance(n.v[2][2]) - c11) +
           absf(computeLuminance(n.v[0][1]) - c11) + absf(computeLuminance(n.v[2][1]) - c11);
}
struct BicubicW { vec2 w0, w1, w2, w3, wB, t, uvTrunc; };
__device__ __forceinline__ vec2 vfloor(vec2 a) { return v2(floorf_(a.x), floorf_(a.y)); }
__device__ __forceinline__ BicubicW bicubicWeights(vec2 iUV) {  // bicubicSampling.inc:74-85
    BicubicW b;
    b.uvTrunc = vfloor(iUV - 0.5f) + 0.5f;
    const vec2 f = iUV - b.uvTrunc, f2 = f * f, f3 = f2 * f;
    b.w0 = -0.5f * f3 + f2 - 0.5f * f;
    b.w1 = 1.5f * f3 - 2.5f * f2 + 1.f;
    b.w2 = -1.5f * f3 + 2.f * f2 + 0.5f * f;
    b.w3 = 0.5f * f3 - 0.5f * f2;
    b.wB = b.w1 + b.w2;
    b.t = b.w2 / b.wB;
    return b;
}

// Block = 32x8 pixels. The current frame is staged with a 2-texel halo (3x3 bilinear taps at texel centres touch
// [-2, +2]); the history with a 6-texel halo, which covers the reprojected 3x3 neighbourhood and the bicubic tap for
// motion up to ~4 pixels - larger motion falls back to global loads per tap corner (tile.cuh).
// Two instantiations of one body, chosen per block after staging:
//   FAST    - the history tile (and with it the current tile, the 3x3 depth neighbourhood and the block's own pixels) lies inside
//             the image, all six images have the target's extent and no staged texel is inf / NaN: clamps, bounds tests, coordinate
//             sanitising and reciprocal range tests are the identity and are left out; 3x3 neighbourhoods read a 4x4 window once
//   !FAST   - image borders, odd configurations, inf / NaN texels: every rule spelled out (the round-1 kernel)
#define TAA_CUR_HALO 2
#define TAA_HIS_HALO 6
#define TAA_CUR_W (32 + 2 * TAA_CUR_HALO)
#define TAA_CUR_H (8 + 2 * TAA_CUR_HALO)
#define TAA_HIS_W (32 + 2 * TAA_HIS_HALO)
#define TAA_HIS_H (8 + 2 * TAA_HIS_HALO)
template <bool TONEMAP, int TECH, bool FAST>
__device__ __forceinline__ void temporalFilterPixel(const TaaParams& p, const float4* sCur, const float4* sHis, const float* sRw) {
    const int bx = blockIdx.x * 32, by = p.y0 + blockIdx.y * 8;
    const TileR11<TAA_CUR_W, TAA_CUR_H> curTile{sCur, bx - TAA_CUR_HALO, by - TAA_CUR_HALO};
    const TileR11<TAA_HIS_W, TAA_HIS_H> hisTile{sHis, bx - TAA_HIS_HALO, by - TAA_HIS_HALO};
    const int ix = bx + (threadIdx.x & 31), iy = by + (threadIdx.x >> 5);
    if (ix >= p.outputImage.w || iy >= p.y1) return;
    const vec2 screenRes = v2((float)p.g->screenResolution[0], (float)p.g->screenResolution[1]);
    const vec2 texelSize = 1.f / v2((float)p.outputImage.w, (float)p.outputImage.h);
    const vec2 iUVf = v2((float)ix, (float)iy);
    const vec2 uv = (iUVf + 0.5f) * texelSize;
    Nb nb;
    sampleNeighbourhood<TONEMAP, FAST>(curTile, p.currentFrame, uv, texelSize, nb);
    // minMaxFromNeighbourhood :54-65. FAST: the neighbourhood has no negative component (hence no -0) and no NaN: FMNMX returns the pinned min / max
    vec3 mn = nb.v[0][0], mx = nb.v[0][0];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (FAST) {
                mn = v3(fmin_nn(mn.x, nb.v[i][j].x), fmin_nn(mn.y, nb.v[i][j].y), fmin_nn(mn.z, nb.v[i][j].z));
                mx = v3(fmax_nn(mx.x, nb.v[i][j].x), fmax_nn(mx.y, nb.v[i][j].y), fmax_nn(mx.z, nb.v[i][j].z));
            } else {
                mn = vmin(mn, nb.v[i][j]); mx = vmax(mx, nb.v[i][j]);
            }
        }
    // resolveColor, temporalFilter.comp:41-57 (weights staged in shared memory by the kernel)
    vec3 currentColor = v3(0.f);
    currentColor = currentColor + nb.v[0][0] * sRw[0]; currentColor = currentColor + nb.v[1][0] * sRw[1]; currentColor = currentColor + nb.v[2][0] * sRw[2];
    currentColor = currentColor + nb.v[0][1] * sRw[3]; currentColor = currentColor + nb.v[1][1] * sRw[4]; currentColor = currentColor + nb.v[2][1] * sRw[5];
    currentColor = currentColor + nb.v[0][2] * sRw[6]; currentColor = currentColor + nb.v[1][2] * sRw[7]; currentColor = currentColor + nb.v[2][2] * sRw[8];

    vec2 motion;
    if (p.useMotionVectorDilation) {  // getClosestFragmentMotion, temporalReprojection.inc:67-83
        float closestDepth = 0.f;
        int offX = 0, offY = 0;
#pragma unroll
        for (int x = -1; x <= 1; x++)
#pragma unroll
            for (int y = -1; y <= 1; y++) {
                const float depth = (FAST || inRange(p.depthBuffer, ix + x, iy + y)) ? loadD32(p.depthBuffer, ix + x, iy + y) : 0.f;
                if (depth > closestDepth) { closestDepth = depth; offX = x; offY = y; }
            }
        motion = (FAST || inRange(p.motionBuffer, ix + offX, iy + offY)) ? loadRG16SNORM(p.motionBuffer, ix + offX, iy + offY) : v2(0.f);
    } else {
        motion = (FAST || inRange(p.motionBuffer, ix, iy)) ? loadRG16SNORM(p.motionBuffer, ix, iy) : v2(0.f);
    }

    // texture(historyBuffer, uv) with the linear clamp sampler
    auto H = [&](float x, float y) {
        if (FAST) {  // finite, bounded coordinates: no sanitising; indices inside the tile need no clamp
            const AxisTap X = axisTapT<false, true>(x, p.historySrc.w, hisTile.x0), Y = axisTapT<false, true>(y, p.historySrc.h, hisTile.y0);
            return tapR11Tile(hisTile, p.historySrc, X, Y);
        }
        return sampleR11LinearClampTile(hisTile, p.historySrc, v2(x, y));
    };
    vec3 historySample;
    const int historySampleTech = TECH >= 0 ? TECH : p.historySampleTech;  // specialisation constant 2 -> template parameter
    if (historySampleTech == 0) {
        { const vec2 hu = uv + motion; historySample = H(hu.x, hu.y); }
    } else if (historySampleTech == 1) {  // 16 tap, bicubicSampling.inc:28-67
        const vec2 pp = iUVf + 0.5f + motion * screenRes;
        const vec2 uvTrunc = vfloor(pp - 0.5f) + 0.5f;
        const vec2 ad = vabs(pp - uvTrunc);
        const vec2 w[4] = {v2(catmullRomWeight1D(ad.x + 1.f), catmullRomWeight1D(ad.y + 1.f)), v2(catmullRomWeight1D(ad.x), catmullRomWeight1D(ad.y)),
                           v2(catmullRomWeight1D(1.f - ad.x), catmullRomWeight1D(1.f - ad.y)), v2(catmullRomWeight1D(2.f - ad.x), catmullRomWeight1D(2.f - ad.y))};
        const vec2 u[4] = {(uvTrunc - 1.f) * texelSize, uvTrunc * texelSize, (uvTrunc + 1.f) * texelSize, (uvTrunc + 2.f) * texelSize};
        vec3 acc = v3(0.f);
        bool first = true;
        for (int yy = 0; yy < 4; yy++)
            for (int xx = 0; xx < 4; xx++) {
                const vec3 term = H(u[xx].x, u[yy].y) * w[xx].x * w[yy].y;
                acc = first ? term : acc + term;
                first = false;
            }
        historySample = acc;
    } else if (historySampleTech == 2) {  // 9 tap :72-107
        const BicubicW b = bicubicWeights(iUVf + 0.5f + motion * screenRes);
        const vec2 uv0 = (b.uvTrunc - 1.f) * texelSize, uvT = (b.uvTrunc + b.t) * texelSize, uv3 = (b.uvTrunc + 2.f) * texelSize;
        historySample = H(uv0.x, uv0.y) * b.w0.x * b.w0.y + H(uv0.x, uvT.y) * b.w0.x * b.wB.y + H(uv0.x, uv3.y) * b.w0.x * b.w3.y +
                        H(uvT.x, uv0.y) * b.wB.x * b.w0.y + H(uvT.x, uvT.y) * b.wB.x * b.wB.y + H(uvT.x, uv3.y) * b.wB.x * b.w3.y +
                        H(uv3.x, uv0.y) * b.w3.x * b.w0.y + H(uv3.x, uvT.y) * b.w3.x * b.wB.y + H(uv3.x, uv3.y) * b.w3.x * b.w3.y;
    } else if (historySampleTech == 3) {  // 5 tap :112-145
        const BicubicW b = bicubicWeights(iUVf + 0.5f + motion * screenRes);
        const vec2 uv0 = (b.uvTrunc - 1.f) * texelSize, uvT = (b.uvTrunc + b.t) * texelSize, uv3 = (b.uvTrunc + 2.f) * texelSize;
        auto T = [&](float x, float y) { return v4(H(x, y), 1.f); };
        const vec4 result = T(uv0.x, uvT.y) * b.w0.x * b.wB.y + T(uvT.x, uv0.y) * b.wB.x * b.w0.y + T(uvT.x, uvT.y) * b.wB.x * b.wB.y +
                            T(uvT.x, uv3.y) * b.wB.x * b.w3.y + T(uv3.x, uvT.y) * b.w3.x * b.wB.y;
        historySample = xyz(result) / result.w;
    } else if (historySampleTech == 4) {  // 1 tap :150-181
        const BicubicW b = bicubicWeights(iUVf + 0.5f + motion * screenRes);
        const vec2 uvT = (b.uvTrunc + b.t) * texelSize;
        const vec3 hs = H(uvT.x, uvT.y);
        const vec4 result = v4(hs + nb.v[0][1] - nb.v[1][1], 1.f) * b.w0.x * b.wB.y + v4(hs + nb.v[1][0] - nb.v[1][1], 1.f) * b.wB.x * b.w0.y +
                            v4(hs, 1.f) * b.wB.x * b.wB.y + v4(hs + nb.v[1][2] - nb.v[1][1], 1.f) * b.wB.x * b.w3.y +
                            v4(hs + nb.v[2][1] - nb.v[1][1], 1.f) * b.w3.x * b.wB.y;
        historySample = xyz(result) / result.w;
    } else {
        historySample = v3(1.f, 0.f, 0.f);
    }
    if (TONEMAP) historySample = taaTonemap(historySample);
    if (p.useClipping) historySample = clipAABB<FAST>(historySample, mn, mx);
    else historySample = vclamp(historySample, mn, mx);
    if (anynan(historySample)) historySample = currentColor;

    const float currentContrast = neighbourhoodContrast(nb);
    // gaussianFilteredNeighbourhood :71-82, used below when the history lies outside the image (same operands either way)
    const vec3 gaussian = nb.v[0][0] * 0.0625f + nb.v[0][2] * 0.0625f + nb.v[2][0] * 0.0625f + nb.v[2][2] * 0.0625f + nb.v[1][0] * 0.125f +
                          nb.v[0][1] * 0.125f + nb.v[1][2] * 0.125f + nb.v[2][1] * 0.125f + nb.v[1][1] * 0.25f;
    const float lastContrast = neighbourhoodContrastAt<TONEMAP, FAST>(hisTile, p.historySrc, uv + motion, texelSize);
    float contrastChange = absf(currentContrast - lastContrast);
    contrastChange = clampf(contrastChange, 0.f, 1.f);
    const float blendMin = 0.03f, blendMax = 0.13f;
    float blendFactor = mixf(blendMax, blendMin, contrastChange);
    if (p.g->cameraCut) blendFactor = 1.f;
    const vec2 ur = uv + motion;
    if (ur.x < 0.f || ur.y < 0.f || ur.x > 1.f || ur.y > 1.f) {  // isUVOutOfImage
        blendFactor = 1.f;
        currentColor = gaussian;
    }
    vec3 color = vmix(historySample, currentColor, blendFactor);
    if (TONEMAP) color = taaTonemapReverse(color);
    const uint32_t packed = packR11G11B10(color);
    if (FAST || inRange(p.historyDst, ix, iy)) ((uint32_t*)p.historyDst.ptr)[texelIndex(p.historyDst, ix, iy)] = packed;
    ((uint32_t*)p.outputImage.ptr)[texelIndex(p.outputImage, ix, iy)] = packed;
}
template <bool TONEMAP, int TECH>
__device__ __noinline__ void temporalFilterPixelGeneric(const TaaParams& p, const float4* sCur, const float4* sHis, const float* sRw) {
    temporalFilterPixel<TONEMAP, TECH, false>(p, sCur, sHis, sRw);
}
template <bool TONEMAP, int TECH>
__global__ void __launch_bounds__(256, 4) temporalFilterKernel(const __grid_constant__ TaaParams p) {
    __shared__ float4 sCur[TAA_CUR_W * TAA_CUR_H];
    __shared__ float4 sHis[TAA_HIS_W * TAA_HIS_H];
    __shared__ float sRw[9];
    if (threadIdx.x < 9) sRw[threadIdx.x] = p.resolveWeights[threadIdx.x];
    const int bx = blockIdx.x * 32, by = p.y0 + blockIdx.y * 8;
    const bool interior = p.sameExtents && tileIsInterior<TAA_HIS_W, TAA_HIS_H>(p.historySrc, bx - TAA_HIS_HALO, by - TAA_HIS_HALO);
    bool special;
    if (interior) {
        special = tileLoadR11<TAA_CUR_W, TAA_CUR_H, true>(sCur, p.currentFrame, bx - TAA_CUR_HALO, by - TAA_CUR_HALO);
        special = tileLoadR11<TAA_HIS_W, TAA_HIS_H, true>(sHis, p.historySrc, bx - TAA_HIS_HALO, by - TAA_HIS_HALO) || special;
    } else {
        special = tileLoadR11<TAA_CUR_W, TAA_CUR_H, false>(sCur, p.currentFrame, bx - TAA_CUR_HALO, by - TAA_CUR_HALO);
        special = tileLoadR11<TAA_HIS_W, TAA_HIS_H, false>(sHis, p.historySrc, bx - TAA_HIS_HALO, by - TAA_HIS_HALO) || special;
    }
    const bool anySpecial = __syncthreads_or(special ? 1 : 0) != 0;
    if (interior && !anySpecial) temporalFilterPixel<TONEMAP, TECH, true>(p, sCur, sHis, sRw);
    else temporalFilterPixelGeneric<TONEMAP, TECH>(p, sCur, sHis, sRw);
}
PLAIN_PASS(launch_temporalFilter, "temporalFilter.comp") {
    TaaParams p;
    p.currentFrame = c.sampled(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.historySrc = c.sampled(3, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.motionBuffer = c.sampled(4, PLAIN_FORMAT_RG16_SNORM);
    p.depthBuffer = c.sampled(5, PLAIN_FORMAT_DEPTH32);
    p.outputImage = c.storage(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.historyDst = c.storage(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.resolveWeights = c.ubuf<float>(6);
    p.g = c.g;
    p.useClipping = c.specBool(0, false);
    p.useMotionVectorDilation = c.specBool(1, false);
    p.historySampleTech = c.spec<int>(2, 0);
    p.useTonemap = c.specBool(3, false);
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < p.outputImage.w || (int)c.exec->dispatch[1] * 8 < p.outputImage.h) { c.fail("temporalFilter.comp: dispatch does not cover the target"); return; }
    c.window(p.outputImage.h, p.y0, p.y1);
    if (p.y1 <= p.y0) return;
    auto same = [&](const ImgView& v) { return v.w == p.outputImage.w && v.h == p.outputImage.h; };
    p.sameExtents = same(p.currentFrame) && same(p.historySrc) && same(p.motionBuffer) && same(p.depthBuffer) && same(p.historyDst);
    dim3 grid(ceilDiv(p.outputImage.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8));
    if (p.useTonemap && p.historySampleTech == 4) PLAIN_LAUNCH(c, (temporalFilterKernel<true, 4>), grid, 256, 0, p);  // the reference's defaults (TAA.h:8-17)
    else if (p.useTonemap) PLAIN_LAUNCH(c, (temporalFilterKernel<true, -1>), grid, 256, 0, p);
    else PLAIN_LAUNCH(c, (temporalFilterKernel<false, -1>), grid, 256, 0, p);
}

// ---------------- colorToLuminance.comp ----------------
// R11G11B10 colour -> R8 luminance (the contrast test of temporalSupersampling.comp reads it with textureGather)
__global__ void __launch_bounds__(256) colorToLuminanceKernel(ImgView dst, ImgView src, int yBegin, int yEnd) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = yBegin + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= dst.w || iy >= yEnd) return;
    const vec3 color = inRange(src, ix, iy) ? loadR11(src, ix, iy) : v3(0.f);  // texelFetch
    ((uint8_t*)dst.ptr)[texelIndex(dst, ix, iy)] = (uint8_t)floatToUnorm8(computeLuminance(color));
}
PLAIN_PASS(launch_colorToLuminance, "colorToLuminance.comp") {
    const ImgView src = c.sampled(0, PLAIN_FORMAT_R11G11B10_UFLOAT), dst = c.storage(1, PLAIN_FORMAT_R8);
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < dst.w || (int)c.exec->dispatch[1] * 8 < dst.h) { c.fail("colorToLuminance.comp: dispatch does not cover the target"); return; }
    int y0, y1;
    c.window(dst.h, y0, y1);
    if (y1 <= y0) return;
    PLAIN_LAUNCH(c, colorToLuminanceKernel, dim3(ceilDiv(dst.w, 32), ceilDiv((unsigned)(y1 - y0), 8)), 256, 0, dst, src, y0, y1);
}

// ---------------- temporalSupersampling.comp ----------------
struct SupersampleParams {
    ImgView currentFrame, lastFrame, targetImage, velocityBuffer, currentDepth, lastDepth, currentLuminance, lastLuminance;
    const plain_global_shader_info* g;
    int y0, y1;
};
__device__ __forceinline__ float minAbsoluteDifference(float s, vec4 v) {  // :22-28, as written (differences of absolute values)
    return fminp(absf(s) - absf(v.x), fminp(absf(s) - absf(v.y), fminp(absf(s) - absf(v.z), absf(s) - absf(v.w))));
}
__device__ __forceinline__ float luminanceBlockDifference(vec4 cur, vec4 last) {  // :30-36
    return minAbsoluteDifference(cur.x, last) + minAbsoluteDifference(cur.y, last) + minAbsoluteDifference(cur.z, last) + minAbsoluteDifference(cur.w, last);
}
// textureGather on an R8 image with the nearest-clamp sampler: (i0,j1), (i1,j1), (i1,j0), (i0,j0), i0 = floor(u*w - 0.5)
__device__ __forceinline__ vec4 gatherR8Clamp(const ImgView& t, vec2 uv) {
    const int x0 = f2i(floorf_(sanitizeCoord(uv.x) * (float)t.w - 0.5f)), y0 = f2i(floorf_(sanitizeCoord(uv.y) * (float)t.h - 0.5f));
    const int xa = iclamp(x0, 0, t.w - 1), xb = iclamp(x0 + 1, 0, t.w - 1), ya = iclamp(y0, 0, t.h - 1), yb = iclamp(y0 + 1, 0, t.h - 1);
    return v4(loadR8(t, xa, yb), loadR8(t, xb, yb), loadR8(t, xb, ya), loadR8(t, xa, ya));
}
// :38-55: closest (largest, reverse z) depth of the 3x3 nearest-clamp taps around uv, linearised
__device__ __forceinline__ float closestNeighbourhoodDepth(const ImgView& depthBuffer, vec2 uv, vec2 texelSize, float nearP, float farP) {
    auto tap = [&](float dx, float dy) {
        return sampleNearest2D<WRAP_CLAMP, float>([&](int x, int y) { return loadD32(depthBuffer, x, y); }, depthBuffer.w, depthBuffer.h, uv + v2(dx, dy) * texelSize, 0.f);
    };
    float closestDepth = tap(-1.f, -1.f);
    closestDepth = fmaxp(tap(0.f, -1.f), closestDepth);
    closestDepth = fmaxp(tap(1.f, -1.f), closestDepth);
    closestDepth = fmaxp(tap(-1.f, 0.f), closestDepth);
    closestDepth = fmaxp(tap(0.f, 0.f), closestDepth);
    closestDepth = fmaxp(tap(1.f, 0.f), closestDepth);
    closestDepth = fmaxp(tap(-1.f, 1.f), closestDepth);
    closestDepth = fmaxp(tap(0.f, 1.f), closestDepth);
    closestDepth = fmaxp(tap(1.f, 1.f), closestDepth);
    return linearizeDepth(closestDepth, nearP, farP);
}
template <bool TONEMAP>
__global__ void __launch_bounds__(256) temporalSupersamplingKernel(const __grid_constant__ SupersampleParams p) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.targetImage.w || iy >= p.y1) return;
    const plain_global_shader_info* g = p.g;
    const vec2 texelSize = 1.f / v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    const vec2 uvCurrent = (v2((float)ix, (float)iy) + v2(0.5f)) * texelSize;
    float closestDepth = 0.f;  // getClosestFragmentMotion, temporalReprojection.inc:67-83
    int offX = 0, offY = 0;
    for (int x = -1; x <= 1; x++)
        for (int y = -1; y <= 1; y++) {
            const float depth = inRange(p.currentDepth, ix + x, iy + y) ? loadD32(p.currentDepth, ix + x, iy + y) : 0.f;
            if (depth > closestDepth) { closestDepth = depth; offX = x; offY = y; }
        }
    const vec2 motion = inRange(p.velocityBuffer, ix + offX, iy + offY) ? loadRG16SNORM(p.velocityBuffer, ix + offX, iy + offY) : v2(0.f);
    const vec2 uvLast = uvCurrent + motion;
    vec3 currentSample = sampleR11LinearClamp(p.currentFrame, uvCurrent);
    vec3 lastSample = sampleR11LinearClamp(p.lastFrame, uvLast);
    if (TONEMAP) {
        currentSample = taaTonemap(currentSample);
        lastSample = taaTonemap(lastSample);
    }
    // acceptLastFrameSample :57-84
    const float contrast = luminanceBlockDifference(gatherR8Clamp(p.currentLuminance, uvCurrent), gatherR8Clamp(p.lastLuminance, uvLast));
    const bool contrastTest = contrast < 0.5f;
    const float currentDepth = closestNeighbourhoodDepth(p.currentDepth, uvCurrent, texelSize, g->nearPlane, g->farPlane);
    const float lastDepth = closestNeighbourhoodDepth(p.lastDepth, uvLast, texelSize, g->nearPlane, g->farPlane);
    const bool depthTest = absf(currentDepth - lastDepth) < 1.f;
    const bool outOfScreen = uvLast.x < 0.f || uvLast.y < 0.f || uvLast.x > 1.f || uvLast.y > 1.f;
    const bool acceptSample = contrastTest && depthTest && !outOfScreen;
    vec3 color = vmix(currentSample, lastSample, acceptSample ? 0.5f : 0.f);
    if (TONEMAP) color = taaTonemapReverse(color);
    storeR11(p.targetImage, ix, iy, color);
}
PLAIN_PASS(launch_temporalSupersampling, "temporalSupersampling.comp") {
    SupersampleParams p;
    p.currentFrame = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.lastFrame = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.targetImage = c.storage(3, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.velocityBuffer = c.sampled(4, PLAIN_FORMAT_RG16_SNORM);
    p.currentDepth = c.sampled(5, PLAIN_FORMAT_DEPTH32);
    p.lastDepth = c.sampled(6, PLAIN_FORMAT_DEPTH32);
    p.currentLuminance = c.sampled(7, PLAIN_FORMAT_R8);
    p.lastLuminance = c.sampled(8, PLAIN_FORMAT_R8);
    p.g = c.g;
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < p.targetImage.w || (int)c.exec->dispatch[1] * 8 < p.targetImage.h) { c.fail("temporalSupersampling.comp: dispatch does not cover the target"); return; }
    c.window(p.targetImage.h, p.y0, p.y1);
    if (p.y1 <= p.y0) return;
    dim3 grid(ceilDiv(p.targetImage.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8));
    if (c.specBool(0, false)) PLAIN_LAUNCH(c, temporalSupersamplingKernel<true>, grid, 256, 0, p);
    else PLAIN_LAUNCH(c, temporalSupersamplingKernel<false>, grid, 256, 0, p);
}

}  // namespace pb
