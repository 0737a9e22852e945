// passes_depth.cu - min/max depth pyramid, sun light matrices, depth downscale.
//   depthHiZPyramid.comp:52-124 (computeMinMax), :130-350 (main); lightMatrix.comp:57-138; depthDownscale.comp:12-20
#include "shader_inc.cuh"

namespace pb {

// ---------------- depthHiZPyramid.comp ----------------
// Result definition: every texel of a level is computed from the complete previous level with
// computeMinMax's taps (2x2, plus the extra row / column / corner when the source extent is odd, :86-121, including the
// '*' of :114 on the corner tap). min/max of binary32 are exact, so the pyramid is bit-exact whatever the schedule.
//
// Schedule: the leading levels whose source extents are even (up to four: a 32x32 depth tile -> 16x16 -> 8x8 -> 4x4
// -> 2x2) are produced by one pass over the depth buffer, each block reducing its tile through shared memory - the
// depth buffer is read exactly once with 64-bit loads. The remaining levels are small; each gets its own launch until
// a level fits one block, which then finishes the chain alone.
struct HizLevels {
    ImgView mip[12];   // target views, level 0 = half resolution
    int count;
};

__device__ __forceinline__ void accumDepth(float d, float& mn, float& mx) {
    const float isSky = (d == 0.f) ? 1.f : 0.f;
    mn = fminp(mn, d + isSky);
    mx = fmaxp(mx, d);
}
__device__ __forceinline__ void accumMinMax(vec2 t, float& mn, float& mx) {
    mn = fminp(mn, t.x + ((t.y == 0.f) ? 1.f : 0.f));
    mx = fmaxp(mx, t.y);
}

template <int FUSED>
__global__ void __launch_bounds__(256) hizFusedKernel(ImgView depth, HizLevels L, int blockRowOffset) {
    __shared__ float2 s0[16][16];
    __shared__ float2 s1[8][8];
    __shared__ float2 s2[4][4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int blockRow = (int)blockIdx.y + blockRowOffset;  // row sharding: block rows of 32 depth rows
    const int x = blockIdx.x * 16 + tx, y = blockRow * 16 + ty;
    float mn = 1.f, mx = 0.f;
    if (x < L.mip[0].w && y < L.mip[0].h) {
        const float* r0 = (const float*)depth.ptr + (size_t)(2 * y) * depth.w + 2 * x;
        const float2 a = __ldg((const float2*)r0), b = __ldg((const float2*)(r0 + depth.w));
        accumDepth(a.x, mn, mx); accumDepth(a.y, mn, mx); accumDepth(b.x, mn, mx); accumDepth(b.y, mn, mx);
        storeRG32F(L.mip[0], x, y, v2(mn, mx));
    }
    if (FUSED < 2) return;
    s0[ty][tx] = make_float2(mn, mx);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int qx = threadIdx.x & 7, qy = threadIdx.x >> 3;
        const int ox = blockIdx.x * 8 + qx, oy = blockRow * 8 + qy;
        float n1 = 1.f, x1 = 0.f;
        if (ox < L.mip[1].w && oy < L.mip[1].h) {
            for (int j = 0; j < 2; j++)
                for (int i = 0; i < 2; i++) { const float2 t = s0[2 * qy + j][2 * qx + i]; accumMinMax(v2(t.x, t.y), n1, x1); }
            storeRG32F(L.mip[1], ox, oy, v2(n1, x1));
        }
        s1[qy][qx] = make_float2(n1, x1);
    }
    if (FUSED < 3) return;
    __syncthreads();
    if (threadIdx.x < 16) {
        const int qx = threadIdx.x & 3, qy = threadIdx.x >> 2;
        const int ox = blockIdx.x * 4 + qx, oy = blockRow * 4 + qy;
        float n2 = 1.f, x2 = 0.f;
        if (ox < L.mip[2].w && oy < L.mip[2].h) {
            for (int j = 0; j < 2; j++)
                for (int i = 0; i < 2; i++) { const float2 t = s1[2 * qy + j][2 * qx + i]; accumMinMax(v2(t.x, t.y), n2, x2); }
            storeRG32F(L.mip[2], ox, oy, v2(n2, x2));
        }
        s2[qy][qx] = make_float2(n2, x2);
    }
    if (FUSED < 4) return;
    __syncthreads();
    if (threadIdx.x < 4) {
        const int qx = threadIdx.x & 1, qy = threadIdx.x >> 1;
        const int ox = blockIdx.x * 2 + qx, oy = blockRow * 2 + qy;
        if (ox < L.mip[3].w && oy < L.mip[3].h) {
            float n3 = 1.f, x3 = 0.f;
            for (int j = 0; j < 2; j++)
                for (int i = 0; i < 2; i++) { const float2 t = s2[2 * qy + j][2 * qx + i]; accumMinMax(v2(t.x, t.y), n3, x3); }
            storeRG32F(L.mip[3], ox, oy, v2(n3, x3));
        }
    }
}

// one texel of a level from its source with the reference's taps; nearest + clamp-to-edge addressing
template <bool FROM_DEPTH>
__device__ __forceinline__ vec2 hizTexel(const ImgView& src, int x, int y) {
    const bool extraRow = (src.h & 1) == 1, extraColumn = (src.w & 1) == 1;
    float mn = 1.f, mx = 0.f;
    auto tap = [&](int ox, int oy, bool cornerQuirk) {
        const int sx = imin(2 * x + ox, src.w - 1), sy = imin(2 * y + oy, src.h - 1);
        if (FROM_DEPTH) {
            const float d = ((const float*)src.ptr)[(size_t)sy * src.w + sx];
            const float isSky = (d == 0.f) ? 1.f : 0.f;
            mn = fminp(mn, cornerQuirk ? d * isSky : d + isSky);  // :114 multiplies on the odd x odd corner texel
            mx = fmaxp(mx, d);
        } else {
            const float2 t = ((const float2*)src.ptr)[(size_t)sy * src.w + sx];
            accumMinMax(v2(t.x, t.y), mn, mx);
        }
    };
    tap(0, 0, false); tap(1, 0, false); tap(0, 1, false); tap(1, 1, false);
    if (extraRow) { tap(0, 2, false); tap(1, 2, false); }
    if (extraColumn) { tap(2, 0, false); tap(2, 1, false); }
    if (extraRow && extraColumn) tap(2, 2, true);
    return v2(mn, mx);
}

template <bool FROM_DEPTH>
__global__ void __launch_bounds__(256) hizLevelKernel(ImgView src, ImgView dst) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= dst.w || y >= dst.h) return;
    storeRG32F(dst, x, y, hizTexel<FROM_DEPTH>(src, x, y));
}

// finishes the chain in one block: levels [first, count), each from the previous one (global memory, block-level sync)
__global__ void __launch_bounds__(1024) hizTailKernel(HizLevels L, int first) {
    for (int level = first; level < L.count; level++) {
        const ImgView src = L.mip[level - 1], dst = L.mip[level];
        const int n = dst.w * dst.h;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int x = i % dst.w, y = i / dst.w;
            storeRG32F(dst, x, y, hizTexel<false>(src, x, y));
        }
        __syncthreads();
    }
}

PLAIN_PASS(launch_depthHiZPyramid, "depthHiZPyramid.comp") {
    const int mipCount = c.spec<int>(0, 0);
    const int resX = c.spec<int>(1, 0), resY = c.spec<int>(2, 0);
    const ImgView depth = c.sampled(13, PLAIN_FORMAT_DEPTH32);
    if (c.failed) return;
    // the reference binds 11 levels (depthHiZPyramid.comp:16-31: up to 4096x4096); a 12th level at binding 11 is this build's
    // extension for BASELINE configs[4] (7680x4320: the half-resolution pyramid has 12 levels), defined level by level like the rest
    if (mipCount < 1 || mipCount > 12) { c.fail("depthHiZPyramid.comp: mip count must be 1..12 (depthHiZPyramid.comp:16-19, + one extension level)"); return; }
    const int bindingCount = mipCount > 11 ? mipCount : 11;
    if (depth.w != resX || depth.h != resY) { c.fail("depthHiZPyramid.comp: depth buffer extent differs from the specialisation constants"); return; }
    HizLevels L;
    L.count = mipCount;
    int srcW = resX, srcH = resY;
    int fused = 0;
    bool evenSoFar = true;
    for (int k = 0; k < mipCount; k++) {
        // the shader's image binding k' = 11 - mipCount + k holds pyramid level k (RenderFrontend.cpp:825-833)
        L.mip[k] = c.storage((uint32_t)(bindingCount - mipCount + k), PLAIN_FORMAT_RG32_SFLOAT);
        if (c.failed) return;
        const int w = std::max(srcW / 2, 1), h = std::max(srcH / 2, 1);
        if (L.mip[k].w != w || L.mip[k].h != h) { c.fail("depthHiZPyramid.comp: pyramid level extent mismatch"); return; }
        evenSoFar = evenSoFar && (srcW % 2 == 0) && (srcH % 2 == 0) && srcW >= 2 && srcH >= 2;
        if (evenSoFar && k < 4) fused = k + 1;
        srcW = w; srcH = h;
    }
    // Row sharding (shard_phase): 1 = only the fused levels, for level-0 rows [row_begin, row_end) (multiples of 16: the
    // rank's own depth rows); 2 = only the remaining levels (replicated on every rank after the last fused level has been
    // all-gathered); 0 = everything.
    const uint32_t phase = c.exec->shardPhase;
    int next = 0;
    if (fused > 0) {
        int y0, y1;
        c.window(L.mip[0].h, y0, y1);
        if (phase == 2) y1 = y0;
        if (y0 % 16 != 0) { c.fail("depthHiZPyramid.comp: row window must start at a multiple of 16 level-0 rows"); return; }
        dim3 grid(ceilDiv(L.mip[0].w, 16), ceilDiv((unsigned)(y1 - y0), 16));
        if (y1 > y0) switch (fused) {
            case 1: PLAIN_LAUNCH(c, hizFusedKernel<1>, grid, 256, 0, depth, L, y0 / 16); break;
            case 2: PLAIN_LAUNCH(c, hizFusedKernel<2>, grid, 256, 0, depth, L, y0 / 16); break;
            case 3: PLAIN_LAUNCH(c, hizFusedKernel<3>, grid, 256, 0, depth, L, y0 / 16); break;
            default: PLAIN_LAUNCH(c, hizFusedKernel<4>, grid, 256, 0, depth, L, y0 / 16); break;
        }
        next = fused;
    } else if (phase != 0) {
        c.fail("depthHiZPyramid.comp: row sharding needs even depth extents (no level can be reduced from a rank's own rows)");
        return;
    }
    if (phase == 1) return;
    while (next < mipCount && (next == 0 || L.mip[next].w * L.mip[next].h > 1024)) {
        dim3 grid(ceilDiv(L.mip[next].w, 16), ceilDiv(L.mip[next].h, 16));
        if (next == 0) PLAIN_LAUNCH(c, hizLevelKernel<true>, grid, 256, 0, depth, L.mip[0]);
        else PLAIN_LAUNCH(c, hizLevelKernel<false>, grid, 256, 0, L.mip[next - 1], L.mip[next]);
        next++;
    }
    if (next < mipCount) PLAIN_LAUNCH(c, hizTailKernel, 1, 1024, 0, L, next);
    // the reference's sync counter (binding 16) is reset to 0 by its last workgroup (:261); it is never raised here
}

// ---------------- lightMatrix.comp:57-138 (single thread in the reference) ----------------
struct M4 { vec4 c[4]; };
__device__ vec4 mulM4(const M4& m, vec4 v) { return vfma(m.c[3], v.w, vfma(m.c[2], v.z, vfma(m.c[1], v.y, m.c[0] * v.x))); }  // the contract's M * v
__device__ M4 mulM4(const M4& a, const M4& b) { M4 r; for (int j = 0; j < 4; j++) r.c[j] = mulM4(a, b.c[j]); return r; }
__device__ float& comp(vec4& v, int i) { return (&v.x)[i]; }

__global__ void lightMatrixKernel(plain_shadow_cascade_info* info, ImgView depthMinMaxLowestMip, const plain_global_shader_info* __restrict__ g,
                                  uint32_t sunShadowCascadeCount, float highestCascadeExtraPadding, float highestCascadeMinFarPlane) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float FLOAT_MAX = 3.402823466e+38f, FLOAT_MIN = 1.175494351e-38f;
    const Globals G = loadGlobals(g);
    M4 coordinateSystemCorrection;
    coordinateSystemCorrection.c[0] = v4(1.0f, 0.0f, 0.0f, 0.0f);
    coordinateSystemCorrection.c[1] = v4(0.0f, 1.0f, 0.0f, 0.0f);
    coordinateSystemCorrection.c[2] = v4(0.0f, 0.0f, -0.5f, 0.f);
    coordinateSystemCorrection.c[3] = v4(0.0f, 0.0f, 0.5f, 1.0f);

    const vec3 forward = -v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2]);
    vec3 up = absf(forward.y) < 0.9999f ? v3(0.f, -1.f, 0.f) : v3(0.f, 0.f, -1.f);
    const vec3 right = cross(forward, up);
    up = cross(right, forward);
    const vec3 rn = normalize(right), un = normalize(up);
    // V = transpose(mat4(vec4(rn,0), vec4(un,0), vec4(forward,0), vec4(0,0,0,1)))
    M4 V;
    V.c[0] = v4(rn.x, un.x, forward.x, 0.f);
    V.c[1] = v4(rn.y, un.y, forward.y, 0.f);
    V.c[2] = v4(rn.z, un.z, forward.z, 0.f);
    V.c[3] = v4(0.f, 0.f, 0.f, 1.f);

    const vec2 depthMinMax = inRange(depthMinMaxLowestMip, 0, 0) ? loadRG32F(depthMinMaxLowestMip, 0, 0) : v2(0.f);
    const float depthMaxLinear = linearizeDepth(depthMinMax.x, G.nearPlane, G.farPlane);
    const float depthMinLinear = linearizeDepth(depthMinMax.y, G.nearPlane, G.farPlane);

    for (uint32_t i = 0; i + 1 < sunShadowCascadeCount; i++)
        info->splits[i] = depthMinLinear + ((depthMaxLinear - depthMinLinear) * (float)((int)i + 1) / (float)sunShadowCascadeCount);  // :51-53

    for (uint32_t i = 0; i < sunShadowCascadeCount; i++) {
        vec3 minP = v3(FLOAT_MAX);
        vec3 maxP = v3(FLOAT_MIN);
        float cascadeMinDepth = (i == 0) ? depthMinLinear : info->splits[i - 1];
        float cascadeMaxDepth = info->splits[i];
        if (i == sunShadowCascadeCount - 1) {
            cascadeMinDepth = G.nearPlane;
            cascadeMaxDepth = fmaxp(depthMaxLinear, highestCascadeMinFarPlane);
        }
        vec3 frustumPoints[8];  // computeFrustumPoints :29-48
        {
            const float nearD = cascadeMinDepth, farD = cascadeMaxDepth;
            const vec3 nearPlaneCenter = G.camPos + G.fwd * nearD;
            const vec3 farPlaneCenter = G.camPos + G.fwd * farD;
            const float heightNear = G.tanFovHalf * nearD, heightFar = G.tanFovHalf * farD;
            const float widthNear = heightNear * G.aspect, widthFar = heightFar * G.aspect;
            frustumPoints[0] = farPlaneCenter + G.up * heightFar + G.right * widthFar;
            frustumPoints[1] = farPlaneCenter + G.up * heightFar - G.right * widthFar;
            frustumPoints[2] = farPlaneCenter - G.up * heightFar + G.right * widthFar;
            frustumPoints[3] = farPlaneCenter - G.up * heightFar - G.right * widthFar;
            frustumPoints[4] = nearPlaneCenter + G.up * heightNear + G.right * widthNear;
            frustumPoints[5] = nearPlaneCenter + G.up * heightNear - G.right * widthNear;
            frustumPoints[6] = nearPlaneCenter - G.up * heightNear + G.right * widthNear;
            frustumPoints[7] = nearPlaneCenter - G.up * heightNear - G.right * widthNear;
        }
        for (int k = 0; k < 8; k++) {
            const vec3 pTransformed = xyz(mulM4(V, v4(frustumPoints[k], 1.f)));
            minP = vmin(minP, pTransformed);
            maxP = vmax(maxP, pTransformed);
        }
        if (i == sunShadowCascadeCount - 1) {
            minP = minP - highestCascadeExtraPadding;
            maxP = maxP + highestCascadeExtraPadding;
        }
        minP = minP - PB_SHADOW_SAMPLE_RADIUS * 2.f;
        maxP = maxP + PB_SHADOW_SAMPLE_RADIUS * 2.f;
        const vec3 scale = v3(2.f) / (maxP - minP);
        const vec3 offset = -0.5f * (maxP + minP) * scale;
        M4 P;
        P.c[0] = v4(scale.x, 0, 0, 0);
        P.c[1] = v4(0, scale.y, 0, 0);
        P.c[2] = v4(0, 0, scale.z, 0);
        P.c[3] = v4(offset.x, offset.y, offset.z, 1.f);
        M4 lm = mulM4(mulM4(coordinateSystemCorrection, P), V);
        for (int cc = 0; cc < 4; cc++)
            for (int r = 0; r < 4; r++) info->lightMatrices[i][cc * 4 + r] = comp(lm.c[cc], r);
        info->lightSpaceScale[i][0] = scale.x;
        info->lightSpaceScale[i][1] = scale.y;
    }
}
PLAIN_PASS(launch_lightMatrix, "lightMatrix.comp") {
    const uint32_t cascades = c.spec<uint32_t>(0, 4);
    plain_shadow_cascade_info* info = c.sbuf<plain_shadow_cascade_info>(0);
    const ImgView lowest = c.storage(1, PLAIN_FORMAT_RG32_SFLOAT);
    if (c.failed) return;
    if (cascades < 1 || cascades > 4) { c.fail("lightMatrix.comp: cascade count must be 1..4"); return; }
    PLAIN_LAUNCH(c, lightMatrixKernel, 1, 32, 0, info, lowest, c.g, cascades, c.push<float>(0), c.push<float>(4));
}

// ---------------- depthDownscale.comp:12-20 ----------------
// uv = (2*x + 0.5) / size lands half a texel inside texel 2*x: the nearest fetch is an exact strided copy.
__global__ void __launch_bounds__(256) depthDownscaleKernel(ImgView dst, ImgView src, int limitX, int limitY, int y0) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dst.w || y >= limitY || x >= limitX) return;
    const int sx = imin(2 * x, src.w - 1), sy = imin(2 * y, src.h - 1);
    storeR16F(dst, x, y, loadD32(src, sx, sy));
}
PLAIN_PASS(launch_depthDownscale, "depthDownscale.comp") {
    const ImgView dst = c.storage(0, PLAIN_FORMAT_R16_SFLOAT);
    const ImgView src = c.sampled(1, PLAIN_FORMAT_DEPTH32);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    int y0, y1;
    c.window(std::min(dst.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(std::min(dst.w, limX), 32), ceilDiv((unsigned)(y1 - y0), 8));
    PLAIN_LAUNCH(c, depthDownscaleKernel, grid, 256, 0, dst, src, limX, y1, y0);
}

}  // namespace pb
