// passes_raster.cu - the rasterisation passes that feed the frame path as a visibility-buffer software
// rasteriser for sm_100a:
//   depthPrepass.vert + depthPrepass.frag   depth (D32F reverse z), motion (RG16_SNORM), geometric normal (RGBA8)
//   sunShadow.vert + sunShadow.frag         one D16 shadow cascade (front faces culled, depth clamp)
//   triangle.vert + gbufferFill.frag        the packed G-buffer (include/plain_frame_types.h): the interpolated inputs and material
//                                           texels of triangle.frag:178-193 for the fragment that won the prepass (depth EQUAL)
// Per pass: (0) rasterVertexKernel, a thread per (draw, vertex): the vertex stage once per vertex into a post-transform cache
// (64 B per entry: clip position + what the fragment stage interpolates, tangent frame already normalised). (1) rasterSetupKernel,
// a thread per triangle: clipping, snapping, culling, the window-space bounding box. Three classes by that box: TINY (<= 1024
// pixels) is rasterised by the set-up thread on the spot; BIG (taller than 64 rows or more than 8192 pixels) is appended to a
// list; the rest is SMALL. (2) rasterCoverKernel<false>, a warp per small triangle: lanes = 32 consecutive pixels of a row,
// exact 64-bit integer edge functions (rows wider than 96 pixels first narrow their span from the edge equations in binary64,
// one pixel of slack), depth from the triangle's affine depth plane in binary64, a read of the visibility texel and - only when
// the fragment would win - one 64-bit atomicMax of (depth bits << 32 | primitive + 1): the depth test GREATER_EQUAL in draw
// order. (3) rasterCoverKernel<true>, persistent warps over (big triangle, 8-row band) pairs, so a wall-sized triangle is
// spread over a few hundred warps. (4) a resolve
// kernel, a thread per pixel: the winning primitive's three cache entries, perspective-correct barycentrics from the same plane
// equations, the fragment stage, one coalesced store per attachment. The G-buffer fill reuses the prepass's visibility
// buffer (the reference rasterises everything twice, with depth test EQUAL the second time).
// The rules the reference leaves to the Vulkan rasteriser are stated in oracle/passes_raster.cpp (header); this file follows them
// operation by operation (binary64 clipping / plane equations, -fmad=false), so the results are bit-identical.
#include "pass_common.cuh"
#include "shader_inc.cuh"

namespace pb {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 crossd(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct ClipV { double x, y, z, w; };

// ---- vertex fetch: VertexInput.h:27-31, VulkanVertexInput.cpp:4-10 ----
struct VertexIn { vec3 pos; vec2 uv; vec3 normal, tangent, bitangent; };
__device__ __forceinline__ float snorm10(uint32_t bits) {
    int v = (int)(bits & 1023u);
    if (v >= 512) v -= 1024;
    return fmaxp((float)v / 511.f, -1.f);
}
__device__ __forceinline__ vec3 unpackA2R10G10B10Snorm(uint32_t p) { return v3(snorm10(p >> 20), snorm10(p >> 10), snorm10(p)); }
__device__ __forceinline__ uint32_t fetchIndex(const RasterDraw& d, uint32_t k) { return d.index32 ? ldg((const uint32_t*)d.indices + k) : (uint32_t)ldg((const uint16_t*)d.indices + k); }
__device__ __forceinline__ VertexIn fetchVertex(const RasterDraw& d, uint32_t index) {
    const uint32_t* p = (const uint32_t*)(d.vertices + (size_t)index * 28);
    VertexIn v;
    v.pos = v3(__uint_as_float(ldg(p)), __uint_as_float(ldg(p + 1)), __uint_as_float(ldg(p + 2)));
    const uint32_t uv = ldg(p + 3);
    v.uv = v2(halfToFloat((uint16_t)(uv & 0xffffu)), halfToFloat((uint16_t)(uv >> 16)));
    v.normal = unpackA2R10G10B10Snorm(ldg(p + 4));
    v.tangent = unpackA2R10G10B10Snorm(ldg(p + 5));
    v.bitangent = unpackA2R10G10B10Snorm(ldg(p + 6));
    return v;
}
__device__ __forceinline__ vec3 mulMat3(const float* m, vec3 v) {  // mat3(model) * v, the contract's M * v
    return vfma(v3(m[8], m[9], m[10]), v.z, vfma(v3(m[4], m[5], m[6]), v.y, v3(m[0], m[1], m[2]) * v.x));
}
__device__ __forceinline__ const RasterDraw& drawOfPrimitive(const RasterDraw* draws, uint32_t drawCount, uint32_t primitive) {
    uint32_t lo = 0, hi = drawCount;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (draws[mid].firstPrimitive <= primitive) lo = mid; else hi = mid; }
    return draws[lo];
}

struct MainPassMatrices { float model[16], mvp[16], mvpPrevious[16]; };  // MainPassMatrices.inc
struct RasterParams {
    const RasterDraw* draws;
    uint32_t drawCount, totalTris;
    uint32_t* triInfo;
    unsigned long long* vis;
    float* vertexCache;
    uint32_t totalVertices;
    int W, H;
    int yMin, yMax;                          // rows of the targets this execution renders (row sharding), inclusive
    int clipNear, clampDepth, cullMode;
    const MainPassMatrices* mainTransforms;
    const float* shadowTransforms;
    const plain_shadow_cascade_info* cascades;
    uint32_t cascade;
    int shadowKind;                          // layout of the cache entries: 1 = sunShadow (uv at [4..5]), 0 = prepass (uv at [11..12])
    const BindlessEntry* bindless;           // material textures: slot == image handle index
    const float* unorm8Table;                // exact float(b) / 255 (pass_common.cuh ShadingTables)
};
// ---- (0) vertex stage, once per (draw, vertex): depthPrepass.vert:28-42, triangle.vert:29-40, sunShadow.vert:29-32 ----
// cache entry (16 floats): [0..3] clip position; KIND 0 (prepass): [4..7] previous clip position, [8..10] N, [11..12] uv;
// KIND 1 (G-buffer fill): [4..6] T, [7..9] B, [10..12] N, [13..14] uv; KIND 2 (shadow): [4..5] uv (uv: the alpha test)
__device__ __forceinline__ const RasterDraw& drawOfVertex(const RasterDraw* draws, uint32_t drawCount, uint32_t entry) {
    uint32_t lo = 0, hi = drawCount;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (draws[mid].firstVertex <= entry) lo = mid; else hi = mid; }
    return draws[lo];
}
template <int KIND>
__global__ void __launch_bounds__(256) rasterVertexKernel(const __grid_constant__ RasterParams p) {
    const uint32_t entry = blockIdx.x * 256 + threadIdx.x;
    if (entry >= p.totalVertices) return;
    const RasterDraw& d = drawOfVertex(p.draws, p.drawCount, entry);
    const VertexIn v = fetchVertex(d, entry - d.firstVertex);
    float4* out = (float4*)(p.vertexCache + (size_t)entry * PLAIN_RASTER_VERTEX_FLOATS);
    if (KIND == 2) {
        const float* L = p.cascades->lightMatrices[p.cascade];
        const float* T = p.shadowTransforms + (size_t)d.push[1] * 16;
        float M[16];  // lightMatrices[cascadeIndex] * transforms[transformIndex]
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const vec4 c = mulm4(L, v4(T[j * 4], T[j * 4 + 1], T[j * 4 + 2], T[j * 4 + 3]));
            M[j * 4] = c.x; M[j * 4 + 1] = c.y; M[j * 4 + 2] = c.z; M[j * 4 + 3] = c.w;
        }
        const vec4 clip = mulm4(M, v4(v.pos, 1.f));
        out[0] = make_float4(clip.x, clip.y, clip.z, clip.w);
        out[1] = make_float4(v.uv.x, v.uv.y, 0.f, 0.f);
        return;
    }
    const MainPassMatrices& tr = p.mainTransforms[d.push[3]];
    const vec4 clip = mulm4(tr.mvp, v4(v.pos, 1.f));
    out[0] = make_float4(clip.x, clip.y, clip.z, clip.w);
    const vec3 N = normalize(mulMat3(tr.model, v.normal));
    if (KIND == 0) {
        const vec4 prev = mulm4(tr.mvpPrevious, v4(v.pos, 1.f));
        out[1] = make_float4(prev.x, prev.y, prev.z, prev.w);
        out[2] = make_float4(N.x, N.y, N.z, v.uv.x);
        out[3] = make_float4(v.uv.y, 0.f, 0.f, 0.f);
    } else {
        const vec3 T = normalize(mulMat3(tr.model, v.tangent)), B = normalize(mulMat3(tr.model, v.bitangent));
        out[1] = make_float4(T.x, T.y, T.z, B.x);
        out[2] = make_float4(B.y, B.z, N.x, N.y);
        out[3] = make_float4(N.z, v.uv.x, v.uv.y, 0.f);
    }
}
__device__ __forceinline__ const float4* cacheEntry(const RasterParams& p, const RasterDraw& d, uint32_t tri, int k) {
    return (const float4*)(p.vertexCache + (size_t)(d.firstVertex + fetchIndex(d, tri * 3 + k)) * PLAIN_RASTER_VERTEX_FLOATS);
}
__device__ __forceinline__ void triangleClipPositions(const RasterParams& p, const RasterDraw& d, uint32_t tri, vec4 clip[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) { const float4 c = cacheEntry(p, d, tri, k)[0]; clip[k] = v4(c.x, c.y, c.z, c.w); }
}

// ---- homogeneous plane equations of one (unclipped) triangle ----
struct TriPlanes { D3 c0, c1, c2, depth; };
__device__ __forceinline__ TriPlanes trianglePlanes(const vec4 clip[3]) {
    D3 v[3];
    double z[3], w[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { v[i] = D3{(double)clip[i].x, (double)clip[i].y, (double)clip[i].w}; z[i] = (double)clip[i].z; w[i] = (double)clip[i].w; }
    TriPlanes t;
    t.c0 = crossd(v[1], v[2]); t.c1 = crossd(v[2], v[0]); t.c2 = crossd(v[0], v[1]);
    const D3 num = D3{t.c0.x * z[0] + t.c1.x * z[1] + t.c2.x * z[2], t.c0.y * z[0] + t.c1.y * z[1] + t.c2.y * z[2], t.c0.z * z[0] + t.c1.z * z[1] + t.c2.z * z[2]};
    const double det = t.c0.z * w[0] + t.c1.z * w[1] + t.c2.z * w[2];
    t.depth = D3{num.x / det, num.y / det, num.z / det};
    return t;
}
__device__ __forceinline__ double pixelNdc(int i, int size) { return ((double)i + 0.5) * (2.0 / (double)size) - 1.0; }  // NDC of a pixel centre
__device__ __forceinline__ void barycentrics(const TriPlanes& t, double px, double py, float l[3]) {
    const double e0 = t.c0.x * px + t.c0.y * py + t.c0.z, e1 = t.c1.x * px + t.c1.y * py + t.c1.z, e2 = t.c2.x * px + t.c2.y * py + t.c2.z;
    const double inv = 1.0 / (e0 + e1 + e2);
    l[0] = (float)(e0 * inv); l[1] = (float)(e1 * inv); l[2] = (float)(e2 * inv);
}
__device__ __forceinline__ float lerp3(const float l[3], float a, float b, float c) { return fmaf_(l[2], c, fmaf_(l[1], b, l[0] * a)); }
__device__ __forceinline__ vec3 lerp3(const float l[3], vec3 a, vec3 b, vec3 c) { return v3(lerp3(l, a.x, b.x, c.x), lerp3(l, a.y, b.y, c.y), lerp3(l, a.z, b.z, c.z)); }

// ---- clipping + snapping: the window-space polygon of one triangle (fanned from its first vertex) ----
struct ScreenPoly { int n; long long X[12], Y[12]; };
__device__ __forceinline__ double clipDistance(int plane, const ClipV& p, double gx, double gy, bool clipNear) {
    switch (plane) {
        case 0: return p.w - 1e-6;
        case 1: return clipNear ? p.w - p.z : 1.0;
        case 2: return gx * p.w - p.x;
        case 3: return gx * p.w + p.x;
        case 4: return gy * p.w - p.y;
        default: return gy * p.w + p.y;
    }
}
__device__ void clipAndSnap(const RasterParams& p, const vec4 clip[3], ScreenPoly& sp) {
    ClipV poly[12], tmp[12];
    int n = 3;
    for (int i = 0; i < 3; i++) poly[i] = ClipV{(double)clip[i].x, (double)clip[i].y, (double)clip[i].z, (double)clip[i].w};
    const double gx = 32768.0 / (double)p.W, gy = 32768.0 / (double)p.H;
    for (int plane = 0; plane < 6 && n >= 3; plane++) {
        bool allIn = true;
        for (int i = 0; i < n; i++) allIn = allIn && clipDistance(plane, poly[i], gx, gy, p.clipNear != 0) >= 0.0;
        if (allIn) continue;
        int m = 0;
        for (int i = 0; i < n; i++) {
            const ClipV a = poly[i], b = poly[(i + 1) % n];
            const double da = clipDistance(plane, a, gx, gy, p.clipNear != 0), db = clipDistance(plane, b, gx, gy, p.clipNear != 0);
            if (da >= 0.0) tmp[m++] = a;
            if ((da >= 0.0) != (db >= 0.0)) {
                const double t = da / (da - db);
                tmp[m++] = ClipV{a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w)};
            }
        }
        n = m;
        for (int i = 0; i < n; i++) poly[i] = tmp[i];
    }
    sp.n = n < 3 ? 0 : n;
    for (int i = 0; i < sp.n; i++) {
        const double inv = 1.0 / poly[i].w;
        const double xs = (poly[i].x * inv * 0.5 + 0.5) * (double)p.W, ys = (poly[i].y * inv * 0.5 + 0.5) * (double)p.H;
        sp.X[i] = (long long)floor(xs * 256.0 + 0.5);
        sp.Y[i] = (long long)floor(ys * 256.0 + 0.5);
    }
}
// one triangle of the fan: orientation, culling, the pixel rectangle it can cover. Returns false when it covers nothing.
struct SubTriangle { long long ax[3], ay[3], ex[3], ey[3], bias[3]; int ix0, ix1, iy0, iy1; };
__device__ __forceinline__ long long llmin(long long a, long long b) { return a < b ? a : b; }
__device__ __forceinline__ long long llmax(long long a, long long b) { return a > b ? a : b; }
__device__ __forceinline__ bool subTriangle(const RasterParams& p, const ScreenPoly& sp, int k, SubTriangle& s) {
    long long x0 = sp.X[0], y0 = sp.Y[0], x1 = sp.X[k], y1 = sp.Y[k], x2 = sp.X[k + 1], y2 = sp.Y[k + 1];
    const long long area2 = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (area2 == 0) return false;
    const bool front = area2 < 0;  // counter clockwise on the y-down screen (VulkanPipeline.cpp:61)
    if ((p.cullMode == PLAIN_CULL_BACK && !front) || (p.cullMode == PLAIN_CULL_FRONT && front)) return false;
    if (area2 < 0) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    s.ax[0] = x0; s.ax[1] = x1; s.ax[2] = x2; s.ay[0] = y0; s.ay[1] = y1; s.ay[2] = y2;
    s.ex[0] = x1 - x0; s.ex[1] = x2 - x1; s.ex[2] = x0 - x2; s.ey[0] = y1 - y0; s.ey[1] = y2 - y1; s.ey[2] = y0 - y2;
#pragma unroll
    for (int e = 0; e < 3; e++) s.bias[e] = (s.ey[e] < 0 || (s.ey[e] == 0 && s.ex[e] > 0)) ? 0 : -1;  // top-left rule
    const long long minX = llmin(x0, llmin(x1, x2)), maxX = llmax(x0, llmax(x1, x2)), minY = llmin(y0, llmin(y1, y2)), maxY = llmax(y0, llmax(y1, y2));
    s.ix0 = (int)llmax(0, (minX - 128 + 255) >> 8); s.ix1 = (int)llmin(p.W - 1, (maxX - 128) >> 8);
    s.iy0 = (int)llmax(p.yMin, (minY - 128 + 255) >> 8); s.iy1 = (int)llmin(p.yMax, (maxY - 128) >> 8);
    return s.ix0 <= s.ix1 && s.iy0 <= s.iy1;
}

// depth of a covered pixel + the depth test: GREATER_EQUAL in draw order = maximum of (depth bits, primitive + 1)
// READ_FIRST: look at the texel before the atomic (it only grows, so a stale value can only cost an atomic): saves most atomics of
// occluded fragments in the warp-parallel paths; the serial tiny path skips it (a dependent load per pixel is all latency there)
// the alpha test of depthPrepass.frag:28-31 / sunShadow.frag:19-22 for draws whose albedo texture has transparent texels
// (out of line and self-contained - it fetches the triangle's uv itself - so that opaque draws carry nothing but the flag)
struct AlphaTest { const RasterDraw* draw; uint32_t tri; };  // draw == nullptr: opaque
__device__ __forceinline__ AlphaTest alphaTestOf(const RasterDraw& d, uint32_t tri) { return AlphaTest{d.alphaTest ? &d : nullptr, tri}; }
__device__ __noinline__ bool alphaTestDiscards(const RasterParams& p, const TriPlanes& tp, AlphaTest a, int ix, int iy) {
    vec2 uv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4* e = cacheEntry(p, *a.draw, a.tri, k);
        if (p.shadowKind) { const float4 u = e[1]; uv[k] = v2(u.x, u.y); }
        else { const float4 n = e[2], u = e[3]; uv[k] = v2(n.w, u.x); }
    }
    float l[3];
    barycentrics(tp, pixelNdc(ix, p.W), pixelNdc(iy, p.H), l);
    const vec2 passUV = v2(lerp3(l, uv[0].x, uv[1].x, uv[2].x), lerp3(l, uv[0].y, uv[1].y, uv[2].y));
    const ImgView albedo = p.bindless[a.draw->push[0]].view;
    const float alpha = sampleLinear2D<WRAP_REPEAT, float>([&](int x, int y) { return ldg(p.unorm8Table + (ldg((const uint32_t*)albedo.ptr + texelIndex(albedo, x, y)) >> 24)); },
                                                             albedo.w, albedo.h, passUV, 0.f);
    return alpha < 0.5f;
}
template <bool READ_FIRST, bool ALPHA>
__device__ __forceinline__ void emitFragment(const RasterParams& p, const TriPlanes& tp, const AlphaTest& alpha, int ix, int iy, double rowDepth, unsigned long long keyLow) {
    if (ALPHA && alpha.draw && alphaTestDiscards(p, tp, alpha, ix, iy)) return;
    float dep = (float)(tp.depth.x * pixelNdc(ix, p.W) + rowDepth);
    if (dep != dep) return;
    if (p.clampDepth) dep = dep < 0.f ? 0.f : (dep > 1.f ? 1.f : dep);
    else if (dep < 0.f || dep > 1.f) return;
    const unsigned long long key = ((unsigned long long)(__float_as_uint(dep) & 0x7fffffffu) << 32) | keyLow;
    unsigned long long* slot = p.vis + (size_t)iy * p.W + ix;
    if (!READ_FIRST || key > *(volatile unsigned long long*)slot) atomicMax(slot, key);
}

// A triangle whose window-space bounding box is at most 1024 pixels is "tiny": the set-up thread rasterises it on the spot.
#define RASTER_TINY_MAX_AREA 1024
// A triangle whose window-space bounding box is taller than 64 rows or larger than 8192 pixels is "big": it is rasterised by the
// persistent kernel in bands of 8 rows (one warp per band), so that a screen-wide wall is spread over a few hundred warps.
#define RASTER_SMALL_MAX_ROWS 64
#define RASTER_SMALL_MAX_AREA 8192
#define RASTER_BIG_BAND_ROWS 8
#define RASTER_BIG_FLAG 0x80000000u
// (1) a thread per triangle: the rows it can cover (y0 | y1 << 16 | big flag); big triangles go to the list behind the counter at triInfo[totalTris]
template <bool ALPHA>
__global__ void __launch_bounds__(128) rasterSetupKernel(const __grid_constant__ RasterParams p) {
    const uint32_t prim = blockIdx.x * 128 + threadIdx.x;
    if (prim >= p.totalTris) return;
    const RasterDraw& d = drawOfPrimitive(p.draws, p.drawCount, prim);
    vec4 clip[3];
    triangleClipPositions(p, d, prim - d.firstPrimitive, clip);
    ScreenPoly sp;
    clipAndSnap(p, clip, sp);
    int y0 = 0x7fffffff, y1 = -1, x0 = 0x7fffffff, x1 = -1;
    for (int k = 1; k + 1 < sp.n; k++) {
        SubTriangle s;
        if (!subTriangle(p, sp, k, s)) continue;
        y0 = imin(y0, s.iy0); y1 = imax(y1, s.iy1);
        x0 = imin(x0, s.ix0); x1 = imax(x1, s.ix1);
    }
    uint32_t info = 0xffffffffu;
    if (y1 >= y0 && (long long)(y1 - y0 + 1) * (x1 - x0 + 1) <= RASTER_TINY_MAX_AREA) {
        // tiny: this thread walks the few pixels itself (no second set-up by a warp); nothing is left for the coverage kernels
        const TriPlanes tp = trianglePlanes(clip);
        const AlphaTest alpha = alphaTestOf(d, prim - d.firstPrimitive);
        const unsigned long long keyLow = (unsigned long long)(prim + 1u);
        for (int k = 1; k + 1 < sp.n; k++) {
            SubTriangle s;
            if (!subTriangle(p, sp, k, s)) continue;
            for (int iy = s.iy0; iy <= s.iy1; iy++) {
                const long long py = (long long)iy * 256 + 128;
                const double rowDepth = tp.depth.y * pixelNdc(iy, p.H) + tp.depth.z;
                long long base[3], step[3];
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    step[e] = s.ey[e] * 256;
                    base[e] = s.ex[e] * (py - s.ay[e]) + s.bias[e] - s.ey[e] * (128 - s.ax[e]);
                }
                for (int ix = s.ix0; ix <= s.ix1; ix++) {
                    bool inside = true;
#pragma unroll
                    for (int e = 0; e < 3; e++) inside = inside && (base[e] - step[e] * (long long)ix >= 0);
                    if (inside) emitFragment<false, ALPHA>(p, tp, alpha, ix, iy, rowDepth, keyLow);
                }
            }
        }
    } else if (y1 >= y0) {
        info = (uint32_t)y0 | ((uint32_t)y1 << 16);
        const int rows = y1 - y0 + 1;
        if (rows > RASTER_SMALL_MAX_ROWS || (long long)rows * (x1 - x0 + 1) > RASTER_SMALL_MAX_AREA) {
            info |= RASTER_BIG_FLAG;
            p.triInfo[p.totalTris + 1 + atomicAdd(&p.triInfo[p.totalTris], 1u)] = prim;
        }
    }
    p.triInfo[prim] = info;
}
// coverage of rows [rowBegin, rowEnd] of one triangle by one warp
template <bool ALPHA>
__device__ void coverTriangle(const RasterParams& p, uint32_t prim, int rowBegin, int rowEnd) {
    const int lane = threadIdx.x & 31;
    const RasterDraw& d = drawOfPrimitive(p.draws, p.drawCount, prim);
    vec4 clip[3];
    triangleClipPositions(p, d, prim - d.firstPrimitive, clip);
    const TriPlanes tp = trianglePlanes(clip);
    const AlphaTest alpha = alphaTestOf(d, prim - d.firstPrimitive);
    ScreenPoly sp;
    clipAndSnap(p, clip, sp);
    const unsigned long long keyLow = (unsigned long long)(prim + 1u);
    for (int k = 1; k + 1 < sp.n; k++) {
        SubTriangle s;
        if (!subTriangle(p, sp, k, s)) continue;
        const int ya = imax(s.iy0, rowBegin), yb = imin(s.iy1, rowEnd);
        if (ya > yb) continue;
        // edge functions in pixels: E_e(ix, iy) = base_e(iy) - step_e * ix, base_e(iy + 1) = base_e(iy) + 256 * ex_e (exact, 64-bit integers)
        long long base[3], step[3], baseStep[3];
        double q0[3], slope[3];  // wide triangles: the real-valued crossing base_e(iy) / step_e = q0_e + slope_e * (iy - ya)
        const bool wide = s.ix1 - s.ix0 >= 96;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            step[e] = s.ey[e] * 256;
            baseStep[e] = s.ex[e] * 256;
            base[e] = s.ex[e] * ((long long)ya * 256 + 128 - s.ay[e]) + s.bias[e] - s.ey[e] * (128 - s.ax[e]);
            q0[e] = 0.0; slope[e] = 0.0;
            if (wide && step[e] != 0) { q0[e] = (double)base[e] / (double)step[e]; slope[e] = (double)baseStep[e] / (double)step[e]; }
        }
        for (int iy = ya; iy <= yb; iy++) {
            const double ny = pixelNdc(iy, p.H);
            long long rowBase[3];
#pragma unroll
            for (int e = 0; e < 3; e++) { rowBase[e] = base[e]; base[e] += baseStep[e]; }
            int xa = s.ix0, xe = s.ix1;
            if (wide) {
                // the span of the row from the real-valued inequality step * ix <= base with one pixel of slack; it only narrows the
                // loop, the exact test below decides
                double lo = (double)xa, hi = (double)xe;
                bool rowEmpty = false;
                const double dy = (double)(iy - ya);
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    const double q = q0[e] + slope[e] * dy;
                    if (step[e] > 0) hi = fmin(hi, q + 1.0);
                    else if (step[e] < 0) lo = fmax(lo, q - 1.0);
                    else if (rowBase[e] < 0) rowEmpty = true;
                }
                if (rowEmpty || !(lo <= hi)) continue;
                xa = imax(xa, (int)floor(lo)); xe = imin(xe, (int)ceil(hi));
            }
            const double rowDepth = tp.depth.y * ny + tp.depth.z;
            for (int xb = xa & ~31; xb <= xe; xb += 32) {
                const int ix = xb + lane;
                if (ix < xa || ix > xe) continue;
                bool inside = true;
#pragma unroll
                for (int e = 0; e < 3; e++) inside = inside && (rowBase[e] - step[e] * (long long)ix >= 0);
                if (inside) emitFragment<true, ALPHA>(p, tp, alpha, ix, iy, rowDepth, keyLow);
            }
        }
    }
}
// (2) BIG = false: a warp per triangle, the ones that fit a band; (3) BIG = true: persistent warps over (listed triangle, band)
template <bool BIG, bool ALPHA>
__global__ void __launch_bounds__(256) rasterCoverKernel(const __grid_constant__ RasterParams p) {
    const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (!BIG) {
        if (warp >= p.totalTris) return;
        const uint32_t info = p.triInfo[warp];
        if (info == 0xffffffffu || (info & RASTER_BIG_FLAG)) return;
        coverTriangle<ALPHA>(p, warp, (int)(info & 0xffffu), (int)(info >> 16));
    } else {
        const uint32_t bigCount = p.triInfo[p.totalTris], warps = gridDim.x * 8u;
        const uint32_t bands = ((uint32_t)p.H + RASTER_BIG_BAND_ROWS - 1) / RASTER_BIG_BAND_ROWS;
        const unsigned long long items = (unsigned long long)bigCount * bands;
        for (unsigned long long item = warp; item < items; item += warps) {
            const uint32_t prim = p.triInfo[p.totalTris + 1 + (uint32_t)(item / bands)], band = (uint32_t)(item % bands);
            const uint32_t info = p.triInfo[prim];
            const int y0 = (int)(info & 0xffffu), y1 = (int)((info & ~RASTER_BIG_FLAG) >> 16);
            const int ra = (int)band * RASTER_BIG_BAND_ROWS, rb = ra + RASTER_BIG_BAND_ROWS - 1;
            if (rb < y0 || ra > y1) continue;
            coverTriangle<ALPHA>(p, prim, imax(ra, y0), imin(rb, y1));
        }
    }
}
template <int KIND>
static void launchCoverage(LaunchCtx& c, const RasterParams& p) {
    cudaMemsetAsync(p.vis, 0, (size_t)p.W * p.H * sizeof(unsigned long long), c.stream);  // attachments are cleared (RenderPass.cpp:95-110)
    if (p.totalTris == 0) return;
    cudaMemsetAsync(p.triInfo + p.totalTris, 0, sizeof(uint32_t), c.stream);
    PLAIN_LAUNCH(c, rasterVertexKernel<KIND>, ceilDiv(p.totalVertices, 256), 256, 0, p);
    if (c.exec->rasterAnyAlphaTest) {  // a draw with a cut-out albedo texture: the variants that carry the alpha test
        PLAIN_LAUNCH(c, rasterSetupKernel<true>, ceilDiv(p.totalTris, 128), 128, 0, p);
        PLAIN_LAUNCH(c, (rasterCoverKernel<false, true>), ceilDiv(p.totalTris, 8), 256, 0, p);
        PLAIN_LAUNCH(c, (rasterCoverKernel<true, true>), (unsigned)c.smCount * 4, 256, 0, p);
    } else {
        PLAIN_LAUNCH(c, rasterSetupKernel<false>, ceilDiv(p.totalTris, 128), 128, 0, p);
        PLAIN_LAUNCH(c, (rasterCoverKernel<false, false>), ceilDiv(p.totalTris, 8), 256, 0, p);
        PLAIN_LAUNCH(c, (rasterCoverKernel<true, false>), (unsigned)c.smCount * 4, 256, 0, p);
    }
}
static bool fillRasterParams(LaunchCtx& c, RasterParams& p, const ImgView& depthTarget) {
    p.draws = c.exec->rasterDraws;
    p.drawCount = (uint32_t)c.exec->draws.size();
    p.totalTris = c.exec->rasterTotalTris;
    p.triInfo = c.exec->rasterTriInfo;
    p.vis = c.exec->rasterVis;
    p.vertexCache = c.exec->rasterVertexCache;
    p.totalVertices = c.exec->rasterTotalVertices;
    p.W = depthTarget.w; p.H = depthTarget.h;
    int y0, y1;
    c.window(p.H, y0, y1);
    p.yMin = y0; p.yMax = y1 - 1;
    p.clipNear = c.pass->clampDepth ? 0 : 1;
    p.clampDepth = c.pass->clampDepth ? 1 : 0;
    p.cullMode = (int)c.pass->cullMode;
    p.mainTransforms = nullptr; p.shadowTransforms = nullptr; p.cascades = nullptr; p.cascade = 0;
    p.shadowKind = 0; p.bindless = c.bindless; p.unorm8Table = c.tables;
    if (!p.vis || (p.totalTris && (!p.draws || !p.triInfo || !p.vertexCache))) { c.fail(c.pass->shader + ": rasteriser scratch missing (render_frame prepares it)"); return false; }
    if (p.W > 32767 || p.H > 32767) { c.fail(c.pass->shader + ": render targets beyond 32767 pixels are not supported"); return false; }
    return true;
}
__device__ __forceinline__ int toSnorm16(float v) { if (v != v) return 0; return (int)floorf_(clampf(v, -1.f, 1.f) * 32767.f + 0.5f); }

// ---------------- depthPrepass.vert:28-42 + depthPrepass.frag:27-49 ----------------
__global__ void __launch_bounds__(256) depthPrepassResolveKernel(const __grid_constant__ RasterParams p, ImgView motionT, ImgView normalT, ImgView depthT, const plain_global_shader_info* __restrict__ g) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.yMin + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.W || iy > p.yMax) return;
    const unsigned long long key = p.vis[(size_t)iy * p.W + ix];
    float depth = 0.f;
    uint32_t motion = 0u, normal = 0u;
    if (key) {
        depth = __uint_as_float((uint32_t)(key >> 32));
        const uint32_t primitive = (uint32_t)(key & 0xffffffffu) - 1u;
        const RasterDraw& d = drawOfPrimitive(p.draws, p.drawCount, primitive);
        vec4 passPos[3], passPosPrevious[3];
        vec3 N[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4* e = cacheEntry(p, d, primitive - d.firstPrimitive, k);
            const float4 a = e[0], b = e[1], n = e[2];
            passPos[k] = v4(a.x, a.y, a.z, a.w);
            passPosPrevious[k] = v4(b.x, b.y, b.z, b.w);
            N[k] = v3(n.x, n.y, n.z);
        }
        float l[3];
        barycentrics(trianglePlanes(passPos), pixelNdc(ix, p.W), pixelNdc(iy, p.H), l);
        const vec3 pos = v3(lerp3(l, passPos[0].x, passPos[1].x, passPos[2].x), lerp3(l, passPos[0].y, passPos[1].y, passPos[2].y), lerp3(l, passPos[0].w, passPos[1].w, passPos[2].w));
        const vec3 posPrev = v3(lerp3(l, passPosPrevious[0].x, passPosPrevious[1].x, passPosPrevious[2].x), lerp3(l, passPosPrevious[0].y, passPosPrevious[1].y, passPosPrevious[2].y),
                                lerp3(l, passPosPrevious[0].w, passPosPrevious[1].w, passPosPrevious[2].w));
        vec2 ndcCurrent = v2(pos.x, pos.y) / pos.z;
        vec2 ndcPrevious = v2(posPrev.x, posPrev.y) / posPrev.z;
        ndcCurrent = ndcCurrent + v2(g->currentFrameCameraJitter[0], g->currentFrameCameraJitter[1]);
        ndcPrevious = ndcPrevious + v2(g->previousFrameCameraJitter[0], g->previousFrameCameraJitter[1]);
        const vec2 mv = (ndcPrevious - ndcCurrent) * v2(0.5f, 0.5f);
        motion = ((uint32_t)toSnorm16(mv.x) & 0xffffu) | ((uint32_t)toSnorm16(mv.y) << 16);
        const vec3 nOut = normalize(lerp3(l, N[0], N[1], N[2])) * 0.5f + 0.5f;  // :48, the geometric normal overwrites the normal-mapped one
        normal = floatToUnorm8(nOut.x) | (floatToUnorm8(nOut.y) << 8) | (floatToUnorm8(nOut.z) << 16);
    }
    const size_t i = (size_t)iy * p.W + ix;
    ((float*)depthT.ptr)[i] = depth;
    ((uint32_t*)motionT.ptr)[i] = motion;
    ((uint32_t*)normalT.ptr)[i] = normal;
}
PLAIN_PASS(launch_depthPrepass, "depthPrepass.vert+depthPrepass.frag") {
    const ImgView motionT = c.target(0, PLAIN_FORMAT_RG16_SNORM), normalT = c.target(1, PLAIN_FORMAT_RGBA8), depthT = c.target(2, PLAIN_FORMAT_DEPTH32);
    if (c.failed) return;
    if (motionT.w != depthT.w || motionT.h != depthT.h || normalT.w != depthT.w || normalT.h != depthT.h) { c.fail("depthPrepass: attachments of different extent"); return; }
    RasterParams p;
    if (!fillRasterParams(c, p, depthT)) return;
    p.mainTransforms = c.sbuf<MainPassMatrices>(0);
    if (c.failed) return;
    launchCoverage<0>(c, p);
    if (p.yMax < p.yMin) return;
    PLAIN_LAUNCH(c, depthPrepassResolveKernel, dim3(ceilDiv(p.W, 32), ceilDiv((unsigned)(p.yMax - p.yMin + 1), 8)), 256, 0, p, motionT, normalT, depthT, c.g);
}

// ---------------- sunShadow.vert:29-32 + sunShadow.frag ----------------
__global__ void __launch_bounds__(256) shadowResolveKernel(const unsigned long long* __restrict__ vis, ImgView shadowMap) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= shadowMap.w || iy >= shadowMap.h) return;
    const unsigned long long key = vis[(size_t)iy * shadowMap.w + ix];
    ((uint16_t*)shadowMap.ptr)[(size_t)iy * shadowMap.w + ix] = key ? (uint16_t)(uint32_t)(__uint_as_float((uint32_t)(key >> 32)) * 65535.f + 0.5f) : (uint16_t)0;
}
PLAIN_PASS(launch_sunShadow, "sunShadow.vert+sunShadow.frag") {
    const ImgView shadowMap = c.target(0, PLAIN_FORMAT_DEPTH16);
    if (c.failed) return;
    RasterParams p;
    if (!fillRasterParams(c, p, shadowMap)) return;
    p.cascades = c.sbuf<plain_shadow_cascade_info>(0);
    p.shadowTransforms = c.sbuf<float>(1);
    p.cascade = c.spec<uint32_t>(0, 0);
    p.shadowKind = 1;
    if (c.failed) return;
    if (p.cascade > 3) { c.fail("sunShadow.vert: cascade index must be 0..3"); return; }
    launchCoverage<2>(c, p);
    PLAIN_LAUNCH(c, shadowResolveKernel, dim3(ceilDiv(p.W, 32), ceilDiv(p.H, 8)), 256, 0, p.vis, shadowMap);
}

// ---------------- triangle.vert:29-40 + the fetches of triangle.frag:178-193 -> packed G-buffer ----------------
__device__ __forceinline__ uint32_t octEncodeSnorm16(vec3 n) {  // inverse of the decode in gbufferShading (include/plain_frame_types.h)
    const float l1 = absf(n.x) + absf(n.y) + absf(n.z);
    float x = n.x / l1, y = n.y / l1;
    if (n.z < 0.f) {
        const float ox = (1.f - absf(y)) * (x >= 0.f ? 1.f : -1.f), oy = (1.f - absf(x)) * (y >= 0.f ? 1.f : -1.f);
        x = ox; y = oy;
    }
    return ((uint32_t)toSnorm16(x) & 0xffffu) | ((uint32_t)toSnorm16(y) << 16);
}
// unorm8 = the exact table float(b) / 255 over the 256 byte values (pass_common.cuh ShadingTables): same bits as the division
__device__ __forceinline__ vec4 sampleRGBA8LinearRepeat(const ImgView& t, vec2 uv, const float* __restrict__ unorm8Table) {
    return sampleLinear2D<WRAP_REPEAT, vec4>([&](int x, int y) {
        const uint32_t v = ldg((const uint32_t*)t.ptr + texelIndex(t, x, y));
        return v4(ldg(unorm8Table + (v & 0xffu)), ldg(unorm8Table + ((v >> 8) & 0xffu)), ldg(unorm8Table + ((v >> 16) & 0xffu)), ldg(unorm8Table + (v >> 24)));
    }, t.w, t.h, uv, v4(0.f));
}
__global__ void __launch_bounds__(256) gbufferFillResolveKernel(const __grid_constant__ RasterParams p, ImgView gbuffer, const BindlessEntry* __restrict__ bindless, const float* __restrict__ unorm8Table) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.yMin + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.W || iy > p.yMax) return;
    const unsigned long long key = p.vis[(size_t)iy * p.W + ix];
    uint4 texel = make_uint4(0u, 0u, 0u, 0u);
    if (key) {
        const uint32_t primitive = (uint32_t)(key & 0xffffffffu) - 1u;
        const RasterDraw& d = drawOfPrimitive(p.draws, p.drawCount, primitive);
        vec4 clip[3];
        vec2 uv[3];
        vec3 T[3], B[3], N[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4* e = cacheEntry(p, d, primitive - d.firstPrimitive, k);
            const float4 a = e[0], b = e[1], c2 = e[2], c3 = e[3];
            clip[k] = v4(a.x, a.y, a.z, a.w);
            T[k] = v3(b.x, b.y, b.z);
            B[k] = v3(b.w, c2.x, c2.y);
            N[k] = v3(c2.z, c2.w, c3.x);
            uv[k] = v2(c3.y, c3.z);
        }
        float l[3];
        barycentrics(trianglePlanes(clip), pixelNdc(ix, p.W), pixelNdc(iy, p.H), l);
        const vec2 passUV = v2(lerp3(l, uv[0].x, uv[1].x, uv[2].x), lerp3(l, uv[0].y, uv[1].y, uv[2].y));
        const vec3 tbnT = lerp3(l, T[0], T[1], T[2]), tbnB = lerp3(l, B[0], B[1], B[2]), tbnN = lerp3(l, N[0], N[1], N[2]);
        const vec4 albedoTexel = sampleRGBA8LinearRepeat(bindless[d.push[0]].view, passUV, unorm8Table);
        const vec4 normalTexel = sampleRGBA8LinearRepeat(bindless[d.push[1]].view, passUV, unorm8Table);
        const vec4 specularTexel = sampleRGBA8LinearRepeat(bindless[d.push[2]].view, passUV, unorm8Table);
        vec3 nrm = v3(normalTexel.x, normalTexel.y, sqrtf_(1.f - normalTexel.x * normalTexel.x + normalTexel.y + normalTexel.y));  // triangle.frag:181, as written
        nrm = nrm * 2.f - 1.f;
        vec3 Nw = normalize(tbnT * nrm.x + tbnB * nrm.y + tbnN * nrm.z);  // passTBN * normalTexelReconstructed
        if (anynan(Nw)) Nw = tbnN;                                        // :190-192
        texel.x = (uint32_t)(key >> 32);
        texel.y = octEncodeSnorm16(Nw);
        texel.z = floatToUnorm8(albedoTexel.x) | (floatToUnorm8(albedoTexel.y) << 8) | (floatToUnorm8(albedoTexel.z) << 16) | (floatToUnorm8(specularTexel.y) << 24);
        texel.w = floatToUnorm8(specularTexel.z);
    }
    ((uint4*)gbuffer.ptr)[(size_t)iy * p.W + ix] = texel;
}
PLAIN_PASS(launch_gbufferFill, "triangle.vert+gbufferFill.frag") {
    const ImgView gbuffer = c.target(0, PLAIN_FORMAT_RGBA32_UINT), depthT = c.target(1, PLAIN_FORMAT_DEPTH32);
    if (c.failed) return;
    if (gbuffer.w != depthT.w || gbuffer.h != depthT.h) { c.fail("gbufferFill: attachments of different extent"); return; }
    if (c.pass->depthFunction != PLAIN_DEPTH_EQUAL) { c.fail("gbufferFill: depth test EQUAL against the prepass expected (RenderFrontend.cpp:1555)"); return; }
    RasterParams p;
    if (!fillRasterParams(c, p, depthT)) return;
    p.mainTransforms = c.sbuf<MainPassMatrices>(17);
    if (c.failed) return;
    for (auto& d : c.exec->draws)  // material textures: bindless slot == image handle index, RGBA8
        for (int k = 0; k < 3; k++) {
            uint32_t index;
            memcpy(&index, d.push + 4 * k, 4);
            if (index >= c.be_imageCount() || c.be_imageFormat(index) != PLAIN_FORMAT_RGBA8) { c.fail("gbufferFill: material textures must be RGBA8 images (bindless index = image handle index)"); return; }
        }
    // no coverage pass: the visibility buffer of the prepass over the same draws decides (depth test EQUAL)
    if (p.totalVertices) PLAIN_LAUNCH(c, rasterVertexKernel<1>, ceilDiv(p.totalVertices, 256), 256, 0, p);
    if (p.yMax < p.yMin) return;
    PLAIN_LAUNCH(c, gbufferFillResolveKernel, dim3(ceilDiv(p.W, 32), ceilDiv((unsigned)(p.yMax - p.yMin + 1), 8)), 256, 0, p, gbuffer, c.bindless, c.tables);
}

}  // namespace pb
