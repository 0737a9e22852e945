// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_raster.cpp - the rasterisation passes that feed the frame path, as a scalar software rasteriser:
//   depthPrepass.vert + depthPrepass.frag   depth (D32F reverse z), motion (RG16_SNORM), geometric normal (RGBA8)
//   sunShadow.vert + sunShadow.frag         one D16 shadow cascade
//   triangle.vert + gbufferFill.frag        the packed G-buffer: the interpolated inputs and material texels of
//                                           triangle.frag:178-193 for the fragment that won the prepass (depth test EQUAL)
// What the reference leaves to the Vulkan rasteriser is defined here (and identically in csrc/passes_raster.cu):
//   * clip-space triangles are clipped (Sutherland-Hodgman, binary64) against w >= 1e-6, the near plane z <= w when depth
//     clamp is off, and a guard band of +-16384 pixels around the viewport centre; the polygon is fanned from its first vertex
//   * window coordinates are snapped to 1/256 pixel, coverage = exact integer edge functions at pixel centres, top-left rule
//   * front face = counter clockwise = negative sum of x_i*y_j - x_j*y_i in window coordinates (VulkanPipeline.cpp:61)
//   * depth and the perspective-correct barycentrics come from the plane equations of the UNCLIPPED triangle in homogeneous
//     coordinates (2-D homogeneous rasterisation), evaluated in binary64 at the pixel centre: with v_i = (x_i, y_i, w_i),
//     c_0 = v_1 x v_2 (cyclic), e_i = c_i . (px, py, 1): lambda_i = e_i * (1 / (e_0 + e_1 + e_2)); depth is the affine function
//     (sum z_i c_i) / det . (px, py, 1) with det = (sum w_i c_i).z (its x and y components vanish analytically), rounded to
//     binary32; fragments outside [0, 1] are dropped (clamped when depth clamp is on)
//   * depth test GREATER_EQUAL in draw order = per pixel the maximum of (depth bits, primitive number): of two fragments at the
//     same depth the one drawn later wins; draws are ordered by draw_meshes call, triangles by index-buffer position
//   * texture fetches (the material texels of gbufferFill, the alpha test of depthPrepass.frag:28-31 / sunShadow.frag:19-22) are
//     bilinear, repeat, mip 0 of RGBA8 images (the reference samples BC-compressed mip chains anisotropically with a mip bias);
//     the alpha test interpolates uv with the pixel's perspective-correct barycentrics and drops the fragment before the depth
//     test when alpha < 0.5; it is skipped for albedo textures without a texel of alpha < 255 (it could not drop anything)
#include <cmath>
#include "backend.h"
#include "shader_inc.h"
#include "shading_hook.h"

namespace orc {

PrepassFragmentHook g_prepassFragmentHook = {nullptr, nullptr};

struct D3 { double x, y, z; };
static D3 crossd(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct ClipV { double x, y, z, w; };

// ---- vertex fetch: VertexInput.h:27-31, VulkanVertexInput.cpp:4-10 ----
struct VertexIn { vec3 pos; vec2 uv; vec3 normal, tangent, bitangent; };
static float snorm10(uint32_t bits) {
    int v = (int)(bits & 1023u);
    if (v >= 512) v -= 1024;
    return max((float)v / 511.f, -1.f);
}
static vec3 unpackA2R10G10B10Snorm(uint32_t p) { return vec3(snorm10(p >> 20), snorm10(p >> 10), snorm10(p)); }
static uint32_t fetchIndex(const Mesh& m, uint32_t k) { return m.index32 ? ((const uint32_t*)m.indices.data())[k] : ((const uint16_t*)m.indices.data())[k]; }
static VertexIn fetchVertex(const Mesh& m, uint32_t index) {
    const uint8_t* p = m.vertices.data() + (size_t)index * 28;
    VertexIn v;
    float f[3]; memcpy(f, p, 12);
    uint16_t h[2]; memcpy(h, p + 12, 4);
    uint32_t n[3]; memcpy(n, p + 16, 12);
    v.pos = vec3(f[0], f[1], f[2]);
    v.uv = vec2(halfToFloat(h[0]), halfToFloat(h[1]));
    v.normal = unpackA2R10G10B10Snorm(n[0]);
    v.tangent = unpackA2R10G10B10Snorm(n[1]);
    v.bitangent = unpackA2R10G10B10Snorm(n[2]);
    return v;
}
static vec3 mulMat3(const mat4& m, vec3 v) { return vfma(m.c[2].xyz(), v.z, vfma(m.c[1].xyz(), v.y, m.c[0].xyz() * v.x)); }  // mat3(model) * v, the contract's M * v

// ---- homogeneous plane equations of one (unclipped) triangle ----
struct TriPlanes { D3 c0, c1, c2, depth; };
static TriPlanes trianglePlanes(const vec4 clip[3]) {
    D3 v[3];
    double z[3], w[3];
    for (int i = 0; i < 3; i++) { v[i] = {(double)clip[i].x, (double)clip[i].y, (double)clip[i].w}; z[i] = (double)clip[i].z; w[i] = (double)clip[i].w; }
    TriPlanes t;
    t.c0 = crossd(v[1], v[2]); t.c1 = crossd(v[2], v[0]); t.c2 = crossd(v[0], v[1]);
    const D3 num = {t.c0.x * z[0] + t.c1.x * z[1] + t.c2.x * z[2], t.c0.y * z[0] + t.c1.y * z[1] + t.c2.y * z[2], t.c0.z * z[0] + t.c1.z * z[1] + t.c2.z * z[2]};
    const double det = t.c0.z * w[0] + t.c1.z * w[1] + t.c2.z * w[2];
    t.depth = {num.x / det, num.y / det, num.z / det};
    return t;
}
static double pixelNdc(int i, int size) { return ((double)i + 0.5) * (2.0 / (double)size) - 1.0; }  // NDC of a pixel centre
static void barycentrics(const TriPlanes& t, double px, double py, float l[3]) {
    const double e0 = t.c0.x * px + t.c0.y * py + t.c0.z, e1 = t.c1.x * px + t.c1.y * py + t.c1.z, e2 = t.c2.x * px + t.c2.y * py + t.c2.z;
    const double inv = 1.0 / (e0 + e1 + e2);
    l[0] = (float)(e0 * inv); l[1] = (float)(e1 * inv); l[2] = (float)(e2 * inv);
}
static float lerp3(const float l[3], float a, float b, float c) { return fma_(l[2], c, fma_(l[1], b, l[0] * a)); }
static vec3 lerp3(const float l[3], vec3 a, vec3 b, vec3 c) { return vec3(lerp3(l, a.x, b.x, c.x), lerp3(l, a.y, b.y, c.y), lerp3(l, a.z, b.z, c.z)); }

// ---- coverage ----
struct RasterTarget { int W, H; bool clipNear, clampDepth; uint32_t cullMode; uint64_t* vis; };
struct AlphaTest { bool enabled; View albedo; vec2 uv[3]; };
static void rasterTriangle(const RasterTarget& rt, const vec4 clip[3], uint32_t primitive, const AlphaTest& alphaTest) {
    const TriPlanes tp = trianglePlanes(clip);
    ClipV poly[12], tmp[12];
    int n = 3;
    for (int i = 0; i < 3; i++) poly[i] = {(double)clip[i].x, (double)clip[i].y, (double)clip[i].z, (double)clip[i].w};
    const double gx = 32768.0 / (double)rt.W, gy = 32768.0 / (double)rt.H;
    auto dist = [&](int plane, const ClipV& p) -> double {
        switch (plane) {
            case 0: return p.w - 1e-6;
            case 1: return rt.clipNear ? p.w - p.z : 1.0;
            case 2: return gx * p.w - p.x;
            case 3: return gx * p.w + p.x;
            case 4: return gy * p.w - p.y;
            default: return gy * p.w + p.y;
        }
    };
    for (int plane = 0; plane < 6 && n >= 3; plane++) {
        bool allIn = true;
        for (int i = 0; i < n; i++) allIn = allIn && dist(plane, poly[i]) >= 0.0;
        if (allIn) continue;
        int m = 0;
        for (int i = 0; i < n; i++) {
            const ClipV a = poly[i], b = poly[(i + 1) % n];
            const double da = dist(plane, a), db = dist(plane, b);
            if (da >= 0.0) tmp[m++] = a;
            if ((da >= 0.0) != (db >= 0.0)) {
                const double t = da / (da - db);
                tmp[m++] = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w)};
            }
        }
        n = m;
        for (int i = 0; i < n; i++) poly[i] = tmp[i];
    }
    if (n < 3) return;
    int64_t X[12], Y[12];
    for (int i = 0; i < n; i++) {
        const double inv = 1.0 / poly[i].w;
        const double xs = (poly[i].x * inv * 0.5 + 0.5) * (double)rt.W, ys = (poly[i].y * inv * 0.5 + 0.5) * (double)rt.H;
        X[i] = (int64_t)std::floor(xs * 256.0 + 0.5);
        Y[i] = (int64_t)std::floor(ys * 256.0 + 0.5);
    }
    for (int k = 1; k + 1 < n; k++) {
        int64_t x0 = X[0], y0 = Y[0], x1 = X[k], y1 = Y[k], x2 = X[k + 1], y2 = Y[k + 1];
        const int64_t area2 = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
        if (area2 == 0) continue;
        const bool front = area2 < 0;
        if ((rt.cullMode == PLAIN_CULL_BACK && !front) || (rt.cullMode == PLAIN_CULL_FRONT && front)) continue;
        if (area2 < 0) { std::swap(x1, x2); std::swap(y1, y2); }
        const int64_t ex[3] = {x1 - x0, x2 - x1, x0 - x2}, ey[3] = {y1 - y0, y2 - y1, y0 - y2};
        const int64_t ax[3] = {x0, x1, x2}, ay[3] = {y0, y1, y2};
        int64_t bias[3];
        for (int e = 0; e < 3; e++) bias[e] = (ey[e] < 0 || (ey[e] == 0 && ex[e] > 0)) ? 0 : -1;  // top-left rule
        const int64_t minX = std::min(x0, std::min(x1, x2)), maxX = std::max(x0, std::max(x1, x2)), minY = std::min(y0, std::min(y1, y2)), maxY = std::max(y0, std::max(y1, y2));
        const int64_t ix0 = std::max<int64_t>(0, (minX - 128 + 255) >> 8), ix1 = std::min<int64_t>(rt.W - 1, (maxX - 128) >> 8);
        const int64_t iy0 = std::max<int64_t>(0, (minY - 128 + 255) >> 8), iy1 = std::min<int64_t>(rt.H - 1, (maxY - 128) >> 8);
        for (int64_t iy = iy0; iy <= iy1; iy++)
            for (int64_t ix = ix0; ix <= ix1; ix++) {
                const int64_t px = ix * 256 + 128, py = iy * 256 + 128;
                bool inside = true;
                for (int e = 0; e < 3; e++) inside = inside && (ex[e] * (py - ay[e]) - ey[e] * (px - ax[e]) + bias[e] >= 0);
                if (!inside) continue;
                const double nx = pixelNdc((int)ix, rt.W), ny = pixelNdc((int)iy, rt.H);
                if (alphaTest.enabled) {
                    float l[3];
                    barycentrics(tp, nx, ny, l);
                    const vec2 passUV(lerp3(l, alphaTest.uv[0].x, alphaTest.uv[1].x, alphaTest.uv[2].x), lerp3(l, alphaTest.uv[0].y, alphaTest.uv[1].y, alphaTest.uv[2].y));
                    if (texture(alphaTest.albedo, s_linearRepeat, passUV).w < 0.5f) continue;  // discard
                }
                float d = (float)(tp.depth.x * nx + (tp.depth.y * ny + tp.depth.z));  // the row term first (it is constant along a row)
                if (d != d) continue;
                if (rt.clampDepth) d = d < 0.f ? 0.f : (d > 1.f ? 1.f : d);
                else if (d < 0.f || d > 1.f) continue;
                const uint64_t key = ((uint64_t)(dm::f2u(d) & 0x7fffffffu) << 32) | (uint64_t)(primitive + 1u);
                uint64_t& slot = rt.vis[(size_t)iy * rt.W + ix];
                if (key > slot) slot = key;
            }
    }
}

struct DrawInfo { const Mesh* mesh; uint32_t firstPrimitive; const uint8_t* push; };
static std::vector<DrawInfo> drawTable(const PassCtx& c) {
    std::vector<DrawInfo> t;
    uint32_t first = 0;
    for (auto& d : c.exec->draws) {
        const Mesh* m = &c.ctx->meshes[d.mesh];
        t.push_back({m, first, d.push});
        first += m->indexCount / 3;
    }
    return t;
}
static const DrawInfo& drawOfPrimitive(const std::vector<DrawInfo>& t, uint32_t primitive) {
    size_t lo = 0, hi = t.size();
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (t[mid].firstPrimitive <= primitive) lo = mid; else hi = mid; }
    return t[lo];
}
static uint32_t pushU32(const DrawInfo& d, int i) { uint32_t v; memcpy(&v, d.push + i * 4, 4); return v; }
static bool hasTransparentTexel(const View& albedo) {  // RGBA8 whose initial data (create_image) had a texel of alpha < 255 in mip 0
    return albedo.valid() && albedo.format() == PLAIN_FORMAT_RGBA8 && albedo.img->transparentTexels;
}
template <typename MatrixOfDraw>
static void rasterDraws(const PassCtx& c, const std::vector<DrawInfo>& draws, const RasterTarget& rt, MatrixOfDraw matrixOfDraw) {
    for (auto& d : draws) {
        const mat4 M = matrixOfDraw(d);
        AlphaTest at;
        at.albedo = c.bindless(pushU32(d, 0));  // push constant 0 of both vertex programs
        at.enabled = hasTransparentTexel(at.albedo);
        for (uint32_t t = 0; t < d.mesh->indexCount / 3; t++) {
            vec4 clip[3];
            for (int k = 0; k < 3; k++) {
                const VertexIn v = fetchVertex(*d.mesh, fetchIndex(*d.mesh, t * 3 + k));
                clip[k] = M * vec4(v.pos, 1.f);
                at.uv[k] = v.uv;
            }
            rasterTriangle(rt, clip, d.firstPrimitive + t, at);
        }
    }
}
static int16_t toSnorm16(float v) { if (v != v) return 0; return (int16_t)(int)dm::floor_(clamp(v, -1.f, 1.f) * 32767.f + 0.5f); }
struct MainPassMatrices { float model[16], mvp[16], mvpPrevious[16]; };  // MainPassMatrices.inc

// ---------------- depthPrepass.vert:28-42 + depthPrepass.frag:27-49 ----------------
ORACLE_PASS(pass_depthPrepass, "depthPrepass.vert+depthPrepass.frag") {
    View motionT = c.target(0), normalT = c.target(1), depthT = c.target(2);
    if (!motionT.valid() || !normalT.valid() || !depthT.valid()) return;
    const MainPassMatrices* transforms = (const MainPassMatrices*)c.sbuf(0);
    const int W = depthT.w(), H = depthT.h();
    std::vector<uint64_t>& vis = c.ctx->visibility[c.exec->targets[2].image.index];
    vis.assign((size_t)W * H, 0);  // attachments are cleared (RenderFrontend.cpp:1718-1720, RenderPass.cpp:95-110)
    const std::vector<DrawInfo> draws = drawTable(c);
    const RasterTarget rt{W, H, !c.pass->clampDepth, c.pass->clampDepth != 0, c.pass->cullMode, vis.data()};
    rasterDraws(c, draws, rt, [&](const DrawInfo& d) { return c.gm4(transforms[pushU32(d, 3)].mvp); });
    const plain_global_shader_info& g = c.g;
    if (g_prepassFragmentHook.beginPass) g_prepassFragmentHook.beginPass(c);
    parallelFor(c.ctx->threads, H, [&](int iy) {
        for (int ix = 0; ix < W; ix++) {
            const uint64_t key = vis[(size_t)iy * W + ix];
            float depth = 0.f;
            int16_t motion[2] = {0, 0};
            uint8_t normal[4] = {0, 0, 0, 0};
            if (key) {
                depth = dm::u2f((uint32_t)(key >> 32));
                const uint32_t primitive = (uint32_t)(key & 0xffffffffu) - 1u;
                const DrawInfo& d = drawOfPrimitive(draws, primitive);
                const MainPassMatrices& tr = transforms[pushU32(d, 3)];
                const mat4 model = c.gm4(tr.model), mvp = c.gm4(tr.mvp), mvpPrevious = c.gm4(tr.mvpPrevious);
                vec4 passPos[3], passPosPrevious[3];
                vec3 N[3];
                for (int k = 0; k < 3; k++) {
                    const VertexIn v = fetchVertex(*d.mesh, fetchIndex(*d.mesh, (primitive - d.firstPrimitive) * 3 + k));
                    passPos[k] = mvp * vec4(v.pos, 1.f);
                    passPosPrevious[k] = mvpPrevious * vec4(v.pos, 1.f);
                    N[k] = normalize(mulMat3(model, v.normal));
                }
                float l[3];
                barycentrics(trianglePlanes(passPos), pixelNdc(ix, W), pixelNdc(iy, H), l);
                const vec3 pos(lerp3(l, passPos[0].x, passPos[1].x, passPos[2].x), lerp3(l, passPos[0].y, passPos[1].y, passPos[2].y), lerp3(l, passPos[0].w, passPos[1].w, passPos[2].w));
                const vec3 posPrev(lerp3(l, passPosPrevious[0].x, passPosPrevious[1].x, passPosPrevious[2].x), lerp3(l, passPosPrevious[0].y, passPosPrevious[1].y, passPosPrevious[2].y),
                                   lerp3(l, passPosPrevious[0].w, passPosPrevious[1].w, passPosPrevious[2].w));
                if (g_prepassFragmentHook.shade) {  // liboracle_refmain.so: the fragment stage is the reference's own depthPrepass.frag main()
                    vec2 mvRef;
                    vec3 nRef;
                    g_prepassFragmentHook.shade(vec4(pos.x, pos.y, 0.f, pos.z), vec4(posPrev.x, posPrev.y, 0.f, posPrev.z), lerp3(l, N[0], N[1], N[2]), mvRef, nRef);
                    motion[0] = toSnorm16(mvRef.x); motion[1] = toSnorm16(mvRef.y);
                    normal[0] = floatToUnorm8(nRef.x); normal[1] = floatToUnorm8(nRef.y); normal[2] = floatToUnorm8(nRef.z);
                    memcpy(depthT.texelPtr(ix, iy, 0), &depth, 4);
                    memcpy(motionT.texelPtr(ix, iy, 0), motion, 4);
                    memcpy(normalT.texelPtr(ix, iy, 0), normal, 4);
                    continue;
                }
                vec2 ndcCurrent = vec2(pos.x, pos.y) / pos.z;
                vec2 ndcPrevious = vec2(posPrev.x, posPrev.y) / posPrev.z;
                ndcCurrent += vec2(g.currentFrameCameraJitter[0], g.currentFrameCameraJitter[1]);
                ndcPrevious += vec2(g.previousFrameCameraJitter[0], g.previousFrameCameraJitter[1]);
                const vec2 mv = (ndcPrevious - ndcCurrent) * vec2(0.5f, 0.5f);
                motion[0] = toSnorm16(mv.x); motion[1] = toSnorm16(mv.y);
                const vec3 nOut = normalize(lerp3(l, N[0], N[1], N[2])) * 0.5f + 0.5f;  // :48, the geometric normal overwrites the normal-mapped one
                normal[0] = floatToUnorm8(nOut.x); normal[1] = floatToUnorm8(nOut.y); normal[2] = floatToUnorm8(nOut.z);
            }
            memcpy(depthT.texelPtr(ix, iy, 0), &depth, 4);
            memcpy(motionT.texelPtr(ix, iy, 0), motion, 4);
            memcpy(normalT.texelPtr(ix, iy, 0), normal, 4);
        }
    });
}

// ---------------- sunShadow.vert:29-32 + sunShadow.frag ----------------
ORACLE_PASS(pass_sunShadow, "sunShadow.vert+sunShadow.frag") {
    View shadowMap = c.target(0);
    if (!shadowMap.valid()) return;
    const uint32_t cascadeIndex = c.spec<uint32_t>(0, 0);
    plain_shadow_cascade_info cascades;
    memcpy(&cascades, c.sbuf(0), sizeof(cascades));
    const float* transforms = (const float*)c.sbuf(1);
    const int W = shadowMap.w(), H = shadowMap.h();
    std::vector<uint64_t>& vis = c.ctx->visibility[c.exec->targets[0].image.index];
    vis.assign((size_t)W * H, 0);
    const std::vector<DrawInfo> draws = drawTable(c);
    const RasterTarget rt{W, H, !c.pass->clampDepth, c.pass->clampDepth != 0, c.pass->cullMode, vis.data()};
    const mat4 lightMatrix = c.gm4(cascades.lightMatrices[cascadeIndex < 4 ? cascadeIndex : 0]);
    rasterDraws(c, draws, rt, [&](const DrawInfo& d) { return lightMatrix * c.gm4(transforms + (size_t)pushU32(d, 1) * 16); });
    for (int iy = 0; iy < H; iy++)
        for (int ix = 0; ix < W; ix++) {
            const uint64_t key = vis[(size_t)iy * W + ix];
            const uint16_t v = key ? (uint16_t)(dm::u2f((uint32_t)(key >> 32)) * 65535.f + 0.5f) : (uint16_t)0;
            memcpy(shadowMap.texelPtr(ix, iy, 0), &v, 2);
        }
}

// ---------------- triangle.vert:29-40 + the fetches of triangle.frag:178-193 -> packed G-buffer ----------------
static uint32_t octEncodeSnorm16(vec3 n) {  // inverse of the decode in gbufferShading (include/plain_frame_types.h)
    const float l1 = abs(n.x) + abs(n.y) + abs(n.z);
    float x = n.x / l1, y = n.y / l1;
    if (n.z < 0.f) {
        const float ox = (1.f - abs(y)) * (x >= 0.f ? 1.f : -1.f), oy = (1.f - abs(x)) * (y >= 0.f ? 1.f : -1.f);
        x = ox; y = oy;
    }
    return (uint32_t)(uint16_t)toSnorm16(x) | ((uint32_t)(uint16_t)toSnorm16(y) << 16);
}
ORACLE_PASS(pass_gbufferFill, "triangle.vert+gbufferFill.frag") {
    View gbuffer = c.target(0), depthT = c.target(1);
    if (!gbuffer.valid() || !depthT.valid()) return;
    const MainPassMatrices* transforms = (const MainPassMatrices*)c.sbuf(17);
    const int W = gbuffer.w(), H = gbuffer.h();
    auto it = c.ctx->visibility.find(c.exec->targets[1].image.index);
    const std::vector<DrawInfo> draws = drawTable(c);
    const bool haveVis = it != c.ctx->visibility.end() && it->second.size() == (size_t)W * H;  // depth test EQUAL against the prepass of the same draws
    parallelFor(c.ctx->threads, H, [&](int iy) {
        for (int ix = 0; ix < W; ix++) {
            uint32_t texel[4] = {0, 0, 0, 0};
            const uint64_t key = haveVis ? it->second[(size_t)iy * W + ix] : 0;
            if (key) {
                const uint32_t primitive = (uint32_t)(key & 0xffffffffu) - 1u;
                const DrawInfo& d = drawOfPrimitive(draws, primitive);
                const MainPassMatrices& tr = transforms[pushU32(d, 3)];
                const mat4 model = c.gm4(tr.model), mvp = c.gm4(tr.mvp);
                vec4 clip[3];
                vec2 uv[3];
                vec3 T[3], B[3], N[3];
                for (int k = 0; k < 3; k++) {
                    const VertexIn v = fetchVertex(*d.mesh, fetchIndex(*d.mesh, (primitive - d.firstPrimitive) * 3 + k));
                    clip[k] = mvp * vec4(v.pos, 1.f);
                    uv[k] = v.uv;
                    T[k] = normalize(mulMat3(model, v.tangent));
                    N[k] = normalize(mulMat3(model, v.normal));
                    B[k] = normalize(mulMat3(model, v.bitangent));
                }
                float l[3];
                barycentrics(trianglePlanes(clip), pixelNdc(ix, W), pixelNdc(iy, H), l);
                const vec2 passUV(lerp3(l, uv[0].x, uv[1].x, uv[2].x), lerp3(l, uv[0].y, uv[1].y, uv[2].y));
                const vec3 tbnT = lerp3(l, T[0], T[1], T[2]), tbnB = lerp3(l, B[0], B[1], B[2]), tbnN = lerp3(l, N[0], N[1], N[2]);
                const vec3 albedoTexel = texture(c.bindless(pushU32(d, 0)), s_linearRepeat, passUV).xyz();
                const vec2 normalTexel = texture(c.bindless(pushU32(d, 1)), s_linearRepeat, passUV).xy();
                const vec3 specularTexel = texture(c.bindless(pushU32(d, 2)), s_linearRepeat, passUV).xyz();
                vec3 nrm(normalTexel.x, normalTexel.y, sqrt(1.f - normalTexel.x * normalTexel.x + normalTexel.y + normalTexel.y));  // triangle.frag:181, as written
                nrm = nrm * 2.f - 1.f;
                vec3 Nw = normalize(tbnT * nrm.x + tbnB * nrm.y + tbnN * nrm.z);  // passTBN * normalTexelReconstructed
                if (isnan(Nw.x) || isnan(Nw.y) || isnan(Nw.z)) Nw = tbnN;       // :190-192
                texel[0] = (uint32_t)(key >> 32);
                texel[1] = octEncodeSnorm16(Nw);
                texel[2] = (uint32_t)floatToUnorm8(albedoTexel.x) | ((uint32_t)floatToUnorm8(albedoTexel.y) << 8) | ((uint32_t)floatToUnorm8(albedoTexel.z) << 16) |
                           ((uint32_t)floatToUnorm8(specularTexel.y) << 24);
                texel[3] = (uint32_t)floatToUnorm8(specularTexel.z);
            }
            memcpy(gbuffer.texelPtr(ix, iy, 0), texel, 16);
        }
    });
}

}  // namespace orc
