// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// sdf_bake.cpp - CPU restatement of the reference's SDF bake of one mesh, Plain/src/AssetPipeline/SceneSDF.cpp:296-514
// (computeSDF) with computePointTrianglesClosestDistance :61-98. The per-mesh preparation (triangle normals, uniform grid,
// ray directions, padded volume) is plainrenderer_b200/host/SdfBakeCommon.h, shared with the CUDA bake like detmath.h.
// PINNED: tests/test_sdf_bake.py checks this function against bricks written by the reference binary itself
// (tests/golden/sdf/*.dds from oracle/_ref/PlainAssetPipeline) - bit-exact.
#include <limits>
#include <string>
#include <thread>
#include "SdfBakeCommon.h"
#include "plain_assets.h"

using namespace sdfbake;

namespace {

// SceneSDF.cpp:61-98: squared distance to the closest triangle (edge or face region), square root at the end
float closestTriangleDistance(V3 p, const std::vector<Triangle>& triangles) {
    float closest = std::numeric_limits<float>::infinity();
    for (const Triangle& t : triangles) {
        const V3 p0 = p - t.v0, p1 = p - t.v1, p2 = p - t.v2;
        const V3 e0 = t.v1 - t.v0, e1 = t.v2 - t.v1, e2 = t.v0 - t.v2;
        const V3 n0 = cross(e0, t.N), n1 = cross(e1, t.N), n2 = cross(e2, t.N);
        auto sign = [](float v) { return (float)((0.f < v) ? 1 : 0) - (float)((v < 0.f) ? 1 : 0); };
        auto neg = [](V3 v) { return v3(-v.x, -v.y, -v.z); };
        const float s0 = sign(dot(n0, neg(p0))), s1 = sign(dot(n1, neg(p1))), s2 = sign(dot(n2, neg(p2)));
        const bool onEdge = s0 + s1 + s2 < 2.f;
        auto clamp01 = [](float v) { return minf(maxf(v, 0.f), 1.f); };
        const float c0 = clamp01(dot(p0, e0) / dot(e0, e0)), c1 = clamp01(dot(p1, e1) / dot(e1, e1)), c2 = clamp01(dot(p2, e2) / dot(e2, e2));
        const V3 d0 = p - (t.v0 + e0 * c0), d1 = p - (t.v1 + e1 * c1), d2 = p - (t.v2 + e2 * c2);
        const float l0 = dot(d0, d0), l1 = dot(d1, d1), l2 = dot(d2, d2);
        float d = onEdge ? minf(minf(l0, l1), l2) : std::fabs(dot(t.N, p0) * dot(t.N, p0));
        d = std::fabs(d);
        closest = minf(closest, d);
    }
    return std::sqrt(std::fabs(closest));
}

uint16_t bakeTexel(const Prepared& P, int x, int y, int z, const uint32_t extent[3]) {
    const V3 origin = cellCenter(x, y, z, (int)extent[0], (int)extent[1], (int)extent[2], P.extends, P.offset);
    float closestTotal = std::numeric_limits<float>::infinity();
    uint32_t backHits = 0;
    for (int ray = 0; ray < kRayCount; ray++) {
        const V3 dir = P.rayDirection[ray];
        float rayClosest = std::numeric_limits<float>::infinity();
        bool backface = false;
        int cell[3];
        pointToCell(origin, P.bbMin, P.bbMax, kGridRes, cell);
        uint32_t ci[3] = {(uint32_t)cell[0], (uint32_t)cell[1], (uint32_t)cell[2]};  // glm::uvec3: stepping below 0 wraps and ends the walk
        V3 pos = origin;
        bool inside = true;
        while (inside) {
            const size_t cellIndex = (size_t)flatten((int)ci[0], (int)ci[1], (int)ci[2], kGridRes, kGridRes);
            const V3 cellMin = P.bbMin + v3((float)ci[0], (float)ci[1], (float)ci[2]) / v3((float)kGridRes, (float)kGridRes, (float)kGridRes) * P.extends;
            const V3 cellMax = cellMin + P.cellSize;
            bool hitInCell = false;
            for (uint32_t k = P.cellStart[cellIndex]; k < P.cellStart[cellIndex + 1]; k++) {
                const Triangle& t = P.triangles[P.cellTriangles[k]];
                const float NoR = dot(t.N, dir);
                if (std::fabs(NoR) < 0.0001f) continue;
                const float D = dot(t.N, t.v0);
                const float tt = (D - dot(t.N, origin)) / NoR;
                if (tt < 0.f) continue;
                const V3 e0 = t.v1 - t.v0, e1 = t.v2 - t.v1, e2 = t.v0 - t.v2;
                const V3 q = origin + dir * tt;
                const float d0 = dot(t.N, cross(q - t.v0, e0)), d1 = dot(t.N, cross(q - t.v1, e1)), d2 = dot(t.N, cross(q - t.v2, e2));
                if (!(d0 >= 0.f && d1 >= 0.f && d2 >= 0.f)) continue;
                const V3 hit = origin + dir * tt;  // origin + t * dir: the same products
                if (!(hit.x <= cellMax.x && hit.x >= cellMin.x && hit.y <= cellMax.y && hit.y >= cellMin.y && hit.z <= cellMax.z && hit.z >= cellMin.z)) continue;
                hitInCell = true;
                if (tt < rayClosest) {
                    rayClosest = tt;
                    backface = dot(dir, t.N) > 0.f;
                }
            }
            if (hitInCell) break;
            // next cell boundary along the ray (SceneSDF.cpp:441-486)
            float step = std::numeric_limits<float>::infinity();
            int axis = 0;
            const float d[3] = {dir.x, dir.y, dir.z}, p[3] = {pos.x, pos.y, pos.z}, mn[3] = {cellMin.x, cellMin.y, cellMin.z}, mx[3] = {cellMax.x, cellMax.y, cellMax.z},
                        cs[3] = {P.cellSize.x, P.cellSize.y, P.cellSize.z};
            for (int c = 0; c < 3; c++) {
                if (d[c] == 0.f) continue;
                float next;
                if (d[c] > 0) { next = mx[c]; next = next == p[c] ? next + cs[c] : next; }
                else { next = mn[c]; next = next == p[c] ? next - cs[c] : next; }
                const float dist = (next - p[c]) / d[c];
                if (dist < step) { step = dist; axis = c; }
            }
            pos = pos + dir * step;  // currentRayPosition += distance * rayDirection
            ci[axis] += d[axis] > 0 ? 1u : 0xffffffffu;
            inside = ci[axis] < (uint32_t)kGridRes;
        }
        if (backface) backHits++;
        closestTotal = minf(closestTotal, rayClosest);
    }
    const float backShare = backHits / (float)kRayCount;
    closestTotal *= backShare > 0.5f ? -1 : 1;
    if (closestTotal == std::numeric_limits<float>::infinity()) closestTotal = closestTriangleDistance(origin, P.triangles);
    return packHalf(closestTotal);
}

thread_local std::string g_error;

}  // namespace

extern "C" {

// the loaders' last_error lives in host/PlainAssets.cpp; the oracle bake has no failure modes beyond bad arguments
int PLAIN_ASSET(sdf_bake)(int, const float* positions, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, const float bbMin[3], const float bbMax[3],
                          const uint32_t extent[3], uint16_t* out, float* outKernelMs) {
    if (!positions || !indices || !out || !extent[0] || !extent[1] || !extent[2]) return 1;
    const Prepared P = prepare(positions, vertexCount, indices, indexCount, bbMin, bbMax);
    const unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < threads; w++)
        pool.emplace_back([&, w]() {
            for (uint32_t z = w; z < extent[2]; z += threads)
                for (uint32_t y = 0; y < extent[1]; y++)
                    for (uint32_t x = 0; x < extent[0]; x++)
                        out[(size_t)x + (size_t)y * extent[0] + (size_t)z * extent[0] * extent[1]] = bakeTexel(P, (int)x, (int)y, (int)z, extent);
        });
    for (auto& t : pool) t.join();
    if (outKernelMs) *outKernelMs = 0.f;
    return 0;
}

}  // extern "C"
