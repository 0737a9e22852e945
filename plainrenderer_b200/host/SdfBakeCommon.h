// SdfBakeCommon.h - host-side preparation of the SDF bake, shared by the CUDA bake (csrc/sdf_bake.cu) and
// the CPU oracle (oracle/sdf_bake.cpp): everything that does not depend on the texel.
//   * triangle list with face normals                        SceneSDF.cpp:262-277, 309-322
//   * 16^3 uniform grid of triangle lists (13-axis SAT)       SceneSDF.cpp:160-293
//   * the 15 x 15 ray directions (the same for every texel)   SceneSDF.cpp:352-367, MathUtils.cpp:4-15
//   * padded volume, brick resolution rule, half packing      sdfUtilities.cpp:5-19, VolumeInfo.cpp:4-9, SceneSDF.cpp:117-131
// The reference is compiled host code (g++ -O2, x86-64, no contraction) on glm's scalar types; every expression below is
// written with the same operand order as glm evaluates it (dot = (x*x' + y*y') + z*z', cross as glm::cross, min/max as
// glm::min/max, normalize = v * (1 / sqrt(dot))), so the prepared data - and with it the bake - reproduces the reference
// binary's bricks bit for bit (tests/golden/sdf). sin/cos/acos come from the same libm the reference binary links.
// Compile without implicit contraction (-ffp-contract=off), like the rest of the host code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace sdfbake {

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator+(V3 a, float s) { return v3(a.x + s, a.y + s, a.z + s); }
inline V3 operator-(V3 a, float s) { return v3(a.x - s, a.y - s, a.z - s); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float minf(float a, float b) { return (b < a) ? b : a; }  // glm::min
inline float maxf(float a, float b) { return (a < b) ? b : a; }  // glm::max
inline V3 vmin(V3 a, V3 b) { return v3(minf(a.x, b.x), minf(a.y, b.y), minf(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return v3(maxf(a.x, b.x), maxf(a.y, b.y), maxf(a.z, b.z)); }
inline V3 vabs(V3 a) { return v3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline V3 normalize(V3 a) { const float inv = 1.f / std::sqrt(dot(a, a)); return a * inv; }

struct Triangle { V3 v0, v1, v2, N; };

static const int kGridRes = 16;       // SceneSDF.cpp:302
static const int kRaysPerAxis = 15;   // SceneSDF.cpp:349
static const int kRayCount = kRaysPerAxis * kRaysPerAxis;

struct Prepared {
    V3 bbMin, bbMax;     // padded bounding box
    V3 extends, offset;  // VolumeInfo of the padded box
    V3 cellSize;         // extends / 16
    std::vector<Triangle> triangles;       // mesh order
    std::vector<uint32_t> cellStart;       // kGridRes^3 + 1 offsets into cellTriangles
    std::vector<uint32_t> cellTriangles;   // triangle indices, per cell in mesh order
    V3 rayDirection[kRayCount];            // index = sampleIndexX * 15 + sampleIndexY
};

// brick resolution: 4 texels per metre, next power of two, clamped to [16, 64] (SceneSDF.cpp:117-131)
inline uint32_t nextPowerOfTwo(uint32_t v) {
    if (v == 0) return 0;  // the reference's bit trick maps 0 to 0 (0 - 1 wraps, all bits set, + 1)
    uint32_t p = 1;
    while (p < v && p < 0x80000000u) p <<= 1;
    return p;
}
inline void brickResolution(const float bbMin[3], const float bbMax[3], uint32_t out[3]) {
    for (int c = 0; c < 3; c++) {
        const float targetRes = (bbMax[c] - bbMin[c]) / 0.25f;
        uint32_t r = nextPowerOfTwo((uint32_t)targetRes);
        r = r < 16u ? 16u : (r > 64u ? 64u : r);
        out[c] = r;
    }
}

// float -> half as glm::packHalf does it (glm/detail/type_half.inl toFloat16): round to nearest with ties away from zero
// ("round 0.5 up" on the magnitude), results below 2^-25 flush to zero, overflow to infinity
inline uint16_t packHalf(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int e = (int)((u >> 23) & 0xffu) - 112;
    uint32_t m = u & 0x7fffffu;
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        m = (m | 0x800000u) >> (1 - e);
        m += 0x1000u;  // half an ulp of the target, then truncate
        return (uint16_t)(sign | (m >> 13));
    }
    if (e == 0xff - 112) return (uint16_t)(m == 0 ? (sign | 0x7c00u) : (sign | 0x7c00u | (m >> 13) | ((m >> 13) == 0 ? 1u : 0u)));
    const uint32_t r = (((uint32_t)e << 23) | m) + 0x1000u;  // a carry out of the mantissa bumps the exponent
    if ((r >> 23) > 30u) return (uint16_t)(sign | 0x7c00u);
    return (uint16_t)(sign | (r >> 13));
}

inline int flatten(int x, int y, int z, int rx, int ry) { return x + y * rx + z * rx * ry; }

// pointToCellIndex, SceneSDF.cpp:240-247
inline void pointToCell(V3 p, V3 bbMin, V3 bbMax, int res, int out[3]) {
    const V3 rel = p - bbMin;
    V3 n = rel / (bbMax - bbMin);
    n = vmin(vmax(n, v3(0.f, 0.f, 0.f)), v3(0.999f, 0.999f, 0.999f));
    const V3 t = n * v3((float)res, (float)res, (float)res);
    out[0] = (int)std::floor(t.x); out[1] = (int)std::floor(t.y); out[2] = (int)std::floor(t.z);
}
// volumeIndexToCellCenter, SceneSDF.cpp:249-254
inline V3 cellCenter(int x, int y, int z, int rx, int ry, int rz, V3 extends, V3 offset) {
    const V3 n = (v3((float)x, (float)y, (float)z) + 0.5f) / v3((float)rx, (float)ry, (float)rz);
    return (n - 0.5f) * extends + offset;
}

inline bool axisSeparates(V3 axis, V3 half, V3 a, V3 b, V3 c) {  // SceneSDF.cpp:160-174
    const float p0 = dot(axis, a), p1 = dot(axis, b), p2 = dot(axis, c);
    const float r = dot(vabs(axis), half);
    const float lo = minf(minf(p0, p1), p2), hi = maxf(maxf(p0, p1), p2);
    return lo > r || hi < -r;
}
// 13-axis separating-axis test of a triangle against a box (centre, full extents), SceneSDF.cpp:179-232
inline bool triangleOverlapsBox(V3 centre, V3 extents, const Triangle& t) {
    const V3 a = t.v0 - centre, b = t.v1 - centre, c = t.v2 - centre;
    const V3 edges[3] = {b - a, c - b, a - c};
    const V3 half = extents * 0.5f;
    const V3 boxAxes[3] = {v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
    for (int e = 0; e < 3; e++)
        for (int k = 0; k < 3; k++)
            if (axisSeparates(cross(boxAxes[k], edges[e]), half, a, b, c)) return false;
    for (int k = 0; k < 3; k++)
        if (axisSeparates(boxAxes[k], half, a, b, c)) return false;
    return !axisSeparates(t.N, half, a, b, c);
}

inline Prepared prepare(const float* positions, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, const float bbMinIn[3], const float bbMaxIn[3]) {
    Prepared P;
    // padSDFBoundingBox (sdfUtilities.cpp:5-19): 7.5 % of the extent, at least 0.5 m
    const V3 lo = v3(bbMinIn[0], bbMinIn[1], bbMinIn[2]), hi = v3(bbMaxIn[0], bbMaxIn[1], bbMaxIn[2]);
    V3 padding = (hi - lo) * 0.075f;  // 0.075f * (max - min): commutative
    padding = vmax(padding, v3(0.5f, 0.5f, 0.5f));
    P.bbMin = lo - padding;
    P.bbMax = hi + padding;
    P.offset = (P.bbMax + P.bbMin) * 0.5f;  // volumeInfoFromBoundingBox
    P.extends = P.bbMax - P.bbMin;
    P.cellSize = P.extends / v3((float)kGridRes, (float)kGridRes, (float)kGridRes);

    const uint32_t triangleCount = indexCount / 3;
    P.triangles.reserve(triangleCount);
    for (uint32_t i = 0; i + 2 < indexCount; i += 3) {
        Triangle t;
        const uint32_t i0 = indices[i], i1 = indices[i + 1], i2 = indices[i + 2];
        if (i0 >= vertexCount || i1 >= vertexCount || i2 >= vertexCount) continue;
        t.v0 = v3(positions[3 * i0], positions[3 * i0 + 1], positions[3 * i0 + 2]);
        t.v1 = v3(positions[3 * i1], positions[3 * i1 + 1], positions[3 * i1 + 2]);
        t.v2 = v3(positions[3 * i2], positions[3 * i2 + 1], positions[3 * i2 + 2]);
        t.N = normalize(cross(t.v0 - t.v2, t.v0 - t.v1));
        P.triangles.push_back(t);
    }
    // uniform grid: a triangle goes into every cell of its bounding range that it overlaps; lists keep the mesh order
    const int cells = kGridRes * kGridRes * kGridRes;
    std::vector<std::vector<uint32_t>> lists((size_t)cells);
    for (uint32_t ti = 0; ti < (uint32_t)P.triangles.size(); ti++) {
        const Triangle& t = P.triangles[ti];
        int c0[3], c1[3];
        pointToCell(vmin(vmin(t.v0, t.v1), t.v2), P.bbMin, P.bbMax, kGridRes, c0);
        pointToCell(vmax(vmax(t.v0, t.v1), t.v2), P.bbMin, P.bbMax, kGridRes, c1);
        for (int x = c0[0]; x <= c1[0]; x++)
            for (int y = c0[1]; y <= c1[1]; y++)
                for (int z = c0[2]; z <= c1[2]; z++) {
                    if (x < 0 || y < 0 || z < 0 || x >= kGridRes || y >= kGridRes || z >= kGridRes) continue;
                    const V3 centre = cellCenter(x, y, z, kGridRes, kGridRes, kGridRes, P.extends, P.offset);
                    if (triangleOverlapsBox(centre, P.cellSize, t)) lists[(size_t)flatten(x, y, z, kGridRes, kGridRes)].push_back(ti);
                }
    }
    P.cellStart.resize((size_t)cells + 1);
    for (int c = 0; c < cells; c++) {
        P.cellStart[(size_t)c] = (uint32_t)P.cellTriangles.size();
        P.cellTriangles.insert(P.cellTriangles.end(), lists[(size_t)c].begin(), lists[(size_t)c].end());
    }
    P.cellStart[(size_t)cells] = (uint32_t)P.cellTriangles.size();
    // ray directions (SceneSDF.cpp:352-367): angles in degrees through directionToVector (MathUtils.cpp:4-15)
    for (int sx = 0; sx < kRaysPerAxis; sx++)
        for (int sy = 0; sy < kRaysPerAxis; sy++) {
            const float sampleX = sx / float(kRaysPerAxis - 1);
            const float sampleY = sy / float(kRaysPerAxis - 1) * 2 - 1;
            const float phi = sampleX * 2.f * 3.1415f;
            const float theta = std::acos(sampleY);
            const float degPhi = phi / 3.1415f * 180.f, degTheta = theta / 3.1415f * 180.f;
            const float radTheta = degTheta * 0.01745329251994329576923690768489f, radPhi = degPhi * 0.01745329251994329576923690768489f;  // glm::radians
            P.rayDirection[sx * kRaysPerAxis + sy] = v3(std::sin(radTheta) * std::cos(radPhi), -std::cos(radTheta), std::sin(radTheta) * std::sin(radPhi));
        }
    return P;
}

}  // namespace sdfbake
