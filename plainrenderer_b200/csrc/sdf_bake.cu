// sdf_bake.cu - the reference's SDF bake of one mesh as a CUDA kernel.
//   Plain/src/AssetPipeline/SceneSDF.cpp:296-514 computeSDF, :61-98 computePointTrianglesClosestDistance
// One block per texel, one thread per ray (15 x 15 = 225 of 256 threads): every ray walks the 16^3 uniform grid of triangle
// lists from the texel centre (host/SdfBakeCommon.h prepares triangles, grid and ray directions exactly as the reference
// does), the block reduces the closest hit and the back-face count, a texel no ray of which hits anything takes the distance
// to the closest triangle (block-parallel minimum). The arithmetic is the reference's operation sequence in binary32 with
// IEEE division and square root and no contraction (-fmad=false): bricks equal the reference binary's byte for byte.
// The reference needs 25 s per 64^3 mesh on one core; the per-ray loop is short, divergent, and reads a few KB of triangles
// that stay in L1/L2 - the kernel is bound by instruction issue, not HBM.
#include <cuda_runtime.h>
#include <string>
#include "SdfBakeCommon.h"
#include "plain_assets.h"

namespace {

using sdfbake::kGridRes;
using sdfbake::kRayCount;

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 operator/(F3 a, F3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator+(F3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
__device__ __forceinline__ F3 operator-(F3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross3(F3 a, F3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }  // glm::min
__device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }  // glm::max
__device__ __forceinline__ float comp(F3 v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

struct BakeParams {
    const float4* triangles;       // 4 x float4 per triangle: v0, v1, v2, N (w unused)
    const uint32_t* cellStart;     // kGridRes^3 + 1
    const uint32_t* cellTriangles;
    const float4* rayDirection;    // kRayCount
    uint32_t triangleCount;
    F3 bbMin, bbMax, extends, offset, cellSize;
    uint32_t ex, ey, ez;
    uint16_t* out;
};

__device__ __forceinline__ uint16_t packHalfDevice(float f) {  // SdfBakeCommon.h packHalf (glm::packHalf semantics)
    const uint32_t u = __float_as_uint(f);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int e = (int)((u >> 23) & 0xffu) - 112;
    uint32_t m = u & 0x7fffffu;
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        m = (m | 0x800000u) >> (1 - e);
        m += 0x1000u;
        return (uint16_t)(sign | (m >> 13));
    }
    if (e == 0xff - 112) return (uint16_t)(m == 0 ? (sign | 0x7c00u) : (sign | 0x7c00u | (m >> 13) | ((m >> 13) == 0 ? 1u : 0u)));
    const uint32_t r = (((uint32_t)e << 23) | m) + 0x1000u;
    if ((r >> 23) > 30u) return (uint16_t)(sign | 0x7c00u);
    return (uint16_t)(sign | (r >> 13));
}

__device__ __forceinline__ void loadTriangle(const float4* t, uint32_t i, F3& v0, F3& v1, F3& v2, F3& N) {
    const float4 a = __ldg(t + 4 * (size_t)i), b = __ldg(t + 4 * (size_t)i + 1), c = __ldg(t + 4 * (size_t)i + 2), n = __ldg(t + 4 * (size_t)i + 3);
    v0 = f3(a.x, a.y, a.z); v1 = f3(b.x, b.y, b.z); v2 = f3(c.x, c.y, c.z); N = f3(n.x, n.y, n.z);
}

__global__ void __launch_bounds__(256) sdfBakeKernel(const __grid_constant__ BakeParams p) {
    __shared__ float sClosest[8];
    __shared__ uint32_t sBack[8];
    const uint32_t texel = blockIdx.x;
    const int x = (int)(texel % p.ex), y = (int)((texel / p.ex) % p.ey), z = (int)(texel / (p.ex * p.ey));
    const float inf = __uint_as_float(0x7f800000u);
    // volumeIndexToCellCenter :249-254
    const F3 n = (f3((float)x, (float)y, (float)z) + 0.5f) / f3((float)p.ex, (float)p.ey, (float)p.ez);
    const F3 origin = (n - 0.5f) * p.extends + p.offset;
    float rayClosest = inf;
    bool backface = false;
    const int ray = (int)threadIdx.x;
    if (ray < kRayCount) {
        const float4 d4 = __ldg(p.rayDirection + ray);
        const F3 dir = f3(d4.x, d4.y, d4.z);
        // pointToCellIndex :240-247
        F3 q = (origin - p.bbMin) / (p.bbMax - p.bbMin);
        q = f3(gmin(gmax(q.x, 0.f), 0.999f), gmin(gmax(q.y, 0.f), 0.999f), gmin(gmax(q.z, 0.f), 0.999f));
        const F3 cellF = q * f3((float)kGridRes, (float)kGridRes, (float)kGridRes);
        uint32_t ci[3] = {(uint32_t)(int)floorf(cellF.x), (uint32_t)(int)floorf(cellF.y), (uint32_t)(int)floorf(cellF.z)};
        F3 pos = origin;
        bool inside = true;
        while (inside) {
            const uint32_t cellIndex = ci[0] + ci[1] * kGridRes + ci[2] * kGridRes * kGridRes;
            const F3 cellMin = p.bbMin + f3((float)ci[0], (float)ci[1], (float)ci[2]) / f3((float)kGridRes, (float)kGridRes, (float)kGridRes) * p.extends;
            const F3 cellMax = cellMin + p.cellSize;
            bool hitInCell = false;
            const uint32_t k1 = __ldg(p.cellStart + cellIndex + 1);
            for (uint32_t k = __ldg(p.cellStart + cellIndex); k < k1; k++) {
                F3 v0, v1, v2, N;
                loadTriangle(p.triangles, __ldg(p.cellTriangles + k), v0, v1, v2, N);
                const float NoR = dot3(N, dir);
                if (fabsf(NoR) < 0.0001f) continue;
                const float D = dot3(N, v0);
                const float t = (D - dot3(N, origin)) / NoR;
                if (t < 0.f) continue;
                const F3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
                const F3 hit = origin + dir * t;
                const float d0 = dot3(N, cross3(hit - v0, e0)), d1 = dot3(N, cross3(hit - v1, e1)), d2 = dot3(N, cross3(hit - v2, e2));
                if (!(d0 >= 0.f && d1 >= 0.f && d2 >= 0.f)) continue;
                if (!(hit.x <= cellMax.x && hit.x >= cellMin.x && hit.y <= cellMax.y && hit.y >= cellMin.y && hit.z <= cellMax.z && hit.z >= cellMin.z)) continue;
                hitInCell = true;
                if (t < rayClosest) {
                    rayClosest = t;
                    backface = dot3(dir, N) > 0.f;
                }
            }
            if (hitInCell) break;
            float step = inf;
            int axis = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float dc = comp(dir, c);
                if (dc == 0.f) continue;
                const float pc = comp(pos, c), cs = comp(p.cellSize, c);
                float next;
                if (dc > 0) { next = comp(cellMax, c); next = next == pc ? next + cs : next; }
                else { next = comp(cellMin, c); next = next == pc ? next - cs : next; }
                const float dist = (next - pc) / dc;
                if (dist < step) { step = dist; axis = c; }
            }
            pos = pos + dir * step;
            const uint32_t delta = comp(dir, axis) > 0 ? 1u : 0xffffffffu;
            if (axis == 0) ci[0] += delta; else if (axis == 1) ci[1] += delta; else ci[2] += delta;
            inside = (axis == 0 ? ci[0] : (axis == 1 ? ci[1] : ci[2])) < (uint32_t)kGridRes;
        }
    }
    // block reduction: minimum of the ray distances (order-independent: no NaN can win a '<'), number of back-face rays
    float m = rayClosest;
    uint32_t back = backface ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = gmin(m, __shfl_xor_sync(0xffffffffu, m, o));
        back += __shfl_xor_sync(0xffffffffu, back, o);
    }
    if ((threadIdx.x & 31) == 0) { sClosest[threadIdx.x >> 5] = m; sBack[threadIdx.x >> 5] = back; }
    __syncthreads();
    float closestTotal = sClosest[0];
    uint32_t backHits = sBack[0];
    for (int w = 1; w < 8; w++) { closestTotal = gmin(closestTotal, sClosest[w]); backHits += sBack[w]; }
    const float backShare = (float)backHits / (float)kRayCount;
    closestTotal *= backShare > 0.5f ? -1.f : 1.f;
    if (closestTotal == inf) {  // uniform per block: no ray hit anything - distance to the closest triangle (:61-98)
        float closest = inf;
        for (uint32_t i = threadIdx.x; i < p.triangleCount; i += blockDim.x) {
            F3 v0, v1, v2, N;
            loadTriangle(p.triangles, i, v0, v1, v2, N);
            const F3 p0 = origin - v0, p1 = origin - v1, p2 = origin - v2;
            const F3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
            const F3 n0 = cross3(e0, N), n1 = cross3(e1, N), n2 = cross3(e2, N);
            auto sign = [](float v) { return (float)((0.f < v) ? 1 : 0) - (float)((v < 0.f) ? 1 : 0); };
            auto neg = [](F3 v) { return f3(-v.x, -v.y, -v.z); };
            const float s0 = sign(dot3(n0, neg(p0))), s1 = sign(dot3(n1, neg(p1))), s2 = sign(dot3(n2, neg(p2)));
            const bool onEdge = s0 + s1 + s2 < 2.f;
            const float c0 = gmin(gmax(dot3(p0, e0) / dot3(e0, e0), 0.f), 1.f), c1 = gmin(gmax(dot3(p1, e1) / dot3(e1, e1), 0.f), 1.f),
                        c2 = gmin(gmax(dot3(p2, e2) / dot3(e2, e2), 0.f), 1.f);
            const F3 q0 = origin - (v0 + e0 * c0), q1 = origin - (v1 + e1 * c1), q2 = origin - (v2 + e2 * c2);
            const float l0 = dot3(q0, q0), l1 = dot3(q1, q1), l2 = dot3(q2, q2);
            float d = onEdge ? gmin(gmin(l0, l1), l2) : fabsf(dot3(N, p0) * dot3(N, p0));
            d = fabsf(d);
            closest = gmin(closest, d);
        }
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) closest = gmin(closest, __shfl_xor_sync(0xffffffffu, closest, o));
        if ((threadIdx.x & 31) == 0) sClosest[threadIdx.x >> 5] = closest;
        __syncthreads();
        closest = sClosest[0];
        for (int w = 1; w < 8; w++) closest = gmin(closest, sClosest[w]);
        closestTotal = __fsqrt_rn(fabsf(closest));
    }
    if (threadIdx.x == 0) p.out[texel] = packHalfDevice(closestTotal);
}

thread_local std::string g_bakeError;

}  // namespace

extern "C" int PLAIN_ASSET(sdf_bake)(int device, const float* positions, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, const float bbMin[3],
                                     const float bbMax[3], const uint32_t extent[3], uint16_t* out, float* outKernelMs) {
    if (!positions || !indices || !out || !extent[0] || !extent[1] || !extent[2]) return 1;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return 1;  // no CUDA device: no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    const sdfbake::Prepared P = sdfbake::prepare(positions, vertexCount, indices, indexCount, bbMin, bbMax);
    std::vector<float4> tri(P.triangles.size() * 4), rays(kRayCount);
    for (size_t i = 0; i < P.triangles.size(); i++) {
        const sdfbake::Triangle& t = P.triangles[i];
        tri[4 * i] = make_float4(t.v0.x, t.v0.y, t.v0.z, 0.f); tri[4 * i + 1] = make_float4(t.v1.x, t.v1.y, t.v1.z, 0.f);
        tri[4 * i + 2] = make_float4(t.v2.x, t.v2.y, t.v2.z, 0.f); tri[4 * i + 3] = make_float4(t.N.x, t.N.y, t.N.z, 0.f);
    }
    for (int r = 0; r < kRayCount; r++) rays[r] = make_float4(P.rayDirection[r].x, P.rayDirection[r].y, P.rayDirection[r].z, 0.f);
    const size_t texels = (size_t)extent[0] * extent[1] * extent[2];
    float4 *dTri = nullptr, *dRays = nullptr;
    uint32_t *dStart = nullptr, *dList = nullptr;
    uint16_t* dOut = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc(&dTri, std::max<size_t>(tri.size(), 1) * sizeof(float4)) != cudaSuccess || cudaMalloc(&dRays, rays.size() * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&dStart, P.cellStart.size() * 4) != cudaSuccess || cudaMalloc(&dList, std::max<size_t>(P.cellTriangles.size(), 1) * 4) != cudaSuccess ||
            cudaMalloc(&dOut, texels * 2) != cudaSuccess) break;
        cudaMemcpy(dTri, tri.data(), tri.size() * sizeof(float4), cudaMemcpyHostToDevice);
        cudaMemcpy(dRays, rays.data(), rays.size() * sizeof(float4), cudaMemcpyHostToDevice);
        cudaMemcpy(dStart, P.cellStart.data(), P.cellStart.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dList, P.cellTriangles.data(), P.cellTriangles.size() * 4, cudaMemcpyHostToDevice);
        BakeParams p;
        p.triangles = dTri; p.cellStart = dStart; p.cellTriangles = dList; p.rayDirection = dRays;
        p.triangleCount = (uint32_t)P.triangles.size();
        auto cv = [](sdfbake::V3 v) { F3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r; };
        p.bbMin = cv(P.bbMin); p.bbMax = cv(P.bbMax); p.extends = cv(P.extends); p.offset = cv(P.offset); p.cellSize = cv(P.cellSize);
        p.ex = extent[0]; p.ey = extent[1]; p.ez = extent[2];
        p.out = dOut;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);
        sdfBakeKernel<<<(unsigned)texels, 256>>>(p);
        cudaEventRecord(e1, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) break;
        if (outKernelMs) cudaEventElapsedTime(outKernelMs, e0, e1);
        if (cudaMemcpy(out, dOut, texels * 2, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(dTri); cudaFree(dRays); cudaFree(dStart); cudaFree(dList); cudaFree(dOut);
    return rc;
}
