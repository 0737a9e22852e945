// passes_gi.cu - SDF instance culling, diffuse SDF sphere trace, spatial/temporal denoise, upscale.
//   sdfCameraFrustumCulling.comp:36-62, sdfCameraTileCulling.comp:37-99, sdfDiffuseTrace.comp:70-207 + SDF.inc:12-184,
//   filterIndirectDiffuseSpatial.comp:21-135, filterIndirectDiffuseTemporal.comp:20-86, indirectLightUpscale.comp:17-71
#include <cstdlib>
#include "shader_inc.cuh"

namespace pb {

// sdfCulling.inc:17-20: strides by the FULL screen resolution even when the trace runs at half resolution
__device__ __forceinline__ uint32_t tileIndexFromTileUV(int tileX, int tileY, const plain_global_shader_info* g) {
    const float t = (float)g->screenResolution[0] / 32.f;
    const float fl = floorf_(t);
    const uint32_t tileCountX = f2u(fl + ((fl < t) ? 1.f : 0.f));  // ceil
    return (uint32_t)tileX + (uint32_t)tileY * tileCountX;
}

// ---------------- sdfCameraFrustumCulling.comp ----------------
// The reference appends with atomicAdd (order undefined); here one warp appends in ascending instance order
// (ballot + prefix popcount), so the list - and with it the 100-instance cut of the tile lists - is deterministic.
__global__ void __launch_bounds__(32) sdfFrustumCullingKernel(const uint32_t* __restrict__ instanceBuffer, const plain_camera_frustum_buffer* __restrict__ frustum, uint32_t* culled,
                                                               size_t culledCapacity, const plain_bounding_box* __restrict__ instanceBBs, const float* __restrict__ influenceRangePtr, uint32_t invocations) {
    const uint32_t instanceCount = instanceBuffer[0];
    const float influenceRange = *influenceRangePtr;
    const uint32_t lane = threadIdx.x;
    uint32_t count = culled[0];
    const uint32_t n = min(instanceCount, invocations);
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t instanceIndex = base + lane;
        bool isInsideFrustum = false;
        if (instanceIndex < n) {
            const plain_bounding_box bb = instanceBBs[instanceIndex];
            const vec3 bbMin = ld3(bb.bbMin), bbMax = ld3(bb.bbMax);
            const vec3 boundingSphereCenter = (bbMax + bbMin) * 0.5f;
            const vec3 bbExtends = (bbMax - bbMin);
            float boundingSphereRadius = fmaxp(fmaxp(bbExtends.x, bbExtends.y), bbExtends.z) * 0.5f;
            boundingSphereRadius += influenceRange;
            isInsideFrustum = true;
            for (int i = 0; i < 6; i++) {
                const vec3 frustumPoint = ld3(frustum->frustumPoints[i]), frustumNormal = ld3(frustum->frustumNormals[i]);
                const bool isOutsidePlane = dot(boundingSphereCenter - frustumPoint, frustumNormal) > boundingSphereRadius;
                isInsideFrustum = isInsideFrustum && !isOutsidePlane;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, isInsideFrustum);
        if (isInsideFrustum) {
            const uint32_t slot = count + __popc(mask & ((1u << lane) - 1u));
            if (1 + slot < culledCapacity) culled[1 + slot] = instanceIndex;
        }
        count += __popc(mask);
    }
    __syncwarp();
    if (lane == 0) culled[0] = count;
}
PLAIN_PASS(launch_sdfFrustumCulling, "sdfCameraFrustumCulling.comp") {
    size_t culledSize = 0;
    const uint32_t* instances = c.sbuf<uint32_t>(0);
    const plain_camera_frustum_buffer* frustum = c.ubuf<plain_camera_frustum_buffer>(1);
    uint32_t* culled = c.sbuf<uint32_t>(2, &culledSize);
    const plain_bounding_box* bbs = c.sbuf<plain_bounding_box>(3);
    const float* influence = c.ubuf<float>(4);
    if (c.failed) return;
    PLAIN_LAUNCH(c, sdfFrustumCullingKernel, 1, 32, 0, instances, frustum, culled, culledSize / 4, bbs, influence, c.exec->dispatch[0] * 64);
}

// ---------------- sdfCameraTileCulling.comp ----------------
// one warp per 32x32-pixel tile; lanes test 32 instances at a time against the tile's cone, bounded by the HiZ depth
// range of the tile, and append in list order up to maxObjectsPerTile (sdfCulling.inc:5)
__global__ void __launch_bounds__(128) sdfTileCullingKernel(const uint32_t* __restrict__ culled, const plain_bounding_box* __restrict__ instanceBBs, plain_culled_instances_per_tile* cullingTiles,
                                                             size_t tileCapacity, const float* __restrict__ influenceRangePtr, ImgView depthMinMaxTexture, const plain_global_shader_info* __restrict__ g,
                                                             int useHiZ, uint32_t tileCountX, uint32_t tileCountY, uint32_t limitX, uint32_t limitY) {
    const uint32_t warpGlobal = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (warpGlobal >= tileCountX * tileCountY) return;
    const int tx = (int)(warpGlobal % tileCountX), ty = (int)(warpGlobal / tileCountX);
    if ((uint32_t)tx >= limitX || (uint32_t)ty >= limitY) return;
    const uint32_t tileIndex = tileIndexFromTileUV(tx, ty, g);
    if ((size_t)tileIndex >= tileCapacity) return;  // out-of-bounds writes are dropped
    const Globals G = loadGlobals(g);
    const float influenceRange = *influenceRangePtr;
    const vec2 res = v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    auto VFromiUV = [&](int ix, int iy) {  // :37-40, normalised by the full screen resolution as in the reference
        const vec2 pixelCoor = (v2((float)ix, (float)iy) / res - 0.5f) * 2.f;
        return viewDirFromNDC(G, pixelCoor);
    };
    const int cullingTileSize = 32;
    const vec3 cameraToPixel = -VFromiUV(tx * cullingTileSize + cullingTileSize / 2, ty * cullingTileSize + cullingTileSize / 2);
    vec3 V_ll = -VFromiUV(tx * cullingTileSize, ty * cullingTileSize);
    vec3 V_ur = -VFromiUV(tx * cullingTileSize + cullingTileSize, ty * cullingTileSize + cullingTileSize);
    V_ll = V_ll / dot(cameraToPixel, V_ll);
    V_ur = V_ur / dot(cameraToPixel, V_ur);
    const float coneRadiusPerMeter = length(V_ll - V_ur) * 0.5f;
    float depthMin = G.nearPlane, depthMax = G.farPlane;
    const vec2 uv = v2((float)tx, (float)ty) / v2((float)tileCountX, (float)tileCountY);
    if (useHiZ) {
        const vec2 depthMinMax = sampleNearest2D<WRAP_CLAMP, vec2>([&](int x, int y) { return loadRG32F(depthMinMaxTexture, x, y); }, depthMinMaxTexture.w, depthMinMaxTexture.h, uv, v2(0.f));
        depthMin = linearizeDepth(depthMinMax.y, G.nearPlane, G.farPlane);
        depthMax = linearizeDepth(depthMinMax.x, G.nearPlane, G.farPlane);
    }
    depthMin *= dot(cameraToPixel, G.fwd);
    depthMax *= dot(cameraToPixel, G.fwd);
    plain_culled_instances_per_tile& tile = cullingTiles[tileIndex];
    const uint32_t culledInstanceCount = culled[0];
    uint32_t objectCount = 0;
    for (uint32_t base = 0; base < culledInstanceCount && objectCount < PLAIN_MAX_OBJECTS_PER_TILE; base += 32) {
        const uint32_t i = base + lane;
        bool pass = false;
        uint32_t instanceIndex = 0;
        if (i < culledInstanceCount) {
            instanceIndex = culled[1 + i];
            const plain_bounding_box bb = instanceBBs[instanceIndex];
            const vec3 bbMin = ld3(bb.bbMin), bbMax = ld3(bb.bbMax);
            const vec3 boundingSphereCenter = (bbMax + bbMin) * 0.5f;
            const vec3 bbExtends = (bbMax - bbMin) * 0.5f;
            float boundingSphereRadius = fmaxp(fmaxp(bbExtends.x, bbExtends.y), bbExtends.z);
            boundingSphereRadius += influenceRange;
            float projection = dot(boundingSphereCenter - G.camPos, cameraToPixel);
            projection = clampf(projection, depthMin, depthMax);
            const float d = length(boundingSphereCenter - (projection * cameraToPixel + G.camPos));
            pass = d < boundingSphereRadius + coneRadiusPerMeter * projection;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (pass) {
            const uint32_t slot = objectCount + __popc(mask & ((1u << lane) - 1u));
            if (slot < PLAIN_MAX_OBJECTS_PER_TILE) tile.indices[slot] = instanceIndex;
        }
        objectCount = min(objectCount + (uint32_t)__popc(mask), (uint32_t)PLAIN_MAX_OBJECTS_PER_TILE);
    }
    if (lane == 0) tile.objectCount = objectCount;
}
PLAIN_PASS(launch_sdfTileCulling, "sdfCameraTileCulling.comp") {
    size_t tilesSize = 0;
    const uint32_t* culled = c.sbuf<uint32_t>(0);
    const plain_bounding_box* bbs = c.sbuf<plain_bounding_box>(1);
    plain_culled_instances_per_tile* tiles = c.sbuf<plain_culled_instances_per_tile>(2, &tilesSize);
    const float* influence = c.ubuf<float>(3);
    const ImgView hiz = c.sampled(4, PLAIN_FORMAT_RG32_SFLOAT);
    if (c.failed) return;
    const uint32_t tcx = c.push<uint32_t>(0), tcy = c.push<uint32_t>(4);
    if (tcx == 0 || tcy == 0) return;
    PLAIN_LAUNCH(c, sdfTileCullingKernel, ceilDiv(tcx * tcy, 4), 128, 0, culled, bbs, tiles, tilesSize / sizeof(plain_culled_instances_per_tile), influence, hiz, c.g,
                 c.specBool(0, false) ? 1 : 0, tcx, tcy, c.exec->dispatch[0] * 8, c.exec->dispatch[1] * 8);
}

// ---------------- sdfDiffuseTrace.comp ----------------
struct TraceInstance {  // one SDFInstance (SDF.inc:4-10) staged in shared memory together with its brick view
    float worldToLocal[16];
    vec3 localExtends;
    vec3 meanAlbedo;
    ImgView sdf;
    vec3 sphereCenter;  // world-space bounding sphere of the brick's box (conservative: radius padded), see instanceBoundingSphere
    float sphereR2;
    // per-instance values of SDF.inc:131-141, evaluated once when the instance is staged (same expressions)
    vec3 localExtendsHalfPadded;   // localExtends * 0.5 + 0.01
    float distanceThreshold;       // length(localExtends / sdfResolution) * 0.25
    float localToGlobalScale;      // 1 / length(worldToLocal[0].xyz)
    vec3 invLocalExtends;          // 1 / localExtends per component: "pos / localExtends" multiplies by it (contract 2)
};
struct TraceResult {
    bool hit;
    float closestHitDistance;
    vec3 hitPos, N;
    int hitCount;
    vec3 albedo;
    // march state of the hit that currently holds closestHitDistance: hit position, normal and albedo (SDF.inc:165-176)
    // are evaluated from it once, after all instances, instead of at every improvement (only the last one survives)
    int winner;  // index into the tile's list, -1: none
    vec3 winnerSamplePos, winnerRayDirection;
    float winnerD, winnerDLast;
};
__device__ __forceinline__ float sampleSDF(const ImgView& sdf, vec3 uv) { return sampleR16FLinearClamp3D(sdf, uv); }  // SDF.inc:12-14
__device__ __forceinline__ vec3 normalFromSDF(vec3 uv, vec3 extends, const ImgView& sdf) {  // SDF.inc:16-25
    const float extendsMax = fmaxp(extends.x, fmaxp(extends.y, extends.z));
    const vec3 extendsNormalized = extends / extendsMax;
    const vec3 epsilon = v3(0.15f) / v3((float)sdf.w, (float)sdf.h, (float)sdf.d) / extendsNormalized;
    return normalize(v3(sampleSDF(sdf, uv + v3(epsilon.x, 0, 0)) - sampleSDF(sdf, uv - v3(epsilon.x, 0, 0)),
                        sampleSDF(sdf, uv + v3(0, epsilon.y, 0)) - sampleSDF(sdf, uv - v3(0, epsilon.y, 0)),
                        sampleSDF(sdf, uv + v3(0, 0, epsilon.z)) - sampleSDF(sdf, uv - v3(0, 0, epsilon.z))));
}
__device__ __forceinline__ bool isPointInAABB(vec3 p, vec3 mn, vec3 mx) {  // SDF.inc:27-35
    return p.x >= mn.x && p.y >= mn.y && p.z >= mn.z && p.x <= mx.x && p.y <= mx.y && p.z <= mx.z;
}
__device__ __forceinline__ bool rayAABBIntersection(vec3 rayOrigin, vec3 rayDirection, vec3 aabbMin, vec3 aabbMax, float& tOut) {  // SDF.inc:42-86
    bool hit = false;
    float t = 100000.f;
    float intersection = rayOrigin.x < 0.f ? aabbMin.x : aabbMax.x;
    const float tx = (intersection - rayOrigin.x) / rayDirection.x;
    vec3 pI = rayOrigin + tx * rayDirection;
    if (tx > 0.f && pI.y >= aabbMin.y && pI.y <= aabbMax.y && pI.z >= aabbMin.z && pI.z <= aabbMax.z) { t = fminp(t, tx); hit = true; }
    intersection = rayOrigin.y < 0.f ? aabbMin.y : aabbMax.y;
    const float ty = (intersection - rayOrigin.y) / rayDirection.y;
    pI = rayOrigin + ty * rayDirection;
    if (ty > 0.f && pI.x >= aabbMin.x && pI.x <= aabbMax.x && pI.z >= aabbMin.z && pI.z <= aabbMax.z) { t = fminp(t, ty); hit = true; }
    intersection = rayOrigin.z < 0.f ? aabbMin.z : aabbMax.z;
    const float tz = (intersection - rayOrigin.z) / rayDirection.z;
    pI = rayOrigin + tz * rayDirection;
    if (tz > 0.f && pI.x >= aabbMin.x && pI.x <= aabbMax.x && pI.y >= aabbMin.y && pI.y <= aabbMax.y) { t = fminp(t, tz); hit = true; }
    tOut = t;
    return hit;
}
// SDF.inc:101-184, split at the head of the sphere-trace loop so that a warp can interleave lanes that are marching
// through different instances (see sdfDiffuseTraceKernel). The per-lane operation sequence is the reference's.
struct MarchState {
    vec3 localSamplePos, rayDirection;
    float hitDistanceLocal, d, dLast;
    int k;
};
// SDF.inc:101-141: transform the ray, clip it to the box, early-outs. Returns true when the lane has to march.
template <typename Inst>
__device__ __forceinline__ bool traceSetup(const Inst& inst, vec3 rayStartWorld, vec3 rayDirectionWorld, const TraceResult& tr, MarchState& st) {
    const vec3 localExtends = inst.localExtends;
    vec3 rayStartLocal = xyz(mulm4(inst.worldToLocal, v4(rayStartWorld, 1.f)));
    const vec3 rayEndLocal = xyz(mulm4(inst.worldToLocal, v4(rayStartWorld + rayDirectionWorld, 1.f)));
    vec3 rayDirection = rayEndLocal - rayStartLocal;
    rayDirection = rayDirection / length(rayDirection);
    const vec3 sdfMaxLocal = localExtends * 0.5f;
    const vec3 sdfMinLocal = -sdfMaxLocal;
    float hitDistanceLocal = 0.f;
    if (!isPointInAABB(rayStartLocal, sdfMinLocal, sdfMaxLocal)) {
        float t;
        if (rayAABBIntersection(rayStartLocal, rayDirection, sdfMinLocal, sdfMaxLocal, t)) {
            rayStartLocal = rayStartLocal + t * rayDirection;
            hitDistanceLocal = t;
        } else {
            return false;
        }
    }
    if (inst.localToGlobalScale * hitDistanceLocal > tr.closestHitDistance) return false;
    st.localSamplePos = rayStartLocal;
    st.rayDirection = rayDirection;
    st.hitDistanceLocal = hitDistanceLocal;
    st.d = 0.f;
    st.dLast = 0.f;
    st.k = 0;
    return true;
}
// one iteration of the loop SDF.inc:144-183. Returns true while the lane keeps marching through this instance.
// `listIndex` is the position of the instance in the tile's list. The reference visits the list in order and replaces the
// closest hit only by a strictly closer one, so of two hits at the same distance the one listed first wins; this kernel
// visits the instances in a different order (sdfDiffuseTraceKernel) and applies that rule explicitly.
template <typename Inst>
__device__ __forceinline__ bool traceStep(const Inst& inst, int listIndex, TraceResult& tr, MarchState& st) {
    if (st.k >= 128) return false;
    const vec3 localExtendsHalf = inst.localExtendsHalfPadded;
    const vec3 localSamplePos = st.localSamplePos;
    if (localSamplePos.x > localExtendsHalf.x || localSamplePos.y > localExtendsHalf.y || localSamplePos.z > localExtendsHalf.z ||
        localSamplePos.x < -localExtendsHalf.x || localSamplePos.y < -localExtendsHalf.y || localSamplePos.z < -localExtendsHalf.z)
        return false;
    const vec3 sampleUV = localSamplePos * inst.invLocalExtends + 0.5f;  // localSamplePos / localExtends + 0.5
    st.dLast = st.d;
    const float d = sampleSDF(inst.sdf, sampleUV);
    st.d = d;
    if (d < inst.distanceThreshold) {
        tr.hit = true;
        const float distanceGlobal = st.hitDistanceLocal * inst.localToGlobalScale;
        if (distanceGlobal < tr.closestHitDistance || (distanceGlobal == tr.closestHitDistance && listIndex < tr.winner)) {
            tr.closestHitDistance = distanceGlobal;
            tr.hitCount = st.k;
            tr.winner = listIndex;
            tr.winnerSamplePos = localSamplePos;
            tr.winnerRayDirection = st.rayDirection;
            tr.winnerD = d;
            tr.winnerDLast = st.dLast;
        }
        return false;
    }
    st.localSamplePos = localSamplePos + st.rayDirection * absf(d);
    st.hitDistanceLocal += absf(d);
    st.k++;
    return true;
}
// SDF.inc:165-176 for the hit that ended up closest
__device__ __forceinline__ void shadeWinner(const TraceInstance& inst, vec3 rayStartWorld, vec3 rayDirectionWorld, TraceResult& tr) {
    const float d = tr.winnerD;
    const float lastStepSizeLocal = d / (1.f - (d - tr.winnerDLast));
    const vec3 hitSamplePos = tr.winnerSamplePos + tr.winnerRayDirection * lastStepSizeLocal;
    const vec3 sampleUV = hitSamplePos * inst.invLocalExtends + 0.5f;
    const vec3 N = normalFromSDF(sampleUV, inst.localExtends, inst.sdf);
    const float* m = inst.worldToLocal;  // transpose(mat3(worldToLocal)) * N
    tr.N = v3(m[0], m[4], m[8]) * N.x + v3(m[1], m[5], m[9]) * N.y + v3(m[2], m[6], m[10]) * N.z;
    tr.albedo = vpow(inst.meanAlbedo, v3(2.2f));
    const float lastStepSizeGlobal = lastStepSizeLocal * inst.localToGlobalScale;
    tr.hitPos = rayStartWorld + rayDirectionWorld * (tr.closestHitDistance + lastStepSizeGlobal);
}

// World-space sphere around the image of the local box [-extends/2, extends/2] under inverse(worldToLocal), padded by
// 0.1 % + 1 mm. Only used to SKIP instances whose box the ray cannot touch (the reference would run its slab test and
// return at SDF.inc:115-126 with no effect), never to accept one; any NaN/inf (singular matrix) disables the skip.
template <typename Inst>
__device__ __forceinline__ void instanceBoundingSphere(Inst& t) {
    const float* m = t.worldToLocal;
    const vec3 a0 = v3(m[0], m[1], m[2]), a1 = v3(m[4], m[5], m[6]), a2 = v3(m[8], m[9], m[10]), tr = v3(m[12], m[13], m[14]);
    const vec3 c12 = cross(a1, a2), c20 = cross(a2, a0), c01 = cross(a0, a1);
    const float invDet = 1.f / dot(a0, c12);
    // rows of inverse(A) are c12, c20, c01 scaled by 1/det: x_w = invA * (x_l - tr)
    auto invA = [&](vec3 v) { return v3(dot(c12, v), dot(c20, v), dot(c01, v)) * invDet; };
    t.sphereCenter = invA(-tr);
    const vec3 h = t.localExtends * 0.5f;
    float r2 = 0.f;
    for (int k = 0; k < 4; k++) {
        const vec3 o = invA(v3(h.x, (k & 1) ? -h.y : h.y, (k & 2) ? -h.z : h.z));
        r2 = fmaxf(r2, dot(o, o));
    }
    const float r = sqrtf(r2) * 1.001f + 0.001f;
    t.sphereR2 = (r2 == r2) ? r * r : dm::nanf_();
}
template <typename Inst>
__device__ __forceinline__ bool rayMissesSphere(const Inst& t, vec3 o, vec3 L, float invLen2) {
    const vec3 oc = t.sphereCenter - o;
    const float oc2 = dot(oc, oc), tca = dot(oc, L);
    const float R2 = t.sphereR2 + 1e-5f * oc2;  // slack for the cancellation in oc2 - tca^2
    if (oc2 <= R2) return false;
    if (tca < 0.f) return true;
    return oc2 - tca * tca * invLen2 > R2;  // false when anything is NaN
}

struct TraceParams {
    ImgView outYSH, outCoCg, depthTexture, normalTexture, skyLut, shadowMap;
    const plain_light_buffer* light;
    const unsigned char* instanceBuffer;  // uvec4 header + SDFInstance[]
    const plain_culled_instances_per_tile* tiles;
    size_t tileCapacity;
    const float* influenceRange;
    const plain_shadow_cascade_info* cascades;
    const plain_global_shader_info* g;
    const BindlessEntry* bindless;
    int strictInfluenceRadiusCutoff, shadowCascadeIndex;
    int groupsX, groupsY;
    int blockRowOffset;  // row sharding: first 16-row block row of this launch
    int variant;         // scheduling experiments (PLAIN_TRACE_VARIANT): bit 0 compact the slab survivors, bit 1 slice long marches, bit 2 lazier refill
};

// ---- corner-replicated SDF bricks (BindlessEntry::corners) ----
// Entry (i, j, k), 0 <= i <= w etc., holds the eight texels a trilinear tap with lower texel (i - 1, j - 1, k - 1) blends under
// clamp-to-edge addressing, as halves in the sampler's order: x = a00 | a10 << 16, y = a01 | a11 << 16 (slice z0), z, w = slice z1.
__global__ void __launch_bounds__(256) buildCornerBrickKernel(uint4* __restrict__ corners, const uint16_t* __restrict__ texels, int w, int h, int d) {
    const int n = (w + 1) * (h + 1) * (d + 1);
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    const int i = e % (w + 1), j = (e / (w + 1)) % (h + 1), k = e / ((w + 1) * (h + 1));
    const int x0 = imax(i - 1, 0), x1 = imin(i, w - 1), y0 = imax(j - 1, 0), y1 = imin(j, h - 1), z0 = imax(k - 1, 0), z1 = imin(k, d - 1);
    auto T = [&](int x, int y, int z) { return (uint32_t)__ldg(texels + ((size_t)z * h + y) * w + x); };
    uint4 c;
    c.x = T(x0, y0, z0) | (T(x1, y0, z0) << 16);
    c.y = T(x0, y1, z0) | (T(x1, y1, z0) << 16);
    c.z = T(x0, y0, z1) | (T(x1, y0, z1) << 16);
    c.w = T(x0, y1, z1) | (T(x1, y1, z1) << 16);
    corners[e] = c;
}
void buildCornerBrick(uint4* corners, const unsigned char* texels, int w, int h, int d, cudaStream_t stream) {
    const int n = (w + 1) * (h + 1) * (d + 1);
    buildCornerBrickKernel<<<ceilDiv((unsigned)n, 256), 256, 0, stream>>>(corners, (const uint16_t*)texels, w, h, d);
}

// One SDFInstance staged for the round-2 tracer: the fields of TraceInstance plus the corner-replicated brick, the albedo of a hit
// (pow(meanAlbedo, 2.2), SDF.inc:174: it depends on the instance only) and whether the instance qualifies for the lean march.
struct TraceInst2 {
    float worldToLocal[16];
    vec3 localExtends;           float distanceThreshold;
    vec3 localExtendsHalfPadded; float localToGlobalScale;
    vec3 invLocalExtends;        float sphereR2;
    vec3 sphereCenter;           int fastOk;
    vec3 albedo;                 int strideY;   // (w + 1)
    vec3 dimsF;                  int strideZ;   // (w + 1) * (h + 1)
    const uint4* corners;        int dimX, dimY;
    ImgView sdf;                 // .d doubles as dimZ
};
// texture(sampler3D R16F, uvw) with the linear + clamp-to-edge sampler from the corner-replicated brick: the set-up and blend of
// image_view.h sampleLinear3D<WRAP_CLAMP> (same expressions, same operands), the eight texels from ONE 128-bit load. The entry
// index clamp(x0 + 1, 0, w) addresses exactly the clamped pair (clamp(x0), clamp(x0 + 1)) for every x0.
// SANITIZE = false: the caller guarantees finite coordinates with |u| <= 65536 (sanitizeCoord is the identity).
template <bool SANITIZE>
__device__ __forceinline__ float sampleCornerBrick(const TraceInst2& t, vec3 uvw) {
    const float fx = fmaf_(SANITIZE ? sanitizeCoord(uvw.x) : uvw.x, t.dimsF.x, -0.5f);
    const float fy = fmaf_(SANITIZE ? sanitizeCoord(uvw.y) : uvw.y, t.dimsF.y, -0.5f);
    const float fz = fmaf_(SANITIZE ? sanitizeCoord(uvw.z) : uvw.z, t.dimsF.z, -0.5f);
    const int ix = floor2i(fx), iy = floor2i(fy), iz = floor2i(fz);  // |f| < 2^31: == f2i(floorf_(f)), and (float)i == floorf_(f)
    const float ax = fx - (float)ix, ay = fy - (float)iy, az = fz - (float)iz;
    const float bx = 1.f - ax, by = 1.f - ay, bz = 1.f - az;
    const int ex = iclamp(ix + 1, 0, t.dimX), ey = iclamp(iy + 1, 0, t.dimY), ez = iclamp(iz + 1, 0, t.sdf.d);
    const uint4 c = __ldg(t.corners + (ez * t.strideZ + ey * t.strideY + ex));
    const float a00 = halfToFloat((uint16_t)(c.x & 0xffffu)), a10 = halfToFloat((uint16_t)(c.x >> 16)), a01 = halfToFloat((uint16_t)(c.y & 0xffffu)), a11 = halfToFloat((uint16_t)(c.y >> 16));
    const float b00 = halfToFloat((uint16_t)(c.z & 0xffffu)), b10 = halfToFloat((uint16_t)(c.z >> 16)), b01 = halfToFloat((uint16_t)(c.w & 0xffffu)), b11 = halfToFloat((uint16_t)(c.w >> 16));
    const float w00 = bx * by, w10 = ax * by, w01 = bx * ay, w11 = ax * ay;
    const float s0 = fmaf_(a11, w11, fmaf_(a01, w01, fmaf_(a10, w10, a00 * w00)));
    const float s1 = fmaf_(b11, w11, fmaf_(b01, w01, fmaf_(b10, w10, b00 * w00)));
    return fmaf_(s1, az, s0 * bz);
}
__device__ __forceinline__ bool finite3(vec3 a) { return absf(a.x) < 3.0e38f && absf(a.y) < 3.0e38f && absf(a.z) < 3.0e38f; }
// One iteration of SDF.inc:144-183 for a lean instance: traceStep with the brick tap from the corner copy and no coordinate
// sanitising. Valid while the march state is finite (checked by the caller after the set-up) and d stays finite; a non-finite
// d raises `poisoned` and leaves tr untouched, and the caller redoes this (ray, instance) with the spelled-out functions.
__device__ __forceinline__ bool traceStepLean(const TraceInst2& inst, int listIndex, TraceResult& tr, MarchState& st, bool& poisoned) {
    if (st.k >= 128) return false;
    const vec3 h = inst.localExtendsHalfPadded;
    const vec3 pos = st.localSamplePos;
    if (absf(pos.x) > h.x || absf(pos.y) > h.y || absf(pos.z) > h.z) return false;  // x > h || x < -h for a finite x
    const vec3 sampleUV = pos * inst.invLocalExtends + 0.5f;
    st.dLast = st.d;
    const float d = sampleCornerBrick<false>(inst, sampleUV);
    if (!(absf(d) < 3.0e38f)) { poisoned = true; return false; }
    st.d = d;
    if (d < inst.distanceThreshold) {
        tr.hit = true;
        const float distanceGlobal = st.hitDistanceLocal * inst.localToGlobalScale;
        if (distanceGlobal < tr.closestHitDistance || (distanceGlobal == tr.closestHitDistance && listIndex < tr.winner)) {
            tr.closestHitDistance = distanceGlobal;
            tr.hitCount = st.k;
            tr.winner = listIndex;
            tr.winnerSamplePos = pos;
            tr.winnerRayDirection = st.rayDirection;
            tr.winnerD = d;
            tr.winnerDLast = st.dLast;
        }
        return false;
    }
    st.localSamplePos = pos + st.rayDirection * absf(d);
    st.hitDistanceLocal += absf(d);
    st.k++;
    return true;
}
// SDF.inc:101-184 for one (ray, instance) with every rule spelled out (coordinate sanitising, the plain brick): the path of
// instances, rays and bricks with non-finite values. Out of line: rare.
__device__ __noinline__ void traceInstanceSpelledOutImpl(const TraceInst2* inst, int listIndex, vec3 rayOrigin, vec3 L, TraceResult* tr) {
    MarchState st;
    if (!traceSetup(*inst, rayOrigin, L, *tr, st)) return;
    while (traceStep(*inst, listIndex, *tr, st)) {}
}
// the caller's TraceResult stays in registers: only a copy has its address taken
__device__ __forceinline__ void traceInstanceSpelledOut(const TraceInst2* inst, int listIndex, vec3 rayOrigin, vec3 L, TraceResult* tr) {
    TraceResult copy = *tr;
    traceInstanceSpelledOutImpl(inst, listIndex, rayOrigin, L, &copy);
    *tr = copy;
}
// SDF.inc:165-176 for the hit that ended up closest (normal from six taps of the corner copy, sanitised like the sampler)
__device__ __forceinline__ void shadeWinner2(const TraceInst2& inst, vec3 rayStartWorld, vec3 rayDirectionWorld, TraceResult& tr) {
    const float d = tr.winnerD;
    const float lastStepSizeLocal = d / (1.f - (d - tr.winnerDLast));
    const vec3 hitSamplePos = tr.winnerSamplePos + tr.winnerRayDirection * lastStepSizeLocal;
    const vec3 uv = hitSamplePos * inst.invLocalExtends + 0.5f;
    vec3 N;
    if (inst.corners) {  // normalFromSDF, SDF.inc:16-25
        const vec3 extends = inst.localExtends;
        const float extendsMax = fmaxp(extends.x, fmaxp(extends.y, extends.z));
        const vec3 extendsNormalized = extends / extendsMax;
        const vec3 epsilon = v3(0.15f) / v3((float)inst.sdf.w, (float)inst.sdf.h, (float)inst.sdf.d) / extendsNormalized;
        N = normalize(v3(sampleCornerBrick<true>(inst, uv + v3(epsilon.x, 0, 0)) - sampleCornerBrick<true>(inst, uv - v3(epsilon.x, 0, 0)),
                         sampleCornerBrick<true>(inst, uv + v3(0, epsilon.y, 0)) - sampleCornerBrick<true>(inst, uv - v3(0, epsilon.y, 0)),
                         sampleCornerBrick<true>(inst, uv + v3(0, 0, epsilon.z)) - sampleCornerBrick<true>(inst, uv - v3(0, 0, epsilon.z))));
    } else {
        N = normalFromSDF(uv, inst.localExtends, inst.sdf);
    }
    const float* m = inst.worldToLocal;  // transpose(mat3(worldToLocal)) * N
    tr.N = v3(m[0], m[4], m[8]) * N.x + v3(m[1], m[5], m[9]) * N.y + v3(m[2], m[6], m[10]) * N.z;
    tr.albedo = inst.albedo;
    const float lastStepSizeGlobal = lastStepSizeLocal * inst.localToGlobalScale;
    tr.hitPos = rayStartWorld + rayDirectionWorld * (tr.closestHitDistance + lastStepSizeGlobal);
}

// State of a hit as its marching lane leaves it for the ray's own thread: what shadeWinner2 needs (SDF.inc:165-176)
struct HitRecord {  // 20 bytes: the local ray direction is recomputed from the ray and the instance (the expression of traceSetup: same operands, same
    vec3 samplePos; //           bits), the march count is not used by the diffuse trace - so the pool holds 416 records instead of 192 in the same space
    float d, dLast;
};
#define TRACE_PAIR_CAP 2560  // (ray, candidate) pairs of one batch of rays; a ray has at most 100
#define TRACE_HIT_POOL 416   // hit records per block (ncu, round 2: an indoor scene reports ~330 per 256 rays; with 192 the winners without a record cost 6 % of the kernel); a block that reports more re-marches those winners
#define TRACE_QUEUE_CAP 896  // marches waiting for a lane (pairs whose ray enters the box), per batch; more are marched on the spot
#define TRACE_LONG_CAP 16    // marches parked between two slices (<= 256: the compaction uses one thread per slot); slicing is off by default
struct LongMarch {           // a march parked between two slices: the whole MarchState
    vec3 pos, dir;
    float hitDistanceLocal, d, dLast;
    int k;
    uint32_t pair;
};
#define TRACE_NO_SLOT 0xffffu
struct MarchJob {            // MarchState at the box entry (SDF.inc:101-141 done) + which pair it belongs to
    vec3 pos, dir;
    float hitDistanceLocal;
    uint32_t pair;           // ray << 8 | list index
};
// Conservative ray / box rejection in the instance's local space, BEFORE the reference's own test (SDF.inc:104-126): slabs against
// the box inflated by 0.1 % + 1 mm + 1e-5 of the origin's magnitude, unnormalised direction, approximate reciprocals. It can only
// say "certainly misses": the reference's test accepts a ray that starts inside the box or whose face intersection lies on the
// closed box, up to the rounding of a few operations on these same operands - far inside the inflation. A rejected pair would have
// returned at SDF.inc:123 with no effect. (oL, dL are the operands traceSetup forms; the compiler shares them.)
__device__ __forceinline__ bool boxCertainlyMissed(const float* m, vec3 halfExtents, vec3 o, vec3 dirWorld) {
    const vec3 oL = xyz(mulm4(m, v4(o, 1.f)));
    const vec3 eL = xyz(mulm4(m, v4(o + dirWorld, 1.f)));
    const vec3 dL = eL - oL;
    const float slack = 1e-3f + 1e-5f * (absf(oL.x) + absf(oL.y) + absf(oL.z));
    const vec3 h = halfExtents * 1.001f + slack;
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(dL.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(dL.y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(dL.z));
    const float ax = (-h.x - oL.x) * ix, bx = (h.x - oL.x) * ix;
    const float ay = (-h.y - oL.y) * iy, by = (h.y - oL.y) * iy;
    const float az = (-h.z - oL.z) * iz, bz = (h.z - oL.z) * iz;
    const float tNear = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));  // fminf / fmaxf drop a NaN (0 * inf on a slab boundary)
    const float tFar = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return tFar < fmaxf(tNear, 0.f);
}

// key of a ray's closest hit: distance bits << 32 | (list index + 1) << 16 | hit record slot. The distance of a hit is a non-negative
// float, so unsigned order is numeric order; the minimum over the hits of a ray is the closest hit, a tie going to the instance
// listed first - the rule of the reference's in-order loop (traceStep). The initial key (10000, nothing) loses against every
// hit closer than 10000 and wins against one at exactly 10000, like `distanceGlobal < closestHitDistance`.
__device__ __forceinline__ unsigned long long traceKey(float distance, int listIndex, uint32_t slot) {
    return ((unsigned long long)dm::f2u(distance) << 32) | ((unsigned long long)(uint32_t)(listIndex + 1) << 16) | (unsigned long long)slot;
}

// Round-2 tracer. A block traces a 16x16-pixel region = 2x2 of the reference's 8x8 workgroups; all four lie in one 32x32
// culling tile (:154), whose instance records are staged once in shared memory. Each 8x8 group keeps its own ray cache for
// the 3x3 resolve (:70-116). Every invocation of a dispatched group traces, also those beyond the image edge.
//
// Per (ray, instance) the operation sequence is exactly the reference's (SDF.inc:101-184); across instances the result of a ray
// is the closest hit with ties going to the instance listed first - what the reference's in-order loop computes - so it does not
// depend on the order in which the instances are visited, on which lane visits them, or on whether an instance whose box lies
// behind the closest hit is visited at all (SDF.inc:141 only ever skips work that cannot win; round 1 established that). The
// work is therefore cut for the machine instead of for the pixel (ncu of the per-pixel kernel: 9 box tests and 2.7 short
// marches per ray, 9 of 32 lanes busy in a march step because a warp lives as long as its slowest ray):
//   B  every thread generates its ray and its candidate mask (bounding-sphere test over the tile's list), all lanes in lockstep
//   C  the (ray, candidate) pairs of the block are compacted into one list (block-wide prefix sum of the candidate counts)
//   D1 the box test of every pair (SDF.inc:101-141 behind a conservative slab rejection) with all 256 threads busy, whatever the
//      spread of candidates per ray; a pair whose ray enters its box leaves a march job (the local ray at the entry) in a queue
//   D2 lanes take march jobs from the queue with an atomic counter; a lane whose march ends takes the next job, so a warp stays
//      populated until the queue is empty. SDF.inc:141 is applied again when a job is taken (against the closest hit found
//      meanwhile). A hit enters the ray's key with one 64-bit atomicMin; its state goes to a hit record for the shading
//   F  every thread shades its own ray from the winning hit record (normal, albedo, shadow or sky), then the resolve
// A march step reads its eight brick texels with one 128-bit load of the corner-replicated brick (TraceInst2).
__global__ void __launch_bounds__(256, 3) sdfDiffuseTraceKernel(const __grid_constant__ TraceParams p) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    TraceInst2* sInst = (TraceInst2*)smemRaw;                                         // [PLAIN_MAX_OBJECTS_PER_TILE]
    unsigned long long* sKey = (unsigned long long*)(sInst + PLAIN_MAX_OBJECTS_PER_TILE);  // [256] closest hit of every ray
    HitRecord* sHitPool = (HitRecord*)(sKey + 256);                                   // [TRACE_HIT_POOL]
    float* sRayO = (float*)(sHitPool + TRACE_HIT_POOL);                               // [256][3]
    float* sRayL = sRayO + 256 * 3;                                                   // [256][3]
    float* sRayNormal = sRayL + 256 * 3;                                              // [4][8][8][3]
    float* sRayDepth = sRayNormal + 256 * 3;                                          // [4][8][8]
    uint32_t* sIncl = (uint32_t*)(sRayDepth + 256);                                   // [256] inclusive prefix sums of the candidate counts
    uint16_t* sPairs = (uint16_t*)(sIncl + 256);                                      // [TRACE_PAIR_CAP] ray << 8 | list index
    uint8_t* sHit = (uint8_t*)(sPairs + TRACE_PAIR_CAP);                              // [256] tr.hit: an instance reported d < threshold, closest or not
    MarchJob* sQueue = (MarchJob*)(sHit + 256);                                       // [TRACE_QUEUE_CAP] marches waiting for a lane
    LongMarch* sLong = (LongMarch*)(sQueue + TRACE_QUEUE_CAP);                        // [TRACE_LONG_CAP] marches parked between two slices
    uint8_t* sLongFlag = (uint8_t*)(sLong + TRACE_LONG_CAP);                          // [TRACE_LONG_CAP] the parked march is unfinished
    uint8_t* sLongList = sLongFlag + TRACE_LONG_CAP;                                  // [TRACE_LONG_CAP] compacted slots of the next slice
    float* sRayColor = (float*)sPairs;                                                // [4][8][8][3], phase F only: the pair list is dead by then
    __shared__ uint32_t sCount, sWarpTotals[8];
    __shared__ int sQueueCount, sQueueNext, sHitCount, sSurvivors, sLongCount;

    const plain_global_shader_info* g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = tid >> 6;                  // which of the 2x2 groups
    const int lx = tid & 7, ly = (tid >> 3) & 7;
    const int blockRow = (int)blockIdx.y + p.blockRowOffset;
    const int gx = blockIdx.x * 2 + (sub & 1), gy = blockRow * 2 + (sub >> 1);
    const bool groupActive = gx < p.groupsX && gy < p.groupsY;
    // tile of the block (identical for its four groups)
    const int tileX = (blockIdx.x * 2) / 4, tileY = (blockRow * 2) / 4;
    const uint32_t tileIndex = tileIndexFromTileUV(tileX, tileY, g);
    const bool tileValid = (size_t)tileIndex < p.tileCapacity;
    if (tid == 0) { sCount = tileValid ? min(p.tiles[tileIndex].objectCount, (uint32_t)PLAIN_MAX_OBJECTS_PER_TILE) : 0u; sHitCount = 0; }
    __syncthreads();
    const uint32_t objectCount = sCount;
    // ---- A: stage the tile's instances ----
    const plain_sdf_instance* instances = (const plain_sdf_instance*)(p.instanceBuffer + 16);
    for (uint32_t i = tid; i < objectCount; i += 256) {
        const plain_sdf_instance in = instances[p.tiles[tileIndex].indices[i]];
        TraceInst2 t;
        bool finite = true;
        for (int k = 0; k < 16; k++) { t.worldToLocal[k] = in.worldToLocal[k]; finite = finite && absf(in.worldToLocal[k]) < 3.0e38f; }
        t.localExtends = ld3(in.localExtends);
        const BindlessEntry be = p.bindless[in.sdfTextureIndex];
        t.sdf = be.view;
        t.corners = (be.format == PLAIN_FORMAT_R16_SFLOAT) ? be.corners : nullptr;
        instanceBoundingSphere(t);
        t.localExtendsHalfPadded = t.localExtends * 0.5f + 0.01f;
        t.distanceThreshold = length(t.localExtends / v3((float)t.sdf.w, (float)t.sdf.h, (float)t.sdf.d)) * 0.25f;
        t.localToGlobalScale = 1.f / length(v3(t.worldToLocal[0], t.worldToLocal[1], t.worldToLocal[2]));
        t.invLocalExtends = 1.f / t.localExtends;
        t.albedo = vpow(ld3(in.meanAlbedo), v3(2.2f));
        t.dimsF = v3((float)t.sdf.w, (float)t.sdf.h, (float)t.sdf.d);
        t.dimX = t.sdf.w; t.dimY = t.sdf.h;
        t.strideY = t.sdf.w + 1; t.strideZ = (t.sdf.w + 1) * (t.sdf.h + 1);
        // lean march: finite transform, extents in [1e-6, 1e6] (texture coordinates of a point inside the padded box stay below 65536),
        // a corner copy of a brick with sane extents
        const vec3 e = t.localExtends;
        t.fastOk = finite && t.corners != nullptr && e.x >= 1e-6f && e.y >= 1e-6f && e.z >= 1e-6f && e.x <= 1e6f && e.y <= 1e6f && e.z <= 1e6f &&
                   t.sdf.w >= 1 && t.sdf.h >= 1 && t.sdf.d >= 1 && t.sdf.w <= 4096 && t.sdf.h <= 4096 && t.sdf.d <= 4096 && absf(t.distanceThreshold) < 3.0e38f && absf(t.localToGlobalScale) < 3.0e38f;
        sInst[i] = t;
    }
    __syncthreads();

    // ---- B: the thread's own ray and its candidate mask ----
    const Globals G = loadGlobals(g);
    const int ix = gx * 8 + lx, iy = gy * 8 + ly;
    vec3 L = v3(0.f), rayOrigin = v3(0.f);
    uint32_t cand[4] = {0u, 0u, 0u, 0u};
    if (groupActive) {  // uniform per warp: a group is two whole warps
        const vec2 uv = v2((float)ix, (float)iy) / v2((float)p.outYSH.w, (float)p.outYSH.h);
        const float depth = sampleNearest2D<WRAP_CLAMP, float>([&](int x, int y) { return loadD32(p.depthTexture, x, y); }, p.depthTexture.w, p.depthTexture.h, uv, 0.f);
        const float depthLinear = linearizeDepth(depth, G.nearPlane, G.farPlane);
        const vec2 pixelNDC = uv * 2.f - 1.f;
        const vec3 V = -viewDirFromNDC(G, pixelNDC);
        const vec3 pWorld = G.camPos + V / dot(V, G.fwd) * depthLinear;
        const ImgView noiseTex = p.bindless[g->noiseTextureIndices[g->frameIndexMod4]].view;
        const vec2 noiseUV = v2((float)ix, (float)iy) / v2((float)noiseTex.w, (float)noiseTex.h);
        const vec2 xi = sampleNearest2D<WRAP_REPEAT, vec2>([&](int x, int y) { return loadRG8(noiseTex, x, y); }, noiseTex.w, noiseTex.h, noiseUV, v2(0.f));
        const vec3 normalTexel = sampleNearest2D<WRAP_CLAMP, vec3>([&](int x, int y) { return loadRGBA8rgb(p.normalTexture, x, y); }, p.normalTexture.w, p.normalTexture.h, uv, v3(0.f));
        const vec3 N = normalTexel * 2.f - 1.f;
        sRayNormal[tid * 3 + 0] = N.x; sRayNormal[tid * 3 + 1] = N.y; sRayNormal[tid * 3 + 2] = N.z;
        sRayDepth[tid] = depthLinear;
        rayOrigin = pWorld + N * 0.2f;
        L = importanceSampleCosine(xi, N);
        const float invLen2 = 1.f / dot(L, L);
        // the bounding-sphere test pays in crowded tiles (it is five times cheaper than the slab rejection of D1a that follows); in a tile
        // with a handful of instances it rejects next to nothing (ncu, round 2: 7.3 instances per tile, 8 pairs per ray survive) and is skipped
        const bool sphereTest = !(p.variant & 32) || objectCount > 12u;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t m = 0u;
            const uint32_t base = (uint32_t)w * 32u;
            if (base < objectCount) {
                const uint32_t n = min(32u, objectCount - base);
                if (sphereTest) {
                    for (uint32_t b = 0; b < n; b++)
                        if (!rayMissesSphere(sInst[base + b], rayOrigin, L, invLen2)) m |= 1u << b;
                } else {
                    m = n >= 32u ? 0xffffffffu : ((1u << n) - 1u);
                }
            }
            cand[w] = m;
        }
    }
    sRayO[tid * 3 + 0] = rayOrigin.x; sRayO[tid * 3 + 1] = rayOrigin.y; sRayO[tid * 3 + 2] = rayOrigin.z;
    sRayL[tid * 3 + 0] = L.x; sRayL[tid * 3 + 1] = L.y; sRayL[tid * 3 + 2] = L.z;
    sKey[tid] = (unsigned long long)dm::f2u(10000.f) << 32;  // closestHitDistance = 10000, no winner
    sHit[tid] = 0;
    // ---- C: block-wide inclusive prefix sum of the candidate counts ----
    const uint32_t cnt = (uint32_t)(__popc(cand[0]) + __popc(cand[1]) + __popc(cand[2]) + __popc(cand[3]));
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) sWarpTotals[warp] = incl;
    __syncthreads();
    {
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w < warp) before += sWarpTotals[w];
        incl += before;
    }
    sIncl[tid] = incl;
    __syncthreads();

    // ---- batches of rays whose pairs fit the list (one batch unless the tile is crowded) ----
    int batchBegin = 0;            // first ray of the batch
    uint32_t pairBase = 0;         // pairs of the rays before the batch
    while (batchBegin < 256) {
        // rays [batchBegin, batchEnd): the longest run whose pairs fit; the prefix sums are monotone, so it is a count
        const bool fits = tid >= batchBegin && incl - pairBase <= (uint32_t)TRACE_PAIR_CAP;
        const int batchEnd = batchBegin + __syncthreads_count(fits ? 1 : 0);
        const int pairEnd = (int)(sIncl[batchEnd - 1] - pairBase);  // pairs in the batch (a single ray always fits: <= 100)
        if (tid == 0) { sQueueCount = 0; sQueueNext = 0; sSurvivors = 0; sLongCount = 0; }
        if (tid >= batchBegin && tid < batchEnd) {  // write this ray's pairs in list order
            uint32_t at = incl - cnt - pairBase;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t m = cand[w];
                while (m) { const int b = __ffs(m) - 1; m &= m - 1; sPairs[at++] = (uint16_t)((tid << 8) | (w * 32 + b)); }
            }
        }
        __syncthreads();
        // ---- D1a: conservative slab rejection of every pair; the survivors are compacted in place (all 256 threads busy) ----
        {
            TraceResult tr;
            tr.hit = false; tr.closestHitDistance = 10000.f; tr.hitCount = 0; tr.hitPos = v3(0.f); tr.N = v3(0.f); tr.albedo = v3(0.f);
            tr.winner = -1; tr.winnerSamplePos = v3(0.f); tr.winnerRayDirection = v3(0.f); tr.winnerD = 0.f; tr.winnerDLast = 0.f;
            // a hit that replaced the closest hit this lane knew of: into the ray's key (the atomicMin decides whether it really is the closest)
            auto recordHit = [&](int ray, int cur) {
                sHit[ray] = 1;
                if (tr.winner != cur) return;  // d < threshold, but not closer than the closest hit known when the pair was taken
                // another lane may have reported a closer hit (or an equal one listed earlier) since: this one cannot win, no record needed
                if ((traceKey(tr.closestHitDistance, cur, 0u) >> 16) >= (sKey[ray] >> 16)) return;
                uint32_t slot = (uint32_t)atomicAdd(&sHitCount, 1);
                if (slot < (uint32_t)TRACE_HIT_POOL) {
                    HitRecord h;
                    h.samplePos = tr.winnerSamplePos; h.d = tr.winnerD; h.dLast = tr.winnerDLast;
                    sHitPool[slot] = h;
                } else {
                    slot = TRACE_NO_SLOT;
                }
                atomicMin(&sKey[ray], traceKey(tr.closestHitDistance, cur, slot));
            };
            // the ray's closest hit so far (other lanes may be improving it: any value read is a valid bound for SDF.inc:141)
            auto loadClosest = [&](int ray) {
                const unsigned long long key = sKey[ray];
                tr.closestHitDistance = dm::u2f((uint32_t)(key >> 32));
                tr.winner = (int)((key >> 16) & 0xffffu) - 1;
                tr.hit = false;
            };
            // a whole march on the spot: the spelled-out path of non-finite values, and lean marches that found a queue full
            auto marchHere = [&](int ray, int cur, vec3 o, vec3 dir, bool lean, MarchState& st) {
                if (lean) {
                    bool poisoned = false;
                    while (traceStepLean(sInst[cur], cur, tr, st, poisoned)) {}
                    lean = !poisoned;
                }
                if (!lean) traceInstanceSpelledOut(&sInst[cur], cur, o, dir, &tr);
                if (tr.hit) recordHit(ray, cur);
            };
            // the reference's box test (SDF.inc:101-141) of one surviving pair; a ray that enters its box leaves a march job
            auto exactBoxTest = [&](uint32_t pr) {
                const int ray = (int)(pr >> 8), cur = (int)(pr & 0xffu);
                const vec3 o = v3(sRayO[ray * 3], sRayO[ray * 3 + 1], sRayO[ray * 3 + 2]), dir = v3(sRayL[ray * 3], sRayL[ray * 3 + 1], sRayL[ray * 3 + 2]);
                const TraceInst2& inst = sInst[cur];
                const bool lean = inst.fastOk && finite3(o) && finite3(dir);
                loadClosest(ray);
                MarchState st;
                if (!lean) { marchHere(ray, cur, o, dir, false, st); return; }
                if (!traceSetup(inst, o, dir, tr, st)) return;
                if (!(finite3(st.localSamplePos) && finite3(st.rayDirection) && absf(st.hitDistanceLocal) < 3.0e38f)) { marchHere(ray, cur, o, dir, false, st); return; }
                const int slot = atomicAdd(&sQueueCount, 1);
                if (slot < TRACE_QUEUE_CAP) {
                    MarchJob j;
                    j.pos = st.localSamplePos; j.dir = st.rayDirection; j.hitDistanceLocal = st.hitDistanceLocal; j.pair = pr;
                    sQueue[slot] = j;
                } else {
                    marchHere(ray, cur, o, dir, true, st);
                }
            };
            const bool compactSurvivors = (p.variant & 1) != 0;
            for (int base = 0; base < pairEnd; base += 256) {  // uniform trip count: the ballots below need whole warps
                const int q = base + tid;
                uint32_t pr = 0;
                bool keep = false;
                if (q < pairEnd) {
                    pr = sPairs[q];
                    const int ray = (int)(pr >> 8), cur = (int)(pr & 0xffu);
                    const vec3 o = v3(sRayO[ray * 3], sRayO[ray * 3 + 1], sRayO[ray * 3 + 2]), dir = v3(sRayL[ray * 3], sRayL[ray * 3 + 1], sRayL[ray * 3 + 2]);
                    const TraceInst2& inst = sInst[cur];
                    const bool lean = inst.fastOk && finite3(o) && finite3(dir);
                    keep = !(lean && boxCertainlyMissed(inst.worldToLocal, inst.localExtends * 0.5f, o, dir));
                }
                if (!compactSurvivors) {  // the survivor's box test right away (fewer barriers, emptier warps)
                    if (keep) exactBoxTest(pr);
                    continue;
                }
                // survivors go to the front of the same list: position base' <= q, so a warp never overwrites a pair that is still unread
                // (every thread reads its pair of this pass before the block-wide barrier, and writes after it)
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                int at = 0;
                if (lane == 0 && m) at = atomicAdd(&sSurvivors, __popc(m));
                at = __shfl_sync(0xffffffffu, at, 0);
                __syncthreads();
                if (keep) sPairs[at + __popc(m & ((1u << lane) - 1u))] = (uint16_t)pr;
            }
            if (compactSurvivors) {
                __syncthreads();
                const int nSurvivors = sSurvivors;
                for (int q = tid; q < nSurvivors; q += 256) exactBoxTest(sPairs[q]);
            }
            __syncthreads();
            // ---- D2: the marches, in three slices of 8 / 24 / 96 steps (SDF.inc:144: at most 128). Lanes take jobs from a list with an
            //      atomic counter; a lane whose march ends (or whose slice is used up) takes the next job, so a warp stays populated until the
            //      list is empty. A march that outlives its slice is parked in sLong and re-compacted with the other long marches for the next
            //      slice, instead of keeping a warp alive for one lane: ncu showed the rare 100-step marches costing as many issue slots as
            //      all the short ones together ----
            int nJobs = min(sQueueCount, TRACE_QUEUE_CAP);
#pragma unroll 1
            for (int round = 0; round < 3; round++) {
                const int sliceEnd = (p.variant & 2) ? (round == 0 ? 8 : (round == 1 ? 32 : 128)) : 128;  // a march of this round stops when st.k reaches it
                // PLAIN_TRACE_VARIANT bits 3 / 4: only the first 4 / 6 warps march (more jobs per lane: the longest lane of a warp is closer to
                // the average), the others wait at the barrier without using issue slots
                const int marchWarps = (p.variant & 8) ? 4 : ((p.variant & 16) ? 6 : 8);
                bool marching = false, done = warp >= marchWarps;
                int ray = 0, cur = 0, slotLong = -1;
                MarchState st;
                st.localSamplePos = v3(0.f); st.rayDirection = v3(0.f); st.hitDistanceLocal = 0.f; st.d = 0.f; st.dLast = 0.f; st.k = 0;
                while (true) {
                    const unsigned marchMask = __ballot_sync(0xffffffffu, marching), idleMask = __ballot_sync(0xffffffffu, !marching && !done);
                    if ((marchMask | idleMask) == 0u) break;
                    const int nMarch = __popc(marchMask), nIdle = __popc(idleMask);
                    if (nIdle > 0 && (nMarch == 0 || ((p.variant & 4) ? (nIdle >= 16 || nIdle >= nMarch) : (nIdle >= 8 || nIdle * 3 >= nMarch)))) {
                        if (!marching && !done) {  // refill
                            const int i = atomicAdd(&sQueueNext, 1);
                            if (i >= nJobs) {
                                done = true;
                            } else if (round == 0) {
                                const MarchJob j = sQueue[i];
                                ray = (int)(j.pair >> 8); cur = (int)(j.pair & 0xffu);
                                loadClosest(ray);
                                // SDF.inc:141 against the closest hit found since the box test (the same product, the same comparison)
                                if (!(sInst[cur].localToGlobalScale * j.hitDistanceLocal > tr.closestHitDistance)) {
                                    st.localSamplePos = j.pos; st.rayDirection = j.dir; st.hitDistanceLocal = j.hitDistanceLocal; st.d = 0.f; st.dLast = 0.f; st.k = 0;
                                    marching = true; slotLong = -1;
                                }
                            } else {
                                slotLong = sLongList[i];
                                const LongMarch j = sLong[slotLong];
                                ray = (int)(j.pair >> 8); cur = (int)(j.pair & 0xffu);
                                loadClosest(ray);
                                st.localSamplePos = j.pos; st.rayDirection = j.dir; st.hitDistanceLocal = j.hitDistanceLocal; st.d = j.d; st.dLast = j.dLast; st.k = j.k;
                                marching = true;
                            }
                        }
                        continue;
                    }
                    if (marching) {
                        bool poisoned = false;
                        marching = traceStepLean(sInst[cur], cur, tr, st, poisoned);
                        if (poisoned) {  // a non-finite brick value: the whole (ray, instance) again, spelled out (tr is untouched so far)
                            const vec3 o = v3(sRayO[ray * 3], sRayO[ray * 3 + 1], sRayO[ray * 3 + 2]), dir = v3(sRayL[ray * 3], sRayL[ray * 3 + 1], sRayL[ray * 3 + 2]);
                            traceInstanceSpelledOut(&sInst[cur], cur, o, dir, &tr);
                        }
                        if (!marching) {
                            if (tr.hit) recordHit(ray, cur);
                            if (slotLong >= 0) sLongFlag[slotLong] = 0;
                        } else if (st.k >= sliceEnd) {  // the slice is used up: park the march (its own slot again if it already has one)
                            if (slotLong < 0) { slotLong = atomicAdd(&sLongCount, 1); if (slotLong >= TRACE_LONG_CAP) slotLong = -1; }
                            if (slotLong >= 0) {
                                LongMarch j;
                                j.pos = st.localSamplePos; j.dir = st.rayDirection; j.hitDistanceLocal = st.hitDistanceLocal; j.d = st.d; j.dLast = st.dLast; j.k = st.k;
                                j.pair = (uint32_t)((ray << 8) | cur);
                                sLong[slotLong] = j;
                                sLongFlag[slotLong] = 1;
                                marching = false;
                            }  // no slot left: the lane keeps the march to its end
                        }
                    }
                }
                __syncthreads();
                // the parked marches, compacted into the list of the next slice
                if (tid == 0) { sQueueNext = 0; sSurvivors = 0; }
                __syncthreads();
                const int nLong = min(sLongCount, TRACE_LONG_CAP);
                if (tid < nLong && sLongFlag[tid]) sLongList[atomicAdd(&sSurvivors, 1)] = (uint8_t)tid;
                __syncthreads();
                nJobs = sSurvivors;
                if (nJobs == 0) break;
            }
        }
        __syncthreads();
        pairBase += (uint32_t)pairEnd;
        batchBegin = batchEnd;
    }

    // ---- F: every thread shades its own ray ----
    if (groupActive) {
        const unsigned long long key = sKey[tid];
        TraceResult tr;
        tr.hit = sHit[tid] != 0;
        tr.closestHitDistance = dm::u2f((uint32_t)(key >> 32));
        tr.winner = (int)((key >> 16) & 0xffffu) - 1;
        tr.hitCount = 0; tr.hitPos = v3(0.f); tr.N = v3(0.f); tr.albedo = v3(0.f);
        tr.winnerSamplePos = v3(0.f); tr.winnerRayDirection = v3(0.f); tr.winnerD = 0.f; tr.winnerDLast = 0.f;
        if (tr.winner >= 0) {
            const uint32_t slot = (uint32_t)(key & 0xffffu);
            if (slot != TRACE_NO_SLOT) {
                const HitRecord h = sHitPool[slot];
                tr.winnerSamplePos = h.samplePos; tr.winnerD = h.d; tr.winnerDLast = h.dLast;
                // the local ray direction of this (ray, instance), SDF.inc:103-107: the expression traceSetup evaluated for the march
                const TraceInst2& wi = sInst[tr.winner];
                const vec3 startLocal = xyz(mulm4(wi.worldToLocal, v4(rayOrigin, 1.f))), endLocal = xyz(mulm4(wi.worldToLocal, v4(rayOrigin + L, 1.f)));
                const vec3 dirLocal = endLocal - startLocal;
                tr.winnerRayDirection = dirLocal / length(dirLocal);
            } else {  // the block ran out of hit records: march the winning pair again (same operations, same state)
                TraceResult again;
                again.hit = false; again.closestHitDistance = 10000.f; again.hitCount = 0; again.winner = -1;
                again.hitPos = v3(0.f); again.N = v3(0.f); again.albedo = v3(0.f);
                again.winnerSamplePos = v3(0.f); again.winnerRayDirection = v3(0.f); again.winnerD = 0.f; again.winnerDLast = 0.f;
                traceInstanceSpelledOut(&sInst[tr.winner], tr.winner, rayOrigin, L, &again);
                tr.winnerSamplePos = again.winnerSamplePos; tr.winnerRayDirection = again.winnerRayDirection; tr.winnerD = again.winnerD; tr.winnerDLast = again.winnerDLast; tr.hitCount = again.hitCount;
            }
            shadeWinner2(sInst[tr.winner], rayOrigin, L, tr);
        }
        vec3 hitColor;
        if (tr.hit) {
            const float shadow = simpleShadow<true>(tr.hitPos, p.cascades->lightMatrices[p.shadowCascadeIndex], p.shadowMap);
            const vec3 sunLight = shadow * p.light->sunStrengthExposed * ld3(p.light->sunColor);
            hitColor = tr.albedo * sunLight;
            bool hitInRange = tr.closestHitDistance < *p.influenceRange;
            hitInRange = hitInRange || !p.strictInfluenceRadiusCutoff;
            const bool selfIntersection = tr.closestHitDistance < 0.0001f;
            if (!hitInRange || selfIntersection) hitColor = v3(0.f);
        } else {
            hitColor = sampleSkyLut(L, p.skyLut);
        }
        sRayColor[tid * 3 + 0] = hitColor.x; sRayColor[tid * 3 + 1] = hitColor.y; sRayColor[tid * 3 + 2] = hitColor.z;
    }
    __syncthreads();
    if (!groupActive) return;
    // resolveColor :70-116 (ray caches indexed [sub][ly][lx] = thread id)
    auto rayIdx = [&](int rx, int ry) { return sub * 64 + ry * 8 + rx; };
    float weightTotal = 1.f;
    vec3 color = v3(sRayColor[tid * 3], sRayColor[tid * 3 + 1], sRayColor[tid * 3 + 2]);
    const vec3 myN = v3(sRayNormal[tid * 3], sRayNormal[tid * 3 + 1], sRayNormal[tid * 3 + 2]);
    const float myDepth = sRayDepth[tid];
    for (int x = -1; x <= 1; x++) {
        for (int y = -1; y <= 1; y++) {
            if (x == 0 && y == 0) continue;
            const int rx = lx + x, ry = ly + y;
            const bool isValidIndex = rx > 0 && ry > 0 && rx < 8 && ry < 8;  // greaterThan(rayIndex, 0): row/column 0 never used (:88)
            if (!isValidIndex) continue;
            const int ri = rayIdx(rx, ry);
            const vec3 nN = v3(sRayNormal[ri * 3], sRayNormal[ri * 3 + 1], sRayNormal[ri * 3 + 2]);
            const float NoN = clampf(dot(myN, nN), 0.f, 1.f);
            const bool normalsMatch = NoN > 0.9f;
            const bool depthMatch = absf(myDepth - sRayDepth[ri]) < 0.5f;
            if (normalsMatch && depthMatch) {
                const float weightX = x == 0 ? 1.f : 0.5f, weightY = y == 0 ? 1.f : 0.5f;
                const float weight = weightX * weightY;
                color = color + weight * v3(sRayColor[ri * 3], sRayColor[ri * 3 + 1], sRayColor[ri * 3 + 2]);
                weightTotal += weight;
            }
        }
    }
    color = color / weightTotal;
    const vec3 YCoCg = linearToYCoCg(color);
    vec4 result_Y_SH = v4(0.f);
    vec2 result_CoCg = v2(0.f);
    result_Y_SH = result_Y_SH + YCoCg.x * directionToSH_L1(L);
    result_CoCg = result_CoCg + v2(YCoCg.y, YCoCg.z);
    if (inRange(p.outYSH, ix, iy)) storeRGBA16F(p.outYSH, ix, iy, 0, result_Y_SH);
    if (inRange(p.outCoCg, ix, iy)) storeRG16F(p.outCoCg, ix, iy, result_CoCg);
}
static size_t traceSharedBytes() {
    static_assert(sizeof(uint16_t) * TRACE_PAIR_CAP >= sizeof(float) * 256 * 3, "the ray colours reuse the pair list");
    return sizeof(TraceInst2) * PLAIN_MAX_OBJECTS_PER_TILE + sizeof(unsigned long long) * 256 + sizeof(HitRecord) * TRACE_HIT_POOL + sizeof(float) * (256 * 3 * 3 + 256) +
           sizeof(uint32_t) * 256 + sizeof(uint16_t) * TRACE_PAIR_CAP + 256 + sizeof(MarchJob) * TRACE_QUEUE_CAP + sizeof(LongMarch) * TRACE_LONG_CAP + 2 * TRACE_LONG_CAP;
}
PLAIN_PASS(launch_sdfDiffuseTrace, "sdfDiffuseTrace.comp") {
    TraceParams p;
    p.outYSH = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.outCoCg = c.storage(1, PLAIN_FORMAT_RG16_SFLOAT);
    p.depthTexture = c.sampled(2, PLAIN_FORMAT_DEPTH32);
    p.normalTexture = c.sampled(3, PLAIN_FORMAT_RGBA8);
    p.skyLut = c.sampled(4, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.shadowMap = c.sampled(10, PLAIN_FORMAT_DEPTH16);
    p.light = c.sbuf<plain_light_buffer>(5);
    p.instanceBuffer = c.sbuf<unsigned char>(6);
    size_t tilesSize = 0;
    p.tiles = c.sbuf<plain_culled_instances_per_tile>(7, &tilesSize);
    p.tileCapacity = tilesSize / sizeof(plain_culled_instances_per_tile);
    p.influenceRange = c.ubuf<float>(8);
    p.cascades = c.sbuf<plain_shadow_cascade_info>(9);
    p.g = c.g;
    p.bindless = c.bindless;
    p.strictInfluenceRadiusCutoff = c.specBool(0, false) ? 1 : 0;
    p.shadowCascadeIndex = c.spec<int>(1, 3);
    p.groupsX = (int)c.exec->dispatch[0];
    p.groupsY = (int)c.exec->dispatch[1];
    if (c.failed) return;
    if (p.shadowCascadeIndex < 0 || p.shadowCascadeIndex > 3) { c.fail("sdfDiffuseTrace.comp: shadow cascade index must be 0..3"); return; }
    if (p.groupsX == 0 || p.groupsY == 0) return;
    int y0, y1;
    c.window(p.groupsY * 8, y0, y1);  // row sharding unit: rows of the (half-res) trace target, in multiples of the 16-row blocks
    if (y0 % 16 != 0) { c.fail("sdfDiffuseTrace.comp: row window must start at a multiple of 16 rows"); return; }
    if (y1 <= y0) return;
    p.blockRowOffset = y0 / 16;
    // default 1: survivors compacted, long marches not sliced (measured slower: profiles/r2_trace_variants.md), eager refill
    static const int variant = getenv("PLAIN_TRACE_VARIANT") ? atoi(getenv("PLAIN_TRACE_VARIANT")) : 1;
    p.variant = variant;
    const size_t smem = traceSharedBytes();  // above the 48 KB static limit: opt in (per device; a host-side attribute, not a stream operation)
    if (cudaFuncSetAttribute(sdfDiffuseTraceKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { c.fail("sdfDiffuseTrace.comp: cannot reserve shared memory"); return; }
    PLAIN_LAUNCH(c, sdfDiffuseTraceKernel, dim3(ceilDiv(p.groupsX, 2), ceilDiv((unsigned)(y1 - y0), 16)), 256, smem, p);
}

// ---------------- sdfDebugVisualisation.comp ----------------
struct DebugVisParams {
    ImgView imageOut, skyLut, shadowMap;
    const plain_light_buffer* light;
    const unsigned char* instanceBuffer;  // uvec4 header + SDFInstance[]
    const plain_culled_instances_per_tile* tiles;
    size_t tileCapacity;
    const plain_shadow_cascade_info* cascades;
    const plain_global_shader_info* g;
    const BindlessEntry* bindless;
    int debugMode, shadowCascadeIndex;
    int y0, y1;
};
// Primary rays through the SDF scene (:74-133). Block = 32x8 pixels of one 32x32 culling tile, whose instance records are
// staged once in shared memory (as in sdfDiffuseTraceKernel); every pixel walks the tile's list in the reference's order.
__global__ void __launch_bounds__(256) sdfDebugVisualisationKernel(const __grid_constant__ DebugVisParams p) {
    __shared__ TraceInstance sInst[PLAIN_MAX_OBJECTS_PER_TILE];
    __shared__ uint32_t sCount;
    const plain_global_shader_info* g = p.g;
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    const uint32_t tileIndex = tileIndexFromTileUV(blockIdx.x, (p.y0 + blockIdx.y * 8) / 32, g);
    const bool tileValid = (size_t)tileIndex < p.tileCapacity;
    if (threadIdx.x == 0) sCount = tileValid ? p.tiles[tileIndex].objectCount : 0u;
    __syncthreads();
    const uint32_t objectCountRaw = sCount;  // the tile culling pass caps it at maxObjectsPerTile
    const uint32_t objectCount = min(objectCountRaw, (uint32_t)PLAIN_MAX_OBJECTS_PER_TILE);
    const plain_sdf_instance* instances = (const plain_sdf_instance*)(p.instanceBuffer + 16);
    for (uint32_t i = threadIdx.x; i < objectCount; i += 256) {
        const plain_sdf_instance in = instances[p.tiles[tileIndex].indices[i]];
        TraceInstance t;
        for (int k = 0; k < 16; k++) t.worldToLocal[k] = in.worldToLocal[k];
        t.localExtends = ld3(in.localExtends);
        t.meanAlbedo = ld3(in.meanAlbedo);
        t.sdf = p.bindless[in.sdfTextureIndex].view;
        t.sphereCenter = v3(0.f);
        t.sphereR2 = 0.f;
        t.localExtendsHalfPadded = t.localExtends * 0.5f + 0.01f;
        t.distanceThreshold = length(t.localExtends / v3((float)t.sdf.w, (float)t.sdf.h, (float)t.sdf.d)) * 0.25f;
        t.localToGlobalScale = 1.f / length(v3(t.worldToLocal[0], t.worldToLocal[1], t.worldToLocal[2]));
        t.invLocalExtends = 1.f / t.localExtends;
        sInst[i] = t;
    }
    __syncthreads();
    if (ix >= p.imageOut.w || iy >= p.y1) return;
    const Globals G = loadGlobals(g);
    const vec2 pixelCoor = (v2((float)ix, (float)iy) / v2((float)g->screenResolution[0], (float)g->screenResolution[1]) - 0.5f) * 2.f;
    const vec3 cameraToPixel = -viewDirFromNDC(G, pixelCoor);
    const vec3 rayStart = G.camPos + G.nearPlane * cameraToPixel;
    TraceResult tr;
    tr.hit = false;
    tr.closestHitDistance = 10000.f;
    tr.hitCount = 0;
    tr.hitPos = v3(0.f); tr.N = v3(0.f); tr.albedo = v3(0.f);
    tr.winner = -1;
    tr.winnerSamplePos = v3(0.f); tr.winnerRayDirection = v3(0.f); tr.winnerD = 0.f; tr.winnerDLast = 0.f;
    for (uint32_t i = 0; i < objectCount; i++) {
        MarchState st;
        if (!traceSetup(sInst[i], rayStart, cameraToPixel, tr, st)) continue;
        while (traceStep(sInst[i], (int)i, tr, st)) {}
    }
    if (tr.winner >= 0) shadeWinner(sInst[tr.winner], rayStart, cameraToPixel, tr);
    const float shadow = simpleShadow<false>(tr.hitPos, p.cascades->lightMatrices[p.shadowCascadeIndex], p.shadowMap);
    vec3 color = v3(0.f);
    if (tr.hit || p.debugMode == 2) {
        if (p.debugMode == 1) {
            vec3 sunLight = p.light->sunStrengthExposed * ld3(p.light->sunColor);
            sunLight = sunLight * shadow;
            const vec3 ambient = v3(0.15f);
            const float NoL = clampf(dot(tr.N, v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2])), 0.f, 1.f);
            color = tr.albedo * (ambient + sunLight * NoL);
        } else if (p.debugMode == 2) {
            const float percentage = (float)objectCountRaw / (float)PLAIN_MAX_OBJECTS_PER_TILE;
            color = percentage >= 1.f ? v3(1.f, 0.f, 0.f) : v3(percentage);
        } else if (p.debugMode == 3) {
            color = tr.N * 0.5f + 0.5f;
        } else if (p.debugMode == 4) {
            color = v3((float)tr.hitCount / 128.f);
        }
    } else {
        color = sampleSkyLut(cameraToPixel, p.skyLut);
    }
    storeR11(p.imageOut, ix, iy, color);
}
PLAIN_PASS(launch_sdfDebugVisualisation, "sdfDebugVisualisation.comp") {
    DebugVisParams p;
    p.imageOut = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.skyLut = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.shadowMap = c.sampled(7, PLAIN_FORMAT_DEPTH16);
    p.light = c.sbuf<plain_light_buffer>(1);
    p.instanceBuffer = c.sbuf<unsigned char>(3);
    size_t tilesSize = 0;
    p.tiles = c.sbuf<plain_culled_instances_per_tile>(4, &tilesSize);
    p.tileCapacity = tilesSize / sizeof(plain_culled_instances_per_tile);
    p.cascades = c.sbuf<plain_shadow_cascade_info>(6);
    p.g = c.g;
    p.bindless = c.bindless;
    p.debugMode = c.spec<int>(0, 0);
    p.shadowCascadeIndex = c.spec<int>(1, 3);
    if (c.failed) return;
    if (p.shadowCascadeIndex < 0 || p.shadowCascadeIndex > 3) { c.fail("sdfDebugVisualisation.comp: shadow cascade index must be 0..3"); return; }
    if ((int)c.exec->dispatch[0] * 8 < p.imageOut.w || (int)c.exec->dispatch[1] * 8 < p.imageOut.h) { c.fail("sdfDebugVisualisation.comp: dispatch does not cover the target"); return; }
    c.window(p.imageOut.h, p.y0, p.y1);
    if (p.y0 % 8 != 0) { c.fail("sdfDebugVisualisation.comp: row window must start at a multiple of 8 rows"); return; }
    if (p.y1 <= p.y0) return;
    PLAIN_LAUNCH(c, sdfDebugVisualisationKernel, dim3(ceilDiv(p.imageOut.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8)), 256, 0, p);
}

// ---------------- filterIndirectDiffuseSpatial.comp ----------------
struct SpatialParams {
    ImgView outYSH, outCoCg, texYSH, texCoCg, depthTexture, normalTexture;
    const plain_global_shader_info* g;
    const ShadingTables* tables;
    int filterIndex;
    int y0, y1;  // rows to produce (row sharding)
};
// depth texel -> world position for a pixel at uv (filterIndirectDiffuseSpatial.comp:21-28); the nearest-sampled depth is passed in
__device__ __forceinline__ vec3 giDepthToWorld(float depth, const Globals& G, vec2 uv) {
    const float depthLinear = linearizeDepth(depth, G.nearPlane, G.farPlane);
    const vec2 pixelNDC = uv * 2.f - 1.f;
    const vec3 cameraToPixel = -viewDirFromNDC(G, pixelNDC);
    return G.camPos + cameraToPixel / dot(cameraToPixel, G.fwd) * depthLinear;
}
// nearest + clamp-to-edge texel of a (sanitised) coordinate: image_view.h nearestTexel + wrapIndex<WRAP_CLAMP>
__device__ __forceinline__ ivec2 nearestClampTexel(vec2 uvSanitized, int w, int h) {
    ivec2 t;
    t.x = iclamp(floor2i(uvSanitized.x * (float)w), 0, w - 1);
    t.y = iclamp(floor2i(uvSanitized.y * (float)h), 0, h - 1);
    return t;
}
template <bool DEPTH_IS_R16F>
__device__ __forceinline__ vec3 giPixelToWorld(const ImgView& depthTexture, const Globals& G, vec2 uv) {
    const ivec2 t = nearestClampTexel(v2(sanitizeCoord(uv.x), sanitizeCoord(uv.y)), depthTexture.w, depthTexture.h);
    return giDepthToWorld(DEPTH_IS_R16F ? loadR16F(depthTexture, t.x, t.y) : loadD32(depthTexture, t.x, t.y), G, uv);
}
// The 32 disc samples come from one xorshift sequence that is the same for every pixel (:60-70): sqrt(rand), cos(angle),
// sin(angle) are tabulated once per context for the seeds a frame can use (ShadingTables::disc; any other seed is computed by
// the block's first warp with the same functions); each pixel only applies its own lengthModifier.
// Per sample, when the depth, Y_SH and CoCg images have the same extent - which is how the frontend creates them - the
// nearest texel is computed once for the three fetches (same expression, same operands: same result). The coordinate is
// NOT sanitised before the nearest + clamp-to-edge lookup: cvt.rmi saturates and maps NaN to 0, so
// clamp(floor2i(u * w), 0, w - 1) returns the same texel for a NaN (0) and for |u| > 65536 (the edge) as the sanitised
// coordinate does. The loop is software-pipelined by hand over two samples (A / B): the three texels of sample i+1 are
// requested before the arithmetic of sample i runs (the position of sample i+1 depends on sample i only through
// lengthModifier, which is known as soon as sample i's coordinate is); Y_SH / CoCg are fetched speculatively (clamped
// addresses are always valid) and only used when the reference would have sampled them.
struct SpatialFetch {
    vec2 uv;        // sampleUV after the border fix-up
    uint32_t depth; // R16F: the half in the low 16 bits; D32F: the binary32 bits
    uint2 ysh;
    uint32_t cocg;
};
template <bool DEPTH_IS_R16F, bool SAME_EXTENT>
__global__ void __launch_bounds__(256, 4) giSpatialFilterKernel(const __grid_constant__ SpatialParams p) {
    __shared__ float4 sDisc[32];
    __shared__ float sVP[16];
    const plain_global_shader_info* g = p.g;
    const uint32_t seedIndex = g->frameIndexMod4 + (uint32_t)p.filterIndex;
    if (seedIndex < PLAIN_DISC_SEEDS) {
        if (threadIdx.x < 32) sDisc[threadIdx.x] = __ldg(&p.tables->disc[seedIndex * 32 + threadIdx.x]);
    } else if (threadIdx.x == 0) {
        uint32_t rngState = wang_hash(seedIndex);
        for (int i = 0; i < 32; i++) {
            const float sq = sqrtf_(rand01(rngState));
            const float angle = 2.f * PV_PI * rand01(rngState);
            sDisc[i] = make_float4(sq, dm::cos(angle), dm::sin(angle), 0.f);
        }
    }
    if (threadIdx.x >= 32 && threadIdx.x < 48) sVP[threadIdx.x - 32] = g->viewProjection[threadIdx.x - 32];
    __syncthreads();
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.outYSH.w || iy >= p.y1) return;
    const Globals G = loadGlobals(g);
    const vec2 texelSize = 1.f / v2((float)p.outYSH.w, (float)p.outYSH.h);
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) * texelSize;
    const vec3 pCenter = giPixelToWorld<DEPTH_IS_R16F>(p.depthTexture, G, uv);
    const vec3 pRight = giPixelToWorld<DEPTH_IS_R16F>(p.depthTexture, G, uv + v2(1.f, 0.f) * texelSize);
    const vec3 pUp = giPixelToWorld<DEPTH_IS_R16F>(p.depthTexture, G, uv + v2(0.f, 1.f) * texelSize);
    const vec3 tangent = normalize(pCenter - pRight);
    const vec3 bitangent = normalize(pCenter - pUp);
    const vec3 N = 2.f * sampleNearest2D<WRAP_CLAMP, vec3>([&](int x, int y) { return loadRGBA8rgb(p.normalTexture, x, y); }, p.normalTexture.w, p.normalTexture.h, uv, v3(0.f)) - 1.f;
    float radiusWorld = 1.5f;
    if (p.filterIndex == 1) radiusWorld = 1.f;
    const float dW = (float)p.depthTexture.w, dH = (float)p.depthTexture.h;
    const int dWm1 = p.depthTexture.w - 1, dHm1 = p.depthTexture.h - 1;
    // position + texel requests of sample i for the current lengthModifier (:72-98)
    auto fetchSample = [&](int i, float lengthModifier) {
        SpatialFetch f;
        const float4 disc = sDisc[i];
        const float d = disc.x * lengthModifier;
        const vec2 offset = v2(disc.y, disc.z) * d;
        const vec3 sampleWorld = pCenter + radiusWorld * (offset.x * tangent + offset.y * bitangent);
        // viewProjection * vec4(sampleWorld, 1): only x, y, w are used
        const float px = fmaf_(sVP[12], 1.f, fmaf_(sVP[8], sampleWorld.z, fmaf_(sVP[4], sampleWorld.y, sVP[0] * sampleWorld.x)));
        const float py = fmaf_(sVP[13], 1.f, fmaf_(sVP[9], sampleWorld.z, fmaf_(sVP[5], sampleWorld.y, sVP[1] * sampleWorld.x)));
        const float pw = fmaf_(sVP[15], 1.f, fmaf_(sVP[11], sampleWorld.z, fmaf_(sVP[7], sampleWorld.y, sVP[3] * sampleWorld.x)));
        const float rw = rcpf_(pw);
        vec2 sampleUV = v2(px * rw, py * rw);
        sampleUV = sampleUV * 0.5f + 0.5f;
        // x < 0 ? alt : x followed by x > 1 ? alt : x selects alt exactly when the first value is outside [0, 1] (NaN keeps itself)
        const float altX = uv.x - offset.x, altY = uv.y - offset.y;
        sampleUV.x = (sampleUV.x < 0.f || sampleUV.x > 1.f) ? altX : sampleUV.x;
        sampleUV.y = (sampleUV.y < 0.f || sampleUV.y > 1.f) ? altY : sampleUV.y;
        f.uv = sampleUV;
        const int tx = iclamp(floor2i(sampleUV.x * dW), 0, dWm1), ty = iclamp(floor2i(sampleUV.y * dH), 0, dHm1);
        const int texel = ty * p.depthTexture.w + tx;
        f.depth = DEPTH_IS_R16F ? (uint32_t)ldg((const uint16_t*)p.depthTexture.ptr + texel) : ldg((const uint32_t*)p.depthTexture.ptr + texel);
        if (SAME_EXTENT) {
            f.ysh = ldg((const uint2*)p.texYSH.ptr + texel);
            f.cocg = ldg((const uint32_t*)p.texCoCg.ptr + texel);
        } else {
            const ivec2 tyS = nearestClampTexel(sampleUV, p.texYSH.w, p.texYSH.h), tcS = nearestClampTexel(sampleUV, p.texCoCg.w, p.texCoCg.h);
            f.ysh = ldg((const uint2*)p.texYSH.ptr + texelIndex(p.texYSH, tyS.x, tyS.y));
            f.cocg = ldg((const uint32_t*)p.texCoCg.ptr + texelIndex(p.texCoCg, tcS.x, tcS.y));
        }
        return f;
    };
    vec4 result_Y_SH = v4(0.f);
    vec2 result_CoCg = v2(0.f);
    float weightTotal = 0.f;
    float lengthModifier = 1.f;
    auto isOutside = [](const SpatialFetch& f) { return f.uv.x < 0.f || f.uv.y < 0.f || f.uv.x > 1.f || f.uv.y > 1.f; };
    auto accumulate = [&](const SpatialFetch& cur, bool outside) {
        const float depth = DEPTH_IS_R16F ? halfToFloat((uint16_t)cur.depth) : dm::u2f(cur.depth);
        const vec3 pixelWorld = giDepthToWorld(depth, G, cur.uv);
        const float distanceToTangentPlane = absf(dot(N, pixelWorld - pCenter));
        // clamp(0.25 / max(dist, 0.0001), 0, 1): the quotient is positive and never NaN (max drops a NaN distance), so the lower
        // clamp is the identity and FMNMX against the non-zero constants returns the bits of the pinned min / max
        float weight = fminf(0.25f / fmaxf(distanceToTangentPlane, 0.0001f), 1.f);
        weight *= weight;
        if (!outside && weight > 0.f) {
            const vec4 sample_Y_SH = v4(halfToFloat((uint16_t)(cur.ysh.x & 0xffffu)), halfToFloat((uint16_t)(cur.ysh.x >> 16)), halfToFloat((uint16_t)(cur.ysh.y & 0xffffu)), halfToFloat((uint16_t)(cur.ysh.y >> 16)));
            const vec2 sample_CoCg = v2(halfToFloat((uint16_t)(cur.cocg & 0xffffu)), halfToFloat((uint16_t)(cur.cocg >> 16)));
            if (!(anynan(sample_Y_SH) || anynan(sample_CoCg))) {
                result_Y_SH = result_Y_SH + weight * sample_Y_SH;
                result_CoCg = result_CoCg + weight * sample_CoCg;
                weightTotal += weight;
            }
        }
    };
#if defined(PV_F32X2)
    // ---- pair path (round 2): two disc samples per iteration in the two halves of packed binary32 arithmetic ----
    // Every multiply / add / fma of the per-sample chain (:72-112) is issued once for samples (i, i + 1) as FMUL2 / FFMA2: the operations,
    // their operands and their order are those of fetchSample / accumulate above, per half. The three reciprocals, the square root and the
    // two divisions of a sample run as their unguarded fast paths (pvec.h rcp2_normal / sqrt2_normal / div2_normal), and the range tests
    // that make those valid are collected in ONE flag per pixel: if any such operand of any sample was outside 2^-60 .. 2^60 the results
    // of this loop are discarded and the spelled-out loop below runs instead (never, for a sane camera and depth buffer).
    // Sample i + 1's radius depends on whether sample i landed off screen (:101-106): it is computed assuming it did not, and recomputed
    // when it did (only pixels whose world-space disc leaves the screen).
    bool pairOk = true;
    {
        const float nearFar = G.nearPlane * G.farPlane, nearMinusFar = G.nearPlane - G.farPlane, tanAspect = G.tanFovHalf * G.aspect;
        auto P = [](float v) { return pk2(v, v); };
        struct PairPos { float2 u, v; int texelA, texelB; bool outsideA, outsideB; };
        auto positions = [&](int i, float lmA, float lmB) {
            PairPos r;
            const float4 dA = sDisc[i], dB = sDisc[i + 1];
            const float2 d = mul2(pk2(dA.x, dB.x), pk2(lmA, lmB));
            const float2 ox = mul2(pk2(dA.y, dB.y), d), oy = mul2(pk2(dA.z, dB.z), d);
            // sampleWorld = pCenter + radiusWorld * (offset.x * tangent + offset.y * bitangent)
            const float2 wx = add2(P(pCenter.x), mul2(P(radiusWorld), add2(mul2(ox, P(tangent.x)), mul2(oy, P(bitangent.x)))));
            const float2 wy = add2(P(pCenter.y), mul2(P(radiusWorld), add2(mul2(ox, P(tangent.y)), mul2(oy, P(bitangent.y)))));
            const float2 wz = add2(P(pCenter.z), mul2(P(radiusWorld), add2(mul2(ox, P(tangent.z)), mul2(oy, P(bitangent.z)))));
            const float2 one = P(1.f);
            const float2 px = __ffma2_rn(P(sVP[12]), one, __ffma2_rn(P(sVP[8]), wz, __ffma2_rn(P(sVP[4]), wy, mul2(P(sVP[0]), wx))));
            const float2 py = __ffma2_rn(P(sVP[13]), one, __ffma2_rn(P(sVP[9]), wz, __ffma2_rn(P(sVP[5]), wy, mul2(P(sVP[1]), wx))));
            const float2 pw = __ffma2_rn(P(sVP[15]), one, __ffma2_rn(P(sVP[11]), wz, __ffma2_rn(P(sVP[7]), wy, mul2(P(sVP[3]), wx))));
            pairOk = pairOk && moderate_(pw.x) && moderate_(pw.y);
            const float2 rw = rcp2_normal(pw);
            float2 su = add2(mul2(mul2(px, rw), P(0.5f)), P(0.5f)), sv = add2(mul2(mul2(py, rw), P(0.5f)), P(0.5f));
            const float2 altX = sub2(P(uv.x), ox), altY = sub2(P(uv.y), oy);
            su.x = (su.x < 0.f || su.x > 1.f) ? altX.x : su.x;
            su.y = (su.y < 0.f || su.y > 1.f) ? altX.y : su.y;
            sv.x = (sv.x < 0.f || sv.x > 1.f) ? altY.x : sv.x;
            sv.y = (sv.y < 0.f || sv.y > 1.f) ? altY.y : sv.y;
            r.u = su; r.v = sv;
            const float2 tu = mul2(su, P(dW)), tv = mul2(sv, P(dH));
            r.texelA = iclamp(floor2i(tv.x), 0, dHm1) * p.depthTexture.w + iclamp(floor2i(tu.x), 0, dWm1);
            r.texelB = iclamp(floor2i(tv.y), 0, dHm1) * p.depthTexture.w + iclamp(floor2i(tu.y), 0, dWm1);
            r.outsideA = su.x < 0.f || sv.x < 0.f || su.x > 1.f || sv.x > 1.f;
            r.outsideB = su.y < 0.f || sv.y < 0.f || su.y > 1.f || sv.y > 1.f;
            return r;
        };
        auto nearest = [&](vec2 at, const ImgView& img) { const ivec2 t = nearestClampTexel(at, img.w, img.h); return texelIndex(img, t.x, t.y); };
#pragma unroll 1
        for (int i = 0; i < 32; i += 2) {
            const float lmA = lengthModifier;
            PairPos q = positions(i, lmA, lmA);
            if (q.outsideA) {  // sample i + 1 starts from the reduced radius (sample i's half is recomputed to the same bits)
                lengthModifier = lmA * 0.98f;
                q = positions(i, lmA, lengthModifier);
            }
            if (q.outsideB) lengthModifier *= 0.98f;
            // the three texels of both samples (clamped addresses are always valid; Y_SH / CoCg are only used where the reference samples them)
            uint32_t depthA, depthB, cocgA, cocgB;
            uint2 yshA, yshB;
            depthA = DEPTH_IS_R16F ? (uint32_t)ldg((const uint16_t*)p.depthTexture.ptr + q.texelA) : ldg((const uint32_t*)p.depthTexture.ptr + q.texelA);
            depthB = DEPTH_IS_R16F ? (uint32_t)ldg((const uint16_t*)p.depthTexture.ptr + q.texelB) : ldg((const uint32_t*)p.depthTexture.ptr + q.texelB);
            if (SAME_EXTENT) {
                yshA = ldg((const uint2*)p.texYSH.ptr + q.texelA); cocgA = ldg((const uint32_t*)p.texCoCg.ptr + q.texelA);
                yshB = ldg((const uint2*)p.texYSH.ptr + q.texelB); cocgB = ldg((const uint32_t*)p.texCoCg.ptr + q.texelB);
            } else {
                const vec2 atA = v2(q.u.x, q.v.x), atB = v2(q.u.y, q.v.y);
                yshA = ldg((const uint2*)p.texYSH.ptr + nearest(atA, p.texYSH)); cocgA = ldg((const uint32_t*)p.texCoCg.ptr + nearest(atA, p.texCoCg));
                yshB = ldg((const uint2*)p.texYSH.ptr + nearest(atB, p.texYSH)); cocgB = ldg((const uint32_t*)p.texCoCg.ptr + nearest(atB, p.texCoCg));
            }
            // giDepthToWorld for both samples (:21-28): linearizeDepth, the view vector through the sample's uv, the world position
            const float2 depth = DEPTH_IS_R16F ? pk2(halfToFloat((uint16_t)depthA), halfToFloat((uint16_t)depthB)) : pk2(dm::u2f(depthA), dm::u2f(depthB));
            const float2 den = add2(P(G.farPlane), mul2(add2(neg2(depth), P(1.f)), P(nearMinusFar)));
            const float2 ndcX = sub2(mul2(q.u, P(2.f)), P(1.f)), ndcY = sub2(mul2(q.v, P(2.f)), P(1.f));
            const float2 kUp = mul2(P(G.tanFovHalf), ndcY), kRight = mul2(P(tanAspect), ndcX);
            float2 Vx = sub2(add2(P(-G.fwd.x), mul2(kUp, P(G.up.x))), mul2(kRight, P(G.right.x)));
            float2 Vy = sub2(add2(P(-G.fwd.y), mul2(kUp, P(G.up.y))), mul2(kRight, P(G.right.y)));
            float2 Vz = sub2(add2(P(-G.fwd.z), mul2(kUp, P(G.up.z))), mul2(kRight, P(G.right.z)));
            const float2 len2 = __ffma2_rn(Vz, Vz, __ffma2_rn(Vy, Vy, mul2(Vx, Vx)));
            pairOk = pairOk && moderate_(den.x) && moderate_(den.y) && moderate_(len2.x) && moderate_(len2.y);
            const float2 depthLinear = div2_normal(P(nearFar), den);
            const float2 rl = rcp2_normal(sqrt2_normal(len2));
            // cameraToPixel = -normalize(V)
            const float2 cx = neg2(mul2(Vx, rl)), cy = neg2(mul2(Vy, rl)), cz = neg2(mul2(Vz, rl));
            const float2 cf = __ffma2_rn(cz, P(G.fwd.z), __ffma2_rn(cy, P(G.fwd.y), mul2(cx, P(G.fwd.x))));
            pairOk = pairOk && moderate_(cf.x) && moderate_(cf.y);
            const float2 rcf = rcp2_normal(cf);
            const float2 wX = add2(P(G.camPos.x), mul2(mul2(cx, rcf), depthLinear));
            const float2 wY = add2(P(G.camPos.y), mul2(mul2(cy, rcf), depthLinear));
            const float2 wZ = add2(P(G.camPos.z), mul2(mul2(cz, rcf), depthLinear));
            const float2 dX = sub2(wX, P(pCenter.x)), dY = sub2(wY, P(pCenter.y)), dZ = sub2(wZ, P(pCenter.z));
            const float2 nd = __ffma2_rn(P(N.z), dZ, __ffma2_rn(P(N.y), dY, mul2(P(N.x), dX)));
            const float2 dist = pk2(fmaxf(absf(nd.x), 0.0001f), fmaxf(absf(nd.y), 0.0001f));
            pairOk = pairOk && dist.x <= 1.1529215e18f && dist.y <= 1.1529215e18f;
            float2 weight = div2_normal(P(0.25f), dist);
            weight = pk2(fminf(weight.x, 1.f), fminf(weight.y, 1.f));
            weight = mul2(weight, weight);
            // accumulate in sample order (:118-135)
            if (!q.outsideA && weight.x > 0.f) {
                const vec4 sample_Y_SH = v4(halfToFloat((uint16_t)(yshA.x & 0xffffu)), halfToFloat((uint16_t)(yshA.x >> 16)), halfToFloat((uint16_t)(yshA.y & 0xffffu)), halfToFloat((uint16_t)(yshA.y >> 16)));
                const vec2 sample_CoCg = v2(halfToFloat((uint16_t)(cocgA & 0xffffu)), halfToFloat((uint16_t)(cocgA >> 16)));
                if (!(anynan(sample_Y_SH) || anynan(sample_CoCg))) {
                    result_Y_SH = result_Y_SH + weight.x * sample_Y_SH;
                    result_CoCg = result_CoCg + weight.x * sample_CoCg;
                    weightTotal += weight.x;
                }
            }
            if (!q.outsideB && weight.y > 0.f) {
                const vec4 sample_Y_SH = v4(halfToFloat((uint16_t)(yshB.x & 0xffffu)), halfToFloat((uint16_t)(yshB.x >> 16)), halfToFloat((uint16_t)(yshB.y & 0xffffu)), halfToFloat((uint16_t)(yshB.y >> 16)));
                const vec2 sample_CoCg = v2(halfToFloat((uint16_t)(cocgB & 0xffffu)), halfToFloat((uint16_t)(cocgB >> 16)));
                if (!(anynan(sample_Y_SH) || anynan(sample_CoCg))) {
                    result_Y_SH = result_Y_SH + weight.y * sample_Y_SH;
                    result_CoCg = result_CoCg + weight.y * sample_CoCg;
                    weightTotal += weight.y;
                }
            }
        }
    }
    if (!pairOk) {  // an operand outside the fast sequences' range somewhere in this pixel: everything again, spelled out
        result_Y_SH = v4(0.f);
        result_CoCg = v2(0.f);
        weightTotal = 0.f;
        lengthModifier = 1.f;
#else
    {
#endif
    SpatialFetch A = fetchSample(0, lengthModifier), B;
#pragma unroll 1
    for (int i = 0; i < 32; i += 2) {
        const bool outsideA = isOutside(A);
        if (outsideA) lengthModifier *= 0.98f;
        B = fetchSample(i + 1, lengthModifier);
        accumulate(A, outsideA);
        const bool outsideB = isOutside(B);
        if (outsideB) lengthModifier *= 0.98f;
        if (i + 2 < 32) A = fetchSample(i + 2, lengthModifier);
        accumulate(B, outsideB);
    }
    }
    weightTotal = fmaxp(weightTotal, 0.00001f);
    result_Y_SH = result_Y_SH / weightTotal;
    result_CoCg = result_CoCg / weightTotal;
    storeRGBA16F(p.outYSH, ix, iy, 0, result_Y_SH);
    if (inRange(p.outCoCg, ix, iy)) storeRG16F(p.outCoCg, ix, iy, result_CoCg);
}
PLAIN_PASS(launch_giSpatialFilter, "filterIndirectDiffuseSpatial.comp") {
    SpatialParams p;
    p.outYSH = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.outCoCg = c.storage(1, PLAIN_FORMAT_RG16_SFLOAT);
    p.texYSH = c.sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.texCoCg = c.sampled(3, PLAIN_FORMAT_RG16_SFLOAT);
    p.depthTexture = c.sampled(4);  // half-res R16F (halfResTrace) or the full-res D32F depth buffer
    p.normalTexture = c.sampled(5, PLAIN_FORMAT_RGBA8);
    p.g = c.g;
    p.tables = (const ShadingTables*)c.tables;
    p.filterIndex = c.spec<int>(0, 0);
    if (c.failed) return;
    const int fmt = c.sampledFormat(4);
    if ((int)c.exec->dispatch[0] * 8 < p.outYSH.w || (int)c.exec->dispatch[1] * 8 < p.outYSH.h) { c.fail("filterIndirectDiffuseSpatial.comp: dispatch does not cover the target"); return; }
    c.window(p.outYSH.h, p.y0, p.y1);
    if (p.y1 <= p.y0) return;
    dim3 grid(ceilDiv(p.outYSH.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8));
    const bool sameExtent = p.texYSH.w == p.depthTexture.w && p.texYSH.h == p.depthTexture.h && p.texCoCg.w == p.depthTexture.w && p.texCoCg.h == p.depthTexture.h;
    if (fmt == PLAIN_FORMAT_R16_SFLOAT && sameExtent) PLAIN_LAUNCH(c, (giSpatialFilterKernel<true, true>), grid, 256, 0, p);
    else if (fmt == PLAIN_FORMAT_R16_SFLOAT) PLAIN_LAUNCH(c, (giSpatialFilterKernel<true, false>), grid, 256, 0, p);
    else if (fmt == PLAIN_FORMAT_DEPTH32 && sameExtent) PLAIN_LAUNCH(c, (giSpatialFilterKernel<false, true>), grid, 256, 0, p);
    else if (fmt == PLAIN_FORMAT_DEPTH32) PLAIN_LAUNCH(c, (giSpatialFilterKernel<false, false>), grid, 256, 0, p);
    else c.fail("filterIndirectDiffuseSpatial.comp: depth binding must be R16F or D32F");
}

// ---------------- filterIndirectDiffuseTemporal.comp ----------------
struct TemporalParams {
    ImgView targetYSH, targetCoCg, historyOutYSH, historyOutCoCg, inputYSH, inputCoCg, historyInYSH, historyInCoCg, velocityCurrent, velocityLastFrame;
    const plain_global_shader_info* g;
    int y0, y1;  // rows to produce (row sharding)
};
__global__ void __launch_bounds__(256) giTemporalFilterKernel(const __grid_constant__ TemporalParams p) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.targetYSH.w || iy >= p.y1) return;
    const plain_global_shader_info* g = p.g;
    const vec2 texelSize = 1.f / v2((float)p.targetYSH.w, (float)p.targetYSH.h);
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) * texelSize;
    const vec4 current_Y_SH = sampleRGBA16FLinearClamp(p.inputYSH, uv);
    const vec2 current_CoCg = sampleRG16FLinearClamp(p.inputCoCg, uv);
    const vec2 motion = sampleLinear2D<WRAP_CLAMP, vec2>([&](int x, int y) { return loadRG16SNORM(p.velocityCurrent, x, y); }, p.velocityCurrent.w, p.velocityCurrent.h, uv, v2(0.f));
    const vec2 uvReprojected = uv + motion;
    vec4 history_Y_SH = sampleRGBA16FLinearClamp(p.historyInYSH, uvReprojected);
    vec2 history_CoCg = sampleRG16FLinearClamp(p.historyInCoCg, uvReprojected);
    const vec2 motionLastFrame = sampleLinear2D<WRAP_REPEAT, vec2>([&](int x, int y) { return loadRG16SNORM(p.velocityLastFrame, x, y); }, p.velocityLastFrame.w, p.velocityLastFrame.h, uvReprojected, v2(0.f));
    const float motionDifference = sqrtf_(absf(length(motion) - length(motionLastFrame)));
    const float K = 10.f;
    const float motionDifferenceFactor = clampf(motionDifference * K, 0.f, 1.f);
    const float alphaDefault = 0.8f;
    float alphaMin = 0.6f;
    alphaMin -= 0.3f * absf(length(current_Y_SH) - length(history_Y_SH));
    alphaMin = fmaxp(alphaMin, 0.f);
    float alpha = mixf(alphaDefault, alphaMin, motionDifferenceFactor);
    const float pixelThreshold = 3.f;
    const vec2 res = v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    const vec2 am = vabs(motion) * res, al = vabs(motionLastFrame) * res;
    if (am.x > pixelThreshold || am.y > pixelThreshold || al.x > pixelThreshold || al.y > pixelThreshold) alpha = alphaMin;
    if (uvReprojected.x < 0.f || uvReprojected.y < 0.f || uvReprojected.x > 1.f || uvReprojected.y > 1.f) alpha = 0.f;
    if (g->cameraCut) alpha = 0.f;
    if (anynan(current_Y_SH) || anynan(current_CoCg)) {
        alpha = 1.f;
        if (anynan(history_Y_SH)) history_Y_SH = v4(0.f);
        if (anynan(history_CoCg)) history_CoCg = v2(0.f);
    }
    const vec4 result_Y_SH = vmix(current_Y_SH, history_Y_SH, alpha);
    const vec2 result_CoCg = vmix(current_CoCg, history_CoCg, alpha);
    storeRGBA16F(p.targetYSH, ix, iy, 0, result_Y_SH);
    if (inRange(p.targetCoCg, ix, iy)) storeRG16F(p.targetCoCg, ix, iy, result_CoCg);
    if (inRange(p.historyOutYSH, ix, iy)) storeRGBA16F(p.historyOutYSH, ix, iy, 0, result_Y_SH);
    if (inRange(p.historyOutCoCg, ix, iy)) storeRG16F(p.historyOutCoCg, ix, iy, result_CoCg);
}
PLAIN_PASS(launch_giTemporalFilter, "filterIndirectDiffuseTemporal.comp") {
    TemporalParams p;
    p.targetYSH = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.targetCoCg = c.storage(1, PLAIN_FORMAT_RG16_SFLOAT);
    p.historyOutYSH = c.storage(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.historyOutCoCg = c.storage(3, PLAIN_FORMAT_RG16_SFLOAT);
    p.inputYSH = c.sampled(4, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.inputCoCg = c.sampled(5, PLAIN_FORMAT_RG16_SFLOAT);
    p.historyInYSH = c.sampled(6, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.historyInCoCg = c.sampled(7, PLAIN_FORMAT_RG16_SFLOAT);
    p.velocityCurrent = c.sampled(8, PLAIN_FORMAT_RG16_SNORM);
    p.velocityLastFrame = c.sampled(9, PLAIN_FORMAT_RG16_SNORM);
    p.g = c.g;
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < p.targetYSH.w || (int)c.exec->dispatch[1] * 8 < p.targetYSH.h) { c.fail("filterIndirectDiffuseTemporal.comp: dispatch does not cover the target"); return; }
    c.window(p.targetYSH.h, p.y0, p.y1);
    if (p.y1 <= p.y0) return;
    PLAIN_LAUNCH(c, giTemporalFilterKernel, dim3(ceilDiv(p.targetYSH.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8)), 256, 0, p);
}

// ---------------- indirectLightUpscale.comp ----------------
struct UpscaleParams {
    ImgView dstYSH, dstCoCg;
    UpscaleSource src;
    const plain_global_shader_info* g;
    int y0, y1;  // rows to produce (row sharding)
};
__global__ void __launch_bounds__(256) giUpscaleKernel(const __grid_constant__ UpscaleParams p) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ix >= p.dstYSH.w || iy >= p.y1) return;
    vec4 result_Y_SH;
    vec2 result_CoCg;
    giUpscalePixel(p.src, p.g, ix, iy, result_Y_SH, result_CoCg);
    storeRGBA16F(p.dstYSH, ix, iy, 0, result_Y_SH);
    if (inRange(p.dstCoCg, ix, iy)) storeRG16F(p.dstCoCg, ix, iy, result_CoCg);
}
PLAIN_PASS(launch_giUpscale, "indirectLightUpscale.comp") {
    UpscaleParams p;
    p.dstYSH = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.dstCoCg = c.storage(1, PLAIN_FORMAT_RG16_SFLOAT);
    p.src.srcYSH = c.sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.src.srcCoCg = c.sampled(3, PLAIN_FORMAT_RG16_SFLOAT);
    p.src.fullResDepth = c.sampled(4, PLAIN_FORMAT_DEPTH32);
    p.src.halfResDepth = c.sampled(5, PLAIN_FORMAT_R16_SFLOAT);
    p.g = c.g;
    if (c.failed) return;
    if ((int)c.exec->dispatch[0] * 8 < p.dstYSH.w || (int)c.exec->dispatch[1] * 8 < p.dstYSH.h) { c.fail("indirectLightUpscale.comp: dispatch does not cover the target"); return; }
    c.window(p.dstYSH.h, p.y0, p.y1);
    if (p.y1 <= p.y0) return;
    PLAIN_LAUNCH(c, giUpscaleKernel, dim3(ceilDiv(p.dstYSH.w, 32), ceilDiv((unsigned)(p.y1 - p.y0), 8)), 256, 0, p);
}

}  // namespace pb
