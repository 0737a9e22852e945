// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_gi.cpp - SDF instance culling, diffuse SDF trace, spatial/temporal denoise, upscale.
#include "backend.h"
#include "shader_inc.h"

namespace orc {

// sdfCulling.inc:17-20
static uint32_t tileIndexFromTileUV(ivec2 tileUV, const plain_global_shader_info& g) {
    uint32_t tileCountX = f2uint(dm::floor_((float)g.screenResolution[0] / 32.f) + ((dm::floor_((float)g.screenResolution[0] / 32.f) < (float)g.screenResolution[0] / 32.f) ? 1.f : 0.f));  // ceil
    return (uint32_t)tileUV.x + (uint32_t)tileUV.y * tileCountX;
}

// ---------------- sdfCameraFrustumCulling.comp:36-62 ----------------
// The reference appends with atomicAdd, so the order of the list is not defined; the oracle (and the CUDA kernel)
// emit instance indices in ascending order.
ORACLE_PASS(pass_sdfCameraFrustumCulling, "sdfCameraFrustumCulling.comp") {
    const uint32_t instanceCount = *(const uint32_t*)c.sbuf(0);
    plain_camera_frustum_buffer fr;
    memcpy(&fr, c.ubuf(1), sizeof(fr));
    uint32_t* culled = (uint32_t*)c.sbuf(2);
    const plain_bounding_box* instanceBBs = (const plain_bounding_box*)c.sbuf(3);
    const float influenceRange = *(const float*)c.ubuf(4);
    const uint32_t invocations = c.exec->dispatch[0] * 64;
    for (uint32_t instanceIndex = 0; instanceIndex < invocations; instanceIndex++) {
        if (instanceIndex >= instanceCount) continue;
        const plain_bounding_box& bb = instanceBBs[instanceIndex];
        vec3 bbMin(bb.bbMin[0], bb.bbMin[1], bb.bbMin[2]), bbMax(bb.bbMax[0], bb.bbMax[1], bb.bbMax[2]);
        vec3 boundingSphereCenter = (bbMax + bbMin) * 0.5f;
        vec3 bbExtends = (bbMax - bbMin);
        float boundingSphereRadius = max(max(bbExtends.x, bbExtends.y), bbExtends.z) * 0.5f;
        boundingSphereRadius += influenceRange;
        bool isInsideFrustum = true;
        for (int i = 0; i < 6; i++) {
            vec3 frustumPoint(fr.frustumPoints[i][0], fr.frustumPoints[i][1], fr.frustumPoints[i][2]);
            vec3 frustumNormal(fr.frustumNormals[i][0], fr.frustumNormals[i][1], fr.frustumNormals[i][2]);
            bool isOutsidePlane = dot(boundingSphereCenter - frustumPoint, frustumNormal) > boundingSphereRadius;
            isInsideFrustum = isInsideFrustum && !isOutsidePlane;
        }
        if (isInsideFrustum) {
            uint32_t indexBufferIndex = culled[0]++;
            culled[1 + indexBufferIndex] = instanceIndex;
        }
    }
}

// ---------------- sdfCameraTileCulling.comp:37-99 ----------------
ORACLE_PASS(pass_sdfCameraTileCulling, "sdfCameraTileCulling.comp") {
    const bool useHiZ = c.specBool(0, false);
    const uint32_t tileCountX = c.push<uint32_t>(0), tileCountY = c.push<uint32_t>(4);
    const uint32_t* culled = (const uint32_t*)c.sbuf(0);
    const plain_bounding_box* instanceBBs = (const plain_bounding_box*)c.sbuf(1);
    size_t tilesSize = 0;
    plain_culled_instances_per_tile* cullingTiles = (plain_culled_instances_per_tile*)c.sbuf(2, &tilesSize);
    const float influenceRange = *(const float*)c.ubuf(3);
    View depthMinMaxTexture = c.sampled(4);
    const plain_global_shader_info& g = c.g;
    vec3 fwd = c.gv3(g.cameraForward), up = c.gv3(g.cameraUp), right = c.gv3(g.cameraRight), camPos = c.gv3(g.cameraPosition);
    auto VFromiUV = [&](ivec2 iUV) {  // :37-40, normalises by the full screen resolution (reference quirk)
        vec2 pixelCoor = (tovec2(iUV) / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]) - 0.5f) * 2.f;
        return calculateViewDirectionFromPixel(pixelCoor, fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
    };
    c.forEachInvocation(8, 8, 1, [&](int tx, int ty, int) {
        ivec2 tileUV(tx, ty);
        if ((uint32_t)tx >= tileCountX || (uint32_t)ty >= tileCountY) return;
        uint32_t tileIndex = tileIndexFromTileUV(tileUV, g);
        if ((size_t)(tileIndex + 1) * sizeof(plain_culled_instances_per_tile) > tilesSize) return;  // out-of-bounds writes are dropped
        plain_culled_instances_per_tile& tile = cullingTiles[tileIndex];
        tile.objectCount = 0;
        const int cullingTileSize = 32;
        vec3 cameraToPixel = -VFromiUV(tileUV * cullingTileSize + ivec2(cullingTileSize, cullingTileSize) / 2);
        vec3 V_ll = -VFromiUV(tileUV * cullingTileSize);
        vec3 V_ur = -VFromiUV(tileUV * cullingTileSize + ivec2(cullingTileSize, cullingTileSize));
        V_ll /= dot(cameraToPixel, V_ll);
        V_ur /= dot(cameraToPixel, V_ur);
        float coneRadiusPerMeter = distance(V_ll, V_ur) * 0.5f;
        float depthMin = g.nearPlane;
        float depthMax = g.farPlane;
        vec2 uv = tovec2(tileUV) / vec2((float)tileCountX, (float)tileCountY);
        if (useHiZ) {
            vec4 depthMinMax = texture(depthMinMaxTexture, s_nearestClamp, uv);
            depthMin = linearizeDepth(depthMinMax.y, g.nearPlane, g.farPlane);
            depthMax = linearizeDepth(depthMinMax.x, g.nearPlane, g.farPlane);
        }
        depthMin *= dot(cameraToPixel, fwd);
        depthMax *= dot(cameraToPixel, fwd);
        const uint32_t culledInstanceCount = culled[0];
        for (uint32_t i = 0; i < culledInstanceCount; i++) {
            const plain_bounding_box& bb = instanceBBs[culled[1 + i]];
            if (tile.objectCount >= PLAIN_MAX_OBJECTS_PER_TILE) break;
            vec3 bbMin(bb.bbMin[0], bb.bbMin[1], bb.bbMin[2]), bbMax(bb.bbMax[0], bb.bbMax[1], bb.bbMax[2]);
            vec3 boundingSphereCenter = (bbMax + bbMin) * 0.5f;
            vec3 bbExtends = (bbMax - bbMin) * 0.5f;
            float boundingSphereRadius = max(max(bbExtends.x, bbExtends.y), bbExtends.z);
            boundingSphereRadius += influenceRange;
            float projection = dot(boundingSphereCenter - camPos, cameraToPixel);
            projection = clamp(projection, depthMin, depthMax);
            float d = distance(boundingSphereCenter, projection * cameraToPixel + camPos);
            if (d < boundingSphereRadius + coneRadiusPerMeter * projection) {
                tile.indices[tile.objectCount] = culled[1 + i];
                tile.objectCount++;
            }
        }
    });
}

// ---------------- SDF.inc ----------------
struct TraceResult {
    bool hit;
    float closestHitDistance;
    vec3 hitPos, N;
    int hitCount;
    vec3 albedo;
};
static float sampleSDF(vec3 uv, const View& sdf) { return texture3D(sdf, s_linearClamp, uv).x; }  // SDF.inc:12-14
static vec3 normalFromSDF(vec3 uv, vec3 extends, const View& sdf) {  // SDF.inc:16-25
    float extendsMax = max(extends.x, max(extends.y, extends.z));
    vec3 extendsNormalized = extends / extendsMax;
    vec3 epsilon = vec3(0.15f) / vec3((float)sdf.w(), (float)sdf.h(), (float)sdf.d()) / extendsNormalized;
    return normalize(vec3(
        sampleSDF(uv + vec3(epsilon.x, 0, 0), sdf) - sampleSDF(uv - vec3(epsilon.x, 0, 0), sdf),
        sampleSDF(uv + vec3(0, epsilon.y, 0), sdf) - sampleSDF(uv - vec3(0, epsilon.y, 0), sdf),
        sampleSDF(uv + vec3(0, 0, epsilon.z), sdf) - sampleSDF(uv - vec3(0, 0, epsilon.z), sdf)));
}
static bool isPointInAABB(vec3 p, vec3 mn, vec3 mx) {  // SDF.inc:27-35
    return p.x >= mn.x && p.y >= mn.y && p.z >= mn.z && p.x <= mx.x && p.y <= mx.y && p.z <= mx.z;
}
struct HitResult { bool hit; float t; };
static HitResult rayAABBIntersection(vec3 rayOrigin, vec3 rayDirection, vec3 aabbMin, vec3 aabbMax) {  // SDF.inc:42-86
    HitResult result;
    result.hit = false;
    result.t = 100000.f;
    float intersection = rayOrigin.x < 0.f ? aabbMin.x : aabbMax.x;
    float tx = (intersection - rayOrigin.x) / rayDirection.x;
    vec3 planeIntersection = rayOrigin + tx * rayDirection;
    if (tx > 0.f && planeIntersection.y >= aabbMin.y && planeIntersection.y <= aabbMax.y && planeIntersection.z >= aabbMin.z && planeIntersection.z <= aabbMax.z) {
        result.t = min(result.t, tx);
        result.hit = true;
    }
    intersection = rayOrigin.y < 0.f ? aabbMin.y : aabbMax.y;
    float ty = (intersection - rayOrigin.y) / rayDirection.y;
    planeIntersection = rayOrigin + ty * rayDirection;
    if (ty > 0.f && planeIntersection.x >= aabbMin.x && planeIntersection.x <= aabbMax.x && planeIntersection.z >= aabbMin.z && planeIntersection.z <= aabbMax.z) {
        result.t = min(result.t, ty);
        result.hit = true;
    }
    intersection = rayOrigin.z < 0.f ? aabbMin.z : aabbMax.z;
    float tz = (intersection - rayOrigin.z) / rayDirection.z;
    planeIntersection = rayOrigin + tz * rayDirection;
    if (tz > 0.f && planeIntersection.x >= aabbMin.x && planeIntersection.x <= aabbMax.x && planeIntersection.y >= aabbMin.y && planeIntersection.y <= aabbMax.y) {
        result.t = min(result.t, tz);
        result.hit = true;
    }
    return result;
}
// SDF.inc:101-184
static void traceRayTroughSDFInstance(const plain_sdf_instance& inst, const mat4& worldToLocal, vec3 rayStartWorld, const View& sdf, vec3 rayDirectionWorld, TraceResult& tr) {
    vec3 localExtends(inst.localExtends[0], inst.localExtends[1], inst.localExtends[2]);
    vec3 rayStartLocal = (worldToLocal * vec4(rayStartWorld, 1.f)).xyz();
    vec3 rayEndLocal = (worldToLocal * vec4(rayStartWorld + rayDirectionWorld, 1.f)).xyz();
    vec3 rayDirection = rayEndLocal - rayStartLocal;
    rayDirection /= length(rayDirection);
    vec3 sdfMaxLocal = localExtends * 0.5f;
    vec3 sdfMinLocal = -sdfMaxLocal;
    float hitDistanceLocal = 0.f;
    if (!isPointInAABB(rayStartLocal, sdfMinLocal, sdfMaxLocal)) {
        HitResult aabbHit = rayAABBIntersection(rayStartLocal, rayDirection, sdfMinLocal, sdfMaxLocal);
        if (aabbHit.hit) {
            rayStartLocal += (aabbHit.t) * rayDirection;
            hitDistanceLocal = aabbHit.t;
        } else {
            return;
        }
    }
    vec3 localSamplePos = rayStartLocal;
    vec3 sdfResolution((float)sdf.w(), (float)sdf.h(), (float)sdf.d());
    float distanceThreshold = length(localExtends / sdfResolution) * 0.25f;
    float dLast = 0.f;
    float d = 0.f;
    float localToGlobalScale = 1.f / length(worldToLocal.c[0].xyz());
    if (localToGlobalScale * hitDistanceLocal > tr.closestHitDistance) return;
    for (int i = 0; i < 128; i++) {
        vec3 localExtendsHalf = localExtends * 0.5f;
        localExtendsHalf += 0.01f;
        if (localSamplePos.x > localExtendsHalf.x || localSamplePos.y > localExtendsHalf.y || localSamplePos.z > localExtendsHalf.z ||
            localSamplePos.x < -localExtendsHalf.x || localSamplePos.y < -localExtendsHalf.y || localSamplePos.z < -localExtendsHalf.z)
            break;
        vec3 sampleUV = localSamplePos / localExtends + 0.5f;
        dLast = d;
        d = texture3D(sdf, s_linearClamp, sampleUV).x;
        if (d < distanceThreshold) {
            tr.hit = true;
            float distanceGlobal = hitDistanceLocal * localToGlobalScale;
            if (distanceGlobal < tr.closestHitDistance) {
                tr.closestHitDistance = distanceGlobal;
                tr.hitCount = i;
                float lastStepSizeLocal = d / (1.f - (d - dLast));
                localSamplePos += rayDirection * lastStepSizeLocal;
                sampleUV = localSamplePos / localExtends + 0.5f;
                tr.N = normalFromSDF(sampleUV, localExtends, sdf);
                tr.N = transpose(mat3_from(worldToLocal)) * tr.N;
                tr.albedo = pow(vec3(inst.meanAlbedo[0], inst.meanAlbedo[1], inst.meanAlbedo[2]), vec3(2.2f));
                float lastStepSizeGlobal = lastStepSizeLocal * localToGlobalScale;
                tr.hitPos = rayStartWorld + rayDirectionWorld * (distanceGlobal + lastStepSizeGlobal);
            }
            break;
        }
        localSamplePos += rayDirection * abs(d);
        hitDistanceLocal += abs(d);
    }
}

// ---------------- sdfDiffuseTrace.comp:70-207 ----------------
ORACLE_PASS(pass_sdfDiffuseTrace, "sdfDiffuseTrace.comp") {
    const bool strictInfluenceRadiusCutoff = c.specBool(0, false);
    const int shadowCascadeIndex = c.spec<int>(1, 3);
    View imageOut_Y_SH = c.storage(0), imageOut_CoCg = c.storage(1);
    View depthTexture = c.sampled(2), normalTexture = c.sampled(3), skyLut = c.sampled(4), shadowMap = c.sampled(10);
    plain_light_buffer light;
    memcpy(&light, c.sbuf(5), sizeof(light));
    const uint8_t* instanceBuffer = c.sbuf(6);
    const plain_sdf_instance* sdfInstances = (const plain_sdf_instance*)(instanceBuffer + 16);
    size_t tilesSize = 0;
    const plain_culled_instances_per_tile* cameraCulledTiles = (const plain_culled_instances_per_tile*)c.sbuf(7, &tilesSize);
    const float influenceRange = *(const float*)c.ubuf(8);
    plain_shadow_cascade_info cascades;
    memcpy(&cascades, c.sbuf(9), sizeof(cascades));
    const plain_global_shader_info& g = c.g;
    View noiseTex = c.bindless((uint32_t)g.noiseTextureIndices[g.frameIndexMod4]);
    vec3 fwd = c.gv3(g.cameraForward), up = c.gv3(g.cameraUp), right = c.gv3(g.cameraRight), camPos = c.gv3(g.cameraPosition);
    const mat4 shadowMatrix = c.gm4(cascades.lightMatrices[shadowCascadeIndex]);

    c.forEachGroup([&](int gx, int gy, int) {
        struct RayInfo { vec3 normal; float depth; vec3 color; };
        RayInfo sharedRays[8][8];
        vec3 rayL[8][8];
        // tileUV = gl_WorkGroupID.xy / (cullingTileSize / 8), :154
        ivec2 tileUV(gx / 4, gy / 4);
        uint32_t tileIndex = tileIndexFromTileUV(tileUV, g);
        plain_culled_instances_per_tile cullingTile;
        memset(&cullingTile, 0, sizeof(cullingTile));
        if ((size_t)(tileIndex + 1) * sizeof(cullingTile) <= tilesSize) cullingTile = cameraCulledTiles[tileIndex];
        for (int ly = 0; ly < 8; ly++)
            for (int lx = 0; lx < 8; lx++) {
                ivec2 iUV(gx * 8 + lx, gy * 8 + ly);
                vec2 uv = tovec2(iUV) / vec2((float)imageOut_Y_SH.w(), (float)imageOut_Y_SH.h());
                float depth = texture(depthTexture, s_nearestClamp, uv).x;
                float depthLinear = linearizeDepth(depth, g.nearPlane, g.farPlane);
                vec2 pixelNDC = uv * 2.f - 1.f;
                vec3 V = -calculateViewDirectionFromPixel(pixelNDC, fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
                vec3 pWorld = camPos + V / dot(V, fwd) * depthLinear;
                vec2 noiseUV = tovec2(iUV) / tovec2(textureSize(noiseTex));
                vec2 xi = texture(noiseTex, s_nearestRepeat, noiseUV).xy();
                vec3 normalTexel = texture(normalTexture, s_nearestClamp, uv).xyz();
                vec3 N = normalTexel * 2.f - 1.f;
                sharedRays[lx][ly].normal = N;
                sharedRays[lx][ly].depth = depthLinear;
                vec3 rayOrigin = pWorld + N * 0.2f;
                vec3 L = importanceSampleCosine(xi, N);
                rayL[lx][ly] = L;

                TraceResult traceResult;
                traceResult.hit = false;
                traceResult.closestHitDistance = 10000.f;
                traceResult.hitCount = 0;
                for (uint32_t i = 0; i < cullingTile.objectCount; i++) {
                    const plain_sdf_instance& instance = sdfInstances[cullingTile.indices[i]];
                    traceRayTroughSDFInstance(instance, c.gm4(instance.worldToLocal), rayOrigin, c.bindless(instance.sdfTextureIndex), L, traceResult);
                }
                vec3 hitColor;
                if (traceResult.hit) {
                    float shadow = simpleShadow(traceResult.hitPos, shadowMatrix, shadowMap, s_nearestWhiteBorder);
                    vec3 sunLight = shadow * light.sunStrengthExposed * vec3(light.sunColor[0], light.sunColor[1], light.sunColor[2]);
                    hitColor = traceResult.albedo * sunLight;
                    bool hitInRange = traceResult.closestHitDistance < influenceRange;
                    hitInRange = hitInRange || !strictInfluenceRadiusCutoff;
                    bool selfIntersection = traceResult.closestHitDistance < 0.0001f;
                    if (!hitInRange || selfIntersection) hitColor = vec3(0.f);
                } else {
                    hitColor = sampleSkyLut(L, skyLut);
                }
                sharedRays[lx][ly].color = hitColor;
            }
        // resolveColor :70-116 (after the barrier)
        for (int ly = 0; ly < 8; ly++)
            for (int lx = 0; lx < 8; lx++) {
                float weightTotal = 1.f;
                vec3 color = sharedRays[lx][ly].color;
                for (int x = -1; x <= 1; x++) {
                    for (int y = -1; y <= 1; y++) {
                        if (x == 0 && y == 0) continue;
                        ivec2 rayIndex(lx + x, ly + y);
                        // greaterThan(rayIndex, 0): row/column 0 of the group is never used as a neighbour (:88)
                        bool isValidIndex = rayIndex.x > 0 && rayIndex.y > 0 && rayIndex.x < 8 && rayIndex.y < 8;
                        if (!isValidIndex) continue;
                        const RayInfo& neighbourRay = sharedRays[rayIndex.x][rayIndex.y];
                        float normalThreshold = 0.9f;
                        float NoN = clamp(dot(sharedRays[lx][ly].normal, neighbourRay.normal), 0.f, 1.f);
                        bool normalsMatch = NoN > normalThreshold;
                        float depthThreshold = 0.5f;
                        bool depthMatch = abs(sharedRays[lx][ly].depth - neighbourRay.depth) < depthThreshold;
                        if (normalsMatch && depthMatch) {
                            float weightX = x == 0 ? 1.f : 0.5f;
                            float weightY = y == 0 ? 1.f : 0.5f;
                            float weight = weightX * weightY;
                            color += weight * neighbourRay.color;
                            weightTotal += weight;
                        }
                    }
                }
                color /= weightTotal;
                vec3 YCoCg = linearToYCoCg(color);
                vec4 result_Y_SH = vec4(0.f);
                vec2 result_CoCg = vec2(0.f);
                result_Y_SH += YCoCg.x * directionToSH_L1(rayL[lx][ly]);
                result_CoCg += vec2(YCoCg.y, YCoCg.z);
                imageOut_Y_SH.store(gx * 8 + lx, gy * 8 + ly, 0, result_Y_SH);
                imageOut_CoCg.store(gx * 8 + lx, gy * 8 + ly, 0, vec4(result_CoCg.x, result_CoCg.y, 0.f, 0.f));
            }
    });
}

// ---------------- sdfDebugVisualisation.comp:74-133 ----------------
ORACLE_PASS(pass_sdfDebugVisualisation, "sdfDebugVisualisation.comp") {
    const int debugMode = c.spec<int>(0, 0);
    const int shadowCascadeIndex = c.spec<int>(1, 3);
    View imageOut = c.storage(0), skyLut = c.sampled(2), shadowMap = c.sampled(7);
    plain_light_buffer light;
    memcpy(&light, c.sbuf(1), sizeof(light));
    const plain_sdf_instance* sdfInstances = (const plain_sdf_instance*)(c.sbuf(3) + 16);
    size_t tilesSize = 0;
    const plain_culled_instances_per_tile* cameraCulledTiles = (const plain_culled_instances_per_tile*)c.sbuf(4, &tilesSize);
    plain_shadow_cascade_info cascades;
    memcpy(&cascades, c.sbuf(6), sizeof(cascades));
    const plain_global_shader_info& g = c.g;
    vec3 fwd = c.gv3(g.cameraForward), up = c.gv3(g.cameraUp), right = c.gv3(g.cameraRight), camPos = c.gv3(g.cameraPosition);
    const mat4 shadowMatrix = c.gm4(cascades.lightMatrices[shadowCascadeIndex]);
    const vec3 sunDirection = c.gv3(g.sunDirection);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= imageOut.w() || iy >= imageOut.h()) return;  // stores outside the image are dropped
        vec2 pixelCoor = (vec2((float)ix, (float)iy) / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]) - 0.5f) * 2.f;
        vec3 cameraToPixel = -calculateViewDirectionFromPixel(pixelCoor, fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
        ivec2 tileUV(ix / 32, iy / 32);
        uint32_t tileIndex = tileIndexFromTileUV(tileUV, g);
        vec3 rayStart = camPos + g.nearPlane * cameraToPixel;
        TraceResult traceResult;
        traceResult.hit = false;
        traceResult.closestHitDistance = 10000.f;
        traceResult.hitCount = 0;  // hitPos / N / albedo / hitCount are undefined in the reference until a hit; pinned to 0
        plain_culled_instances_per_tile cullingTile;
        memset(&cullingTile, 0, sizeof(cullingTile));
        if ((size_t)(tileIndex + 1) * sizeof(cullingTile) <= tilesSize) cullingTile = cameraCulledTiles[tileIndex];
        for (uint32_t i = 0; i < cullingTile.objectCount; i++) {
            const plain_sdf_instance& instance = sdfInstances[cullingTile.indices[i]];
            traceRayTroughSDFInstance(instance, c.gm4(instance.worldToLocal), rayStart, c.bindless(instance.sdfTextureIndex), cameraToPixel, traceResult);
        }
        float shadow = simpleShadow(traceResult.hitPos, shadowMatrix, shadowMap, s_nearestBlackBorder);
        vec3 color = vec3(0.f);
        if (traceResult.hit || debugMode == 2) {
            if (debugMode == 1) {
                vec3 sunLight = light.sunStrengthExposed * vec3(light.sunColor[0], light.sunColor[1], light.sunColor[2]);
                sunLight = sunLight * shadow;
                vec3 ambient = vec3(0.15f);
                float NoL = clamp(dot(traceResult.N, sunDirection), 0.f, 1.f);
                color = traceResult.albedo * (ambient + sunLight * NoL);
            } else if (debugMode == 2) {
                float percentage = (float)cullingTile.objectCount / (float)PLAIN_MAX_OBJECTS_PER_TILE;
                color = percentage >= 1.f ? vec3(1.f, 0.f, 0.f) : vec3(percentage);
            } else if (debugMode == 3) {
                color = traceResult.N * 0.5f + 0.5f;
            } else if (debugMode == 4) {
                color = vec3((float)traceResult.hitCount / 128.f);
            }
        } else {
            color = sampleSkyLut(cameraToPixel, skyLut);
        }
        imageOut.store(ix, iy, 0, vec4(color, 1.f));
    });
}

// ---------------- filterIndirectDiffuseSpatial.comp:21-135 ----------------
ORACLE_PASS(pass_filterIndirectDiffuseSpatial, "filterIndirectDiffuseSpatial.comp") {
    const int filterIndex = c.spec<int>(0, 0);
    View imageOut_Y_SH = c.storage(0), imageOut_CoCg = c.storage(1);
    View texture_Y_SH = c.sampled(2), texture_CoCg = c.sampled(3), depthTexture = c.sampled(4), normalTexture = c.sampled(5);
    const plain_global_shader_info& g = c.g;
    vec3 fwd = c.gv3(g.cameraForward), up = c.gv3(g.cameraUp), right = c.gv3(g.cameraRight), camPos = c.gv3(g.cameraPosition);
    const mat4 viewProjection = c.gm4(g.viewProjection);
    auto pixelToWorld = [&](vec2 uv) {
        float depth = texture(depthTexture, s_nearestClamp, uv).x;
        float depthLinear = linearizeDepth(depth, g.nearPlane, g.farPlane);
        vec2 pixelNDC = uv * 2.f - 1.f;
        vec3 cameraToPixel = -calculateViewDirectionFromPixel(pixelNDC, fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
        return camPos + cameraToPixel / dot(cameraToPixel, fwd) * depthLinear;
    };
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= imageOut_Y_SH.w() || iy >= imageOut_Y_SH.h()) return;  // stores would be dropped anyway
        vec2 texelSize = 1.f / vec2((float)imageOut_Y_SH.w(), (float)imageOut_Y_SH.h());
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) * texelSize;
        vec3 pCenter = pixelToWorld(uv);
        vec3 pRight = pixelToWorld(uv + vec2(1.f, 0.f) * texelSize);
        vec3 pUp = pixelToWorld(uv + vec2(0.f, 1.f) * texelSize);
        vec3 tangent = normalize(pCenter - pRight);
        vec3 bitangent = normalize(pCenter - pUp);
        vec3 N = 2.f * texture(normalTexture, s_nearestClamp, uv).xyz() - 1.f;
        int sampleCount = 32;
        vec4 result_Y_SH = vec4(0.f);
        vec2 result_CoCg = vec2(0.f);
        float weightTotal = 0.f;
        uint rngState = wang_hash(g.frameIndexMod4 + (uint)filterIndex);
        float radiusWorld = 1.5f;
        if (filterIndex == 1) radiusWorld = 1.f;
        float lengthModifier = 1.f;
        for (int i = 0; i < sampleCount; i++) {
            float d = sqrt(rand(rngState)) * lengthModifier;
            float angle = 2.f * pi * rand(rngState);
            vec2 offset = vec2(cos(angle), sin(angle)) * d;
            vec3 sampleWorld = pCenter + radiusWorld * (offset.x * tangent + offset.y * bitangent);
            vec4 sampleProjected = viewProjection * vec4(sampleWorld, 1.f);
            vec2 sampleUV = sampleProjected.xy() / sampleProjected.w;
            sampleUV = sampleUV * 0.5f + 0.5f;
            sampleUV.x = sampleUV.x < 0.f ? uv.x - offset.x : sampleUV.x;
            sampleUV.y = sampleUV.y < 0.f ? uv.y - offset.y : sampleUV.y;
            sampleUV.x = sampleUV.x > 1.f ? uv.x - offset.x : sampleUV.x;
            sampleUV.y = sampleUV.y > 1.f ? uv.y - offset.y : sampleUV.y;
            vec3 pixelWorld = pixelToWorld(sampleUV);
            float distanceToTangentPlane = abs(dot(N, pixelWorld - pCenter));
            float maxDistance = 0.25f;
            float weight = clamp(maxDistance / max(distanceToTangentPlane, 0.0001f), 0.f, 1.f);
            weight *= weight;
            if (sampleUV.x < 0.f || sampleUV.y < 0.f || sampleUV.x > 1.f || sampleUV.y > 1.f) {
                weight = 0.f;
                lengthModifier *= 0.98f;
            }
            if (weight > 0.f) {
                vec4 sample_Y_SH = texture(texture_Y_SH, s_nearestClamp, sampleUV);
                vec2 sample_CoCg = texture(texture_CoCg, s_nearestClamp, sampleUV).xy();
                if (any_isnan(sample_Y_SH) || any_isnan(sample_CoCg)) {
                } else {
                    result_Y_SH += weight * sample_Y_SH;
                    result_CoCg += weight * sample_CoCg;
                    weightTotal += weight;
                }
            }
        }
        weightTotal = max(weightTotal, 0.00001f);
        result_Y_SH /= weightTotal;
        result_CoCg /= weightTotal;
        imageOut_Y_SH.store(ix, iy, 0, result_Y_SH);
        imageOut_CoCg.store(ix, iy, 0, vec4(result_CoCg.x, result_CoCg.y, 0.f, 0.f));
    });
}

// ---------------- filterIndirectDiffuseTemporal.comp:20-86 ----------------
ORACLE_PASS(pass_filterIndirectDiffuseTemporal, "filterIndirectDiffuseTemporal.comp") {
    View targetOut_Y_SH = c.storage(0), targetOut_CoCg = c.storage(1), historyOut_Y_SH = c.storage(2), historyOut_CoCg = c.storage(3);
    View input_Y_SH = c.sampled(4), input_CoCg = c.sampled(5), historyIn_Y_SH = c.sampled(6), historyIn_CoCg = c.sampled(7);
    View velocityCurrent = c.sampled(8), velocityLastFrame = c.sampled(9);
    const plain_global_shader_info& g = c.g;
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= targetOut_Y_SH.w() || iy >= targetOut_Y_SH.h()) return;
        vec2 texelSize = 1.f / vec2((float)targetOut_Y_SH.w(), (float)targetOut_Y_SH.h());
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) * texelSize;
        vec4 current_Y_SH = texture(input_Y_SH, s_linearClamp, uv);
        vec2 current_CoCg = texture(input_CoCg, s_linearClamp, uv).xy();
        vec2 motion = texture(velocityCurrent, s_linearClamp, uv).xy();
        vec2 uvReprojected = uv + motion;
        vec4 history_Y_SH = texture(historyIn_Y_SH, s_linearClamp, uvReprojected);
        vec2 history_CoCg = texture(historyIn_CoCg, s_linearClamp, uvReprojected).xy();
        vec2 motionLastFrame = texture(velocityLastFrame, s_linearRepeat, uvReprojected).xy();
        float motionDifference = sqrt(abs(length(motion) - length(motionLastFrame)));
        float K = 10.f;
        float motionDifferenceFactor = clamp(motionDifference * K, 0.f, 1.f);
        float alphaDefault = 0.8f;
        float alphaMin = 0.6f;
        alphaMin -= 0.3f * abs(length(current_Y_SH) - length(history_Y_SH));
        alphaMin = max(alphaMin, 0.f);
        float alpha = mix(alphaDefault, alphaMin, motionDifferenceFactor);
        float pixelThreshold = 3.f;
        vec2 res((float)g.screenResolution[0], (float)g.screenResolution[1]);
        vec2 am = abs(motion) * res, al = abs(motionLastFrame) * res;
        if (am.x > pixelThreshold || am.y > pixelThreshold || al.x > pixelThreshold || al.y > pixelThreshold) alpha = alphaMin;
        if (uvReprojected.x < 0.f || uvReprojected.y < 0.f || uvReprojected.x > 1.f || uvReprojected.y > 1.f) alpha = 0.f;
        if (g.cameraCut) alpha = 0.f;
        if (any_isnan(current_Y_SH) || any_isnan(current_CoCg)) {
            alpha = 1.f;
            if (any_isnan(history_Y_SH)) history_Y_SH = vec4(0.f);
            if (any_isnan(history_CoCg)) history_CoCg = vec2(0.f);
        }
        vec4 result_Y_SH = mix(current_Y_SH, history_Y_SH, alpha);
        vec2 result_CoCg = mix(current_CoCg, history_CoCg, alpha);
        targetOut_Y_SH.store(ix, iy, 0, result_Y_SH);
        targetOut_CoCg.store(ix, iy, 0, vec4(result_CoCg.x, result_CoCg.y, 0.f, 0.f));
        historyOut_Y_SH.store(ix, iy, 0, result_Y_SH);
        historyOut_CoCg.store(ix, iy, 0, vec4(result_CoCg.x, result_CoCg.y, 0.f, 0.f));
    });
}

// ---------------- indirectLightUpscale.comp:17-71 ----------------
ORACLE_PASS(pass_indirectLightUpscale, "indirectLightUpscale.comp") {
    View fullResDst_Y_SH = c.storage(0), fullResDst_CoCg = c.storage(1);
    View halfResSrc_Y_SH = c.sampled(2), halfResSrc_CoCg = c.sampled(3), fullResDepthTex = c.sampled(4), halfResDepth = c.sampled(5);
    const plain_global_shader_info& g = c.g;
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= fullResDst_Y_SH.w() || iy >= fullResDst_Y_SH.h()) return;
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
        vec4 result_Y_SH = vec4(0.f);
        vec2 result_CoCg = vec2(0.f);
        float fullResDepth = texture(fullResDepthTex, s_nearestClamp, uv).x;
        fullResDepth = linearizeDepth(fullResDepth, g.nearPlane, g.farPlane);
        vec2 halfResTexelSize = 1.f / tovec2(textureSize(halfResDepth));
        vec4 depthSamples = textureGather(halfResDepth, s_nearestClamp, uv);
        depthSamples.x = linearizeDepth(depthSamples.x, g.nearPlane, g.farPlane);
        depthSamples.y = linearizeDepth(depthSamples.y, g.nearPlane, g.farPlane);
        depthSamples.z = linearizeDepth(depthSamples.z, g.nearPlane, g.farPlane);
        depthSamples.w = linearizeDepth(depthSamples.w, g.nearPlane, g.farPlane);
        float minDepthDiff = 1000.f;
        vec2 closestDepthTexel = vec2(0.f);
        float edgeDepthThreshold = 0.5f;
        bool isEdge = false;
        vec2 offsets[4] = {vec2(0, 1), vec2(1, 1), vec2(1, 0), vec2(0, 0)};
        for (int i = 0; i < 4; i++) {
            float depthDiff = abs(depthSamples[i] - fullResDepth);
            isEdge = isEdge || depthDiff > edgeDepthThreshold;
            if (depthDiff < minDepthDiff) {
                minDepthDiff = depthDiff;
                closestDepthTexel = offsets[i];
            }
        }
        vec2 uvClosestTexel = uv + closestDepthTexel * halfResTexelSize;
        if (isEdge) {
            result_Y_SH = texture(halfResSrc_Y_SH, s_nearestClamp, uvClosestTexel);
            result_CoCg = texture(halfResSrc_CoCg, s_nearestClamp, uvClosestTexel).xy();
        } else {
            result_Y_SH = texture(halfResSrc_Y_SH, s_linearClamp, uv);
            result_CoCg = texture(halfResSrc_CoCg, s_linearClamp, uv).xy();
        }
        fullResDst_Y_SH.store(ix, iy, 0, result_Y_SH);
        fullResDst_CoCg.store(ix, iy, 0, vec4(result_CoCg.x, result_CoCg.y, 0.f, 0.f));
    });
}

}  // namespace orc

// the functions above that restate the reference's GLSL include files, behind the batch entry points the reference-pinning test uses
#define INC_NS orc
#define INC_PREFIX oracle_inc_
#define INC_IS_REFERENCE 0
#define INC_PART_SDF 1
#include "inc_eval.h"
