// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_depth.cpp - min/max depth pyramid, sun light matrices, depth downscale.
#include "backend.h"
#include "shader_inc.h"

namespace orc {

// depthHiZPyramid.comp:52-124
static vec2 computeMinMax(const View& src, ivec2 upperLeft, bool fromDepthBuffer, ivec2 srcRes, bool extraRow, bool extraColumn) {
    float depthMin = 1.f;
    float depthMax = 0.f;
    vec2 texelSize = vec2(1.f) / tovec2(srcRes);
    vec2 upperLeftUV = tovec2(upperLeft) * texelSize;
    upperLeftUV += texelSize * 0.5f;
    auto tap = [&](vec2 offset, bool cornerQuirk) {
        vec2 uv = upperLeftUV + offset * texelSize;
        vec4 t = texture(src, s_nearestClamp, uv);
        if (fromDepthBuffer) {
            float depthTexel = t.x;
            float isSky = (depthTexel == 0.f) ? 1.f : 0.f;
            // :114 multiplies instead of adding for the odd x odd corner texel
            depthMin = min(depthMin, cornerQuirk ? depthTexel * isSky : depthTexel + isSky);
            depthMax = max(depthMax, depthTexel);
        } else {
            depthMin = min(depthMin, t.x + ((t.y == 0.f) ? 1.f : 0.f));
            depthMax = max(depthMax, t.y);
        }
    };
    tap(vec2(0, 0), false); tap(vec2(1, 0), false); tap(vec2(0, 1), false); tap(vec2(1, 1), false);
    if (extraRow) { tap(vec2(0, 2), false); tap(vec2(1, 2), false); }
    if (extraColumn) { tap(vec2(2, 0), false); tap(vec2(2, 1), false); }
    if (extraRow && extraColumn) tap(vec2(2, 2), true);
    return vec2(depthMin, depthMax);
}

// depthHiZPyramid.comp:130-350. The reference computes mips 0-5 per workgroup and the rest in the last workgroup to
// finish; a level whose source size is odd reads texels owned by a neighbouring workgroup (extra row/column), which
// races in the reference. The oracle defines the result level by level: every texel of a level is computed from the
// complete previous level.
ORACLE_PASS(pass_depthHiZPyramid, "depthHiZPyramid.comp") {
    const int mipCount = c.spec<int>(0, 0);
    const int depthBufferResX = c.spec<int>(1, 0);
    const int depthBufferResY = c.spec<int>(2, 0);
    View depthBuffer = c.sampled(13);
    View pyramidTexture = c.sampled(15);
    uint32_t* syncCounter = (uint32_t*)c.sbuf(16);
    bool fromDepthBuffer = true;
    int srcMipLevel = 0;
    ivec2 srcMipRes(depthBufferResX, depthBufferResY);
    ivec2 currentMipRes(max(srcMipRes.x / 2, 1), max(srcMipRes.y / 2, 1));
    // 11 bindings in the reference (:21-31); a 12th level at binding 11 is the extension for 7680x4320
    const int bindingCount = mipCount > 11 ? mipCount : 11;
    for (int k = 0; k < bindingCount; k++) {
        if (!(mipCount >= bindingCount - k)) continue;
        View target = c.storage((uint32_t)k);
        View src = depthBuffer;
        if (!fromDepthBuffer) { src = pyramidTexture; src.mip = pyramidTexture.mip + srcMipLevel; }
        const bool extraRow = (srcMipRes.y % 2) == 1, extraColumn = (srcMipRes.x % 2) == 1;
        const ivec2 sres = srcMipRes, cres = currentMipRes;
        const bool fdb = fromDepthBuffer;
        parallelFor(c.ctx->threads, cres.y, [&](int y) {
            for (int x = 0; x < cres.x; x++) {
                vec2 minMax = computeMinMax(src, ivec2(x * 2, y * 2), fdb, sres, extraRow, extraColumn);
                target.store(x, y, 0, vec4(minMax.x, minMax.y, 0.f, 0.f));
            }
        });
        if (!fromDepthBuffer) srcMipLevel++;
        fromDepthBuffer = false;
        srcMipRes = currentMipRes;
        currentMipRes = ivec2(max(srcMipRes.x / 2, 1), max(srcMipRes.y / 2, 1));
    }
    if (syncCounter) *syncCounter = 0;  // :261
}

// ---------------- lightMatrix.comp:57-138 ----------------
ORACLE_PASS(pass_lightMatrix, "lightMatrix.comp") {
    const uint32_t sunShadowCascadeCount = c.spec<uint32_t>(0, 4);
    plain_shadow_cascade_info* info = (plain_shadow_cascade_info*)c.sbuf(0);
    View depthMinMaxLowestMip = c.storage(1);
    const float highestCascadeExtraPadding = c.push<float>(0);
    const float highestCascadeMinFarPlane = c.push<float>(4);
    const plain_global_shader_info& g = c.g;
    const float FLOAT_MAX = 3.402823466e+38f, FLOAT_MIN = 1.175494351e-38f;
    vec3 camPos = c.gv3(g.cameraPosition), camFwd = c.gv3(g.cameraForward), camUp = c.gv3(g.cameraUp), camRight = c.gv3(g.cameraRight);

    mat4 coordinateSystemCorrection;
    coordinateSystemCorrection.c[0] = vec4(1.0f, 0.0f, 0.0f, 0.0f);
    coordinateSystemCorrection.c[1] = vec4(0.0f, 1.0f, 0.0f, 0.0f);
    coordinateSystemCorrection.c[2] = vec4(0.0f, 0.0f, -0.5f, 0.f);
    coordinateSystemCorrection.c[3] = vec4(0.0f, 0.0f, 0.5f, 1.0f);

    mat4 V = mat4_diag(1.f);
    vec3 forward = -c.gv3(g.sunDirection);
    vec3 up = abs(forward.y) < 0.9999f ? vec3(0.f, -1.f, 0.f) : vec3(0.f, 0.f, -1.f);
    vec3 right = cross(forward, up);
    up = cross(right, forward);
    vec3 rn = normalize(right), un = normalize(up);
    V.c[0] = vec4(rn, V.c[0].w);
    V.c[1] = vec4(un, V.c[1].w);
    V.c[2] = vec4(forward, V.c[2].w);
    V.c[3].w = 1.f;
    V = transpose(V);

    vec4 depthMinMax = depthMinMaxLowestMip.fetch(0, 0);
    float depthMaxLinear = linearizeDepth(depthMinMax.x, g.nearPlane, g.farPlane);
    float depthMinLinear = linearizeDepth(depthMinMax.y, g.nearPlane, g.farPlane);

    for (uint32_t i = 0; i + 1 < sunShadowCascadeCount; i++)
        info->splits[i] = depthMinLinear + ((depthMaxLinear - depthMinLinear) * (float)((int)i + 1) / (float)sunShadowCascadeCount);  // computeCascadeSplit :51-53

    for (uint32_t i = 0; i < sunShadowCascadeCount; i++) {
        vec3 minP = vec3(FLOAT_MAX);
        vec3 maxP = vec3(FLOAT_MIN);
        float cascadeMinDepth = (i == 0) ? depthMinLinear : info->splits[i - 1];
        float cascadeMaxDepth = info->splits[i];
        if (i == 0) cascadeMinDepth = depthMinLinear;
        if (i == sunShadowCascadeCount - 1) {
            cascadeMinDepth = g.nearPlane;
            cascadeMaxDepth = max(depthMaxLinear, highestCascadeMinFarPlane);
        }
        // computeFrustumPoints :29-48
        vec3 frustumPoints[8];
        {
            float near = cascadeMinDepth, far = cascadeMaxDepth;
            vec3 nearPlaneCenter = camPos + camFwd * near;
            vec3 farPlaneCenter = camPos + camFwd * far;
            float heightNear = g.cameraTanFovHalf * near;
            float heightFar = g.cameraTanFovHalf * far;
            float widthNear = heightNear * g.cameraAspectRatio;
            float widthFar = heightFar * g.cameraAspectRatio;
            frustumPoints[0] = farPlaneCenter + camUp * heightFar + camRight * widthFar;
            frustumPoints[1] = farPlaneCenter + camUp * heightFar - camRight * widthFar;
            frustumPoints[2] = farPlaneCenter - camUp * heightFar + camRight * widthFar;
            frustumPoints[3] = farPlaneCenter - camUp * heightFar - camRight * widthFar;
            frustumPoints[4] = nearPlaneCenter + camUp * heightNear + camRight * widthNear;
            frustumPoints[5] = nearPlaneCenter + camUp * heightNear - camRight * widthNear;
            frustumPoints[6] = nearPlaneCenter - camUp * heightNear + camRight * widthNear;
            frustumPoints[7] = nearPlaneCenter - camUp * heightNear - camRight * widthNear;
        }
        for (int k = 0; k < 8; k++) {
            vec3 pTransformed = (V * vec4(frustumPoints[k], 1.f)).xyz();
            minP = min(minP, pTransformed);
            maxP = max(maxP, pTransformed);
        }
        if (i == sunShadowCascadeCount - 1) {
            minP -= highestCascadeExtraPadding;
            maxP += highestCascadeExtraPadding;
        }
        minP -= shadowSampleRadius * 2.f;
        maxP += shadowSampleRadius * 2.f;
        vec3 scale = vec3(2.f) / (maxP - minP);
        vec3 offset = -0.5f * (maxP + minP) * scale;
        mat4 P;
        P.c[0] = vec4(scale.x, 0, 0, 0);
        P.c[1] = vec4(0, scale.y, 0, 0);
        P.c[2] = vec4(0, 0, scale.z, 0);
        P.c[3] = vec4(offset.x, offset.y, offset.z, 1.f);
        mat4 lm = coordinateSystemCorrection * P * V;
        for (int cc = 0; cc < 4; cc++)
            for (int r = 0; r < 4; r++) info->lightMatrices[i][cc * 4 + r] = lm.c[cc][r];
        info->lightSpaceScale[i][0] = scale.x;
        info->lightSpaceScale[i][1] = scale.y;
    }
}

// ---------------- depthDownscale.comp:12-20 ----------------
ORACLE_PASS(pass_depthDownscale, "depthDownscale.comp") {
    View halfResDst = c.storage(0);
    View fullResSrc = c.sampled(1);
    c.forEachInvocation(8, 8, 1, [&](int x, int y, int) {
        vec2 texelSize = 1.f / tovec2(textureSize(fullResSrc));
        vec2 uv = (vec2((float)(x * 2), (float)(y * 2)) + 0.5f) * texelSize;
        float depth = texture(fullResSrc, s_nearestClamp, uv).x;
        halfResDst.store(x, y, 0, vec4(depth, 0.f, 0.f, 0.f));
    });
}

}  // namespace orc
