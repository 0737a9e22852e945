// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_post.cpp - froxel volumetrics, TAA resolve, bloom chain.
#include "backend.h"
#include "shader_inc.h"

namespace orc {

static vec3 froxelWorldPos(const plain_global_shader_info& g, vec3 uv, float maxDistance, vec3* outV) {
    vec3 fwd(g.cameraForward[0], g.cameraForward[1], g.cameraForward[2]), up(g.cameraUp[0], g.cameraUp[1], g.cameraUp[2]);
    vec3 right(g.cameraRight[0], g.cameraRight[1], g.cameraRight[2]), camPos(g.cameraPosition[0], g.cameraPosition[1], g.cameraPosition[2]);
    vec3 ndc = 2.f * (uv - 0.5f);
    vec3 V = calculateViewDirectionFromPixel(vec2(ndc.x, ndc.y), fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
    if (outV) *outV = V;
    return camPos - V / dot(-V, fwd) * froxelUVToDepth(uv.z, maxDistance);
}

// ---------------- froxelVolumeMaterial.comp:17-44 ----------------
ORACLE_PASS(pass_froxelVolumeMaterial, "froxelVolumeMaterial.comp") {
    View materialVolume = c.storage(0);
    View noiseTexture = c.sampled(1);
    plain_volumetric_lighting_settings s;
    memcpy(&s, c.ubuf(2), sizeof(s));
    const plain_global_shader_info& g = c.g;
    c.forEachInvocation(4, 4, 4, [&](int x, int y, int z) {
        if (x >= materialVolume.w() || y >= materialVolume.h() || z >= materialVolume.d()) return;
        vec3 volumeRes((float)materialVolume.w(), (float)materialVolume.h(), (float)materialVolume.d());
        vec3 uv = (vec3((float)x, (float)y, (float)z) + 0.5f + s.sampleOffset) / volumeRes;
        vec3 posWorld = froxelWorldPos(g, uv, s.maxDistance, nullptr);
        float noiseScale = 0.5f;
        vec3 noiseSample = posWorld * noiseScale + vec3(s.windSampleOffset[0], s.windSampleOffset[1], s.windSampleOffset[2]);
        float noise = texture3D(noiseTexture, s_linearRepeat, noiseSample).x;
        vec3 scatteringCoefficient(s.scatteringCoefficients[0], s.scatteringCoefficients[1], s.scatteringCoefficients[2]);
        float absorptionCoefficient = s.absorptionCoefficient;
        float densityMultiplier = s.baseDensity;
        densityMultiplier += s.densityNoiseRange * (noise - 0.5f);
        densityMultiplier = max(densityMultiplier, 0.f);
        scatteringCoefficient *= densityMultiplier;
        absorptionCoefficient *= densityMultiplier;
        materialVolume.store(x, y, z, vec4(scatteringCoefficient, absorptionCoefficient));
    });
}

// ---------------- froxelLightScattering.comp:31-64 ----------------
ORACLE_PASS(pass_froxelLightScattering, "froxelLightScattering.comp") {
    View outVolume = c.storage(0);
    View sunShadowMap = c.sampled(1), materialVolume = c.sampled(2);
    plain_shadow_cascade_info cascades;
    memcpy(&cascades, c.sbuf(3), sizeof(cascades));
    plain_light_buffer light;
    memcpy(&light, c.sbuf(4), sizeof(light));
    plain_volumetric_lighting_settings s;
    memcpy(&s, c.ubuf(5), sizeof(s));
    const plain_global_shader_info& g = c.g;
    const mat4 lightMatrix = c.gm4(cascades.lightMatrices[2]);  // hard-coded cascade 2 (:45)
    c.forEachInvocation(4, 4, 4, [&](int x, int y, int z) {
        if (x >= outVolume.w() || y >= outVolume.h() || z >= outVolume.d()) return;
        vec3 volumeRes((float)outVolume.w(), (float)outVolume.h(), (float)outVolume.d());
        vec3 uv = (vec3((float)x, (float)y, (float)z) + 0.5f + s.sampleOffset) / volumeRes;
        // ndc = 2*uv - 1 here (:40), 2*(uv - 0.5) in the other froxel passes
        vec3 ndc = 2.f * uv - 1.f;
        vec3 fwd = c.gv3(g.cameraForward);
        vec3 V = calculateViewDirectionFromPixel(vec2(ndc.x, ndc.y), fwd, c.gv3(g.cameraUp), c.gv3(g.cameraRight), g.cameraTanFovHalf, g.cameraAspectRatio);
        vec3 posWorld = c.gv3(g.cameraPosition) - V / dot(-V, fwd) * froxelUVToDepth(uv.z, s.maxDistance);
        float shadow = simpleShadow(posWorld, lightMatrix, sunShadowMap, s_nearestBlackBorder);
        float sunStrength = shadow * light.sunStrengthExposed;
        vec3 L = c.gv3(g.sunDirection);
        float VoL = dot(-V, L);
        float phase = phaseGreenstein(VoL, s.phaseFunctionG);
        vec4 sa = materialVolume.fetch(x, y, z);
        vec3 scatteringCoefficient = sa.xyz();
        float absorptionCoefficient = sa.w;
        vec3 constantAmbientLighting = vec3(0.02f);
        vec3 inscattering = (sunStrength * phase * vec3(light.sunColor[0], light.sunColor[1], light.sunColor[2]) + constantAmbientLighting) * scatteringCoefficient;
        vec3 extinctionCoefficient = scatteringCoefficient + absorptionCoefficient;
        float transmittance = computeLuminance(extinctionCoefficient);
        outVolume.store(x, y, z, vec4(inscattering, transmittance));
    });
}

// ---------------- volumeLightingReprojection.comp:19-62 ----------------
ORACLE_PASS(pass_volumeLightingReprojection, "volumeLightingReprojection.comp") {
    View targetImage = c.storage(0);
    View inputVolume = c.sampled(1), historyVolume = c.sampled(2);
    plain_volumetric_lighting_settings s;
    memcpy(&s, c.ubuf(3), sizeof(s));
    const plain_global_shader_info& g = c.g;
    const mat4 viewProjectionPrevious = c.gm4(g.viewProjectionPrevious);
    c.forEachInvocation(4, 4, 4, [&](int x, int y, int z) {
        if (x >= targetImage.w() || y >= targetImage.h() || z >= targetImage.d()) return;
        vec4 current = inputVolume.fetch(x, y, z);
        vec3 volumeRes((float)targetImage.w(), (float)targetImage.h(), (float)targetImage.d());
        vec3 uv = (vec3((float)x, (float)y, (float)z) + 0.5f) / volumeRes;
        vec3 posWorld = froxelWorldPos(g, uv, s.maxDistance, nullptr);
        vec4 ndcPrevious = viewProjectionPrevious * vec4(posWorld, 1.f);
        vec3 ndcP = ndcPrevious.xyz() / ndcPrevious.w;
        vec3 camPosPrev = c.gv3(g.cameraPositionPrevious);
        vec3 V_history = normalize(camPosPrev - posWorld);
        float historyDistance = distance(posWorld, camPosPrev);
        float historyDepth = historyDistance * dot(-V_history, c.gv3(g.cameraForwardPrevious));
        vec3 historyUV = vec3(ndcP.x * 0.5f + 0.5f, ndcP.y * 0.5f + 0.5f, depthToFroxelUVZ(historyDepth, s.maxDistance));
        vec4 history = texture3D(historyVolume, s_linearClamp, historyUV);
        float alpha = 0.95f;
        if (historyUV.x > 1.f || historyUV.y > 1.f || historyUV.z > 1.f || historyUV.x < 0.f || historyUV.y < 0.f || historyUV.z < 0.f) alpha = 0.f;
        if (g.cameraCut) history = current;
        vec4 result = mix(current, history, alpha);
        targetImage.store(x, y, z, result);
    });
}

// ---------------- volumetricLightingIntegration.comp:18-43 ----------------
ORACLE_PASS(pass_volumetricLightingIntegration, "volumetricLightingIntegration.comp") {
    View integrationVolume = c.storage(0);
    View scatteringTransmittanceVolume = c.sampled(1);
    plain_volumetric_lighting_settings s;
    memcpy(&s, c.ubuf(2), sizeof(s));
    c.forEachInvocation(8, 8, 1, [&](int x, int y, int) {
        if (x >= integrationVolume.w() || y >= integrationVolume.h()) return;
        vec3 inscatteringTotal = vec3(0.f);
        float transmittance = 1.f;
        const int resZ = integrationVolume.d();
        // the reference loops z <= res.z (:28); iteration z == res.z fetches and stores out of range and has no effect
        for (int z = 0; z <= resZ; z++) {
            vec4 it = scatteringTransmittanceVolume.fetch(x, y, z);
            float depthStart = froxelUVToDepth((float)z / (float)resZ, s.maxDistance);
            float depthEnd = froxelUVToDepth((float)(z + 1) / (float)resZ, s.maxDistance);
            float segmentLength = depthEnd - depthStart;
            vec3 inscattering = integrateInscattering(it.xyz(), vec3(it.w), segmentLength);
            inscatteringTotal += inscattering;
            transmittance *= exp(-it.w * segmentLength);
            integrationVolume.store(x, y, z, vec4(inscatteringTotal, transmittance));
        }
    });
}

// ---------------- temporalFilter.comp + temporalReprojection.inc + bicubicSampling.inc ----------------
struct Nb { vec3 v[3][3]; };
static vec3 taaTonemap(vec3 color) { return color / (1.f + computeLuminance(color)); }          // temporalReprojection.inc:34-36
static vec3 taaTonemapReverse(vec3 color) { return color / (1.f - computeLuminance(color)); }   // :38-40
static Nb sampleNeighbourhood(const View& tex, vec2 uv, vec2 texelSize, bool useTonemapping) {   // :42-52
    Nb n;
    for (int x = -1; x <= 1; x++)
        for (int y = -1; y <= 1; y++) {
            vec3 color = texture(tex, s_linearClamp, uv + texelSize * vec2((float)x, (float)y)).xyz();
            color = useTonemapping ? taaTonemap(color) : color;
            n.v[x + 1][y + 1] = color;
        }
    return n;
}
static vec3 clipAABB(vec3 target, vec3 bbMin, vec3 bbMax) {  // :8-30
    const vec3 epsilon = vec3(0.0001f);
    vec3 center = 0.5f * (bbMax + bbMin);
    vec3 extend = 0.5f * (bbMax - bbMin) + epsilon;
    vec3 toTarget = target - center;
    vec3 toTargetNorm = toTarget / extend;
    vec3 a = abs(toTargetNorm);
    float maxComponent = max(a.x, max(a.y, a.z));
    if (maxComponent < 1.f) return target;
    return center + toTarget / maxComponent;
}
static float catmullRomWeight1D(float d) {  // bicubicSampling.inc:4-17 (second branch uses the signed d, as written)
    float d1 = abs(d);
    float d2 = d1 * d1;
    float d3 = d2 * d1;
    if (d1 <= 1.f) return (1.f / 6.f) * (9.f * d3 - 15.f * d2 + 6.f);
    else if (d1 <= 2.f) return (1.f / 6.f) * (-3.f * d3 + 15.f * d2 - 24.f * d + 12.f);
    return 0.f;
}
static float computeNeighbourhoodContrast(const Nb& n) {  // temporalFilter.comp:59-69
    float c11 = computeLuminance(n.v[1][1]);
    return abs(computeLuminance(n.v[0][0]) - c11) + abs(computeLuminance(n.v[1][0]) - c11) + abs(computeLuminance(n.v[2][0]) - c11) +
           abs(computeLuminance(n.v[0][2]) - c11) + abs(computeLuminance(n.v[1][2]) - c11) + abs(computeLuminance(n.v[2][2]) - c11) +
           abs(computeLuminance(n.v[0][1]) - c11) + abs(computeLuminance(n.v[2][1]) - c11);
}
struct BicubicW { vec2 w0, w1, w2, w3, wB, t, uvTrunc; };
static BicubicW bicubicWeights(vec2 iUV) {  // bicubicSampling.inc:74-85
    BicubicW b;
    b.uvTrunc = floor(iUV - 0.5f) + 0.5f;
    vec2 f = iUV - b.uvTrunc;
    vec2 f2 = f * f;
    vec2 f3 = f2 * f;
    b.w0 = -0.5f * f3 + f2 - 0.5f * f;
    b.w1 = 1.5f * f3 - 2.5f * f2 + 1.f;
    b.w2 = -1.5f * f3 + 2.f * f2 + 0.5f * f;
    b.w3 = 0.5f * f3 - 0.5f * f2;
    b.wB = b.w1 + b.w2;
    b.t = b.w2 / b.wB;
    return b;
}

// history sample of temporalFilter.comp:104-127: bilinear, or the bicubic samplers of bicubicSampling.inc:28-181 at the
// reprojected pixel position p = iUV + 0.5 + motion * screenResolution
static vec3 sampleHistory(int historySampleTech, const View& historyBufferSrc, vec2 uv, vec2 motion, vec2 p, vec2 texelSize, const Nb& nb) {
    vec3 historySample;
    if (historySampleTech == 0) {
        historySample = texture(historyBufferSrc, s_linearClamp, uv + motion).xyz();
    } else if (historySampleTech == 1) {  // 16 tap, bicubicSampling.inc:28-67
        vec2 uvTrunc = floor(p - 0.5f) + 0.5f;
        vec2 d = p - uvTrunc;
        vec2 ad = abs(d);
        vec2 w[4] = {vec2(catmullRomWeight1D(ad.x + 1.f), catmullRomWeight1D(ad.y + 1.f)), vec2(catmullRomWeight1D(ad.x), catmullRomWeight1D(ad.y)),
                     vec2(catmullRomWeight1D(1.f - ad.x), catmullRomWeight1D(1.f - ad.y)), vec2(catmullRomWeight1D(2.f - ad.x), catmullRomWeight1D(2.f - ad.y))};
        vec2 u[4] = {(uvTrunc - 1.f) * texelSize, uvTrunc * texelSize, (uvTrunc + 1.f) * texelSize, (uvTrunc + 2.f) * texelSize};
        vec3 acc = vec3(0.f);
        bool first = true;
        for (int yy = 0; yy < 4; yy++)
            for (int xx = 0; xx < 4; xx++) {
                vec3 term = texture(historyBufferSrc, s_linearClamp, vec2(u[xx].x, u[yy].y)).xyz() * w[xx].x * w[yy].y;
                acc = first ? term : acc + term;
                first = false;
            }
        historySample = acc;
    } else if (historySampleTech == 2) {  // 9 tap :72-107
        BicubicW b = bicubicWeights(p);
        vec2 uv0 = (b.uvTrunc - 1.f) * texelSize, uvT = (b.uvTrunc + b.t) * texelSize, uv3 = (b.uvTrunc + 2.f) * texelSize;
        auto T = [&](float x, float y) { return texture(historyBufferSrc, s_linearClamp, vec2(x, y)).xyz(); };
        historySample = T(uv0.x, uv0.y) * b.w0.x * b.w0.y + T(uv0.x, uvT.y) * b.w0.x * b.wB.y + T(uv0.x, uv3.y) * b.w0.x * b.w3.y +
                        T(uvT.x, uv0.y) * b.wB.x * b.w0.y + T(uvT.x, uvT.y) * b.wB.x * b.wB.y + T(uvT.x, uv3.y) * b.wB.x * b.w3.y +
                        T(uv3.x, uv0.y) * b.w3.x * b.w0.y + T(uv3.x, uvT.y) * b.w3.x * b.wB.y + T(uv3.x, uv3.y) * b.w3.x * b.w3.y;
    } else if (historySampleTech == 3) {  // 5 tap :112-145
        BicubicW b = bicubicWeights(p);
        vec2 uv0 = (b.uvTrunc - 1.f) * texelSize, uvT = (b.uvTrunc + b.t) * texelSize, uv3 = (b.uvTrunc + 2.f) * texelSize;
        auto T = [&](float x, float y) { return vec4(texture(historyBufferSrc, s_linearClamp, vec2(x, y)).xyz(), 1.f); };
        vec4 result = T(uv0.x, uvT.y) * b.w0.x * b.wB.y + T(uvT.x, uv0.y) * b.wB.x * b.w0.y + T(uvT.x, uvT.y) * b.wB.x * b.wB.y +
                      T(uvT.x, uv3.y) * b.wB.x * b.w3.y + T(uv3.x, uvT.y) * b.w3.x * b.wB.y;
        historySample = result.xyz() / result.w;
    } else if (historySampleTech == 4) {  // 1 tap :150-181
        BicubicW b = bicubicWeights(p);
        vec2 uvT = (b.uvTrunc + b.t) * texelSize;
        vec3 hs = texture(historyBufferSrc, s_linearClamp, uvT).xyz();
        vec4 result = vec4(hs + nb.v[0][1] - nb.v[1][1], 1.f) * b.w0.x * b.wB.y + vec4(hs + nb.v[1][0] - nb.v[1][1], 1.f) * b.wB.x * b.w0.y +
                      vec4(hs, 1.f) * b.wB.x * b.wB.y + vec4(hs + nb.v[1][2] - nb.v[1][1], 1.f) * b.wB.x * b.w3.y +
                      vec4(hs + nb.v[2][1] - nb.v[1][1], 1.f) * b.w3.x * b.wB.y;
        historySample = result.xyz() / result.w;
    } else {
        historySample = vec3(1.f, 0.f, 0.f);
    }
    return historySample;
}

ORACLE_PASS(pass_temporalFilter, "temporalFilter.comp") {
    const bool useClipping = c.specBool(0, false);
    const bool useMotionVectorDilation = c.specBool(1, false);
    const int historySampleTech = c.spec<int>(2, 0);
    const bool useTonemap = c.specBool(3, false);
    View currentFrame = c.sampled(0), historyBufferSrc = c.sampled(3), motionBuffer = c.sampled(4), depthBuffer = c.sampled(5);
    View outputImage = c.storage(1), historyBufferDst = c.storage(2);
    float rw[9];
    memcpy(rw, c.ubuf(6), sizeof(rw));
    const plain_global_shader_info& g = c.g;
    const vec2 screenRes((float)g.screenResolution[0], (float)g.screenResolution[1]);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= outputImage.w() || iy >= outputImage.h()) return;
        ivec2 iUV(ix, iy);
        vec2 texelSize = 1.f / vec2((float)outputImage.w(), (float)outputImage.h());
        vec2 uv = (tovec2(iUV) + 0.5f) * texelSize;
        Nb nb = sampleNeighbourhood(currentFrame, uv, texelSize, useTonemap);
        vec3 mn = nb.v[0][0], mx = nb.v[0][0];  // minMaxFromNeighbourhood, temporalReprojection.inc:54-65
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) { mn = min(mn, nb.v[i][j]); mx = max(mx, nb.v[i][j]); }
        // resolveColor :41-57, weights w{x}_{y} in buffer order w0_0 w1_0 w2_0 w0_1 ...
        vec3 currentColor = vec3(0.f);
        currentColor += nb.v[0][0] * rw[0]; currentColor += nb.v[1][0] * rw[1]; currentColor += nb.v[2][0] * rw[2];
        currentColor += nb.v[0][1] * rw[3]; currentColor += nb.v[1][1] * rw[4]; currentColor += nb.v[2][1] * rw[5];
        currentColor += nb.v[0][2] * rw[6]; currentColor += nb.v[1][2] * rw[7]; currentColor += nb.v[2][2] * rw[8];

        vec2 motion;
        if (useMotionVectorDilation) {  // getClosestFragmentMotion, temporalReprojection.inc:67-83
            float closestDepth = 0.f;
            ivec2 closestDepthOffset(0, 0);
            for (int x = -1; x <= 1; x++)
                for (int y = -1; y <= 1; y++) {
                    float depth = depthBuffer.fetch(ix + x, iy + y).x;
                    if (depth > closestDepth) { closestDepth = depth; closestDepthOffset = ivec2(x, y); }
                }
            motion = motionBuffer.fetch(ix + closestDepthOffset.x, iy + closestDepthOffset.y).xy();
        } else {
            motion = motionBuffer.fetch(ix, iy).xy();
        }

        vec3 historySample = sampleHistory(historySampleTech, historyBufferSrc, uv, motion, tovec2(iUV) + 0.5f + motion * screenRes, texelSize, nb);
        if (useTonemap) historySample = taaTonemap(historySample);
        if (useClipping) historySample = clipAABB(historySample, mn, mx);
        else historySample = clamp(historySample, mn, mx);
        if (isnan(historySample.x) || isnan(historySample.y) || isnan(historySample.z)) historySample = currentColor;

        float currentContrast = computeNeighbourhoodContrast(nb);
        Nb lastNb = sampleNeighbourhood(historyBufferSrc, uv + motion, texelSize, useTonemap);
        float lastContrast = computeNeighbourhoodContrast(lastNb);
        float contrastChange = abs(currentContrast - lastContrast);
        contrastChange = clamp(contrastChange, 0.f, 1.f);
        float blendMin = 0.03f;
        float blendMax = 0.13f;
        float blendFactor = mix(blendMax, blendMin, contrastChange);
        if (g.cameraCut) blendFactor = 1.f;
        vec2 ur = uv + motion;
        if (ur.x < 0.f || ur.y < 0.f || ur.x > 1.f || ur.y > 1.f) {  // isUVOutOfImage
            blendFactor = 1.f;
            // gaussianFilteredNeighbourhood :71-82
            currentColor = nb.v[0][0] * 0.0625f + nb.v[0][2] * 0.0625f + nb.v[2][0] * 0.0625f + nb.v[2][2] * 0.0625f + nb.v[1][0] * 0.125f +
                           nb.v[0][1] * 0.125f + nb.v[1][2] * 0.125f + nb.v[2][1] * 0.125f + nb.v[1][1] * 0.25f;
        }
        vec3 color = mix(historySample, currentColor, blendFactor);
        if (useTonemap) color = taaTonemapReverse(color);
        historyBufferDst.store(ix, iy, 0, vec4(color, 1.f));
        outputImage.store(ix, iy, 0, vec4(color, 1.f));
    });
}

// ---------------- colorToLuminance.comp:13-21 ----------------
ORACLE_PASS(pass_colorToLuminance, "colorToLuminance.comp") {
    View srcTexture = c.sampled(0), dstImage = c.storage(1);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= dstImage.w() || iy >= dstImage.h()) return;
        vec3 color = srcTexture.fetch(ix, iy).xyz();
        float l = computeLuminance(color);
        dstImage.store(ix, iy, 0, vec4(l, 0.f, 0.f, 0.f));
    });
}

// ---------------- temporalSupersampling.comp:21-110 ----------------
static float minAbsoluteDifference(float s, vec4 v) {  // :22-28 (as written: differences of absolute values, not absolute differences)
    return min(abs(s) - abs(v.x), min(abs(s) - abs(v.y), min(abs(s) - abs(v.z), abs(s) - abs(v.w))));
}
static float computeLuminanceBlockDifference(vec4 cur, vec4 last) {  // :30-36
    return minAbsoluteDifference(cur.x, last) + minAbsoluteDifference(cur.y, last) + minAbsoluteDifference(cur.z, last) + minAbsoluteDifference(cur.w, last);
}
static float getClosestNeighbourhoodDepth(const View& depthBuffer, vec2 uv, const plain_global_shader_info& g) {  // :38-55
    vec2 texelSize = 1.f / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
    float closestDepth = texture(depthBuffer, s_nearestClamp, uv + vec2(-1.f, -1.f) * texelSize).x;
    static const int off[8][2] = {{0, -1}, {1, -1}, {-1, 0}, {0, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    for (int i = 0; i < 8; i++) closestDepth = max(texture(depthBuffer, s_nearestClamp, uv + vec2((float)off[i][0], (float)off[i][1]) * texelSize).x, closestDepth);
    return linearizeDepth(closestDepth, g.nearPlane, g.farPlane);
}
ORACLE_PASS(pass_temporalSupersampling, "temporalSupersampling.comp") {
    const bool useTonemap = c.specBool(0, false);
    View currentFrame = c.sampled(1), lastFrame = c.sampled(2), velocityBuffer = c.sampled(4), currentDepthBuffer = c.sampled(5), lastDepthBuffer = c.sampled(6);
    View currentLuminanceTexture = c.sampled(7), lastLuminanceTexture = c.sampled(8);
    View targetImage = c.storage(3);
    const plain_global_shader_info& g = c.g;
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix >= targetImage.w() || iy >= targetImage.h()) return;
        vec2 texelSize = 1.f / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
        vec2 uvCurrent = (vec2((float)ix, (float)iy) + vec2(0.5f)) * texelSize;
        // getClosestFragmentMotion, temporalReprojection.inc:67-83
        float closestDepth = 0.f;
        ivec2 closestDepthOffset(0, 0);
        for (int x = -1; x <= 1; x++)
            for (int y = -1; y <= 1; y++) {
                float depth = currentDepthBuffer.fetch(ix + x, iy + y).x;
                if (depth > closestDepth) { closestDepth = depth; closestDepthOffset = ivec2(x, y); }
            }
        vec2 motion = velocityBuffer.fetch(ix + closestDepthOffset.x, iy + closestDepthOffset.y).xy();
        vec2 uvLast = uvCurrent + motion;
        vec3 currentSample = texture(currentFrame, s_linearClamp, uvCurrent).xyz();
        vec3 lastSample = texture(lastFrame, s_linearClamp, uvLast).xyz();
        if (useTonemap) {
            currentSample = taaTonemap(currentSample);
            lastSample = taaTonemap(lastSample);
        }
        // acceptLastFrameSample :57-84
        vec4 currentLuminance = textureGather(currentLuminanceTexture, s_nearestClamp, uvCurrent);
        vec4 lastLuminance = textureGather(lastLuminanceTexture, s_nearestClamp, uvLast);
        float contrast = computeLuminanceBlockDifference(currentLuminance, lastLuminance);
        bool contrastTest = contrast < 0.5f;
        float currentDepth = getClosestNeighbourhoodDepth(currentDepthBuffer, uvCurrent, g);
        float lastDepth = getClosestNeighbourhoodDepth(lastDepthBuffer, uvLast, g);
        float depthDifference = abs(currentDepth - lastDepth);
        bool depthTest = depthDifference < 1.f;
        bool outOfScreen = uvLast.x < 0.f || uvLast.y < 0.f || uvLast.x > 1.f || uvLast.y > 1.f;
        bool acceptSample = contrastTest && depthTest && !outOfScreen;
        float blendFactor = acceptSample ? 0.5f : 0.f;
        vec3 color = mix(currentSample, lastSample, blendFactor);
        if (useTonemap) color = taaTonemapReverse(color);
        targetImage.store(ix, iy, 0, vec4(color, 1.f));
    });
}

// ---------------- bloomDownsample.comp:12-50 ----------------
ORACLE_PASS(pass_bloomDownsample, "bloomDownsample.comp") {
    View target = c.storage(0);
    View source = c.sampled(1);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix > target.w() || iy > target.h()) return;  // '>' as in the reference (:16); the extra store is dropped
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) / vec2((float)target.w(), (float)target.h());
        vec2 texelSize = 1.f / tovec2(textureSize(source));
        vec3 color = vec3(0.f);
        auto T = [&](float ox, float oy) { return texture(source, s_linearClamp, uv + texelSize * vec2(ox, oy)).xyz(); };
        color += texture(source, s_linearClamp, uv).xyz() * 0.125f;
        color += T(0.5f, 0.5f) * 0.125f; color += T(0.5f, -0.5f) * 0.125f; color += T(-0.5f, 0.5f) * 0.125f; color += T(-0.5f, -0.5f) * 0.125f;
        color += T(1.5f, 0.f) * 0.0625f; color += T(-1.5f, 0.f) * 0.0625f; color += T(0.f, 1.5f) * 0.0625f; color += T(0.f, -1.5f) * 0.0625f;
        color += T(1.5f, 1.5f) * 0.03125f; color += T(1.5f, -1.5f) * 0.03125f; color += T(-1.5f, 1.5f) * 0.03125f; color += T(-1.5f, -1.5f) * 0.03125f;
        target.store(ix, iy, 0, vec4(color, 0.f));
    });
}

// ---------------- bloomUpsample.comp:19-58 ----------------
ORACLE_PASS(pass_bloomUpsample, "bloomUpsample.comp") {
    View target = c.storage(0);
    View targetPreviousMip = c.sampled(1), source = c.sampled(2);
    const bool isLowestMip = c.specBool(0, false);
    const float blurRadius = c.push<float>(0);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix > target.w() || iy > target.h()) return;
        vec2 texelSize = 1.f / tovec2(textureSize(source));
        vec2 sampleStepSize = blurRadius * texelSize;
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) / vec2((float)target.w(), (float)target.h());
        vec3 color = vec3(0.f);
        auto S = [&](float ox, float oy) { return texture(source, s_linearClamp, uv + sampleStepSize * vec2(ox, oy)).xyz(); };
        color += texture(source, s_linearClamp, uv).xyz() * 0.25f;
        color += S(1.f, 0.f) * 0.125f; color += S(-1.f, 0.f) * 0.125f; color += S(0.f, 1.f) * 0.125f; color += S(0.f, -1.f) * 0.125f;
        color += S(1.f, 1.f) * 0.0625f; color += S(1.f, -1.f) * 0.0625f; color += S(-1.f, 1.f) * 0.0625f; color += S(-1.f, -1.f) * 0.0625f;
        if (!isLowestMip) {
            auto P = [&](float ox, float oy) { return texture(targetPreviousMip, s_linearClamp, uv + texelSize * vec2(ox, oy)).xyz(); };
            color += P(0.5f, 0.5f) * 0.25f; color += P(0.5f, -0.5f) * 0.25f; color += P(-0.5f, 0.5f) * 0.25f; color += P(-0.5f, -0.5f) * 0.25f;
        }
        target.store(ix, iy, 0, vec4(color, 0.f));
    });
}

// ---------------- applyBloom.comp:16-31 ----------------
ORACLE_PASS(pass_applyBloom, "applyBloom.comp") {
    View target = c.storage(0);
    View bloomTexture = c.sampled(1);
    const float bloomStrength = c.push<float>(0);
    c.forEachInvocation(8, 8, 1, [&](int ix, int iy, int) {
        if (ix > target.w() || iy > target.h()) return;
        vec2 uv = (vec2((float)ix, (float)iy) + 0.5f) / vec2((float)target.w(), (float)target.h());
        vec3 bloom = texture(bloomTexture, s_linearClamp, uv).xyz();
        vec3 scene = target.fetch(ix, iy).xyz();
        vec3 color = mix(scene, bloom, bloomStrength);
        target.store(ix, iy, 0, vec4(color, 0.f));
    });
}

}  // namespace orc

// the functions above that restate the reference's GLSL include files, behind the batch entry points the reference-pinning test uses
#define INC_NS orc
#define INC_PREFIX oracle_inc_
#define INC_IS_REFERENCE 0
#define INC_PART_TAA 1
#include "inc_eval.h"
