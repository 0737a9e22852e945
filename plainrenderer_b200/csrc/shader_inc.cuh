// shader_inc.cuh - device-side statements of the reference's GLSL include files (resources/shaders/*.inc) used by
// the pass kernels. Each function cites the lines it follows; the operation order is the numeric contract of
// DESIGN.md (left-to-right binary32, no contraction: compile with -fmad=false).
#pragma once
#include "pass_common.cuh"

namespace pb {

// ---- colorConversion.inc ----
PV_HD float linearTosRGB1(float l) {  // :5-13
    const float lo = l * 12.92f;
    const float hi = (dm::pow(absf(l), 1.0f / 2.4f) * 1.055f) - 0.055f;
    return l <= 0.0031308f ? lo : hi;
}
PV_HD vec3 linearTosRGB(vec3 c) { return v3(linearTosRGB1(c.x), linearTosRGB1(c.y), linearTosRGB1(c.z)); }
PV_HD float sRGBToLinear1(float s) {  // :15-23; vec3 / float in the shader: multiplies by the rounded reciprocal (contract 2)
    const float lo = s * rcpf_(12.92f);
    const float hi = dm::pow(absf(s + 0.055f) * rcpf_(1.055f), 2.4f);
    return s <= 0.004045f ? lo : hi;
}
PV_HD vec3 sRGBToLinear(vec3 c) { return v3(sRGBToLinear1(c.x), sRGBToLinear1(c.y), sRGBToLinear1(c.z)); }
PV_HD vec3 linearToYCoCg(vec3 l) {  // :26-31
    return v3(l.x * 0.25f + 0.5f * l.y + 0.25f * l.z, l.x * 0.5f - 0.5f * l.z, -l.x * 0.25f + 0.5f * l.y - 0.25f * l.z);
}
PV_HD vec3 YCoCgToLinear(vec3 c) { return v3(c.x + c.y - c.z, c.x + c.z, c.x - c.y - c.z); }  // :33-38

// ---- tonemapping.inc:17-49 ----
PV_HD vec3 RRTAndODTFit(vec3 v) {
    const vec3 a = v * (v + 0.0245786f) - 0.000090537f;
    const vec3 b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
PV_HD vec3 ACESFitted(vec3 color) {
    color = v3(0.59719f * color.x + 0.35458f * color.y + 0.04823f * color.z,
               0.07600f * color.x + 0.90834f * color.y + 0.01566f * color.z,
               0.02840f * color.x + 0.13383f * color.y + 0.83777f * color.z);
    color = RRTAndODTFit(color);
    color = v3(1.60475f * color.x + -0.53108f * color.y + -0.07367f * color.z,
               -0.10208f * color.x + 1.10813f * color.y + -0.00605f * color.z,
               -0.00327f * color.x + -0.07276f * color.y + 1.07602f * color.z);
    return vclamp(color, 0.f, 1.f);
}

// ---- noise.inc ----
PV_HD vec3 hash32(vec2 q) {  // :14-24
    const uint32_t UI0 = 1597334673U, UI1 = 3812015801U, UI2 = 2798796415U;
    const uint32_t qx = (uint32_t)f2i(q.x), qy = (uint32_t)f2i(q.y);
    uint32_t nx = qx * UI0, ny = qy * UI1, nz = qx * UI2;
    const uint32_t h = nx ^ ny ^ nz;
    nx = h * UI0; ny = h * UI1; nz = h * UI2;
    const float UIF = 1.0f / (float)0xffffffffU;
    return v3((float)nx, (float)ny, (float)nz) * UIF;
}
PV_HD uint32_t xorshift32(uint32_t& state) {  // :28-35
    state ^= (state << 13);
    state ^= (state >> 17);
    state ^= (state << 5);
    return state;
}
PV_HD uint32_t wang_hash(uint32_t seed) {  // :38-46
    seed = (seed ^ 61) ^ (seed >> 16);
    seed *= 9;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2d;
    seed = seed ^ (seed >> 15);
    return seed;
}
PV_HD float rand01(uint32_t& state) {  // :49-54
    const uint32_t x = xorshift32(state);
    state = x;
    return clampf((float)x * dm::u2f(0x2f800004u), 0.f, 1.f);
}

// ---- dither.inc:6-12 ----
PV_HD vec3 ditherRGB8(vec3 c, int ux, int uy, float g_time) {
    const vec2 a = v2((float)ux * g_time, (float)uy * g_time);
    const vec2 b = v2(((float)ux + 165.f) * g_time, ((float)uy + 1292.f) * g_time);
    vec3 noise = hash32(v2((float)f2u(a.x), (float)f2u(a.y)));
    noise = noise + hash32(v2((float)f2u(b.x), (float)f2u(b.y)));
    noise = noise - 1.f;
    noise = noise / 255.f;
    return c + noise;
}

PV_HD float computeLuminance(vec3 color) { return dot(color, v3(0.21f, 0.72f, 0.07f)); }  // luminance.inc:5-7
PV_HD float linearizeDepth(float depth, float nearP, float farP) { return nearP * farP / (farP + (-depth + 1.f) * (nearP - farP)); }  // linearDepth.inc:5-8

// screenToWorld.inc:4-9
PV_HD vec3 calculateViewDirectionFromPixel(vec2 pixelNDC, vec3 cameraForward, vec3 cameraUp, vec3 cameraRight, float cameraTanFovHalf, float aspectRatio) {
    vec3 V = -cameraForward;
    V = V + cameraTanFovHalf * pixelNDC.y * cameraUp;
    V = V - cameraTanFovHalf * aspectRatio * pixelNDC.x * cameraRight;
    return normalize(V);
}
PV_HD vec3 viewDirFromNDC(const Globals& G, vec2 ndc) { return calculateViewDirectionFromPixel(ndc, G.fwd, G.up, G.right, G.tanFovHalf, G.aspect); }

// indirectLightUpscale.comp:17-71 for one full-resolution pixel: depth-aware upscale of the half-resolution GI (Y_SH RGBA16F + CoCg RG16F).
// Shared by giUpscaleKernel (passes_gi.cu) and by the shading kernel when the backend folds the upscale into its consumer (triangle.frag:294-320).
struct UpscaleSource { ImgView srcYSH, srcCoCg, fullResDepth, halfResDepth; };
__device__ __forceinline__ void giUpscalePixel(const UpscaleSource& p, const plain_global_shader_info* __restrict__ g, int ix, int iy, vec4& result_Y_SH, vec2& result_CoCg) {
    const float nearP = g->nearPlane, farP = g->farPlane;
    const vec2 uv = (v2((float)ix, (float)iy) + 0.5f) / v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    float fullResDepth = sampleNearest2D<WRAP_CLAMP, float>([&](int x, int y) { return loadD32(p.fullResDepth, x, y); }, p.fullResDepth.w, p.fullResDepth.h, uv, 0.f);
    fullResDepth = linearizeDepth(fullResDepth, nearP, farP);
    const vec2 halfResTexelSize = 1.f / v2((float)p.halfResDepth.w, (float)p.halfResDepth.h);
    // textureGather: (i0,j1), (i1,j1), (i1,j0), (i0,j0) with i0 = floor(u*w - 0.5), clamp-to-edge
    const int gx0 = f2i(floorf_(sanitizeCoord(uv.x) * (float)p.halfResDepth.w - 0.5f));
    const int gy0 = f2i(floorf_(sanitizeCoord(uv.y) * (float)p.halfResDepth.h - 0.5f));
    auto G4 = [&](int x, int y) { return loadR16F(p.halfResDepth, iclamp(x, 0, p.halfResDepth.w - 1), iclamp(y, 0, p.halfResDepth.h - 1)); };
    float depthSamples[4] = {G4(gx0, gy0 + 1), G4(gx0 + 1, gy0 + 1), G4(gx0 + 1, gy0), G4(gx0, gy0)};
    for (int i = 0; i < 4; i++) depthSamples[i] = linearizeDepth(depthSamples[i], nearP, farP);
    float minDepthDiff = 1000.f;
    vec2 closestDepthTexel = v2(0.f);
    const float edgeDepthThreshold = 0.5f;
    bool isEdge = false;
    const float offX[4] = {0.f, 1.f, 1.f, 0.f}, offY[4] = {1.f, 1.f, 0.f, 0.f};
    for (int i = 0; i < 4; i++) {
        const float depthDiff = absf(depthSamples[i] - fullResDepth);
        isEdge = isEdge || depthDiff > edgeDepthThreshold;
        if (depthDiff < minDepthDiff) {
            minDepthDiff = depthDiff;
            closestDepthTexel = v2(offX[i], offY[i]);
        }
    }
    const vec2 uvClosestTexel = uv + closestDepthTexel * halfResTexelSize;
    if (isEdge) {
        result_Y_SH = sampleNearest2D<WRAP_CLAMP, vec4>([&](int x, int y) { return loadRGBA16F(p.srcYSH, x, y); }, p.srcYSH.w, p.srcYSH.h, uvClosestTexel, v4(0.f));
        result_CoCg = sampleNearest2D<WRAP_CLAMP, vec2>([&](int x, int y) { return loadRG16F(p.srcCoCg, x, y); }, p.srcCoCg.w, p.srcCoCg.h, uvClosestTexel, v2(0.f));
    } else {
        result_Y_SH = sampleRGBA16FLinearClamp(p.srcYSH, uv);
        result_CoCg = sampleRG16FLinearClamp(p.srcCoCg, uv);
    }
}

// ---- brdf.inc ----
PV_HD float D_GGX(float NoH, float r) {  // :4-8
    const float a = NoH * r;
    const float k = r / (1.0f - NoH * NoH + a * a);
    return k * k * (1.0f / PV_PI);
}
PV_HD float Visibility(float NoV, float NoL, float r) {  // :21-26
    const float r_2 = r * r;
    const float v1 = NoL * sqrtf_(NoV * NoV * (1.f - r_2) + r_2);
    const float v2 = NoV * sqrtf_(NoL * NoL * (1.f - r_2) + r_2);
    return 0.5f / (v1 + v2);
}
PV_HD vec3 F_Schlick(vec3 f0, vec3 f90, float VoH) { return f0 + (f90 - f0) * dm::pow(1.f - VoH, 5.f); }  // :34-36
PV_HD vec3 DisneyDiffuse(vec3 diffuseColor, float NoL, float VoH, float NoV, float r) {  // :39-46
    const float energyBias = mixf(0.f, 0.5f, r);
    const float energyFactor = mixf(1.f, 1.f / 1.51f, r);
    const float fresnelDiffuse90Biased = energyBias + 2.f * VoH * VoH * r;
    return diffuseColor / PV_PI * F_Schlick(v3(1.f), v3(fresnelDiffuse90Biased), NoL) * F_Schlick(v3(1.f), v3(fresnelDiffuse90Biased), NoV) * energyFactor;
}
PV_HD vec3 CoDWWIIDiffuse(vec3 diffuseColor, float NoL, float VoH, float NoV, float NoH, float r) {  // :49-58
    const float f0Diffuse = VoH + dm::pow(1.f - VoH, 5.f);
    const float f1 = (1.f - 0.75f * dm::pow(1.f - NoL, 5.f)) * (1.f - 0.75f * dm::pow(1.f - NoV, 5.f));
    const float g = dm::log2(2.f / (r * r) - 1.f) / 18.f;
    const float t = clampf(2.2f * g - 0.5f, 0.f, 1.f);
    const float fd = f0Diffuse + (f1 - f0Diffuse) * t;
    const float fb = (34.5f * g * g - 59.f * g + 24.5f) * VoH * dm::pow(2.f, -fmaxp(73.2f * g - 21.2f, 8.9f) * sqrtf_(NoH));
    return diffuseColor / PV_PI * (fd + fb);
}
PV_HD float Titanfall2DiffuseSingleComponent(float NoL, float LoV, float NoV, float NoH, float r) {  // :60-66
    const float facing = 0.5f + 0.5f * LoV;
    const float rough = facing * (0.9f - 0.4f * facing) * (0.5f + NoH) / fmaxp(NoH, 0.03f);
    const float smoothDiffuse = 1.05f * (1.f - dm::pow(1.f - NoL, 5.f)) * (1.f - dm::pow(1.f - NoV, 5.f));
    return 1.f / PV_PI * mixf(smoothDiffuse, rough, r);
}
PV_HD vec3 Titanfall2Diffuse(vec3 diffuseColor, float NoL, float LoV, float NoV, float NoH, float r) {  // :68-72
    const float single = Titanfall2DiffuseSingleComponent(NoL, LoV, NoV, NoH, r);
    const float multi = 0.1159f * r;
    return diffuseColor * (single + diffuseColor * multi);
}
PV_HD vec3 GGXSingleScattering(float r, vec3 f0, float NoH, float NoV, float VoH, float NoL) {  // :74-79
    const float D = D_GGX(NoH, r);
    const float Vis = Visibility(NoV, NoL, r);
    const vec3 F = F_Schlick(f0, v3(1.f), VoH);
    return D * Vis * F;
}

// ---- SphericalHarmonics.inc:5-15 ----
PV_HD vec4 directionToSH_L1(vec3 V) {
    const float s3 = sqrtf_(3.f), sp = sqrtf_(PV_PI);
    return normalize(v4(1.f / (2.f * sp), -s3 * V.y / (2.f * sp), s3 * V.z / (2.f * sp), -s3 * V.x / (2.f * sp)));
}
PV_HD vec3 dominantDirectionFromSH_L1(vec4 c) { return v3(-c.w, -c.y, c.z); }

// ---- sampling.inc ----
PV_HD vec3 hemisphereToWorld(vec3 sampleHemisphere, vec3 N) {  // :13-22 / :35-44
    const vec3 up = absf(N.z) < 0.999f ? v3(0.f, 0.f, 1.f) : v3(1.f, 0.f, 0.f);
    const vec3 tangent = normalize(cross(up, N));
    const vec3 bitangent = cross(N, tangent);
    vec3 sampleWorld = v3(0.f);
    sampleWorld = sampleWorld + sampleHemisphere.x * tangent;
    sampleWorld = sampleWorld + sampleHemisphere.y * bitangent;
    sampleWorld = sampleWorld + sampleHemisphere.z * N;
    return sampleWorld;
}
PV_HD vec3 importanceSampleGGX(vec2 xi, float r, vec3 N) {  // :4-23
    const float r_2 = r * r;
    const float cosTheta = sqrtf_((1.f - xi.y) / (1.f + (r_2 * r_2 - 1.f) * xi.y));
    const float sinTheta = sqrtf_(1.f - cosTheta * cosTheta);
    const float phi = 2.f * PV_PI * xi.x;
    return hemisphereToWorld(v3(dm::cos(phi) * sinTheta, dm::sin(phi) * sinTheta, cosTheta), N);
}
PV_HD vec3 importanceSampleCosine(vec2 xi, vec3 N) {  // :25-45
    const float phi = 2.f * PV_PI * xi.y;
    const float cosTheta = sqrtf_(xi.x);
    const float sinTheta = sqrtf_(1.f - xi.x);
    return hemisphereToWorld(v3(dm::cos(phi) * sinTheta, dm::sin(phi) * sinTheta, cosTheta), N);
}
PV_HD float radicalInverse_VdC(uint32_t bits) {  // :47-54
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
PV_HD vec2 hammersley2d(uint32_t i, uint32_t N) { return v2((float)i / (float)N, radicalInverse_VdC(i)); }

// ---- sky.inc ----
struct AtmosphereCoefficients { vec3 scatterRayleigh, scatterMie, extinction; };
PV_HD AtmosphereCoefficients calculateCoefficients(float height, const plain_atmosphere_settings& a) {  // :12-44
    const float rayleighFactor = dm::exp(-height * (1.f / 8.f));
    const float mieFactor = dm::exp(-height * (1.f / 1.2f));
    const float ozoneFactor = fmaxp(0.f, 1.f - absf(height - 25.f) / 15.f);
    AtmosphereCoefficients c;
    c.scatterRayleigh = rayleighFactor * ld3(a.scatteringRayleighGround);
    c.scatterMie = v3(mieFactor) * a.scatteringMieGround;
    c.extinction = rayleighFactor * ld3(a.extinctionRayleighGround) + v3(mieFactor * a.extinctionMieGround) + ozoneFactor * ld3(a.ozoneExtinction);
    return c;
}
struct Intersection { vec3 pos; float distance; bool hitEarth; };
PV_HD Intersection rayEarthIntersection(vec3 P, vec3 D, vec3 C, float earthRadius, float atmosphere) {  // :62-83
    const vec3 L = C - P;
    const float t_ca = dot(L, D);
    const float d = sqrtf_(dot(L, L) - t_ca * t_ca);
    const float t_hc_earth = sqrtf_(earthRadius * earthRadius - d * d);
    const float t_earth = t_ca - t_hc_earth;
    const float r = earthRadius + atmosphere;
    const float t_hc_atmosphere = sqrtf_(r * r - d * d);
    const float t_atmosphere = t_ca + absf(t_hc_atmosphere);
    Intersection result;
    result.hitEarth = t_earth >= 0.f;
    const float t = result.hitEarth ? t_earth : t_atmosphere;
    result.distance = t;
    result.pos = P + t * D;
    return result;
}
PV_HD vec2 toSkyLut(vec3 V) {  // :85-94
    const float theta = dm::acos(-(V.y));
    float y = theta / PV_PI;
    const float y_lowRange = y * 2.f - 1.f;
    const float y_lowRangeScaled = signf(y_lowRange) * sqrtf_(absf(y_lowRange));
    y = y_lowRangeScaled * 0.5f + 0.5f;
    const float phi = -dm::atan2(V.z, V.x);
    return v2(phi / (2.f * 3.1415f) + 0.5f, y);
}
PV_HD vec3 fromSkyLut(vec2 uv) {  // :96-103
    float theta = (1.f - uv.y) - 0.5f;
    theta = signf(theta) * theta * theta * 2.f;
    theta *= PV_PI;
    theta += PV_PI * 0.5f;
    const float phi = (-uv.x + 0.5f) * 2.f * PV_PI;
    return v3(dm::sin(theta) * dm::cos(phi), dm::cos(theta), dm::sin(theta) * dm::sin(phi));
}
PV_HD vec2 computeLutUV(float height, float atmosphereHeight, vec3 up, vec3 direction) { return v2(height / atmosphereHeight, dot(up, direction) * 0.5f + 0.5f); }  // :105-110
PV_HD vec3 sampleSkyLut(vec3 V, const ImgView& skyLut) {  // :112-116
    vec2 uv = toSkyLut(V);
    uv.y = clampf(uv.y, 0.005f, 0.995f);
    return sampleR11LinearRepeat(skyLut, uv);
}

// ---- volumeShading.inc ----
PV_HD float phaseGreenstein(float VoL, float g) { return (1.f - g * g) / (4.f * PV_PI * dm::pow(1.f + g * g - 2.f * g * VoL, 1.5f)); }  // :4-6
PV_HD float phaseRayleigh(float VoL) { return 3.f / (16.f * PV_PI) * (1.f + VoL * VoL); }  // :14-16
PV_HD float cornetteShanksPhase(float VoL, float g) {  // :18-22
    const float nominator = 3.f / (8.f * PV_PI) * (1.f - g * g) * (1.f + VoL * VoL);
    const float denominator = (2.f + g * g) * dm::pow(1.f + g * g - 2.f * g * VoL, 1.5f);
    return nominator / denominator;
}
PV_HD vec3 integrateInscattering(vec3 inscattering, vec3 extinctionCoefficients, float length) {  // :25-27
    return (inscattering - inscattering * vexp(-extinctionCoefficients * length)) / vmax(extinctionCoefficients, v3(0.00001f));
}

// ---- volumetricFroxelLighting.inc ----
#define PB_MAX_VOLUMETRIC_LIGHTING_DEPTH 30.f  // :4
PV_HD float froxelUVToDepth(float uvZ, float maxDistance) {  // :23-31
    const float remaped = (dm::exp(3.f * uvZ) - 1.f) / (dm::exp(3.f) - 1.f);
    return remaped * maxDistance;
}
PV_HD float depthToFroxelUVZ(float depth, float maxDistance) {  // :33-41
    const float linear = depth / maxDistance;
    return dm::log(linear * (dm::exp(3.f) - 1.f) + 1.f) / 3.f;
}
PV_HD vec4 volumeTextureLookup(vec2 screenUV, float depth, const ImgView& froxelTexture, float maxDistance) {  // :43-49
    return sampleRGBA16FLinearClamp3D(froxelTexture, v3(screenUV.x, screenUV.y, depthToFroxelUVZ(depth, maxDistance)));
}
PV_HD vec3 applyInscatteringTransmittance(vec3 originalColor, vec4 it) { return originalColor * it.w + xyz(it); }  // :51-53

// ---- sunShadowCascades.inc ----
#define PB_SHADOW_SAMPLE_RADIUS 0.03f  // :5
// :13-20 with a nearest sampler on a D16 map; BORDER_WHITE selects the border colour
template <bool BORDER_WHITE> PV_HD float simpleShadow(vec3 posWorld, const float* lightMatrix, const ImgView& shadowMap) {
    vec4 posLightSpace = mulm4(lightMatrix, v4(posWorld, 1.f));
    posLightSpace = posLightSpace / posLightSpace.w;
    const vec2 xy = v2(posLightSpace.x, posLightSpace.y) * 0.5f + 0.5f;
    const float actualDepth = clampf(posLightSpace.z, 0.f, 1.f);
    const float shadowMapDepth = sampleNearest2D<WRAP_BORDER, float>([&](int x, int y) { return loadD16(shadowMap, x, y); }, shadowMap.w, shadowMap.h, xy, BORDER_WHITE ? 1.f : 0.f);
    return actualDepth > shadowMapDepth ? 1.f : 0.f;
}

}  // namespace pb
