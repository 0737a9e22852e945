// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// shader_inc.h - scalar C++ restatement of the reference's GLSL include files (resources/shaders/*.inc).
// Every function cites the lines it follows. All literals are binary32 (GLSL has no implicit doubles).
#pragma once
#include "image.h"

namespace orc {

// ---- colorConversion.inc ----
inline vec3 linearTosRGB(vec3 linear) {  // colorConversion.inc:5-13
    vec3 lo = linear * 12.92f;
    vec3 hi = (pow(abs(linear), vec3(1.0f / 2.4f)) * 1.055f) - 0.055f;
    return vec3(linear.x <= 0.0031308f ? lo.x : hi.x, linear.y <= 0.0031308f ? lo.y : hi.y, linear.z <= 0.0031308f ? lo.z : hi.z);
}
inline vec3 sRGBToLinear(vec3 s) {  // colorConversion.inc:15-23
    vec3 lo = s / 12.92f;
    vec3 hi = pow(abs(s + 0.055f) / 1.055f, vec3(2.4f));
    return vec3(s.x <= 0.004045f ? lo.x : hi.x, s.y <= 0.004045f ? lo.y : hi.y, s.z <= 0.004045f ? lo.z : hi.z);
}
inline vec3 linearToYCoCg(vec3 l) {  // colorConversion.inc:26-31
    return vec3(l.x * 0.25f + 0.5f * l.y + 0.25f * l.z, l.x * 0.5f - 0.5f * l.z, -l.x * 0.25f + 0.5f * l.y - 0.25f * l.z);
}
inline vec3 YCoCgToLinear(vec3 c) {  // colorConversion.inc:33-38
    return vec3(c.x + c.y - c.z, c.x + c.z, c.x - c.y - c.z);
}

// ---- tonemapping.inc:17-49 ----
inline vec3 RRTAndODTFit(vec3 v) {
    vec3 a = v * (v + 0.0245786f) - 0.000090537f;
    vec3 b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
inline vec3 ACESFitted(vec3 color) {
    // transpose(ACESInputMat) * color: the GLSL initialiser lists columns, the transpose makes them rows
    color = vec3(0.59719f * color.x + 0.35458f * color.y + 0.04823f * color.z,
                 0.07600f * color.x + 0.90834f * color.y + 0.01566f * color.z,
                 0.02840f * color.x + 0.13383f * color.y + 0.83777f * color.z);
    color = RRTAndODTFit(color);
    color = vec3(1.60475f * color.x + -0.53108f * color.y + -0.07367f * color.z,
                 -0.10208f * color.x + 1.10813f * color.y + -0.00605f * color.z,
                 -0.00327f * color.x + -0.07276f * color.y + 1.07602f * color.z);
    return clamp(color, 0.f, 1.f);
}

// ---- noise.inc ----
inline vec3 hash32(vec2 q) {  // noise.inc:14-24
    const uint UI0 = 1597334673U, UI1 = 3812015801U, UI2 = 2798796415U;
    uint qx = (uint)f2int(q.x), qy = (uint)f2int(q.y);
    uint nx = qx * UI0, ny = qy * UI1, nz = qx * UI2;
    uint h = nx ^ ny ^ nz;
    nx = h * UI0; ny = h * UI1; nz = h * UI2;
    const float UIF = 1.0f / (float)0xffffffffU;
    return vec3((float)nx, (float)ny, (float)nz) * UIF;
}
inline uint xorshift32(uint& state) {  // noise.inc:28-35
    state ^= (state << 13);
    state ^= (state >> 17);
    state ^= (state << 5);
    return state;
}
inline uint wang_hash(uint seed) {  // noise.inc:38-46
    seed = (seed ^ 61) ^ (seed >> 16);
    seed *= 9;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2d;
    seed = seed ^ (seed >> 15);
    return seed;
}
inline float rand(uint& state) {  // noise.inc:49-54
    uint x = xorshift32(state);
    state = x;
    return clamp((float)x * uintBitsToFloat(0x2f800004u), 0.f, 1.f);
}

// ---- dither.inc:6-12. hash32(uvec2(...)): the uvec2 is converted back to vec2 by the call ----
inline vec3 ditherRGB8(vec3 c, ivec2 uv, float g_time) {
    vec2 a = vec2((float)uv.x * g_time, (float)uv.y * g_time);
    vec2 b = vec2(((float)uv.x + 165.f) * g_time, ((float)uv.y + 1292.f) * g_time);
    vec3 noise = hash32(vec2((float)f2uint(a.x), (float)f2uint(a.y)));
    noise += hash32(vec2((float)f2uint(b.x), (float)f2uint(b.y)));
    noise -= 1.f;
    noise /= 255.f;
    return c + noise;
}

// ---- luminance.inc:5-7 ----
inline float computeLuminance(vec3 color) { return dot(color, vec3(0.21f, 0.72f, 0.07f)); }

// ---- linearDepth.inc:5-8 ----
inline float linearizeDepth(float depth, float near, float far) { return near * far / (far + (-depth + 1.f) * (near - far)); }

// ---- screenToWorld.inc:4-9 ----
inline vec3 calculateViewDirectionFromPixel(vec2 pixelNDC, vec3 cameraForward, vec3 cameraUp, vec3 cameraRight, float cameraTanFovHalf, float aspectRatio) {
    vec3 V = -cameraForward;
    V += cameraTanFovHalf * pixelNDC.y * cameraUp;
    V -= cameraTanFovHalf * aspectRatio * pixelNDC.x * cameraRight;
    return normalize(V);
}

// ---- brdf.inc ----
inline float D_GGX(float NoH, float r) {  // brdf.inc:4-8
    float a = NoH * r;
    float k = r / (1.0f - NoH * NoH + a * a);
    return k * k * (1.0f / pi);
}
inline float Visibility(float NoV, float NoL, float r) {  // brdf.inc:21-26
    float r_2 = r * r;
    float v1 = NoL * sqrt(NoV * NoV * (1.f - r_2) + r_2);
    float v2 = NoV * sqrt(NoL * NoL * (1.f - r_2) + r_2);
    return 0.5f / (v1 + v2);
}
inline vec3 F_Schlick(vec3 f0, vec3 f90, float VoH) { return f0 + (f90 - f0) * pow(1.f - VoH, 5.f); }  // brdf.inc:34-36
inline vec3 DisneyDiffuse(vec3 diffuseColor, float NoL, float VoH, float NoV, float r) {  // brdf.inc:39-46
    float energyBias = mix(0.f, 0.5f, r);
    float energyFactor = mix(1.f, 1.f / 1.51f, r);
    float fresnelDiffuse90Biased = energyBias + 2.f * VoH * VoH * r;
    return diffuseColor / pi * F_Schlick(vec3(1.f), vec3(fresnelDiffuse90Biased), NoL) * F_Schlick(vec3(1.f), vec3(fresnelDiffuse90Biased), NoV) * energyFactor;
}
inline vec3 CoDWWIIDiffuse(vec3 diffuseColor, float NoL, float VoH, float NoV, float NoH, float r) {  // brdf.inc:49-58
    float f0Diffuse = VoH + pow(1.f - VoH, 5.f);
    float f1 = (1.f - 0.75f * pow(1.f - NoL, 5.f)) * (1.f - 0.75f * pow(1.f - NoV, 5.f));
    float g = log2(2.f / (r * r) - 1.f) / 18.f;
    float t = clamp(2.2f * g - 0.5f, 0.f, 1.f);
    float fd = f0Diffuse + (f1 - f0Diffuse) * t;
    float fb = (34.5f * g * g - 59.f * g + 24.5f) * VoH * pow(2.f, -max(73.2f * g - 21.2f, 8.9f) * sqrt(NoH));
    return diffuseColor / pi * (fd + fb);
}
inline float Titanfall2DiffuseSingleComponent(float NoL, float LoV, float NoV, float NoH, float r) {  // brdf.inc:60-66
    float facing = 0.5f + 0.5f * LoV;
    float rough = facing * (0.9f - 0.4f * facing) * (0.5f + NoH) / max(NoH, 0.03f);
    float smoothDiffuse = 1.05f * (1.f - pow(1.f - NoL, 5.f)) * (1.f - pow(1.f - NoV, 5.f));
    return 1.f / pi * mix(smoothDiffuse, rough, r);
}
inline vec3 Titanfall2Diffuse(vec3 diffuseColor, float NoL, float LoV, float NoV, float NoH, float r) {  // brdf.inc:68-72
    float single = Titanfall2DiffuseSingleComponent(NoL, LoV, NoV, NoH, r);
    float multi = 0.1159f * r;
    return diffuseColor * (single + diffuseColor * multi);
}
inline vec3 GGXSingleScattering(float r, vec3 f0, float NoH, float NoV, float VoH, float NoL) {  // brdf.inc:74-79
    float D = D_GGX(NoH, r);
    float Vis = Visibility(NoV, NoL, r);
    vec3 F = F_Schlick(f0, vec3(1.f), VoH);
    return D * Vis * F;
}

// ---- SphericalHarmonics.inc:5-15 ----
inline vec4 directionToSH_L1(vec3 V) {
    return normalize(vec4(1.f / (2.f * sqrt(pi)), -sqrt(3.f) * V.y / (2.f * sqrt(pi)), sqrt(3.f) * V.z / (2.f * sqrt(pi)), -sqrt(3.f) * V.x / (2.f * sqrt(pi))));
}
inline vec3 dominantDirectionFromSH_L1(vec4 c) { return vec3(-c.w, -c.y, c.z); }

// ---- sampling.inc ----
inline vec3 importanceSampleGGX(vec2 xi, float r, vec3 N) {  // sampling.inc:4-23
    float r_2 = r * r;
    float cosTheta = sqrt((1.f - xi.y) / (1.f + (r_2 * r_2 - 1.f) * xi.y));
    float sinTheta = sqrt(1.f - cosTheta * cosTheta);
    float phi = 2.f * pi * xi.x;
    vec3 sampleHemisphere = vec3(cos(phi) * sinTheta, sin(phi) * sinTheta, cosTheta);
    vec3 up = abs(N.z) < 0.999f ? vec3(0.f, 0.f, 1.f) : vec3(1.f, 0.f, 0.f);
    vec3 tangent = normalize(cross(up, N));
    vec3 bitangent = cross(N, tangent);
    vec3 sampleWorld = vec3(0.f);
    sampleWorld += sampleHemisphere.x * tangent;
    sampleWorld += sampleHemisphere.y * bitangent;
    sampleWorld += sampleHemisphere.z * N;
    return sampleWorld;
}
inline vec3 importanceSampleCosine(vec2 xi, vec3 N) {  // sampling.inc:25-45
    float phi = 2.f * pi * xi.y;
    float cosTheta = sqrt(xi.x);
    float sinTheta = sqrt(1.f - xi.x);
    vec3 sampleHemisphere = vec3(cos(phi) * sinTheta, sin(phi) * sinTheta, cosTheta);
    vec3 up = abs(N.z) < 0.999f ? vec3(0.f, 0.f, 1.f) : vec3(1.f, 0.f, 0.f);
    vec3 tangent = normalize(cross(up, N));
    vec3 bitangent = cross(N, tangent);
    vec3 sampleWorld = vec3(0.f);
    sampleWorld += sampleHemisphere.x * tangent;
    sampleWorld += sampleHemisphere.y * bitangent;
    sampleWorld += sampleHemisphere.z * N;
    return sampleWorld;
}
inline float radicalInverse_VdC(uint bits) {  // sampling.inc:47-54
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
inline vec2 hammersley2d(uint i, uint N) { return vec2((float)i / (float)N, radicalInverse_VdC(i)); }  // sampling.inc:56-58

// ---- sky.inc ----
struct AtmosphereCoefficients { vec3 scatterRayleigh, scatterMie, extinction; };
inline AtmosphereCoefficients calculateCoefficients(float height, const plain_atmosphere_settings& a) {  // sky.inc:12-44
    float rayleighFactor = exp(-height * (1.f / 8.f));
    float mieFactor = exp(-height * (1.f / 1.2f));
    float ozoneFactor = max(0.f, 1.f - abs(height - 25.f) / 15.f);
    AtmosphereCoefficients c;
    vec3 sR(a.scatteringRayleighGround[0], a.scatteringRayleighGround[1], a.scatteringRayleighGround[2]);
    vec3 eR(a.extinctionRayleighGround[0], a.extinctionRayleighGround[1], a.extinctionRayleighGround[2]);
    vec3 oz(a.ozoneExtinction[0], a.ozoneExtinction[1], a.ozoneExtinction[2]);
    c.scatterRayleigh = rayleighFactor * sR;
    c.scatterMie = vec3(mieFactor) * a.scatteringMieGround;
    c.extinction = rayleighFactor * eR + vec3(mieFactor * a.extinctionMieGround) + ozoneFactor * oz;
    return c;
}
struct Intersection { vec3 pos; float distance; bool hitEarth; };
inline Intersection rayEarthIntersection(vec3 P, vec3 D, vec3 C, float earthRadius, float atmosphere) {  // sky.inc:62-83
    vec3 L = C - P;
    float t_ca = dot(L, D);
    float d = sqrt(dot(L, L) - t_ca * t_ca);
    float t_hc_earth = sqrt(earthRadius * earthRadius - d * d);
    float t_earth = t_ca - t_hc_earth;
    float r = earthRadius + atmosphere;
    float t_hc_atmosphere = sqrt(r * r - d * d);
    float t_atmosphere = t_ca + abs(t_hc_atmosphere);
    Intersection result;
    result.hitEarth = t_earth >= 0.f;  // false when t_earth is NaN (ray misses the earth)
    float t = result.hitEarth ? t_earth : t_atmosphere;
    result.distance = t;
    result.pos = P + t * D;
    return result;
}
inline vec2 toSkyLut(vec3 V) {  // sky.inc:85-94
    float theta = acos(-(V.y));
    float y = theta / pi;
    float y_lowRange = y * 2.f - 1.f;
    float y_lowRangeScaled = sign(y_lowRange) * sqrt(abs(y_lowRange));
    y = y_lowRangeScaled * 0.5f + 0.5f;
    float phi = -atan(V.z, V.x);
    return vec2(phi / (2.f * 3.1415f) + 0.5f, y);
}
inline vec3 fromSkyLut(vec2 uv) {  // sky.inc:96-103
    float theta = (1.f - uv.y) - 0.5f;
    theta = sign(theta) * theta * theta * 2.f;
    theta *= pi;
    theta += pi * 0.5f;
    float phi = (-uv.x + 0.5f) * 2.f * pi;
    return vec3(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi));
}
inline vec2 computeLutUV(float height, float atmosphereHeight, vec3 up, vec3 direction) {  // sky.inc:105-110
    return vec2(height / atmosphereHeight, dot(up, direction) * 0.5f + 0.5f);
}
inline vec3 sampleSkyLut(vec3 V, const View& skyLut) {  // sky.inc:112-116
    vec2 uv = toSkyLut(V);
    uv.y = clamp(uv.y, 0.005f, 0.995f);
    return texture(skyLut, s_linearRepeat, uv).xyz();
}

// ---- volumeShading.inc ----
inline float phaseGreenstein(float VoL, float g) { return (1.f - g * g) / (4.f * pi * pow(1.f + g * g - 2.f * g * VoL, 1.5f)); }  // :4-6
inline float phaseRayleigh(float VoL) { return 3.f / (16.f * pi) * (1.f + VoL * VoL); }  // :14-16
inline float cornetteShanksPhase(float VoL, float g) {  // :18-22
    float nominator = 3.f / (8.f * pi) * (1.f - g * g) * (1.f + VoL * VoL);
    float denominator = (2.f + g * g) * pow(1.f + g * g - 2.f * g * VoL, 1.5f);
    return nominator / denominator;
}
inline vec3 integrateInscattering(vec3 inscattering, vec3 extinctionCoefficients, float length) {  // :25-27
    return (inscattering - inscattering * exp(-extinctionCoefficients * length)) / max(extinctionCoefficients, 0.00001f);
}

// ---- volumetricFroxelLighting.inc ----
static const float maxVolumetricLightingDepth = 30.f;  // :4
inline float froxelUVToDepth(float uvZ, float maxDistance) {  // :23-31, k = 3, exponential distribution
    float remaped = (exp(3.f * uvZ) - 1.f) / (exp(3.f) - 1.f);
    return remaped * maxDistance;
}
inline float depthToFroxelUVZ(float depth, float maxDistance) {  // :33-41
    float linear = depth / maxDistance;
    return log(linear * (exp(3.f) - 1.f) + 1.f) / 3.f;
}
inline vec4 volumeTextureLookup(vec2 screenUV, float depth, const View& froxelTexture, float maxDistance) {  // :43-49
    vec3 uv(screenUV.x, screenUV.y, depthToFroxelUVZ(depth, maxDistance));
    return texture3D(froxelTexture, s_linearClamp, uv);
}
inline vec3 applyInscatteringTransmittance(vec3 originalColor, vec4 it) { return originalColor * it.w + it.xyz(); }  // :51-53

// ---- sunShadowCascades.inc ----
static const float shadowSampleRadius = 0.03f;  // :5
inline float simpleShadow(vec3 posWorld, const mat4& lightMatrix, const View& shadowMap, const Sampler& s) {  // :13-20
    vec4 posLightSpace = lightMatrix * vec4(posWorld, 1.f);
    posLightSpace /= posLightSpace.w;
    vec2 xy = posLightSpace.xy() * 0.5f + 0.5f;
    float actualDepth = clamp(posLightSpace.z, 0.f, 1.f);
    float shadowMapDepth = texture(shadowMap, s, xy).x;
    return actualDepth > shadowMapDepth ? 1.f : 0.f;
}

}  // namespace orc
