// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_shading.cpp - tonemapping.comp, brdfLut.comp and the deferred recast of triangle.frag + sky.frag +
// sunSprite.frag ("gbufferShading.comp").
#include <map>
#include <mutex>
#include <tuple>
#include "backend.h"
#include "shader_inc.h"
#include "shading_hook.h"

namespace orc {

// ---------------- tonemapping.comp:17-27 ----------------
ORACLE_PASS(pass_tonemapping, "tonemapping.comp") {
    View imageOut = c.storage(0);
    View imageIn = c.sampled(1);
    const float g_time = c.g.time;
    c.forEachInvocation(8, 8, 1, [&](int x, int y, int) {
        ivec2 uv(x, y);
        vec3 linearColor = imageIn.fetch(uv).xyz();
        vec3 tonemapped = ACESFitted(linearColor);
        vec3 sRGB = linearTosRGB(tonemapped);
        sRGB = ditherRGB8(sRGB, uv, g_time);
        imageOut.store(uv, vec4(sRGB, 1.f));
    });
}

// ---------------- brdfLut.comp:20-101 ----------------
// The LUT is a pure function of (diffuse BRDF, extent, format) and costs 512 x 512 x 1024 software-transcendental samples (about 14 s on 8 host
// threads), paid by every frontend a test creates: the texels of the first evaluation in a process are kept and copied afterwards. Test
// infrastructure only - the first evaluation of each configuration is the full computation below, and it is what the CUDA kernel is held against.
static std::mutex g_brdfLutMutex;
static std::map<std::tuple<int, int, int, uint32_t>, std::vector<uint8_t>> g_brdfLutCache;
ORACLE_PASS(pass_brdfLut, "brdfLut.comp") {
    View lut = c.storage(0);
    const int diffuseBRDF = c.spec<int>(0, 0);
    const bool wholeImage = (int)c.exec->dispatch[0] * 8 >= lut.w() && (int)c.exec->dispatch[1] * 8 >= lut.h() && lut.d() == 1;
    const auto key = std::make_tuple(diffuseBRDF, lut.w(), lut.h(), lut.format());
    std::vector<uint8_t>& texels = lut.img->mips[(size_t)lut.mip].data;
    if (wholeImage) {
        std::lock_guard<std::mutex> lock(g_brdfLutMutex);
        auto it = g_brdfLutCache.find(key);
        if (it != g_brdfLutCache.end() && it->second.size() == texels.size()) { texels = it->second; return; }
    }
    c.forEachInvocation(8, 8, 1, [&](int ux, int uy, int) {
        float r = (float)ux / (float)lut.w();
        r = max(r, 0.0001f);
        float NoV = max((float)uy, 0.1f) / (float)lut.h();
        vec3 V = vec3(sqrt(1.0f - NoV * NoV), 0.f, NoV);
        vec3 N = vec3(0.f, 0.f, 1.f);
        const int samples = 1024;
        vec3 result = vec3(0.f);
        for (int i = 0; i < samples; i++) {
            vec2 xi = hammersley2d((uint)i, (uint)samples);
            {  // specular
                vec3 H = importanceSampleGGX(xi, r, N);
                vec3 L = 2.f * dot(V, H) * H - V;
                float VoH = max(dot(V, H), 0.f);
                float NoH = max(H.z, 0.f);
                float NoL = max(L.z, 0.f);
                if (NoL > 0.f) {
                    float F_c = pow(1.f - VoH, 5.f);
                    float Vis = Visibility(NoV, NoL, r);
                    float k = Vis * VoH * NoL / NoH;
                    result.x += F_c * k;
                    result.y += k;
                }
            }
            {  // diffuse
                vec3 L = importanceSampleCosine(xi, N);
                vec3 H = normalize(V + L);
                float VoH = clamp(dot(V, H), 0.f, 1.f);
                float NoL = max(L.z, 0.f);
                float NoH = max(H.z, 0.f);
                vec3 F0Diffuse = vec3(0.04f);
                float fresnelInOut = (1.f - F_Schlick(F0Diffuse, vec3(1.f), NoV).x) * (1.f - F_Schlick(F0Diffuse, vec3(1.f), NoL).x);
                if (diffuseBRDF == 0) {
                    result.z += (1.f / pi) * fresnelInOut;
                } else if (diffuseBRDF == 1) {
                    result.z += DisneyDiffuse(vec3(1.f), NoL, VoH, NoV, r).x * fresnelInOut;
                } else if (diffuseBRDF == 2) {
                    result.z += CoDWWIIDiffuse(vec3(1.f), NoL, VoH, NoV, NoH, r).x * fresnelInOut;
                } else if (diffuseBRDF == 3) {
                    float LoV = clamp(dot(L, V), 0.f, 1.f);
                    result.z += Titanfall2DiffuseSingleComponent(NoL, LoV, NoV, NoH, r) * fresnelInOut;
                }
            }
        }
        result /= (float)samples;
        result.x *= 4.f;
        result.y *= 4.f;
        lut.store(ux, uy, 0, vec4(result, 0.f));
    });
    if (wholeImage) {
        std::lock_guard<std::mutex> lock(g_brdfLutMutex);
        g_brdfLutCache[key] = texels;
    }
}

// ---------------- deferred shading over the packed G-buffer ----------------
// G-buffer texel layout: include/plain_frame_types.h. Octahedral decode of the shading normal:
//   f = snorm16 pair; n = (f.x, f.y, 1 - |f.x| - |f.y|); t = max(-n.z, 0); n.xy += (n.xy >= 0) ? -t : t; N = normalize(n)
struct GBufferTexel {
    float depth;
    vec3 N;
    vec3 albedoTexel;
    float specG, specB;
};
static GBufferTexel decodeGBuffer(const View& gb, int x, int y) {
    uint32_t t[4];
    gb.loadUint4(x, y, t);
    GBufferTexel g;
    g.depth = dm::u2f(t[0]);
    vec2 f(snorm16ToFloat((int16_t)(t[1] & 0xffffu)), snorm16ToFloat((int16_t)(t[1] >> 16)));
    vec3 n(f.x, f.y, 1.f - abs(f.x) - abs(f.y));
    float tt = max(-n.z, 0.f);
    n.x += (n.x >= 0.f) ? -tt : tt;
    n.y += (n.y >= 0.f) ? -tt : tt;
    g.N = normalize(n);
    g.albedoTexel = vec3(unorm8ToFloat(t[2] & 0xff), unorm8ToFloat((t[2] >> 8) & 0xff), unorm8ToFloat((t[2] >> 16) & 0xff));
    g.specG = unorm8ToFloat((t[2] >> 24) & 0xff);
    g.specB = unorm8ToFloat(t[3] & 0xff);
    return g;
}

ShadeGeometryHook g_shadeGeometryHook = {nullptr, nullptr};

struct ShadingResources {
    View gbuffer, brdfLut, shadowMaps[4], ySH, coCg, volumetricLUT, skyLut, transmissionLut, noise;
    plain_light_buffer light;
    plain_shadow_cascade_info cascades;
    plain_volumetric_lighting_settings vol;
    plain_global_shader_info g;
    int diffuseBRDF, directMultiscatterBRDF, indirectLightingTech;
    bool geometricAA;
    uint32_t sunShadowCascadeCount;
    mat4 sunSpriteModel;
    bool hasSunSprite;
};

// triangle.frag:92-120
static float calcShadow(const ShadingResources& R, vec2 fragCoord, vec3 pos, const View& shadowMap, const mat4& lightMatrix, int cascade) {
    vec4 posLightSpace = lightMatrix * vec4(pos, 1.f);
    posLightSpace /= posLightSpace.w;
    vec2 plsXY = posLightSpace.xy() * 0.5f + 0.5f;
    float actualDepth = clamp(posLightSpace.z, 0.f, 1.f);
    vec2 noiseUV = fragCoord / tovec2(textureSize(R.noise));
    float noise = texture(R.noise, s_nearestRepeat, noiseUV).x;
    vec2 offsetScale = shadowSampleRadius * vec2(R.cascades.lightSpaceScale[cascade][0], R.cascades.lightSpaceScale[cascade][1]);
    float shadow = 0.f;
    float sampleCount = 12.f;
    for (int i = 0; (float)i < sampleCount; i++) {
        float d = ((float)i + 0.5f * noise) / sampleCount;
        d = sqrt(d);
        float angle = noise * 2.f * pi + 2.f * pi * (float)i / sampleCount;
        vec2 offset = vec2(cos(angle), sin(angle));
        offset *= offsetScale * d;
        vec2 samplePosition = plsXY + offset;
        float depthTexel = texture(shadowMap, s_nearestBlackBorder, samplePosition).x;  // shadowTest, triangle.frag:84-87
        shadow += (actualDepth >= depthTexel) ? 1.f : 0.f;
    }
    return shadow / sampleCount;
}

// triangle.frag:123-131
static float ReflectedEnergyAverage(float roughness) {
    float smoothness = 1.f - sqrt(roughness);
    float r = -0.0761947f - 0.383026f * smoothness;
    r = 1.04997f + smoothness * r;
    r = 0.409255f + smoothness * r;
    return min(0.999f, r);
}

// triangle.frag:146-175
static vec3 computeSpecularMultiscatteringLobe(const ShadingResources& R, float r, float NoL, vec3 f0, vec3 singleScatteringLobe, vec3 brdfLut) {
    vec3 multiScatteringLobe;
    float energyOutgoing = brdfLut.y;
    vec3 fresnelAverage = f0 + (1.f - f0) / 21.f;
    if (R.directMultiscatterBRDF == 0) {
        float energyAverage = ReflectedEnergyAverage(r);
        float energyIncoming = texture(R.brdfLut, s_linearClamp, vec2(r, NoL)).y;
        float multiScatteringLobeUnscaled = (1.f - energyIncoming) * (1.f - energyOutgoing) / (3.1415f * (1.f - energyAverage));
        vec3 multiScatteringScaling = (fresnelAverage * fresnelAverage * energyAverage) / (1.f - fresnelAverage * (1.f - energyAverage));
        multiScatteringLobe = multiScatteringLobeUnscaled * multiScatteringScaling;
    } else if (R.directMultiscatterBRDF == 1) {
        multiScatteringLobe = vec3((1.f - energyOutgoing) / pi);
        vec3 multiScatteringScaling = (fresnelAverage * fresnelAverage * energyOutgoing) / (1.f - fresnelAverage * (1.f - energyOutgoing));
        multiScatteringLobe *= multiScatteringScaling;
    } else if (R.directMultiscatterBRDF == 2) {
        multiScatteringLobe = f0 * (1.f / energyOutgoing - 1.f) * singleScatteringLobe;
    } else {
        multiScatteringLobe = vec3(0.f);
    }
    return multiScatteringLobe;
}

// GeometricAA.inc:4-20 with dFdxFine/dFdyFine on the G-buffer: the quad is pixels (2i,2j)..(2i+1,2j+1),
// dFdx = right - left in the pixel's row, dFdy = bottom - top in its column; a partner outside the image
// is replaced by the pixel itself (difference 0).
static float modifiedRoughnessGeometricAA(const View& gb, int x, int y, float r) {
    float kappa = 0.18f;
    float pixelVariance = 0.5f;
    float pxVar2 = pixelVariance * pixelVariance;
    int x0 = x & ~1, y0 = y & ~1;
    int x1 = (x0 + 1 < gb.w()) ? x0 + 1 : x0;
    int y1 = (y0 + 1 < gb.h()) ? y0 + 1 : y0;
    vec3 N_U = decodeGBuffer(gb, x1, y).N - decodeGBuffer(gb, x0, y).N;
    vec3 N_V = decodeGBuffer(gb, x, y1).N - decodeGBuffer(gb, x, y0).N;
    float lengthN_U2 = dot(N_U, N_U);
    float lengthN_V2 = dot(N_V, N_V);
    float variance = pxVar2 * (lengthN_V2 + lengthN_U2);
    float kernelRoughness2 = min(2.f * variance, kappa);
    return clamp(sqrt(r * r + kernelRoughness2), 0.f, 1.f);
}

static vec3 shadeGeometry(const ShadingResources& R, int x, int y, const GBufferTexel& gbt, vec3 cameraToPixel) {
    const plain_global_shader_info& g = R.g;
    vec2 fragCoord((float)x + 0.5f, (float)y + 0.5f);
    vec3 cameraForward(g.cameraForward[0], g.cameraForward[1], g.cameraForward[2]);
    vec3 cameraPosition(g.cameraPosition[0], g.cameraPosition[1], g.cameraPosition[2]);
    // passPos: world position reconstructed from depth (as sdfDiffuseTrace.comp:122-126)
    float depthLinear = linearizeDepth(gbt.depth, g.nearPlane, g.farPlane);
    vec3 passPos = cameraPosition + cameraToPixel / dot(cameraToPixel, cameraForward) * depthLinear;
    if (g_shadeGeometryHook.shade) {  // liboracle_refmain.so: the rest of this function is the reference's own triangle.frag main()
        FragmentInputs in;
        in.x = x; in.y = y; in.passPos = passPos; in.N = gbt.N; in.albedoTexel = gbt.albedoTexel; in.specG = gbt.specG; in.specB = gbt.specB;
        const View& gb = R.gbuffer;  // the quad's partners as modifiedRoughnessGeometricAA below defines them
        const int x0 = x & ~1, y0 = y & ~1, x1 = (x0 + 1 < gb.w()) ? x0 + 1 : x0, y1 = (y0 + 1 < gb.h()) ? y0 + 1 : y0;
        in.dxN[0] = decodeGBuffer(gb, x0, y).N; in.dxN[1] = decodeGBuffer(gb, x1, y).N;
        in.dyN[0] = decodeGBuffer(gb, x, y0).N; in.dyN[1] = decodeGBuffer(gb, x, y1).N;
        return g_shadeGeometryHook.shade(in);
    }

    // triangle.frag:184-193
    float metalic = gbt.specB;
    float r = gbt.specG;
    r = max(r * r, 0.0045f);
    vec3 albedo = sRGBToLinear(gbt.albedoTexel);
    vec3 diffuseColor = (1.f - metalic) * albedo;
    vec3 N = gbt.N;
    vec3 L = normalize(vec3(g.sunDirection[0], g.sunDirection[1], g.sunDirection[2]));
    vec3 V = cameraPosition - passPos;
    float pixelDepth = dot(V, -cameraForward);
    V = normalize(V);
    vec3 H = normalize(V + L);
    if (R.geometricAA) r = modifiedRoughnessGeometricAA(R.gbuffer, x, y, r);

    const float NoH = max(dot(N, H), 0.f);
    const float NoL = clamp(dot(N, L), 0.f, 1.f);
    const float VoH = abs(dot(V, H));
    const float LoV = max(dot(L, V), 0.f);
    float NoV = abs(dot(N, V));
    NoV = max(NoV, 0.0001f);
    const vec3 f0 = mix(vec3(0.04f), albedo, metalic);

    // sun light, triangle.frag:222-241
    float sunShadow = 0.f;
    int cascadeIndex = 0;
    for (uint32_t cascade = 0; cascade + 1 < R.sunShadowCascadeCount; cascade++) cascadeIndex += (pixelDepth >= R.cascades.splits[cascade]) ? 1 : 0;
    {
        mat4 lm;
        for (int cc = 0; cc < 4; cc++) lm.c[cc] = vec4(R.cascades.lightMatrices[cascadeIndex][cc * 4], R.cascades.lightMatrices[cascadeIndex][cc * 4 + 1], R.cascades.lightMatrices[cascadeIndex][cc * 4 + 2], R.cascades.lightMatrices[cascadeIndex][cc * 4 + 3]);
        sunShadow = calcShadow(R, fragCoord, passPos, R.shadowMaps[cascadeIndex], lm, cascadeIndex);
    }
    vec3 sunColor(R.light.sunColor[0], R.light.sunColor[1], R.light.sunColor[2]);
    vec3 directLighting = max(dot(N, L), 0.f) * sunShadow * sunColor;
    vec3 brdfLut = texture(R.brdfLut, s_linearClamp, vec2(r, NoV)).xyz();

    // direct diffuse, triangle.frag:243-285
    vec3 diffuseDirect;
    vec3 diffuseBRDFIntegral = vec3(1.f);
    if (R.diffuseBRDF == 0) {
        diffuseDirect = diffuseColor / pi * directLighting;
        diffuseBRDFIntegral = vec3(brdfLut.z);
    } else if (R.diffuseBRDF == 1) {
        diffuseDirect = DisneyDiffuse(diffuseColor, NoL, VoH, NoV, r) * directLighting;
        diffuseBRDFIntegral = vec3(brdfLut.z);
    } else if (R.diffuseBRDF == 2) {
        diffuseDirect = CoDWWIIDiffuse(diffuseColor, NoL, VoH, NoV, NoH, r) * directLighting;
        diffuseBRDFIntegral = vec3(brdfLut.z);
    } else {
        diffuseDirect = Titanfall2Diffuse(diffuseColor, NoL, LoV, NoV, NoH, r) * directLighting;
        float multiIntegral = 0.1159f * r * pi * 2.f;
        vec3 F0Diffuse = vec3(0.04f);
        multiIntegral *= (1.f - F_Schlick(F0Diffuse, vec3(1.f), NoV).x);
        multiIntegral *= 0.94291f;
        diffuseBRDFIntegral = min(vec3(brdfLut.z) + diffuseColor * multiIntegral, vec3(1.f));
    }
    diffuseDirect *= (1.f - F_Schlick(f0, vec3(1.f), NoV)) * (1.f - (F_Schlick(f0, vec3(1.f), NoL)));

    // direct specular, triangle.frag:287-290
    vec3 singleScatteringLobe = GGXSingleScattering(r, f0, NoH, NoV, VoH, NoL);
    vec3 multiScatteringLobe = computeSpecularMultiscatteringLobe(R, r, NoL, f0, singleScatteringLobe, brdfLut);
    vec3 specularDirect = directLighting * (singleScatteringLobe + multiScatteringLobe);

    vec3 lightingIndirect;
    if (R.indirectLightingTech == 0) {  // triangle.frag:295-322
        vec2 screenUV = fragCoord / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
        vec4 irradiance_Y_SH = texture(R.ySH, s_nearestClamp, screenUV);
        float irradiance_Y = dot(irradiance_Y_SH, directionToSH_L1(N));
        vec2 irradiance_CoCg = texture(R.coCg, s_nearestClamp, screenUV).xy();
        vec3 irradiance = YCoCgToLinear(vec3(irradiance_Y, irradiance_CoCg.x, irradiance_CoCg.y));
        vec3 diffuseIndirect = irradiance * diffuseColor * diffuseBRDFIntegral;

        vec3 dominantDirection = dominantDirectionFromSH_L1(irradiance_Y_SH);
        float dominantDirectionLength = length(dominantDirection);
        dominantDirectionLength = clamp(dominantDirectionLength, 0.01f, 1.f);
        float r_indirect = mix(1.f, r, sqrt(dominantDirectionLength));
        vec3 L_indirect = dominantDirection / dominantDirectionLength;
        vec3 H_indirect = normalize(L_indirect + V);
        float NoH_indirect = max(dot(N, H_indirect), 0.f);
        float NoL_indirect = max(dot(N, L_indirect), 0.f);
        float VoH_indirect = max(dot(V, H_indirect), 0.f);
        vec3 singleScattering_indirect = GGXSingleScattering(r_indirect, f0, NoH_indirect, NoV, VoH_indirect, NoL_indirect);
        vec3 multiScattering_indirect = computeSpecularMultiscatteringLobe(R, r_indirect, NoL_indirect, f0, singleScattering_indirect, brdfLut);
        vec3 specularIndirect = (singleScattering_indirect + multiScattering_indirect) * YCoCgToLinear(vec3(irradiance_Y_SH.x, irradiance_CoCg.x, irradiance_CoCg.y));
        lightingIndirect = diffuseIndirect + specularIndirect;
    } else {  // constant ambient, triangle.frag:324-333
        float ambientStrength = 0.003f;
        vec3 irradiance = vec3(ambientStrength) * R.light.sunStrengthExposed;
        vec3 reflection = vec3(ambientStrength) * R.light.sunStrengthExposed;
        vec3 singleScattering = mix(vec3(brdfLut.x), vec3(brdfLut.y), f0);
        vec3 diffuseIndirect = irradiance * diffuseColor * diffuseBRDFIntegral;
        vec3 specularIndirect = singleScattering * reflection;
        lightingIndirect = diffuseIndirect + specularIndirect;
    }
    vec3 color = (diffuseDirect + specularDirect) * R.light.sunStrengthExposed + lightingIndirect;

    // applyVolumetricLighting, triangle.frag:133-144
    vec2 noiseUV = fragCoord / tovec2(textureSize(R.noise));
    vec2 noise = texture(R.noise, s_nearestRepeat, noiseUV).xy();
    noise -= 0.5f;
    noise *= 0.013f;
    vec2 screenUV = fragCoord / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
    screenUV += noise;
    vec4 inscatteringTransmittance = volumeTextureLookup(screenUV, pixelDepth, R.volumetricLUT, R.vol.maxDistance);
    return applyInscatteringTransmittance(color, inscatteringTransmittance);
}

// sky.frag:20-33 then sunSprite.frag:23-43 blended additively (RenderPass.cpp:112-124: src*srcAlpha + dst*dstAlpha,
// R11G11B10 has no alpha so dstAlpha reads 1) onto the value already stored in the R11G11B10 target.
static vec3 shadeSky(const ShadingResources& R, int x, int y, vec3 cameraToPixel) {
    const plain_global_shader_info& g = R.g;
    vec2 fragCoord((float)x + 0.5f, (float)y + 0.5f);
    vec2 res((float)g.screenResolution[0], (float)g.screenResolution[1]);
    vec3 V = cameraToPixel;
    vec3 color = sampleSkyLut(V, R.skyLut);
    vec2 d = fragCoord * res;
    color = ditherRGB8(color, ivec2(f2int(d.x), f2int(d.y)), g.time);
    vec2 screenUV = fragCoord / res;
    vec4 inscatteringTransmittance = volumeTextureLookup(screenUV, maxVolumetricLightingDepth, R.volumetricLUT, R.vol.maxDistance);
    color = applyInscatteringTransmittance(color, inscatteringTransmittance);
    color = unpackR11G11B10(packR11G11B10(color));  // render-target store before the blend

    if (R.hasSunSprite) {
        // invert mat3(model) = Rot * diag(s, s, 1): the quad point that projects onto this pixel's ray
        const mat4& M = R.sunSpriteModel;
        vec3 c0 = M.c[0].xyz(), c1 = M.c[1].xyz(), c2 = M.c[2].xyz();
        float s2 = dot(c0, c0);
        vec3 q(dot(c0, V) / s2, dot(c1, V) / s2, dot(c2, V));
        if (q.z < 0.f) {
            vec2 posCentered(q.x / -q.z, q.y / -q.z);
            if (abs(posCentered.x) <= 1.f && abs(posCentered.y) <= 1.f) {
                float distanceFromCenter = dot(posCentered, posCentered);
                if (!(distanceFromCenter > 1.f)) {
                    vec3 passWorldPos = c0 * posCentered.x + c1 * posCentered.y + c2 * -1.f;
                    float bias = 0.002f;
                    vec3 Vs = normalize(passWorldPos + vec3(0.f, bias, 0.f));
                    vec2 lutUV = computeLutUV(0.f, 100.f, vec3(0.f, -1.f, 0.f), Vs);
                    vec3 transmission = texture(R.transmissionLut, s_linearClamp, lutUV).xyz();
                    float mu = sqrt(1.f - distanceFromCenter);
                    vec3 limb = pow(vec3(mu), vec3(0.482f, 0.511f, 0.643f));
                    vec3 sun = R.light.sunStrengthExposed * transmission * limb;
                    float alpha = 1.f - distanceFromCenter;
                    alpha *= alpha;
                    color = sun * alpha + color;
                }
            }
        }
    }
    return color;
}

ORACLE_PASS(pass_gbufferShading, "gbufferShading.comp") {
    ShadingResources R;
    R.gbuffer = c.sampled(0);
    R.brdfLut = c.sampled(3);
    for (int i = 0; i < 4; i++) R.shadowMaps[i] = c.sampled(9 + i);
    R.ySH = c.sampled(15);
    R.coCg = c.sampled(16);
    R.volumetricLUT = c.sampled(18);
    R.skyLut = c.sampled(21);
    R.transmissionLut = c.sampled(22);
    View colorOut = c.storage(20);
    memcpy(&R.light, c.sbuf(7), sizeof(R.light));
    memcpy(&R.cascades, c.sbuf(8), sizeof(R.cascades));
    memcpy(&R.vol, c.ubuf(19), sizeof(R.vol));
    R.g = c.g;
    R.noise = c.bindless((uint32_t)c.g.noiseTextureIndices[c.g.frameIndexMod4]);
    R.diffuseBRDF = c.spec<int>(0, 0);
    R.directMultiscatterBRDF = c.spec<int>(1, 0);
    R.geometricAA = c.specBool(2, false);
    R.indirectLightingTech = c.spec<int>(3, 0);
    R.sunShadowCascadeCount = c.spec<uint32_t>(4, 4);
    R.hasSunSprite = c.exec->pushConstants.size() >= 64;
    if (R.hasSunSprite) {
        float m[16];
        memcpy(m, c.exec->pushConstants.data(), 64);
        R.sunSpriteModel = c.gm4(m);
    }
    const plain_global_shader_info& g = c.g;
    vec3 fwd = c.gv3(g.cameraForward), up = c.gv3(g.cameraUp), right = c.gv3(g.cameraRight);
    if (g_shadeGeometryHook.beginPass) g_shadeGeometryHook.beginPass(c);
    c.forEachInvocation(8, 8, 1, [&](int x, int y, int) {
        if (x >= colorOut.w() || y >= colorOut.h()) return;
        GBufferTexel gbt = decodeGBuffer(R.gbuffer, x, y);
        vec2 uv = vec2((float)x + 0.5f, (float)y + 0.5f) / vec2((float)g.screenResolution[0], (float)g.screenResolution[1]);
        vec2 pixelNDC = uv * 2.f - 1.f;
        vec3 cameraToPixel = -calculateViewDirectionFromPixel(pixelNDC, fwd, up, right, g.cameraTanFovHalf, g.cameraAspectRatio);
        vec3 color = (gbt.depth == 0.f) ? shadeSky(R, x, y, cameraToPixel) : shadeGeometry(R, x, y, gbt, cameraToPixel);
        colorOut.store(x, y, 0, vec4(color, 1.f));
    });
}

}  // namespace orc
