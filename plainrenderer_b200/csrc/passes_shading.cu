// passes_shading.cu - deferred recast of the forward shading pass over the packed G-buffer, plus brdfLut.comp.
//   "gbufferShading.comp" = triangle.frag:76-341 (fragment stage) + sky.frag:20-33 + sunSprite.frag:23-43 as one
//   full-screen kernel; brdfLut.comp:20-101.
//
// G-buffer texel (include/plain_frame_types.h), one 128-bit load per pixel:
//   x depth bits | y shading normal, octahedral 2 x SNORM16 | z albedo sRGB8 + roughness << 24 | w metalness
// Thread mapping: a warp shades a 16x2 pixel tile (lane = (y & 1) * 16 + x), so the 2x2 quads that stand in for
// dFdxFine/dFdyFine (GeometricAA.inc:10-11) live inside one warp: the horizontal partner is lane ^ 1, the vertical
// partner lane ^ 16, and the normal differences come from two shuffles instead of three more G-buffer loads.
#include "shader_inc.cuh"

namespace pb {

__global__ void buildShadingTablesKernel(ShadingTables* t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 256) {
        t->unorm8[i] = unorm8((uint32_t)i);
        t->srgbToLinear[i] = sRGBToLinear1(unorm8((uint32_t)i));
    }
    if (i < 256 * 12) {
        const float noise = unorm8((uint32_t)(i / 12));
        const int tap = i % 12;
        const float sampleCount = 12.f;
        float d = ((float)tap + 0.5f * noise) / sampleCount;
        d = sqrtf_(d);
        const float angle = noise * 2.f * PV_PI + 2.f * PV_PI * (float)tap / sampleCount;
        t->pcf[i] = make_float4(dm::cos(angle), dm::sin(angle), d, 0.f);
    }
    if (i < PLAIN_DISC_SEEDS) {  // one thread per seed: the xorshift sequence is serial
        uint32_t rngState = wang_hash((uint32_t)i);
        for (int k = 0; k < 32; k++) {
            const float sq = sqrtf_(rand01(rngState));
            const float angle = 2.f * PV_PI * rand01(rngState);
            t->disc[i * 32 + k] = make_float4(sq, dm::cos(angle), dm::sin(angle), 0.f);
        }
    }
}
void buildShadingTables(ShadingTables* deviceTables, cudaStream_t stream) { buildShadingTablesKernel<<<12, 256, 0, stream>>>(deviceTables); }

struct ShadingParams {
    ImgView gbuffer, brdfLut, shadowMaps[4], ySH, coCg, volumetricLUT, skyLut, transmissionLut, colorOut;
    const plain_light_buffer* light;
    const plain_shadow_cascade_info* cascades;
    const plain_volumetric_lighting_settings* vol;
    const plain_global_shader_info* g;
    const BindlessEntry* bindless;
    const ShadingTables* tables;
    int diffuseBRDF, directMultiscatterBRDF, geometricAA, indirectLightingTech;
    uint32_t sunShadowCascadeCount;
    float sunSpriteModel[16];
    int hasSunSprite;
    UpscaleSource upscale;  // FUSED: the inputs of the indirectLightUpscale.comp execution this kernel absorbed (ySH / coCg are then only extents)
};

__device__ __forceinline__ vec3 decodeOctNormal(uint32_t packed) {
    const vec2 f = v2(snorm16((int16_t)(packed & 0xffffu)), snorm16((int16_t)(packed >> 16)));
    vec3 n = v3(f.x, f.y, 1.f - absf(f.x) - absf(f.y));
    const float tt = fmaxp(-n.z, 0.f);
    n.x += (n.x >= 0.f) ? -tt : tt;
    n.y += (n.y >= 0.f) ? -tt : tt;
    return normalize(n);
}

__device__ __forceinline__ float ReflectedEnergyAverage(float roughness) {  // triangle.frag:123-131
    const float smoothness = 1.f - sqrtf_(roughness);
    float r = -0.0761947f - 0.383026f * smoothness;
    r = 1.04997f + smoothness * r;
    r = 0.409255f + smoothness * r;
    return fminp(0.999f, r);
}

template <int MULTI>
__device__ __forceinline__ vec3 computeSpecularMultiscatteringLobe(const ShadingParams& p, float r, float NoL, vec3 f0, vec3 singleScatteringLobe, vec3 brdfLut) {  // :146-175
    const int mode = MULTI >= 0 ? MULTI : p.directMultiscatterBRDF;
    const float energyOutgoing = brdfLut.y;
    const vec3 fresnelAverage = f0 + (1.f - f0) / 21.f;
    if (mode == 0) {
        const float energyAverage = ReflectedEnergyAverage(r);
        const float energyIncoming = sampleRGBA16FLinearClamp(p.brdfLut, v2(r, NoL)).y;
        const float multiScatteringLobeUnscaled = (1.f - energyIncoming) * (1.f - energyOutgoing) / (3.1415f * (1.f - energyAverage));
        const vec3 multiScatteringScaling = (fresnelAverage * fresnelAverage * energyAverage) / (1.f - fresnelAverage * (1.f - energyAverage));
        return multiScatteringLobeUnscaled * multiScatteringScaling;
    } else if (mode == 1) {
        vec3 multiScatteringLobe = v3((1.f - energyOutgoing) / PV_PI);
        const vec3 multiScatteringScaling = (fresnelAverage * fresnelAverage * energyOutgoing) / (1.f - fresnelAverage * (1.f - energyOutgoing));
        return multiScatteringLobe * multiScatteringScaling;
    } else if (mode == 2) {
        return f0 * (1.f / energyOutgoing - 1.f) * singleScatteringLobe;
    }
    return v3(0.f);
}

// triangle.frag:92-120, 12-tap spiral PCF with a per-pixel blue-noise rotation. cos/sin/sqrt of tap i depend only on the
// 8-bit noise texel and i: they come from the exact table (ShadingTables::pcf).
__device__ __forceinline__ float calcShadow(const ImgView& shadowMap, const float* lightMatrix, vec2 lightSpaceScale, vec3 pos, const float4* __restrict__ pcfRow) {
    vec4 posLightSpace = mulm4(lightMatrix, v4(pos, 1.f));
    posLightSpace = posLightSpace / posLightSpace.w;
    const vec2 plsXY = v2(posLightSpace.x, posLightSpace.y) * 0.5f + 0.5f;
    const float actualDepth = clampf(posLightSpace.z, 0.f, 1.f);
    const vec2 offsetScale = PB_SHADOW_SAMPLE_RADIUS * lightSpaceScale;
    float shadow = 0.f;
    const float sampleCount = 12.f;
    // shadowTest (:84-87) compares actualDepth with the D16 texel decoded as float(code) / 65535. The decode is monotonic in the code, so
    // the twelve comparisons are integer comparisons against ONE threshold: the largest code whose decoded value does not exceed
    // actualDepth (found among three candidates with the same IEEE division: truncation of actualDepth * 65535 is off by at most one).
    // actualDepth is in [0, 1] (a NaN was clamped to 0), so the black border (0) always passes.
    const int c0 = (int)(actualDepth * 65535.f);
    const int c1 = imin(c0 + 1, 65535);
    const int codeMax = ((float)c1 / 65535.f <= actualDepth) ? c1 : (((float)c0 / 65535.f <= actualDepth) ? c0 : c0 - 1);
    const uint16_t* texels = (const uint16_t*)shadowMap.ptr;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const float4 e = __ldg(pcfRow + i);  // {cos(angle), sin(angle), sqrt(d)}
        vec2 offset = v2(e.x, e.y);
        offset = offset * (offsetScale * e.z);
        const vec2 samplePosition = plsXY + offset;
        const ivec2 t = nearestTexelUnsanitized(samplePosition, shadowMap.w, shadowMap.h);  // nearest + border sampler (image_view.h)
        const bool inside = (unsigned)t.x < (unsigned)shadowMap.w && (unsigned)t.y < (unsigned)shadowMap.h;
        const bool lit = inside ? ((int)ldg(texels + t.y * shadowMap.w + t.x) <= codeMax) : true;
        shadow += lit ? 1.f : 0.f;
    }
    return shadow / sampleCount;
}

template <int DIFFUSE, int MULTI, int GEOAA, int TECH, bool FUSED>
__device__ __forceinline__ vec3 shadeGeometry(const ShadingParams& p, const Globals& G, int x, int y, uint4 texel, vec3 N, vec3 N_U, vec3 N_V, vec3 cameraToPixel, uint32_t noiseBytes) {
    const plain_global_shader_info* g = p.g;
    const vec2 fragCoord = v2((float)x + 0.5f, (float)y + 0.5f);
    const float depth = dm::u2f(texel.x);
    const float depthLinear = linearizeDepth(depth, G.nearPlane, G.farPlane);
    const vec3 passPos = G.camPos + cameraToPixel / dot(cameraToPixel, G.fwd) * depthLinear;

    // triangle.frag:184-193
    const ShadingTables* T = p.tables;
    const float metalic = __ldg(&T->unorm8[texel.w & 0xffu]);
    float r = __ldg(&T->unorm8[texel.z >> 24]);
    r = fmaxp(r * r, 0.0045f);
    const vec3 albedo = v3(__ldg(&T->srgbToLinear[texel.z & 0xffu]), __ldg(&T->srgbToLinear[(texel.z >> 8) & 0xffu]), __ldg(&T->srgbToLinear[(texel.z >> 16) & 0xffu]));
    const vec2 noiseTexel = v2(__ldg(&T->unorm8[noiseBytes & 0xffu]), __ldg(&T->unorm8[noiseBytes >> 8]));
    const vec3 diffuseColor = (1.f - metalic) * albedo;
    const vec3 L = normalize(v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2]));
    vec3 V = G.camPos - passPos;
    const float pixelDepth = dot(V, -G.fwd);
    V = normalize(V);
    const vec3 H = normalize(V + L);
    if (GEOAA >= 0 ? GEOAA != 0 : p.geometricAA != 0) {  // GeometricAA.inc:4-20
        const float kappa = 0.18f, pixelVariance = 0.5f;
        const float pxVar2 = pixelVariance * pixelVariance;
        const float lengthN_U2 = dot(N_U, N_U), lengthN_V2 = dot(N_V, N_V);
        const float variance = pxVar2 * (lengthN_V2 + lengthN_U2);
        const float kernelRoughness2 = fminp(2.f * variance, kappa);
        r = clampf(sqrtf_(r * r + kernelRoughness2), 0.f, 1.f);
    }
    const float NoH = fmaxp(dot(N, H), 0.f);
    const float NoL = clampf(dot(N, L), 0.f, 1.f);
    const float VoH = absf(dot(V, H));
    const float LoV = fmaxp(dot(L, V), 0.f);
    float NoV = absf(dot(N, V));
    NoV = fmaxp(NoV, 0.0001f);
    const vec3 f0 = vmix(v3(0.04f), albedo, metalic);

    // sun light, :222-241
    int cascadeIndex = 0;
    for (uint32_t cascade = 0; cascade + 1 < p.sunShadowCascadeCount; cascade++) cascadeIndex += (pixelDepth >= p.cascades->splits[cascade]) ? 1 : 0;
    const float sunShadow = calcShadow(p.shadowMaps[cascadeIndex], p.cascades->lightMatrices[cascadeIndex],
                                       v2(p.cascades->lightSpaceScale[cascadeIndex][0], p.cascades->lightSpaceScale[cascadeIndex][1]), passPos, T->pcf + (noiseBytes & 0xffu) * 12);
    const float sunStrengthExposed = p.light->sunStrengthExposed;
    const vec3 sunColor = ld3(p.light->sunColor);
    const vec3 directLighting = fmaxp(dot(N, L), 0.f) * sunShadow * sunColor;
    const vec3 brdfLut = xyz(sampleRGBA16FLinearClamp(p.brdfLut, v2(r, NoV)));

    // direct diffuse, :243-285
    const int diffuseMode = DIFFUSE >= 0 ? DIFFUSE : p.diffuseBRDF;
    vec3 diffuseDirect;
    vec3 diffuseBRDFIntegral = v3(1.f);
    if (diffuseMode == 0) {
        diffuseDirect = diffuseColor / PV_PI * directLighting;
        diffuseBRDFIntegral = v3(brdfLut.z);
    } else if (diffuseMode == 1) {
        diffuseDirect = DisneyDiffuse(diffuseColor, NoL, VoH, NoV, r) * directLighting;
        diffuseBRDFIntegral = v3(brdfLut.z);
    } else if (diffuseMode == 2) {
        diffuseDirect = CoDWWIIDiffuse(diffuseColor, NoL, VoH, NoV, NoH, r) * directLighting;
        diffuseBRDFIntegral = v3(brdfLut.z);
    } else {
        diffuseDirect = Titanfall2Diffuse(diffuseColor, NoL, LoV, NoV, NoH, r) * directLighting;
        float multiIntegral = 0.1159f * r * PV_PI * 2.f;
        multiIntegral *= (1.f - F_Schlick(v3(0.04f), v3(1.f), NoV).x);
        multiIntegral *= 0.94291f;
        diffuseBRDFIntegral = vmin(v3(brdfLut.z) + diffuseColor * multiIntegral, v3(1.f));
    }
    diffuseDirect = diffuseDirect * ((1.f - F_Schlick(f0, v3(1.f), NoV)) * (1.f - (F_Schlick(f0, v3(1.f), NoL))));

    // direct specular, :287-290
    const vec3 singleScatteringLobe = GGXSingleScattering(r, f0, NoH, NoV, VoH, NoL);
    const vec3 multiScatteringLobe = computeSpecularMultiscatteringLobe<MULTI>(p, r, NoL, f0, singleScatteringLobe, brdfLut);
    const vec3 specularDirect = directLighting * (singleScatteringLobe + multiScatteringLobe);

    vec3 lightingIndirect;
    const vec2 screenRes = v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    if ((TECH >= 0 ? TECH : p.indirectLightingTech) == 0) {  // :295-322
        const vec2 screenUV = fragCoord / screenRes;
        vec4 irradiance_Y_SH;
        vec2 irradiance_CoCg;
        if (FUSED) {
            // the upscale pass folded into its consumer (indirectLightUpscale.comp:17-71 -> triangle.frag:294-320): the nearest texel the two
            // lookups below would have fetched is computed here for that texel, and rounded through binary16 exactly as its store +
            // load would have done. 28 bytes per pixel (12 written, 12 + 4 read) and one launch less; the launcher checked that the
            // full-resolution GI images have the colour target's extent, so both lookups address the same texel
            ivec2 zero; zero.x = 0; zero.y = 0;
            const ivec2 t = sampleNearest2D<WRAP_CLAMP, ivec2>([&](int tx, int ty) { ivec2 r; r.x = tx; r.y = ty; return r; }, p.ySH.w, p.ySH.h, screenUV, zero);  // the sampler's own texel selection
            vec4 up_Y_SH;
            vec2 up_CoCg;
            giUpscalePixel(p.upscale, g, t.x, t.y, up_Y_SH, up_CoCg);
            irradiance_Y_SH = v4(halfToFloat(floatToHalf(up_Y_SH.x)), halfToFloat(floatToHalf(up_Y_SH.y)), halfToFloat(floatToHalf(up_Y_SH.z)), halfToFloat(floatToHalf(up_Y_SH.w)));
            irradiance_CoCg = v2(halfToFloat(floatToHalf(up_CoCg.x)), halfToFloat(floatToHalf(up_CoCg.y)));
        } else {
            irradiance_Y_SH = sampleNearest2D<WRAP_CLAMP, vec4>([&](int tx, int ty) { return loadRGBA16F(p.ySH, tx, ty); }, p.ySH.w, p.ySH.h, screenUV, v4(0.f));
            irradiance_CoCg = sampleNearest2D<WRAP_CLAMP, vec2>([&](int tx, int ty) { return loadRG16F(p.coCg, tx, ty); }, p.coCg.w, p.coCg.h, screenUV, v2(0.f));
        }
        const float irradiance_Y = dot(irradiance_Y_SH, directionToSH_L1(N));
        const vec3 irradiance = YCoCgToLinear(v3(irradiance_Y, irradiance_CoCg.x, irradiance_CoCg.y));
        const vec3 diffuseIndirect = irradiance * diffuseColor * diffuseBRDFIntegral;

        const vec3 dominantDirection = dominantDirectionFromSH_L1(irradiance_Y_SH);
        float dominantDirectionLength = length(dominantDirection);
        dominantDirectionLength = clampf(dominantDirectionLength, 0.01f, 1.f);
        const float r_indirect = mixf(1.f, r, sqrtf_(dominantDirectionLength));
        const vec3 L_indirect = dominantDirection / dominantDirectionLength;
        const vec3 H_indirect = normalize(L_indirect + V);
        const float NoH_indirect = fmaxp(dot(N, H_indirect), 0.f);
        const float NoL_indirect = fmaxp(dot(N, L_indirect), 0.f);
        const float VoH_indirect = fmaxp(dot(V, H_indirect), 0.f);
        const vec3 singleScattering_indirect = GGXSingleScattering(r_indirect, f0, NoH_indirect, NoV, VoH_indirect, NoL_indirect);
        const vec3 multiScattering_indirect = computeSpecularMultiscatteringLobe<MULTI>(p, r_indirect, NoL_indirect, f0, singleScattering_indirect, brdfLut);
        const vec3 specularIndirect = (singleScattering_indirect + multiScattering_indirect) * YCoCgToLinear(v3(irradiance_Y_SH.x, irradiance_CoCg.x, irradiance_CoCg.y));
        lightingIndirect = diffuseIndirect + specularIndirect;
    } else {  // constant ambient, :324-333
        const float ambientStrength = 0.003f;
        const vec3 irradiance = v3(ambientStrength) * sunStrengthExposed;
        const vec3 reflection = v3(ambientStrength) * sunStrengthExposed;
        const vec3 singleScattering = vmix(v3(brdfLut.x), v3(brdfLut.y), f0);
        const vec3 diffuseIndirect = irradiance * diffuseColor * diffuseBRDFIntegral;
        const vec3 specularIndirect = singleScattering * reflection;
        lightingIndirect = diffuseIndirect + specularIndirect;
    }
    const vec3 color = (diffuseDirect + specularDirect) * sunStrengthExposed + lightingIndirect;

    // applyVolumetricLighting, :133-144
    vec2 noise = noiseTexel;
    noise = noise - 0.5f;
    noise = noise * 0.013f;
    vec2 screenUV = fragCoord / screenRes;
    screenUV = screenUV + noise;
    const vec4 inscatteringTransmittance = volumeTextureLookup(screenUV, pixelDepth, p.volumetricLUT, p.vol->maxDistance);
    return applyInscatteringTransmittance(color, inscatteringTransmittance);
}

// sky.frag:20-33, then sunSprite.frag:23-43 blended additively onto the stored R11G11B10 value (RenderPass.cpp:112-124)
__device__ __forceinline__ vec3 shadeSky(const ShadingParams& p, int x, int y, vec3 cameraToPixel) {
    const plain_global_shader_info* g = p.g;
    const vec2 fragCoord = v2((float)x + 0.5f, (float)y + 0.5f);
    const vec2 res = v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    const vec3 V = cameraToPixel;
    vec3 color = sampleSkyLut(V, p.skyLut);
    const vec2 d = fragCoord * res;
    color = ditherRGB8(color, f2i(d.x), f2i(d.y), g->time);
    const vec2 screenUV = fragCoord / res;
    const vec4 inscatteringTransmittance = volumeTextureLookup(screenUV, PB_MAX_VOLUMETRIC_LIGHTING_DEPTH, p.volumetricLUT, p.vol->maxDistance);
    color = applyInscatteringTransmittance(color, inscatteringTransmittance);
    color = unpackR11G11B10(packR11G11B10(color));  // render-target store before the blend

    if (p.hasSunSprite) {
        // invert mat3(model) = Rot * diag(s, s, 1): the point of the sprite quad that projects onto this pixel's ray
        const float* M = p.sunSpriteModel;
        const vec3 c0 = v3(M[0], M[1], M[2]), c1 = v3(M[4], M[5], M[6]), c2 = v3(M[8], M[9], M[10]);
        const float s2 = dot(c0, c0);
        const vec3 q = v3(dot(c0, V) / s2, dot(c1, V) / s2, dot(c2, V));
        if (q.z < 0.f) {
            const vec2 posCentered = v2(q.x / -q.z, q.y / -q.z);
            if (absf(posCentered.x) <= 1.f && absf(posCentered.y) <= 1.f) {
                const float distanceFromCenter = dot(posCentered, posCentered);
                if (!(distanceFromCenter > 1.f)) {
                    const vec3 passWorldPos = c0 * posCentered.x + c1 * posCentered.y + c2 * -1.f;
                    const float bias = 0.002f;
                    const vec3 Vs = normalize(passWorldPos + v3(0.f, bias, 0.f));
                    const vec2 lutUV = computeLutUV(0.f, 100.f, v3(0.f, -1.f, 0.f), Vs);
                    const vec3 transmission = sampleR11LinearClamp(p.transmissionLut, lutUV);
                    const float mu = sqrtf_(1.f - distanceFromCenter);
                    const vec3 limb = vpow(v3(mu), v3(0.482f, 0.511f, 0.643f));
                    const vec3 sun = p.light->sunStrengthExposed * transmission * limb;
                    float alpha = 1.f - distanceFromCenter;
                    alpha *= alpha;
                    color = sun * alpha + color;
                }
            }
        }
    }
    return color;
}

template <int DIFFUSE, int MULTI, int GEOAA, int TECH, bool FUSED = false>
__global__ void __launch_bounds__(256, 4) gbufferShadingKernel(const __grid_constant__ ShadingParams p, int limitX, int limitY, int y0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + (warp & 1) * 16 + (lane & 15);
    const int y = y0 + blockIdx.y * 8 + (warp >> 1) * 2 + (lane >> 4);  // y0: first row of the launch (even; row sharding)
    const int W = p.colorOut.w, Hh = p.colorOut.h;
    const bool inImage = x < W && y < Hh;                        // quad partners are loaded even when they lie outside the row window
    const bool inside = inImage && x < limitX && y < limitY;
    uint4 texel = make_uint4(0, 0, 0, 0);
    if (inImage) texel = loadU4(p.gbuffer, x, y);
    const vec3 N = decodeOctNormal(texel.y);
    // quad partners: dFdx = right - left within the pixel pair, dFdy = bottom - top; a partner outside the image is the pixel itself
    const bool rightOk = ((x | 1) < W), bottomOk = ((y | 1) < Hh);
    const float nxo = __shfl_xor_sync(0xffffffffu, N.x, 1), nyo = __shfl_xor_sync(0xffffffffu, N.y, 1), nzo = __shfl_xor_sync(0xffffffffu, N.z, 1);
    const float nxv = __shfl_xor_sync(0xffffffffu, N.x, 16), nyv = __shfl_xor_sync(0xffffffffu, N.y, 16), nzv = __shfl_xor_sync(0xffffffffu, N.z, 16);
    if (!inside) return;
    vec3 N_U = v3(0.f), N_V = v3(0.f);
    if (rightOk) N_U = (x & 1) ? (N - v3(nxo, nyo, nzo)) : (v3(nxo, nyo, nzo) - N);
    if (bottomOk) N_V = (y & 1) ? (N - v3(nxv, nyv, nzv)) : (v3(nxv, nyv, nzv) - N);
    if (!rightOk) N_U = N - N;
    if (!bottomOk) N_V = N - N;

    const plain_global_shader_info* g = p.g;
    const Globals G = loadGlobals(g);
    const vec2 uv = v2((float)x + 0.5f, (float)y + 0.5f) / v2((float)g->screenResolution[0], (float)g->screenResolution[1]);
    const vec2 pixelNDC = uv * 2.f - 1.f;
    const vec3 cameraToPixel = -viewDirFromNDC(G, pixelNDC);
    vec3 color;
    if (dm::u2f(texel.x) == 0.f) {
        color = shadeSky(p, x, y, cameraToPixel);
    } else {
        // blue noise of this frame (global.inc noiseTextureIndices), nearest + repeat at fragCoord / textureSize
        const ImgView noiseTex = p.bindless[g->noiseTextureIndices[g->frameIndexMod4]].view;
        const vec2 noiseUV = v2((float)x + 0.5f, (float)y + 0.5f) / v2((float)noiseTex.w, (float)noiseTex.h);
        const uint32_t noiseBytes = sampleNearest2D<WRAP_REPEAT, uint32_t>([&](int tx, int ty) { return (uint32_t)ldg((const uint16_t*)noiseTex.ptr + texelIndex(noiseTex, tx, ty)); }, noiseTex.w, noiseTex.h, noiseUV, 0u);
        color = shadeGeometry<DIFFUSE, MULTI, GEOAA, TECH, FUSED>(p, G, x, y, texel, N, N_U, N_V, cameraToPixel, noiseBytes);
    }
    storeR11(p.colorOut, x, y, color);
}

PLAIN_PASS(launch_gbufferShading, "gbufferShading.comp") {
    ShadingParams p;
    p.gbuffer = c.sampled(0, PLAIN_FORMAT_RGBA32_UINT);
    p.brdfLut = c.sampled(3, PLAIN_FORMAT_RGBA16_SFLOAT);
    for (int i = 0; i < 4; i++) p.shadowMaps[i] = c.sampled(9 + i, PLAIN_FORMAT_DEPTH16);
    p.ySH = c.sampled(15, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.coCg = c.sampled(16, PLAIN_FORMAT_RG16_SFLOAT);
    p.volumetricLUT = c.sampled(18, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.skyLut = c.sampled(21, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.transmissionLut = c.sampled(22, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.colorOut = c.storage(20, PLAIN_FORMAT_R11G11B10_UFLOAT);
    p.light = c.sbuf<plain_light_buffer>(7);
    p.cascades = c.sbuf<plain_shadow_cascade_info>(8);
    p.vol = c.ubuf<plain_volumetric_lighting_settings>(19);
    p.g = c.g;
    p.bindless = c.bindless;
    p.tables = (const ShadingTables*)c.tables;
    p.diffuseBRDF = c.spec<int>(0, 0);
    p.directMultiscatterBRDF = c.spec<int>(1, 0);
    p.geometricAA = c.specBool(2, false) ? 1 : 0;
    p.indirectLightingTech = c.spec<int>(3, 0);
    p.sunShadowCascadeCount = c.spec<uint32_t>(4, 4);
    p.hasSunSprite = c.exec->pushConstants.size() >= 64 ? 1 : 0;
    memset(p.sunSpriteModel, 0, sizeof(p.sunSpriteModel));
    if (p.hasSunSprite) memcpy(p.sunSpriteModel, c.exec->pushConstants.data(), 64);
    if (c.failed) return;
    if (p.gbuffer.w != p.colorOut.w || p.gbuffer.h != p.colorOut.h) { c.fail("gbufferShading.comp: G-buffer and colour target extents differ"); return; }
    if (p.sunShadowCascadeCount < 1 || p.sunShadowCascadeCount > 4) { c.fail("gbufferShading.comp: cascade count must be 1..4"); return; }
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    int y0, y1;
    c.window(std::min(p.colorOut.h, limY), y0, y1);
    if (y0 % 2 != 0) { c.fail("gbufferShading.comp: row window must start at an even row (2x2 quads)"); return; }
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(p.colorOut.w, 32), ceilDiv((unsigned)(y1 - y0), 8));
    // the backend folded the execution that produces ySH / coCg (indirectLightUpscale.comp) into this one: its inputs travel in p.upscale
    bool fused = false;
    if (c.exec->fusedProducer >= 0 && p.indirectLightingTech == 0 && p.ySH.w == p.colorOut.w && p.ySH.h == p.colorOut.h && p.coCg.w == p.colorOut.w && p.coCg.h == p.colorOut.h) {
        LaunchCtx u = c;
        u.exec = c.be_exec(c.exec->fusedProducer);
        u.pass = c.be_pass(u.exec->pass);
        p.upscale.srcYSH = u.sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
        p.upscale.srcCoCg = u.sampled(3, PLAIN_FORMAT_RG16_SFLOAT);
        p.upscale.fullResDepth = u.sampled(4, PLAIN_FORMAT_DEPTH32);
        p.upscale.halfResDepth = u.sampled(5, PLAIN_FORMAT_R16_SFLOAT);
        if (u.failed) { c.fail(u.error); return; }
        fused = true;
    } else if (c.exec->fusedProducer >= 0) { c.fail("gbufferShading.comp: an upscale execution was folded into a configuration that does not read it"); return; }
    if (fused && p.diffuseBRDF == 2 && p.directMultiscatterBRDF == 0 && p.geometricAA)
        PLAIN_LAUNCH(c, (gbufferShadingKernel<2, 0, 1, 0, true>), grid, 256, 0, p, limX, y1, y0);
    else if (fused)
        PLAIN_LAUNCH(c, (gbufferShadingKernel<-1, -1, -1, 0, true>), grid, 256, 0, p, limX, y1, y0);
    else if (p.diffuseBRDF == 2 && p.directMultiscatterBRDF == 0 && p.geometricAA && p.indirectLightingTech == 0)
        PLAIN_LAUNCH(c, (gbufferShadingKernel<2, 0, 1, 0>), grid, 256, 0, p, limX, y1, y0);
    else if (p.diffuseBRDF == 2 && p.directMultiscatterBRDF == 0 && p.geometricAA && p.indirectLightingTech == 1)
        PLAIN_LAUNCH(c, (gbufferShadingKernel<2, 0, 1, 1>), grid, 256, 0, p, limX, y1, y0);
    else
        PLAIN_LAUNCH(c, (gbufferShadingKernel<-1, -1, -1, -1>), grid, 256, 0, p, limX, y1, y0);
}

// ---------------- brdfLut.comp:20-101 (startup / on diffuse-BRDF change) ----------------
__global__ void __launch_bounds__(64) brdfLutKernel(ImgView lut, int diffuseBRDF, int limitX, int limitY) {
    const int ux = blockIdx.x * 8 + threadIdx.x, uy = blockIdx.y * 8 + threadIdx.y;
    if (ux >= lut.w || uy >= lut.h || ux >= limitX || uy >= limitY) return;
    float r = (float)ux / (float)lut.w;
    r = fmaxp(r, 0.0001f);
    const float NoV = fmaxp((float)uy, 0.1f) / (float)lut.h;
    const vec3 V = v3(sqrtf_(1.0f - NoV * NoV), 0.f, NoV);
    const vec3 N = v3(0.f, 0.f, 1.f);
    const int samples = 1024;
    vec3 result = v3(0.f);
    for (int i = 0; i < samples; i++) {
        const vec2 xi = hammersley2d((uint32_t)i, (uint32_t)samples);
        {  // specular
            const vec3 H = importanceSampleGGX(xi, r, N);
            const vec3 L = 2.f * dot(V, H) * H - V;
            const float VoH = fmaxp(dot(V, H), 0.f);
            const float NoH = fmaxp(H.z, 0.f);
            const float NoL = fmaxp(L.z, 0.f);
            if (NoL > 0.f) {
                const float F_c = dm::pow(1.f - VoH, 5.f);
                const float Vis = Visibility(NoV, NoL, r);
                const float k = Vis * VoH * NoL / NoH;
                result.x += F_c * k;
                result.y += k;
            }
        }
        {  // diffuse
            const vec3 L = importanceSampleCosine(xi, N);
            const vec3 H = normalize(V + L);
            const float VoH = clampf(dot(V, H), 0.f, 1.f);
            const float NoL = fmaxp(L.z, 0.f);
            const float NoH = fmaxp(H.z, 0.f);
            const vec3 F0Diffuse = v3(0.04f);
            const float fresnelInOut = (1.f - F_Schlick(F0Diffuse, v3(1.f), NoV).x) * (1.f - F_Schlick(F0Diffuse, v3(1.f), NoL).x);
            if (diffuseBRDF == 0) {
                result.z += (1.f / PV_PI) * fresnelInOut;
            } else if (diffuseBRDF == 1) {
                result.z += DisneyDiffuse(v3(1.f), NoL, VoH, NoV, r).x * fresnelInOut;
            } else if (diffuseBRDF == 2) {
                result.z += CoDWWIIDiffuse(v3(1.f), NoL, VoH, NoV, NoH, r).x * fresnelInOut;
            } else if (diffuseBRDF == 3) {
                const float LoV = clampf(dot(L, V), 0.f, 1.f);
                result.z += Titanfall2DiffuseSingleComponent(NoL, LoV, NoV, NoH, r) * fresnelInOut;
            }
        }
    }
    result = result / (float)samples;
    result.x *= 4.f;
    result.y *= 4.f;
    storeRGBA16F(lut, ux, uy, 0, v4(result, 0.f));
}
PLAIN_PASS(launch_brdfLut, "brdfLut.comp") {
    const ImgView lut = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    if (c.failed) return;
    dim3 grid(c.exec->dispatch[0], c.exec->dispatch[1]);
    PLAIN_LAUNCH(c, brdfLutKernel, grid, dim3(8, 8), 0, lut, c.spec<int>(0, 0), (int)grid.x * 8, (int)grid.y * 8);
}

}  // namespace pb
