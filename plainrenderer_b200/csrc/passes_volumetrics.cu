// passes_volumetrics.cu - froxel volumetric lighting.
//   froxelVolumeMaterial.comp:17-44, froxelLightScattering.comp:31-64, volumeLightingReprojection.comp:19-62,
//   volumetricLightingIntegration.comp:18-43, volumetricFroxelLighting.inc:1-55
// Froxel volumes are RGBA16F, x fastest: a warp covers 32 consecutive froxels of one row (256 contiguous bytes).
#include "shader_inc.cuh"

namespace pb {

__device__ __forceinline__ vec3 froxelWorldPos(const Globals& G, vec3 uv, float maxDistance) {
    const vec3 ndc = 2.f * (uv - 0.5f);
    const vec3 V = viewDirFromNDC(G, v2(ndc.x, ndc.y));
    return G.camPos - V / dot(-V, G.fwd) * froxelUVToDepth(uv.z, maxDistance);
}

// yBegin / limY: froxel rows [yBegin, limY) of this launch (row sharding: the rank's band of froxel rows + a few rows of overlap)
#define FROXEL_COORDS(vol)                                                                                     \
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = yBegin + blockIdx.y * 4 + (threadIdx.x >> 5), z = blockIdx.z; \
    if (x >= (vol).w || y >= (vol).h || z >= (vol).d || x >= limX || y >= limY || z >= limZ) return;

// ---------------- froxelVolumeMaterial.comp ----------------
__global__ void __launch_bounds__(128) froxelVolumeMaterialKernel(ImgView materialVolume, ImgView noiseTexture, const plain_volumetric_lighting_settings* __restrict__ sp,
                                                                   const plain_global_shader_info* __restrict__ g, int limX, int limY, int limZ, int yBegin) {
    FROXEL_COORDS(materialVolume)
    const plain_volumetric_lighting_settings s = *sp;
    const Globals G = loadGlobals(g);
    const vec3 volumeRes = v3((float)materialVolume.w, (float)materialVolume.h, (float)materialVolume.d);
    const vec3 uv = (v3((float)x, (float)y, (float)z) + 0.5f + s.sampleOffset) / volumeRes;
    const vec3 posWorld = froxelWorldPos(G, uv, s.maxDistance);
    const float noiseScale = 0.5f;
    const vec3 noiseSample = posWorld * noiseScale + ld3(s.windSampleOffset);
    const float noise = sampleLinear3D<WRAP_REPEAT, float>([&](int tx, int ty, int tz) { return loadR8(noiseTexture, tx, ty, tz); }, noiseTexture.w, noiseTexture.h, noiseTexture.d, noiseSample, 0.f);
    vec3 scatteringCoefficient = ld3(s.scatteringCoefficients);
    float absorptionCoefficient = s.absorptionCoefficient;
    float densityMultiplier = s.baseDensity;
    densityMultiplier += s.densityNoiseRange * (noise - 0.5f);
    densityMultiplier = fmaxp(densityMultiplier, 0.f);
    scatteringCoefficient = scatteringCoefficient * densityMultiplier;
    absorptionCoefficient *= densityMultiplier;
    storeRGBA16F(materialVolume, x, y, z, v4(scatteringCoefficient, absorptionCoefficient));
}
PLAIN_PASS(launch_froxelVolumeMaterial, "froxelVolumeMaterial.comp") {
    const ImgView vol = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT), noise = c.sampled(1, PLAIN_FORMAT_R8);
    const plain_volumetric_lighting_settings* s = c.ubuf<plain_volumetric_lighting_settings>(2);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 4, limY = (int)c.exec->dispatch[1] * 4, limZ = (int)c.exec->dispatch[2] * 4;
    int y0, y1;
    c.window(std::min(vol.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(vol.w, 32), ceilDiv((unsigned)(y1 - y0), 4), vol.d);
    PLAIN_LAUNCH(c, froxelVolumeMaterialKernel, grid, 128, 0, vol, noise, s, c.g, limX, y1, limZ, y0);
}

// ---------------- froxelLightScattering.comp ----------------
__global__ void __launch_bounds__(128) froxelLightScatteringKernel(ImgView outVolume, ImgView sunShadowMap, ImgView materialVolume, const plain_shadow_cascade_info* __restrict__ cascades,
                                                                    const plain_light_buffer* __restrict__ light, const plain_volumetric_lighting_settings* __restrict__ sp,
                                                                    const plain_global_shader_info* __restrict__ g, int limX, int limY, int limZ, int yBegin) {
    FROXEL_COORDS(outVolume)
    const plain_volumetric_lighting_settings s = *sp;
    const Globals G = loadGlobals(g);
    const vec3 volumeRes = v3((float)outVolume.w, (float)outVolume.h, (float)outVolume.d);
    const vec3 uv = (v3((float)x, (float)y, (float)z) + 0.5f + s.sampleOffset) / volumeRes;
    const vec3 ndc = 2.f * uv - 1.f;  // :40 (the other froxel passes use 2 * (uv - 0.5))
    const vec3 V = viewDirFromNDC(G, v2(ndc.x, ndc.y));
    const vec3 posWorld = G.camPos - V / dot(-V, G.fwd) * froxelUVToDepth(uv.z, s.maxDistance);
    const float shadow = simpleShadow<false>(posWorld, cascades->lightMatrices[2], sunShadowMap);  // hard-coded cascade 2 (:45)
    const float sunStrength = shadow * light->sunStrengthExposed;
    const vec3 L = v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2]);
    const float VoL = dot(-V, L);
    const float phase = phaseGreenstein(VoL, s.phaseFunctionG);
    const vec4 sa = inRange(materialVolume, x, y, z) ? loadRGBA16F(materialVolume, x, y, z) : v4(0.f);
    const vec3 scatteringCoefficient = xyz(sa);
    const float absorptionCoefficient = sa.w;
    const vec3 constantAmbientLighting = v3(0.02f);
    const vec3 inscattering = (sunStrength * phase * ld3(light->sunColor) + constantAmbientLighting) * scatteringCoefficient;
    const vec3 extinctionCoefficient = scatteringCoefficient + absorptionCoefficient;
    const float transmittance = computeLuminance(extinctionCoefficient);
    storeRGBA16F(outVolume, x, y, z, v4(inscattering, transmittance));
}
PLAIN_PASS(launch_froxelLightScattering, "froxelLightScattering.comp") {
    const ImgView out = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    const ImgView shadow = c.sampled(1, PLAIN_FORMAT_DEPTH16), material = c.sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    const plain_shadow_cascade_info* cascades = c.sbuf<plain_shadow_cascade_info>(3);
    const plain_light_buffer* light = c.sbuf<plain_light_buffer>(4);
    const plain_volumetric_lighting_settings* s = c.ubuf<plain_volumetric_lighting_settings>(5);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 4, limY = (int)c.exec->dispatch[1] * 4, limZ = (int)c.exec->dispatch[2] * 4;
    int y0, y1;
    c.window(std::min(out.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(out.w, 32), ceilDiv((unsigned)(y1 - y0), 4), out.d);
    PLAIN_LAUNCH(c, froxelLightScatteringKernel, grid, 128, 0, out, shadow, material, cascades, light, s, c.g, limX, y1, limZ, y0);
}

// ---------------- volumeLightingReprojection.comp ----------------
__global__ void __launch_bounds__(128) volumeLightingReprojectionKernel(ImgView targetImage, ImgView inputVolume, ImgView historyVolume, const plain_volumetric_lighting_settings* __restrict__ sp,
                                                                         const plain_global_shader_info* __restrict__ g, int limX, int limY, int limZ, int yBegin) {
    FROXEL_COORDS(targetImage)
    const float maxDistance = sp->maxDistance;
    const Globals G = loadGlobals(g);
    const vec4 current = inRange(inputVolume, x, y, z) ? loadRGBA16F(inputVolume, x, y, z) : v4(0.f);
    const vec3 volumeRes = v3((float)targetImage.w, (float)targetImage.h, (float)targetImage.d);
    const vec3 uv = (v3((float)x, (float)y, (float)z) + 0.5f) / volumeRes;
    const vec3 posWorld = froxelWorldPos(G, uv, maxDistance);
    const vec4 ndcPrevious = mulm4(g->viewProjectionPrevious, v4(posWorld, 1.f));
    const vec3 ndcP = xyz(ndcPrevious) / ndcPrevious.w;
    const vec3 camPosPrev = v3(g->cameraPositionPrevious[0], g->cameraPositionPrevious[1], g->cameraPositionPrevious[2]);
    const vec3 V_history = normalize(camPosPrev - posWorld);
    const float historyDistance = length(posWorld - camPosPrev);
    const float historyDepth = historyDistance * dot(-V_history, v3(g->cameraForwardPrevious[0], g->cameraForwardPrevious[1], g->cameraForwardPrevious[2]));
    const vec3 historyUV = v3(ndcP.x * 0.5f + 0.5f, ndcP.y * 0.5f + 0.5f, depthToFroxelUVZ(historyDepth, maxDistance));
    vec4 history = sampleRGBA16FLinearClamp3D(historyVolume, historyUV);
    float alpha = 0.95f;
    if (historyUV.x > 1.f || historyUV.y > 1.f || historyUV.z > 1.f || historyUV.x < 0.f || historyUV.y < 0.f || historyUV.z < 0.f) alpha = 0.f;
    if (g->cameraCut) history = current;
    storeRGBA16F(targetImage, x, y, z, vmix(current, history, alpha));
}
PLAIN_PASS(launch_volumeLightingReprojection, "volumeLightingReprojection.comp") {
    const ImgView target = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    const ImgView input = c.sampled(1, PLAIN_FORMAT_RGBA16_SFLOAT), history = c.sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    const plain_volumetric_lighting_settings* s = c.ubuf<plain_volumetric_lighting_settings>(3);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 4, limY = (int)c.exec->dispatch[1] * 4, limZ = (int)c.exec->dispatch[2] * 4;
    int y0, y1;
    c.window(std::min(target.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(target.w, 32), ceilDiv((unsigned)(y1 - y0), 4), target.d);
    PLAIN_LAUNCH(c, volumeLightingReprojectionKernel, grid, 128, 0, target, input, history, s, c.g, limX, y1, limZ, y0);
}

// ---------------- volumetricLightingIntegration.comp ----------------
// front-to-back scan along z, one thread per froxel column. The reference loops z <= res.z (:28): the extra iteration
// fetches and stores out of range and has no effect.
__global__ void __launch_bounds__(128) volumetricLightingIntegrationKernel(ImgView integrationVolume, ImgView scatteringTransmittanceVolume, const plain_volumetric_lighting_settings* __restrict__ sp,
                                                                            int limX, int limY, int yBegin) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = yBegin + blockIdx.y * 4 + (threadIdx.x >> 5);
    if (x >= integrationVolume.w || y >= integrationVolume.h || x >= limX || y >= limY) return;
    const float maxDistance = sp->maxDistance;
    vec3 inscatteringTotal = v3(0.f);
    float transmittance = 1.f;
    const int resZ = integrationVolume.d;
    float depthStart = froxelUVToDepth(0.f / (float)resZ, maxDistance);
    for (int z = 0; z < resZ; z++) {
        const vec4 it = inRange(scatteringTransmittanceVolume, x, y, z) ? loadRGBA16F(scatteringTransmittanceVolume, x, y, z) : v4(0.f);
        const float depthEnd = froxelUVToDepth((float)(z + 1) / (float)resZ, maxDistance);
        const float segmentLength = depthEnd - depthStart;
        const vec3 inscattering = integrateInscattering(xyz(it), v3(it.w), segmentLength);
        inscatteringTotal = inscatteringTotal + inscattering;
        transmittance *= dm::exp(-it.w * segmentLength);
        storeRGBA16F(integrationVolume, x, y, z, v4(inscatteringTotal, transmittance));
        depthStart = depthEnd;  // froxelUVToDepth(z / resZ) of the next slice is the same expression as this slice's end
    }
}
PLAIN_PASS(launch_volumetricLightingIntegration, "volumetricLightingIntegration.comp") {
    const ImgView integration = c.storage(0, PLAIN_FORMAT_RGBA16_SFLOAT), src = c.sampled(1, PLAIN_FORMAT_RGBA16_SFLOAT);
    const plain_volumetric_lighting_settings* s = c.ubuf<plain_volumetric_lighting_settings>(2);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    int y0, y1;
    c.window(std::min(integration.h, limY), y0, y1);
    if (y1 <= y0) return;
    dim3 grid(ceilDiv(integration.w, 32), ceilDiv((unsigned)(y1 - y0), 4));
    PLAIN_LAUNCH(c, volumetricLightingIntegrationKernel, grid, 128, 0, integration, src, s, limX, y1, y0);
}

}  // namespace pb
