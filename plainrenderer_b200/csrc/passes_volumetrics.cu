// passes_volumetrics.cu - froxel volumetric lighting.
//   froxelVolumeMaterial.comp:17-44, froxelLightScattering.comp:31-64, volumeLightingReprojection.comp:19-62,
//   volumetricLightingIntegration.comp:18-43, volumetricFroxelLighting.inc:1-55
// Froxel volumes are RGBA16F, x fastest: a warp covers 32 consecutive froxels of one row (256 contiguous bytes).
// The per-froxel bodies live in froxel_inc.cuh (host+device); this file holds the four per-pass kernels and the fused column
// kernel the backend launches instead of them when the four executions form one chain (backend.cu planFusions).
#include <cstdlib>
#include "froxel_inc.cuh"

namespace pb {

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
    storeRGBA16F(materialVolume, x, y, z, froxelMaterialAt(s, noiseTexture, posWorld));
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
    const vec4 sa = inRange(materialVolume, x, y, z) ? loadRGBA16F(materialVolume, x, y, z) : v4(0.f);
    storeRGBA16F(outVolume, x, y, z, froxelScatteringAt(g, s, cascades, light, sunShadowMap, V, posWorld, sa));
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
    storeRGBA16F(targetImage, x, y, z, froxelReprojectionAt(g, maxDistance, historyVolume, posWorld, current));
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
        const vec4 terms = froxelSegmentTerms(it, segmentLength);
        inscatteringTotal = inscatteringTotal + xyz(terms);
        transmittance *= terms.w;
        storeRGBA16F(integrationVolume, x, y, z, v4(inscatteringTotal, transmittance));
        depthStart = depthEnd;  // froxelUVToDepth(z / resZ) of the next slice is the same expression as this slice's end
    }
}

// ---------------- the four passes as ONE launch (backend.cu planFusions hands the chain over as ExecRecord::fusedRun) ----------------
// Block = 8 columns x ZLANES z lanes of one froxel row (16 by default: 128 threads); the phases are in froxel_inc.cuh (froxelBlockPrologue / Phase1 / Phase2 / Phase3), one call
// per thread and phase with a barrier in between.
// Algorithmic bytes per froxel: history read 8 + history write 8 + integrated write 8 = 24 (the four passes: 8 + 16 + 24 + 16 = 64); the material
// and scattering volumes are not written (tests that compare them run unfused: plain_set_pass_fusion_enabled).
template <int ZLANES> __global__ void __launch_bounds__(FROXEL_COLS * ZLANES, 1024 / (FROXEL_COLS * ZLANES)) froxelColumnKernel(const __grid_constant__ FroxelFusedParams p) {
    __shared__ FroxelBlockShared sh;
    const int tid = threadIdx.x, blockX = blockIdx.x, y = p.yBegin + blockIdx.y;
    const plain_volumetric_lighting_settings s = *p.settings;
    const Globals G = loadGlobals(p.in.g);
    froxelBlockPrologue<ZLANES>(sh, p, G, s, tid, blockX, y);
    __syncthreads();
    froxelBlockPhase1<ZLANES, false>(sh, p, G, s, tid, blockX, y);
    __syncthreads();
    froxelBlockPhase2(sh, p, tid, blockX);
    __syncthreads();
    froxelBlockPhase3<ZLANES>(sh, p, tid, blockX, y);
}
// c = the chain's last execution (volumetricLightingIntegration.comp); c.exec->fusedRun = material, scattering, reprojection, integration
static void launchFroxelColumns(LaunchCtx& c) {
    if (c.exec->fusedRun.size() != 4) { c.fail("froxel chain: a fused run of four executions expected"); return; }
    LaunchCtx m[4] = {c, c, c, c};
    for (int k = 0; k < 4; k++) {
        m[k].exec = c.be_exec(c.exec->fusedRun[(size_t)k]);
        m[k].pass = c.be_pass(m[k].exec->pass);
    }
    FroxelFusedParams p;
    p.materialVolume = m[0].storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.scatteringVolume = m[1].storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.in.noiseTexture = m[0].sampled(1, PLAIN_FORMAT_R8);
    p.settings = m[0].ubuf<plain_volumetric_lighting_settings>(2);
    p.in.sunShadowMap = m[1].sampled(1, PLAIN_FORMAT_DEPTH16);
    p.in.cascades = m[1].sbuf<plain_shadow_cascade_info>(3);
    p.in.light = m[1].sbuf<plain_light_buffer>(4);
    p.historyTarget = m[2].storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.in.historyVolume = m[2].sampled(2, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.integrationVolume = m[3].storage(0, PLAIN_FORMAT_RGBA16_SFLOAT);
    p.in.g = c.g;
    for (int k = 0; k < 4; k++) if (m[k].failed) { c.fail(m[k].error); return; }
    auto sameExtent = [&](const ImgView& v) { return v.w == p.historyTarget.w && v.h == p.historyTarget.h && v.d == p.historyTarget.d; };
    if (!sameExtent(p.materialVolume) || !sameExtent(p.scatteringVolume) || !sameExtent(p.integrationVolume) || p.historyTarget.d > FROXEL_MAX_DEPTH) { c.fail("froxel chain: fused volumes of different extents"); return; }
    int y0, y1;
    c.window(p.historyTarget.h, y0, y1);
    if (y1 <= y0) return;
    p.yBegin = y0;
    // z lanes per block, A / B switch: 64 / 32 / 16 (default). Measured at 3840x2160 (profiles/r4_froxel_fusion.md): 0.430 / 0.357 / 0.338 ms - eight
    // 128-thread blocks per SM cover each other's barrier phases (prologue, running sums), two 512-thread blocks do not
    static const int zLanes = getenv("PLAIN_FROXEL_ZLANES") ? atoi(getenv("PLAIN_FROXEL_ZLANES")) : 16;
    const dim3 grid(ceilDiv(p.historyTarget.w, FROXEL_COLS), (unsigned)(y1 - y0));
    if (zLanes == 64) PLAIN_LAUNCH(c, froxelColumnKernel<64>, grid, FROXEL_COLS * 64, 0, p);
    else if (zLanes == 32) PLAIN_LAUNCH(c, froxelColumnKernel<32>, grid, FROXEL_COLS * 32, 0, p);
    else PLAIN_LAUNCH(c, froxelColumnKernel<16>, grid, FROXEL_COLS * 16, 0, p);
}
PLAIN_PASS(launch_volumetricLightingIntegration, "volumetricLightingIntegration.comp") {
    if (!c.exec->fusedRun.empty()) { launchFroxelColumns(c); return; }
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
