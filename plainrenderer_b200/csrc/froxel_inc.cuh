// froxel_inc.cuh - the per-froxel bodies of the four volumetric-lighting shaders, host+device so that
// tests/test_froxel_fusion_cpu.py can run the fused column path on the CPU and hold it against the oracle bit for bit.
//   froxelVolumeMaterial.comp:17-44, froxelLightScattering.comp:31-64, volumeLightingReprojection.comp:19-62,
//   volumetricLightingIntegration.comp:18-43, volumetricFroxelLighting.inc:1-55
// Two users (passes_volumetrics.cu): the four per-pass kernels (one thread per froxel, every value formed where the shader forms
// it) and the fused column kernel, which forms the SAME values in a different place: everything that depends on (x, y) only - the
// three view directions of a column and V / dot(-V, forward) - once per column, everything that depends on z only - the three
// froxelUVToDepth tables - once per block. Under the contract (no contraction, IEEE operations, left to right) a value does not
// depend on which thread computes it, so the bits are those of the per-pass kernels.
#pragma once
#include "shader_inc.cuh"

namespace pb {

// a texel stored as RGBA16F and read back: what the next pass of the chain sees
PV_HD vec4 roundRGBA16F(vec4 c) { return v4(halfToFloat(floatToHalf(c.x)), halfToFloat(floatToHalf(c.y)), halfToFloat(floatToHalf(c.z)), halfToFloat(floatToHalf(c.w))); }

PV_HD vec3 froxelWorldPos(const Globals& G, vec3 uv, float maxDistance) {
    const vec3 ndc = 2.f * (uv - 0.5f);
    const vec3 V = viewDirFromNDC(G, v2(ndc.x, ndc.y));
    return G.camPos - V / dot(-V, G.fwd) * froxelUVToDepth(uv.z, maxDistance);
}

// trilinear sample of the R8 density noise with repeat addressing (sampleLinear3D<WRAP_REPEAT> of image_view.h: same set-up, same blend). When all
// three extents are powers of two (the 32^3 volume of Volumetrics.cpp:60-75) the six index wraps are masks instead of integer remainders:
// for size = 2^k, i & (size - 1) IS the non-negative remainder of the two's complement i - an integer identity, no arithmetic changes.
PV_HD float froxelNoiseSample(const ImgView& t, vec3 uvw) {
    const int w = t.w, h = t.h, d = t.d;
    if (((w & (w - 1)) | (h & (h - 1)) | (d & (d - 1))) != 0 || w <= 0 || h <= 0 || d <= 0)
        return sampleLinear3D<WRAP_REPEAT, float>([&](int tx, int ty, int tz) { return loadR8(t, tx, ty, tz); }, w, h, d, uvw, 0.f);
    const Bilerp b = bilerpSetup(v2(uvw.x, uvw.y), w, h);
    const float fz = fmaf_(sanitizeCoord(uvw.z), (float)d, -0.5f);
    const float z0f = floorf_(fz);
    const float az = fz - z0f, bz = 1.f - az;
    const int z0i = f2i(z0f);
    const int x0 = b.x0 & (w - 1), x1 = (b.x0 + 1) & (w - 1), y0 = b.y0 & (h - 1), y1 = (b.y0 + 1) & (h - 1), z0 = z0i & (d - 1), z1 = (z0i + 1) & (d - 1);
    const float a00 = loadR8(t, x0, y0, z0), a10 = loadR8(t, x1, y0, z0), a01 = loadR8(t, x0, y1, z0), a11 = loadR8(t, x1, y1, z0);
    const float b00 = loadR8(t, x0, y0, z1), b10 = loadR8(t, x1, y0, z1), b01 = loadR8(t, x0, y1, z1), b11 = loadR8(t, x1, y1, z1);
    const float s0 = vfma(a11, b.w11, vfma(a01, b.w01, vfma(a10, b.w10, a00 * b.w00)));
    const float s1 = vfma(b11, b.w11, vfma(b01, b.w01, vfma(b10, b.w10, b00 * b.w00)));
    return vfma(s1, az, s0 * bz);
}

// ---------------- froxelVolumeMaterial.comp:24-43, from the world position on ----------------
PV_HD vec4 froxelMaterialAt(const plain_volumetric_lighting_settings& s, const ImgView& noiseTexture, vec3 posWorld) {
    const float noiseScale = 0.5f;
    const vec3 noiseSample = posWorld * noiseScale + ld3(s.windSampleOffset);
    const float noise = froxelNoiseSample(noiseTexture, noiseSample);
    vec3 scatteringCoefficient = ld3(s.scatteringCoefficients);
    float absorptionCoefficient = s.absorptionCoefficient;
    float densityMultiplier = s.baseDensity;
    densityMultiplier += s.densityNoiseRange * (noise - 0.5f);
    densityMultiplier = fmaxp(densityMultiplier, 0.f);
    scatteringCoefficient = scatteringCoefficient * densityMultiplier;
    absorptionCoefficient *= densityMultiplier;
    return v4(scatteringCoefficient, absorptionCoefficient);
}

// ---------------- froxelLightScattering.comp:45-63, from the view direction and the world position on ----------------
// sa = the material texel of this froxel (scattering coefficients, absorption coefficient)
PV_HD vec4 froxelScatteringAt(const plain_global_shader_info* g, const plain_volumetric_lighting_settings& s, const plain_shadow_cascade_info* cascades,
                              const plain_light_buffer* light, const ImgView& sunShadowMap, vec3 V, vec3 posWorld, vec4 sa) {
    const float shadow = simpleShadow<false>(posWorld, cascades->lightMatrices[2], sunShadowMap);  // hard-coded cascade 2 (:45)
    const float sunStrength = shadow * light->sunStrengthExposed;
    const vec3 L = v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2]);
    const float VoL = dot(-V, L);
    const float phase = phaseGreenstein(VoL, s.phaseFunctionG);
    const vec3 scatteringCoefficient = xyz(sa);
    const float absorptionCoefficient = sa.w;
    const vec3 constantAmbientLighting = v3(0.02f);
    const vec3 inscattering = (sunStrength * phase * ld3(light->sunColor) + constantAmbientLighting) * scatteringCoefficient;
    const vec3 extinctionCoefficient = scatteringCoefficient + absorptionCoefficient;
    const float transmittance = computeLuminance(extinctionCoefficient);
    return v4(inscattering, transmittance);
}

// ---------------- volumeLightingReprojection.comp:36-61, from the world position on ----------------
// current = the scattering texel of this froxel
PV_HD vec4 froxelReprojectionAt(const plain_global_shader_info* g, float maxDistance, const ImgView& historyVolume, vec3 posWorld, vec4 current) {
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
    return vmix(current, history, alpha);
}

// ---------------- volumetricLightingIntegration.comp:30-40: what one slice adds, before the running sums ----------------
// it = the reprojected texel of this froxel; returns (inscattering of the segment, transmittance factor of the segment)
PV_HD vec4 froxelSegmentTerms(vec4 it, float segmentLength) {
    const vec3 inscattering = integrateInscattering(xyz(it), v3(it.w), segmentLength);
    return v4(inscattering, dm::exp(-it.w * segmentLength));
}

// ================= the fused column path =================
// uv of a froxel centre along one axis: (i + 0.5 + sampleOffset) / res as the material and scattering shaders form it, (i + 0.5) / res as the
// reprojection does (a vector / vector division multiplies by the correctly rounded reciprocal, pvec.h)
PV_HD float froxelAxisUVJittered(int i, float sampleOffset, float res) { return (((float)i + 0.5f) + sampleOffset) * rcpf_(res); }
PV_HD float froxelAxisUV(int i, float res) { return ((float)i + 0.5f) * rcpf_(res); }

// what a column (x, y) contributes to the world positions of its froxels: V and W = V / dot(-V, forward), so that posWorld = camPos - W * depth(z).
// variant 0 = froxelVolumeMaterial.comp (jittered uv, ndc = 2 * (uv - 0.5)), 1 = froxelLightScattering.comp (jittered uv, ndc = 2 * uv - 1, :40),
// 2 = volumeLightingReprojection.comp (uv of the froxel centre, ndc = 2 * (uv - 0.5))
PV_HD void froxelColumnSetup(const Globals& G, int variant, int x, int y, float sampleOffset, float resX, float resY, vec3& V, vec3& W) {
    const float ux = variant == 2 ? froxelAxisUV(x, resX) : froxelAxisUVJittered(x, sampleOffset, resX);
    const float uy = variant == 2 ? froxelAxisUV(y, resY) : froxelAxisUVJittered(y, sampleOffset, resY);
    const vec2 ndc = variant == 1 ? v2(2.f * ux - 1.f, 2.f * uy - 1.f) : v2(2.f * (ux - 0.5f), 2.f * (uy - 0.5f));
    V = viewDirFromNDC(G, ndc);
    W = V / dot(-V, G.fwd);
}
// entry j of a block's depth tables for a volume of resZ slices: [0, resZ) = froxelUVToDepth of the jittered slice centre (material, scattering),
// [resZ, 2 resZ) = of the slice centre (reprojection), [2 resZ, 3 resZ] = of the slice boundaries z / resZ (integration: start and end of a segment)
PV_HD float froxelDepthTableEntry(int j, int resZ, float sampleOffset, float maxDistance) {
    float uvZ;
    if (j < resZ) uvZ = froxelAxisUVJittered(j, sampleOffset, (float)resZ);
    else if (j < 2 * resZ) uvZ = froxelAxisUV(j - resZ, (float)resZ);
    else uvZ = (float)(j - 2 * resZ) / (float)resZ;
    return froxelUVToDepth(uvZ, maxDistance);
}

struct FroxelFusedInputs {
    ImgView noiseTexture, sunShadowMap, historyVolume;
    const plain_shadow_cascade_info* cascades;
    const plain_light_buffer* light;
    const plain_global_shader_info* g;
};
// material -> scattering -> reprojection of ONE froxel with the column's and the slice's shared values; every intermediate texel is rounded
// through binary16 as its store + load would. material / scattering / reprojected are the three texels as the per-pass kernels store them.
PV_HD void froxelFusedTexel(const FroxelFusedInputs& in, const Globals& G, const plain_volumetric_lighting_settings& s, vec3 Wm, vec3 Vs, vec3 Ws, vec3 Wr,
                            float depthJittered, float depthCentre, vec4& material, vec4& scattering, vec4& reprojected) {
    material = roundRGBA16F(froxelMaterialAt(s, in.noiseTexture, G.camPos - Wm * depthJittered));
    scattering = roundRGBA16F(froxelScatteringAt(in.g, s, in.cascades, in.light, in.sunShadowMap, Vs, G.camPos - Ws * depthJittered, material));
    reprojected = roundRGBA16F(froxelReprojectionAt(in.g, s.maxDistance, in.historyVolume, G.camPos - Wr * depthCentre, scattering));
}


// ---- one block of the fused launch: 8 columns x ZLANES z lanes of one froxel row, a thread owns the froxels (x, y, z = lane, lane + ZLANES, ..) of
// its column. The phases are functions of the thread index so that the kernel (passes_volumetrics.cu froxelColumnKernel: one call per thread and
// phase, __syncthreads between the phases) and the CPU check (tests/emul/froxel_fusion_host.cu: a loop over the thread indices per phase) run the
// same statements. ----
#define FROXEL_COLS 8
// ZLANES (template parameter of the phases): z lanes of a block, 64 / 32 / 16 = 512 / 256 / 128 threads; the smaller the block, the more blocks
// share an SM and cover each other's barrier phases (the froxels a thread owns grow accordingly: z = lane, lane + ZLANES, ..)
#define FROXEL_MAX_DEPTH 128  // planFusions (backend.cu) does not fuse deeper volumes
struct FroxelBlockShared {
    float depth[3 * FROXEL_MAX_DEPTH + 1];          // froxelDepthTableEntry
    float column[3][FROXEL_COLS][6];                // per variant and column: V.xyz, W.xyz (froxelColumnSetup)
    float4 terms[FROXEL_MAX_DEPTH][FROXEL_COLS];    // per slice and column: the segment's terms (phase 1), then the running sums (phase 2)
};
struct FroxelFusedParams {
    ImgView historyTarget, integrationVolume;       // the two volumes the launch writes
    ImgView materialVolume, scatteringVolume;       // written only by froxelBlockPhase1<true> (the product leaves them alone; the CPU check compares them)
    FroxelFusedInputs in;
    const plain_volumetric_lighting_settings* settings;
    int yBegin;
};
// prologue: 24 threads form the three (V, W) pairs of the block's 8 columns, the other warps the three depth tables (3 d + 1 exponentials per block
// instead of four to five per froxel)
template <int ZLANES> PV_HD void froxelBlockPrologue(FroxelBlockShared& sh, const FroxelFusedParams& p, const Globals& G, const plain_volumetric_lighting_settings& s, int tid, int blockX, int y) {
    const int resX = p.historyTarget.w, resY = p.historyTarget.h, resZ = p.historyTarget.d;
    if (tid < 3 * FROXEL_COLS) {
        const int variant = tid / FROXEL_COLS, xl = tid % FROXEL_COLS;
        vec3 V, W;
        froxelColumnSetup(G, variant, blockX * FROXEL_COLS + xl, y, s.sampleOffset, (float)resX, (float)resY, V, W);
        float* c = sh.column[variant][xl];
        c[0] = V.x; c[1] = V.y; c[2] = V.z; c[3] = W.x; c[4] = W.y; c[5] = W.z;
    } else if (tid >= 32) {
        for (int j = tid - 32; j <= 3 * resZ; j += FROXEL_COLS * ZLANES - 32) sh.depth[j] = froxelDepthTableEntry(j, resZ, s.sampleOffset, s.maxDistance);
    }
}
// phase 1: material -> scattering -> reprojection of the thread's froxels in registers; the reprojected texel (= next frame's history) is stored, the
// slice's integration terms are left in shared memory
template <int ZLANES, bool WRITE_INTERMEDIATES> PV_HD void froxelBlockPhase1(FroxelBlockShared& sh, const FroxelFusedParams& p, const Globals& G, const plain_volumetric_lighting_settings& s, int tid, int blockX, int y) {
    const int xl = tid % FROXEL_COLS, zl = tid / FROXEL_COLS, x = blockX * FROXEL_COLS + xl;
    const int resZ = p.historyTarget.d;
    if (x >= p.historyTarget.w) return;
    const float* cm = sh.column[0][xl];
    const float* cs = sh.column[1][xl];
    const float* cr = sh.column[2][xl];
    const vec3 Wm = v3(cm[3], cm[4], cm[5]), Vs = v3(cs[0], cs[1], cs[2]), Ws = v3(cs[3], cs[4], cs[5]), Wr = v3(cr[3], cr[4], cr[5]);
    for (int z = zl; z < resZ; z += ZLANES) {
        vec4 material, scattering, reprojected;
        froxelFusedTexel(p.in, G, s, Wm, Vs, Ws, Wr, sh.depth[z], sh.depth[resZ + z], material, scattering, reprojected);
        storeRGBA16F(p.historyTarget, x, y, z, reprojected);
        if (WRITE_INTERMEDIATES) {
            storeRGBA16F(p.materialVolume, x, y, z, material);
            storeRGBA16F(p.scatteringVolume, x, y, z, scattering);
        }
        const float segmentLength = sh.depth[2 * resZ + z + 1] - sh.depth[2 * resZ + z];
        const vec4 terms = froxelSegmentTerms(reprojected, segmentLength);
        float4 t;
        t.x = terms.x; t.y = terms.y; t.z = terms.z; t.w = terms.w;
        sh.terms[z][xl] = t;
    }
}
// phase 2: one thread per column forms the running sums front to back - the only serial part of volumetricLightingIntegration.comp:28-41, three
// additions and one multiplication per slice
PV_HD void froxelBlockPhase2(FroxelBlockShared& sh, const FroxelFusedParams& p, int tid, int blockX) {
    if (tid >= FROXEL_COLS || blockX * FROXEL_COLS + tid >= p.historyTarget.w) return;
    const int resZ = p.historyTarget.d;
    float tx = 0.f, ty = 0.f, tz = 0.f, transmittance = 1.f;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
    for (int z = 0; z < resZ; z++) {
        float4 t = sh.terms[z][tid];
        tx = tx + t.x; ty = ty + t.y; tz = tz + t.z;
        transmittance *= t.w;
        t.x = tx; t.y = ty; t.z = tz; t.w = transmittance;
        sh.terms[z][tid] = t;
    }
}
// phase 3: every thread stores its froxels of the integrated volume
template <int ZLANES> PV_HD void froxelBlockPhase3(const FroxelBlockShared& sh, const FroxelFusedParams& p, int tid, int blockX, int y) {
    const int xl = tid % FROXEL_COLS, zl = tid / FROXEL_COLS, x = blockX * FROXEL_COLS + xl;
    if (x >= p.historyTarget.w) return;
    for (int z = zl; z < p.historyTarget.d; z += ZLANES) {
        const float4 t = sh.terms[z][xl];
        storeRGBA16F(p.integrationVolume, x, y, z, v4(t.x, t.y, t.z, t.w));
    }
}

}  // namespace pb
