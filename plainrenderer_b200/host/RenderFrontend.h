// RenderFrontend.h - host-side mirror of the reference's frame-path callers above the C-ABI:
// RenderFrontend (Plain/src/Runtime/Rendering/RenderFrontend.{h,cpp}) and the technique drivers SDFGI, TAA, Sky,
// Volumetrics, Bloom (Plain/src/Runtime/Rendering/Techniques/*). Same pass order, shader names, binding numbers,
// specialisation constants, ping-pong indices and host-computed buffers; the rasterisation passes are replaced by
// uploads of their outputs, and the forward/sky/sun-sprite graphic passes by one "gbufferShading.comp" execution.
#pragma once
#include <array>
#include <string>
#include <vector>
#include "HostMath.h"
#include "RenderBackend.h"
#include "plain_frame_types.h"

// ---- settings, defaults as in the reference ----
enum class DiffuseBRDF : int { Lambert = 0, Disney = 1, CoDWWII = 2, Titanfall2 = 3 };
enum class DirectSpecularMultiscattering : int { McAuley = 0, Simplified = 1, ScaledGGX = 2, None = 3 };
enum class IndirectLightingTech : int { SDFTrace = 0, ConstantAmbient = 1 };
struct ShadingConfig {  // RenderFrontend.h:32-38
    DiffuseBRDF diffuseBRDF = DiffuseBRDF::CoDWWII;
    DirectSpecularMultiscattering directMultiscatter = DirectSpecularMultiscattering::McAuley;
    IndirectLightingTech indirectLightingTech = IndirectLightingTech::SDFTrace;
    bool useGeometryAA = true;
    int sunShadowCascadeCount = 3;
};
struct SDFTraceSettings {  // SDFGI.h:17-29
    bool halfResTrace = true;
    bool strictInfluenceRadiusCutoff = true;
    float traceInfluenceRadius = 5.f;
    float additionalSunShadowMapPadding = 3.f;
};
enum class SDFVisualisationMode : int { None = 0, VisualizeSDF = 1, CameraTileUsage = 2, SDFNormals = 3, RaymarchingCount = 4 };  // SDFGI.h:9
struct SDFDebugSettings {  // SDFGI.h:11-15
    SDFVisualisationMode visualisationMode = SDFVisualisationMode::None;
    bool showCameraTileUsageWithHiZ = true;
    bool useInfluenceRadiusForDebug = false;
};
enum class HistorySamplingTech : int { Bilinear = 0, Bicubic16Tap = 1, Bicubic9Tap = 2, Bicubic5Tap = 3, Bicubic1Tap = 4 };
struct TAASettings {  // TAA.h:8-17
    bool enabled = true;
    bool useSeparateSupersampling = false;
    bool useClipping = true;
    bool useMotionVectorDilation = true;
    HistorySamplingTech historySamplingTech = HistorySamplingTech::Bicubic1Tap;
    bool supersampleUseTonemapping = true;
    bool filterUseTonemapping = true;
    bool useMipBias = true;
};
struct BloomSettings { bool enabled = true; float strength = 0.05f; float radius = 1.5f; };  // Bloom.h:5-9
struct VolumetricsSettings {  // Volumetrics.h:5-13
    float scatteringCoefficients[3] = {1.f, 1.f, 1.f};
    float maxDistance = 30.f;
    float absorptionCoefficient = 1.f;
    float baseDensity = 0.003f;
    float densityNoiseRange = 0.008f;
    float densityNoiseScale = 0.5f;
    float phaseFunctionG = 0.2f;
};
struct WindSettings { hm::Vec3 vector = hm::Vec3(0.f); float speed = 0.15f; };  // Volumetrics.h:15-18
struct CameraExtrinsic {  // Camera.h:4-9
    hm::Vec3 position = hm::Vec3(0.f, -1.f, -5.f);
    hm::Vec3 forward = hm::Vec3(0.f, 0.f, -1.f);
    hm::Vec3 right = hm::Vec3(1.f, 0.f, 0.f);
    hm::Vec3 up = hm::Vec3(0.f, -1.f, 0.f);
};
struct CameraIntrinsic { float fov = 35.f; float aspectRatio = 1.f; float near = 0.1f; float far = 300.f; };  // Camera.h:11-16

struct FrameRenderTargets { ImageHandle colorBuffer, motionBuffer, depthBuffer; };  // FrameRenderTargets.h
struct ViewFrustum {  // ViewFrustum.h:6-30
    hm::Vec3 l_l_n, l_l_f, l_u_n, l_u_f, r_l_n, r_l_f, r_u_n, r_u_f;
    hm::Vec3 top, bot, right, left, near, far;
};
ViewFrustum computeViewFrustum(const CameraExtrinsic& e, const CameraIntrinsic& i);
ViewFrustum computeOrthogonalFrustumFittedToCamera(const ViewFrustum& cameraFrustum, hm::Vec3 lightDirection);  // ViewFrustum.cpp:231-271
hm::Mat4 viewMatrixFromCameraExtrinsic(const CameraExtrinsic& e);      // Camera.cpp:4-12
hm::Mat4 projectionMatrixFromCameraIntrinsic(const CameraIntrinsic& i);  // Camera.cpp:14-27

// frame counters of the reference's main loop (FrameIndex.cpp:12-18), owned by the frontend instead of a global
struct FrameIndex {
    size_t frameIndex = 0;
    void markNewFrame() { frameIndex++; }
    size_t mod2() const { return frameIndex % 2; }
    size_t mod8() const { return frameIndex % 8; }
};
hm::Vec2 hammersley2D(uint32_t index);  // MathUtils.cpp:25-70
hm::Vec3 directionToVector(hm::Vec2 directionDegrees);  // MathUtils.cpp:4-15
uint32_t mipCountFromResolution(uint32_t w, uint32_t h, uint32_t d);
hm::AABB padSDFBoundingBox(const hm::AABB& bb);  // sdfUtilities.cpp:5-19

// scene objects with an SDF (RuntimeScene.h RenderObject + MeshFrontend.h)
struct Material { uint32_t albedoTextureIndex = 0, normalTextureIndex = 0, specularTextureIndex = 0; };  // MeshFrontend.h: bindless slots
struct MeshFrontend { int sdfTextureIndex = -1; hm::Vec3 meanAlbedo = hm::Vec3(0.5f); hm::AABB localBB; MeshHandle backendHandle; Material material; };
struct RenderObject { uint32_t mesh = 0; hm::AABB bbWorld; hm::Mat4 modelMatrix; hm::Mat4 previousModelMatrix; };
struct DefaultTextures { ImageHandle diffuse, specular, normal; };  // RenderFrontend.cpp:86-150
bool isAxisAlignedBoundingBoxIntersectingViewFrustum(const ViewFrustum& frustum, const hm::AABB& bb);  // Culling.cpp:5-42

struct SDFTraceDependencies {
    FrameRenderTargets currentFrame, previousFrame;
    ViewFrustum cameraFrustum;
    ImageHandle depthHalfRes, worldSpaceNormals, skyLut, shadowMap, depthMinMaxPyramid;
    StorageBufferHandle lightBuffer, sunShadowInfoBuffer;
};

class SDFGI {
public:
    void init(RenderBackend& b, int w, int h, const SDFTraceSettings& s, const SDFDebugSettings& debug, int sunShadowCascadeIndex);
    void renderSDFVisualization(RenderBackend& b, ImageHandle target, const SDFTraceDependencies& d, const SDFDebugSettings& debug, const SDFTraceSettings& s) const;
    void updateSDFDebugSettings(RenderBackend& b, const SDFDebugSettings& debug, int sunShadowCascadeIndex);
    void updateSDFScene(RenderBackend& b, const std::vector<RenderObject>& scene, const std::vector<MeshFrontend>& meshes);
    struct IndirectLightingImages { ImageHandle Y_SH, CoCg; };
    IndirectLightingImages getIndirectLightingResults(bool tracedHalfRes) const;
    void computeIndirectLighting(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s, const FrameIndex& fi) const;
    ImageHandle m_indirectDiffuse_Y_SH[2], m_indirectDiffuse_CoCg[2], m_indirectDiffuseHistory_Y_SH[2], m_indirectDiffuseHistory_CoCg[2];
    ImageHandle m_indirectLightingFullRes_Y_SH, m_indirectLightingFullRes_CoCg;
    StorageBufferHandle m_sdfInstanceBuffer, m_sdfCameraFrustumCulledInstances, m_sdfInstanceWorldBBBuffer, m_sdfCameraCulledTiles;
    UniformBufferHandle m_cameraFrustumBuffer, m_sdfTraceInfluenceRangeBuffer;
private:
    void sdfInstanceCulling(RenderBackend& b, const SDFTraceDependencies& d, int targetW, int targetH, float influenceRadius, bool hiZ) const;
    void diffuseSDFTrace(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s) const;
    void filterIndirectDiffuse(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s, const FrameIndex& fi) const;
    uint32_t m_sdfInstanceCount = 0;
    RenderPassHandle m_diffuseSDFTracePass, m_indirectDiffuseFilterSpatialPass[2], m_indirectDiffuseFilterTemporalPass, m_indirectLightingUpscale;
    RenderPassHandle m_sdfCameraFrustumCulling, m_sdfCameraTileCulling, m_sdfCameraTileCullingHiZ, m_sdfDebugVisualisationPass;
};

class TAA {
public:
    void init(RenderBackend& b, int w, int h, const TAASettings& s);
    void updateSettings(RenderBackend& b, const TAASettings& s);
    void computeTemporalSuperSampling(RenderBackend& b, const FrameRenderTargets& currentFrame, const FrameRenderTargets& lastFrame, ImageHandle target, const FrameIndex& fi) const;
    void computeTemporalFilter(RenderBackend& b, ImageHandle colorSrc, const FrameRenderTargets& currentFrame, ImageHandle target, const FrameIndex& fi) const;
    hm::Vec2 computeProjectionMatrixJitter(const FrameIndex& fi) const;
    hm::Mat4 applyProjectionMatrixJitter(const hm::Mat4& projection, hm::Vec2 offset) const;
    void updateTaaResolveWeights(RenderBackend& b, hm::Vec2 cameraJitterInPixels);
    ImageHandle m_historyBuffers[2], m_sceneLuminance[2];
    std::array<float, 9> m_lastResolveWeights{};
private:
    RenderPassHandle m_temporalFilterPass, m_temporalSupersamplingPass, m_colorToLuminancePass;
    UniformBufferHandle m_taaResolveWeightBuffer;
};

class Sky {
public:
    void init(RenderBackend& b);
    void updateTransmissionLut(RenderBackend& b) const;
    void updateSkyLut(RenderBackend& b, StorageBufferHandle lightBuffer, const plain_atmosphere_settings& a) const;
    // model matrix of the sun sprite quad (Sky.cpp:247-264), passed to the shading pass as push constants
    hm::Mat4 sunSpriteModelMatrix(hm::Vec2 sunDirectionDegrees) const;
    ImageHandle m_skyTransmissionLut, m_skyMultiscatterLut, m_skyLut;
private:
    RenderPassHandle m_skyTransmissionLutPass, m_skyMultiscatterLutPass, m_skyLutPass;
    UniformBufferHandle m_atmosphereSettingsBuffer;
};

class Volumetrics {
public:
    void init(RenderBackend& b, int w, int h, uint32_t noiseSeed);
    struct Dependencies { ImageHandle shadowMap; StorageBufferHandle sunShadowInfoBuffer, lightBuffer; };
    void computeVolumetricLighting(RenderBackend& b, const VolumetricsSettings& s, const WindSettings& wind, const Dependencies& d, const FrameIndex& fi, float deltaTime);
    ImageHandle m_volumeMaterialVolume, m_scatteringTransmittanceVolume, m_volumetricLightingHistory[2], m_volumetricIntegrationVolume, m_perlinNoise3D;
    UniformBufferHandle m_volumetricsSettingsUniforms;
private:
    RenderPassHandle m_froxelVolumeMaterialPass, m_froxelScatteringTransmittancePass, m_volumetricLightingIntegration, m_volumetricLightingReprojection;
    hm::Vec3 m_windSampleOffset = hm::Vec3(0.f);
};

class Bloom {
public:
    void init(RenderBackend& b);
    void computeBloom(RenderBackend& b, ImageHandle targetImage, const BloomSettings& s) const;
    mutable ImageHandle m_lastDownscaleTexture, m_lastUpscaleTexture;  // transient handles of the last frame (tests)
private:
    std::vector<RenderPassHandle> m_bloomDownsamplePasses, m_bloomUpsamplePasses;
    RenderPassHandle m_applyBloomPass;
};

class RenderFrontend {
public:
    void setup(int device, uint32_t width, uint32_t height, uint32_t noiseSeed);
    void shutdown();
    // per frame, in the order of the reference's main loop (main.cpp:79-90)
    void markNewFrame(float time, float deltaTime);  // Timer::markNewFrame + FrameIndex::markNewFrame
    void prepareNewFrame();                          // newFrame + prepareRenderpasses (the ordered pass list)
    void setCameraExtrinsic(const CameraExtrinsic& e);
    void prepareForDrawcalls();                      // updateGlobalShaderInfo
    void renderScene(const std::vector<RenderObject>& scene);
    void renderFrame();
    bool renderFrameSegment(plain_exchange* pending);  // row-sharded frames: run up to the next exchange
    uint32_t m_shardRank = 0, m_shardCount = 1;

    // with m_rasterInputs the depth prepass, the sun shadow cascades and the G-buffer are rasterised from the
    // registered meshes by the backend's graphic passes instead of being uploaded
    bool m_rasterInputs = false;
    bool m_motionRowsOnly = false;  // row-sharded, uploaded inputs: only this rank's rows of the motion vectors were uploaded (all-gathered with the pyramid level)
    void setMeshGeometry(uint32_t mesh, const MeshBinary& geometry, const Material* material);  // registerMeshes :456-531 (geometry + material part)
    DefaultTextures m_defaultTextures;
    uint32_t m_currentMainPassDrawcallCount = 0, m_currentShadowPassDrawcallCount = 0;
    uint32_t registerSdfMesh(const uint16_t* r16fTexels, uint32_t rx, uint32_t ry, uint32_t rz, const hm::AABB& localBB, hm::Vec3 meanAlbedo);
    const FrameRenderTargets& currentTargets() const { return m_frameRenderTargets[m_sceneRenderTargetIndex]; }

    RenderBackend backend;
    ShadingConfig m_shadingConfig;
    SDFTraceSettings m_sdfTraceSettings;
    SDFDebugSettings m_sdfDebugSettings;
    TAASettings m_taaSettings;
    BloomSettings m_bloomSettings;
    VolumetricsSettings m_volumetricsSettings;
    WindSettings m_windSettings;
    plain_atmosphere_settings m_atmosphereSettings;
    hm::Vec2 m_sunDirection;  // degrees (phi, theta), RenderFrontend.h:141
    CameraIntrinsic m_cameraIntrinsic;
    plain_global_shader_info m_globalShaderInfo;
    FrameIndex m_frameIndex;

    uint32_t m_screenWidth = 0, m_screenHeight = 0;
    FrameRenderTargets m_frameRenderTargets[2];
    int m_sceneRenderTargetIndex = 0;
    ImageHandle m_postProcessBuffers[2], m_brdfLut, m_minMaxDepthPyramid, m_depthHalfRes;
    // raster-pass outputs the caller uploads every frame; two of each, indexed like m_frameRenderTargets, so that the upload
    // of frame N+1 (copy engine) overlaps the passes of frame N instead of waiting for them to release a single image
    ImageHandle m_worldSpaceNormalImages[2], m_gbuffers[2], m_motionBuffers[3];
    int m_motionBufferIndex = 0;
    ImageHandle worldSpaceNormalImage() const { return m_worldSpaceNormalImages[m_sceneRenderTargetIndex]; }
    ImageHandle gbuffer() const { return m_gbuffers[m_sceneRenderTargetIndex]; }
    std::vector<ImageHandle> m_shadowMaps, m_noiseTextures;
    StorageBufferHandle m_histogramBuffer, m_lightBuffer, m_histogramPerTileBuffer, m_depthPyramidSyncBuffer, m_sunShadowInfoBuffer;
    UniformBufferHandle m_globalUniformBuffer;
    SDFGI m_sdfGi;
    TAA m_taa;
    Sky m_sky;
    Volumetrics m_volumetrics;
    Bloom m_bloom;
    std::vector<MeshFrontend> m_frontendMeshes;
    ImageHandle m_lastTonemapSource;

private:
    void initImages(uint32_t noiseSeed);
    void initBuffers();
    void initRenderpasses();
    void prepareRenderpasses();
    void computeColorBufferHistogram(ImageHandle lastFrameColor, bool exchange = true);
    void computeExposure();
    void computeDepthPyramid(ImageHandle depthBuffer, const FrameRenderTargets* alsoDownscale = nullptr, bool holdRest = false);
    ComputePassExecution m_pendingPyramidRest;
    ExchangeRequest m_pendingHistogramExchange;
    void computeSunLightMatrices();
    void downscaleDepth(const FrameRenderTargets& current, bool exchange = true);
    void shadeGBuffer(ImageHandle colorTarget);
    void renderDepthPrepass(ImageHandle depth, ImageHandle normal, ImageHandle motion);  // :792-802
    void renderSunShadowCascades();                                                     // :760-775
    void fillGBuffer(ImageHandle gbuffer, ImageHandle depth);                           // the raster half of renderForwardShading :840-...
    void computeTonemapping(ImageHandle src);
    void computeBRDFLut();
    void updateGlobalShaderInfo();
    ShaderDescription createDepthPyramidShaderDescription(uint32_t* outThreadgroupCount) const;
    void computeSinglePassMipChainDispatchCount(uint32_t w, uint32_t h, uint32_t mipCount, uint32_t maxMipCount, uint32_t out[2]) const;

    CameraExtrinsic m_cameraExtrinsic;
    bool m_frameOpen = false;
    hm::Mat4 m_viewProjectionMatrix = hm::Mat4::zero();
    ViewFrustum m_cameraFrustum, m_sunShadowFrustum;
    float m_time = 0.f, m_deltaTime = 0.016f;
    RenderPassHandle m_shadingPass, m_brdfLutPass, m_histogramPerTilePass, m_histogramResetPass, m_histogramCombinePass, m_preExposeLightsPass;
    RenderPassHandle m_depthPyramidPass, m_lightMatrixPass, m_tonemappingPass, m_depthDownscalePass;
    RenderPassHandle m_depthPrePass, m_gbufferFillPass, m_shadowPasses[4];
    StorageBufferHandle m_mainPassTransformsBuffer, m_shadowPassTransformsBuffer;
    SamplerHandle m_samplers[8];
};
