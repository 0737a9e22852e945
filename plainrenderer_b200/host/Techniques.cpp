// Techniques.cpp - host-side mirrors of the reference technique drivers (Plain/src/Runtime/Rendering/Techniques/):
// SDFGI.cpp:44-630, TAA.cpp:85-202, Sky.cpp:10-352, Volumetrics.cpp:17-247, Bloom.cpp:8-144. They create the same
// passes (shader names + specialisation constants) and images, and emit the same executions per frame.
#include <cstring>
#include "RenderFrontend.h"

static const uint32_t SS = PLAIN_USAGE_SAMPLED | PLAIN_USAGE_STORAGE;
static uint32_t ceilDivU(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static void setRows(RenderBackend& b, ComputePassExecution& e, uint32_t divisor, uint32_t rows, uint32_t extend = 0) { b.band(divisor, rows, extend, &e.rowBegin, &e.rowEnd); }
static ExchangeRequest exchangeRows(uint32_t kind, const char* name, std::initializer_list<ImageHandle> images, uint32_t mip, uint32_t divisor, uint32_t halo = 0) {
    ExchangeRequest x;
    x.kind = kind;
    x.name = name;
    x.haloRows = halo;
    for (ImageHandle h : images) { x.images.push_back(h); x.mips.push_back(mip); x.divisors.push_back(divisor); }
    return x;
}
template <typename T> static SpecialisationConstant specConst(uint32_t location, const T& v) { return SpecialisationConstant{location, dataToCharArray(&v, sizeof(T))}; }
static RenderPassHandle makePass(RenderBackend& b, const char* name, const char* shader, std::vector<SpecialisationConstant> consts = {}) {
    ComputePassDescription d;
    d.name = name;
    d.shaderDescription.srcPathRelative = shader;
    d.shaderDescription.specialisationConstants = std::move(consts);
    return b.createComputePass(d);
}

// =============================== SDFGI ===============================
static const uint32_t sdfCameraCullingTileSize = 32;
static const uint32_t maxObjectCountMainScene = 1200;  // SceneConfig.h

static std::vector<SpecialisationConstant> sdfDebugSpecConstants(const SDFDebugSettings& debug, int sunShadowCascadeIndex) {  // SDFGI.cpp:13-29
    return {specConst(0, (int)debug.visualisationMode), specConst(1, sunShadowCascadeIndex)};
}

void SDFGI::init(RenderBackend& b, int w, int h, const SDFTraceSettings& s, const SDFDebugSettings& debug, int sunShadowCascadeIndex) {
    const uint32_t tw = s.halfResTrace ? w / 2 : w, th = s.halfResTrace ? h / 2 : h;
    for (int i = 0; i < 2; i++) {
        m_indirectDiffuseHistory_Y_SH[i] = b.createImage(imageDesc2D(tw, th, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
        m_indirectDiffuse_Y_SH[i] = b.createImage(imageDesc2D(tw, th, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
        m_indirectDiffuse_CoCg[i] = b.createImage(imageDesc2D(tw, th, PLAIN_FORMAT_RG16_SFLOAT, SS), nullptr, 0);
        m_indirectDiffuseHistory_CoCg[i] = b.createImage(imageDesc2D(tw, th, PLAIN_FORMAT_RG16_SFLOAT, SS), nullptr, 0);
    }
    m_indirectLightingFullRes_Y_SH = b.createImage(imageDesc2D(w, h, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    m_indirectLightingFullRes_CoCg = b.createImage(imageDesc2D(w, h, PLAIN_FORMAT_RG16_SFLOAT, SS), nullptr, 0);
    m_sdfInstanceBuffer = b.createStorageBuffer(maxObjectCountMainScene * sizeof(plain_sdf_instance) + sizeof(uint32_t) * 4);
    m_sdfCameraFrustumCulledInstances = b.createStorageBuffer(maxObjectCountMainScene * sizeof(uint32_t) + sizeof(uint32_t));
    m_cameraFrustumBuffer = b.createUniformBuffer(sizeof(plain_camera_frustum_buffer));
    m_sdfInstanceWorldBBBuffer = b.createStorageBuffer(maxObjectCountMainScene * sizeof(plain_bounding_box));
    // the reference sizes this for 1920x1080 (SDFGI.cpp:146-151); sized by the actual resolution here. The tile index
    // strides by the full-resolution tile count (sdfCulling.inc:17-20), so full-res tile counts bound it.
    const size_t tileCount = (size_t)ceilDivU(w, sdfCameraCullingTileSize) * ceilDivU(h, sdfCameraCullingTileSize);
    m_sdfCameraCulledTiles = b.createStorageBuffer(tileCount * sizeof(plain_culled_instances_per_tile));
    m_sdfTraceInfluenceRangeBuffer = b.createUniformBuffer(sizeof(float));
    const uint32_t strict = s.strictInfluenceRadiusCutoff ? 1u : 0u;
    m_diffuseSDFTracePass = makePass(b, "Indirect diffuse SDF trace", "sdfDiffuseTrace.comp", {specConst(0, strict), specConst(1, sunShadowCascadeIndex)});
    for (int i = 0; i < 2; i++) m_indirectDiffuseFilterSpatialPass[i] = makePass(b, "Indirect diffuse spatial filter", "filterIndirectDiffuseSpatial.comp", {specConst(0, i)});
    m_indirectDiffuseFilterTemporalPass = makePass(b, "Indirect diffuse temporal filter", "filterIndirectDiffuseTemporal.comp");
    m_indirectLightingUpscale = makePass(b, "Indirect lighting upscale", "indirectLightUpscale.comp");
    m_sdfCameraFrustumCulling = makePass(b, "SDF camera frustum culling", "sdfCameraFrustumCulling.comp");
    m_sdfCameraTileCulling = makePass(b, "SDF camera tile culling", "sdfCameraTileCulling.comp", {specConst(0, 0u)});
    m_sdfCameraTileCullingHiZ = makePass(b, "SDF camera tile culling", "sdfCameraTileCulling.comp", {specConst(0, 1u)});
    m_sdfDebugVisualisationPass = makePass(b, "SDF debug visualisation", "sdfDebugVisualisation.comp", sdfDebugSpecConstants(debug, sunShadowCascadeIndex));
}

void SDFGI::updateSDFDebugSettings(RenderBackend& b, const SDFDebugSettings& debug, int sunShadowCascadeIndex) {  // SDFGI.cpp:371-373
    ShaderDescription d;
    d.srcPathRelative = "sdfDebugVisualisation.comp";
    d.specialisationConstants = sdfDebugSpecConstants(debug, sunShadowCascadeIndex);
    b.updateComputePassShaderDescription(m_sdfDebugVisualisationPass, d);
}

void SDFGI::renderSDFVisualization(RenderBackend& b, ImageHandle target, const SDFTraceDependencies& d, const SDFDebugSettings& debug, const SDFTraceSettings& s) const {  // SDFGI.cpp:334-369
    const float sdfInfluenceRadius = debug.useInfluenceRadiusForDebug ? s.traceInfluenceRadius : 0.f;
    const ImageDescription targetDescription = b.getImageDescription(m_indirectLightingFullRes_CoCg);
    const bool useHiZCulling = debug.visualisationMode == SDFVisualisationMode::CameraTileUsage && debug.showCameraTileUsageWithHiZ;
    sdfInstanceCulling(b, d, (int)targetDescription.width, (int)targetDescription.height, sdfInfluenceRadius, useHiZCulling);
    ComputePassExecution e;
    e.genericInfo.handle = m_sdfDebugVisualisationPass;
    e.genericInfo.resources.storageImages = {ImageResource(target, 0, 0)};
    e.genericInfo.resources.sampledImages = {ImageResource(d.skyLut, 0, 2), ImageResource(d.shadowMap, 0, 7)};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(d.lightBuffer, true, 1), StorageBufferResource(m_sdfInstanceBuffer, true, 3), StorageBufferResource(m_sdfCameraCulledTiles, true, 4),
                                              StorageBufferResource(m_sdfCameraFrustumCulledInstances, true, 5), StorageBufferResource(d.sunShadowInfoBuffer, true, 6)};
    e.dispatchCount[0] = ceilDivU(targetDescription.width, 8);
    e.dispatchCount[1] = ceilDivU(targetDescription.height, 8);
    b.setComputePassExecution(e);
}

void SDFGI::updateSDFScene(RenderBackend& b, const std::vector<RenderObject>& scene, const std::vector<MeshFrontend>& meshes) {  // SDFGI.cpp:260-313
    std::vector<plain_bounding_box> instanceWorldBBs;
    std::vector<plain_sdf_instance> instanceData;
    for (const RenderObject& obj : scene) {
        const MeshFrontend& mesh = meshes[obj.mesh];
        if (mesh.sdfTextureIndex < 0) continue;
        const hm::AABB paddedWorldBB = padSDFBoundingBox(obj.bbWorld);
        plain_bounding_box wbb{};
        wbb.bbMin[0] = paddedWorldBB.min.x; wbb.bbMin[1] = paddedWorldBB.min.y; wbb.bbMin[2] = paddedWorldBB.min.z;
        wbb.bbMax[0] = paddedWorldBB.max.x; wbb.bbMax[1] = paddedWorldBB.max.y; wbb.bbMax[2] = paddedWorldBB.max.z;
        instanceWorldBBs.push_back(wbb);
        plain_sdf_instance inst{};
        inst.sdfTextureIndex = (uint32_t)mesh.sdfTextureIndex;
        const hm::AABB paddedLocalBB = padSDFBoundingBox(mesh.localBB);
        const hm::Vec3 ext = paddedLocalBB.max - paddedLocalBB.min;
        inst.localExtends[0] = ext.x; inst.localExtends[1] = ext.y; inst.localExtends[2] = ext.z;
        inst.meanAlbedo[0] = mesh.meanAlbedo.x; inst.meanAlbedo[1] = mesh.meanAlbedo.y; inst.meanAlbedo[2] = mesh.meanAlbedo.z;
        const hm::Vec3 bbOffset = (paddedLocalBB.min + paddedLocalBB.max) * 0.5f;
        const hm::Mat4 worldToLocal = hm::inverse(obj.modelMatrix * hm::translate(bbOffset));
        std::memcpy(inst.worldToLocal, worldToLocal.m, sizeof(float) * 16);
        instanceData.push_back(inst);
    }
    std::vector<uint8_t> bufferData(sizeof(plain_sdf_instance) * instanceData.size() + sizeof(uint32_t) * 4, 0);
    m_sdfInstanceCount = (uint32_t)instanceData.size();
    std::memcpy(bufferData.data(), &m_sdfInstanceCount, sizeof(uint32_t));
    if (!instanceData.empty()) std::memcpy(bufferData.data() + 16, instanceData.data(), sizeof(plain_sdf_instance) * instanceData.size());
    b.setStorageBufferData(m_sdfInstanceBuffer, bufferData.data(), bufferData.size());
    if (!instanceWorldBBs.empty()) b.setStorageBufferData(m_sdfInstanceWorldBBBuffer, instanceWorldBBs.data(), instanceWorldBBs.size() * sizeof(plain_bounding_box));
}

SDFGI::IndirectLightingImages SDFGI::getIndirectLightingResults(bool tracedHalfRes) const {
    IndirectLightingImages r;
    if (tracedHalfRes) { r.Y_SH = m_indirectLightingFullRes_Y_SH; r.CoCg = m_indirectLightingFullRes_CoCg; }
    else { r.Y_SH = m_indirectDiffuseHistory_Y_SH[0]; r.CoCg = m_indirectDiffuseHistory_CoCg[0]; }
    return r;
}

void SDFGI::computeIndirectLighting(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s, const FrameIndex& fi) const {
    diffuseSDFTrace(b, d, s);
    filterIndirectDiffuse(b, d, s, fi);
}

void SDFGI::diffuseSDFTrace(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s) const {  // SDFGI.cpp:380-419
    const ImageDescription target = b.getImageDescription(m_indirectDiffuse_CoCg[0]);
    sdfInstanceCulling(b, d, (int)target.width, (int)target.height, s.traceInfluenceRadius, true);
    ComputePassExecution e;
    e.genericInfo.handle = m_diffuseSDFTracePass;
    e.genericInfo.resources.storageImages = {ImageResource(m_indirectDiffuse_Y_SH[0], 0, 0), ImageResource(m_indirectDiffuse_CoCg[0], 0, 1)};
    e.genericInfo.resources.sampledImages = {ImageResource(d.currentFrame.depthBuffer, 0, 2), ImageResource(d.worldSpaceNormals, 0, 3), ImageResource(d.skyLut, 0, 4),
                                             ImageResource(d.shadowMap, 0, 10)};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(d.lightBuffer, true, 5), StorageBufferResource(m_sdfInstanceBuffer, true, 6),
                                              StorageBufferResource(m_sdfCameraCulledTiles, true, 7), StorageBufferResource(d.sunShadowInfoBuffer, true, 9)};
    e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_sdfTraceInfluenceRangeBuffer, 8)};
    e.dispatchCount[0] = ceilDivU(target.width, 8);
    e.dispatchCount[1] = ceilDivU(target.height, 8);
    const uint32_t div = s.halfResTrace ? 2 : 1;  // the GI buffers' rows per full-resolution row
    setRows(b, e, div, e.dispatchCount[1] * 8);
    b.setComputePassExecution(e);
    // the spatial filter gathers in a world-space disc (no bound in pixels): every rank needs the whole traced image
    b.addExchange(exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "giTrace", {m_indirectDiffuse_Y_SH[0], m_indirectDiffuse_CoCg[0]}, 0, div));
}

void SDFGI::filterIndirectDiffuse(RenderBackend& b, const SDFTraceDependencies& d, const SDFTraceSettings& s, const FrameIndex& fi) const {  // SDFGI.cpp:421-536
    (void)fi;  // the reference computes historySrcIndex from FrameIndex but binds fixed indices (SDFGI.cpp:457-474)
    const ImageHandle depthSrc = s.halfResTrace ? d.depthHalfRes : d.currentFrame.depthBuffer;
    const ImageDescription target = b.getImageDescription(m_indirectDiffuse_Y_SH[1]);
    const uint32_t gx = ceilDivU(target.width, 8), gy = ceilDivU(target.height, 8);
    const uint32_t div = s.halfResTrace ? 2 : 1;
    {   // spatial filter on input
        ComputePassExecution e;
        e.genericInfo.handle = m_indirectDiffuseFilterSpatialPass[0];
        e.genericInfo.resources.storageImages = {ImageResource(m_indirectDiffuse_Y_SH[1], 0, 0), ImageResource(m_indirectDiffuse_CoCg[1], 0, 1)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_indirectDiffuse_Y_SH[0], 0, 2), ImageResource(m_indirectDiffuse_CoCg[0], 0, 3), ImageResource(depthSrc, 0, 4),
                                                 ImageResource(d.worldSpaceNormals, 0, 5)};
        e.dispatchCount[0] = gx; e.dispatchCount[1] = gy;
        // the temporal filter reads its input bilinearly at the pixel's own position, one row beyond the band: overlapped computation
        // (two more rows of the filter on either side; its inputs are gathered or within the uploaded halo) instead of a halo exchange
        setRows(b, e, div, target.height, 2);
        b.setComputePassExecution(e);
    }
    {   // temporal filter
        ComputePassExecution e;
        e.genericInfo.handle = m_indirectDiffuseFilterTemporalPass;
        e.genericInfo.resources.storageImages = {ImageResource(m_indirectDiffuse_Y_SH[0], 0, 0), ImageResource(m_indirectDiffuse_CoCg[0], 0, 1),
                                                 ImageResource(m_indirectDiffuseHistory_Y_SH[1], 0, 2), ImageResource(m_indirectDiffuseHistory_CoCg[1], 0, 3)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_indirectDiffuse_Y_SH[1], 0, 4), ImageResource(m_indirectDiffuse_CoCg[1], 0, 5),
                                                 ImageResource(m_indirectDiffuseHistory_Y_SH[0], 0, 6), ImageResource(m_indirectDiffuseHistory_CoCg[0], 0, 7),
                                                 ImageResource(d.currentFrame.motionBuffer, 0, 8), ImageResource(d.previousFrame.motionBuffer, 0, 9)};
        e.dispatchCount[0] = gx; e.dispatchCount[1] = gy;
        setRows(b, e, div, target.height);
        b.setComputePassExecution(e);
        b.addExchange(exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "giTemporal", {m_indirectDiffuseHistory_Y_SH[1], m_indirectDiffuseHistory_CoCg[1]}, 0, div));  // input of the second spatial filter
    }
    {   // spatial filter on history
        ComputePassExecution e;
        e.genericInfo.handle = m_indirectDiffuseFilterSpatialPass[1];
        e.genericInfo.resources.storageImages = {ImageResource(m_indirectDiffuseHistory_Y_SH[0], 0, 0), ImageResource(m_indirectDiffuseHistory_CoCg[0], 0, 1)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_indirectDiffuseHistory_Y_SH[1], 0, 2), ImageResource(m_indirectDiffuseHistory_CoCg[1], 0, 3), ImageResource(depthSrc, 0, 4),
                                                 ImageResource(d.worldSpaceNormals, 0, 5)};
        e.dispatchCount[0] = gx; e.dispatchCount[1] = gy;
        // read by the upscale and, reprojected, by the NEXT frame's temporal filter: the all-gather is deferred (it overlaps upscale /
        // froxels / shading / TAA / bloom). The upscale runs 8 full-resolution rows beyond the band (for the shading pass) and gathers
        // half-resolution rows floor(y / 2 - 0.25) .. + 1 plus the closest-depth texel one row further: 5 rows beyond the band here,
        // computed by this rank itself (6 with a spare; the filter's own inputs are gathered or inside the 16 uploaded halo rows)
        setRows(b, e, div, target.height, s.halfResTrace ? 6 : 8);  // full-resolution trace: the shading pass reads this image itself, 8 rows beyond the band
        b.setComputePassExecution(e);
        ExchangeRequest x = exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "giSpatial1", {m_indirectDiffuseHistory_Y_SH[0], m_indirectDiffuseHistory_CoCg[0]}, 0, div);
        x.deferred = true;
        b.addExchange(x);
    }
    if (s.halfResTrace) {  // upscale
        ComputePassExecution e;
        e.genericInfo.handle = m_indirectLightingUpscale;
        e.genericInfo.resources.storageImages = {ImageResource(m_indirectLightingFullRes_Y_SH, 0, 0), ImageResource(m_indirectLightingFullRes_CoCg, 0, 1)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_indirectDiffuseHistory_Y_SH[0], 0, 2), ImageResource(m_indirectDiffuseHistory_CoCg[0], 0, 3),
                                                 ImageResource(d.currentFrame.depthBuffer, 0, 4), ImageResource(d.depthHalfRes, 0, 5)};
        const ImageDescription full = b.getImageDescription(m_indirectLightingFullRes_Y_SH);
        e.dispatchCount[0] = ceilDivU(full.width, 8);
        e.dispatchCount[1] = ceilDivU(full.height, 8);
        setRows(b, e, 1, full.height, 8);  // the shading pass runs 8 rows beyond the band
        b.setComputePassExecution(e);
    }
}

void SDFGI::sdfInstanceCulling(RenderBackend& b, const SDFTraceDependencies& d, int targetW, int targetH, float influenceRadius, bool hiZ) const {  // SDFGI.cpp:538-630
    {
        const ViewFrustum& f = d.cameraFrustum;
        plain_camera_frustum_buffer fd{};
        auto set = [&](int i, hm::Vec3 p, hm::Vec3 n) {
            fd.frustumPoints[i][0] = p.x; fd.frustumPoints[i][1] = p.y; fd.frustumPoints[i][2] = p.z;
            fd.frustumNormals[i][0] = n.x; fd.frustumNormals[i][1] = n.y; fd.frustumNormals[i][2] = n.z;
        };
        set(0, f.l_u_f, f.top); set(1, f.l_l_f, f.bot); set(2, f.l_l_n, f.near); set(3, f.l_l_f, f.far); set(4, f.l_l_f, f.left); set(5, f.r_l_f, f.right);
        b.setUniformBufferData(m_cameraFrustumBuffer, &fd, sizeof(fd));
        uint32_t zero = 0;
        b.setStorageBufferData(m_sdfCameraFrustumCulledInstances, &zero, sizeof(zero));
        ComputePassExecution e;
        e.genericInfo.handle = m_sdfCameraFrustumCulling;
        e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_sdfInstanceBuffer, true, 0), StorageBufferResource(m_sdfCameraFrustumCulledInstances, false, 2),
                                                  StorageBufferResource(m_sdfInstanceWorldBBBuffer, true, 3)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_cameraFrustumBuffer, 1), UniformBufferResource(m_sdfTraceInfluenceRangeBuffer, 4)};
        e.dispatchCount[0] = ceilDivU(m_sdfInstanceCount, 64);
        b.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = hiZ ? m_sdfCameraTileCullingHiZ : m_sdfCameraTileCulling;
        const uint32_t tileCount[2] = {ceilDivU((uint32_t)targetW, sdfCameraCullingTileSize), ceilDivU((uint32_t)targetH, sdfCameraCullingTileSize)};
        e.dispatchCount[0] = ceilDivU(tileCount[0], 8);
        e.dispatchCount[1] = ceilDivU(tileCount[1], 8);
        e.pushConstants = dataToCharArray(tileCount, sizeof(tileCount));
        e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_sdfCameraFrustumCulledInstances, true, 0), StorageBufferResource(m_sdfInstanceWorldBBBuffer, true, 1),
                                                  StorageBufferResource(m_sdfCameraCulledTiles, false, 2)};
        b.setUniformBufferData(m_sdfTraceInfluenceRangeBuffer, &influenceRadius, sizeof(influenceRadius));
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_sdfTraceInfluenceRangeBuffer, 3)};
        const uint32_t depthPyramidMipLevel = 4;  // log2(32) - 1, SDFGI.cpp:621-623
        e.genericInfo.resources.sampledImages = {ImageResource(d.depthMinMaxPyramid, depthPyramidMipLevel, 4)};
        b.setComputePassExecution(e);
    }
}

// =============================== TAA ===============================
void TAA::init(RenderBackend& b, int w, int h, const TAASettings& s) {
    for (int i = 0; i < 2; i++) m_historyBuffers[i] = b.createImage(imageDesc2D(w, h, PLAIN_FORMAT_R11G11B10_UFLOAT, SS), nullptr, 0);
    m_taaResolveWeightBuffer = b.createUniformBuffer(sizeof(float) * 9);
    const uint32_t clip = s.useClipping, dil = s.useMotionVectorDilation, tm = s.filterUseTonemapping;
    m_temporalFilterPass = makePass(b, "Temporal filtering", "temporalFilter.comp", {specConst(0, clip), specConst(1, dil), specConst(2, (int)s.historySamplingTech), specConst(3, tm)});
    const uint32_t ssTm = s.supersampleUseTonemapping;
    m_temporalSupersamplingPass = makePass(b, "Temporal supersampling", "temporalSupersampling.comp", {specConst(0, ssTm)});
    for (int i = 0; i < 2; i++) m_sceneLuminance[i] = b.createImage(imageDesc2D(w, h, PLAIN_FORMAT_R8, SS), nullptr, 0);
    m_colorToLuminancePass = makePass(b, "Color to Luminance", "colorToLuminance.comp");
}
void TAA::updateSettings(RenderBackend& b, const TAASettings& s) {  // TAA.cpp:80-83
    const uint32_t clip = s.useClipping, dil = s.useMotionVectorDilation, tm = s.filterUseTonemapping, ssTm = s.supersampleUseTonemapping;
    ShaderDescription filter;
    filter.srcPathRelative = "temporalFilter.comp";
    filter.specialisationConstants = {specConst(0, clip), specConst(1, dil), specConst(2, (int)s.historySamplingTech), specConst(3, tm)};
    b.updateComputePassShaderDescription(m_temporalFilterPass, filter);
    ShaderDescription supersampling;
    supersampling.srcPathRelative = "temporalSupersampling.comp";
    supersampling.specialisationConstants = {specConst(0, ssTm)};
    b.updateComputePassShaderDescription(m_temporalSupersamplingPass, supersampling);
}
void TAA::computeTemporalSuperSampling(RenderBackend& b, const FrameRenderTargets& cur, const FrameRenderTargets& last, ImageHandle target, const FrameIndex& fi) const {  // TAA.cpp:85-137
    const ImageDescription td = b.getImageDescription(target);
    const size_t m2 = fi.mod2();
    const ImageHandle currentLuminance = m_sceneLuminance[m2], historyLuminance = m_sceneLuminance[(m2 + 1) % 2];
    {   // scene luminance
        ComputePassExecution e;
        e.genericInfo.handle = m_colorToLuminancePass;
        e.genericInfo.resources.storageImages = {ImageResource(currentLuminance, 0, 1)};
        e.genericInfo.resources.sampledImages = {ImageResource(cur.colorBuffer, 0, 0)};
        e.dispatchCount[0] = ceilDivU(td.width, 8);
        e.dispatchCount[1] = ceilDivU(td.height, 8);
        b.setComputePassExecution(e);
    }
    {   // temporal supersampling
        ComputePassExecution e;
        e.genericInfo.handle = m_temporalSupersamplingPass;
        e.genericInfo.resources.storageImages = {ImageResource(target, 0, 3)};
        e.genericInfo.resources.sampledImages = {ImageResource(cur.colorBuffer, 0, 1), ImageResource(last.colorBuffer, 0, 2), ImageResource(cur.motionBuffer, 0, 4), ImageResource(cur.depthBuffer, 0, 5),
                                                 ImageResource(last.depthBuffer, 0, 6), ImageResource(currentLuminance, 0, 7), ImageResource(historyLuminance, 0, 8)};
        e.dispatchCount[0] = ceilDivU(td.width, 8);
        e.dispatchCount[1] = ceilDivU(td.height, 8);
        b.setComputePassExecution(e);
    }
}
void TAA::computeTemporalFilter(RenderBackend& b, ImageHandle colorSrc, const FrameRenderTargets& cur, ImageHandle target, const FrameIndex& fi) const {  // TAA.cpp:139-166
    const size_t m2 = fi.mod2();
    ComputePassExecution e;
    e.genericInfo.handle = m_temporalFilterPass;
    e.genericInfo.resources.storageImages = {ImageResource(target, 0, 1), ImageResource(m_historyBuffers[(m2 + 1) % 2], 0, 2)};
    e.genericInfo.resources.sampledImages = {ImageResource(colorSrc, 0, 0), ImageResource(m_historyBuffers[m2], 0, 3), ImageResource(cur.motionBuffer, 0, 4), ImageResource(cur.depthBuffer, 0, 5)};
    e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_taaResolveWeightBuffer, 6)};
    const ImageDescription td = b.getImageDescription(target);
    e.dispatchCount[0] = ceilDivU(td.width, 8);
    e.dispatchCount[1] = ceilDivU(td.height, 8);
    setRows(b, e, 1, td.height, 4);  // 4 rows beyond the band: the first bloom downsample reads +-3 rows of the resolved image
    b.setComputePassExecution(e);
    // next frame's resolve reads the history at reprojected positions
    ExchangeRequest x = exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "taaHistory", {m_historyBuffers[(m2 + 1) % 2]}, 0, 1);
    x.deferred = true;
    b.addExchange(x);
}
hm::Vec2 TAA::computeProjectionMatrixJitter(const FrameIndex& fi) const {  // TAA.cpp:168-170
    hm::Vec2 h = hammersley2D((uint32_t)fi.mod8());
    hm::Vec2 r; r.x = 2.f * h.x - 1.f; r.y = 2.f * h.y - 1.f;
    return r;
}
hm::Mat4 TAA::applyProjectionMatrixJitter(const hm::Mat4& projection, hm::Vec2 offset) const {  // TAA.cpp:172-179
    hm::Mat4 j = projection;
    j.at(2, 0) = offset.x;
    j.at(2, 1) = offset.y;
    return j;
}
void TAA::updateTaaResolveWeights(RenderBackend& b, hm::Vec2 jitter) {  // TAA.cpp:181-202
    std::array<float, 9> weights{};
    int index = 0;
    float totalWeight = 0.f;
    for (int y = -1; y <= 1; y++)
        for (int x = -1; x <= 1; x++) {
            const float dx = jitter.x - (float)x, dy = jitter.y - (float)y;
            const float d = dm::sqrt_(dx * dx + dy * dy);
            const float w = dm::exp(-2.29f * d * d);
            weights[index++] = w;
            totalWeight += w;
        }
    for (float& w : weights) w /= totalWeight;
    m_lastResolveWeights = weights;
    b.setUniformBufferData(m_taaResolveWeightBuffer, weights.data(), sizeof(float) * 9);
}

// =============================== Sky ===============================
void Sky::init(RenderBackend& b) {
    m_skyTransmissionLut = b.createImage(imageDesc2D(128, 128, PLAIN_FORMAT_R11G11B10_UFLOAT, SS), nullptr, 0);
    m_skyMultiscatterLut = b.createImage(imageDesc2D(32, 32, PLAIN_FORMAT_R11G11B10_UFLOAT, SS), nullptr, 0);
    m_skyLut = b.createImage(imageDesc2D(200, 100, PLAIN_FORMAT_R11G11B10_UFLOAT, SS), nullptr, 0);
    m_atmosphereSettingsBuffer = b.createUniformBuffer(sizeof(plain_atmosphere_settings));
    m_skyTransmissionLutPass = makePass(b, "Sky transmission lut", "skyTransmissionLut.comp");
    m_skyMultiscatterLutPass = makePass(b, "Sky multiscatter lut", "skyMultiscatterLut.comp");
    m_skyLutPass = makePass(b, "Sky lut", "skyLut.comp");
}
void Sky::updateTransmissionLut(RenderBackend& b) const {  // Sky.cpp:260-272
    ComputePassExecution e;
    e.genericInfo.handle = m_skyTransmissionLutPass;
    e.genericInfo.resources.storageImages = {ImageResource(m_skyTransmissionLut, 0, 0)};
    e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_atmosphereSettingsBuffer, 1)};
    e.dispatchCount[0] = 128 / 8; e.dispatchCount[1] = 128 / 8;
    b.setComputePassExecution(e);
}
void Sky::updateSkyLut(RenderBackend& b, StorageBufferHandle lightBuffer, const plain_atmosphere_settings& a) const {  // Sky.cpp:274-316
    b.setUniformBufferData(m_atmosphereSettingsBuffer, &a, sizeof(a));
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_skyMultiscatterLutPass;
        e.genericInfo.resources.storageImages = {ImageResource(m_skyMultiscatterLut, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_skyTransmissionLut, 0, 1)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_atmosphereSettingsBuffer, 3)};
        e.dispatchCount[0] = 32 / 8; e.dispatchCount[1] = 32 / 8;
        b.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_skyLutPass;
        e.genericInfo.resources.storageImages = {ImageResource(m_skyLut, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_skyTransmissionLut, 0, 1), ImageResource(m_skyMultiscatterLut, 0, 2)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_atmosphereSettingsBuffer, 4)};
        e.genericInfo.resources.storageBuffers = {StorageBufferResource(lightBuffer, true, 5)};
        e.dispatchCount[0] = 200 / 8; e.dispatchCount[1] = 100 / 8;  // 25 x 12: rows 96-99 stay unwritten (Sky.cpp:311-312)
        b.setComputePassExecution(e);
    }
}
hm::Mat4 Sky::sunSpriteModelMatrix(hm::Vec2 sunDirection) const {  // Sky.cpp:247-259
    const float sunAngularDiameter = 0.535f;
    const float spriteScale = dm::tan(hm::radians(sunAngularDiameter * 0.5f));
    const hm::Mat4 scaleM = hm::scale(hm::Vec3(spriteScale, spriteScale, 1.f));
    const hm::Mat4 lat = hm::rotate(hm::radians(sunDirection.y + 90.f), hm::Vec3(-1.f, 0.f, 0.f));
    const hm::Mat4 lon = hm::rotate(hm::radians(sunDirection.x + -90.f), hm::Vec3(0.f, -1.f, 0.f));
    return lon * lat * scaleM;
}

// =============================== Volumetrics ===============================
static uint32_t hashU(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
void Volumetrics::init(RenderBackend& b, int w, int h, uint32_t noiseSeed) {
    const uint32_t fx = ceilDivU((uint32_t)w, 8), fy = ceilDivU((uint32_t)h, 8), fz = 64;  // computeVolumetricLightingFroxelResolution
    m_scatteringTransmittanceVolume = b.createImage(imageDesc3D(fx, fy, fz, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    m_volumetricIntegrationVolume = b.createImage(imageDesc3D(fx, fy, fz, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    m_volumeMaterialVolume = b.createImage(imageDesc3D(fx, fy, fz, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    for (int i = 0; i < 2; i++) m_volumetricLightingHistory[i] = b.createImage(imageDesc3D(fx, fy, fz, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    // 32^3 R8 density noise: fixture input (the reference's Perlin noise is seeded by C rand()); smooth periodic value noise
    const int N = 32, L = 8;
    std::vector<uint8_t> noise((size_t)N * N * N);
    auto lattice = [&](int x, int y, int z) { return (float)(hashU(noiseSeed ^ hashU((uint32_t)((x % L) + L * ((y % L) + L * (z % L))))) & 0xffff) / 65535.f; };
    for (int z = 0; z < N; z++)
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++) {
                float fxx = (float)x * L / N, fyy = (float)y * L / N, fzz = (float)z * L / N;
                int x0 = (int)fxx, y0 = (int)fyy, z0 = (int)fzz;
                float tx = fxx - x0, ty = fyy - y0, tz = fzz - z0;
                float v = 0.f;
                for (int k = 0; k < 8; k++) {
                    int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
                    v += lattice(x0 + dx, y0 + dy, z0 + dz) * (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
                }
                noise[(size_t)(z * N + y) * N + x] = (uint8_t)(v * 255.f + 0.5f);
            }
    m_perlinNoise3D = b.createImage(imageDesc3D(N, N, N, PLAIN_FORMAT_R8, PLAIN_USAGE_SAMPLED), noise.data(), noise.size());
    m_volumetricsSettingsUniforms = b.createUniformBuffer(sizeof(plain_volumetric_lighting_settings));
    m_froxelVolumeMaterialPass = makePass(b, "Froxel volume material", "froxelVolumeMaterial.comp");
    m_froxelScatteringTransmittancePass = makePass(b, "Froxel light scattering", "froxelLightScattering.comp");
    m_volumetricLightingReprojection = makePass(b, "Volumetric lighting reprojection", "volumeLightingReprojection.comp");
    m_volumetricLightingIntegration = makePass(b, "Volumetric light integration", "volumetricLightingIntegration.comp");
}
void Volumetrics::computeVolumetricLighting(RenderBackend& b, const VolumetricsSettings& s, const WindSettings& wind, const Dependencies& d, const FrameIndex& fi, float deltaTime) {  // Volumetrics.cpp:136-247
    plain_volumetric_lighting_settings u{};
    u.sampleOffset = hammersley2D((uint32_t)fi.mod8()).x - 0.5f;
    m_windSampleOffset = m_windSampleOffset + wind.vector * wind.speed * deltaTime;
    u.windSampleOffset[0] = m_windSampleOffset.x; u.windSampleOffset[1] = m_windSampleOffset.y; u.windSampleOffset[2] = m_windSampleOffset.z;
    for (int i = 0; i < 3; i++) u.scatteringCoefficients[i] = s.scatteringCoefficients[i];
    u.maxDistance = s.maxDistance; u.absorptionCoefficient = s.absorptionCoefficient; u.baseDensity = s.baseDensity;
    u.densityNoiseRange = s.densityNoiseRange; u.densityNoiseScale = s.densityNoiseScale; u.phaseFunctionG = s.phaseFunctionG;
    b.setUniformBufferData(m_volumetricsSettingsUniforms, &u, sizeof(u));
    const ImageDescription fd = b.getImageDescription(m_volumeMaterialVolume);
    const uint32_t g4[3] = {ceilDivU(fd.width, 4), ceilDivU(fd.height, 4), ceilDivU(fd.depth, 4)};
    // Row sharding: every pass works on the rank's band of froxel rows (8 screen rows each) extended by froxelOverlap rows,
    // which covers what the rank's pixels read from the integrated volume (shading runs on the band +-8 screen rows = 1 froxel
    // row, its lookup is jittered by +-0.0065 of the screen height = 1.8 froxel rows at 4K, plus the trilinear footprint). The
    // reprojection reads LAST frame's result at reprojected positions, anywhere in the volume: the band rows of this frame's
    // result are all-gathered for the next frame.
    const uint32_t froxelRowPixels = 8, froxelOverlap = 5;
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_froxelVolumeMaterialPass;
        e.genericInfo.resources.storageImages = {ImageResource(m_volumeMaterialVolume, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_perlinNoise3D, 0, 1)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_volumetricsSettingsUniforms, 2)};
        for (int i = 0; i < 3; i++) e.dispatchCount[i] = g4[i];
        setRows(b, e, froxelRowPixels, fd.height, froxelOverlap);
        b.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_froxelScatteringTransmittancePass;
        e.genericInfo.resources.storageImages = {ImageResource(m_scatteringTransmittanceVolume, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(d.shadowMap, 0, 1), ImageResource(m_volumeMaterialVolume, 0, 2)};
        e.genericInfo.resources.storageBuffers = {StorageBufferResource(d.sunShadowInfoBuffer, true, 3), StorageBufferResource(d.lightBuffer, true, 4)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_volumetricsSettingsUniforms, 5)};
        for (int i = 0; i < 3; i++) e.dispatchCount[i] = g4[i];
        setRows(b, e, froxelRowPixels, fd.height, froxelOverlap);
        b.setComputePassExecution(e);
    }
    const size_t m2 = fi.mod2();
    const ImageHandle reprojectionTarget = m_volumetricLightingHistory[m2];
    const ImageHandle reprojectionHistory = m_volumetricLightingHistory[(m2 + 1) % 2];
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_volumetricLightingReprojection;
        e.genericInfo.resources.storageImages = {ImageResource(reprojectionTarget, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(m_scatteringTransmittanceVolume, 0, 1), ImageResource(reprojectionHistory, 0, 2)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_volumetricsSettingsUniforms, 3)};
        for (int i = 0; i < 3; i++) e.dispatchCount[i] = g4[i];
        setRows(b, e, froxelRowPixels, fd.height, froxelOverlap);
        b.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_volumetricLightingIntegration;
        e.genericInfo.resources.storageImages = {ImageResource(m_volumetricIntegrationVolume, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(reprojectionTarget, 0, 1)};
        e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_volumetricsSettingsUniforms, 2)};
        e.dispatchCount[0] = ceilDivU(fd.width, 8);
        e.dispatchCount[1] = ceilDivU(fd.height, 8);
        setRows(b, e, froxelRowPixels, fd.height, froxelOverlap);
        b.setComputePassExecution(e);
    }
    // next frame's reprojection history: every rank contributes the froxel rows of its band (in every z slice)
    ExchangeRequest x = exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "froxelHistory", {reprojectionTarget}, 0, froxelRowPixels);
    x.deferred = true;
    b.addExchange(x);
}

// =============================== Bloom ===============================
static const int bloomMipCount = 6;
void Bloom::init(RenderBackend& b) {
    for (int i = 0; i < bloomMipCount - 1; i++) m_bloomDownsamplePasses.push_back(makePass(b, ("Bloom downsample mip " + std::to_string(i + 1)).c_str(), "bloomDownsample.comp"));
    for (int i = 0; i < bloomMipCount - 1; i++) {
        const uint32_t isLowestMip = i == 0 ? 1u : 0u;
        m_bloomUpsamplePasses.push_back(makePass(b, ("Bloom Upsample mip " + std::to_string(bloomMipCount - 2 - i)).c_str(), "bloomUpsample.comp", {specConst(0, isLowestMip)}));
    }
    m_applyBloomPass = makePass(b, "Apply bloom", "applyBloom.comp");
}
void Bloom::computeBloom(RenderBackend& b, ImageHandle targetImage, const BloomSettings& s) const {  // Bloom.cpp:56-144
    const ImageDescription td = b.getImageDescription(targetImage);
    const int width = (int)td.width, height = (int)td.height;
    const ImageDescription desc = imageDesc2D(width, height, PLAIN_FORMAT_R11G11B10_UFLOAT, SS, PLAIN_MIPS_MANUAL, bloomMipCount);
    const ImageHandle downscaleTexture = b.createTemporaryImage(desc);
    auto mipRes = [&](int mip, int v) { int r = v / (1 << mip); return r > 1 ? r : 1; };  // resolutionFromMip, MathUtils.cpp:21-23
    for (int i = 0; i < (int)m_bloomDownsamplePasses.size(); i++) {
        ComputePassExecution e;
        e.genericInfo.handle = m_bloomDownsamplePasses[i];
        const int sourceMip = i, targetMip = i + 1;
        e.genericInfo.resources.storageImages = {ImageResource(downscaleTexture, targetMip, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(i == 0 ? targetImage : downscaleTexture, sourceMip, 1)};
        e.dispatchCount[0] = ceilDivU((uint32_t)mipRes(targetMip, width), 8);
        e.dispatchCount[1] = ceilDivU((uint32_t)mipRes(targetMip, height), 8);
        if (i == 0) setRows(b, e, 2, (uint32_t)mipRes(1, height));  // mip 1 from the rank's own rows; the small mips below are computed by every rank
        b.setComputePassExecution(e);
        if (i == 0) b.addExchange(exchangeRows(PLAIN_EXCHANGE_ALLGATHER_ROWS, "bloomMip1", {downscaleTexture}, 1, 2));
    }
    const ImageHandle upscaleTexture = b.createTemporaryImage(desc);
    for (int i = 0; i < (int)m_bloomUpsamplePasses.size(); i++) {
        ComputePassExecution e;
        e.genericInfo.handle = m_bloomUpsamplePasses[i];
        const int targetMip = bloomMipCount - 2 - i, sourceMip = targetMip + 1;
        e.genericInfo.resources.storageImages = {ImageResource(upscaleTexture, targetMip, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(upscaleTexture, sourceMip, 1), ImageResource(downscaleTexture, sourceMip, 2)};
        e.dispatchCount[0] = ceilDivU((uint32_t)mipRes(targetMip, width), 8);
        e.dispatchCount[1] = ceilDivU((uint32_t)mipRes(targetMip, height), 8);
        e.pushConstants = dataToCharArray(&s.radius, sizeof(s.radius));
        if (targetMip == 0) setRows(b, e, 1, (uint32_t)height, 1);  // +-1 row: applyBloom samples it bilinearly at the pixel centre
        // mip 1 reads only replicated mips (>= 2): every rank computes the rows its mip-0 window samples (the band +-1 row is
        // +-1 row of mip 1; the 9-tap tent reaches 1.5 texels + the bilinear footprint: 5 rows cover it)
        if (targetMip == 1) setRows(b, e, 2, (uint32_t)mipRes(1, height), 5);
        b.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_applyBloomPass;
        e.genericInfo.resources.storageImages = {ImageResource(targetImage, 0, 0)};
        e.genericInfo.resources.sampledImages = {ImageResource(upscaleTexture, 0, 1)};
        e.dispatchCount[0] = ceilDivU((uint32_t)width, 8);
        e.dispatchCount[1] = ceilDivU((uint32_t)height, 8);
        e.pushConstants = dataToCharArray(&s.strength, sizeof(s.strength));
        setRows(b, e, 1, (uint32_t)height);
        b.setComputePassExecution(e);
    }
    m_lastDownscaleTexture = downscaleTexture;
    m_lastUpscaleTexture = upscaleTexture;
}
