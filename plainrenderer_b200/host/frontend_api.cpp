// frontend_api.cpp - C entry points (include/plain_frontend.h) over the RenderFrontend mirror.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include "RenderFrontend.h"
#include "plain_frontend.h"

struct plain_frontend {
    RenderFrontend fe;
    std::vector<RenderObject> scene;
    std::string lastError;
    plain_frontend_settings settings;
};

#define FE_TRY(fe, body)                          \
    try { body; return 0; }                       \
    catch (const std::exception& e) { (fe)->lastError = e.what(); return 1; }

extern "C" {

void PLAIN_FE(default_settings)(plain_frontend_settings* s, uint32_t width, uint32_t height) {
    std::memset(s, 0, sizeof(*s));
    s->width = width; s->height = height;
    s->diffuse_brdf = 2; s->direct_multiscatter = 0; s->indirect_lighting_tech = 0; s->use_geometry_aa = 1; s->sun_shadow_cascade_count = 3;
    s->half_res_trace = 1; s->strict_influence_radius_cutoff = 1; s->trace_influence_radius = 5.f;
    s->taa_enabled = 1; s->taa_use_clipping = 1; s->taa_use_motion_vector_dilation = 1; s->taa_history_sampling_tech = 4; s->taa_filter_use_tonemapping = 1;
    s->bloom_enabled = 1; s->bloom_strength = 0.05f; s->bloom_radius = 1.5f;
    s->sun_direction_deg[0] = 0.f; s->sun_direction_deg[1] = 0.f;
    s->camera_fov_deg = 35.f; s->camera_near = 0.1f; s->camera_far = 300.f;
    s->noise_seed = 0x504c4149u;
    s->taa_use_separate_supersampling = 0; s->taa_supersample_use_tonemapping = 1;
    s->sdf_debug_mode = 0; s->sdf_debug_show_tile_usage_with_hiz = 1; s->sdf_debug_use_influence_radius = 0;
    s->raster_inputs = 0;
}

int PLAIN_FE(create)(int device, const plain_frontend_settings* s, plain_frontend** out) {
    if (!s || !out) return 1;
    plain_frontend* p = new plain_frontend();
    p->settings = *s;
    RenderFrontend& f = p->fe;
    f.m_shadingConfig.diffuseBRDF = (DiffuseBRDF)s->diffuse_brdf;
    f.m_shadingConfig.directMultiscatter = (DirectSpecularMultiscattering)s->direct_multiscatter;
    f.m_shadingConfig.indirectLightingTech = (IndirectLightingTech)s->indirect_lighting_tech;
    f.m_shadingConfig.useGeometryAA = s->use_geometry_aa != 0;
    f.m_shadingConfig.sunShadowCascadeCount = s->sun_shadow_cascade_count;
    f.m_sdfTraceSettings.halfResTrace = s->half_res_trace != 0;
    f.m_sdfTraceSettings.strictInfluenceRadiusCutoff = s->strict_influence_radius_cutoff != 0;
    f.m_sdfTraceSettings.traceInfluenceRadius = s->trace_influence_radius;
    f.m_taaSettings.enabled = s->taa_enabled != 0;
    f.m_taaSettings.useClipping = s->taa_use_clipping != 0;
    f.m_taaSettings.useMotionVectorDilation = s->taa_use_motion_vector_dilation != 0;
    f.m_taaSettings.historySamplingTech = (HistorySamplingTech)s->taa_history_sampling_tech;
    f.m_taaSettings.filterUseTonemapping = s->taa_filter_use_tonemapping != 0;
    f.m_taaSettings.useSeparateSupersampling = s->taa_use_separate_supersampling != 0;
    f.m_taaSettings.supersampleUseTonemapping = s->taa_supersample_use_tonemapping != 0;
    f.m_sdfDebugSettings.visualisationMode = (SDFVisualisationMode)s->sdf_debug_mode;
    f.m_sdfDebugSettings.showCameraTileUsageWithHiZ = s->sdf_debug_show_tile_usage_with_hiz != 0;
    f.m_sdfDebugSettings.useInfluenceRadiusForDebug = s->sdf_debug_use_influence_radius != 0;
    f.m_rasterInputs = s->raster_inputs != 0;
    f.m_bloomSettings.enabled = s->bloom_enabled != 0;
    f.m_bloomSettings.strength = s->bloom_strength;
    f.m_bloomSettings.radius = s->bloom_radius;
    f.m_cameraIntrinsic.fov = s->camera_fov_deg;
    f.m_cameraIntrinsic.near = s->camera_near;
    f.m_cameraIntrinsic.far = s->camera_far;
    f.m_shardRank = s->shard_rank;
    f.m_shardCount = s->shard_count > 1 ? s->shard_count : 1;
    try {
        f.setup(device, s->width, s->height, s->noise_seed);
        f.m_sunDirection.x = s->sun_direction_deg[0];
        f.m_sunDirection.y = s->sun_direction_deg[1];
    } catch (const std::exception& e) {
        static std::string createError;
        createError = e.what();
        fprintf(stderr, "plain_frontend create failed: %s\n", e.what());
        delete p;
        return 1;
    }
    *out = p;
    return 0;
}
void PLAIN_FE(destroy)(plain_frontend* fe) {
    if (!fe) return;
    fe->fe.shutdown();
    delete fe;
}
const char* PLAIN_FE(last_error)(plain_frontend* fe) { return fe ? fe->lastError.c_str() : "null frontend"; }
plain_ctx* PLAIN_FE(backend)(plain_frontend* fe) { return fe->fe.backend.context(); }

int PLAIN_FE(register_sdf_mesh)(plain_frontend* fe, const uint16_t* texels, uint32_t rx, uint32_t ry, uint32_t rz, const float mn[3], const float mx[3], const float albedo[3], uint32_t* out) {
    FE_TRY(fe, {
        hm::AABB bb;
        bb.min = hm::Vec3(mn[0], mn[1], mn[2]);
        bb.max = hm::Vec3(mx[0], mx[1], mx[2]);
        *out = fe->fe.registerSdfMesh(texels, rx, ry, rz, bb, hm::Vec3(albedo[0], albedo[1], albedo[2]));
    });
}
int PLAIN_FE(set_mesh_geometry)(plain_frontend* fe, uint32_t mesh, const plain_mesh_binary* g, uint32_t albedoTexture, uint32_t normalTexture, uint32_t specularTexture) {
    FE_TRY(fe, {
        MeshBinary mb;
        mb.indexCount = g->index_count; mb.vertexCount = g->vertex_count;
        const bool wide = !(g->index_count < 65535u);
        mb.indexBuffer.resize((size_t)g->index_count * (wide ? 2 : 1));
        std::memcpy(mb.indexBuffer.data(), g->index_buffer, (size_t)g->index_count * (wide ? 4 : 2));
        mb.vertexBuffer.assign((const uint8_t*)g->vertex_buffer, (const uint8_t*)g->vertex_buffer + (size_t)g->vertex_count * 28);
        Material m;
        m.albedoTextureIndex = albedoTexture; m.normalTextureIndex = normalTexture; m.specularTextureIndex = specularTexture;
        fe->fe.setMeshGeometry(mesh, mb, &m);
    });
}
int PLAIN_FE(set_scene)(plain_frontend* fe, uint32_t n, const uint32_t* meshes, const float* mats, const float* bmin, const float* bmax) {
    FE_TRY(fe, {
        fe->scene.clear();
        for (uint32_t i = 0; i < n; i++) {
            RenderObject o;
            o.mesh = meshes[i];
            std::memcpy(o.modelMatrix.m, mats + 16 * i, sizeof(float) * 16);
            o.previousModelMatrix = o.modelMatrix;  // static objects (RuntimeScene.h keeps the last frame's matrix per object)
            o.bbWorld.min = hm::Vec3(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2]);
            o.bbWorld.max = hm::Vec3(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2]);
            fe->scene.push_back(o);
        }
    });
}

static void beginFrame(plain_frontend* fe, const plain_camera_extrinsic* cam, float time, float deltaTime, const plain_frame_inputs* in) {
    RenderFrontend& f = fe->fe;
    // a row-sharded caller that uploads only its rows also uploads only its rows of the motion vectors: the ranks all-gather them over
    // NVLink (they are read at reprojected positions) instead of every rank pulling the whole buffer through the host
    f.m_motionRowsOnly = in && in->motion && f.backend.shard.active() && !(in->row_begin == 0 && in->row_end == 0);
    f.markNewFrame(time, deltaTime);
    f.prepareNewFrame();
    if (in) {  // what depthPrepass / sunShadow / the G-buffer producer write during the frame
        const FrameRenderTargets& t = f.currentTargets();
        const size_t W = f.m_screenWidth, H = f.m_screenHeight;
        const uint32_t r0 = (in->row_begin == 0 && in->row_end == 0) ? 0u : (in->row_begin < H ? in->row_begin : (uint32_t)H);
        const uint32_t r1 = (in->row_begin == 0 && in->row_end == 0) ? (uint32_t)H : (in->row_end < H ? in->row_end : (uint32_t)H);
        auto upRows = [&](ImageHandle h, const void* p, size_t bytesPerTexel) {  // rows [r0, r1) of a full-frame host buffer
            if (!p || r1 <= r0) return;
            const size_t pitch = W * bytesPerTexel;
            const unsigned char* src = (const unsigned char*)p + pitch * r0;
            if (in->async_upload) f.backend.writeImageRowsAsync(h, 0, r0, r1, src, pitch * (r1 - r0));
            else { f.backend.writeImageRowsAsync(h, 0, r0, r1, src, pitch * (r1 - r0)); f.backend.waitForGPUIdle(); }
        };
        auto up = [&](ImageHandle h, const void* p, size_t size) {
            if (!p) return;
            if (in->async_upload) f.backend.writeImageAsync(h, 0, p, size);
            else f.backend.writeImage(h, 0, p, size);
        };
        upRows(t.depthBuffer, in->depth, 4);
        upRows(f.worldSpaceNormalImage(), in->normal, 4);
        upRows(f.gbuffer(), in->gbuffer, 16);
        if (f.m_motionRowsOnly) upRows(t.motionBuffer, in->motion, 4);
        else up(t.motionBuffer, in->motion, W * H * 4);
        for (int i = 0; i < 4; i++) up(f.m_shadowMaps[i], in->shadow_maps[i], (size_t)2048 * 2048 * 2);
    }
    CameraExtrinsic e;
    e.position = hm::Vec3(cam->position[0], cam->position[1], cam->position[2]);
    e.forward = hm::Vec3(cam->forward[0], cam->forward[1], cam->forward[2]);
    e.right = hm::Vec3(cam->right[0], cam->right[1], cam->right[2]);
    e.up = hm::Vec3(cam->up[0], cam->up[1], cam->up[2]);
    f.setCameraExtrinsic(e);
    f.prepareForDrawcalls();
    f.renderScene(fe->scene);
}
int PLAIN_FE(render_frame)(plain_frontend* fe, const plain_camera_extrinsic* cam, float time, float deltaTime, const plain_frame_inputs* in) {
    FE_TRY(fe, {
        if (fe->fe.backend.shard.active()) throw std::runtime_error("render_frame: a row-sharded frontend is driven with begin_frame / run_segment");
        beginFrame(fe, cam, time, deltaTime, in);
        fe->fe.renderFrame();
    });
}
int PLAIN_FE(begin_frame)(plain_frontend* fe, const plain_camera_extrinsic* cam, float time, float deltaTime, const plain_frame_inputs* in) {
    FE_TRY(fe, { beginFrame(fe, cam, time, deltaTime, in); });
}
int PLAIN_FE(run_segment)(plain_frontend* fe, plain_exchange* out) {
    try {
        std::memset(out, 0, sizeof(*out));
        return fe->fe.renderFrameSegment(out) ? 1 : 0;  // 1: perform the exchange in *out, then call again
    } catch (const std::exception& e) {
        fe->lastError = e.what();
        return -1;
    }
}
int PLAIN_FE(set_peer_exchange)(plain_frontend* fe, int32_t enabled) {
    fe->fe.backend.m_peerExchange = enabled != 0;
    return 0;
}
void PLAIN_FE(shard_band)(uint32_t fullHeight, uint32_t count, uint32_t rank, uint32_t divisor, uint32_t rows, uint32_t* a, uint32_t* b) {
    uint32_t y0 = 0, y1 = fullHeight;
    if (count > 1) shardBandRows(fullHeight, count, rank, &y0, &y1);
    if (divisor < 1) divisor = 1;
    uint32_t lo = y0 / divisor, hi = (y1 + divisor - 1) / divisor;
    if (hi > rows) hi = rows;
    if (lo > hi) lo = hi;
    *a = lo; *b = hi;
}
int PLAIN_FE(read_output)(plain_frontend* fe, void* out, size_t size, int32_t asyncPinned) {
    FE_TRY(fe, {
        ImageHandle h = fe->fe.backend.getSwapchainInputImage();
        if (asyncPinned) fe->fe.backend.readImageAsync(h, 0, out, size);
        else fe->fe.backend.readImage(h, 0, out, size);
    });
}
int PLAIN_FE(read_output_rows)(plain_frontend* fe, void* outFullFrame, uint32_t r0, uint32_t r1, int32_t asyncPinned) {
    FE_TRY(fe, {
        ImageHandle h = fe->fe.backend.getSwapchainInputImage();
        const size_t pitch = (size_t)fe->fe.m_screenWidth * 4;
        if (r1 > fe->fe.m_screenHeight) r1 = fe->fe.m_screenHeight;
        if (r1 > r0) fe->fe.backend.readImageRowsAsync(h, 0, r0, r1, (unsigned char*)outFullFrame + pitch * r0, pitch * (r1 - r0));
        if (!asyncPinned) fe->fe.backend.waitForGPUIdle();
    });
}

int PLAIN_FE(get_image)(plain_frontend* fe, const char* name, plain_image_handle* out) {
    RenderFrontend& f = fe->fe;
    std::map<std::string, ImageHandle> m = {
        {"color0", f.m_frameRenderTargets[0].colorBuffer}, {"color1", f.m_frameRenderTargets[1].colorBuffer},
        {"depth0", f.m_frameRenderTargets[0].depthBuffer}, {"depth1", f.m_frameRenderTargets[1].depthBuffer},
        {"motion0", f.m_motionBuffers[0]}, {"motion1", f.m_motionBuffers[1]}, {"motion2", f.m_motionBuffers[2]},
        {"post0", f.m_postProcessBuffers[0]}, {"post1", f.m_postProcessBuffers[1]}, {"normal", f.worldSpaceNormalImage()}, {"gbuffer", f.gbuffer()},
        {"depthHalf", f.m_depthHalfRes}, {"hiz", f.m_minMaxDepthPyramid}, {"brdfLut", f.m_brdfLut},
        {"skyTransmission", f.m_sky.m_skyTransmissionLut}, {"skyMultiscatter", f.m_sky.m_skyMultiscatterLut}, {"skyLut", f.m_sky.m_skyLut},
        {"shadow0", f.m_shadowMaps[0]}, {"shadow1", f.m_shadowMaps[1]}, {"shadow2", f.m_shadowMaps[2]}, {"shadow3", f.m_shadowMaps[3]},
        {"giY0", f.m_sdfGi.m_indirectDiffuse_Y_SH[0]}, {"giY1", f.m_sdfGi.m_indirectDiffuse_Y_SH[1]}, {"giC0", f.m_sdfGi.m_indirectDiffuse_CoCg[0]}, {"giC1", f.m_sdfGi.m_indirectDiffuse_CoCg[1]},
        {"giHistY0", f.m_sdfGi.m_indirectDiffuseHistory_Y_SH[0]}, {"giHistY1", f.m_sdfGi.m_indirectDiffuseHistory_Y_SH[1]},
        {"giHistC0", f.m_sdfGi.m_indirectDiffuseHistory_CoCg[0]}, {"giHistC1", f.m_sdfGi.m_indirectDiffuseHistory_CoCg[1]},
        {"giFullY", f.m_sdfGi.m_indirectLightingFullRes_Y_SH}, {"giFullC", f.m_sdfGi.m_indirectLightingFullRes_CoCg},
        {"froxelMaterial", f.m_volumetrics.m_volumeMaterialVolume}, {"froxelScatter", f.m_volumetrics.m_scatteringTransmittanceVolume},
        {"froxelHist0", f.m_volumetrics.m_volumetricLightingHistory[0]}, {"froxelHist1", f.m_volumetrics.m_volumetricLightingHistory[1]},
        {"froxelIntegration", f.m_volumetrics.m_volumetricIntegrationVolume}, {"taaHist0", f.m_taa.m_historyBuffers[0]}, {"taaHist1", f.m_taa.m_historyBuffers[1]},
        {"taaLum0", f.m_taa.m_sceneLuminance[0]}, {"taaLum1", f.m_taa.m_sceneLuminance[1]},
        {"bloomDown", f.m_bloom.m_lastDownscaleTexture}, {"bloomUp", f.m_bloom.m_lastUpscaleTexture}, {"output", f.backend.getSwapchainInputImage()},
    };
    auto it = m.find(name);
    if (it == m.end()) { fe->lastError = std::string("unknown image '") + name + "'"; return 1; }
    *out = it->second;
    return 0;
}
int PLAIN_FE(get_storage_buffer)(plain_frontend* fe, const char* name, plain_handle* out) {
    RenderFrontend& f = fe->fe;
    std::map<std::string, StorageBufferHandle> m = {
        {"histogram", f.m_histogramBuffer}, {"histogramPerTile", f.m_histogramPerTileBuffer}, {"light", f.m_lightBuffer}, {"sunShadowInfo", f.m_sunShadowInfoBuffer},
        {"sdfInstances", f.m_sdfGi.m_sdfInstanceBuffer}, {"sdfCulled", f.m_sdfGi.m_sdfCameraFrustumCulledInstances}, {"sdfTiles", f.m_sdfGi.m_sdfCameraCulledTiles},
        {"sdfWorldBBs", f.m_sdfGi.m_sdfInstanceWorldBBBuffer},
    };
    auto it = m.find(name);
    if (it == m.end()) { fe->lastError = std::string("unknown buffer '") + name + "'"; return 1; }
    *out = it->second.index;
    return 0;
}
int PLAIN_FE(get_global_shader_info)(plain_frontend* fe, void* out) { std::memcpy(out, &fe->fe.m_globalShaderInfo, sizeof(plain_global_shader_info)); return 0; }
int PLAIN_FE(get_resolve_weights)(plain_frontend* fe, float out[9]) { std::memcpy(out, fe->fe.m_taa.m_lastResolveWeights.data(), sizeof(float) * 9); return 0; }
// ---- host-side functions, for tests against the reference's own host code ----
static CameraExtrinsic extrinsicOf(const plain_camera_extrinsic* c) {
    CameraExtrinsic e;
    e.position = hm::Vec3(c->position[0], c->position[1], c->position[2]);
    e.forward = hm::Vec3(c->forward[0], c->forward[1], c->forward[2]);
    e.right = hm::Vec3(c->right[0], c->right[1], c->right[2]);
    e.up = hm::Vec3(c->up[0], c->up[1], c->up[2]);
    return e;
}
static CameraIntrinsic intrinsicOf(float fov, float aspect, float nearPlane, float farPlane) {
    CameraIntrinsic i;
    i.fov = fov; i.aspectRatio = aspect; i.near = nearPlane; i.far = farPlane;
    return i;
}
static void put3(float* out, hm::Vec3 v) { out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void PLAIN_FE(host_hammersley2d)(uint32_t index, float out[2]) { const hm::Vec2 h = hammersley2D(index); out[0] = h.x; out[1] = h.y; }
void PLAIN_FE(host_direction_to_vector)(const float deg[2], float out[3]) { hm::Vec2 d; d.x = deg[0]; d.y = deg[1]; put3(out, directionToVector(d)); }
uint32_t PLAIN_FE(host_mip_count_from_resolution)(uint32_t w, uint32_t h, uint32_t d) { return mipCountFromResolution(w, h, d); }
void PLAIN_FE(host_camera_matrices)(const plain_camera_extrinsic* c, float fov, float aspect, float nearPlane, float farPlane, float outView[16], float outProjection[16]) {
    const hm::Mat4 v = viewMatrixFromCameraExtrinsic(extrinsicOf(c)), p = projectionMatrixFromCameraIntrinsic(intrinsicOf(fov, aspect, nearPlane, farPlane));
    std::memcpy(outView, v.m, 64);
    std::memcpy(outProjection, p.m, 64);
}
void PLAIN_FE(host_view_frustum)(const plain_camera_extrinsic* c, float fov, float aspect, float nearPlane, float farPlane, float outPoints[24], float outNormals[18]) {
    const ViewFrustum f = computeViewFrustum(extrinsicOf(c), intrinsicOf(fov, aspect, nearPlane, farPlane));
    const hm::Vec3 p[8] = {f.l_l_n, f.l_l_f, f.l_u_n, f.l_u_f, f.r_l_n, f.r_l_f, f.r_u_n, f.r_u_f}, n[6] = {f.top, f.bot, f.right, f.left, f.near, f.far};
    for (int i = 0; i < 8; i++) put3(outPoints + 3 * i, p[i]);
    for (int i = 0; i < 6; i++) put3(outNormals + 3 * i, n[i]);
}
static ViewFrustum frustumIn(const float points[24], const float normals[18]) {
    ViewFrustum f;
    hm::Vec3* p[8] = {&f.l_l_n, &f.l_l_f, &f.l_u_n, &f.l_u_f, &f.r_l_n, &f.r_l_f, &f.r_u_n, &f.r_u_f};
    hm::Vec3* n[6] = {&f.top, &f.bot, &f.right, &f.left, &f.near, &f.far};
    for (int i = 0; i < 8; i++) *p[i] = hm::Vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    for (int i = 0; i < 6; i++) *n[i] = hm::Vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    return f;
}
void PLAIN_FE(host_orthogonal_frustum_fitted_to_camera)(const float points[24], const float normals[18], const float lightDirection[3], float outPoints[24], float outNormals[18]) {
    const ViewFrustum f = computeOrthogonalFrustumFittedToCamera(frustumIn(points, normals), hm::Vec3(lightDirection[0], lightDirection[1], lightDirection[2]));
    const hm::Vec3 p[8] = {f.l_l_n, f.l_l_f, f.l_u_n, f.l_u_f, f.r_l_n, f.r_l_f, f.r_u_n, f.r_u_f}, n[6] = {f.top, f.bot, f.right, f.left, f.near, f.far};
    for (int i = 0; i < 8; i++) put3(outPoints + 3 * i, p[i]);
    for (int i = 0; i < 6; i++) put3(outNormals + 3 * i, n[i]);
}
int PLAIN_FE(host_aabb_intersects_frustum)(const float points[24], const float normals[18], const float bbMin[3], const float bbMax[3]) {
    const ViewFrustum f = frustumIn(points, normals);
    hm::AABB bb;
    bb.min = hm::Vec3(bbMin[0], bbMin[1], bbMin[2]);
    bb.max = hm::Vec3(bbMax[0], bbMax[1], bbMax[2]);
    return isAxisAlignedBoundingBoxIntersectingViewFrustum(f, bb) ? 1 : 0;
}
void PLAIN_FE(host_pad_sdf_bounding_box)(const float bbMin[3], const float bbMax[3], float outMin[3], float outMax[3]) {
    hm::AABB bb;
    bb.min = hm::Vec3(bbMin[0], bbMin[1], bbMin[2]);
    bb.max = hm::Vec3(bbMax[0], bbMax[1], bbMax[2]);
    const hm::AABB r = padSDFBoundingBox(bb);
    put3(outMin, r.min); put3(outMax, r.max);
}
void PLAIN_FE(host_sdf_world_to_local)(const float model[16], const float bbOffset[3], float out[16]) {  // as SDFGI::updateSDFScene (Techniques.cpp)
    hm::Mat4 m;
    std::memcpy(m.m, model, 64);
    const hm::Mat4 r = hm::inverse(m * hm::translate(hm::Vec3(bbOffset[0], bbOffset[1], bbOffset[2])));
    std::memcpy(out, r.m, 64);
}
int PLAIN_FE(get_drawcall_counts)(plain_frontend* fe, uint32_t out[2]) {
    out[0] = fe->fe.m_currentMainPassDrawcallCount;
    out[1] = fe->fe.m_currentShadowPassDrawcallCount;
    return 0;
}
int PLAIN_FE(set_exposure)(plain_frontend* fe, float previousFrameExposure) {
    FE_TRY(fe, {
        plain_light_buffer lb{};
        lb.previousFrameExposure = previousFrameExposure;
        lb.sunStrengthExposed = fe->fe.m_globalShaderInfo.sunStrength * previousFrameExposure;
        lb.sunColor[0] = lb.sunColor[1] = lb.sunColor[2] = 1.f;
        fe->fe.backend.setStorageBufferData(fe->fe.m_lightBuffer, &lb, sizeof(lb));
    });
}

}  // extern "C"
