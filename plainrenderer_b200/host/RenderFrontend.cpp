// RenderFrontend.cpp - see RenderFrontend.h. Follows Plain/src/Runtime/Rendering/RenderFrontend.cpp: setup :156-192,
// pass list :313-406, camera/jitter :423-454, histogram :707-754, exposure :776-790, depth pyramid :804-838 and
// :1770-1827, light matrices :840-871, depth downscale :873-892, forward shading :894-929 (recast), tonemapping
// :931-945, global shader info :1158-1184, images/buffers :1186-1516.
#include "RenderFrontend.h"
#include <cstring>

static const uint32_t shadowMapRes = 2048;
static const uint32_t brdfLutRes = 512;
static const uint32_t nHistogramBins = 128;
static const int maxSunShadowCascadeCount = 4;
static const uint32_t histogramTileSizeX = 32, histogramTileSizeY = 32;
static const uint32_t noiseTextureCount = 4, noiseTextureWidth = 32, noiseTextureHeight = 32;
static const float histogramMinValue = 0.001f, histogramMaxValue = 200000.f;  // createHistogramSettings :1063-1073

static uint32_t ceilDiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static void setRows(RenderBackend& b, ComputePassExecution& e, uint32_t divisor, uint32_t rows, uint32_t extend = 0) { b.band(divisor, rows, extend, &e.rowBegin, &e.rowEnd); }
static ExchangeRequest gatherRows(const char* name, std::initializer_list<ImageHandle> images, uint32_t mip, uint32_t divisor) {
    ExchangeRequest x;
    x.kind = PLAIN_EXCHANGE_ALLGATHER_ROWS;
    x.name = name;
    for (ImageHandle h : images) { x.images.push_back(h); x.mips.push_back(mip); x.divisors.push_back(divisor); }
    return x;
}

// ---- small host utilities (MathUtils.cpp, sdfUtilities.cpp, ViewFrustum.cpp) ----
static float radicalInverseBase2(uint32_t in) {
    uint32_t out = (in << 16) | (in >> 16);
    out = ((out & 0x00ff00ffu) << 8) | ((out & 0xff00ff00u) >> 8);
    out = ((out & 0x0f0f0f0fu) << 4) | ((out & 0xf0f0f0f0u) >> 4);
    out = ((out & 0x33333333u) << 2) | ((out & 0xccccccccu) >> 2);
    out = ((out & 0x55555555u) << 1) | ((out & 0xaaaaaaaau) >> 1);
    return (float)out * 2.3283064365386963e-10f;
}
static float radicalInverseBase3(uint32_t in) {
    const float inverseBase = 1.f / 3.f;
    uint32_t reversedDigits = 0, current = in;
    float inverseBasePowerN = 1.f;
    while (current) {
        uint32_t next = current / 3;
        reversedDigits = reversedDigits * 3 + (current - next * 3);
        inverseBasePowerN *= inverseBase;
        current = next;
    }
    return (float)reversedDigits * inverseBasePowerN;
}
hm::Vec2 hammersley2D(uint32_t index) { hm::Vec2 r; r.x = radicalInverseBase2(index); r.y = radicalInverseBase3(index); return r; }
hm::Vec3 directionToVector(hm::Vec2 d) {
    float theta = hm::radians(d.y), phi = hm::radians(d.x);
    return hm::Vec3(dm::sin(theta) * dm::cos(phi), -dm::cos(theta), dm::sin(theta) * dm::sin(phi));
}
uint32_t mipCountFromResolution(uint32_t w, uint32_t h, uint32_t d) {
    uint32_t m = w > h ? w : h;
    if (d > m) m = d;
    uint32_t n = 1;
    while (m > 1) { m >>= 1; n++; }
    return n;
}
hm::AABB padSDFBoundingBox(const hm::AABB& bb) {
    hm::Vec3 padding = (bb.max - bb.min) * 0.075f;
    padding = hm::vmax(padding, hm::Vec3(0.5f));
    hm::AABB p;
    p.min = bb.min - padding;
    p.max = bb.max + padding;
    return p;
}
ViewFrustum computeViewFrustum(const CameraExtrinsic& e, const CameraIntrinsic& in) {
    using namespace hm;
    Vec3 nearC = e.position + e.forward * in.near, farC = e.position + e.forward * in.far;
    float tanFoV = dm::tan(radians(in.fov) * 0.5f);
    float hn = tanFoV * in.near, hf = tanFoV * in.far, wn = hn * in.aspectRatio, wf = hf * in.aspectRatio;
    ViewFrustum f;
    f.r_u_f = farC + e.up * hf + e.right * wf;   f.l_u_f = farC + e.up * hf - e.right * wf;
    f.r_l_f = farC - e.up * hf + e.right * wf;   f.l_l_f = farC - e.up * hf - e.right * wf;
    f.r_u_n = nearC + e.up * hn + e.right * wn;  f.l_u_n = nearC + e.up * hn - e.right * wn;
    f.r_l_n = nearC - e.up * hn + e.right * wn;  f.l_l_n = nearC - e.up * hn - e.right * wn;
    f.top = normalize(cross(f.r_u_f - f.r_u_n, f.r_u_n - f.l_u_n));
    f.bot = normalize(cross(f.r_l_n - f.l_l_n, f.r_l_f - f.r_l_n));
    f.right = normalize(cross(f.r_u_n - f.r_l_n, f.r_l_f - f.r_l_n));
    f.left = normalize(cross(f.l_l_f - f.l_l_n, f.l_u_n - f.l_l_n));
    f.near = normalize(cross(f.r_u_n - f.r_l_n, f.r_l_n - f.l_l_n));
    f.far = normalize(cross(f.r_l_f - f.l_l_f, f.r_u_f - f.r_l_f));
    return f;
}
static void computeViewFrustumNormals(ViewFrustum& f) {  // ViewFrustum.cpp:39-51
    using namespace hm;
    f.top = normalize(cross(f.r_u_f - f.r_u_n, f.r_u_n - f.l_u_n));
    f.bot = normalize(cross(f.r_l_n - f.l_l_n, f.r_l_f - f.r_l_n));
    f.right = normalize(cross(f.r_u_n - f.r_l_n, f.r_l_f - f.r_l_n));
    f.left = normalize(cross(f.l_l_f - f.l_l_n, f.l_u_n - f.l_l_n));
    f.near = normalize(cross(f.r_u_n - f.r_l_n, f.r_l_n - f.l_l_n));
    f.far = normalize(cross(f.r_l_f - f.l_l_f, f.r_u_f - f.r_l_f));
}
// the orthographic box, in the light's view space, around the eight corners of the camera frustum (ViewFrustum.cpp:231-271)
ViewFrustum computeOrthogonalFrustumFittedToCamera(const ViewFrustum& c, hm::Vec3 lightDirection) {
    using namespace hm;
    const Vec3 up = dm::abs_(lightDirection.y) < 0.999f ? Vec3(0.f, -1.f, 0.f) : Vec3(0.f, 0.f, -1.f);
    // glm::lookAt(eye = -lightDirection, center = 0, up), right-handed
    const Vec3 eye = -lightDirection;
    const Vec3 fw = normalize(Vec3(0.f) - eye), s = normalize(cross(fw, up)), u = cross(s, fw);
    Mat4 V = Mat4::identity();
    V.at(0, 0) = s.x; V.at(1, 0) = s.y; V.at(2, 0) = s.z;
    V.at(0, 1) = u.x; V.at(1, 1) = u.y; V.at(2, 1) = u.z;
    V.at(0, 2) = -fw.x; V.at(1, 2) = -fw.y; V.at(2, 2) = -fw.z;
    V.at(3, 0) = -dot(s, eye); V.at(3, 1) = -dot(u, eye); V.at(3, 2) = dot(fw, eye);
    const float inf = dm::inff_();
    Vec3 maxP(-inf), minP(inf);
    const Vec3 corners[8] = {c.l_l_f, c.l_l_n, c.r_l_f, c.r_l_n, c.l_u_f, c.l_u_n, c.r_u_f, c.r_u_n};
    for (const Vec3& p : corners) {
        const Vec4 t = V * Vec4(p, 1.f);
        minP = vmin(minP, Vec3(t.x, t.y, t.z));
        maxP = vmax(maxP, Vec3(t.x, t.y, t.z));
    }
    const Vec3 scale(2.f / (maxP.x - minP.x), 2.f / (maxP.y - minP.y), 2.f / (maxP.z - minP.z));
    const Vec3 offset = (maxP + minP) * -0.5f * scale;
    Mat4 clip = Mat4::identity();
    clip.at(0, 0) = scale.x; clip.at(1, 1) = scale.y; clip.at(2, 2) = scale.z;
    clip.at(3, 0) = offset.x; clip.at(3, 1) = offset.y; clip.at(3, 2) = offset.z;
    const Mat4 clipToWorld = inverse(clip * V);
    auto corner = [&](float x, float y, float z) { const Vec4 t = clipToWorld * Vec4(x, y, z, 1.f); return Vec3(t.x, t.y, t.z); };
    ViewFrustum r;
    r.l_l_n = corner(-1, -1, -1); r.r_l_n = corner(1, -1, -1); r.l_u_n = corner(-1, 1, -1); r.r_u_n = corner(1, 1, -1);
    r.l_l_f = corner(-1, -1, 1); r.r_l_f = corner(1, -1, 1); r.l_u_f = corner(-1, 1, 1); r.r_u_f = corner(1, 1, 1);
    computeViewFrustumNormals(r);
    return r;
}
hm::Mat4 viewMatrixFromCameraExtrinsic(const CameraExtrinsic& e) {  // Camera.cpp:4-12
    hm::Mat4 v = hm::Mat4::identity();
    v.at(0, 0) = e.right.x; v.at(0, 1) = e.right.y; v.at(0, 2) = e.right.z;
    v.at(1, 0) = e.up.x; v.at(1, 1) = e.up.y; v.at(1, 2) = e.up.z;
    v.at(2, 0) = -e.forward.x; v.at(2, 1) = -e.forward.y; v.at(2, 2) = -e.forward.z;
    v = hm::transpose(v);
    return v * hm::translate(-e.position);
}
hm::Mat4 projectionMatrixFromCameraIntrinsic(const CameraIntrinsic& in) {  // Camera.cpp:14-27: y flip, reverse z
    hm::Mat4 p = hm::perspective(hm::radians(in.fov), in.aspectRatio, in.near, in.far);
    hm::Mat4 c = hm::Mat4::identity();
    c.at(1, 1) = -1.f; c.at(2, 2) = -0.5f; c.at(3, 2) = 0.5f;
    return c * p;
}

// deterministic fixture noise (the reference seeds its blue/Perlin noise with C rand())
static uint32_t pcg(uint32_t& s) {
    s = s * 747796405u + 2891336453u;
    uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}

void RenderFrontend::setup(int device, uint32_t width, uint32_t height, uint32_t noiseSeed) {
    m_screenWidth = width;
    m_screenHeight = height;
    m_cameraIntrinsic.aspectRatio = (float)width / (float)height;
    m_sunDirection.x = 0.f; m_sunDirection.y = 0.f;
    std::memset(&m_globalShaderInfo, 0, sizeof(m_globalShaderInfo));
    // GlobalShaderInfo defaults, ResourceDescriptions.h:174-203
    m_globalShaderInfo.sunDirection[1] = -1.f;
    m_globalShaderInfo.cameraRight[0] = 1.f; m_globalShaderInfo.cameraUp[1] = -1.f;
    m_globalShaderInfo.cameraForward[2] = -1.f; m_globalShaderInfo.cameraForwardPrevious[2] = -1.f;
    m_globalShaderInfo.cameraTanFovHalf = 1.f; m_globalShaderInfo.cameraAspectRatio = 1.f;
    m_globalShaderInfo.nearPlane = 0.1f; m_globalShaderInfo.farPlane = 100.f;
    m_globalShaderInfo.sunStrength = 128000.f; m_globalShaderInfo.exposureOffset = 1.f;
    m_globalShaderInfo.exposureAdaptionSpeedEvPerSec = 2.f; m_globalShaderInfo.deltaTime = 0.016f;
    // AtmosphereSettings defaults, Sky.h:6-15
    plain_atmosphere_settings& a = m_atmosphereSettings;
    a.scatteringRayleighGround[0] = 0.0058f; a.scatteringRayleighGround[1] = 0.0135f; a.scatteringRayleighGround[2] = 0.0331f;
    a.earthRadius = 6371.f;
    for (int i = 0; i < 3; i++) a.extinctionRayleighGround[i] = a.scatteringRayleighGround[i];
    a.atmosphereHeight = 100.f;
    a.ozoneExtinction[0] = 0.000650f; a.ozoneExtinction[1] = 0.001881f; a.ozoneExtinction[2] = 0.000085f;
    a.scatteringMieGround = 0.006f;
    a.extinctionMieGround = 1.11f * a.scatteringMieGround;
    a.mieScatteringExponent = 0.76f;

    backend.setup(device, width, height);
    if (m_shardCount > 1) {
        if ((height + PLAIN_SHARD_ROW_UNIT - 1) / PLAIN_SHARD_ROW_UNIT < m_shardCount) throw std::runtime_error("row sharding: fewer 32-row units than ranks");
        if (width % 16 != 0 || height % 16 != 0) throw std::runtime_error("row sharding needs a resolution that is a multiple of 16 (four HiZ levels reduced from a rank's own rows)");
        if (m_taaSettings.useSeparateSupersampling || m_sdfDebugSettings.visualisationMode != SDFVisualisationMode::None)
            throw std::runtime_error("row sharding: the separate temporal supersampling pass and the SDF debug visualisation are single-GPU only");
        backend.shard.rank = m_shardRank; backend.shard.count = m_shardCount; backend.shard.fullHeight = height;
        shardBandRows(height, m_shardCount, m_shardRank, &backend.shard.y0, &backend.shard.y1);
    }
    // the 8 global samplers, RenderFrontend.cpp:1300-1397 (binding order global.inc:35-42)
    const plain_sampler_desc descs[8] = {
        {PLAIN_SAMPLER_LINEAR, PLAIN_WRAP_REPEAT, 1, 8.f, PLAIN_BORDER_WHITE, 20},   // anisotropicRepeat
        {PLAIN_SAMPLER_NEAREST, PLAIN_WRAP_COLOR, 0, 0.f, PLAIN_BORDER_BLACK, 20},   // nearestBlackBorder
        {PLAIN_SAMPLER_LINEAR, PLAIN_WRAP_REPEAT, 0, 0.f, PLAIN_BORDER_WHITE, 20},   // linearRepeat
        {PLAIN_SAMPLER_LINEAR, PLAIN_WRAP_CLAMP, 0, 0.f, PLAIN_BORDER_WHITE, 20},    // linearClamp
        {PLAIN_SAMPLER_NEAREST, PLAIN_WRAP_CLAMP, 0, 8.f, PLAIN_BORDER_BLACK, 20},   // nearestClamp
        {PLAIN_SAMPLER_LINEAR, PLAIN_WRAP_COLOR, 0, 8.f, PLAIN_BORDER_WHITE, 20},    // linearWhiteBorder
        {PLAIN_SAMPLER_NEAREST, PLAIN_WRAP_REPEAT, 0, 8.f, PLAIN_BORDER_BLACK, 20},  // nearestRepeat
        {PLAIN_SAMPLER_NEAREST, PLAIN_WRAP_COLOR, 0, 0.f, PLAIN_BORDER_WHITE, 20},   // nearestWhiteBorder
    };
    for (int i = 0; i < 8; i++) m_samplers[i] = backend.createSampler(descs[i]);
    initImages(noiseSeed);
    initBuffers();
    initRenderpasses();
    m_sky.init(backend);
    m_bloom.init(backend);
    m_volumetrics.init(backend, (int)width, (int)height, noiseSeed ^ 0x9e3779b9u);
    m_taa.init(backend, (int)width, (int)height, m_taaSettings);
    m_sdfGi.init(backend, (int)width, (int)height, m_sdfTraceSettings, m_sdfDebugSettings, m_shadingConfig.sunShadowCascadeCount - 1);

    RenderPassResources globalResources;  // setupGlobalShaderInfoResources :295-311
    globalResources.uniformBuffers = {UniformBufferResource(m_globalUniformBuffer, 0)};
    for (uint32_t i = 0; i < 8; i++) globalResources.samplers.push_back(SamplerResource(m_samplers[i], i + 1));
    backend.setGlobalDescriptorSetResources(globalResources);

    backend.newFrame();
    computeBRDFLut();
    backend.prepareForDrawcallRecording();
    backend.renderFrame(false);
}


void RenderFrontend::shutdown() { backend.shutdown(); }

void RenderFrontend::initImages(uint32_t noiseSeed) {
    const uint32_t w = m_screenWidth, h = m_screenHeight;
    const uint32_t SS = PLAIN_USAGE_SAMPLED | PLAIN_USAGE_STORAGE;
    for (int i = 0; i < 2; i++) m_postProcessBuffers[i] = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_R11G11B10_UFLOAT, SS), nullptr, 0);
    for (int i = 0; i < maxSunShadowCascadeCount; i++)
        m_shadowMaps.push_back(backend.createImage(imageDesc2D(shadowMapRes, shadowMapRes, PLAIN_FORMAT_DEPTH16, PLAIN_USAGE_ATTACHMENT | PLAIN_USAGE_SAMPLED), nullptr, 0));
    m_brdfLut = backend.createImage(imageDesc2D(brdfLutRes, brdfLutRes, PLAIN_FORMAT_RGBA16_SFLOAT, SS), nullptr, 0);
    m_minMaxDepthPyramid = backend.createImage(imageDesc2D(w / 2, h / 2, PLAIN_FORMAT_RG32_SFLOAT, SS, PLAIN_MIPS_FULL_CHAIN), nullptr, 0);
    uint32_t s = noiseSeed;
    for (uint32_t i = 0; i < noiseTextureCount; i++) {
        std::vector<uint8_t> noise(noiseTextureWidth * noiseTextureHeight * 2);
        for (auto& v : noise) v = (uint8_t)(pcg(s) & 0xff);
        m_noiseTextures.push_back(backend.createImage(imageDesc2D(noiseTextureWidth, noiseTextureHeight, PLAIN_FORMAT_RG8, PLAIN_USAGE_SAMPLED), noise.data(), noise.size()));
        m_globalShaderInfo.noiseTextureIndices[i] = (int32_t)backend.getImageGlobalTextureArrayIndex(m_noiseTextures.back());
    }
    for (int i = 0; i < 2; i++) m_worldSpaceNormalImages[i] = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_RGBA8, PLAIN_USAGE_ATTACHMENT | PLAIN_USAGE_SAMPLED), nullptr, 0);
    m_depthHalfRes = backend.createImage(imageDesc2D(w / 2, h / 2, PLAIN_FORMAT_R16_SFLOAT, SS), nullptr, 0);
    // packed G-buffer: the post-raster inputs of triangle.frag (new in this build)
    for (int i = 0; i < 2; i++) m_gbuffers[i] = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_RGBA32_UINT, PLAIN_USAGE_ATTACHMENT | PLAIN_USAGE_SAMPLED), nullptr, 0);
    for (int i = 0; i < 3; i++) m_motionBuffers[i] = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_RG16_SNORM, PLAIN_USAGE_ATTACHMENT | PLAIN_USAGE_SAMPLED), nullptr, 0);
    for (int i = 0; i < 2; i++) {  // initRenderTargets :1399-1447
        m_frameRenderTargets[i].motionBuffer = m_motionBuffers[i];
        m_frameRenderTargets[i].colorBuffer = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_R11G11B10_UFLOAT, PLAIN_USAGE_ATTACHMENT | SS), nullptr, 0);
        m_frameRenderTargets[i].depthBuffer = backend.createImage(imageDesc2D(w, h, PLAIN_FORMAT_DEPTH32, PLAIN_USAGE_ATTACHMENT | PLAIN_USAGE_SAMPLED), nullptr, 0);
    }
}

void RenderFrontend::initBuffers() {
    m_histogramBuffer = backend.createStorageBuffer(nHistogramBins * sizeof(uint32_t));
    plain_light_buffer initialLight{};
    m_lightBuffer = backend.createStorageBuffer(sizeof(plain_light_buffer), &initialLight);
    // the reference sizes this for 1920x1080 (:1069-1070 FIXME); sized by the actual resolution here
    const size_t tiles = (size_t)ceilDiv(m_screenWidth, histogramTileSizeX) * ceilDiv(m_screenHeight, histogramTileSizeY);
    m_histogramPerTileBuffer = backend.createStorageBuffer(tiles * nHistogramBins * sizeof(uint32_t));
    uint32_t zero = 0;
    m_depthPyramidSyncBuffer = backend.createStorageBuffer(sizeof(uint32_t), &zero);
    m_sunShadowInfoBuffer = backend.createStorageBuffer(sizeof(plain_shadow_cascade_info));
    m_globalUniformBuffer = backend.createUniformBuffer(sizeof(plain_global_shader_info));
    const size_t maxObjectCountMainScene = 1200;  // SceneConfig.h
    m_mainPassTransformsBuffer = backend.createStorageBuffer(maxObjectCountMainScene * 3 * sizeof(hm::Mat4));  // MainPassMatrices {model, mvp, mvpPrevious}
    m_shadowPassTransformsBuffer = backend.createStorageBuffer(maxObjectCountMainScene * sizeof(hm::Mat4));
}

template <typename T> static SpecialisationConstant specConst(uint32_t location, const T& v) { return SpecialisationConstant{location, dataToCharArray(&v, sizeof(T))}; }

ShaderDescription RenderFrontend::createDepthPyramidShaderDescription(uint32_t* outThreadgroupCount) const {
    ShaderDescription desc;
    desc.srcPathRelative = "depthHiZPyramid.comp";
    const uint32_t width = m_screenWidth / 2, height = m_screenHeight / 2;
    const uint32_t depthMipCount = mipCountFromResolution(width, height, 1);
    uint32_t dc[2];
    computeSinglePassMipChainDispatchCount(width, height, depthMipCount, depthMipCount > 11 ? depthMipCount : 11, dc);
    *outThreadgroupCount = dc[0] * dc[1];
    desc.specialisationConstants = {specConst(0, depthMipCount), specConst(1, m_screenWidth), specConst(2, m_screenHeight), specConst(3, *outThreadgroupCount)};
    return desc;
}

void RenderFrontend::computeSinglePassMipChainDispatchCount(uint32_t w, uint32_t h, uint32_t mipCount, uint32_t maxMipCount, uint32_t out[2]) const {
    const uint32_t unusedMips = maxMipCount > mipCount ? maxMipCount - mipCount : 0;
    if (unusedMips >= 6) { out[0] = 1; out[1] = 1; return; }
    const uint32_t localThreadGroupExtent = 32u >> unusedMips;
    out[0] = ceilDiv(w, localThreadGroupExtent);
    out[1] = ceilDiv(h, localThreadGroupExtent);
}

void RenderFrontend::initRenderpasses() {
    {   // depth prepass :1716-1735, shadow cascades :1563-1590, and the raster half of the main pass :1537-1561 (depth test EQUAL)
        GraphicPassDescription d;
        d.attachments = {Attachment{PLAIN_FORMAT_RG16_SNORM, PLAIN_LOAD_OP_CLEAR}, Attachment{PLAIN_FORMAT_RGBA8, PLAIN_LOAD_OP_CLEAR}, Attachment{PLAIN_FORMAT_DEPTH32, PLAIN_LOAD_OP_CLEAR}};
        d.depthTest.function = PLAIN_DEPTH_GREATER_EQUAL; d.depthTest.write = true;
        d.name = "Depth prepass";
        d.rasterization.cullMode = PLAIN_CULL_BACK;
        d.shaderDescriptions.vertex.srcPathRelative = "depthPrepass.vert";
        d.shaderDescriptions.fragment.srcPathRelative = "depthPrepass.frag";
        m_depthPrePass = backend.createGraphicPass(d);
        GraphicPassDescription m;
        m.attachments = {Attachment{PLAIN_FORMAT_RGBA32_UINT, PLAIN_LOAD_OP_CLEAR}, Attachment{PLAIN_FORMAT_DEPTH32, PLAIN_LOAD_OP_LOAD}};
        m.depthTest.function = PLAIN_DEPTH_EQUAL; m.depthTest.write = true;
        m.name = "G-buffer fill";
        m.rasterization.cullMode = PLAIN_CULL_BACK;
        m.shaderDescriptions.vertex.srcPathRelative = "triangle.vert";
        m.shaderDescriptions.fragment.srcPathRelative = "gbufferFill.frag";
        m_gbufferFillPass = backend.createGraphicPass(m);
        for (uint32_t cascade = 0; cascade < (uint32_t)maxSunShadowCascadeCount; cascade++) {
            GraphicPassDescription s;
            s.name = "Shadow map cascade " + std::to_string(cascade);
            s.attachments = {Attachment{PLAIN_FORMAT_DEPTH16, PLAIN_LOAD_OP_CLEAR}};
            s.shaderDescriptions.vertex.srcPathRelative = "sunShadow.vert";
            s.shaderDescriptions.fragment.srcPathRelative = "sunShadow.frag";
            s.shaderDescriptions.vertex.specialisationConstants = {specConst(0, cascade)};
            s.depthTest.function = PLAIN_DEPTH_GREATER_EQUAL; s.depthTest.write = true;
            s.rasterization.cullMode = PLAIN_CULL_FRONT;
            s.rasterization.clampDepth = true;
            m_shadowPasses[cascade] = backend.createGraphicPass(s);
        }
        // createDefaultTextures :86-150 (the normal default is two-channel RG8 {128, 128} there; RGBA8 here, the material textures
        // of the G-buffer fill are RGBA8)
        const uint8_t diffuse[4] = {255, 255, 255, 255}, specular[4] = {0, 128, 255, 0}, normal[4] = {128, 128, 255, 255};
        m_defaultTextures.diffuse = backend.createImage(imageDesc2D(1, 1, PLAIN_FORMAT_RGBA8, PLAIN_USAGE_SAMPLED), diffuse, 4);
        m_defaultTextures.specular = backend.createImage(imageDesc2D(1, 1, PLAIN_FORMAT_RGBA8, PLAIN_USAGE_SAMPLED), specular, 4);
        m_defaultTextures.normal = backend.createImage(imageDesc2D(1, 1, PLAIN_FORMAT_RGBA8, PLAIN_USAGE_SAMPLED), normal, 4);
    }
    {   // deferred recast of the "Forward shading" graphic pass; constants as createForwardPassShaderDescription :1099-1135
        ComputePassDescription d;
        d.name = "Forward shading";
        d.shaderDescription.srcPathRelative = "gbufferShading.comp";
        const int diffuse = (int)m_shadingConfig.diffuseBRDF, multi = (int)m_shadingConfig.directMultiscatter, tech = (int)m_shadingConfig.indirectLightingTech;
        const uint32_t geoAA = m_shadingConfig.useGeometryAA ? 1u : 0u;
        const uint32_t cascades = (uint32_t)m_shadingConfig.sunShadowCascadeCount;
        d.shaderDescription.specialisationConstants = {specConst(0, diffuse), specConst(1, multi), specConst(2, geoAA), specConst(3, tech), specConst(4, cascades)};
        m_shadingPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "BRDF Lut creation";
        d.shaderDescription.srcPathRelative = "brdfLut.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, (int)m_shadingConfig.diffuseBRDF)};
        m_brdfLutPass = backend.createComputePass(d);
    }
    const uint32_t maxTileCount = ceilDiv(m_screenWidth, histogramTileSizeX) * ceilDiv(m_screenHeight, histogramTileSizeY);
    {
        ComputePassDescription d;
        d.name = "Histogram per tile";
        d.shaderDescription.srcPathRelative = "histogramPerTile.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, nHistogramBins), specConst(1, histogramMinValue), specConst(2, histogramMaxValue), specConst(3, (int)maxTileCount)};
        m_histogramPerTilePass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Histogram reset";
        d.shaderDescription.srcPathRelative = "histogramReset.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, nHistogramBins)};
        m_histogramResetPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Histogram combine tiles";
        d.shaderDescription.srcPathRelative = "histogramCombineTiles.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, nHistogramBins), specConst(1, (int)maxTileCount)};
        m_histogramCombinePass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Pre-expose lights";
        d.shaderDescription.srcPathRelative = "preExposeLights.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, (int)nHistogramBins), specConst(1, histogramMinValue), specConst(2, histogramMaxValue)};
        m_preExposeLightsPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Depth min/max pyramid creation";
        uint32_t threadgroupCount = 0;
        d.shaderDescription = createDepthPyramidShaderDescription(&threadgroupCount);
        m_depthPyramidPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Compute light matrix";
        d.shaderDescription.srcPathRelative = "lightMatrix.comp";
        d.shaderDescription.specialisationConstants = {specConst(0, (uint32_t)m_shadingConfig.sunShadowCascadeCount)};
        m_lightMatrixPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Tonemapping";
        d.shaderDescription.srcPathRelative = "tonemapping.comp";
        m_tonemappingPass = backend.createComputePass(d);
    }
    {
        ComputePassDescription d;
        d.name = "Depth downscale";
        d.shaderDescription.srcPathRelative = "depthDownscale.comp";
        m_depthDownscalePass = backend.createComputePass(d);
    }
}

void RenderFrontend::markNewFrame(float time, float deltaTime) {
    m_time = time;
    m_deltaTime = deltaTime;
    m_frameIndex.markNewFrame();
}

void RenderFrontend::prepareNewFrame() {
    m_currentMainPassDrawcallCount = 0;    // :275-277
    m_currentShadowPassDrawcallCount = 0;
    backend.newFrame();
    backend.beginRecording();  // executions are kept on the host and sent segment by segment in renderFrameSegment
    prepareRenderpasses();
    backend.endRecording();
    backend.prepareForDrawcallRecording();
}

// the ordered pass list, RenderFrontend.cpp:313-406
void RenderFrontend::prepareRenderpasses() {
    const FrameRenderTargets previousRenderTarget = m_frameRenderTargets[m_sceneRenderTargetIndex];
    m_sceneRenderTargetIndex = (m_sceneRenderTargetIndex + 1) % 2;
    // the motion buffer rotates through three images: the one frame N-1 wrote is still read by frame N (velocityLastFrame of
    // the GI temporal filter), so with two the upload of frame N+1's motion vectors would have to wait for frame N to end
    m_motionBufferIndex = (m_motionBufferIndex + 1) % 3;
    m_frameRenderTargets[m_sceneRenderTargetIndex].motionBuffer = m_motionBuffers[m_motionBufferIndex];
    const FrameRenderTargets currentRenderTarget = m_frameRenderTargets[m_sceneRenderTargetIndex];

    auto fillOutSdfGiDependencies = [&]() {  // :1075-1092
        SDFTraceDependencies dep;
        dep.currentFrame = currentRenderTarget;
        dep.previousFrame = previousRenderTarget;
        dep.cameraFrustum = m_cameraFrustum;
        dep.depthHalfRes = m_depthHalfRes;
        dep.worldSpaceNormals = worldSpaceNormalImage();
        dep.skyLut = m_sky.m_skyLut;
        dep.shadowMap = m_shadowMaps[m_shadingConfig.sunShadowCascadeCount - 1];
        dep.lightBuffer = m_lightBuffer;
        dep.sunShadowInfoBuffer = m_sunShadowInfoBuffer;
        dep.depthMinMaxPyramid = m_minMaxDepthPyramid;
        return dep;
    };
    if (m_sdfDebugSettings.visualisationMode != SDFVisualisationMode::None) {  // :321-340, primary rays through the SDF scene instead of the frame
        if (m_rasterInputs) renderDepthPrepass(currentRenderTarget.depthBuffer, worldSpaceNormalImage(), currentRenderTarget.motionBuffer);  // else: uploaded
        computeDepthPyramid(currentRenderTarget.depthBuffer);
        computeColorBufferHistogram(m_postProcessBuffers[0]);
        m_sky.updateTransmissionLut(backend);
        computeExposure();
        m_sky.updateSkyLut(backend, m_lightBuffer, m_atmosphereSettings);
        computeSunLightMatrices();
        if (m_rasterInputs) renderSunShadowCascades();  // else: uploaded
        m_sdfGi.renderSDFVisualization(backend, m_postProcessBuffers[0], fillOutSdfGiDependencies(), m_sdfDebugSettings, m_sdfTraceSettings);
        computeTonemapping(m_postProcessBuffers[0]);
        return;
    }

    // row-sharded: the half-resolution depth is produced next to the pyramid's local levels and gathered in the same exchange
    const bool downscaleWithPyramid = backend.shard.active() && m_shadingConfig.indirectLightingTech == IndirectLightingTech::SDFTrace && m_sdfTraceSettings.halfResTrace;
    // row-sharded with uploaded inputs: the rank's share of the histogram, of the pyramid and of the half-resolution depth need nothing
    // but last frame's colour and this frame's depth, so they are recorded first and their two exchanges (row all-gather, counter
    // all-reduce) share ONE barrier; everything up to the SDF trace - exposure, the three sky LUTs, the small pyramid levels, light
    // matrices, the froxel chain, instance culling - then forms one submission whose independent chains run side by side
    const bool exchangesFirst = backend.shard.active() && !m_rasterInputs;
    if (exchangesFirst) {
        computeColorBufferHistogram(previousRenderTarget.colorBuffer, false);
        computeDepthPyramid(currentRenderTarget.depthBuffer, downscaleWithPyramid ? &currentRenderTarget : nullptr, true);
        backend.addExchange(m_pendingHistogramExchange);
        m_sky.updateTransmissionLut(backend);
        computeExposure();
        m_sky.updateSkyLut(backend, m_lightBuffer, m_atmosphereSettings);
        backend.setComputePassExecution(m_pendingPyramidRest);
    } else {
        computeColorBufferHistogram(previousRenderTarget.colorBuffer);
        m_sky.updateTransmissionLut(backend);
        computeExposure();
        m_sky.updateSkyLut(backend, m_lightBuffer, m_atmosphereSettings);
        // depth / motion / normal / G-buffer / shadow maps of this frame: rasterised from the meshes (m_rasterInputs) or uploaded
        if (m_rasterInputs) renderDepthPrepass(currentRenderTarget.depthBuffer, worldSpaceNormalImage(), currentRenderTarget.motionBuffer);
        computeDepthPyramid(currentRenderTarget.depthBuffer, downscaleWithPyramid ? &currentRenderTarget : nullptr);
    }
    computeSunLightMatrices();
    if (m_rasterInputs) renderSunShadowCascades();
    Volumetrics::Dependencies vd;
    vd.lightBuffer = m_lightBuffer;
    vd.shadowMap = m_shadowMaps[m_shadingConfig.sunShadowCascadeCount - 1];
    vd.sunShadowInfoBuffer = m_sunShadowInfoBuffer;
    // row-sharded: a submission ends at every exchange, so the froxel chain (which depends on the light matrices and the exposure only)
    // is recorded next to the SDF trace instead of behind the GI chain - the backend's scheduler runs it on a side stream beside the
    // trace, as it does in the unsharded frame's single submission. Same passes, same inputs, same bits.
    const bool volumetricsFirst = backend.shard.active();
    if (volumetricsFirst) m_volumetrics.computeVolumetricLighting(backend, m_volumetricsSettings, m_windSettings, vd, m_frameIndex, m_deltaTime);
    if (m_shadingConfig.indirectLightingTech == IndirectLightingTech::SDFTrace) {
        if (m_sdfTraceSettings.halfResTrace && !downscaleWithPyramid) downscaleDepth(currentRenderTarget);
        m_sdfGi.computeIndirectLighting(backend, fillOutSdfGiDependencies(), m_sdfTraceSettings, m_frameIndex);
    }
    if (!volumetricsFirst) m_volumetrics.computeVolumetricLighting(backend, m_volumetricsSettings, m_windSettings, vd, m_frameIndex, m_deltaTime);

    if (m_rasterInputs) fillGBuffer(gbuffer(), currentRenderTarget.depthBuffer);
    shadeGBuffer(currentRenderTarget.colorBuffer);  // renderForwardShading + m_sky.renderSky

    ImageHandle currentSrc = currentRenderTarget.colorBuffer;
    if (m_taaSettings.enabled) {
        if (m_taaSettings.useSeparateSupersampling) {
            m_taa.computeTemporalSuperSampling(backend, currentRenderTarget, previousRenderTarget, m_postProcessBuffers[0], m_frameIndex);
            currentSrc = m_postProcessBuffers[0];
        }
        m_taa.computeTemporalFilter(backend, currentSrc, currentRenderTarget, m_postProcessBuffers[1], m_frameIndex);
        currentSrc = m_postProcessBuffers[1];
    }
    if (m_bloomSettings.enabled) m_bloom.computeBloom(backend, currentSrc, m_bloomSettings);
    computeTonemapping(currentSrc);
}

void RenderFrontend::setCameraExtrinsic(const CameraExtrinsic& extrinsic) {
    std::memcpy(m_globalShaderInfo.previousFrameCameraJitter, m_globalShaderInfo.currentFrameCameraJitter, sizeof(float) * 2);
    m_cameraExtrinsic = extrinsic;
    const hm::Mat4 viewMatrix = viewMatrixFromCameraExtrinsic(extrinsic);
    const hm::Mat4 projectionMatrix = projectionMatrixFromCameraIntrinsic(m_cameraIntrinsic);
    if (m_taaSettings.enabled) {
        const hm::Vec2 jitterInPixels = m_taa.computeProjectionMatrixJitter(m_frameIndex);
        m_taa.updateTaaResolveWeights(backend, jitterInPixels);
        m_globalShaderInfo.currentFrameCameraJitter[0] = jitterInPixels.x * (1.f / (float)m_screenWidth);
        m_globalShaderInfo.currentFrameCameraJitter[1] = jitterInPixels.y * (1.f / (float)m_screenHeight);
        hm::Vec2 off;
        off.x = m_globalShaderInfo.currentFrameCameraJitter[0]; off.y = m_globalShaderInfo.currentFrameCameraJitter[1];
        m_viewProjectionMatrix = m_taa.applyProjectionMatrixJitter(projectionMatrix, off) * viewMatrix;
    } else {
        m_globalShaderInfo.currentFrameCameraJitter[0] = 0.f; m_globalShaderInfo.currentFrameCameraJitter[1] = 0.f;
        m_viewProjectionMatrix = projectionMatrix * viewMatrix;
    }
    std::memcpy(m_globalShaderInfo.viewProjectionPrevious, m_globalShaderInfo.viewProjection, sizeof(float) * 16);
    std::memcpy(m_globalShaderInfo.viewProjection, m_viewProjectionMatrix.m, sizeof(float) * 16);
    CameraIntrinsic in = m_cameraIntrinsic;
    m_cameraFrustum = computeViewFrustum(extrinsic, in);
    m_sunShadowFrustum = computeOrthogonalFrustumFittedToCamera(m_cameraFrustum, directionToVector(m_sunDirection));  // updateShadowFrustum :1054-1056
}

void RenderFrontend::prepareForDrawcalls() { updateGlobalShaderInfo(); }

void RenderFrontend::updateGlobalShaderInfo() {
    plain_global_shader_info& g = m_globalShaderInfo;
    const hm::Vec3 sun = directionToVector(m_sunDirection);
    g.sunDirection[0] = sun.x; g.sunDirection[1] = sun.y; g.sunDirection[2] = sun.z; g.sunDirection[3] = 0.f;
    std::memcpy(g.cameraPositionPrevious, g.cameraPosition, sizeof(float) * 4);
    g.cameraPosition[0] = m_cameraExtrinsic.position.x; g.cameraPosition[1] = m_cameraExtrinsic.position.y; g.cameraPosition[2] = m_cameraExtrinsic.position.z; g.cameraPosition[3] = 1.f;
    g.deltaTime = m_deltaTime;
    g.time = m_time;
    g.nearPlane = m_cameraIntrinsic.near;
    g.farPlane = m_cameraIntrinsic.far;
    g.cameraRight[0] = m_cameraExtrinsic.right.x; g.cameraRight[1] = m_cameraExtrinsic.right.y; g.cameraRight[2] = m_cameraExtrinsic.right.z; g.cameraRight[3] = 0.f;
    g.cameraUp[0] = m_cameraExtrinsic.up.x; g.cameraUp[1] = m_cameraExtrinsic.up.y; g.cameraUp[2] = m_cameraExtrinsic.up.z; g.cameraUp[3] = 0.f;
    std::memcpy(g.cameraForwardPrevious, g.cameraForward, sizeof(float) * 4);
    g.cameraForward[0] = m_cameraExtrinsic.forward.x; g.cameraForward[1] = m_cameraExtrinsic.forward.y; g.cameraForward[2] = m_cameraExtrinsic.forward.z; g.cameraForward[3] = 0.f;
    g.cameraTanFovHalf = dm::tan(hm::radians(m_cameraIntrinsic.fov) * 0.5f);
    g.cameraAspectRatio = m_cameraIntrinsic.aspectRatio;
    g.screenResolution[0] = (int32_t)m_screenWidth; g.screenResolution[1] = (int32_t)m_screenHeight;
    g.mipBias = (m_taaSettings.enabled && m_taaSettings.useMipBias) ? dm::log2(0.5f) : 0.f;
    backend.setUniformBufferData(m_globalUniformBuffer, &g, sizeof(g));
}

bool isAxisAlignedBoundingBoxIntersectingViewFrustum(const ViewFrustum& f, const hm::AABB& bb) {  // Culling.cpp:5-42
    const hm::Vec3 planes[6][2] = {{f.l_u_f, f.top}, {f.l_l_f, f.bot}, {f.l_u_n, f.near}, {f.l_u_f, f.far}, {f.l_u_f, f.left}, {f.r_u_f, f.right}};
    for (int i = 0; i < 6; i++) {
        bool isBBOutsidePlane = true;
        for (int k = 0; k < 8; k++) {
            const hm::Vec3 bp((k & 1) ? bb.max.x : bb.min.x, (k & 2) ? bb.max.y : bb.min.y, (k & 4) ? bb.max.z : bb.min.z);
            isBBOutsidePlane = isBBOutsidePlane && hm::dot(bp - planes[i][0], planes[i][1]) > 0.f;
        }
        if (isBBOutsidePlane) return false;
    }
    return true;
}

// RenderFrontend.cpp:533-683: SDF instance table, then - when the raster passes run here - frustum culling, the per-draw matrices
// and the draw calls of the depth prepass, the G-buffer fill (the reference's main pass) and the shadow cascades
void RenderFrontend::renderScene(const std::vector<RenderObject>& scene) {
    m_sdfGi.updateSDFScene(backend, scene, m_frontendMeshes);
    if (!m_rasterInputs) return;
    struct MainPassPushConstants { uint32_t albedoTextureIndex, normalTextureIndex, specularTextureIndex, transformIndex; };
    struct MainPassMatrices { hm::Mat4 model, mvp, mvpPrevious; };
    std::vector<MainPassPushConstants> mainPassPushConstants;
    std::vector<MeshHandle> mainPassCulledMeshes;
    std::vector<MainPassMatrices> mainPassMatrices;
    hm::Mat4 previousViewProjection;
    std::memcpy(previousViewProjection.m, m_globalShaderInfo.viewProjectionPrevious, sizeof(float) * 16);
    for (const RenderObject& obj : scene) {
        const MeshFrontend& mesh = m_frontendMeshes[obj.mesh];
        if (mesh.backendHandle.index == PLAIN_INVALID_INDEX) continue;  // an SDF-only mesh
        if (!isAxisAlignedBoundingBoxIntersectingViewFrustum(m_cameraFrustum, obj.bbWorld)) continue;
        m_currentMainPassDrawcallCount++;
        mainPassCulledMeshes.push_back(mesh.backendHandle);
        mainPassPushConstants.push_back({mesh.material.albedoTextureIndex, mesh.material.normalTextureIndex, mesh.material.specularTextureIndex, (uint32_t)mainPassMatrices.size()});
        mainPassMatrices.push_back({obj.modelMatrix, m_viewProjectionMatrix * obj.modelMatrix, previousViewProjection * obj.previousModelMatrix});
    }
    // only the prepass draws are needed for the SDF debug visualisation (:588-601)
    const bool renderingSDFVisualisation = m_sdfDebugSettings.visualisationMode != SDFVisualisationMode::None;
    if (!renderingSDFVisualisation) backend.drawMeshes(mainPassCulledMeshes, (const char*)mainPassPushConstants.data(), m_gbufferFillPass, 0);
    backend.drawMeshes(mainPassCulledMeshes, (const char*)mainPassPushConstants.data(), m_depthPrePass, 0);
    if (!mainPassMatrices.empty()) backend.setStorageBufferData(m_mainPassTransformsBuffer, mainPassMatrices.data(), sizeof(MainPassMatrices) * mainPassMatrices.size());
    // shadow pass (:607-652): coarse culling against the orthographic frustum fitted to the camera frustum, its near plane pushed
    // 10 km towards the sun so that casters outside the view are kept (the normals are those of the unshifted box, as in the reference)
    struct ShadowPushConstants { uint32_t albedoTextureIndex, transformIndex; };
    std::vector<MeshHandle> shadowCulledMeshes;
    std::vector<ShadowPushConstants> shadowPushConstantData;
    std::vector<hm::Mat4> shadowModelMatrices;
    const hm::Vec3 nearPlaneOffset = directionToVector(m_sunDirection) * 10000.f;
    m_sunShadowFrustum.l_l_n = m_sunShadowFrustum.l_l_n + nearPlaneOffset;
    m_sunShadowFrustum.r_l_n = m_sunShadowFrustum.r_l_n + nearPlaneOffset;
    m_sunShadowFrustum.l_u_n = m_sunShadowFrustum.l_u_n + nearPlaneOffset;
    m_sunShadowFrustum.r_u_n = m_sunShadowFrustum.r_u_n + nearPlaneOffset;
    for (const RenderObject& obj : scene) {
        const MeshFrontend& mesh = m_frontendMeshes[obj.mesh];
        if (mesh.backendHandle.index == PLAIN_INVALID_INDEX) continue;
        if (!isAxisAlignedBoundingBoxIntersectingViewFrustum(m_sunShadowFrustum, obj.bbWorld)) continue;
        m_currentShadowPassDrawcallCount++;
        shadowCulledMeshes.push_back(mesh.backendHandle);
        shadowPushConstantData.push_back({mesh.material.albedoTextureIndex, (uint32_t)shadowModelMatrices.size()});
        shadowModelMatrices.push_back(obj.modelMatrix);
    }
    for (int shadowPass = 0; shadowPass < m_shadingConfig.sunShadowCascadeCount; shadowPass++)
        backend.drawMeshes(shadowCulledMeshes, (const char*)shadowPushConstantData.data(), m_shadowPasses[shadowPass], 0);
    if (!shadowModelMatrices.empty()) backend.setStorageBufferData(m_shadowPassTransformsBuffer, shadowModelMatrices.data(), sizeof(hm::Mat4) * shadowModelMatrices.size());
}

void RenderFrontend::renderDepthPrepass(ImageHandle depth, ImageHandle normal, ImageHandle motion) {  // :792-802
    GraphicPassExecution e;
    e.genericInfo.handle = m_depthPrePass;
    e.targets = {RenderTarget{motion, 0}, RenderTarget{normal, 0}, RenderTarget{depth, 0}};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_mainPassTransformsBuffer, true, 0)};
    // row sharding: the rank's band + 16 rows (what a sharded caller uploads otherwise: the stencils of HiZ, trace, shading, TAA read
    // that far); the motion vectors are read at reprojected positions by the GI temporal filter and the TAA resolve: all-gathered
    backend.band(1, m_screenHeight, 16, &e.rowBegin, &e.rowEnd);
    backend.setGraphicPassExecution(e);
    if (backend.shard.active()) {
        ExchangeRequest x;
        x.kind = PLAIN_EXCHANGE_ALLGATHER_ROWS;
        x.name = "motion";
        x.images.push_back(motion); x.mips.push_back(0); x.divisors.push_back(1);
        backend.addExchange(x);
    }
}
void RenderFrontend::renderSunShadowCascades() {  // :760-775
    for (int i = 0; i < m_shadingConfig.sunShadowCascadeCount; i++) {
        GraphicPassExecution e;
        e.genericInfo.handle = m_shadowPasses[i];
        e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_sunShadowInfoBuffer, true, 0), StorageBufferResource(m_shadowPassTransformsBuffer, true, 1)};
        e.targets = {RenderTarget{m_shadowMaps[i], 0}};
        backend.setGraphicPassExecution(e);
    }
}
// the raster half of the reference's main pass (renderForwardShading: triangle.vert/.frag with depth test EQUAL, :1537-1561): the
// interpolated inputs and material texels land in the packed G-buffer, the shading itself runs in gbufferShading.comp
void RenderFrontend::fillGBuffer(ImageHandle gbuffer, ImageHandle depth) {
    GraphicPassExecution e;
    e.genericInfo.handle = m_gbufferFillPass;
    e.targets = {RenderTarget{gbuffer, 0}, RenderTarget{depth, 0}};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_mainPassTransformsBuffer, true, 17)};
    backend.band(1, m_screenHeight, 16, &e.rowBegin, &e.rowEnd);  // the same rows as the prepass whose visibility buffer it resolves
    backend.setGraphicPassExecution(e);
}
void RenderFrontend::setMeshGeometry(uint32_t mesh, const MeshBinary& geometry, const Material* material) {
    if (mesh >= m_frontendMeshes.size()) throw std::runtime_error("setMeshGeometry: invalid mesh");
    m_frontendMeshes[mesh].backendHandle = backend.createMeshes({geometry})[0];
    Material m;
    m.albedoTextureIndex = backend.getImageGlobalTextureArrayIndex(m_defaultTextures.diffuse);
    m.normalTextureIndex = backend.getImageGlobalTextureArrayIndex(m_defaultTextures.normal);
    m.specularTextureIndex = backend.getImageGlobalTextureArrayIndex(m_defaultTextures.specular);
    if (material) {
        if (material->albedoTextureIndex != PLAIN_INVALID_INDEX) m.albedoTextureIndex = material->albedoTextureIndex;
        if (material->normalTextureIndex != PLAIN_INVALID_INDEX) m.normalTextureIndex = material->normalTextureIndex;
        if (material->specularTextureIndex != PLAIN_INVALID_INDEX) m.specularTextureIndex = material->specularTextureIndex;
    }
    m_frontendMeshes[mesh].material = m;
}

void RenderFrontend::renderFrame() {  // RenderFrontend.cpp:685-705
    plain_exchange x;
    while (renderFrameSegment(&x)) {}  // an unsharded frame has no exchanges: one segment
}
// runs the recorded passes up to the next exchange (row sharding); returns false when the frame has been submitted completely
bool RenderFrontend::renderFrameSegment(plain_exchange* pending) {
    if (!m_frameOpen) {
        m_frameOpen = true;
        m_globalShaderInfo.frameIndex++;
        m_globalShaderInfo.frameIndexMod2 = m_globalShaderInfo.frameIndex % 2;
        m_globalShaderInfo.frameIndexMod3 = m_globalShaderInfo.frameIndex % 3;
        m_globalShaderInfo.frameIndexMod4 = m_globalShaderInfo.frameIndex % 4;
    }
    if (backend.runSegment(pending)) return true;
    m_globalShaderInfo.cameraCut = 0;
    m_frameOpen = false;
    return false;
}

uint32_t RenderFrontend::registerSdfMesh(const uint16_t* texels, uint32_t rx, uint32_t ry, uint32_t rz, const hm::AABB& localBB, hm::Vec3 meanAlbedo) {
    ImageHandle img = backend.createImage(imageDesc3D(rx, ry, rz, PLAIN_FORMAT_R16_SFLOAT, PLAIN_USAGE_SAMPLED), texels, (size_t)rx * ry * rz * 2);
    MeshFrontend m;
    m.sdfTextureIndex = (int)backend.getImageGlobalTextureArrayIndex(img);
    m.meanAlbedo = meanAlbedo;
    m.localBB = localBB;
    m_frontendMeshes.push_back(m);
    return (uint32_t)m_frontendMeshes.size() - 1;
}

void RenderFrontend::computeColorBufferHistogram(ImageHandle lastFrameColor, bool exchange) {
    StorageBufferResource histogramPerTileResource(m_histogramPerTileBuffer, false, 0);
    StorageBufferResource histogramResource(m_histogramBuffer, false, 1);
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_histogramPerTilePass;
        e.genericInfo.resources.storageBuffers = {histogramPerTileResource, StorageBufferResource(m_lightBuffer, true, 3)};
        e.genericInfo.resources.sampledImages = {ImageResource(lastFrameColor, 0, 2)};
        e.dispatchCount[0] = ceilDiv(m_screenWidth, histogramTileSizeX);
        e.dispatchCount[1] = ceilDiv(m_screenHeight, histogramTileSizeY);
        setRows(backend, e, histogramTileSizeY, e.dispatchCount[1]);  // a rank bins the tile rows of its own band
        backend.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_histogramResetPass;
        e.genericInfo.resources.storageBuffers = {histogramResource};
        e.dispatchCount[0] = ceilDiv(nHistogramBins, 64);
        backend.setComputePassExecution(e);
    }
    {
        ComputePassExecution e;
        e.genericInfo.handle = m_histogramCombinePass;
        e.genericInfo.resources.storageBuffers = {histogramPerTileResource, histogramResource};
        e.dispatchCount[0] = ceilDiv(m_screenWidth, histogramTileSizeX) * ceilDiv(m_screenHeight, histogramTileSizeY);
        e.dispatchCount[1] = ceilDiv(nHistogramBins, 64);
        if (backend.shard.active()) {  // sums the rank's own tiles; the 128 partial counters are all-reduced below
            uint32_t a, b;
            backend.band(histogramTileSizeY, ceilDiv(m_screenHeight, histogramTileSizeY), 0, &a, &b);
            e.rowBegin = a * ceilDiv(m_screenWidth, histogramTileSizeX);
            e.rowEnd = b * ceilDiv(m_screenWidth, histogramTileSizeX);
        }
        backend.setComputePassExecution(e);
    }
    ExchangeRequest x;
    x.kind = PLAIN_EXCHANGE_ALLREDUCE_SUM_U32;
    x.name = "histogram";
    x.buffer = m_histogramBuffer.index;
    x.elementCount = nHistogramBins;
    if (exchange) backend.addExchange(x);
    else m_pendingHistogramExchange = x;
}

void RenderFrontend::computeExposure() {
    ComputePassExecution e;
    e.genericInfo.handle = m_preExposeLightsPass;
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_histogramBuffer, false, 1), StorageBufferResource(m_lightBuffer, false, 0)};
    e.genericInfo.resources.sampledImages = {ImageResource(m_sky.m_skyTransmissionLut, 0, 2)};
    backend.setComputePassExecution(e);
}

void RenderFrontend::computeDepthPyramid(ImageHandle depthBuffer, const FrameRenderTargets* alsoDownscale, bool holdRest) {
    ComputePassExecution e;
    e.genericInfo.handle = m_depthPyramidPass;
    const uint32_t width = m_screenWidth / 2, height = m_screenHeight / 2;
    const uint32_t mipCount = mipCountFromResolution(width, height, 1);
    // 11 in the reference (RenderFrontend.cpp:806); 7680x4320 (BASELINE configs[4]) needs a 12th level, bound at binding 11
    const uint32_t maxMipCount = mipCount > 11 ? mipCount : 11;
    uint32_t dc[2];
    computeSinglePassMipChainDispatchCount(width, height, mipCount, maxMipCount, dc);
    e.dispatchCount[0] = dc[0]; e.dispatchCount[1] = dc[1];
    e.genericInfo.resources.sampledImages = {ImageResource(depthBuffer, 0, 13), ImageResource(m_minMaxDepthPyramid, 0, 15)};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_depthPyramidSyncBuffer, false, 16)};
    const uint32_t unusedMipCount = maxMipCount > mipCount ? maxMipCount - mipCount : 0;
    for (uint32_t i = 0; i < maxMipCount; i++) {
        const uint32_t mipLevel = i >= unusedMipCount ? i - unusedMipCount : 0;
        e.genericInfo.resources.storageImages.push_back(ImageResource(m_minMaxDepthPyramid, mipLevel, i));
    }
    if (!backend.shard.active()) {
        backend.setComputePassExecution(e);
        if (alsoDownscale) downscaleDepth(*alsoDownscale, true);
        return;
    }
    // sharded: the levels reduced from a rank's own depth rows (up to four: band boundaries are multiples of 32 rows), then
    // an all-gather of the last of them, then the remaining small levels on every rank
    uint32_t fused = 0, sw = m_screenWidth, sh = m_screenHeight;
    for (uint32_t k = 0; k < mipCount && k < 4 && sw % 2 == 0 && sh % 2 == 0 && sw >= 2 && sh >= 2; k++) { fused = k + 1; sw /= 2; sh /= 2; }
    ComputePassExecution local = e;
    local.shardPhase = 1;
    setRows(backend, local, 2, height);
    backend.setComputePassExecution(local);
    ExchangeRequest x = gatherRows("hiz", {m_minMaxDepthPyramid}, fused - 1, 2u << (fused - 1));
    if (alsoDownscale) {  // the depth downscale does not depend on the pyramid: its all-gather shares this exchange's barrier
        downscaleDepth(*alsoDownscale, false);
        x.name = "hiz+depthHalf";
        x.images.push_back(m_depthHalfRes); x.mips.push_back(0); x.divisors.push_back(2);
    }
    if (m_motionRowsOnly && !m_rasterInputs) {  // the uploaded band of motion vectors rides on the same barrier
        x.name += "+motion";
        x.images.push_back(m_frameRenderTargets[m_sceneRenderTargetIndex].motionBuffer); x.mips.push_back(0); x.divisors.push_back(1);
    }
    backend.addExchange(x);
    ComputePassExecution rest = e;
    rest.shardPhase = 2;
    if (holdRest) m_pendingPyramidRest = rest;  // emitted by the caller after the exchange that follows (prepareRenderpasses)
    else backend.setComputePassExecution(rest);
}

void RenderFrontend::computeSunLightMatrices() {
    ComputePassExecution e;
    e.genericInfo.handle = m_lightMatrixPass;
    const uint32_t depthPyramidMipCount = mipCountFromResolution(m_screenWidth / 2, m_screenHeight / 2, 1);
    e.genericInfo.resources.storageImages = {ImageResource(m_minMaxDepthPyramid, depthPyramidMipCount - 1, 1)};
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_sunShadowInfoBuffer, false, 0)};
    float pc[2];
    pc[0] = m_sdfTraceSettings.traceInfluenceRadius;
    pc[1] = m_volumetricsSettings.maxDistance;
    if (!m_sdfTraceSettings.strictInfluenceRadiusCutoff) pc[0] += m_sdfTraceSettings.additionalSunShadowMapPadding;
    e.pushConstants = dataToCharArray(pc, sizeof(pc));
    backend.setComputePassExecution(e);
}

void RenderFrontend::downscaleDepth(const FrameRenderTargets& current, bool exchange) {
    ComputePassExecution e;
    e.genericInfo.handle = m_depthDownscalePass;
    e.dispatchCount[0] = ceilDiv(m_screenWidth / 2, 8);
    e.dispatchCount[1] = ceilDiv(m_screenHeight / 2, 8);
    e.genericInfo.resources.storageImages = {ImageResource(m_depthHalfRes, 0, 0)};
    e.genericInfo.resources.sampledImages = {ImageResource(current.depthBuffer, 0, 1)};
    setRows(backend, e, 2, m_screenHeight / 2);
    backend.setComputePassExecution(e);
    if (exchange) backend.addExchange(gatherRows("depthHalf", {m_depthHalfRes}, 0, 2));  // the spatial GI filter reads it at arbitrary rows
}

// renderForwardShading :894-929 + Sky::renderSky (Sky.cpp:318-352), as one full-screen pass over the G-buffer.
// Bindings 3,7,8,9-12,15,16,18,19 are triangle.frag's; 0 = G-buffer, 20 = colour target, 21 = sky LUT (sky.frag
// binding 0), 22 = transmission LUT (sunSprite.frag binding 1). Push constants = SunSpriteMatrices (Sky.cpp:256-262).
void RenderFrontend::shadeGBuffer(ImageHandle colorTarget) {
    ComputePassExecution e;
    e.genericInfo.handle = m_shadingPass;
    const SDFGI::IndirectLightingImages indirect = m_sdfGi.getIndirectLightingResults(m_sdfTraceSettings.halfResTrace);
    e.genericInfo.resources.storageBuffers = {StorageBufferResource(m_lightBuffer, true, 7), StorageBufferResource(m_sunShadowInfoBuffer, true, 8)};
    e.genericInfo.resources.sampledImages = {ImageResource(gbuffer(), 0, 0), ImageResource(m_brdfLut, 0, 3), ImageResource(indirect.Y_SH, 0, 15), ImageResource(indirect.CoCg, 0, 16),
                                             ImageResource(m_volumetrics.m_volumetricIntegrationVolume, 0, 18), ImageResource(m_sky.m_skyLut, 0, 21),
                                             ImageResource(m_sky.m_skyTransmissionLut, 0, 22)};
    for (uint32_t i = 0; i < (uint32_t)maxSunShadowCascadeCount; i++) e.genericInfo.resources.sampledImages.push_back(ImageResource(m_shadowMaps[i], 0, 9 + i));
    e.genericInfo.resources.uniformBuffers = {UniformBufferResource(m_volumetrics.m_volumetricsSettingsUniforms, 19)};
    e.genericInfo.resources.storageImages = {ImageResource(colorTarget, 0, 20)};
    struct SunSpriteMatrices { hm::Mat4 model, mvp; } sun;
    sun.model = m_sky.sunSpriteModelMatrix(m_sunDirection);
    sun.mvp = hm::Mat4::identity();  // only the model matrix is used by the deferred recast; set in renderFrame order
    e.pushConstants = dataToCharArray(&sun, sizeof(sun));
    e.dispatchCount[0] = ceilDiv(m_screenWidth, 8);
    e.dispatchCount[1] = ceilDiv(m_screenHeight, 8);
    setRows(backend, e, 1, m_screenHeight, 8);  // 8 extra rows on both sides: the TAA resolve (+-4 rows, for bloom) reads +-2 rows of colour
    backend.setComputePassExecution(e);
}

void RenderFrontend::computeTonemapping(ImageHandle src) {
    ComputePassExecution e;
    e.genericInfo.handle = m_tonemappingPass;
    e.genericInfo.resources.storageImages = {ImageResource(backend.getSwapchainInputImage(), 0, 0)};
    e.genericInfo.resources.sampledImages = {ImageResource(src, 0, 1)};
    e.dispatchCount[0] = ceilDiv(m_screenWidth, 8);
    e.dispatchCount[1] = ceilDiv(m_screenHeight, 8);
    setRows(backend, e, 1, m_screenHeight);
    backend.setComputePassExecution(e);
    m_lastTonemapSource = src;
}

void RenderFrontend::computeBRDFLut() {
    ComputePassExecution e;
    e.genericInfo.handle = m_brdfLutPass;
    e.genericInfo.resources.storageImages = {ImageResource(m_brdfLut, 0, 0)};
    e.dispatchCount[0] = brdfLutRes / 8;
    e.dispatchCount[1] = brdfLutRes / 8;
    backend.setComputePassExecution(e);
}
