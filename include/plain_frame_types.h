/* plain_frame_types.h - byte layouts of the buffers the frame-path passes read and write.
 *
 * These are the std140/std430 blocks declared in the reference shaders; the host side (RenderFrontend mirror) fills
 * them and both backends (CUDA, CPU oracle) interpret the same bytes. Each struct cites the block it restates.
 */
#ifndef PLAIN_FRAME_TYPES_H
#define PLAIN_FRAME_TYPES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* global.inc:4-33 (std140) == GlobalShaderInfo, ResourceDescriptions.h:174-203. Matrices are column-major. */
typedef struct {
    float viewProjection[16];
    float viewProjectionPrevious[16];
    float sunDirection[4];
    float cameraPosition[4];
    float cameraPositionPrevious[4];
    float cameraRight[4];
    float cameraUp[4];
    float cameraForward[4];
    float cameraForwardPrevious[4];
    int32_t noiseTextureIndices[4];
    float currentFrameCameraJitter[2];
    float previousFrameCameraJitter[2];
    int32_t screenResolution[2];
    float cameraTanFovHalf;
    float cameraAspectRatio;
    float nearPlane;
    float farPlane;
    float sunStrength; /* sunIlluminanceLux on the host side */
    float exposureOffset;
    float exposureAdaptionSpeedEvPerSec;
    float deltaTime;
    float time;
    float mipBias;
    uint32_t cameraCut; /* bool */
    uint32_t frameIndex;
    uint32_t frameIndexMod2;
    uint32_t frameIndexMod3;
    uint32_t frameIndexMod4;
} plain_global_shader_info; /* 340 bytes */

/* lightBuffer.inc:4-8 (std430) */
typedef struct {
    float sunColor[3];
    float previousFrameExposure;
    float sunStrengthExposed;
} plain_light_buffer; /* 20 bytes */

/* sunShadowCascades.inc:7-11 (std430) */
typedef struct {
    float splits[4];
    float lightMatrices[4][16];
    float lightSpaceScale[4][2];
} plain_shadow_cascade_info; /* 304 bytes */

/* SDF.inc:4-10 == SDFGI.h:31-37 */
typedef struct {
    float localExtends[3];
    uint32_t sdfTextureIndex;
    float meanAlbedo[3];
    float padding;
    float worldToLocal[16];
} plain_sdf_instance; /* 96 bytes */

/* sdfCulling.inc:5-15 */
#define PLAIN_MAX_OBJECTS_PER_TILE 100
#define PLAIN_SDF_CULLING_TILE_SIZE 32
typedef struct {
    uint32_t objectCount;
    uint32_t indices[PLAIN_MAX_OBJECTS_PER_TILE];
} plain_culled_instances_per_tile; /* 404 bytes */
typedef struct {
    float bbMin[3]; float padding1;
    float bbMax[3]; float padding2;
} plain_bounding_box; /* 32 bytes */

/* sky.inc:1-10 (std140) == AtmosphereSettings, Sky.h:6-15 */
typedef struct {
    float scatteringRayleighGround[3];
    float earthRadius;
    float extinctionRayleighGround[3];
    float atmosphereHeight;
    float ozoneExtinction[3];
    float scatteringMieGround;
    float extinctionMieGround;
    float mieScatteringExponent;
} plain_atmosphere_settings; /* 56 bytes */

/* volumetricFroxelLighting.inc:6-16 (std140) == VolumetricsBufferContents, Volumetrics.h:51-59 */
typedef struct {
    float windSampleOffset[3];
    float sampleOffset;
    float scatteringCoefficients[3];
    float maxDistance;
    float absorptionCoefficient;
    float baseDensity;
    float densityNoiseRange;
    float densityNoiseScale;
    float phaseFunctionG;
} plain_volumetric_lighting_settings; /* 52 bytes */

/* sdfCameraFrustumCulling.comp:13-16 (std140) */
typedef struct {
    float frustumPoints[6][4];
    float frustumNormals[6][4];
} plain_camera_frustum_buffer; /* 192 bytes */

/* Packed G-buffer texel, PLAIN_FORMAT_RGBA32_UINT (defined by this build as the post-raster inputs of
 * triangle.frag:76-78,178-193):
 *   x: depth, float bits (D32F reverse-Z, 0 = sky)
 *   y: shading normal N (after normal mapping), octahedral, 2 x SNORM16 (x low half)
 *   z: albedo texel sRGB R | G<<8 | B<<16, specular-texture G (roughness) << 24
 *   w: specular-texture B (metalness) in bits 0-7, rest 0
 */

#ifdef __cplusplus
}
#endif
#endif
