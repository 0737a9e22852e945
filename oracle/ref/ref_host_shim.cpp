// ORACLE - test infrastructure only. C entry points over the reference's OWN host-side functions that feed the frame path
// (Camera.cpp, ViewFrustum.cpp, Culling.cpp, MathUtils.cpp, sdfUtilities.cpp, CompressedTypes.cpp), compiled from the sources where
// they lie under /root/reference by oracle/build_ref.sh into oracle/_ref/libref_host.so. tests/test_host_vs_reference.py checks the
// host mirror (plainrenderer_b200/host/) against them. This file contains no reference code: it only calls it.
#include "pch.h"
#include "Runtime/Rendering/Camera.h"
#include "Runtime/Rendering/ViewFrustum.h"
#include "Runtime/Rendering/Culling.h"
#include "Common/Utilities/MathUtils.h"
#include "Common/sdfUtilities.h"
#include "Common/CompressedTypes.h"

static Camera makeCamera(const float* pos, const float* fwd, const float* right, const float* up, float fov, float aspect, float nearPlane, float farPlane) {
    Camera c;
    c.extrinsic.position = glm::vec3(pos[0], pos[1], pos[2]);
    c.extrinsic.forward = glm::vec3(fwd[0], fwd[1], fwd[2]);
    c.extrinsic.right = glm::vec3(right[0], right[1], right[2]);
    c.extrinsic.up = glm::vec3(up[0], up[1], up[2]);
    c.intrinsic.fov = fov; c.intrinsic.aspectRatio = aspect; c.intrinsic.near = nearPlane; c.intrinsic.far = farPlane;
    return c;
}
static void put(float* out, const glm::vec3& v) { out[0] = v.x; out[1] = v.y; out[2] = v.z; }
static void frustumOut(const ViewFrustum& f, float* points, float* normals) {
    const glm::vec3 p[8] = {f.points.l_l_n, f.points.l_l_f, f.points.l_u_n, f.points.l_u_f, f.points.r_l_n, f.points.r_l_f, f.points.r_u_n, f.points.r_u_f};
    const glm::vec3 n[6] = {f.normals.top, f.normals.bot, f.normals.right, f.normals.left, f.normals.near, f.normals.far};
    for (int i = 0; i < 8; i++) put(points + 3 * i, p[i]);
    for (int i = 0; i < 6; i++) put(normals + 3 * i, n[i]);
}
static ViewFrustum frustumIn(const float* points, const float* normals) {
    ViewFrustum f;
    glm::vec3* p[8] = {&f.points.l_l_n, &f.points.l_l_f, &f.points.l_u_n, &f.points.l_u_f, &f.points.r_l_n, &f.points.r_l_f, &f.points.r_u_n, &f.points.r_u_f};
    glm::vec3* n[6] = {&f.normals.top, &f.normals.bot, &f.normals.right, &f.normals.left, &f.normals.near, &f.normals.far};
    for (int i = 0; i < 8; i++) *p[i] = glm::vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    for (int i = 0; i < 6; i++) *n[i] = glm::vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    return f;
}

extern "C" {
__attribute__((visibility("default"))) void ref_hammersley2D(uint32_t index, float out[2]) { const glm::vec2 h = hammersley2D(index); out[0] = h.x; out[1] = h.y; }
__attribute__((visibility("default"))) void ref_directionToVector(const float anglesDeg[2], float out[3]) { put(out, directionToVector(glm::vec2(anglesDeg[0], anglesDeg[1]))); }
__attribute__((visibility("default"))) uint32_t ref_mipCountFromResolution(uint32_t w, uint32_t h, uint32_t d) { return mipCountFromResolution(w, h, d); }
__attribute__((visibility("default"))) void ref_cameraMatrices(const float* pos, const float* fwd, const float* right, const float* up, float fov, float aspect, float nearPlane, float farPlane,
                                                                float outView[16], float outProjection[16]) {
    const Camera c = makeCamera(pos, fwd, right, up, fov, aspect, nearPlane, farPlane);
    const glm::mat4 v = viewMatrixFromCameraExtrinsic(c.extrinsic), p = projectionMatrixFromCameraIntrinsic(c.intrinsic);
    memcpy(outView, &v[0][0], 64);  // column-major, as glm stores it
    memcpy(outProjection, &p[0][0], 64);
}
__attribute__((visibility("default"))) void ref_viewFrustum(const float* pos, const float* fwd, const float* right, const float* up, float fov, float aspect, float nearPlane, float farPlane,
                                                             float outPoints[24], float outNormals[18]) {
    frustumOut(computeViewFrustum(makeCamera(pos, fwd, right, up, fov, aspect, nearPlane, farPlane)), outPoints, outNormals);
}
__attribute__((visibility("default"))) void ref_orthogonalFrustumFittedToCamera(const float points[24], const float normals[18], const float lightDirection[3], float outPoints[24], float outNormals[18]) {
    frustumOut(computeOrthogonalFrustumFittedToCamera(frustumIn(points, normals), glm::vec3(lightDirection[0], lightDirection[1], lightDirection[2])), outPoints, outNormals);
}
__attribute__((visibility("default"))) int ref_aabbIntersectsFrustum(const float points[24], const float normals[18], const float bbMin[3], const float bbMax[3]) {
    AxisAlignedBoundingBox bb;
    bb.min = glm::vec3(bbMin[0], bbMin[1], bbMin[2]);
    bb.max = glm::vec3(bbMax[0], bbMax[1], bbMax[2]);
    return isAxisAlignedBoundingBoxIntersectingViewFrustum(frustumIn(points, normals), bb) ? 1 : 0;
}
__attribute__((visibility("default"))) void ref_padSDFBoundingBox(const float bbMin[3], const float bbMax[3], float outMin[3], float outMax[3]) {
    AxisAlignedBoundingBox bb;
    bb.min = glm::vec3(bbMin[0], bbMin[1], bbMin[2]);
    bb.max = glm::vec3(bbMax[0], bbMax[1], bbMax[2]);
    const AxisAlignedBoundingBox r = padSDFBoundingBox(bb);
    put(outMin, r.min); put(outMax, r.max);
}
// SDFGI.cpp:288-292 computes worldToLocal = glm::inverse(modelMatrix * glm::translate(glm::mat4(1.f), bbOffset)) inline in
// SDFGI::updateSDFScene (a method that needs the Vulkan backend); the same glm calls on the same operands
__attribute__((visibility("default"))) void ref_sdfWorldToLocal(const float model[16], const float bbOffset[3], float out[16]) {
    glm::mat4 m;
    memcpy(&m[0][0], model, 64);
    const glm::mat4 r = glm::inverse(m * glm::translate(glm::mat4(1.f), glm::vec3(bbOffset[0], bbOffset[1], bbOffset[2])));
    memcpy(out, &r[0][0], 64);
}
__attribute__((visibility("default"))) uint32_t ref_vec3ToNormalizedR10B10G10A2(const float v[3]) { return vec3ToNormalizedR10B10G10A2(glm::vec3(v[0], v[1], v[2])).value; }
}
