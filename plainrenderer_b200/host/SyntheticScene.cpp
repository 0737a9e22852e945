// SyntheticScene.cpp - stand-in for the parts before the frame path that are out of scope here: the .plain/SDF
// loader (a seeded scene of oriented boxes with analytic SDF bricks, brick resolution by the reference's rule
// SceneSDF.cpp:120-131) and the raster passes depthPrepass / sunShadow (a CPU ray caster writing what they would
// write: D32F reverse-Z depth, RG16_SNORM motion as depthPrepass.frag:33-41, RGBA8 geometric normal as
// depthPrepass.frag:46-48, D16 shadow maps fitted with lightMatrix.comp's formulas) plus the packed G-buffer
// (include/plain_frame_types.h). Host-only; none of this is on the timed frame path.
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include "RenderFrontend.h"
#include "plain_frontend.h"

using namespace hm;

struct SynthMesh { Vec3 half; uint8_t albedo[3]; uint8_t roughness, metal; uint32_t res[3]; std::vector<uint16_t> sdf; AABB localBB; };
struct SynthObject { uint32_t mesh; Mat4 model, worldToLocal; AABB bbWorld; };
struct plain_synthetic_scene {
    std::vector<SynthMesh> meshes;
    std::vector<SynthObject> objects;
};

static uint32_t rngNext(uint32_t& s) { s = s * 747796405u + 2891336453u; uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u; return (w >> 22u) ^ w; }
static float rngFloat(uint32_t& s) { return (float)(rngNext(s) >> 8) * (1.f / 16777216.f); }
static uint32_t hash2(uint32_t a, uint32_t b) { uint32_t x = a * 0x9e3779b1u ^ (b + 0x7f4a7c15u + (a << 6) + (a >> 2)); x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

static uint16_t floatToHalfBits(float f) {  // round to nearest even
    uint32_t u = dm::f2u(f), sign = (u >> 16) & 0x8000u, au = u & 0x7fffffffu;
    if (au >= 0x47800000u) return (uint16_t)(sign | (au > 0x7f800000u ? 0x7e00u : 0x7c00u));
    if (au < 0x38800000u) {
        if (au < 0x33000000u) return (uint16_t)sign;
        uint32_t m = (au & 0x7fffffu) | 0x800000u;
        int shift = 126 - (int)(au >> 23);
        uint32_t v = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (v & 1u))) v++;
        return (uint16_t)(sign | v);
    }
    uint32_t v = ((au - 0x38000000u) >> 13), rem = au & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (v & 1u))) v++;
    return (uint16_t)(sign | v);
}
static uint32_t nextPowerOfTwo(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

static void bakeBoxSdf(SynthMesh& m) {
    m.localBB.min = -m.half;
    m.localBB.max = m.half;
    const Vec3 ext = m.localBB.max - m.localBB.min;
    const float e[3] = {ext.x, ext.y, ext.z};
    for (int c = 0; c < 3; c++) {
        uint32_t r = nextPowerOfTwo((uint32_t)(e[c] / 0.25f));
        m.res[c] = r < 16 ? 16 : (r > 64 ? 64 : r);
    }
    const AABB padded = padSDFBoundingBox(m.localBB);
    const Vec3 pe = padded.max - padded.min;
    m.sdf.resize((size_t)m.res[0] * m.res[1] * m.res[2]);
    for (uint32_t z = 0; z < m.res[2]; z++)
        for (uint32_t y = 0; y < m.res[1]; y++)
            for (uint32_t x = 0; x < m.res[0]; x++) {
                Vec3 p(padded.min.x + ((float)x + 0.5f) / (float)m.res[0] * pe.x, padded.min.y + ((float)y + 0.5f) / (float)m.res[1] * pe.y, padded.min.z + ((float)z + 0.5f) / (float)m.res[2] * pe.z);
                Vec3 q(std::fabs(p.x) - m.half.x, std::fabs(p.y) - m.half.y, std::fabs(p.z) - m.half.z);
                Vec3 qp(q.x > 0 ? q.x : 0, q.y > 0 ? q.y : 0, q.z > 0 ? q.z : 0);
                float inside = q.x > q.y ? (q.x > q.z ? q.x : q.z) : (q.y > q.z ? q.y : q.z);
                float d = length(qp) + (inside < 0.f ? inside : 0.f);
                m.sdf[((size_t)z * m.res[1] + y) * m.res[0] + x] = floatToHalfBits(d);
            }
}

static void addObject(plain_synthetic_scene& s, uint32_t mesh, Vec3 position, float yawDeg) {
    SynthObject o;
    o.mesh = mesh;
    o.model = translate(position) * rotate(radians(yawDeg), Vec3(0.f, 1.f, 0.f));
    o.worldToLocal = inverse(o.model);
    const Vec3 h = s.meshes[mesh].half;
    o.bbWorld.min = Vec3(1e30f);
    o.bbWorld.max = Vec3(-1e30f);
    for (int k = 0; k < 8; k++) {
        Vec4 p = o.model * Vec4((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z, 1.f);
        o.bbWorld.min = vmin(o.bbWorld.min, Vec3(p.x, p.y, p.z));
        o.bbWorld.max = vmax(o.bbWorld.max, Vec3(p.x, p.y, p.z));
    }
    s.objects.push_back(o);
}

// up is -y (CameraExtrinsic default up = (0,-1,0)). An atrium of about 30 x 15 x 18 m: floor slab, two side walls,
// two rows of pillars, crates on the floor and a few beams overhead.
static void buildScene(plain_synthetic_scene& s, uint32_t seed, uint32_t nInstances) {
    uint32_t r = seed ? seed : 0x504c4149u;
    auto addMesh = [&](Vec3 half) {
        SynthMesh m;
        m.half = half;
        for (int c = 0; c < 3; c++) m.albedo[c] = (uint8_t)(255.f * (0.2f + 0.6f * rngFloat(r)));
        m.roughness = (uint8_t)(255.f * (0.2f + 0.7f * rngFloat(r)));
        m.metal = rngFloat(r) < 0.1f ? 255 : 0;
        bakeBoxSdf(m);
        s.meshes.push_back(std::move(m));
        return (uint32_t)s.meshes.size() - 1;
    };
    const uint32_t floorMesh = addMesh(Vec3(15.f, 0.25f, 9.f));
    const uint32_t wallMesh = addMesh(Vec3(15.f, 7.5f, 0.3f));
    const uint32_t pillarMesh = addMesh(Vec3(0.45f, 5.f, 0.45f));
    const uint32_t beamMesh = addMesh(Vec3(0.35f, 0.3f, 7.f));
    std::vector<uint32_t> crateMeshes;
    for (int i = 0; i < 12; i++) crateMeshes.push_back(addMesh(Vec3(0.3f + 1.2f * rngFloat(r), 0.3f + 1.0f * rngFloat(r), 0.3f + 1.2f * rngFloat(r))));
    addObject(s, floorMesh, Vec3(0.f, 0.25f, 0.f), 0.f);
    if (nInstances >= 3) {
        addObject(s, wallMesh, Vec3(0.f, -7.5f, -9.3f), 0.f);
        addObject(s, wallMesh, Vec3(0.f, -7.5f, 9.3f), 0.f);
    }
    uint32_t remaining = nInstances > (uint32_t)s.objects.size() ? nInstances - (uint32_t)s.objects.size() : 0;
    const uint32_t nPillars = remaining / 5 < 16 ? remaining / 5 : 16, nBeams = remaining / 10 < 8 ? remaining / 10 : 8;
    for (uint32_t i = 0; i < nPillars; i++) {
        float x = -13.f + 26.f * (float)(i / 2) / (float)((nPillars + 1) / 2 > 1 ? (nPillars + 1) / 2 - 1 : 1);
        addObject(s, pillarMesh, Vec3(x, -5.f, (i & 1) ? 5.5f : -5.5f), 0.f);
    }
    for (uint32_t i = 0; i < nBeams; i++) addObject(s, beamMesh, Vec3(-12.f + 24.f * ((float)i + 0.5f) / (float)nBeams, -10.3f, 0.f), 0.f);
    while (s.objects.size() < nInstances) {
        uint32_t m = crateMeshes[rngNext(r) % crateMeshes.size()];
        Vec3 pos(-13.f + 26.f * rngFloat(r), -s.meshes[m].half.y, -7.5f + 15.f * rngFloat(r));
        addObject(s, m, pos, 90.f * rngFloat(r));
    }
}

struct Hit { float tEnter, tExit; int object; Vec3 normalWorld; };
static bool castRay(const plain_synthetic_scene& s, Vec3 o, Vec3 d, Hit& hit) {
    hit.tEnter = 1e30f;
    hit.object = -1;
    for (size_t i = 0; i < s.objects.size(); i++) {
        const SynthObject& ob = s.objects[i];
        const Vec3 h = s.meshes[ob.mesh].half;
        Vec4 ol = ob.worldToLocal * Vec4(o, 1.f), dl = ob.worldToLocal * Vec4(d, 0.f);
        const float oo[3] = {ol.x, ol.y, ol.z}, dd[3] = {dl.x, dl.y, dl.z}, hh[3] = {h.x, h.y, h.z};
        float t0 = 0.f, t1 = 1e30f;
        int axis = -1;
        bool miss = false;
        for (int a = 0; a < 3; a++) {
            if (std::fabs(dd[a]) < 1e-12f) {
                if (std::fabs(oo[a]) > hh[a]) { miss = true; break; }
                continue;
            }
            float inv = 1.f / dd[a];
            float ta = (-hh[a] - oo[a]) * inv, tb = (hh[a] - oo[a]) * inv;
            if (ta > tb) { float t = ta; ta = tb; tb = t; }
            if (ta > t0) { t0 = ta; axis = a; }
            if (tb < t1) t1 = tb;
            if (t0 > t1) { miss = true; break; }
        }
        if (miss || axis < 0 || t0 >= hit.tEnter) continue;
        hit.tEnter = t0;
        hit.tExit = t1;
        hit.object = (int)i;
        Vec3 nl(0.f);
        (&nl.x)[axis] = dd[axis] > 0.f ? -1.f : 1.f;
        Vec4 nw = ob.model * Vec4(nl, 0.f);
        hit.normalWorld = normalize(Vec3(nw.x, nw.y, nw.z));
    }
    return hit.object >= 0;
}

static Mat4 viewMatrix(const plain_camera_extrinsic& c) {
    Mat4 v = Mat4::identity();
    v.at(0, 0) = c.right[0]; v.at(0, 1) = c.right[1]; v.at(0, 2) = c.right[2];
    v.at(1, 0) = c.up[0]; v.at(1, 1) = c.up[1]; v.at(1, 2) = c.up[2];
    v.at(2, 0) = -c.forward[0]; v.at(2, 1) = -c.forward[1]; v.at(2, 2) = -c.forward[2];
    v = transpose(v);
    return v * translate(Vec3(-c.position[0], -c.position[1], -c.position[2]));
}
static Mat4 projectionMatrix(const plain_frontend_settings& st) {
    Mat4 p = perspective(radians(st.camera_fov_deg), (float)st.width / (float)st.height, st.camera_near, st.camera_far);
    Mat4 c = Mat4::identity();
    c.at(1, 1) = -1.f; c.at(2, 2) = -0.5f; c.at(3, 2) = 0.5f;
    return c * p;
}
static int16_t toSnorm16(float v) { v = v < -1.f ? -1.f : (v > 1.f ? 1.f : v); float s = v * 32767.f; return (int16_t)(s >= 0.f ? s + 0.5f : s - 0.5f); }
static uint8_t toUnorm8(float v) { v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v); return (uint8_t)(v * 255.f + 0.5f); }
static uint32_t octEncode(Vec3 n) {
    float l1 = std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z);
    float x = n.x / l1, y = n.y / l1;
    if (n.z < 0.f) {
        float ox = (1.f - std::fabs(y)) * (x >= 0.f ? 1.f : -1.f), oy = (1.f - std::fabs(x)) * (y >= 0.f ? 1.f : -1.f);
        x = ox; y = oy;
    }
    return (uint32_t)(uint16_t)toSnorm16(x) | ((uint32_t)(uint16_t)toSnorm16(y) << 16);
}

static void parallelRows(int threads, int rows, const std::function<void(int)>& fn) {
    int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > rows) t = rows;
    std::vector<std::thread> pool;
    for (int k = 0; k < t; k++) pool.emplace_back([=, &fn]() { for (int y = k; y < rows; y += t) fn(y); });
    for (auto& th : pool) th.join();
}

extern "C" {

int PLAIN_FE(synthetic_scene_create)(uint32_t seed, uint32_t n_instances, plain_synthetic_scene** out) {
    if (!out) return 1;
    plain_synthetic_scene* s = new plain_synthetic_scene();
    buildScene(*s, seed, n_instances < 1 ? 1 : n_instances);
    *out = s;
    return 0;
}
void PLAIN_FE(synthetic_scene_destroy)(plain_synthetic_scene* s) { delete s; }

int PLAIN_FE(synthetic_scene_attach)(plain_synthetic_scene* s, plain_frontend* fe) {
    std::vector<uint32_t> meshIds;
    for (auto& m : s->meshes) {
        const float mn[3] = {m.localBB.min.x, m.localBB.min.y, m.localBB.min.z}, mx[3] = {m.localBB.max.x, m.localBB.max.y, m.localBB.max.z};
        const float al[3] = {m.albedo[0] / 255.f, m.albedo[1] / 255.f, m.albedo[2] / 255.f};
        uint32_t id = 0;
        if (PLAIN_FE(register_sdf_mesh)(fe, m.sdf.data(), m.res[0], m.res[1], m.res[2], mn, mx, al, &id)) return 1;
        meshIds.push_back(id);
    }
    std::vector<uint32_t> idx;
    std::vector<float> mats, bmin, bmax;
    for (auto& o : s->objects) {
        idx.push_back(meshIds[o.mesh]);
        mats.insert(mats.end(), o.model.m, o.model.m + 16);
        bmin.insert(bmin.end(), {o.bbWorld.min.x, o.bbWorld.min.y, o.bbWorld.min.z});
        bmax.insert(bmax.end(), {o.bbWorld.max.x, o.bbWorld.max.y, o.bbWorld.max.z});
    }
    return PLAIN_FE(set_scene)(fe, (uint32_t)idx.size(), idx.data(), mats.data(), bmin.data(), bmax.data());
}

int PLAIN_FE(synthetic_scene_render_inputs)(plain_synthetic_scene* s, const plain_frontend_settings* st, const plain_camera_extrinsic* cam, const plain_camera_extrinsic* prevCam,
                                            uint32_t frame_index, void* depthOut, void* motionOut, void* normalOut, void* gbufferOut, void* const shadowMaps[4], int32_t threads) {
    const int W = (int)st->width, H = (int)st->height;
    const Vec3 camPos(cam->position[0], cam->position[1], cam->position[2]), fwd(cam->forward[0], cam->forward[1], cam->forward[2]);
    const Vec3 up(cam->up[0], cam->up[1], cam->up[2]), right(cam->right[0], cam->right[1], cam->right[2]);
    const float tanHalf = dm::tan(radians(st->camera_fov_deg) * 0.5f), aspect = (float)W / (float)H;
    Vec2 jitter;
    jitter.x = 0.f; jitter.y = 0.f;
    if (st->taa_enabled) {
        Vec2 hv = hammersley2D(frame_index % 8);
        jitter.x = (2.f * hv.x - 1.f) / (float)W;
        jitter.y = (2.f * hv.y - 1.f) / (float)H;
    }
    const Mat4 proj = projectionMatrix(*st);
    const Mat4 vpUnjittered = proj * viewMatrix(*cam);
    const Mat4 vpPrevUnjittered = proj * viewMatrix(prevCam ? *prevCam : *cam);
    float* depth = (float*)depthOut;
    int16_t* motion = (int16_t*)motionOut;
    uint8_t* normal = (uint8_t*)normalOut;
    uint32_t* gb = (uint32_t*)gbufferOut;
    std::vector<float> rowMin(H, 1.f), rowMax(H, 0.f);
    parallelRows(threads, H, [&](int y) {
        for (int x = 0; x < W; x++) {
            const size_t i = (size_t)y * W + x;
            const float nx = ((float)x + 0.5f) / (float)W * 2.f - 1.f + jitter.x, ny = ((float)y + 0.5f) / (float)H * 2.f - 1.f + jitter.y;
            const Vec3 dir = normalize(fwd - up * (tanHalf * ny) + right * (tanHalf * aspect * nx));
            Hit hit;
            float d = 0.f;
            uint32_t texel[4] = {0, 0, 0, 0};
            int16_t mv[2] = {0, 0};
            uint8_t nrm[4] = {0, 0, 0, 0};
            if (castRay(*s, camPos, dir, hit)) {
                const Vec3 p = camPos + dir * hit.tEnter;
                const Vec4 clip = vpUnjittered * Vec4(p, 1.f);
                d = clip.z / clip.w;
                if (d <= 0.f || d > 1.f) d = 0.f;
            }
            if (d > 0.f) {
                const Vec3 p = camPos + dir * hit.tEnter;
                const Vec4 c0 = vpUnjittered * Vec4(p, 1.f), c1 = vpPrevUnjittered * Vec4(p, 1.f);
                mv[0] = toSnorm16((c1.x / c1.w - c0.x / c0.w) * 0.5f);
                mv[1] = toSnorm16((c1.y / c1.w - c0.y / c0.w) * 0.5f);
                const Vec3 n = hit.normalWorld;
                nrm[0] = toUnorm8(n.x * 0.5f + 0.5f); nrm[1] = toUnorm8(n.y * 0.5f + 0.5f); nrm[2] = toUnorm8(n.z * 0.5f + 0.5f);
                const SynthMesh& m = s->meshes[s->objects[hit.object].mesh];
                // normal-mapped shading normal: geometric normal perturbed by a 0.25 m world-space cell hash
                const uint32_t cell = hash2(hash2((uint32_t)(int)std::floor(p.x * 4.f), (uint32_t)(int)std::floor(p.y * 4.f)), (uint32_t)(int)std::floor(p.z * 4.f));
                Vec3 pert(((cell & 0xff) / 255.f - 0.5f) * 0.16f, (((cell >> 8) & 0xff) / 255.f - 0.5f) * 0.16f, (((cell >> 16) & 0xff) / 255.f - 0.5f) * 0.16f);
                const Vec3 ns = normalize(n + pert);
                const float shade = 0.85f + 0.15f * ((cell >> 24) / 255.f);
                texel[0] = dm::f2u(d);
                texel[1] = octEncode(ns);
                texel[2] = (uint32_t)(uint8_t)(m.albedo[0] * shade) | ((uint32_t)(uint8_t)(m.albedo[1] * shade) << 8) | ((uint32_t)(uint8_t)(m.albedo[2] * shade) << 16) | ((uint32_t)m.roughness << 24);
                texel[3] = m.metal;
                if (d < rowMin[y]) rowMin[y] = d;
                if (d > rowMax[y]) rowMax[y] = d;
            }
            if (depth) depth[i] = d;
            if (motion) { motion[i * 2] = mv[0]; motion[i * 2 + 1] = mv[1]; }
            if (normal) std::memcpy(normal + i * 4, nrm, 4);
            if (gb) std::memcpy(gb + i * 4, texel, 16);
        }
    });
    bool anyShadow = false;
    for (int cidx = 0; cidx < 4; cidx++) anyShadow = anyShadow || (shadowMaps && shadowMaps[cidx]);
    if (!anyShadow) return 0;

    // light matrices with lightMatrix.comp:57-138's formulas from the depth range of this frame
    float dMin = 1.f, dMax = 0.f;
    for (int y = 0; y < H; y++) { if (rowMin[y] < dMin) dMin = rowMin[y]; if (rowMax[y] > dMax) dMax = rowMax[y]; }
    auto linearize = [&](float dd) { return st->camera_near * st->camera_far / (st->camera_far + (-dd + 1.f) * (st->camera_near - st->camera_far)); };
    const float depthMaxLinear = linearize(dMin), depthMinLinear = linearize(dMax);
    const int cascades = st->sun_shadow_cascade_count;
    const Vec2 sunDeg = {st->sun_direction_deg[0], st->sun_direction_deg[1]};
    const Vec3 sunDir = directionToVector(sunDeg);
    const Vec3 lf = -sunDir;
    Vec3 lup = std::fabs(lf.y) < 0.9999f ? Vec3(0.f, -1.f, 0.f) : Vec3(0.f, 0.f, -1.f);
    const Vec3 lright = cross(lf, lup);
    lup = cross(lright, lf);
    Mat4 V = Mat4::identity();
    const Vec3 rn = normalize(lright), un = normalize(lup);
    V.at(0, 0) = rn.x; V.at(0, 1) = rn.y; V.at(0, 2) = rn.z;
    V.at(1, 0) = un.x; V.at(1, 1) = un.y; V.at(1, 2) = un.z;
    V.at(2, 0) = lf.x; V.at(2, 1) = lf.y; V.at(2, 2) = lf.z;
    V = transpose(V);
    float splits[4] = {0, 0, 0, 0};
    for (int i = 0; i + 1 < cascades; i++) splits[i] = depthMinLinear + ((depthMaxLinear - depthMinLinear) * (float)(i + 1) / (float)cascades);
    const float padding = st->strict_influence_radius_cutoff ? st->trace_influence_radius : st->trace_influence_radius + 3.f;
    for (int ci = 0; ci < cascades; ci++) {
        if (!shadowMaps[ci]) continue;
        float nearD = ci == 0 ? depthMinLinear : splits[ci - 1], farD = splits[ci];
        if (ci == cascades - 1) { nearD = st->camera_near; farD = depthMaxLinear > 30.f ? depthMaxLinear : 30.f; }
        Vec3 mn(3.402823466e+38f), mx(1.175494351e-38f);
        for (int k = 0; k < 8; k++) {
            const float dist = (k & 4) ? nearD : farD, hh = tanHalf * dist, ww = hh * aspect;
            const Vec3 p = camPos + fwd * dist + up * ((k & 2) ? -hh : hh) + right * ((k & 1) ? -ww : ww);
            const Vec4 t = V * Vec4(p, 1.f);
            mn = vmin(mn, Vec3(t.x, t.y, t.z));
            mx = vmax(mx, Vec3(t.x, t.y, t.z));
        }
        if (ci == cascades - 1) { mn = mn - Vec3(padding); mx = mx + Vec3(padding); }
        mn = mn - Vec3(0.06f);
        mx = mx + Vec3(0.06f);
        const Vec3 sc(2.f / (mx.x - mn.x), 2.f / (mx.y - mn.y), 2.f / (mx.z - mn.z));
        const Vec3 off = (mx + mn) * sc * -0.5f;
        Mat4 P = Mat4::zero();
        P.at(0, 0) = sc.x; P.at(1, 1) = sc.y; P.at(2, 2) = sc.z; P.at(3, 0) = off.x; P.at(3, 1) = off.y; P.at(3, 2) = off.z; P.at(3, 3) = 1.f;
        Mat4 corr = Mat4::identity();
        corr.at(2, 2) = -0.5f; corr.at(3, 2) = 0.5f;
        const Mat4 LM = corr * P * V, invLM = inverse(LM);
        uint16_t* sm = (uint16_t*)shadowMaps[ci];
        const int R = 2048;
        parallelRows(threads, R, [&](int y) {
            for (int x = 0; x < R; x++) {
                const float lx = ((float)x + 0.5f) / (float)R * 2.f - 1.f, ly = ((float)y + 0.5f) / (float)R * 2.f - 1.f;
                const Vec4 a = invLM * Vec4(lx, ly, 1.f, 1.f), b = invLM * Vec4(lx, ly, 0.f, 1.f);
                const Vec3 pa(a.x / a.w, a.y / a.w, a.z / a.w), pb(b.x / b.w, b.y / b.w, b.z / b.w);
                const Vec3 d = normalize(pb - pa);
                Hit hit;
                uint16_t v = 0;
                if (castRay(*s, pa - d * 1000.f, d, hit)) {
                    const Vec3 p = pa - d * 1000.f + d * hit.tExit;  // back face, as the front-face-culled shadow pass
                    const Vec4 c = LM * Vec4(p, 1.f);
                    float z = c.z / c.w;
                    z = z < 0.f ? 0.f : (z > 1.f ? 1.f : z);
                    v = (uint16_t)(z * 65535.f + 0.5f);
                }
                sm[(size_t)y * R + x] = v;
            }
        });
    }
    return 0;
}

}  // extern "C"
