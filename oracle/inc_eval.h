// ORACLE - test infrastructure only. Batch evaluation entry points over the functions of the reference's GLSL include files, compiled
// twice from this one header:
//   * oracle/ref/ref_glsl_shim.cpp  (INC_IS_REFERENCE 1): namespace refglsl = the reference's OWN text (oracle/_ref/libref_glsl.so)
//   * oracle/inc_eval_*.cpp / passes_gi.cpp / passes_post.cpp (INC_IS_REFERENCE 0): namespace orc = the oracle's restatement (liboracle.so)
// tests/test_oracle_vs_reference_glsl.py feeds both the same random inputs and compares the outputs bit for bit. Where the two sides spell
// a call differently (a global uniform against a parameter, a GLSL struct against the C-ABI struct) the branch is in this file.
//
//   int <prefix>eval(const char* name, const float* in, int n_in, float* out, int n_out, int count)
// evaluates `name` on `count` items of n_in floats each into n_out floats each (unsigned values travel as their bit patterns); returns 0,
// 1 for an unknown name or a wrong arity. Images for the functions that sample: <prefix>set_image(slot, format, w, h, d, data, bytes).
#include <string.h>
#include <stdint.h>

#ifndef INC_CAT
#define INC_CAT2(a, b) a##b
#define INC_CAT(a, b) INC_CAT2(a, b)
#endif
#define INC_FN(name) INC_CAT(INC_PREFIX, name)

namespace INC_NS {
namespace inc_eval {
using gl::vec2; using gl::vec3; using gl::vec4;

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline vec3 V3(const float* p) { return vec3(p[0], p[1], p[2]); }
static inline vec2 V2(const float* p) { return vec2(p[0], p[1]); }
static inline void put(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static inline void put(float* o, vec2 v) { o[0] = v.x; o[1] = v.y; }
static inline void put(float* o, vec4 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

#if defined(INC_PART_PURE)
#define CASE(fname, NI, NO) if (!strcmp(name, fname)) { if (n_in != NI || n_out != NO) return 1; for (int i = 0; i < count; i++) { const float* a = in + (size_t)i * NI; float* o = out + (size_t)i * NO;
#define END } return 0; }
static int evalPure(const char* name, const float* in, int n_in, float* out, int n_out, int count) {
    CASE("D_GGX", 2, 1) o[0] = D_GGX(a[0], a[1]); END
    CASE("Visibility", 3, 1) o[0] = Visibility(a[0], a[1], a[2]); END
    CASE("F_Schlick", 7, 3) put(o, F_Schlick(V3(a), V3(a + 3), a[6])); END
    CASE("DisneyDiffuse", 7, 3) put(o, DisneyDiffuse(V3(a), a[3], a[4], a[5], a[6])); END
    CASE("CoDWWIIDiffuse", 8, 3) put(o, CoDWWIIDiffuse(V3(a), a[3], a[4], a[5], a[6], a[7])); END
    CASE("Titanfall2DiffuseSingleComponent", 5, 1) o[0] = Titanfall2DiffuseSingleComponent(a[0], a[1], a[2], a[3], a[4]); END
    CASE("Titanfall2Diffuse", 8, 3) put(o, Titanfall2Diffuse(V3(a), a[3], a[4], a[5], a[6], a[7])); END
    CASE("GGXSingleScattering", 8, 3) put(o, GGXSingleScattering(a[0], V3(a + 1), a[4], a[5], a[6], a[7])); END
    CASE("RRTAndODTFit", 3, 3) put(o, RRTAndODTFit(V3(a))); END
    CASE("ACESFitted", 3, 3) put(o, ACESFitted(V3(a))); END
    CASE("linearTosRGB", 3, 3) put(o, linearTosRGB(V3(a))); END
    CASE("sRGBToLinear", 3, 3) put(o, sRGBToLinear(V3(a))); END
    CASE("linearToYCoCg", 3, 3) put(o, linearToYCoCg(V3(a))); END
    CASE("YCoCgToLinear", 3, 3) put(o, YCoCgToLinear(V3(a))); END
    CASE("directionToSH_L1", 3, 4) put(o, directionToSH_L1(V3(a))); END
    CASE("dominantDirectionFromSH_L1", 4, 3) put(o, dominantDirectionFromSH_L1(vec4(a[0], a[1], a[2], a[3]))); END
    CASE("importanceSampleGGX", 6, 3) put(o, importanceSampleGGX(V2(a), a[2], V3(a + 3))); END
    CASE("importanceSampleCosine", 5, 3) put(o, importanceSampleCosine(V2(a), V3(a + 2))); END
    CASE("radicalInverse_VdC", 1, 1) o[0] = radicalInverse_VdC(f2u(a[0])); END
    CASE("hammersley2d", 2, 2) put(o, hammersley2d(f2u(a[0]), f2u(a[1]))); END
    CASE("hash32", 2, 3) put(o, hash32(V2(a))); END
    CASE("wang_hash", 1, 1) o[0] = u2f(wang_hash(f2u(a[0]))); END
    CASE("xorshift32", 1, 2) { gl::uint st = f2u(a[0]); const gl::uint r = xorshift32(st); o[0] = u2f(r); o[1] = u2f(st); } END
    CASE("rand", 1, 2) { gl::uint st = f2u(a[0]); o[0] = rand(st); o[1] = u2f(st); } END
#if INC_IS_REFERENCE
    CASE("ditherRGB8", 6, 3) g_time = a[5]; put(o, ditherRGB8(V3(a), gl::ivec2((int)a[3], (int)a[4]))); END
#else
    CASE("ditherRGB8", 6, 3) put(o, ditherRGB8(V3(a), gl::ivec2((int)a[3], (int)a[4]), a[5])); END
#endif
    CASE("computeLuminance", 3, 1) o[0] = computeLuminance(V3(a)); END
    CASE("linearizeDepth", 3, 1) o[0] = linearizeDepth(a[0], a[1], a[2]); END
    CASE("calculateViewDirectionFromPixel", 13, 3) put(o, calculateViewDirectionFromPixel(V2(a), V3(a + 2), V3(a + 5), V3(a + 8), a[11], a[12])); END
    CASE("phaseGreenstein", 2, 1) o[0] = phaseGreenstein(a[0], a[1]); END
    CASE("phaseRayleigh", 1, 1) o[0] = phaseRayleigh(a[0]); END
    CASE("cornetteShanksPhase", 2, 1) o[0] = cornetteShanksPhase(a[0], a[1]); END
    CASE("integrateInscattering", 7, 3) put(o, integrateInscattering(V3(a), V3(a + 3), a[6])); END
    CASE("calculateCoefficients", 15, 9) {
        // a[1..14]: scatteringRayleighGround(3) earthRadius extinctionRayleighGround(3) atmosphereHeight ozoneExtinction(3) scatteringMieGround extinctionMieGround mieScatteringExponent
#if INC_IS_REFERENCE
        AtmosphereSettings s;
        s.scatteringRayleighGround = V3(a + 1); s.earthRadius = a[4]; s.extinctionRayleighGround = V3(a + 5); s.atmosphereHeight = a[8];
        s.ozoneExtinction = V3(a + 9); s.scatteringMieGround = a[12]; s.extinctionMieGround = a[13]; s.mieScatteringExponent = a[14];
#else
        plain_atmosphere_settings s;
        memset(&s, 0, sizeof(s));
        for (int k = 0; k < 3; k++) { s.scatteringRayleighGround[k] = a[1 + k]; s.extinctionRayleighGround[k] = a[5 + k]; s.ozoneExtinction[k] = a[9 + k]; }
        s.earthRadius = a[4]; s.atmosphereHeight = a[8]; s.scatteringMieGround = a[12]; s.extinctionMieGround = a[13]; s.mieScatteringExponent = a[14];
#endif
        const AtmosphereCoefficients c = calculateCoefficients(a[0], s);
        put(o, c.scatterRayleigh); put(o + 3, c.scatterMie); put(o + 6, c.extinction);
    } END
    CASE("rayEarthIntersection", 11, 5) { const Intersection r = rayEarthIntersection(V3(a), V3(a + 3), V3(a + 6), a[9], a[10]); put(o, r.pos); o[3] = r.distance; o[4] = r.hitEarth ? 1.f : 0.f; } END
    CASE("toSkyLut", 3, 2) put(o, toSkyLut(V3(a))); END
    CASE("fromSkyLut", 2, 3) put(o, fromSkyLut(V2(a))); END
    CASE("computeLutUV", 8, 2) put(o, computeLutUV(a[0], a[1], V3(a + 2), V3(a + 5))); END
    return 1;
}
#undef CASE
#undef END
#endif  // INC_PART_PURE

// ---- images for the sampling functions ----
#if defined(INC_PART_SDF) || defined(INC_PART_TAA)
static orc::Image gImages[4];
static int setImage(int slot, uint32_t format, int w, int h, int d, const void* data, size_t bytes) {
    if (slot < 0 || slot >= 4) return 1;
    plain_image_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.width = (uint32_t)w; desc.height = (uint32_t)h; desc.depth = (uint32_t)d;
    desc.type = d > 1 ? PLAIN_IMAGE_TYPE_3D : PLAIN_IMAGE_TYPE_2D;
    desc.format = format;
    desc.mip_count = PLAIN_MIPS_ONE;
    gImages[slot].allocate(desc);
    if (bytes != gImages[slot].mips[0].data.size()) return 1;
    memcpy(gImages[slot].mips[0].data.data(), data, bytes);
    return 0;
}
static orc::View viewOf(int slot) { orc::View v; v.img = &gImages[slot]; v.mip = 0; return v; }
#endif

#if defined(INC_PART_SDF)
// item: instance {localExtends(3), meanAlbedo(3), worldToLocal(16 column-major)} = 22 floats, rayStart(3), rayDirection(3), closestHitDistance(1) = 29 in
// out: hit, closestHitDistance, hitPos(3), N(3), hitCount, albedo(3) = 12 (TraceResult after ONE instance, starting from {false, closest, 0, 0, 0, 0})
static int evalSdf(const char* name, const float* in, int n_in, float* out, int n_out, int count) {
    if (!strcmp(name, "isPointInAABB")) {
        if (n_in != 9 || n_out != 1) return 1;
        for (int i = 0; i < count; i++) { const float* a = in + (size_t)i * 9; out[i] = isPointInAABB(V3(a), V3(a + 3), V3(a + 6)) ? 1.f : 0.f; }
        return 0;
    }
    if (!strcmp(name, "rayAABBIntersection")) {
        if (n_in != 12 || n_out != 2) return 1;
        for (int i = 0; i < count; i++) { const float* a = in + (size_t)i * 12; const HitResult r = rayAABBIntersection(V3(a), V3(a + 3), V3(a + 6), V3(a + 9)); out[2 * i] = r.hit ? 1.f : 0.f; out[2 * i + 1] = r.t; }
        return 0;
    }
    if (!strcmp(name, "normalFromSDF")) {  // uv(3), extends(3) -> N(3), brick in image slot 0
        if (n_in != 6 || n_out != 3) return 1;
        const orc::View sdf = viewOf(0);
        for (int i = 0; i < count; i++) {
            const float* a = in + (size_t)i * 6;
#if INC_IS_REFERENCE
            put(out + 3 * i, normalFromSDF(V3(a), V3(a + 3), &sdf));
#else
            put(out + 3 * i, normalFromSDF(V3(a), V3(a + 3), sdf));
#endif
        }
        return 0;
    }
    if (!strcmp(name, "traceRayTroughSDFInstance")) {
        if (n_in != 29 || n_out != 12) return 1;
        const orc::View sdf = viewOf(0);
        for (int i = 0; i < count; i++) {
            const float* a = in + (size_t)i * 29;
            float* o = out + (size_t)i * 12;
            gl::mat4 m;
            for (int c = 0; c < 4; c++) m.c[c] = vec4(a[6 + c * 4], a[7 + c * 4], a[8 + c * 4], a[9 + c * 4]);
            TraceResult tr;
            tr.hit = false; tr.closestHitDistance = a[28]; tr.hitPos = vec3(0.f); tr.N = vec3(0.f); tr.hitCount = 0; tr.albedo = vec3(0.f);
#if INC_IS_REFERENCE
            SDFInstance inst;
            inst.localExtends = V3(a); inst.sdfTextureIndex = 0; inst.meanAlbedo = V3(a + 3); inst.padding = 0.f; inst.worldToLocal = m;
            traceRayTroughSDFInstance(inst, V3(a + 22), &sdf, V3(a + 25), tr);
#else
            plain_sdf_instance inst;
            memset(&inst, 0, sizeof(inst));
            for (int k = 0; k < 3; k++) { inst.localExtends[k] = a[k]; inst.meanAlbedo[k] = a[3 + k]; }
            for (int k = 0; k < 16; k++) inst.worldToLocal[k] = a[6 + k];
            traceRayTroughSDFInstance(inst, m, V3(a + 22), sdf, V3(a + 25), tr);
#endif
            o[0] = tr.hit ? 1.f : 0.f; o[1] = tr.closestHitDistance; put(o + 2, tr.hitPos); put(o + 5, tr.N); o[8] = (float)tr.hitCount; put(o + 9, tr.albedo);
        }
        return 0;
    }
    return 1;
}
#endif  // INC_PART_SDF

#if defined(INC_PART_TAA)
static int evalTaa(const char* name, const float* in, int n_in, float* out, int n_out, int count) {
    if (!strcmp(name, "catmullRomWeight1D")) {
        if (n_in != 1 || n_out != 1) return 1;
        for (int i = 0; i < count; i++) out[i] = catmullRomWeight1D(in[i]);
        return 0;
    }
    if (!strcmp(name, "clipAABB")) {
        if (n_in != 9 || n_out != 3) return 1;
        for (int i = 0; i < count; i++) { const float* a = in + (size_t)i * 9; put(out + 3 * i, clipAABB(V3(a), V3(a + 3), V3(a + 6))); }
        return 0;
    }
    if (!strcmp(name, "tonemap") || !strcmp(name, "tonemapReverse")) {
        if (n_in != 3 || n_out != 3) return 1;
        const bool rev = name[7] == 'R';
        for (int i = 0; i < count; i++) {
#if INC_IS_REFERENCE
            put(out + 3 * i, rev ? tonemapReverse(V3(in + 3 * i)) : tonemap(V3(in + 3 * i)));
#else
            put(out + 3 * i, rev ? taaTonemapReverse(V3(in + 3 * i)) : taaTonemap(V3(in + 3 * i)));
#endif
        }
        return 0;
    }
    // sampleNeighbourhood + minMaxFromNeighbourhood of the R11G11B10 image in slot 1: uv(2), texelSize(2), useTonemapping(1) -> 27 + 6
    if (!strcmp(name, "sampleNeighbourhood")) {
        if (n_in != 5 || n_out != 33) return 1;
        const orc::View tex = viewOf(1);
        for (int i = 0; i < count; i++) {
            const float* a = in + (size_t)i * 5;
            float* o = out + (size_t)i * 33;
#if INC_IS_REFERENCE
            const Nb33 nb = sampleNeighbourhood(&tex, &orc::s_linearClamp, V2(a), V2(a + 2), a[4] != 0.f);
            const Vec3x2 mm = minMaxFromNeighbourhood(nb);
            for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) put(o + (x * 3 + y) * 3, nb[x][y]);
            put(o + 27, mm[0]); put(o + 30, mm[1]);
#else
            const Nb nb = sampleNeighbourhood(tex, V2(a), V2(a + 2), a[4] != 0.f);
            vec3 mn = nb.v[0][0], mx = nb.v[0][0];
            for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) { mn = min(mn, nb.v[x][y]); mx = max(mx, nb.v[x][y]); put(o + (x * 3 + y) * 3, nb.v[x][y]); }
            put(o + 27, mn); put(o + 30, mx);
#endif
        }
        return 0;
    }
    // history sample of temporalFilter.comp:104-127 on the R11G11B10 image in slot 1, neighbourhood sampled from slot 2 at uv:
    // tech(1), uv(2), motion(2), screenResolution(2) -> 3. iUV = uv * screenResolution - 0.5 is passed as the pixel index (2 more floats)
    if (!strcmp(name, "historySample")) {
        if (n_in != 9 || n_out != 3) return 1;
        const orc::View hist = viewOf(1), cur = viewOf(2);
        for (int i = 0; i < count; i++) {
            const float* a = in + (size_t)i * 9;
            const int tech = (int)a[0];
            const vec2 iUVf(a[1], a[2]), motion(a[3], a[4]), screenRes(a[5], a[6]), texelSize(a[7], a[8]);
            const vec2 uv = (iUVf + 0.5f) * texelSize;
            vec3 r;
#if INC_IS_REFERENCE
            const Nb33 nb = sampleNeighbourhood(&cur, &orc::s_linearClamp, uv, texelSize, true);
            if (tech == 0) r = texture(sampler2D(&hist, g_sampler_linearClamp), uv + motion).xyz();
            else {
                const vec2 uvReprojected = iUVf + 0.5f + motion * screenRes;
                if (tech == 1) r = bicubicSample16Tap(&hist, g_sampler_linearClamp, uvReprojected, texelSize);
                else if (tech == 2) r = bicubicSample9Tap(&hist, g_sampler_linearClamp, uvReprojected, texelSize);
                else if (tech == 3) r = bicubicSample5Tap(&hist, g_sampler_linearClamp, uvReprojected, texelSize);
                else r = bicubicSample1Tap(&hist, g_sampler_linearClamp, uvReprojected, texelSize, nb);
            }
#else
            const Nb nb = sampleNeighbourhood(cur, uv, texelSize, true);
            r = sampleHistory(tech, hist, uv, motion, iUVf + 0.5f + motion * screenRes, texelSize, nb);
#endif
            put(out + 3 * i, r);
        }
        return 0;
    }
    return 1;
}
#endif  // INC_PART_TAA

}  // namespace inc_eval
}  // namespace INC_NS

extern "C" {
#if defined(INC_PART_PURE)
__attribute__((visibility("default"))) int INC_FN(eval_pure)(const char* name, const float* in, int n_in, float* out, int n_out, int count) { return INC_NS::inc_eval::evalPure(name, in, n_in, out, n_out, count); }
#endif
#if defined(INC_PART_SDF)
__attribute__((visibility("default"))) int INC_FN(eval_sdf)(const char* name, const float* in, int n_in, float* out, int n_out, int count) { return INC_NS::inc_eval::evalSdf(name, in, n_in, out, n_out, count); }
__attribute__((visibility("default"))) int INC_FN(set_image_sdf)(int slot, uint32_t format, int w, int h, int d, const void* data, size_t bytes) { return INC_NS::inc_eval::setImage(slot, format, w, h, d, data, bytes); }
#endif
#if defined(INC_PART_TAA)
__attribute__((visibility("default"))) int INC_FN(eval_taa)(const char* name, const float* in, int n_in, float* out, int n_out, int count) { return INC_NS::inc_eval::evalTaa(name, in, n_in, out, n_out, count); }
__attribute__((visibility("default"))) int INC_FN(set_image_taa)(int slot, uint32_t format, int w, int h, int d, const void* data, size_t bytes) { return INC_NS::inc_eval::setImage(slot, format, w, h, d, data, bytes); }
#endif
}
