// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// glsl.h - the GLSL built-ins the reference shaders use, as scalar C++. Operation order is the numeric
// contract of DESIGN.md section "Numeric contract" (version 2): every built-in lowers to the IEEE binary32
// operation sequence written here (compiled with -ffp-contract=off: nothing contracts implicitly):
//   * dot, matrix * vector, mix and the sampler's blends are fma chains (fma_ = one correctly rounded fmaf)
//   * a division with a vector operand multiplies by the correctly rounded reciprocal of the divisor
//     (v / s = v * (1/s), v / w = v * (1/w) per component, s / v = s * (1/v)); float / float is IEEE division
//   * transcendental functions come from plainrenderer_b200/csrc/detmath.h (the pinned libm).
#pragma once
#include <stdint.h>
#include "detmath.h"

namespace gl {

typedef uint32_t uint;

struct vec3;
struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} explicit vec2(float a) : x(a), y(a) {}
    vec2 xy() const { return *this; } inline vec3 xyx() const; };
struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(vec2 v, float c) : x(v.x), y(v.y), z(c) {}
    float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; }
    vec2 xy() const { return vec2(x, y); } vec3 xyz() const { return *this; } };
inline vec3 vec2::xyx() const { return vec3(x, y, x); }
struct vec4 { float x, y, z, w; vec4() : x(0), y(0), z(0), w(0) {} vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {} vec4(vec2 v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; }
    vec3 xyz() const { return vec3(x, y, z); } vec2 xy() const { return vec2(x, y); } };
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(int a) : x(a), y(a) {} };
struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {} };
struct uvec3 { uint x, y, z; uvec3() : x(0), y(0), z(0) {} uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {} };

// reciprocal of a divisor: the correctly rounded 1/x. -DPLAIN_EMULATE_APPROX_RCP turns it into rcp.approx in the error-model build of the
// "fast" contract (liboracle_sfu.so) - the experiment that showed why the fast contract must NOT do that: sdfDiffuseTrace.comp:120 and
// sdfCameraTileCulling.comp:75 sample with NEAREST filtering at uv = iUV / size, exactly on texel borders, so one ulp in the reciprocal
// moves a quarter of the pixels to the neighbouring texel (tests/test_fast_contract_emulation.py documents the numbers)
inline float rcp_(float x) {
#if defined(DM_FAST) && defined(PLAIN_EMULATE_APPROX_RCP)
    return dm::hw_rcp(x);
#else
    return 1.f / x;
#endif
}
#define GL_VEC_OPS(V, N)                                                                                                \
    inline V operator+(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] + (&b.x)[i]; return r; }      \
    inline V operator-(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] - (&b.x)[i]; return r; }      \
    inline V operator*(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * (&b.x)[i]; return r; }      \
    inline V operator/(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * rcp_((&b.x)[i]); return r; } \
    inline V operator+(V a, float b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] + b; return r; }          \
    inline V operator-(V a, float b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] - b; return r; }          \
    inline V operator*(V a, float b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * b; return r; }          \
    inline V operator/(V a, float b) { V r; const float rb = rcp_(b); for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * rb; return r; } \
    inline V operator+(float a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = a + (&b.x)[i]; return r; }          \
    inline V operator-(float a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = a - (&b.x)[i]; return r; }          \
    inline V operator*(float a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = a * (&b.x)[i]; return r; }          \
    inline V operator/(float a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = a * rcp_((&b.x)[i]); return r; }  \
    inline V operator-(V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = -(&a.x)[i]; return r; }                      \
    inline V& operator+=(V& a, V b) { a = a + b; return a; }                                                            \
    inline V& operator-=(V& a, V b) { a = a - b; return a; }                                                            \
    inline V& operator*=(V& a, V b) { a = a * b; return a; }                                                            \
    inline V& operator/=(V& a, V b) { a = a / b; return a; }                                                            \
    inline V& operator+=(V& a, float b) { a = a + b; return a; }                                                        \
    inline V& operator-=(V& a, float b) { a = a - b; return a; }                                                        \
    inline V& operator*=(V& a, float b) { a = a * b; return a; }                                                        \
    inline V& operator/=(V& a, float b) { a = a / b; return a; }
GL_VEC_OPS(vec2, 2)
GL_VEC_OPS(vec3, 3)
GL_VEC_OPS(vec4, 4)
// fused multiply-add: one rounding. vfma(a, s, c) = a * s + c per component
inline float fma_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline float vfma(float a, float s, float c) { return fma_(a, s, c); }
inline vec2 vfma(vec2 a, float s, vec2 c) { return vec2(fma_(a.x, s, c.x), fma_(a.y, s, c.y)); }
inline vec3 vfma(vec3 a, float s, vec3 c) { return vec3(fma_(a.x, s, c.x), fma_(a.y, s, c.y), fma_(a.z, s, c.z)); }
inline vec4 vfma(vec4 a, float s, vec4 c) { return vec4(fma_(a.x, s, c.x), fma_(a.y, s, c.y), fma_(a.z, s, c.z), fma_(a.w, s, c.w)); }

inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator*(ivec2 a, int b) { return ivec2(a.x * b, a.y * b); }
inline ivec2 operator/(ivec2 a, int b) { return ivec2(a.x / b, a.y / b); }
inline vec2 tovec2(ivec2 a) { return vec2((float)a.x, (float)a.y); }

// scalar built-ins. min(x,y) = y < x ? y : x ; max(x,y) = x < y ? y : x, except that a NaN operand is dropped
// (GLSL leaves NaN undefined; GPUs return the other operand, and the frame path relies on it: preExposeLights.comp:72
// max(targetEV100, 10) must recover from the NaN mean of an empty histogram)
inline float min(float x, float y) { return dm::isnan_(x) ? y : (dm::isnan_(y) ? x : ((y < x) ? y : x)); }
inline float max(float x, float y) { return dm::isnan_(x) ? y : (dm::isnan_(y) ? x : ((x < y) ? y : x)); }
// float -> int / uint conversion, saturating, NaN -> 0 (GLSL: undefined out of range; pinned to what cvt.rzi does)
inline int f2int(float f) { if (dm::isnan_(f)) return 0; if (f >= 2147483648.f) return 2147483647; if (f <= -2147483648.f) return (int)0x80000000; return (int)f; }
inline uint32_t f2uint(float f) { if (dm::isnan_(f)) return 0u; if (f >= 4294967296.f) return 0xffffffffu; if (f <= 0.f) return 0u; return (uint32_t)f; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline int max(int x, int y) { return (x < y) ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float abs(float x) { return dm::abs_(x); }
inline float sign(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
inline float floor(float x) { return dm::floor_(x); }
inline float fract(float x) { return x - dm::floor_(x); }
inline float sqrt(float x) { return dm::sqrt_(x); }
inline float mix(float a, float b, float t) { return fma_(b, t, a * (1.f - t)); }
inline float exp(float x) { return dm::exp(x); }
inline float exp2(float x) { return dm::exp2(x); }
inline float log(float x) { return dm::log(x); }
inline float log2(float x) { return dm::log2(x); }
inline float pow(float x, float y) { return dm::pow(x, y); }
inline float sin(float x) { return dm::sin(x); }
inline float cos(float x) { return dm::cos(x); }
inline float acos(float x) { return dm::acos(x); }
inline float atan(float y, float x) { return dm::atan2(y, x); }
inline bool isnan(float x) { return dm::isnan_(x); }
inline float uintBitsToFloat(uint u) { return dm::u2f(u); }

#define GL_MAP1(V, N, fn) inline V fn(V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fn((&a.x)[i]); return r; }
#define GL_MAP2(V, N, fn) inline V fn(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fn((&a.x)[i], (&b.x)[i]); return r; }
#define GL_MAP2S(V, N, fn) inline V fn(V a, float b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fn((&a.x)[i], b); return r; }
GL_MAP1(vec2, 2, abs) GL_MAP1(vec3, 3, abs) GL_MAP1(vec4, 4, abs)
GL_MAP1(vec2, 2, floor) GL_MAP1(vec3, 3, exp) GL_MAP1(vec2, 2, sqrt)
GL_MAP2(vec2, 2, min) GL_MAP2(vec3, 3, min) GL_MAP2(vec4, 4, min)
GL_MAP2(vec2, 2, max) GL_MAP2(vec3, 3, max) GL_MAP2(vec4, 4, max)
GL_MAP2S(vec2, 2, min) GL_MAP2S(vec3, 3, min) GL_MAP2S(vec2, 2, max) GL_MAP2S(vec3, 3, max)
GL_MAP2(vec3, 3, pow)
inline vec3 clamp(vec3 x, float lo, float hi) { return vec3(clamp(x.x, lo, hi), clamp(x.y, lo, hi), clamp(x.z, lo, hi)); }
inline vec3 clamp(vec3 x, vec3 lo, vec3 hi) { return vec3(clamp(x.x, lo.x, hi.x), clamp(x.y, lo.y, hi.y), clamp(x.z, lo.z, hi.z)); }
inline vec2 clamp(vec2 x, float lo, float hi) { return vec2(clamp(x.x, lo, hi), clamp(x.y, lo, hi)); }
inline vec3 mix(vec3 a, vec3 b, float t) { return vfma(b, t, a * (1.f - t)); }
inline vec4 mix(vec4 a, vec4 b, float t) { return vfma(b, t, a * (1.f - t)); }
inline vec2 mix(vec2 a, vec2 b, float t) { return vfma(b, t, a * (1.f - t)); }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return vec3(mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z)); }

// dot = fma(a.w,b.w, fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))), left to right
inline float dot(vec2 a, vec2 b) { return fma_(a.y, b.y, a.x * b.x); }
inline float dot(vec3 a, vec3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
inline float dot(vec4 a, vec4 b) { return fma_(a.w, b.w, fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x))); }
inline float length(vec2 a) { return sqrt(dot(a, a)); }
inline float length(vec3 a) { return sqrt(dot(a, a)); }
inline float length(vec4 a) { return sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
// normalize(v) = v / length(v) = v * (1 / sqrt(dot(v, v)))
#if defined(DM_FAST)
inline vec3 normalize(vec3 a) { return a * dm::hw_rsqrt(dot(a, a)); }
inline vec4 normalize(vec4 a) { return a * dm::hw_rsqrt(dot(a, a)); }
#else
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec4 normalize(vec4 a) { return a / length(a); }
#endif
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// reflect(I, N) = I - (2 * dot(N, I)) * N
inline vec3 reflect(vec3 I, vec3 N) { return I - (2.f * dot(N, I)) * N; }
inline bool any_isnan(vec3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
inline bool any_isnan(vec4 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z) || isnan(a.w); }
inline bool any_isnan(vec2 a) { return isnan(a.x) || isnan(a.y); }

// column-major 4x4: c[col][row], like GLSL/glm
struct mat4 {
    vec4 c[4];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat4 mat4_diag(float d) { mat4 m; m.c[0] = vec4(d, 0, 0, 0); m.c[1] = vec4(0, d, 0, 0); m.c[2] = vec4(0, 0, d, 0); m.c[3] = vec4(0, 0, 0, d); return m; }
// M * v = fma(c3, v.w, fma(c2, v.z, fma(c1, v.y, c0*v.x))) per row
inline vec4 operator*(const mat4& m, vec4 v) { return vfma(m.c[3], v.w, vfma(m.c[2], v.z, vfma(m.c[1], v.y, m.c[0] * v.x))); }
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for (int j = 0; j < 4; j++) r.c[j] = a * b.c[j]; return r; }
inline mat4 transpose(const mat4& m) { mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.c[i][j] = m.c[j][i]; return r; }

struct mat3 {
    vec3 c[3];
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 mat3_from(const mat4& m) { mat3 r; for (int i = 0; i < 3; i++) r.c[i] = m.c[i].xyz(); return r; }
inline vec3 operator*(const mat3& m, vec3 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline mat3 transpose(const mat3& m) { mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.c[i][j] = m.c[j][i]; return r; }

static const float pi = 3.1415926535f;  // global.inc:44

}  // namespace gl
