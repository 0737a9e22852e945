// ORACLE - test infrastructure only. What the reference's GLSL include files need beyond oracle/glsl.h to compile as C++
// (oracle/ref/ref_glsl_shim.cpp): integer vectors, mixed int / float operands (GLSL converts implicitly), mat3 from columns /
// from a mat4, bool vectors, array-typed values, and the texture types bound to the oracle's own image sampler
// (oracle/image.h) - the stand-in for the Vulkan driver's sampler on both sides of the comparison.
// Every arithmetic operation is oracle/glsl.h's (the numeric contract): this header only adds spelling.
#pragma once
#include "glsl.h"
#include "image.h"

namespace refglsl {
using namespace gl;
using orc::View;
using orc::Sampler;

// ---- integer / unsigned vectors (float -> int conversions saturate like the contract's f2int / f2uint) ----
struct uvec2; struct uvec3;
struct ivec2 : gl::ivec2 { using gl::ivec2::ivec2; ivec2() {} ivec2(gl::ivec2 v) : gl::ivec2(v) {} explicit ivec2(gl::vec2 v) : gl::ivec2(f2int(v.x), f2int(v.y)) {} explicit inline ivec2(const uvec2& v); ivec2 xy() const { return *this; } };
struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {} explicit ivec3(gl::vec3 v) : x(f2int(v.x)), y(f2int(v.y)), z(f2int(v.z)) {}
    ivec3(ivec2 v, int c) : x(v.x), y(v.y), z(c) {} explicit inline ivec3(const uvec3& v); inline ivec3(const uvec2& v, int c); };
struct ivec4 { int x, y, z, w; ivec4() : x(0), y(0), z(0), w(0) {} ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {} int operator[](int i) const { return (&x)[i]; } };
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} explicit uvec2(gl::vec2 v) : x(f2uint(v.x)), y(f2uint(v.y)) {} uvec2 xy() const { return *this; } };
struct uvec3 { uint x, y, z; uvec3() : x(0), y(0), z(0) {} uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {} uvec3(ivec3 v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}  /* uvec3 v = imageSize(image3D): GLSL converts implicitly */
    uvec2 xy() const { return uvec2(x, y); } uvec3 xyz() const { return *this; } };
inline ivec2::ivec2(const uvec2& v) : gl::ivec2((int)v.x, (int)v.y) {}
inline ivec3::ivec3(const uvec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
inline ivec3::ivec3(const uvec2& v, int c) : x((int)v.x), y((int)v.y), z(c) {}
inline uvec3 operator*(uvec3 a, uvec3 b) { return uvec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline uvec3 operator*(uint a, uvec3 b) { return uvec3(a * b.x, a * b.y, a * b.z); }
// vec3(uvec3), vec2(uvec2): GLSL constructors the converter leaves as written
struct vec3 : gl::vec3 { using gl::vec3::vec3; vec3() {} vec3(gl::vec3 v) : gl::vec3(v) {} explicit vec3(uvec3 n) : gl::vec3((float)n.x, (float)n.y, (float)n.z) {} gl::vec2 yz() const { return gl::vec2(y, z); }
    vec3(float a, gl::vec2 b) : gl::vec3(a, b.x, b.y) {} };
struct vec2 : gl::vec2 { using gl::vec2::vec2; vec2() {} vec2(gl::vec2 v) : gl::vec2(v) {} vec2(uvec2 n) : gl::vec2((float)n.x, (float)n.y) {}  /* GLSL converts uvec2 -> vec2 implicitly (dither.inc:7) */
    explicit vec2(gl::ivec2 n) : gl::vec2((float)n.x, (float)n.y) {} };
struct vec4 : gl::vec4 { using gl::vec4::vec4; vec4() {} vec4(gl::vec4 v) : gl::vec4(v) {} vec4(float a, gl::vec3 v) : gl::vec4(a, v.x, v.y, v.z) {} vec4(gl::vec2 a, gl::vec2 b) : gl::vec4(a.x, a.y, b.x, b.y) {}
    vec4(float a, float b, gl::vec2 c) : gl::vec4(a, b, c.x, c.y) {} };
inline ivec2 operator*(ivec2 a, int b) { return ivec2(a.x * b, a.y * b); }
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(ivec2 a, ivec2 b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator/(ivec2 a, int b) { return ivec2(a.x / b, a.y / b); }
inline ivec2 operator/(ivec2 a, ivec2 b) { return ivec2(a.x / b.x, a.y / b.y); }
inline ivec2 operator*(ivec2 a, ivec2 b) { return ivec2(a.x * b.x, a.y * b.y); }
inline ivec2 max(ivec2 a, ivec2 b) { return ivec2(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y); }
inline ivec2 min(ivec2 a, ivec2 b) { return ivec2(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y); }
inline gl::vec2 operator*(ivec2 a, gl::vec2 b) { return gl::vec2((float)a.x * b.x, (float)a.y * b.y); }
inline gl::vec2 operator*(ivec2 a, float b) { return gl::vec2((float)a.x * b, (float)a.y * b); }
inline gl::vec2 operator+(ivec2 a, gl::vec2 b) { return gl::vec2((float)a.x + b.x, (float)a.y + b.y); }

// ---- mixed int / float operands ----
#define REF_INT_OPS(V)                                                          \
    inline V operator+(V a, int b) { return a + (float)b; }                     \
    inline V operator-(V a, int b) { return a - (float)b; }                     \
    inline V operator*(V a, int b) { return a * (float)b; }                     \
    inline V operator/(V a, int b) { return a / (float)b; }                     \
    inline V operator+(int a, V b) { return (float)a + b; }                     \
    inline V operator-(int a, V b) { return (float)a - b; }                     \
    inline V operator*(int a, V b) { return (float)a * b; }                     \
    inline V operator/(int a, V b) { return (float)a / b; }
REF_INT_OPS(gl::vec2) REF_INT_OPS(gl::vec3) REF_INT_OPS(gl::vec4)
using gl::max; using gl::min; using gl::clamp;  // keep the float overloads visible next to the mixed ones declared here
inline float max(int a, float b) { return gl::max((float)a, b); }
inline float max(float a, int b) { return gl::max(a, (float)b); }
inline float min(int a, float b) { return gl::min((float)a, b); }
inline float min(float a, int b) { return gl::min(a, (float)b); }
inline float clamp(float x, int lo, int hi) { return gl::clamp(x, (float)lo, (float)hi); }
inline float clamp(float x, float lo, int hi) { return gl::clamp(x, lo, (float)hi); }
inline gl::vec3 clamp(gl::vec3 x, int lo, int hi) { return gl::clamp(x, (float)lo, (float)hi); }
inline gl::vec3 operator/(gl::vec3 a, ivec3 b) { return a / gl::vec3((float)b.x, (float)b.y, (float)b.z); }

// ---- matrices ----
struct mat4 : gl::mat4 {  // mat4(d), mat4 M = {{..}, {..}, {..}, {..}} (columns)
    mat4() {}
    mat4(const gl::mat4& m) : gl::mat4(m) {}
    explicit mat4(float d) : gl::mat4(gl::mat4_diag(d)) {}
    mat4(gl::vec4 c0, gl::vec4 c1, gl::vec4 c2, gl::vec4 c3) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
};
struct mat3 : gl::mat3 {
    mat3() {}
    mat3(const gl::mat3& m) : gl::mat3(m) {}
    mat3(gl::vec3 c0, gl::vec3 c1, gl::vec3 c2) { c[0] = c0; c[1] = c1; c[2] = c2; }
    explicit mat3(const gl::mat4& m) { for (int i = 0; i < 3; i++) c[i] = m.c[i].xyz(); }
};

// ---- bool vectors ----
struct bvec3 { bool x, y, z; };
inline bvec3 greaterThan(gl::vec3 a, gl::vec3 b) { return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec3 lessThan(gl::vec3 a, gl::vec3 b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bool any(bvec3 v) { return v.x || v.y || v.z; }

// ---- arrays GLSL spells as types ----
struct Nb33 { gl::vec3 v[3][3]; gl::vec3* operator[](int i) { return v[i]; } const gl::vec3* operator[](int i) const { return v[i]; } };
struct Vec3x2 { gl::vec3 v[2]; gl::vec3& operator[](int i) { return v[i]; } const gl::vec3& operator[](int i) const { return v[i]; } };

// ---- textures: the oracle's sampler ----
typedef const View* texture2D;
typedef const View* texture3D;
typedef const Sampler* sampler;
struct sampler2D { const View* t; const Sampler* s; sampler2D(const View* tex, const Sampler* smp) : t(tex), s(smp) {} };
struct sampler3D { const View* t; const Sampler* s; sampler3D(const View* tex, const Sampler* smp) : t(tex), s(smp) {} };
inline vec4 texture(sampler2D s, gl::vec2 uv) { return orc::texture(*s.t, *s.s, uv); }
inline vec4 texture(sampler3D s, gl::vec3 uvw) { return orc::texture3D(*s.t, *s.s, uvw); }
inline vec4 texelFetch(sampler2D s, ivec2 uv, int) { return s.t->fetch(uv); }
inline ivec3 textureSize(sampler3D s, int) { return ivec3(s.t->w(), s.t->h(), s.t->d()); }
static const Sampler* const g_sampler_linearClamp = &orc::s_linearClamp;
static const Sampler* const g_sampler_linearRepeat = &orc::s_linearRepeat;

}  // namespace refglsl
