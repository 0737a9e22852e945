// HostMath.h - small column-major vector/matrix helpers for the host side (the reference uses glm here).
// Transcendentals come from detmath.h so that host-computed inputs (projection, jitter weights, sun direction)
// are the same bits on every machine and toolchain.
#pragma once
#include <cmath>
#include <cstdint>
#include "detmath.h"

namespace hm {

struct Vec2 { float x = 0, y = 0; };
struct Vec3 { float x = 0, y = 0, z = 0; Vec3() {} Vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit Vec3(float a) : x(a), y(a), z(a) {} };
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; Vec4() {} Vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} Vec4(Vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {} };

inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(float s, Vec3 a) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(Vec3 a) { return dm::sqrt_(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { float l = length(a); return Vec3(a.x / l, a.y / l, a.z / l); }
inline Vec3 vmin(Vec3 a, Vec3 b) { return Vec3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
inline Vec3 vmax(Vec3 a, Vec3 b) { return Vec3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

struct Mat4 {
    float m[16];  // m[col * 4 + row]
    float& at(int col, int row) { return m[col * 4 + row]; }
    float at(int col, int row) const { return m[col * 4 + row]; }
    static Mat4 identity() { Mat4 r{}; for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.f : 0.f; return r; }
    static Mat4 zero() { Mat4 r{}; for (int i = 0; i < 16; i++) r.m[i] = 0.f; return r; }
};
inline Vec4 operator*(const Mat4& a, Vec4 v) {
    Vec4 r;
    r.x = a.at(0, 0) * v.x + a.at(1, 0) * v.y + a.at(2, 0) * v.z + a.at(3, 0) * v.w;
    r.y = a.at(0, 1) * v.x + a.at(1, 1) * v.y + a.at(2, 1) * v.z + a.at(3, 1) * v.w;
    r.z = a.at(0, 2) * v.x + a.at(1, 2) * v.y + a.at(2, 2) * v.z + a.at(3, 2) * v.w;
    r.w = a.at(0, 3) * v.x + a.at(1, 3) * v.y + a.at(2, 3) * v.z + a.at(3, 3) * v.w;
    return r;
}
inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; c++) {
        Vec4 col = a * Vec4(b.at(c, 0), b.at(c, 1), b.at(c, 2), b.at(c, 3));
        r.at(c, 0) = col.x; r.at(c, 1) = col.y; r.at(c, 2) = col.z; r.at(c, 3) = col.w;
    }
    return r;
}
inline Mat4 transpose(const Mat4& a) { Mat4 r; for (int c = 0; c < 4; c++) for (int w = 0; w < 4; w++) r.at(c, w) = a.at(w, c); return r; }
inline Mat4 translate(Vec3 t) { Mat4 r = Mat4::identity(); r.at(3, 0) = t.x; r.at(3, 1) = t.y; r.at(3, 2) = t.z; return r; }
inline Mat4 scale(Vec3 s) { Mat4 r = Mat4::identity(); r.at(0, 0) = s.x; r.at(1, 1) = s.y; r.at(2, 2) = s.z; return r; }
// rotation by angle (radians) about a unit axis
inline Mat4 rotate(float angle, Vec3 axis) {
    float c = dm::cos(angle), s = dm::sin(angle);
    Vec3 a = normalize(axis);
    Vec3 t = a * (1.f - c);
    Mat4 r = Mat4::identity();
    r.at(0, 0) = c + t.x * a.x;       r.at(0, 1) = t.x * a.y + s * a.z; r.at(0, 2) = t.x * a.z - s * a.y;
    r.at(1, 0) = t.y * a.x - s * a.z; r.at(1, 1) = c + t.y * a.y;       r.at(1, 2) = t.y * a.z + s * a.x;
    r.at(2, 0) = t.z * a.x + s * a.y; r.at(2, 1) = t.z * a.y - s * a.x; r.at(2, 2) = c + t.z * a.z;
    return r;
}
// right-handed perspective, clip z in [-1, 1] (what glm::perspective gives without GLM_FORCE_DEPTH_ZERO_TO_ONE)
inline Mat4 perspective(float fovyRadians, float aspect, float zNear, float zFar) {
    float tanHalf = dm::tan(fovyRadians * 0.5f);
    Mat4 r = Mat4::zero();
    r.at(0, 0) = 1.f / (aspect * tanHalf);
    r.at(1, 1) = 1.f / tanHalf;
    r.at(2, 2) = -(zFar + zNear) / (zFar - zNear);
    r.at(2, 3) = -1.f;
    r.at(3, 2) = -(2.f * zFar * zNear) / (zFar - zNear);
    return r;
}
// general 4x4 inverse (cofactor expansion, double accumulation)
inline Mat4 inverse(const Mat4& a) {
    const float* m = a.m;
    double inv[16];
    inv[0] = (double)m[5] * m[10] * m[15] - (double)m[5] * m[11] * m[14] - (double)m[9] * m[6] * m[15] + (double)m[9] * m[7] * m[14] + (double)m[13] * m[6] * m[11] - (double)m[13] * m[7] * m[10];
    inv[4] = -(double)m[4] * m[10] * m[15] + (double)m[4] * m[11] * m[14] + (double)m[8] * m[6] * m[15] - (double)m[8] * m[7] * m[14] - (double)m[12] * m[6] * m[11] + (double)m[12] * m[7] * m[10];
    inv[8] = (double)m[4] * m[9] * m[15] - (double)m[4] * m[11] * m[13] - (double)m[8] * m[5] * m[15] + (double)m[8] * m[7] * m[13] + (double)m[12] * m[5] * m[11] - (double)m[12] * m[7] * m[9];
    inv[12] = -(double)m[4] * m[9] * m[14] + (double)m[4] * m[10] * m[13] + (double)m[8] * m[5] * m[14] - (double)m[8] * m[6] * m[13] - (double)m[12] * m[5] * m[10] + (double)m[12] * m[6] * m[9];
    inv[1] = -(double)m[1] * m[10] * m[15] + (double)m[1] * m[11] * m[14] + (double)m[9] * m[2] * m[15] - (double)m[9] * m[3] * m[14] - (double)m[13] * m[2] * m[11] + (double)m[13] * m[3] * m[10];
    inv[5] = (double)m[0] * m[10] * m[15] - (double)m[0] * m[11] * m[14] - (double)m[8] * m[2] * m[15] + (double)m[8] * m[3] * m[14] + (double)m[12] * m[2] * m[11] - (double)m[12] * m[3] * m[10];
    inv[9] = -(double)m[0] * m[9] * m[15] + (double)m[0] * m[11] * m[13] + (double)m[8] * m[1] * m[15] - (double)m[8] * m[3] * m[13] - (double)m[12] * m[1] * m[11] + (double)m[12] * m[3] * m[9];
    inv[13] = (double)m[0] * m[9] * m[14] - (double)m[0] * m[10] * m[13] - (double)m[8] * m[1] * m[14] + (double)m[8] * m[2] * m[13] + (double)m[12] * m[1] * m[10] - (double)m[12] * m[2] * m[9];
    inv[2] = (double)m[1] * m[6] * m[15] - (double)m[1] * m[7] * m[14] - (double)m[5] * m[2] * m[15] + (double)m[5] * m[3] * m[14] + (double)m[13] * m[2] * m[7] - (double)m[13] * m[3] * m[6];
    inv[6] = -(double)m[0] * m[6] * m[15] + (double)m[0] * m[7] * m[14] + (double)m[4] * m[2] * m[15] - (double)m[4] * m[3] * m[14] - (double)m[12] * m[2] * m[7] + (double)m[12] * m[3] * m[6];
    inv[10] = (double)m[0] * m[5] * m[15] - (double)m[0] * m[7] * m[13] - (double)m[4] * m[1] * m[15] + (double)m[4] * m[3] * m[13] + (double)m[12] * m[1] * m[7] - (double)m[12] * m[3] * m[5];
    inv[14] = -(double)m[0] * m[5] * m[14] + (double)m[0] * m[6] * m[13] + (double)m[4] * m[1] * m[14] - (double)m[4] * m[2] * m[13] - (double)m[12] * m[1] * m[6] + (double)m[12] * m[2] * m[5];
    inv[3] = -(double)m[1] * m[6] * m[11] + (double)m[1] * m[7] * m[10] + (double)m[5] * m[2] * m[11] - (double)m[5] * m[3] * m[10] - (double)m[9] * m[2] * m[7] + (double)m[9] * m[3] * m[6];
    inv[7] = (double)m[0] * m[6] * m[11] - (double)m[0] * m[7] * m[10] - (double)m[4] * m[2] * m[11] + (double)m[4] * m[3] * m[10] + (double)m[8] * m[2] * m[7] - (double)m[8] * m[3] * m[6];
    inv[11] = -(double)m[0] * m[5] * m[11] + (double)m[0] * m[7] * m[9] + (double)m[4] * m[1] * m[11] - (double)m[4] * m[3] * m[9] - (double)m[8] * m[1] * m[7] + (double)m[8] * m[3] * m[5];
    inv[15] = (double)m[0] * m[5] * m[10] - (double)m[0] * m[6] * m[9] - (double)m[4] * m[1] * m[10] + (double)m[4] * m[2] * m[9] + (double)m[8] * m[1] * m[6] - (double)m[8] * m[2] * m[5];
    double det = (double)m[0] * inv[0] + (double)m[1] * inv[4] + (double)m[2] * inv[8] + (double)m[3] * inv[12];
    Mat4 r;
    double id = 1.0 / det;
    for (int i = 0; i < 16; i++) r.m[i] = (float)(inv[i] * id);
    return r;
}

struct AABB { Vec3 min, max; };

}  // namespace hm
