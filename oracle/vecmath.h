// oracle/vecmath.h — tiny float3 algebra for the CPU oracle (TEST INFRASTRUCTURE ONLY).
// Never included by the product path (rfw_rs_b200/csrc).  Compiled with -ffp-contract=off so
// every expression is evaluated as written (no FMA contraction), like the reference's GLSL/Rust.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
static inline V3 vmin(V3 a, V3 b) { return V3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
static inline V3 vmax(V3 a, V3 b) { return V3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
static inline V3 mix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline V3 reflect(V3 I, V3 N) { return I - N * (2.0f * dot(N, I)); }
static inline float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int32_t f2i(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float i2f(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// column-major 4x4 (glam Mat4 layout, crates/rfw-math/src/lib.rs:1-33)
struct M4 {
    float m[16];
    float at(int r, int c) const { return m[c * 4 + r]; }
};
// M * (p, 1).  Summation order (c0 x + c1 y) + (c2 z + c3): the order of `mat4 * vec4` is not specified by GLSL; this is the
// one glm (the library the reference vendors, and what oracle/_ref compiles the shaders against) evaluates, so object-space
// ray origins (ray_gen.comp:340) are bit-identical to the reference-as-compiled-here.  glam sums left to right; the two differ
// in the last bit only.
static inline V3 xform_point(const M4& M, V3 p) {
    return V3((M.at(0, 0) * p.x + M.at(0, 1) * p.y) + (M.at(0, 2) * p.z + M.at(0, 3)),
              (M.at(1, 0) * p.x + M.at(1, 1) * p.y) + (M.at(1, 2) * p.z + M.at(1, 3)),
              (M.at(2, 0) * p.x + M.at(2, 1) * p.y) + (M.at(2, 2) * p.z + M.at(2, 3)));
}
static inline V3 xform_vec(const M4& M, V3 p) {
    return V3(M.at(0, 0) * p.x + M.at(0, 1) * p.y + M.at(0, 2) * p.z,
              M.at(1, 0) * p.x + M.at(1, 1) * p.y + M.at(1, 2) * p.z,
              M.at(2, 0) * p.x + M.at(2, 1) * p.y + M.at(2, 2) * p.z);
}
// general inverse by cofactors, evaluated in double then rounded once (glam's Mat4::inverse is the
// same cofactor expansion in f32; the rounding difference is far below the 1e-4 parity tolerance)
static inline bool invert(const M4& A, M4& out) {
    double a[16], inv[16];
    for (int i = 0; i < 16; i++) a[i] = A.m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0.0 || !std::isfinite(det)) return false;
    double id = 1.0 / det;
    for (int i = 0; i < 16; i++) out.m[i] = (float)(inv[i] * id);
    return true;
}
static inline M4 transpose(const M4& A) {
    M4 t;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) t.m[c * 4 + r] = A.m[r * 4 + c];
    return t;
}

}  // namespace orc
