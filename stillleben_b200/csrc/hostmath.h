// hostmath.h — host-side matrix helpers of the frame set-up (column-major float[16], Magnum order).
// These evaluate what the reference computes on the CPU per frame before it issues GL calls:
//   transformation chain / normal matrix   src/shaders/render_shader.cpp:233-249
//   frustum corners + shadow matrix        src/render_pass.cpp:69-211
// The operation order of mvp() is part of the numerical contract: products are
// formed in double precision, ((a0*b0 + a1*b1) + a2*b2) + a3*b3, and rounded to float once.
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <algorithm>

namespace hm {

struct Vec3 { float x, y, z; };
inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { float l = length(a); return v3(a.x / l, a.y / l, a.z / l); }
inline Vec3 vmin(Vec3 a, Vec3 b) { return v3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
inline Vec3 vmax(Vec3 a, Vec3 b) { return v3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }

struct Mat4 {
    float m[16];
    float at(int r, int c) const { return m[c * 4 + r]; }
    float& at(int r, int c) { return m[c * 4 + r]; }
};
inline Mat4 identity() { Mat4 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
inline Mat4 load(const float* p) { Mat4 r; std::memcpy(r.m, p, sizeof r.m); return r; }
inline Mat4 mul(const Mat4& a, const Mat4& b) {   // float product, k ascending
    Mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += a.at(i, k) * b.at(k, c);
            r.at(i, c) = s;
        }
    return r;
}
inline void mul4(const Mat4& a, const float v[4], float o[4]) {
    for (int i = 0; i < 4; ++i) o[i] = a.at(i, 0) * v[0] + a.at(i, 1) * v[1] + a.at(i, 2) * v[2] + a.at(i, 3) * v[3];
}
inline Vec3 transform_point(const Mat4& a, Vec3 p) {
    float v[4] = {p.x, p.y, p.z, 1.0f}, o[4];
    mul4(a, v, o);
    return v3(o[0] / o[3], o[1] / o[3], o[2] / o[3]);
}
inline Mat4 inverted_rigid(const Mat4& a) {
    Mat4 r = identity();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.at(i, j) = a.at(j, i);
    Vec3 t = v3(a.at(0, 3), a.at(1, 3), a.at(2, 3));
    for (int i = 0; i < 3; ++i) r.at(i, 3) = -(r.at(i, 0) * t.x + r.at(i, 1) * t.y + r.at(i, 2) * t.z);
    return r;
}
// general inverse: Gauss-Jordan with partial pivoting in double precision
inline Mat4 inverted(const Mat4& a) {
    double w[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { w[r][c] = a.at(r, c); w[r][c + 4] = (r == c) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
        if (p != c) for (int k = 0; k < 8; ++k) std::swap(w[p][k], w[c][k]);
        double d = w[c][c];
        for (int k = 0; k < 8; ++k) w[c][k] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) {
            double f = w[r][c];
            for (int k = 0; k < 8; ++k) w[r][k] -= f * w[c][k];
        }
    }
    Mat4 o;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o.at(r, c) = (float)w[r][c + 4];
    return o;
}
// cofactor matrix of the upper-left 3x3 (Magnum's normalMatrix()), column-major 3x3 out
inline void normal_matrix(const Mat4& a, float* o9) {
    float m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = a.at(r, c);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
            o9[c * 3 + r] = m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1];
        }
}
// contract C2: mvp = P * (V * (world * pre)) in double, rounded once
inline void mul44d(const double* a, const double* b, double* c) {
    for (int col = 0; col < 4; ++col)
        for (int r = 0; r < 4; ++r)
            c[col * 4 + r] = ((a[0 * 4 + r] * b[col * 4 + 0] + a[1 * 4 + r] * b[col * 4 + 1]) + a[2 * 4 + r] * b[col * 4 + 2]) +
                             a[3 * 4 + r] * b[col * 4 + 3];
}
inline void mvp(const Mat4& P, const Mat4& V, const Mat4& world, const Mat4& pre, float* out) {
    double p[16], v[16], w[16], m[16], mw[16], mc[16], r[16];
    for (int i = 0; i < 16; ++i) { p[i] = P.m[i]; v[i] = V.m[i]; w[i] = world.m[i]; m[i] = pre.m[i]; }
    mul44d(w, m, mw); mul44d(v, mw, mc); mul44d(p, mc, r);
    for (int i = 0; i < 16; ++i) out[i] = (float)r[i];
}

// The same products with the double-precision operands kept by the caller (the frame's P and V and an object's
// world * pre are shared by many draws): bit-identical to mvp(). `v == nullptr` stands for V = identity, whose
// product is exact and is skipped.
inline void to_double(const Mat4& a, double* o) { for (int i = 0; i < 16; ++i) o[i] = a.m[i]; }
inline void world_pre_d(const Mat4& world, const Mat4& pre, double* mw) {
    double w[16], m[16];
    to_double(world, w); to_double(pre, m);
    mul44d(w, m, mw);
}
inline void mvp_d(const double* p, const double* v, const double* mw, float* out) {
    double mc[16], r[16];
    if (v) { mul44d(v, mw, mc); mul44d(p, mc, r); } else mul44d(p, mw, r);
    for (int i = 0; i < 16; ++i) out[i] = (float)r[i];
}

}  // namespace hm
