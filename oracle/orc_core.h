// orc_core.h — shared small types of the CPU ORACLE (test infrastructure only).
//
// The oracle is a CPU restatement of the reference's sl::RenderPass::render path. It is the
// CHECKER for the CUDA implementation and must never be linked, imported or called by the
// product (stillleben_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may use it.
//
// PARITY STATUS — pinned on the reference itself wherever the reference can be compiled or imported without a GL stack
// (oracle/build_ref.py -> oracle/_ref/, tests/test_oracle_ref.py, tests/test_glsl_ref.py):
//   * programmable stages: the reference's GLSL sources (render_shader.vert/.frag, shadow_shader.vert, tone_map_shader.frag,
//     ssao_shader.frag, ssao_apply_shader.frag, background_*.vert/.frag, cubemap_shader_*.frag, brdf_shader.frag) are compiled
//     VERBATIM as C++ and run against vertex_stage / fragment_stage / tone map / SSAO / background / light-map texel
//     functions of this oracle on >= 10^4 random inputs per program;
//   * config-4 masks and pose gradients: the reference's own torch extension (python/src/diff.cu + bridge_diff.cpp) and its
//     python/stillleben/diff.py (goldens + live calls);
//   * mesh front end: the reference's consolidate.cpp + compute_tangents.cpp + vendored CgltfImporter (byte for byte);
//   * camera model: the reference's camera_model.py (goldens).
// NOT pinned (no GL implementation exists in this environment, and the reference's tests hold no numeric vectors for it):
// the GL FIXED FUNCTION — clipping, viewport snap, coverage rule, depth quantisation, texture filtering / LOD selection.
// These follow the GL 4.5 core specification under the written contract C1-C7 of DESIGN.md; the reference's known-answer
// assertions (tests/basic.cpp) are re-asserted in tests/test_oracle_pins.py.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../include/slb.h"

namespace orc {

struct V2 { float x, y; };
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct V4 {
    float x, y, z, w;
    V4() : x(0), y(0), z(0), w(0) {}
    V4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    V4(const V3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    V3 xyz() const { return V3(x, y, z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
// GLSL normalize(): x * inversesqrt(dot(x,x)); a zero vector gives NaN/inf like GL does.
inline V3 normalize(V3 a) { float l = length(a); return V3(a.x / l, a.y / l, a.z / l); }
inline V3 vmax(V3 a, V3 b) { return V3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
inline V3 vmin(V3 a, V3 b) { return V3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
inline V4 operator+(V4 a, V4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline V4 operator*(V4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline V3 mix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }

// Column-major 4x4, element (row r, col c) at m[c*4+r] — Magnum::Matrix4 memory order.
struct M4 {
    float m[16];
    float at(int r, int c) const { return m[c * 4 + r]; }
    float& at(int r, int c) { return m[c * 4 + r]; }
    static M4 identity() {
        M4 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r;
    }
    static M4 from(const float* p) { M4 r; std::memcpy(r.m, p, sizeof r.m); return r; }
};
// float mat*mat in the plain left-to-right order (Magnum's operator*, RectangularMatrix.h)
inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += a.at(rr, k) * b.at(k, c);
            r.at(rr, c) = s;
        }
    return r;
}
inline V4 mul(const M4& a, V4 v) {
    V4 r;
    for (int rr = 0; rr < 4; ++rr)
        r[rr] = a.at(rr, 0) * v.x + a.at(rr, 1) * v.y + a.at(rr, 2) * v.z + a.at(rr, 3) * v.w;
    return r;
}
inline V3 transform_point(const M4& a, V3 p) {  // Magnum Matrix4::transformPoint (affine part + divide)
    V4 r = mul(a, V4(p, 1.0f));
    return V3(r.x / r.w, r.y / r.w, r.z / r.w);
}
inline V3 mul3(const float* m9 /*column-major 3x3*/, V3 v) {
    return V3(m9[0] * v.x + m9[3] * v.y + m9[6] * v.z, m9[1] * v.x + m9[4] * v.y + m9[7] * v.z,
              m9[2] * v.x + m9[5] * v.y + m9[8] * v.z);
}
// rigid inverse: R^T, -R^T t (Magnum Matrix4::invertedRigid)
inline M4 inverted_rigid(const M4& a) {
    M4 r = M4::identity();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.at(i, j) = a.at(j, i);
    V3 t(a.at(0, 3), a.at(1, 3), a.at(2, 3));
    for (int i = 0; i < 3; ++i) r.at(i, 3) = -(r.at(i, 0) * t.x + r.at(i, 1) * t.y + r.at(i, 2) * t.z);
    return r;
}
// general inverse via double-precision Gauss-Jordan (Magnum Matrix4::inverted is the adjugate
// formula in float; this is used only for the frustum corners of the shadow fit)
inline M4 inverted(const M4& a) {
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
    M4 o;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o.at(r, c) = (float)w[r][c + 4];
    return o;
}
// cofactor matrix of the upper-left 3x3 (Magnum Matrix4::normalMatrix() == comatrix;
// contrib/magnum/src/Magnum/Math/Matrix4.h:943-947), column-major 3x3 out.
inline void normal_matrix(const M4& a, float* o9) {
    float m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = a.at(r, c);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
            o9[c * 3 + r] = m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1];
        }
}

// ---- the 68-byte consolidated vertex (reference: src/mesh_tools/consolidate.cpp:53-61) ----
#pragma pack(push, 1)
struct Vertex68 {
    float pos[3];
    float uv[2];
    float color[4];
    float tangent[4];
    uint32_t vertex_index;
    float normal[3];
};
#pragma pack(pop)
static_assert(sizeof(Vertex68) == SLB_VERTEX_STRIDE, "vertex stride");

// ---- textures ----
struct Texture {
    int kind = SLB_TEXTURE_2D;
    int w = 0, h = 0, ch = 4;  // stored as RGBA8 always (RGB gets alpha 255)
    int wrap_s = SLB_WRAP_REPEAT, wrap_t = SLB_WRAP_REPEAT;
    int min_filter = SLB_FILTER_LINEAR_MIPMAP_LINEAR, mag_filter = SLB_FILTER_LINEAR;
    bool has_alpha = false;
    struct Level { int w, h; std::vector<uint8_t> px; };
    std::vector<Level> levels;
};

struct Material {
    float base_color[4];
    float emissive[4];
    float metallic, roughness;
    int tex[5];  // base, normal, metallic-roughness, emissive, occlusion; -1 = none
};

struct Mesh {
    std::vector<Vertex68> verts;
    std::vector<uint32_t> indices;
    std::vector<slb_submesh> submeshes;
    std::vector<Material> materials;
    std::vector<Texture> textures;
    float bbox_min[3], bbox_max[3];
};

struct CubeLevel { int size; std::vector<float> px; /* [6][size][size][4] */ };
struct LightMap {
    std::vector<CubeLevel> env;        // 512^2 full mip chain
    std::vector<CubeLevel> irradiance; // 32^2, one level
    std::vector<CubeLevel> prefilter;  // 128^2, 5 levels
    int lut_size = 0;
    std::vector<std::vector<float>> lut; // mip chain of [size][size][4]
    int n_lights = 0;
    float light_directions[SLB_NUM_LIGHTS][3];
    float light_colors[SLB_NUM_LIGHTS][3];
};

// texture helpers implemented in orc_texture.cpp
void build_texture(Texture& t, const slb_image* img, int kind);
V4 sample_texture_2d(const Texture& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy);
V4 sample_texture_rect_linear(const Texture& t, float x, float y);   // clamp-to-border transparent
V4 sample_texture_rect_nearest(const Texture& t, float x, float y);
V4 sample_cube_lod(const std::vector<CubeLevel>& cube, V3 dir, float lod);
V4 sample_lut(const LightMap& lm, float u, float v);
void build_cube_mips(std::vector<CubeLevel>& cube);

// light map precompute implemented in orc_lightmap.cpp
LightMap* lightmap_create(const slb_lightmap_desc* d, int env_size, int irr_size, int pre_size, int lut_size,
                          int n_samples);

}  // namespace orc
