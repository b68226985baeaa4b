// k_frag.cuh — programmable stages of the reference as device functions:
//   vertex stage     src/shaders/render_shader.vert:57-95
//   geometry stage   src/shaders/render_shader.geom:13-35   (barycentric basis, flat vertex ids)
//   fragment stage   src/shaders/render_shader.frag:225-412
// plus the texture units they use (GL 4.5 core §8.14: LOD selection, bilinear / trilinear filtering,
// wrap modes; §8.13: cube-map face selection, seamless filtering), written out explicitly in fp32
// so the results do not depend on the 8-bit filter weights of the hardware texture path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "k_contract.cuh"
#include "slb_dev.h"

namespace slbk {

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ f3 normalize3(f3 a) { float r = rsqrtf(dot3(a, a)); return mk3(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ f3 max3(f3 a, f3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
__device__ __forceinline__ f3 mix3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// column-major 4x4 times (x,y,z,1)
__device__ __forceinline__ float4 mul_m4_p(const float* __restrict__ m, float x, float y, float z, float w) {
    return make_float4(m[0] * x + m[4] * y + m[8] * z + m[12] * w, m[1] * x + m[5] * y + m[9] * z + m[13] * w,
                       m[2] * x + m[6] * y + m[10] * z + m[14] * w, m[3] * x + m[7] * y + m[11] * z + m[15] * w);
}
__device__ __forceinline__ f3 mul_m3(const float* __restrict__ m9, f3 v) {
    return mk3(m9[0] * v.x + m9[3] * v.y + m9[6] * v.z, m9[1] * v.x + m9[4] * v.y + m9[7] * v.z,
               m9[2] * v.x + m9[5] * v.y + m9[8] * v.z);
}

// ---------------------------------------------------------------------------------------------
// texture units
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int wrap_index(int i, int n, int mode, bool& border) {
    switch (mode) {
        case SLB_WRAP_REPEAT: { if ((n & (n - 1)) == 0) return i & (n - 1); int m = i % n; return m < 0 ? m + n : m; }
        case SLB_WRAP_MIRRORED_REPEAT: { int p = 2 * n; int m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m; }
        case SLB_WRAP_CLAMP_TO_BORDER: if (i < 0 || i >= n) { border = true; return 0; } return i;
        default: return min(max(i, 0), n - 1);
    }
}
__device__ __forceinline__ float4 tex_load(const DTexture& t, int level, int lw, int xi, int yi) {
    uchar4 p = __ldg(reinterpret_cast<const uchar4*>(t.px) + t.level_off[level] + (size_t)yi * lw + xi);
    const float k = 1.0f / 255.0f;
    return make_float4(p.x * k, p.y * k, p.z * k, p.w * k);
}
__device__ __forceinline__ float4 tex_fetch(const DTexture& t, int level, int lw, int lh, int x, int y) {
    bool border = false;
    int xi = wrap_index(x, lw, t.wrap_s, border);
    int yi = wrap_index(y, lh, t.wrap_t, border);
    if (border) return make_float4(0.f, 0.f, 0.f, 0.f);
    return tex_load(t, level, lw, xi, yi);
}
__device__ __forceinline__ float4 tex_sample_level(const DTexture& t, int level, float u, float v, bool normalised, bool linear) {
    int lw = max(1, t.w >> level), lh = max(1, t.h >> level);
    float xs = normalised ? u * (float)lw : u, ys = normalised ? v * (float)lh : v;
    if (!linear) return tex_fetch(t, level, lw, lh, (int)floorf(xs), (int)floorf(ys));
    float x = xs - 0.5f, y = ys - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    float4 t00, t10, t01, t11;
    if (t.wrap_s == SLB_WRAP_REPEAT && t.wrap_t == SLB_WRAP_REPEAT && !(lw & (lw - 1)) && !(lh & (lh - 1))) {
        // the common case (glTF default wrap, power-of-two levels): four masks instead of eight mode switches
        const int x0 = i0 & (lw - 1), x1 = (i0 + 1) & (lw - 1), y0 = j0 & (lh - 1), y1 = (j0 + 1) & (lh - 1);
        t00 = tex_load(t, level, lw, x0, y0); t10 = tex_load(t, level, lw, x1, y0);
        t01 = tex_load(t, level, lw, x0, y1); t11 = tex_load(t, level, lw, x1, y1);
    } else {
        t00 = tex_fetch(t, level, lw, lh, i0, j0); t10 = tex_fetch(t, level, lw, lh, i0 + 1, j0);
        t01 = tex_fetch(t, level, lw, lh, i0, j0 + 1); t11 = tex_fetch(t, level, lw, lh, i0 + 1, j0 + 1);
    }
    return t00 * ((1 - a) * (1 - b)) + t10 * (a * (1 - b)) + t01 * ((1 - a) * b) + t11 * (a * b);
}
// texture2D() with implicit derivatives (du/dx etc. in normalised units per pixel)
static __device__ float4 tex_sample_2d(const DTexture& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy) {
    const int max_level = t.n_levels - 1;
    float W = (float)t.w, H = (float)t.h;
    float rx = sqrtf(dudx * W * dudx * W + dvdx * H * dvdx * H);
    float ry = sqrtf(dudy * W * dudy * W + dvdy * H * dvdy * H);
    float rho = fmaxf(rx, ry);
    float lambda = log2f(rho);
    bool mag_linear = (t.mag_filter == SLB_FILTER_LINEAR);
    int mf = t.min_filter;
    float c = (mag_linear && (mf == SLB_FILTER_NEAREST_MIPMAP_NEAREST || mf == SLB_FILTER_NEAREST_MIPMAP_LINEAR)) ? 0.5f : 0.0f;
    if (!(lambda > c)) return tex_sample_level(t, 0, u, v, true, mag_linear);
    bool lin = (mf == SLB_FILTER_LINEAR || mf == SLB_FILTER_LINEAR_MIPMAP_NEAREST || mf == SLB_FILTER_LINEAR_MIPMAP_LINEAR);
    if (mf == SLB_FILTER_NEAREST || mf == SLB_FILTER_LINEAR) return tex_sample_level(t, 0, u, v, true, lin);
    if (mf == SLB_FILTER_NEAREST_MIPMAP_NEAREST || mf == SLB_FILTER_LINEAR_MIPMAP_NEAREST) {
        int d = (lambda <= 0.5f) ? 0 : min((int)ceilf(lambda + 0.5f) - 1, max_level);
        return tex_sample_level(t, d, u, v, true, lin);
    }
    float lc = fminf(lambda, (float)max_level);
    int d1 = (int)floorf(lc);
    int d2 = min(d1 + 1, max_level);
    float f = lc - (float)d1;
    float4 s1 = tex_sample_level(t, d1, u, v, true, lin);
    if (d2 == d1 || f == 0.0f) return s1;
    float4 s2 = tex_sample_level(t, d2, u, v, true, lin);
    return s1 * (1.0f - f) + s2 * f;
}
__device__ __forceinline__ float4 tex_sample_rect(const DTexture& t, float x, float y) {
    return tex_sample_level(t, 0, x, y, false, t.mag_filter == SLB_FILTER_LINEAR);
}

// cube maps: face order +X,-X,+Y,-Y,+Z,-Z (GL 4.5 table 8.19)
__device__ __forceinline__ void cube_face_coords(f3 d, int& face, float& s, float& t) {
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { if (d.x >= 0) { face = 0; sc = -d.z; tc = -d.y; } else { face = 1; sc = d.z; tc = -d.y; } ma = ax; }
    else if (ay >= az)        { if (d.y >= 0) { face = 2; sc = d.x;  tc = d.z;  } else { face = 3; sc = d.x; tc = -d.z; } ma = ay; }
    else                      { if (d.z >= 0) { face = 4; sc = d.x;  tc = -d.y; } else { face = 5; sc = -d.x; tc = -d.y; } ma = az; }
    s = 0.5f * (sc / ma + 1.0f);
    t = 0.5f * (tc / ma + 1.0f);
}
__device__ __forceinline__ f3 cube_face_dir(int face, float s, float t) {
    float a = 2.0f * s - 1.0f, b = 2.0f * t - 1.0f;
    switch (face) {
        case 0: return mk3(1, -b, -a);
        case 1: return mk3(-1, -b, a);
        case 2: return mk3(a, 1, b);
        case 3: return mk3(a, -1, -b);
        case 4: return mk3(a, -b, 1);
        default: return mk3(-a, -b, -1);
    }
}
__device__ __forceinline__ float4 cube_tap(const DCubeLevel& l, int face, int x, int y) {
    int n = l.size;
    if (x < 0 || x >= n || y < 0 || y >= n) {   // seamless: re-project through the texel centre
        f3 d = cube_face_dir(face, (x + 0.5f) / n, (y + 0.5f) / n);
        float s, t; cube_face_coords(d, face, s, t);
        x = min(max((int)floorf(s * n), 0), n - 1);
        y = min(max((int)floorf(t * n), 0), n - 1);
    }
    return __ldg(l.px + ((size_t)face * n + y) * n + x);
}
__device__ __forceinline__ float4 cube_sample_level(const DCubeLevel& l, f3 dir) {
    int face; float s, t; cube_face_coords(dir, face, s, t);
    float x = s * l.size - 0.5f, y = t * l.size - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    float4 t00 = cube_tap(l, face, i0, j0), t10 = cube_tap(l, face, i0 + 1, j0);
    float4 t01 = cube_tap(l, face, i0, j0 + 1), t11 = cube_tap(l, face, i0 + 1, j0 + 1);
    return t00 * ((1 - a) * (1 - b)) + t10 * (a * (1 - b)) + t01 * ((1 - a) * b) + t11 * (a * b);
}
__device__ __forceinline__ float4 cube_sample_lod(const DCubeLevel* levels, int n_levels, f3 dir, float lod) {
    int max_level = n_levels - 1;
    float lc = fminf(fmaxf(lod, 0.0f), (float)max_level);
    int d1 = (int)floorf(lc);
    int d2 = min(d1 + 1, max_level);
    float f = lc - (float)d1;
    float4 s1 = cube_sample_level(levels[d1], dir);
    if (d2 == d1 || f == 0.0f) return s1;
    float4 s2 = cube_sample_level(levels[d2], dir);
    return s1 * (1.0f - f) + s2 * f;
}
__device__ __forceinline__ float4 lut_sample(const DLightMap& lm, float u, float v) {
    int n = lm.lut_size;
    float x = u * n - 0.5f, y = v * n - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    auto at = [&](int xi, int yi) { xi = min(max(xi, 0), n - 1); yi = min(max(yi, 0), n - 1); return __ldg(lm.lut + (size_t)yi * n + xi); };
    return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) + at(i0 + 1, j0 + 1) * (a * b);
}

// ---------------------------------------------------------------------------------------------
// vertex stage (render_shader.vert:57-95)
// ---------------------------------------------------------------------------------------------
struct VSOut {
    float u, v;
    f3 nW;
    float4 objc;   // xyz object frame, w = camera z
    f3 wc, cc;
    float su, sv;  // sticker coordinates
};
__device__ __forceinline__ void vertex_stage(const DFrame& f, const DDraw& d, uint32_t vi, VSOut& o, uint32_t& vertex_id) {
    float4 p4 = __ldg(d.pos4 + vi);
    vertex_id = __float_as_uint(p4.w);
    float4 a0 = __ldg(d.attr + 3 * (size_t)vi), a1 = __ldg(d.attr + 3 * (size_t)vi + 1);
    float4 oc4 = mul_m4_p(d.meshToObject, p4.x, p4.y, p4.z, 1.0f);
    o.objc = make_float4(oc4.x / oc4.w, oc4.y / oc4.w, oc4.z / oc4.w, 1.0f);
    float4 wc4 = mul_m4_p(d.objectToWorld, oc4.x, oc4.y, oc4.z, oc4.w);
    o.wc = mk3(wc4.x / wc4.w, wc4.y / wc4.w, wc4.z / wc4.w);
    float4 cc4 = mul_m4_p(f.V, wc4.x, wc4.y, wc4.z, wc4.w);
    o.cc = mk3(cc4.x / cc4.w, cc4.y / cc4.w, cc4.z / cc4.w);
    o.objc.w = o.cc.z;
    f3 n = mk3(a0.z, a0.w, a1.x);
    o.nW = normalize3(mul_m3(d.normalToWorld, n));
    o.u = a0.x; o.v = a0.y;
    if (d.sticker) {
        float4 sp = mul_m4_p(d.stickerProj, oc4.x, oc4.y, oc4.z, oc4.w);
        o.su = (sp.x / sp.w - d.stickerRange[0]) / d.stickerRange[2];
        o.sv = (sp.y / sp.w - d.stickerRange[1]) / d.stickerRange[3];
    } else { o.su = -1.0f; o.sv = -1.0f; }
}

// interpolated world-space tangent and bitangent (render_shader.vert:75-84) — only evaluated for draws with
// a normal texture, so the common path does not carry them
static __device__ __noinline__ void tangent_frame(const DDraw& d, const uint32_t vi[3], const float bary[3], f3& tW, f3& bW) {
    tW = bW = mk3(0.f, 0.f, 0.f);
    for (int j = 0; j < 3; ++j) {
        float4 a0 = __ldg(d.attr + 3 * (size_t)vi[j]), a1 = __ldg(d.attr + 3 * (size_t)vi[j] + 1), a2 = __ldg(d.attr + 3 * (size_t)vi[j] + 2);
        f3 n = normalize3(mul_m3(d.normalToWorld, mk3(a0.z, a0.w, a1.x)));
        f3 t = normalize3(mul_m3(d.normalToWorld, mk3(a1.y, a1.z, a1.w)));
        f3 b = normalize3(cross3(n, t)) * a2.x;
        tW = tW + t * bary[j]; bW = bW + b * bary[j];
    }
}

// perspective-correct barycentrics w.r.t. the ORIGINAL triangle from the three edge-function values of a
// sub-triangle (render_shader.geom:13-35: smooth-interpolated barycentric basis). `unit_basis`: the sub-triangle
// IS the unclipped primitive, its vertices carry the basis (1,0,0),(0,1,0),(0,0,1) and the combination is the identity.
__device__ __forceinline__ void bary_from_weights(const SubTri& st, const PolyV& a, const PolyV& b, const PolyV& c, long long w0,
                                                  long long w1, long long w2, bool unit_basis, float out[3]) {
    float b0 = __ll2float_rn(w0) * st.inv2A, b1 = __ll2float_rn(w1) * st.inv2A, b2 = __ll2float_rn(w2) * st.inv2A;
    float g0 = b0 * a.invw, g1 = b1 * b.invw, g2 = b2 * c.invw;
    float s = g0 + g1 + g2;
    float q0 = g0 / s, q1 = g1 / s, q2 = g2 / s;
    if (unit_basis) { out[0] = q0; out[1] = q1; out[2] = q2; return; }
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] = q0 * a.b[j] + q1 * b.b[j] + q2 * c.b[j];
}

// The three snapped vertices of one sub-triangle of a primitive of draw d. `kbyte` is the low byte of the
// visibility key: fan index k in bits 0..2 and, for a clipped primitive, the slot + 1 of the ClipRec the setup
// kernel published in bits 3..7 — the clipped, snapped polygon is then read back instead of re-clipped
// (returns 2). Unclipped primitives are re-set-up in registers (returns 1); 0 = culled. pm[] receives the three
// mesh-space vertices (xyz + one-based vertex id in w).
__device__ __forceinline__ int fetch_subtri(const DFrame& f, const DDraw& d, uint32_t seq, int kbyte, const uint32_t vi[3], float4 pm[3],
                                            PolyV& a, PolyV& b, PolyV& c) {
    pm[0] = __ldg(d.pos4 + vi[0]); pm[1] = __ldg(d.pos4 + vi[1]); pm[2] = __ldg(d.pos4 + vi[2]);
    const int k = kbyte & 7, slot = kbyte >> 3;
    if (slot) {
        const ClipRec& cr = f.clip[slot - 1];
        if (k < 1 || k + 1 >= cr.n) return 0;
        const DPolyV &va = cr.v[0], &vb = cr.v[k], &vc = cr.v[k + 1];
        a.X = va.X; a.Y = va.Y; a.z = va.z; a.invw = va.invw; a.b[0] = va.b[0]; a.b[1] = va.b[1]; a.b[2] = va.b[2];
        b.X = vb.X; b.Y = vb.Y; b.z = vb.z; b.invw = vb.invw; b.b[0] = vb.b[0]; b.b[1] = vb.b[1]; b.b[2] = vb.b[2];
        c.X = vc.X; c.Y = vc.Y; c.z = vc.z; c.invw = vc.invw; c.b[0] = vc.b[0]; c.b[1] = vc.b[1]; c.b[2] = vc.b[2];
        return 2;
    }
    const float3 q0 = make_float3(pm[0].x, pm[0].y, pm[0].z), q1 = make_float3(pm[1].x, pm[1].y, pm[1].z), q2 = make_float3(pm[2].x, pm[2].y, pm[2].z);
    int r = setup_subtri_fast(d.mvp, q0, q1, q2, f.W, f.H, k, a, b, c);
    if (r >= 0) return r;
    (void)seq;   // clipped, but the frame's ClipRec table was full: clip again. Temporaries: the out-of-line call takes
    // addresses, and a, b, c must stay in registers on the hot path.
    PolyV ta, tb, tc;
    if (!resetup_clipped(d.mvp, q0, q1, q2, f.W, f.H, k, ta, tb, tc)) return 0;
    a = ta; b = tb; c = tc;
    return 2;
}

struct FragIn {
    float u, v, u_dx, v_dx, u_dy, v_dy;   // uv at the pixel and the per-pixel finite differences
    f3 nW; float4 objc; f3 wc, cc; float su, sv;
    bool front;
};

// General vertex-stage path (a transformation with a projective last row): the vertex stage runs on the three
// vertices and its outputs are interpolated, one vertex at a time (sum order ((v0*b0 + v1*b1) + v2*b2)).
static __device__ __noinline__ void interpolate_general(const DFrame& f, const DDraw& d, const uint32_t vi[3], const float bary[3],
                                                        FragIn& in) {
    in.su = in.sv = 0.f;
    in.wc = in.cc = mk3(0.f, 0.f, 0.f);
    in.objc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < 3; ++j) {
        VSOut v; uint32_t id;
        vertex_stage(f, d, vi[j], v, id);
        const float w = bary[j];
        in.wc = in.wc + v.wc * w; in.cc = in.cc + v.cc * w;
        in.objc = in.objc + v.objc * w;
        in.su += v.su * w; in.sv += v.sv * w;
    }
}

// Inputs of the fragment stage at pixel (px,py): the interpolated outputs of the vertex stage
// (render_shader.vert:57-95). For affine transformation chains (every rigid / scaled pose: DRAW_AFFINE) the
// positions are interpolated in the mesh frame and transformed ONCE — mathematically identical because the
// weights sum to one, and a third of the arithmetic; normals and sticker coordinates are non-linear per vertex
// and stay per vertex. dFdx / dFdy of uv come from the 2x2 quad partner evaluated on the same primitive
// (helper invocation); its edge-function values are the pixel's own plus / minus one exact integer step.
template <bool LEAN = false>
__device__ __forceinline__ void shade_inputs(const DFrame& f, const DDraw& d, const SubTri& st, const PolyV& a, const PolyV& b,
                                             const PolyV& c, const uint32_t vi[3], const float4 pm[3], bool unit_basis, int px, int py,
                                             bool want_derivs, FragIn& in, float bary[3], uint32_t vid[3]) {
    long long w0, w1, w2; subtri_weights(st, px, py, w0, w1, w2);
    bary_from_weights(st, a, b, c, w0, w1, w2, unit_basis, bary);
    float bx[3] = {0.f, 0.f, 0.f}, by[3] = {0.f, 0.f, 0.f};
    if (want_derivs) {
        const long long sx = (px & 1) ? -256 : 256, sy = (py & 1) ? -256 : 256;   // towards the quad partner
        bary_from_weights(st, a, b, c, w0 - sx * (st.cy - st.by), w1 - sx * (st.ay - st.cy), w2 - sx * (st.by - st.ay), unit_basis, bx);
        bary_from_weights(st, a, b, c, w0 + sy * (st.cx - st.bx), w1 + sy * (st.ax - st.cx), w2 + sy * (st.bx - st.ax), unit_basis, by);
    }
    float ux = 0.f, vx = 0.f, uy = 0.f, vy = 0.f;
    in.u = in.v = 0.f;
    in.nW = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // uv and the (per-vertex normalised) world-space normal
        vid[j] = __float_as_uint(pm[j].w);
        const float4 a0 = __ldg(d.attr + 3 * (size_t)vi[j]), a1 = __ldg(d.attr + 3 * (size_t)vi[j] + 1);
        const float w = bary[j];
        in.u += a0.x * w; in.v += a0.y * w;
        in.nW = in.nW + normalize3(mul_m3(d.normalToWorld, mk3(a0.z, a0.w, a1.x))) * w;
        ux += a0.x * bx[j]; vx += a0.y * bx[j]; uy += a0.x * by[j]; vy += a0.y * by[j];
    }
    if (LEAN || (d.flags & DRAW_AFFINE)) {
        const float mx = pm[0].x * bary[0] + pm[1].x * bary[1] + pm[2].x * bary[2];
        const float my = pm[0].y * bary[0] + pm[1].y * bary[1] + pm[2].y * bary[2];
        const float mz = pm[0].z * bary[0] + pm[1].z * bary[1] + pm[2].z * bary[2];
        const float* M = d.meshToObject;
        const float ox = M[0] * mx + M[4] * my + M[8] * mz + M[12], oy = M[1] * mx + M[5] * my + M[9] * mz + M[13],
                    oz = M[2] * mx + M[6] * my + M[10] * mz + M[14];
        const float* O = d.objectToWorld;
        in.wc = mk3(O[0] * ox + O[4] * oy + O[8] * oz + O[12], O[1] * ox + O[5] * oy + O[9] * oz + O[13],
                    O[2] * ox + O[6] * oy + O[10] * oz + O[14]);
        const float* V = f.V;
        in.cc = mk3(V[0] * in.wc.x + V[4] * in.wc.y + V[8] * in.wc.z + V[12], V[1] * in.wc.x + V[5] * in.wc.y + V[9] * in.wc.z + V[13],
                    V[2] * in.wc.x + V[6] * in.wc.y + V[10] * in.wc.z + V[14]);
        in.objc = make_float4(ox, oy, oz, in.cc.z);
        in.su = in.sv = -1.0f;
        if (!LEAN && d.sticker) {   // projective per vertex (render_shader.vert:86-92), then interpolated
            in.su = in.sv = 0.f;
            for (int j = 0; j < 3; ++j) {
                const float qx = M[0] * pm[j].x + M[4] * pm[j].y + M[8] * pm[j].z + M[12], qy = M[1] * pm[j].x + M[5] * pm[j].y + M[9] * pm[j].z + M[13],
                            qz = M[2] * pm[j].x + M[6] * pm[j].y + M[10] * pm[j].z + M[14];
                const float4 sp = mul_m4_p(d.stickerProj, qx, qy, qz, 1.0f);
                in.su += (sp.x / sp.w - d.stickerRange[0]) / d.stickerRange[2] * bary[j];
                in.sv += (sp.y / sp.w - d.stickerRange[1]) / d.stickerRange[3] * bary[j];
            }
        }
    } else if (!LEAN) {   // rare: copies, so that the out-of-line call does not pin `in`, `vi`, `bary` to local memory
        FragIn tmp;
        const uint32_t vi2[3] = {vi[0], vi[1], vi[2]};
        const float b2[3] = {bary[0], bary[1], bary[2]};
        interpolate_general(f, d, vi2, b2, tmp);
        in.wc = tmp.wc; in.cc = tmp.cc; in.objc = tmp.objc; in.su = tmp.su; in.sv = tmp.sv;
    }
    in.front = st.twoA < 0;   // FrontFace = CW (render_pass.cpp:330)
    in.u_dx = in.v_dx = in.u_dy = in.v_dy = 0.0f;
    if (want_derivs) {
        float sgx = (px & 1) ? -1.0f : 1.0f, sgy = (py & 1) ? -1.0f : 1.0f;
        in.u_dx = sgx * (ux - in.u); in.v_dx = sgx * (vx - in.v);
        in.u_dy = sgy * (uy - in.u); in.v_dy = sgy * (vy - in.v);
    }
}

__device__ __forceinline__ float4 sample_mat(const DTexture* t, const FragIn& in) {
    return tex_sample_2d(*t, in.u, in.v, in.u_dx, in.v_dx, in.u_dy, in.v_dy);
}
__device__ __forceinline__ float pow22(float x) { return x > 0.0f ? exp2f(2.2f * __log2f(x)) : 0.0f; }
__device__ __forceinline__ float4 to_linear(float4 c) { return make_float4(pow22(c.x), pow22(c.y), pow22(c.z), c.w); }
__device__ __forceinline__ bool draw_has_textures(const DDraw& d) { return d.tex[0] || d.tex[1] || d.tex[2] || d.tex[3] || d.tex[4]; }

// base colour incl. alpha (render_shader.frag:237-246)
__device__ __forceinline__ float4 base_color(const DDraw& d, const FragIn& in) {
    float4 bc = make_float4(d.base_color[0], d.base_color[1], d.base_color[2], d.base_color[3]);
    if (d.tex[0]) { float4 t = to_linear(sample_mat(d.tex[0], in)); bc = make_float4(bc.x * t.x, bc.y * t.y, bc.z * t.z, bc.w * t.w); }
    return bc;
}

__device__ __forceinline__ float DistributionGGX(f3 N, f3 H, float roughness) {
    float a = roughness * roughness, a2 = a * a;
    float NdotH = fmaxf(dot3(N, H), 0.0f), NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = 3.141592653589793f * denom * denom;
    return a2 / denom;
}
__device__ __forceinline__ float GeometrySchlickGGX(float NdotV, float roughness) {
    float r = roughness + 1.0f, k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}

// sampler2DArrayShadow (render_shader.frag:321-337): each tap is a linear filter of the comparison
// `ref <= stored` with stored = d24 / 16777215. All 16 taps of a light share `ref`, so the comparison is
// turned into an integer one ONCE: shadow_threshold() returns the smallest d24 whose float quotient
// reaches ref (exact, the quotient is monotonic in d24) and every tap compares integers.
__device__ __forceinline__ uint32_t shadow_threshold(float ref) {
    // Closed form of "smallest d24 with fl32(d24 / 16777215) >= ref": the rounded quotient reaches ref exactly when
    // d24 / C lies at or above the midpoint between ref and its float predecessor (on the midpoint itself only if
    // ties-to-even lands on ref, i.e. ref's mantissa is even). With midpoint = P * 2^-s (P odd, < 2^25) this is
    // d24 >= ceil(C * P / 2^s) in 64-bit integers — no division, no search loop. NaN never passes (0x1000000).
    if (!(ref == ref)) return 0x1000000u;
    const uint32_t bits = __float_as_uint(clampf(ref, 0.0f, 1.0f)) & 0x7FFFFFFFu;
    if (bits == 0u) return 0u;
    const uint32_t e = bits >> 23, m = bits & 0x7FFFFFu;
    uint32_t P; int s; bool odd;
    if (e == 0u) { P = 2u * m - 1u; s = 150; odd = m & 1u; }                                  // denormal
    else if (m) { const uint32_t M = m | 0x800000u; P = 2u * M - 1u; s = 151 - (int)e; odd = M & 1u; }
    else { P = 0x1FFFFFFu; s = 152 - (int)e; odd = false; }                                  // power of two: the predecessor is half as far
    if (s >= 62) return 1u;
    const unsigned long long N = 16777215ull * P, mask = (1ull << s) - 1ull;
    uint32_t d = (uint32_t)((N + mask) >> s);
    if ((N & mask) == 0ull && odd) ++d;
    return d;
}
// 4x4 PCF taps at offsets {-1.5,-0.5,0.5,1.5} texels, each a 2x2 bilinear filter of the comparison: the 16
// taps share their fractional position, so the sum separates into weights (1-a, 1, 1, 1, a) x (1-b, 1, 1, 1, b)
// over the 5x5 texel neighbourhood — 25 compares instead of 64 (clamp-to-edge per texel index).
__device__ __forceinline__ float shadow_pcf16(const uint32_t* __restrict__ map, const uint32_t* __restrict__ mask, float u, float v, uint32_t thr,
                                              uint32_t tagbits) {
    const int N = SLB_SHADOW_RES;
    float x = u * N - 2.0f, y = v * N - 2.0f;            // (u - 1.5/N) * N - 0.5
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    const float wx[5] = {1.0f - a, 1.0f, 1.0f, 1.0f, a};
    const float wy[5] = {1.0f - b, 1.0f, 1.0f, 1.0f, b};
    const uint32_t thr_t = tagbits | thr;   // texels are generation tag << 24 | d24 (DFrame::shadow_tagbits)
    float sum = 0.0f;
    bool touched = true;
    if (mask) {   // the (at most 2 x 2) texel blocks under the clamped 5x5 footprint: all clear -> every texel is untouched = lit
        const int SH = SLB_SHADOW_MASK_SHIFT, RW = SLB_SHADOW_MASK_ROW;
        const int bx0 = min(max(i0, 0), N - 1) >> SH, bx1 = min(max(i0 + 4, 0), N - 1) >> SH;
        const int by0 = min(max(j0, 0), N - 1) >> SH, by1 = min(max(j0 + 4, 0), N - 1) >> SH;
        const uint32_t m = ((__ldg(mask + by0 * RW + (bx0 >> 5)) >> (bx0 & 31)) | (__ldg(mask + by0 * RW + (bx1 >> 5)) >> (bx1 & 31)) |
                            (__ldg(mask + by1 * RW + (bx0 >> 5)) >> (bx0 & 31)) | (__ldg(mask + by1 * RW + (bx1 >> 5)) >> (bx1 & 31))) & 1u;
        touched = m != 0u;
    }
    if (touched) {
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) {
            const uint32_t* row = map + (size_t)min(max(j0 + jj, 0), N - 1) * N;
            float r = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 5; ++ii)
                r += (__ldg(row + min(max(i0 + ii, 0), N - 1)) >= thr_t) ? wx[ii] : 0.0f;   // stale / untouched texels carry a larger tag: lit
            sum += r * wy[jj];
        }
    } else {   // the same sums with every comparison true (same operation order: bit-identical to the taps)
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) {
            float r = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 5; ++ii) r += wx[ii];
            sum += r * wy[jj];
        }
    }
    return thr > 0xFFFFFFu ? 0.0f : sum * (1.0f / 16.0f);   // NaN reference: no tap passes
}

// fragment stage (render_shader.frag:225-412); the discards are evaluated by the rasteriser
// LEAN = the sub-batch uses none of: normal / metallic-roughness / emissive / occlusion textures, stickers, light maps,
// projective transformation chains (decided on the host per sub-batch). The lean instantiation compiles those paths
// out: fewer live registers in the common case (base-colour textures, analytic lights, PCF shadows).
template <bool LEAN = false>
__device__ __forceinline__ void fragment_stage(const DFrame& f, const DDraw& d, const FragIn& in, const uint32_t vi[3],
                                               const float bary[3], float4& out_color, float4& out_normal) {
    const float PI = 3.141592653589793f;
    float4 baseColor = base_color(d, in);
    if (!LEAN && d.sticker && in.su >= 0 && in.sv >= 0 && in.su < 1 && in.sv < 1) {
        float4 sc = to_linear(tex_sample_rect(*d.sticker, in.su * d.sticker->w, in.sv * d.sticker->h));
        float a = sc.w;
        baseColor = make_float4(baseColor.x * (1 - a) + sc.x * a, baseColor.y * (1 - a) + sc.y * a, baseColor.z * (1 - a) + sc.z * a,
                                baseColor.w * (1 - a) + sc.w * a);
    }
    f3 normal;
    if (!LEAN && d.tex[1]) {
        float4 t = sample_mat(d.tex[1], in);
        f3 tW, bW;
        {
            const uint32_t vi2[3] = {vi[0], vi[1], vi[2]};
            const float b2[3] = {bary[0], bary[1], bary[2]};
            f3 t2, bb2;
            tangent_frame(d, vi2, b2, t2, bb2);   // out of line: operate on copies (see shade_inputs)
            tW = t2; bW = bb2;
        }
        normal = normalize3(tW * (t.x * 2.0f - 1.0f) + bW * (t.y * 2.0f - 1.0f) + in.nW * (t.z * 2.0f - 1.0f));
    } else normal = in.nW;
    if (!in.front) normal = -normal;

    f3 camPos = mk3(f.camPos[0], f.camPos[1], f.camPos[2]);
    f3 cameraDirection = normalize3(camPos - in.wc);
    f3 I = -cameraDirection;
    f3 reflDir = I - normal * (2.0f * dot3(normal, I));
    float NoV = clampf(dot3(normal, cameraDirection), 1e-5f, 1.0f);

    float roughness = d.roughness, metallic = d.metallic;
    if (!LEAN && d.tex[2]) { float4 t = sample_mat(d.tex[2], in); roughness *= t.y; metallic *= t.z; }
    roughness = fmaxf(roughness, 0.045f);
    float occlusion = 1.0f;
    if (!LEAN && d.tex[4]) occlusion = sample_mat(d.tex[4], in).x;
    f3 emissive = mk3(d.emissive[0], d.emissive[1], d.emissive[2]);
    if (!LEAN && d.tex[3]) { float4 t = to_linear(sample_mat(d.tex[3], in)); emissive = emissive * mk3(t.x, t.y, t.z); }

    f3 color = mk3(0.f, 0.f, 0.f);
    f3 bc = mk3(baseColor.x, baseColor.y, baseColor.z);
    f3 c_diff = bc * (1.0f - 0.04f) * (1.0f - metallic);
    f3 F0 = mix3(mk3(0.04f, 0.04f, 0.04f), bc, metallic);
    f3 Fr = max3(mk3(1.0f - roughness, 1.0f - roughness, 1.0f - roughness), F0) - F0;
    const float omv = 1.0f - NoV, omv2 = omv * omv;
    f3 k_S = F0 + Fr * (omv2 * omv2 * omv);

#pragma unroll 1
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i) {
        if (!f.lightActive[i]) continue;
        float4 pc = mul_m4_p(f.shadowMat[i], in.wc.x, in.wc.y, in.wc.z, 1.0f);
        float pcx = 0.5f * (pc.x / pc.w) + 0.5f, pcy = 0.5f * (pc.y / pc.w) + 0.5f, pcz = 0.5f * (pc.z / pc.w) + 0.5f;
        const float inverseShadow = shadow_pcf16(f.shadowMap[i], f.shadowMask[i], pcx, pcy, shadow_threshold(pcz - 0.00003f), f.shadow_tagbits);

        f3 L = normalize3(mk3(-f.lightDir[i][0], -f.lightDir[i][1], -f.lightDir[i][2]));
        f3 H = normalize3(cameraDirection + L);
        f3 radiance = mk3(f.lightCol[i][0], f.lightCol[i][1], f.lightCol[i][2]);
        float NDF = DistributionGGX(normal, H, roughness);
        float NdotL = fmaxf(dot3(normal, L), 0.0f);
        float G = GeometrySchlickGGX(NdotL, roughness) * GeometrySchlickGGX(fmaxf(dot3(normal, cameraDirection), 0.0f), roughness);
        f3 nominator = k_S * (NDF * G);
        float denominator = 4.0f * NoV * NdotL;
        f3 specular = nominator / fmaxf(denominator, 0.001f);
        f3 kD = (mk3(1.f, 1.f, 1.f) - k_S) * (1.0f - metallic);
        color = color + (kD * bc / PI + specular) * radiance * (inverseShadow * NdotL);
    }
    color = color + mk3(f.ambient[0], f.ambient[1], f.ambient[2]) * bc;

    if (!LEAN && f.lm) {
        const DLightMap& lm = *f.lm;
        float4 fab = lut_sample(lm, NoV, roughness);
        float4 rad = cube_sample_lod(lm.pre, 5, reflDir, roughness * 4.0f);
        float4 irr = cube_sample_lod(&lm.irr, 1, normal, 0.0f);
        f3 radiance = mk3(rad.x, rad.y, rad.z), irradiance = mk3(irr.x, irr.y, irr.z);
        f3 one = mk3(1.f, 1.f, 1.f);
        f3 FssEss = k_S * fab.x + mk3(fab.y, fab.y, fab.y);
        float Ems = 1.0f - (fab.x + fab.y);
        f3 F_avg = F0 + (one - F0) / 21.0f;
        f3 FmsEms = FssEss * F_avg * Ems / (one - F_avg * Ems);
        f3 k_D = c_diff * (one - FssEss - FmsEms);
        f3 selfColor = FssEss * radiance + (FmsEms + k_D) * irradiance;
        color = color + selfColor * occlusion;
    }
    color = color + emissive;

    out_color = make_float4(color.x, color.y, color.z, baseColor.w);
    f3 nc = mk3(f.V[0] * normal.x + f.V[4] * normal.y + f.V[8] * normal.z, f.V[1] * normal.x + f.V[5] * normal.y + f.V[9] * normal.z,
                f.V[2] * normal.x + f.V[6] * normal.y + f.V[10] * normal.z);
    nc = normalize3(nc);
    out_normal = make_float4(nc.x, nc.y, nc.z, dot3(normal, cameraDirection));
}

// tone map (tone_map_shader.frag:102-131): RGB -> Yxy, exposure, -> RGB, ACES, RGBA8 (linear: the
// gamma line of the reference is overwritten)
__device__ __forceinline__ float aces1(float x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    float v = (x * (a * x + b)) / (x * (c * x + d) + e);
    return (v != v) ? 0.0f : clampf(v, 0.0f, 1.0f);
}
__device__ __forceinline__ unsigned unorm8(float v) {
    if (!(v == v)) return 0u;
    return (unsigned)__float2int_rn(clampf(v, 0.0f, 1.0f) * 255.0f);
}
__device__ __forceinline__ uchar4 tone_map(float4 hdr, float manual_exposure, const float* avg) {
    float X = 0.4124564f * hdr.x + 0.3575761f * hdr.y + 0.1804375f * hdr.z;
    float Y = 0.2126729f * hdr.x + 0.7151522f * hdr.y + 0.0721750f * hdr.z;
    float Z = 0.0193339f * hdr.x + 0.1191920f * hdr.y + 0.9503041f * hdr.z;
    float inv = 1.0f / (X + Y + Z);
    float yY = Y, yx = X * inv, yy = Y * inv;
    if (manual_exposure >= 0) yY *= manual_exposure;
    else {
        float lum = 0.1f * (0.2125f * (avg[0] / avg[3]) + 0.7154f * (avg[1] / avg[3]) + 0.0721f * (avg[2] / avg[3]));
        yY /= (9.6f * lum + 0.0001f);
    }
    float x2 = yY * yx / yy, y2 = yY, z2 = yY * (1.0f - yx - yy) / yy;
    float r = 3.2404542f * x2 - 1.5371385f * y2 - 0.4985314f * z2;
    float g = -0.9692660f * x2 + 1.8760108f * y2 + 0.0415560f * z2;
    float b = 0.0556434f * x2 - 0.2040259f * y2 + 1.0572252f * z2;
    return make_uchar4((unsigned char)unorm8(aces1(r)), (unsigned char)unorm8(aces1(g)), (unsigned char)unorm8(aces1(b)),
                       (unsigned char)unorm8(hdr.w));
}

}  // namespace slbk
