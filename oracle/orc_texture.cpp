// orc_texture.cpp — ORACLE (test infrastructure): texture storage, mip chains and GL-style
// sampling written out explicitly (SURVEY §8c "bilinear/trilinear/cube sampling written out").
//
// Follows: GL 4.5 core spec §8.14 (texture minification / magnification, LOD selection),
// the reference's texture set-up in src/mesh.cpp:644-663 (RGB8/RGBA8, glTF sampler filters and
// wrap, glGenerateMipmap), src/context.cpp:560-640 (rect textures: clamp-to-border transparent),
// src/light_map.cpp:379 (seamless cube filtering).
#include "orc_core.h"

namespace orc {

// One step of the mip chain. For even sizes this is the 2x2 box glGenerateMipmap uses on
// power-of-two textures; for odd sizes it is the polyphase box (weights d-i, d, i+1 over
// three taps) so every source texel carries the same total weight. Integer arithmetic with
// round-half-up so that the CUDA implementation can be bit-identical.
static void taps(int s, int d, int i, int idx[3], int wgt[3], int& total) {
    if (s == 1) { idx[0] = idx[1] = idx[2] = 0; wgt[0] = 1; wgt[1] = 0; wgt[2] = 0; total = 1; return; }
    if ((s & 1) == 0) { idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 1; wgt[0] = 1; wgt[1] = 1; wgt[2] = 0; total = 2; return; }
    idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 2;
    wgt[0] = d - i; wgt[1] = d; wgt[2] = i + 1; total = s;
}

void build_texture(Texture& t, const slb_image* img, int kind) {
    t.kind = kind;
    t.w = img->width; t.h = img->height; t.ch = img->channels;
    t.wrap_s = img->wrap_s; t.wrap_t = img->wrap_t;
    t.min_filter = img->min_filter; t.mag_filter = img->mag_filter;
    t.has_alpha = (img->channels == 4);
    Texture::Level l0; l0.w = t.w; l0.h = t.h; l0.px.resize((size_t)t.w * t.h * 4);
    const uint8_t* src = (const uint8_t*)img->pixels;
    for (size_t i = 0; i < (size_t)t.w * t.h; ++i) {
        for (int c = 0; c < 3; ++c) l0.px[i * 4 + c] = src[i * t.ch + c];
        l0.px[i * 4 + 3] = (t.ch == 4) ? src[i * 4 + 3] : 255;
    }
    t.levels.clear();
    t.levels.push_back(std::move(l0));
    if (kind != SLB_TEXTURE_2D) return;
    while (t.levels.back().w > 1 || t.levels.back().h > 1) {
        const Texture::Level& s = t.levels.back();
        Texture::Level d; d.w = std::max(1, s.w >> 1); d.h = std::max(1, s.h >> 1);
        d.px.resize((size_t)d.w * d.h * 4);
        for (int y = 0; y < d.h; ++y) {
            int iy[3], wy[3], ty; taps(s.h, d.h, y, iy, wy, ty);
            for (int x = 0; x < d.w; ++x) {
                int ix[3], wx[3], tx; taps(s.w, d.w, x, ix, wx, tx);
                for (int c = 0; c < 4; ++c) {
                    uint32_t acc = 0;
                    for (int b = 0; b < 3; ++b)
                        for (int a = 0; a < 3; ++a)
                            acc += (uint32_t)(wy[b] * wx[a]) * s.px[((size_t)iy[b] * s.w + ix[a]) * 4 + c];
                    uint32_t tot = (uint32_t)(tx * ty);
                    d.px[((size_t)y * d.w + x) * 4 + c] = (uint8_t)((acc + tot / 2) / tot);
                }
            }
        }
        t.levels.push_back(std::move(d));
    }
}

static inline int wrap_index(int i, int n, int mode, bool& border) {
    switch (mode) {
        case SLB_WRAP_REPEAT: { int m = i % n; return m < 0 ? m + n : m; }
        case SLB_WRAP_MIRRORED_REPEAT: {
            int p = 2 * n; int m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m;
        }
        case SLB_WRAP_CLAMP_TO_BORDER:
            if (i < 0 || i >= n) { border = true; return 0; }
            return i;
        default: return std::min(std::max(i, 0), n - 1);
    }
}

static inline V4 texel(const Texture::Level& l, int x, int y) {
    const uint8_t* p = &l.px[((size_t)y * l.w + x) * 4];
    return V4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
}

static V4 fetch_wrapped(const Texture& t, const Texture::Level& l, int x, int y) {
    bool border = false;
    int xi = wrap_index(x, l.w, t.wrap_s, border);
    int yi = wrap_index(y, l.h, t.wrap_t, border);
    if (border) return V4(0, 0, 0, 0);
    return texel(l, xi, yi);
}

// sample one level with texel-space coordinates (x = u*w for normalised textures)
static V4 sample_level(const Texture& t, int level, float xs, float ys, bool linear) {
    const Texture::Level& l = t.levels[level];
    if (!linear) return fetch_wrapped(t, l, (int)std::floor(xs), (int)std::floor(ys));
    float x = xs - 0.5f, y = ys - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    V4 t00 = fetch_wrapped(t, l, i0, j0), t10 = fetch_wrapped(t, l, i0 + 1, j0);
    V4 t01 = fetch_wrapped(t, l, i0, j0 + 1), t11 = fetch_wrapped(t, l, i0 + 1, j0 + 1);
    return t00 * ((1 - a) * (1 - b)) + t10 * (a * (1 - b)) + t01 * ((1 - a) * b) + t11 * (a * b);
}

V4 sample_texture_2d(const Texture& t, float u, float v, float dudx, float dvdx, float dudy, float dvdy) {
    const int max_level = (int)t.levels.size() - 1;
    float W = (float)t.w, H = (float)t.h;
    float rx = std::sqrt(dudx * W * dudx * W + dvdx * H * dvdx * H);
    float ry = std::sqrt(dudy * W * dudy * W + dvdy * H * dvdy * H);
    float rho = std::max(rx, ry);
    float lambda = std::log2(rho);  // rho == 0 -> -inf -> magnification
    bool mag_linear = (t.mag_filter == SLB_FILTER_LINEAR);
    float c = (mag_linear && (t.min_filter == SLB_FILTER_NEAREST_MIPMAP_NEAREST ||
                              t.min_filter == SLB_FILTER_NEAREST_MIPMAP_LINEAR)) ? 0.5f : 0.0f;
    if (!(lambda > c)) return sample_level(t, 0, u * W, v * H, mag_linear);
    int mf = t.min_filter;
    bool lin = (mf == SLB_FILTER_LINEAR || mf == SLB_FILTER_LINEAR_MIPMAP_NEAREST || mf == SLB_FILTER_LINEAR_MIPMAP_LINEAR);
    if (mf == SLB_FILTER_NEAREST || mf == SLB_FILTER_LINEAR) return sample_level(t, 0, u * W, v * H, lin);
    if (mf == SLB_FILTER_NEAREST_MIPMAP_NEAREST || mf == SLB_FILTER_LINEAR_MIPMAP_NEAREST) {
        int d = (lambda <= 0.5f) ? 0 : std::min((int)std::ceil(lambda + 0.5f) - 1, max_level);
        const Texture::Level& l = t.levels[d];
        return sample_level(t, d, u * l.w, v * l.h, lin);
    }
    float lc = std::min(lambda, (float)max_level);
    int d1 = (int)std::floor(lc);
    int d2 = std::min(d1 + 1, max_level);
    float f = lc - (float)d1;
    const Texture::Level& l1 = t.levels[d1];
    V4 s1 = sample_level(t, d1, u * l1.w, v * l1.h, lin);
    if (d2 == d1 || f == 0.0f) return s1;
    const Texture::Level& l2 = t.levels[d2];
    V4 s2 = sample_level(t, d2, u * l2.w, v * l2.h, lin);
    return s1 * (1.0f - f) + s2 * f;
}

V4 sample_texture_rect_linear(const Texture& t, float x, float y) { return sample_level(t, 0, x, y, t.mag_filter == SLB_FILTER_LINEAR); }
V4 sample_texture_rect_nearest(const Texture& t, float x, float y) { return sample_level(t, 0, x, y, false); }

// ---- cube maps --------------------------------------------------------------------------
// Face order +X,-X,+Y,-Y,+Z,-Z and (sc,tc,ma) selection: GL 4.5 spec table 8.19.
static inline void cube_face_coords(V3 d, int& face, float& s, float& t) {
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { if (d.x >= 0) { face = 0; sc = -d.z; tc = -d.y; } else { face = 1; sc = d.z; tc = -d.y; } ma = ax; }
    else if (ay >= az)        { if (d.y >= 0) { face = 2; sc = d.x;  tc = d.z;  } else { face = 3; sc = d.x; tc = -d.z; } ma = ay; }
    else                      { if (d.z >= 0) { face = 4; sc = d.x;  tc = -d.y; } else { face = 5; sc = -d.x; tc = -d.y; } ma = az; }
    s = 0.5f * (sc / ma + 1.0f);
    t = 0.5f * (tc / ma + 1.0f);
}
// inverse: direction through the point (s,t) in [0,1]^2 (may lie outside for seamless taps)
static inline V3 cube_face_dir(int face, float s, float t) {
    float a = 2.0f * s - 1.0f, b = 2.0f * t - 1.0f;
    switch (face) {
        case 0: return V3(1, -b, -a);
        case 1: return V3(-1, -b, a);
        case 2: return V3(a, 1, b);
        case 3: return V3(a, -1, -b);
        case 4: return V3(a, -b, 1);
        default: return V3(-a, -b, -1);
    }
}
static inline V4 cube_texel(const CubeLevel& l, int face, int x, int y) {
    const float* p = &l.px[(((size_t)face * l.size + y) * l.size + x) * 4];
    return V4(p[0], p[1], p[2], p[3]);
}
// Seamless tap: a texel index outside the face is re-projected through its centre direction
// onto the neighbouring face and the nearest texel there is used (corner taps, where two
// indices are out of range, land on one of the adjacent faces — GL averages the three corner
// texels instead; documented deviation, affects 24 texel corners per level).
static V4 cube_tap(const CubeLevel& l, int face, int x, int y) {
    int n = l.size;
    if (x >= 0 && x < n && y >= 0 && y < n) return cube_texel(l, face, x, y);
    V3 d = cube_face_dir(face, (x + 0.5f) / n, (y + 0.5f) / n);
    int f2; float s, t; cube_face_coords(d, f2, s, t);
    int xi = std::min(std::max((int)std::floor(s * n), 0), n - 1);
    int yi = std::min(std::max((int)std::floor(t * n), 0), n - 1);
    return cube_texel(l, f2, xi, yi);
}
static V4 sample_cube_level(const CubeLevel& l, V3 dir) {
    int face; float s, t; cube_face_coords(dir, face, s, t);
    float x = s * l.size - 0.5f, y = t * l.size - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    V4 t00 = cube_tap(l, face, i0, j0), t10 = cube_tap(l, face, i0 + 1, j0);
    V4 t01 = cube_tap(l, face, i0, j0 + 1), t11 = cube_tap(l, face, i0 + 1, j0 + 1);
    return t00 * ((1 - a) * (1 - b)) + t10 * (a * (1 - b)) + t01 * ((1 - a) * b) + t11 * (a * b);
}
V4 sample_cube_lod(const std::vector<CubeLevel>& cube, V3 dir, float lod) {
    int max_level = (int)cube.size() - 1;
    float lc = std::min(std::max(lod, 0.0f), (float)max_level);
    int d1 = (int)std::floor(lc);
    int d2 = std::min(d1 + 1, max_level);
    float f = lc - (float)d1;
    V4 s1 = sample_cube_level(cube[d1], dir);
    if (d2 == d1 || f == 0.0f) return s1;
    V4 s2 = sample_cube_level(cube[d2], dir);
    return s1 * (1.0f - f) + s2 * f;
}
void build_cube_mips(std::vector<CubeLevel>& cube) {
    while (cube.back().size > 1) {
        const CubeLevel& s = cube.back();
        CubeLevel d; d.size = s.size / 2; d.px.resize((size_t)6 * d.size * d.size * 4);
        for (int f = 0; f < 6; ++f)
            for (int y = 0; y < d.size; ++y)
                for (int x = 0; x < d.size; ++x)
                    for (int c = 0; c < 4; ++c) {
                        auto at = [&](int xx, int yy) { return s.px[(((size_t)f * s.size + yy) * s.size + xx) * 4 + c]; };
                        d.px[(((size_t)f * d.size + y) * d.size + x) * 4 + c] =
                            0.25f * ((at(2 * x, 2 * y) + at(2 * x + 1, 2 * y)) + (at(2 * x, 2 * y + 1) + at(2 * x + 1, 2 * y + 1)));
                    }
        cube.push_back(std::move(d));
    }
}

// BRDF LUT: texture2D(lightMapBRDFLUT, vec2(NoV, roughness)) (render_shader.frag:378). The LUT is
// a smooth 512^2 function with clamp-to-edge + linear filtering (light_map.cpp:577-602); the
// implicit-LOD mip selection is replaced by level-0 bilinear (documented deviation: the LUT is
// band-limited far below one texel, so box-filtered mips differ by < 1e-4).
V4 sample_lut(const LightMap& lm, float u, float v) {
    int n = lm.lut_size;
    const std::vector<float>& l = lm.lut[0];
    float x = u * n - 0.5f, y = v * n - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    auto at = [&](int xi, int yi) {
        xi = std::min(std::max(xi, 0), n - 1); yi = std::min(std::max(yi, 0), n - 1);
        const float* p = &l[((size_t)yi * n + xi) * 4];
        return V4(p[0], p[1], p[2], p[3]);
    };
    int i0 = (int)fx, j0 = (int)fy;
    return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) +
           at(i0 + 1, j0 + 1) * (a * b);
}

}  // namespace orc
