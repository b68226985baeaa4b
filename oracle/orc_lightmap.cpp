// orc_lightmap.cpp — ORACLE (test infrastructure only; see orc_core.h for the rules).
//
// CPU restatement of the load-time IBL precompute LightMap::load() runs on the GPU
// (reference: src/light_map.cpp:376-611) and of the four GLSL programs it uses:
//   equirect -> cube      src/shaders/cubemap_shader_equirectangular.frag:10-27
//   irradiance            src/shaders/cubemap_shader_irradiance.frag:14-45
//   GGX prefilter         src/shaders/cubemap_shader_prefilter.frag:77-117
//   BRDF LUT              src/shaders/brdf_shader.frag:74-118
// The reference renders a unit cube seen from the origin through the six CUBE_MAP_SIDES views
// (light_map.cpp:185-192) with a 90 degree projection; the fragment at face texel (i,j) therefore
// receives WorldPos = the point of the cube face whose GL cube-map coordinates are
// ((i+0.5)/n, (j+0.5)/n) — i.e. rendering into a face and sampling that face are inverse maps, so
// no rasterisation is needed here: every face texel evaluates the shader at cube_face_dir().
//
// Documented deviations (driver-defined behaviour in the reference):
//  * texture(equirectangularMap, uv) uses implicit LOD + max anisotropy on a mip-chained RGB32F
//    image (light_map.cpp:168-174); restated as level-0 bilinear, clamp-to-edge.
//  * texture(environmentMap, sampleVec) in the irradiance loop uses implicit derivatives between
//    neighbouring fragments; restated as the constant LOD log2(env_size / irradiance_size).
#include "orc_core.h"

#include <omp.h>

namespace orc {

static const float PI = 3.14159265359f;  // the GLSL constant, not M_PI

static inline V3 face_dir(int face, float s, float t) {
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

static V3 equirect_sample(const float* img, int W, int H, float u, float v) {
    float x = u * W - 0.5f, y = v * H - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    auto at = [&](int i, int j) {
        i = std::min(std::max(i, 0), W - 1); j = std::min(std::max(j, 0), H - 1);
        const float* p = img + ((size_t)j * W + i) * 3; return V3(p[0], p[1], p[2]);
    };
    return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) +
           at(i0 + 1, j0 + 1) * (a * b);
}

static float radical_inverse(uint32_t bits) {
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
static V3 importance_sample_ggx(float xi_x, float xi_y, V3 N, float roughness) {
    float a = roughness * roughness;
    float phi = 2.0f * PI * xi_x;
    float cosTheta = std::sqrt((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
    V3 H(std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, cosTheta);
    V3 up = std::fabs(N.z) < 0.999f ? V3(0, 0, 1) : V3(1, 0, 0);
    V3 tangent = normalize(cross(up, N));
    V3 bitangent = cross(N, tangent);
    return normalize(tangent * H.x + bitangent * H.y + N * H.z);
}
static float distribution_ggx(V3 N, V3 H, float roughness) {
    float a = roughness * roughness, a2 = a * a;
    float NdotH = std::max(dot(N, H), 0.0f), NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = PI * denom * denom;
    return a2 / denom;
}
static float geometry_schlick_ibl(float NdotV, float roughness) {
    float k = (roughness * roughness) / 2.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}

// ---- the four fragment programs, one output texel each (WorldPos = un-normalised cube-face position) ----
static V3 equirect_texel(const slb_lightmap_desc* d, V3 world_pos) {      // cubemap_shader_equirectangular.frag:10-27
    V3 v = normalize(world_pos);
    float u = std::atan2(v.y, v.x) * 0.1591f + 0.5f;
    float w = std::asin(v.z) * 0.3183f + 0.5f;
    return equirect_sample(d->equirect_rgb, d->width, d->height, u, w);
}
static V3 irradiance_texel(const std::vector<CubeLevel>& env, V3 world_pos, float lod) {   // cubemap_shader_irradiance.frag:14-45
    V3 N = normalize(world_pos);
    V3 irradiance(0.0f);
    V3 up(0, 1, 0);
    V3 right = cross(up, N);
    up = cross(N, right);
    const float sampleDelta = 0.020f;
    float nrSamples = 0.0f;
    for (float phi = 0.0f; phi < 2.0f * PI; phi += sampleDelta)
        for (float theta = 0.0f; theta < 0.5f * PI; theta += sampleDelta) {
            V3 ts(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
            V3 sv = right * ts.x + up * ts.y + N * ts.z;
            V4 c = sample_cube_lod(env, sv, lod);
            irradiance += V3(c.x, c.y, c.z) * std::cos(theta) * std::sin(theta);
            nrSamples += 1.0f;
        }
    return irradiance * PI * (1.0f / nrSamples);
}
static V3 prefilter_texel(const std::vector<CubeLevel>& env, V3 world_pos, float roughness, int n_samples, int env_size) {   // cubemap_shader_prefilter.frag:77-114
    V3 N = normalize(world_pos);
    V3 R = N, V = R;
    V3 color(0.0f);
    float totalWeight = 0.0f;
    for (uint32_t i = 0; i < (uint32_t)n_samples; ++i) {
        float xi_x = (float)i / (float)n_samples, xi_y = radical_inverse(i);
        V3 H = importance_sample_ggx(xi_x, xi_y, N, roughness);
        V3 L = normalize(H * (2.0f * dot(V, H)) - V);
        float NdotL = std::max(dot(N, L), 0.0f);
        if (NdotL > 0.0f) {
            float D = distribution_ggx(N, H, roughness);
            float NdotH = std::max(dot(N, H), 0.0f);
            float HdotV = std::max(dot(H, V), 0.0f);
            float pdf = D * NdotH / (4.0f * HdotV) + 0.0001f;
            float resolution = (float)env_size;
            float saTexel = 4.0f * PI / (6.0f * resolution * resolution);
            float saSample = 1.0f / ((float)n_samples * pdf + 0.0001f);
            float mipLevel = roughness == 0.0f ? 0.0f : 0.5f * std::log2(saSample / saTexel);
            V4 c = sample_cube_lod(env, L, mipLevel);
            color += V3(c.x, c.y, c.z) * NdotL;
            totalWeight += NdotL;
        }
    }
    return color / totalWeight;
}
static void brdf_texel(float NdotV, float roughness, int n_samples, float& A, float& B) {   // brdf_shader.frag:74-118
    V3 V(std::sqrt(1.0f - NdotV * NdotV), 0.0f, NdotV);
    A = 0.0f; B = 0.0f;
    V3 N(0, 0, 1);
    for (uint32_t i = 0; i < (uint32_t)n_samples; ++i) {
        float xi_x = (float)i / (float)n_samples, xi_y = radical_inverse(i);
        V3 H = importance_sample_ggx(xi_x, xi_y, N, roughness);
        V3 L = normalize(H * (2.0f * dot(V, H)) - V);
        float NdotL = std::max(L.z, 0.0f), NdotH = std::max(H.z, 0.0f), VdotH = std::max(dot(V, H), 0.0f);
        if (NdotL > 0.0f) {
            float G = geometry_schlick_ibl(std::max(dot(N, L), 0.0f), roughness) *
                      geometry_schlick_ibl(std::max(dot(N, V), 0.0f), roughness);
            float G_Vis = (G * VdotH) / (NdotH * NdotV);
            float Fc = std::pow(1.0f - VdotH, 5.0f);
            A += (1.0f - Fc) * G_Vis;
            B += Fc * G_Vis;
        }
    }
    A /= (float)n_samples; B /= (float)n_samples;
}

LightMap* lightmap_create(const slb_lightmap_desc* d, int env_size, int irr_size, int pre_size, int lut_size,
                          int n_samples) {
    LightMap* lm = new LightMap;
    lm->n_lights = std::min(std::max(d->n_lights, 0), SLB_NUM_LIGHTS);
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i)
        for (int k = 0; k < 3; ++k) {
            lm->light_directions[i][k] = d->light_directions[i][k];
            lm->light_colors[i][k] = d->light_colors[i][k];
        }

    // ---- equirect -> cube (light_map.cpp:394-429) + full mip chain ----
    {
        CubeLevel l0; l0.size = env_size; l0.px.resize((size_t)6 * env_size * env_size * 4);
        #pragma omp parallel for collapse(2)
        for (int f = 0; f < 6; ++f)
            for (int y = 0; y < env_size; ++y)
                for (int x = 0; x < env_size; ++x) {
                    V3 c = equirect_texel(d, face_dir(f, (x + 0.5f) / env_size, (y + 0.5f) / env_size));
                    float* p = &l0.px[(((size_t)f * env_size + y) * env_size + x) * 4];
                    p[0] = c.x; p[1] = c.y; p[2] = c.z; p[3] = 1.0f;
                }
        lm->env.push_back(std::move(l0));
        build_cube_mips(lm->env);
    }

    // ---- irradiance (light_map.cpp:449-505) ----
    {
        CubeLevel l0; l0.size = irr_size; l0.px.resize((size_t)6 * irr_size * irr_size * 4);
        const float lod = std::log2((float)env_size / (float)irr_size);
        #pragma omp parallel for collapse(2) schedule(dynamic, 4)
        for (int f = 0; f < 6; ++f)
            for (int y = 0; y < irr_size; ++y)
                for (int x = 0; x < irr_size; ++x) {
                    V3 irradiance = irradiance_texel(lm->env, face_dir(f, (x + 0.5f) / irr_size, (y + 0.5f) / irr_size), lod);
                    float* p = &l0.px[(((size_t)f * irr_size + y) * irr_size + x) * 4];
                    p[0] = irradiance.x; p[1] = irradiance.y; p[2] = irradiance.z; p[3] = 1.0f;
                }
        lm->irradiance.push_back(std::move(l0));
    }

    // ---- GGX prefilter, 5 mips, roughness = mip / 4 (light_map.cpp:507-575) ----
    {
        const int MAX_MIP_LEVELS = 5;
        for (int mip = 0; mip < MAX_MIP_LEVELS; ++mip) {
            int n = std::max(1, (int)(pre_size * std::pow(0.5f, (float)mip)));
            float roughness = (float)mip / (float)(MAX_MIP_LEVELS - 1);
            CubeLevel l; l.size = n; l.px.resize((size_t)6 * n * n * 4);
            #pragma omp parallel for collapse(2) schedule(dynamic, 4)
            for (int f = 0; f < 6; ++f)
                for (int y = 0; y < n; ++y)
                    for (int x = 0; x < n; ++x) {
                        V3 color = prefilter_texel(lm->env, face_dir(f, (x + 0.5f) / n, (y + 0.5f) / n), roughness, n_samples, env_size);
                        float* p = &l.px[(((size_t)f * n + y) * n + x) * 4];
                        p[0] = color.x; p[1] = color.y; p[2] = color.z; p[3] = 1.0f;
                    }
            lm->prefilter.push_back(std::move(l));
        }
    }

    // ---- BRDF LUT (light_map.cpp:577-602, brdf_shader.frag:74-118) ----
    {
        lm->lut_size = lut_size;
        std::vector<float> l0((size_t)lut_size * lut_size * 4);
        #pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < lut_size; ++y)
            for (int x = 0; x < lut_size; ++x) {
                float NdotV = (x + 0.5f) / lut_size, roughness = (y + 0.5f) / lut_size;
                float A, B; brdf_texel(NdotV, roughness, n_samples, A, B);
                float* p = &l0[((size_t)y * lut_size + x) * 4];
                p[0] = A; p[1] = B; p[2] = 0.0f; p[3] = 1.0f;
            }
        lm->lut.push_back(std::move(l0));
    }
    return lm;
}

}  // namespace orc

using namespace orc;

extern "C" {

// sizes <= 0 select the reference's: env 512, irradiance 32, prefilter 128, LUT 512, 1024 samples
void* orc_lightmap_create(const slb_lightmap_desc* d, int env_size, int irr_size, int pre_size, int lut_size, int n_samples) {
    return lightmap_create(d, env_size > 0 ? env_size : 512, irr_size > 0 ? irr_size : 32, pre_size > 0 ? pre_size : 128,
                           lut_size > 0 ? lut_size : 512, n_samples > 0 ? n_samples : 1024);
}
// Build an oracle light map from precomputed maps (lets the renderer be checked independently of
// the precompute: both sides are fed the SAME maps). env0: [6][e][e][4]; irr: [6][i][i][4];
// pre: 5 levels packed; lut: [l][l][4].
void* orc_lightmap_from_maps(const slb_lightmap_desc* d, const float* env0, int env_size, const float* irr, int irr_size,
                             const float* pre, int pre_size, const float* lut, int lut_size) {
    LightMap* lm = new LightMap;
    lm->n_lights = std::min(std::max(d->n_lights, 0), SLB_NUM_LIGHTS);
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i)
        for (int k = 0; k < 3; ++k) { lm->light_directions[i][k] = d->light_directions[i][k]; lm->light_colors[i][k] = d->light_colors[i][k]; }
    CubeLevel e; e.size = env_size; e.px.assign(env0, env0 + (size_t)6 * env_size * env_size * 4);
    lm->env.push_back(std::move(e)); build_cube_mips(lm->env);
    CubeLevel ir; ir.size = irr_size; ir.px.assign(irr, irr + (size_t)6 * irr_size * irr_size * 4);
    lm->irradiance.push_back(std::move(ir));
    const float* p = pre;
    for (int mip = 0; mip < 5; ++mip) {
        int n = std::max(1, pre_size >> mip);
        CubeLevel l; l.size = n; l.px.assign(p, p + (size_t)6 * n * n * 4); p += (size_t)6 * n * n * 4;
        lm->prefilter.push_back(std::move(l));
    }
    lm->lut_size = lut_size;
    lm->lut.push_back(std::vector<float>(lut, lut + (size_t)lut_size * lut_size * 4));
    return lm;
}
// which: 0 env level 0, 1 irradiance, 2 prefilter (all mips packed), 3 LUT. Returns #floats (out may be NULL).
size_t orc_lightmap_read(const void* h, int which, float* out) {
    const LightMap* lm = (const LightMap*)h;
    std::vector<float> tmp;
    const std::vector<float>* src = nullptr;
    if (which == 0) src = &lm->env[0].px;
    else if (which == 1) src = &lm->irradiance[0].px;
    else if (which == 2) { for (const CubeLevel& l : lm->prefilter) tmp.insert(tmp.end(), l.px.begin(), l.px.end()); src = &tmp; }
    else if (which == 3) src = &lm->lut[0];
    if (!src) return 0;
    if (out) std::memcpy(out, src->data(), src->size() * sizeof(float));
    return src->size();
}
void orc_lightmap_sizes(const void* h, int sizes[4]) {
    const LightMap* lm = (const LightMap*)h;
    sizes[0] = lm->env[0].size; sizes[1] = lm->irradiance[0].size; sizes[2] = lm->prefilter[0].size; sizes[3] = lm->lut_size;
}
void orc_lightmap_destroy(void* h) { delete (LightMap*)h; }

// per-texel test hooks (see orc_test_hooks.h): which = 0 equirect -> cube, 1 irradiance, 2 prefilter; world_pos: n x 3
// cube-face positions; `lm` supplies the environment cube for 1 and 2, `d` the equirect image for 0
void orc_test_lightmap_texels(int which, const slb_lightmap_desc* d, const void* lm_h, const float* world_pos, int n, float roughness,
                              int n_samples, float irradiance_lod, float* rgb_out) {
    const LightMap* lm = (const LightMap*)lm_h;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
        V3 wp(world_pos[3 * i], world_pos[3 * i + 1], world_pos[3 * i + 2]), c;
        if (which == 0) c = equirect_texel(d, wp);
        else if (which == 1) c = irradiance_texel(lm->env, wp, irradiance_lod);
        else c = prefilter_texel(lm->env, wp, roughness, n_samples, lm->env[0].size);
        rgb_out[3 * i] = c.x; rgb_out[3 * i + 1] = c.y; rgb_out[3 * i + 2] = c.z;
    }
}
void orc_test_brdf_lut(const float* ndotv_roughness, int n, int n_samples, float* ab_out) {
    for (int i = 0; i < n; ++i) brdf_texel(ndotv_roughness[2 * i], ndotv_roughness[2 * i + 1], n_samples, ab_out[2 * i], ab_out[2 * i + 1]);
}
// filtered cube look-up of an oracle light map: which = 0 env, 1 irradiance, 2 prefilter (the harness binds the compiled
// reference shaders' samplerCube to this)
void orc_test_cube_sample(const void* lm_h, int which, const float* dir, float lod, float* rgba) {
    const LightMap* lm = (const LightMap*)lm_h;
    const std::vector<CubeLevel>& cube = which == 0 ? lm->env : (which == 1 ? lm->irradiance : lm->prefilter);
    V4 c = sample_cube_lod(cube, V3(dir[0], dir[1], dir[2]), lod);
    rgba[0] = c.x; rgba[1] = c.y; rgba[2] = c.z; rgba[3] = c.w;
}

}  // extern "C"
