// glslref_harness.cpp — runs the REFERENCE'S OWN SHADER TEXT on the CPU (test infrastructure only).
//
// oracle/build_ref.py translates the GLSL files under /root/reference/src/shaders with oracle/glsl_to_cpp.py (syntax-only
// edits, see there) into oracle/_ref/gen/*.inc; each is included below as the body of a struct, so uniforms / inputs /
// outputs are its data members and main() a member function, compiled against oracle/glsl_compat.h. This file binds the
// samplers to the oracle's texture units (texture filtering is GL fixed function: oracle/orc_texture.cpp) and exposes one
// `glslref_*` entry point per program with the same argument structs as the oracle's `orc_test_*` hooks
// (oracle/orc_test_hooks.h). Built together with the oracle sources into oracle/_ref/libglslref.so.
#include <cstring>
#include <random>
#include <vector>

#include "glsl_compat.h"
#include "orc_core.h"
#include "orc_test_hooks.h"

extern "C" void orc_test_cube_sample(const void* lm_h, int which, const float* dir, float lod, float* rgba);

namespace glsl {

#undef M_PI
#define M_PI 3.141592653589793f
#define NUM_LIGHTS 3
// TextureInput of src/shaders/render_shader.cpp:34-47
#define BASE_COLOR_TEXTURE 0
#define NORMAL_TEXTURE 1
#define METALLIC_ROUGHNESS_TEXTURE 2
#define EMISSIVE_TEXTURE 3
#define OCCLUSION_TEXTURE 4
#define discard { _discarded = true; return; }

struct RenderVert {
    vec4 gl_Position;
#include "_ref/gen/render_shader.glsl.inc"
#include "_ref/gen/render_shader.vert.inc"
};
struct RenderFrag {
    vec4 gl_FragCoord; bool gl_FrontFacing = true; bool _discarded = false;
#include "_ref/gen/render_shader.glsl.inc"
#include "_ref/gen/render_shader.frag.inc"
};
#undef baseColorFactor
#undef emissiveFactor
#undef alphaCutoff
#undef metallicFactor
#undef roughnessFactor
#undef DIELECTRIC_SPECULAR
#undef MIN_ROUGHNESS
struct ToneMapFrag {
    vec4 gl_FragCoord;
#include "_ref/gen/tone_map_shader.frag.inc"
};
#undef RGB_TO_LUM
struct SsaoFrag {
    vec4 gl_FragCoord;
#include "_ref/gen/ssao_shader.frag.inc"
};
struct SsaoApplyFrag {
    vec4 gl_FragCoord;
#include "_ref/gen/ssao_apply_shader.frag.inc"
};
struct BackgroundVert {
    vec4 gl_Position;
#include "_ref/gen/background_shader.vert.inc"
};
struct BackgroundFrag {
#include "_ref/gen/background_shader.frag.inc"
};
struct BackgroundCubeVert {
    vec4 gl_Position;
#include "_ref/gen/background_cube_shader.vert.inc"
};
struct BackgroundCubeFrag {
#include "_ref/gen/background_cube_shader.frag.inc"
};
struct ShadowVert {
    vec4 gl_Position;
#define UNIFORM_TRANSFORMATION 0
#include "_ref/gen/shadow_shader.vert.inc"
};
struct EquirectFrag {
#include "_ref/gen/cubemap_shader_equirectangular.frag.inc"
};
struct IrradianceFrag {
#include "_ref/gen/cubemap_shader_irradiance.frag.inc"
};
struct PrefilterFrag {
#include "_ref/gen/cubemap_shader_prefilter.frag.inc"
};
struct BrdfFrag {
#include "_ref/gen/brdf_shader.frag.inc"
};

static mat4 to_mat4(const float* m) { mat4 r; for (int c = 0; c < 4; ++c) r.c[c] = vec4(m[c * 4], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]); return r; }
static mat3 to_mat3(const float* m) { mat3 r; for (int c = 0; c < 3; ++c) r.c[c] = vec3(m[c * 3], m[c * 3 + 1], m[c * 3 + 2]); return r; }
static vec4 from(const orc::V4& v) { return vec4(v.x, v.y, v.z, v.w); }

// a float RGBA image bound as a rectangle texture, linear filter, clamp to edge (the render targets: render_pass.cpp:347-365)
static sampler2DRect rect_image(const float* img, int W, int H) {
    sampler2DRect s;
    s.size = ivec2(W, H);
    s.fetch = [=](ivec2 p) { if (p.x < 0 || p.x >= W || p.y < 0 || p.y >= H) return vec4(0.0f); const float* q = img + ((size_t)p.y * W + p.x) * 4; return vec4(q[0], q[1], q[2], q[3]); };
    s.sample = [=](vec2 c) {
        float x = c.x - 0.5f, y = c.y - 0.5f;
        float fx = std::floor(x), fy = std::floor(y), a = x - fx, b = y - fy;
        int i0 = (int)fx, j0 = (int)fy;
        auto at = [&](int i, int j) { i = i < 0 ? 0 : (i > W - 1 ? W - 1 : i); j = j < 0 ? 0 : (j > H - 1 ? H - 1 : j); const float* q = img + ((size_t)j * W + i) * 4; return vec4(q[0], q[1], q[2], q[3]); };
        return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) + at(i0 + 1, j0 + 1) * (a * b);
    };
    return s;
}

}  // namespace glsl

using namespace glsl;

extern "C" {

// ---- render_shader.vert ----
void glslref_vertex(const orc_vert_uniforms* u, const void* verts68, int n, orc_vert_out* out) {
    const orc::Vertex68* v = (const orc::Vertex68*)verts68;
    for (int i = 0; i < n; ++i) {
        RenderVert s;
        s.meshToObject = to_mat4(u->mesh_to_object); s.objectToWorld = to_mat4(u->object_to_world);
        s.projectionMatrix = to_mat4(u->projection); s.worldToCam = to_mat4(u->world_to_cam);
        s.normalToWorld = to_mat3(u->normal_to_world); s.normalToCam = to_mat3(u->normal_to_cam);
        s.stickerProjection = to_mat4(u->sticker_projection);
        s.stickerRange = vec4(u->sticker_range[0], u->sticker_range[1], u->sticker_range[2], u->sticker_range[3]);
        s.position = vec4(v[i].pos[0], v[i].pos[1], v[i].pos[2], 1.0f);   // a vec3 attribute read as vec4: w = 1
        s.textureCoords = vec2(v[i].uv[0], v[i].uv[1]);
        s.vertexColors = vec4(v[i].color[0], v[i].color[1], v[i].color[2], v[i].color[3]);
        s.normal = vec3(v[i].normal[0], v[i].normal[1], v[i].normal[2]);
        s.tangent = vec4(v[i].tangent[0], v[i].tangent[1], v[i].tangent[2], v[i].tangent[3]);
        s.vertexIndex = v[i].vertex_index;
        s.main();
        orc_vert_out& r = out[i];
        std::memset(&r, 0, sizeof r);
        const auto& d = s.primitiveData;
        r.uv[0] = d.interpolatedTextureCoords.x; r.uv[1] = d.interpolatedTextureCoords.y;
        for (int k = 0; k < 3; ++k) {
            r.normal_cam[k] = d.normalInCam[k]; r.normal_w[k] = d.normalInWorld[k]; r.tangent_w[k] = d.tangentInWorld[k];
            r.bitangent_w[k] = d.bitangentInWorld[k]; r.world[k] = d.worldCoordinates[k]; r.cam[k] = d.camCoordinates[k];
        }
        for (int k = 0; k < 4; ++k) { r.objc[k] = d.objectCoordinates[k]; r.position[k] = s.vsPosition[k]; }
        r.sticker[0] = d.stickerCoordinates.x; r.sticker[1] = d.stickerCoordinates.y;
        r.vertex_id = s.gsVertexIndex;
    }
}

// ---- shadow_shader.vert ----
void glslref_shadow_vertex(const float* transformation, const float* pos3, int n, float* clip4) {
    for (int i = 0; i < n; ++i) {
        ShadowVert s; s.transformation = to_mat4(transformation);
        s.position = vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        s.main();
        for (int k = 0; k < 4; ++k) clip4[4 * i + k] = s.gl_Position[k];
    }
}

// ---- render_shader.frag ----
void glslref_fragment(const orc_frag_uniforms* u, const orc_frag_in* in, int n, orc_frag_out* out) {
    const int N = SLB_SHADOW_RES;
    std::vector<float> zero_peel;
    for (int i = 0; i < n; ++i) {
        const orc_frag_in& fi = in[i];
        RenderFrag s;
        for (int k = 0; k < 3; ++k) s.materialParameters[k] = vec4(u->material[4 * k], u->material[4 * k + 1], u->material[4 * k + 2], u->material[4 * k + 3]);
        s.availableTextures = u->available_textures;
        s.lightMapAvailable = u->light_map_available;
        for (int k = 0; k < 3; ++k) {
            s.lightDirections[k] = vec3(u->light_directions[3 * k], u->light_directions[3 * k + 1], u->light_directions[3 * k + 2]);
            s.lightColors[k] = vec3(u->light_colors[3 * k], u->light_colors[3 * k + 1], u->light_colors[3 * k + 2]);
            s.shadowMatrices[k] = to_mat4(u->shadow_matrices + 16 * k);
        }
        s.ambientLight = vec3(u->ambient[0], u->ambient[1], u->ambient[2]);
        s.classIndex = u->class_index; s.instanceIndex = u->instance_index;
        s.camPosition = vec3(u->cam_position[0], u->cam_position[1], u->cam_position[2]);
        s.worldToCam = to_mat4(u->world_to_cam);
        // samplers: material textures with the implicit derivatives of THIS fragment's quad
        const float dudx = fi.uv_dx[0] - fi.uv[0], dvdx = fi.uv_dx[1] - fi.uv[1], dudy = fi.uv_dy[0] - fi.uv[0], dvdy = fi.uv_dy[1] - fi.uv[1];
        sampler2D* mats[5] = {&s.baseColorTexture, &s.normalTexture, &s.metallicRoughnessTexture, &s.emissiveTexture, &s.occlusionTexture};
        for (int k = 0; k < 5; ++k) {
            const orc::Texture* t = (const orc::Texture*)u->tex[k];
            mats[k]->sample = [=](vec2 uv) { return from(orc::sample_texture_2d(*t, uv.x, uv.y, dudx, dvdx, dudy, dvdy)); };
        }
        const orc::LightMap* lm = (const orc::LightMap*)u->light_map;
        s.lightMapBRDFLUT.sample = [=](vec2 uv) { return from(orc::sample_lut(*lm, uv.x, uv.y)); };
        s.lightMapIrradiance.sample_lod = [=](vec3 d, float lod) { return from(orc::sample_cube_lod(lm->irradiance, orc::V3(d.x, d.y, d.z), lod)); };
        s.lightMapPrefilter.sample_lod = [=](vec3 d, float lod) { return from(orc::sample_cube_lod(lm->prefilter, orc::V3(d.x, d.y, d.z), lod)); };
        const orc::Texture* st = (const orc::Texture*)u->sticker;
        if (st) {
            s.stickerTexture.size = ivec2(st->w, st->h);
            s.stickerTexture.sample = [=](vec2 c) { return from(orc::sample_texture_rect_linear(*st, c.x, c.y)); };
        } else {
            // no sticker bound: whatever texture unit 8 holds is sampled in the reference; the restatement skips the block
            // (SURVEY A.9). A transparent texture makes mix(baseColor, sticker, 0) the identity, i.e. the same thing.
            s.stickerTexture.size = ivec2(1, 1);
            s.stickerTexture.sample = [](vec2) { return vec4(0.0f); };
        }
        const float* peel = u->peel; const int W = u->width;
        s.depthTexture.fetch = [=](ivec2 p) { if (!peel) return vec4(0.0f); const float* q = peel + ((size_t)p.y * W + p.x) * 4; return vec4(q[0], q[1], q[2], q[3]); };
        s.shadowMap.size = ivec3(N, N, 3);
        const uint32_t* const* maps = u->shadow_map;
        // sampler2DArrayShadow, linear filter, compare LEQUAL, clamp to edge (render_pass.cpp:271-292): the reference value is
        // clamped to [0,1], each of the 2x2 texels compares, the results are filtered (GL 4.5 §8.23)
        s.shadowMap.compare = [=](vec4 p) {
            const uint32_t* map = maps[(int)p.z];
            float ref = p.w < 0.0f ? 0.0f : (p.w > 1.0f ? 1.0f : p.w);
            float x = p.x * N - 0.5f, y = p.y * N - 0.5f;
            float fx = std::floor(x), fy = std::floor(y), a = x - fx, b = y - fy;
            int i0 = (int)fx, j0 = (int)fy;
            auto cmp = [&](int ii, int jj) {
                ii = ii < 0 ? 0 : (ii > N - 1 ? N - 1 : ii); jj = jj < 0 ? 0 : (jj > N - 1 ? N - 1 : jj);
                float stored = map ? (float)map[(size_t)jj * N + ii] / 16777215.0f : 1.0f;
                return ref <= stored ? 1.0f : 0.0f;
            };
            return cmp(i0, j0) * ((1 - a) * (1 - b)) + cmp(i0 + 1, j0) * (a * (1 - b)) + cmp(i0, j0 + 1) * ((1 - a) * b) + cmp(i0 + 1, j0 + 1) * (a * b);
        };
        // inputs
        auto& d = s.fragmentData;
        d.interpolatedTextureCoords = vec2(fi.uv[0], fi.uv[1]);
        d.normalInWorld = vec3(fi.normal_w[0], fi.normal_w[1], fi.normal_w[2]);
        d.tangentInWorld = vec3(fi.tangent_w[0], fi.tangent_w[1], fi.tangent_w[2]);
        d.bitangentInWorld = vec3(fi.bitangent_w[0], fi.bitangent_w[1], fi.bitangent_w[2]);
        d.objectCoordinates = vec4(fi.objc[0], fi.objc[1], fi.objc[2], fi.objc[3]);
        d.worldCoordinates = vec3(fi.world[0], fi.world[1], fi.world[2]);
        d.camCoordinates = vec3(fi.cam[0], fi.cam[1], fi.cam[2]);
        d.stickerCoordinates = vec2(fi.sticker[0], fi.sticker[1]);
        s.gl_FragCoord = vec4(fi.frag_x, fi.frag_y, 0.5f, 1.0f);
        s.gl_FrontFacing = fi.front_facing != 0;
        s.g_vertexIndices = uvec3(fi.vertex_ids[0], fi.vertex_ids[1], fi.vertex_ids[2]);
        s.g_barycentricCoeffs = vec3(fi.bary[0], fi.bary[1], fi.bary[2]);
        s.main();
        orc_frag_out& o = out[i];
        std::memset(&o, 0, sizeof o);
        if (s._discarded) { o.discarded = 1; continue; }
        for (int k = 0; k < 4; ++k) { o.color[k] = s.color[k]; o.objc[k] = s.objectCoordinatesOut[k]; o.camc[k] = s.camCoordinatesOut[k]; o.normal[k] = s.normalOut[k]; }
        o.class_index = s.classIndexOut; o.instance_index = s.instanceIndexOut;
        o.vertex_ids[0] = s.vertexIndices.x; o.vertex_ids[1] = s.vertexIndices.y; o.vertex_ids[2] = s.vertexIndices.z;
        for (int k = 0; k < 3; ++k) o.bary[k] = s.barycentricCoeffs[k];
    }
}

// ---- tone_map_shader.frag: RGBA32F colour -> RGBA8 unorm (GL: round to nearest; NaN converts to 0) ----
void glslref_tonemap(const float* hdr, int n, float manual_exposure, const float* avg, uint8_t* rgba8) {
    for (int i = 0; i < n; ++i) {
        ToneMapFrag s;
        s.manualExposure = manual_exposure;
        const float* px = hdr + 4 * (size_t)i;
        s.rgbSampler.fetch = [=](ivec2, int) { return vec4(px[0], px[1], px[2], px[3]); };
        s.luminanceSampler.levels = 1;
        s.luminanceSampler.fetch = [=](ivec2, int) { return vec4(avg[0], avg[1], avg[2], avg[3]); };
        s.gl_FragCoord = vec4(0.5f, 0.5f, 0.5f, 1.0f);
        s.main();
        for (int k = 0; k < 4; ++k) {
            float v = s.outputColor[k];
            if (!(v == v)) v = 0.0f;
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            rgba8[4 * (size_t)i + k] = (uint8_t)std::lrintf(v * 255.0f);
        }
    }
}

// ---- ssao_shader.frag / ssao_apply_shader.frag over a whole frame ----
// kernel + noise: the constructor code of src/shaders/ssao_shader.cpp:72-112 restated in orc_render.cpp (ssao_tables) is
// passed in by the test; here only the shader text runs.
void glslref_ssao(const float* camc, const float* normals, int W, int H, const float* projection, const float* noise16x3, const float* kernel64x3, float* ao) {
    #pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            SsaoFrag s;
            s.positions = rect_image(camc, W, H);
            s.normals = rect_image(normals, W, H);
            // 4x4 RGB32F noise texture, nearest, repeat (ssao_shader.cpp:88-93)
            s.noiseSampler.sample = [=](vec2 uv) {
                int i = (int)std::floor(uv.x * 4.0f), j = (int)std::floor(uv.y * 4.0f);
                i = ((i % 4) + 4) % 4; j = ((j % 4) + 4) % 4;
                const float* q = noise16x3 + 3 * (j * 4 + i); return vec4(q[0], q[1], q[2], 1.0f);
            };
            for (int k = 0; k < 64; ++k) s.samples[k] = vec3(kernel64x3[3 * k], kernel64x3[3 * k + 1], kernel64x3[3 * k + 2]);
            s.projection = to_mat4(projection);
            s.gl_FragCoord = vec4(px + 0.5f, py + 0.5f, 0.5f, 1.0f);
            s.main();
            ao[(size_t)py * W + px] = s.ao;
        }
}
void glslref_ssao_apply(const float* hdr, const float* ao, const float* camc, int W, int H, float* out) {
    #pragma omp parallel for
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            SsaoApplyFrag s;
            s.rgbSampler.fetch = [=](ivec2 p, int) { const float* q = hdr + ((size_t)p.y * W + p.x) * 4; return vec4(q[0], q[1], q[2], q[3]); };
            s.aoSampler.fetch = [=](ivec2 p, int) { if (p.x < 0 || p.x >= W || p.y < 0 || p.y >= H) return vec4(0.0f); return vec4(ao[(size_t)p.y * W + p.x], 0.0f, 0.0f, 1.0f); };
            s.coordinateSampler = rect_image(camc, W, H);
            s.gl_FragCoord = vec4(px + 0.5f, py + 0.5f, 0.5f, 1.0f);
            s.main();
            for (int k = 0; k < 4; ++k) out[((size_t)py * W + px) * 4 + k] = s.outputColor[k];
        }
}

// ---- background_shader.{vert,frag}: the vertex stage runs on the full-screen triangle's corners, the varying is
// interpolated linearly to the pixel centre (no perspective: w = 1) ----
void glslref_background_image(const void* tex, int W, int H, float* out) {
    const orc::Texture* t = (const orc::Texture*)tex;
    // corners of NDC space through the vertex shader: textureCoords is affine in position, so two corners define it
    BackgroundVert v00, v11;
    v00.position = vec2(-1.0f, -1.0f); v00.main();
    v11.position = vec2(1.0f, 1.0f); v11.main();
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const float fx = (px + 0.5f) / W, fy = (py + 0.5f) / H;   // window y grows with NDC y (memory row = GL row)
            BackgroundFrag s;
            s.textureCoords = vec2(v00.textureCoords.x * (1.0f - fx) + v11.textureCoords.x * fx, v00.textureCoords.y * (1.0f - fy) + v11.textureCoords.y * fy);
            s.rgb.size = ivec2(t->w, t->h);
            s.rgb.sample = [=](vec2 c) { return from(orc::sample_texture_rect_linear(*t, c.x, c.y)); };
            s.main();
            for (int k = 0; k < 4; ++k) out[((size_t)py * W + px) * 4 + k] = s.fragmentColor[k];
        }
}
// ---- background_cube_shader.vert: NDC position of cube-surface points (the fragment stage then samples the cube map with
// the interpolated position itself) ----
void glslref_skybox_project(const float* projection, const float* view, const float* pos3, int n, float* ndc_xy_w) {
    for (int i = 0; i < n; ++i) {
        BackgroundCubeVert s; s.projection = to_mat4(projection); s.view = to_mat4(view);
        s.position = vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        s.main();
        ndc_xy_w[3 * i] = s.gl_Position.x / s.gl_Position.w; ndc_xy_w[3 * i + 1] = s.gl_Position.y / s.gl_Position.w; ndc_xy_w[3 * i + 2] = s.gl_Position.w;
    }
}

// ---- light-map precompute programs: one texel per call site (WorldPos given) ----
void glslref_lightmap_texels(int which, const slb_lightmap_desc* d, const void* lm_h, const float* world_pos, int n, float roughness,
                             float irradiance_lod, float* rgb_out) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
        vec3 wp(world_pos[3 * i], world_pos[3 * i + 1], world_pos[3 * i + 2]);
        vec4 c;
        auto env = [=](vec3 dir, float lod) { float q[4]; float dd[3] = {dir.x, dir.y, dir.z}; orc_test_cube_sample(lm_h, 0, dd, lod, q); return vec4(q[0], q[1], q[2], q[3]); };
        if (which == 0) {
            EquirectFrag s; s.WorldPos = wp;
            const float* img = d->equirect_rgb; const int W = d->width, H = d->height;
            // level-0 bilinear, clamp to edge (documented deviation: the reference samples with implicit LOD + anisotropy)
            s.equirectangularMap.sample = [=](vec2 uv) {
                float x = uv.x * W - 0.5f, y = uv.y * H - 0.5f;
                float fx = std::floor(x), fy = std::floor(y), a = x - fx, b = y - fy;
                int i0 = (int)fx, j0 = (int)fy;
                auto at = [&](int ii, int jj) { ii = ii < 0 ? 0 : (ii > W - 1 ? W - 1 : ii); jj = jj < 0 ? 0 : (jj > H - 1 ? H - 1 : jj); const float* p = img + ((size_t)jj * W + ii) * 3; return vec4(p[0], p[1], p[2], 1.0f); };
                return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) + at(i0 + 1, j0 + 1) * (a * b);
            };
            s.main(); c = s.FragColor;
        } else if (which == 1) {
            IrradianceFrag s; s.WorldPos = wp;
            s.environmentMap.sample_lod = env; s.environmentMap.implicit_lod = irradiance_lod;
            s.main(); c = s.FragColor;
        } else {
            PrefilterFrag s; s.WorldPos = wp; s.roughness = roughness;
            s.environmentMap.sample_lod = env;
            s.main(); c = s.FragColor;
        }
        rgb_out[3 * i] = c.x; rgb_out[3 * i + 1] = c.y; rgb_out[3 * i + 2] = c.z;
    }
}
void glslref_brdf_lut(const float* ndotv_roughness, int n, float* ab_out) {
    for (int i = 0; i < n; ++i) {
        BrdfFrag s; s.TexCoords = vec2(ndotv_roughness[2 * i], ndotv_roughness[2 * i + 1]);
        s.main();
        ab_out[2 * i] = s.FragColor.x; ab_out[2 * i + 1] = s.FragColor.y;
    }
}

}  // extern "C"
