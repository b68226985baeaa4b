// orc_test_hooks.h — plain-C interface of the per-stage test hooks of the oracle and of the compiled reference shaders
// (oracle/_ref/libglslref.so). TEST INFRASTRUCTURE ONLY. Both sides take the SAME structs, so tests/test_glsl_ref.py can
// feed identical random inputs to `orc_test_*` (the oracle's restatement) and `glslref_*` (the reference's GLSL text
// compiled as C++) and compare the outputs.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_frag_uniforms {
    float material[12];            // materialParameters[3]: baseColor, emissive, (alphaCutoff, metallic, roughness, 0)
    uint32_t available_textures;   // bit i = TextureInput i bound (render_shader.cpp:34-47: base 0, normal 1, mr 2, emissive 3, occlusion 4)
    uint32_t light_map_available;
    float light_directions[9], light_colors[9];
    float shadow_matrices[48];     // 3 column-major mat4
    float ambient[3];
    uint32_t class_index, instance_index;
    float cam_position[3];
    float world_to_cam[16];
    const void* tex[5];            // oracle Texture handles (orc_texture_create, SLB_TEXTURE_2D) or NULL
    const void* light_map;         // oracle LightMap handle or NULL
    const void* sticker;           // oracle Texture handle (SLB_TEXTURE_RECT) or NULL
    const uint32_t* shadow_map[3]; // d24 planes, SLB_SHADOW_RES^2, or NULL
    const float* peel;             // previous layer's coordinate buffer HxWx4 (w = depth) or NULL (== zeros)
    int32_t width, height;
} orc_frag_uniforms;

typedef struct orc_frag_in {       // the interpolated DataBridge (render_shader.glsl) + built-ins of one fragment
    float uv[2], uv_dx[2], uv_dy[2];   // uv here and at the quad neighbours in x / y (implicit derivatives)
    float normal_w[3], tangent_w[3], bitangent_w[3];
    float objc[4];
    float world[3], cam[3];
    float sticker[2];
    int32_t front_facing;
    float frag_x, frag_y;          // gl_FragCoord.xy
    uint32_t vertex_ids[3];
    float bary[3];
} orc_frag_in;

typedef struct orc_frag_out {
    float color[4], objc[4], camc[4], normal[4];
    uint32_t class_index, instance_index, vertex_ids[3];
    float bary[3];
    int32_t discarded;
} orc_frag_out;

typedef struct orc_vert_uniforms {
    float mesh_to_object[16], object_to_world[16], world_to_cam[16], projection[16];
    float normal_to_world[9], normal_to_cam[9];
    float sticker_projection[16], sticker_range[4];
} orc_vert_uniforms;

typedef struct orc_vert_out {
    float uv[2], normal_cam[3], normal_w[3], tangent_w[3], bitangent_w[3], objc[4], world[3], cam[3], sticker[2];
    float position[4];
    uint32_t vertex_id;
} orc_vert_out;

#ifdef __cplusplus
}
#endif
