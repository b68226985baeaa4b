/* slb.h — C ABI of the B200-native multi-target rasteriser ("slb" = stillleben-b200).
 *
 * This is the drop-in boundary for ONE path of AIS-Bonn/stillleben:
 *     sl::RenderPass::render(Scene&)          (reference: src/render_pass.cpp:303-796)
 * plus the result hand-off that follows it  (reference: src/cuda_interop.cpp:83-208,
 *                                            python/src/py_magnum.cpp:17-46).
 *
 * The reference has no FFI for this path (it is a C++ class surface + pybind11),
 * so every entry point below cites the reference C++ interface it replaces.  All functions
 * are extern "C", take plain pointers and sizes, return an int status (SLB_OK == 0) and never
 * throw; slb_last_error() gives the message.  No torch / Magnum / CUDA-runtime types appear in
 * a signature: a CUDA stream is passed as void* (cudaStream_t), device buffers as void*.
 *
 * Streams: every entry point with a `stream` argument queues its work there (NULL = the context's own stream). The
 * per-context scratch is shared between calls, so the library ORDERS calls itself: the stream of a call first waits for
 * whatever the previous call on this context queued, whatever stream that used. Calls that read back, update or destroy
 * (slb_result_read*, slb_mesh_*, slb_*_destroy, slb_ctx_synchronize) wait for the context's own stream and for the stream
 * of the latest call. A context must not be used from two host threads at once.
 *
 * Conventions (identical to the reference):
 *   - all matrices are 4x4 float32 COLUMN-MAJOR (Magnum::Matrix4 memory order)
 *   - camera frame: +x right, +y down, +z forward; memory row r == GL window y == r
 *   - vertex stream: 68-byte interleaved records produced by the reference's
 *     consolidateMesh()  (reference: src/mesh_tools/consolidate.cpp:53-61)
 */
#ifndef SLB_H
#define SLB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLB_ABI_VERSION 1
#define SLB_NUM_LIGHTS 3          /* reference: include/stillleben/common.h:17 (NumLights)      */
#define SLB_VERTEX_STRIDE 68      /* reference: src/mesh_tools/consolidate.cpp:53-61             */
#define SLB_SHADOW_RES 2048       /* reference: src/render_pass.cpp:271                          */
#define SLB_INVALID_COORD 3000.0f /* reference: src/render_pass.cpp:316 (clear value "invalid")  */

/* status codes */
enum {
    SLB_OK = 0,
    SLB_ERR_INVALID_ARGUMENT = 1, /* reference: std::invalid_argument -> Python ValueError        */
    SLB_ERR_RUNTIME = 2,          /* reference: std::runtime_error / logic_error -> RuntimeError  */
    SLB_ERR_CUDA = 3,             /* reference: std::abort() on CUDA/GL failure (cuda_interop.cpp:99) */
    SLB_ERR_OUT_OF_MEMORY = 4
};

/* The eight colour attachments of the reference frame buffer, in attachment order
 * (reference: src/shaders/common.glsl:14-21, src/render_pass.cpp:347-365). */
enum {
    SLB_TARGET_RGB = 0,          /* RGBA8   4 B/px  tone-mapped colour                              */
    SLB_TARGET_COORD = 1,        /* RGBA32F 16 B/px xyz = object-frame point, w = camera z (depth)  */
    SLB_TARGET_CLASS = 2,        /* R16UI   2 B/px                                                  */
    SLB_TARGET_INSTANCE = 3,     /* R16UI   2 B/px                                                  */
    SLB_TARGET_NORMAL = 4,       /* RGBA32F 16 B/px xyz = camera-space unit normal, w = N.V         */
    SLB_TARGET_VERTEX_INDEX = 5, /* RGBA32UI 16 B/px xyz = three one-based vertex ids               */
    SLB_TARGET_BARY = 6,         /* RGBA32F 16 B/px xyz = perspective-correct barycentrics          */
    SLB_TARGET_CAM_COORD = 7,    /* RGBA32F 16 B/px (x,y,z,1) camera-frame point                    */
    SLB_NUM_TARGETS = 8
};
#define SLB_TARGETS_SIX 0x1Fu /* rgb + coord/depth + class + instance + normals  = 40 B/px */
#define SLB_TARGETS_ALL 0xFFu /* all eight attachments                           = 88 B/px */

/* texture sampler enums (values follow glTF / GL numeric order loosely; own namespace) */
enum {
    SLB_WRAP_REPEAT = 0,
    SLB_WRAP_CLAMP_TO_EDGE = 1,
    SLB_WRAP_MIRRORED_REPEAT = 2,
    SLB_WRAP_CLAMP_TO_BORDER = 3 /* border colour transparent black (context.cpp:597-599) */
};
enum {
    SLB_FILTER_NEAREST = 0,
    SLB_FILTER_LINEAR = 1,
    SLB_FILTER_NEAREST_MIPMAP_NEAREST = 2,
    SLB_FILTER_LINEAR_MIPMAP_NEAREST = 3,
    SLB_FILTER_NEAREST_MIPMAP_LINEAR = 4,
    SLB_FILTER_LINEAR_MIPMAP_LINEAR = 5
};
enum {
    SLB_TEXTURE_2D = 0,  /* normalised coords + full mip chain  (reference: GL::Texture2D, mesh.cpp:656-663) */
    SLB_TEXTURE_RECT = 1 /* pixel coords, clamp-to-border transparent (reference: GL::RectangleTexture, context.cpp:597-599) */
};

typedef struct slb_ctx slb_ctx;
typedef struct slb_mesh slb_mesh;
typedef struct slb_texture slb_texture;
typedef struct slb_lightmap slb_lightmap;
typedef struct slb_result slb_result;

/* ---- context -------------------------------------------------------------------------- */

/* Replaces sl::Context::CreateCUDA(device) (reference: src/context.cpp:411, python/src/py_context.cpp:34).
 * EGL/GL bring-up disappears; this selects the CUDA device and creates the work stream. */
int slb_ctx_create(int device, slb_ctx** out);
void slb_ctx_destroy(slb_ctx* ctx);
/* Last error message for ctx (or for the failed slb_ctx_create when ctx == NULL). */
const char* slb_last_error(const slb_ctx* ctx);
int slb_abi_version(void);
int slb_ctx_device(const slb_ctx* ctx);
/* Block until all work queued on the context's streams is done. */
int slb_ctx_synchronize(slb_ctx* ctx);

/* Page-locked host memory for slb_render_batch_host destinations / upload sources (cudaHostAlloc): D2H copies
 * into pageable memory cannot overlap with rendering. (New: the reference reads results with glGetTexImage into
 * pageable Magnum::Image2D storage, python/src/py_magnum.cpp:33-45.) */
int slb_host_alloc(slb_ctx* ctx, size_t bytes, void** out);
void slb_host_free(slb_ctx* ctx, void* ptr);

/* ---- assets --------------------------------------------------------------------------- */

typedef struct slb_image {
    const void* pixels; /* host OR device pointer (UVA); rows bottom-up exactly as the importers
                           deliver them: texel (i,j) at pixels[(j*width+i)*channels] */
    int32_t width, height;
    int32_t channels; /* 3 or 4, uint8 (reference accepts RGB8/RGBA8 only: mesh.cpp:644-653) */
    int32_t wrap_s, wrap_t;
    int32_t min_filter, mag_filter;
} slb_image;

/* One draw of the reference's object->draw() loop (reference: src/object.cpp:101-140, Drawable). */
typedef struct slb_submesh {
    uint32_t index_offset; /* first index into the consolidated index buffer */
    uint32_t index_count;  /* multiple of 3 */
    int32_t material;      /* index into materials[], -1 = context default material (context.cpp:382-384) */
    uint32_t reserved;
} slb_submesh;

/* What RenderShader::setMaterial() derives from Trade::MaterialData BEFORE the per-object
 * override (reference: src/shaders/render_shader.cpp:332-418).  The importer-dependent
 * defaulting (metallic 0.04 / roughness 0.5 / 1.0 with texture / glTF factor) is applied by
 * the caller that parsed the file; the per-object override is applied inside the library. */
typedef struct slb_material {
    float base_color[4];
    float emissive[4];
    float metallic;
    float roughness;
    int32_t tex_base_color; /* index into images[] or -1 */
    int32_t tex_normal;
    int32_t tex_metallic_roughness;
    int32_t tex_emissive;
    int32_t tex_occlusion;
    int32_t reserved;
} slb_material;

/* Replaces Mesh::loadVisual(): VBO/IBO/texture upload (reference: src/mesh.cpp:624-745).
 * vertices: n_vertices * 68 bytes; indices: n_indices u32.  bbox_min/max = Mesh::m_bbox (mesh
 * frame, before pretransform). Host or device source pointers are both accepted (UVA). */
int slb_mesh_upload(slb_ctx* ctx, const void* vertices, uint32_t n_vertices, const uint32_t* indices,
                    uint32_t n_indices, const slb_submesh* submeshes, uint32_t n_submeshes,
                    const slb_material* materials, uint32_t n_materials, const slb_image* images,
                    uint32_t n_images, const float bbox_min[3], const float bbox_max[3], slb_mesh** out);
/* Replaces Mesh::recompileMesh() (reference: src/mesh.cpp:818-821): re-upload of the whole 68-byte stream, nothing
 * recomputed. */
int slb_mesh_update_vertices(slb_ctx* ctx, slb_mesh* mesh, const void* vertices, uint32_t n_vertices);
/* Replaces Mesh::updateVertexPositionsAndColors / updateVertexPositions / updateVertexColors (reference:
 * src/mesh.cpp:747-761,823-855; python/src/py_mesh.cpp:100-212): point[id-1] += position_update[i] (n x 3 floats, may be
 * NULL), colour[id-1] += color_update[i] (n x 4 floats, may be NULL); vertex_ids are the ONE-BASED vertex ids the
 * vertex-index target reports. After a position update the vertex normals are recomputed like Mesh::recomputeNormals
 * (area-weighted face normals, src/mesh.cpp:763-816) — on the device, the vertex stream never leaves HBM. The bounding
 * box is NOT updated (the reference keeps the load-time box). The arrays may be host or device memory. */
int slb_mesh_update_positions_and_colors(slb_ctx* ctx, slb_mesh* mesh, const int32_t* vertex_ids, uint32_t n,
                                         const float* position_update, const float* color_update, void* stream);
/* Replaces Mesh::setVertexPositions (reference: src/mesh.cpp:857-870): all positions replaced, normals recomputed.
 * SLB_ERR_INVALID_ARGUMENT if n_vertices differs ("Number of new vertices should match the existing mesh vertices"). */
int slb_mesh_set_positions(slb_ctx* ctx, slb_mesh* mesh, const float* positions, uint32_t n_vertices, void* stream);
/* Replaces Mesh::setVertexColors (reference: src/mesh.cpp:872-885). */
int slb_mesh_set_colors(slb_ctx* ctx, slb_mesh* mesh, const float* colors, uint32_t n_vertices, void* stream);
/* Replaces Mesh::recomputeNormals (reference: src/mesh.cpp:763-816). */
int slb_mesh_recompute_normals(slb_ctx* ctx, slb_mesh* mesh, void* stream);
/* The current 68-byte vertex stream (Mesh::meshPoints / meshNormals / meshColors views, reference: src/mesh.cpp:930-998);
 * vertices_out: n_vertices * 68 bytes of host or device memory. Synchronous. */
int slb_mesh_read_vertices(slb_ctx* ctx, const slb_mesh* mesh, void* vertices_out, uint32_t n_vertices);
void slb_mesh_destroy(slb_ctx* ctx, slb_mesh* mesh);

/* Replaces Context::loadTexture / loadTexture2D and the sl.Texture / sl.Texture2D constructors
 * (reference: src/context.cpp:560-640, python/src/py_magnum.cpp:115-198). */
int slb_texture_create(slb_ctx* ctx, const slb_image* image, int kind, slb_texture** out);
void slb_texture_destroy(slb_ctx* ctx, slb_texture* tex);
/* Read back mip level `level` as RGBA8 (host_out may be NULL to query the size). Returns the number of
 * levels (> 0) or a negative/zero status on error. The mip chain is what glGenerateMipmap produces in
 * the reference (src/mesh.cpp:661-663); tests pin it bit-exactly against the oracle. */
int slb_texture_read_level(slb_ctx* ctx, const slb_texture* tex, int level, int32_t* width, int32_t* height, void* host_out);

/* Light map = the four GPU objects LightMap::load() leaves behind (reference:
 * src/light_map.cpp:266-611, include/stillleben/light_map.h:33-56):
 *   env cube 512^2 RGBA32F (mip chain), irradiance cube 32^2, prefilter cube 128^2 x 5 mips,
 *   BRDF LUT 512^2 RG (stored RGBA32F) + up to 3 directional lights parsed from the .ibl file.
 * slb_lightmap_create runs the precompute on the device from a lat-long float RGB image
 * (rows bottom-up, GL addressing), i.e. it replaces the cubemap_shader_* / brdf_shader passes. */
typedef struct slb_lightmap_desc {
    const float* equirect_rgb; /* host, width*height*3 float32 */
    int32_t width, height;
    int32_t n_lights;                        /* <= SLB_NUM_LIGHTS (light_map.cpp:328-346) */
    float light_directions[SLB_NUM_LIGHTS][3];
    float light_colors[SLB_NUM_LIGHTS][3];
} slb_lightmap_desc;
int slb_lightmap_create(slb_ctx* ctx, const slb_lightmap_desc* desc, slb_lightmap** out);
/* Same with explicit map sizes / sample count (values <= 0 select the reference's 512 / 32 / 128 / 512 and
 * 1024 samples, src/light_map.cpp:381,451,510,580); reduced sizes keep parity tests fast. */
int slb_lightmap_create_ex(slb_ctx* ctx, const slb_lightmap_desc* desc, int env_size, int irradiance_size, int prefilter_size,
                           int lut_size, int n_samples, slb_lightmap** out);
/* The same light map from maps precomputed elsewhere (rank 0's slb_lightmap_create read back with slb_lightmap_read and
 * shipped in the one load-time broadcast of the asset arena): env0 [6][e][e][4], irradiance [6][i][i][4],
 * prefilter = five levels packed [6][p>>l][p>>l][4], lut [l][l][4] floats, host or device memory. desc supplies the
 * lights only (equirect_rgb is ignored). */
int slb_lightmap_create_from_maps(slb_ctx* ctx, const slb_lightmap_desc* desc, const float* env0, int env_size,
                                  const float* irradiance, int irradiance_size, const float* prefilter, int prefilter_size,
                                  const float* lut, int lut_size, slb_lightmap** out);
/* sizes[4] = env, irradiance, prefilter (level 0), LUT edge lengths */
int slb_lightmap_sizes(const slb_lightmap* lm, int32_t sizes[4]);
/* Read back the precomputed maps (tests / oracle cross-checks). which: 0 env cube level 0
 * (6*512*512*4), 1 irradiance (6*32*32*4), 2 prefilter all mips packed, 3 BRDF LUT (512*512*4). */
int slb_lightmap_read(slb_ctx* ctx, const slb_lightmap* lm, int which, float* host_out, size_t n_floats);
void slb_lightmap_destroy(slb_ctx* ctx, slb_lightmap* lm);

/* ---- scene descriptor (everything RenderPass::render reads; SURVEY §8b "inputs") ------- */

typedef struct slb_object_desc {
    const slb_mesh* mesh;
    float pose[16];         /* objectToWorld  = Object::pose()                 (render_pass.cpp:590) */
    float pretransform[16]; /* meshToObject   = Mesh::pretransform()           (object.cpp:83)       */
    uint32_t class_index;   /* Mesh::classIndex()    <= 65535                   (mesh.cpp:1083-1089)  */
    uint32_t instance_index;/* Object::instanceIndex() <= 65535                 (object.cpp:376-382)  */
    float metallic;         /* per-object override, < 0 = none                  (object.h:277-278)    */
    float roughness;
    int32_t casts_shadows;  /* Object::castsShadows()                           (render_pass.cpp:447) */
    int32_t visible;        /* result of the DrawPredicate, evaluated by caller (render_pass.cpp:444,587) */
    const slb_texture* sticker_texture; /* RECT texture or NULL                 (render_pass.cpp:605) */
    float sticker_projection[16];       /* Object::stickerViewProjection()      (object.cpp:494-513)  */
    float sticker_range[4];             /* min.x, min.y, size.x, size.y (raw Range2D)                 */
} slb_object_desc;

typedef struct slb_scene_desc {
    int32_t width, height;  /* Scene::viewport() */
    float projection[16];   /* camera().projectionMatrix()                      (scene.cpp:222-258)   */
    float world_to_cam[16]; /* camera().cameraMatrix()                                                 */
    /* manual lighting (ignored when light_map != NULL: render_shader.cpp:270-315) */
    float light_directions[SLB_NUM_LIGHTS][3];
    float light_colors[SLB_NUM_LIGHTS][3];
    float ambient_light[3];
    const slb_lightmap* light_map;
    /* background plane (render_pass.cpp:545-582); drawn iff dot(size,size) > 0 */
    float background_plane_size[2];
    float background_plane_pose[16];
    const slb_texture* background_plane_texture; /* TEXTURE_2D or NULL */
    const slb_texture* background_image;         /* TEXTURE_RECT or NULL (render_pass.cpp:637-647) */
    float manual_exposure;                       /* < 0 = auto exposure (tone_map_shader.frag:109-123) */
    int32_t ssao_enabled;                        /* RenderPass::ssaoEnabled() (render_pass.h:150) */
    const slb_object_desc* objects;
    int32_t n_objects;
} slb_scene_desc;

/* ---- results --------------------------------------------------------------------------- */

/* Replaces RenderPass::Result + CUDATexture/CUDAMapper (reference: include/stillleben/render_pass.h:48-78,
 * include/stillleben/cuda_interop.h:17-69).  A result holds n_frames frames of the selected
 * targets as dense linear device arrays [n_frames][height][width][channels] — the layout the
 * reference's accessors produce with cudaMemcpy2DFromArray (py_magnum.cpp:17-30) — so map/unmap
 * and the per-accessor copy disappear.
 * external_ptrs: optional array of SLB_NUM_TARGETS device pointers owned by the caller (e.g.
 * torch tensors); entries for targets in target_mask must then be non-NULL. NULL = library
 * allocates. */
int slb_result_create(slb_ctx* ctx, int32_t width, int32_t height, int32_t n_frames, uint32_t target_mask,
                      void* const* external_ptrs, slb_result** out);
/* Device pointer + bytes per pixel of every target (NULL / 0 for targets not in the mask). */
int slb_result_ptrs(const slb_result* res, void* ptrs[SLB_NUM_TARGETS], size_t bytes_per_pixel[SLB_NUM_TARGETS]);
/* Copy one target of frames [first_frame, first_frame + n_frames) to host memory (replaces the CPU
 * branch of extract(), py_magnum.cpp:33-45). */
int slb_result_read(slb_ctx* ctx, const slb_result* res, int target, int32_t first_frame, int32_t n_frames,
                    void* host_out, size_t host_bytes);
/* Pre-tone-map HDR colour (RGBA32F) of a frame, kept only if slb_ctx_set_option(KEEP_HDR) is on. */
int slb_result_read_hdr(slb_ctx* ctx, const slb_result* res, int32_t frame, float* host_out, size_t n_floats);
void slb_result_destroy(slb_ctx* ctx, slb_result* res);

/* ---- the hot path ------------------------------------------------------------------------ */

/* Replaces RenderPass::render(Scene&, result, depthBufferResult, predicate) for a BATCH of
 * independent scenes (reference: src/render_pass.cpp:303; python/src/py_render_pass.cpp:252-258).
 * Scene i is rendered into frame first_frame + i of `result`.  depth_peel (may be NULL) is the
 * reference's depthBufferResult: frame first_frame + i of it supplies the previous layer's
 * coord.w (render_shader.frag:229-233).  All scenes of one call must share width/height with
 * the result.  Work is queued on `stream` (cudaStream_t as void*, NULL = the context's own
 * stream); the call returns without synchronising. */
int slb_render_batch(slb_ctx* ctx, const slb_scene_desc* scenes, int32_t n_scenes, slb_result* result,
                     int32_t first_frame, const slb_result* depth_peel, void* stream);

/* Same path, host buffers end to end: renders n_scenes scenes in internal sub-batches and copies
 * the selected targets into caller-provided HOST arrays host_ptrs[t] of layout
 * [n_scenes][H][W][C] (pinned memory recommended), overlapping D2H copies of sub-batch k with
 * the rendering of sub-batch k+1.  Synchronous: returns when host_ptrs are complete. */
int slb_render_batch_host(slb_ctx* ctx, const slb_scene_desc* scenes, int32_t n_scenes, uint32_t target_mask,
                          void* const host_ptrs[SLB_NUM_TARGETS]);

/* ---- statistics / options ---------------------------------------------------------------- */

typedef struct slb_stats {
    uint64_t kernel_launches;  /* kernels launched by this library since ctx creation */
    uint64_t frames_rendered;
    uint64_t triangles_submitted;
    uint64_t triangles_binned; /* (sub)triangle-tile pairs emitted by the binner, last batch */
    uint64_t bytes_h2d;        /* descriptor uploads */
    uint64_t bytes_d2h;
    float last_kernel_ms[8];   /* when option TIME_KERNELS is on: clears, setup, scan, emit, raster, shade/store, ssao,
                                  post — summed over the render calls since the previous slb_ctx_get_stats */
} slb_stats;
int slb_ctx_get_stats(slb_ctx* ctx, slb_stats* out);

enum {
    SLB_OPT_TIME_KERNELS = 1, /* record per-kernel CUDA-event times (adds syncs; off by default) */
    SLB_OPT_KEEP_HDR = 2,     /* keep the HDR colour buffer readable after render             */
    SLB_OPT_MAX_SUBBATCH = 3, /* frames per internal sub-batch (default 64)                   */
    /* Raster path selection (tuning / testing; results are bit-identical for every setting):
     * triangles whose pixel box holds at most DIRECT_MAX pixels are rasterised by one thread of the
     * setup kernel, up to WARP_MAX pixels by one warp of it; larger ones take the tiled path. 0/0 =
     * everything through the tiled path. Defaults 128 / 4096. */
    SLB_OPT_DIRECT_MAX = 4,
    SLB_OPT_WARP_MAX = 5,
    /* 1 (default): sub-batches that use no normal / metallic-roughness / emissive / occlusion textures, stickers,
     * light maps or projective transformations run the lean instantiation of the shade kernel; 0: always the full one. */
    SLB_OPT_LEAN_SHADE = 6,
    /* 1 (default): the first few huge sub-triangles of a camera view (the background plane, close-up faces) are
     * resolved per pixel inside the shade kernel instead of being tile-binned; 0: everything large is tile-binned.
     * Results are bit-identical. */
    SLB_OPT_HUGE_IN_SHADE = 7,
    /* 1 (default): those huge sub-triangles are shaded from per-frame records holding the vertex stage's outputs
     * (k_huge_prepare) instead of a per-pixel re-set-up. ID maps identical; float targets equal within rounding. */
    SLB_OPT_HUGE_PREPARE = 8,
    /* 1 (default): every shadow map carries a block-occupancy mask (one bit per 8x8 texels); PCF footprints over
     * untouched blocks skip their 25 taps. Results are bit-identical. */
    SLB_OPT_SHADOW_MASK = 9,
    /* 1 (default): slb_render_batch queues the first phase of every sub-batch (upload, clear, set-up, scan) on an auxiliary stream of the
       context, so that it overlaps the shade pass of the previous sub-batch (measured: +4 % on config C3); 0: everything on one stream */
    SLB_OPT_OVERLAP = 10
};
int slb_ctx_set_option(slb_ctx* ctx, int option, int64_t value);

/* ---- config 4: render-and-compare helpers (reference: python/src/diff.cu, bridge_diff.cpp) -- */

/* valid[h,w] = 0 iff pixel is an object pixel and a 3x3 neighbour belongs to a different
 * non-zero instance with smaller depth (reference: python/src/diff.cu:13-99).
 * instance_index: int16 HxW, depth: float32 HxW, valid_out: uint8 HxW; all DEVICE pointers. */
int slb_diff_sobel_valid_mask(slb_ctx* ctx, const int16_t* instance_index, const float* depth,
                              uint8_t* valid_out, int32_t height, int32_t width, void* stream);
/* 3x3 dilation of a per-object mask gated on the Sobel-valid mask, copying a neighbour's object
 * coordinate into newly covered pixels (reference: python/src/diff.cu:101-193).
 * mask/valid: uint8 HxW; coords: float32 HxWx3 (pixel stride coord_stride floats). */
int slb_diff_dilate_object_mask(slb_ctx* ctx, const uint8_t* mask, const uint8_t* valid, const float* coords,
                                int32_t coord_stride, uint8_t* mask_out, float* coords_out, int32_t height,
                                int32_t width, void* stream);

/* Fused render-and-compare backward: gradient of an objective w.r.t. the locally linearised object poses
 * T(alpha,beta,gamma,a,b,c) = T0 [[1,-gamma,beta,a],[gamma,1,-alpha,b],[-beta,alpha,1,c],[0,0,0,1]]
 * from dObjective/dImage. Replaces the per-object Python loop of the reference
 * (python/stillleben/diff.py:73-127 compute_image_space_gradients, :355-523
 * backpropagate_gradient_to_poses; masks as python/src/diff.cu) with one pass over the pixels.
 * rgb: uint8 HxWx4, instance_index: int16 HxW, coord_depth: float32 HxWx4 (object xyz, w = depth),
 * grad_image: float32 3xHxW, grad_out: float32 n_objects x 6 — all DEVICE pointers.
 * projection (16 floats) and poses (n_objects x 16), column-major as in slb_scene_desc, and
 * instance_ids (n_objects) are HOST arrays. */
int slb_diff_pose_grad(slb_ctx* ctx, const uint8_t* rgb, const int16_t* instance_index, const float* coord_depth,
                       const float* grad_image, const float* projection, const float* poses,
                       const int32_t* instance_ids, int32_t n_objects, float* grad_out, int32_t height,
                       int32_t width, void* stream);

/* ---- camera noise model (reference: python/stillleben/camera_model.py) ---------------------- */

enum {                          /* stages of process_deterministic, in order (camera_model.py:224-262) */
    SLB_CAM_CHROMATIC = 1,      /* chromatic_aberration :46-73  */
    SLB_CAM_BLUR = 2,           /* blur(blur_sigma) if blur_sigma > 0 :106-119 */
    SLB_CAM_EXPOSURE = 4,       /* exposure :121-131 */
    SLB_CAM_NOISE = 8,          /* noise(a, b) if do_noise :133-161 */
    SLB_CAM_CLAMP = 16,         /* clamp to [0,1] :249 */
    SLB_CAM_HUE = 32,           /* color_jitter :163-222 */
    SLB_CAM_POST_BLUR = 64,     /* blur(0.4) + clamp :253-260 */
    SLB_CAM_ALL = 127
};
typedef struct slb_camera_params {
    float chromatic_translation[3][2]; /* (tx, ty) per channel, normalised [-1,1] image units */
    float chromatic_scaling[3];
    float blur_sigma;
    float exposure_deltaS;
    int32_t do_noise;
    float noise_a, noise_b;            /* var = a * y (Poissonian part), std b (Gaussian part) */
    float hue_shift;                   /* -0.5 .. 0.5 */
    uint32_t stages;                   /* SLB_CAM_* mask; SLB_CAM_ALL = process_deterministic */
    uint64_t seed;                     /* counter-RNG seed of the noise stage */
} slb_camera_params;

/* Applies the camera model to n images of H x W pixels in one launch pair. in_format 0: planar float32
 * [n][3][H][W] in [0,1] (what the reference takes); 1: RGBA8 [n][H][W][4] (the render target, /255 folded in).
 * out: planar float32 [n][3][H][W]. in / out are DEVICE pointers (out may not alias in); params is a HOST
 * array of n records. */
int slb_camera_model(slb_ctx* ctx, const void* in, int32_t in_format, float* out, int32_t n_images, int32_t height,
                     int32_t width, const slb_camera_params* params, void* stream);

/* ---- batched PNG encoder (reference: src/image_saver.cpp, python/src/py_image_saver.cpp:37-99) -------- */

/* Upper bound of one encoded file for an image of the given shape. */
size_t slb_png_bound(int32_t height, int32_t width, int32_t channels, int32_t bytes_per_channel);
/* Encodes n images of identical shape into n complete PNG files in DEVICE memory. images: n contiguous
 * HxWxC arrays, uint8 (C = 1, 3, 4: grey, RGB, RGBA) or, with bytes_per_channel = 2, 16-bit HxW (C = 1) as the
 * reference's binding accepts them; row 0 is the top row of the file. File i is written to
 * out + i * out_stride (out_stride >= slb_png_bound(...)) and its size to sizes[i]; images / out / sizes are
 * DEVICE pointers. */
int slb_png_encode(slb_ctx* ctx, const void* images, int32_t n_images, int32_t height, int32_t width, int32_t channels,
                   int32_t bytes_per_channel, uint8_t* out, size_t out_stride, uint32_t* sizes, void* stream);

/* ---- batched JPEG encoder (reference: src/image_saver.cpp:55-97 -> AnyImageConverter -> JpegImageConverter = libjpeg,
 * jpeg_set_defaults + jpeg_set_quality(80): baseline, 4:2:0, slow-integer DCT, Annex K Huffman tables) ------------- */

/* Upper bound of one encoded file for an image of the given shape (every block at its longest code, every byte stuffed). */
size_t slb_jpeg_bound(int32_t height, int32_t width, int32_t channels);
/* Encodes n images of identical shape into n complete JFIF files in DEVICE memory, byte-identical to what libjpeg writes for
 * the same pixels and quality. images: n contiguous HxWxC uint8 arrays, C = 1 (grey), 3 (RGB) or 4 (RGBA: alpha ignored, as
 * JpegImageConverter does); row 0 is the top row of the file. quality: 1..100 (the reference's converter default is 80).
 * File i is written to out + i * out_stride and its size to sizes[i]. out_stride may be smaller than slb_jpeg_bound() (the
 * worst case is several times the raw image; real files are a fraction of it): an image whose file does not fit gets
 * sizes[i] = 0 and nothing is written beyond its stride — the caller retries that image with a larger stride.
 * images / out / sizes are DEVICE pointers. */
int slb_jpeg_encode(slb_ctx* ctx, const void* images, int32_t n_images, int32_t height, int32_t width, int32_t channels,
                    int32_t quality, uint8_t* out, size_t out_stride, uint32_t* sizes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLB_H */
