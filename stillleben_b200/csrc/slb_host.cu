// slb_host.cu — the C ABI of include/slb.h: context, asset upload, result buffers and the host side
// of the hot path (marshalling a batch of scene descriptors into DFrame/DDraw arrays and queueing
// the kernels). Replaces the host part of sl::RenderPass::render (reference: src/render_pass.cpp:303-796)
// and of Mesh::loadVisual / LightMap::load / CUDATexture (src/mesh.cpp:624-745, src/light_map.cpp:266-611,
// src/cuda_interop.cpp:83-208). No exception crosses the boundary; every entry point returns a status.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/slb.h"
#include "hostmath.h"
#include "kernels.h"
#include "slb_dev.h"

using namespace slbk;

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes, bool zero = false) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaDeviceSynchronize(); cudaFree(p); }   // growth is rare; queued work of an earlier sub-batch may still use the old block
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        if (zero) e = cudaMemset(p, 0, want);
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct slb_texture {
    DTexture h; DTexture* d = nullptr; uint8_t* px = nullptr;
};
struct slb_mesh {
    float4* pos4 = nullptr; float4* attr = nullptr; uint32_t* idx = nullptr;
    float4* col4 = nullptr;                                   // vertex colours (never read by the renderer)
    // vertex-edit path (built on first use): vertex -> faces adjacency in ascending face order (CSR), per-face scratch
    uint32_t* adj_off = nullptr; uint32_t* adj_face = nullptr; float4* face_n = nullptr;
    uint32_t n_vertices = 0, n_indices = 0;
    std::vector<slb_submesh> submeshes;
    std::vector<slb_material> materials;
    std::vector<slb_texture*> textures;
    float bbox_min[3], bbox_max[3];
};
struct slb_lightmap {
    DLightMap h; DLightMap* d = nullptr;
    std::vector<void*> allocs;
    int env_size = 0, irr_size = 0, pre_size = 0, lut_size = 0;
};
struct slb_result {
    int32_t W = 0, H = 0, n_frames = 0; uint32_t mask = 0;
    void* ptrs[SLB_NUM_TARGETS]; bool owned[SLB_NUM_TARGETS];
    float4* hdr = nullptr;   // [n_frames][H][W], only with SLB_OPT_KEEP_HDR
};
static const size_t kTargetBpp[SLB_NUM_TARGETS] = {4, 16, 2, 2, 16, 16, 16, 16};

enum { ST_SHADOW = 0, ST_COUNT, ST_SCAN, ST_EMIT, ST_RASTER, ST_SHADE, ST_SSAO, ST_POST, ST_N };

// Everything one sub-batch owns on the device and in pinned host memory. A context has TWO sets, so that the
// host can marshal, upload and queue the set-up of sub-batch k+1 while it waits for the scan totals of sub-batch k
// (slb_render_batch): the one host synchronisation per sub-batch no longer idles the GPU.
struct Pending {   // a sub-batch whose first phase (clear, set-up, scan) is queued and whose second phase is not
    bool active = false;
    int n = 0, W = 0, H = 0, first_frame = 0;
    size_t npx = 0;
    RasterGrid grid;
    uint32_t n_tiles = 0, normal_cap = 0, huge_cap = 0;
    uint64_t total_bin_tris = 0;
    slb_result* result = nullptr;
    bool post = false;
};
struct Scratch {
    DevBuf frames_d, draws_d, chunk_base_d, views_d, bdraws_d, scan_sums, active_tiles, scan_totals, survivors;
    DevBuf clip_recs, clip_counts, huge_recs, huge_counts, huge_shade, shadow_mask;
    DevBuf tile_count, tile_off, pairs, keys, hdr, scratch_normal, scratch_cam, zplane, ao, avg, mip_a, mip_b, shadow_maps;
    void* staging = nullptr; size_t staging_cap = 0;   // pinned
    uint32_t* total_pinned = nullptr;                   // mapped pinned: the scan writes its totals here
    uint32_t shadow_gen = 255;   // generation of the shadow-map pool (DFrame::shadow_tagbits); 255 = clear before the next use
    void* shadow_pool_at = nullptr; size_t shadow_pool_cap = 0;
    void* batch = nullptr; void (*batch_delete)(void*) = nullptr;   // host arrays of the sub-batch (Batch, defined below)
    cudaEvent_t scanned = nullptr;
    cudaEvent_t p1_done = nullptr, p2_done = nullptr;   // two-stream pipeline: first phase queued (aux stream) / second phase queued (render stream)
    bool p2_recorded = false;
    Pending pending;
    void release() {
        DevBuf* bufs[] = {&frames_d, &draws_d, &chunk_base_d, &views_d, &bdraws_d, &scan_sums, &active_tiles, &scan_totals, &survivors, &clip_recs,
                          &clip_counts, &huge_recs, &huge_counts, &huge_shade, &shadow_mask, &tile_count, &tile_off, &pairs, &keys, &hdr, &scratch_normal, &scratch_cam,
                          &zplane, &ao, &avg, &mip_a, &mip_b, &shadow_maps};
        for (DevBuf* b : bufs) b->release();
        if (staging) cudaFreeHost(staging);
        if (total_pinned) cudaFreeHost(total_pinned);
        if (batch && batch_delete) batch_delete(batch);
        if (scanned) cudaEventDestroy(scanned);
        if (p1_done) cudaEventDestroy(p1_done);
        if (p2_done) cudaEventDestroy(p2_done);
        p1_done = p2_done = nullptr; p2_recorded = false;
        staging = nullptr; total_pinned = nullptr; batch = nullptr; scanned = nullptr;
    }
};

struct slb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // first phases of the sub-batch pipeline when SLB_OPT_OVERLAP is on (slb_render_batch)
    cudaEvent_t call_start = nullptr;
    bool overlap = true;
    // All per-context scratch (two Scratch sets, shadow pool + generation tags, staging, diff / camera / PNG buffers) is
    // shared between calls, so calls are ORDERED even when they name different streams: enter_stream() makes the
    // stream of an entry point wait for everything the previous entry point queued. last_stream = where that was.
    cudaStream_t last_stream = nullptr;
    cudaEvent_t order_event = nullptr;
    std::string err;
    bool time_kernels = false, keep_hdr = false;
    int max_subbatch = 64;
    // largest pixel box (in pixels) the setup kernel rasterises directly with one thread / with one warp; larger
    // triangles take the tiled path (SLB_OPT_DIRECT_MAX, SLB_OPT_WARP_MAX)
    int direct_max = 128, warp_max = 4096;
    bool lean_shade = true;
    bool huge_in_shade = true;   // camera views resolve their first SLB_HUGE_PER_VIEW huge sub-triangles in the shade kernel
    bool huge_prepare = true;    // ... and shade them from per-frame HugeShade records (k_huge_prepare) instead of a per-pixel re-set-up
    bool shadow_mask = true;     // block-occupancy masks of the shadow maps: PCF footprints over untouched blocks skip their 25 taps
    slb_stats stats;
    // assets owned by the context
    slb_mesh* plane = nullptr;
    Scratch scr[2];
    DevBuf diff_params, diff_partial, cam_params, cam_mid, png_rows, png_info, png_offsets;
    DevBuf jpeg_coefs, jpeg_off, jpeg_total, jpeg_stream, jpeg_header;
    int jpeg_key[4] = {0, 0, 0, 0};   // (H, W, channels, quality) the header / constant tables were built for
    int jpeg_header_len = 0;
    JpegQuant jpeg_quant;
    bool png_tables = false;
    // timing
    struct Ev { int stage; cudaEvent_t a, b; };
    std::vector<Ev> events;
    std::vector<cudaEvent_t> event_pool;
    float time_acc[16] = {0};   // stage times already drained from `events` (bounded: see StageTimer)
    // host-render slots
    slb_result* slot[2] = {nullptr, nullptr};
    cudaEvent_t slot_rendered[2] = {nullptr, nullptr}, slot_copied[2] = {nullptr, nullptr};
};
static thread_local std::string g_create_error;

// The stream an entry point queues its work on (NULL = the context's own), ordered after the previous entry point's work.
static cudaStream_t enter_stream(slb_ctx* ctx, void* stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (ctx->last_stream && s != ctx->last_stream && ctx->order_event) {
        cudaEventRecord(ctx->order_event, ctx->last_stream);
        cudaStreamWaitEvent(s, ctx->order_event, 0);
    }
    ctx->last_stream = s;
    return s;
}
// Wait for everything queued through this context, on whichever stream it went (read-back, destroy, update paths).
static cudaError_t sync_ctx(slb_ctx* ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (ctx->last_stream && ctx->last_stream != ctx->stream) { cudaError_t e2 = cudaStreamSynchronize(ctx->last_stream); if (e == cudaSuccess) e = e2; }
    if (ctx->aux_stream) { cudaError_t e2 = cudaStreamSynchronize(ctx->aux_stream); if (e == cudaSuccess) e = e2; }   // already joined into the render stream; free
    return e;
}

static int fail(slb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SLB_ERR_OUT_OF_MEMORY : SLB_ERR_CUDA,        \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                                \
    } while (0)

static int ensure_staging(slb_ctx* ctx, Scratch& S, size_t bytes) {
    if (bytes <= S.staging_cap) return SLB_OK;
    if (S.staging) cudaFreeHost(S.staging);
    S.staging = nullptr; S.staging_cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CU(cudaHostAlloc(&S.staging, want, cudaHostAllocDefault));
    S.staging_cap = want;
    return SLB_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static void ssao_tables(float* noise, float* kernel) {   // reference: src/shaders/ssao_shader.cpp:72-112
    std::mt19937 random{0xdeadbeef};
    std::uniform_real_distribution<float> rf(0.0f, 1.0f);
    for (int i = 0; i < 16; ++i) { float a = 2.0f * rf(random) - 1.0f; float b = 2.0f * rf(random) - 1.0f; noise[i * 3] = a; noise[i * 3 + 1] = b; noise[i * 3 + 2] = 0.0f; }
    for (int i = 0; i < 64; ++i) {
        float a = 2.0f * rf(random) - 1.0f; float b = 2.0f * rf(random) - 1.0f; float c = rf(random);
        hm::Vec3 s = hm::normalize(hm::v3(a, b, c)) * rf(random);
        float scale = (float)i / 64.0f;
        float t = scale * scale;
        s = s * (0.1f * (1.0f - t) + 1.0f * t);
        kernel[i * 3] = s.x; kernel[i * 3 + 1] = s.y; kernel[i * 3 + 2] = s.z;
    }
}

extern "C" int slb_abi_version(void) { return SLB_ABI_VERSION; }

extern "C" const char* slb_last_error(const slb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int slb_ctx_device(const slb_ctx* ctx) { return ctx ? ctx->device : -1; }

extern "C" int slb_ctx_create(int device, slb_ctx** out) {
    slb_ctx* ctx = nullptr;
    if (!out) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(ctx, SLB_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_ctx_create: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(ctx, SLB_ERR_RUNTIME, std::string("stillleben_b200 is built for sm_100a only; device is ") + prop.name);
    slb_ctx* c = new slb_ctx;
    c->device = device;
    std::memset(&c->stats, 0, sizeof c->stats);
    ctx = c;
    cudaError_t e1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaError_t e2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&c->call_start, cudaEventDisableTiming);
    cudaError_t e3 = cudaEventCreateWithFlags(&c->order_event, cudaEventDisableTiming);
    for (Scratch& S : c->scr) {
        if (e3 == cudaSuccess) e3 = cudaHostAlloc((void**)&S.total_pinned, 64, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&S.scanned, cudaEventDisableTiming);
        if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&S.p1_done, cudaEventDisableTiming);
        if (e3 == cudaSuccess) e3 = cudaEventCreateWithFlags(&S.p2_done, cudaEventDisableTiming);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { for (Scratch& S : c->scr) S.release(); delete c; ctx = nullptr; return fail(ctx, SLB_ERR_CUDA, "slb_ctx_create: stream/pinned allocation failed"); }
    float noise[48], kernel[192];
    ssao_tables(noise, kernel);
    upload_ssao_tables(noise, kernel);
    // background plane: Magnum Primitives::planeSolid(TextureCoordinates) — strip (1,-1)(1,1)(-1,-1)(-1,1),
    // +Z normal, no tangent / vertex-id attribute (contrib/magnum/src/Magnum/Primitives/Plane.cpp:36-60)
    {
        struct V68 { float pos[3], uv[2], color[4], tangent[4]; uint32_t id; float normal[3]; } v[4];
        static_assert(sizeof(V68) == SLB_VERTEX_STRIDE, "vertex stride");
        const float P[4][2] = {{1, -1}, {1, 1}, {-1, -1}, {-1, 1}};
        const float U[4][2] = {{1, 0}, {1, 1}, {0, 0}, {0, 1}};
        std::memset(v, 0, sizeof v);
        for (int i = 0; i < 4; ++i) {
            v[i].pos[0] = P[i][0]; v[i].pos[1] = P[i][1]; v[i].uv[0] = U[i][0]; v[i].uv[1] = U[i][1];
            v[i].color[3] = 1.0f; v[i].tangent[3] = 1.0f; v[i].normal[2] = 1.0f; v[i].id = 0;
        }
        const uint32_t idx[6] = {0, 1, 2, 2, 1, 3};
        slb_submesh sm = {0, 6, -1, 0};
        const float bmin[3] = {-1, -1, 0}, bmax[3] = {1, 1, 0};
        int rc = slb_mesh_upload(c, v, 4, idx, 6, &sm, 1, nullptr, 0, nullptr, 0, bmin, bmax, &c->plane);
        if (rc != SLB_OK) { g_create_error = c->err; slb_ctx_destroy(c); return rc; }
    }
    *out = c;
    return SLB_OK;
}

extern "C" int slb_ctx_synchronize(slb_ctx* ctx) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(sync_ctx(ctx));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return SLB_OK;
}

extern "C" void slb_ctx_destroy(slb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->plane) slb_mesh_destroy(ctx, ctx->plane);
    for (int i = 0; i < 2; ++i) {
        if (ctx->slot[i]) slb_result_destroy(ctx, ctx->slot[i]);
        if (ctx->slot_rendered[i]) cudaEventDestroy(ctx->slot_rendered[i]);
        if (ctx->slot_copied[i]) cudaEventDestroy(ctx->slot_copied[i]);
    }
    for (Scratch& S : ctx->scr) S.release();
    DevBuf* bufs[] = {&ctx->diff_params, &ctx->diff_partial, &ctx->cam_params, &ctx->cam_mid, &ctx->png_rows, &ctx->png_info, &ctx->png_offsets,
                      &ctx->jpeg_coefs, &ctx->jpeg_off, &ctx->jpeg_total, &ctx->jpeg_stream, &ctx->jpeg_header};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : ctx->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->order_event) cudaEventDestroy(ctx->order_event);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->call_start) cudaEventDestroy(ctx->call_start);
    delete ctx;
}

extern "C" int slb_ctx_set_option(slb_ctx* ctx, int option, int64_t value) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    switch (option) {
        case SLB_OPT_TIME_KERNELS: ctx->time_kernels = value != 0; return SLB_OK;
        case SLB_OPT_KEEP_HDR: ctx->keep_hdr = value != 0; return SLB_OK;
        case SLB_OPT_MAX_SUBBATCH:
            if (value < 1 || value > 1024) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "SLB_OPT_MAX_SUBBATCH must be in [1, 1024]");
            ctx->max_subbatch = (int)value;
            for (int i = 0; i < 2; ++i) if (ctx->slot[i]) { slb_result_destroy(ctx, ctx->slot[i]); ctx->slot[i] = nullptr; }
            return SLB_OK;
        case SLB_OPT_DIRECT_MAX:
        case SLB_OPT_WARP_MAX:
            if (value < 0 || value > 4096) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "SLB_OPT_DIRECT_MAX / SLB_OPT_WARP_MAX must be in [0, 4096]");
            (option == SLB_OPT_DIRECT_MAX ? ctx->direct_max : ctx->warp_max) = (int)value;
            return SLB_OK;
        case SLB_OPT_LEAN_SHADE: ctx->lean_shade = value != 0; return SLB_OK;
        case SLB_OPT_HUGE_PREPARE: ctx->huge_prepare = value != 0; return SLB_OK;
        case SLB_OPT_SHADOW_MASK: ctx->shadow_mask = value != 0; return SLB_OK;
        case SLB_OPT_OVERLAP: ctx->overlap = value != 0; return SLB_OK;
        case SLB_OPT_HUGE_IN_SHADE: ctx->huge_in_shade = value != 0; return SLB_OK;
        default: return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "unknown option");
    }
}

static void collect_times(slb_ctx* ctx);
extern "C" int slb_ctx_get_stats(slb_ctx* ctx, slb_stats* out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    collect_times(ctx);   // waits for the stage events recorded since the previous call (SLB_OPT_TIME_KERNELS)
    *out = ctx->stats;
    return SLB_OK;
}

extern "C" int slb_host_alloc(slb_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out || bytes == 0) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return SLB_OK;
}
extern "C" void slb_host_free(slb_ctx* ctx, void* ptr) {
    if (!ptr) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFreeHost(ptr);
}

// ---------------------------------------------------------------------------------------------
// textures / meshes
// ---------------------------------------------------------------------------------------------
static bool valid_image(const slb_image* im) {
    return im && im->pixels && im->width > 0 && im->height > 0 && (im->channels == 3 || im->channels == 4) &&
           im->wrap_s >= 0 && im->wrap_s <= 3 && im->wrap_t >= 0 && im->wrap_t <= 3 && im->min_filter >= 0 && im->min_filter <= 5 &&
           (im->mag_filter == SLB_FILTER_NEAREST || im->mag_filter == SLB_FILTER_LINEAR);
}

extern "C" int slb_texture_create(slb_ctx* ctx, const slb_image* image, int kind, slb_texture** out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!valid_image(image)) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_texture_create: unsupported image (RGB8/RGBA8 only)");
    if (kind != SLB_TEXTURE_2D && kind != SLB_TEXTURE_RECT) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_texture_create: bad kind");
    CU(cudaSetDevice(ctx->device));
    slb_texture* t = new slb_texture;
    std::memset(&t->h, 0, sizeof t->h);
    t->h.w = image->width; t->h.h = image->height;
    t->h.wrap_s = image->wrap_s; t->h.wrap_t = image->wrap_t; t->h.min_filter = image->min_filter; t->h.mag_filter = image->mag_filter;
    t->h.kind = kind; t->h.has_alpha = image->channels == 4;
    int lw = image->width, lh = image->height, nl = 0; size_t total = 0;
    std::vector<int> ws, hs;
    for (;;) {
        t->h.level_off[nl] = (uint32_t)total; ws.push_back(lw); hs.push_back(lh);
        total += (size_t)lw * lh; ++nl;
        if (kind != SLB_TEXTURE_2D || (lw == 1 && lh == 1) || nl == SLB_MAX_LEVELS) break;
        lw = lw > 1 ? lw >> 1 : 1; lh = lh > 1 ? lh >> 1 : 1;
    }
    t->h.n_levels = nl;
    void* raw = nullptr;
    size_t raw_bytes = (size_t)image->width * image->height * image->channels;
    cudaError_t e = cudaMalloc((void**)&t->px, total * 4);
    if (e == cudaSuccess) e = cudaMalloc(&raw, raw_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d, sizeof(DTexture));
    if (e == cudaSuccess) e = cudaMemcpyAsync(raw, image->pixels, raw_bytes, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) {
        launch_expand_rgba((const uint8_t*)raw, image->channels, t->px, (size_t)image->width * image->height, ctx->stream);
        for (int l = 1; l < nl; ++l)
            launch_mip_level(t->px + (size_t)t->h.level_off[l - 1] * 4, ws[l - 1], hs[l - 1], t->px + (size_t)t->h.level_off[l] * 4, ws[l], hs[l], ctx->stream);
        ctx->stats.kernel_launches += nl;
        t->h.px = t->px;
        e = cudaMemcpyAsync(t->d, &t->h, sizeof(DTexture), cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = sync_ctx(ctx);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (raw) cudaFree(raw);
    if (e != cudaSuccess) {
        if (t->px) cudaFree(t->px);
        if (t->d) cudaFree(t->d);
        delete t;
        return fail(ctx, e == cudaErrorMemoryAllocation ? SLB_ERR_OUT_OF_MEMORY : SLB_ERR_CUDA, std::string("slb_texture_create: ") + cudaGetErrorString(e));
    }
    *out = t;
    return SLB_OK;
}
// read back one level as RGBA8 (tests pin the mip chain bit-exactly against the oracle)
extern "C" int slb_texture_read_level(slb_ctx* ctx, const slb_texture* tex, int level, int32_t* w, int32_t* h, void* host_out) {
    if (!ctx || !tex || level < 0 || level >= tex->h.n_levels) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_texture_read_level: bad level");
    int lw = tex->h.w >> level, lh = tex->h.h >> level;
    lw = lw > 0 ? lw : 1; lh = lh > 0 ? lh : 1;
    if (w) *w = lw;
    if (h) *h = lh;
    if (host_out) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaMemcpy(host_out, tex->px + (size_t)tex->h.level_off[level] * 4, (size_t)lw * lh * 4, cudaMemcpyDeviceToHost));
    }
    return tex->h.n_levels;
}

extern "C" void slb_texture_destroy(slb_ctx* ctx, slb_texture* tex) {
    if (!tex) return;
    if (ctx) { cudaSetDevice(ctx->device); sync_ctx(ctx); }
    if (tex->px) cudaFree(tex->px);
    if (tex->d) cudaFree(tex->d);
    delete tex;
}

extern "C" int slb_mesh_upload(slb_ctx* ctx, const void* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                               const slb_submesh* submeshes, uint32_t n_submeshes, const slb_material* materials, uint32_t n_materials,
                               const slb_image* images, uint32_t n_images, const float bbox_min[3], const float bbox_max[3], slb_mesh** out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!vertices || !indices || n_vertices == 0 || !submeshes || n_submeshes == 0 || !bbox_min || !bbox_max)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_upload: missing vertices / indices / submeshes / bbox");
    for (uint32_t i = 0; i < n_submeshes; ++i) {
        const slb_submesh& s = submeshes[i];
        if (s.index_count % 3 != 0 || (uint64_t)s.index_offset + s.index_count > n_indices)
            return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_upload: submesh index range invalid");
        if (s.material >= (int32_t)n_materials) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_upload: submesh material out of range");
    }
    for (uint32_t i = 0; i < n_materials; ++i) {
        const int32_t tx[5] = {materials[i].tex_base_color, materials[i].tex_normal, materials[i].tex_metallic_roughness,
                               materials[i].tex_emissive, materials[i].tex_occlusion};
        for (int k = 0; k < 5; ++k)
            if (tx[k] >= (int32_t)n_images) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_upload: material texture index out of range");
    }
    CU(cudaSetDevice(ctx->device));
    slb_mesh* m = new slb_mesh;
    m->n_vertices = n_vertices; m->n_indices = n_indices;
    m->submeshes.assign(submeshes, submeshes + n_submeshes);
    if (n_materials) m->materials.assign(materials, materials + n_materials);
    for (int k = 0; k < 3; ++k) { m->bbox_min[k] = bbox_min[k]; m->bbox_max[k] = bbox_max[k]; }
    void* raw = nullptr;
    cudaError_t e = cudaMalloc((void**)&m->pos4, (size_t)n_vertices * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->attr, (size_t)n_vertices * 3 * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->col4, (size_t)n_vertices * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->idx, (size_t)(n_indices ? n_indices : 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&raw, (size_t)n_vertices * SLB_VERTEX_STRIDE + sizeof(uint32_t));
    uint32_t* d_max = raw ? reinterpret_cast<uint32_t*>((uint8_t*)raw + (size_t)n_vertices * SLB_VERTEX_STRIDE) : nullptr;   // 68 n is 4-byte aligned
    uint32_t h_max = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(raw, vertices, (size_t)n_vertices * SLB_VERTEX_STRIDE, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->idx, indices, (size_t)n_indices * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) {
        launch_repack_vertices((const uint8_t*)raw, n_vertices, m->pos4, m->attr, m->col4, ctx->stream);
        launch_index_max(m->idx, n_indices, d_max, ctx->stream);
        ctx->stats.kernel_launches += 2;
        e = cudaMemcpyAsync(&h_max, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = sync_ctx(ctx);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (raw) cudaFree(raw);
    int rc = SLB_OK;
    if (e != cudaSuccess) rc = fail(ctx, e == cudaErrorMemoryAllocation ? SLB_ERR_OUT_OF_MEMORY : SLB_ERR_CUDA, std::string("slb_mesh_upload: ") + cudaGetErrorString(e));
    // every kernel dereferences pos4[idx] / attr[idx] unchecked: reject an index buffer that points past the vertices
    else if (n_indices && h_max >= n_vertices) rc = fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_upload: index value out of range (>= n_vertices)");
    // textures:
    for (uint32_t i = 0; rc == SLB_OK && i < n_images; ++i) {
        slb_texture* t = nullptr;
        rc = slb_texture_create(ctx, &images[i], SLB_TEXTURE_2D, &t);
        if (rc == SLB_OK) m->textures.push_back(t);
    }
    if (rc != SLB_OK) { std::string keep = ctx->err; slb_mesh_destroy(ctx, m); ctx->err = keep; return rc; }
    *out = m;
    return SLB_OK;
}

extern "C" int slb_mesh_update_vertices(slb_ctx* ctx, slb_mesh* mesh, const void* vertices, uint32_t n_vertices) {
    if (!ctx || !mesh || !vertices) return SLB_ERR_INVALID_ARGUMENT;
    if (n_vertices != mesh->n_vertices) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_update_vertices: vertex count must not change");
    CU(cudaSetDevice(ctx->device));
    void* raw = nullptr;
    CU(cudaMalloc(&raw, (size_t)n_vertices * SLB_VERTEX_STRIDE));
    cudaError_t e = cudaMemcpyAsync(raw, vertices, (size_t)n_vertices * SLB_VERTEX_STRIDE, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) {
        launch_repack_vertices((const uint8_t*)raw, n_vertices, mesh->pos4, mesh->attr, mesh->col4, ctx->stream);
        ctx->stats.kernel_launches += 1;
        e = sync_ctx(ctx);
    }
    cudaFree(raw);
    if (e != cudaSuccess) return fail(ctx, SLB_ERR_CUDA, std::string("slb_mesh_update_vertices: ") + cudaGetErrorString(e));
    return SLB_OK;
}

// vertex -> faces adjacency in ascending face order (the summation order of Mesh::recomputeNormals, mesh.cpp:776-816),
// built once per mesh on the host from the resident index buffer
static int mesh_build_adjacency(slb_ctx* ctx, slb_mesh* m) {
    if (m->adj_off) return SLB_OK;
    std::vector<uint32_t> idx(m->n_indices), off(m->n_vertices + 1, 0u), faces(m->n_indices);
    CU(cudaMemcpy(idx.data(), m->idx, (size_t)m->n_indices * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < m->n_indices; ++i) off[idx[i] + 1]++;
    for (uint32_t v = 0; v < m->n_vertices; ++v) off[v + 1] += off[v];
    std::vector<uint32_t> cur(off.begin(), off.end() - 1);
    for (uint32_t i = 0; i < m->n_indices; ++i) faces[cur[idx[i]]++] = i / 3;     // face order ascending; a face listing a vertex twice counts twice
    CU(cudaMalloc((void**)&m->adj_off, off.size() * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&m->adj_face, (faces.size() ? faces.size() : 1) * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&m->face_n, (size_t)(m->n_indices / 3 ? m->n_indices / 3 : 1) * sizeof(float4)));
    CU(cudaMemcpy(m->adj_off, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->adj_face, faces.data(), faces.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return SLB_OK;
}
static cudaStream_t mesh_stream(slb_ctx* ctx, void* stream) { return enter_stream(ctx, stream); }

extern "C" int slb_mesh_recompute_normals(slb_ctx* ctx, slb_mesh* mesh, void* stream) {
    if (!ctx || !mesh) return SLB_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    int rc = mesh_build_adjacency(ctx, mesh);
    if (rc != SLB_OK) return rc;
    cudaStream_t s = mesh_stream(ctx, stream);
    launch_recompute_normals(mesh->pos4, mesh->idx, mesh->n_indices / 3, mesh->adj_off, mesh->adj_face, mesh->face_n, mesh->n_vertices, mesh->attr, s);
    ctx->stats.kernel_launches += 2;
    CU(cudaGetLastError());
    return SLB_OK;
}

extern "C" int slb_mesh_update_positions_and_colors(slb_ctx* ctx, slb_mesh* mesh, const int32_t* vertex_ids, uint32_t n,
                                                    const float* position_update, const float* color_update, void* stream) {
    if (!ctx || !mesh || (n && !vertex_ids)) return SLB_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = mesh_stream(ctx, stream);
    if (n && (position_update || color_update)) {
        // the three arrays may live in host or device memory (UVA): stage them on the device
        const size_t b_ids = (size_t)n * 4, b_pos = position_update ? (size_t)n * 12 : 0, b_col = color_update ? (size_t)n * 16 : 0;
        uint8_t* stage = nullptr;
        CU(cudaMallocAsync((void**)&stage, b_ids + b_pos + b_col + 4, s));
        uint32_t* d_err = reinterpret_cast<uint32_t*>(stage + b_ids + b_pos + b_col);
        CU(cudaMemcpyAsync(stage, vertex_ids, b_ids, cudaMemcpyDefault, s));
        if (b_pos) CU(cudaMemcpyAsync(stage + b_ids, position_update, b_pos, cudaMemcpyDefault, s));
        if (b_col) CU(cudaMemcpyAsync(stage + b_ids + b_pos, color_update, b_col, cudaMemcpyDefault, s));
        CU(cudaMemsetAsync(d_err, 0, 4, s));
        launch_vertex_delta(reinterpret_cast<const int32_t*>(stage), n, b_pos ? reinterpret_cast<const float*>(stage + b_ids) : nullptr,
                            b_col ? reinterpret_cast<const float*>(stage + b_ids + b_pos) : nullptr, mesh->pos4, mesh->col4, mesh->n_vertices, d_err, s);
        ctx->stats.kernel_launches += 1;
        uint32_t h_err = 0;
        CU(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaFreeAsync(stage, s));
        CU(cudaStreamSynchronize(s));
        if (h_err) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_update_positions_and_colors: vertex id out of range (ids are one-based)");
    }
    if (position_update && n) return slb_mesh_recompute_normals(ctx, mesh, stream);    // mesh.cpp:840-841
    return SLB_OK;
}

extern "C" int slb_mesh_set_positions(slb_ctx* ctx, slb_mesh* mesh, const float* positions, uint32_t n_vertices, void* stream) {
    if (!ctx || !mesh || !positions) return SLB_ERR_INVALID_ARGUMENT;
    if (n_vertices != mesh->n_vertices)   // mesh.cpp:862-863
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "Number of new vertices should match the existing mesh vertices");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = mesh_stream(ctx, stream);
    float* stage = nullptr;
    CU(cudaMallocAsync((void**)&stage, (size_t)n_vertices * 12, s));
    CU(cudaMemcpyAsync(stage, positions, (size_t)n_vertices * 12, cudaMemcpyDefault, s));
    launch_set_positions(stage, n_vertices, mesh->pos4, s);
    ctx->stats.kernel_launches += 1;
    CU(cudaFreeAsync(stage, s));
    return slb_mesh_recompute_normals(ctx, mesh, stream);                                // mesh.cpp:868
}

extern "C" int slb_mesh_set_colors(slb_ctx* ctx, slb_mesh* mesh, const float* colors, uint32_t n_vertices, void* stream) {
    if (!ctx || !mesh || !colors) return SLB_ERR_INVALID_ARGUMENT;
    if (n_vertices != mesh->n_vertices)   // mesh.cpp:878-879
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "Number of new vertices should match the existing mesh vertices for vertex color update");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = mesh_stream(ctx, stream);
    float* stage = nullptr;
    CU(cudaMallocAsync((void**)&stage, (size_t)n_vertices * 16, s));
    CU(cudaMemcpyAsync(stage, colors, (size_t)n_vertices * 16, cudaMemcpyDefault, s));
    launch_set_colors(stage, n_vertices, mesh->col4, s);
    ctx->stats.kernel_launches += 1;
    CU(cudaFreeAsync(stage, s));
    return SLB_OK;
}

extern "C" int slb_mesh_read_vertices(slb_ctx* ctx, const slb_mesh* mesh, void* vertices_out, uint32_t n_vertices) {
    if (!ctx || !mesh || !vertices_out) return SLB_ERR_INVALID_ARGUMENT;
    if (n_vertices != mesh->n_vertices) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_mesh_read_vertices: vertex count mismatch");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());      // edits may have been queued on a caller stream
    uint8_t* raw = nullptr;
    CU(cudaMalloc((void**)&raw, (size_t)n_vertices * SLB_VERTEX_STRIDE));
    launch_unpack_vertices(mesh->pos4, mesh->attr, mesh->col4, n_vertices, raw, ctx->stream);
    ctx->stats.kernel_launches += 1;
    cudaError_t e = cudaMemcpyAsync(vertices_out, raw, (size_t)n_vertices * SLB_VERTEX_STRIDE, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = sync_ctx(ctx);
    cudaFree(raw);
    if (e != cudaSuccess) return fail(ctx, SLB_ERR_CUDA, std::string("slb_mesh_read_vertices: ") + cudaGetErrorString(e));
    return SLB_OK;
}

extern "C" void slb_mesh_destroy(slb_ctx* ctx, slb_mesh* mesh) {
    if (!mesh) return;
    if (ctx) { cudaSetDevice(ctx->device); sync_ctx(ctx); }
    for (slb_texture* t : mesh->textures) slb_texture_destroy(ctx, t);
    if (mesh->pos4) cudaFree(mesh->pos4);
    if (mesh->attr) cudaFree(mesh->attr);
    if (mesh->idx) cudaFree(mesh->idx);
    if (mesh->col4) cudaFree(mesh->col4);
    if (mesh->adj_off) cudaFree(mesh->adj_off);
    if (mesh->adj_face) cudaFree(mesh->adj_face);
    if (mesh->face_n) cudaFree(mesh->face_n);
    delete mesh;
}

// ---------------------------------------------------------------------------------------------
// light maps
// ---------------------------------------------------------------------------------------------
static int lightmap_finish(slb_ctx* ctx, slb_lightmap* lm) {
    CU(cudaMemcpyAsync(lm->d, &lm->h, sizeof(DLightMap), cudaMemcpyHostToDevice, ctx->stream));
    return SLB_OK;
}
static int lm_alloc(slb_ctx* ctx, slb_lightmap* lm, size_t bytes, void** out) {
    CU(cudaMalloc(out, bytes));
    lm->allocs.push_back(*out);
    return SLB_OK;
}
static int lightmap_build_env_mips(slb_ctx* ctx, slb_lightmap* lm, float4* level0, int size) {
    lm->h.n_env = 0;
    float4* cur = level0; int s = size;
    for (;;) {
        lm->h.env[lm->h.n_env].px = cur; lm->h.env[lm->h.n_env].size = s; lm->h.n_env++;
        if (s == 1 || lm->h.n_env == 12) break;
        float4* nxt = nullptr;
        int rc = lm_alloc(ctx, lm, (size_t)6 * (s / 2) * (s / 2) * sizeof(float4), (void**)&nxt);
        if (rc != SLB_OK) return rc;
        launch_cube_mip(cur, s, nxt, ctx->stream);
        ctx->stats.kernel_launches += 1;
        cur = nxt; s /= 2;
    }
    return SLB_OK;
}

// sizes <= 0 select the reference's (512 / 32 / 128 / 512, 1024 samples)
extern "C" int slb_lightmap_create_ex(slb_ctx* ctx, const slb_lightmap_desc* desc, int env_size, int irr_size, int pre_size, int lut_size,
                                      int n_samples, slb_lightmap** out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!desc || !desc->equirect_rgb || desc->width <= 0 || desc->height <= 0 || desc->n_lights < 0 || desc->n_lights > SLB_NUM_LIGHTS)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_create: bad descriptor");
    env_size = env_size > 0 ? env_size : 512; irr_size = irr_size > 0 ? irr_size : 32; pre_size = pre_size > 0 ? pre_size : 128;
    lut_size = lut_size > 0 ? lut_size : 512; n_samples = n_samples > 0 ? n_samples : 1024;
    if ((env_size & (env_size - 1)) || (pre_size & (pre_size - 1)) || pre_size < 16)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_create: cube sizes must be powers of two (prefilter >= 16)");
    CU(cudaSetDevice(ctx->device));
    slb_lightmap* lm = new slb_lightmap;
    std::memset(&lm->h, 0, sizeof lm->h);
    lm->env_size = env_size; lm->irr_size = irr_size; lm->pre_size = pre_size; lm->lut_size = lut_size;
    lm->h.n_lights = desc->n_lights;
    std::memcpy(lm->h.light_directions, desc->light_directions, sizeof lm->h.light_directions);
    std::memcpy(lm->h.light_colors, desc->light_colors, sizeof lm->h.light_colors);
    int rc = SLB_OK;
    float* eq = nullptr; float4* env0 = nullptr; float4* irr = nullptr; float4* lut = nullptr;
    size_t eq_bytes = (size_t)desc->width * desc->height * 3 * sizeof(float);
    do {
        if ((rc = lm_alloc(ctx, lm, sizeof(DLightMap), (void**)&lm->d)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, eq_bytes, (void**)&eq)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, (size_t)6 * env_size * env_size * sizeof(float4), (void**)&env0)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, (size_t)6 * irr_size * irr_size * sizeof(float4), (void**)&irr)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, (size_t)lut_size * lut_size * sizeof(float4), (void**)&lut)) != SLB_OK) break;
        cudaError_t e = cudaMemcpyAsync(eq, desc->equirect_rgb, eq_bytes, cudaMemcpyDefault, ctx->stream);
        if (e != cudaSuccess) { rc = fail(ctx, SLB_ERR_CUDA, cudaGetErrorString(e)); break; }
        launch_equirect_to_cube(eq, desc->width, desc->height, env0, env_size, ctx->stream);
        ctx->stats.kernel_launches += 1;
        if ((rc = lightmap_build_env_mips(ctx, lm, env0, env_size)) != SLB_OK) break;
        lm->h.irr.px = irr; lm->h.irr.size = irr_size;
        lm->h.lut = lut; lm->h.lut_size = lut_size;
        for (int mip = 0; mip < 5; ++mip) {
            int n = pre_size >> mip;
            float4* p = nullptr;
            if ((rc = lm_alloc(ctx, lm, (size_t)6 * n * n * sizeof(float4), (void**)&p)) != SLB_OK) break;
            lm->h.pre[mip].px = p; lm->h.pre[mip].size = n;
        }
        if (rc != SLB_OK) break;
        if ((rc = lightmap_finish(ctx, lm)) != SLB_OK) break;   // env chain visible to the convolution kernels
        launch_irradiance(lm->d, irr, irr_size, log2f((float)env_size / (float)irr_size), ctx->stream);
        for (int mip = 0; mip < 5; ++mip)
            launch_prefilter(lm->d, const_cast<float4*>(lm->h.pre[mip].px), lm->h.pre[mip].size, (float)mip / 4.0f, n_samples, (float)env_size, ctx->stream);
        launch_brdf_lut(lut, lut_size, n_samples, ctx->stream);
        ctx->stats.kernel_launches += 7;
        cudaError_t e2 = sync_ctx(ctx);
        if (e2 == cudaSuccess) e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { rc = fail(ctx, SLB_ERR_CUDA, std::string("slb_lightmap_create: ") + cudaGetErrorString(e2)); break; }
    } while (0);
    if (rc != SLB_OK) { std::string keep = ctx->err; slb_lightmap_destroy(ctx, lm); ctx->err = keep; return rc; }
    *out = lm;
    return SLB_OK;
}
extern "C" int slb_lightmap_create(slb_ctx* ctx, const slb_lightmap_desc* desc, slb_lightmap** out) {
    return slb_lightmap_create_ex(ctx, desc, 0, 0, 0, 0, 0, out);
}

// A light map from maps that were precomputed elsewhere (another rank's slb_lightmap_create, read back with
// slb_lightmap_read and broadcast in the asset arena): copies them, rebuilds the environment's mip chain.
extern "C" int slb_lightmap_create_from_maps(slb_ctx* ctx, const slb_lightmap_desc* desc, const float* env0, int env_size, const float* irradiance,
                                             int irr_size, const float* prefilter, int pre_size, const float* lut, int lut_size, slb_lightmap** out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!desc || desc->n_lights < 0 || desc->n_lights > SLB_NUM_LIGHTS || !env0 || !irradiance || !prefilter || !lut)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_create_from_maps: bad arguments");
    if (env_size <= 0 || irr_size <= 0 || lut_size <= 0 || (env_size & (env_size - 1)) || (pre_size & (pre_size - 1)) || pre_size < 16)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_create_from_maps: cube sizes must be powers of two (prefilter >= 16)");
    CU(cudaSetDevice(ctx->device));
    slb_lightmap* lm = new slb_lightmap;
    std::memset(&lm->h, 0, sizeof lm->h);
    lm->env_size = env_size; lm->irr_size = irr_size; lm->pre_size = pre_size; lm->lut_size = lut_size;
    lm->h.n_lights = desc->n_lights;
    std::memcpy(lm->h.light_directions, desc->light_directions, sizeof lm->h.light_directions);
    std::memcpy(lm->h.light_colors, desc->light_colors, sizeof lm->h.light_colors);
    int rc = SLB_OK;
    do {
        float4 *e0 = nullptr, *ir = nullptr, *lu = nullptr;
        const size_t be = (size_t)6 * env_size * env_size * sizeof(float4), bi = (size_t)6 * irr_size * irr_size * sizeof(float4),
                     bl = (size_t)lut_size * lut_size * sizeof(float4);
        if ((rc = lm_alloc(ctx, lm, sizeof(DLightMap), (void**)&lm->d)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, be, (void**)&e0)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, bi, (void**)&ir)) != SLB_OK) break;
        if ((rc = lm_alloc(ctx, lm, bl, (void**)&lu)) != SLB_OK) break;
        cudaError_t e = cudaMemcpyAsync(e0, env0, be, cudaMemcpyDefault, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ir, irradiance, bi, cudaMemcpyDefault, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lu, lut, bl, cudaMemcpyDefault, ctx->stream);
        if (e != cudaSuccess) { rc = fail(ctx, SLB_ERR_CUDA, cudaGetErrorString(e)); break; }
        if ((rc = lightmap_build_env_mips(ctx, lm, e0, env_size)) != SLB_OK) break;
        lm->h.irr.px = ir; lm->h.irr.size = irr_size;
        lm->h.lut = lu; lm->h.lut_size = lut_size;
        const float* src = prefilter;
        for (int mip = 0; mip < 5 && rc == SLB_OK; ++mip) {
            const int n = pre_size >> mip;
            float4* p = nullptr;
            if ((rc = lm_alloc(ctx, lm, (size_t)6 * n * n * sizeof(float4), (void**)&p)) != SLB_OK) break;
            e = cudaMemcpyAsync(p, src, (size_t)6 * n * n * sizeof(float4), cudaMemcpyDefault, ctx->stream);
            if (e != cudaSuccess) { rc = fail(ctx, SLB_ERR_CUDA, cudaGetErrorString(e)); break; }
            lm->h.pre[mip].px = p; lm->h.pre[mip].size = n;
            src += (size_t)6 * n * n * 4;
        }
        if (rc != SLB_OK) break;
        if ((rc = lightmap_finish(ctx, lm)) != SLB_OK) break;
        cudaError_t e2 = sync_ctx(ctx);
        if (e2 == cudaSuccess) e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { rc = fail(ctx, SLB_ERR_CUDA, std::string("slb_lightmap_create_from_maps: ") + cudaGetErrorString(e2)); break; }
    } while (0);
    if (rc != SLB_OK) { std::string keep = ctx->err; slb_lightmap_destroy(ctx, lm); ctx->err = keep; return rc; }
    *out = lm;
    return SLB_OK;
}

extern "C" int slb_lightmap_read(slb_ctx* ctx, const slb_lightmap* lm, int which, float* host_out, size_t n_floats) {
    if (!ctx || !lm || !host_out) return SLB_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(sync_ctx(ctx));
    if (which == 0 || which == 1 || which == 3) {
        const void* src = which == 0 ? (const void*)lm->h.env[0].px : which == 1 ? (const void*)lm->h.irr.px : (const void*)lm->h.lut;
        size_t n = which == 0 ? (size_t)6 * lm->env_size * lm->env_size * 4 : which == 1 ? (size_t)6 * lm->irr_size * lm->irr_size * 4
                                                                                         : (size_t)lm->lut_size * lm->lut_size * 4;
        if (n_floats < n) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_read: buffer too small");
        CU(cudaMemcpy(host_out, src, n * sizeof(float), cudaMemcpyDeviceToHost));
        return SLB_OK;
    }
    if (which == 2) {
        size_t ofs = 0;
        for (int mip = 0; mip < 5; ++mip) {
            size_t n = (size_t)6 * lm->h.pre[mip].size * lm->h.pre[mip].size * 4;
            if (n_floats < ofs + n) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_read: buffer too small");
            CU(cudaMemcpy(host_out + ofs, lm->h.pre[mip].px, n * sizeof(float), cudaMemcpyDeviceToHost));
            ofs += n;
        }
        return SLB_OK;
    }
    return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_lightmap_read: which must be 0..3");
}
extern "C" int slb_lightmap_sizes(const slb_lightmap* lm, int32_t sizes[4]) {
    if (!lm || !sizes) return SLB_ERR_INVALID_ARGUMENT;
    sizes[0] = lm->env_size; sizes[1] = lm->irr_size; sizes[2] = lm->pre_size; sizes[3] = lm->lut_size;
    return SLB_OK;
}

extern "C" void slb_lightmap_destroy(slb_ctx* ctx, slb_lightmap* lm) {
    if (!lm) return;
    if (ctx) { cudaSetDevice(ctx->device); sync_ctx(ctx); }
    for (void* p : lm->allocs) cudaFree(p);
    delete lm;
}

// ---------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------
extern "C" int slb_result_create(slb_ctx* ctx, int32_t width, int32_t height, int32_t n_frames, uint32_t target_mask,
                                 void* const* external_ptrs, slb_result** out) {
    if (!ctx || !out) return SLB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (width <= 0 || height <= 0 || n_frames <= 0 || (target_mask & ~SLB_TARGETS_ALL) || target_mask == 0)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_create: bad size or target mask");
    CU(cudaSetDevice(ctx->device));
    slb_result* r = new slb_result;
    r->W = width; r->H = height; r->n_frames = n_frames; r->mask = target_mask;
    for (int t = 0; t < SLB_NUM_TARGETS; ++t) { r->ptrs[t] = nullptr; r->owned[t] = false; }
    for (int t = 0; t < SLB_NUM_TARGETS; ++t) {
        if (!(target_mask & (1u << t))) continue;
        if (external_ptrs) {
            if (!external_ptrs[t]) { slb_result_destroy(ctx, r); return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_create: external pointer missing for a requested target"); }
            r->ptrs[t] = external_ptrs[t];
        } else {
            cudaError_t e = cudaMalloc(&r->ptrs[t], (size_t)n_frames * width * height * kTargetBpp[t]);
            if (e != cudaSuccess) { slb_result_destroy(ctx, r); return fail(ctx, SLB_ERR_OUT_OF_MEMORY, std::string("slb_result_create: ") + cudaGetErrorString(e)); }
            r->owned[t] = true;
        }
    }
    if (ctx->keep_hdr) {
        cudaError_t e = cudaMalloc((void**)&r->hdr, (size_t)n_frames * width * height * sizeof(float4));
        if (e != cudaSuccess) { slb_result_destroy(ctx, r); return fail(ctx, SLB_ERR_OUT_OF_MEMORY, "slb_result_create: hdr allocation failed"); }
    }
    *out = r;
    return SLB_OK;
}

extern "C" int slb_result_ptrs(const slb_result* res, void* ptrs[SLB_NUM_TARGETS], size_t bytes_per_pixel[SLB_NUM_TARGETS]) {
    if (!res) return SLB_ERR_INVALID_ARGUMENT;
    for (int t = 0; t < SLB_NUM_TARGETS; ++t) {
        if (ptrs) ptrs[t] = res->ptrs[t];
        if (bytes_per_pixel) bytes_per_pixel[t] = res->ptrs[t] ? kTargetBpp[t] : 0;
    }
    return SLB_OK;
}

extern "C" int slb_result_read(slb_ctx* ctx, const slb_result* res, int target, int32_t first_frame, int32_t n_frames, void* host_out,
                               size_t host_bytes) {
    if (!ctx || !res || !host_out) return SLB_ERR_INVALID_ARGUMENT;
    if (target < 0 || target >= SLB_NUM_TARGETS || !res->ptrs[target]) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_read: target not in the result");
    if (first_frame < 0 || n_frames <= 0 || first_frame + n_frames > res->n_frames) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_read: frame range");
    size_t per = (size_t)res->W * res->H * kTargetBpp[target];
    if (host_bytes < per * n_frames) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_read: host buffer too small");
    CU(cudaSetDevice(ctx->device));
    CU(sync_ctx(ctx));
    CU(cudaMemcpy(host_out, (const uint8_t*)res->ptrs[target] + per * first_frame, per * n_frames, cudaMemcpyDeviceToHost));
    ctx->stats.bytes_d2h += per * n_frames;
    return SLB_OK;
}

extern "C" int slb_result_read_hdr(slb_ctx* ctx, const slb_result* res, int32_t frame, float* host_out, size_t n_floats) {
    if (!ctx || !res || !host_out) return SLB_ERR_INVALID_ARGUMENT;
    if (!res->hdr) return fail(ctx, SLB_ERR_RUNTIME, "slb_result_read_hdr: result was created without SLB_OPT_KEEP_HDR");
    if (frame < 0 || frame >= res->n_frames || n_floats < (size_t)res->W * res->H * 4) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_result_read_hdr: bad frame / size");
    CU(cudaSetDevice(ctx->device));
    CU(sync_ctx(ctx));
    CU(cudaMemcpy(host_out, res->hdr + (size_t)frame * res->W * res->H, (size_t)res->W * res->H * sizeof(float4), cudaMemcpyDeviceToHost));
    return SLB_OK;
}

extern "C" void slb_result_destroy(slb_ctx* ctx, slb_result* res) {
    if (!res) return;
    if (ctx) { cudaSetDevice(ctx->device); sync_ctx(ctx); cudaStreamSynchronize(ctx->copy_stream); }
    for (int t = 0; t < SLB_NUM_TARGETS; ++t) if (res->owned[t] && res->ptrs[t]) cudaFree(res->ptrs[t]);
    if (res->hdr) cudaFree(res->hdr);
    delete res;
}

// ---------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------
namespace {

struct Batch {
    void clear() {   // keeps the capacity of every array: a context marshals thousands of sub-batches
        frames.clear(); draws.clear(); views.clear(); bdraws.clear(); chunk_draw.clear(); world_pre.clear();
        n_chunks = n_shadow_maps = 0; n_tris = 0;
        fused = true; any_ssao = any_auto = any_bg = any_frag_test = false; lean = true;
    }
    std::vector<double> world_pre;        // per (object, sub-mesh) of the current frame: world * pre in double (shared by its views)
    std::vector<DFrame> frames;
    std::vector<DDraw> draws;
    std::vector<DView> views;             // camera views [0, n) then shadow views
    std::vector<DBinDraw> bdraws;         // camera + shadow draws in creation order
    std::vector<uint32_t> chunk_draw;     // setup chunk -> bin draw
    uint32_t n_chunks = 0, n_shadow_maps = 0;
    uint64_t n_tris = 0;
    bool fused = true, any_ssao = false, any_auto = false, any_bg = false, any_frag_test = false;
    bool lean = true;                     // no draw needs the full-featured shade kernel (launch_shade)
};

static inline uint32_t chunks_of(uint32_t n_tris) { return (n_tris + SLB_SETUP_CHUNK - 1) / SLB_SETUP_CHUNK; }

// reference: src/render_pass.cpp:69-129
static void frustum_corners(const slb_scene_desc& sc, hm::Vec3 corners[8]) {
    using namespace hm;
    Mat4 P = load(sc.projection), V = load(sc.world_to_cam);
    Mat4 Pinv = inverted(P);
    float nearv = -1.0f, farv = 1.0f;
    if (sc.n_objects > 0) {
        float nearObj = std::numeric_limits<float>::infinity(), farObj = -nearObj;
        for (int i = 0; i < sc.n_objects; ++i) {
            const slb_object_desc& o = sc.objects[i];
            const slb_mesh* m = o.mesh;
            Mat4 pre = load(o.pretransform);
            Vec3 lo = transform_point(pre, v3(m->bbox_min[0], m->bbox_min[1], m->bbox_min[2]));
            Vec3 hi = transform_point(pre, v3(m->bbox_max[0], m->bbox_max[1], m->bbox_max[2]));
            Vec3 center = (lo + hi) * 0.5f;
            float radius = length(hi - lo) / 2;
            Vec3 objInCam = transform_point(mul(V, load(o.pose)), center);
            Vec3 np = transform_point(P, objInCam - v3(0, 0, radius));
            Vec3 fp = transform_point(P, objInCam + v3(0, 0, radius));
            nearObj = std::min(nearObj, np.z);
            farObj = std::max(farObj, fp.z);
        }
        nearv = std::max(std::max(-1.0f, nearObj), nearv);
        farv = std::min(farObj, farv);
    }
    const float hc[8][3] = {{-1, 1, nearv}, {1, 1, nearv}, {1, -1, nearv}, {-1, -1, nearv},
                            {-1, 1, farv},  {1, 1, farv},  {1, -1, farv},  {-1, -1, farv}};
    Mat4 camToWorld = inverted_rigid(V);
    for (int i = 0; i < 8; ++i) {
        float v[4] = {hc[i][0], hc[i][1], hc[i][2], 1.0f}, a[4], b[4];
        mul4(Pinv, v, a); mul4(camToWorld, a, b);
        corners[i] = v3(b[0] / b[3], b[1] / b[3], b[2] / b[3]);
    }
}
// reference: src/render_pass.cpp:131-211
static hm::Mat4 shadow_matrix(const slb_scene_desc& sc, const hm::Vec3 corners[8], hm::Vec3 lightDirection) {
    using namespace hm;
    Vec3 z = normalize(lightDirection);
    Vec3 x = normalize(cross(z, v3(0, 0, 1)));
    Vec3 y = normalize(cross(z, x));
    Mat4 camToWorld = identity();
    const float xs[3] = {x.x, x.y, x.z}, ys[3] = {y.x, y.y, y.z}, zs[3] = {z.x, z.y, z.z};
    for (int r = 0; r < 3; ++r) { camToWorld.at(r, 0) = xs[r]; camToWorld.at(r, 1) = ys[r]; camToWorld.at(r, 2) = zs[r]; }
    Mat4 worldToCam = inverted_rigid(camToWorld);
    const float inf = std::numeric_limits<float>::infinity();
    Vec3 mn = v3(inf, inf, inf), mx = v3(-inf, -inf, -inf);
    for (int i = 0; i < 8; ++i) { Vec3 c = transform_point(worldToCam, corners[i]); mn = vmin(mn, c); mx = vmax(mx, c); }
    float nearv = mn.z, farv = mx.z;
    float meanZ = (nearv + farv) / 2.0f;
    float spread = farv - meanZ;
    farv = meanZ + 5.0f * spread;
    nearv = meanZ - 5.0f * spread;
    float L = mn.x, R = mx.x, T = mn.y, B = mx.y;
    if (sc.n_objects > 0) {
        Vec3 lo_all = v3(inf, inf, inf), hi_all = v3(-inf, -inf, -inf);
        for (int i = 0; i < sc.n_objects; ++i) {
            const slb_object_desc& o = sc.objects[i];
            const slb_mesh* m = o.mesh;
            Mat4 pre = load(o.pretransform);
            Vec3 lo = transform_point(pre, v3(m->bbox_min[0], m->bbox_min[1], m->bbox_min[2]));
            Vec3 hi = transform_point(pre, v3(m->bbox_max[0], m->bbox_max[1], m->bbox_max[2]));
            float radius = length(hi - lo) / 2;
            Vec3 c = transform_point(mul(worldToCam, load(o.pose)), (lo + hi) * 0.5f);
            lo_all = vmin(lo_all, c - v3(radius, radius, radius));
            hi_all = vmax(hi_all, c + v3(radius, radius, radius));
        }
        L = std::max(L, lo_all.x); R = std::min(R, hi_all.x);
        T = std::max(T, lo_all.y); B = std::min(B, hi_all.y);
    }
    Mat4 Pm; std::memset(Pm.m, 0, sizeof Pm.m);
    Pm.at(0, 0) = 2.0f / (R - L); Pm.at(1, 1) = 2.0f / (B - T); Pm.at(2, 2) = 2.0f / (farv - nearv);
    Pm.at(0, 3) = -(R + L) / (R - L); Pm.at(1, 3) = -(B + T) / (B - T); Pm.at(2, 3) = -(farv + nearv) / (farv - nearv);
    Pm.at(3, 3) = 1.0f;
    return mul(Pm, worldToCam);
}

static void fill_material(DDraw& d, const slb_mesh* mesh, int material, float ovr_metallic, float ovr_roughness) {
    for (int i = 0; i < 5; ++i) d.tex[i] = nullptr;
    bool tex0_alpha = false;
    if (mesh && material >= 0 && material < (int)mesh->materials.size()) {
        const slb_material& m = mesh->materials[material];
        std::memcpy(d.base_color, m.base_color, 16); std::memcpy(d.emissive, m.emissive, 16);
        d.metallic = m.metallic; d.roughness = m.roughness;
        const int32_t tx[5] = {m.tex_base_color, m.tex_normal, m.tex_metallic_roughness, m.tex_emissive, m.tex_occlusion};
        for (int i = 0; i < 5; ++i)
            if (tx[i] >= 0 && tx[i] < (int)mesh->textures.size()) d.tex[i] = mesh->textures[tx[i]]->d;
        if (d.tex[0]) tex0_alpha = mesh->textures[tx[0]]->h.has_alpha != 0;
    } else {   // context default material: 0x3bd267ff_srgbaf (reference: src/context.cpp:382-384)
        d.base_color[0] = 0.04373503f; d.base_color[1] = 0.6444797f; d.base_color[2] = 0.13563333f; d.base_color[3] = 1.0f;
        d.emissive[0] = d.emissive[1] = d.emissive[2] = d.emissive[3] = 0.0f;
        d.metallic = 0.04f; d.roughness = 0.5f;
    }
    if (ovr_metallic >= 0.0f) d.metallic = ovr_metallic;      // render_shader.cpp:372-377
    if (ovr_roughness >= 0.0f) d.roughness = ovr_roughness;
    d.flags = (tex0_alpha || d.base_color[3] < 0.5f) ? DRAW_FRAG_TEST : 0u;
}

}  // namespace

static int validate_scene(slb_ctx* ctx, const slb_scene_desc& sc, const slb_result* result) {
    if (sc.width != result->W || sc.height != result->H) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: scene viewport differs from the result size");
    if (sc.n_objects < 0 || (sc.n_objects > 0 && !sc.objects)) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: objects missing");
    for (int i = 0; i < sc.n_objects; ++i) {
        const slb_object_desc& o = sc.objects[i];
        if (!o.mesh) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: object without mesh");
        if (o.class_index > 0xFFFFu) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "class index out of range (max 65535)");          // mesh.cpp:1083-1089
        if (o.instance_index > 0xFFFFu) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "instance index out of range (max 65535)");    // object.cpp:376-382
        if (o.sticker_texture && o.sticker_texture->h.kind != SLB_TEXTURE_RECT) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "sticker texture must be a RECT texture");
    }
    if (sc.background_image && sc.background_image->h.kind != SLB_TEXTURE_RECT) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "background image must be a RECT texture");
    if (sc.background_plane_texture && sc.background_plane_texture->h.kind != SLB_TEXTURE_2D) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "plane texture must be a 2D texture");
    return SLB_OK;
}

// marshal scenes [s0, s0+n) into a Batch (host arrays only; device pointers are patched in later)
static void build_batch(slb_ctx* ctx, const slb_scene_desc* scenes, int n, const slb_result* result, int first_frame,
                        const slb_result* depth_peel, Batch& b) {
    using namespace hm;
    const int W = result->W, H = result->H;
    const int tiles_x = (W + SLB_TILE - 1) / SLB_TILE, tiles_y = (H + SLB_TILE - 1) / SLB_TILE;
    b.frames.resize(n);
    b.views.resize(n);
    {   // exact sizes up front: no reallocation while the draw lists grow
        size_t n_draws = 0, n_bdraws = 0, n_chunks = 0;
        for (int j = 0; j < n; ++j) {
            size_t subs = 0, chunks = 0;
            for (int i = 0; i < scenes[j].n_objects; ++i)
                for (const slb_submesh& sm : scenes[j].objects[i].mesh->submeshes) { ++subs; chunks += chunks_of(sm.index_count / 3); }
            n_draws += subs + 1; n_bdraws += subs * (1 + SLB_NUM_LIGHTS) + 1; n_chunks += chunks * (1 + SLB_NUM_LIGHTS) + 1;
        }
        b.draws.reserve(n_draws); b.bdraws.reserve(n_bdraws); b.chunk_draw.reserve(n_chunks);
    }
    for (int j = 0; j < n; ++j) {
        const slb_scene_desc& sc = scenes[j];
        DFrame& f = b.frames[j];
        std::memset(&f, 0, sizeof f);
        f.W = W; f.H = H; f.tiles_x = tiles_x; f.tiles_y = tiles_y;
        {
            DView& cv = b.views[j];
            std::memset(&cv, 0, sizeof cv);
            cv.W = W; cv.H = H; cv.tiles_x = tiles_x; cv.tiles_y = tiles_y;
            cv.tile_base = (uint32_t)j * tiles_x * tiles_y; cv.shadow = 0; cv.frame = (uint32_t)j;
        }
        Mat4 P = load(sc.projection), V = load(sc.world_to_cam);
        double Pd[16], Vd[16];
        to_double(P, Pd); to_double(V, Vd);
        b.world_pre.clear();
        std::memcpy(f.P, P.m, 64); std::memcpy(f.V, V.m, 64);
        Mat4 Pinv = inverted(P); std::memcpy(f.Pinv, Pinv.m, 64);
        Mat4 camToWorld = inverted_rigid(V);
        f.camPos[0] = camToWorld.at(0, 3); f.camPos[1] = camToWorld.at(1, 3); f.camPos[2] = camToWorld.at(2, 3);
        f.manual_exposure = sc.manual_exposure;
        f.ssao = sc.ssao_enabled ? 1 : 0;
        const slb_lightmap* lm = sc.light_map;
        f.lm = lm ? lm->d : nullptr;
        bool anyLight = false;
        for (int i = 0; i < SLB_NUM_LIGHTS; ++i) {
            float dir[3] = {0, 0, 0}, col[3] = {0, 0, 0};
            if (lm) {   // render_shader.cpp:270-296
                if (i < lm->h.n_lights) { std::memcpy(dir, lm->h.light_directions[i], 12); std::memcpy(col, lm->h.light_colors[i], 12); }
            } else { std::memcpy(dir, sc.light_directions[i], 12); std::memcpy(col, sc.light_colors[i], 12); }
            std::memcpy(f.lightDir[i], dir, 12); std::memcpy(f.lightCol[i], col, 12);
            bool colZero = col[0] == 0 && col[1] == 0 && col[2] == 0, dirZero = dir[0] == 0 && dir[1] == 0 && dir[2] == 0;
            f.lightActive[i] = !(colZero || dirZero);
            anyLight |= f.lightActive[i] != 0;
        }
        if (!lm) std::memcpy(f.ambient, sc.ambient_light, 12);
        f.peel = depth_peel ? (const float*)depth_peel->ptrs[SLB_TARGET_COORD] + (size_t)(first_frame + j) * W * H * 4 : nullptr;
        f.bg_image = sc.background_image ? sc.background_image->d : nullptr;
        for (int t = 0; t < SLB_NUM_TARGETS; ++t)
            f.out[t] = result->ptrs[t] ? (uint8_t*)result->ptrs[t] + (size_t)(first_frame + j) * W * H * kTargetBpp[t] : nullptr;
        const bool fused = !f.ssao && f.manual_exposure >= 0 && !f.bg_image && !f.lm && !ctx->keep_hdr;
        b.fused &= fused;
        b.any_ssao |= f.ssao != 0; b.any_auto |= f.manual_exposure < 0; b.any_bg |= (f.bg_image || f.lm);

        // ---- draw list in submission order (render_pass.cpp:545-622) ----
        f.draw_begin = (uint32_t)b.draws.size();
        uint32_t prim = 0;
        const bool frag_all = f.peel != nullptr;
        auto push_draw = [&](DDraw& d, uint32_t n_tris) {
            d.n_tris = n_tris; d.prim_base = prim; prim += n_tris;
            d.frame = (uint32_t)j;
            Mat4 o2w = load(d.objectToWorld), m2o = load(d.meshToObject);
            double mw[16];
            world_pre_d(o2w, m2o, mw);                       // kept for this object's shadow views (same operands, same bits)
            b.world_pre.insert(b.world_pre.end(), mw, mw + 16);
            mvp_d(Pd, Vd, mw, d.mvp);
            normal_matrix(mul(o2w, m2o), d.normalToWorld);
            if (frag_all) d.flags |= DRAW_FRAG_TEST;
            auto affine = [](const float* m) { return m[3] == 0.0f && m[7] == 0.0f && m[11] == 0.0f && m[15] == 1.0f; };
            if (affine(d.meshToObject) && affine(d.objectToWorld) && affine(f.V)) d.flags |= DRAW_AFFINE;
            b.lean &= (d.flags & DRAW_AFFINE) && !d.tex[1] && !d.tex[2] && !d.tex[3] && !d.tex[4] && !d.sticker && !f.lm;
            b.any_frag_test |= (d.flags & DRAW_FRAG_TEST) != 0;
            b.n_tris += n_tris;
            DBinDraw bd; std::memset(&bd, 0, sizeof bd);
            bd.pos4 = d.pos4; bd.idx = d.idx; bd.n_tris = n_tris; bd.prim_base = d.prim_base; bd.view = (uint32_t)j;
            bd.draw = (uint32_t)b.draws.size(); bd.flags = d.flags;
            std::memcpy(bd.mvp, d.mvp, 64);
            bd.chunk_base = b.n_chunks; b.n_chunks += chunks_of(n_tris);
            b.chunk_draw.insert(b.chunk_draw.end(), chunks_of(n_tris), (uint32_t)b.bdraws.size());
            b.bdraws.push_back(bd);
            b.draws.push_back(d);
        };
        if (sc.background_plane_size[0] * sc.background_plane_size[0] + sc.background_plane_size[1] * sc.background_plane_size[1] > 0) {
            DDraw d; std::memset(&d, 0, sizeof d);
            const slb_mesh* pm = ctx->plane;
            d.pos4 = pm->pos4; d.attr = pm->attr; d.idx = pm->idx;
            Mat4 scale = identity();
            scale.at(0, 0) = sc.background_plane_size[0] / 2.0f; scale.at(1, 1) = sc.background_plane_size[1] / 2.0f;
            Mat4 o2w = mul(load(sc.background_plane_pose), scale);
            std::memcpy(d.objectToWorld, o2w.m, 64);
            Mat4 I = identity(); std::memcpy(d.meshToObject, I.m, 64); std::memcpy(d.stickerProj, I.m, 64);
            fill_material(d, nullptr, -1, -1.0f, -1.0f);
            if (sc.background_plane_texture) {
                d.base_color[0] = d.base_color[1] = d.base_color[2] = d.base_color[3] = 1.0f;
                d.tex[0] = sc.background_plane_texture->d;
                d.flags = sc.background_plane_texture->h.has_alpha ? DRAW_FRAG_TEST : 0u;
            } else { d.base_color[0] = 0.0f; d.base_color[1] = 0.8f; d.base_color[2] = 0.0f; d.base_color[3] = 1.0f; }
            d.stickerRange[2] = d.stickerRange[3] = 1e-6f;
            push_draw(d, 2);
        }
        for (int i = 0; i < sc.n_objects; ++i) {
            const slb_object_desc& o = sc.objects[i];
            if (!o.visible) continue;
            const slb_mesh* mesh = o.mesh;
            for (const slb_submesh& sm : mesh->submeshes) {
                DDraw d; std::memset(&d, 0, sizeof d);
                d.pos4 = mesh->pos4; d.attr = mesh->attr; d.idx = mesh->idx + sm.index_offset;
                std::memcpy(d.meshToObject, o.pretransform, 64); std::memcpy(d.objectToWorld, o.pose, 64);
                fill_material(d, mesh, sm.material, o.metallic, o.roughness);
                d.class_index = o.class_index; d.instance_index = o.instance_index;
                d.sticker = o.sticker_texture ? o.sticker_texture->d : nullptr;
                std::memcpy(d.stickerProj, o.sticker_projection, 64);
                d.stickerRange[0] = o.sticker_range[0]; d.stickerRange[1] = o.sticker_range[1];
                d.stickerRange[2] = std::max(1e-6f, o.sticker_range[2]); d.stickerRange[3] = std::max(1e-6f, o.sticker_range[3]);
                push_draw(d, sm.index_count / 3);
            }
        }
        f.draw_end = (uint32_t)b.draws.size();
        f.n_prims = prim;
        {   // Encode the draw in the sequence number when it fits 32 bits: seq = local draw << shift | triangle keeps the
            // submission order (all the visibility key needs) and saves the shade kernel its search over prim_base.
            uint32_t max_tris = 1;
            for (uint32_t di = f.draw_begin; di < f.draw_end; ++di) max_tris = std::max(max_tris, b.draws[di].n_tris);
            uint32_t shift = 1;
            while (shift < 32 && (1ull << shift) < max_tris) ++shift;
            const uint64_t n_draws = f.draw_end - f.draw_begin;
            f.seq_shift = 0;
            if (n_draws && shift < 32 && ((n_draws - 1) << shift) + max_tris <= 0xFFFFFFFFull) {
                f.seq_shift = shift;
                const size_t bd0 = b.bdraws.size() - n_draws;   // this frame's camera bin draws were pushed together with its draws
                for (uint32_t li = 0; li < n_draws; ++li) {
                    b.draws[f.draw_begin + li].prim_base = li << shift;
                    b.bdraws[bd0 + li].prim_base = li << shift;
                }
            }
        }

        // ---- shadow views (render_pass.cpp:407-460): one 2048^2 depth-only view per active light ----
        if (anyLight) {
            Vec3 corners[8]; frustum_corners(sc, corners);
            const int stiles = SLB_SHADOW_RES / SLB_TILE;
            for (int li = 0; li < SLB_NUM_LIGHTS; ++li) {
                if (!f.lightActive[li]) continue;
                Mat4 sm = shadow_matrix(sc, corners, v3(f.lightDir[li][0], f.lightDir[li][1], f.lightDir[li][2]));
                std::memcpy(f.shadowMat[li], sm.m, 64);
                const uint32_t slot = b.n_shadow_maps++;
                f.shadowMap[li] = reinterpret_cast<const uint32_t*>((uintptr_t)slot);   // patched to a pointer once the pool is sized
                DView sv; std::memset(&sv, 0, sizeof sv);
                sv.W = sv.H = SLB_SHADOW_RES; sv.tiles_x = sv.tiles_y = stiles;
                sv.tile_base = (uint32_t)n * tiles_x * tiles_y + slot * (uint32_t)(stiles * stiles);
                sv.shadow = 1; sv.frame = (uint32_t)j;
                const uint32_t view_index = (uint32_t)n + slot;
                b.views.push_back(sv);
                uint32_t sprim = 0;
                double smd[16];
                to_double(sm, smd);
                size_t wp = (f.draw_end - f.draw_begin) - [&] {   // first world * pre entry of an object: after the plane's, if drawn
                    size_t subs = 0;
                    for (int i = 0; i < sc.n_objects; ++i) if (sc.objects[i].visible) subs += sc.objects[i].mesh->submeshes.size();
                    return subs; }();
                for (int i = 0; i < sc.n_objects; ++i) {
                    const slb_object_desc& o = sc.objects[i];
                    if (!o.visible) continue;
                    const slb_mesh* mesh = o.mesh;
                    if (!o.casts_shadows) { wp += mesh->submeshes.size(); continue; }
                    for (const slb_submesh& sub : mesh->submeshes) {
                        DBinDraw bd; std::memset(&bd, 0, sizeof bd);
                        bd.pos4 = mesh->pos4; bd.idx = mesh->idx + sub.index_offset; bd.n_tris = sub.index_count / 3;
                        bd.prim_base = sprim; sprim += bd.n_tris; bd.view = view_index;
                        mvp_d(smd, nullptr, b.world_pre.data() + 16 * wp++, bd.mvp);   // == mvp(sm, identity, pose, pretransform)
                        bd.chunk_base = b.n_chunks; b.n_chunks += chunks_of(bd.n_tris);
                        b.chunk_draw.insert(b.chunk_draw.end(), chunks_of(bd.n_tris), (uint32_t)b.bdraws.size());
                        b.bdraws.push_back(bd);
                    }
                }
            }
        }
    }
}

static cudaEvent_t get_event(slb_ctx* ctx) {
    if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
// completed stage timings -> ctx->time_acc, events back to the pool
static void drain_events(slb_ctx* ctx) {
    for (auto& e : ctx->events) {
        cudaEventSynchronize(e.b);
        float ms = 0; cudaEventElapsedTime(&ms, e.a, e.b);
        ctx->time_acc[e.stage] += ms;
        ctx->event_pool.push_back(e.a); ctx->event_pool.push_back(e.b);
    }
    ctx->events.clear();
}
struct StageTimer {
    slb_ctx* ctx; cudaStream_t s; int stage; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(slb_ctx* c, cudaStream_t st, int stage_) : ctx(c), s(st), stage(stage_) {
        if (ctx->time_kernels) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, s); }
    }
    ~StageTimer() {
        if (!ctx->time_kernels) return;
        cudaEventRecord(b, s); ctx->events.push_back({stage, a, b});
        if (ctx->events.size() >= 4096) drain_events(ctx);   // bounded even if nobody ever asks for the statistics
    }
};

static int subbatch_setup_scan(slb_ctx* ctx, Scratch& S, cudaStream_t s);
// Phase 1 of a sub-batch on scratch set `set`: marshal, upload, clear, set-up (with the direct raster paths), scan.
// Ends with the event the second phase waits on before it reads the scan totals.
static int subbatch_phase1(slb_ctx* ctx, int set, const slb_scene_desc* scenes, int n, slb_result* result, int first_frame,
                           const slb_result* depth_peel, cudaStream_t s) {
    Scratch& S = ctx->scr[set];
    if (!S.batch) { S.batch = new Batch; S.batch_delete = [](void* p) { delete static_cast<Batch*>(p); }; }
    Batch& b = *static_cast<Batch*>(S.batch);   // arrays keep their capacity from one sub-batch to the next
    b.clear();
    build_batch(ctx, scenes, n, result, first_frame, depth_peel, b);
    const int W = result->W, H = result->H;
    const size_t npx = (size_t)W * H;
    const uint32_t tiles_per_cam = (uint32_t)b.frames[0].tiles_x * b.frames[0].tiles_y;
    const uint32_t tiles_per_shadow = (uint32_t)(SLB_SHADOW_RES / SLB_TILE) * (SLB_SHADOW_RES / SLB_TILE);
    RasterGrid grid;
    grid.n_cam_tiles = tiles_per_cam * (uint32_t)n; grid.tiles_per_cam = tiles_per_cam; grid.n_cam_views = (uint32_t)n;
    grid.tiles_per_shadow = tiles_per_shadow; grid.n_active = 0;
    const uint32_t n_tiles = grid.n_cam_tiles + tiles_per_shadow * b.n_shadow_maps;

    // ---- device scratch ----
    CU(S.frames_d.reserve(b.frames.size() * sizeof(DFrame)));
    CU(S.draws_d.reserve((b.draws.size() + 1) * sizeof(DDraw)));
    CU(S.views_d.reserve(b.views.size() * sizeof(DView)));
    CU(S.bdraws_d.reserve((b.bdraws.size() + 1) * sizeof(DBinDraw)));
    CU(S.chunk_base_d.reserve((b.chunk_draw.size() + 1) * 4));
    CU(S.tile_count.reserve((size_t)n_tiles * 4, true));
    CU(S.tile_off.reserve(((size_t)n_tiles + 1) * 4));
    CU(S.scan_sums.reserve(((size_t)n_tiles / 4096 + 2) * 8));
    CU(S.active_tiles.reserve((size_t)n_tiles * sizeof(ActiveTile)));
    CU(S.scan_totals.reserve(64));
    CU(S.keys.reserve(npx * n * 8));
    CU(S.clip_recs.reserve((size_t)n * SLB_MAX_CLIP * sizeof(ClipRec)));
    CU(S.clip_counts.reserve((size_t)n * 4));
    CU(S.huge_recs.reserve((size_t)n * SLB_HUGE_PER_VIEW * sizeof(HugeRec)));
    CU(S.huge_counts.reserve((size_t)n * 4));
    CU(S.shadow_maps.reserve((size_t)b.n_shadow_maps * SLB_SHADOW_RES * SLB_SHADOW_RES * 4));
    if (ctx->huge_in_shade && ctx->huge_prepare) CU(S.huge_shade.reserve((size_t)n * SLB_HUGE_PER_VIEW * sizeof(HugeShade)));
    if (ctx->shadow_mask) CU(S.shadow_mask.reserve(((size_t)b.n_shadow_maps + 1) * SLB_SHADOW_MASK_WORDS * 4));
    const bool post = !b.fused;
    if (post) {
        if (!result->hdr) CU(S.hdr.reserve(npx * n * 16));
        CU(S.ao.reserve(npx * n * 4));
        CU(S.avg.reserve((size_t)n * 16));
        if (!result->ptrs[SLB_TARGET_NORMAL] && b.any_ssao) CU(S.scratch_normal.reserve(npx * n * 16));
        if (!result->ptrs[SLB_TARGET_CAM_COORD] && b.any_ssao) CU(S.scratch_cam.reserve(npx * n * 16));
        if (b.any_ssao) CU(S.zplane.reserve(npx * n * 4));
        if (b.any_auto) { CU(S.mip_a.reserve((npx / 4 + W + H + 4) * n * 16)); CU(S.mip_b.reserve((npx / 16 + W + H + 4) * n * 16)); }
    }
    // ---- patch device pointers into the host arrays ----
    uint32_t* smaps = S.shadow_maps.as<uint32_t>();
    const size_t smap_elems = (size_t)SLB_SHADOW_RES * SLB_SHADOW_RES;
    for (int j = 0; j < n; ++j) {
        DFrame& f = b.frames[j];
        f.keys = S.keys.as<uint64_t>() + npx * j;
        b.views[j].out = f.keys;
        f.clip = S.clip_recs.as<ClipRec>() + (size_t)j * SLB_MAX_CLIP;
        f.clip_count = S.clip_counts.as<uint32_t>() + j;
        if (ctx->huge_in_shade) {
            f.huge = S.huge_recs.as<HugeRec>() + (size_t)j * SLB_HUGE_PER_VIEW; f.huge_n = S.huge_counts.as<uint32_t>() + j;
            b.views[j].huge = S.huge_recs.as<HugeRec>() + (size_t)j * SLB_HUGE_PER_VIEW; b.views[j].huge_n = S.huge_counts.as<uint32_t>() + j;
            if (ctx->huge_prepare) f.huge_shade = S.huge_shade.as<HugeShade>() + (size_t)j * SLB_HUGE_PER_VIEW;
        }
        f.fused_tonemap = b.fused ? 1 : 0;
        for (int li = 0; li < SLB_NUM_LIGHTS; ++li) {
            const uintptr_t slot = (uintptr_t)f.shadowMap[li];
            f.shadowMask[li] = (f.lightActive[li] && ctx->shadow_mask) ? S.shadow_mask.as<uint32_t>() + (size_t)SLB_SHADOW_MASK_WORDS * slot : nullptr;
            f.shadowMap[li] = f.lightActive[li] ? smaps + smap_elems * slot : nullptr;
        }
        f.scratch_normal = (float4*)f.out[SLB_TARGET_NORMAL];
        f.scratch_cam = (float4*)f.out[SLB_TARGET_CAM_COORD];
        if (post) {
            f.hdr = result->hdr ? result->hdr + npx * (first_frame + j) : S.hdr.as<float4>() + npx * j;
            f.ao = S.ao.as<float>() + npx * j;
            f.avg = S.avg.as<float>() + 4 * j;
            if (!f.scratch_normal && b.any_ssao) f.scratch_normal = S.scratch_normal.as<float4>() + npx * j;
            if (!f.scratch_cam && b.any_ssao) f.scratch_cam = S.scratch_cam.as<float4>() + npx * j;
            if (b.any_ssao) f.zplane = S.zplane.as<float>() + npx * j;
        }
    }
    // generation tag of the shadow maps of this sub-batch: a real clear only when the pool moved / grew or the 8-bit
    // generation wraps
    const bool pool_changed = S.shadow_maps.p != S.shadow_pool_at || S.shadow_maps.cap != S.shadow_pool_cap;
    bool clear_shadow_pool = false;
    if (b.n_shadow_maps) {
        if (pool_changed || S.shadow_gen >= 255) { clear_shadow_pool = true; S.shadow_gen = 0; }
        S.shadow_pool_at = S.shadow_maps.p; S.shadow_pool_cap = S.shadow_maps.cap;
        ++S.shadow_gen;
    }
    const uint32_t shadow_tagbits = (255u - S.shadow_gen) << 24;
    for (int j = 0; j < n; ++j) b.frames[j].shadow_tagbits = shadow_tagbits;
    for (uint32_t sidx = 0; sidx < b.n_shadow_maps; ++sidx) {
        b.views[n + sidx].out = smaps + smap_elems * sidx; b.views[n + sidx].tagbits = shadow_tagbits;
        b.views[n + sidx].mask = ctx->shadow_mask ? S.shadow_mask.as<uint32_t>() + (size_t)SLB_SHADOW_MASK_WORDS * sidx : nullptr;
    }

    // ---- one staged upload ----
    const size_t sz[5] = {b.frames.size() * sizeof(DFrame), b.draws.size() * sizeof(DDraw), b.views.size() * sizeof(DView),
                          b.bdraws.size() * sizeof(DBinDraw), b.chunk_draw.size() * 4};
    const void* src[5] = {b.frames.data(), b.draws.data(), b.views.data(), b.bdraws.data(), b.chunk_draw.data()};
    void* dst[5] = {S.frames_d.p, S.draws_d.p, S.views_d.p, S.bdraws_d.p, S.chunk_base_d.p};
    size_t total_sz = 0;
    for (int i = 0; i < 5; ++i) total_sz += (sz[i] + 63) & ~(size_t)63;
    int rc = ensure_staging(ctx, S, total_sz + 64);
    if (rc != SLB_OK) return rc;
    {
        uint8_t* st = (uint8_t*)S.staging;
        size_t at = 0;
        for (int i = 0; i < 5; ++i) {
            if (sz[i]) {
                std::memcpy(st + at, src[i], sz[i]);
                CU(cudaMemcpyAsync(dst[i], st + at, sz[i], cudaMemcpyHostToDevice, s));
            }
            at += (sz[i] + 63) & ~(size_t)63;
            ctx->stats.bytes_h2d += sz[i];
        }
    }
    CU(cudaMemsetAsync(S.clip_counts.p, 0, (size_t)n * 4, s));
    CU(cudaMemsetAsync(S.huge_counts.p, 0, (size_t)n * 4, s));
    CU(cudaMemsetAsync(S.tile_count.p, 0, (size_t)n_tiles * 4, s));
    {   // "nothing drawn": keys = all ones, shadow d24 >= 0xFFFFFF; only non-empty tiles get a raster warp
        StageTimer t(ctx, s, ST_SHADOW);
        CU(cudaMemsetAsync(S.keys.p, 0xFF, npx * n * 8, s));
        if (clear_shadow_pool) CU(cudaMemsetAsync(S.shadow_maps.p, 0xFF, S.shadow_maps.cap, s));   // whole pool: every slot starts stale
        if (ctx->shadow_mask && b.n_shadow_maps) CU(cudaMemsetAsync(S.shadow_mask.p, 0, (size_t)b.n_shadow_maps * SLB_SHADOW_MASK_WORDS * 4, s));
    }
    // ---- setup (camera + shadow views together) -> scan ----
    Pending& P = S.pending;
    P.active = true; P.n = n; P.W = W; P.H = H; P.first_frame = first_frame; P.npx = npx; P.grid = grid; P.n_tiles = n_tiles;
    P.result = result; P.post = post;
    P.total_bin_tris = 0;
    for (const DBinDraw& bd : b.bdraws) P.total_bin_tris += bd.n_tris;
    return subbatch_setup_scan(ctx, S, s);
}

// set-up + scan of the sub-batch pending on S (also the retry after a survivor-buffer overflow)
static int subbatch_setup_scan(slb_ctx* ctx, Scratch& S, cudaStream_t s) {
    Pending& P = S.pending;
    Batch& b = *static_cast<Batch*>(S.batch);
    // ordinary survivors in [0, normal_cap), huge ones (whole block each in the emit pass) in the last fifth
    const size_t want = (size_t)P.total_bin_tris + P.total_bin_tris / 4 + 8192;
    if (S.survivors.cap < want * sizeof(PairRec)) CU(S.survivors.reserve(want * sizeof(PairRec)));
    const uint32_t surv_capacity = (uint32_t)std::min<size_t>(S.survivors.cap / sizeof(PairRec), 0xFFFFFFF0u);
    P.huge_cap = surv_capacity / 5;
    P.normal_cap = surv_capacity - P.huge_cap;
    CU(cudaMemsetAsync(S.scan_totals.p, 0, 64, s));   // [0] survivor count, [1] overflow flag, [2] huge survivor count
    {
        StageTimer t(ctx, s, ST_COUNT);
        launch_setup(S.views_d.as<DView>(), S.frames_d.as<DFrame>(), S.bdraws_d.as<DBinDraw>(), S.chunk_base_d.as<uint32_t>(), b.n_chunks,
                     S.tile_count.as<uint32_t>(), S.survivors.as<PairRec>(), S.scan_totals.as<uint32_t>(), P.normal_cap, P.huge_cap,
                     ctx->direct_max, ctx->warp_max, s);
    }
    {
        StageTimer t(ctx, s, ST_SCAN);
        launch_scan(S.tile_count.as<uint32_t>(), S.tile_off.as<uint32_t>(), S.active_tiles.as<ActiveTile>(),
                    S.scan_sums.as<unsigned long long>(), S.total_pinned, S.scan_totals.as<uint32_t>(), P.n_tiles, s);
    }
    // The scan writes its totals straight into page-locked host memory (UVA): a D2H memcpy here would queue
    // behind the result copies of the previous sub-batch on the copy engine and serialise the pipeline.
    CU(cudaEventRecord(S.scanned, s));
    return SLB_OK;
}

// Phase 2: wait for the scan totals (the pair buffer, the emit grid and the raster grid are sized from them), then
// emit -> raster -> shade -> post passes.
static int subbatch_phase2(slb_ctx* ctx, int set, cudaStream_t s) {
    Scratch& S = ctx->scr[set];
    Pending& P = S.pending;
    if (!P.active) return SLB_OK;
    P.active = false;
    Batch& b = *static_cast<Batch*>(S.batch);
    const int n = P.n, W = P.W, H = P.H, first_frame = P.first_frame;
    const size_t npx = P.npx;
    slb_result* result = P.result;
    const bool post = P.post;
    RasterGrid grid = P.grid;
    const DFrame* frames_d = S.frames_d.as<DFrame>();
    const DDraw* draws_d = S.draws_d.as<DDraw>();
    const DView* views_d = S.views_d.as<DView>();
    for (int attempt = 0;; ++attempt) {
        CU(cudaEventSynchronize(S.scanned));
        if (S.total_pinned[3] == 0) break;
        // more clipped sub-triangles than triangles + 8192 (pathological): grow the survivor buffer and redo the pass
        if (attempt == 4) return fail(ctx, SLB_ERR_RUNTIME, "slb_render_batch: survivor buffer overflow");
        P.total_bin_tris = P.total_bin_tris * 2 + 65536;
        CU(cudaMemsetAsync(S.tile_count.p, 0, (size_t)P.n_tiles * 4, s));
        CU(cudaMemsetAsync(S.clip_counts.p, 0, (size_t)n * 4, s));
        CU(cudaMemsetAsync(S.huge_counts.p, 0, (size_t)n * 4, s));
        int rc = subbatch_setup_scan(ctx, S, s);
        if (rc != SLB_OK) return rc;
    }
    const uint32_t normal_cap = P.normal_cap;
    uint32_t total_pairs = 0, n_survivors = 0, n_huge = 0;
    total_pairs = S.total_pinned[0];
    grid.n_active = S.total_pinned[1];
    n_survivors = S.total_pinned[2];
    n_huge = S.total_pinned[4];
    CU(S.pairs.reserve(((size_t)total_pairs + 1) * sizeof(PairRec)));
    {
        StageTimer t(ctx, s, ST_EMIT);
        launch_emit(views_d, S.survivors.as<PairRec>(), n_survivors, S.survivors.as<PairRec>() + normal_cap, n_huge,
                    S.tile_count.as<uint32_t>(), S.pairs.as<PairRec>(), total_pairs, s);
    }
    {
        StageTimer t(ctx, s, ST_RASTER);
        launch_raster(b.any_frag_test, views_d, frames_d, draws_d, S.active_tiles.as<ActiveTile>(), S.pairs.as<PairRec>(), grid, s);
    }
    {
        StageTimer t(ctx, s, ST_SHADE);
        launch_shade(frames_d, draws_d, n, W, H, b.lean && ctx->lean_shade, s);
    }
    ctx->stats.kernel_launches += (b.n_chunks ? 1 : 0) + 3 + (n_survivors ? 1 : 0) + (n_huge ? 1 : 0) + (grid.n_active ? 1 : 0) + 1;
    if (post) {
        if (b.any_auto) {   // 1x1 level of the mip chain of the HDR target, taken before background / SSAO
            StageTimer t(ctx, s, ST_POST);
            const float4* src = result->hdr ? result->hdr + npx * first_frame : S.hdr.as<float4>();
            size_t src_stride = npx;
            int sw = W, sh = H; bool flip = false;
            while (sw > 1 || sh > 1) {
                int dw = sw > 1 ? sw >> 1 : 1, dh = sh > 1 ? sh >> 1 : 1;
                const bool last = dw == 1 && dh == 1;
                float4* dst = last ? S.avg.as<float4>() : (flip ? S.mip_b.as<float4>() : S.mip_a.as<float4>());
                size_t dst_stride = last ? 1 : (size_t)dw * dh;
                launch_downsample(src, sw, sh, dst, n, src_stride, dst_stride, s);
                ctx->stats.kernel_launches += 1;
                src = dst; src_stride = dst_stride; sw = dw; sh = dh; flip = !flip;
            }
            if (W == 1 && H == 1) CU(cudaMemcpyAsync(S.avg.p, src, (size_t)n * 16, cudaMemcpyDeviceToDevice, s));
        }
        if (b.any_bg) {
            StageTimer t(ctx, s, ST_POST);
            launch_background(frames_d, n, W, H, s);
            ctx->stats.kernel_launches += 1;
        }
        if (b.any_ssao) {
            StageTimer t(ctx, s, ST_SSAO);
            launch_ssao(frames_d, n, W, H, s);
            ctx->stats.kernel_launches += 1;
        }
        {
            StageTimer t(ctx, s, ST_POST);
            launch_ssao_apply_tonemap(frames_d, n, W, H, s);
            ctx->stats.kernel_launches += 1;
        }
    }
    CU(cudaGetLastError());
    ctx->stats.frames_rendered += n;
    ctx->stats.triangles_submitted += b.n_tris;
    ctx->stats.triangles_binned = total_pairs;
    return SLB_OK;
}

// the two phases back to back (one sub-batch at a time: slb_render_batch_host)
static int render_subbatch(slb_ctx* ctx, const slb_scene_desc* scenes, int n, slb_result* result, int first_frame,
                           const slb_result* depth_peel, cudaStream_t s) {
    int rc = subbatch_phase1(ctx, 0, scenes, n, result, first_frame, depth_peel, s);
    if (rc != SLB_OK) return rc;
    return subbatch_phase2(ctx, 0, s);
}

// Stage times are recorded as events on the stream and only read here, on demand: no synchronisation inside the
// render calls, so timing a run does not drain the sub-batch pipeline between calls.
static void collect_times(slb_ctx* ctx) {
    drain_events(ctx);
    for (int i = 0; i < 8; ++i) { ctx->stats.last_kernel_ms[i] = i < ST_N ? ctx->time_acc[i] : 0.0f; ctx->time_acc[i] = 0.0f; }
}

extern "C" int slb_render_batch(slb_ctx* ctx, const slb_scene_desc* scenes, int32_t n_scenes, slb_result* result, int32_t first_frame,
                                const slb_result* depth_peel, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!scenes || n_scenes <= 0 || !result) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: scenes / result missing");
    if (first_frame < 0 || first_frame + n_scenes > result->n_frames) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: frame range exceeds the result");
    if (depth_peel && (depth_peel->W != result->W || depth_peel->H != result->H || depth_peel->n_frames < first_frame + n_scenes ||
                       !depth_peel->ptrs[SLB_TARGET_COORD]))
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch: depth-peel result must hold the coord target at the same size");
    if (ctx->keep_hdr && !result->hdr) return fail(ctx, SLB_ERR_RUNTIME, "slb_render_batch: SLB_OPT_KEEP_HDR is on but the result was created without it");
    for (int i = 0; i < n_scenes; ++i) { int rc = validate_scene(ctx, scenes[i], result); if (rc != SLB_OK) return rc; }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = enter_stream(ctx, stream);
    // Software pipeline over the two scratch sets: the first phase of sub-batch k+1 (marshal, upload, clear, set-up,
    // scan) is queued BEFORE the host waits for the scan totals of sub-batch k, so the GPU always has work while the
    // host sizes and queues the second phase (emit, raster, shade). Stream order: P1(0) P1(1) P2(0) P1(2) P2(1) ...
    // SLB_OPT_OVERLAP (default on, more than one sub-batch): the first phases are queued on the context's auxiliary stream, the second
    // phases on `s`, so that the set-up of sub-batch k+1 can fill the tail and the launch gaps of the shade pass of sub-batch k (the two
    // touch different scratch sets). Ordering: the auxiliary stream starts after everything `s` holds at the call; P2(k) waits for P1(k)
    // (p1_done); P1(k+2), which reuses the scratch set of k, waits for P2(k) (p2_done, also across calls). Every P1 is followed by its P2
    // on `s`, so `s` alone still orders the call for the caller, for enter_stream() and for sync_ctx().
    const bool two_streams = ctx->overlap && n_scenes > ctx->max_subbatch && ctx->aux_stream;
    cudaStream_t s1 = two_streams ? ctx->aux_stream : s;
    if (two_streams) { CU(cudaEventRecord(ctx->call_start, s)); CU(cudaStreamWaitEvent(s1, ctx->call_start, 0)); }
    auto phase1 = [&](int k, int at, int n) -> int {
        Scratch& S = ctx->scr[k & 1];
        if (two_streams && S.p2_recorded) { if (cudaStreamWaitEvent(s1, S.p2_done, 0) != cudaSuccess) return fail(ctx, SLB_ERR_CUDA, "cudaStreamWaitEvent"); }
        int r = subbatch_phase1(ctx, k & 1, scenes + at, n, result, first_frame + at, depth_peel, s1);
        if (r == SLB_OK && two_streams && cudaEventRecord(S.p1_done, s1) != cudaSuccess) r = fail(ctx, SLB_ERR_CUDA, "cudaEventRecord");
        return r;
    };
    auto phase2 = [&](int k) -> int {
        Scratch& S = ctx->scr[k & 1];
        if (two_streams && cudaStreamWaitEvent(s, S.p1_done, 0) != cudaSuccess) return fail(ctx, SLB_ERR_CUDA, "cudaStreamWaitEvent");
        int r = subbatch_phase2(ctx, k & 1, s);
        if (r == SLB_OK) { if (cudaEventRecord(S.p2_done, s) != cudaSuccess) r = fail(ctx, SLB_ERR_CUDA, "cudaEventRecord"); else S.p2_recorded = true; }
        return r;
    };
    int k = 0, rc = SLB_OK;
    for (int at = 0; at < n_scenes && rc == SLB_OK; at += ctx->max_subbatch, ++k) {
        const int n = std::min(ctx->max_subbatch, n_scenes - at);
        rc = phase1(k, at, n);
        if (rc == SLB_OK && k > 0) rc = phase2(k - 1);
    }
    if (rc == SLB_OK && k > 0) rc = phase2(k - 1);
    if (rc != SLB_OK) {   // leave no half-issued sub-batch behind
        cudaStreamSynchronize(s1);
        cudaStreamSynchronize(s);
        ctx->scr[0].pending.active = ctx->scr[1].pending.active = false;
        return rc;
    }
    return SLB_OK;
}

extern "C" int slb_render_batch_host(slb_ctx* ctx, const slb_scene_desc* scenes, int32_t n_scenes, uint32_t target_mask,
                                     void* const host_ptrs[SLB_NUM_TARGETS]) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!scenes || n_scenes <= 0 || !host_ptrs || target_mask == 0 || (target_mask & ~SLB_TARGETS_ALL))
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch_host: bad arguments");
    for (int t = 0; t < SLB_NUM_TARGETS; ++t)
        if ((target_mask & (1u << t)) && !host_ptrs[t]) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_render_batch_host: host pointer missing for a requested target");
    CU(cudaSetDevice(ctx->device));
    const int W = scenes[0].width, H = scenes[0].height;
    for (int i = 0; i < 2; ++i) {
        slb_result* r = ctx->slot[i];
        if (r && (r->W != W || r->H != H || r->mask != target_mask || r->n_frames != ctx->max_subbatch || (ctx->keep_hdr && !r->hdr))) {
            slb_result_destroy(ctx, r); ctx->slot[i] = nullptr;
        }
        if (!ctx->slot[i]) { int rc = slb_result_create(ctx, W, H, ctx->max_subbatch, target_mask, nullptr, &ctx->slot[i]); if (rc != SLB_OK) return rc; }
        if (!ctx->slot_rendered[i]) { CU(cudaEventCreateWithFlags(&ctx->slot_rendered[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&ctx->slot_copied[i], cudaEventDisableTiming)); }
    }
    for (int i = 0; i < n_scenes; ++i) { int rc = validate_scene(ctx, scenes[i], ctx->slot[0]); if (rc != SLB_OK) return rc; }
    const size_t npx = (size_t)W * H;
    int k = 0;
    bool used[2] = {false, false};
    const bool dbg = getenv("SLB_DEBUG_TIMING") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now();
    for (int at = 0, n = 0; at < n_scenes; at += n, ++k) {
        // the first sub-batch is short so that the device->host copies (the bottleneck of this entry point) start early
        n = std::min(k == 0 ? std::min(ctx->max_subbatch, 16) : ctx->max_subbatch, n_scenes - at);
        const int sl = k & 1;
        if (used[sl]) CU(cudaStreamWaitEvent(ctx->stream, ctx->slot_copied[sl], 0));   // slot free again?
        const double t0 = now();
        int rc = render_subbatch(ctx, scenes + at, n, ctx->slot[sl], 0, nullptr, ctx->stream);
        if (rc != SLB_OK) return rc;
        const double t1 = now();
        CU(cudaEventRecord(ctx->slot_rendered[sl], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->slot_rendered[sl], 0));
        for (int t = 0; t < SLB_NUM_TARGETS; ++t) {
            if (!(target_mask & (1u << t))) continue;
            const size_t per = npx * kTargetBpp[t];
            CU(cudaMemcpyAsync((uint8_t*)host_ptrs[t] + per * at, ctx->slot[sl]->ptrs[t], per * n, cudaMemcpyDeviceToHost, ctx->copy_stream));
            ctx->stats.bytes_d2h += per * n;
        }
        CU(cudaEventRecord(ctx->slot_copied[sl], ctx->copy_stream));
        used[sl] = true;
        if (dbg) fprintf(stderr, "[slb] sub-batch %d: render_subbatch host %.2f ms (from %.2f), copies queued at %.2f\n", k, t1 - t0, t0 - t_begin, now() - t_begin);
    }
    if (dbg) fprintf(stderr, "[slb] all queued at %.2f ms\n", now() - t_begin);
    CU(sync_ctx(ctx));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    if (dbg) fprintf(stderr, "[slb] done at %.2f ms\n", now() - t_begin);
    return SLB_OK;
}

// ---------------------------------------------------------------------------------------------
// config 4 helpers
// ---------------------------------------------------------------------------------------------
extern "C" int slb_diff_sobel_valid_mask(slb_ctx* ctx, const int16_t* instance_index, const float* depth, uint8_t* valid_out,
                                         int32_t height, int32_t width, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!instance_index || !depth || !valid_out || height <= 0 || width <= 0) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_diff_sobel_valid_mask: bad arguments");
    CU(cudaSetDevice(ctx->device));
    launch_sobel_valid_mask(instance_index, depth, valid_out, height, width, enter_stream(ctx, stream));
    ctx->stats.kernel_launches += 1;
    CU(cudaGetLastError());
    return SLB_OK;
}
extern "C" int slb_diff_dilate_object_mask(slb_ctx* ctx, const uint8_t* mask, const uint8_t* valid, const float* coords,
                                           int32_t coord_stride, uint8_t* mask_out, float* coords_out, int32_t height, int32_t width,
                                           void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!mask || !valid || !coords || !mask_out || !coords_out || coord_stride < 3 || height <= 0 || width <= 0)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_diff_dilate_object_mask: bad arguments");
    CU(cudaSetDevice(ctx->device));
    launch_dilate_object_mask(mask, valid, coords, coord_stride, mask_out, coords_out, height, width, enter_stream(ctx, stream));
    ctx->stats.kernel_launches += 1;
    CU(cudaGetLastError());
    return SLB_OK;
}

extern "C" int slb_diff_pose_grad(slb_ctx* ctx, const uint8_t* rgb, const int16_t* instance_index, const float* coord_depth,
                                  const float* grad_image, const float* projection, const float* poses, const int32_t* instance_ids,
                                  int32_t n_objects, float* grad_out, int32_t height, int32_t width, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!rgb || !instance_index || !coord_depth || !grad_image || !projection || !grad_out || n_objects < 0 || height <= 0 || width <= 0 ||
        (n_objects > 0 && (!poses || !instance_ids)))
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_diff_pose_grad: bad arguments");
    if (n_objects == 0) return SLB_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = enter_stream(ctx, stream);
    // parameter block: P (row-major), then per object T0 (row-major) + instance id
    const size_t n_par = 16 + (size_t)17 * n_objects;
    std::vector<float> par(n_par);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) par[r * 4 + c] = projection[c * 4 + r];
    for (int o = 0; o < n_objects; ++o) {
        float* T = par.data() + 16 + (size_t)17 * o;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T[r * 4 + c] = poses[(size_t)o * 16 + c * 4 + r];
        const int32_t id = (int32_t)(int16_t)instance_ids[o];   // the instance target is read as int16 (py_render_pass.cpp:103-127)
        std::memcpy(T + 16, &id, 4);
    }
    CU(ctx->diff_params.reserve(n_par * sizeof(float)));
    CU(ctx->diff_partial.reserve(pose_grad_partial_floats(n_objects, height, width) * sizeof(float)));
    // pageable source: the copy is staged by the runtime before the call returns, so `par` may go out of scope
    CU(cudaMemcpyAsync(ctx->diff_params.p, par.data(), n_par * sizeof(float), cudaMemcpyHostToDevice, s));
    launch_pose_grad(rgb, instance_index, coord_depth, grad_image, ctx->diff_params.as<float>(), n_objects, ctx->diff_partial.as<float>(),
                     grad_out, height, width, s);
    ctx->stats.kernel_launches += 2;
    CU(cudaGetLastError());
    return SLB_OK;
}

// ---------------------------------------------------------------------------------------------
// camera noise model
// ---------------------------------------------------------------------------------------------
extern "C" int slb_camera_model(slb_ctx* ctx, const void* in, int32_t in_format, float* out, int32_t n_images, int32_t height,
                                int32_t width, const slb_camera_params* params, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!in || !out || !params || n_images <= 0 || height <= 0 || width <= 0 || (in_format != 0 && in_format != 1) || (const void*)out == in)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_camera_model: bad arguments");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = enter_stream(ctx, stream);
    const size_t img_floats = (size_t)3 * height * width;
    bool any_post = false;
    for (int i = 0; i < n_images; ++i) any_post |= (params[i].stages & SLB_CAM_POST_BLUR) != 0;
    for (int i = 0; i < n_images; ++i)
        if (((params[i].stages & SLB_CAM_POST_BLUR) != 0) != any_post)
            return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_camera_model: SLB_CAM_POST_BLUR must be set for all images of a call or for none");
    CU(ctx->cam_params.reserve((size_t)n_images * sizeof(slb_camera_params)));
    CU(cudaMemcpyAsync(ctx->cam_params.p, params, (size_t)n_images * sizeof(slb_camera_params), cudaMemcpyHostToDevice, s));
    float* mid = out;
    if (any_post) { CU(ctx->cam_mid.reserve((size_t)n_images * img_floats * sizeof(float))); mid = ctx->cam_mid.as<float>(); }
    launch_camera_stage1(in_format == 0 ? (const float*)in : nullptr, in_format == 1 ? (const uint8_t*)in : nullptr, mid,
                         ctx->cam_params.as<slb_camera_params>(), n_images, height, width, s);
    if (any_post) launch_camera_stage2(mid, out, 0.4f, n_images, height, width, s);
    ctx->stats.kernel_launches += any_post ? 2 : 1;
    CU(cudaGetLastError());
    return SLB_OK;
}

// ---------------------------------------------------------------------------------------------
// batched PNG encoder
// ---------------------------------------------------------------------------------------------
extern "C" size_t slb_png_bound(int32_t height, int32_t width, int32_t channels, int32_t bytes_per_channel) {
    if (height <= 0 || width <= 0 || channels <= 0 || bytes_per_channel <= 0) return 0;
    return png_file_bound(height, width, channels, bytes_per_channel);
}
extern "C" int slb_png_encode(slb_ctx* ctx, const void* images, int32_t n_images, int32_t height, int32_t width, int32_t channels,
                              int32_t bytes_per_channel, uint8_t* out, size_t out_stride, uint32_t* sizes, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!images || !out || !sizes || n_images <= 0 || height <= 0 || width <= 0)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_png_encode: bad arguments");
    const bool ok8 = bytes_per_channel == 1 && (channels == 1 || channels == 3 || channels == 4);
    const bool ok16 = bytes_per_channel == 2 && channels == 1;
    if (!ok8 && !ok16)   // py_image_saver.cpp:50-95
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_png_encode: images must be uint8 HxW, HxWx3, HxWx4 or 16-bit HxW");
    if (out_stride < png_file_bound(height, width, channels, bytes_per_channel))
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_png_encode: out_stride is smaller than slb_png_bound()");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = enter_stream(ctx, stream);
    if (!ctx->png_tables) { png_upload_tables(); ctx->png_tables = true; }
    const size_t rows = (size_t)n_images * height * png_segments(width, channels, bytes_per_channel);   // deflate blocks
    CU(ctx->png_rows.reserve(rows * png_seg_bound()));
    CU(ctx->png_info.reserve(rows * png_row_info_bytes()));
    CU(ctx->png_offsets.reserve(rows * 4));
    launch_png_encode((const uint8_t*)images, n_images, height, width, channels, bytes_per_channel, ctx->png_rows.as<uint8_t>(), ctx->png_info.p,
                      ctx->png_offsets.as<uint32_t>(), out, out_stride, sizes, s);
    ctx->stats.kernel_launches += 3;
    CU(cudaGetLastError());
    return SLB_OK;
}

// ---- batched JPEG encoder ---------------------------------------------------------------------
extern "C" size_t slb_jpeg_bound(int32_t height, int32_t width, int32_t channels) {
    if (height <= 0 || width <= 0 || (channels != 1 && channels != 3 && channels != 4)) return 0;
    return jpeg_file_bound(height, width, channels);
}
extern "C" int slb_jpeg_encode(slb_ctx* ctx, const void* images, int32_t n_images, int32_t height, int32_t width, int32_t channels,
                               int32_t quality, uint8_t* out, size_t out_stride, uint32_t* sizes, void* stream) {
    if (!ctx) return SLB_ERR_INVALID_ARGUMENT;
    if (!images || !out || !sizes || n_images <= 0 || height <= 0 || width <= 0 || out_stride == 0)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_jpeg_encode: bad arguments");
    if (channels != 1 && channels != 3 && channels != 4)   // JpegImageConverter: R8Unorm / RGB8Unorm (RGBA8Unorm with alpha dropped)
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_jpeg_encode: images must be uint8 HxW, HxWx3 or HxWx4");
    if (height > 65535 || width > 65535) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_jpeg_encode: JPEG dimensions are limited to 65535");
    if (quality < 1 || quality > 100) return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_jpeg_encode: quality must be in 1..100");
    if (jpeg_blocks(height, width, channels) > 2500000u)   // bit offsets are 32-bit: 1658 bits per block at most
        return fail(ctx, SLB_ERR_INVALID_ARGUMENT, "slb_jpeg_encode: image too large (more than 2.5 M blocks, about 100 megapixels in colour)");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = enter_stream(ctx, stream);
    const int key[4] = {height, width, channels, quality};
    if (memcmp(key, ctx->jpeg_key, sizeof(key)) != 0) {   // quantisation / Huffman tables in constant memory + the file header of this shape
        const std::vector<uint8_t> header = jpeg_prepare(height, width, channels, quality, &ctx->jpeg_quant, s);
        CU(ctx->jpeg_header.reserve(header.size()));
        CU(cudaMemcpyAsync(ctx->jpeg_header.p, header.data(), header.size(), cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));   // `header` is a local
        ctx->jpeg_header_len = (int)header.size();
        memcpy(ctx->jpeg_key, key, sizeof(key));
    }
    // scratch per image: coefficients (2 B each) + worst-case bit stream; batches are cut so the scratch stays below 512 MB
    const size_t per_image = jpeg_coef_bytes(height, width, channels) + jpeg_stream_words(height, width, channels) * 4;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_images, ((size_t)512 << 20) / per_image));
    CU(ctx->jpeg_coefs.reserve((size_t)chunk * jpeg_coef_bytes(height, width, channels)));
    CU(ctx->jpeg_stream.reserve((size_t)chunk * jpeg_stream_words(height, width, channels) * 4));
    CU(ctx->jpeg_off.reserve((size_t)chunk * jpeg_blocks(height, width, channels) * 4));
    CU(ctx->jpeg_total.reserve((size_t)chunk * 4));
    const size_t image_bytes = (size_t)height * width * channels;
    for (int at = 0; at < n_images; at += chunk) {
        const int n = std::min(chunk, n_images - at);
        launch_jpeg_encode((const uint8_t*)images + (size_t)at * image_bytes, n, height, width, channels, ctx->jpeg_quant, ctx->jpeg_coefs.as<int16_t>(),
                           ctx->jpeg_off.as<uint32_t>(), ctx->jpeg_total.as<uint32_t>(), ctx->jpeg_stream.as<uint32_t>(),
                           ctx->jpeg_header.as<uint8_t>(), ctx->jpeg_header_len, out + (size_t)at * out_stride, out_stride, sizes + at, s);
        ctx->stats.kernel_launches += 4;
    }
    CU(cudaGetLastError());
    return SLB_OK;
}
