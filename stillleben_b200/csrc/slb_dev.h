// slb_dev.h — plain-old-data records shared by the host marshalling code (slb_host.cu) and the
// kernels. Everything here lives in HBM; layouts are chosen for the kernels, not for the caller:
//   * vertex stream is split at upload into pos4[] (xyz + one-based vertex id, one 128-bit load in
//     triangle setup) and attr[] (uv, normal, tangent: three 128-bit loads in the shade kernel);
//     the reference's 68-byte interleaved record (src/mesh_tools/consolidate.cpp:53-61) is only the
//     upload format,
//   * textures are RGBA8 with the whole mip chain in one allocation,
//   * per-batch scene state is two flat arrays (DFrame[], DDraw[]) uploaded with one copy.
#pragma once
#include <stdint.h>

#include "../../include/slb.h"

#define SLB_TILE 8                 // fine-raster tile: 8x8 pixels, one warp, two pixels per lane
#ifndef SLB_SETUP_CHUNK
#define SLB_SETUP_CHUNK 256        // triangles per block in the setup/bin kernels
#endif
#ifndef SLB_SETUP_MINB
#define SLB_SETUP_MINB 5           // resident set-up blocks per SM the register budget is chosen for (48 registers at 256 threads)
#endif
#define SLB_MAX_LEVELS 16

struct DTexture {
    const uint8_t* px;             // RGBA8, level l starts at texel level_off[l]
    int32_t w, h, n_levels;
    int32_t wrap_s, wrap_t, min_filter, mag_filter;
    int32_t kind, has_alpha;
    uint32_t level_off[SLB_MAX_LEVELS];
};

struct DCubeLevel { const float4* px; int32_t size; int32_t pad; };
struct DLightMap {
    DCubeLevel env[12]; int32_t n_env;
    DCubeLevel irr;
    DCubeLevel pre[5];
    const float4* lut; int32_t lut_size;
    int32_t n_lights;
    float light_directions[SLB_NUM_LIGHTS][3];
    float light_colors[SLB_NUM_LIGHTS][3];
};

struct DDraw {
    const float4* pos4;            // xyz, w = __uint_as_float(one-based vertex id)
    const float4* attr;            // 3 x float4 per vertex: (u,v,nx,ny) (nz,tx,ty,tz) (tw,0,0,0)
    const uint32_t* idx;           // already offset to the sub-mesh's first index
    uint32_t n_tris;
    uint32_t prim_base;            // sequence number of triangle 0 within the frame (increasing in submission order)
    uint32_t frame;                // frame index within the batch
    uint32_t pad0;
    float mvp[16];
    float meshToObject[16], objectToWorld[16];
    float normalToWorld[9];
    float base_color[4], emissive[4];
    float metallic, roughness;
    const DTexture* tex[5];        // base, normal, metallic-roughness, emissive, occlusion
    const DTexture* sticker;
    float stickerProj[16];
    float stickerRange[4];
    uint32_t class_index, instance_index;
    uint32_t flags;                // DRAW_*
    uint32_t pad;
};
enum {
    DRAW_FRAG_TEST = 1u,           // coverage depends on the fragment stage (alpha test / depth peel)
    DRAW_AFFINE = 2u               // meshToObject, objectToWorld and the view matrix all have the last row (0,0,0,1)
};

// What the binner needs of one draw. Camera views and shadow views go through the SAME setup / bin / raster
// kernels: a shadow map is just another view with its own tile grid, front faces culled and depth-only output.
struct DBinDraw {
    const float4* pos4;
    const uint32_t* idx;
    uint32_t n_tris;
    uint32_t chunk_base;           // first setup chunk of this draw within the batch
    uint32_t prim_base;            // sequence number of triangle 0 within its view
    uint32_t view;                 // index into DView[]
    uint32_t draw;                 // index into DDraw[] (camera views; used by the fragment-test path)
    uint32_t flags;                // DRAW_*
    float mvp[16];
};
// A huge sub-triangle of a camera view that is resolved per pixel inside the shade kernel instead of going through
// count / scan / emit / raster: its prepared edge set-up (k_contract.cuh SubTri, 64 B), key bits and pixel box.
#define SLB_HUGE_PER_VIEW 16       // more than this many per view: the rest takes the tiled path (both merge by minimum)
struct __align__(16) HugeRec {
    int32_t ax, ay, bx, by, cx, cy;
    float az, bz, cz;
    int32_t s;                     // orientation sign
    float inv2A;
    int32_t bias0, bias1, bias2;
    uint32_t seq, kbyte;
    int16_t px0, py0, px1, py1;    // pixel box (inclusive)
};
static_assert(sizeof(HugeRec) == 80, "HugeRec must be 80 bytes");

// What the shade kernel needs to shade a pixel of a huge sub-triangle WITHOUT re-doing the triangle's set-up: the vertex
// stage's outputs of the three ORIGINAL vertices (computed once per frame by k_huge_prepare instead of once per pixel — the
// background plane alone covers most of a table-top frame) plus the perspective terms of the sub-triangle's vertices.
struct __align__(16) HugeShade {
    float invw[3];                 // 1/w of the sub-triangle's vertices a, b, c
    uint32_t fast;                 // 1: usable (affine chain, no sticker); 0: the pixel takes the generic path
    float basis[3][3];             // barycentric basis of a, b, c w.r.t. the original triangle (identity if unclipped)
    uint32_t unit_basis, draw, front;
    uint32_t vid[3], vi[3];        // one-based vertex ids (the render target) and the three indices into the vertex buffers
    float objc[3][3], wc[3][3], cc[3][3], nW[3][3];   // per original vertex: object / world / camera position, world normal (normalised)
    float uv[3][2];
};
static_assert(sizeof(HugeShade) == 256, "HugeShade must be 256 bytes");

// Coarse occupancy of a shadow map: one bit per 16x16 texel block, set by every rasteriser that MAY write a texel of the
// block in this sub-batch. A PCF footprint whose blocks are all clear consists of untouched (= lit) texels only, so the 25
// taps need no loads (bit-identical result). word = by * SLB_SHADOW_MASK_ROW + (bx >> 5), bit = bx & 31. 2 KB per map: a
// set-up block accumulates its bits in shared memory and flushes the non-zero words with one RED.OR each.
#define SLB_SHADOW_MASK_SHIFT 4
#define SLB_SHADOW_MASK_ROW ((SLB_SHADOW_RES >> SLB_SHADOW_MASK_SHIFT) / 32)
#define SLB_SHADOW_MASK_WORDS ((SLB_SHADOW_RES >> SLB_SHADOW_MASK_SHIFT) * SLB_SHADOW_MASK_ROW)

struct DView {
    int32_t W, H, tiles_x, tiles_y;
    uint32_t tile_base;            // first tile of this view in the batch-wide tile arrays
    int32_t shadow;                // 1: cull front faces, write d24 only (render_pass.cpp:426-460)
    uint32_t frame;                // DFrame index (camera views)
    uint32_t tagbits;              // shadow views: generation tag << 24, stored above every d24 (see DFrame::shadow_tagbits)
    void* out;                     // uint64 keys[H*W] (camera) or uint32 tag | d24 [H*W] (shadow)
    uint32_t* mask;                // shadow views: SLB_SHADOW_MASK_WORDS words of block-occupancy bits (null: feature off)
    HugeRec* huge;                 // camera views: SLB_HUGE_PER_VIEW slots (null: feature off / shadow view)
    uint32_t* huge_n;              // how many were claimed (may exceed the capacity: the excess went to the tiled path)
};

// A primitive that needed polygon clipping, written once by the binner so that the raster's fragment test
// and the shade kernel fetch the clipped, snapped polygon instead of re-clipping it per pixel.
struct DPolyV { int32_t X, Y; float z, invw; float b[3]; };
struct ClipRec { uint32_t seq; int32_t n; DPolyV v[10]; };
#define SLB_MAX_CLIP 31            // per frame (slot + 1 fits 5 bits of the key); primitives beyond this are re-clipped where they are used

struct DFrame {
    int32_t W, H, tiles_x, tiles_y;
    uint32_t seq_shift;            // > 0: sequence number = (draw - draw_begin) << seq_shift | triangle (order preserving); 0: prim_base search
    uint32_t draw_begin, draw_end;
    uint32_t n_prims;
    float P[16], V[16], Pinv[16];
    float camPos[3];
    float manual_exposure;
    float lightDir[SLB_NUM_LIGHTS][3], lightCol[SLB_NUM_LIGHTS][3];
    int32_t lightActive[SLB_NUM_LIGHTS];
    int32_t ssao;
    float ambient[3];
    int32_t fused_tonemap;         // 1: shade kernel tone-maps and stores rgb itself (no post passes needed)
    // Shadow-map texels are tag << 24 | d24 with tag = 255 - generation, the generation counting the sub-batches that
    // reused the map pool since its last clear. Newer generations are SMALLER, so RED.MIN lets this sub-batch's depths
    // replace stale ones, and a texel nobody wrote this time compares >= any threshold of the current tag — "lit",
    // exactly like a cleared texel. The 1 GB memset per sub-batch is needed only once every 255 sub-batches.
    uint32_t shadow_tagbits;
    float shadowMat[SLB_NUM_LIGHTS][16];
    const uint32_t* shadowMap[SLB_NUM_LIGHTS];
    const uint32_t* shadowMask[SLB_NUM_LIGHTS];   // block-occupancy bits of each map (null: always take the taps)
    const DLightMap* lm;
    const float* peel;             // previous layer's coord target (HxWx4) or null
    const DTexture* bg_image;
    ClipRec* clip;                 // SLB_MAX_CLIP records
    uint32_t* clip_count;
    const HugeRec* huge;           // this frame's huge sub-triangles, resolved per pixel by the shade kernel
    const uint32_t* huge_n;
    HugeShade* huge_shade;         // SLB_HUGE_PER_VIEW records, filled by k_huge_prepare between set-up and shading
    uint64_t* keys;                // H*W visibility keys
    float4* hdr;                   // H*W pre-tone-map colour (post-pass path only)
    float4* scratch_normal;        // used by SSAO when the normal / cam-coord targets are not requested
    float4* scratch_cam;
    float* zplane;                 // camera-space z of every pixel as a dense plane (SSAO taps: 8 pixels per 32-byte sector instead of 2)
    float* ao;
    float* avg;                    // 4 floats: 1x1 mip level for auto exposure
    void* out[SLB_NUM_TARGETS];    // this frame's slice of each requested target (null = not requested)
};

// One (tile, sub-triangle) pair of the binner: the snapped sub-triangle itself, 48 bytes so that a
// tile's list is a 16-byte-aligned contiguous run that cp.async.bulk can stage into shared memory.
struct __align__(16) PairRec {
    int32_t ax, ay, bx, by, cx, cy;   // 24.8 fixed-point window coordinates
    float az, bz, cz;                 // window z in [0,1]
    uint32_t seq;                     // primitive sequence number within the frame
    uint32_t k_flags;                 // bits 0..2: fan index k of the sub-triangle (1..7); bits 3..7: ClipRec slot + 1 of a clipped
                                      // primitive (0: unclipped / not published); bit 8: needs fragment test; bits 16..: view
    uint32_t draw;                    // index into DDraw[] (fragment-test path only)
};
static_assert(sizeof(PairRec) == 48, "PairRec must be 48 bytes");

// a non-empty tile: what one raster warp needs to start (written by the scan's fix-up pass)
struct ActiveTile { uint32_t tile, beg, count, pad; };

// visibility key: depth24 << 40 | seq << 8 | (clip slot + 1) << 3 | k   (min == GL_LESS + first draw wins on ties)
#define SLB_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
