// k_render.cu — the per-frame hot path as sm_100a kernels:
//   k_bin<false/true>   triangle setup + tile binning (count pass / emit pass)
//   k_scan              exclusive scan of the per-tile counts
//   k_raster            warp-per-tile fine raster: per-tile pair lists staged through TMA bulk copies
//                       (cp.async.bulk + mbarrier) into shared memory, warp-ballot edge tests,
//                       per-lane z-compare, 64-bit visibility keys
//   (shadow maps are extra depth-only VIEWS of the same three kernels: front faces culled, d24 output)
//   k_shade             vertex + fragment stage of the visible fragment of every pixel, fused tone
//                       map, all render targets stored with 128-bit coalesced writes
// Replaces steps 6-10 and 14 of sl::RenderPass::render (reference: src/render_pass.cpp:407-622,696-710).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "k_frag.cuh"
#include "kernels.h"
#include "slb_dev.h"

using namespace slbk;

// ---------------------------------------------------------------------------------------------
// setup + binning
// ---------------------------------------------------------------------------------------------
// can any pixel centre of tile (tx,ty) be covered? (conservative: tests the most-inside corner per edge)
__device__ __forceinline__ bool tile_may_overlap(const SubTri& t, int tx, int ty, int W, int H) {
    int x0 = tx * SLB_TILE, y0 = ty * SLB_TILE;
    int x1 = min(x0 + SLB_TILE - 1, W - 1), y1 = min(y0 + SLB_TILE - 1, H - 1);
    // d(s*w0)/dpx = -s*(cy-by), d(s*w0)/dpy = s*(cx-bx), etc.
    int cx, cy;
    cx = ((long long)t.s * -(t.cy - t.by) > 0) ? x1 : x0; cy = ((long long)t.s * (t.cx - t.bx) > 0) ? y1 : y0;
    if (t.s * edge_fn(t.bx, t.by, t.cx, t.cy, cx * 256 + 128, cy * 256 + 128) + t.bias0 < 0) return false;
    cx = ((long long)t.s * -(t.ay - t.cy) > 0) ? x1 : x0; cy = ((long long)t.s * (t.ax - t.cx) > 0) ? y1 : y0;
    if (t.s * edge_fn(t.cx, t.cy, t.ax, t.ay, cx * 256 + 128, cy * 256 + 128) + t.bias1 < 0) return false;
    cx = ((long long)t.s * -(t.by - t.ay) > 0) ? x1 : x0; cy = ((long long)t.s * (t.bx - t.ax) > 0) ? y1 : y0;
    if (t.s * edge_fn(t.ax, t.ay, t.bx, t.by, cx * 256 + 128, cy * 256 + 128) + t.bias2 < 0) return false;
    return true;
}

struct BigEntry { PairRec rec; int tx0, ty0, ntx, nty; };
#ifndef SLB_MID_GROUP
#define SLB_MID_GROUP 8     // lanes that rasterise one mid-size triangle together in the set-up kernel
#endif
#define SLB_BIG_QUEUE 48
#define SLB_BIG_TILES 24   // sub-triangles touching more tiles than this are binned by the whole block

template <bool EMIT>
__device__ __forceinline__ void bin_pair(uint32_t tile, const PairRec& rec, uint32_t* __restrict__ tile_count,
                                         const uint32_t* __restrict__ tile_off, PairRec* __restrict__ pairs, uint32_t capacity) {
    if (EMIT) {
        // the scan left the END of the tile's segment in tile_count: one atomic yields the global slot
        uint32_t at = atomicSub(tile_count + tile, 1u) - 1u;
        if (at < capacity) pairs[at] = rec;
    } else {
        atomicAdd(tile_count + tile, 1u);
    }
}

// Warp-aggregated variants: neighbouring triangles of a mesh mostly fall into the same tile, so the lanes of a
// warp that target the same tile are grouped with __match_any_sync and only the group leader touches the
// counter (one atomic per distinct tile per warp instead of one per lane). Must be called by converged lanes.
__device__ __forceinline__ void count_pair_agg(uint32_t tile, bool valid, uint32_t* __restrict__ tile_count) {
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(mask, valid ? tile : 0xFFFFFFFFu);
    if (valid && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(tile_count + tile, (uint32_t)__popc(peers));
}
__device__ __forceinline__ void emit_pair_agg(uint32_t tile, bool valid, const PairRec& rec, uint32_t* __restrict__ tile_count,
                                              PairRec* __restrict__ pairs, uint32_t capacity) {
    const unsigned mask = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(mask, valid ? tile : 0xFFFFFFFFu);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == (unsigned)leader) base = atomicSub(tile_count + tile, (uint32_t)__popc(peers));   // old value = end of free space
    base = __shfl_sync(mask, base, leader);
    if (valid) {
        const uint32_t at = base - 1u - (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (at < capacity) pairs[at] = rec;
    }
}

// Tile walk of one snapped sub-triangle: COUNT pass (EMIT = false) increments the per-tile counters, EMIT pass
// writes the PairRec into each tile's segment. Sub-triangles touching more than SLB_BIG_TILES tiles are queued
// for the whole block. Returns the number of pairs produced (a queued one reports 1).
template <bool EMIT>
__device__ __forceinline__ int bin_tiles(const PairRec& rec, long long twoA, int px0, int py0, int px1, int py1, int W, int H, int tiles_x,
                                         uint32_t tile_base, uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_off,
                                         PairRec* __restrict__ pairs, uint32_t capacity, BigEntry* s_big, int* s_nbig) {
    const int tx0 = px0 / SLB_TILE, tx1 = px1 / SLB_TILE, ty0 = py0 / SLB_TILE, ty1 = py1 / SLB_TILE;
    const int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
    if (ntx * nty == 1) {
        bin_pair<EMIT>(tile_base + ty0 * tiles_x + tx0, rec, tile_count, tile_off, pairs, capacity);
        return 1;
    }
    if (ntx * nty == 2) {
        bin_pair<EMIT>(tile_base + ty0 * tiles_x + tx0, rec, tile_count, tile_off, pairs, capacity);
        bin_pair<EMIT>(tile_base + ty1 * tiles_x + tx1, rec, tile_count, tile_off, pairs, capacity);
        return 2;
    }
    if (ntx * nty > SLB_BIG_TILES) {
        int q = atomicAdd(s_nbig, 1);
        if (q < SLB_BIG_QUEUE) {
            s_big[q].rec = rec; s_big[q].tx0 = tx0; s_big[q].ty0 = ty0; s_big[q].ntx = ntx; s_big[q].nty = nty;
            return 1;
        }
    }
    SubTri st;
    st.ax = rec.ax; st.ay = rec.ay; st.bx = rec.bx; st.by = rec.by; st.cx = rec.cx; st.cy = rec.cy;
    st.s = twoA > 0 ? 1 : -1;
    st.bias0 = top_left(rec.cx - rec.bx, rec.cy - rec.by, st.s) ? 0 : -1;
    st.bias1 = top_left(rec.ax - rec.cx, rec.ay - rec.cy, st.s) ? 0 : -1;
    st.bias2 = top_left(rec.bx - rec.ax, rec.by - rec.ay, st.s) ? 0 : -1;
    int n = 0;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx)
            if (tile_may_overlap(st, tx, ty, W, H)) {
                bin_pair<EMIT>(tile_base + ty * tiles_x + tx, rec, tile_count, tile_off, pairs, capacity);
                ++n;
            }
    return n;
}
// the tail every bin kernel shares: the block walks the tiles of its queued large sub-triangles together
template <bool EMIT>
__device__ __forceinline__ void bin_big_queue(const BigEntry* s_big, int nbig, const DView* __restrict__ views,
                                              uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_off,
                                              PairRec* __restrict__ pairs, uint32_t capacity) {
    for (int q = 0; q < nbig; ++q) {
        const BigEntry& e = s_big[q];
        const DView& v = views[e.rec.k_flags >> 16];
        SubTri st;
        make_subtri(e.rec.ax, e.rec.ay, e.rec.bx, e.rec.by, e.rec.cx, e.rec.cy, e.rec.az, e.rec.bz, e.rec.cz, st);
        const int n = e.ntx * e.nty;
        for (int i = threadIdx.x; i < n; i += SLB_SETUP_CHUNK) {
            int tx = e.tx0 + i % e.ntx, ty = e.ty0 + i / e.ntx;
            if (tile_may_overlap(st, tx, ty, v.W, v.H))
                bin_pair<EMIT>(v.tile_base + ty * v.tiles_x + tx, e.rec, tile_count, tile_off, pairs, capacity);
        }
    }
}

// pixel box of a snapped triangle inside a W x H viewport; false if it contains no pixel centre
__device__ __forceinline__ bool pixel_box(int ax, int ay, int bx, int by, int cx, int cy, int W, int H, int& px0, int& py0, int& px1, int& py1) {
    const int xmin = min(ax, min(bx, cx)), xmax = max(ax, max(bx, cx));
    const int ymin = min(ay, min(by, cy)), ymax = max(ay, max(by, cy));
    px0 = max(0, (xmin - 128 + 255) >> 8); px1 = min(W - 1, (xmax - 128) >> 8);
    py0 = max(0, (ymin - 128 + 255) >> 8); py1 = min(H - 1, (ymax - 128) >> 8);
    return px0 <= px1 && py0 <= py1;
}

// survivor test of one snapped sub-triangle (setup pass): fills `rec` and the tile range if it survives
__device__ __forceinline__ bool survivor_test(int ax, int ay, int bx, int by, int cx, int cy, float az, float bz, float cz, uint32_t seq,
                                              uint32_t k_flags, uint32_t draw, const DView& v, PairRec& rec, long long& twoA, int& px0,
                                              int& py0, int& px1, int& py1) {
    if (!pixel_box(ax, ay, bx, by, cx, cy, v.W, v.H, px0, py0, px1, py1)) return false;   // dies before any 64-bit arithmetic
    twoA = edge_fn(ax, ay, bx, by, cx, cy);
    if (twoA == 0) return false;
    if (v.shadow && twoA < 0) return false;   // shadow views cull FRONT faces (render_pass.cpp:428-429)
    rec.ax = ax; rec.ay = ay; rec.bx = bx; rec.by = by; rec.cx = cx; rec.cy = cy;
    rec.az = az; rec.bz = bz; rec.cz = cz;
    rec.seq = seq; rec.k_flags = k_flags; rec.draw = draw;
    return true;
}
__device__ __forceinline__ bool setup_subtri_count(int ax, int ay, int bx, int by, int cx, int cy, float az, float bz, float cz, uint32_t seq,
                                                   uint32_t k_flags, uint32_t draw, const DView& v, uint32_t* __restrict__ tile_count,
                                                   BigEntry* s_big, int* s_nbig, PairRec& rec) {
    long long twoA; int px0, py0, px1, py1;
    if (!survivor_test(ax, ay, bx, by, cx, cy, az, bz, cz, seq, k_flags, draw, v, rec, twoA, px0, py0, px1, py1)) return false;
    return bin_tiles<false>(rec, twoA, px0, py0, px1, py1, v.W, v.H, v.tiles_x, v.tile_base, tile_count, nullptr, nullptr, 0, s_big, s_nbig) > 0;
}

// Survivors of the tiled path. Ordinary ones are compacted into survivors[0 .. normal_cap); HUGE ones (more than
// SLB_HUGE_TILES tiles: the background plane, close-up faces) go to survivors[normal_cap + i] and are emitted by
// a kernel that gives each of them a whole block, so a handful of screen-filling triangles never serialises
// on one thread block. counters: [0] ordinary survivors, [1] overflow flag, [2] huge survivors.
#define SLB_HUGE_TILES 2
struct SurvOut { PairRec* survivors; uint32_t* counters; uint32_t normal_cap, huge_cap; };
__device__ __forceinline__ void append_huge(const SurvOut& so, const PairRec& rec) {
    const uint32_t at = atomicAdd(&so.counters[2], 1u);
    if (at < so.huge_cap) so.survivors[so.normal_cap + at] = rec; else atomicExch(&so.counters[1], 1u);
}
// Camera views resolve their first SLB_HUGE_PER_VIEW huge sub-triangles (the background plane, close-up faces) per
// pixel in the shade kernel: no tile counting, no pairs, no raster warp for them. Returns false when the view has no
// slot left (or the feature is off): the caller then sends the record down the tiled path; both routes merge by min.
// (out of line and by value: rare, and the caller's record must not be pinned to local memory)
static __device__ __noinline__ bool claim_huge_slot(HugeRec* huge, uint32_t* huge_n, const PairRec rec, int px0, int py0, int px1, int py1) {
    if (!huge || (rec.k_flags & 0x100u)) return false;   // fragment-tested draws need the raster's discard logic
    const uint32_t slot = atomicAdd(huge_n, 1u);
    if (slot >= SLB_HUGE_PER_VIEW) return false;
    SubTri st;
    make_subtri(rec.ax, rec.ay, rec.bx, rec.by, rec.cx, rec.cy, rec.az, rec.bz, rec.cz, st);
    HugeRec h;
    h.ax = st.ax; h.ay = st.ay; h.bx = st.bx; h.by = st.by; h.cx = st.cx; h.cy = st.cy;
    h.az = st.az; h.bz = st.bz; h.cz = st.cz; h.s = st.s; h.inv2A = st.inv2A;
    h.bias0 = st.bias0; h.bias1 = st.bias1; h.bias2 = st.bias2;
    h.seq = rec.seq; h.kbyte = rec.k_flags & 0xffu;
    h.px0 = (int16_t)px0; h.py0 = (int16_t)py0; h.px1 = (int16_t)px1; h.py1 = (int16_t)py1;
    huge[slot] = h;
    return true;
}
__device__ __forceinline__ int tiles_of_box(int px0, int py0, int px1, int py1) {
    return (px1 / SLB_TILE - px0 / SLB_TILE + 1) * (py1 / SLB_TILE - py0 / SLB_TILE + 1);
}

// primitives that need polygon clipping (rare: the background plane, triangles crossing the near plane): the
// clipped polygon is published once for the fragment test / shade kernel; survivors are appended one by one
static __device__ __noinline__ void setup_clipped_prim(const float* mvp, float3 p0, float3 p1, float3 p2, uint32_t seq, uint32_t flags,
                                                       uint32_t draw, const DView& v, const DFrame* fr, uint32_t* __restrict__ tile_count,
                                                       const SurvOut so /* by value: a reference would pin the kernel's copy to local memory */,
                                                       BigEntry* s_big, int* s_nbig) {
    PrimSetup ps;
    if (!setup_prim(mvp, p0, p1, p2, v.W, v.H, ps)) return;
    uint32_t slot_bits = 0;   // (ClipRec slot + 1) << 3, carried in the key so the consumers index the record directly
    if (fr) {
        uint32_t slot = atomicAdd(fr->clip_count, 1u);
        if (slot < SLB_MAX_CLIP) {
            slot_bits = (slot + 1u) << 3;
            ClipRec& cr = fr->clip[slot];
            cr.seq = seq; cr.n = ps.n;
            for (int i = 0; i < ps.n; ++i) {
                cr.v[i].X = ps.v[i].X; cr.v[i].Y = ps.v[i].Y; cr.v[i].z = ps.v[i].z; cr.v[i].invw = ps.v[i].invw;
                cr.v[i].b[0] = ps.v[i].b[0]; cr.v[i].b[1] = ps.v[i].b[1]; cr.v[i].b[2] = ps.v[i].b[2];
            }
        }
    }
    for (int k = 1; k + 1 < ps.n; ++k) {
        const PolyV &a = ps.v[0], &b = ps.v[k], &c = ps.v[k + 1];
        PairRec rec;
        long long twoA; int px0, py0, px1, py1;
        if (!survivor_test(a.X, a.Y, b.X, b.Y, c.X, c.Y, a.z, b.z, c.z, seq, (uint32_t)k | slot_bits | flags, draw, v, rec, twoA, px0, py0, px1, py1))
            continue;
        const bool huge = tiles_of_box(px0, py0, px1, py1) > SLB_HUGE_TILES;
        if (huge && claim_huge_slot(v.huge, v.huge_n, rec, px0, py0, px1, py1)) continue;   // resolved in the shade kernel
        if (bin_tiles<false>(rec, twoA, px0, py0, px1, py1, v.W, v.H, v.tiles_x, v.tile_base, tile_count, nullptr, nullptr, 0, s_big, s_nbig) > 0) {
            if (huge) { append_huge(so, rec); continue; }
            uint32_t at = atomicAdd(&so.counters[0], 1u);
            if (at < so.normal_cap) so.survivors[at] = rec; else atomicExch(&so.counters[1], 1u);
        }
    }
}

// Direct path for SMALL unclipped triangles (the bulk of a 16k-triangle mesh at table-top distance covers a
// handful of pixels): rasterised by ONE thread right in the setup kernel, every covered pixel merged into the
// view's output with a fire-and-forget L2 atomic (RED.MIN: u64 visibility key, or u32 d24 for a shadow view).
// No survivor record, no per-tile pair, no raster warp. The result is the same minimum the tiled path takes,
// so coverage, depth and ids stay bit-identical (contract C6/C7); all edge arithmetic fits 32 bits because the
// triangle spans < 64 px (coordinates taken relative to vertex a).
__device__ __forceinline__ void red_min_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_min_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.min.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// G = 1: the thread walks the triangle's pixel box alone.  G = 32: the warp walks it together, lane l taking the
// box pixels l, l+32, ... in row-major order (mid-size triangles, e.g. a 16k-triangle mesh seen from a 2048^2
// shadow view), so the cost per triangle is box/32 iterations instead of one iteration per touched tile.
// shadow views: mark the texel blocks under a pixel box in the map's block-occupancy mask (see SLB_SHADOW_MASK_WORDS). Called by
// a whole warp with the bounding box of the triangles its lanes are about to rasterise (a superset of what they write, which
// is all the mask promises): the l-th participating lane takes block rows l, l + n_lanes, ... of the box, so the marking is a handful of parallel RED.ORs per warp.
__device__ __forceinline__ void mark_shadow_blocks(uint32_t* __restrict__ mask, int px0, int py0, int px1, int py1, int lane, int n_lanes) {
    const int bx0 = px0 >> SLB_SHADOW_MASK_SHIFT, bx1 = px1 >> SLB_SHADOW_MASK_SHIFT;
    for (int by = (py0 >> SLB_SHADOW_MASK_SHIFT) + lane; by <= (py1 >> SLB_SHADOW_MASK_SHIFT); by += n_lanes)
        for (int w = bx0 >> 5; w <= bx1 >> 5; ++w) {
            const int lo = max(bx0 - w * 32, 0), hi = min(bx1 - w * 32, 31);
            atomicOr(mask + by * SLB_SHADOW_MASK_ROW + w, (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo));   // result unused: RED.OR
        }
}
template <bool SHADOW, int G>
__device__ __forceinline__ void raster_direct(int ax, int ay, int bx, int by, int cx, int cy, float az, float bz, float cz, uint32_t seq,
                                              void* __restrict__ out, int W, int H, int lane, uint32_t tagbits, uint32_t* __restrict__ mask) {
    const int xmin = min(ax, min(bx, cx)), xmax = max(ax, max(bx, cx));
    const int ymin = min(ay, min(by, cy)), ymax = max(ay, max(by, cy));
    const int px0 = max(0, (xmin - 128 + 255) >> 8), px1 = min(W - 1, (xmax - 128) >> 8);
    const int py0 = max(0, (ymin - 128 + 255) >> 8), py1 = min(H - 1, (ymax - 128) >> 8);
    if (SHADOW && mask) {   // occupancy mask of the map: the box of the warp's triangles, one reduction per bound
        if (G == 1) {
            const unsigned act = __activemask();
            const int wx0 = __reduce_min_sync(act, px0), wy0 = __reduce_min_sync(act, py0), wx1 = __reduce_max_sync(act, px1), wy1 = __reduce_max_sync(act, py1);
            mark_shadow_blocks(mask, wx0, wy0, wx1, wy1, __popc(act & ((1u << (threadIdx.x & 31)) - 1u)), __popc(act));   // rank among the active lanes
        } else mark_shadow_blocks(mask, px0, py0, px1, py1, lane, G);
    }
    const int rbx = bx - ax, rby = by - ay, rcx = cx - ax, rcy = cy - ay;
    const int twoA = rbx * rcy - rby * rcx;
    const int sg = twoA > 0 ? 1 : -1;
    const float inv2A = __frcp_rn(__int2float_rn(twoA));
    const int bias0 = top_left(cx - bx, cy - by, sg) ? 0 : -1;
    const int bias1 = top_left(ax - cx, ay - cy, sg) ? 0 : -1;
    const int bias2 = top_left(bx - ax, by - ay, sg) ? 0 : -1;
    const int qx = px0 * 256 + 128 - ax, qy = py0 * 256 + 128 - ay;
    const int e0r = sg * ((rcx - rbx) * (qy - rby) - (rcy - rby) * (qx - rbx)) + bias0;   // s*edge(b,c,p) + bias at (px0, py0)
    const int e1r = sg * (rcy * (qx - rcx) - rcx * (qy - rcy)) + bias1;                   // s*edge(c,a,p) + bias
    const int e2r = sg * (rbx * qy - rby * qx) + bias2;                                   // s*edge(a,b,p) + bias
    const int dx0 = -sg * (rcy - rby) * 256, dy0 = sg * (rcx - rbx) * 256;
    const int dx1 = sg * rcy * 256, dy1 = -sg * rcx * 256;
    const int dx2 = -sg * rby * 256, dy2 = sg * rbx * 256;
    const float zb = __fsub_rn(bz, az), zc = __fsub_rn(cz, az);
    const unsigned long long lowkey = ((unsigned long long)seq << 8) | 1ull;
    auto plot = [&](int e1, int e2, int x, int y) {   // contract C7 on a covered pixel, merged with RED.MIN
        const int w1 = sg * (e1 - bias1), w2 = sg * (e2 - bias2);
        const float q1 = __fmul_rn(__int2float_rn(w1), inv2A), q2 = __fmul_rn(__int2float_rn(w2), inv2A);
        float z = __fmaf_rn(q2, zc, __fmaf_rn(q1, zb, az));
        z = fminf(fmaxf(z, 0.0f), 1.0f);
        const uint32_t d24 = __float2uint_rn(__fmul_rn(z, 16777215.0f));
        if (SHADOW) red_min_u32(reinterpret_cast<uint32_t*>(out) + (size_t)y * W + x, tagbits | d24);   // generation tag above the depth
        else red_min_u64(reinterpret_cast<unsigned long long*>(out) + (size_t)y * W + x, ((unsigned long long)d24 << 40) | lowkey);
    };
    if (G == 1) {
        int f0 = e0r, f1 = e1r, f2 = e2r;
        for (int y = py0; y <= py1; ++y, f0 += dy0, f1 += dy1, f2 += dy2) {
            int e0 = f0, e1 = f1, e2 = f2;
            for (int x = px0; x <= px1; ++x, e0 += dx0, e1 += dx1, e2 += dx2)
                if ((e0 | e1 | e2) >= 0) plot(e1, e2, x, y);
        }
    } else {
        const int bw = px1 - px0 + 1, n = bw * (py1 - py0 + 1);
        const int qstep = G / bw, rstep = G - qstep * bw;   // advancing G pixels in row-major order = (qstep rows, rstep columns)
        int x = lane % bw, y = lane / bw;
        for (int i = lane; i < n; i += G) {
            const int e0 = e0r + x * dx0 + y * dy0, e1 = e1r + x * dx1 + y * dy1, e2 = e2r + x * dx2 + y * dy2;
            if ((e0 | e1 | e2) >= 0) plot(e1, e2, px0 + x, py0 + y);
            x += rstep; y += qstep;
            if (x >= bw) { x -= bw; ++y; }
        }
    }
}

// PASS 1 — triangle setup: one thread per triangle, one block per 256-triangle chunk of one draw (camera or
// shadow view). Transforms, clips, snaps (contract C1-C6), counts the (tile, sub-triangle) pairs per tile and
// appends every surviving sub-triangle, already snapped, to the compact survivors[] array.
__global__ void __launch_bounds__(SLB_SETUP_CHUNK, SLB_SETUP_MINB) k_setup(const DView* __restrict__ views, const DFrame* __restrict__ frames,
                                                              const DBinDraw* __restrict__ bdraws, const uint32_t* __restrict__ chunk_draw,
                                                              uint32_t* __restrict__ tile_count, const SurvOut so, int direct_max, int warp_max) {
    __shared__ float s_mvp[16];
    __shared__ BigEntry s_big[SLB_BIG_QUEUE];
    __shared__ int s_nbig;
    __shared__ uint32_t s_wcount[SLB_SETUP_CHUNK / 32], s_base;
    __shared__ int s_dq[10][SLB_SETUP_CHUNK];   // direct-path queue (structure of arrays: conflict-free)
    __shared__ int s_ndirect, s_nmid, s_mid_next;
    const uint32_t di = __ldg(chunk_draw + blockIdx.x);   // host-built table: setup chunk -> bin draw
    const DBinDraw& d = bdraws[di];
    if (threadIdx.x < 16) s_mvp[threadIdx.x] = d.mvp[threadIdx.x];
    if (threadIdx.x == 32) s_nbig = 0;
    if (threadIdx.x == 33) { s_ndirect = 0; s_nmid = 0; s_mid_next = 0; }
    __syncthreads();
    const DView& v = views[d.view];
    const uint32_t tri = (blockIdx.x - d.chunk_base) * SLB_SETUP_CHUNK + threadIdx.x;
    PairRec rec;
    bool has_rec = false;
    uint32_t dt0 = 0xFFFFFFFFu, dt1 = 0xFFFFFFFFu;   // tiles of a one / two tile survivor
    if (tri < d.n_tris) {
        const uint32_t* ip = d.idx + 3 * (size_t)tri;
        const uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
        const float4 p0 = __ldg(d.pos4 + i0), p1 = __ldg(d.pos4 + i1), p2 = __ldg(d.pos4 + i2);
        ClipV c0, c1, c2;
        xform_clip(s_mvp, p0.x, p0.y, p0.z, c0);
        xform_clip(s_mvp, p1.x, p1.y, p1.z, c1);
        xform_clip(s_mvp, p2.x, p2.y, p2.z, c2);
        const int fc0 = frustum_code(c0), fc1 = frustum_code(c1), fc2 = frustum_code(c2);
        if (!(fc0 & fc1 & fc2)) {
            const uint32_t seq = d.prim_base + tri;
            const uint32_t flags = ((d.flags & DRAW_FRAG_TEST) ? 0x100u : 0u) | (d.view << 16);
            if (!inside_frustum(fc0 | fc1 | fc2, c0, c1, c2) && (need_mask(c0) | need_mask(c1) | need_mask(c2))) {
                setup_clipped_prim(s_mvp, make_float3(p0.x, p0.y, p0.z), make_float3(p1.x, p1.y, p1.z), make_float3(p2.x, p2.y, p2.z), seq, flags,
                                   d.draw, v, v.shadow ? nullptr : &frames[v.frame], tile_count, so, s_big, &s_nbig);
            } else {
                const float hw = 0.5f * (float)v.W, hh = 0.5f * (float)v.H;
                PolyV a, b, c;
                if (project_vertex(c0, hw, hh, a) && project_vertex(c1, hw, hh, b) && project_vertex(c2, hw, hh, c)) {
                    long long twoA; int px0, py0, px1, py1;
                    if (survivor_test(a.X, a.Y, b.X, b.Y, c.X, c.Y, a.z, b.z, c.z, seq, 1u | flags, d.draw, v, rec, twoA, px0, py0, px1, py1)) {
                        const int tx0 = px0 / SLB_TILE, tx1 = px1 / SLB_TILE, ty0 = py0 / SLB_TILE, ty1 = py1 / SLB_TILE;
                        const int ext_x = max(a.X, max(b.X, c.X)) - min(a.X, min(b.X, c.X)), ext_y = max(a.Y, max(b.Y, c.Y)) - min(a.Y, min(b.Y, c.Y));
                        if (!(flags & 0x100u) && ext_x < 16384 && ext_y < 16384 && (px1 - px0 + 1) * (py1 - py0 + 1) <= warp_max) {
                            // small / mid-size triangle: queued for the direct path below (never becomes a survivor). Small
                            // ones fill the queue from the front (one thread each), mid-size ones from the back (one warp each).
                            // (Counting-sorting the small ones by box size to even out the warps was measured: no gain.)
                            const bool small = (px1 - px0 + 1) * (py1 - py0 + 1) <= direct_max;
                            const int q = small ? atomicAdd(&s_ndirect, 1) : SLB_SETUP_CHUNK - 1 - atomicAdd(&s_nmid, 1);
                            s_dq[0][q] = a.X; s_dq[1][q] = a.Y; s_dq[2][q] = b.X; s_dq[3][q] = b.Y; s_dq[4][q] = c.X; s_dq[5][q] = c.Y;
                            s_dq[6][q] = __float_as_int(a.z); s_dq[7][q] = __float_as_int(b.z); s_dq[8][q] = __float_as_int(c.z);
                            s_dq[9][q] = (int)seq;
                        } else if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= 2) {   // one or two tiles: counted warp-aggregated below
                            has_rec = true;
                            dt0 = v.tile_base + ty0 * v.tiles_x + tx0; dt1 = v.tile_base + ty1 * v.tiles_x + tx1;
                        } else if (!claim_huge_slot(v.huge, v.huge_n, rec, px0, py0, px1, py1)) {   // more than two tiles: huge
                            has_rec = bin_tiles<false>(rec, twoA, px0, py0, px1, py1, v.W, v.H, v.tiles_x, v.tile_base, tile_count, nullptr, nullptr,
                                                       0, s_big, &s_nbig) > 0;
                            if (has_rec) { append_huge(so, rec); has_rec = false; }
                        }
                    }
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, dt0 != 0xFFFFFFFFu)) {   // rare now that small triangles take the direct path: skip the match
        count_pair_agg(dt0, dt0 != 0xFFFFFFFFu, tile_count);
        count_pair_agg(dt1, dt1 != 0xFFFFFFFFu && dt1 != dt0, tile_count);
    }
    // compact the block's survivors: one global atomic per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, has_rec);
    if (lane == 0) s_wcount[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < SLB_SETUP_CHUNK / 32; ++w) { uint32_t c = s_wcount[w]; s_wcount[w] = total; total += c; }
        s_base = total ? atomicAdd(&so.counters[0], total) : 0u;
    }
    __syncthreads();
    if (has_rec) {
        const uint32_t at = s_base + s_wcount[warp] + __popc(m & ((1u << lane) - 1u));
        if (at < so.normal_cap) so.survivors[at] = rec; else atomicExch(&so.counters[1], 1u);
    }
    bin_big_queue<false>(s_big, min(s_nbig, SLB_BIG_QUEUE), views, tile_count, nullptr, nullptr, 0);
    // direct path: the block's small triangles, densely packed, one per thread ...
#define SLB_DQ(q) s_dq[0][q], s_dq[1][q], s_dq[2][q], s_dq[3][q], s_dq[4][q], s_dq[5][q], __int_as_float(s_dq[6][q]), \
                  __int_as_float(s_dq[7][q]), __int_as_float(s_dq[8][q]), (uint32_t)s_dq[9][q]
    if ((int)threadIdx.x < s_ndirect) {
        const int q = threadIdx.x;
        if (v.shadow) raster_direct<true, 1>(SLB_DQ(q), v.out, v.W, v.H, 0, v.tagbits, v.mask);
        else raster_direct<false, 1>(SLB_DQ(q), v.out, v.W, v.H, 0, 0u, nullptr);
    }
    // ... then the mid-size ones, one per GROUP of SLB_MID_GROUP lanes (four triangles in flight per warp), handed out
    // dynamically. A whole warp per triangle replicates the triangle's set-up 32 times and leaves most lanes idle on boxes of
    // a few dozen pixels; one thread per triangle makes the warp wait for its largest box. Groups of 8 sit in between.
    const int nmid = s_nmid;
    constexpr int MG = SLB_MID_GROUP;
    const int gl = lane & (MG - 1);
    const unsigned gmask = (MG == 32 ? 0xffffffffu : ((1u << MG) - 1u)) << (lane & ~(MG - 1));
    for (;;) {
        int m = 0;
        if (gl == 0) m = atomicAdd(&s_mid_next, 1);
        m = __shfl_sync(gmask, m, 0, MG);
        if (m >= nmid) break;
        const int q = SLB_SETUP_CHUNK - 1 - m;
        if (v.shadow) raster_direct<true, MG>(SLB_DQ(q), v.out, v.W, v.H, gl, v.tagbits, v.mask);
        else raster_direct<false, MG>(SLB_DQ(q), v.out, v.W, v.H, gl, 0u, nullptr);
    }
#undef SLB_DQ
}

// PASS 2 — emit: one thread per ordinary SURVIVOR (dense: no culled triangles, no vertex fetch, no transform). Ordinary
// survivors touch one or two tiles (everything larger went to the huge list, PASS 2b), so each thread writes its record
// into at most two tile segments, with one atomic per distinct tile per warp.
__global__ void __launch_bounds__(SLB_SETUP_CHUNK, 4) k_emit(const DView* __restrict__ views, const PairRec* __restrict__ survivors,
                                                             uint32_t n_survivors, uint32_t* __restrict__ tile_count,
                                                             PairRec* __restrict__ pairs, uint32_t capacity) {
    static_assert(SLB_HUGE_TILES == 2, "k_emit handles one- and two-tile records only");
    const uint32_t i = blockIdx.x * SLB_SETUP_CHUNK + threadIdx.x;
    PairRec rec;
    uint32_t dt0 = 0xFFFFFFFFu, dt1 = 0xFFFFFFFFu;
    if (i < n_survivors) {
        rec = survivors[i];
        const DView& v = views[rec.k_flags >> 16];
        int px0, py0, px1, py1;
        pixel_box(rec.ax, rec.ay, rec.bx, rec.by, rec.cx, rec.cy, v.W, v.H, px0, py0, px1, py1);
        const int tx0 = px0 / SLB_TILE, tx1 = px1 / SLB_TILE, ty0 = py0 / SLB_TILE, ty1 = py1 / SLB_TILE;
        dt0 = v.tile_base + ty0 * v.tiles_x + tx0; dt1 = v.tile_base + ty1 * v.tiles_x + tx1;   // equal for a one-tile record
    }
    emit_pair_agg(dt0, dt0 != 0xFFFFFFFFu, rec, tile_count, pairs, capacity);
    emit_pair_agg(dt1, dt1 != 0xFFFFFFFFu && dt1 != dt0, rec, tile_count, pairs, capacity);
}

// PASS 2b — one block per HUGE survivor: the block's threads stride over the tiles of its bounding box.
#define SLB_HUGE_THREADS 64
__global__ void __launch_bounds__(SLB_HUGE_THREADS) k_emit_huge(const DView* __restrict__ views, const PairRec* __restrict__ huge,
                                                               uint32_t* __restrict__ tile_count, PairRec* __restrict__ pairs, uint32_t capacity) {
    const PairRec rec = huge[blockIdx.x];
    const DView& v = views[rec.k_flags >> 16];
    int px0, py0, px1, py1;
    pixel_box(rec.ax, rec.ay, rec.bx, rec.by, rec.cx, rec.cy, v.W, v.H, px0, py0, px1, py1);
    const int tx0 = px0 / SLB_TILE, ty0 = py0 / SLB_TILE;
    const int ntx = px1 / SLB_TILE - tx0 + 1, nty = py1 / SLB_TILE - ty0 + 1;
    SubTri st;
    make_subtri(rec.ax, rec.ay, rec.bx, rec.by, rec.cx, rec.cy, rec.az, rec.bz, rec.cz, st);
    for (int i = threadIdx.x; i < ntx * nty; i += SLB_HUGE_THREADS) {
        const int tx = tx0 + i % ntx, ty = ty0 + i / ntx;
        if (tile_may_overlap(st, tx, ty, v.W, v.H)) bin_pair<true>(v.tile_base + ty * v.tiles_x + tx, rec, tile_count, nullptr, pairs, capacity);
    }
}

// Exclusive scan of the n tile counts into off[0..n] (off[n] = total pairs) fused with the compaction of the
// non-empty tiles into active[]: the scanned value packs (pair count, non-empty flag) into 64 bits, so one scan
// yields both the tile's pair offset and its slot in the active list. Three kernels: block-local scans of 4096
// items, a scan of the block totals, and a fix-up pass that also writes the ActiveTile records.
#define SLB_SCAN_ITEMS 4
#define SLB_SCAN_BLOCK (1024 * SLB_SCAN_ITEMS)
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_warp, unsigned long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + inc - v;
}
__device__ __forceinline__ unsigned long long pack_count(uint32_t c) { return (unsigned long long)c | ((unsigned long long)(c != 0) << 32); }
__global__ void __launch_bounds__(1024) k_scan_local(const uint32_t* __restrict__ count, unsigned long long* __restrict__ block_sums, uint32_t n) {
    __shared__ unsigned long long s_warp[33];
    const uint32_t base = blockIdx.x * SLB_SCAN_BLOCK + threadIdx.x * SLB_SCAN_ITEMS;
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < SLB_SCAN_ITEMS; ++i) sum += pack_count((base + i < n) ? count[base + i] : 0u);
    unsigned long long total;
    block_exclusive_scan(sum, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan_sums(unsigned long long* __restrict__ block_sums, uint32_t n_blocks, uint32_t* __restrict__ totals,
                                                    const uint32_t* __restrict__ counters) {
    __shared__ unsigned long long s_warp[33];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t at = 0; at < n_blocks; at += 1024) {
        uint32_t i = at + threadIdx.x;
        unsigned long long v = i < n_blocks ? block_sums[i] : 0ull;
        unsigned long long total;
        unsigned long long ex = block_exclusive_scan(v, s_warp, total);
        if (i < n_blocks) block_sums[i] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {   // pairs, active tiles, survivors, survivor overflow flag -> mapped host memory
        totals[0] = (uint32_t)s_carry; totals[1] = (uint32_t)(s_carry >> 32); totals[2] = counters[0]; totals[3] = counters[1]; totals[4] = counters[2];
    }
}
__global__ void __launch_bounds__(1024) k_scan_fix(uint32_t* __restrict__ count, const unsigned long long* __restrict__ block_sums,
                                                   uint32_t* __restrict__ off, ActiveTile* __restrict__ active, uint32_t n) {
    __shared__ unsigned long long s_warp[33];
    const uint32_t base = blockIdx.x * SLB_SCAN_BLOCK + threadIdx.x * SLB_SCAN_ITEMS;
    uint32_t c[SLB_SCAN_ITEMS];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < SLB_SCAN_ITEMS; ++i) { c[i] = (base + i < n) ? count[base + i] : 0u; sum += pack_count(c[i]); }
    unsigned long long total;
    unsigned long long run = block_sums[blockIdx.x] + block_exclusive_scan(sum, s_warp, total);
#pragma unroll
    for (int i = 0; i < SLB_SCAN_ITEMS; ++i) {
        if (base + i < n) {
            off[base + i] = (uint32_t)run;
            if (c[i]) {
                ActiveTile a; a.tile = base + i; a.beg = (uint32_t)run; a.count = c[i]; a.pad = 0; active[(uint32_t)(run >> 32)] = a;
                count[base + i] = (uint32_t)run + c[i];   // becomes the emit pass's cursor (end of the segment)
            }
        }
        run += pack_count(c[i]);
    }
    if (base <= n && n < base + SLB_SCAN_ITEMS) off[n] = (uint32_t)(run);   // run == total here only for the thread covering index n
}

// ---------------------------------------------------------------------------------------------
// fine raster
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

#define SLB_RASTER_WARPS 8
#define SLB_RASTER_CHUNK 32
#define SLB_RASTER_SMALL 24   // sub-triangles whose in-tile pixel box is at most this big are rasterised by ONE lane

// fragment-stage discards that decide coverage: depth peel + alpha test (render_shader.frag:229-246)
__device__ __noinline__ bool frag_discard(const DFrame& f, const DDraw& d, uint32_t seq, int k /* low key byte */, int px, int py) {
    const uint32_t tri = seq - d.prim_base;
    const uint32_t* ip = d.idx + 3 * (size_t)tri;
    const uint32_t vi[3] = {__ldg(ip), __ldg(ip + 1), __ldg(ip + 2)};
    PolyV a, b, c;
    float4 pm[3];
    const int how = fetch_subtri(f, d, seq, k, vi, pm, a, b, c);
    if (!how) return true;
    SubTri st;
    if (!make_subtri(a.X, a.Y, b.X, b.Y, c.X, c.Y, a.z, b.z, c.z, st)) return true;
    FragIn in; float bary[3]; uint32_t vid[3];
    shade_inputs(f, d, st, a, b, c, vi, pm, how == 1, px, py, d.tex[0] != nullptr, in, bary, vid);
    if (f.peel && in.objc.w - 0.00001f <= f.peel[((size_t)py * f.W + px) * 4 + 3]) return true;
    if ((d.flags & DRAW_FRAG_TEST) && base_color(d, in).w < 0.5f) return true;
    return false;
}

// One warp per 8x8 tile. The tile's PairRec run is staged through TMA bulk copies (double buffered, 32
// records per stage). Each lane then owns one record of the stage: sub-triangles whose pixel box inside the
// tile is small (the common case for 16k-triangle meshes) are walked by that lane alone, with per-lane
// z-compare through a 64-bit atomicMin on the tile's key buffer in shared memory; the remaining (large)
// records are found with a warp ballot and processed by the whole warp, two pixels per lane.
template <bool FRAG>
__global__ void __launch_bounds__(SLB_RASTER_WARPS * 32) k_raster(const DView* __restrict__ views, const DFrame* __restrict__ frames,
                                                                   const DDraw* __restrict__ draws, const ActiveTile* __restrict__ active,
                                                                   const PairRec* __restrict__ pairs, const RasterGrid g) {
    __shared__ __align__(128) PairRec s_rec[SLB_RASTER_WARPS][2][SLB_RASTER_CHUNK];
    __shared__ unsigned long long s_key[SLB_RASTER_WARPS][SLB_TILE * SLB_TILE];
    __shared__ __align__(8) uint64_t s_bar[SLB_RASTER_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t at = blockIdx.x * SLB_RASTER_WARPS + warp;
    if (at >= g.n_active) return;
    const ActiveTile act = active[at];   // only non-empty tiles get a warp; the outputs were pre-filled with "empty"
    const uint32_t t = act.tile;
    uint32_t vi, tl;   // camera views come first in the tile index space, then the shadow views
    if (t < g.n_cam_tiles) { vi = t / g.tiles_per_cam; tl = t - vi * g.tiles_per_cam; }
    else { const uint32_t u = t - g.n_cam_tiles; vi = u / g.tiles_per_shadow; tl = u - vi * g.tiles_per_shadow; vi += g.n_cam_views; }
    const DView& v = views[vi];
    const DFrame& f = frames[v.frame];
    const int W = v.W, H = v.H;
    const int tx = tl % v.tiles_x, ty = tl / v.tiles_x;
    const int x_lo = tx * SLB_TILE, y_lo = ty * SLB_TILE;
    const int x_hi = min(x_lo + SLB_TILE - 1, W - 1), y_hi = min(y_lo + SLB_TILE - 1, H - 1);
    const int lx = lane & 7, ly = lane >> 3;
    const uint32_t beg = act.beg, n = act.count;
    unsigned long long* keys = s_key[warp];
    {   // start from what the setup kernel's direct path already merged into this tile of the output
        const int gx = x_lo + lx;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gy = y_lo + ly + 4 * h;
            unsigned long long k0 = SLB_KEY_EMPTY;
            if (gx < W && gy < H) {
                if (v.shadow) {
                    const uint32_t g0 = reinterpret_cast<const uint32_t*>(v.out)[(size_t)gy * W + gx];
                    if ((g0 & 0xFF000000u) == v.tagbits) k0 = (unsigned long long)(g0 & 0xFFFFFFu) << 40;   // written in this generation
                } else {
                    k0 = reinterpret_cast<const unsigned long long*>(v.out)[(size_t)gy * W + gx];
                }
            }
            keys[lane + 32 * h] = k0;
        }
    }

    if (n > 0) {
        uint64_t* bar = s_bar[warp];
        if (lane == 0) {
            mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        const uint32_t nchunks = (n + SLB_RASTER_CHUNK - 1) / SLB_RASTER_CHUNK;
        if (lane == 0) {
            uint32_t cnt = min(n, (uint32_t)SLB_RASTER_CHUNK);
            mbar_expect_tx(&bar[0], cnt * (uint32_t)sizeof(PairRec));
            bulk_g2s(&s_rec[warp][0][0], pairs + beg, cnt * (uint32_t)sizeof(PairRec), &bar[0]);
        }
        for (uint32_t c = 0; c < nchunks; ++c) {
            const int b = c & 1;
            if (lane == 0 && c + 1 < nchunks) {   // prefetch the next stage into the other buffer
                uint32_t cnt = min(n - (c + 1) * SLB_RASTER_CHUNK, (uint32_t)SLB_RASTER_CHUNK);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bar[b ^ 1], cnt * (uint32_t)sizeof(PairRec));
                bulk_g2s(&s_rec[warp][b ^ 1][0], pairs + beg + (size_t)(c + 1) * SLB_RASTER_CHUNK, cnt * (uint32_t)sizeof(PairRec),
                         &bar[b ^ 1]);
            }
            mbar_wait(&bar[b], (c >> 1) & 1);
            const int cnt = (int)min(n - c * SLB_RASTER_CHUNK, (uint32_t)SLB_RASTER_CHUNK);
            // ---- one record per lane ----
            bool big = false;
            // edge functions (orientation-normalised, top-left bias folded in) at the tile's first pixel centre and
            // their per-pixel steps; depth plane; low key bits — kept in registers for the warp-wide path below
            long long E0 = 0, E1 = 0, E2 = 0;
            int sx0 = 0, sy0 = 0, sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
            float z_a = 0.f, z_b = 0.f, z_c = 0.f, inv2A = 0.f;
            int sb = 0;
            uint32_t r_seq = 0, r_kf = 0, r_draw = 0;
            if (lane < cnt) {
                const PairRec r = s_rec[warp][b][lane];
                r_seq = r.seq; r_kf = r.k_flags; r_draw = r.draw;
                const int xmin = min(r.ax, min(r.bx, r.cx)), xmax = max(r.ax, max(r.bx, r.cx));
                const int ymin = min(r.ay, min(r.by, r.cy)), ymax = max(r.ay, max(r.by, r.cy));
                const int px0 = max(x_lo, (xmin - 128 + 255) >> 8), px1 = min(x_hi, (xmax - 128) >> 8);
                const int py0 = max(y_lo, (ymin - 128 + 255) >> 8), py1 = min(y_hi, (ymax - 128) >> 8);
                if (px0 <= px1 && py0 <= py1) {
                    const long long twoA = edge_fn(r.ax, r.ay, r.bx, r.by, r.cx, r.cy);
                    const int sg = twoA > 0 ? 1 : -1;
                    inv2A = __frcp_rn(__ll2float_rn(twoA));
                    const int bias0 = top_left(r.cx - r.bx, r.cy - r.by, sg) ? 0 : -1;
                    const int bias1 = top_left(r.ax - r.cx, r.ay - r.cy, sg) ? 0 : -1;
                    const int bias2 = top_left(r.bx - r.ax, r.by - r.ay, sg) ? 0 : -1;
                    sx0 = -sg * (r.cy - r.by); sy0 = sg * (r.cx - r.bx);     // d(s*w_i)/dpx, d(s*w_i)/dpy in units of 1/256 px
                    sx1 = -sg * (r.ay - r.cy); sy1 = sg * (r.ax - r.cx);
                    sx2 = -sg * (r.by - r.ay); sy2 = sg * (r.bx - r.ax);
                    z_a = r.az; z_b = __fsub_rn(r.bz, r.az); z_c = __fsub_rn(r.cz, r.az);
                    sb = (sg < 0 ? 1 : 0) | (bias1 ? 2 : 0) | (bias2 ? 4 : 0);
                    const unsigned long long lowkey = ((unsigned long long)r.seq << 8) | (r.k_flags & 0xffu);
                    if ((px1 - px0 + 1) * (py1 - py0 + 1) > SLB_RASTER_SMALL) {
                        big = true;
                        const int ox = x_lo * 256 + 128, oy = y_lo * 256 + 128;
                        E0 = sg * edge_fn(r.bx, r.by, r.cx, r.cy, ox, oy) + bias0;
                        E1 = sg * edge_fn(r.cx, r.cy, r.ax, r.ay, ox, oy) + bias1;
                        E2 = sg * edge_fn(r.ax, r.ay, r.bx, r.by, ox, oy) + bias2;
                    } else {
                        const int cx0 = px0 * 256 + 128, cy0 = py0 * 256 + 128;
                        const long long a0 = sg * edge_fn(r.bx, r.by, r.cx, r.cy, cx0, cy0) + bias0;
                        const long long a1 = sg * edge_fn(r.cx, r.cy, r.ax, r.ay, cx0, cy0) + bias1;
                        const long long a2 = sg * edge_fn(r.ax, r.ay, r.bx, r.by, cx0, cy0) + bias2;
                        auto plot = [&](long long w1, long long w2, int x, int y) {   // w_i = s*w_i + bias_i at a covered pixel
                            w1 -= bias1; w2 -= bias2;
                            if (sg < 0) { w1 = -w1; w2 = -w2; }
                            float q1 = __fmul_rn(__ll2float_rn(w1), inv2A), q2 = __fmul_rn(__ll2float_rn(w2), inv2A);
                            float z = __fmaf_rn(q2, z_c, __fmaf_rn(q1, z_b, z_a));
                            z = fminf(fmaxf(z, 0.0f), 1.0f);
                            const unsigned long long key = ((unsigned long long)__float2uint_rn(__fmul_rn(z, 16777215.0f)) << 40) | lowkey;
                            unsigned long long* slot = keys + (y - y_lo) * SLB_TILE + (x - x_lo);
                            if (key >= *(volatile unsigned long long*)slot) return;
                            if (FRAG) {
                                if ((r.k_flags & 0x100u) && frag_discard(f, draws[r.draw], r.seq, r.k_flags & 0xff, x, y)) return;
                            }
                            atomicMin(slot, key);
                        };
                        if (xmax - xmin < 16384 && ymax - ymin < 16384) {
                            // the whole triangle is smaller than 64 px: every edge value inside its box fits 32 bits
                            int e0r = (int)a0, e1r = (int)a1, e2r = (int)a2;
                            const int dx0 = sx0 * 256, dy0 = sy0 * 256, dx1 = sx1 * 256, dy1 = sy1 * 256, dx2 = sx2 * 256, dy2 = sy2 * 256;
                            for (int y = py0; y <= py1; ++y, e0r += dy0, e1r += dy1, e2r += dy2) {
                                int e0 = e0r, e1 = e1r, e2 = e2r;
                                for (int x = px0; x <= px1; ++x, e0 += dx0, e1 += dx1, e2 += dx2)
                                    if ((e0 | e1 | e2) >= 0) plot((long long)e1, (long long)e2, x, y);
                            }
                        } else {
                            const long long dx0 = (long long)sx0 * 256, dy0 = (long long)sy0 * 256, dx1 = (long long)sx1 * 256,
                                            dy1 = (long long)sy1 * 256, dx2 = (long long)sx2 * 256, dy2 = (long long)sy2 * 256;
                            long long e0r = a0, e1r = a1, e2r = a2;
                            for (int y = py0; y <= py1; ++y, e0r += dy0, e1r += dy1, e2r += dy2) {
                                long long e0 = e0r, e1 = e1r, e2 = e2r;
                                for (int x = px0; x <= px1; ++x, e0 += dx0, e1 += dx1, e2 += dx2)
                                    if ((e0 | e1 | e2) >= 0) plot(e1, e2, x, y);
                            }
                        }
                    }
                }
            }
            unsigned bigmask = __ballot_sync(0xffffffffu, big);
            __syncwarp();
            // ---- large records: the owning lane broadcasts its set-up, every lane tests its two pixels ----
            while (bigmask) {
                const int j = __ffs(bigmask) - 1;
                bigmask &= bigmask - 1;
                const long long b0 = __shfl_sync(0xffffffffu, E0, j), b1 = __shfl_sync(0xffffffffu, E1, j), b2 = __shfl_sync(0xffffffffu, E2, j);
                const int tx0 = __shfl_sync(0xffffffffu, sx0, j), ty0 = __shfl_sync(0xffffffffu, sy0, j);
                const int tx1 = __shfl_sync(0xffffffffu, sx1, j), ty1 = __shfl_sync(0xffffffffu, sy1, j);
                const int tx2 = __shfl_sync(0xffffffffu, sx2, j), ty2 = __shfl_sync(0xffffffffu, sy2, j);
                const float za = __shfl_sync(0xffffffffu, z_a, j), zb = __shfl_sync(0xffffffffu, z_b, j), zc = __shfl_sync(0xffffffffu, z_c, j);
                const float i2a = __shfl_sync(0xffffffffu, inv2A, j);
                const int sbj = __shfl_sync(0xffffffffu, sb, j);
                const uint32_t seqj = __shfl_sync(0xffffffffu, r_seq, j), kfj = __shfl_sync(0xffffffffu, r_kf, j);
                const uint32_t drawj = __shfl_sync(0xffffffffu, r_draw, j);
                const unsigned long long lowkey = ((unsigned long long)seqj << 8) | (kfj & 0xffu);
                const long long e0 = b0 + (long long)(tx0 * lx + ty0 * ly) * 256, e1 = b1 + (long long)(tx1 * lx + ty1 * ly) * 256,
                                e2 = b2 + (long long)(tx2 * lx + ty2 * ly) * 256;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long g0 = e0 + (long long)ty0 * (1024 * h), g1 = e1 + (long long)ty1 * (1024 * h), g2 = e2 + (long long)ty2 * (1024 * h);
                    const int x = x_lo + lx, y = y_lo + ly + 4 * h;
                    if ((g0 | g1 | g2) < 0 || x > x_hi || y > y_hi) continue;
                    long long w1 = g1 - ((sbj & 2) ? -1 : 0), w2 = g2 - ((sbj & 4) ? -1 : 0);
                    if (sbj & 1) { w1 = -w1; w2 = -w2; }
                    float q1 = __fmul_rn(__ll2float_rn(w1), i2a), q2 = __fmul_rn(__ll2float_rn(w2), i2a);
                    float z = __fmaf_rn(q2, zc, __fmaf_rn(q1, zb, za));
                    z = fminf(fmaxf(z, 0.0f), 1.0f);
                    const unsigned long long key = ((unsigned long long)__float2uint_rn(__fmul_rn(z, 16777215.0f)) << 40) | lowkey;
                    unsigned long long* slot = keys + (ly + 4 * h) * SLB_TILE + lx;
                    if (key >= *slot) continue;
                    if (FRAG) {
                        if ((kfj & 0x100u) && frag_discard(f, draws[drawj], seqj, kfj & 0xff, x, y)) continue;
                    }
                    *slot = key;
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    const int px = x_lo + lx;
    if (px < W) {
        if (v.shadow) {   // depth-only view: the d24 plane the PCF lookup reads (cleared to 0xFFFFFF where nothing was drawn)
            uint32_t* out = reinterpret_cast<uint32_t*>(v.out);
            const unsigned long long k0 = keys[ly * SLB_TILE + lx], k1 = keys[(ly + 4) * SLB_TILE + lx];
            static_assert(SLB_TILE == 8 && SLB_SHADOW_MASK_SHIFT >= 3, "a raster tile lies inside one block of the shadow-map occupancy mask");
            if (v.mask) {   // one atomic per tile, by the first active lane
                const unsigned act = __activemask();
                const unsigned any = __ballot_sync(act, k0 != SLB_KEY_EMPTY || k1 != SLB_KEY_EMPTY);
                const int bx = tx >> (SLB_SHADOW_MASK_SHIFT - 3), by = ty >> (SLB_SHADOW_MASK_SHIFT - 3);
                if (any && lane == __ffs(act) - 1) atomicOr(v.mask + by * SLB_SHADOW_MASK_ROW + (bx >> 5), 1u << (bx & 31));
            }
            if (y_lo + ly < H) out[(size_t)(y_lo + ly) * W + px] = v.tagbits | (k0 == SLB_KEY_EMPTY ? 0xFFFFFFu : (uint32_t)(k0 >> 40));
            if (y_lo + ly + 4 < H) out[(size_t)(y_lo + ly + 4) * W + px] = v.tagbits | (k1 == SLB_KEY_EMPTY ? 0xFFFFFFu : (uint32_t)(k1 >> 40));
        } else {
            unsigned long long* out = reinterpret_cast<unsigned long long*>(v.out);
            if (y_lo + ly < H) out[(size_t)(y_lo + ly) * W + px] = keys[ly * SLB_TILE + lx];
            if (y_lo + ly + 4 < H) out[(size_t)(y_lo + ly + 4) * W + px] = keys[(ly + 4) * SLB_TILE + lx];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shade + multi-render-target store
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_draw_by_prim(const DDraw* __restrict__ draws, uint32_t lo, uint32_t hi, uint32_t seq) {
    --hi;   // last draw in [lo, hi] with prim_base <= seq
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (draws[mid].prim_base <= seq) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Once per frame and huge sub-triangle (instead of once per pixel of it): the per-pixel re-set-up of k_shade's generic path —
// index / vertex fetch, clip-space transform or ClipRec read, the vertex stage of the three original vertices
// (render_shader.vert:57-95) — with the results parked in a HugeShade record for the pixels to interpolate.
__global__ void __launch_bounds__(32) k_huge_prepare(const DFrame* __restrict__ frames, const DDraw* __restrict__ draws) {
    const DFrame& f = frames[blockIdx.x];
    if (!f.huge || !f.huge_shade) return;
    const int nh = min((int)__ldg(f.huge_n), SLB_HUGE_PER_VIEW);
    if ((int)threadIdx.x >= nh) return;
    const HugeRec h = f.huge[threadIdx.x];
    HugeShade hs;
    hs.fast = 0u;
    const uint32_t seq = h.seq;
    const uint32_t di = f.seq_shift ? f.draw_begin + (seq >> f.seq_shift) : find_draw_by_prim(draws, f.draw_begin, f.draw_end, seq);
    const DDraw& d = draws[di];
    const uint32_t tri = seq - d.prim_base;
    const uint32_t* ip = d.idx + 3 * (size_t)tri;
    const uint32_t vi[3] = {__ldg(ip), __ldg(ip + 1), __ldg(ip + 2)};
    PolyV va, vb, vc; float4 pm[3];
    const int how = fetch_subtri(f, d, seq, (int)h.kbyte, vi, pm, va, vb, vc);
    if (how && (d.flags & DRAW_AFFINE) && !d.sticker) {
        hs.fast = 1u;
        hs.invw[0] = va.invw; hs.invw[1] = vb.invw; hs.invw[2] = vc.invw;
        for (int j = 0; j < 3; ++j) { hs.basis[0][j] = va.b[j]; hs.basis[1][j] = vb.b[j]; hs.basis[2][j] = vc.b[j]; }
        hs.unit_basis = how == 1 ? 1u : 0u;
        hs.draw = di;
        hs.front = h.s < 0 ? 1u : 0u;   // FrontFace = CW (render_pass.cpp:330): twoA < 0
        for (int j = 0; j < 3; ++j) {
            VSOut o; uint32_t id;
            vertex_stage(f, d, vi[j], o, id);
            hs.vid[j] = id; hs.vi[j] = vi[j];
            hs.objc[j][0] = o.objc.x; hs.objc[j][1] = o.objc.y; hs.objc[j][2] = o.objc.z;
            hs.wc[j][0] = o.wc.x; hs.wc[j][1] = o.wc.y; hs.wc[j][2] = o.wc.z;
            hs.cc[j][0] = o.cc.x; hs.cc[j][1] = o.cc.y; hs.cc[j][2] = o.cc.z;
            hs.nW[j][0] = o.nW.x; hs.nW[j][1] = o.nW.y; hs.nW[j][2] = o.nW.z;
            hs.uv[j][0] = o.u; hs.uv[j][1] = o.v;
        }
    }
    f.huge_shade[threadIdx.x] = hs;
}

template <int THREADS, int MINB, bool LEAN>
__global__ void __launch_bounds__(THREADS, MINB) k_shade(const DFrame* __restrict__ frames, const DDraw* __restrict__ draws) {
    const DFrame& f = frames[blockIdx.z];
    const int W = f.W, H = f.H;
    // A warp covers an 8x4 pixel block, not a 32x1 strip: fewer draws / texture rows / shadow texel rows per warp
    // (measured 29.3 vs 32.2 ms per 1024-frame step). Float targets still store full 128-byte lines per row segment.
#ifndef SLB_WARP_W
#define SLB_WARP_W 8
#endif
    static_assert(THREADS == 256 || THREADS == 128, "the warp layout assumes 32x8 or 32x4 pixel blocks");
    constexpr int BH = THREADS / 32;   // block height in pixels (block width is 32)
    constexpr int WW = SLB_WARP_W, WH = 32 / WW, WPR = 32 / WW;   // warp block width / height, warps per block row
    const int wq = threadIdx.x >> 5, lq = threadIdx.x & 31;
    const int px = blockIdx.x * 32 + (wq % WPR) * WW + (lq % WW), py = blockIdx.y * BH + (wq / WPR) * WH + (lq / WW);
    // huge sub-triangles of this view (resolved per pixel below): staged once per block, only those whose pixel box
    // meets the block's 32x8 pixels
    __shared__ HugeRec s_huge[SLB_HUGE_PER_VIEW];
    __shared__ int s_src[SLB_HUGE_PER_VIEW];
    __shared__ int s_nh;
    // the pixel's visibility key is requested BEFORE the block stages its huge records: its latency overlaps the two barriers
    const bool in_frame = px < W && py < H;
    const size_t p = (size_t)py * W + px;
    unsigned long long key = in_frame ? f.keys[p] : SLB_KEY_EMPTY;
    if (threadIdx.x == 0) s_nh = 0;
    __syncthreads();
    if (f.huge && threadIdx.x < SLB_HUGE_PER_VIEW && (int)threadIdx.x < min((int)__ldg(f.huge_n), SLB_HUGE_PER_VIEW)) {
        const HugeRec h = f.huge[threadIdx.x];
        const int bx0 = blockIdx.x * 32, by0 = blockIdx.y * BH;
        if (h.px1 >= bx0 && h.px0 <= bx0 + 31 && h.py1 >= by0 && h.py0 <= by0 + BH - 1) { const int at = atomicAdd(&s_nh, 1); s_huge[at] = h; s_src[at] = threadIdx.x; }
    }
    __syncthreads();
    if (!in_frame) return;
    int best = -1;                       // the staged huge record that owns this pixel, if one does
    long long bw0 = 0, bw1 = 0, bw2 = 0;   // its edge-function values here (reused for the barycentrics)
    {   // coverage (C6) and depth (C7) of the staged huge sub-triangles at this pixel, merged by minimum (every covering
        // record is evaluated, exactly as the tiled path would: snapped fan triangles may overlap in degenerate cases)
        const int nh = s_nh;
        for (int i = 0; i < nh; ++i) {
            const HugeRec& h = s_huge[i];
            if (px < h.px0 || px > h.px1 || py < h.py0 || py > h.py1) continue;
            SubTri st;
            st.ax = h.ax; st.ay = h.ay; st.bx = h.bx; st.by = h.by; st.cx = h.cx; st.cy = h.cy;
            st.az = h.az; st.bz = h.bz; st.cz = h.cz; st.s = h.s; st.inv2A = h.inv2A;
            st.bias0 = h.bias0; st.bias1 = h.bias1; st.bias2 = h.bias2; st.twoA = h.s;
            long long w0, w1, w2;
            subtri_weights(st, px, py, w0, w1, w2);
            if (!subtri_covers(st, w0, w1, w2)) continue;
            const unsigned long long cand = ((unsigned long long)subtri_depth24(st, w1, w2) << 40) | ((unsigned long long)h.seq << 8) | h.kbyte;
            if (cand < key) { key = cand; best = i; bw0 = w0; bw1 = w1; bw2 = w2; }
        }
        if (nh && !f.fused_tonemap) f.keys[p] = key;   // the post passes (sky box / background image) test coverage on the keys
    }
    // the target pointers are fetched where they are used (not held in registers across the set-up code)
    auto store_geometry = [&](float4 coord, unsigned short cls, unsigned short inst, uint4 vidx, float4 bary4, float4 cam) {
        if (float4* o = reinterpret_cast<float4*>(f.out[SLB_TARGET_COORD])) o[p] = coord;
        if (unsigned short* o = reinterpret_cast<unsigned short*>(f.out[SLB_TARGET_CLASS])) o[p] = cls;
        if (unsigned short* o = reinterpret_cast<unsigned short*>(f.out[SLB_TARGET_INSTANCE])) o[p] = inst;
        if (uint4* o = reinterpret_cast<uint4*>(f.out[SLB_TARGET_VERTEX_INDEX])) o[p] = vidx;
        if (float4* o = reinterpret_cast<float4*>(f.out[SLB_TARGET_BARY])) o[p] = bary4;
        if (f.scratch_cam) f.scratch_cam[p] = cam;   // == out[CAM_COORD] when requested
        if (f.zplane) f.zplane[p] = cam.z;
    };

    float4 hdr = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
    bool shaded = false;
    if (key != SLB_KEY_EMPTY) {
        FragIn in; float bary[3]; uint32_t vid[3], vi[3];
        const DDraw* dp = nullptr;
        // (the record is read through the read-only path: every lane of a warp that sits on the plane reads the same 256 bytes,
        // which stay in L1 for the frame — no staging copy, no block-wide barrier before the first pixel can start)
        const HugeShade* hsp = (best >= 0 && f.huge_shade) ? f.huge_shade + s_src[best] : nullptr;
        if (hsp && __ldg(&hsp->fast)) {
            // huge sub-triangle: interpolate the vertex-stage outputs k_huge_prepare parked for it; the barycentrics come from
            // the edge-function values the coverage test above already has (same arithmetic as bary_from_weights)
            const HugeShade& hs = *hsp;      // fields are fetched where they are used
            const HugeRec& h = s_huge[best];
            dp = &draws[hs.draw];
            auto bary_at = [&](long long w0, long long w1, long long w2, float out[3]) {
                const float b0 = __ll2float_rn(w0) * h.inv2A, b1 = __ll2float_rn(w1) * h.inv2A, b2 = __ll2float_rn(w2) * h.inv2A;
                const float g0 = b0 * hs.invw[0], g1 = b1 * hs.invw[1], g2 = b2 * hs.invw[2];
                const float sgm = g0 + g1 + g2;
                const float q0 = g0 / sgm, q1 = g1 / sgm, q2 = g2 / sgm;
                if (hs.unit_basis) { out[0] = q0; out[1] = q1; out[2] = q2; return; }
#pragma unroll
                for (int j = 0; j < 3; ++j) out[j] = q0 * hs.basis[0][j] + q1 * hs.basis[1][j] + q2 * hs.basis[2][j];
            };
            bary_at(bw0, bw1, bw2, bary);
            auto lerp = [&](const float a[3][3], int c) { return a[0][c] * bary[0] + a[1][c] * bary[1] + a[2][c] * bary[2]; };
            in.objc = make_float4(lerp(hs.objc, 0), lerp(hs.objc, 1), lerp(hs.objc, 2), 0.f);
            in.wc = mk3(lerp(hs.wc, 0), lerp(hs.wc, 1), lerp(hs.wc, 2));
            in.cc = mk3(lerp(hs.cc, 0), lerp(hs.cc, 1), lerp(hs.cc, 2));
            in.objc.w = in.cc.z;
            in.nW = mk3(lerp(hs.nW, 0), lerp(hs.nW, 1), lerp(hs.nW, 2));
            in.u = hs.uv[0][0] * bary[0] + hs.uv[1][0] * bary[1] + hs.uv[2][0] * bary[2];
            in.v = hs.uv[0][1] * bary[0] + hs.uv[1][1] * bary[1] + hs.uv[2][1] * bary[2];
            in.su = in.sv = -1.0f;
            in.front = hs.front != 0u;
            in.u_dx = in.v_dx = in.u_dy = in.v_dy = 0.0f;
            if (LEAN ? dp->tex[0] != nullptr : draw_has_textures(*dp)) {   // dFdx / dFdy of uv: the quad partner's values on the same primitive
                const long long sx = (px & 1) ? -256 : 256, sy = (py & 1) ? -256 : 256;
                float bx[3], by[3];
                bary_at(bw0 - sx * (h.cy - h.by), bw1 - sx * (h.ay - h.cy), bw2 - sx * (h.by - h.ay), bx);
                bary_at(bw0 + sy * (h.cx - h.bx), bw1 + sy * (h.ax - h.cx), bw2 + sy * (h.bx - h.ax), by);
                const float sgx = (px & 1) ? -1.0f : 1.0f, sgy = (py & 1) ? -1.0f : 1.0f;
                in.u_dx = sgx * (hs.uv[0][0] * bx[0] + hs.uv[1][0] * bx[1] + hs.uv[2][0] * bx[2] - in.u);
                in.v_dx = sgx * (hs.uv[0][1] * bx[0] + hs.uv[1][1] * bx[1] + hs.uv[2][1] * bx[2] - in.v);
                in.u_dy = sgy * (hs.uv[0][0] * by[0] + hs.uv[1][0] * by[1] + hs.uv[2][0] * by[2] - in.u);
                in.v_dy = sgy * (hs.uv[0][1] * by[0] + hs.uv[1][1] * by[1] + hs.uv[2][1] * by[2] - in.v);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) { vid[j] = hs.vid[j]; vi[j] = hs.vi[j]; }
        } else {
            const uint32_t seq = (uint32_t)(key >> 8);
            const int k = (int)(key & 0xffu);
            // the draw is encoded in the sequence number when the frame's draws fit (host: build_batch), else searched
            const DDraw& d = draws[f.seq_shift ? f.draw_begin + (seq >> f.seq_shift) : find_draw_by_prim(draws, f.draw_begin, f.draw_end, seq)];
            const uint32_t tri = seq - d.prim_base;
            const uint32_t* ip = d.idx + 3 * (size_t)tri;
            vi[0] = __ldg(ip); vi[1] = __ldg(ip + 1); vi[2] = __ldg(ip + 2);
            PolyV va, vb, vc;
            SubTri st;
            float4 pm[3];
            const int how = fetch_subtri(f, d, seq, k, vi, pm, va, vb, vc);
            if (how && make_subtri(va.X, va.Y, vb.X, vb.Y, vc.X, vc.Y, va.z, vb.z, vc.z, st)) {
                shade_inputs<LEAN>(f, d, st, va, vb, vc, vi, pm, how == 1, px, py, LEAN ? d.tex[0] != nullptr : draw_has_textures(d), in, bary, vid);
                dp = &d;
            }
        }
        if (dp) {
            const DDraw& d = *dp;
            // geometry targets first: their registers are free before the lighting code runs
            store_geometry(in.objc, (unsigned short)d.class_index, (unsigned short)d.instance_index, make_uint4(vid[0], vid[1], vid[2], 0u),
                           make_float4(bary[0], bary[1], bary[2], 1.0f), make_float4(in.cc.x, in.cc.y, in.cc.z, 1.0f));
            fragment_stage<LEAN>(f, d, in, vi, bary, hdr, nrm);
            shaded = true;
        }
    }
    if (!shaded) {   // clear values (render_pass.cpp:316,523-532)
        const float4 inval = make_float4(SLB_INVALID_COORD, SLB_INVALID_COORD, SLB_INVALID_COORD, SLB_INVALID_COORD);
        store_geometry(inval, 0, 0, make_uint4(0u, 0u, 0u, 0u), make_float4(0.f, 0.f, 0.f, 1.0f), inval);
    }
    if (f.fused_tonemap) {
        if (f.out[SLB_TARGET_RGB]) reinterpret_cast<uchar4*>(f.out[SLB_TARGET_RGB])[p] = tone_map(hdr, f.manual_exposure, nullptr);
    } else {
        f.hdr[p] = hdr;
    }
    if (f.scratch_normal) f.scratch_normal[p] = nrm;   // == out[NORMAL] when that target is requested
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace slbk {

void launch_setup(const DView* views, const DFrame* frames, const DBinDraw* bdraws, const uint32_t* chunk_draw, uint32_t n_chunks,
                  uint32_t* tile_count, PairRec* survivors, uint32_t* counters, uint32_t normal_cap, uint32_t huge_cap, int direct_max,
                  int warp_max, cudaStream_t s) {
    if (n_chunks == 0) return;
    SurvOut so;
    so.survivors = survivors; so.counters = counters; so.normal_cap = normal_cap; so.huge_cap = huge_cap;
    k_setup<<<n_chunks, SLB_SETUP_CHUNK, 0, s>>>(views, frames, bdraws, chunk_draw, tile_count, so, direct_max, max(direct_max, warp_max));
}
void launch_emit(const DView* views, const PairRec* survivors, uint32_t n_survivors, const PairRec* huge, uint32_t n_huge,
                 uint32_t* tile_count, PairRec* pairs, uint32_t capacity, cudaStream_t s) {
    if (n_survivors)
        k_emit<<<(n_survivors + SLB_SETUP_CHUNK - 1) / SLB_SETUP_CHUNK, SLB_SETUP_CHUNK, 0, s>>>(views, survivors, n_survivors, tile_count, pairs,
                                                                                               capacity);
    if (n_huge) k_emit_huge<<<n_huge, SLB_HUGE_THREADS, 0, s>>>(views, huge, tile_count, pairs, capacity);
}
void launch_scan(uint32_t* count, uint32_t* off, ActiveTile* active, unsigned long long* block_sums, uint32_t* totals,
                 const uint32_t* counters, uint32_t n, cudaStream_t s) {
    const uint32_t n_blocks = (n + 1 + SLB_SCAN_BLOCK - 1) / SLB_SCAN_BLOCK;   // n + 1: some thread must own index n (the total)
    k_scan_local<<<n_blocks, 1024, 0, s>>>(count, block_sums, n);
    k_scan_sums<<<1, 1024, 0, s>>>(block_sums, n_blocks, totals, counters);
    k_scan_fix<<<n_blocks, 1024, 0, s>>>(count, block_sums, off, active, n);
}
void launch_raster(bool frag_test, const DView* views, const DFrame* frames, const DDraw* draws, const ActiveTile* active,
                   const PairRec* pairs, RasterGrid g, cudaStream_t s) {
    if (g.n_active == 0) return;
    const unsigned grid = (g.n_active + SLB_RASTER_WARPS - 1) / SLB_RASTER_WARPS;
    if (frag_test) k_raster<true><<<grid, SLB_RASTER_WARPS * 32, 0, s>>>(views, frames, draws, active, pairs, g);
    else k_raster<false><<<grid, SLB_RASTER_WARPS * 32, 0, s>>>(views, frames, draws, active, pairs, g);
}
#ifndef SLB_SHADE_MINB
#define SLB_SHADE_MINB 4
#endif
#ifndef SLB_SHADE_MINB_FULL
#define SLB_SHADE_MINB_FULL SLB_SHADE_MINB
#endif
void launch_shade(const DFrame* frames, const DDraw* draws, int n_frames, int W, int H, bool lean, cudaStream_t s) {
    // 256 threads = 32 x 8 pixels. Launch bounds measured on B200 (ms per 1024-frame step): (256,2) 118 regs 43.8,
    // (128,5) 96 regs 39.4, (256,3) 80 regs 36.3, (256,4) 64 regs 33.4 — the kernel is latency bound, occupancy wins
    // even with ~150 B of spills.
    // `lean`: no material textures beyond base colour, no stickers, no light map, affine chains only (see fragment_stage)
    k_huge_prepare<<<n_frames, 32, 0, s>>>(frames, draws);   // no-op for frames without huge records
#ifndef SLB_SHADE_THREADS
#define SLB_SHADE_THREADS 256
#endif
    constexpr int T = SLB_SHADE_THREADS, BH = T / 32, MB = SLB_SHADE_MINB * 256 / T, MBF = SLB_SHADE_MINB_FULL * 256 / T;
    if (lean) k_shade<T, MB, true><<<dim3((W + 31) / 32, (H + BH - 1) / BH, n_frames), T, 0, s>>>(frames, draws);
    else k_shade<T, MBF, false><<<dim3((W + 31) / 32, (H + BH - 1) / BH, n_frames), T, 0, s>>>(frames, draws);
}

}  // namespace slbk
