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
#define SLB_BIG_QUEUE 48
#define SLB_BIG_TILES 24   // sub-triangles touching more tiles than this are binned by the whole block

template <bool EMIT>
__device__ __forceinline__ void bin_pair(uint32_t tile, const PairRec& rec, uint32_t* __restrict__ tile_count,
                                         const uint32_t* __restrict__ tile_off, PairRec* __restrict__ pairs, uint32_t capacity) {
    if (EMIT) {
        uint32_t slot = atomicSub(tile_count + tile, 1u) - 1u;   // the count pass left the tile's total here
        uint32_t at = __ldg(tile_off + tile) + slot;
        if (at < capacity) pairs[at] = rec;
    } else {
        atomicAdd(tile_count + tile, 1u);
    }
}

template <bool EMIT>
__global__ void __launch_bounds__(SLB_SETUP_CHUNK, 4) k_bin(const DView* __restrict__ views, const DFrame* __restrict__ frames,
                                                         const DBinDraw* __restrict__ bdraws, const uint32_t* __restrict__ chunk_draw,
                                                         uint32_t* __restrict__ tile_count,
                                                         const uint32_t* __restrict__ tile_off, PairRec* __restrict__ pairs,
                                                         uint32_t capacity) {
    __shared__ float s_mvp[16];
    __shared__ BigEntry s_big[SLB_BIG_QUEUE];
    __shared__ int s_nbig;
    const uint32_t di = __ldg(chunk_draw + blockIdx.x);   // host-built table: setup chunk -> bin draw
    const DBinDraw& d = bdraws[di];
    if (threadIdx.x < 16) s_mvp[threadIdx.x] = d.mvp[threadIdx.x];
    if (threadIdx.x == 32) s_nbig = 0;
    __syncthreads();
    const DView& v = views[d.view];
    const int W = v.W, H = v.H, tiles_x = v.tiles_x;
    const uint32_t tile_base = v.tile_base;
    const bool shadow = v.shadow != 0;
    const uint32_t tri = (blockIdx.x - d.chunk_base) * SLB_SETUP_CHUNK + threadIdx.x;
    if (tri < d.n_tris) {
        const uint32_t* ip = d.idx + 3 * (size_t)tri;
        uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
        float4 p0 = __ldg(d.pos4 + i0), p1 = __ldg(d.pos4 + i1), p2 = __ldg(d.pos4 + i2);
        PrimSetup ps;
        bool clipped = false;
        if (setup_prim(s_mvp, make_float3(p0.x, p0.y, p0.z), make_float3(p1.x, p1.y, p1.z), make_float3(p2.x, p2.y, p2.z), W, H, ps, &clipped)) {
            if (EMIT && clipped && !shadow) {   // publish the clipped polygon once for the fragment test and the shade kernel
                const DFrame& f = frames[v.frame];
                uint32_t slot = atomicAdd(f.clip_count, 1u);
                if (slot < SLB_MAX_CLIP) {
                    ClipRec& cr = f.clip[slot];
                    cr.seq = d.prim_base + tri; cr.n = ps.n;
                    for (int i = 0; i < ps.n; ++i) {
                        cr.v[i].X = ps.v[i].X; cr.v[i].Y = ps.v[i].Y; cr.v[i].z = ps.v[i].z; cr.v[i].invw = ps.v[i].invw;
                        cr.v[i].b[0] = ps.v[i].b[0]; cr.v[i].b[1] = ps.v[i].b[1]; cr.v[i].b[2] = ps.v[i].b[2];
                    }
                }
            }
            for (int k = 1; k + 1 < ps.n; ++k) {
                SubTri st;
                if (!make_subtri(ps, k, st)) continue;
                if (shadow && st.twoA < 0) continue;   // shadow views cull FRONT faces (render_pass.cpp:428-429)
                int px0, py0, px1, py1;
                if (!subtri_pixel_bbox(st, W, H, px0, py0, px1, py1)) continue;
                int tx0 = px0 / SLB_TILE, tx1 = px1 / SLB_TILE, ty0 = py0 / SLB_TILE, ty1 = py1 / SLB_TILE;
                int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
                PairRec rec;
                rec.ax = st.ax; rec.ay = st.ay; rec.bx = st.bx; rec.by = st.by; rec.cx = st.cx; rec.cy = st.cy;
                rec.az = st.az; rec.bz = st.bz; rec.cz = st.cz;
                rec.seq = d.prim_base + tri;
                rec.k_flags = (uint32_t)k | ((d.flags & DRAW_FRAG_TEST) ? 0x100u : 0u);
                rec.draw = d.draw;
                if (ntx * nty == 1) {
                    bin_pair<EMIT>(tile_base + ty0 * tiles_x + tx0, rec, tile_count, tile_off, pairs, capacity);
                } else if (ntx * nty <= SLB_BIG_TILES) {
                    const bool test = ntx * nty > 2;
                    for (int ty = ty0; ty <= ty1; ++ty)
                        for (int tx = tx0; tx <= tx1; ++tx)
                            if (!test || tile_may_overlap(st, tx, ty, W, H))
                                bin_pair<EMIT>(tile_base + ty * tiles_x + tx, rec, tile_count, tile_off, pairs, capacity);
                } else {
                    int q = atomicAdd(&s_nbig, 1);
                    if (q < SLB_BIG_QUEUE) {
                        s_big[q].rec = rec; s_big[q].tx0 = tx0; s_big[q].ty0 = ty0; s_big[q].ntx = ntx; s_big[q].nty = nty;
                    } else {   // queue full: this thread walks the tiles itself
                        for (int ty = ty0; ty <= ty1; ++ty)
                            for (int tx = tx0; tx <= tx1; ++tx)
                                if (tile_may_overlap(st, tx, ty, W, H))
                                    bin_pair<EMIT>(tile_base + ty * tiles_x + tx, rec, tile_count, tile_off, pairs, capacity);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int nbig = min(s_nbig, SLB_BIG_QUEUE);
    for (int q = 0; q < nbig; ++q) {   // large sub-triangles: all threads of the block share the tile walk
        const BigEntry& e = s_big[q];
        SubTri st;
        make_subtri(e.rec.ax, e.rec.ay, e.rec.bx, e.rec.by, e.rec.cx, e.rec.cy, e.rec.az, e.rec.bz, e.rec.cz, st);
        const int n = e.ntx * e.nty;
        for (int i = threadIdx.x; i < n; i += SLB_SETUP_CHUNK) {
            int tx = e.tx0 + i % e.ntx, ty = e.ty0 + i / e.ntx;
            if (tile_may_overlap(st, tx, ty, W, H))
                bin_pair<EMIT>(tile_base + ty * tiles_x + tx, e.rec, tile_count, tile_off, pairs, capacity);
        }
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
__global__ void __launch_bounds__(1024) k_scan_sums(unsigned long long* __restrict__ block_sums, uint32_t n_blocks, uint32_t* __restrict__ totals) {
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
    if (threadIdx.x == 0) { totals[0] = (uint32_t)s_carry; totals[1] = (uint32_t)(s_carry >> 32); }   // pairs, active tiles
}
__global__ void __launch_bounds__(1024) k_scan_fix(const uint32_t* __restrict__ count, const unsigned long long* __restrict__ block_sums,
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
            if (c[i]) { ActiveTile a; a.tile = base + i; a.beg = (uint32_t)run; a.count = c[i]; a.pad = 0; active[(uint32_t)(run >> 32)] = a; }
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
#define SLB_RASTER_SMALL 16   // sub-triangles whose in-tile pixel box is at most this big are rasterised by ONE lane

// fragment-stage discards that decide coverage: depth peel + alpha test (render_shader.frag:229-246)
__device__ __noinline__ bool frag_discard(const DFrame& f, const DDraw& d, uint32_t seq, int k, int px, int py) {
    const uint32_t tri = seq - d.prim_base;
    const uint32_t* ip = d.idx + 3 * (size_t)tri;
    const uint32_t vi[3] = {__ldg(ip), __ldg(ip + 1), __ldg(ip + 2)};
    PolyV a, b, c;
    if (!fetch_subtri(f, d, tri, seq, k, vi, a, b, c)) return true;
    SubTri st;
    if (!make_subtri(a.X, a.Y, b.X, b.Y, c.X, c.Y, a.z, b.z, c.z, st)) return true;
    FragIn in; float bary[3]; uint32_t vid[3];
    shade_inputs(f, d, st, a, b, c, vi, px, py, d.tex[0] != nullptr, in, bary, vid);
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
    keys[lane] = SLB_KEY_EMPTY; keys[lane + 32] = SLB_KEY_EMPTY;

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
            if (lane < cnt) {
                const PairRec r = s_rec[warp][b][lane];
                SubTri st;
                make_subtri(r.ax, r.ay, r.bx, r.by, r.cx, r.cy, r.az, r.bz, r.cz, st);
                int px0, py0, px1, py1;
                if (subtri_pixel_bbox(st, W, H, px0, py0, px1, py1)) {
                    px0 = max(px0, x_lo); px1 = min(px1, x_hi); py0 = max(py0, y_lo); py1 = min(py1, y_hi);
                    if (px0 <= px1 && py0 <= py1) {
                        if ((px1 - px0 + 1) * (py1 - py0 + 1) > SLB_RASTER_SMALL) big = true;
                        else {
                            const int cx0 = px0 * 256 + 128, cy0 = py0 * 256 + 128;
                            long long a0 = st.s * edge_fn(st.bx, st.by, st.cx, st.cy, cx0, cy0) + st.bias0;
                            long long a1 = st.s * edge_fn(st.cx, st.cy, st.ax, st.ay, cx0, cy0) + st.bias1;
                            long long a2 = st.s * edge_fn(st.ax, st.ay, st.bx, st.by, cx0, cy0) + st.bias2;
                            const long long dx0 = (long long)(-st.s * (st.cy - st.by)) * 256, dy0 = (long long)(st.s * (st.cx - st.bx)) * 256;
                            const long long dx1 = (long long)(-st.s * (st.ay - st.cy)) * 256, dy1 = (long long)(st.s * (st.ax - st.cx)) * 256;
                            const long long dx2 = (long long)(-st.s * (st.by - st.ay)) * 256, dy2 = (long long)(st.s * (st.bx - st.ax)) * 256;
                            const float dzb = __fsub_rn(st.bz, st.az), dzc = __fsub_rn(st.cz, st.az);
                            const unsigned long long lowkey = ((unsigned long long)r.seq << 8) | (r.k_flags & 0xffu);
                            for (int y = py0; y <= py1; ++y, a0 += dy0, a1 += dy1, a2 += dy2) {
                                long long e0 = a0, e1 = a1, e2 = a2;
                                for (int x = px0; x <= px1; ++x, e0 += dx0, e1 += dx1, e2 += dx2) {
                                    if ((e0 | e1 | e2) < 0) continue;
                                    long long w1 = e1 - st.bias1, w2 = e2 - st.bias2;
                                    if (st.s < 0) { w1 = -w1; w2 = -w2; }
                                    float q1 = __fmul_rn(__ll2float_rn(w1), st.inv2A), q2 = __fmul_rn(__ll2float_rn(w2), st.inv2A);
                                    float z = __fmaf_rn(q2, dzc, __fmaf_rn(q1, dzb, st.az));
                                    z = fminf(fmaxf(z, 0.0f), 1.0f);
                                    const unsigned long long key = ((unsigned long long)__float2uint_rn(__fmul_rn(z, 16777215.0f)) << 40) | lowkey;
                                    unsigned long long* slot = keys + (y - y_lo) * SLB_TILE + (x - x_lo);
                                    if (FRAG) {
                                        if ((r.k_flags & 0x100u) && (key >= *(volatile unsigned long long*)slot ||
                                                                     frag_discard(f, draws[r.draw], r.seq, r.k_flags & 0xff, x, y)))
                                            continue;
                                    }
                                    atomicMin(slot, key);
                                }
                            }
                        }
                    }
                }
            }
            unsigned bigmask = __ballot_sync(0xffffffffu, big);
            __syncwarp();
            // ---- large records: the whole warp, two pixels per lane ----
            while (bigmask) {
                const int j = __ffs(bigmask) - 1;
                bigmask &= bigmask - 1;
                const PairRec& r = s_rec[warp][b][j];
                SubTri st;
                make_subtri(r.ax, r.ay, r.bx, r.by, r.cx, r.cy, r.az, r.bz, r.cz, st);
                const unsigned long long lowkey = ((unsigned long long)r.seq << 8) | (r.k_flags & 0xffu);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int x = x_lo + lx, y = y_lo + ly + 4 * h;
                    if (x > x_hi || y > y_hi) continue;
                    long long w0, w1, w2; subtri_weights(st, x, y, w0, w1, w2);
                    if (!subtri_covers(st, w0, w1, w2)) continue;
                    const unsigned long long key = ((unsigned long long)subtri_depth24(st, w1, w2) << 40) | lowkey;
                    unsigned long long* slot = keys + (ly + 4 * h) * SLB_TILE + lx;
                    if (key >= *slot) continue;
                    if (FRAG) {
                        if ((r.k_flags & 0x100u) && frag_discard(f, draws[r.draw], r.seq, r.k_flags & 0xff, x, y)) continue;
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
            if (y_lo + ly < H) out[(size_t)(y_lo + ly) * W + px] = k0 == SLB_KEY_EMPTY ? 0xFFFFFFu : (uint32_t)(k0 >> 40);
            if (y_lo + ly + 4 < H) out[(size_t)(y_lo + ly + 4) * W + px] = k1 == SLB_KEY_EMPTY ? 0xFFFFFFu : (uint32_t)(k1 >> 40);
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

__global__ void __launch_bounds__(256) k_shade(const DFrame* __restrict__ frames, const DDraw* __restrict__ draws) {
    const DFrame& f = frames[blockIdx.z];
    const int W = f.W, H = f.H;
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= W || py >= H) return;
    const size_t p = (size_t)py * W + px;
    const unsigned long long key = f.keys[p];
    float4* const o_coord = reinterpret_cast<float4*>(f.out[SLB_TARGET_COORD]);
    unsigned short* const o_cls = reinterpret_cast<unsigned short*>(f.out[SLB_TARGET_CLASS]);
    unsigned short* const o_inst = reinterpret_cast<unsigned short*>(f.out[SLB_TARGET_INSTANCE]);
    uint4* const o_vidx = reinterpret_cast<uint4*>(f.out[SLB_TARGET_VERTEX_INDEX]);
    float4* const o_bary = reinterpret_cast<float4*>(f.out[SLB_TARGET_BARY]);

    float4 hdr = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
    bool shaded = false;
    if (key != SLB_KEY_EMPTY) {
        const uint32_t seq = (uint32_t)(key >> 8);
        const int k = (int)(key & 0xffu);
        const DDraw& d = draws[find_draw_by_prim(draws, f.draw_begin, f.draw_end, seq)];
        const uint32_t tri = seq - d.prim_base;
        const uint32_t* ip = d.idx + 3 * (size_t)tri;
        const uint32_t vi[3] = {__ldg(ip), __ldg(ip + 1), __ldg(ip + 2)};
        PolyV va, vb, vc;
        SubTri st;
        if (fetch_subtri(f, d, tri, seq, k, vi, va, vb, vc) && make_subtri(va.X, va.Y, vb.X, vb.Y, vc.X, vc.Y, va.z, vb.z, vc.z, st)) {
            FragIn in; float bary[3]; uint32_t vid[3];
            shade_inputs(f, d, st, va, vb, vc, vi, px, py, draw_has_textures(d), in, bary, vid);
            // geometry targets first: their registers are free before the lighting code runs
            if (o_coord) o_coord[p] = in.objc;
            if (o_cls) o_cls[p] = (unsigned short)d.class_index;
            if (o_inst) o_inst[p] = (unsigned short)d.instance_index;
            if (o_vidx) o_vidx[p] = make_uint4(vid[0], vid[1], vid[2], 0u);
            if (o_bary) o_bary[p] = make_float4(bary[0], bary[1], bary[2], 1.0f);
            if (f.scratch_cam) f.scratch_cam[p] = make_float4(in.cc.x, in.cc.y, in.cc.z, 1.0f);   // == out[CAM_COORD] when requested
            fragment_stage(f, d, in, vi, bary, hdr, nrm);
            shaded = true;
        }
    }
    if (!shaded) {   // clear values (render_pass.cpp:316,523-532)
        const float4 inval = make_float4(SLB_INVALID_COORD, SLB_INVALID_COORD, SLB_INVALID_COORD, SLB_INVALID_COORD);
        if (o_coord) o_coord[p] = inval;
        if (o_cls) o_cls[p] = 0;
        if (o_inst) o_inst[p] = 0;
        if (o_vidx) o_vidx[p] = make_uint4(0u, 0u, 0u, 0u);
        if (o_bary) o_bary[p] = make_float4(0.f, 0.f, 0.f, 1.0f);
        if (f.scratch_cam) f.scratch_cam[p] = inval;
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

void launch_bin(bool emit, const DView* views, const DFrame* frames, const DBinDraw* bdraws, const uint32_t* chunk_draw,
                uint32_t n_chunks, uint32_t* tile_count, const uint32_t* tile_off, PairRec* pairs, uint32_t capacity, cudaStream_t s) {
    if (n_chunks == 0) return;
    if (emit) k_bin<true><<<n_chunks, SLB_SETUP_CHUNK, 0, s>>>(views, frames, bdraws, chunk_draw, tile_count, tile_off, pairs, capacity);
    else k_bin<false><<<n_chunks, SLB_SETUP_CHUNK, 0, s>>>(views, frames, bdraws, chunk_draw, tile_count, tile_off, pairs, capacity);
}
void launch_scan(const uint32_t* count, uint32_t* off, ActiveTile* active, unsigned long long* block_sums, uint32_t* totals, uint32_t n,
                 cudaStream_t s) {
    const uint32_t n_blocks = (n + 1 + SLB_SCAN_BLOCK - 1) / SLB_SCAN_BLOCK;   // n + 1: some thread must own index n (the total)
    k_scan_local<<<n_blocks, 1024, 0, s>>>(count, block_sums, n);
    k_scan_sums<<<1, 1024, 0, s>>>(block_sums, n_blocks, totals);
    k_scan_fix<<<n_blocks, 1024, 0, s>>>(count, block_sums, off, active, n);
}
void launch_raster(bool frag_test, const DView* views, const DFrame* frames, const DDraw* draws, const ActiveTile* active,
                   const PairRec* pairs, RasterGrid g, cudaStream_t s) {
    if (g.n_active == 0) return;
    const unsigned grid = (g.n_active + SLB_RASTER_WARPS - 1) / SLB_RASTER_WARPS;
    if (frag_test) k_raster<true><<<grid, SLB_RASTER_WARPS * 32, 0, s>>>(views, frames, draws, active, pairs, g);
    else k_raster<false><<<grid, SLB_RASTER_WARPS * 32, 0, s>>>(views, frames, draws, active, pairs, g);
}
void launch_shade(const DFrame* frames, const DDraw* draws, int n_frames, int W, int H, cudaStream_t s) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, n_frames);
    k_shade<<<grid, 256, 0, s>>>(frames, draws);
}

}  // namespace slbk
