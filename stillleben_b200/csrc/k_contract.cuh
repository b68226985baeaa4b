// k_contract.cuh — the fixed-function front end (clip, viewport, snap, edge functions, depth) as
// device functions. These replace what the GL driver does between the reference's vertex and
// fragment stages (state: src/render_pass.cpp:325-332,373,502,523; GL 4.5 core §13.5-13.6, §14.6,
// §17.3) and implement the numerical contract C1-C7 of DESIGN.md "Raster contract" with explicit
// round-to-nearest intrinsics, so that coverage, depth24 and therefore every ID map are
// bit-reproducible regardless of compiler contraction settings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slbk {

constexpr float kGuard = 16.0f;   // guard band: clip x,y only beyond |x| > 16 w   (C4)

struct ClipV { float x, y, z, w; float b[3]; };
struct PolyV { int32_t X, Y; float z, invw; float b[3]; };
struct PrimSetup { int n; PolyV v[10]; };

// C3: clip = M * (p,1), one fma chain per row
__device__ __forceinline__ void xform_clip(const float* __restrict__ m, float px, float py, float pz, ClipV& c) {
    c.x = __fmaf_rn(m[0], px, __fmaf_rn(m[4], py, __fmaf_rn(m[8], pz, m[12])));
    c.y = __fmaf_rn(m[1], px, __fmaf_rn(m[5], py, __fmaf_rn(m[9], pz, m[13])));
    c.z = __fmaf_rn(m[2], px, __fmaf_rn(m[6], py, __fmaf_rn(m[10], pz, m[14])));
    c.w = __fmaf_rn(m[3], px, __fmaf_rn(m[7], py, __fmaf_rn(m[11], pz, m[15])));
}
__device__ __forceinline__ float plane_dist(const ClipV& v, int plane) {   // C4
    switch (plane) {
        case 0: return __fadd_rn(v.z, v.w);
        case 1: return __fsub_rn(v.w, v.z);
        case 2: return __fmaf_rn(kGuard, v.w, v.x);
        case 3: return __fmaf_rn(kGuard, v.w, -v.x);
        case 4: return __fmaf_rn(kGuard, v.w, v.y);
        default: return __fmaf_rn(kGuard, v.w, -v.y);
    }
}
__device__ __forceinline__ int frustum_code(const ClipV& v) {
    return (v.x < -v.w ? 1 : 0) | (v.x > v.w ? 2 : 0) | (v.y < -v.w ? 4 : 0) | (v.y > v.w ? 8 : 0) |
           (v.z < -v.w ? 16 : 0) | (v.z > v.w ? 32 : 0);
}
__device__ __forceinline__ int need_mask(const ClipV& v) {
    int need = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) if (!(plane_dist(v, p) >= 0.0f)) need |= (1 << p);
    return need;
}

// All three vertices inside the true frustum (codes == 0) and finite: then no clip plane of C4 can be violated
// (-w <= x <= w implies w >= 0 and 16 w +- x >= 0, ...), so need_mask() is known to be 0 without evaluating its
// 18 plane distances. Non-finite coordinates compare false everywhere, so they are sent to the full test.
__device__ __forceinline__ bool inside_frustum(int codes, const ClipV& a, const ClipV& b, const ClipV& c) {
    const float sum = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + ((c.x + c.y) + (c.z + c.w));
    return codes == 0 && (sum - sum) == 0.0f;
}

static __device__ __noinline__ int clip_poly(const ClipV* in, int n, int plane, ClipV* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const ClipV& a = in[i];
        const ClipV& b = in[(i + 1) % n];
        float da = plane_dist(a, plane), db = plane_dist(b, plane);
        bool ia = da >= 0.0f, ib = db >= 0.0f;
        if (ia) out[m++] = a;
        if (ia != ib) {
            const ClipV& I = ia ? a : b;   // interpolate from the inside vertex (watertight)
            const ClipV& O = ia ? b : a;
            float dI = ia ? da : db, dO = ia ? db : da;
            float t = __fdiv_rn(dI, __fsub_rn(dI, dO));
            ClipV r;
            r.x = __fmaf_rn(t, __fsub_rn(O.x, I.x), I.x);
            r.y = __fmaf_rn(t, __fsub_rn(O.y, I.y), I.y);
            r.z = __fmaf_rn(t, __fsub_rn(O.z, I.z), I.z);
            r.w = __fmaf_rn(t, __fsub_rn(O.w, I.w), I.w);
#pragma unroll
            for (int k = 0; k < 3; ++k) r.b[k] = __fmaf_rn(t, __fsub_rn(O.b[k], I.b[k]), I.b[k]);
            out[m++] = r;
        }
    }
    return m;
}

// C5: perspective divide, viewport transform, snap to 1/256 pixel (round-to-nearest-even)
__device__ __forceinline__ bool project_vertex(const ClipV& v, float hw, float hh, PolyV& o) {
    if (!(v.w > 0.0f)) return false;
    float invw = __frcp_rn(v.w);   // correctly rounded 1/w (== __fdiv_rn(1, w))
    float xw = __fmaf_rn(__fmul_rn(v.x, invw), hw, hw);
    float yw = __fmaf_rn(__fmul_rn(v.y, invw), hh, hh);
    o.X = __float2int_rn(__fmul_rn(xw, 256.0f));
    o.Y = __float2int_rn(__fmul_rn(yw, 256.0f));
    o.z = __fmaf_rn(__fmul_rn(v.z, invw), 0.5f, 0.5f);
    o.invw = invw;
    o.b[0] = v.b[0]; o.b[1] = v.b[1]; o.b[2] = v.b[2];
    return true;
}

// slow path: polygon clipping against the planes in `need`
static __device__ __noinline__ bool setup_clipped(const ClipV c[3], int need, float hw, float hh, PrimSetup& ps) {
    ClipV bufA[10], bufB[10];
    ClipV* cur = bufA; ClipV* nxt = bufB;
    int n = 3;
    for (int i = 0; i < 3; ++i) cur[i] = c[i];
    for (int p = 0; p < 6 && n >= 3; ++p) {
        if (!(need & (1 << p))) continue;
        n = clip_poly(cur, n, p, nxt);
        ClipV* t = cur; cur = nxt; nxt = t;
    }
    ps.n = 0;
    if (n < 3) return false;
    for (int i = 0; i < n; ++i)
        if (!project_vertex(cur[i], hw, hh, ps.v[i])) return false;
    ps.n = n;
    return true;
}

// clip + viewport + snap. Returns false if the primitive is culled.
__device__ __forceinline__ bool setup_prim(const float* __restrict__ mvp, float3 p0, float3 p1, float3 p2, int W, int H,
                                           PrimSetup& ps, bool* clipped = nullptr) {
    ClipV c[3];
    xform_clip(mvp, p0.x, p0.y, p0.z, c[0]);
    xform_clip(mvp, p1.x, p1.y, p1.z, c[1]);
    xform_clip(mvp, p2.x, p2.y, p2.z, c[2]);
    ps.n = 0;
    if (frustum_code(c[0]) & frustum_code(c[1]) & frustum_code(c[2])) return false;
    int need = need_mask(c[0]) | need_mask(c[1]) | need_mask(c[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) { c[i].b[0] = c[i].b[1] = c[i].b[2] = 0.0f; c[i].b[i] = 1.0f; }
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    if (clipped) *clipped = need != 0;
    if (need) return setup_clipped(c, need, hw, hh, ps);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (!project_vertex(c[i], hw, hh, ps.v[i])) return false;
    ps.n = 3;
    return true;
}

// Sub-triangle k of the (possibly clipped) primitive with its three vertices returned BY VALUE, so the
// common unclipped case (k == 1) stays in registers; only the clipped case touches local arrays.
static __device__ __noinline__ bool setup_subtri_clipped(const ClipV c[3], int need, float hw, float hh, int k, PolyV& a, PolyV& b,
                                                         PolyV& cc) {
    PrimSetup ps;
    if (!setup_clipped(c, need, hw, hh, ps)) return false;
    if (k < 1 || k + 1 >= ps.n) return false;
    a = ps.v[0]; b = ps.v[k]; cc = ps.v[k + 1];
    return true;
}
// Fast path of the per-pixel re-setup: returns 1 with (a,b,cc) filled when the primitive needs no clipping
// (then k must be 1), 0 when it is culled, and -1 when it needs clipping (the caller then fetches the
// binner's ClipRec or falls back to setup_subtri_clipped through resetup_clipped()).
__device__ __forceinline__ int setup_subtri_fast(const float* __restrict__ mvp, float3 p0, float3 p1, float3 p2, int W, int H, int k,
                                                 PolyV& a, PolyV& b, PolyV& cc) {
    ClipV c[3];
    xform_clip(mvp, p0.x, p0.y, p0.z, c[0]);
    xform_clip(mvp, p1.x, p1.y, p1.z, c[1]);
    xform_clip(mvp, p2.x, p2.y, p2.z, c[2]);
    const int f0 = frustum_code(c[0]), f1 = frustum_code(c[1]), f2 = frustum_code(c[2]);
    if (f0 & f1 & f2) return 0;
    if (!inside_frustum(f0 | f1 | f2, c[0], c[1], c[2]) && (need_mask(c[0]) | need_mask(c[1]) | need_mask(c[2]))) return -1;
    if (k != 1) return 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { c[i].b[0] = c[i].b[1] = c[i].b[2] = 0.0f; c[i].b[i] = 1.0f; }
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    return (project_vertex(c[0], hw, hh, a) && project_vertex(c[1], hw, hh, b) && project_vertex(c[2], hw, hh, cc)) ? 1 : 0;
}
static __device__ __noinline__ bool resetup_clipped(const float* __restrict__ mvp, float3 p0, float3 p1, float3 p2, int W, int H, int k,
                                                    PolyV& a, PolyV& b, PolyV& cc) {
    ClipV c[3];
    xform_clip(mvp, p0.x, p0.y, p0.z, c[0]);
    xform_clip(mvp, p1.x, p1.y, p1.z, c[1]);
    xform_clip(mvp, p2.x, p2.y, p2.z, c[2]);
    int need = need_mask(c[0]) | need_mask(c[1]) | need_mask(c[2]);
    for (int i = 0; i < 3; ++i) { c[i].b[0] = c[i].b[1] = c[i].b[2] = 0.0f; c[i].b[i] = 1.0f; }
    return setup_subtri_clipped(c, need, 0.5f * (float)W, 0.5f * (float)H, k, a, b, cc);
}

__device__ __forceinline__ long long edge_fn(int ax, int ay, int bx, int by, int px, int py) {
    return (long long)(bx - ax) * (long long)(py - ay) - (long long)(by - ay) * (long long)(px - ax);
}
// C6: top-left rule in (x, row) coordinates for orientation sign s
__device__ __forceinline__ bool top_left(int dx, int dy, int s) {
    return (dy == 0 && (long long)s * dx > 0) || ((long long)s * dy < 0);
}

// A snapped sub-triangle ready for coverage / depth evaluation.
struct SubTri {
    int ax, ay, bx, by, cx, cy;
    float az, bz, cz;
    long long twoA;
    int s;
    float inv2A;
    int bias0, bias1, bias2;   // 0 or -1
};
__device__ __forceinline__ bool make_subtri(int ax, int ay, int bx, int by, int cx, int cy, float az, float bz, float cz,
                                            SubTri& t) {
    t.ax = ax; t.ay = ay; t.bx = bx; t.by = by; t.cx = cx; t.cy = cy; t.az = az; t.bz = bz; t.cz = cz;
    t.twoA = edge_fn(ax, ay, bx, by, cx, cy);
    if (t.twoA == 0) return false;
    t.s = t.twoA > 0 ? 1 : -1;
    t.inv2A = __frcp_rn(__ll2float_rn(t.twoA));
    t.bias0 = top_left(cx - bx, cy - by, t.s) ? 0 : -1;   // w0 <-> edge b->c
    t.bias1 = top_left(ax - cx, ay - cy, t.s) ? 0 : -1;   // w1 <-> edge c->a
    t.bias2 = top_left(bx - ax, by - ay, t.s) ? 0 : -1;   // w2 <-> edge a->b
    return true;
}
__device__ __forceinline__ bool make_subtri(const PrimSetup& ps, int k, SubTri& t) {
    const PolyV &a = ps.v[0], &b = ps.v[k], &c = ps.v[k + 1];
    return make_subtri(a.X, a.Y, b.X, b.Y, c.X, c.Y, a.z, b.z, c.z, t);
}
__device__ __forceinline__ void subtri_weights(const SubTri& t, int px, int py, long long& w0, long long& w1, long long& w2) {
    int cx = px * 256 + 128, cy = py * 256 + 128;
    w0 = edge_fn(t.bx, t.by, t.cx, t.cy, cx, cy);
    w1 = edge_fn(t.cx, t.cy, t.ax, t.ay, cx, cy);
    w2 = edge_fn(t.ax, t.ay, t.bx, t.by, cx, cy);
}
__device__ __forceinline__ bool subtri_covers(const SubTri& t, long long w0, long long w1, long long w2) {
    return (t.s * w0 + t.bias0 >= 0) && (t.s * w1 + t.bias1 >= 0) && (t.s * w2 + t.bias2 >= 0);
}
// C7: depth24 from screen-space barycentrics
__device__ __forceinline__ uint32_t subtri_depth24(const SubTri& t, long long w1, long long w2) {
    float b1 = __fmul_rn(__ll2float_rn(w1), t.inv2A), b2 = __fmul_rn(__ll2float_rn(w2), t.inv2A);
    float z = __fmaf_rn(b2, __fsub_rn(t.cz, t.az), __fmaf_rn(b1, __fsub_rn(t.bz, t.az), t.az));
    z = fminf(fmaxf(z, 0.0f), 1.0f);
    return __float2uint_rn(__fmul_rn(z, 16777215.0f));
}
// pixel bounding box of a sub-triangle (pixel centres inside the snapped bbox), clamped to the viewport
__device__ __forceinline__ bool subtri_pixel_bbox(const SubTri& t, int W, int H, int& px0, int& py0, int& px1, int& py1) {
    int xmin = min(t.ax, min(t.bx, t.cx)), xmax = max(t.ax, max(t.bx, t.cx));
    int ymin = min(t.ay, min(t.by, t.cy)), ymax = max(t.ay, max(t.by, t.cy));
    px0 = max(0, (xmin - 128 + 255) >> 8); px1 = min(W - 1, (xmax - 128) >> 8);
    py0 = max(0, (ymin - 128 + 255) >> 8); py1 = min(H - 1, (ymax - 128) >> 8);
    return px0 <= px1 && py0 <= py1;
}

}  // namespace slbk
