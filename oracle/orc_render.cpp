// orc_render.cpp — ORACLE (test infrastructure only; see orc_core.h header for the rules).
//
// CPU restatement of sl::RenderPass::render (reference: src/render_pass.cpp:303-796) and of the
// GLSL programs it runs:
//   vertex stage      src/shaders/render_shader.vert:57-95
//   geometry stage    src/shaders/render_shader.geom:13-35
//   fragment stage    src/shaders/render_shader.frag:225-412
//   uniforms          src/shaders/render_shader.cpp:233-460
//   shadow pass       src/render_pass.cpp:69-211,407-460, src/shaders/shadow_shader.vert:10-13
//   background        src/render_pass.cpp:637-660, src/shaders/background_*.{vert,frag}
//   SSAO              src/shaders/ssao_shader.{frag,cpp}, ssao_apply_shader.frag
//   tone map          src/shaders/tone_map_shader.frag:102-131
// Fixed-function stages (clip, viewport, rasterise, depth) follow the GL 4.5 core spec
// §13.5-13.6, §14.6, §17.3 under the NUMERICAL CONTRACT written in DESIGN.md §"Raster contract":
// 8 sub-pixel bits, top-left rule in (x, row) coordinates, 24-bit depth, LESS, first draw wins.
// PARITY STATUS: programmable stages pinned on the reference's GLSL text compiled as C++ (tests/test_glsl_ref.py); the GL
// fixed function follows the written contract (see orc_core.h).
#include "orc_core.h"
#include "orc_test_hooks.h"

#include <cstdio>
#include <limits>
#include <random>
#include <omp.h>

namespace orc {

// ---------------------------------------------------------------------------------------
// Geometry front end (numerical contract C1-C7)
// ---------------------------------------------------------------------------------------
static const float kGuard = 16.0f;  // guard band in NDC units: clip only beyond |x| > 16 w

static void mul44d(const double* a, const double* b, double* c) {  // column-major, C2: fixed order
    for (int col = 0; col < 4; ++col)
        for (int r = 0; r < 4; ++r)
            c[col * 4 + r] = ((a[0 * 4 + r] * b[col * 4 + 0] + a[1 * 4 + r] * b[col * 4 + 1]) + a[2 * 4 + r] * b[col * 4 + 2]) +
                             a[3 * 4 + r] * b[col * 4 + 3];
}
static void to_d(const float* f, double* d) { for (int i = 0; i < 16; ++i) d[i] = f[i]; }
static void to_f(const double* d, float* f) { for (int i = 0; i < 16; ++i) f[i] = (float)d[i]; }

struct ClipV { float x, y, z, w; float b[3]; };
struct PolyV { int32_t X, Y; float z, invw; float b[3]; };
struct PrimSetup { int n; PolyV v[10]; };

static inline void xform_clip(const float* m, const float* p, ClipV& c) {  // C3: fma chain
    c.x = std::fmaf(m[0], p[0], std::fmaf(m[4], p[1], std::fmaf(m[8], p[2], m[12])));
    c.y = std::fmaf(m[1], p[0], std::fmaf(m[5], p[1], std::fmaf(m[9], p[2], m[13])));
    c.z = std::fmaf(m[2], p[0], std::fmaf(m[6], p[1], std::fmaf(m[10], p[2], m[14])));
    c.w = std::fmaf(m[3], p[0], std::fmaf(m[7], p[1], std::fmaf(m[11], p[2], m[15])));
}
static inline float plane_dist(const ClipV& v, int plane) {  // C4
    switch (plane) {
        case 0: return v.z + v.w;                      // near  z >= -w
        case 1: return v.w - v.z;                      // far   z <=  w
        case 2: return std::fmaf(kGuard, v.w, v.x);    // x >= -G w
        case 3: return std::fmaf(kGuard, v.w, -v.x);   // x <=  G w
        case 4: return std::fmaf(kGuard, v.w, v.y);
        default: return std::fmaf(kGuard, v.w, -v.y);
    }
}
static int clip_poly(const ClipV* in, int n, int plane, ClipV* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const ClipV& a = in[i];
        const ClipV& b = in[(i + 1) % n];
        float da = plane_dist(a, plane), db = plane_dist(b, plane);
        bool ia = da >= 0.0f, ib = db >= 0.0f;
        if (ia) out[m++] = a;
        if (ia != ib) {
            const ClipV& I = ia ? a : b;  // always interpolate from the inside vertex (watertight)
            const ClipV& O = ia ? b : a;
            float dI = ia ? da : db, dO = ia ? db : da;
            float t = dI / (dI - dO);
            ClipV r;
            r.x = std::fmaf(t, O.x - I.x, I.x);
            r.y = std::fmaf(t, O.y - I.y, I.y);
            r.z = std::fmaf(t, O.z - I.z, I.z);
            r.w = std::fmaf(t, O.w - I.w, I.w);
            for (int k = 0; k < 3; ++k) r.b[k] = std::fmaf(t, O.b[k] - I.b[k], I.b[k]);
            out[m++] = r;
        }
    }
    return m;
}

// clip + viewport transform + snap. Returns false if the primitive is culled.
static bool setup_prim(const float* mvp, const float* p0, const float* p1, const float* p2, int W, int H, PrimSetup& ps) {
    ClipV c[3];
    xform_clip(mvp, p0, c[0]); xform_clip(mvp, p1, c[1]); xform_clip(mvp, p2, c[2]);
    int code_and = 0x3f, need = 0;
    for (int i = 0; i < 3; ++i) {
        const ClipV& v = c[i];
        int fc = (v.x < -v.w ? 1 : 0) | (v.x > v.w ? 2 : 0) | (v.y < -v.w ? 4 : 0) | (v.y > v.w ? 8 : 0) |
                 (v.z < -v.w ? 16 : 0) | (v.z > v.w ? 32 : 0);
        code_and &= fc;
        for (int p = 0; p < 6; ++p) if (!(plane_dist(v, p) >= 0.0f)) need |= (1 << p);
        c[i].b[0] = c[i].b[1] = c[i].b[2] = 0.0f; c[i].b[i] = 1.0f;
    }
    ps.n = 0;
    if (code_and) return false;  // all three outside one frustum plane
    ClipV bufA[10], bufB[10];
    ClipV* cur = bufA; ClipV* nxt = bufB;
    int n = 3;
    for (int i = 0; i < 3; ++i) cur[i] = c[i];
    if (need) {
        for (int p = 0; p < 6 && n >= 3; ++p) {
            if (!(need & (1 << p))) continue;
            n = clip_poly(cur, n, p, nxt);
            std::swap(cur, nxt);
        }
        if (n < 3) return false;
    }
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    for (int i = 0; i < n; ++i) {
        const ClipV& v = cur[i];
        if (!(v.w > 0.0f)) return false;
        float invw = 1.0f / v.w;
        float xw = std::fmaf(v.x * invw, hw, hw);
        float yw = std::fmaf(v.y * invw, hh, hh);
        PolyV& o = ps.v[i];
        o.X = (int32_t)std::lrintf(xw * 256.0f);   // C5: round-to-nearest-even, 8 sub-pixel bits
        o.Y = (int32_t)std::lrintf(yw * 256.0f);
        o.z = std::fmaf(v.z * invw, 0.5f, 0.5f);
        o.invw = invw;
        o.b[0] = v.b[0]; o.b[1] = v.b[1]; o.b[2] = v.b[2];
    }
    ps.n = n;
    return true;
}

static inline int64_t edge_fn(int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t px, int32_t py) {
    return (int64_t)(bx - ax) * (int64_t)(py - ay) - (int64_t)(by - ay) * (int64_t)(px - ax);
}
static inline bool top_left(int32_t dx, int32_t dy, int s) {  // C6
    return (dy == 0 && (int64_t)s * dx > 0) || ((int64_t)s * dy < 0);
}

struct SubTri {
    const PolyV *a, *b, *c;
    int64_t twoA; int s; float inv2A;
    int64_t bias0, bias1, bias2;
};
static inline bool make_subtri(const PrimSetup& ps, int k, SubTri& t) {
    t.a = &ps.v[0]; t.b = &ps.v[k]; t.c = &ps.v[k + 1];
    t.twoA = edge_fn(t.a->X, t.a->Y, t.b->X, t.b->Y, t.c->X, t.c->Y);
    if (t.twoA == 0) return false;
    t.s = t.twoA > 0 ? 1 : -1;
    t.inv2A = 1.0f / (float)t.twoA;
    // w0 <-> edge b->c, w1 <-> edge c->a, w2 <-> edge a->b
    t.bias0 = top_left(t.c->X - t.b->X, t.c->Y - t.b->Y, t.s) ? 0 : -1;
    t.bias1 = top_left(t.a->X - t.c->X, t.a->Y - t.c->Y, t.s) ? 0 : -1;
    t.bias2 = top_left(t.b->X - t.a->X, t.b->Y - t.a->Y, t.s) ? 0 : -1;
    return true;
}
static inline void subtri_weights(const SubTri& t, int px, int py, int64_t& w0, int64_t& w1, int64_t& w2) {
    int32_t cx = px * 256 + 128, cy = py * 256 + 128;
    w0 = edge_fn(t.b->X, t.b->Y, t.c->X, t.c->Y, cx, cy);
    w1 = edge_fn(t.c->X, t.c->Y, t.a->X, t.a->Y, cx, cy);
    w2 = edge_fn(t.a->X, t.a->Y, t.b->X, t.b->Y, cx, cy);
}
static inline bool subtri_covers(const SubTri& t, int64_t w0, int64_t w1, int64_t w2) {
    return (t.s * w0 + t.bias0 >= 0) && (t.s * w1 + t.bias1 >= 0) && (t.s * w2 + t.bias2 >= 0);
}
static inline uint32_t subtri_depth24(const SubTri& t, int64_t w1, int64_t w2) {  // C7
    float b1 = (float)w1 * t.inv2A, b2 = (float)w2 * t.inv2A;
    float z = std::fmaf(b2, t.c->z - t.a->z, std::fmaf(b1, t.b->z - t.a->z, t.a->z));
    z = std::min(std::max(z, 0.0f), 1.0f);
    return (uint32_t)std::lrintf(z * 16777215.0f);
}
// perspective-correct barycentrics w.r.t. the ORIGINAL triangle at pixel (px,py) of sub-triangle t
static inline void subtri_bary(const SubTri& t, int px, int py, float out[3]) {
    int64_t w0, w1, w2; subtri_weights(t, px, py, w0, w1, w2);
    float b0 = (float)w0 * t.inv2A, b1 = (float)w1 * t.inv2A, b2 = (float)w2 * t.inv2A;
    float g0 = b0 * t.a->invw, g1 = b1 * t.b->invw, g2 = b2 * t.c->invw;
    float s = g0 + g1 + g2;
    float q0 = g0 / s, q1 = g1 / s, q2 = g2 / s;
    for (int j = 0; j < 3; ++j) out[j] = q0 * t.a->b[j] + q1 * t.b->b[j] + q2 * t.c->b[j];
}

// ---------------------------------------------------------------------------------------
// Frame set-up
// ---------------------------------------------------------------------------------------
struct Draw {
    const Mesh* mesh = nullptr;        // nullptr for the background plane
    uint32_t index_offset = 0, n_tris = 0;
    uint32_t prim_base = 0;
    M4 meshToObject, objectToWorld;
    float mvp[16];
    float normalToWorld[9];
    Material mat;
    const Texture* tex[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t class_index = 0, instance_index = 0;
    const Texture* sticker = nullptr;
    M4 stickerProj; float stickerRange[4];
    bool alpha_tested = false;
};

static Vertex68 g_plane_verts[4];
static uint32_t g_plane_idx[6] = {0, 1, 2, 2, 1, 3};
static void init_plane() {
    // Magnum Primitives::planeSolid(TextureCoordinates): triangle strip (1,-1) (1,1) (-1,-1) (-1,1)
    // (contrib/magnum/src/Magnum/Primitives/Plane.cpp:36-60); no tangent / vertex-id attributes.
    const float P[4][2] = {{1, -1}, {1, 1}, {-1, -1}, {-1, 1}};
    const float U[4][2] = {{1, 0}, {1, 1}, {0, 0}, {0, 1}};
    for (int i = 0; i < 4; ++i) {
        Vertex68& v = g_plane_verts[i];
        std::memset(&v, 0, sizeof v);
        v.pos[0] = P[i][0]; v.pos[1] = P[i][1]; v.pos[2] = 0;
        v.uv[0] = U[i][0]; v.uv[1] = U[i][1];
        v.color[3] = 1.0f; v.tangent[3] = 1.0f;
        v.normal[2] = 1.0f; v.vertex_index = 0;
    }
}

struct Frame {
    int W, H;
    M4 P, V;
    V3 camPos;
    V3 lightDir[SLB_NUM_LIGHTS], lightCol[SLB_NUM_LIGHTS];
    bool lightActive[SLB_NUM_LIGHTS];
    V3 ambient;
    const LightMap* lm;
    M4 shadowMat[SLB_NUM_LIGHTS];
    std::vector<uint32_t> shadowMap[SLB_NUM_LIGHTS];  // d24, SLB_SHADOW_RES^2
    std::vector<Draw> draws;
    const float* peel;  // HxWx4 or null
};

static const Vertex68* draw_verts(const Draw& d) { return d.mesh ? d.mesh->verts.data() : g_plane_verts; }
static const uint32_t* draw_indices(const Draw& d) { return (d.mesh ? d.mesh->indices.data() : g_plane_idx) + d.index_offset; }

// reference: src/render_pass.cpp:69-129
static void frustum_corners(const slb_scene_desc& sc, V3 corners[8]) {
    M4 P = M4::from(sc.projection), V = M4::from(sc.world_to_cam);
    M4 Pinv = inverted(P);
    float nearv = -1.0f, farv = 1.0f;
    if (sc.n_objects > 0) {
        float nearObj = std::numeric_limits<float>::infinity(), farObj = -nearObj;
        for (int i = 0; i < sc.n_objects; ++i) {
            const slb_object_desc& o = sc.objects[i];
            const Mesh* m = (const Mesh*)o.mesh;
            M4 pre = M4::from(o.pretransform);
            V3 lo = transform_point(pre, V3(m->bbox_min[0], m->bbox_min[1], m->bbox_min[2]));
            V3 hi = transform_point(pre, V3(m->bbox_max[0], m->bbox_max[1], m->bbox_max[2]));
            V3 center = (lo + hi) * 0.5f;   // Range3D::center() = (min+max)/2
            float radius = length(hi - lo) / 2;
            V3 objInCam = transform_point(mul(V, M4::from(o.pose)), center);
            V3 np = transform_point(P, objInCam - V3(0, 0, radius));
            V3 fp = transform_point(P, objInCam + V3(0, 0, radius));
            nearObj = std::min(nearObj, np.z);
            farObj = std::max(farObj, fp.z);
        }
        nearv = std::max(std::max(-1.0f, nearObj), nearv);
        farv = std::min(farObj, farv);
    }
    const float hc[8][3] = {{-1, 1, nearv}, {1, 1, nearv}, {1, -1, nearv}, {-1, -1, nearv},
                            {-1, 1, farv},  {1, 1, farv},  {1, -1, farv},  {-1, -1, farv}};
    M4 camToWorld = inverted_rigid(V);
    for (int i = 0; i < 8; ++i) {
        V4 p = mul(camToWorld, mul(Pinv, V4(hc[i][0], hc[i][1], hc[i][2], 1.0f)));
        corners[i] = V3(p.x / p.w, p.y / p.w, p.z / p.w);
    }
}
// reference: src/render_pass.cpp:131-211
static M4 shadow_matrix(const slb_scene_desc& sc, const V3 corners[8], V3 lightDirection) {
    V3 z = normalize(lightDirection);
    V3 x = normalize(cross(z, V3(0, 0, 1)));
    V3 y = normalize(cross(z, x));
    M4 camToWorld = M4::identity();
    for (int r = 0; r < 3; ++r) { camToWorld.at(r, 0) = x[r]; camToWorld.at(r, 1) = y[r]; camToWorld.at(r, 2) = z[r]; }
    M4 worldToCam = inverted_rigid(camToWorld);
    const float inf = std::numeric_limits<float>::infinity();
    V3 mn(inf), mx(-inf);
    for (int i = 0; i < 8; ++i) { V3 c = transform_point(worldToCam, corners[i]); mn = vmin(mn, c); mx = vmax(mx, c); }
    float nearv = mn.z, farv = mx.z;
    float meanZ = (nearv + farv) / 2.0f;
    float spread = farv - meanZ;
    farv = meanZ + 5.0f * spread;
    nearv = meanZ - 5.0f * spread;
    float L = mn.x, R = mx.x, T = mn.y, B = mx.y;
    if (sc.n_objects > 0) {
        V3 maxObj(inf), minObj(-inf);  // (sic: names swapped in the reference)
        for (int i = 0; i < sc.n_objects; ++i) {
            const slb_object_desc& o = sc.objects[i];
            const Mesh* m = (const Mesh*)o.mesh;
            M4 pre = M4::from(o.pretransform);
            V3 lo = transform_point(pre, V3(m->bbox_min[0], m->bbox_min[1], m->bbox_min[2]));
            V3 hi = transform_point(pre, V3(m->bbox_max[0], m->bbox_max[1], m->bbox_max[2]));
            float radius = length(hi - lo) / 2;
            V3 c = transform_point(mul(worldToCam, M4::from(o.pose)), (lo + hi) * 0.5f);
            maxObj = vmin(maxObj, c - V3(radius));
            minObj = vmax(minObj, c + V3(radius));
        }
        L = std::max(L, maxObj.x); R = std::min(R, minObj.x);
        T = std::max(T, maxObj.y); B = std::min(B, minObj.y);
    }
    M4 Pm; std::memset(Pm.m, 0, sizeof Pm.m);
    Pm.at(0, 0) = 2.0f / (R - L); Pm.at(1, 1) = 2.0f / (B - T); Pm.at(2, 2) = 2.0f / (farv - nearv);
    Pm.at(0, 3) = -(R + L) / (R - L); Pm.at(1, 3) = -(B + T) / (B - T); Pm.at(2, 3) = -(farv + nearv) / (farv - nearv);
    Pm.at(3, 3) = 1.0f;
    return mul(Pm, worldToCam);
}

static void resolve_material(Draw& d, const Mesh* mesh, int material, float ovr_metallic, float ovr_roughness) {
    Material m;
    if (mesh && material >= 0 && material < (int)mesh->materials.size()) m = mesh->materials[material];
    else {
        // context default material (src/context.cpp:382-384): 0x3bd267ff_srgbaf, no textures
        m.base_color[0] = 0.04373503f; m.base_color[1] = 0.6444797f; m.base_color[2] = 0.13563333f; m.base_color[3] = 1.0f;
        m.emissive[0] = m.emissive[1] = m.emissive[2] = 0.0f; m.emissive[3] = 0.0f;
        m.metallic = 0.04f; m.roughness = 0.5f;
        for (int i = 0; i < 5; ++i) m.tex[i] = -1;
    }
    if (ovr_metallic >= 0.0f) m.metallic = ovr_metallic;    // render_shader.cpp:372-377
    if (ovr_roughness >= 0.0f) m.roughness = ovr_roughness;
    d.mat = m;
    for (int i = 0; i < 5; ++i)
        d.tex[i] = (mesh && m.tex[i] >= 0 && m.tex[i] < (int)mesh->textures.size()) ? &mesh->textures[m.tex[i]] : nullptr;
    d.alpha_tested = (d.tex[0] && d.tex[0]->has_alpha) || d.mat.base_color[3] < 0.5f;
}

static void make_mvp(const M4& P, const M4& V, const M4& world, const M4& pre, float* mvp) {
    double p[16], v[16], w[16], m[16], mw[16], mc[16], r[16];
    to_d(P.m, p); to_d(V.m, v); to_d(world.m, w); to_d(pre.m, m);
    mul44d(w, m, mw); mul44d(v, mw, mc); mul44d(p, mc, r);
    to_f(r, mvp);
}

// ---------------------------------------------------------------------------------------
// Vertex + fragment stages
// ---------------------------------------------------------------------------------------
struct VSOut {
    V2 uv; V3 nW, tW, bW; V4 objc; V3 wc, cc; V2 sticker;
};
static void vertex_stage(const Frame& f, const Draw& d, const Vertex68& v, VSOut& o) {
    V4 position(v.pos[0], v.pos[1], v.pos[2], 1.0f);
    V4 oc4 = mul(d.meshToObject, position);
    o.objc = V4(oc4.x / oc4.w, oc4.y / oc4.w, oc4.z / oc4.w, 1.0f);
    V4 wc4 = mul(d.objectToWorld, oc4);
    o.wc = V3(wc4.x / wc4.w, wc4.y / wc4.w, wc4.z / wc4.w);
    V4 cc4 = mul(f.V, wc4);
    o.cc = V3(cc4.x / cc4.w, cc4.y / cc4.w, cc4.z / cc4.w);
    o.objc.w = o.cc.z;
    V3 n(v.normal[0], v.normal[1], v.normal[2]), t(v.tangent[0], v.tangent[1], v.tangent[2]);
    o.nW = normalize(mul3(d.normalToWorld, n));
    o.tW = normalize(mul3(d.normalToWorld, t));
    o.bW = normalize(cross(o.nW, o.tW)) * v.tangent[3];
    o.uv = V2{v.uv[0], v.uv[1]};
    V4 sp = mul(d.stickerProj, oc4);
    o.sticker = V2{(sp.x / sp.w - d.stickerRange[0]) / d.stickerRange[2], (sp.y / sp.w - d.stickerRange[1]) / d.stickerRange[3]};
}

static inline V3 powv(V3 v, float e) { return V3(std::pow(v.x, e), std::pow(v.y, e), std::pow(v.z, e)); }
static inline V3 reflectv(V3 I, V3 N) { return I - N * (2.0f * dot(N, I)); }

static float DistributionGGX(V3 N, V3 H, float roughness) {
    float a = roughness * roughness, a2 = a * a;
    float NdotH = std::max(dot(N, H), 0.0f), NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = 3.141592653589793f * denom * denom;
    return a2 / denom;
}
static float GeometrySchlickGGX(float NdotV, float roughness) {
    float r = roughness + 1.0f, k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
static float GeometrySmith(V3 N, V3 V, V3 L, float roughness) {
    return GeometrySchlickGGX(std::max(dot(N, L), 0.0f), roughness) * GeometrySchlickGGX(std::max(dot(N, V), 0.0f), roughness);
}

// sampler2DArrayShadow lookup: linear filter of (ref <= stored), clamp-to-edge (Appendix A.8)
static float shadow_tap(const std::vector<uint32_t>& map, float u, float v, float ref) {
    const int N = SLB_SHADOW_RES;
    ref = clampf(ref, 0.0f, 1.0f);
    float x = u * N - 0.5f, y = v * N - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    auto cmp = [&](int i, int j) {
        i = std::min(std::max(i, 0), N - 1); j = std::min(std::max(j, 0), N - 1);
        float stored = (float)map[(size_t)j * N + i] / 16777215.0f;
        return ref <= stored ? 1.0f : 0.0f;
    };
    return cmp(i0, j0) * ((1 - a) * (1 - b)) + cmp(i0 + 1, j0) * (a * (1 - b)) + cmp(i0, j0 + 1) * ((1 - a) * b) +
           cmp(i0 + 1, j0 + 1) * (a * b);
}

struct FragIn {
    V2 uv, uv_dx, uv_dy;  // uv at the pixel, and at its quad neighbours in x / y
    V3 nW, tW, bW; V4 objc; V3 wc, cc; V2 sticker;
    bool front;
};
struct FragOut {
    V4 color; V4 objc; V4 camc; V4 normal; uint32_t cls, inst; uint32_t vid[3]; float bary[3];
};

static inline V4 sample_mat(const Texture* t, const FragIn& in) {
    float sx = (in.uv_dx.x - in.uv.x), sy = (in.uv_dx.y - in.uv.y);
    float tx = (in.uv_dy.x - in.uv.x), ty = (in.uv_dy.y - in.uv.y);
    return sample_texture_2d(*t, in.uv.x, in.uv.y, sx, sy, tx, ty);
}
static inline V4 to_linear(V4 c) { return V4(std::pow(c.x, 2.2f), std::pow(c.y, 2.2f), std::pow(c.z, 2.2f), c.w); }

// base colour incl. alpha — shared by the discard test and the shading (render_shader.frag:237-246)
static V4 base_color(const Draw& d, const FragIn& in) {
    V4 bc(d.mat.base_color[0], d.mat.base_color[1], d.mat.base_color[2], d.mat.base_color[3]);
    if (d.tex[0]) { V4 t = to_linear(sample_mat(d.tex[0], in)); bc = V4(bc.x * t.x, bc.y * t.y, bc.z * t.z, bc.w * t.w); }
    return bc;
}

// reference: src/shaders/render_shader.frag:225-412 (the discards are evaluated by the caller)
static void fragment_stage(const Frame& f, const Draw& d, const FragIn& in, FragOut& out) {
    const float PI = 3.141592653589793f;
    V4 baseColor = base_color(d, in);
    if (d.sticker && in.sticker.x >= 0 && in.sticker.y >= 0 && in.sticker.x < 1 && in.sticker.y < 1) {
        V4 sc = to_linear(sample_texture_rect_linear(*d.sticker, in.sticker.x * d.sticker->w, in.sticker.y * d.sticker->h));
        float a = sc.w;
        baseColor = V4(mixf(baseColor.x, sc.x, a), mixf(baseColor.y, sc.y, a), mixf(baseColor.z, sc.z, a), mixf(baseColor.w, sc.w, a));
    }
    V3 normal;
    if (d.tex[1]) {
        V4 t = sample_mat(d.tex[1], in);
        V3 n(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
        normal = normalize(in.tW * n.x + in.bW * n.y + in.nW * n.z);
    } else normal = in.nW;
    if (!in.front) normal = -normal;

    V3 cameraDirection = normalize(f.camPos - in.wc);
    V3 lightDir = reflectv(-cameraDirection, normal);
    float NoV = clampf(dot(normal, cameraDirection), 1e-5f, 1.0f);

    float roughness = d.mat.roughness, metallic = d.mat.metallic;
    if (d.tex[2]) { V4 t = sample_mat(d.tex[2], in); roughness *= t.y; metallic *= t.z; }
    roughness = std::max(roughness, 0.045f);
    float occlusion = 1.0f;
    if (d.tex[4]) occlusion = sample_mat(d.tex[4], in).x;
    V3 emissive(d.mat.emissive[0], d.mat.emissive[1], d.mat.emissive[2]);
    if (d.tex[3]) { V4 t = to_linear(sample_mat(d.tex[3], in)); emissive = emissive * V3(t.x, t.y, t.z); }

    V3 color(0.0f);
    V3 bc(baseColor.x, baseColor.y, baseColor.z);
    V3 c_diff = bc * (1.0f - 0.04f) * (1.0f - metallic);
    V3 F0 = mix(V3(0.04f), bc, metallic);
    V3 Fr = vmax(V3(1.0f - roughness), F0) - F0;
    V3 k_S = F0 + Fr * std::pow(1.0f - NoV, 5.0f);

    const float shadowMapScale = 1.0f / (float)SLB_SHADOW_RES;
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i) {
        if (!f.lightActive[i]) continue;
        V4 pc = mul(f.shadowMat[i], V4(in.wc, 1.0f));
        pc = V4(pc.x / pc.w, pc.y / pc.w, pc.z / pc.w, 1.0f);
        pc = V4(0.5f * pc.x + 0.5f, 0.5f * pc.y + 0.5f, 0.5f * pc.z + 0.5f, 1.0f);
        float inverseShadow = 0.0f;
        for (float y = -1.5f; y <= 1.5f; y += 1.0f)
            for (float x = -1.5f; x <= 1.5f; x += 1.0f)
                inverseShadow += shadow_tap(f.shadowMap[i], pc.x + x * shadowMapScale, pc.y + y * shadowMapScale, pc.z - 0.00003f);
        inverseShadow /= 16.0f;

        V3 L = normalize(-f.lightDir[i]);
        V3 H = normalize(cameraDirection + L);
        V3 radiance = f.lightCol[i];
        float NDF = DistributionGGX(normal, H, roughness);
        float G = GeometrySmith(normal, cameraDirection, L, roughness);
        V3 nominator = k_S * (NDF * G);
        float denominator = 4.0f * NoV * std::max(dot(normal, L), 0.0f);
        V3 specular = nominator / std::max(denominator, 0.001f);
        V3 kD = (V3(1.0f) - k_S) * (1.0f - metallic);
        float NdotL = std::max(dot(normal, L), 0.0f);
        color += (kD * bc / PI + specular) * radiance * (inverseShadow * NdotL);
    }
    color += f.ambient * bc;

    if (f.lm) {
        V4 fab = sample_lut(*f.lm, NoV, roughness);
        float lodLevel = roughness * 4.0f;
        V4 rad = sample_cube_lod(f.lm->prefilter, lightDir, lodLevel);
        V4 irr = sample_cube_lod(f.lm->irradiance, normal, 0.0f);
        V3 radiance(rad.x, rad.y, rad.z), irradiance(irr.x, irr.y, irr.z);
        V3 FssEss = k_S * fab.x + V3(fab.y);
        float Ems = 1.0f - (fab.x + fab.y);
        V3 F_avg = F0 + (V3(1.0f) - F0) / 21.0f;
        V3 FmsEms = FssEss * F_avg * Ems / (V3(1.0f) - F_avg * Ems);
        V3 k_D = c_diff * (V3(1.0f) - FssEss - FmsEms);
        V3 selfColor = FssEss * radiance + (FmsEms + k_D) * irradiance;
        color += selfColor * occlusion;
    }
    color += emissive;

    out.color = V4(color, baseColor.w);
    out.objc = in.objc;
    out.camc = V4(in.cc, 1.0f);
    out.cls = d.class_index; out.inst = d.instance_index;
    // mat3(worldToCam) * normal
    V3 nc(f.V.at(0, 0) * normal.x + f.V.at(0, 1) * normal.y + f.V.at(0, 2) * normal.z,
          f.V.at(1, 0) * normal.x + f.V.at(1, 1) * normal.y + f.V.at(1, 2) * normal.z,
          f.V.at(2, 0) * normal.x + f.V.at(2, 1) * normal.y + f.V.at(2, 2) * normal.z);
    nc = normalize(nc);
    out.normal = V4(nc, dot(normal, cameraDirection));
}

// Interpolate the vertex-stage outputs of one primitive at pixel (px,py), incl. the uv of the two
// quad neighbours (helper-invocation semantics: same primitive, extrapolated; "fine" derivatives).
static void interpolate(const Frame& f, const Draw& d, const SubTri& st, const VSOut vs[3], int px, int py, FragIn& in, float bary[3]) {
    subtri_bary(st, px, py, bary);
    auto lerp3 = [&](V3 a, V3 b, V3 c) { return a * bary[0] + b * bary[1] + c * bary[2]; };
    in.uv = V2{vs[0].uv.x * bary[0] + vs[1].uv.x * bary[1] + vs[2].uv.x * bary[2],
               vs[0].uv.y * bary[0] + vs[1].uv.y * bary[1] + vs[2].uv.y * bary[2]};
    in.nW = lerp3(vs[0].nW, vs[1].nW, vs[2].nW);
    in.tW = lerp3(vs[0].tW, vs[1].tW, vs[2].tW);
    in.bW = lerp3(vs[0].bW, vs[1].bW, vs[2].bW);
    in.wc = lerp3(vs[0].wc, vs[1].wc, vs[2].wc);
    in.cc = lerp3(vs[0].cc, vs[1].cc, vs[2].cc);
    for (int k = 0; k < 4; ++k) in.objc[k] = vs[0].objc[k] * bary[0] + vs[1].objc[k] * bary[1] + vs[2].objc[k] * bary[2];
    in.sticker = V2{vs[0].sticker.x * bary[0] + vs[1].sticker.x * bary[1] + vs[2].sticker.x * bary[2],
                    vs[0].sticker.y * bary[0] + vs[1].sticker.y * bary[1] + vs[2].sticker.y * bary[2]};
    in.front = st.twoA < 0;  // FrontFace = CW (render_pass.cpp:330, Appendix A.4)
    float bx[3], by[3];
    subtri_bary(st, px ^ 1, py, bx);   // quad partner in x  (dFdx = +-(partner - self))
    subtri_bary(st, px, py ^ 1, by);
    float sgx = (px & 1) ? -1.0f : 1.0f, sgy = (py & 1) ? -1.0f : 1.0f;
    V2 ux{vs[0].uv.x * bx[0] + vs[1].uv.x * bx[1] + vs[2].uv.x * bx[2], vs[0].uv.y * bx[0] + vs[1].uv.y * bx[1] + vs[2].uv.y * bx[2]};
    V2 uy{vs[0].uv.x * by[0] + vs[1].uv.x * by[1] + vs[2].uv.x * by[2], vs[0].uv.y * by[0] + vs[1].uv.y * by[1] + vs[2].uv.y * by[2]};
    in.uv_dx = V2{in.uv.x + sgx * (ux.x - in.uv.x), in.uv.y + sgx * (ux.y - in.uv.y)};
    in.uv_dy = V2{in.uv.x + sgy * (uy.x - in.uv.x), in.uv.y + sgy * (uy.y - in.uv.y)};
    (void)f; (void)d;
}

static void fetch_vs(const Frame& f, const Draw& d, uint32_t tri, VSOut vs[3], uint32_t vid[3], const float* pos[3]) {
    const Vertex68* verts = draw_verts(d);
    const uint32_t* idx = draw_indices(d) + 3 * (size_t)tri;
    for (int k = 0; k < 3; ++k) {
        const Vertex68& v = verts[idx[k]];
        vertex_stage(f, d, v, vs[k]);
        vid[k] = v.vertex_index;
        pos[k] = v.pos;
    }
}

// ---------------------------------------------------------------------------------------
// Post passes
// ---------------------------------------------------------------------------------------
static std::vector<float> downsample(const std::vector<float>& s, int sw, int sh, int& dw, int& dh) {
    dw = std::max(1, sw >> 1); dh = std::max(1, sh >> 1);
    std::vector<float> d((size_t)dw * dh * 4);
    auto taps = [](int sN, int dN, int i, int idx[3], float w[3]) {
        if (sN == 1) { idx[0] = idx[1] = idx[2] = 0; w[0] = 1; w[1] = w[2] = 0; return; }
        if ((sN & 1) == 0) { idx[0] = 2 * i; idx[1] = idx[2] = 2 * i + 1; w[0] = w[1] = 0.5f; w[2] = 0; return; }
        idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 2;
        w[0] = (float)(dN - i) / sN; w[1] = (float)dN / sN; w[2] = (float)(i + 1) / sN;
    };
    for (int y = 0; y < dh; ++y) {
        int iy[3]; float wy[3]; taps(sh, dh, y, iy, wy);
        for (int x = 0; x < dw; ++x) {
            int ix[3]; float wx[3]; taps(sw, dw, x, ix, wx);
            for (int c = 0; c < 4; ++c) {
                double acc = 0;
                for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a)
                    acc += (double)(wy[b] * wx[a]) * s[((size_t)iy[b] * sw + ix[a]) * 4 + c];
                d[((size_t)y * dw + x) * 4 + c] = (float)acc;
            }
        }
    }
    return d;
}

// linear-filtered rectangle texture read of an RGBA32F HxWx4 image, clamp-to-edge
static V4 rect_linear(const float* img, int W, int H, float xs, float ys) {
    float x = xs - 0.5f, y = ys - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    auto at = [&](int i, int j) {
        i = std::min(std::max(i, 0), W - 1); j = std::min(std::max(j, 0), H - 1);
        const float* p = img + ((size_t)j * W + i) * 4; return V4(p[0], p[1], p[2], p[3]);
    };
    return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) + at(i0 + 1, j0 + 1) * (a * b);
}

// reference: src/shaders/ssao_shader.cpp:72-112 (mt19937(0xdeadbeef) noise + kernel)
static void ssao_tables(V3 noise[16], V3 kernel[64]) {
    std::mt19937 random{0xdeadbeef};
    std::uniform_real_distribution<float> rf(0.0f, 1.0f);
    for (int i = 0; i < 16; ++i) { float a = 2.0f * rf(random) - 1.0f; float b = 2.0f * rf(random) - 1.0f; noise[i] = V3(a, b, 0.0f); }
    for (int i = 0; i < 64; ++i) {
        float a = 2.0f * rf(random) - 1.0f; float b = 2.0f * rf(random) - 1.0f; float c = rf(random);
        V3 s(a, b, c);
        s = normalize(s) * rf(random);
        float scale = (float)i / 64.0f;
        float t = scale * scale;
        s = s * (0.1f * (1.0f - t) + 1.0f * t);  // Math::lerp(0.1, 1.0, t)
        kernel[i] = s;
    }
}

static inline float smoothstep01(float x) { float t = clampf(x, 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }

static V3 aces(V3 x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    V3 r;
    for (int i = 0; i < 3; ++i) {
        float v = (x[i] * (a * x[i] + b)) / (x[i] * (c * x[i] + d) + e);
        r[i] = (v != v) ? 0.0f : clampf(v, 0.0f, 1.0f);   // clamp(NaN) -> 0 (Hard parts: NaN on black)
    }
    return r;
}

static inline uint8_t unorm8(float v) {
    if (!(v == v)) return 0;
    v = clampf(v, 0.0f, 1.0f);
    return (uint8_t)std::lrintf(v * 255.0f);
}


// tone map of one pixel (tone_map_shader.frag:102-131); avg = texel (0,0) of the last mip level (auto exposure only)
static void tone_map_pixel(const float* hdr4, float manual_exposure, const float avg[4], uint8_t* rgba8) {
    V3 c(hdr4[0], hdr4[1], hdr4[2]);
    V3 xyz(0.4124564f * c.x + 0.3575761f * c.y + 0.1804375f * c.z, 0.2126729f * c.x + 0.7151522f * c.y + 0.0721750f * c.z,
           0.0193339f * c.x + 0.1191920f * c.y + 0.9503041f * c.z);
    float inv = 1.0f / (xyz.x + xyz.y + xyz.z);
    V3 Yxy(xyz.y, xyz.x * inv, xyz.y * inv);
    if (manual_exposure >= 0) Yxy.x *= manual_exposure;
    else {
        float lum = 0.1f * (0.2125f * (avg[0] / avg[3]) + 0.7154f * (avg[1] / avg[3]) + 0.0721f * (avg[2] / avg[3]));
        Yxy.x /= (9.6f * lum + 0.0001f);
    }
    V3 x2(Yxy.x * Yxy.y / Yxy.z, Yxy.x, Yxy.x * (1.0f - Yxy.y - Yxy.z) / Yxy.z);
    V3 r(3.2404542f * x2.x - 1.5371385f * x2.y - 0.4985314f * x2.z, -0.9692660f * x2.x + 1.8760108f * x2.y + 0.0415560f * x2.z,
         0.0556434f * x2.x - 0.2040259f * x2.y + 1.0572252f * x2.z);
    V3 a = aces(r);
    rgba8[0] = unorm8(a.x); rgba8[1] = unorm8(a.y); rgba8[2] = unorm8(a.z); rgba8[3] = unorm8(hdr4[3]);
}

// SSAO (ssao_shader.frag:20-57) over a whole frame: camc / normals are the HxWx4 float targets, P the projection
static void ssao_pass(const float* camc, const float* normals, int W, int H, const M4& P, float* ao, int n_threads) {
    V3 noise[16], kernel[64]; ssao_tables(noise, kernel);
    #pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t p = (size_t)py * W + px;
            V3 fragPos(camc[p * 4], camc[p * 4 + 1], camc[p * 4 + 2]);
            V3 nrm(normals[p * 4], normals[p * 4 + 1], normals[p * 4 + 2]);
            if (nrm.x == 0 && nrm.y == 0 && nrm.z == 0) { ao[p] = 1.0f; continue; }  // background: NaN path -> no occlusion
            V3 normal = normalize(nrm);
            V3 randomVec = normalize(noise[(py & 3) * 4 + (px & 3)]);
            V3 tangent = normalize(randomVec - normal * dot(randomVec, normal));
            V3 bitangent = cross(normal, tangent);
            float occlusion = 0.0f;
            for (int i = 0; i < 64; ++i) {
                V3 s = kernel[i];
                V3 samplePos = tangent * s.x + bitangent * s.y + normal * s.z;
                samplePos = fragPos + samplePos * 0.1f;
                V4 off = mul(P, V4(samplePos, 1.0f));
                float ox = off.x / off.w * 0.5f + 0.5f, oy = off.y / off.w * 0.5f + 0.5f;
                float sampleDepth = rect_linear(camc, W, H, ox * W, oy * H).z;
                float rangeCheck = smoothstep01(0.1f / std::fabs(fragPos.z - sampleDepth));
                occlusion += (sampleDepth <= samplePos.z - 0.0025f ? 1.0f : 0.0f) * rangeCheck;
            }
            ao[p] = 1.0f - (occlusion / 64.0f);
        }
}
// bilateral blur + apply (ssao_apply_shader.frag:29-76): out rgb = hdr rgb * blurred ao, alpha untouched
static void ssao_apply_pass(const float* hdr, const float* ao, const float* camc, int W, int H, float* outc, int n_threads) {
    #pragma omp parallel for num_threads(n_threads)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t p = (size_t)py * W + px;
            float center_d = rect_linear(camc, W, H, (float)px, (float)py).z;
            float result = 0.0f, w_total = 0.0f;
            const float BlurSigma = 3.0f * 0.5f, BlurFalloff = 1.0f / (2.0f * BlurSigma * BlurSigma);
            for (int x = -2; x < 2; ++x)
                for (int y = -2; y < 2; ++y) {
                    int ux = px + x, uy = py + y;
                    float c = (ux >= 0 && ux < W && uy >= 0 && uy < H) ? ao[(size_t)uy * W + ux] : 0.0f;  // texelFetch out of range -> 0
                    float dd = rect_linear(camc, W, H, (float)ux, (float)uy).z;
                    float r = std::sqrt((float)(x * x + y * y));
                    float ddiff = (dd - center_d) * 300.0f;
                    float w = std::exp2(-r * r * BlurFalloff - ddiff * ddiff);
                    w_total += w;
                    result += c * w;
                }
            float a = result / w_total;
            outc[p * 4 + 0] = hdr[p * 4 + 0] * a; outc[p * 4 + 1] = hdr[p * 4 + 1] * a; outc[p * 4 + 2] = hdr[p * 4 + 2] * a;
            outc[p * 4 + 3] = hdr[p * 4 + 3];
        }
}
// direction the sky box is sampled with at NDC (xn, yn): the cube position interpolated over the cube faces == the view
// ray in the un-translated world frame, R^T (Pinv ndc)  (background_cube_shader.vert:14-20)
static V3 skybox_dir(const M4& Pinv, const M4& V, float xn, float yn) {
    V4 q = mul(Pinv, V4(xn, yn, 1.0f, 1.0f));
    V3 dc(q.x / q.w, q.y / q.w, q.z / q.w);
    return V3(V.at(0, 0) * dc.x + V.at(1, 0) * dc.y + V.at(2, 0) * dc.z, V.at(0, 1) * dc.x + V.at(1, 1) * dc.y + V.at(2, 1) * dc.z,
              V.at(0, 2) * dc.x + V.at(1, 2) * dc.y + V.at(2, 2) * dc.z);
}
// background image texel for pixel (px, py) (background_shader.vert:12-14, .frag:10-16)
static V4 background_image_texel(const Texture& bg, int px, int py, int W, int H) {
    float tx = ((px + 0.5f) / W), ty = 1.0f - ((py + 0.5f) / H);   // textureCoords = (x_ndc, -y_ndc)/2 + 0.5 at the pixel centre
    int ix = (int)(tx * bg.w), iy = (int)(ty * bg.h);              // ivec2(textureCoords * texSize)
    return sample_texture_rect_linear(bg, (float)ix, (float)iy);
}

}  // namespace orc

using namespace orc;

// =======================================================================================
// C interface of the oracle
// =======================================================================================
extern "C" {

void* orc_mesh_create(const void* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices,
                      const slb_submesh* submeshes, uint32_t n_submeshes, const slb_material* materials,
                      uint32_t n_materials, const slb_image* images, uint32_t n_images, const float bbox_min[3],
                      const float bbox_max[3]) {
    Mesh* m = new Mesh;
    m->verts.resize(n_vertices);
    std::memcpy(m->verts.data(), vertices, (size_t)n_vertices * sizeof(Vertex68));
    m->indices.assign(indices, indices + n_indices);
    m->submeshes.assign(submeshes, submeshes + n_submeshes);
    for (uint32_t i = 0; i < n_materials; ++i) {
        Material mm; const slb_material& s = materials[i];
        std::memcpy(mm.base_color, s.base_color, 16); std::memcpy(mm.emissive, s.emissive, 16);
        mm.metallic = s.metallic; mm.roughness = s.roughness;
        mm.tex[0] = s.tex_base_color; mm.tex[1] = s.tex_normal; mm.tex[2] = s.tex_metallic_roughness;
        mm.tex[3] = s.tex_emissive; mm.tex[4] = s.tex_occlusion;
        m->materials.push_back(mm);
    }
    m->textures.resize(n_images);
    for (uint32_t i = 0; i < n_images; ++i) build_texture(m->textures[i], &images[i], SLB_TEXTURE_2D);
    for (int k = 0; k < 3; ++k) { m->bbox_min[k] = bbox_min[k]; m->bbox_max[k] = bbox_max[k]; }
    return m;
}
void orc_mesh_update_vertices(void* mesh, const void* vertices, uint32_t n_vertices) {
    Mesh* m = (Mesh*)mesh;
    m->verts.resize(n_vertices);
    std::memcpy(m->verts.data(), vertices, (size_t)n_vertices * sizeof(Vertex68));
}
// Mesh::recomputeNormals (reference: src/mesh.cpp:763-816), literally: per face cross = (v1 - v2) x (v1 - v3), area = |cross|,
// normal = cross.normalized() (= cross * (1 / length), Magnum Vector::normalized); per vertex the sum of normal * area over
// its faces in ascending face order, normalised. A zero-area face yields NaN (0 * inf) and poisons its vertices, as there.
void orc_mesh_recompute_normals(void* mesh) {
    Mesh* m = (Mesh*)mesh;
    const size_t nf = m->indices.size() / 3, nv = m->verts.size();
    std::vector<V3> fn(nf);
    auto dot_rn = [](V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; };
    for (size_t f = 0; f < nf; ++f) {
        const float* p1 = m->verts[m->indices[3 * f]].pos; const float* p2 = m->verts[m->indices[3 * f + 1]].pos; const float* p3 = m->verts[m->indices[3 * f + 2]].pos;
        V3 a(p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]), b(p1[0] - p3[0], p1[1] - p3[1], p1[2] - p3[2]);
        V3 c(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
        float area = std::sqrt(dot_rn(c, c));
        float inv = 1.0f / area;
        fn[f] = V3((c.x * inv) * area, (c.y * inv) * area, (c.z * inv) * area);
    }
    std::vector<V3> acc(nv, V3(0.0f));
    for (size_t f = 0; f < nf; ++f)
        for (int k = 0; k < 3; ++k) { V3& n = acc[m->indices[3 * f + k]]; n = V3(n.x + fn[f].x, n.y + fn[f].y, n.z + fn[f].z); }
    for (size_t v = 0; v < nv; ++v) {
        float inv = 1.0f / std::sqrt(dot_rn(acc[v], acc[v]));
        m->verts[v].normal[0] = acc[v].x * inv; m->verts[v].normal[1] = acc[v].y * inv; m->verts[v].normal[2] = acc[v].z * inv;
    }
}
// Mesh::updateVertexPositionsAndColors (reference: src/mesh.cpp:823-855); ids are one-based. Returns -1 on a bad id.
int orc_mesh_update_positions_and_colors(void* mesh, const int32_t* ids, uint32_t n, const float* dpos, const float* dcol) {
    Mesh* m = (Mesh*)mesh;
    for (uint32_t i = 0; i < n; ++i) {
        int32_t v = ids[i] - 1;
        if (v < 0 || (size_t)v >= m->verts.size()) return -1;
        if (dpos) for (int k = 0; k < 3; ++k) m->verts[v].pos[k] = m->verts[v].pos[k] + dpos[3 * (size_t)i + k];
        if (dcol) for (int k = 0; k < 4; ++k) m->verts[v].color[k] = m->verts[v].color[k] + dcol[4 * (size_t)i + k];
    }
    if (dpos && n) orc_mesh_recompute_normals(mesh);
    return 0;
}
void orc_mesh_read_vertices(const void* mesh, void* out68) {
    const Mesh* m = (const Mesh*)mesh;
    std::memcpy(out68, m->verts.data(), m->verts.size() * sizeof(Vertex68));
}
void orc_mesh_destroy(void* m) { delete (Mesh*)m; }

void* orc_texture_create(const slb_image* image, int kind) {
    Texture* t = new Texture; build_texture(*t, image, kind); return t;
}
void orc_texture_destroy(void* t) { delete (Texture*)t; }
// read back one mip level (RGBA8) — lets tests pin the CUDA mip chain bit-exactly
int orc_texture_level(const void* tex, int level, int* w, int* h, uint8_t* out) {
    const Texture* t = (const Texture*)tex;
    if (level < 0 || level >= (int)t->levels.size()) return -1;
    *w = t->levels[level].w; *h = t->levels[level].h;
    if (out) std::memcpy(out, t->levels[level].px.data(), t->levels[level].px.size());
    return (int)t->levels.size();
}

// test hook: the d24 shadow maps of the last orc_render call (tests/test_gl_ref.py compares them with GL's depth layers)
static bool g_keep_shadow_maps = false;
static std::vector<uint32_t> g_kept_shadow_map[SLB_NUM_LIGHTS];
void orc_test_keep_shadow_maps(int on) { g_keep_shadow_maps = on != 0; if (!on) for (auto& m : g_kept_shadow_map) m.clear(); }
size_t orc_test_shadow_map(int light, uint32_t* out) {   // -> number of texels (0: that light was inactive); out may be NULL
    if (light < 0 || light >= SLB_NUM_LIGHTS) return 0;
    const std::vector<uint32_t>& m = g_kept_shadow_map[light];
    if (out && !m.empty()) std::memcpy(out, m.data(), m.size() * sizeof(uint32_t));
    return m.size();
}
// Render one scene. The handles inside `sc` (mesh, textures, light map) are ORACLE handles.
// out[t]: host arrays in the layouts of slb.h (NULL to skip); hdr_out: HxWx4 float pre-tone-map
// colour (after background + SSAO), may be NULL; peel: HxWx4 float coord buffer of the previous
// depth-peel layer or NULL.
int orc_render(const slb_scene_desc* scp, const float* peel, void* const out[SLB_NUM_TARGETS], float* hdr_out, int n_threads) {
    const slb_scene_desc& sc = *scp;
    static bool plane_init = false;
    if (!plane_init) { init_plane(); plane_init = true; }
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    Frame f;
    f.W = sc.width; f.H = sc.height; f.peel = peel;
    const int W = f.W, H = f.H;
    f.P = M4::from(sc.projection); f.V = M4::from(sc.world_to_cam);
    { M4 inv = inverted_rigid(f.V); f.camPos = V3(inv.at(0, 3), inv.at(1, 3), inv.at(2, 3)); }
    f.lm = (const LightMap*)sc.light_map;
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i) {
        f.lightDir[i] = V3(0.0f); f.lightCol[i] = V3(0.0f);
        if (f.lm) {
            if (i < f.lm->n_lights) {
                f.lightDir[i] = V3(f.lm->light_directions[i][0], f.lm->light_directions[i][1], f.lm->light_directions[i][2]);
                f.lightCol[i] = V3(f.lm->light_colors[i][0], f.lm->light_colors[i][1], f.lm->light_colors[i][2]);
            }
        } else {
            f.lightDir[i] = V3(sc.light_directions[i][0], sc.light_directions[i][1], sc.light_directions[i][2]);
            f.lightCol[i] = V3(sc.light_colors[i][0], sc.light_colors[i][1], sc.light_colors[i][2]);
        }
        bool colZero = f.lightCol[i].x == 0 && f.lightCol[i].y == 0 && f.lightCol[i].z == 0;
        bool dirZero = f.lightDir[i].x == 0 && f.lightDir[i].y == 0 && f.lightDir[i].z == 0;
        f.lightActive[i] = !(colZero || dirZero);
    }
    f.ambient = f.lm ? V3(0.0f) : V3(sc.ambient_light[0], sc.ambient_light[1], sc.ambient_light[2]);

    // ---- draw list in submission order (render_pass.cpp:545-622) ----
    uint32_t prim = 0;
    if (sc.background_plane_size[0] * sc.background_plane_size[0] + sc.background_plane_size[1] * sc.background_plane_size[1] > 0) {
        Draw d;
        d.mesh = nullptr; d.index_offset = 0; d.n_tris = 2; d.prim_base = prim; prim += 2;
        d.meshToObject = M4::identity();
        M4 scale = M4::identity();
        scale.at(0, 0) = sc.background_plane_size[0] / 2.0f; scale.at(1, 1) = sc.background_plane_size[1] / 2.0f;
        d.objectToWorld = mul(M4::from(sc.background_plane_pose), scale);
        resolve_material(d, nullptr, -1, -1.0f, -1.0f);
        const Texture* pt = (const Texture*)sc.background_plane_texture;
        if (pt) { d.mat.base_color[0] = d.mat.base_color[1] = d.mat.base_color[2] = d.mat.base_color[3] = 1.0f; d.tex[0] = pt; d.alpha_tested = pt->has_alpha; }
        else { d.mat.base_color[0] = 0.0f; d.mat.base_color[1] = 0.8f; d.mat.base_color[2] = 0.0f; d.mat.base_color[3] = 1.0f; }
        d.class_index = 0; d.instance_index = 0;
        d.sticker = nullptr; d.stickerProj = M4::identity();
        d.stickerRange[0] = d.stickerRange[1] = 0; d.stickerRange[2] = d.stickerRange[3] = 1e-6f;
        f.draws.push_back(d);
    }
    for (int i = 0; i < sc.n_objects; ++i) {
        const slb_object_desc& o = sc.objects[i];
        if (!o.visible) continue;
        const Mesh* mesh = (const Mesh*)o.mesh;
        for (const slb_submesh& sm : mesh->submeshes) {
            Draw d;
            d.mesh = mesh; d.index_offset = sm.index_offset; d.n_tris = sm.index_count / 3; d.prim_base = prim; prim += d.n_tris;
            d.meshToObject = M4::from(o.pretransform); d.objectToWorld = M4::from(o.pose);
            resolve_material(d, mesh, sm.material, o.metallic, o.roughness);
            d.class_index = o.class_index; d.instance_index = o.instance_index;
            d.sticker = (const Texture*)o.sticker_texture;
            d.stickerProj = M4::from(o.sticker_projection);
            d.stickerRange[0] = o.sticker_range[0]; d.stickerRange[1] = o.sticker_range[1];
            d.stickerRange[2] = std::max(1e-6f, o.sticker_range[2]); d.stickerRange[3] = std::max(1e-6f, o.sticker_range[3]);
            f.draws.push_back(d);
        }
    }
    for (Draw& d : f.draws) {
        make_mvp(f.P, f.V, d.objectToWorld, d.meshToObject, d.mvp);
        normal_matrix(mul(d.objectToWorld, d.meshToObject), d.normalToWorld);
    }

    // ---- shadow pass (render_pass.cpp:407-460) ----
    bool anyLight = f.lightActive[0] || f.lightActive[1] || f.lightActive[2];
    if (anyLight) {
        V3 corners[8]; frustum_corners(sc, corners);
        const int N = SLB_SHADOW_RES;
        for (int li = 0; li < SLB_NUM_LIGHTS; ++li) {
            if (!f.lightActive[li]) continue;
            f.shadowMat[li] = shadow_matrix(sc, corners, f.lightDir[li]);
            std::vector<uint32_t>& map = f.shadowMap[li];
            map.assign((size_t)N * N, 0xFFFFFFu);
            struct SDraw { const Mesh* mesh; uint32_t off, ntri; float mvp[16]; };
            std::vector<SDraw> sd;
            for (int i = 0; i < sc.n_objects; ++i) {
                const slb_object_desc& o = sc.objects[i];
                if (!o.visible || !o.casts_shadows) continue;
                const Mesh* mesh = (const Mesh*)o.mesh;
                for (const slb_submesh& sm : mesh->submeshes) {
                    SDraw s; s.mesh = mesh; s.off = sm.index_offset; s.ntri = sm.index_count / 3;
                    make_mvp(f.shadowMat[li], M4::identity(), M4::from(o.pose), M4::from(o.pretransform), s.mvp);
                    sd.push_back(s);
                }
            }
            // min() is order independent: parallelise over triangles with an atomic min
            uint32_t* mp = map.data();
            for (const SDraw& s : sd) {
                const Vertex68* verts = s.mesh->verts.data();
                const uint32_t* idx = s.mesh->indices.data() + s.off;
                #pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads)
                for (uint32_t t = 0; t < s.ntri; ++t) {
                    PrimSetup ps;
                    if (!setup_prim(s.mvp, verts[idx[3 * t]].pos, verts[idx[3 * t + 1]].pos, verts[idx[3 * t + 2]].pos, N, N, ps)) continue;
                    for (int k = 1; k + 1 < ps.n; ++k) {
                        SubTri st; if (!make_subtri(ps, k, st)) continue;
                        if (st.twoA < 0) continue;  // cull FRONT faces (render_pass.cpp:428-429)
                        int32_t xmin = std::min(st.a->X, std::min(st.b->X, st.c->X)), xmax = std::max(st.a->X, std::max(st.b->X, st.c->X));
                        int32_t ymin = std::min(st.a->Y, std::min(st.b->Y, st.c->Y)), ymax = std::max(st.a->Y, std::max(st.b->Y, st.c->Y));
                        int px0 = std::max(0, (xmin - 128 + 255) >> 8), px1 = std::min(N - 1, (xmax - 128) >> 8);
                        int py0 = std::max(0, (ymin - 128 + 255) >> 8), py1 = std::min(N - 1, (ymax - 128) >> 8);
                        for (int py = py0; py <= py1; ++py)
                            for (int px = px0; px <= px1; ++px) {
                                int64_t w0, w1, w2; subtri_weights(st, px, py, w0, w1, w2);
                                if (!subtri_covers(st, w0, w1, w2)) continue;
                                uint32_t d24 = subtri_depth24(st, w1, w2);
                                uint32_t* dst = &mp[(size_t)py * N + px];
                                uint32_t cur = __atomic_load_n(dst, __ATOMIC_RELAXED);
                                while (d24 < cur && !__atomic_compare_exchange_n(dst, &cur, d24, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                            }
                    }
                }
            }
        }
        if (g_keep_shadow_maps) for (int li = 0; li < SLB_NUM_LIGHTS; ++li) g_kept_shadow_map[li] = f.lightActive[li] ? f.shadowMap[li] : std::vector<uint32_t>();
    }

    // ---- main pass: visibility (key = depth24 << 32 | primitive sequence number) ----
    std::vector<uint64_t> keys((size_t)W * H, ~0ull);
    {
        uint64_t* kp = keys.data();
        for (const Draw& d : f.draws) {
            const Vertex68* verts = draw_verts(d);
            const uint32_t* idx = draw_indices(d);
            const bool need_frag = d.alpha_tested || f.peel;
            #pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads)
            for (uint32_t t = 0; t < d.n_tris; ++t) {
                PrimSetup ps;
                if (!setup_prim(d.mvp, verts[idx[3 * t]].pos, verts[idx[3 * t + 1]].pos, verts[idx[3 * t + 2]].pos, W, H, ps)) continue;
                VSOut vs[3]; bool have_vs = false;
                for (int k = 1; k + 1 < ps.n; ++k) {
                    SubTri st; if (!make_subtri(ps, k, st)) continue;
                    int32_t xmin = std::min(st.a->X, std::min(st.b->X, st.c->X)), xmax = std::max(st.a->X, std::max(st.b->X, st.c->X));
                    int32_t ymin = std::min(st.a->Y, std::min(st.b->Y, st.c->Y)), ymax = std::max(st.a->Y, std::max(st.b->Y, st.c->Y));
                    int px0 = std::max(0, (xmin - 128 + 255) >> 8), px1 = std::min(W - 1, (xmax - 128) >> 8);
                    int py0 = std::max(0, (ymin - 128 + 255) >> 8), py1 = std::min(H - 1, (ymax - 128) >> 8);
                    for (int py = py0; py <= py1; ++py)
                        for (int px = px0; px <= px1; ++px) {
                            int64_t w0, w1, w2; subtri_weights(st, px, py, w0, w1, w2);
                            if (!subtri_covers(st, w0, w1, w2)) continue;
                            if (need_frag) {
                                if (!have_vs) { uint32_t vid[3]; const float* pos[3]; fetch_vs(f, d, t, vs, vid, pos); have_vs = true; }
                                FragIn in; float bary[3];
                                interpolate(f, d, st, vs, px, py, in, bary);
                                // depth peeling (render_shader.frag:229-233)
                                if (f.peel && in.objc.w - 0.00001f <= f.peel[((size_t)py * W + px) * 4 + 3]) continue;
                                // alpha test (render_shader.frag:242-246), alphaCutoff hard-coded 0.5
                                if (d.alpha_tested && base_color(d, in).w < 0.5f) continue;
                            }
                            // depth test LESS + first draw wins on ties == min over (depth24, sequence)
                            uint64_t key = ((uint64_t)subtri_depth24(st, w1, w2) << 32) | (uint64_t)(d.prim_base + t);
                            uint64_t* dst = &kp[(size_t)py * W + px];
                            uint64_t cur = __atomic_load_n(dst, __ATOMIC_RELAXED);
                            while (key < cur && !__atomic_compare_exchange_n(dst, &cur, key, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                        }
                }
            }
        }
    }

    // ---- main pass: shade the visible fragment of every pixel; clear values elsewhere ----
    std::vector<float> hdr((size_t)W * H * 4, 0.0f);
    std::vector<float> coord((size_t)W * H * 4), normals((size_t)W * H * 4, 0.0f), camc((size_t)W * H * 4), bary_t((size_t)W * H * 4, 0.0f);
    std::vector<uint16_t> cls((size_t)W * H, 0), inst((size_t)W * H, 0);
    std::vector<uint32_t> vidx((size_t)W * H * 4, 0);
    #pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t p = (size_t)py * W + px;
            uint64_t key = keys[p];
            if (key == ~0ull) {
                for (int k = 0; k < 4; ++k) { coord[p * 4 + k] = SLB_INVALID_COORD; camc[p * 4 + k] = SLB_INVALID_COORD; }
                // GL clears the RGB32F-typed attachment 6 with (0,0,0) -> alpha reads 1 only for
                // 3-channel formats; the texture is RGBA32F, cleared by clearColor(6, 0x00000000_rgbf)
                // = Color3 -> Color4 with alpha 1 (render_pass.cpp:530)
                bary_t[p * 4 + 3] = 1.0f;
                continue;
            }
            uint32_t seq = (uint32_t)(key & 0xffffffffu);
            // find draw
            size_t lo = 0, hi = f.draws.size() - 1;
            while (lo < hi) { size_t mid = (lo + hi + 1) / 2; if (f.draws[mid].prim_base <= seq) lo = mid; else hi = mid - 1; }
            const Draw& d = f.draws[lo];
            uint32_t t = seq - d.prim_base;
            VSOut vs[3]; uint32_t vid[3]; const float* pos[3];
            fetch_vs(f, d, t, vs, vid, pos);
            PrimSetup ps; setup_prim(d.mvp, pos[0], pos[1], pos[2], W, H, ps);
            SubTri st; bool found = false;
            for (int k = 1; k + 1 < ps.n && !found; ++k) {
                if (!make_subtri(ps, k, st)) continue;
                int64_t w0, w1, w2; subtri_weights(st, px, py, w0, w1, w2);
                if (subtri_covers(st, w0, w1, w2)) found = true;
            }
            if (!found) continue;  // cannot happen: the key came from this primitive
            FragIn in; float bary[3];
            interpolate(f, d, st, vs, px, py, in, bary);
            FragOut o; fragment_stage(f, d, in, o);
            for (int k = 0; k < 4; ++k) { hdr[p * 4 + k] = o.color[k]; coord[p * 4 + k] = o.objc[k]; normals[p * 4 + k] = o.normal[k]; camc[p * 4 + k] = o.camc[k]; }
            cls[p] = (uint16_t)o.cls; inst[p] = (uint16_t)o.inst;
            for (int k = 0; k < 3; ++k) { vidx[p * 4 + k] = vid[k]; bary_t[p * 4 + k] = bary[k]; }
            vidx[p * 4 + 3] = 0;          // uvec3 output into RGBA32UI: 4th component undefined in GL; 0 here
            bary_t[p * 4 + 3] = 1.0f;     // vec3 output into RGBA32F: missing alpha is 1
        }

    // ---- auto-exposure average: 1x1 level of the mip chain, taken BEFORE background/SSAO ----
    float avg[4] = {0, 0, 0, 0};
    if (sc.manual_exposure < 0) {
        std::vector<float> lvl = hdr; int lw = W, lh = H;
        while (lw > 1 || lh > 1) { int dw, dh; lvl = downsample(lvl, lw, lh, dw, dh); lw = dw; lh = dh; }
        for (int k = 0; k < 4; ++k) avg[k] = lvl[k];
    }

    // ---- background (render_pass.cpp:637-660) ----
    const Texture* bg = (const Texture*)sc.background_image;
    if (bg) {
        // Full-screen quad at z_ndc = 0 (gl_Position.xywz = (pos,1,0)) with depth test LESS still
        // enabled: it replaces colour attachment 0 wherever the stored depth is > 0.5 — i.e. also
        // over geometry farther than ~0.2 m. Literal restatement of the reference.
        const uint32_t quad_d24 = (uint32_t)std::lrintf(0.5f * 16777215.0f);
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                size_t p = (size_t)py * W + px;
                uint32_t d24 = (keys[p] == ~0ull) ? 0xFFFFFFu : (uint32_t)(keys[p] >> 32);
                if (!(quad_d24 < d24)) continue;
                V4 c = background_image_texel(*bg, px, py, W, H);
                hdr[p * 4 + 0] = c.x; hdr[p * 4 + 1] = c.y; hdr[p * 4 + 2] = c.z; hdr[p * 4 + 3] = 0.0f;
            }
    } else if (f.lm) {
        // skybox at depth == far, LEQUAL: fills pixels the geometry did not touch
        M4 Pinv = inverted(f.P);
        #pragma omp parallel for num_threads(n_threads)
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                size_t p = (size_t)py * W + px;
                if (keys[p] != ~0ull) continue;
                V3 dw = skybox_dir(Pinv, f.V, 2.0f * (px + 0.5f) / W - 1.0f, 2.0f * (py + 0.5f) / H - 1.0f);
                V4 c = sample_cube_lod(f.lm->env, dw, 0.0f);
                hdr[p * 4 + 0] = c.x; hdr[p * 4 + 1] = c.y; hdr[p * 4 + 2] = c.z; hdr[p * 4 + 3] = 0.0f;
            }
    }

    // ---- SSAO (render_pass.cpp:662-694) ----
    if (sc.ssao_enabled) {
        std::vector<float> ao((size_t)W * H, 1.0f);
        ssao_pass(camc.data(), normals.data(), W, H, f.P, ao.data(), n_threads);
        std::vector<float> outc = hdr;
        ssao_apply_pass(hdr.data(), ao.data(), camc.data(), W, H, outc.data(), n_threads);
        hdr.swap(outc);
    }

    // ---- tone map (tone_map_shader.frag:102-131) ----
    std::vector<uint8_t> rgb((size_t)W * H * 4);
    for (size_t p = 0; p < (size_t)W * H; ++p) tone_map_pixel(&hdr[p * 4], sc.manual_exposure, avg, &rgb[p * 4]);

    if (out) {
        if (out[SLB_TARGET_RGB]) std::memcpy(out[SLB_TARGET_RGB], rgb.data(), rgb.size());
        if (out[SLB_TARGET_COORD]) std::memcpy(out[SLB_TARGET_COORD], coord.data(), coord.size() * 4);
        if (out[SLB_TARGET_CLASS]) std::memcpy(out[SLB_TARGET_CLASS], cls.data(), cls.size() * 2);
        if (out[SLB_TARGET_INSTANCE]) std::memcpy(out[SLB_TARGET_INSTANCE], inst.data(), inst.size() * 2);
        if (out[SLB_TARGET_NORMAL]) std::memcpy(out[SLB_TARGET_NORMAL], normals.data(), normals.size() * 4);
        if (out[SLB_TARGET_VERTEX_INDEX]) std::memcpy(out[SLB_TARGET_VERTEX_INDEX], vidx.data(), vidx.size() * 4);
        if (out[SLB_TARGET_BARY]) std::memcpy(out[SLB_TARGET_BARY], bary_t.data(), bary_t.size() * 4);
        if (out[SLB_TARGET_CAM_COORD]) std::memcpy(out[SLB_TARGET_CAM_COORD], camc.data(), camc.size() * 4);
    }
    if (hdr_out) std::memcpy(hdr_out, hdr.data(), hdr.size() * 4);
    return 0;
}

// =======================================================================================
// Per-stage test hooks (oracle/orc_test_hooks.h): the restatement's vertex / fragment / post stages on caller-supplied
// inputs, so that tests/test_glsl_ref.py can hold them against the reference's GLSL compiled as C++ (oracle/_ref/libglslref.so).
// =======================================================================================
void orc_test_vertex(const orc_vert_uniforms* u, const void* verts68, int n, orc_vert_out* out) {
    Frame f; Draw d;
    f.V = M4::from(u->world_to_cam); f.P = M4::from(u->projection);
    d.meshToObject = M4::from(u->mesh_to_object); d.objectToWorld = M4::from(u->object_to_world);
    std::memcpy(d.normalToWorld, u->normal_to_world, sizeof d.normalToWorld);
    d.stickerProj = M4::from(u->sticker_projection);
    for (int k = 0; k < 4; ++k) d.stickerRange[k] = u->sticker_range[k];
    const Vertex68* v = (const Vertex68*)verts68;
    for (int i = 0; i < n; ++i) {
        VSOut o; vertex_stage(f, d, v[i], o);
        orc_vert_out& r = out[i];
        std::memset(&r, 0, sizeof r);
        r.uv[0] = o.uv.x; r.uv[1] = o.uv.y;
        for (int k = 0; k < 3; ++k) { r.normal_w[k] = o.nW[k]; r.tangent_w[k] = o.tW[k]; r.bitangent_w[k] = o.bW[k]; r.world[k] = o.wc[k]; r.cam[k] = o.cc[k]; }
        for (int k = 0; k < 4; ++k) r.objc[k] = o.objc[k];
        r.sticker[0] = o.sticker.x; r.sticker[1] = o.sticker.y;
        r.vertex_id = v[i].vertex_index;
        // clip position as the raster front end computes it (contract C2/C3: mvp formed once, one fma chain per row)
        float mvp[16]; make_mvp(f.P, f.V, d.objectToWorld, d.meshToObject, mvp);
        ClipV c; xform_clip(mvp, v[i].pos, c);
        r.position[0] = c.x; r.position[1] = c.y; r.position[2] = c.z; r.position[3] = c.w;
    }
}

void orc_test_fragment(const orc_frag_uniforms* u, const orc_frag_in* in, int n, orc_frag_out* out) {
    Frame f; Draw d;
    f.W = u->width; f.H = u->height; f.peel = u->peel;
    f.V = M4::from(u->world_to_cam);
    f.camPos = V3(u->cam_position[0], u->cam_position[1], u->cam_position[2]);
    f.lm = u->light_map_available ? (const LightMap*)u->light_map : nullptr;
    for (int i = 0; i < SLB_NUM_LIGHTS; ++i) {
        f.lightDir[i] = V3(u->light_directions[3 * i], u->light_directions[3 * i + 1], u->light_directions[3 * i + 2]);
        f.lightCol[i] = V3(u->light_colors[3 * i], u->light_colors[3 * i + 1], u->light_colors[3 * i + 2]);
        bool colZero = f.lightCol[i].x == 0 && f.lightCol[i].y == 0 && f.lightCol[i].z == 0;
        bool dirZero = f.lightDir[i].x == 0 && f.lightDir[i].y == 0 && f.lightDir[i].z == 0;
        f.lightActive[i] = !(colZero || dirZero);
        f.shadowMat[i] = M4::from(u->shadow_matrices + 16 * i);
        if (u->shadow_map[i]) f.shadowMap[i].assign(u->shadow_map[i], u->shadow_map[i] + (size_t)SLB_SHADOW_RES * SLB_SHADOW_RES);
        else f.shadowMap[i].assign((size_t)SLB_SHADOW_RES * SLB_SHADOW_RES, 0xFFFFFFu);
    }
    f.ambient = V3(u->ambient[0], u->ambient[1], u->ambient[2]);
    std::memcpy(d.mat.base_color, u->material, 16); std::memcpy(d.mat.emissive, u->material + 4, 16);
    d.mat.metallic = u->material[9]; d.mat.roughness = u->material[10];
    for (int i = 0; i < 5; ++i) d.tex[i] = (u->available_textures & (1u << i)) ? (const Texture*)u->tex[i] : nullptr;
    d.alpha_tested = (d.tex[0] && d.tex[0]->has_alpha) || d.mat.base_color[3] < 0.5f;
    d.class_index = u->class_index; d.instance_index = u->instance_index;
    d.sticker = (const Texture*)u->sticker;
    for (int i = 0; i < n; ++i) {
        const orc_frag_in& s = in[i];
        FragIn fi;
        fi.uv = V2{s.uv[0], s.uv[1]}; fi.uv_dx = V2{s.uv_dx[0], s.uv_dx[1]}; fi.uv_dy = V2{s.uv_dy[0], s.uv_dy[1]};
        fi.nW = V3(s.normal_w[0], s.normal_w[1], s.normal_w[2]); fi.tW = V3(s.tangent_w[0], s.tangent_w[1], s.tangent_w[2]);
        fi.bW = V3(s.bitangent_w[0], s.bitangent_w[1], s.bitangent_w[2]);
        fi.objc = V4(s.objc[0], s.objc[1], s.objc[2], s.objc[3]);
        fi.wc = V3(s.world[0], s.world[1], s.world[2]); fi.cc = V3(s.cam[0], s.cam[1], s.cam[2]);
        fi.sticker = V2{s.sticker[0], s.sticker[1]};
        fi.front = s.front_facing != 0;
        orc_frag_out& o = out[i];
        std::memset(&o, 0, sizeof o);
        // the two discards, evaluated where the raster stage of orc_render evaluates them (render_shader.frag:229-246)
        const int px = (int)s.frag_x, py = (int)s.frag_y;
        const float prev = f.peel ? f.peel[((size_t)py * f.W + px) * 4 + 3] : 0.0f;   // no peel layer bound: a zero texture (render_pass.cpp:395-402)
        if (fi.objc.w - 0.00001f <= prev) { o.discarded = 1; continue; }
        if (d.alpha_tested && base_color(d, fi).w < 0.5f) { o.discarded = 1; continue; }
        FragOut fo; fragment_stage(f, d, fi, fo);
        for (int k = 0; k < 4; ++k) { o.color[k] = fo.color[k]; o.objc[k] = fo.objc[k]; o.camc[k] = fo.camc[k]; o.normal[k] = fo.normal[k]; }
        o.class_index = fo.cls; o.instance_index = fo.inst;
        for (int k = 0; k < 3; ++k) { o.vertex_ids[k] = s.vertex_ids[k]; o.bary[k] = s.bary[k]; }   // flat / smooth pass-through (render_shader.geom)
    }
}

void orc_test_tonemap(const float* hdr, int n, float manual_exposure, const float* avg, uint8_t* rgba8) {
    for (int i = 0; i < n; ++i) tone_map_pixel(hdr + 4 * (size_t)i, manual_exposure, avg, rgba8 + 4 * (size_t)i);
}
void orc_test_ssao(const float* camc, const float* normals, int W, int H, const float* projection, float* ao) {
    ssao_pass(camc, normals, W, H, M4::from(projection), ao, omp_get_max_threads());
}
void orc_test_ssao_apply(const float* hdr, const float* ao, const float* camc, int W, int H, float* out) {
    ssao_apply_pass(hdr, ao, camc, W, H, out, omp_get_max_threads());
}
void orc_test_ssao_tables(float* noise16x3, float* kernel64x3) {
    V3 noise[16], kernel[64]; ssao_tables(noise, kernel);
    for (int i = 0; i < 16; ++i) for (int k = 0; k < 3; ++k) noise16x3[3 * i + k] = noise[i][k];
    for (int i = 0; i < 64; ++i) for (int k = 0; k < 3; ++k) kernel64x3[3 * i + k] = kernel[i][k];
}
void orc_test_skybox_dir(const float* projection, const float* world_to_cam, const float* ndc_xy, int n, float* dir) {
    M4 Pinv = inverted(M4::from(projection)), V = M4::from(world_to_cam);
    for (int i = 0; i < n; ++i) { V3 d = skybox_dir(Pinv, V, ndc_xy[2 * i], ndc_xy[2 * i + 1]); dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z; }
}
// frustum corners + shadow matrix of one scene / light (render_pass.cpp:69-211) from plain arrays (column-major matrices):
// the same signature as ref_shadow_setup of oracle/ref_frame_harness.cpp, which runs the reference's own two functions
void orc_test_shadow_setup(const float* projection, const float* world_to_cam, int n_objects, const float* poses, const float* pretransforms,
                           const float* bbox_min, const float* bbox_max, const float* light_dir, float* corners_out, float* shadow_out) {
    std::vector<Mesh> meshes(n_objects);
    std::vector<slb_object_desc> objs(n_objects);
    for (int i = 0; i < n_objects; ++i) {
        std::memcpy(meshes[i].bbox_min, bbox_min + 3 * i, 12); std::memcpy(meshes[i].bbox_max, bbox_max + 3 * i, 12);
        std::memset(&objs[i], 0, sizeof objs[i]);
        objs[i].mesh = (const slb_mesh*)&meshes[i];
        std::memcpy(objs[i].pose, poses + 16 * i, 64); std::memcpy(objs[i].pretransform, pretransforms + 16 * i, 64);
    }
    slb_scene_desc sc; std::memset(&sc, 0, sizeof sc);
    std::memcpy(sc.projection, projection, 64); std::memcpy(sc.world_to_cam, world_to_cam, 64);
    sc.objects = objs.data(); sc.n_objects = n_objects;
    V3 corners[8];
    frustum_corners(sc, corners);
    for (int i = 0; i < 8; ++i) { corners_out[3 * i] = corners[i].x; corners_out[3 * i + 1] = corners[i].y; corners_out[3 * i + 2] = corners[i].z; }
    const M4 sm = shadow_matrix(sc, corners, V3(light_dir[0], light_dir[1], light_dir[2]));
    std::memcpy(shadow_out, sm.m, 64);
}
void orc_test_background_image(const void* tex, int W, int H, float* out /* HxWx4 */) {
    const Texture& bg = *(const Texture*)tex;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            V4 c = background_image_texel(bg, px, py, W, H);
            float* o = out + ((size_t)py * W + px) * 4; o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = 0.0f;
        }
}

}  // extern "C"
