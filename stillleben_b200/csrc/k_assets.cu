// k_assets.cu — load-time kernels:
//   vertex repack + texture upload / mip chains  replaces Mesh::loadVisual (reference: src/mesh.cpp:624-745)
//   IBL precompute                               replaces LightMap::load's GL passes (reference:
//       src/light_map.cpp:376-611, src/shaders/cubemap_shader_{equirectangular,irradiance,prefilter}.frag,
//       src/shaders/brdf_shader.frag)
#include <cuda_runtime.h>
#include <stdint.h>

#include "k_frag.cuh"
#include "kernels.h"

using namespace slbk;

// 68-byte interleaved record {pos3, uv2, color4, tangent4, id, normal3} -> pos4[] + attr[3]
__global__ void k_repack(const uint8_t* __restrict__ v68, uint32_t n, float4* __restrict__ pos4, float4* __restrict__ attr,
                         float4* __restrict__ col4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* v = reinterpret_cast<const float*>(v68 + (size_t)i * SLB_VERTEX_STRIDE);   // 68 = 17 words: 4-byte aligned
    pos4[i] = make_float4(v[0], v[1], v[2], v[13]);                                         // v[13] = vertex id bits
    attr[3 * (size_t)i + 0] = make_float4(v[3], v[4], v[14], v[15]);                        // u, v, nx, ny
    attr[3 * (size_t)i + 1] = make_float4(v[16], v[9], v[10], v[11]);                       // nz, tx, ty, tz
    attr[3 * (size_t)i + 2] = make_float4(v[12], 0.f, 0.f, 0.f);                            // tw
    col4[i] = make_float4(v[5], v[6], v[7], v[8]);   // vertex colours: never read by the renderer, kept for Mesh::colors / update_colors
}
// the inverse: pos4[] + attr[] + col4[] -> the 68-byte record (read-back for Mesh::points / normals / colors accessors)
__global__ void k_unpack(const float4* __restrict__ pos4, const float4* __restrict__ attr, const float4* __restrict__ col4, uint32_t n,
                         uint8_t* __restrict__ v68) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* v = reinterpret_cast<float*>(v68 + (size_t)i * SLB_VERTEX_STRIDE);
    const float4 p = pos4[i], a0 = attr[3 * (size_t)i], a1 = attr[3 * (size_t)i + 1], a2 = attr[3 * (size_t)i + 2], c = col4[i];
    v[0] = p.x; v[1] = p.y; v[2] = p.z; v[13] = p.w;
    v[3] = a0.x; v[4] = a0.y; v[14] = a0.z; v[15] = a0.w;
    v[16] = a1.x; v[9] = a1.y; v[10] = a1.z; v[11] = a1.w; v[12] = a2.x;
    v[5] = c.x; v[6] = c.y; v[7] = c.z; v[8] = c.w;
}
// largest index value of an index buffer (upload-time validation: an index >= n_vertices would be an out-of-bounds read in
// every kernel that dereferences pos4[idx])
__global__ void k_index_max(const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, idx[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ---- vertex edit path (reference: Mesh::updateVertexPositionsAndColors / setVertexPositions / recomputeNormals,
// src/mesh.cpp:763-878) ----
// point[id-1] += update (ids are one-based vertex ids, mesh.cpp:830-835); colour likewise (:845-850). err: out-of-range id seen.
__global__ void k_vertex_delta(const int32_t* __restrict__ ids, uint32_t n, const float* __restrict__ dpos, const float* __restrict__ dcol,
                               float4* __restrict__ pos4, float4* __restrict__ col4, uint32_t n_vertices, uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = ids[i] - 1;
    if (v < 0 || (uint32_t)v >= n_vertices) { atomicExch(err, 1u); return; }
    if (dpos) {   // float atomics: a vertex listed twice accumulates both updates (the reference does so sequentially)
        atomicAdd(&pos4[v].x, dpos[3 * (size_t)i]); atomicAdd(&pos4[v].y, dpos[3 * (size_t)i + 1]); atomicAdd(&pos4[v].z, dpos[3 * (size_t)i + 2]);
    }
    if (dcol) {
        atomicAdd(&col4[v].x, dcol[4 * (size_t)i]); atomicAdd(&col4[v].y, dcol[4 * (size_t)i + 1]);
        atomicAdd(&col4[v].z, dcol[4 * (size_t)i + 2]); atomicAdd(&col4[v].w, dcol[4 * (size_t)i + 3]);
    }
}
__global__ void k_set_positions(const float* __restrict__ p3, uint32_t n, float4* __restrict__ pos4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos4[i].x = p3[3 * (size_t)i]; pos4[i].y = p3[3 * (size_t)i + 1]; pos4[i].z = p3[3 * (size_t)i + 2];
}
__global__ void k_set_colors(const float* __restrict__ c4, uint32_t n, float4* __restrict__ col4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    col4[i] = make_float4(c4[4 * (size_t)i], c4[4 * (size_t)i + 1], c4[4 * (size_t)i + 2], c4[4 * (size_t)i + 3]);
}
// recomputeNormals pass 1 (mesh.cpp:776-803): per face, cross = (v1 - v2) x (v1 - v3), area = |cross|, normal = cross.normalized();
// stored: normal * area. Explicit round-to-nearest intrinsics (no fma contraction): bit-identical with the oracle's twin,
// incl. the NaN a zero-area face produces (0 * inf), which the reference propagates into the vertex normal.
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
__global__ void k_face_normals(const float4* __restrict__ pos4, const uint32_t* __restrict__ idx, uint32_t n_faces, float4* __restrict__ face_n) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const float4 v1 = pos4[idx[3 * (size_t)f]], v2 = pos4[idx[3 * (size_t)f + 1]], v3 = pos4[idx[3 * (size_t)f + 2]];
    const float ax = __fsub_rn(v1.x, v2.x), ay = __fsub_rn(v1.y, v2.y), az = __fsub_rn(v1.z, v2.z);
    const float bx = __fsub_rn(v1.x, v3.x), by = __fsub_rn(v1.y, v3.y), bz = __fsub_rn(v1.z, v3.z);
    const float cx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(by, az)), cy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(bz, ax)),
                cz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(bx, ay));
    const float area = __fsqrt_rn(dot3_rn(cx, cy, cz, cx, cy, cz));
    const float inv = __fdiv_rn(1.0f, area);                     // Vector::normalized() = *this * lengthInverted()
    face_n[f] = make_float4(__fmul_rn(__fmul_rn(cx, inv), area), __fmul_rn(__fmul_rn(cy, inv), area), __fmul_rn(__fmul_rn(cz, inv), area), 0.f);
}
// pass 2 (mesh.cpp:805-818): per vertex, sum of its faces' (normal * area) in ascending face order, normalised; written into the
// normal slot of the attribute stream
__global__ void k_vertex_normals(const uint32_t* __restrict__ adj_off, const uint32_t* __restrict__ adj_face, const float4* __restrict__ face_n,
                                 uint32_t n_vertices, float4* __restrict__ attr) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vertices) return;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (uint32_t k = adj_off[v]; k < adj_off[v + 1]; ++k) {
        const float4 fn = face_n[adj_face[k]];
        nx = __fadd_rn(nx, fn.x); ny = __fadd_rn(ny, fn.y); nz = __fadd_rn(nz, fn.z);
    }
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot3_rn(nx, ny, nz, nx, ny, nz)));
    nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
    attr[3 * (size_t)v].z = nx; attr[3 * (size_t)v].w = ny; attr[3 * (size_t)v + 1].x = nz;
}

__global__ void k_expand_rgba(const uint8_t* __restrict__ src, int channels, uint8_t* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uchar4 o;
    o.x = src[i * channels]; o.y = src[i * channels + 1]; o.z = src[i * channels + 2];
    o.w = channels == 4 ? src[i * 4 + 3] : 255;
    reinterpret_cast<uchar4*>(dst)[i] = o;
}

// integer box filter with round-half-up; even sizes: 2x2 box, odd sizes: polyphase box
__device__ __forceinline__ void itaps(int s, int d, int i, int idx[3], int wgt[3], int& total) {
    if (s == 1) { idx[0] = idx[1] = idx[2] = 0; wgt[0] = 1; wgt[1] = 0; wgt[2] = 0; total = 1; return; }
    if ((s & 1) == 0) { idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 1; wgt[0] = 1; wgt[1] = 1; wgt[2] = 0; total = 2; return; }
    idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 2;
    wgt[0] = d - i; wgt[1] = d; wgt[2] = i + 1; total = s;
}
__global__ void k_mip_level(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int ix[3], wx[3], tx, iy[3], wy[3], ty;
    itaps(sw, dw, x, ix, wx, tx); itaps(sh, dh, y, iy, wy, ty);
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) {
            uint32_t w = (uint32_t)(wy[b] * wx[a]);
            uchar4 p = reinterpret_cast<const uchar4*>(src)[(size_t)iy[b] * sw + ix[a]];
            acc[0] += w * p.x; acc[1] += w * p.y; acc[2] += w * p.z; acc[3] += w * p.w;
        }
    uint32_t tot = (uint32_t)(tx * ty);
    reinterpret_cast<uchar4*>(dst)[(size_t)y * dw + x] =
        make_uchar4((acc[0] + tot / 2) / tot, (acc[1] + tot / 2) / tot, (acc[2] + tot / 2) / tot, (acc[3] + tot / 2) / tot);
}

__global__ void k_cube_mip(const float4* __restrict__ src, int ssize, float4* __restrict__ dst) {
    int dsize = ssize / 2;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= dsize || y >= dsize) return;
    auto at = [&](int xx, int yy) { return src[((size_t)f * ssize + yy) * ssize + xx]; };
    float4 a = at(2 * x, 2 * y), b = at(2 * x + 1, 2 * y), c = at(2 * x, 2 * y + 1), d = at(2 * x + 1, 2 * y + 1);
    float4 o;
    o.x = 0.25f * ((a.x + b.x) + (c.x + d.x)); o.y = 0.25f * ((a.y + b.y) + (c.y + d.y));
    o.z = 0.25f * ((a.z + b.z) + (c.z + d.z)); o.w = 0.25f * ((a.w + b.w) + (c.w + d.w));
    dst[((size_t)f * dsize + y) * dsize + x] = o;
}

// ---- IBL precompute -----------------------------------------------------------------------
#define SLB_PI_GLSL 3.14159265359f

__device__ __forceinline__ f3 equirect_sample(const float* __restrict__ img, int W, int H, float u, float v) {
    float x = u * W - 0.5f, y = v * H - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    auto at = [&](int i, int j) {
        i = min(max(i, 0), W - 1); j = min(max(j, 0), H - 1);
        const float* p = img + ((size_t)j * W + i) * 3; return mk3(p[0], p[1], p[2]);
    };
    return at(i0, j0) * ((1 - a) * (1 - b)) + at(i0 + 1, j0) * (a * (1 - b)) + at(i0, j0 + 1) * ((1 - a) * b) + at(i0 + 1, j0 + 1) * (a * b);
}
__global__ void k_equirect_to_cube(const float* __restrict__ eq, int ew, int eh, float4* __restrict__ cube, int size) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= size || y >= size) return;
    f3 v = normalize3(cube_face_dir(f, (x + 0.5f) / size, (y + 0.5f) / size));
    float u = atan2f(v.y, v.x) * 0.1591f + 0.5f;
    float w = asinf(v.z) * 0.3183f + 0.5f;
    f3 c = equirect_sample(eq, ew, eh, u, w);
    cube[((size_t)f * size + y) * size + x] = make_float4(c.x, c.y, c.z, 1.0f);
}

__global__ void k_irradiance(const DLightMap* __restrict__ lm, float4* __restrict__ out, int size, float lod) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= size || y >= size) return;
    f3 N = normalize3(cube_face_dir(f, (x + 0.5f) / size, (y + 0.5f) / size));
    f3 irradiance = mk3(0.f, 0.f, 0.f);
    f3 up = mk3(0.f, 1.f, 0.f);
    f3 right = cross3(up, N);
    up = cross3(N, right);
    const float sampleDelta = 0.020f;
    float nrSamples = 0.0f;
    for (float phi = 0.0f; phi < 2.0f * SLB_PI_GLSL; phi += sampleDelta)
        for (float theta = 0.0f; theta < 0.5f * SLB_PI_GLSL; theta += sampleDelta) {
            f3 ts = mk3(sinf(theta) * cosf(phi), sinf(theta) * sinf(phi), cosf(theta));
            f3 sv = right * ts.x + up * ts.y + N * ts.z;
            float4 c = cube_sample_lod(lm->env, lm->n_env, sv, lod);
            irradiance = irradiance + mk3(c.x, c.y, c.z) * (cosf(theta) * sinf(theta));
            nrSamples += 1.0f;
        }
    irradiance = irradiance * SLB_PI_GLSL * (1.0f / nrSamples);
    out[((size_t)f * size + y) * size + x] = make_float4(irradiance.x, irradiance.y, irradiance.z, 1.0f);
}

__device__ __forceinline__ float radical_inverse(uint32_t bits) { return (float)__brev(bits) * 2.3283064365386963e-10f; }
__device__ __forceinline__ f3 importance_sample_ggx(float xi_x, float xi_y, f3 N, float roughness) {
    float a = roughness * roughness;
    float phi = 2.0f * SLB_PI_GLSL * xi_x;
    float cosTheta = sqrtf((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    f3 H = mk3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    f3 up = fabsf(N.z) < 0.999f ? mk3(0.f, 0.f, 1.f) : mk3(1.f, 0.f, 0.f);
    f3 tangent = normalize3(cross3(up, N));
    f3 bitangent = cross3(N, tangent);
    return normalize3(tangent * H.x + bitangent * H.y + N * H.z);
}
__device__ __forceinline__ float distribution_ggx_ibl(f3 N, f3 H, float roughness) {
    float a = roughness * roughness, a2 = a * a;
    float NdotH = fmaxf(dot3(N, H), 0.0f), NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = SLB_PI_GLSL * denom * denom;
    return a2 / denom;
}
__global__ void k_prefilter(const DLightMap* __restrict__ lm, float4* __restrict__ out, int size, float roughness, int n_samples,
                            float resolution) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, f = blockIdx.z;
    if (x >= size || y >= size) return;
    f3 N = normalize3(cube_face_dir(f, (x + 0.5f) / size, (y + 0.5f) / size));
    f3 V = N;
    f3 color = mk3(0.f, 0.f, 0.f);
    float totalWeight = 0.0f;
    for (uint32_t i = 0; i < (uint32_t)n_samples; ++i) {
        float xi_x = (float)i / (float)n_samples, xi_y = radical_inverse(i);
        f3 H = importance_sample_ggx(xi_x, xi_y, N, roughness);
        f3 L = normalize3(H * (2.0f * dot3(V, H)) - V);
        float NdotL = fmaxf(dot3(N, L), 0.0f);
        if (NdotL > 0.0f) {
            float D = distribution_ggx_ibl(N, H, roughness);
            float NdotH = fmaxf(dot3(N, H), 0.0f), HdotV = fmaxf(dot3(H, V), 0.0f);
            float pdf = D * NdotH / (4.0f * HdotV) + 0.0001f;
            float saTexel = 4.0f * SLB_PI_GLSL / (6.0f * resolution * resolution);
            float saSample = 1.0f / ((float)n_samples * pdf + 0.0001f);
            float mipLevel = roughness == 0.0f ? 0.0f : 0.5f * log2f(saSample / saTexel);
            float4 c = cube_sample_lod(lm->env, lm->n_env, L, mipLevel);
            color = color + mk3(c.x, c.y, c.z) * NdotL;
            totalWeight += NdotL;
        }
    }
    color = color / totalWeight;
    out[((size_t)f * size + y) * size + x] = make_float4(color.x, color.y, color.z, 1.0f);
}
__device__ __forceinline__ float geometry_schlick_ibl(float NdotV, float roughness) {
    float k = (roughness * roughness) / 2.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
__global__ void k_brdf_lut(float4* __restrict__ out, int size, int n_samples) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= size || y >= size) return;
    float NdotV = (x + 0.5f) / size, roughness = (y + 0.5f) / size;
    f3 V = mk3(sqrtf(1.0f - NdotV * NdotV), 0.0f, NdotV);
    f3 N = mk3(0.f, 0.f, 1.f);
    float A = 0.0f, B = 0.0f;
    for (uint32_t i = 0; i < (uint32_t)n_samples; ++i) {
        float xi_x = (float)i / (float)n_samples, xi_y = radical_inverse(i);
        f3 H = importance_sample_ggx(xi_x, xi_y, N, roughness);
        f3 L = normalize3(H * (2.0f * dot3(V, H)) - V);
        float NdotL = fmaxf(L.z, 0.0f), NdotH = fmaxf(H.z, 0.0f), VdotH = fmaxf(dot3(V, H), 0.0f);
        if (NdotL > 0.0f) {
            float G = geometry_schlick_ibl(fmaxf(dot3(N, L), 0.0f), roughness) * geometry_schlick_ibl(fmaxf(dot3(N, V), 0.0f), roughness);
            float G_Vis = (G * VdotH) / (NdotH * NdotV);
            float Fc = powf(1.0f - VdotH, 5.0f);
            A += (1.0f - Fc) * G_Vis;
            B += Fc * G_Vis;
        }
    }
    out[(size_t)y * size + x] = make_float4(A / (float)n_samples, B / (float)n_samples, 0.0f, 1.0f);
}

namespace slbk {

void launch_repack_vertices(const uint8_t* verts68, uint32_t n, float4* pos4, float4* attr, float4* col4, cudaStream_t s) {
    if (n) k_repack<<<(n + 255) / 256, 256, 0, s>>>(verts68, n, pos4, attr, col4);
}
void launch_unpack_vertices(const float4* pos4, const float4* attr, const float4* col4, uint32_t n, uint8_t* verts68, cudaStream_t s) {
    if (n) k_unpack<<<(n + 255) / 256, 256, 0, s>>>(pos4, attr, col4, n, verts68);
}
void launch_index_max(const uint32_t* idx, uint32_t n, uint32_t* out, cudaStream_t s) {
    if (n) k_index_max<<<min((n + 255) / 256, 1184u), 256, 0, s>>>(idx, n, out);
}
void launch_vertex_delta(const int32_t* ids, uint32_t n, const float* dpos, const float* dcol, float4* pos4, float4* col4, uint32_t n_vertices,
                         uint32_t* err, cudaStream_t s) {
    if (n) k_vertex_delta<<<(n + 255) / 256, 256, 0, s>>>(ids, n, dpos, dcol, pos4, col4, n_vertices, err);
}
void launch_set_positions(const float* p3, uint32_t n, float4* pos4, cudaStream_t s) {
    if (n) k_set_positions<<<(n + 255) / 256, 256, 0, s>>>(p3, n, pos4);
}
void launch_set_colors(const float* c4, uint32_t n, float4* col4, cudaStream_t s) {
    if (n) k_set_colors<<<(n + 255) / 256, 256, 0, s>>>(c4, n, col4);
}
void launch_recompute_normals(const float4* pos4, const uint32_t* idx, uint32_t n_faces, const uint32_t* adj_off, const uint32_t* adj_face,
                              float4* face_n, uint32_t n_vertices, float4* attr, cudaStream_t s) {
    if (n_faces) k_face_normals<<<(n_faces + 255) / 256, 256, 0, s>>>(pos4, idx, n_faces, face_n);
    if (n_vertices) k_vertex_normals<<<(n_vertices + 255) / 256, 256, 0, s>>>(adj_off, adj_face, face_n, n_vertices, attr);
}
void launch_expand_rgba(const uint8_t* src, int channels, uint8_t* dst, size_t n_texels, cudaStream_t s) {
    if (n_texels) k_expand_rgba<<<(unsigned)((n_texels + 255) / 256), 256, 0, s>>>(src, channels, dst, n_texels);
}
void launch_mip_level(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, cudaStream_t s) {
    dim3 block(16, 16), grid((dw + 15) / 16, (dh + 15) / 16);
    k_mip_level<<<grid, block, 0, s>>>(src, sw, sh, dst, dw, dh);
}
void launch_cube_mip(const float4* src, int ssize, float4* dst, cudaStream_t s) {
    int d = ssize / 2;
    dim3 block(16, 16), grid((d + 15) / 16, (d + 15) / 16, 6);
    k_cube_mip<<<grid, block, 0, s>>>(src, ssize, dst);
}
void launch_equirect_to_cube(const float* equirect, int ew, int eh, float4* cube, int size, cudaStream_t s) {
    dim3 block(16, 16), grid((size + 15) / 16, (size + 15) / 16, 6);
    k_equirect_to_cube<<<grid, block, 0, s>>>(equirect, ew, eh, cube, size);
}
void launch_irradiance(const DLightMap* lm, float4* out, int size, float lod, cudaStream_t s) {
    dim3 block(8, 8), grid((size + 7) / 8, (size + 7) / 8, 6);
    k_irradiance<<<grid, block, 0, s>>>(lm, out, size, lod);
}
void launch_prefilter(const DLightMap* lm, float4* out, int size, float roughness, int n_samples, float env_resolution, cudaStream_t s) {
    dim3 block(8, 8), grid((size + 7) / 8, (size + 7) / 8, 6);
    k_prefilter<<<grid, block, 0, s>>>(lm, out, size, roughness, n_samples, env_resolution);
}
void launch_brdf_lut(float4* out, int size, int n_samples, cudaStream_t s) {
    dim3 block(16, 16), grid((size + 15) / 16, (size + 15) / 16);
    k_brdf_lut<<<grid, block, 0, s>>>(out, size, n_samples);
}

}  // namespace slbk
