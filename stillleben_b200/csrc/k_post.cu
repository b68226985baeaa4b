// k_post.cu — full-screen passes that follow the geometry pass when they are enabled:
//   background image / IBL sky box     reference: src/render_pass.cpp:637-660, src/shaders/background_*.{vert,frag}
//   auto-exposure average (mip chain)  reference: src/render_pass.cpp:632-635, tone_map_shader.frag:114-122
//   SSAO + bilateral blur/apply        reference: src/render_pass.cpp:662-694, ssao_shader.frag:20-57, ssao_apply_shader.frag:29-76
//   tone map                           reference: src/render_pass.cpp:696-710, tone_map_shader.frag:102-131
// When none of them is needed the shade kernel tone-maps in registers and these kernels never run.
#include <cuda_runtime.h>
#include <stdint.h>

#include "k_frag.cuh"
#include <type_traits>

#include "kernels.h"

using namespace slbk;

__constant__ float c_ssao_noise[16 * 3];
__constant__ float c_ssao_kernel[64 * 3];

// linear-filtered rectangle-texture read of the z component of the camCoordinates image (clamp-to-edge), served
// from the dense z plane the shade kernel writes next to it
__device__ __forceinline__ float rect_linear_z(const float* __restrict__ img, int W, int H, float xs, float ys) {
    float x = xs - 0.5f, y = ys - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx, j0 = (int)fy;
    // clamp-to-edge: min(max(i, 0), n - 1) as ONE instruction each (VIMNMX with the relu modifier)
    const int i1 = __vimin_s32_relu(i0 + 1, W - 1), j1 = __vimin_s32_relu(j0 + 1, H - 1);
    i0 = __vimin_s32_relu(i0, W - 1); j0 = __vimin_s32_relu(j0, H - 1);
    // texel offsets as UNSIGNED 32-bit values (a frame is far below 2^32 pixels): one IMAD.WIDE.U32 per address instead of a
    // sign-extended 64-bit add chain
    const unsigned r0 = (unsigned)(j0 * W), r1 = (unsigned)(j1 * W);
    float z00 = __ldg(img + (r0 + (unsigned)i0)), z10 = __ldg(img + (r0 + (unsigned)i1));
    float z01 = __ldg(img + (r1 + (unsigned)i0)), z11 = __ldg(img + (r1 + (unsigned)i1));
    return z00 * ((1 - a) * (1 - b)) + z10 * (a * (1 - b)) + z01 * ((1 - a) * b) + z11 * (a * b);
}

__global__ void __launch_bounds__(256) k_background(const DFrame* __restrict__ frames) {
    const DFrame& f = frames[blockIdx.z];
    const int W = f.W, H = f.H;
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= W || py >= H) return;
    const size_t p = (size_t)py * W + px;
    const unsigned long long key = f.keys[p];
    if (f.bg_image) {
        // full-screen quad at z_ndc = 0 under depth func LESS: wins wherever stored depth24 > d24(0.5)
        const uint32_t quad_d24 = __float2uint_rn(0.5f * 16777215.0f);
        uint32_t d24 = (key == SLB_KEY_EMPTY) ? 0xFFFFFFu : (uint32_t)(key >> 40);
        if (!(quad_d24 < d24)) return;
        const DTexture& bg = *f.bg_image;
        // exact division / multiplication: the texel index is a floor of this value
        float tx = __fdiv_rn(px + 0.5f, (float)W), ty = __fsub_rn(1.0f, __fdiv_rn(py + 0.5f, (float)H));
        int ix = (int)__fmul_rn(tx, (float)bg.w), iy = (int)__fmul_rn(ty, (float)bg.h);
        float4 c = tex_sample_rect(bg, (float)ix, (float)iy);
        f.hdr[p] = make_float4(c.x, c.y, c.z, 0.0f);
    } else if (f.lm) {
        if (key != SLB_KEY_EMPTY) return;
        float xn = 2.0f * (px + 0.5f) / W - 1.0f, yn = 2.0f * (py + 0.5f) / H - 1.0f;
        float4 q = mul_m4_p(f.Pinv, xn, yn, 1.0f, 1.0f);
        f3 dc = mk3(q.x / q.w, q.y / q.w, q.z / q.w);
        f3 dw = mk3(f.V[0] * dc.x + f.V[1] * dc.y + f.V[2] * dc.z, f.V[4] * dc.x + f.V[5] * dc.y + f.V[6] * dc.z,
                    f.V[8] * dc.x + f.V[9] * dc.y + f.V[10] * dc.z);
        float4 c = cube_sample_lod(f.lm->env, f.lm->n_env, dw, 0.0f);
        f.hdr[p] = make_float4(c.x, c.y, c.z, 0.0f);
    }
}

// one glGenerateMipmap step of an RGBA32F image: 2x2 box for even sizes, polyphase box for odd sizes
__device__ __forceinline__ void mip_taps(int sN, int dN, int i, int idx[3], float w[3]) {
    if (sN == 1) { idx[0] = idx[1] = idx[2] = 0; w[0] = 1; w[1] = w[2] = 0; return; }
    if ((sN & 1) == 0) { idx[0] = 2 * i; idx[1] = idx[2] = 2 * i + 1; w[0] = w[1] = 0.5f; w[2] = 0; return; }
    idx[0] = 2 * i; idx[1] = 2 * i + 1; idx[2] = 2 * i + 2;
    w[0] = (float)(dN - i) / sN; w[1] = (float)dN / sN; w[2] = (float)(i + 1) / sN;
}
__global__ void k_downsample(const float4* __restrict__ src, int sw, int sh, float4* __restrict__ dst, int dw, int dh, size_t src_stride,
                             size_t dst_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float4* s = src + blockIdx.z * src_stride;
    int ix[3], iy[3]; float wx[3], wy[3];
    mip_taps(sw, dw, x, ix, wx); mip_taps(sh, dh, y, iy, wy);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) {
            double w = (double)__fmul_rn(wy[b], wx[a]);
            float4 v = s[(size_t)iy[b] * sw + ix[a]];
            a0 += w * v.x; a1 += w * v.y; a2 += w * v.z; a3 += w * v.w;
        }
    dst[blockIdx.z * dst_stride + (size_t)y * dw + x] = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
}

// 1 / y as ONE MUFU.RCP (flush-to-zero approximate reciprocal, the instruction the compiler's fast division is built around,
// without its range fix-up for denormal or > 2^126 operands: there the SSAO results are clamped / irrelevant)
__device__ __forceinline__ float rcp_fast(float y) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y)); return r; }
__device__ __forceinline__ float smoothstep01(float x) { float t = clampf(x, 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }

__global__ void __launch_bounds__(256) k_ssao(const DFrame* __restrict__ frames) {
    const DFrame& f = frames[blockIdx.z];
    const int W = f.W, H = f.H;
    // a warp covers an 8x4 pixel block: the 64 taps of neighbouring pixels land in neighbouring texels
    const int wq = threadIdx.x >> 5, lq = threadIdx.x & 31;
    const int px = blockIdx.x * 32 + (wq & 3) * 8 + (lq & 7), py = blockIdx.y * 8 + (wq >> 2) * 4 + (lq >> 3);
    if (px >= W || py >= H) return;
    const size_t p = (size_t)py * W + px;
    if (!f.ssao) { f.ao[p] = 1.0f; return; }
    float4 c4 = f.scratch_cam[p], n4 = f.scratch_normal[p];
    if (n4.x == 0 && n4.y == 0 && n4.z == 0) { f.ao[p] = 1.0f; return; }
    f3 fragPos = mk3(c4.x, c4.y, c4.z);
    f3 normal = normalize3(mk3(n4.x, n4.y, n4.z));
    const float* nz = c_ssao_noise + ((py & 3) * 4 + (px & 3)) * 3;
    f3 randomVec = normalize3(mk3(nz[0], nz[1], nz[2]));
    f3 tangent = normalize3(randomVec - normal * dot3(randomVec, normal));
    f3 bitangent = cross3(normal, tangent);
    // offset = P (fragPos + 0.1 TBN s) is linear in the kernel sample s: project the frame once, not every sample
    // (rows x, y, w of P; ssao_shader.frag:38-44)
    const float* P = f.P;
    float base[3], ct[3], cb[3], cn[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int row = r == 2 ? 3 : r;
        base[r] = P[row] * fragPos.x + P[4 + row] * fragPos.y + P[8 + row] * fragPos.z + P[12 + row];
        ct[r] = 0.1f * (P[row] * tangent.x + P[4 + row] * tangent.y + P[8 + row] * tangent.z);
        cb[r] = 0.1f * (P[row] * bitangent.x + P[4 + row] * bitangent.y + P[8 + row] * bitangent.z);
        cn[r] = 0.1f * (P[row] * normal.x + P[4 + row] * normal.y + P[8 + row] * normal.z);
    }
    const float zt = 0.1f * tangent.z, zb = 0.1f * bitangent.z, zn = 0.1f * normal.z;
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    // A pinhole projection has the last row (0, 0, 1, 0): then w of the projected sample IS its camera z (base[2] = fragPos.z,
    // ct[2] = zt, ... bit for bit: multiplications by 0 and 1), and its three FMAs per tap are dropped. Uniform per frame.
    const bool pinhole_w = P[3] == 0.0f && P[7] == 0.0f && P[11] == 1.0f && P[15] == 0.0f;
    float occlusion = 0.0f;
    auto taps = [&](auto pinhole) {
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
            const float sx = c_ssao_kernel[i * 3], sy = c_ssao_kernel[i * 3 + 1], sz = c_ssao_kernel[i * 3 + 2];
            const float ox = base[0] + ct[0] * sx + cb[0] * sy + cn[0] * sz, oy = base[1] + ct[1] * sx + cb[1] * sy + cn[1] * sz;
            const float sample_z = fragPos.z + zt * sx + zb * sy + zn * sz;
            const float ow = decltype(pinhole)::value ? sample_z : base[2] + ct[2] * sx + cb[2] * sy + cn[2] * sz;
            const float inv = rcp_fast(ow);
            float sampleDepth = rect_linear_z(f.zplane, W, H, ox * inv * hw + hw, oy * inv * hh + hh);
            float rangeCheck = smoothstep01(0.1f * rcp_fast(fabsf(fragPos.z - sampleDepth)));
            occlusion += (sampleDepth <= sample_z - 0.0025f ? 1.0f : 0.0f) * rangeCheck;
        }
    };
    if (pinhole_w) taps(std::true_type{}); else taps(std::false_type{});
    f.ao[p] = 1.0f - (occlusion / 64.0f);
}

__global__ void __launch_bounds__(256) k_ssao_apply_tonemap(const DFrame* __restrict__ frames) {
    const DFrame& f = frames[blockIdx.z];
    const int W = f.W, H = f.H;
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= W || py >= H) return;
    const size_t p = (size_t)py * W + px;
    float4 hdr = f.hdr[p];
    if (f.ssao) {
        // The 16 taps and the centre read the z plane at INTEGER rectangle coordinates, i.e. on texel corners: each is the
        // bilinear mean (weights exactly 1/4) of a 2x2 texel block, and the 17 blocks share a 5x5 neighbourhood. Load those
        // 25 clamped texels once (instead of 68 loads with their own clamps) and form every depth with the summation order of
        // rect_linear_z — bit-identical.
        float z[5][5];
        {
            unsigned col[5], row[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                col[k] = (unsigned)__vimin_s32_relu(px - 3 + k, W - 1);
                row[k] = (unsigned)(__vimin_s32_relu(py - 3 + k, H - 1) * W);
            }
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int i = 0; i < 5; ++i) z[j][i] = __ldg(f.zplane + (row[j] + col[i]));
        }
        auto corner_z = [&](int x, int y) {   // rect_linear_z(px + x, py + y): texels (px+x-1 .. px+x) x (py+y-1 .. py+y)
            return __fmaf_rn(z[y + 3][x + 3], 0.25f, __fmaf_rn(z[y + 3][x + 2], 0.25f, __fmaf_rn(z[y + 2][x + 3], 0.25f, z[y + 2][x + 2] * 0.25f)));
        };
        const float center_d = corner_z(0, 0);
        float result = 0.0f, w_total = 0.0f;
        const float BlurSigma = 3.0f * 0.5f, BlurFalloff = 1.0f / (2.0f * BlurSigma * BlurSigma);
#pragma unroll
        for (int x = -2; x < 2; ++x)
#pragma unroll
            for (int y = -2; y < 2; ++y) {
                int ux = px + x, uy = py + y;
                float c = (ux >= 0 && ux < W && uy >= 0 && uy < H) ? f.ao[(size_t)uy * W + ux] : 0.0f;
                float dd = corner_z(x, y);
                float r = sqrtf((float)(x * x + y * y));
                float ddiff = (dd - center_d) * 300.0f;
                float w = exp2f(-r * r * BlurFalloff - ddiff * ddiff);
                w_total += w;
                result += c * w;
            }
        float a = result / w_total;
        hdr.x *= a; hdr.y *= a; hdr.z *= a;
        f.hdr[p] = hdr;
    }
    if (f.out[SLB_TARGET_RGB]) reinterpret_cast<uchar4*>(f.out[SLB_TARGET_RGB])[p] = tone_map(hdr, f.manual_exposure, f.avg);
}

namespace slbk {

void upload_ssao_tables(const float* noise16x3, const float* kernel64x3) {
    cudaMemcpyToSymbol(c_ssao_noise, noise16x3, sizeof(float) * 48);
    cudaMemcpyToSymbol(c_ssao_kernel, kernel64x3, sizeof(float) * 192);
}
static dim3 frame_grid(int n_frames, int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8, n_frames); }
void launch_background(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s) {
    k_background<<<frame_grid(n_frames, W, H), 256, 0, s>>>(frames);
}
void launch_downsample(const float4* src, int sw, int sh, float4* dst, int n_frames, size_t src_stride, size_t dst_stride, cudaStream_t s) {
    int dw = sw > 1 ? sw >> 1 : 1, dh = sh > 1 ? sh >> 1 : 1;
    dim3 block(16, 16), grid((dw + 15) / 16, (dh + 15) / 16, n_frames);
    k_downsample<<<grid, block, 0, s>>>(src, sw, sh, dst, dw, dh, src_stride, dst_stride);
}
void launch_ssao(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s) {
    k_ssao<<<frame_grid(n_frames, W, H), 256, 0, s>>>(frames);
}
void launch_ssao_apply_tonemap(const DFrame* frames, int n_frames, int W, int H, cudaStream_t s) {
    k_ssao_apply_tonemap<<<frame_grid(n_frames, W, H), 256, 0, s>>>(frames);
}

}  // namespace slbk
