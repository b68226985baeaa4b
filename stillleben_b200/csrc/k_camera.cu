// k_camera.cu — the camera noise model as two fused kernels (reference: python/stillleben/camera_model.py,
// a chain of ~40 PyTorch ops per image: affine_grid / grid_sample, two conv2d, pointwise exposure, torch.poisson /
// normal_, a gather-based RGB<->HSV round trip). SURVEY 8(f-4): the step right after the render path in dataset
// generation.
//   k_cam_stage1: chromatic aberration (bilinear resample with reflection padding, per channel) evaluated into a
//                 (32+4)x(8+4) shared-memory tile -> 5x5 Gaussian blur (zero padded) -> exposure -> Poissonian-
//                 Gaussian noise (Philox counter RNG) -> clamp -> hue shift.        camera_model.py:46-222
//   k_cam_stage2: the post blur (sigma 0.4, 5x5) and the final clamp.               camera_model.py:255-260
// Images are planar float [n][3][H][W] as the reference takes them, or straight from the render target
// (RGBA8 [n][H][W][4], /255 folded in). Every stage can be switched off (stages mask), so the same two
// kernels also serve the reference's individual entry points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

#define CBX 32
#define CBY 8
#define CHALO 2

namespace {

__device__ __forceinline__ float reflect_clip(float c, int size) {   // ATen reflect_coordinates(c, -1, 2*size-1) + clip
    const float mn = -0.5f, span = (float)size;
    c = fabsf(c - mn);
    const float extra = fmodf(c, span);
    const int flips = (int)floorf(c / span);
    const float r = (flips & 1) ? span - extra + mn : extra + mn;
    return fminf(fmaxf(r, 0.0f), (float)(size - 1));
}

struct Src {   // one image in either layout
    const float* f; const uint8_t* u8; int H, W; size_t plane;
    __device__ __forceinline__ float at(int ch, int y, int x) const {
        if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
        return u8 ? (float)u8[((size_t)y * W + x) * 4 + ch] * (1.0f / 255.0f) : f[ch * plane + (size_t)y * W + x];
    }
};

// chromatic aberration of channel ch at output pixel (y, x); outside the image the following convolution pads with 0
__device__ __forceinline__ float ca_sample(const Src& s, const slb_camera_params& p, bool enabled, int ch, int y, int x) {
    if (y < 0 || y >= s.H || x < 0 || x >= s.W) return 0.0f;
    if (!enabled) return s.at(ch, y, x);
    const float xs = (float)(2 * x + 1) / (float)s.W - 1.0f, ys = (float)(2 * y + 1) / (float)s.H - 1.0f;
    const float gx = p.chromatic_scaling[ch] * xs + p.chromatic_translation[ch][0];
    const float gy = p.chromatic_scaling[ch] * ys + p.chromatic_translation[ch][1];
    const float ix = reflect_clip(((gx + 1.0f) * (float)s.W - 1.0f) * 0.5f, s.W);
    const float iy = reflect_clip(((gy + 1.0f) * (float)s.H - 1.0f) * 0.5f, s.H);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float ax = ix - x0, ay = iy - y0;
    const int xi = (int)x0, yi = (int)y0;
    return s.at(ch, yi, xi) * (1.0f - ax) * (1.0f - ay) + s.at(ch, yi, xi + 1) * ax * (1.0f - ay) +
           s.at(ch, yi + 1, xi) * (1.0f - ax) * ay + s.at(ch, yi + 1, xi + 1) * ax * ay;
}

__device__ __forceinline__ void gaussian5(float sigma, float g[5][5]) {   // camera_model.py:76-104
    const float var = sigma * sigma;
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float d = (float)((i - 2) * (i - 2) + (j - 2) * (j - 2));
            g[j][i] = (1.0f / (2.0f * 3.14159265358979f * var)) * expf(-d / (2.0f * var));
            sum += g[j][i];
        }
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) g[j][i] /= sum;
}

// ---- Philox4x32-10 counter RNG: stream = (seed, pixel, channel), no state in memory -----------------------
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
struct Rng {
    uint4 ctr; uint2 key; uint4 buf; int have;
    __device__ Rng(uint64_t seed, uint64_t stream) : have(0) {
        ctr = make_uint4(0u, 0u, (uint32_t)stream, (uint32_t)(stream >> 32));
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        buf = make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ float uniform() {   // (0, 1]
        if (!have) { buf = philox(ctr, key); ++ctr.x; have = 4; }
        const uint32_t v = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
        --have;
        return ((float)(v >> 8) + 1.0f) * (1.0f / 16777216.0f);
    }
};
// Poisson(lambda): multiplication method below 10, Hoermann's transformed rejection (PTRS) above
__device__ float poisson(Rng& rng, float lam) {
    if (!(lam > 0.0f)) return 0.0f;
    if (lam < 10.0f) {
        const float L = expf(-lam);
        float k = 0.0f, prod = rng.uniform();
        while (prod > L) { k += 1.0f; prod *= rng.uniform(); }
        return k;
    }
    const float slam = sqrtf(lam), loglam = logf(lam);
    const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (int it = 0; it < 64; ++it) {
        const float U = rng.uniform() - 0.5f, V = rng.uniform();
        const float us = 0.5f - fabsf(U);
        const float k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return k;
        if (k < 0.0f || (us < 0.013f && V > us)) continue;
        if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -lam + k * loglam - lgammaf(k + 1.0f)) return k;
    }
    return floorf(lam + 0.5f);
}
__device__ __forceinline__ float gaussian(Rng& rng) {
    const float u1 = rng.uniform(), u2 = rng.uniform();
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// camera_model.py:163-222 on one pixel
__device__ __forceinline__ void hue_shift(float& R, float& G, float& B, float shift) {
    const float M = fmaxf(R, fmaxf(G, B)), m = fminf(R, fminf(G, B)), C = M - m;
    int arg = 0;                                        // torch.max returns the FIRST maximal channel
    if (G > R) arg = 1;
    if (B > (arg ? G : R)) arg = 2;
    float h = 0.0f;
    if (C != 0.0f) h = arg == 0 ? (G - B) / C + 0.0f : arg == 1 ? (B - R) / C + 2.0f : (R - G) / C + 4.0f;
    h *= 60.0f;
    if (h < 0.0f) h += 360.0f;
    h += shift * 360.0f;
    if (h < 0.0f) h += 360.0f;
    if (h > 360.0f) h -= 360.0f;
    h /= 60.0f;
    const float X = C * (1.0f - fabsf(fmodf(h, 2.0f) - 1.0f));
    const int oc = min(max((int)h, 0), 5);
    float r, g, b;
    switch (oc) {
        case 0: r = C; g = X; b = 0.f; break;
        case 1: r = X; g = C; b = 0.f; break;
        case 2: r = 0.f; g = C; b = X; break;
        case 3: r = 0.f; g = X; b = C; break;
        case 4: r = X; g = 0.f; b = C; break;
        default: r = C; g = 0.f; b = X; break;
    }
    R = r + m; G = g + m; B = b + m;
}

__global__ void __launch_bounds__(CBX * CBY) k_cam_stage1(const float* __restrict__ in_f, const uint8_t* __restrict__ in_u8,
                                                          float* __restrict__ out, const slb_camera_params* __restrict__ params,
                                                          int H, int W) {
    __shared__ float s_t[3][CBY + 2 * CHALO][CBX + 2 * CHALO];
    const int img = blockIdx.z;
    const slb_camera_params p = params[img];
    const size_t plane = (size_t)H * W;
    Src src;
    src.f = in_f ? in_f + (size_t)img * 3 * plane : nullptr;
    src.u8 = in_u8 ? in_u8 + (size_t)img * 4 * plane : nullptr;
    src.H = H; src.W = W; src.plane = plane;
    const int bx = blockIdx.x * CBX, by = blockIdx.y * CBY;
    const int tid = threadIdx.y * CBX + threadIdx.x;
    const bool do_ca = p.stages & SLB_CAM_CHROMATIC, do_blur = (p.stages & SLB_CAM_BLUR) && p.blur_sigma > 0.0f;
    if (do_blur) {
        for (int t = tid; t < 3 * (CBY + 2 * CHALO) * (CBX + 2 * CHALO); t += CBX * CBY) {
            const int ch = t / ((CBY + 2 * CHALO) * (CBX + 2 * CHALO)), rem = t % ((CBY + 2 * CHALO) * (CBX + 2 * CHALO));
            const int ly = rem / (CBX + 2 * CHALO), lx = rem % (CBX + 2 * CHALO);
            s_t[ch][ly][lx] = ca_sample(src, p, do_ca, ch, by + ly - CHALO, bx + lx - CHALO);
        }
        __syncthreads();
    }
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= W || y >= H) return;
    float v[3];
    if (do_blur) {
        float g[5][5];
        gaussian5(p.blur_sigma, g);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int i = 0; i < 5; ++i) acc += g[j][i] * s_t[ch][threadIdx.y + j][threadIdx.x + i];
            v[ch] = acc;
        }
    } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = ca_sample(src, p, do_ca, ch, y, x);
    }
    if (p.stages & SLB_CAM_EXPOSURE) {
        const float e = expf(p.exposure_deltaS);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = 1.0f / (1.0f + e * (1.0f / (v[ch] + 0.0001f) - 1.0f));
    }
    if ((p.stages & SLB_CAM_NOISE) && p.do_noise) {
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
            Rng rng(p.seed, ((uint64_t)img * 3 + ch) * plane + (size_t)y * W + x);
            float pv = v[ch];
            if (p.noise_a > 0.0f) { const float chi = 1.0f / p.noise_a; pv = poisson(rng, chi * v[ch]) / chi; }
            const float gv = p.noise_b > 0.0f ? gaussian(rng) * p.noise_b : 0.0f;
            v[ch] = fminf(fmaxf(pv + gv, 0.0f), 1.0f);
        }
    }
    if (p.stages & SLB_CAM_CLAMP) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = fminf(fmaxf(v[ch], 0.0f), 1.0f);
    }
    if (p.stages & SLB_CAM_HUE) hue_shift(v[0], v[1], v[2], p.hue_shift);
    float* o = out + (size_t)img * 3 * plane + (size_t)y * W + x;
    o[0] = v[0]; o[plane] = v[1]; o[2 * plane] = v[2];
}

__global__ void __launch_bounds__(CBX * CBY) k_cam_stage2(const float* __restrict__ in, float* __restrict__ out, float sigma, int H, int W) {
    __shared__ float s_t[3][CBY + 2 * CHALO][CBX + 2 * CHALO];
    const size_t plane = (size_t)H * W;
    const float* src = in + (size_t)blockIdx.z * 3 * plane;
    const int bx = blockIdx.x * CBX, by = blockIdx.y * CBY;
    const int tid = threadIdx.y * CBX + threadIdx.x;
    for (int t = tid; t < 3 * (CBY + 2 * CHALO) * (CBX + 2 * CHALO); t += CBX * CBY) {
        const int ch = t / ((CBY + 2 * CHALO) * (CBX + 2 * CHALO)), rem = t % ((CBY + 2 * CHALO) * (CBX + 2 * CHALO));
        const int ly = rem / (CBX + 2 * CHALO), lx = rem % (CBX + 2 * CHALO);
        const int yy = by + ly - CHALO, xx = bx + lx - CHALO;
        s_t[ch][ly][lx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? src[ch * plane + (size_t)yy * W + xx] : 0.0f;
    }
    __syncthreads();
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= W || y >= H) return;
    float g[5][5];
    gaussian5(sigma, g);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int i = 0; i < 5; ++i) acc += g[j][i] * s_t[ch][threadIdx.y + j][threadIdx.x + i];
        out[(size_t)blockIdx.z * 3 * plane + ch * plane + (size_t)y * W + x] = fminf(fmaxf(acc, 0.0f), 1.0f);
    }
}

}  // namespace

namespace slbk {
void launch_camera_stage1(const float* in_f, const uint8_t* in_u8, float* out, const slb_camera_params* params, int n, int H, int W,
                          cudaStream_t s) {
    dim3 grid((W + CBX - 1) / CBX, (H + CBY - 1) / CBY, n), block(CBX, CBY);
    k_cam_stage1<<<grid, block, 0, s>>>(in_f, in_u8, out, params, H, W);
}
void launch_camera_stage2(const float* in, float* out, float sigma, int n, int H, int W, cudaStream_t s) {
    dim3 grid((W + CBX - 1) / CBX, (H + CBY - 1) / CBY, n), block(CBX, CBY);
    k_cam_stage2<<<grid, block, 0, s>>>(in, out, sigma, H, W);
}
}  // namespace slbk
