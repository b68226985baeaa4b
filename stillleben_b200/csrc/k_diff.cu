// k_diff.cu — the two mask kernels of the render-and-compare backward pass
// (reference: python/src/diff.cu:13-99 generateSobelValidMaskKernel, :101-193 dilateObjectMaskKernel;
// bridge: python/src/bridge_diff.cpp:13-157). One thread per pixel over a clamped 3x3 window that is
// staged in shared memory (34x10 halo tile per 32x8 block); no device synchronisation afterwards.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

#define DBX 32
#define DBY 8

__device__ __forceinline__ int clampi(int v, int n) { v = v > 0 ? v : 0; return v < n - 1 ? v : n - 1; }

__global__ void __launch_bounds__(DBX * DBY) k_sobel_valid(const int16_t* __restrict__ inst, const float* __restrict__ depth,
                                                           uint8_t* __restrict__ valid, int H, int W) {
    __shared__ int16_t s_i[DBY + 2][DBX + 2];
    __shared__ float s_d[DBY + 2][DBX + 2];
    const int bx = blockIdx.x * DBX, by = blockIdx.y * DBY;
    for (int t = threadIdx.y * DBX + threadIdx.x; t < (DBX + 2) * (DBY + 2); t += DBX * DBY) {
        int ly = t / (DBX + 2), lx = t % (DBX + 2);
        size_t q = (size_t)clampi(by + ly - 1, H) * W + clampi(bx + lx - 1, W);
        s_i[ly][lx] = inst[q]; s_d[ly][lx] = depth[q];
    }
    __syncthreads();
    const int c = bx + threadIdx.x, r = by + threadIdx.y;
    if (c >= W || r >= H) return;
    const int16_t cur = s_i[threadIdx.y + 1][threadIdx.x + 1];
    uint8_t ok = 1;
    if (cur != 0) {
        const float d = s_d[threadIdx.y + 1][threadIdx.x + 1];
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y) {
                int16_t o = s_i[threadIdx.y + y][threadIdx.x + x];
                if (o != cur && o != 0 && s_d[threadIdx.y + y][threadIdx.x + x] < d) ok = 0;
            }
    }
    valid[(size_t)r * W + c] = ok;
}

__global__ void __launch_bounds__(DBX * DBY) k_dilate(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ valid,
                                                      const float* __restrict__ coords, int cs, uint8_t* __restrict__ mask_out,
                                                      float* __restrict__ coords_out, int H, int W) {
    __shared__ uint8_t s_m[DBY + 2][DBX + 2], s_v[DBY + 2][DBX + 2];
    const int bx = blockIdx.x * DBX, by = blockIdx.y * DBY;
    for (int t = threadIdx.y * DBX + threadIdx.x; t < (DBX + 2) * (DBY + 2); t += DBX * DBY) {
        int ly = t / (DBX + 2), lx = t % (DBX + 2);
        size_t q = (size_t)clampi(by + ly - 1, H) * W + clampi(bx + lx - 1, W);
        s_m[ly][lx] = mask[q]; s_v[ly][lx] = valid[q];
    }
    __syncthreads();
    const int c = bx + threadIdx.x, r = by + threadIdx.y;
    if (c >= W || r >= H) return;
    const size_t p = (size_t)r * W + c;
    uint8_t om = s_m[threadIdx.y + 1][threadIdx.x + 1];
    size_t src = p;
    if (om == 0) {
        bool allValid = true, allBackground = true;
        for (int x = 0; x < 3; ++x)
            for (int y = 0; y < 3; ++y) {   // the reference's `break` leaves only the inner loop
                if (s_m[threadIdx.y + y][threadIdx.x + x] != 0) {
                    allBackground = false;
                    src = (size_t)clampi(r + y - 1, H) * W + clampi(c + x - 1, W);
                }
                if (s_v[threadIdx.y + y][threadIdx.x + x] == 0) { allValid = false; break; }
            }
        if (!allBackground && allValid) om = 1;
    }
    mask_out[p] = om;
    coords_out[p * 3] = coords[src * cs]; coords_out[p * 3 + 1] = coords[src * cs + 1]; coords_out[p * 3 + 2] = coords[src * cs + 2];
}

// ---------------------------------------------------------------------------------------------
// Fused render-and-compare backward (reference: python/stillleben/diff.py:73-127 compute_image_space_gradients +
// :355-523 backpropagate_gradient_to_poses — there a Python loop over the objects with ~30 torch ops, two kernel
// launches and a device synchronisation per object). Here: ONE pass over the pixels for ALL objects.
//   * the sobel-valid mask of the pixel's 3x3 neighbourhood is recomputed from a (32+4)x(8+4) instance / depth
//     tile in shared memory (no intermediate mask image),
//   * image gradients are the reference's central differences of rgb/255 (zero padded), zero where invalid,
//   * a pixel contributes to object o if it shows o, or (dilation, diff.cu:101-193) if all nine neighbours are
//     valid and one of them shows o — then with THAT neighbour's object coordinates (last one in x-outer /
//     y-inner order), its own image gradient and its own dLoss/dImage,
//   * per (pixel, object): y = T0 x, d(xy)/dX from the projection rows (row 2 as the divisor, exactly as the
//     reference), dX/d(alpha, beta, gamma, a, b, c) = T0 G_k x, contracted to six numbers,
//   * block-level shuffle reduction per object present in the block -> partial[block][object][6]; a second
//     kernel sums the partials in a fixed order in double (deterministic, unlike float atomics).
// Matrices arrive row-major (P[r*4+c], T0[r*4+c]).
#define PGX 32
#define PGY 8
__global__ void __launch_bounds__(PGX * PGY) k_pose_grad(const uint8_t* __restrict__ rgb, const int16_t* __restrict__ inst,
                                                         const float* __restrict__ coord, const float* __restrict__ grad_img,
                                                         const float* __restrict__ params /* P[16], then per object T0[16] + id */,
                                                         int n_obj, float* __restrict__ partial, int H, int W) {
    __shared__ int16_t s_i[PGY + 4][PGX + 4];
    __shared__ float s_d[PGY + 4][PGX + 4];
    __shared__ uint8_t s_v[PGY + 2][PGX + 2];
    __shared__ float s_red[PGX * PGY / 32][6];
    const int bx = blockIdx.x * PGX, by = blockIdx.y * PGY;
    const int tid = threadIdx.y * PGX + threadIdx.x;
    for (int t = tid; t < (PGX + 4) * (PGY + 4); t += PGX * PGY) {
        const int ly = t / (PGX + 4), lx = t % (PGX + 4);
        // clamp-to-edge twice: neighbour q of p is clamp(p + d), and q's own neighbours are clamp(q + d)
        const size_t q = (size_t)clampi(by + ly - 2, H) * W + clampi(bx + lx - 2, W);
        s_i[ly][lx] = inst[q]; s_d[ly][lx] = coord[q * 4 + 3];
    }
    __syncthreads();
    // valid mask of the (32+2)x(8+2) pixels around the block, each from ITS clamped 3x3 neighbourhood
    for (int t = tid; t < (PGX + 2) * (PGY + 2); t += PGX * PGY) {
        const int ly = t / (PGX + 2), lx = t % (PGX + 2);
        const int gr = clampi(by + ly - 1, H), gc = clampi(bx + lx - 1, W);   // the pixel this entry stands for
        const int cy = gr - by + 2, cx = gc - bx + 2;                         // its place in the +-2 tile
        const int16_t cur = s_i[cy][cx];
        uint8_t ok = 1;
        if (cur != 0) {
            const float d = s_d[cy][cx];
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    const int ny = clampi(gr + y, H) - by + 2, nx = clampi(gc + x, W) - bx + 2;
                    const int16_t o = s_i[ny][nx];
                    if (o != cur && o != 0 && s_d[ny][nx] < d) ok = 0;
                }
        }
        s_v[ly][lx] = ok;
    }
    __syncthreads();
    const int c = bx + threadIdx.x, r = by + threadIdx.y;
    const bool in_image = c < W && r < H;
    const size_t N = (size_t)H * W, p = (size_t)min(r, H - 1) * W + min(c, W - 1);
    // neighbourhood in the reference's visiting order (x outer, y inner), clamped
    int16_t ids[9]; bool all_valid = true;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y) {
            ids[x * 3 + y] = s_i[clampi(r + y - 1, H) - by + 2][clampi(c + x - 1, W) - bx + 2];
            all_valid &= s_v[clampi(r + y - 1, H) - by + 1][clampi(c + x - 1, W) - bx + 1] != 0;
        }
    const int16_t own = ids[4];
    const bool valid = s_v[min(r, H - 1) - by + 1][min(c, W - 1) - bx + 1] != 0;
    // s[j] = sum_c dLoss/dI_c * dI_c/d(x_j): the image-side factor, shared by every object the pixel feeds
    float s0 = 0.0f, s1 = 0.0f;
    if (in_image && valid) {
        const float kx = (float)W * 0.25f, ky = (float)H * 0.25f, k255 = 1.0f / 255.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float l = c > 0 ? (float)rgb[(p - 1) * 4 + ch] * k255 : 0.0f, rr = c + 1 < W ? (float)rgb[(p + 1) * 4 + ch] * k255 : 0.0f;
            const float u = r > 0 ? (float)rgb[(p - W) * 4 + ch] * k255 : 0.0f, dn = r + 1 < H ? (float)rgb[(p + W) * 4 + ch] * k255 : 0.0f;
            const float g = grad_img[ch * N + p];
            s0 += g * -((rr - l) * kx);
            s1 += g * -((dn - u) * ky);
        }
    }
    const float* P = params;
    const int lane = tid & 31, warp = tid >> 5;
    for (int o = 0; o < n_obj; ++o) {
        const float* T = params + 16 + 17 * o;
        const int16_t id = (int16_t)__float_as_int(__ldg(T + 16));
        int hit = -1;
        if (in_image) {
            if (own == id) hit = 4;
            else if (all_valid) {
#pragma unroll
                for (int n = 0; n < 9; ++n) if (ids[n] == id) hit = n;
            }
        }
        if (!__syncthreads_or(hit >= 0)) continue;
        float g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (hit >= 0 && (s0 != 0.0f || s1 != 0.0f)) {
            const int sx = hit / 3, sy = hit - 3 * sx;
            const size_t q = (size_t)clampi(r + sy - 1, H) * W + clampi(c + sx - 1, W);
            const float4 oc = *reinterpret_cast<const float4*>(coord + q * 4);
            const float x0 = oc.x, x1 = oc.y, x2 = oc.z;
            float y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = __ldg(T + i * 4) * x0 + __ldg(T + i * 4 + 1) * x1 + __ldg(T + i * 4 + 2) * x2 + __ldg(T + i * 4 + 3);
            float Py[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) Py[i] = __ldg(P + i * 4) * y[0] + __ldg(P + i * 4 + 1) * y[1] + __ldg(P + i * 4 + 2) * y[2] + __ldg(P + i * 4 + 3) * y[3];
            const float inv = 1.0f / Py[2], inv2 = -inv * inv;
            float t[3];   // t_i = sum_j s_j * d(xy_j)/dX_i
#pragma unroll
            for (int i = 0; i < 3; ++i)
                t[i] = s0 * (__ldg(P + i) * inv + __ldg(P + 8 + i) * inv2 * Py[0]) + s1 * (__ldg(P + 4 + i) * inv + __ldg(P + 8 + i) * inv2 * Py[1]);
            // u = T0[:3,:3]^T t; rotations: (G x) as cross products, translations: the columns of T0
            float u[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) u[k] = t[0] * __ldg(T + k) + t[1] * __ldg(T + 4 + k) + t[2] * __ldg(T + 8 + k);
            g6[0] = -x2 * u[1] + x1 * u[2];
            g6[1] = x2 * u[0] - x0 * u[2];
            g6[2] = -x1 * u[0] + x0 * u[1];
            g6[3] = u[0]; g6[4] = u[1]; g6[5] = u[2];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) g6[k] += __shfl_xor_sync(0xffffffffu, g6[k], off);
            if (lane == 0) s_red[warp][k] = g6[k];
        }
        __syncthreads();
        if (tid < 6) {
            float sum = 0.0f;
            for (int w = 0; w < PGX * PGY / 32; ++w) sum += s_red[w][tid];
            partial[((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * n_obj + o) * 6 + tid] = sum;
        }
        __syncthreads();
    }
}
// out[o][k] = sum over blocks of partial[block][o][k], fixed order, double accumulation
__global__ void __launch_bounds__(256) k_pose_grad_reduce(const float* __restrict__ partial, int n_blocks, int n_obj, float* __restrict__ out) {
    __shared__ double s_acc[256];
    const int ok = blockIdx.x;   // object * 6 + k
    double acc = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) acc += (double)partial[(size_t)b * n_obj * 6 + ok];
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) s_acc[threadIdx.x] += s_acc[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[ok] = (float)s_acc[0];
}

namespace slbk {
void launch_sobel_valid_mask(const int16_t* inst, const float* depth, uint8_t* valid, int H, int W, cudaStream_t s) {
    dim3 grid((W + DBX - 1) / DBX, (H + DBY - 1) / DBY), block(DBX, DBY);
    k_sobel_valid<<<grid, block, 0, s>>>(inst, depth, valid, H, W);
}
void launch_dilate_object_mask(const uint8_t* mask, const uint8_t* valid, const float* coords, int coord_stride, uint8_t* mask_out,
                               float* coords_out, int H, int W, cudaStream_t s) {
    dim3 grid((W + DBX - 1) / DBX, (H + DBY - 1) / DBY), block(DBX, DBY);
    k_dilate<<<grid, block, 0, s>>>(mask, valid, coords, coord_stride, mask_out, coords_out, H, W);
}
size_t pose_grad_partial_floats(int n_obj, int H, int W) {
    return (size_t)((W + PGX - 1) / PGX) * ((H + PGY - 1) / PGY) * n_obj * 6;
}
void launch_pose_grad(const uint8_t* rgb, const int16_t* inst, const float* coord, const float* grad_img, const float* params, int n_obj,
                      float* partial, float* out, int H, int W, cudaStream_t s) {
    dim3 grid((W + PGX - 1) / PGX, (H + PGY - 1) / PGY), block(PGX, PGY);
    cudaMemsetAsync(partial, 0, pose_grad_partial_floats(n_obj, H, W) * sizeof(float), s);
    k_pose_grad<<<grid, block, 0, s>>>(rgb, inst, coord, grad_img, params, n_obj, partial, H, W);
    k_pose_grad_reduce<<<n_obj * 6, 256, 0, s>>>(partial, (int)(grid.x * grid.y), n_obj, out);
}
}  // namespace slbk
