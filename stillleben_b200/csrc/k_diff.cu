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
}  // namespace slbk
