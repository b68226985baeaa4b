// orc_diff.cpp — ORACLE (test infrastructure only; see orc_core.h for the rules).
//
// CPU restatement of the two render-and-compare mask kernels of the reference
// (python/src/diff.cu:13-99 generateSobelValidMaskKernel, :101-193 dilateObjectMaskKernel).
// The reference stages a (32+2)x(32+2) window in shared memory with clamp-to-edge halo loads;
// semantically every pixel looks at its clamped 3x3 neighbourhood. Loop order (x outer, y inner)
// and the inner-loop-only `break` of the dilate kernel are kept because they decide WHICH
// neighbour's coordinate is copied.
#include <cstdint>
#include <cstddef>

static inline int clampi(int v, int n) { v = v > 0 ? v : 0; return v < n - 1 ? v : n - 1; }

// The reference ships TWO implementations of both mask functions and they differ: the CUDA kernels
// (diff.cu, the branch the product follows: the module is CUDA-only, diff.py:5-9) treat every pixel, with a
// clamped 3x3 window walked column-offset-outer; the CPU branch of the bridge (bridge_diff.cpp:36-63,105-152)
// skips the one-pixel image border (valid stays 1, dilated mask / coordinates stay 0 there), walks the window
// row-offset-outer, and writes the neighbour's coordinate only when the pixel really is dilated.
// variant 0 = CUDA kernels (default), 1 = CPU branch. The CPU branch exists so that the restatement can be pinned
// against oracle/_ref/diff (the reference's own extension) in a container without a GPU.
static int g_variant = 0;

extern "C" {

void orc_diff_set_variant(int v) { g_variant = v; }

static void sobel_valid_mask_cpu_branch(const int16_t* inst, const float* depth, uint8_t* valid, int H, int W) {
    for (size_t p = 0; p < (size_t)H * W; ++p) valid[p] = 1;
    for (int h = 1; h < H - 1; ++h)
        for (int w = 1; w < W - 1; ++w) {
            int16_t cur = inst[(size_t)h * W + w];
            if (cur == 0) continue;
            float d = depth[(size_t)h * W + w];
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    size_t q = (size_t)(h + x) * W + (w + y);
                    if (inst[q] != cur && inst[q] != 0 && depth[q] < d) valid[(size_t)h * W + w] = 0;
                }
        }
}
static void dilate_object_mask_cpu_branch(const uint8_t* mask, const uint8_t* valid, const float* coords, int cs, uint8_t* mask_out,
                                          float* coords_out, int H, int W) {
    for (size_t p = 0; p < (size_t)H * W; ++p) { mask_out[p] = 0; coords_out[p * 3] = coords_out[p * 3 + 1] = coords_out[p * 3 + 2] = 0.0f; }
    for (int h = 1; h < H - 1; ++h)
        for (int w = 1; w < W - 1; ++w) {
            size_t p = (size_t)h * W + w;
            mask_out[p] = mask[p];
            for (int k = 0; k < 3; ++k) coords_out[p * 3 + k] = coords[p * cs + k];
            if (mask[p] != 0) continue;
            bool allValid = true, allBackground = true;
            float c3[3] = {0, 0, 0};
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    size_t q = (size_t)(h + x) * W + (w + y);
                    if (mask[q] != 0) { allBackground = false; for (int k = 0; k < 3; ++k) c3[k] = coords[q * cs + k]; }
                    if (valid[q] == 0) { allValid = false; break; }
                }
            if (allBackground || !allValid) continue;
            mask_out[p] = 1;
            for (int k = 0; k < 3; ++k) coords_out[p * 3 + k] = c3[k];
        }
}

void orc_diff_sobel_valid_mask(const int16_t* inst, const float* depth, uint8_t* valid, int H, int W) {
    if (g_variant == 1) { sobel_valid_mask_cpu_branch(inst, depth, valid, H, W); return; }
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            uint8_t ok = 1;
            int16_t cur = inst[(size_t)r * W + c];
            if (cur != 0) {
                float d = depth[(size_t)r * W + c];
                for (int x = -1; x <= 1; ++x)
                    for (int y = -1; y <= 1; ++y) {
                        size_t q = (size_t)clampi(r + y, H) * W + clampi(c + x, W);
                        if (inst[q] != cur && inst[q] != 0 && depth[q] < d) ok = 0;
                    }
            }
            valid[(size_t)r * W + c] = ok;
        }
}

void orc_diff_dilate_object_mask(const uint8_t* mask, const uint8_t* valid, const float* coords, int coord_stride,
                                 uint8_t* mask_out, float* coords_out /* HxWx3 dense */, int H, int W) {
    if (g_variant == 1) { dilate_object_mask_cpu_branch(mask, valid, coords, coord_stride, mask_out, coords_out, H, W); return; }
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            size_t p = (size_t)r * W + c;
            uint8_t om = mask[p];
            float oc[3] = {coords[p * coord_stride], coords[p * coord_stride + 1], coords[p * coord_stride + 2]};
            if (om == 0) {
                bool allValid = true, allBackground = true;
                for (int x = -1; x <= 1; ++x)
                    for (int y = -1; y <= 1; ++y) {
                        size_t q = (size_t)clampi(r + y, H) * W + clampi(c + x, W);
                        if (mask[q] != 0) {
                            allBackground = false;
                            oc[0] = coords[q * coord_stride]; oc[1] = coords[q * coord_stride + 1]; oc[2] = coords[q * coord_stride + 2];
                        }
                        if (valid[q] == 0) { allValid = false; break; }
                    }
                if (!allBackground && allValid) om = 1;
            }
            mask_out[p] = om;
            coords_out[p * 3] = oc[0]; coords_out[p * 3 + 1] = oc[1]; coords_out[p * 3 + 2] = oc[2];
        }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Render-and-compare backward (reference: python/stillleben/diff.py:73-127 compute_image_space_gradients,
// :355-523 backpropagate_gradient_to_poses). Literal float32 restatement of the tensor program, object by
// object, pixel by pixel, with the products taken in the order of its bmm chain:
//   grad_x[c] = -(rgb[c][x+1] - rgb[c][x-1]) / (2/W*2), grad_y likewise with H (zero padded), zeroed where
//     the sobel-valid mask is false                                                           (:101-124)
//   mask, coords = dilate_object_mask(instance == idx, valid, coordinates)                   (:401-411)
//   y = T0 (coords, 1);  den = P[2,:] . y  (the reference really uses row 2)     (:431-443)
//   g_coord[j][i] = P[j,i] / den - P[2,i] / den^2 * (P[j,:] . y)
//   g_pose[i][k] = (T0 G_k x)[i] for the six generators alpha, beta, gamma, a, b, c           (:446-483)
//   out[obj][k] = sum_pixels sum_c grad_in[c] * ((g_xy[c][:] g_coord) g_pose)[k]               (:485-521)
// Matrices are row-major here (P[r*4+c], T0[r*4+c]) as torch holds them.
extern "C" void orc_diff_pose_grad(const uint8_t* rgb /* HxWx4 */, const int16_t* inst, const float* coord /* HxWx4, w = depth */,
                                   const float* grad_img /* 3xHxW */, const float* P, const float* poses /* n_obj x 16 */,
                                   const int32_t* instance_ids, int n_obj, float* out /* n_obj x 6 */, int H, int W) {
    const size_t N = (size_t)H * W;
    uint8_t* valid = new uint8_t[N];
    float* depth = new float[N];
    for (size_t p = 0; p < N; ++p) depth[p] = coord[p * 4 + 3];
    orc_diff_sobel_valid_mask(inst, depth, valid, H, W);
    float* gx = new float[3 * N];
    float* gy = new float[3 * N];
    const float sx = 2.0f / (float)W * 2.0f, sy = 2.0f / (float)H * 2.0f;   // "sobel /= 2 / _w * 2"
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < H; ++r)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)r * W + x;
                auto px = [&](int rr, int xx) -> float {
                    if (rr < 0 || rr >= H || xx < 0 || xx >= W) return 0.0f;
                    return (float)rgb[((size_t)rr * W + xx) * 4 + c] / 255.0f;
                };
                float dx = -1.0f / sx * px(r, x - 1) + 1.0f / sx * px(r, x + 1);
                float dy = -1.0f / sy * px(r - 1, x) + 1.0f / sy * px(r + 1, x);
                gx[c * N + p] = valid[p] ? -dx : 0.0f;
                gy[c * N + p] = valid[p] ? -dy : 0.0f;
            }
    uint8_t* mask = new uint8_t[N];
    uint8_t* mask_d = new uint8_t[N];
    float* oc = new float[3 * N];
    for (int o = 0; o < n_obj; ++o) {
        const float* T = poses + 16 * o;
        for (size_t p = 0; p < N; ++p) mask[p] = (int32_t)inst[p] == instance_ids[o];
        orc_diff_dilate_object_mask(mask, valid, coord, 4, mask_d, oc, H, W);
        double acc[6] = {0, 0, 0, 0, 0, 0};   // torch sums N x 6 floats pairwise; double keeps the oracle order-free
        for (size_t p = 0; p < N; ++p) {
            if (!mask_d[p]) continue;
            const float x[4] = {oc[p * 3], oc[p * 3 + 1], oc[p * 3 + 2], 1.0f};
            float y[4];
            for (int r = 0; r < 4; ++r) y[r] = T[r * 4] * x[0] + T[r * 4 + 1] * x[1] + T[r * 4 + 2] * x[2] + T[r * 4 + 3] * x[3];
            float Py[3];
            for (int r = 0; r < 3; ++r) Py[r] = P[r * 4] * y[0] + P[r * 4 + 1] * y[1] + P[r * 4 + 2] * y[2] + P[r * 4 + 3] * y[3];
            float gcoord[2][3];
            for (int j = 0; j < 2; ++j)
                for (int i = 0; i < 3; ++i)
                    gcoord[j][i] = P[j * 4 + i] * (1.0f / Py[2]) + (P[2 * 4 + i] * (-1.0f / (Py[2] * Py[2]))) * Py[j];
            // generators applied to x, then T0; homogeneous row dropped
            const float gen[6][4] = {{0.0f, -x[2], x[1], 0.0f}, {x[2], 0.0f, -x[0], 0.0f}, {-x[1], x[0], 0.0f, 0.0f},
                                     {x[3], 0.0f, 0.0f, 0.0f},  {0.0f, x[3], 0.0f, 0.0f},  {0.0f, 0.0f, x[3], 0.0f}};
            float gpose[3][6];
            for (int k = 0; k < 6; ++k)
                for (int i = 0; i < 3; ++i)
                    gpose[i][k] = T[i * 4] * gen[k][0] + T[i * 4 + 1] * gen[k][1] + T[i * 4 + 2] * gen[k][2] + T[i * 4 + 3] * gen[k][3];
            float A[3][3];   // g_xy [3x2] @ g_coord [2x3]
            for (int c = 0; c < 3; ++c)
                for (int i = 0; i < 3; ++i) A[c][i] = gx[c * N + p] * gcoord[0][i] + gy[c * N + p] * gcoord[1][i];
            for (int k = 0; k < 6; ++k) {
                float g = 0.0f;
                for (int c = 0; c < 3; ++c) {
                    float Bck = A[c][0] * gpose[0][k] + A[c][1] * gpose[1][k] + A[c][2] * gpose[2][k];
                    g += grad_img[c * N + p] * Bck;
                }
                acc[k] += (double)g;
            }
        }
        for (int k = 0; k < 6; ++k) out[o * 6 + k] = (float)acc[k];
    }
    delete[] valid; delete[] depth; delete[] gx; delete[] gy; delete[] mask; delete[] mask_d; delete[] oc;
}
