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

extern "C" {

void orc_diff_sobel_valid_mask(const int16_t* inst, const float* depth, uint8_t* valid, int H, int W) {
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
