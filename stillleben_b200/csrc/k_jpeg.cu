// Batched baseline JPEG encoder on the device: n images of one shape -> n complete JFIF files in device memory.
//
// Replaces the JPEG leg of the reference's image saver (src/image_saver.cpp:55-97: AnyImageConverter -> Magnum's
// JpegImageConverter = libjpeg, jpeg_set_defaults + jpeg_set_quality(80, TRUE)). libjpeg is deterministic integer code, so
// the files are reproduced BYTE FOR BYTE (oracle/jpeg_np.py restates the algorithm and is itself checked against libjpeg's
// output through PIL; tests/test_gpu_jpeg.py holds these kernels to it):
//   k_jpeg_blocks   one thread per 8x8 block in scan (MCU-interleaved) order: RGB -> YCbCr in 16-bit fixed point
//                   (jccolor.c), 2x2 chroma box filter with the alternating 1,2 bias (jcsample.c:h2v2_downsample), edge
//                   replication (jcprepct.c), level shift, slow-integer forward DCT (jfdctint.c), quantisation
//                   (jcdctmgr.c), coefficients stored in zigzag order (int16)
//   k_jpeg_scan     one block per image: entropy-coded length of every block (jchuff.c:encode_one_block with the Annex K
//                   tables; DC prediction per component; dummy edge blocks of jccoefct.c) and their exclusive prefix sum
//   k_jpeg_emit     one thread per block: the same walk again, this time ORing the code words into the image's bit stream
//                   at the block's bit offset (MSB first; only the first and last word of a block are shared)
//   k_jpeg_finish   one block per image: pad the last byte with 1-bits, stuff 0xFF -> 0xFF 0x00 (chunked prefix sum of
//                   the expansion), prepend the header (jcmarker.c order, built on the host), append EOI, write the size
// HBM traffic: the image is read once (plus the 2x2 chroma footprint), 2 B per coefficient written and read twice, the
// stream written and read once, the file written once.
#include <cstdint>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace slbk {

namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t kLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                            18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables: code counts per length (1..16) and symbols in code order
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
    0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
    0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
    0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
    0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

void quant_table(const uint8_t* base, int quality, uint8_t out[64]) {   // jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int i = 0; i < 64; ++i) {
        int v = (base[i] * s + 50) / 100;
        out[i] = (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
    }
}
void derive(const uint8_t bits[16], const uint8_t* vals, uint32_t lut[256]) {   // jchuff.c:jpeg_make_c_derived_tbl -> code << 8 | length
    memset(lut, 0, 256 * sizeof(uint32_t));
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) lut[vals[k++]] = (code++ << 8) | (uint32_t)len;
        code <<= 1;
    }
}
void put_marker(std::vector<uint8_t>& o, uint8_t tag, const std::vector<uint8_t>& payload) {
    o.push_back(0xFF); o.push_back(tag);
    const size_t n = payload.size() + 2;
    o.push_back((uint8_t)(n >> 8)); o.push_back((uint8_t)n);
    o.insert(o.end(), payload.begin(), payload.end());
}

}  // namespace

// geometry of one image in blocks
struct JpegGeom {
    int W, H, C;          // C: bytes per pixel of the input (1, 3, 4)
    int ncomp;            // 1 (grey) or 3
    int mw, mh;           // MCUs per row / column (8x8 for grey, 16x16 for colour)
    int wb, hb;           // luma blocks that contain image samples (ceil(W/8), ceil(H/8)): the rest of the MCU grid is dummy
    int n_blocks;         // blocks per image in scan order
};

// (the quantisation divisors depend on the quality and travel as a kernel PARAMETER, JpegQuant: contexts / streams encoding at
// different qualities never share state; the Huffman tables below are the same for every call)
__constant__ uint32_t c_huff[4][256];    // DC luma, AC luma, DC chroma, AC chroma: code << 8 | length

#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
// one 1-D pass of jfdctint.c on eight values with stride `st`
template <bool FIRST>
__device__ __forceinline__ void fdct_pass(int* d, int st) {
    const int t0 = d[0] + d[7 * st], t7 = d[0] - d[7 * st], t1 = d[st] + d[6 * st], t6 = d[st] - d[6 * st];
    const int t2 = d[2 * st] + d[5 * st], t5 = d[2 * st] - d[5 * st], t3 = d[3 * st] + d[4 * st], t4 = d[3 * st] - d[4 * st];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int N = FIRST ? 13 - 2 : 13 + 2;
    if (FIRST) { d[0] = (t10 + t11) << 2; d[4 * st] = (t10 - t11) << 2; }
    else { d[0] = DESCALE(t10 + t11, 2); d[4 * st] = DESCALE(t10 - t11, 2); }
    int z1 = (t12 + t13) * 4433;
    d[2 * st] = DESCALE(z1 + t13 * 6270, N);
    d[6 * st] = DESCALE(z1 + t12 * (-15137), N);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
    d[7 * st] = DESCALE(a4 + z1 + z3, N);
    d[5 * st] = DESCALE(a5 + z2 + z4, N);
    d[3 * st] = DESCALE(a6 + z2 + z3, N);
    d[st] = DESCALE(a7 + z1 + z4, N);
}

// scan-order block index -> (component, block x, block y); colour MCU = Y00 Y01 Y10 Y11 Cb Cr
__device__ __forceinline__ void block_pos(const JpegGeom& g, int b, int& comp, int& bx, int& by) {
    if (g.ncomp == 1) { comp = 0; bx = b % g.mw; by = b / g.mw; return; }
    const int m = b / 6, k = b - 6 * m, mx = m % g.mw, my = m / g.mw;
    if (k < 4) { comp = 0; bx = 2 * mx + (k & 1); by = 2 * my + (k >> 1); }
    else { comp = k - 3; bx = mx; by = my; }
}

__global__ void __launch_bounds__(128) k_jpeg_blocks(const uint8_t* __restrict__ images, JpegGeom g, const JpegQuant q, int n_images,
                                                     int16_t* __restrict__ coefs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_images * g.n_blocks) return;
    const int img = (int)(t / g.n_blocks), b = (int)(t - (long long)img * g.n_blocks);
    int comp, bx, by;
    block_pos(g, b, comp, bx, by);
    int16_t* out = coefs + (size_t)t * 64;
    if (comp == 0 && (bx >= g.wb || by >= g.hb)) {   // dummy block (jccoefct.c): AC = 0; its DC is resolved by the entropy coder
#pragma unroll
        for (int k = 0; k < 64; k += 8) *reinterpret_cast<uint4*>(out + k) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const uint8_t* src = images + (size_t)img * g.H * g.W * g.C;
    int d[64];
    if (g.ncomp == 1) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint8_t* row = src + (size_t)min(by * 8 + y, g.H - 1) * g.W;
#pragma unroll
            for (int x = 0; x < 8; ++x) d[y * 8 + x] = (int)row[min(bx * 8 + x, g.W - 1)] - 128;
        }
    } else if (comp == 0) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint8_t* row = src + (size_t)min(by * 8 + y, g.H - 1) * g.W * g.C;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const uint8_t* p = row + (size_t)min(bx * 8 + x, g.W - 1) * g.C;
                d[y * 8 + x] = ((19595 * p[0] + 38470 * p[1] + 7471 * p[2] + 32768) >> 16) - 128;
            }
        }
    } else {
        const int ch = (g.H + 1) >> 1;   // chroma rows that exist; below them the last one is replicated (expand_bottom_edge)
        const int k0 = comp == 1 ? -11059 : 32768, k1 = comp == 1 ? -21709 : -27439, k2 = comp == 1 ? 32768 : -5329;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int cy = min(by * 8 + y, ch - 1);
            const uint8_t* r0 = src + (size_t)(2 * cy) * g.W * g.C;
            const uint8_t* r1 = src + (size_t)min(2 * cy + 1, g.H - 1) * g.W * g.C;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const int cx = bx * 8 + x;
                const size_t x0 = (size_t)min(2 * cx, g.W - 1) * g.C, x1 = (size_t)min(2 * cx + 1, g.W - 1) * g.C;
                auto conv = [&](const uint8_t* p) { return (k0 * p[0] + k1 * p[1] + k2 * p[2] + (128 << 16) + 32767) >> 16; };
                d[y * 8 + x] = ((conv(r0 + x0) + conv(r0 + x1) + conv(r1 + x0) + conv(r1 + x1) + ((cx & 1) ? 2 : 1)) >> 2) - 128;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) fdct_pass<true>(d + 8 * r, 1);
#pragma unroll
    for (int c = 0; c < 8; ++c) fdct_pass<false>(d + c, 8);
    const int qt = comp ? 1 : 0;
    int16_t z[64];
    // zigzag position k <- natural index (static indices: d[] and z[] stay in registers)
#define Q(k, nat) { const int v = d[nat], dv = q.div[qt][nat]; const int mag = (abs(v) + (dv >> 1)) / dv; z[k] = (int16_t)(v < 0 ? -mag : mag); }
    Q(0, 0) Q(1, 1) Q(2, 8) Q(3, 16) Q(4, 9) Q(5, 2) Q(6, 3) Q(7, 10)
    Q(8, 17) Q(9, 24) Q(10, 32) Q(11, 25) Q(12, 18) Q(13, 11) Q(14, 4) Q(15, 5)
    Q(16, 12) Q(17, 19) Q(18, 26) Q(19, 33) Q(20, 40) Q(21, 48) Q(22, 41) Q(23, 34)
    Q(24, 27) Q(25, 20) Q(26, 13) Q(27, 6) Q(28, 7) Q(29, 14) Q(30, 21) Q(31, 28)
    Q(32, 35) Q(33, 42) Q(34, 49) Q(35, 56) Q(36, 57) Q(37, 50) Q(38, 43) Q(39, 36)
    Q(40, 29) Q(41, 22) Q(42, 15) Q(43, 23) Q(44, 30) Q(45, 37) Q(46, 44) Q(47, 51)
    Q(48, 58) Q(49, 59) Q(50, 52) Q(51, 45) Q(52, 38) Q(53, 31) Q(54, 39) Q(55, 46)
    Q(56, 53) Q(57, 60) Q(58, 61) Q(59, 54) Q(60, 47) Q(61, 55) Q(62, 62) Q(63, 63)
#undef Q
#pragma unroll
    for (int k = 0; k < 64; k += 8) {
        uint4 w;
        w.x = (uint16_t)z[k] | ((uint32_t)(uint16_t)z[k + 1] << 16); w.y = (uint16_t)z[k + 2] | ((uint32_t)(uint16_t)z[k + 3] << 16);
        w.z = (uint16_t)z[k + 4] | ((uint32_t)(uint16_t)z[k + 5] << 16); w.w = (uint16_t)z[k + 6] | ((uint32_t)(uint16_t)z[k + 7] << 16);
        *reinterpret_cast<uint4*>(out + k) = w;
    }
}

// DC value a block is coded with: its own, or for a dummy luma block the one of the block jccoefct.c copies it from
// (right edge: the block to its left; bottom edge: the MCU's Y01, itself possibly a right-edge dummy)
__device__ __forceinline__ int block_dc(const JpegGeom& g, const int16_t* __restrict__ img_coefs, int b) {
    if (g.ncomp == 3) {
        const int m = b / 6;
        int k = b - 6 * m;
        if (k < 4) {
            const int mx = m % g.mw, my = m / g.mw;
            if (2 * my + (k >> 1) >= g.hb) k = 1;
            if (2 * mx + (k & 1) >= g.wb) k -= 1;
            b = 6 * m + k;
        }
    }
    return img_coefs[(size_t)b * 64];
}
// previous block of the same component in scan order (-1: none, predictor 0)
__device__ __forceinline__ int prev_block(const JpegGeom& g, int b) {
    if (g.ncomp == 1) return b - 1;
    const int m = b / 6, k = b - 6 * m;
    if (k >= 1 && k < 4) return b - 1;
    if (m == 0) return -1;
    return k == 0 ? 6 * (m - 1) + 3 : b - 6;
}

// jchuff.c:encode_one_block as a token walk: `put(code, length)` for every code word
template <class Put>
__device__ __forceinline__ void encode_block(const JpegGeom& g, const int16_t* __restrict__ img_coefs, int b, Put put) {
    const int comp = g.ncomp == 1 ? 0 : ((b % 6) < 4 ? 0 : 1);
    const uint32_t* dc_lut = c_huff[comp ? 2 : 0];
    const uint32_t* ac_lut = c_huff[comp ? 3 : 1];
    const int pb = prev_block(g, b);
    const int dc = block_dc(g, img_coefs, b), diff = dc - (pb >= 0 ? block_dc(g, img_coefs, pb) : 0);
    {
        const int mag = abs(diff), n = 32 - __clz(mag);
        const uint32_t e = dc_lut[n];
        put(e >> 8, e & 0xffu);
        if (n) put((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << n) - 1u), n);
    }
    const uint4* cp = reinterpret_cast<const uint4*>(img_coefs + (size_t)b * 64);
    int run = 0;
#pragma unroll 1
    for (int k8 = 0; k8 < 8; ++k8) {
        const uint4 w = __ldg(cp + k8);
        if (k8 && !(w.x | w.y | w.z | w.w)) { run += 8; continue; }
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k8 == 0 && j == 0) continue;
            const int v = (int)(int16_t)(ws[j >> 1] >> ((j & 1) * 16));
            if (v == 0) { ++run; continue; }
            while (run > 15) { const uint32_t e = ac_lut[0xF0]; put(e >> 8, e & 0xffu); run -= 16; }
            const int mag = abs(v), n = 32 - __clz(mag);
            const uint32_t e = ac_lut[(run << 4) + n];
            put(e >> 8, e & 0xffu);
            put((uint32_t)(v < 0 ? v - 1 : v) & ((1u << n) - 1u), n);
            run = 0;
        }
    }
    if (run) { const uint32_t e = ac_lut[0]; put(e >> 8, e & 0xffu); }
}

// one block of 1024 threads per image: bit length of every block, exclusive prefix sum -> bit_off[img][b], total bits
__global__ void __launch_bounds__(1024) k_jpeg_scan(JpegGeom g, const int16_t* __restrict__ coefs, uint32_t* __restrict__ bit_off,
                                                    uint32_t* __restrict__ total_bits) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int img = blockIdx.x;
    const int16_t* ic = coefs + (size_t)img * g.n_blocks * 64;
    uint32_t* off = bit_off + (size_t)img * g.n_blocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < g.n_blocks; base += 1024) {
        const int b = base + threadIdx.x;
        uint32_t len = 0;
        if (b < g.n_blocks) encode_block(g, ic, b, [&](uint32_t, uint32_t n) { len += n; });
        uint32_t v = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += u; }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry, wbase = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        if (b < g.n_blocks) off[b] = carry + wbase + v - len;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) total_bits[img] = s_carry;
}

// one thread per block: code words ORed into the image's MSB-first bit stream (32-bit words, stream_words per image, zeroed)
__global__ void __launch_bounds__(128) k_jpeg_emit(JpegGeom g, int n_images, const int16_t* __restrict__ coefs, const uint32_t* __restrict__ bit_off,
                                                   uint32_t* __restrict__ stream, size_t stream_words) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_images * g.n_blocks) return;
    const int img = (int)(t / g.n_blocks), b = (int)(t - (long long)img * g.n_blocks);
    uint32_t* words = stream + (size_t)img * stream_words;
    const uint32_t start = bit_off[t];
    uint32_t word = start >> 5;
    int fill = (int)(start & 31u);            // bits of `acc` in use (the leading ones belong to the previous block: zero here)
    unsigned long long acc = 0ull;
    bool first = true;
    encode_block(g, coefs + (size_t)img * g.n_blocks * 64, b, [&](uint32_t code, uint32_t n) {
        acc = (acc << n) | code;
        fill += (int)n;
        if (fill >= 32) {
            const uint32_t out = (uint32_t)(acc >> (fill - 32));
            if (first) { atomicOr(words + word, out); first = false; } else words[word] = out;   // interior words have one writer
            ++word;
            fill -= 32;
            acc &= (1ull << fill) - 1ull;
        }
    });
    if (fill) atomicOr(words + word, (uint32_t)(acc << (32 - fill)));
}

// one block per image: pad, stuff, frame
__global__ void __launch_bounds__(1024) k_jpeg_finish(const uint32_t* __restrict__ stream, size_t stream_words, const uint32_t* __restrict__ total_bits,
                                                      const uint8_t* __restrict__ header, int header_len, uint8_t* __restrict__ out, size_t out_stride,
                                                      uint32_t* __restrict__ sizes) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int img = blockIdx.x;
    const uint32_t* words = stream + (size_t)img * stream_words;
    uint8_t* o = out + (size_t)img * out_stride;
    for (int i = threadIdx.x; i < header_len && (size_t)i < out_stride; i += 1024) o[i] = header[i];
    const uint32_t bits = total_bits[img], n_bytes = (bits + 7u) >> 3, pad = (8u - (bits & 7u)) & 7u;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    uint8_t* body = o + header_len;
    for (uint32_t base = 0; base < n_bytes; base += 1024) {
        const uint32_t j = base + threadIdx.x;
        uint32_t byte = 0, cnt = 0;
        if (j < n_bytes) {
            byte = (__ldg(words + (j >> 2)) >> (24 - 8 * (j & 3))) & 0xffu;
            if (j == n_bytes - 1) byte |= (1u << pad) - 1u;   // flush_bits: fill with ones
            cnt = byte == 0xffu ? 2u : 1u;
        }
        uint32_t v = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, s); if ((threadIdx.x & 31) >= s) v += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, w, s); if (threadIdx.x >= s) w += u; }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry, wbase = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        if (j < n_bytes) {
            const uint32_t at = carry + wbase + v - cnt;
            if ((size_t)header_len + at + cnt <= out_stride) {   // (a file that does not fit is reported below, never written past its stride)
                body[at] = (uint8_t)byte;
                if (cnt == 2u) body[at + 1] = 0;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t end = s_carry;
        if ((size_t)header_len + end + 2u <= out_stride) {
            body[end] = 0xFF; body[end + 1] = 0xD9;   // EOI
            sizes[img] = (uint32_t)header_len + end + 2u;
        } else sizes[img] = 0u;
    }
}

// ---- host side ---------------------------------------------------------------------------------
static JpegGeom make_geom(int H, int W, int channels) {
    JpegGeom g;
    g.W = W; g.H = H; g.C = channels; g.ncomp = channels == 1 ? 1 : 3;
    g.wb = (W + 7) / 8; g.hb = (H + 7) / 8;
    if (g.ncomp == 1) { g.mw = g.wb; g.mh = g.hb; g.n_blocks = g.mw * g.mh; }
    else { g.mw = (W + 15) / 16; g.mh = (H + 15) / 16; g.n_blocks = g.mw * g.mh * 6; }
    return g;
}
// Worst case of one block: DC 9 + 11 bits, 63 x (16 + 10) bits of AC = 1658 bits -> 208 bytes before stuffing.
size_t jpeg_stream_words(int H, int W, int channels) { return ((size_t)make_geom(H, W, channels).n_blocks * 208 + 3) / 4 + 1; }
size_t jpeg_coef_bytes(int H, int W, int channels) { return (size_t)make_geom(H, W, channels).n_blocks * 64 * sizeof(int16_t); }
size_t jpeg_blocks(int H, int W, int channels) { return (size_t)make_geom(H, W, channels).n_blocks; }
size_t jpeg_file_bound(int H, int W, int channels) { return 1024 + jpeg_stream_words(H, W, channels) * 4 * 2 + 2; }   // every byte stuffed

// tables for `quality` into constant memory; returns the file header (jcmarker.c: write_file_header, write_frame_header,
// write_scan_header) for images of this shape
std::vector<uint8_t> jpeg_prepare(int H, int W, int channels, int quality, JpegQuant* quant, cudaStream_t s) {
    const int ncomp = channels == 1 ? 1 : 3;
    uint8_t q[2][64];
    quant_table(kLumaQ, quality, q[0]);
    quant_table(kChromaQ, quality, q[1]);
    for (int t = 0; t < 2; ++t) for (int i = 0; i < 64; ++i) quant->div[t][i] = (uint16_t)(q[t][i] * 8);   // 8 * Q in NATURAL order (jcdctmgr.c divisors)
    uint32_t lut[4][256];
    derive(kDcLumaBits, kDcVals, lut[0]); derive(kAcLumaBits, kAcLumaVals, lut[1]);
    derive(kDcChromaBits, kDcVals, lut[2]); derive(kAcChromaBits, kAcChromaVals, lut[3]);
    cudaMemcpyToSymbolAsync(c_huff, lut, sizeof(lut), 0, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);   // the host array above is stack storage

    std::vector<uint8_t> h = {0xFF, 0xD8};
    put_marker(h, 0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < (ncomp == 1 ? 1 : 2); ++t) {
        std::vector<uint8_t> p = {(uint8_t)t};
        for (int k = 0; k < 64; ++k) p.push_back(q[t][kZigzag[k]]);
        put_marker(h, 0xDB, p);
    }
    std::vector<uint8_t> sof = {8, (uint8_t)(H >> 8), (uint8_t)H, (uint8_t)(W >> 8), (uint8_t)W};
    if (ncomp == 1) sof.insert(sof.end(), {1, 1, 0x11, 0});
    else sof.insert(sof.end(), {3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    put_marker(h, 0xC0, sof);
    auto dht = [&](uint8_t id, const uint8_t* bits, const uint8_t* vals, int nvals) {
        std::vector<uint8_t> p = {id};
        p.insert(p.end(), bits, bits + 16);
        p.insert(p.end(), vals, vals + nvals);
        put_marker(h, 0xC4, p);
    };
    dht(0x00, kDcLumaBits, kDcVals, 12);
    dht(0x10, kAcLumaBits, kAcLumaVals, 162);
    if (ncomp == 3) { dht(0x01, kDcChromaBits, kDcVals, 12); dht(0x11, kAcChromaBits, kAcChromaVals, 162); }
    if (ncomp == 1) put_marker(h, 0xDA, {1, 1, 0x00, 0, 63, 0});
    else put_marker(h, 0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    return h;
}

void launch_jpeg_encode(const uint8_t* images, int n, int H, int W, int channels, const JpegQuant& quant, int16_t* coefs, uint32_t* bit_off, uint32_t* total_bits,
                        uint32_t* stream, const uint8_t* header_dev, int header_len, uint8_t* out, size_t out_stride, uint32_t* sizes,
                        cudaStream_t s) {
    const JpegGeom g = make_geom(H, W, channels);
    const size_t words = jpeg_stream_words(H, W, channels);
    const long long total = (long long)n * g.n_blocks;
    cudaMemsetAsync(stream, 0, (size_t)n * words * 4, s);
    k_jpeg_blocks<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(images, g, quant, n, coefs);
    k_jpeg_scan<<<n, 1024, 0, s>>>(g, coefs, bit_off, total_bits);
    k_jpeg_emit<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(g, n, coefs, bit_off, stream, words);
    k_jpeg_finish<<<n, 1024, 0, s>>>(stream, words, total_bits, header_dev, header_len, out, out_stride, sizes);
}

}  // namespace slbk
