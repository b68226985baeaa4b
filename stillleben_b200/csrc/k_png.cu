// k_png.cu — batched PNG encoder on the device (reference: src/image_saver.cpp + python/src/py_image_saver.cpp:37-99,
// a pool of CPU threads each running libpng through Magnum's AnyImageConverter; SURVEY 8(f-4): "at > 10 k fps the
// saver becomes the bottleneck"). Every image of a batch becomes a complete, standard PNG file in device memory:
//   k_png_rows      one thread per (image, scanline, 512-byte segment): PNG filter 1 (Sub) + deflate with the fixed
//                   Huffman code and run matches at distance 1 / bytes-per-pixel, written as one byte-aligned
//                   deflate block per segment (fixed block, end-of-block, empty stored block = a zlib "sync
//                   flush"), plus the block's Adler-32 partial sums and the CRC-32 of its compressed bytes;
//   k_png_finalize  one block per image: block offsets (prefix sum), ordered tree combination of the Adler-32 /
//                   CRC-32 parts (GF(2) polynomial arithmetic as in zlib's crc32_combine), signature, IHDR, IDAT
//                   header / trailer, IEND;
//   k_png_gather    one block per deflate block: moves its bytes to their place in the file.
// Formats as the reference's binding accepts them: uint8 HxW, HxWx3, HxWx4 and 16-bit HxW (big-endian samples).
// Row 0 of the tensor is the top row of the file (the binding's flipud + Magnum's bottom-up rows cancel).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"

namespace {

#define PNG_POLY 0xEDB88320u

__constant__ uint32_t c_crc_table[256];
__constant__ uint32_t c_x2n[32];               // x^(2^k) mod P, reflected (zlib crc32.c x2n_table)
__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {   // a(x) * b(x) mod P, reflected bit order
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ PNG_POLY : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t x2nmodp(uint64_t n, unsigned k) {    // x^(n * 2^k) mod P
    uint32_t p = 1u << 31;
    while (n) { if (n & 1) p = multmodp(c_x2n[k & 31], p); n >>= 1; ++k; }
    return p;
}
__device__ __forceinline__ uint32_t crc_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) { return multmodp(x2nmodp(len2, 3), crc1) ^ crc2; }
__device__ __forceinline__ uint32_t crc_bytes(uint32_t crc, const uint8_t* p, size_t n) {
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = c_crc_table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

struct BitWriter {   // LSB-first bit packing into the scanline's byte range
    uint8_t* out; size_t pos; uint64_t acc; int nbits;
    __device__ __forceinline__ void put(uint32_t v, int n) {
        acc |= (uint64_t)v << nbits; nbits += n;
        while (nbits >= 8) { out[pos++] = (uint8_t)acc; acc >>= 8; nbits -= 8; }
    }
    __device__ __forceinline__ void align() { if (nbits) { out[pos++] = (uint8_t)acc; acc = 0; nbits = 0; } }
};
__device__ __forceinline__ void put_literal(BitWriter& w, uint32_t lit) {        // RFC 1951 3.2.6
    if (lit < 144) w.put(rev_bits(0x30 + lit, 8), 8); else w.put(rev_bits(0x190 + (lit - 144), 9), 9);
}
__device__ __forceinline__ void put_match(BitWriter& w, int len, int dist) {     // dist in 1..4: codes 0..3, no extra bits
    int c = 0;
    while (c < 28 && (int)c_len_base[c + 1] <= len) ++c;
    const int sym = 257 + c;
    if (sym < 280) w.put(rev_bits(sym - 256, 7), 7); else w.put(rev_bits(0xC0 + (sym - 280), 8), 8);
    if (c_len_extra[c]) w.put((uint32_t)(len - c_len_base[c]), c_len_extra[c]);
    w.put(rev_bits((uint32_t)(dist - 1), 5), 5);
}

struct RowInfo { uint32_t bytes, crc, adler_a, len; uint64_t adler_b; };

// sample byte b of pixel x of the scanline as the PNG stores it (16-bit samples big-endian)
__device__ __forceinline__ uint8_t raw_byte(const uint8_t* row, int i, int bpc) { return bpc == 2 ? row[i ^ 1] : row[i]; }

// one thread per (image, scanline, segment): a scanline is cut into segments of PNG_SEG data bytes, each its own
// byte-aligned deflate block, so that a batch of 64 VGA frames exposes > 100 k independent encoders
#define PNG_SEG 512
__global__ void __launch_bounds__(128) k_png_rows(const uint8_t* __restrict__ images, int n, int H, int W, int channels, int bpc, int n_seg,
                                                  uint8_t* __restrict__ rows_out, size_t seg_bound, RowInfo* __restrict__ info) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * H * n_seg) return;
    const size_t rowi = t / n_seg;
    const int seg = (int)(t - rowi * n_seg);
    const int bpp = channels * bpc;
    const size_t row_bytes = (size_t)W * bpp;
    const uint8_t* row = images + rowi * row_bytes;
    const size_t beg = (size_t)seg * PNG_SEG, end = min(beg + (size_t)PNG_SEG, row_bytes);
    BitWriter w;
    w.out = rows_out + t * seg_bound; w.pos = 0; w.acc = 0; w.nbits = 0;
    w.put(0u, 1); w.put(1u, 2);                               // BFINAL = 0, BTYPE = 01 (fixed Huffman)
    // filtered scanline: filter-type byte 1 (Sub) in front of the first segment, then raw[i] - raw[i - bpp]
    const size_t seg_len = (end - beg) + (seg == 0 ? 1 : 0);  // filtered bytes this block carries
    uint64_t sa = 0, sb = 0;                                  // Adler partial sums of this block
    if (seg == 0) { put_literal(w, 1u); sa = 1; sb = seg_len; }
    auto filt = [&](size_t i) -> uint32_t {
        const uint32_t cur = raw_byte(row, (int)i, bpc), left = i >= (size_t)bpp ? raw_byte(row, (int)(i - bpp), bpc) : 0u;
        return (cur - left) & 0xffu;
    };
    size_t i = beg;
    while (i < end) {
        const uint32_t v = filt(i);
        // run matches inside the segment: the filtered bytes repeat with period 1 (flat colour) or bpp (constant gradient)
        int best = 0, dist = 0;
        for (int d = 1; d <= bpp; d += (bpp > 1 ? bpp - 1 : 1)) {
            if (i >= beg + (size_t)d) {
                int l = 0;
                while (l < 258 && i + l < end && filt(i + l) == filt(i + l - d)) ++l;
                if (l > best) { best = l; dist = d; }
            }
            if (bpp == 1) break;
        }
        if (best >= 4) {
            put_match(w, best, dist);
            for (int k = 0; k < best; ++k) { const uint32_t b = filt(i + k); sa += b; sb += (uint64_t)(end - (i + k)) * b; }
            i += best;
        } else {
            put_literal(w, v);
            sa += v; sb += (uint64_t)(end - i) * v;
            ++i;
        }
    }
    w.put(0u, 7);                                             // end of block (symbol 256)
    w.put(0u, 1); w.put(0u, 2); w.align();                    // empty stored block: byte alignment ("sync flush")
    w.out[w.pos++] = 0x00; w.out[w.pos++] = 0x00; w.out[w.pos++] = 0xFF; w.out[w.pos++] = 0xFF;
    RowInfo ri;
    ri.bytes = (uint32_t)w.pos;
    ri.crc = crc_bytes(0u, w.out, w.pos);
    ri.adler_a = (uint32_t)(sa % 65521u);                     // sums over this block's bytes only (Adler's initial 1 comes later)
    ri.adler_b = sb % 65521u;
    ri.len = (uint32_t)seg_len;
    info[t] = ri;
}

// ---------------------------------------------------------------------------------------------
// Warp-per-segment encoder: the same greedy parse and the same bytes as k_png_rows, computed in parallel.
//   1. every lane filters 16 bytes of the segment into shared memory;
//   2. run lengths at distance 1 and bpp for every position: "next mismatch" by a per-lane backward walk + a warp
//      suffix-min over the lanes; best(i), nxt(i) = i + (best >= 4 ? best : 1);
//   3. the greedy token starts are the orbit of position 0 under nxt: pointer jumping (10 doubling rounds);
//   4. token bit lengths -> warp prefix sum -> every lane ORs its tokens into the block image in shared memory;
//   5. CRC-32 of the block by right-aligned per-lane chunks (the register of a chunk is shifted into place with one
//      multiplication by x^(8 c m) from a table), Adler-32 partial sums by a warp reduction.
// A byte-serial thread per segment runs at 4 of 32 lanes (profiles/r01_d_aux_kernels.md); this keeps all lanes busy.
#define PNG_WARPS 8
#define PNG_MAX_CHUNK 20                                 // ceil(max block bytes / 32)
__constant__ uint32_t c_pow[PNG_MAX_CHUNK + 1][32];      // x^(8 c m) mod P, reflected: shifts a CRC register over c*m bytes
struct SegShared {
    uint8_t f[PNG_SEG + 16];
    uint16_t jump[PNG_SEG + 2];
    uint16_t len[PNG_SEG + 2];                            // best match length at the position (0..258), bit 15: distance is bpp
    uint8_t mark[PNG_SEG + 2];
    uint32_t out[160];
};
__device__ __forceinline__ void or_bits(uint32_t* out, uint32_t bitpos, uint32_t v, int nbits) {
    const uint32_t w = bitpos >> 5, sh = bitpos & 31;
    atomicOr(out + w, v << sh);
    if (sh + nbits > 32) atomicOr(out + w + 1, v >> (32 - sh));
}
__global__ void __launch_bounds__(PNG_WARPS * 32) k_png_rows_warp(const uint8_t* __restrict__ images, int n_images, int H, int W, int channels,
                                                                  int bpc, int n_seg, uint8_t* __restrict__ rows_out, size_t seg_bound,
                                                                  RowInfo* __restrict__ info) {
    __shared__ SegShared s_all[PNG_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t t = (size_t)blockIdx.x * PNG_WARPS + warp;
    if (t >= (size_t)n_images * H * n_seg) return;
    SegShared& S = s_all[warp];
    const size_t rowi = t / n_seg;
    const int seg = (int)(t - rowi * n_seg);
    const int bpp = channels * bpc;
    const size_t row_bytes = (size_t)W * bpp;
    const uint8_t* row = images + rowi * row_bytes;
    const int beg = seg * PNG_SEG, n = (int)min((size_t)PNG_SEG, row_bytes - (size_t)beg);   // data bytes of this block
    // 1. filtered bytes
    for (int r = lane; r < n; r += 32) {
        const int i = beg + r;
        const uint32_t cur = raw_byte(row, i, bpc), left = i >= bpp ? raw_byte(row, i - bpp, bpc) : 0u;
        S.f[r] = (uint8_t)(cur - left);
    }
    for (int w = lane; w < 160; w += 32) S.out[w] = 0u;
    __syncwarp();
    // 2. run lengths: nz_d[r] = first position j >= r where f[j] != f[j - d] (or j < d, or j == n)
    const int r0 = lane * 16;
    int best_len[16];
    {
        int nz1 = PNG_SEG + 1, nzb = PNG_SEG + 1;   // first mismatch inside this lane's 16 positions
        for (int k = 15; k >= 0; --k) {
            const int r = r0 + k;
            if (r >= n) { nz1 = nzb = min(r, n); continue; }
            if (r < 1 || S.f[r] != S.f[r - 1]) nz1 = r;
            if (r < bpp || S.f[r] != S.f[r - bpp]) nzb = r;
        }
        // suffix-min over the lanes: the first mismatch at or after the START of lane l's block
        int s1 = nz1, sb_ = nzb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_down_sync(0xffffffffu, s1, o), b = __shfl_down_sync(0xffffffffu, sb_, o);
            if (lane + o < 32) { s1 = min(s1, a); sb_ = min(sb_, b); }
        }
        int c1 = __shfl_down_sync(0xffffffffu, s1, 1), cb = __shfl_down_sync(0xffffffffu, sb_, 1);   // first mismatch after this lane's block
        if (lane == 31) { c1 = n; cb = n; }
        c1 = min(c1, n); cb = min(cb, n);
        for (int k = 15; k >= 0; --k) {
            const int r = r0 + k;
            best_len[k] = 0;
            if (r >= n) continue;
            if (r < 1 || S.f[r] != S.f[r - 1]) c1 = r;
            if (r < bpp || S.f[r] != S.f[r - bpp]) cb = r;
            const int l1 = min(258, c1 - r), lb = bpp > 1 ? min(258, cb - r) : 0;
            int best = l1, far = 0;
            if (lb > l1) { best = lb; far = 1; }
            best_len[k] = best;
            S.len[r] = (uint16_t)(best | (far << 15));
            S.jump[r] = (uint16_t)min(n, r + (best >= 4 ? best : 1));
            S.mark[r] = r == 0 ? 1 : 0;
        }
        if (lane == 0) { S.jump[n] = (uint16_t)n; S.mark[n] = 0; }
    }
    __syncwarp();
    // 3. orbit of position 0 under nxt by pointer jumping
    for (int round = 0; round < 10; ++round) {
        int nj[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int r = r0 + k;
            nj[k] = n;
            if (r < n) {
                const int j = S.jump[r];
                if (S.mark[r]) S.mark[j] = 1;
                nj[k] = S.jump[j];
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) if (r0 + k < n) S.jump[r0 + k] = (uint16_t)nj[k];
        __syncwarp();
    }
    // 4. token codes and bit offsets
    uint32_t code[16]; uint8_t nb[16];
    int local_bits = 0;
    uint32_t sa = 0; unsigned long long sbw = 0;      // Adler partial sums over this lane's filtered bytes
    const int L = n + (seg == 0 ? 1 : 0);              // filtered bytes this block carries
    const int first = seg == 0 ? 1 : 0;                // index of data byte 0 inside the block
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = r0 + k;
        nb[k] = 0; code[k] = 0;
        if (r >= n) continue;
        const uint32_t v = S.f[r];
        sa += v; sbw += (unsigned long long)(L - (r + first)) * v;
        if (!S.mark[r]) continue;
        const int len = best_len[k];
        if (len >= 4) {
            int c = 0;
            while (c < 28 && (int)c_len_base[c + 1] <= len) ++c;
            const int sym = 257 + c;
            uint32_t bits; int nbits;
            if (sym < 280) { bits = rev_bits(sym - 256, 7); nbits = 7; } else { bits = rev_bits(0xC0 + (sym - 280), 8); nbits = 8; }
            const int eb = c_len_extra[c];
            if (eb) { bits |= (uint32_t)(len - c_len_base[c]) << nbits; nbits += eb; }
            const int dist = (S.len[r] >> 15) ? bpp : 1;
            bits |= rev_bits((uint32_t)(dist - 1), 5) << nbits; nbits += 5;
            code[k] = bits; nb[k] = (uint8_t)nbits;
        } else if (v < 144) { code[k] = rev_bits(0x30 + v, 8); nb[k] = 8; }
        else { code[k] = rev_bits(0x190 + (v - 144), 9); nb[k] = 9; }
        local_bits += nb[k];
    }
    int inc = local_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += a; }
    const int total_bits = __shfl_sync(0xffffffffu, inc, 31);
    const int head_bits = 3 + (seg == 0 ? 8 : 0);
    uint32_t bitpos = (uint32_t)(head_bits + inc - local_bits);
#pragma unroll
    for (int k = 0; k < 16; ++k) if (nb[k]) { or_bits(S.out, bitpos, code[k], nb[k]); bitpos += nb[k]; }
    if (lane == 0) {
        or_bits(S.out, 0u, 2u, 3);                                             // BFINAL = 0, BTYPE = 01
        if (seg == 0) or_bits(S.out, 3u, rev_bits(0x30 + 1, 8), 8);            // filter-type byte 1 as a literal
    }
    // end of block (7 zero bits), empty stored block header (3 zero bits), pad to a byte, 00 00 FF FF
    const int nbytes = (head_bits + total_bits + 10 + 7) >> 3;
    const int total = nbytes + 4;
    __syncwarp();
    if (lane == 0) {
        const uint32_t o = (uint32_t)(nbytes + 2) * 8u;
        or_bits(S.out, o, 0xFFFFu, 16);
    }
    __syncwarp();
    // 5. copy out, CRC-32, Adler-32
    const uint8_t* ob = reinterpret_cast<const uint8_t*>(S.out);
    uint8_t* dst = rows_out + t * seg_bound;
    for (int i = lane; i < total; i += 32) dst[i] = ob[i];
    const int c = (total + 31) >> 5;                                           // chunk bytes per lane, right aligned
    const int stop = total - (31 - lane) * c, start = stop - c;
    uint32_t reg = (start <= 0 && stop > 0) ? 0xFFFFFFFFu : 0u;                // the chunk holding byte 0 carries the initial register
    for (int i = max(start, 0); i < stop; ++i) reg = c_crc_table[(reg ^ ob[i]) & 0xff] ^ (reg >> 8);
    if (stop <= 0) reg = 0u;
    uint32_t shifted = multmodp(c_pow[c][31 - lane], reg);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) shifted ^= __shfl_xor_sync(0xffffffffu, shifted, o);
    if (seg == 0 && lane == 0) { sa += 1u; sbw += (unsigned long long)L; }      // the filter byte (value 1, first in the block)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sbw += __shfl_xor_sync(0xffffffffu, sbw, o); }
    if (lane == 0) {
        RowInfo ri;
        ri.bytes = (uint32_t)total;
        ri.crc = ~shifted;
        ri.adler_a = sa % 65521u;
        ri.adler_b = sbw % 65521u;
        ri.len = (uint32_t)L;
        info[t] = ri;
    }
}

__device__ __forceinline__ void put_be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

// One block per image. CRC-32 and Adler-32 of a concatenation are associative combinations of the parts
// ((crc, length) and (sum, weighted sum, length) monoids), so the image's deflate blocks are reduced by a tree in
// shared memory; their file offsets are a prefix sum of their sizes.
#define PNG_FIN_THREADS 256
__global__ void __launch_bounds__(PNG_FIN_THREADS) k_png_finalize(const RowInfo* __restrict__ info, int H, int W, int channels, int bpc,
                                                                  int n_seg, uint8_t* __restrict__ out, size_t out_stride,
                                                                  uint32_t* __restrict__ sizes, uint32_t* __restrict__ row_offset) {
    __shared__ uint32_t s_crc[PNG_FIN_THREADS], s_bytes[PNG_FIN_THREADS], s_a[PNG_FIN_THREADS], s_b[PNG_FIN_THREADS];
    __shared__ uint64_t s_len[PNG_FIN_THREADS];
    __shared__ uint32_t s_base[PNG_FIN_THREADS];
    const int img = blockIdx.x, tid = threadIdx.x;
    const int n_blocks = H * n_seg;
    const int per = (n_blocks + PNG_FIN_THREADS - 1) / PNG_FIN_THREADS;
    const int b0 = min(tid * per, n_blocks), b1 = min(b0 + per, n_blocks);
    const RowInfo* my = info + (size_t)img * n_blocks;
    // sequential combination of this thread's run of blocks
    uint32_t crc = 0, bytes = 0; uint64_t a = 0, b = 0, len = 0;
    for (int r = b0; r < b1; ++r) {
        const RowInfo ri = my[r];
        crc = bytes ? crc_combine(crc, ri.crc, ri.bytes) : ri.crc;
        bytes += ri.bytes;
        b = (b + (uint64_t)ri.len % 65521u * a + ri.adler_b) % 65521u;
        a = (a + ri.adler_a) % 65521u;
        len += ri.len;
    }
    s_crc[tid] = crc; s_bytes[tid] = bytes; s_a[tid] = (uint32_t)a; s_b[tid] = (uint32_t)b; s_len[tid] = len;
    __syncthreads();
    if (tid == 0) {   // exclusive prefix of the byte counts (256 adds)
        uint32_t run = 0;
        for (int t = 0; t < PNG_FIN_THREADS; ++t) { s_base[t] = run; run += s_bytes[t]; }
    }
    __syncthreads();
    {   // file offsets of this thread's blocks: 33 (signature + IHDR) + 10 (IDAT length, type, zlib header) + prefix
        uint32_t off = 43u + s_base[tid];
        for (int r = b0; r < b1; ++r) { row_offset[(size_t)img * n_blocks + r] = off; off += my[r].bytes; }
    }
    // ordered tree reduction: element t absorbs element t + stride (which follows it in the stream)
    for (int stride = 1; stride < PNG_FIN_THREADS; stride <<= 1) {
        if ((tid & (2 * stride - 1)) == 0) {
            const int o = tid + stride;
            if (s_bytes[o]) s_crc[tid] = s_bytes[tid] ? crc_combine(s_crc[tid], s_crc[o], s_bytes[o]) : s_crc[o];
            s_bytes[tid] += s_bytes[o];
            s_b[tid] = (uint32_t)((s_b[tid] + (s_len[o] % 65521u) * s_a[tid] + s_b[o]) % 65521u);
            s_a[tid] = (s_a[tid] + s_a[o]) % 65521u;
            s_len[tid] += s_len[o];
        }
        __syncthreads();
    }
    if (tid != 0) return;
    uint8_t* f = out + (size_t)img * out_stride;
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    for (int i = 0; i < 8; ++i) f[i] = sig[i];
    put_be32(f + 8, 13u);
    uint8_t* ih = f + 12;
    ih[0] = 'I'; ih[1] = 'H'; ih[2] = 'D'; ih[3] = 'R';
    put_be32(ih + 4, (uint32_t)W); put_be32(ih + 8, (uint32_t)H);
    ih[12] = (uint8_t)(8 * bpc);
    ih[13] = channels == 1 ? 0 : channels == 3 ? 2 : 6;       // colour type: grey, RGB, RGBA
    ih[14] = 0; ih[15] = 0; ih[16] = 0;
    put_be32(f + 29, crc_bytes(0u, ih, 17));
    // IDAT: zlib header, the deflate blocks, final empty stored block, Adler-32
    uint8_t* idat = f + 33;                                   // length field at +0, type at +4, data from +8
    idat[4] = 'I'; idat[5] = 'D'; idat[6] = 'A'; idat[7] = 'T';
    idat[8] = 0x78; idat[9] = 0x01;
    const uint32_t body = s_bytes[0];
    uint32_t c = crc_combine(crc_bytes(0u, idat + 4, 6), s_crc[0], body);
    const uint32_t A = (1u + s_a[0]) % 65521u, B = (uint32_t)((s_len[0] % 65521u + s_b[0]) % 65521u);   // Adler starts at A = 1, B = 0
    uint8_t* tail = idat + 10 + body;
    tail[0] = 0x01; tail[1] = 0x00; tail[2] = 0x00; tail[3] = 0xFF; tail[4] = 0xFF;   // BFINAL = 1 stored, empty
    put_be32(tail + 5, (B << 16) | A);
    c = crc_combine(c, crc_bytes(0u, tail, 9), 9);
    put_be32(tail + 9, c);
    put_be32(idat, 2u + body + 9u);                           // IDAT data length
    uint8_t* iend = tail + 13;
    put_be32(iend, 0u);
    iend[4] = 'I'; iend[5] = 'E'; iend[6] = 'N'; iend[7] = 'D';
    put_be32(iend + 8, 0xAE426082u);
    sizes[img] = 43u + body + 13u + 12u;
}

__global__ void __launch_bounds__(128) k_png_gather(const uint8_t* __restrict__ rows_out, size_t seg_bound, const RowInfo* __restrict__ info,
                                                    const uint32_t* __restrict__ row_offset, int blocks_per_image, uint8_t* __restrict__ out,
                                                    size_t out_stride) {
    const size_t t = blockIdx.x;                              // (image * H + row) * n_seg + segment
    const uint32_t nbytes = info[t].bytes;
    const uint8_t* src = rows_out + t * seg_bound;
    uint8_t* dst = out + (t / blocks_per_image) * out_stride + row_offset[t];
    for (uint32_t i = threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

namespace slbk {

void png_upload_tables() {
    uint32_t table[256];
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ PNG_POLY : c >> 1;
        table[i] = c;
    }
    cudaMemcpyToSymbol(c_crc_table, table, sizeof table);
    // x^(2^k) mod P by repeated squaring (host copy of multmodp)
    auto mul = [](uint32_t a, uint32_t b) {
        uint32_t m = 1u << 31, p = 0;
        for (;;) {
            if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
            m >>= 1;
            b = (b & 1) ? (b >> 1) ^ PNG_POLY : b >> 1;
        }
        return p;
    };
    uint32_t x2n[32], p = 1u << 30;
    x2n[0] = p;
    for (int k = 1; k < 32; ++k) x2n[k] = p = mul(p, p);
    cudaMemcpyToSymbol(c_x2n, x2n, sizeof x2n);
    // x^(8 c m) for the warp encoder's chunk shifts: x^8 by squaring x three times, then powers by multiplication
    static uint32_t pw[PNG_MAX_CHUNK + 1][32];
    const uint32_t x8 = x2n[3];
    for (int c = 0; c <= PNG_MAX_CHUNK; ++c) {
        uint32_t xc = 1u << 31;                                   // x^0
        for (int i = 0; i < c; ++i) xc = mul(xc, x8);             // x^(8 c)
        uint32_t acc = 1u << 31;
        for (int m = 0; m < 32; ++m) { pw[c][m] = acc; acc = mul(acc, xc); }
    }
    cudaMemcpyToSymbol(c_pow, pw, sizeof pw);
}
int png_segments(int W, int channels, int bpc) { return (int)(((size_t)W * channels * bpc + PNG_SEG - 1) / PNG_SEG); }
size_t png_seg_bound() { return ((size_t)(PNG_SEG + 1) * 9 + 7) / 8 + 16; }
size_t png_file_bound(int H, int W, int channels, int bpc) {
    return 33 + 10 + (size_t)H * png_segments(W, channels, bpc) * png_seg_bound() + 13 + 12 + 16;
}
size_t png_row_info_bytes() { return sizeof(RowInfo); }
void launch_png_encode(const uint8_t* images, int n, int H, int W, int channels, int bpc, uint8_t* rows_scratch, void* row_info,
                       uint32_t* row_offset, uint8_t* out, size_t out_stride, uint32_t* sizes, cudaStream_t s) {
    const int n_seg = png_segments(W, channels, bpc);
    const size_t blocks = (size_t)n * H * n_seg, sb = png_seg_bound();
    static const bool serial = getenv("SLB_PNG_SERIAL") != nullptr;   // the byte-serial reference encoder (same bytes)
    if (serial) k_png_rows<<<(unsigned)((blocks + 127) / 128), 128, 0, s>>>(images, n, H, W, channels, bpc, n_seg, rows_scratch, sb, (RowInfo*)row_info);
    else k_png_rows_warp<<<(unsigned)((blocks + PNG_WARPS - 1) / PNG_WARPS), PNG_WARPS * 32, 0, s>>>(images, n, H, W, channels, bpc, n_seg, rows_scratch, sb, (RowInfo*)row_info);
    k_png_finalize<<<n, PNG_FIN_THREADS, 0, s>>>((const RowInfo*)row_info, H, W, channels, bpc, n_seg, out, out_stride, sizes, row_offset);
    k_png_gather<<<(unsigned)blocks, 128, 0, s>>>(rows_scratch, sb, (const RowInfo*)row_info, row_offset, H * n_seg, out, out_stride);
}

}  // namespace slbk
