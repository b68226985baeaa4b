"""ORACLE (test infrastructure only: imported by tests/ — never by the product).

CPU restatement of the baseline JPEG writer behind the reference's image saver: src/image_saver.cpp:55-97 hands every
image to Magnum's AnyImageConverter, which for *.jpg / *.jpeg runs JpegImageConverter = libjpeg with
jpeg_set_defaults() + jpeg_set_quality(80, force_baseline) (Magnum's default jpegQuality 0.8). libjpeg is a third-party
dependency absent from /root/reference; its published algorithm (IJG jcparam.c / jccolor.c / jcsample.c / jfdctint.c /
jcdctmgr.c / jchuff.c / jcmarker.c, release 6b lineage as in libjpeg-turbo) is restated here:
  * quantisation tables: Annex K tables scaled by 200 - 2q, (base * s + 50) / 100 clamped to 1..255;
  * RGB -> YCbCr in 16-bit fixed point (FIX(0.299) = 19595 ...), chroma 2x2 box-averaged with the alternating 1,2 bias
    (h2v2_downsample), edges replicated to whole MCUs; grey images: one component;
  * forward DCT = the "slow integer" LL&M transform (CONST_BITS 13, PASS1_BITS 2), quantised by (|c| + q/2) / q with q = 8 Q;
  * Annex K Huffman tables, MSB-first bit stream, 0xFF byte stuffing, padding with 1-bits;
  * marker order SOI APP0(JFIF 1.01, density 1:1) DQT.. SOF0 DHT.. SOS <scan> EOI.
PINNED: tests/test_jpeg_oracle.py checks that these are EXACTLY the bytes libjpeg itself writes for the same pixels
(PIL's encoder is libjpeg-turbo: Image.save(quality=80, subsampling="4:2:0")) — so the restatement is tied to the library
the reference calls, not only to a decoder's tolerance. The CUDA encoder (stillleben_b200/csrc/k_jpeg.cu) must reproduce
these bytes. Vectorised numpy; any image size.
"""
import numpy as np

STD_LUMA_Q = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                       18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100,
                       103, 99], np.int64)
STD_CHROMA_Q = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99]
                        + [99] * 32, np.int64)
ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56,
                   57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d],
           [0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
            0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
            0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
            0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
            0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
            0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
            0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
            0xfa])
AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77],
             [0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
              0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
              0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
              0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
              0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
              0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
              0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
              0xfa])


def quant_table(base, quality=80):
    """jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)."""
    quality = min(max(int(quality), 1), 100)
    s = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((base * s + 50) // 100, 1, 255)


def derive_codes(bits, vals):
    """jchuff.c:jpeg_make_c_derived_tbl — canonical codes: (code, length) per symbol."""
    code, size, k = {}, {}, 0
    c = 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            code[vals[k]], size[vals[k]] = c, length
            c += 1
            k += 1
        c <<= 1
    return code, size


def huff_lut(table):
    """(code << 8 | length) per symbol, 0 for unused symbols — the 256-entry table form the kernels use."""
    code, size = derive_codes(*table)
    lut = np.zeros(256, np.uint32)
    for s in code:
        lut[s] = (code[s] << 8) | size[s]
    return lut


def fdct_islow(block):
    """jfdctint.c (6b lineage) on [..., 8, 8] int64 samples already level-shifted by -128; output scaled by 8."""
    C = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069, f2053=16819,
             f2562=20995, f3072=25172)

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    def pass_(d, first):
        t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
        t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
        t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
        t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
        t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
        o = [None] * 8
        if first:
            o[0], o[4] = (t10 + t11) << 2, (t10 - t11) << 2
            n = 13 - 2
        else:
            o[0], o[4] = descale(t10 + t11, 2), descale(t10 - t11, 2)
            n = 13 + 2
        z1 = (t12 + t13) * C["f0541"]
        o[2] = descale(z1 + t13 * C["f0765"], n)
        o[6] = descale(z1 + t12 * (-C["f1847"]), n)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * C["f1175"]
        t4, t5, t6, t7 = t4 * C["f0298"], t5 * C["f2053"], t6 * C["f3072"], t7 * C["f1501"]
        z1, z2, z3, z4 = z1 * (-C["f0899"]), z2 * (-C["f2562"]), z3 * (-C["f1961"]) + z5, z4 * (-C["f0390"]) + z5
        o[7], o[5], o[3], o[1] = descale(t4 + z1 + z3, n), descale(t5 + z2 + z4, n), descale(t6 + z2 + z3, n), descale(t7 + z1 + z4, n)
        return np.stack(o, -1)

    rows = pass_(block.astype(np.int64), True)                       # pass 1: along each row
    cols = pass_(np.swapaxes(rows, -1, -2), False)                   # pass 2: along each column
    return np.swapaxes(cols, -1, -2)


def quantize(coef, q):
    """jcdctmgr.c:forward_DCT — divisor 8 Q, round half away from zero by (|c| + d/2) / d."""
    d = (q.reshape(8, 8) << 3)
    mag = (np.abs(coef) + (d >> 1)) // d
    return np.where(coef < 0, -mag, mag)


def _blocks(plane, bw, bh):
    """plane [H, W] -> [bh, bw, 8, 8] with edge replication up to the block grid (jcprepct.c / jcsample.c edge expansion)."""
    H, W = plane.shape
    ys = np.minimum(np.arange(bh * 8), H - 1)
    xs = np.minimum(np.arange(bw * 8), W - 1)
    p = plane[ys][:, xs]
    return p.reshape(bh, 8, bw, 8).transpose(0, 2, 1, 3)


def components(img):
    """uint8 [H,W] or [H,W,3|4] (row 0 = top) -> list of (blocks [bh,bw,8,8] int64 level-shifted, quant id), MCU geometry."""
    if img.ndim == 2:
        H, W = img.shape
        bw, bh = (W + 7) // 8, (H + 7) // 8
        return [(_blocks(img.astype(np.int64), bw, bh) - 128, 0)], (bw, bh, 1)
    H, W = img.shape[:2]
    r, g, b = (img[..., k].astype(np.int64) for k in range(3))
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16
    mw, mh = (W + 15) // 16, (H + 15) // 16
    out = [(_blocks(y, 2 * mw, 2 * mh) - 128, 0)]
    # jcprepct.c / jcsample.c: the colour rows are replicated only up to an EVEN height and, per row, to the right up to the
    # chroma block grid at full resolution; the DOWNSAMPLED rows are then replicated down to the iMCU grid
    ch = (H + 1) // 2
    ys = np.minimum(np.arange(2 * ch), H - 1)
    xs = np.minimum(np.arange(mw * 16), W - 1)
    bias = np.tile(np.array([1, 2]), mw * 4)[None, :]                  # h2v2_downsample: bias 1, 2, 1, 2 along a row
    rows = np.minimum(np.arange(mh * 8), ch - 1)
    for c in (cb, cr):
        p = c[ys][:, xs]
        ds = ((p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + bias) >> 2)[rows]
        out.append((ds.reshape(mh, 8, mw, 8).transpose(0, 2, 1, 3) - 128, 1))
    return out, (mw, mh, 3)


def coefficients(img, quality=80):
    """Quantised coefficients in ZIGZAG order, in scan (MCU-interleaved) block order: int64 [n_blocks, 64], component id per block."""
    comps, (mw, mh, nc) = components(img)
    qt = [quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)]
    qc = [quantize(fdct_islow(blk), qt[qid]).reshape(blk.shape[0], blk.shape[1], 64)[..., ZIGZAG] for blk, qid in comps]
    if nc == 1:
        return qc[0].reshape(-1, 64), np.zeros(mw * mh, np.int64), qt
    Y = qc[0].reshape(mh, 2, mw, 2, 64).transpose(0, 2, 1, 3, 4).reshape(mh, mw, 4, 64).copy()  # Y00 Y01 Y10 Y11 per MCU
    # jccoefct.c:compress_data — luma blocks of the last MCU column / row that lie wholly outside the image are "dummy blocks":
    # AC = 0, DC = that of the previous block of the MCU (right edge: the block to the left; bottom edge: the MCU's Y01)
    H, W = img.shape[:2]
    wb, hb = (W + 7) // 8, (H + 7) // 8
    if wb & 1:
        for yi in (0, 1):
            Y[:, mw - 1, 2 * yi + 1, 1:] = 0
            Y[:, mw - 1, 2 * yi + 1, 0] = Y[:, mw - 1, 2 * yi, 0]
    if hb & 1:
        for bi in (0, 1):
            Y[mh - 1, :, 2 + bi, 1:] = 0
            Y[mh - 1, :, 2 + bi, 0] = Y[mh - 1, :, 1, 0]
    mcu = np.concatenate([Y, qc[1][:, :, None, :], qc[2][:, :, None, :]], axis=2)               # + Cb + Cr
    comp = np.tile(np.array([0, 0, 0, 0, 1, 2]), mw * mh)
    return mcu.reshape(-1, 64), comp, qt


def _nbits(v):
    return int(abs(int(v))).bit_length()


def scan_bits(coefs, comp):
    """jchuff.c:encode_one_block over every block in scan order -> list of (value, n_bits) tokens, MSB-first."""
    dc_tabs = [derive_codes(*DC_LUMA), derive_codes(*DC_CHROMA)]
    ac_tabs = [derive_codes(*AC_LUMA), derive_codes(*AC_CHROMA)]
    last = {0: 0, 1: 0, 2: 0}
    toks = []
    for blk, c in zip(coefs, comp):
        t = 0 if c == 0 else 1
        dcode, dsize = dc_tabs[t]
        acode, asize = ac_tabs[t]
        diff = int(blk[0]) - last[int(c)]
        last[int(c)] = int(blk[0])
        n = _nbits(diff)
        toks.append((dcode[n], dsize[n]))
        if n:
            toks.append(((diff if diff >= 0 else diff - 1) & ((1 << n) - 1), n))
        run = 0
        for k in range(1, 64):
            v = int(blk[k])
            if v == 0:
                run += 1
                continue
            while run > 15:
                toks.append((acode[0xF0], asize[0xF0]))
                run -= 16
            n = _nbits(v)
            toks.append((acode[(run << 4) + n], asize[(run << 4) + n]))
            toks.append(((v if v >= 0 else v - 1) & ((1 << n) - 1), n))
            run = 0
        if run:
            toks.append((acode[0], asize[0]))
    return toks


def pack_scan(toks):
    acc, n = 0, 0
    for v, k in toks:
        acc = (acc << k) | v
        n += k
    pad = (-n) % 8
    acc = (acc << pad) | ((1 << pad) - 1)                               # flush_bits: fill with ones
    raw = acc.to_bytes((n + pad) // 8, "big") if n + pad else b""
    return raw.replace(b"\xff", b"\xff\x00")


def _marker(tag, payload):
    return bytes([0xFF, tag]) + (len(payload) + 2).to_bytes(2, "big") + payload


def header(W, H, ncomp, qt):
    """jcmarker.c: write_file_header + write_frame_header + write_scan_header for a baseline image."""
    out = b"\xff\xd8" + _marker(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    for i in range(1 if ncomp == 1 else 2):
        out += _marker(0xDB, bytes([i]) + bytes(int(v) for v in qt[i][ZIGZAG]))
    if ncomp == 1:
        sof = bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([1, 1, 0x11, 0])
    else:
        sof = bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1])
    out += _marker(0xC0, sof)
    tabs = [(0x00, DC_LUMA), (0x10, AC_LUMA)] + ([] if ncomp == 1 else [(0x01, DC_CHROMA), (0x11, AC_CHROMA)])
    for tid, (bits, vals) in tabs:
        out += _marker(0xC4, bytes([tid]) + bytes(bits) + bytes(vals))
    if ncomp == 1:
        out += _marker(0xDA, bytes([1, 1, 0x00, 0, 63, 0]))
    else:
        out += _marker(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def encode(img, quality=80):
    """uint8 [H,W] / [H,W,3] / [H,W,4] (alpha ignored), row 0 = top -> the complete JFIF file as bytes."""
    img = np.asarray(img)
    assert img.dtype == np.uint8
    coefs, comp, qt = coefficients(img, quality)
    H, W = img.shape[:2]
    return header(W, H, 1 if img.ndim == 2 else 3, qt) + pack_scan(scan_bits(coefs, comp)) + b"\xff\xd9"
