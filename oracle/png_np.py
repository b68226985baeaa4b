"""ORACLE (test infrastructure only: imported by tests/ — never by the product).

Byte-for-byte CPU restatement of the batched PNG encoder (stillleben_b200/csrc/k_png.cu), which replaces the
reference's thread-pool image saver (src/image_saver.cpp, python/src/py_image_saver.cpp:37-99 — libpng behind
Magnum's AnyImageConverter). The reference's exact bytes are libpng's and are not reproducible without it; what
IS defined is the decoded image: row 0 of the tensor is the top row, uint8 HxW / HxWx3 / HxWx4 or 16-bit HxW.
So parity = (a) this restatement decodes, with an independent decoder (PIL, zlib), to exactly the input pixels,
and (b) the kernels produce exactly these bytes. Stream layout: PNG filter 1 (Sub) on every scanline; one
fixed-Huffman deflate block per 512-byte scanline segment with run matches (length >= 4) at distance 1 or bytes-per-pixel,
followed by an empty stored block (byte alignment); a final empty stored block; zlib header 78 01.
Pure-Python loops: small images only.
"""
import struct
import zlib

import numpy as np

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


class BitWriter:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, v, n):
        self.acc |= v << self.n
        self.n += n
        while self.n >= 8:
            self.out.append(self.acc & 0xFF)
            self.acc >>= 8
            self.n -= 8

    def align(self):
        if self.n:
            self.out.append(self.acc & 0xFF)
            self.acc, self.n = 0, 0


def rev(v, n):
    return int(format(v, f"0{n}b")[::-1], 2)


def put_literal(w, lit):
    if lit < 144:
        w.put(rev(0x30 + lit, 8), 8)
    else:
        w.put(rev(0x190 + lit - 144, 9), 9)


def put_match(w, length, dist):
    c = 0
    while c < 28 and LEN_BASE[c + 1] <= length:
        c += 1
    sym = 257 + c
    if sym < 280:
        w.put(rev(sym - 256, 7), 7)
    else:
        w.put(rev(0xC0 + sym - 280, 8), 8)
    if LEN_EXTRA[c]:
        w.put(length - LEN_BASE[c], LEN_EXTRA[c])
    w.put(rev(dist - 1, 5), 5)


SEG = 512       # data bytes per deflate block (k_png.cu: PNG_SEG)


def encode_row(raw, bpp):
    """raw: the scanline's bytes as PNG stores them. Returns the scanline's byte-aligned deflate blocks (one per
    SEG-byte segment, no match reaches back across a segment start) and its filtered bytes."""
    n = len(raw)
    f = [(raw[i] - (raw[i - bpp] if i >= bpp else 0)) & 0xFF for i in range(n)]
    out = bytearray()
    for beg in range(0, max(n, 1), SEG):
        end = min(beg + SEG, n)
        w = BitWriter()
        w.put(0, 1); w.put(1, 2)
        if beg == 0:
            put_literal(w, 1)
        i = beg
        while i < end:
            best, dist = 0, 0
            for d in ([1] if bpp == 1 else [1, bpp]):
                if i < beg + d:
                    continue
                l = 0
                while l < 258 and i + l < end and f[i + l] == f[i + l - d]:
                    l += 1
                if l > best:
                    best, dist = l, d
            if best >= 4:
                put_match(w, best, dist)
                i += best
            else:
                put_literal(w, f[i])
                i += 1
        w.put(0, 7)
        w.put(0, 1); w.put(0, 2); w.align()
        out += w.out + b"\x00\x00\xff\xff"
    return bytes(out), bytes([1] + f)


def chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def encode(img):
    """img: uint8 HxW / HxWx3 / HxWx4 or uint16 / int16 HxW -> PNG file bytes."""
    img = np.asarray(img)
    if img.dtype in (np.uint16, np.int16):
        assert img.ndim == 2
        bpc, channels = 2, 1
        rows = img.astype(">u2" if img.dtype == np.uint16 else ">i2").view(np.uint8).reshape(img.shape[0], -1)
    else:
        assert img.dtype == np.uint8
        bpc, channels = 1, (1 if img.ndim == 2 else img.shape[2])
        rows = img.reshape(img.shape[0], -1)
    H, W = img.shape[:2]
    bpp = bpc * channels
    body, filtered = bytearray(b"\x78\x01"), bytearray()
    for r in range(H):
        blk, filt = encode_row(rows[r].tolist(), bpp)
        body += blk
        filtered += filt
    body += b"\x01\x00\x00\xff\xff" + struct.pack(">I", zlib.adler32(bytes(filtered)) & 0xFFFFFFFF)
    ihdr = struct.pack(">IIBBBBB", W, H, 8 * bpc, {1: 0, 3: 2, 4: 6}[channels], 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", bytes(body)) + chunk(b"IEND", b"")
