"""Reader of the dump oracle/_ref/meshtool writes (oracle/ref_meshtool.cpp) — test infrastructure only."""
import struct

import numpy as np

# Magnum enums (contrib/magnum/src/Magnum/Sampler.h): SamplerFilter {Nearest, Linear}, SamplerMipmap {Base, Nearest, Linear},
# SamplerWrapping {Repeat, MirroredRepeat, ClampToEdge, ClampToBorder, MirrorClampToEdge}
FILTER = ["nearest", "linear"]
MIPMAP = ["base", "nearest", "linear"]
WRAP = ["repeat", "mirrored_repeat", "clamp_to_edge", "clamp_to_border", "mirror_clamp_to_edge"]


def read(path):
    b = open(path, "rb").read()
    at = 0

    def take(fmt):
        nonlocal at
        v = struct.unpack_from("<" + fmt, b, at)
        at += struct.calcsize("<" + fmt)
        return v

    magic, stride, nv, ni = take("IIII")
    assert magic == 0x534C4D31
    out = {"vertices": np.frombuffer(b, np.uint8, nv * stride, at).reshape(nv, stride).copy()}
    at += nv * stride
    out["indices"] = np.frombuffer(b, np.uint32, ni, at).copy()
    at += ni * 4
    (nsub,) = take("I")
    out["submeshes"] = np.frombuffer(b, np.int32, nsub * 3, at).reshape(nsub, 3).astype(np.int64)
    at += nsub * 12
    (nmat,) = take("I")
    mats = []
    for _ in range(nmat):
        (ok,) = take("I")
        if not ok:
            mats.append([np.nan] * 16)
            continue
        vals = take("10f")
        tex = take("5i")
        (tm,) = take("f")
        mats.append(list(vals) + list(tex) + [tm])
    out["materials"] = np.array(mats, np.float64).reshape(nmat, 16)
    (ntex,) = take("I")
    out["textures"] = np.array([take("iIIIII") for _ in range(ntex)], np.int64).reshape(ntex, 6)
    (nimg,) = take("I")
    for i in range(nimg):
        w, h, ch = take("III")
        if ch:
            out[f"image{i}"] = np.frombuffer(b, np.uint8, w * h * ch, at).reshape(h, w, ch).copy()
            at += w * h * ch
    assert at == len(b)
    return out
