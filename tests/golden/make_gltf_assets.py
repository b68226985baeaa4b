"""Writes the small glTF assets under tests/golden/assets/ (our own files, not the reference's) that exercise the mesh
front end beyond the reference's two test assets: missing normals / tangents / UVs, several primitives per mesh, nested
TRS nodes, float vertex colours, u8 / u16 / u32 indices, several materials with all five texture kinds and non-default
samplers, factors equal to the glTF default (the importer quirk of SURVEY A.10), a non-mesh node.
tests/golden/make_ref_golden.py then runs the REFERENCE's consolidation on them (oracle/_ref/meshtool).

    python tests/golden/make_gltf_assets.py
"""
import io
import json
import os
import struct

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "assets")


class Builder:
    def __init__(self):
        self.bin = bytearray()
        self.js = {"asset": {"version": "2.0"}, "buffers": [{}], "bufferViews": [], "accessors": [], "meshes": [], "nodes": [],
                   "scenes": [{"nodes": []}], "scene": 0, "materials": [], "textures": [], "images": [], "samplers": []}

    def view(self, data):
        while len(self.bin) % 4:
            self.bin += b"\0"
        self.js["bufferViews"].append({"buffer": 0, "byteOffset": len(self.bin), "byteLength": len(data)})
        self.bin += data
        return len(self.js["bufferViews"]) - 1

    def accessor(self, arr, kind):
        ct = {np.dtype(np.float32): 5126, np.dtype(np.uint32): 5125, np.dtype(np.uint16): 5123, np.dtype(np.uint8): 5121}[arr.dtype]
        a = {"bufferView": self.view(arr.tobytes()), "componentType": ct, "count": len(arr), "type": kind}
        if kind == "VEC3" and arr.dtype == np.float32:
            a["min"], a["max"] = arr.min(0).tolist(), arr.max(0).tolist()
        self.js["accessors"].append(a)
        return len(self.js["accessors"]) - 1

    def image(self, px):
        buf = io.BytesIO()
        Image.fromarray(px).save(buf, format="PNG")
        self.js["images"].append({"bufferView": self.view(buf.getvalue()), "mimeType": "image/png"})
        return len(self.js["images"]) - 1

    def write(self, name):
        self.js["buffers"][0]["byteLength"] = len(self.bin)
        js = {k: v for k, v in self.js.items() if v != []}
        jb = json.dumps(js).encode()
        jb += b" " * (-len(jb) % 4)
        bb = bytes(self.bin) + b"\0" * (-len(self.bin) % 4)
        total = 12 + 8 + len(jb) + 8 + len(bb)
        with open(os.path.join(OUT, name), "wb") as f:
            f.write(struct.pack("<4sII", b"glTF", 2, total))
            f.write(struct.pack("<I4s", len(jb), b"JSON") + jb)
            f.write(struct.pack("<I4s", len(bb), b"BIN\0") + bb)


def grid(nu, nv, seed):
    """A bumpy open patch: positions, uv, indices."""
    rng = np.random.RandomState(seed)
    u, v = np.meshgrid(np.linspace(0, 1, nu), np.linspace(0, 1, nv), indexing="xy")
    z = 0.15 * np.sin(5 * u + seed) * np.cos(4 * v) + 0.01 * rng.rand(nv, nu)
    pos = np.stack([u - 0.5, v - 0.5, z], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([u * 1.7 + 0.1 * v, v * 1.3], -1).reshape(-1, 2).astype(np.float32)
    idx = []
    for j in range(nv - 1):
        for i in range(nu - 1):
            a = j * nu + i
            idx += [a, a + 1, a + nu, a + 1, a + nu + 1, a + nu]
    return pos, uv, np.array(idx)


def smooth(pos, idx):
    tri = idx.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    n = np.zeros_like(pos)
    for k in range(3):
        np.add.at(n, tri[:, k], fn)
    return (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)


def tex(seed, n, channels=3):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:n, 0:n]
    px = np.stack([(127 + 120 * np.sin(x / (2.0 + k) + seed) * np.cos(y / (3.0 + k))) for k in range(channels)], -1)
    px += rng.randint(-6, 7, size=px.shape)
    px = np.clip(px, 0, 255).astype(np.uint8)
    if channels == 4:
        px[..., 3] = np.where(((x // 4) + (y // 4)) % 2 == 0, 255, 40)
    return px


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- asset 1: everything the front end has to generate or carry ----
    b = Builder()
    b.js["samplers"] = [{"magFilter": 9728, "minFilter": 9985, "wrapS": 33071, "wrapT": 33648}, {}]
    for k, ch in enumerate((3, 3, 3, 3, 4, 3)):
        b.image(tex(10 + k, 16 if k else 32, ch))
    b.js["textures"] = [{"sampler": 0, "source": 0}, {"sampler": 1, "source": 1}, {"source": 2}, {"sampler": 1, "source": 3},
                        {"sampler": 1, "source": 4}, {"sampler": 0, "source": 5}]
    b.js["materials"] = [
        {"pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.8, 0.7, 1.0], "baseColorTexture": {"index": 0},
                                  "metallicRoughnessTexture": {"index": 2}, "metallicFactor": 0.5, "roughnessFactor": 0.8},
         "normalTexture": {"index": 1}, "occlusionTexture": {"index": 3}, "emissiveTexture": {"index": 5},
         "emissiveFactor": [0.3, 0.2, 0.1]},
        {"pbrMetallicRoughness": {"metallicFactor": 1.0, "roughnessFactor": 1.0}},            # glTF defaults: not emitted by the importer
        {"pbrMetallicRoughness": {"baseColorTexture": {"index": 4}, "metallicRoughnessTexture": {"index": 2}}},
        {"pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.4, 0.6, 0.3], "metallicFactor": 0.0, "roughnessFactor": 0.25}},
    ]
    prims = []
    # (a) full attribute set, u16 indices
    pos, uv, idx = grid(7, 5, 1)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (len(pos), 1))
    tan = np.tile(np.array([[1, 0, 0, -1]], np.float32), (len(pos), 1))
    col = np.random.RandomState(3).rand(len(pos), 4).astype(np.float32)
    prims.append({"attributes": {"POSITION": b.accessor(pos, "VEC3"), "NORMAL": b.accessor(nrm, "VEC3"), "TANGENT": b.accessor(tan, "VEC4"),
                                 "TEXCOORD_0": b.accessor(uv, "VEC2"), "COLOR_0": b.accessor(col, "VEC4")},
                  "indices": b.accessor(idx.astype(np.uint16), "SCALAR"), "material": 0})
    # (b) no tangents, UVs present (computeTangents), u32 indices. (Normals must be present: the reference ABORTS on a
    # mesh without them — consolidate.cpp:84-87 passes the vertex count as the attribute's array size and Magnum asserts
    # "Normal can't be an array attribute".)
    pos, uv, idx = grid(6, 6, 2)
    prims.append({"attributes": {"POSITION": b.accessor(pos + np.float32(0.2), "VEC3"), "NORMAL": b.accessor(smooth(pos, idx), "VEC3"),
                                 "TEXCOORD_0": b.accessor(uv, "VEC2")},
                  "indices": b.accessor(idx.astype(np.uint32), "SCALAR"), "material": 2})
    # (c) positions + normals only (tangents stay zero, uv zero), u8 indices, no material
    pos, uv, idx = grid(4, 4, 3)
    prims.append({"attributes": {"POSITION": b.accessor(pos, "VEC3"), "NORMAL": b.accessor(smooth(pos, idx), "VEC3")},
                  "indices": b.accessor(idx.astype(np.uint8), "SCALAR")})
    b.js["meshes"] = [{"primitives": [prims[0], prims[1]]}, {"primitives": [prims[2]]}, {"primitives": [dict(prims[0], material=3)]}]
    b.js["nodes"] = [
        {"children": [1, 2], "translation": [0.1, -0.2, 0.3], "rotation": [0.18257419, 0.36514837, 0.54772256, 0.73029674], "scale": [1.5, 0.5, 2.0]},
        {"mesh": 0, "rotation": [0.0, 0.70710678, 0.0, 0.70710678]},
        {"children": [3], "matrix": [0.0, 1.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.5, 0.25, -0.125, 1.0]},
        {"mesh": 1, "scale": [2.0, 2.0, 2.0], "children": [4]},
        {"name": "empty"},
        {"mesh": 2, "translation": [1.0, 2.0, 3.0]},
    ]
    b.js["scenes"][0]["nodes"] = [0, 5]
    b.write("kitchen_sink.glb")
    # ---- asset 2: a single textured patch with a normal map etc. (used by the renderer parity variants too) ----
    b = Builder()
    b.js["samplers"] = [{}]
    for k in range(5):
        b.image(tex(20 + k, 32, 3))
    b.js["textures"] = [{"sampler": 0, "source": k} for k in range(5)]
    b.js["materials"] = [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 2}},
                          "normalTexture": {"index": 1}, "occlusionTexture": {"index": 4}, "emissiveTexture": {"index": 3},
                          "emissiveFactor": [1.0, 1.0, 1.0]}]
    pos, uv, idx = grid(9, 9, 5)
    b.js["meshes"] = [{"primitives": [{"attributes": {"POSITION": b.accessor(pos, "VEC3"), "NORMAL": b.accessor(smooth(pos, idx), "VEC3"),
                                                      "TEXCOORD_0": b.accessor(uv, "VEC2")},
                                       "indices": b.accessor(idx.astype(np.uint16), "SCALAR"), "material": 0}]}]
    b.js["nodes"] = [{"mesh": 0}]
    b.js["scenes"][0]["nodes"] = [0]
    b.write("pbr_patch.glb")
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
