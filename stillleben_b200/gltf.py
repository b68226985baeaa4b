"""Minimal glTF 2.0 / GLB front end producing the consolidated vertex stream the hot path consumes.

This is the caller side of the path, not the path itself: it restates what the
reference does between `importer.openFile()` and `Mesh::loadVisual()`:
  * one sub-mesh per (node, primitive), node transforms baked into positions / normals / tangents,
    32-bit indices rebased per sub-mesh, one-based vertex ids, tangent.w forced to 1
    (reference: src/mesh_tools/consolidate.cpp:53-61,212-335),
  * smooth normals when missing (consolidate.cpp:78-88), per-vertex tangents from UV deltas when
    missing, zero without UVs (src/mesh_tools/compute_tangents.cpp:53-110),
  * import conventions of the vendored importers: v <- 1 - v, image rows bottom-up,
  * material defaulting of RenderShader::setMaterial incl. the CgltfImporter quirk that factors equal
    to the glTF default 1.0 are not emitted (src/shaders/render_shader.cpp:355-383).
Only what the reference's two test assets and typical YCB-style exports use is supported: float
POSITION/NORMAL/TANGENT/TEXCOORD_0, u8/u16/u32 indices, PNG/JPEG images (through PIL), embedded or
external buffers.
"""
import base64
import json
import os
import struct

import numpy as np

from . import abi
from .desc import ImageData, MaterialData, MeshData

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}
_WRAP = {10497: abi.WRAP_REPEAT, 33071: abi.WRAP_CLAMP_TO_EDGE, 33648: abi.WRAP_MIRRORED_REPEAT}
_FILTER = {9728: abi.FILTER_NEAREST, 9729: abi.FILTER_LINEAR, 9984: abi.FILTER_NEAREST_MIPMAP_NEAREST,
           9985: abi.FILTER_LINEAR_MIPMAP_NEAREST, 9986: abi.FILTER_NEAREST_MIPMAP_LINEAR, 9987: abi.FILTER_LINEAR_MIPMAP_LINEAR}


def _quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def _node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], np.float64).reshape(4, 4).T     # glTF matrices are column-major
    m = np.eye(4)
    if "scale" in node:
        m[:3, :3] = np.diag(node["scale"])
    if "rotation" in node:
        m[:3, :3] = _quat_to_mat(node["rotation"]) @ m[:3, :3]
    if "translation" in node:
        m[:3, 3] = node["translation"]
    return m


class _Doc:
    def __init__(self, path):
        self.dir = os.path.dirname(os.path.abspath(path))
        raw = open(path, "rb").read()
        self.bin_chunk = None
        if raw[:4] == b"glTF":
            _, _, length = struct.unpack("<4sII", raw[:12])
            at = 12
            while at < length:
                clen, ctype = struct.unpack("<I4s", raw[at:at + 8])
                data = raw[at + 8:at + 8 + clen]
                if ctype == b"JSON":
                    self.js = json.loads(data)
                elif ctype == b"BIN\x00":
                    self.bin_chunk = data
                at += 8 + clen
        else:
            self.js = json.loads(raw)
        self.buffers = []
        for b in self.js.get("buffers", []):
            uri = b.get("uri")
            if uri is None:
                self.buffers.append(self.bin_chunk)
            elif uri.startswith("data:"):
                self.buffers.append(base64.b64decode(uri.split(",", 1)[1]))
            else:
                self.buffers.append(open(os.path.join(self.dir, uri), "rb").read())

    def accessor(self, i):
        a = self.js["accessors"][i]
        bv = self.js["bufferViews"][a["bufferView"]]
        dt, nc = _COMP[a["componentType"]], _NCOMP[a["type"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or np.dtype(dt).itemsize * nc
        buf = self.buffers[bv["buffer"]]
        out = np.empty((a["count"], nc), dt)
        item = np.dtype(dt).itemsize * nc
        if stride == item:
            out[:] = np.frombuffer(buf, dt, a["count"] * nc, off).reshape(a["count"], nc)
        else:
            for k in range(a["count"]):
                out[k] = np.frombuffer(buf, dt, nc, off + k * stride)
        return out

    def image(self, i):
        from PIL import Image
        import io
        im = self.js["images"][i]
        if "uri" in im:
            if im["uri"].startswith("data:"):
                pil = Image.open(io.BytesIO(base64.b64decode(im["uri"].split(",", 1)[1])))
            else:
                pil = Image.open(os.path.join(self.dir, im["uri"]))
        else:
            bv = self.js["bufferViews"][im["bufferView"]]
            off = bv.get("byteOffset", 0)
            pil = Image.open(io.BytesIO(self.buffers[bv["buffer"]][off:off + bv["byteLength"]]))
        pil = pil.convert("RGBA" if pil.mode in ("RGBA", "LA", "P") and "A" in pil.getbands() else "RGB")
        return np.ascontiguousarray(np.asarray(pil)[::-1])     # StbImageImporter: row 0 = bottom row


def _smooth_normals(pos, idx):
    tri = idx.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    n = np.zeros_like(pos)
    for k in range(3):
        np.add.at(n, tri[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    return (n / np.maximum(ln, 1e-30)).astype(np.float32)


def _compute_tangents(pos, uv, idx):
    """compute_tangents.cpp:53-110: per-face tangent from UV deltas, averaged per vertex, normalised."""
    tri = idx.reshape(-1, 3)
    d1, d2 = pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]]
    u1, u2 = uv[tri[:, 1]] - uv[tri[:, 0]], uv[tri[:, 2]] - uv[tri[:, 0]]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = 1.0 / (u1[:, 0] * u2[:, 1] - u1[:, 1] * u2[:, 0])
        t = (d1 * u2[:, 1:2] - d2 * u1[:, 1:2]) * r[:, None]
    acc = np.zeros_like(pos)
    deg = np.zeros(len(pos))
    for k in range(3):
        np.add.at(acc, tri[:, k], t)
        np.add.at(deg, tri[:, k], 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = acc / deg[:, None]
        acc = acc / np.linalg.norm(acc, axis=1, keepdims=True)
    return acc.astype(np.float32)


def load(path, name=None):
    """Load a .gltf / .glb file into a MeshData (consolidated, reference vertex layout)."""
    doc = _Doc(path)
    js = doc.js
    verts, inds, subs = [], [], []
    v_off = i_off = 0

    def visit(ni, parent):
        nonlocal v_off, i_off
        node = js["nodes"][ni]
        m = parent @ _node_matrix(node)
        if "mesh" in node:
            for prim in js["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4 or "indices" not in prim:
                    continue        # consolidate.cpp:72-76: non-triangle / non-indexed sub-meshes are ignored
                at = prim["attributes"]
                pos = doc.accessor(at["POSITION"]).astype(np.float32)
                idx = doc.accessor(prim["indices"]).astype(np.uint32).reshape(-1)
                uv = doc.accessor(at["TEXCOORD_0"]).astype(np.float32) if "TEXCOORD_0" in at else None
                if uv is not None:
                    uv = uv.copy()
                    uv[:, 1] = 1.0 - uv[:, 1]
                nrm = doc.accessor(at["NORMAL"]).astype(np.float32) if "NORMAL" in at else _smooth_normals(pos, idx)
                if "TANGENT" in at:
                    tan = doc.accessor(at["TANGENT"]).astype(np.float32)[:, :3]
                elif uv is not None:
                    tan = _compute_tangents(pos, uv, idx)
                else:
                    tan = np.zeros_like(pos)
                v = np.zeros(len(pos), abi.VERTEX_DTYPE)
                R, t = m[:3, :3], m[:3, 3]
                v["position"] = (pos.astype(np.float64) @ R.T + t).astype(np.float32)
                v["normal"] = (nrm.astype(np.float64) @ R.T).astype(np.float32)          # transformVector, not renormalised
                v["tangent"][:, :3] = (tan.astype(np.float64) @ R.T).astype(np.float32)
                v["tangent"][:, 3] = 1.0                                                 # consolidate.cpp:278
                v["uv"] = uv if uv is not None else 0.0
                v["color"] = 1.0
                verts.append(v)
                inds.append(idx + v_off)
                subs.append((i_off, len(idx), prim.get("material", -1)))
                v_off += len(pos)
                i_off += len(idx)
        for c in node.get("children", []):
            visit(c, m)

    scene = js["scenes"][js.get("scene", 0)]
    for ni in scene["nodes"]:
        visit(ni, np.eye(4))
    if not verts:
        raise ValueError(f"{path}: no triangle meshes")      # reference: Mesh::LoadException (src/mesh.cpp:244-248)
    vertices = np.concatenate(verts)
    vertices["vertex_index"] = np.arange(1, len(vertices) + 1, dtype=np.uint32)
    indices = np.concatenate(inds)

    # textures -> images with their sampler state; only RGB8 / RGBA8 are accepted (mesh.cpp:644-653)
    images, tex_to_image = [], {}
    for ti, tex in enumerate(js.get("textures", [])):
        smp = js.get("samplers", [{}])[tex["sampler"]] if "sampler" in tex else {}
        px = doc.image(tex["source"])
        tex_to_image[ti] = len(images)
        images.append(ImageData(px, _WRAP.get(smp.get("wrapS", 10497), abi.WRAP_REPEAT), _WRAP.get(smp.get("wrapT", 10497), abi.WRAP_REPEAT),
                                _FILTER.get(smp.get("minFilter", 9987), abi.FILTER_LINEAR_MIPMAP_LINEAR),
                                _FILTER.get(smp.get("magFilter", 9729), abi.FILTER_LINEAR)))

    def tex_index(info):
        return tex_to_image.get(info["index"], -1) if info else -1

    materials = []
    for mat in js.get("materials", []):
        pbr = mat.get("pbrMetallicRoughness", {})
        mr_tex = tex_index(pbr.get("metallicRoughnessTexture"))
        metallic, roughness = (1.0, 1.0) if mr_tex >= 0 else (0.04, 0.5)        # render_shader.cpp:355-363
        if pbr.get("metallicFactor", 1.0) != 1.0:                              # importer omits the glTF default 1.0
            metallic = float(pbr["metallicFactor"])
        if pbr.get("roughnessFactor", 1.0) != 1.0:
            roughness = float(pbr["roughnessFactor"])
        em = list(mat.get("emissiveFactor", [0.0, 0.0, 0.0]))
        materials.append(MaterialData(tuple(float(x) for x in pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])),
                                      (float(em[0]), float(em[1]), float(em[2]), 0.0), metallic, roughness,
                                      tex_index(pbr.get("baseColorTexture")), tex_index(mat.get("normalTexture")), mr_tex,
                                      tex_index(mat.get("emissiveTexture")), tex_index(mat.get("occlusionTexture"))))
    subs = [(o, c, m if 0 <= m < len(materials) else -1) for o, c, m in subs]
    return MeshData(vertices, indices, subs, materials, images, name=name or os.path.basename(path))
