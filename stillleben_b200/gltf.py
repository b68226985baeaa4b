"""Minimal glTF 2.0 / GLB front end producing the consolidated vertex stream the hot path consumes.

This is the caller side of the path, not the path itself: it restates what the
reference does between `importer.openFile()` and `Mesh::loadVisual()`:
  * one sub-mesh per (node, primitive), node transforms baked into positions / normals / tangents,
    32-bit indices rebased per sub-mesh, one-based vertex ids, tangent.w forced to 1
    (reference: src/mesh_tools/consolidate.cpp:53-61,212-335),
  * per-vertex tangents from UV deltas when missing, zero without UVs (src/mesh_tools/compute_tangents.cpp:53-110);
    missing normals are an error like in the reference (its smooth-normal branch, consolidate.cpp:78-88, aborts),
  * all arithmetic in float32 in Magnum's operation order: tests/test_oracle_ref.py holds this module byte for byte
    against the reference's own consolidation run on the same assets (oracle/_ref/meshtool),
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


_F = np.float32


def _quat_to_mat(q):
    """Magnum Math::Quaternion::toMatrix in float32, operation for operation (contrib/magnum/src/Magnum/Math/Quaternion.h:779-791);
    returns M[row, col]."""
    x, y, z, w = (_F(v) for v in q)
    one, two = _F(1), _F(2)
    c0 = (one - two * (y * y) - two * (z * z), two * x * y + two * z * w, two * x * z - two * y * w)
    c1 = (two * x * y - two * z * w, one - two * (x * x) - two * (z * z), two * y * z + two * x * w)
    c2 = (two * x * z + two * y * w, two * y * z - two * x * w, one - two * (x * x) - two * (y * y))
    return np.array([c0, c1, c2], np.float32).T


def _matmul32(a, b):
    """Magnum RectangularMatrix::operator* in float32: every element is ((0 + a_r0 b_0c) + a_r1 b_1c) + ... in that order."""
    out = np.zeros((4, 4), np.float32)
    for pos in range(4):
        out = (out + np.outer(a[:, pos], b[pos, :]).astype(np.float32)).astype(np.float32)
    return out


def _node_matrix(node):
    """Trade::ObjectData3D::transformation() of the vendored CgltfImporter (float32): the node matrix as given, or
    Matrix4::from(rotation.toMatrix(), translation) * Matrix4::scaling(scale)
    (contrib/magnum/src/Magnum/Trade/ObjectData3D.cpp:72-79, CgltfImporter.cpp:1398-1414)."""
    if "matrix" in node:
        return np.array(node["matrix"], np.float32).reshape(4, 4).T.copy()     # glTF matrices are column-major
    q = np.array(node.get("rotation", [0, 0, 0, 1]), np.float32)
    if abs(float(np.dot(q, q)) - 1.0) >= 2e-5:                                  # !isNormalized() -> renormalised
        q = (q * (_F(1) / np.sqrt(np.dot(q, q), dtype=np.float32))).astype(np.float32)
    rt = np.eye(4, dtype=np.float32)
    rt[:3, :3] = _quat_to_mat(q)
    rt[:3, 3] = np.array(node.get("translation", [0, 0, 0]), np.float32)
    sc = np.eye(4, dtype=np.float32)
    sc[0, 0], sc[1, 1], sc[2, 2] = (_F(v) for v in node.get("scale", [1, 1, 1]))
    return _matmul32(rt, sc)


def _transform32(m, v, w):
    """Matrix4::transformPoint (w = 1) / transformVector (w = 0) over an array of float32 vectors, Magnum's summation order."""
    out = np.zeros((len(v), 4), np.float32)
    cols = (v[:, 0], v[:, 1], v[:, 2], np.full(len(v), w, np.float32))
    for pos in range(4):
        out = out + cols[pos][:, None] * m[None, :, pos]
    if w:
        return (out[:, :3] / out[:, 3:4]).astype(np.float32)
    return np.ascontiguousarray(out[:, :3])


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


def _normalized32(v):
    ln = np.sqrt(np.einsum("ij,ij->i", v, v, dtype=np.float32), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (v * (_F(1) / ln)[:, None]).astype(np.float32)


def _smooth_normals(pos, idx):
    """Magnum MeshTools::generateSmoothNormals (contrib/magnum/src/Magnum/MeshTools/GenerateNormals.cpp:101-222): the face
    cross product (area) weighted by the interior angle at the vertex, accumulated per vertex in triangle order, normalised."""
    tri = idx.reshape(-1, 3)
    v0, v1, v2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    cr = np.cross(v2 - v1, v0 - v1).astype(np.float32)
    v10n, v20n, v21n = _normalized32(v1 - v0), _normalized32(v2 - v0), _normalized32(v2 - v1)
    bad = np.isnan(v10n).any(1) | np.isnan(v20n).any(1) | np.isnan(v21n).any(1)
    a0 = np.arccos(np.clip(np.einsum("ij,ij->i", v10n, v20n, dtype=np.float32), -1, 1)).astype(np.float32)
    a1 = np.arccos(np.clip(np.einsum("ij,ij->i", -v10n, v21n, dtype=np.float32), -1, 1)).astype(np.float32)
    a2 = (_F(np.pi) - a0 - a1).astype(np.float32)
    ang = np.stack([a0, a1, a2], 1)
    ang[bad] = 0
    n = np.zeros_like(pos)
    np.add.at(n, tri.reshape(-1), (cr[:, None, :] * ang[:, :, None]).reshape(-1, 3))     # in index order = triangle order per vertex
    return _normalized32(n)


def _compute_tangents(pos, uv, idx):
    """compute_tangents.cpp:53-110: per-face tangent from UV deltas, summed per vertex in face order, divided by the
    vertex degree, normalised (float32 throughout)."""
    tri = idx.reshape(-1, 3)
    d1, d2 = pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]]
    u1, u2 = uv[tri[:, 1]] - uv[tri[:, 0]], uv[tri[:, 2]] - uv[tri[:, 0]]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (_F(1) / (u1[:, 0] * u2[:, 1] - u1[:, 1] * u2[:, 0])).astype(np.float32)
        t = ((d1 * u2[:, 1:2] - d2 * u1[:, 1:2]) * r[:, None]).astype(np.float32)
    acc = np.zeros_like(pos)
    deg = np.zeros(len(pos), np.float32)
    np.add.at(acc, tri.reshape(-1), np.repeat(t, 3, axis=0))
    np.add.at(deg, tri.reshape(-1), 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = (acc / deg[:, None]).astype(np.float32)
    return _normalized32(acc)


def load(path, name=None, generate_missing_normals=False):
    """Load a .gltf / .glb file into a MeshData (consolidated, reference vertex layout)."""
    doc = _Doc(path)
    js = doc.js
    verts, inds, subs = [], [], []
    v_off = i_off = 0

    def visit(ni, parent):
        nonlocal v_off, i_off
        node = js["nodes"][ni]
        m = _matmul32(parent, _node_matrix(node))
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
                if "NORMAL" in at:
                    nrm = doc.accessor(at["NORMAL"]).astype(np.float32)
                elif generate_missing_normals:
                    nrm = _smooth_normals(pos, idx)
                else:
                    # the reference does not survive such a mesh: consolidate.cpp:84-87 passes the vertex count as the new
                    # attribute's ARRAY size and Magnum aborts ("Normal can't be an array attribute") — probe: oracle/_ref/meshtool
                    raise ValueError(f"{path}: a primitive has no NORMAL attribute (the reference aborts on such meshes; "
                                     "pass generate_missing_normals=True for Magnum-style smooth normals)")
                if "TANGENT" in at:
                    tan = doc.accessor(at["TANGENT"]).astype(np.float32)[:, :3]
                elif uv is not None:
                    tan = _compute_tangents(pos, uv, idx)
                else:
                    tan = np.zeros_like(pos)
                v = np.zeros(len(pos), abi.VERTEX_DTYPE)                                 # the consolidated buffer is zero-initialised
                v["position"] = _transform32(m, pos, 1.0)                                # transformPoint (consolidate.cpp:255)
                v["normal"] = _transform32(m, nrm, 0.0)                                  # transformVector, not renormalised (:294)
                v["tangent"][:, :3] = _transform32(m, tan, 0.0)
                v["tangent"][:, 3] = 1.0                                                 # consolidate.cpp:278
                if uv is not None:
                    v["uv"] = uv
                if "COLOR_0" in at:                                                      # Color4 only (transferAttribute<Color4>, :270)
                    col = doc.accessor(at["COLOR_0"])
                    if col.dtype != np.float32 or col.shape[1] != 4:
                        raise ValueError(f"{path}: COLOR_0 must be float VEC4 (the reference asserts on any other format)")
                    v["color"] = col
                verts.append(v)
                inds.append(idx + v_off)
                subs.append((i_off, len(idx), prim.get("material", -1)))
                v_off += len(pos)
                i_off += len(idx)
        for c in node.get("children", []):
            visit(c, m)

    scene = js["scenes"][js.get("scene", 0)]
    for ni in scene["nodes"]:
        visit(ni, np.eye(4, dtype=np.float32))
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
                                      (float(em[0]), float(em[1]), float(em[2]), 1.0), metallic, roughness,   # emissiveColor(): Color3 -> Color4
                                      tex_index(pbr.get("baseColorTexture")), tex_index(mat.get("normalTexture")), mr_tex,
                                      tex_index(mat.get("emissiveTexture")), tex_index(mat.get("occlusionTexture"))))
    subs = [(o, c, m if 0 <= m < len(materials) else -1) for o, c, m in subs]
    return MeshData(vertices, indices, subs, materials, images, name=name or os.path.basename(path))
