"""Stanford PLY front end: `sl.Mesh('obj_000001.ply')` — the mesh format of the BOP / T-LESS / LINEMOD model sets.

The reference imports PLY through Magnum's plugin set (StanfordImporter / Assimp, src/mesh.cpp:176-248) and consolidates
the result like every other format (src/mesh_tools/consolidate.cpp:53-335); this loader produces the same MeshData record
gltf.load() does. Supported: ascii, binary_little_endian and binary_big_endian files; vertex properties x y z, nx ny nz,
red green blue [alpha] (uchar or float), s t / u v / texture_u texture_v; faces as `vertex_indices` / `vertex_index`
lists (polygons are fanned), extra properties and elements are skipped. Vertex colours land in the colour slot of the
68-byte stream (the renderer does not read it, exactly like the reference's shader); a `comment TextureFile <name>` next
to texture coordinates becomes the base-colour texture. Missing normals: Magnum-style smooth normals.
Not pinned against the reference (convenience of the Python surface, outside the §8 hot path).
"""
import os
import struct

import numpy as np

from . import abi
from .desc import ImageData, MaterialData, MeshData
from .gltf import _compute_tangents, _smooth_normals

_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
          "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def load(path, name=None):
    path = os.fspath(path)
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        raise RuntimeError(f"Could not load mesh {path}: not a PLY file")
    header = raw[:end].decode("ascii", "replace").splitlines()
    body = raw[raw.index(b"\n", end) + 1:]
    fmt, elements, texture_file = None, [], None
    for line in header[1:]:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "comment" and len(tok) >= 3 and tok[1].lower() == "texturefile":
            texture_file = " ".join(tok[2:])
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property" and elements:
            if tok[1] == "list":
                elements[-1]["props"].append(("list", _TYPES[tok[2]], _TYPES[tok[3]], tok[4]))
            else:
                elements[-1]["props"].append(("scalar", _TYPES[tok[1]], None, tok[2]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise RuntimeError(f"Could not load mesh {path}: unsupported PLY format {fmt}")
    order = ">" if fmt == "binary_big_endian" else "<"
    vertex, faces = None, []
    if fmt == "ascii":
        tokens = body.split()
        at = 0
    else:
        at = 0
    for el in elements:
        scalar_only = all(p[0] == "scalar" for p in el["props"])
        if fmt == "ascii":
            if scalar_only:
                n = el["count"] * len(el["props"])
                data = np.array(tokens[at:at + n], dtype=np.float64).reshape(el["count"], len(el["props"]))
                at += n
                table = {p[3]: data[:, k] for k, p in enumerate(el["props"])}
            else:
                table, rows = None, []
                for _ in range(el["count"]):
                    row = {}
                    for kind, t0, t1, pname in el["props"]:
                        if kind == "list":
                            cnt = int(tokens[at]); at += 1
                            row[pname] = [int(float(x)) for x in tokens[at:at + cnt]]; at += cnt
                        else:
                            row[pname] = float(tokens[at]); at += 1
                    rows.append(row)
        else:
            if scalar_only:
                dt = np.dtype([(p[3], order + p[1]) for p in el["props"]])
                arr = np.frombuffer(body, dt, el["count"], at)
                at += dt.itemsize * el["count"]
                table = {p[3]: arr[p[3]] for p in el["props"]}
            else:
                table, rows = None, []
                for _ in range(el["count"]):
                    row = {}
                    for kind, t0, t1, pname in el["props"]:
                        if kind == "list":
                            cdt = np.dtype(order + t0)
                            cnt = int(np.frombuffer(body, cdt, 1, at)[0]); at += cdt.itemsize
                            idt = np.dtype(order + t1)
                            row[pname] = np.frombuffer(body, idt, cnt, at).astype(np.int64).tolist(); at += idt.itemsize * cnt
                        else:
                            sdt = np.dtype(order + t0)
                            row[pname] = float(np.frombuffer(body, sdt, 1, at)[0]); at += sdt.itemsize
                    rows.append(row)
        if el["name"] == "vertex":
            if table is None:
                raise RuntimeError(f"Could not load mesh {path}: list properties on vertices are not supported")
            vertex = table
        elif el["name"] == "face":
            key = "vertex_indices" if any(p[3] == "vertex_indices" for p in el["props"]) else "vertex_index"
            if table is not None:
                raise RuntimeError(f"Could not load mesh {path}: faces without an index list")
            for row in rows:
                poly = row[key]
                for k in range(1, len(poly) - 1):
                    faces.append((poly[0], poly[k], poly[k + 1]))
    if vertex is None or not faces:
        raise RuntimeError(f"Could not load mesh {path}: no triangle meshes")
    n = len(vertex["x"])
    idx = np.asarray(faces, np.uint32).reshape(-1)
    if idx.max() >= n:
        raise RuntimeError(f"Could not load mesh {path}: face index out of range")
    v = np.zeros(n, abi.VERTEX_DTYPE)
    v["position"] = np.stack([vertex["x"], vertex["y"], vertex["z"]], 1).astype(np.float32)
    if all(k in vertex for k in ("nx", "ny", "nz")):
        v["normal"] = np.stack([vertex["nx"], vertex["ny"], vertex["nz"]], 1).astype(np.float32)
    else:
        v["normal"] = np.nan_to_num(_smooth_normals(np.ascontiguousarray(v["position"]), idx))
    uv = None
    for a, b in (("s", "t"), ("u", "v"), ("texture_u", "texture_v")):
        if a in vertex and b in vertex:
            uv = np.stack([vertex[a], vertex[b]], 1).astype(np.float32)
            break
    if uv is not None:
        v["uv"] = uv
        v["tangent"][:, :3] = np.nan_to_num(_compute_tangents(np.ascontiguousarray(v["position"]), uv, idx))
    v["tangent"][:, 3] = 1.0
    if all(k in vertex for k in ("red", "green", "blue")):
        scale = 255.0 if np.asarray(vertex["red"]).dtype.kind in "ui" or (fmt == "ascii" and max(vertex["red"].max(), vertex["green"].max(), vertex["blue"].max()) > 1.0) else 1.0
        col = np.stack([vertex["red"], vertex["green"], vertex["blue"], vertex.get("alpha", np.full(n, scale))], 1).astype(np.float32) / np.float32(scale)
        v["color"] = col
    v["vertex_index"] = np.arange(1, n + 1, dtype=np.uint32)
    images, tex = [], -1
    if uv is not None and texture_file:
        tpath = os.path.join(os.path.dirname(path), texture_file)
        if os.path.isfile(tpath):
            from PIL import Image
            pil = Image.open(tpath)
            pil = pil.convert("RGBA" if "A" in pil.getbands() else "RGB")
            images.append(ImageData(np.ascontiguousarray(np.asarray(pil)[::-1])))
            tex = 0
    material = MaterialData(tex_base_color=tex)
    return MeshData(v, idx, [(0, len(idx), 0)], [material], images, name=name or os.path.basename(path))


def _write_test_ply(path, positions, faces, fmt="ascii", normals=None, colors=None, uv=None, texture_file=None):
    """Writer used by the tests (and handy for stand-in assets): triangles or polygons, optional normals / uchar colours / uv."""
    n = len(positions)
    lines = ["ply", f"format {fmt} 1.0"]
    if texture_file:
        lines.append(f"comment TextureFile {texture_file}")
    lines += [f"element vertex {n}", "property float x", "property float y", "property float z"]
    if normals is not None:
        lines += ["property float nx", "property float ny", "property float nz"]
    if colors is not None:
        lines += ["property uchar red", "property uchar green", "property uchar blue"]
    if uv is not None:
        lines += ["property float texture_u", "property float texture_v"]
    lines += [f"element face {len(faces)}", "property list uchar int vertex_indices", "end_header"]
    head = ("\n".join(lines) + "\n").encode()
    with open(path, "wb") as f:
        f.write(head)
        if fmt == "ascii":
            for i in range(n):
                row = [f"{x:.7g}" for x in positions[i]]
                if normals is not None:
                    row += [f"{x:.7g}" for x in normals[i]]
                if colors is not None:
                    row += [str(int(c)) for c in colors[i]]
                if uv is not None:
                    row += [f"{x:.7g}" for x in uv[i]]
                f.write((" ".join(row) + "\n").encode())
            for poly in faces:
                f.write((" ".join([str(len(poly))] + [str(int(i)) for i in poly]) + "\n").encode())
        else:
            o = "<" if fmt == "binary_little_endian" else ">"
            for i in range(n):
                f.write(struct.pack(o + "3f", *positions[i]))
                if normals is not None:
                    f.write(struct.pack(o + "3f", *normals[i]))
                if colors is not None:
                    f.write(struct.pack("3B", *[int(c) for c in colors[i]]))
                if uv is not None:
                    f.write(struct.pack(o + "2f", *uv[i]))
            for poly in faces:
                f.write(struct.pack("B", len(poly)) + struct.pack(o + f"{len(poly)}i", *[int(i) for i in poly]))
