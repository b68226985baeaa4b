"""Wavefront OBJ (+ MTL) front end: `sl.Mesh('…/textured.obj')` as examples/ycb.py:42 uses it.

The reference imports OBJ through Magnum's AssimpImporter (src/mesh.cpp:176-248) and then runs the same consolidation as
for glTF (src/mesh_tools/consolidate.cpp:53-335), so the result here is the same MeshData record gltf.load() produces:
68-byte vertices (position, uv, normal, tangent with w = 1, colour, one-based vertex id), u32 indices, one sub-mesh per
`usemtl` group. What Assimp does for the YCB models and this loader repeats:
  * a vertex per distinct (v, vt, vn) triple (aiProcess_JoinIdenticalVertices), polygons fanned into triangles
    (aiProcess_Triangulate), negative (relative) indices resolved;
  * `Kd` -> base colour, `map_Kd` -> base colour texture (row 0 = bottom row, as every imported image), `Ke` -> emissive;
    OBJ texture coordinates already have their origin at the bottom left, so they are NOT flipped (glTF ones are);
  * faces without `vn`: Magnum-style smooth normals (gltf._smooth_normals); tangents from the UV deltas
    (compute_tangents.cpp:53-110) when the mesh has texture coordinates.
Not pinned against the reference (Assimp is not in the tree): OBJ support is a convenience of the Python surface, outside
the §8 hot path.
"""
import os

import numpy as np

from . import abi
from .desc import ImageData, MaterialData, MeshData
from .gltf import _compute_tangents, _smooth_normals


def _load_mtl(path):
    mats, cur = {}, None
    if not os.path.isfile(path):
        return mats
    for line in open(path, errors="ignore"):
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        key = tok[0]
        if key == "newmtl":
            cur = mats.setdefault(" ".join(tok[1:]), {})
        elif cur is None:
            continue
        elif key in ("Kd", "Ke"):
            cur[key] = tuple(float(x) for x in tok[1:4])
        elif key == "d":
            cur["d"] = float(tok[1])
        elif key == "map_Kd":
            cur["map_Kd"] = os.path.join(os.path.dirname(path), tok[-1])       # options (-s, -o …) precede the file name
    return mats


def load(path, name=None):
    path = os.fspath(path)
    pos, uvs, nrms = [], [], []
    mtl = {}
    groups = {}                    # material name -> list of (v, vt, vn) corner triples, three per triangle
    order = []
    cur = None

    def group(mat):
        if mat not in groups:
            groups[mat] = []
            order.append(mat)
        return groups[mat]

    corners = group(None)
    for line in open(path, errors="ignore"):
        tok = line.split()
        if not tok:
            continue
        key = tok[0]
        if key == "v":
            pos.append([float(x) for x in tok[1:4]])
        elif key == "vt":
            uvs.append([float(tok[1]), float(tok[2]) if len(tok) > 2 else 0.0])
        elif key == "vn":
            nrms.append([float(x) for x in tok[1:4]])
        elif key == "mtllib":
            mtl.update(_load_mtl(os.path.join(os.path.dirname(path), " ".join(tok[1:]))))
        elif key == "usemtl":
            cur = " ".join(tok[1:])
            corners = group(cur)
        elif key == "f":
            face = []
            for c in tok[1:]:
                parts = (c.split("/") + ["", ""])[:3]
                iv = int(parts[0])
                it = int(parts[1]) if parts[1] else 0
                inn = int(parts[2]) if parts[2] else 0
                face.append((iv - 1 if iv > 0 else len(pos) + iv,
                             (it - 1 if it > 0 else len(uvs) + it) if it else -1,
                             (inn - 1 if inn > 0 else len(nrms) + inn) if inn else -1))
            for k in range(1, len(face) - 1):
                corners.extend((face[0], face[k], face[k + 1]))
    order = [m for m in order if groups[m]]
    if not order:
        raise RuntimeError(f"Could not load mesh {path}: no faces")
    P = np.asarray(pos, np.float32).reshape(-1, 3)
    T = np.asarray(uvs, np.float32).reshape(-1, 2)
    N = np.asarray(nrms, np.float32).reshape(-1, 3)

    verts, inds, subs, materials, images = [], [], [], [], []
    image_of = {}
    v_off = i_off = 0
    for mat in order:
        tri = np.asarray(groups[mat], np.int64).reshape(-1, 3)
        uniq, inv = np.unique(tri, axis=0, return_inverse=True)
        # keep first-occurrence order (Assimp numbers joined vertices in the order the faces reference them)
        first = np.full(len(uniq), len(tri), np.int64)
        np.minimum.at(first, inv.reshape(-1), np.arange(len(tri)))
        rank = np.argsort(first, kind="stable")
        remap = np.empty(len(uniq), np.int64)
        remap[rank] = np.arange(len(uniq))
        uniq, idx = uniq[rank], remap[inv.reshape(-1)].astype(np.uint32)
        v = np.zeros(len(uniq), abi.VERTEX_DTYPE)
        v["position"] = P[uniq[:, 0]]
        has_uv = len(T) > 0 and (uniq[:, 1] >= 0).all()
        if has_uv:
            v["uv"] = T[uniq[:, 1]]
        if len(N) > 0 and (uniq[:, 2] >= 0).all():
            v["normal"] = N[uniq[:, 2]]
        else:
            v["normal"] = np.nan_to_num(_smooth_normals(v["position"], idx))
        if has_uv:
            v["tangent"][:, :3] = np.nan_to_num(_compute_tangents(np.ascontiguousarray(v["position"]), np.ascontiguousarray(v["uv"]), idx))
        v["tangent"][:, 3] = 1.0
        m = mtl.get(mat, {})
        tex = -1
        if "map_Kd" in m and os.path.isfile(m["map_Kd"]):
            if m["map_Kd"] not in image_of:
                from PIL import Image
                pil = Image.open(m["map_Kd"])
                pil = pil.convert("RGBA" if "A" in pil.getbands() else "RGB")
                image_of[m["map_Kd"]] = len(images)
                images.append(ImageData(np.ascontiguousarray(np.asarray(pil)[::-1])))
            tex = image_of[m["map_Kd"]]
        kd = m.get("Kd", (1.0, 1.0, 1.0))
        ke = m.get("Ke", (0.0, 0.0, 0.0))
        materials.append(MaterialData((kd[0], kd[1], kd[2], m.get("d", 1.0)), (ke[0], ke[1], ke[2], 1.0), tex_base_color=tex))
        verts.append(v)
        inds.append(idx + np.uint32(v_off))
        subs.append((i_off, len(idx), len(materials) - 1))
        v_off += len(v)
        i_off += len(idx)
    vertices = np.concatenate(verts)
    vertices["vertex_index"] = np.arange(1, len(vertices) + 1, dtype=np.uint32)
    return MeshData(vertices, np.concatenate(inds), subs, materials, images, name=name or os.path.basename(path))
