"""ctypes front end of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from stillleben_b200 import abi
from stillleben_b200.desc import DescBatch, ImageData, LightMapData, MeshData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "slb.h"))
    if (not force and os.path.exists(ORACLE_SO)
            and os.path.getmtime(ORACLE_SO) >= max(os.path.getmtime(s) for s in srcs)):
        return ORACLE_SO
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(abi.Submesh), C.c_uint32,
                                      C.POINTER(abi.Material), C.c_uint32, C.POINTER(abi.Image), C.c_uint32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_mesh_destroy.argtypes = [C.c_void_p]
        L.orc_mesh_update_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_texture_create.restype = C.c_void_p
        L.orc_texture_create.argtypes = [C.POINTER(abi.Image), C.c_int]
        L.orc_texture_destroy.argtypes = [C.c_void_p]
        L.orc_texture_level.restype = C.c_int
        L.orc_texture_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.POINTER(abi.SceneDesc), C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
        L.orc_lightmap_create.restype = C.c_void_p
        L.orc_lightmap_create.argtypes = [C.POINTER(abi.LightmapDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_lightmap_from_maps.restype = C.c_void_p
        L.orc_lightmap_from_maps.argtypes = [C.POINTER(abi.LightmapDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_lightmap_read.restype = C.c_size_t
        L.orc_lightmap_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_lightmap_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_lightmap_destroy.argtypes = [C.c_void_p]
        L.orc_diff_sobel_valid_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_diff_dilate_object_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_int, C.c_int]
        _lib = L
    return _lib


def lightmap_desc(lm: LightMapData):
    d = abi.LightmapDesc()
    eq = np.ascontiguousarray(lm.equirect, np.float32)
    d.equirect_rgb = eq.ctypes.data
    d.height, d.width = eq.shape[0], eq.shape[1]
    d.n_lights = len(lm.light_directions)
    for i in range(d.n_lights):
        for k in range(3):
            d.light_directions[i][k] = float(lm.light_directions[i][k])
            d.light_colors[i][k] = float(lm.light_colors[i][k])
    return d, eq


class OracleAssets:
    """Creates (and caches) oracle handles for MeshData / ImageData / LightMapData objects."""

    def __init__(self, lightmap_sizes=(0, 0, 0, 0, 0)):
        self.L = lib()
        self.handles = {}
        self.keep = []
        self.lightmap_sizes = lightmap_sizes
        self.lightmap_maps = {}   # id(LightMapData) -> (env0, irr, pre, lut) numpy, to feed from_maps

    def handle_of(self, obj):
        key = id(obj)
        if key in self.handles:
            return self.handles[key]
        if isinstance(obj, MeshData):
            h = self._mesh(obj)
        elif isinstance(obj, ImageData):
            img = obj.to_c()
            h = self.L.orc_texture_create(C.byref(img), obj.kind)
        elif isinstance(obj, LightMapData):
            h = self._lightmap(obj)
        else:
            raise TypeError(type(obj))
        self.handles[key] = h
        self.keep.append(obj)
        return h

    def _mesh(self, m: MeshData):
        subs = (abi.Submesh * len(m.submeshes))(*[abi.Submesh(o, c, mat, 0) for o, c, mat in m.submeshes])
        mats = (abi.Material * max(1, len(m.materials)))(*[x.to_c() for x in m.materials])
        imgs = (abi.Image * max(1, len(m.images)))(*[x.to_c() for x in m.images])
        v = np.ascontiguousarray(m.vertices)
        bmin = (C.c_float * 3)(*m.bbox_min.tolist())
        bmax = (C.c_float * 3)(*m.bbox_max.tolist())
        return self.L.orc_mesh_create(v.ctypes.data, len(v), m.indices.ctypes.data, len(m.indices), subs,
                                      len(m.submeshes), mats, len(m.materials), imgs, len(m.images), bmin, bmax)

    def set_lightmap_maps(self, lm, env0, irr, pre, lut):
        self.lightmap_maps[id(lm)] = tuple(np.ascontiguousarray(a, np.float32) for a in (env0, irr, pre, lut))

    def _lightmap(self, lm: LightMapData):
        d, eq = lightmap_desc(lm)
        if id(lm) in self.lightmap_maps:
            env0, irr, pre, lut = self.lightmap_maps[id(lm)]
            pre_size = int(round(np.sqrt(pre.size / (6 * 4) / sum(0.25 ** k for k in range(5)))))
            return self.L.orc_lightmap_from_maps(C.byref(d), env0.ctypes.data, env0.shape[1], irr.ctypes.data,
                                                 irr.shape[1], pre.ctypes.data, pre_size, lut.ctypes.data, lut.shape[0])
        return self.L.orc_lightmap_create(C.byref(d), *self.lightmap_sizes)

    def read_lightmap(self, lm):
        h = self.handle_of(lm)
        sizes = (C.c_int * 4)()
        self.L.orc_lightmap_sizes(h, sizes)
        out = []
        for which in range(4):
            n = self.L.orc_lightmap_read(h, which, None)
            a = np.empty(n, np.float32)
            self.L.orc_lightmap_read(h, which, a.ctypes.data)
            out.append(a)
        e, i, l = sizes[0], sizes[1], sizes[3]
        return out[0].reshape(6, e, e, 4), out[1].reshape(6, i, i, 4), out[2], out[3].reshape(l, l, 4)


def render(scene, assets=None, peel=None, n_threads=0, want_hdr=True):
    """Render one SceneSpec with the oracle -> dict of numpy arrays named as abi.TARGET_NAMES (+ 'hdr')."""
    assets = assets or OracleAssets()
    batch = DescBatch([scene], assets.handle_of)
    H, W = scene.height, scene.width
    outs, ptrs = {}, (C.c_void_p * abi.NUM_TARGETS)()
    for t, (dt, ch) in enumerate(abi.TARGET_FORMATS):
        a = np.zeros((H, W, ch), dt)
        outs[abi.TARGET_NAMES[t]] = a
        ptrs[t] = a.ctypes.data
    hdr = np.zeros((H, W, 4), np.float32) if want_hdr else None
    peel_p = None
    if peel is not None:
        peel = np.ascontiguousarray(peel, np.float32)
        peel_p = peel.ctypes.data
    rc = lib().orc_render(batch.ptr, peel_p, ptrs, hdr.ctypes.data if want_hdr else None, n_threads)
    assert rc == 0
    if want_hdr:
        outs["hdr"] = hdr
    return outs
