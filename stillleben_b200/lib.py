"""Loader of the native library (stillleben_b200/libslb.so) and thin object wrappers over the C ABI.

The product path has NO fallback: if the CUDA library is missing or cannot create a context this
module raises. Nothing here imports or calls the CPU oracle.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .desc import DescBatch, ImageData, LightMapData, MeshData

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslb.so")

_lib = None


class SlbError(RuntimeError):
    pass


def load():
    """dlopen libslb.so and attach prototypes for every symbol of include/slb.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlbError(f"{LIB_PATH} is missing: build it with `make -C stillleben_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
        _lib = abi.bind(C.CDLL(LIB_PATH))
        if _lib.slb_abi_version() != abi.SLB_ABI_VERSION:
            raise SlbError("libslb.so ABI version mismatch")
    return _lib


def _raise(lib, ctx, rc, what):
    msg = lib.slb_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == abi.ERR_INVALID_ARGUMENT:
        raise ValueError(f"{what}: {msg}")      # reference: std::invalid_argument -> ValueError
    raise SlbError(f"{what} failed (status {rc}): {msg}")


class Result:
    """RenderPass::Result replacement: n_frames frames of the selected targets in device memory."""

    def __init__(self, ctx, width, height, n_frames, target_mask=abi.TARGETS_SIX, torch_tensors=False):
        self.ctx, self.W, self.H, self.n_frames, self.mask = ctx, width, height, n_frames, target_mask
        self.tensors = None
        ext = None
        if torch_tensors:
            import torch
            dev = torch.device("cuda", ctx.device)
            tdt = {np.uint8: torch.uint8, np.float32: torch.float32, np.uint16: torch.int16, np.uint32: torch.int32}
            self.tensors = {}
            ext = (C.c_void_p * abi.NUM_TARGETS)()
            for t, (dt, ch) in enumerate(abi.TARGET_FORMATS):
                if target_mask & (1 << t):
                    ten = torch.empty((n_frames, height, width, ch), dtype=tdt[dt], device=dev)
                    self.tensors[t] = ten
                    ext[t] = ten.data_ptr()
        h = C.c_void_p()
        rc = ctx.lib.slb_result_create(ctx.h, width, height, n_frames, target_mask, ext, C.byref(h))
        if rc != abi.OK:
            _raise(ctx.lib, ctx.h, rc, "slb_result_create")
        self.h = h

    def numpy(self, target, first_frame=0, n_frames=None):
        n = self.n_frames - first_frame if n_frames is None else n_frames
        dt, ch = abi.TARGET_FORMATS[target]
        out = np.empty((n, self.H, self.W, ch), dt)
        rc = self.ctx.lib.slb_result_read(self.ctx.h, self.h, target, first_frame, n, out.ctypes.data, out.nbytes)
        if rc != abi.OK:
            _raise(self.ctx.lib, self.ctx.h, rc, "slb_result_read")
        return out

    def hdr(self, frame=0):
        out = np.empty((self.H, self.W, 4), np.float32)
        rc = self.ctx.lib.slb_result_read_hdr(self.ctx.h, self.h, frame, out.ctypes.data, out.size)
        if rc != abi.OK:
            _raise(self.ctx.lib, self.ctx.h, rc, "slb_result_read_hdr")
        return out

    def frame_dict(self, frame=0):
        """All targets of one frame as numpy arrays keyed like the oracle's output."""
        return {abi.TARGET_NAMES[t]: self.numpy(t, frame, 1)[0] for t in range(abi.NUM_TARGETS) if self.mask & (1 << t)}

    def ptrs(self):
        p = (C.c_void_p * abi.NUM_TARGETS)()
        b = (C.c_size_t * abi.NUM_TARGETS)()
        self.ctx.lib.slb_result_ptrs(self.h, p, b)
        return list(p), list(b)

    def close(self):
        if self.h:
            self.ctx.lib.slb_result_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """sl::Context replacement for the render path: one CUDA device, one work stream."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.slb_ctx_create(device, C.byref(h))
        if rc != abi.OK:
            _raise(self.lib, None, rc, "slb_ctx_create")
        self.h = h
        self.device = device
        self._handles = {}
        self._kinds = {}         # id(obj) -> "mesh" | "texture" | "lightmap" (for release / close)
        self._keep = {}          # id(obj) -> obj: keeps the host arrays alive as long as the device handle exists
        self._pinned = []
        self.lightmap_sizes = (0, 0, 0, 0, 0)

    # ---- assets -------------------------------------------------------------------------------
    def handle_of(self, obj):
        key = id(obj)
        if key in self._handles:
            return self._handles[key]
        if isinstance(obj, MeshData):
            h, kind = self._upload_mesh(obj), "mesh"
        elif isinstance(obj, ImageData):
            h, kind = self._create_texture(obj), "texture"
        elif isinstance(obj, LightMapData):
            h, kind = self._create_lightmap(obj), "lightmap"
        else:
            raise TypeError(type(obj))
        self._handles[key] = h
        self._kinds[key] = kind
        self._keep[key] = obj
        return h

    def release(self, obj):
        """Free the device copy of a mesh / texture / light map (it is uploaded again if it is used again)."""
        key = id(obj)
        h = self._handles.pop(key, None)
        if h is None or not self.h:
            return
        kind = self._kinds.pop(key)
        self._keep.pop(key, None)
        {"mesh": self.lib.slb_mesh_destroy, "texture": self.lib.slb_texture_destroy, "lightmap": self.lib.slb_lightmap_destroy}[kind](self.h, h)

    def _upload_mesh(self, m):
        subs = (abi.Submesh * len(m.submeshes))(*[abi.Submesh(o, c, mat, 0) for o, c, mat in m.submeshes])
        mats = (abi.Material * max(1, len(m.materials)))(*[x.to_c() for x in m.materials])
        imgs = (abi.Image * max(1, len(m.images)))(*[x.to_c() for x in m.images])
        v = np.ascontiguousarray(m.vertices)
        bmin = (C.c_float * 3)(*np.asarray(m.bbox_min, np.float32).tolist())
        bmax = (C.c_float * 3)(*np.asarray(m.bbox_max, np.float32).tolist())
        h = C.c_void_p()
        rc = self.lib.slb_mesh_upload(self.h, v.ctypes.data, len(v), m.indices.ctypes.data, len(m.indices), subs,
                                      len(m.submeshes), mats, len(m.materials), imgs, len(m.images), bmin, bmax, C.byref(h))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_upload")
        return h.value

    def update_vertices(self, mesh):
        v = np.ascontiguousarray(mesh.vertices)
        rc = self.lib.slb_mesh_update_vertices(self.h, self.handle_of(mesh), v.ctypes.data, len(v))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_update_vertices")

    @staticmethod
    def _ptr(a):
        """Device pointer of a CUDA torch tensor, host pointer of a numpy array, None for None."""
        if a is None:
            return None
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def update_positions_and_colors(self, mesh, vertex_ids, position_update=None, color_update=None, stream=None):
        """Mesh::updateVertexPositionsAndColors on the device copy: ids int32 (one-based), updates float32 n x 3 / n x 4;
        numpy arrays or CUDA tensors."""
        rc = self.lib.slb_mesh_update_positions_and_colors(self.h, self.handle_of(mesh), self._ptr(vertex_ids), len(vertex_ids),
                                                           self._ptr(position_update), self._ptr(color_update), stream)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_update_positions_and_colors")

    def set_positions(self, mesh, positions, stream=None):
        rc = self.lib.slb_mesh_set_positions(self.h, self.handle_of(mesh), self._ptr(positions), len(positions), stream)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_set_positions")

    def set_colors(self, mesh, colors, stream=None):
        rc = self.lib.slb_mesh_set_colors(self.h, self.handle_of(mesh), self._ptr(colors), len(colors), stream)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_set_colors")

    def recompute_normals(self, mesh, stream=None):
        rc = self.lib.slb_mesh_recompute_normals(self.h, self.handle_of(mesh), stream)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_recompute_normals")

    def read_vertices(self, mesh):
        """The device copy's current 68-byte vertex stream as a structured numpy array."""
        out = np.empty(len(mesh.vertices), abi.VERTEX_DTYPE)
        rc = self.lib.slb_mesh_read_vertices(self.h, self.handle_of(mesh), out.ctypes.data, len(out))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_mesh_read_vertices")
        return out

    def _create_texture(self, img):
        ci = img.to_c()
        h = C.c_void_p()
        rc = self.lib.slb_texture_create(self.h, C.byref(ci), img.kind, C.byref(h))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_texture_create")
        return h.value

    def texture_level(self, img, level):
        h = self.handle_of(img)
        w, hh = C.c_int32(), C.c_int32()
        n = self.lib.slb_texture_read_level(self.h, h, level, C.byref(w), C.byref(hh), None)
        if n <= 0:
            _raise(self.lib, self.h, n, "slb_texture_read_level")
        out = np.empty((hh.value, w.value, 4), np.uint8)
        self.lib.slb_texture_read_level(self.h, h, level, C.byref(w), C.byref(hh), out.ctypes.data)
        return n, out

    def _create_lightmap(self, lm):
        d = abi.LightmapDesc()
        eq = np.ascontiguousarray(lm.equirect, np.float32)
        d.equirect_rgb = eq.ctypes.data
        d.height, d.width = eq.shape[0], eq.shape[1]
        d.n_lights = len(lm.light_directions)
        for i in range(d.n_lights):
            for k in range(3):
                d.light_directions[i][k] = float(lm.light_directions[i][k])
                d.light_colors[i][k] = float(lm.light_colors[i][k])
        h = C.c_void_p()
        if lm.maps is not None:                        # precomputed on another rank, received in the asset broadcast
            env0, irr, pre, lut = (np.ascontiguousarray(a, np.float32) for a in lm.maps)
            pre_size = int(round(np.sqrt(pre.size / (6 * 4) / sum(0.25 ** k for k in range(5)))))
            rc = self.lib.slb_lightmap_create_from_maps(self.h, C.byref(d), env0.ctypes.data, env0.shape[1], irr.ctypes.data, irr.shape[1],
                                                        pre.ctypes.data, pre_size, lut.ctypes.data, lut.shape[0], C.byref(h))
        else:
            rc = self.lib.slb_lightmap_create_ex(self.h, C.byref(d), *self.lightmap_sizes, C.byref(h))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_lightmap_create")
        return h.value

    def read_lightmap(self, lm):
        """(env level 0, irradiance, prefilter packed, LUT) as numpy arrays."""
        h = self.handle_of(lm)
        sizes = (C.c_int32 * 4)()
        self.lib.slb_lightmap_sizes(h, sizes)
        e, i, p, l = sizes
        counts = [6 * e * e * 4, 6 * i * i * 4, sum(6 * (p >> m) ** 2 * 4 for m in range(5)), l * l * 4]
        out = []
        for which, n in enumerate(counts):
            a = np.empty(n, np.float32)
            rc = self.lib.slb_lightmap_read(self.h, h, which, a.ctypes.data, n)
            if rc != abi.OK:
                _raise(self.lib, self.h, rc, "slb_lightmap_read")
            out.append(a)
        return out[0].reshape(6, e, e, 4), out[1].reshape(6, i, i, 4), out[2], out[3].reshape(l, l, 4)

    # ---- rendering ----------------------------------------------------------------------------
    def descs(self, scenes):
        return DescBatch(scenes, self.handle_of)

    def render(self, scenes, result=None, first_frame=0, depth_peel=None, target_mask=abi.TARGETS_SIX, stream=None, descs=None):
        """slb_render_batch: queue the batch; returns the Result (not synchronised)."""
        descs = descs or self.descs(scenes)
        if result is None:
            result = Result(self, scenes[0].width, scenes[0].height, descs.n, target_mask)
        rc = self.lib.slb_render_batch(self.h, descs.ptr, descs.n, result.h, first_frame,
                                       depth_peel.h if depth_peel is not None else None, stream)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_render_batch")
        return result

    def render_host(self, descs, host_arrays, target_mask=abi.TARGETS_SIX):
        """slb_render_batch_host: host buffers end to end (synchronous)."""
        ptrs = (C.c_void_p * abi.NUM_TARGETS)()
        for t in range(abi.NUM_TARGETS):
            if target_mask & (1 << t):
                a = host_arrays[t]
                ptrs[t] = a if isinstance(a, int) else a.ctypes.data
        rc = self.lib.slb_render_batch_host(self.h, descs.ptr, descs.n, target_mask, ptrs)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_render_batch_host")

    def host_alloc(self, shape, dtype):
        """numpy array over page-locked host memory (slb_host_alloc); freed with the context."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        rc = self.lib.slb_host_alloc(self.h, n, C.byref(p))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_host_alloc")
        self._pinned.append(p.value)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def synchronize(self):
        rc = self.lib.slb_ctx_synchronize(self.h)
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_ctx_synchronize")

    def set_option(self, option, value):
        rc = self.lib.slb_ctx_set_option(self.h, option, int(value))
        if rc != abi.OK:
            _raise(self.lib, self.h, rc, "slb_ctx_set_option")

    def stats(self):
        s = abi.Stats()
        self.lib.slb_ctx_get_stats(self.h, C.byref(s))
        return s

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.slb_host_free(self.h, p)
            self._pinned = []
            for obj in list(self._keep.values()):       # meshes, textures, light maps: device memory goes with the context
                self.release(obj)
            self.lib.slb_ctx_destroy(self.h)
            self.h = None
