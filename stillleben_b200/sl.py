"""Host-side mirror of the reference's Python surface for the render path.

    from stillleben_b200 import sl          # instead of:  import stillleben as sl

Same names, argument meanings, defaults, tensor dtypes / shapes and error behaviour as the pybind11 module
of the reference for everything RenderPass::render reads or returns:
  sl.init / sl.init_cuda            python/src/py_context.cpp:19-100
  sl.Mesh                           python/src/py_mesh.cpp:314-525, src/mesh.cpp:1000-1089
  sl.Object                         python/src/py_object.cpp:23-197
  sl.Scene                          python/src/py_scene.cpp:48-425, src/scene.cpp:195-330,453-470
  sl.LightMap                       python/src/py_light_map.cpp:19-52, src/light_map.cpp:62-152,266-376
  sl.Texture / sl.Texture2D         python/src/py_magnum.cpp:115-198
  sl.RenderPass / RenderPassResult  python/src/py_render_pass.cpp:92-279
Matrices are 4x4 float tensors indexed m[row, col] (as the reference's toTorch<Matrix4>). Everything here is
host logic over the C ABI (stillleben_b200.lib); rendering always runs in the CUDA library. Physics entry
points raise RuntimeError: PhysX stays on the host and is not part of this build.
"""
import math
import os
import warnings

import numpy as np
import torch

from . import abi, gltf
from . import lib as _lib
from .desc import (ImageData, LightMapData, MeshData, ObjectSpec, SceneSpec, fov_projection, intrinsics_projection,
                   inverted_rigid, look_at_pose)

_ctx = None
_cuda_index = 0


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
def init():
    """sl.init(): the reference creates a GL context without CUDA interop; here rendering always runs on a
    CUDA device, so this is init_cuda(0) with results returned as CPU tensors."""
    _init(0, False)


def init_cuda(device_index=0, use_cuda=True):
    _init(device_index, use_cuda)


_use_cuda = True


def _init(device_index, use_cuda):
    global _ctx, _cuda_index, _use_cuda
    if _ctx is not None:
        if device_index != _cuda_index or use_cuda != _use_cuda:      # py_context.cpp:36-44: only warns
            warnings.warn("stillleben context was already created with different CUDA settings")
        return
    _ctx = _lib.Context(device_index)
    _cuda_index, _use_cuda = device_index, use_cuda


def _context():
    if _ctx is None:
        raise RuntimeError("Call sl::init() first")                   # py_context.cpp:69-75
    return _ctx


def _np44(m):
    a = m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
    a = np.asarray(a, np.float32)
    if a.shape != (4, 4):
        raise ValueError("expected a 4x4 matrix")
    return a.copy()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


# ---------------------------------------------------------------------------------------------
# textures
# ---------------------------------------------------------------------------------------------
def _image_from(arg):
    if isinstance(arg, (str, os.PathLike)):
        from PIL import Image
        img = Image.open(arg)
        img = img.convert("RGBA" if "A" in img.getbands() else "RGB")
        return np.ascontiguousarray(np.asarray(img)[::-1])            # row 0 = bottom row
    t = arg.detach().cpu() if isinstance(arg, torch.Tensor) else torch.as_tensor(arg)
    if t.dim() != 3 or t.shape[2] != 3 or t.dtype != torch.uint8:
        raise ValueError("expected a HxWx3 uint8 CPU tensor")         # py_magnum.cpp:131-151
    return np.ascontiguousarray(t.numpy())                            # tensor row 0 becomes GL row 0


class Texture:
    """Rectangle texture (pixel coordinates, linear filter): background image / sticker. Loaded from a file it clamps to a transparent
    border (Context::loadTexture, src/context.cpp:596-598); built from a tensor it keeps GL's default for rectangle textures,
    clamp-to-edge (py_magnum.cpp:147-151 sets no wrapping) - the difference shows on the first texel row / column of a background
    image, which background_shader.frag samples on texel CORNERS (tests/test_gl_ref.py runs that shader on a real GL)."""

    def __init__(self, arg):
        wrap = abi.WRAP_CLAMP_TO_BORDER if isinstance(arg, (str, os.PathLike)) else abi.WRAP_CLAMP_TO_EDGE
        self.image = ImageData(_image_from(arg), wrap_s=wrap, wrap_t=wrap, min_filter=abi.FILTER_LINEAR, mag_filter=abi.FILTER_LINEAR,
                               kind=abi.TEXTURE_RECT)


class Texture2D:
    """Normalised-coordinate texture with a full mip chain: background plane texture."""

    def __init__(self, arg):
        self.image = ImageData(_image_from(arg), wrap_s=abi.WRAP_REPEAT, wrap_t=abi.WRAP_REPEAT,
                               min_filter=abi.FILTER_LINEAR_MIPMAP_LINEAR, mag_filter=abi.FILTER_LINEAR, kind=abi.TEXTURE_2D)


# ---------------------------------------------------------------------------------------------
# light map
# ---------------------------------------------------------------------------------------------
def _read_rgbe(path):
    """Radiance .hdr reader (RLE + flat), returns float32 HxWx3 with row 0 = top."""
    with open(path, "rb") as f:
        data = f.read()
    pos = data.index(b"\n\n") + 2
    end = data.index(b"\n", pos)
    dims = data[pos:end].split()
    H, W = int(dims[1]), int(dims[3])
    p = end + 1
    out = np.zeros((H, W, 4), np.uint8)
    buf = np.frombuffer(data, np.uint8)
    for y in range(H):
        if W >= 8 and W < 32768 and buf[p] == 2 and buf[p + 1] == 2 and (int(buf[p + 2]) << 8 | int(buf[p + 3])) == W:
            p += 4
            for c in range(4):
                x = 0
                while x < W:
                    n = int(buf[p]); p += 1
                    if n > 128:
                        n -= 128
                        out[y, x:x + n, c] = buf[p]; p += 1
                    else:
                        out[y, x:x + n, c] = buf[p:p + n]; p += n
                    x += n
        else:
            out[y] = buf[p:p + 4 * W].reshape(W, 4); p += 4 * W
    e = out[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    return out[..., :3].astype(np.float32) * scale[..., None]


class LightMap:
    """Image-based lighting: a lat-long HDR map (+ up to three directional lights from an sIBL .ibl file)."""

    def __init__(self, path=None):
        self.data = None
        self.path = None
        if path is not None:
            self.load(path)

    def load(self, path):
        path = os.fspath(path)
        lights_d, lights_c = [], []
        img_path = path
        if path.endswith(".ibl"):                                     # light_map.cpp:62-152: Corrade ini
            sections, cur = {}, None
            for line in open(path, errors="ignore"):
                line = line.strip()
                if line.startswith("[") and line.endswith("]"):
                    cur = line[1:-1]; sections[cur] = {}
                elif "=" in line and cur is not None:
                    k, v = line.split("=", 1)
                    sections[cur][k.strip()] = v.strip().strip('"')
            # LightMap::load + IBLSpec::load + LightSpec::load (light_map.cpp:55-151,266-345); a file the reference refuses makes
            # its constructor throw "Could not load light map <path>" (:259-262)
            ref = sections.get("Reflection")
            if ref is None or "REFfile" not in ref or "REFmap" not in ref or int(ref["REFmap"]) != 1:
                raise RuntimeError("Could not load light map " + path)
            img_path = os.path.join(os.path.dirname(path), ref["REFfile"])
            for sec, pre in (("Sun", "SUN"), ("Light1", "LIGHT"), ("Light2", "LIGHT")):
                if sec not in sections:
                    continue
                s = sections[sec]
                multi = np.float32(float(s[pre + "multi"])) if pre + "multi" in s else np.float32(1.0)
                col = np.ones(3, np.float32)                               # Color3{1.0f}: only a present colour is divided by 255
                if pre + "color" in s:
                    parts = s[pre + "color"].split(",")
                    if len(parts) != 3:
                        continue                                           # "Invalid light spec": this light is skipped
                    col = np.array([float(x) for x in parts], np.float32) / np.float32(255)
                col = multi * col
                u = np.float32(float(s[pre + "u"])) if pre + "u" in s else np.float32(0)
                v = np.float32(float(s[pre + "v"])) if pre + "v" in s else np.float32(0)
                theta, phi = float((u + np.float32(0.5)) * np.float32(math.pi) * np.float32(2)), float(v * np.float32(math.pi))
                lights_d.append((-np.array([math.cos(phi) * math.sin(theta), math.sin(phi) * math.sin(theta), math.cos(theta)])).tolist())
                lights_c.append(col.tolist())
        if img_path.lower().endswith(".hdr"):
            eq = _read_rgbe(img_path)
        elif img_path.lower().endswith(".npy"):
            eq = np.load(img_path).astype(np.float32)
        else:
            from PIL import Image
            eq = np.asarray(Image.open(img_path).convert("RGB"), np.float32) / 255.0
        self.data = LightMapData(np.ascontiguousarray(eq[::-1]), lights_d[:3], lights_c[:3])     # rows bottom-up
        self.path = path
        return True

    @staticmethod
    def from_equirect(equirect_bottom_up, light_directions=(), light_colors=()):
        """Extension: build a light map from a float32 HxWx3 lat-long image (row 0 = bottom)."""
        lm = LightMap()
        lm.data = LightMapData(np.ascontiguousarray(equirect_bottom_up, np.float32), [list(d) for d in light_directions],
                               [list(c) for c in light_colors])
        return lm


# ---------------------------------------------------------------------------------------------
# mesh
# ---------------------------------------------------------------------------------------------
class Range3D:
    """Magnum::Range3D as exposed by py_magnum.cpp:52-80 (min, max, center, size, diagonal)."""

    def __init__(self, lo, hi):
        self.min, self.max = _t(lo), _t(hi)

    def __repr__(self):
        return "Range3D(Vector(%g, %g, %g),Vector(%g, %g, %g))" % (*self.min.tolist(), *self.max.tolist())

    @property
    def center(self):
        return (self.min + self.max) / 2

    @property
    def size(self):
        return self.max - self.min

    @property
    def diagonal(self):
        return float(torch.linalg.norm(self.size))


_Range3D = Range3D


class Mesh:
    def __init__(self, filename, visual=True, physics=True, flags=None):
        _context()
        if isinstance(filename, MeshData):
            self.data, self.filename = filename, filename.name
        else:
            self.filename = os.fspath(filename)
            if self.filename.lower().endswith(".obj"):
                from . import objfile
                self.data = objfile.load(self.filename)
            elif self.filename.lower().endswith(".ply"):
                from . import plyfile
                self.data = plyfile.load(self.filename)
            elif self.filename.lower().endswith((".gltf", ".glb")):
                self.data = gltf.load(self.filename)
            else:
                raise RuntimeError(f"Could not load mesh {self.filename}: only glTF / GLB / OBJ / PLY are supported by this build")
        self._scale = 1.0
        self._rigid = np.eye(4, dtype=np.float32)
        self._class_index = 1                                          # mesh.h:300
        # physics=True is the reference default; V-HACD / PhysX cooking is skipped here (no PhysX in this build)

    @staticmethod
    def load_threaded(filenames, visual=True, physics=True, flags=None):
        return [Mesh(f, visual, physics) for f in filenames]

    @staticmethod
    def from_data(mesh_data):
        """Extension: wrap an in-memory consolidated mesh (stillleben_b200.desc.MeshData)."""
        return Mesh(mesh_data)

    # ---- pretransform (src/mesh.cpp:1000-1081) ----
    @property
    def pretransform(self):
        s = np.eye(4, dtype=np.float32)
        s[:3, :3] *= self._scale
        return _t(s @ self._rigid)

    @pretransform.setter
    def pretransform(self, m):
        m = _np44(m)
        u, w, vt = np.linalg.svd(m[:3, :3].astype(np.float64))
        if w.max() - w.min() > 1e-5:
            raise ValueError("Scaling is not uniform")
        self._scale = float((w.max() + w.min()) / 2.0)
        self._rigid = np.eye(4, dtype=np.float32)
        self._rigid[:3, :3] = (u @ vt).astype(np.float32)
        self._rigid[:3, 3] = (1.0 / self._scale) * m[:3, 3]

    def _pre_np(self):
        return self.pretransform.numpy()

    @property
    def bbox(self):
        p = self._pre_np().astype(np.float64)
        lo = p[:3, :3] @ self.data.bbox_min + p[:3, 3]
        hi = p[:3, :3] @ self.data.bbox_max + p[:3, 3]
        return Range3D(lo, hi)                                       # mesh.cpp:1075-1081: corners as transformed, not re-sorted

    def center_bbox(self):
        c = (self.data.bbox_min + self.data.bbox_max) / 2
        self._rigid[:3, 3] = -(self._rigid[:3, :3] @ c)

    def scale_to_bbox_diagonal(self, target_diagonal, mode="exact"):
        diag = float(np.linalg.norm(self.data.bbox_max - self.data.bbox_min))
        scale = target_diagonal / diag
        if mode == "exact":
            self._scale = scale
        elif mode == "order_of_magnitude":
            self._scale = float(10.0 ** round(math.log10(scale)))
        else:
            raise ValueError("invalid value for mode argument")

    @property
    def class_index(self):
        return self._class_index

    @class_index.setter
    def class_index(self, index):
        if index < 0 or index > 0xFFFF:
            raise ValueError("Mesh::setClassIndex(): out of range")    # mesh.cpp:1083-1089
        self._class_index = int(index)

    # ---- geometry access / editing (py_mesh.cpp:100-260,409-500; src/mesh.cpp:747-885) ----
    # Edits are applied to the DEVICE copy (slb_mesh_update_positions_and_colors / slb_mesh_set_positions: add, recompute
    # the area-weighted normals, no re-upload) and mirrored lazily into self.data when an accessor asks for them.
    def _sync_from_device(self):
        ctx = _context()
        if getattr(self, "_device_dirty", False) and id(self.data) in ctx._handles:
            self.data.vertices[:] = ctx.read_vertices(self.data)
            self._device_dirty = False

    @property
    def points(self):
        self._sync_from_device()
        return torch.from_numpy(self.data.vertices["position"].copy())

    @property
    def normals(self):
        self._sync_from_device()
        return torch.from_numpy(self.data.vertices["normal"].copy())

    @property
    def colors(self):
        self._sync_from_device()
        return torch.from_numpy(self.data.vertices["color"].copy())

    @property
    def faces(self):
        return torch.from_numpy(self.data.indices.astype(np.int32))

    @staticmethod
    def _check_update(vertex_indices, update, width, what):
        # the argument checks of py_mesh.cpp:100-212 (same messages)
        if vertex_indices.dim() != 1:
            raise ValueError("vertex_indices (1st argument) should be one dimensional")
        if update.dim() != 2:
            raise ValueError(f"{what} (2nd argument) should be two dimensional")
        if vertex_indices.size(0) != update.size(0):
            raise ValueError("vertices_index  and vertices_update should be of same size")
        if update.size(1) != width:
            raise ValueError(f"{what} should be of shape (N,{width})")
        if vertex_indices.device.type != "cpu" or update.device.type != "cpu":
            raise ValueError(f"vertex_indices and {what} should be CPU tensors")
        if not vertex_indices.is_contiguous():
            raise ValueError("vertex_indices should be contiguous tensor\n(use vertex_indices.contiguous() before passing vertex_indices as an argument)")
        if not update.is_contiguous():
            raise ValueError(f"{what} should be contiguous tensor\n(use {what}.contiguous() before passing {what} as an argument)")

    def _edit(self, ids, dpos, dcol):
        ctx = _context()
        ids_np = np.ascontiguousarray(ids.numpy(), np.int32)
        p = None if dpos is None else np.ascontiguousarray(dpos.numpy(), np.float32)
        c = None if dcol is None else np.ascontiguousarray(dcol.numpy(), np.float32)
        ctx.handle_of(self.data)                  # uploads the mesh if this is its first use
        self._sync_from_device()
        ctx.update_positions_and_colors(self.data, ids_np, p, c)
        self._device_dirty = True

    def update_positions(self, vertex_indices, position_update):
        self._check_update(vertex_indices, position_update, 3, "position_update")
        self._edit(vertex_indices, position_update, None)

    def update_colors(self, vertex_indices, color_update):
        self._check_update(vertex_indices, color_update, 4, "color_update")
        self._edit(vertex_indices, None, color_update)

    def update_positions_and_colors(self, vertex_indices, position_update, color_update):
        self._check_update(vertex_indices, position_update, 3, "position_update")
        self._check_update(vertex_indices, color_update, 4, "color_update")
        self._edit(vertex_indices, position_update, color_update)

    def set_new_positions(self, new_positions):
        if not new_positions.is_contiguous():
            raise ValueError("new_positions should be contiguous tensor\n(use new_positions.contiguous() before passing new_positions as an argument)")
        p = np.ascontiguousarray(new_positions.detach().cpu().numpy(), np.float32).reshape(-1, 3)
        ctx = _context()
        ctx.handle_of(self.data)
        ctx.set_positions(self.data, p)           # ValueError on a size mismatch, as Mesh::setVertexPositions
        self._device_dirty = True

    def set_new_colors(self, new_colors):
        if not new_colors.is_contiguous():
            raise ValueError("new_colors should be contiguous tensor \n(use new_colors.contiguous() before passing new_colors as an argument)")
        c = np.ascontiguousarray(new_colors.detach().cpu().numpy(), np.float32).reshape(-1, 4)
        ctx = _context()
        ctx.handle_of(self.data)
        ctx.set_colors(self.data, c)
        self._device_dirty = True


class MeshCache:
    """sl.MeshCache (python/src/py_mesh.cpp:516-530, src/mesh_cache.cpp): meshes by file name, so that Scene.deserialize()
    re-uses loaded meshes (and their device copies) instead of importing the file again."""

    def __init__(self):
        _context()
        self._meshes = {}

    def add(self, meshes):
        for m in meshes:
            self._meshes[m.filename] = m

    # mapping protocol used by Scene.deserialize
    def __contains__(self, filename):
        return filename in self._meshes

    def __getitem__(self, filename):
        return self._meshes[filename]

    def __setitem__(self, filename, mesh):
        self._meshes[filename] = mesh


# ---------------------------------------------------------------------------------------------
# object
# ---------------------------------------------------------------------------------------------
class Object:
    def __init__(self, mesh, options=None):
        self.mesh = mesh
        self._pose = np.eye(4, dtype=np.float32)
        self._instance_index = 0
        self.specular_color = torch.ones(4)
        self.shininess = 80.0
        self.metallic = -1.0                                           # object.h:277-278: < 0 = use the material's
        self.roughness = -1.0
        self.casts_shadows = True
        self.sticker_rotation = torch.tensor([0.0, 0.0, 0.0, 1.0])    # quaternion x,y,z,w
        self.sticker_range = torch.zeros(4)                            # min.x, min.y, max.x, max.y (Range2D)
        self.sticker_texture = None
        self.static = False
        self.mass = 1.0
        # rigid-body state: stored only (PhysX is not part of this build); py_object.cpp:120-197
        self.density = 500.0
        self.linear_velocity = torch.zeros(3)
        self.angular_velocity = torch.zeros(3)
        self.linear_velocity_limit = 1e3

    def pose(self):
        return _t(self._pose)

    def set_pose(self, pose):
        self._pose = _np44(pose)

    @property
    def instance_index(self):
        return self._instance_index

    @instance_index.setter
    def instance_index(self, index):
        if index < 0 or index > 0xFFFF:
            raise ValueError("Object::setInstanceIndex(): out of range")      # object.cpp:376-382
        self._instance_index = int(index)

    def _sticker_projection(self):
        """object.cpp:494-513: orthographic (2x/d, 2y/d, z+2, 1) of the rotated object point, d = bbox diagonal."""
        x, y, z, w = [float(v) for v in self.sticker_rotation]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float32)
        d = self.mesh.bbox.diagonal
        proj = np.eye(4, dtype=np.float32)
        proj[0, 0] = proj[1, 1] = 2.0 / d
        proj[2, 3] = 2.0
        rot = np.eye(4, dtype=np.float32)
        rot[:3, :3] = R
        return proj @ rot


# ---------------------------------------------------------------------------------------------
# scene
# ---------------------------------------------------------------------------------------------
class Scene:
    def __init__(self, viewport_size):
        _context()
        self._W, self._H = int(viewport_size[0]), int(viewport_size[1])
        self._projection = fov_projection(self._W, self._H)            # scene.cpp:138: 58 degree default
        self._camera_pose = np.eye(4, dtype=np.float32)
        self._objects = []
        self._light_directions = torch.zeros(3, 3)
        self._light_colors = torch.tensor([[300.0, 300.0, 300.0], [0, 0, 0], [0, 0, 0]])     # scene.h:225-230
        self.ambient_light = torch.zeros(3)
        self.light_map = None
        self.background_image = None
        self.background_color = torch.tensor([1.0, 1.0, 1.0, 1.0])
        self.background_plane_pose = torch.eye(4)
        self.background_plane_size = torch.zeros(2)
        self.background_plane_texture = None
        self.manual_exposure = -1.0                                    # scene.h:198: < 0 = auto exposure
        self._rng = np.random.RandomState(0)

    # ---- camera ----
    @property
    def viewport(self):
        return (self._W, self._H)

    def camera_pose(self):
        return _t(self._camera_pose)

    def set_camera_pose(self, pose):
        self._camera_pose = _np44(pose)

    def set_camera_look_at(self, position, look_at, up=(0.0, 0.0, 1.0)):
        as3 = lambda v: np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, np.float32)
        self._camera_pose = look_at_pose(as3(position), as3(look_at), as3(up))

    def set_camera_intrinsics(self, fx, fy, cx, cy):
        self._projection = intrinsics_projection(fx, fy, cx, cy, self._W, self._H)

    def set_camera_hfov(self, hfov):
        self._projection = fov_projection(self._W, self._H, math.degrees(hfov))

    def set_camera_projection(self, P):
        self._projection = _np44(P)

    def projection_matrix(self):
        return _t(self._projection)

    def min_dist_for_object_diameter(self, diameter):
        P = self._projection
        return float(max(P[0, 0] * diameter / 2.0, P[1, 1] * diameter / 2.0))      # src/pose.cpp:24-34

    # ---- objects ----
    def add_object(self, obj):
        if obj.instance_index == 0:
            obj.instance_index = len(self._objects) + 1                # scene.cpp:285-287
        self._objects.append(obj)

    def remove_object(self, obj):
        self._objects.remove(obj)

    @property
    def objects(self):
        return list(self._objects)

    def load_visual(self):
        ctx = _context()
        for o in self._objects:
            ctx.handle_of(o.mesh.data)

    # ---- lights ----
    @property
    def light_directions(self):
        return self._light_directions

    @light_directions.setter
    def light_directions(self, d):
        self._light_directions.copy_(torch.as_tensor(d, dtype=torch.float32).reshape(3, 3))

    @property
    def light_colors(self):
        return self._light_colors

    @light_colors.setter
    def light_colors(self, c):
        self._light_colors.copy_(torch.as_tensor(c, dtype=torch.float32).reshape(3, 3))

    @property
    def light_position(self):
        warnings.warn("light_position is deprecated, use light_directions instead.", DeprecationWarning)
        return -self._light_directions[0]

    @light_position.setter
    def light_position(self, p):
        warnings.warn("light_position is deprecated, use light_directions instead.", DeprecationWarning)
        self._light_directions.zero_()
        self._light_directions[0] = -torch.as_tensor(p, dtype=torch.float32)

    def choose_random_light_direction(self):
        # scene.cpp:453-470: from above and from the camera side, sampled in the camera frame
        r = np.array([self._rng.normal(), -abs(self._rng.normal()), -abs(self._rng.normal())], np.float32)
        r /= np.linalg.norm(r)
        d_world = self._camera_pose[:3, :3] @ (-r)
        self._light_directions.zero_()
        self._light_directions[0] = torch.from_numpy(d_world.astype(np.float32))

    def choose_random_light_position(self):
        warnings.warn("choose_random_light_position() is deprecated", DeprecationWarning)     # py_scene.cpp:350-352: sets nothing

    def _no_physics(self, *a, **k):
        raise RuntimeError("physics is not available in this build (PhysX stays host-side, SURVEY §3.2)")

    simulate = check_collisions = find_noncolliding_pose = load_physics = _no_physics

    def simulate_tabletop_scene(self, vis_cb=None):
        """scene.cpp:612-759 drops the objects onto a z = 0.04 table through the origin with PhysX. PhysX is not part of this
        build: this is a NON-PHYSICAL placement sampler with the same outputs (background_plane_pose set like :650-658,
        every non-static object given a pose above the table): random orientation, bounding spheres resting on the table,
        positions rejection-sampled so the spheres do not intersect. Objects do not lean on each other."""
        warnings.warn("simulate_tabletop_scene(): PhysX is not part of this build; objects are placed by a non-physical "
                      "sampler (random orientations, non-intersecting bounding spheres on the table)")
        rng = self._rng
        top = 0.04
        a = rng.uniform(-math.pi, math.pi)
        plane = np.eye(4, dtype=np.float32)
        plane[:2, :2] = [[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]]
        plane[2, 3] = top
        if all(not o.static for o in self._objects):
            self.background_plane_pose = _t(plane)
        placed = []
        for o in self._objects:
            if o.static:
                continue
            bb = o.mesh.bbox
            r = bb.diagonal / 2.0
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            R = quat_to_matrix(torch.tensor([q[0], q[1], q[2], q[3]], dtype=torch.float32)).numpy()
            spread = r
            while True:
                xy = rng.uniform(-spread, spread, size=2)
                if all(np.hypot(*(xy - pxy)) >= r + pr for pxy, pr in placed):
                    break
                spread *= 1.1
            placed.append((xy, r))
            centre = np.array([xy[0], xy[1], top + r], np.float32)
            pose = np.eye(4, dtype=np.float32)
            pose[:3, :3] = R
            pose[:3, 3] = centre - R @ bb.center.numpy()               # Matrix4::from(q, pos) * translation(-bbox centre), :676
            o.set_pose(pose)

    # ---- serialization (scene.cpp:761-868, object.cpp:384-460, mesh.cpp:1091-1115: Corrade configuration text) ----
    def serialize(self):
        f = lambda v: " ".join("%.9g" % float(x) for x in np.asarray(v, np.float64).reshape(-1))
        quat = lambda R: f(matrix_to_quat(_t(R)).numpy())
        out = ["viewport=%d %d" % (self._W, self._H), "projection=" + f(self._projection),
               "cameraPosition=" + f(self._camera_pose[:3, 3]), "cameraRotation=" + quat(self._camera_pose[:3, :3]),
               "ambientLight=" + f(self.ambient_light), "numObjects=%d" % len(self._objects)]
        if self.light_map is not None and self.light_map.path:
            out.append("lightMap=" + self.light_map.path)
        out += ["backgroundPlanePose=" + f(_np44(self.background_plane_pose)), "backgroundPlaneSize=" + f(self.background_plane_size),
                "manualExposure=%.9g" % float(self.manual_exposure)]
        for i in range(3):
            out += ["[light]", "direction=" + f(self._light_directions[i]), "color=" + f(self._light_colors[i])]
        for o in self._objects:
            out += ["[object]", "pose=" + f(o._pose), "instanceIndex=%d" % o.instance_index, "specularColor=" + f(o.specular_color),
                    "shininess=%.9g" % o.shininess, "roughness=%.9g" % float(o.roughness), "metallic=%.9g" % float(o.metallic),
                    "casts_shadows=%s" % ("true" if o.casts_shadows else "false"), "stickerRange=" + f(o.sticker_range),
                    "stickerRotation=" + f(o.sticker_rotation), "static=%s" % ("true" if o.static else "false"),
                    "density=%.9g" % float(o.density), "linear_velocity_limit=%.9g" % float(o.linear_velocity_limit),
                    "[object/mesh]", "filename=" + str(o.mesh.filename), "classIndex=%d" % o.mesh.class_index,
                    "scale=%.9g" % o.mesh._scale, "rigidPretransform=" + f(o.mesh._rigid)]
        return "\n".join(out) + "\n"

    def deserialize(self, text, cache=None):
        groups, cur = [("", {})], None
        for line in text.splitlines():
            line = line.strip()
            if not line or line.startswith(("#", ";")):
                continue
            if line.startswith("[") and line.endswith("]"):
                groups.append((line[1:-1], {}))
            elif "=" in line:
                k, v = line.split("=", 1)
                groups[-1][1][k.strip()] = v.strip()
        vec = lambda v: np.array([float(x) for x in v.split()], np.float32)
        top = groups[0][1]
        if "viewport" in top:
            self._W, self._H = (int(x) for x in top["viewport"].split())
        if "projection" in top:
            self._projection = vec(top["projection"]).reshape(4, 4)
        if "cameraPosition" in top and "cameraRotation" in top:
            pose = np.eye(4, dtype=np.float32)
            pose[:3, :3] = quat_to_matrix(torch.from_numpy(vec(top["cameraRotation"]))).numpy()
            pose[:3, 3] = vec(top["cameraPosition"])
            self._camera_pose = pose
        lights = [g for n, g in groups if n == "light"]
        if "lightPosition" in top:                                             # scene.cpp:816-820 (legacy files)
            p = vec(top["lightPosition"])
            self._light_directions.zero_()
            self._light_directions[0] = torch.from_numpy(-p / np.linalg.norm(p))
            self._light_colors.zero_()
            self._light_colors[0] = torch.tensor([0.0, 0.8, 0.0])
        elif lights:
            self._light_directions.zero_()
            self._light_colors.zero_()
            for i, g in enumerate(lights[:3]):
                self._light_directions[i] = torch.from_numpy(vec(g["direction"]))
                self._light_colors[i] = torch.from_numpy(vec(g["color"]))
        if "ambientLight" in top:
            self.ambient_light = torch.from_numpy(vec(top["ambientLight"]))
        if "lightMap" in top:
            self.light_map = LightMap(top["lightMap"])
        if "backgroundPlanePose" in top:
            self.background_plane_pose = torch.from_numpy(vec(top["backgroundPlanePose"]).reshape(4, 4))
        if "backgroundPlaneSize" in top:
            self.background_plane_size = torch.from_numpy(vec(top["backgroundPlaneSize"]))
        if "manualExposure" in top:
            self.manual_exposure = float(top["manualExposure"])
        self._objects = []
        meshes = cache if cache is not None else {}
        it = iter(range(len(groups)))
        for gi in it:
            name, g = groups[gi]
            if name != "object":
                continue
            if gi + 1 >= len(groups) or groups[gi + 1][0] != "object/mesh":
                raise RuntimeError("Did not find mesh subgroup in object")   # object.cpp:410-411
            mg = groups[gi + 1][1]
            key = mg["filename"]
            if key in meshes:
                mesh = meshes[key]                                 # mesh_cache.cpp: a cached mesh is used as it is
            else:
                mesh = Mesh(key)                                   # Mesh::deserialize (mesh.cpp:1099-1115)
                if "classIndex" in mg:
                    mesh.class_index = int(mg["classIndex"])
                if "scale" in mg:
                    mesh._scale = float(mg["scale"])
                if "rigidPretransform" in mg:
                    mesh._rigid = vec(mg["rigidPretransform"]).reshape(4, 4)
                meshes[key] = mesh                                 # (without a caller's cache: a local one, scene.cpp:852-857)
            o = Object(mesh)
            if "pose" in g:
                o._pose = vec(g["pose"]).reshape(4, 4)
            if "instanceIndex" in g:
                o.instance_index = int(g["instanceIndex"])
            if "specularColor" in g:
                o.specular_color = torch.from_numpy(vec(g["specularColor"]))
            for k, attr in (("shininess", "shininess"), ("roughness", "roughness"), ("metallic", "metallic"), ("density", "density"),
                            ("linear_velocity_limit", "linear_velocity_limit")):
                if k in g:
                    setattr(o, attr, float(g[k]))
            if "casts_shadows" in g:
                o.casts_shadows = g["casts_shadows"] == "true"
            if "static" in g:
                o.static = g["static"] == "true"
            if "stickerRange" in g:
                o.sticker_range = torch.from_numpy(vec(g["stickerRange"]))
            if "stickerRotation" in g:
                o.sticker_rotation = torch.from_numpy(vec(g["stickerRotation"]))
            self.add_object(o)

    # ---- marshalling ----
    def _spec(self, ssao_enabled, predicate):
        objs = []
        for o in self._objects:
            visible = True if predicate is None else bool(predicate(o))
            rng = [float(v) for v in o.sticker_range]
            objs.append(ObjectSpec(o.mesh.data, pose=o._pose, pretransform=o.mesh._pre_np(), class_index=o.mesh.class_index,
                                   instance_index=o.instance_index, metallic=float(o.metallic), roughness=float(o.roughness),
                                   casts_shadows=bool(o.casts_shadows), visible=visible,
                                   sticker_texture=o.sticker_texture.image if o.sticker_texture is not None else None,
                                   sticker_projection=o._sticker_projection() if o.sticker_texture is not None else np.eye(4, dtype=np.float32),
                                   sticker_range=(rng[0], rng[1], rng[2] - rng[0], rng[3] - rng[1])))
        sc = SceneSpec(self._W, self._H, self._projection, inverted_rigid(self._camera_pose), objs,
                       light_directions=self._light_directions.numpy().copy(), light_colors=self._light_colors.numpy().copy(),
                       ambient_light=tuple(float(v) for v in self.ambient_light),
                       light_map=self.light_map.data if self.light_map is not None else None,
                       background_plane_size=tuple(float(v) for v in self.background_plane_size),
                       background_plane_pose=_np44(self.background_plane_pose),
                       background_plane_texture=self.background_plane_texture.image if self.background_plane_texture is not None else None,
                       background_image=self.background_image.image if self.background_image is not None else None,
                       manual_exposure=float(self.manual_exposure), ssao_enabled=bool(ssao_enabled))
        return sc


# ---------------------------------------------------------------------------------------------
# render pass
# ---------------------------------------------------------------------------------------------
class RenderPassResult:
    """Accessors return FRESH tensors (they outlive the result and the next render), on cuda:<device> after
    init_cuda(use_cuda=True), on the CPU otherwise — same shapes and dtypes as py_render_pass.cpp:20-223."""

    def __init__(self):
        self._res = None

    def _ensure(self, W, H):
        ctx = _context()
        if self._res is None or (self._res.W, self._res.H) != (W, H):
            self._res = _lib.Result(ctx, W, H, 1, abi.TARGETS_ALL, torch_tensors=True)
        return self._res

    def _tensor(self, target):
        if self._res is None:
            raise RuntimeError("RenderPassResult is empty: render into it first")
        _context().synchronize()
        t = self._res.tensors[target][0].clone()
        return t if _use_cuda else t.cpu()

    def rgb(self):
        return self._tensor(abi.TARGET_RGB)

    def class_index(self):
        return self._tensor(abi.TARGET_CLASS)

    def instance_index(self):
        return self._tensor(abi.TARGET_INSTANCE)

    def coordinates(self):
        return self._tensor(abi.TARGET_COORD)[:, :, 0:3]

    def depth(self):
        return self._tensor(abi.TARGET_COORD)[:, :, 3]

    def coordDepth(self):
        return self._tensor(abi.TARGET_COORD)

    def normals(self):
        return self._tensor(abi.TARGET_NORMAL)

    def vertex_indices(self):
        return self._tensor(abi.TARGET_VERTEX_INDEX)[:, :, 0:3]

    def barycentric_coeffs(self):
        return self._tensor(abi.TARGET_BARY)[:, :, 0:3]

    def cam_coordinates(self):
        return self._tensor(abi.TARGET_CAM_COORD)


class RenderPass:
    def __init__(self, shading="pbr"):
        _context()
        if shading not in ("pbr", "phong", "flat"):
            raise ValueError("unknown shading type specified")        # py_render_pass.cpp:244
        self.shading = shading            # stored, unused by the reference renderer too
        self.ssao_enabled = True          # render_pass.h:150
        self._result = RenderPassResult()

    def render(self, scene, result=None, depth_peel=None, predicate=None):
        ctx = _context()
        res = result if result is not None else self._result       # a second render with result=None overwrites the first
        r = res._ensure(scene._W, scene._H)
        spec = scene._spec(self.ssao_enabled, predicate)
        ctx.render([spec], result=r, depth_peel=depth_peel._res if depth_peel is not None else None)
        return res


def view(scene):
    """The reference opens an X11 viewer; headless this returns immediately (SURVEY §3.2 caveat ii)."""
    return None


def quat_to_matrix(quat):
    """py_magnum.cpp:83-99: [x y z w] (normalised first) -> 3x3 rotation matrix."""
    if quat.dim() != 1 or quat.size(0) != 4:
        raise ValueError("Quaternion tensor should be one-dimensional tensor of size 4")
    x, y, z, w = (quat.detach().cpu().double() / quat.detach().cpu().double().norm()).tolist()
    return torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float32)


def matrix_to_quat(matrix):
    """py_magnum.cpp:100-113: 3x3 rotation matrix -> [x y z w] (Magnum's Quaternion::fromMatrix branch order)."""
    m = matrix.detach().cpu().double().numpy()
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, s / 4]
    else:
        i = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0) * 2
        q = [0.0, 0.0, 0.0, (m[k, j] - m[j, k]) / s]
        q[i], q[j], q[k] = s / 4, (m[j, i] + m[i, j]) / s, (m[k, i] + m[i, k]) / s
    return torch.tensor(q, dtype=torch.float32)


def render_debug_image(scene):
    """src/debug.cpp:19-58 draws every object's coordinate frame (DebugTools::ObjectRenderer3D: unit x / y / z axes in red /
    green / blue) over a transparent background into an RGBA8 image. A debugging aid outside the render path: the three axis
    segments are clipped at the near plane, projected with the scene's camera and drawn on the host."""
    from PIL import Image, ImageDraw
    W, H = scene._W, scene._H
    img = Image.new("RGBA", (W, H), (0, 0, 0, 0))
    draw = ImageDraw.Draw(img)
    PV = scene._projection.astype(np.float64) @ inverted_rigid(scene._camera_pose).astype(np.float64)
    near = 1e-3
    for o in scene._objects:
        M = PV @ o._pose.astype(np.float64)
        for axis, col in ((0, (255, 0, 0, 255)), (1, (0, 255, 0, 255)), (2, (0, 0, 255, 255))):
            a, b = M[:, 3].copy(), M[:, 3] + M[:, axis]
            if a[3] < near and b[3] < near:
                continue
            if a[3] < near:
                a = b + (a - b) * ((b[3] - near) / (b[3] - a[3]))
            if b[3] < near:
                b = a + (b - a) * ((a[3] - near) / (a[3] - b[3]))
            pa, pb = a[:2] / a[3], b[:2] / b[3]
            draw.line([((pa[0] + 1) * W / 2, (pa[1] + 1) * H / 2), ((pb[0] + 1) * W / 2, (pb[1] + 1) * H / 2)], fill=col, width=1)
    t = torch.from_numpy(np.asarray(img).copy())
    return t.to(f"cuda:{_cuda_index}") if _use_cuda else t


from . import diff  # noqa: E402,F401  (sl.diff.backpropagate_gradient_to_poses etc., as the reference's stillleben.diff)
