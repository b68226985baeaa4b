"""Backend-neutral scene description (plain numpy) and its marshalling into include/slb.h structs.

These classes hold exactly the state RenderPass::render() reads from Scene / Object / Mesh
(SURVEY §8b "inputs the call reads").  ``build_scene_descs`` turns them into ``slb_scene_desc``
arrays; the handle lookup is injected so the same marshalling serves the CUDA library and — in the
tests — the CPU oracle.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi


@dataclass
class ImageData:
    pixels: np.ndarray  # uint8 [H, W, 3|4], row 0 = GL row 0 (bottom row of an imported picture)
    wrap_s: int = abi.WRAP_REPEAT
    wrap_t: int = abi.WRAP_REPEAT
    min_filter: int = abi.FILTER_LINEAR_MIPMAP_LINEAR
    mag_filter: int = abi.FILTER_LINEAR
    kind: int = abi.TEXTURE_2D

    def to_c(self):
        px = np.ascontiguousarray(self.pixels, dtype=np.uint8)
        self._keep = px
        return abi.Image(px.ctypes.data, px.shape[1], px.shape[0], px.shape[2], self.wrap_s, self.wrap_t,
                         self.min_filter, self.mag_filter)


@dataclass
class MaterialData:
    base_color: tuple = (1.0, 1.0, 1.0, 1.0)
    emissive: tuple = (0.0, 0.0, 0.0, 0.0)
    metallic: float = 0.04   # defaults of RenderShader::setMaterial (render_shader.cpp:355-356)
    roughness: float = 0.5
    tex_base_color: int = -1
    tex_normal: int = -1
    tex_metallic_roughness: int = -1
    tex_emissive: int = -1
    tex_occlusion: int = -1

    def to_c(self):
        return abi.Material((C.c_float * 4)(*self.base_color), (C.c_float * 4)(*self.emissive), self.metallic,
                            self.roughness, self.tex_base_color, self.tex_normal, self.tex_metallic_roughness,
                            self.tex_emissive, self.tex_occlusion, 0)


@dataclass
class MeshData:
    vertices: np.ndarray            # structured array of abi.VERTEX_DTYPE
    indices: np.ndarray             # uint32 [n_idx]
    submeshes: List[tuple]          # (index_offset, index_count, material)
    materials: List[MaterialData] = field(default_factory=list)
    images: List[ImageData] = field(default_factory=list)
    bbox_min: Optional[np.ndarray] = None
    bbox_max: Optional[np.ndarray] = None
    name: str = ""

    def __post_init__(self):
        assert self.vertices.dtype == abi.VERTEX_DTYPE
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32)
        if self.bbox_min is None:
            self.bbox_min = self.vertices["position"].min(axis=0).astype(np.float32)
            self.bbox_max = self.vertices["position"].max(axis=0).astype(np.float32)

    @property
    def n_triangles(self):
        return sum(c // 3 for _, c, _ in self.submeshes)

    def geometry_bytes(self):
        """Algorithmic geometry bytes of one draw of this mesh: verts*68 + idx*4."""
        return len(self.vertices) * abi.VERTEX_STRIDE + sum(c for _, c, _ in self.submeshes) * 4


@dataclass
class LightMapData:
    equirect: np.ndarray  # float32 [H, W, 3], row 0 = GL row 0
    light_directions: list = field(default_factory=list)  # up to 3 (x,y,z)
    light_colors: list = field(default_factory=list)
    # precomputed (env level 0 [6,e,e,4], irradiance [6,i,i,4], prefilter packed, LUT [l,l,4]) — set on ranks that received the
    # maps in the asset broadcast instead of running the precompute themselves (dist.broadcast_assets)
    maps: Optional[tuple] = None


@dataclass
class ObjectSpec:
    mesh: MeshData
    pose: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))          # row-major m[r,c]
    pretransform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    class_index: int = 1          # Mesh::classIndex default (mesh.h:300)
    instance_index: int = 0       # 0 = auto (position in scene, 1-based; scene.cpp:285-287)
    metallic: float = -1.0
    roughness: float = -1.0
    casts_shadows: bool = True
    visible: bool = True
    sticker_texture: Optional[ImageData] = None
    sticker_projection: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    sticker_range: tuple = (0.0, 0.0, 0.0, 0.0)


@dataclass
class SceneSpec:
    width: int
    height: int
    projection: np.ndarray
    world_to_cam: np.ndarray
    objects: List[ObjectSpec] = field(default_factory=list)
    # Scene defaults: colours {(300,300,300),0,0}, all directions 0 => no light (scene.h:225-230)
    light_directions: np.ndarray = field(default_factory=lambda: np.zeros((3, 3), np.float32))
    light_colors: np.ndarray = field(
        default_factory=lambda: np.array([[300, 300, 300], [0, 0, 0], [0, 0, 0]], np.float32))
    ambient_light: tuple = (0.0, 0.0, 0.0)
    light_map: Optional[LightMapData] = None
    background_plane_size: tuple = (0.0, 0.0)
    background_plane_pose: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    background_plane_texture: Optional[ImageData] = None
    background_image: Optional[ImageData] = None
    manual_exposure: float = -1.0
    ssao_enabled: bool = True


def intrinsics_projection(fx, fy, cx, cy, W, H, n=0.1, f=10.0):
    """Scene::setCameraIntrinsics (reference: src/scene.cpp:222-253), returned row-major m[r,c]."""
    f32 = np.float32
    fx, fy, cx, cy, W, H, n, f = map(f32, (fx, fy, cx, cy, W, H, n, f))
    L = -cx * n / fx
    R = (W - cx) * n / fx
    T = -cy * n / fy
    B = (H - cy) * n / fy
    cols = [
        [f32(2.0) * n / (R - L), 0, 0, 0],
        [0, f32(2.0) * n / (B - T), 0, 0],
        [(R + L) / (L - R), (T + B) / (T - B), (f + n) / (f - n), 1],
        [0, 0, (f32(2.0) * f * n) / (n - f), 0],
    ]
    return np.array(cols, dtype=np.float32).T.copy()


def fov_projection(W, H, fov_deg=58.0):
    """Scene::setCameraFromFOV (reference: src/scene.cpp:260-271); default 58 deg (scene.cpp:138)."""
    fx = W / (2.0 * np.tan(np.deg2rad(fov_deg) / 2.0))
    return intrinsics_projection(fx, fx, W / 2, H / 2, W, H)


def look_at_pose(position, look_at, up=(0, 0, 1)):
    """Scene::setCameraLookAt (reference: src/scene.cpp:205-215): camera pose (camera -> world)."""
    position = np.asarray(position, np.float32)
    z = np.asarray(look_at, np.float32) - position
    z = z / np.linalg.norm(z)
    x = np.cross(z, np.asarray(up, np.float32))
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    y = y / np.linalg.norm(y)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, position
    return m


def inverted_rigid(m):
    r = np.eye(4, dtype=np.float32)
    r[:3, :3] = m[:3, :3].T
    r[:3, 3] = -(m[:3, :3].T @ m[:3, 3])
    return r


class DescBatch:
    """Owns the ctypes memory of an array of slb_scene_desc (keeps everything alive)."""

    def __init__(self, scenes, handle_of):
        self.n = len(scenes)
        self.scenes = (abi.SceneDesc * self.n)()
        self._objs = []
        for i, sc in enumerate(scenes):
            d = self.scenes[i]
            d.width, d.height = sc.width, sc.height
            d.projection = abi.mat4_to_c(sc.projection)
            d.world_to_cam = abi.mat4_to_c(sc.world_to_cam)
            ld = np.asarray(sc.light_directions, np.float32).reshape(3, 3)
            lc = np.asarray(sc.light_colors, np.float32).reshape(3, 3)
            for k in range(3):
                for j in range(3):
                    d.light_directions[k][j] = float(ld[k, j])
                    d.light_colors[k][j] = float(lc[k, j])
            for j in range(3):
                d.ambient_light[j] = float(sc.ambient_light[j])
            d.light_map = handle_of(sc.light_map) if sc.light_map is not None else None
            d.background_plane_size[0] = float(sc.background_plane_size[0])
            d.background_plane_size[1] = float(sc.background_plane_size[1])
            d.background_plane_pose = abi.mat4_to_c(sc.background_plane_pose)
            d.background_plane_texture = (handle_of(sc.background_plane_texture)
                                          if sc.background_plane_texture is not None else None)
            d.background_image = handle_of(sc.background_image) if sc.background_image is not None else None
            d.manual_exposure = float(sc.manual_exposure)
            d.ssao_enabled = 1 if sc.ssao_enabled else 0
            objs = (abi.ObjectDesc * max(1, len(sc.objects)))()
            for k, o in enumerate(sc.objects):
                od = objs[k]
                od.mesh = handle_of(o.mesh)
                od.pose = abi.mat4_to_c(o.pose)
                od.pretransform = abi.mat4_to_c(o.pretransform)
                od.class_index = int(o.class_index)
                od.instance_index = int(o.instance_index) if o.instance_index else k + 1
                od.metallic = float(o.metallic)
                od.roughness = float(o.roughness)
                od.casts_shadows = 1 if o.casts_shadows else 0
                od.visible = 1 if o.visible else 0
                od.sticker_texture = handle_of(o.sticker_texture) if o.sticker_texture is not None else None
                od.sticker_projection = abi.mat4_to_c(o.sticker_projection)
                for j in range(4):
                    od.sticker_range[j] = float(o.sticker_range[j])
            self._objs.append(objs)
            d.objects = C.cast(objs, C.POINTER(abi.ObjectDesc))
            d.n_objects = len(sc.objects)

    @property
    def ptr(self):
        return C.cast(self.scenes, C.POINTER(abi.SceneDesc))
