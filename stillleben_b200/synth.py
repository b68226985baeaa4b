"""Procedural stand-in assets and the synthetic scenes of BASELINE.json's configs.

YCB meshes, sIBL light maps and the reference's test assets are not available on the GPU box, so
the benchmark and the parity tests use deterministic procedural stand-ins: a unit cube
equal to tests/cube.glb (24 vertices, 12 triangles, per-face normals, baseColor 0.8, metallic 0,
roughness 0.4), a pool of ~16 k-triangle parametric shapes (YCB google_16k scale) with procedural
textures, and a procedural sky + sun equirect.  Everything is seeded and numpy-only.
"""
import numpy as np

from . import abi
from .desc import (ImageData, LightMapData, MaterialData, MeshData, ObjectSpec, SceneSpec, fov_projection,
                   intrinsics_projection, inverted_rigid, look_at_pose)


# ------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------
def _pack(pos, nrm, uv, tan, indices, material, images=(), name=""):
    n = len(pos)
    v = np.zeros(n, dtype=abi.VERTEX_DTYPE)
    v["position"] = pos
    v["uv"] = uv
    v["color"] = (1.0, 1.0, 1.0, 1.0)
    v["tangent"][:, :3] = tan
    v["tangent"][:, 3] = 1.0            # consolidate.cpp:278 forces tangent.w = 1
    v["vertex_index"] = np.arange(1, n + 1, dtype=np.uint32)   # one-based (consolidate.cpp:333-335)
    v["normal"] = nrm
    idx = np.asarray(indices, np.uint32).reshape(-1)
    return MeshData(v, idx, [(0, len(idx), 0)], [material], list(images), name=name)


def cube_mesh():
    """Blender default cube (+-1) as in the reference's tests/cube.glb: 24 verts / 36 indices."""
    faces = [  # (normal, u axis, v axis)
        ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, -1, 0), (0, 0, 1)),
        ((0, 1, 0), (-1, 0, 0), (0, 0, 1)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
        ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (1, 0, 0), (0, -1, 0)),
    ]
    pos, nrm, uv, tan, idx = [], [], [], [], []
    for n, a, b in faces:
        n, a, b = map(lambda t: np.array(t, np.float32), (n, a, b))
        base = len(pos)
        for (s, t) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(n + s * a + t * b)
            nrm.append(n)
            uv.append(((s + 1) / 2, (t + 1) / 2))
            tan.append(a)
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    mat = MaterialData(base_color=(0.8, 0.8, 0.8, 1.0), metallic=0.0, roughness=0.4)
    return _pack(np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32),
                 np.array(tan, np.float32), idx, mat, name="cube")


def _grid_mesh(P, nu, nv, material, images, name, wrap_u=True):
    """Parametric surface P[(nv+1), (nu+1), 3] -> mesh with smooth normals and d/du tangents."""
    H, Wd = nv + 1, nu + 1
    pos = P.reshape(-1, 3).astype(np.float32)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    a = (jj * Wd + ii).reshape(-1)
    b = a + 1
    c = a + Wd
    d = c + 1
    idx = np.stack([a, b, d, a, d, c], axis=1).reshape(-1).astype(np.uint32)
    # area-weighted smooth normals; seam vertices are welded by position for the accumulation
    tri = idx.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    nrm = np.zeros_like(pos)
    for k in range(3):
        np.add.at(nrm, tri[:, k], fn)
    if wrap_u:
        N = nrm.reshape(H, Wd, 3)
        s = N[:, 0] + N[:, -1]
        N[:, 0] = s
        N[:, -1] = s
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(ln > 1e-20, nrm / np.maximum(ln, 1e-20), np.array([0, 0, 1], np.float32))
    dPu = np.gradient(P, axis=1).reshape(-1, 3)
    tan = dPu - nrm * np.sum(dPu * nrm, axis=1, keepdims=True)
    lt = np.linalg.norm(tan, axis=1, keepdims=True)
    tan = np.where(lt > 1e-20, tan / np.maximum(lt, 1e-20), np.array([1, 0, 0], np.float32))
    uu, vv = np.meshgrid(np.linspace(0, 1, Wd), np.linspace(0, 1, H))
    uv = np.stack([uu, vv], axis=-1).reshape(-1, 2).astype(np.float32)
    return _pack(pos, nrm.astype(np.float32), uv, tan.astype(np.float32), idx, material, images, name=name)


def procedural_texture(seed, size=512):
    """Deterministic RGB uint8 [size,size,3] pattern (stripes x checker x low-frequency noise)."""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    base = rng.uniform(0.15, 0.95, size=3).astype(np.float32)
    alt = rng.uniform(0.05, 0.95, size=3).astype(np.float32)
    k = rng.randint(3, 12)
    pat = ((np.floor(x * k) + np.floor(y * k)) % 2).astype(np.float32)
    stripes = 0.5 + 0.5 * np.sin(2 * np.pi * (x * rng.randint(2, 9) + y * rng.randint(0, 5)))
    m = np.clip(0.6 * pat + 0.4 * stripes, 0, 1)[..., None]
    img = base * m + alt * (1 - m)
    img *= (0.85 + 0.15 * np.sin(2 * np.pi * (3 * x + 2 * y)))[..., None]
    return np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8)


def shape_mesh(kind, seed, nu=128, nv=64, textured=True, tex_size=512):
    """One of the stand-in shapes; (nu+1)*(nv+1) = 8385 vertices, 2*nu*nv = 16384 triangles."""
    rng = np.random.RandomState(seed)
    u = np.linspace(0, 2 * np.pi, nu + 1, dtype=np.float64)
    v = np.linspace(0, 1, nv + 1, dtype=np.float64)
    U, V = np.meshgrid(u, v)
    if kind == "blob":
        th = V * np.pi
        r = 1.0 + sum(rng.uniform(0.03, 0.12) * np.sin(rng.randint(1, 5) * U + rng.uniform(0, 6)) *
                      np.sin(rng.randint(1, 4) * th + rng.uniform(0, 6)) for _ in range(4))
        P = np.stack([r * np.sin(th) * np.cos(U), r * np.sin(th) * np.sin(U), r * np.cos(th)], -1)
    elif kind == "torus":
        R, r = 1.0, rng.uniform(0.25, 0.45)
        ph = V * 2 * np.pi
        P = np.stack([(R + r * np.cos(ph)) * np.cos(U), (R + r * np.cos(ph)) * np.sin(U), r * np.sin(ph)], -1)
    elif kind == "can":
        # closed cylinder: profile radius over v with flat caps
        hgt = rng.uniform(1.2, 2.4)
        t = V
        rad = np.where(t < 0.15, t / 0.15, np.where(t > 0.85, (1 - t) / 0.15, 1.0))
        z = np.where(t < 0.15, 0.0, np.where(t > 0.85, 1.0, (t - 0.15) / 0.7)) * hgt - hgt / 2
        P = np.stack([rad * np.cos(U), rad * np.sin(U), z], -1)
    elif kind == "box":
        # superellipsoid (rounded box)
        e = rng.uniform(0.2, 0.5)
        th = V * np.pi - np.pi / 2
        sx, sy, sz = rng.uniform(0.6, 1.4, size=3)
        spow = lambda w, m: np.sign(w) * np.abs(w) ** m
        P = np.stack([sx * spow(np.cos(th), e) * spow(np.cos(U), e), sy * spow(np.cos(th), e) * spow(np.sin(U), e),
                      sz * spow(np.sin(th), e)], -1)
    elif kind == "bottle":
        t = V
        rad = 0.5 + 0.35 * np.cos(np.pi * np.clip((t - 0.55) / 0.3, 0, 1)) * (t > 0.55) + 0.35 * (t <= 0.55)
        rad = np.where(t < 0.05, rad * t / 0.05, np.where(t > 0.97, rad * (1 - t) / 0.03, rad))
        P = np.stack([rad * np.cos(U), rad * np.sin(U), (t - 0.5) * 2.6], -1)
    else:
        raise ValueError(kind)
    images = []
    if textured:
        images.append(ImageData(procedural_texture(seed * 7 + 1, tex_size)))
        mat = MaterialData(base_color=(1, 1, 1, 1), metallic=0.0, roughness=0.6, tex_base_color=0)
    else:
        c = rng.uniform(0.1, 0.9, size=3)
        mat = MaterialData(base_color=(float(c[0]), float(c[1]), float(c[2]), 1.0), metallic=float(rng.uniform(0, 1)),
                           roughness=float(rng.uniform(0.2, 0.9)))
    return _grid_mesh(P.astype(np.float32), nu, nv, mat, images, f"{kind}{seed}")


SHAPE_KINDS = ["blob", "torus", "can", "box", "bottle"]


def mesh_pool(n=21, nu=128, nv=64, tex_size=512):
    """Pool of n stand-in meshes shared by all scenes of a batch."""
    return [shape_mesh(SHAPE_KINDS[i % len(SHAPE_KINDS)], 100 + i, nu, nv, textured=(i % 3 != 2), tex_size=tex_size)
            for i in range(n)]


def normalising_pretransform(mesh, diagonal):
    """Mesh::centerBBox + scaleToBBoxDiagonal (reference: src/mesh.cpp:1000-1045): scaling(s) * rigid."""
    c = (mesh.bbox_min + mesh.bbox_max) / 2
    d = float(np.linalg.norm(mesh.bbox_max - mesh.bbox_min))
    s = diagonal / d
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] *= s
    m[:3, 3] = -c * s
    return m


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float32)


def procedural_equirect(width=1024, height=512, seed=7):
    """Sky gradient + Gaussian sun, float32 [H, W, 3], row 0 = bottom (GL addressing)."""
    rng = np.random.RandomState(seed)
    v, u = np.mgrid[0:height, 0:width].astype(np.float32)
    u = (u + 0.5) / width
    v = (v + 0.5) / height
    elev = (v - 0.5) * np.pi                      # asin(z): -pi/2 (bottom) .. pi/2 (top)
    sky = np.stack([0.25 + 0.35 * (1 - np.clip(elev, 0, 2)), 0.45 + 0.25 * (1 - np.clip(elev, 0, 2)),
                    0.9 * np.ones_like(elev)], -1)
    ground = np.stack([0.22 * np.ones_like(elev), 0.2 * np.ones_like(elev), 0.17 * np.ones_like(elev)], -1)
    img = np.where((elev > 0)[..., None], sky, ground).astype(np.float32)
    su, sv = rng.uniform(0.1, 0.9), rng.uniform(0.62, 0.8)
    du = np.minimum(np.abs(u - su), 1 - np.abs(u - su))
    sun = 60.0 * np.exp(-((du * 2) ** 2 + (v - sv) ** 2) / (2 * 0.012 ** 2))
    img += sun[..., None] * np.array([1.0, 0.93, 0.8], np.float32)
    # light direction as LightMap::load derives it from (u,v) (light_map.cpp:314-326)
    theta = (su + 0.5) * 2 * np.pi
    phi = (1.0 - sv) * np.pi
    pos = np.array([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], np.float32)
    return img.astype(np.float32), -pos


# ------------------------------------------------------------------------------------------
# scenes of the configs
# ------------------------------------------------------------------------------------------
def config1_scene(size=128, cube=None):
    """C1: cube.glb stand-in, 128x128, look-at (4,0,0)->(0,0,0), one light, exposure 1, SSAO off."""
    cube = cube or cube_mesh()
    pose = look_at_pose((4, 0, 0), (0, 0, 0))
    ldir = np.zeros((3, 3), np.float32)
    ldir[0] = np.array([-1, -1, -1], np.float32) / np.sqrt(3.0)
    lcol = np.zeros((3, 3), np.float32)
    lcol[0] = 3.0
    return SceneSpec(size, size, fov_projection(size, size), inverted_rigid(pose), [ObjectSpec(cube, instance_index=1)],
                     light_directions=ldir, light_colors=lcol, manual_exposure=1.0, ssao_enabled=False)


YCB_INTRINSICS = (1066.778, 1067.487, 312.9869, 241.3109)   # examples/ycb.py:32


def tabletop_scene(pool, seed, n_objects=20, width=640, height=480, light_map=None, ssao=False, manual_exposure=1.0,
                   n_lights=1, plane=True, intrinsics=YCB_INTRINSICS):
    """Random table-top heap stand-in (no PhysX): objects scattered above a 3x3 m plane at z=0, camera on
    a ring looking at the heap.  scene s of config C3 uses seed 1000+s."""
    rng = np.random.RandomState(seed)
    az = rng.uniform(0, 2 * np.pi)
    dist = rng.uniform(0.9, 1.25)
    cam_pos = np.array([dist * np.cos(az), dist * np.sin(az), rng.uniform(0.55, 0.9)], np.float32)
    pose = look_at_pose(cam_pos, (0.0, 0.0, 0.08))
    if intrinsics is not None and (width, height) == (640, 480):
        proj = intrinsics_projection(*intrinsics, width, height)
    else:
        proj = fov_projection(width, height)
    objects = []
    for k in range(n_objects):
        mesh = pool[rng.randint(len(pool))]
        diag = rng.uniform(0.08, 0.30)
        r = 0.42 * np.sqrt(rng.uniform())
        a = rng.uniform(0, 2 * np.pi)
        p = np.eye(4, dtype=np.float32)
        p[:3, :3] = random_rotation(rng)
        p[:3, 3] = (r * np.cos(a), r * np.sin(a), diag * 0.5 + rng.uniform(0.0, 0.18))
        objects.append(ObjectSpec(mesh, pose=p, pretransform=normalising_pretransform(mesh, diag), class_index=1 + (k % 21),
                                  instance_index=k + 1, metallic=float(rng.uniform(0, 1)), roughness=float(rng.uniform(0, 1))))
    ldir = np.zeros((3, 3), np.float32)
    lcol = np.zeros((3, 3), np.float32)
    for i in range(n_lights):
        d = np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), -1.0], np.float32)
        ldir[i] = d / np.linalg.norm(d)
        lcol[i] = rng.uniform(1.5, 4.0)
    sc = SceneSpec(width, height, proj, inverted_rigid(pose), objects, light_directions=ldir, light_colors=lcol,
                   ambient_light=(0.25, 0.25, 0.25), light_map=light_map, manual_exposure=manual_exposure,
                   ssao_enabled=ssao)
    if plane:
        sc.background_plane_size = (3.0, 3.0)
    return sc
