"""Scenes shared by the CPU and GPU tests, built from the committed fixtures in tests/golden/."""
import copy
import os

import numpy as np

from stillleben_b200 import abi, synth
from stillleben_b200.desc import (ImageData, MaterialData, MeshData, ObjectSpec, SceneSpec, fov_projection, inverted_rigid,
                                  look_at_pose)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load_mesh(name):
    if name in _cache:
        return _cache[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    verts = np.ascontiguousarray(z["vertices"]).view(abi.VERTEX_DTYPE).reshape(-1)
    mats = []
    for r in z["materials"]:
        mats.append(MaterialData(tuple(float(x) for x in r[0:4]), tuple(float(x) for x in r[4:8]), float(r[8]), float(r[9]),
                                 int(r[10]), int(r[11]), int(r[12]), int(r[13]), int(r[14])))
    images = []
    i = 0
    while f"image{i}" in z:
        s = z[f"image{i}_sampler"]
        images.append(ImageData(np.ascontiguousarray(z[f"image{i}"]), int(s[0]), int(s[1]), int(s[2]), int(s[3])))
        i += 1
    m = MeshData(verts, z["indices"], [tuple(int(x) for x in s) for s in z["submeshes"]], mats, images,
                 z["bbox_min"].astype(np.float32), z["bbox_max"].astype(np.float32), name)
    _cache[name] = m
    return m


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def cube_test_scene(width=640, height=480):
    """The reference's "vertex indices" test (tests/basic.cpp:375-453): cube.glb at the origin, 640x480,
    camera at (4,0,0) looking at the origin; a light direction is set (chooseRandomLightDirection there,
    a fixed one here) so the colour target is not black."""
    cube = load_mesh("cube_glb_mesh")
    pose = look_at_pose((4, 0, 0), (0, 0, 0))
    ldir = np.zeros((3, 3), np.float32)
    ldir[0] = np.array([-1, -1, -1], np.float32) / np.sqrt(3.0)
    lcol = np.zeros((3, 3), np.float32)
    lcol[0] = 3.0
    return SceneSpec(width, height, fov_projection(width, height), inverted_rigid(pose), [ObjectSpec(cube, instance_index=1)],
                     light_directions=ldir, light_colors=lcol, manual_exposure=1.0, ssao_enabled=False)


def bunny_test_scene(width=640, height=480, lit=False):
    """The reference's "render" test (tests/basic.cpp:108-261): bunny centred, scaled to diagonal 0.5,
    placed at (0,0,minimumDistanceForObjectDiameter), instance index 0xFFFF, default (identity) camera
    pose, default lights (none active), default SSAO on, auto exposure."""
    bunny = load_mesh("bunny_mesh")
    pre = synth.normalising_pretransform(bunny, 0.5)
    P = fov_projection(width, height)
    diag = 0.5
    distance = max(P[0, 0] * diag / 2.0, P[1, 1] * diag / 2.0)     # src/pose.cpp:24-34
    pose = np.eye(4, dtype=np.float32)
    pose[2, 3] = distance
    sc = SceneSpec(width, height, P, np.eye(4, dtype=np.float32),
                   [ObjectSpec(bunny, pose=pose, pretransform=pre, class_index=1, instance_index=0xFFFF)],
                   manual_exposure=-1.0, ssao_enabled=True)
    if lit:     # not part of the reference test: a light so that the colour target is not black
        sc.light_directions[0] = np.array([0.3, 0.5, 0.8], np.float32) / np.linalg.norm([0.3, 0.5, 0.8])
        sc.light_colors[0] = 3.0
        sc.ambient_light = (0.2, 0.2, 0.2)
    return sc


def small_tabletop_scene():
    pool = synth.mesh_pool(5, nu=32, nv=16, tex_size=64)
    return synth.tabletop_scene(pool, 4242, n_objects=6, width=160, height=120, intrinsics=None)


# ---------------------------------------------------------------------------------------------
# scene variants covering the edge cases of the path (used by the GPU parity tests)
# ---------------------------------------------------------------------------------------------
def small_pool():
    if "pool" not in _cache:
        _cache["pool"] = synth.mesh_pool(6, nu=48, nv=24, tex_size=64)
    return _cache["pool"]


def alpha_mesh():
    """A textured blob whose RGBA texture has transparent holes -> exercises the alpha-test discard."""
    if "alpha" not in _cache:
        m = synth.shape_mesh("blob", 77, nu=48, nv=24, textured=True, tex_size=64)
        rgb = m.images[0].pixels
        yy, xx = np.mgrid[0:64, 0:64]
        a = np.where(((xx // 8) + (yy // 8)) % 2 == 0, 255, 30).astype(np.uint8)
        m.images[0] = ImageData(np.ascontiguousarray(np.dstack([rgb, a])))
        _cache["alpha"] = m
    return _cache["alpha"]


def light_map_data():
    from stillleben_b200.desc import LightMapData
    if "lm" not in _cache:
        eq, sun_dir = synth.procedural_equirect(128, 64, seed=7)
        _cache["lm"] = LightMapData(eq, [sun_dir.tolist()], [[2.0, 1.9, 1.7]])
    return _cache["lm"]


def variant(name, width=320, height=240):
    pool = small_pool()
    base = dict(n_objects=6, width=width, height=height, intrinsics=None)
    if name == "tabletop":
        return synth.tabletop_scene(pool, 11, **base)
    if name == "three_lights":
        return synth.tabletop_scene(pool, 12, n_lights=3, **base)
    if name == "ssao":
        return synth.tabletop_scene(pool, 13, ssao=True, **base)
    if name == "auto_exposure":
        return synth.tabletop_scene(pool, 14, manual_exposure=-1.0, **base)
    if name == "no_plane_no_light":
        sc = synth.tabletop_scene(pool, 15, plane=False, n_lights=0, **base)
        return sc
    if name == "empty":
        return synth.tabletop_scene(pool, 16, n_objects=0, width=width, height=height, intrinsics=None, plane=False)
    if name == "ibl":
        return synth.tabletop_scene(pool, 17, light_map=light_map_data(), ssao=True, **base)
    if name == "alpha_test":
        sc = synth.tabletop_scene(pool, 18, **base)
        m = alpha_mesh()
        for o in sc.objects[:3]:
            o.mesh = m
            o.pretransform = synth.normalising_pretransform(m, 0.3)
        return sc
    if name == "sticker":
        sc = synth.tabletop_scene(pool, 19, **base)
        st = ImageData(synth.procedural_texture(5, 32), kind=abi.TEXTURE_RECT)
        o = sc.objects[0]
        o.sticker_texture = st
        d = float(np.linalg.norm(o.mesh.bbox_max - o.mesh.bbox_min)) * float(o.pretransform[0, 0])
        proj = np.eye(4, dtype=np.float32)            # orthographic: (2x/d, 2y/d, z+2, 1)  
        proj[0, 0] = proj[1, 1] = 2.0 / d
        proj[2, 3] = 2.0
        o.sticker_projection = proj
        o.sticker_range = (-0.5, -0.5, 1.0, 1.0)
        return sc
    if name == "background_image":
        sc = synth.tabletop_scene(pool, 20, plane=False, **base)
        sc.background_image = ImageData(synth.procedural_texture(9, 96), kind=abi.TEXTURE_RECT, mag_filter=abi.FILTER_NEAREST)
        return sc
    if name == "plane_texture":
        sc = synth.tabletop_scene(pool, 21, **base)
        sc.background_plane_texture = ImageData(synth.procedural_texture(3, 128))
        return sc
    if name == "near_clip":
        # camera 6 cm above the 3x3 m plane looking along it: the plane crosses the near plane and
        # extends behind the camera -> homogeneous clipping + guard band
        sc = synth.tabletop_scene(pool, 22, **base)
        pose = look_at_pose((0.9, 0.1, 0.06), (0.0, 0.0, 0.10))
        sc.world_to_cam = inverted_rigid(pose)
        return sc
    if name == "predicate":
        sc = synth.tabletop_scene(pool, 23, **base)
        sc.objects[1].visible = False
        sc.objects[2].casts_shadows = False
        return sc
    if name == "id_limits":
        sc = synth.tabletop_scene(pool, 24, **base)
        sc.objects[0].instance_index = 65535
        sc.objects[0].class_index = 65535
        return sc
    if name == "multi_submesh":
        # one mesh drawn as two sub-meshes with different materials (per-sub-mesh draws, shared vertex buffer); one of
        # its instances does not cast shadows, another object is hidden: exercises the draw / shadow-draw bookkeeping
        sc = synth.tabletop_scene(pool, 26, n_lights=2, **base)
        src = sc.objects[0].mesh
        n_idx = len(src.indices)
        cut = (n_idx // 6) * 3
        mats = [src.materials[0], MaterialData(base_color=(0.9, 0.2, 0.1, 1.0), metallic=0.3, roughness=0.7)]
        two = MeshData(src.vertices, src.indices, [(0, cut, 0), (cut, n_idx - cut, 1)], mats, list(src.images), name="two_materials")
        for k in (0, 2, 4):
            sc.objects[k].mesh = two
            sc.objects[k].pretransform = synth.normalising_pretransform(two, 0.25)
        sc.objects[2].casts_shadows = False
        sc.objects[3].visible = False
        return sc
    if name == "low_poly_closeup":
        # a 64-triangle blob filling the view: several dozen huge triangles per frame -> the first 16 are resolved in the
        # shade kernel, the rest fall back to the tiled path, a few are medium / small (warp and thread paths)
        m = synth.shape_mesh("blob", 5, nu=8, nv=4, textured=True, tex_size=64)
        sc = synth.tabletop_scene(pool, 27, n_objects=2, width=width, height=height, intrinsics=None)
        for k, o in enumerate(sc.objects):
            o.mesh = m
            o.pretransform = synth.normalising_pretransform(m, 1.1 - 0.5 * k)
            o.pose = np.eye(4, dtype=np.float32)
            o.pose[:3, 3] = (0.15 * k, -0.1 * k, 0.35 + 0.2 * k)
        return sc
    if name == "pbr_textures":
        # normal map + tangent frame, metallic-roughness, emissive and occlusion textures (render_shader.frag:259-298) on the
        # reference-consolidated assets: pbr_patch (all five textures, computed tangents) and kitchen_sink (four sub-meshes,
        # nested node transforms, an RGBA base colour -> alpha test, a sub-mesh without material, non-default samplers)
        sc = synth.tabletop_scene(pool, 28, n_lights=2, **base)
        for k, mesh_name in ((0, "pbr_patch_mesh"), (1, "kitchen_sink_mesh"), (2, "pbr_patch_mesh"), (3, "kitchen_sink_mesh")):
            m = load_mesh(mesh_name)
            o = sc.objects[k]
            o.mesh = m
            o.pretransform = synth.normalising_pretransform(m, 0.35 if mesh_name == "pbr_patch_mesh" else 0.9)
            o.pose = np.array(o.pose, np.float32)
            o.pose[2, 3] += 0.12                        # lift the open patches off the table so both faces are seen
            o.metallic = o.roughness = -1.0             # material factors x textures (no per-object override)
        sc.objects[2].metallic, sc.objects[2].roughness = 0.9, 0.3      # per-object overrides on top of the textures
        return sc
    if name == "pbr_textures_ibl":
        sc = variant("pbr_textures", width, height)
        sc.light_map = light_map_data()
        sc.ssao_enabled = True
        return sc
    if name == "projective":
        # non-affine transformation chains: a pretransform and a pose with a projective last row, and a general projection
        # matrix (Scene::setCameraProjection accepts any 4x4, scene.cpp:255-258) -> the per-vertex vertex-stage path
        sc = synth.tabletop_scene(pool, 29, **base)
        for k in (0, 2):
            pre = np.array(sc.objects[k].pretransform, np.float32)
            pre[3, :3] = (0.15, -0.1, 0.2)
            sc.objects[k].pretransform = pre
        pose = np.array(sc.objects[1].pose, np.float32)
        pose[3, :3] = (0.02, 0.03, -0.02)
        sc.objects[1].pose = pose
        P = np.array(sc.projection, np.float32)
        P[0, 1] = 0.05; P[1, 0] = -0.03                 # skew
        sc.projection = P
        return sc
    if name == "c2_shape":
        # config C2 at its real shape: 10 objects incl. the bunny, 640x480, the ycb.py intrinsics, IBL + SSAO +
        # auto exposure, metallic / roughness ~ U(0,1)
        rng = np.random.RandomState(1234)
        sc = synth.tabletop_scene(pool, 30, n_objects=10, width=640, height=480, intrinsics=(1066.778, 1067.487, 312.9869, 241.3109),
                                  light_map=light_map_data(), ssao=True, manual_exposure=-1.0)
        bunny = load_mesh("bunny_mesh")
        for k, o in enumerate(sc.objects):
            if k % 3 == 0:
                o.mesh = bunny
                o.pretransform = synth.normalising_pretransform(bunny, float(rng.uniform(0.08, 0.30)))
            o.metallic, o.roughness = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
        return sc
    if name == "c5_shape":
        # config C5's frame: 1920x1080, 64 objects, IBL + SSAO + 3 shadow lights
        return synth.tabletop_scene(pool, 31, n_objects=64, width=1920, height=1080, intrinsics=None, n_lights=3, ssao=True,
                                    light_map=light_map_data())
    if name == "odd_viewport":
        return synth.tabletop_scene(pool, 25, n_objects=5, width=203, height=117, intrinsics=None)
    raise KeyError(name)


def gl_scene_of(name):
    """The variant as the OpenGL reference harness takes it (tests/test_gl_ref.py, tests/golden/make_gl_golden.py)."""
    sc = variant(name)
    if name == "projective":
        sc.objects = [copy.copy(o) for o in sc.objects]
        pose = np.array(sc.objects[1].pose, np.float32)
        pose[3, :3] = 0.0
        sc.objects[1].pose = pose
    return sc


def single_level_copy(sc):
    """The scene with every material / plane texture minified by GL_LINEAR (level 0 only): takes LOD selection out of the comparison."""
    sc = copy.copy(sc)
    sc.objects = [copy.copy(o) for o in sc.objects]
    memo = {}
    for o in sc.objects:
        if id(o.mesh) not in memo:
            m = copy.copy(o.mesh)
            m.images = [copy.copy(im) for im in m.images]
            for im in m.images:
                im.min_filter = abi.FILTER_LINEAR
            memo[id(o.mesh)] = m
        o.mesh = memo[id(o.mesh)]
    if sc.background_plane_texture is not None:
        sc.background_plane_texture = copy.copy(sc.background_plane_texture)
        sc.background_plane_texture.min_filter = abi.FILTER_LINEAR
    return sc


def gl_post_scene(name):
    """The two post-pass scenes of the OpenGL goldens (tests/golden/make_gl_golden.py, tests/test_gl_golden.py): 'ssao' = the ssao variant with
    level-0 texture filtering; 'ibl' = the ibl variant (SSAO on) with level-0 filtering and roughness in quarters (whole prefilter levels)."""
    sc = single_level_copy(variant(name))
    if name == "ibl":
        for k, ob in enumerate(sc.objects):
            ob.roughness = (0.25, 0.5, 0.75, 1.0)[k % 4]
        sc.light_map = copy.copy(sc.light_map)
    return sc


def fuzz_scene(seed):
    """Random corner cases of the fixed-function stage: viewports from 1x1 to 333x77, 1 - 8 objects, 0 - 3 shadow lights, with / without the
    plane, hidden objects / non-casters, and five camera set-ups — inside the heap (geometry crossing the near plane and behind the camera),
    a 100 - 150 degree field of view, 6 m away (sub-pixel triangles), grazing 1 - 5 cm above the plane, and the generator's own."""
    from stillleben_b200.desc import fov_projection
    rng = np.random.RandomState(seed)
    W, H = [(320, 240), (64, 48), (17, 13), (200, 120), (3, 2), (1, 1), (333, 77)][rng.randint(7)]
    sc = synth.tabletop_scene(small_pool(), seed, n_objects=int(rng.randint(1, 9)), width=W, height=H, intrinsics=None,
                              n_lights=int(rng.randint(0, 4)), plane=bool(rng.randint(2)))
    mode = rng.randint(5)
    if mode == 0:
        pos = rng.uniform(-0.3, 0.3, 3) * [1, 1, 0.3] + [0, 0, 0.15]
        sc.world_to_cam = inverted_rigid(look_at_pose(pos, rng.uniform(-0.3, 0.3, 3) + [0, 0, 0.1]))
    elif mode == 1:
        sc.projection = fov_projection(W, H, float(rng.uniform(100, 150)))
    elif mode == 2:
        sc.world_to_cam = inverted_rigid(look_at_pose(np.array([5.0 * np.cos(seed), 5.0 * np.sin(seed), 3.0]), (0, 0, 0.1)))
    elif mode == 3:
        pos = np.array([rng.uniform(0.5, 1.2), rng.uniform(-0.3, 0.3), rng.uniform(0.01, 0.05)])
        sc.world_to_cam = inverted_rigid(look_at_pose(pos, (0, 0, 0.05)))
    for ob in sc.objects:
        if rng.rand() < 0.2:
            ob.casts_shadows = False
        if rng.rand() < 0.1:
            ob.visible = False
    return single_level_copy(sc)


VARIANTS = ["tabletop", "three_lights", "ssao", "auto_exposure", "no_plane_no_light", "empty", "ibl", "alpha_test", "sticker",
            "background_image", "plane_texture", "near_clip", "predicate", "id_limits", "odd_viewport", "multi_submesh", "low_poly_closeup",
            "pbr_textures", "pbr_textures_ibl", "projective", "c2_shape", "c5_shape"]
