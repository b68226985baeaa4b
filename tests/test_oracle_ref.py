"""Pins of the oracle on the REFERENCE'S OWN CODE (no GPU needed):

  * golden vectors produced by running the reference's python/stillleben/diff.py + its compiled extension
    (tests/golden/make_ref_golden.py -> tests/golden/golden_diff.npz),
  * live calls into oracle/_ref/diff/libstillleben_diff_python.so (python/src/bridge_diff.cpp + diff.cu compiled from
    /root/reference by oracle/build_ref.py; the prebuilt file travels, the sources do not) on random inputs,
  * the reference's consolidated vertex stream (src/mesh_tools/consolidate.cpp run through oracle/_ref/meshtool ->
    tests/golden/*_ref.npz) against the repo's glTF front end,
  * the reference's own computeFrustumCorners / computeShadowMapMatrix (src/render_pass.cpp, cut out at build time and compiled with
    the reference's GL-less Magnum -> oracle/_ref/libframeref.so) against the oracle's frustum_corners / shadow_matrix,
  * the reference's GLSL programs compiled verbatim as C++ (oracle/_ref/libglslref.so) against the oracle's
    vertex / fragment / tone-map / SSAO / background / light-map restatements on random inputs (tests/test_glsl_ref.py).

The reference's extension has a CPU and a CUDA branch that differ at the image border and in the window walk order
(see oracle/orc_diff.cpp); here the oracle's "CPU branch" variant is pinned, tests/test_gpu_diff.py pins the CUDA one.
"""
import os
import sys

import numpy as np
import pytest
import torch

import diff_ref
import fixtures
import oracle_util as ou

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture()
def cpu_branch():
    ou.lib().orc_diff_set_variant(1)
    yield
    ou.lib().orc_diff_set_variant(0)


def _cases():
    g = fixtures.load_golden("golden_diff")
    for i in range(int(g["n_cases"])):
        yield i, {k[len(f"c{i}_"):]: v for k, v in g.items() if k.startswith(f"c{i}_")}


def test_masks_match_reference_goldens(cpu_branch):
    n_invalid = 0
    for i, c in _cases():
        inst, coord4 = np.ascontiguousarray(c["inst"]), np.ascontiguousarray(c["coord4"])
        valid = diff_ref.masks(inst, np.ascontiguousarray(coord4[..., 3]))
        assert np.array_equal(valid.astype(bool), c["valid"]), i
        n_invalid += int((~c["valid"]).sum())
        for o, idx in enumerate(c["ids"]):
            m, oc = diff_ref.dilate((inst == idx).astype(np.uint8), valid, coord4)
            assert np.array_equal(m, c["dilated_mask"][o]), (i, o)
            assert np.array_equal(oc, c["dilated_coords"][o]), (i, o)            # which neighbour's coordinate is copied
    assert n_invalid > 20                                                        # the occlusion branch is exercised


def test_pose_grad_matches_reference_goldens(cpu_branch):
    """orc_diff_pose_grad == python/stillleben/diff.py:355-523 run on the reference's own extension."""
    for i, c in _cases():
        got = diff_ref.oracle_pose_grad(*(np.ascontiguousarray(c[k]) for k in ("rgb", "inst", "coord4", "grad", "P", "poses", "ids")))
        ref = c["pose_grad"]
        assert np.abs(ref).max() > 1.0
        np.testing.assert_allclose(got, ref, rtol=5e-4, atol=5e-5 * np.abs(ref).max(), err_msg=f"case {i}")


def test_image_space_gradients_match_reference_goldens(cpu_branch):
    """compute_image_space_gradients (diff.py:73-127): F.conv2d with the scaled sobel rows == the central differences
    the oracle and the kernels use."""
    for i, c in _cases():
        img = c["rgb"][..., :3].astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
        H, W = img.shape[1:]
        px, py = np.pad(img, ((0, 0), (0, 0), (1, 1))), np.pad(img, ((0, 0), (1, 1), (0, 0)))
        gx = -(px[:, :, 2:] - px[:, :, :-2]) / np.float32(2.0 / W * 2.0)
        gy = -(py[:, 2:, :] - py[:, :-2, :]) / np.float32(2.0 / H * 2.0)
        gx[:, ~c["valid"]] = 0
        gy[:, ~c["valid"]] = 0
        np.testing.assert_allclose(gx, c["grad_x"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(gy, c["grad_y"], rtol=1e-5, atol=1e-4)


def test_apply_pose_delta_matches_reference():
    from stillleben_b200 import diff
    g = fixtures.load_golden("golden_diff")
    pose, delta = torch.from_numpy(g["apd_pose"]), torch.from_numpy(g["apd_delta"])
    np.testing.assert_allclose(diff.apply_pose_delta(pose, delta).numpy(), g["apd_ortho"], atol=2e-6)
    np.testing.assert_allclose(diff.apply_pose_delta(pose, delta, orthonormalize=False).numpy(), g["apd_raw"], atol=1e-6)
    np.testing.assert_allclose(diff.apply_pose_delta(pose[2], delta[2]).numpy(), g["apd_single"], atol=2e-6)


def _ref_ext():
    import build_ref
    ext = build_ref.load_diff()
    if ext is None:
        pytest.skip("oracle/_ref/diff not built (python oracle/build_ref.py diff)")
    return ext


@pytest.mark.parametrize("seed,H,W", [(3, 33, 47), (4, 64, 64), (5, 9, 200), (6, 3, 3), (7, 2, 5)])
def test_masks_match_reference_extension_live(cpu_branch, seed, H, W):
    """The reference's compiled bridge (CPU branch) against the oracle on fresh random maps, incl. degenerate sizes."""
    ext = _ref_ext()
    rng = np.random.RandomState(seed)
    inst = rng.randint(0, 4, size=(H, W)).astype(np.int16)
    depth = rng.uniform(0.5, 2.0, size=(H, W)).astype(np.float32)
    coord4 = rng.normal(size=(H, W, 4)).astype(np.float32)
    coord = np.ascontiguousarray(coord4[..., :3])
    valid_ref = ext.generate_sobel_valid_mask(torch.from_numpy(inst), torch.from_numpy(depth))
    valid = diff_ref.masks(inst, depth)
    assert np.array_equal(valid.astype(bool), valid_ref.numpy())
    for idx in (1, 2, 3):
        m_ref, c_ref = ext.dilate_object_mask(torch.from_numpy(inst == idx), valid_ref, torch.from_numpy(coord))
        m, oc = diff_ref.dilate((inst == idx).astype(np.uint8), valid, coord4)
        assert np.array_equal(m, m_ref.numpy())
        assert np.array_equal(oc, c_ref.numpy())


# ---------------------------------------------------------------------------------------------------------------------
# mesh front end: stillleben_b200/gltf.py against the reference's own consolidation (oracle/_ref/meshtool -> *_mesh.npz)
# ---------------------------------------------------------------------------------------------------------------------
def _assert_same_mesh(mine, ref, full_images=True):
    assert np.array_equal(mine.vertices.view(np.uint8).reshape(len(mine.vertices), -1), ref.vertices.view(np.uint8).reshape(len(ref.vertices), -1))
    assert np.array_equal(mine.indices, ref.indices)
    assert [tuple(s) for s in mine.submeshes] == [tuple(s) for s in ref.submeshes]
    assert len(mine.materials) == len(ref.materials)
    for a, b in zip(mine.materials, ref.materials):
        np.testing.assert_array_equal(np.float32(a.base_color), np.float32(b.base_color))
        np.testing.assert_array_equal(np.float32(a.emissive), np.float32(b.emissive))
        assert np.float32(a.metallic) == np.float32(b.metallic) and np.float32(a.roughness) == np.float32(b.roughness)
        assert (a.tex_base_color, a.tex_normal, a.tex_metallic_roughness, a.tex_emissive, a.tex_occlusion) == \
               (b.tex_base_color, b.tex_normal, b.tex_metallic_roughness, b.tex_emissive, b.tex_occlusion)
    assert len(mine.images) == len(ref.images)
    for a, b in zip(mine.images, ref.images):
        assert (a.wrap_s, a.wrap_t, a.min_filter, a.mag_filter) == (b.wrap_s, b.wrap_t, b.min_filter, b.mag_filter)
        if full_images:
            assert np.array_equal(a.pixels, b.pixels)
    np.testing.assert_array_equal(mine.bbox_min, ref.bbox_min)
    np.testing.assert_array_equal(mine.bbox_max, ref.bbox_max)


@pytest.mark.parametrize("asset", ["kitchen_sink", "pbr_patch"])
def test_gltf_front_end_matches_reference_consolidation(asset):
    """Byte for byte: 68-byte vertices (node transforms baked in Magnum's float32 order, computed tangents, colours, zero
    initialisation), indices, sub-mesh table in Object::loadVisual order, materials as RenderShader::setMaterial resolves
    them (incl. the importer dropping factors equal to the glTF default), one texture per glTF texture with its sampler."""
    from stillleben_b200 import gltf
    mine = gltf.load(os.path.join(fixtures.GOLDEN, "assets", asset + ".glb"))
    _assert_same_mesh(mine, fixtures.load_mesh(asset + "_mesh"))


@pytest.mark.parametrize("path,fixture", [("tests/cube.glb", "cube_glb_mesh"), ("tests/stanford_bunny/scene.gltf", "bunny_mesh")])
def test_gltf_front_end_matches_reference_on_its_own_assets(path, fixture):
    import hashlib
    from stillleben_b200 import gltf
    src = os.path.join("/root/reference", path)
    if not os.path.exists(src):
        pytest.skip("the reference's test assets are not on this machine")
    mine = gltf.load(src)
    _assert_same_mesh(mine, fixtures.load_mesh(fixture), full_images=False)
    z = np.load(os.path.join(fixtures.GOLDEN, fixture + ".npz"))
    for i, im in enumerate(mine.images):     # full-size image bytes as the reference's StbImageImporter delivers them (bottom-up rows)
        assert hashlib.sha256(im.pixels.tobytes()).digest() == z[f"sha256_image{i}"].tobytes()


def test_gltf_without_normals_is_an_error_like_in_the_reference(tmp_path):
    """The reference aborts on such a mesh (consolidate.cpp:84-87, verified with oracle/_ref/meshtool); here: ValueError."""
    import json
    from stillleben_b200 import gltf
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    idx = np.array([0, 1, 2], np.uint16)
    blob = pos.tobytes() + idx.tobytes() + b"\0\0"
    import base64
    js = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
          "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
          "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}],
          "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}], "nodes": [{"mesh": 0}], "scenes": [{"nodes": [0]}], "scene": 0}
    p = tmp_path / "no_normals.gltf"
    p.write_text(json.dumps(js))
    with pytest.raises(ValueError, match="NORMAL"):
        gltf.load(str(p))
    m = gltf.load(str(p), generate_missing_normals=True)
    np.testing.assert_allclose(m.vertices["normal"], [[0, 0, 1]] * 3, atol=1e-6)


def _frame_ref():
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libframeref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libframeref.so not built (python oracle/build_ref.py frame)")
    return C.CDLL(so)


def test_frustum_corners_and_shadow_matrix_match_the_reference_functions():
    """oracle/orc_render.cpp:frustum_corners / shadow_matrix against the reference's OWN computeFrustumCorners /
    computeShadowMapMatrix (src/render_pass.cpp:67-211, cut out at build time and compiled with the reference's Magnum:
    oracle/_ref/libframeref.so) on random table-top scenes, cameras and light directions — incl. the empty scene
    (near / far = -1 / 1) and objects partly outside the frustum."""
    import ctypes as C
    from stillleben_b200 import desc
    ref, orc = _frame_ref(), ou.lib()
    fp = C.POINTER(C.c_float)
    args = [fp, fp, C.c_int, fp, fp, fp, fp, fp, fp, fp]
    ref.ref_shadow_setup.argtypes = args
    orc.orc_test_shadow_setup.argtypes = args
    rng = np.random.RandomState(5)
    worst = 0.0
    for trial in range(300):
        n = int(rng.choice([0, 1, 2, 5, 20]))
        W, H = (640, 480) if trial % 2 else (1920, 1080)
        if trial % 3:
            P = desc.intrinsics_projection(1066.778 * W / 640, 1067.487 * W / 640, 312.9869 * W / 640, 241.3109 * H / 480, W, H)
        else:
            P = desc.fov_projection(W, H, float(rng.uniform(30, 90)))
        eye = rng.uniform(-1, 1, 3) * [0.8, 0.8, 0.3] + [0, 0, 0.9]
        V = desc.inverted_rigid(desc.look_at_pose(eye, rng.uniform(-0.2, 0.2, 3)))
        poses, pres, bmin, bmax = [], [], [], []
        for _ in range(n):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            x, y, z, w = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            pose = np.eye(4, dtype=np.float32); pose[:3, :3] = R; pose[:3, 3] = rng.uniform(-0.6, 0.6, 3) * [1, 1, 0.3]
            s = float(rng.uniform(0.02, 3.0))
            pre = np.eye(4, dtype=np.float32) * s; pre[3, 3] = 1; pre[:3, 3] = rng.uniform(-0.05, 0.05, 3)
            lo = rng.uniform(-0.1, 0.0, 3) / s
            hi = lo + rng.uniform(0.02, 0.3, 3) / s
            poses.append(pose.T.reshape(-1)); pres.append(pre.T.reshape(-1)); bmin.append(lo); bmax.append(hi)      # column-major
        as_f = lambda a, k: np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1) if len(a) else np.zeros(k, np.float32))
        poses, pres, bmin, bmax = as_f(poses, 16), as_f(pres, 16), as_f(bmin, 3), as_f(bmax, 3)
        L = rng.normal(size=3).astype(np.float32)
        L[2] = -abs(L[2]) - 0.2
        Pc, Vc = np.ascontiguousarray(P.T.reshape(-1), np.float32), np.ascontiguousarray(V.T.reshape(-1), np.float32)
        out = {}
        for name, fn in (("ref", ref.ref_shadow_setup), ("orc", orc.orc_test_shadow_setup)):
            corners, sm = np.zeros(24, np.float32), np.zeros(16, np.float32)
            fn(*(a.ctypes.data_as(fp) for a in (Pc, Vc)), n, *(a.ctypes.data_as(fp) for a in (poses, pres, bmin, bmax, L, corners, sm)))
            out[name] = (corners.copy(), sm.copy())
        (c_ref, m_ref), (c_orc, m_orc) = out["ref"], out["orc"]
        assert np.isfinite(c_ref).all() and np.isfinite(m_ref).all()
        scale_c, scale_m = np.abs(c_ref).max() + 1e-6, np.abs(m_ref).max() + 1e-6
        worst = max(worst, np.abs(c_orc - c_ref).max() / scale_c, np.abs(m_orc - m_ref).max() / scale_m)
        np.testing.assert_allclose(c_orc, c_ref, atol=1e-4 * scale_c, rtol=0)       # float32 inverses: Magnum's cofactor expansion vs the oracle's
        np.testing.assert_allclose(m_orc, m_ref, atol=1e-4 * scale_m, rtol=0)
    assert worst < 1e-4            # measured: 3.0e-5


def _host_ref():
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libhostref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libhostref.so not built (python oracle/build_ref.py host)")
    lib = C.CDLL(so)
    fp = C.POINTER(C.c_float)
    lib.ref_projection.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, fp]
    lib.ref_look_at.argtypes = [fp, fp, fp, fp]
    lib.ref_set_camera_pose.argtypes = [fp]
    lib.ref_mesh_normalise.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float, fp, fp]
    lib.ref_mesh_set_pretransform.argtypes = [fp, fp, fp, fp]
    lib.ref_sticker_projection.argtypes = [fp, fp, fp, fp, fp]
    return lib, fp


def _f(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))


def test_camera_helpers_match_the_reference_scene_functions():
    """desc.intrinsics_projection / fov_projection / look_at_pose (what sl.Scene.set_camera_* use) against the reference's own
    Scene::setCameraIntrinsics / setCameraFromFOV / setCameraLookAt (src/scene.cpp:192-271, compiled from the reference source)."""
    from stillleben_b200 import desc
    lib, fp = _host_ref()
    rng = np.random.RandomState(11)
    out = np.zeros(16, np.float32)
    for W, H in ((640, 480), (1920, 1080), (320, 240), (127, 33)):
        for _ in range(20):
            fx, fy = rng.uniform(200, 2500, 2)
            cx, cy = rng.uniform(0.3, 0.7) * W, rng.uniform(0.3, 0.7) * H
            lib.ref_projection(W, H, 0, fx, fy, cx, cy, out.ctypes.data_as(fp))
            np.testing.assert_allclose(desc.intrinsics_projection(fx, fy, cx, cy, W, H), out.reshape(4, 4).T, rtol=2e-6, atol=1e-7)
            fov = float(rng.uniform(20, 110))
            lib.ref_projection(W, H, 1, np.deg2rad(fov), 0, 0, 0, out.ctypes.data_as(fp))
            np.testing.assert_allclose(desc.fov_projection(W, H, fov), out.reshape(4, 4).T, rtol=5e-6, atol=1e-7)
    for _ in range(100):
        pos, at = rng.uniform(-2, 2, 3), rng.uniform(-0.5, 0.5, 3)
        up = (0.0, 0.0, 1.0) if rng.rand() < 0.7 else rng.normal(size=3)
        assert lib.ref_look_at(_f(pos).ctypes.data_as(fp), _f(at).ctypes.data_as(fp), _f(up).ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
        np.testing.assert_allclose(desc.look_at_pose(pos, at, up), out.reshape(4, 4).T, atol=2e-6)
    # setCameraPose rejects non-rigid poses (scene.cpp:192-200); the look-at poses above are rigid by construction
    bad = np.eye(4, dtype=np.float32); bad[0, 0] = 2.0
    assert lib.ref_set_camera_pose(_f(bad.T).ctypes.data_as(fp)) == 1
    assert lib.ref_set_camera_pose(_f(np.eye(4).T).ctypes.data_as(fp)) == 0


def test_mesh_pretransform_logic_matches_the_reference_mesh_functions(monkeypatch):
    """sl.Mesh.center_bbox / scale_to_bbox_diagonal / pretransform setter / bbox against the reference's Mesh::centerBBox /
    scaleToBBoxDiagonal / setPretransform / bbox (src/mesh.cpp:1019-1081), and sl.Object._sticker_projection against
    Object::stickerViewProjection (src/object.cpp:494-513) — all compiled from the reference source."""
    import torch
    from stillleben_b200 import sl
    monkeypatch.setattr(sl, "_ctx", object())                         # host logic only
    lib, fp = _host_ref()
    rng = np.random.RandomState(12)
    base = fixtures.load_mesh("cube_glb_mesh")
    pre_out, bbox_out = np.zeros(16, np.float32), np.zeros(6, np.float32)
    for trial in range(60):
        lo = rng.uniform(-3, 0, 3).astype(np.float32)
        hi = (lo + rng.uniform(0.01, 4, 3)).astype(np.float32)
        md = type(base)(base.vertices, base.indices, base.submeshes, base.materials, base.images, lo.copy(), hi.copy(), "m")
        mesh = sl.Mesh.from_data(md)
        center, mode, target = bool(trial % 2), trial % 3, float(rng.uniform(0.05, 2.0))
        if center:
            mesh.center_bbox()
        if mode:
            mesh.scale_to_bbox_diagonal(target, "exact" if mode == 1 else "order_of_magnitude")
        lib.ref_mesh_normalise(_f(lo).ctypes.data_as(fp), _f(hi).ctypes.data_as(fp), int(center), mode, target, pre_out.ctypes.data_as(fp),
                               bbox_out.ctypes.data_as(fp))
        ref_pre = pre_out.reshape(4, 4).T
        np.testing.assert_allclose(mesh.pretransform.numpy(), ref_pre, rtol=2e-6, atol=1e-6 * max(1.0, np.abs(ref_pre).max()))
        bb = mesh.bbox
        np.testing.assert_allclose(np.concatenate([bb.min.numpy(), bb.max.numpy()]), bbox_out, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(bbox_out).max()))
        # sticker projection of an object of this mesh
        q = rng.normal(size=4).astype(np.float32); q /= np.linalg.norm(q)
        obj = sl.Object(mesh)
        obj.sticker_rotation = torch.from_numpy(q)
        obj.sticker_texture = object()
        lib.ref_sticker_projection(_f(lo).ctypes.data_as(fp), _f(hi).ctypes.data_as(fp), _f(mesh.pretransform.numpy().T).ctypes.data_as(fp),
                                   _f(q).ctypes.data_as(fp), pre_out.ctypes.data_as(fp))
        ref_sp = pre_out.reshape(4, 4).T
        np.testing.assert_allclose(obj._sticker_projection(), ref_sp, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(ref_sp).max()))
    # setPretransform: uniform scale x rigid accepted and split like the reference; anisotropic scale rejected
    scale_out, rigid_out = np.zeros(1, np.float32), np.zeros(16, np.float32)
    for trial in range(40):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = sl.quat_to_matrix(torch.tensor(q, dtype=torch.float32)).numpy()
        s = float(rng.uniform(0.01, 20))
        m = np.eye(4, dtype=np.float32); m[:3, :3] = s * R; m[:3, 3] = rng.uniform(-1, 1, 3)
        mesh = sl.Mesh.from_data(base)
        mesh.pretransform = torch.from_numpy(m)
        assert lib.ref_mesh_set_pretransform(_f(m.T).ctypes.data_as(fp), scale_out.ctypes.data_as(fp), rigid_out.ctypes.data_as(fp), pre_out.ctypes.data_as(fp)) == 0
        assert abs(mesh._scale - float(scale_out[0])) < 3e-6 * s
        np.testing.assert_allclose(mesh._rigid, rigid_out.reshape(4, 4).T, atol=2e-5 * max(1.0, 1.0 / s))
        np.testing.assert_allclose(mesh.pretransform.numpy(), pre_out.reshape(4, 4).T, atol=3e-6 * max(1.0, s))
    m = np.eye(4, dtype=np.float32); m[0, 0] = 2.0
    assert lib.ref_mesh_set_pretransform(_f(m.T).ctypes.data_as(fp), scale_out.ctypes.data_as(fp), rigid_out.ctypes.data_as(fp), pre_out.ctypes.data_as(fp)) == 1
    with pytest.raises(ValueError, match="not uniform"):
        sl.Mesh.from_data(base).pretransform = torch.from_numpy(m)


def test_vertex_edit_twin_matches_the_reference_mesh_functions():
    """The oracle's vertex-edit twin (orc_mesh_update_positions_and_colors / recompute_normals — which the DEVICE path
    reproduces bit for bit, tests/test_gpu_assets.py) against the reference's own Mesh::updateVertexPositionsAndColors /
    setVertexPositions / recomputeNormals (src/mesh.cpp:763-870, compiled from the reference source): positions, colours and the
    area-weighted normals, incl. the one-based ids and the size check of setVertexPositions."""
    import ctypes as C
    from stillleben_b200 import abi, synth
    lib, fp = _host_ref()
    ip, up = C.POINTER(C.c_int), C.POINTER(C.c_uint)
    lib.ref_mesh_edit.argtypes = [fp, fp, fp, C.c_int, up, C.c_int, ip, C.c_int, fp, fp]
    L = ou.lib()
    L.orc_mesh_update_positions_and_colors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.orc_mesh_read_vertices.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_mesh_recompute_normals.argtypes = [C.c_void_p]
    rng = np.random.RandomState(21)
    for mesh in (synth.shape_mesh("blob", 9, nu=24, nv=12, textured=True, tex_size=32), fixtures.load_mesh("cube_glb_mesh"), fixtures.load_mesh("kitchen_sink_mesh")):
        assets = ou.OracleAssets()
        h = assets.handle_of(mesh)
        n = len(mesh.vertices)
        ids = (rng.permutation(n)[: max(1, n // 2)] + 1).astype(np.int32)
        dpos = (rng.normal(size=(len(ids), 3)) * 0.02).astype(np.float32)
        dcol = rng.rand(len(ids), 4).astype(np.float32)
        pos, nrm, col = (np.ascontiguousarray(mesh.vertices[k]).copy() for k in ("position", "normal", "color"))
        idx = np.ascontiguousarray(mesh.indices, np.uint32)
        assert lib.ref_mesh_edit(pos.ctypes.data_as(fp), nrm.ctypes.data_as(fp), col.ctypes.data_as(fp), n, idx.ctypes.data_as(up), len(idx),
                                 ids.ctypes.data_as(ip), len(ids), dpos.ctypes.data_as(fp), dcol.ctypes.data_as(fp)) == 0
        assert L.orc_mesh_update_positions_and_colors(h, ids.ctypes.data, len(ids), dpos.ctypes.data, dcol.ctypes.data) == 0
        got = np.empty(n, abi.VERTEX_DTYPE)
        L.orc_mesh_read_vertices(h, got.ctypes.data)
        assert np.array_equal(got["position"], pos) and np.array_equal(got["color"], col)
        ok = np.isfinite(nrm).all(1)                                   # (vertices without faces / zero-area fans: NaN on both sides)
        assert np.array_equal(np.isfinite(got["normal"]).all(1), ok)
        np.testing.assert_allclose(got["normal"][ok], nrm[ok], atol=2e-6)
        assert np.abs(nrm[ok] - mesh.vertices["normal"][ok]).max() > 1e-3    # the edit did change them
        # setVertexPositions
        newp = (pos * np.float32(1.07) + np.float32(0.01)).astype(np.float32)
        assert lib.ref_mesh_edit(pos.ctypes.data_as(fp), nrm.ctypes.data_as(fp), col.ctypes.data_as(fp), n, idx.ctypes.data_as(up), len(idx),
                                 None, n, newp.ctypes.data_as(fp), None) == 0
        expect = got.copy(); expect["position"] = newp
        L.orc_mesh_update_vertices(h, expect.ctypes.data, n)
        L.orc_mesh_recompute_normals(h)
        L.orc_mesh_read_vertices(h, got.ctypes.data)
        assert np.array_equal(pos, newp)
        np.testing.assert_allclose(got["normal"][ok], nrm[ok], atol=2e-6)
        # "Number of new vertices should match the existing mesh vertices" (mesh.cpp:861-862)
        assert lib.ref_mesh_edit(pos.ctypes.data_as(fp), nrm.ctypes.data_as(fp), col.ctypes.data_as(fp), n, idx.ctypes.data_as(up), len(idx),
                                 None, n - 1, newp.ctypes.data_as(fp), None) == 1


def test_ibl_reader_matches_the_reference_light_map_parser(tmp_path, monkeypatch):
    """sl.LightMap's .ibl reader against the reference's own IBLSpec::load / LightSpec::load / addLight (src/light_map.cpp:55-151,
    314-345, compiled from the reference source with Corrade's Configuration): the reflection file, and direction + colour of
    the sun and the two extra lights, over files with quotes, spaces, missing keys and missing groups."""
    import ctypes as C
    from stillleben_b200 import sl
    monkeypatch.setattr(sl, "_ctx", object())
    lib, fp = _host_ref()
    lib.ref_ibl_lights.argtypes = [C.c_char_p, C.c_char_p, C.c_int, fp, fp, fp]
    np.save(tmp_path / "env.npy", np.ones((8, 16, 3), np.float32))
    rng = np.random.RandomState(4)
    cases = []
    for k in range(12):
        sec = ['[Header]', 'ICOfile = "x.jpg"', '[Reflection]', 'REFfile = "env.npy"' if k % 2 else "REFfile=env.npy", "REFmap = 1",
               f"REFgamma = {rng.uniform(1, 2.2):.3f}", f"REFmulti = {rng.uniform(0.5, 2):.2f}"]
        groups = [("Sun", "SUN"), ("Light1", "LIGHT"), ("Light2", "LIGHT")][: 1 + k % 3] if k % 4 != 3 else [("Light1", "LIGHT")]
        for name, pre in groups:
            sec.append(f"[{name}]")
            if k % 5 != 4:
                r, g, b = rng.randint(0, 256, 3)
                sec.append(f"{pre}color = {r},{g},{b}" if k % 2 else f"{pre}color={r}, {g} ,{b}")
            if k % 3 != 2:
                sec.append(f"{pre}multi = {rng.uniform(0.2, 5):.3f}")
            sec.append(f"{pre}u = {rng.uniform(-0.5, 0.5):.4f}")
            if k % 7 != 6:
                sec.append(f"{pre}v = {rng.uniform(0.05, 0.95):.4f}")
        cases.append("\n".join(sec) + "\n")
    for k, text in enumerate(cases):
        path = tmp_path / f"case{k}.ibl"
        path.write_text(text)
        ref_file = C.create_string_buffer(256)
        gm, dirs, cols = np.zeros(2, np.float32), np.zeros(9, np.float32), np.zeros(9, np.float32)
        n = lib.ref_ibl_lights(str(path).encode(), ref_file, 256, gm.ctypes.data_as(fp), dirs.ctypes.data_as(fp), cols.ctypes.data_as(fp))
        assert n >= 1 and ref_file.value == b"env.npy", (k, n, ref_file.value)
        lm = sl.LightMap(str(path))
        assert len(lm.data.light_directions) == n, k
        np.testing.assert_allclose(np.asarray(lm.data.light_directions, np.float32).reshape(-1), dirs[: 3 * n], atol=2e-6, err_msg=text)
        np.testing.assert_allclose(np.asarray(lm.data.light_colors, np.float32).reshape(-1), cols[: 3 * n], rtol=2e-6, atol=1e-7, err_msg=text)
    # files the reference refuses: no [Reflection] group, a mapping mode other than 1
    (tmp_path / "bad1.ibl").write_text("[Sun]\nSUNu = 0.1\nSUNv = 0.2\n")
    (tmp_path / "bad2.ibl").write_text('[Reflection]\nREFfile = "env.npy"\nREFmap = 2\n')
    for name in ("bad1.ibl", "bad2.ibl"):
        gm, dirs, cols = np.zeros(2, np.float32), np.zeros(9, np.float32), np.zeros(9, np.float32)
        assert lib.ref_ibl_lights(str(tmp_path / name).encode(), C.create_string_buffer(256), 256, gm.ctypes.data_as(fp), dirs.ctypes.data_as(fp),
                                  cols.ctypes.data_as(fp)) == -1
        with pytest.raises((RuntimeError, KeyError)):
            sl.LightMap(str(tmp_path / name))


def test_ssao_tables_match_reference_generator():
    """src/shaders/ssao_shader.cpp:72-112 (the noise / kernel loops of the SSAOShader constructor, std::mt19937{0xdeadbeef}) cut out
    of the reference and run here (oracle/_ref/libhostref.so:ref_ssao_tables) against the oracle's tables — the ones
    tests/test_glsl_ref.py feeds to the verbatim ssao_shader.frag and the product's slb_host.cu:ssao_tables restates."""
    import ctypes as C
    import oracle_util as ou
    lib, _ = _host_ref()
    if not hasattr(lib, "ref_ssao_tables"):
        pytest.skip("oracle/_ref/libhostref.so predates ref_ssao_tables (python oracle/build_ref.py host --force)")
    lib.ref_ssao_tables.argtypes = [C.c_void_p, C.c_void_p]
    n_ref, k_ref = np.zeros((16, 3), np.float32), np.zeros((64, 3), np.float32)
    lib.ref_ssao_tables(n_ref.ctypes.data, k_ref.ctypes.data)
    orc = C.CDLL(ou.ORACLE_SO)
    orc.orc_test_ssao_tables.argtypes = [C.c_void_p, C.c_void_p]
    n_orc, k_orc = np.zeros((16, 3), np.float32), np.zeros((64, 3), np.float32)
    orc.orc_test_ssao_tables(n_orc.ctypes.data, k_orc.ctypes.data)
    assert np.abs(n_ref).max() > 0.5 and np.abs(k_ref).max() > 0.3          # the cut really ran
    np.testing.assert_array_equal(n_orc, n_ref)
    np.testing.assert_allclose(k_orc, k_ref, rtol=2e-6, atol=1e-7)          # normalized(): Magnum's 1/sqrt(dot) vs the oracle's
