"""Load-time kernels and the config-4 mask kernels against the oracle."""
import ctypes as C

import numpy as np
import pytest

import fixtures
import oracle_util as ou
from stillleben_b200 import abi, synth
from stillleben_b200.desc import ImageData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(64, 64, 3), (37, 101, 4), (1, 9, 3), (128, 2, 4)])
def test_mip_chain_bit_exact(gpu_ctx, shape):
    rng = np.random.RandomState(shape[0] * 7 + shape[1])
    img = ImageData(rng.randint(0, 256, size=shape).astype(np.uint8))
    L = ou.lib()
    ci = img.to_c()
    h = L.orc_texture_create(C.byref(ci), abi.TEXTURE_2D)
    w, hh = C.c_int(), C.c_int()
    n_levels = L.orc_texture_level(h, 0, C.byref(w), C.byref(hh), None)
    n_gpu, _ = gpu_ctx.texture_level(img, 0)
    assert n_gpu == n_levels
    for lvl in range(n_levels):
        L.orc_texture_level(h, lvl, C.byref(w), C.byref(hh), None)
        ref = np.empty((hh.value, w.value, 4), np.uint8)
        L.orc_texture_level(h, lvl, C.byref(w), C.byref(hh), ref.ctypes.data)
        _, got = gpu_ctx.texture_level(img, lvl)
        assert got.shape == ref.shape and (got == ref).all(), lvl
    L.orc_texture_destroy(h)


def test_lightmap_precompute_matches_oracle(gpu_ctx, oracle_assets):
    # light_map.cpp:376-611 at reduced sizes (64 / 16 / 32 / 64, 64 samples)
    lm = fixtures.light_map_data()
    g = gpu_ctx.read_lightmap(lm)
    r = oracle_assets.read_lightmap(lm)
    names = ["env", "irradiance", "prefilter", "lut"]
    for name, a, b in zip(names, g, r):
        assert a.shape == b.shape, name
        np.testing.assert_allclose(a, b, rtol=2e-3, atol=2e-4, err_msg=name)


def test_vertex_update_path(gpu_ctx):
    # Mesh::updateVertexPositionsAndColors / recompileMesh (src/mesh.cpp:763-855): re-upload, re-render
    mesh = synth.shape_mesh("blob", 5, nu=24, nv=12, textured=False)
    scene = synth.tabletop_scene([mesh], 3, n_objects=2, width=160, height=120, intrinsics=None)
    before = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    b0 = before.frame_dict(0)
    mesh.vertices["position"] *= 0.5
    gpu_ctx.update_vertices(mesh)
    after = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL).frame_dict(0)
    assets = ou.OracleAssets()
    ref = ou.render(scene, assets, want_hdr=False)
    import parity
    parity.assert_parity(after, ref, rgb_outliers=4)
    assert (after["instance_index"] != b0["instance_index"]).any()


def _diff_inputs(gpu_ctx):
    scene = fixtures.variant("tabletop")
    res = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    f = res.frame_dict(0)
    return f["instance_index"][..., 0].view(np.int16).copy(), np.ascontiguousarray(f["coord"][..., 3]), np.ascontiguousarray(f["coord"])


def test_diff_kernels_match_oracle(gpu_ctx):
    import torch
    inst, depth, coord = _diff_inputs(gpu_ctx)
    H, W = inst.shape
    L = ou.lib()
    valid_ref = np.zeros((H, W), np.uint8)
    L.orc_diff_sobel_valid_mask(inst.ctypes.data, depth.ctypes.data, valid_ref.ctypes.data, H, W)
    dev = torch.device("cuda", 0)
    t_inst, t_depth = torch.from_numpy(inst).to(dev), torch.from_numpy(depth).to(dev)
    t_valid = torch.empty((H, W), dtype=torch.uint8, device=dev)
    rc = gpu_ctx.lib.slb_diff_sobel_valid_mask(gpu_ctx.h, t_inst.data_ptr(), t_depth.data_ptr(), t_valid.data_ptr(), H, W, None)
    assert rc == 0
    gpu_ctx.synchronize()
    assert (t_valid.cpu().numpy() == valid_ref).all()
    assert 0 < (valid_ref == 0).sum() < H * W
    for obj in (1, 2, 3):
        mask = (inst == obj).astype(np.uint8)
        mo_ref, co_ref = np.zeros((H, W), np.uint8), np.zeros((H, W, 3), np.float32)
        L.orc_diff_dilate_object_mask(mask.ctypes.data, valid_ref.ctypes.data, coord.ctypes.data, 4, mo_ref.ctypes.data, co_ref.ctypes.data, H, W)
        t_mask, t_coord = torch.from_numpy(mask).to(dev), torch.from_numpy(coord).to(dev)
        t_mo = torch.empty((H, W), dtype=torch.uint8, device=dev)
        t_co = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        rc = gpu_ctx.lib.slb_diff_dilate_object_mask(gpu_ctx.h, t_mask.data_ptr(), t_valid.data_ptr(), t_coord.data_ptr(), 4, t_mo.data_ptr(),
                                                     t_co.data_ptr(), H, W, None)
        assert rc == 0
        gpu_ctx.synchronize()
        assert (t_mo.cpu().numpy() == mo_ref).all()
        assert (t_co.cpu().numpy().view(np.uint32) == co_ref.view(np.uint32)).all()
        assert mo_ref.sum() >= mask.sum()


def test_vertex_edit_path_recomputes_normals_like_the_reference(gpu_ctx):
    """Mesh::updateVertexPositionsAndColors / setVertexPositions / recomputeNormals (src/mesh.cpp:763-870) on the device copy:
    positions += update (one-based ids), colours += update, area-weighted vertex normals recomputed — the vertex stream read
    back must equal the oracle twin's BIT FOR BIT, and the edited mesh must shade like the oracle's edited mesh."""
    import parity
    mesh = synth.shape_mesh("blob", 9, nu=24, nv=12, textured=True, tex_size=32)
    scene = synth.tabletop_scene([mesh], 5, n_objects=2, width=160, height=120, intrinsics=None)
    assets = ou.OracleAssets()
    L = ou.lib()
    L.orc_mesh_update_positions_and_colors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.orc_mesh_read_vertices.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_mesh_recompute_normals.argtypes = [C.c_void_p]
    before = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL).frame_dict(0)
    rng = np.random.RandomState(2)
    n = len(mesh.vertices)
    ids = rng.permutation(n)[: n // 2].astype(np.int32) + 1
    dpos = (rng.normal(size=(len(ids), 3)) * 0.02).astype(np.float32)
    dcol = rng.rand(len(ids), 4).astype(np.float32)
    gpu_ctx.update_positions_and_colors(mesh, ids, dpos, dcol)
    h = assets.handle_of(mesh)
    assert L.orc_mesh_update_positions_and_colors(h, ids.ctypes.data, len(ids), dpos.ctypes.data, dcol.ctypes.data) == 0
    ref_v = np.empty(n, abi.VERTEX_DTYPE)
    L.orc_mesh_read_vertices(h, ref_v.ctypes.data)
    got_v = gpu_ctx.read_vertices(mesh)
    for f in abi.VERTEX_DTYPE.names:
        assert np.array_equal(got_v[f], ref_v[f], equal_nan=True), f
    assert np.abs(got_v["normal"] - mesh.vertices["normal"]).max() > 1e-3          # the normals did change
    np.testing.assert_allclose(np.linalg.norm(got_v["normal"], axis=1), 1.0, atol=1e-5)
    assert np.array_equal(got_v["color"][ids - 1], mesh.vertices["color"][ids - 1] + dcol)
    after = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL).frame_dict(0)
    ref = ou.render(scene, assets, want_hdr=False)
    parity.assert_parity(after, ref, rgb_outliers=4)
    assert (after["normals"] != before["normals"]).any()
    # setVertexPositions: everything replaced, normals again; device-resident update arrays are accepted as well
    import torch
    newp = (got_v["position"] * np.float32(1.1)).astype(np.float32)
    gpu_ctx.set_positions(mesh, torch.from_numpy(newp).cuda())
    upd = np.zeros(n, np.float32)
    expect = ref_v.copy(); expect["position"] = newp
    ou.lib().orc_mesh_update_vertices(h, expect.ctypes.data, n)
    L.orc_mesh_recompute_normals(h)
    L.orc_mesh_read_vertices(h, ref_v.ctypes.data)
    got_v = gpu_ctx.read_vertices(mesh)
    assert np.array_equal(got_v["position"], newp) and np.array_equal(got_v["normal"], ref_v["normal"], equal_nan=True)
    del upd
    # error behaviour: a vertex id outside 1..n, a wrong vertex count (std::invalid_argument -> ValueError)
    with pytest.raises(ValueError):
        gpu_ctx.update_positions_and_colors(mesh, np.array([n + 1], np.int32), np.zeros((1, 3), np.float32))
    with pytest.raises(ValueError, match="Number of new vertices"):
        gpu_ctx.set_positions(mesh, np.zeros((n - 1, 3), np.float32))


def test_upload_rejects_out_of_range_indices(gpu_ctx):
    from stillleben_b200.desc import MeshData
    src = synth.shape_mesh("blob", 3, nu=8, nv=4, textured=False)
    bad = src.indices.copy(); bad[5] = len(src.vertices)
    with pytest.raises(ValueError, match="index value out of range"):
        gpu_ctx.handle_of(MeshData(src.vertices.copy(), bad, list(src.submeshes), list(src.materials), []))


def test_asset_release_frees_and_reuploads(gpu_ctx):
    import torch
    mesh = synth.shape_mesh("blob", 4, nu=64, nv=32, textured=True, tex_size=256)
    scene = synth.tabletop_scene([mesh], 6, n_objects=1, width=64, height=48, intrinsics=None)
    a = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL).frame_dict(0)
    free0 = torch.cuda.mem_get_info(0)[0]
    gpu_ctx.release(mesh)
    assert id(mesh) not in gpu_ctx._handles and torch.cuda.mem_get_info(0)[0] >= free0
    b = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL).frame_dict(0)        # uploaded again on use
    for k in a:
        assert (a[k].view(np.uint8) == b[k].view(np.uint8)).all(), k


def _all_face_positions(size):
    """WorldPos of every texel of a size^2 cube map, [6][size][size][3] (GL cube-face orientation, light_map.cpp:185-192)."""
    c = (np.arange(size, dtype=np.float64) + 0.5) / size * 2.0 - 1.0
    a, b = np.meshgrid(c, c, indexing="xy")                     # a along x (s), b along y (t)
    one = np.ones_like(a)
    faces = [(one, -b, -a), (-one, -b, a), (a, one, b), (a, -one, -b), (a, -b, one), (-a, -b, -one)]
    return np.ascontiguousarray(np.stack([np.stack(f, -1) for f in faces]), np.float32)


def test_lightmap_precompute_at_the_reference_sizes(gpu_ctx):
    """LightMap::load at the sizes the reference uses (light_map.cpp:381,451,510,580: environment 512^2 with mips,
    irradiance 32^2 at 0.02 rad steps, prefilter 128^2 x 5 mips and BRDF LUT 512^2 with 1024 samples each): the full
    environment, irradiance and prefilter cubes and 20 000 random LUT texels against the oracle's per-texel functions (the ones tests/test_glsl_ref.py holds against the reference's shader text)."""
    from stillleben_b200.desc import LightMapData
    src = fixtures.light_map_data()
    lm = LightMapData(src.equirect.copy(), list(src.light_directions), list(src.light_colors))      # a fresh object: not the cached small one
    saved = gpu_ctx.lightmap_sizes
    gpu_ctx.lightmap_sizes = (0, 0, 0, 0, 0)                                                         # the reference's sizes
    try:
        env, irr, pre, lut = gpu_ctx.read_lightmap(lm)
    finally:
        gpu_ctx.lightmap_sizes = saved
    assert env.shape == (6, 512, 512, 4) and irr.shape == (6, 32, 32, 4) and lut.shape == (512, 512, 4)
    L = ou.lib()
    L.orc_test_lightmap_texels.argtypes = [C.c_int, C.POINTER(abi.LightmapDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p]
    L.orc_test_brdf_lut.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    d, eq = ou.lightmap_desc(lm)
    h = L.orc_lightmap_create(C.byref(d), 512, 4, 4, 4, 4)           # full environment + mip chain; the small maps are unused
    n = L.orc_lightmap_read(h, 0, None)
    env_ref = np.empty(n, np.float32); L.orc_lightmap_read(h, 0, env_ref.ctypes.data)
    np.testing.assert_allclose(env, env_ref.reshape(env.shape), rtol=1e-4, atol=1e-5)
    wp = _all_face_positions(32)                                  # EVERY irradiance texel: 6144 x 24 806 environment samples
    ref = np.zeros((6, 32, 32, 3), np.float32)
    L.orc_test_lightmap_texels(1, C.byref(d), h, wp.ctypes.data, 6 * 32 * 32, 0.0, 1024, float(np.log2(512 / 32)), ref.ctypes.data)
    np.testing.assert_allclose(irr[..., :3], ref, rtol=2e-3, atol=2e-4)
    off = 0
    for mip in range(5):                                          # every prefilter texel of every roughness level
        size = 128 >> mip
        level = pre[off:off + 6 * size * size * 4].reshape(6, size, size, 4); off += level.size
        wp = _all_face_positions(size)
        ref = np.zeros((6, size, size, 3), np.float32)
        L.orc_test_lightmap_texels(2, C.byref(d), h, wp.ctypes.data, 6 * size * size, mip / 4.0, 1024, 0.0, ref.ctypes.data)
        np.testing.assert_allclose(level[..., :3], ref, rtol=2e-3, atol=2e-4, err_msg=f"prefilter mip {mip}")
    rng = np.random.RandomState(0)
    x, y = rng.randint(0, 512, 20000), rng.randint(0, 512, 20000)
    uv = np.ascontiguousarray(np.stack([(x + 0.5) / 512, (y + 0.5) / 512], 1), np.float32)
    ref = np.zeros((20000, 2), np.float32)
    L.orc_test_brdf_lut(uv.ctypes.data, 20000, 1024, ref.ctypes.data)
    np.testing.assert_allclose(lut[y, x, :2], ref, rtol=2e-3, atol=2e-4)
    L.orc_lightmap_destroy(h)
    gpu_ctx.release(lm)
