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
    parity.assert_parity(after, ref, rgb_outlier_frac=1e-3)
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
