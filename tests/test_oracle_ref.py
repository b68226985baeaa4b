"""Pins of the oracle on the REFERENCE'S OWN CODE (no GPU needed):

  * golden vectors produced by running the reference's python/stillleben/diff.py + its compiled extension
    (tests/golden/make_ref_golden.py -> tests/golden/golden_diff.npz),
  * live calls into oracle/_ref/diff/libstillleben_diff_python.so (python/src/bridge_diff.cpp + diff.cu compiled from
    /root/reference by oracle/build_ref.py; the prebuilt file travels, the sources do not) on random inputs,
  * the reference's consolidated vertex stream (src/mesh_tools/consolidate.cpp run through oracle/_ref/meshtool ->
    tests/golden/*_ref.npz) against the repo's glTF front end,
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
