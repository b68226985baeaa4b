"""Config 4 (BASELINE.json): the fused render-and-compare backward on the GPU against the CPU oracle, and the
reference's own gradient test (tests/test_grad.py) replayed through the `sl` / `diff` mirrors."""
import numpy as np
import pytest
import torch

import diff_ref
import fixtures
from stillleben_b200 import abi, diff, sl

pytestmark = pytest.mark.gpu


def gpu_pose_grad(ctx, rgb, inst, coord4, grad, P, poses, ids):
    dev = torch.device("cuda", 0)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (rgb, inst, coord4, grad)]
    out = torch.full((len(ids), 6), float("nan"), dtype=torch.float32, device=dev)
    Pc = np.ascontiguousarray(np.asarray(P, np.float32).T)                       # column-major for the ABI
    Tc = np.ascontiguousarray(np.asarray(poses, np.float32).transpose(0, 2, 1))
    idv = np.ascontiguousarray(ids, np.int32)
    rc = ctx.lib.slb_diff_pose_grad(ctx.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), Pc.ctypes.data,
                                    Tc.ctypes.data, idv.ctypes.data, len(ids), out.data_ptr(), inst.shape[0], inst.shape[1], None)
    assert rc == 0, ctx.lib.slb_last_error(ctx.h)
    ctx.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("seed,shape", [(0, (60, 80)), (1, (37, 131)), (2, (8, 32))])
def test_pose_grad_matches_oracle_on_synthetic_maps(gpu_ctx, seed, shape):
    args = diff_ref.synthetic_inputs(seed, H=shape[0], W=shape[1])
    ref = diff_ref.oracle_pose_grad(*args)
    got = gpu_pose_grad(gpu_ctx, *args)
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3 * max(1.0, np.abs(ref).max()))
    again = gpu_pose_grad(gpu_ctx, *args)
    assert np.array_equal(got, again)                                           # fixed-order reduction: deterministic


def test_pose_grad_matches_oracle_on_a_render(gpu_ctx):
    scene = fixtures.small_tabletop_scene()
    res = gpu_ctx.render([scene], target_mask=abi.TARGETS_ALL)
    gpu_ctx.synchronize()
    fr = res.frame_dict(0)
    rgb, inst, coord4 = fr["rgb"], fr["instance_index"].view(np.int16).reshape(fr["rgb"].shape[:2]), fr["coord"]
    H, W = inst.shape
    grad = np.random.RandomState(5).normal(size=(3, H, W)).astype(np.float32)
    P = np.asarray(scene.projection, np.float32)
    poses = np.stack([np.asarray(o.pose, np.float32) for o in scene.objects])
    ids = np.array([o.instance_index for o in scene.objects], np.int32)
    ref = diff_ref.oracle_pose_grad(rgb, inst, coord4, grad, P, poses, ids)
    got = gpu_pose_grad(gpu_ctx, rgb, inst, coord4, grad, P, poses, ids)
    assert np.abs(ref).max() > 0
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3 * np.abs(ref).max())


def gaussian_pyramid_grad(observed, rendered):
    """tests/test_grad.py:21-62 (get_gaussian_pyramid_comparison) with an all-ones mask; cv2.getGaussianKernel(3, 3)
    written out."""
    F = torch.nn.functional
    H, W = observed.shape[:2]
    g = torch.exp(-(torch.arange(3.0) - 1.0) ** 2 / (2 * 3.0 ** 2))
    g = g / g.sum()
    kernel = (g[:, None] * g[None, :]).view(1, 1, 3, 3).float()
    obs = observed.permute(2, 0, 1).contiguous()
    rnd = torch.nn.Parameter(rendered.permute(2, 0, 1).contiguous())
    r = F.conv2d(rnd.view(3, 1, H, W), kernel, padding=1)
    o = F.conv2d(obs.view(3, 1, H, W), kernel, padding=1)
    diff2 = ((r - o).squeeze()) ** 2
    in_1 = diff2.view(3, 1, H, W)
    l1 = F.conv2d(in_1, kernel, padding=1).view(1, -1)
    in_2 = F.interpolate(in_1, scale_factor=0.5, mode="bilinear")
    l2 = F.conv2d(in_2, kernel, padding=1).view(1, -1)
    in_3 = F.interpolate(in_2, scale_factor=0.5, mode="bilinear")
    l3 = F.conv2d(in_3, kernel, padding=1).view(1, -1)
    loss = torch.cat([l1, l2, l3], dim=1).sum() / float(H * W)
    loss.backward()
    return rnd.grad.view(3, H, W).clone(), float(loss.detach())


def test_gradient_sign_like_the_reference_test_grad():
    """tests/test_grad.py:93-153: bunny at the literal pose, +0.01 along each of the six pose parameters, render,
    Gaussian-pyramid loss against the unperturbed image, backpropagate: the gradient w.r.t. the perturbed
    parameter must be positive (descending it moves the object back)."""
    sl.init_cuda(0)
    scene = sl.Scene((640, 480))
    scene.set_camera_intrinsics(1066.778, 1067.487, 312.9869, 241.3109)
    mesh = sl.Mesh.from_data(fixtures.load_mesh("bunny_mesh"))
    mesh.center_bbox()
    mesh.scale_to_bbox_diagonal(0.5, "order_of_magnitude")
    obj = sl.Object(mesh)
    scene.add_object(obj)
    pose = torch.tensor([[0.0596, 0.8315, -0.5523, -0.0651], [0.4715, 0.4642, 0.7498, -0.06036],
                         [0.8798, -0.3051, -0.3644, 0.80551], [0.0, 0.0, 0.0, 1.0]])
    U, _S, Vh = torch.linalg.svd(pose[:3, :3])
    pose[:3, :3] = U @ Vh
    obj.set_pose(pose)
    torch.manual_seed(0)
    scene.choose_random_light_direction()
    renderer = sl.RenderPass()
    gt_rgb = renderer.render(scene).rgb()[:, :, :3].float().cpu() / 255.0
    assert float(gt_rgb.sum()) > 0
    for param in range(6):
        gt_pose = obj.pose().clone()
        delta_gt = torch.zeros(6)
        delta_gt[param] = 0.01
        obj.set_pose(diff.apply_pose_delta(gt_pose, delta_gt))
        rendered = renderer.render(scene)
        rnd_rgb = rendered.rgb()[:, :, :3].float().cpu() / 255.0
        grad_wrt_img, loss = gaussian_pyramid_grad(gt_rgb, rnd_rgb)
        assert loss > 0
        delta = diff.backpropagate_gradient_to_poses(scene, rendered, grad_wrt_img)
        assert delta.shape == (1, 6)
        assert float(delta[0][param]) > 0, (param, delta)
        obj.set_pose(gt_pose)


def _ref_ext():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import build_ref
    ext = build_ref.load_diff()
    if ext is None:
        pytest.skip("oracle/_ref/diff not built")
    return ext


@pytest.mark.parametrize("seed,H,W", [(0, 64, 96), (1, 480, 640), (2, 32, 32)])
def test_mask_kernels_match_the_reference_cuda_kernels(gpu_ctx, seed, H, W):
    """python/src/diff.cu (the reference's own kernels, compiled for sm_100a into oracle/_ref/diff) run on the same device
    maps as k_sobel_valid / k_dilate and the oracle: all three must agree bit for bit. Sizes are multiples of the
    reference's 32x32 blocks — for other sizes its halo load leaves shared memory uninitialised at the right / bottom
    image edge (diff.cu:36-64: only block-boundary threads load the halo)."""
    ext = _ref_ext()
    sl.init_cuda(0)
    dev = torch.device("cuda", 0)
    rgb, inst, coord4, grad, P, poses, ids = diff_ref.synthetic_inputs(seed, H=H, W=W, n_obj=max(4, H * W // 4000))
    inst_t, depth_t = torch.from_numpy(inst).to(dev), torch.from_numpy(np.ascontiguousarray(coord4[..., 3])).to(dev)
    coord_t = torch.from_numpy(np.ascontiguousarray(coord4[..., :3])).to(dev)
    valid_ref = ext.generate_sobel_valid_mask(inst_t, depth_t)
    valid_mine = diff.generate_sobel_valid_mask(inst_t, depth_t)
    valid_orc = diff_ref.masks(inst, np.ascontiguousarray(coord4[..., 3])).astype(bool)
    assert torch.equal(valid_ref, valid_mine)
    assert np.array_equal(valid_ref.cpu().numpy(), valid_orc)
    assert (~valid_orc).sum() > 0 or H * W <= 1024          # occluding neighbours exist (not in the single-block case)
    for idx in ids[:-1][:8]:
        mask_t = inst_t == int(idx)
        m_ref, c_ref = ext.dilate_object_mask(mask_t, valid_ref, coord_t)
        m_mine, c_mine = diff.dilate_object_mask(mask_t, valid_mine, coord_t)
        m_orc, c_orc = diff_ref.dilate((inst == idx).astype(np.uint8), valid_orc.astype(np.uint8), coord4)
        assert torch.equal(m_ref, m_mine) and np.array_equal(m_ref.cpu().numpy(), m_orc)
        assert torch.equal(c_ref, c_mine) and np.array_equal(c_ref.cpu().numpy(), c_orc)
