"""Host-side mirror of the reference's `stillleben.diff` render-and-compare surface (config 4 of BASELINE.json).

    from stillleben_b200 import sl, diff       # instead of:  import stillleben as sl; sl.diff....

Same names, argument meanings, return shapes and dtypes as python/stillleben/diff.py and the bridge
python/src/bridge_diff.cpp:13-157:
  generate_sobel_valid_mask(instance_index, depth)            bridge_diff.cpp:13-69   -> bool HxW
  dilate_object_mask(object_mask, valid_mask, coordinates)    bridge_diff.cpp:71-157  -> (bool HxW, float HxWx3)
  compute_image_space_gradients(scene, render_result)         diff.py:73-127          -> (3xHxW, 3xHxW, bool HxW)
  backpropagate_gradient_to_poses(scene, render_result, g)    diff.py:355-523         -> N x 6 float (CPU)
  apply_pose_delta(pose, delta, orthonormalize=True)          diff.py:525-590
All image work runs in the CUDA library through the C ABI (slb_diff_*): the reference's per-object Python loop
(~30 torch ops, two kernel launches and a device synchronisation per object) is ONE fused pass over the pixels
(`slb_diff_pose_grad`). Inputs on the CPU are moved to the context's device first, results follow the
reference's placement (masks / gradients on the input device, pose gradients on the CPU).
"""
import numpy as np
import torch

from . import sl as _sl


def _ctx():
    return _sl._context()


def _dev():
    return torch.device("cuda", _sl._cuda_index)


def _stream():
    """The library work is queued on torch's CURRENT stream of the device, i.e. behind the torch ops that produced the input
    tensors (.to(), .contiguous(), autograd outputs) and ahead of the ones that consume the outputs."""
    return torch.cuda.current_stream(_dev()).cuda_stream


def _check(rc, what):
    if rc != 0:
        ctx = _ctx()
        raise RuntimeError(f"{what}: {ctx.lib.slb_last_error(ctx.h).decode()}")


def generate_sobel_valid_mask(instance_index, depth):
    ctx = _ctx()
    src = instance_index.device
    inst = instance_index.squeeze().to(_dev(), torch.int16).contiguous()
    dep = depth.squeeze().to(_dev(), torch.float32).contiguous()
    H, W = inst.shape
    valid = torch.empty((H, W), dtype=torch.uint8, device=_dev())
    _check(ctx.lib.slb_diff_sobel_valid_mask(ctx.h, inst.data_ptr(), dep.data_ptr(), valid.data_ptr(), H, W, _stream()), "generate_sobel_valid_mask")
    ctx.synchronize()
    return valid.bool().to(src)


def dilate_object_mask(object_mask, valid_mask, coordinates):
    ctx = _ctx()
    src = object_mask.device
    mask = object_mask.squeeze().to(_dev(), torch.uint8).contiguous()
    valid = valid_mask.squeeze().to(_dev(), torch.uint8).contiguous()
    coords = coordinates.to(_dev(), torch.float32)
    H, W = mask.shape
    stride = coords.stride(1) if coords.stride(2) == 1 and coords.stride(0) == W * coords.stride(1) else 0
    if stride < 3:                                   # a view of the 4-channel coordinate buffer is used in place
        coords = coords.contiguous()
        stride = 3
    mask_out = torch.empty((H, W), dtype=torch.uint8, device=_dev())
    coords_out = torch.empty((H, W, 3), dtype=torch.float32, device=_dev())
    _check(ctx.lib.slb_diff_dilate_object_mask(ctx.h, mask.data_ptr(), valid.data_ptr(), coords.data_ptr(), stride, mask_out.data_ptr(),
                                               coords_out.data_ptr(), H, W, _stream()), "dilate_object_mask")
    ctx.synchronize()
    return mask_out.bool().to(src), coords_out.to(src)


def compute_image_space_gradients(scene, render_result):
    """Gradient of intensity w.r.t. the 2D pixel position: central differences of rgb / 255 scaled by W/4 and H/4,
    negated, zero where the sobel-valid mask is false (diff.py:73-127)."""
    rgb = render_result.rgb()[:, :, :3]
    device = rgb.device
    rgb_float = rgb.permute(2, 0, 1).float() / 255.0
    _c, _h, _w = rgb_float.shape
    pad = torch.nn.functional.pad
    px = pad(rgb_float, (1, 1, 0, 0))
    py = pad(rgb_float, (0, 0, 1, 1))
    grad_x = -(px[:, :, 2:] - px[:, :, :-2]) / (2 / _w * 2)
    grad_y = -(py[:, 2:, :] - py[:, :-2, :]) / (2.0 / _h * 2)
    valid = generate_sobel_valid_mask(render_result.instance_index().squeeze(), render_result.depth().squeeze()).to(device)
    grad_x[:, ~valid] = 0
    grad_y[:, ~valid] = 0
    return grad_x, grad_y, valid


def backpropagate_gradient_to_poses(scene, render_result, grad_objective_wrt_rnd_img, visualize_grad=False):
    """N x 6 gradient of the objective w.r.t. the locally linearised poses (alpha, beta, gamma, a, b, c) of
    scene.objects, from the 3xHxW gradient w.r.t. the rendered image (diff.py:355-523)."""
    ctx = _ctx()
    dev = _dev()
    objects = list(scene.objects)
    out = torch.zeros(len(objects), 6)
    if not objects:
        return out
    rgb = render_result.rgb().to(dev).contiguous()
    inst = render_result.instance_index().to(dev, torch.int16).contiguous()
    coord = render_result.coordDepth().to(dev, torch.float32).contiguous()
    H, W = rgb.shape[0], rgb.shape[1]
    g = grad_objective_wrt_rnd_img.to(dev, torch.float32).contiguous()
    if tuple(g.shape) != (3, H, W):
        raise ValueError("grad_objective_wrt_rnd_img must be a 3xHxW tensor")
    P = np.ascontiguousarray(np.asarray(scene.projection_matrix().cpu(), np.float32).T)            # column-major for the ABI
    poses = np.ascontiguousarray(np.stack([np.asarray(o.pose().cpu(), np.float32).T for o in objects]))
    ids = np.asarray([o.instance_index for o in objects], np.int32)
    d_out = torch.empty((len(objects), 6), dtype=torch.float32, device=dev)
    _check(ctx.lib.slb_diff_pose_grad(ctx.h, rgb.data_ptr(), inst.data_ptr(), coord.data_ptr(), g.data_ptr(), P.ctypes.data,
                                      poses.ctypes.data, ids.ctypes.data, len(objects), d_out.data_ptr(), H, W, _stream()),
           "backpropagate_gradient_to_poses")
    ctx.synchronize()
    return d_out.cpu()


def apply_pose_delta(pose, delta, orthonormalize=True):
    """pose @ [[1,-g,b,a],[g,1,-al,b'],[-b,al,1,c],[0,0,0,1]], optionally re-orthonormalised with an SVD
    (diff.py:525-590); batched (Bx4x4, Bx6) or single."""
    batched = pose.dim() == 3
    if batched:
        assert delta.dim() == 2
    else:
        assert delta.dim() == 1
        pose, delta = pose.unsqueeze(0), delta.unsqueeze(0)
    device = pose.device
    pose, delta = pose.cpu().float(), delta.cpu().float()
    B = pose.size(0)
    dm = torch.zeros(B, 4, 4)
    dm[:, 0, 0] = 1.0; dm[:, 0, 1] = -delta[:, 2]; dm[:, 0, 2] = delta[:, 1]
    dm[:, 1, 0] = delta[:, 2]; dm[:, 1, 1] = 1.0; dm[:, 1, 2] = -delta[:, 0]
    dm[:, 2, 0] = -delta[:, 1]; dm[:, 2, 1] = delta[:, 0]; dm[:, 2, 2] = 1.0
    dm[:, :3, 3] = delta[:, 3:]
    dm[:, 3, 3] = 1.0
    new_poses = torch.matmul(pose, dm)
    if orthonormalize:
        for b in range(B):
            U, _S, Vh = torch.linalg.svd(new_poses[b, :3, :3])
            new_poses[b, :3, :3] = torch.matmul(U, Vh)
    if not batched:
        new_poses = new_poses[0]
    return new_poses.to(device)
