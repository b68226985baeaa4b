"""One invocation of each auxiliary kernel family at 640x480 x 64 images (for `ncu` captures under profiles/):
camera model (k_cam_stage1/2), PNG encoder (k_png_rows / finalize / gather), pose-gradient backward (k_pose_grad)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import diff_ref  # noqa: E402
from stillleben_b200 import abi, camera_model, image_saver, sl, synth  # noqa: E402

sl.init_cuda(0)
ctx = sl._context()
pool = synth.mesh_pool(21)
scenes = [synth.tabletop_scene(pool, 1000 + s) for s in range(64)]
res = ctx.render(scenes, target_mask=abi.TARGETS_SIX)
ctx.synchronize()
rgb = torch.from_numpy(res.numpy(abi.TARGET_RGB)).cuda()
for _ in range(2):
    camera_model.process_batch(rgb, [camera_model.random_parameters() for _ in range(64)])
    image_saver.encode_batch(rgb)
rgb1, inst, coord4, grad, P, poses, ids = diff_ref.synthetic_inputs(0, H=480, W=640, n_obj=20)
t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rgb1, inst, coord4, grad)]
out = torch.zeros((len(ids), 6), device="cuda")
Pc, Tc, idv = np.ascontiguousarray(P.T), np.ascontiguousarray(poses.transpose(0, 2, 1)), np.ascontiguousarray(ids, np.int32)
for _ in range(2):
    ctx.lib.slb_diff_pose_grad(ctx.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), Pc.ctypes.data, Tc.ctypes.data,
                               idv.ctypes.data, len(ids), out.data_ptr(), 480, 640, None)
ctx.synchronize()
print("ok")
