"""Config 4 timing: ms per fused render-and-compare backward (slb_diff_pose_grad) at 640x480 with 21 objects,
next to the CPU oracle of the same computation.  python tools/bench_diff.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import diff_ref  # noqa: E402
from stillleben_b200 import lib  # noqa: E402

ctx = lib.Context(0)
rgb, inst, coord4, grad, P, poses, ids = args = diff_ref.synthetic_inputs(0, H=480, W=640, n_obj=20)
dev = torch.device("cuda", 0)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (rgb, inst, coord4, grad)]
out = torch.zeros((len(ids), 6), device=dev)
Pc, Tc, idv = np.ascontiguousarray(P.T), np.ascontiguousarray(poses.transpose(0, 2, 1)), np.ascontiguousarray(ids, np.int32)


def run():
    ctx.lib.slb_diff_pose_grad(ctx.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), Pc.ctypes.data, Tc.ctypes.data,
                               idv.ctypes.data, len(ids), out.data_ptr(), 480, 640, None)


for _ in range(5):
    run()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    run()
ctx.synchronize()
gpu_ms = (time.perf_counter() - t0) * 1e3 / 200
t0 = time.perf_counter()
ref = diff_ref.oracle_pose_grad(*args)
cpu_ms = (time.perf_counter() - t0) * 1e3
err = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
print(f"pose_grad 640x480, {len(ids)} objects: {gpu_ms:.3f} ms per backward on the GPU, {cpu_ms:.1f} ms CPU oracle (1 core), max rel err {err:.2e}")
