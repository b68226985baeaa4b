"""Config 4 timing: ms per fused render-and-compare backward (slb_diff_pose_grad) at 640x480 with 21 objects, next to
(a) the reference's OWN CUDA kernels for the mask part of the same computation (oracle/_ref/diff: python/src/diff.cu +
bridge_diff.cpp compiled from the reference — generate_sobel_valid_mask once + dilate_object_mask per object, the calls
diff.py:355-523 makes before its per-object tensor program) and (b) the CPU oracle.  python tools/bench_diff.py"""
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
ref_ms = None
sys.path.insert(0, os.path.join(ROOT, "oracle"))
try:
    import build_ref
    ext = build_ref.load_diff()
except Exception as e:      # noqa: BLE001
    ext, ref_note = None, f"reference extension not available ({e})"
if ext is not None:
    inst_t = t[1]
    depth_t = t[2][..., 3].contiguous()
    coord_t = t[2][..., :3].contiguous()
    masks = [inst_t == int(i) for i in ids]

    def run_ref():
        valid = ext.generate_sobel_valid_mask(inst_t, depth_t)
        for m in masks:
            ext.dilate_object_mask(m, valid, coord_t)

    for _ in range(3):
        run_ref()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        run_ref()
    torch.cuda.synchronize()
    ref_ms = (time.perf_counter() - t0) * 1e3 / 50
    ref_note = f"{ref_ms:.3f} ms for the reference's own mask kernels alone (1 + {len(ids)} launches of diff.cu, before diff.py's per-object tensor program)"
print(ref_note)
print(f"pose_grad 640x480, {len(ids)} objects: {gpu_ms:.3f} ms per backward on the GPU, {cpu_ms:.1f} ms CPU oracle (1 core), max rel err {err:.2e}")
