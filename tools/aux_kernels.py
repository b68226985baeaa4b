"""One invocation of each auxiliary kernel family at 640x480 x 64 images (for `ncu` captures under profiles/):
camera model (k_cam_stage1/2), PNG encoder (k_png_rows_warp / finalize / gather), JPEG encoder (k_jpeg_*), pose-gradient
backward (k_pose_grad). Without ncu it also prints wall-clock throughputs (CUDA events, 10 repetitions)."""
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
    image_saver.encode_batch_jpeg(rgb)
rgb1, inst, coord4, grad, P, poses, ids = diff_ref.synthetic_inputs(0, H=480, W=640, n_obj=20)
t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rgb1, inst, coord4, grad)]
out = torch.zeros((len(ids), 6), device="cuda")
Pc, Tc, idv = np.ascontiguousarray(P.T), np.ascontiguousarray(poses.transpose(0, 2, 1)), np.ascontiguousarray(ids, np.int32)
for _ in range(2):
    ctx.lib.slb_diff_pose_grad(ctx.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), Pc.ctypes.data, Tc.ctypes.data,
                               idv.ctypes.data, len(ids), out.data_ptr(), 480, 640, None)
ctx.synchronize()
if os.environ.get("SLB_AUX_TIMING", "1") == "1" and "--no-timing" not in sys.argv:
    import json

    def timed(fn, reps=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # a NON-default stream: the legacy default stream's handle is 0, which the ABI reads as "the context's own stream" —
    # the kernels would then not run on the stream the timing events are recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    n = rgb.shape[0]
    # device-side encode only (the files stay in HBM): the ABI calls behind encode_batch / encode_batch_jpeg
    png_out = torch.empty((n, ctx.lib.slb_png_bound(480, 640, 4, 1)), dtype=torch.uint8, device="cuda")
    jpg_out = torch.empty((n, 640 * 480 * 2), dtype=torch.uint8, device="cuda")
    sizes = torch.empty((n,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms_png = timed(lambda: ctx.lib.slb_png_encode(ctx.h, rgb.data_ptr(), n, 480, 640, 4, 1, png_out.data_ptr(), png_out.shape[1], sizes.data_ptr(), st))
    png_kb = float(sizes.float().mean()) / 1e3
    ms_jpg = timed(lambda: ctx.lib.slb_jpeg_encode(ctx.h, rgb.data_ptr(), n, 480, 640, 4, 80, jpg_out.data_ptr(), jpg_out.shape[1], sizes.data_ptr(), st))
    jpg_kb = float(sizes.float().mean()) / 1e3
    ms_cam = timed(lambda: camera_model.process_batch(rgb, [camera_model.random_parameters() for _ in range(n)]), 3)
    print(json.dumps({"images": n, "shape": "640x480 RGBA8 rendered frames",
                      "png_encode": {"ms_per_batch": round(ms_png, 3), "images_per_s": round(n / ms_png * 1e3), "mean_file_kb": round(png_kb, 1)},
                      "jpeg_encode_q80": {"ms_per_batch": round(ms_jpg, 3), "images_per_s": round(n / ms_jpg * 1e3), "mean_file_kb": round(jpg_kb, 1)},
                      "camera_model_python_wrapper": {"ms_per_batch": round(ms_cam, 3), "images_per_s": round(n / ms_cam * 1e3)}}))
print("ok")
