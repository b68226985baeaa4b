"""Frames/s of the other BASELINE.json configs (parity-test cases, not bench lines) for DESIGN.md:
  C2  640x480, 10 objects, PBR + IBL light map, SSAO on, auto exposure      (examples/ycb.py shape)
  C5  1920x1080, 64 objects, PBR + IBL + SSAO + 3 shadow lights
with the per-stage CUDA-event times.   python tools/bench_configs.py [n_scenes]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stillleben_b200 import abi, lib, synth  # noqa: E402
from stillleben_b200.desc import LightMapData  # noqa: E402

NAMES = ["memset", "setup", "scan", "emit", "raster", "shade_store", "ssao", "post"]


def run(ctx, scenes, W, H, mask, label, steps=3, subbatch=None):
    if subbatch:
        ctx.set_option(abi.OPT_MAX_SUBBATCH, subbatch)
    res = lib.Result(ctx, W, H, len(scenes), mask)
    descs = ctx.descs(scenes)
    for _ in range(2):
        ctx.render(scenes, result=res, descs=descs)
    ctx.synchronize()
    ctx.set_option(abi.OPT_TIME_KERNELS, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = np.zeros(8)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        ctx.render(scenes, result=res, descs=descs)
    ctx.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stage += np.array(list(ctx.stats().last_kernel_ms))
    ctx.set_option(abi.OPT_TIME_KERNELS, 0)
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": label, "frames_per_s": len(scenes) / ms * 1e3, "ms_per_frame": ms / len(scenes),
                      "stage_ms_per_frame": {n: round(float(v) / steps / len(scenes), 4) for n, v in zip(NAMES, stage)}}))
    res.close()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ctx = lib.Context(0)
    pool = synth.mesh_pool(21)
    eq, sun = synth.procedural_equirect()
    lm = LightMapData(eq, [sun.tolist()], [[2.0, 1.9, 1.7]])
    c2 = [synth.tabletop_scene(pool, 2000 + s, n_objects=10, light_map=lm, ssao=True, manual_exposure=-1.0) for s in range(n)]
    run(ctx, c2, 640, 480, abi.TARGETS_ALL, "C2 640x480 10obj IBL+SSAO+auto exposure, 8 targets")
    c5 = [synth.tabletop_scene(pool, 3000 + s, n_objects=64, width=1920, height=1080, light_map=None, ssao=True, manual_exposure=1.0,
                               n_lights=3, intrinsics=None) for s in range(max(4, n // 8))]
    run(ctx, c5, 1920, 1080, abi.TARGETS_ALL, "C5 1920x1080 64obj SSAO + 3 shadow lights, 8 targets", subbatch=8)
    c5i = [synth.tabletop_scene(pool, 3000 + s, n_objects=64, width=1920, height=1080, light_map=lm, ssao=True, manual_exposure=1.0,
                                intrinsics=None) for s in range(max(4, n // 8))]
    run(ctx, c5i, 1920, 1080, abi.TARGETS_ALL, "C5 1920x1080 64obj IBL (1 light) + SSAO, 8 targets", subbatch=8)


if __name__ == "__main__":
    main()
