"""A/B timing of library options on the C3 workload in ONE process (same box, same clocks): per-stage ms per step for every
combination given on the command line.   python tools/ab_bench.py [--scenes 256] [--steps 3] name=opt:val,opt:val ..."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from stillleben_b200 import abi, lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--config", default="C3")
ap.add_argument("configs", nargs="*", default=["default="])
a = ap.parse_args()
ctx = lib.Context(0)
pool = bench.build_pool()
cfg = bench.CONFIGS[a.config]
scenes = bench.build_scenes(a.config, pool, bench.build_light_map(a.config), 0, a.scenes)
descs = ctx.descs(scenes)
res = lib.Result(ctx, cfg["W"], cfg["H"], a.scenes, abi.TARGETS_SIX)
ctx.set_option(abi.OPT_TIME_KERNELS, 1)
names = ["clear", "bin_count", "scan", "emit", "raster", "shade", "ssao", "post"]
OPTS = {n[4:].lower(): getattr(abi, n) for n in dir(abi) if n.startswith("OPT_")}
defaults = {"overlap": 1, "huge_prepare": 1, "shadow_mask": 1, "direct_max": 128, "warp_max": 4096, "lean_shade": 1, "huge_in_shade": 1}
out = {}
for cfg in a.configs:
    name, _, spec = cfg.partition("=")
    for k, v in defaults.items():
        ctx.set_option(OPTS[k], v)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split(":")
        ctx.set_option(OPTS[k], int(v))
    for _ in range(2):
        ctx.render(scenes, result=res, descs=descs)
    ctx.synchronize(); ctx.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(a.steps):
        ctx.render(scenes, result=res, descs=descs)
    ctx.synchronize(); e1.record(); torch.cuda.synchronize()
    st = ctx.stats()
    row = {n: round(st.last_kernel_ms[i] / a.steps, 3) for i, n in enumerate(names)}
    row["wall_ms"] = round(e0.elapsed_time(e1) / a.steps, 3)
    row["fps"] = round(a.scenes / (row["wall_ms"] / 1e3), 1)
    out[name] = row
    print(name, row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_bench.json"), "w"), indent=1)
