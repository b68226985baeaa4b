"""Does the GPU gain from two render streams running side by side? Two contexts (each its own stream and scratch), two host threads, each
rendering half of the C3 batch, against one context rendering all of it.   python tools/two_ctx_probe.py [--scenes 512]"""
import argparse, json, os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from stillleben_b200 import abi, lib

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=512)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
pool = bench.build_pool()
cfg = bench.CONFIGS["C3"]
scenes = bench.build_scenes("C3", pool, None, 0, a.scenes)
out = {}
def make(n_ctx):
    ctxs = [lib.Context(0) for _ in range(n_ctx)]
    per = a.scenes // n_ctx
    jobs = []
    for i, c in enumerate(ctxs):
        sc = scenes[i * per:(i + 1) * per]
        jobs.append((c, sc, c.descs(sc), lib.Result(c, cfg["W"], cfg["H"], per, abi.TARGETS_SIX)))
    return jobs
def step(jobs):
    def work(j):
        c, sc, d, r = j
        c.render(sc, result=r, descs=d); c.synchronize()
    ts = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in ts: t.start()
    for t in ts: t.join()
for n_ctx in (1, 2, 4):
    jobs = make(n_ctx)
    for _ in range(2): step(jobs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): step(jobs)
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out[n_ctx] = {"ms_per_step": ms, "fps": a.scenes / ms * 1e3}
    print(n_ctx, "contexts:", out[n_ctx], flush=True)
    for c, *_ in jobs: c.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "two_ctx_probe.json"), "w"), indent=1)
