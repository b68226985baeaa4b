#!/usr/bin/env python
"""bench.py — multi-target frames/sec of the RenderPass::render hot path (BASELINE.json metric).

Workload (config C3 of SURVEY.md §8d): a FIXED batch of 1024 random table-top scenes, 20 objects each
from a pool of 21 stand-in meshes (16 384 triangles each), 640x480, six render targets (40 B/px), one
shadow-casting directional light + ambient, manual exposure 1, SSAO off.  A "step" renders the whole
batch once; with N GPUs the batch is sharded by scene index (strong scaling, no steady-state collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : frames/s with the scene descriptors in host memory and all outputs left in HBM (device-timed).
`e2e`    : frames/s through slb_render_batch_host — descriptors host->device AND all six targets
           device->host (pinned) inside the timed region.
`--impl reference` times the CPU implementation of the same path (the OpenMP oracle: the reference's GL
path cannot be built here) on all host cores over a bounded sample of the same scenes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCENES = 1024
N_OBJECTS = 20
W, H = 640, 480
POOL = 21
BYTES_PER_PX = 40
METRIC = "multi-target frames/sec at 640x480x20obj"


def build_scenes(pool, lo, hi):
    from stillleben_b200 import synth
    return [synth.tabletop_scene(pool, 1000 + s, n_objects=N_OBJECTS, width=W, height=H) for s in range(lo, hi)]


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max((float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_fps(pool, n_sample, n_threads=0, repeats=1):
    """Oracle (CPU restatement of the reference path) on `n_sample` scenes of the workload -> frames/s."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    assets = ou.OracleAssets()
    scenes = build_scenes(pool, 0, n_sample)
    ou.render(scenes[0], assets, n_threads=n_threads, want_hdr=False)     # builds oracle textures, warms caches
    t0 = time.perf_counter()
    for _ in range(repeats):
        for sc in scenes:
            ou.render(sc, assets, n_threads=n_threads, want_hdr=False)
    dt = time.perf_counter() - t0
    return n_sample * repeats / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from stillleben_b200 import synth
    pool = synth.mesh_pool(POOL)
    cores = os.cpu_count() or 1
    sample = 4
    times = []
    for i in range(args.warmup + args.steps):
        fps, dt = cpu_reference_fps(pool, sample, n_threads=cores)   # explicit: torchrun exports OMP_NUM_THREADS=1
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    fps = sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C3: {N_SCENES} scenes x {N_OBJECTS} objects, {W}x{H}, six targets", "sample_scenes_per_step": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} scenes of the workload per step, OpenMP oracle on all host cores"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=N_SCENES)
    ap.add_argument("--subbatch", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from stillleben_b200 import abi, lib, synth
    from stillleben_b200 import dist as sdist

    rank, world, local = sdist.env_rank_world()
    numa_node = sdist.bind_to_gpu_numa_node(local) if world > 1 else None   # host buffers next to the rank's GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # ---- load: rank 0 builds the pool, ONE broadcast of the asset arena ----
    # NCCL logs its version banner to stdout when the first communicator comes up: keep fd 1 pointed at stderr until
    # then, so that rank 0's stdout carries exactly ONE JSON line.
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        pool = synth.mesh_pool(POOL) if rank == 0 else None
        pool = sdist.broadcast_meshes(pool, 0, dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    ctx = lib.Context(local)
    if args.subbatch:
        ctx.set_option(abi.OPT_MAX_SUBBATCH, args.subbatch)
    lo, hi = sdist.shard_range(args.scenes, rank, world)
    n_local = hi - lo
    scenes = build_scenes(pool, lo, hi)
    descs = ctx.descs(scenes)                       # host-side scene descriptors (what a caller hands over)
    result = lib.Result(ctx, W, H, n_local, abi.TARGETS_SIX)
    tstream = torch.cuda.Stream(device=dev)            # work is queued on this (non-default) stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step():
        ctx.render(scenes, result=result, stream=stream, descs=descs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.set_option(abi.OPT_TIME_KERNELS, 1)
    st0 = ctx.stats()
    launches0 = st0.kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(8)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    stage_ms += np.array(list(ctx.stats().last_kernel_ms))    # stage events of the whole timed region, read after it
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = ctx.stats().kernel_launches - launches0
    ctx.set_option(abi.OPT_TIME_KERNELS, 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = args.scenes * args.steps / (ms_max * 1e-3)

    # ---- end to end: host descriptors in, all six targets back in pinned host memory ----
    e2e = None
    if not args.no_e2e:
        chunk = min(1024, n_local)                                          # one call per step and rank
        host = {}
        for tgt, (dt_, ch) in enumerate(abi.TARGET_FORMATS):
            if abi.TARGETS_SIX & (1 << tgt):
                host[tgt] = ctx.host_alloc((chunk, H, W, ch), dt_)          # page-locked (slb_host_alloc)
        host_ptrs = {k: v.ctypes.data for k, v in host.items()}
        chunk_descs = [ctx.descs(scenes[a:a + chunk]) for a in range(0, n_local, chunk)]
        d2h = n_local * W * H * BYTES_PER_PX
        h2d0 = ctx.stats().bytes_h2d

        def e2e_step():
            for cd in chunk_descs:
                ctx.render_host(cd, host_ptrs, abi.TARGETS_SIX)

        e2e_step()
        h2d = ctx.stats().bytes_h2d - h2d0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        t = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": args.scenes * args.steps / float(t.item()), "unit": "frames/s", "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "note": f"slb_render_batch_host, page-locked host buffers (slb_host_alloc), {chunk}-scene calls; bound by the device->host link "
                       "(12.29 MB per frame; 57 GB/s measured D2H on this pool = 4660 frames/s per GPU)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the shade + multi-render-target store kernel (CUDA events on the launching stream) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    sub = ctx.stats()
    traffic = None                                        # DRAM bytes of one k_shade launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = (tj["dram_read_bytes_per_launch"] + tj["dram_write_bytes_per_launch"]) / tj["frames_per_launch"]
    n_sub = -(-n_local // (args.subbatch or 64))
    shade_ms_per_launch = stage_ms[5] / (args.steps * n_sub)
    frames_per_launch = n_local / n_sub
    alg_bytes = frames_per_launch * W * H * BYTES_PER_PX
    achieved = alg_bytes / (shade_ms_per_launch * 1e-3) / 1e9 if shade_ms_per_launch > 0 else 0.0
    names = ["shadow", "bin_count", "scan", "bin_emit", "raster", "shade_store", "ssao", "post"]
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C3: fixed batch of {args.scenes} scenes x {N_OBJECTS} objects ({POOL}-mesh pool, 16384 tris each), "
                                   f"{W}x{H}, six targets (40 B/px), 1 shadow light + ambient, exposure 1, SSAO off",
                       "scenes_per_gpu": n_local, "subbatch": args.subbatch or 64, "rank0_numa_node": numa_node,
                       "l2": "outputs per step (%.1f GB) exceed L2; no flush needed" % (n_local * W * H * BYTES_PER_PX / 1e9)},
            "clocks": sampler.summary(), "gpu_launches": int(launches),
            "stage_ms_per_step": {n: float(v / args.steps) for n, v in zip(names, stage_ms)},
            "roofline": {"bound": "hbm", "kernel": "k_shade (shade + MRT store)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic * frames_per_launch if traffic else None,
                         "traffic_source": "profiles/traffic.json (ncu --set full, dram read + write bytes per k_shade launch, scaled to this launch size)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": shade_ms_per_launch},
            "triangles_per_frame": int(sub.triangles_submitted / max(1, sub.frames_rendered))}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:
        pool0 = pool
        cores = os.cpu_count() or 1
        sample = 24
        fps, dt = cpu_reference_fps(pool0, sample, n_threads=cores)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} scenes of the workload, OpenMP oracle ({dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
