#!/usr/bin/env python
"""bench.py — multi-target frames/sec of the RenderPass::render hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C2|C5]

Workloads (`config.workload` in the JSON line names the one that ran, C3 is the default and the headline):
  C3  a FIXED batch of 1024 random table-top scenes x 20 objects, 640x480, six render targets (40 B/px), one shadow-casting
      directional light + ambient, manual exposure 1, SSAO off
  C2  examples/ycb.py's shape: 512 scenes x 10 objects, 640x480, ycb.py intrinsics, PBR + IBL light map (sun from the map),
      SSAO on, auto exposure
  C5  256 scenes x 64 objects, 1920x1080, PBR + IBL + SSAO + 3 shadow lights
Mesh pool for all three: 20 procedural stand-ins of 16 384 triangles (YCB google_16k scale) + the reference's own Stanford
bunny (69 451 triangles, tests/golden/bunny_mesh.npz), shared by all scenes. A "step" renders the whole batch once; with N
GPUs the batch is sharded by scene index (strong scaling, no steady-state collective; ONE broadcast of meshes + IBL maps).

`value`  : frames/s with the scene descriptors in host memory and all outputs left in HBM (device-timed, max over ranks).
`e2e`    : frames/s through slb_render_batch_host — descriptors host->device AND all six targets device->host (pinned)
           inside the timed region; `d2h_link_gbs` is a pinned device->host copy of the same size class measured in the same
           process right before it (the ceiling of that number).
`roofline`: the shade + MRT-store kernel against the measured HBM peak; `binner`: the set-up / bin kernel.
`--impl reference` times the CPU implementation of the same path (the OpenMP oracle, one scene per host thread — the
reference's GL path cannot be built here) over a bounded sample of the same scenes.
"""
import argparse
import concurrent.futures
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POOL = 21
BYTES_PER_PX = 40            # six targets; main() switches both to the eight-target set (88 B/px) for --targets eight
TARGET_MASK = 0x1F
TARGETS_TEXT = "six targets (40 B/px)"
METRIC = "multi-target frames/sec at 640x480x20obj"
YCB = (1066.778, 1067.487, 312.9869, 241.3109)
CONFIGS = {
    "C3": dict(n_scenes=1024, n_objects=20, W=640, H=480, n_lights=1, ibl=False, ssao=False, exposure=1.0, subbatch=64, seed0=1000,
               intrinsics=YCB, e2e_chunk=1024, text="1 shadow light + ambient, exposure 1, SSAO off"),
    "C2": dict(n_scenes=512, n_objects=10, W=640, H=480, n_lights=1, ibl=True, ssao=True, exposure=-1.0, subbatch=64, seed0=2000,
               intrinsics=YCB, e2e_chunk=512, text="PBR + IBL light map (sun light from the map), SSAO on, auto exposure"),
    "C5": dict(n_scenes=256, n_objects=64, W=1920, H=1080, n_lights=3, ibl=True, ssao=True, exposure=1.0, subbatch=8, seed0=3000,
               intrinsics=None, e2e_chunk=32, text="PBR + IBL + SSAO + 3 shadow lights, exposure 1"),
}


def workload_string(name, n_scenes):
    c = CONFIGS[name]
    return (f"{name}: fixed batch of {n_scenes} scenes x {c['n_objects']} objects (pool: 20 procedural 16384-triangle stand-ins + the "
            f"Stanford bunny, 69451 triangles), {c['W']}x{c['H']}, {TARGETS_TEXT}, {c['text']}")


def build_pool():
    from stillleben_b200 import abi, synth
    from stillleben_b200.desc import ImageData, MaterialData, MeshData
    pool = synth.mesh_pool(POOL - 1)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bunny_mesh.npz"))     # the reference's tests/stanford_bunny after ITS consolidation
    mats = [MaterialData(tuple(float(x) for x in r[0:4]), tuple(float(x) for x in r[4:8]), float(r[8]), float(r[9]), int(r[10]), int(r[11]),
                         int(r[12]), int(r[13]), int(r[14])) for r in z["materials"]]
    s = z["image0_sampler"]
    images = [ImageData(np.ascontiguousarray(z["image0"]), int(s[0]), int(s[1]), int(s[2]), int(s[3]))]
    pool.append(MeshData(np.ascontiguousarray(z["vertices"]).view(abi.VERTEX_DTYPE).reshape(-1), z["indices"],
                         [tuple(int(x) for x in sm) for sm in z["submeshes"]], mats, images, z["bbox_min"].astype(np.float32),
                         z["bbox_max"].astype(np.float32), "stanford_bunny"))
    return pool


def build_light_map(name):
    from stillleben_b200 import synth
    from stillleben_b200.desc import LightMapData
    c = CONFIGS[name]
    if not c["ibl"]:
        return None
    eq, sun = synth.procedural_equirect()
    dirs, cols = [sun.tolist()], [[2.0, 1.9, 1.7]]
    for k in range(1, c["n_lights"]):                  # Light1 / Light2 of an sIBL file (light_map.cpp:104-152)
        d = np.array([0.5 * (-1) ** k, 0.4 * k - 0.3, -1.0])
        dirs.append((d / np.linalg.norm(d)).tolist())
        cols.append([0.8, 0.8, 0.9])
    return LightMapData(eq, dirs, cols)


def build_scenes(name, pool, light_map, lo, hi):
    from stillleben_b200 import synth
    c = CONFIGS[name]
    return [synth.tabletop_scene(pool, c["seed0"] + s, n_objects=c["n_objects"], width=c["W"], height=c["H"], light_map=light_map,
                                 ssao=c["ssao"], manual_exposure=c["exposure"], n_lights=c["n_lights"], intrinsics=c["intrinsics"])
            for s in range(lo, hi)]


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


class CpuBaseline:
    """The CPU restatement of the path (oracle) on the first `n_sample` scenes of the workload, ONE SCENE PER HOST THREAD
    (scenes are independent, exactly how the work shards across GPUs): descriptors, oracle assets and output arrays are built
    once, outside the timed calls."""

    def __init__(self, name, pool, n_sample, light_map):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ctypes as C
        import oracle_util as ou
        from stillleben_b200 import abi
        from stillleben_b200.desc import DescBatch
        self.ou, self.C, self.abi = ou, C, abi
        self.assets = ou.OracleAssets()
        if light_map is not None:          # the light-map precompute is load-time work on both sides: reduced sizes keep the CPU set-up short
            self.assets.lightmap_sizes = (128, 16, 32, 64, 64)
        self.scenes = build_scenes(name, pool, light_map, 0, n_sample)
        self.jobs = []
        for sc in self.scenes:
            batch = DescBatch([sc], self.assets.handle_of)
            ptrs = (C.c_void_p * abi.NUM_TARGETS)()
            keep = []
            for t, (dt, ch) in enumerate(abi.TARGET_FORMATS):
                if TARGET_MASK & (1 << t):
                    a = np.zeros((sc.height, sc.width, ch), dt)
                    keep.append(a)
                    ptrs[t] = a.ctypes.data
            self.jobs.append((batch, ptrs, keep))
        self.L = ou.lib()
        self.cores = os.cpu_count() or 1
        self._one(self.jobs[0])                    # warms caches, builds the plane

    def _one(self, job):
        rc = self.L.orc_render(job[0].ptr, None, job[1], None, 1)      # one OpenMP thread inside: the parallelism is across scenes
        assert rc == 0

    def run(self):
        """-> (frames/s, seconds)"""
        t0 = time.perf_counter()
        with concurrent.futures.ThreadPoolExecutor(self.cores) as ex:   # ctypes releases the GIL during the call
            list(ex.map(self._one, self.jobs))
        dt = time.perf_counter() - t0
        return len(self.jobs) / dt, dt


class GlBaseline:
    """The reference's own shaders on a CPU OpenGL implementation: oracle/_ref/glref (the reference's GLSL text, #define header and
    uniform setters compiled in by oracle/build_ref.py; Mesa llvmpipe = the libGL bundled with Nsight Compute in this image) — what
    SURVEY 8(d) names as the preferred CPU baseline. Scene-parallel like the port: `workers` glref processes, each with llvmpipe's
    rasteriser in its own thread only (LP_NUM_THREADS=0), render their share of the first `n_sample` scenes of the workload.
    Timed region = RenderPass::render per frame (shadow passes ... tone map, glFinish) as the harness clocks it; context creation,
    shader compilation, mesh / texture upload and read-back are outside, as asset building is for the port."""

    def __init__(self, name, pool, n_sample, light_map):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import tempfile
        import glref_util
        why = glref_util.available()
        if why:
            raise RuntimeError(why)
        self.g = glref_util
        self.cores = os.cpu_count() or 1
        self.tmp = tempfile.TemporaryDirectory()
        scenes = build_scenes(name, pool, light_map, 0, n_sample)
        self.n = len(scenes)
        self.workers = min(self.cores, self.n)
        self.jobs = [[] for _ in range(self.workers)]
        for k, sc in enumerate(scenes):
            src, dst = os.path.join(self.tmp.name, f"s{k}.bin"), os.path.join(self.tmp.name, f"o{k}.bin")
            glref_util.dump(sc, src, None, (128, 16, 32, 64))          # light-map sizes as the port's baseline (load-time work, untimed)
            self.jobs[k % self.workers] += [src, dst]
        self.env = glref_util.gl_env({"GLREF_TIMING": "1", "LP_NUM_THREADS": "0"})

    def _worker(self, files):
        p = subprocess.run([self.g.GLREF] + files, env=self.env, capture_output=True, text=True, timeout=1800)
        if p.returncode != 0:
            raise RuntimeError("glref failed: " + p.stderr[-2000:])
        ms = [float(ln.split()[-1]) for ln in p.stderr.splitlines() if "render_ms" in ln]
        assert len(ms) == len(files) // 2, p.stderr[-2000:]
        return sum(ms) * 1e-3

    def run(self):
        """-> (frames/s, seconds): the slowest worker's summed render time bounds the step"""
        with concurrent.futures.ThreadPoolExecutor(self.workers) as ex:
            busy = list(ex.map(self._worker, self.jobs))
        dt = max(busy)
        return self.n / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    n_scenes = args.scenes or c["n_scenes"]
    cores = os.cpu_count() or 1
    used = min(cores, 64)            # one scene per host thread / glref process; bounded so a very wide host does not need 100+ GL contexts
    sample = min(n_scenes, max(8, used if c["W"] <= 640 else used // 2))
    pool, light_map = build_pool(), build_light_map(args.config)

    def timed(base, steps, warmup):
        times = []
        for i in range(warmup + steps):
            _, dt = base.run()
            if i >= warmup:
                times.append(dt)
        dt = sum(times) / len(times)
        return sample / dt, dt

    # the port (OpenMP oracle): always available
    port_fps, port_dt = timed(CpuBaseline(args.config, pool, sample, light_map), args.steps, args.warmup)
    kind, fps, dt = "port", port_fps, port_dt
    what = f"scenes 0..{sample - 1} of the workload per step, OpenMP oracle, one scene per host thread"
    extra = {}
    # the reference's own shaders on Mesa llvmpipe (oracle/_ref/glref), when the image has them: this is then the line's value. A step
    # costs seconds per frame and core, so it gets fewer steps (the frames are the same every step; there is no noise to average out).
    if args.ref_arm != "port":
        try:
            gl_steps, gl_warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
            fps, dt = timed(GlBaseline(args.config, pool, sample, light_map), gl_steps, gl_warmup)
            kind = "reference"
            what = (f"scenes 0..{sample - 1} of the workload per step; the reference's GLSL programs + uniform setters (oracle/_ref/glref) on Mesa llvmpipe, "
                    f"one process per scene slot, single-threaded rasteriser, RenderPass::render region only ({gl_steps} steps after {gl_warmup} warm-up; "
                    f"the render loop around the shaders is restated as GL calls, DESIGN 5)")
            extra = {"port": {"value": port_fps, "unit": "frames/s", "ms_per_step": port_dt * 1e3,
                              "sample": "same scenes, OpenMP oracle, one scene per host thread"}}
        except Exception as e:                      # no Mesa libGL / harness on this machine: the port stands
            extra = {"reference_gl_unavailable": str(e)[:300]}
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.config, n_scenes)},
            "cpu_baseline": dict({"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": what}, **extra),
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_d2h(torch, dev, n_bytes=1 << 28, reps=4):
    """Pinned device->host copy bandwidth of this process / GPU, GB/s (the ceiling of the end-to-end number)."""
    src = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def measure_d2h_sustained(torch, dev, barrier, n_bytes=1 << 28, reps=8):
    """Pinned device->host bandwidth of this rank WHILE every rank copies (barrier, then `reps` back-to-back copies timed as
    one interval): what a rank's link delivers under the contention the end-to-end path sees. GB/s."""
    src = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * n_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scenes", type=int, default=0, help="batch size (default: the config's)")
    ap.add_argument("--subbatch", type=int, default=0)
    ap.add_argument("--targets", default="six", choices=["six", "eight"], help="six (40 B/px, the headline) or all eight targets (88 B/px)")
    ap.add_argument("--ref-arm", default="auto", choices=["auto", "gl", "port"],
                    help="--impl reference: the reference's shaders on Mesa llvmpipe when available (auto / gl), or the OpenMP oracle port only")
    ap.add_argument("--overlap", default="on", choices=["on", "off"], help="two-stream sub-batch pipeline in the timed region (SLB_OPT_OVERLAP)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.targets == "eight":
        global BYTES_PER_PX, TARGET_MASK, TARGETS_TEXT
        BYTES_PER_PX, TARGET_MASK, TARGETS_TEXT = 88, 0xFF, "all eight targets (88 B/px)"
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from stillleben_b200 import abi, lib
    from stillleben_b200 import dist as sdist

    c = CONFIGS[args.config]
    W, H = c["W"], c["H"]
    n_scenes = args.scenes or c["n_scenes"]
    subbatch = args.subbatch or c["subbatch"]
    rank, world, local = sdist.env_rank_world()
    numa_node = sdist.bind_to_gpu_numa_node(local) if world > 1 else None   # host buffers next to the rank's GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = lib.Context(local)
    ctx.set_option(abi.OPT_MAX_SUBBATCH, subbatch)
    # ---- load: rank 0 builds the pool and runs the IBL precompute, ONE broadcast of the asset arena (meshes + IBL maps) ----
    # NCCL logs its version banner to stdout when the first communicator comes up: keep fd 1 pointed at stderr until
    # then, so that rank 0's stdout carries exactly ONE JSON line.
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        pool, lms = None, None
        if rank == 0:
            pool = build_pool()
            lm = build_light_map(args.config)
            lms = [sdist.precompute_light_map(ctx, lm)] if (lm is not None and world > 1) else ([lm] if lm is not None else [])
        pool, lms = sdist.broadcast_assets(pool, lms, 0, dev)
        light_map = lms[0] if lms else None
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    lo, hi = sdist.shard_range(n_scenes, rank, world)
    n_local = hi - lo
    scenes = build_scenes(args.config, pool, light_map, lo, hi)
    descs = ctx.descs(scenes)                       # host-side scene descriptors (what a caller hands over)
    result = lib.Result(ctx, W, H, n_local, TARGET_MASK)
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

    two_streams = args.overlap == "on"
    ctx.set_option(abi.OPT_OVERLAP, 1 if two_streams else 0)
    for _ in range(args.warmup):
        step()
    st0 = ctx.stats()
    launches0 = st0.kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = ctx.stats().kernel_launches - launches0
    # ---- stage pass: the same steps again on ONE stream with a CUDA event pair around every stage. With the two-stream pipeline
    # (SLB_OPT_OVERLAP) a stage's event interval on its stream also contains the other stream's kernels (the block dispatcher drains one
    # grid before the next), so per-kernel durations - stage_ms_per_step, roofline, binner - are taken here, right after the timed
    # region, same process, same inputs, same clocks; `pipeline.ms_per_step_single_stream` is this pass's own step time.
    stage_steps = max(1, min(args.steps, 5))
    ctx.set_option(abi.OPT_OVERLAP, 0)
    ctx.set_option(abi.OPT_TIME_KERNELS, 1)
    step()
    torch.cuda.synchronize()
    ctx.stats()                                               # drains the events of the settling step
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(stage_steps):
        step()
    es1.record()
    torch.cuda.synchronize()
    stage_ms = np.array(list(ctx.stats().last_kernel_ms))     # stage events of the stage pass, read after it
    serial_ms_per_step = es0.elapsed_time(es1) / stage_steps
    ctx.set_option(abi.OPT_TIME_KERNELS, 0)
    ctx.set_option(abi.OPT_OVERLAP, 1 if two_streams else 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_scenes * args.steps / (ms_max * 1e-3)

    # ---- end to end: host descriptors in, all six targets back in pinned host memory ----
    e2e = None
    if not args.no_e2e:
        # The host-buffer path is bound by each rank's device->host link, and on a multi-GPU box the links are NOT equal while all
        # ranks copy (profiles/r02_multi_gpu.md: 11.6 vs 18.2 GB/s at 8 GPUs): the fixed batch is sharded in proportion to the
        # sustained bandwidth every rank measures under that contention, so that all ranks finish together.
        e2e_scenes, shard_sizes = scenes, [n_local]
        if world > 1:
            d2h_gbs = measure_d2h_sustained(torch, dev, barrier)
            bw = torch.tensor([d2h_gbs], dtype=torch.float64, device=dev)
            allbw = [torch.zeros_like(bw) for _ in range(world)]
            dist.all_gather(allbw, bw)
            weights = [float(x.item()) for x in allbw]
            shard_sizes = sdist.shard_sizes_weighted(n_scenes, weights)
            lo_e, hi_e = sdist.shard_range_weighted(n_scenes, rank, weights)
            e2e_scenes = build_scenes(args.config, pool, light_map, lo_e, hi_e)
        else:
            d2h_gbs = measure_d2h(torch, dev)
        n_e2e = len(e2e_scenes)                                              # (the device-resident leg above keeps the equal shards)
        chunk = max(1, min(c["e2e_chunk"], n_e2e))                         # scenes per call
        host = {}
        for tgt, (dt_, ch) in enumerate(abi.TARGET_FORMATS):
            if TARGET_MASK & (1 << tgt):
                host[tgt] = ctx.host_alloc((chunk, H, W, ch), dt_)          # page-locked (slb_host_alloc)
        host_ptrs = {k: v.ctypes.data for k, v in host.items()}
        chunk_descs = [ctx.descs(e2e_scenes[a:a + chunk]) for a in range(0, n_e2e, chunk)]
        d2h = n_e2e * W * H * BYTES_PER_PX
        h2d0 = ctx.stats().bytes_h2d

        def e2e_step():
            for cd in chunk_descs:
                ctx.render_host(cd, host_ptrs, TARGET_MASK)

        e2e_step()
        h2d = ctx.stats().bytes_h2d - h2d0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        t = torch.tensor([dt_e2e, -d2h_gbs], dtype=torch.float64, device=dev)
        tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        e2e_fps = n_scenes * args.steps / float(t[0].item())
        link = -float(t[1].item())                                           # the slowest rank's link
        per_gpu_gbs = e2e_fps / world * W * H * BYTES_PER_PX / 1e9
        e2e = {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
               "d2h_link_gbs": link, "d2h_achieved_gbs_per_gpu": per_gpu_gbs, "frac_of_d2h_link": per_gpu_gbs / link if link else None,
               "shard_sizes": shard_sizes, "sharding": "equal" if world == 1 else "proportional to each rank's sustained D2H bandwidth measured while all ranks copy",
               "note": f"slb_render_batch_host, page-locked host buffers (slb_host_alloc), {chunk}-scene calls; every frame returns "
                       f"{W * H * BYTES_PER_PX / 1e6:.2f} MB over the device->host link, whose pinned-copy bandwidth (d2h_link_gbs, "
                       "measured in this process; the minimum over ranks probing at the same time) bounds this number"}

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
    ncu = {}
    mpath = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")          # written by tools/ncu_metrics.py from the committed capture
    if os.path.exists(mpath):
        ncu = json.load(open(mpath)).get(args.config, {})
    n_sub = -(-n_local // subbatch)
    frames_per_launch = n_local / n_sub
    shade_ms = stage_ms[5] / (stage_steps * n_sub)
    alg_bytes = frames_per_launch * W * H * BYTES_PER_PX
    achieved = alg_bytes / (shade_ms * 1e-3) / 1e9 if shade_ms > 0 else 0.0
    setup_ms = stage_ms[1] / (stage_steps * n_sub)
    geom_bytes = sum(sum(o.mesh.geometry_bytes() for o in sc.objects) for sc in scenes[:64]) / min(64, len(scenes)) * (1 + c["n_lights"])
    setup_gbs = geom_bytes * frames_per_launch / (setup_ms * 1e-3) / 1e9 if setup_ms > 0 else 0.0
    names = ["shadow", "bin_count", "scan", "bin_emit", "raster", "shade_store", "ssao", "post"]
    ks, kb = ncu.get("k_shade", {}), ncu.get("k_setup", {})
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.config, n_scenes), "scenes_per_gpu": n_local, "subbatch": subbatch,
                       "rank0_numa_node": numa_node,
                       "l2": "outputs per step (%.1f GB) exceed L2; no flush needed" % (n_local * W * H * BYTES_PER_PX / 1e9)},
            "clocks": sampler.summary(), "gpu_launches": int(launches),
            "pipeline": {"two_streams": two_streams, "ms_per_step_single_stream": serial_ms_per_step, "stage_pass_steps": stage_steps,
                         "note": "value / ms_per_step: the timed region, sub-batch set-up and shade passes on two streams (SLB_OPT_OVERLAP) unless --overlap off; "
                                 "stage_ms_per_step, roofline and binner: CUDA events around every stage in a single-stream pass of the same steps "
                                 "right after it (on two streams a stage's event interval also contains the other stream's kernels)"},
            "stage_ms_per_step": {n: float(v / stage_steps) for n, v in zip(names, stage_ms)},
            "roofline": {"bound": "hbm", "kernel": "k_shade (shade + MRT store)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "measured": "CUDA events on the launching stream, single-stream stage pass (see pipeline.note)",
                         "frac": achieved / peak,
                         "traffic": ks["dram_bytes_per_frame"] * frames_per_launch if "dram_bytes_per_frame" in ks else None,
                         "traffic_source": "profiles/r02_ncu_metrics.json (ncu --set full: dram read + write bytes of one k_shade launch, per frame, scaled to this launch size)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": shade_ms,
                         # what actually bounds the kernel: issue slots and memory latency, from the committed ncu capture
                         "issue_slot_util_pct_ncu": ks.get("issue_pct"), "warp_instructions_per_32_pixels_ncu":
                             (ks["warp_inst_per_frame"] / (W * H / 32.0)) if "warp_inst_per_frame" in ks else None},
            "binner": {"kernel": "k_setup (triangle set-up, direct raster, tile count)", "achieved": setup_gbs, "unit": "GB/s", "frac": setup_gbs / peak,
                       "algorithmic_bytes_per_launch": geom_bytes * frames_per_launch, "ms_per_launch": setup_ms,
                       "algorithmic_bytes": "sum over drawn sub-meshes and views (camera + shadow) of n_verts*68 + n_idx*4 (SURVEY 8d B_geom)",
                       "issue_slot_util_pct_ncu": kb.get("issue_pct"), "active_threads_per_instruction_ncu": kb.get("threads_per_inst"),
                       "l2_hit_rate_pct": kb.get("l2_hit_pct"), "l2_hit_source": "profiles/r02_ncu_metrics.json (ncu lts__t_sector_hit_rate.pct)"},
            "triangles_per_frame": int(sub.triangles_submitted / max(1, sub.frames_rendered))}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        sample = min(n_scenes, max(8, (cores if W <= 640 else cores // 2)))
        base = CpuBaseline(args.config, pool, sample, build_light_map(args.config))
        best = max(base.run() for _ in range(2))
        line["cpu_baseline"] = {"value": best[0], "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"scenes 0..{sample - 1} of the workload, OpenMP oracle, one scene per host thread ({best[1]:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
