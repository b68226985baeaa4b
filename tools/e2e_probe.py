import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stillleben_b200 import abi, lib, synth
ctx = lib.Context(0)
pool = synth.mesh_pool(21)
scenes = [synth.tabletop_scene(pool, 1000 + s) for s in range(128)]
descs = ctx.descs(scenes)
host = {t: ctx.host_alloc((128, 480, 640, ch), dt) for t, (dt, ch) in enumerate(abi.TARGET_FORMATS) if abi.TARGETS_SIX & (1 << t)}
ptrs = {k: v.ctypes.data for k, v in host.items()}
ctx.render_host(descs, ptrs)
ctx.render_host(descs, ptrs)
os.environ["SLB_DEBUG_TIMING"] = "1"
t = time.perf_counter(); ctx.render_host(descs, ptrs); print("render_host 128 scenes", (time.perf_counter() - t) * 1e3, "ms")
