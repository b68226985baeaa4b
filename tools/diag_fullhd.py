import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures, oracle_util as ou, parity
from stillleben_b200 import abi, lib, synth
ctx = lib.Context(0); ctx.lightmap_sizes = (64, 16, 32, 64, 64)
scene = synth.tabletop_scene(fixtures.small_pool(), 31, n_objects=8, width=1920, height=1080, intrinsics=None, n_lights=3, ssao=True)
ref = ou.render(scene)
for prep in (1, 0):
    for mask in (1, 0):
        ctx.set_option(abi.OPT_HUGE_PREPARE, prep); ctx.set_option(abi.OPT_SHADOW_MASK, mask); ctx.set_option(abi.OPT_KEEP_HDR, 1)
        res = ctx.render([scene], target_mask=abi.TARGETS_ALL); ctx.synchronize()
        gpu = res.frame_dict(0); gpu["hdr"] = res.hdr(0)
        st = parity.compare(gpu, ref)
        d = np.abs(gpu["rgb"].astype(int) - ref["rgb"].astype(int)).max(-1)
        ys, xs = np.nonzero(d > 1)
        print("prep", prep, "mask", mask, st["rgb"], "hdr bad", st["hdr"]["bad"], "where", list(zip(ys[:6].tolist(), xs[:6].tolist())),
              "inst at", [int(gpu["instance_index"][y, x, 0]) for y, x in zip(ys[:6], xs[:6])])
