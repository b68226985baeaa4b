"""Development probe: render a few scenes with the CUDA library and the oracle, print the parity stats."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_util as ou
import parity
from stillleben_b200 import abi, lib, synth

ctx = lib.Context(0)
ctx.set_option(abi.OPT_KEEP_HDR, 1)
A = ou.OracleAssets()
pool = synth.mesh_pool(6, nu=64, nv=32, tex_size=128)
scenes = [("c1", synth.config1_scene())]
for s in range(3):
    scenes.append((f"tab{s}", synth.tabletop_scene(pool, 1000 + s, n_objects=8)))
sc = synth.tabletop_scene(pool, 2000, n_objects=6, ssao=True)
scenes.append(("ssao", sc))
sc = synth.tabletop_scene(pool, 2001, n_objects=6, manual_exposure=-1.0)
scenes.append(("autoexp", sc))
for name, sc in scenes:
    t = time.time()
    res = ctx.render([sc], target_mask=abi.TARGETS_ALL)
    ctx.synchronize()
    g = res.frame_dict(0)
    g["hdr"] = res.hdr(0)
    tg = time.time() - t
    t = time.time()
    r = ou.render(sc, A)
    to = time.time() - t
    st = parity.compare(g, r)
    print(name, f"gpu {tg*1e3:.1f} ms oracle {to*1e3:.1f} ms")
    for k, v in st.items():
        print("   ", k, v)
    if "--save" in sys.argv:
        from PIL import Image
        os.makedirs("gpurun_out", exist_ok=True)
        Image.fromarray(g["rgb"][..., :3]).save(f"gpurun_out/{name}_gpu.png")
        Image.fromarray(r["rgb"][..., :3]).save(f"gpurun_out/{name}_orc.png")
print("stats", ctx.stats().kernel_launches, ctx.stats().triangles_binned)
