"""Prints the parity statistics (CUDA vs oracle) of every fixture variant: the measured bound behind the outlier budget."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures, oracle_util as ou, parity
from stillleben_b200 import abi, lib
ctx = lib.Context(0); ctx.lightmap_sizes = (64, 16, 32, 64, 64)
rows = {}
for name in fixtures.VARIANTS:
    scene = fixtures.variant(name)
    assets = ou.OracleAssets()
    if scene.light_map is not None:
        assets.set_lightmap_maps(scene.light_map, *ctx.read_lightmap(scene.light_map))
    ctx.set_option(abi.OPT_KEEP_HDR, 1)
    res = ctx.render([scene], target_mask=abi.TARGETS_ALL); ctx.synchronize()
    gpu = res.frame_dict(0); gpu["hdr"] = res.hdr(0)
    ref = ou.render(scene, assets)
    st = parity.compare(gpu, ref)
    rows[name] = {"rgb_over1": st["rgb"]["over1"], "rgb_max": st["rgb"]["max"], "hdr_bad": st["hdr"]["bad"], "n": st["rgb"]["n"],
                  "float_bad": {k: v["bad"] for k, v in st.items() if "bad" in v and k != "hdr"},
                  "id_mismatch": sum(v["mismatch"] for v in st.values() if "mismatch" in v)}
    print(name, rows[name])
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_parity_stats.json"), "w"), indent=1)
