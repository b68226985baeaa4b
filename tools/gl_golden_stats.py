"""Prints, for every committed OpenGL golden (tests/golden/gl_ref_*.npz), how far the CUDA path is from it: pixels with a different id /
coverage, largest coordinate difference, RGBA8 pixels beyond one level.   python tools/gl_golden_stats.py   (needs a GPU)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures, test_gl_golden as T
from stillleben_b200 import abi, lib

ctx = lib.Context(0)
ctx.lightmap_sizes = (64, 16, 32, 64, 64)
out = {}
def stats(frame, z):
    depth = z["coord"][..., 3] if "coord" in z.files else z["depth"]
    bad = (((depth == abi.INVALID_COORD) != (frame["coord"][..., 3] == abi.INVALID_COORD)) | np.any(z["instance_index"] != frame["instance_index"], axis=-1)
           | np.any(z["class_index"] != frame["class_index"], axis=-1) | np.any(z["vertex_index"] != frame["vertex_index"][..., :3], axis=-1))
    ok = ~bad
    d8 = np.abs(z["rgb"].astype(int) - frame["rgb"].astype(int)).max(-1)[ok]
    row = {"pixels": int(bad.size), "id_or_coverage_differs": int(bad.sum()), "depth_max_abs": float(np.abs(depth - frame["coord"][..., 3])[ok].max()),
           "rgba8_beyond_1_level": int((d8 > 1).sum()), "rgba8_beyond_8_levels": int((d8 > 8).sum())}
    if "normals" in z.files:
        row["normals_beyond_1e-2"] = int((np.abs(z["normals"].astype(np.float32) - frame["normals"]).max(-1)[ok] > 1e-2).sum())
    return row
def render(sc):
    res = ctx.render([sc], target_mask=abi.TARGETS_ALL); ctx.synchronize(); return res.frame_dict(0)
for name in T.GL_GOLDEN:
    out[name] = stats(render(fixtures.single_level_copy(fixtures.gl_scene_of(name))), np.load(os.path.join(fixtures.GOLDEN, f"gl_ref_{name}.npz")))
for name in ("ssao", "ibl"):
    out["post_" + name] = stats(render(T.post_scene(name)), np.load(os.path.join(fixtures.GOLDEN, f"gl_ref_post_{name}.npz")))
out["bench_c3"] = stats(render(T.bench_scene()), np.load(os.path.join(fixtures.GOLDEN, "gl_ref_bench_c3.npz")))
for k, v in out.items():
    print(k, v)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gl_golden_stats.json"), "w"), indent=1)
