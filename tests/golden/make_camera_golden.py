"""Generates tests/golden/golden_camera_model.npz by importing the REFERENCE's own camera model
(/root/reference/python/stillleben/camera_model.py, pure PyTorch) in this container — it cannot travel to the
GPU box, the vectors do. The module is loaded under a stub package so that `from . import profiling` resolves
without the reference's compiled extension.

    python tests/golden/make_camera_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/python/stillleben"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_camera_model.npz")


def load_reference():
    pkg = types.ModuleType("slref")
    pkg.__path__ = [REF]
    sys.modules["slref"] = pkg
    for name in ("profiling", "camera_model"):
        spec = importlib.util.spec_from_file_location(f"slref.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"slref.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["slref.camera_model"]


def test_image(seed, H, W):
    """Smooth colourful image with edges and a few saturated / black regions (hue and exposure corner cases)."""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.stack([0.5 + 0.5 * np.sin(x / 7.0 + k) * np.cos(y / 5.0 - k) for k in range(3)], 0)
    img += 0.15 * rng.rand(3, H, W).astype(np.float32)
    img = np.clip(img, 0, 1)
    img[:, :6, :9] = 0.0
    img[:, -5:, -7:] = 1.0
    img[0, 10:20, 10:20] = 1.0; img[1, 10:20, 10:20] = 0.0; img[2, 10:20, 10:20] = 0.0
    img[:, 20:24, 30:40] = 0.5                      # grey: C == 0 branch of the hue conversion
    return img.astype(np.float32)


def main():
    cm = load_reference()
    torch.manual_seed(0)
    out = {}
    cases = [
        dict(H=48, W=64, tr=[[0.002, -0.001], [0.0, 0.0], [-0.0015, 0.002]], sc=[1.002, 1.0, 0.998], blur=1.3, dS=0.7, hue=0.03),
        dict(H=37, W=53, tr=[[0.0, 0.0], [0.0, 0.0], [0.0, 0.0]], sc=[1.0, 1.0, 1.0], blur=0.0, dS=-1.5, hue=-0.05),
        dict(H=40, W=40, tr=[[-0.002, 0.002], [0.001, 0.001], [0.002, -0.002]], sc=[0.998, 1.001, 1.002], blur=2.8, dS=0.0, hue=0.0),
    ]
    for i, c in enumerate(cases):
        img = test_image(i, c["H"], c["W"])
        t = torch.from_numpy(img)
        tr, sc = torch.tensor(c["tr"]), torch.tensor(c["sc"])
        out[f"in{i}"] = img
        out[f"par{i}"] = np.array(sum(c["tr"], []) + c["sc"] + [c["blur"], c["dS"], c["hue"]], np.float32)
        out[f"ca{i}"] = cm.chromatic_aberration(t, tr, sc).numpy()
        out[f"blur{i}"] = cm.blur(t, max(c["blur"], 0.4)).numpy()
        out[f"exp{i}"] = cm.exposure(t, c["dS"]).numpy()
        out[f"hue{i}"] = cm.color_jitter(t.clone(), c["hue"]).numpy()
        out[f"full{i}"] = cm.process_deterministic(t.clone(), tr, sc, c["blur"], c["dS"], False, 0.0, 0.0, c["hue"]).numpy()
    # noise statistics of the reference sampler (torch.poisson + normal) for the distribution test
    flat = torch.full((3, 256, 256), 0.37)
    for j, (a, b) in enumerate([(0.03, 0.015), (0.002, 0.0), (0.0, 0.01)]):
        n = cm.noise(flat.clone(), a, b)
        out[f"noise_par{j}"] = np.array([a, b], np.float32)
        out[f"noise_mean_var{j}"] = np.array([float(n.mean()), float(n.var())], np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.startswith("full")})


if __name__ == "__main__":
    main()
