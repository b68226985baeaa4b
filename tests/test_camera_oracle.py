"""The numpy oracle of the camera noise model against vectors produced by the REFERENCE's own module
(tests/golden/golden_camera_model.npz <- tests/golden/make_camera_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import camera_model_np as cm  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_camera_model.npz"))


def params(i):
    p = G[f"par{i}"]
    return p[:6].reshape(3, 2), p[6:9], float(p[9]), float(p[10]), float(p[11])


def test_stages_match_the_reference_module():
    for i in range(3):
        img = G[f"in{i}"]
        tr, sc, bl, dS, hue = params(i)
        np.testing.assert_allclose(cm.chromatic_aberration(img, tr, sc), G[f"ca{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.blur(img, max(bl, 0.4)), G[f"blur{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.exposure(img, dS), G[f"exp{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.color_jitter(img, hue), G[f"hue{i}"], atol=2e-5)


def test_process_deterministic_matches_the_reference_module():
    for i in range(3):
        tr, sc, bl, dS, hue = params(i)
        got = cm.process_deterministic(G[f"in{i}"], tr, sc, bl, dS, hue)
        # the hue conversion is discontinuous where two channels tie; a few pixels may take the other branch
        bad = np.abs(got - G[f"full{i}"]) > 1e-4
        assert bad.mean() < 2e-3, bad.mean()


def test_noise_moments_describe_the_reference_sampler():
    for j in range(3):
        a, b = map(float, G[f"noise_par{j}"])
        mean, var = G[f"noise_mean_var{j}"]
        m, v = cm.noise_moments(0.37, a, b)
        assert abs(mean - m) < 2e-3 and abs(var - v) < 0.05 * max(v, 1e-6)
