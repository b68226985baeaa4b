"""The fused camera-model kernels (slb_camera_model) against vectors produced by the REFERENCE's own module
(tests/golden/golden_camera_model.npz), against the numpy oracle on other sizes, and the noise stage against
the moments of the reference's sampler."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import camera_model_np as cmo  # noqa: E402
from stillleben_b200 import camera_model as cm  # noqa: E402
from stillleben_b200 import sl  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "golden_camera_model.npz"))


def params(i):
    p = G[f"par{i}"]
    return torch.from_numpy(p[:6].reshape(3, 2).copy()), torch.from_numpy(p[6:9].copy()), float(p[9]), float(p[10]), float(p[11])


def test_stages_match_the_reference_vectors():
    sl.init_cuda(0)
    for i in range(3):
        img = torch.from_numpy(G[f"in{i}"])
        tr, sc, bl, dS, hue = params(i)
        np.testing.assert_allclose(cm.chromatic_aberration(img, tr, sc).numpy(), G[f"ca{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.blur(img, max(bl, 0.4)).numpy(), G[f"blur{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.exposure(img, dS).numpy(), G[f"exp{i}"], atol=1e-5)
        np.testing.assert_allclose(cm.color_jitter(img, hue).numpy(), G[f"hue{i}"], atol=3e-5)


def test_process_deterministic_matches_the_reference_vectors():
    sl.init_cuda(0)
    for i in range(3):
        tr, sc, bl, dS, hue = params(i)
        got = cm.process_deterministic(torch.from_numpy(G[f"in{i}"]).cuda(), tr, sc, bl, dS, False, 0.0, 0.0, hue)
        assert got.is_cuda and got.shape == G[f"full{i}"].shape
        bad = np.abs(got.cpu().numpy() - G[f"full{i}"]) > 1e-4            # hue branch flips where two channels tie
        assert bad.mean() < 2e-3, bad.mean()


def test_batch_from_render_target_matches_oracle():
    sl.init_cuda(0)
    rng = np.random.RandomState(1)
    n, H, W = 3, 75, 133                                                    # ragged w.r.t. the 32x8 blocks
    u8 = rng.randint(0, 256, size=(n, H, W, 4)).astype(np.uint8)
    ps = [dict(chromatic_translation=torch.tensor(rng.uniform(-0.002, 0.002, (3, 2)), dtype=torch.float32),
               chromatic_scaling=torch.tensor(rng.uniform(0.998, 1.002, 3), dtype=torch.float32), blur_sigma=[1.7, 0.0, 2.9][k],
               exposure_deltaS=[-1.0, 0.4, 1.1][k], do_noise=False, noise_a=0.0, noise_b=0.0, hue_shift=[0.04, -0.02, 0.0][k]) for k in range(n)]
    got = cm.process_batch(torch.from_numpy(u8).cuda(), ps).cpu().numpy()
    for k in range(n):
        img = u8[k, :, :, :3].transpose(2, 0, 1).astype(np.float32) / 255.0
        ref = cmo.process_deterministic(img, ps[k]["chromatic_translation"].numpy(), ps[k]["chromatic_scaling"].numpy(), ps[k]["blur_sigma"],
                                        ps[k]["exposure_deltaS"], ps[k]["hue_shift"])
        bad = np.abs(got[k] - ref) > 1e-4
        assert bad.mean() < 2e-3, (k, bad.mean())


def test_noise_distribution_matches_the_reference_sampler():
    sl.init_cuda(0)
    torch.manual_seed(0)
    flat = torch.full((3, 256, 256), 0.37).cuda()
    for j in range(3):
        a, b = map(float, G[f"noise_par{j}"])
        n = cm.noise(flat, a, b)
        mean, var = float(n.mean()), float(n.var())
        ref_mean, ref_var = G[f"noise_mean_var{j}"]
        assert abs(mean - ref_mean) < 2e-3, (j, mean, ref_mean)
        assert abs(var - ref_var) < 0.05 * max(ref_var, 1e-6), (j, var, ref_var)
        if a > 0:   # Poissonian part: values sit on the lattice k * a
            k = (n / a).cpu()
            if b == 0:
                assert float((k - k.round()).abs().max()) < 1e-2
    # small-lambda branch of the sampler (multiplication method): dark pixel, strong signal-dependent noise
    dark = torch.full((3, 128, 128), 0.05).cuda()
    n = cm.noise(dark, 0.04, 0.0)
    assert abs(float(n.mean()) - 0.05) < 2e-3 and abs(float(n.var()) - 0.04 * 0.05) < 0.1 * 0.04 * 0.05
    # reproducible under the torch seed, different across calls
    torch.manual_seed(7); n1 = cm.noise(flat, 0.03, 0.01)
    torch.manual_seed(7); n2 = cm.noise(flat, 0.03, 0.01)
    n3 = cm.noise(flat, 0.03, 0.01)
    assert torch.equal(n1, n2) and not torch.equal(n1, n3)
