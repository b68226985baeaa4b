"""Host-side mirror of the reference's `stillleben.camera_model` (python/stillleben/camera_model.py).

    from stillleben_b200 import camera_model      # instead of:  from stillleben import camera_model

Same entry points, argument meanings and 3xHxW float return values: chromatic_aberration, blur, exposure, noise,
color_jitter, process_deterministic, process_image. The reference runs ~40 PyTorch ops per image; here every
entry point is one call of the fused CUDA kernels behind `slb_camera_model` (k_camera.cu), and `process_batch`
pushes a whole batch (optionally straight from the RGBA8 render target) through them in one launch pair.
The noise stage draws from a counter-based Philox generator seeded from torch's global RNG, so results are
reproducible under torch.manual_seed but are a different sample than torch.poisson / normal_ would give.
"""
import math
import random

import torch

from . import abi
from . import sl as _sl

__all__ = ["chromatic_aberration", "blur", "exposure", "noise", "color_jitter", "process_deterministic", "process_image",
           "process_batch", "random_parameters"]


def _stream(ctx):
    """torch's CURRENT stream on the context's device: the library work is ordered after the torch ops that produced the
    inputs (stack / to / contiguous) instead of racing them on a private stream."""
    import torch
    return torch.cuda.current_stream(torch.device("cuda", ctx.device)).cuda_stream


def _params(stages, chromatic_translation=None, chromatic_scaling=None, blur_sigma=0.0, exposure_deltaS=0.0, do_noise=False,
            noise_a=0.0, noise_b=0.0, hue_shift=0.0, seed=None):
    p = abi.CameraParams()
    tr = torch.zeros(3, 2) if chromatic_translation is None else torch.as_tensor(chromatic_translation, dtype=torch.float32).cpu()
    sc = torch.ones(3) if chromatic_scaling is None else torch.as_tensor(chromatic_scaling, dtype=torch.float32).cpu()
    for c in range(3):
        p.chromatic_translation[c][0], p.chromatic_translation[c][1] = float(tr[c, 0]), float(tr[c, 1])
        p.chromatic_scaling[c] = float(sc[c])
    p.blur_sigma, p.exposure_deltaS = float(blur_sigma), float(exposure_deltaS)
    p.do_noise, p.noise_a, p.noise_b = int(bool(do_noise)), float(noise_a), float(noise_b)
    p.hue_shift, p.stages = float(hue_shift), int(stages)
    p.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
    return p


def _run(images, params, u8=False):
    """images: float [n,3,H,W] (or uint8 [n,H,W,4] with u8=True) on any device; returns float [n,3,H,W] there."""
    ctx = _sl._context()
    dev = torch.device("cuda", _sl._cuda_index)
    src = images.device
    x = images.to(dev).contiguous() if u8 else images.to(dev, torch.float32).contiguous()
    n = x.shape[0]
    H, W = (x.shape[1], x.shape[2]) if u8 else (x.shape[2], x.shape[3])
    out = torch.empty((n, 3, H, W), dtype=torch.float32, device=dev)
    arr = (abi.CameraParams * n)(*params)
    rc = ctx.lib.slb_camera_model(ctx.h, x.data_ptr(), 1 if u8 else 0, out.data_ptr(), n, H, W, arr, _stream(ctx))
    if rc != 0:
        raise RuntimeError(ctx.lib.slb_last_error(ctx.h).decode())
    ctx.synchronize()
    return out.to(src)


def _check(rgb):
    assert rgb.dim() == 3, "input tensor has invalid size {}".format(rgb.size())
    assert rgb.size(0) == 3, "input tensor has invalid size {}".format(rgb.size())


def chromatic_aberration(rgb, translations, scaling):
    _check(rgb)
    return _run(rgb.unsqueeze(0), [_params(abi.CAM_CHROMATIC, translations, scaling, seed=0)])[0]


def blur(rgb, sigma):
    return _run(rgb.unsqueeze(0), [_params(abi.CAM_BLUR, blur_sigma=sigma, seed=0)])[0]


def exposure(rgb, deltaS):
    return _run(rgb.unsqueeze(0), [_params(abi.CAM_EXPOSURE, exposure_deltaS=deltaS, seed=0)])[0]


def noise(rgb, a, b):
    return _run(rgb.unsqueeze(0), [_params(abi.CAM_NOISE, do_noise=True, noise_a=a, noise_b=b)])[0]


def color_jitter(tensor_img, hue_shift):
    assert tensor_img.size(0) == 3
    return _run(tensor_img.unsqueeze(0), [_params(abi.CAM_HUE, hue_shift=hue_shift, seed=0)])[0]


def process_deterministic(rgb, chromatic_translation, chromatic_scaling, blur_sigma, exposure_deltaS, do_noise, noise_a, noise_b,
                          hue_shift, seed=None):
    _check(rgb)
    p = _params(abi.CAM_ALL, chromatic_translation, chromatic_scaling, blur_sigma, exposure_deltaS, do_noise, noise_a, noise_b, hue_shift, seed)
    return _run(rgb.unsqueeze(0), [p])[0]


def random_parameters():
    """The parameter distribution of process_image (camera_model.py:264-286)."""
    hue_jitter = 0.05
    return dict(chromatic_translation=torch.empty(3, 2).uniform_(-0.002, 0.002), chromatic_scaling=torch.empty(3).uniform_(0.998, 1.002),
                blur_sigma=random.uniform(0.0, 3.0) if random.random() > 0.3 else 0.0, exposure_deltaS=random.uniform(-2, 1.2),
                do_noise=random.random() > 0.3, noise_a=random.random() * 0.04, noise_b=random.random() * 0.02,
                hue_shift=random.uniform(-hue_jitter, hue_jitter))


def process_image(rgb):
    _check(rgb)
    return process_deterministic(rgb, **random_parameters())


def process_batch(images, parameters=None):
    """Extension: n images in one launch pair. `images` is float [n,3,H,W] in [0,1] or the renderer's uint8
    [n,H,W,4] colour target; `parameters` a list of n dicts as random_parameters() returns (default: random)."""
    u8 = images.dtype == torch.uint8
    n = images.shape[0]
    parameters = parameters or [random_parameters() for _ in range(n)]
    return _run(images, [_params(abi.CAM_ALL, **p) for p in parameters], u8=u8)


del math
