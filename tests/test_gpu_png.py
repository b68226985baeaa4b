"""The batched PNG encoder on the GPU (slb_png_encode): byte-identical to its CPU restatement (oracle/png_np.py) on
small images of every accepted format, and at full size through the property that matters — an independent decoder
(PIL) returns exactly the rendered pixels; plus the ImageSaver mirror used the way the reference documents it."""
import io
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import png_np  # noqa: E402
import test_png_oracle as tpo  # noqa: E402
from stillleben_b200 import image_saver, sl  # noqa: E402

pytestmark = pytest.mark.gpu
Image = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("name", sorted(tpo.images()))
def test_kernel_bytes_equal_the_restatement(name):
    sl.init_cuda(0)
    img = tpo.images()[name]
    t = torch.from_numpy(img.view(np.int16) if img.dtype == np.uint16 else img)
    batch = torch.stack([t, t.flip(0), t])                       # three files per call, two identical
    files = image_saver.encode_batch(batch.cuda())
    ref = png_np.encode(img)
    assert files[0] == ref and files[2] == ref
    assert files[1] == png_np.encode(img[::-1])


def test_full_size_batch_round_trips_through_pil():
    sl.init_cuda(0)
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:480, 0:640]
    base = np.stack([(xx * 255 // 639), (yy * 255 // 479), ((xx + yy) % 256), np.full_like(xx, 255)], -1).astype(np.uint8)
    imgs = np.stack([np.roll(base, 17 * k, axis=1) for k in range(8)])
    imgs[3, 100:300, 200:500, :3] = rng.randint(0, 256, (200, 300, 3))       # an incompressible patch
    imgs[5, :, :, :3] = 0                                                     # a flat frame
    files = image_saver.encode_batch(torch.from_numpy(imgs).cuda())
    for k, data in enumerate(files):
        got = np.asarray(Image.open(io.BytesIO(data)))
        assert np.array_equal(got, imgs[k]), k
    assert len(files[5]) < 40_000 and len(files[3]) > 150_000
    ids = (rng.randint(0, 21, (4, 30, 40)).repeat(16, 1).repeat(16, 2)).astype(np.int16)
    for k, data in enumerate(image_saver.encode_batch(torch.from_numpy(ids).cuda())):
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))).astype(np.int16), ids[k])


def test_image_saver_like_the_reference_docstring(tmp_path):
    sl.init_cuda(0)
    saver = image_saver.ImageSaver()
    with pytest.raises(RuntimeError, match="__enter__"):                     # py_image_saver.cpp:39-40
        saver.save(torch.zeros(4, 4, 3, dtype=torch.uint8), str(tmp_path / "x.png"))
    with image_saver.ImageSaver() as saver:
        with pytest.raises(ValueError):
            saver.save(torch.zeros(640, 480, 3), str(tmp_path / "float.png"))   # "Color images need to have type uint8"
        with pytest.raises(ValueError):
            saver.save(torch.zeros(4, 4, 2, dtype=torch.uint8), str(tmp_path / "c2.png"))
        rgb = (torch.arange(48 * 64 * 3) % 251).to(torch.uint8).view(48, 64, 3).cuda()
        depth = (torch.arange(48 * 64) % 3000).to(torch.int16).view(48, 64)
        for i in range(70):                                                    # more than one internal flush
            saver.save(rgb, str(tmp_path / f"rgb{i}.png"))
        saver.save(depth, str(tmp_path / "d.png"))
    assert np.array_equal(np.asarray(Image.open(tmp_path / "rgb69.png")), rgb.cpu().numpy())
    assert np.array_equal(np.asarray(Image.open(tmp_path / "d.png")).astype(np.int16), depth.numpy())
