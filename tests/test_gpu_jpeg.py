"""The batched JPEG encoder on the GPU (slb_jpeg_encode): the files must be byte-identical to the CPU restatement
(oracle/jpeg_np.py), which tests/test_jpeg_oracle.py pins on libjpeg's own output — and, directly, to what libjpeg (through PIL)
writes for the same pixels at full size; plus the ImageSaver mirror choosing JPEG from the file extension like the reference's
AnyImageConverter (src/image_saver.cpp:55-97)."""
import io
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import jpeg_np  # noqa: E402
import test_jpeg_oracle as tjo  # noqa: E402
from stillleben_b200 import image_saver, sl  # noqa: E402

pytestmark = pytest.mark.gpu
Image = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("shape", tjo.SHAPES)
def test_kernel_bytes_equal_the_restatement(shape):
    sl.init_cuda(0)
    imgs = [tjo.make_image(*shape, kind, seed=sum(shape)) for kind in ("smooth", "noise", "flat")]
    files = image_saver.encode_batch_jpeg(torch.from_numpy(np.stack(imgs)).cuda())
    for img, data in zip(imgs, files):
        assert data == jpeg_np.encode(img)


@pytest.mark.parametrize("quality", [1, 50, 95, 100])
def test_qualities(quality):
    sl.init_cuda(0)
    img = tjo.make_image(40, 56, 3, "smooth", 3)
    assert image_saver.encode_batch_jpeg(torch.from_numpy(img[None]).cuda(), quality)[0] == tjo.libjpeg_bytes(img, quality)


def test_full_size_batch_equals_libjpeg():
    """640x480 RGB / RGBA / grey batches: every file is exactly libjpeg's; a noise frame exercises the stride retry."""
    sl.init_cuda(0)
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:480, 0:640]
    base = np.stack([(xx * 255 // 639), (yy * 255 // 479), ((xx + yy) % 256)], -1).astype(np.uint8)
    imgs = np.stack([np.roll(base, 17 * k, axis=1) for k in range(6)])
    imgs[2] = rng.randint(0, 256, (480, 640, 3))                              # incompressible: larger than the first stride
    imgs[4] = 255
    files = image_saver.encode_batch_jpeg(torch.from_numpy(imgs).cuda(), first_stride=100_000)
    for k, data in enumerate(files):
        assert data == tjo.libjpeg_bytes(imgs[k]), k
    assert len(files[2]) > 150_000 and len(files[4]) < 12_000          # the noise frame did not fit the first stride
    rgba = np.concatenate([imgs, np.full((6, 480, 640, 1), 200, np.uint8)], -1)
    assert image_saver.encode_batch_jpeg(torch.from_numpy(rgba[:2]).cuda()) == files[:2]
    grey = imgs[..., 1].copy()
    for k, data in enumerate(image_saver.encode_batch_jpeg(torch.from_numpy(grey).cuda())):
        assert data == tjo.libjpeg_bytes(grey[k]), k


def test_full_hd_odd_size():
    sl.init_cuda(0)
    rng = np.random.RandomState(1)
    img = np.clip(rng.normal(128, 40, (1, 1080, 1917, 3)), 0, 255).astype(np.uint8)     # 1080 = 67.5 MCU rows, 1917: ragged columns
    assert image_saver.encode_batch_jpeg(torch.from_numpy(img).cuda())[0] == tjo.libjpeg_bytes(img[0])


def test_image_saver_picks_jpeg_from_the_extension(tmp_path):
    sl.init_cuda(0)
    img = tjo.make_image(48, 64, 3, "smooth", 9)
    with image_saver.ImageSaver() as saver:
        saver.save(torch.from_numpy(img), str(tmp_path / "a.jpg"))
        saver.save(torch.from_numpy(img), str(tmp_path / "b.JPEG"))
        saver.save(torch.from_numpy(img), str(tmp_path / "c.png"))
        saver.save(torch.from_numpy(img[..., 0].copy()), str(tmp_path / "d.jpeg"))
        with pytest.raises(ValueError):
            saver.save(torch.zeros(8, 8, dtype=torch.int16), str(tmp_path / "e.jpg"))
    assert open(tmp_path / "a.jpg", "rb").read() == tjo.libjpeg_bytes(img) == open(tmp_path / "b.JPEG", "rb").read()
    assert open(tmp_path / "d.jpeg", "rb").read() == tjo.libjpeg_bytes(img[..., 0].copy())
    assert np.array_equal(np.asarray(Image.open(tmp_path / "c.png")), img)


def test_argument_checks():
    sl.init_cuda(0)
    ctx = sl._context()
    assert ctx.lib.slb_jpeg_bound(480, 640, 2) == 0 and ctx.lib.slb_jpeg_bound(480, 640, 3) > 480 * 640
    x = torch.zeros(8, 8, 3, dtype=torch.uint8, device="cuda")
    out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), 1, 8, 8, 3, 0, out.data_ptr(), 4096, sizes.data_ptr(), None) != 0      # quality 0
    assert ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), 1, 8, 8, 2, 80, out.data_ptr(), 4096, sizes.data_ptr(), None) != 0     # 2 channels
    assert ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), 1, 20000, 20000, 3, 80, out.data_ptr(), 4096, sizes.data_ptr(), None) != 0   # > 2.5 M blocks
    assert ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), 1, 8, 8, 3, 80, out.data_ptr(), 100, sizes.data_ptr(), None) == 0      # too small a stride
    ctx.synchronize()
    assert int(sizes[0]) == 0 and int(out[100:].sum()) == 0


def test_two_contexts_with_different_qualities_do_not_share_tables():
    """The quantisation divisors travel as a kernel parameter: a second context encoding at another quality in between
    must not change what the first one writes (its header / table cache stays valid)."""
    from stillleben_b200 import lib
    sl.init_cuda(0)
    ctx_a = sl._context()
    ctx_b = lib.Context(0)
    img = tjo.make_image(48, 64, 3, "smooth", 11)
    x = torch.from_numpy(img[None]).cuda()

    def encode(ctx, quality):
        out = torch.zeros((1, 1 << 16), dtype=torch.uint8, device="cuda")
        sizes = torch.zeros(1, dtype=torch.int32, device="cuda")
        assert ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), 1, 48, 64, 3, quality, out.data_ptr(), out.shape[1], sizes.data_ptr(), None) == 0
        ctx.synchronize()
        return out[0, :int(sizes[0])].cpu().numpy().tobytes()

    try:
        assert encode(ctx_a, 80) == tjo.libjpeg_bytes(img, 80)
        assert encode(ctx_b, 35) == tjo.libjpeg_bytes(img, 35)
        assert encode(ctx_a, 80) == tjo.libjpeg_bytes(img, 80)
        assert encode(ctx_b, 35) == tjo.libjpeg_bytes(img, 35)
    finally:
        ctx_b.close()
