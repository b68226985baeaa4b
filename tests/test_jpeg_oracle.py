"""oracle/jpeg_np.py pinned on libjpeg itself: the restatement must produce EXACTLY the file libjpeg writes (through PIL, whose
encoder is libjpeg-turbo) for the same pixels at the reference's settings (quality 80, 4:2:0, baseline, standard tables:
what Magnum's JpegImageConverter asks for, src/image_saver.cpp:55-97). Covers whole / ragged MCU grids (dummy blocks, edge
replication), 1x1, grey, RGB, noise (long codes, 0xFF stuffing) and smooth content (runs, EOB, ZRL)."""
import io
import os
import sys

import numpy as np
import pytest
from PIL import Image

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import jpeg_np  # noqa: E402


def libjpeg_bytes(img, quality=80):
    b = io.BytesIO()
    Image.fromarray(img).save(b, format="JPEG", quality=quality, optimize=False, **({"subsampling": "4:2:0"} if img.ndim == 3 else {}))
    return b.getvalue()


def make_image(H, W, C, kind, seed):
    rng = np.random.RandomState(seed)
    if kind == "noise":
        img = rng.randint(0, 256, (H, W, C))
    elif kind == "flat":
        img = np.full((H, W, C), 255) * (np.arange(C) != 1)            # saturated: long runs, stuffing candidates
    else:
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.stack([128 + 100 * np.sin(xx / 9.0 + k) + 20 * np.cos(yy / 5.0 * k) for k in range(1, C + 1)], -1) + rng.normal(0, 6, (H, W, C))
    img = np.clip(img, 0, 255).astype(np.uint8)
    return img[..., 0] if C == 1 else img


SHAPES = [(16, 16, 3), (17, 23, 3), (25, 40, 3), (40, 25, 3), (8, 8, 3), (1, 1, 3), (64, 64, 1), (13, 9, 1), (48, 64, 3), (33, 70, 3), (120, 160, 3)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["smooth", "noise", "flat"])
def test_restatement_writes_libjpegs_bytes(shape, kind):
    img = make_image(*shape, kind, seed=sum(shape))
    assert jpeg_np.encode(img) == libjpeg_bytes(img)


@pytest.mark.parametrize("quality", [1, 25, 50, 75, 95, 100])
def test_other_qualities(quality):
    img = make_image(40, 56, 3, "smooth", 3)
    assert jpeg_np.encode(img, quality) == libjpeg_bytes(img, quality)


def test_alpha_is_ignored_and_file_decodes():
    img = make_image(32, 32, 3, "smooth", 5)
    rgba = np.concatenate([img, np.full((32, 32, 1), 77, np.uint8)], -1)
    data = jpeg_np.encode(rgba)
    assert data == jpeg_np.encode(img)
    dec = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(int)
    assert np.abs(dec - img.astype(int)).mean() < 6
