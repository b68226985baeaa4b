"""The PNG encoder's CPU restatement (oracle/png_np.py) produces standard PNG files: an independent decoder
(PIL; zlib for the raw stream) returns exactly the input pixels for every format the reference's saver accepts
(python/src/py_image_saver.cpp:50-95), and rejects nothing."""
import io
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import png_np  # noqa: E402

Image = pytest.importorskip("PIL.Image")


def images(seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:37, 0:53]
    smooth = ((np.sin(xx / 9.0) * 0.5 + 0.5) * 200 + yy).astype(np.uint8)
    out = {
        "grey_noise": rng.randint(0, 256, (21, 34)).astype(np.uint8),
        "grey_flat_runs": np.repeat(rng.randint(0, 4, (16, 6)).astype(np.uint8), 50, axis=1),       # runs up to > 258
        "rgb_smooth": np.stack([smooth, smooth[::-1], smooth.T[:37, :53] if False else (smooth // 2)], -1),
        "rgba_mixed": np.concatenate([rng.randint(0, 256, (18, 25, 3)).astype(np.uint8), np.full((18, 25, 1), 255, np.uint8)], -1),
        "rgb_flat": np.full((9, 300, 3), 7, np.uint8),
        "u16_ids": (rng.randint(0, 5, (24, 40)) * 13000).astype(np.uint16),
        "one_pixel": np.array([[200]], np.uint8),
        "segmentation": np.repeat(np.repeat(rng.randint(0, 21, (6, 8)).astype(np.uint8), 20, 0), 20, 1),
    }
    return out


def decode(png_bytes):
    im = Image.open(io.BytesIO(png_bytes))
    im.load()
    return np.asarray(im)


@pytest.mark.parametrize("name", sorted(images()))
def test_oracle_png_decodes_to_the_input(name):
    img = images()[name]
    data = png_np.encode(img)
    got = decode(data)
    assert got.shape == img.shape
    assert np.array_equal(got.astype(np.int64), img.astype(np.int64))


def test_stream_structure():
    img = images()["rgb_flat"]
    data = png_np.encode(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[-12:] == b"\x00\x00\x00\x00IEND\xaeB`\x82"
    idat_len = int.from_bytes(data[33:37], "big")
    assert data[37:41] == b"IDAT"
    raw = zlib.decompress(data[41:41 + idat_len])
    assert len(raw) == img.shape[0] * (1 + img.shape[1] * 3) and raw[0] == 1       # filter type Sub on every scanline
    assert len(data) < img.nbytes // 4                                              # run matches do compress
