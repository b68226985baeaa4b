"""Host-side mirror of the reference's `sl.ImageSaver` (python/src/py_image_saver.cpp:19-135, src/image_saver.cpp).

    from stillleben_b200.image_saver import ImageSaver
    with ImageSaver() as saver:
        saver.save(result.rgb(), "/tmp/rgb.png")            # HxWx3 / HxWx4 uint8, HxW uint8 or int16

Same calling convention and argument checks as the reference (which runs libpng on a pool of CPU threads);
here the images are ENCODED ON THE GPU in batches (`slb_png_encode`: one standard PNG file per image, built in
device memory) and only the finished files cross PCIe; a small thread pool writes them to disk. `save()`
returns immediately; files are complete when the `with` block exits (as in the reference).
`encode_batch()` is the batched extension: [n, H, W(, C)] tensor -> list of `bytes`.
File names ending in .jpg / .jpeg are written as baseline JPEG (`slb_jpeg_encode`: the bytes libjpeg writes at the reference's
converter default, quality 80), everything else as PNG — the choice AnyImageConverter makes from the extension
(src/image_saver.cpp:55-97).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import torch

from . import sl as _sl


def _stream(ctx):
    """torch's CURRENT stream on the context's device: the library work is ordered after the torch ops that produced the
    inputs (stack / to / contiguous) instead of racing them on a private stream."""
    import torch
    return torch.cuda.current_stream(torch.device("cuda", ctx.device)).cuda_stream


def encode_batch(images):
    """images: uint8 [n,H,W], [n,H,W,3], [n,H,W,4] or int16 / uint16 [n,H,W] (any device) -> list of n PNG files (bytes)."""
    ctx = _sl._context()
    dev = torch.device("cuda", _sl._cuda_index)
    if images.dtype == torch.uint8:
        if images.dim() == 3:
            channels = 1
        elif images.dim() == 4 and images.size(3) in (3, 4):
            channels = images.size(3)
        else:
            raise ValueError("Color images need to have shape HxWx3 or HxWx4")          # py_image_saver.cpp:52-53
        bpc = 1
    elif images.dtype in (torch.int16, torch.uint16):
        if images.dim() != 3:
            raise ValueError("Grayscale images need to be byte or short type")          # :94
        channels, bpc = 1, 2
    else:
        raise ValueError("Color images need to have type uint8" if images.dim() == 4 else "Grayscale images need to be byte or short type")
    x = images.to(dev).contiguous()
    n, H, W = x.shape[0], x.shape[1], x.shape[2]
    bound = ctx.lib.slb_png_bound(H, W, channels, bpc)
    stride = (bound + 255) // 256 * 256
    out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    sizes = torch.empty((n,), dtype=torch.int32, device=dev)
    rc = ctx.lib.slb_png_encode(ctx.h, x.data_ptr(), n, H, W, channels, bpc, out.data_ptr(), stride, sizes.data_ptr(), _stream(ctx))
    if rc != 0:
        raise RuntimeError(ctx.lib.slb_last_error(ctx.h).decode())
    ctx.synchronize()
    sz = sizes.cpu().tolist()
    top = max(sz)
    host = out[:, :top].cpu().numpy()                      # only the used prefix of every file crosses PCIe
    return [host[i, :sz[i]].tobytes() for i in range(n)]


def encode_batch_jpeg(images, quality=80, first_stride=None):
    """images: uint8 [n,H,W], [n,H,W,3] or [n,H,W,4] (alpha ignored), any device -> list of n JFIF files (bytes).
    first_stride: bytes reserved per file on the first attempt (default: half the raw image; files that do not fit are
    encoded again with the worst-case bound)."""
    ctx = _sl._context()
    dev = torch.device("cuda", _sl._cuda_index)
    if images.dtype != torch.uint8:
        raise ValueError("JPEG images need to have type uint8")          # JpegImageConverter: R8Unorm / RGB8Unorm only
    if images.dim() == 3:
        channels = 1
    elif images.dim() == 4 and images.size(3) in (3, 4):
        channels = images.size(3)
    else:
        raise ValueError("Color images need to have shape HxWx3 or HxWx4")
    x = images.to(dev).contiguous()
    n, H, W = x.shape[0], x.shape[1], x.shape[2]
    bound = ctx.lib.slb_jpeg_bound(H, W, channels)
    stride = min(bound, first_stride or (H * W * channels // 2 + 4096))      # real files are a fraction of the raw image; the bound is the retry size
    while True:
        stride = (stride + 255) // 256 * 256
        out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
        sizes = torch.empty((n,), dtype=torch.int32, device=dev)
        rc = ctx.lib.slb_jpeg_encode(ctx.h, x.data_ptr(), n, H, W, channels, int(quality), out.data_ptr(), stride, sizes.data_ptr(), _stream(ctx))
        if rc != 0:
            raise RuntimeError(ctx.lib.slb_last_error(ctx.h).decode())
        ctx.synchronize()
        sz = sizes.cpu().tolist()
        if min(sz) > 0 or stride >= bound:
            break
        stride = bound                                     # some file did not fit (sizes[i] == 0): once more with the worst case
    top = max(sz)
    host = out[:, :top].cpu().numpy()
    return [host[i, :sz[i]].tobytes() for i in range(n)]


def _is_jpeg(path):
    return path.lower().endswith((".jpg", ".jpeg"))


class ImageSaver:
    MAX_PENDING = 64

    def __init__(self):
        self._active = False
        self._pending = []           # (tensor, path)
        self._pool = None
        self._futures = []

    def __enter__(self):
        _sl._context()
        self._active = True
        self._pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4))
        return self

    def save(self, input, path):
        if not self._active:
            raise RuntimeError("Call __enter__() first")                                 # py_image_saver.cpp:39-40
        if input.dim() == 3:
            if input.size(2) not in (3, 4):
                raise ValueError("Color images need to have shape HxWx3 or HxWx4")
            if input.dtype != torch.uint8:
                raise ValueError("Color images need to have type uint8")
        elif input.dim() == 2:
            if input.dtype not in (torch.uint8, torch.int16):
                raise ValueError("Grayscale images need to be byte or short type")
        else:
            raise ValueError("Color images need to have shape HxWx3 or HxWx4")
        if _is_jpeg(os.fspath(path)) and input.dtype != torch.uint8:
            raise ValueError("JPEG images need to have type uint8")
        self._pending.append((input.detach().clone(), os.fspath(path)))      # a snapshot, like input.flipud() in py_image_saver.cpp:44
        if len(self._pending) >= self.MAX_PENDING:
            self._flush()

    def _flush(self):
        groups = {}
        for t, p in self._pending:
            groups.setdefault((tuple(t.shape), t.dtype, _is_jpeg(p)), []).append((t, p))
        self._pending = []
        for (_, _, jpeg), items in groups.items():
            batch = torch.stack([t for t, _ in items])
            files = encode_batch_jpeg(batch) if jpeg else encode_batch(batch)
            for data, (_, p) in zip(files, items):
                self._futures.append(self._pool.submit(_write, p, data))

    def __exit__(self, *exc):
        try:
            self._flush()
            for f in self._futures:
                f.result()
        finally:
            self._futures = []
            self._pool.shutdown(wait=True)
            self._pool, self._active = None, False
        return False


def _write(path, data):
    with open(path, "wb") as fh:
        fh.write(data)
