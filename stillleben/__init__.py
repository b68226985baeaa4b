"""`import stillleben as sl` — the reference's Python package name, served by the B200-native implementation.

The reference builds this package around a pybind11 module (python/stillleben/__init__.py:12-42: `from
.lib.libstillleben_python import *`). Here the same names come from `stillleben_b200.sl` (host logic over the C ABI of
include/slb.h; rendering always runs in stillleben_b200/libslb.so), so scripts written for the reference — its own
tests/test_python.py and tests/test_grad.py, examples/ycb.py — run unmodified:

    import stillleben as sl
    sl.init_cuda()
    ...
    result = sl.RenderPass().render(scene)

Not provided (out of scope for the render path): Animator, ImageLoader, Viewer, the `extension`
build helper, PhysX. `Scene.simulate_tabletop_scene()` arranges the objects with a NON-PHYSICAL placement
sampler and says so in a warning (see stillleben_b200/sl.py); `simulate()` / `check_collisions()` raise.
"""
import torch  # noqa: F401  (the reference imports torch first as well)

from stillleben_b200 import camera_model, diff  # noqa: F401
from stillleben_b200.image_saver import ImageSaver  # noqa: F401
from stillleben_b200.sl import (LightMap, Mesh, MeshCache, Object, Range3D, RenderPass, RenderPassResult, Scene, Texture, Texture2D,  # noqa: F401
                                init, init_cuda, matrix_to_quat, quat_to_matrix, render_debug_image, view)
from . import losses, profiling  # noqa: F401

__all__ = ["init", "init_cuda", "render_debug_image", "ImageSaver", "LightMap", "Mesh", "MeshCache", "Object", "Range3D", "RenderPass",
           "RenderPassResult", "Scene", "Texture", "Texture2D", "view", "camera_model", "diff", "losses", "quat_to_matrix", "matrix_to_quat"]
