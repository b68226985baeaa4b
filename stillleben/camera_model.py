"""stillleben.camera_model — see stillleben_b200/camera_model.py (fused CUDA kernels behind the reference's entry points)."""
from stillleben_b200.camera_model import *  # noqa: F401,F403
from stillleben_b200.camera_model import __all__  # noqa: F401
