"""stillleben.diff — see stillleben_b200/diff.py (the render-and-compare backward as one fused CUDA pass)."""
from stillleben_b200.diff import (apply_pose_delta, backpropagate_gradient_to_poses, compute_image_space_gradients,  # noqa: F401
                                  dilate_object_mask, generate_sobel_valid_mask)

__all__ = ["compute_image_space_gradients", "backpropagate_gradient_to_poses", "apply_pose_delta"]
