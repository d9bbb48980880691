"""Mirror of the hot-path helpers of the reference's utils/util_funcs.py."""
from __future__ import annotations

import torch

from .engine import WrongInputException, as_torch, get_plan, require_cuda_f32, infer_scales


def multi_scale_like_depth(image, depth_ms):
    """reference utils/util_funcs.py:163-175: target pyramid (tf.image.resize bilinear,
    half-pixel centres) sized like each depth map.  image [B,H,W,3] -> list of [B,H_s,W_s,3]."""
    image = as_torch(image)
    depth_ms = [as_torch(d) for d in depth_ms]
    require_cuda_f32(image=image)
    B, H, W, _ = image.shape
    plan = get_plan(image.device.index or 0, B, 1, H, W, infer_scales(H, depth_ms))
    # the pyramid kernel wants a source tensor too: pass the target as a 1-frame source view
    return plan.build_pyramids(image.unsqueeze(1), image)


def multi_scale_like_flow(image, flow_ms):
    """reference utils/util_funcs.py:178-190: the target resized to each flow level's size.
    image [B,H,W,3], flow_ms list of [B,N,H_s,W_s,2] -> list of [B,H_s,W_s,3]."""
    image = as_torch(image)
    flow_ms = [as_torch(f) for f in flow_ms]
    require_cuda_f32(image=image)
    B, H, W, _ = image.shape
    scales = []
    for f in flow_ms:
        if f.shape[2] <= 0 or H % f.shape[2]:
            raise WrongInputException(f"flow height {f.shape[2]} does not divide the image height {H}")
        scales.append(H // f.shape[2])
    plan = get_plan(image.device.index or 0, B, 1, H, W, scales)
    return plan.build_pyramids(image.unsqueeze(1), image)


def safe_reciprocal_number(src_tensor):
    """reference utils/util_funcs.py:157-160 (the step before the path; plain elementwise
    torch, kept so that callers find the same helper)."""
    src_tensor = as_torch(src_tensor)
    mask = (src_tensor > 0.00001).to(src_tensor.dtype)
    return (1.0 / src_tensor) * mask


def safe_reciprocal_number_ms(src_ms):
    """reference utils/util_funcs.py:146-154."""
    return [safe_reciprocal_number(s) for s in src_ms]
