"""Mirror of the hot-path functions of the reference's utils/convert_pose.py."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from .engine import WrongInputException, as_torch, get_plan, require_cuda_f32


def pose_rvec2matr_batch_tf(poses):
    """reference utils/convert_pose.py:32-71: [B,N,6] (t, rotation vector) -> [B,N,4,4]
    (negated skew matrix, identity below |theta| < 1e-8).  Forward only."""
    poses = as_torch(poses)
    require_cuda_f32(poses=poses)
    B, N, _ = poses.shape
    plan = get_plan(poses.device.index or 0, B, N, 2, 2, [1])
    return plan.pose_rvec2matr(poses)


pose_rvec2matr_batch = pose_rvec2matr_batch_tf


def pose_matr2rvec_batch(poses, invert=False):
    """reference utils/convert_pose.py:151-168: [B,N,4,4] -> [B,N,6] = (t, rotation vector).
    invert=True applies tf.linalg.inv first (losses.py:120: stereo_T_LR -> T_RL).  Forward only:
    the reference feeds it dataset transforms, never predictions."""
    poses = as_torch(poses)
    require_cuda_f32(poses=poses)
    if poses.dim() != 4 or tuple(poses.shape[-2:]) != (4, 4):
        raise WrongInputException(f"poses must be [batch, numsrc, 4, 4], got {tuple(poses.shape)}")
    poses = poses.contiguous()
    B, N = poses.shape[:2]
    out = torch.empty((B, N, 6), dtype=torch.float32, device=poses.device)
    st = C.c_void_p(torch.cuda.current_stream(poses.device).cuda_stream)
    _cabi.check(_cabi.lib().xpt_pose_matr2rvec(poses.device.index or 0, poses.data_ptr(), B * N, int(bool(invert)),
                                               out.data_ptr(), st))
    return out
