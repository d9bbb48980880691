"""Mirror of the one hot-path function of the reference's utils/convert_pose.py."""
from __future__ import annotations

from .engine import as_torch, get_plan, require_cuda_f32


def pose_rvec2matr_batch_tf(poses):
    """reference utils/convert_pose.py:32-71: [B,N,6] (t, rotation vector) -> [B,N,4,4]
    (negated skew matrix, identity below |theta| < 1e-8).  Forward only."""
    poses = as_torch(poses)
    require_cuda_f32(poses=poses)
    B, N, _ = poses.shape
    plan = get_plan(poses.device.index or 0, B, N, 2, 2, [1])
    return plan.pose_rvec2matr(poses)


pose_rvec2matr_batch = pose_rvec2matr_batch_tf
