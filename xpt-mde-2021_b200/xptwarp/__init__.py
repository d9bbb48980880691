"""xptwarp -- B200-native view synthesis + photometric/smoothness loss (host side).

The arithmetic lives in libxptwarp.so (hand-written sm_100a CUDA behind the C-ABI of
include/xptwarp.h); this package mirrors the reference's Python call surface:

    reference                                             here
    model/synthesize/synthesize_base.SynthesizeMultiScale xptwarp.SynthesizeMultiScale
    model/loss_and_metric/losses.TotalLoss                xptwarp.TotalLoss
    model/loss_and_metric/losses.*LossMultiScale          xptwarp.losses.*
    model/loss_and_metric/loss_factory.loss_factory       xptwarp.loss_factory
    utils/convert_pose.pose_rvec2matr_batch_tf            xptwarp.pose_rvec2matr_batch_tf
    utils/convert_pose.pose_matr2rvec_batch               xptwarp.pose_matr2rvec_batch
    utils/util_funcs.multi_scale_like_depth               xptwarp.multi_scale_like_depth
    model/synthesize/flow_warping.FlowWarpMultiScale      xptwarp.FlowWarpMultiScale
    utils/util_funcs.multi_scale_like_flow                xptwarp.multi_scale_like_flow
"""
from .engine import Plan, WrongInputException, get_plan  # noqa: F401
from .synthesize import SynthesizeMultiScale  # noqa: F401
from .flow_warping import FlowWarpMultiScale  # noqa: F401
from .losses import (TotalLoss, PhotometricLossMultiScale, SmoothenessLossMultiScale, StereoDepthLoss,  # noqa: F401
                     StereoPoseLoss, MonoDepth2LossMultiScale, MoALossMultiScale, CombinedLossMultiScale,
                     FlowWarpLossMultiScale, L2Regularizer)
from .loss_factory import loss_factory, check_loss_dependency  # noqa: F401
from .convert_pose import pose_rvec2matr_batch_tf, pose_matr2rvec_batch  # noqa: F401
from .util_funcs import (multi_scale_like_depth, multi_scale_like_flow, safe_reciprocal_number,  # noqa: F401
                         safe_reciprocal_number_ms)
