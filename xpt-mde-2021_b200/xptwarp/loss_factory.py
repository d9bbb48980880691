"""Mirror of the reference's model/loss_and_metric/loss_factory.py."""
from __future__ import annotations

import numpy as np

from . import losses as lm


def loss_factory(dataset_cfg, loss_weights, scale_weights, stereo=False, weights_to_regularize=None, batch_size=1):
    """reference loss_factory.py:6-52.  `stereo` / `batch_size` are explicit here: the reference
    binds them to opts.* at import time; batch_size is the GLOBAL batch."""
    scale_weights = np.asarray(scale_weights, dtype=np.float32).reshape(-1, 1)
    loss_pool = {
        "L1": lm.PhotometricLossMultiScale("L1", scale_weights),
        "L1_R": lm.PhotometricLossMultiScale("L1", scale_weights, key_suffix="_R"),
        "SSIM": lm.PhotometricLossMultiScale("SSIM", scale_weights),
        "SSIM_R": lm.PhotometricLossMultiScale("SSIM", scale_weights, key_suffix="_R"),
        "smoothe": lm.SmoothenessLossMultiScale(scale_weights),
        "smoothe_R": lm.SmoothenessLossMultiScale(scale_weights, key_suffix="_R"),
        "md2L1": lm.MonoDepth2LossMultiScale("L1", scale_weights),
        "md2L1_R": lm.MonoDepth2LossMultiScale("L1", scale_weights, key_suffix="_R"),
        "md2SSIM": lm.MonoDepth2LossMultiScale("SSIM", scale_weights),
        "md2SSIM_R": lm.MonoDepth2LossMultiScale("SSIM", scale_weights, key_suffix="_R"),
        "moaL1": lm.MoALossMultiScale("L1", scale_weights),
        "moaL1_R": lm.MoALossMultiScale("L1", scale_weights, key_suffix="_R"),
        "moaSSIM": lm.MoALossMultiScale("SSIM", scale_weights),
        "moaSSIM_R": lm.MoALossMultiScale("SSIM", scale_weights, key_suffix="_R"),
        "stereoL1": lm.StereoDepthLoss("L1", scale_weights),
        "stereoSSIM": lm.StereoDepthLoss("SSIM", scale_weights),
        "stereoPose": lm.StereoPoseLoss(),
        "cmbL1": lm.CombinedLossMultiScale("L1", scale_weights),
        "cmbL1_R": lm.CombinedLossMultiScale("L1", scale_weights, key_suffix="_R"),
        "cmbSSIM": lm.CombinedLossMultiScale("SSIM", scale_weights),
        "cmbSSIM_R": lm.CombinedLossMultiScale("SSIM", scale_weights, key_suffix="_R"),
        "flowL2": lm.FlowWarpLossMultiScale("L2", scale_weights),
        "flowL2_R": lm.FlowWarpLossMultiScale("L2", scale_weights, key_suffix="_R"),
        "flow_reg": lm.L2Regularizer(weights_to_regularize),
    }
    losses, weights = dict(), dict()
    for name, weight in loss_weights.items():
        if weight == 0.:
            continue
        if not check_loss_dependency(name, dataset_cfg):
            continue
        losses[name] = loss_pool[name]
        weights[name] = weight
    return lm.TotalLoss(losses, weights, stereo, batch_size)


def check_loss_dependency(loss_key, dataset_cfg):
    """reference loss_factory.py:55-74."""
    loss_dependency = [(["L1", "SSIM", "smoothe", "flowL2", "flow_reg"], ["image", "intrinsic"]),
                       (["L1_R", "SSIM_R", "smoothe_R", "flowL2_R"], ["image_R", "intrinsic_R"]),
                       (["stereoL1", "stereoSSIM", "stereoPose"],
                        ["image", "intrinsic", "image_R", "intrinsic_R", "stereo_T_LR"])]
    dependents = []
    for loss_names, data_names in loss_dependency:
        if loss_key in loss_names:
            dependents = data_names
    return all(dep in dataset_cfg for dep in dependents)
