"""Mirror of the reference's model/loss_and_metric/losses.py call surface for the hot path:
TotalLoss, PhotometricLossMultiScale, SmoothenessLossMultiScale (+ the per-scale functions of
loss_util.py through the same kernels)."""
from __future__ import annotations

import torch

from . import _cabi
from .engine import WrongInputException, as_torch, get_plan, infer_scales, require_cuda_f32
from .synthesize import SynthesizeMultiScale
from .util_funcs import multi_scale_like_depth

_FUSED_SET = ("L1", "SSIM", "smoothe")


def _scale_weights_list(scale_weights):
    if scale_weights is None:
        return None
    if isinstance(scale_weights, torch.Tensor):
        return [float(v) for v in scale_weights.reshape(-1).tolist()]
    return [float(v) for v in list(getattr(scale_weights, "reshape", lambda *_: scale_weights)(-1))]


class _TotalLossFn(torch.autograd.Function):
    """loss + every gradient in ONE fused launch; backward only rescales the stored gradients
    (the loss is linear in its upstream gradient)."""

    @staticmethod
    def forward(ctx, plan, want_grad, image5d, intrinsic, pose, *maps):
        S = plan.S
        depth_ms, disp_ms = maps[:S], (maps[S:] if len(maps) > S else None)
        source, target = image5d[:, :-1], image5d[:, -1]
        r = plan.total_loss(source, target, intrinsic, depth_ms, disp_ms, pose, want_grad=want_grad)
        losses = r["losses"]
        if want_grad:
            ctx.shapes = [t.shape for t in maps]
            ctx.have_disp = disp_ms is not None
            ctx.save_for_backward(r["d_pose"], *r["d_depth_ms"], *(r["d_disp_ms"] if disp_ms is not None else ()))
        ctx.want_grad = want_grad
        by_type = losses[1:4].clone()
        ctx.mark_non_differentiable(by_type)
        return losses[0].clone(), by_type

    @staticmethod
    def backward(ctx, g_total, _g_by_type):
        if not ctx.want_grad:
            raise RuntimeError("TotalLoss was evaluated without gradients")
        d_pose, *d_maps = ctx.saved_tensors
        outs = [g_total * d.reshape(s) for d, s in zip(d_maps, ctx.shapes)]
        return (None, None, None, None, g_total * d_pose, *outs)


class _PhotoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, method, S, *ts):
        synth_ms, target_ms = ts[:S], ts[S:]
        loss, _ = plan.photometric_loss(method, synth_ms, target_ms)
        ctx.plan, ctx.method, ctx.S = plan, method, S
        ctx.save_for_backward(*ts)
        return loss

    @staticmethod
    def backward(ctx, g):
        ts = ctx.saved_tensors
        S = ctx.S
        _, d_synth = ctx.plan.photometric_loss(ctx.method, ts[:S], ts[S:], grad_loss_batch=g.reshape(-1).contiguous(),
                                               want_grad=True)
        return (None, None, None, *d_synth, *([None] * S))


class _SmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, S, *ts):
        loss, _ = plan.smoothness_loss(ts[:S], ts[S:])
        ctx.plan, ctx.S = plan, S
        ctx.save_for_backward(*ts)
        return loss

    @staticmethod
    def backward(ctx, g):
        ts = ctx.saved_tensors
        S = ctx.S
        _, d_disp = ctx.plan.smoothness_loss(ts[:S], ts[S:], grad_loss_batch=g.reshape(-1).contiguous(), want_grad=True)
        d_disp = [d.reshape(t.shape) for d, t in zip(d_disp, ts[:S])]
        return (None, None, *d_disp, *([None] * S))


class LossBase:
    """reference losses.py:143-154."""

    def __call__(self, features, predictions, augm_data):
        raise NotImplementedError()


class PhotometricLoss(LossBase):
    """reference losses.py:157-172."""
    _METHODS = {"L1": _cabi.XPT_PHOTO_L1, "L2": _cabi.XPT_PHOTO_L2, "SSIM": _cabi.XPT_PHOTO_SSIM}

    def __init__(self, method, scale_weights, key_suffix=""):
        if method not in self._METHODS:
            raise WrongInputException("Wrong photometric loss name: " + method)
        self.method = method
        self.key_suffix = key_suffix
        self.scale_weights = scale_weights


class PhotometricLossMultiScale(PhotometricLoss):
    """reference losses.py:175-195: mean over sources of the per-pixel term, weighted sum over
    scales -> [batch] (the reference returns [batch, 1]; only its sum is ever used)."""

    def __call__(self, features, predictions, augm_data):
        target_ms = [as_torch(t) for t in augm_data["target_ms" + self.key_suffix]]
        synth_ms = [as_torch(t) for t in augm_data["synth_target_ms" + self.key_suffix]]
        require_cuda_f32(synth_target_ms=synth_ms, target_ms=target_ms)
        B, N, H, W, _ = synth_ms[0].shape
        scales = [H // s.shape[2] for s in synth_ms]
        plan = get_plan(synth_ms[0].device.index or 0, B, N, H, W, scales, _scale_weights_list(self.scale_weights))
        return _PhotoFn.apply(plan, self._METHODS[self.method], plan.S, *synth_ms, *target_ms)


class SmoothenessLossMultiScale(LossBase):
    """reference losses.py:386-440."""

    def __init__(self, scale_weights, key_suffix=""):
        self.key_suffix = key_suffix
        self.scale_weights = scale_weights

    def __call__(self, features, predictions, augm_data):
        disp_ms = [as_torch(t) for t in predictions["disp_ms" + self.key_suffix]]
        target_ms = [as_torch(t) for t in augm_data["target_ms" + self.key_suffix]]
        require_cuda_f32(disp_ms=disp_ms, target_ms=target_ms)
        B, H, W, _ = target_ms[0].shape
        scales = [H // t.shape[1] for t in target_ms]
        plan = get_plan(target_ms[0].device.index or 0, B, 1, H, W, scales, _scale_weights_list(self.scale_weights))
        return _SmoothFn.apply(plan, plan.S, *disp_ms, *target_ms)


class _OutsideHotPath(LossBase):
    def __init__(self, name):
        self.name = name

    def __call__(self, features, predictions, augm_data):
        raise WrongInputException(f"loss {self.name!r} is outside the B200 hot path of this build "
                                  "(SURVEY.md section 8f lists it as a next row)")


class TotalLoss:
    """reference losses.py:14-55.  `batch_size` is the GLOBAL batch of compute_average_loss."""

    def __init__(self, loss_objects=None, loss_weights=None, stereo=False, batch_size=1):
        self.loss_objects = loss_objects
        self.loss_weights = loss_weights
        self.stereo = stereo
        self.batch_size = batch_size

    def _fused_ok(self, predictions, features):
        if not self.loss_objects or any(k not in _FUSED_SET for k in self.loss_objects):
            return False
        if self.stereo and ("image5d_R" in features):
            return False
        sw = None
        for obj in self.loss_objects.values():
            w = _scale_weights_list(getattr(obj, "scale_weights", None))
            if sw is not None and w != sw:
                return False
            sw = w
        return "depth_ms" in predictions and "pose" in predictions and "flow_ms" not in predictions

    def __call__(self, predictions, features):
        """
        :param predictions: {"depth_ms": [...], "disp_ms": [...], "pose": [batch, numsrc, 6]}
        :param features: {"image5d": [batch, snippet, height, width, 3], "intrinsic": [batch, 3, 3]}
        :return: total_loss (scalar), loss_by_type {name: unweighted mean}
        """
        if self._fused_ok(predictions, features):
            return self._call_fused(predictions, features)
        augm_data = self.append_data(features, predictions)
        losses, loss_by_type = [], dict()
        for loss_name in self.loss_objects:
            loss_batch = self.loss_objects[loss_name](features, predictions, augm_data)
            loss_mean = loss_batch.sum() / self.batch_size          # tf.nn.compute_average_loss
            losses.append(loss_mean * self.loss_weights[loss_name])
            loss_by_type[loss_name] = loss_mean
        return torch.stack(losses).sum(), loss_by_type

    def _call_fused(self, predictions, features):
        image5d = as_torch(features["image5d"])
        intrinsic = as_torch(features["intrinsic"])
        depth_ms = [as_torch(d) for d in predictions["depth_ms"]]
        pose = as_torch(predictions["pose"])
        require_cuda_f32(image5d=image5d, intrinsic=intrinsic, pose=pose, depth_ms=depth_ms)
        B, F, H, W, _ = image5d.shape
        w = {k: float(self.loss_weights[k]) for k in self.loss_objects}
        sw = _scale_weights_list(next(iter(self.loss_objects.values())).scale_weights)
        plan = get_plan(image5d.device.index or 0, B, F - 1, H, W, infer_scales(H, depth_ms), sw,
                        w.get("L1", 0.0), w.get("SSIM", 0.0), w.get("smoothe", 0.0), self.batch_size)
        maps = list(depth_ms)
        if "smoothe" in w:
            maps += [as_torch(d) for d in predictions["disp_ms"]]
        want_grad = torch.is_grad_enabled() and any(t.requires_grad for t in [pose, *maps])
        total, by_type = _TotalLossFn.apply(plan, want_grad, image5d, intrinsic, pose, *maps)
        names = ("L1", "SSIM", "smoothe")
        return total, {k: by_type[names.index(k)] for k in self.loss_objects}

    def append_data(self, features, predictions, suffix=""):
        """reference losses.py:57-103 (the target frame is the LAST one of the snippet)."""
        image5d = as_torch(features["image5d" + suffix])
        intrinsic = as_torch(features["intrinsic" + suffix])
        source_image = image5d[:, :-1]
        target_image = image5d[:, -1]
        augm_data = {"source" + suffix: source_image, "target" + suffix: target_image}
        if ("depth_ms" + suffix in predictions) and ("pose" + suffix in predictions):
            pred_depth_ms = predictions["depth_ms" + suffix]
            augm_data["target_ms" + suffix] = multi_scale_like_depth(target_image, pred_depth_ms)
            augm_data["synth_target_ms" + suffix] = SynthesizeMultiScale()(source_image, intrinsic, pred_depth_ms,
                                                                           predictions["pose" + suffix])
        if "flow_ms" + suffix in predictions:
            raise WrongInputException("flow_ms: FlowWarpMultiScale is outside the B200 hot path of this build")
        return augm_data
