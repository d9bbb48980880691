"""Mirror of the reference's model/loss_and_metric/losses.py call surface for the hot path:
TotalLoss, PhotometricLossMultiScale, SmoothenessLossMultiScale (+ the per-scale functions of
loss_util.py through the same kernels)."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _cabi
from .convert_pose import pose_matr2rvec_batch
from .engine import WrongInputException, as_torch, get_plan, infer_scales, require_cuda_f32, scale_tensors
from .flow_warping import FlowWarpMultiScale, infer_flow_scales
from .synthesize import SynthesizeMultiScale
from .util_funcs import multi_scale_like_depth, multi_scale_like_flow

# losses the fused tile kernel evaluates (one launch per group): the temporal set of each eye, the two stereo
# syntheses of StereoDepthLoss; StereoPoseLoss is a tiny kernel of its own
_TEMPORAL = ("L1", "SSIM", "smoothe")
_FUSED_SET = _TEMPORAL + tuple(n + "_R" for n in _TEMPORAL) + ("stereoL1", "stereoSSIM", "stereoPose")


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
    def forward(ctx, plan, want_grad, split, source, target, intrinsic, pose, *maps):
        """split > 0: the batch holds two groups of snippets (first `split`, rest) -- both eyes of a rig in one
        launch -- and the by-type means come back per group, [2, 3], from the per-snippet losses."""
        S = plan.S
        depth_ms, disp_ms = maps[:S], (maps[S:] if len(maps) > S else None)
        r = plan.total_loss(source, target, intrinsic, depth_ms, disp_ms, pose, want_grad=want_grad,
                            want_loss_batch=split > 0)
        losses = r["losses"]
        if want_grad:
            ctx.shapes = [t.shape for t in maps]
            ctx.have_disp = disp_ms is not None
            ctx.save_for_backward(r["d_pose"], *r["d_depth_ms"], *(r["d_disp_ms"] if disp_ms is not None else ()))
        ctx.want_grad = want_grad
        if split > 0:
            lb = r["loss_batch"].reshape(3, plan.B)
            inv_gb = 1.0 / float(plan.cfg.global_batch)
            by_type = torch.stack([lb[:, :split].sum(dim=1), lb[:, split:].sum(dim=1)]) * inv_gb
        else:
            by_type = losses[1:4].clone()
        ctx.mark_non_differentiable(by_type)
        return losses[0].clone(), by_type

    @staticmethod
    def backward(ctx, g_total, _g_by_type):
        if not ctx.want_grad:
            raise RuntimeError("TotalLoss was evaluated without gradients")
        # the stored gradients are for an upstream of 1; the actual upstream (a device scalar) is applied to all of
        # them in ONE launch of the library instead of one PyTorch multiply per tensor
        d_pose, *d_maps = scale_tensors(list(ctx.saved_tensors), g_total)
        outs = [d.reshape(s) for d, s in zip(d_maps, ctx.shapes)]
        return (None, None, None, None, None, None, d_pose, *outs)


class _StereoPoseFn(torch.autograd.Function):
    """StereoPoseLoss forward + both gradients in one tiny launch (xpt_stereo_pose_loss)."""

    @staticmethod
    def forward(ctx, T_LR, pose_lr, pose_rl):
        T_LR, pose_lr, pose_rl = T_LR.contiguous(), pose_lr.contiguous(), pose_rl.contiguous()
        B, n = pose_lr.shape[:2]
        loss = torch.empty((B,), dtype=torch.float32, device=pose_lr.device)
        d_lr, d_rl = torch.empty_like(pose_lr), torch.empty_like(pose_rl)
        st = C.c_void_p(torch.cuda.current_stream(pose_lr.device).cuda_stream)
        _cabi.check(_cabi.lib().xpt_stereo_pose_loss(pose_lr.device.index or 0, T_LR.data_ptr(), pose_lr.data_ptr(),
                                                     pose_rl.data_ptr(), B, n, loss.data_ptr(), None,
                                                     d_lr.data_ptr(), d_rl.data_ptr(), st))
        ctx.save_for_backward(d_lr, d_rl)
        return loss

    @staticmethod
    def backward(ctx, g):
        d_lr, d_rl = ctx.saved_tensors
        g = g.reshape(-1, 1, 1)
        return None, g * d_lr, g * d_rl


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


def _scale_by_snippet(g, grads):
    """upstream dL/d loss_batch [B] applied to per-snippet gradients computed for an all-ones upstream"""
    g = g.reshape(-1)
    return [d * g.view(-1, *([1] * (d.dim() - 1))) for d in grads]


class _PhotoMinFn(torch.autograd.Function):
    """MonoDepth2 / MoA: min over sources per pixel and channel at full resolution (xpt_photometric_min_loss).
    When a gradient is wanted the forward call already runs the gradient kernel (for an all-ones upstream: the loss is
    linear in it), so the backward is a per-snippet rescale instead of a second pass over every scale."""

    @staticmethod
    def forward(ctx, plan, method, S, have_stereo, target, *ts):
        synth_ms = ts[:S]
        stereo_ms = ts[S:] if have_stereo else None
        want = any(ctx.needs_input_grad[5:])
        loss, d_synth, d_stereo, _ = plan.photometric_min_loss(method, synth_ms, stereo_ms, target, want_grad=want)
        ctx.S, ctx.have_stereo, ctx.want = S, have_stereo, want
        if want:
            ctx.save_for_backward(*d_synth, *(d_stereo if have_stereo else ()))
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, None, None, *_scale_by_snippet(g, ctx.saved_tensors))


class _PhotoPairFn(torch.autograd.Function):
    """The "L1" and the "SSIM" loss object of one loss set (moaL1 + moaSSIM, md2L1 + md2SSIM: xpt_photometric_min_pair_loss;
    cmbL1 + cmbSSIM: xpt_photometric_cmb_pair_loss) in ONE launch.  Returns their weighted contribution to the total loss
    and the two unweighted means; the gradient of the contribution is formed by the same launch (the weights are known
    here, as in _TotalLossFn) and the backward applies the upstream scalar to all stored gradients in one launch of the
    library.  `extra`: the stereo syntheses (MoA), none (MonoDepth2) or the flow-warped view (Combined; no gradient)."""

    @staticmethod
    def forward(ctx, plan, kind, S, w_l1, w_ssim, batch_size, target, *ts):
        synth_ms, extra = ts[:S], ts[S:]
        want = any(ctx.needs_input_grad[7:])
        c1, c2 = w_l1 / batch_size, w_ssim / batch_size
        if kind == "cmb":
            loss2, d_synth = plan.photometric_cmb_pair_loss(synth_ms, extra[0], target, c1, c2, want_grad=want)
            grads = list(d_synth) + [None] if want else None
            n_saved = S
        else:
            loss2, d_synth, d_stereo = plan.photometric_min_pair_loss(synth_ms, extra if extra else None, target, c1, c2,
                                                                      want_grad=want)
            grads = list(d_synth) + (list(d_stereo) if extra else []) if want else None
            n_saved = len(grads) if want else 0
        ctx.want, ctx.n_saved, ctx.n_in = want, n_saved, len(ts)
        if want:
            ctx.save_for_backward(*[g for g in grads if g is not None])
        means = loss2.sum(dim=1) / batch_size                       # tf.nn.compute_average_loss, per loss
        ctx.mark_non_differentiable(means)
        return means[0] * w_l1 + means[1] * w_ssim, means

    @staticmethod
    def backward(ctx, g, _g_means):
        if not ctx.want:
            raise RuntimeError("the loss pair was evaluated without gradients")
        scaled = scale_tensors(list(ctx.saved_tensors), g)
        return (None, None, None, None, None, None, None, *scaled, *([None] * (ctx.n_in - ctx.n_saved)))


class _PhotoCmbFn(torch.autograd.Function):
    """CombinedLossMultiScale: static term where it beats the flow term (xpt_photometric_cmb_loss).  The flow-warped
    view only enters through a comparison, so it receives no gradient (as in the reference's graph).  Forward and
    gradient share one launch, like _PhotoMinFn."""

    @staticmethod
    def forward(ctx, plan, method, target, warped, *synth_ms):
        want = any(ctx.needs_input_grad[4:])
        loss, d_synth, _ = plan.photometric_cmb_loss(method, synth_ms, warped, target, want_grad=want)
        if want:
            ctx.save_for_backward(*d_synth)
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, None, *_scale_by_snippet(g, ctx.saved_tensors))


class _L2RegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *weights):
        loss, _ = plan.l2_regularizer(weights)
        ctx.plan = plan
        ctx.save_for_backward(*weights)
        return loss

    @staticmethod
    def backward(ctx, g):
        _, d_w = ctx.plan.l2_regularizer(ctx.saved_tensors, grad_loss=g.reshape(1).contiguous(), want_grad=True)
        return (None, *d_w)


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


class MonoDepth2LossMultiScale(PhotometricLoss):
    """reference losses.py:198-232: every scale's synthesis is up-sampled to the original size, the per-pixel
    photometric term is taken against the full-resolution target and the minimum over the sources is averaged."""

    STEREO_KEY = None

    def _inputs(self, augm_data):
        synth_ms = [as_torch(t) for t in augm_data["synth_target_ms" + self.key_suffix]]
        stereo_ms = [as_torch(t) for t in augm_data[self.STEREO_KEY]] if self.STEREO_KEY else None
        target = as_torch(augm_data["target" + self.key_suffix])
        require_cuda_f32(synth_target_ms=synth_ms, target=target)
        B, N, H, W = synth_ms[0].shape[0], synth_ms[0].shape[1], target.shape[1], target.shape[2]
        scales = [H // s.shape[2] for s in synth_ms]
        plan = get_plan(target.device.index or 0, B, N, H, W, scales, _scale_weights_list(self.scale_weights))
        return plan, target, synth_ms, stereo_ms

    def __call__(self, features, predictions, augm_data):
        plan, target, synth_ms, stereo_ms = self._inputs(augm_data)
        return _PhotoMinFn.apply(plan, self._METHODS[self.method], plan.S, stereo_ms is not None, target,
                                 *synth_ms, *(stereo_ms or ()))


class MoALossMultiScale(MonoDepth2LossMultiScale):
    """reference losses.py:282-321: minimum over the temporal syntheses AND the stereo synthesis.  As in the
    reference the stereo entry is always augm_data["stereo_synth_ms"] (the left one), also for key_suffix "_R"."""

    STEREO_KEY = "stereo_synth_ms"


class CombinedLossMultiScale(PhotometricLoss):
    """reference losses.py:235-279: the static (depth + pose) photometric term of every scale, kept only where it
    is smaller than the optical-flow term of warped_target_ms[0]; both compared at the original resolution."""

    def _inputs(self, augm_data):
        synth_ms = [as_torch(t) for t in augm_data["synth_target_ms" + self.key_suffix]]
        warped = as_torch(augm_data["warped_target_ms" + self.key_suffix][0])
        target = as_torch(augm_data["target" + self.key_suffix])
        require_cuda_f32(synth_target_ms=synth_ms, warped_target_ms=warped, target=target)
        B, N, H, W = synth_ms[0].shape[0], synth_ms[0].shape[1], target.shape[1], target.shape[2]
        scales = [H // s.shape[2] for s in synth_ms]
        plan = get_plan(target.device.index or 0, B, N, H, W, scales, _scale_weights_list(self.scale_weights))
        return plan, target, synth_ms, [warped.detach()]

    def __call__(self, features, predictions, augm_data):
        plan, target, synth_ms, (warped,) = self._inputs(augm_data)
        return _PhotoCmbFn.apply(plan, self._METHODS[self.method], target, warped, *synth_ms)


class FlowWarpLossMultiScale(PhotometricLoss):
    """reference losses.py:497-519: photometric term ("L2" in the pool) between every flow-warped view and the
    target resized like the flow, mean over sources, weighted sum over the flow scales."""

    def __call__(self, features, predictions, augm_data):
        flow_target_ms = [as_torch(t) for t in augm_data["flow_target_ms" + self.key_suffix]]
        warped_ms = [as_torch(t) for t in augm_data["warped_target_ms" + self.key_suffix]]
        require_cuda_f32(warped_target_ms=warped_ms, flow_target_ms=flow_target_ms)
        B, N = warped_ms[0].shape[:2]
        # the plan's "full resolution" is the first flow level: only the level sizes matter to the loss kernel
        H, W = warped_ms[0].shape[2], warped_ms[0].shape[3]
        scales = [H // t.shape[2] for t in warped_ms]
        plan = get_plan(warped_ms[0].device.index or 0, B, N, H, W, scales, _scale_weights_list(self.scale_weights))
        return _PhotoFn.apply(plan, self._METHODS[self.method], plan.S, *warped_ms, *flow_target_ms)


class L2Regularizer(LossBase):
    """reference losses.py:522-534 ("flow_reg"): sum of tf.nn.l2_loss over the given weights, tiled to [batch]."""

    def __init__(self, weights_to_regularize):
        self.weights = weights_to_regularize

    def __call__(self, features, predictions, augm_data):
        if not self.weights:
            raise WrongInputException("flow_reg needs weights_to_regularize (loss_factory(..., weights_to_regularize=...))")
        ws = [as_torch(w) for w in self.weights]
        require_cuda_f32(weights_to_regularize=ws)
        batch = as_torch(features["image5d"]).shape[0]
        plan = get_plan(ws[0].device.index or 0, 1, 1, 8, 8, [1])
        return _L2RegFn.apply(plan, *ws).expand(batch)


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


class StereoDepthLoss(PhotometricLoss):
    """reference losses.py:443-478: photometric loss of the left view synthesised from the right frame plus
    that of the right view synthesised from the left frame (both N = 1), summed per scale, then merged."""

    def __init__(self, method, scale_weights):
        super().__init__(method, scale_weights)

    def __call__(self, features, predictions, augm_data):
        left = PhotometricLossMultiScale(self.method, self.scale_weights)(
            features, predictions, {"synth_target_ms": augm_data["stereo_synth_ms"], "target_ms": augm_data["target_ms"]})
        right = PhotometricLossMultiScale(self.method, self.scale_weights)(
            features, predictions, {"synth_target_ms": augm_data["stereo_synth_ms_R"], "target_ms": augm_data["target_ms_R"]})
        return left + right


class StereoPoseLoss(LossBase):
    """reference losses.py:481-495."""

    def __call__(self, features, predictions, augm_data):
        T = as_torch(features["stereo_T_LR"])
        lr, rl = as_torch(predictions["pose_LR"]), as_torch(predictions["pose_RL"])
        require_cuda_f32(stereo_T_LR=T, pose_LR=lr, pose_RL=rl)
        return _StereoPoseFn.apply(T, lr, rl)


class TotalLoss:
    """reference losses.py:14-55.  `batch_size` is the GLOBAL batch of compute_average_loss."""

    def __init__(self, loss_objects=None, loss_weights=None, stereo=False, batch_size=1):
        self.loss_objects = loss_objects
        self.loss_weights = loss_weights
        self.stereo = stereo
        self.batch_size = batch_size

    def _stereo_on(self, features):
        """losses.py:38: the right eye and the stereo syntheses are used when the rig flag is set and the
        dataset carries right frames (the reference tests the key "image_R" and then reads "image5d_R")."""
        return bool(self.stereo) and ("image_R" in features or "image5d_R" in features)

    @staticmethod
    def _is_stock(name, obj):
        """the fused launch stands for the loss OBJECT the reference's loop would call, so the object under a name
        must be the stock one (type, method, key suffix): TotalLoss({"L1": MonoDepth2LossMultiScale(...)}) takes the
        generic loop below instead of silently being scored as plain L1"""
        sfx = "_R" if name.endswith("_R") else ""
        base = name[:-2] if sfx else name
        if base in ("L1", "SSIM"):
            return type(obj) is PhotometricLossMultiScale and obj.method == base and obj.key_suffix == sfx
        if base == "smoothe":
            return type(obj) is SmoothenessLossMultiScale and obj.key_suffix == sfx
        if base in ("stereoL1", "stereoSSIM"):
            return type(obj) is StereoDepthLoss and obj.method == base[6:]
        if base == "stereoPose":
            return type(obj) is StereoPoseLoss
        return False

    def _min_pairs(self):
        """{L1 name: SSIM name, SSIM name: None} for every stock (moa|md2|cmb)L1 + (moa|md2|cmb)SSIM pair of one eye with
        the same scale weights: the pair shares one launch (xpt_photometric_min_pair_loss / xpt_photometric_cmb_pair_loss)"""
        pairs = {}
        for sfx in ("", "_R"):
            for base, cls in (("moa", MoALossMultiScale), ("md2", MonoDepth2LossMultiScale), ("cmb", CombinedLossMultiScale)):
                n1, n2 = base + "L1" + sfx, base + "SSIM" + sfx
                o1, o2 = self.loss_objects.get(n1), self.loss_objects.get(n2)
                if (type(o1) is cls and type(o2) is cls and o1.method == "L1" and o2.method == "SSIM"
                        and o1.key_suffix == sfx and o2.key_suffix == sfx
                        and _scale_weights_list(o1.scale_weights) == _scale_weights_list(o2.scale_weights)):
                    pairs[n1], pairs[n2] = n2, None
        return pairs

    def _fused_ok(self, predictions, features):
        if not self.loss_objects or any(k not in _FUSED_SET for k in self.loss_objects):
            return False
        if not all(self._is_stock(k, o) for k, o in self.loss_objects.items()):
            return False
        sw = None
        for obj in self.loss_objects.values():
            w = _scale_weights_list(getattr(obj, "scale_weights", None))
            if w is None:
                continue
            if sw is not None and w != sw:
                return False
            sw = w
        have_depth = "depth_ms" in predictions or "depth_logit_ms" in predictions
        return sw is not None and have_depth and "pose" in predictions and "flow_ms" not in predictions

    def __call__(self, predictions, features):
        """
        :param predictions: {"depth_ms": [...], "disp_ms": [...], "pose": [batch, numsrc, 6]}
                            (+ the same keys with "_R", "pose_LR", "pose_RL" on a stereo rig)
        :param features: {"image5d": [batch, snippet, height, width, 3], "intrinsic": [batch, 3, 3]}
                         (+ "image5d_R", "intrinsic_R", "stereo_T_LR" [batch, 4, 4] on a stereo rig)
        :return: total_loss (scalar), loss_by_type {name: unweighted mean}
        """
        if self._fused_ok(predictions, features):
            return self._call_fused(predictions, features)
        augm_data = self.append_data(features, predictions)
        if self._stereo_on(features):
            augm_data.update(self.append_data(features, predictions, "_R"))
            augm_data.update(self.synethesize_stereo(features, predictions, augm_data))
        losses, loss_by_type = [], dict()
        pairs = self._min_pairs()
        for loss_name in self.loss_objects:
            if loss_name in pairs:
                if pairs[loss_name] is None:                          # the SSIM half: evaluated with its L1 partner
                    continue
                # moaL1 + moaSSIM (md2L1 + md2SSIM, cmbL1 + cmbSSIM) of one eye: ONE launch for both loss objects
                ssim_name, obj = pairs[loss_name], self.loss_objects[loss_name]
                plan, target, synth_ms, extra = obj._inputs(augm_data)
                part, means = _PhotoPairFn.apply(plan, "cmb" if type(obj) is CombinedLossMultiScale else "min", plan.S,
                                                 float(self.loss_weights[loss_name]), float(self.loss_weights[ssim_name]),
                                                 float(self.batch_size), target, *synth_ms, *(extra or ()))
                losses.append(part)
                loss_by_type[loss_name], loss_by_type[ssim_name] = means[0], means[1]
                continue
            loss_batch = self.loss_objects[loss_name](features, predictions, augm_data)
            loss_mean = loss_batch.sum() / self.batch_size          # tf.nn.compute_average_loss
            losses.append(loss_mean * self.loss_weights[loss_name])
            loss_by_type[loss_name] = loss_mean
        return torch.stack(losses).sum(), loss_by_type

    def _fused_group(self, source, target, intrinsic, depth_ms, disp_ms, pose, w_l1, w_ssim, w_smooth, sw, split=0,
                     logit=False):
        """one fused launch: (total contribution, [L1, SSIM, smoothe] unweighted means -- [2, 3] per group of
        snippets when split > 0)"""
        B, N, H, W, _ = source.shape
        plan = get_plan(source.device.index or 0, B, N, H, W, infer_scales(H, depth_ms), sw, w_l1, w_ssim, w_smooth,
                        self.batch_size, _cabi.XPT_FLAG_DEPTH_LOGIT if logit else 0)
        maps = list(depth_ms) + (list(disp_ms) if (w_smooth != 0.0 and disp_ms is not None) else [])
        want_grad = torch.is_grad_enabled() and any(t.requires_grad for t in [pose, *maps])
        return _TotalLossFn.apply(plan, want_grad, int(split), source, target, intrinsic, pose, *maps)

    @staticmethod
    def _same_launch(groups):
        """both eyes COULD share a launch (same weights, same shapes, disparities given for both or for neither) -- but
        only by concatenating every input of the two eyes with torch.cat first.  Measured on the stereo LOSS_RIGID_T1 step at
        config-2 size, replayed as a CUDA graph: 0.43 ms with the shared launch + cat, 0.39 ms with one launch per eye and no
        copy (profiles/r02_stereo_t1_eyes.txt), so one launch per eye is the default and XPT_EYES=cat keeps the A/B."""
        if len(groups) != 2 or os.environ.get("XPT_EYES") != "cat":
            return False
        (img_l, K_l, dep_l, dsp_l, pose_l, w_l), (img_r, K_r, dep_r, dsp_r, pose_r, w_r) = groups[""], groups["_R"]
        return (w_l == w_r and img_l.shape == img_r.shape and pose_l.shape == pose_r.shape
                and (dsp_l is None) == (dsp_r is None) and img_l.device == img_r.device
                and all(a.shape == b.shape for a, b in zip(dep_l, dep_r)))

    def _call_fused(self, predictions, features):
        w = {k: float(self.loss_weights[k]) for k in self.loss_objects}
        sw = next(v for v in (_scale_weights_list(getattr(o, "scale_weights", None)) for o in self.loss_objects.values())
                  if v is not None)
        stereo = self._stereo_on(features)
        totals, by = [], {}
        eyes, groups, logits = {}, {}, {}
        for sfx in ("", "_R"):
            names = [n + sfx for n in _TEMPORAL if n + sfx in w]
            need = names or (stereo and ("stereoL1" in w or "stereoSSIM" in w))
            if not need:
                continue
            if sfx and not stereo:
                raise WrongInputException(f"losses {names} need a stereo rig: TotalLoss(stereo=True) and image5d_R in features")
            image5d, K = as_torch(features["image5d" + sfx]), as_torch(features["intrinsic" + sfx])
            # "depth_logit_ms": the depth net's output BEFORE its last op; the kernel applies InverseSigmoidActivation
            # (model_factory.py:133-137) at load and the gradient comes back on the logits
            logit = ("depth_ms" + sfx) not in predictions
            depth_ms = [as_torch(d) for d in predictions[("depth_logit_ms" if logit else "depth_ms") + sfx]]
            require_cuda_f32(**{"image5d" + sfx: image5d, "intrinsic" + sfx: K, "depth_ms" + sfx: depth_ms})
            eyes[sfx] = (image5d, K, depth_ms)
            if not names:
                continue
            pose = as_torch(predictions["pose" + sfx])
            # no "disp_ms" in predictions: the kernel derives it from depth_ms (what model_wrappers.py:47-48 computes
            # with safe_reciprocal_number_ms) and returns the whole gradient on depth_ms
            disp_ms = ([as_torch(d) for d in predictions["disp_ms" + sfx]]
                       if ("smoothe" + sfx in w and "disp_ms" + sfx in predictions) else None)
            groups[sfx] = (image5d, K, depth_ms, disp_ms, pose,
                           (w.get("L1" + sfx, 0.0), w.get("SSIM" + sfx, 0.0), w.get("smoothe" + sfx, 0.0)))
            logits[sfx] = logit
        if any(logits.values()) and stereo:
            raise WrongInputException("depth_logit_ms is served for the temporal losses of one eye; a stereo rig needs depth_ms")
        if self._same_launch(groups):
            # both eyes of the rig in ONE fused launch (twice the batch): same kernels, half the launches and a
            # better filled last wave; the by-type means come back per eye from the per-snippet losses
            (img_l, K_l, dep_l, dsp_l, pose_l, wts), (img_r, K_r, dep_r, dsp_r, pose_r, _) = groups[""], groups["_R"]
            B = img_l.shape[0]
            cat = lambda a, b: torch.cat([a, b], dim=0)
            disp = None if dsp_l is None else [cat(a, b) for a, b in zip(dsp_l, dsp_r)]
            total, by_eye = self._fused_group(cat(img_l[:, :-1], img_r[:, :-1]), cat(img_l[:, -1], img_r[:, -1]), cat(K_l, K_r),
                                              [cat(a, b) for a, b in zip(dep_l, dep_r)], disp, cat(pose_l, pose_r),
                                              *wts, sw, split=B)
            totals.append(total)
            for e, sfx in enumerate(("", "_R")):
                for i, n in enumerate(_TEMPORAL):
                    if n + sfx in w:
                        by[n + sfx] = by_eye[e, i]
        else:
            for sfx, (image5d, K, depth_ms, disp_ms, pose, wts) in groups.items():
                total, by_type = self._fused_group(image5d[:, :-1], image5d[:, -1], K, depth_ms, disp_ms, pose, *wts, sw,
                                                   logit=logits.get(sfx, False))
                totals.append(total)
                for i, n in enumerate(_TEMPORAL):
                    if n + sfx in w:
                        by[n + sfx] = by_type[i]
        if "stereoL1" in w or "stereoSSIM" in w:
            # StereoDepthLoss (losses.py:443-478) over the two syntheses of losses.py:105-140: the fused kernel with
            # ONE source frame (the other eye's target), the rig transform as the pose and -- like the reference --
            # the LEFT intrinsics for both directions.  One launch per direction (the loss is their sum): no copies.
            if not stereo or "stereo_T_LR" not in features:
                raise WrongInputException("stereoL1 / stereoSSIM need TotalLoss(stereo=True), image5d_R and stereo_T_LR")
            T = as_torch(features["stereo_T_LR"])
            require_cuda_f32(stereo_T_LR=T)
            (img_l, K_l, depth_l), (img_r, _, depth_r) = eyes[""], eyes["_R"]
            ws = (w.get("stereoL1", 0.0), w.get("stereoSSIM", 0.0), 0.0)
            if (os.environ.get("XPT_EYES") == "cat" and img_l.shape == img_r.shape
                    and all(a.shape == b.shape for a, b in zip(depth_l, depth_r))):       # (A/B: see _same_launch)
                rig = torch.cat([pose_matr2rvec_batch(T.unsqueeze(1), invert=True), pose_matr2rvec_batch(T.unsqueeze(1))], dim=0)
                tgt = torch.cat([img_l[:, -1], img_r[:, -1]], dim=0)         # [2B,H,W,3]: left targets, then right targets
                src = torch.cat([img_r[:, -1:], img_l[:, -1:]], dim=0)       # the other eye's frame as the single source
                total, by_type = self._fused_group(src, tgt, torch.cat([K_l, K_l], dim=0),
                                                   [torch.cat([a, b], dim=0) for a, b in zip(depth_l, depth_r)], None, rig, *ws, sw)
                totals.append(total)
                st_l1, st_ssim = by_type[0], by_type[1]
            else:
                parts = []
                for src_img, tgt_img, depth_ms, inv in ((img_r, img_l, depth_l, True), (img_l, img_r, depth_r, False)):
                    rig = pose_matr2rvec_batch(T.unsqueeze(1), invert=inv)
                    parts.append(self._fused_group(src_img[:, -1:], tgt_img[:, -1], K_l, depth_ms, None, rig, *ws, sw))
                totals += [parts[0][0], parts[1][0]]
                st_l1, st_ssim = parts[0][1][0] + parts[1][1][0], parts[0][1][1] + parts[1][1][1]
            if "stereoL1" in w:
                by["stereoL1"] = st_l1
            if "stereoSSIM" in w:
                by["stereoSSIM"] = st_ssim
        if "stereoPose" in w:
            mean = self.loss_objects["stereoPose"](features, predictions, None).sum() / self.batch_size
            by["stereoPose"] = mean
            totals.append(mean * w["stereoPose"])
        return torch.stack(totals).sum(), {k: by[k] for k in self.loss_objects}

    def synethesize_stereo(self, features, predictions, augm_data):
        """reference losses.py:105-140 (the method name is the reference's)."""
        synth_stereo = dict()
        if ("stereo_T_LR" not in features) or ("depth_ms" not in predictions):
            return synth_stereo
        T = as_torch(features["stereo_T_LR"])
        K = as_torch(features["intrinsic"])
        synth_stereo["stereo_synth_ms"] = SynthesizeMultiScale()(
            augm_data["target_R"].unsqueeze(1), K, predictions["depth_ms"], pose_matr2rvec_batch(T.unsqueeze(1), invert=True))
        synth_stereo["stereo_synth_ms_R"] = SynthesizeMultiScale()(
            augm_data["target"].unsqueeze(1), K, predictions["depth_ms_R"], pose_matr2rvec_batch(T.unsqueeze(1)))
        return synth_stereo

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
            pred_flow_ms = predictions["flow_ms" + suffix]
            # flows have lower resolution than depths, so "target_ms" is not appropriate for flows (losses.py:97)
            augm_data["flow_target_ms" + suffix] = multi_scale_like_flow(target_image, pred_flow_ms)
            augm_data["warped_target_ms" + suffix] = FlowWarpMultiScale()(source_image, pred_flow_ms)
        return augm_data
