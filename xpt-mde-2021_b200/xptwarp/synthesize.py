"""Mirror of the reference's model/synthesize/synthesize_base.py call surface."""
from __future__ import annotations

import torch

from .engine import WrongInputException, as_torch, get_plan, infer_scales, require_cuda_f32


class _SynthesizeFn(torch.autograd.Function):
    """Differentiable wrapper: backward = xpt_synthesize_backward (what tape.gradient does
    through synthesize_base.py + bilinear_interp.py in the reference, train_val.py:85)."""

    @staticmethod
    def forward(ctx, plan, want_mask, source, intrinsic, pose, *depth_ms):
        synth, mask = plan.synthesize(source, intrinsic, depth_ms, pose, want_mask=want_mask)
        ctx.plan = plan
        ctx.save_for_backward(source, intrinsic, pose, *depth_ms)
        outs = tuple(synth) + (tuple(mask) if want_mask else ())
        if want_mask:
            ctx.mark_non_differentiable(*mask)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        source, intrinsic, pose, *depth_ms = ctx.saved_tensors
        S = plan.S
        g = []
        for l in range(S):
            gl = grads[l]
            if gl is None:
                h, w = plan.level_hw(l)
                gl = torch.zeros((plan.B, plan.N, h, w, 3), dtype=torch.float32, device=plan.device)
            g.append(gl)
        d_depth, d_pose, d_source = plan.synthesize_backward(source, intrinsic, depth_ms, pose, g,
                                                             want_source_grad=ctx.needs_input_grad[2])
        d_depth = [d.reshape(t.shape) for d, t in zip(d_depth, depth_ms)]
        return (None, None, d_source, None, d_pose, *d_depth)


class SynthesizeMultiScale:
    """reference model/synthesize/synthesize_base.py:10-29."""

    def __call__(self, source_image, intrinsic, pred_depth_ms, pred_pose, return_mask=False):
        """
        :param source_image: [batch, numsrc, height, width, 3]
        :param intrinsic: [batch, 3, 3]
        :param pred_depth_ms: list of [batch, height/scale, width/scale, 1]
        :param pred_pose: [batch, numsrc, 6] twist (t, rotation vector), target -> source
        :return: list of [batch, numsrc, height/scale, width/scale, 3]
                 (+ list of validity masks [batch, numsrc, h, w, 1] when return_mask)
        """
        source_image, intrinsic, pred_pose = as_torch(source_image), as_torch(intrinsic), as_torch(pred_pose)
        pred_depth_ms = [as_torch(d) for d in pred_depth_ms]
        if source_image.dim() != 5 or source_image.shape[-1] != 3:
            raise WrongInputException(f"source_image must be [batch, numsrc, height, width, 3], got {tuple(source_image.shape)}")
        require_cuda_f32(source_image=source_image, intrinsic=intrinsic, pred_pose=pred_pose, pred_depth_ms=pred_depth_ms)
        B, N, H, W, _ = source_image.shape
        if tuple(pred_pose.shape) != (B, N, 6):
            raise WrongInputException(f"pred_pose must be [{B}, {N}, 6], got {tuple(pred_pose.shape)}")
        plan = get_plan(source_image.device.index or 0, B, N, H, W, infer_scales(H, pred_depth_ms))
        outs = _SynthesizeFn.apply(plan, bool(return_mask), source_image, intrinsic, pred_pose, *pred_depth_ms)
        synth = list(outs[:plan.S])
        if return_mask:
            return synth, list(outs[plan.S:])
        return synth
