"""Mirror of the reference's model/synthesize/flow_warping.py call surface."""
from __future__ import annotations

import torch

from .engine import WrongInputException, as_torch, get_plan, require_cuda_f32


def infer_flow_scales(H: int, flow_ms):
    """scale = H // H_s of each flow level, read from static shapes like flow_warping.py:44-45."""
    scales = []
    for f in flow_ms:
        hs = f.shape[2]
        if hs <= 0 or H % hs:
            raise WrongInputException(f"flow height {hs} does not divide the image height {H}")
        scales.append(H // hs)
    return scales


class _FlowWarpFn(torch.autograd.Function):
    """backward = xpt_flow_warp_backward (tape.gradient through flow_warping.py + bilinear_interp.py)."""

    @staticmethod
    def forward(ctx, plan, source, *flow_ms):
        warped, _ = plan.flow_warp(source, flow_ms)
        ctx.plan = plan
        ctx.save_for_backward(source, *flow_ms)
        return tuple(warped)

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        source, *flow_ms = ctx.saved_tensors
        g = []
        for l in range(plan.S):
            gl = grads[l]
            if gl is None:
                h, w = plan.level_hw(l)
                gl = torch.zeros((plan.B, plan.N, h, w, 3), dtype=torch.float32, device=plan.device)
            g.append(gl)
        want_flow = any(ctx.needs_input_grad[2:])
        d_flow, d_source = plan.flow_warp_backward(source, flow_ms, g, want_flow_grad=want_flow,
                                                   want_source_grad=ctx.needs_input_grad[1])
        d_flow = [d.reshape(t.shape) for d, t in zip(d_flow, flow_ms)] if want_flow else [None] * plan.S
        return (None, d_source, *d_flow)


class FlowWarpMultiScale:
    """reference model/synthesize/flow_warping.py:11-71."""

    def __call__(self, source_image, flow_ms):
        """
        :param source_image: source images [batch, numsrc, height, width, 3]
        :param flow_ms: optical flow from source to target in multi scale,
                        list of [batch, numsrc, height/scale, width/scale, 2]
        :return: reconstructed target view in multi scale, list of [batch, numsrc, height/scale, width/scale, 3]
        """
        source_image = as_torch(source_image)
        flow_ms = [as_torch(f) for f in flow_ms]
        if source_image.dim() != 5 or source_image.shape[-1] != 3:
            raise WrongInputException(f"source_image must be [batch, numsrc, height, width, 3], got {tuple(source_image.shape)}")
        require_cuda_f32(source_image=source_image, flow_ms=flow_ms)
        B, N, H, W, _ = source_image.shape
        for l, f in enumerate(flow_ms):
            if f.dim() != 5 or f.shape[0] != B or f.shape[1] != N or f.shape[4] != 2:
                raise WrongInputException(f"flow_ms[{l}] must be [{B}, {N}, h, w, 2], got {tuple(f.shape)}")
        plan = get_plan(source_image.device.index or 0, B, N, H, W, infer_flow_scales(H, flow_ms))
        return list(_FlowWarpFn.apply(plan, source_image, *flow_ms))
