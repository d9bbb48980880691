"""Host-side engine: plans (xpt_ctx handles), pointer marshalling, output allocation.

PyTorch is used here for device memory, streams and autograd plumbing only; all
arithmetic happens in libxptwarp.so.  Inputs may be torch CUDA tensors or any
``__dlpack__`` producer (zero-copy, see dlpack.py)."""
from __future__ import annotations

import collections
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import XPT_MAX_SCALES, XptConfig, XptFrames, XptLossOutputs, ptr_array
from .dlpack import kDLCUDA, view_of


class WrongInputException(Exception):
    """Same name and role as the reference's utils/util_class.py:11-13."""

    def __init__(self, msg):
        super().__init__(msg)


def as_torch(x) -> torch.Tensor:
    """torch tensor as-is; any other DLPack producer is imported zero-copy."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule":
        v = view_of(x)            # validates dtype and reads the device without consuming the capsule
        if v.device_type != kDLCUDA:
            raise WrongInputException("xptwarp needs CUDA tensors (DLPack device_type kDLCUDA); there is no CPU path")
        return torch.from_dlpack(x)
    raise WrongInputException(f"unsupported tensor type {type(x).__name__}: need torch.Tensor or __dlpack__")


def _check_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise WrongInputException(f"{name} must live on a CUDA device: xptwarp has no CPU fallback")
    if t.dtype != torch.float32:
        raise WrongInputException(f"{name} must be float32, got {t.dtype}")


def require_cuda_f32(**named):
    """Entry-point guard: every tensor must already live on a CUDA device as float32."""
    for name, t in named.items():
        for i, x in enumerate(t if isinstance(t, (list, tuple)) else [t]):
            _check_cuda_f32(x, name if not isinstance(t, (list, tuple)) else f"{name}[{i}]")


def _dense(t: torch.Tensor, name: str) -> torch.Tensor:
    """The C-ABI takes dense tensors and the library never copies behind the caller's back (a silent
    .contiguous() would be a hidden read + write of the tensor through a PyTorch kernel on the hot path)."""
    _check_cuda_f32(t, name)
    if not t.is_contiguous():
        raise WrongInputException(f"{name} must be contiguous (shape {tuple(t.shape)}, strides {t.stride()}): "
                                  "xptwarp does not copy its inputs; call .contiguous() where the tensor is produced")
    return t


def scale_tensors(tensors, scale: torch.Tensor):
    """[t * scale for t in tensors] in ONE launch (xpt_scale_tensors); `scale` is a device scalar"""
    ts = [t if t.is_contiguous() else t.contiguous() for t in tensors]
    outs = [torch.empty_like(t) for t in ts]
    n = len(ts)
    if n == 0:
        return outs
    dev = ts[0].device
    sc = scale.reshape(-1)[:1].to(device=dev, dtype=torch.float32).contiguous()
    src = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    dst = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
    cnt = (C.c_int64 * n)(*[t.numel() for t in ts])
    _cabi.check(_cabi.lib().xpt_scale_tensors(dev.index or 0, src, dst, cnt, n, sc.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return outs


def _frame_view(t: torch.Tensor, name: str, lead: int) -> torch.Tensor:
    """[...lead dims..., H, W, 3] with dense inner (H, W, 3): leading dims may be strided views."""
    _check_cuda_f32(t, name)
    H, W, Cn = t.shape[-3:]
    if Cn != 3:
        raise WrongInputException(f"{name}: channel-last RGB expected, got {tuple(t.shape)}")
    if t.stride(-1) != 1 or t.stride(-2) != 3 or t.stride(-3) != W * 3:
        raise WrongInputException(f"{name}: the inner (H, W, 3) dimensions must be dense, got strides {t.stride()}: "
                                  "xptwarp does not copy its inputs (leading batch / frame dimensions may be strided views)")
    return t


class Plan:
    """Owns one xpt_ctx (shapes, weights and device scratch bound at creation)."""

    def __init__(self, device_index, B, N, H, W, scales, scale_weights, w_l1, w_ssim, w_smooth,
                 global_batch, flags=0, img_grad_factor=4.0):
        if len(scales) > XPT_MAX_SCALES:
            raise WrongInputException(f"at most {XPT_MAX_SCALES} scales")
        cfg = XptConfig()
        cfg.batch, cfg.num_src, cfg.height, cfg.width = B, N, H, W
        cfg.num_scales = len(scales)
        for i, s in enumerate(scales):
            cfg.scales[i] = int(s)
            cfg.scale_weights[i] = float(scale_weights[i])
        cfg.w_l1, cfg.w_ssim, cfg.w_smooth = float(w_l1), float(w_ssim), float(w_smooth)
        cfg.img_grad_factor = float(img_grad_factor)
        cfg.global_batch = int(global_batch)
        cfg.device = int(device_index)
        cfg.flags = int(flags)
        self.cfg = cfg
        self.B, self.N, self.H, self.W = B, N, H, W
        self.scales = tuple(int(s) for s in scales)
        self.S = len(scales)
        self.device = torch.device("cuda", device_index)
        self._lib = _cabi.lib()
        h = C.c_void_p()
        _cabi.check(self._lib.xpt_create(C.byref(h), C.byref(cfg)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self._lib.xpt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------
    def level_hw(self, l):
        return self.H // self.scales[l], self.W // self.scales[l]

    def stream(self):
        if not getattr(self, "pinned", False) and torch.cuda.is_current_stream_capturing():
            self.pinned = True        # a capture now holds pointers into this plan's scratch: keep the ctx alive
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launches(self):
        return self._lib.xpt_last_launch_count(self.handle)

    def scratch_bytes(self):
        return self._lib.xpt_scratch_bytes(self.handle)

    def geometry_slot_shared(self):
        return bool(self._lib.xpt_geometry_slot_shared(self.handle))

    # ---- multi-rank (reference distributer.py:93-110) ---------------------------------------
    def comm_init(self, dist=None):
        """Create this plan's NCCL communicator: rank 0 makes the rendezvous id, torch.distributed (any backend)
        carries its 128 bytes to the other ranks -- plumbing only, the collectives themselves are the library's."""
        import torch.distributed as td
        dist = dist or td
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _cabi.check(self._lib.xpt_comm_unique_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8)
        backend = dist.get_backend()
        if backend == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        _cabi.check(self._lib.xpt_comm_init(self.handle, buf, world, rank))

    def comm_destroy(self):
        """Release the communicator and peer mappings (call before torch.distributed.destroy_process_group)."""
        _cabi.check(self._lib.xpt_comm_destroy(self.handle))

    def comm_status(self):
        """(uses_peer_memory, exchange_error) of this plan's multi-rank loss exchange; synchronising."""
        p2p, err = C.c_int(0), C.c_int(0)
        _cabi.check(self._lib.xpt_comm_status(self.handle, C.byref(p2p), C.byref(err)))
        return bool(p2p.value), bool(err.value)

    def allreduce(self, tensors):
        """in-place sum over the ranks of dense fp32 device tensors, one NCCL group on the current stream"""
        ts = [_dense(t, f"tensors[{i}]") for i, t in enumerate(tensors)]
        n = len(ts)
        ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in ts])
        cnt = (C.c_int64 * max(n, 1))(*[t.numel() for t in ts])
        _cabi.check(self._lib.xpt_allreduce(self.handle, ptrs, cnt, n, self.stream()))
        return tensors

    def _frames(self, source, target, intrinsic):
        f = XptFrames()
        f.source = source.data_ptr()
        f.source_batch_stride = source.stride(0)
        f.source_frame_stride = source.stride(1)
        if target is not None:
            f.target = target.data_ptr()
            f.target_batch_stride = target.stride(0)
        f.intrinsic = intrinsic.data_ptr()
        return f

    def _level_list(self, ts: Sequence[torch.Tensor], name, last):
        if len(ts) != self.S:
            raise WrongInputException(f"{name}: expected {self.S} scales, got {len(ts)}")
        out = []
        for l, t in enumerate(ts):
            h, w = self.level_hw(l)
            t = _dense(t, f"{name}[{l}]")
            if t.numel() != self.B * h * w * last:
                raise WrongInputException(f"{name}[{l}]: expected {self.B}x..x{h}x{w} ({last} per pixel), got {tuple(t.shape)}")
            out.append(t)
        return out

    def _empty_levels(self, *lead_last):
        lead, last = lead_last
        return [torch.empty((self.B, *lead, *self.level_hw(l), last), dtype=torch.float32, device=self.device)
                for l in range(self.S)]

    # ---- C-ABI calls ----------------------------------------------------------
    def pose_rvec2matr(self, pose):
        pose = _dense(pose, "pose")
        out = torch.empty((self.B, self.N, 4, 4), dtype=torch.float32, device=self.device)
        _cabi.check(self._lib.xpt_pose_rvec2matr(self.handle, pose.data_ptr(), out.data_ptr(), self.stream()))
        return out

    def build_pyramids(self, source, target, intrinsic=None):
        source = _frame_view(source, "source", 2)
        target = _frame_view(target, "target", 1)
        K = intrinsic if intrinsic is not None else torch.zeros(self.B, 3, 3, device=self.device)
        tgt = self._empty_levels((), 3)
        f = self._frames(source, target, _dense(K, "intrinsic"))
        arr = ptr_array([t.data_ptr() for t in tgt])
        _cabi.check(self._lib.xpt_build_pyramids(self.handle, C.byref(f), C.byref(arr), self.stream()))
        return tgt

    def synthesize(self, source, intrinsic, depth_ms, pose, want_mask=False):
        source = _frame_view(source, "source", 2)
        intrinsic, pose = _dense(intrinsic, "intrinsic"), _dense(pose, "pose")
        depth_ms = self._level_list(depth_ms, "depth_ms", 1)
        synth = self._empty_levels((self.N,), 3)
        mask = self._empty_levels((self.N,), 1) if want_mask else None
        f = self._frames(source, None, intrinsic)
        d = ptr_array([t.data_ptr() for t in depth_ms])
        s = ptr_array([t.data_ptr() for t in synth])
        m = ptr_array([t.data_ptr() for t in mask]) if want_mask else None
        _cabi.check(self._lib.xpt_synthesize(self.handle, C.byref(f), C.byref(d), pose.data_ptr(), C.byref(s),
                                             C.byref(m) if want_mask else None, self.stream()))
        return synth, mask

    def synthesize_backward(self, source, intrinsic, depth_ms, pose, grad_synth_ms, want_source_grad=False):
        source = _frame_view(source, "source", 2)
        intrinsic, pose = _dense(intrinsic, "intrinsic"), _dense(pose, "pose")
        depth_ms = self._level_list(depth_ms, "depth_ms", 1)
        grad_synth_ms = self._level_list(grad_synth_ms, "grad_synth_ms", self.N * 3)
        d_depth = self._empty_levels((), 1)
        d_pose = torch.empty((self.B, self.N, 6), dtype=torch.float32, device=self.device)
        d_source = (torch.empty((self.B, self.N, self.H, self.W, 3), dtype=torch.float32, device=self.device)
                    if want_source_grad else None)
        f = self._frames(source, None, intrinsic)
        _cabi.check(self._lib.xpt_synthesize_backward(
            self.handle, C.byref(f), C.byref(ptr_array([t.data_ptr() for t in depth_ms])), pose.data_ptr(),
            C.byref(ptr_array([t.data_ptr() for t in grad_synth_ms])),
            C.byref(ptr_array([t.data_ptr() for t in d_depth])), d_pose.data_ptr(),
            d_source.data_ptr() if want_source_grad else None, self.stream()))
        return d_depth, d_pose, d_source

    def photometric_loss(self, method, synth_ms, target_ms, grad_loss_batch=None, want_grad=False):
        synth_ms = self._level_list(synth_ms, "synth_target_ms", self.N * 3)
        target_ms = self._level_list(target_ms, "target_ms", 3)
        loss = torch.empty((self.B,), dtype=torch.float32, device=self.device)
        d_synth = self._empty_levels((self.N,), 3) if want_grad else None
        g = _dense(grad_loss_batch, "grad_loss_batch") if grad_loss_batch is not None else None
        _cabi.check(self._lib.xpt_photometric_loss(
            self.handle, int(method), C.byref(ptr_array([t.data_ptr() for t in synth_ms])),
            C.byref(ptr_array([t.data_ptr() for t in target_ms])), loss.data_ptr(),
            g.data_ptr() if g is not None else None,
            C.byref(ptr_array([t.data_ptr() for t in d_synth])) if want_grad else None, self.stream()))
        return loss, d_synth

    def photometric_min_loss(self, method, synth_ms, stereo_synth_ms, target, grad_loss_batch=None, want_grad=False):
        """xpt_photometric_min_loss: per-pixel min over the N (+1 stereo) sources at full resolution."""
        synth_ms = self._level_list(synth_ms, "synth_target_ms", self.N * 3)
        have_st = stereo_synth_ms is not None
        if have_st:
            stereo_synth_ms = self._level_list(stereo_synth_ms, "stereo_synth_ms", 3)
        target = _frame_view(target, "target", 1)
        loss = torch.empty((self.B,), dtype=torch.float32, device=self.device)
        d_synth = self._empty_levels((self.N,), 3) if want_grad else None
        d_stereo = self._empty_levels((1,), 3) if (want_grad and have_st) else None
        g = _dense(grad_loss_batch, "grad_loss_batch") if grad_loss_batch is not None else None
        _cabi.check(self._lib.xpt_photometric_min_loss(
            self.handle, int(method), C.byref(ptr_array([t.data_ptr() for t in synth_ms])),
            C.byref(ptr_array([t.data_ptr() for t in stereo_synth_ms])) if have_st else None,
            target.data_ptr(), target.stride(0), loss.data_ptr(), g.data_ptr() if g is not None else None,
            C.byref(ptr_array([t.data_ptr() for t in d_synth])) if want_grad else None,
            C.byref(ptr_array([t.data_ptr() for t in d_stereo])) if d_stereo is not None else None, self.stream()))
        return loss, d_synth, d_stereo, (target, synth_ms, stereo_synth_ms)

    def photometric_min_pair_loss(self, synth_ms, stereo_synth_ms, target, grad_l1=1.0, grad_ssim=1.0, want_grad=False):
        """xpt_photometric_min_pair_loss: the L1 and the SSIM min-over-sources loss of one loss set in ONE launch;
        the gradient is grad_l1 * dL1 + grad_ssim * dSSIM (the two losses' upstream weights)."""
        synth_ms = self._level_list(synth_ms, "synth_target_ms", self.N * 3)
        have_st = stereo_synth_ms is not None
        if have_st:
            stereo_synth_ms = self._level_list(stereo_synth_ms, "stereo_synth_ms", 3)
        target = _frame_view(target, "target", 1)
        loss = torch.empty((2, self.B), dtype=torch.float32, device=self.device)
        d_synth = self._empty_levels((self.N,), 3) if want_grad else None
        d_stereo = self._empty_levels((1,), 3) if (want_grad and have_st) else None
        _cabi.check(self._lib.xpt_photometric_min_pair_loss(
            self.handle, C.byref(ptr_array([t.data_ptr() for t in synth_ms])),
            C.byref(ptr_array([t.data_ptr() for t in stereo_synth_ms])) if have_st else None,
            target.data_ptr(), target.stride(0), loss[0].data_ptr(), loss[1].data_ptr(), float(grad_l1), float(grad_ssim),
            C.byref(ptr_array([t.data_ptr() for t in d_synth])) if want_grad else None,
            C.byref(ptr_array([t.data_ptr() for t in d_stereo])) if d_stereo is not None else None, self.stream()))
        return loss, d_synth, d_stereo

    def photometric_cmb_loss(self, method, synth_ms, warped, target, grad_loss_batch=None, want_grad=False):
        """xpt_photometric_cmb_loss: static term where it beats the optical-flow term (CombinedLossMultiScale)."""
        synth_ms = self._level_list(synth_ms, "synth_target_ms", self.N * 3)
        warped = _dense(warped, "warped_target_ms[0]")
        if warped.dim() != 5 or warped.shape[0] != self.B or warped.shape[1] != self.N or warped.shape[4] != 3:
            raise WrongInputException(f"warped_target_ms[0]: expected [{self.B},{self.N},h,w,3], got {tuple(warped.shape)}")
        target = _frame_view(target, "target", 1)
        loss = torch.empty((self.B,), dtype=torch.float32, device=self.device)
        d_synth = self._empty_levels((self.N,), 3) if want_grad else None
        g = _dense(grad_loss_batch, "grad_loss_batch") if grad_loss_batch is not None else None
        _cabi.check(self._lib.xpt_photometric_cmb_loss(
            self.handle, int(method), C.byref(ptr_array([t.data_ptr() for t in synth_ms])), warped.data_ptr(),
            int(warped.shape[2]), int(warped.shape[3]), target.data_ptr(), target.stride(0), loss.data_ptr(),
            g.data_ptr() if g is not None else None,
            C.byref(ptr_array([t.data_ptr() for t in d_synth])) if want_grad else None, self.stream()))
        return loss, d_synth, (target, synth_ms, warped)

    def photometric_cmb_pair_loss(self, synth_ms, warped, target, grad_l1=1.0, grad_ssim=1.0, want_grad=False):
        """xpt_photometric_cmb_pair_loss: cmbL1 + cmbSSIM of one eye in ONE launch."""
        synth_ms = self._level_list(synth_ms, "synth_target_ms", self.N * 3)
        warped = _dense(warped, "warped_target_ms[0]")
        if warped.dim() != 5 or warped.shape[0] != self.B or warped.shape[1] != self.N or warped.shape[4] != 3:
            raise WrongInputException(f"warped_target_ms[0]: expected [{self.B},{self.N},h,w,3], got {tuple(warped.shape)}")
        target = _frame_view(target, "target", 1)
        loss = torch.empty((2, self.B), dtype=torch.float32, device=self.device)
        d_synth = self._empty_levels((self.N,), 3) if want_grad else None
        _cabi.check(self._lib.xpt_photometric_cmb_pair_loss(
            self.handle, C.byref(ptr_array([t.data_ptr() for t in synth_ms])), warped.data_ptr(),
            int(warped.shape[2]), int(warped.shape[3]), target.data_ptr(), target.stride(0), loss[0].data_ptr(),
            loss[1].data_ptr(), float(grad_l1), float(grad_ssim),
            C.byref(ptr_array([t.data_ptr() for t in d_synth])) if want_grad else None, self.stream()))
        return loss, d_synth

    def flow_warp(self, source, flow_ms, want_mask=False):
        """xpt_flow_warp (FlowWarpMultiScale.__call__); the plan's scales are the flow scales."""
        source = _frame_view(source, "source", 2)
        flow_ms = self._level_list(flow_ms, "flow_ms", self.N * 2)
        warped = self._empty_levels((self.N,), 3)
        mask = self._empty_levels((self.N,), 1) if want_mask else None
        f = XptFrames()
        f.source, f.source_batch_stride, f.source_frame_stride = source.data_ptr(), source.stride(0), source.stride(1)
        _cabi.check(self._lib.xpt_flow_warp(
            self.handle, C.byref(f), C.byref(ptr_array([t.data_ptr() for t in flow_ms])),
            C.byref(ptr_array([t.data_ptr() for t in warped])),
            C.byref(ptr_array([t.data_ptr() for t in mask])) if want_mask else None, self.stream()))
        return warped, mask

    def flow_warp_backward(self, source, flow_ms, grad_warped_ms, want_flow_grad=True, want_source_grad=False):
        source = _frame_view(source, "source", 2)
        flow_ms = self._level_list(flow_ms, "flow_ms", self.N * 2)
        grad_warped_ms = self._level_list(grad_warped_ms, "grad_warped_ms", self.N * 3)
        d_flow = self._empty_levels((self.N,), 2) if want_flow_grad else None
        d_source = (torch.empty((self.B, self.N, self.H, self.W, 3), dtype=torch.float32, device=self.device)
                    if want_source_grad else None)
        f = XptFrames()
        f.source, f.source_batch_stride, f.source_frame_stride = source.data_ptr(), source.stride(0), source.stride(1)
        _cabi.check(self._lib.xpt_flow_warp_backward(
            self.handle, C.byref(f), C.byref(ptr_array([t.data_ptr() for t in flow_ms])),
            C.byref(ptr_array([t.data_ptr() for t in grad_warped_ms])),
            C.byref(ptr_array([t.data_ptr() for t in d_flow])) if want_flow_grad else None,
            d_source.data_ptr() if want_source_grad else None, self.stream()))
        return d_flow, d_source

    def l2_regularizer(self, weights, grad_loss=None, want_grad=False):
        """xpt_l2_regularizer: sum_i sum(w_i^2)/2 -> [1]; with want_grad also grad_loss * w_i per tensor."""
        ws = [_dense(w, f"weights[{i}]") for i, w in enumerate(weights)]
        n = len(ws)
        wp = (C.c_void_p * max(n, 1))(*[w.data_ptr() for w in ws])
        cnt = (C.c_int64 * max(n, 1))(*[w.numel() for w in ws])
        loss = torch.empty((1,), dtype=torch.float32, device=self.device)
        d_w = [torch.empty_like(w) for w in ws] if want_grad else None
        dp = (C.c_void_p * max(n, 1))(*[d.data_ptr() for d in d_w]) if want_grad else None
        g = _dense(grad_loss, "grad_loss") if grad_loss is not None else None
        _cabi.check(self._lib.xpt_l2_regularizer(self.handle, wp, cnt, n, loss.data_ptr(),
                                                 g.data_ptr() if g is not None else None, dp, self.stream()))
        return loss, d_w

    def smoothness_loss(self, disp_ms, target_ms, grad_loss_batch=None, want_grad=False):
        disp_ms = self._level_list(disp_ms, "disp_ms", 1)
        target_ms = self._level_list(target_ms, "target_ms", 3)
        loss = torch.empty((self.B,), dtype=torch.float32, device=self.device)
        d_disp = self._empty_levels((), 1) if want_grad else None
        g = _dense(grad_loss_batch, "grad_loss_batch") if grad_loss_batch is not None else None
        _cabi.check(self._lib.xpt_smoothness_loss(
            self.handle, C.byref(ptr_array([t.data_ptr() for t in disp_ms])),
            C.byref(ptr_array([t.data_ptr() for t in target_ms])), loss.data_ptr(),
            g.data_ptr() if g is not None else None,
            C.byref(ptr_array([t.data_ptr() for t in d_disp])) if want_grad else None, self.stream()))
        return loss, d_disp

    def bind_total_loss(self, source, target, intrinsic, depth_ms, disp_ms, pose, want_grad=True,
                        want_synth=False, want_mask=False, want_target_ms=False, want_source_grad=False,
                        want_loss_batch=False, grad_scale=1.0, out: Optional[dict] = None) -> "BoundTotalLoss":
        """Marshal one xpt_total_loss call once; .run() then only crosses the C-ABI.  `out` may carry
        preallocated output tensors (same keys as the result dict) to be reused."""
        source = _frame_view(source, "source", 2)
        target = _frame_view(target, "target", 1)
        intrinsic, pose = _dense(intrinsic, "intrinsic"), _dense(pose, "pose")
        depth_ms = self._level_list(depth_ms, "depth_ms", 1)
        have_disp = disp_ms is not None
        if have_disp:
            disp_ms = self._level_list(disp_ms, "disp_ms", 1)
        # disp_ms None with a smoothe weight: the kernel derives disp = safe_reciprocal_number(depth)
        # (utils/util_funcs.py:146-160) and folds d_disp into d_depth
        r = out if out is not None else {}
        dev, f32 = self.device, torch.float32

        def need(key, make):
            if key not in r or r[key] is None:
                r[key] = make()
            return r[key]
        o = XptLossOutputs()
        o.grad_scale = float(grad_scale)
        o.losses = need("losses", lambda: torch.empty(4, dtype=f32, device=dev)).data_ptr()
        if want_loss_batch:
            o.loss_batch = need("loss_batch", lambda: torch.empty((3, self.B), dtype=f32, device=dev)).data_ptr()

        def put(field, key, lead, last):
            ts = need(key, lambda: self._empty_levels(lead, last))
            for l, t in enumerate(ts):
                getattr(o, field)[l] = t.data_ptr()
        if want_synth:
            put("synth_ms", "synth_ms", (self.N,), 3)
        if want_mask:
            put("mask_ms", "mask_ms", (self.N,), 1)
        if want_target_ms:
            put("target_ms", "target_ms", (), 3)
        if want_grad:
            put("d_depth_ms", "d_depth_ms", (), 1)
            if have_disp or self.cfg.w_smooth == 0.0:
                put("d_disp_ms", "d_disp_ms", (), 1)
            o.d_pose = need("d_pose", lambda: torch.empty((self.B, self.N, 6), dtype=f32, device=dev)).data_ptr()
            if want_source_grad:
                o.d_source = need("d_source", lambda: torch.empty((self.B, self.N, self.H, self.W, 3), dtype=f32,
                                                                  device=dev)).data_ptr()
        f = self._frames(source, target, intrinsic)
        d = ptr_array([t.data_ptr() for t in depth_ms])
        dd = ptr_array([t.data_ptr() for t in disp_ms]) if have_disp else None
        keep = (source, target, intrinsic, pose, depth_ms, disp_ms)
        return BoundTotalLoss(self, f, d, dd, pose.data_ptr(), o, r, keep)

    def total_loss(self, *args, **kwargs):
        """xpt_total_loss (see bind_total_loss for the arguments); returns the dict of outputs."""
        return self.bind_total_loss(*args, **kwargs).run()

    def profile_begin(self, max_records, kernel=0):
        """kernel: 0 = the fused tile kernel, 1 = the tiled pyramid kernel (XPT_PROFILE_*)"""
        _cabi.check(self._lib.xpt_profile_select(self.handle, int(kernel)))
        _cabi.check(self._lib.xpt_profile_begin(self.handle, int(max_records)))

    def profile_end(self, capacity):
        buf = (C.c_float * capacity)()
        n = self._lib.xpt_profile_end(self.handle, buf, capacity)
        if n < 0:
            _cabi.check(n)
        return [buf[i] for i in range(n)]


class BoundTotalLoss:
    """A fully marshalled xpt_total_loss call (ctypes structs prebuilt, tensors kept alive)."""
    __slots__ = ("plan", "_f", "_d", "_dd", "_pose", "_o", "out", "_keep", "_fn", "_h")

    def __init__(self, plan, f, d, dd, pose_ptr, o, out, keep):
        self.plan, self._f, self._d, self._dd, self._pose, self._o = plan, f, d, dd, pose_ptr, o
        self.out, self._keep = out, keep
        self._fn, self._h = plan._lib.xpt_total_loss, plan.handle

    def run(self, stream=None):
        st = self.plan.stream() if stream is None else stream
        rc = self._fn(self._h, C.byref(self._f), C.byref(self._d), C.byref(self._dd) if self._dd is not None else None,
                      self._pose, C.byref(self._o), st)
        if rc != 0:
            _cabi.check(rc)
        return self.out


_PLANS: "collections.OrderedDict[tuple, Plan]" = collections.OrderedDict()
# Large enough that one rig + flow loss set (about 8 plans per step) and a validation batch never evict a live plan;
# a plan that a CUDA-graph capture has used is PINNED: the captured graph keeps raw pointers into the plan's scratch
# (source pyramid, partial sums), so its ctx must outlive every replay.
_MAX_PLANS = 64


def get_plan(device_index, B, N, H, W, scales, scale_weights=None, w_l1=0.0, w_ssim=0.0, w_smooth=0.0,
             global_batch=0, flags=0, img_grad_factor=4.0) -> Plan:
    scales = tuple(int(s) for s in scales)
    sw = tuple(float(w) for w in (scale_weights if scale_weights is not None else [1.0] * len(scales)))
    key = (device_index, B, N, H, W, scales, sw, float(w_l1), float(w_ssim), float(w_smooth),
           int(global_batch) or B, int(flags), float(img_grad_factor))
    p = _PLANS.get(key)
    if p is None:
        p = Plan(device_index, B, N, H, W, scales, sw, w_l1, w_ssim, w_smooth, int(global_batch) or B, flags,
                 img_grad_factor)
        _PLANS[key] = p
        if len(_PLANS) > _MAX_PLANS:
            # only drop the cache's reference: an autograd graph may still hold the plan for its backward;
            # Plan.__del__ destroys the xpt_ctx when the last reference goes.  Pinned plans are never dropped.
            for k in [k for k, v in _PLANS.items() if not getattr(v, "pinned", False)][:len(_PLANS) - _MAX_PLANS]:
                del _PLANS[k]
    else:
        _PLANS.move_to_end(key)
    return p


def infer_scales(H: int, depth_ms: Sequence[torch.Tensor]) -> List[int]:
    """scale = H // H_s, read from static shapes like synthesize_base.py:61-64."""
    scales = []
    for d in depth_ms:
        hs = d.shape[1]
        if hs <= 0 or H % hs:
            raise WrongInputException(f"depth height {hs} does not divide the image height {H}")
        scales.append(H // hs)
    return scales
