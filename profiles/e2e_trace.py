"""Device-side timeline of the host-buffer entry point (XPT_HOST_TRACE=1) + wall time per call.
Run on the GPU box: XPT_HOST_TRACE=1 python profiles/e2e_trace.py [cfg2|cfg3]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from xptwarp import _cabi
from oracle import xpt_oracle as orc
B, H, W = {"cfg2": (8, 128, 384), "cfg3": (16, 256, 832)}[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
N, S = 4, 4
feats, preds = orc.make_inputs(B, H, W, seed=3)
plan = xptwarp.get_plan(0, B, N, H, W, [1, 2, 4, 8], [1, 1, 1, 1], 0.5, 0.5, 1.0, B, _cabi.XPT_FLAG_GRAPH)
himg = feats["image5d"].contiguous().pin_memory()
hK, hpose = feats["intrinsic"].contiguous().pin_memory(), preds["pose"].contiguous().pin_memory()
hdepth = [d.contiguous().pin_memory() for d in preds["depth_ms"]]
hdisp = [d.contiguous().pin_memory() for d in preds["disp_ms"]]
hl, hdp = torch.zeros(4).pin_memory(), torch.zeros(B, N, 6).pin_memory()
hdd = [torch.zeros_like(d).pin_memory() for d in hdepth]; hds = [torch.zeros_like(d).pin_memory() for d in hdepth]
fr = _cabi.XptFrames()
fr.source, fr.source_batch_stride, fr.source_frame_stride = himg.data_ptr(), himg.stride(0), himg.stride(1)
fr.target, fr.target_batch_stride = himg.data_ptr() + N * himg.stride(1) * 4, himg.stride(0)
fr.intrinsic = hK.data_ptr()
o = _cabi.XptLossOutputs(); o.losses, o.d_pose, o.grad_scale = hl.data_ptr(), hdp.data_ptr(), 1.0
for s_ in range(S): o.d_depth_ms[s_], o.d_disp_ms[s_] = hdd[s_].data_ptr(), hds[s_].data_ptr()
dp, sp = _cabi.ptr_array([d.data_ptr() for d in hdepth]), _cabi.ptr_array([d.data_ptr() for d in hdisp])
st = torch.cuda.Stream(); torch.cuda.set_stream(st); stream = plan.stream()
def step():
    _cabi.check(plan._lib.xpt_total_loss_host(plan.handle, C.byref(fr), C.byref(dp), C.byref(sp), hpose.data_ptr(), C.byref(o), stream))
for _ in range(6): step()
t0 = time.perf_counter()
for _ in range(20): step()
print(f"wall per call: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
