"""Device time of the min-over-sources / combined loss kernel (k_photo_min) at config-2 frame size.
Run on the GPU box: python profiles/minloss_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from oracle import xpt_oracle as orc
B, H, W, N = 8, 128, 384, 4
feats, preds = orc.make_inputs(B, H, W, N=N, seed=3)
src, tgt = feats["image5d"][:, :-1].cuda(), feats["image5d"][:, -1].cuda()
synth = xptwarp.SynthesizeMultiScale()(src, feats["intrinsic"].cuda(), [d.cuda() for d in preds["depth_ms"]], preds["pose"].cuda())
# the stereo view: a DIFFERENT image (source 0's synthesis shifted by three pixels) -- a copy of source 0 would tie with it
# at every pixel, which no real rig produces
stereo = [torch.roll(s[:, :1], shifts=3, dims=3).contiguous() for s in synth]
warped0 = synth[2].clone()
import sys as _s
from xptwarp import _cabi
plan = xptwarp.get_plan(0, B, N, H, W, [1, 2, 4, 8], [1, 1, 1, 1], flags=_cabi.XPT_FLAG_MIN_TILES if "tiles" in _s.argv else 0)
print("kernel:", "k_photo_min (tiles)" if "tiles" in _s.argv else "k_min_strip")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for method, mname in ((0, "L1"), (2, "SSIM")):
    for grad in (False, True):
        t_md2 = timed(lambda: plan.photometric_min_loss(method, synth, None, tgt, want_grad=grad))
        t_moa = timed(lambda: plan.photometric_min_loss(method, synth, stereo, tgt, want_grad=grad))
        t_cmb = timed(lambda: plan.photometric_cmb_loss(method, synth, warped0, tgt, want_grad=grad))
        print(f"{mname:4s} grad={int(grad)}  md2 {t_md2:7.1f} us   moa {t_moa:7.1f} us   cmb {t_cmb:7.1f} us")
for grad in (() if "tiles" in _s.argv else (False, True)):
    t_md2 = timed(lambda: plan.photometric_min_pair_loss(synth, None, tgt, 1.0, 1.0, want_grad=grad))
    t_moa = timed(lambda: plan.photometric_min_pair_loss(synth, stereo, tgt, 1.0, 1.0, want_grad=grad))
    print(f"PAIR grad={int(grad)}  md2 {t_md2:7.1f} us   moa {t_moa:7.1f} us   (L1 + SSIM in one launch)")
