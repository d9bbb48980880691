"""One configuration of k_photo_min in a loop (for ncu): python profiles/minloss_one.py [L1|SSIM] [md2|moa|cmb]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from oracle import xpt_oracle as orc
B, H, W, N = 8, 128, 384, 4
method = {"L1": 0, "SSIM": 2, "PAIR": 3}[sys.argv[1] if len(sys.argv) > 1 else "SSIM"]
kind = sys.argv[2] if len(sys.argv) > 2 else "moa"
feats, preds = orc.make_inputs(B, H, W, N=N, seed=3)
src, tgt = feats["image5d"][:, :-1].cuda(), feats["image5d"][:, -1].cuda()
synth = xptwarp.SynthesizeMultiScale()(src, feats["intrinsic"].cuda(), [d.cuda() for d in preds["depth_ms"]], preds["pose"].cuda())
stereo = [s[:, :1].contiguous() for s in synth]
plan = xptwarp.get_plan(0, B, N, H, W, [1, 2, 4, 8], [1, 1, 1, 1])
for _ in range(8):
    if kind == "cmb":
        plan.photometric_cmb_loss(method, synth, synth[2].clone(), tgt, want_grad=True)
    elif method == 3:
        plan.photometric_min_pair_loss(synth, stereo if kind == "moa" else None, tgt, 1.0, 1.0, want_grad=True)
    else:
        plan.photometric_min_loss(method, synth, stereo if kind == "moa" else None, tgt, want_grad=True)
torch.cuda.synchronize()
