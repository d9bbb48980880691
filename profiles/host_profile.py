"""cProfile of the public-API step (TotalLoss forward + backward) on a stereo rig: where the HOST time goes.
Run on the GPU box: python profiles/host_profile.py [LOSS_RIGID_T1|mono]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import numpy as np, torch, xptwarp
from oracle import xpt_oracle as orc
B, H, W = 8, 128, 384
mono = len(sys.argv) > 1 and sys.argv[1] == "mono"
if mono:
    feats, preds = orc.make_inputs(B, H, W, seed=3)
    lw = {"L1": .5, "SSIM": .5, "smoothe": 1.}
    cfg = {"image": 1, "intrinsic": 1}
else:
    feats, preds = orc.make_stereo_inputs(B, H, W, seed=3)
    lw = {"L1": .5, "L1_R": .5, "SSIM": .5, "SSIM_R": .5, "smoothe": 1., "smoothe_R": 1., "stereoL1": .01, "stereoSSIM": .01, "stereoPose": 1.}
    cfg = {"image": 1, "intrinsic": 1, "image_R": 1, "intrinsic_R": 1, "stereo_T_LR": 1}
f = {k: v.cuda() for k, v in feats.items()}
p = {k: ([t.cuda().requires_grad_(True) for t in v] if isinstance(v, list) else v.cuda().requires_grad_(True)) for k, v in preds.items()}
tot = xptwarp.loss_factory(cfg, lw, np.array([1., 1., 1., 1.]), stereo=not mono, batch_size=B)
def step():
    for v in p.values():
        for t in (v if isinstance(v, list) else [v]):
            t.grad = None
    total, _ = tot(p, f)
    total.backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print(f"{(time.perf_counter()-t0)/50*1e3:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
