"""Lean A/B timing of the fused kernel: 'full fwd+bwd' only, kernel time (xpt_profile_*) and step time.
Run on the GPU box:  XPTWARP_LIB=... python profiles/ab_time.py cfg2 [cfg3 ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from oracle import xpt_oracle as orc

SH = {"cfg2": (8, 128, 384), "cfg3": (16, 256, 832), "cfg4": (64, 128, 384), "cfg5": (128, 384, 1280)}
for name in sys.argv[1:] or ["cfg2"]:
    B, H, W = SH[name]
    feats, preds = orc.make_inputs(B, H, W, seed=5)
    f = {k: v.cuda() for k, v in feats.items()}
    p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]], "pose": preds["pose"].cuda()}
    img = f["image5d"]
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    plan = xptwarp.get_plan(0, B, 4, H, W, [1, 2, 4, 8], [1, 1, 1, 1], 0.5, 0.5, 1.0, B)
    call = plan.bind_total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"],
                                want_grad=True, want_synth=False, want_mask=False, want_source_grad=False)
    big = B * H * W > 2e7
    n = 100 if B * H * W < 1e6 else (10 if big else 40)
    for _ in range(2000 if B * H * W < 1e6 else (20 if big else 300)): call.run()
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        plan.profile_begin(n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): call.run()
        e1.record(); torch.cuda.synchronize()
        d = sorted(plan.profile_end(n))
        res.append(f"k p50 {1e3*d[len(d)//2]:.1f} mean {1e3*sum(d)/len(d):.1f} step {1e3*e0.elapsed_time(e1)/n:.1f}")
    print(name, os.path.basename(os.environ.get("XPTWARP_LIB", "tree")), " | ".join(res), "us", flush=True)
