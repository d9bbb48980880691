"""Ablation timing of the fused kernel (CUDA events via xpt_profile_*): which term costs what.
Run on the GPU box:  python profiles/phase_split.py [cfg2|cfg3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from oracle import xpt_oracle as orc

wl = {"cfg2": (8, 128, 384), "cfg3": (16, 256, 832)}[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B, H, W = wl
feats, preds = orc.make_inputs(B, H, W, seed=5)
f = {k: v.cuda() for k, v in feats.items()}
p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]], "pose": preds["pose"].cuda()}
img = f["image5d"]
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for name, (w1, ws, wm, grad, synth) in {
    "full fwd+bwd": (0.5, 0.5, 1.0, True, False), "forward only": (0.5, 0.5, 1.0, False, False),
    "L1 only fwd+bwd": (0.5, 0.0, 0.0, True, False), "SSIM only fwd+bwd": (0.0, 0.5, 0.0, True, False),
    "smooth only fwd+bwd": (0.0, 0.0, 1.0, True, False), "L1+SSIM (no smooth) fwd+bwd": (0.5, 0.5, 0.0, True, False), "full + synth/mask out": (0.5, 0.5, 1.0, True, True),
    "full + dL/dsource": (0.5, 0.5, 1.0, True, "src"),
}.items():
    plan = xptwarp.get_plan(0, B, 4, H, W, [1, 2, 4, 8], [1, 1, 1, 1], w1, ws, wm, B)
    call = plan.bind_total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"],
                                want_grad=grad, want_synth=bool(synth) and synth != "src", want_mask=bool(synth) and synth != "src",
                                want_source_grad=(synth == "src"))
    for _ in range(5): call.run()
    torch.cuda.synchronize()
    plan.profile_begin(30)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): call.run()
    e1.record(); torch.cuda.synchronize()
    d = plan.profile_end(30)
    print(f"{name:26s} fused kernel {1e3*sum(d)/len(d):8.1f} us   step {1e3*e0.elapsed_time(e1)/30:8.1f} us  ({B*H*W/(sum(d)/len(d)*1e-3)/1e9:.2f} Gpx/s kernel)")
