"""Device time of one training step (pyramids + k_strip + epilogue, CUDA-graph replay) at config 2 and config 3 for
the library selected by XPTWARP_LIB (default: in-tree).  usage: python profiles/strip_time.py [tiles]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import xptwarp  # noqa: E402
from xptwarp.engine import infer_scales  # noqa: E402
from oracle import xpt_oracle as orc  # noqa: E402

flags = 2 | (0 if "tiles" in sys.argv[1:] else 8)
out = []
for (B, H, W, iters) in ((8, 128, 384, 100), (16, 256, 832, 30)):
    feats, preds = orc.make_inputs(B, H, W, seed=5)
    f = {k: v.cuda() for k, v in feats.items()}
    p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]], "pose": preds["pose"].cuda()}
    img = f["image5d"]
    lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
    plan = xptwarp.get_plan(0, B, 4, H, W, infer_scales(H, preds["depth_ms"]), sw, lw["L1"], lw["SSIM"], lw["smoothe"], B, flags)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(10):
            plan.total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"], want_grad=True)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            plan.total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"], want_grad=True)
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out.append(f"B{B} {H}x{W}: {ms * 1e3:7.1f} us {B * H * W / ms / 1e6:5.2f} Gpx/s")
print(os.environ.get("XPTWARP_LIB", "tree").split("_")[-1], " | ".join(out), flush=True)
