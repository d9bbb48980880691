"""A/B of the streaming strip kernel (k_strip, default) against the round-1 tile kernel (XPT_FLAG_TILES) on the same
inputs: max relative differences of every output, then device time of both on config 2 / config 3.
usage: python profiles/strip_debug.py [quick]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xptwarp  # noqa: E402
from xptwarp.engine import infer_scales  # noqa: E402
from oracle import xpt_oracle as orc  # noqa: E402

STRIP = 8


def rel(a, b):
    a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(B, H, W, N, S, lw, sw, derive=False, seed=1):
    feats, preds = orc.make_inputs(B, H, W, N=N, n_scales=S, seed=seed)
    f = {k: v.cuda() for k, v in feats.items()}
    p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]],
         "pose": preds["pose"].cuda()}
    img = f["image5d"]
    out = {}
    for name, flags in (("strip", STRIP), ("tiles", 0)):
        plan = xptwarp.get_plan(0, B, N, H, W, infer_scales(H, preds["depth_ms"]), sw, lw.get("L1", 0.0),
                                lw.get("SSIM", 0.0), lw.get("smoothe", 0.0), B, flags)
        r = plan.total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], None if derive else p["disp_ms"],
                            p["pose"], want_grad=True)
        torch.cuda.synchronize()
        out[name] = r
    a, b = out["strip"], out["tiles"]
    msg = [f"B{B} {H}x{W} N{N} S{S} derive={int(derive)} lw={lw}:"]
    msg.append("losses " + " ".join(f"{rel(a['losses'][k], b['losses'][k]):.1e}" for k in range(4)))
    msg.append("d_pose %.1e" % rel(a["d_pose"], b["d_pose"]))
    for s in range(S):
        msg.append(f"d_depth[{s}] {rel(a['d_depth_ms'][s], b['d_depth_ms'][s]):.1e}")
        if not derive and a.get("d_disp_ms") is not None:
            msg.append(f"d_disp[{s}] {rel(a['d_disp_ms'][s], b['d_disp_ms'][s]):.1e}")
    print(" ".join(msg), flush=True)
    return out


def timeit(B, H, W, N=4, S=4, iters=50):
    feats, preds = orc.make_inputs(B, H, W, N=N, n_scales=S, seed=5)
    f = {k: v.cuda() for k, v in feats.items()}
    p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]],
         "pose": preds["pose"].cuda()}
    img = f["image5d"]
    lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
    for name, flags in (("strip", STRIP | 2), ("tiles", 2)):
        plan = xptwarp.get_plan(0, B, N, H, W, infer_scales(H, preds["depth_ms"]), sw, lw["L1"], lw["SSIM"], lw["smoothe"],
                                B, flags)
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
        print(f"time B{B} {H}x{W}: {name} {ms*1e3:.1f} us/step  {B*H*W/ms/1e6:.2f} Gpx/s", flush=True)


if __name__ == "__main__":
    T1, SW = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
    run(1, 16, 16, 1, 1, T1, [1.0])
    run(2, 32, 64, 4, 4, T1, SW)
    run(2, 32, 64, 4, 4, {"L1": 1.0}, SW)
    run(2, 32, 64, 4, 4, {"SSIM": 1.0}, SW)
    run(2, 32, 64, 4, 4, {"smoothe": 1.0}, SW)
    run(2, 40, 72, 3, 2, T1, [1.0, 0.5])
    run(1, 72, 88, 2, 4, T1, SW)
    run(2, 64, 96, 4, 4, T1, SW, derive=True)
    run(3, 128, 384, 4, 4, T1, SW)
    run(2, 128, 384, 1, 4, T1, SW)
    if len(sys.argv) < 2:
        timeit(8, 128, 384)
        timeit(16, 256, 832, iters=20)
