"""Alternate two different input sets through ONE plan many times (eager and graph replay) and compare every
result with the first evaluation of that set: catches stale camera geometry in the constant bank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import torch, xptwarp
from xptwarp import _cabi
from oracle import xpt_oracle as orc
B, H, W = 8, 128, 384
bad = 0
for flags in (0, _cabi.XPT_FLAG_GRAPH):
    plan = xptwarp.get_plan(0, B, 4, H, W, [1, 2, 4, 8], [1, 1, 1, 1], 0.5, 0.5, 1.0, B, flags)
    sets, calls, ref = [], [], []
    for seed in (1, 2, 3):
        f, p = orc.make_inputs(B, H, W, seed=seed)
        f = {k: v.cuda() for k, v in f.items()}
        p = {"depth_ms": [d.cuda() for d in p["depth_ms"]], "disp_ms": [d.cuda() for d in p["disp_ms"]], "pose": p["pose"].cuda()}
        img = f["image5d"]
        sets.append((f, p))
        calls.append(plan.bind_total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"], want_grad=True))
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    for c in calls:
        r = c.run(); torch.cuda.synchronize()
        ref.append((r["losses"].clone(), r["d_pose"].clone(), r["d_depth_ms"][0].clone()))
    for i in range(600):
        k = (i * 7 + i // 3) % 3
        r = calls[k].run()
        if i % 50 == 49 or i < 12:
            torch.cuda.synchronize()
            ok = torch.equal(r["losses"], ref[k][0]) and torch.equal(r["d_pose"], ref[k][1]) and torch.equal(r["d_depth_ms"][0], ref[k][2])
            bad += 0 if ok else 1
    torch.cuda.synchronize()
print("alternating inputs:", "OK" if bad == 0 else f"{bad} MISMATCHES")
