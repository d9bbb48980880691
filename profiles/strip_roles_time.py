"""Busy cycles per tick of every warp role of k_strip (needs the -DXPT_STRIP_PROF build of the library:
   make -C xpt-mde-2021_b200/csrc EXTRA=-DXPT_STRIP_PROF OUT=../../profiles/variants/libxptwarp_prof.so
   XPTWARP_LIB=profiles/variants/libxptwarp_prof.so python profiles/strip_roles_time.py [cfg2|cfg3])."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import xptwarp  # noqa: E402
from xptwarp import _cabi  # noqa: E402
from xptwarp.engine import infer_scales  # noqa: E402
from oracle import xpt_oracle as orc  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
B, H, W = {"cfg2": (8, 128, 384), "cfg3": (16, 256, 832)}[wl]
feats, preds = orc.make_inputs(B, H, W, seed=5)
f = {k: v.cuda() for k, v in feats.items()}
p = {"depth_ms": [d.cuda() for d in preds["depth_ms"]], "disp_ms": [d.cuda() for d in preds["disp_ms"]], "pose": preds["pose"].cuda()}
img = f["image5d"]
lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
plan = xptwarp.get_plan(0, B, 4, H, W, infer_scales(H, preds["depth_ms"]), sw, lw["L1"], lw["SSIM"], lw["smoothe"], B, 8)
for _ in range(5):
    plan.total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"], p["pose"], want_grad=True)
torch.cuda.synchronize()
lib = _cabi.lib()
n = 148 * 32
buf = (C.c_longlong * n)()
fn = lib.xpt_debug_strip_busy
fn.argtypes = [C.c_void_p, C.c_int]
assert fn(buf, n) == 0
a = np.array(buf[:], dtype=np.float64).reshape(148, 32)[:, :24]
names = ["L", "X", "O0", "O1"] + [f"G{i}" for i in range(4)] + [f"Y{i}" for i in range(4)] + [f"S{i // 3}{i % 3}" for i in range(12)]
tot = a.max()
print(f"{wl}: busy cycles per warp (mean over CTAs), share of the busiest warp")
for i, nm in enumerate(names):
    print(f"  {nm:4s} {a[:, i].mean():12.0f}  {a[:, i].mean() / a.mean(axis=0).max():5.2f}")
