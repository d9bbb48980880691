"""Shared test helpers (CPU and GPU tests)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["case_small_t1", "case_n2_t2", "case_adv_t1"]


def load_case(name, kind="f32"):
    return np.load(os.path.join(GOLDEN, f"{name}_{kind}.npz"))


def case_inputs(g, device="cpu", dtype=torch.float32):
    S = sum(1 for k in g.files if k.startswith("depth_"))
    cv = lambda a: torch.tensor(a, dtype=dtype, device=device)
    feats = {"image5d": cv(g["image5d"]), "intrinsic": cv(g["intrinsic"])}
    preds = {"depth_ms": [cv(g[f"depth_{s}"]) for s in range(S)],
             "disp_ms": [cv(g[f"disp_{s}"]) for s in range(S)], "pose": cv(g["pose"])}
    lw = dict(zip(g["loss_names"].tolist(), [float(w) for w in g["loss_weights"]]))
    sw = [float(w) for w in g["scale_weights"]]
    return feats, preds, lw, sw, int(g["global_batch"])


def relerr(a, b):
    """max-abs difference relative to the largest reference magnitude."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
