"""Wall time per training-style step (TotalLoss forward + backward through the public API) of every loss set of
the reference's config on a stereo rig, config-2 frame size.  Run on the GPU box: python profiles/loss_sets.py [B H W]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
import numpy as np, torch, xptwarp
from xptwarp import engine
from oracle import xpt_oracle as orc

B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 128, 384)
SETS = {
    "LOSS_RIGID_T1": {"L1": .5, "L1_R": .5, "SSIM": .5, "SSIM_R": .5, "smoothe": 1., "smoothe_R": 1., "stereoL1": .01, "stereoSSIM": .01, "stereoPose": 1.},
    "LOSS_RIGID_T2": {"L1": .5, "L1_R": .5, "SSIM": .5, "SSIM_R": .5, "smoothe": 20., "smoothe_R": 20., "stereoL1": .5, "stereoSSIM": .5, "stereoPose": 1.},
    "LOSS_RIGID_MOA_WST": {"moaL1": 5., "moaL1_R": 5., "moaSSIM": .5, "moaSSIM_R": .5, "smoothe": 20., "smoothe_R": 20., "stereoL1": .5, "stereoSSIM": .5, "stereoPose": 1.},
    "LOSS_RIGID_MD2": {"md2L1": .5, "md2L1_R": .5, "md2SSIM": .5, "md2SSIM_R": .5, "smoothe": 1., "smoothe_R": 1., "stereoL1": .5, "stereoSSIM": .5, "stereoPose": 1.},
    "LOSS_RIGID_COMB": {"cmbL1": 5., "cmbL1_R": 5., "cmbSSIM": .5, "cmbSSIM_R": .5, "smoothe": 20., "smoothe_R": 20., "stereoL1": .5, "stereoSSIM": .5, "stereoPose": 1.},
    "LOSS_FLOW": {"flowL2": 1., "flowL2_R": 1.},
}
feats, preds = orc.make_stereo_inputs(B, H, W, seed=3)
flow = {"flow_ms": orc.make_flow(B, H, W, seed=4), "flow_ms_R": orc.make_flow(B, H, W, seed=5)}
f = {k: v.cuda() for k, v in feats.items()}
cfg = {"image": 1, "intrinsic": 1, "image_R": 1, "intrinsic_R": 1, "stereo_T_LR": 1}
for name, lw in SETS.items():
    if os.environ.get("XPT_LS_ONLY") and name != os.environ["XPT_LS_ONLY"]:
        continue
    src = dict(preds)
    if name == "LOSS_RIGID_COMB":
        src.update(flow)
    if name == "LOSS_FLOW":
        src = dict(flow)
    p = {k: ([t.cuda().requires_grad_(True) for t in v] if isinstance(v, list) else v.cuda().requires_grad_(True)) for k, v in src.items()}
    tot = xptwarp.loss_factory(cfg, lw, np.array([1., 1., 1., 1.]), stereo=True, batch_size=B)

    def step():
        for v in p.values():
            for t in (v if isinstance(v, list) else [v]):
                t.grad = None
        total, _ = tot(p, f)
        total.backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    n = int(os.environ.get("XPT_LS_STEPS", "10"))
    t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    # the same step captured once with torch.cuda.graph and replayed (no Python on the path)
    gdt = float("nan")
    try:
        for v in p.values():
            for t in (v if isinstance(v, list) else [v]):
                t.grad = None
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
            for v in p.values():
                for t in (v if isinstance(v, list) else [v]):
                    t.grad = None
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            total, _ = tot(p, f)
            total.backward()
        for _ in range(3): graph.replay()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n): graph.replay()
        torch.cuda.synchronize()
        gdt = (time.perf_counter() - t0) / n
    except Exception as e:      # a set whose objects cannot be captured is reported, not fatal
        print("   graph capture failed:", type(e).__name__, str(e)[:120])
        torch.cuda.synchronize()
    print(f"{name:20s} eager {dt*1e3:7.2f} ms/step ({2*B*H*W/dt/1e9:6.3f} Gpixel/s both eyes)   "
          f"torch.cuda.graph replay {gdt*1e3:7.2f} ms/step ({2*B*H*W/gdt/1e9:6.3f} Gpixel/s)   plans cached {len(engine._PLANS)}")
