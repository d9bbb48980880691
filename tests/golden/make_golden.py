"""Generate the golden fixtures in tests/golden/*.npz.

Runs ONLY in the build container, where the reference checkout is mounted at
/root/reference:   python tests/golden/make_golden.py

It executes the reference's own, unmodified Python source for the hot path
(SynthesizeMultiScale, loss_factory -> TotalLoss, pose_rvec2matr_batch_tf/_np,
photometric_loss_*) over the TensorFlow-API shim in tf_shim.py and stores
inputs + outputs (+ gradients from torch.autograd standing in for
tape.gradient).  See tf_shim.py for what this does and does not pin.
The GPU box has no /root/reference: tests read only the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("XPT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf_shim  # noqa: E402

F64 = "--f64" in sys.argv
if F64:
    tf_shim.FLOAT = torch.float64
tf_shim.install()
sys.path.insert(0, REF)

from model.synthesize.synthesize_base import SynthesizeMultiScale  # noqa: E402
from model.loss_and_metric.loss_factory import loss_factory  # noqa: E402
import model.loss_and_metric.loss_util as lsu  # noqa: E402
import utils.convert_pose as cp  # noqa: E402
import utils.util_funcs as uf  # noqa: E402
from oracle import xpt_oracle as orc  # noqa: E402  (input generator only)

DT = torch.float64 if F64 else torch.float32
LOSS_SETS = {
    # reference config-example.py:76-89 (entries the dataset keys keep), :70-71
    "T1": ({"L1": 0.5, "L1_R": 0.5, "SSIM": 0.5, "SSIM_R": 0.5, "smoothe": 1.0, "smoothe_R": 1.0,
            "stereoL1": 0.01, "stereoSSIM": 0.01, "stereoPose": 1.0}, np.array([1.0, 1.0, 1.0, 1.0])),
    "T2": ({"L1": 0.5, "L1_R": 0.5, "SSIM": 0.5, "SSIM_R": 0.5, "smoothe": 20.0, "smoothe_R": 20.0,
            "stereoL1": 0.5, "stereoSSIM": 0.5, "stereoPose": 1.0}, np.array([0.4, 0.8, 1.2, 1.6])),
}


def make_inputs(*a, **k):
    """fp32-rounded inputs in both runs, so the f64 vectors are the exact-arithmetic
    answer for the very same inputs."""
    feats, preds = orc.make_inputs(*a, dtype=torch.float32, **k)
    cv = lambda t: t.to(DT)
    return ({k_: cv(v) for k_, v in feats.items()},
            {"depth_ms": [cv(d) for d in preds["depth_ms"]], "disp_ms": [cv(d) for d in preds["disp_ms"]],
             "pose": cv(preds["pose"])})


def np_(t):
    return t.detach().cpu().numpy()


def run_case(name, B, H, W, N, loss_set, seed, adversarial=False, global_batch=None):
    feats, preds = make_inputs(B, H, W, N=N, seed=seed, adversarial=adversarial)
    src = feats["image5d"][:, :-1].clone().requires_grad_(True)
    image5d = torch.cat([src, feats["image5d"][:, -1:]], dim=1)
    depth = [d.clone().requires_grad_(True) for d in preds["depth_ms"]]
    disp = [d.clone().requires_grad_(True) for d in preds["disp_ms"]]
    pose = preds["pose"].clone().requires_grad_(True)
    loss_weights, scale_weights = LOSS_SETS[loss_set]
    gb = B if global_batch is None else global_batch
    total_obj = loss_factory({"image": 1, "intrinsic": 1}, loss_weights, scale_weights,
                             stereo=False, batch_size=gb)
    features = {"image5d": image5d, "intrinsic": feats["intrinsic"]}
    predictions = {"depth_ms": depth, "disp_ms": disp, "pose": pose}
    total, by_type = total_obj(predictions, features)
    total.backward()
    augm = total_obj.append_data(features, predictions)
    out = {
        "B": B, "H": H, "W": W, "N": N, "global_batch": gb, "loss_set": loss_set,
        "loss_names": np.array(list(total_obj.loss_objects.keys())),
        "loss_weights": np.array([total_obj.loss_weights[k] for k in total_obj.loss_objects]),
        "scale_weights": scale_weights,
        "total": np_(total), "d_pose": np_(pose.grad),
    }
    for k, v in by_type.items():
        out["loss_" + k] = np_(v)
    for s in range(len(depth)):
        out[f"d_depth_{s}"] = np_(depth[s].grad)
        out[f"d_disp_{s}"] = np_(disp[s].grad)
    if not F64:
        out.update({"image5d": np_(feats["image5d"]), "intrinsic": np_(feats["intrinsic"]),
                    "pose": np_(preds["pose"]), "d_source": np_(src.grad)})
        for s in range(len(depth)):
            out[f"depth_{s}"] = np_(preds["depth_ms"][s])
            out[f"disp_{s}"] = np_(preds["disp_ms"][s])
            out[f"synth_{s}"] = np_(augm["synth_target_ms"][s])
            out[f"target_{s}"] = np_(augm["target_ms"][s])
    path = os.path.join(HERE, f"{name}_{'f64' if F64 else 'f32'}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: float(v) for k, v in by_type.items()}, float(total))


STEREO_SETS = {
    # reference config-example.py:76-121 as written (stereo rigs: STEREO=True, config-example.py:20)
    "T1": LOSS_SETS["T1"][0], "T2": LOSS_SETS["T2"][0],
    "MOA_WST": {"moaL1": 5.0, "moaL1_R": 5.0, "moaSSIM": 0.5, "moaSSIM_R": 0.5, "smoothe": 20.0, "smoothe_R": 20.0,
                "stereoL1": 0.5, "stereoSSIM": 0.5, "stereoPose": 1.0},
    "MD2": {"md2L1": 0.5, "md2L1_R": 0.5, "md2SSIM": 0.5, "md2SSIM_R": 0.5, "smoothe": 1.0, "smoothe_R": 1.0,
            "stereoL1": 0.5, "stereoSSIM": 0.5, "stereoPose": 1.0},
    # config-example.py:90-95 LOSS_RIGID_COMB and :108-111 LOSS_FLOW
    "COMB": {"cmbL1": 5.0, "cmbL1_R": 5.0, "cmbSSIM": 0.5, "cmbSSIM_R": 0.5, "smoothe": 20.0, "smoothe_R": 20.0,
             "stereoL1": 0.5, "stereoSSIM": 0.5, "stereoPose": 1.0},
    "FLOW": {"flowL2": 1.0, "flowL2_R": 1.0, "flow_reg": 4e-7},
}
_PRED_KEYS = ("depth_ms", "disp_ms", "pose", "depth_ms_R", "disp_ms_R", "pose_R", "pose_LR", "pose_RL",
              "flow_ms", "flow_ms_R")


def run_stereo_case(name, B, H, W, N, loss_set, scale_weights, seed, global_batch=None, flow=None):
    """TotalLoss(stereo=True) of the reference on a stereo rig: temporal losses of both eyes, the two stereo
    syntheses (losses.py:105-140), StereoDepthLoss / StereoPoseLoss / MoA / MonoDepth2 as configured."""
    feats, preds = orc.make_stereo_inputs(B, H, W, N=N, seed=seed, dtype=torch.float32)
    wreg = None
    if flow:      # "with": rigid predictions + PWC-shaped flows (LOSS_RIGID_COMB); "only": FlowNet training (LOSS_FLOW)
        if flow == "only":
            preds = {}
            g = torch.Generator().manual_seed(seed + 99)
            wreg = [(torch.randn(*shp, generator=g, dtype=torch.float64) * 3).float().to(DT).requires_grad_(True)
                    for shp in ((3, 3, 8, 16), (16,), (5, 7))]
        preds["flow_ms"] = orc.make_flow(B, H, W, N=N, seed=seed, dtype=torch.float32)
        preds["flow_ms_R"] = orc.make_flow(B, H, W, N=N, seed=seed + 13, dtype=torch.float32)
    cv = lambda t: t.to(DT)
    feats = {k: cv(v) for k, v in feats.items()}
    preds = {k: ([cv(t).clone().requires_grad_(True) for t in v] if isinstance(v, list)
                 else cv(v).clone().requires_grad_(True)) for k, v in preds.items()}
    feats["image_R"] = feats["image5d_R"]             # losses.py:38 tests this key
    gb = B if global_batch is None else global_batch
    cfg = {"image": 1, "intrinsic": 1, "image_R": 1, "intrinsic_R": 1, "stereo_T_LR": 1}
    total_obj = loss_factory(cfg, STEREO_SETS[loss_set], np.asarray(scale_weights, dtype=np.float64), stereo=True,
                             weights_to_regularize=wreg, batch_size=gb)
    total, by_type = total_obj(preds, feats)
    total.backward()
    out = {"B": B, "H": H, "W": W, "N": N, "global_batch": gb, "loss_set": loss_set,
           "loss_names": np.array(list(total_obj.loss_objects.keys())),
           "loss_weights": np.array([total_obj.loss_weights[k] for k in total_obj.loss_objects]),
           "scale_weights": np.asarray(scale_weights, dtype=np.float64), "total": np_(total)}
    for k, v in by_type.items():
        out["loss_" + k] = np_(v)
    gz = lambda t: np_(torch.zeros_like(t) if t.grad is None else t.grad)
    for k in _PRED_KEYS:
        if k not in preds:
            continue
        v = preds[k]
        if isinstance(v, list):
            for s_, t in enumerate(v):
                out[f"d_{k}_{s_}"] = gz(t)
        else:
            out["d_" + k] = gz(v)
    if wreg is not None:
        for i, w_ in enumerate(wreg):
            out[f"d_wreg_{i}"] = gz(w_)
            if not F64:
                out[f"in_wreg_{i}"] = np_(w_)
    if flow:      # the flow-warped views and their targets (augm_data of losses.py:95-101)
        augm = total_obj.append_data(feats, preds)
        for s_, (wt, ft) in enumerate(zip(augm["warped_target_ms"], augm["flow_target_ms"])):
            out[f"warped_{s_}"] = np_(wt)
            out[f"flow_target_{s_}"] = np_(ft)
    if not F64:
        for k, v in feats.items():
            if k != "image_R":
                out["in_" + k] = np_(v)
        for k in _PRED_KEYS:
            if k not in preds:
                continue
            v = preds[k]
            if isinstance(v, list):
                for s_, t in enumerate(v):
                    out[f"in_{k}_{s_}"] = np_(t)
            else:
                out["in_" + k] = np_(v)
    path = os.path.join(HERE, f"{name}_{'f64' if F64 else 'f32'}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: float(v) for k, v in by_type.items()}, float(total))


def reference_class(path, name, namespace):
    """exec ONE class of a reference module that cannot be imported whole (model_factory.py pulls in the Keras nets):
    its unmodified source lines, cut out with ast, run over the shim"""
    import ast
    src = open(os.path.join(REF, path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == name)
    code = "\n".join(src.splitlines()[node.lineno - 1:node.end_lineno])
    exec(compile(code, os.path.join(REF, path), "exec"), namespace)
    return namespace[name]


def run_logit_case(name, B, H, W, N, loss_set, seed):
    """the depth net's last op is InverseSigmoidActivation (model_factory.py:133-137), then
    predict_batch forms disp_ms = safe_reciprocal_number_ms(depth_ms) (model_wrappers.py:47-48), then TotalLoss.
    Inputs: logits; outputs: losses and dL/dlogit (autograd through the reference's activation and reciprocal)."""
    import tensorflow as tf
    act = reference_class("model/build_model/model_factory.py", "InverseSigmoidActivation", {"tf": tf, "uf": uf})()
    feats, preds = make_inputs(B, H, W, N=N, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    logit = [(torch.rand(d.shape, generator=g, dtype=torch.float64) * 6.0 - 4.5).float().to(DT).requires_grad_(True)
             for d in preds["depth_ms"]]                          # depth = 1 / (sigmoid + 0.01) in [1.2, 46]
    depth = [act(x) for x in logit]
    disp = uf.safe_reciprocal_number_ms(depth)
    pose = preds["pose"].clone().requires_grad_(True)
    loss_weights, scale_weights = LOSS_SETS[loss_set]
    total_obj = loss_factory({"image": 1, "intrinsic": 1}, loss_weights, scale_weights, stereo=False, batch_size=B)
    features = {"image5d": feats["image5d"], "intrinsic": feats["intrinsic"]}
    total, by_type = total_obj({"depth_ms": depth, "disp_ms": disp, "pose": pose}, features)
    total.backward()
    out = {"B": B, "H": H, "W": W, "N": N, "global_batch": B, "loss_set": loss_set,
           "loss_names": np.array(list(total_obj.loss_objects.keys())),
           "loss_weights": np.array([total_obj.loss_weights[k] for k in total_obj.loss_objects]),
           "scale_weights": scale_weights, "total": np_(total), "d_pose": np_(pose.grad)}
    for k, v in by_type.items():
        out["loss_" + k] = np_(v)
    for s_ in range(len(logit)):
        out[f"d_logit_{s_}"] = np_(logit[s_].grad)
        out[f"act_depth_{s_}"] = np_(depth[s_])
    if not F64:
        out.update({"image5d": np_(feats["image5d"]), "intrinsic": np_(feats["intrinsic"]), "pose": np_(preds["pose"])})
        for s_ in range(len(logit)):
            out[f"logit_{s_}"] = np_(logit[s_])
    path = os.path.join(HERE, f"{name}_{'f64' if F64 else 'f32'}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: float(v) for k, v in by_type.items()}, float(total))


def run_pieces():
    """Per-function vectors: pose conversion (tf and numpy twins), per-pixel
    photometric terms, safe reciprocal, SynthesizeMultiScale alone."""
    g = torch.Generator().manual_seed(7)
    poses = (torch.rand(3, 4, 6, generator=g, dtype=torch.float64) * 2 - 1).to(DT)
    T_tf = cp.pose_rvec2matr_batch_tf(poses)
    T_np = cp.pose_rvec2matr_batch_np(np_(poses).astype(np.float64))
    synt = (torch.rand(2, 3, 9, 11, 3, generator=g, dtype=torch.float64) * 2 - 1).to(DT)
    synt[:, :, 2:4, 3:6] = 0          # black (invalid) pixels
    synt[:, 1, 0, :] = 0
    orig = (torch.rand(2, 9, 11, 3, generator=g, dtype=torch.float64) * 2 - 1).to(DT)
    depth = (torch.rand(2, 4, 6, 1, generator=g, dtype=torch.float64) * 3).to(DT)
    depth[0, 0, 0, 0] = 0.000001
    feats, preds = make_inputs(2, 16, 24, N=2, n_scales=2, seed=11)
    synth_ms = SynthesizeMultiScale()(feats["image5d"][:, :-1], feats["intrinsic"],
                                      preds["depth_ms"], preds["pose"])
    tgt_ms = uf.multi_scale_like_depth(feats["image5d"][:, -1], preds["depth_ms"])
    out = {
        "poses": np_(poses), "T_tf": np_(T_tf), "T_np": T_np,
        "synt": np_(synt), "orig": np_(orig),
        "l1_map": np_(lsu.photometric_loss_l1(synt, orig, False)),
        "l2_map": np_(lsu.photometric_loss_l2(synt, orig, False)),
        "ssim_map": np_(lsu.photometric_loss_ssim(synt, orig, False)),
        "l1": np_(lsu.photometric_loss_l1(synt, orig)), "ssim": np_(lsu.photometric_loss_ssim(synt, orig)),
        "depth": np_(depth), "disp": np_(uf.safe_reciprocal_number(depth)),
        "syn_image5d": np_(feats["image5d"]), "syn_intrinsic": np_(feats["intrinsic"]),
        "syn_pose": np_(preds["pose"]),
    }
    for s in range(2):
        out[f"syn_depth_{s}"] = np_(preds["depth_ms"][s])
        out[f"syn_synth_{s}"] = np_(synth_ms[s])
        out[f"syn_target_{s}"] = np_(tgt_ms[s])
    path = os.path.join(HERE, f"pieces_{'f64' if F64 else 'f32'}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    if "--logit-only" in sys.argv:
        run_logit_case("logit_t1", B=2, H=32, W=64, N=4, loss_set="T1", seed=111)
        sys.exit(0)
    run_pieces()
    run_logit_case("logit_t1", B=2, H=32, W=64, N=4, loss_set="T1", seed=111)
    run_case("case_small_t1", B=2, H=32, W=64, N=4, loss_set="T1", seed=101)
    run_case("case_n2_t2", B=3, H=48, W=40, N=2, loss_set="T2", seed=202, global_batch=6)
    run_case("case_adv_t1", B=2, H=32, W=48, N=4, loss_set="T1", seed=303, adversarial=True)
    run_stereo_case("stereo_t1", B=2, H=32, W=64, N=2, loss_set="T1", scale_weights=[1, 1, 1, 1], seed=404)
    run_stereo_case("stereo_t2", B=2, H=32, W=40, N=3, loss_set="T2", scale_weights=[0.4, 0.8, 1.2, 1.6], seed=505, global_batch=4)
    run_stereo_case("stereo_moa", B=2, H=32, W=64, N=2, loss_set="MOA_WST", scale_weights=[1, 1, 1, 1], seed=606)
    run_stereo_case("stereo_md2", B=2, H=32, W=48, N=3, loss_set="MD2", scale_weights=[1, 1, 1, 1], seed=707)
    run_stereo_case("stereo_comb", B=2, H=64, W=96, N=2, loss_set="COMB", scale_weights=[0.4, 0.8, 1.2, 1.6], seed=808,
                    flow="with")
    run_stereo_case("stereo_flow", B=2, H=64, W=96, N=3, loss_set="FLOW", scale_weights=[1, 1, 1, 1], seed=909,
                    global_batch=4, flow="only")
