"""Oracle vs the golden vectors produced by the reference's own Python source
(tests/golden/make_golden.py).  Pins the oracle before anything trusts it."""
import numpy as np
import pytest
import torch

from oracle import xpt_oracle as orc
from helpers import (CASES, FLOW_CASES, FLOW_KEYS, PRED_KEYS, STEREO_CASES, case_inputs, case_reg_weights, golden_grad,
                     load_case, relerr, stereo_case_inputs)


def test_pieces_pose_and_photometric_maps():
    g = load_case("pieces")
    T = orc.pose_rvec2matr_batch(torch.tensor(g["poses"])).numpy()
    assert np.abs(T - g["T_tf"]).max() < 1e-6
    # the reference's numpy twin (utils/convert_pose.py:74-111), run in fp64
    assert np.abs(T - g["T_np"]).max() < 1e-6
    synt, orig = torch.tensor(g["synt"]), torch.tensor(g["orig"])
    assert np.abs(orc.photometric_loss_l1(synt, orig, False).numpy() - g["l1_map"]).max() < 1e-6
    assert np.abs(orc.photometric_loss_l2(synt, orig, False).numpy() - g["l2_map"]).max() < 1e-6
    assert np.abs(orc.photometric_loss_ssim(synt, orig, False).numpy() - g["ssim_map"]).max() < 2e-5
    assert relerr(orc.photometric_loss_l1(synt, orig).numpy(), g["l1"]) < 1e-6
    assert relerr(orc.photometric_loss_ssim(synt, orig).numpy(), g["ssim"]) < 1e-5
    d = torch.tensor(g["depth"])
    assert np.allclose(orc.safe_reciprocal_number(d).numpy(), g["disp"], rtol=1e-6)


def test_pieces_synthesize_multi_scale():
    g = load_case("pieces")
    img = torch.tensor(g["syn_image5d"])
    depth_ms = [torch.tensor(g[f"syn_depth_{s}"]) for s in range(2)]
    out = orc.synthesize_multi_scale(img[:, :-1], torch.tensor(g["syn_intrinsic"]), depth_ms,
                                     torch.tensor(g["syn_pose"]))
    tgt = orc.multi_scale_like_depth(img[:, -1], depth_ms)
    for s in range(2):
        assert np.abs(out[s].numpy() - g[f"syn_synth_{s}"]).max() < 1e-5
        assert np.abs(tgt[s].numpy() - g[f"syn_target_{s}"]).max() < 1e-6


@pytest.mark.parametrize("name", CASES)
def test_total_loss_and_gradients_fp32(name):
    g = load_case(name)
    feats, preds, lw, sw, gb = case_inputs(g)
    r = orc.loss_and_grads(feats, preds, lw, sw, gb, want_source_grad=True)
    assert relerr(r["total"].numpy(), g["total"]) < 1e-5
    for k in lw:
        assert relerr(r["by_type"][k].numpy(), g["loss_" + k]) < 1e-5
    for s in range(len(preds["depth_ms"])):
        assert np.abs(r["synth_ms"][s].numpy() - g[f"synth_{s}"]).max() < 1e-5
        assert np.abs(r["target_ms"][s].numpy() - g[f"target_{s}"]).max() < 1e-6
        assert relerr(r["d_depth_ms"][s].numpy(), g[f"d_depth_{s}"]) < 1e-4
        assert relerr(r["d_disp_ms"][s].numpy(), g[f"d_disp_{s}"]) < 1e-5
    assert relerr(r["d_pose"].numpy(), g["d_pose"]) < 1e-4
    assert relerr(r["d_source"].numpy(), g["d_source"]) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_total_loss_and_gradients_fp64(name):
    g, g64 = load_case(name), load_case(name, "f64")
    feats, preds, lw, sw, gb = case_inputs(g, dtype=torch.float64)
    r = orc.loss_and_grads(feats, preds, lw, sw, gb)
    assert relerr(r["total"].numpy(), g64["total"]) < 1e-12
    assert relerr(r["d_pose"].numpy(), g64["d_pose"]) < 1e-9
    for s in range(len(preds["depth_ms"])):
        assert relerr(r["d_depth_ms"][s].numpy(), g64[f"d_depth_{s}"]) < 1e-9
        assert relerr(r["d_disp_ms"][s].numpy(), g64[f"d_disp_{s}"]) < 1e-9


@pytest.mark.parametrize("name", STEREO_CASES)
@pytest.mark.parametrize("kind", ["f32", "f64"])
def test_stereo_total_loss_and_gradients(name, kind):
    """TotalLoss(stereo=True): both eyes' temporal losses, the two stereo syntheses, StereoDepthLoss,
    StereoPoseLoss, MoA and MonoDepth2 min-over-sources -- against the reference's own source."""
    g32, g = load_case(name), load_case(name, kind)
    dt = torch.float64 if kind == "f64" else torch.float32
    feats, preds, lw, sw, gb = stereo_case_inputs(g32, dtype=dt)
    r = orc.stereo_loss_and_grads(feats, preds, lw, sw, gb)
    ltol, gtol = (1e-12, 1e-9) if kind == "f64" else (1e-5, 1e-4)
    assert relerr(r["total"].numpy(), g["total"]) < ltol
    for k in lw:
        assert relerr(r["by_type"][k].numpy(), g["loss_" + k]) < max(ltol, 2e-5 if kind == "f32" else 0), k
    for k in PRED_KEYS:
        ref = golden_grad(g, k)
        got = r["grads"][k]
        if isinstance(ref, list):
            for s in range(len(ref)):
                assert relerr(got[s].numpy(), ref[s]) < gtol, (k, s)
        else:
            assert relerr(got.numpy(), ref) < gtol, k


@pytest.mark.parametrize("name", FLOW_CASES)
@pytest.mark.parametrize("kind", ["f32", "f64"])
def test_flow_total_loss_and_gradients(name, kind):
    """FlowWarpMultiScale + flowL2 + flow_reg (LOSS_FLOW) and CombinedLossMultiScale (LOSS_RIGID_COMB) on a stereo
    rig -- against the reference's own source (flow_warping.py, losses.py:235-279,497-534)."""
    g32, g = load_case(name), load_case(name, kind)
    dt = torch.float64 if kind == "f64" else torch.float32
    feats, preds, lw, sw, gb = stereo_case_inputs(g32, dtype=dt)
    wreg = case_reg_weights(g32, dtype=dt)
    r = orc.stereo_loss_and_grads(feats, preds, lw, sw, gb, weights_to_regularize=wreg)
    ltol, gtol = (1e-12, 1e-9) if kind == "f64" else (1e-5, 1e-4)
    assert relerr(r["total"].numpy(), g["total"]) < ltol
    for k in lw:
        assert relerr(r["by_type"][k].numpy(), g["loss_" + k]) < max(ltol, 2e-5 if kind == "f32" else 0), k
    for s in range(len(preds["flow_ms"])):
        assert np.abs(r["augm"]["warped_target_ms"][s].numpy() - g[f"warped_{s}"]).max() < (1e-5 if kind == "f32" else 1e-12)
        assert np.abs(r["augm"]["flow_target_ms"][s].numpy() - g[f"flow_target_{s}"]).max() < (1e-6 if kind == "f32" else 1e-12)
    for k in PRED_KEYS + FLOW_KEYS:
        if k not in preds:
            continue
        ref, got = golden_grad(g, k), r["grads"][k]
        if isinstance(ref, list):
            for s in range(len(ref)):
                assert relerr(got[s].numpy(), ref[s]) < gtol, (k, s)
        else:
            assert relerr(got.numpy(), ref) < gtol, k
    if wreg is not None:
        for i, w in enumerate(r["grads"]["weights_to_regularize"]):
            assert relerr(w.numpy(), g[f"d_wreg_{i}"]) < gtol


def test_pose_matr2rvec_round_trip():
    """reference test_pose_matr2rvec_batch (convert_pose.py:256-271): rvec -> matrix -> rvec for U(-1,1)."""
    g = torch.Generator().manual_seed(3)
    p = torch.rand(8, 5, 6, generator=g, dtype=torch.float64) * 2 - 1
    assert float((orc.pose_matr2rvec_batch(orc.pose_rvec2matr_batch(p)) - p).abs().max()) < 1e-9


def test_gradients_match_finite_differences_fp64():
    """No reference test pins gradients: check the autograd answer
    against central differences on pose in fp64."""
    feats, preds = orc.make_inputs(1, 16, 24, N=2, seed=5, dtype=torch.float64)
    lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
    r = orc.loss_and_grads(feats, preds, lw, sw)
    eps = 1e-6
    for idx in [(0, 0, 0), (0, 1, 2), (0, 0, 3), (0, 1, 5)]:
        vals = []
        for sgn in (+1, -1):
            p = {k: v for k, v in preds.items()}
            pose = preds["pose"].clone()
            pose[idx] += sgn * eps
            p["pose"] = pose
            vals.append(orc.total_loss(p, feats, lw, sw)[0].item())
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - r["d_pose"][idx].item()) < 1e-5 * max(1.0, abs(fd)) + 1e-7


def test_depth_activation_against_the_reference_class():
    """oracle.inverse_sigmoid_activation against vectors produced by the reference's own
    InverseSigmoidActivation source (model_factory.py:133-137, cut out and executed by make_golden.py), and the
    total loss + dL/dlogit through it."""
    import torch
    from helpers import load_case
    from oracle import xpt_oracle as orc
    for kind, dt, tol in (("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-12)):
        g = load_case("logit_t1", kind)
        g32 = load_case("logit_t1", "f32")
        S = sum(1 for k in g32.files if k.startswith("logit_"))
        logit = [torch.tensor(g32[f"logit_{s}"], dtype=dt).requires_grad_(True) for s in range(S)]
        depth = [orc.inverse_sigmoid_activation(x) for x in logit]
        for s in range(S):
            assert relerr(depth[s].detach().numpy(), g[f"act_depth_{s}"]) < tol
        feats = {"image5d": torch.tensor(g32["image5d"], dtype=dt), "intrinsic": torch.tensor(g32["intrinsic"], dtype=dt)}
        pose = torch.tensor(g32["pose"], dtype=dt).requires_grad_(True)
        lw = dict(zip(g["loss_names"].tolist(), [float(w) for w in g["loss_weights"]]))
        total, by_type = orc.total_loss({"depth_ms": depth, "disp_ms": [orc.safe_reciprocal_number(d) for d in depth],
                                         "pose": pose}, feats, lw, [float(w) for w in g["scale_weights"]], int(g["global_batch"]))
        total.backward()
        assert relerr(total.detach().numpy(), g["total"]) < (1e-5 if kind == "f32" else 1e-9)
        for s in range(S):
            assert relerr(logit[s].grad.numpy(), g[f"d_logit_{s}"]) < (1e-4 if kind == "f32" else 1e-8), s
