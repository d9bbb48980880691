"""CPU oracle for the xpt-mde view-synthesis + photometric/smoothness loss path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path runs on
hand-written CUDA kernels and fails loudly if they are missing.

What it is: an op-for-op restatement, on torch-CPU tensors, of the reference's
TensorFlow graph for this path, so that autograd supplies the gradients the
reference gets from ``tape.gradient`` (reference ``model/train_val.py:85``).
``dtype`` selects fp32 (the parity target) or fp64 (the tie-breaker).

Pinning status (see DESIGN.md "Oracle"):
  * pinned by the reference's own known-answer tests (tests/test_oracle_kat.py):
    scale_intrinsic, transform_to_source, bilinear weights / valid mask /
    reconstruction, Rodrigues sign + angle + translation, interior 3x3 mean.
  * pinned against the reference's *own Python source* executed in the build
    container over a TensorFlow-API shim (tests/golden/make_golden.py ->
    tests/golden/*.npz): op order, shapes, slicing, weights, aggregation.
  * PARITY UNPINNED against real TensorFlow 2.4.1 kernels: TF is not
    installable here (no network), so tf.image.resize half-pixel behaviour,
    SAME avg-pool border divisors and TF's autodiff rules are restated from the
    TF op definitions, not observed.

Each function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# constants read inside the path (reference config-example.py:67-71, 76-89)
IMAGE_GRADIENT_FACTOR = 4.0
SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2
LOSS_RIGID_T1 = {"L1": 0.5, "SSIM": 0.5, "smoothe": 1.0}
LOSS_RIGID_T2 = {"L1": 0.5, "SSIM": 0.5, "smoothe": 20.0}
SCALE_WEIGHT_T1 = (1.0, 1.0, 1.0, 1.0)
SCALE_WEIGHT_T2 = (0.4, 0.8, 1.2, 1.6)


# --------------------------------------------------------------------------
# pose  (reference utils/convert_pose.py:32-71)
# --------------------------------------------------------------------------
def pose_rvec2matr_batch(poses: torch.Tensor) -> torch.Tensor:
    """[B,N,6] (t, rotation vector) -> [B,N,4,4]; negated skew matrix
    (utils/convert_pose.py:53-56), identity where |theta| < 1e-8 (:65)."""
    B, N, _ = poses.shape
    p = poses.unsqueeze(-1)                       # [B,N,6,1]
    trans = p[:, :, :3]
    uvec = p[:, :, 3:]
    unorm = torch.linalg.vector_norm(uvec, dim=2, keepdim=True)   # [B,N,1,1]
    uvec = uvec / unorm
    w1, w2, w3 = uvec[:, :, 0:1], uvec[:, :, 1:2], uvec[:, :, 2:3]
    z = torch.zeros_like(w1)
    w_hat = torch.cat([z, w3, -w2, -w3, z, w1, w2, -w1, z], dim=2).reshape(B, N, 3, 3)
    eye = torch.eye(3, dtype=poses.dtype).expand(B, N, 3, 3)
    tmp = eye + w_hat * torch.sin(unorm) + torch.matmul(w_hat, w_hat) * (1 - torch.cos(unorm))
    rot = torch.where(unorm.abs() < 1e-8, eye, tmp)
    tmat = torch.cat([rot, trans], dim=3)
    last = torch.tensor([0, 0, 0, 1], dtype=poses.dtype).expand(B, N, 1, 4)
    return torch.cat([tmat, last], dim=2)


# --------------------------------------------------------------------------
# pyramids  (reference synthesize_base.py:74-85, util_funcs.py:163-175)
# --------------------------------------------------------------------------
def pose_matr2rvec_batch(poses: torch.Tensor) -> torch.Tensor:
    """utils/convert_pose.py:151-168: [B,N,4,4] -> [B,N,6] = (t, rvec).  theta = acos((tr R - 1)/2);
    axis from the antisymmetric part (same flipped sign convention as rvec2matr); |theta| < 1e-5 -> axis/2."""
    R = poses[:, :, :3, :3]
    theta = torch.acos((R[:, :, 0, 0] + R[:, :, 1, 1] + R[:, :, 2, 2] - 1.0) / 2.0).unsqueeze(-1)
    axis = torch.stack([R[:, :, 1, 2] - R[:, :, 2, 1], R[:, :, 2, 0] - R[:, :, 0, 2],
                        R[:, :, 0, 1] - R[:, :, 1, 0]], dim=-1)
    rvec = torch.where(theta.abs() < 0.00001, axis / 2.0, axis / (2 * torch.sin(theta)) * theta)
    return torch.cat([poses[:, :, :3, 3], rvec], dim=-1)


def resize_bilinear_tf(img_nhwc: torch.Tensor, size_hw: Tuple[int, int]) -> torch.Tensor:
    """tf.image.resize(method="bilinear") of TF2: half-pixel centres, no antialias.
    img [M,H,W,C] -> [M,h,w,C].  Identity when the size is unchanged."""
    M, H, W, C = img_nhwc.shape
    h, w = size_hw
    if (h, w) == (H, W):
        return img_nhwc
    x = img_nhwc.permute(0, 3, 1, 2)
    y = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False, antialias=False)
    return y.permute(0, 2, 3, 1)


def multi_scale_like_depth(image: torch.Tensor, depth_ms: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """utils/util_funcs.py:163-175: target pyramid sized like each depth map."""
    return [resize_bilinear_tf(image, (d.shape[1], d.shape[2])) for d in depth_ms]


def safe_reciprocal_number(x: torch.Tensor) -> torch.Tensor:
    """utils/util_funcs.py:157-160."""
    mask = (x > 0.00001).to(x.dtype)
    return (1.0 / x) * mask


def inverse_sigmoid_activation(x: torch.Tensor) -> torch.Tensor:
    """model/build_model/model_factory.py:133-137 InverseSigmoidActivation: the depth nets' last op,
    depth = safe_reciprocal_number(sigmoid(x) + 0.01)."""
    return safe_reciprocal_number(torch.sigmoid(x) + 0.01)


# --------------------------------------------------------------------------
# synthesis  (reference synthesize_base.py:39-178, bilinear_interp.py:7-147)
# --------------------------------------------------------------------------
def scale_intrinsic(intrinsic: torch.Tensor, scale) -> torch.Tensor:
    """synthesize_base.py:66-71."""
    B = intrinsic.shape[0]
    scaled = intrinsic[:, :2, :] / scale
    const = torch.tensor([[[0, 0, 1]]], dtype=intrinsic.dtype).expand(B, 1, 3)
    return torch.cat([scaled, const], dim=1)


def pixel_meshgrid(height: int, width: int, dtype) -> torch.Tensor:
    """synthesize_base.py:114-124 -> [3, H*W], rows (u, v, 1), u fastest."""
    v = torch.arange(height, dtype=dtype)
    u = torch.arange(width, dtype=dtype)
    vg, ug = torch.meshgrid(v, u, indexing="ij")
    return torch.stack([ug.reshape(-1), vg.reshape(-1), torch.ones(height * width, dtype=dtype)], dim=0)


def pixel2cam(pixel_coords: torch.Tensor, depth: torch.Tensor, intrinsic: torch.Tensor) -> torch.Tensor:
    """synthesize_base.py:126-146 -> [B,4,P]."""
    B = depth.shape[0]
    d = depth.reshape(B, 1, -1)
    cam = torch.matmul(torch.linalg.inv(intrinsic), pixel_coords)      # tensordot over K's last dim
    cam = cam * d
    ones = torch.ones(B, 1, cam.shape[2], dtype=cam.dtype)
    return torch.cat([cam, ones], dim=1)


def transform_to_source(tgt_coords: torch.Tensor, t2s_pose: torch.Tensor) -> torch.Tensor:
    """synthesize_base.py:149-159 -> [B,N,4,P]."""
    N = t2s_pose.shape[1]
    x = tgt_coords.unsqueeze(1).expand(-1, N, -1, -1)
    return torch.matmul(t2s_pose, x)


def cam2pixel(cam_coords: torch.Tensor, intrinsic: torch.Tensor) -> torch.Tensor:
    """synthesize_base.py:161-178 -> [B,N,3,P]; no positive-depth test."""
    K = intrinsic.unsqueeze(1)
    pix = torch.matmul(K, cam_coords[:, :, :3, :])
    return pix / (pix[:, :, 2:3, :] + 1e-10)


def neighbor_int_pixels(pixel_coords, height, width):
    """bilinear_interp.py:34-50 -> (u_floor, u_ceil, v_floor, v_ceil) [B,N,4,P]."""
    u = pixel_coords[:, :, 0:1, :]
    uf = torch.floor(u)
    uc = torch.clamp(uf + 1, 0, width - 1)
    uf = torch.clamp(uf, 0, width - 1)
    v = pixel_coords[:, :, 1:2, :]
    vf = torch.floor(v)
    vc = torch.clamp(vf + 1, 0, height - 1)
    vf = torch.clamp(vf, 0, height - 1)
    return torch.cat([uf, uc, vf, vc], dim=2)


def make_valid_mask(pixel_floorceil, valid_mask, batch):
    """bilinear_interp.py:53-76 -> [B,N,1,P] float {0,1}."""
    uf, uc = pixel_floorceil[:, :, 0:1], pixel_floorceil[:, :, 1:2]
    vf, vc = pixel_floorceil[:, :, 2:3], pixel_floorceil[:, :, 3:4]
    mask = (uf + 1 == uc) & (vf + 1 == vc)
    if valid_mask is not None:
        nz = valid_mask.reshape(batch, 1, 1, -1) != 0
        mask = mask & nz
    return mask.to(pixel_floorceil.dtype)


def calc_neighbor_weights(pixel_coords, pixel_floorceil, valid_mask):
    """bilinear_interp.py:79-100 -> (w_ufvf, w_ufvc, w_ucvf, w_ucvc) [B,N,4,P]."""
    u, v = pixel_coords[:, :, 0:1], pixel_coords[:, :, 1:2]
    uf, uc = pixel_floorceil[:, :, 0:1], pixel_floorceil[:, :, 1:2]
    vf, vc = pixel_floorceil[:, :, 2:3], pixel_floorceil[:, :, 3:4]
    w_uf, w_uc = uc - u, u - uf
    w_vf, w_vc = vc - v, v - vf
    w = torch.cat([w_uf * w_vf, w_uf * w_vc, w_uc * w_vf, w_uc * w_vc], dim=2)
    return w * valid_mask


def sample_neighbor_images(source_image, pixel_floorceil):
    """bilinear_interp.py:103-133 -> [B,N,4,P,C] (gather_nd with batch_dims=2)."""
    B, N, H, W, C = source_image.shape
    idx = pixel_floorceil.detach().to(torch.int64)
    uf, uc, vf, vc = idx[:, :, 0], idx[:, :, 1], idx[:, :, 2], idx[:, :, 3]
    flat = source_image.reshape(B, N, H * W, C)

    def g(vv, uu):
        lin = (vv * W + uu).unsqueeze(-1).expand(-1, -1, -1, C)
        return torch.gather(flat, 2, lin)
    return torch.stack([g(vf, uf), g(vc, uf), g(vf, uc), g(vc, uc)], dim=2)


def bilinear_interpolation(image, pixel_coords, valid_mask=None, return_mask=False):
    """bilinear_interp.py:7-32.  image [B,N,H,W,C], coords [B,N,>=2,P]."""
    B, N, H, W, C = image.shape
    # NaN coordinates: tf.cast(nan, int32) is undefined; they are invalid anyway
    # (SURVEY.md A.3), so make the index finite and the mask false.
    finite = torch.isfinite(pixel_coords[:, :, 0:1]) & torch.isfinite(pixel_coords[:, :, 1:2])
    safe_coords = torch.where(finite, pixel_coords[:, :, :2], torch.full_like(pixel_coords[:, :, :2], -10.0))
    fc = neighbor_int_pixels(safe_coords, H, W)
    mask = make_valid_mask(fc, valid_mask, B)
    weights = calc_neighbor_weights(safe_coords, fc, mask)
    sampled = sample_neighbor_images(image, fc)
    merged = (sampled * weights.unsqueeze(-1)).sum(dim=2)
    out = merged.reshape(B, N, H, W, C)
    if return_mask:
        return out, mask.reshape(B, N, H, W, 1)
    return out


def warp_pixel_coords(depth_sc, poses_matr, intrinsic_sc):
    """synthesize_base.py:106-112."""
    B, H, W, _ = depth_sc.shape
    grid = pixel_meshgrid(H, W, depth_sc.dtype)
    cam = pixel2cam(grid, depth_sc, intrinsic_sc)
    src = transform_to_source(cam, poses_matr)
    return cam2pixel(src, intrinsic_sc)


def synthesize_single_scale(source_image, intrinsic, depth_sc, poses_matr, return_mask=False):
    """synthesize_base.py:39-58."""
    B, N, H, W, C = source_image.shape
    _, Hs, Ws, _ = depth_sc.shape
    scale = int(H // Hs)
    K_sc = scale_intrinsic(intrinsic, scale)
    src_sc = resize_bilinear_tf(source_image.reshape(B * N, H, W, C), (Hs, Ws)).reshape(B, N, Hs, Ws, C)
    coords = warp_pixel_coords(depth_sc, poses_matr, K_sc)
    return bilinear_interpolation(src_sc, coords, depth_sc, return_mask=return_mask)


def synthesize_multi_scale(source_image, intrinsic, pred_depth_ms, pred_pose, return_mask=False):
    """synthesize_base.py:10-29."""
    T = pose_rvec2matr_batch(pred_pose)
    outs = [synthesize_single_scale(source_image, intrinsic, d, T, return_mask) for d in pred_depth_ms]
    if return_mask:
        return [o[0] for o in outs], [o[1] for o in outs]
    return outs


def flow_to_pixel_coordinates(flow):
    """flow_warping.py:51-71: [B,N,h,w,2] -> [B,N,2,h*w] (grid - flow)."""
    B, N, h, w, _ = flow.shape
    v = torch.arange(h, dtype=flow.dtype)
    u = torch.arange(w, dtype=flow.dtype)
    vg, ug = torch.meshgrid(v, u, indexing="ij")
    grid = torch.stack([ug, vg], dim=0).reshape(1, 1, 2, -1)
    uvflow = flow.reshape(B, N, -1, 2).permute(0, 1, 3, 2)
    return grid - uvflow


def flow_warp_multi_scale(source_image, flow_ms):
    """flow_warping.py:11-49."""
    B, N, H, W, C = source_image.shape
    outs = []
    for flow in flow_ms:
        h, w = flow.shape[2], flow.shape[3]
        src = resize_bilinear_tf(source_image.reshape(B * N, H, W, C), (h, w)).reshape(B, N, h, w, C)
        outs.append(bilinear_interpolation(src, flow_to_pixel_coordinates(flow)))
    return outs


# --------------------------------------------------------------------------
# photometric terms  (reference loss_util.py)
# --------------------------------------------------------------------------
def _avg_pool_same_3x3(x5: torch.Tensor) -> torch.Tensor:
    """AveragePooling3D(pool=(1,3,3), strides=1, padding="SAME") on [B,N,H,W,C]:
    divisor counts in-image taps only (loss_util.py:78)."""
    B, N, H, W, C = x5.shape
    x = x5.permute(0, 1, 4, 2, 3).reshape(B * N * C, 1, H, W)
    y = F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False)
    return y.reshape(B, N, C, H, W).permute(0, 1, 3, 4, 2)


def photometric_loss_l1(synt_target, orig_target, reduce=True):
    """loss_util.py:6-25."""
    orig = orig_target.unsqueeze(1)
    gray = synt_target.mean(dim=-1, keepdim=True)
    err_mask = gray == 0
    err = (synt_target - orig).abs()
    err = torch.where(err_mask, torch.zeros((), dtype=err.dtype), err)
    if reduce:
        err = err.mean(dim=(1, 2, 3, 4))
    return err


def photometric_loss_l2(synt_target, orig_target, reduce=True):
    """loss_util.py:29-48."""
    orig = orig_target.unsqueeze(1)
    gray = synt_target.mean(dim=-1, keepdim=True)
    err_mask = gray == 0
    err = (synt_target - orig) ** 2
    err = torch.where(err_mask, torch.zeros((), dtype=err.dtype), err)
    if reduce:
        err = err.mean(dim=(1, 2, 3, 4))
    return err


def photometric_loss_ssim(synt_target, orig_target, reduce=True):
    """loss_util.py:52-96."""
    N = synt_target.shape[1]
    x = orig_target.unsqueeze(1).expand(-1, N, -1, -1, -1)
    y = synt_target
    gray = y.mean(dim=-1, keepdim=True)
    err_mask = gray == 0
    mu_x = _avg_pool_same_3x3(x)
    mu_y = _avg_pool_same_3x3(y)
    sigma_x = _avg_pool_same_3x3(x ** 2) - mu_x ** 2
    sigma_y = _avg_pool_same_3x3(y ** 2) - mu_y ** 2
    sigma_xy = _avg_pool_same_3x3(x * y) - mu_x * mu_y
    ssim_n = (2 * mu_x * mu_y + SSIM_C1) * (2 * sigma_xy + SSIM_C2)
    ssim_d = (mu_x ** 2 + mu_y ** 2 + SSIM_C1) * (sigma_x + sigma_y + SSIM_C2)
    ssim = ssim_n / ssim_d
    ssim = torch.clamp((1 - ssim) / 2, 0, 1)
    ssim = torch.where(err_mask, torch.zeros((), dtype=ssim.dtype), ssim)
    if reduce:
        ssim = ssim.mean(dim=(1, 2, 3, 4))
    return ssim


_PHOTO = {"L1": photometric_loss_l1, "L2": photometric_loss_l2, "SSIM": photometric_loss_ssim}


def merge_multi_scale_losses(losses: Sequence[torch.Tensor], scale_weights) -> torch.Tensor:
    """losses.py:147-154: [S,B]^T @ [S,1] -> [B,1]."""
    sw = torch.as_tensor(scale_weights, dtype=losses[0].dtype).reshape(-1, 1)
    return torch.matmul(torch.stack(list(losses), dim=0).t(), sw)


def photometric_loss_multi_scale(method, synth_target_ms, target_ms, scale_weights):
    """losses.py:175-195 (plain mean over sources, inside the reduce_mean)."""
    fn = _PHOTO[method]
    return merge_multi_scale_losses([fn(s, t) for s, t in zip(synth_target_ms, target_ms)], scale_weights)


def smootheness_loss(disp, image, grad_factor=IMAGE_GRADIENT_FACTOR):
    """losses.py:409-440."""
    def gx(t):
        return t[:, :, :-1, :] - t[:, :, 1:, :]

    def gy(t):
        return t[:, :-1, :, :] - t[:, 1:, :, :]
    wx = torch.exp(-(gx(image) * grad_factor).abs().mean(dim=3, keepdim=True))
    wy = torch.exp(-(gy(image) * grad_factor).abs().mean(dim=3, keepdim=True))
    sx = 0.5 * (gx(disp) * wx).abs().mean(dim=(1, 2, 3))
    sy = 0.5 * (gy(disp) * wy).abs().mean(dim=(1, 2, 3))
    return sx + sy


def smootheness_loss_multi_scale(disp_ms, target_ms, scale_weights):
    """losses.py:386-407 (each scale divided by its scale factor, :401-402)."""
    orig_w = target_ms[0].shape[2]
    losses = []
    for disp, image in zip(disp_ms, target_ms):
        scale = orig_w / image.shape[2]
        losses.append(smootheness_loss(disp, image) / scale)
    return merge_multi_scale_losses(losses, scale_weights)


def monodepth2_loss_multi_scale(method, synth_target_ms, target, scale_weights):
    """losses.py:198-232: upsample each scale to full-res, min over sources."""
    fn = _PHOTO[method]
    Ho, Wo = target.shape[1:3]
    losses = []
    for synt in synth_target_ms:
        B, N, h, w, C = synt.shape
        up = resize_bilinear_tf(synt.reshape(B * N, h, w, C), (Ho, Wo)).reshape(B, N, Ho, Wo, C)
        per_pix = fn(up, target, False)
        losses.append(torch.amin(per_pix, dim=1).mean(dim=(1, 2, 3)))
    return merge_multi_scale_losses(losses, scale_weights)


# --------------------------------------------------------------------------
# TotalLoss  (reference losses.py:14-103)
# --------------------------------------------------------------------------
def append_data(features: Dict, predictions: Dict, suffix: str = "") -> Dict:
    """losses.py:57-103 (target frame is the LAST one of the snippet, :77-78)."""
    image5d = features["image5d" + suffix]
    source, target = image5d[:, :-1], image5d[:, -1]
    augm = {"source" + suffix: source, "target" + suffix: target}
    if "depth_ms" + suffix in predictions and "pose" + suffix in predictions:
        augm["target_ms" + suffix] = multi_scale_like_depth(target, predictions["depth_ms" + suffix])
        augm["synth_target_ms" + suffix] = synthesize_multi_scale(source, features["intrinsic" + suffix],
                                                                  predictions["depth_ms" + suffix],
                                                                  predictions["pose" + suffix])
    if "flow_ms" + suffix in predictions:
        augm["flow_target_ms" + suffix] = [resize_bilinear_tf(target, (f.shape[2], f.shape[3]))
                                           for f in predictions["flow_ms" + suffix]]
        augm["warped_target_ms" + suffix] = flow_warp_multi_scale(source, predictions["flow_ms" + suffix])
    return augm


def synthesize_stereo(features: Dict, predictions: Dict, augm: Dict) -> Dict:
    """losses.py:105-140: left target from the right frame with T_RL = inv(T_LR), right target from the left
    frame with T_LR; both through the rvec round trip of convert_pose.py:151-168 and -- as the reference
    does -- both with the LEFT intrinsics."""
    out = {}
    if "stereo_T_LR" not in features or "depth_ms" not in predictions:
        return out
    T_LR = features["stereo_T_LR"]
    pose_RL = pose_matr2rvec_batch(torch.linalg.inv(T_LR).unsqueeze(1))
    out["stereo_synth_ms"] = synthesize_multi_scale(augm["target_R"].unsqueeze(1), features["intrinsic"],
                                                    predictions["depth_ms"], pose_RL)
    pose_LR = pose_matr2rvec_batch(T_LR.unsqueeze(1))
    out["stereo_synth_ms_R"] = synthesize_multi_scale(augm["target"].unsqueeze(1), features["intrinsic"],
                                                      predictions["depth_ms_R"], pose_LR)
    return out


def stereo_depth_loss(method, augm: Dict, scale_weights) -> torch.Tensor:
    """losses.py:443-478: per-scale photometric loss of both stereo syntheses, summed, then merged."""
    fn = _PHOTO[method]
    left = [fn(s, t) for s, t in zip(augm["stereo_synth_ms"], augm["target_ms"])]
    right = [fn(s, t) for s, t in zip(augm["stereo_synth_ms_R"], augm["target_ms_R"])]
    return merge_multi_scale_losses([l + r for l, r in zip(left, right)], scale_weights)


def stereo_pose_loss(features: Dict, predictions: Dict) -> torch.Tensor:
    """losses.py:481-495: MSE over the 6 twist components of both directions, mean over numsrc -> [B]."""
    T = features["stereo_T_LR"].unsqueeze(1)
    lr_true, rl_true = pose_matr2rvec_batch(T), pose_matr2rvec_batch(torch.linalg.inv(T))
    loss = ((lr_true - predictions["pose_LR"]) ** 2).mean(dim=-1) + ((rl_true - predictions["pose_RL"]) ** 2).mean(dim=-1)
    return loss.mean(dim=1)


def moa_loss_multi_scale(method, temp_synth_ms, stereo_synth_ms, target, scale_weights):
    """losses.py:282-321: per-pixel min over the N temporal and the stereo synthesis, at full resolution."""
    fn = _PHOTO[method]
    Ho, Wo = target.shape[1:3]

    def up(x):
        B, N, h, w, C = x.shape
        return resize_bilinear_tf(x.reshape(B * N, h, w, C), (Ho, Wo)).reshape(B, N, Ho, Wo, C)
    losses = []
    for temp, stro in zip(temp_synth_ms, stereo_synth_ms):
        both = torch.cat([fn(up(temp), target, False), fn(up(stro), target, False)], dim=1)
        losses.append(torch.amin(both, dim=1).mean(dim=(1, 2, 3)))
    return merge_multi_scale_losses(losses, scale_weights)


def combined_loss_multi_scale(method, synth_target_ms, warped_target_ms, target, scale_weights):
    """losses.py:235-279 CombinedLossMultiScale: per pixel/channel/source, the static (depth+pose) loss of every
    scale counts only where it is smaller than the optical-flow loss of flow level 0; both compared at full
    resolution.  The mask is a constant of the graph (tf.cast of a comparison): no gradient reaches the flow."""
    fn = _PHOTO[method]
    Ho, Wo = target.shape[1:3]

    def up(x):
        B, N, h, w, C = x.shape
        return resize_bilinear_tf(x.reshape(B * N, h, w, C), (Ho, Wo)).reshape(B, N, Ho, Wo, C)
    flow_loss = fn(up(warped_target_ms[0]), target, False)
    losses = []
    for synt in synth_target_ms:
        static = fn(up(synt), target, False)
        mask = (static < flow_loss).to(static.dtype)
        losses.append((static * mask).mean(dim=(1, 2, 3, 4)))
    return merge_multi_scale_losses(losses, scale_weights)


def l2_regularizer(weights: Sequence[torch.Tensor], batch: int) -> torch.Tensor:
    """losses.py:522-534 L2Regularizer: sum_w tf.nn.l2_loss(w) = sum(w^2)/2, tiled to [batch]."""
    loss = sum((w ** 2).sum() / 2 for w in weights)
    return loss.reshape(1).repeat(batch)


def total_loss(predictions: Dict, features: Dict, loss_weights: Dict[str, float],
               scale_weights: Sequence[float], global_batch: Optional[int] = None,
               return_augm: bool = False, stereo: bool = False,
               weights_to_regularize: Optional[Sequence[torch.Tensor]] = None):
    """losses.py:26-55: per-type [B] loss -> sum/global_batch -> weighted sum.
    Returns (total, {name: unweighted mean}[, augm_data])."""
    augm = append_data(features, predictions)
    if stereo and "image5d_R" in features:             # losses.py:38-42
        augm.update(append_data(features, predictions, "_R"))
        augm.update(synthesize_stereo(features, predictions, augm))
    B = features["image5d"].shape[0]
    gb = B if global_batch is None else global_batch
    by_type, weighted = {}, []
    for name, w in loss_weights.items():
        if w == 0.0:
            continue                                   # loss_factory.py:41-43
        sfx = "_R" if name.endswith("_R") else ""
        base = name[:-2] if sfx else name
        if base in ("L1", "SSIM"):
            lb = photometric_loss_multi_scale(base, augm["synth_target_ms" + sfx], augm["target_ms" + sfx], scale_weights)
        elif base == "smoothe":
            # no disp_ms: what model_wrappers.py:47-48 derives from depth_ms (value 0, not the reference's NaN, at depth 0)
            disp = predictions.get("disp_ms" + sfx)
            if disp is None:
                disp = [torch.where(d > 0.00001, 1.0 / torch.where(d > 0.00001, d, torch.ones_like(d)), torch.zeros_like(d))
                        for d in predictions["depth_ms" + sfx]]
            lb = smootheness_loss_multi_scale(disp, augm["target_ms" + sfx], scale_weights)
        elif base in ("md2L1", "md2SSIM"):
            lb = monodepth2_loss_multi_scale(base[3:], augm["synth_target_ms" + sfx], augm["target" + sfx], scale_weights)
        elif base in ("moaL1", "moaSSIM"):
            # losses.py:293-295: the stereo synthesis is ALWAYS the left one ("stereo_synth_ms"), also for "_R"
            lb = moa_loss_multi_scale(base[3:], augm["synth_target_ms" + sfx], augm["stereo_synth_ms"],
                                      augm["target" + sfx], scale_weights)
        elif base in ("stereoL1", "stereoSSIM"):
            lb = stereo_depth_loss(base[6:], augm, scale_weights)
        elif base == "stereoPose":
            lb = stereo_pose_loss(features, predictions)
        elif base == "flowL2":
            lb = photometric_loss_multi_scale("L2", augm["warped_target_ms" + sfx], augm["flow_target_ms" + sfx],
                                              scale_weights)
        elif base in ("cmbL1", "cmbSSIM"):
            lb = combined_loss_multi_scale(base[3:], augm["synth_target_ms" + sfx], augm["warped_target_ms" + sfx],
                                           augm["target" + sfx], scale_weights)
        elif name == "flow_reg":
            lb = l2_regularizer(weights_to_regularize, B)
        else:
            raise ValueError(f"oracle: loss {name!r} is outside the hot path")
        mean = lb.sum() / gb                           # tf.nn.compute_average_loss
        by_type[name] = mean
        weighted.append(mean * w)
    total = torch.stack(weighted).sum()
    if return_augm:
        return total, by_type, augm
    return total, by_type


# --------------------------------------------------------------------------
# synthetic workload   -- shared by tests and bench
# --------------------------------------------------------------------------
def make_inputs(B: int, H: int, W: int, N: int = 4, n_scales: int = 4, seed: int = 20211,
                dtype=torch.float32, adversarial: bool = False) -> Tuple[Dict, Dict]:
    """Seeded synthetic snippet batch: smooth images in [-1,1], depth in [2,60] m
    with 1% exact zeros, small poses with |omega| >= 1e-3, pinhole K."""
    g = torch.Generator().manual_seed(seed)

    def U(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float64) * (hi - lo) + lo

    low = U(B * (N + 1), 3, max(H // 8, 2), max(W // 8, 2))
    img = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    img = (img + 0.1 * U(B * (N + 1), 3, H, W)).clamp(-1, 1)
    image5d = img.reshape(B, N + 1, 3, H, W).permute(0, 1, 3, 4, 2).contiguous()

    if adversarial:
        depth0 = U(B, 1, H, W, lo=1.0, hi=80.0)
    else:
        fld = U(B, 1, max(H // 16, 2), max(W // 16, 2))
        fld = F.interpolate(fld, size=(H, W), mode="bilinear", align_corners=False)
        depth0 = torch.exp(math.log(2.0) + (fld + 1) * 0.5 * (math.log(60.0) - math.log(2.0)))
    zero = U(B, 1, H, W, lo=0.0, hi=1.0) < 0.01
    depth0 = torch.where(zero, torch.zeros((), dtype=torch.float64), depth0)
    depth_ms = []
    for s in range(n_scales):
        sc = 2 ** s
        d = depth0[:, :, ::sc, ::sc]
        if s > 0:
            d = d * (1 + 0.02 * U(*d.shape))
        depth_ms.append(d.permute(0, 2, 3, 1).contiguous())

    tscale, rscale = (1.5, 0.3) if adversarial else (0.3, 0.03)
    t = U(B, N, 3) * tscale
    w = U(B, N, 3) * rscale
    nrm = w.norm(dim=-1, keepdim=True)
    w = torch.where(nrm < 1e-3, w + 2e-3, w)
    pose = torch.cat([t, w], dim=-1)

    f = W * U(B, lo=0.5, hi=0.7)
    K = torch.zeros(B, 3, 3, dtype=torch.float64)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = W / 2 + U(B, lo=-4, hi=4)
    K[:, 1, 2] = H / 2 + U(B, lo=-4, hi=4)
    K[:, 2, 2] = 1.0

    depth_ms = [d.to(dtype) for d in depth_ms]
    features = {"image5d": image5d.to(dtype), "intrinsic": K.to(dtype)}
    # NOTE: the reference's safe_reciprocal_number gives (1/0)*0 = NaN on the exact
    # zeros planted above (a net never emits them); the synthetic disparity is 0 there.
    disp_ms = [torch.where(d > 1e-5, 1.0 / torch.where(d > 1e-5, d, torch.ones_like(d)),
                           torch.zeros_like(d)) for d in depth_ms]
    predictions = {"depth_ms": depth_ms, "disp_ms": disp_ms,
                   "pose": pose.to(dtype)}
    return features, predictions


def make_stereo_inputs(B: int, H: int, W: int, N: int = 4, n_scales: int = 4, seed: int = 20211,
                       dtype=torch.float32) -> Tuple[Dict, Dict]:
    """make_inputs for a stereo rig: left + right snippets, the rig transform T_LR (baseline ~0.5 m along x, a
    slight toe-in) and the two rig-pose predictions the reference's StereoPoseLoss consumes."""
    fl, pl = make_inputs(B, H, W, N, n_scales, seed, dtype)
    fr, pr = make_inputs(B, H, W, N, n_scales, seed + 7919, dtype)
    g = torch.Generator().manual_seed(seed + 1)
    rig = torch.zeros(B, 1, 6, dtype=torch.float64)
    rig[:, 0, 0] = 0.5 + 0.05 * (torch.rand(B, generator=g, dtype=torch.float64) - 0.5)
    rig[:, 0, 1:3] = 0.02 * (torch.rand(B, 2, generator=g, dtype=torch.float64) - 0.5)
    rig[:, 0, 3:] = 0.02 * (torch.rand(B, 3, generator=g, dtype=torch.float64) - 0.5) + 0.004
    T_LR = pose_rvec2matr_batch(rig)[:, 0]
    features = {"image5d": fl["image5d"], "intrinsic": fl["intrinsic"], "image5d_R": fr["image5d"],
                "intrinsic_R": (fl["intrinsic"].double() * (1 + 0.01 * torch.rand(B, 1, 1, generator=g, dtype=torch.float64))
                                ).to(dtype).clone(),
                "stereo_T_LR": T_LR.to(dtype)}
    features["intrinsic_R"][:, 2, :] = fl["intrinsic"][:, 2, :]
    noise = lambda: 0.01 * (torch.rand(B, 1, 6, generator=g, dtype=torch.float64) - 0.5)
    predictions = dict(pl)
    predictions.update({"depth_ms_R": pr["depth_ms"], "disp_ms_R": pr["disp_ms"], "pose_R": pr["pose"],
                        "pose_LR": (rig + noise()).to(dtype),
                        "pose_RL": (pose_matr2rvec_batch(torch.linalg.inv(T_LR).unsqueeze(1)) + noise()).to(dtype)})
    return features, predictions


def make_flow(B: int, H: int, W: int, N: int = 4, n_scales: int = 4, seed: int = 20211, dtype=torch.float32,
              first_scale: int = 4, magnitude: float = 3.0) -> List[torch.Tensor]:
    """Seeded optical-flow pyramid shaped like PWC-Net's output (flow_net.py:44-48, :349-350): list of
    [B,N,H/s,W/s,2], s = first_scale * 2^k; a smooth field of a few pixels plus per-pixel noise, so that some
    samples leave the image (invalid) and some land exactly on integer coordinates (flow 0 at ~1 % of pixels)."""
    g = torch.Generator().manual_seed(seed + 4242)

    def U(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float64) * (hi - lo) + lo
    out = []
    for k in range(n_scales):
        s = first_scale * 2 ** k
        h, w = H // s, W // s
        low = U(B * N, 2, max(h // 4, 1), max(w // 4, 1))
        fld = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) * magnitude
        fld = fld + 0.2 * U(B * N, 2, h, w)
        zero = U(B * N, 1, h, w, lo=0.0, hi=1.0) < 0.01
        fld = torch.where(zero, torch.zeros((), dtype=torch.float64), fld)
        out.append(fld.reshape(B, N, 2, h, w).permute(0, 1, 3, 4, 2).contiguous().to(dtype))
    return out


_GRAD_KEYS = ("depth_ms", "disp_ms", "pose", "depth_ms_R", "disp_ms_R", "pose_R", "pose_LR", "pose_RL",
              "flow_ms", "flow_ms_R")


def stereo_loss_and_grads(features: Dict, predictions: Dict, loss_weights, scale_weights,
                          global_batch: Optional[int] = None, weights_to_regularize=None):
    """total_loss(stereo=True) + autograd w.r.t. every prediction.  Returns dict(total, by_type, grads{key})."""
    preds = {}
    for k in _GRAD_KEYS:
        if k not in predictions:
            continue
        v = predictions[k]
        preds[k] = ([t.detach().clone().requires_grad_(True) for t in v] if isinstance(v, (list, tuple))
                    else v.detach().clone().requires_grad_(True))
    wreg = None if weights_to_regularize is None else [w.detach().clone().requires_grad_(True) for w in weights_to_regularize]
    total, by_type, augm = total_loss(preds, features, loss_weights, scale_weights, global_batch, True, stereo=True,
                                      weights_to_regularize=wreg)
    total.backward()
    gz = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
    grads = {k: ([gz(t) for t in v] if isinstance(v, list) else gz(v)) for k, v in preds.items()}
    if wreg is not None:
        grads["weights_to_regularize"] = [gz(w) for w in wreg]
    return {"total": total.detach(), "by_type": {k: v.detach() for k, v in by_type.items()}, "grads": grads,
            "augm": {k: ([t.detach() for t in v] if isinstance(v, list) else v.detach()) for k, v in augm.items()}}


def loss_and_grads(features: Dict, predictions: Dict, loss_weights, scale_weights,
                   global_batch: Optional[int] = None, want_source_grad: bool = False):
    """Forward + autograd backward, the oracle for tape.gradient (train_val.py:85).
    Returns dict(total, by_type, synth_ms, d_depth_ms, d_disp_ms, d_pose[, d_source])."""
    depth = [d.detach().clone().requires_grad_(True) for d in predictions["depth_ms"]]
    disp = [d.detach().clone().requires_grad_(True) for d in predictions["disp_ms"]]
    pose = predictions["pose"].detach().clone().requires_grad_(True)
    # gradient w.r.t. the SOURCE frames only (what GatherNd->ScatterNd gives in TF);
    # the target frame is data: no gradient is consumed there.
    src = features["image5d"][:, :-1].detach().clone().requires_grad_(want_source_grad)
    img = torch.cat([src, features["image5d"][:, -1:].detach()], dim=1)
    feats = {"image5d": img, "intrinsic": features["intrinsic"]}
    preds = {"depth_ms": depth, "disp_ms": disp, "pose": pose}
    total, by_type, augm = total_loss(preds, feats, loss_weights, scale_weights, global_batch, True)
    total.backward()

    def gz(t):
        return torch.zeros_like(t) if t.grad is None else t.grad
    out = {
        "total": total.detach(), "by_type": {k: v.detach() for k, v in by_type.items()},
        "synth_ms": [s.detach() for s in augm["synth_target_ms"]],
        "target_ms": [t.detach() for t in augm["target_ms"]],
        "d_depth_ms": [gz(d) for d in depth], "d_disp_ms": [gz(d) for d in disp], "d_pose": gz(pose),
    }
    if want_source_grad:
        out["d_source"] = gz(src)
    return out
