"""The reference's own known-answer tests for the hot path, re-expressed against
the oracle (the reference's test scripts need TensorFlow + a display, and several
call APIs that no longer exist -- SURVEY.md section 4).  Each test cites the
reference test it restates."""
import numpy as np
import torch

from oracle import xpt_oracle as orc


def test_scale_intrinsic():
    # reference model/synthesize/test_synthesizing.py:149-163
    K = torch.tensor([[8, 0, 4], [0, 8, 4], [0, 0, 1]], dtype=torch.float32).expand(8, 3, 3)
    Ks = orc.scale_intrinsic(K, 2)
    assert np.allclose(Ks[:, :2].numpy(), K[:, :2].numpy() / 2)
    assert np.allclose(Ks[:, -1].numpy(), K[:, -1].numpy())


def test_pixel2cam_shape_and_values():
    # reference test_synthesizing.py:166-183 (shape only there; values added here)
    B, H, W = 8, 4, 4
    grid = orc.pixel_meshgrid(H, W, torch.float32)
    assert grid.shape == (3, H * W)
    assert grid[0, :W].tolist() == [0, 1, 2, 3] and grid[1, :W].tolist() == [0, 0, 0, 0]   # u fastest
    K = torch.tensor([[4, 0, H / 2], [0, 4, W / 2], [0, 0, 1]], dtype=torch.float32).expand(B, 3, 3)
    depth = torch.full((B, H, W, 1), 2.0)
    cam = orc.pixel2cam(grid, depth, K)
    assert cam.shape == (B, 4, H * W)
    assert np.allclose(cam[0, 2].numpy(), 2.0) and np.allclose(cam[0, 3].numpy(), 1.0)
    assert np.allclose(cam[0, 0, :W].numpy(), (np.arange(W) - H / 2) / 4 * 2)


def test_transform_to_source():
    # reference test_synthesizing.py:186-208: T = 2*I3 with t = 1 -> Y = 2X + 1
    B, P, N = 8, 6, 3
    coords = np.arange(1, 4 * P + 1).reshape(P, 4).T.astype(np.float32)
    coords[3] = 1
    coords = torch.tensor(np.tile(coords, (B, 1, 1)))
    T = np.identity(4) * 2
    T[:3, 3] = 1
    T[3, 3] = 1
    T = torch.tensor(np.tile(T, (B, N, 1, 1)), dtype=torch.float32)
    out = orc.transform_to_source(coords, T)
    assert out.shape == (B, N, 4, P)
    assert np.allclose(coords[2, :3].numpy() * 2 + 1, out[2, 1, :3].numpy())


def test_pixel_weighting():
    # reference test_synthesizing.py:211-255
    B, N, H, W = 8, 4, 5, 5
    rng = np.random.default_rng(0)
    pc = rng.uniform(0.1, 3.9, (B, N, 3, H * W))
    pc[:, :, :, 0] = -1.5
    pc[:, :, :, 1] = 7
    cu, cv = 0.2, 0.7
    pc[:, :, 0, 3] = 2 + cu
    pc[:, :, 1, 3] = 3 + cv
    pc[:, :, 2, :] = 1
    pc_t = torch.tensor(pc, dtype=torch.float32)
    fc = orc.neighbor_int_pixels(pc_t, H, W)
    assert np.allclose(np.floor(pc[:, :, 0, 2:]), fc[:, :, 0, 2:].numpy())
    assert np.allclose(np.ceil(pc[:, :, 1, 2:]), fc[:, :, 3, 2:].numpy())
    mask = orc.make_valid_mask(fc, None, B)
    w = orc.calc_neighbor_weights(pc_t, fc, mask).numpy()
    assert np.allclose(w[:, :, 0, 3], (1 - cu) * (1 - cv), atol=1e-6)
    assert np.allclose(w[:, :, 1, 3], (1 - cu) * cv, atol=1e-6)
    assert np.allclose(w[:, :, 2, 3], cu * (1 - cv), atol=1e-6)
    assert np.allclose(w[:, :, 3, 3], cu * cv, atol=1e-6)
    ws = w.sum(axis=2)
    assert (np.isclose(ws, 0, atol=1e-6) | np.isclose(ws, 1, atol=1e-6)).all()
    assert np.allclose(ws[:, :, 0], 0) and np.allclose(ws[:, :, 1], 0)        # out-of-image


def test_reconstruct_bilinear_interp():
    # reference test_synthesizing.py:258-301
    B, N, H, W = 8, 4, 5, 5
    pc = np.stack(np.meshgrid(np.arange(0, H), np.arange(0, W)), axis=0).reshape(1, 1, 2, 5, 5).astype(np.float32)
    u_add = 1.3
    pc[0, 0, 0] += u_add
    pc = torch.tensor(np.tile(pc, (B, N, 1, 1, 1)).reshape(B, N, 2, H * W))
    fc = orc.neighbor_int_pixels(pc, H, W)
    mask = orc.make_valid_mask(fc, None, B).numpy()
    exp = np.zeros((B, N, H, W))
    exp[:, :, :4, :3] = 1
    assert np.allclose(exp.reshape(B, N, 1, H * W), mask)
    image = np.meshgrid(np.arange(0, H), np.arange(0, W))[0].reshape(1, 1, H, W, 1)
    image = torch.tensor(np.tile(image, (B, N, 1, 1, 3)).astype(np.float32))
    depth = torch.ones(B, H, W, 1)
    recon = orc.bilinear_interpolation(image, pc, depth).numpy()
    expected = (image.numpy() + u_add) * exp.reshape(B, N, H, W, 1)
    assert np.allclose(recon, expected, atol=1e-5)


def test_average_pool_3d_interior():
    # reference model/loss_and_metric/losses.py:541-559
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 20, 20, 3, generator=g)
    mu = orc._avg_pool_same_3x3(x)
    assert np.isclose(x[0, 0, 10:13, 10:13, 1].mean().item(), mu[0, 0, 11, 11, 1].item(), atol=1e-6)
    # border divisors 4 / 6 (not pinned by the reference)
    assert np.isclose(x[0, 0, 0:2, 0:2, 0].mean().item(), mu[0, 0, 0, 0, 0].item(), atol=1e-6)
    assert np.isclose(x[0, 0, 0:2, 4:7, 0].mean().item(), mu[0, 0, 0, 5, 0].item(), atol=1e-6)


def test_pose_rvec2matr_batch():
    # reference utils/convert_pose.py:222-242
    g = torch.Generator().manual_seed(2)
    poses = torch.rand(8, 4, 6, generator=g) * 2 - 1
    T = orc.pose_rvec2matr_batch(poses).numpy()
    p0, m0 = poses[3, 2].numpy(), T[3, 2]
    assert np.allclose(p0[:3], m0[:3, 3])
    assert np.isclose(np.arccos((np.trace(m0[:3, :3]) - 1) / 2), np.linalg.norm(p0[3:]), atol=1e-5)
    assert np.allclose(m0[3], [0, 0, 0, 1])


def test_rotation_sign():
    # reference utils/tests.py:62-76: omega = (0,0,pi/3) -> [[c,s,0],[-s,c,0],[0,0,1]]
    a = np.pi / 3
    T = orc.pose_rvec2matr_batch(torch.tensor([[[0, 0, 0, 0, 0, a]]], dtype=torch.float64)).numpy()[0, 0]
    R = np.array([[np.cos(a), np.sin(a), 0], [-np.sin(a), np.cos(a), 0], [0, 0, 1]])
    assert np.allclose(T[:3, :3], R)


def test_zero_rotation_is_identity():
    # utils/convert_pose.py:65 (tf.where on |theta| < 1e-8), forward only
    T = orc.pose_rvec2matr_batch(torch.tensor([[[1.0, 2.0, 3.0, 0, 0, 0]]])).numpy()[0, 0]
    assert np.allclose(T[:3, :3], np.eye(3)) and np.allclose(T[:3, 3], [1, 2, 3])


def test_flow_warp_simple():
    # reference model/build_model/flow_net.py:204-237: constant flow (1.5, 3.5)
    # -> interior equals the 4-neighbour average
    B, N, H, W = 1, 1, 12, 14
    g = torch.Generator().manual_seed(3)
    img = torch.rand(B, N, H, W, 3, generator=g)
    flow = torch.zeros(B, N, H, W, 2)
    flow[..., 0], flow[..., 1] = 1.5, 3.5
    out = orc.flow_warp_multi_scale(img, [flow])[0]
    i, j = 6, 7
    exp = 0.25 * (img[0, 0, i - 4, j - 2] + img[0, 0, i - 3, j - 2] + img[0, 0, i - 4, j - 1] + img[0, 0, i - 3, j - 1])
    assert np.allclose(out[0, 0, i, j].numpy(), exp.numpy(), atol=1e-6)


def test_pyramid_is_centre_2x2_mean():
    # tf.image.resize bilinear half-pixel, integer factor s
    g = torch.Generator().manual_seed(4)
    img = torch.rand(1, 16, 24, 3, generator=g)
    for s in (2, 4, 8):
        out = orc.resize_bilinear_tf(img, (16 // s, 24 // s))
        a = s // 2 - 1
        exp = 0.25 * (img[:, a::s, a::s] + img[:, a + 1::s, a::s] + img[:, a::s, a + 1::s] + img[:, a + 1::s, a + 1::s])
        assert np.allclose(out.numpy(), exp.numpy(), atol=1e-6)
