"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/xptwarp.h
declares, the ctypes structs match the header, DLPack unwrapping is zero-copy, the host logic
mirrors the reference's factory/validation behaviour, and the product path refuses to run without
a CUDA device (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xptwarp.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from xptwarp import _cabi
    return _cabi


def test_library_exports_every_declared_symbol(built):
    text = open(HEADER).read()
    declared = set(re.findall(r"XPT_API\s+[\w\s\*]+?\b(xpt_\w+)\s*\(", text))
    assert len(declared) >= 16
    assert declared == set(built.SYMBOLS), declared ^ set(built.SYMBOLS)
    lib = built.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.xpt_version() == int(re.search(r"#define XPT_VERSION (\d+)", text).group(1))
    assert lib.xpt_status_string(-2) == b"XPT_BAD_SHAPE"


def test_struct_layouts_match_the_header(built, tmp_path):
    """sizeof/offsetof of the ctypes mirrors equal the C compiler's view of include/xptwarp.h."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "xptwarp.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(xpt_config), offsetof(xpt_config, scale_weights),'
                   'offsetof(xpt_config, flags), sizeof(xpt_frames), sizeof(xpt_loss_outputs),'
                   'offsetof(xpt_loss_outputs, d_pose), offsetof(xpt_loss_outputs, grad_scale));return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals == [C.sizeof(built.XptConfig), built.XptConfig.scale_weights.offset, built.XptConfig.flags.offset,
                    C.sizeof(built.XptFrames), C.sizeof(built.XptLossOutputs), built.XptLossOutputs.d_pose.offset,
                    built.XptLossOutputs.grad_scale.offset]


def test_no_cpu_fallback(built):
    """Without a CUDA device the library refuses (XPT_NO_DEVICE); with CPU tensors the host raises."""
    import xptwarp
    if not torch.cuda.is_available():
        with pytest.raises(built.XptError) as e:
            xptwarp.get_plan(0, 2, 4, 32, 64, [1, 2, 4, 8])
        assert e.value.status == -4
    img = torch.zeros(1, 2, 16, 16, 3)
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.SynthesizeMultiScale()(img, torch.eye(3)[None], [torch.ones(1, 16, 16, 1)], torch.zeros(1, 2, 6))


def test_create_rejects_bad_shapes(built):
    lib = built.lib()
    cfg = built.XptConfig()
    cfg.batch, cfg.num_src, cfg.height, cfg.width, cfg.num_scales = 1, 4, 30, 64, 2
    cfg.scales[0], cfg.scales[1] = 1, 4                   # 30 % 4 != 0
    h = C.c_void_p()
    assert lib.xpt_create(C.byref(h), C.byref(cfg)) == -2
    assert b"does not divide" in lib.xpt_last_error()
    cfg.num_src = 99
    assert lib.xpt_create(C.byref(h), C.byref(cfg)) == -2
    assert lib.xpt_create(None, C.byref(cfg)) == -1


def test_dlpack_view_is_zero_copy():
    from xptwarp.dlpack import kDLCPU, view_of
    t = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    v = view_of(t[:, 1:])                     # strided view with an offset
    assert v.ptr == t[:, 1:].data_ptr() and v.shape == (2, 2, 4) and v.strides == (12, 4, 1)
    assert v.device_type == kDLCPU and v.is_dense_from(1) and not v.is_dense_from(0)
    with pytest.raises(TypeError):
        view_of(torch.zeros(3, dtype=torch.float64))
    with pytest.raises(TypeError):
        view_of(object())


def test_loss_factory_mirrors_reference_filtering():
    import xptwarp
    from xptwarp.loss_factory import check_loss_dependency
    weights = {"L1": 0.5, "L1_R": 0.5, "SSIM": 0.5, "SSIM_R": 0.5, "smoothe": 1.0, "smoothe_R": 1.0,
               "stereoL1": 0.01, "stereoSSIM": 0.01, "stereoPose": 1.0, "md2L1": 0.0}
    tl = xptwarp.loss_factory({"image": 1, "intrinsic": 1}, weights, np.array([1, 1, 1, 1.0]), batch_size=4)
    assert list(tl.loss_objects) == ["L1", "SSIM", "smoothe"] and tl.batch_size == 4
    assert tl.loss_weights == {"L1": 0.5, "SSIM": 0.5, "smoothe": 1.0}
    assert check_loss_dependency("stereoL1", {"image", "intrinsic", "image_R", "intrinsic_R", "stereo_T_LR"})
    assert not check_loss_dependency("L1_R", {"image", "intrinsic"})
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.losses.PhotometricLossMultiScale("L3", None)
    # the optical-flow rows of the pool (loss_factory.py:27-29): flowL2 is the "L2" photometric term; flow_reg
    # without weights_to_regularize refuses to run
    tl2 = xptwarp.loss_factory({"image": 1, "intrinsic": 1}, {"flowL2": 1.0, "md2L1": 0.5, "flow_reg": 4e-7, "cmbSSIM": 0.5},
                               np.ones(4))
    assert isinstance(tl2.loss_objects["md2L1"], xptwarp.MonoDepth2LossMultiScale)
    assert isinstance(tl2.loss_objects["flowL2"], xptwarp.FlowWarpLossMultiScale) and tl2.loss_objects["flowL2"].method == "L2"
    assert isinstance(tl2.loss_objects["cmbSSIM"], xptwarp.CombinedLossMultiScale)
    with pytest.raises(xptwarp.WrongInputException):
        tl2.loss_objects["flow_reg"]({"image5d": None}, None, None)
    # the whole stereo pool of config-example.py:76-121 is served on a rig dataset
    rig = {"image": 1, "intrinsic": 1, "image_R": 1, "intrinsic_R": 1, "stereo_T_LR": 1}
    tl3 = xptwarp.loss_factory(rig, dict(weights, moaL1=5.0, moaSSIM_R=0.5), np.ones(4), stereo=True, batch_size=4)
    assert list(tl3.loss_objects) == ["L1", "L1_R", "SSIM", "SSIM_R", "smoothe", "smoothe_R", "stereoL1", "stereoSSIM",
                                      "stereoPose", "moaL1", "moaSSIM_R"]
    assert isinstance(tl3.loss_objects["stereoPose"], xptwarp.StereoPoseLoss) and tl3.stereo


def test_infer_scales_and_shard_bounds():
    from xptwarp.distributed import shard_bounds
    from xptwarp.engine import WrongInputException, infer_scales
    assert infer_scales(128, [torch.zeros(1, 128 // s, 4, 1) for s in (1, 2, 4, 8)]) == [1, 2, 4, 8]
    with pytest.raises(WrongInputException):
        infer_scales(128, [torch.zeros(1, 50, 4, 1)])
    spans = [shard_bounds(64, r, 8) for r in range(8)]
    assert spans[0] == (0, 8) and spans[-1] == (56, 64)
    spans = [shard_bounds(10, r, 4) for r in range(4)]
    assert [hi - lo for lo, hi in spans] == [3, 3, 2, 2] and spans[-1][1] == 10
    with pytest.raises(ValueError):
        shard_bounds(8, 4, 4)


def test_flow_call_surface_rejects_bad_inputs():
    """FlowWarpMultiScale / multi_scale_like_flow / the flow losses refuse CPU tensors (no CPU path), wrong ranks and
    flow sizes that do not divide the image -- with the reference's WrongInputException, before any kernel runs."""
    import xptwarp
    from xptwarp.flow_warping import infer_flow_scales
    src = torch.zeros(2, 3, 32, 64, 3)
    flows = [torch.zeros(2, 3, 32 // s, 64 // s, 2) for s in (4, 8)]
    assert infer_flow_scales(32, flows) == [4, 8]
    with pytest.raises(xptwarp.WrongInputException):
        infer_flow_scales(32, [torch.zeros(2, 3, 5, 8, 2)])
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.FlowWarpMultiScale()(src, flows)                          # CPU tensors
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.FlowWarpMultiScale()(src[:, :, :, :, :2], flows)          # not RGB
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.multi_scale_like_flow(src[:, 0], flows)                   # CPU tensor
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.L2Regularizer(None)({"image5d": src}, None, None)          # no weights given
    with pytest.raises(xptwarp.WrongInputException):
        xptwarp.L2Regularizer([torch.zeros(3)])({"image5d": src}, None, None)   # CPU weights


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "xpt-mde-2021_b200"))
    from oracle import xpt_oracle as orc
    from xptwarp.distributed import allreduce_gradient_buckets, allreduce_losses, shard_batch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    feats, preds = orc.make_inputs(4, 16, 24, N=2, seed=77)
    f, p = shard_batch(feats, preds, rank, world)
    # per-rank losses normalised by the GLOBAL batch (the oracle stands in for the kernels on CPU)
    total, by_type = orc.total_loss(p, f, orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1, global_batch=4)
    vec = torch.stack([total, by_type["L1"], by_type["SSIM"], by_type["smoothe"]]).detach()
    allreduce_losses(vec)
    bucket = torch.full((1000,), float(rank + 1))
    for w in allreduce_gradient_buckets([bucket]):
        w.wait()
    q.put((rank, vec.tolist(), float(bucket[0])))
    dist.destroy_process_group()


def test_sharded_losses_allreduce_to_the_global_batch_mean_gloo():
    """world_size 2 over gloo: shard + all-reduce(sum) of per-rank losses == single-process loss."""
    import torch.multiprocessing as mp
    from oracle import xpt_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    feats, preds = orc.make_inputs(4, 16, 24, N=2, seed=77)
    total, by_type = orc.total_loss(preds, feats, orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1, global_batch=4)
    want = [float(total), float(by_type["L1"]), float(by_type["SSIM"]), float(by_type["smoothe"])]
    for rank, vec, b0 in res:
        assert np.allclose(vec, want, rtol=1e-5), (rank, vec, want)
        assert b0 == 3.0


def test_check_dlpack_status_codes(built):
    """xpt_check_dlpack: the dtype / device / contiguity codes of SURVEY 8b live on the C side and need no GPU."""
    from xptwarp import dlpack
    lib = built.lib()

    def managed(t):
        cap = t.__dlpack__()
        ptr = dlpack._PyCapsule_GetPointer(cap, b"dltensor")
        return cap, ptr
    cap, ptr = managed(torch.zeros(2, 3, 4, dtype=torch.float64))
    assert lib.xpt_check_dlpack(ptr, 0, 0) == built.XPT_BAD_DTYPE
    assert b"float32" in lib.xpt_last_error()
    cap, ptr = managed(torch.zeros(2, 3, 4, dtype=torch.float32))
    assert lib.xpt_check_dlpack(ptr, 0, 0) == built.XPT_BAD_DEVICE          # a CPU tensor: no CPU path
    assert lib.xpt_check_dlpack(None, 0, 0) == built.XPT_BAD_ARGUMENT
    assert lib.xpt_status_string(built.XPT_NOT_CONTIGUOUS) == b"XPT_NOT_CONTIGUOUS"
    assert lib.xpt_status_string(built.XPT_NCCL_ERROR) == b"XPT_NCCL_ERROR"
    del cap


def test_strided_inputs_are_refused_not_copied(built):
    """engine._dense / _frame_view raise on non-dense inputs instead of calling .contiguous() behind the caller"""
    from xptwarp import engine
    t = torch.zeros(4, 6, 8)[:, ::2]
    for fn in (lambda: engine._dense(t, "x"), lambda: engine._frame_view(torch.zeros(2, 8, 8, 6)[..., ::2], "img", 1)):
        with pytest.raises(engine.WrongInputException):
            fn()


def test_fused_path_is_gated_on_the_loss_objects(built):
    """TotalLoss takes the fused launch only when the object under each name is the stock one."""
    import xptwarp
    from xptwarp import losses as lm
    sw = np.ones((4, 1), dtype=np.float32)
    stock = xptwarp.loss_factory({"image": 1, "intrinsic": 1}, {"L1": 0.5, "SSIM": 0.5, "smoothe": 1.0}, sw, batch_size=2)
    preds = {"depth_ms": [None], "pose": None}
    assert stock._fused_ok(preds, {})
    odd = lm.TotalLoss({"L1": lm.MonoDepth2LossMultiScale("SSIM", sw)}, {"L1": 1.0}, False, 2)
    assert not odd._fused_ok(preds, {})
    swapped = lm.TotalLoss({"L1": lm.PhotometricLossMultiScale("SSIM", sw)}, {"L1": 1.0}, False, 2)
    assert not swapped._fused_ok(preds, {})
    wrong_eye = lm.TotalLoss({"L1_R": lm.PhotometricLossMultiScale("L1", sw)}, {"L1_R": 1.0}, True, 2)
    assert not wrong_eye._fused_ok(preds, {})


def test_loss_pairs_share_a_launch_only_when_stock(built):
    """TotalLoss._min_pairs: (moa|md2|cmb)L1 + ...SSIM of ONE eye, stock objects with equal scale weights -> one pair launch
    (xpt_photometric_min_pair_loss / xpt_photometric_cmb_pair_loss); anything else keeps the reference's loop."""
    import xptwarp
    rig = {"image": 1, "intrinsic": 1, "image_R": 1, "intrinsic_R": 1, "stereo_T_LR": 1}
    lw = {"moaL1": 8.5, "moaL1_R": 8.5, "moaSSIM": 0.15, "moaSSIM_R": 0.15, "md2L1": 1.0, "cmbL1_R": 1.0, "cmbSSIM_R": 0.5,
          "smoothe": 1.0, "stereoPose": 1.0}
    tl = xptwarp.loss_factory(rig, lw, np.ones(4), stereo=True, batch_size=4)
    pairs = tl._min_pairs()
    assert pairs == {"moaL1": "moaSSIM", "moaSSIM": None, "moaL1_R": "moaSSIM_R", "moaSSIM_R": None,
                     "cmbL1_R": "cmbSSIM_R", "cmbSSIM_R": None}           # md2L1 has no SSIM partner
    # different scale weights, a swapped method, a subclass or a foreign object under the name: no pair
    tl.loss_objects["moaSSIM"].scale_weights = np.array([1, 2, 3, 4.0])
    assert "moaL1" not in tl._min_pairs() and "moaL1_R" in tl._min_pairs()
    tl.loss_objects["moaSSIM_R"] = xptwarp.MonoDepth2LossMultiScale("SSIM", np.ones(4), "_R")
    assert "moaL1_R" not in tl._min_pairs()
    tl.loss_objects["cmbL1_R"] = xptwarp.CombinedLossMultiScale("SSIM", np.ones(4), "_R")
    assert "cmbL1_R" not in tl._min_pairs()


def test_round2_entry_points_refuse_null_arguments(built):
    """The round-2 entry points validate before they touch CUDA: a NULL ctx (all a machine without a GPU can offer) is
    XPT_BAD_ARGUMENT with a message, never a crash."""
    lib = built.lib()
    assert lib.xpt_total_loss_host_end(None) == built.XPT_BAD_ARGUMENT
    assert b"ctx is NULL" in lib.xpt_last_error()
    fr, o = built.XptFrames(), built.XptLossOutputs()
    empty = built.ptr_array([0])
    assert lib.xpt_total_loss_host_begin(None, C.byref(fr), C.byref(empty), None, None, C.byref(o), None) == built.XPT_BAD_ARGUMENT
    assert lib.xpt_photometric_min_pair_loss(None, C.byref(empty), None, None, 0, None, None, 1.0, 1.0, None, None,
                                             None) == built.XPT_BAD_ARGUMENT
    assert lib.xpt_photometric_cmb_pair_loss(None, C.byref(empty), None, 4, 4, None, 0, None, None, 1.0, 1.0, None,
                                             None) == built.XPT_BAD_ARGUMENT
    assert lib.xpt_comm_status(None, None, None) == built.XPT_BAD_ARGUMENT
