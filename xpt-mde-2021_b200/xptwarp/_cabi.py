"""ctypes binding of libxptwarp.so (include/xptwarp.h).  Nothing else in the package
touches the shared library.  There is NO fallback: if the library is missing or no
sm_100 device is visible, the first call raises."""
from __future__ import annotations

import ctypes as C
import os

XPT_MAX_SCALES = 8
XPT_FLAG_UNFUSED = 1
XPT_FLAG_GRAPH = 2
XPT_FLAG_NO_PIPELINE = 4
XPT_FLAG_STRIP = 8
XPT_FLAG_ALLREDUCE = 16
XPT_FLAG_DEPTH_LOGIT = 32
XPT_FLAG_MIN_TILES = 64
XPT_OK, XPT_BAD_ARGUMENT, XPT_BAD_SHAPE, XPT_CUDA_ERROR, XPT_NO_DEVICE, XPT_OUT_OF_MEMORY = 0, -1, -2, -3, -4, -5
XPT_BAD_DTYPE, XPT_BAD_DEVICE, XPT_NOT_CONTIGUOUS, XPT_NCCL_ERROR = -6, -7, -8, -9
XPT_PHOTO_L1, XPT_PHOTO_L2, XPT_PHOTO_SSIM = 0, 1, 2

_FP = C.POINTER(C.c_float)
FloatPtrArray = _FP * XPT_MAX_SCALES


class XptConfig(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("num_src", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("num_scales", C.c_int32), ("scales", C.c_int32 * XPT_MAX_SCALES),
        ("scale_weights", C.c_float * XPT_MAX_SCALES),
        ("w_l1", C.c_float), ("w_ssim", C.c_float), ("w_smooth", C.c_float),
        ("img_grad_factor", C.c_float), ("global_batch", C.c_int32), ("device", C.c_int32),
        ("flags", C.c_uint32),
    ]


class XptFrames(C.Structure):
    _fields_ = [
        ("source", C.c_void_p), ("source_batch_stride", C.c_int64), ("source_frame_stride", C.c_int64),
        ("target", C.c_void_p), ("target_batch_stride", C.c_int64), ("intrinsic", C.c_void_p),
    ]


class XptLossOutputs(C.Structure):
    _fields_ = [
        ("losses", C.c_void_p), ("loss_batch", C.c_void_p),
        ("synth_ms", C.c_void_p * XPT_MAX_SCALES), ("mask_ms", C.c_void_p * XPT_MAX_SCALES),
        ("target_ms", C.c_void_p * XPT_MAX_SCALES), ("d_depth_ms", C.c_void_p * XPT_MAX_SCALES),
        ("d_disp_ms", C.c_void_p * XPT_MAX_SCALES), ("d_pose", C.c_void_p), ("d_source", C.c_void_p),
        ("grad_scale", C.c_float),
    ]


PtrArray = C.c_void_p * XPT_MAX_SCALES

# every symbol include/xptwarp.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "xpt_version": (C.c_int, []),
    "xpt_last_error": (C.c_char_p, []),
    "xpt_status_string": (C.c_char_p, [C.c_int]),
    "xpt_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(XptConfig)]),
    "xpt_destroy": (None, [C.c_void_p]),
    "xpt_get_config": (C.c_int, [C.c_void_p, C.POINTER(XptConfig)]),
    "xpt_scratch_bytes": (C.c_size_t, [C.c_void_p]),
    "xpt_pose_rvec2matr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xpt_pose_matr2rvec": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "xpt_stereo_pose_loss": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xpt_build_pyramids": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.c_void_p]),
    "xpt_synthesize": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.c_void_p,
                                 C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p]),
    "xpt_synthesize_backward": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.c_void_p,
                                          C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "xpt_photometric_loss": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p,
                                       C.c_void_p, C.POINTER(PtrArray), C.c_void_p]),
    "xpt_photometric_min_loss": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(PtrArray), C.POINTER(PtrArray),
                                           C.c_void_p]),
    "xpt_photometric_min_pair_loss": (C.c_int, [C.c_void_p, C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p, C.c_int64,
                                                C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.POINTER(PtrArray),
                                                C.POINTER(PtrArray), C.c_void_p]),
    "xpt_photometric_cmb_loss": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PtrArray), C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(PtrArray),
                                           C.c_void_p]),
    "xpt_photometric_cmb_pair_loss": (C.c_int, [C.c_void_p, C.POINTER(PtrArray), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.POINTER(PtrArray),
                                                C.c_void_p]),
    "xpt_flow_warp": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.POINTER(PtrArray),
                                C.POINTER(PtrArray), C.c_void_p]),
    "xpt_flow_warp_backward": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.POINTER(PtrArray),
                                         C.POINTER(PtrArray), C.c_void_p, C.c_void_p]),
    "xpt_l2_regularizer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "xpt_smoothness_loss": (C.c_int, [C.c_void_p, C.POINTER(PtrArray), C.POINTER(PtrArray), C.c_void_p,
                                      C.c_void_p, C.POINTER(PtrArray), C.c_void_p]),
    "xpt_total_loss": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.POINTER(PtrArray),
                                 C.c_void_p, C.POINTER(XptLossOutputs), C.c_void_p]),
    "xpt_total_loss_host": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.POINTER(PtrArray),
                                      C.c_void_p, C.POINTER(XptLossOutputs), C.c_void_p]),
    "xpt_total_loss_host_begin": (C.c_int, [C.c_void_p, C.POINTER(XptFrames), C.POINTER(PtrArray), C.POINTER(PtrArray),
                                      C.c_void_p, C.POINTER(XptLossOutputs), C.c_void_p]),
    "xpt_total_loss_host_end": (C.c_int, [C.c_void_p]),
    "xpt_profile_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "xpt_profile_select": (C.c_int, [C.c_void_p, C.c_int]),
    "xpt_profile_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "xpt_last_launch_count": (C.c_int, [C.c_void_p]),
    "xpt_geometry_slot_shared": (C.c_int, [C.c_void_p]),
    "xpt_check_dlpack": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "xpt_scale_tensors": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "xpt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "xpt_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "xpt_comm_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xpt_allreduce": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "xpt_comm_destroy": (C.c_int, [C.c_void_p]),
    "xpt_comm_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

# XPTWARP_LIB: explicit path of another BUILD of the same library (profiling ablations); never a fallback
LIB_PATH = os.environ.get("XPTWARP_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libxptwarp.so")
_lib = None


class XptError(RuntimeError):
    """A libxptwarp call returned a negative xpt_status."""

    def __init__(self, status, message):
        super().__init__(f"libxptwarp: {message} (status {status})")
        self.status = status


def lib():
    """Load libxptwarp.so once; fail loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C xpt-mde-2021_b200/csrc`.  xptwarp has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        L = lib()
        raise XptError(status, L.xpt_last_error().decode() or L.xpt_status_string(status).decode())


def ptr_array(ptrs):
    arr = PtrArray()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
