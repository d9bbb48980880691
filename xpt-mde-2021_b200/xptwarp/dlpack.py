"""Zero-copy tensor exchange: any object that implements ``__dlpack__`` (torch,
TensorFlow via ``tf.experimental.dlpack.to_dlpack``, CuPy, JAX ...) is unwrapped to
a raw device pointer + shape + strides for the C-ABI.  The capsule is kept alive
for the duration of the call and never consumed, so the producer keeps ownership.

Layout of DLManagedTensor follows dlpack.h v0.8 (the struct is ABI-stable)."""
from __future__ import annotations

import ctypes as C

kDLCPU, kDLCUDA, kDLCUDAHost = 1, 2, 3
kDLFloat = 2


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_PyCapsule_IsValid = C.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = C.c_int
_PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]


class DeviceView:
    """Borrowed view of a device tensor: pointer (bytes offset applied), shape, strides in elements."""
    __slots__ = ("ptr", "shape", "strides", "device_type", "device_id", "_keepalive")

    def __init__(self, ptr, shape, strides, device_type, device_id, keepalive):
        self.ptr, self.shape, self.strides = ptr, tuple(shape), tuple(strides)
        self.device_type, self.device_id, self._keepalive = device_type, device_id, keepalive

    def is_dense_from(self, dim):
        """dims >= dim are C-contiguous."""
        expect = 1
        for d in range(len(self.shape) - 1, dim - 1, -1):
            if self.shape[d] != 1 and self.strides[d] != expect:
                return False
            expect *= self.shape[d]
        return True


def from_capsule(capsule, keepalive=None) -> DeviceView:
    if not _PyCapsule_IsValid(capsule, b"dltensor"):
        raise TypeError("expected an unconsumed 'dltensor' PyCapsule")
    mt = C.cast(_PyCapsule_GetPointer(capsule, b"dltensor"), C.POINTER(DLManagedTensor)).contents
    t = mt.dl_tensor
    if (t.dtype.code, t.dtype.bits, t.dtype.lanes) != (kDLFloat, 32, 1):
        raise TypeError(f"xptwarp needs float32 tensors, got DLPack dtype code={t.dtype.code} bits={t.dtype.bits}")
    shape = [t.shape[i] for i in range(t.ndim)]
    if t.strides:
        strides = [t.strides[i] for i in range(t.ndim)]
    else:
        strides, acc = [0] * t.ndim, 1
        for i in range(t.ndim - 1, -1, -1):
            strides[i] = acc
            acc *= shape[i]
    return DeviceView((t.data or 0) + t.byte_offset, shape, strides, t.device.device_type, t.device.device_id,
                      (capsule, keepalive))


def view_of(obj) -> DeviceView:
    """obj: anything with __dlpack__ (or an already-made capsule)."""
    if type(obj).__name__ == "PyCapsule":
        return from_capsule(obj)
    if not hasattr(obj, "__dlpack__"):
        raise TypeError(f"{type(obj).__name__} does not implement __dlpack__")
    return from_capsule(obj.__dlpack__(), keepalive=obj)
