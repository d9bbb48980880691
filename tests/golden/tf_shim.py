"""A minimal TensorFlow-API shim over torch-CPU, used ONLY to generate golden vectors.

TensorFlow 2.4.1 cannot be installed in the build container (no network), but the
reference is pure Python: with this shim registered as ``tensorflow`` the
reference's own, unmodified source files (``model/synthesize/*.py``,
``model/loss_and_metric/*.py``, ``utils/convert_pose.py``, ``utils/util_funcs.py``)
import and run.  What is thereby pinned is everything the reference's Python
decides: op order, slicing, shapes, tiling, masks, weights, aggregation.  What is
NOT pinned is TensorFlow's kernels themselves: each ``tf.*`` op below is restated
from the TF 2.4 op definition (deliberately with different torch primitives than
``oracle/xpt_oracle.py`` uses, so that the two are independent restatements).

Tensors are ``torch.Tensor`` with ``get_shape()`` patched on, so ``tape.gradient``
is emulated by ``torch.autograd``.  Set ``FLOAT`` to ``torch.float64`` before
importing reference modules to generate fp64 vectors.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

FLOAT = torch.float32


class _Shape(list):
    def as_list(self):
        return list(self)


def _get_shape(self):
    return _Shape(int(s) for s in self.shape)


torch.Tensor.get_shape = _get_shape          # generator process only


def _dt(dtype):
    if dtype is None:
        return None
    if dtype in ("float32", np.float32) or dtype is torch.float32:
        return FLOAT
    if dtype in ("int32", np.int32) or dtype is torch.int32:
        return torch.int64            # index arithmetic; value-identical
    if dtype is torch.bool or dtype is bool:
        return torch.bool
    if dtype is torch.float64:
        return torch.float64
    raise TypeError(f"tf_shim: dtype {dtype!r}")


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_dt(dtype))
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack([_t(v) for v in x], dim=0)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], (list, tuple)) and len(x[0]) \
            and isinstance(x[0][0], torch.Tensor):
        return torch.stack([_t(v) for v in x], dim=0)      # nested lists of tensors (tf auto-packing)
    a = np.asarray(x)
    if dtype is not None:
        return torch.as_tensor(a).to(_dt(dtype))
    if a.dtype.kind == "f":
        return torch.as_tensor(a).to(FLOAT)
    return torch.as_tensor(a)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


# ---- elementwise / shape ops ------------------------------------------------
def constant(value, dtype=None, **_):
    return _t(value, dtype)


convert_to_tensor = constant


def cast(x, dtype):
    return _t(x).to(_dt(dtype))


def reshape(x, shape, **_):
    return _t(x).reshape(tuple(int(s) for s in shape))


def expand_dims(x, axis, **_):
    return _t(x).unsqueeze(int(axis))


def squeeze(x, axis=None, **_):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(int(axis))


def tile(x, multiples, **_):
    return _t(x).repeat(*[int(m) for m in multiples])


def concat(values, axis, **_):
    return torch.cat([_t(v) for v in values], dim=int(axis))


def stack(values, axis=0, **_):
    return torch.stack([_t(v) for v in values], dim=int(axis))


def transpose(x, perm=None, **_):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*[int(p) for p in perm])


def slice_(x, begin, size, **_):
    x = _t(x)
    idx = []
    for b, s in zip(begin, size):
        idx.append(slice(int(b), None if s == -1 else int(b) + int(s)))
    return x[tuple(idx)]


def ones(shape, dtype=None, **_):
    return torch.ones(tuple(int(s) for s in shape), dtype=_dt(dtype) or FLOAT)


def zeros(shape, dtype=None, **_):
    return torch.zeros(tuple(int(s) for s in shape), dtype=_dt(dtype) or FLOAT)


def eye(n, **_):
    return torch.eye(int(n), dtype=FLOAT)


def range_(start, limit=None, delta=1, dtype=None, **_):
    if limit is None:
        start, limit = 0, start
    return torch.arange(start, limit, delta, dtype=_dt(dtype) or torch.int64)


def meshgrid(*args, indexing="xy", **_):
    ts = [_t(a) for a in args]
    if indexing == "xy" and len(ts) == 2:
        g1, g0 = torch.meshgrid(ts[1], ts[0], indexing="ij")
        return [g0, g1]
    return list(torch.meshgrid(*ts, indexing="ij"))


def floor(x, **_):
    return torch.floor(_t(x))


def clip_by_value(x, lo, hi, **_):
    # gradient passes for lo <= x <= hi (TF _ClipByValueGrad), same as torch.clamp
    return torch.clamp(_t(x), lo, hi)


def abs_(x, **_):
    return torch.abs(_t(x))


def square(x, **_):
    x = _t(x)
    return x * x


def exp(x, **_):
    return torch.exp(_t(x))


def sigmoid(x, **_):
    """tf.math.sigmoid: 1 / (1 + exp(-x))"""
    return 1.0 / (1.0 + torch.exp(-_t(x)))


def sin(x, **_):
    return torch.sin(_t(x))


def cos(x, **_):
    return torch.cos(_t(x))


def equal(a, b, **_):
    return _t(a) == (b if not isinstance(b, torch.Tensor) else b)


def not_equal(a, b, **_):
    return _t(a) != b


def logical_and(a, b, **_):
    return torch.logical_and(a, b)


def where(cond, x=None, y=None, **_):
    x, y = _t(x), _t(y)
    if x.dtype != y.dtype:
        x = x.to(y.dtype) if y.is_floating_point() else x
        y = y.to(x.dtype)
    return torch.where(cond, x, y)


def reduce_mean(x, axis=None, keepdims=False, **_):
    x = _t(x)
    ax = _axes(axis)
    if ax is None:
        return x.mean()
    return x.mean(dim=ax, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False, **_):
    x = _t(x)
    ax = _axes(axis)
    if ax is None:
        return x.sum()
    return x.sum(dim=ax, keepdim=keepdims)


def reduce_min(x, axis=None, keepdims=False, **_):
    # TF's reduce_min gradient is split equally among ties; so is torch.amin's
    x = _t(x)
    return torch.amin(x, dim=_axes(axis), keepdim=keepdims)


def norm(x, axis=None, keepdims=False, **_):
    x = _t(x)
    return torch.sqrt((x * x).sum(dim=_axes(axis), keepdim=keepdims))


def matmul(a, b, **_):
    return torch.matmul(_t(a), _t(b))


def tensordot(a, b, axes, **_):
    a, b = _t(a), _t(b)
    (ax_a,), (ax_b,) = axes
    return torch.tensordot(a, b, dims=([int(ax_a)], [int(ax_b)]))


def linalg_inv(x, **_):
    return torch.linalg.inv(_t(x))


def gather_nd(params, indices, batch_dims=0, **_):
    """tf.gather_nd with batch_dims=2 and index depth 2: params [B,N,H,W,C],
    indices [B,N,P,2] -> [B,N,P,C]."""
    params, indices = _t(params), _t(indices).to(torch.int64)
    assert batch_dims == 2 and indices.shape[-1] == 2
    B, N, P, _ = indices.shape
    bi = torch.arange(B).reshape(B, 1, 1).expand(B, N, P)
    ni = torch.arange(N).reshape(1, N, 1).expand(B, N, P)
    return params[bi, ni, indices[..., 0], indices[..., 1]]


def image_resize(images, size, method="bilinear", **_):
    """tf.image.resize, TF2 semantics (half_pixel_centers=True, antialias=False),
    following tensorflow/core/kernels/image/resize_bilinear_op.cc:
    in = (out + 0.5) * scale - 0.5; lower = max(floor(in), 0);
    upper = min(ceil(in), size-1); lerp = in - floor(in);
    top = tl + (tr - tl) * x_lerp; out = top + (bottom - top) * y_lerp."""
    x = _t(images)
    M, H, W, C = x.shape
    h, w = int(size[0]), int(size[1])
    if method == "nearest":
        # TF2 nearest with half_pixel_centers: floor((out + 0.5) * scale)
        yi = torch.clamp(torch.floor((torch.arange(h, dtype=torch.float64) + 0.5) * (H / h)).long(), max=H - 1)
        xi = torch.clamp(torch.floor((torch.arange(w, dtype=torch.float64) + 0.5) * (W / w)).long(), max=W - 1)
        return x[:, yi][:, :, xi]
    assert method == "bilinear"

    def weights(out_size, in_size):
        scale = in_size / out_size
        pos = (torch.arange(out_size, dtype=torch.float32) + 0.5) * np.float32(scale) - 0.5
        fl = torch.floor(pos)
        lower = torch.clamp(fl, min=0).long()
        upper = torch.clamp(torch.ceil(pos), max=in_size - 1).long()
        return lower, upper, (pos - fl).to(x.dtype)
    yl, yu, yw = weights(h, H)
    xl, xu, xw = weights(w, W)
    xw = xw.reshape(1, 1, w, 1)
    yw = yw.reshape(1, h, 1, 1)
    top_rows, bot_rows = x[:, yl], x[:, yu]
    tl, tr = top_rows[:, :, xl], top_rows[:, :, xu]
    bl, br = bot_rows[:, :, xl], bot_rows[:, :, xu]
    top = tl + (tr - tl) * xw
    bottom = bl + (br - bl) * xw
    return top + (bottom - top) * yw


class AveragePooling3D:
    """tf.keras.layers.AveragePooling3D; SAME padding divides by the number of
    in-bounds taps (TF AvgPool excludes padding from the count)."""

    def __init__(self, pool_size, strides=1, padding="SAME", **_):
        assert tuple(pool_size) == (1, 3, 3) and strides == 1 and padding == "SAME"

    def __call__(self, x):
        x = _t(x)                               # [B,N,H,W,C]
        B, N, H, W, C = x.shape
        pad = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))
        ones_ = torch.nn.functional.pad(torch.ones(H, W, dtype=x.dtype), (1, 1, 1, 1))
        acc = torch.zeros_like(x)
        cnt = torch.zeros(H, W, dtype=x.dtype)
        for dy in range(3):
            for dx in range(3):
                acc = acc + pad[:, :, dy:dy + H, dx:dx + W]
                cnt = cnt + ones_[dy:dy + H, dx:dx + W]
        return acc / cnt.reshape(1, 1, H, W, 1)


class Lambda:
    def __init__(self, fn, name=None, **_):
        self.fn = fn

    def __call__(self, inputs):
        return self.fn(inputs)


def compute_average_loss(per_example_loss, global_batch_size=None, **_):
    return _t(per_example_loss).sum() / global_batch_size


def random_uniform(shape, minval=0, maxval=1, **_):
    return torch.rand(tuple(shape), dtype=FLOAT) * (maxval - minval) + minval


def linalg_trace(x, **_):
    """tf.linalg.trace: sum of the main diagonal of the innermost matrices."""
    x = _t(x)
    return torch.einsum("...ii->...", x)


def acos(x, **_):
    return torch.arccos(_t(x))


def keras_mse(y_true, y_pred):
    """tf.keras.losses.MSE: mean of squared differences over the LAST axis."""
    d = _t(y_pred) - _t(y_true)
    return (d * d).sum(dim=-1) / d.shape[-1]


def l2_loss(t, **_):
    """tf.nn.l2_loss: sum(t ** 2) / 2 (TF op definition)."""
    t = _t(t)
    return (t * t).sum() / 2


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"tf_shim: {name} is outside the hot path")
    return f


def install(opts_overrides=None):
    """Register fake ``tensorflow``, ``tensorflow_addons``, ``quaternion`` and
    ``config`` modules in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.bool, tf.uint8 = torch.float32, torch.int32, torch.bool, torch.uint8
    tf.Tensor = torch.Tensor
    for name, fn in dict(
        constant=constant, convert_to_tensor=convert_to_tensor, cast=cast, reshape=reshape,
        expand_dims=expand_dims, squeeze=squeeze, tile=tile, concat=concat, stack=stack,
        transpose=transpose, slice=slice_, ones=ones, zeros=zeros, eye=eye, range=range_,
        meshgrid=meshgrid, floor=floor, clip_by_value=clip_by_value, abs=abs_, square=square,
        exp=exp, sin=sin, cos=cos, equal=equal, not_equal=not_equal, logical_and=logical_and,
        where=where, reduce_mean=reduce_mean, reduce_sum=reduce_sum, reduce_min=reduce_min,
        norm=norm, matmul=matmul, tensordot=tensordot, gather_nd=gather_nd,
    ).items():
        setattr(tf, name, fn)
    tf.linalg = types.SimpleNamespace(inv=linalg_inv, trace=linalg_trace)
    tf.math = types.SimpleNamespace(not_equal=not_equal, equal=equal, sin=sin, cos=cos,
                                    acos=acos, is_nan=torch.isnan, sigmoid=sigmoid,
                                    count_nonzero=_unsupported("math.count_nonzero"))
    tf.image = types.SimpleNamespace(resize=image_resize,
                                     convert_image_dtype=_unsupported("image.convert_image_dtype"))
    tf.nn = types.SimpleNamespace(compute_average_loss=compute_average_loss,
                                  l2_loss=l2_loss, avg_pool=_unsupported("nn.avg_pool"))
    tf.random = types.SimpleNamespace(uniform=random_uniform, normal=_unsupported("random.normal"))

    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Lambda = Lambda
    layers.AveragePooling3D = AveragePooling3D
    keras.layers = layers
    keras.losses = types.SimpleNamespace(MSE=keras_mse)
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = layers

    tfa = types.ModuleType("tensorflow_addons")
    sys.modules["tensorflow_addons"] = tfa
    sys.modules["quaternion"] = types.ModuleType("quaternion")

    # the git-ignored config.py of the reference: only the constants the path reads
    # (config-example.py:22,42-43,67,253)
    cfg = types.ModuleType("config")
    opts = types.SimpleNamespace(IMAGE_GRADIENT_FACTOR=4, ENABLE_SHAPE_DECOR=False, SNIPPET_LEN=5,
                                 STEREO=False, BATCH_SIZE=1, PER_REPLICA_BATCH=1)
    for k, v in (opts_overrides or {}).items():
        setattr(opts, k, v)
    cfg.opts = opts
    sys.modules["config"] = cfg
    return tf
