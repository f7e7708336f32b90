"""TEST INFRASTRUCTURE -- a minimal stand-in for the ``tensorflow`` module, backed by torch tensors.

TensorFlow is not installed in this image nor on the GPU box, so ``encodermap_b200/tf_adapter.py`` (the
``tf.custom_gradient`` half of the drop-in boundary) could not be executed at all in round 1.
This module provides exactly the ``tf`` surface that the adapter and the reference's CALLERS of the hot path touch
(loss_functions.py:266-296, 335-369, 917-942; models/layers.py:204-215, 957-986, 1252-1267), with TensorFlow's
semantics for each symbol, so that those callers + ``tf_adapter.install()`` run end to end against libemk on a GPU.

* a "tf.Tensor" is a ``torch.Tensor`` (CPU or CUDA); gradients are taken with ``GradientTape`` below, which records
  through torch autograd -- ``tf.custom_gradient`` becomes a ``torch.autograd.Function`` whose backward calls the
  user's ``grad_fn``, i.e. the adapter's backward closures are what produces every gradient;
* ``tf.py_function(func, inp, Tout)`` calls ``func`` eagerly (what TF does in eager mode);
* ``tf.experimental.dlpack`` is torch's DLPack pair.

Install with ``fake_tf.install()`` (registers ``tensorflow`` and ``tensorflow.keras...`` in ``sys.modules``); never
imported by the product (tests/test_abi.py greps for that).
"""
from __future__ import annotations

import sys
import types
from typing import Any, Callable, Sequence

import numpy as np
import torch
import torch.utils.dlpack

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
bool_ = torch.bool


class Tensor(torch.Tensor):
    """torch tensor with the few tf.Tensor behaviours the reference's function bodies rely on: negative-step slices
    (``x[:, k::-1]``; torch refuses them), ``get_shape()`` and a ``numpy()`` that works on graph-attached tensors.
    torch propagates the subclass through every operation, so results stay ``Tensor``."""

    def get_shape(self):
        return tuple(self.shape)

    # tf.Tensor is an immutable value: `x *= s` REBINDS x to a new tensor (the reference scales its losses that way,
    # loss_functions.py:283, 929); torch would modify storage in place, which autograd forbids on custom-op outputs
    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __itruediv__(self, other):
        return self / other

    def numpy(self):
        return torch.Tensor.numpy(self.detach().cpu().as_subclass(torch.Tensor))

    def __getitem__(self, idx):
        items = idx if isinstance(idx, tuple) else (idx,)
        if not any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in items):
            return super().__getitem__(idx)
        x, out, dim = self, [], 0
        for i in items:
            if isinstance(i, slice) and i.step is not None and i.step < 0:
                pos = torch.tensor(list(range(*i.indices(x.shape[dim]))), dtype=torch.long, device=x.device)
                x = x.index_select(dim, pos)
                out.append(slice(None))
            else:
                out.append(i)
            if i is not None:
                dim += 1
        return torch.Tensor.__getitem__(x, tuple(out))


def _wrap(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t

_DEFAULT_DEVICE = ["cpu"]


def set_default_device(device) -> None:
    """Where ``convert_to_tensor`` places python / numpy data (TF: the GPU when one is visible)."""
    _DEFAULT_DEVICE[0] = device


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    return {"float32": torch.float32, "float64": torch.float64, "bool": torch.bool, "int32": torch.int32,
            np.float32: torch.float32, np.float64: torch.float64}[dtype]


def _stack_nested(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and len(x):
        items = [_stack_nested(i) for i in x]
        tensors = [i for i in items if isinstance(i, torch.Tensor)]
        if tensors:
            ref = tensors[0]
            return torch.stack([i if isinstance(i, torch.Tensor) else torch.as_tensor(np.asarray(j), dtype=ref.dtype, device=ref.device)
                                for i, j in zip(items, x)])
    return None


def convert_to_tensor(x, dtype=None, name=None):
    if isinstance(x, (list, tuple)):
        nested = _stack_nested(x)            # lists of tensors keep dtype, device and the autograd graph (tf packs them)
        if nested is not None:
            x = nested
    if isinstance(x, torch.Tensor):
        return _wrap(x if dtype is None else x.to(_dt(dtype)))
    a = np.asarray(x)
    if dtype is None and a.dtype == np.float64 and not isinstance(x, np.ndarray):
        dtype = torch.float32    # python floats / lists of floats become float32 in TF; numpy arrays keep their dtype
    return _wrap(torch.as_tensor(a, dtype=_dt(dtype), device=_DEFAULT_DEVICE[0]))


constant = convert_to_tensor


def is_tensor(x) -> bool:
    return isinstance(x, torch.Tensor)


is_numeric_tensor = is_tensor


def zeros(shp, dtype=None):
    shp = _shape_list(shp)
    return _wrap(torch.zeros(shp, dtype=_dt(dtype) or _WORK_DTYPE[0], device=_DEFAULT_DEVICE[0]))


def eye(n, dtype=None):
    return _wrap(torch.eye(int(n), dtype=_dt(dtype) or _WORK_DTYPE[0], device=_DEFAULT_DEVICE[0]))


_WORK_DTYPE = [torch.float32]


def set_work_dtype(dtype) -> None:
    """dtype of tf.zeros / tf.eye / tf.cast(.., float32) results: float32 as in TF, or float64 for the CPU comparison of the
    reference's function bodies with the float64 oracle (the reference hard-codes float32 in a few casts)."""
    _WORK_DTYPE[0] = dtype


def norm(x, axis=None, keepdims=False):
    return torch.sqrt(torch.sum(x * x, dim=axis, keepdim=keepdims))


def equal(a, b):
    return a == b


def boolean_mask(x, mask, axis=0):
    assert axis == 1 and np.ndim(mask) == 2   # tf.boolean_mask(x, mask, axis=1) with a rank-2 mask flattens dims (1, 2) row-major
    return x[:, torch.from_numpy(np.asarray(mask)).to(x.device)]


# ---- elementwise / shape ops (the subset the callers use) ---------------------------------------------------------
def _shape_list(shp):
    if isinstance(shp, torch.Tensor):
        shp = shp.tolist()
    if not isinstance(shp, (list, tuple)):
        shp = [shp]
    return [int(v) for v in shp]


def shape(x):
    return torch.tensor(list(x.shape), dtype=torch.int64)


def reshape(x, shp):
    return x.reshape(_shape_list(shp))


def concat(values: Sequence[Tensor], axis: int, name=None):
    return torch.cat(list(values), dim=axis)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def reduce_mean(x, axis=None, keepdims=False):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def square(x):
    return x * x


def sqrt(x):
    return torch.sqrt(x)


def abs(x):  # noqa: A001 - mirrors tf.abs
    return torch.abs(x)


def minimum(a, b):
    # forward value only matters here; TF's tie rule for the gradient is documented in oracle/em_oracle.py
    return torch.minimum(a, b if isinstance(b, torch.Tensor) else torch.as_tensor(b, dtype=a.dtype, device=a.device))


def maximum(a, b):
    return torch.maximum(a, b if isinstance(b, torch.Tensor) else torch.as_tensor(b, dtype=a.dtype, device=a.device))


def sin(x):
    return torch.sin(x)


def cos(x):
    return torch.cos(x)


def add(a, b):
    return a + b


def zeros_like(x):
    return torch.zeros_like(x)


def ones_like(x):
    return torch.ones_like(x)


def where(c, a, b):
    return torch.where(c, a, b)


def cast(x, dtype):
    d = _dt(dtype)
    return x.to(_WORK_DTYPE[0] if d == torch.float32 else d)


def broadcast_to(x, shp):
    return x.expand(_shape_list(shp))


def gather(params, indices, axis=0, batch_dims=0):
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long, device=params.device)
    return params.index_select(axis, idx)


def matmul(a, b):
    return torch.matmul(a, b)


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.T


def cond(pred, true_fn: Callable, false_fn: Callable, name=None):
    return true_fn() if bool(pred) else false_fn()


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def stack(values, axis=0):
    return torch.stack(list(values), dim=axis)


# ---- tf.custom_gradient / tf.py_function / GradientTape --------------------------------------------------------------
def custom_gradient(f: Callable) -> Callable:
    """``f(*args) -> (outputs, grad_fn)`` with ``grad_fn(*upstream) -> gradient(s) w.r.t. the tensor args``.

    TensorFlow's contract (tensorflow/python/ops/custom_gradient.py): inside ``f`` no gradient is recorded through the
    operations of the forward pass; the returned ``grad_fn`` IS the gradient.  Non-tensor arguments are passed through;
    ``grad_fn`` returns one gradient per positional argument (or a single tensor for a single argument)."""

    def wrapper(*args):
        tensor_pos = [i for i, a in enumerate(args) if isinstance(a, torch.Tensor)]

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *tensors):
                full = list(args)
                for pos, t in zip(tensor_pos, tensors):
                    full[pos] = t.detach()
                with torch.no_grad():
                    out, grad_fn = f(*full)
                ctx.grad_fn = grad_fn
                ctx.multi = isinstance(out, (tuple, list))
                return tuple(out) if ctx.multi else out

            @staticmethod
            def backward(ctx, *upstream):
                with torch.no_grad():
                    grads = ctx.grad_fn(*upstream)
                if not isinstance(grads, (tuple, list)):
                    grads = (grads,)
                grads = list(grads)
                if len(grads) != len(args):
                    raise ValueError(f"custom_gradient: grad_fn returned {len(grads)} gradients for {len(args)} arguments")
                return tuple(grads[pos] for pos in tensor_pos)

        return _Fn.apply(*[args[i] for i in tensor_pos])

    wrapper.__name__ = getattr(f, "__name__", "custom_gradient_op")
    return wrapper


def py_function(func: Callable, inp: Sequence[Any], Tout):
    out = func(*inp)
    if isinstance(Tout, (list, tuple)):
        return list(out)
    return out


class GradientTape:
    """``with tf.GradientTape() as tape: ...; tape.gradient(loss, variables)`` on top of torch autograd."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, t):
        if isinstance(t, torch.Tensor) and not t.requires_grad:
            t.requires_grad_(True)

    def gradient(self, target, sources):
        single = isinstance(sources, torch.Tensor)
        src = [sources] if single else list(sources)
        grads = torch.autograd.grad(target, src, allow_unused=True)
        return grads[0] if single else list(grads)


def Variable(initial_value, trainable=True, dtype=None, name=None):
    t = convert_to_tensor(initial_value, dtype=dtype).detach().clone()
    return _wrap(t.requires_grad_(bool(trainable)))


# ---- namespaces --------------------------------------------------------------------------------------------------------
class _Debugging:
    @staticmethod
    def assert_all_finite(x, message=""):
        if isinstance(x, torch.Tensor) and not bool(torch.isfinite(x).all()):
            raise FloatingPointError(message)   # TF raises InvalidArgumentError; a FloatingPointError is the torch adapter's twin
        return x

    @staticmethod
    def assert_rank(x, rank):
        assert x.dim() == rank

    @staticmethod
    def is_numeric_tensor(x):
        return isinstance(x, torch.Tensor)


class _Dlpack:
    @staticmethod
    def to_dlpack(t):
        return torch.utils.dlpack.to_dlpack(t.detach().contiguous())

    @staticmethod
    def from_dlpack(capsule):
        return _wrap(torch.utils.dlpack.from_dlpack(capsule))


class _Experimental:
    dlpack = _Dlpack()


class _Linalg:
    @staticmethod
    def diag_part(x):
        return torch.diagonal(x, dim1=-2, dim2=-1)


class _Math:
    @staticmethod
    def equal(a, b):
        r = a == b
        return r if isinstance(r, torch.Tensor) else torch.tensor(bool(r))

    @staticmethod
    def mod(a, b):
        if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
            return torch.remainder(a, b)
        return torch.tensor(a % b)

    @staticmethod
    def floormod(a, b):
        return _Math.mod(a, b)


class _KerasLayer:
    """tf.keras.layers.Layer: ``layer(x)`` dispatches to ``call`` (no build / no weights needed on this path)."""

    def __init__(self, *a, **kw):
        pass

    def __call__(self, *a, **kw):
        return self.call(*a, **kw)


class _Concatenate(_KerasLayer):
    def __init__(self, axis=-1, name=None):
        self.axis = axis

    def call(self, xs):
        return torch.cat(list(xs), dim=self.axis)


class _KerasBackend:
    @staticmethod
    def constant(value, dtype=None, shape=None, name=None):   # noqa: A002
        return torch.as_tensor(value, dtype=_dt(dtype))


debugging = _Debugging()
experimental = _Experimental()
linalg = _Linalg()
math = _Math()


def install() -> types.ModuleType:
    """Register this module as ``tensorflow`` (plus the keras sub-modules the reference imports)."""
    me = sys.modules[__name__]
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer, layers.Concatenate = _KerasLayer, _Concatenate
    backend = types.ModuleType("tensorflow.keras.backend")
    backend.constant = _KerasBackend.constant
    keras.layers, keras.backend = layers, backend
    me.keras = keras
    sys.modules["tensorflow"] = me
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = layers
    sys.modules["tensorflow.keras.backend"] = backend
    return me


def uninstall() -> None:
    for k in ("tensorflow", "tensorflow.keras", "tensorflow.keras.layers", "tensorflow.keras.backend"):
        if sys.modules.get(k) is not None and getattr(sys.modules[k], "__name__", "").startswith(("oracle.fake_tf", "tensorflow")):
            sys.modules.pop(k, None)
