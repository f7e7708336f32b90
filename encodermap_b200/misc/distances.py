"""Drop-in for ``encodermap.misc.distances`` (reference file encodermap/misc/distances.py).

Same callables, same argument meaning; tensors are CUDA torch tensors and every result comes from
libemk.so.  Non-tensor inputs (lists, ndarrays) are moved to the current CUDA device, as the
reference converts them with ``tf.convert_to_tensor``."""
from __future__ import annotations

from math import pi
from typing import Callable

import numpy as np
import torch

from .. import _ops


def _as_cuda(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    t = torch.as_tensor(np.asarray(x), dtype=torch.float32)
    return t.cuda()


def sigmoid(sig: float, a: float, b: float) -> Callable:
    """Reference: encodermap/misc/distances.py:66-88.  The closure accepts a CUDA tensor (kernel) or a
    python number / ndarray (evaluated on the host, exactly as the reference's pure-python closure
    does -- its tests call it with floats, tests/test_pairwise_distances.py:191-193)."""

    def func(r):
        if isinstance(r, torch.Tensor):
            return _ops.Sigmoid.apply(_ops.f32c(_ops.require_cuda(r, "r")), sig, a, b)
        return 1 - (1 + (2 ** (a / b) - 1) * (r / sig) ** a) ** (-b / a)

    return func


def periodic_distance(a, b, periodicity: float = 2 * pi) -> torch.Tensor:
    """Reference: encodermap/misc/distances.py:113-141.  Operands broadcast like the reference's."""
    a, b = _as_cuda(a), _as_cuda(b)
    if a.shape != b.shape:
        a, b = torch.broadcast_tensors(a, b)
    return _ops.PeriodicDistance.apply(_ops.f32c(_ops.require_cuda(a, "a")), _ops.f32c(_ops.require_cuda(b, "b")), periodicity)


def pairwise_dist_periodic(positions, periodicity: float) -> torch.Tensor:
    """Reference: encodermap/misc/distances.py:144-176.  (n,d) -> (n,n), differentiable w.r.t. ``positions``
    with TensorFlow's autodiff conventions (in the models it only ever sees input data)."""
    positions = _as_cuda(positions)
    if positions.requires_grad:
        return _ops.PairwiseDistPeriodic.apply(positions, float(periodicity))
    return _ops.pairwise_dist_periodic_raw(positions, periodicity)


def pairwise_dist(positions, squared: bool = False, flat: bool = False) -> torch.Tensor:
    """Reference: encodermap/misc/distances.py:179-255.  Rank-2 input gains a leading batch axis;
    ``flat`` returns the strict upper triangle in row-major order."""
    return _ops.PairwiseDist.apply(_as_cuda(positions), bool(squared), bool(flat), None, None, None)
