"""Drop-in for the three hot layers of ``encodermap.models.layers`` (reference file
encodermap/models/layers.py: ``PeriodicInput`` :174-215, ``BackMapLayer`` :912-986,
``PairwiseDistances`` :1164-1267).  Constructor arguments and ``get_config`` keys follow the
reference; the layers are ``torch.nn.Module``s here (the TF adapter wraps the same entry points as
``tf.custom_gradient`` functions, see encodermap_b200/tf_adapter.py)."""
from __future__ import annotations

from typing import Any, Tuple

import torch

from .. import _ops
from ..parameters import ADCParameters, Parameters


class PeriodicInput(torch.nn.Module):
    """(rows, d) -> (rows, 2d) = [sin x, cos x] with x rescaled to radians when periodicity != 2 pi."""

    def __init__(self, parameters, print_name: str, trainable: bool = False, **kwargs: Any) -> None:
        super().__init__()
        self.p = parameters
        self.print_name = print_name

    def get_config(self) -> dict:
        return {"parameters": dict(vars(self.p)), "print_name": self.print_name, "trainable": False}

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return _ops.PeriodicInputFn.apply(inputs, self.p.periodicity)

    call = forward


class BackMapLayer(torch.nn.Module):
    """(distances (b,n-1), angles (b,n-2), dihedrals (b,n-3)) -> Cartesian (b,n,3).

    Reference semantics (layers.py:957-986): bond lengths are the BATCH MEAN of ``distances``; the chain
    is laid out in the plane and both halves are curled into 3-D by the dihedrals (+ pi).  One fused
    kernel does all of it; the backward is the exact VJP from force/torque prefix sums."""

    def __init__(self, left_split: int, right_split: int) -> None:
        super().__init__()
        self.left_split = left_split
        self.right_split = right_split

    @classmethod
    def from_config(cls, config: dict) -> "BackMapLayer":
        return cls(left_split=config.pop("left_split"), right_split=config.pop("right_split"))

    def get_config(self) -> dict:
        return {"left_split": self.left_split, "right_split": self.right_split}

    def forward(self, inputs: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]) -> torch.Tensor:
        distances, angles, dihedrals = inputs
        n = int(angles.shape[1]) + 2
        if (self.left_split, self.right_split) != (n // 2 - 1, (n - 3) // 2):
            raise ValueError(f"BackMapLayer(left_split={self.left_split}, right_split={self.right_split}) does not match "
                             f"{n} atoms: expected ({n // 2 - 1},{(n - 3) // 2}) (models/models.py:661-671)")
        lengths = mean_lengths(distances)
        return _ops.BackMap.apply(lengths, angles, dihedrals)

    call = forward


class _ColumnMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.rows = x.shape[0]
        return _ops.column_mean_raw(x)[None]

    @staticmethod
    def backward(ctx, g):
        return (g / ctx.rows).expand(ctx.rows, -1)


def mean_lengths(distances: torch.Tensor) -> torch.Tensor:
    """tf.expand_dims(tf.reduce_mean(distances, 0), 0) of the reference (layers.py:970) -> (1, n-1)."""
    return _ColumnMean.apply(distances)


def back_map(distances: torch.Tensor, angles: torch.Tensor, dihedrals: torch.Tensor) -> torch.Tensor:
    """Functional form of BackMapLayer."""
    return _ops.BackMap.apply(mean_lengths(distances), angles, dihedrals)


class PairwiseDistances(torch.nn.Module):
    """inputs[:, start:stop:step] -> flat upper-triangle pairwise distances (b, n_sel (n_sel-1)/2)."""

    def __init__(self, parameters, print_name: str, trainable: bool = False, **kwargs: Any) -> None:
        super().__init__()
        self.p = parameters
        self.print_name = print_name
        if getattr(self.p, "reconstruct_sidechains", False):
            raise NotImplementedError("PairwiseDistances with reconstruct_sidechains=True (gathered side-chain atoms, "
                                      "layers.py:1190-1208) is outside the hot path built here")

    def get_config(self) -> dict:
        return {"parameters": dict(vars(self.p)), "print_name": self.print_name, "trainable": False,
                "sidechain_info": getattr(self.p, "sidechain_info", None)}

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return _ops.PairwiseDist.apply(inputs, False, True, self.p.cartesian_pwd_start, self.p.cartesian_pwd_stop,
                                       self.p.cartesian_pwd_step)

    call = forward
