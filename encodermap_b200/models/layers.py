"""Drop-in for the hot layers of ``encodermap.models.layers`` (reference file
encodermap/models/layers.py: ``PeriodicInput`` :174-215, ``BackMapLayerWithSidechains`` :218-843, ``BackMapLayer`` :912-986,
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


class BackMapLayerWithSidechains(torch.nn.Module):
    """(central_distances, central_angles, central_dihedrals, side_distances, side_angles, side_dihedrals) -> Cartesian
    (b, 3 n_residues + n_side_atoms, 3): backbone first, then the side-chain atoms residue by residue.

    Reference semantics (layers.py:218-843): ``feature_description[-1]`` maps the 1-based residue number to its number of
    side-chain dihedrals; the atoms are laid out in a plane and every bond angle and dihedral is then set one after the other by
    rotating the atoms behind it.  One kernel per direction does the whole sequence per frame (``emk_sidechain_backmap(_bwd)``).
    Descriptions the reference's constructor cannot build (not exactly one end residue without side chain, ...) raise
    ``ValueError`` here."""

    def __init__(self, feature_description: Any) -> None:
        super().__init__()
        self.feature_description = feature_description
        info = feature_description[-1]
        n_residues = max(list(info.keys()))
        assert sorted(info.keys()) == list(range(1, n_residues + 1)), (
            f"Currently the `feature_indices[-1]` dict needs to contain monotonous "
            f"increasing keys. Starting from 1 {feature_description[-1].keys()=}"
        )
        self.counts = [int(info[k]) for k in range(1, n_residues + 1)]
        self._plans = {}
        self.n_sidechains = sum(v + 1 for v in self.counts if v > 0)
        self.n_atoms = 3 * n_residues + self.n_sidechains
        self._plan_for(None)     # validates the description on the host

    def _plan_for(self, device):
        plan = self._plans.get(device)
        if plan is None:
            plan = _ops.SidechainPlan(self.counts, device)
            self._plans[device] = plan
        return plan

    def get_config(self) -> dict:
        return {"feature_description": self.feature_description}

    # the plans hold library handles and device memory: they are rebuilt on demand, never copied or pickled
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_plans"] = {}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._plans = {}

    @classmethod
    def from_config(cls, config: dict) -> "BackMapLayerWithSidechains":
        fd = {int(k): {int(kk): vv for kk, vv in v.items()} for k, v in config.pop("feature_description").items()}
        return cls(feature_description=fd)

    def forward(self, inputs) -> torch.Tensor:
        inputs = tuple(inputs)
        return _ops.SidechainBackmap.apply(self._plan_for(inputs[0].device), *inputs)

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
        self.indices = None
        self._index_dev = {}
        if getattr(self.p, "reconstruct_sidechains", False):
            # the sliced backbone plus one atom per residue with a side chain (layers.py:1188-1208)
            assert hasattr(self.p, "sidechain_info"), (
                "The provided parameters ask for sidechains to be reconstructed, "
                "but don't contain a 'sidechain_info' attribute."
            )
            info = self.p.sidechain_info[-1]
            counts = [info[k] for k in sorted(info.keys())]
            self.indices = _ops.sidechain_pairwise_indices(counts, self.p.cartesian_pwd_start, self.p.cartesian_pwd_stop,
                                                           self.p.cartesian_pwd_step)

    def get_config(self) -> dict:
        return {"parameters": dict(vars(self.p)), "print_name": self.print_name, "trainable": False,
                "sidechain_info": getattr(self.p, "sidechain_info", None)}

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.indices is not None:
            index = self._index_dev.get(inputs.device)
            if index is None:
                if len(self.indices) and (self.indices.min() < 0 or self.indices.max() >= inputs.shape[1]):
                    # tf.gather on a GPU returns zeros for such an index (on a CPU it raises): refuse instead
                    raise IndexError(f"PairwiseDistances: atom index {int(self.indices.max())} outside the {inputs.shape[1]} atoms "
                                     f"of the input (layers.py:1196-1207 advances by the side-chain dihedral counts)")
                index = torch.as_tensor(self.indices, dtype=torch.int32, device=inputs.device)
                self._index_dev[inputs.device] = index
            return _ops.PairwiseDist.apply(_ops.GatherAtoms.apply(inputs, index), False, True, None, None, None)
        return _ops.PairwiseDist.apply(inputs, False, True, self.p.cartesian_pwd_start, self.p.cartesian_pwd_stop,
                                       self.p.cartesian_pwd_step)

    call = forward
