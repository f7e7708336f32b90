"""Drop-in for the TF part of ``encodermap.misc.backmapping`` (reference file
encodermap/misc/backmapping.py:179-309, 1873-1912, 1950-1968)."""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib, _ops


def split_and_reverse_dihedrals(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference: encodermap/misc/backmapping.py:179-214.  Index lists come from libemk (bit-exact contract)."""
    n_atoms = int(x.shape[1]) + 3
    _, _, dl, dr = _lib.backmap_split_indices(n_atoms)
    dev = x.device
    return x[:, torch.from_numpy(dl.astype("int64")).to(dev)], x[:, torch.from_numpy(dr.astype("int64")).to(dev)]


def split_and_reverse_cartesians(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference: encodermap/misc/backmapping.py:217-256."""
    la, ra, _, _ = _lib.backmap_split_indices(int(x.shape[1]))
    dev = x.device
    return x[:, torch.from_numpy(la.astype("int64")).to(dev)], x[:, torch.from_numpy(ra.astype("int64")).to(dev)]


def dihedrals_to_cartesian_tf_layers(dihedrals: torch.Tensor, cartesians: torch.Tensor, left_iteration_counter: int,
                                     right_iteration_counter: int) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:259-309.  The iteration counters are implied by the
    shapes; they are checked against the reference's formula (models/models.py:661-671)."""
    n = int(dihedrals.shape[-1]) + 3
    if (left_iteration_counter, right_iteration_counter) != (n // 2 - 1, (n - 3) // 2):
        raise ValueError(f"iteration counters ({left_iteration_counter},{right_iteration_counter}) do not match "
                         f"{n} atoms: expected ({n // 2 - 1},{(n - 3) // 2})")
    return _ops.DihedralsToCartesian.apply(dihedrals, cartesians, 0)


def dihedral_to_cartesian_tf_one_way_layers(dihedrals: torch.Tensor, cartesian: torch.Tensor, n: int) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1873-1912 (``n`` must equal dihedrals.shape[-1])."""
    if n != int(dihedrals.shape[-1]):
        raise ValueError("n must equal dihedrals.shape[-1]")
    return _ops.DihedralsToCartesian.apply(dihedrals, cartesian, 1)


def rotation_matrix(axis_unit_vec: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1950-1968 (applied to row vectors on the right)."""
    return _ops.rotation_matrix_raw(axis_unit_vec, angle)
