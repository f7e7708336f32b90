"""Drop-in for the TF1 twins the TF2 layers call (reference file encodermap/encodermap_tf1/backmapping.py)."""
from __future__ import annotations

from math import pi

import torch

from .. import _ops


def chain_in_plane(lengths: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:97-119.  ``lengths`` is (1,n-1) or (b,n-1)."""
    return _ops.ChainInPlane.apply(lengths, angles)


def dihedrals_to_cartesian_tf(dihedrals: torch.Tensor, cartesian: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:164-195.  A rank-2 ``cartesian`` is shared
    by all frames (the reference tiles it)."""
    return _ops.DihedralsToCartesian.apply(dihedrals, cartesian, 0)


def dihedral_to_cartesian_tf_one_way(dihedrals: torch.Tensor, cartesian: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:198-214."""
    return _ops.DihedralsToCartesian.apply(dihedrals, cartesian, 1)


# ---- generation side, TF1 signatures: atoms are named, not indexed (reference encodermap_tf1/backmapping.py:256-318) ----
def _positions(atom_names, name):
    # the TF1 loops start at atom 1 (`for i in range(1, len(atom_names))`)
    return [i for i in range(1, len(atom_names)) if atom_names[i] == name]


def guess_sp2_atom(cartesians: torch.Tensor, atom_names, bond_partner: str, angle_to_previous: float, bond_length: float) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:256-281."""
    if cartesians.shape[1] != len(atom_names):
        raise AssertionError(f"cartesians.shape={tuple(cartesians.shape)} len(atom_names)={len(atom_names)}")
    return _ops.guess_sp2_raw(cartesians, _positions(atom_names, bond_partner), angle_to_previous, bond_length)


def guess_amide_H(cartesians: torch.Tensor, atom_names) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:284-285."""
    return guess_sp2_atom(cartesians, atom_names, "N", 123 / 180 * pi, 1.10)


def guess_amide_O(cartesians: torch.Tensor, atom_names) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:288-289."""
    return guess_sp2_atom(cartesians, atom_names, "C", 121 / 180 * pi, 1.24)


def merge_cartesians(central_cartesians: torch.Tensor, central_atom_names, H_cartesians: torch.Tensor, O_cartesians: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/encodermap_tf1/backmapping.py:292-318."""
    return _ops.merge_cartesians_raw(central_cartesians, _positions(central_atom_names, "N"), _positions(central_atom_names, "C"),
                                     H_cartesians, O_cartesians)
