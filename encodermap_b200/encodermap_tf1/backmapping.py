"""Drop-in for the TF1 twins the TF2 layers call (reference file encodermap/encodermap_tf1/backmapping.py)."""
from __future__ import annotations

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
