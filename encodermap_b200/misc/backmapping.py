"""Drop-in for the TF part of ``encodermap.misc.backmapping`` (reference file
encodermap/misc/backmapping.py:179-309, 1873-1912, 1920-1990)."""
from __future__ import annotations

from math import pi
from typing import Sequence, Tuple

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


# ---- generation side: amide H / carbonyl O guessed from the backbone (forward only, as on the reference's generate path) ----
AMIDE_H = (123 / 180 * pi, 1.10)   # angle to the previous bond, bond length: misc/backmapping.py:1943-1944
AMIDE_O = (121 / 180 * pi, 1.24)   # misc/backmapping.py:1946-1947


def guess_sp2_atom(cartesians: torch.Tensor, indices: Sequence[int], angle_to_previous: float, bond_length: float) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1920-1941.  One atom per entry of ``indices`` (the centre atom), in the plane
    of its two neighbours, ``angle_to_previous`` away from the bond to the previous atom."""
    return _ops.guess_sp2_raw(cartesians, indices, angle_to_previous, bond_length)


def guess_amide_H(cartesians: torch.Tensor, N_indices: Sequence[int]) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1943-1944 (the first nitrogen gets no hydrogen: ``N_indices[1::]``)."""
    return guess_sp2_atom(cartesians, list(N_indices)[1::], *AMIDE_H)


def guess_amide_O(cartesians: torch.Tensor, C_indices: Sequence[int]) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1946-1947."""
    return guess_sp2_atom(cartesians, list(C_indices), *AMIDE_O)


def merge_cartesians(central_cartesians: torch.Tensor, N_indices: Sequence[int], O_indices: Sequence[int], H_cartesians: torch.Tensor,
                     O_cartesians: torch.Tensor) -> torch.Tensor:
    """Reference: encodermap/misc/backmapping.py:1970-1990: after atom i >= 1 comes the next H if ``i in N_indices[1::]``,
    else the next O if ``i in O_indices``; the reference's closing assert on the atom count is EmkError(EMK_E_SHAPE) here."""
    return _ops.merge_cartesians_raw(central_cartesians, list(N_indices)[1::], list(O_indices), H_cartesians, O_cartesians)


def backbone_with_amide_atoms(cartesians: torch.Tensor, N_indices: Sequence[int], C_indices: Sequence[int]) -> torch.Tensor:
    """``merge_cartesians(c, N, C, guess_amide_H(c, N), guess_amide_O(c, C))`` as ONE launch (the H and O arrays never exist)."""
    return _ops.backbone_amide_raw(cartesians, list(N_indices)[1::], list(C_indices), *AMIDE_H, *AMIDE_O)


# ---- topology-aware back-mapping: the rotation loop of mdtraj_backmapping -------------------------------------------------------
def near_and_far_sides(n_atoms: int, bonds, edges):
    """Atoms on either side of every edge of a tree-like bond graph: ``(near_sides, far_sides)``, lists of sorted integer arrays,
    edge[0] on the near side and edge[1] on the far side.  Reference: ``_get_near_and_far_networkx``
    (encodermap/misc/rotate.py:409-511), which removes the edge from a networkx graph and takes the two connected components;
    here a plain breadth-first search (no networkx), same membership.  Raises as the reference does when removing the edge does
    not split the graph in two (rings, several chains)."""
    import numpy as np

    adj = [[] for _ in range(n_atoms)]
    for a, b in np.asarray(bonds, dtype=np.int64).reshape(-1, 2):
        adj[a].append(int(b))
        adj[b].append(int(a))
    near_sides, far_sides = [], []
    for u, v in np.asarray(edges, dtype=np.int64).reshape(-1, 2):
        u, v = int(u), int(v)
        if v not in adj[u]:
            raise Exception(f"Seems like the edge {(u, v)} is not part of the graph.")
        seen = np.zeros(n_atoms, dtype=bool)
        seen[v] = True
        stack = [v]
        while stack:                       # everything reachable from edge[1] without crossing the edge
            a = stack.pop()
            for b in adj[a]:
                if (a == v and b == u) or seen[b]:
                    continue
                seen[b] = True
                stack.append(b)
        if seen[u]:
            raise Exception(f"Splitting at edge {(u, v)} does not work: the two atoms are still connected (ring?).")
        far = np.flatnonzero(seen)
        other = np.zeros(n_atoms, dtype=bool)
        other[u] = True
        stack = [u]
        while stack:
            a = stack.pop()
            for b in adj[a]:
                if (a == u and b == v) or other[b]:
                    continue
                other[b] = True
                stack.append(b)
        if other.sum() + far.size != n_atoms:
            raise Exception("Protein might be cyclic or contain more than 1 chain.")
        near_sides.append(np.flatnonzero(other))
        far_sides.append(far)
    return near_sides, far_sides


def set_dihedrals(xyz: torch.Tensor, dihedral_indices, bond_indices, far_sides, dihedrals: torch.Tensor) -> torch.Tensor:
    """The rotation loop of ``mdtraj_backmapping`` (encodermap/misc/backmapping.py:1661-1690 for the backbone, :1722-1745 for the
    side chains -- concatenate the lists, backbone first): every frame starts from ``xyz`` ((n_atoms, 3) shared, or one structure
    per frame) and has its dihedrals set to ``dihedrals[frame]`` one after the other by rotating the far side of each central
    bond.  Returns (frames, n_atoms, 3); what is left to the caller is the topology work around it (which atoms, which bonds:
    mdtraj / networkx in the reference) and writing the trajectory."""
    return _ops.set_dihedrals_raw(xyz, dihedral_indices, bond_indices, far_sides, dihedrals)
