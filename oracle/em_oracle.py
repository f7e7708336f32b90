"""CPU oracle for the EncoderMap training hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, op for op and in the reference's own order of operations, the
algorithms on the hot path of AG-Peter/encodermap in torch-CPU.
It exists so that the CUDA kernels can be checked; it is never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import it;
  * nothing under ``encodermap_b200/`` imports it, and the product path fails loudly
    when the CUDA library is missing.

Parity pin status
-----------------
The reference's arithmetic lives in TensorFlow (``tensorflow>=2.15``, un-pinned,
``setup.py:48`` of the reference), which is not installed in this image, so the reference
itself cannot be run.  The restatement is pinned two ways (see ``tests/test_oracle_*``):

  1. against every known-answer vector the reference's own tests hold for this path
     (``tests/test_pairwise_distances.py``, ``tests/test_losses.py``,
     ``tests/test_dihedral_to_cartesian.py``, ``tests/test_backmapping_em1_em2.py``);
  2. against golden outputs produced by executing the reference's *own function
     bodies* (extracted from ``/root/reference`` with ``ast``) on a numpy-backed
     stand-in for the ``tf`` namespace -- ``tools/gen_golden.py`` -> ``tests/golden/*.npz``.

What neither covers -- TensorFlow's own kernels and its autodiff tie rules -- stays
**parity unpinned**: every gradient, and the float32 rounding of Eigen's pow/sqrt/GEMM.
Gradients here come from torch autograd in float64 over the restated forward.

All functions take/return torch tensors; ``dtype`` follows the inputs (the reference
computes in float32; float64 evaluation of the same algorithm is the parity target for
the kernels).
"""

from __future__ import annotations

import math
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch

pi = math.pi


def _t(x, dtype=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    arr = np.asarray(x)
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t if dtype is None else t.to(dtype)


# ----------------------------------------------------------------------------------------
# encodermap/misc/distances.py
# ----------------------------------------------------------------------------------------


def sigmoid(sig: float, a: float, b: float) -> Callable:
    """Sketch-map sigmoid closure.  Reference: encodermap/misc/distances.py:66-88
    (TF1 twin encodermap_tf1/misc.py:115-124):  1 - (1 + (2^(a/b) - 1) (r/sig)^a)^(-b/a)."""

    def func(r):
        return 1 - (1 + (2 ** (a / b) - 1) * (r / sig) ** a) ** (-b / a)

    return func


def periodic_distance(a, b, periodicity: float = 2 * pi):
    """Minimum-image distance.  Reference: encodermap/misc/distances.py:113-141."""
    a, b = _t(a), _t(b)
    d = torch.abs(b - a)
    return torch.minimum(d, periodicity - d)


def pairwise_dist_periodic(positions, periodicity: float):
    """All-pairs periodic distance through the (N,N,D) broadcast tensor.
    Reference: encodermap/misc/distances.py:144-176 (eps 1e-12 on zero components and on
    the result)."""
    positions = _t(positions)
    assert positions.dim() == 2
    vecs = periodic_distance(positions[:, None, :], positions[None, :, :], periodicity)
    mask = (vecs == 0.0).to(torch.float32).to(vecs.dtype)
    vecs = vecs + mask * 1e-12
    return torch.sqrt(torch.sum(torch.square(vecs), dim=2)) + 1.0e-12


def pairwise_dist(positions, squared: bool = False, flat: bool = False):
    """Gram-form Euclidean distance matrix.  Reference: encodermap/misc/distances.py:179-255
    (TF1 twin encodermap_tf1/misc.py:143-188).  Rank-2 input gains a leading batch axis."""
    positions = _t(positions)
    if positions.dim() == 2:
        positions = positions[None]
    gram = torch.matmul(positions, positions.transpose(1, 2))
    sq = torch.diagonal(gram, dim1=1, dim2=2)
    dist = sq[:, None, :] - 2.0 * gram + sq[:, :, None]
    dist = torch.clamp_min(dist, 0.0)
    if flat:
        n = positions.shape[1]
        keep = np.ones((n, n), dtype=bool)
        keep[np.tril_indices(n)] = False
        dist = dist[:, torch.from_numpy(keep)]
    if not squared:
        mask = (dist == 0.0).to(dist.dtype)
        dist = dist + mask * 1e-16
        dist = torch.sqrt(dist)
        dist = dist * (1.0 - mask)
    return dist


def periodic_distance_vjp_tf(a, b, periodicity, grad_out):
    """VJP of ``periodic_distance`` under TENSORFLOW's autodiff tie rules, stated explicitly (torch's own autograd
    splits the gradient of ``minimum`` evenly on ties, TensorFlow does not): ``abs' = sign`` (0 at 0) and
    ``minimum(x, y)`` routes the whole gradient to ``x`` where ``x <= y`` (tensorflow/python/ops/math_grad.py,
    ``_MinimumGrad`` -> ``_MaximumMinimumGrad(op, grad, math_ops.less_equal)``).
    For d = |b - a|, out = minimum(d, P - d):  d(out)/dd = +1 where d <= P - d, else -1.  Returns (grad_a, grad_b)."""
    a, b, g = np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(grad_out, np.float64)
    diff = b - a
    d = np.abs(diff)
    branch = np.where(d <= periodicity - d, 1.0, -1.0)
    gb = g * branch * np.sign(diff)
    return -gb, gb


def pairwise_dist_periodic_vjp_tf(positions, periodicity, grad_out):
    """VJP of ``pairwise_dist_periodic`` (encodermap/misc/distances.py:164-175) w.r.t. ``positions`` under TensorFlow's
    tie rules (see ``periodic_distance_vjp_tf``): a = x_i (axis 1 expanded), b = x_j (axis 0 expanded),
    V = min(d, P - d) + 1e-12 [V == 0], Dist = sqrt(sum_k V^2) + 1e-12, so d(Dist_ij)/d(V_ijk) = V_ijk / sqrt(sum V^2)."""
    x, g = np.asarray(positions, np.float64), np.asarray(grad_out, np.float64)
    a, b = x[:, None, :], x[None, :, :]
    diff = b - a
    d = np.abs(diff)
    v = np.minimum(d, periodicity - d)
    v = v + (v == 0.0) * 1e-12
    s = np.sqrt(np.sum(v * v, axis=2))
    gv = g[:, :, None] * v / s[:, :, None]
    gdiff = gv * np.where(d <= periodicity - d, 1.0, -1.0) * np.sign(diff)
    return gdiff.sum(axis=0) - gdiff.sum(axis=1)   # b = x_j collects +, a = x_i collects -


# ----------------------------------------------------------------------------------------
# encodermap/loss_functions/loss_functions.py
# ----------------------------------------------------------------------------------------

DEFAULT_SIG = (4.5, 12, 6, 1, 2, 6)  # encodermap/parameters/parameters.py:620


def sigmoid_loss(periodicity: float = 2 * pi, dist_sig_parameters: Sequence[float] = DEFAULT_SIG) -> Callable:
    """Reference: encodermap/loss_functions/loss_functions.py:301-369 (TF1 twin
    ``distance_cost`` encodermap_tf1/misc.py:88-112).  Mean over all N^2 ordered pairs."""
    sig_h = sigmoid(*dist_sig_parameters[:3])
    sig_l = sigmoid(*dist_sig_parameters[3:])

    def sigmoid_loss_func(y_true, y_pred):
        r_h, r_l = _t(y_true), _t(y_pred)
        if periodicity == float("inf"):
            dist_h = pairwise_dist(r_h)
        else:
            dist_h = pairwise_dist_periodic(r_h, periodicity)
        dist_l = pairwise_dist(r_l)
        return torch.mean(torch.square(sig_h(dist_h) - sig_l(dist_l)))

    return sigmoid_loss_func


def distance_loss_value(y_true, latent, periodicity=2 * pi, dist_sig_parameters=DEFAULT_SIG,
                        distance_cost_scale: Optional[float] = 500.0):
    """The arithmetic of ``distance_loss_func`` once the encoder output is known.
    Reference: encodermap/loss_functions/loss_functions.py:266-296 (tuple inputs are
    concatenated over their first three members, scale None => 0)."""
    if isinstance(y_true, (tuple, list)):
        y_true = torch.cat([_t(y) for y in y_true[:3]], dim=1)
    if distance_cost_scale is None:
        return torch.zeros((), dtype=_t(latent).dtype)
    return sigmoid_loss(periodicity, dist_sig_parameters)(y_true, latent) * distance_cost_scale


def cartesian_distance_loss_value(y_true, latent, cartesian_dist_sig_parameters=DEFAULT_SIG,
                                  cartesian_distance_cost_scale: Optional[float] = 1.0):
    """Reference: encodermap/loss_functions/loss_functions.py:917-942 (periodicity forced
    to inf, second sigmoid parameter set)."""
    if cartesian_distance_cost_scale is None:
        return torch.zeros((), dtype=_t(latent).dtype)
    f = sigmoid_loss(float("inf"), cartesian_dist_sig_parameters)
    return f(y_true, latent) * cartesian_distance_cost_scale


# ----------------------------------------------------------------------------------------
# encodermap/models/layers.py
# ----------------------------------------------------------------------------------------


def periodic_input(x, periodicity: float = 2 * pi):
    """Reference: encodermap/models/layers.py:204-215 (twin models/models.py:3345-3350)."""
    x = _t(x)
    if periodicity != 2 * pi:
        x = x / periodicity * 2 * pi
    return torch.cat([torch.sin(x), torch.cos(x)], dim=1)


def pairwise_distances_layer(x, start=None, stop=None, step=None):
    """Reference: encodermap/models/layers.py:1252-1267 (no-sidechain branch)."""
    x = _t(x)
    return pairwise_dist(x[:, start:stop:step], flat=True)


# ----------------------------------------------------------------------------------------
# encodermap/encodermap_tf1/backmapping.py and encodermap/misc/backmapping.py
# ----------------------------------------------------------------------------------------


def straight_tetrahedral_chain(n_atoms=None, bond_lengths=None) -> np.ndarray:
    """Reference: encodermap/encodermap_tf1/backmapping.py:71-94 (float32 output)."""
    dx = math.cos(70.63 / 180 * pi)
    dy = math.sin(70.63 / 180 * pi)
    if n_atoms and not bond_lengths:
        xyz = np.zeros((n_atoms, 3), dtype=np.float32)
        idx = np.repeat(np.arange(int(n_atoms / 2) + 1), 2)
        xyz[:, 0] = idx[1 : n_atoms + 1] + dx * idx[0:n_atoms]
        xyz[:, 1] = dy * idx[0:n_atoms]
    elif (bond_lengths and not n_atoms) or n_atoms == len(bond_lengths) + 1:
        n_bonds = len(bond_lengths)
        n_atoms = n_atoms or n_bonds + 1
        dxs = bond_lengths * np.tile([1, dx], int(n_atoms / 2))[:n_bonds]
        dys = bond_lengths * np.tile([0, dy], int(n_atoms / 2))[:n_bonds]
        xyz = np.zeros((n_atoms, 3), dtype=np.float32)
        xyz[1:, 0] = np.cumsum(dxs)
        xyz[1:, 1] = np.cumsum(dys)
    else:
        raise ValueError("input not compatible")
    return xyz


def chain_in_plane(lengths, angles):
    """Planar zig-zag chain.  Reference: encodermap/encodermap_tf1/backmapping.py:97-119.
    ``lengths`` is (1, n-1) (broadcast over the batch) or (B, n-1); ``angles`` (B, n-2)."""
    lengths, angles = _t(lengths), _t(angles)
    batch = angles.shape[0]
    prev = torch.zeros(batch, dtype=angles.dtype)
    xs = [torch.zeros(batch, dtype=angles.dtype)]
    ys = [torch.zeros(batch, dtype=angles.dtype)]
    sign = 1
    i = -1
    for i in range(angles.shape[1]):
        xs.append(xs[-1] + lengths[:, i] * torch.cos(prev))
        ys.append(ys[-1] + lengths[:, i] * torch.sin(prev) * sign)
        prev = pi - angles[:, i] - prev
        sign *= -1
    xs.append(xs[-1] + lengths[:, i + 1] * torch.cos(prev))
    ys.append(ys[-1] + lengths[:, i + 1] * torch.sin(prev) * sign)
    xs = torch.stack(xs, dim=1)
    ys = torch.stack(ys, dim=1)
    return torch.stack([xs, ys, torch.zeros_like(xs)], dim=2)


def rotation_matrix(axis_unit_vec, angle):
    """Rodrigues matrix, applied to ROW vectors on the right.
    Reference: encodermap/misc/backmapping.py:1950-1968 (twin encodermap_tf1/misc.py:286-304)."""
    u, angle = _t(axis_unit_vec), _t(angle)
    ang = angle[:, None, None]
    eye = torch.eye(3, dtype=u.dtype)[None]
    z = torch.zeros(u.shape[0], dtype=u.dtype)
    cross = torch.stack(
        [
            torch.stack([z, -u[:, 2], u[:, 1]], dim=0),
            torch.stack([u[:, 2], z, -u[:, 0]], dim=0),
            torch.stack([-u[:, 1], u[:, 0], z], dim=0),
        ],
        dim=0,
    ).permute(2, 0, 1)
    r = torch.cos(ang) * eye
    r = r + torch.sin(ang) * cross
    uu = u[:, :, None]
    r = r + (1 - torch.cos(ang)) * torch.matmul(uu, uu.transpose(1, 2))
    return r


def dihedral_to_cartesian_one_way(dihedrals, cartesian, n: Optional[int] = None):
    """One tail of the chain.  Reference: encodermap/misc/backmapping.py:1873-1912 (explicit
    sqrt-of-sum norm; TF1 twin encodermap_tf1/backmapping.py:198-214 uses tf.norm)."""
    dihedrals, cartesian = _t(dihedrals), _t(cartesian)
    if n is None:
        n = dihedrals.shape[-1]
    dihedrals = -dihedrals
    rotated = cartesian[:, 1:]
    collected = [cartesian[:, :1]]
    for i in range(n):
        collected.append(rotated[:, 0:1])
        axis = rotated[:, 1] - rotated[:, 0]
        axis = axis / torch.sqrt(torch.sum(torch.square(axis), dim=1))[:, None]
        offset = rotated[:, 1:2]
        rotated = offset + torch.matmul(rotated[:, 1:] - offset, rotation_matrix(axis, dihedrals[:, i]))
    collected.append(rotated)
    return torch.cat(collected, dim=1)


def split_and_reverse_dihedrals(x):
    """Reference: encodermap/misc/backmapping.py:179-214."""
    x = _t(x)
    middle = int(int(x.shape[1]) / 2)
    if x.shape[1] % 2 == 0:
        return torch.flip(x[:, :middle], dims=[1]), x[:, middle:]
    return torch.flip(x[:, : middle + 1], dims=[1]), x[:, middle + 1 :]


def split_and_reverse_cartesians(x):
    """Reference: encodermap/misc/backmapping.py:217-256."""
    x = _t(x)
    split = int(int(x.shape[1]) / 2)
    return torch.flip(x[:, : split + 2], dims=[1]), x[:, split - 1 :]


def split_indices_tf1(n_atoms: int):
    """The TF1 slicing the reference's split test pins the TF2 helpers against.
    Reference: encodermap/encodermap_tf1/backmapping.py:175-181 and
    tests/test_backmapping_em1_em2.py:2130-2137.  Returns numpy index arrays
    (left_atoms, left_dihedrals, right_atoms, right_dihedrals)."""
    atoms = np.arange(n_atoms)
    dih = np.arange(n_atoms - 3)
    split = int(int(n_atoms) / 2)
    return (atoms[split + 1 :: -1], dih[split - 2 :: -1], atoms[split - 1 :], dih[split - 1 :])


def dihedrals_to_cartesian_layers(dihedrals, cartesians, left_iteration_counter: int, right_iteration_counter: int):
    """Two-sided build.  Reference: encodermap/misc/backmapping.py:259-309."""
    dihedrals, cartesians = _t(dihedrals), _t(cartesians)
    if cartesians.dim() == 2:
        cartesians = cartesians[None].expand(dihedrals.shape[0], -1, -1)
    c_left, c_right = split_and_reverse_cartesians(cartesians)
    d_left, d_right = split_and_reverse_dihedrals(dihedrals)
    new_left = dihedral_to_cartesian_one_way(d_left, c_left, left_iteration_counter)
    new_right = dihedral_to_cartesian_one_way(d_right, c_right, right_iteration_counter)
    return torch.cat([torch.flip(new_left, dims=[1]), new_right[:, 3:]], dim=1)


def dihedrals_to_cartesian_tf1(dihedrals, cartesian):
    """TF1 twin with its own slicing.  Reference: encodermap/encodermap_tf1/backmapping.py:164-195."""
    dihedrals, cartesian = _t(dihedrals), _t(cartesian)
    if cartesian.dim() == 2:
        cartesian = cartesian[None].expand(dihedrals.shape[0], -1, -1)
    n_atoms = cartesian.shape[1]
    la, ld, ra, rd = split_indices_tf1(n_atoms)
    c_right = cartesian[:, torch.from_numpy(ra.copy())]
    d_right = dihedrals[:, torch.from_numpy(rd.copy())]
    c_left = cartesian[:, torch.from_numpy(la.copy())]
    d_left = dihedrals[:, torch.from_numpy(ld.copy())]
    new_right = dihedral_to_cartesian_one_way(d_right, c_right)
    new_left = dihedral_to_cartesian_one_way(d_left, c_left)
    return torch.cat([torch.flip(new_left, dims=[1]), new_right[:, 3:]], dim=1)


def split_counts(n_atoms: int) -> Tuple[int, int]:
    """(left_split, right_split) loop counts.  Reference: encodermap/models/models.py:661-671."""
    return n_atoms // 2 - 1, (n_atoms - 3) // 2


def back_map_layer(distances, angles, dihedrals, left_split: Optional[int] = None, right_split: Optional[int] = None):
    """Reference: encodermap/models/layers.py:957-986: batch-mean bond lengths ->
    chain_in_plane -> dihedrals + pi -> two-sided build."""
    distances, angles, dihedrals = _t(distances), _t(angles), _t(dihedrals)
    n_atoms = distances.shape[1] + 1
    if left_split is None or right_split is None:
        left_split, right_split = split_counts(n_atoms)
    lengths = torch.mean(distances, dim=0)[None]
    chain = chain_in_plane(lengths, angles)
    return dihedrals_to_cartesian_layers(dihedrals + pi, chain, left_split, right_split)


# ----------------------------------------------------------------------------------------
# helpers used by the tests and the CPU baseline (not reference functions)
# ----------------------------------------------------------------------------------------


# ----------------------------------------------------------------------------------------
# encodermap/models/layers.py: BackMapLayerWithSidechains
# ----------------------------------------------------------------------------------------


def sidechain_topology(counts: Sequence[int]) -> dict:
    """Index tables of BackMapLayerWithSidechains for ``counts[r]`` side-chain dihedrals in residue r + 1.
    Reference: encodermap/models/layers.py:234-474 (numpy twin misc/backmapping.py:571-798).  Atom order: the 3 n backbone
    atoms, then per residue with a side chain its ``count + 1`` atoms.  ``True`` in a mask row = the atom stays where it is.

    The reference's construction only closes (its hstack of 3 n - 1 rows) when exactly one of the first / last residue has no
    side chain, and an interior residue without one re-uses the rows of the previous residue: restated as is.
    """
    counts = [int(c) for c in counts]
    n_res = len(counts)
    n_bb = 3 * n_res
    n_side = sum(c + 1 for c in counts if c > 0)
    if n_side == 0:
        raise ValueError("no side chain at all: the reference layer cannot be built (side_angle_indices undefined, layers.py:477)")
    central_left = np.tri(n_bb - 1, n_bb, 0).astype(bool)                                      # :249-253
    right_rows = [np.zeros((1, n_side), bool)]                                                  # :254-256
    filled = 0                  # side-chain atoms placed so far ("count", :257)
    next_atom = n_bb            # index of the next side chain's first atom ("count2 - 1", :258)
    side_triplets, side_quads, side_rows_of_dihedrals = [], [], []
    last_rows = None
    for r, c in enumerate(counts):                                                              # :280-349
        if c == 0:
            if r == 0 or r == n_res - 1:
                continue
            if last_rows is None:
                raise ValueError("a residue without side chain before the first side chain: NameError in the reference (:287)")
            right_rows.append(last_rows)
            continue
        n_, ca_ = 3 * r, 3 * r + 1
        chain = [n_, ca_] + list(range(next_atom, next_atom + c + 1))    # N, CA, CB, CG, ...
        side_rows_of_dihedrals += list(range(filled, filled + c))                              # :289-291
        for q in range(c + 1):
            side_triplets.append(chain[q:q + 3])                                                # :298-300, 312-314, 322-328
            if q < c:
                side_quads.append(chain[q:q + 4])                                               # :302-309, 316-319, 329-337
        filled += c + 1
        next_atom += c + 1
        last_rows = np.zeros((3, n_side), bool)                                                 # :340-345
        last_rows[:, :filled] = True
        right_rows.append(last_rows)
    right_rows.append(np.ones((1, n_side), bool))                                               # :357-359
    right = np.vstack(right_rows)
    if right.shape[0] != n_bb - 1:
        raise ValueError(f"the reference's index construction needs exactly one of the first / last residue without a side "
                         f"chain ({right.shape[0]} rows for {n_bb - 1} backbone bonds, layers.py:370-372)")
    central_mask = np.hstack([central_left, right])                                             # :370-372
    blocks = []
    for c in counts:                                                                            # :373-391
        if c > 0:
            blocks.append((np.tri(c + 1, c + 2, 0) + 1)[:, 1:])
    side_right = np.zeros((n_side, n_side))
    o = 0
    for b in blocks:
        side_right[o:o + b.shape[0], o:o + b.shape[1]] = b
        o += b.shape[0]
    side_mask = np.hstack([np.ones((n_side, n_bb), bool), (side_right % 2) == 0])
    bb = np.arange(n_bb)
    central_triplets = np.stack([bb[:-2], bb[1:-1], bb[2:]], 1)                                 # :261-267
    central_quads = np.stack([bb[:-3], bb[1:-2], bb[2:-1], bb[3:]], 1)                          # :269-277
    return {
        "n_atoms": n_bb + n_side, "n_side": n_side, "n_res": n_res,
        "central_mask": central_mask,
        "central_angle_mask": central_mask[1:],                                                 # :410
        "side_angle_mask": side_mask,                                                           # :415
        "dihedral_mask": np.vstack([central_mask[1:-1], side_mask[np.array(side_rows_of_dihedrals, dtype=int)]]),   # :423-428
        "central_angle_triplets": central_triplets,
        "side_angle_triplets": np.array(side_triplets).reshape(-1, 3),
        "dihedral_quadruplets": np.vstack([central_quads, np.array(side_quads).reshape(-1, 4)]),
        "counts": np.array(counts),
    }


def _rotation_about(angle, direction, point):
    """Batch of 3x3 rotations by ``angle`` about ``direction`` (normalised here) and the translation that makes them rotations
    about ``point``.  Reference: _rotation_matrices, encodermap/models/layers.py:859-899."""
    u = direction / torch.sqrt((direction ** 2).sum(1, keepdim=True))
    ca, sa = torch.cos(angle), torch.sin(angle)
    eye = torch.eye(3, dtype=angle.dtype).expand(angle.shape[0], 3, 3)
    rot = eye * ca[:, None, None] + u[:, :, None] * u[:, None, :] * (1.0 - ca)[:, None, None]
    us = u * sa[:, None]
    zero = torch.zeros_like(sa)
    skew = torch.stack([torch.stack([zero, -us[:, 2], us[:, 1]], 1), torch.stack([us[:, 2], zero, -us[:, 0]], 1),
                        torch.stack([-us[:, 1], us[:, 0], zero], 1)], 1)
    rot = rot + skew
    shift = point - torch.einsum("bij,bj->bi", rot, point)
    return rot, shift


#: a measured bond angle whose cosine is this close to +-1 is treated as a constant in the backward pass: acos is not
#: differentiable there (TensorFlow yields 0, a huge number or NaN depending on how the float32 cosine happened to round)
STRAIGHT_EPS = 1e-12


def backmap_with_sidechains(counts: Sequence[int], inputs, topology: Optional[dict] = None):
    """BackMapLayerWithSidechains.call: (central_distances, central_angles, central_dihedrals, side_distances, side_angles,
    side_dihedrals) -> (batch, n_atoms, 3).  Reference: encodermap/models/layers.py:533-843; differentiable (torch autograd
    over this restatement is the gradient oracle).  All atoms start in the z = 0 plane: the backbone on the x axis at the
    running sum of its bond lengths, every side chain straight up in y from its CA (:593-648); then every bond angle
    (backbone: about +z, :654-717; side chains: about -z, :720-783) and every dihedral (:786-841) is set one after the other by
    rotating the atoms whose mask entry is False about the pivot / bond, by |target - measured| for the bond angles and
    target - measured for the dihedrals."""
    topo = topology or sidechain_topology(counts)
    cd, ca_, cdih, sd, sa_, sdih = [_t(x) for x in inputs]
    dt = cd.dtype
    nb = cd.shape[0]
    counts = [int(c) for c in topo["counts"]]
    n_bb = 3 * len(counts)
    zero = torch.zeros(nb, 1, dtype=dt)
    xs_c = torch.cat([zero, torch.cumsum(cd, 1)], 1)                                            # :593-606
    xs_s, ys_s = [], []
    j = 0
    for r, c in enumerate(counts):                                                              # :607-628
        if c > 0:
            for n in range(c + 1):
                xs_s.append(xs_c[:, 3 * r + 1])
                ys_s.append(sd[:, j - n:j + 1].sum(1))
                j += 1
    xs = torch.cat([xs_c, torch.stack(xs_s, 1)], 1)
    ys = torch.cat([torch.zeros(nb, n_bb, dtype=dt), torch.stack(ys_s, 1)], 1)
    xyz = torch.stack([xs, ys, torch.zeros_like(xs)], 2)                                        # :635-648

    def apply(xyz, rot, shift, mask_row):
        moved = torch.einsum("bij,bnj->bni", rot, xyz) + shift[:, None, :]
        keep = torch.from_numpy(np.ascontiguousarray(mask_row))[None, :, None]
        return torch.where(keep, xyz, moved)

    def set_angles(xyz, targets, masks, triplets, z):
        axis = torch.tensor([[0.0, 0.0, z]], dtype=dt).expand(nb, 3)
        for i in range(masks.shape[0]):
            a, b, c = (xyz[:, int(k)] for k in triplets[i])
            ba, bc = a - b, c - b
            t = (ba * bc).sum(1) / (torch.sqrt((ba ** 2).sum(1)) * torch.sqrt((bc ** 2).sum(1)))
            t = torch.clamp(t, -1.0, 1.0)
            t = torch.where(1.0 - t.detach() ** 2 < STRAIGHT_EPS, t.detach(), t)
            angle = torch.abs(targets[:, i] - torch.acos(t))
            rot, shift = _rotation_about(angle, axis, b)
            xyz = apply(xyz, rot, shift, masks[i])
        return xyz

    xyz = set_angles(xyz, ca_, topo["central_angle_mask"], topo["central_angle_triplets"], 1.0)
    xyz = set_angles(xyz, sa_, topo["side_angle_mask"], topo["side_angle_triplets"], -1.0)
    dih = torch.cat([cdih, sdih], 1)                                                            # :546-552
    for i in range(topo["dihedral_mask"].shape[0]):
        a, b, c, d = (xyz[:, int(k)] for k in topo["dihedral_quadruplets"][i])
        b1, b2, b3 = b - a, c - b, d - c
        c1 = torch.linalg.cross(b2, b3)
        c2 = torch.linalg.cross(b1, b2)
        p1 = (b1 * c1).sum(1) * torch.sqrt((b2 * b2).sum(1))
        p2 = (c1 * c2).sum(1)
        angle = dih[:, i] - torch.atan2(p1, p2)
        rot, shift = _rotation_about(angle, c - b, b)
        xyz = apply(xyz, rot, shift, topo["dihedral_mask"][i])
    return xyz


def sidechain_pairwise_indices(counts: Sequence[int], start=None, stop=None, step=None) -> np.ndarray:
    """Atoms PairwiseDistances selects when side chains are reconstructed: the sliced backbone plus one atom index per residue
    with a side chain.  Reference: encodermap/models/layers.py:1188-1208 (the running index advances by the residue's number of
    side-chain DIHEDRALS, one less than its atoms: restated as is)."""
    n_res = len(counts)
    first = np.arange(3 * n_res)[start:stop:step]
    atom = 3 * n_res + 1
    extra = []
    for c in counts:
        if c == 0:
            continue
        atom += int(c)
        extra.append(atom)
    return np.concatenate([first, np.array(extra, dtype=first.dtype)])


# ---- generation side: guessed amide H / carbonyl O and the merge --------------------------------------
def guess_sp2_atom(cartesians, indices, angle_to_previous: float, bond_length: float):
    """Reference: encodermap/misc/backmapping.py:1920-1941.  ``cartesians[:, i + 1]`` past the last atom raises in
    TensorFlow and the reference falls back to atom i - 2 (:1926-1929); negative positions wrap as Python indexing does."""
    x = _t(cartesians)
    n = x.shape[1]
    added = []
    for i in indices:
        prev_vec = x[:, i - 1] - x[:, i]
        next_vec = (x[:, i + 1] if i + 1 < n else x[:, i - 2]) - x[:, i]
        axis = torch.linalg.cross(prev_vec, next_vec)
        axis = axis / torch.sqrt((axis * axis).sum(dim=1, keepdim=True))
        ang = torch.full((x.shape[0],), angle_to_previous, dtype=x.dtype)
        bond = torch.matmul(prev_vec[:, None, :], rotation_matrix(axis, ang))[:, 0, :]
        bond = bond * (bond_length / torch.sqrt((bond * bond).sum(dim=1, keepdim=True)))
        added.append(x[:, i] + bond)
    return torch.stack(added, dim=1)


def guess_amide_H(cartesians, N_indices):
    """Reference: encodermap/misc/backmapping.py:1943-1944."""
    return guess_sp2_atom(cartesians, list(N_indices)[1::], 123 / 180 * pi, 1.10)


def guess_amide_O(cartesians, C_indices):
    """Reference: encodermap/misc/backmapping.py:1946-1947."""
    return guess_sp2_atom(cartesians, list(C_indices), 121 / 180 * pi, 1.24)


def merge_cartesians(central_cartesians, N_indices, O_indices, H_cartesians, O_cartesians):
    """Reference: encodermap/misc/backmapping.py:1970-1990."""
    c, h, o = _t(central_cartesians), _t(H_cartesians), _t(O_cartesians)
    n_tail, o_set = set(list(N_indices)[1::]), set(O_indices)
    out = [c[:, 0]]
    h_i = o_i = 0
    for i in range(1, c.shape[1]):
        out.append(c[:, i])
        if i in n_tail:
            out.append(h[:, h_i])
            h_i += 1
        elif i in o_set:
            out.append(o[:, o_i])
            o_i += 1
    out = torch.stack(out, dim=1)
    assert out.shape[1] == c.shape[1] + h.shape[1] + o.shape[1]
    return out


# ---- topology-aware back-mapping: the rotation loop of mdtraj_backmapping (numpy; the reference's loop is numpy too) ------------
def dihedral_np(xyz: np.ndarray, indices) -> float:
    """Reference: encodermap/misc/rotate.py:547-581 (adapted from MDTraj there; numba twin misc/backmapping.py:329-352)."""
    a, b, c, d = (xyz[i] for i in indices)
    b1, b2, b3 = b - a, c - b, d - c
    c1, c2 = np.cross(b2, b3), np.cross(b1, b2)
    p1 = (b1 * c1).sum(-1) * (b2 * b2).sum(-1) ** 0.5
    p2 = (c1 * c2).sum(-1)
    return np.arctan2(p1, p2)


def rotation_matrix_about(angle: float, direction: np.ndarray, pivot: np.ndarray) -> np.ndarray:
    """4x4 homogeneous rotation about an axis through ``pivot``: transformations.rotation_matrix, restated by the reference as
    _rotmat_jit (encodermap/misc/backmapping.py:356-381)."""
    sina, cosa = np.sin(angle), np.cos(angle)
    u = direction / (direction ** 2).sum() ** 0.5
    r = np.identity(3) * cosa + np.outer(u, u) * (1.0 - cosa)
    us = u * sina
    r = r + np.array([[0.0, -us[2], us[1]], [us[2], 0.0, -us[0]], [-us[1], us[0], 0.0]])
    m = np.identity(4)
    m[:3, :3] = r
    m[:3, 3] = pivot - r @ pivot
    return m


def set_dihedrals(xyz, dihedral_indices, bond_indices, far_sides, dihedrals) -> np.ndarray:
    """Reference: encodermap/misc/backmapping.py:1661-1690 (and :1722-1745 for the side-chain dihedrals), in float64."""
    xyz = np.asarray(xyz, dtype=np.float64)
    dihedrals = np.asarray(dihedrals, dtype=np.float64)
    frames = dihedrals.shape[0]
    new_xyz = np.repeat(xyz[None], frames, 0).copy() if xyz.ndim == 2 else xyz.copy()
    new_xyz = np.pad(new_xyz, ((0, 0), (0, 0), (0, 1)), mode="constant", constant_values=1)
    for i in range(frames):
        for j in range(dihedrals.shape[1]):
            far_side, dihedral, bond = far_sides[j], dihedral_indices[j], bond_indices[j]
            angle = dihedrals[i, j] - dihedral_np(new_xyz[i, :, :3], dihedral)
            direction = np.diff(new_xyz[i, bond, :3], axis=0).flatten()
            rotmat = rotation_matrix_about(angle, direction, new_xyz[i, bond[0], :3])
            new_xyz[i, far_side, :3] = rotmat.dot(new_xyz[i, far_side].T).T[:, :3]
    return new_xyz[..., :3]


def sigmoid_loss_and_grad(y_true, y_pred, periodicity=2 * pi, sig=DEFAULT_SIG, dtype=torch.float64):
    """Loss and dL/d(y_pred) by autograd over the restated forward (the reference relies on
    tf.GradientTape; only the latent side needs a gradient)."""
    h = _t(y_true).to(dtype)
    z = _t(y_pred).to(dtype).clone().requires_grad_(True)
    loss = sigmoid_loss(periodicity, sig)(h, z)
    (g,) = torch.autograd.grad(loss, z)
    return loss.detach(), g


def sigmoid_loss_tiles(y_true, y_pred, periodicity, sig, tile_begin: int, tile_end: int, tile: int = 128,
                       dtype=torch.float64):
    """Partial loss / gradient over a slice of the upper-triangular tile list (row-major over
    tile rows), normalised by the full N^2 -- the quantity one rank of the sharded evaluation
    produces.  Sum over a partition of the tile list == full result."""
    h = _t(y_true).to(dtype)
    z = _t(y_pred).to(dtype)
    n = h.shape[0]
    nt = (n + tile - 1) // tile
    sig_h, sig_l = sigmoid(*sig[:3]), sigmoid(*sig[3:])
    loss = torch.zeros((), dtype=dtype)
    grad = torch.zeros_like(z)
    t = 0
    for ti in range(nt):
        for tj in range(ti, nt):
            if tile_begin <= t < tile_end:
                ri = slice(ti * tile, min(n, (ti + 1) * tile))
                rj = slice(tj * tile, min(n, (tj + 1) * tile))
                zi = z[ri].clone().requires_grad_(True)
                zj = z[rj].clone().requires_grad_(True)
                if periodicity == float("inf"):
                    dh = torch.sqrt(torch.sum((h[ri][:, None] - h[rj][None]) ** 2, dim=2))
                else:
                    dh = torch.sqrt(torch.sum(periodic_distance(h[ri][:, None], h[rj][None], periodicity) ** 2, dim=2))
                d2 = torch.sum((zi[:, None] - zj[None]) ** 2, dim=2)
                m = (d2 == 0).to(dtype)
                dl = torch.sqrt(d2 + m) * (1 - m)
                w = 1.0 if ti == tj else 2.0
                part = w * torch.sum((sig_h(dh) - sig_l(dl)) ** 2) / (n * n)
                gi, gj = torch.autograd.grad(part, (zi, zj))
                loss = loss + part.detach()
                grad[ri] += gi
                grad[rj] += gj
            t += 1
    return loss, grad


def dihedral_of(p0, p1, p2, p3):
    """Signed dihedral of four points in the convention of mdtraj's ``compute_dihedrals``
    (b1.(b2 x b3)|b2| , (b1 x b2).(b2 x b3)) -- the convention the reference's "recomputed
    dihedrals == requested" check uses (tests/test_losses.py:663-703)."""
    b1, b2, b3 = p1 - p0, p2 - p1, p3 - p2
    c1 = torch.cross(b2, b3, dim=-1)
    c2 = torch.cross(b1, b2, dim=-1)
    y = torch.sum(b1 * c1, dim=-1) * torch.linalg.norm(b2, dim=-1)
    x = torch.sum(c1 * c2, dim=-1)
    return torch.atan2(y, x)


def dihedral_vjp_from_xyz(xyz, grad_xyz):
    """Closed form of d<grad_xyz, BackMapLayer(...)>/d(dihedrals), float64, evaluated ON THE GIVEN coordinates
    (test infrastructure for chains where float32 coordinates limit what any backward can know: at 60 nm from the
    origin a float32 coordinate carries 4e-6 nm of rounding, 3e-5 relative on a 0.14 nm bond vector).
    Dihedral d twists the end of the chain that does not hold the three middle atoms (reference
    misc/backmapping.py:259-309: the middle atoms keep their planar positions) about its bond:
    left of the anchor (d < n/2 - 1) atoms 0..d about the bond (d+1, d+2), else atoms d+3.. about the bond (d+1, d+2).
    Checked against float64 autograd of back_map_layer to 1e-11 (tests/test_oracle_kats.py)."""
    x = np.asarray(xyz, dtype=np.float64)
    g = np.asarray(grad_xyz, dtype=np.float64)
    n = x.shape[-2]
    dr0 = n // 2 - 1
    out = np.zeros(x.shape[:-2] + (n - 3,))
    for d in range(n - 3):
        if d < dr0:
            tq = np.cross(x[..., : d + 1, :] - x[..., d + 1 : d + 2, :], g[..., : d + 1, :]).sum(-2)
            u = x[..., d + 2, :] - x[..., d + 1, :]
            out[..., d] = -(u * tq).sum(-1) / np.linalg.norm(u, axis=-1)
        else:
            k = d + 2
            tq = np.cross(x[..., k + 1 :, :] - x[..., k : k + 1, :], g[..., k + 1 :, :]).sum(-2)
            u = x[..., k, :] - x[..., k - 1, :]
            out[..., d] = (u * tq).sum(-1) / np.linalg.norm(u, axis=-1)
    return out
