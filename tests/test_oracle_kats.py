"""Pin the oracle against the known-answer vectors the reference's OWN tests hold for the path.

Each test names the reference test it restates.  CPU only.
"""
import math

import numpy as np
import pytest
import torch
from scipy.spatial.distance import cdist

from oracle import em_oracle as O

pi = math.pi
DEFAULT_SIG = (4.5, 12, 6, 1, 2, 6)


def _np_sigmoid(r, sig, a, b):
    return 1 - (1 + (2 ** (a / b) - 1) * (r / sig) ** a) ** (-b / a)


def _min_image_metric(periodicity):
    # reference tests/test_pairwise_distances.py:65-75
    def func(i, j):
        dx = np.linalg.norm(j - i)
        if dx > periodicity * 0.5:
            dx = dx - periodicity
        if dx <= -periodicity * 0.5:
            dx = dx + periodicity
        return np.abs(dx)

    return func


def test_periodic_distance_kat():
    # reference tests/test_pairwise_distances.py:98-108 (exact equality)
    a = np.array([0.0, 0.0, 0.0])
    b = np.array([pi / 2, pi, 3 / 2 * pi])
    d = O.periodic_distance(a, b, 2 * pi).numpy()
    assert np.array_equal(d, np.array([pi / 2, pi, pi / 2]))


def test_periodic_distance_docstring_kat():
    # reference encodermap/misc/distances.py:131-137
    d = O.periodic_distance(np.array([[1.5], [1.5]]), np.array([[-3.1], [-3.1]])).numpy()
    np.testing.assert_allclose(d, [[1.68318531], [1.68318531]], atol=1e-8)


def test_pairwise_dist_periodic_kats():
    # reference tests/test_pairwise_distances.py:140-148
    pts = np.array([[1 / 8, 1 / 2], [7 / 8, 1 / 2]], dtype=np.float32)
    np.testing.assert_allclose(O.pairwise_dist_periodic(pts, 1).numpy(), [[0, 1 / 4], [1 / 4, 0]], atol=1e-6)
    np.testing.assert_allclose(O.pairwise_dist_periodic(pts, float("inf")).numpy(), [[0, 6 / 8], [6 / 8, 0]], atol=1e-6)


def test_pairwise_dist_kat_and_flat_order():
    # reference tests/test_pairwise_distances.py:150-153 and :164-168 (pins pair order)
    d = O.pairwise_dist(np.array([[1 / 8, 1 / 2], [7 / 8, 1 / 2]], dtype=np.float32)).numpy()
    np.testing.assert_allclose(d, [[[0, 6 / 8], [6 / 8, 0]]], atol=1e-6)
    f = O.pairwise_dist(np.array([[0, 0], [1, 0], [0, 1]], dtype=np.float32), flat=True).numpy()
    np.testing.assert_allclose(f, [[1, 1, 2 ** 0.5]], atol=1e-6)


def test_periodic_many_points():
    # reference tests/test_pairwise_distances.py:110-135
    metric = _min_image_metric(360.0)
    pts = np.vstack([20.0, 40.0, 340.0]).astype(np.float32)
    np.testing.assert_allclose(cdist(pts, pts, metric), O.pairwise_dist_periodic(pts, 360.0).numpy(), atol=1e-5)
    np.random.seed(1)
    pts = np.random.random((100, 1)).astype(np.float32) * 360.0
    np.testing.assert_allclose(cdist(pts, pts, metric), O.pairwise_dist_periodic(pts, 360.0).numpy(), atol=1e-5)


def test_periodic_equals_nonperiodic_for_huge_period():
    # reference tests/test_pairwise_distances.py:155-162
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(10, 3)).astype(np.float32)
    np.testing.assert_allclose(O.pairwise_dist_periodic(pts, 10000).numpy().reshape(-1),
                               O.pairwise_dist(pts).numpy().reshape(-1), atol=1e-5)


def test_sigmoid_of_pairwise_vs_scipy():
    # reference tests/test_pairwise_distances.py:80-93
    np.random.seed(1)
    pts = np.random.random((100, 2)).astype(np.float32) * 360.0
    ref = cdist(pts, pts, lambda i, j: _np_sigmoid(np.linalg.norm(j - i), *DEFAULT_SIG[:3]))
    got = O.sigmoid(*DEFAULT_SIG[:3])(O.pairwise_dist(pts)).numpy().squeeze()
    np.testing.assert_allclose(ref, got, atol=1e-3)


def test_sigmoid_scalar_kat():
    # reference tests/test_pairwise_distances.py:186-193: em1 sigmoid(1.5, 5, 12, 2) == em2 sigmoid(5, 12, 2)(1.5)
    assert O.sigmoid(5, 12, 2)(1.5) == _np_sigmoid(1.5, 5, 12, 2)


@pytest.mark.parametrize("periodic", [False, True])
def test_sigmoid_loss_vs_scipy(periodic):
    # reference tests/test_losses.py:195-280 (re-seeded; the reference draws unseeded)
    rng = np.random.default_rng(7)
    if periodic:
        highd = (rng.random((256, 51)).astype("float32") * 2 * np.pi) - np.pi
        per = 2 * np.pi

        def hmetric(i, j):
            return _np_sigmoid(_min_image_metric(per)(i, j), *DEFAULT_SIG[:3])
    else:
        highd = rng.random((256, 51)).astype("float32") * 100
        per = float("inf")

        def hmetric(i, j):
            return _np_sigmoid(np.linalg.norm(j - i), *DEFAULT_SIG[:3])
    lowd = rng.random((256, 2)).astype("float32") * 10
    cost = O.sigmoid_loss(per, DEFAULT_SIG)(highd, lowd).item()
    if periodic:
        # the reference's numpy restatement (tests/test_losses.py:259-271)
        d = np.abs(highd[None] - highd[:, None])
        d = np.minimum(d, 2 * np.pi - d)
        d = d + np.equal(d, 0.0) * 1e-16
        sig_h = _np_sigmoid(np.linalg.norm(d, axis=2), *DEFAULT_SIG[:3])
    else:
        sig_h = cdist(highd, highd, hmetric)
    sig_l = cdist(lowd, lowd, lambda i, j: _np_sigmoid(np.linalg.norm(j - i), *DEFAULT_SIG[3:]))
    np.testing.assert_allclose(cost, np.mean(np.square(sig_h - sig_l)), rtol=1e-5)


def test_behavioural_zeros():
    # reference tests/test_losses.py:311-318 (uniform in + uniform latent => 0) and :897-904
    x = np.ones((32, 7), dtype=np.float32) * 0.3
    z = np.ones((32, 2), dtype=np.float32) * 1.7
    assert O.sigmoid_loss(2 * pi, DEFAULT_SIG)(x, z).item() == 0.0
    y = np.random.default_rng(3).random((20, 6)).astype(np.float32)
    assert O.cartesian_distance_loss_value(y, y, (1, 1, 1, 1, 1, 1)).item() == 0.0
    assert O.cartesian_distance_loss_value(np.zeros((20, 6), np.float32), np.zeros((20, 2), np.float32), (1, 1, 1, 1, 1, 1)).item() == 0.0


def test_straight_tetrahedral_chain_kat():
    # reference tests/test_dihedral_to_cartesian.py:186-197
    want = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.6633345, 1.8867929, 0.0], [4.6633344, 1.8867929, 0.0],
            [4.995002, 2.8301892, 0.0], [6.995002, 2.8301892, 0.0], [7.990003, 5.6603785, 0.0]]
    np.testing.assert_allclose(O.straight_tetrahedral_chain(bond_lengths=[1, 2, 3, 1, 2, 3]), want, rtol=1e-6)


def test_helix_kat_one_way(golden):
    # reference tests/test_dihedral_to_cartesian.py:98-153 (33x3 table, atol 1e-4; upstream skips it
    # because it predates the two-sided split -- it pins the one-way algorithm)
    g = golden["backmapping"]
    start = torch.from_numpy(O.straight_tetrahedral_chain(33)).double()[None].expand(2, -1, -1)
    out = O.dihedral_to_cartesian_one_way(torch.from_numpy(g["helix_dih"]), start).numpy()
    np.testing.assert_allclose(out[0], g["helix_kat"], atol=1e-4)


def test_split_indices_match_tf1_slicing():
    # reference tests/test_backmapping_em1_em2.py:2115-2156 (exact equality, random chain lengths)
    rng = np.random.default_rng(11)
    for n in list(rng.integers(6, 1000, size=10)) + [6, 7, 8, 9, 300, 1500]:
        n = int(n)
        atoms = torch.arange(n)[None]
        dih = torch.arange(n - 3)[None]
        la, ld, ra, rd = O.split_indices_tf1(n)
        cl, cr = O.split_and_reverse_cartesians(atoms)
        dl, dr = O.split_and_reverse_dihedrals(dih)
        assert np.array_equal(cl[0].numpy(), la) and np.array_equal(cr[0].numpy(), ra)
        assert np.array_equal(dl[0].numpy(), ld) and np.array_equal(dr[0].numpy(), rd)
        left, right = O.split_counts(n)
        assert left == len(ld) and right == len(rd)


def test_split_doctest_examples():
    # reference encodermap/misc/backmapping.py:186-199 (3 residues: 6 dihedrals -> 3 reversed + 3)
    d = torch.tensor([[3.69533481, 5.64050171, 5.60165278, 5.12605805, 0.22550092, 4.34644107]], dtype=torch.float64)
    left, right = O.split_and_reverse_dihedrals(d)
    assert left[0].tolist() == [5.60165278, 5.64050171, 3.69533481]
    assert right[0].tolist() == [5.12605805, 0.22550092, 4.34644107]
    c = torch.arange(27.0).reshape(1, 9, 3)
    cl, cr = O.split_and_reverse_cartesians(c)
    assert cl.shape == (1, 6, 3) and cr.shape == (1, 6, 3)
    assert torch.equal(cl[:, 0], cr[:, 2]) and torch.equal(cl[:, 1], cr[:, 1]) and torch.equal(cl[:, 2], cr[:, 0])


def test_backmap_realises_internal_coordinates():
    # the invariant behind reference tests/test_losses.py:663-703: recomputed bond lengths, angles and
    # dihedrals of the back-mapped chain equal the requested ones
    rng = np.random.default_rng(5)
    B, n = 3, 45
    dist = torch.from_numpy(rng.uniform(0.13, 0.15, size=(B, n - 1)))
    ang = torch.from_numpy(rng.uniform(1.9, 2.2, size=(B, n - 2)))
    dih = torch.from_numpy(rng.uniform(-pi, pi, size=(B, n - 3)))
    xyz = O.back_map_layer(dist, ang, dih)
    bonds = xyz[:, 1:] - xyz[:, :-1]
    np.testing.assert_allclose(torch.linalg.norm(bonds, dim=-1).numpy(), dist.mean(0)[None].expand(B, -1).numpy(), atol=1e-12)
    cosang = -(bonds[:, 1:] * bonds[:, :-1]).sum(-1) / (torch.linalg.norm(bonds[:, 1:], dim=-1) * torch.linalg.norm(bonds[:, :-1], dim=-1))
    np.testing.assert_allclose(torch.acos(cosang).numpy(), ang.numpy(), atol=1e-10)
    got = O.dihedral_of(xyz[:, :-3], xyz[:, 1:-2], xyz[:, 2:-1], xyz[:, 3:])
    diff = (got - dih + pi) % (2 * pi) - pi
    assert diff.abs().max().item() < 1e-9


def test_dihedral_vjp_closed_form_matches_autograd():
    """The closed-form dihedral VJP used to judge long chains equals float64 autograd of the restated BackMapLayer."""
    rng = np.random.default_rng(5)
    for n in (7, 8, 40, 41):
        b = 2
        dist = rng.uniform(0.13, 0.15, size=(b, n - 1))
        ang = rng.uniform(1.9, 2.2, size=(b, n - 2))
        dih = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(b, n - 3))).requires_grad_(True)
        w = rng.normal(size=(b, n, 3))
        xyz = O.back_map_layer(torch.from_numpy(dist), torch.from_numpy(ang), dih)
        (xyz * torch.from_numpy(w)).sum().backward()
        got = O.dihedral_vjp_from_xyz(xyz.detach().numpy(), w)
        np.testing.assert_allclose(got, dih.grad.numpy(), rtol=1e-9, atol=1e-11)


def _moving_pivot_vjp(xyz, g, lengths, angles):
    """Sequential float64 statement of the backward kernel's algorithm (backmap.cu, "backward, version 3"): both ends
    are walked towards the anchor, the wrench (S, M) of the end is kept about the atom the walk has reached, and the
    three terms of a cut are  -<unit(e2), M>,  <unit(e1 x e2), M>,  -<unit(e1), S>;  left of the anchor an angle /
    length also moves the whole molecule along the planar chain."""
    n = xyz.shape[0]
    dr0 = n // 2 - 1
    rD, rA, rL = np.zeros(n), np.zeros(n), np.zeros(n)          # per cut
    tot = {}
    for side in (0, 1):
        steps = dr0 if side == 0 else n - 1 - dr0
        a = (lambda i: i) if side == 0 else (lambda i: n - 1 - i)
        S, M = np.zeros(3), np.zeros(3)
        for i in range(steps):
            S = S + g[a(i)]
            e1 = xyz[a(i + 1)] - xyz[a(i)]
            M = M - np.cross(e1, S)
            j2 = a(i + 2)
            e2 = xyz[min(max(j2, 0), n - 1)] - xyz[a(i + 1)]
            cut = i if side == 0 else n - 2 - i
            with np.errstate(all="ignore"):
                rD[cut] = -np.dot(e2, M) / np.linalg.norm(e2)
                c = np.cross(e1, e2)
                rA[cut] = np.dot(c, M) / np.linalg.norm(c)
            rL[cut] = -np.dot(e1, S) / np.linalg.norm(e1)
        tot[side] = (S, M)
    # left-of-anchor planar terms: total wrench about x_{dr0}, planar chain from (lengths, angles)
    F = tot[0][0] + tot[1][0] + g[dr0]
    Mref = tot[0][1] + tot[1][1]
    pos, ang_dir = np.zeros(2), 0.0
    for k in range(dr0):
        d = np.array([np.cos(ang_dir), np.sin(ang_dir)])
        pos = pos + lengths[k] * d                                 # planar atom k+1
        tz = Mref[2] + ((xyz[dr0][0] - pos[0]) * F[1] - (xyz[dr0][1] - pos[1]) * F[0])
        rA[k] += -tz if k & 1 else tz
        rL[k] += d[0] * F[0] + d[1] * F[1]
        ang_dir += -((-1) ** k) * (np.pi - angles[k])              # turn at atom k+1
    gD = np.array([rD[d] if d < dr0 else rD[d + 2] for d in range(n - 3)])
    gA = np.array([rA[j] if j < dr0 else rA[j + 1] for j in range(n - 2)])
    gL = np.array([rL[k] for k in range(n - 1)])
    return gD, gA, gL


def test_moving_pivot_backward_formulas():
    """The formulas the CUDA backward implements, stated sequentially in float64, equal float64 autograd of the
    restated BackMapLayer (per-frame bond lengths so that d/d(lengths) is checked too)."""
    rng = np.random.default_rng(11)
    for n in (6, 7, 8, 9, 30, 31):
        dist = torch.from_numpy(rng.uniform(0.13, 0.15, size=(1, n - 1))).requires_grad_(True)
        ang = torch.from_numpy(rng.uniform(1.9, 2.2, size=(1, n - 2))).requires_grad_(True)
        dih = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(1, n - 3))).requires_grad_(True)
        w = rng.normal(size=(1, n, 3))
        xyz = O.back_map_layer(dist, ang, dih)
        (xyz * torch.from_numpy(w)).sum().backward()
        gD, gA, gL = _moving_pivot_vjp(xyz.detach().numpy()[0], w[0], dist.detach().numpy()[0], ang.detach().numpy()[0])
        np.testing.assert_allclose(gD, dih.grad.numpy()[0], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(gA, ang.grad.numpy()[0], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(gL, dist.grad.numpy()[0], rtol=1e-8, atol=1e-10)


# ---- start-chain VJP of dihedrals_to_cartesian: the algorithm of d2c_chain_bwd_kernel (backmap.cu) in numpy float64 ----
def _cv_frame(a, b, c):
    bc = c - b; ab = b - a
    lbc = np.linalg.norm(bc); u = bc / lbc
    nraw = np.cross(ab, u); ln = np.linalg.norm(nraw); nn = nraw / ln
    m = np.cross(nn, u)
    return u, nn, m, lbc, ln, ab

def _cv_frame_vjp(a, b, c, ub, nb, mb):
    """adjoint of (u, n, m) = _cv_frame(a, b, c)"""
    u, nn, m, lbc, ln, ab = _cv_frame(a, b, c)
    nb = nb + np.cross(u, mb)            # m = n x u
    ub = ub + np.cross(mb, nn)
    nrawb = (nb - nn * np.dot(nn, nb)) / ln
    abb = np.cross(u, nrawb)             # nraw = ab x u
    ub = ub + np.cross(nrawb, ab)
    bcb = (ub - u * np.dot(u, ub)) / lbc
    return -abb, abb - bcb, bcb          # a, b, c

def _cv_place_vjp(a, b, c, d, db):
    """d = place(a,b,c; L,theta,Delta): adjoint w.r.t. (a,b,c) and the three scalars, everything read off the FINAL points"""
    u, nn, m, _, _, _ = _cv_frame(a, b, c)
    w = d - c; L = np.linalg.norm(w)
    ct = -np.dot(w, u) / L; st = np.sqrt(max(0.0, 1 - ct * ct))
    x, y = np.dot(w, m), np.dot(w, nn)
    cd, sd = x / (L * st), y / (L * st)
    Lb = np.dot(db, w) / L
    tb = np.dot(db, L * (st * u + ct * (cd * m + sd * nn)))
    Db = np.dot(db, L * st * (-sd * m + cd * nn))
    ub = -L * ct * db; mb = L * st * cd * db; nb = L * st * sd * db
    ab_, bb_, cb_ = _cv_frame_vjp(a, b, c, ub, nb, mb)
    return ab_, bb_, cb_ + db, Lb, tb, Db

def _cv_internal_vjp(a, b, c, d, Lb, tb, Db):
    """adjoint of (L, theta, Delta)(a,b,c,d) with the conventions of place()"""
    u, nn, m, _, _, _ = _cv_frame(a, b, c)
    w = d - c; L = np.linalg.norm(w)
    q = -np.dot(w, u) / L; st = np.sqrt(max(1e-300, 1 - q * q))
    x, y = np.dot(w, m), np.dot(w, nn)
    wb = Lb * w / L
    qb = -tb / st
    wb = wb + qb * (-u / L + np.dot(w, u) * w / L ** 3)
    ub = qb * (-w / L)
    r2 = x * x + y * y
    xb, yb = -y / r2 * Db, x / r2 * Db
    wb = wb + xb * m + yb * nn
    mb = xb * w; nb = yb * w
    ab_, bb_, cb_ = _cv_frame_vjp(a, b, c, ub, nb, mb)
    return ab_, bb_, cb_ - wb, wb

def _cv_one_way(start, final, g):
    """grad w.r.t. the start chain of one one-way build (float64): start, final, g are (m,3) in chain order"""
    m_ = start.shape[0]
    sbar = np.zeros_like(start)
    acc = np.zeros_like(start)           # propagated adjoints of the final points
    for k in range(m_ - 1, 2, -1):
        db = g[k] + acc[k]
        ab_, bb_, cb_, Lb, tb, Db = _cv_place_vjp(final[k - 3], final[k - 2], final[k - 1], final[k], db)
        acc[k - 3] += ab_; acc[k - 2] += bb_; acc[k - 1] += cb_
        sa, sb, sc, sd = _cv_internal_vjp(start[k - 3], start[k - 2], start[k - 1], start[k], Lb, tb, Db)
        sbar[k - 3] += sa; sbar[k - 2] += sb; sbar[k - 1] += sc; sbar[k] += sd
    for k in range(3):
        sbar[k] += g[k] + acc[k]
    return sbar

def _start_chain_vjp(start, final, g, one_way):
    n = start.shape[0]
    if one_way:
        return _cv_one_way(start, final, g)
    s = n // 2
    out = np.zeros_like(start)
    li = np.arange(s + 1, -1, -1)                       # left chain order
    out[li] += _cv_one_way(start[li], final[li], g[li])
    ri = np.arange(s - 1, n)
    gr = g[ri].copy(); gr[:3] = 0.0                     # the right build's first three outputs are dropped
    out[ri] += _cv_one_way(start[ri], final[ri], gr)
    return out



def test_start_chain_vjp_prototype():
    """Twisting about bonds keeps lengths and angles and shifts dihedrals, so the build is a placement recursion on the
    start chain's internal coordinates; its reverse mode (what the CUDA kernel runs) equals float64 autograd of the
    restated reference for the one-way and the two-sided op."""
    rng = np.random.default_rng(0)
    for n, one_way in ((4, True), (6, True), (9, True), (4, False), (5, False), (12, False), (13, False), (40, False)):
        b = 2
        start = O.straight_tetrahedral_chain(n).astype(np.float64)[None] + rng.normal(scale=0.05, size=(b, n, 3))
        dih = rng.uniform(-np.pi, np.pi, size=(b, n - 3))
        w = rng.normal(size=(b, n, 3))
        st = torch.from_numpy(start).requires_grad_(True)
        fn = (lambda d, c: O.dihedral_to_cartesian_one_way(d, c)) if one_way else O.dihedrals_to_cartesian_tf1
        out = fn(torch.from_numpy(dih), st)
        (out * torch.from_numpy(w)).sum().backward()
        got = np.stack([_start_chain_vjp(start[f], out.detach().numpy()[f], w[f], one_way) for f in range(b)])
        np.testing.assert_allclose(got, st.grad.numpy(), rtol=1e-9, atol=1e-11)
