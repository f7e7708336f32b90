#!/usr/bin/env python
"""Float64 prototype of the back-mapping maths the CUDA kernels implement, checked against the
oracle (run: python tools/proto_backmap.py).  Three things are established here:

 1. forward: BackMapLayer == NeRF placement from internal coordinates, anchored on the three
    middle atoms of the planar chain (an SE(3) prefix product of LOCAL transforms);
 2. backward: dL/d(dihedral), dL/d(angle), dL/d(length) from prefix sums of force and torque
    (rigid-body hinge/twist/slide motions) -- no autodiff tape;
 3. the general rotation-scan form of dihedrals_to_cartesian on an arbitrary start chain.

Conventions are found by construction and verified numerically; nothing here is product code.
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import em_oracle as O  # noqa: E402

pi = math.pi


def planar_chain(L, theta):
    """psi_0 = 0, psi_{k+1} = psi_k - (-1)^k (pi - theta_k); c_{k+1} = c_k + L_k (cos psi_k, sin psi_k, 0)."""
    n = len(L) + 1
    psi = np.zeros(n - 1)
    for k in range(n - 2):
        psi[k + 1] = psi[k] - (-1) ** k * (pi - theta[k])
    c = np.zeros((n, 3))
    for k in range(n - 1):
        c[k + 1] = c[k] + L[k] * np.array([math.cos(psi[k]), math.sin(psi[k]), 0.0])
    return c, psi


def rot_axis(u, w):
    """Column-vector rotation by angle w about unit axis u."""
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return math.cos(w) * np.eye(3) + math.sin(w) * K + (1 - math.cos(w)) * np.outer(u, u)


def side_orders(n):
    s = n // 2
    left = list(range(s + 1, -1, -1))      # atoms s+1, s, ..., 0
    right = list(range(s - 1, n))          # atoms s-1, s, ..., n-1
    nd = n - 3
    m = nd // 2
    if nd % 2 == 0:
        dl, dr = list(range(m - 1, -1, -1)), list(range(m, nd))
    else:
        dl, dr = list(range(m, -1, -1)), list(range(m + 1, nd))
    return left, right, dl, dr


def d2c_scan(c, delta):
    """General form: out_{t(k)} = C_{k-3}(c_{t(k)}), C_i = C_{i-1} o A_i,
    A_i = rotation by +delta_i about the axis c_{t(i+1)} -> c_{t(i+2)} through c_{t(i+2)}."""
    n = len(c)
    out = c.copy()
    left, right, dl, dr = side_orders(n)
    for atoms, dih in ((left, dl), (right, dr)):
        R, tau = np.eye(3), np.zeros(3)
        for i, di in enumerate(dih):
            a, b = c[atoms[i + 1]], c[atoms[i + 2]]
            u = (b - a) / np.linalg.norm(b - a)
            Ri = rot_axis(u, delta[di])
            ti = b - Ri @ b
            R, tau = R @ Ri, R @ ti + tau
            out[atoms[i + 3]] = R @ c[atoms[i + 3]] + tau
    return out


def nerf_forward(L, theta, phi):
    """BackMapLayer as NeRF: anchor atoms s-1, s, s+1 at their planar positions, then place every
    further atom D after (A,B,C) with |CD| = L, angle(B,C,D) = theta, dihedral(A,B,C,D) = phi
    (mdtraj sign), carrying the local frame explicitly (no cross products of positions)."""
    n = len(L) + 1
    s = n // 2
    c, psi = planar_chain(L, theta)
    out = np.zeros((n, 3))
    out[s - 1 : s + 2] = c[s - 1 : s + 2]
    left, right, dl, dr = side_orders(n)
    for side, (atoms, dih) in enumerate(((left, dl), (right, dr))):
        # frame at atom t2: x along bond t1->t2, z = plane normal of (t0,t1,t2), y = z x x
        t0, t1, t2 = c[atoms[0]], c[atoms[1]], c[atoms[2]]
        x = (t2 - t1) / np.linalg.norm(t2 - t1)
        zz = np.cross(t1 - t0, t2 - t1)
        zz /= np.linalg.norm(zz)
        y = np.cross(zz, x)
        R = np.stack([x, y, zz], axis=1)
        p = t2.copy()
        for i, di in enumerate(dih):
            k_prev, k_next = atoms[i + 2], atoms[i + 3]
            bond = min(k_prev, k_next)                 # bond index between the two atoms
            ang = k_prev - 1                            # theta index: angle at atom k_prev is theta[k_prev-1]
            g = pi - theta[ang]
            w = phi[di]
            # local step: twist about x by w (mdtraj sign), then bend about z by g
            cw, sw, cg, sg = math.cos(w), math.sin(w), math.cos(g), math.sin(g)
            Rx = np.array([[1, 0, 0], [0, cw, -sw], [0, sw, cw]])
            Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
            R = R @ Rx @ Rz
            p = p + L[bond] * R[:, 0]
            out[k_next] = p
    return out


def mech_backward(L, theta, xyz, g):
    """Gradients of sum(g * xyz) w.r.t. (phi, theta, L) by rigid-body mechanics.
    F_k = sum_{j<=k} g_j, T_k = sum_{j<=k} x_j x g_j (prefix sums; suffix = total - prefix)."""
    n = len(xyz)
    s = n // 2
    c, psi = planar_chain(L, theta)
    F = np.cumsum(g, axis=0)
    T = np.cumsum(np.cross(xyz, g), axis=0)
    Ftot, Ttot = F[-1], T[-1]

    def lo(k):   # sums over atoms <= k
        return (F[k], T[k]) if k >= 0 else (np.zeros(3), np.zeros(3))

    def hi(k):   # sums over atoms >= k
        f, t = lo(k - 1)
        return Ftot - f, Ttot - t

    def torque_about(f, t, piv):
        return t - np.cross(piv, f)

    gphi = np.zeros(n - 3)
    gth = np.zeros(n - 2)
    gL = np.zeros(n - 1)
    zhat = np.array([0.0, 0.0, 1.0])
    # dihedral d involves atoms d, d+1, d+2, d+3; axis d+1 -> d+2
    left, right, dl, dr = side_orders(n)
    for d in range(n - 3):
        a, b = xyz[d + 1], xyz[d + 2]
        u = (b - a) / np.linalg.norm(b - a)
        if d in dr:      # right side: atoms >= d+3 rotate by +dphi about u (through b)
            f, t = hi(d + 3)
            gphi[d] = u @ torque_about(f, t, b)
        else:            # left side: atoms <= d rotate about the reversed axis (b -> a) through a
            f, t = lo(d)
            gphi[d] = -u @ torque_about(f, t, a)
    # angle j is at atom h = j+1
    for j in range(n - 2):
        h = j + 1
        nrm = np.cross(xyz[h] - xyz[h - 1], xyz[h + 1] - xyz[h])
        nrm /= np.linalg.norm(nrm)
        if h >= s:       # right of / at the middle atom: atoms >= h+1 hinge about nrm through x_h
            f, t = hi(h + 1)
            gth[j] = -(nrm @ torque_about(f, t, xyz[h]))
        else:            # left: everything follows the planar in-plane rotation about c_h (z axis),
            f, t = lo(h - 1)  # and atoms <= h-1 additionally hinge back about nrm through x_h
            sgn = (-1) ** j
            gth[j] = sgn * (zhat @ torque_about(Ftot, Ttot, c[h])) + (nrm @ torque_about(f, t, xyz[h]))
    # bond k joins atoms k, k+1
    for k in range(n - 1):
        bdir = (xyz[k + 1] - xyz[k]) / np.linalg.norm(xyz[k + 1] - xyz[k])
        if k >= s:       # right side (incl. bond s): atoms >= k+1 slide along the bond
            f, _ = hi(k + 1)
            gL[k] = bdir @ f
        elif k == s - 1:  # bond between the two anchored atoms s-1, s: planar shift of atoms >= s; atoms <= s-1 stay
            f, _ = hi(k + 1)
            gL[k] = bdir @ f
        else:            # left: planar shift of everything by the planar bond direction, atoms <= k slide back
            pdir = np.array([math.cos(psi[k]), math.sin(psi[k]), 0.0])
            f, _ = lo(k)
            gL[k] = pdir @ Ftot - bdir @ f
    return gphi, gth, gL


def main():
    rng = np.random.default_rng(0)
    worst = {}
    for n in (6, 7, 8, 9, 10, 11, 12, 13, 30, 31, 64, 301):
        L = rng.uniform(0.13, 0.15, size=n - 1)
        theta = rng.uniform(1.9, 2.2, size=n - 2)
        phi = rng.uniform(-pi, pi, size=n - 3)
        Lt = torch.from_numpy(L)[None].clone().requires_grad_(True)
        tht = torch.from_numpy(theta)[None].clone().requires_grad_(True)
        pht = torch.from_numpy(phi)[None].clone().requires_grad_(True)
        chain = O.chain_in_plane(Lt, tht)
        left_c, right_c = O.split_counts(n)
        ref = O.dihedrals_to_cartesian_layers(pht + pi, chain, left_c, right_c)
        gout = rng.normal(size=(n, 3))
        (ref[0] * torch.from_numpy(gout)).sum().backward()
        refx = ref[0].detach().numpy()

        c, _ = planar_chain(L, theta)
        e = {}
        e["planar"] = np.abs(c - chain[0].detach().numpy()).max()
        e["scan"] = np.abs(d2c_scan(c, phi + pi) - refx).max()
        e["nerf"] = np.abs(nerf_forward(L, theta, phi) - refx).max()
        gphi, gth, gL = mech_backward(L, theta, refx, gout)
        sc = lambda a, b: np.abs(a - b).max() / max(1e-30, np.abs(b).max())  # noqa: E731
        e["gphi"] = sc(gphi, pht.grad[0].numpy())
        e["gtheta"] = sc(gth, tht.grad[0].numpy())
        e["gL"] = sc(gL, Lt.grad[0].numpy())
        print(n, {k: f"{v:.1e}" for k, v in e.items()})
        for k, v in e.items():
            worst[k] = max(worst.get(k, 0), v)
    print("worst", {k: f"{v:.1e}" for k, v in worst.items()})
    assert all(v < 1e-8 for v in worst.values()), worst


if __name__ == "__main__":
    main()
