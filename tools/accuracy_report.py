"""Accuracy of the CUDA path against the float64 oracle, next to the reference's own float32 arithmetic
(oracle evaluated in float32, same op order as the reference).  Prints the three distances SURVEY.md H1 asks for.
Run on a GPU box:  python tools/accuracy_report.py"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _ops  # noqa: E402
from encodermap_b200.models.layers import back_map  # noqa: E402
from oracle import em_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
pi = math.pi
print("back-mapping (random coil, bond 0.13-0.15 nm, angle 1.9-2.2 rad, dihedral U(-pi,pi)); max |dx| in nm")
print(f"{'atoms':>6} {'extent':>8} {'ours-f64':>10} {'ref32-f64':>10} {'ours-ref32':>10}")
for n, b in ((300, 8), (1500, 4)):
    rng = np.random.default_rng(n)
    dist = rng.uniform(0.13, 0.15, size=(b, n - 1)).astype(np.float32)
    ang = rng.uniform(1.9, 2.2, size=(b, n - 2)).astype(np.float32)
    dih = rng.uniform(-pi, pi, size=(b, n - 3)).astype(np.float32)
    f64 = O.back_map_layer(dist.astype(np.float64), ang.astype(np.float64), dih.astype(np.float64)).numpy()
    f32 = O.back_map_layer(dist, ang, dih).numpy().astype(np.float64)
    ours = back_map(*(torch.from_numpy(v).to(dev) for v in (dist, ang, dih))).cpu().numpy().astype(np.float64)
    print(f"{n:6d} {np.ptp(f64[..., 0]).max():8.1f} {np.abs(ours - f64).max():10.2e} {np.abs(f32 - f64).max():10.2e} {np.abs(ours - f32).max():10.2e}")

print("\nsigmoid cost (periodic, clustered data, default parameters); relative errors vs float64")
print(f"{'N':>6} {'D':>5} {'loss ours':>10} {'loss ref32':>10} {'grad ours':>10} {'grad ref32':>10}")
for n, d in ((256, 51), (1024, 256)):
    rng = np.random.default_rng(n)
    centres = rng.uniform(-pi, pi, size=(8, d))
    h = (centres[rng.integers(0, 8, n)] + rng.normal(scale=0.05, size=(n, d))).astype(np.float32)
    z = (rng.normal(size=(n, 2)) * 3).astype(np.float32)
    l64, g64 = O.sigmoid_loss_and_grad(h, z, 2 * pi, O.DEFAULT_SIG, dtype=torch.float64)
    l32, g32 = O.sigmoid_loss_and_grad(h, z, 2 * pi, O.DEFAULT_SIG, dtype=torch.float32)
    lo, go = _ops.sigmoid_cost_raw(torch.from_numpy(h).to(dev), torch.from_numpy(z).to(dev), 2 * pi, O.DEFAULT_SIG)
    rel = lambda a, b: np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) / np.linalg.norm(np.asarray(b, dtype=np.float64))  # noqa: E731
    print(f"{n:6d} {d:5d} {abs(lo.item() - l64.item()) / l64.item():10.2e} {abs(l32.item() - l64.item()) / l64.item():10.2e} "
          f"{rel(go.cpu().numpy(), g64.numpy()):10.2e} {rel(g32.numpy(), g64.numpy()):10.2e}")

print("\nback-mapping gradients d<w,xyz>/d(dihedrals), d/d(angles): norm-wise relative error vs float64 autograd of the restated")
print("reference; 'floor' = exact float64 VJP evaluated on float32-rounded coordinates (what any backward fed float32 xyz can know);")
print("'ref32' = float32 autograd of the reference's own op order")
print(f"{'atoms':>6} {'dih ours':>10} {'dih floor':>10} {'dih ref32':>10} {'ang ours':>10} {'ang ref32':>10}")
rel = lambda a, b: np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) / np.linalg.norm(np.asarray(b, dtype=np.float64))  # noqa: E731
for n, b in ((300, 4), (1500, 2)):
    rng = np.random.default_rng(100 + n)
    dist = rng.uniform(0.13, 0.15, size=(b, n - 1)).astype(np.float32)
    ang = rng.uniform(1.9, 2.2, size=(b, n - 2)).astype(np.float32)
    dih = rng.uniform(-pi, pi, size=(b, n - 3)).astype(np.float32)
    w = rng.normal(size=(b, n, 3))
    res = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        a = torch.from_numpy(ang).to(dt).requires_grad_(True)
        h = torch.from_numpy(dih).to(dt).requires_grad_(True)
        x = O.back_map_layer(torch.from_numpy(dist).to(dt), a, h)
        (x * torch.from_numpy(w).to(dt)).sum().backward()
        res[tag] = (a.grad.double().numpy(), h.grad.double().numpy(), x.detach().double().numpy())
    ag, hg = (torch.from_numpy(v).to(dev).requires_grad_(True) for v in (ang, dih))
    (back_map(torch.from_numpy(dist).to(dev), ag, hg) * torch.from_numpy(w).to(dev, torch.float32)).sum().backward()
    floor = O.dihedral_vjp_from_xyz(res["f64"][2].astype(np.float32), w)
    print(f"{n:6d} {rel(hg.grad.cpu().numpy(), res['f64'][1]):10.2e} {rel(floor, res['f64'][1]):10.2e} {rel(res['f32'][1], res['f64'][1]):10.2e} "
          f"{rel(ag.grad.cpu().numpy(), res['f64'][0]):10.2e} {rel(res['f32'][0], res['f64'][0]):10.2e}")
