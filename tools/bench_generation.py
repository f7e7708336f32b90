"""Generation-side back-mapping on the GPU (python tools/bench_generation.py): guessed amide atoms + merge on a large batch, and the
rotation loop of mdtraj_backmapping (set_dihedrals) on a 500-residue chain with side chains, next to the numpy restatement of
the reference's Python loop timed on two frames."""
import math
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from encodermap_b200.misc import backmapping as B  # noqa: E402
from oracle import em_oracle as O  # noqa: E402
from _timing import eager_time  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6450.3
g = torch.Generator(device=dev).manual_seed(3)
for frames, n in ((65536, 300), (65536, 1500)):
    xyz = torch.randn(frames, n, 3, device=dev, generator=g)
    n_idx, c_idx = np.arange(n)[::3], np.arange(n)[2::3]
    ms = eager_time(lambda: B.backbone_with_amide_atoms(xyz, n_idx, c_idx), reps=10)
    n_out = n + (n // 3 - 1) + n // 3
    nbytes = frames * 12 * (n + n_out)
    print(f"backbone_with_amide_atoms ({frames} x {n} -> {n_out} atoms): {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s  {nbytes / ms / 1e6 / HBM:.3f} of HBM "
          f"(incl. the host-side plan upload of every call)")

rng = np.random.default_rng(0)
n_res = 500
bonds, kind, side_quads = [], [], []
for r in range(n_res):
    base = len(kind)
    kind += ["N", "CA", "C"]
    if r:
        bonds.append((prev_c, base))
    bonds += [(base, base + 1), (base + 1, base + 2)]
    prev_c = base + 2
n_bb = len(kind)
for r in range(n_res):
    chain = [3 * r, 3 * r + 1]
    for _ in range(int(rng.integers(0, 5))):
        kind.append("S")
        bonds.append((chain[-1], len(kind) - 1))
        chain.append(len(kind) - 1)
    side_quads += [chain[k:k + 4] for k in range(len(chain) - 3)]
n_atoms = len(kind)
quads = np.vstack([np.array([[k, k + 1, k + 2, k + 3] for k in range(n_bb - 3)]), np.array(side_quads).reshape(-1, 4)])
t0 = time.perf_counter()
_, fars = B.near_and_far_sides(n_atoms, bonds, quads[:, 1:3])
t_far = time.perf_counter() - t0
start = np.cumsum(rng.normal(scale=0.09, size=(n_atoms, 3)), axis=0).astype(np.float32)
far_total = sum(len(f) for f in fars)
for frames in (64, 1024, 8192):
    targets = torch.from_numpy(rng.uniform(-math.pi, math.pi, size=(frames, len(quads))).astype(np.float32)).to(dev)
    s = torch.from_numpy(start).to(dev)
    ms = eager_time(lambda: B.set_dihedrals(s, quads, quads[:, 1:3], fars, targets), reps=3, warm=1)
    print(f"set_dihedrals ({frames} frames x {n_atoms} atoms, {len(quads)} dihedrals, {far_total} far-side atoms per frame): {ms:9.2f} ms  "
          f"{frames / ms * 1e3:10.0f} frames/s  {frames * far_total / ms / 1e6:8.2f} G atom-rotations/s")
t0 = time.perf_counter()
O.set_dihedrals(start.astype(np.float64), quads, quads[:, 1:3], fars, rng.uniform(-math.pi, math.pi, size=(2, len(quads))))
t_cpu = (time.perf_counter() - t0) / 2
print(f"numpy restatement of the reference's loop: {t_cpu:.2f} s per frame ({1 / t_cpu:.2f} frames/s, one core); far sides by BFS: {t_far:.2f} s once")
