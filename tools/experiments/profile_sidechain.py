import sys, numpy as np, torch
sys.path.insert(0, '.')
from encodermap_b200 import _ops
gold = np.load('tests/golden/sidechains.npz')
counts = [int(c) for c in gold['ub_like_counts']]
dev = torch.device('cuda:0')
plan = _ops.SidechainPlan(counts, dev)
rng = np.random.default_rng(0)
frames = 256
n_res, n_side = len(counts), sum(c + 1 for c in counts if c > 0)
inp = [rng.uniform(0.13, 0.16, size=(frames, 3 * n_res - 1)), rng.uniform(1.85, 2.25, size=(frames, 3 * n_res - 2)),
       rng.uniform(-np.pi, np.pi, size=(frames, 3 * n_res - 3)), rng.uniform(0.13, 0.19, size=(frames, n_side)),
       rng.uniform(1.80, 2.20, size=(frames, n_side)), rng.uniform(-np.pi, np.pi, size=(frames, sum(counts)))]
inp = [torch.as_tensor(v.astype(np.float32), device=dev) for v in inp]
g = torch.randn(frames, plan.n_atoms, 3, device=dev)
for _ in range(3):
    _ops.sidechain_backmap_raw(plan, inp)
    _ops.sidechain_backmap_bwd_raw(plan, inp, g)
torch.cuda.synchronize()
