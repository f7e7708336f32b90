"""Where the cycles of one backbone step of the side-chain backward pass go (needs a libemk built with the clock64 probes of
this experiment; see profiles/r02_sidechain_backmap.txt for the result)."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from encodermap_b200 import _ops, _lib
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
_, saved = _ops.sidechain_backmap_raw(plan, inp, save_state=True)
L = _lib.lib()
_ops.sidechain_backmap_bwd_raw(plan, inp, g, saved=saved)
torch.cuda.synchronize()
L.emk_debug_sc_prof(None, 1)
_ops.sidechain_backmap_bwd_raw(plan, inp, g, saved=saved)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 8)()
L.emk_debug_sc_prof(out, 0)
v = list(out)
n = v[4]
print("backbone steps timed:", n)
print("thread 0 per step: apply + reduce %.0f, barrier 1 %.0f, sums %.0f, finish %.0f, store + rest %.0f, barrier 2 %.0f" % tuple(v[q] / n for q in (0, 1, 2, 3, 5, 6)))
print("thread 32: next inverse rotation %.0f" % (v[7] / n))
