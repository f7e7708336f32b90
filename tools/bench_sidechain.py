"""Back-mapping with side chains on the GPU (python tools/bench_sidechain.py): BackMapLayerWithSidechains forward and backward at
the batch sizes the model trains with and on a large batch, next to the CPU restatement of the reference's layer (oracle,
torch float64, the whole batch vectorised the way TensorFlow runs it) timed on the same topology."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from encodermap_b200 import _ops  # noqa: E402
from oracle import em_oracle as O  # noqa: E402
from _timing import eager_time, graph_time  # noqa: E402

dev = torch.device("cuda:0")
gold = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "sidechains.npz")
rng = np.random.default_rng(0)


def make_inputs(counts, frames):
    n_res, n_side = len(counts), sum(c + 1 for c in counts if c > 0)
    return [rng.uniform(0.13, 0.16, size=(frames, 3 * n_res - 1)).astype(np.float32),
            rng.uniform(1.85, 2.25, size=(frames, 3 * n_res - 2)).astype(np.float32),
            rng.uniform(-np.pi, np.pi, size=(frames, 3 * n_res - 3)).astype(np.float32),
            rng.uniform(0.13, 0.19, size=(frames, n_side)).astype(np.float32),
            rng.uniform(1.80, 2.20, size=(frames, n_side)).astype(np.float32),
            rng.uniform(-np.pi, np.pi, size=(frames, sum(counts))).astype(np.float32)]


topologies = {"ubiquitin-sized (76 residues)": [int(c) for c in gold["ub_like_counts"]]}
big = [int(c) for c in rng.integers(0, 5, size=300)]
big[0], big[-1] = 3, 0
topologies["300 residues"] = big
for name, counts in topologies.items():
    plan = _ops.SidechainPlan(counts, dev)
    print(f"{name}: {plan.n_atoms} atoms, {plan.n_ops} sequential rotations per frame")
    for frames in (256, 1024, 16384):
        inputs = [torch.as_tensor(v, device=dev) for v in make_inputs(counts, frames)]
        gout = torch.randn(frames, plan.n_atoms, 3, device=dev)
        fwd = graph_time(lambda: _ops.sidechain_backmap_raw(plan, inputs), reps=3, replays=3)
        bwd = graph_time(lambda: _ops.sidechain_backmap_bwd_raw(plan, inputs, gout), reps=3, replays=3)
        _, saved = _ops.sidechain_backmap_raw(plan, inputs, save_state=True)
        fwd_s = graph_time(lambda: _ops.sidechain_backmap_raw(plan, inputs, save_state=True), reps=3, replays=3)
        bwd_s = graph_time(lambda: _ops.sidechain_backmap_bwd_raw(plan, inputs, gout, saved=saved), reps=3, replays=3)
        print(f"  {frames:6d} frames: forward {fwd:8.3f} ms ({frames / fwd * 1e3:10.0f} frames/s)   backward incl. its own forward "
              f"{bwd:8.3f} ms   |  training: forward keeping its state {fwd_s:8.3f} ms + backward {bwd_s:8.3f} ms")
    cpu_frames = 64
    ci = [torch.tensor(v.astype(np.float64), requires_grad=True) for v in make_inputs(counts, cpu_frames)]
    t0 = time.perf_counter()
    out = O.backmap_with_sidechains(counts, ci)
    t1 = time.perf_counter()
    out.sum().backward()
    t2 = time.perf_counter()
    print(f"  CPU restatement of the layer ({torch.get_num_threads()} threads, {cpu_frames} frames): forward {(t1 - t0) * 1e3:8.1f} ms "
          f"({cpu_frames / (t1 - t0):8.0f} frames/s), backward {(t2 - t1) * 1e3:8.1f} ms")
