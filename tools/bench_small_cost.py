"""Per-batch sigmoid cost at training-step sizes (python tools/bench_small_cost.py); EMK_CLUSTER=1 disables the K-split."""
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from encodermap_b200 import _lib, _ops  # noqa: E402
from _timing import eager_time, graph_time  # noqa: E402

dev = torch.device("cuda:0")
SIG = (4.5, 12, 6, 1, 2, 6)
for n, d, per in ((256, 3, float("inf")), (256, 1024, 2 * math.pi), (512, 1024, 2 * math.pi), (1024, 1024, 2 * math.pi),
                  (1024, 4950, float("inf")), (1024, 44850, float("inf")), (2048, 1024, 2 * math.pi), (3072, 1024, 2 * math.pi), (4096, 1024, 2 * math.pi), (8192, 1024, 2 * math.pi)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = (torch.rand(n, d, device=dev, generator=g) * 2 - 1) * math.pi
    z = torch.randn(n, 2, device=dev, generator=g)
    ms_eager = eager_time(lambda: _ops.sigmoid_cost_raw(x, z, per, SIG))
    ms = graph_time(lambda: _ops.sigmoid_cost_raw(x, z, per, SIG))     # device time (incl. the two output memsets)
    picked_small = n <= 1024 and 8 < d <= 2048
    _lib.set_option("cost_small_tile_max_rows", 0 if picked_small else 1 << 20)
    ms_other = graph_time(lambda: _ops.sigmoid_cost_raw(x, z, per, SIG))   # the tile shape the dispatcher did NOT pick
    _lib.set_option("cost_small_tile_max_rows", 1024)
    pairs = n * (n + 1) / 2
    instr = pairs * ((4 if per < 1e30 else 2) * d + 60)
    print(f"EMK_CLUSTER={os.environ.get('EMK_CLUSTER', 'auto'):>4} n={n:5d} d={d:5d} {'periodic' if per < 1e30 else 'euclid  '}: {ms * 1e3:8.1f} us (graph replay; eager from Python {ms_eager * 1e3:6.1f} us; "
          f"other tile shape {ms_other * 1e3:8.1f} us)  {pairs / ms / 1e6:8.2f} Gpairs/s  {instr / (ms * 1e-3) / (148 * 128 * 1.965e9):.3f} of FP32 issue roofline")
