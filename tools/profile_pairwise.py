"""Profiling helper: PairwiseDistances forward / backward and the fused Cartesian loss at 65 536 x 100 selected atoms."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _ops  # noqa: E402

dev = torch.device("cuda:0")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device=dev).manual_seed(0)
xyz = torch.randn(b, 300, 3, device=dev, generator=g)
tgt = torch.randn(b, 300, 3, device=dev, generator=g)
go = torch.randn(b, 4950, device=dev, generator=g)
for _ in range(2):   # launch order per iteration: forward, backward (+ a torch fill), fused loss
    out = _ops.pairwise_dist_raw(xyz, False, True, 1, None, 3)
    gx = _ops.pairwise_dist_bwd_raw(xyz, go, False, True, 1, None, 3)
    _ops.cartesian_pair_loss_raw(xyz, tgt, 1, None, 3, "mean_abs", 0.0, True)
torch.cuda.synchronize()
print("ok", out.shape, gx.shape)
