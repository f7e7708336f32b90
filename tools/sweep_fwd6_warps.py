"""Sweep the warps-per-CTA option of the lane-per-frame forward kernel at whole-wave batch sizes."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _lib, _ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
_lib.set_option("backmap_fwd6_min_batch", 0)
for warps in (8, 12, 14, 16, 18, 20):
    _lib.set_option("backmap_fwd6_warps", warps)
    for waves in (1, 4):
        b = 148 * (warps // 2) * 32 * waves
        g = torch.Generator(device=dev).manual_seed(1)
        lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
        ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
        dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
        for _ in range(2):
            _ops.backmap_raw(lengths, ang, dih)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _ops.backmap_raw(lengths, ang, dih)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"warps {warps} waves {waves} n={n} b={b}: {ms:.3f} ms  {b / ms / 1e3:.2f} Mframes/s  {b * (4 * (2 * n - 5) + 12 * n) / ms / 1e6 / 6464.3:.3f} of HBM", flush=True)
        del ang, dih
