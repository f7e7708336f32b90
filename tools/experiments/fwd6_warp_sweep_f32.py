"""Does the float32 first pass of backmap_fwd6_kernel gain from more resident warps?  (65 536 x 1 500, no fall-backs: limit 64 nm)"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from encodermap_b200 import _lib, _ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.set_option("backmap_fwd6_min_batch", 0)
n = 1500
for b in (1 << 16, 1 << 18):
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
    for ext, warps in ((0, 16), (0, 20), (64, 12), (64, 16), (64, 20)):
        _lib.set_option("backmap_fwd6_f32_extent_nm", ext)
        _lib.set_option("backmap_fwd6_warps", warps)
        for _ in range(2):
            out = _ops.backmap_raw(lengths, ang, dih)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = _ops.backmap_raw(lengths, ang, dih)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"b={b} {'f32' if ext else 'f64'} warps={warps}: {ms:.3f} ms  {b / ms / 1e3:.1f} Mframes/s  {b * 29980 / ms / 1e6 / 6450.3:.3f} of HBM", flush=True)
