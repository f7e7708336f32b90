"""Profiling helper: a few forward back-mapping launches of one variant.  python tools/profile_backmap_fwd.py [5|6] [frames] [atoms]"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _lib, _ops  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 6
b = int(sys.argv[2]) if len(sys.argv) > 2 else 56832       # 148 SMs x 12 tile pairs x 32 frames: exactly one wave of fwd6
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
_lib.set_option("backmap_fwd6_min_batch", 0 if variant == 6 else -1)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    e0.record()
    out = _ops.backmap_raw(lengths, ang, dih)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"variant {variant} n={n} b={b}: {ms:.3f} ms  {b / ms / 1e3:.2f} Mframes/s  {b * (4 * (2 * n - 5) + 12 * n) / ms / 1e6 / 6464.3:.3f} of HBM")
