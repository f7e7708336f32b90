"""Forward back-mapping: chunk-scan kernel (fwd5) against the lane-per-frame kernel (fwd6) over batch sizes and chain lengths."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _lib, _ops  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6450.3


def run(n, b, variant, reps=5):
    # variant 5: chunk scan (float64); 6: lane per frame, float32 first pass; 64: lane per frame, float64 chain only
    _lib.set_option("backmap_fwd6_min_batch", -1 if variant == 5 else 0)
    _lib.set_option("backmap_fwd6_f32_extent_nm", 16 if variant == 6 else 0)
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
    for _ in range(2):
        out = _ops.backmap_raw(lengths, ang, dih)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = _ops.backmap_raw(lengths, ang, dih)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpf = 4 * ((n - 2) + (n - 3)) + 12 * n
    return ms, b / ms / 1e3, b * bpf / ms / 1e6 / HBM, out


if __name__ == "__main__":
    shapes = [(1500, 1 << 16), (1500, 1 << 18), (1500, 1 << 14), (1500, 4096), (1500, 1024), (300, 1 << 16), (300, 4096), (300, 1024), (3000, 8192), (1000, 32768)]
    for n, b in shapes:
        r5 = run(n, b, 5)
        r64 = run(n, b, 64)
        r6 = run(n, b, 6)
        diff = (r5[3] - r6[3]).abs().max().item()
        print(f"n={n:5d} b={b:7d}: fwd5 {r5[0]:8.3f} ms {r5[2]:.3f} of HBM | fwd6/f64 {r64[0]:8.3f} ms {r64[1]:8.2f} Mfr/s {r64[2]:.3f} | "
              f"fwd6/f32 {r6[0]:8.3f} ms {r6[1]:8.2f} Mfr/s {r6[2]:.3f} of HBM | max |f32 - fwd5| {diff:.2e}", flush=True)
