import math, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from encodermap_b200 import _lib, _ops
dev = torch.device("cuda:0")
_lib.set_option("backmap_fwd6_min_batch", 0)
for n, b in ((1500, 1 << 16), (300, 1 << 16), (3000, 8192)):
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
    for ext in (0, 8, 16, 32, 64):
        _lib.set_option("backmap_fwd6_f32_extent_nm", ext)
        for _ in range(2): out = _ops.backmap_raw(lengths, ang, dih)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = _ops.backmap_raw(lengths, ang, dih)
        e1.record(); torch.cuda.synchronize()
        mid = out[:, n // 2][:, None]
        extent = (out - mid).norm(dim=2).max(dim=1).values
        print(f"n={n} b={b} ext_limit={ext:2d}: {e0.elapsed_time(e1)/5:.3f} ms   frames with extent > 16 nm: {(extent > 16).float().mean().item():.4f}  median extent {extent.median().item():.1f}")
