import math, sys, torch
sys.path.insert(0, ".")
from encodermap_b200 import _lib, _ops
dev = torch.device("cuda:0")
_lib.set_option("backmap_fwd6_min_batch", 0)
_lib.set_option("backmap_fwd6_f32_extent_nm", 64)
n, b = 1500, 1 << 18
g = torch.Generator(device=dev).manual_seed(1)
lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
for _ in range(3): out = _ops.backmap_raw(lengths, ang, dih)
torch.cuda.synchronize(); print("ok")
