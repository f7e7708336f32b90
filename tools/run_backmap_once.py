"""One forward + backward of the back-mapping on cfg5-shaped data (profiling helper)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _ops  # noqa: E402

n, b = 1500, 1 << 15
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).requires_grad_(True)
dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).requires_grad_(True)
w = torch.randn(b, n, 3, device=dev, generator=g)
xyz = _ops.BackMap.apply(lengths, ang, dih)
xyz.backward(w)                                   # angle + dihedral gradients (use_backbone_angles=True)
xyz2 = _ops.BackMap.apply(lengths, ang.detach(), dih)
xyz2.backward(w)                                  # dihedral gradients only (the ADC default)
torch.cuda.synchronize()
print("ok", xyz.shape)
