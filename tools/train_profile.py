"""Eager training steps of the harness for an ncu launch list (python tools/train_profile.py cfg1|adc|adc_fused)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import train_harness as H  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
g = torch.Generator(device=dev).manual_seed(7)
if which == "cfg1":
    data = (torch.rand(4096, 1024, device=dev, generator=g) * 2 - 1) * math.pi
    m = H.EncoderMapStep(1024, H.Parameters()).to(dev)
    batch = (data,)
else:
    n, b = 300, 1024
    dist = 0.13 + 0.02 * torch.rand(b, n - 1, device=dev, generator=g)
    ang = 1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)
    dih = (torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi
    with torch.no_grad():
        cart = H.BackMapLayer(n // 2 - 1, (n - 3) // 2)((dist, ang, dih))
    m = H.ADCStep(n, H.ADCParameters(cartesian_pwd_start=1, cartesian_pwd_step=3, use_backbone_angles=True), fused_cartesian=(which == "adc_fused")).to(dev)
    batch = (ang, dih, cart, dist)
print(H.time_steps(m, lambda it: batch, steps=3, warmup=2))
