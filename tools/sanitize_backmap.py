"""Small back-mapping / pairwise workload for compute-sanitizer (memcheck, racecheck, initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_backmap.py"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import ADCParameters, _ops  # noqa: E402
from encodermap_b200.encodermap_tf1 import chain_in_plane, dihedrals_to_cartesian_tf  # noqa: E402
from encodermap_b200.models.layers import PairwiseDistances  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
for n, b in ((4, 3), (31, 7), (300, 13), (700, 5), (1500, 14)):
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).requires_grad_(True)
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).requires_grad_(True)
    w = torch.randn(b, n, 3, device=dev, generator=g)
    xyz = _ops.BackMap.apply(lengths, ang, dih)
    xyz.backward(w)
    xyz2 = _ops.BackMap.apply(lengths, ang.detach(), dih)
    xyz2.backward(w)
    c = chain_in_plane(lengths, ang)
    c.backward(w)
    d = dihedrals_to_cartesian_tf(dih, c.detach())
    d.backward(w)
    if n >= 30:
        for sel in ((1, None, 3), (None, None, None)):
            p = ADCParameters(cartesian_pwd_start=sel[0], cartesian_pwd_stop=sel[1], cartesian_pwd_step=sel[2])
            x = xyz.detach().clone().requires_grad_(True)
            if x.shape[1] // (sel[2] or 1) <= 400:
                out = PairwiseDistances(p, "pd")(x)
                out.sum().backward()
torch.cuda.synchronize()
print("sanitize workload done")
