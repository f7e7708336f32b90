import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from encodermap_b200 import _ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.rand(1024, 4950, device=dev, generator=g) * 3
z = torch.randn(1024, 2, device=dev, generator=g)
for _ in range(3): _ops.sigmoid_cost_raw(x, z, float("inf"), (4.5, 12, 6, 1, 2, 6))
torch.cuda.synchronize(); print("ok")
