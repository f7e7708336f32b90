"""Round-2 kernels under compute-sanitizer (memcheck, racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_round2.py
narrow-input cost, warp-per-frame PairwiseDistances forward / backward (1..4 chunks), fused Cartesian loss (one and four warps
per frame, every variant, both target kinds), cartesian_distance_loss from coordinates, generation-side atoms, lane-per-frame
back-mapping (float64 chain and the float32 first pass with fall-back), back-mapping with side chains and the atom gather."""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import ADCParameters, _lib, _ops  # noqa: E402
from encodermap_b200.loss_functions.loss_functions import cartesian_distance_loss_from_coordinates  # noqa: E402
from encodermap_b200.misc import backmapping as B  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
SIG = (0.3, 6, 6, 1, 4, 6)
for n, d, l, per in ((256, 3, 2, float("inf")), (700, 8, 3, 1.0), (65, 1, 12, 2 * math.pi)):
    x = torch.rand(n, d, device=dev, generator=g)
    z = torch.randn(n, l, device=dev, generator=g)
    _ops.sigmoid_cost_raw(x, z, per, SIG)
for b, n in ((5, 2), (7, 33), (9, 64), (6, 97), (3, 128), (700, 40)):
    x = torch.randn(b, n, 3, device=dev, generator=g)
    for sq in (False, True):
        out = _ops.pairwise_dist_raw(x, sq, True)
        _ops.pairwise_dist_bwd_raw(x, torch.randn_like(out), sq, True)
for b in (5, 3000):
    for n, sel in ((30, (1, None, 3)), (300, (1, None, 3)), (300, (None, None, None)), (600, (1, None, 3))):
        if b > 100 and n > 300:
            continue
        xyz = torch.randn(b, n, 3, device=dev, generator=g)
        tgt = torch.randn(b, n, 3, device=dev, generator=g)
        pairs = _ops.pairwise_dist_raw(tgt, False, True, *sel)
        for variant in ("mean_abs", "mean_square", "mean_norm"):
            _ops.cartesian_pair_loss_raw(xyz, tgt, *sel, variant, 0.1, True, True)
            _ops.cartesian_pair_loss_raw(xyz, pairs, *sel, variant, 0.0, True, False)
xyz = torch.randn(300, 300, 3, device=dev, generator=g)
z = torch.randn(300, 2, device=dev, generator=g).requires_grad_(True)
for sel in ((1, None, 3), (None, None, 2)):
    p = ADCParameters(cartesian_pwd_start=sel[0], cartesian_pwd_stop=sel[1], cartesian_pwd_step=sel[2], cartesian_distance_cost_scale=1.0)
    cartesian_distance_loss_from_coordinates(None, p)(xyz, z).backward()
bb = torch.randn(70, 30, 3, device=dev, generator=g)
B.backbone_with_amide_atoms(bb, np.arange(30)[::3], np.arange(30)[2::3])
B.merge_cartesians(bb, np.arange(30)[::3], np.arange(30)[2::3], B.guess_amide_H(bb, np.arange(30)[::3]), B.guess_amide_O(bb, np.arange(30)[2::3]))
_lib.set_option("backmap_fwd6_min_batch", 0)
for ext in (0, 16):
    _lib.set_option("backmap_fwd6_f32_extent_nm", ext)
    for n, b in ((16, 70), (300, 200), (1500, 45)):
        lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
        ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).contiguous()
        dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).contiguous()
        dih[0] = math.pi   # an extended chain: the float32 pass falls back
        _ops.backmap_raw(lengths, ang, dih)
# back-mapping with side chains: forward (with and without kept state), backward (both), gathered pairwise distances,
# set_dihedrals; a frame count above the grid so the frame loop of a CTA runs more than once is covered by the GPU tests
for counts, frames in (([3, 4, 0], 3), ([0, 2, 1, 4], 2), ([2, 0, 4, 1, 0, 0, 3, 2, 1, 4, 2, 0], 40)):
    plan = _ops.SidechainPlan(counts, dev)
    n_res, n_side = len(counts), sum(c + 1 for c in counts if c > 0)
    ins = [0.13 + 0.03 * torch.rand(frames, 3 * n_res - 1, device=dev, generator=g), 1.9 + 0.3 * torch.rand(frames, 3 * n_res - 2, device=dev, generator=g),
           (torch.rand(frames, 3 * n_res - 3, device=dev, generator=g) * 2 - 1) * math.pi, 0.13 + 0.05 * torch.rand(frames, n_side, device=dev, generator=g),
           1.8 + 0.4 * torch.rand(frames, n_side, device=dev, generator=g), (torch.rand(frames, sum(counts), device=dev, generator=g) * 2 - 1) * math.pi]
    xyz, saved = _ops.sidechain_backmap_raw(plan, ins, save_state=True)
    _ops.sidechain_backmap_raw(plan, ins)
    go = torch.randn_like(xyz)
    _ops.sidechain_backmap_bwd_raw(plan, ins, go, saved=saved)
    _ops.sidechain_backmap_bwd_raw(plan, ins, go, needs=(False, True, True, False, True, True))
    index = torch.as_tensor(_ops.sidechain_pairwise_indices(counts, 1, None, 3), dtype=torch.int32, device=dev)
    if int(index.max()) < plan.n_atoms:
        sel = _ops.gather_atoms_raw(xyz, index)
        _ops.gather_atoms_bwd_raw(torch.randn_like(sel), index, plan.n_atoms)
torch.cuda.synchronize()
print("sanitize workload done")
