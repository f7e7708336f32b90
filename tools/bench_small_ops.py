"""Achieved HBM bandwidth of the small bandwidth-bound ops at config shapes (python tools/bench_small_ops.py)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from encodermap_b200 import ADCParameters, Parameters, _lib, _ops  # noqa: E402
from encodermap_b200.misc.distances import periodic_distance  # noqa: E402
from _timing import graph_time  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6450.3   # MEASURED_PEAKS.json hbm_gbs
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=20):
    return graph_time(fn, reps)     # device time: below ~40 us an eager call from Python measures the host


def report(name, ms, nbytes):
    print(f"{name:46s} {ms * 1e3:9.1f} us  {nbytes / ms / 1e6:8.1f} GB/s  {nbytes / ms / 1e6 / HBM:6.3f} of HBM")


L = _lib.lib()
st = lambda t: _lib.stream_of(t)  # noqa: E731
for rows, d in ((4096, 1024), (65536, 1024)):
    x = (torch.rand(rows, d, device=dev, generator=g) * 2 - 1) * math.pi
    out = torch.empty(rows, 2 * d, device=dev)
    go = torch.randn(rows, 2 * d, device=dev, generator=g)
    gx = torch.empty_like(x)
    a = [_lib.DL(v) for v in (x, out, go, gx)]
    report(f"periodic_input fwd ({rows}x{d})", timeit(lambda: L.emk_dl_periodic_input(a[0], 2 * math.pi, a[1], st(x))), 12 * rows * d)
    report(f"periodic_input bwd ({rows}x{d})", timeit(lambda: L.emk_dl_periodic_input_bwd(a[0], 2 * math.pi, a[2], a[3], st(x))), 16 * rows * d)
    y = (torch.rand(rows, d, device=dev, generator=g) * 2 - 1) * math.pi
    o2 = torch.empty_like(x)
    b = [_lib.DL(v) for v in (x, y, o2)]
    report(f"periodic_distance fwd ({rows}x{d})", timeit(lambda: L.emk_dl_periodic_distance(b[0], b[1], 2 * math.pi, b[2], st(x))), 12 * rows * d)
for bsz, n, sel in ((1024, 300, (1, None, 3)), (1024, 300, (None, None, None)), (65536, 300, (1, None, 3))):
    xyz = torch.randn(bsz, n, 3, device=dev, generator=g)
    ns = len(range(*slice(*sel).indices(n)))
    npair = ns * (ns - 1) // 2
    out = torch.empty(bsz, npair, device=dev)
    go = torch.randn(bsz, npair, device=dev, generator=g)
    gx = torch.zeros_like(xyz)
    a = [_lib.DL(v) for v in (xyz, out, go, gx)]
    i = lambda v: _lib.NONE_INDEX if v is None else v  # noqa: E731
    report(f"pairwise flat fwd (b={bsz}, n_sel={ns})", timeit(lambda: L.emk_dl_pairwise_dist(a[0], i(sel[0]), i(sel[1]), i(sel[2]), 0, 1, a[1], st(xyz))), 4 * bsz * npair + 12 * bsz * ns)
    report(f"pairwise flat bwd (b={bsz}, n_sel={ns})", timeit(lambda: L.emk_dl_pairwise_dist_bwd(a[0], i(sel[0]), i(sel[1]), i(sel[2]), 0, 1, a[2], a[3], st(xyz))), 4 * bsz * npair + 24 * bsz * ns)
dist = torch.rand(65536, 1499, device=dev, generator=g)
o = torch.empty(1499, device=dev)
a = [_lib.DL(dist), _lib.DL(o)]
report("column_mean (65536x1499)", timeit(lambda: L.emk_dl_column_mean(a[0], a[1], st(dist))), 4 * dist.numel())
# fused Cartesian branch (PairwiseDistances("output") + cartesian_loss + gradient in one launch)
for bsz, n, sel in ((1024, 300, (1, None, 3)), (65536, 300, (1, None, 3))):
    xyz = torch.randn(bsz, n, 3, device=dev, generator=g)
    tgt = torch.randn(bsz, n, 3, device=dev, generator=g)
    ns = len(range(*slice(*sel).indices(n)))
    ms = timeit(lambda: _ops.cartesian_pair_loss_raw(xyz, tgt, *sel, "mean_abs", 0.0, True))
    # bytes of the unfused composition it replaces: two pair matrices written + read, their gradient written + read
    report(f"fused cartesian loss+grad (b={bsz}, n_sel={ns})", ms, bsz * (24 * ns + 12 * n))
