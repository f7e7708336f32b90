"""Quick timing of the back-mapping kernels (fwd, fwd+bwd) at config-3 and config-5 shapes."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import _ops  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6464.3


def run(n, b, reps=5):
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).requires_grad_(True)
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).requires_grad_(True)
    w = torch.randn(b, n, 3, device=dev, generator=g)
    with torch.no_grad():
        for _ in range(2):
            _ops.BackMap.apply(lengths, ang, dih)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _ops.BackMap.apply(lengths, ang, dih)
        e1.record()
        torch.cuda.synchronize()
    fwd = e0.elapsed_time(e1) / reps
    # backward kernel alone, preallocated outputs, straight through the C ABI
    from encodermap_b200 import _lib
    with torch.no_grad():
        xyz = _ops.BackMap.apply(lengths, ang, dih)
    ga, gd = torch.empty_like(ang), torch.empty_like(dih)
    args = [_lib.DL(v) for v in (lengths, ang.detach(), xyz, w, ga, gd)]

    def time_bwd(with_angles):
        def bwd():
            _lib.check(_lib.lib().emk_dl_backmap_bwd(args[0], args[1], args[2], args[3], args[4] if with_angles else None, args[5], None, _lib.stream_of(xyz)))
        bwd(); bwd()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            bwd()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    bwd_d, bwd_ad = time_bwd(False), time_bwd(True)
    bf = 4 * ((n - 2) + (n - 3)) + 12 * n
    bd = 24 * n + 4 * (n - 3)                      # dihedral-only backward: xyz + grad_xyz in, grad_dihedrals out
    bad = bd + 4 * (n - 1) + 8 * (n - 2)           # + lengths, angles in, grad_angles out
    print(f"n={n} b={b}: fwd {fwd:.3f} ms {b / fwd / 1e3:.2f} Mframes/s {b * bf / fwd / 1e6 / HBM:.3f} of HBM | "
          f"bwd(dih) {bwd_d:.3f} ms {b * bd / bwd_d / 1e6 / HBM:.3f} of HBM | bwd(ang+dih) {bwd_ad:.3f} ms {b * bad / bwd_ad / 1e6 / HBM:.3f} of HBM | "
          f"fwd+bwd(dih) {b / (fwd + bwd_d) / 1e3:.2f} Mframes/s")


for n, b in ((300, 1024), (300, 65536), (1500, 65536), (999, 32768), (3000, 8192)):
    run(n, b)


def autograd_path(n=1500, b=32768, reps=4):
    import time
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)).requires_grad_(True)
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi).requires_grad_(True)
    w = torch.randn(b, n, 3, device=dev, generator=g)
    for it in range(reps):
        ang.grad = dih.grad = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        xyz = _ops.BackMap.apply(lengths, ang, dih)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        xyz.backward(w)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"autograd n={n} b={b}: fwd {1e3 * (t1 - t0):.3f} ms  bwd {1e3 * (t2 - t1):.3f} ms")


autograd_path()


def standalone(n=1500, b=32768, reps=5):
    """chain_in_plane and dihedrals_to_cartesian on an explicit start chain (the ops the reference's eager code and tests call)."""
    from encodermap_b200.encodermap_tf1 import chain_in_plane, dihedrals_to_cartesian_tf

    g = torch.Generator(device=dev).manual_seed(2)
    lengths = (0.13 + 0.02 * torch.rand(1, n - 1, device=dev, generator=g)).contiguous()
    ang = (1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g))
    dih = ((torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn):
        fn(); fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.no_grad():
        t_chain = timeit(lambda: chain_in_plane(lengths, ang))
        chain = chain_in_plane(lengths, ang)
        t_d2c = timeit(lambda: dihedrals_to_cartesian_tf(dih, chain))
    bc = 4 * (n - 2) + 12 * n
    bd = 4 * (n - 3) + 24 * n
    print(f"standalone n={n} b={b}: chain_in_plane {t_chain:.3f} ms {b * bc / t_chain / 1e6 / HBM:.3f} of HBM | "
          f"dihedrals_to_cartesian {t_d2c:.3f} ms {b * bd / t_d2c / 1e6 / HBM:.3f} of HBM")


standalone()
standalone(300, 65536)
