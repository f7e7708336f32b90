"""world_size-2 gloo test of the multi-GPU host logic (tile-range partition + all-reduce + frame
sharding).  The per-rank evaluator is the oracle restricted to the rank's tiles -- the CUDA kernel
cannot run here -- so what is tested is the plumbing in encodermap_b200/parallel.py."""
import math
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import em_oracle as O

SIG = (4.5, 12, 6, 1, 2, 6)


def _oracle_partial(high, low, periodicity, sig, tile_range):
    """Oracle evaluation of libemk's 128x64 tile list slice [begin, end)."""
    from encodermap_b200 import _lib

    with torch.enable_grad():   # also called from inside an autograd.Function.forward, where grad mode is off
        return _oracle_partial_impl(high, low, periodicity, sig, tile_range)


def _oracle_partial_impl(high, low, periodicity, sig, tile_range):
    from encodermap_b200 import _lib

    h, z = high.double(), low.double()
    n = h.shape[0]
    sig_h, sig_l = O.sigmoid(*sig[:3]), O.sigmoid(*sig[3:])
    loss = torch.zeros(1, dtype=torch.float64)
    grad = torch.zeros_like(z)
    for t in range(*tile_range):
        i, j = _lib.pair_tile_decode(n, t)
        ri, rj = slice(i * 128, min(n, i * 128 + 128)), slice(j * 64, min(n, j * 64 + 64))
        zi, zj = z[ri].clone().requires_grad_(True), z[rj].clone().requires_grad_(True)
        dh = torch.sqrt(torch.sum(O.periodic_distance(h[ri][:, None], h[rj][None], periodicity) ** 2, dim=2))
        d2 = torch.sum((zi[:, None] - zj[None]) ** 2, dim=2)
        m = (d2 == 0).double()
        dl = torch.sqrt(d2 + m) * (1 - m)
        w = 1.0 if j // 2 == i else 2.0
        part = w * torch.sum((sig_h(dh) - sig_l(dl)) ** 2) / (n * n)
        gi, gj = torch.autograd.grad(part, (zi, zj))
        loss += part.detach()
        grad[ri] += gi
        grad[rj] += gj
    return loss, grad


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from encodermap_b200 import parallel

        rng = np.random.default_rng(5)  # same data on every rank (inputs are replicated)
        n, d = 300, 12
        high = torch.from_numpy(rng.uniform(-math.pi, math.pi, size=(n, d)))
        low = torch.from_numpy(rng.normal(size=(n, 2)))
        loss, grad = parallel.sharded_sigmoid_cost(high, low, 2 * math.pi, SIG, partial_fn=_oracle_partial)
        fr = parallel.frame_range(1001, rank, world)
        # data-parallel form: every rank owns half of the rows
        rows = n // world
        zl = low[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
        dp = parallel.data_parallel_sigmoid_cost(high[rank * rows:(rank + 1) * rows], zl, 2 * math.pi, SIG, partial_fn=_oracle_partial)
        dp.backward()
        # a training step's parameter gradient: replicated weights, local-mean term + the global-batch cost, then the
        # host framework's data-parallel MEAN all-reduce (DDP / Horovod / tf.distribute).  grad_reduction="mean" must make
        # that average equal the single-process gradient on the global batch (the round-1 version was world_size too small).
        w = torch.from_numpy(np.random.default_rng(6).normal(size=(d, 2)) * 0.3).requires_grad_(True)
        xl = high[rank * rows:(rank + 1) * rows]
        zl2 = torch.tanh(xl @ w)
        (500.0 * parallel.data_parallel_sigmoid_cost(xl, zl2, 2 * math.pi, SIG, partial_fn=_oracle_partial) + (zl2 ** 2).mean()).backward()
        wg = w.grad.clone()
        dist.all_reduce(wg)
        wg /= world
        w2 = w.detach().clone().requires_grad_(True)
        zl3 = torch.tanh(xl @ w2)
        (500.0 * parallel.data_parallel_sigmoid_cost(xl, zl3, 2 * math.pi, SIG, partial_fn=_oracle_partial, grad_reduction="sum")
         + (zl3 ** 2).mean() / world).backward()
        wg_sum = w2.grad.clone()
        dist.all_reduce(wg_sum)    # a framework that SUMS the ranks' gradients (local means pre-divided by the caller)
        # host -> device replication from one slice per rank (7 rows over 2 ranks: a ragged last slice)
        xh = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        assert torch.equal(parallel.replicate_from_host(xh, torch.device("cpu")), xh)
        q.put((rank, loss.item(), grad.numpy(), parallel.tile_range(n, rank, world), fr, dp.item(), zl.grad.numpy(), wg.numpy(), wg_sum.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_cost_world2_gloo():
    from encodermap_b200 import _build

    _build.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    high = rng.uniform(-math.pi, math.pi, size=(300, 12))
    low = rng.normal(size=(300, 2))
    lref, gref = O.sigmoid_loss_and_grad(high, low, 2 * math.pi, SIG)
    # single-process training-step gradient on the global batch
    wref = torch.from_numpy(np.random.default_rng(6).normal(size=(12, 2)) * 0.3).requires_grad_(True)
    zz = torch.tanh(torch.from_numpy(high) @ wref)
    (500.0 * O.sigmoid_loss(2 * math.pi, SIG)(high, zz) + (zz ** 2).mean()).backward()
    for rank, loss, grad, tr, fr, dp_loss, dp_grad, wg, wg_sum in results:
        assert np.linalg.norm(wg - wref.grad.numpy()) <= 1e-9 * np.linalg.norm(wref.grad.numpy())
        assert np.linalg.norm(wg_sum - wref.grad.numpy()) <= 1e-9 * np.linalg.norm(wref.grad.numpy())
        np.testing.assert_allclose(loss, lref.item(), rtol=1e-9)          # every rank holds the reduced result
        assert np.linalg.norm(grad - gref.numpy()) <= 1e-8 * np.linalg.norm(gref.numpy())
        np.testing.assert_allclose(dp_loss, lref.item(), rtol=1e-6)       # float32 scalar out of the autograd op
        mine = 2 * gref.numpy()[rank * 150:(rank + 1) * 150]   # grad_reduction="mean" (default): scaled by world_size = 2
        assert np.linalg.norm(dp_grad - mine) <= 1e-6 * np.linalg.norm(mine)
    (b0, e0), (b1, e1) = results[0][3], results[1][3]
    assert b0 == 0 and e0 == b1 and abs((e0 - b0) - (e1 - b1)) <= 1
    assert results[0][4] == (0, 501) and results[1][4] == (501, 1001)
