"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

Only two paths shard:

* the full-set sigmoid cost -- inputs replicated, the upper-triangular pair-tile list cut into
  ``world`` contiguous equal-count ranges (``emk_pair_tile_range``), one all-reduce(sum) of the
  float64 loss and the (n, latent) float32 gradient afterwards;
* back-mapping -- contiguous frame ranges, no communication (mean bond lengths are passed
  replicated).

The collective runs on NCCL over NVLink/NVSwitch on GPUs and on gloo in the CPU tests of the host
logic; ``partial_fn`` lets those tests inject a per-rank evaluator.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def tile_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    return _lib.pair_tile_range(n_rows, rank, world)


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_cost(loss: torch.Tensor, grad: Optional[torch.Tensor], group=None):
    """Sum the per-rank partial (loss, grad) of one evaluation.  512 KB at N = 65 536: latency-bound."""
    dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    if grad is not None:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return loss, grad


def tile_shard(n_rows: int, group=None):
    """(tile_range of this rank, reduce function) for SigmoidCost."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    return tile_range(n_rows, rank, world), (lambda loss, grad: allreduce_cost(loss, grad, group))


def sharded_sigmoid_cost(high: torch.Tensor, low: torch.Tensor, periodicity: float, sig, group=None,
                         partial_fn: Optional[Callable] = None):
    """Evaluate this rank's slice of the pair tiles and all-reduce.  Returns (loss float64[1], grad)."""
    from . import _ops

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tr = tile_range(int(high.shape[0]), rank, world)
    if partial_fn is None:
        loss, grad = _ops.sigmoid_cost_raw(high, low, periodicity, sig, tr, True)
    else:
        loss, grad = partial_fn(high, low, periodicity, sig, tr)
    return allreduce_cost(loss, grad, group)


class _DataParallelCost(torch.autograd.Function):
    """Per-batch cost inside data-parallel training: every rank owns n/G rows of the
    batch.  One exchange each way: all-gather the high-d rows and the latent, evaluate this rank's slice of the
    pair tiles of the FULL batch, sum loss and dL/dz over ranks, keep the gradient rows this rank owns."""

    @staticmethod
    def forward(ctx, high_local, low_local, periodicity, sig, group, partial_fn):
        from . import _ops

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        rows = high_local.shape[0]
        high = torch.empty((rows * world, high_local.shape[1]), dtype=high_local.dtype, device=high_local.device)
        low = torch.empty((rows * world, low_local.shape[1]), dtype=low_local.dtype, device=low_local.device)
        dist.all_gather_into_tensor(high, high_local.contiguous(), group=group)
        dist.all_gather_into_tensor(low, low_local.detach().contiguous(), group=group)
        tr = tile_range(rows * world, rank, world)
        if partial_fn is None:
            loss, grad = _ops.sigmoid_cost_raw(high, low, periodicity, sig, tr, True)
        else:
            loss, grad = partial_fn(high, low, periodicity, sig, tr)
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        mine = torch.empty_like(low_local)
        try:
            dist.reduce_scatter_tensor(mine, grad.contiguous(), op=dist.ReduceOp.SUM, group=group)
        except (RuntimeError, NotImplementedError):  # gloo has no reduce-scatter: all-reduce and slice
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
            mine = grad[rank * rows:(rank + 1) * rows].clone()
        ctx.save_for_backward(mine)
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (mine,) = ctx.saved_tensors
        return None, mine * grad_output, None, None, None, None


def data_parallel_sigmoid_cost(high_local: torch.Tensor, low_local: torch.Tensor, periodicity: float, sig, group=None,
                               partial_fn: Optional[Callable] = None) -> torch.Tensor:
    """Sigmoid cost of the GLOBAL batch (all ranks' rows, equal counts per rank), differentiable w.r.t. this rank's
    latent rows.  The value is the same on every rank; averaging of the dense-layer gradients is the host
    framework's usual data-parallel all-reduce."""
    return _DataParallelCost.apply(high_local, low_local, periodicity, tuple(sig), group, partial_fn)


def replicate_from_host(x_host: torch.Tensor, device: torch.device, group=None) -> torch.Tensor:
    """Device copy of a host tensor that every rank holds (the replicated input of the full-set cost), built from
    one slice per rank: each rank copies rows [r n / G, (r+1) n / G) over its own PCIe link and the slices are
    all-gathered over NVLink.  At 8 ranks this moves 1/8 of the bytes per host link (268 MB -> 34 MB at 65 536 x 1 024)
    and the exchange runs at NVSwitch speed.  Falls back to a plain copy for a single rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return x_host.to(device, non_blocking=True)
    rank = dist.get_rank(group)
    n = x_host.shape[0]
    per = (n + world - 1) // world
    out = torch.empty((per * world,) + tuple(x_host.shape[1:]), dtype=x_host.dtype, device=device)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    mine = out[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[:hi - lo].copy_(x_host[lo:hi], non_blocking=True)
    if hi - lo < per:
        mine[hi - lo:].zero_()
    dist.all_gather_into_tensor(out, mine, group=group)   # in place: `mine` is this rank's slice of `out`
    return out[:n]
