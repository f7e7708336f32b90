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
