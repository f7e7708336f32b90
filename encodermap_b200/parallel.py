"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for rendezvous, libemk's own NCCL communicator
(``emk_comm_*``, include/emk.h) for the exchange steps of the hot path.

Only two paths shard:

* the full-set sigmoid cost -- inputs replicated, the upper-triangular pair-tile list cut into ``world`` contiguous
  equal-count ranges (``emk_pair_tile_range``), afterwards ONE fused NCCL launch that sums the float64 loss and the
  (n, latent) float32 gradient (``emk_comm_allreduce``);
* back-mapping -- contiguous frame ranges, no communication (mean bond lengths are passed replicated).

Inside data-parallel training every rank owns n/G rows of the batch (section 8e row 2): ``data_parallel_sigmoid_cost``
all-gathers the rows (one fused launch for high-d and latent rows), evaluates this rank's slice of the pair tiles of the
GLOBAL batch, and finishes with one fused {all-reduce loss, reduce-scatter dL/dz}.

Backends: NCCL on GPUs (``init_comm`` builds libemk's communicator from a unique id shipped through the torch process
group; without it the same collectives go through ``torch.distributed``).  ``gloo`` is supported explicitly for the CPU
tests of this host logic (it has no reduce-scatter: all-reduce + slice); ``partial_fn`` lets those tests inject a
per-rank evaluator.  Any other failure propagates -- nothing is caught and papered over.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

_COMM = {"world": 0, "rank": -1}


def tile_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    return _lib.pair_tile_range(n_rows, rank, world)


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


# ---------------------------------------------------------------------------------------------------
# libemk's communicator
# ---------------------------------------------------------------------------------------------------
def init_comm(group=None, force_single: bool = False) -> bool:
    """Create libemk's NCCL communicator over the ranks of the DEFAULT process group (collective; the CUDA device of
    this rank must be current).  Returns True when the communicator is up.  A no-op for a single rank unless
    ``force_single`` (tests: a one-rank communicator still runs every collective through NCCL)."""
    if not dist.is_initialized():
        return False
    if group is not None and group is not dist.group.WORLD:
        raise ValueError("libemk's communicator spans the default process group; pass sub-groups to the collectives instead")
    world, rank = dist.get_world_size(), dist.get_rank()
    if _COMM["world"] == world:
        return True
    if world == 1 and not force_single:
        return False
    if dist.get_backend() != "nccl":
        raise RuntimeError(f"init_comm needs the nccl backend (got {dist.get_backend()!r})")
    L = _lib.lib()
    uid = (ctypes.c_char * 128)()
    if rank == 0:
        _lib.check(L.emk_comm_unique_id(uid))
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    with torch.cuda.device(torch.cuda.current_device()):
        _lib.check(L.emk_comm_init(rank, world, box[0]))
    _COMM.update(world=world, rank=rank)
    return True


def destroy_comm() -> None:
    if _COMM["world"]:
        _lib.check(_lib.lib().emk_comm_destroy())
        _COMM.update(world=0, rank=-1)


def _use_emk_comm(t: torch.Tensor, group) -> bool:
    return (_COMM["world"] > 0 and t.is_cuda and (group is None or group is dist.group.WORLD)
            and dist.get_world_size(group) == _COMM["world"])


def _is_gloo(group) -> bool:
    return dist.get_backend(group) == "gloo"


def allreduce_cost(loss: torch.Tensor, grad: Optional[torch.Tensor], group=None):
    """Sum the per-rank partial (loss float64[1], grad (n,l) float32) of one evaluation, in place.  512 KB at N = 65 536:
    latency-bound, so the two tensors travel in one launch when libemk's communicator is up."""
    if _use_emk_comm(loss, group):
        g = grad if grad is not None else None
        if g is not None and (not g.is_contiguous() or g.dtype != torch.float32):
            raise _lib.EmkError(-5, "allreduce_cost: the gradient must be a contiguous float32 tensor")
        with torch.cuda.device(loss.device):
            _lib.check(_lib.lib().emk_comm_allreduce(ctypes.c_void_p(loss.data_ptr()),
                                                     ctypes.c_void_p(g.data_ptr()) if g is not None else None,
                                                     g.numel() if g is not None else 0, _lib.stream_of(loss)))
        return loss, grad
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


# ---------------------------------------------------------------------------------------------------
# per-batch cost inside data-parallel training
# ---------------------------------------------------------------------------------------------------
def _gather_rows(high_local: torch.Tensor, low_local: torch.Tensor, group):
    world = dist.get_world_size(group)
    rows = high_local.shape[0]
    high = torch.empty((rows * world, high_local.shape[1]), dtype=high_local.dtype, device=high_local.device)
    low = torch.empty((rows * world, low_local.shape[1]), dtype=low_local.dtype, device=low_local.device)
    hl, ll = high_local.contiguous(), low_local.detach().contiguous()
    if _use_emk_comm(hl, group) and hl.dtype == torch.float32 and ll.dtype == torch.float32:
        with torch.cuda.device(hl.device):
            _lib.check(_lib.lib().emk_comm_allgather2(ctypes.c_void_p(hl.data_ptr()), ctypes.c_void_p(high.data_ptr()), hl.numel(),
                                                      ctypes.c_void_p(ll.data_ptr()), ctypes.c_void_p(low.data_ptr()), ll.numel(),
                                                      _lib.stream_of(hl)))
    else:
        dist.all_gather_into_tensor(high, hl, group=group)
        dist.all_gather_into_tensor(low, ll, group=group)
    return high, low


def _reduce_cost_scatter(loss: torch.Tensor, grad: torch.Tensor, rows: int, group):
    """loss <- sum over ranks; returns this rank's `rows` rows of the summed gradient."""
    rank = dist.get_rank(group)
    mine = torch.empty((rows, grad.shape[1]), dtype=grad.dtype, device=grad.device)
    grad = grad.contiguous()
    if _use_emk_comm(grad, group) and grad.dtype == torch.float32:
        with torch.cuda.device(grad.device):
            _lib.check(_lib.lib().emk_comm_reduce_cost_scatter(ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(grad.data_ptr()),
                                                               ctypes.c_void_p(mine.data_ptr()), mine.numel(), _lib.stream_of(grad)))
        return loss, mine
    dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    if _is_gloo(group):   # gloo implements no reduce-scatter: all-reduce and keep this rank's rows
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        mine.copy_(grad[rank * rows:(rank + 1) * rows])
    else:
        dist.reduce_scatter_tensor(mine, grad, op=dist.ReduceOp.SUM, group=group)
    return loss, mine


class _DataParallelCost(torch.autograd.Function):
    """Per-batch cost inside data-parallel training: every rank owns n/G rows of the
    batch.  One exchange each way: all-gather the high-d rows and the latent, evaluate this rank's slice of the
    pair tiles of the FULL batch, sum loss and dL/dz over ranks, keep the gradient rows this rank owns."""

    @staticmethod
    def forward(ctx, high_local, low_local, periodicity, sig, group, partial_fn, grad_scale):
        from . import _ops

        _ops._reject_high_grad(ctx.needs_input_grad[0])
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        rows = high_local.shape[0]
        high, low = _gather_rows(high_local, low_local, group)
        tr = tile_range(rows * world, rank, world)
        if partial_fn is None:
            loss, grad = _ops.sigmoid_cost_raw(high, low, periodicity, sig, tr, True)
        else:
            loss, grad = partial_fn(high, low, periodicity, sig, tr)
        loss, mine = _reduce_cost_scatter(loss, grad, rows, group)
        ctx.save_for_backward(mine)
        ctx.grad_scale = grad_scale
        ctx.low_dtype = low_local.dtype
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (mine,) = ctx.saved_tensors
        return None, (mine * (grad_output * ctx.grad_scale)).to(ctx.low_dtype), None, None, None, None, None


def data_parallel_sigmoid_cost(high_local: torch.Tensor, low_local: torch.Tensor, periodicity: float, sig, group=None,
                               partial_fn: Optional[Callable] = None, grad_reduction: str = "mean") -> torch.Tensor:
    """Sigmoid cost of the GLOBAL batch (all ranks' rows, equal counts per rank), differentiable w.r.t. this rank's
    latent rows.  The VALUE is the global-batch cost, identical on every rank.

    ``grad_reduction`` says how the host framework combines the ranks' PARAMETER gradients afterwards:

    * ``"mean"`` (default; DistributedDataParallel, Horovod, ``tf.distribute`` all average): the local gradient
      d(cost)/d(low_local) is scaled by ``world_size``, so that after the framework's mean all-reduce the parameter
      gradient of this term equals the single-process gradient on the global batch -- the other loss terms are local
      means, which average to the global mean on their own.  Without the factor the distance cost would be
      down-weighted by ``world_size`` (``distance_cost_scale = 500`` would act like ``500 / G``).
    * ``"sum"``: the framework sums the ranks' parameter gradients; the local gradient is returned as it is."""
    if grad_reduction not in ("mean", "sum"):
        raise ValueError("grad_reduction must be 'mean' or 'sum'")
    scale = float(dist.get_world_size(group)) if grad_reduction == "mean" else 1.0
    return _DataParallelCost.apply(high_local, low_local, periodicity, tuple(sig), group, partial_fn, scale)


def average_gradients(params, group=None) -> None:
    """The data-parallel all-reduce of the dense layers' gradients (what DDP / Horovod / tf.distribute do for the host
    framework; SURVEY.md 8e last row: "replicas only"): one flat float32 buffer, one collective, mean over ranks, in place.
    Goes through libemk's communicator when it is up (capturable in a CUDA graph next to the hot ops), else through
    torch.distributed."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    world = dist.get_world_size(group)
    flat = torch._utils._flatten_dense_tensors(grads)
    if _use_emk_comm(flat, group) and flat.dtype == torch.float32:
        with torch.cuda.device(flat.device):
            _lib.check(_lib.lib().emk_comm_allreduce(None, ctypes.c_void_p(flat.data_ptr()), flat.numel(), _lib.stream_of(flat)))
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    for g, v in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(v)


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
