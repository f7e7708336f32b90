"""torch.autograd.Function wrappers around the libemk entry points.

torch is plumbing here: it owns device memory, streams and the autograd graph; every number on
the hot path is produced by a kernel in libemk.so, reached through ctypes with DLPack tensors.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DL, EmkError, check, f32c, require_cuda, sig_array, stream_of


def _empty_like_shape(ref: torch.Tensor, shape, dtype=torch.float32) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=ref.device)


# ---------------------------------------------------------------------------------------------------
# fused sigmoid cost
# ---------------------------------------------------------------------------------------------------
def sigmoid_cost_raw(high: torch.Tensor, low: torch.Tensor, periodicity: float, sig: Sequence[float],
                     tile_range: Optional[Tuple[int, int]] = None, need_grad: bool = True):
    """One fused launch: returns (loss float64[1] tensor, grad (n,l) float32 or None), both partial over
    `tile_range` (default: all tiles) and already normalised by n^2."""
    require_cuda(high, "y_true")
    require_cuda(low, "y_pred")
    high, low = f32c(high), f32c(low)
    if high.dim() != 2 or low.dim() != 2:
        raise EmkError(-4, f"sigmoid cost needs rank-2 inputs, got {tuple(high.shape)} and {tuple(low.shape)}")
    n = high.shape[0]
    if tile_range is None:
        tile_range = (0, _lib.pair_tile_count(n))
    loss = torch.empty(1, dtype=torch.float64, device=high.device)
    grad = torch.empty_like(low) if need_grad else None
    flags = _lib.EMK_COST_ZERO_OUTPUTS | (0 if need_grad else _lib.EMK_COST_NO_GRAD)
    with torch.cuda.device(high.device):
        check(_lib.lib().emk_dl_sigmoid_cost(DL(high), DL(low), float(periodicity), sig_array(sig), tile_range[0], tile_range[1],
                                             DL(loss), DL(grad), flags, stream_of(high)))
    return loss, grad


_CHUNK_CACHE = {}
_SIDE_STREAMS = {}
_CACHE_LOCK = threading.Lock()   # both caches are filled lazily and may be reached from several host threads


def _reject_high_grad(needs_grad: bool) -> None:
    """The reference's sigmoid_loss is differentiable through pairwise_dist(_periodic)(y_true) as well; every caller
    on the hot path feeds input DATA there and the fused kernel does not produce that gradient.
    Dropping it silently would train a different model, so a y_true that requires grad is an error."""
    if needs_grad:
        raise EmkError(-7, "sigmoid cost: y_true requires grad, but the fused kernel only differentiates w.r.t. y_pred "
                           "(detach y_true, or build the cost from pairwise_dist(_periodic) + sigmoid, which are differentiable)")


def _row_chunk_tiles(n: int, rows_per_chunk: int):
    """[(first_row, tile_begin)] of consecutive row chunks (multiples of the kernel's 1024-row bands): tile ids are
    band-major, a tile of the band starting at row r touches rows and columns >= r only, so tiles from tile_begin on
    need nothing before first_row.  Host-side bisection over emk_pair_tile_decode, cached per (n, rows_per_chunk)."""
    key = (n, rows_per_chunk)
    with _CACHE_LOCK:
        cached = _CHUNK_CACHE.get(key)
    if cached is None:
        total = _lib.pair_tile_count(n)
        out = []
        for first_row in range(0, n, rows_per_chunk):
            block = first_row // 128        # first 128-row tile row of the chunk
            lo, hi = 0, total               # smallest tile whose tile row is >= block (tile rows are band-monotone)
            while lo < hi:
                mid = (lo + hi) // 2
                if _lib.pair_tile_decode(n, mid)[0] >= block:
                    hi = mid
                else:
                    lo = mid + 1
            out.append((first_row, lo))
        with _CACHE_LOCK:
            cached = _CHUNK_CACHE.setdefault(key, out)
    return cached


def sigmoid_cost_streamed(high_host: torch.Tensor, low: torch.Tensor, periodicity: float, sig: Sequence[float],
                          need_grad: bool = True, rows_per_chunk: int = 8192, tile_range: Optional[Tuple[int, int]] = None,
                          interleave_group=None):
    """Same result as ``sigmoid_cost_raw`` with the high-d input in (pinned) HOST memory: rows are copied to the device in
    chunks from the LAST row backwards on a side stream while the pair tiles that only need the rows already there run
    on the current stream (tile ids are band-major, so every chunk of rows unlocks one contiguous tile range).  The copy
    of a 65 536 x 1 024 input (268 MB, ~5 ms over PCIe) hides completely behind the first 16 ms of tiles.

    With a ``tile_range`` (one rank's share of a multi-GPU evaluation) only the rows that range touches are copied --
    tiles from tile t on need no row before the band of t, so the last of 8 ranks moves 35 % of the matrix -- and every
    rank streams its own rows over its own host link behind its own tiles: no exchange of inputs between GPUs at all.

    With ``interleave_group`` (a torch.distributed group of G ranks that all hold the same pinned host tensor; ``tile_range`` must
    be None) the work is split the other way round: every rank takes 1/G of the tiles of EVERY row chunk, copies 1/G of every
    chunk over its own host link and the chunk is completed by an all-gather over NVLink on the side stream -- all ranks start
    after the first chunk and the remaining copies hide behind tiles on every rank, where a contiguous tile range makes the rank
    that owns the first band wait for the whole input.  The partial results still have to be summed by the caller."""
    require_cuda(low, "y_pred")
    if high_host.is_cuda:
        return sigmoid_cost_raw(high_host, low, periodicity, sig, tile_range, need_grad)
    if high_host.dtype != torch.float32 or not high_host.is_contiguous() or high_host.dim() != 2:
        raise EmkError(-4, "streamed sigmoid cost needs a contiguous rank-2 float32 host tensor")
    low = f32c(low)
    n, d = high_host.shape
    rows_per_chunk = max(1024, (rows_per_chunk // 1024) * 1024)
    world = rank = 0
    if interleave_group is not None:
        import torch.distributed as dist

        world, rank = dist.get_world_size(interleave_group), dist.get_rank(interleave_group)
        if tile_range is not None:
            raise ValueError("interleave_group and tile_range are mutually exclusive")
        if world == 1 or n % rows_per_chunk != 0 or rows_per_chunk % world != 0 or d % 4 != 0:
            # ragged chunks cannot be all-gathered in equal slices: contiguous tile ranges, every rank streams what it needs
            tile_range = _lib.pair_tile_range(n, rank, world)
            world = 0
    if d % 4 != 0:
        # the kernel would re-pad the whole (n, d) matrix into TMA-legal scratch on every chunk call (and read rows the
        # side stream is still writing): one plain copy, then the ordinary path, which pads once
        return sigmoid_cost_raw(high_host.to(low.device, non_blocking=True), low, periodicity, sig, tile_range, need_grad)
    chunks = _row_chunk_tiles(n, rows_per_chunk)
    total = _lib.pair_tile_count(n)
    tb, te = (0, total) if tile_range is None else tile_range
    dev = low.device
    high = torch.empty((n, d), dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float64, device=dev)
    grad = torch.empty_like(low) if need_grad else None
    base_flags = 0 if need_grad else _lib.EMK_COST_NO_GRAD
    main = torch.cuda.current_stream(dev)
    with _CACHE_LOCK:
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)   # one copy stream per device, reused by every call
    side.wait_stream(main)
    first = True
    with torch.cuda.device(dev):
        if te <= tb:   # an empty share: the outputs are still defined (zero)
            loss.zero_()
            if grad is not None:
                grad.zero_()
        for idx in range(len(chunks) - 1, -1, -1):
            r0, t0 = chunks[idx]
            r1 = min(n, r0 + rows_per_chunk)
            t1 = chunks[idx + 1][1] if idx + 1 < len(chunks) else total
            if t1 <= tb or te <= tb:
                break         # every remaining chunk lies before this rank's first tile: its rows are never read
            t0, t1 = max(t0, tb), min(t1, te)
            with torch.cuda.stream(side):
                if world > 1:
                    per = (r1 - r0) // world
                    mine = high[r0 + rank * per:r0 + (rank + 1) * per]
                    mine.copy_(high_host[r0 + rank * per:r0 + (rank + 1) * per], non_blocking=True)
                    dist.all_gather_into_tensor(high[r0:r1], mine, group=interleave_group)   # in place: `mine` is this rank's slice
                else:
                    high[r0:r1].copy_(high_host[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            main.wait_event(ev)
            if world > 1:      # this rank's share of the chunk's tiles
                cnt = t1 - t0
                t0, t1 = t0 + cnt * rank // world, t0 + cnt * (rank + 1) // world
            if t1 > t0:
                flags = base_flags | (_lib.EMK_COST_ZERO_OUTPUTS if first else 0)
                check(_lib.lib().emk_dl_sigmoid_cost(DL(high), DL(low), float(periodicity), sig_array(sig), t0, t1,
                                                     DL(loss), DL(grad), flags, stream_of(low)))
                first = False
    high.record_stream(side)   # allocated on the current stream, written on the side stream
    return loss, grad


class SigmoidCostStreamed(torch.autograd.Function):
    """SigmoidCost with a pinned host tensor as the high-d input (copy overlapped with the pair tiles)."""

    @staticmethod
    def forward(ctx, high_host, low, periodicity, sig, tile_range=None, reduce_fn=None, interleave_group=None):
        _reject_high_grad(ctx.needs_input_grad[0])
        loss, grad = sigmoid_cost_streamed(high_host, low, periodicity, sig, ctx.needs_input_grad[1], tile_range=tile_range,
                                           interleave_group=interleave_group)
        if reduce_fn is not None:  # multi-GPU: sum the partial results of all ranks
            loss, grad = reduce_fn(loss, grad)
        ctx.save_for_backward(grad)
        ctx.low_dtype = low.dtype
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        g = None if grad is None else (grad * grad_output).to(ctx.low_dtype)
        return None, g, None, None, None, None, None


class SigmoidCost(torch.autograd.Function):
    """loss = mean_ij (s_h(D^h_ij) - s_l(D^l_ij))^2 ; forward and dL/d(low) come out of the same launch."""

    @staticmethod
    def forward(ctx, high, low, periodicity, sig, tile_range, reduce_fn):
        _reject_high_grad(ctx.needs_input_grad[0])
        need_grad = ctx.needs_input_grad[1]
        loss, grad = sigmoid_cost_raw(high, low, periodicity, sig, tile_range, need_grad)
        if reduce_fn is not None:  # multi-GPU: sum the partial results of all ranks
            loss, grad = reduce_fn(loss, grad)
        ctx.save_for_backward(grad)
        ctx.low_dtype = low.dtype
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        g = None if grad is None else (grad * grad_output).to(ctx.low_dtype)
        return None, g, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# raw forward / backward calls (no autograd): shared by the torch adapter below and by tf_adapter.py
# ---------------------------------------------------------------------------------------------------
def _slice_count(n: int, start, stop, step) -> int:
    return len(range(*slice(start, stop, step).indices(n)))


def _idx(v) -> int:
    return _lib.NONE_INDEX if v is None else int(v)


def pairwise_dist_raw(x: torch.Tensor, squared=False, flat=False, start=None, stop=None, step=None) -> torch.Tensor:
    require_cuda(x, "positions")
    x = f32c(x)
    if x.dim() not in (2, 3):
        raise EmkError(-4, f"pairwise_dist needs rank 2 or 3, got {tuple(x.shape)}")
    n_all = x.shape[0] if x.dim() == 2 else x.shape[1]
    b = 1 if x.dim() == 2 else x.shape[0]
    n = _slice_count(n_all, start, stop, step)
    shape = (b, n * (n - 1) // 2) if flat else (b, n, n)
    out = _empty_like_shape(x, shape)
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_pairwise_dist(DL(x), _idx(start), _idx(stop), _idx(step), int(squared), int(flat), DL(out), stream_of(x)))
    return out


def pairwise_dist_bwd_raw(x: torch.Tensor, grad_out: torch.Tensor, squared=False, flat=False, start=None, stop=None, step=None) -> torch.Tensor:
    x, grad_out = f32c(x), f32c(grad_out)
    gx = torch.zeros_like(x)   # unselected atoms (strided selection) receive zero
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_pairwise_dist_bwd(DL(x), _idx(start), _idx(stop), _idx(step), int(squared), int(flat), DL(grad_out), DL(gx), stream_of(x)))
    return gx


def pairwise_dist_periodic_raw(x: torch.Tensor, periodicity: float) -> torch.Tensor:
    require_cuda(x, "positions")
    x = f32c(x)
    assert x.dim() == 2  # the reference asserts rank 2 (encodermap/misc/distances.py:161)
    out = _empty_like_shape(x, (x.shape[0], x.shape[0]))
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_pairwise_dist_periodic(DL(x), float(periodicity), DL(out), stream_of(x)))
    return out


def pairwise_dist_periodic_bwd_raw(x: torch.Tensor, periodicity: float, out: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    x, out, grad_out = f32c(x), f32c(out), f32c(grad_out)
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_pairwise_dist_periodic_bwd(DL(x), float(periodicity), DL(out), DL(grad_out), DL(gx), stream_of(x)))
    return gx


def periodic_distance_raw(a: torch.Tensor, b: torch.Tensor, periodicity: float) -> torch.Tensor:
    a, b = f32c(require_cuda(a, "a")), f32c(require_cuda(b, "b"))
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(_lib.lib().emk_dl_periodic_distance(DL(a), DL(b), float(periodicity), DL(out), stream_of(a)))
    return out


def periodic_distance_bwd_raw(a: torch.Tensor, b: torch.Tensor, periodicity: float, grad_out: torch.Tensor):
    a, b, grad_out = f32c(a), f32c(b), f32c(grad_out)
    ga, gb = torch.empty_like(a), torch.empty_like(b)
    with torch.cuda.device(a.device):
        check(_lib.lib().emk_dl_periodic_distance_bwd(DL(a), DL(b), float(periodicity), DL(grad_out), DL(ga), DL(gb), stream_of(a)))
    return ga, gb


def sigmoid_raw(r: torch.Tensor, sig: float, a: float, b: float) -> torch.Tensor:
    r = f32c(require_cuda(r, "r"))
    out = torch.empty_like(r)
    with torch.cuda.device(r.device):
        check(_lib.lib().emk_dl_sigmoid(DL(r), float(sig), float(a), float(b), DL(out), stream_of(r)))
    return out


def sigmoid_bwd_raw(r: torch.Tensor, sig: float, a: float, b: float, grad_out: torch.Tensor) -> torch.Tensor:
    r, grad_out = f32c(r), f32c(grad_out)
    gr = torch.empty_like(r)
    with torch.cuda.device(r.device):
        check(_lib.lib().emk_dl_sigmoid_bwd(DL(r), float(sig), float(a), float(b), DL(grad_out), DL(gr), stream_of(r)))
    return gr


def periodic_input_raw(x: torch.Tensor, periodicity: float) -> torch.Tensor:
    require_cuda(x, "inputs")
    x = f32c(x)
    if x.dim() != 2:
        raise EmkError(-4, f"PeriodicInput needs rank-2 input, got {tuple(x.shape)}")
    out = _empty_like_shape(x, (x.shape[0], 2 * x.shape[1]))
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_periodic_input(DL(x), float(periodicity), DL(out), stream_of(x)))
    return out


def periodic_input_bwd_raw(x: torch.Tensor, periodicity: float, grad_out: torch.Tensor) -> torch.Tensor:
    x, grad_out = f32c(x), f32c(grad_out)
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_periodic_input_bwd(DL(x), float(periodicity), DL(grad_out), DL(gx), stream_of(x)))
    return gx


def rotation_matrix_raw(axis: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    require_cuda(axis, "axis_unit_vec")
    axis, angle = f32c(axis), f32c(angle)
    out = _empty_like_shape(axis, (axis.shape[0], 3, 3))
    with torch.cuda.device(axis.device):
        check(_lib.lib().emk_dl_rotation_matrix(DL(axis), DL(angle), DL(out), stream_of(axis)))
    return out


def _host_indices(idx):
    """index list -> (numpy int64 array kept alive by the caller, ctypes pointer, count)"""
    import numpy as np

    a = np.ascontiguousarray(np.asarray(idx.detach().cpu() if isinstance(idx, torch.Tensor) else idx, dtype=np.int64).reshape(-1))
    return a, a.ctypes.data_as(_lib.c_i64p), int(a.size)


def guess_sp2_raw(xyz: torch.Tensor, indices, angle: float, bond_length: float) -> torch.Tensor:
    """(b, n, 3) backbone, centre-atom indices -> (b, len(indices), 3) guessed atoms (reference misc/backmapping.py:1920-1941)."""
    require_cuda(xyz, "cartesians")
    xyz = f32c(xyz)
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise EmkError(-4, f"guess_sp2_atom needs (b, n_atoms, 3) coordinates, got {tuple(xyz.shape)}")
    keep, ptr, cnt = _host_indices(indices)
    out = _empty_like_shape(xyz, (xyz.shape[0], cnt, 3))
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_guess_sp2_atoms(xyz.data_ptr(), xyz.shape[0], xyz.shape[1], ptr, cnt, float(angle), float(bond_length),
                                             out.data_ptr(), stream_of(xyz)))
    return out


def merge_cartesians_raw(central: torch.Tensor, h_after, o_after, h_xyz: torch.Tensor, o_xyz: torch.Tensor) -> torch.Tensor:
    """Reference merge loop (misc/backmapping.py:1970-1990) with the two membership lists it tests against."""
    require_cuda(central, "central_cartesians")
    central, h_xyz, o_xyz = f32c(central), f32c(require_cuda(h_xyz, "H_cartesians")), f32c(require_cuda(o_xyz, "O_cartesians"))
    kh, ph, nh = _host_indices(h_after)
    ko, po, no = _host_indices(o_after)
    out = _empty_like_shape(central, (central.shape[0], central.shape[1] + h_xyz.shape[1] + o_xyz.shape[1], 3))
    with torch.cuda.device(central.device):
        check(_lib.lib().emk_merge_cartesians(central.data_ptr(), central.shape[0], central.shape[1], ph, nh, po, no, h_xyz.data_ptr(),
                                              h_xyz.shape[1], o_xyz.data_ptr(), o_xyz.shape[1], out.data_ptr(), stream_of(central)))
    return out


def backbone_amide_raw(central: torch.Tensor, h_after, o_after, h_angle: float, h_length: float, o_angle: float, o_length: float) -> torch.Tensor:
    """guess_amide_H + guess_amide_O + merge_cartesians in one launch."""
    require_cuda(central, "central_cartesians")
    central = f32c(central)
    kh, ph, nh = _host_indices(h_after)
    ko, po, no = _host_indices(o_after)
    n_out = int(_lib.lib().emk_merged_atom_count(central.shape[1], ph, nh, po, no))
    if n_out < 0:
        raise EmkError(-5, "backbone_with_amide_atoms: index outside the backbone")
    out = _empty_like_shape(central, (central.shape[0], n_out, 3))
    with torch.cuda.device(central.device):
        check(_lib.lib().emk_backbone_amide_atoms(central.data_ptr(), central.shape[0], central.shape[1], ph, nh, po, no, float(h_angle),
                                                  float(h_length), float(o_angle), float(o_length), out.data_ptr(), n_out,
                                                  stream_of(central)))
    return out


def set_dihedrals_raw(start: torch.Tensor, quads, bonds, far_sides, targets: torch.Tensor) -> torch.Tensor:
    """The rotation loop of mdtraj_backmapping (reference misc/backmapping.py:1661-1690, 1722-1745) on the GPU.
    ``start`` (n_atoms, 3) | (frames, n_atoms, 3); ``quads`` (D, 4), ``bonds`` (D, 2) integer arrays; ``far_sides`` a list of D
    integer index arrays; ``targets`` (frames, D) radians  ->  (frames, n_atoms, 3)."""
    import numpy as np

    require_cuda(start, "xyz")
    require_cuda(targets, "dihedrals")
    start, targets = f32c(start), f32c(targets)
    if start.dim() == 2:
        start = start[None]
    if start.dim() != 3 or start.shape[2] != 3 or targets.dim() != 2:
        raise EmkError(-4, f"set_dihedrals needs (n_atoms, 3) | (frames, n_atoms, 3) coordinates and (frames, D) targets, got {tuple(start.shape)}, {tuple(targets.shape)}")
    frames, d = int(targets.shape[0]), int(targets.shape[1])
    if start.shape[0] not in (1, frames):
        raise EmkError(-4, f"set_dihedrals: {start.shape[0]} start structures for {frames} frames")
    q = np.ascontiguousarray(np.asarray(quads, dtype=np.int32).reshape(-1, 4))
    bd = np.ascontiguousarray(np.asarray(bonds, dtype=np.int32).reshape(-1, 2))
    if len(q) != d or len(bd) != d or len(far_sides) != d:
        raise EmkError(-4, f"set_dihedrals: {d} target columns, {len(q)} dihedral quadruplets, {len(bd)} bonds, {len(far_sides)} far sides")
    off = np.zeros(d + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(fs) for fs in far_sides])
    far = np.ascontiguousarray(np.concatenate([np.asarray(fs, dtype=np.int32).reshape(-1) for fs in far_sides]) if d else np.zeros(0, np.int32))
    out = _empty_like_shape(start, (frames, start.shape[1], 3))
    p32 = lambda a: a.ctypes.data_as(_lib.c_i32p)  # noqa: E731
    with torch.cuda.device(start.device):
        check(_lib.lib().emk_set_dihedrals(start.data_ptr(), start.shape[0], start.shape[1], p32(q), p32(bd), p32(off), p32(far), d,
                                           targets.data_ptr(), frames, out.data_ptr(), stream_of(start)))
    return out


class SidechainPlan:
    """Host + device step table of BackMapLayerWithSidechains for one topology (``emk_sidechain_plan_create``).  ``counts[r]`` =
    number of side-chain dihedrals of residue r + 1 (``feature_description[-1]``, reference models/layers.py:234-500).  The plan
    is bound to the CUDA device that is current when it is built; without a device it can still be inspected."""

    def __init__(self, counts, device: Optional[torch.device] = None):
        import numpy as np
        import weakref

        self.counts = np.ascontiguousarray(np.asarray(counts, dtype=np.int32).reshape(-1))
        handle = C.c_void_p()
        L = _lib.lib()
        if device is not None and torch.device(device).type == "cuda":
            ctx = torch.cuda.device(device)
        else:
            import contextlib

            ctx = contextlib.nullcontext()
        with ctx:
            rc = L.emk_sidechain_plan_create(len(self.counts), self.counts.ctypes.data_as(_lib.c_i32p), C.byref(handle))
        if rc == -7:      # EMK_E_UNSUPPORTED: the descriptions the reference's constructor cannot build either
            raise ValueError(L.emk_last_error().decode(errors="replace"))
        check(rc)
        self.handle = handle
        self._finalizer = weakref.finalize(self, L.emk_sidechain_plan_destroy, handle)
        info = (C.c_int64 * 10)()
        check(L.emk_sidechain_plan_info(handle, info))
        self.n_atoms, self.n_side, self.n_ops, self.n_residues = (int(v) for v in info[:4])
        self.columns = tuple(int(v) for v in info[4:10])
        self.saved_size = int(L.emk_sidechain_saved_size(handle))
        self.device = torch.device(device) if device is not None else None

    def ops(self):
        """(n_steps, 12) int32: kind, a, b, c, d, column, lo0, hi0, lo1, hi1, 0, 0 (see include/emk.h)."""
        import numpy as np

        out = np.zeros((self.n_ops, 12), dtype=np.int32)
        check(_lib.lib().emk_sidechain_plan_ops(self.handle, out.ctypes.data_as(_lib.c_i32p)))
        return out


def _six_inputs(plan: SidechainPlan, inputs):
    if len(inputs) != 6:
        raise EmkError(-4, f"BackMapLayerWithSidechains takes six inputs, got {len(inputs)}")
    ts = [f32c(require_cuda(t, name)) for t, name in zip(inputs, ("central_distances", "central_angles", "central_dihedrals",
                                                                   "side_distances", "side_angles", "side_dihedrals"))]
    frames = ts[0].shape[0]
    for t, cols in zip(ts, plan.columns):
        if t.dim() != 2 or t.shape[0] != frames or t.shape[1] != cols:
            raise EmkError(-4, f"BackMapLayerWithSidechains: input shapes {[tuple(t.shape) for t in ts]} do not match the topology "
                               f"(columns {plan.columns})")
    return ts


def _dl_array(dls):
    arr = (C.c_void_p * len(dls))()
    for k, d in enumerate(dls):
        arr[k] = d.ptr
    return arr


def sidechain_backmap_raw(plan: SidechainPlan, inputs, save_state: bool = False):
    """-> xyz, or (xyz, saved_state) with save_state=True: the float64 block the backward pass takes over instead of repeating the
    forward pass (sin / cos of every rotation, the coordinates before rounding)."""
    ts = _six_inputs(plan, inputs)
    out = _empty_like_shape(ts[0], (ts[0].shape[0], plan.n_atoms, 3))
    saved = _empty_like_shape(ts[0], (ts[0].shape[0], plan.saved_size), torch.float64) if save_state else None
    dls = [DL(t) for t in ts]
    with torch.cuda.device(ts[0].device):
        check(_lib.lib().emk_dl_sidechain_backmap(plan.handle, _dl_array(dls), DL(out), DL(saved), stream_of(ts[0])))
    return (out, saved) if save_state else out


def sidechain_backmap_bwd_raw(plan: SidechainPlan, inputs, grad_xyz: torch.Tensor, needs=(True,) * 6, saved: Optional[torch.Tensor] = None):
    ts = _six_inputs(plan, inputs)
    g = f32c(grad_xyz)
    grads = [torch.empty_like(t) if need else None for t, need in zip(ts, needs)]
    dls, gdls = [DL(t) for t in ts], [DL(t) for t in grads]
    with torch.cuda.device(ts[0].device):
        check(_lib.lib().emk_dl_sidechain_backmap_bwd(plan.handle, _dl_array(dls), DL(g), DL(saved), _dl_array(gdls), stream_of(ts[0])))
    return grads


class SidechainBackmap(torch.autograd.Function):
    """BackMapLayerWithSidechains.call with its exact VJP (reference models/layers.py:533-843)."""

    @staticmethod
    def forward(ctx, plan, *inputs):
        ctx.plan = plan
        if any(ctx.needs_input_grad[1:]):
            out, saved = sidechain_backmap_raw(plan, inputs, save_state=True)
            ctx.save_for_backward(*inputs, saved)
        else:
            out = sidechain_backmap_raw(plan, inputs)
            ctx.save_for_backward(*inputs)
        return out

    @staticmethod
    def backward(ctx, grad_xyz):
        saved = ctx.saved_tensors[6] if len(ctx.saved_tensors) > 6 else None
        grads = sidechain_backmap_bwd_raw(ctx.plan, ctx.saved_tensors[:6], grad_xyz, ctx.needs_input_grad[1:], saved)
        return (None, *grads)


def gather_atoms_raw(xyz: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    require_cuda(xyz, "xyz")
    xyz = f32c(xyz)
    out = _empty_like_shape(xyz, (xyz.shape[0], index.shape[0], 3))
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_gather_atoms(DL(xyz), DL(index), DL(out), stream_of(xyz)))
    return out


def gather_atoms_bwd_raw(grad_out: torch.Tensor, index: torch.Tensor, n_atoms: int) -> torch.Tensor:
    require_cuda(grad_out, "grad_out")
    g = f32c(grad_out)
    gx = _empty_like_shape(g, (g.shape[0], n_atoms, 3))
    with torch.cuda.device(g.device):
        check(_lib.lib().emk_dl_gather_atoms_bwd(DL(g), DL(index), DL(gx), stream_of(g)))
    return gx


class GatherAtoms(torch.autograd.Function):
    """tf.gather(params=inputs, indices=..., axis=1) of PairwiseDistances (reference models/layers.py:1260-1265)."""

    @staticmethod
    def forward(ctx, xyz, index):
        ctx.save_for_backward(index)
        ctx.n_atoms = xyz.shape[1]
        return gather_atoms_raw(xyz, index)

    @staticmethod
    def backward(ctx, grad_out):
        (index,) = ctx.saved_tensors
        return gather_atoms_bwd_raw(grad_out, index, ctx.n_atoms), None


def sidechain_pairwise_indices(counts, start=None, stop=None, step=None):
    """Atom selection of PairwiseDistances with reconstruct_sidechains (reference models/layers.py:1188-1208), int64 numpy."""
    import numpy as np

    c = np.ascontiguousarray(np.asarray(counts, dtype=np.int32).reshape(-1))
    args = [_lib.NONE_INDEX if v is None else int(v) for v in (start, stop, step)]
    L = _lib.lib()
    n = int(L.emk_sidechain_pairwise_indices(len(c), c.ctypes.data_as(_lib.c_i32p), *args, None))
    if n < 0:
        raise EmkError(-2, "sidechain_pairwise_indices: bad arguments")
    out = np.zeros(n, dtype=np.int64)
    L.emk_sidechain_pairwise_indices(len(c), c.ctypes.data_as(_lib.c_i32p), *args, out.ctypes.data_as(_lib.c_i64p))
    return out


def column_mean_raw(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, "distances")
    x = f32c(x)
    out = _empty_like_shape(x, (x.shape[1],))
    with torch.cuda.device(x.device):
        check(_lib.lib().emk_dl_column_mean(DL(x), DL(out), stream_of(x)))
    return out


def _lengths_2d(lengths: torch.Tensor, b: int, n: int) -> torch.Tensor:
    lengths = f32c(lengths)
    if lengths.dim() == 1:
        lengths = lengths[None]
    if lengths.shape[-1] != n - 1 or lengths.shape[0] not in (1, b):
        raise EmkError(-4, f"lengths must be (1,{n - 1}) or ({b},{n - 1}), got {tuple(lengths.shape)}")
    return lengths


def backmap_raw(lengths: torch.Tensor, angles: torch.Tensor, dihedrals: torch.Tensor) -> torch.Tensor:
    """(lengths (1|b, n-1), angles (b, n-2), dihedrals (b, n-3)) -> xyz (b, n, 3); the +pi of BackMapLayer.call is inside."""
    require_cuda(angles, "angles")
    angles, dihedrals = f32c(angles), f32c(dihedrals)
    b, n = angles.shape[0], angles.shape[1] + 2
    lengths = _lengths_2d(lengths, b, n)
    xyz = _empty_like_shape(angles, (b, n, 3))
    with torch.cuda.device(angles.device):
        check(_lib.lib().emk_dl_backmap(DL(lengths), DL(angles), DL(dihedrals), DL(xyz), stream_of(angles)))
    return xyz


def backmap_bwd_raw(lengths: torch.Tensor, angles: torch.Tensor, xyz: torch.Tensor, grad_xyz: torch.Tensor,
                    need_lengths: bool = False, need_angles: bool = True, need_dihedrals: bool = True):
    """Exact VJP from the final coordinates -> (grad_lengths (1|b, n-1) | None, grad_angles | None, grad_dihedrals | None)."""
    angles, xyz, grad_xyz = f32c(angles), f32c(xyz), f32c(grad_xyz)
    b, n = xyz.shape[0], xyz.shape[1]
    lengths = _lengths_2d(lengths, b, n)
    ga = torch.empty_like(angles) if need_angles else None
    gd = _empty_like_shape(xyz, (b, n - 3)) if need_dihedrals else None
    gl = _empty_like_shape(xyz, (b, n - 1)) if need_lengths else None
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_backmap_bwd(DL(lengths), DL(angles), DL(xyz), DL(grad_xyz), DL(ga), DL(gd), DL(gl), stream_of(xyz)))
    if gl is not None and lengths.shape[0] == 1:
        gl = gl.sum(dim=0, keepdim=True)  # shared bond lengths: every frame contributes
    return gl, ga, gd


def chain_in_plane_raw(lengths: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    require_cuda(angles, "angles")
    angles = f32c(angles)
    b, n = angles.shape[0], angles.shape[1] + 2
    lengths = _lengths_2d(lengths, b, n)
    xyz = _empty_like_shape(angles, (b, n, 3))
    with torch.cuda.device(angles.device):
        check(_lib.lib().emk_dl_chain_in_plane(DL(lengths), DL(angles), DL(xyz), stream_of(angles)))
    return xyz


def chain_in_plane_bwd_raw(lengths: torch.Tensor, angles: torch.Tensor, grad_xyz: torch.Tensor, need_lengths: bool = True,
                           need_angles: bool = True):
    angles, grad_xyz = f32c(angles), f32c(grad_xyz)
    b, n = angles.shape[0], angles.shape[1] + 2
    lengths = _lengths_2d(lengths, b, n)
    ga = torch.empty_like(angles) if need_angles else None
    gl = _empty_like_shape(angles, (b, n - 1)) if need_lengths else None
    with torch.cuda.device(angles.device):
        check(_lib.lib().emk_dl_chain_in_plane_bwd(DL(lengths), DL(angles), DL(grad_xyz), DL(ga), DL(gl), stream_of(angles)))
    if gl is not None and lengths.shape[0] == 1:
        gl = gl.sum(dim=0, keepdim=True)
    return gl, ga


def d2c_raw(dihedrals: torch.Tensor, cartesian: torch.Tensor, one_way: int) -> torch.Tensor:
    require_cuda(dihedrals, "dihedrals")
    dihedrals, cartesian = f32c(dihedrals), f32c(cartesian)
    b, n = dihedrals.shape[0], dihedrals.shape[1] + 3
    xyz = _empty_like_shape(dihedrals, (b, n, 3))
    with torch.cuda.device(dihedrals.device):
        check(_lib.lib().emk_dl_dihedrals_to_cartesian(DL(dihedrals), DL(cartesian), int(one_way), DL(xyz), stream_of(dihedrals)))
    return xyz


def d2c_bwd_raw(xyz: torch.Tensor, grad_xyz: torch.Tensor, one_way: int) -> torch.Tensor:
    xyz, grad_xyz = f32c(xyz), f32c(grad_xyz)
    gd = _empty_like_shape(xyz, (xyz.shape[0], xyz.shape[1] - 3))
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_dihedrals_to_cartesian_bwd(DL(xyz), DL(grad_xyz), int(one_way), DL(gd), stream_of(xyz)))
    return gd


def d2c_chain_bwd_raw(cartesian: torch.Tensor, xyz: torch.Tensor, grad_xyz: torch.Tensor, one_way: int) -> torch.Tensor:
    """Gradient w.r.t. the START chain; a rank-2 (shared) chain receives the sum over frames, as the reference's tiling implies.
    One thread per frame in float64 (a correctness path: the models use the fused BackMapLayer op, whose gradient w.r.t. the
    planar chain never exists as a tensor)."""
    cartesian, xyz, grad_xyz = f32c(cartesian), f32c(xyz), f32c(grad_xyz)
    gc = torch.empty_like(xyz)
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_dihedrals_to_cartesian_chain_bwd(DL(cartesian), DL(xyz), DL(grad_xyz), int(one_way), DL(gc), stream_of(xyz)))
    if cartesian.dim() == 2:
        gc = gc.sum(dim=0)
    return gc


# ---------------------------------------------------------------------------------------------------
# torch.autograd adapters
# ---------------------------------------------------------------------------------------------------
class PairwiseDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, squared, flat, start, stop, step):
        out = pairwise_dist_raw(x, squared, flat, start, stop, step)
        ctx.save_for_backward(f32c(x))
        ctx.args = (squared, flat, start, stop, step)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        return pairwise_dist_bwd_raw(x, grad_out, *ctx.args), None, None, None, None, None


class PairwiseDistPeriodic(torch.autograd.Function):
    """pairwise_dist_periodic with the reference's autodiff conventions (encodermap/misc/distances.py:144-176)."""

    @staticmethod
    def forward(ctx, x, periodicity):
        out = pairwise_dist_periodic_raw(x, periodicity)
        ctx.save_for_backward(f32c(x), out)
        ctx.periodicity = float(periodicity)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, out = ctx.saved_tensors
        return pairwise_dist_periodic_bwd_raw(x, ctx.periodicity, out, grad_out), None


class PeriodicDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, periodicity):
        out = periodic_distance_raw(a, b, periodicity)
        ctx.save_for_backward(a, b)
        ctx.periodicity = float(periodicity)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        ga, gb = periodic_distance_bwd_raw(a, b, ctx.periodicity, grad_out)
        return ga, gb, None


class Sigmoid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, sig, a, b):
        out = sigmoid_raw(r, sig, a, b)
        ctx.save_for_backward(r)
        ctx.params = (float(sig), float(a), float(b))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (r,) = ctx.saved_tensors
        return sigmoid_bwd_raw(r, *ctx.params, grad_out), None, None, None


class PeriodicInputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, periodicity):
        out = periodic_input_raw(x, periodicity)
        ctx.save_for_backward(f32c(x))
        ctx.periodicity = float(periodicity)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        return periodic_input_bwd_raw(x, ctx.periodicity, grad_out), None


class BackMap(torch.autograd.Function):
    """(lengths, angles, dihedrals) -> xyz with the exact VJP from force/torque prefix sums."""

    @staticmethod
    def forward(ctx, lengths, angles, dihedrals):
        xyz = backmap_raw(lengths, angles, dihedrals)
        ctx.save_for_backward(lengths, angles, xyz)
        return xyz

    @staticmethod
    def backward(ctx, grad_xyz):
        lengths, angles, xyz = ctx.saved_tensors
        need_l, need_a, need_d = ctx.needs_input_grad
        gl, ga, gd = backmap_bwd_raw(lengths, angles, xyz, grad_xyz, need_l, need_a, need_d)
        if gl is not None:
            gl = gl.reshape(lengths.shape) if gl.numel() == lengths.numel() else gl
        return gl, ga, gd


class ChainInPlane(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lengths, angles):
        xyz = chain_in_plane_raw(lengths, angles)
        ctx.save_for_backward(lengths, angles)
        return xyz

    @staticmethod
    def backward(ctx, grad_xyz):
        lengths, angles = ctx.saved_tensors
        need_l, need_a = ctx.needs_input_grad
        gl, ga = chain_in_plane_bwd_raw(lengths, angles, grad_xyz, need_l, need_a)
        if gl is not None:
            gl = gl.reshape(lengths.shape) if gl.numel() == lengths.numel() else gl
        return gl, ga


class DihedralsToCartesian(torch.autograd.Function):
    """Arbitrary start chain, differentiable w.r.t. the dihedrals and the start chain (a rank-2 chain is shared by
    all frames, as the reference tiles it: its gradient is the sum over frames)."""

    @staticmethod
    def forward(ctx, dihedrals, cartesian, one_way):
        xyz = d2c_raw(dihedrals, cartesian, one_way)
        ctx.save_for_backward(xyz, f32c(cartesian))
        ctx.one_way = int(one_way)
        return xyz

    @staticmethod
    def backward(ctx, grad_xyz):
        xyz, cartesian = ctx.saved_tensors
        need_d, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gd = d2c_bwd_raw(xyz, grad_xyz, ctx.one_way) if need_d else None
        gc = d2c_chain_bwd_raw(cartesian, xyz, grad_xyz, ctx.one_way) if need_c else None
        return gd, gc, None


# ---------------------------------------------------------------------------------------------------
# fused Cartesian branch: pairwise distances of the selected atoms + cartesian loss (+ clash count)
# ---------------------------------------------------------------------------------------------------
CART_VARIANTS = {"mean_abs": 0, "mean_square": 1, "mean_norm": 2}


def cartesian_pair_loss_raw(xyz: torch.Tensor, target: torch.Tensor, start=None, stop=None, step=None, variant: str = "mean_abs",
                            clash_distance: float = 0.0, need_grad: bool = True, need_clashes: bool = False):
    """One launch: (loss_sum float64[1], d(loss_sum)/d(xyz) (b,n,3) | None, clashes (b) int64 | None).  ``target`` is either the
    input coordinates (b,n,3) or their flat pair distances (b, n_sel (n_sel-1)/2); ``loss_sum`` is un-normalised (sum over
    frames and pairs of |diff| / diff^2, or sum over frames of the pair-wise 2-norm)."""
    require_cuda(xyz, "cartesians")
    require_cuda(target, "target")
    xyz, target = f32c(xyz), f32c(target)
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise EmkError(-4, f"cartesian pair loss needs (b, n_atoms, 3) coordinates, got {tuple(xyz.shape)}")
    if variant not in CART_VARIANTS:
        raise ValueError(f"cartesian_cost_variant {variant} not available")
    loss = torch.zeros(1, dtype=torch.float64, device=xyz.device)
    grad = torch.empty_like(xyz) if need_grad else None
    clashes = torch.empty(xyz.shape[0], dtype=torch.int64, device=xyz.device) if need_clashes else None
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_cartesian_pair_loss(DL(xyz), _idx(start), _idx(stop), _idx(step), DL(target), CART_VARIANTS[variant],
                                                    float(clash_distance), DL(loss), DL(grad), DL(clashes), stream_of(xyz)))
    return loss, grad, clashes


class CartesianPairLoss(torch.autograd.Function):
    """mean over frames (and pairs) of the cartesian cost between the pair distances of ``xyz`` and of ``target``;
    forward value and d/d(xyz) come out of the same launch."""

    @staticmethod
    def forward(ctx, xyz, target, start, stop, step, variant):
        loss, grad, _ = cartesian_pair_loss_raw(xyz, target, start, stop, step, variant, 0.0, ctx.needs_input_grad[0])
        n_sel = _slice_count(xyz.shape[1], start, stop, step)
        count = xyz.shape[0] * (1 if variant == "mean_norm" else max(1, n_sel * (n_sel - 1) // 2))
        ctx.save_for_backward(grad)
        ctx.count = count
        ctx.dtype = xyz.dtype
        return (loss[0] / count).to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        g = None if grad is None else (grad * (grad_output / ctx.count)).to(ctx.dtype)
        return g, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# cartesian_distance_loss straight from the coordinates
# ---------------------------------------------------------------------------------------------------
def cartesian_distance_cost_raw(xyz: torch.Tensor, low: torch.Tensor, sig: Sequence[float], start=None, stop=None, step=None,
                                tile_range: Optional[Tuple[int, int]] = None, need_grad: bool = True):
    """(loss float64[1], d(loss)/d(low) | None) of the sketch-map cost between the flat pair distances of the selected atoms of
    every frame (high-d side, Euclidean) and the latent rows; the (frames, n_pairs) matrix only exists as library scratch."""
    require_cuda(xyz, "cartesians")
    require_cuda(low, "y_pred")
    xyz, low = f32c(xyz), f32c(low)
    if xyz.dim() != 3 or xyz.shape[2] != 3 or low.dim() != 2:
        raise EmkError(-4, f"cartesian distance cost needs (b, n_atoms, 3) coordinates and a rank-2 latent, got {tuple(xyz.shape)} and {tuple(low.shape)}")
    if tile_range is None:
        tile_range = (0, _lib.pair_tile_count(xyz.shape[0]))
    loss = torch.empty(1, dtype=torch.float64, device=xyz.device)
    grad = torch.empty_like(low) if need_grad else None
    flags = _lib.EMK_COST_ZERO_OUTPUTS | (0 if need_grad else _lib.EMK_COST_NO_GRAD)
    with torch.cuda.device(xyz.device):
        check(_lib.lib().emk_dl_cartesian_distance_cost(DL(xyz), _idx(start), _idx(stop), _idx(step), DL(low), sig_array(sig), tile_range[0],
                                                        tile_range[1], DL(loss), DL(grad), flags, stream_of(xyz)))
    return loss, grad


class CartesianDistanceCost(torch.autograd.Function):
    """sigmoid cost of (pair distances of the input coordinates, latent): differentiable w.r.t. the latent; the coordinates are
    input data (a ``cartesians`` that requires grad is refused, as ``y_true`` is in SigmoidCost)."""

    @staticmethod
    def forward(ctx, xyz, low, sig, start, stop, step, tile_range, reduce_fn):
        _reject_high_grad(ctx.needs_input_grad[0])
        loss, grad = cartesian_distance_cost_raw(xyz, low, sig, start, stop, step, tile_range, ctx.needs_input_grad[1])
        if reduce_fn is not None:
            loss, grad = reduce_fn(loss, grad)
        ctx.save_for_backward(grad)
        ctx.low_dtype = low.dtype
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        g = None if grad is None else (grad * grad_output).to(ctx.low_dtype)
        return None, g, None, None, None, None, None, None
