"""Drop-in for the three hot loss factories of ``encodermap.loss_functions.loss_functions``.

Reference: encodermap/loss_functions/loss_functions.py -- ``sigmoid_loss`` :301-369,
``distance_loss`` :200-298, ``cartesian_distance_loss`` :873-944.  Signatures, closure names
(``distance_loss_func`` / ``cartesian_distance_loss_func`` -- the models look losses up by name,
models/models.py:2256-2258), scale handling and the finite-value assertion are kept; the N x N
arithmetic is one fused kernel launch (emk_sigmoid_cost) that also produces dL/d(latent).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .. import _ops
from ..parameters import ADCParameters, Parameters


class _FiniteCheck:
    """``tf.debugging.assert_all_finite`` of the reference (loss_functions.py:293-295, 364-366, 939-941) without a
    device synchronisation per step.

    ``mode`` True: read the scalar back now (one synchronisation per call, the reference's behaviour to the step);
    ``"deferred"`` (default): the finite flag of this call is copied to pinned host memory behind the kernels and
    examined when the NEXT call comes in (or by ``flush()``), so a NaN cost still stops training with the reference's
    message, one step late and at no cost to the launch pipeline; False: never.  While a CUDA graph is being captured
    nothing can be read back, so the check is skipped (the scalar still carries the NaN to whoever consumes it)."""

    def __init__(self, mode, message: str) -> None:
        if mode not in (True, False, "deferred"):
            raise ValueError("check_finite must be True, False or 'deferred'")
        self.mode, self.message = mode, message
        self._pending = None   # (pinned flag tensor, event)

    def flush(self) -> None:
        if self._pending is not None:
            flag, event = self._pending
            self._pending = None
            event.synchronize()
            if not bool(flag.item()):
                raise FloatingPointError(self.message)

    def __call__(self, x: torch.Tensor) -> None:
        if self.mode is False:
            return
        if x.is_cuda and torch.cuda.is_current_stream_capturing():
            return
        if self.mode is True or not x.is_cuda:
            if not torch.isfinite(x).all():
                raise FloatingPointError(self.message)
            return
        self.flush()
        flag = torch.empty((), dtype=torch.bool, pin_memory=True)
        flag.copy_(torch.isfinite(x.detach()).all(), non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(x.device))
        self._pending = (flag, event)


def sigmoid_loss(
    parameters=None,
    periodicity_overwrite: Optional[float] = None,
    dist_dig_parameters_overwrite: Optional[Sequence[float]] = None,
    *,
    process_group=None,
    data_parallel: bool = False,
    check_finite="deferred",
) -> Callable:
    """Sigmoid loss closure.  Reference: encodermap/loss_functions/loss_functions.py:301-369.

    Extra keyword-only arguments (not in the reference): ``process_group`` shards the pair tiles of
    one evaluation over the ranks of a torch.distributed group (inputs replicated, one all-reduce of
    the loss and dL/d(latent)); ``check_finite`` selects how the reference's finite assertion runs: ``"deferred"``
    (default -- checked one call late, no synchronisation), ``True`` (checked now, one synchronisation per call) or
    ``False``; the closure's ``flush_finite_check()`` examines the last pending flag.  ``data_parallel=True`` (with a
    ``process_group``) is the form for data-parallel TRAINING: ``y_true`` / ``y_pred`` are this rank's rows of the global
    batch, the value is the cost of the global batch and the gradient is pre-scaled for a mean-reducing framework
    (``parallel.data_parallel_sigmoid_cost``).  ``y_true`` may also be a PINNED
    host tensor: it is then copied to the device in row chunks on a side stream while the pair tiles that need only the
    rows already there run.

    Narrower than the reference in one respect: the cost is differentiable w.r.t. ``y_pred`` only.  ``y_true`` is input
    data in every caller on the hot path (models/models.py:2419-2422, loss_functions.py:277-287); a ``y_true`` that
    requires grad raises instead of silently dropping that gradient.  Any latent width works (``n_neurons[-1]``,
    parameters/parameters.py:612)."""
    p = Parameters() if parameters is None else parameters
    periodicity = periodicity_overwrite if periodicity_overwrite is not None else p.periodicity
    sig = tuple(dist_dig_parameters_overwrite) if dist_dig_parameters_overwrite is not None else tuple(p.dist_sig_parameters)
    finite = _FiniteCheck(check_finite, "Sigmoid cost became infinite or NaN.")

    def sigmoid_loss_func(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
        tile_range, reduce_fn = None, None
        if data_parallel and process_group is not None:
            from ..parallel import data_parallel_sigmoid_cost

            cost = data_parallel_sigmoid_cost(y_true, y_pred, periodicity, sig, group=process_group)
            finite(cost)
            return cost
        if process_group is not None:
            from ..parallel import tile_shard

            tile_range, reduce_fn = tile_shard(int(y_true.shape[0]), process_group)
        if not y_true.is_cuda and y_true.is_pinned():
            # high-d input still in pinned host memory (not a CPU path -- unpinned CPU tensors are rejected like everywhere
            # else).  One or two ranks: streamed to the device behind the pair tiles, every rank only the rows its own tile
            # range touches.  More ranks: the rank that owns the first tile band needs every row before its first tile, so
            # streaming would expose a full 1/1 copy on that rank; instead every rank copies 1/G of the rows over its own
            # host link and the slices are all-gathered over NVLink (parallel.replicate_from_host).
            if process_group is not None and torch.distributed.get_world_size(process_group) > 2:
                # interleaved split: every rank takes 1/G of the tiles of every row chunk, copies 1/G of every chunk over its
                # own host link and the chunk is completed by an all-gather over NVLink behind the tiles of the previous one
                cost = _ops.SigmoidCostStreamed.apply(y_true, y_pred, periodicity, sig, None, reduce_fn, process_group)
            else:
                cost = _ops.SigmoidCostStreamed.apply(y_true, y_pred, periodicity, sig, tile_range, reduce_fn)
        else:
            cost = _ops.SigmoidCost.apply(y_true, y_pred, periodicity, sig, tile_range, reduce_fn)
        finite(cost)
        return cost

    sigmoid_loss_func.flush_finite_check = finite.flush
    return sigmoid_loss_func


def _latent_of(model) -> Callable:
    enc = getattr(model, "encoder", None)
    if enc is None:
        raise Exception("model has no `encoder` attribute: cannot find the bottleneck/latent layer")
    return enc


def distance_loss(model, parameters=None, callback=None, *, process_group=None, data_parallel: bool = False,
                  check_finite="deferred") -> Callable:
    """Encodermap distance_loss.  Reference: encodermap/loss_functions/loss_functions.py:200-298.
    ``model.encoder`` maps the (tuple of) inputs to the latent; ``callback`` is accepted for signature
    compatibility (summary writing stays in the host framework)."""
    p = Parameters() if parameters is None else parameters
    latent = _latent_of(model)
    dist_loss = sigmoid_loss(p, process_group=process_group, data_parallel=data_parallel, check_finite=False)   # one check, below
    finite = _FiniteCheck(check_finite, "Dist cost became infinite or NaN.")

    def distance_loss_func(y_true, y_pred=None) -> torch.Tensor:
        distance_loss_func.name = "distance_loss"
        try:
            y_pred = latent(y_true, training=True)
        except TypeError:
            y_pred = latent(y_true)
        if isinstance(y_true, tuple):
            y_true = torch.cat(y_true[:3], dim=1)
        if p.distance_cost_scale is not None:
            dist_cost = dist_loss(y_true, y_pred) * p.distance_cost_scale
        else:
            dist_cost = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        finite(dist_cost)
        return dist_cost

    distance_loss_func.flush_finite_check = finite.flush
    return distance_loss_func


def cartesian_distance_loss(model, parameters=None, callback=None, *, process_group=None, data_parallel: bool = False,
                            check_finite="deferred") -> Callable:
    """Encodermap cartesian distance loss.  Reference: encodermap/loss_functions/loss_functions.py:873-944
    (non-periodic, ``cartesian_dist_sig_parameters``; called as ``(input pairwise distances, latent)``,
    models/models.py:2419-2422)."""
    p = ADCParameters() if parameters is None else parameters
    dist_loss = sigmoid_loss(p, periodicity_overwrite=float("inf"),
                             dist_dig_parameters_overwrite=p.cartesian_dist_sig_parameters, process_group=process_group,
                             data_parallel=data_parallel, check_finite=False)
    finite = _FiniteCheck(check_finite, "Cartesian distance cost became infinite or NaN.")

    def cartesian_distance_loss_func(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
        cartesian_distance_loss_func.name = "cartesian_distance_loss"
        if p.cartesian_distance_cost_scale is not None:
            dist_cost = dist_loss(y_true, y_pred) * p.cartesian_distance_cost_scale
        else:
            dist_cost = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        finite(dist_cost)
        return dist_cost

    cartesian_distance_loss_func.flush_finite_check = finite.flush
    return cartesian_distance_loss_func


def cartesian_distance_loss_from_coordinates(model, parameters=None, callback=None, *, process_group=None, check_finite="deferred") -> Callable:
    """``cartesian_distance_loss`` fed with the input COORDINATES instead of their stored pair distances.

    Reference composition: ``inp_pair = PairwiseDistances(p, "input")(inp_cartesians)`` (models/models.py:837-839) followed by
    ``cartesian_distance_loss(model, p)(inp_pair, latent)`` (:2419-2422, loss_functions.py:873-944).  ``f(cartesians, y_pred)``
    returns the same value; the (batch, n_pairs) matrix -- the largest activation of the ADC step, 20 MB at 1 024 x 4 950 and
    184 MB at 1 024 x 44 850 -- is library scratch that is gone when the call returns.  Atom selection: ``cartesian_pwd_start /
    stop / step`` as in PairwiseDistances.  ``process_group`` shards the pair tiles over ranks as in ``sigmoid_loss``."""
    p = ADCParameters() if parameters is None else parameters
    sig = tuple(p.cartesian_dist_sig_parameters)
    finite = _FiniteCheck(check_finite, "Cartesian distance cost became infinite or NaN.")

    def cartesian_distance_loss_func(cartesians: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
        cartesian_distance_loss_func.name = "cartesian_distance_loss"
        if p.cartesian_distance_cost_scale is not None:
            tile_range, reduce_fn = None, None
            if process_group is not None:
                from ..parallel import tile_shard

                tile_range, reduce_fn = tile_shard(int(cartesians.shape[0]), process_group)
            dist_cost = _ops.CartesianDistanceCost.apply(cartesians, y_pred, sig, p.cartesian_pwd_start, p.cartesian_pwd_stop,
                                                         p.cartesian_pwd_step, tile_range, reduce_fn) * p.cartesian_distance_cost_scale
        else:
            dist_cost = torch.zeros((), dtype=torch.float32, device=y_pred.device)
        finite(dist_cost)
        return dist_cost

    cartesian_distance_loss_func.flush_finite_check = finite.flush
    return cartesian_distance_loss_func


def fused_cartesian_loss(model=None, scale_callback=None, parameters=None, log_callback=None, *, check_finite="deferred") -> Callable:
    """The Cartesian branch of the ADC step as ONE op: ``PairwiseDistances`` on the back-mapped coordinates + ``cartesian_loss``
    (reference: encodermap/models/layers.py:1252-1267 + encodermap/loss_functions/loss_functions.py:947-1067, called as
    ``cartesian_loss_func(inp_pair, out_pair)`` in models/models.py:2385-2387).

    ``f(y_true, out_cartesians)``: ``out_cartesians`` are the BackMapLayer output (b, n_atoms, 3); ``y_true`` is either the
    input coordinates (b, n_atoms, 3) -- their pair distances are then formed on the fly -- or the stored input pair distances
    (b, n_pairs), as the reference passes them.  Atom selection (``cartesian_pwd_start/stop/step``), cost variant
    (``cartesian_cost_variant``), ``cartesian_cost_reference`` and the current scale (``scale_callback.current_cartesian_cost_scale``
    or ``cartesian_cost_scale``) follow the reference.  Neither (b, n_pairs) matrix of the output side exists in memory;
    d(cost)/d(out_cartesians) comes out of the same launch and feeds ``BackMapLayer``'s backward."""
    p = ADCParameters() if parameters is None else parameters
    finite = _FiniteCheck(check_finite, "Cartesian cost became infinite or NaN.")

    def cartesian_loss_func(y_true: torch.Tensor, out_cartesians: torch.Tensor) -> torch.Tensor:
        scale = scale_callback.current_cartesian_cost_scale if scale_callback is not None else p.cartesian_cost_scale
        cost = _ops.CartesianPairLoss.apply(out_cartesians, y_true, p.cartesian_pwd_start, p.cartesian_pwd_stop, p.cartesian_pwd_step,
                                            p.cartesian_cost_variant)
        cost = cost / getattr(p, "cartesian_cost_reference", 1) * scale
        finite(cost)
        return cost

    cartesian_loss_func.flush_finite_check = finite.flush
    return cartesian_loss_func


def clash_count(cartesians: torch.Tensor, clash_distance: float = 0.1, parameters=None) -> torch.Tensor:
    """Pairs of (selected) atoms closer than ``clash_distance`` per frame, (b) int64: the quantity ``ADCClashMetric.update_state``
    logs (reference: encodermap/callbacks/metrics.py:512-520, 0.1 nm for C-alpha selections) without the (b, n_pairs) matrix.
    The metric looks at ALL atoms the output layer produced unless ``parameters`` carries a selection."""
    start = stop = step = None
    if parameters is not None:
        start, stop, step = parameters.cartesian_pwd_start, parameters.cartesian_pwd_stop, parameters.cartesian_pwd_step
    _, _, clashes = _ops.cartesian_pair_loss_raw(cartesians, cartesians, start, stop, step, "mean_abs", clash_distance, False, True)
    return clashes
