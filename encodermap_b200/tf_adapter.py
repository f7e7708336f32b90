"""TensorFlow side of the drop-in boundary: every hot-path callable of the reference as a
``tf.custom_gradient`` function with the reference's signature, on top of the same libemk entry points the torch adapter
uses, and ``install()`` which rebinds them into the reference's modules so that ``em.EncoderMap`` /
``AngleDihedralCartesianEncoderMap`` train unchanged (their loss closures and layers call the rebound names).

TensorFlow is not installed in this image nor on the GPU box.  The module imports ``tensorflow`` lazily, so the tests can
register a torch-backed stand-in (TF's semantics for the ~30 symbols touched here and by the reference's callers; it
lives with the test infrastructure) under that name: ``tests/test_tf_adapter.py`` then runs the reference's caller bodies +
``install()`` end to end against libemk on the GPU and compares with the oracle.  With a real TensorFlow the hazards are
(SURVEY.md H6):

* Keras traces ``train_step`` into a graph, so every ctypes call sits inside ``tf.py_function``; shapes are restored with
  ``tf.reshape`` from the input shapes because ``py_function`` outputs have none;
* TF's DLPack export does not order its compute stream with anyone else's: ``_enter()`` waits for the device ONCE before
  the first libemk launch of a call, and ``_leave()`` waits for libemk's own stream only (not the device) before the
  results go back.  Both waits disappear in the production route, a TF custom op that hands TF's own ``cudaStream_t`` to
  the C ABI (INTEGRATION.md shows that binding);
* the reference hides all GPUs unless ``ENCODERMAP_ENABLE_GPU=True`` is set *before* ``import encodermap``
  (``encodermap/__init__.py:189-206``) -- ``install()`` checks it;
* outputs are allocated with torch and returned to TF through DLPack (zero copy).

Gradients: each ``backward`` below calls the matching libemk backward kernel on the saved inputs/outputs -- no torch
autograd graph is kept alive between TF's forward and backward passes.
"""
from __future__ import annotations

import os
import threading
from math import pi

import torch

from . import _ops

tf = None   # bound by _require_tf()


def _require_tf():
    global tf
    if tf is None:
        try:
            import tensorflow as _tf
        except Exception as e:  # noqa: BLE001
            raise ImportError("encodermap_b200.tf_adapter needs TensorFlow >= 2.13 (not installed in this environment)") from e
        tf = _tf
    return tf


def _reset_tf_binding() -> None:
    """Tests: forget the bound module (a different stand-in may be registered next)."""
    global tf
    tf = None


# ---- tensor hand-over ---------------------------------------------------------------------------------------------------
def _to_torch(t):
    """tf.Tensor (GPU) -> torch tensor sharing memory."""
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t: torch.Tensor):
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def _enter() -> None:
    # TF's producer stream is not visible through DLPack: its pending kernels must have finished before libemk reads.
    # (Without a CUDA device there is nothing to wait for -- and nothing to compute with: the libemk call raises.)
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _leave() -> None:
    # libemk launched on torch's current stream; TF's consumers run on TF's stream: wait for OUR stream only
    if torch.cuda.is_available():
        torch.cuda.current_stream().synchronize()


def _eager(fn, inputs, n_out):
    """Run ``fn(*torch_tensors) -> tuple of torch float32 tensors`` eagerly from inside a (possibly traced) TF function."""

    def body(*tf_inputs):
        _enter()
        outs = fn(*[_to_torch(t) for t in tf_inputs])
        _leave()
        return [_to_tf(o) for o in outs]

    return tf.py_function(body, inputs, [tf.float32] * n_out)


def _f32(x):
    return tf.convert_to_tensor(x, dtype=tf.float32)


# ---- encodermap/misc/distances.py ------------------------------------------------------------------------------------------
def sigmoid(sig, a, b):
    """``encodermap.misc.distances.sigmoid`` (:66-88): tensors go through libemk, python numbers / ndarrays are
    evaluated on the host exactly as the reference's pure-python closure does (tests/test_pairwise_distances.py:191-193)."""
    _require_tf()

    def func(r):
        if not tf.is_tensor(r):
            return 1 - (1 + (2 ** (a / b) - 1) * (r / sig) ** a) ** (-b / a)

        @tf.custom_gradient
        def op(x):
            (out,) = _eager(lambda t: (_ops.sigmoid_raw(t, sig, a, b),), [x], 1)
            out = tf.reshape(out, tf.shape(x))

            def backward(g):
                (gr,) = _eager(lambda t, g_: (_ops.sigmoid_bwd_raw(t, sig, a, b, g_),), [x, g], 1)
                return tf.reshape(gr, tf.shape(x))

            return out, backward

        return op(_f32(r))

    return func


def periodic_distance(a, b, periodicity=2 * pi):
    """``encodermap.misc.distances.periodic_distance`` (:113-141); operands broadcast like the reference's."""
    _require_tf()
    a, b = _f32(a), _f32(b)
    out_shape = tf.shape(a + b)            # broadcast shape (values unused)
    # the broadcast itself is a TF op: its own gradient sums over the broadcast axes
    a_b, b_b = tf.broadcast_to(a, out_shape), tf.broadcast_to(b, out_shape)

    @tf.custom_gradient
    def op(x, y):
        (out,) = _eager(lambda s, t: (_ops.periodic_distance_raw(s, t, periodicity),), [x, y], 1)
        out = tf.reshape(out, out_shape)

        def backward(g):
            ga, gb = _eager(lambda s, t, g_: _ops.periodic_distance_bwd_raw(s, t, periodicity, g_), [x, y, g], 2)
            return tf.reshape(ga, out_shape), tf.reshape(gb, out_shape)

        return out, backward

    return op(a_b, b_b)


def pairwise_dist_periodic(positions, periodicity):
    """``encodermap.misc.distances.pairwise_dist_periodic`` (:144-176): (n,d) -> (n,n); TensorFlow's autodiff tie rules
    in the backward kernel (abs' = sign, minimum routes to its first operand)."""
    _require_tf()
    positions = _f32(positions)
    assert len(positions.shape) == 2   # distances.py:161

    @tf.custom_gradient
    def op(x):
        n = tf.shape(x)[0]
        (out,) = _eager(lambda t: (_ops.pairwise_dist_periodic_raw(t, periodicity),), [x], 1)
        out = tf.reshape(out, [n, n])

        def backward(g):
            (gx,) = _eager(lambda t, o, g_: (_ops.pairwise_dist_periodic_bwd_raw(t, periodicity, o, g_),), [x, out, g], 1)
            return tf.reshape(gx, tf.shape(x))

        return out, backward

    return op(positions)


def _pairwise_op(x, squared, flat, start, stop, step, rank3):
    @tf.custom_gradient
    def op(x):
        (out,) = _eager(lambda t: (_ops.pairwise_dist_raw(t, squared, flat, start, stop, step),), [x], 1)
        # py_function outputs carry no shape: restore it from what libemk returned (rank-2 input gains a batch axis)
        b = tf.shape(x)[0] if rank3 else 1
        n_all = int(x.shape[1] if rank3 else x.shape[0])
        n = len(range(*slice(start, stop, step).indices(n_all)))
        out = tf.reshape(out, [b, n * (n - 1) // 2] if flat else [b, n, n])

        def backward(g):
            (gx,) = _eager(lambda t, g_: (_ops.pairwise_dist_bwd_raw(t, g_, squared, flat, start, stop, step),), [x, g], 1)
            return tf.reshape(gx, tf.shape(x))

        return out, backward

    return op(x)


def pairwise_dist(positions, squared=False, flat=False):
    """``encodermap.misc.distances.pairwise_dist`` (:179-255): accepts ndarray / list / tensor of rank 2 or 3."""
    _require_tf()
    positions = _f32(positions)
    return _pairwise_op(positions, bool(squared), bool(flat), None, None, None, len(positions.shape) == 3)


# ---- encodermap/loss_functions/loss_functions.py ------------------------------------------------------------------------------
def sigmoid_loss(parameters=None, periodicity_overwrite=None, dist_dig_parameters_overwrite=None):
    """``encodermap.loss_functions.loss_functions.sigmoid_loss`` (:301-369): one fused libemk launch gives the cost and
    dL/d(y_pred).  ``distance_loss`` (:200-298) and ``cartesian_distance_loss`` (:873-944) call this factory through the
    module global, so rebinding it is enough for both."""
    _require_tf()
    periodicity = periodicity_overwrite if periodicity_overwrite is not None else getattr(parameters, "periodicity", 2 * pi)
    sig = tuple(dist_dig_parameters_overwrite if dist_dig_parameters_overwrite is not None
                else getattr(parameters, "dist_sig_parameters", (4.5, 12, 6, 1, 2, 6)))

    @tf.custom_gradient
    def _cost(y_true, y_pred):
        def run(h, z):
            loss, grad = _ops.sigmoid_cost_raw(h, z, periodicity, sig)
            return loss.to(torch.float32), grad

        loss, grad = _eager(run, [y_true, y_pred], 2)
        loss = tf.reshape(loss, [])
        grad = tf.reshape(grad, tf.shape(y_pred))

        def backward(upstream):
            # the high-d side is input data in every caller: no gradient is produced for it
            return tf.zeros_like(y_true), upstream * grad

        return loss, backward

    def sigmoid_loss_func(y_true, y_pred):
        cost = _cost(_f32(y_true), _f32(y_pred))
        tf.debugging.assert_all_finite(cost, message="Sigmoid cost became infinite or NaN.")
        return cost

    return sigmoid_loss_func


def cartesian_distance_loss_from_coordinates(model, parameters=None, callback=None):
    """``cartesian_distance_loss`` (loss_functions.py:873-944) fed with the input COORDINATES instead of the stored pair matrix
    (the composition models/models.py:837-839 + :2419-2422): ``f(inp_cartesians, latent)``.  The (batch, n_pairs) matrix is
    library scratch.  Not installed by ``install()`` -- the reference's model builders pass ``inp_pair``; a caller opts in."""
    _require_tf()
    p = parameters
    sig = tuple(getattr(p, "cartesian_dist_sig_parameters", (4.5, 12, 6, 1, 2, 6)))
    sel = (getattr(p, "cartesian_pwd_start", None), getattr(p, "cartesian_pwd_stop", None), getattr(p, "cartesian_pwd_step", None))

    @tf.custom_gradient
    def _cost(cartesians, y_pred):
        def run(x, z):
            loss, grad = _ops.cartesian_distance_cost_raw(x, z, sig, *sel)
            return loss.to(torch.float32), grad

        loss, grad = _eager(run, [cartesians, y_pred], 2)
        loss = tf.reshape(loss, [])
        grad = tf.reshape(grad, tf.shape(y_pred))

        def backward(upstream):
            return tf.zeros_like(cartesians), upstream * grad      # the coordinates are input data

        return loss, backward

    def cartesian_distance_loss_func(cartesians, y_pred):
        scale = getattr(p, "cartesian_distance_cost_scale", 1)
        cost = _cost(_f32(cartesians), _f32(y_pred)) * scale if scale is not None else tf.constant(0.0)
        tf.debugging.assert_all_finite(cost, message="Cartesian distance cost became infinite or NaN.")
        return cost

    return cartesian_distance_loss_func


def fused_cartesian_loss(model=None, scale_callback=None, parameters=None, log_callback=None):
    """``PairwiseDistances("output")`` + ``cartesian_loss`` (models/layers.py:1252-1267 + loss_functions.py:947-1067, called as
    ``cartesian_loss_func(inp_pair, out_pair)`` in models/models.py:2385-2387) as ONE op on the back-mapped coordinates:
    ``f(inp_cartesians | inp_pair, out_cartesians)``; d(cost)/d(out_cartesians) comes out of the same launch.  Opt-in like
    ``cartesian_distance_loss_from_coordinates``."""
    _require_tf()
    p = parameters
    sel = (getattr(p, "cartesian_pwd_start", None), getattr(p, "cartesian_pwd_stop", None), getattr(p, "cartesian_pwd_step", None))
    variant = getattr(p, "cartesian_cost_variant", "mean_abs")

    @tf.custom_gradient
    def _cost(target, out_cartesians):
        n_atoms = int(out_cartesians.shape[1])
        n_sel = len(range(*slice(*sel).indices(n_atoms)))
        frames = tf.cast(tf.shape(out_cartesians)[0], tf.float32)
        count = frames * (1.0 if variant == "mean_norm" else float(max(1, n_sel * (n_sel - 1) // 2)))

        def run(t, x):
            loss, grad, _ = _ops.cartesian_pair_loss_raw(x, t, *sel, variant, 0.0, True)
            return loss.to(torch.float32), grad

        loss, grad = _eager(run, [target, out_cartesians], 2)
        loss = tf.reshape(loss, []) / count
        grad = tf.reshape(grad, tf.shape(out_cartesians)) / count

        def backward(upstream):
            return tf.zeros_like(target), upstream * grad

        return loss, backward

    def cartesian_loss_func(y_true, out_cartesians):
        scale = scale_callback.current_cartesian_cost_scale if scale_callback is not None else getattr(p, "cartesian_cost_scale", 1)
        cost = _cost(_f32(y_true), _f32(out_cartesians)) / getattr(p, "cartesian_cost_reference", 1) * scale
        tf.debugging.assert_all_finite(cost, message="Cartesian cost became infinite or NaN.")
        return cost

    return cartesian_loss_func


# ---- encodermap/models/layers.py ---------------------------------------------------------------------------------------------
def periodic_input(inputs, periodicity):
    """``PeriodicInput.call`` (models/layers.py:204-215): (rows, d) -> (rows, 2d) = [sin x, cos x]."""
    _require_tf()

    @tf.custom_gradient
    def op(x):
        (out,) = _eager(lambda t: (_ops.periodic_input_raw(t, periodicity),), [x], 1)
        out = tf.reshape(out, [tf.shape(x)[0], 2 * int(x.shape[1])])

        def backward(g):
            (gx,) = _eager(lambda t, g_: (_ops.periodic_input_bwd_raw(t, periodicity, g_),), [x, g], 1)
            return tf.reshape(gx, tf.shape(x))

        return out, backward

    return op(_f32(inputs))


def pairwise_distances(inputs, start=None, stop=None, step=None):
    """``PairwiseDistances.call`` without the side-chain gather (models/layers.py:1252-1267): the atom selection
    ``inputs[:, start:stop:step]`` is a stride inside the kernel, not a copy."""
    _require_tf()
    return _pairwise_op(_f32(inputs), False, True, start, stop, step, True)


def back_map(distances, angles, dihedrals):
    """``BackMapLayer.call`` (models/layers.py:957-986) as one differentiable op: batch-mean bond lengths,
    ``chain_in_plane``, ``dihedrals + pi``, ``dihedrals_to_cartesian_tf_layers`` -- one forward and one backward kernel."""
    _require_tf()

    @tf.custom_gradient
    def op(distances, angles, dihedrals):
        def run(d_, a_, p_):
            return (_ops.backmap_raw(_ops.column_mean_raw(d_)[None], a_, p_),)

        (xyz,) = _eager(run, [distances, angles, dihedrals], 1)
        xyz = tf.reshape(xyz, [tf.shape(angles)[0], int(angles.shape[1]) + 2, 3])

        def backward(g):
            def run_b(d_, a_, x_, g_):
                gl, ga, gd = _ops.backmap_bwd_raw(_ops.column_mean_raw(d_)[None], a_, x_, g_, True, True, True)
                return (gl / d_.shape[0]).expand(d_.shape[0], -1).contiguous(), ga, gd   # d(mean)/d(row) = 1/rows

            gdist, ga, gd = _eager(run_b, [distances, angles, xyz, g], 3)
            return tf.reshape(gdist, tf.shape(distances)), tf.reshape(ga, tf.shape(angles)), tf.reshape(gd, tf.shape(dihedrals))

        return xyz, backward

    return op(_f32(distances), _f32(angles), _f32(dihedrals))


_SIDECHAIN_PLANS = {}
_SIDECHAIN_LOCK = threading.Lock()      # py_function bodies may run on TensorFlow's executor threads


def _sidechain_plan(counts, device):
    key = (tuple(int(c) for c in counts), str(device))
    with _SIDECHAIN_LOCK:
        plan = _SIDECHAIN_PLANS.get(key)
        if plan is None:
            plan = _ops.SidechainPlan(key[0], device)
            _SIDECHAIN_PLANS[key] = plan
    return plan


def back_map_with_sidechains(feature_description, inputs):
    """``BackMapLayerWithSidechains.call`` (models/layers.py:533-843) as one differentiable op: the six inputs (central
    distances / angles / dihedrals, side distances / angles / dihedrals) -> (batch, n_atoms, 3); one forward and one backward
    kernel instead of a dozen TensorFlow kernels per bond angle and dihedral."""
    _require_tf()
    info = feature_description[-1]
    counts = [int(info[k]) for k in sorted(info.keys())]
    n_atoms = 3 * len(counts) + sum(c + 1 for c in counts if c > 0)

    @tf.custom_gradient
    def op(*ins):
        def run(*ts):
            return (_ops.sidechain_backmap_raw(_sidechain_plan(counts, ts[0].device), ts),)

        (xyz,) = _eager(run, list(ins), 1)
        xyz = tf.reshape(xyz, [tf.shape(ins[0])[0], n_atoms, 3])

        def backward(g):
            def run_b(*ts):
                return tuple(_ops.sidechain_backmap_bwd_raw(_sidechain_plan(counts, ts[0].device), ts[:6], ts[6]))

            grads = _eager(run_b, list(ins) + [g], 6)
            return tuple(tf.reshape(gr, tf.shape(t)) for gr, t in zip(grads, ins))

        return xyz, backward

    return op(*[_f32(t) for t in inputs])


def gathered_pairwise_distances(inputs, indices):
    """``PairwiseDistances.call`` with reconstructed side chains (models/layers.py:1260-1265): tf.gather of the selected atoms +
    flat pairwise distances, both in libemk."""
    _require_tf()
    idx_host = [int(i) for i in indices]
    n_sel = len(idx_host)
    inputs = _f32(inputs)
    n_atoms = int(inputs.shape[1])
    if n_sel and (min(idx_host) < 0 or max(idx_host) >= n_atoms):
        raise IndexError(f"PairwiseDistances: atom index {max(idx_host)} outside the {n_atoms} atoms of the input")

    @tf.custom_gradient
    def op(x):
        def run(t):
            index = torch.as_tensor(idx_host, dtype=torch.int32, device=t.device)
            return (_ops.pairwise_dist_raw(_ops.gather_atoms_raw(t, index), False, True, None, None, None),)

        (out,) = _eager(run, [x], 1)
        out = tf.reshape(out, [tf.shape(x)[0], n_sel * (n_sel - 1) // 2])

        def backward(g):
            def run_b(t, g_):
                index = torch.as_tensor(idx_host, dtype=torch.int32, device=t.device)
                sel = _ops.gather_atoms_raw(t, index)
                g_sel = _ops.pairwise_dist_bwd_raw(sel, g_, False, True, None, None, None)
                return (_ops.gather_atoms_bwd_raw(g_sel, index, t.shape[1]),)

            (gx,) = _eager(run_b, [x, g], 1)
            return tf.reshape(gx, tf.shape(x))

        return out, backward

    return op(inputs)


# ---- encodermap/encodermap_tf1/backmapping.py, encodermap/misc/backmapping.py ---------------------------------------------
def chain_in_plane(lengths, angles):
    """``encodermap.encodermap_tf1.backmapping.chain_in_plane`` (:97-119); ``lengths`` (1,n-1) or (b,n-1)."""
    _require_tf()

    @tf.custom_gradient
    def op(lengths, angles):
        (xyz,) = _eager(lambda l_, a_: (_ops.chain_in_plane_raw(l_, a_),), [lengths, angles], 1)
        xyz = tf.reshape(xyz, [tf.shape(angles)[0], int(angles.shape[1]) + 2, 3])

        def backward(g):
            gl, ga = _eager(lambda l_, a_, g_: _ops.chain_in_plane_bwd_raw(l_, a_, g_, True, True), [lengths, angles, g], 2)
            return tf.reshape(gl, tf.shape(lengths)), tf.reshape(ga, tf.shape(angles))

        return xyz, backward

    return op(_f32(lengths), _f32(angles))


def _d2c(dihedrals, cartesian, one_way):
    @tf.custom_gradient
    def op(dihedrals, cartesian):
        (xyz,) = _eager(lambda d_, c_: (_ops.d2c_raw(d_, c_, one_way),), [dihedrals, cartesian], 1)
        xyz = tf.reshape(xyz, [tf.shape(dihedrals)[0], int(dihedrals.shape[1]) + 3, 3])

        def backward(g):
            def run_b(c_, x_, g_):
                return _ops.d2c_bwd_raw(x_, g_, one_way), _ops.d2c_chain_bwd_raw(c_, x_, g_, one_way)

            gd, gc = _eager(run_b, [cartesian, xyz, g], 2)
            return tf.reshape(gd, tf.shape(dihedrals)), tf.reshape(gc, tf.shape(cartesian))

        return xyz, backward

    return op(_f32(dihedrals), _f32(cartesian))


def dihedrals_to_cartesian_tf(dihedrals, cartesian):
    """``encodermap.encodermap_tf1.backmapping.dihedrals_to_cartesian_tf`` (:164-195); a rank-2 ``cartesian`` is shared by
    all frames (the reference tiles it)."""
    _require_tf()
    return _d2c(dihedrals, cartesian, 0)


def dihedrals_to_cartesian_tf_layers(dihedrals, cartesians, left_iteration_counter, right_iteration_counter):
    """``encodermap.misc.backmapping.dihedrals_to_cartesian_tf_layers`` (:259-309); the counters are implied by the shapes
    and checked against the reference's formula (models/models.py:661-671)."""
    _require_tf()
    n = int(dihedrals.shape[-1]) + 3
    if (left_iteration_counter, right_iteration_counter) != (n // 2 - 1, (n - 3) // 2):
        raise ValueError(f"iteration counters ({left_iteration_counter},{right_iteration_counter}) do not match {n} atoms: "
                         f"expected ({n // 2 - 1},{(n - 3) // 2})")
    return _d2c(dihedrals, cartesians, 0)


def dihedral_to_cartesian_tf_one_way(dihedrals, cartesian):
    """``encodermap.encodermap_tf1.backmapping.dihedral_to_cartesian_tf_one_way`` (:198-214)."""
    _require_tf()
    return _d2c(dihedrals, cartesian, 1)


def dihedral_to_cartesian_tf_one_way_layers(dihedrals, cartesian, n):
    """``encodermap.misc.backmapping.dihedral_to_cartesian_tf_one_way_layers`` (:1873-1912)."""
    _require_tf()
    if n != int(dihedrals.shape[-1]):
        raise ValueError("n must equal dihedrals.shape[-1]")
    return _d2c(dihedrals, cartesian, 1)


def rotation_matrix(axis_unit_vec, angle):
    """``encodermap.misc.backmapping.rotation_matrix`` (:1950-1968); forward only (the scan kernels never build it)."""
    _require_tf()
    (out,) = _eager(lambda a_, g_: (_ops.rotation_matrix_raw(a_, g_),), [_f32(axis_unit_vec), _f32(angle)], 1)
    return tf.reshape(out, [tf.shape(angle)[0], 3, 3])


# ---- generation side: guessed amide H / O, merge (encodermap/misc/backmapping.py:1920-1990; forward only) ---------------------------
def _indices(idx):
    import numpy as np

    return np.asarray(idx.numpy() if hasattr(idx, "numpy") else idx, dtype=np.int64).reshape(-1).tolist()


def guess_sp2_atom(cartesians, indices, angle_to_previous, bond_length):
    """``encodermap.misc.backmapping.guess_sp2_atom`` (:1920-1941)."""
    _require_tf()
    idx = _indices(indices)
    (out,) = _eager(lambda x: (_ops.guess_sp2_raw(x, idx, angle_to_previous, bond_length),), [_f32(cartesians)], 1)
    return tf.reshape(out, [tf.shape(cartesians)[0], len(idx), 3])


def guess_amide_H(cartesians, N_indices):
    """``encodermap.misc.backmapping.guess_amide_H`` (:1943-1944)."""
    return guess_sp2_atom(cartesians, _indices(N_indices)[1::], 123 / 180 * pi, 1.10)


def guess_amide_O(cartesians, C_indices):
    """``encodermap.misc.backmapping.guess_amide_O`` (:1946-1947)."""
    return guess_sp2_atom(cartesians, _indices(C_indices), 121 / 180 * pi, 1.24)


def merge_cartesians(central_cartesians, N_indices, O_indices, H_cartesians, O_cartesians):
    """``encodermap.misc.backmapping.merge_cartesians`` (:1970-1990)."""
    _require_tf()
    h_after, o_after = _indices(N_indices)[1::], _indices(O_indices)
    n_out = int(central_cartesians.shape[1]) + int(H_cartesians.shape[1]) + int(O_cartesians.shape[1])
    (out,) = _eager(lambda c, h, o: (_ops.merge_cartesians_raw(c, h_after, o_after, h, o),),
                    [_f32(central_cartesians), _f32(H_cartesians), _f32(O_cartesians)], 1)
    return tf.reshape(out, [tf.shape(central_cartesians)[0], n_out, 3])


# ---- installation ----------------------------------------------------------------------------------------------------------------
# (module, name) -> replacement; these are the `from ... import` sites of the reference (loss_functions.py:43-48,
# models/layers.py:47-54, models/models.py:49-58, autoencoder/autoencoder.py:64-80): a name imported into a module is a
# separate binding, so every importing module is patched, not only the defining one.
def _rebind_table():
    return {
        "encodermap.misc.distances": dict(sigmoid=sigmoid, periodic_distance=periodic_distance,
                                          pairwise_dist_periodic=pairwise_dist_periodic, pairwise_dist=pairwise_dist),
        "encodermap.loss_functions.loss_functions": dict(sigmoid_loss=sigmoid_loss, sigmoid=sigmoid, periodic_distance=periodic_distance,
                                                         pairwise_dist_periodic=pairwise_dist_periodic, pairwise_dist=pairwise_dist),
        "encodermap.encodermap_tf1.backmapping": dict(chain_in_plane=chain_in_plane, dihedrals_to_cartesian_tf=dihedrals_to_cartesian_tf,
                                                      dihedral_to_cartesian_tf_one_way=dihedral_to_cartesian_tf_one_way),
        "encodermap.misc.backmapping": dict(dihedrals_to_cartesian_tf_layers=dihedrals_to_cartesian_tf_layers,
                                            dihedral_to_cartesian_tf_one_way_layers=dihedral_to_cartesian_tf_one_way_layers,
                                            rotation_matrix=rotation_matrix, guess_sp2_atom=guess_sp2_atom, guess_amide_H=guess_amide_H,
                                            guess_amide_O=guess_amide_O, merge_cartesians=merge_cartesians),
        "encodermap.models.layers": dict(pairwise_dist=pairwise_dist, chain_in_plane=chain_in_plane,
                                         dihedrals_to_cartesian_tf_layers=dihedrals_to_cartesian_tf_layers),
        "encodermap.models.models": dict(pairwise_dist=pairwise_dist, chain_in_plane=chain_in_plane,
                                         dihedrals_to_cartesian_tf=dihedrals_to_cartesian_tf),
        "encodermap.autoencoder.autoencoder": dict(pairwise_dist=pairwise_dist, chain_in_plane=chain_in_plane,
                                                   dihedrals_to_cartesian_tf=dihedrals_to_cartesian_tf),
    }


def _layer_calls():
    def periodic_input_call(self, inputs):
        return periodic_input(inputs, self.p.periodicity)

    def pairwise_distances_call(self, inputs):
        if getattr(self.p, "reconstruct_sidechains", False):
            return gathered_pairwise_distances(inputs, self.indices)      # the gather of layers.py:1260-1265 and the distances
        return pairwise_distances(inputs, self.p.cartesian_pwd_start, self.p.cartesian_pwd_stop, self.p.cartesian_pwd_step)

    def back_map_call(self, inputs):
        distances, angles, dihedrals = inputs
        n = int(angles.shape[1]) + 2
        if (self.left_split, self.right_split) != (n // 2 - 1, (n - 3) // 2):
            raise ValueError(f"BackMapLayer(left_split={self.left_split}, right_split={self.right_split}) does not match {n} atoms")
        return back_map(distances, angles, dihedrals)

    def back_map_with_sidechains_call(self, inputs):
        return back_map_with_sidechains(self.feature_description, inputs)

    return {"PeriodicInput": periodic_input_call, "PairwiseDistances": pairwise_distances_call, "BackMapLayer": back_map_call,
            "BackMapLayerWithSidechains": back_map_with_sidechains_call}


def install(enable_layers: bool = True, require_gpu_env: bool = True) -> dict:
    """Rebind the hot-path names inside an importable ``encodermap`` package.  Call it BEFORE constructing
    ``EncoderMap`` / ``AngleDihedralCartesianEncoderMap``: the loss closures capture ``sigmoid_loss(p)`` at construction
    (loss_functions.py:263, 917-921).  Returns {"module.name": original object} so that ``uninstall`` can restore it.
    Modules of the table that the installed encodermap does not have are skipped (TF1-only or trimmed installs)."""
    _require_tf()
    if require_gpu_env and os.environ.get("ENCODERMAP_ENABLE_GPU", "False") != "True":
        raise RuntimeError("set ENCODERMAP_ENABLE_GPU=True before importing encodermap: it hides all GPUs otherwise")
    import importlib

    saved = {}
    for modname, names in _rebind_table().items():
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        for name, repl in names.items():
            if hasattr(mod, name):
                saved[f"{modname}.{name}"] = getattr(mod, name)
                setattr(mod, name, repl)
    if enable_layers:
        try:
            layers = importlib.import_module("encodermap.models.layers")
        except ImportError:
            layers = None
        if layers is not None:
            for cls_name, call in _layer_calls().items():
                cls = getattr(layers, cls_name, None)
                if cls is not None:
                    saved[f"encodermap.models.layers.{cls_name}.call"] = cls.call
                    cls.call = call
    return saved


def uninstall(saved: dict) -> None:
    import importlib

    for key, obj in saved.items():
        if key.endswith(".call"):
            modname, cls_name, _ = key.rsplit(".", 2)
            setattr(getattr(importlib.import_module(modname), cls_name), "call", obj)
        else:
            modname, name = key.rsplit(".", 1)
            setattr(importlib.import_module(modname), name, obj)
