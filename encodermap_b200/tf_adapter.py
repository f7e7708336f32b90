"""TensorFlow side of the drop-in: the same libemk entry points wrapped as ``tf.custom_gradient``
functions with the reference's signatures, and ``install()`` which rebinds them into the reference's
modules so that ``em.EncoderMap`` / ``AngleDihedralCartesianEncoderMap`` train unchanged.

TensorFlow is not installed in the build image nor on the GPU box, so this module is
import-guarded and **untested here**; the torch adapter (``_ops.py``) exercises the identical C entry
points.  Hazards encoded below:

* Keras traces ``train_step`` into a graph, so every ctypes call sits inside ``tf.py_function``;
* TF's DLPack export does not synchronise its compute stream: we synchronise the device before launching
  on our own stream and again before handing results back;
* the reference hides all GPUs unless ``ENCODERMAP_ENABLE_GPU=True`` is set *before* ``import encodermap``
  (``encodermap/__init__.py:189-206``) -- ``install()`` checks it;
* outputs are allocated with torch and returned to TF through DLPack (zero copy).
"""
from __future__ import annotations

import os
from math import pi

try:  # pragma: no cover - TensorFlow is absent in this image
    import tensorflow as tf
except Exception:  # noqa: BLE001
    tf = None

import torch

from . import _ops


def _require_tf():
    if tf is None:
        raise ImportError("encodermap_b200.tf_adapter needs TensorFlow >= 2.13 (not installed in this environment)")


def _to_torch(t):
    """tf.Tensor (GPU) -> torch tensor sharing memory."""
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t: torch.Tensor):
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def _eager(fn, inputs, n_out):
    """Run ``fn(*torch_tensors) -> tuple of torch tensors`` eagerly from inside a traced graph."""

    def body(*tf_inputs):
        torch.cuda.synchronize()  # TF's producer stream is not visible through DLPack
        outs = fn(*[_to_torch(t) for t in tf_inputs])
        torch.cuda.synchronize()
        return [_to_tf(o) for o in outs]

    return tf.py_function(body, inputs, [tf.float32] * n_out)


def sigmoid_loss(parameters=None, periodicity_overwrite=None, dist_dig_parameters_overwrite=None):
    """``encodermap.loss_functions.loss_functions.sigmoid_loss`` (:301-369) on libemk."""
    _require_tf()
    periodicity = periodicity_overwrite if periodicity_overwrite is not None else getattr(parameters, "periodicity", 2 * pi)
    sig = tuple(dist_dig_parameters_overwrite if dist_dig_parameters_overwrite is not None
                else getattr(parameters, "dist_sig_parameters", (4.5, 12, 6, 1, 2, 6)))

    @tf.custom_gradient
    def sigmoid_loss_func(y_true, y_pred):
        def run(h, z):
            loss, grad = _ops.sigmoid_cost_raw(h, z, periodicity, sig)
            return loss.to(torch.float32), grad

        loss, grad = _eager(run, [y_true, y_pred], 2)
        loss = tf.reshape(loss, [])
        grad = tf.reshape(grad, tf.shape(y_pred))

        def backward(upstream):
            return tf.zeros_like(y_true), upstream * grad  # the high-d side is input data

        tf.debugging.assert_all_finite(loss, message="Sigmoid cost became infinite or NaN.")
        return loss, backward

    return sigmoid_loss_func


def back_map(distances, angles, dihedrals):
    """``BackMapLayer.call`` (encodermap/models/layers.py:957-986) as one differentiable TF op."""
    _require_tf()

    @tf.custom_gradient
    def op(distances, angles, dihedrals):
        lengths = tf.expand_dims(tf.reduce_mean(distances, 0), 0)

        def run(l_, a_, d_):
            return (_ops.BackMap.apply(l_, a_, d_),)

        (xyz,) = _eager(run, [lengths, angles, dihedrals], 1)
        xyz = tf.reshape(xyz, tf.concat([tf.shape(angles)[:1], [tf.shape(angles)[1] + 2, 3]], 0))

        def backward(g):
            def run_b(l_, a_, x_, g_):
                b, n = x_.shape[0], x_.shape[1]
                ga = torch.empty_like(a_)
                gd = torch.empty((b, n - 3), dtype=torch.float32, device=x_.device)
                gl = torch.empty((b, n - 1), dtype=torch.float32, device=x_.device)
                from . import _lib

                with torch.cuda.device(x_.device):
                    _lib.check(_lib.lib().emk_dl_backmap_bwd(_lib.DL(l_), _lib.DL(a_), _lib.DL(x_), _lib.DL(g_.contiguous()),
                                                              _lib.DL(ga), _lib.DL(gd), _lib.DL(gl), _lib.stream_of(x_)))
                return gl.sum(0, keepdim=True), ga, gd

            gl, ga, gd = _eager(run_b, [lengths, angles, xyz, g], 3)
            rows = tf.cast(tf.shape(distances)[0], tf.float32)
            return tf.broadcast_to(tf.reshape(gl, [1, -1]) / rows, tf.shape(distances)), tf.reshape(ga, tf.shape(angles)), tf.reshape(gd, tf.shape(dihedrals))

        return xyz, backward

    return op(distances, angles, dihedrals)


def pairwise_dist(positions, squared=False, flat=False):
    """``encodermap.misc.distances.pairwise_dist`` (:179-255)."""
    _require_tf()
    positions = tf.convert_to_tensor(positions, dtype=tf.float32)

    @tf.custom_gradient
    def op(x):
        (out,) = _eager(lambda t: (_ops.PairwiseDist.apply(t, bool(squared), bool(flat), None, None, None),), [x], 1)

        def backward(g):
            def run_b(t, g_):
                t = t.detach().requires_grad_(True)
                with torch.enable_grad():
                    o = _ops.PairwiseDist.apply(t, bool(squared), bool(flat), None, None, None)
                (gx,) = torch.autograd.grad(o, t, g_.reshape(o.shape))
                return (gx,)

            (gx,) = _eager(run_b, [x, g], 1)
            return tf.reshape(gx, tf.shape(x))

        return out, backward

    return op(positions)


def install(enable_layers: bool = True):
    """Rebind the hot-path names inside an importable ``encodermap`` package (call before constructing
    ``EncoderMap`` / ``AngleDihedralCartesianEncoderMap``: the loss closures capture at construction,
    loss_functions.py:263, 917-921)."""
    _require_tf()
    if os.environ.get("ENCODERMAP_ENABLE_GPU", "False") != "True":
        raise RuntimeError("set ENCODERMAP_ENABLE_GPU=True before importing encodermap: it hides all GPUs otherwise")
    import encodermap.loss_functions.loss_functions as lf
    import encodermap.misc.distances as dists
    import encodermap.models.layers as layers
    import encodermap.models.models as models

    lf.sigmoid_loss = sigmoid_loss  # distance_loss / cartesian_distance_loss call it through the module global
    dists.pairwise_dist = pairwise_dist
    lf.pairwise_dist = pairwise_dist
    layers.pairwise_dist = pairwise_dist
    if enable_layers:
        def _call(self, inputs):
            distances, angles, dihedrals = inputs
            return back_map(distances, angles, dihedrals)

        layers.BackMapLayer.call = _call
        models.BackMapLayer = layers.BackMapLayer
