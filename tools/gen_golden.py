#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE's own function bodies.

TensorFlow is not installed in this image, so ``import encodermap`` fails.  The hot-path
functions, however, are short pure functions of a dozen ``tf.*`` primitives.  This script

  1. parses the reference source files under ``/root/reference`` with ``ast`` and pulls out
     the hot-path function definitions by name (nothing is copied into the repo -- the code
     is compiled and executed in memory, straight from where it lies);
  2. executes them against ``_TFShim`` -- a numpy-backed stand-in for the handful of ``tf``
     symbols they touch (float64 or float32);
  3. writes inputs and outputs to ``tests/golden/*.npz``.

The oracle (``oracle/em_oracle.py``) and the CUDA kernels are then tested against these
files.  ``/root/reference`` only exists in the build container, so this script runs there
only; the fixtures it writes are committed.

    python tools/gen_golden.py            # rewrites tests/golden/*.npz
"""

from __future__ import annotations

import ast
import math
import os
import sys
from pathlib import Path

import numpy as np

REF = Path(os.environ.get("EMK_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


# ----------------------------------------------------------------------------------------
# numpy-backed stand-in for the tf namespace used by the hot-path functions
# ----------------------------------------------------------------------------------------


class T(np.ndarray):
    """ndarray that also answers the few tf.Tensor methods the reference calls."""

    def get_shape(self):
        return self.shape

    def numpy(self):
        return np.asarray(self)

    # tf.Tensor is immutable: `r += x` rebinds r to a new (broadcast) tensor
    def __iadd__(self, other):
        return self + other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


def _w(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    return a.view(T)


class _Debugging:
    @staticmethod
    def is_numeric_tensor(x):
        return isinstance(x, T)

    @staticmethod
    def assert_all_finite(x, message=""):
        assert np.all(np.isfinite(np.asarray(x))), message
        return x

    @staticmethod
    def assert_rank(x, rank):
        assert np.ndim(x) == rank


class _Linalg:
    @staticmethod
    def diag_part(x):
        return _w(np.diagonal(x, axis1=-2, axis2=-1).copy())

    @staticmethod
    def cross(a, b):
        return _w(np.cross(a, b))


class _Math:
    @staticmethod
    def equal(a, b):
        return _w(np.equal(a, b))

    @staticmethod
    def mod(a, b):
        return _w(np.mod(a, b))


class _Errors:
    # `cartesians[:, i + 1]` past the last atom raises InvalidArgumentError in TensorFlow; numpy raises IndexError
    InvalidArgumentError = IndexError


class _TFShim:
    """Only what distances.py / loss_functions.py / backmapping.py / layers.py use on the path."""

    debugging = _Debugging()
    errors = _Errors()
    linalg = _Linalg()
    math = _Math()
    float32 = np.float32

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)

    # construction ---------------------------------------------------------------
    def convert_to_tensor(self, x, dtype=None):
        if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], (list, tuple)):
            x = np.array([[np.asarray(e) for e in row] for row in x])
        a = np.asarray(x)
        if a.dtype.kind == "f" or a.dtype.kind in "iu":
            a = a.astype(self.dtype) if a.dtype.kind == "f" else a
        return _w(a)

    def constant(self, x, dtype=None):
        return self.convert_to_tensor(x)

    def is_numeric_tensor(self, x):
        return isinstance(x, T)

    def zeros(self, shape, dtype=None):
        return _w(np.zeros(tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, int) else (shape,), self.dtype))

    def ones_like(self, x):
        return _w(np.ones_like(x))

    def eye(self, n):
        return _w(np.eye(n, dtype=self.dtype))

    def shape(self, x):
        return np.asarray(np.shape(x))

    # elementwise ----------------------------------------------------------------
    def abs(self, x):
        return _w(np.abs(x))

    def minimum(self, a, b):
        return _w(np.minimum(a, b))

    def maximum(self, a, b):
        return _w(np.maximum(a, b))

    def square(self, x):
        return _w(np.square(x))

    def sqrt(self, x):
        return _w(np.sqrt(x))

    def cos(self, x):
        return _w(np.cos(x))

    def sin(self, x):
        return _w(np.sin(x))

    def add(self, a, b):
        return _w(np.add(a, b))

    def equal(self, a, b):
        return _w(np.equal(a, b))

    def where(self, c, a, b):
        return _w(np.where(c, a, b))

    def cast(self, x, dtype):
        # the reference casts boolean masks to "float32"/np.float32; keep the working dtype so
        # the float64 evaluation stays float64 (the mask holds 0/1 exactly either way)
        return _w(np.asarray(x).astype(self.dtype))

    def to_float(self, x):
        return self.cast(x, None)

    # reductions / shape ops -------------------------------------------------------
    def reduce_sum(self, x, axis=None):
        return _w(np.sum(x, axis=axis))

    def reduce_mean(self, x, axis=None):
        return _w(np.mean(x, axis=axis))

    def norm(self, x, axis=None, keepdims=False):
        return _w(np.sqrt(np.sum(np.square(x), axis=axis, keepdims=keepdims)))

    def expand_dims(self, x, axis):
        return _w(np.expand_dims(x, axis))

    def transpose(self, x, perm=None):
        return _w(np.transpose(x, perm))

    def matmul(self, a, b):
        return _w(np.matmul(a, b))

    def stack(self, xs, axis=0):
        return _w(np.stack([np.asarray(x) for x in xs], axis=axis))

    def concat(self, xs, axis):
        return _w(np.concatenate([np.asarray(x) for x in xs], axis=axis))

    def tile(self, x, multiples):
        return _w(np.tile(x, [int(m) for m in multiples]))

    def boolean_mask(self, x, mask, axis=0):
        # tf.boolean_mask(x, mask, axis=1) with a rank-2 mask flattens dims (1,2) row-major
        assert axis == 1 and np.ndim(mask) == 2
        return _w(np.asarray(x)[:, np.asarray(mask)])

    def cond(self, pred, true_fn, false_fn, name=None):
        return true_fn() if bool(np.asarray(pred)) else false_fn()


# ----------------------------------------------------------------------------------------
# pull function definitions out of the reference sources
# ----------------------------------------------------------------------------------------


def _extract(path: Path, names, ns, take_last=()):
    """Compile the named top-level (or class-level) function definitions of ``path`` into ``ns`` (the first definition of a
    name, or the last for names in ``take_last``: functions preceded by ``@overload`` stubs)."""
    tree = ast.parse(path.read_text())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and (node.name not in found or node.name in take_last):
            node.decorator_list = []  # drop @tf.function / @overload
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            found[node.name] = node
    missing = set(names) - set(found)
    assert not missing, f"{path}: {missing} not found"
    mod = ast.Module(body=[found[n] for n in names], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, str(path), "exec"), ns)


def load_reference(dtype):
    tf = _TFShim(dtype)
    ns = {"tf": tf, "np": np, "pi": math.pi, "cos": math.cos, "sin": math.sin, "math": math,
          "overload": lambda f: f, "Union": None, "Number": None, "Callable": None, "Optional": None}
    _extract(REF / "encodermap/misc/distances.py",
             ["sigmoid", "periodic_distance_np", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"], ns)
    _extract(REF / "encodermap/misc/backmapping.py",
             ["split_and_reverse_dihedrals", "split_and_reverse_cartesians", "dihedrals_to_cartesian_tf_layers",
              "dihedral_to_cartesian_tf_one_way_layers", "rotation_matrix", "guess_sp2_atom", "guess_amide_H", "guess_amide_O",
              "merge_cartesians"], ns)
    ns1 = dict(ns)  # TF1 twins live in their own namespace (same names, different bodies)
    _extract(REF / "encodermap/encodermap_tf1/backmapping.py",
             ["straight_tetrahedral_chain", "chain_in_plane", "dihedrals_to_cartesian_tf",
              "dihedral_to_cartesian_tf_one_way"], ns1)
    _extract(REF / "encodermap/encodermap_tf1/misc.py", ["rotation_matrix", "distance_cost"], ns1)
    ns1["sigmoid_tf1"] = None
    # sigmoid_loss: the closure needs a Parameters-like object
    ns["Parameters"] = lambda: type("P", (), {"periodicity": 2 * math.pi, "dist_sig_parameters": (4.5, 12, 6, 1, 2, 6)})()
    _extract(REF / "encodermap/loss_functions/loss_functions.py", ["sigmoid_loss"], ns)
    ns["chain_in_plane"] = ns1["chain_in_plane"]
    # BackMapLayer.call / PeriodicInput.call / PairwiseDistances.call bodies
    ns["Concatenate"] = lambda axis, name=None: (lambda xs: tf.concat(xs, axis))
    tree = ast.parse((REF / "encodermap/models/layers.py").read_text())
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("BackMapLayer", "PeriodicInput", "PairwiseDistances")]:
        for fn in cls.body:
            if isinstance(fn, ast.FunctionDef) and fn.name == "call":
                fn.name = f"{cls.name}_call"
                fn.decorator_list = []
                fn.returns = None
                for a in fn.args.args:
                    a.annotation = None
                mod = ast.Module(body=[fn], type_ignores=[])
                ast.fix_missing_locations(mod)
                exec(compile(mod, "layers.py", "exec"), ns)
    return tf, ns, ns1


class _Self:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(20261017)
    tf64, R, R1 = load_reference(np.float64)
    tf32, R32, R132 = load_reference(np.float32)
    sig_default = (4.5, 12, 6, 1, 2, 6)

    # ---- distances ------------------------------------------------------------------------
    g = {}
    r = np.abs(rng.normal(size=64)) * 3
    for name, p in {"h_default": (4.5, 12, 6), "l_default": (1, 2, 6), "cube_h": (0.2, 3, 6), "nb_h": (0.3, 6, 6),
                    "nb_l": (1, 4, 6), "odd": (1.3, 2.5, 3.7)}.items():
        g[f"sigmoid_{name}_params"] = np.array(p, dtype=np.float64)
        g[f"sigmoid_{name}_out"] = np.asarray(R["sigmoid"](*p)(tf64.convert_to_tensor(r)))
    g["sigmoid_r"] = r
    a = rng.uniform(-math.pi, math.pi, size=(7, 5))
    b = rng.uniform(-math.pi, math.pi, size=(7, 5))
    g["perdist_a"], g["perdist_b"] = a, b
    g["perdist_2pi"] = np.asarray(R["periodic_distance"](tf64.convert_to_tensor(a), tf64.convert_to_tensor(b), 2 * math.pi))
    g["perdist_360"] = np.asarray(R["periodic_distance"](tf64.convert_to_tensor(a * 50), tf64.convert_to_tensor(b * 50), 360.0))
    g["perdist_inf"] = np.asarray(R["periodic_distance"](tf64.convert_to_tensor(a), tf64.convert_to_tensor(b), float("inf")))
    x = rng.uniform(-math.pi, math.pi, size=(48, 37))
    g["pwp_x"] = x
    g["pwp_2pi"] = np.asarray(R["pairwise_dist_periodic"](tf64.convert_to_tensor(x), 2 * math.pi))
    g["pwp_1"] = np.asarray(R["pairwise_dist_periodic"](tf64.convert_to_tensor(x), 1.0))
    y = rng.normal(size=(40, 3))
    g["pw_x2"] = y
    g["pw_2d"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y)))
    g["pw_2d_sq"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y), squared=True))
    g["pw_2d_flat"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y), flat=True))
    y3 = rng.normal(size=(5, 17, 3))
    g["pw_x3"] = y3
    g["pw_3d"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y3)))
    g["pw_3d_flat"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y3), flat=True))
    g["pw_3d_flat_sq"] = np.asarray(R["pairwise_dist"](tf64.convert_to_tensor(y3), flat=True, squared=True))
    # float32 run of the same functions (error band of the reference's own precision)
    g["pw_2d_f32"] = np.asarray(R32["pairwise_dist"](tf32.convert_to_tensor(y.astype(np.float32))))
    g["pwp_2pi_f32"] = np.asarray(R32["pairwise_dist_periodic"](tf32.convert_to_tensor(x.astype(np.float32)), 2 * math.pi))
    np.savez_compressed(OUT / "distances.npz", **g)

    # ---- sigmoid loss ------------------------------------------------------------------------
    g = {}
    cases = {
        "periodic_256x51": dict(n=256, d=51, per=2 * math.pi, sig=sig_default, hscale=None, lscale=10.0),
        "nonperiodic_256x51": dict(n=256, d=51, per=float("inf"), sig=sig_default, hscale=100.0, lscale=10.0),
        "cube_256x3": dict(n=256, d=3, per=float("inf"), sig=(0.2, 3, 6, 1, 2, 6), hscale=1.0, lscale=1.0),
        "nb_200x8": dict(n=200, d=8, per=float("inf"), sig=(0.3, 6, 6, 1, 4, 6), hscale=1.0, lscale=1.0),
        "periodic_clustered_300x64": dict(n=300, d=64, per=2 * math.pi, sig=sig_default, hscale=None, lscale=3.0, clustered=True),
        "generic_130x20": dict(n=130, d=20, per=3.0, sig=(1.3, 2.5, 3.7, 0.8, 1.7, 2.9), hscale=None, lscale=2.0),
        "latent3_150x10": dict(n=150, d=10, per=2 * math.pi, sig=sig_default, hscale=None, lscale=2.0, lat=3),
    }
    for name, c in cases.items():
        n, d = c["n"], c["d"]
        if c.get("clustered"):
            centres = rng.uniform(-math.pi, math.pi, size=(6, d))
            h = centres[rng.integers(0, 6, size=n)] + rng.normal(scale=0.05, size=(n, d))
            h = (h + math.pi) % (2 * math.pi) - math.pi
        elif c["hscale"] is None:
            h = rng.uniform(-0.5, 0.5, size=(n, d)) * (c["per"] if np.isfinite(c["per"]) else 1.0)
        else:
            h = rng.uniform(size=(n, d)) * c["hscale"]
        low = rng.uniform(size=(n, c.get("lat", 2))) * c["lscale"]
        h = h.astype(np.float32).astype(np.float64)  # float32-representable inputs
        low = low.astype(np.float32).astype(np.float64)
        f = R["sigmoid_loss"](None, periodicity_overwrite=c["per"], dist_dig_parameters_overwrite=c["sig"])
        g[f"{name}_high"], g[f"{name}_low"] = h, low
        g[f"{name}_per"] = np.float64(c["per"])
        g[f"{name}_sig"] = np.array(c["sig"], dtype=np.float64)
        g[f"{name}_loss"] = np.float64(f(tf64.convert_to_tensor(h), tf64.convert_to_tensor(low)))
        # TF1 twin must agree exactly (reference tests/test_losses.py:226,277)
        R1["pairwise_dist"], R1["pairwise_dist_periodic"] = R["pairwise_dist"], R["pairwise_dist_periodic"]
        R1["sigmoid"] = lambda r_, s_, a_, b_: R["sigmoid"](s_, a_, b_)(r_)
        if c["per"] == float("inf"):
            l1 = R1["distance_cost"](tf64.convert_to_tensor(h), tf64.convert_to_tensor(low), *c["sig"], c["per"])
            assert float(l1) == float(g[f"{name}_loss"])
        f32 = R32["sigmoid_loss"](None, periodicity_overwrite=c["per"], dist_dig_parameters_overwrite=c["sig"])
        g[f"{name}_loss_f32"] = np.float64(f32(tf32.convert_to_tensor(h.astype(np.float32)), tf32.convert_to_tensor(low.astype(np.float32))))
    np.savez_compressed(OUT / "sigmoid_loss.npz", **g)

    # ---- back-mapping ----------------------------------------------------------------------------
    g = {}
    g["tetra_7"] = np.asarray(R1["straight_tetrahedral_chain"](bond_lengths=[1, 2, 3, 1, 2, 3]))
    g["tetra_33"] = np.asarray(R1["straight_tetrahedral_chain"](33))
    for n_atoms, batch in ((9, 4), (12, 3), (30, 5), (31, 5), (300, 2)):
        dist = rng.uniform(0.13, 0.15, size=(batch, n_atoms - 1)).astype(np.float32).astype(np.float64)
        ang = rng.uniform(1.9, 2.2, size=(batch, n_atoms - 2)).astype(np.float32).astype(np.float64)
        dih = rng.uniform(-math.pi, math.pi, size=(batch, n_atoms - 3)).astype(np.float32).astype(np.float64)
        k = f"n{n_atoms}"
        g[f"{k}_dist"], g[f"{k}_ang"], g[f"{k}_dih"] = dist, ang, dih
        lengths = np.mean(dist, axis=0)[None]
        chain = R1["chain_in_plane"](tf64.convert_to_tensor(lengths), tf64.convert_to_tensor(ang))
        g[f"{k}_chain"] = np.asarray(chain)
        chain_b = R1["chain_in_plane"](tf64.convert_to_tensor(dist), tf64.convert_to_tensor(ang))
        g[f"{k}_chain_perframe_lengths"] = np.asarray(chain_b)
        left, right = n_atoms // 2 - 1, (n_atoms - 3) // 2
        d_in = tf64.convert_to_tensor(dih + math.pi)
        out_layers = R["dihedrals_to_cartesian_tf_layers"](d_in, chain, left, right)
        g[f"{k}_d2c_layers"] = np.asarray(out_layers)
        out_tf1 = R1["dihedrals_to_cartesian_tf"](d_in, chain)
        g[f"{k}_d2c_tf1"] = np.asarray(out_tf1)
        me = _Self(left_split=left, right_split=right)
        out_layer = R["BackMapLayer_call"](me, (tf64.convert_to_tensor(dist), tf64.convert_to_tensor(ang), tf64.convert_to_tensor(dih)))
        g[f"{k}_backmaplayer"] = np.asarray(out_layer)
        # float32 evaluation of the same reference code (its own error band)
        out32 = R32["BackMapLayer_call"](me, (tf32.convert_to_tensor(dist.astype(np.float32)),
                                              tf32.convert_to_tensor(ang.astype(np.float32)),
                                              tf32.convert_to_tensor(dih.astype(np.float32))))
        g[f"{k}_backmaplayer_f32"] = np.asarray(out32)
        # split index construction (integer, bit-exact)
        ia = np.arange(n_atoms)[None]
        idh = np.arange(n_atoms - 3)[None]
        cl, cr = R["split_and_reverse_cartesians"](tf64.convert_to_tensor(ia))
        dl, dr = R["split_and_reverse_dihedrals"](tf64.convert_to_tensor(idh))
        g[f"{k}_split_atoms_left"], g[f"{k}_split_atoms_right"] = np.asarray(cl)[0].astype(np.int64), np.asarray(cr)[0].astype(np.int64)
        g[f"{k}_split_dih_left"], g[f"{k}_split_dih_right"] = np.asarray(dl)[0].astype(np.int64), np.asarray(dr)[0].astype(np.int64)
    # helix KAT of the reference's (skipped) one-way test: tests/test_dihedral_to_cartesian.py:98-153
    phi = (57.8 / 180) * math.pi + math.pi
    psi = (47.0 / 180) * math.pi + math.pi
    dih = np.array([[phi, psi, 0.0] * 10] * 2)
    start = R1["straight_tetrahedral_chain"](33)
    g["helix_dih"] = dih
    g["helix_oneway"] = np.asarray(R1["dihedral_to_cartesian_tf_one_way"](tf64.convert_to_tensor(dih), tf64.convert_to_tensor(np.tile(start[None], (2, 1, 1)))))
    g["helix_twosided"] = np.asarray(R1["dihedrals_to_cartesian_tf"](tf64.convert_to_tensor(dih), tf64.convert_to_tensor(start)))
    # the reference's own 33x3 helix known-answer table (float32 literals in its test file)
    tree = ast.parse((REF / "tests/test_dihedral_to_cartesian.py").read_text())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "test_straight_to_helix_array":
            for st in node.body:
                if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "result":
                    g["helix_kat"] = np.array(ast.literal_eval(st.value.args[0].left))[0]
    assert np.abs(g["helix_oneway"][0] - g["helix_kat"]).max() < 1e-4  # reference's own atol
    # rotation matrix
    ax = rng.normal(size=(6, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    angs = rng.uniform(-math.pi, math.pi, size=6)
    g["rot_axis"], g["rot_angle"] = ax, angs
    g["rot_out"] = np.asarray(R["rotation_matrix"](tf64.convert_to_tensor(ax), tf64.convert_to_tensor(angs)))
    assert np.array_equal(g["rot_out"], np.asarray(R1["rotation_matrix"](tf64.convert_to_tensor(ax), tf64.convert_to_tensor(angs))))
    np.savez_compressed(OUT / "backmapping.npz", **g)

    # ---- layers ----------------------------------------------------------------------------------
    g = {}
    xin = rng.uniform(-math.pi, math.pi, size=(6, 9))
    g["pi_x"] = xin
    g["pi_2pi"] = np.asarray(R["PeriodicInput_call"](_Self(p=_Self(periodicity=2 * math.pi), print_name="x"), tf64.convert_to_tensor(xin)))
    g["pi_360"] = np.asarray(R["PeriodicInput_call"](_Self(p=_Self(periodicity=360.0), print_name="x"), tf64.convert_to_tensor(xin * 50)))
    xyz = rng.normal(size=(4, 30, 3))
    g["pd_xyz"] = xyz
    for tag, (a_, b_, c_) in {"ca": (1, None, 3), "all": (None, None, None), "odd": (2, 25, 4)}.items():
        p = _Self(reconstruct_sidechains=False, cartesian_pwd_start=a_, cartesian_pwd_stop=b_, cartesian_pwd_step=c_)
        g[f"pd_{tag}"] = np.asarray(R["PairwiseDistances_call"](_Self(p=p), tf64.convert_to_tensor(xyz)))
    np.savez_compressed(OUT / "layers.npz", **g)

    # ---- generation side: guessed amide H / O and the merge (own generator: the files above stay byte-identical) ----
    g = {}
    rng2 = np.random.default_rng(20261018)
    for n_atoms, batch in ((9, 3), (30, 4), (300, 2)):
        dist = rng2.uniform(0.13, 0.15, size=(batch, n_atoms - 1)).astype(np.float32).astype(np.float64)
        ang = rng2.uniform(1.9, 2.2, size=(batch, n_atoms - 2)).astype(np.float32).astype(np.float64)
        dih = rng2.uniform(-math.pi, math.pi, size=(batch, n_atoms - 3)).astype(np.float32).astype(np.float64)
        me = _Self(left_split=n_atoms // 2 - 1, right_split=(n_atoms - 3) // 2)
        xyz = np.asarray(R["BackMapLayer_call"](me, (tf64.convert_to_tensor(dist), tf64.convert_to_tensor(ang), tf64.convert_to_tensor(dih))))
        xyz = xyz.astype(np.float32).astype(np.float64)   # float32-representable backbone
        n_idx, c_idx = np.arange(n_atoms)[::3], np.arange(n_atoms)[2::3]   # reference tests/test_backmapping_em1_em2.py:571-591
        k = f"n{n_atoms}"
        g[f"{k}_xyz"] = xyz
        h = R["guess_amide_H"](tf64.convert_to_tensor(xyz), n_idx)
        o = R["guess_amide_O"](tf64.convert_to_tensor(xyz), c_idx)
        g[f"{k}_H"], g[f"{k}_O"] = np.asarray(h), np.asarray(o)
        g[f"{k}_merged"] = np.asarray(R["merge_cartesians"](tf64.convert_to_tensor(xyz), n_idx, c_idx, h, o))
        h32 = R32["guess_amide_H"](tf32.convert_to_tensor(xyz.astype(np.float32)), n_idx)
        g[f"{k}_H_f32"] = np.asarray(h32)
    # an irregular selection: arbitrary centre atoms incl. the last one (falls back to atom i - 2) and generic angle / length
    xyz = g["n30_xyz"]
    sel = np.array([1, 4, 5, 17, 28, 29])
    g["n30_sel"] = sel
    g["n30_sp2_generic"] = np.asarray(R["guess_sp2_atom"](tf64.convert_to_tensor(xyz), sel, 1.9, 0.101))
    # ---- the rotation loop of mdtraj_backmapping (misc/backmapping.py:1661-1690): its primitives are the reference's own --
    # _dihedral / _displacement (misc/rotate.py:547-601), _rotmat_jit (misc/backmapping.py:356-381, the reference's restatement of
    # transformations.rotation_matrix, which is not installed) and _get_near_and_far_networkx (misc/rotate.py:409-511), extracted
    # and executed unmodified; the eight-line loop around them is restated here line by line.
    import networkx as nx

    gen = {"np": np, "nx": nx, "Optional": None, "Union": None, "md": None, "jit": lambda *a, **k: (lambda f: f)}
    _extract(REF / "encodermap/misc/rotate.py", ["_displacement", "_dihedral", "_get_near_and_far_networkx"], gen,
             take_last=("_get_near_and_far_networkx",))
    _extract(REF / "encodermap/misc/backmapping.py", ["_rotmat_jit"], gen)
    # _rotmat_jit builds the matrix in float32 (its three dtype="float32" literals); the sequential path that actually runs
    # calls transformations.rotation_matrix, which is float64.  Both are evaluated: the function as it stands, and the same
    # source with those literals promoted.
    src = (REF / "encodermap/misc/backmapping.py").read_text()
    node = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "_rotmat_jit")
    node.decorator_list = []
    gen64 = dict(gen)
    exec(compile(ast.unparse(node).replace("'float32'", "'float64'").replace("np.float32", "np.float64"), "_rotmat_jit[float64]", "exec"), gen64)
    # a branched, ring-free "protein": backbone of 3 R atoms, an O on every C, an H on every N but the first, side chains of
    # 0..4 atoms on the CA (what mdtraj's bond graph looks like without rings)
    n_res = 12
    bonds, names = [], []
    for r in range(n_res):
        base = len(names)
        names += ["N", "CA", "C"]
        if r > 0:
            bonds.append((prev_c, base))
        bonds += [(base, base + 1), (base + 1, base + 2)]
        prev_c = base + 2
    n_bb = len(names)
    for r in range(n_res):
        n_i, ca_i, c_i = 3 * r, 3 * r + 1, 3 * r + 2
        names.append("O"); bonds.append((c_i, len(names) - 1))
        if r > 0:
            names.append("H"); bonds.append((n_i, len(names) - 1))
        parent = ca_i
        for _ in range(int(rng2.integers(0, 5))):
            names.append("S"); bonds.append((parent, len(names) - 1)); parent = len(names) - 1
    n_all = len(names)
    graph = nx.Graph()
    graph.add_nodes_from(range(n_all))
    graph.add_edges_from(bonds)
    # backbone dihedrals psi / omega / phi: quadruplets of consecutive backbone atoms, central bond = atoms 1, 2 of each
    quads = np.array([[k, k + 1, k + 2, k + 3] for k in range(n_bb - 3)])
    # side-chain dihedrals: N-CA-S1-S2, CA-S1-S2-S3, ... wherever the chain is long enough
    adj = {a: [] for a in range(n_all)}
    for a, b in bonds:
        adj[a].append(b); adj[b].append(a)
    side_quads = []
    for r in range(n_res):
        chain = [3 * r, 3 * r + 1]
        nxt = [b for b in adj[3 * r + 1] if names[b] == "S"]
        while nxt:
            chain.append(nxt[0])
            nxt = [b for b in adj[chain[-1]] if names[b] == "S" and b not in chain]
        side_quads += [chain[k:k + 4] for k in range(len(chain) - 3)]
    all_quads = np.vstack([quads, np.array(side_quads).reshape(-1, 4)])
    bond_idx = all_quads[:, 1:3]
    near_sides, far_sides = gen["_get_near_and_far_networkx"](graph, bond_idx)
    start = np.cumsum(rng2.normal(scale=0.09, size=(n_all, 3)), axis=0).astype(np.float32).astype(np.float64)
    targets = rng2.uniform(-math.pi, math.pi, size=(5, len(all_quads))).astype(np.float32).astype(np.float64)
    def run_loop(rotmat_fn):
        new_xyz = np.repeat(start[None], 5, 0)
        new_xyz = np.pad(new_xyz, ((0, 0), (0, 0), (0, 1)), mode="constant", constant_values=1)       # :1631-1633
        for i in range(targets.shape[0]):                                                            # :1673-1690
            for j in range(targets.shape[1]):
                far_side, dihedral, bond = far_sides[j], all_quads[j], bond_idx[j]
                target_angle = targets[i, j]
                current_angle = gen["_dihedral"](new_xyz[i, :, :3], dihedral)[0][0]
                angle = target_angle - current_angle
                direction = np.diff(new_xyz[i, bond, :3], axis=0).flatten()
                pivot_point = new_xyz[i, bond[0], :3]
                rotmat = rotmat_fn(angle, direction.copy(), pivot_point)
                new_xyz[i, far_side, :3] = rotmat.dot(new_xyz[i, far_side].T).T[:, :3]
        return new_xyz

    new_xyz = run_loop(gen64["_rotmat_jit"])
    new_xyz_f32rot = run_loop(gen["_rotmat_jit"])
    reached = np.array([[gen["_dihedral"](new_xyz[i, :, :3], q)[0][0] for q in all_quads] for i in range(5)])
    # every dihedral is left where it was put unless a LATER rotation moves one of its atoms relative to the others; for this
    # ordering (backbone along the chain, side chains afterwards) all of them survive
    dev = np.abs((reached - targets + math.pi) % (2 * math.pi) - math.pi)
    assert dev.max() < 1e-9, dev.max()
    assert np.abs(new_xyz_f32rot - new_xyz).max() < 2e-5
    g["sd_out_f32rot"] = new_xyz_f32rot[..., :3]
    g["sd_bonds"] = np.array(bonds)
    g["sd_quads"], g["sd_bond_idx"] = all_quads, bond_idx
    g["sd_far_offsets"] = np.concatenate([[0], np.cumsum([len(f_) for f_ in far_sides])])
    g["sd_far_atoms"] = np.concatenate([np.sort(np.asarray(f_)) for f_ in far_sides])
    g["sd_near_sizes"] = np.array([len(n_) for n_ in near_sides])
    g["sd_start"], g["sd_targets"], g["sd_out"] = start, targets, new_xyz[..., :3]
    np.savez_compressed(OUT / "generation.npz", **g)
    sidechain_section()
    for f in sorted(OUT.glob("*.npz")):
        print(f"{f.name}: {f.stat().st_size} bytes")



# ----------------------------------------------------------------------------------------
# side-chain back-mapping: BackMapLayerWithSidechains and its numpy twin
# ----------------------------------------------------------------------------------------


class _TensorArray:
    """tf.TensorArray as the layer uses it: write returns the array, read / stack."""

    work_dtype = np.float64       # the evaluation dtype of the shim that owns this class (TensorFlow: the array's dtype argument)

    def __init__(self, dtype=None, size=0, clear_after_read=False):
        self.items = {}

    def write(self, i, value):
        self.items[int(i)] = np.asarray(value, dtype=self.work_dtype if np.asarray(value).dtype.kind == "f" else None)
        return self

    def read(self, i):
        return _w(self.items[int(i)])

    def stack(self):
        return _w(np.stack([self.items[k] for k in sorted(self.items)], axis=0))


def _set_shape(self, shape):
    assert tuple(self.shape) == tuple(int(v) for v in shape), (self.shape, shape)


T.set_shape = _set_shape


class _KerasBackend:
    @staticmethod
    def batch_dot(a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.ndim == 2 and b.ndim == 2:
            return _w(np.sum(a * b, axis=1, keepdims=True))
        if a.ndim == 3 and b.ndim == 2:
            return _w(np.einsum("bij,bj->bi", a, b))
        return _w(np.einsum("bij,bjk->bik", a, b))


class _TFShimSide(_TFShim):
    """The further tf symbols BackMapLayerWithSidechains (models/layers.py:218-908) touches."""

    bool = np.bool_
    int32 = np.int32
    TensorArray = _TensorArray
    keras = type("K", (), {"backend": _KerasBackend()})()

    class linalg(_Linalg):
        @staticmethod
        def diag(x, k=0):
            x = np.asarray(x)
            out = np.zeros(x.shape + (x.shape[-1],), x.dtype)
            idx = np.arange(x.shape[-1])
            out[..., idx, idx] = x
            return _w(out)

    def __init__(self, dtype):
        super().__init__(dtype)
        self.TensorArray = type("TensorArray", (_TensorArray,), {"work_dtype": self.dtype.type})

    def constant(self, x, shape=None, dtype=None):
        a = np.asarray(x)
        if dtype is np.bool_ or dtype is np.int32:
            return _w(a.astype(dtype))
        return _w(a.astype(self.dtype))

    def zeros(self, shape, dtype=None):
        return _w(np.zeros(tuple(int(v) for v in shape), self.dtype))

    def pad(self, x, paddings, constant_values=0):
        return _w(np.pad(np.asarray(x), paddings, mode="constant", constant_values=constant_values))

    def repeat(self, x, repeats, axis=None):
        return _w(np.repeat(np.asarray(x), int(repeats), axis=axis))

    def gather(self, params, indices, axis=None, batch_dims=0):
        assert batch_dims == 0
        return _w(np.take(np.asarray(params), np.asarray(indices), axis=axis))

    def where(self, c, a=None, b=None):
        if a is None:
            return _w(np.argwhere(np.asarray(c)))
        return _w(np.where(c, a, b))

    def clip_by_value(self, x, clip_value_min, clip_value_max):
        return _w(np.clip(x, clip_value_min, clip_value_max))

    def squeeze(self, x):
        return _w(np.squeeze(x))

    def acos(self, x):
        return _w(np.arccos(x))

    def atan2(self, y, x):
        return _w(np.arctan2(y, x))

    def einsum(self, eq, *xs):
        return _w(np.einsum(eq, *[np.asarray(x) for x in xs]))

    def transpose(self, x, perm=None):
        return _w(np.transpose(np.asarray(x), perm))


def _class_methods(path, cls_name, wanted, ns, drop_super=True):
    tree = ast.parse(path.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in wanted:
            fn.name = f"{cls_name}_{fn.name.strip('_')}"
            fn.decorator_list, fn.returns = [], None
            for a in fn.args.args:
                a.annotation = None
            if drop_super:   # `super().__init__()` needs the class cell; the Keras base initialiser does nothing the path needs
                fn.body = [st for st in fn.body if "super()" not in ast.unparse(st)]
            mod = ast.Module(body=[fn], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, str(path), "exec"), ns)


SIDECHAIN_CASES = {
    "metlysgly": [3, 4, 0],                                  # the docstring example of _full_backmapping_np
    "first_empty": [0, 2, 1, 4],                             # residue 1 without side chain (the other admissible end)
    "twelve": [2, 0, 4, 1, 0, 0, 3, 2, 1, 4, 2, 0],          # interior glycines after a side chain
    "ub_like": [3, 2, 2, 4, 1, 2, 1, 1, 2, 0, 4, 1, 2, 2, 2, 2, 1, 2, 1, 1, 2, 1, 2, 2, 1, 2, 4, 0, 4, 2, 3, 2, 4, 2, 0, 2, 1,
                1, 2, 3, 3, 4, 2, 2, 2, 0, 0, 4, 3, 2, 2, 2, 0, 4, 1, 2, 1, 2, 4, 2, 2, 3, 4, 2, 1, 1, 2, 4, 2, 1, 2, 4, 2, 4, 0, 0],
}


def sidechain_section():
    import itertools
    import types

    from scipy.linalg import block_diag

    g = {}
    rng = np.random.default_rng(20261019)
    tf = _TFShimSide(np.float64)
    ns = {"tf": tf, "np": np, "itertools": itertools, "block_diag": block_diag, "Any": None}
    _extract(REF / "encodermap/models/layers.py", ["_batch_fro", "_rotation_matrices", "_unit_vector"], ns)
    _class_methods(REF / "encodermap/models/layers.py", "BackMapLayerWithSidechains", ("__init__", "call"), ns)
    _class_methods(REF / "encodermap/models/layers.py", "PairwiseDistances", ("__init__",), ns)
    tf32 = _TFShimSide(np.float32)
    ns32 = {"tf": tf32, "np": np, "itertools": itertools, "block_diag": block_diag, "Any": None}
    _extract(REF / "encodermap/models/layers.py", ["_batch_fro", "_rotation_matrices", "_unit_vector"], ns32)
    _class_methods(REF / "encodermap/models/layers.py", "BackMapLayerWithSidechains", ("__init__", "call"), ns32)

    # the numpy twin imports matplotlib, transformations and encodermap.misc.rotate inside its body: stand-ins for the three.
    # transformations.rotation_matrix (C. Gohlke, not installed) = the reference's own restatement _rotmat_jit
    # (misc/backmapping.py:356-381) with its float32 literals promoted, as in the generation section above.
    gen = {"np": np, "Optional": None, "Union": None}
    _extract(REF / "encodermap/misc/rotate.py", ["_displacement", "_dihedral"], gen)
    src = (REF / "encodermap/misc/backmapping.py").read_text()
    node = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "_rotmat_jit")
    node.decorator_list = []
    exec(compile(ast.unparse(node).replace("'float32'", "'float64'").replace("np.float32", "np.float64"), "_rotmat_jit[float64]", "exec"), gen)

    class _Ax:
        def plot(self, *a, **k):
            pass

    class _Fig:
        def savefig(self, *a, **k):
            pass

    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = lambda **k: (_Fig(), (_Ax(), _Ax(), _Ax()))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    tr = types.ModuleType("transformations")
    tr.rotation_matrix = lambda angle, direction, point: gen["_rotmat_jit"](
        float(angle), np.array(direction, dtype=np.float64), np.asarray(point, dtype=np.float64))
    em = types.ModuleType("encodermap")
    em_misc = types.ModuleType("encodermap.misc")
    em_rot = types.ModuleType("encodermap.misc.rotate")
    em_rot._dihedral = gen["_dihedral"]
    fakes = {"matplotlib": mpl, "matplotlib.pyplot": plt, "transformations": tr, "encodermap": em, "encodermap.misc": em_misc,
             "encodermap.misc.rotate": em_rot}
    saved = {k: sys.modules.get(k) for k in fakes}
    sys.modules.update(fakes)
    nsn = {"np": np, "Sequence": None, "Union": None, "Literal": None, "BytesIO": None, "overload": lambda f: f}
    _extract(REF / "encodermap/misc/backmapping.py", ["_full_backmapping_np"], nsn, take_last=("_full_backmapping_np",))

    try:
        for tag, counts in SIDECHAIN_CASES.items():
            fd = {-1: {k + 1: int(v) for k, v in enumerate(counts)}}
            n_res = len(counts)
            n_side = sum(v + 1 for v in counts if v > 0)
            batch = 2 if tag == "ub_like" else 3
            f32 = lambda a: a.astype(np.float32).astype(np.float64)
            inputs = [
                f32(rng.uniform(0.13, 0.16, size=(batch, 3 * n_res - 1))),
                f32(rng.uniform(1.85, 2.25, size=(batch, 3 * n_res - 2))),
                f32(rng.uniform(-math.pi, math.pi, size=(batch, 3 * n_res - 3))),
                f32(rng.uniform(0.13, 0.19, size=(batch, n_side))),
                f32(rng.uniform(1.80, 2.20, size=(batch, n_side))),
                f32(rng.uniform(-math.pi, math.pi, size=(batch, sum(counts)))),
            ]
            layer = _Self()
            ns["BackMapLayerWithSidechains_init"](layer, fd)
            g[f"{tag}_counts"] = np.array(counts, dtype=np.int32)
            for k, v in zip(("cd", "ca", "cdih", "sd", "sa", "sdih"), inputs):
                g[f"{tag}_in_{k}"] = v
            g[f"{tag}_central_mask"] = np.asarray(layer.central_distance_indices)
            g[f"{tag}_central_angle_mask"] = np.asarray(layer.central_angle_indices)
            g[f"{tag}_side_angle_mask"] = np.asarray(layer.side_angle_indices)
            g[f"{tag}_dihedral_mask"] = np.asarray(layer.dihedral_indices)
            g[f"{tag}_central_angle_triplets"] = np.asarray(layer.central_angle_index_triplets)
            g[f"{tag}_side_angle_triplets"] = np.asarray(layer.sidechain_angle_index_triplets)
            g[f"{tag}_dihedral_quadruplets"] = np.asarray(layer.dihedral_index_quadruplets)
            out_layer = np.asarray(ns["BackMapLayerWithSidechains_call"](layer, tuple(tf.convert_to_tensor(v) for v in inputs)))
            out_np, _, idx = nsn["_full_backmapping_np"](fd, *inputs, return_indices=True)
            # the TensorFlow layer and the numpy twin are the same algorithm; they differ by ~1e-8 nm even in float64 because
            # every bond angle is measured on a still straight triplet, where acos turns a 1e-16 rounding difference of the
            # cosine into sqrt(2e-16) = 1.4e-8 rad (in float32, as the layer runs in the reference: 3.5e-4 rad)
            assert np.abs(out_layer - out_np).max() < 1e-6, np.abs(out_layer - out_np).max()
            g[f"{tag}_out"] = out_layer
            # the same layer body evaluated in float32, as the reference runs it: how far its own output is from the float64
            # evaluation (every bond angle measured on a straight triplet costs acos(-1 + 6e-8) = pi - 3.5e-4)
            layer32 = _Self()
            ns32["BackMapLayerWithSidechains_init"](layer32, fd)
            out32 = np.asarray(ns32["BackMapLayerWithSidechains_call"](layer32, tuple(tf32.convert_to_tensor(v.astype(np.float32)) for v in inputs)))
            assert out32.dtype == np.float32, out32.dtype
            g[f"{tag}_out_f32"] = out32
            g[f"{tag}_out_np"] = np.asarray(out_np)
            for k, v in idx.items():
                g[f"{tag}_np_{k}"] = np.asarray(v)
            # the atom selection of PairwiseDistances when side chains are reconstructed (models/layers.py:1188-1208)
            for sel_tag, (a, b, c) in {"ca": (1, None, 3), "all": (None, None, None)}.items():
                pl = _Self(p=_Self(reconstruct_sidechains=True, sidechain_info=fd, cartesian_pwd_start=a, cartesian_pwd_stop=b,
                                   cartesian_pwd_step=c))
                ns["PairwiseDistances_init"](pl, pl.p, "x")
                g[f"{tag}_pwd_indices_{sel_tag}"] = np.asarray(pl.indices)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    np.savez_compressed(OUT / "sidechains.npz", **g)

if __name__ == "__main__":
    if not REF.exists():
        sys.exit(f"{REF} not present: golden vectors can only be regenerated in the build container")
    main()
