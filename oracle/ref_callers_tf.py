"""TEST INFRASTRUCTURE -- an importable ``encodermap`` package skeleton holding the CALLERS of the hot path, so that
``encodermap_b200.tf_adapter.install()`` has the reference's module layout to rebind and the callers can be run.

Two sources for the caller bodies:

* ``mode="reference"`` (build container only): the functions / ``call`` methods are pulled out of ``/root/reference``
  with ``ast`` and compiled in memory, unmodified, straight from where they lie (nothing is copied into the repo) --
  the same technique as ``tools/gen_golden.py``;
* ``mode="restated"`` (everywhere, incl. the GPU box where ``/root/reference`` does not exist): the same callers
  restated below against the ``tf`` API, each citing the reference lines it follows.
  ``tests/test_tf_adapter.py::test_restated_callers_match_reference_bodies`` pins the restatement to the extracted
  bodies on CPU (identical values in float64).

In both modes every module also holds the reference's names of the hot OPS (``pairwise_dist``, ``chain_in_plane``, ...):
real bodies in reference mode, and in restated mode placeholders that raise ``UnpatchedReferenceOp`` -- after
``install()`` none of them may be reached, which is exactly what the GPU test asserts by running to completion.
"""
from __future__ import annotations

import ast
import sys
import types
from math import pi
from pathlib import Path

REF = Path("/root/reference")

MODULES = ["encodermap", "encodermap.misc", "encodermap.misc.distances", "encodermap.misc.backmapping", "encodermap.encodermap_tf1",
           "encodermap.encodermap_tf1.backmapping", "encodermap.loss_functions", "encodermap.loss_functions.loss_functions",
           "encodermap.models", "encodermap.models.layers", "encodermap.models.models", "encodermap.parameters",
           "encodermap.parameters.parameters"]


class UnpatchedReferenceOp(RuntimeError):
    pass


def _placeholder(name):
    def op(*a, **kw):
        raise UnpatchedReferenceOp(f"{name}: the reference's own TensorFlow implementation was reached -- tf_adapter.install() did not rebind it")

    op.__name__ = name
    return op


# ---- restated callers (source text, compiled into the skeleton modules so that their globals ARE those modules) -----------
RESTATED = {
    # encodermap/loss_functions/loss_functions.py:200-298 (distance_loss), :301-369 (sigmoid_loss), :873-944 (cartesian_distance_loss)
    "encodermap.loss_functions.loss_functions": '''
def _do_nothing():
    pass


def _summary_cost(name, cost):
    pass   # tf.summary.scalar inside a name scope: stays in the host framework


def sigmoid_loss(parameters=None, periodicity_overwrite=None, dist_dig_parameters_overwrite=None):
    p = Parameters() if parameters is None else parameters
    periodicity = periodicity_overwrite if periodicity_overwrite is not None else p.periodicity
    dist_sig_parameters = dist_dig_parameters_overwrite if dist_dig_parameters_overwrite is not None else p.dist_sig_parameters

    def sigmoid_loss_func(y_true, y_pred):
        if periodicity == float("inf"):
            dist_h = pairwise_dist(y_true)
        else:
            dist_h = pairwise_dist_periodic(y_true, periodicity)
        dist_l = pairwise_dist(y_pred)
        sig_h = sigmoid(*dist_sig_parameters[:3])(dist_h)
        sig_l = sigmoid(*dist_sig_parameters[3:])(dist_l)
        cost = tf.reduce_mean(tf.square(sig_h - sig_l))
        tf.debugging.assert_all_finite(cost, message="Sigmoid cost became infinite or NaN.")
        return cost

    return sigmoid_loss_func


def distance_loss(model, parameters=None, callback=None):
    p = Parameters() if parameters is None else parameters
    latent = model.encoder                      # both branches of the reference's layer-count check pick model.encoder
    write_bool = K.constant(False, "bool", name="log_bool") if callback is None else callback.log_bool
    dist_loss = sigmoid_loss(p)                 # module global: captured at construction (:263)

    def distance_loss_func(y_true, y_pred=None):
        distance_loss_func.name = "distance_loss"
        y_pred = latent(y_true, training=True)
        if isinstance(y_true, tuple):           # functional model: (angles, dihedrals[, side dihedrals])
            y_true = tf.concat(y_true[:3], axis=1)
        if p.distance_cost_scale is not None:
            dist_cost = dist_loss(y_true, y_pred)
            dist_cost *= p.distance_cost_scale
        else:
            dist_cost = 0.0
        tf.cond(write_bool, true_fn=lambda: _summary_cost("Distance Cost", dist_cost), false_fn=lambda: _do_nothing(), name="Cost")
        tf.debugging.assert_all_finite(dist_cost, message="Dist cost became infinite or NaN.")
        return dist_cost

    return distance_loss_func


def cartesian_distance_loss(model, parameters=None, callback=None):
    p = ADCParameters() if parameters is None else parameters
    write_bool = K.constant(False, "bool", name="log_bool") if callback is None else callback.log_bool
    dist_loss = sigmoid_loss(p, periodicity_overwrite=float("inf"), dist_dig_parameters_overwrite=p.cartesian_dist_sig_parameters)

    def cartesian_distance_loss_func(y_true, y_pred):
        cartesian_distance_loss_func.name = "cartesian_distance_loss"
        if p.cartesian_distance_cost_scale is not None:
            dist_cost = dist_loss(y_true, y_pred)
            dist_cost *= p.cartesian_distance_cost_scale
        else:
            dist_cost = 0.0
        tf.cond(write_bool, true_fn=lambda: _summary_cost("Cartesian Distance Cost", dist_cost), false_fn=lambda: _do_nothing(), name="Cost")
        tf.debugging.assert_all_finite(dist_cost, message="Cartesian distance cost became infinite or NaN.")
        return dist_cost

    return cartesian_distance_loss_func
''',
    # encodermap/models/layers.py:204-215 (PeriodicInput.call), :957-986 (BackMapLayer.call), :1252-1267 (PairwiseDistances.call)
    "encodermap.models.layers": '''
class PeriodicInput(Layer):
    def __init__(self, parameters, print_name, trainable=False):
        self.p, self.print_name = parameters, print_name

    def call(self, inputs):
        outputs = inputs
        if self.p.periodicity != 2 * pi:
            outputs = outputs / self.p.periodicity * 2 * pi
        outputs = Concatenate(axis=1, name=f"{self.print_name}_Concat")([tf.sin(outputs), tf.cos(outputs)])
        return outputs


class BackMapLayer(Layer):
    def __init__(self, left_split, right_split):
        self.left_split, self.right_split = left_split, right_split

    def call(self, inputs):
        distances, angles, dihedrals = inputs
        out = tf.expand_dims(tf.reduce_mean(distances, 0), 0)     # mean bond lengths over the batch (:970)
        out = chain_in_plane(out, angles)
        out_dihedrals = tf.add(dihedrals, pi)
        out = dihedrals_to_cartesian_tf_layers(out_dihedrals, out, left_iteration_counter=self.left_split,
                                               right_iteration_counter=self.right_split)
        return out


class BackMapLayerWithSidechains(Layer):
    # the constructor's index tables (:234-500) only feed `call`, and `call` (:533-843) IS the hot op: a placeholder here, the
    # reference's own body is exercised by tools/gen_golden.py (tests/golden/sidechains.npz)
    def __init__(self, feature_description):
        self.feature_description = feature_description

    call = _placeholder("encodermap.models.layers.BackMapLayerWithSidechains.call")


class PairwiseDistances(Layer):
    def __init__(self, parameters, print_name, trainable=False):
        self.p, self.print_name = parameters, print_name
        if self.p.reconstruct_sidechains:                                   # :1188-1208
            n_residues = max(list(self.p.sidechain_info[-1].keys()))
            self.indices = np.arange(n_residues * 3)[self.p.cartesian_pwd_start : self.p.cartesian_pwd_stop : self.p.cartesian_pwd_step]
            atom = n_residues * 3 + 1
            indices = []
            for residue, n_sidechains_in_residue in self.p.sidechain_info[-1].items():
                if n_sidechains_in_residue == 0:
                    continue
                atom += n_sidechains_in_residue
                indices.append(atom)
            self.indices = np.concatenate([self.indices, indices])

    def call(self, inputs):
        if not self.p.reconstruct_sidechains:
            out = inputs[:, self.p.cartesian_pwd_start : self.p.cartesian_pwd_stop : self.p.cartesian_pwd_step]
        else:
            out = tf.gather(params=inputs, indices=self.indices, axis=1, batch_dims=0)
        out = pairwise_dist(out, flat=True)
        return out
''',
}

# which hot-op names each module of the reference holds (its own definitions or `from ... import` bindings)
OP_NAMES = {
    "encodermap.misc.distances": ["sigmoid", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"],
    "encodermap.misc.backmapping": ["dihedrals_to_cartesian_tf_layers", "dihedral_to_cartesian_tf_one_way_layers", "rotation_matrix",
                                    "split_and_reverse_dihedrals", "split_and_reverse_cartesians", "guess_sp2_atom", "guess_amide_H",
                                    "guess_amide_O", "merge_cartesians"],
    "encodermap.encodermap_tf1.backmapping": ["chain_in_plane", "dihedrals_to_cartesian_tf", "dihedral_to_cartesian_tf_one_way"],
    "encodermap.loss_functions.loss_functions": ["sigmoid", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"],
    "encodermap.models.layers": ["pairwise_dist", "chain_in_plane", "dihedrals_to_cartesian_tf_layers"],
    "encodermap.models.models": ["pairwise_dist", "chain_in_plane", "dihedrals_to_cartesian_tf"],
}
# where each op is DEFINED in the reference (file, names) -- reference mode extracts them from there
OP_SOURCES = {
    "encodermap.misc.distances": ("encodermap/misc/distances.py", ["sigmoid", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"]),
    "encodermap.misc.backmapping": ("encodermap/misc/backmapping.py",
                                    ["split_and_reverse_dihedrals", "split_and_reverse_cartesians", "dihedrals_to_cartesian_tf_layers",
                                     "dihedral_to_cartesian_tf_one_way_layers", "rotation_matrix", "guess_sp2_atom", "guess_amide_H",
                                     "guess_amide_O", "merge_cartesians"]),
    "encodermap.encodermap_tf1.backmapping": ("encodermap/encodermap_tf1/backmapping.py",
                                              ["chain_in_plane", "dihedrals_to_cartesian_tf", "dihedral_to_cartesian_tf_one_way"]),
}
CALLER_SOURCES = {
    "encodermap.loss_functions.loss_functions": ("encodermap/loss_functions/loss_functions.py",
                                                 ["sigmoid_loss", "distance_loss", "cartesian_distance_loss"]),
}
LAYER_CALLS = ("encodermap/models/layers.py", ["PeriodicInput", "BackMapLayer", "PairwiseDistances"])


def _strip(node):
    node.decorator_list = []
    node.returns = None
    for a in node.args.args + node.args.kwonlyargs:
        a.annotation = None
    return node


def _extract_functions(path: Path, names, ns):
    tree = ast.parse(path.read_text())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            found[node.name] = _strip(node)
    missing = set(names) - set(found)
    assert not missing, f"{path}: {missing} not found"
    mod = ast.Module(body=[found[n] for n in names], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, str(path), "exec"), ns)


def _extract_layer_calls(path: Path, class_names, ns):
    """class skeletons (constructor = the attributes `call` reads) with the reference's own ``call`` body"""
    tree = ast.parse(path.read_text())
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in class_names]:
        call = next(f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "call")
        mod = ast.Module(body=[_strip(call)], type_ignores=[])
        ast.fix_missing_locations(mod)
        tmp = {}
        exec(compile(mod, str(path), "exec"), ns, tmp)
        if cls.name == "BackMapLayer":
            def init(self, left_split, right_split):
                self.left_split, self.right_split = left_split, right_split
        elif cls.name == "PairwiseDistances":
            # the reference's own constructor (the side-chain atom selection, :1188-1208) minus its Keras base-class call
            ref_init = next(f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "__init__")
            ref_init.body = [st for st in ref_init.body if "super()" not in ast.unparse(st) and not isinstance(st, ast.Expr)]
            ref_init.name = "ref_init"
            imod = ast.Module(body=[_strip(ref_init)], type_ignores=[])
            ast.fix_missing_locations(imod)
            itmp = {}
            exec(compile(imod, str(path), "exec"), ns, itmp)

            def init(self, parameters, print_name, trainable=False, _ref_init=itmp["ref_init"]):
                self.p, self.print_name = parameters, print_name
                _ref_init(self, parameters, print_name, trainable)
        else:
            def init(self, parameters, print_name, trainable=False):
                self.p, self.print_name = parameters, print_name
        ns[cls.name] = type(cls.name, (ns["Layer"],), {"__init__": init, "call": tmp["call"]})


class _P:
    """Parameters / ADCParameters with the reference's defaults for the fields the callers read
    (encodermap/parameters/parameters.py:611-638, 794-828)."""

    def __init__(self, **kw):
        self.periodicity = 2 * pi
        self.dist_sig_parameters = (4.5, 12, 6, 1, 2, 6)
        self.distance_cost_scale = 500
        self.cartesian_dist_sig_parameters = (4.5, 12, 6, 1, 2, 6)
        self.cartesian_distance_cost_scale = 1
        self.cartesian_pwd_start = self.cartesian_pwd_stop = self.cartesian_pwd_step = None
        self.reconstruct_sidechains = False
        self.__dict__.update(kw)


def build(tf, mode: str = "restated") -> dict:
    """Create the skeleton package in ``sys.modules`` and return {name: module}."""
    assert mode in ("restated", "reference")
    if mode == "reference" and not REF.exists():
        raise FileNotFoundError(f"{REF} is only present in the build container")
    mods = {}
    for name in MODULES:
        m = types.ModuleType(name)
        m.__path__ = []   # a package: sub-module imports resolve through sys.modules
        mods[name] = m
        sys.modules[name] = m
    for name, m in mods.items():
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(mods[parent], child, m)
    common = {"tf": tf, "pi": pi, "np": __import__("numpy"), "K": tf.keras.backend, "Layer": tf.keras.layers.Layer,
              "Concatenate": tf.keras.layers.Concatenate, "Parameters": _P, "ADCParameters": _P, "_placeholder": _placeholder,
              "cos": __import__("math").cos, "sin": __import__("math").sin,
              # typing names that appear in nested annotations / decorators of the reference's bodies
              "overload": (lambda f: f), "Union": None, "Number": None, "Callable": None, "Optional": None, "Sequence": None}
    mods["encodermap.parameters.parameters"].__dict__.update(Parameters=_P, ADCParameters=_P)
    for m in mods.values():
        m.__dict__.update(common)
    # hot ops: real bodies (reference mode) or placeholders
    if mode == "reference":
        for modname, (rel, names) in OP_SOURCES.items():
            _extract_functions(REF / rel, names, mods[modname].__dict__)
        tf1 = mods["encodermap.encodermap_tf1.backmapping"].__dict__
        misc_ns = {"tf": tf, "np": common["np"]}
        _extract_functions(REF / "encodermap/encodermap_tf1/misc.py", ["rotation_matrix"], misc_ns)
        tf1["rotation_matrix"] = misc_ns["rotation_matrix"]
    else:
        for modname, (_, names) in OP_SOURCES.items():
            for n in names:
                mods[modname].__dict__[n] = _placeholder(f"{modname}.{n}")
    # `from ... import` bindings of the ops into the caller modules
    defined_in = {n: m for m, (_, names) in OP_SOURCES.items() for n in names}
    for modname, names in OP_NAMES.items():
        for n in names:
            if n not in mods[modname].__dict__ or modname not in OP_SOURCES:
                mods[modname].__dict__[n] = mods[defined_in[n]].__dict__[n]
    # callers
    if mode == "reference":
        for modname, (rel, names) in CALLER_SOURCES.items():
            ns = mods[modname].__dict__
            ns["_do_nothing"] = lambda: None
            ns["_summary_cost"] = lambda name, cost: None
            _extract_functions(REF / rel, names, ns)
        _extract_layer_calls(REF / LAYER_CALLS[0], LAYER_CALLS[1], mods["encodermap.models.layers"].__dict__)
    else:
        for modname, src in RESTATED.items():
            exec(compile(src, f"<restated {modname}>", "exec"), mods[modname].__dict__)
    for n in ("BackMapLayer", "PairwiseDistances", "PeriodicInput", "BackMapLayerWithSidechains"):
        if n in mods["encodermap.models.layers"].__dict__:
            mods["encodermap.models.models"].__dict__[n] = mods["encodermap.models.layers"].__dict__[n]
    return mods


def remove() -> None:
    for name in MODULES:
        sys.modules.pop(name, None)
