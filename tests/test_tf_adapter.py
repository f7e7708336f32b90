"""The TensorFlow half of the drop-in boundary (encodermap_b200/tf_adapter.py), executed without TensorFlow.

``oracle/fake_tf.py`` stands in for the ``tensorflow`` module (torch-backed, TF semantics for the symbols touched) and
``oracle/ref_callers_tf.py`` provides an importable ``encodermap`` skeleton holding the CALLERS of the hot path:

* CPU, build container only: the restated callers are pinned to the reference's own function bodies (extracted with
  ``ast`` from /root/reference and run unmodified);
* CPU, everywhere: ``install()`` rebinds every import site and nothing computes on the CPU;
* GPU: callers + ``install()`` run against libemk and match the float64 oracle, value and gradient; in restated mode
  the reference's own hot ops are placeholders that raise, so reaching the end proves nothing fell back to them.
"""
import math

import numpy as np
import pytest
import torch

from oracle import em_oracle as O
from oracle import fake_tf, ref_callers_tf

pi = math.pi
SIG = (4.5, 12, 6, 1, 2, 6)


@pytest.fixture()
def tf_env():
    """fake tensorflow registered, adapter bound to it; everything is undone afterwards"""
    from encodermap_b200 import tf_adapter

    tf = fake_tf.install()
    tf_adapter._reset_tf_binding()
    yield tf, tf_adapter
    ref_callers_tf.remove()
    fake_tf.set_default_device("cpu")
    fake_tf.set_work_dtype(torch.float32)
    fake_tf.uninstall()
    tf_adapter._reset_tf_binding()


class _Encoder:
    """model.encoder of the loss factories: tanh(concat(inputs) @ w)"""

    def __init__(self, tf, w):
        self.tf, self.w = tf, w
        self.layers = [None, None]

    def encoder(self, x, training=False):
        if isinstance(x, tuple):
            x = self.tf.concat(x[:3], axis=1)
        return torch.tanh(x @ self.w)


def _oracle_ops_as_tf():
    """the oracle's restatement of the hot ops, callable on fake-tf tensors (float64, torch autograd)"""
    return {
        "sigmoid": O.sigmoid, "periodic_distance": O.periodic_distance, "pairwise_dist_periodic": O.pairwise_dist_periodic,
        "pairwise_dist": O.pairwise_dist, "chain_in_plane": O.chain_in_plane,
        "dihedrals_to_cartesian_tf_layers": O.dihedrals_to_cartesian_layers,
        "dihedrals_to_cartesian_tf": O.dihedrals_to_cartesian_tf1,
    }


@pytest.mark.skipif(not ref_callers_tf.REF.exists(), reason="/root/reference is only present in the build container")
def test_restated_callers_match_reference_bodies(tf_env):
    """The callers restated in oracle/ref_callers_tf.py give the same numbers as the reference's own bodies (extracted,
    unmodified) -- both on the fake tf, float64, CPU; hot ops: the reference's own in one run, the oracle's in the other."""
    tf, _ = tf_env
    fake_tf.set_work_dtype(torch.float64)
    rng = np.random.default_rng(3)
    n_atoms, b = 12, 6
    ang = rng.uniform(1.9, 2.2, (b, n_atoms - 2))
    dih = rng.uniform(-pi, pi, (b, n_atoms - 3))
    dist = rng.uniform(0.13, 0.15, (b, n_atoms - 1))
    w0 = rng.normal(size=(2 * n_atoms - 5, 2)) * 0.3
    pair = np.abs(rng.normal(size=(b, 15)))
    side_xyz = np.random.default_rng(4).normal(size=(2, 17, 3))
    results = {}
    for mode in ("reference", "restated"):
        mods = ref_callers_tf.build(tf, mode)
        if mode == "restated":
            for m in mods.values():
                for name, fn in _oracle_ops_as_tf().items():
                    if name in m.__dict__:
                        m.__dict__[name] = fn
        lf, layers = mods["encodermap.loss_functions.loss_functions"], mods["encodermap.models.layers"]
        w = tf.Variable(w0)
        model = _Encoder(tf, w)
        p = ref_callers_tf._P(cartesian_dist_sig_parameters=(0.5, 6, 6, 1, 2, 6), cartesian_distance_cost_scale=3.0,
                              cartesian_pwd_start=1, cartesian_pwd_step=3)
        out = {}
        f = lf.distance_loss(model, p)
        assert f.__name__ == "distance_loss_func"
        loss = f((tf.convert_to_tensor(ang), tf.convert_to_tensor(dih)))
        out["distance_loss"] = loss.item()
        out["distance_loss_grad"] = torch.autograd.grad(loss, w)[0].numpy()
        fc = lf.cartesian_distance_loss(model, p)
        assert fc.__name__ == "cartesian_distance_loss_func"
        z = tf.Variable(rng.normal(size=(b, 2)) if mode == "reference" else results["reference"]["z"])
        out["z"] = z.detach().numpy().copy()
        lc = fc(tf.convert_to_tensor(pair), z)
        out["cartesian_distance_loss"] = lc.item()
        out["cartesian_distance_loss_grad"] = torch.autograd.grad(lc, z)[0].numpy()
        assert lf.distance_loss(model, ref_callers_tf._P(distance_cost_scale=None))((tf.convert_to_tensor(ang), tf.convert_to_tensor(dih))) == 0.0
        xyz = layers.BackMapLayer(n_atoms // 2 - 1, (n_atoms - 3) // 2)((tf.convert_to_tensor(dist), tf.convert_to_tensor(ang), tf.convert_to_tensor(dih)))
        out["backmap"] = xyz.numpy()
        out["pairwise"] = layers.PairwiseDistances(p, "pd")(xyz).numpy()
        out["periodic_input"] = layers.PeriodicInput(ref_callers_tf._P(periodicity=360.0), "pi")(tf.convert_to_tensor(dih * 50)).numpy()
        # PairwiseDistances with reconstructed side chains: the atom selection of its constructor and the gather of its call
        ps = ref_callers_tf._P(cartesian_pwd_start=1, cartesian_pwd_step=3, reconstruct_sidechains=True,
                               sidechain_info={-1: {1: 2, 2: 0, 3: 1, 4: 0}})
        pds = layers.PairwiseDistances(ps, "pds")
        out["side_indices"] = np.asarray(pds.indices)
        out["side_pairwise"] = pds(tf.convert_to_tensor(side_xyz)).numpy()
        results[mode] = out
        ref_callers_tf.remove()
    for key, ref in results["reference"].items():
        np.testing.assert_allclose(results["restated"][key], ref, rtol=1e-9, atol=1e-12, err_msg=key)


def test_install_rebinds_every_import_site(tf_env):
    tf, adapter = tf_env
    mods = ref_callers_tf.build(tf, "restated")
    before = {k: dict(m.__dict__) for k, m in mods.items()}
    calls_before = {c: mods["encodermap.models.layers"].__dict__[c].__dict__["call"] for c in ("PeriodicInput", "PairwiseDistances", "BackMapLayer")}
    with pytest.raises(RuntimeError, match="ENCODERMAP_ENABLE_GPU"):
        adapter.install()
    saved = adapter.install(require_gpu_env=False)
    lf, d = mods["encodermap.loss_functions.loss_functions"], mods["encodermap.misc.distances"]
    layers, models = mods["encodermap.models.layers"], mods["encodermap.models.models"]
    for mod, names in ((d, ["sigmoid", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"]),
                       (lf, ["sigmoid_loss", "sigmoid", "periodic_distance", "pairwise_dist_periodic", "pairwise_dist"]),
                       (layers, ["pairwise_dist", "chain_in_plane", "dihedrals_to_cartesian_tf_layers"]),
                       (models, ["pairwise_dist", "chain_in_plane", "dihedrals_to_cartesian_tf"]),
                       (mods["encodermap.encodermap_tf1.backmapping"], ["chain_in_plane", "dihedrals_to_cartesian_tf", "dihedral_to_cartesian_tf_one_way"]),
                       (mods["encodermap.misc.backmapping"], ["dihedrals_to_cartesian_tf_layers", "dihedral_to_cartesian_tf_one_way_layers", "rotation_matrix"])):
        for name in names:
            assert getattr(mod, name) is getattr(adapter, name), f"{mod.__name__}.{name} was not rebound"
    for cls in ("PeriodicInput", "PairwiseDistances", "BackMapLayer"):
        assert getattr(layers, cls).call is not calls_before[cls]
        assert getattr(models, cls) is getattr(layers, cls)      # the class object models.py imported is the patched one
    # the loss factories of the reference now build on the adapter's sigmoid_loss (module global, captured at construction)
    rng = np.random.default_rng(0)
    w = tf.Variable(rng.normal(size=(9, 2)).astype(np.float32))
    f = lf.distance_loss(_Encoder(tf, w), ref_callers_tf._P())
    # ... and there is no CPU fallback behind it: CPU tensors are refused by libemk's Python layer
    from encodermap_b200 import EmkError

    with pytest.raises(EmkError):
        f(tf.convert_to_tensor(rng.normal(size=(10, 9)).astype(np.float32)))
    # python scalars through sigmoid stay on the host like the reference's closure (tests/test_pairwise_distances.py:191-193)
    assert d.sigmoid(4.5, 12, 6)(4.5) == pytest.approx(0.5)
    adapter.uninstall(saved)
    for k, m in mods.items():
        for name, obj in before[k].items():
            if callable(obj) and not isinstance(obj, type):
                assert m.__dict__[name] is obj, f"{k}.{name} not restored"
    for cls in calls_before:
        assert getattr(layers, cls).call is calls_before[cls]


# ---------------------------------------------------------------------------------------------------
# GPU: callers + install() against libemk
# ---------------------------------------------------------------------------------------------------
def _relnorm(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture()
def tf_gpu(tf_env, cuda_device):
    tf, adapter = tf_env
    fake_tf.set_default_device("cuda:0")
    mods = ref_callers_tf.build(tf, "restated")     # the reference's own hot ops are placeholders that raise
    saved = adapter.install(require_gpu_env=False)
    yield tf, adapter, mods
    adapter.uninstall(saved)


@pytest.mark.gpu
def test_tf_adapter_loss_closures_on_gpu(tf_gpu):
    """distance_loss / cartesian_distance_loss of the reference (caller bodies unchanged) on the rebound sigmoid_loss."""
    tf, adapter, mods = tf_gpu
    lf = mods["encodermap.loss_functions.loss_functions"]
    rng = np.random.default_rng(4)
    ang = rng.uniform(1.9, 2.2, (180, 28)).astype(np.float32)
    dih = rng.uniform(-pi, pi, (180, 27)).astype(np.float32)
    w0 = (rng.normal(size=(55, 2)) * 0.3).astype(np.float32)
    w = tf.Variable(w0)
    f = lf.distance_loss(_Encoder(tf, w), ref_callers_tf._P())
    with tf.GradientTape() as tape:
        loss = f((tf.convert_to_tensor(ang), tf.convert_to_tensor(dih)))
    gw = tape.gradient(loss, w)
    wd = torch.from_numpy(w0).double().requires_grad_(True)
    x64 = torch.from_numpy(np.concatenate([ang, dih], 1)).double()
    lref = O.sigmoid_loss(2 * pi, SIG)(x64, torch.tanh(x64 @ wd)) * 500
    lref.backward()
    np.testing.assert_allclose(loss.item(), lref.item(), rtol=1e-5)
    assert _relnorm(gw.cpu().numpy(), wd.grad.numpy()) < 2e-5
    # cartesian distance loss at the ADC shape: (1024, 4950) pair distances, value and dL/dz
    n, d = 1024, 4950
    centres = rng.uniform(0.4, 8.0, size=(8, d))
    pair = (centres[rng.integers(0, 8, n)] + rng.normal(scale=4.5 / math.sqrt(2 * d), size=(n, d))).astype(np.float32)
    lat = (rng.normal(size=(n, 2)) * 1.5).astype(np.float32)
    fc = lf.cartesian_distance_loss(object(), ref_callers_tf._P(cartesian_distance_cost_scale=3.0))
    z = tf.Variable(lat)
    with tf.GradientTape() as tape:
        lc = fc(tf.convert_to_tensor(pair), z)
    gz = tape.gradient(lc, z)
    lref, gref = O.sigmoid_loss_and_grad(pair, lat, float("inf"), SIG)
    np.testing.assert_allclose(lc.item(), 3.0 * lref.item(), rtol=1e-5)
    assert _relnorm(gz.cpu().numpy(), 3.0 * gref.numpy()) < 1e-5
    # the reference's finite assertion still fires through the adapter
    bad = lat.copy()
    bad[7, 0] = np.nan
    with pytest.raises(FloatingPointError, match="infinite or NaN"):
        fc(tf.convert_to_tensor(pair), tf.convert_to_tensor(bad))


@pytest.mark.gpu
def test_tf_adapter_adc_branch_on_gpu(tf_gpu):
    """The Cartesian branch of the ADC model through the reference's layer classes (call rebound by install()):
    PeriodicInput -> dense -> BackMapLayer -> PairwiseDistances -> mean |difference| ; gradient w.r.t. the dense weights."""
    tf, adapter, mods = tf_gpu
    layers = mods["encodermap.models.layers"]
    rng = np.random.default_rng(9)
    n, b = 48, 20
    p = ref_callers_tf._P(cartesian_pwd_start=1, cartesian_pwd_step=3)
    dist = rng.uniform(0.13, 0.15, (b, n - 1)).astype(np.float32)
    ang = rng.uniform(1.9, 2.2, (b, n - 2)).astype(np.float32)
    dih = rng.uniform(-pi, pi, (b, n - 3)).astype(np.float32)
    target = np.abs(rng.normal(size=(b, 16 * 15 // 2))).astype(np.float32)
    wa0 = (rng.normal(size=(2 * (n - 3), n - 2)) * 0.02).astype(np.float32)
    wd0 = (rng.normal(size=(2 * (n - 3), n - 3)) * 0.5).astype(np.float32)
    wa, wd = tf.Variable(wa0), tf.Variable(wd0)
    with tf.GradientTape() as tape:
        feat = layers.PeriodicInput(p, "dihedrals")(tf.convert_to_tensor(dih))
        out_ang = tf.convert_to_tensor(ang) + feat @ wa
        out_dih = feat @ wd
        xyz = layers.BackMapLayer(n // 2 - 1, (n - 3) // 2)((tf.convert_to_tensor(dist), out_ang, out_dih))
        pw = layers.PairwiseDistances(p, "pairwise")(xyz)
        loss = tf.reduce_mean(tf.abs(pw - tf.convert_to_tensor(target)))
    ga, gd = tape.gradient(loss, [wa, wd])
    # float64 oracle of the same composition
    wa64, wd64 = torch.from_numpy(wa0).double().requires_grad_(True), torch.from_numpy(wd0).double().requires_grad_(True)
    feat64 = O.periodic_input(torch.from_numpy(dih).double(), 2 * pi)
    xyz64 = O.back_map_layer(torch.from_numpy(dist).double(), torch.from_numpy(ang).double() + feat64 @ wa64, feat64 @ wd64)
    pw64 = O.pairwise_distances_layer(xyz64, 1, None, 3)
    l64 = (pw64 - torch.from_numpy(target).double()).abs().mean()
    l64.backward()
    assert np.abs(xyz.detach().cpu().numpy() - xyz64.detach().numpy()).max() < 1e-4      # nm
    np.testing.assert_allclose(loss.item(), l64.item(), rtol=1e-5)
    assert _relnorm(ga.cpu().numpy(), wa64.grad.numpy()) < 5e-5
    assert _relnorm(gd.cpu().numpy(), wd64.grad.numpy()) < 5e-5
    # the same branch with PairwiseDistances + loss fused into one op (opt-in): same value, same gradients
    p.cartesian_cost_variant, p.cartesian_cost_scale, p.cartesian_cost_reference = "mean_abs", 1, 1
    fused = adapter.fused_cartesian_loss(None, None, p)
    with tf.GradientTape() as tape:
        feat = layers.PeriodicInput(p, "dihedrals")(tf.convert_to_tensor(dih))
        xyz = layers.BackMapLayer(n // 2 - 1, (n - 3) // 2)((tf.convert_to_tensor(dist), tf.convert_to_tensor(ang) + feat @ wa, feat @ wd))
        loss_f = fused(tf.convert_to_tensor(target), xyz)
    ga_f, gd_f = tape.gradient(loss_f, [wa, wd])
    np.testing.assert_allclose(loss_f.item(), l64.item(), rtol=1e-5)
    assert _relnorm(ga_f.cpu().numpy(), wa64.grad.numpy()) < 5e-5 and _relnorm(gd_f.cpu().numpy(), wd64.grad.numpy()) < 5e-5
    # cartesian_distance_loss fed with coordinates (8f-2) against the reference composition pair matrix -> loss
    z0 = rng.normal(size=(b, 2)).astype(np.float32)
    z = tf.Variable(z0)
    p.cartesian_dist_sig_parameters, p.cartesian_distance_cost_scale = (0.6, 6, 3, 1, 2, 6), 2.0
    with tf.GradientTape() as tape:
        lc = adapter.cartesian_distance_loss_from_coordinates(None, p)(xyz.detach(), z)
    gz = tape.gradient(lc, z)
    pairs64 = O.pairwise_distances_layer(xyz.detach().cpu().double(), 1, None, 3)
    lref, gref = O.sigmoid_loss_and_grad(pairs64.numpy(), z0, float("inf"), p.cartesian_dist_sig_parameters)
    np.testing.assert_allclose(lc.item(), 2.0 * lref.item(), rtol=1e-5)
    assert _relnorm(gz.cpu().numpy(), 2.0 * gref.numpy()) < 5e-5


@pytest.mark.gpu
def test_tf_adapter_sidechain_branch_on_gpu(tf_gpu):
    """The Cartesian branch with reconstructed side chains through the reference's layer classes (calls rebound by install()):
    BackMapLayerWithSidechains -> PairwiseDistances (gathered atoms) -> mean |difference|; gradients w.r.t. variables added to
    the angle and dihedral inputs.  The reference's own BackMapLayerWithSidechains.call is a placeholder that raises here."""
    tf, adapter, mods = tf_gpu
    layers = mods["encodermap.models.layers"]
    rng = np.random.default_rng(11)
    counts = [2, 0, 4, 1, 3, 0]
    fd = {-1: {k + 1: c for k, c in enumerate(counts)}}
    n_res, n_side, b = len(counts), sum(c + 1 for c in counts if c > 0), 7
    vals = [rng.uniform(0.13, 0.16, (b, 3 * n_res - 1)), rng.uniform(1.85, 2.25, (b, 3 * n_res - 2)), rng.uniform(-pi, pi, (b, 3 * n_res - 3)),
            rng.uniform(0.13, 0.19, (b, n_side)), rng.uniform(1.8, 2.2, (b, n_side)), rng.uniform(-pi, pi, (b, sum(counts)))]
    vals = [v.astype(np.float32) for v in vals]
    p = ref_callers_tf._P(cartesian_pwd_start=1, cartesian_pwd_step=3, reconstruct_sidechains=True, sidechain_info=fd)
    offs0 = [np.zeros((1, v.shape[1]), np.float32) for v in vals]
    offs = [tf.Variable(o) for o in offs0]
    pd_layer = layers.PairwiseDistances(p, "pairwise")
    n_sel = len(pd_layer.indices)
    target = np.abs(rng.normal(size=(b, n_sel * (n_sel - 1) // 2))).astype(np.float32)
    with tf.GradientTape() as tape:
        xyz = layers.BackMapLayerWithSidechains(fd)(tuple(tf.convert_to_tensor(v) + o for v, o in zip(vals, offs)))
        pw = pd_layer(xyz)
        loss = tf.reduce_mean(tf.abs(pw - tf.convert_to_tensor(target)))
    grads = tape.gradient(loss, offs)
    o64 = [torch.zeros(1, v.shape[1], dtype=torch.float64, requires_grad=True) for v in vals]
    xyz64 = O.backmap_with_sidechains(counts, [torch.from_numpy(v).double() + o for v, o in zip(vals, o64)])
    pw64 = O.pairwise_dist(xyz64[:, torch.as_tensor(O.sidechain_pairwise_indices(counts, 1, None, 3))], flat=True)
    l64 = (pw64 - torch.from_numpy(target).double()).abs().mean()
    l64.backward()
    assert np.abs(xyz.detach().cpu().numpy() - xyz64.detach().numpy()).max() < 1e-4      # nm
    np.testing.assert_allclose(loss.item(), l64.item(), rtol=1e-5)
    for g, o in zip(grads, o64):
        assert _relnorm(g.cpu().numpy(), o.grad.numpy()) < 5e-5


@pytest.mark.gpu
def test_tf_adapter_standalone_ops_on_gpu(tf_gpu):
    """every remaining callable of SURVEY.md 8b through the adapter: value and gradient against the float64 oracle"""
    tf, adapter, mods = tf_gpu
    d = mods["encodermap.misc.distances"]
    tf1 = mods["encodermap.encodermap_tf1.backmapping"]
    mb = mods["encodermap.misc.backmapping"]
    rng = np.random.default_rng(21)

    def check(fn_gpu, fn_ref, inputs, rtol=2e-5, atol=1e-6):
        xs = [tf.Variable(np.asarray(a, np.float32)) for a in inputs]
        with tf.GradientTape() as tape:
            out = fn_gpu(*xs)
        wgt = rng.normal(size=tuple(out.shape))
        grads = tape.gradient(tf.reduce_sum(out * tf.convert_to_tensor(wgt.astype(np.float32))), xs)
        xr = [torch.from_numpy(np.asarray(a, np.float32)).double().requires_grad_(True) for a in inputs]
        ref = fn_ref(*xr)
        (ref * torch.from_numpy(wgt)).sum().backward()
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=rtol, atol=atol)
        for g, r in zip(grads, xr):
            assert _relnorm(g.cpu().numpy(), r.grad.numpy()) < 5e-5

    x = rng.uniform(-pi, pi, (33, 17))
    check(lambda t: d.pairwise_dist_periodic(t, 2 * pi), lambda t: O.pairwise_dist_periodic(t, 2 * pi), [x])
    check(lambda t: d.pairwise_dist(t), lambda t: O.pairwise_dist(t), [rng.normal(size=(40, 3))])
    check(lambda t: d.pairwise_dist(t, flat=True), lambda t: O.pairwise_dist(t, flat=True), [rng.normal(size=(5, 17, 3))])
    check(lambda t: d.pairwise_dist(t, squared=True), lambda t: O.pairwise_dist(t, squared=True), [rng.normal(size=(3, 9, 3))])
    check(lambda s, t: d.periodic_distance(s, t, 2 * pi), lambda s, t: O.periodic_distance(s, t, 2 * pi),
          [rng.uniform(-pi, pi, (50, 7)), rng.uniform(-pi, pi, (50, 7))])
    check(lambda s, t: d.periodic_distance(s[:, None, :], t[None, :, :], 1.0), lambda s, t: O.periodic_distance(s[:, None, :], t[None, :, :], 1.0),
          [rng.uniform(0, 1, (6, 4)), rng.uniform(0, 1, (5, 4))])                      # broadcast form (distances.py:164-168)
    for params in ((4.5, 12, 6), (1, 2, 6), (0.2, 3, 6), (1.3, 2.5, 3.7)):
        check(lambda t: d.sigmoid(*params)(t), lambda t: O.sigmoid(*params)(t), [np.abs(rng.normal(size=64)) * 3 + 0.01], rtol=5e-5)
    n, b = 30, 4
    lengths = rng.uniform(0.13, 0.15, (1, n - 1))
    ang = rng.uniform(1.9, 2.2, (b, n - 2))
    dih = rng.uniform(-pi, pi, (b, n - 3))
    check(tf1.chain_in_plane, O.chain_in_plane, [lengths, ang], atol=1e-5)
    chain = O.chain_in_plane(torch.from_numpy(lengths), torch.from_numpy(ang)).numpy()
    check(tf1.dihedrals_to_cartesian_tf, O.dihedrals_to_cartesian_tf1, [dih, chain], atol=1e-5)
    check(lambda a_, c_: mb.dihedrals_to_cartesian_tf_layers(a_, c_, n // 2 - 1, (n - 3) // 2),
          lambda a_, c_: O.dihedrals_to_cartesian_layers(a_, c_, n // 2 - 1, (n - 3) // 2), [dih, chain], atol=1e-5)
    check(tf1.dihedral_to_cartesian_tf_one_way, O.dihedral_to_cartesian_one_way, [dih, chain], atol=1e-5)
    axis = rng.normal(size=(6, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    got = mb.rotation_matrix(tf.convert_to_tensor(axis.astype(np.float32)), tf.convert_to_tensor(rng.uniform(-pi, pi, 6).astype(np.float32)))
    assert tuple(got.shape) == (6, 3, 3)
    # generation side (misc/backmapping.py:1920-1990), called as reference tests/test_backmapping_em1_em2.py:571-591 does
    xyz = O.back_map_layer(torch.from_numpy(np.repeat(lengths, b, 0)), torch.from_numpy(ang), torch.from_numpy(dih)).numpy()
    n_idx, c_idx = np.arange(n)[::3], np.arange(n)[2::3]
    x_tf = tf.convert_to_tensor(xyz.astype(np.float32))
    h, o = mb.guess_amide_H(x_tf, n_idx), mb.guess_amide_O(x_tf, c_idx)
    merged = mb.merge_cartesians(x_tf, n_idx, c_idx, h, o)
    x32 = xyz.astype(np.float32).astype(np.float64)
    want = O.merge_cartesians(x32, n_idx, c_idx, O.guess_amide_H(x32, n_idx), O.guess_amide_O(x32, c_idx)).numpy()
    assert tuple(merged.shape) == want.shape
    assert np.abs(merged.cpu().numpy() - want).max() < 1e-5
