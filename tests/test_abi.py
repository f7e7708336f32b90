"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/emk.h declares, the
ctypes table covers them all, and the host-only index helpers are bit-exact against the oracle and
the golden vectors.  No compute entry point is called (there is no GPU here)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import em_oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def L():
    from encodermap_b200 import _build, _lib

    _build.build()
    return _lib


def declared_symbols():
    text = (ROOT / "include" / "emk.h").read_text()
    return sorted(set(re.findall(r"EMK_API\s+[\w\s\*]+?\b(emk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(L):
    lib = L.lib()
    names = declared_symbols()
    assert len(names) >= 45
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/emk.h but not exported by libemk.so"
    assert sorted(L.SIGNATURES) == names, "ctypes signature table and include/emk.h disagree"
    assert lib.emk_version() == 100
    assert lib.emk_build_info().startswith(b"sm_100a")


def test_library_is_self_contained(L):
    """No libcudart/libtorch dependency: static cudart, plain C ABI."""
    import subprocess

    out = subprocess.run(["ldd", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "libtorch" not in out and "libc10" not in out


def test_sass_is_sm100a_with_tma_and_packed_fp32(L):
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    r = subprocess.run(["cuobjdump", "-sass", "-fun", "pair_tile_kernel", str(L.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0 or "Function" not in r.stdout:
        r = subprocess.run(["cuobjdump", "-sass", str(L.LIB_PATH)], capture_output=True, text=True)
    sass = r.stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "sm_100" in sass
    assert "UTMALDG" in sass, "TMA tensor loads missing from the pair-tile kernel"
    assert "FFMA2" in sass and "FADD2" in sass, "packed FP32 instructions missing from the pair-tile kernel"


def test_triu_pair_order_bit_exact(L):
    for n in (0, 1, 2, 3, 7, 100):
        i, j = L.triu_pair_indices(n)
        wi, wj = np.triu_indices(n, k=1)
        assert np.array_equal(i, wi) and np.array_equal(j, wj)
    # the order the reference's boolean mask produces (distances.py:240-242), via the oracle
    x = np.random.default_rng(0).normal(size=(1, 9, 3))
    flat = O.pairwise_dist(x, flat=True).numpy()[0]
    full = O.pairwise_dist(x).numpy()[0]
    i, j = L.triu_pair_indices(9)
    assert np.array_equal(flat, full[i, j])


def test_split_indices_bit_exact(L, golden):
    g = golden["backmapping"]
    for n in (9, 12, 30, 31, 300):
        la, ra, dl, dr = L.backmap_split_indices(n)
        assert np.array_equal(la, g[f"n{n}_split_atoms_left"]) and np.array_equal(ra, g[f"n{n}_split_atoms_right"])
        assert np.array_equal(dl, g[f"n{n}_split_dih_left"]) and np.array_equal(dr, g[f"n{n}_split_dih_right"])
    # against the TF1 slicing for random chain lengths (reference tests/test_backmapping_em1_em2.py:2115-2156)
    rng = np.random.default_rng(3)
    for n in [int(v) for v in rng.integers(6, 1000, size=20)] + [4, 5, 6, 7, 1500]:
        la, ra, dl, dr = L.backmap_split_indices(n)
        wla, wdl, wra, wdr = O.split_indices_tf1(n)
        if n >= 6:
            assert np.array_equal(la, wla) and np.array_equal(ra, wra) and np.array_equal(dl, wdl) and np.array_equal(dr, wdr)
        assert (len(dl), len(dr)) == O.split_counts(n)
        cl, cr = O.split_and_reverse_cartesians(torch.arange(n)[None])
        tl, tr = O.split_and_reverse_dihedrals(torch.arange(n - 3)[None])
        assert np.array_equal(la, cl[0].numpy()) and np.array_equal(ra, cr[0].numpy())
        assert np.array_equal(dl, tl[0].numpy()) and np.array_equal(dr, tr[0].numpy())


def test_pair_tile_enumeration(L):
    for n in (1, 63, 64, 65, 128, 129, 200, 1000, 1100, 4096, 5000, 9000):
        tr, tc = -(-n // 128), -(-n // 64)
        want = {(i, j) for i in range(tr) for j in range(2 * i, tc)}
        assert L.pair_tile_count(n) == len(want)
        got = [L.pair_tile_decode(n, t) for t in range(len(want))]
        assert len(set(got)) == len(got) and set(got) == want      # a bijection onto the upper-triangular tiles
        # band-major order: 8 tile rows per band, column-major inside a band
        bands = [i // 8 for i, _ in got]
        assert bands == sorted(bands)
        for (i0, j0), (i1, j1) in zip(got, got[1:]):
            if i0 // 8 == i1 // 8:
                assert (j1, i1) > (j0, i0)
        # every unordered pair (i <= j) of rows is covered exactly once by the tile list
        if n <= 200:
            cover = np.zeros((n, n), dtype=int)
            for (i, j) in want:
                r0, r1, c0, c1 = i * 128, min(n, i * 128 + 128), j * 64, min(n, j * 64 + 64)
                cover[r0:r1, c0:c1] += 1
                if j // 2 != i:
                    cover[c0:c1, r0:r1] += 1
            assert (cover == 1).all()
    assert L.pair_tile_count(65536) == 262656
    seen = {L.pair_tile_decode(65536, t) for t in range(0, 262656, 97)} | {L.pair_tile_decode(65536, 262655)}
    assert all(0 <= i < 512 and 2 * i <= j < 1024 for i, j in seen) and len(seen) == len(range(0, 262656, 97)) + 1
    for world in (1, 2, 3, 4, 8):
        ranges = [L.pair_tile_range(65536, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 262656
        assert all(ranges[k][1] == ranges[k + 1][0] for k in range(world - 1))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_host_helpers_report_errors(L):
    lib = L.lib()
    b, e = ctypes.c_int64(), ctypes.c_int64()
    assert lib.emk_pair_tile_range(100, 3, 2, ctypes.byref(b), ctypes.byref(e)) == -6
    assert b"rank" in lib.emk_last_error()
    assert lib.emk_pair_tile_decode(100, 10 ** 9, ctypes.byref(b), ctypes.byref(e)) == -6
    assert lib.emk_backmap_split_counts(2, (ctypes.c_int64 * 4)()) == -4
    with pytest.raises(L.EmkError):
        L.pair_tile_range(100, 5, 2)


def test_product_path_refuses_cpu_tensors(L):
    """No CPU fallback: the public API raises on CPU tensors instead of computing something else."""
    import encodermap_b200 as em
    from encodermap_b200.loss_functions import sigmoid_loss
    from encodermap_b200.misc.distances import pairwise_dist_periodic
    from encodermap_b200.models.layers import back_map

    with pytest.raises(em.EmkError):
        sigmoid_loss()(torch.zeros(4, 3), torch.zeros(4, 2))
    with pytest.raises(em.EmkError):
        back_map(torch.zeros(2, 5), torch.zeros(2, 4), torch.zeros(2, 3))
    if not torch.cuda.is_available():
        with pytest.raises((em.EmkError, RuntimeError, AssertionError)):
            pairwise_dist_periodic(np.zeros((4, 3), np.float32), 1.0)


def test_product_does_not_import_oracle():
    """Nothing under encodermap_b200/ may import, call or link the oracle."""
    pat = re.compile(r"^\s*(from\s+\.*oracle|import\s+oracle|from\s+\S*em_oracle|import\s+\S*em_oracle)|em_oracle|oracle/", re.M)
    for path in (ROOT / "encodermap_b200").rglob("*.py"):
        assert not pat.search(path.read_text()), f"{path} reaches into oracle/"
    for path in (ROOT / "encodermap_b200" / "csrc").glob("*"):
        assert "#include \"../../oracle" not in path.read_text() and "em_oracle" not in path.read_text()


def test_row_chunk_tile_ranges_for_streamed_input(L):
    """Host logic of sigmoid_cost_streamed: tile ids are band-major, so the tiles from chunk c's tile_begin on touch no row
    (and no column) before the chunk's first row -- checked exhaustively against emk_pair_tile_decode."""
    from encodermap_b200 import _ops

    for n, rows in ((300, 1024), (2500, 1024), (4097, 2048), (9000, 3072)):
        chunks = _ops._row_chunk_tiles(n, rows)
        total = L.pair_tile_count(n)
        assert chunks[0] == (0, 0)
        for (r0, t0), nxt in zip(chunks, chunks[1:] + [(n, total)]):
            assert t0 <= nxt[1]
            for t in range(t0, nxt[1]):
                i, j = L.pair_tile_decode(n, t)
                assert i * 128 >= r0 and j * 64 >= r0
            if t0 > 0:
                i, _ = L.pair_tile_decode(n, t0 - 1)
                assert i * 128 < r0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints exactly one JSON line with the keys
    of the measurement contract; it runs here because it only times the oracle port on the host cores."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sigmoid_cost_pairs_per_s_fwd_bwd" and d["unit"] == "unique pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_streamed_cost_control_flow(L, monkeypatch):
    """Host logic of sigmoid_cost_streamed with the CUDA pieces stubbed out (no GPU here): the launches walk the tile list
    from the last row chunk to the first, the ranges chain without gaps, and only the first launch zeroes the outputs."""
    from encodermap_b200 import _ops

    calls = []

    class FakeStream:
        def wait_stream(self, s): pass
        def wait_event(self, e): pass

    class FakeEvent:
        def record(self, s): pass

    class Ctx:
        def __init__(self, *a): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False

    real = L.lib()

    class FakeLib:
        def __getattr__(self, k):
            return getattr(real, k)

        def emk_dl_sigmoid_cost(self, high, low, per, sig, t0, t1, loss, grad, flags, st):
            calls.append((t0, t1, flags))
            return 0

    fake = FakeLib()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda d=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", Ctx)
    monkeypatch.setattr(torch.cuda, "device", Ctx)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setattr(_ops, "require_cuda", lambda t, n: t)
    monkeypatch.setattr(_ops, "stream_of", lambda t: None)
    monkeypatch.setattr(_ops, "DL", lambda t: t)
    monkeypatch.setattr(L, "lib", lambda: fake)
    monkeypatch.setattr(_ops, "_SIDE_STREAMS", {})
    n = 5000
    _ops.sigmoid_cost_streamed(torch.zeros(n, 8), torch.zeros(n, 2), 6.28, (4.5, 12, 6, 1, 2, 6), True, 1024)
    total = int(real.emk_pair_tile_count(n))
    assert calls[0][1] == total and calls[-1][0] == 0 and len(calls) == 5
    assert all(a[0] == b[1] for a, b in zip(calls, calls[1:]))
    assert calls[0][2] & L.EMK_COST_ZERO_OUTPUTS and not any(c[2] & L.EMK_COST_ZERO_OUTPUTS for c in calls[1:])


def test_streamed_cost_interleaved_split(L, monkeypatch):
    """More than two ranks with the input in pinned host memory: every rank takes 1/G of the tiles of EVERY row chunk (walked from
    the last chunk to the first), copies 1/G of the chunk's rows and completes the chunk by an all-gather.  Host logic only
    (CUDA and torch.distributed stubbed): the ranks' launches tile every chunk's range without gaps or overlap."""
    import torch.distributed as dist

    from encodermap_b200 import _ops

    class FakeStream:
        def wait_stream(self, s): pass
        def wait_event(self, e): pass

    class FakeEvent:
        def record(self, s): pass

    class Ctx:
        def __init__(self, *a): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False

    real = L.lib()
    calls, gathers = [], []

    class FakeLib:
        def __getattr__(self, k):
            return getattr(real, k)

        def emk_dl_sigmoid_cost(self, high, low, per, sig, t0, t1, loss, grad, flags, st):
            calls.append((t0, t1, flags))
            return 0

    state = {"rank": 0}
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda d=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", Ctx)
    monkeypatch.setattr(torch.cuda, "device", Ctx)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setattr(_ops, "require_cuda", lambda t, n: t)
    monkeypatch.setattr(_ops, "stream_of", lambda t: None)
    monkeypatch.setattr(_ops, "DL", lambda t: t)
    monkeypatch.setattr(L, "lib", lambda: FakeLib())
    monkeypatch.setattr(_ops, "_SIDE_STREAMS", {})
    monkeypatch.setattr(dist, "get_world_size", lambda g=None: 4)
    monkeypatch.setattr(dist, "get_rank", lambda g=None: state["rank"])
    monkeypatch.setattr(dist, "all_gather_into_tensor", lambda out, inp, group=None: gathers.append((tuple(out.shape), tuple(inp.shape))))
    n, rows = 8192, 2048
    total = int(real.emk_pair_tile_count(n))
    per_rank = []
    for rank in range(4):
        state["rank"] = rank
        calls.clear()
        gathers.clear()
        _ops.sigmoid_cost_streamed(torch.zeros(n, 8), torch.zeros(n, 2), 6.28, (4.5, 12, 6, 1, 2, 6), True, rows, interleave_group="g")
        assert gathers == [((rows, 8), (rows // 4, 8))] * (n // rows)         # one all-gather per chunk, equal slices
        assert calls[0][2] & L.EMK_COST_ZERO_OUTPUTS and not any(c[2] & L.EMK_COST_ZERO_OUTPUTS for c in calls[1:])
        per_rank.append([(a, b) for a, b, _ in calls])
    assert all(len(c) == n // rows for c in per_rank)
    covered = 0
    for k in range(n // rows):                       # chunk k (from the last one backwards): ranks 0..3 chain
        parts = [per_rank[r][k] for r in range(4)]
        assert all(parts[r][1] == parts[r + 1][0] for r in range(3))
        covered += parts[3][1] - parts[0][0]
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    assert covered == total and per_rank[3][0][1] == total and per_rank[0][-1][0] == 0
    # ragged shapes fall back to contiguous tile ranges (every rank streams the rows its own range touches)
    state["rank"] = 1
    calls.clear()
    gathers.clear()
    _ops.sigmoid_cost_streamed(torch.zeros(5000, 8), torch.zeros(5000, 2), 6.28, (4.5, 12, 6, 1, 2, 6), True, 1024, interleave_group="g")
    b, e = L.pair_tile_range(5000, 1, 4)
    assert not gathers and min(a for a, _, _ in calls) == b and max(t for _, t, _ in calls) == e


def test_merged_atom_count_follows_the_reference_loop(L):
    """emk_merged_atom_count (host only): atom 0, then every atom i >= 1 followed by a hydrogen if i is in h_after, else by an
    oxygen if i is in o_after -- the loop of reference misc/backmapping.py:1970-1990 -- incl. its quirks: atom 0 never gets a
    partner, an atom in both lists gets the hydrogen only."""
    import ctypes

    import numpy as np

    def count(n, h_after, o_after):
        h = np.asarray(h_after, dtype=np.int64)
        o = np.asarray(o_after, dtype=np.int64)
        return int(L.lib().emk_merged_atom_count(n, h.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), h.size,
                                           o.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), o.size))

    def reference_loop(n, h_after, o_after):
        total = 1
        for i in range(1, n):
            total += 1
            if i in h_after:
                total += 1
            elif i in o_after:
                total += 1
        return total

    rng = np.random.default_rng(4)
    cases = [(9, list(range(0, 9, 3))[1:], list(range(2, 9, 3))), (30, [], []), (30, [0], [0]), (12, [3, 6, 9], [3, 5, 8, 11]), (1, [], [])]
    for _ in range(20):
        n = int(rng.integers(3, 60))
        cases.append((n, rng.choice(n, size=int(rng.integers(0, n)), replace=False).tolist(),
                      rng.choice(n, size=int(rng.integers(0, n)), replace=False).tolist()))
    for n, h, o in cases:
        assert count(n, h, o) == reference_loop(n, h, o), (n, h, o)
    assert count(10, [10], []) == -1 and count(10, [], [-1]) == -1      # outside the backbone


def test_near_and_far_sides_against_networkx_on_random_trees():
    """Product host logic (breadth-first far sides) against the reference's method -- remove the edge from a networkx graph, take
    the two connected components (misc/rotate.py:444-505) -- on random trees; membership must be identical."""
    nx = pytest.importorskip("networkx")
    from encodermap_b200.misc.backmapping import near_and_far_sides

    rng = np.random.default_rng(12)
    for _ in range(25):
        n = int(rng.integers(2, 80))
        bonds = [(int(rng.integers(0, k)), k) for k in range(1, n)]          # a random tree
        edges = [bonds[k] if rng.random() < 0.5 else bonds[k][::-1] for k in rng.choice(len(bonds), size=min(10, len(bonds)), replace=False)]
        near, far = near_and_far_sides(n, bonds, edges)
        for (u, v), nr, fr in zip(edges, near, far):
            g = nx.Graph()
            g.add_nodes_from(range(n))
            g.add_edges_from(bonds)
            g.remove_edge(u, v)
            comps = list(nx.connected_components(g))
            assert len(comps) == 2
            comp_u = next(c for c in comps if u in c)
            comp_v = next(c for c in comps if v in c)
            assert set(nr.tolist()) == comp_u and set(fr.tolist()) == comp_v
            assert np.all(np.diff(nr) > 0) and np.all(np.diff(fr) > 0)      # sorted, unique


def test_sidechain_plan_equals_the_reference_masks(L, golden):
    """emk_sidechain_plan_create against the boolean masks / index tables the reference's constructor builds
    (models/layers.py:234-474, run by tools/gen_golden.py): bit-exact, as ranges."""
    from encodermap_b200 import _ops

    g = golden["sidechains"]
    for tag in ("metlysgly", "first_empty", "twelve", "ub_like"):
        counts = g[f"{tag}_counts"]
        plan = _ops.SidechainPlan(counts)
        ops = plan.ops()
        masks = np.ones((plan.n_ops, plan.n_atoms), dtype=bool)
        for k, o in enumerate(ops):
            assert 0 <= o[6] <= o[7] <= plan.n_atoms and 0 <= o[8] <= o[9] <= plan.n_atoms
            masks[k, o[6]:o[7]] = False
            masks[k, o[8]:o[9]] = False
        want = np.vstack([g[f"{tag}_central_angle_mask"], g[f"{tag}_side_angle_mask"], g[f"{tag}_dihedral_mask"]])
        assert np.array_equal(masks, want), tag
        n_ang = len(g[f"{tag}_central_angle_mask"]) + len(g[f"{tag}_side_angle_mask"])
        assert np.array_equal(ops[:n_ang, 1:4], np.vstack([g[f"{tag}_central_angle_triplets"], g[f"{tag}_side_angle_triplets"]]))
        assert (ops[:n_ang, 4] == -1).all()
        assert np.array_equal(ops[n_ang:, 1:5], g[f"{tag}_dihedral_quadruplets"])
        n_ca, n_sa, n_cd = len(g[f"{tag}_central_angle_mask"]), len(g[f"{tag}_side_angle_mask"]), 3 * len(counts) - 3
        assert np.array_equal(ops[:, 0], np.concatenate([np.zeros(n_ca), np.ones(n_sa), np.full(n_cd, 2), np.full(plan.n_ops - n_ang - n_cd, 3)]))
        for kind in range(4):    # every input column is used exactly once, in order
            cols = ops[ops[:, 0] == kind, 5]
            assert np.array_equal(cols, np.arange(len(cols)))
        assert plan.columns == tuple(g[f"{tag}_in_{k}"].shape[1] for k in ("cd", "ca", "cdih", "sd", "sa", "sdih"))
        assert plan.n_atoms == g[f"{tag}_out"].shape[1]
        for sel, (a, b, c) in {"ca": (1, None, 3), "all": (None, None, None)}.items():
            assert np.array_equal(_ops.sidechain_pairwise_indices(counts, a, b, c), g[f"{tag}_pwd_indices_{sel}"])


def test_sidechain_plan_refuses_what_the_reference_cannot_build(L):
    from encodermap_b200 import _ops
    from encodermap_b200.models.layers import BackMapLayerWithSidechains

    for bad in ([3, 4, 2], [0, 2, 0], [0, 0, 3, 1], [0, 0, 0]):
        with pytest.raises(ValueError):
            _ops.SidechainPlan(bad)
    with pytest.raises(AssertionError):
        BackMapLayerWithSidechains({-1: {1: 2, 3: 0}})
    layer = BackMapLayerWithSidechains({-1: {1: 3, 2: 4, 3: 0}})
    assert layer.n_atoms == 18 and layer.n_sidechains == 9
    import copy
    import pickle

    for clone in (copy.deepcopy(layer), pickle.loads(pickle.dumps(layer))):     # library handles are rebuilt, not copied
        assert clone.counts == layer.counts and clone._plans == {} and clone._plan_for(None).n_atoms == 18
    again = BackMapLayerWithSidechains.from_config({"feature_description": {"-1": {"1": 3, "2": 4, "3": 0}}})
    assert again.counts == layer.counts and again.get_config()["feature_description"] == {-1: {1: 3, 2: 4, 3: 0}}


def test_sidechain_plan_equals_the_oracle_topology_on_random_descriptions(L):
    """300 random admissible residue descriptions (either end bare, interior residues with 0..6 side-chain dihedrals): the step
    table of the library, expanded to masks, equals the oracle's restatement of the reference constructor (itself pinned to the
    reference's own constructor on four descriptions, tests/test_oracle_golden.py) bit for bit."""
    from oracle import em_oracle as O

    from encodermap_b200 import _ops

    rng = np.random.default_rng(77)
    for trial in range(300):
        n_res = int(rng.integers(2, 40))
        counts = [int(v) for v in rng.integers(0, 7, size=n_res)]
        if trial % 2:
            counts[0], counts[-1] = 0, max(1, counts[-1])
            if n_res > 2:
                counts[1] = max(1, counts[1])        # the reference needs a side chain before the first interior bare residue
        else:
            counts[0], counts[-1] = max(1, counts[0]), 0
        topo = O.sidechain_topology(counts)
        plan = _ops.SidechainPlan(counts)
        ops = plan.ops()
        masks = np.ones((plan.n_ops, plan.n_atoms), dtype=bool)
        for k, o in enumerate(ops):
            masks[k, o[6]:o[7]] = False
            masks[k, o[8]:o[9]] = False
        want = np.vstack([topo["central_angle_mask"], topo["side_angle_mask"], topo["dihedral_mask"]])
        assert np.array_equal(masks, want), counts
        n_ang = len(topo["central_angle_mask"]) + len(topo["side_angle_mask"])
        assert np.array_equal(ops[:n_ang, 1:4], np.vstack([topo["central_angle_triplets"], topo["side_angle_triplets"]])), counts
        assert np.array_equal(ops[n_ang:, 1:5], topo["dihedral_quadruplets"]), counts
        assert np.array_equal(_ops.sidechain_pairwise_indices(counts, 1, None, 3), O.sidechain_pairwise_indices(counts, 1, None, 3))
