"""CPU-only tests: C-ABI surface, host logic (HVG selection, sharding, collectives over gloo), generator."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    """libscarf_b200.so loads without a GPU and exports exactly what include/scarf_b200.h declares."""
    from scarf_b200 import lib

    header = open(os.path.join(ROOT, "include", "scarf_b200.h")).read()
    declared = set(re.findall(r"\b(scf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), name
    assert lib.version() == 100


def test_abi_argument_errors_do_not_need_a_gpu():
    from scarf_b200 import lib

    with pytest.raises(ValueError, match="null pointer"):
        lib.call("scf_csr_row_sums", None, None, None, None, 0, None, None, None, None)
    with pytest.raises(ValueError, match="method"):
        lib.call("scf_knn_l2", 8, 1, 8, 10, 4, 4, 2, -1, 8, 8, 7, None, 0, None)
    assert b"method" in lib.raw("scf_last_error")()


def _knn_plan(nq, nref, dim, k):
    from scarf_b200 import lib

    out = (ctypes.c_int32 * 16)()
    rc = lib.raw("scf_knn_plan")(nq, nref, dim, k, ctypes.cast(out, ctypes.c_void_p))
    names = ("kc", "kchunks", "pair", "units", "rounds", "units_rem", "nsplit", "nlists", "groups", "ref_tiles", "q_tiles",
             "stages", "issuers", "smem_kib", "max_cands")
    return rc, dict(zip(names, list(out)))


@pytest.mark.parametrize("shape", [(100_000, 100_000, 50, 11), (125_000, 1_000_000, 100, 21), (1_000_000, 1_000_000, 100, 21),
                                   (500_000, 4_000_000, 100, 11), (300, 500_000, 40, 11), (1, 10, 5, 3), (7, 40_001, 130, 24),
                                   (40_000, 400_000, 100, 21), (5000, 5000, 189, 24)])
def test_knn_plan_invariants(shape, monkeypatch):
    """scf_knn_plan (host only): the schedule the tensor-core kNN would run.  Rounds cover whole query tiles and only
    when the reference operand exceeds L2; the lists a row can end up with fit the re-rank kernel; the pipeline is deeper
    than a step; the CTA-pair kernel (opt-in) keeps four lists of 16 per row."""
    nq, nref, dim, k = shape
    for pair in ("0", "1"):
        monkeypatch.setenv("SCF_KNN_PAIR", pair)
        rc, p = _knn_plan(nq, nref, dim, k)
        assert rc == 0, shape
        assert p["pair"] == int(pair)
        units_max = 74 if p["pair"] else 148
        assert 1 <= p["units"] <= units_max and 1 <= p["units_rem"] <= p["units"]
        assert p["rounds"] * p["units"] <= p["q_tiles"] and (p["q_tiles"] - p["rounds"] * p["units"]) < max(p["units"], p["q_tiles"] + 1)
        assert p["nlists"] == p["nsplit"] * p["groups"] and p["nlists"] * p["kc"] <= p["max_cands"]
        assert p["groups"] == (4 if p["pair"] else 2) and p["kc"] in (16, 32) and (not p["pair"] or p["kc"] == 16)
        assert p["kchunks"] == (dim + 3 + 63) // 64 and p["stages"] >= p["kchunks"] + 1 and p["smem_kib"] <= 227
        assert p["issuers"] == (2 if p["kchunks"] >= 2 else 1)
        ref_step = 256 if p["pair"] else 128
        assert p["ref_tiles"] == -(-nref // ref_step) and p["q_tiles"] == -(-nq // 256)
        fits_l2 = p["ref_tiles"] * ref_step * p["kchunks"] * 64 * 2 <= 48 << 20
        assert (p["rounds"] == 0) if fits_l2 else (p["rounds"] == p["q_tiles"] // p["units"])


def test_knn_plan_rejects_shapes_of_the_fp64_kernel():
    assert _knn_plan(1000, 1000, 190, 11)[0] == 1   # dim + 3 > 192
    assert _knn_plan(1000, 1000, 20, 25)[0] == 1    # k > 24


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: importing the binding without the built library is an ImportError."""
    code = ("import importlib.util, sys; spec = importlib.util.spec_from_file_location('lib', sys.argv[1]);"
            "m = importlib.util.module_from_spec(spec); m.__file__ = sys.argv[2]; spec.loader.exec_module(m)")
    fake = tmp_path / "scarf_b200"
    fake.mkdir()
    src = open(os.path.join(ROOT, "scarf_b200", "lib.py")).read()
    (fake / "lib.py").write_text(src)
    r = subprocess.run([sys.executable, "-c", code, str(fake / "lib.py"), str(fake / "lib.py")], capture_output=True,
                       text=True)
    assert r.returncode != 0 and "ImportError" in r.stderr and "no CPU fallback" in r.stderr


def test_hvg_host_logic_matches_oracle(pbmc):
    """Product LOWESS / HVG choice (scarf_b200/hvg.py) vs the oracle restatement on the PBMC statistics."""
    from oracle import pipeline as P
    from scarf_b200 import hvg

    counts, cell_idx = pbmc["counts"], pbmc["cell_idx"]
    feat_I = P.gene_ncells(counts) > 20
    n_counts, _ = P.cell_totals(counts)
    hv_o, st = P.mark_hvgs(counts, cell_idx, feat_I, gene_names=pbmc["names"], top_n=100, return_stats=True)
    c_var = np.full(counts.shape[1], np.nan)
    c_var[feat_I] = hvg.remove_trend(st["avg"][feat_I], st["sigmas"][feat_I])
    np.testing.assert_allclose(c_var[feat_I], st["c_var"][feat_I], rtol=1e-10)
    hv = hvg.choose_hvgs(st["normed_n"], st["nz_mean"], c_var, feat_I, pbmc["names"], 100, int(0.01 * 892))
    assert np.array_equal(hv, hv_o)
    # blacklist really removes genes (the reference's default regex)
    assert not any(str(n).startswith(("MT-", "RPS", "RPL")) for n in pbmc["names"][hv])


def test_hvg_device_functions_match_host(pbmc):
    """remove_trend_device / choose_hvgs_device (torch, here on CPU tensors) == the numpy statements."""
    import torch

    from oracle import pipeline as P
    from scarf_b200 import hvg

    counts, cell_idx = pbmc["counts"], pbmc["cell_idx"]
    feat_I = P.gene_ncells(counts) > 20
    hv_o, st = P.mark_hvgs(counts, cell_idx, feat_I, gene_names=pbmc["names"], top_n=100, return_stats=True)
    t = {k: torch.from_numpy(np.asarray(v, dtype=np.float64)) for k, v in st.items()}
    fI = torch.from_numpy(feat_I)
    c_var = hvg.remove_trend_device(t["avg"], t["sigmas"], select=fI)
    c_var = torch.where(fI, c_var, torch.full_like(c_var, float("nan")))
    ref = hvg.remove_trend(st["avg"][feat_I], st["sigmas"][feat_I])
    np.testing.assert_allclose(c_var[fI].numpy(), ref, rtol=1e-12)
    keep = torch.from_numpy(hvg.blacklist_keep_mask(pbmc["names"], counts.shape[1]))
    hv = hvg.choose_hvgs_device(t["normed_n"], t["nz_mean"], c_var, fI & keep, 100, int(0.01 * 892))
    assert np.array_equal(hv.numpy(), hv_o)


def test_lowess_ties_and_small_windows():
    from oracle.lowess import lowess as lowess_o
    from scarf_b200.hvg import _lowess, _lowess_numpy

    rng = np.random.default_rng(5)
    x = np.sort(rng.normal(size=60))
    x[10] = x[9]
    x[30:33] = x[30]
    y = np.sin(x) + 0.1 * rng.normal(size=60)
    y[20] += 3.0  # outlier exercises the robustness iterations
    np.testing.assert_allclose(_lowess(y, x, 0.2, 100), lowess_o(y, x, frac=0.2, it=100), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(_lowess_numpy(y, x, 0.2, 100), lowess_o(y, x, frac=0.2, it=100), rtol=1e-9, atol=1e-12)
    x2 = rng.gamma(2.0, 1.0, size=200)  # unsorted input, even n, the size mark_hvgs uses
    y2 = np.log1p(x2) + 0.05 * rng.normal(size=200)
    np.testing.assert_allclose(_lowess(y2, x2, 0.1, 100), lowess_o(y2, x2, frac=0.1, it=100), rtol=1e-9, atol=1e-12)


def test_shard_plan():
    from scarf_b200.dist import ShardPlan

    p = ShardPlan.make(100_000, 8, 1000)
    assert p.starts[0] == 0 and p.stops[-1] == 100_000
    assert all(a == b for a, b in zip(p.stops[:-1], p.starts[1:]))
    assert all(s % 1000 == 0 for s in p.starts)
    assert max(b - a for a, b in zip(p.starts, p.stops)) - min(b - a for a, b in zip(p.starts, p.stops)) <= 1000
    p = ShardPlan.make(2500, 4, 1000)  # fewer chunks than ranks: trailing ranks are empty
    assert [b - a for a, b in zip(p.starts, p.stops)] == [1000, 1000, 500, 0]


def test_synth_shards_are_slices_of_the_global_matrix():
    from scarf_b200 import synth

    full = synth.make_counts_scipy(900, 500, 8, seed=3, block=300)
    part = synth.make_counts_scipy(300, 500, 8, seed=3, block=300, row_start=600)
    assert (full[600:] != part).nnz == 0
    assert full.indices.dtype == np.int32 and np.all(np.diff(full.indptr) > 0)
    assert full.has_sorted_indices


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as td
sys.path.insert(0, sys.argv[1])
from scarf_b200.dist import Comm, ShardPlan
td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
c = Comm()
plan = ShardPlan.make(2500, 2, 1000)
a, b = plan.rows(c.rank)
rows = torch.arange(a, b, dtype=torch.float32)[:, None] * torch.ones(1, 3)
counts = c.allgather_counts(b - a, "cpu")
assert counts == [2000, 500], counts
allr = c.allgather_rows(rows, counts)
assert allr.shape == (2500, 3) and torch.equal(allr[:, 0], torch.arange(2500, dtype=torch.float32))
g = torch.full((4,), 2 ** 40 + c.rank, dtype=torch.int64)
c.allreduce_sum_(g)
assert int(g[0]) == 2 ** 41 + 1            # int64 fixed-point sums are exact
m = torch.tensor([1.0 + c.rank, 5.0 - c.rank]); c.allreduce_min_(m); assert m.tolist() == [1.0, 4.0]
z = torch.tensor([c.rank], dtype=torch.int32); c.allreduce_max_(z); assert int(z) == 1
# sharded write of the graph arrays (scarf/knn_utils.py:54-59,108-117): rank 0 creates, every rank writes its own rows
import numpy as np
from types import SimpleNamespace
from scarf_b200.datastore import write_graph_arrays
from scarf_b200.zarr_store import open_group
k, path = 3, sys.argv[4]
if c.rank == 0:
    open_group(path, "w")
c.barrier()
gi = torch.arange(a, b, dtype=torch.int64)
res = SimpleNamespace(n_cells=2500, row_offset=a, indices=gi[:, None] * 10 + torch.arange(k),
                      distances=(gi[:, None] + torch.arange(k) / 4.0).to(torch.float32),
                      edges=torch.stack([gi.repeat_interleave(k), (gi[:, None] * 10 + torch.arange(k)).flatten()], 1),
                      weights=(gi.repeat_interleave(k) / 2.0).to(torch.float32))
nb = write_graph_arrays(open_group(path, "r+"), "A/knn__3", "A/knn__3/graph__1.0__1.5", res, c, batch_size=1000)
assert nb == (b - a) * k * (8 + 8 + 16 + 8), nb
if c.rank == 0:
    z = open_group(path, "r")
    full = np.arange(2500)
    assert z["A/knn__3"]["indices"].dtype == np.dtype("u8") and z["A/knn__3"]["indices"].chunks == (1000, 3)
    assert np.array_equal(z["A/knn__3"]["indices"][:], full[:, None] * 10 + np.arange(k))
    assert np.array_equal(z["A/knn__3"]["distances"][:], full[:, None] + np.arange(k) / 4.0)
    e = z["A/knn__3/graph__1.0__1.5"]["edges"]
    assert e.shape == (7500, 2) and e.chunks == (3000, 2) and np.array_equal(e[:][:, 0], np.repeat(full, k))
    assert np.array_equal(z["A/knn__3/graph__1.0__1.5"]["weights"][:], np.repeat(full, k) / 2.0)
c.barrier(); td.destroy_process_group(); print("ok")
'''


def test_comm_world2_gloo(tmp_path):
    """The N>1 host logic (uneven all-gather, exact int64 all-reduce, min/max floor exchange, the sharded write of the
    graph arrays) on CPU over gloo."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(tmp_path / "g.zarr")],
                              stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err


def test_subset_hash_and_shard_spans():
    """The subset hash is the reference's (scarf/assay.py:317-329: Python's hash of the two index tuples, an int), and
    DataStore._shard cuts the selected cells into chunk-aligned blocks with the raw-row span that holds each."""
    from types import SimpleNamespace

    from scarf_b200.datastore import DataStore, create_subset_hash

    ci, fi = np.array([0, 2, 5, 9], dtype=np.int64), np.array([1, 3], dtype=np.int64)
    assert create_subset_hash(ci, fi) == hash(tuple([hash(tuple(ci)), hash(tuple(fi))]))
    assert isinstance(create_subset_hash(ci, fi), int) and create_subset_hash(ci, fi) != create_subset_hash(ci[:3], fi)
    cells = np.arange(0, 7000, 2)  # 3500 selected cells out of 7000 raw rows
    spans = [DataStore._shard(SimpleNamespace(_world=3, _rank=r, cells=SimpleNamespace(N=7000)), cells, 1000)
             for r in range(3)]
    assert [(s, e) for s, e, _, _ in spans] == [(0, 2000), (2000, 3000), (3000, 3500)]
    assert spans[1][2:] == (4000, 5999) and spans[2][2:] == (6000, 6999)
    assert DataStore._shard(SimpleNamespace(_world=1, _rank=0, cells=SimpleNamespace(N=7000)), cells, 1000) == \
        (0, 3500, 0, 7000)


def test_csr_sortedness_check_on_cpu_tensors():
    """CsrDevice.check_sorted (device-agnostic torch code): unsorted or duplicated column ids inside a row are refused,
    empty rows and row boundaries are not mistaken for a descent."""
    import torch

    from scarf_b200.ops import CsrDevice

    ip = torch.tensor([0, 3, 3, 5, 6], dtype=torch.int64)
    ok = CsrDevice(ip, torch.tensor([1, 4, 9, 0, 2, 0], dtype=torch.int32), torch.ones(6, dtype=torch.int32), 4, 10)
    ok.check_sorted()
    for bad_ix in ([1, 9, 4, 0, 2, 0], [1, 4, 4, 0, 2, 0]):
        bad = CsrDevice(ip, torch.tensor(bad_ix, dtype=torch.int32), torch.ones(6, dtype=torch.int32), 4, 10)
        with pytest.raises(ValueError, match="ascend strictly"):
            bad.check_sorted()


def test_graph_parameter_resolution_equals_reference_method():
    """a12: GraphDataStore._set_graph_params (scarf/datastore/graph_datastore.py:63-363) was executed on stub stores
    (tests/golden/make_ref_function_goldens.py -> ref_graph_params.json): fresh store, explicit values, values cached by
    an earlier run at every level of the tree, partial overrides, and the error cases as they really fall out (the
    pca_cell_key column is only checked when `dims` is not given).  DataStore._set_graph_params resolves the same tuples
    on the same stubs."""
    import json
    import os
    from types import SimpleNamespace

    from conftest import GOLDEN
    from scarf_b200.datastore import DataStore

    class Tree(dict):
        pass

    cells = SimpleNamespace(columns=["I", "ids", "names", "sub", "RNA_nCounts"],
                            get_dtype=lambda c: bool if c in ("I", "sub") else float)
    with open(os.path.join(GOLDEN, "ref_graph_params.json")) as f:
        scenarios = json.load(f)
    assert len(scenarios) == 13
    for sc in scenarios:
        stub = SimpleNamespace(zw=Tree({k: SimpleNamespace(attrs=dict(v)) for k, v in sc["tree"].items()}), cells=cells)
        if "raises" in sc:
            with pytest.raises({"ValueError": ValueError, "TypeError": TypeError}[sc["raises"]]):
                DataStore._set_graph_params(stub, "RNA", "I", "hvgs", **sc["kwargs"])
        else:
            got = DataStore._set_graph_params(stub, "RNA", "I", "hvgs", **sc["kwargs"])
            assert list(got) == sc["result"], (sc["kwargs"], got, sc["result"])
            assert [type(v) for v in got] == [type(v) for v in sc["result"]]


def test_public_signatures_equal_the_reference():
    """(b): parameter names, order and defaults of the four drop-in entry points, as recorded from the reference's
    source (scarf/datastore/{datastore,graph_datastore,mapping_datastore}.py, parsed with `ast` when the goldens were
    made: tests/golden/ref_signatures.json)."""
    import inspect
    import json
    import os

    from conftest import GOLDEN
    from scarf_b200.datastore import DataStore

    with open(os.path.join(GOLDEN, "ref_signatures.json")) as f:
        ref = json.load(f)
    assert sorted(ref) == ["load_graph", "make_graph", "mark_hvgs", "run_mapping"]
    for name, params in ref.items():
        sig = [p for p in list(inspect.signature(getattr(DataStore, name)).parameters.values())[1:]
               if p.kind != p.VAR_KEYWORD]
        assert [p.name for p in sig] == [a for a, _ in params], name
        for p, (a, default) in zip(sig, params):
            if default is None:
                assert p.default is inspect.Parameter.empty, (name, a)
            else:
                assert p.default == eval(default, {"np": np}), (name, a, p.default, default)


def test_mark_hvgs_glue_on_cpu_tensors(pbmc, monkeypatch):
    """graph.mark_hvgs_csr's tensor route (what DataStore.mark_hvgs runs: statistics wanted, or min_var / max_var /
    keep_bounds given) with the statistics kernel replaced by the oracle's numbers, so that the glue -- argument
    passing, blacklist, trend removal, choice, returned statistics -- runs here without a GPU."""
    import torch

    from oracle import pipeline as P
    from scarf_b200 import graph
    from scarf_b200.ops import CsrDevice

    counts, cell_idx, names = pbmc["counts"], pbmc["cell_idx"], pbmc["names"]
    feat_I = P.gene_ncells(counts) > 20
    hv_o, st_o = P.mark_hvgs(counts, cell_idx, feat_I, gene_names=names, top_n=100, return_stats=True)
    full = {k: torch.from_numpy(np.nan_to_num(np.asarray(st_o[k], dtype=np.float64)))
            for k in ("normed_n", "normed_tot", "sigmas", "avg", "nz_mean")}
    monkeypatch.setattr(graph, "hvg_gene_stats", lambda *a, **k: dict(full))
    csr = CsrDevice(torch.zeros(893, dtype=torch.int64), torch.zeros(0, dtype=torch.int32),
                    torch.zeros(0, dtype=torch.int32), 892, counts.shape[1])  # never read: the statistics are patched
    n_counts = torch.ones(892, dtype=torch.float64)
    cells = torch.from_numpy(cell_idx)
    mask, st = graph.mark_hvgs_csr(csr, cells, feat_I, n_counts, 892, gene_names=names, top_n=100, return_stats=True)
    assert np.array_equal(mask, hv_o) and set(st) >= {"c_var", "avg", "sigmas", "normed_n", "nz_mean", "normed_tot"}
    np.testing.assert_allclose(st["c_var"][feat_I], st_o["c_var"][feat_I], rtol=1e-10)
    kw = dict(min_cells=8, max_cells=700.0, min_mean=-3.0, max_mean=2.0)
    hi = float(np.log2(np.nanpercentile(st_o["c_var"][feat_I], 99.8)))  # cuts the very top of the corrected variances
    for extra in (dict(top_n=50), dict(top_n=50, keep_bounds=True), dict(top_n=50, max_var=hi),
                  dict(top_n=50, min_var=0.0, max_var=hi)):
        got = graph.mark_hvgs_csr(csr, cells, feat_I, n_counts, 892, gene_names=names, **kw, **extra)
        want = P.choose_hvgs(st_o["normed_n"], st_o["nz_mean"], st_o["c_var"], feat_I, names, **kw, **extra)
        assert np.array_equal(got, want) and want.sum() > 0, extra
    with pytest.raises(ValueError, match="greater than 0"):
        graph.mark_hvgs_csr(csr, cells, feat_I, n_counts, 892, gene_names=names, top_n=0)


def test_datastore_mark_hvgs_front_end_on_a_stub_store(pbmc, monkeypatch):
    """DataStore.mark_hvgs (reference signature: cell_key None, max_cells, keep_bounds, min_var / max_var) driven on a
    stub store with CPU tensors and the statistics kernel patched with the oracle's numbers: the columns written to the
    feature table and the HVG mask are the oracle's."""
    import torch
    from types import SimpleNamespace

    from oracle import pipeline as P
    from scarf_b200 import graph
    from scarf_b200.datastore import DataStore, RNAassay
    from scarf_b200.ops import CsrDevice

    counts, cell_idx, names = pbmc["counts"], pbmc["cell_idx"], pbmc["names"]
    g = counts.shape[1]
    feat_I = P.gene_ncells(counts) > 20
    hv_o, st_o = P.mark_hvgs(counts, cell_idx, feat_I, gene_names=names, top_n=100, return_stats=True)
    full = {k: torch.from_numpy(np.nan_to_num(np.asarray(st_o[k], dtype=np.float64)))
            for k in ("normed_n", "normed_tot", "sigmas", "avg", "nz_mean")}
    monkeypatch.setattr(graph, "hvg_gene_stats", lambda *a, **k: dict(full))
    keep = np.zeros(892, dtype=bool)
    keep[cell_idx] = True
    written = {}

    def insert(name, values, fill_value=np.nan, key="I", overwrite=False):
        written[name] = np.asarray(values)

    feats = SimpleNamespace(fetch_all=lambda c: {"I": feat_I, "names": names}[c], insert=insert, N=g)
    assay = object.__new__(RNAassay)
    assay.feats, assay.name, assay.sf = feats, "RNA", 1000
    assay._span = (0, 892)  # the whole (empty) matrix counts as loaded
    assay._span_csr = CsrDevice(torch.zeros(893, dtype=torch.int64), torch.zeros(0, dtype=torch.int32),
                                torch.zeros(0, dtype=torch.int32), 892, g)
    assay.chunk_rows, assay.n_rows, assay.n_cols = 1000, 892, g
    assay.cells = SimpleNamespace(fetch_all=lambda c: np.ones(892))  # RNA_nCounts (unused: statistics are patched)
    cells = SimpleNamespace(columns=["I", "ids", "names"], N=892, active_index=lambda k: np.where(keep)[0])
    store = SimpleNamespace(cells=cells, _defaultAssay="RNA", _get_assay=lambda a: assay, device=torch.device("cpu"),
                            comm=None, _world=1, _rank=0, _barrier=lambda: None,
                            _shard=lambda ci, al: (0, int(ci.size), 0, 892))
    assert DataStore.mark_hvgs(store, top_n=100) is None  # cell_key None -> "I", min_cells None -> int(0.01 * N)
    assert np.array_equal(written["I__hvgs"], hv_o[feat_I]) and written["I__hvgs"].sum() == 100
    assert set(written) == {"I__normed_tot", "I__avg", "I__nz_mean", "I__sigmas", "I__normed_n", "I__c_var__200__0.1",
                            "I__hvgs"}
    np.testing.assert_allclose(written["I__c_var__200__0.1"], st_o["c_var"][feat_I], rtol=1e-10)
    DataStore.mark_hvgs(store, top_n=30, max_cells=600.0, keep_bounds=True, hvg_key_name="few", show_plot=False)
    want = P.choose_hvgs(st_o["normed_n"], st_o["nz_mean"], st_o["c_var"], feat_I, names, top_n=30, min_cells=8,
                         max_cells=600.0, keep_bounds=True)
    assert np.array_equal(written["I__few"], want[feat_I]) and want.sum() == 31
    with pytest.raises(ValueError, match="not found in cell metadata"):
        DataStore.mark_hvgs(store, cell_key="nope")


def test_datastore_load_graph_on_a_store_without_a_gpu(tmp_path):
    """DataStore.load_graph reads the Zarr layout make_graph writes (knn__k/indices, graph__lc__bw/{edges,weights}) and
    returns what the reference's load_graph returned on the same arrays (tests/golden/ref_functions.npz)."""
    import os
    from types import SimpleNamespace

    from conftest import GOLDEN
    from scarf_b200.datastore import DataStore
    from scarf_b200.zarr_store import open_group

    ref = np.load(os.path.join(GOLDEN, "ref_functions.npz"))
    n, k = int(ref["graph_n"]), int(ref["graph_k"])
    root = open_group(str(tmp_path / "g.zarr"), "w")
    knn = "RNA/normed__I__hvgs/reduction__pca__11__I/ann__l2__50__50__48__4466/knn__5"
    gl = f"{knn}/graph__1.0__1.5"
    root.create_group(gl)
    a = root[knn].create_dataset("indices", (n, k), "u8", (1000, k))
    a[:] = ref["graph_edges"][:, 1].reshape(n, k)
    e = root[gl].create_dataset("edges", (n * k, 2), "u8", (1000 * k, 2))
    e[:] = ref["graph_edges"]
    w = root[gl].create_dataset("weights", (n * k,), "f8", (1000 * k,))
    w[:] = ref["graph_weights"]
    store = SimpleNamespace(zw=root, _get_latest_keys=lambda a_, c, f: ("RNA", "I", "hvgs"),
                            _get_latest_graph_loc=lambda a_, c, f: gl)
    for tag, kw in (("default", {}), ("sym_upper", dict(symmetric=True, upper_only=True)),
                    ("raw_k3", dict(symmetric=False, use_k=3)), ("sym_k0", dict(symmetric=True, use_k=0))):
        g = DataStore.load_graph(store, **kw)
        assert np.array_equal(np.asarray(g.todense()), ref[f"graph_{tag}"]), tag
    with pytest.raises(ValueError, match="not found in zarr location"):
        DataStore.load_graph(store, graph_loc="RNA/none")


def test_datastore_constructor_follows_the_reference(tmp_path, monkeypatch):
    """DataStore.__init__ (scarf/datastore/datastore.py:46-90, base_datastore.py:77-186): the reference's parameter
    list, default-assay resolution (explicit -> `defaultAssay` attribute -> the only assay), the every-open cell filter
    on nFeatures, the error cases.  Runs on CPU tensors: the store already holds every first-open column, so no kernel
    is needed (the CUDA check is patched out for this test only)."""
    import inspect
    import json
    import os

    import scipy.sparse as sp
    import torch

    from scarf_b200.datastore import DataStore
    from scarf_b200.zarr_store import open_group

    params = list(inspect.signature(DataStore.__init__).parameters.values())[1:]
    assert [p.name for p in params if p.kind == p.POSITIONAL_OR_KEYWORD] == [
        "zarr_loc", "assay_types", "default_assay", "min_features_per_cell", "min_cells_per_feature", "mito_pattern",
        "ribo_pattern", "nthreads", "zarr_mode", "workspace", "synchronizer"]
    rng = np.random.default_rng(5)
    m = sp.random(40, 30, density=0.4, random_state=5, format="csr", dtype=np.float64)
    m.data = np.ceil(m.data * 5).astype(np.uint32)
    m = m.astype(np.uint32)
    m.sort_indices()
    dense = m.toarray()
    loc = str(tmp_path / "s.zarr")
    root = open_group(loc, "w")

    def put(grp, name, arr):
        grp.create_dataset(name, arr.shape, arr.dtype, (100000,))[:] = arr

    cg = root.create_group("cellData")
    put(cg, "I", np.ones(40, dtype=bool)), put(cg, "ids", np.array([f"c{i}" for i in range(40)]))
    put(cg, "names", np.array([f"c{i}" for i in range(40)]))
    put(cg, "RNA_nCounts", dense.sum(1).astype(np.float64)), put(cg, "RNA_nFeatures", (dense > 0).sum(1).astype(np.float64))
    for name in ("RNA", "ADT"):
        ag = root.create_group(name)
        ag.attrs["is_assay"] = True
        fg = ag.create_group("featureData")
        put(fg, "I", np.ones(30, dtype=bool)), put(fg, "ids", np.array([f"g{i}" for i in range(30)]))
        put(fg, "names", np.array([f"g{i}" for i in range(30)]))
        put(fg, "nCells", (dense > 0).sum(0).astype(np.float64)), put(fg, "dropOuts", 40.0 - (dense > 0).sum(0))
        rg = ag.create_group("counts_csr")
        rg.attrs["shape"] = [40, 30]
        put(rg, "indptr", m.indptr.astype(np.int64)), put(rg, "indices", m.indices.astype(np.int32))
        put(rg, "data", m.data.astype(np.uint32))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    with pytest.raises(ValueError, match="more than one assay"):
        DataStore(loc, device="cpu")
    with pytest.raises(ValueError, match="was not found"):
        DataStore(loc, default_assay="rna", device="cpu")
    with pytest.raises(ValueError, match="'r' or 'r\\+'"):
        DataStore(loc, default_assay="RNA", zarr_mode="w", device="cpu")
    with pytest.raises(NotImplementedError):
        DataStore(loc, {"RNA": "ATAC"}, "RNA", device="cpu")
    n_feat = (dense > 0).sum(1)
    cut = int(np.sort(n_feat)[5])  # a threshold below the median: the filter applies (base_datastore.py:384-399)
    ds = DataStore(loc, None, "RNA", cut, 20, None, None, 4, "r+", device="cpu")  # positional, as the reference takes them
    assert ds._defaultAssay == "RNA" and ds.nthreads == 4 and ds.assay_names == ["ADT", "RNA"]
    assert json.load(open(os.path.join(loc, ".zattrs")))["defaultAssay"] == "RNA"
    assert np.array_equal(ds.cells.fetch_all("I"), n_feat > cut) and 0 < (n_feat > cut).sum() < 40
    assert np.array_equal(ds.RNA.csr.indices.numpy(), m.indices) and ds.RNA.csr.n_cols == 30
    # reopening: the remembered default assay, and the filter runs again with the default threshold of 10 (it is
    # applied at every open, base_datastore.py:384-399); a threshold above the median of the kept cells changes nothing
    ds2 = DataStore(loc, device="cpu")
    assert cut < 10 <= np.median(n_feat[n_feat > cut])
    assert ds2._defaultAssay == "RNA" and np.array_equal(ds2.cells.fetch_all("I"), n_feat > 10)
    ds3 = DataStore(loc, min_features_per_cell=10 ** 6, device="cpu")
    assert np.array_equal(ds3.cells.fetch_all("I"), n_feat > 10)
