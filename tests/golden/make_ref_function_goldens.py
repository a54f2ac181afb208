"""Generates tests/golden/ref_functions.npz by EXECUTING reference functions of the path on seeded inputs.

The reference package cannot be imported here (`import scarf` needs dask / zarr / numcodecs / hnswlib), but the
functions below are plain numpy / pandas code: their source is cut out of the reference files with `ast` and executed
unchanged (nothing is copied into this repository; the GPU box never runs this script).

    python tests/golden/make_ref_function_goldens.py        (needs /root/reference)

Functions executed (reference file:lines) and the scope-table rows they pin:
  norm_lib_size, norm_lib_size_log   scarf/assay.py:41-64       a4  (library-size normalisation, log1p)
  clean_array                        scarf/utils.py:143-153     a6  (mu / sigma clean-up)
  fix_knn_query                      scarf/ann.py:31-52         a10 (self removal in a k+1 query result)
  _order_features                    scarf/mapping_utils.py:98-145  a14 (source / target feature alignment)
  fit_lowess                         scarf/feat_utils.py:11-45  a3  (log-log binning, minimum-variance gene per bin,
                                     corrected variance) -- executed with statsmodels' `lowess` (not installed) replaced
                                     by the restated smoother oracle/lowess.py: everything around the smoother is the
                                     reference's own code
  RNAassay.mark_hvgs                 scarf/assay.py:945-1063    a3  (HVG choice) -- the method body after
                                     `set_summary_stats`, run on a stub assay whose feature table serves columns from
                                     a dict through the reference's own MetaData.sift / multi_sift / grep /
                                     get_index_by / index_to_bool (scarf/metadata.py:339-394,483-533,569-584)
  AnnStream._fit_pca, transform_z    scarf/ann.py:191-192,207-256  a8 (the block loop around sklearn's IncrementalPCA:
                                     first block kept as the end reservoir, carry-over of blocks shorter than
                                     dims + 1, one extra component dropped) -- run on a stub stream of numpy blocks
                                     with the installed scikit-learn (1.9; the reference pins 1.6.1, same algorithm)
  GraphDataStore._store_to_sparse, load_graph   scarf/datastore/graph_datastore.py:474-511,1022-1075  f-4 (use_k
                                     clamping and row-wise truncation, g + g.T - g * g.T, upper triangle) -- run on a
                                     stub store holding the edges / weights arrays in a dict
  smoothen_dists                     scarf/knn_utils.py:89-159  a11 (the chunk loop: chunk-local row ids shifted by
                                     `last_row`, per-chunk float32 cast, zero-weight entries set to the smallest
                                     non-zero weight seen) -- run with umap-learn's two numba functions (not
                                     installed) replaced by the restatements that reproduce the reference's
                                     `knn_weights.npy` golden (oracle/pipeline.py), on a stub store of numpy arrays
  GraphDataStore._set_graph_params, _choose_reduction_method   scarf/datastore/graph_datastore.py:24-363  a12 (explicit ->
                                     cached -> default resolution of make_graph's parameters and the group names built
                                     from them) -- run on stub stores (a dict of nodes with attrs); the scenarios and
                                     the resolved tuples go to tests/golden/ref_graph_params.json
The per-cell scalar of the renormalised branch (`RNAassay.normed`, scarf/assay.py:814-823) is inline code of a method
that needs a store; the three lines are restated below where the scalar is built.
"""
import ast
import os
from types import SimpleNamespace
from typing import Tuple

import numpy as np
import pandas as pd

REF = "/root/reference/scarf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_functions.npz")


def ref_function(rel_path, name, extra=None, cls=None):
    """The function `name` of a reference file (a method of class `cls` when given), compiled from its own source."""
    with open(os.path.join(REF, rel_path)) as f:
        src = f.read()
    body = ast.parse(src).body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "pd": pd, "Tuple": Tuple, "daskArrayType": object, "__name__": "ref"}
    ns.update(extra or {})
    exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, rel_path), "exec"), ns)
    return ns[name]


rng = np.random.default_rng(20261017)
out = {}

# ---- a4: normalisation ------------------------------------------------------------------------------------------
norm_lib_size = ref_function("assay.py", "norm_lib_size")
norm_lib_size_log = ref_function("assay.py", "norm_lib_size_log")
counts = ((rng.random((60, 90)) < 0.15) * rng.integers(1, 200, (60, 90))).astype(np.uint32)
counts[7] = 0  # a cell without any count
feat_idx = np.sort(rng.choice(90, 25, replace=False))
cell_idx = np.sort(rng.choice(60, 50, replace=False))
cell_idx = np.union1d(cell_idx, [7])
sub = counts[cell_idx][:, feat_idx]
scalar = sub.sum(axis=1)  # assay.py:819-823: counts.sum(axis=1) over the feature subset, zeros replaced by one
scalar[scalar == 0] = 1
assay = SimpleNamespace(sf=1000, scalar=scalar)
out.update(norm_counts=counts, norm_cell_idx=cell_idx, norm_feat_idx=feat_idx,
           norm_lib_size=np.asarray(norm_lib_size(assay, sub), dtype=np.float64),
           norm_lib_size_log=np.asarray(norm_lib_size_log(assay, sub), dtype=np.float64))
n_counts = counts.sum(axis=1).astype(np.float64)  # renormalize_subset=False: scalar = nCounts of the cells
n_counts[n_counts == 0] = 1
assay_nc = SimpleNamespace(sf=1000, scalar=n_counts[cell_idx])
out.update(norm_n_counts=n_counts, norm_lib_size_log_ncounts=np.asarray(norm_lib_size_log(assay_nc, sub), dtype=np.float64))

# ---- a6: clean_array ----------------------------------------------------------------------------------------------
clean_array = ref_function("utils.py", "clean_array")
x = rng.normal(size=40)
x[[1, 5, 9, 13, 17, 21]] = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-300]
out.update(clean_in=x, clean_fill0=clean_array(x.copy()), clean_fill1=clean_array(x.copy(), 1))

# ---- a10: fix_knn_query -------------------------------------------------------------------------------------------
fix_knn_query = ref_function("ann.py", "fix_knn_query")
n, k1 = 30, 6
ind = np.stack([rng.choice(200, k1, replace=False) for _ in range(n)]).astype(np.uint64)
ref_idx = np.arange(100, 100 + n).astype(np.uint64)
dist = np.sort(rng.random((n, k1)).astype(np.float32), axis=1)
for r in range(n):
    ind[r][ind[r] == ref_idx[r]] = 999  # no accidental self
    mode = r % 3
    if mode == 0:
        ind[r, 0] = ref_idx[r]  # self found first (the normal case)
    elif mode == 1:
        ind[r, 1 + r % (k1 - 1)] = ref_idx[r]  # self found further down (ties / approximate search)
    # mode 2: self not found at all -> the last neighbour is dropped
fi, fd, n_mis = fix_knn_query(ind, dist, ref_idx)
out.update(fix_ind=ind, fix_dist=dist, fix_ref_idx=ref_idx, fix_out_ind=fi, fix_out_dist=fd, fix_n_mis=np.int64(n_mis))

# ---- a14: _order_features -----------------------------------------------------------------------------------------
order = ref_function("mapping_utils.py", "_order_features", {"logger": SimpleNamespace(warning=lambda *a, **k: None),
                                                            "controlled_compute": None})
s_ids = np.array([f"G{i:03d}" for i in range(40)])
t_ids = np.array([f"G{i:03d}" for i in rng.permutation(60)[:35]] + ["X1", "X2"])  # reordered, some missing, some extra
s_feat_ids = s_ids[np.sort(rng.choice(40, 18, replace=False))]
mk = lambda ids: SimpleNamespace(feats=SimpleNamespace(fetch_all=lambda col, ids=ids: ids))
s_idx, t_re_idx = order(mk(s_ids), mk(t_ids), s_feat_ids, filter_null=False, exclude_missing=False, nthreads=1)
out.update(order_s_ids=s_ids, order_t_ids=t_ids, order_s_feat_ids=s_feat_ids, order_s_idx=np.asarray(s_idx, dtype=np.int64),
           order_t_re_idx=np.asarray(t_re_idx, dtype=np.int64))

# ---- a3: fit_lowess (binning + corrected variance; the smoother itself is the oracle's restatement) -------------------
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.lowess import lowess as oracle_lowess  # noqa: E402

sm = types.ModuleType("statsmodels.nonparametric.smoothers_lowess")
sm.lowess = lambda endog, exog, return_sorted=False, frac=2.0 / 3.0, it=3: oracle_lowess(endog, exog, frac=frac, it=it)
for name in ("statsmodels", "statsmodels.nonparametric"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["statsmodels.nonparametric.smoothers_lowess"] = sm
fit_lowess = ref_function("feat_utils.py", "fit_lowess")
n_genes = 3000
mean = rng.gamma(0.3, 1.0, n_genes) + 1e-4  # gene means over four decades, variance ~ mean * (1 + mean * dispersion)
var = mean * (1.0 + mean * rng.gamma(2.0, 0.5, n_genes)) * np.exp(rng.normal(0.0, 0.3, n_genes))
out.update(lowess_avg=mean, lowess_var=var, lowess_c_var=np.asarray(fit_lowess(mean, var, 200, 0.1), dtype=np.float64))

# ---- a3: RNAassay.mark_hvgs (the HVG choice) -------------------------------------------------------------------------
import re
from typing import Any, Iterable, List, Optional

quiet = SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None)
md_ns = {"re": re, "Iterable": Iterable, "List": List, "Any": Any, "Optional": Optional, "logger": quiet,
         "_all_true": ref_function("metadata.py", "_all_true")}


class FeatureTable:
    """Columns from a dict; every other method is the reference's MetaData code."""

    def __init__(self, cols):
        self.cols, self.N = dict(cols), len(cols["I"])

    def fetch_all(self, column):
        return self.cols[column]

    def insert(self, column_name, values, fill_value=np.nan, key="I", overwrite=False, location="primary"):
        self.cols[column_name] = np.asarray(values)


for m in ("sift", "multi_sift", "grep", "get_index_by", "index_to_bool"):
    setattr(FeatureTable, m, ref_function("metadata.py", m, md_ns, cls="MetaData"))
mark_hvgs = ref_function("assay.py", "mark_hvgs", {"logger": quiet}, cls="RNAassay")

g = 1200
names = np.array([f"Gene{i}" for i in range(g)], dtype=object)
for i, nm in zip(rng.choice(g, 14, replace=False), ["MT-CO1", "mt-Co2", "Rps3", "RPL13", "Mrps7", "MRPL2", "Ccnb1", "HLA-A",
                                                     "H2-K1", "Hist1h1a", "hist2h2aa", "xMT-1", "RPS", "Hla-dra"]):
    names[i] = nm
names = names.astype("U")
feat_I = rng.random(g) < 0.85
normed_n = np.floor(rng.gamma(2.0, 200.0, g))
nz_mean = rng.gamma(1.5, 1.0, g) + 0.05
c_var = np.exp(rng.normal(0.0, 0.6, g))
c_var[rng.choice(g, 6, replace=False)] = c_var[0]  # ties at arbitrary places
c_var[~feat_I] = 0.0  # what MetaData.insert leaves outside the active rows (fill value of remove_trend's column)
cases = [dict(top_n=100, min_cells=8, max_cells=np.inf, min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf),
         dict(top_n=40, min_cells=50, max_cells=900.0, min_mean=-2.0, max_mean=1.5, min_var=-np.inf, max_var=np.inf),
         dict(top_n=5000, min_cells=8, max_cells=np.inf, min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf),
         dict(top_n=10, min_cells=8, max_cells=np.inf, min_mean=-np.inf, max_mean=np.inf, min_var=-0.5, max_var=1.0),
         # keep_bounds=True: >= / <= in every sift; bounds placed ON values of the data so that it matters
         dict(top_n=60, min_cells=float(np.sort(normed_n)[300]), max_cells=float(np.sort(normed_n)[1100]),
              min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf, keep_bounds=True)]
out.update(hvg_names=names, hvg_feat_I=feat_I, hvg_normed_n=normed_n, hvg_nz_mean=nz_mean, hvg_c_var=c_var,
           hvg_blacklist=np.array("^MT-|^RPS|^RPL|^MRPS|^MRPL|^CCN|^HLA-|^H2-|^HIST"))
for ci, case in enumerate(cases):
    feats = FeatureTable({"I": feat_I, "names": names, "I__normed_n": normed_n, "I__nz_mean": nz_mean,
                          "I__c_var__200__0.1": c_var})
    stub = SimpleNamespace(feats=feats, set_summary_stats=lambda ck, nb, lf: ("I_", "c_var__200__0.1"))
    # col_renamer gives f"{identifier}_{x}": identifier "I_" -> columns "I__normed_n", ... (scarf/assay.py:1007-1011)
    case = dict(case)
    keep_bounds = case.pop("keep_bounds", False)
    mark_hvgs(stub, cell_key="I", n_bins=200, lowess_frac=0.1, blacklist=str(out["hvg_blacklist"]), hvg_key_name="hvgs",
              keep_bounds=keep_bounds, show_plot=False, **case)
    out[f"hvg_case{ci}_params"] = np.array([case[k_] for k_ in ("top_n", "min_cells", "max_cells", "min_mean", "max_mean",
                                                               "min_var", "max_var")] + [float(keep_bounds)],
                                           dtype=np.float64)
    out[f"hvg_case{ci}_mask"] = feats.cols["I__hvgs"].astype(bool)
    print("hvg case", ci, int(out[f"hvg_case{ci}_mask"].sum()))

# ---- a8: AnnStream._fit_pca (IncrementalPCA block loop) --------------------------------------------------------------
ann_ns = {"logger": quiet}
fit_pca = ref_function("ann.py", "_fit_pca", ann_ns, cls="AnnStream")
transform_z = ref_function("ann.py", "transform_z", ann_ns, cls="AnnStream")
n_pca, h_pca, dims_pca, bs = 2005, 40, 7, 1000  # blocks of 1000, 1000 and 5 rows: the last one is shorter than dims + 1
lat = rng.normal(size=(n_pca, 10)) * (3.0 * 0.8 ** np.arange(10))
x_pca = np.abs(lat @ rng.normal(size=(10, h_pca)) + rng.normal(size=(n_pca, h_pca)))
mu_pca, sigma_pca = x_pca.mean(axis=0), x_pca.std(axis=0)


class Stream:
    def __init__(self):
        self.dims, self.batchSize, self.nCells, self.mu, self.sigma = dims_pca, bs, n_pca, mu_pca, sigma_pca

    def iter_blocks(self, msg=""):
        for s_ in range(0, n_pca, bs):
            yield x_pca[s_:s_ + bs]

    transform_z = transform_z


st = Stream()
fit_pca(st, False, np.ones(n_pca, dtype=bool))
out.update(pca_x=x_pca, pca_mu=mu_pca, pca_sigma=sigma_pca, pca_dims=np.int64(dims_pca), pca_batch=np.int64(bs),
           pca_loadings=st.loadings.copy())
print("ipca loadings", st.loadings.shape)

# ---- f-4: _store_to_sparse + load_graph ---------------------------------------------------------------------------------
from scipy.sparse import coo_matrix, csr_matrix  # noqa: E402

gd_ns = {"logger": SimpleNamespace(debug=lambda *a, **k: None), "csr_matrix": csr_matrix, "coo_matrix": coo_matrix,
         "Optional": Optional}
store_to_sparse = ref_function("datastore/graph_datastore.py", "_store_to_sparse", gd_ns, cls="GraphDataStore")
load_graph = ref_function("datastore/graph_datastore.py", "load_graph", gd_ns, cls="GraphDataStore")
n_g, k_g = 60, 5
nbrs = np.stack([rng.choice(np.delete(np.arange(n_g), r), k_g, replace=False) for r in range(n_g)])
edges = np.stack([np.repeat(np.arange(n_g), k_g), nbrs.reshape(-1)], axis=1).astype(np.uint64)
weights = np.sort(rng.random((n_g, k_g)), axis=1)[:, ::-1].reshape(-1).copy()


class GraphStore:
    zw = {"g": {"edges": edges, "weights": weights}}
    _store_to_sparse = store_to_sparse

    def _get_graph_ncells_k(self, loc):
        return n_g, k_g

    def _get_latest_keys(self, a, c, f):
        return "RNA", "I", "hvgs"


out.update(graph_edges=edges, graph_weights=weights, graph_n=np.int64(n_g), graph_k=np.int64(k_g))
for tag, kw in (("default", dict()), ("sym_upper", dict(symmetric=True, upper_only=True)),
                ("sym_full", dict(symmetric=True, upper_only=False)), ("raw_k3", dict(symmetric=False, use_k=3)),
                ("sym_k0", dict(symmetric=True, use_k=0)), ("sym_upper_k9", dict(symmetric=True, upper_only=True, use_k=9))):
    gmat = load_graph(GraphStore(), graph_loc="g", **kw)
    out[f"graph_{tag}"] = np.asarray(gmat.todense())
    print("graph", tag, type(gmat).__name__, gmat.nnz)

# ---- a11: smoothen_dists (chunk loop around the two umap-learn functions) ------------------------------------------------
from oracle import pipeline as oracle_pipeline  # noqa: E402

umap_mod = types.ModuleType("umap.umap_")
umap_mod.smooth_knn_dist = oracle_pipeline.smooth_knn_dist_vec
umap_mod.compute_membership_strengths = oracle_pipeline.compute_membership_strengths
sys.modules.setdefault("umap", types.ModuleType("umap"))
sys.modules["umap.umap_"] = umap_mod
created = {}


def create_zarr_dataset(store, name, chunks, dtype, shape):
    created[name] = np.zeros(shape, dtype=np.uint64 if isinstance(dtype, tuple) else dtype)
    return created[name]


smoothen = ref_function("knn_utils.py", "smoothen_dists", {"create_zarr_dataset": create_zarr_dataset,
                                                           "tqdmbar": lambda it, **k: it,
                                                           "_is_umap_version_new": lambda: False})
n_s, k_s, chunk = 2300, 6, 1000  # three chunks, the last one ragged
idx_s = np.stack([rng.choice(np.delete(np.arange(n_s), r), k_s, replace=False) for r in range(n_s)]).astype(np.uint64)
dist_s = np.sort(rng.gamma(2.0, 1.0, (n_s, k_s)), axis=1).astype(np.float32).astype(np.float64)
dist_s[[5, 1200, 2299], -2:] = [4.0e4, 9.0e5]  # two very far neighbours
dist_s[77] = dist_s[77, 0]  # all neighbours equally far
dist_s[9, :3] = 0.0  # zero distances (duplicate cells)
dist_s[1500] = 0.0
# umap's membership function zeroes an entry whose neighbour id equals the row id INSIDE THE CHUNK it is called on:
# with chunks this hits global ids that are no self loops; such zeros are then raised to the smallest non-zero weight
idx_s[1007, 4] = 7
idx_s[2100, 0] = 100
idx_s[3, 2] = 3
smoothen(None, idx_s, dist_s, 1.0, 1.5, chunk)
out.update(smooth_idx=idx_s, smooth_dist=dist_s, smooth_chunk=np.int64(chunk), smooth_edges=created["edges"].copy(),
           smooth_weights=created["weights"].copy())
print("smoothen_dists: entries at the floor:", int((created["weights"] == created["weights"].min()).sum()),
      "floor", created["weights"].min())

# ---- a12: _set_graph_params ----------------------------------------------------------------------------------------------
import json

gp_ns = {"logger": SimpleNamespace(debug=lambda *a, **k: None, info=lambda *a, **k: None), "Assay": object}
set_graph_params = ref_function("datastore/graph_datastore.py", "_set_graph_params", gp_ns, cls="GraphDataStore")
choose_reduction = ref_function("datastore/graph_datastore.py", "_choose_reduction_method", gp_ns, cls="GraphDataStore")


class RNAassay:  # the class NAME is what `_choose_reduction_method` looks at
    pass


def graph_store(tree):
    cells = SimpleNamespace(columns=["I", "ids", "names", "sub", "RNA_nCounts"],
                            get_dtype=lambda c: bool if c in ("I", "sub") else float)
    return SimpleNamespace(zw={k_: SimpleNamespace(attrs=dict(v)) for k_, v in tree.items()}, cells=cells,
                           _get_assay=lambda a: RNAassay(), _choose_reduction_method=choose_reduction)  # (a staticmethod there)


base = "RNA/normed__I__hvgs"
red = f"{base}/reduction__pca__25__I"
ann = f"{red}/ann__l2__63__70__48__99"
knn = f"{ann}/knn__21"
cached_tree = {base: {"subset_params": {"log_transform": False, "renormalize_subset": True}, "latest_reduction": red},
               red: {"latest_ann": ann, "latest_kmeans": f"{red}/kmeans__300__99"},
               ann: {"latest_knn": knn}, knn: {"latest_graph": f"{knn}/graph__2.0__1.25"}}
scenarios = [
    {"tree": {}, "kwargs": {}},
    {"tree": {}, "kwargs": {"dims": 60, "k": 40, "reduction_method": "PCA"}},
    {"tree": {}, "kwargs": {"dims": 30, "pca_cell_key": "sub", "ann_m": 20, "ann_efc": 10, "ann_ef": 12, "rand_state": 7,
                            "n_centroids": 50, "local_connectivity": 2.5, "bandwidth": 0.5, "log_transform": False,
                            "renormalize_subset": False, "ann_metric": "l2"}},
    {"tree": cached_tree, "kwargs": {}},
    {"tree": cached_tree, "kwargs": {"k": 11}},
    {"tree": cached_tree, "kwargs": {"dims": 11}},
    {"tree": cached_tree, "kwargs": {"log_transform": True, "bandwidth": 3.0}},
    # pca_cell_key is validated only when dims is NOT given (the check sits in the else of an inner if,
    # graph_datastore.py:188-208): the outcomes below are recorded, not presumed
    {"tree": {}, "kwargs": {"dims": 5, "pca_cell_key": "nope"}},
    {"tree": {}, "kwargs": {"dims": 5, "pca_cell_key": "RNA_nCounts"}},
    {"tree": {}, "kwargs": {"pca_cell_key": "nope"}},
    {"tree": {}, "kwargs": {"pca_cell_key": "RNA_nCounts"}},
    {"tree": {}, "kwargs": {"pca_cell_key": "sub"}},
    {"tree": {}, "kwargs": {"reduction_method": "umap"}},
]
for sc in scenarios:
    try:
        res_t = set_graph_params(graph_store(sc["tree"]), "RNA", "I", "hvgs", **sc["kwargs"])
        sc["result"] = [bool(v) if isinstance(v, (bool, np.bool_)) else v for v in res_t]
    except (ValueError, TypeError) as err:
        sc["raises"] = type(err).__name__
with open(os.path.join(os.path.dirname(OUT), "ref_graph_params.json"), "w") as f:
    json.dump(scenarios, f, indent=1)
print("graph params scenarios:", len(scenarios))

# ---- (b): public signatures ----------------------------------------------------------------------------------------------
signatures = {}
for rel, cls, name in (("datastore/graph_datastore.py", "GraphDataStore", "make_graph"),
                       ("datastore/graph_datastore.py", "GraphDataStore", "load_graph"),
                       ("datastore/mapping_datastore.py", "MappingDatastore", "run_mapping"),
                       ("datastore/datastore.py", "DataStore", "mark_hvgs")):
    with open(os.path.join(REF, rel)) as f:
        cdef = next(n for n in ast.parse(f.read()).body if isinstance(n, ast.ClassDef) and n.name == cls)
    fdef = next(n for n in cdef.body if isinstance(n, ast.FunctionDef) and n.name == name)
    args = fdef.args.args[1:]
    defaults = [None] * (len(args) - len(fdef.args.defaults)) + [ast.unparse(d) for d in fdef.args.defaults]
    signatures[name] = [[a.arg, d] for a, d in zip(args, defaults)]
with open(os.path.join(os.path.dirname(OUT), "ref_signatures.json"), "w") as f:
    json.dump(signatures, f, indent=1)

np.savez_compressed(OUT, **out)
print("ok", OUT, os.path.getsize(OUT), "bytes;", "mismatching self rows:", int(n_mis), "; missing target features:",
      int((np.asarray(t_re_idx) == -1).sum()))
