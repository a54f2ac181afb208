"""Pins the CPU oracle on the reference's own golden vectors (scarf/tests/test_datastore.py:66-79).

The goldens were produced by the reference chain auto_filter_cells -> mark_hvgs(top_n=100) ->
make_graph(feat_key='hvgs') (scarf/tests/fixtures_datastore.py:58-73) on the PBMC fixture.
hnswlib is approximate, so the index pin is a stated recall / identical-row fraction.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pipeline as P


@pytest.fixture(scope="module")
def pbmc_oracle(pbmc):
    counts, cell_idx = pbmc["counts"], pbmc["cell_idx"]
    feat_I = P.gene_ncells(counts) > 20  # scarf/assay.py:225 (min_cells_per_feature default)
    hvgs = P.mark_hvgs(counts, cell_idx, feat_I, gene_names=pbmc["names"], top_n=100)
    out = P.make_graph(counts, cell_idx, hvgs, dims=11, k=11, return_all=True)
    out["hvgs"] = hvgs
    return out


def test_weights_golden(pbmc):
    """App. A.7 restatement on the reference's own indices/distances -> knn_weights.npy (1e-5)."""
    edges, w = P.smoothen_dists(pbmc["indices"], pbmc["distances"], 1.0, 1.5, 1000)
    assert np.abs(w - pbmc["weights"]).max() < 1e-5
    assert np.array_equal(edges[:, 1].reshape(pbmc["indices"].shape), pbmc["indices"])
    assert np.array_equal(edges[:, 0], np.repeat(np.arange(808, dtype=np.uint64), 11))


def test_weights_scalar_equals_vectorised(pbmc):
    _, w0 = P.smoothen_dists(pbmc["indices"][:200], pbmc["distances"][:200], vectorised=False)
    _, w1 = P.smoothen_dists(pbmc["indices"][:200], pbmc["distances"][:200], vectorised=True)
    assert np.array_equal(w0, w1)


def test_hvg_count(pbmc_oracle):
    assert pbmc_oracle["hvgs"].sum() == 100


def test_chain_indices_golden(pbmc, pbmc_oracle):
    """Whole chain vs knn_indices.npy: exact search vs the reference's HNSW graph."""
    idx, gi = pbmc_oracle["indices"], pbmc["indices"]
    recall = np.mean([len(set(a) & set(b)) / gi.shape[1] for a, b in zip(idx, gi)])
    assert recall > 0.995
    assert np.mean((idx == gi).all(axis=1)) > 0.98


def test_chain_distances_golden(pbmc, pbmc_oracle):
    """Rows whose neighbour list is identical must reproduce knn_distances.npy (test bar 1e-3)."""
    idx, gi = pbmc_oracle["indices"], pbmc["indices"]
    same = (idx == gi).all(axis=1)
    assert np.abs(pbmc_oracle["distances"][same] - pbmc["distances"][same]).max() < 1e-3


def test_chain_weights_golden(pbmc, pbmc_oracle):
    same = np.repeat((pbmc_oracle["indices"] == pbmc["indices"]).all(axis=1), 11)
    assert np.abs(pbmc_oracle["weights"][same] - pbmc["weights"][same]).max() < 1e-4


def test_exact_knn_c_equals_numpy():
    rng = np.random.default_rng(0)
    y = rng.normal(size=(300, 7)).astype(np.float32)
    y[10] = y[3]  # duplicate -> tie broken by index
    i0, d0 = P.exact_knn(y, y, 5, self_offset=0)
    i1, d1 = P.exact_knn_numpy(y, y, 5, self_offset=0)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    i0, d0 = P.exact_knn(y[:50], y, 3, self_offset=-1)
    i1, d1 = P.exact_knn_numpy(y[:50], y, 3, self_offset=-1)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert np.array_equal(i0[:, 0], np.where(np.arange(50) == 10, 3, np.arange(50)))


def test_sign_rule_matches_sklearn_svd_flip():
    from sklearn.utils.extmath import svd_flip

    rng = np.random.default_rng(1)
    u, s, vt = np.linalg.svd(rng.normal(size=(40, 12)), full_matrices=False)
    _, vt_f = svd_flip(u, vt, u_based_decision=False)
    assert np.allclose(P.sign_rule(vt), vt_f)


def test_run_mapping_golden(pbmc, pbmc_oracle):
    """run_mapping self-map (fixtures_datastore.py:146-153, save_k=3) -> cell_attributes.csv: mapping_scores
    (test_datastore.py:160-163, tolerance 1e-2).  The stored golden predates the division by save_k that
    get_mapping_score does today (mapping_datastore.py:282), hence per_k=False; exact PCA / exact search differ from
    IncrementalPCA / hnswlib on a handful of neighbour lists, so the pin is the fraction of cells within tolerance."""
    counts, cell_idx = pbmc["counts"], pbmc["cell_idx"]
    ids = np.arange(counts.shape[1])
    s_idx, t_re = P.order_features(ids, ids, np.where(pbmc_oracle["hvgs"])[0])
    assert np.array_equal(s_idx, np.where(pbmc_oracle["hvgs"])[0]) and np.array_equal(t_re, s_idx)
    xt = P.aligned_target(counts, cell_idx, t_re)
    assert np.array_equal(xt, pbmc_oracle["x"])  # self-map: the aligned target IS the reference's normalised data
    idx, dist, _ = P.run_mapping(pbmc_oracle["embedding"], pbmc_oracle["loadings"], pbmc_oracle["mu"],
                                 pbmc_oracle["sigma"], xt, save_k=3)
    assert np.array_equal(idx[:, 0], np.arange(len(cell_idx), dtype=idx.dtype))  # every cell finds itself first
    sc = P.mapping_score(idx, dist, len(cell_idx), per_k=False)
    assert (np.abs(sc - pbmc["mapping_scores"]) < 1e-2).mean() > 0.98


def test_aligned_target_missing_features_and_order():
    """Features the target lacks become columns of 1.0; present ones follow the SOURCE order (mapping_utils.py:200-211)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    t = sp.csr_matrix(rng.poisson(1.0, size=(6, 9)).astype(np.uint32))
    source_ids = np.array(["g%d" % i for i in range(12)])
    target_ids = np.array(["g7", "g2", "x1", "g5", "g0", "x2", "g9", "g3", "x3"])
    s_idx, t_re = P.order_features(source_ids, target_ids, source_ids[[0, 2, 4, 5, 9, 11]])
    assert s_idx.tolist() == [0, 2, 4, 5, 9, 11] and t_re.tolist() == [4, 1, -1, 3, 6, -1]
    a = P.aligned_target(t, np.arange(6), t_re)
    assert np.all(a[:, [2, 5]] == 1.0)
    dense = t.toarray().astype(np.float64)
    sub = dense[:, [4, 1, 3, 6]]
    tot = sub.sum(1, keepdims=True)
    tot[tot == 0] = 1
    np.testing.assert_allclose(a[:, [0, 1, 3, 4]], np.log1p(1000 * sub / tot), rtol=1e-12)


# ---- goldens produced by EXECUTING reference functions (tests/golden/make_ref_function_goldens.py) -----------------
@pytest.fixture(scope="module")
def ref_fn():
    return np.load(os.path.join(GOLDEN, "ref_functions.npz"))


def test_normalisation_equals_reference_functions(ref_fn):
    """a4: norm_lib_size / norm_lib_size_log (scarf/assay.py:41-64) executed from the reference's source; the oracle
    restatement gives the same float64 bits, for the renormalised scalar and for nCounts."""
    import scipy.sparse as sp

    c = sp.csr_matrix(ref_fn["norm_counts"])
    ci, fi = ref_fn["norm_cell_idx"], ref_fn["norm_feat_idx"]
    assert np.array_equal(P.normed_hvg(c, ci, fi), ref_fn["norm_lib_size_log"])
    assert np.array_equal(P.normed_hvg(c, ci, fi, log_transform=False), ref_fn["norm_lib_size"])
    assert np.array_equal(P.normed_hvg(c, ci, fi, renormalize_subset=False, n_counts=ref_fn["norm_n_counts"]),
                          ref_fn["norm_lib_size_log_ncounts"])
    assert (ref_fn["norm_lib_size_log"][list(ci).index(7)] == 0).all()  # the empty cell: scalar 0 -> 1, values 0


def test_clean_array_equals_reference_function(ref_fn):
    """a6: clean_array (scarf/utils.py:143-153): oracle and the product's tensor version."""
    import torch

    from scarf_b200 import graph

    for fill, key in ((0, "clean_fill0"), (1, "clean_fill1")):
        assert np.array_equal(P.clean_array(ref_fn["clean_in"].copy(), fill), ref_fn[key])
        got = graph.clean_array(torch.from_numpy(ref_fn["clean_in"].copy()), float(fill)).numpy()
        assert np.array_equal(got, ref_fn[key])


def test_fix_knn_query_equals_reference_function(ref_fn):
    """a10: fix_knn_query (scarf/ann.py:31-52) on rows with the self hit first, further down, and missing."""
    from scarf_b200.graph import fix_knn_query

    i, d, n_mis = fix_knn_query(ref_fn["fix_ind"], ref_fn["fix_dist"], ref_fn["fix_ref_idx"])
    assert np.array_equal(i, ref_fn["fix_out_ind"]) and np.array_equal(d, ref_fn["fix_out_dist"])
    assert n_mis == int(ref_fn["fix_n_mis"]) == 20


def test_order_features_equals_reference_function(ref_fn):
    """a14: _order_features (scarf/mapping_utils.py:98-145, defaults): oracle and the product's host routine."""
    from scarf_b200.graph import order_features

    s_ids, t_ids, want = ref_fn["order_s_ids"], ref_fn["order_t_ids"], ref_fn["order_t_re_idx"]
    s_idx, t_re = P.order_features(s_ids, t_ids, ref_fn["order_s_feat_ids"])
    assert np.array_equal(s_idx, ref_fn["order_s_idx"]) and np.array_equal(t_re, want)
    assert np.array_equal(order_features(s_ids, t_ids, ref_fn["order_s_idx"]), want) and (want == -1).sum() == 7
    with pytest.raises(ValueError, match="None of the features"):
        order_features(s_ids, np.array(["Q1", "Q2"]), ref_fn["order_s_idx"])


def test_fit_lowess_equals_reference_function(ref_fn):
    """a3: fit_lowess (scarf/feat_utils.py:11-45) executed from the reference's source around the restated smoother:
    histogram edges, bin membership, first-minimum gene of every bin and exp(log var - fit) agree with the oracle
    restatement to 2 ulp (the reference raises e to a scalar at a time, numpy's vectorised pow differs in the last bit
    for 5 % of the genes), and with the product's host routine (native LOWESS, other summation order) to 1e-10."""
    from scarf_b200 import hvg

    a, b, want = ref_fn["lowess_avg"], ref_fn["lowess_var"], ref_fn["lowess_c_var"]
    np.testing.assert_allclose(P.fit_lowess(a, b, 200, 0.1), want, rtol=5e-16, atol=0)
    np.testing.assert_allclose(hvg.fit_lowess(a, b, 200, 0.1), want, rtol=1e-10)
    assert np.isfinite(want).all() and want.min() > 0


@pytest.mark.parametrize("case", [0, 1, 2, 3, 4])
def test_hvg_choice_equals_reference_method(ref_fn, case):
    """a3: RNAassay.mark_hvgs (scarf/assay.py:945-1063) executed on a stub assay through the reference's own
    MetaData.sift / multi_sift / grep / get_index_by: default top-n rule, log2 mean bounds with max_cells, top_n larger
    than the number of eligible genes, explicit min_var / max_var (top_n then ignored), keep_bounds=True with bounds
    placed on values of the data (61 genes for top_n = 60: the threshold gene is kept).  Gene names include lower-case
    spellings the blacklist still catches (names and pattern are upper-cased, `re.match`).  Oracle, product host
    routine and product tensor routine give the same mask."""
    import torch

    from scarf_b200 import hvg

    top_n, min_cells, max_cells, min_mean, max_mean, min_var, max_var, keep_bounds = ref_fn[f"hvg_case{case}_params"]
    want = ref_fn[f"hvg_case{case}_mask"]
    names, feat_I = ref_fn["hvg_names"], ref_fn["hvg_feat_I"]
    nn, nz, cv = ref_fn["hvg_normed_n"], ref_fn["hvg_nz_mean"], ref_fn["hvg_c_var"]
    bl = str(ref_fn["hvg_blacklist"])
    kw = dict(top_n=int(top_n), min_cells=min_cells, max_cells=max_cells, min_mean=min_mean, max_mean=max_mean,
              min_var=min_var, max_var=max_var, keep_bounds=bool(keep_bounds))
    assert want.sum() == [100, 40, 997, 565, 61][case]
    assert np.array_equal(P.choose_hvgs(nn, nz, cv, feat_I, names, blacklist=bl, **kw), want)
    assert np.array_equal(hvg.choose_hvgs(nn, nz, cv, feat_I, names, blacklist=bl, **kw), want)
    keep = hvg.blacklist_keep_mask(names, len(names), bl)
    assert (~keep).sum() == 13  # every planted name but "xMT-1" (the pattern is anchored at the start)
    t = lambda a: torch.from_numpy(np.asarray(a))
    got = hvg.choose_hvgs_device(t(nn), t(nz), t(cv), t(feat_I & keep), **kw).numpy()
    assert np.array_equal(got, want)
    with pytest.raises(ValueError, match="greater than 0"):
        hvg.choose_hvgs(nn, nz, cv, feat_I, names, top_n=0)


def test_ipca_loop_equals_reference_method(ref_fn):
    """a8: AnnStream._fit_pca (scarf/ann.py:207-256) executed on a stub stream of numpy blocks -- first block held back
    as the end reservoir, a last block shorter than dims + 1 carried over, dims + 1 components fitted and the last one
    dropped -- against the oracle's restatement of the loop (same scikit-learn underneath)."""
    got = P.ipca_loadings(ref_fn["pca_x"], ref_fn["pca_mu"], ref_fn["pca_sigma"], int(ref_fn["pca_dims"]),
                          int(ref_fn["pca_batch"]))
    assert got.shape == ref_fn["pca_loadings"].shape == (40, 7)
    np.testing.assert_allclose(got, ref_fn["pca_loadings"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag,kw", [("default", {}), ("sym_upper", dict(symmetric=True, upper_only=True)),
                                    ("sym_full", dict(symmetric=True, upper_only=False)),
                                    ("raw_k3", dict(symmetric=False, use_k=3)), ("sym_k0", dict(symmetric=True, use_k=0)),
                                    ("sym_upper_k9", dict(symmetric=True, upper_only=True, use_k=9))])
def test_load_graph_equals_reference_method(ref_fn, tag, kw):
    """f-4: GraphDataStore._store_to_sparse + load_graph (scarf/datastore/graph_datastore.py:474-511,1022-1075)
    executed on a stub store: the defaults return the stored directed graph, use_k is clamped into [1, k] and keeps the
    first use_k entries of every row, symmetric=True gives g + g.T - g * g.T, upper_only its upper triangle."""
    from scarf_b200.graph import graph_to_sparse

    g = graph_to_sparse(ref_fn["graph_edges"], ref_fn["graph_weights"], int(ref_fn["graph_n"]), int(ref_fn["graph_k"]), **kw)
    assert np.array_equal(np.asarray(g.todense()), ref_fn[f"graph_{tag}"])


def test_smoothen_dists_loop_equals_reference_function(ref_fn):
    """a11: smoothen_dists (scarf/knn_utils.py:89-159) executed around the restated umap functions on a stub store:
    three chunks (the last ragged), chunk-local 'neighbour id == row id' zeroing that hits non-self entries in later
    chunks, the floor at the smallest non-zero weight, zero distances.  The oracle's loop gives the same arrays."""
    e, w = P.smoothen_dists(ref_fn["smooth_idx"], ref_fn["smooth_dist"], 1.0, 1.5, int(ref_fn["smooth_chunk"]))
    assert np.array_equal(e, ref_fn["smooth_edges"]) and np.array_equal(w, ref_fn["smooth_weights"])
    wr = ref_fn["smooth_weights"].reshape(-1, 6)
    floor = wr.min()
    assert wr[1007, 4] == floor and wr[2100, 0] == floor and wr[3, 2] == floor  # the planted chunk-local collisions
    assert (wr == floor).sum() == 9 and (wr == 0).sum() == 0  # + chance collisions (e.g. row 1017 -> cell 17) + the floor's source
