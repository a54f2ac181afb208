"""Pins the CPU oracle on the reference's own golden vectors (scarf/tests/test_datastore.py:66-79).

The goldens were produced by the reference chain auto_filter_cells -> mark_hvgs(top_n=100) ->
make_graph(feat_key='hvgs') (scarf/tests/fixtures_datastore.py:58-73) on the PBMC fixture.
hnswlib is approximate, so the index pin is a stated recall / identical-row fraction.
"""
import numpy as np
import pytest

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
