"""The Scarf-compatible front end (scarf_b200/datastore.py): Zarr layout, parameter resolution, error behaviour.
The store logic is tested on CPU; everything that computes needs the GPU."""
import json
import os

import numpy as np
import pytest


def test_zarr_store_roundtrip_and_metadata(tmp_path):
    from scarf_b200.zarr_store import open_group

    root = open_group(str(tmp_path / "s.zarr"), "w")
    g = root.create_group("RNA/normed__I__hvgs/reduction__pca__11__I")
    assert "RNA" in root and "RNA/normed__I__hvgs" in root
    a = g.create_dataset("reduction", (2500, 7), "f8", (1000,))
    assert a.chunks == (1000, 7)  # a 1-tuple of chunks is completed to full width like zarr does
    x = np.arange(2500 * 7, dtype=np.float64).reshape(2500, 7)
    a[:] = x
    a[1200:1300] = -x[1200:1300]
    x[1200:1300] *= -1
    b = root["RNA/normed__I__hvgs/reduction__pca__11__I"]["reduction"]
    assert np.array_equal(b[:], x) and np.array_equal(b[990:2010], x[990:2010])
    meta = json.load(open(os.path.join(a.path, ".zarray")))
    assert meta == {"chunks": [1000, 7], "compressor": None, "dtype": "<f8", "fill_value": 0, "filters": None,
                    "order": "C", "shape": [2500, 7], "zarr_format": 2}
    assert sorted(os.listdir(a.path)) == [".zarray", "0.0", "1.0", "2.0"]
    e = g.create_dataset("edges", (33, 2), "u8", (22,))
    e[:] = np.arange(66).reshape(33, 2)
    assert e[:].dtype == np.uint64 and e[30:33].tolist() == [[60, 61], [62, 63], [64, 65]]
    s = g.create_dataset("ids", (3,), "U5", (100000,))
    s[:] = np.array(["a", "bcd", "efghi"])
    assert root["RNA/normed__I__hvgs/reduction__pca__11__I/ids"][:].tolist() == ["a", "bcd", "efghi"]
    g.attrs["latest_ann"] = "x/y"
    g.attrs["subset_params"] = {"log_transform": True}
    assert root["RNA/normed__I__hvgs/reduction__pca__11__I"].attrs["latest_ann"] == "x/y"
    assert json.load(open(os.path.join(g.path, ".zgroup"))) == {"zarr_format": 2}


def test_fix_knn_query_matches_reference_semantics():
    from scarf_b200.graph import fix_knn_query

    ind = np.array([[0, 5, 6, 7], [9, 1, 8, 7], [4, 5, 6, 7], [3, 3, 2, 1]])
    dist = np.arange(16, dtype=np.float32).reshape(4, 4)
    i, d, n_mis = fix_knn_query(ind, dist, np.array([0, 1, 2, 3]))
    assert i.tolist() == [[5, 6, 7], [9, 8, 7], [4, 5, 6], [3, 2, 1]]
    assert d.tolist() == [[1, 2, 3], [4, 6, 7], [8, 9, 10], [13, 14, 15]]
    assert n_mis == 2


@pytest.mark.gpu
def test_datastore_chain_layout_and_cache(tmp_path, pbmc):
    """mark_hvgs -> make_graph -> load_graph -> run_mapping on the PBMC fixture: Scarf's Zarr tree (names, dtypes,
    chunks, latest_* attributes), the reference goldens through the front end, re-run = cache hit, cached-parameter
    resolution, error behaviour."""
    import scipy.sparse as sp

    from oracle import pipeline as P
    from scarf_b200.datastore import DataStore

    counts = pbmc["counts"]
    ids = np.array([f"ENSG{i:08d}" for i in range(counts.shape[1])])
    ds = DataStore.from_csr(str(tmp_path / "pbmc.zarr"), counts, ids, feature_names=pbmc["names"])
    assert np.array_equal(ds.cells.fetch_all("RNA_nCounts"), np.asarray(counts.sum(1)).ravel())
    # the reference fixture filters cells with auto_filter_cells (out of scope): install its kept set as `I`
    keep = np.zeros(892, dtype=bool)
    keep[pbmc["cell_idx"]] = True
    a = ds.cells.z.create_dataset("I", keep.shape, bool, (100000,))
    a[:] = keep
    with pytest.raises(ValueError, match="You have to choose which features"):
        ds.make_graph()
    ds.mark_hvgs(top_n=100)
    hv = ds.RNA.feats.fetch_all("I__hvgs")
    feat_I = P.gene_ncells(counts) > 20
    assert np.array_equal(hv, P.mark_hvgs(counts, pbmc["cell_idx"], feat_I, gene_names=pbmc["names"], top_n=100))
    with pytest.raises(NotImplementedError):
        ds.make_graph(feat_key="hvgs", harmonize=True, batch_columns=["x"])
    ds.make_graph(feat_key="hvgs")
    base = "RNA/normed__I__hvgs"
    red = f"{base}/reduction__pca__11__I"
    ann = f"{red}/ann__l2__50__50__48__4466"
    knn = f"{ann}/knn__11"
    gl = f"{knn}/graph__1.0__1.5"
    z = ds.zw
    assert z[base].attrs["latest_reduction"] == red and z[red].attrs["latest_ann"] == ann
    assert z[red].attrs["latest_kmeans"] == f"{red}/kmeans__1000__4466"
    assert z[ann].attrs["latest_knn"] == knn and z[knn].attrs["latest_graph"] == gl
    assert z[ann].attrs["isHarmonized"] is False
    assert z["RNA"].attrs["latest_feat_key"] == "hvgs" and z["RNA"].attrs["latest_cell_key"] == "I"
    assert z[base].attrs["subset_params"] == {"log_transform": True, "renormalize_subset": True}
    for loc, name, shape, dtype, chunks in (
            (base, "mu", (100,), "<f8", (100000,)), (base, "sigma", (100,), "<f8", (100000,)),
            (red, "reduction", (100, 11), "<f8", (1000, 100)),  # chunks = data.chunksize (graph_datastore.py:921-928)
             (knn, "indices", (808, 11), "<u8", (1000, 11)),
            (knn, "distances", (808, 11), "<f8", (1000, 11)), (gl, "edges", (8888, 2), "<u8", (11000, 2)),
            (gl, "weights", (8888,), "<f8", (11000,)),
            (f"{red}/kmeans__1000__4466", "cluster_centers", (808, 11), "<f8", (1000, 1000)),
            (f"{red}/kmeans__1000__4466", "cluster_labels", (808,), "<f8", (100000,))):
        arr = z[loc][name]
        assert (arr.shape, arr.dtype.str, arr.chunks) == (shape, dtype, chunks), (loc, name)
    idx = z[knn]["indices"][:]
    recall = np.mean([len(set(a_) & set(b_)) / 11 for a_, b_ in zip(idx.astype(np.int64), pbmc["indices"].astype(np.int64))])
    assert recall > 0.99
    edges, weights = z[gl]["edges"][:], z[gl]["weights"][:]
    assert np.array_equal(edges[:, 0], np.repeat(np.arange(808, dtype=np.uint64), 11))
    assert np.array_equal(edges[:, 1].reshape(808, 11), idx)
    _, w_o = P.smoothen_dists(idx, z[knn]["distances"][:], 1.0, 1.5, 1000)
    assert np.abs(weights - w_o).max() < 1e-5
    g = ds.load_graph(symmetric=True, upper_only=True)
    assert sp.issparse(g) and g.shape == (808, 808) and (g - sp.triu(g)).nnz == 0
    assert ds.load_graph().nnz == 808 * 11  # the reference's defaults (None): the stored, directed graph
    # Assay.save_normalized_data: the materialised normalised matrix (row a5) against the oracle, 1e-5 relative
    loc = ds.save_normalized_data(feat_key="hvgs")
    arr = z[loc.rsplit("/", 1)[0]]["data"]
    data = arr[:]
    x_o = P.normed_hvg(counts, pbmc["cell_idx"], np.where(hv)[0])
    assert loc == f"{base}/data" and data.shape == x_o.shape and arr.chunks == (1000, 100) and arr.dtype.str == "<f8"
    np.testing.assert_allclose(data, x_o, rtol=1e-5, atol=1e-7)
    g5 = ds.load_graph(symmetric=False, use_k=5)
    assert g5.nnz == 808 * 5
    # second call: cache hit (nothing rewritten), cached k is picked up when k is not given
    stamp = os.path.getmtime(os.path.join(z[knn]["indices"].path, "0.0"))
    ds.make_graph(feat_key="hvgs")
    assert os.path.getmtime(os.path.join(z[knn]["indices"].path, "0.0")) == stamp
    ds.make_graph(feat_key="hvgs", k=7)
    assert z[ann].attrs["latest_knn"] == f"{ann}/knn__7" and z[f"{ann}/knn__7"]["indices"].shape == (808, 7)
    ds.make_graph(feat_key="hvgs")  # k=None now resolves to the cached 7
    assert z[ann].attrs["latest_knn"] == f"{ann}/knn__7"
    ds.make_graph(feat_key="hvgs", k=11)
    # run_mapping: the self-map golden
    with pytest.raises(ValueError, match="cannot be sample as"):
        ds.run_mapping(target_assay=ds.RNA, target_name="selfmap", target_feat_key="hvgs")
    ds.run_mapping(target_assay=ds.RNA, target_name="selfmap", target_feat_key="hvgs_self", save_k=3)
    pi, pd_ = z["RNA/projections/selfmap"]["indices"], z["RNA/projections/selfmap"]["distances"]
    assert (pi.shape, pi.dtype.str, pi.chunks, pd_.dtype.str) == ((808, 3), "<u8", (1000, 3), "<f8")
    sc = P.mapping_score(pi[:], pd_[:], 808, per_k=False)
    assert (np.abs(sc - pbmc["mapping_scores"]) < 1e-2).mean() > 0.98
    # run_mapping finds the finished graph in the store (mapping_datastore.py:143 -> the cached branches of make_graph,
    # graph_datastore.py:797-881): loadings / mu / sigma / embedding are loaded, nothing is recomputed or rewritten
    stamp = os.path.getmtime(os.path.join(z[knn]["indices"].path, "0.0"))
    ds.run_mapping(target_assay=ds.RNA, target_name="selfmap2", target_feat_key="hvgs_self", save_k=3)
    assert ds.last_make_graph_timing == {"cache_hit": True}
    assert os.path.getmtime(os.path.join(z[knn]["indices"].path, "0.0")) == stamp
    assert np.array_equal(z["RNA/projections/selfmap2"]["indices"][:], pi[:])
    assert np.array_equal(z["RNA/projections/selfmap2"]["distances"][:], pd_[:])
    # the subset hash is the reference's (an int, assay.py:317-329); a different cell subset makes the group stale
    from scarf_b200.datastore import create_subset_hash
    assert z[base].attrs["subset_hash"] == create_subset_hash(pbmc["cell_idx"], np.where(hv)[0])
    ann_obj = ds.make_graph(feat_key="hvgs", return_ann_object=True)
    assert ann_obj.k == 11 and ann_obj.loadings.shape == (100, 11) and ann_obj.kmeans.cluster_centers_.shape == (808, 11)
    one = ann_obj.reducer(np.zeros(100))
    assert one.shape == (11,)
    ki, kd = ann_obj.transform_ann(ann_obj.reducer(np.zeros((4, 100))), k=3)
    assert ki.shape == (4, 3) and ki.dtype == np.uint64
