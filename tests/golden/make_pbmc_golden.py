"""Generates tests/golden/pbmc_*.np[yz] from the reference's own test fixtures.

Run in the build container (needs /root/reference; the GPU box never runs this):
    PYTHONPATH=/root/repo python tests/golden/make_pbmc_golden.py

Inputs (reference, read-only):
  scarf/tests/datasets/1K_pbmc_citeseq.zarr.tar.gz   raw counts 892 x 36601 u32 (Blosc lz4 bitshuffle)
  scarf/tests/datasets/cell_attributes.csv           index = the 808 cells kept by auto_filter_cells
  scarf/tests/datasets/knn_{indices,distances,weights}.npy   goldens of test_datastore.py:66-79
Outputs:
  pbmc_counts.npz   CSR of the raw counts (+ gene names, kept cell ids)
  pbmc_knn_indices.npy / pbmc_knn_distances.npy / pbmc_knn_weights.npy   copies of the goldens
  pbmc_mapping_scores.npy   cell_attributes.csv column mapping_scores (run_mapping self-map, save_k=3)
"""
import os
import tarfile
import tempfile

import numpy as np
import pandas as pd
import scipy.sparse as sp

from oracle.blosc_shim import read_zarr_array

REF = "/root/reference/scarf/tests/datasets"
OUT = os.path.dirname(os.path.abspath(__file__))

with tempfile.TemporaryDirectory() as tmp:
    tarfile.open(os.path.join(REF, "1K_pbmc_citeseq.zarr.tar.gz"), "r:gz").extractall(tmp)
    counts = sp.csr_matrix(read_zarr_array(os.path.join(tmp, "RNA", "counts")))
    names = read_zarr_array(os.path.join(tmp, "RNA", "featureData", "names"))
attrs = pd.read_csv(os.path.join(REF, "cell_attributes.csv"), index_col=0)
assert np.array_equal(np.asarray(counts.sum(1)).ravel()[attrs.index.values], attrs.RNA_nCounts.values)
assert counts.data.max() < 65536
np.savez_compressed(os.path.join(OUT, "pbmc_counts.npz"), indptr=counts.indptr.astype(np.int64),
                    indices=counts.indices.astype(np.int32), data=counts.data.astype(np.uint16),
                    shape=np.array(counts.shape), names=names.astype("U"), cell_idx=attrs.index.values.astype(np.int64))
for f in ("indices", "distances", "weights"):
    np.save(os.path.join(OUT, f"pbmc_knn_{f}.npy"), np.load(os.path.join(REF, f"knn_{f}.npy")))
# run_mapping golden (test_datastore.py:160-163): self-mapping scores of the 808 cells
np.save(os.path.join(OUT, "pbmc_mapping_scores.npy"), attrs.mapping_scores.values.astype(np.float64))
print("ok", counts.shape, counts.nnz)
