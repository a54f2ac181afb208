import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pbmc():
    """The reference's PBMC test dataset as CSR + its golden kNN arrays (tests/golden/make_pbmc_golden.py)."""
    import scipy.sparse as sp

    d = np.load(os.path.join(GOLDEN, "pbmc_counts.npz"))
    counts = sp.csr_matrix((d["data"].astype(np.uint32), d["indices"], d["indptr"]), shape=tuple(d["shape"]))
    return {
        "counts": counts,
        "names": d["names"],
        "cell_idx": d["cell_idx"],
        "indices": np.load(os.path.join(GOLDEN, "pbmc_knn_indices.npy")),
        "distances": np.load(os.path.join(GOLDEN, "pbmc_knn_distances.npy")),
        "weights": np.load(os.path.join(GOLDEN, "pbmc_knn_weights.npy")),
        "mapping_scores": np.load(os.path.join(GOLDEN, "pbmc_mapping_scores.npy")),
    }
