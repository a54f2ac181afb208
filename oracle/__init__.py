"""CPU oracle for the Scarf ``make_graph`` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or the timed CPU baseline), never as the thing shipped.  The product
path (``scarf_b200``) never imports this package and fails loudly if its CUDA
library is missing.

Parity pin status (see DESIGN.md "Oracle"):
  * edge weights (App. A.7)      -- pinned on the reference's own golden vectors
    ``knn_indices.npy`` / ``knn_distances.npy`` -> ``knn_weights.npy``.
  * whole chain (gene stats -> LOWESS -> HVG -> normalise -> mu/sigma ->
    IncrementalPCA in Scarf's block order -> exact kNN) -- pinned on the
    reference's PBMC fixture against ``knn_indices.npy`` / ``knn_distances.npy``
    (hnswlib is approximate, so the pin is a stated recall, not bit equality).
  * the exact-covariance PCA used as the tight PCA oracle is a different
    estimator from IncrementalPCA; its agreement with the
    IncrementalPCA path is measured and stated, not assumed.
"""
