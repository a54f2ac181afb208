"""BASELINE.json configs[4] (C5): run_mapping of 500k synthetic query cells onto the 1M-cell reference graph on one GPU.
Reference: C3 (1M cells x 30k genes, 2k HVGs, D = 100, k = 21); queries: further cells of the same generative model
(same gene model, disjoint generator blocks).  Times the mapping core (row sums, normalise with the reference's
mu / sigma, tensor-core projection, exact kNN without self handling) with CUDA events for save_k = 3 (the reference
default, scarf/datastore/mapping_datastore.py:40) and 11, and checks a sample of query rows against an FP64 brute
force over all 1M reference cells (indices and float32 distances bit for bit).

    python tools/c5_mapping_probe.py [n_ref n_query]   ->  one JSON line
"""
import json
import sys
import time

import numpy as np
import torch

from scarf_b200 import graph, synth

n_ref, n_q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000, 500_000)
genes, hvgs, dims, k, factors, block = 30_000, 2_000, 100, 21, 115, 2000
dev = torch.device("cuda:0")
t0 = time.time()
ref = synth.make_counts_device(n_ref, genes, factors, seed=4466, device=dev, block=block)
q_start = (n_ref + block - 1) // block * block
qry = synth.make_counts_device(n_q, genes, factors, seed=4466, device=dev, block=block, row_start=q_start)
torch.cuda.synchronize()
t_gen = time.time() - t0

n_counts, _ = graph.cell_totals(ref)
feat_I = graph.gene_ncells(ref) > 20
hv = graph.mark_hvgs_csr(ref, None, feat_I, n_counts, n_ref, top_n=hvgs, as_tensor=True,
                         keep_mask=torch.ones(genes, dtype=torch.bool, device=dev))
res = graph.make_graph_csr(ref, None, hv, dims=dims, k=k)
torch.cuda.synchronize()
feat_idx = np.asarray(res.feat_idx)


def mapping(save_k):
    return graph.run_mapping_csr(qry, None, feat_idx, res.mu, res.sigma, res.loadings, res.embedding_all, res.dims,
                                 save_k=save_k)


out = {"workload": f"C5: {n_q} queries -> {n_ref}-cell reference, {genes} genes, H={len(feat_idx)}, D={res.dims}",
       "synth_s": round(t_gen, 1), "query_nnz": qry.nnz}
for save_k in (3, 11):
    mapping(save_k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m = mapping(save_k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    out[f"k{save_k}"] = {"ms": round(ms, 1), "queries_per_s": round(n_q / ms * 1e3),
                         "knn_tflops_algorithmic": round(2.0 * n_q * n_ref * res.dims / ms / 1e9, 1)}
    # sample check against the oracle's definition: float32(sum in float64 of (a - b)^2), order (distance, index)
    rows = torch.randint(0, n_q, (48,), generator=torch.Generator().manual_seed(save_k)).to(dev)
    a = m.embedding[rows, : res.dims].double()
    b = res.embedding_all[:, : res.dims]
    d = torch.empty((rows.numel(), n_ref), dtype=torch.float32, device=dev)
    for lo in range(0, n_ref, 100_000):
        bb = b[lo: lo + 100_000].double()
        d[:, lo: lo + 100_000] = ((a[:, None, :] - bb[None, :, :]) ** 2).sum(-1).float()
    order = torch.argsort(d, dim=1, stable=True)[:, :save_k]  # stable: ties keep the lower index first
    want_d = torch.gather(d, 1, order)
    ok = bool(torch.equal(order, m.indices[rows])) and bool(torch.equal(want_d, m.distances[rows]))
    out[f"k{save_k}"]["sample_rows_bit_exact"] = ok
    del d, order, want_d
print(json.dumps(out))
