"""Developer probe (GPU): run the C2-style pipeline once and report kNN guard statistics on the real embedding."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import graph, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 50
k = int(sys.argv[3]) if len(sys.argv) > 3 else 11
csr = synth.make_counts_device(n, 30000, 65, seed=4466, device="cuda", block=2000)
n_counts, _ = graph.cell_totals(csr)
feat_I = (graph.gene_ncells(csr) > 20).cpu().numpy()
t0 = time.perf_counter()
hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000)
torch.cuda.synchronize()
print("mark_hvgs wall ms", 1e3 * (time.perf_counter() - t0))
res = graph.make_graph_csr(csr, None, hv, dims=dims, k=k, knn_method=0 if n <= 20000 else 1)
y = res.embedding
nrm = y[:, :dims].norm(dim=1)
print("norms: median %.2f max %.2f" % (nrm.median().item(), nrm.max().item()))
print("eigenvalues", res.eigenvalues[:5].tolist(), "...", res.eigenvalues[-3:].tolist())
d = res.distances
print("kth dist: median %.2f min %.3f ; 1st dist median %.2f" % (d[:, -1].median().item(), d[:, -1].min().item(), d[:, 0].median().item()))
for method in (1,):
    st = {}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    i1, d1 = ops.knn_l2(y, y, dims, k, self_offset=0, method=method, stats=st)
    e1.record()
    torch.cuda.synchronize()
    print("method", method, "ms", e0.elapsed_time(e1), "guard fails", int(st["guard_fail_rows"].item()))
bm = nrm.max().item()
eps = (2.0 * (2.0 * 2**-11 + 2**-22) + 2**-13) * nrm * bm + (2**-14) * bm * bm
print("eps median %.3f max %.3f" % (eps.median().item(), eps.max().item()))
