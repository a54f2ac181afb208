"""Developer probe (GPU): where the time of mark_hvgs_csr goes (wall clock with synchronisation per section)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import graph, hvg, ops, synth  # noqa: E402

n = 100000
csr = synth.make_counts_device(n, 30000, 65, seed=4466, device="cuda", block=2000)
n_counts, _ = graph.cell_totals(csr)
feat_I = graph.gene_ncells(csr) > 20
keep = torch.ones(30000, dtype=torch.bool, device="cuda")


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:28s} {1e3 * (t1 - t0):7.3f} ms")
    return t1


for rep in range(3):
    print("rep", rep)
    torch.cuda.synchronize()
    t = time.perf_counter()
    st = graph.hvg_gene_stats(csr, None, n_counts, n, None, as_numpy=False)
    t = tick("hvg_gene_stats", t)
    a, s = st["avg"][feat_I], st["sigmas"][feat_I]
    t = tick("mask index", t)
    cv = hvg.remove_trend_device(a, s)
    t = tick("remove_trend_device", t)
    c_var = torch.full((30000,), float("nan"), dtype=torch.float64, device="cuda")
    c_var[feat_I] = cv
    t = tick("scatter c_var", t)
    m = hvg.choose_hvgs_device(st["normed_n"], st["nz_mean"], c_var, feat_I & keep, 2000, 1000)
    t = tick("choose_hvgs_device", t)
    t0 = time.perf_counter()
    hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True, keep_mask=keep)
    tick("mark_hvgs_csr total", t0)
    t0 = time.perf_counter()
    res = graph.make_graph_csr(csr, None, hv, dims=50, k=11, gram_mode=3, knn_method=1)
    tick("make_graph_csr total", t0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True, keep_mask=keep)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=50))
