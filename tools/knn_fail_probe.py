"""Developer probe (GPU): guard failures of the tensor-core kNN on the embedding of a bench workload -- rows that fail the
guard, rows left for the FP64 scan after the collect pass.  usage: python tools/knn_fail_probe.py C3 [cells]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from scarf_b200 import graph, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = dict(bench.WORKLOADS[name])
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["cells"]
dev = torch.device("cuda:0")
csr = synth.make_counts_device(n, cfg["genes"], cfg["factors"], seed=bench.SEED, device=dev, block=bench.GEN_BLOCK, row_start=0)
n_counts, _ = graph.cell_totals(csr)
feat_i = graph.gene_ncells(csr) > 20
keep = torch.ones(cfg["genes"], dtype=torch.bool, device=dev)
hv = graph.mark_hvgs_csr(csr, None, feat_i, n_counts, n, top_n=cfg["hvgs"], as_tensor=True, keep_mask=keep)
res = graph.make_graph_csr(csr, None, hv, dims=cfg["dims"], k=cfg["k"], gram_mode=3, knn_method=1)
y = res.embedding_all
for nq in (n, n // 8):
    st = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.knn_l2(y[:nq], y, res.dims, cfg["k"], self_offset=0, method=1, stats=st)
    e0.record()
    ops.knn_l2(y[:nq], y, res.dims, cfg["k"], self_offset=0, method=1, stats=st)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {nq} queries x {n} refs, dims {res.dims}, k {cfg['k']}: {e0.elapsed_time(e1):.2f} ms, guard fail rows "
          f"{int(st['guard_fail_rows'].item())}, rows left for the FP64 scan {int(st['guard_rest_rows'].item())}", flush=True)
