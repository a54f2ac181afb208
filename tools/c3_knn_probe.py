"""Developer probe (GPU): kNN stage on the real C3 embedding (1M cells, D=100, k=21): time, guard failures."""
import sys
import torch
sys.path.insert(0, ".")
from scarf_b200 import graph, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
csr = synth.make_counts_device(n, 30_000, 115, seed=4466, device=dev, block=2000)
n_counts, _ = graph.cell_totals(csr)
feat_I = graph.gene_ncells(csr) > 20
hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True,
                         keep_mask=torch.ones(30_000, dtype=torch.bool, device=dev))
_orig = graph.eig_topk
_st = {}
graph.eig_topk = lambda cov, dims, **kw: _orig(cov, dims, stats=_st)
res = graph.make_graph_csr(csr, None, hv, dims=100, k=21, gram_mode=3, knn_method=1)
print("eig stats", _st)
del csr
y = res.embedding_all
print("evals", [round(float(x), 3) for x in res.eigenvalues[[0, 1, 10, 30, 50, 60, 70, 80, 99]]])
print("|y|^2 mean", float((y * y).sum(1).mean()), "d_k mean", float(res.distances[:, -1].mean()), "d_1 mean", float(res.distances[:, 0].mean()))
for rep in range(2):
    st = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, dist = ops.knn_l2(y, y, 100, 21, self_offset=0, method=1, stats=st)
    e1.record()
    torch.cuda.synchronize()
    print("knn ms", e0.elapsed_time(e1), "guard fails", int(st["guard_fail_rows"].item()))
