"""Developer probe (GPU): run one C2 step and dump the covariance matrix handed to the eigensolver plus a sample of
the embedding / kNN distances to gpurun_out/ for offline (CPU) experiments."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import graph, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dev = torch.device("cuda", 0)
csr = synth.make_counts_device(n, 30_000, 65, seed=4466, device=dev, block=2000)
n_counts, _ = graph.cell_totals(csr)
feat_I = graph.gene_ncells(csr) > 20
hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True,
                         keep_mask=torch.ones(30_000, dtype=torch.bool, device=dev))
captured = {}
orig = graph.eig_topk


def spy(cov, dims, **kw):
    captured["cov"] = cov.clone()
    st = {}
    out = orig(cov, dims, stats=st)
    captured["stats"] = st
    return out


graph.eig_topk = spy
res = graph.make_graph_csr(csr, None, hv, dims=50, k=11, gram_mode=3, knn_method=1)
torch.cuda.synchronize()
print("eig stats", captured["stats"])
np.save("gpurun_out/c2_cov.npy", captured["cov"].cpu().numpy())
np.save("gpurun_out/c2_y_sample.npy", res.embedding[:20000, :50].cpu().numpy())
np.save("gpurun_out/c2_dist_sample.npy", res.distances[:20000].cpu().numpy())
np.save("gpurun_out/c2_evals.npy", res.eigenvalues.cpu().numpy())
y = res.embedding[:, :50]
print("|y|^2 mean", float((y * y).sum(1).mean()), "max", float((y * y).sum(1).max()))
print("d_k mean", float(res.distances[:, -1].mean()), "d_1 mean", float(res.distances[:, 0].mean()))
print("evals", res.eigenvalues[:5].tolist(), res.eigenvalues[-3:].tolist())
