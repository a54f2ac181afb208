"""Developer tool (GPU): runs the normalise -> Gram part of one bench workload and saves the covariance handed to the
eigensolver (float64 .npy) so that the solver's schedule can be studied off-line (tools/eig_schedule_prototype.py,
tools/eig_probe.py).  usage: python tools/dump_cov.py C3 gpurun_out/c3_cov.npy [cells]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from scarf_b200 import graph, lib, ops, synth  # noqa: E402


class Stop(Exception):
    pass


def main():
    name, out = sys.argv[1], sys.argv[2]
    cfg = dict(bench.WORKLOADS[name])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["cells"]
    dev = torch.device("cuda:0")
    csr = synth.make_counts_device(n, cfg["genes"], cfg["factors"], seed=bench.SEED, device=dev, block=bench.GEN_BLOCK,
                                   row_start=0)
    n_counts, _ = graph.cell_totals(csr)
    feat_i = graph.gene_ncells(csr) > 20
    keep = torch.ones(cfg["genes"], dtype=torch.bool, device=dev)
    hv = graph.mark_hvgs_csr(csr, None, feat_i, n_counts, n, top_n=cfg["hvgs"], as_tensor=True, keep_mask=keep)
    real = ops.eig_topk

    def grab(g_fx, n_cols, dims, scale, *a, **k):
        cov = g_fx[:n_cols, :n_cols].to(torch.float64) * scale
        np.save(out, cov.cpu().numpy())
        print(f"saved {out}: {tuple(cov.shape)}, dims {dims}, scale {scale:.6e}, n {n}", flush=True)
        raise Stop()

    ops.eig_topk = grab
    try:
        graph.make_graph_csr(csr, None, hv, dims=cfg["dims"], k=cfg["k"], gram_mode=3, knn_method=1)
    except Stop:
        pass
    ops.eig_topk = real


if __name__ == "__main__":
    main()
