"""Multi-GPU == single-GPU check (launched by tests/test_multigpu.py or by hand under torchrun):
every rank builds the same synthetic matrix, runs make_graph (+ mark_hvgs, run_mapping) once alone on its GPU and once
sharded over all ranks, and compares its shard of the sharded result bit for bit with the single-GPU rows."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scarf_b200 import graph, synth  # noqa: E402
from scarf_b200.dist import Comm, ShardPlan  # noqa: E402
from scarf_b200.ops import CsrDevice  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n, g, top_n, dims, k = int(os.environ.get("MG_CELLS", "23500")), 8000, 800, 30, 11
    full = synth.make_counts_device(n, g, 45, seed=21, device=dev, block=500)
    nq = 3000
    tgt = synth.make_counts_device(nq, g, 45, seed=22, device=dev, block=500)

    def run(csr, tcsr, comm, n_total):
        n_counts, _ = graph.cell_totals(csr)
        feat_I = graph.gene_ncells(csr, comm) > 20
        hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n_total, top_n=top_n, comm=comm, as_tensor=True)
        res = graph.make_graph_csr(csr, None, hv, dims=dims, k=k, comm=comm, gram_mode=3, knn_method=1)
        t_col = np.where(hv.cpu().numpy())[0]
        mp = graph.run_mapping_csr(tcsr, None, t_col, res.mu, res.sigma, res.loadings, res.embedding_all, res.dims,
                                   save_k=3, comm=comm)
        return hv, res, mp

    hv1, r1, m1 = run(full, tgt, None, n)
    td.init_process_group("nccl", device_id=dev)
    comm = Comm()
    plan, qplan = ShardPlan.make(n, world, 1000), ShardPlan.make(nq, world, 1000)
    (a, b), (qa, qb) = plan.rows(rank), qplan.rows(rank)

    def shard(c, a, b):
        ip = c.indptr[a:b + 1]
        return CsrDevice((ip - ip[0]).contiguous(), c.indices[ip[0]:ip[-1]].contiguous(),
                         c.data[ip[0]:ip[-1]].contiguous(), b - a, c.n_cols)

    hvp, rp, mp = run(shard(full, a, b), shard(tgt, qa, qb), comm, n)
    ok = True

    def check(name, x, y):
        nonlocal ok
        same = bool(torch.equal(x, y))
        ok &= same
        if not same:
            d = (x.double() - y.double()).abs().max().item() if x.shape == y.shape else "shape"
            print(f"[rank {rank}] MISMATCH {name}: max |diff| {d}", flush=True)

    check("hvgs", hv1, hvp)
    check("mu", r1.mu, rp.mu), check("sigma", r1.sigma, rp.sigma)
    check("loadings", r1.loadings, rp.loadings)
    check("embedding", r1.embedding[a:b], rp.embedding)
    check("embedding_all", r1.embedding_all, rp.embedding_all)
    check("indices", r1.indices[a:b], rp.indices), check("distances", r1.distances[a:b], rp.distances)
    check("edges", r1.edges[a * k:b * k], rp.edges), check("weights", r1.weights[a * k:b * k], rp.weights)
    check("map indices", m1.indices[qa:qb], mp.indices), check("map distances", m1.distances[qa:qb], mp.distances)
    flag = torch.tensor([1 if ok else 0], device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_OK" if int(flag.item()) == 1 else "MULTIGPU_FAIL", f"world={world} cells={n}", flush=True)
    td.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
