"""Multi-GPU == single-GPU check (launched by tests/test_multigpu.py or by hand under torchrun):
every rank builds the same synthetic matrix, runs make_graph (+ mark_hvgs, run_mapping) once alone on its GPU and once
sharded over all ranks, and compares its shard of the sharded result bit for bit with the single-GPU rows.
MG_ONE_GPU=1: all ranks share cuda:0 and talk over gloo (host-staged collectives) -- the same sharded code path, the
same kernels, on a box with a single GPU.  Then the Scarf-style front end: the sharded DataStore (every rank loads its
rows, writes its chunks) must leave the same store as a single-rank one."""
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
    one_gpu = os.environ.get("MG_ONE_GPU") == "1"
    if one_gpu:
        local = 0
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
    if one_gpu:
        td.init_process_group("gloo")
    else:
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
    ok &= datastore_check(full, tgt, hv1, r1, m1, comm, dev, rank, top_n, dims, k)
    flag = torch.tensor([1 if ok else 0], device=dev)
    comm.allreduce_min_(flag)
    if rank == 0:
        print("MULTIGPU_OK" if int(flag.item()) == 1 else "MULTIGPU_FAIL", f"world={world} cells={n}", flush=True)
    td.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def datastore_check(full, tgt, hv1, r1, m1, comm, dev, rank, top_n, dims, k):
    """DataStore(comm=...) on a store in a directory all ranks see: mark_hvgs + make_graph + run_mapping sharded; rank 0
    compares the arrays in the store with the single-rank device results (widened like the store widens them)."""
    import shutil
    import tempfile

    from scarf_b200.datastore import DataStore

    path = os.path.join(tempfile.gettempdir(), "scarf_b200_mg_store.zarr")
    tpath = os.path.join(tempfile.gettempdir(), "scarf_b200_mg_target.zarr")
    g = full.n_cols
    ids = [f"g{i}" for i in range(g)]
    if rank == 0:
        for p_, c_ in ((path, full), (tpath, tgt)):
            shutil.rmtree(p_, ignore_errors=True)
            DataStore.from_csr(p_, synth.to_scipy(c_), ids, device=dev, min_features_per_cell=0)
    comm.barrier()
    ds = DataStore(path, device=dev, comm=comm, min_features_per_cell=0)
    ds.mark_hvgs(top_n=top_n, min_cells=None, show_plot=False)
    ds.make_graph(feat_key="hvgs", dims=dims, k=k, n_centroids=50)
    tds = DataStore(tpath, device=dev, comm=comm, min_features_per_cell=0)
    ds.run_mapping(tds.RNA, "tgt", "hvgs_tgt", save_k=3)
    comm.barrier()
    ok = True
    if rank == 0:
        z = ds.zw
        base = f"RNA/normed__I__hvgs/reduction__pca__{dims}__I/ann__l2__50__50__48__4466"
        knn = f"{base}/knn__{k}"
        want = {f"{knn}/indices": r1.indices.cpu().numpy().astype("u8"),
                f"{knn}/distances": r1.distances.cpu().numpy().astype("f8"),
                f"{knn}/graph__1.0__1.5/edges": r1.edges.cpu().numpy().astype("u8"),
                f"{knn}/graph__1.0__1.5/weights": r1.weights.cpu().numpy().astype("f8"),
                f"{base}/embedding": r1.embedding_all[:, :dims].cpu().numpy(),
                "RNA/projections/tgt/indices": m1.indices.cpu().numpy().astype("u8"),
                "RNA/projections/tgt/distances": m1.distances.cpu().numpy().astype("f8")}
        import numpy as np
        for loc, w in want.items():
            grp, name = loc.rsplit("/", 1)
            got = z[grp][name][:]
            same = got.shape == w.shape and np.array_equal(got, w)
            ok &= same
            if not same:
                print(f"[datastore] MISMATCH {loc}: {got.shape} vs {w.shape}", flush=True)
        hv_store = ds.RNA.feats.fetch_all("I__hvgs")
        same = np.array_equal(hv_store, hv1.cpu().numpy())
        ok &= same
        if not same:
            print("[datastore] MISMATCH I__hvgs", flush=True)
        # a second make_graph finds its groups (cache hit) and run_mapping does not recompute the graph
        ds.make_graph(feat_key="hvgs", dims=dims, k=k, n_centroids=50)
    else:
        ds.make_graph(feat_key="hvgs", dims=dims, k=k, n_centroids=50)
    ok &= ds.last_make_graph_timing.get("cache_hit") is True
    comm.barrier()
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True), shutil.rmtree(tpath, ignore_errors=True)
    return ok


if __name__ == "__main__":
    main()
