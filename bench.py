#!/usr/bin/env python
"""make_graph benchmark (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm  (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's algorithm on host cores

One "step" = mark_hvgs + make_graph on one synthetic CSR batch: raw CSR -> weighted kNN graph.  As in the
reference, the per-cell nCounts and the per-gene nCells mask `I` are attributes the DataStore computed when the
store was created (scarf/datastore/base_datastore.py:324-401, scarf/assay.py:201-225): both arms receive them.
`value`  : cells/s with the CSR shard already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public API with HOST (pinned) CSR buffers, H2D + D2H inside the timed region.
`roofline`: the dominant kernel (exact kNN) timed live with CUDA events on its launch stream.
`cpu_baseline`: the CPU oracle (restated reference path, "port") on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on for one GPU
    "C2": dict(cells=100_000, genes=30_000, hvgs=2_000, dims=50, k=11, factors=65),
    "C1": dict(cells=5_000, genes=20_000, hvgs=2_000, dims=25, k=11, factors=40),
    "C3": dict(cells=1_000_000, genes=30_000, hvgs=2_000, dims=100, k=21, factors=115),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=None, help="cells per GPU (overrides the workload)")
    ap.add_argument("--gram-mode", type=int, default=int(os.environ.get("SCF_GRAM_MODE", "3")))
    ap.add_argument("--knn-method", type=int, default=int(os.environ.get("SCF_KNN_METHOD", "1")))
    ap.add_argument("--cpu-sample", type=int, default=4000, help="cells in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}


def measure_f16_peak(torch, dev):
    """cuBLAS FP16 dense GEMM 8192^3 (FP32 accumulate), best of 10 -- the method MEASURED_PEAKS.json uses for bf16;
    the kNN kernel issues tcgen05.mma kind::f16, which runs at the same rate: TFLOP/s."""
    a = torch.randn((8192, 8192), device=dev, dtype=torch.float16)
    b = torch.randn((8192, 8192), device=dev, dtype=torch.float16)
    for _ in range(3):
        a @ b
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full summary (profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_full_summary.json")
    if os.path.exists(p):
        d = json.load(open(p)).get(kernel_key)
        if d:
            return d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """Samples whose wall-clock stamp falls inside [t0, t1] (the timed region); all samples if none does."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        use = inside or rows
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if use:
            out = {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": max(r[2] for r in use),
                   "reasons": sorted(reasons), "samples": len(use),
                   "window": "timed region" if inside else "whole run (no sample fell inside the timed region)"}
        return out


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, sample, threads):
    """The reference's algorithm on host cores (oracle port): mark_hvgs + make_graph on `sample` cells of the
    workload (same generator / seed, first `sample` rows).  Returns (seconds, cells)."""
    import numpy as np
    from threadpoolctl import threadpool_limits

    from oracle import pipeline as P
    from scarf_b200 import synth

    m = synth.make_counts_scipy(sample, cfg["genes"], cfg["factors"], seed=4466, block=2000)
    P.exact_knn(np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float32), 1)  # builds/loads the C part untimed
    cell_idx = np.arange(sample)
    n_counts, _ = P.cell_totals(m)     # DataStore-creation attributes: outside the timed region on both arms
    feat_I = P.gene_ncells(m) > 20
    t0 = time.perf_counter()
    with threadpool_limits(limits=threads):
        hv = P.mark_hvgs(m, cell_idx, feat_I, top_n=min(cfg["hvgs"], int(feat_I.sum()) - 1), n_counts=n_counts)
        P.make_graph(m, cell_idx, hv, dims=cfg["dims"], k=cfg["k"], pca="ipca", knn_threads=threads)
    return time.perf_counter() - t0, sample


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(cfg, min(1000, args.cpu_sample), threads)
    times = []
    for _ in range(args.steps):
        dt, n = cpu_reference_run(cfg, args.cpu_sample, threads)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = args.cpu_sample / (ms / 1e3)
    sample = f"first {args.cpu_sample} cells of the workload (same generator/seed), full gene set"
    print(json.dumps({
        "impl": "reference", "metric": "make_graph_cells_per_s", "value": val, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cfg, max(args.gpus, 1)),
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (scarf 0.32.3) cannot be imported here (dask/zarr/hnswlib/umap-learn absent, no network): "
                "this arm times the CPU restatement of its algorithm (oracle/, IncrementalPCA in Scarf's block order + "
                "exact kNN + umap smoothing) on all host cores",
    }))


def workload_config(args, cfg, world):
    return {"workload": f"{args.workload}: synthetic {cfg['cells']} cells/GPU x {cfg['genes']} genes raw CSR, "
                        f"{cfg['hvgs']} HVGs, dims={cfg['dims']}, k={cfg['k']} (mark_hvgs + make_graph)",
            "cells_per_gpu": cfg["cells"], "total_cells": cfg["cells"] * world, "genes": cfg["genes"],
            "hvgs": cfg["hvgs"], "dims": cfg["dims"], "k": cfg["k"],
            "l2": "inputs larger than L2 (CSR shard >> 126 MB is re-read from HBM every step)",
            "gram_mode": args.gram_mode, "knn_method": args.knn_method}


# ---------------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as td

    from scarf_b200 import graph, lib, synth
    from scarf_b200.dist import Comm
    from scarf_b200.ops import CsrDevice

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    comm = Comm()
    n_local = cfg["cells"]
    n_total = n_local * world
    block = 2000
    csr = synth.make_counts_device(n_local, cfg["genes"], cfg["factors"], seed=4466, device=dev, block=block,
                                   row_start=rank * n_local)
    torch.cuda.synchronize()
    nnz = csr.nnz
    timers = []

    n_counts, _ = graph.cell_totals(csr)  # DataStore-creation attributes (see the module docstring)
    feat_I = graph.gene_ncells(csr, comm) > 20  # bool device tensor
    keep = torch.ones(cfg["genes"], dtype=torch.bool, device=dev)  # synthetic gene names never hit the blacklist
    torch.cuda.synchronize()

    eig_stats = {}

    def step(c, tm=None):
        if tm is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            tm.append(("step_start", e))
        hv = graph.mark_hvgs_csr(c, None, feat_I, n_counts, n_total, top_n=cfg["hvgs"], comm=comm, as_tensor=True,
                                 keep_mask=keep)
        return graph.make_graph_csr(c, None, hv, dims=cfg["dims"], k=cfg["k"], comm=comm, gram_mode=args.gram_mode,
                                    knn_method=args.knn_method, timers=tm, stats=eig_stats)

    def timed(fn, steps):
        comm.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        comm.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        comm.allreduce_max_(ms)
        return float(ms.item())

    res = None
    for _ in range(max(args.warmup, 1)):
        res = step(csr)
    torch.cuda.synchronize()
    # stored values that fall on the selected features (the compact matrix the normalise stage reads): counted once,
    # outside the timed region, for the HBM-side roofline of that stage
    fmask = torch.zeros(cfg["genes"], dtype=torch.bool, device=dev)
    fmask[torch.from_numpy(res.feat_idx).to(dev)] = True
    hvg_nnz = 0
    for s0 in range(0, nnz, 1 << 26):
        hvg_nnz += int(fmask[csr.indices[s0:s0 + (1 << 26)].long()].sum())
    del fmask
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)  # nvidia-smi needs a moment before its first sample
    lib.LAUNCHES["n"] = 0
    if args.profiler_range:
        torch.cuda.profiler.start()
    wall0 = time.time()
    ms_total = timed(lambda: step(csr, timers), args.steps)
    wall1 = time.time()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    launches = lib.LAUNCHES["n"]
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = n_total / (ms_step / 1e3)

    # per-stage CUDA-event times of the timed steps (events were recorded on the launching stream)
    kernel_ms = [a.elapsed_time(b) for name, (a, b) in ((n_, e_) for n_, e_ in timers if n_ == "knn_kernel_events")]
    timers = [t for t in timers if t[0] != "knn_kernel_events"]
    stage_ms = {}
    for (n0, a), (n1, b) in zip(timers[:-1], timers[1:]):
        if n1 != "step_start":
            stage_ms.setdefault("mark_hvgs" if n1 == "start" else n1, []).append(a.elapsed_time(b))
    stage_ms = {k_: sum(v) / len(v) for k_, v in stage_ms.items()}
    pk = peaks()
    knn_ms = stage_ms.get("knn", float("nan"))  # whole entry point
    knn_kernel_ms = sum(kernel_ms) / len(kernel_ms) if kernel_ms else knn_ms  # knn_tc_kernel alone (events in the library)
    knn_flop = 2.0 * n_total * cfg["dims"] * n_local  # 2*N_ref*D per query, true D
    f16_peak_run = measure_f16_peak(torch, dev)
    # the driver-measured dense 16-bit peak is the denominator (burst figure: the kNN entry point runs for a few ms);
    # the in-run cuBLAS FP16 number is reported next to it
    tensor_peak = pk["bf16_tflops"]
    # the fused top-k' epilogue has to read every FP32 accumulator out of TMEM: 4 B per (query, reference) pair at the
    # 64 B/clk/SM tcgen05.ld rate (B300_MICROARCH.md) bounds the kernel from below whatever the tensor pipe does
    sm_hz = 1e6 * float((clocks or {}).get("sm_max_mhz") or 1965.0)
    tmem_floor_ms = 4.0 * n_total * n_local / (64.0 * 148 * sm_hz) * 1e3
    roofline = {"kernel": "knn_tc_kernel (tcgen05 kind::f16 distance contraction with fused top-k', the dominant kernel of "
                          "scf_knn_l2; timed with CUDA events recorded by the library around this launch alone; "
                          "entry_point_* = the whole call incl. operand prep, FP64 re-rank and guard repair)",
                "bound": "tensor", "achieved": knn_flop / (knn_kernel_ms * 1e-3) / 1e12, "peak": tensor_peak,
                "unit": "TFLOP/s", "frac": knn_flop / (knn_kernel_ms * 1e-3) / 1e12 / tensor_peak,
                "traffic": ncu_traffic("knn_tc_kernel"),
                "peak_note": f"dense 16-bit tensor peak from MEASURED_PEAKS.json ({pk['src']}; bf16 cuBLAS 8192^3, burst); "
                             f"cuBLAS FP16 8192^3 measured in this run: {f16_peak_run:.1f} TFLOP/s; HBM {pk['hbm_gbs']} GB/s. "
                             "achieved = 2*N_query*N_ref*D (true D, no padding credit) / CUDA-event time of the kernel",
                "ms_per_launch": knn_kernel_ms,
                "entry_point_ms": knn_ms,
                "entry_point_achieved": knn_flop / (knn_ms * 1e-3) / 1e12,
                "tmem_readout_floor_ms": tmem_floor_ms,
                "hbm_side": {k_: {"ms": stage_ms.get(k_), "algorithmic_GBs": v / (stage_ms[k_] * 1e-3) / 1e9,
                                  "frac_of_hbm_peak": v / (stage_ms[k_] * 1e-3) / 1e9 / pk["hbm_gbs"]}
                             for k_, v in (("normalise", 12.0 * hvg_nnz + 8.0 * n_local + 4.0 * 2048 * n_local *
                                            (2 if args.gram_mode == 3 else 1)),) if stage_ms.get(k_)}}
    csr_bytes = 8.0 * nnz + 8.0 * (n_local + 1)
    stages = {k_: round(v, 4) for k_, v in stage_ms.items()}

    # ---- e2e: host (pinned) CSR -> device -> graph -> host ----
    e2e = None
    if not args.no_e2e:
        h_ip, h_ix, h_dv = (t.cpu().pin_memory() for t in (csr.indptr, csr.indices, csr.data))
        out_host = {}

        def e2e_step():
            c = CsrDevice(h_ip.to(dev, non_blocking=True), h_ix.to(dev, non_blocking=True),
                          h_dv.to(dev, non_blocking=True), n_local, cfg["genes"])
            r = step(c)
            for name in ("indices", "distances", "weights"):
                t = getattr(r, name)
                if name not in out_host:
                    out_host[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                out_host[name].copy_(t, non_blocking=True)

        e2e_step()
        ms_e2e = timed(e2e_step, args.steps) / args.steps
        h2d = sum(t.numel() * t.element_size() for t in (h_ip, h_ix, h_dv))
        d2h = sum(t.numel() * t.element_size() for t in out_host.values())
        e2e = {"value": n_total / (ms_e2e / 1e3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e}

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dt, n = cpu_reference_run(cfg, args.cpu_sample, threads)
        cpu_base = {"value": n / dt, "unit": "cells/s", "cores": threads, "kind": "port",
                    "sample": f"first {n} cells of the workload (same generator/seed), one pass, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps({
            "metric": "make_graph_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 gene / column statistics, 3xTF32 Gram and projection, f16 tensor-core kNN candidates + f64 re-rank)",
            "data": "synthetic", "config": workload_config(args, cfg, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_base, "stage_ms": stages, "eig": eig_stats,
            "csr_bytes_per_gpu": csr_bytes, "nnz_per_cell": nnz / n_local,
        }))
    if world > 1:
        td.destroy_process_group()


def main():
    args = parse()
    cfg = dict(WORKLOADS[args.workload])
    if args.cells:
        cfg["cells"] = args.cells
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                                   "29531", os.path.abspath(__file__)] + sys.argv[1:])
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
