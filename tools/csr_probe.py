"""Developer probe (GPU): CUDA-event timings of the CSR-side kernels on a C2-sized shard, with the algorithmic
bytes of SURVEY 8(d) against the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from scarf_b200 import graph, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dev = torch.device("cuda", 0)
csr = synth.make_counts_device(n, 30_000, 65, seed=4466, device=dev, block=2000)
nnz = csr.nnz
peak = 6455.3
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
n_counts, _ = graph.cell_totals(csr)
feat_I = graph.gene_ncells(csr) > 20
hv = graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True,
                         keep_mask=torch.ones(30_000, dtype=torch.bool, device=dev))
rank_t = torch.cumsum(hv, dim=0, dtype=torch.int32)
col_map = torch.where(hv, rank_t - 1, torch.full_like(rank_t, -1)).contiguous()
csr_bytes = 8.0 * nnz + 8.0 * n


def timed(name, fn, bytes_, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:34s} {best:8.3f} ms  {bytes_ / best / 1e6:8.1f} GB/s alg  {bytes_ / best / 1e6 / peak:5.2f} of HBM peak")
    return out


timed("row_sums (all genes)", lambda: ops.csr_row_sums(csr), csr_bytes + 12 * n)
row_sum, row_nnz = timed("row_sums (HVG subset)", lambda: ops.csr_row_sums(csr, None, col_map), csr_bytes + 12 * n)
timed("gene_stats plain (3 RED/value)", lambda: ops.csr_gene_stats(csr, None, n_counts, 1000.0, windowed=False), csr_bytes + 8 * n)
timed("gene_stats windowed (smem RMW)", lambda: ops.csr_gene_stats(csr, None, n_counts, 1000.0, windowed=True), csr_bytes + 8 * n)
timed("gene_ncells plain", lambda: ops.csr_gene_stats(csr, None, None, with_moments=False, windowed=False), csr_bytes)
timed("gene_ncells windowed", lambda: ops.csr_gene_stats(csr, None, None, with_moments=False, windowed=True), csr_bytes)
ref_sums = None
for variant in ("0", "1", "2", "3", "4", None):  # None: the library's default
    if variant is None:
        os.environ.pop("SCF_COMPACT_VARIANT", None)
    else:
        os.environ["SCF_COMPACT_VARIANT"] = variant
    comp = timed(f"hvg_compact (variant {variant})", lambda: ops.csr_hvg_compact(csr, None, col_map, 2000, row_sum, row_nnz),
                 csr_bytes + 8 * n + 12.0 * float(row_nnz.sum()))
    row_off, cols, xs, sx, sxx = comp
    if ref_sums is None:
        ref_sums = (sx.clone(), sxx.clone())
    assert torch.equal(sx, ref_sums[0]) and torch.equal(sxx, ref_sums[1]), "column sums differ between variants"
hnnz = int(cols.numel())
print("hvg nnz per cell", hnnz / n)
mu = torch.zeros(2000, dtype=torch.float64, device=dev)
sigma = torch.ones(2000, dtype=torch.float64, device=dev)
z = torch.empty((n, 2048), dtype=torch.float32, device=dev)
z_lo = torch.empty_like(z)
for seg in ("128", "256", "512", "1024", "2048", None):
    if seg is None:
        os.environ.pop("SCF_DENSE_SEG", None)
    else:
        os.environ["SCF_DENSE_SEG"] = seg
    timed(f"hvg_dense_scale (Z + Z_lo) seg {seg}", lambda: ops.hvg_dense_scale(row_off, cols, xs, 2000, z, mu, sigma, z_lo=z_lo),
          12.0 * hnnz + 8 * n + 2 * 4.0 * 2048 * n)
timed("hvg_dense_scale (Z only)", lambda: ops.hvg_dense_scale(row_off, cols, xs, 2000, z, mu, sigma),
      12.0 * hnnz + 8 * n + 4.0 * 2048 * n)
timed("csr_norm_scale (Z + Z_lo)", lambda: ops.csr_norm_scale(csr, None, col_map, 2000, row_sum, z, mu=mu, sigma=sigma, z_lo=z_lo),
      csr_bytes + 2 * 4.0 * 2048 * n)
v32 = torch.randn((2000, 52), dtype=torch.float32, device=dev)
timed("project D=50 (FP32 SIMT)", lambda: ops.project(z, n, 2000, v32, 50), 4.0 * 2048 * n + 4 * 64 * n)
timed("project D=50 (3xTF32 tcgen05)", lambda: ops.project(z, n, 2000, v32, 50, z_lo=z_lo), 2 * 4.0 * 2048 * n + 4 * 64 * n)
v100 = torch.randn((2000, 100), dtype=torch.float32, device=dev)
timed("project D=100 (FP32 SIMT)", lambda: ops.project(z, n, 2000, v100, 100), 4.0 * 2048 * n + 4 * 128 * n)
timed("project D=100 (3xTF32 tcgen05)", lambda: ops.project(z, n, 2000, v100, 100, z_lo=z_lo), 2 * 4.0 * 2048 * n + 4 * 128 * n)
st = {}
cov = torch.randn((2000, 2000), dtype=torch.float64, device=dev)
timed("mark_hvgs_csr (whole)", lambda: graph.mark_hvgs_csr(csr, None, feat_I, n_counts, n, top_n=2000, as_tensor=True,
                                                           keep_mask=torch.ones(30_000, dtype=torch.bool, device=dev)), csr_bytes)
