"""Seeded synthetic scRNA-seq counts (SURVEY.md App. C recipe) as CSR, generated block-wise with torch
on any device.  Used by the tests (small, CPU), the benchmark (on the GPU) and smoke().

Model: gene base rates ~ Gamma(0.3); L latent factors with sparse (4 %) N(0,1) gene loadings and
geometrically decaying strengths 1.2 * 0.93**l; per-cell library size ~ logN(log 4000, 0.4);
counts ~ Poisson(lib * softmax_j(log base_j + sum_l s_l f_il W_lj)).  Rows are generated in blocks of
``block`` cells seeded by (seed, global block id), so any shard of the global matrix can be produced
independently and bit-identically on any rank.
"""
from __future__ import annotations

import numpy as np
import torch

# NOTE: nothing here may import .ops / .lib at module level: the CPU arm of bench.py uses this generator and must not
# map the CUDA library


def gene_model(n_genes: int, n_factors: int, seed: int):
    rng = np.random.default_rng(seed)
    base = rng.gamma(0.3, 1.0, size=n_genes) + 1e-6
    w = rng.normal(size=(n_factors, n_genes)) * (rng.random((n_factors, n_genes)) < 0.04)
    strength = 1.2 * 0.93 ** np.arange(n_factors)
    return np.log(base).astype(np.float32), (w * strength[:, None]).astype(np.float32)


def _block_counts(log_base, w, n, block_id, seed, lib_mean, lib_sigma, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1_000_003 + int(block_id))
    f = torch.randn((n, w.shape[0]), generator=g, device=device, dtype=torch.float32)
    lib = torch.exp(torch.randn((n, 1), generator=g, device=device, dtype=torch.float32) * lib_sigma
                    + float(np.log(lib_mean)))
    logits = f @ w + log_base
    lam = torch.softmax(logits, dim=1) * lib
    return torch.poisson(lam, generator=g)


def make_counts_torch(n_cells: int, n_genes: int, n_factors: int = 65, seed: int = 4466, device="cuda",
                      block: int = 4096, row_start: int = 0, lib_mean: float = 4000.0, lib_sigma: float = 0.4):
    """-> (indptr int64 [n+1], indices int32, data int32-storage uint32) on ``device`` for global rows
    [row_start, row_start + n_cells); row_start must be a multiple of ``block``."""
    assert row_start % block == 0
    device = torch.device(device)
    lb, w = gene_model(n_genes, n_factors, seed)
    lb, w = torch.from_numpy(lb).to(device), torch.from_numpy(w).to(device)
    counts_per_row, idx_parts, val_parts = [], [], []
    for s in range(0, n_cells, block):
        n = min(block, n_cells - s)
        full = _block_counts(lb, w, block, (row_start + s) // block, seed, lib_mean, lib_sigma, device)[:n]
        nz = full > 0
        counts_per_row.append(nz.sum(dim=1))
        r, c = torch.nonzero(nz, as_tuple=True)  # row-major order -> columns ascending inside a row
        idx_parts.append(c.to(torch.int32))
        val_parts.append(full[r, c].to(torch.int32))
        del full, nz, r, c
    per_row = torch.cat(counts_per_row)
    indptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(per_row, 0)
    return indptr, torch.cat(idx_parts), torch.cat(val_parts)


def make_counts_device(n_cells, n_genes, n_factors=65, seed=4466, device="cuda", **kw):
    from .ops import CsrDevice

    ip, ix, dv = make_counts_torch(n_cells, n_genes, n_factors, seed, device, **kw)
    return CsrDevice(ip, ix, dv, n_cells, n_genes)


def make_counts_scipy(n_cells, n_genes, n_factors=65, seed=4466, **kw):
    """CPU generation -> scipy.sparse.csr_matrix (uint32), for tests and the CPU baseline."""
    import scipy.sparse as sp

    ip, ix, dv = make_counts_torch(n_cells, n_genes, n_factors, seed, "cpu", **kw)
    return sp.csr_matrix((dv.numpy().astype(np.uint32), ix.numpy(), ip.numpy()), shape=(n_cells, n_genes))


def to_scipy(csr):
    import scipy.sparse as sp

    return sp.csr_matrix((csr.data.cpu().numpy().view(np.uint32), csr.indices.cpu().numpy(),
                          csr.indptr.cpu().numpy()), shape=(csr.n_rows, csr.n_cols))
