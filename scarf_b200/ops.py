"""Torch-tensor front end of the C-ABI: one function per kernel, device tensors in and out.

Every function launches on the current torch CUDA stream and returns without synchronising.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import lib

U32 = torch.int32  # uint32 counts are carried in int32 storage (same bits); only the kernels read them


@dataclass
class CsrDevice:
    """Raw counts in CSR form on one GPU (what ``Assay.to_raw_sparse`` yields, scarf/assay.py:175-199)."""
    indptr: torch.Tensor   # int64 [n_rows + 1]
    indices: torch.Tensor  # int32 [nnz], ascending inside a row
    data: torch.Tensor     # uint32 [nnz] in int32 storage
    n_rows: int
    n_cols: int

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    @property
    def device(self):
        return self.indptr.device

    def check_sorted(self) -> None:
        """Raises ValueError unless the column ids ascend strictly inside every row (the CSR contract of the library:
        the windowed gene statistics locate a row's gene windows by binary search).  One pass, for construction time."""
        nnz = self.nnz
        if nnz < 2:
            return
        step = 1 << 26
        starts = torch.zeros(nnz + 1, dtype=torch.bool, device=self.device)
        starts[self.indptr.clamp(max=nnz)] = True  # first position of every row (empty rows repeat a position)
        for lo in range(1, nnz, step):
            hi = min(nnz, lo + step)
            bad = (self.indices[lo:hi] <= self.indices[lo - 1:hi - 1]) & ~starts[lo:hi]
            if bool(bad.any()):
                raise ValueError("CSR column ids must ascend strictly inside every row (scipy: `m.sort_indices()` after "
                                 "`m.sum_duplicates()`)")

    @staticmethod
    def from_host(indptr, indices, data, shape, device="cuda", non_blocking=False, validate=False) -> "CsrDevice":
        import numpy as np

        def as_t(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            return torch.from_numpy(a.view(np.int32) if dt == np.uint32 else a)

        ip = as_t(indptr, np.int64).to(device, non_blocking=non_blocking)
        ix = as_t(indices, np.int32).to(device, non_blocking=non_blocking)
        dv = as_t(data, np.uint32).to(device, non_blocking=non_blocking)
        out = CsrDevice(ip, ix, dv, int(shape[0]), int(shape[1]))
        if validate:
            out.check_sorted()
        return out

    @staticmethod
    def from_scipy(m, device="cuda") -> "CsrDevice":
        m = m.tocsr()
        m.sort_indices()
        return CsrDevice.from_host(m.indptr, m.indices, m.data, m.shape, device)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t.data_ptr()


def _chk(t, dtype, name):
    if t is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------- dense store -> CSR
def dense_to_csr(block: torch.Tensor, n_cols: int):
    """Non-zero values of the first ``n_cols`` columns of a dense uint32 row block ``[rows, ld]`` (int32 storage,
    ld % 4 == 0) -> (row_nnz int64 [rows], indices int32 [nnz], data uint32-in-int32 [nnz]), columns ascending in a
    row.  One host synchronisation (the number of stored values sizes the outputs)."""
    _chk(block, U32, "block")
    rows, ld = int(block.shape[0]), int(block.stride(0))
    cnt = torch.empty(rows, dtype=torch.int64, device=block.device)
    lib.call("scf_dense_row_nnz", _ptr(block), rows, n_cols, ld, _ptr(cnt), _stream())
    ptr = torch.cumsum(cnt, 0) - cnt
    nnz = int(cnt.sum().item())
    idx = torch.empty(nnz, dtype=torch.int32, device=block.device)
    val = torch.empty(nnz, dtype=U32, device=block.device)
    if nnz:
        lib.call("scf_dense_to_csr", _ptr(block), rows, n_cols, ld, _ptr(ptr), _ptr(idx), _ptr(val), _stream())
    return cnt, idx, val


def csr_from_dense_zarr(arr, device="cuda", row_range=None, n_threads=None) -> CsrDevice:
    """Reads ``<assay>/counts`` of a Scarf store (dense uint32 N x G, chunks (rows, cols), Blosc or plain;
    scarf/writers.py:164-204) into a CSR on ``device``: per block of chunk rows the column chunks are decoded on host
    threads (the library's decoder runs without the GIL), copied into a dense staging block in HBM and converted by
    ``scf_dense_row_nnz`` / ``scf_dense_to_csr``.  ``row_range=(lo, hi)`` reads a shard (chunk-row aligned ``lo``)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    if arr.ndim != 2 or arr.dtype.kind not in "ui" or arr.dtype.itemsize > 4:
        raise TypeError(f"raw counts must be a 2-D integer array of at most 32 bits, got {arr.dtype} {arr.shape}")
    n, g = arr.shape
    lo, hi = (0, n) if row_range is None else row_range
    cr, cc = arr.chunks
    if lo % cr:
        raise ValueError(f"row_range must start on a chunk row boundary ({cr})")
    dev = torch.device(device)
    ld = round_up(max(g, 1), 4)
    stage = torch.zeros((cr, ld), dtype=U32, device=dev)
    n_cc = (g + cc - 1) // cc
    counts, idxs, vals = [], [], []

    def load(ci, cj):
        c = arr.read_chunk((ci, cj))
        c = np.ascontiguousarray(c, dtype=np.uint32) if c.dtype != np.uint32 else c
        return torch.from_numpy(c.view(np.int32))

    with ThreadPoolExecutor(max_workers=n_threads or min(16, os.cpu_count() or 1)) as pool:
        for ci in range(lo // cr, (hi + cr - 1) // cr):
            r0, r1 = ci * cr, min((ci + 1) * cr, hi)
            for cj, t in enumerate(pool.map(lambda j: load(ci, j), range(n_cc))):
                w = min(cc, g - cj * cc)
                stage[:, cj * cc: cj * cc + w].copy_(t[:, :w], non_blocking=False)
            c, i, v = dense_to_csr(stage[: r1 - r0], g)
            counts.append(c), idxs.append(i), vals.append(v)
    cnt = torch.cat(counts) if counts else torch.zeros(0, dtype=torch.int64, device=dev)
    indptr = torch.zeros(cnt.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt, 0, out=indptr[1:])
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt, device=dev)
    return CsrDevice(indptr, cat(idxs, torch.int32), cat(vals, U32), hi - lo, g)


# ------------------------------------------------------------------------------------------ K0
def csr_row_sums(csr: CsrDevice, row_ids=None, col_map=None):
    """(sum float64, nnz int32) per selected row over the columns with col_map >= 0 (all if None)."""
    _chk(row_ids, torch.int64, "row_ids"), _chk(col_map, torch.int32, "col_map")
    n = csr.n_rows if row_ids is None else int(row_ids.numel())
    s = torch.empty(n, dtype=torch.float64, device=csr.device)
    c = torch.empty(n, dtype=torch.int32, device=csr.device)
    lib.call("scf_csr_row_sums", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
             _ptr(col_map), _ptr(s), _ptr(c), _stream())
    return s, c


def csr_gene_stats(csr: CsrDevice, row_ids=None, row_div=None, sf=1000.0, out=None, with_moments=True,
                   windowed=True):
    """Per-gene (nnz uint64-in-int64, sum f64, sumsq f64) of sf*c/row_div over the selected rows; accumulates into out."""
    _chk(row_ids, torch.int64, "row_ids"), _chk(row_div, torch.float64, "row_div")
    n = csr.n_rows if row_ids is None else int(row_ids.numel())
    if out is None:
        nnz = torch.zeros(csr.n_cols, dtype=torch.int64, device=csr.device)
        sm = torch.zeros(csr.n_cols, dtype=torch.float64, device=csr.device) if with_moments else None
        sq = torch.zeros(csr.n_cols, dtype=torch.float64, device=csr.device) if with_moments else None
    else:
        nnz, sm, sq = out
    if windowed:
        ws_bytes = int(lib.raw("scf_csr_gene_stats_workspace_bytes")(n, csr.n_cols))
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=csr.device)
        lib.call("scf_csr_gene_stats_windowed", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
                 csr.n_cols, _ptr(row_div), float(sf), _ptr(nnz), _ptr(sm), _ptr(sq), ws.data_ptr(), ws_bytes,
                 _stream(), launches=2)
    else:
        lib.call("scf_csr_gene_stats", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
                 csr.n_cols, _ptr(row_div), float(sf), _ptr(nnz), _ptr(sm), _ptr(sq), _stream())
    return nnz, sm, sq


def hvg_select(nnz, sm, sq, feat_i, keep, m_cells, n_cells_total, n_bins, lowess_frac, top_n, min_cells, max_cells,
               min_mean, max_mean):
    """Fused trend removal + HVG choice on the per-gene statistics (one CTA).  ``feat_i`` / ``keep``: bool device
    vectors (keep may be None); bounds are the final, strict ones (+-inf = open).  -> (mask bool [G], col_map int32 [G],
    n_selected int32 [1]), all on the device, no synchronisation."""
    g = int(nnz.numel())
    dev = nnz.device
    _chk(nnz, torch.int64, "nnz"), _chk(sm, torch.float64, "sum"), _chk(sq, torch.float64, "sumsq")
    hv = torch.empty(g, dtype=torch.bool, device=dev)
    col_map = torch.empty(g, dtype=torch.int32, device=dev)
    n_sel = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.raw("scf_hvg_select_workspace_bytes")(g))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    fi = feat_i.to(torch.bool).contiguous()
    kp = keep.to(torch.bool).contiguous() if keep is not None else None
    lib.call("scf_hvg_select", _ptr(nnz), _ptr(sm), _ptr(sq), fi.data_ptr(), kp.data_ptr() if kp is not None else None,
             g, float(m_cells), float(n_cells_total), int(n_bins), float(lowess_frac), int(top_n), float(min_cells),
             float(max_cells), float(min_mean), float(max_mean), hv.data_ptr(), _ptr(col_map), _ptr(n_sel),
             ws.data_ptr(), ws_bytes, _stream())
    return hv, col_map, n_sel


def lowess(endog, exog, valid=None, frac=0.1, it=100):
    """Robust LOWESS fit at every usable point, on the device (float64 vectors of <= 512 points; NaN elsewhere)."""
    _chk(endog, torch.float64, "endog"), _chk(exog, torch.float64, "exog"), _chk(valid, torch.uint8, "valid")
    out = torch.empty_like(endog)
    lib.call("scf_lowess", _ptr(endog), _ptr(exog), _ptr(valid), int(endog.numel()), float(frac), int(it), _ptr(out),
             _stream())
    return out


# ------------------------------------------------------------------------------------------ K1
def csr_hvg_colstats(csr: CsrDevice, row_ids, col_map, n_cols, row_sum, sf=1000.0, log_transform=True, n_rep=32):
    """int64 fixed-point (sum x, sum x^2) per selected column; divide by 2**lib.COLSTAT_SHIFT."""
    n = csr.n_rows if row_ids is None else int(row_ids.numel())
    acc = torch.zeros((2, n_rep, n_cols), dtype=torch.int64, device=csr.device)
    lib.call("scf_csr_hvg_colstats", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
             _ptr(col_map), _ptr(row_sum), float(sf), int(bool(log_transform)), int(n_cols), int(n_rep),
             acc[0].data_ptr(), acc[1].data_ptr(), _stream())
    tot = acc.sum(dim=1)
    return tot[0], tot[1]


def csr_hvg_compact(csr: CsrDevice, row_ids, col_map, n_cols, row_sum, row_nnz, sf=1000.0, log_transform=True,
                    n_rep=32):
    """One scan: compact normalised matrix of the selected non-zero entries + fixed-point column sums.
    ``row_nnz`` = int32 per-row count from :func:`csr_row_sums` with the same ``col_map``.
    -> (row_off int64 [n+1], cols int32, xs float64, sum_fx, sumsq_fx)."""
    n = csr.n_rows if row_ids is None else int(row_ids.numel())
    row_off = torch.zeros(n + 1, dtype=torch.int64, device=csr.device)
    torch.cumsum(row_nnz, dim=0, out=row_off[1:])
    total = int(row_off[-1].item())
    cols = torch.empty(max(total, 1), dtype=torch.int32, device=csr.device)
    xs = torch.empty(max(total, 1), dtype=torch.float64, device=csr.device)
    acc = torch.zeros((2, n_rep, n_cols), dtype=torch.int64, device=csr.device)
    lib.call("scf_csr_hvg_compact", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
             _ptr(col_map), _ptr(row_sum), float(sf), int(bool(log_transform)), _ptr(row_off), _ptr(cols), _ptr(xs),
             int(n_cols), int(n_rep), acc[0].data_ptr(), acc[1].data_ptr(), _stream())
    tot = acc.sum(dim=1)
    return row_off, cols, xs, tot[0], tot[1]


def hvg_dense_scale(row_off, cols, xs, n_cols, z, mu=None, sigma=None, z_lo=None):
    """Z rows [0, n) (and z_lo) from the compact matrix of :func:`csr_hvg_compact`; same values as csr_norm_scale."""
    n = int(row_off.numel()) - 1
    assert z.dtype == torch.float32 and z.stride(1) == 1 and z.shape[0] >= n
    lib.call("scf_hvg_dense_scale", _ptr(row_off), _ptr(cols), _ptr(xs), n, int(n_cols), _ptr(mu), _ptr(sigma),
             z.data_ptr(), z_lo.data_ptr() if z_lo is not None else None, int(z.stride(0)), _stream())
    return z


def csr_norm_scale(csr: CsrDevice, row_ids, col_map, n_cols, row_sum, z, sf=1000.0, log_transform=True, mu=None,
                   sigma=None, missing_fill=None, z_lo=None):
    """Writes rows [0, n_sel) of the float32 matrix z (row stride z.stride(0)) with (x - mu)/sigma; ``z_lo``
    (optional, same shape and stride) receives z - tf32_trunc(z) for the 3xTF32 Gram."""
    n = csr.n_rows if row_ids is None else int(row_ids.numel())
    _chk(z, torch.float32, "z"), _chk(mu, torch.float64, "mu"), _chk(sigma, torch.float64, "sigma")
    assert z.dim() == 2 and z.stride(1) == 1 and z.shape[0] >= n and z.is_cuda
    if z_lo is not None:
        assert z_lo.dtype == torch.float32 and z_lo.shape == z.shape and z_lo.stride() == z.stride()
    lib.call("scf_csr_norm_scale", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), _ptr(row_ids), n,
             _ptr(col_map), int(n_cols), _ptr(row_sum), float(sf), int(bool(log_transform)), _ptr(mu), _ptr(sigma),
             _ptr(missing_fill), z.data_ptr(), z_lo.data_ptr() if z_lo is not None else None, int(z.stride(0)),
             _stream())
    return z


# ------------------------------------------------------------------------------------------ K2 / K4
def gram_accumulate(z, n_rows, n_cols, g_fx=None, mode=0, z_lo=None):
    """g_fx (int64, << lib.GRAM_SHIFT) += Z[:n_rows, :n_cols]^T Z[:n_rows, :n_cols], upper triangle; call
    :func:`gram_symmetrize` after the last accumulation / all-reduce.  g_fx is allocated [ldz, ldz] so that the
    tensor-core epilogue never needs to clip a row segment."""
    assert z.dtype == torch.float32 and z.stride(1) == 1
    if g_fx is None:
        ld = round_up(n_cols, 32)
        g_fx = torch.zeros((ld, ld), dtype=torch.int64, device=z.device)
    lib.call("scf_gram_accumulate", z.data_ptr(), z_lo.data_ptr() if z_lo is not None else None, int(z.stride(0)),
             int(n_rows), int(n_cols), _ptr(g_fx), int(g_fx.stride(0)), int(mode), _stream())
    return g_fx


def gram_symmetrize(g_fx, n_cols):
    lib.call("scf_gram_symmetrize", _ptr(g_fx), int(n_cols), int(g_fx.stride(0)), _stream())
    return g_fx


def project(z, n_rows, n_cols, v, dims, y=None, ldy=None, z_lo=None):
    """y[:n_rows, :dims] = z[:n_rows, :n_cols] @ v[:n_cols, :dims]; pad columns of y are zeroed.  With the second plane
    ``z_lo`` (= z - tf32_trunc(z), the 3xTF32 Gram operand) the product runs on the tensor cores."""
    assert z.dtype == torch.float32 and v.dtype == torch.float32 and v.stride(1) == 1
    if y is None:
        ldy = ldy or round_up(dims, 32)
        y = torch.empty((n_rows, ldy), dtype=torch.float32, device=z.device)
    if z_lo is not None and dims <= 128 and z.stride(0) % 32 == 0 and y.stride(0) % 4 == 0:
        assert z_lo.dtype == torch.float32 and z_lo.stride() == z.stride()
        ws_bytes = int(lib.raw("scf_project_tc_workspace_bytes")(int(z.stride(0)), int(dims)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
        lib.call("scf_project_tc", z.data_ptr(), z_lo.data_ptr(), int(z.stride(0)), int(n_rows), int(n_cols),
                 v.data_ptr(), int(v.stride(0)), int(dims), y.data_ptr(), int(y.stride(0)), ws.data_ptr(), ws_bytes,
                 _stream(), launches=2)
        return y
    lib.call("scf_project", z.data_ptr(), int(z.stride(0)), int(n_rows), int(n_cols), v.data_ptr(), int(v.stride(0)),
             int(dims), y.data_ptr(), int(y.stride(0)), _stream())
    return y


def eig_topk(g_fx, n_cols, dims, scale, col_mean=None, mean_weight=0.0, tol=1e-8, max_rounds=24, stats=None,
             ld32=None):
    """K3 (scf_eig_topk): top-``dims`` eigenpairs of ``g_fx[:n_cols, :n_cols] * scale - mean_weight * outer(col_mean)``
    (g_fx: the mirrored int64 fixed-point Gram).  -> (eigenvalues float64 [dims] descending, eigenvectors float64
    [n_cols, dims], sklearn sign rule[, float32 copy [n_cols, ld32] with zero pad columns when ``ld32`` is given]).
    Native Chebyshev-filtered subspace iteration; one stream synchronisation per round.  ``stats`` (dict) receives
    rounds / residual / restarts."""
    assert g_fx.dtype == torch.int64 and g_fx.is_cuda and g_fx.stride(1) == 1
    dev = g_fx.device
    ws_bytes = int(lib.raw("scf_eig_topk_workspace_bytes")(int(n_cols), int(dims)))
    if ws_bytes < 0:
        raise NotImplementedError(f"scf_eig_topk: dims = {dims} of {n_cols} features is outside the native solver "
                                  "(dims + 8 <= 160 columns)")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    evals = torch.empty(dims, dtype=torch.float64, device=dev)
    evecs = torch.empty((n_cols, dims), dtype=torch.float64, device=dev)
    v32 = torch.empty((n_cols, int(ld32)), dtype=torch.float32, device=dev) if ld32 else None
    report = torch.zeros(176, dtype=torch.float64).pin_memory()
    _chk(col_mean, torch.float64, "col_mean")
    try:
        lib.call("scf_eig_topk", g_fx.data_ptr(), int(g_fx.stride(0)), int(n_cols), float(scale), _ptr(col_mean),
                 float(mean_weight), int(dims), float(tol), int(max_rounds), evals.data_ptr(), evecs.data_ptr(),
                 _ptr(v32), int(ld32 or 0), report.data_ptr(), ws.data_ptr(), ws_bytes, _stream(), launches=0)
    finally:
        lib.LAUNCHES["n"] += int(report[168])
        if stats is not None:
            stats["eig_rounds"], stats["eig_residual"] = int(report[5]), float(report[0])
            stats["eig_restarts"], stats["eig_robust"] = int(report[6]), bool(report[7])
            stats["eig_min_pivot"] = float(report[4])
            stats["eig_ritz_solver"] = "tridiagonal" if report[169] == 1.0 else "jacobi"
    return (evals, evecs, v32) if ld32 else (evals, evecs)


def sym_eig_small(a, info=None):
    """(eigenvalues ascending, eigenvectors as columns) of a small symmetric PSD float64 matrix on the device: the
    one-CTA Jacobi kernel the native eigensolver uses for its Rayleigh-Ritz matrices (n <= scf_sym_eig_max_n())."""
    n = int(a.shape[0])
    if n > int(lib.raw("scf_sym_eig_max_n")()):
        raise NotImplementedError(f"scf_sym_eig_jacobi: n = {n} exceeds {int(lib.raw('scf_sym_eig_max_n')())}")
    _chk(a, torch.float64, "a")
    a = a.contiguous()
    w = torch.empty(n, dtype=torch.float64, device=a.device)
    v = torch.empty((n, n), dtype=torch.float64, device=a.device)
    lib.call("scf_sym_eig_jacobi", _ptr(a), n, int(a.stride(0)), _ptr(w), _ptr(v), int(v.stride(0)),
             _ptr(info), _stream())
    return w, v


def sym_eig_tridiag(a):
    """Same result as :func:`sym_eig_small` from the tridiagonal kernel (scf_sym_eig_tridiag) -> (w, v, ok): ``ok`` is a
    device int32 scalar, 1 when the eigenvectors are orthonormal to 1e-9 (else the Jacobi kernel has to be used)."""
    n = int(a.shape[0])
    _chk(a, torch.float64, "a")
    a = a.contiguous()
    w = torch.empty(n, dtype=torch.float64, device=a.device)
    v = torch.empty((n, n), dtype=torch.float64, device=a.device)
    work = torch.empty(2 * n * n + 3 * n, dtype=torch.float64, device=a.device)
    ok = torch.zeros(1, dtype=torch.int32, device=a.device)
    lib.call("scf_sym_eig_tridiag", _ptr(a), n, int(a.stride(0)), _ptr(w), _ptr(v), int(v.stride(0)), _ptr(work),
             _ptr(ok), _stream())
    return w, v, ok


# ------------------------------------------------------------------------------------------ K5
def knn_l2(q, ref, dim, k, self_offset=-1, method=0, stats=None, kernel_events=None):
    """Exact kNN (squared L2).  q, ref: float32 [n, ld] sharing the row stride.  -> (int64 idx, float32 dist).
    ``stats`` (optional dict) receives ``guard_fail_rows`` / ``guard_rest_rows`` as device scalar tensors (method 1).
    ``kernel_events``: optional pair of recorded ``torch.cuda.Event(enable_timing=True)``; the library re-records them
    around the tensor-core kernel of this call alone (measurement hook of bench.py)."""
    assert q.dtype == torch.float32 and ref.dtype == torch.float32
    assert q.stride(1) == 1 and ref.stride(1) == 1 and q.stride(0) == ref.stride(0)
    nq, nref = int(q.shape[0]), int(ref.shape[0])
    idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    ws_bytes = int(lib.raw("scf_knn_workspace_bytes")(nq, nref, int(dim), int(k), int(method)))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=q.device)
    if kernel_events is not None and method == 1:
        lib.call("scf_knn_time_next_call", kernel_events[0].cuda_event, kernel_events[1].cuda_event, launches=0)
    lib.call("scf_knn_l2", q.data_ptr(), nq, ref.data_ptr(), nref, int(dim), int(q.stride(0)), int(k),
             int(self_offset), idx.data_ptr(), dist.data_ptr(), int(method), ws.data_ptr(), ws_bytes, _stream())
    if stats is not None:
        off = int(lib.raw("scf_knn_fail_count_offset")(nq, nref, int(dim), int(k), int(method)))
        stats["guard_fail_rows"] = ws[off:off + 4].view(torch.int32).clone() if off >= 0 else None
        # rows the tensor-core collect pass could not settle either (FP64 scan over all references)
        stats["guard_rest_rows"] = ws[off + 128:off + 132].view(torch.int32).clone() if off >= 0 else None
    return idx, dist


# ------------------------------------------------------------------------------------------ K6
def chunk_sums(dist, row_offset, chunk_size, n_chunks_total):
    """float64 sum of the distances of every chunk this shard touches (other entries stay 0)."""
    n, k = dist.shape
    out = torch.zeros(n_chunks_total, dtype=torch.float64, device=dist.device)
    lib.call("scf_chunk_sums", dist.data_ptr(), n, k, int(row_offset), int(chunk_size), out.data_ptr(), _stream())
    return out


def smooth_knn(dist, chunk_mean, lc=1.0, bw=1.5, row_offset=0, chunk_size=1000):
    """umap smooth_knn_dist per row -> (sigma, rho) float32."""
    n, k = dist.shape
    assert dist.dtype == torch.float32 and dist.is_contiguous() and chunk_mean.dtype == torch.float32
    sigma = torch.empty(n, dtype=torch.float32, device=dist.device)
    rho = torch.empty(n, dtype=torch.float32, device=dist.device)
    lib.call("scf_smooth_knn", dist.data_ptr(), n, k, float(lc), float(bw), int(row_offset), int(chunk_size),
             chunk_mean.data_ptr(), sigma.data_ptr(), rho.data_ptr(), _stream())
    return sigma, rho


def membership_coo(idx, dist, sigma, rho, row_offset, chunk_size, n_chunks_total):
    """compute_membership_strengths + COO rows -> edges int64 [n*k,2], weights float32 [n*k],
    per-chunk (min non-zero weight, has-zero flag) indexed by global chunk id."""
    n, k = idx.shape
    dev = idx.device
    assert idx.dtype == torch.int64 and idx.is_contiguous() and dist.is_contiguous()
    edges = torch.empty((n * k, 2), dtype=torch.int64, device=dev)
    weights = torch.empty(n * k, dtype=torch.float32, device=dev)
    cmin = torch.full((n_chunks_total,), math.inf, dtype=torch.float32, device=dev)
    czero = torch.zeros(n_chunks_total, dtype=torch.int32, device=dev)
    lib.call("scf_membership_coo", idx.data_ptr(), dist.data_ptr(), sigma.data_ptr(), rho.data_ptr(), n, k,
             int(row_offset), int(chunk_size), edges.data_ptr(), weights.data_ptr(), cmin.data_ptr(),
             czero.data_ptr(), _stream())
    return edges, weights, cmin, czero


def fill_zero_weights(weights, floor):
    lib.call("scf_fill_zero_weights", weights.data_ptr(), int(weights.numel()), float(floor), _stream())
    return weights


def graph_symmetrize(idx, weights, use_k, upper_only):
    """COO entries of g + g.T - g * g.T (optionally its upper triangle) of the kNN graph given as neighbour ids int64
    [n, k] and weights float64 [n, k] on the device -> (rows int64, cols int64, vals float64), unused slots have row -1."""
    n, k = idx.shape
    assert idx.dtype == torch.int64 and weights.dtype == torch.float64 and idx.is_contiguous() and weights.is_contiguous()
    m = 2 * n * int(use_k)
    rows = torch.empty(m, dtype=torch.int64, device=idx.device)
    cols = torch.empty(m, dtype=torch.int64, device=idx.device)
    vals = torch.empty(m, dtype=torch.float64, device=idx.device)
    lib.call("scf_graph_symmetrize", idx.data_ptr(), weights.data_ptr(), n, k, int(use_k), int(bool(upper_only)),
             rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), _stream())
    return rows, cols, vals
