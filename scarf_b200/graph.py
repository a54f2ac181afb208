"""The make_graph hot path on device-resident CSR shards.

``mark_hvgs_csr`` and ``make_graph_csr`` are the numeric cores that the Scarf-compatible
``DataStore`` (scarf_b200/datastore.py) calls; they take and return torch tensors so that the
benchmark can time them with the inputs already in HBM.  Reference call stack: SURVEY.md 3.1-3.3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import hvg as hvg_host
from . import lib, ops
from .dist import Comm
from .ops import CsrDevice, round_up

SF = 1000.0  # RNAassay.sf (scarf/assay.py:776)


# =============================================================================================
# statistics that the reference computes when a DataStore is first opened / in mark_hvgs
# =============================================================================================
def cell_totals(csr: CsrDevice):
    """nCounts (float64) and nFeatures (int32) of every stored cell (scarf/datastore/base_datastore.py:345-366)."""
    return ops.csr_row_sums(csr)


def gene_ncells(csr: CsrDevice, comm: Comm | None = None):
    """Per-feature nCells over ALL stored cells (scarf/assay.py:201-225)."""
    nnz, _, _ = ops.csr_gene_stats(csr, None, None, with_moments=False)
    if comm is not None:
        comm.allreduce_sum_(nnz)
    return nnz


def hvg_gene_stats(csr: CsrDevice, cell_idx, n_counts, n_cells_total: int, comm: Comm | None = None, as_numpy=True):
    """RNAassay.set_feature_stats (scarf/assay.py:830-897): per-gene normed_n / normed_tot / sigmas /
    avg / nz_mean of ``sf*c/nCounts`` over the ``cell_idx`` rows: float64 vectors over all genes."""
    row_div = n_counts[cell_idx] if cell_idx is not None else n_counts
    nnz, sm, sq = ops.csr_gene_stats(csr, cell_idx, row_div.contiguous(), SF)
    m = csr.n_rows if cell_idx is None else int(cell_idx.numel())
    if comm is not None and comm.world > 1:
        mt = torch.tensor([m], dtype=torch.int64, device=csr.device)
        comm.allreduce_sum_(nnz), comm.allreduce_sum_(sm), comm.allreduce_sum_(sq), comm.allreduce_sum_(mt)
        m = int(mt.item())
    m = float(m)
    n = nnz.to(torch.float64)
    mean = sm / m
    var = torch.clamp(sq / m - mean * mean, min=0.0)  # population variance (dask var, ddof 0)
    nz_mean = torch.where(n > 0, sm / torch.clamp(n, min=1.0), torch.zeros_like(sm))
    out = {"normed_n": n, "normed_tot": sm, "sigmas": var, "avg": sm / float(n_cells_total), "nz_mean": nz_mean}
    return {k: v.cpu().numpy() for k, v in out.items()} if as_numpy else out


def mark_hvgs_csr(csr: CsrDevice, cell_idx, feat_I, n_counts, n_cells_total, gene_names=None, top_n=500,
                  min_cells=None, max_cells=np.inf, min_mean=-np.inf, max_mean=np.inf, n_bins=200, lowess_frac=0.1,
                  blacklist=hvg_host.DEFAULT_BLACKLIST, comm: Comm | None = None, return_stats=False,
                  as_tensor=False, keep_mask=None, min_var=-np.inf, max_var=np.inf, keep_bounds=False):
    """DataStore.mark_hvgs (scarf/datastore/datastore.py:223-314) -> bool mask over all genes (numpy, or a device
    tensor with ``as_tensor``).  The per-gene vectors never leave the GPU; only the <= n_bins binned points of the
    trend fit visit the host (LOWESS).  ``keep_mask`` = precomputed blacklist survivors (static per dataset).
    ``min_var`` / ``max_var`` (log2; scarf/assay.py:1014-1063): a finite ``min_var`` replaces the top-n rule;
    ``keep_bounds``: bounds inclusive (MetaData.sift).  Both leave the fused kernel for the tensor formulation."""
    if min_cells is None:
        min_cells = int(0.01 * n_cells_total)  # datastore.py:291
    dev = csr.device
    if min_var == -np.inf and top_n < 1:
        raise ValueError("ERROR: Please provide a value greater than 0 for `top_n` parameter")  # assay.py:1030-1033
    if (csr.indptr.is_cuda and not return_stats and min_var == -np.inf and max_var == np.inf and not keep_bounds
            and n_bins <= 512):
        # device path: one statistics pass + ONE fused kernel for trend removal and choice (no synchronisation)
        row_div = n_counts[cell_idx] if cell_idx is not None else n_counts
        nnz, sm, sq = ops.csr_gene_stats(csr, cell_idx, row_div.contiguous(), SF)
        m = csr.n_rows if cell_idx is None else int(cell_idx.numel())
        if comm is not None and comm.world > 1:
            mt = torch.tensor([m], dtype=torch.int64, device=dev)
            comm.allreduce_sum_(nnz), comm.allreduce_sum_(sm), comm.allreduce_sum_(sq), comm.allreduce_sum_(mt)
            m = int(mt.item())
        if keep_mask is None:
            keep_mask = hvg_host.blacklist_keep_mask(gene_names, csr.n_cols, blacklist)
        to_dev = lambda x: (x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x, dtype=bool))).to(dev)
        lo_mean = 2.0 ** min_mean if min_mean != -np.inf else -np.inf  # log2 thresholds (scarf/assay.py:1014-1021)
        hi_mean = 2.0 ** max_mean if max_mean != np.inf else np.inf
        mask, _, _ = ops.hvg_select(nnz, sm, sq, to_dev(feat_I), to_dev(keep_mask), m, n_cells_total, n_bins,
                                    lowess_frac, top_n, min_cells, max_cells, lo_mean, hi_mean)
        return mask if as_tensor else mask.cpu().numpy()
    st = hvg_gene_stats(csr, cell_idx, n_counts, n_cells_total, comm, as_numpy=False)
    if keep_mask is None:
        keep_mask = hvg_host.blacklist_keep_mask(gene_names, csr.n_cols, blacklist)
    feat_I_t = feat_I if torch.is_tensor(feat_I) else torch.from_numpy(np.asarray(feat_I, dtype=bool))
    keep_t = keep_mask if torch.is_tensor(keep_mask) else torch.from_numpy(np.asarray(keep_mask, dtype=bool))
    feat_I_t, keep_t = feat_I_t.to(dev), keep_t.to(dev)
    c_var = hvg_host.remove_trend_device(st["avg"], st["sigmas"], n_bins, lowess_frac, select=feat_I_t)
    c_var = torch.where(feat_I_t, c_var, torch.full_like(c_var, math.nan))
    mask = hvg_host.choose_hvgs_device(st["normed_n"], st["nz_mean"], c_var, feat_I_t & keep_t, top_n, min_cells,
                                       max_cells, min_mean, max_mean, min_var, max_var, keep_bounds)
    out = mask if as_tensor else mask.cpu().numpy()
    if return_stats:
        st = {k: v.cpu().numpy() for k, v in st.items()}
        st["c_var"] = c_var.cpu().numpy()
        return out, st
    return out


# =============================================================================================
# make_graph
# =============================================================================================
@dataclass
class GraphResult:
    """Device tensors produced by :func:`make_graph_csr` for this rank's rows."""
    n_cells: int            # selected cells over all ranks
    row_offset: int         # global (selected-row) id of this rank's first row
    feat_idx: np.ndarray    # ascending gene ids of the features used
    mu: torch.Tensor        # float64 [H]
    sigma: torch.Tensor     # float64 [H]
    loadings: torch.Tensor  # float64 [H, dims]
    eigenvalues: torch.Tensor
    embedding: torch.Tensor  # float32 [n_local, ldy] (pad columns zero) -- what hnswlib would have indexed
    embedding_all: torch.Tensor  # the embedding of ALL selected cells (== embedding on one GPU): run_mapping's index
    dims: int
    k: int
    indices: torch.Tensor    # int64 [n_local, k]
    distances: torch.Tensor  # float32 [n_local, k] squared L2
    edges: torch.Tensor      # int64 [n_local*k, 2]
    weights: torch.Tensor    # float32 [n_local*k]


def clean_array(x, fill_val=0.0):
    """scarf/utils.py:143-153: NaN -> 0, then zeros -> ``fill_val``.  As executed, the reference's ``np.nan_to_num``
    has already turned +-inf into +-the largest float64 when its ``x == inf`` line runs, so infinities do NOT become
    zero (tests/golden/ref_functions.npz, made by running the reference function); same here."""
    x = torch.nan_to_num(x)
    return torch.where(x == 0, torch.full_like(x, fill_val), x)


def clamp_dims(dims, n_cells, batch_size):
    """scarf/ann.py:173-185."""
    if dims > n_cells:
        dims = n_cells
    if dims >= batch_size:
        dims = batch_size - 1
    return dims


def sign_rule(vt):
    """sklearn svd_flip(u_based_decision=False): the largest-|.| entry of every component is positive."""
    j = vt.abs().argmax(dim=1, keepdim=True)
    s = torch.sign(torch.gather(vt, 1, j))
    s[s == 0] = 1
    return vt * s


def _finish_eig(w, v, dims):
    w = torch.flip(w[-dims:], dims=[0])
    vt = sign_rule(torch.flip(v[:, -dims:], dims=[1]).T.contiguous())
    return w, vt.T.contiguous()


def _cheb_filter(cov, x, degree, cut, top):
    """Scaled Chebyshev filter p(C) x (Zhou & Saad): damps the spectrum in [0, cut] (covariances are PSD), keeps the
    component at ``top`` near unit size, amplifies everything above ``cut`` like T_degree.  ``cut`` / ``top`` may be
    Python floats or 0-dim device tensors: the coefficients of the three-term recurrence are evaluated on the device
    in closed form (sigma_j = T_{j-1}(t) / T_j(t), t = (top - c) / e), so the filter never synchronises."""
    dev, dt = cov.device, cov.dtype
    cut = torch.as_tensor(cut, dtype=dt, device=dev)
    top = torch.as_tensor(top, dtype=dt, device=dev)
    e = 0.5 * cut  # centre c == half width e: the damped interval is [0, cut]
    t = (top - e) / e
    r = 1.0 / (t + torch.sqrt(t * t - 1.0))  # exp(-acosh t)
    j = torch.arange(1, degree + 1, dtype=dt, device=dev)
    sig = r * (1.0 + r ** (2.0 * (j - 1.0))) / (1.0 + r ** (2.0 * j))  # sigma_1 .. sigma_degree
    a = 2.0 * sig[1:] / e          # step j: y_{j+1} = a_j (C - c) y_j - d_j y_{j-1}
    d = sig[:-1] * sig[1:]
    cs = cov - torch.diag_embed(e.expand(cov.shape[0]))
    y = (cs @ x) * (sig[0] / e)
    for i in range(degree - 1):
        y_new = (cs @ y).mul_(a[i]).addcmul_(x, d[i], value=-1.0)
        x, y = y, y_new
    return y


def _orthonormalise(y):
    """Orthonormal basis of range(y): Householder QR.  (Cholesky-QR is cheaper but breaks down here: after a filter
    the unconverged tail columns of the block are dominated by leaked top eigenvectors and become nearly parallel.)"""
    return torch.linalg.qr(y)[0]


def _cholqr2(y):
    """Orthonormal basis of range(y) by column scaling + two Cholesky-QR passes (GEMM, a b x b Cholesky and a
    triangular solve each: a fraction of the Householder QR's latency).  Safe for a filtered block of Ritz vectors
    under eig_topk's degree rule (columns are nearly parallel to at most ~1e4 : 1, i.e. cond^2 <= 1e8); a start block
    filtered from random columns (cond ~ 1e11) needs the Householder QR.  -> (q, bad): ``bad`` is a device scalar,
    non-zero when a Cholesky factorisation broke down (the caller then repeats with Householder)."""
    y = y / y.norm(dim=0, keepdim=True)
    bad = None
    for _ in range(2):
        l, info = torch.linalg.cholesky_ex(y.T @ y)
        bad = info if bad is None else bad + info
        y = torch.linalg.solve_triangular(l, y.T, upper=False).T
    return y.contiguous(), bad


def _cheb_growth(x):
    """|T_m(x)|^(1/m) for large m: x + sqrt(x^2 - 1) (x >= 1)."""
    return x + math.sqrt(max(x * x - 1.0, 0.0))


def eig_topk(cov, dims, tol=1e-8, degree=6, max_rounds=5, stats=None):
    """K3: top-``dims`` eigenpairs of the symmetric PSD float64 matrix ``cov`` (replicated on every rank; the inputs
    are bit-identical after the integer all-reduce and the start block is seeded, so every rank gets the same
    loadings).

    Only ``dims`` << H pairs are needed, so instead of a full tridiagonalisation (cuSOLVER syevd: ~28 ms at H = 2000
    on B200 -- two thousand dependent BLAS-2 panels) this runs Chebyshev-filtered subspace iteration on a
    ``b = 2*dims + 64`` wide block: a polynomial of C (GEMMs) that damps the spectrum below the block, an
    orthonormalisation (Cholesky-QR2; Householder QR as the checked fallback), and a Rayleigh-Ritz step per round.  It stops when every kept pair has a residual
    ``|C v - lambda v| <= tol * lambda_max`` (angle to the exact eigenvector <= residual * lambda_max / eigengap:
    1e-8 keeps the angle below 1e-5 rad for relative gaps down to 1e-3; the Gram matrix itself carries a 2e-5
    relative error from the 3xTF32 tensor-core accumulation).

    Filter degree.  The wanted spectrum is wide (lambda_1 / lambda_dims ~ 50), and T_m grows like rho^m faster at
    lambda_1 than at lambda_dims (rho ~ 60 per degree here).  From a random start every column mixes all
    eigenvectors, so the weak directions survive only down to eps * rho^m: the first filter keeps m = ``degree``.
    After a Rayleigh-Ritz rotation column i is its own Ritz vector with eps-sized leakage along the strong ones; the
    filter blows that leakage up by rho^m, and the Householder QR still resolves the weak direction to
    eps * (eps * rho^m).  Each later round therefore takes the largest m with rho^m <= 1e20 (rho from the current
    Ritz values): the iteration cannot bury the weak pairs, whatever the spectrum.  Small matrices, and the (never
    observed) case of max_rounds without convergence, take the full ``eigh``."""
    h = cov.shape[0]
    b = 2 * dims + 64
    if h < 4 * b:
        w, v = torch.linalg.eigh(cov)
        return _finish_eig(w, v, dims)
    g = torch.Generator(device=cov.device)
    g.manual_seed(4466)
    q0 = torch.randn((h, b), dtype=torch.float64, device=cov.device, generator=g)
    # first filter without a Rayleigh-Ritz step: cut at the mean eigenvalue (the wanted ones lie above it), upper
    # bound from the 1-norm; both stay on the device
    trace = torch.diagonal(cov).sum()
    top0 = cov.abs().sum(dim=0).max()
    zero = torch.zeros((), dtype=torch.int32, device=cov.device)

    def start_block(householder):
        """Orthonormal basis of the filtered random block.  One degree-``degree`` filter leaves it with a condition
        number ~1e11: Householder QR territory (~1 ms of dependent panels at C2).  Filtering in steps of degree 3 with
        a Cholesky-QR2 after each keeps every step at cond <= T_3(t) = 4 t^3 - 3 t, t = (bound - cut / 2) / (cut / 2)
        (9e6 at C2, inside Cholesky-QR2's range of ~1e8); the product of the steps is nearly as sharp a filter (same
        number of rounds at C2) and all of it is GEMM-shaped.  Whether the range held is CHECKED, not assumed: the
        Cholesky status and the measured orthonormality defect travel with the first round's synchronisation, and a
        failed check repeats the start with Householder."""
        if householder:
            return _orthonormalise(_cheb_filter(cov, q0, degree, trace / h, top0)), zero, zero.to(torch.float64)
        x, bad = q0, zero
        for d in [3] * (degree // 3) + ([degree % 3] if degree % 3 else []):
            x, bd = _cholqr2(_cheb_filter(cov, x, d, trace / h, top0))
            bad = bad + bd
        defect = (x.T @ x - torch.eye(b, dtype=x.dtype, device=x.device)).abs().max()
        return x, bad, defect

    # Gate (one small synchronisation; the host waits for the Gram anyway at the end of round 1): the Cholesky route
    # only when T_3 at the upper bound of lambda_1 -- min(1-norm, Frobenius norm), which overestimates lambda_1 up to
    # ~3x, T_3 up to ~30x -- stays below 3e9.  A wrong guess costs time, never accuracy (the check above).
    bound, cut0 = torch.stack([torch.minimum(top0, torch.linalg.matrix_norm(cov)), trace / h]).tolist()
    t0 = (bound - 0.5 * cut0) / (0.5 * cut0) if cut0 > 0.0 else math.inf
    chol_start = degree >= 3 and 4.0 * t0 ** 3 <= 3e9
    q, chol_bad, start_defect = start_block(householder=not chol_start)
    if stats is not None:
        stats["eig_start"] = "cholesky-qr2" if chol_start else "householder"
    y_prev = None
    rounds = 0
    while rounds < max_rounds:
        rounds += 1
        aq = cov @ q
        t = torch.nan_to_num(q.T @ aq)  # a broken-down Cholesky-QR must reach the check below, not make eigh throw
        w, s = ops.sym_eig_small(t)  # one-CTA Jacobi for the shrunk blocks of the later rounds, library eigh else
        top, wt = s[:, -dims:], w[-dims:]
        v = q @ top
        res_t = (aq @ top - v * wt).norm(dim=0).max() / w[-1]
        # the one synchronisation of the round: residual + the Ritz values that fix the next filter
        width = int(q.shape[1])
        keep = min(width, dims + 32)
        res, th_min, th_max, th_dims, bulk, th_keep, bulk_keep, bad, defect = torch.stack(
            [res_t, w[0], w[-1], wt[0], (trace - w.sum()) / (h - width), w[-keep],
             (trace - w[-keep:].sum()) / (h - keep), chol_bad.to(torch.float64), start_defect]).tolist()
        if bad != 0.0 or res != res or not defect <= 1e-9:
            # the Cholesky-QR of this round's basis broke down (or the start block is not orthonormal): Householder,
            # same round again
            q = _orthonormalise(y_prev) if y_prev is not None else start_block(householder=True)[0]
            chol_bad, start_defect = zero, torch.zeros_like(start_defect)
            if stats is not None:
                stats["eig_householder_repeats"] = stats.get("eig_householder_repeats", 0) + 1
            rounds -= 1
            continue
        if stats is not None:
            stats["eig_rounds"], stats["eig_residual"] = rounds, res
        if res <= tol:
            return _finish_eig(wt, v, dims)
        if rounds == max_rounds or not (th_dims > 0.0 and th_max > th_dims):
            break
        # The wide start block is only needed to catch the wanted directions: once the block is rotated to Ritz
        # vectors, and if the spectrum has a gap below the wanted pairs (theta_{dims+32} < 0.8 theta_dims), the later
        # rounds keep the top dims + 32 Ritz vectors only.  The damped interval then ends at the smallest KEPT Ritz
        # value, closer to lambda_dims: faster convergence for half the GEMM / QR / eigh work.  Without such a gap
        # (wanted pairs inside a dense bulk) the full block stays.
        if keep < width and th_keep < 0.8 * th_dims:
            s = s[:, -keep:]
            th_min, bulk = th_keep, bulk_keep
        # damped interval [0, cut]: the block's smallest Ritz value, or the mean of the spectrum outside the block
        # when that is larger (it never exceeds lambda_{b+1}); kept clear of the wanted Ritz values
        cut = max(th_min, min(bulk, 0.5 * (th_min + th_dims)))
        cut = min(max(cut, 1e-3 * th_dims), 0.9 * th_dims)
        e = 0.5 * cut
        rho = _cheb_growth((th_max - e) / e) / _cheb_growth((th_dims - e) / e)
        m = int(max(2, min(32, math.floor(math.log(1e20) / math.log(max(rho, 1.0 + 1e-9))))))
        # Ritz vectors in DESCENDING order: the (Gram-Schmidt-like) orthonormalisation then takes the strong
        # directions first and removes their leakage from the weak columns, so the new basis stays aligned with the
        # eigen-directions and the next Rayleigh-Ritz matrix is nearly diagonal (few Jacobi sweeps)
        y_prev = _cheb_filter(cov, q @ torch.flip(s, dims=[1]), m, cut, th_max)
        q, chol_bad = _cholqr2(y_prev)
    w, v = torch.linalg.eigh(cov)
    if stats is not None:
        stats["eig_rounds"] = -1 - stats.get("eig_rounds", 0)
    return _finish_eig(w, v, dims)


def normalise_stats(csr, cell_idx, col_map, n_feat, comm, log_transform=True, renormalize_subset=True,
                    n_counts=None):
    """Row scalars, mu / sigma of the normalised feature matrix (graph_datastore.py:767-796) and -- from the same
    scan -- the compact matrix of normalised values that :func:`ops.hvg_dense_scale` turns into Z."""
    row_sum, row_nnz = ops.csr_row_sums(csr, cell_idx, col_map)  # scarf/assay.py:814-823
    if not renormalize_subset:
        row_sum = (n_counts[cell_idx] if cell_idx is not None else n_counts).contiguous()
    row_off, cols, xs, sx, sxx = ops.csr_hvg_compact(csr, cell_idx, col_map, n_feat, row_sum, row_nnz, SF,
                                                     log_transform)
    n_local = csr.n_rows if cell_idx is None else int(cell_idx.numel())
    n = n_local
    if comm.world > 1:
        cnt = torch.tensor([n_local], dtype=torch.int64, device=csr.device)
        comm.allreduce_sum_(sx), comm.allreduce_sum_(sxx), comm.allreduce_sum_(cnt)
        n = int(cnt.item())
    n = float(n)
    scale = 2.0 ** -lib.COLSTAT_SHIFT
    mean = sx.to(torch.float64) * scale / n
    var = torch.clamp(sxx.to(torch.float64) * scale / n - mean * mean, min=0.0)
    mu = clean_array(mean)
    sigma = clean_array(torch.sqrt(var), 1.0)
    return (row_off, cols, xs), mu, sigma, int(n)


def make_graph_csr(csr: CsrDevice, cell_idx, feat_mask, dims=11, k=11, lc=1.0, bw=1.5, batch_size=1000,
                   log_transform=True, renormalize_subset=True, n_counts=None, comm: Comm | None = None,
                   gram_mode=0, knn_method=0, loadings=None, mu=None, sigma=None, timers=None,
                   pca_rows=None, stats=None) -> GraphResult:
    """normalise -> mu/sigma -> Z -> Gram -> eig -> project -> exact kNN -> edge weights for this rank's rows.

    ``timers``: optional list; (stage name, torch.cuda.Event) pairs are appended at every stage boundary.
    ``stats``: optional dict; receives the eigensolver's route / rounds / residual (see :func:`eig_topk`).
    ``cell_idx``: int64 device tensor of the local CSR rows to use (None = all).  ``feat_mask``: bool over all
    genes (numpy), e.g. from :func:`mark_hvgs_csr`.  With ``comm.world > 1`` every rank passes its own shard and
    receives its own rows of the global graph; neighbour ids are global selected-row ids.
    ``pca_rows``: optional int64 device tensor of positions among this rank's selected rows; the PCA is then fitted on
    those cells only (``pca_cell_key``, scarf/datastore/graph_datastore.py:764 -> AnnStream._fit_pca ``use_for_pca``,
    scarf/ann.py:215-228) -- z-scaling still uses the mu / sigma of all selected cells, the covariance is centred on
    the subset's own mean like the PCA estimator does -- while every selected cell is projected and searched.
    """
    comm = comm or Comm()
    dev = csr.device

    def mark(name):  # CUDA events on the launching stream; read by the caller after its own synchronize
        if timers is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timers.append((name, e))

    mark("start")
    mask_t = (feat_mask if torch.is_tensor(feat_mask) else torch.from_numpy(np.asarray(feat_mask, dtype=bool))).to(dev)
    rank_t = torch.cumsum(mask_t, dim=0, dtype=torch.int32)
    col_map = torch.where(mask_t, rank_t - 1, torch.full_like(rank_t, -1)).contiguous()
    feat_idx_t = torch.nonzero(mask_t).flatten()
    n_feat = int(feat_idx_t.numel())
    if n_feat == 0:
        raise ValueError("make_graph: no features selected")
    n_local = csr.n_rows if cell_idx is None else int(cell_idx.numel())

    # ---- normalisation scalars, mu, sigma (collective 2a) ----
    compact, mu_d, sigma_d, n_total = normalise_stats(csr, cell_idx, col_map, n_feat, comm, log_transform,
                                                      renormalize_subset, n_counts)
    if mu is not None:  # run_mapping reuses the reference's mu / sigma
        mu_d, sigma_d = mu, sigma
    counts = comm.allgather_counts(n_local, dev)
    row_offset = int(sum(counts[: comm.rank]))
    k = min(k, n_total - 1)  # scarf/ann.py:85-86
    dims = clamp_dims(dims, n_total, batch_size)
    mark("stats")

    # ---- Z (K1) ----
    ldz = round_up(n_feat, 128)
    z = torch.empty((max(n_local, 1), ldz), dtype=torch.float32, device=dev)
    z_lo = torch.empty_like(z) if (gram_mode == 3 and loadings is None) else None
    ops.hvg_dense_scale(*compact, n_feat, z, mu_d, sigma_d, z_lo=z_lo)
    del compact
    mark("normalise")

    # ---- PCA: Gram (K2, collective 2b) + eigensolve (K3) ----
    if loadings is None:
        if pca_rows is None:
            g_fx = ops.gram_accumulate(z, n_local, n_feat, mode=gram_mode, z_lo=z_lo)
            n_pca, col_mean = n_total, None
        else:  # PCA on a subset of the cells: Gram of the gathered rows, centred on their own mean
            zs = z.index_select(0, pca_rows)
            zs_lo = z_lo.index_select(0, pca_rows) if z_lo is not None else None
            g_fx = ops.gram_accumulate(zs, int(pca_rows.numel()), n_feat, mode=gram_mode, z_lo=zs_lo)
            moments = torch.cat([zs[:, :n_feat].sum(dim=0, dtype=torch.float64),
                                 torch.tensor([float(pca_rows.numel())], dtype=torch.float64, device=dev)])
            comm.allreduce_sum_(moments)
            n_pca = int(round(float(moments[-1].item())))
            col_mean = moments[:-1] / max(n_pca, 1)
            dims = clamp_dims(dims, n_pca, batch_size)
            del zs, zs_lo
        comm.allreduce_sum_(g_fx)
        ops.gram_symmetrize(g_fx, n_feat)
        mark("gram")
        cov = g_fx[:n_feat, :n_feat].to(torch.float64) * 2.0 ** -lib.GRAM_SHIFT
        if col_mean is not None:
            cov = cov - float(n_pca) * torch.outer(col_mean, col_mean)
        cov = cov / max(n_pca - 1, 1)
        evals, load = eig_topk(cov, dims, stats=stats)
        mark("eig")
    else:
        load, evals = loadings, torch.zeros(dims, dtype=torch.float64, device=dev)
    ldv = round_up(dims, 4)
    v32 = torch.zeros((n_feat, ldv), dtype=torch.float32, device=dev)
    v32[:, :dims] = load.to(torch.float32)

    # ---- embedding (K4) + exchange (collective 3) ----
    y = ops.project(z, n_local, n_feat, v32, dims, z_lo=z_lo)  # tensor cores when the 3xTF32 plane exists
    del z, z_lo
    y_all = comm.allgather_rows(y, counts)
    mark("project")

    # ---- exact kNN (K5) ----
    kev = None
    if timers is not None and knn_method == 1:  # the tensor kernel alone, for the roofline of bench.py
        kev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        kev[0].record(), kev[1].record()  # creates the handles; the library records them again around the kernel
        timers.append(("knn_kernel_events", kev))
    idx, dist = ops.knn_l2(y, y_all, dims, k, self_offset=row_offset, method=knn_method, kernel_events=kev)
    mark("knn")

    # ---- edge weights (K6, collective 4) ----
    edges, weights = smoothen_dists(idx, dist, lc, bw, row_offset, batch_size, n_total, comm)
    mark("weights")

    return GraphResult(n_total, row_offset, feat_idx_t.cpu().numpy(), mu_d, sigma_d, load, evals, y, y_all, dims, k,
                       idx, dist, edges, weights)


def smoothen_dists(idx, dist, lc, bw, row_offset, chunk_size, n_total, comm: Comm | None = None):
    """scarf/knn_utils.py:89-159 on device for rows [row_offset, row_offset + n) -> (edges, weights)."""
    comm = comm or Comm()
    n, k = idx.shape
    n_chunks = (n_total + chunk_size - 1) // chunk_size
    csum = ops.chunk_sums(dist, row_offset, chunk_size, n_chunks)
    comm.allreduce_sum_(csum)
    rows_in_chunk = torch.full((n_chunks,), float(chunk_size), dtype=torch.float64, device=idx.device)
    rows_in_chunk[-1] = float(n_total - (n_chunks - 1) * chunk_size)
    cmean = (csum / (rows_in_chunk * k)).to(torch.float32)
    sigma, rho = ops.smooth_knn(dist, cmean, lc, bw, row_offset, chunk_size)
    edges, weights, cmin, czero = ops.membership_coo(idx, dist, sigma, rho, row_offset, chunk_size, n_chunks)
    comm.allreduce_min_(cmin), comm.allreduce_max_(czero)
    sel = czero > 0
    if bool(sel.any()):  # zero weights := min(1, smallest non-zero weight of the chunks that hold a zero)
        floor = min(1.0, float(cmin[sel].min()))
        ops.fill_zero_weights(weights, floor)
    return edges, weights


# =============================================================================================
# run_mapping
# =============================================================================================
@dataclass
class MappingResult:
    indices: torch.Tensor     # int64 [nq_local, save_k] ids of reference cells (global selected-row ids)
    distances: torch.Tensor   # float32 [nq_local, save_k] squared L2
    embedding: torch.Tensor   # float32 [nq_local, ld] target cells in the reference's PCA space
    mu: torch.Tensor
    sigma: torch.Tensor


def run_mapping_csr(target: CsrDevice, target_cell_idx, t_col_of_feature, ref_mu, ref_sigma, ref_loadings,
                    ref_embedding_all, dims, save_k=3, use_ref_mu=True, use_ref_sigma=True, log_transform=True,
                    renormalize_subset=True, n_counts=None, comm: Comm | None = None) -> MappingResult:
    """MappingDatastore.run_mapping's numeric core (scarf/datastore/mapping_datastore.py:118-208 with
    mapping_utils.align_features :148-214) for this rank's target cells.

    ``t_col_of_feature``: int array [H], for every feature of the reference graph (source order) its column in the
    target matrix or -1 when the target lacks it (``_order_features``, exclude_missing=False); such features get
    the value 1.0 before z-scaling (mapping_utils.py:208).  The target is normalised with the reference's
    ``subset_params`` over the features it shares with the source.  ``ref_embedding_all`` is the reference embedding
    of ALL reference cells (every rank holds it after make_graph's all-gather); no self handling in the query.
    ``use_ref_mu`` / ``use_ref_sigma`` False: z-scale with the target's own column mean / std instead
    (mapping_datastore.py:177-190)."""
    comm = comm or Comm()
    dev = target.device
    t_col = np.asarray(t_col_of_feature, dtype=np.int64)
    n_feat = int(t_col.size)
    present = t_col >= 0
    if not present.any():
        raise ValueError("ERROR: None of the features from reference were found in the target data")
    cm = np.full(target.n_cols, -1, dtype=np.int32)
    cm[t_col[present]] = np.where(present)[0].astype(np.int32)
    col_map = torch.from_numpy(cm).to(dev)
    fill = np.where(present, np.nan, 1.0)
    missing_fill = torch.from_numpy(fill).to(dev) if (~present).any() else None
    n_local = target.n_rows if target_cell_idx is None else int(target_cell_idx.numel())
    if renormalize_subset:
        row_sum, _ = ops.csr_row_sums(target, target_cell_idx, col_map)
    else:
        row_sum = (n_counts[target_cell_idx] if target_cell_idx is not None else n_counts).contiguous()
    mu_d, sigma_d = ref_mu, ref_sigma
    if not (use_ref_mu and use_ref_sigma):
        sx, sxx = ops.csr_hvg_colstats(target, target_cell_idx, col_map, n_feat, row_sum, SF, log_transform)
        n = n_local
        if comm.world > 1:
            cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
            comm.allreduce_sum_(sx), comm.allreduce_sum_(sxx), comm.allreduce_sum_(cnt)
            n = int(cnt.item())
        scale = 2.0 ** -lib.COLSTAT_SHIFT
        mean = sx.to(torch.float64) * scale / n
        var = torch.clamp(sxx.to(torch.float64) * scale / n - mean * mean, min=0.0)
        miss = torch.from_numpy(~present).to(dev)
        mean = torch.where(miss, torch.ones_like(mean), mean)  # a column of ones: mean 1, std 0
        var = torch.where(miss, torch.zeros_like(var), var)
        if not use_ref_mu:
            mu_d = clean_array(mean)
        if not use_ref_sigma:
            sigma_d = clean_array(torch.sqrt(var), 1.0)
    ldz = round_up(n_feat, 128)
    z = torch.empty((max(n_local, 1), ldz), dtype=torch.float32, device=dev)
    z_lo = torch.empty_like(z)  # second 3xTF32 plane: the projection runs on the tensor cores
    ops.csr_norm_scale(target, target_cell_idx, col_map, n_feat, row_sum, z, SF, log_transform, mu_d, sigma_d,
                       missing_fill=missing_fill, z_lo=z_lo)
    ldv = round_up(dims, 4)
    v32 = torch.zeros((n_feat, ldv), dtype=torch.float32, device=dev)
    v32[:, :dims] = ref_loadings[:, :dims].to(torch.float32)
    y = ops.project(z, n_local, n_feat, v32, dims, ldy=int(ref_embedding_all.stride(0)), z_lo=z_lo)
    del z, z_lo
    save_k = min(save_k, int(ref_embedding_all.shape[0]))
    idx, dist = ops.knn_l2(y, ref_embedding_all, dims, save_k, self_offset=-1, method=1)
    return MappingResult(idx, dist, y, mu_d, sigma_d)


# =============================================================================================
# small pieces of the AnnStream contract
# =============================================================================================
def graph_to_sparse(edges, weights, n_cells, k, use_k=None, symmetric=None, upper_only=None):
    """The stored COO graph as a scipy CSR matrix: ``_store_to_sparse`` (scarf/datastore/graph_datastore.py:474-511:
    ``use_k`` clamped into [1, k], the first ``use_k`` of every row's k entries kept) and ``load_graph``'s
    symmetrisation ``g + g.T - g * g.T`` with the optional upper triangle (graph_datastore.py:1052-1075)."""
    from scipy.sparse import csr_matrix, triu

    use_k = k if use_k is None else min(max(int(use_k), 1), k)
    if use_k != k:
        keep = np.tile([True] * use_k + [False] * (k - use_k), n_cells)
        edges, weights = edges[keep], weights[keep]
    g = csr_matrix((weights, (edges[:, 0].astype(np.int64), edges[:, 1].astype(np.int64))), shape=(n_cells, n_cells))
    if symmetric is True:
        g = g + g.T - g.multiply(g.T)
        if upper_only is True:
            g = triu(g)
    return g


def order_features(source_ids, target_ids, source_feat_idx) -> np.ndarray:
    """``_order_features`` with its defaults exclude_missing=False / filter_null=False (scarf/mapping_utils.py:98-145):
    for every source feature the graph was built on (``source_feat_idx``: ascending positions in the source feature
    table), its column in the target matrix, matched by feature id, or -1 when the target lacks it."""
    pos = {v: i for i, v in enumerate(target_ids)}
    t_col = np.array([pos.get(source_ids[i], -1) for i in source_feat_idx], dtype=np.int64)
    if t_col.size == 0 or np.all(t_col == -1):
        raise ValueError("ERROR: None of the features from reference were found in the target data")
    return t_col


def fix_knn_query(indices: np.ndarray, distances: np.ndarray, ref_idx: np.ndarray):
    """scarf/ann.py:31-52: drop each query's own hit from a (k+1)-neighbour result -- column 0 when it is the query
    itself, else wherever the query is found, else the last column.  Returns (indices, distances, n_not_first)."""
    n, k1 = indices.shape
    first = indices[:, 0] != ref_idx
    pos = np.zeros(n, dtype=np.int64)
    hit = indices == ref_idx[:, None]
    found = hit.any(axis=1)
    pos[found] = hit[found].argmax(axis=1)
    pos[~found] = k1 - 1
    keep = np.ones((n, k1), dtype=bool)
    keep[np.arange(n), pos] = False
    return indices[keep].reshape(n, k1 - 1), distances[keep].reshape(n, k1 - 1), int(first.sum())


def fit_kmeans(embedding_all, dims, n_clusters, rand_state=4466, n_iter=10):
    """The ``kmeans__<n>__<seed>`` arrays make_graph always writes (scarf/ann.py:328-346, read back by run_umap /
    run_tsne for initialisation, graph_datastore.py:427-457).  The reference fits sklearn MiniBatchKMeans; its result
    is not pinned by any reference test, so this is a deterministic Lloyd iteration on the GPU: seeded choice of
    start rows, assignment = exact 1-nearest-centre search with the kNN kernel, update = segment means.  Every rank
    holds the full embedding and gets identical centres and labels.  -> (centres float32 [n_clusters, dims],
    labels int64 [n_cells])."""
    n = int(embedding_all.shape[0])
    n_clusters = int(max(2, min(n_clusters, n)))
    dev = embedding_all.device
    g = torch.Generator(device="cpu")
    g.manual_seed(int(rand_state))
    start = torch.sort(torch.randperm(n, generator=g)[:n_clusters]).values.to(dev)
    ld = int(embedding_all.stride(0))
    centres = embedding_all[start].clone()  # [n_clusters, ld], pad columns zero
    labels = None
    for it in range(n_iter + 1):
        idx, _ = ops.knn_l2(embedding_all, centres, dims, 1, self_offset=-1, method=1)
        labels = idx[:, 0]
        if it == n_iter:
            break
        sums = torch.zeros((n_clusters, ld), dtype=torch.float64, device=dev)
        sums.index_add_(0, labels, embedding_all.to(torch.float64))
        cnt = torch.bincount(labels, minlength=n_clusters).to(torch.float64)
        new = (sums / cnt.clamp(min=1.0)[:, None]).to(torch.float32)
        centres = torch.where((cnt > 0)[:, None], new, centres).contiguous()  # an empty cluster keeps its centre
    return centres[:, :dims].contiguous(), labels
