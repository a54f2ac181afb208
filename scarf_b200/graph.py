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
    if ldz * 24 > 227 * 1024:
        # scf_hvg_dense_scale / scf_csr_norm_scale keep mu, sigma and one output row in shared memory (24 B per
        # column): ~9,600 features.  The path is built for HVG-sized feature sets (the H x H Gram grows quadratically too)
        raise NotImplementedError(f"make_graph: {n_feat} features selected; the GPU path handles up to "
                                  f"{227 * 1024 // 24 // 128 * 128} (use a feature selection, e.g. mark_hvgs, instead of "
                                  "feat_key='I' on a whole transcriptome)")
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
        # K3, native (scf_eig_topk): cov = G 2^-shift / (n - 1) [- n / (n - 1) mean mean^T for a PCA on a cell subset]
        den = float(max(n_pca - 1, 1))
        evals, load, v32 = ops.eig_topk(g_fx, n_feat, dims, 2.0 ** -lib.GRAM_SHIFT / den, col_mean,
                                        float(n_pca) / den if col_mean is not None else 0.0, stats=stats,
                                        ld32=round_up(dims, 4))
        mark("eig")
    else:
        load, evals = loadings, torch.zeros(dims, dtype=torch.float64, device=dev)
        v32 = torch.zeros((n_feat, round_up(dims, 4)), dtype=torch.float32, device=dev)
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
def graph_to_sparse(edges, weights, n_cells, k, use_k=None, symmetric=None, upper_only=None, device=None):
    """The stored COO graph as a scipy CSR matrix: ``_store_to_sparse`` (scarf/datastore/graph_datastore.py:474-511:
    ``use_k`` clamped into [1, k], the first ``use_k`` of every row's k entries kept) and ``load_graph``'s
    symmetrisation ``g + g.T - g * g.T`` with the optional upper triangle (graph_datastore.py:1052-1075).  With a CUDA
    ``device`` the symmetrisation arithmetic runs in ``scf_graph_symmetrize`` (the stored layout -- row i owns entries
    [i k, (i + 1) k) -- is what the kernel needs); scipy then only assembles the container.  ``device=None`` (a store
    read without a GPU, the CPU tests) evaluates the same expression with scipy."""
    from scipy.sparse import csr_matrix, triu

    use_k = k if use_k is None else min(max(int(use_k), 1), k)
    if symmetric is True and device is not None and torch.device(device).type == "cuda":
        dev = torch.device(device)
        idx = torch.from_numpy(np.ascontiguousarray(edges[:, 1].astype(np.int64)).reshape(n_cells, k)).to(dev)
        w = torch.from_numpy(np.ascontiguousarray(np.asarray(weights, dtype=np.float64)).reshape(n_cells, k)).to(dev)
        rows, cols, vals = ops.graph_symmetrize(idx, w, use_k, upper_only is True)
        rows, cols, vals = rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy()
        keep = rows >= 0
        return csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n_cells, n_cells))
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
    is not pinned by any reference test, so this is a deterministic k-means on the GPU: k-means++ seeding (D^2
    sampling by inverse CDF with host-seeded uniforms, so every rank draws the same rows), then Lloyd iterations --
    assignment = exact 1-nearest-centre search with the kNN kernel, update = segment means.  Every rank holds the
    full embedding and gets identical centres and labels.  -> (centres float32 [n_clusters, dims], labels int64
    [n_cells])."""
    n = int(embedding_all.shape[0])
    n_clusters = int(max(2, min(n_clusters, n)))
    dev = embedding_all.device
    ld = int(embedding_all.stride(0))
    rng = np.random.default_rng(int(rand_state))
    trials = 2 + int(math.log(n_clusters))  # greedy k-means++ (Arthur & Vassilvitskii; sklearn's n_local_trials)
    u = torch.from_numpy(rng.random((n_clusters, trials))).to(dev)  # host-seeded uniforms: the same draws on every rank
    # Seeding runs on a seeded subsample, as MiniBatchKMeans does (init_size = 3 * batch_size, at least 3 * n_clusters;
    # four times that here): n_clusters sequential D^2-sampling steps over all cells cost 0.6 s at 100k cells x 1,000
    # centres, the Lloyd iterations below -- on all cells -- are what fixes the centres.
    m = min(n, 4 * max(3 * n_clusters, 3072))
    sub = (torch.from_numpy(np.sort(rng.choice(n, size=m, replace=False))).to(dev) if m < n
           else torch.arange(n, dtype=torch.int64, device=dev))
    y64 = embedding_all[sub, :dims].to(torch.float64)
    start = torch.empty(n_clusters, dtype=torch.int64, device=dev)  # positions in the subsample
    start[0] = int(float(u[0, 0].item()) * m) % m
    mind2 = ((y64 - y64[start[0]]) ** 2).sum(dim=1)
    norm2 = (y64 * y64).sum(dim=1)
    # `batch` centres per step (1 up to 128 centres: plain greedy k-means++): the loop is bound by the ~15 small launches
    # of a step, not by their arithmetic.  Every centre of a batch is the best of its own `trials` draws against the
    # centres chosen in the steps before.
    batch = max(1, n_clusters // 128)
    c = 1
    while c < n_clusters:
        b = min(batch, n_clusters - c)
        cdf = torch.cumsum(mind2, dim=0)
        cand = torch.searchsorted(cdf, (u[c:c + b] * cdf[-1]).reshape(-1)).clamp(max=m - 1)  # D^2 sampling, b * trials draws
        d2 = torch.addmm(norm2[:, None] + norm2[cand][None, :], y64, y64[cand].T, alpha=-2.0).clamp_(min=0.0)  # [m, b * trials]
        d2[cand, torch.arange(b * trials, device=dev)] = 0.0  # a candidate is at distance zero from itself, exactly
        d2 = torch.minimum(d2, mind2[:, None]).reshape(m, b, trials)
        best = torch.argmin(d2.sum(dim=0), dim=1)  # per centre: the draw that lowers the potential most (first on ties)
        start[c:c + b] = cand.reshape(b, trials).gather(1, best[:, None])[:, 0]
        mind2 = d2.gather(2, best[None, :, None].expand(m, b, 1))[:, :, 0].min(dim=1).values.contiguous()
        c += b
    start = sub[start]
    if n_clusters == n:  # one centre per cell: the sampling above cannot draw a row twice only while mind2 > 0
        start = torch.arange(n, dtype=torch.int64, device=dev)
    centres = embedding_all[torch.sort(start).values].clone()  # [n_clusters, ld], pad columns zero
    labels = None
    emb64 = embedding_all.to(torch.float64)
    for it in range(n_iter + 1):
        idx, _ = ops.knn_l2(embedding_all, centres, dims, 1, self_offset=-1, method=1)
        labels = idx[:, 0]
        if it == n_iter:
            break
        # segment means in a fixed order (no floating-point atomics: every rank must get the same bits): rows sorted by
        # label (stable), then one sequential float64 sum per (cluster, column) -- torch.segment_reduce gives every
        # output element to one thread.
        order = torch.argsort(labels, stable=True)
        cnt_i = torch.bincount(labels, minlength=n_clusters)
        sums = torch.segment_reduce(emb64[order], "sum", lengths=cnt_i, axis=0, unsafe=True)
        cnt = cnt_i.to(torch.float64)
        new = (sums / cnt.clamp(min=1.0)[:, None]).to(torch.float32)
        centres = torch.where((cnt > 0)[:, None], new, centres).contiguous()  # an empty cluster keeps its centre
    return centres[:, :dims].contiguous(), labels
