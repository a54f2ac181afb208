"""float64 CPU restatement of Scarf's ``mark_hvgs`` + ``make_graph`` arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines (relative to /root/reference) it restates.  Inputs are
``scipy.sparse.csr_matrix`` raw counts (what ``Assay.to_raw_sparse`` returns,
scarf/assay.py:175-199); dense intermediates are float64 like the reference's.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np
import scipy.sparse as sp

from .lowess import lowess

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_BLACKLIST = "^MT-|^RPS|^RPL|^MRPS|^MRPL|^CCN|^HLA-|^H2-|^HIST"  # scarf/datastore/datastore.py:235


# ----------------------------------------------------------------------------------------------
# A.1  per-cell totals / per-gene nCells              scarf/datastore/base_datastore.py:345-366
# ----------------------------------------------------------------------------------------------
def cell_totals(counts: sp.csr_matrix):
    """``nCounts`` (sum over all genes) and ``nFeatures`` (genes > 0), stored as f8."""
    counts = counts.tocsr()
    n_counts = np.asarray(counts.sum(axis=1)).ravel().astype(np.float64)
    n_feats = np.asarray((counts > 0).sum(axis=1)).ravel().astype(np.float64)
    return n_counts, n_feats


def gene_ncells(counts: sp.csr_matrix):
    """``nCells`` per feature over ALL stored cells (scarf/assay.py:201-225)."""
    return np.asarray((counts > 0).sum(axis=0)).ravel()


# ----------------------------------------------------------------------------------------------
# A.2  HVG statistics                                            scarf/assay.py:830-897,41-51
# ----------------------------------------------------------------------------------------------
def gene_stats(counts: sp.csr_matrix, cell_idx, feat_idx, n_counts, n_cells_total, sf=1000):
    """Per-gene nnz / sum / population variance of ``sf*c/nCounts`` over ``cell_idx`` rows.

    ``avg`` divides by the TOTAL number of cells in the store (scarf/assay.py:874), not by
    ``len(cell_idx)``.  Returns dict of float64 arrays of length ``len(feat_idx)``.
    """
    sub = counts[cell_idx][:, feat_idx].astype(np.float64).tocsr()
    scal = sp.diags(sf / n_counts[cell_idx])
    v = (scal @ sub).tocsc()
    n = np.asarray((v > 0).sum(axis=0)).ravel().astype(np.float64)
    tot = np.asarray(v.sum(axis=0)).ravel()
    m = len(cell_idx)
    mean = tot / m
    # population variance, two-pass like dask/numpy var(ddof=0) on the dense block
    v2 = v.copy()
    v2.data = v2.data ** 2
    ex2 = np.asarray(v2.sum(axis=0)).ravel() / m
    # two-pass refinement: sum((x-mean)^2) = sum_nz (x-mean)^2 + (m-nnz)*mean^2
    csc = v
    d = csc.data - np.repeat(mean, np.diff(csc.indptr))
    ssd = np.add.reduceat(np.append(d * d, 0.0), np.minimum(csc.indptr[:-1], len(d)))
    ssd[np.diff(csc.indptr) == 0] = 0.0
    var = (ssd + (m - np.diff(csc.indptr)) * mean ** 2) / m
    del ex2
    nz_mean = np.divide(tot, n, out=np.zeros_like(tot), where=n != 0)
    return {"normed_n": n, "normed_tot": tot, "sigmas": var, "avg": tot / n_cells_total, "nz_mean": nz_mean}


# ----------------------------------------------------------------------------------------------
# A.3  trend removal                       scarf/metadata.py:586-617, scarf/feat_utils.py:11-45
# ----------------------------------------------------------------------------------------------
def fit_lowess(a, b, n_bins=200, lowess_frac=0.1):
    la, lb = np.log(np.asarray(a, dtype=np.float64)), np.log(np.asarray(b, dtype=np.float64))
    edges = np.histogram(la, bins=n_bins)[1]
    edges[-1] += 0.1
    bin_members, bx, by = [], [], []
    for i in range(n_bins):
        idx = np.where((la >= edges[i]) & (la < edges[i + 1]))[0]
        if len(idx) == 0:
            continue
        g = idx[np.argmin(lb[idx])]  # first minimum, like pandas idxmin
        bin_members.append(idx)
        by.append(lb[g])
        bx.append(la[g])
    fit = lowess(np.array(by), np.array(bx), frac=lowess_frac, it=100)
    out = np.zeros(len(la))
    for f, idx in zip(fit, bin_members):
        out[idx] = np.e ** (lb[idx] - f)
    return out


def remove_trend(avg, sigmas, n_bins=200, lowess_frac=0.1, fill_value=0.0):
    a, b = np.asarray(avg, dtype=float), np.asarray(sigmas, dtype=float)
    idx = a > 0
    ret = np.repeat(float(fill_value), len(a))
    ret[idx] = fit_lowess(a[idx], b[idx], n_bins, lowess_frac)
    return ret


# ----------------------------------------------------------------------------------------------
# A.4  HVG choice                    scarf/assay.py:1014-1063, scarf/datastore/datastore.py:282-314
# ----------------------------------------------------------------------------------------------
def choose_hvgs(normed_n, nz_mean, c_var, feat_I, gene_names=None, top_n=500, min_cells=0, max_cells=np.inf,
                min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf, blacklist=DEFAULT_BLACKLIST,
                keep_bounds=False):
    """The HVG choice of RNAassay.mark_hvgs (scarf/assay.py:1014-1063) on full-length per-gene vectors: log2 bounds
    exponentiated when finite, blacklist through MetaData.grep (names and pattern upper-cased, re.match;
    scarf/metadata.py:569-584), multi_sift bounds (strict unless ``keep_bounds``; metadata.py:483-533), and either
    the (top_n + 1)-th largest eligible corrected variance or ``2**min_var`` as the threshold.  Pinned on the
    reference method executed on a stub assay (tests/golden/make_ref_function_goldens.py)."""
    G = len(normed_n)
    if max_mean != np.inf:
        max_mean = 2 ** max_mean
    if max_var != np.inf:
        max_var = 2 ** max_var
    if min_mean != -np.inf:
        min_mean = 2 ** min_mean
    if min_var != -np.inf:
        min_var = 2 ** min_var
    if blacklist != "" and gene_names is not None:
        pat = re.compile(blacklist.upper())
        bl = np.array([pat.match(str(x).upper()) is None for x in gene_names])
    else:
        bl = np.ones(G, dtype=bool)

    def sift(v, lo, hi):
        return (v >= lo) & (v <= hi) if keep_bounds else (v > lo) & (v < hi)

    with np.errstate(invalid="ignore"):
        if min_var == -np.inf:
            if top_n < 1:
                raise ValueError("ERROR: Please provide a value greater than 0 for `top_n` parameter")
            idx = sift(normed_n, min_cells, max_cells) & sift(nz_mean, min_mean, max_mean) & feat_I & bl
            n_valid = idx.sum()
            if top_n > n_valid:
                top_n = n_valid - 1
            min_var = np.sort(c_var[idx])[::-1][top_n]
        hvgs = sift(normed_n, min_cells, max_cells) & sift(nz_mean, min_mean, max_mean) & sift(c_var, min_var, max_var)
        return hvgs & feat_I & bl


def mark_hvgs(counts, cell_idx, feat_I, gene_names=None, top_n=500, min_cells=None, max_cells=np.inf,
              n_bins=200, lowess_frac=0.1, blacklist=DEFAULT_BLACKLIST, n_counts=None, return_stats=False):
    """Boolean HVG mask over all genes.  ``feat_I`` is the feature ``I`` column (bool, all genes)."""
    n_total = counts.shape[0]
    if n_counts is None:
        n_counts, _ = cell_totals(counts)
    if min_cells is None:
        min_cells = int(0.01 * n_total)
    feat_idx = np.where(feat_I)[0]
    st = gene_stats(counts, cell_idx, feat_idx, n_counts, n_total)
    c_var_I = remove_trend(st["avg"], st["sigmas"], n_bins, lowess_frac)
    G = counts.shape[1]

    def full(v):
        o = np.full(G, np.nan)
        o[feat_idx] = v
        return o

    normed_n, nz_mean, c_var = full(st["normed_n"]), full(st["nz_mean"]), full(c_var_I)
    hvgs = choose_hvgs(normed_n, nz_mean, c_var, feat_I, gene_names, top_n, min_cells, max_cells, blacklist=blacklist)
    if return_stats:
        return hvgs, {"c_var": c_var, "normed_n": normed_n, "nz_mean": nz_mean, **{k: full(v) for k, v in st.items()}}
    return hvgs


# ----------------------------------------------------------------------------------------------
# A.5  normalise, mu/sigma, PCA
# ----------------------------------------------------------------------------------------------
def normed_hvg(counts, cell_idx, feat_idx, sf=1000, log_transform=True, renormalize_subset=True, n_counts=None):
    """Dense f8 (cells x HVGs): scarf/assay.py:814-826 with norm_lib_size_log (assay.py:54-64)."""
    sub = np.asarray(counts[cell_idx][:, feat_idx].todense(), dtype=np.float64)
    if renormalize_subset:
        scalar = sub.sum(axis=1)
        scalar[scalar == 0] = 1
    else:
        scalar = n_counts[cell_idx]
    val = sf * sub / scalar.reshape(-1, 1)
    return np.log1p(val) if log_transform else val


def clean_array(x, fill_val=0):
    """scarf/utils.py:143-153."""
    x = np.nan_to_num(x, copy=True)
    x[(x == np.inf) | (x == -np.inf)] = 0
    x[x == 0] = fill_val
    return x


def mu_sigma(x):
    """scarf/datastore/graph_datastore.py:767-796 (dask mean/std, ddof 0)."""
    return clean_array(x.mean(axis=0)), clean_array(x.std(axis=0), 1)


def clamp_dims(dims, n_cells, batch_size):
    """scarf/ann.py:173-185."""
    if dims > n_cells:
        dims = n_cells
    if dims >= batch_size:
        dims = batch_size - 1
    return dims


def ipca_loadings(x, mu, sigma, dims, batch_size=1000):
    """Scarf's IncrementalPCA fit (scarf/ann.py:207-256): z-scaled blocks fed 1,2,..,n-1 then
    block 0 (the ``end_reservoir``) plus carry-over, ``n_components=dims+1``, last one dropped."""
    from sklearn.decomposition import IncrementalPCA

    pca = IncrementalPCA(n_components=dims + 1, batch_size=batch_size)
    end_reservoir, carry = [], []
    for s in range(0, x.shape[0], batch_size):
        i = (x[s:s + batch_size] - mu) / sigma
        if len(carry) > 0:
            i = np.vstack((carry, i))
            carry = []
        if len(i) < dims + 1:
            carry = i
            continue
        if len(end_reservoir) == 0:
            end_reservoir = i
            continue
        pca.partial_fit(i, check_input=False)
    i = np.vstack((end_reservoir, carry)) if len(carry) > 0 else end_reservoir
    pca.partial_fit(i, check_input=False)
    return pca.components_[:-1, :].T.copy()


def sign_rule(vt):
    """sklearn ``svd_flip(u_based_decision=False)``: each component is signed so that its
    largest-|.| entry is positive (sklearn/utils/extmath.py svd_flip)."""
    vt = np.array(vt, dtype=np.float64, copy=True)
    j = np.argmax(np.abs(vt), axis=1)
    s = np.sign(vt[np.arange(vt.shape[0]), j])
    s[s == 0] = 1
    return vt * s[:, None]


def exact_pca_loadings(z, dims):
    """Exact-PCA equivalent: top-``dims`` eigenvectors of the covariance of Z
    with the sklearn sign rule.  Returns (loadings HxD, eigenvalues of ZtZ-centred/(n-1))."""
    n = z.shape[0]
    m = z.mean(axis=0)
    c = (z.T @ z - n * np.outer(m, m)) / (n - 1)
    w, v = np.linalg.eigh(c)
    order = np.argsort(w)[::-1][:dims]
    return sign_rule(v[:, order].T).T.copy(), w[order]


# ----------------------------------------------------------------------------------------------
# A.6  exact kNN  (defines what "bit-exact indices" means; replaces hnswlib l2, scarf/ann.py:14-52)
# ----------------------------------------------------------------------------------------------
_knn_lib = None


def _load_knn_lib():
    global _knn_lib
    if _knn_lib is None:
        so = os.path.join(_HERE, "_build", "liboracle.so")
        src = os.path.join(_HERE, "oracle_c.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build_c()
        _knn_lib = ctypes.CDLL(so)
        _knn_lib.oracle_knn_exact.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int32]
        _knn_lib.oracle_knn_exact.restype = ctypes.c_int32
    return _knn_lib


def build_c():
    os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o",
                           os.path.join(_HERE, "_build", "liboracle.so"), os.path.join(_HERE, "oracle_c.c"), "-lm"])


def exact_knn(queries, refs, k, self_offset=0, nthreads=0):
    """k nearest refs per query by ``d = float32(sum_t (double(a_t)-double(b_t))**2)`` (sequential
    in t), ordered by (d, index).  ``self_offset >= 0``: query i is ref ``i+self_offset`` and is
    excluded (what fix_knn_query leaves, scarf/ann.py:31-52); ``-1``: nothing excluded
    (run_mapping, scarf/mapping_datastore.py:198-208).  Returns (uint64 idx, float32 sq-dist)."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    r = np.ascontiguousarray(refs, dtype=np.float32)
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    dist = np.empty((q.shape[0], k), dtype=np.float32)
    lib = _load_knn_lib()
    rc = lib.oracle_knn_exact(q.ctypes.data, q.shape[0], r.ctypes.data, r.shape[0], q.shape[1], k,
                              self_offset, idx.ctypes.data, dist.ctypes.data, nthreads)
    if rc != 0:
        raise ValueError(f"oracle_knn_exact failed: {rc}")
    return idx.astype(np.uint64), dist


def exact_knn_numpy(queries, refs, k, self_offset=0):
    """Pure-numpy twin of :func:`exact_knn` for small cases (cross-checks the C code)."""
    q = np.asarray(queries, dtype=np.float32).astype(np.float64)
    r = np.asarray(refs, dtype=np.float32).astype(np.float64)
    idx = np.empty((q.shape[0], k), dtype=np.uint64)
    dist = np.empty((q.shape[0], k), dtype=np.float32)
    for i in range(q.shape[0]):
        acc = np.zeros(r.shape[0])
        for t in range(q.shape[1]):
            df = q[i, t] - r[:, t]
            acc = acc + df * df
        d32 = acc.astype(np.float32)
        if self_offset >= 0:
            d32[i + self_offset] = np.inf
        order = np.lexsort((np.arange(r.shape[0]), d32))[:k]
        idx[i], dist[i] = order, d32[order]
    return idx, dist


# ----------------------------------------------------------------------------------------------
# A.7  edge weights    scarf/knn_utils.py:89-159 + umap-learn smooth_knn_dist / membership strengths
# ----------------------------------------------------------------------------------------------
SMOOTH_K_TOLERANCE = 1e-5
MIN_K_DIST_SCALE = 1e-3


def smooth_knn_dist(distances, k, n_iter=64, local_connectivity=1.0, bandwidth=1.0):
    """umap-learn (>=0.5) ``smooth_knn_dist`` restated with its float32 locals
    (psum, lo, mid, hi are numba float32; rho/result float32 arrays).  ``distances`` is one
    CHUNK (the per-chunk ``mean_distances`` is part of the reference's behaviour)."""
    f32 = np.float32
    dist = np.ascontiguousarray(distances, dtype=f32)
    n = dist.shape[0]
    target = f32(np.log2(k) * bandwidth)
    rho = np.zeros(n, dtype=f32)
    result = np.zeros(n, dtype=f32)
    mean_distances = f32(np.mean(dist))
    for i in range(n):
        lo, hi, mid = f32(0.0), f32(np.inf), f32(1.0)
        row = dist[i]
        nz = row[row > 0.0]
        if nz.shape[0] >= local_connectivity:
            index = int(np.floor(local_connectivity))
            interp = f32(local_connectivity - index)
            if index > 0:
                rho[i] = nz[index - 1]
                if interp > SMOOTH_K_TOLERANCE:
                    rho[i] += interp * (nz[index] - nz[index - 1])
            else:
                rho[i] = interp * nz[0]
        elif nz.shape[0] > 0:
            rho[i] = np.max(nz)
        for _ in range(n_iter):
            d = row[1:] - rho[i]
            with np.errstate(over="ignore", divide="ignore"):
                terms = np.where(d > 0, np.exp(-(d / mid), dtype=f32), f32(1.0)).astype(f32)
            psum = f32(0.0)
            for t in terms:  # sequential float32 accumulation like the numba loop
                psum = f32(psum + t)
            if np.fabs(psum - target) < SMOOTH_K_TOLERANCE:
                break
            if psum > target:
                hi = mid
                mid = f32((lo + hi) / f32(2.0))
            else:
                lo = mid
                if hi == np.inf:
                    mid = f32(mid * f32(2))
                else:
                    mid = f32((lo + hi) / f32(2.0))
        result[i] = mid
        if rho[i] > 0.0:
            mean_i = f32(np.mean(row))
            if result[i] < MIN_K_DIST_SCALE * mean_i:
                result[i] = f32(MIN_K_DIST_SCALE * mean_i)
        else:
            if result[i] < MIN_K_DIST_SCALE * mean_distances:
                result[i] = f32(MIN_K_DIST_SCALE * mean_distances)
    return result, rho


def smooth_knn_dist_vec(distances, k, n_iter=64, local_connectivity=1.0, bandwidth=1.0):
    """Row-vectorised twin of :func:`smooth_knn_dist` (same float32 steps, all rows at once) for
    sizes where the scalar loop is too slow.  Only ``local_connectivity`` with an integer part
    >= 1 or == 0 handled like the scalar code."""
    f32 = np.float32
    dist = np.ascontiguousarray(distances, dtype=f32)
    n, kk = dist.shape
    target = f32(np.log2(k) * bandwidth)
    mean_distances = f32(np.mean(dist))
    rho = np.zeros(n, dtype=f32)
    pos = dist > 0
    cnt = pos.sum(axis=1)
    index = int(np.floor(local_connectivity))
    interp = f32(local_connectivity - index)
    # compact positive entries to the front, keeping order
    order = np.argsort(~pos, axis=1, kind="stable")
    nzs = np.take_along_axis(dist, order, axis=1)
    enough = cnt >= local_connectivity
    if index > 0:
        r = nzs[:, index - 1].copy()
        if interp > SMOOTH_K_TOLERANCE:
            nxt = nzs[:, min(index, kk - 1)]
            r = (r + interp * (nxt - r)).astype(f32)
    else:
        r = (interp * nzs[:, 0]).astype(f32)
    rho[enough] = r[enough]
    some = (~enough) & (cnt > 0)
    rho[some] = np.where(pos, dist, -np.inf).max(axis=1)[some]
    lo = np.zeros(n, dtype=f32)
    hi = np.full(n, np.inf, dtype=f32)
    mid = np.ones(n, dtype=f32)
    active = np.ones(n, dtype=bool)
    d = (dist[:, 1:] - rho[:, None]).astype(f32)
    for _ in range(n_iter):
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            terms = np.where(d > 0, np.exp(-(d / mid[:, None]), dtype=f32), f32(1.0)).astype(f32)
        psum = np.zeros(n, dtype=f32)
        for j in range(terms.shape[1]):
            psum = (psum + terms[:, j]).astype(f32)
        done = np.fabs(psum - target) < SMOOTH_K_TOLERANCE
        active &= ~done
        if not active.any():
            break
        gt = psum > target
        up = active & gt
        dn = active & ~gt
        hi = np.where(up, mid, hi)
        lo = np.where(dn, mid, lo)
        with np.errstate(invalid="ignore", over="ignore"):
            half = ((lo + hi) / f32(2.0)).astype(f32)
        new_mid = np.where(up, half, np.where(dn, np.where(hi == np.inf, (mid * f32(2)).astype(f32), half), mid))
        mid = new_mid.astype(f32)
    result = mid.copy()
    mean_i = dist.mean(axis=1).astype(f32)
    floor_pos = (MIN_K_DIST_SCALE * mean_i).astype(f32)
    floor_zero = f32(MIN_K_DIST_SCALE * mean_distances)
    result = np.where(rho > 0, np.maximum(result, floor_pos), np.maximum(result, floor_zero)).astype(f32)
    return result, rho


def compute_membership_strengths(knn_indices, knn_dists, sigmas, rhos):
    """umap-learn ``compute_membership_strengths`` (float32 val).  ``i`` is the CHUNK-LOCAL row,
    ``knn_indices`` holds GLOBAL ids."""
    f32 = np.float32
    n, k = knn_indices.shape
    d = (np.asarray(knn_dists, dtype=f32) - rhos[:, None]).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        vals = np.exp(-(d / sigmas[:, None]), dtype=f32)
    vals = np.where((d <= 0) | (sigmas[:, None] == 0), f32(1.0), vals)
    vals = np.where(np.asarray(knn_indices).astype(np.int64) == np.arange(n)[:, None], f32(0.0), vals).astype(f32)
    rows = np.repeat(np.arange(n, dtype=np.int32), k)
    cols = np.asarray(knn_indices).reshape(-1).astype(np.int32)
    return rows, cols, vals.reshape(-1)


def smoothen_dists(indices, distances, lc=1.0, bw=1.5, chunk_size=1000, vectorised=True):
    """scarf/knn_utils.py:89-159 -> (edges uint64 (N*k,2), weights float64 (N*k,))."""
    n, k = indices.shape
    edges = np.empty((n * k, 2), dtype=np.uint64)
    weights = np.empty(n * k, dtype=np.float64)
    global_min = 1
    null_idx = []
    fn = smooth_knn_dist_vec if vectorised else smooth_knn_dist
    for s in range(0, n, chunk_size):
        ki = indices[s:s + chunk_size]
        kv = np.asarray(distances[s:s + chunk_size]).astype(np.float32, order="C")
        sig, rho = fn(kv, k=k, local_connectivity=lc, bandwidth=bw)
        rows, cols, vals = compute_membership_strengths(ki, kv, sig, rho)
        edges[s * k:(s + len(ki)) * k, 0] = rows.astype(np.int64) + s
        edges[s * k:(s + len(ki)) * k, 1] = cols
        weights[s * k:(s + len(ki)) * k] = vals
        nidx = vals == 0
        if nidx.sum() > 0:
            mv = vals[~nidx].min()
            if mv < global_min:
                global_min = mv
        null_idx.extend(nidx)
    null_idx = np.array(null_idx, dtype=bool)
    weights[null_idx] = global_min
    return edges, weights


# ----------------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------------
def make_graph(counts, cell_idx, hvg_mask, dims=11, k=11, lc=1.0, bw=1.5, batch_size=1000, pca="ipca",
               return_all=False, knn_threads=0, use_for_pca=None):
    """Restated ``make_graph(feat_key='hvgs')`` on CSR counts: normalise -> mu/sigma -> PCA -> exact
    kNN (self excluded) -> edge weights.  ``pca``: 'ipca' (reference's estimator) or 'exact'.
    ``use_for_pca``: bool mask over the selected cells (``pca_cell_key``, scarf/ann.py:215-228): the PCA is fitted
    on those rows of the z-scaled matrix only (exact route), every cell is projected."""
    feat_idx = np.where(hvg_mask)[0]
    x = normed_hvg(counts, cell_idx, feat_idx)
    mu, sigma = mu_sigma(x)
    k = min(k, x.shape[0] - 1)  # scarf/ann.py:85-86
    dims = clamp_dims(dims, x.shape[0], batch_size)
    z = (x - mu) / sigma
    if use_for_pca is not None:
        dims = clamp_dims(dims, int(np.sum(use_for_pca)), batch_size)
        loadings, _ = exact_pca_loadings(z[np.asarray(use_for_pca, dtype=bool)], dims)
    elif pca == "ipca":
        loadings = ipca_loadings(x, mu, sigma, dims, batch_size)
    else:
        loadings, _ = exact_pca_loadings(z, dims)
    y = z @ loadings
    idx, dist = exact_knn(y.astype(np.float32), y.astype(np.float32), k, self_offset=0, nthreads=knn_threads)
    edges, weights = smoothen_dists(idx, dist.astype(np.float64), lc, bw, batch_size)
    out = {"indices": idx, "distances": dist.astype(np.float64), "edges": edges, "weights": weights}
    if return_all:
        out.update({"x": x, "mu": mu, "sigma": sigma, "loadings": loadings, "embedding": y})
    return out


# ----------------------------------------------------------------------------------------------
# A.8  run_mapping                     scarf/mapping_utils.py:98-214, scarf/datastore/mapping_datastore.py:95-208
# ----------------------------------------------------------------------------------------------
def order_features(source_ids, target_ids, source_feat_ids):
    """``_order_features`` with exclude_missing=False, filter_null=False (the defaults, mapping_utils.py:98-145):
    for every source feature used by the graph, its column in the target matrix or -1 when the target lacks it."""
    pos = {v: i for i, v in enumerate(target_ids)}
    wanted = set(source_feat_ids)
    s_idx = np.array([i for i, v in enumerate(source_ids) if v in wanted], dtype=np.int64)
    t_re_idx = np.array([pos.get(source_ids[i], -1) for i in s_idx], dtype=np.int64)
    if np.all(t_re_idx == -1):
        raise ValueError("ERROR: None of the features from reference were found in the target data")
    return s_idx, t_re_idx


def aligned_target(target_counts, target_cell_idx, t_re_idx, sf=1000, log_transform=True, renormalize_subset=True,
                   n_counts=None):
    """``align_features`` (mapping_utils.py:186-213): the target normalised over the features it shares with the
    source (``target.normed(cells, sorted_t_idx, **source subset_params)``), columns in the SOURCE order, features
    the target lacks filled with 1.0."""
    present = t_re_idx != -1
    sorted_t_idx = np.array(sorted(t_re_idx[present]))
    normed = normed_hvg(target_counts, target_cell_idx, sorted_t_idx, sf, log_transform, renormalize_subset, n_counts)
    unsorter = np.argsort(np.argsort(t_re_idx[present]))
    a = np.ones((normed.shape[0], t_re_idx.size))
    a[:, np.where(present)[0]] = normed[:, unsorter]
    return a


def run_mapping(ref_embedding, loadings, mu, sigma, target_x, save_k=3, ref_mu=True, ref_sigma=True, nthreads=0):
    """The projection loop (mapping_datastore.py:177-208): z-scale the aligned target with the reference's mu /
    sigma (or the target's own when ``ref_mu`` / ``ref_sigma`` are False), embed with the reference loadings, query
    the reference index for ``save_k`` neighbours -- no self handling.  Exact search stands in for hnswlib."""
    if not ref_mu:
        mu = clean_array(target_x.mean(axis=0))
    if not ref_sigma:
        sigma = clean_array(target_x.std(axis=0), 1)
    y_q = ((target_x - mu) / sigma) @ loadings
    idx, dist = exact_knn(y_q.astype(np.float32), np.asarray(ref_embedding, dtype=np.float32), save_k, self_offset=-1,
                          nthreads=nthreads)
    return idx, dist.astype(np.float64), y_q


def mapping_score(indices, distances, n_ref, multiplier=1000.0, per_k=True):
    """``get_mapping_score`` with the defaults (mapping_datastore.py:255-285), one group; ``per_k=False`` reproduces
    the normalisation the reference's stored golden (cell_attributes.csv: mapping_scores) was generated with."""
    w = 1.0 / (np.log1p(distances) + 1.0)
    ms = np.zeros(n_ref)
    np.add.at(ms, np.asarray(indices, dtype=np.int64).ravel(), w.ravel())
    ms = multiplier * ms / (indices.shape[0] * (indices.shape[1] if per_k else 1))
    return np.log1p(ms)
