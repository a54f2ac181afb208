"""Host-side part of ``mark_hvgs``: trend removal and HVG choice on the per-gene statistics.

O(G) work on ~30k numbers: it stays on the host like the reference's
(scarf/metadata.py:586-617, scarf/feat_utils.py:11-45, scarf/assay.py:1014-1063); the per-gene
statistics themselves come from the CSR kernels.
"""
from __future__ import annotations

import re

import numpy as np

DEFAULT_BLACKLIST = "^MT-|^RPS|^RPL|^MRPS|^MRPL|^CCN|^HLA-|^H2-|^HIST"  # scarf/datastore/datastore.py:235


def _lowess(y, x, frac, it):
    """statsmodels' ``lowess(endog, exog, frac, it, delta=0, return_sorted=False)`` (scarf/feat_utils.py:38-40)
    through the library's native host routine (statsmodels' own is Cython); :func:`_lowess_numpy` is the
    all-numpy statement of the same algorithm the tests compare it with."""
    from . import lib

    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(y)
    lib.call("scf_host_lowess", y.ctypes.data, x.ctypes.data, int(y.size), float(frac), int(it), out.ctypes.data,
             launches=0)
    return out


def _lowess_numpy(y, x, frac, it):
    """Cleveland's robust LOWESS, all windows evaluated at once."""
    order = np.argsort(x, kind="stable")
    x, y = x[order], y[order]
    n = x.size
    k = int(frac * n + 1e-10)
    if not 2 <= k <= n:
        raise ValueError("lowess: frac * n must be within [2, n]")
    left = np.zeros(n, dtype=np.int64)
    lo = 0
    for i in range(n):  # the k-wide window slides right while x[i] is nearer to its far end
        while lo + k < n and x[i] > 0.5 * (x[lo] + x[lo + k]):
            lo += 1
        left[i] = lo
    first = np.arange(n)  # tied x reuse the fit of the first point of the run (delta = 0)
    for i in range(1, n):
        if x[i] == x[i - 1]:
            first[i] = first[i - 1]
    win = left[first][:, None] + np.arange(k)[None, :]
    xs, ys = x[win], y[win]
    xi = x[first][:, None]
    radius = np.maximum(xi[:, 0] - xs[:, 0], xs[:, -1] - xi[:, 0])[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.abs(xs - xi) / radius
    tri = (1.0 - u ** 3) ** 3
    tri[~np.isfinite(tri)] = 0.0
    rw = np.ones(n)
    fit = np.zeros(n)
    prev_rw = None
    for _ in range(it + 1):
        if prev_rw is not None and np.array_equal(rw, prev_rw):
            break  # fixed point reached bit for bit: every later iteration would reproduce `fit`
        prev_rw = rw
        w = tri * rw[win]
        sw = w.sum(axis=1, keepdims=True)
        ok = sw[:, 0] > 0
        wn = np.divide(w, sw, out=np.zeros_like(w), where=sw > 0)
        xm = (wn * xs).sum(axis=1, keepdims=True)
        dx = xs - xm
        sq = (wn * dx * dx).sum(axis=1, keepdims=True)
        slope_ok = sq > 1e-12
        p = np.where(slope_ok, wn * (1.0 + (xi - xm) * dx / np.where(slope_ok, sq, 1.0)), wn)
        fit = np.where(ok, (p * ys).sum(axis=1), y[first])
        r = np.abs(y - fit)
        med = np.median(r)
        r = (r > 0).astype(np.float64) if med == 0 else r / (6.0 * med)
        r = np.minimum(r, 1.0)
        rw = (1.0 - r * r) ** 2
    out = np.empty(n)
    out[order] = fit
    return out


def fit_lowess(a, b, n_bins=200, lowess_frac=0.1):
    """scarf/feat_utils.py:11-45: bin log(a), take the min-log(b) gene per bin, LOWESS through those,
    return exp(log b - fit(bin)) per gene."""
    la, lb = np.log(a), np.log(b)
    edges = np.histogram(la, bins=n_bins)[1]
    edges[-1] += 0.1
    which = np.searchsorted(edges, la, side="right") - 1  # edges[i] <= la < edges[i+1]
    which[(which < 0) | (which >= n_bins)] = -1
    # min-log(b) gene of every non-empty bin (first one on ties, like argmin over the bin's members)
    order = np.lexsort((lb, which))
    order = order[which[order] >= 0]
    w_sorted = which[order]
    firsts = order[np.r_[True, w_sorted[1:] != w_sorted[:-1]]] if order.size else order
    bins = which[firsts]
    fit = _lowess(lb[firsts], la[firsts], lowess_frac, 100)
    fit_of_bin = np.full(n_bins, np.nan)
    fit_of_bin[bins] = fit
    out = np.zeros(la.size)
    sel = which >= 0
    out[sel] = np.exp(lb[sel] - fit_of_bin[which[sel]])
    return out


def remove_trend(avg, sigmas, n_bins=200, lowess_frac=0.1, fill_value=0.0):
    """scarf/metadata.py:586-617."""
    avg, sigmas = np.asarray(avg, dtype=np.float64), np.asarray(sigmas, dtype=np.float64)
    ret = np.full(avg.size, float(fill_value))
    pos = avg > 0
    ret[pos] = fit_lowess(avg[pos], sigmas[pos], n_bins, lowess_frac)
    return ret


def _sift(v, lo, hi, keep_bounds):
    """MetaData.sift (scarf/metadata.py:483-505); works on numpy arrays and torch tensors."""
    return (v >= lo) & (v <= hi) if keep_bounds else (v > lo) & (v < hi)


def choose_hvgs(normed_n, nz_mean, c_var, feat_I, gene_names=None, top_n=500, min_cells=0, max_cells=np.inf,
                min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf, blacklist=DEFAULT_BLACKLIST,
                keep_bounds=False):
    """scarf/assay.py:1014-1063 + MetaData.multi_sift (scarf/metadata.py:483-533): bounds strict unless
    ``keep_bounds``; with ``min_var == -inf`` (the default) the threshold is the (top_n + 1)-th largest corrected
    variance of the eligible genes, otherwise ``top_n`` is ignored and ``2**min_var`` is the threshold;
    ``c_var < 2**max_var`` applies in both modes.  Mean / variance bounds are given in log2 and only exponentiated when
    finite.  All vectors cover every gene (NaN outside ``feat_I``).  Pinned on the executed reference method
    (tests/golden/ref_functions.npz)."""
    g = normed_n.size
    keep = blacklist_keep_mask(gene_names, g, blacklist)
    min_mean = 2.0 ** min_mean if min_mean != -np.inf else min_mean
    max_mean = 2.0 ** max_mean if max_mean != np.inf else max_mean
    max_var = 2.0 ** max_var if max_var != np.inf else max_var
    with np.errstate(invalid="ignore"):
        idx = _sift(normed_n, min_cells, max_cells, keep_bounds) & _sift(nz_mean, min_mean, max_mean, keep_bounds)
        idx &= feat_I & keep
        if min_var == -np.inf:
            if top_n < 1:
                raise ValueError("ERROR: Please provide a value greater than 0 for `top_n` parameter")
            n_valid = int(idx.sum())
            if top_n > n_valid:
                top_n = n_valid - 1
            min_var = np.sort(c_var[idx])[::-1][top_n]
        else:
            min_var = 2.0 ** min_var
        hv = idx & _sift(c_var, min_var, max_var, keep_bounds)
    return hv


# =============================================================================================
# The same two steps on device tensors: the per-gene vectors stay on the GPU (they come from the CSR kernels), only
# the <= n_bins binned points cross to the host for the LOWESS fit.  Used by graph.mark_hvgs_csr.
# =============================================================================================
def blacklist_keep_mask(gene_names, n_genes, blacklist=DEFAULT_BLACKLIST):
    """bool numpy mask of the genes that survive the blacklist regex (static per dataset).  As the reference does it
    (MetaData.grep, scarf/metadata.py:569-584): names AND pattern are upper-cased, the pattern must match at the
    start of the name (``re.match``)."""
    if blacklist and gene_names is not None:
        pat = re.compile(blacklist.upper())
        return np.fromiter((pat.match(str(x).upper()) is None for x in gene_names), dtype=bool, count=n_genes)
    return np.ones(n_genes, dtype=bool)


def remove_trend_device(avg, sigmas, n_bins=200, lowess_frac=0.1, select=None):
    """:func:`remove_trend` on float64 device vectors (any device); returns a device vector of the same length.
    ``select`` (bool tensor, optional) restricts the fit to those entries (the feature ``I`` column); the others get
    0.  Written without boolean indexing: on a GPU nothing synchronises (the <= n_bins binned points are fitted by
    the one-CTA device LOWESS, ``scf_lowess``); CPU tensors use the native host routine."""
    import torch

    n = avg.numel()
    dev = avg.device
    pos = avg > 0 if select is None else (avg > 0) & select
    nan = torch.full_like(avg, float("nan"))
    la = torch.where(pos, torch.log(torch.where(pos, avg, torch.ones_like(avg))), nan)
    lb = torch.where(pos, torch.log(torch.where(pos, sigmas, torch.ones_like(sigmas))), nan)
    inf = torch.full_like(avg, float("inf"))
    first = torch.where(pos, la, inf).min()
    last = torch.where(pos, la, -inf).max()
    same = first == last  # np.histogram widens a zero-width range by +-0.5
    first, last = torch.where(same, first - 0.5, first), torch.where(same, last + 0.5, last)
    # np.histogram's edges for `bins=n_bins`: linspace(first, last, n_bins + 1) = arange * step + first with the last
    # edge set to `last` exactly; the reference then widens the last edge by 0.1 (feat_utils.py:25-27)
    step = (last - first) / n_bins
    edges = torch.arange(n_bins + 1, dtype=torch.float64, device=dev) * step + first
    edges[-1] = last + 0.1
    which = torch.bucketize(torch.where(pos, la, inf), edges, right=True) - 1  # edges[i] <= la < edges[i+1]
    valid = pos & (which >= 0) & (which < n_bins)
    which = torch.where(valid, which, torch.full_like(which, n_bins))  # invalid -> a bin past the end
    # min-log(b) gene of every non-empty bin, the first one on ties (argmin over the bin's members)
    binmin = torch.full((n_bins + 1,), float("inf"), dtype=torch.float64, device=dev)
    binmin.scatter_reduce_(0, which, torch.where(valid, lb, inf), "amin", include_self=True)
    ids = torch.arange(n, device=dev)
    cand = torch.where(valid & (lb == binmin[which]), ids, torch.full_like(ids, n))
    firsts = torch.full((n_bins + 1,), n, dtype=ids.dtype, device=dev)
    firsts.scatter_reduce_(0, which, cand, "amin", include_self=True)
    firsts = firsts[:n_bins]
    safe = firsts.clamp(max=max(n - 1, 0))
    if avg.is_cuda and n_bins <= 512:  # device LOWESS (one CTA): the whole trend removal stays asynchronous
        from . import ops

        occupied = firsts < n
        fit_t = ops.lowess(lb[safe].contiguous(), la[safe].contiguous(), occupied.to(torch.uint8), lowess_frac, 100)
        fit_t = torch.cat([fit_t, torch.full((1,), float("nan"), dtype=torch.float64, device=dev)])
        val = torch.exp(lb - fit_t[which])
        return torch.where(valid, val, torch.zeros_like(val))
    pts = torch.stack([(firsts < n).to(torch.float64), la[safe], lb[safe]]).cpu().numpy()  # the one synchronisation
    bins = np.where(pts[0] > 0)[0]
    fit_of_bin = np.full(n_bins + 1, np.nan)
    if bins.size:
        fit_of_bin[bins] = _lowess(pts[2][bins], pts[1][bins], lowess_frac, 100)
    fit_t = torch.from_numpy(fit_of_bin).to(dev)
    val = torch.exp(lb - fit_t[which])
    return torch.where(valid, val, torch.zeros_like(val))


def choose_hvgs_device(normed_n, nz_mean, c_var, eligible, top_n=500, min_cells=0, max_cells=np.inf,
                       min_mean=-np.inf, max_mean=np.inf, min_var=-np.inf, max_var=np.inf, keep_bounds=False):
    """:func:`choose_hvgs` on device vectors; ``eligible`` = feat_I & blacklist-keep (bool tensor).  No
    synchronisation: the (top_n+1)-th largest corrected variance is picked with a device-side index."""
    import torch

    min_mean = 2.0 ** min_mean if min_mean != -np.inf else min_mean
    max_mean = 2.0 ** max_mean if max_mean != np.inf else max_mean
    max_var = 2.0 ** max_var if max_var != np.inf else max_var
    idx = _sift(normed_n, min_cells, max_cells, keep_bounds) & _sift(nz_mean, min_mean, max_mean, keep_bounds) & eligible
    idx = idx & ~torch.isnan(c_var)
    if min_var == -np.inf:
        if top_n < 1:
            raise ValueError("ERROR: Please provide a value greater than 0 for `top_n` parameter")
        ninf = torch.full_like(c_var, float("-inf"))
        cv = torch.sort(torch.where(idx, c_var, ninf), descending=True).values
        n_valid = idx.sum()
        kk = torch.minimum(torch.full_like(n_valid, int(top_n)), n_valid - 1).clamp(min=0)  # assay.py:1035-1040
        thr = cv[kk]
        return idx & ((c_var >= thr) & (c_var <= max_var) if keep_bounds else (c_var > thr) & (c_var < max_var))
    return idx & _sift(c_var, 2.0 ** min_var, max_var, keep_bounds)
