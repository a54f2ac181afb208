"""LOWESS restatement (test infrastructure; see oracle/__init__.py).

The reference calls ``statsmodels.nonparametric.smoothers_lowess.lowess(endog,
exog, frac=0.1, it=100, return_sorted=False)`` (scarf/feat_utils.py:22,38-40;
statsmodels is unpinned in requirements.txt:20 and absent from this image).
This is a plain-loop restatement of the published algorithm (Cleveland 1979,
as implemented by statsmodels' ``_lowess.pyx`` with ``delta=0``):

  k = int(frac*n + 1e-10) nearest neighbours; tricube kernel weights times the
  robustness weights, normalised to sum 1; weighted local *linear* fit
  evaluated at x_i; then ``it`` robustifying passes with bisquare weights on
  residuals scaled by 6*median|r|.

Pinned transitively: with this function in the chain the oracle reproduces the
reference's PBMC ``knn_indices.npy`` (tests/test_oracle_golden.py).
"""
import numpy as np


def lowess(endog, exog, frac=2.0 / 3.0, it=3):
    y = np.asarray(endog, dtype=np.float64)
    x = np.asarray(exog, dtype=np.float64)
    order = np.argsort(x, kind="mergesort")
    x, y = x[order], y[order]
    n = x.shape[0]
    k = int(frac * n + 1e-10)
    if not 2 <= k <= n:
        raise ValueError("Invalid frac: need 2 <= int(frac*n) <= n")
    y_fit = np.zeros(n)
    resid_w = np.ones(n)
    for _ in range(it + 1):
        left, right = 0, k
        y_fit = np.zeros(n)
        i = 0
        while i < n:
            # slide the k-wide window so that it holds the k nearest neighbours of x[i]
            while right < n and x[i] > (x[left] + x[right]) / 2.0:
                left += 1
                right += 1
            xs = x[left:right]
            radius = max(x[i] - x[left], x[right - 1] - x[i])
            with np.errstate(divide="ignore", invalid="ignore"):
                dist = np.abs(xs - x[i]) / radius
            w = (1.0 - dist ** 3) ** 3
            w = np.where(np.isfinite(w), w, 0.0) * resid_w[left:right]
            sw = w.sum()
            if sw <= 0.0:
                y_fit[i] = y[i]
            else:
                w = w / sw
                xm = np.dot(w, xs)
                sq = np.dot(w, (xs - xm) ** 2)
                p = w.copy()
                if sq > 1e-12:
                    p = w * (1.0 + (x[i] - xm) * (xs - xm) / sq)
                y_fit[i] = np.dot(p, y[left:right])
            # delta = 0: tied x values reuse the fit of the first of the run
            j = i + 1
            while j < n and x[j] == x[i]:
                y_fit[j] = y_fit[i]
                j += 1
            i = j
        r = np.abs(y - y_fit)
        med = np.median(r)
        if med == 0:
            r = (r > 0).astype(np.float64)
        else:
            r = r / (6.0 * med)
        r[r >= 1.0] = 1.0
        resid_w = (1.0 - r ** 2) ** 2
    out = np.empty(n)
    out[order] = y_fit
    return out
