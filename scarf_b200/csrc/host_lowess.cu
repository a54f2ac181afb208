// Host-side robust LOWESS used by mark_hvgs' trend removal (scarf/feat_utils.py:22,38-40 calls statsmodels'
// lowess(endog, exog, frac, it, delta=0, return_sorted=False), a Cython routine; this is its native counterpart
// here).  O(it * n * k) on <= 200 points: microseconds, no reason to involve the GPU.
#include <math.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "common.cuh"

extern "C" int32_t scf_host_lowess(const double* endog, const double* exog, int64_t n, double frac, int32_t it,
                                   double* out) {
  SCF_ARG(endog && exog && out, "null pointer");
  SCF_ARG(n >= 2 && it >= 0, "bad sizes");
  const int64_t k = (int64_t)(frac * (double)n + 1e-10);
  SCF_ARG(k >= 2 && k <= n, "frac * n must be within [2, n]");
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return exog[a] < exog[b]; });
  std::vector<double> x(n), y(n), fit(n, 0.0), rw(n, 1.0), w(k), r(n), tmp(n), tri((size_t)n * k);
  std::vector<int64_t> lefts(n);
  for (int64_t i = 0; i < n; ++i) x[i] = exog[order[i]], y[i] = endog[order[i]];
  {  // windows and tricube weights do not change between the robustness passes
    int64_t left = 0, right = k;
    for (int64_t i = 0; i < n; ++i) {
      while (right < n && x[i] > 0.5 * (x[left] + x[right])) ++left, ++right;  // k nearest neighbours of x[i]
      lefts[i] = left;
      const double radius = std::max(x[i] - x[left], x[right - 1] - x[i]);
      for (int64_t j = 0; j < k; ++j) {
        const double d = fabs(x[left + j] - x[i]) / radius;
        double t = 1.0 - d * d * d;
        t = t * t * t;
        tri[(size_t)i * k + j] = isfinite(t) ? t : 0.0;
      }
    }
  }
  for (int32_t pass = 0; pass <= it; ++pass) {
    int64_t i = 0;
    while (i < n) {
      const int64_t left = lefts[i];
      double sw = 0.0;
      for (int64_t j = 0; j < k; ++j) {
        w[j] = tri[(size_t)i * k + j] * rw[left + j];
        sw += w[j];
      }
      if (!(sw > 0.0)) {
        fit[i] = y[i];
      } else {
        double xm = 0.0;
        for (int64_t j = 0; j < k; ++j) w[j] /= sw, xm += w[j] * x[left + j];
        double sq = 0.0;
        for (int64_t j = 0; j < k; ++j) sq += w[j] * (x[left + j] - xm) * (x[left + j] - xm);
        double f = 0.0;
        for (int64_t j = 0; j < k; ++j) {
          const double p = sq > 1e-12 ? w[j] * (1.0 + (x[i] - xm) * (x[left + j] - xm) / sq) : w[j];
          f += p * y[left + j];
        }
        fit[i] = f;
      }
      int64_t j = i + 1;  // delta = 0: tied x reuse the fit of the first point of the run
      while (j < n && x[j] == x[i]) fit[j++] = fit[i];
      i = j;
    }
    for (int64_t q = 0; q < n; ++q) r[q] = fabs(y[q] - fit[q]);
    tmp = r;
    std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
    double med = tmp[n / 2];
    if ((n & 1) == 0) med = 0.5 * (med + *std::max_element(tmp.begin(), tmp.begin() + n / 2));
    for (int64_t q = 0; q < n; ++q) {
      double v = med == 0.0 ? (r[q] > 0.0 ? 1.0 : 0.0) : r[q] / (6.0 * med);
      v = std::min(v, 1.0);
      rw[q] = (1.0 - v * v) * (1.0 - v * v);  // bisquare
    }
  }
  for (int64_t i = 0; i < n; ++i) out[order[i]] = fit[i];
  return 0;
}
