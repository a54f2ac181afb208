// Device-side robust LOWESS: the same algorithm as scf_host_lowess (host_lowess.cu), run by ONE CTA so that
// mark_hvgs' trend removal (scarf/feat_utils.py:22,38-40 -> statsmodels lowess, frac 0.1, it 100, delta 0) needs no
// device->host round trip.  <= 512 points (the reference bins the genes into 200); every arithmetic step uses the
// round-to-nearest intrinsics in the host routine's order (no FMA contraction), so both give the same fit.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kMaxN = 512;
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) lowess_kernel(const double* __restrict__ endog,
                                                          const double* __restrict__ exog,
                                                          const uint8_t* __restrict__ valid, int n_in, double frac, int it,
                                                          double* __restrict__ out) {
  __shared__ double x[kMaxN], y[kMaxN], fit[kMaxN], rw[kMaxN], r[kMaxN];
  __shared__ int src[kMaxN], lefts[kMaxN], first[kMaxN];
  __shared__ int s_n;
  __shared__ double s_med[2];
  const int tid = threadIdx.x;
  // ---- compact the usable points (input order), then stable rank sort by x ----
  if (tid == 0) {
    int m = 0;
    for (int i = 0; i < n_in; ++i)
      if (!valid || valid[i]) src[m++] = i;
    s_n = m;
  }
  for (int i = tid; i < n_in; i += kThreads) out[i] = CUDART_NAN;
  __syncthreads();
  const int n = s_n;
  const int k = (int)(frac * (double)n + 1e-10);
  if (n < 2 || k < 2 || k > n) return;  // the host routine reports this as an argument error: the fit stays NaN
  for (int i = tid; i < n; i += kThreads) r[i] = exog[src[i]];  // r = unsorted x (scratch)
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    const double xi = r[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (r[j] < xi) || (r[j] == xi && j < i);
    lefts[rank] = src[i];  // lefts = sorted source ids (scratch)
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    src[i] = lefts[i];
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    x[i] = exog[src[i]];
    y[i] = endog[src[i]];
    rw[i] = 1.0;
    fit[i] = 0.0;
  }
  __syncthreads();
  if (tid == 0) {  // sliding k-nearest window and runs of tied x: sequential like the host code, n steps
    int left = 0, right = k;
    for (int i = 0; i < n; ++i) {
      while (right < n && x[i] > __dmul_rn(0.5, __dadd_rn(x[left], x[right]))) ++left, ++right;
      lefts[i] = left;
      first[i] = (i > 0 && x[i] == x[i - 1]) ? first[i - 1] : i;
    }
  }
  __syncthreads();
  for (int pass = 0; pass <= it; ++pass) {
    for (int i = tid; i < n; i += kThreads) {
      if (first[i] != i) continue;
      const int left = lefts[i];
      const double xi = x[i];
      const double radius = fmax(__dsub_rn(xi, x[left]), __dsub_rn(x[left + k - 1], xi));
      auto weight = [&](int j) {  // tricube(|x_j - x_i| / radius) * robustness weight, un-normalised
        const double d = __ddiv_rn(fabs(__dsub_rn(x[left + j], xi)), radius);
        double t = __dsub_rn(1.0, __dmul_rn(__dmul_rn(d, d), d));
        t = __dmul_rn(__dmul_rn(t, t), t);
        if (!isfinite(t)) t = 0.0;
        return __dmul_rn(t, rw[left + j]);
      };
      double sw = 0.0;
      for (int j = 0; j < k; ++j) sw = __dadd_rn(sw, weight(j));
      double f;
      if (!(sw > 0.0)) {
        f = y[i];
      } else {
        double xm = 0.0;
        for (int j = 0; j < k; ++j) xm = __dadd_rn(xm, __dmul_rn(__ddiv_rn(weight(j), sw), x[left + j]));
        double sq = 0.0;
        for (int j = 0; j < k; ++j) {
          const double dx = __dsub_rn(x[left + j], xm);
          sq = __dadd_rn(sq, __dmul_rn(__dmul_rn(__ddiv_rn(weight(j), sw), dx), dx));
        }
        f = 0.0;
        for (int j = 0; j < k; ++j) {
          const double w = __ddiv_rn(weight(j), sw);
          const double p =
              sq > 1e-12 ? __dmul_rn(w, __dadd_rn(1.0, __ddiv_rn(__dmul_rn(__dsub_rn(xi, xm), __dsub_rn(x[left + j], xm)), sq)))
                         : w;
          f = __dadd_rn(f, __dmul_rn(p, y[left + j]));
        }
      }
      fit[i] = f;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
      if (first[i] != i) fit[i] = fit[first[i]];
    }
    __syncthreads();
    if (pass == it) break;  // the weights of a further pass are never used
    for (int i = tid; i < n; i += kThreads) r[i] = fabs(__dsub_rn(y[i], fit[i]));
    __syncthreads();
    // median by rank: the order statistics n/2 (and n/2 - 1 for even n), ties broken by position
    for (int i = tid; i < n; i += kThreads) {
      const double ri = r[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (r[j] < ri) || (r[j] == ri && j < i);
      if (rank == n / 2) s_med[0] = ri;
      if (rank == n / 2 - 1) s_med[1] = ri;
    }
    __syncthreads();
    const double med = (n & 1) ? s_med[0] : __dmul_rn(0.5, __dadd_rn(s_med[0], s_med[1]));
    for (int i = tid; i < n; i += kThreads) {
      double v = med == 0.0 ? (r[i] > 0.0 ? 1.0 : 0.0) : __ddiv_rn(r[i], __dmul_rn(6.0, med));
      v = fmin(v, 1.0);
      const double u = __dsub_rn(1.0, __dmul_rn(v, v));
      rw[i] = __dmul_rn(u, u);  // bisquare
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += kThreads) out[src[i]] = fit[i];
}

}  // namespace

extern "C" int32_t scf_lowess(const double* endog, const double* exog, const uint8_t* valid, int32_t n, double frac,
                              int32_t it, double* out, void* stream) {
  SCF_ARG(endog && exog && out, "null pointer");
  SCF_ARG(n >= 1 && n <= kMaxN && it >= 0, "n must be within [1, 512]");
  lowess_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(endog, exog, valid, n, frac, it, out);
  return scf_check_launch("scf_lowess");
}
