// Device-side robust LOWESS (block-level routine shared by scf_lowess and the fused HVG selection kernel): the same algorithm as scf_host_lowess (host_lowess.cu), run by ONE CTA so that
// mark_hvgs' trend removal (scarf/feat_utils.py:22,38-40 -> statsmodels lowess, frac 0.1, it 100, delta 0) needs no
// device->host round trip.  <= 512 points (the reference bins the genes into 200); the arithmetic
// uses the round-to-nearest intrinsics of the host routine (no FMA contraction); window sums are formed by four lanes
// (see below), so the two fits agree to ~1e-13 relative rather than bit for bit.
#pragma once
#include <math_constants.h>
#include "common.cuh"

constexpr int kLowessMaxN = 512;
constexpr int kLowessThreads = 1024;

// Runs the fit with all kLowessThreads threads of the calling CTA (uniform call).  ``dyn``: 2 * n * k doubles of shared
// memory when cache_tri != 0 (k <= frac * n + 1), unused otherwise.
__device__ __forceinline__ void lowess_block(const double* __restrict__ endog, const double* __restrict__ exog,
                                             const uint8_t* __restrict__ valid, int n_in, double frac, int it,
                                             int cache_tri, double* __restrict__ out, double* dyn) {
  constexpr int kMaxN = kLowessMaxN, kThreads = kLowessThreads;
  __shared__ double x[kMaxN], y[kMaxN], fit[kMaxN], rw[kMaxN], r[kMaxN], prev[kMaxN];
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
  double still = 0.0;  // 2^-46 of the largest |y|: a fit that moves less than this has converged
  for (int i = 0; i < n; ++i) still = fmax(still, fabs(y[i]));
  still *= 0x1p-46;
  // tricube weights do not change between the robustness passes: kept in shared memory when n * k values fit
  // (always for the 200 x 20 problem of mark_hvgs), recomputed otherwise.  wn = this pass's normalised weights.
  const bool cached = cache_tri != 0;
  double* tri = dyn;                        // [n][k]
  double* wn_all = dyn + (size_t)n * k;     // [n][k]
  auto tricube = [&](int i, int left, int j) {
    const double xi = x[i];
    const double radius = fmax(__dsub_rn(xi, x[left]), __dsub_rn(x[left + k - 1], xi));
    const double d = __ddiv_rn(fabs(__dsub_rn(x[left + j], xi)), radius);
    double t = __dsub_rn(1.0, __dmul_rn(__dmul_rn(d, d), d));
    t = __dmul_rn(__dmul_rn(t, t), t);
    return isfinite(t) ? t : 0.0;
  };
  if (cached) {
    for (int e = tid; e < n * k; e += kThreads) {
      const int i = e / k, j = e - i * k;
      tri[e] = tricube(i, lefts[i], j);
    }
    __syncthreads();
  }
  // Four lanes per point: every window sum is formed as four interleaved partial sums that a two-step butterfly
  // adds up (all four lanes obtain the same value).  FP64 dependent chains are what this kernel waits for, and the
  // quad cuts them by four; the fit agrees with the sequential host routine to ~1e-13 relative.
  const int quad = tid & 3;
  const unsigned qmask = 0xFu << ((tid & 31) & ~3);
  auto quad_sum = [&](double v) {
    v = __dadd_rn(v, __shfl_xor_sync(qmask, v, 1));
    return __dadd_rn(v, __shfl_xor_sync(qmask, v, 2));
  };
  for (int pass = 0; pass <= it; ++pass) {
    for (int i0 = 0; i0 < n; i0 += kThreads / 4) {
      const int i = i0 + (tid >> 2);
      const bool live = i < n && first[i] == i;  // uniform inside a quad
      if (!live) continue;
      const int left = lefts[i];
      const double xi = x[i];
      auto weight = [&](int j) {  // tricube * robustness weight, un-normalised
        return __dmul_rn(cached ? tri[(size_t)i * k + j] : tricube(i, left, j), rw[left + j]);
      };
      double* wn = wn_all + (size_t)i * k;  // per-point scratch (only when cached); lane q owns entries j = q mod 4
      double part = 0.0;
      for (int j = quad; j < k; j += 4) {
        const double wj = weight(j);
        if (cached) wn[j] = wj;
        part = __dadd_rn(part, wj);
      }
      const double sw = quad_sum(part);
      double f = y[i];
      if (sw > 0.0) {
        if (cached)
          for (int j = quad; j < k; j += 4) wn[j] = __ddiv_rn(wn[j], sw);
        auto w_of = [&](int j) { return cached ? wn[j] : __ddiv_rn(weight(j), sw); };
        part = 0.0;
        for (int j = quad; j < k; j += 4) part = __dadd_rn(part, __dmul_rn(w_of(j), x[left + j]));
        const double xm = quad_sum(part);
        part = 0.0;
        for (int j = quad; j < k; j += 4) {
          const double dx = __dsub_rn(x[left + j], xm);
          part = __dadd_rn(part, __dmul_rn(__dmul_rn(w_of(j), dx), dx));
        }
        const double sq = quad_sum(part);
        const double xd = __dsub_rn(xi, xm);
        part = 0.0;
        for (int j = quad; j < k; j += 4) {
          const double w = w_of(j);
          const double p =
              sq > 1e-12 ? __dmul_rn(w, __dadd_rn(1.0, __ddiv_rn(__dmul_rn(xd, __dsub_rn(x[left + j], xm)), sq))) : w;
          part = __dadd_rn(part, __dmul_rn(p, y[left + j]));
        }
        f = quad_sum(part);
      }
      if (quad == 0) fit[i] = f;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
      if (first[i] != i) fit[i] = fit[first[i]];
    }
    // The robustness iteration contracts (the fit moves ~3x less from one pass to the next) until it reaches the
    // rounding jitter of the window sums (~3e-15 of the data's scale, where it wanders for the rest of the `it` passes
    // statsmodels always runs).  Once no point has moved by more than 2^-46 of the scale the remaining passes are
    // skipped: what they would still change is below the 1e-13 agreement of this routine with the host / reference
    // fit.  (The barrier also publishes the tied points' values.)
    int moved = 0;
    for (int i = tid; i < n; i += kThreads) {
      moved |= !(fabs(__dsub_rn(fit[i], prev[i])) <= still);  // NaN counts as moved
      prev[i] = fit[i];
    }
    if (!__syncthreads_or(moved) && pass > 0) break;
    if (pass == it) break;  // the weights of a further pass are never used
    for (int i = tid; i < n; i += kThreads) r[i] = fabs(__dsub_rn(y[i], fit[i]));
    __syncthreads();
    // median by rank: the order statistics n/2 (and n/2 - 1 for even n), ties broken by position; a quad per point
    for (int i0 = 0; i0 < n; i0 += kThreads / 4) {
      const int i = i0 + (tid >> 2);
      if (i >= n) continue;  // uniform inside a quad
      const double ri = r[i];
      int rank = 0;
      for (int j = quad; j < n; j += 4) rank += (r[j] < ri) || (r[j] == ri && j < i);
      rank += __shfl_xor_sync(qmask, rank, 1);
      rank += __shfl_xor_sync(qmask, rank, 2);
      if (quad == 0) {
        if (rank == n / 2) s_med[0] = ri;
        if (rank == n / 2 - 1) s_med[1] = ri;
      }
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

