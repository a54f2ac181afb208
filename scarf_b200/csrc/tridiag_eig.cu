// All eigenpairs of a small symmetric matrix (n <= 160) in ONE CTA, the fast path of the Rayleigh-Ritz steps of
// scf_eig_topk (eig_topk.cu): Householder tridiagonalisation in shared memory, eigenvalues by multisection on the Sturm
// count, eigenvectors by inverse iteration (one thread per vector, Gaussian elimination with partial pivoting as in
// LAPACK dstein / dlagtf), back-transformation with the stored reflectors.  ~0.15 ms at n = 128, where the one-sided
// Jacobi kernel (jacobi_eig.cu) needs 2.3 ms: a Jacobi sweep is n - 1 dependent steps of ~2,500 cycles on one SM and
// eight sweeps are needed.  Inverse iteration without re-orthogonalisation is accurate while the eigenvalues are
// separated by more than ~1e-7 |T| (cross-contamination eps |T| / gap); the kernel measures the orthogonality of what it
// produced and reports it in `ok`: the caller then runs the Jacobi kernel, which returns at once when *ok == 1.
// Deterministic: fixed reduction orders, seeded start vectors.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

__device__ long long g_te_clk[8];  // developer timing (SCF_EIG_DEBUG=1): cycles of the phases of the last launch
constexpr int TE_MAXN = 160;
constexpr int TE_THREADS = 1024;

__device__ __forceinline__ double group8_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// numbers of eigenvalues of the tridiagonal (d, e2 = e^2) below x[0..3): sign changes of the characteristic polynomial
// recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (no division: one FMA on the dependent chain per row;
// rescaled every eight rows against overflow).  Three independent chains per thread: the instruction latency of one
// chain (~10 cycles per instruction with four warps on the SM) hides behind the other two.
__device__ __forceinline__ void sturm_count3(const double* __restrict__ d, const double* __restrict__ e2, int n,
                                             const double (&x)[3], int (&cnt)[3]) {
  double p0[3], p1[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p0[c] = 1.0, p1[c] = d[0] - x[c];
    cnt[c] = p1[c] < 0.0;
  }
  int i = 1;
  while (i < n) {
    const int stop = min(n, i + 8);
    for (; i < stop; ++i) {
      const double di = d[i], ei = e2[i - 1];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double p2 = fma(di - x[c], p1[c], -ei * p0[c]);
        if (p2 == 0.0) p2 = p1[c] < 0.0 ? 1e-300 : -1e-300;  // a zero counts as a sign change (eigenvalue <= x)
        cnt[c] += (p2 < 0.0) != (p1[c] < 0.0);
        p0[c] = p1[c], p1[c] = p2;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double ap = fabs(p1[c]);
      if (ap > 1e100 || ap < 1e-100) {
        const double sc = ap > 1.0 ? 1e-100 : 1e100;
        p0[c] *= sc, p1[c] *= sc;
      }
    }
  }
}

// reflectors k with max(0, (k - 6) / 8) == T, applied to the rows t >= T of the column held in xr
template <int T, int RMAX>
__device__ __forceinline__ void bt_range(double (&xr)[RMAX], const double* __restrict__ A, int ld, int n, int sub,
                                         unsigned gmask) {
  const int k_hi = min(n - 3, T == RMAX - 1 ? n - 3 : 8 * T + 13), k_lo = T == 0 ? 0 : 8 * T + 6;
  for (int k = k_hi; k >= k_lo; --k) {
    const double* vcol = A + (size_t)(k + 1) * ld + k;  // v_i = vcol[i * ld], row k + 1 + i
    double dot = 0.0;
#pragma unroll
    for (int t = T; t < RMAX; ++t) {
      const int r = sub + 8 * t;
      if (r > k && r < n) dot = fma(vcol[(size_t)(r - k - 1) * ld], xr[t], dot);
    }
    dot = 2.0 * group8_sum(dot, gmask);
#pragma unroll
    for (int t = T; t < RMAX; ++t) {
      const int r = sub + 8 * t;
      if (r > k && r < n) xr[t] = fma(-dot, vcol[(size_t)(r - k - 1) * ld], xr[t]);
    }
  }
}
template <int T, int RMAX>
__device__ __forceinline__ void bt_sweep(double (&xr)[RMAX], const double* __restrict__ A, int ld, int n, int sub,
                                         unsigned gmask) {
  bt_range<T, RMAX>(xr, A, ld, n, sub, gmask);
  if constexpr (T > 0) bt_sweep<T - 1, RMAX>(xr, A, ld, n, sub, gmask);
}

__global__ void __launch_bounds__(TE_THREADS, 1) tridiag_eig_kernel(const double* __restrict__ a, int n, int64_t lda,
                                                                   double* __restrict__ evals,
                                                                   double* __restrict__ evecs, int64_t ldv,
                                                                   int descending, double* __restrict__ work,
                                                                   int* __restrict__ ok) {
  extern __shared__ __align__(16) double A[];  // [n][n + 1]
  __shared__ double s_d[TE_MAXN], s_e[TE_MAXN], s_e2[TE_MAXN], s_v[TE_MAXN], s_p[TE_MAXN], s_q[TE_MAXN];
  __shared__ double s_lam[TE_MAXN];
  __shared__ double s_red[32];
  __shared__ double s_scal[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ld = n + 1;
  const int grp = tid >> 3, sub = tid & 7;
  const unsigned gmask = 0xFFu << (lane & ~7);
  for (int e = tid; e < n * n; e += TE_THREADS) {
    const int r = e / n, c = e - r * n;
    A[r * ld + c] = 0.5 * (a[(int64_t)r * lda + c] + a[(int64_t)c * lda + r]);
  }
  __syncthreads();
  long long t_prev = clock64();
  auto stamp = [&](int slot) {
    if (tid == 0) {
      const long long t = clock64();
      g_te_clk[slot] = t - t_prev;
      t_prev = t;
    }
  };

  // ---- Householder tridiagonalisation: A <- H_k A H_k, H_k = I - 2 v v^T on rows / columns k+1 .. n-1 ----
  for (int k = 0; k + 2 < n; ++k) {
    const int m = n - k - 1;
    double* col = A + (size_t)(k + 1) * ld + k;  // x_i = col[i * ld]
    if (warp == 0) {
      double s = 0.0;
      for (int i = lane; i < m; i += 32) s = fma(col[i * ld], col[i * ld], s);
      s = warp_sum(s);
      const double x0 = col[0];
      const double alpha = s > 0.0 ? -copysign(sqrt(s), x0) : 0.0;
      const double v0 = x0 - alpha;
      const double vn2 = s - x0 * x0 + v0 * v0;  // |v|^2
      const double inv = vn2 > 0.0 ? rsqrt(vn2) : 0.0;
      for (int i = lane; i < m; i += 32) s_v[i] = (i == 0 ? v0 : col[i * ld]) * inv;
      if (lane == 0) s_scal[0] = alpha;
    }
    __syncthreads();
    // p = S v, S = A[k+1.., k+1..]; eight threads per row
    for (int i = grp; i < m; i += TE_THREADS / 8) {
      const double* row = A + (size_t)(k + 1 + i) * ld + (k + 1);
      double acc = 0.0;
      for (int j = sub; j < m; j += 8) acc = fma(row[j], s_v[j], acc);
      acc = group8_sum(acc, gmask);
      if (sub == 0) s_p[i] = acc;
    }
    __syncthreads();
    // K = v^T p (every warp computes it for itself), q = 2 (p - K v)
    {
      double kk = 0.0;
      for (int i = lane; i < m; i += 32) kk = fma(s_v[i], s_p[i], kk);
      kk = warp_sum(kk);
      for (int i = tid; i < m; i += TE_THREADS) s_q[i] = 2.0 * (s_p[i] - kk * s_v[i]);
    }
    __syncthreads();
    // S -= v q^T + q v^T; the reflector replaces column k below the sub-diagonal, e_k = alpha
    for (int i = grp; i < m; i += TE_THREADS / 8) {  // eight threads per row, no index arithmetic per element
      double* row = A + (size_t)(k + 1 + i) * ld + (k + 1);
      const double vi = s_v[i], qi = s_q[i];
      for (int j = sub; j < m; j += 8) row[j] -= fma(vi, s_q[j], qi * s_v[j]);
    }
    for (int i = tid; i < m; i += TE_THREADS) col[i * ld] = s_v[i];
    if (tid == 0) s_e[k] = s_scal[0];
    __syncthreads();
  }
  stamp(0);
  for (int i = tid; i < n; i += TE_THREADS) s_d[i] = A[i * ld + i];
  if (tid == 0 && n >= 2) s_e[n - 2] = A[(size_t)(n - 1) * ld + (n - 2)];
  __syncthreads();
  for (int i = tid; i < n; i += TE_THREADS) s_e2[i] = i + 1 < n ? s_e[i] * s_e[i] : 0.0;
  // Gershgorin bounds of the spectrum
  {
    double lo = 1e300, hi = -1e300;
    for (int i = tid; i < n; i += TE_THREADS) {
      const double r = (i > 0 ? fabs(s_e[i - 1]) : 0.0) + (i + 1 < n ? fabs(s_e[i]) : 0.0);
      lo = fmin(lo, s_d[i] - r), hi = fmax(hi, s_d[i] + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(SCF_FULL, lo, o));
      hi = fmax(hi, __shfl_xor_sync(SCF_FULL, hi, o));
    }
    __syncthreads();  // s_e2 complete, s_red free
    if (lane == 0) s_red[warp] = lo;
    __syncthreads();
    if (tid == 0) {
      double l = 1e300;
      for (int w = 0; w < TE_THREADS / 32; ++w) l = fmin(l, s_red[w]);
      s_scal[1] = l;
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = hi;
    __syncthreads();
    if (tid == 0) {
      double h2 = -1e300;
      for (int w = 0; w < TE_THREADS / 32; ++w) h2 = fmax(h2, s_red[w]);
      const double span = fmax(h2 - s_scal[1], 1e-300);
      s_scal[2] = h2 + 1e-12 * span, s_scal[1] -= 1e-12 * span;
    }
    __syncthreads();
  }
  // ---- eigenvalue j (ascending): one thread per eigenvalue, the bracket is cut in four per round (three interleaved
  //      Sturm counts); 29 rounds = 58 bits below the Gershgorin span.  More threads per eigenvalue (multisection over
  //      eight lanes) were measured slower: the SM's FP64 rate, not latency, then sets the time ----
  if (tid < n) {
    const int j = tid;
    double lo = s_scal[1], hi = s_scal[2];
    for (int it = 0; it < 29; ++it) {
      const double w = 0.25 * (hi - lo);
      const double x[3] = {lo + w, lo + 2.0 * w, lo + 3.0 * w};
      int cnt[3];
      sturm_count3(s_d, s_e2, n, x, cnt);
      // count(x) <= j  <=>  x <= lambda_j: the bracket becomes the quarter that holds lambda_j
      const int q = (cnt[0] <= j) + (cnt[1] <= j) + (cnt[2] <= j);
      const double nlo = q == 0 ? lo : (q == 1 ? x[0] : (q == 2 ? x[1] : x[2]));
      hi = q == 0 ? x[0] : (q == 1 ? x[1] : (q == 2 ? x[2] : hi));
      lo = nlo;
    }
    s_lam[j] = 0.5 * (lo + hi);
  }
  __syncthreads();
  stamp(1);

  // ---- eigenvectors of the tridiagonal matrix by inverse iteration: thread j solves (T - lam_j) x = b three times ----
  if (tid < n) {
    const int j = tid;
    const double lam = s_lam[j];
    double dd[TE_MAXN], du[TE_MAXN], du2[TE_MAXN], x[TE_MAXN];
    unsigned long long st = 0x9E3779B97F4A7C15ull * (unsigned long long)(j + 1);
    for (int i = 0; i < n; ++i) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      x[i] = (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
    const double tiny = 1e-300;
    for (int it = 0; it < 2; ++it) {  // two solves reach the accuracy of three (measured; LAPACK's dstein stops on growth)
      // forward elimination with partial pivoting (row i against row i + 1), applied to the right-hand side as it goes
      double di = s_d[0] - lam, ui = n > 1 ? s_e[0] : 0.0, bi = x[0];
      for (int i = 0; i + 1 < n; ++i) {
        const double li = s_e[i];                       // sub-diagonal entry of row i + 1
        double dn = s_d[i + 1] - lam;                   // diagonal of row i + 1
        const double un = i + 2 < n ? s_e[i + 1] : 0.0; // super-diagonal of row i + 1
        double bn = x[i + 1];
        if (fabs(di) >= fabs(li)) {
          if (di == 0.0) di = tiny;
          const double mlt = li / di;
          dd[i] = di, du[i] = ui, du2[i] = 0.0, x[i] = bi;
          di = dn - mlt * ui, ui = un, bi = bn - mlt * bi;
        } else {  // swap the rows
          const double mlt = di / li;
          dd[i] = li, du[i] = dn, du2[i] = un, x[i] = bn;
          di = ui - mlt * dn, ui = -mlt * un, bi = bi - mlt * bn;
        }
      }
      if (di == 0.0) di = tiny;
      dd[n - 1] = di, x[n - 1] = bi;
      // back substitution
      x[n - 1] = x[n - 1] / dd[n - 1];
      if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / dd[n - 2];
      for (int i = n - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
      // normalise (scaled against overflow: the solution of a nearly singular system is huge)
      double mx = 0.0;
      for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(x[i]));
      const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
      double s2 = 0.0;
      for (int i = 0; i < n; ++i) {
        x[i] *= sc;
        s2 = fma(x[i], x[i], s2);
      }
      const double inv = rsqrt(s2);
      for (int i = 0; i < n; ++i) x[i] *= inv;
    }
    for (int i = 0; i < n; ++i) work[(size_t)i * n + j] = x[i];
  }
  __syncthreads();
  stamp(2);

  // ---- back-transformation S = H_0 ... H_{n-3} X: eight threads hold one column in registers ----
  constexpr int RMAX = TE_MAXN / 8;
  for (int c0 = 0; c0 < n; c0 += TE_THREADS / 8) {
    const int c = c0 + grp;
    double xr[RMAX];
#pragma unroll
    for (int t = 0; t < RMAX; ++t) {
      const int r = sub + 8 * t;
      xr[t] = (c < n && r < n) ? work[(size_t)r * n + c] : 0.0;
    }
    // reflector k touches the rows r > k only: for k in [8 T + 6, 8 T + 13] the register rows t < T (r = sub + 8 t <= k
    // for every lane) are skipped at compile time (bt_range<T>), which halves the work of the triangular sweep
    bt_sweep<RMAX - 1, RMAX>(xr, A, ld, n, sub, gmask);
    if (c < n) {
      const int cc = descending ? n - 1 - c : c;
#pragma unroll
      for (int t = 0; t < RMAX; ++t) {
        const int r = sub + 8 * t;
        if (r < n) evecs[(size_t)r * ldv + cc] = xr[t];
      }
      if (sub == 0) evals[cc] = s_lam[c];
    }
  }
  __syncthreads();  // every thread is done with the reflectors in A; evecs written by this CTA are visible to it
  __threadfence_block();
  stamp(3);
  // ---- orthogonality check: S^T S = I to 1e-9, else the caller's Jacobi kernel takes over ----
  for (int e = tid; e < n * n; e += TE_THREADS) {
    const int r = e / n, c = e - r * n;
    A[r * ld + c] = evecs[(size_t)r * ldv + c];
  }
  __syncthreads();
  double worst = 0.0;
  for (int e = tid; e < n * n; e += TE_THREADS) {
    const int i = e / n, j = e - i * n;
    if (j < i) continue;
    double s = 0.0;
    for (int r = 0; r < n; ++r) s = fma(A[r * ld + i], A[r * ld + j], s);
    worst = fmax(worst, fabs(s - (i == j ? 1.0 : 0.0)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(SCF_FULL, worst, o));
  if (lane == 0) s_red[warp] = worst;
  __syncthreads();
  if (tid == 0) {
    double w = 0.0;
    for (int i = 0; i < TE_THREADS / 32; ++i) w = fmax(w, s_red[i]);
    *ok = (w <= 1e-9) ? 1 : 0;  // NaN fails the comparison
  }
  stamp(4);
}

}  // namespace

// evals / evecs as scf_sym_eig_jacobi (descending != 0: descending order); work: n * n doubles of device scratch;
// ok (device int): 1 when the eigenvectors are orthonormal to 1e-9, 0 when the caller has to fall back
int32_t tridiag_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int descending,
                           double* work, int* ok, cudaStream_t stream) {
  SCF_ARG(a && evals && evecs && work && ok, "null pointer");
  SCF_ARG(n >= 3 && n <= TE_MAXN && lda >= n && ldv >= n, "n must be within [3, 160]");
  const size_t smem = (size_t)n * (n + 1) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(tridiag_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_eig_topk(tridiagonal): %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  tridiag_eig_kernel<<<1, TE_THREADS, smem, stream>>>(a, n, lda, evals, evecs, ldv, descending, work, ok);
  if (getenv("SCF_EIG_DEBUG")) {
    long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_te_clk, sizeof(h));
    fprintf(stderr, "[tridiag n=%d] cycles: tridiagonalise %lld, multisection %lld, inverse iteration %lld, "
            "back-transform %lld, check %lld\n", n, h[0], h[1], h[2], h[3], h[4]);
  }
  return scf_check_launch("scf_eig_topk(tridiagonal)");
}

extern "C" int32_t scf_sym_eig_tridiag(const double* a, int32_t n, int64_t lda, double* evals, double* evecs, int64_t ldv,
                                       double* work, int32_t* ok, void* stream) {
  return tridiag_eig_launch(a, n, lda, evals, evecs, ldv, 0, work, ok, (cudaStream_t)stream);
}
