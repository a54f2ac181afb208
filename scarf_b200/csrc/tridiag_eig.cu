// All eigenpairs of a small symmetric matrix (n <= 160), the fast path of the Rayleigh-Ritz steps of scf_eig_topk
// (eig_topk.cu), in four launches:
//   tridiag_eig_kernel    one CTA: Householder tridiagonalisation (matrix in registers for n <= 128)
//   tridiag_vec_kernel    a warp per eigenvalue over the whole GPU: multisection on the Sturm count (32 points per round),
//                         eigenvector of the tridiagonal matrix by inverse iteration (Gaussian elimination with partial
//                         pivoting as in LAPACK dstein / dlagtf)
//   tridiag_back_kernel   a warp per eigenvector: back-transformation with the stored reflectors
//   tridiag_check_kernel  orthogonality of the result
// The one-sided Jacobi kernel (jacobi_eig.cu) needs 2.3 ms at n = 128: a Jacobi sweep is n - 1 dependent steps of ~2,500
// cycles on one SM and eight sweeps are needed.  Inverse iteration without re-orthogonalisation is accurate while the
// eigenvalues are separated by more than ~1e-7 |T| (cross-contamination eps |T| / gap); the check kernel measures the
// orthogonality of the result and reports it in `ok`: the caller then runs the Jacobi kernel, which returns at once
// when *ok == 1.  Deterministic: fixed reduction orders, seeded start vectors.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

__device__ long long g_te_clk[12];  // developer timing (SCF_EIG_DEBUG=1): cycles of the phases of the last launch
constexpr int TE_MAXN = 160;
constexpr int TE_THREADS = 1024;

__device__ __forceinline__ double group8_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// NT > 0: n <= 8 NT <= 128, the matrix lives in REGISTERS during the tridiagonalisation (eight threads per row, thread
// (row i, sub) holds the columns sub + 8 t, t < NT); NT == 0: n <= 160, the matrix in shared memory ([n][n + 1] doubles
// of dynamic shared memory).  Reflector k is stored as H_k = I - tau_k u u^T: u in vwork[k * n ..], tau_k in tauwork[k].
template <int NT>
__global__ void __launch_bounds__(TE_THREADS, 1) tridiag_eig_kernel(const double* __restrict__ a, int n, int64_t lda,
                                                                   double* __restrict__ dework,
                                                                   double* __restrict__ vwork,
                                                                   double* __restrict__ tauwork, int* __restrict__ ok) {
  extern __shared__ __align__(16) double A[];  // [n][n + 1] (NT == 0)
  __shared__ double s_d[TE_MAXN], s_e[TE_MAXN], s_v[TE_MAXN], s_p[TE_MAXN], s_q[TE_MAXN];
  __shared__ double s_x[2][128];  // NT > 0: the pivot column of the current / next step
  __shared__ double s_scal[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ld = n + 1;
  const int grp = tid >> 3, sub = tid & 7;
  const unsigned gmask = 0xFFu << (lane & ~7);
  long long t_prev = clock64();
  auto stamp = [&](int slot) {
    if (tid == 0) {
      const long long t = clock64();
      g_te_clk[slot] = t - t_prev;
      t_prev = t;
    }
  };
  if constexpr (NT > 0) {
    // ---- Householder tridiagonalisation, matrix in registers: step k applies H_k = I - tau u u^T (u = x - alpha e_1, x the
    //      column k below the diagonal) from both sides to the trailing block S.  Four phases, a CTA barrier after each:
    //        1  warp 0: |x|^2, alpha, tau, u -> shared memory          3  warp 0: K = tau / 2 u^T p, q = p - K u
    //        2  p = tau S u (eight threads per row)                     4  S -= u q^T + q u^T; the owners of column k + 1
    //                                                                      publish it as the next step's x
    //      The vector FP64 rate of the SM (~16 lanes per clock on this part) bounds phases 2 and 4; scalars are therefore
    //      computed ONCE (by warp 0) -- every warp computing them for itself to save the barriers costs more FP64 issue
    //      slots than the matrix work.  Shared memory sees vectors only (eight distinct words per warp and load); the
    //      shared-memory version below spends its time in bank conflicts on the matrix rows.
    const int i = grp;  // this thread's row
    double am[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = sub + 8 * t;
      am[t] = (i < n && j < n) ? 0.5 * (a[(int64_t)i * lda + j] + a[(int64_t)j * lda + i]) : 0.0;
    }
    double* s_u = s_v;  // names of the shared-memory version's vectors reused
    if (tid < 128) s_x[0][tid] = 0.0, s_x[1][tid] = 0.0;
    for (int j = tid; j < TE_MAXN; j += TE_THREADS) s_p[j] = 0.0, s_q[j] = 0.0, s_u[j] = 0.0;
    __syncthreads();
    if (sub == 0 && i < n) s_x[0][i] = am[0];  // column 0
    __syncthreads();
    for (int k = 0; k + 2 < n; ++k) {
      const double* x = s_x[k & 1];
      const int k1 = k + 1;
      const int t0 = k1 >> 3;                   // the columns of the register slots t < t0 are all <= k: dead
      const bool live_warp = 4 * warp + 3 > k;  // some row of this warp is > k
      // ---- 1 ----
      if (warp == 0) {
        double s = 0.0;
        for (int r = k1 + lane; r < n; r += 32) s = fma(x[r], x[r], s);
        s = warp_sum(s);
        const double x0 = x[k1];
        double alpha = 0.0, tau = 0.0;
        if (s > 0.0) {
          const double rs = rsqrt(s);
          double sq = s * rs;
          sq = fma(fma(-sq, sq, s), 0.5 * rs, sq);  // sqrt(s): one Newton step on s * rsqrt(s)
          alpha = x0 >= 0.0 ? -sq : sq;
          tau = 1.0 / fma(fabs(x0), sq, s);         // 2 / |u|^2, |u|^2 = 2 (s + |x0| sqrt(s))
        }
        for (int r = k + lane; r < n; r += 32) {
          const double ur = r > k1 ? x[r] : (r == k1 ? x0 - alpha : 0.0);
          s_u[r] = ur;
          if (r > k) vwork[(size_t)k * n + (r - k1)] = ur;  // reflector k: rows k + 1 ...
        }
        if (lane == 0) s_scal[0] = tau, s_e[k] = alpha, tauwork[k] = tau;
      }
      __syncthreads();
      // ---- 2 ----
      const double tau = s_scal[0];
      if (live_warp) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < t0) continue;
          const double uj = s_u[sub + 8 * t];  // zero for the columns <= k
          if (t & 1) a1 = fma(am[t], uj, a1);
          else a0 = fma(am[t], uj, a0);
        }
        const double pi = tau * group8_sum(a0 + a1, gmask);
        if (sub == 0) s_p[i] = i > k ? pi : 0.0;
      }
      __syncthreads();
      // ---- 3 ----
      if (warp == 0) {
        double kk = 0.0;
        for (int r = k1 + lane; r < n; r += 32) kk = fma(s_u[r], s_p[r], kk);
        kk = 0.5 * tau * warp_sum(kk);
        for (int r = k + lane; r < n; r += 32) s_q[r] = r > k ? fma(-kk, s_u[r], s_p[r]) : 0.0;
      }
      __syncthreads();
      // ---- 4 ----
      if (live_warp) {
        const double ui = i > k ? s_u[i] : 0.0, qi = i > k ? s_q[i] : 0.0;
        const int tsel = (sub == (k1 & 7)) ? t0 : -1;  // the register slot that holds column k + 1
        double* xn = s_x[k1 & 1];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < t0) continue;
          const int j = sub + 8 * t;
          am[t] = fma(-ui, s_q[j], fma(-qi, s_u[j], am[t]));  // u, q are zero for the columns <= k
          if (t == tsel && i > k1) xn[i] = am[t];             // column k + 1 of the rows below: the next step's x
        }
      }
      __syncthreads();
    }
    stamp(0);
    // the tridiagonal matrix: the diagonal entries are final once their row has left the trailing block
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = sub + 8 * t;
      if (i < n && j == i) s_d[i] = am[t];
      if (i == n - 1 && j == n - 2) s_e[n - 2] = am[t];
    }
  } else {
    for (int e = tid; e < n * n; e += TE_THREADS) {
      const int r = e / n, c = e - r * n;
      A[r * ld + c] = 0.5 * (a[(int64_t)r * lda + c] + a[(int64_t)c * lda + r]);
    }
    __syncthreads();
    __syncthreads();
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
      for (int i = tid; i < m; i += TE_THREADS) vwork[(size_t)k * n + i] = s_v[i];  // reflector k: rows k + 1 + i
      if (tid == 0) s_e[k] = s_scal[0], tauwork[k] = 2.0;  // unit vector: H = I - 2 v v^T
      __syncthreads();
    }
    stamp(0);
    for (int i = tid; i < n; i += TE_THREADS) s_d[i] = A[i * ld + i];
    if (tid == 0 && n >= 2) s_e[n - 2] = A[(size_t)(n - 1) * ld + (n - 2)];
  }
  __syncthreads();
  for (int i = tid; i < n; i += TE_THREADS) dework[i] = s_d[i], dework[n + i] = i + 1 < n ? s_e[i] : 0.0;
  if (tid == 0) *ok = 1;  // cleared by the check kernel
}

// Eigenpairs of the tridiagonal matrix (d, e) over the whole GPU: one WARP per eigenvalue j (ascending).
//   eigenvalue   multisection on the Sturm count: the 32 lanes evaluate the count at 32 interior points of the bracket,
//                which shrinks 33-fold per round; 12 rounds = 60 bits below the Gershgorin span.  One sequence per lane,
//                so the phase runs at the latency of the dependent FMA chain (n steps per round) instead of the FP64
//                issue rate of one SM: inside the single CTA (one thread per eigenvalue, 29 rounds of three counts)
//                it took 0.24 ms at n = 128.
//   eigenvector  inverse iteration by lane 0 (Gaussian elimination with partial pivoting as in LAPACK dstein / dlagtf,
//                two solves from a seeded start vector), work arrays in shared memory.
// number of eigenvalues of the tridiagonal (d, e2 = e^2) below x: sign changes of the characteristic polynomial
// recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (no division: one FMA on the dependent chain per row).  The sign
// bookkeeping -- an exact zero takes the sign opposite to its predecessor, i.e. counts as a change: the next value,
// -e^2 p_{i-1}, then has that sign and adds none -- runs on the high words in the integer pipe.  e2 must be > 0 (clamped
// by the caller): with an exact zero a vanishing value would stay zero.  Rescaled every eight rows against overflow.
__device__ __forceinline__ int sturm_count1(const double* __restrict__ d, const double* __restrict__ e2, int n, double x) {
  double p0 = 1.0, p1 = d[0] - x;
  int h1 = __double2hiint(p1);
  int cnt = (unsigned)h1 >> 31;
  for (int i = 1; i < n; ++i) {
    const double p2 = fma(d[i] - x, p1, -e2[i - 1] * p0);
    int h2 = __double2hiint(p2);
    const bool zero = (((unsigned)h2 << 1) | (unsigned)__double2loint(p2)) == 0u;
    h2 = zero ? (h1 ^ (int)0x80000000) : h2;
    cnt += (unsigned)(h2 ^ h1) >> 31;
    p0 = p1, p1 = p2, h1 = h2;
    if ((i & 7) == 0) {
      const unsigned ex = ((unsigned)h1 >> 20) & 0x7FFu;  // biased exponent: rescale outside 2^+-332 (~1e+-100)
      if (ex > 1023u + 332u || (ex < 1023u - 332u && ex != 0u)) {
        const double sc = ex > 1023u ? 1e-100 : 1e100;
        p0 *= sc, p1 *= sc;
      }
    }
  }
  return cnt;
}

constexpr int TV_WARPS = 4;
__global__ void __launch_bounds__(32 * TV_WARPS) tridiag_vec_kernel(int n, const double* __restrict__ dework,
                                                                   double* __restrict__ evals, int descending,
                                                                   double* __restrict__ xwork) {
  __shared__ double s_d[TE_MAXN], s_e[TE_MAXN], s_e2[TE_MAXN];
  __shared__ double s_w[TV_WARPS][4][TE_MAXN];  // per warp: dd, du, du2, x of the elimination
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += blockDim.x) s_d[i] = dework[i], s_e[i] = dework[n + i];
  __syncthreads();
  // Gershgorin bounds of the spectrum (every warp for itself: the same values in every warp of every CTA)
  double lo = 1e300, hi = -1e300;
  for (int i = lane; i < n; i += 32) {
    const double r = (i > 0 ? fabs(s_e[i - 1]) : 0.0) + (i + 1 < n ? fabs(s_e[i]) : 0.0);
    lo = fmin(lo, s_d[i] - r), hi = fmax(hi, s_d[i] + r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(SCF_FULL, lo, o));
    hi = fmax(hi, __shfl_xor_sync(SCF_FULL, hi, o));
  }
  const double span = fmax(hi - lo, 1e-300);
  hi += 1e-12 * span, lo -= 1e-12 * span;
  // squared off-diagonals, kept above (1e-18 span)^2: moves no eigenvalue by more than 1e-18 span and keeps the
  // recurrence of the Sturm count from sticking at zero when the matrix splits exactly
  for (int i = tid; i < n; i += blockDim.x) s_e2[i] = i + 1 < n ? fmax(s_e[i] * s_e[i], 1e-36 * span * span) : 0.0;
  __syncthreads();
  const int j = blockIdx.x * TV_WARPS + warp;
  if (j >= n) return;
  for (int it = 0; it < 12; ++it) {
    const double w = (hi - lo) * (1.0 / 33.0);
    const int cnt = sturm_count1(s_d, s_e2, n, fma((double)(lane + 1), w, lo));
    // count(x) <= j  <=>  x <= lambda_j; the counts are monotone in x: the lanes that satisfy it are a prefix
    const int q = __popc(__ballot_sync(SCF_FULL, cnt <= j));
    const double nlo = q == 0 ? lo : fma((double)q, w, lo);
    hi = q == 32 ? hi : fma((double)(q + 1), w, lo);
    lo = nlo;
  }
  const double lam = 0.5 * (lo + hi);
  if (lane == 0) {
    double* dd = s_w[warp][0];
    double* du = s_w[warp][1];
    double* du2 = s_w[warp][2];
    double* x = s_w[warp][3];
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
    for (int i = 0; i < n; ++i) xwork[(size_t)i * n + j] = x[i];
    evals[descending ? n - 1 - j : j] = lam;
  }
}

// Back-transformation S = H_0 ... H_{n-3} X over the whole GPU: one warp per column (lane l holds the rows l + 32 t),
// four columns per CTA.  The reflectors are staged in shared memory first -- read from global memory inside the loop, every reflector cost an L2 round trip (~1,000 cycles each, 63 us
// at n = 128, measured).  Each reflector then costs the warp one dot product and one update: 2 ROWS FMAs per lane and a
// shuffle reduction, n - 2 of them in sequence.
template <int ROWS>
__global__ void __launch_bounds__(128) tridiag_back_kernel(int n, const double* __restrict__ xwork,
                                                           const double* __restrict__ vwork,
                                                           const double* __restrict__ tauwork,
                                                           double* __restrict__ evecs, int64_t ldv, int descending) {
  extern __shared__ __align__(16) double s_vt[];  // [n - 2][n]: reflector k in row k (n - k - 1 entries used)
  double* s_tau = s_vt + (size_t)(n - 2) * n;
  // all copies in flight at once (cp.async, 8 bytes each): the staging costs one L2 round trip, not one per element
  for (int e = threadIdx.x; e < (n - 2) * n; e += blockDim.x) {
    const int k = e / n, i = e - k * n;
    if (i < n - k - 1)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(s_vt + e)),
                   "l"(vwork + e)
                   : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  for (int k = threadIdx.x; k < n - 2; k += blockDim.x) s_tau[k] = tauwork[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n) return;
  double xr[ROWS];
#pragma unroll
  for (int t = 0; t < ROWS; ++t) {
    const int r = lane + 32 * t;
    xr[t] = r < n ? xwork[(size_t)r * n + c] : 0.0;
  }
  for (int k = n - 3; k >= 0; --k) {
    const double* v = s_vt + (size_t)k * n - (k + 1);  // v[r] = entry at row r (r > k)
    double vr[ROWS];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < ROWS; ++t) {
      const int r = lane + 32 * t;
      vr[t] = (r > k && r < n) ? v[r] : 0.0;
      dot = fma(vr[t], xr[t], dot);
    }
    dot = s_tau[k] * warp_sum(dot);
#pragma unroll
    for (int t = 0; t < ROWS; ++t) xr[t] = fma(-dot, vr[t], xr[t]);
  }
  const int cc = descending ? n - 1 - c : c;
#pragma unroll
  for (int t = 0; t < ROWS; ++t) {
    const int r = lane + 32 * t;
    if (r < n) evecs[(size_t)r * ldv + cc] = xr[t];
  }
}

// Orthogonality check: S^T S = I to 1e-9, else *ok = 0 and the caller's Jacobi kernel takes over.  One thread per pair
// (i <= j) of columns.
__global__ void __launch_bounds__(256) tridiag_check_kernel(int n, const double* __restrict__ evecs, int64_t ldv,
                                                            int* __restrict__ ok) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = e / n, j = e - i * n;
  if (i >= n || j < i) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = 0;
  for (; r + 16 <= n; r += 16) {  // sixteen rows (32 loads) in flight: the loop is bound by the L2 latency
    double vi[16], vj[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      vi[u] = evecs[(size_t)(r + u) * ldv + i];
      vj[u] = evecs[(size_t)(r + u) * ldv + j];
    }
#pragma unroll
    for (int u = 0; u < 16; u += 4) {
      s0 = fma(vi[u], vj[u], s0);
      s1 = fma(vi[u + 1], vj[u + 1], s1);
      s2 = fma(vi[u + 2], vj[u + 2], s2);
      s3 = fma(vi[u + 3], vj[u + 3], s3);
    }
  }
  for (; r < n; ++r) s0 = fma(evecs[(size_t)r * ldv + i], evecs[(size_t)r * ldv + j], s0);
  const double dev = fabs((s0 + s1) + (s2 + s3) - (i == j ? 1.0 : 0.0));
  if (!(dev <= 1e-9)) *ok = 0;  // NaN fails the comparison
}

}  // namespace

// evals / evecs as scf_sym_eig_jacobi (descending != 0: descending order); work: 2 * n * n + 3 * n doubles of device scratch;
// ok (device int): 1 when the eigenvectors are orthonormal to 1e-9, 0 when the caller has to fall back
int32_t tridiag_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int descending,
                           double* work, int* ok, cudaStream_t stream) {
  SCF_ARG(a && evals && evecs && work && ok, "null pointer");
  SCF_ARG(n >= 3 && n <= TE_MAXN && lda >= n && ldv >= n, "n must be within [3, 160]");
  double* xwork = work;
  double* vwork = work + (size_t)n * n;
  double* tauwork = vwork + (size_t)n * n;
  double* dework = tauwork + n;  // diagonal and off-diagonal of the tridiagonal matrix
  if (n <= 128) {  // matrix in registers
    auto fn = n <= 64 ? tridiag_eig_kernel<8> : (n <= 96 ? tridiag_eig_kernel<12> : tridiag_eig_kernel<16>);
    fn<<<1, TE_THREADS, 0, stream>>>(a, n, lda, dework, vwork, tauwork, ok);
  } else {
    const size_t smem = (size_t)n * (n + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(tridiag_eig_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      scf_set_error("scf_eig_topk(tridiagonal): %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
    tridiag_eig_kernel<0><<<1, TE_THREADS, smem, stream>>>(a, n, lda, dework, vwork, tauwork, ok);
  }
  tridiag_vec_kernel<<<(n + TV_WARPS - 1) / TV_WARPS, 32 * TV_WARPS, 0, stream>>>(n, dework, evals, descending, xwork);
  {
    const size_t bsmem = ((size_t)(n - 2) * n + n) * sizeof(double);
    auto bk = n <= 128 ? tridiag_back_kernel<4> : tridiag_back_kernel<5>;
    cudaError_t e = cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
    if (e != cudaSuccess) {
      scf_set_error("scf_eig_topk(tridiagonal): %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
    bk<<<(n + 3) / 4, 128, bsmem, stream>>>(n, xwork, vwork, tauwork, evecs, ldv, descending);
  }
  tridiag_check_kernel<<<(n * n + 255) / 256, 256, 0, stream>>>(n, evecs, ldv, ok);
  if (getenv("SCF_EIG_DEBUG")) {
    long long h[12];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_te_clk, sizeof(h));
    fprintf(stderr, "[tridiag n=%d] cycles: tridiagonalise %lld\n", n, h[0]);
  }
  return scf_check_launch("scf_eig_topk(tridiagonal)");
}

extern "C" int32_t scf_sym_eig_tridiag(const double* a, int32_t n, int64_t lda, double* evals, double* evecs, int64_t ldv,
                                       double* work, int32_t* ok, void* stream) {
  return tridiag_eig_launch(a, n, lda, evals, evecs, ldv, 0, work, ok, (cudaStream_t)stream);
}
