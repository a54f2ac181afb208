// K3: top-`dims` eigenpairs of the PCA covariance -- what sklearn's IncrementalPCA fit stands for on this path
// (scarf/ann.py:207-256) -- entirely in this library: no cuSOLVER, no cuBLAS, no framework kernels.
//
// Chebyshev-filtered subspace iteration (Zhou & Saad) on a block of b = dims + 32 columns, all arithmetic FP64:
//   start     X = seeded random block; a few filter steps of degree 3, each followed by a Cholesky-QR2
//   round     Rayleigh-Ritz:  AQ = C Q,  T = Q^T AQ,  T = S diag(theta) S^T (tridiag_eig.cu; Jacobi as the checked
//             fallback),  V = Q S,  AV = AQ S
//             residual max_j |AV_j - theta_j V_j| / theta_1 over the wanted pairs;  converged -> sign rule -> done
//             else filter of degree m (largest m with rho^m <= 1e20, rho from the Ritz values) applied to V, whose
//             first step needs no product with C (C V = AV is there already), then Cholesky-QR2
// The host runs the schedule and synchronises the stream once per round to read the residual and the Ritz values
// (a few hundred bytes); everything else stays on the device.
//
// Kernels (all deterministic: fixed-order reductions, no floating-point atomics -- every rank of a sharded run gets
// bit-identical loadings from its bit-identical copy of the Gram matrix):
//   eig_cov_kernel        covariance from the fixed-point Gram (optionally centred on a column mean)
//   eig_dgemm_mma_kernel  out = alpha A B + gamma P + delta Q (DMMA), 64-row tiles x all columns, split K with an ordered
//                         reduction by the last CTA of a tile (serves the filter steps and the tall x small rotations)
//   eig_gram_kernel       S = X^T Y over row chunks; eig_gram_reduce_kernel adds the partial tiles in chunk order
//   eig_chol_kernel       one CTA: column scaling, blocked Cholesky of the b x b Gram in shared memory, triangular inverse
//                         (one column per group of eight threads) -> the factor W with (X W)^T (X W) = I
//   tridiag_* / jacobi_eig_kernel   (tridiag_eig.cu, jacobi_eig.cu) all eigenpairs of the b x b Rayleigh-Ritz matrix
//   eig_resid_* / eig_sign_kernel   residual norms and sklearn's sign rule
#include <math.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"

int32_t jacobi_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int* info,
                          int descending, const int* skip_if, cudaStream_t stream);
int32_t tridiag_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int descending,
                           double* work, int* ok, cudaStream_t stream);

namespace {

constexpr int EG_MAXB = 160;       // widest block: 160 x 161 doubles of shared memory in the Cholesky kernel
constexpr int EG_BUFFER = 32;      // columns carried beyond the wanted ones
constexpr int EG_TM = 64;          // rows of a GEMM tile
constexpr int EG_KT = 16;          // K step of the GEMM
constexpr int EG_THREADS = 256;

// ---------------------------------------------------------------------------------------------- covariance
// cov = gram_fx * scale - mean_w * mean mean^T   (mean may be null), [h, ldc] row major, pad columns zero
__global__ void __launch_bounds__(256) eig_cov_kernel(const long long* __restrict__ g, int64_t ldg, int h, double scale,
                                                      const double* __restrict__ mean, double mean_w,
                                                      double* __restrict__ cov, int64_t ldc) {
  const int64_t total = (int64_t)h * ldc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldc, c = e - r * ldc;
    double v = 0.0;
    if (c < h) {
      v = (double)g[r * ldg + c] * scale;
      if (mean) v -= mean_w * mean[r] * mean[c];
    }
    cov[e] = v;
  }
}

// trace and the largest absolute row sum (an upper bound of the largest eigenvalue): norms[0] = trace, norms[1] = bound
__global__ void __launch_bounds__(256) eig_rowsum_kernel(const double* __restrict__ cov, int64_t ldc, int h,
                                                         double* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= h) return;
  double s = 0.0;
  for (int c = lane; c < h; c += 32) s += fabs(cov[(int64_t)r * ldc + c]);
  s = warp_sum(s);
  if (lane == 0) rowsum[r] = s;
}
__global__ void __launch_bounds__(1024) eig_norms_kernel(const double* __restrict__ cov, int64_t ldc, int h,
                                                         const double* __restrict__ rowsum, double* __restrict__ norms) {
  __shared__ double s_tr[32], s_mx[32];
  double tr = 0.0, mx = 0.0;
  for (int r = threadIdx.x; r < h; r += 1024) {
    tr += cov[(int64_t)r * ldc + r];
    mx = fmax(mx, rowsum[r]);
  }
  tr = warp_sum(tr);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(SCF_FULL, mx, o));
  if ((threadIdx.x & 31) == 0) s_tr[threadIdx.x >> 5] = tr, s_mx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, m = 0.0;
    for (int w = 0; w < 32; ++w) t += s_tr[w], m = fmax(m, s_mx[w]);
    norms[0] = t, norms[1] = m;
  }
}

// seeded start block: standard normals from a counter hash (splitmix64 + Box-Muller); pad columns zero
__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) eig_init_kernel(double* __restrict__ x, int h, int b, int64_t ldb,
                                                       unsigned long long seed) {
  const int64_t total = (int64_t)h * ldb;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldb, c = e - r * ldb;
    double v = 0.0;
    if (c < b) {
      const unsigned long long k = splitmix(seed ^ (unsigned long long)(r * (int64_t)b + c));
      const unsigned long long k2 = splitmix(k);
      const double u1 = ((double)(k >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
      const double u2 = (double)(k2 >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
      v = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    x[e] = v;
  }
}

// ---------------------------------------------------------------------------------------------- GEMM
// out[M, N] = alpha * A[M, K] B[K, N] + gamma * P[M, N] + delta * Q[M, N]      (P, Q may be null)
// All matrices row major; B, P, Q, out share the row stride ldn (a multiple of 16 >= 16 NJ, pad columns zero).
// grid = (row tiles of 64, K splits).  A CTA accumulates its K range for the whole width, writes the partial tile to
// `part` and bumps the tile's counter; the CTA that arrives last adds the partials in split order and applies the
// epilogue (deterministic).
//
// FP64 tensor path (mma.sync m8n8k4, SASS DMMA): a warp instruction does 256 FMAs where a DFMA does 32.  The SM's vector
// FP64 pipe (~16-32 lanes / clk) bounded the SIMT version of this kernel (round 2, first half) at 12.6 TFLOP/s; this
// one reaches 16.8.  Four warps per CTA, warp w owns rows 16 w .. 16 w + 15 of the 64-row tile (two 8-row
// MMA tiles) x all N columns.  Fragments (PTX ISA): A 8 x 4 row major -- lane holds A[lane / 4][lane % 4]; B 4 x 8 column
// major -- lane holds B[lane % 4][lane / 4]; C 8 x 8 -- lane holds C[lane / 4][2 (lane % 4) + {0, 1}].  Shared-memory row
// strides are = 4 (mod 16) doubles so that the 16 lanes of a half warp hit 16 distinct 8-byte banks.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int NJ>
__global__ void __launch_bounds__(128) eig_dgemm_mma_kernel(const double* __restrict__ a, int64_t lda,
                                                            const double* __restrict__ bm, int64_t ldn, int m, int k,
                                                            double alpha, const double* __restrict__ pm, double gamma,
                                                            const double* __restrict__ qm, double delta,
                                                            double* __restrict__ out, double* __restrict__ part,
                                                            unsigned int* __restrict__ counters, int k_per_split) {
  constexpr int N = NJ * 16, NT = N / 8, THREADS = 128;
  constexpr int SA = EG_KT + 4, SB = N + 4;
  constexpr int B_PER = EG_KT * N / THREADS;
  extern __shared__ __align__(16) double eg_smem[];
  double (*sa)[EG_TM][SA] = reinterpret_cast<double (*)[EG_TM][SA]>(eg_smem);
  double (*sb)[EG_KT][SB] = reinterpret_cast<double (*)[EG_KT][SB]>(eg_smem + 2 * EG_TM * SA);
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row / column
  const int row0 = blockIdx.x * EG_TM;
  const int k0 = blockIdx.y * k_per_split, k1 = min(k, k0 + k_per_split);
  double acc[2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int a_r = tid >> 1, a_c = (tid & 1) * 8;  // A tile 64 x 16: eight consecutive k of one row per thread
  double ra[8];
  double rb[B_PER];
  auto load_tiles = [&](int kk) {
    const int gr = row0 + a_r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gk = kk + a_c + i;
      ra[i] = (gr < m && gk < k1) ? a[(int64_t)gr * lda + gk] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + THREADS * i;
      const int br = e / N, bc = e - br * N;
      const int gk = kk + br;
      rb[i] = gk < k1 ? bm[(int64_t)gk * ldn + bc] : 0.0;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sa[buf][a_r][a_c + i] = ra[i];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + THREADS * i;
      const int br = e / N, bc = e - br * N;
      sb[buf][br][bc] = rb[i];
    }
  };
  int buf = 0;
  if (k0 < k1) {
    load_tiles(k0);
    store_tiles(0);
  }
  __syncthreads();
  for (int kk = k0; kk < k1; kk += EG_KT) {
    const bool more = kk + EG_KT < k1;
    if (more) load_tiles(kk + EG_KT);
#pragma unroll
    for (int k4 = 0; k4 < EG_KT / 4; ++k4) {
      const double a0 = sa[buf][16 * w + fr][4 * k4 + fc], a1 = sa[buf][16 * w + 8 + fr][4 * k4 + fc];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double b = sb[buf][4 * k4 + fc][8 * nt + fr];
        dmma884(acc[0][nt], a0, b);
        dmma884(acc[1][nt], a1, b);
      }
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const int nsplit = gridDim.y;
  if (nsplit > 1) {
    double* mine = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(EG_TM * N);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        *reinterpret_cast<double2*>(mine + ((size_t)((w * 2 + mt) * NT + nt) * 32 + lane) * 2) =
            make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(counters + blockIdx.x, 1u) == (unsigned)(nsplit - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    for (int s = 0; s < nsplit; ++s) {  // fixed order: the sum does not depend on which CTA came last
      const double* src = part + ((size_t)s * gridDim.x + blockIdx.x) * (size_t)(EG_TM * N);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(src + ((size_t)((w * 2 + mt) * NT + nt) * 32 + lane) * 2));
          acc[mt][nt][0] += v.x, acc[mt][nt][1] += v.y;
        }
    }
    if (tid == 0) counters[blockIdx.x] = 0u;  // ready for the next launch
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int gr = row0 + 16 * w + 8 * mt + fr;
    if (gr >= m) continue;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int64_t o = (int64_t)gr * ldn + 8 * nt + 2 * fc + h2;
        double v = alpha * acc[mt][nt][h2];
        if (pm) v = fma(gamma, pm[o], v);
        if (qm) v = fma(delta, qm[o], v);
        out[o] = v;
      }
  }
}

// ---------------------------------------------------------------------------------------------- Gram
// s[b, b] (row stride lds) = X^T Y over the h rows; grid = (tiles of 64 x 64 outputs, row chunks); with more than one
// chunk the partial tiles go to `part` and eig_gram_reduce_kernel adds them in chunk order.
__global__ void __launch_bounds__(EG_THREADS) eig_gram_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                              int64_t ldn, int h, int b, double* __restrict__ s,
                                                              int64_t lds, double* __restrict__ part, int rows_per_chunk) {
  __shared__ __align__(16) double sx[EG_KT][64];
  __shared__ __align__(16) double sy[EG_KT][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles = (b + 63) / 64;
  const int ti = blockIdx.x / tiles, tj = blockIdx.x - ti * tiles;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(h, r0 + rows_per_chunk);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int rr = r0; rr < r1; rr += EG_KT) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // 16 rows x 64 columns of each operand: element e = tid + 256 i
      const int e = tid + EG_THREADS * i;
      const int lr = e >> 6, lc = e & 63;
      const int gr = rr + lr;
      const int cx = ti * 64 + lc, cy = tj * 64 + lc;
      sx[lr][lc] = (gr < r1 && cx < b) ? x[(int64_t)gr * ldn + cx] : 0.0;
      sy[lr][lc] = (gr < r1 && cy < b) ? y[(int64_t)gr * ldn + cy] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < EG_KT; ++t) {
      double xv[4], yv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = sx[t][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) yv[j] = sy[t][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(xv[i], yv[j], acc[i][j]);
    }
  }
  const int nchunk = gridDim.y;
  if (nchunk > 1) {
    double* mine = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mine[(ty * 4 + i) * 64 + tx + 16 * j] = acc[i][j];
    return;  // eig_gram_reduce_kernel adds the partial tiles in chunk order
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * 64 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj * 64 + tx + 16 * j;
      if (gi < b && gj < b) s[(int64_t)gi * lds + gj] = acc[i][j];
    }
  }
}

// The partial tiles of eig_gram_kernel summed in chunk order (deterministic), one thread per output element: its
// loads are independent, so they are in flight together.  (The last CTA of a tile used to walk the ~37 chunks one L2 round
// trip after the other: 25 of the kernel's 34 us.)
__global__ void __launch_bounds__(256) eig_gram_reduce_kernel(const double* __restrict__ part, int n_tiles, int nchunk, int b,
                                                             double* __restrict__ s, int64_t lds) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // tile * 4096 + element
  if (e >= n_tiles * 4096) return;
  const int tile = e >> 12, el = e & 4095;
  const int tiles = (b + 63) / 64;
  const int ti = tile / tiles, tj = tile - ti * tiles;
  const int gi = ti * 64 + (el >> 6), gj = tj * 64 + (el & 63);
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < nchunk; ++c) acc += __ldcg(part + ((size_t)c * n_tiles + tile) * 4096 + el);
  if (gi < b && gj < b) s[(int64_t)gi * lds + gj] = acc;
}

// ---------------------------------------------------------------------------------------------- Cholesky factor
// One CTA.  s = X^T X (b x b).  d_i = 1 / sqrt(s_ii), s' = D s D = L L^T, W = D L^-T (upper triangular, [b, ldw] row
// major, columns >= b and rows >= b zero up to ldw): (X W)^T (X W) = I.  flags[0] |= 1 when a pivot is not positive
// (the block is numerically rank deficient: the caller repeats with the eigenvalue-based factor); flags[1] (as double
// bits in flagsd[0]) keeps the smallest pivot seen.
__global__ void __launch_bounds__(1024, 1) eig_chol_kernel(const double* __restrict__ s, int64_t lds, int b,
                                                           double shift, double* __restrict__ w, int64_t ldw,
                                                           int* __restrict__ flags, double* __restrict__ min_pivot) {
  extern __shared__ __align__(16) double sm[];  // [b][b + 1]: lower triangle L, upper triangle (L^-1)^T
  __shared__ double sd[EG_MAXB];      // column scaling d_i = 1 / sqrt(s_ii)
  __shared__ double s_invd[EG_MAXB];  // 1 / L_ii
  __shared__ double s_tmp[EG_MAXB];
  __shared__ int s_bad;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ld = (b + 1) | 1;  // odd row stride: rows and columns are both read without bank conflicts
  const int grp = tid >> 3, sub = tid & 7, ngrp = nt >> 3;  // eight threads share one dot product
  const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
  if (tid == 0) s_bad = 0;
  for (int i = tid; i < b; i += nt) {
    const double v = s[(int64_t)i * lds + i];
    sd[i] = v > 0.0 ? rsqrt(v) : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < b * b; e += nt) {
    const int r = e / b, c = e - r * b;
    // `shift` on the unit diagonal (first pass of an orthonormalisation): a block filtered from an orthonormal one is
    // ill conditioned up to ~1e7-1e8, its Gram up to 1e16 -- rounding alone can push a pivot below zero.  With the shift
    // the factor exists, X W comes out with a condition number of ~sqrt(1 + shift cond(X)^2) (<~ 1e3), and the second,
    // unshifted pass restores orthonormality to ~1e-10 or better (shifted Cholesky-QR, Fukaya et al.).
    sm[r * ld + c] = 0.5 * (s[(int64_t)r * lds + c] + s[(int64_t)c * lds + r]) * sd[r] * sd[c] + (r == c ? shift : 0.0);
  }
  __syncthreads();
  if (shift == 0.0) {
    // Second pass of an orthonormalisation: when the block is orthonormal to 1e-5 already (the usual case once the
    // first pass worked on a well-conditioned block), one Newton-Schulz step W = D (I - E / 2), E = S' - I, brings it
    // to ~E^2 without a factorisation: (X W)^T (X W) = I - 3/4 E^2 + O(E^3).
    double emax = 0.0;
    for (int e = tid; e < b * b; e += nt) {
      const int r = e / b, c = e - r * b;
      if (r != c) emax = fmax(emax, fabs(sm[r * ld + c]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(SCF_FULL, emax, o));
    if ((tid & 31) == 0) s_tmp[tid >> 5] = emax;
    __syncthreads();
    emax = 0.0;
    for (int w2 = 0; w2 < (nt >> 5); ++w2) emax = fmax(emax, s_tmp[w2]);
    __syncthreads();
    if (emax <= 1e-5) {
      for (int e = tid; e < b * (int)ldw; e += nt) {
        const int i = e / (int)ldw, j = e - i * (int)ldw;
        double v = 0.0;
        if (j < b) v = sd[i] * (i == j ? 1.0 : -0.5 * sm[i * ld + j]);
        w[(int64_t)i * ldw + j] = v;
      }
      return;
    }
  }
  // Blocked (left-looking) Cholesky, eight columns at a time -- three CTA barriers per block instead of two per column:
  //   (1) panel update   P = S'[jb.., jb..jb+8) - L[jb.., 0..jb) L[jb..jb+8, 0..jb)^T, one thread per entry
  //   (2) the 8 x 8 diagonal block of P factored by warp 0 (row i in lane i, columns exchanged by shuffles)
  //   (3) the rows below it solved against that block, one thread per row
  double minp = 1e300;
  for (int jb = 0; jb < b; jb += 8) {
    const int nb = min(8, b - jb);
    if (jb > 0) {
      for (int idx = tid; idx < (b - jb) * 8; idx += nt) {
        const int r = jb + (idx >> 3), cc = idx & 7;
        if (cc < nb && jb + cc <= r) {
          const double* lr = sm + (size_t)r * ld;
          const double* lc = sm + (size_t)(jb + cc) * ld;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          for (int k = 0; k < jb; k += 4) {  // jb is a multiple of 8
            a0 = fma(lr[k], lc[k], a0);
            a1 = fma(lr[k + 1], lc[k + 1], a1);
            a2 = fma(lr[k + 2], lc[k + 2], a2);
            a3 = fma(lr[k + 3], lc[k + 3], a3);
          }
          sm[(size_t)r * ld + jb + cc] -= (a0 + a1) + (a2 + a3);
        }
      }
      __syncthreads();
    }
    if (tid < 32) {
      const int lane = tid, i = lane & 7;  // lanes >= 8 mirror lanes 0..7 (every lane takes part in the shuffles)
      double d[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        d[c] = (i < nb && c <= i) ? sm[(size_t)(jb + i) * ld + jb + c] : (c == i ? 1.0 : 0.0);
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // The scaled matrix has a unit diagonal.  A filtered block may be ill conditioned up to ~1e7 (its Gram 1e14):
        // tiny positive pivots are expected in the first pass and repaired by the second; only a non-positive (or NaN)
        // pivot means the factorisation broke down.
        const double piv = __shfl_sync(SCF_FULL, d[j], j);
        const bool okp = piv > 0.0;
        if (j < nb) {
          bad |= !okp;
          minp = fmin(minp, piv);
        }
        const double inv = okp ? rsqrt(piv) : 0.0;
        const double lij = i == j ? (okp ? piv * inv : 1.0) : d[j] * inv;  // rows i < j hold zeros in d[j]
        d[j] = lij;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) {
          const double lcj = __shfl_sync(SCF_FULL, lij, c);
          if (i >= c) d[c] = fma(-lij, lcj, d[c]);
        }
      }
      __syncwarp();  // the mirror lanes have read the block before it is overwritten
      if (lane < nb) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= i) sm[(size_t)(jb + i) * ld + jb + c] = d[c];
        double dii = 1.0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c == i) dii = d[c];
        s_invd[jb + i] = __drcp_rn(dii);
      }
      if (lane == 0 && bad) s_bad = 1;
    }
    __syncthreads();
    for (int r = jb + nb + tid; r < b; r += nt) {
      double x[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = c < nb ? sm[(size_t)r * ld + jb + c] : 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < nb) {
          double a = x[c];
#pragma unroll
          for (int k = 0; k < c; ++k) a = fma(-x[k], sm[(size_t)(jb + c) * ld + jb + k], a);
          x[c] = a * s_invd[jb + c];
        }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < nb) sm[(size_t)r * ld + jb + c] = x[c];
    }
    __syncthreads();
  }
  // X = L^-1, one column per group of eight threads (no CTA barrier: the columns are independent, the groups run at
  // their own pace): X[r][c] = -(sum_{k=c}^{r-1} L[r][k] X[k][c]) / L[r][r], X[k][c] kept at sm[c][k] (upper triangle,
  // which the factor does not use), X[c][c] = s_invd[c]
  for (int c = grp; c < b; c += ngrp) {
    for (int r = c + 1; r < b; ++r) {
      double acc = 0.0;
      for (int k = c + sub; k < r; k += 8)
        acc = fma(sm[(size_t)r * ld + k], k == c ? s_invd[c] : sm[(size_t)c * ld + k], acc);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
      if (sub == 0) sm[(size_t)c * ld + r] = -acc * s_invd[r];
      __syncwarp(gmask);
    }
  }
  __syncthreads();
  // W[i][j] = d_i * (L^-1)[j][i] for i <= j  (upper triangular); (L^-1)[j][i] is stored at (i, j), its diagonal in s_invd
  for (int e = tid; e < b * (int)ldw; e += nt) {
    const int i = e / (int)ldw, j = e - i * (int)ldw;
    double v = 0.0;
    if (j < b && i <= j) v = sd[i] * (i == j ? s_invd[i] : sm[i * ld + j]);
    w[(int64_t)i * ldw + j] = v;
  }
  if (tid == 0) {
    if (s_bad) atomicOr(flags, 1);
    *min_pivot = fmin(*min_pivot, minp);
  }
}

// Eigenvalue-based factor (SVQB), the fallback for a numerically rank-deficient block: with s' = D s D = U diag(l) U^T
// (Jacobi), W = D U diag(max(l, eps l_max))^-1/2.  Built from the Jacobi output by this small kernel.
__global__ void __launch_bounds__(256) eig_svqb_scale_kernel(const double* __restrict__ s, int64_t lds, int b,
                                                             double* __restrict__ sc, int64_t ldsc,
                                                             double* __restrict__ d) {
  // sc = D s D with d_i = 1 / sqrt(s_ii)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < b * b; e += gridDim.x * blockDim.x) {
    const int r = e / b, c = e - r * b;
    const double dr = s[(int64_t)r * lds + r] > 0.0 ? rsqrt(s[(int64_t)r * lds + r]) : 0.0;
    const double dc = s[(int64_t)c * lds + c] > 0.0 ? rsqrt(s[(int64_t)c * lds + c]) : 0.0;
    sc[(int64_t)r * ldsc + c] = 0.5 * (s[(int64_t)r * lds + c] + s[(int64_t)c * lds + r]) * dr * dc;
    if (c == 0) d[r] = dr;
  }
}
__global__ void __launch_bounds__(256) eig_svqb_factor_kernel(const double* __restrict__ u, int64_t ldu,
                                                              const double* __restrict__ lam, const double* __restrict__ d,
                                                              int b, double* __restrict__ w, int64_t ldw) {
  // lam ascending; W[i][j] = d_i u[i][j] / sqrt(max(lam_j, 1e-14 lam_max))
  const double lmax = lam[b - 1];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < b * (int)ldw; e += gridDim.x * blockDim.x) {
    const int i = e / (int)ldw, j = e - i * (int)ldw;
    double v = 0.0;
    if (j < b) v = d[i] * u[(int64_t)i * ldu + j] * rsqrt(fmax(lam[j], 1e-14 * lmax));
    w[(int64_t)i * ldw + j] = v;
  }
}

// ---------------------------------------------------------------------------------------------- residual + output
// partial sums of squares of AV_j - theta_j V_j over row chunks (deterministic two-phase reduction)
__global__ void __launch_bounds__(256) eig_resid_partial_kernel(const double* __restrict__ v, const double* __restrict__ av,
                                                                int64_t ldn, int h, int dims,
                                                                const double* __restrict__ theta,
                                                                double* __restrict__ part, int rows_per_chunk) {
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(h, r0 + rows_per_chunk);
  for (int j = threadIdx.x; j < dims; j += blockDim.x) {
    const double th = theta[j];
    double s = 0.0;
    for (int r = r0; r < r1; ++r) {
      const double d = av[(int64_t)r * ldn + j] - th * v[(int64_t)r * ldn + j];
      s = fma(d, d, s);
    }
    part[(size_t)blockIdx.x * dims + j] = s;
  }
}
// host block: [0] residual (max_j |r_j| / theta_1), [1] trace, [2] 1-norm bound, [3] flags, [4] min pivot,
// [8 .. 8 + b) Ritz values (descending)
__global__ void __launch_bounds__(256) eig_resid_final_kernel(const double* __restrict__ part, int nchunk, int dims, int b,
                                                              const double* __restrict__ theta,
                                                              const double* __restrict__ norms,
                                                              const int* __restrict__ flags,
                                                              const double* __restrict__ min_pivot,
                                                              double* __restrict__ report) {
  __shared__ double s_mx[8];
  double mx = 0.0;
  for (int j = threadIdx.x; j < dims; j += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * dims + j];
    mx = fmax(mx, sqrt(s));
    if (!(s == s)) mx = (__longlong_as_double(0x7ff0000000000000LL));  // NaN must not pass as converged
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(SCF_FULL, mx, o));
  if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < 8; ++w) m = fmax(m, s_mx[w]);
    report[0] = theta[0] > 0.0 ? m / theta[0] : (__longlong_as_double(0x7ff0000000000000LL));
    report[1] = norms[0], report[2] = norms[1];
    report[3] = (double)flags[0];
    report[4] = *min_pivot;
    report[169] = (double)flags[4];  // 1: the Ritz pairs of this round came from the tridiagonal kernel
  }
  for (int j = threadIdx.x; j < b; j += blockDim.x) report[8 + j] = theta[j];
}

// sklearn's svd_flip(u_based_decision=False): the entry of largest magnitude of every component is positive (first
// such entry on ties).  One warp per component; out [h, dims] row major.
__global__ void __launch_bounds__(256) eig_sign_kernel(const double* __restrict__ v, int64_t ldn, int h, int dims,
                                                       double* __restrict__ out, float* __restrict__ out32,
                                                       int64_t ld32) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= dims) {  // pad columns of the float32 copy
    if (out32 && j < ld32)
      for (int r = lane; r < h; r += 32) out32[(int64_t)r * ld32 + j] = 0.f;
    return;
  }
  double best = -1.0;
  int bi = 0x7fffffff;
  for (int r = lane; r < h; r += 32) {
    const double a = fabs(v[(int64_t)r * ldn + j]);
    if (a > best) best = a, bi = r;  // ascending r inside a lane: the first maximum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(SCF_FULL, best, o);
    const int oi = __shfl_xor_sync(SCF_FULL, bi, o);
    if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
  }
  const double sg = v[(int64_t)bi * ldn + j] < 0.0 ? -1.0 : 1.0;
  for (int r = lane; r < h; r += 32) {
    const double x = sg * v[(int64_t)r * ldn + j];
    out[(int64_t)r * dims + j] = x;
    if (out32) out32[(int64_t)r * ld32 + j] = (float)x;
  }
}

// ---------------------------------------------------------------------------------------------- host side
constexpr int EG_NTALL = 5;  // tall (h x ldn) buffers the schedule rotates through

struct Layout {
  int b, nj, ldn;
  int64_t ldc;
  int nsplit, k_per_split, row_tiles, gram_chunks, gram_rows, gram_tiles, resid_chunks, resid_rows;
  size_t off_cov, off_tall[EG_NTALL], off_part, off_gpart, off_s, off_w, off_t, off_u, off_x, off_lam, off_theta, off_rowsum,
      off_norms, off_rpart, off_d, off_report, off_counters, off_flags, off_minp, total;
};

bool make_layout(int h, int dims, Layout& L) {
  if (h < 1 || dims < 1 || dims > h) return false;
  int b = std::min(dims + EG_BUFFER, h);
  if (b > 128 && dims + 24 <= 128) b = 128;  // 64 column pairs: the Jacobi kernel's fast configuration (16 threads a pair)
  if (b > EG_MAXB) b = std::min(EG_MAXB, h);
  if (b < dims + std::min(8, h - dims)) return false;  // dims too large for the shared-memory Cholesky
  L.b = b;
  L.nj = (b + 15) / 16;
  L.ldn = L.nj * 16;
  L.ldc = ((int64_t)h + 7) / 8 * 8;
  L.row_tiles = (h + EG_TM - 1) / EG_TM;
  // K splits of the big products: AT MOST two CTAs per SM, so that no SM gets a third one while others hold two -- the
  // kernel runs at the SM's FP64 rate, and with 32 row tiles x 10 splits = 320 CTAs on 148 SMs the SMs with three CTAs
  // set the time (62 us; 32 x 9 = 288: see DESIGN.md for the measurement)
  int want = std::max(1, (2 * SCF_NUM_SMS) / L.row_tiles);
  L.k_per_split = std::max(EG_KT * 4, ((h + want - 1) / want + EG_KT - 1) / EG_KT * EG_KT);
  L.nsplit = (h + L.k_per_split - 1) / L.k_per_split;
  L.gram_tiles = ((b + 63) / 64) * ((b + 63) / 64);
  L.gram_chunks = std::max(1, std::min((h + 63) / 64, SCF_NUM_SMS / L.gram_tiles));
  L.gram_rows = ((h + L.gram_chunks - 1) / L.gram_chunks + EG_KT - 1) / EG_KT * EG_KT;
  L.gram_chunks = (h + L.gram_rows - 1) / L.gram_rows;
  L.resid_chunks = std::min(SCF_NUM_SMS, (h + 31) / 32);
  L.resid_rows = (h + L.resid_chunks - 1) / L.resid_chunks;
  L.resid_chunks = (h + L.resid_rows - 1) / L.resid_rows;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  const size_t tall = (size_t)h * L.ldn * 8, small = (size_t)EG_MAXB * L.ldn * 8;
  L.off_cov = o, o = al(o + (size_t)h * L.ldc * 8);
  for (int i = 0; i < EG_NTALL; ++i) L.off_tall[i] = o, o = al(o + tall);
  L.off_part = o, o = al(o + (size_t)L.nsplit * L.row_tiles * EG_TM * L.ldn * 8);
  L.off_gpart = o, o = al(o + (size_t)L.gram_chunks * L.gram_tiles * 4096 * 8);
  L.off_s = o, o = al(o + small);
  L.off_w = o, o = al(o + small);
  L.off_t = o, o = al(o + small);
  L.off_u = o, o = al(o + small);
  L.off_x = o, o = al(o + (size_t)(2 * EG_MAXB * EG_MAXB + 3 * EG_MAXB) * 8);  // scratch of the tridiagonal eigensolver
  L.off_lam = o, o = al(o + EG_MAXB * 8);
  L.off_theta = o, o = al(o + EG_MAXB * 8);
  L.off_rowsum = o, o = al(o + (size_t)h * 8);
  L.off_norms = o, o = al(o + 64);
  L.off_rpart = o, o = al(o + (size_t)L.resid_chunks * EG_MAXB * 8);
  L.off_d = o, o = al(o + EG_MAXB * 8);
  L.off_report = o, o = al(o + (size_t)(8 + EG_MAXB + 8) * 8);
  L.off_counters = o, o = al(o + (size_t)(L.row_tiles + L.gram_tiles + 8) * 4);
  L.off_flags = o, o = al(o + 64);
  L.off_minp = o, o = al(o + 64);
  L.total = o;
  return true;
}

struct Ctx {
  Layout L;
  int h, dims;
  unsigned char* ws;
  cudaStream_t st;
  mutable int launches = 0;  // kernels launched by this call (reported in host_report[168])
  double* p(size_t off) const { return reinterpret_cast<double*>(ws + off); }
  double* tall(int i) const { return p(L.off_tall[i]); }
  // a tall buffer that is none of the given ones
  double* spare(const double* a = nullptr, const double* b2 = nullptr, const double* c = nullptr,
                const double* d = nullptr) const {
    for (int i = 0; i < EG_NTALL; ++i) {
      double* t = tall(i);
      if (t != a && t != b2 && t != c && t != d) return t;
    }
    return nullptr;
  }
};

typedef void (*GemmFn)(const double*, int64_t, const double*, int64_t, int, int, double, const double*, double,
                       const double*, double, double*, double*, unsigned int*, int);
int gemm_threads(int) { return 128; }
GemmFn gemm_for(int nj) {
  switch (nj) {
    case 1: return eig_dgemm_mma_kernel<1>;
    case 2: return eig_dgemm_mma_kernel<2>;
    case 3: return eig_dgemm_mma_kernel<3>;
    case 4: return eig_dgemm_mma_kernel<4>;
    case 5: return eig_dgemm_mma_kernel<5>;
    case 6: return eig_dgemm_mma_kernel<6>;
    case 7: return eig_dgemm_mma_kernel<7>;
    case 8: return eig_dgemm_mma_kernel<8>;
    case 9: return eig_dgemm_mma_kernel<9>;
    default: return eig_dgemm_mma_kernel<10>;
  }
}

size_t gemm_smem(int nj) { return (size_t)(2 * EG_TM * (EG_KT + 4) + 2 * EG_KT * (nj * 16 + 4)) * 8; }

// out = alpha * A B + gamma * P + delta * Q with A = [m, k]; k == 0: the element-wise epilogue alone
void gemm(const Ctx& c, const double* a, int64_t lda, int m, int k, const double* bm, double alpha, const double* pm,
          double gamma, const double* qm, double delta, double* out) {
  const Layout& L = c.L;
  int nsplit = 1, kps = std::max(k, 1);
  if (k > 512) nsplit = L.nsplit, kps = L.k_per_split;  // the products with the covariance
  const int tiles = (m + EG_TM - 1) / EG_TM;
  ++c.launches;
  gemm_for(L.nj)<<<dim3((unsigned)tiles, (unsigned)nsplit), gemm_threads(L.nj), gemm_smem(L.nj), c.st>>>(
      a, lda, bm, L.ldn, m, k, alpha, pm, gamma, qm, delta, out, c.p(L.off_part),
      reinterpret_cast<unsigned int*>(c.ws + L.off_counters), kps);
}

void gram(const Ctx& c, const double* x, const double* y, double* s) {
  const Layout& L = c.L;
  ++c.launches;
  eig_gram_kernel<<<dim3((unsigned)L.gram_tiles, (unsigned)L.gram_chunks), EG_THREADS, 0, c.st>>>(
      x, y, L.ldn, c.h, L.b, s, L.ldn, c.p(L.off_gpart), L.gram_rows);
  if (L.gram_chunks > 1) {
    ++c.launches;
    eig_gram_reduce_kernel<<<(L.gram_tiles * 4096 + 255) / 256, 256, 0, c.st>>>(c.p(L.off_gpart), L.gram_tiles, L.gram_chunks,
                                                                                 L.b, s, L.ldn);
  }
}

// q <- orthonormal basis of range(y) (two passes); `robust`: eigenvalue-based factor instead of Cholesky
int32_t orthonormalise(const Ctx& c, double* y, double* tmp, double** result, bool robust) {
  const Layout& L = c.L;
  double* src = y;
  double* dst = tmp;
  for (int pass = 0; pass < 2; ++pass) {
    gram(c, src, src, c.p(L.off_s));
    c.launches += robust ? 3 : 1;
    if (!robust) {
      const size_t smem = (size_t)L.b * ((L.b + 1) | 1) * 8;
      eig_chol_kernel<<<1, 1024, smem, c.st>>>(c.p(L.off_s), L.ldn, L.b, pass == 0 ? 1e-10 : 0.0, c.p(L.off_w), L.ldn,
                                                reinterpret_cast<int*>(c.ws + L.off_flags), c.p(L.off_minp));
    } else {
      eig_svqb_scale_kernel<<<32, 256, 0, c.st>>>(c.p(L.off_s), L.ldn, L.b, c.p(L.off_t), L.ldn, c.p(L.off_d));
      int32_t rc = jacobi_eig_launch(c.p(L.off_t), L.b, L.ldn, c.p(L.off_lam), c.p(L.off_u), L.ldn, nullptr, 0, nullptr,
                                     c.st);
      if (rc) return rc;
      eig_svqb_factor_kernel<<<32, 256, 0, c.st>>>(c.p(L.off_u), L.ldn, c.p(L.off_lam), c.p(L.off_d), L.b,
                                                    c.p(L.off_w), L.ldn);
    }
    gemm(c, src, L.ldn, c.h, L.b, c.p(L.off_w), 1.0, nullptr, 0.0, nullptr, 0.0, dst);
    std::swap(src, dst);
  }
  *result = src;  // after two passes: back in y's buffer
  return scf_check_launch("scf_eig_topk(orthonormalise)");
}

double cheb_growth(double x) { return x + sqrt(std::max(x * x - 1.0, 0.0)); }

// Scaled Chebyshev filter of the given degree applied to x0 (h x b): damps [0, cut], keeps the component at `top` near
// unit size.  When cx0 (= C x0) is given the first step takes it instead of a product with C.  Buffers: the three
// tall matrices bufs[0..2] rotate; returns the one that holds the result.
double* cheb_filter(const Ctx& c, const double* x0, const double* cx0, int degree, double cut, double top,
                    double* bufs[3]) {
  const Layout& L = c.L;
  const double* cov = c.p(L.off_cov);
  const double e = 0.5 * cut, t = (top - e) / e;
  const double r = 1.0 / (t + sqrt(std::max(t * t - 1.0, 0.0)));
  auto sig = [&](int j) { return r * (1.0 + pow(r, 2.0 * (j - 1.0))) / (1.0 + pow(r, 2.0 * j)); };  // sigma_j, j >= 1
  // y_1 = (sigma_1 / e) (C - e) x0
  const double s1 = sig(1) / e;
  double* y = bufs[0];
  if (cx0) {  // y = s1 cx0 - s1 e x0: the GEMM kernel's epilogue on an empty product
    gemm(c, cov, L.ldc, c.h, 0, x0, 0.0, cx0, s1, x0, -s1 * e, y);
  } else {
    gemm(c, cov, L.ldc, c.h, c.h, x0, s1, x0, -s1 * e, nullptr, 0.0, y);
  }
  const double* xp = x0;
  for (int j = 1; j < degree; ++j) {
    // y_{j+1} = a (C - e) y_j - d y_{j-1},  a = 2 sigma_{j+1} / e,  d = sigma_j sigma_{j+1}
    const double sj = sig(j), sj1 = sig(j + 1);
    const double a = 2.0 * sj1 / e, d = sj * sj1;
    double* yn = nullptr;
    for (int i = 0; i < 3; ++i)
      if (bufs[i] != y && bufs[i] != xp) yn = bufs[i];
    gemm(c, cov, L.ldc, c.h, c.h, y, a, y, -a * e, xp, -d, yn);
    xp = y;
    y = yn;
  }
  return y;
}

}  // namespace

extern "C" int64_t scf_eig_topk_workspace_bytes(int32_t h, int32_t dims) {
  Layout L;
  if (!make_layout(h, dims, L)) return -1;
  return (int64_t)L.total;
}

extern "C" int32_t scf_eig_topk(const int64_t* gram_fx, int64_t ldg, int32_t h, double scale, const double* col_mean,
                                double mean_weight, int32_t dims, double tol, int32_t max_rounds, double* evals,
                                double* evecs, float* evecs_f32, int64_t ld32, double* host_report, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  SCF_ARG(gram_fx && evals && evecs && host_report && workspace, "null pointer");
  SCF_ARG(h >= 1 && dims >= 1 && dims <= h && ldg >= h, "bad sizes");
  SCF_ARG(!evecs_f32 || ld32 >= dims, "ld32 < dims");
  Ctx c;
  if (!make_layout(h, dims, c.L)) {
    scf_set_error("scf_eig_topk: dims = %d does not fit the %d-column block of the solver", dims, EG_MAXB);
    return 1;
  }
  SCF_ARG(workspace_bytes >= (int64_t)c.L.total, "workspace too small");
  const Layout& L = c.L;
  c.h = h, c.dims = dims, c.ws = (unsigned char*)workspace, c.st = (cudaStream_t)stream;
  const int b = L.b;
  cudaError_t e = cudaFuncSetAttribute(eig_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)EG_MAXB * ((EG_MAXB + 1) | 1) * 8));
  if (e == cudaSuccess) e = cudaMemsetAsync(c.ws + L.off_counters, 0, (size_t)(L.row_tiles + L.gram_tiles + 8) * 4, c.st);
  if (e == cudaSuccess) e = cudaMemsetAsync(c.ws + L.off_flags, 0, 64, c.st);
  if (e != cudaSuccess) {
    scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  double* cov = c.p(L.off_cov);
  eig_cov_kernel<<<4 * SCF_NUM_SMS, 256, 0, c.st>>>((const long long*)gram_fx, ldg, h, scale, col_mean, mean_weight, cov,
                                                    L.ldc);
  eig_rowsum_kernel<<<(h + 7) / 8, 256, 0, c.st>>>(cov, L.ldc, h, c.p(L.off_rowsum));
  eig_norms_kernel<<<1, 1024, 0, c.st>>>(cov, L.ldc, h, c.p(L.off_rowsum), c.p(L.off_norms));
  int32_t rc = scf_check_launch("scf_eig_topk(cov)");
  if (rc) return rc;
  double norms[2];
  e = cudaMemcpyAsync(norms, c.p(L.off_norms), 16, cudaMemcpyDeviceToHost, c.st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
  if (e != cudaSuccess) {
    scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  const double trace = norms[0], top0 = norms[1];
  memset(host_report, 0, (size_t)(8 + EG_MAXB + 8) * 8);
  host_report[1] = trace, host_report[2] = top0;
  if (!(trace > 0.0) || !(top0 > 0.0) || !(trace == trace) || isinf(top0)) {
    // an all-zero (every selected feature constant) or non-finite covariance has no PCA
    scf_set_error("scf_eig_topk: the covariance is zero or not finite (trace %g, norm %g)", trace, top0);
    return 2;
  }
  for (int i = 1; i <= 10; ++i) {
    e = cudaFuncSetAttribute(gemm_for(i), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem(i));
    if (e != cudaSuccess) {
      scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
  }
  double* q = nullptr;
  bool robust = false;
  int rounds = 0, restarts = 0;
  double res = INFINITY;
  for (;;) {  // (re)start: the second attempt orthonormalises with the eigenvalue-based factor
    const double big = 1e300;
    e = cudaMemcpyAsync(c.p(L.off_minp), &big, 8, cudaMemcpyHostToDevice, c.st);
    if (e == cudaSuccess) e = cudaMemsetAsync(c.ws + L.off_flags, 0, 64, c.st);
    if (e != cudaSuccess) {
      scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
    q = c.tall(0);
    eig_init_kernel<<<2 * SCF_NUM_SMS, 256, 0, c.st>>>(q, h, b, L.ldn, 4466ull);
    // start: filter steps of degree 3 (cut at the mean eigenvalue: the wanted ones lie above it), each followed by an
    // orthonormalisation -- a degree-3 step keeps the block's condition number within the Cholesky-QR2 range
    for (int sidx = 0; sidx < 4; ++sidx) {
      double* fb[3];
      fb[0] = c.spare(q), fb[1] = c.spare(q, fb[0]), fb[2] = c.spare(q, fb[0], fb[1]);
      double* y = cheb_filter(c, q, nullptr, 3, trace / h, top0, fb);
      rc = orthonormalise(c, y, q, &q, robust);  // q's buffer is free once the filter has run
      if (rc) return rc;
    }
    bool restart = false;
    for (rounds = 1; rounds <= max_rounds; ++rounds) {
      double* aq = c.spare(q);
      double* v = c.spare(q, aq);
      double* av = c.spare(q, aq, v);
      gemm(c, cov, L.ldc, h, h, q, 1.0, nullptr, 0.0, nullptr, 0.0, aq);
      gram(c, q, aq, c.p(L.off_t));
      // Ritz pairs of T: tridiagonalisation + multisection + inverse iteration (tridiag_eig.cu); the Jacobi kernel is
      // launched behind it and returns at once unless that result failed its orthogonality check
      int* tri_ok = reinterpret_cast<int*>(c.ws + L.off_flags) + 4;
      if (b >= 3) {
        rc = tridiag_eig_launch(c.p(L.off_t), b, L.ldn, c.p(L.off_theta), c.p(L.off_u), L.ldn, 1, c.p(L.off_x), tri_ok,
                                c.st);
        if (rc) return rc;
        c.launches += 4;  // tridiagonalisation, eigenpairs of the tridiagonal matrix, back-transformation, check
      }
      rc = jacobi_eig_launch(c.p(L.off_t), b, L.ldn, c.p(L.off_theta), c.p(L.off_u), L.ldn, nullptr, 1,
                             b >= 3 ? tri_ok : nullptr, c.st);
      if (rc) return rc;
      c.launches += 3;  // Jacobi + the two residual kernels below
      // V = Q S, AV = AQ S (S = Ritz vectors, descending Ritz values)
      gemm(c, q, L.ldn, h, b, c.p(L.off_u), 1.0, nullptr, 0.0, nullptr, 0.0, v);
      gemm(c, aq, L.ldn, h, b, c.p(L.off_u), 1.0, nullptr, 0.0, nullptr, 0.0, av);
      eig_resid_partial_kernel<<<L.resid_chunks, 256, 0, c.st>>>(v, av, L.ldn, h, dims, c.p(L.off_theta),
                                                                 c.p(L.off_rpart), L.resid_rows);
      eig_resid_final_kernel<<<1, 256, 0, c.st>>>(c.p(L.off_rpart), L.resid_chunks, dims, b, c.p(L.off_theta),
                                                  c.p(L.off_norms), reinterpret_cast<int*>(c.ws + L.off_flags),
                                                  c.p(L.off_minp), c.p(L.off_report));
      rc = scf_check_launch("scf_eig_topk(rayleigh-ritz)");
      if (rc) return rc;
      e = cudaMemcpyAsync(host_report, c.p(L.off_report), (size_t)(8 + EG_MAXB + 8) * 8, cudaMemcpyDeviceToHost, c.st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);  // the one synchronisation of the round
      if (e != cudaSuccess) {
        scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
        return -(int32_t)e;
      }
      res = host_report[0];
      const bool chol_bad = host_report[3] != 0.0;
      if ((chol_bad || !(res == res) || isinf(res)) && !robust) {
        restart = true;  // a Cholesky factorisation broke down: once more with the eigenvalue-based factor
        break;
      }
      const double* th = host_report + 8;
      if (res <= tol) {
        e = cudaMemcpyAsync(evals, c.p(L.off_theta), (size_t)dims * 8, cudaMemcpyDeviceToDevice, c.st);
        if (e != cudaSuccess) {
          scf_set_error("scf_eig_topk: %s", cudaGetErrorString(e));
          return -(int32_t)e;
        }
        eig_sign_kernel<<<(int)((std::max<int64_t>(dims, evecs_f32 ? ld32 : 0) + 7) / 8), 256, 0, c.st>>>(
            v, L.ldn, h, dims, evecs, evecs_f32, ld32);
        host_report[5] = (double)rounds, host_report[6] = (double)restarts, host_report[7] = robust ? 1.0 : 0.0;
        host_report[168] = (double)(c.launches + 5);  // + covariance, row sums, norms, start block, sign rule
        return scf_check_launch("scf_eig_topk(sign)");
      }
      if (rounds == max_rounds) break;
      const double th_max = th[0], th_dims = th[dims - 1], th_min = th[b - 1];
      if (!(th_dims > 0.0 && th_max >= th_dims)) break;
      double sum = 0.0;
      for (int i = 0; i < b; ++i) sum += th[i];
      const double bulk = h > b ? (trace - sum) / (h - b) : 0.0;
      // damped interval [0, cut]: the block's smallest Ritz value, or the mean of the spectrum outside the block when
      // that is larger (it never exceeds lambda_{b+1}); kept clear of the wanted Ritz values
      double cut = std::max(th_min, std::min(bulk, 0.5 * (th_min + th_dims)));
      cut = std::min(std::max(cut, 1e-3 * th_dims), 0.9 * th_dims);
      const double eh = 0.5 * cut;
      const double rho = cheb_growth((th_max - eh) / eh) / cheb_growth((th_dims - eh) / eh);
      int m = (int)floor(log(1e20) / log(std::max(rho, 1.0 + 1e-9)));
      m = std::max(2, std::min(32, m));
      // the filter rotates through the three tall buffers that are not its inputs (q and aq are free now)
      double* fb[3];
      fb[0] = c.spare(v, av), fb[1] = c.spare(v, av, fb[0]), fb[2] = c.spare(v, av, fb[0], fb[1]);
      double* y = cheb_filter(c, v, av, m, cut, th_max, fb);
      rc = orthonormalise(c, y, c.spare(v, av, y), &q, robust);
      if (rc) return rc;
    }
    if (restart && restarts == 0) {
      ++restarts, robust = true;
      continue;
    }
    break;
  }
  host_report[5] = (double)-rounds, host_report[6] = (double)restarts, host_report[7] = robust ? 1.0 : 0.0;
  scf_set_error("scf_eig_topk: no convergence in %d rounds (residual %g, tolerance %g)", max_rounds, res, tol);
  return 3;
}
