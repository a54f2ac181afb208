// Small symmetric eigensolver for the later Rayleigh-Ritz steps of the PCA eigensolve (graph.eig_topk): all eigenpairs
// of a symmetric positive semi-definite n x n matrix in ONE CTA with the matrix resident in shared memory.
//
// cuSOLVER's syevd spends ~1.5 ms on an 82 x 82 matrix (a few hundred dependent tiny kernels: tridiagonalisation, 81
// reflector applications for the back-transformation).  The Ritz matrices of the second and later rounds are small
// (dims + 32 columns), well conditioned and nearly diagonal (the basis is a filtered set of Ritz vectors), which is
// the case one-sided (Hestenes) Jacobi is made for: the columns of W = T are rotated pairwise until they are mutually
// orthogonal, then W = V * Lambda, i.e. the column norms are the eigenvalues and the normalised columns the
// eigenvectors.  n/2 disjoint column pairs are rotated at once (round-robin ordering, n - 1 steps per sweep, eight
// threads per pair); convergence is quadratic once the columns are nearly orthogonal (2-3 sweeps here).  A sweep moves
// the whole matrix through shared memory n - 1 times (3 n^3 * 8 bytes), so the kernel pays off for n <~ 100 only:
// scarf_b200.ops.sym_eig_small keeps the library eigh for the wide first-round block.  Deterministic: every rank
// computes bit-identical pairs from bit-identical input.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int JE_THREADS = 1024;
// threads per column pair: 16 while n / 2 pairs x 16 fit one CTA (n <= 128), else 8
constexpr int JE_MAXN = 168;   // 168 * 168 * 8 B = 226 KB of shared memory
constexpr int JE_MAX_SWEEPS = 40;

// sum over the TPP lanes of a pair; only the lanes of that group take part (a warp can hold idle groups)
template <int TPP>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = TPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

template <int JE_TPP>
__global__ void __launch_bounds__(JE_TPP == 16 ? JE_THREADS : 704, 1) jacobi_eig_kernel(const double* __restrict__ a, int n, int64_t lda,
                                                                   double* __restrict__ evals,
                                                                   double* __restrict__ evecs, int64_t ldv,
                                                                   int* __restrict__ info, int descending,
                                                                   const int* __restrict__ skip_if) {
  if (skip_if && *skip_if == 1) return;  // the tridiagonal kernel (tridiag_eig.cu) has delivered: nothing to do
  extern __shared__ __align__(16) double w[];  // column major, n rows x ne columns (ne = n rounded up to even)
  __shared__ int s_rotated;
  __shared__ double s_norm[JE_MAXN + 1];
  __shared__ int s_rank[JE_MAXN + 1];
  const int tid = threadIdx.x;
  const int ne = n + (n & 1);
  const int nthreads = blockDim.x;
  for (int e = tid; e < n * ne; e += nthreads) {
    const int col = e / n, row = e - col * n;
    // symmetrised read: the caller's matrix is symmetric up to rounding
    w[e] = col < n ? 0.5 * (a[(int64_t)row * lda + col] + a[(int64_t)col * lda + row]) : 0.0;
  }
  if (tid == 0) s_rotated = 0;
  __syncthreads();
  const int pair = tid / JE_TPP, sub = tid % JE_TPP;
  const int npairs = ne / 2;
  const bool active = pair < npairs;
  const unsigned gmask = (JE_TPP == 16 ? 0xFFFFu : 0xFFu) << ((tid & 31) & ~(JE_TPP - 1));
  constexpr int RMAX = JE_TPP == 16 ? 8 : 21;  // rows of a column one thread holds (n <= 128 resp. 168)
  int sweeps = 0;
  for (; sweeps < JE_MAX_SWEEPS; ++sweeps) {
    for (int step = 0; step < ne - 1; ++step) {
      if (active) {
        // round-robin tournament: column ne-1 stays, the others rotate
        int ci, cj;
        if (pair == 0) {
          ci = ne - 1, cj = step;
        } else {
          ci = (step + pair) % (ne - 1);
          cj = (step - pair + (ne - 1)) % (ne - 1);
        }
        double* wi = w + (size_t)ci * n;
        double* wj = w + (size_t)cj * n;
        // the thread's rows of both columns stay in registers between the dot products and the rotation
        double xr[RMAX], yr[RMAX];
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
        for (int t = 0; t < RMAX; ++t) {
          const int r = sub + t * JE_TPP;
          xr[t] = r < n ? wi[r] : 0.0;
          yr[t] = r < n ? wj[r] : 0.0;
          alpha = fma(xr[t], xr[t], alpha);
          beta = fma(yr[t], yr[t], beta);
          gamma = fma(xr[t], yr[t], gamma);
        }
        alpha = group_sum<JE_TPP>(alpha, gmask), beta = group_sum<JE_TPP>(beta, gmask);
        gamma = group_sum<JE_TPP>(gamma, gmask);
        // (thresholds compared in squared form: FP64 square roots and divisions are ~40-instruction sequences and sit
        // on the critical path of every one of the n - 1 dependent steps of a sweep)
        const double ab = alpha * beta, g2 = gamma * gamma;
        if (g2 > 1e-30 * ab && alpha > 0.0 && beta > 0.0) {  // uniform inside the group
          // tan of the rotation angle: t = 2 gamma / (d + sign(d) sqrt(d^2 + 4 gamma^2)), d = beta - alpha (the smaller
          // root: |t| <= 1); sqrt(q) = q rsqrt(q), the division as a reciprocal
          const double d = beta - alpha;
          const double q = fma(d, d, 4.0 * g2);
          const double t = 2.0 * gamma * __drcp_rn(d + copysign(q * rsqrt(q), d));
          const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
#pragma unroll
          for (int u = 0; u < RMAX; ++u) {
            const int r = sub + u * JE_TPP;
            if (r < n) {
              wi[r] = c * xr[u] - sn * yr[u];
              wj[r] = sn * xr[u] + c * yr[u];
            }
          }
          if (sub == 0 && g2 > 1e-26 * ab) s_rotated = 1;  // benign race: any writer wins
        }
      }
      __syncthreads();
    }
    const int again = s_rotated;
    __syncthreads();
    if (tid == 0) s_rotated = 0;
    __syncthreads();
    if (!again) {
      ++sweeps;
      break;
    }
  }
  // eigenvalues = column norms; ascending order by rank (ties by column index)
  for (int c = tid; c < n; c += nthreads) {
    double s2 = 0.0;
    for (int r = 0; r < n; ++r) s2 = fma(w[(size_t)c * n + r], w[(size_t)c * n + r], s2);
    s_norm[c] = sqrt(s2);
  }
  __syncthreads();
  for (int c = tid; c < n; c += nthreads) {
    int rank = 0;
    const double mine = s_norm[c];
    for (int o = 0; o < n; ++o) rank += (s_norm[o] < mine) || (s_norm[o] == mine && o < c);
    if (descending) rank = n - 1 - rank;
    s_rank[c] = rank;
    evals[rank] = mine;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += nthreads) {
    const int c = e / n, r = e - c * n;
    const double nv = s_norm[c];
    // a zero column can only come from an exactly singular matrix: its direction is undefined, emit a unit vector
    evecs[(int64_t)r * ldv + s_rank[c]] = nv > 0.0 ? w[e] / nv : (r == c ? 1.0 : 0.0);
  }
  if (tid == 0 && info) *info = sweeps < JE_MAX_SWEEPS ? sweeps : -1;
}

}  // namespace

int32_t jacobi_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int* info,
                          int descending, const int* skip_if, cudaStream_t stream);

extern "C" int32_t scf_sym_eig_max_n(void) { return JE_MAXN; }

extern "C" int32_t scf_sym_eig_jacobi(const double* a, int32_t n, int64_t lda, double* evals, double* evecs, int64_t ldv,
                                      int32_t* info, void* stream) {
  return jacobi_eig_launch(a, n, lda, evals, evecs, ldv, info, 0, nullptr, (cudaStream_t)stream);
}

// descending != 0: eigenvalues (and the matching eigenvector columns) in descending order (eig_topk.cu)
int32_t jacobi_eig_launch(const double* a, int n, int64_t lda, double* evals, double* evecs, int64_t ldv, int* info,
                          int descending, const int* skip_if, cudaStream_t stream) {
  SCF_ARG(a && evals && evecs, "null pointer");
  SCF_ARG(n >= 1 && n <= JE_MAXN && lda >= n && ldv >= n, "n must be within [1, 168]");
  const size_t smem = (size_t)n * (n + (n & 1)) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(jacobi_eig_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(jacobi_eig_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_sym_eig_jacobi: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  // one group of 16 (n <= 128) or 8 threads per column pair: a small matrix runs with few warps (cheaper barriers)
  const int npairs = (n + 1) / 2;
  const int tpp = npairs * 16 <= JE_THREADS ? 16 : 8;
  int threads = (npairs * tpp + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > JE_THREADS ? JE_THREADS : threads);
  if (tpp == 16)
    jacobi_eig_kernel<16><<<1, threads, smem, stream>>>(a, n, lda, evals, evecs, ldv, info, descending, skip_if);
  else
    jacobi_eig_kernel<8><<<1, threads, smem, stream>>>(a, n, lda, evals, evecs, ldv, info, descending, skip_if);
  return scf_check_launch("scf_sym_eig_jacobi");
}
