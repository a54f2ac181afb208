// K6: umap-learn smooth_knn_dist + compute_membership_strengths + COO assembly, as Scarf calls them
// per chunk (scarf/knn_utils.py:89-159).  All arithmetic is float32 like the numba originals.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr float kSmoothTol = 1e-5f;
constexpr float kMinKDistScale = 1e-3f;
constexpr int kMaxK = 256;

// one CTA per chunk: float64 sum of the chunk's distances held by this shard (a chunk that straddles two
// shards is completed by a sum all-reduce; the caller divides by rows_in_chunk * k and rounds to float32)
__global__ void __launch_bounds__(256) chunk_sums_kernel(const float* __restrict__ dist, int64_t n, int k,
                                                         int64_t row_offset, int64_t chunk_size,
                                                         double* __restrict__ chunk_sum) {
  const int64_t first_chunk = row_offset / chunk_size;
  const int64_t c = first_chunk + blockIdx.x;
  int64_t g0 = c * chunk_size, g1 = g0 + chunk_size;
  int64_t r0 = max(g0 - row_offset, (int64_t)0), r1 = min(g1 - row_offset, n);
  double acc = 0.0;
  for (int64_t e = r0 * k + threadIdx.x; e < r1 * k; e += 256) acc += (double)dist[e];
  __shared__ double sh[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    chunk_sum[c] = t;
  }
}

__global__ void __launch_bounds__(128) smooth_knn_kernel(const float* __restrict__ dist, int64_t n, int k,
                                                         float lc, float bandwidth, int64_t row_offset,
                                                         int64_t chunk_size, const float* __restrict__ chunk_mean,
                                                         float* __restrict__ sigma, float* __restrict__ rho_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* row = dist + r * k;
  const float target = log2f((float)k) * bandwidth;
  // rho: distance to the local_connectivity-th nearest neighbour with a positive distance
  int nnz = 0;
  float mx = 0.f, sum = 0.f;
  for (int j = 0; j < k; ++j) {
    const float d = row[j];
    sum += d;
    if (d > 0.f) {
      ++nnz;
      mx = fmaxf(mx, d);
    }
  }
  auto nz_at = [&](int want) {  // want-th (0-based) positive entry in row order
    int seen = 0;
    for (int j = 0; j < k; ++j)
      if (row[j] > 0.f) {
        if (seen == want) return row[j];
        ++seen;
      }
    return 0.f;
  };
  float rho = 0.f;
  if ((float)nnz >= lc) {
    const int index = (int)floorf(lc);
    const float interp = lc - (float)index;
    if (index > 0) {
      rho = nz_at(index - 1);
      if (interp > kSmoothTol && index < nnz) rho += interp * (nz_at(index) - rho);
    } else {
      rho = interp * nz_at(0);
    }
  } else if (nnz > 0) {
    rho = mx;
  }
  float lo = 0.f, hi = CUDART_INF_F, mid = 1.f;
  for (int it = 0; it < 64; ++it) {
    float psum = 0.f;
    for (int j = 1; j < k; ++j) {  // j starts at 1: SURVEY fact 5
      const float d = row[j] - rho;
      psum += d > 0.f ? expf(-(d / mid)) : 1.f;
    }
    if (fabsf(psum - target) < kSmoothTol) break;
    if (psum > target) {
      hi = mid;
      mid = (lo + hi) / 2.f;
    } else {
      lo = mid;
      mid = hi == CUDART_INF_F ? mid * 2.f : (lo + hi) / 2.f;
    }
  }
  float res = mid;
  if (rho > 0.f) {
    const float mean_i = sum / (float)k;
    res = fmaxf(res, kMinKDistScale * mean_i);
  } else {
    res = fmaxf(res, kMinKDistScale * chunk_mean[(row_offset + r) / chunk_size]);
  }
  sigma[r] = res;
  rho_out[r] = rho;
}

__device__ __forceinline__ void atomic_min_pos(float* addr, float v) {  // v > 0
  atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__global__ void __launch_bounds__(256) membership_kernel(const int64_t* __restrict__ idx,
                                                         const float* __restrict__ dist,
                                                         const float* __restrict__ sigma,
                                                         const float* __restrict__ rho, int64_t n, int k,
                                                         int64_t row_offset, int64_t chunk_size,
                                                         int64_t* __restrict__ edges, float* __restrict__ weights,
                                                         float* __restrict__ chunk_min,
                                                         int32_t* __restrict__ chunk_has_zero) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * k) return;
  const int64_t r = e / k;
  const int64_t g = row_offset + r;
  const int64_t local = g % chunk_size;
  const int64_t nb = idx[e];
  const float sg = sigma[r];
  const float d = dist[e] - rho[r];
  float w;
  if (nb == local)
    w = 0.f;
  else if (d <= 0.f || sg == 0.f)
    w = 1.f;
  else
    w = expf(-(d / sg));
  edges[2 * e] = g;
  edges[2 * e + 1] = nb;
  weights[e] = w;
  // per-chunk minimum of the non-zero weights / "has a zero" flag.  A full warp whose 32 consecutive edges lie in
  // one chunk (all but the warps that straddle a chunk boundary or the tail) reduces first: one atomic per warp
  // instead of one per edge on the same handful of addresses.
  const int64_t c = g / chunk_size;
  const unsigned active = __activemask();
  if (active == SCF_FULL && __match_any_sync(SCF_FULL, c) == SCF_FULL) {
    float wmin = w > 0.f ? w : CUDART_INF_F;
    int zero = w == 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wmin = fminf(wmin, __shfl_xor_sync(SCF_FULL, wmin, o));
      zero |= __shfl_xor_sync(SCF_FULL, zero, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (zero) chunk_has_zero[c] = 1;
      if (wmin < CUDART_INF_F) atomic_min_pos(chunk_min + c, wmin);
    }
  } else if (w == 0.f) {
    chunk_has_zero[c] = 1;
  } else {
    atomic_min_pos(chunk_min + c, w);
  }
}

__global__ void fill_zero_kernel(float* __restrict__ w, int64_t n, float v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n && w[e] == 0.f) w[e] = v;
}

// load_graph's symmetrisation g + g^T - g o g^T (scarf/datastore/graph_datastore.py:1052-1075) on the stored kNN graph:
// one thread per directed edge (i -> j, weight w) of the first use_k neighbours of every row; the reverse edge's
// weight w' is looked up in row j (0 when j does not list i).  s = (w + w') - w w' in float64, the two steps rounded
// separately like the sparse-matrix expression.  Slot 2e holds (i, j, s); slot 2e + 1 holds the mirrored entry
// (j, i, s) when the reverse edge is absent (otherwise that entry is produced by the reverse edge's own thread) and
// row -1 when it is not needed.  upper_only: only the entries with row <= column survive (scipy.sparse.triu).
__global__ void graph_symmetrize_kernel(const int64_t* __restrict__ idx, const double* __restrict__ w, int64_t n, int k,
                                        int use_k, int upper_only, int64_t* __restrict__ out_row,
                                        int64_t* __restrict__ out_col, double* __restrict__ out_val) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * use_k) return;
  const int64_t i = e / use_k;
  const int slot = (int)(e - i * use_k);
  const int64_t j = idx[i * k + slot];
  const double wij = w[i * k + slot];
  double wji = 0.0;
  bool reverse = false;
  if (j >= 0 && j < n)
    for (int t = 0; t < use_k; ++t)
      if (idx[j * k + t] == i) {
        wji = __dadd_rn(wji, w[j * k + t]);  // duplicates add up, as in the COO -> CSR conversion
        reverse = true;
      }
  const double sv = __dsub_rn(__dadd_rn(wij, wji), __dmul_rn(wij, wji));
  int64_t r0 = i, c0 = j, r1 = -1, c1 = -1;
  if (!reverse) r1 = j, c1 = i;
  if (upper_only) {
    if (r0 > c0) r0 = -1;
    if (r1 > c1) r1 = -1;
  }
  out_row[2 * e] = r0, out_col[2 * e] = c0, out_val[2 * e] = sv;
  out_row[2 * e + 1] = r1, out_col[2 * e + 1] = c1, out_val[2 * e + 1] = sv;
}

}  // namespace

extern "C" int32_t scf_graph_symmetrize(const int64_t* idx, const double* weights, int64_t n, int32_t k, int32_t use_k,
                                        int32_t upper_only, int64_t* out_row, int64_t* out_col, double* out_val,
                                        void* stream) {
  SCF_ARG(idx && weights && out_row && out_col && out_val, "null pointer");
  SCF_ARG(n >= 0 && k > 0 && use_k >= 1 && use_k <= k, "bad sizes");
  if (n == 0) return 0;
  graph_symmetrize_kernel<<<(unsigned)((n * use_k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      idx, weights, n, k, use_k, upper_only, out_row, out_col, out_val);
  return scf_check_launch("scf_graph_symmetrize");
}

extern "C" int32_t scf_chunk_sums(const float* dist, int64_t n, int32_t k, int64_t row_offset, int64_t chunk_size,
                                  double* chunk_sum, void* stream) {
  SCF_ARG(dist && chunk_sum, "null pointer");
  SCF_ARG(n >= 0 && k > 0 && row_offset >= 0 && chunk_size > 0, "bad sizes");
  if (n == 0) return 0;
  const int64_t c0 = row_offset / chunk_size, c1 = (row_offset + n - 1) / chunk_size;
  chunk_sums_kernel<<<(unsigned)(c1 - c0 + 1), 256, 0, (cudaStream_t)stream>>>(dist, n, k, row_offset, chunk_size,
                                                                                chunk_sum);
  return scf_check_launch("scf_chunk_sums");
}

extern "C" int32_t scf_smooth_knn(const float* dist, int64_t n, int32_t k, float local_connectivity, float bandwidth,
                                  int64_t row_offset, int64_t chunk_size, const float* chunk_mean, float* sigma,
                                  float* rho, void* stream) {
  SCF_ARG(dist && chunk_mean && sigma && rho, "null pointer");
  SCF_ARG(n >= 0 && k > 0 && k <= kMaxK && row_offset >= 0 && chunk_size > 0, "bad sizes");
  SCF_ARG(local_connectivity >= 0.f, "local_connectivity < 0");
  if (n == 0) return 0;
  smooth_knn_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      dist, n, k, local_connectivity, bandwidth, row_offset, chunk_size, chunk_mean, sigma, rho);
  return scf_check_launch("scf_smooth_knn");
}

extern "C" int32_t scf_membership_coo(const int64_t* idx, const float* dist, const float* sigma, const float* rho,
                                      int64_t n, int32_t k, int64_t row_offset, int64_t chunk_size, int64_t* edges,
                                      float* weights, float* chunk_min, int32_t* chunk_has_zero, void* stream) {
  SCF_ARG(idx && dist && sigma && rho && edges && weights && chunk_min && chunk_has_zero, "null pointer");
  SCF_ARG(n >= 0 && k > 0 && row_offset >= 0 && chunk_size > 0, "bad sizes");
  if (n == 0) return 0;
  membership_kernel<<<(unsigned)((n * k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      idx, dist, sigma, rho, n, k, row_offset, chunk_size, edges, weights, chunk_min, chunk_has_zero);
  return scf_check_launch("scf_membership_coo");
}

extern "C" int32_t scf_fill_zero_weights(float* weights, int64_t n, float floor_value, void* stream) {
  SCF_ARG(weights, "null pointer");
  if (n <= 0) return 0;
  fill_zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weights, n, floor_value);
  return scf_check_launch("scf_fill_zero_weights");
}
