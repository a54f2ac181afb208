// K0 / K1: CSR row kernels (warp per cell, persistent grid).  HBM-bound: the algorithmic traffic
// is 8 B per stored value (int32 gene id + uint32 count) + 8 B of indptr per cell.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

__device__ __forceinline__ int64_t row_of(const int64_t* row_ids, int64_t r) { return row_ids ? row_ids[r] : r; }

// Calls f(gene, count) for every stored value of one CSR row, 32 lanes x 4 loads in flight.
template <typename F>
__device__ __forceinline__ void warp_row_scan(const int32_t* __restrict__ indices, const uint32_t* __restrict__ data,
                                              int64_t s, int64_t e, int lane, F f) {
  for (int64_t p = s + lane; p < e; p += 128) {
    int32_t g[4];
    uint32_t c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t pp = p + 32 * u;
      const bool ok = pp < e;
      g[u] = ok ? ld_stream(indices + pp) : -1;
      c[u] = ok ? ld_stream(data + pp) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (g[u] >= 0) f(g[u], c[u]);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) row_sums_kernel(const int64_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices,
                                                            const uint32_t* __restrict__ data,
                                                            const int64_t* __restrict__ row_ids, int64_t n_sel,
                                                            const int32_t* __restrict__ col_map,
                                                            double* __restrict__ out_sum, int32_t* __restrict__ out_nnz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    unsigned long long acc = 0;
    int cnt = 0;
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      const bool sel = col_map ? (__ldg(col_map + g) >= 0) : true;
      if (sel) {
        acc += c;
        cnt += c > 0;
      }
    });
    acc = warp_sum(acc);
    cnt = warp_sum(cnt);
    if (lane == 0) {
      out_sum[r] = (double)acc;
      if (out_nnz) out_nnz[r] = cnt;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gene_stats_kernel(const int64_t* __restrict__ indptr,
                                                              const int32_t* __restrict__ indices,
                                                              const uint32_t* __restrict__ data,
                                                              const int64_t* __restrict__ row_ids, int64_t n_sel,
                                                              const double* __restrict__ row_div, double sf,
                                                              unsigned long long* __restrict__ gene_nnz,
                                                              double* __restrict__ gene_sum,
                                                              double* __restrict__ gene_sumsq) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    const double dv = row_div ? row_div[r] : 1.0;
    const bool norm = row_div != nullptr;
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      if (c == 0) return;
      // norm_lib_size: sf * counts / scalar  (scarf/assay.py:51), evaluated left to right in float64
      const double v = norm ? __ddiv_rn(sf * (double)c, dv) : (double)c;
      atomicAdd(gene_nnz + g, 1ull);
      if (gene_sum) atomicAdd(gene_sum + g, v);
      if (gene_sumsq) atomicAdd(gene_sumsq + g, v * v);
    });
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm_value(uint32_t c, double s, double sf, bool log_transform) {
  const double v = __ddiv_rn(sf * (double)c, s);
  return log_transform ? log1p(v) : v;
}

__global__ void __launch_bounds__(kThreads) hvg_colstats_kernel(const int64_t* __restrict__ indptr,
                                                                const int32_t* __restrict__ indices,
                                                                const uint32_t* __restrict__ data,
                                                                const int64_t* __restrict__ row_ids, int64_t n_sel,
                                                                const int32_t* __restrict__ col_map,
                                                                const double* __restrict__ row_sum, double sf,
                                                                int log_transform, long long* __restrict__ sum_fx,
                                                                long long* __restrict__ sumsq_fx, int n_cols,
                                                                int n_rep) {
  // same-address atomics serialise in L2: every CTA adds into one of n_rep copies of the H-long accumulators
  sum_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  sumsq_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    double sc = row_sum[r];
    if (sc == 0.0) sc = 1.0;  // scalar[scalar == 0] = 1  (scarf/assay.py:821-823)
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      const int col = __ldg(col_map + g);
      if (col < 0 || c == 0) return;
      const double x = norm_value(c, sc, sf, log_transform != 0);
      atomicAdd((unsigned long long*)(sum_fx + col), (unsigned long long)to_fx(x, SCF_COLSTAT_SHIFT));
      atomicAdd((unsigned long long*)(sumsq_fx + col), (unsigned long long)to_fx(x * x, SCF_COLSTAT_SHIFT));
    });
  }
}

// ---------------------------------------------------------------------------------------------
// Z rows go straight to global memory: the warp first writes the "no count" row ((0 - mu)/sigma, identical for
// every cell, kept in shared memory), then overwrites the <= n_hvg entries the cell really has.  Both are plain
// stores that merge in L2, so HBM sees each Z row once; no per-warp row buffer means ~40 resident warps per SM to
// cover the latency of the CSR stream.  smem: double mu[ldz] | double sigma[ldz] | float base[ldz] | float base_lo[ldz]
__device__ __forceinline__ float tf32_low_part(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kThreads) norm_scale_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const uint32_t* __restrict__ data,
    const int64_t* __restrict__ row_ids, int64_t n_sel, const int32_t* __restrict__ col_map, int n_cols,
    const double* __restrict__ row_sum, double sf, int log_transform, const double* __restrict__ mu,
    const double* __restrict__ sigma, const double* __restrict__ missing_fill, float* __restrict__ z,
    float* __restrict__ z_lo, int64_t ldz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_mu = reinterpret_cast<double*>(smem_raw);
  double* s_sigma = s_mu + ldz;
  float* s_base = reinterpret_cast<float*>(s_sigma + ldz);
  float* s_base_lo = s_base + ldz;
  for (int j = threadIdx.x; j < ldz; j += kThreads) {
    const double m = (j < n_cols && mu) ? mu[j] : 0.0;
    const double sd = (j < n_cols && sigma) ? sigma[j] : 1.0;
    double x0 = 0.0;
    if (j < n_cols && missing_fill) {
      const double f = missing_fill[j];
      if (f == f) x0 = f;
    }
    s_mu[j] = m;
    s_sigma[j] = sd;
    const float b = j < n_cols ? (float)__ddiv_rn(x0 - m, sd) : 0.f;
    s_base[j] = b;
    s_base_lo[j] = tf32_low_part(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int nvec = (int)(ldz >> 2);
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    double sc = row_sum[r];
    if (sc == 0.0) sc = 1.0;
    float* zr = z + r * ldz;
    float* zl = z_lo ? z_lo + r * ldz : nullptr;
    for (int j = lane; j < nvec; j += 32) {
      reinterpret_cast<float4*>(zr)[j] = reinterpret_cast<const float4*>(s_base)[j];
      if (zl) reinterpret_cast<float4*>(zl)[j] = reinterpret_cast<const float4*>(s_base_lo)[j];
    }
    __syncwarp();  // orders the base row before the overwrites below (different lanes hit the same addresses)
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      const int col = __ldg(col_map + g);
      if (col < 0) return;
      if (missing_fill) {  // a column the target matrix does not really have keeps its fill value
        const double f = __ldg(missing_fill + col);
        if (f == f) return;
      }
      const double x = norm_value(c, sc, sf, log_transform != 0);
      const float v = (float)__ddiv_rn(x - s_mu[col], s_sigma[col]);
      zr[col] = v;
      if (zl) zl[col] = tf32_low_part(v);  // what the tensor core's 19-bit read of z drops (3xTF32 low plane)
    });
  }
}

// ---------------------------------------------------------------------------------------------
// K1a': one pass that (i) compacts the selected, non-zero entries of every row into a small CSR of normalised
// values x = log1p(sf*c/row_sum) (col int32, x float64, row offsets given) and (ii) accumulates the fixed-point
// column sums of x and x^2.  The selected entries are ~10 % of a row and scattered over the lanes, so doing the
// FP64 division + log1p where they are found would run it at 1/8 lane occupancy; here each warp first packs them
// (ballot + popc) into a shared-memory batch and then evaluates full batches.  K1b' reads this compact matrix.
constexpr int kBatch = 512;  // entries per warp batch (flushed when fewer than 128 slots remain)

__global__ void __launch_bounds__(kThreads) hvg_compact_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const uint32_t* __restrict__ data,
    const int64_t* __restrict__ row_ids, int64_t n_sel, const int32_t* __restrict__ col_map,
    const double* __restrict__ row_sum, double sf, int log_transform, const int64_t* __restrict__ row_off,
    int32_t* __restrict__ out_col, double* __restrict__ out_x, long long* __restrict__ sum_fx,
    long long* __restrict__ sumsq_fx, int n_cols, int n_rep) {
  __shared__ int32_t s_col[kWarpsPerCta][kBatch];
  __shared__ uint32_t s_cnt[kWarpsPerCta][kBatch];
  sum_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  sumsq_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int32_t* bc = s_col[warp];
  uint32_t* bn = s_cnt[warp];
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    double sc = row_sum[r];
    if (sc == 0.0) sc = 1.0;  // scalar[scalar == 0] = 1  (scarf/assay.py:821-823)
    int64_t cursor = row_off[r];
    int pending = 0;
    auto flush = [&]() {
      __syncwarp();
      for (int i = lane; i < pending; i += 32) {
        const int col = bc[i];
        const double x = norm_value(bn[i], sc, sf, log_transform != 0);
        atomicAdd((unsigned long long*)(sum_fx + col), (unsigned long long)to_fx(x, SCF_COLSTAT_SHIFT));
        atomicAdd((unsigned long long*)(sumsq_fx + col), (unsigned long long)to_fx(x * x, SCF_COLSTAT_SHIFT));
        out_col[cursor + i] = col;
        out_x[cursor + i] = x;
      }
      cursor += pending;
      pending = 0;
      __syncwarp();
    };
    for (int64_t p = s + lane; p - lane < e; p += 128) {
      int32_t g[4];
      uint32_t c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t pp = p + 32 * u;
        const bool ok = pp < e;
        g[u] = ok ? ld_stream(indices + pp) : -1;
        c[u] = ok ? ld_stream(data + pp) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = g[u] >= 0 ? __ldg(col_map + g[u]) : -1;
        const bool hit = col >= 0 && c[u] != 0u;
        const unsigned m = __ballot_sync(SCF_FULL, hit);
        if (hit) {
          const int pos = pending + __popc(m & lt_mask);
          bc[pos] = col;
          bn[pos] = c[u];
        }
        pending += __popc(m);
      }
      if (pending > kBatch - 128) flush();
    }
    flush();
  }
}

// K1b': Z (and the 3xTF32 low plane) from the compact matrix: base row from shared memory, then the row's entries.
__global__ void __launch_bounds__(kThreads) hvg_dense_scale_kernel(
    const int64_t* __restrict__ row_off, const int32_t* __restrict__ cols, const double* __restrict__ xs,
    int64_t n_sel, int n_cols, const double* __restrict__ mu, const double* __restrict__ sigma,
    float* __restrict__ z, float* __restrict__ z_lo, int64_t ldz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_mu = reinterpret_cast<double*>(smem_raw);
  double* s_sigma = s_mu + ldz;
  float* s_base = reinterpret_cast<float*>(s_sigma + ldz);
  float* s_base_lo = s_base + ldz;
  for (int j = threadIdx.x; j < ldz; j += kThreads) {
    const double m = (j < n_cols && mu) ? mu[j] : 0.0;
    const double sd = (j < n_cols && sigma) ? sigma[j] : 1.0;
    s_mu[j] = m;
    s_sigma[j] = sd;
    const float b = j < n_cols ? (float)__ddiv_rn(0.0 - m, sd) : 0.f;
    s_base[j] = b;
    s_base_lo[j] = tf32_low_part(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int nvec = (int)(ldz >> 2);
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    float* zr = z + r * ldz;
    float* zl = z_lo ? z_lo + r * ldz : nullptr;
    for (int j = lane; j < nvec; j += 32) {
      reinterpret_cast<float4*>(zr)[j] = reinterpret_cast<const float4*>(s_base)[j];
      if (zl) reinterpret_cast<float4*>(zl)[j] = reinterpret_cast<const float4*>(s_base_lo)[j];
    }
    __syncwarp();
    const int64_t e = row_off[r + 1];
    for (int64_t p = row_off[r] + lane; p < e; p += 32) {
      const int col = ld_stream(cols + p);
      const double x = __ldcs(xs + p);
      const float v = (float)__ddiv_rn(x - s_mu[col], s_sigma[col]);
      zr[col] = v;
      if (zl) zl[col] = tf32_low_part(v);
    }
  }
}

int grid_for(const void* kernel, int threads, size_t smem) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (per_sm < 1) per_sm = 1;
  return SCF_NUM_SMS * per_sm;
}

}  // namespace

extern "C" int32_t scf_csr_row_sums(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                    const int64_t* row_ids, int64_t n_sel, const int32_t* col_map, double* out_sum,
                                    int32_t* out_nnz, void* stream) {
  SCF_ARG(indptr && indices && data && out_sum, "null pointer");
  SCF_ARG(n_sel >= 0, "n_sel < 0");
  if (n_sel == 0) return 0;
  const int grid = grid_for((const void*)row_sums_kernel, kThreads, 0);
  row_sums_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_sel, col_map,
                                                                out_sum, out_nnz);
  return scf_check_launch("scf_csr_row_sums");
}

extern "C" int32_t scf_csr_gene_stats(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                      const int64_t* row_ids, int64_t n_sel, int32_t n_genes, const double* row_div,
                                      double sf, unsigned long long* gene_nnz, double* gene_sum, double* gene_sumsq,
                                      void* stream) {
  SCF_ARG(indptr && indices && data && gene_nnz, "null pointer");
  SCF_ARG(n_sel >= 0 && n_genes > 0, "bad sizes");
  if (n_sel == 0) return 0;
  const int grid = grid_for((const void*)gene_stats_kernel, kThreads, 0);
  gene_stats_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_sel, row_div, sf,
                                                                  gene_nnz, gene_sum, gene_sumsq);
  return scf_check_launch("scf_csr_gene_stats");
}

extern "C" int32_t scf_csr_hvg_colstats(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                        const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                                        const double* row_sum, double sf, int32_t log_transform, int32_t n_cols,
                                        int32_t n_rep, int64_t* sum_fx, int64_t* sumsq_fx, void* stream) {
  SCF_ARG(indptr && indices && data && col_map && row_sum && sum_fx && sumsq_fx, "null pointer");
  SCF_ARG(n_sel >= 0 && n_cols > 0 && n_rep > 0, "bad sizes");
  if (n_sel == 0) return 0;
  const int grid = grid_for((const void*)hvg_colstats_kernel, kThreads, 0);
  hvg_colstats_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_sel, col_map,
                                                                    row_sum, sf, log_transform, (long long*)sum_fx,
                                                                    (long long*)sumsq_fx, n_cols, n_rep);
  return scf_check_launch("scf_csr_hvg_colstats");
}

extern "C" int32_t scf_csr_norm_scale(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                      const int64_t* row_ids, int64_t n_sel, const int32_t* col_map, int32_t n_cols,
                                      const double* row_sum, double sf, int32_t log_transform, const double* mu,
                                      const double* sigma, const double* missing_fill, float* z, float* z_lo,
                                      int64_t ldz, void* stream) {
  SCF_ARG(indptr && indices && data && col_map && row_sum && z, "null pointer");
  SCF_ARG(n_sel >= 0 && n_cols > 0 && ldz >= n_cols && (ldz & 3) == 0, "bad sizes (ldz must be a multiple of 4)");
  if (n_sel == 0) return 0;
  const size_t smem = (size_t)ldz * (8 + 8 + 4 + 4);
  SCF_ARG(smem <= 227 * 1024, "ldz too large for the shared-memory row buffers");
  cudaError_t e = cudaFuncSetAttribute(norm_scale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_csr_norm_scale: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  const int grid = grid_for((const void*)norm_scale_kernel, kThreads, smem);
  norm_scale_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_sel, col_map,
                                                                     n_cols, row_sum, sf, log_transform, mu, sigma,
                                                                     missing_fill, z, z_lo, ldz);
  return scf_check_launch("scf_csr_norm_scale");
}

extern "C" int32_t scf_csr_hvg_compact(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                       const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                                       const double* row_sum, double sf, int32_t log_transform, const int64_t* row_off,
                                       int32_t* out_col, double* out_x, int32_t n_cols, int32_t n_rep, int64_t* sum_fx,
                                       int64_t* sumsq_fx, void* stream) {
  SCF_ARG(indptr && indices && data && col_map && row_sum && row_off && out_col && out_x && sum_fx && sumsq_fx,
          "null pointer");
  SCF_ARG(n_sel >= 0 && n_cols > 0 && n_rep > 0, "bad sizes");
  if (n_sel == 0) return 0;
  const int grid = grid_for((const void*)hvg_compact_kernel, kThreads, 0);
  hvg_compact_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_sel, col_map,
                                                                   row_sum, sf, log_transform, row_off, out_col, out_x,
                                                                   (long long*)sum_fx, (long long*)sumsq_fx, n_cols,
                                                                   n_rep);
  return scf_check_launch("scf_csr_hvg_compact");
}

extern "C" int32_t scf_hvg_dense_scale(const int64_t* row_off, const int32_t* cols, const double* xs, int64_t n_sel,
                                       int32_t n_cols, const double* mu, const double* sigma, float* z, float* z_lo,
                                       int64_t ldz, void* stream) {
  SCF_ARG(row_off && cols && xs && z, "null pointer");
  SCF_ARG(n_sel >= 0 && n_cols > 0 && ldz >= n_cols && (ldz & 3) == 0, "bad sizes (ldz must be a multiple of 4)");
  if (n_sel == 0) return 0;
  const size_t smem = (size_t)ldz * (8 + 8 + 4 + 4);
  SCF_ARG(smem <= 227 * 1024, "ldz too large for the shared-memory column tables");
  cudaError_t e = cudaFuncSetAttribute(hvg_dense_scale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_hvg_dense_scale: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  const int grid = grid_for((const void*)hvg_dense_scale_kernel, kThreads, smem);
  hvg_dense_scale_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(row_off, cols, xs, n_sel, n_cols, mu, sigma, z,
                                                                          z_lo, ldz);
  return scf_check_launch("scf_hvg_dense_scale");
}
