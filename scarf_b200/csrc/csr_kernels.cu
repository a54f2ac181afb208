// K0 / K1: CSR row kernels (warp per cell, persistent grid).  HBM-bound: the algorithmic traffic
// is 8 B per stored value (int32 gene id + uint32 count) + 8 B of indptr per cell.
#include <math_constants.h>
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kCompactVariant = 2;  // scf_csr_hvg_compact: see the variant table at its entry point

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

// Counts are small integers and the divisor is fixed inside a row, so the FP64 division + log1p (~250 instructions,
// the bulk of what these kernels issued) is evaluated once per row for the counts 1..kNormTab -- two per lane, into a
// per-warp shared-memory table -- and looked up per stored value; larger counts take the direct evaluation.  Same
// function on the same arguments: bit-identical values.
constexpr int kNormTab = 64;
__device__ __forceinline__ void fill_norm_table(double* tab, int lane, double s, double sf, bool log_transform) {
  __syncwarp();  // the previous row's lookups are done
  tab[lane] = norm_value((uint32_t)lane + 1u, s, sf, log_transform);
  tab[lane + 32] = norm_value((uint32_t)lane + 33u, s, sf, log_transform);
  __syncwarp();
}
__device__ __forceinline__ double norm_lookup(const double* tab, uint32_t c, double s, double sf, bool log_transform) {
  return c - 1u < (uint32_t)kNormTab ? tab[c - 1u] : norm_value(c, s, sf, log_transform);
}
__device__ __forceinline__ unsigned long long colstat_fx(double v) {  // == to_fx(v, SCF_COLSTAT_SHIFT): exact scaling
  return (unsigned long long)__double2ll_rn(v * (double)(1ull << SCF_COLSTAT_SHIFT));
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
  __shared__ double s_tab[kWarpsPerCta][kNormTab];
  sum_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  sumsq_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  const int lane = threadIdx.x & 31;
  double* tab = s_tab[threadIdx.x >> 5];
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    double sc = row_sum[r];
    if (sc == 0.0) sc = 1.0;  // scalar[scalar == 0] = 1  (scarf/assay.py:821-823)
    fill_norm_table(tab, lane, sc, sf, log_transform != 0);
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      const int col = __ldg(col_map + g);
      if (col < 0 || c == 0) return;
      const double x = norm_lookup(tab, c, sc, sf, log_transform != 0);
      atomicAdd((unsigned long long*)(sum_fx + col), colstat_fx(x));
      atomicAdd((unsigned long long*)(sumsq_fx + col), colstat_fx(x * x));
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
  __shared__ double s_tab[kWarpsPerCta][kNormTab];
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
    fill_norm_table(s_tab[warp], lane, sc, sf, log_transform != 0);
    __syncwarp();  // orders the base row before the overwrites below (different lanes hit the same addresses)
    warp_row_scan(indices, data, s, e, lane, [&](int32_t g, uint32_t c) {
      const int col = __ldg(col_map + g);
      if (col < 0) return;
      if (missing_fill) {  // a column the target matrix does not really have keeps its fill value
        const double f = __ldg(missing_fill + col);
        if (f == f) return;
      }
      const double x = norm_lookup(s_tab[warp], c, sc, sf, log_transform != 0);
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
//
// Column sums: one 64-bit reduction per value and statistic into one of n_rep copies in L2 (106 M at C2).
// Measured at C2 (100k cells, tools/csr_probe.py; 1.25 ms before): the per-row lookup table above 1.13 ms; eight
// stored values per lane and step 1.04 ms; the next stretch of the row loading while the current one is looked up
// and packed 0.85 ms (PREFETCH 1); a third stage that also keeps the col_map look-ups of the next stretch in flight
// 0.83 ms (PREFETCH 2: no further gain, the two load latencies are hidden by then).  Per-CTA shared-memory
// accumulators for the column sums (two 32-bit ATOMS.ADD with carry per sum -- shared memory has no native 64-bit
// add -- flushed once per CTA) were SLOWER (1.47 ms): the L2 reductions are not what bounds the kernel, and the
// accumulators cost occupancy.
// dynamic smem: [kWarpsPerCta][kNormTab] double | [kWarpsPerCta][BATCH] int32 col | [kWarpsPerCta][BATCH] uint32 count
template <int UNROLL, int BATCH, int MIN_CTAS, int PREFETCH>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) hvg_compact_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const uint32_t* __restrict__ data,
    const int64_t* __restrict__ row_ids, int64_t n_sel, const int32_t* __restrict__ col_map,
    const double* __restrict__ row_sum, double sf, int log_transform, const int64_t* __restrict__ row_off,
    int32_t* __restrict__ out_col, double* __restrict__ out_x, long long* __restrict__ sum_fx,
    long long* __restrict__ sumsq_fx, int n_cols, int n_rep) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_tab = reinterpret_cast<double*>(smem_raw);
  int32_t* s_col = reinterpret_cast<int32_t*>(s_tab + kWarpsPerCta * kNormTab);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_col + kWarpsPerCta * BATCH);
  sum_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  sumsq_fx += (int64_t)(blockIdx.x % n_rep) * n_cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int32_t* bc = s_col + warp * BATCH;
  uint32_t* bn = s_cnt + warp * BATCH;
  double* tab = s_tab + warp * kNormTab;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    const int64_t row = row_of(row_ids, r);
    const int64_t s = indptr[row], e = indptr[row + 1];
    double sc = row_sum[r];
    if (sc == 0.0) sc = 1.0;  // scalar[scalar == 0] = 1  (scarf/assay.py:821-823)
    int64_t cursor = row_off[r];
    int pending = 0;
    fill_norm_table(tab, lane, sc, sf, log_transform != 0);
    auto flush = [&]() {
      __syncwarp();
      for (int i = lane; i < pending; i += 32) {
        const int col = bc[i];
        const double x = norm_lookup(tab, bn[i], sc, sf, log_transform != 0);
        const unsigned long long fx = colstat_fx(x), fxx = colstat_fx(x * x);
        atomicAdd((unsigned long long*)(sum_fx + col), fx);
        atomicAdd((unsigned long long*)(sumsq_fx + col), fxx);
        out_col[cursor + i] = col;
        out_x[cursor + i] = x;
      }
      cursor += pending;
      pending = 0;
      __syncwarp();
    };
    // PREFETCH 0: load, look up, pack.  1: the next stretch of the row is in flight while this one is looked up and
    // packed.  2: three stages -- stretch i + 2 loading, the col_map look-ups of stretch i + 1 in flight, stretch i
    // packed: neither of the two dependent global latencies is exposed.
    constexpr int S = 32 * UNROLL;
    int32_t col[UNROLL], g1[UNROLL], g2[UNROLL];
    uint32_t c[UNROLL], c1[UNROLL], c2[UNROLL];
    auto load = [&](int32_t* gg, uint32_t* cc, int64_t p) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t pp = p + 32 * u;
        const bool ok = pp < e;
        gg[u] = ok ? ld_stream(indices + pp) : -1;
        cc[u] = ok ? ld_stream(data + pp) : 0u;
      }
    };
    auto lookup = [&](int32_t* out, const int32_t* gg) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) out[u] = gg[u] >= 0 ? __ldg(col_map + gg[u]) : -1;
    };
    const int64_t p0 = s + lane;
    if (PREFETCH == 1) load(g1, c1, p0);
    if (PREFETCH == 2) {
      load(g1, c, p0);
      load(g2, c2, p0 + S);
      lookup(col, g1);
    }
    for (int64_t p = p0; p - lane < e; p += S) {
      if (PREFETCH == 0) {
        load(g1, c, p);
        lookup(col, g1);
      } else if (PREFETCH == 1) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) g2[u] = g1[u], c[u] = c1[u];
        load(g1, c1, p + S);
        lookup(col, g2);
      } else {
        load(g1, c1, p + 2 * S);
        lookup(g2, g2);  // column ids of the next stretch replace its gene ids; consumed one iteration later
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const bool hit = col[u] >= 0 && c[u] != 0u;
        const unsigned m = __ballot_sync(SCF_FULL, hit);
        if (hit) {
          const int pos = pending + __popc(m & lt_mask);
          bc[pos] = col[u];
          bn[pos] = c[u];
        }
        pending += __popc(m);
      }
      if (pending > BATCH - S) flush();
      if (PREFETCH == 2) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) col[u] = g2[u], c[u] = c2[u], g2[u] = g1[u], c2[u] = c1[u];
      }
    }
    flush();
  }
}

constexpr int kDenseSeg = 512;  // columns per segment of hvg_dense_scale when two planes are written
constexpr int kDensePf = 3;     // chunks of 32 entries loaded ahead of the one being written

// K1b': Z (and the 3xTF32 low plane) from the compact matrix: base row from shared memory, then the row's entries.
__global__ void __launch_bounds__(kThreads) hvg_dense_scale_kernel(
    const int64_t* __restrict__ row_off, const int32_t* __restrict__ cols, const double* __restrict__ xs,
    int64_t n_sel, int n_cols, const double* __restrict__ mu, const double* __restrict__ sigma,
    float* __restrict__ z, float* __restrict__ z_lo, int64_t ldz, int seg_vecs) {
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
  // A row is written segment by segment: the "no count" values of 4 * seg_vecs columns, then at once the entries the
  // cell has in those columns (the compact matrix keeps a row's columns ascending).  Written as a whole row first and
  // patched afterwards, part of the row had left L2 by the time the patches arrived: 3.3 GB of DRAM traffic for
  // 2.3 GB of algorithmic bytes (ncu launch list of the step).  The chunk of entries in hand and the next ones stay in
  // registers (a queue of kDensePf + 1 chunks), so no step waits on a load it has just issued.
  const int nvec = (int)(ldz >> 2);
  for (int64_t r = warp0; r < n_sel; r += nwarps) {
    float* zr = z + r * ldz;
    float* zl = z_lo ? z_lo + r * ldz : nullptr;
    const int64_t e = row_off[r + 1];
    int64_t p = row_off[r];  // first entry of the chunk in hand (warp-uniform)
    auto load = [&](int64_t q, int& c, double& x) {
      const bool ok = q + lane < e;
      c = ok ? ld_stream(cols + q + lane) : INT_MAX;
      x = ok ? __ldcs(xs + q + lane) : 0.0;
    };
    if (!zl) {  // one plane: the rows in flight fit in L2, whole-row order with four loads ahead of the stores
      for (int j = lane; j < nvec; j += 32)
        reinterpret_cast<float4*>(zr)[j] = reinterpret_cast<const float4*>(s_base)[j];
      __syncwarp();
      for (int64_t q = p + lane; q < e; q += 32 * 4) {
        int c4[4];
        double x4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t qq = q + 32 * u;
          c4[u] = qq < e ? ld_stream(cols + qq) : -1;
          x4[u] = qq < e ? __ldcs(xs + qq) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c4[u] >= 0) zr[c4[u]] = (float)__ddiv_rn(x4[u] - s_mu[c4[u]], s_sigma[c4[u]]);
      }
      continue;
    }
    int colq[kDensePf + 1];  // the chunk in hand and the kDensePf chunks after it
    double xq[kDensePf + 1];
#pragma unroll
    for (int u = 0; u <= kDensePf; ++u) load(p + 32 * u, colq[u], xq[u]);
    for (int v0 = 0; v0 < nvec; v0 += seg_vecs) {
      const int v1 = min(v0 + seg_vecs, nvec);
      for (int j = v0 + lane; j < v1; j += 32) {
        reinterpret_cast<float4*>(zr)[j] = reinterpret_cast<const float4*>(s_base)[j];
        if (zl) reinterpret_cast<float4*>(zl)[j] = reinterpret_cast<const float4*>(s_base_lo)[j];
      }
      __syncwarp();  // orders the base values before the entries below (other lanes, same addresses)
      const int seg0 = 4 * v0, seg1 = 4 * v1;
      while (p < e) {
        const int col = colq[0];
        if (col >= seg0 && col < seg1) {
          const float v = (float)__ddiv_rn(xq[0] - s_mu[col], s_sigma[col]);
          zr[col] = v;
          if (zl) zl[col] = tf32_low_part(v);
        }
        if (__shfl_sync(SCF_FULL, col, 31) >= seg1) break;  // the chunk reaches into the next segment: kept in hand
        p += 32;
#pragma unroll
        for (int u = 0; u < kDensePf; ++u) colq[u] = colq[u + 1], xq[u] = xq[u + 1];
        load(p + 32 * kDensePf, colq[kDensePf], xq[kDensePf]);
      }
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
  // variants {stored values per lane and step, batch entries per warp, CTAs per SM aimed at, pipeline depth}
  struct Variant { int batch; const void* fn; };
  static const Variant kVariants[] = {
      {512, (const void*)hvg_compact_kernel<4, 512, 4, 0>}, {512, (const void*)hvg_compact_kernel<8, 512, 5, 0>},
      {512, (const void*)hvg_compact_kernel<8, 512, 3, 1>}, {256, (const void*)hvg_compact_kernel<4, 256, 4, 1>},
      {512, (const void*)hvg_compact_kernel<8, 512, 3, 2>},
  };
  constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
  const char* env = getenv("SCF_COMPACT_VARIANT");  // developer switch (tools/csr_probe.py sweeps it)
  int variant = env ? atoi(env) : kCompactVariant;
  if (variant < 0 || variant >= kNumVariants) variant = kCompactVariant;
  const size_t smem = (size_t)kWarpsPerCta * (kNormTab * 8 + kVariants[variant].batch * 8);
  const void* fn = kVariants[variant].fn;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_csr_hvg_compact: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  const int grid = grid_for(fn, kThreads, smem);
  long long* sfx = (long long*)sum_fx;
  long long* sqfx = (long long*)sumsq_fx;
  void* args[] = {&indptr, &indices, &data, &row_ids, &n_sel, &col_map, &row_sum, &sf, &log_transform, &row_off,
                  &out_col, &out_x, &sfx, &sqfx, &n_cols, &n_rep};
  e = cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), args, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    scf_set_error("scf_csr_hvg_compact: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
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
  // (one plane: whole-row order inside the kernel, 0.31 ms at C2 against 0.44 ms segmented)
  const char* env = getenv("SCF_DENSE_SEG");  // developer switch (columns per segment, two-plane mode)
  int seg = env ? atoi(env) : kDenseSeg;
  if (seg < 4 || seg > ldz) seg = (int)ldz;
  hvg_dense_scale_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(row_off, cols, xs, n_sel, n_cols, mu, sigma, z,
                                                                          z_lo, ldz, (seg + 3) / 4);
  return scf_check_launch("scf_hvg_dense_scale");
}
