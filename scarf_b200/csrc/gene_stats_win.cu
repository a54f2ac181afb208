// K0b, windowed: per-gene nnz / sum / sum of squares without one L2 reduction per stored value.
//
// The per-value reductions of gene_stats_kernel (csr_kernels.cu) run at the SM's reduction issue rate (~1.3 cycles
// per lane and request, B300_MICROARCH.md "Atomics"): 3 ms for the 265 M stored values of a 100k-cell shard, a tenth of
// the HBM roofline.  Within ONE row every gene occurs once, so a whole CTA can add the entries of a row into shared
// memory accumulators with plain read-modify-writes (no atomics: distinct addresses), rows follow each other
// separated by a CTA barrier.  Shared memory holds a window of GS_W genes (nnz u32, sum f64, sumsq f64 = 20 B per
// gene), so the gene axis is cut into windows and a CTA owns (block of rows) x (window); rows are sorted by gene, the
// window's entries of a row are one contiguous range whose bounds a small search kernel writes first.  Several small
// CTAs share an SM (default: 1024-gene windows, 128 threads, eight CTAs): while some wait for their next batch of rows
// to arrive from HBM or at their row barrier, the others accumulate.
#include <stdlib.h>
#include "common.cuh"

namespace {

// Tunables (template parameters of the kernel): GS_W genes per window (20 B of accumulators each), GS_THREADS per CTA
// (few warps: the per-row bookkeeping is paid per warp), GS_PER_LANE entries per thread and row held in registers (the
// rest of a long row: direct loads), GS_BATCH rows whose entries are in flight together, GS_MINB CTAs per SM.

__device__ __forceinline__ int64_t row_of(const int64_t* row_ids, int64_t r) { return row_ids ? row_ids[r] : r; }

// seg[r][w] = first position of row r (absolute, in indices / data) whose gene id is >= w * GS_W; seg[r][nwin] = row end
__global__ void gene_window_bounds_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                          const int64_t* __restrict__ row_ids, int64_t n_sel, int nwin,
                                          int GS_W, int64_t* __restrict__ seg) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sel * (nwin + 1)) return;
  const int64_t r = t / (nwin + 1);
  const int w = (int)(t - r * (nwin + 1));
  const int64_t row = row_of(row_ids, r);
  const int64_t s = indptr[row], e = indptr[row + 1];
  const int32_t key = w * GS_W;
  int64_t lo = s, hi = e;  // lower bound of key in indices[s, e)
  if (w == nwin) lo = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(indices + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  seg[t] = lo;
}

template <int GS_W, int GS_THREADS, int GS_PER_LANE, int GS_BATCH, int GS_MINB>
__global__ void __launch_bounds__(GS_THREADS, GS_MINB) gene_stats_win_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const uint32_t* __restrict__ data,
    const int64_t* __restrict__ row_ids, int64_t n_sel, const double* __restrict__ row_div, double sf,
    const int64_t* __restrict__ seg, int nwin, int n_genes, int64_t rows_per_block,
    unsigned long long* __restrict__ gene_nnz, double* __restrict__ gene_sum, double* __restrict__ gene_sumsq) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_mom = reinterpret_cast<double2*>(smem_raw);  // (sum, sum of squares) side by side: one 16-byte access
  uint32_t* s_n = reinterpret_cast<uint32_t*>(s_mom + GS_W);
  const int tid = threadIdx.x;
  const int w = blockIdx.y;
  const int g0 = w * GS_W;
  for (int j = tid; j < GS_W; j += GS_THREADS) {
    s_mom[j] = make_double2(0.0, 0.0);
    s_n[j] = 0u;
  }
  __syncthreads();
  const bool norm = row_div != nullptr;
  const bool moments = gene_sum != nullptr;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(n_sel, r_begin + rows_per_block);
  // per-row range of this window and divisor of the rows of a batch: fetched by GS_BATCH threads one batch ahead
  __shared__ long long s_base[2][GS_BATCH];
  __shared__ int s_cnt[2][GS_BATCH];
  __shared__ double s_dv[2][GS_BATCH];
  auto fetch_meta = [&](int64_t rb, int buf) {
    if (tid < GS_BATCH) {
      const int64_t r = rb + tid;
      long long base = 0;
      int cnt = 0;
      double d = 1.0;
      if (r < r_end) {
        base = __ldg(seg + r * (nwin + 1) + w);
        cnt = (int)(__ldg(seg + r * (nwin + 1) + w + 1) - base);
        // norm_lib_size = sf * counts / scalar (scarf/assay.py:51): one division per ROW (sf / scalar) and a multiply
        // per stored value; differs from the per-value division by at most one unit in the last place
        if (norm) d = __ddiv_rn(sf, __ldg(row_div + r));
      }
      s_base[buf][tid] = base, s_cnt[buf][tid] = cnt, s_dv[buf][tid] = d;
    }
  };
  fetch_meta(r_begin, 0);
  __syncthreads();
  int buf = 0;
  for (int64_t rb = r_begin; rb < r_end; rb += GS_BATCH, buf ^= 1) {
    // ---- issue the loads of this thread's entries of the batch (all in flight together) ----
    int32_t g[GS_BATCH][GS_PER_LANE];
    uint32_t c[GS_BATCH][GS_PER_LANE];
#pragma unroll
    for (int k = 0; k < GS_BATCH; ++k) {
      const long long base = s_base[buf][k];
      const int cnt = s_cnt[buf][k];
#pragma unroll
      for (int u = 0; u < GS_PER_LANE; ++u) {
        g[k][u] = 0, c[k][u] = 0u;
        if (tid + u * GS_THREADS < cnt) {
          g[k][u] = ld_stream(indices + base + tid + u * GS_THREADS);
          c[k][u] = ld_stream(data + base + tid + u * GS_THREADS);
        }
      }
    }
    fetch_meta(rb + GS_BATCH, buf ^ 1);  // visible after the first row barrier below
    // ---- values: counts * (sf / scalar) ----
    double v[GS_BATCH][GS_PER_LANE];
    if (moments) {
#pragma unroll
      for (int k = 0; k < GS_BATCH; ++k)
#pragma unroll
        for (int u = 0; u < GS_PER_LANE; ++u)
          if (c[k][u] != 0u)  // whole warps hold no entry in their second slot: they skip the division
            v[k][u] = norm ? (double)c[k][u] * s_dv[buf][k] : (double)c[k][u];
          else
            v[k][u] = 0.0;
    }
    // ---- accumulate row by row: genes are unique inside a row, so plain read-modify-writes never collide ----
#pragma unroll
    for (int k = 0; k < GS_BATCH; ++k) {
      if (rb + k >= r_end) break;  // uniform
#pragma unroll
      for (int u = 0; u < GS_PER_LANE; ++u)
        if (c[k][u] != 0u) {  // explicit zeros carry no statistics (and mark the unused register slots)
          const int j = g[k][u] - g0;
          s_n[j] += 1u;
          if (moments) {
            double2 m = s_mom[j];
            m.x += v[k][u];
            m.y += v[k][u] * v[k][u];
            s_mom[j] = m;
          }
        }
      const int cnt = s_cnt[buf][k];
      if (cnt > GS_PER_LANE * GS_THREADS) {  // long row (rare, uniform): the tail straight from memory
        const long long base = s_base[buf][k];
        const double d = s_dv[buf][k];
        for (int e = tid + GS_PER_LANE * GS_THREADS; e < cnt; e += GS_THREADS) {
          const uint32_t cc = ld_stream(data + base + e);
          if (cc != 0u) {
            const int j = ld_stream(indices + base + e) - g0;
            s_n[j] += 1u;
            if (moments) {
              const double vv = norm ? (double)cc * d : (double)cc;
              double2 m = s_mom[j];
              m.x += vv;
              m.y += vv * vv;
              s_mom[j] = m;
            }
          }
        }
      }
      __syncthreads();  // the next row may touch the same genes
    }
  }
  // ---- flush the window: one reduction per gene and CTA ----
  for (int j = tid; j < GS_W; j += GS_THREADS) {
    const int gidx = g0 + j;
    if (gidx < n_genes && s_n[j] != 0u) {
      atomicAdd(gene_nnz + gidx, (unsigned long long)s_n[j]);
      if (moments) {
        atomicAdd(gene_sum + gidx, s_mom[j].x);
        if (gene_sumsq) atomicAdd(gene_sumsq + gidx, s_mom[j].y);
      }
    }
  }
}

struct GsVariant {
  int w, threads, minb, batch;
  void (*kernel)(const int64_t*, const int32_t*, const uint32_t*, const int64_t*, int64_t, const double*, double,
                 const int64_t*, int, int, int64_t, unsigned long long*, double*, double*);
};
// measured on a 100k-cell shard (B200): 1.52 ms / 1.67 ms / 1.84 ms; wider batches, narrower windows and more entries
// per lane were all slower (2.1 - 3.0 ms), and so were shared-memory atomics in place of the row barriers (FP64 and
// 64-bit shared atomics are CAS loops, ATOMS.CAST.SPIN.64: 2.2 - 2.7 ms, round 2)
const GsVariant kVariants[] = {
    {1024, 128, 8, 8, gene_stats_win_kernel<1024, 128, 1, 8, 8>},
    {2048, 128, 5, 8, gene_stats_win_kernel<2048, 128, 2, 8, 5>},
    {4096, 256, 2, 8, gene_stats_win_kernel<4096, 256, 2, 8, 2>},
};
constexpr int kDefaultVariant = 0;
const GsVariant& gs_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCF_GS_VARIANT");  // developer switch (tools/csr_probe.py sweeps it)
    v = e ? atoi(e) : kDefaultVariant;
    if (v < 0 || v >= (int)(sizeof(kVariants) / sizeof(kVariants[0]))) v = kDefaultVariant;
  }
  return kVariants[v];
}

}  // namespace

extern "C" int64_t scf_csr_gene_stats_workspace_bytes(int64_t n_sel, int32_t n_genes) {
  const int64_t nwin = (n_genes + 1024 - 1) / 1024;  // the narrowest window any variant uses
  return n_sel * (nwin + 1) * 8;
}

extern "C" int32_t scf_csr_gene_stats_windowed(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                               const int64_t* row_ids, int64_t n_sel, int32_t n_genes,
                                               const double* row_div, double sf, unsigned long long* gene_nnz,
                                               double* gene_sum, double* gene_sumsq, void* workspace,
                                               int64_t workspace_bytes, void* stream) {
  SCF_ARG(indptr && indices && data && gene_nnz && workspace, "null pointer");
  SCF_ARG(n_sel >= 0 && n_genes > 0, "bad sizes");
  SCF_ARG(gene_sum || !gene_sumsq, "gene_sumsq needs gene_sum");
  SCF_ARG(workspace_bytes >= scf_csr_gene_stats_workspace_bytes(n_sel, n_genes), "workspace too small");
  if (n_sel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const GsVariant& var = gs_variant();
  const int nwin = (n_genes + var.w - 1) / var.w;
  int64_t* seg = (int64_t*)workspace;
  const int64_t nseg = n_sel * (nwin + 1);
  gene_window_bounds_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(indptr, indices, row_ids, n_sel, nwin, var.w,
                                                                          seg);
  int32_t rc = scf_check_launch("scf_csr_gene_stats_windowed(bounds)");
  if (rc) return rc;
  const size_t smem = (size_t)var.w * (8 + 8 + 4);
  cudaError_t e = cudaFuncSetAttribute(var.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_csr_gene_stats_windowed: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  // (row blocks) x (windows) CTAs, minb per SM: a few waves of them for balance
  int64_t blocks = (int64_t)(var.minb * SCF_NUM_SMS * 4 + nwin - 1) / nwin;
  int64_t rows_per_block = (n_sel + blocks - 1) / blocks;
  rows_per_block = (rows_per_block + var.batch - 1) / var.batch * var.batch;
  if (rows_per_block < var.batch) rows_per_block = var.batch;
  blocks = (n_sel + rows_per_block - 1) / rows_per_block;
  var.kernel<<<dim3((unsigned)blocks, (unsigned)nwin), var.threads, smem, st>>>(
      indptr, indices, data, row_ids, n_sel, row_div, sf, seg, nwin, n_genes, rows_per_block, gene_nnz, gene_sum,
      gene_sumsq);
  return scf_check_launch("scf_csr_gene_stats_windowed");
}
