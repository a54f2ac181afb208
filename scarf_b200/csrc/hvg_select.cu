// mark_hvgs after the statistics pass, fused: trend removal (MetaData.remove_trend -> fit_lowess,
// scarf/metadata.py:586-617, scarf/feat_utils.py:11-45) and the HVG choice (RNAassay.mark_hvgs,
// scarf/assay.py:1014-1063 with MetaData.multi_sift, scarf/metadata.py:483-533) on the per-gene vectors, in ONE CTA.
//
// The arithmetic is O(G) on ~30k genes and was ~120 tiny library launches (0.7 ms of a 17 ms step); here the gene
// vectors make a few passes through one CTA: statistics and logs, range of the log means, equal-width bins with the
// reference's edge arithmetic (np.histogram edges, last edge + 0.1), per-bin minimum-variance gene (first one on ties),
// LOWESS through those points (lowess_dev.cuh), exp(log var - fit(bin)), the strict bounds, and the (top_n + 1)-th
// largest corrected variance by a radix select on the order-preserving bit pattern.  Same formulas, in float64, as the
// tensor-op formulation in scarf_b200/hvg.py (which stays for CPU tensors and for callers that want the statistics).
#include <math_constants.h>
#include "lowess_dev.cuh"

namespace {

constexpr int HS_THREADS = kLowessThreads;
constexpr int HS_MAX_BINS = kLowessMaxN;

struct HvgSelectParams {
  const unsigned long long* nnz;
  const double* sum;
  const double* sumsq;
  const uint8_t* feat_i;   // bool per gene: the feature table's `I`
  const uint8_t* keep;     // bool per gene: blacklist survivors (nullable = all)
  int n_genes, n_bins, top_n;
  double m_cells, n_cells_total, lowess_frac;
  double min_cells, max_cells, min_mean, max_mean;  // strict bounds; +-inf = open
  // scratch, n_genes each
  double* la;
  double* lb;
  double* cvar;
  int* which;
  // scratch for the LOWESS points, HS_MAX_BINS each
  double* px;
  double* py;
  double* pfit;
  uint8_t* pvalid;
  // outputs
  uint8_t* hv;       // bool per gene
  int32_t* col_map;  // rank among the selected genes, -1 elsewhere (the col_map of the CSR kernels)
  int32_t* n_sel;    // number of selected genes
};

__device__ __forceinline__ unsigned long long f64_key(double v) {  // order-preserving bits (any sign)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(HS_THREADS, 1) hvg_select_kernel(const HvgSelectParams p) {
  extern __shared__ double dyn[];
  __shared__ double s_edges[HS_MAX_BINS + 1];
  __shared__ unsigned long long s_binmin[HS_MAX_BINS];
  __shared__ int s_first[HS_MAX_BINS];
  __shared__ double s_red[2][32];
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_rank;
  __shared__ int s_scan[HS_THREADS];
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.n_genes, nb = p.n_bins;
  const double inf = CUDART_INF;

  // ---- P1: statistics (RNAassay.set_feature_stats, scarf/assay.py:860-897), logs, range of log(avg) ----
  double lo = inf, hi = -inf;
  for (int g = tid; g < G; g += HS_THREADS) {
    const double tot = p.sum[g];
    const double mean = tot / p.m_cells;
    // population variance (dask var, ddof 0); separate multiply and subtract like the tensor-op formulation
    const double var = fmax(__dsub_rn(p.sumsq[g] / p.m_cells, __dmul_rn(mean, mean)), 0.0);
    const double avg = tot / p.n_cells_total;
    const bool pos = avg > 0.0 && p.feat_i[g] != 0;
    const double a = pos ? log(avg) : CUDART_NAN;
    p.la[g] = a;
    p.lb[g] = pos ? log(var) : CUDART_NAN;
    if (pos) lo = fmin(lo, a), hi = fmax(hi, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(SCF_FULL, lo, o));
    hi = fmax(hi, __shfl_xor_sync(SCF_FULL, hi, o));
  }
  if (lane == 0) s_red[0][warp] = lo, s_red[1][warp] = hi;
  for (int b = tid; b < nb; b += HS_THREADS) s_binmin[b] = ~0ull, s_first[b] = G;
  __syncthreads();
  lo = s_red[0][0], hi = s_red[1][0];
  for (int w = 1; w < HS_THREADS / 32; ++w) lo = fmin(lo, s_red[0][w]), hi = fmax(hi, s_red[1][w]);
  // np.histogram widens a zero-width range by +-0.5; edges = arange * step + first, last edge = last + 0.1
  if (lo == hi) lo = __dsub_rn(lo, 0.5), hi = __dadd_rn(hi, 0.5);
  const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)nb);
  for (int i = tid; i <= nb; i += HS_THREADS)
    s_edges[i] = i == nb ? __dadd_rn(hi, 0.1) : __dadd_rn(__dmul_rn((double)i, step), lo);
  __syncthreads();

  // ---- P2: bin of every gene (edges[i] <= la < edges[i+1]), minimum log-variance per bin ----
  for (int g = tid; g < G; g += HS_THREADS) {
    const double a = p.la[g];
    int w = -1;
    if (a == a) {  // number of edges <= a, minus one (torch.bucketize(right=True) - 1)
      int l = 0, r = nb + 1;
      while (l < r) {
        const int mid = (l + r) >> 1;
        if (s_edges[mid] <= a) l = mid + 1;
        else r = mid;
      }
      w = l - 1;
      if (w < 0 || w >= nb) w = -1;
    }
    p.which[g] = w;
    if (w >= 0) atomicMin(&s_binmin[w], f64_key(p.lb[g]));
  }
  __syncthreads();
  for (int g = tid; g < G; g += HS_THREADS) {  // the first gene (smallest index) that attains the bin's minimum
    const int w = p.which[g];
    if (w >= 0 && f64_key(p.lb[g]) == s_binmin[w]) atomicMin(&s_first[w], g);
  }
  __syncthreads();
  for (int b = tid; b < nb; b += HS_THREADS) {
    const int g = s_first[b];
    p.pvalid[b] = g < G;
    p.px[b] = g < G ? p.la[g] : 0.0;
    p.py[b] = g < G ? p.lb[g] : 0.0;
  }
  __threadfence_block();
  __syncthreads();

  // ---- P3: LOWESS through the binned points (frac, 100 robustness passes) ----
  const long long kmax = (long long)(p.lowess_frac * (double)nb + 1e-10) + 1;
  const int cache = (size_t)2 * nb * (kmax > 0 ? kmax : 1) * sizeof(double) <= 160 * 1024 ? 1 : 0;
  lowess_block(p.py, p.px, p.pvalid, nb, p.lowess_frac, 100, cache, p.pfit, dyn);
  __threadfence_block();
  __syncthreads();

  // ---- P4: corrected variance, bounds, number of eligible genes ----
  int local = 0;
  for (int g = tid; g < G; g += HS_THREADS) {
    const int w = p.which[g];
    double cv = w >= 0 ? exp(p.lb[g] - p.pfit[w]) : 0.0;  // genes outside the fit get 0 (fill_value)
    if (p.feat_i[g] == 0) cv = CUDART_NAN;
    p.cvar[g] = cv;
    const double n = (double)p.nnz[g];
    const double nzm = n > 0.0 ? p.sum[g] / fmax(n, 1.0) : 0.0;
    const bool ok = n > p.min_cells && n < p.max_cells && nzm > p.min_mean && nzm < p.max_mean && p.feat_i[g] != 0 &&
                    (p.keep == nullptr || p.keep[g] != 0) && cv == cv;
    p.which[g] = ok ? 1 : 0;  // `which` is free now: eligibility flag
    local += ok;
  }
  local = warp_sum(local);
  if (tid == 0) s_count = 0;
  __syncthreads();
  if (lane == 0 && local) atomicAdd(&s_count, local);
  __syncthreads();
  const int n_valid = s_count;
  // threshold = the (kk + 1)-th largest corrected variance of the eligible genes, kk = min(top_n, n_valid - 1) >= 0
  // (scarf/assay.py:1035-1040); selected = eligible & c_var > threshold
  long long kk = min((long long)p.top_n, (long long)n_valid - 1);
  if (kk < 0) kk = 0;
  unsigned long long thr_key = 0;  // key of -inf-like: everything eligible passes when nothing is eligible
  if (n_valid > 0) {
    // radix select from the most significant byte: find the key with exactly kk eligible keys above it
    if (tid == 0) s_prefix = 0, s_rank = kk;
    __syncthreads();
    for (int byte = 7; byte >= 0; --byte) {
      for (int i = tid; i < 256; i += HS_THREADS) s_hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long hi_mask = byte == 7 ? 0ull : (~0ull << (8 * (byte + 1)));
      for (int g = tid; g < G; g += HS_THREADS)
        if (p.which[g]) {
          const unsigned long long key = f64_key(p.cvar[g]);
          if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> (8 * byte)) & 255], 1u);
        }
      __syncthreads();
      if (tid == 0) {
        long long rank = s_rank;  // number of keys (inside the current prefix) that must lie above the answer
        int d = 255;
        for (; d > 0; --d) {
          if (rank < (long long)s_hist[d]) break;
          rank -= s_hist[d];
        }
        s_rank = rank;
        s_prefix = prefix | ((unsigned long long)d << (8 * byte));
      }
      __syncthreads();
    }
    thr_key = s_prefix;
  }
  // ---- P5: mask, rank of every selected gene (exclusive scan over the genes in index order) ----
  const int per = (G + HS_THREADS - 1) / HS_THREADS;
  const int g0 = tid * per, g1 = min(G, g0 + per);
  int cnt = 0;
  for (int g = g0; g < g1; ++g) {
    const bool sel = p.which[g] && f64_key(p.cvar[g]) > thr_key;
    p.hv[g] = sel;
    cnt += sel;
  }
  s_scan[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < HS_THREADS; o <<= 1) {  // Hillis-Steele inclusive scan
    const int v = tid >= o ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  int pos = s_scan[tid] - cnt;
  for (int g = g0; g < g1; ++g) p.col_map[g] = p.hv[g] ? pos++ : -1;
  if (tid == HS_THREADS - 1) *p.n_sel = s_scan[tid];
}

}  // namespace

extern "C" int64_t scf_hvg_select_workspace_bytes(int32_t n_genes) {
  return (int64_t)n_genes * (3 * 8 + 4) + (int64_t)HS_MAX_BINS * (3 * 8 + 1) + 256;
}

extern "C" int32_t scf_hvg_select(const unsigned long long* gene_nnz, const double* gene_sum, const double* gene_sumsq,
                                  const uint8_t* feat_i, const uint8_t* keep, int32_t n_genes, double m_cells,
                                  double n_cells_total, int32_t n_bins, double lowess_frac, int32_t top_n,
                                  double min_cells, double max_cells, double min_mean, double max_mean, uint8_t* hv,
                                  int32_t* col_map, int32_t* n_sel, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
  SCF_ARG(gene_nnz && gene_sum && gene_sumsq && feat_i && hv && col_map && n_sel && workspace, "null pointer");
  SCF_ARG(n_genes > 0 && n_bins >= 1 && n_bins <= HS_MAX_BINS && top_n >= 0 && m_cells > 0 && n_cells_total > 0,
          "bad sizes (n_bins must be within [1, 512])");
  SCF_ARG(workspace_bytes >= scf_hvg_select_workspace_bytes(n_genes), "workspace too small");
  HvgSelectParams p;
  p.nnz = gene_nnz, p.sum = gene_sum, p.sumsq = gene_sumsq, p.feat_i = feat_i, p.keep = keep;
  p.n_genes = n_genes, p.n_bins = n_bins, p.top_n = top_n;
  p.m_cells = m_cells, p.n_cells_total = n_cells_total, p.lowess_frac = lowess_frac;
  p.min_cells = min_cells, p.max_cells = max_cells, p.min_mean = min_mean, p.max_mean = max_mean;
  unsigned char* ws = (unsigned char*)workspace;
  p.la = (double*)ws, ws += (size_t)n_genes * 8;
  p.lb = (double*)ws, ws += (size_t)n_genes * 8;
  p.cvar = (double*)ws, ws += (size_t)n_genes * 8;
  p.px = (double*)ws, ws += (size_t)HS_MAX_BINS * 8;
  p.py = (double*)ws, ws += (size_t)HS_MAX_BINS * 8;
  p.pfit = (double*)ws, ws += (size_t)HS_MAX_BINS * 8;
  p.which = (int*)ws, ws += (size_t)n_genes * 4;
  p.pvalid = (uint8_t*)ws;
  p.hv = hv, p.col_map = col_map, p.n_sel = n_sel;
  const long long kmax = (long long)(lowess_frac * (double)n_bins + 1e-10) + 1;
  size_t dyn = (size_t)2 * n_bins * (kmax > 0 ? kmax : 1) * sizeof(double);
  if (dyn > 160 * 1024) dyn = 0;
  cudaError_t e = cudaFuncSetAttribute(hvg_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) {
    scf_set_error("scf_hvg_select: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  hvg_select_kernel<<<1, HS_THREADS, dyn, (cudaStream_t)stream>>>(p);
  return scf_check_launch("scf_hvg_select");
}
