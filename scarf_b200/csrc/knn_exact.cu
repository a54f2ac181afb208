// K5, method 0: brute-force kNN with the oracle's arithmetic in FP64 (oracle/oracle_c.c defines it):
//   d(a,b) = (float) sum_t ((double)a_t - (double)b_t)^2, t ascending, separate mul and add (no FMA),
//   neighbours ordered by (d, index).
// It is the exactness anchor of the library: the tcgen05 path (knn_tc.cu) re-ranks its candidates with
// the same arithmetic and sends every query whose guard band cannot be proven through this kernel.
#include <float.h>
#include <algorithm>
#include "common.cuh"
#include "knn_common.cuh"

namespace {

constexpr int TQ = 64, TR = 64, DK = 32;

// dynamic smem layout: Qs[dim_pad][TQ] | Rs[DK][TR+1] | Ds[TQ][TR+1] | Ld[TQ][k] | Li[TQ][k]
// q_ids (nullable): list of query rows to compute (results go to those rows); n_ids_dev (nullable): device
// scalar holding the length of that list (the tcgen05 path's fail list, no host round trip).
__global__ void __launch_bounds__(256) knn_exact_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_ids,
                                                        const int* __restrict__ n_ids_dev, int64_t nq,
                                                        const float* __restrict__ ref, int64_t nref, int dim,
                                                        int64_t ld, int64_t ldr, int k, int64_t self_offset,
                                                        int64_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim_pad = (dim + DK - 1) / DK * DK;
  float* Qs = reinterpret_cast<float*>(smem_raw);
  float* Rs = Qs + (size_t)dim_pad * TQ;
  float* Ds = Rs + DK * (TR + 1);
  float* Ld = Ds + TQ * (TR + 1);
  int* Li = reinterpret_cast<int*>(Ld + (size_t)TQ * k);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (n_ids_dev) nq = *n_ids_dev;
  const int64_t n_qtiles = (nq + TQ - 1) / TQ;
 for (int64_t item = blockIdx.x; item < n_qtiles; item += gridDim.x) {
  const int64_t q0 = item * TQ;
  const int64_t ref_begin = 0, ref_end = nref;
  __syncthreads();  // previous item's lists / Qs fully consumed
  // query tile, transposed, zero padded
  for (int e = tid; e < dim_pad * TQ; e += 256) {
    const int t = e % dim_pad, qq = e / dim_pad;
    float v = 0.f;
    if (q0 + qq < nq && t < dim) {
      const int64_t qi = q_ids ? q_ids[q0 + qq] : q0 + qq;
      v = q[qi * ld + t];
    }
    Qs[t * TQ + qq] = v;
  }
  int cnt = 0;  // threads 0..TQ-1: entries in this query's list
  int64_t self = -1;
  if (tid < TQ && q0 + tid < nq && self_offset >= 0) self = (q_ids ? q_ids[q0 + tid] : q0 + tid) + self_offset;
  __syncthreads();
  for (int64_t r0 = ref_begin; r0 < ref_end; r0 += TR) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int t0 = 0; t0 < dim_pad; t0 += DK) {
      // ref chunk: TR rows x DK dims, coalesced along t
      for (int e = tid; e < TR * DK; e += 256) {
        const int t = e % DK, rr = e / DK;
        float v = 0.f;
        if (r0 + rr < ref_end && t0 + t < dim) v = __ldg(ref + (r0 + rr) * ldr + t0 + t);
        Rs[t * (TR + 1) + rr] = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int t = 0; t < DK; ++t) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (double)Qs[(t0 + t) * TQ + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = (double)Rs[t * (TR + 1) + tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double df = __dsub_rn(a[i], b[j]);
            acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(df, df));
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ds[(ty * 4 + i) * (TR + 1) + tx * 4 + j] = (float)acc[i][j];
    __syncthreads();
    if (tid < TQ && q0 + tid < nq) {
      float* ld_ = Ld + (size_t)tid * k;
      int* li_ = Li + (size_t)tid * k;
      const int lim = (int)min((int64_t)TR, ref_end - r0);
      for (int rr = 0; rr < lim; ++rr) {
        const int64_t j = r0 + rr;
        if (j == self) continue;
        const float d = Ds[tid * (TR + 1) + rr];
        if (cnt == k && !(d < ld_[k - 1])) continue;  // equal d with a larger index never wins
        int p = cnt < k ? cnt : k - 1;
        while (p > 0 && d < ld_[p - 1]) {
          ld_[p] = ld_[p - 1];
          li_[p] = li_[p - 1];
          --p;
        }
        ld_[p] = d;
        li_[p] = (int)j;
        if (cnt < k) ++cnt;
      }
    }
    // the next tile's first __syncthreads (after loading Rs) orders these reads of Ds before its rewrite
  }
  if (tid < TQ && q0 + tid < nq) {
    const int64_t qi = q_ids ? q_ids[q0 + tid] : q0 + tid;
    for (int p = 0; p < k; ++p) {
      out_idx[qi * k + p] = p < cnt ? (int64_t)Li[(size_t)tid * k + p] : -1;
      out_dist[qi * k + p] = p < cnt ? Ld[(size_t)tid * k + p] : FLT_MAX;
    }
  }
 }
}

}  // namespace

size_t knn_exact_smem(int dim, int k) {
  const int dim_pad = (dim + DK - 1) / DK * DK;
  return sizeof(float) * ((size_t)dim_pad * TQ + DK * (TR + 1) + TQ * (TR + 1)) + (size_t)TQ * k * 8;
}

static int32_t exact_prepare(int dim, int k, size_t& smem) {
  smem = knn_exact_smem(dim, k);
  if (smem > 227 * 1024) {
    scf_set_error("scf_knn_l2: dim/k too large for the exact kernel (%zu B of shared memory)", smem);
    return 1;
  }
  cudaError_t e = cudaFuncSetAttribute(knn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  return 0;
}

int32_t knn_exact_launch(const float* q, const int64_t* q_ids, const int* n_ids_dev, int64_t nq, const float* ref,
                         int64_t nref, int dim, int64_t ld, int64_t ldr, int k, int64_t self_offset, int64_t* out_idx,
                         float* out_dist, cudaStream_t stream) {
  if (nq == 0) return 0;
  size_t smem;
  int32_t rc = exact_prepare(dim, k, smem);
  if (rc) return rc;
  // with a device-side count the grid is a fixed persistent size and every CTA strides over the list
  const int64_t tiles = (nq + TQ - 1) / TQ;
  const unsigned grid = (unsigned)(n_ids_dev ? (tiles < 4 * SCF_NUM_SMS ? tiles : 4 * SCF_NUM_SMS) : tiles);
  knn_exact_kernel<<<grid, 256, smem, stream>>>(q, q_ids, n_ids_dev, nq, ref, nref, dim, ld, ldr, k, self_offset,
                                                out_idx, out_dist);
  return scf_check_launch("scf_knn_l2(exact)");
}

// ---------------------------------------------------------------------------------------------------------
// Repair of the rows the tcgen05 path could not prove (knn_tc.cu): the re-rank kernel leaves, per failed row,
// the (distance, index) key of its current k-th candidate.  The true k nearest all have keys <= that key, so one
// threshold scan over the references with the oracle's FP64 arithmetic collects them (normally k plus a handful);
// a warp then picks the k smallest.  Rows whose list overflows (or that had fewer than k candidates, or that do not
// fit the scratch) go through knn_exact_kernel.
//   scratch: [FIX_ROWS] int cnt | [FIX_ROWS][FIX_LIST] u64 keys | [nq] i64 slow_ids | int slow_count
namespace {

constexpr int FIX_QG = 8;         // failed rows per work item (each staged reference is used FIX_QG times)
constexpr int FIX_LIST = 256;     // collected keys per row
constexpr int FIX_ROWS = KNN_FIX_CAP;
constexpr int FIX_MAXDIM = 128;

__global__ void __launch_bounds__(256) knn_fix_scan_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_ids,
                                                           const unsigned long long* __restrict__ q_keys,
                                                           const int* __restrict__ n_ids_dev,
                                                           const float* __restrict__ ref, int64_t nref, int dim,
                                                           int64_t ld, int64_t self_offset, int refs_per_split,
                                                           int nsplit, int* __restrict__ cnt,
                                                           unsigned long long* __restrict__ lists) {
  __shared__ double aq[FIX_QG][FIX_MAXDIM];
  __shared__ unsigned long long kth[FIX_QG];
  __shared__ long long selfs[FIX_QG];
  const int n = min(*n_ids_dev, FIX_ROWS);
  const int ngroups = (n + FIX_QG - 1) / FIX_QG;
  for (int item = blockIdx.x; item < ngroups * nsplit; item += gridDim.x) {
    const int g = item / nsplit, sp = item % nsplit;
    __syncthreads();
    for (int e = threadIdx.x; e < FIX_QG * dim; e += blockDim.x) {
      const int qq = e / dim, t = e % dim;
      const int slot = g * FIX_QG + qq;
      aq[qq][t] = slot < n ? (double)q[q_ids[slot] * ld + t] : 0.0;
    }
    if (threadIdx.x < FIX_QG) {
      const int slot = g * FIX_QG + threadIdx.x;
      kth[threadIdx.x] = slot < n ? q_keys[slot] : 0ull;  // key 0: nothing passes
      selfs[threadIdx.x] = (slot < n && self_offset >= 0) ? q_ids[slot] + self_offset : -1;
    }
    __syncthreads();
    const int64_t r0 = (int64_t)sp * refs_per_split, r1 = min(r0 + (int64_t)refs_per_split, nref);
    for (int64_t j = r0 + threadIdx.x; j < r1; j += blockDim.x) {
      const float* b = ref + j * ld;
      double acc[FIX_QG];
#pragma unroll
      for (int qq = 0; qq < FIX_QG; ++qq) acc[qq] = 0.0;
      int t = 0;
      for (; t + 4 <= dim; t += 4) {  // rows are 16-byte aligned (ld is a multiple of 4 floats, checked by the host)
        const float4 v = __ldg(reinterpret_cast<const float4*>(b + t));
        const double bv[4] = {(double)v.x, (double)v.y, (double)v.z, (double)v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int qq = 0; qq < FIX_QG; ++qq) {
            const double df = __dsub_rn(aq[qq][t + u], bv[u]);
            acc[qq] = __dadd_rn(acc[qq], __dmul_rn(df, df));
          }
      }
      for (; t < dim; ++t) {
        const double bvv = (double)__ldg(b + t);
#pragma unroll
        for (int qq = 0; qq < FIX_QG; ++qq) {
          const double df = __dsub_rn(aq[qq][t], bvv);
          acc[qq] = __dadd_rn(acc[qq], __dmul_rn(df, df));
        }
      }
#pragma unroll
      for (int qq = 0; qq < FIX_QG; ++qq) {
        const unsigned long long key = ((unsigned long long)__float_as_uint((float)acc[qq]) << 32) | (unsigned)j;
        if (key <= kth[qq] && j != selfs[qq]) {
          const int slot = g * FIX_QG + qq;
          const int pos = atomicAdd(cnt + slot, 1);
          if (pos < FIX_LIST) lists[(size_t)slot * FIX_LIST + pos] = key;
        }
      }
    }
  }
}

// one warp per failed row: the k smallest collected keys; rows that cannot be finished here go on the slow list
__global__ void __launch_bounds__(256) knn_fix_select_kernel(const int64_t* __restrict__ q_ids,
                                                             const int* __restrict__ n_ids_dev, int k,
                                                             const int* __restrict__ cnt,
                                                             const unsigned long long* __restrict__ lists,
                                                             int64_t* __restrict__ out_idx, float* __restrict__ out_dist,
                                                             int64_t* __restrict__ slow_ids, int* __restrict__ slow_count) {
  const int nfail = *n_ids_dev;
  const int lane = threadIdx.x & 31;
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < nfail; w += gridDim.x * 8) {
    const int c = w < FIX_ROWS ? cnt[w] : -1;
    if (c < k || c > FIX_LIST) {  // overflow, beyond the scratch, or an unusable threshold
      if (lane == 0) slow_ids[atomicAdd(slow_count, 1)] = q_ids[w];
      continue;
    }
    const int64_t qi = q_ids[w];
    unsigned long long key[FIX_LIST / 32];
#pragma unroll
    for (int u = 0; u < FIX_LIST / 32; ++u) {
      const int e = lane + 32 * u;
      key[u] = e < c ? lists[(size_t)w * FIX_LIST + e] : ~0ull;
    }
    for (int r = 0; r < k; ++r) {
      unsigned long long best = key[0];
#pragma unroll
      for (int u = 1; u < FIX_LIST / 32; ++u) best = best < key[u] ? best : key[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(SCF_FULL, best, o);
        best = other < best ? other : best;
      }
#pragma unroll
      for (int u = 0; u < FIX_LIST / 32; ++u)
        if (key[u] == best) key[u] = ~0ull;  // keys are unique (distinct reference ids)
      if (lane == 0) {
        out_idx[qi * k + r] = (int64_t)(best & 0xffffffffull);
        out_dist[qi * k + r] = __uint_as_float((unsigned)(best >> 32));
      }
    }
  }
}

}  // namespace

size_t knn_exact_fix_scratch_bytes(int64_t nq, int k) {
  (void)k;
  return (size_t)FIX_ROWS * 4 + (size_t)FIX_ROWS * FIX_LIST * 8 + (size_t)nq * 8 + 256;
}

int32_t knn_exact_fix_launch(const float* q, const int64_t* q_ids, const unsigned long long* q_keys,
                             const int* n_ids_dev, int64_t nq_max, const float* ref, int64_t nref, int dim, int64_t ld,
                             int k, int64_t self_offset, int64_t* out_idx, float* out_dist, void* scratch,
                             cudaStream_t stream) {
  if (nq_max == 0) return 0;
  size_t smem;
  int32_t rc = exact_prepare(dim, k, smem);
  if (rc) return rc;
  unsigned char* sc = (unsigned char*)scratch;
  int* cnt = (int*)sc;
  unsigned long long* lists = (unsigned long long*)(sc + (size_t)FIX_ROWS * 4);
  int64_t* slow_ids = (int64_t*)(sc + (size_t)FIX_ROWS * 4 + (size_t)FIX_ROWS * FIX_LIST * 8);
  int* slow_count = (int*)(sc + (size_t)FIX_ROWS * 4 + (size_t)FIX_ROWS * FIX_LIST * 8 + (size_t)nq_max * 8);
  cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)FIX_ROWS * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(slow_count, 0, 4, stream);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2(fix): %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  const bool scan_ok = dim <= FIX_MAXDIM && (ld & 3) == 0 && k <= FIX_LIST;
  if (scan_ok) {
    int refs_per_split = (int)std::max<int64_t>(1024, (nref + 63) / 64);
    refs_per_split = (refs_per_split + 255) / 256 * 256;
    const int nsplit = (int)((nref + refs_per_split - 1) / refs_per_split);
    knn_fix_scan_kernel<<<4 * SCF_NUM_SMS, 256, 0, stream>>>(q, q_ids, q_keys, n_ids_dev, ref, nref, dim, ld,
                                                             self_offset, refs_per_split, nsplit, cnt, lists);
    rc = scf_check_launch("scf_knn_l2(fix,scan)");
    if (rc) return rc;
  }  // otherwise every count stays 0 < k and all failed rows take the slow path
  knn_fix_select_kernel<<<SCF_NUM_SMS, 256, 0, stream>>>(q_ids, n_ids_dev, k, cnt, lists, out_idx, out_dist, slow_ids,
                                                         slow_count);
  rc = scf_check_launch("scf_knn_l2(fix,select)");
  if (rc) return rc;
  const int64_t tiles = (nq_max + TQ - 1) / TQ;
  const unsigned grid = (unsigned)(tiles < 4 * SCF_NUM_SMS ? tiles : 4 * SCF_NUM_SMS);
  knn_exact_kernel<<<grid, 256, smem, stream>>>(q, slow_ids, slow_count, nq_max, ref, nref, dim, ld, ld, k, self_offset,
                                                out_idx, out_dist);
  return scf_check_launch("scf_knn_l2(fix,slow)");
}
