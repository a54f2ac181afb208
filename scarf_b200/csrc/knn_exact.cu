// K5, method 0: brute-force kNN with the oracle's arithmetic in FP64 (oracle/oracle_c.c defines it):
//   d(a,b) = (float) sum_t ((double)a_t - (double)b_t)^2, t ascending, separate mul and add (no FMA),
//   neighbours ordered by (d, index).
// It is the exactness anchor of the library: the tcgen05 path (knn_tc.cu) re-ranks its candidates with
// the same arithmetic and sends every query whose guard band cannot be proven through this kernel.
#include <float.h>
#include "common.cuh"
#include "knn_common.cuh"

namespace {

constexpr int TQ = 64, TR = 64, DK = 32;

// dynamic smem layout: Qs[dim_pad][TQ] | Rs[DK][TR+1] | Ds[TQ][TR+1] | Ld[TQ][k] | Li[TQ][k]
// q_ids (nullable): list of query rows to compute (results go to those rows); n_ids_dev (nullable): device
// scalar holding the length of that list (the tcgen05 path's fail list, no host round trip).
// nsplit > 1: the references are cut into nsplit ranges, work item = (query tile, range), partial top-k lists go
// to part_d / part_i [(tile * nsplit + range) * TQ + row][k] and knn_merge_kernel picks the final k: this keeps
// all SMs busy when only a few query rows are recomputed.  gate: 0 = always run, 1 = run only if the device-side
// count is <= gate_cap (split mode), 2 = run only if it is > gate_cap (one item per query tile).
__global__ void __launch_bounds__(256) knn_exact_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_ids,
                                                        const int* __restrict__ n_ids_dev, int64_t nq,
                                                        const float* __restrict__ ref, int64_t nref, int dim,
                                                        int64_t ld, int64_t ldr, int k, int64_t self_offset,
                                                        int64_t* __restrict__ out_idx, float* __restrict__ out_dist,
                                                        int nsplit, float* __restrict__ part_d,
                                                        int* __restrict__ part_i, int gate, int gate_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim_pad = (dim + DK - 1) / DK * DK;
  float* Qs = reinterpret_cast<float*>(smem_raw);
  float* Rs = Qs + (size_t)dim_pad * TQ;
  float* Ds = Rs + DK * (TR + 1);
  float* Ld = Ds + TQ * (TR + 1);
  int* Li = reinterpret_cast<int*>(Ld + (size_t)TQ * k);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (n_ids_dev) nq = *n_ids_dev;
  if ((gate == 1 && nq > gate_cap) || (gate == 2 && nq <= gate_cap)) return;
  const int64_t n_qtiles = (nq + TQ - 1) / TQ;
  const int64_t refs_per_split = ((nref + nsplit - 1) / nsplit + TR - 1) / TR * TR;
 for (int64_t item = blockIdx.x; item < n_qtiles * nsplit; item += gridDim.x) {
  const int64_t q0 = (item / nsplit) * TQ;
  const int sp = (int)(item % nsplit);
  const int64_t ref_begin = sp * refs_per_split, ref_end = min(ref_begin + refs_per_split, nref);
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
    if (nsplit == 1) {
      const int64_t qi = q_ids ? q_ids[q0 + tid] : q0 + tid;
      for (int p = 0; p < k; ++p) {
        out_idx[qi * k + p] = p < cnt ? (int64_t)Li[(size_t)tid * k + p] : -1;
        out_dist[qi * k + p] = p < cnt ? Ld[(size_t)tid * k + p] : FLT_MAX;
      }
    } else {
      const size_t base = ((size_t)item * TQ + tid) * k;
      for (int p = 0; p < k; ++p) {
        part_i[base + p] = p < cnt ? Li[(size_t)tid * k + p] : -1;
        part_d[base + p] = p < cnt ? Ld[(size_t)tid * k + p] : FLT_MAX;
      }
    }
  }
 }
}

// one warp per recomputed query: final k of the nsplit partial lists, ordered by (distance, index)
__global__ void __launch_bounds__(256) knn_merge_kernel(const int64_t* __restrict__ q_ids,
                                                        const int* __restrict__ n_ids_dev, int k, int nsplit,
                                                        const float* __restrict__ part_d,
                                                        const int* __restrict__ part_i, int64_t* __restrict__ out_idx,
                                                        float* __restrict__ out_dist, int gate_cap) {
  const int n = *n_ids_dev;
  if (n > gate_cap) return;
  const int lane = threadIdx.x & 31;
  for (int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < n; w += (int64_t)gridDim.x * 8) {
    const int64_t tile = w / TQ, rowi = w % TQ;
    const int64_t qi = q_ids[w];
    unsigned long long last = 0;  // keys are unique (distinct ids): strictly increasing picks
    bool first = true;
    for (int r = 0; r < k; ++r) {
      unsigned long long best = ~0ull;
      for (int c = lane; c < nsplit * k; c += 32) {
        const int sp = c / k, e = c % k;
        const size_t at = (((size_t)tile * nsplit + sp) * TQ + rowi) * k + e;
        const int j = part_i[at];
        if (j < 0) continue;
        const unsigned long long key = ((unsigned long long)__float_as_uint(part_d[at]) << 32) | (unsigned)j;
        if ((first || key > last) && key < best) best = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(SCF_FULL, best, o);
        best = other < best ? other : best;
      }
      if (lane == 0) {
        out_idx[qi * k + r] = best == ~0ull ? -1 : (int64_t)(best & 0xffffffffull);
        out_dist[qi * k + r] = best == ~0ull ? FLT_MAX : __uint_as_float((unsigned)(best >> 32));
      }
      last = best;
      first = false;
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
                                                out_idx, out_dist, 1, nullptr, nullptr, 0, 0);
  return scf_check_launch("scf_knn_l2(exact)");
}

// Recompute the rows listed in q_ids[0 .. *n_ids_dev): few rows -> reference-split work items + merge,
// many rows -> one work item per query tile.  scratch: knn_exact_fix_scratch_bytes(k).
size_t knn_exact_fix_scratch_bytes(int k) { return (size_t)KNN_FIX_CAP * KNN_FIX_NSPLIT * k * 8; }

int32_t knn_exact_fix_launch(const float* q, const int64_t* q_ids, const int* n_ids_dev, int64_t nq_max,
                             const float* ref, int64_t nref, int dim, int64_t ld, int k, int64_t self_offset,
                             int64_t* out_idx, float* out_dist, void* scratch, cudaStream_t stream) {
  if (nq_max == 0) return 0;
  size_t smem;
  int32_t rc = exact_prepare(dim, k, smem);
  if (rc) return rc;
  float* part_d = (float*)scratch;
  int* part_i = (int*)((unsigned char*)scratch + (size_t)KNN_FIX_CAP * KNN_FIX_NSPLIT * k * 4);
  knn_exact_kernel<<<2 * SCF_NUM_SMS, 256, smem, stream>>>(q, q_ids, n_ids_dev, nq_max, ref, nref, dim, ld, ld, k,
                                                           self_offset, out_idx, out_dist, KNN_FIX_NSPLIT, part_d,
                                                           part_i, 1, KNN_FIX_CAP);
  rc = scf_check_launch("scf_knn_l2(fix,split)");
  if (rc) return rc;
  knn_merge_kernel<<<SCF_NUM_SMS, 256, 0, stream>>>(q_ids, n_ids_dev, k, KNN_FIX_NSPLIT, part_d, part_i, out_idx,
                                                    out_dist, KNN_FIX_CAP);
  rc = scf_check_launch("scf_knn_l2(fix,merge)");
  if (rc) return rc;
  const int64_t tiles = (nq_max + TQ - 1) / TQ;
  const unsigned grid = (unsigned)(tiles < 4 * SCF_NUM_SMS ? tiles : 4 * SCF_NUM_SMS);
  knn_exact_kernel<<<grid, 256, smem, stream>>>(q, q_ids, n_ids_dev, nq_max, ref, nref, dim, ld, ld, k, self_offset,
                                                out_idx, out_dist, 1, nullptr, nullptr, 2, KNN_FIX_CAP);
  return scf_check_launch("scf_knn_l2(fix,full)");
}
