// K5, method 1: exact kNN = tcgen05 TF32 candidate generation + exact FP64 re-rank + proven guard band.
//
//   score(i,j) = |b_j|^2 - 2 a_i.b_j  (= d(i,j) - |a_i|^2) is produced by ONE tensor-core contraction: the
//   operands are augmented along K,   A' = [a, 1, 1, 1, 0..],  B' = [-2b, n_hi, n_mid, n_lo, 0..]  with
//   |b|^2 = n_hi + n_mid + n_lo split into TF32-exact pieces, so the epilogue has no per-element arithmetic
//   besides the top-k' filter.  Pad reference rows carry n_hi = 1e30 and never win.
//
//   kernel 1 (knn_prep)    builds A', B' (values pre-rounded to TF32, round-to-nearest), |a|^2, max|b|.
//   kernel 2 (knn_tc)      one CTA per (128-query tile, reference split): TMA -> smem (128B swizzle) ->
//                          tcgen05.mma kind::tf32 (M=128, N=256, K=8) -> TMEM (2 x 256 columns, double
//                          buffered) -> 4 epilogue warps, one query row per thread, keep the k' smallest.
//   kernel 3 (knn_rerank)  one warp per query: the oracle's FP64 distance for every candidate, order by
//                          (float32 distance, index), and the guard: everything the tensor cores rejected is
//                          provably farther than the k-th kept neighbour, else the row goes on the fail list.
//   kernel 4 (knn_exact)   FP64 brute force for the fail list (device-side count, no host round trip).
#include <float.h>
#include "common.cuh"
#include "knn_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;        // queries per CTA (= TMEM lanes)
constexpr int BN = 256;        // references per MMA / accumulator tile (TMEM columns)
constexpr int KCH = 32;        // float32 per 128-byte swizzle row
constexpr int A_CHUNK_BYTES = BM * 128;
constexpr int B_STAGE_BYTES = BN * 128;
constexpr int NTHREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue
constexpr float PAD_NORM = 1e30f;

// error model of the tensor-core score (see DESIGN.md "kNN guard band"):
//   inputs rounded to TF32 (RN, rel 2^-11 each) -> |2 a.b - 2 a~.b~| <= 2 (2*2^-11 + 2^-22) |a||b|
//   FP32 accumulation of <= 131 terms, any order, truncating adder (2^-23 per add, x4 margin)
//   -> 2^-14 (2|a||b| + |b|^2)
__device__ __host__ inline double eps_c1() { return 2.0 * (2.0 * 0x1p-11 + 0x1p-22) + 0x1p-13; }
__device__ __host__ inline double eps_c2() { return 0x1p-14 + 0x1p-28; }

__device__ __forceinline__ float to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------------------------------------- prep
// one warp per row; rows [0, nq_pad) of qop and [0, nr_pad) of rop
__global__ void __launch_bounds__(256) knn_prep_kernel(const float* __restrict__ q, int64_t nq, int64_t nq_pad,
                                                       const float* __restrict__ ref, int64_t nref, int64_t nr_pad,
                                                       int dim, int64_t ld, int kp, float* __restrict__ qop,
                                                       float* __restrict__ rop, double* __restrict__ qnorm2,
                                                       float* __restrict__ bmax) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  float local_bmax = 0.f;
  for (int64_t r = w0; r < nq_pad + nr_pad; r += nw) {
    const bool is_q = r < nq_pad;
    const int64_t row = is_q ? r : r - nq_pad;
    const bool live = is_q ? row < nq : row < nref;
    const float* src = (is_q ? q : ref) + row * ld;
    float* dst = (is_q ? qop : rop) + row * kp;
    double n2 = 0.0;
    for (int t = lane; t < kp; t += 32) {
      if (t >= dim && t < dim + 3) continue;  // the three augmentation slots are written below
      float v = 0.f;
      if (live && t < dim) {
        const float x = src[t];
        n2 += (double)x * (double)x;
        v = is_q ? to_tf32_rn(x) : -2.f * to_tf32_rn(x);
      }
      dst[t] = v;
    }
    n2 = warp_sum(n2);
    if (lane < 3) {
      float aug = 0.f;
      if (is_q) {
        aug = live ? 1.f : 0.f;
      } else if (live) {
        const float hi = tf32_trunc((float)n2);
        const float mid = tf32_trunc((float)(n2 - (double)hi));
        const float lo = tf32_trunc((float)(n2 - (double)hi - (double)mid));
        aug = lane == 0 ? hi : (lane == 1 ? mid : lo);
      } else {
        aug = lane == 0 ? PAD_NORM : 0.f;
      }
      dst[dim + lane] = aug;
    }
    if (lane == 0 && live) {
      if (is_q)
        qnorm2[row] = n2;
      else
        local_bmax = fmaxf(local_bmax, (float)sqrt(n2) * 1.0000002f);
    }
  }
  if (lane == 0 && local_bmax > 0.f) atomicMax(reinterpret_cast<int*>(bmax), __float_as_int(local_bmax));
}

// ---------------------------------------------------------------------------------------------- main
// Per-thread UNSORTED candidate list in shared memory, entry e of row r at [e * BM + r] (conflict free).
// thr = largest kept score, pmax = its slot: a new candidate overwrites that slot, then the maximum is
// recomputed with kc independent loads (no dependent insertion-sort chain).
__device__ __forceinline__ void list_replace_max(float* ls, int* li, int row, int kc, float s, int j, float& thr,
                                                 int& pmax) {
  ls[pmax * BM + row] = s;
  li[pmax * BM + row] = j;
  float mx = ls[row];
  int pm = 0;
#pragma unroll 4
  for (int e = 1; e < kc; ++e) {
    const float x = ls[e * BM + row];
    if (x > mx) mx = x, pm = e;
  }
  thr = mx;
  pmax = pm;
}

// v[c] for a runtime c in [0,32): 31 selects instead of a local-memory round trip
__device__ __forceinline__ float select32(const uint32_t (&v)[32], int c) {
  uint32_t a[16], b[8], d[4], e[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (c & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (c & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = (c & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) e[i] = (c & 8) ? d[2 * i + 1] : d[2 * i];
  return __uint_as_float((c & 16) ? e[1] : e[0]);
}

// One 32-column chunk of one query row: block-min prefilter, then every lane drains its own hits concurrently.
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], int jbase, float* ls, int* li, int row, int kc,
                                           float& thr, int& pmax) {
  float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]),
        m3 = __uint_as_float(v[3]);
#pragma unroll
  for (int c = 4; c < 32; c += 4) {
    m0 = fminf(m0, __uint_as_float(v[c]));
    m1 = fminf(m1, __uint_as_float(v[c + 1]));
    m2 = fminf(m2, __uint_as_float(v[c + 2]));
    m3 = fminf(m3, __uint_as_float(v[c + 3]));
  }
  const float m = fminf(fminf(m0, m1), fminf(m2, m3));
  if (!__any_sync(SCF_FULL, m < thr)) return;  // warp-uniform: no candidate in this chunk for any row
  uint32_t mask = 0;
  if (m < thr) {
#pragma unroll
    for (int c = 0; c < 32; ++c) mask |= (__uint_as_float(v[c]) < thr) ? (1u << c) : 0u;
  }
  while (__any_sync(SCF_FULL, mask != 0u)) {
    if (mask) {
      const int c = __ffs(mask) - 1;
      mask &= mask - 1;
      const float s = select32(v, c);
      if (s < thr) list_replace_max(ls, li, row, kc, s, jbase + c, thr, pmax);
    }
  }
}

struct KnnTcParams {
  int kchunks;          // Kp / 32
  int stages;           // B pipeline depth
  int kc;               // candidates kept per (query, split)
  int n_ref_tiles;      // nr_pad / BN
  int tiles_per_split;
  int nsplit;
  float* cand_score;    // [nq_pad, nsplit, kc]  (unsorted)
  int* cand_idx;
  float* cand_tau;      // [nq_pad, nsplit]  largest kept score = lower bound of every rejected score
};

__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                             const __grid_constant__ CUtensorMap tmap_r,
                                                             const KnnTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the 128-byte swizzle is a function of the shared-memory address: tiles must sit on 1024-byte boundaries
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = sA + (size_t)p.kchunks * A_CHUNK_BYTES;
  float* ls = reinterpret_cast<float*>(sB + (size_t)p.stages * B_STAGE_BYTES);
  int* li = reinterpret_cast<int*>(ls + (size_t)p.kc * BM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(li + (size_t)p.kc * BM);
  uint64_t* a_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int split = blockIdx.y;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(tile_begin + p.tiles_per_split, p.n_ref_tiles);
  const int ntiles = max(tile_end - tile_begin, 0);

  if (threadIdx.x == 0) {
    tc::mbar_init(a_full, 1);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tmem_full + a, 1);
      tc::mbar_init(tmem_empty + a, BM);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_q);
    tc::tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_expect_tx(a_full, (uint32_t)(p.kchunks * A_CHUNK_BYTES));
      for (int c = 0; c < p.kchunks; ++c) tc::tma_load_2d(sA + (size_t)c * A_CHUNK_BYTES, &tmap_q, a_full, c * KCH, q0);
      int it = 0;
      for (int t = tile_begin; t < tile_end; ++t)
        for (int c = 0; c < p.kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          tc::mbar_wait(empty + s, ph ^ 1u);
          tc::mbar_expect_tx(full + s, B_STAGE_BYTES);
          tc::tma_load_2d(sB + (size_t)s * B_STAGE_BYTES, &tmap_r, full + s, c * KCH, t * BN);
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(BM, BN, false, false);
      tc::mbar_wait(a_full, 0);
      tc::tc_fence_after();
      int it = 0;
      for (int lt = 0; lt < ntiles; ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_ph = (uint32_t)(lt >> 1) & 1u;
        tc::mbar_wait(tmem_empty + acc, acc_ph ^ 1u);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int c = 0; c < p.kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          tc::mbar_wait(full + s, ph);
          tc::tc_fence_after();
          const uint64_t da = tc::umma_desc_k_sw128(sA + (size_t)c * A_CHUNK_BYTES);
          const uint64_t db = tc::umma_desc_k_sw128(sB + (size_t)s * B_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < KCH / 8; ++kk)  // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
            tc::umma_tf32(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (c | kk) != 0);
          tc::umma_commit(empty + s);  // smem stage reusable once these MMAs have read it
        }
        tc::umma_commit(tmem_full + acc);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: top-k' per query row =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    for (int e = 0; e < p.kc; ++e) {
      ls[e * BM + row] = FLT_MAX;
      li[e * BM + row] = -1;
    }
    float thr = FLT_MAX;
    int pmax = 0;
    for (int lt = 0; lt < ntiles; ++lt) {
      const int acc = lt & 1;
      const uint32_t acc_ph = (uint32_t)(lt >> 1) & 1u;
      tc::mbar_wait(tmem_full + acc, acc_ph);
      tc::tc_fence_after();
      const int j0 = (tile_begin + lt) * BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      uint32_t v0[32], v1[32];
      tc::tmem_ld32(t_row, v0);
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; cc += 2) {  // TMEM load of the next chunk overlaps the scan of this one
        tc::tmem_ld_wait();
        tc::tmem_ld32(t_row + (uint32_t)((cc + 1) * 32), v1);
        scan_chunk(v0, j0 + cc * 32, ls, li, row, p.kc, thr, pmax);
        tc::tmem_ld_wait();
        if (cc + 2 < BN / 32) tc::tmem_ld32(t_row + (uint32_t)((cc + 2) * 32), v0);
        scan_chunk(v1, j0 + (cc + 1) * 32, ls, li, row, p.kc, thr, pmax);
      }
      tc::tc_fence_before();
      tc::mbar_arrive(tmem_empty + acc);
    }
    // candidates out: [query, split, kc]
    const size_t base = ((size_t)(q0 + row) * p.nsplit + split) * p.kc;
    for (int e = 0; e < p.kc; ++e) {
      p.cand_score[base + e] = ls[e * BM + row];
      p.cand_idx[base + e] = li[e * BM + row];
    }
    p.cand_tau[(size_t)(q0 + row) * p.nsplit + split] = thr;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- re-rank
constexpr int MAXU = 8;  // candidates per lane: nsplit * kc <= 256

__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ q, int64_t nq,
                                                         const float* __restrict__ ref, int64_t nref, int dim,
                                                         int64_t ld, int k, int64_t self_offset, int kc, int nsplit,
                                                         const float* __restrict__ cand_score,
                                                         const int* __restrict__ cand_idx,
                                                         const float* __restrict__ cand_tau,
                                                         const double* __restrict__ qnorm2,
                                                         const float* __restrict__ bmax, int64_t* __restrict__ out_idx,
                                                         float* __restrict__ out_dist, int64_t* __restrict__ fail_ids,
                                                         int* __restrict__ fail_count) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const int ncand = kc * nsplit;
  const int64_t self = self_offset >= 0 ? qi + self_offset : -1;
  const float* a = q + qi * ld;
  unsigned long long key[MAXU];
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    key[u] = ~0ull;
    const int c = lane + 32 * u;
    if (c < ncand) {
      const int j = cand_idx[(size_t)qi * ncand + c];
      if (j >= 0 && j < nref && j != self) {
        const float* b = ref + (int64_t)j * ld;
        double acc = 0.0;
        for (int t = 0; t < dim; ++t) {  // the oracle's arithmetic: sequential, separate multiply and add
          const double df = __dsub_rn((double)a[t], (double)__ldg(b + t));
          acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
        key[u] = ((unsigned long long)__float_as_uint((float)acc) << 32) | (unsigned)j;
      }
    }
  }
  // smallest kept-list threshold over the splits: every rejected reference scored >= tau
  float tau = FLT_MAX;
  for (int s = lane; s < nsplit; s += 32) tau = fminf(tau, cand_tau[(size_t)qi * nsplit + s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tau = fminf(tau, __shfl_xor_sync(SCF_FULL, tau, o));
  unsigned long long last = ~0ull;
  bool enough = true;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = key[0];
#pragma unroll
    for (int u = 1; u < MAXU; ++u) best = best < key[u] ? best : key[u];
    unsigned long long wbest = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(SCF_FULL, wbest, o);
      wbest = other < wbest ? other : wbest;
    }
    if (wbest == ~0ull) {
      enough = false;
      break;
    }
#pragma unroll
    for (int u = 0; u < MAXU; ++u)
      if (key[u] == wbest) key[u] = ~0ull;  // ids are unique, so exactly one slot matches
    if (lane == 0) {
      out_idx[qi * k + r] = (int64_t)(wbest & 0xffffffffull);
      out_dist[qi * k + r] = __uint_as_float((unsigned)(wbest >> 32));
    }
    last = wbest;
  }
  if (lane == 0) {
    bool ok = enough;
    if (ok && tau < 1e29f) {  // lists were full: something was rejected, prove it is farther than the k-th kept
      const double an2 = qnorm2[qi], bm = (double)*bmax;
      const double eps = eps_c1() * sqrt(an2) * bm + eps_c2() * bm * bm;
      const double lower = ((double)tau - eps + an2) * (1.0 - 1e-12);  // bound on any rejected exact distance
      const float dk = __uint_as_float((unsigned)(last >> 32));
      ok = lower > 0.0 && dk < __double2float_rd(lower);  // strict: a tie would be decided by the index
    }
    if (!ok) fail_ids[atomicAdd(fail_count, 1)] = qi;
  }
}

int pick_kc(int k) {
  const int need = k + 1;  // the query itself may be among the candidates
  if (need <= 12) return 16;
  if (need <= 25) return 32;
  if (need <= 50) return 64;
  return 0;
}

struct Plan {
  int kp, kchunks, kc, stages, nsplit, tiles_per_split, n_ref_tiles;
  int64_t nq_pad, nr_pad;
  size_t smem;
  // workspace offsets (bytes)
  size_t off_qop, off_rop, off_qn, off_cs, off_ci, off_tau, off_fail, off_misc, total;
};

bool make_plan(int64_t nq, int64_t nref, int dim, int k, Plan& pl) {
  pl.kc = pick_kc(k);
  pl.kp = (dim + 3 + KCH - 1) / KCH * KCH;
  pl.kchunks = pl.kp / KCH;
  if (pl.kc == 0 || pl.kchunks > 4) return false;  // k > 48 or dim > 125: method 0 handles those
  pl.nq_pad = (nq + BM - 1) / BM * BM;
  pl.nr_pad = (nref + BN - 1) / BN * BN;
  pl.n_ref_tiles = (int)(pl.nr_pad / BN);
  const int64_t qtiles = pl.nq_pad / BM;
  // reference splits: fill the 148 SMs evenly (1 CTA per SM), keep >= 4 tiles per split, <= 8 splits
  int best_s = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 8 && s * 4 <= std::max(pl.n_ref_tiles, 4) && s * pl.kc <= 32 * MAXU; ++s) {
    const int64_t ctas = qtiles * s;
    const double eff = (double)ctas / (double)((ctas + SCF_NUM_SMS - 1) / SCF_NUM_SMS * SCF_NUM_SMS);
    if (eff > best_eff + 0.03) best_eff = eff, best_s = s;
  }
  pl.nsplit = best_s;
  pl.tiles_per_split = (pl.n_ref_tiles + pl.nsplit - 1) / pl.nsplit;
  pl.nsplit = (pl.n_ref_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
  auto smem_for = [&](int stages) {
    return (size_t)pl.kchunks * A_CHUNK_BYTES + (size_t)stages * B_STAGE_BYTES + (size_t)pl.kc * BM * 8 +
           (size_t)(1 + 2 * stages + 4) * 8 + 16;
  };
  pl.stages = 4;
  while (pl.stages > 2 && smem_for(pl.stages) + 1024 > 227 * 1024) --pl.stages;
  pl.smem = smem_for(pl.stages) + 1024;  // slack for the 1024-byte alignment of the swizzled tiles
  if (pl.smem > 227 * 1024) return false;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  pl.off_qop = o, o = al(o + (size_t)pl.nq_pad * pl.kp * 4);
  pl.off_rop = o, o = al(o + (size_t)pl.nr_pad * pl.kp * 4);
  pl.off_qn = o, o = al(o + (size_t)pl.nq_pad * 8);
  pl.off_cs = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * pl.kc * 4);
  pl.off_ci = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * pl.kc * 4);
  pl.off_tau = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * 4);
  pl.off_fail = o, o = al(o + (size_t)nq * 8);
  pl.off_misc = o, o = al(o + 256);
  pl.total = o;
  return true;
}

}  // namespace

int64_t knn_tc_fail_count_offset(int64_t nq, int64_t nref, int dim, int k) {
  Plan pl;
  if (!make_plan(nq, nref, dim, k, pl)) return -1;
  return (int64_t)pl.off_misc;
}

int64_t knn_tc_workspace_bytes(int64_t nq, int64_t nref, int dim, int k) {
  Plan pl;
  if (!make_plan(nq, nref, dim, k, pl)) return 0;
  return (int64_t)pl.total;
}

int32_t knn_tc_launch(const float* q, int64_t nq, const float* ref, int64_t nref, int dim, int64_t ld, int k,
                      int64_t self_offset, int64_t* out_idx, float* out_dist, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream) {
  Plan pl;
  if (!make_plan(nq, nref, dim, k, pl))  // k > 48 or dim > 125: outside the tensor-core kernel's shapes
    return knn_exact_launch(q, nullptr, nullptr, nq, ref, nref, dim, ld, ld, k, self_offset, out_idx, out_dist,
                            stream);
  if (!workspace || workspace_bytes < (int64_t)pl.total) {
    scf_set_error("scf_knn_l2: workspace too small (%lld < %zu)", (long long)workspace_bytes, pl.total);
    return 1;
  }
  unsigned char* ws = (unsigned char*)workspace;
  float* qop = (float*)(ws + pl.off_qop);
  float* rop = (float*)(ws + pl.off_rop);
  double* qn = (double*)(ws + pl.off_qn);
  float* cs = (float*)(ws + pl.off_cs);
  int* ci = (int*)(ws + pl.off_ci);
  float* ctau = (float*)(ws + pl.off_tau);
  int64_t* fail_ids = (int64_t*)(ws + pl.off_fail);
  int* fail_count = (int*)(ws + pl.off_misc);
  float* bmax = (float*)(ws + pl.off_misc + 64);
  cudaError_t e = cudaMemsetAsync(ws + pl.off_misc, 0, 256, stream);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  knn_prep_kernel<<<4 * SCF_NUM_SMS, 256, 0, stream>>>(q, nq, pl.nq_pad, ref, nref, pl.nr_pad, dim, ld, pl.kp, qop, rop,
                                                       qn, bmax);
  int32_t rc = scf_check_launch("scf_knn_l2(prep)");
  if (rc) return rc;
  CUtensorMap tq, tr;
  rc = scf_make_tmap_2d_f32(&tq, qop, (uint64_t)pl.nq_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BM);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f32(&tr, rop, (uint64_t)pl.nr_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BN);
  if (rc) return rc;
  KnnTcParams prm;
  prm.kchunks = pl.kchunks, prm.stages = pl.stages, prm.kc = pl.kc, prm.n_ref_tiles = pl.n_ref_tiles;
  prm.tiles_per_split = pl.tiles_per_split, prm.nsplit = pl.nsplit, prm.cand_score = cs, prm.cand_idx = ci, prm.cand_tau = ctau;
  e = cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  dim3 grid((unsigned)(pl.nq_pad / BM), (unsigned)pl.nsplit);
  knn_tc_kernel<<<grid, NTHREADS, pl.smem, stream>>>(tq, tr, prm);
  rc = scf_check_launch("scf_knn_l2(tcgen05)");
  if (rc) return rc;
  knn_rerank_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(q, nq, ref, nref, dim, ld, k, self_offset, pl.kc,
                                                                  pl.nsplit, cs, ci, ctau, qn, bmax, out_idx, out_dist,
                                                                  fail_ids, fail_count);
  rc = scf_check_launch("scf_knn_l2(rerank)");
  if (rc) return rc;
  return knn_exact_launch(q, fail_ids, fail_count, nq, ref, nref, dim, ld, ld, k, self_offset, out_idx, out_dist,
                          stream);
}
