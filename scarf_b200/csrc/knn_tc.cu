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
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "knn_common.cuh"
#include "tc_common.cuh"

namespace {

#ifndef SCF_KNN_DEBUG
#define SCF_KNN_DEBUG 0  // compile-time developer switches for the epilogue: 4 / 8 = timing experiments, 16 = counters
#endif
__device__ unsigned long long g_dbg[8];  // SCF_KNN_DEBUG & 16: event counters (developer diagnostics)

constexpr int BM = 128;        // queries per MMA (= TMEM lanes)
constexpr int QT = 2;          // query tiles per CTA: both reuse every reference tile staged in shared memory
constexpr int BN = 128;        // references per MMA / accumulator tile (TMEM columns)
constexpr int KCH = 32;        // float32 per 128-byte swizzle row
constexpr int A_CHUNK_BYTES = BM * 128;
constexpr int B_STAGE_BYTES = BN * 128;
constexpr int NTHREADS = 320;  // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-9: epilogue
constexpr int NLISTS = QT * BM;  // one candidate list per query row of the CTA
constexpr float PAD_NORM = 1e30f;

// error model of the tensor-core score (see DESIGN.md "kNN guard band"), a~ = tf32(a), da = a - a~ (known exactly):
//   |2 a.b - 2 a~.b~| = 2 |da.b + a~.db| <= 2 (|da| |b| + |a| |db|)      with |da| per query, max |b|, max |db|
//   tf32 x tf32 products are exact in FP32; accumulation of <= 131 terms in any order with a truncating adder
//   (2^-23 per add, x4 margin) -> 2^-14 (2|a||b| + |b|^2); the |b|^2 split leaves < 2^-28 |b|^2
__device__ __host__ inline double eps_acc() { return 0x1p-14 + 0x1p-28; }

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
                                                       float* __restrict__ qerr, float* __restrict__ bmax) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  float local_bmax = 0.f, local_dbmax = 0.f;  // bmax[0] = max |b|, bmax[1] = max |b - tf32(b)|
  for (int64_t r = w0; r < nq_pad + nr_pad; r += nw) {
    const bool is_q = r < nq_pad;
    const int64_t row = is_q ? r : r - nq_pad;
    const bool live = is_q ? row < nq : row < nref;
    const float* src = (is_q ? q : ref) + row * ld;
    float* dst = (is_q ? qop : rop) + row * kp;
    double n2 = 0.0, e2 = 0.0;
    for (int t = lane; t < kp; t += 32) {
      if (t >= dim && t < dim + 3) continue;  // the three augmentation slots are written below
      float v = 0.f;
      if (live && t < dim) {
        const float x = src[t];
        const float xr = to_tf32_rn(x);
        const double dx = (double)x - (double)xr;
        n2 += (double)x * (double)x;
        e2 += dx * dx;
        v = is_q ? xr : -2.f * xr;
      }
      dst[t] = v;
    }
    n2 = warp_sum(n2);
    e2 = warp_sum(e2);
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
      if (is_q) {
        qnorm2[row] = n2;
        qerr[row] = (float)sqrt(e2) * 1.0000002f;
      } else {
        local_bmax = fmaxf(local_bmax, (float)sqrt(n2) * 1.0000002f);
        local_dbmax = fmaxf(local_dbmax, (float)sqrt(e2) * 1.0000002f);
      }
    }
  }
  if (lane == 0 && local_bmax > 0.f) atomicMax(reinterpret_cast<int*>(bmax), __float_as_int(local_bmax));
  if (lane == 0 && local_dbmax > 0.f) atomicMax(reinterpret_cast<int*>(bmax + 1), __float_as_int(local_dbmax));
}

// ---------------------------------------------------------------------------------------------- main
// Per-thread UNSORTED candidate list: the KC scores live in registers, the ids in shared memory (entry e of
// list l at [e * NLISTS + l], conflict free).  The slot number is carried in the low mantissa bits of every kept
// score (<= 31 ulp perturbation; the threshold is rounded down past the tag, so every rejected score is >= thr),
// so one FMNMX3 tree yields both the largest kept score and the slot that holds it: a new candidate overwrites it.
template <int KC>
struct CandList {
  static constexpr uint32_t SLOT_MASK = KC - 1;
  float r[KC];
  float tmax;  // largest kept (tagged) score: its low bits name the slot
  float thr;   // comparison threshold: tmax with the tag rounded DOWN, so scores equal to a kept one never pass
  template <int LO, int N>
  __device__ __forceinline__ float max_tree() const {  // ternary tree over r[LO .. LO+N) -> FMNMX3
    if constexpr (N == 1) {
      return r[LO];
    } else if constexpr (N == 2) {
      return fmaxf(r[LO], r[LO + 1]);
    } else {
      constexpr int A = (N + 2) / 3, B = (N - A + 1) / 2, C = N - A - B;
      return fmaxf(fmaxf(max_tree<LO, A>(), max_tree<LO + A, B>()), max_tree<LO + A + B, C>());
    }
  }
  __device__ __forceinline__ void recompute() {
    tmax = max_tree<0, KC>();
    const uint32_t b = __float_as_uint(tmax);
    thr = __uint_as_float((b & 0x80000000u) ? (b | SLOT_MASK) : (b & ~SLOT_MASK));
  }
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < KC; ++e) r[e] = __uint_as_float((0x7F7FFFFFu & ~SLOT_MASK) | (uint32_t)e);
    recompute();
  }
  __device__ __forceinline__ void replace_max(float s, int j, int* li_slot) {
    const uint32_t pmax = __float_as_uint(tmax) & SLOT_MASK;
    li_slot[pmax * NLISTS] = j;
    const float tagged = __uint_as_float((__float_as_uint(s) & ~SLOT_MASK) | pmax);
    const uint32_t onehot = 1u << pmax;  // selects, not a dynamically indexed store (keeps r[] in registers)
#pragma unroll
    for (int e = 0; e < KC; ++e) r[e] = (onehot & (1u << e)) ? tagged : r[e];
    recompute();
  }
};

__device__ __forceinline__ float min8(const uint32_t* v) {
  const float a = fminf(fminf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
  const float b = fminf(fminf(__uint_as_float(v[3]), __uint_as_float(v[4])), __uint_as_float(v[5]));
  return fminf(fminf(a, b), fminf(__uint_as_float(v[6]), __uint_as_float(v[7])));
}

// One 32-column chunk of one query row.  Group-of-8 minima prefilter (FMNMX3); rows with a candidate stage the
// values of their hit groups in shared memory ([column][thread], conflict free) and set a 32-bit hit mask; then
// every lane drains its own hits concurrently from the staged copy (dynamic column index without local memory).
// All branches that contain warp votes are warp-uniform.  Returns after the registers v are dead, so the caller
// can issue the next TMEM load before calling drain_hits().
template <int KC>
__device__ __forceinline__ uint32_t stage_hits(const uint32_t (&v)[32], float thr, float* st_slot, int dbg) {
  if (dbg & 8) return 0u;  // timing experiment: TMEM traffic only
  float g[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = min8(v + 8 * i);
  const float m = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
  if (!__any_sync(SCF_FULL, m < thr) || (dbg & 4)) return 0u;  // no candidate in this chunk for any row of the warp
  const bool cnt = (dbg & 16) && (threadIdx.x & 31) == 0;
  if (cnt) atomicAdd(&g_dbg[0], 1ull);
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!__any_sync(SCF_FULL, g[i] < thr)) continue;
    if (cnt) atomicAdd(&g_dbg[1], 1ull);
    if (g[i] < thr) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float x = __uint_as_float(v[8 * i + c]);
        if (x < thr) {
          st_slot[(8 * i + c) * NLISTS] = x;
          mask |= 1u << (8 * i + c);
        }
      }
    }
  }
  return mask;
}

template <int KC>
__device__ __forceinline__ void drain_hits(uint32_t mask, int jbase, CandList<KC>& cl, int* li_slot,
                                           const float* st_slot, int dbg) {
  const bool cnt = (dbg & 16) && (threadIdx.x & 31) == 0;
  while (__any_sync(SCF_FULL, mask != 0u)) {
    if (cnt) atomicAdd(&g_dbg[2], 1ull);
    if (mask) {
      const int c = __ffs(mask) - 1;
      mask &= mask - 1;
      const float sc = st_slot[c * NLISTS];
      if (sc < cl.thr) {
        if (dbg & 16) atomicAdd(&g_dbg[3], 1ull);
        cl.replace_max(sc, jbase + c, li_slot);
      }
    }
  }
}

struct KnnTcParams {
  int kchunks;          // Kp / 32
  int ksteps_last;      // 8-wide K steps that hold data in the last chunk (all-zero padding steps are skipped)
  int stages;           // B pipeline depth
  int n_ref_tiles;      // nr_pad / BN
  int tiles_per_split;
  int nsplit;
  float* cand_score;    // [nq_pad, nsplit, KC]  (unsorted)
  int* cand_idx;
  float* cand_tau;      // [nq_pad, nsplit]  largest kept score = lower bound of every rejected score
  int nq;               // live query rows (pad rows keep nothing)
  int flags;            // developer switches (SCF_KNN_FLAGS): 2 = back off in waits, 4/8 = timing experiments, 16 = counters
  // collect pass (repair of guard failures): query row r of this launch is failed row r of the fail list
  const int* fail_count;   // device count of failed rows (rows >= min(count, FIXTC_ROWS) do nothing)
  const float* fix_thr;    // [FIXTC_ROWS] score threshold: every reference scoring below it is collected
  int* fix_cnt;            // [FIXTC_ROWS] collected so far
  int* fix_list;           // [FIXTC_ROWS][FIXTC_CAP] reference ids
};

constexpr int FIXTC_ROWS = 8192;  // failed rows repaired by the tensor-core collect pass (the rest: FP64 scan)
constexpr int FIXTC_CAP = 64;     // collected references per failed row
constexpr int FIXTC_NSPLIT = 16;  // reference ranges per query tile in the collect pass

template <int KC, bool COLLECT>
__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                             const __grid_constant__ CUtensorMap tmap_r,
                                                             const KnnTcParams p) {
  int n_fix = 0;
  if constexpr (COLLECT) {  // whole CTA leaves before any barrier / TMEM setup when its query tile is empty
    n_fix = min(*p.fail_count, FIXTC_ROWS);
    if ((int)blockIdx.x * (QT * BM) >= n_fix) return;
  }
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the 128-byte swizzle is a function of the shared-memory address: tiles must sit on 1024-byte boundaries
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                             // [QT][kchunks][128 rows x 128 B]
  unsigned char* sB = sA + (size_t)QT * p.kchunks * A_CHUNK_BYTES;      // [stages][128 rows x 128 B]
  int* li = reinterpret_cast<int*>(sB + (size_t)p.stages * B_STAGE_BYTES);  // candidate ids [KC][NLISTS]
  float* st = reinterpret_cast<float*>(li + (size_t)KC * NLISTS);           // staged chunk values [32][NLISTS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(st + (size_t)32 * NLISTS);
  uint64_t* a_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t backoff = (p.flags & 2) ? 32u : 0u;
  const int q0 = blockIdx.x * (QT * BM);
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
      tc::mbar_init(tmem_empty + a, NLISTS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_q);
    tc::tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);  // 2 buffers x QT tiles x 128 columns
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_expect_tx(a_full, (uint32_t)(QT * p.kchunks * A_CHUNK_BYTES));
      for (int t = 0; t < QT; ++t)
        for (int c = 0; c < p.kchunks; ++c)
          tc::tma_load_2d(sA + (size_t)(t * p.kchunks + c) * A_CHUNK_BYTES, &tmap_q, a_full, c * KCH, q0 + t * BM);
      int it = 0;
      for (int t = tile_begin; t < tile_end; ++t)
        for (int c = 0; c < p.kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          tc::mbar_wait(empty + s, ph ^ 1u, backoff);
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
        tc::mbar_wait(tmem_empty + acc, acc_ph ^ 1u, backoff);
        tc::tc_fence_after();
        for (int c = 0; c < p.kchunks; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          tc::mbar_wait(full + s, ph, backoff);
          tc::tc_fence_after();
          const uint64_t db = tc::umma_desc_k_sw128(sB + (size_t)s * B_STAGE_BYTES);
          const int nk = c + 1 == p.kchunks ? p.ksteps_last : KCH / 8;
#pragma unroll
          for (int t = 0; t < QT; ++t) {
            const uint64_t da = tc::umma_desc_k_sw128(sA + (size_t)(t * p.kchunks + c) * A_CHUNK_BYTES);
            const uint32_t d_tmem = tmem_base + (uint32_t)((acc * QT + t) * BN);
#pragma unroll
            for (int kk = 0; kk < KCH / 8; ++kk)  // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
              if (kk < nk) tc::umma_tf32(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (c | kk) != 0);
          }
          tc::umma_commit(empty + s);  // smem stage reusable once these MMAs have read it
        }
        tc::umma_commit(tmem_full + acc);  // both accumulators of this reference tile complete
      }
    }
  } else {
    // ===================== epilogue: one query row per thread, top-k' of the whole reference range ===========
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int qt = (warp - 2) >> 2;    // which of the CTA's query tiles
    const int row = quarter * 32 + lane;
    int* li_slot = li + qt * BM + row;
    float* st_slot = st + qt * BM + row;
    if constexpr (COLLECT) {
      // fixed per-row threshold, append-only: every reference whose score is below it goes on the row's list
      const int slot = q0 + qt * BM + row;
      const float thr = slot < n_fix ? p.fix_thr[slot] : -FLT_MAX;
      for (int lt = 0; lt < ntiles; ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_ph = (uint32_t)(lt >> 1) & 1u;
        tc::mbar_wait(tmem_full + acc, acc_ph);
        tc::tc_fence_after();
        const int j0 = (tile_begin + lt) * BN;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * QT + qt) * BN);
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t v[32];
          tc::tmem_ld32(t_row + (uint32_t)(cc * 32), v);
          tc::tmem_ld_wait();
          float g[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) g[i] = min8(v + 8 * i);
          const float m = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
          if (!__any_sync(SCF_FULL, m < thr)) continue;
          if (m < thr) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (__uint_as_float(v[c]) < thr) {
                const int pos = atomicAdd(p.fix_cnt + slot, 1);
                if (pos < FIXTC_CAP) p.fix_list[(size_t)slot * FIXTC_CAP + pos] = j0 + cc * 32 + c;
              }
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(tmem_empty + acc);
      }
    } else {
    CandList<KC> cl;
      cl.init();
      if (q0 + qt * BM + row >= p.nq) cl.thr = -FLT_MAX;  // pad row: never a candidate
  #pragma unroll
      for (int e = 0; e < KC; ++e) li_slot[e * NLISTS] = -1;
      for (int lt = 0; lt < ntiles; ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_ph = (uint32_t)(lt >> 1) & 1u;
        tc::mbar_wait(tmem_full + acc, acc_ph);
        tc::tc_fence_after();
        const int j0 = (tile_begin + lt) * BN;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * QT + qt) * BN);
        uint32_t v[32];
        tc::tmem_ld32(t_row, v);
  #pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          tc::tmem_ld_wait();
          const uint32_t mask = stage_hits<KC>(v, cl.thr, st_slot, SCF_KNN_DEBUG);
          // v is dead: the TMEM load of the next chunk overlaps the drain of this one
          if (cc + 1 < BN / 32) tc::tmem_ld32(t_row + (uint32_t)((cc + 1) * 32), v);
          drain_hits<KC>(mask, j0 + cc * 32, cl, li_slot, st_slot, SCF_KNN_DEBUG);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(tmem_empty + acc);
      }
      // candidates out: [query, split, KC]
      const size_t sub = (size_t)(q0 + qt * BM + row) * p.nsplit + split;
  #pragma unroll
      for (int e = 0; e < KC; ++e) {
        p.cand_score[sub * KC + e] = cl.r[e];
        p.cand_idx[sub * KC + e] = li_slot[e * NLISTS];
      }
      p.cand_tau[sub] = cl.thr;
    }
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
                                                         const float* __restrict__ qerr,
                                                         const float* __restrict__ bmax, int64_t* __restrict__ out_idx,
                                                         float* __restrict__ out_dist, int64_t* __restrict__ fail_ids,
                                                         unsigned long long* __restrict__ fail_keys,
                                                         float* __restrict__ fix_thr, int* __restrict__ fail_count) {
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
    const double an2 = qnorm2[qi], an = sqrt(an2), bm = (double)bmax[0], dbm = (double)bmax[1];
    const double eps = 2.0 * ((double)qerr[qi] * bm + an * dbm) * (1.0 + 1e-6) + eps_acc() * (2.0 * an * bm + bm * bm);
    const float dk = __uint_as_float((unsigned)(last >> 32));
    if (ok && tau < 1e29f) {  // lists were full: something was rejected, prove it is farther than the k-th kept
      const double lower = ((double)tau - eps + an2) * (1.0 - 1e-12);  // bound on any rejected exact distance
      ok = lower > 0.0 && dk < __double2float_rd(lower);  // strict: a tie would be decided by the index
    }
    if (!ok) {
      // Repair: every true neighbour has an exact distance <= dk, hence a tensor-core score below
      // dk - |a|^2 + eps; the collect pass gathers exactly those references (or, if there were fewer than k
      // candidates, the FP64 scan takes the row).
      const int slot = atomicAdd(fail_count, 1);
      fail_ids[slot] = qi;
      fail_keys[slot] = enough ? last : ~0ull;
      if (slot < FIXTC_ROWS) {
        const double t = (double)dk * (1.0 + 0x1p-22) - an2 + eps;
        fix_thr[slot] = enough ? __double2float_ru(t + fabs(t) * 1e-9) : FLT_MAX;
      }
    }
  }
}

// rows of the augmented query operand of the failed queries, compacted for the collect pass
__global__ void __launch_bounds__(256) knn_fix_gather_kernel(const float* __restrict__ qop, int kp,
                                                             const int64_t* __restrict__ fail_ids,
                                                             const int* __restrict__ fail_count,
                                                             float* __restrict__ qfix) {
  const int n = min(*fail_count, FIXTC_ROWS);
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    const float* src = qop + fail_ids[r] * kp;
    for (int t = lane; t < kp; t += 32) qfix[(size_t)r * kp + t] = src[t];
  }
}

// one warp per failed row: the oracle's FP64 distance of every collected reference, k smallest by (distance, index);
// rows that cannot be finished here (list overflow, too few entries, beyond the collect capacity) are handed on
__global__ void __launch_bounds__(256) knn_fix_finish_kernel(const float* __restrict__ q, const float* __restrict__ ref,
                                                             int dim, int64_t ld, int k, int64_t self_offset,
                                                             const int64_t* __restrict__ fail_ids,
                                                             const unsigned long long* __restrict__ fail_keys,
                                                             const int* __restrict__ fail_count,
                                                             const int* __restrict__ fix_cnt,
                                                             const int* __restrict__ fix_list,
                                                             int64_t* __restrict__ out_idx, float* __restrict__ out_dist,
                                                             int64_t* __restrict__ rest_ids,
                                                             unsigned long long* __restrict__ rest_keys,
                                                             int* __restrict__ rest_count) {
  const int nfail = *fail_count;
  const int lane = threadIdx.x & 31;
  for (int w = blockIdx.x * 8 + (threadIdx.x >> 5); w < nfail; w += gridDim.x * 8) {
    const int64_t qi = fail_ids[w];
    const int c = w < FIXTC_ROWS ? fix_cnt[w] : -1;
    bool done = false;
    if (c >= k && c <= FIXTC_CAP) {
      const int64_t self = self_offset >= 0 ? qi + self_offset : -1;
      const float* a = q + qi * ld;
      unsigned long long key[FIXTC_CAP / 32];
      int live = 0;
#pragma unroll
      for (int u = 0; u < FIXTC_CAP / 32; ++u) {
        key[u] = ~0ull;
        const int e = lane + 32 * u;
        if (e < c) {
          const int j = fix_list[(size_t)w * FIXTC_CAP + e];
          if (j != self) {
            const float* b = ref + (int64_t)j * ld;
            double acc = 0.0;
            for (int t = 0; t < dim; ++t) {
              const double df = __dsub_rn((double)a[t], (double)__ldg(b + t));
              acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
            key[u] = ((unsigned long long)__float_as_uint((float)acc) << 32) | (unsigned)j;
            ++live;
          }
        }
      }
      live = warp_sum(live);
      if (live >= k) {
        done = true;
        for (int r = 0; r < k; ++r) {
          unsigned long long best = key[0];
#pragma unroll
          for (int u = 1; u < FIXTC_CAP / 32; ++u) best = best < key[u] ? best : key[u];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(SCF_FULL, best, o);
            best = other < best ? other : best;
          }
#pragma unroll
          for (int u = 0; u < FIXTC_CAP / 32; ++u)
            if (key[u] == best) key[u] = ~0ull;
          if (lane == 0) {
            out_idx[qi * k + r] = (int64_t)(best & 0xffffffffull);
            out_dist[qi * k + r] = __uint_as_float((unsigned)(best >> 32));
          }
        }
      }
    }
    if (!done && lane == 0) {
      const int slot = atomicAdd(rest_count, 1);
      rest_ids[slot] = qi;
      rest_keys[slot] = fail_keys[w];
    }
  }
}

int pick_kc(int k) {
  const int need = k + 1;  // the query itself may be among the candidates
  if (need <= 12) return 16;
  if (need <= 25) return 32;
  return 0;
}

struct Plan {
  int kp, kchunks, kc, stages, nsplit, tiles_per_split, n_ref_tiles;
  int64_t nq_pad, nr_pad;
  size_t smem;
  // workspace offsets (bytes)
  size_t off_qop, off_rop, off_qn, off_qe, off_cs, off_ci, off_tau, off_fail, off_fkey, off_misc, off_fix, off_qfix,
      off_fthr, off_fcnt, off_flist, off_rest, off_rkey, total;
};

bool make_plan(int64_t nq, int64_t nref, int dim, int k, Plan& pl) {
  pl.kc = pick_kc(k);
  pl.kp = (dim + 3 + KCH - 1) / KCH * KCH;
  pl.kchunks = pl.kp / KCH;
  if (pl.kc == 0 || pl.kchunks > 4) return false;  // k > 24 or dim > 125: method 0 handles those
  pl.nq_pad = (nq + QT * BM - 1) / (QT * BM) * (QT * BM);
  pl.nr_pad = (nref + BN - 1) / BN * BN;
  pl.n_ref_tiles = (int)(pl.nr_pad / BN);
  const int64_t ctas = pl.nq_pad / (QT * BM);
  // One candidate list per query over the WHOLE reference range keeps the number of list updates at
  // k' ln(N/k'); the references are split across CTAs only when there are too few query tiles to fill the GPU.
  int s = 1;
  if (ctas < SCF_NUM_SMS) s = (int)std::min<int64_t>((SCF_NUM_SMS + ctas - 1) / ctas, 8);
  s = std::max(1, std::min(s, pl.n_ref_tiles / 8));
  while (s > 1 && s * pl.kc > 32 * MAXU) --s;
  pl.tiles_per_split = (pl.n_ref_tiles + s - 1) / s;
  pl.nsplit = (pl.n_ref_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
  auto smem_for = [&](int stages) {
    return (size_t)QT * pl.kchunks * A_CHUNK_BYTES + (size_t)stages * B_STAGE_BYTES + (size_t)pl.kc * NLISTS * 4 +
           (size_t)32 * NLISTS * 4 + (size_t)(1 + 2 * stages + 4) * 8 + 64;
  };
  pl.stages = 6;
  while (pl.stages > 2 && smem_for(pl.stages) + 1024 > 227 * 1024) --pl.stages;
  pl.smem = smem_for(pl.stages) + 1024;  // slack for the 1024-byte alignment of the swizzled tiles
  if (pl.smem > 227 * 1024) return false;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  pl.off_qop = o, o = al(o + (size_t)pl.nq_pad * pl.kp * 4);
  pl.off_rop = o, o = al(o + (size_t)pl.nr_pad * pl.kp * 4);
  pl.off_qn = o, o = al(o + (size_t)pl.nq_pad * 8);
  pl.off_qe = o, o = al(o + (size_t)pl.nq_pad * 4);
  pl.off_cs = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * pl.kc * 4);
  pl.off_ci = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * pl.kc * 4);
  pl.off_tau = o, o = al(o + (size_t)pl.nq_pad * pl.nsplit * 4);
  pl.off_fail = o, o = al(o + (size_t)nq * 8);
  pl.off_fkey = o, o = al(o + (size_t)nq * 8);
  pl.off_misc = o, o = al(o + 256);
  pl.off_fix = o, o = al(o + knn_exact_fix_scratch_bytes(nq, k));
  pl.off_qfix = o, o = al(o + (size_t)FIXTC_ROWS * pl.kp * 4);
  pl.off_fthr = o, o = al(o + (size_t)FIXTC_ROWS * 4);
  pl.off_fcnt = o, o = al(o + (size_t)FIXTC_ROWS * 4);
  pl.off_flist = o, o = al(o + (size_t)FIXTC_ROWS * FIXTC_CAP * 4);
  pl.off_rest = o, o = al(o + (size_t)nq * 8);
  pl.off_rkey = o, o = al(o + (size_t)nq * 8);
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
  if (!make_plan(nq, nref, dim, k, pl))  // k > 24 or dim > 125: outside the tensor-core kernel's shapes
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
  float* qe = (float*)(ws + pl.off_qe);
  float* cs = (float*)(ws + pl.off_cs);
  int* ci = (int*)(ws + pl.off_ci);
  float* ctau = (float*)(ws + pl.off_tau);
  int64_t* fail_ids = (int64_t*)(ws + pl.off_fail);
  unsigned long long* fail_keys = (unsigned long long*)(ws + pl.off_fkey);
  int* fail_count = (int*)(ws + pl.off_misc);
  float* bmax = (float*)(ws + pl.off_misc + 64);
  int* rest_count = (int*)(ws + pl.off_misc + 128);
  float* qfix = (float*)(ws + pl.off_qfix);
  float* fix_thr = (float*)(ws + pl.off_fthr);
  int* fix_cnt = (int*)(ws + pl.off_fcnt);
  int* fix_list = (int*)(ws + pl.off_flist);
  int64_t* rest_ids = (int64_t*)(ws + pl.off_rest);
  unsigned long long* rest_keys = (unsigned long long*)(ws + pl.off_rkey);
  cudaError_t e = cudaMemsetAsync(ws + pl.off_misc, 0, 256, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(fix_cnt, 0, (size_t)FIXTC_ROWS * 4, stream);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  knn_prep_kernel<<<4 * SCF_NUM_SMS, 256, 0, stream>>>(q, nq, pl.nq_pad, ref, nref, pl.nr_pad, dim, ld, pl.kp, qop, rop,
                                                       qn, qe, bmax);
  int32_t rc = scf_check_launch("scf_knn_l2(prep)");
  if (rc) return rc;
  CUtensorMap tq, tr;
  rc = scf_make_tmap_2d_f32(&tq, qop, (uint64_t)pl.nq_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BM);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f32(&tr, rop, (uint64_t)pl.nr_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BN);
  if (rc) return rc;
  KnnTcParams prm;
  prm.kchunks = pl.kchunks, prm.stages = pl.stages, prm.n_ref_tiles = pl.n_ref_tiles;
  prm.nq = (int)nq;
  prm.ksteps_last = ((dim + 3 + 7) / 8) - (pl.kchunks - 1) * (KCH / 8);
  prm.tiles_per_split = pl.tiles_per_split, prm.nsplit = pl.nsplit, prm.cand_score = cs, prm.cand_idx = ci, prm.cand_tau = ctau;
  {
    const char* f = getenv("SCF_KNN_FLAGS");
    prm.flags = f ? atoi(f) : 2;
  }
  prm.fail_count = nullptr, prm.fix_thr = nullptr, prm.fix_cnt = nullptr, prm.fix_list = nullptr;
  auto kern = pl.kc == 16 ? knn_tc_kernel<16, false> : knn_tc_kernel<32, false>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  dim3 grid((unsigned)(pl.nq_pad / (QT * BM)), (unsigned)pl.nsplit);
  kern<<<grid, NTHREADS, pl.smem, stream>>>(tq, tr, prm);
  rc = scf_check_launch("scf_knn_l2(tcgen05)");
  if (rc) return rc;
  knn_rerank_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(q, nq, ref, nref, dim, ld, k, self_offset, pl.kc,
                                                                  pl.nsplit, cs, ci, ctau, qn, qe, bmax, out_idx, out_dist,
                                                                  fail_ids, fail_keys, fix_thr, fail_count);
  rc = scf_check_launch("scf_knn_l2(rerank)");
  if (rc) return rc;
  if (SCF_KNN_DEBUG & 16) {
    unsigned long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_dbg, sizeof(h));
    const double chunks = (double)pl.nq_pad / 32.0 * (double)pl.nr_pad / 32.0;  // (32 rows x 32 columns) units
    fprintf(stderr, "[knn dbg] nsplit %d kc %d warp-chunks %.3g hit-chunks %llu (%.1f%%) hit-groups %llu drain-iters %llu inserts %llu (%.1f / query)\n",
            pl.nsplit, pl.kc, chunks, h[0], 100.0 * h[0] / chunks, h[1], h[2], h[3], (double)h[3] / (double)nq);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(g_dbg, h, sizeof(h));
  }
  // ---- repair of the rows whose guard could not be proven: tensor-core collect pass, then FP64 on the few left ----
  knn_fix_gather_kernel<<<SCF_NUM_SMS, 256, 0, stream>>>(qop, pl.kp, fail_ids, fail_count, qfix);
  rc = scf_check_launch("scf_knn_l2(fix,gather)");
  if (rc) return rc;
  CUtensorMap tf;
  rc = scf_make_tmap_2d_f32(&tf, qfix, (uint64_t)FIXTC_ROWS, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BM);
  if (rc) return rc;
  KnnTcParams fp = prm;
  fp.nq = FIXTC_ROWS;
  fp.nsplit = std::max(1, std::min(FIXTC_NSPLIT, pl.n_ref_tiles));
  fp.tiles_per_split = (pl.n_ref_tiles + fp.nsplit - 1) / fp.nsplit;
  fp.nsplit = (pl.n_ref_tiles + fp.tiles_per_split - 1) / fp.tiles_per_split;
  fp.fail_count = fail_count, fp.fix_thr = fix_thr, fp.fix_cnt = fix_cnt, fp.fix_list = fix_list;
  auto fkern = knn_tc_kernel<16, true>;
  e = cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  fkern<<<dim3(FIXTC_ROWS / (QT * BM), (unsigned)fp.nsplit), NTHREADS, pl.smem, stream>>>(tf, tr, fp);
  rc = scf_check_launch("scf_knn_l2(fix,collect)");
  if (rc) return rc;
  knn_fix_finish_kernel<<<SCF_NUM_SMS, 256, 0, stream>>>(q, ref, dim, ld, k, self_offset, fail_ids, fail_keys, fail_count,
                                                         fix_cnt, fix_list, out_idx, out_dist, rest_ids, rest_keys,
                                                         rest_count);
  rc = scf_check_launch("scf_knn_l2(fix,finish)");
  if (rc) return rc;
  return knn_exact_fix_launch(q, rest_ids, rest_keys, rest_count, nq, ref, nref, dim, ld, k, self_offset, out_idx,
                              out_dist, ws + pl.off_fix, stream);
}
