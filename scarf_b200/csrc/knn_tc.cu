// K5, method 1: exact kNN = tcgen05 FP16 candidate generation + exact FP64 re-rank + proven guard band.
//
//   score(i,j) = |b_j|^2 - 2 a_i.b_j  (= d(i,j) - |a_i|^2) is produced by ONE tensor-core contraction: the
//   operands are augmented along K,   A' = [a, 1, 1, 1, 0..],  B' = [-2b, n_hi, n_mid, n_lo, 0..]  with
//   |b|^2 = n_hi + n_mid + n_lo split into FP16-exact pieces, so the epilogue has no per-element arithmetic
//   besides the top-k' filter.
//
//   Operands are FP16 (kind::f16, FP32 accumulate), not TF32: both formats keep an 11-bit significand, so the
//   guard band is the same, but an FP16 instruction covers K = 16 instead of 8 for the same 32 bytes per operand
//   row: half the instructions, and the reference stream through L2 / shared memory halves too.
//   FP16's 5-bit exponent is handled by an exact power-of-two scale s (both operands are multiplied by s before
//   rounding, s^2 max|b|^2 <= 2^15): distances scale by s^2, the guard works in scaled units.
//
//   kernel 0 (knn_range)   max |b|^2 over the references, max |a_t| over the queries -> s.
//   kernel 1 (knn_prep)    builds A', B' (s * value rounded to FP16, round-to-nearest), |a|^2, max|b|, rounding norms.
//   kernel 2 (knn_tc)      one CTA per SM walks its share of the (256-query tile, 128-reference tile) space (SegWalk:
//                          rounds with all CTAs in step, then a balanced remainder): TMA -> smem (128B swizzle) ->
//                          tcgen05.mma kind::f16 (M=128, N=128, K=16), one issuing warp per query tile -> TMEM
//                          (2 x 2 x 128 columns, double buffered) -> 16 epilogue warps, one query row x 64 columns per
//                          thread, keep the k' smallest.  knn_tc_pair_kernel: the same over CTA pairs (cta_group::2),
//                          opt-in.
//   kernel 3 (knn_rerank)  one warp per query: the oracle's FP64 distance for every candidate that can still reach the k
//                          nearest, order by (float32 distance, index), and the guard: everything the tensor cores
//                          rejected is provably farther than the k-th kept neighbour, else the row goes on the fail list.
//   repair                 collect pass (kernel 2 with a fixed per-row threshold, append-only) + FP64 selection for the
//                          fail list; the FP64 brute force (knn_exact) for what is left (device-side counts, no host
//                          round trip).
// Developer switches (environment): SCF_KNN_PAIR=1 (CTA-pair kernel), SCF_KNN_ISSUERS=1|2, SCF_KNN_ROUNDS=0|1,
// SCF_KNN_FLAGS; compile-time: SCF_KNN_DEBUG (timing skeletons, counters), SCF_KNN_HB (hit-buffer depth).
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "knn_common.cuh"
#include "tc_common.cuh"

namespace {

#ifndef SCF_KNN_DEBUG
#define SCF_KNN_DEBUG 0  // compile-time developer switches: 4 / 8 / 32 / 64 = timing experiments (no drain / read-out only / no read-out / no MMA), 16 = counters
#endif
__device__ unsigned long long g_dbg[8];  // SCF_KNN_DEBUG & 16: event counters (developer diagnostics)

constexpr int BM = 128;        // queries per MMA (= TMEM lanes)
constexpr int KCH = 64;        // FP16 values per 128-byte swizzle row
constexpr int KSTEP = 16;      // K per tcgen05.mma kind::f16 (32 bytes of every operand row)
constexpr int A_CHUNK_BYTES = BM * 128;
constexpr int ACC_COLS = 256;  // TMEM columns of one accumulator buffer: QT query tiles x BN references
constexpr int GCOLS = 64;      // accumulator columns one epilogue thread examines per reference tile
constexpr int NEPI_WARPS = 16;
constexpr int NTHREADS = 96 + 32 * NEPI_WARPS;  // warp 0: TMA, warp 1: MMA of query tile 0 + TMEM owner, warps 2-17: epilogue,
                                                // warp 18: MMA of query tile 1
constexpr int NTHREADS_PAIR = 64 + 32 * NEPI_WARPS;  // knn_tc_pair_kernel: warp 0 TMA, warp 1 MMA (leader CTA), warps 2-17 epilogue
constexpr int NL = 32 * NEPI_WARPS;             // epilogue threads = candidate lists per CTA
#ifndef SCF_KNN_HB
#define SCF_KNN_HB 8
#endif
constexpr int HB = SCF_KNN_HB;  // buffered hits per epilogue thread
constexpr int NDONE = 16;      // ring of "all MMAs of step t have completed" barriers (step t uses entry t % 16)

// error model of the tensor-core score (see DESIGN.md "kNN guard band"), all in scaled units (a := s a, b := s b),
// a~ = fp16(a), da = a - a~ (known exactly):
//   |2 a.b - 2 a~.b~| = 2 |da.b + a~.db| <= 2 (|da| |b| + |a| |db|)      with |da| per query, max |b|, max |db|
//   fp16 x fp16 products are exact in FP32; accumulation of kp terms in any order with a truncating adder
//   (2^-23 per add, x4 margin) -> 4 kp 2^-23 (2|a||b| + |b|^2); the three-way FP16 split of |b|^2 leaves less than
//   2^-30 |b|^2 + 2^-24 (the last piece may be subnormal)
__device__ __host__ inline double eps_acc(int kp) { return 4.0 * kp * 0x1p-23 + 0x1p-30; }
constexpr double EPS_SPLIT_ABS = 0x1p-24;

// the power-of-two operand scale: s^2 max|b|^2 <= 2^15 and s max|a_t| <= 2^14 (FP16 holds up to 65504)
__device__ __forceinline__ float knn_scale(const float* __restrict__ range) {
  const float bn2 = range[0], amax = range[1];
  int e = 0;  // s = 2^e
  if (bn2 > 0.f) {
    int eb;
    frexpf(bn2, &eb);  // bn2 < 2^eb
    e = (15 - eb) >> 1;  // floor: s^2 bn2 < 2^(2e + eb) <= 2^15
  }
  if (amax > 0.f) {
    int ea;
    frexpf(amax, &ea);  // amax < 2^ea
    e = min(e, 14 - ea);
  }
  e = max(-60, min(e, 14));  // large e only blows tiny data up into the normal FP16 range
  return ldexpf(1.f, e);
}

// Eight lanes per row, four rows per warp and step (kernels 0 and 1): one row per warp and step left the loads of a row
// (128 floats: one instruction per lane) and the warp-wide FP64 reductions behind them on the critical path of every
// row -- 0.65 + 1.18 ms for the 1.1 M rows of a C3 shard, a tenth of the HBM rate.
__device__ __forceinline__ float group8_sumf(float v, unsigned mask) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__device__ __forceinline__ double group8_sumd(double v, unsigned mask) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__global__ void __launch_bounds__(256) knn_range_kernel(const float* __restrict__ q, int64_t nq,
                                                        const float* __restrict__ ref, int64_t nref, int dim, int64_t ld,
                                                        float* __restrict__ range) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const unsigned gmask = 0xFFu << (lane & ~7);
  const int64_t g0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3), ng = (int64_t)gridDim.x * 32;
  float bn2 = 0.f, amax = 0.f;
  for (int64_t r = g0; r < nq + nref; r += ng) {
    const bool is_q = r < nq;
    const float* src = is_q ? q + r * ld : ref + (r - nq) * ld;
    float n2 = 0.f, m = 0.f;
    for (int t = sub; t < dim; t += 8) {
      const float x = src[t];
      n2 = fmaf(x, x, n2);
      m = fmaxf(m, fabsf(x));
    }
    if (is_q) {
      amax = fmaxf(amax, m);
    } else {
      n2 = group8_sumf(n2, gmask);
      bn2 = fmaxf(bn2, n2 * 1.001f);  // margin for the float32 summation
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(SCF_FULL, amax, o));
    bn2 = fmaxf(bn2, __shfl_xor_sync(SCF_FULL, bn2, o));
  }
  if (lane == 0) {
    // non-negative floats order like their bit patterns; NaN / inf inputs saturate the scale, the FP64 re-rank and
    // the guard then send such rows to the exact path
    if (bn2 > 0.f) atomicMax(reinterpret_cast<int*>(range), __float_as_int(fminf(bn2, 3e38f)));
    if (amax > 0.f) atomicMax(reinterpret_cast<int*>(range + 1), __float_as_int(fminf(amax, 3e38f)));
  }
}

// ---------------------------------------------------------------------------------------------- prep
// eight lanes per row; rows [0, nq_pad) of qop and [0, nr_pad) of rop.  Pad rows are all zero: pad queries never keep a
// candidate (threshold -FLT_MAX), pad references are rejected by index in the epilogue.
__global__ void __launch_bounds__(256) knn_prep_kernel(const float* __restrict__ q, int64_t nq, int64_t nq_pad,
                                                       const float* __restrict__ ref, int64_t nref, int64_t nr_pad,
                                                       int dim, int64_t ld, int kp, __half* __restrict__ qop,
                                                       __half* __restrict__ rop, double* __restrict__ qnorm2,
                                                       float* __restrict__ qerr, const float* __restrict__ range,
                                                       float* __restrict__ bmax) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const unsigned gmask = 0xFFu << (lane & ~7);
  const int64_t g0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3), ng = (int64_t)gridDim.x * 32;
  const float sc = knn_scale(range);
  float local_bmax = 0.f, local_dbmax = 0.f;  // bmax[0] = max |b|, bmax[1] = max |b - fp16(b)|  (scaled units)
  const int64_t total = nq_pad + nr_pad;
  // (every lane of a warp runs the same number of iterations: the group reductions below need all 32 lanes converged)
  for (int64_t r0 = g0 - (lane >> 3); r0 < total; r0 += ng) {
    const int64_t r = r0 + (lane >> 3);
    const bool in_range = r < total;
    const bool is_q = r < nq_pad;
    const int64_t row = is_q ? r : r - nq_pad;
    const bool live = in_range && (is_q ? row < nq : row < nref);
    const float* src = (is_q ? q : ref) + row * ld;
    __half* dst = (is_q ? qop : rop) + row * kp;
    double n2 = 0.0, e2 = 0.0;
    if (in_range)
      for (int t = sub; t < kp; t += 8) {
        if (t >= dim && t < dim + 3) continue;  // the three augmentation slots are written below
        __half v = __float2half_rn(0.f);
        if (live && t < dim) {
          const float x = src[t] * sc;  // exact: sc is a power of two
          const __half xr = __float2half_rn(x);
          const double dx = (double)x - (double)__half2float(xr);
          n2 += (double)x * (double)x;
          e2 += dx * dx;
          v = is_q ? xr : __float2half_rn(-2.f * __half2float(xr));  // exact doubling
        }
        dst[t] = v;
      }
    n2 = group8_sumd(n2, gmask);
    e2 = group8_sumd(e2, gmask);
    if (in_range && sub < 3) {
      float aug = 0.f;
      if (is_q) {
        aug = live ? 1.f : 0.f;
      } else if (live) {
        const float hi = __half2float(__float2half_rn((float)n2));
        const float mid = __half2float(__float2half_rn((float)(n2 - (double)hi)));
        const float lo = __half2float(__float2half_rn((float)(n2 - (double)hi - (double)mid)));
        aug = sub == 0 ? hi : (sub == 1 ? mid : lo);
      }
      dst[dim + sub] = __float2half_rn(aug);
    }
    if (sub == 0 && live) {
      if (is_q) {
        qnorm2[row] = n2;
        qerr[row] = (float)sqrt(e2) * 1.0000002f;
      } else {
        local_bmax = fmaxf(local_bmax, (float)sqrt(n2) * 1.0000002f);
        local_dbmax = fmaxf(local_dbmax, (float)sqrt(e2) * 1.0000002f);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    local_bmax = fmaxf(local_bmax, __shfl_xor_sync(SCF_FULL, local_bmax, o));
    local_dbmax = fmaxf(local_dbmax, __shfl_xor_sync(SCF_FULL, local_dbmax, o));
  }
  if (lane == 0 && local_bmax > 0.f) atomicMax(reinterpret_cast<int*>(bmax), __float_as_int(local_bmax));
  if (lane == 0 && local_dbmax > 0.f) atomicMax(reinterpret_cast<int*>(bmax + 1), __float_as_int(local_dbmax));
  if (blockIdx.x == 0 && threadIdx.x == 0) bmax[2] = sc * sc;  // distances in scaled units = s^2 * distance
}

// ---------------------------------------------------------------------------------------------- main
// Per-thread UNSORTED candidate list: the KC scores live in registers, the ids in global memory (entry e of the list
// at gid[e]; written only when a candidate is inserted, L2 absorbs it).  The slot number is carried in the low
// mantissa bits of every kept score (<= 31 ulp perturbation; the threshold is rounded down past the tag, so every
// rejected score is >= thr), so one FMNMX3 tree yields both the largest kept score and the slot that holds it: a new
// candidate overwrites it.
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
  __device__ __forceinline__ void replace_max(float s, int j, int* __restrict__ gid) {
    const uint32_t pmax = __float_as_uint(tmax) & SLOT_MASK;
    gid[pmax] = j;
    const float tagged = __uint_as_float((__float_as_uint(s) & ~SLOT_MASK) | pmax);
    const uint32_t onehot = 1u << pmax;  // selects, not a dynamically indexed store (keeps r[] in registers)
#pragma unroll
    for (int e = 0; e < KC; ++e) r[e] = (onehot & (1u << e)) ? tagged : r[e];
    recompute();
  }
};

struct KnnTcParams {
  int kchunks;          // Kp / 64
  int ksteps_last;      // 16-wide K steps that hold data in the last chunk (all-zero padding steps are skipped)
  int stages;           // B pipeline depth
  int n_ref_tiles;      // nr_pad / BN
  int tiles_per_split;
  int nlists;           // candidate lists per query row: nsplit (reference ranges) x HALVES (column halves)
  float* cand_score;    // [nq_pad, nlists, KC]  (unsorted)
  int* cand_idx;
  float* cand_tau;      // [nq_pad, nlists]  largest kept score = lower bound of every rejected score
  int balanced;         // 0: static grid (blockIdx.x = query tile, blockIdx.y = reference range: the collect pass);
                        // 1: the schedule of SegWalk below
  int rounds;           //    whole query tiles every unit (CTA / CTA pair) sweeps first, all units in step
  int units_rem;        //    units that share the remaining query tiles (the others are done after their rounds)
  long long rem_total;  //    tile steps of the remaining query tiles = (n_q_tiles - rounds * units) * n_ref_tiles
  int nq;               // live query rows (pad rows keep nothing)
  int nref;             // live reference rows (pad rows are rejected by index)
  int flags;            // developer switches (SCF_KNN_FLAGS): 2 = back off in waits
  // collect pass (repair of guard failures): query row r of this launch is failed row r of the fail list
  const int* fail_count;   // device count of failed rows (rows >= min(count, FIXTC_ROWS) do nothing)
  const float* fix_thr;    // [FIXTC_ROWS] score threshold: every reference scoring below it is collected
  int* fix_cnt;            // [FIXTC_ROWS] collected so far
  int* fix_list;           // [FIXTC_ROWS][FIXTC_CAP] reference ids
};

constexpr int FIXTC_ROWS = 32768;  // failed rows repaired by the tensor-core collect pass (the rest: FP64 scan); ~1 % of the rows fail on the C3 embedding
constexpr int FIXTC_CAP = 512;    // collected references per failed row (64 left 825 of the 1,140 failed rows of the C3 embedding to the FP64 scan: 24 ms)
constexpr int FIXTC_NSPLIT = 16;  // reference ranges per query tile in the collect pass

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }  // one FMNMX3
__device__ __forceinline__ float min4(const uint32_t* v) {
  return fminf(fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2])), __uint_as_float(v[3]));
}
__device__ __forceinline__ float min16(const uint32_t* v) {  // eight FMNMX3 / FMNMX
  const float t0 = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
  const float t1 = fmin3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
  const float t2 = fmin3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
  const float t3 = fmin3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
  const float t4 = fmin3(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
  return fminf(fmin3(t0, t1, t2), fmin3(t3, t4, __uint_as_float(v[15])));
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// Epilogue state of one thread: the candidate list of (its query row, its 64-column group) plus a small buffer of
// hits in shared memory.  A score below the threshold is APPENDED to the buffer (three predicated instructions); the
// buffers of a warp are drained together -- every lane inserts its own entries into its own register list at the same
// time -- when one of them is nearly full.  In the steady state of a search hits are rare and scattered over the
// lanes: an immediate insert costs the whole warp ~KC + 30 instructions for one lane's benefit, the buffered one is
// shared by all lanes that have collected something since the last drain.  The threshold is stale between drains
// (never too small): entries are checked again against the fresh one when they are drained.
template <int KC>
struct Epi {
  CandList<KC> cl;
  float thr;        // comparison threshold in use (>= cl.thr; with two lists per row also <= the partner's)
  int cnt;          // buffered hits
  float* hs;        // hit scores  [HB][NL] (this thread's column)
  int* hc;          // hit reference ids
  int* gid;         // this list's candidate ids in global memory
  int nref;
  __device__ __forceinline__ void drain() {
    const bool count = (SCF_KNN_DEBUG & 16) && (threadIdx.x & 31) == 0;
    if (count) atomicAdd(&g_dbg[3], 1ull);
    for (int i = 0; __any_sync(SCF_FULL, i < cnt); ++i) {
      if (count) atomicAdd(&g_dbg[4], 1ull);
      if (i < cnt) {
        const float sc = hs[i * NL];
        const int j = hc[i * NL];
        if (SCF_KNN_DEBUG & 16) atomicAdd(&g_dbg[5], 1ull);
        if (sc < cl.thr && j < nref) {
          if (SCF_KNN_DEBUG & 16) atomicAdd(&g_dbg[6], 1ull);
          cl.replace_max(sc, j, gid);
        }
      }
    }
    cnt = 0;
  }
  // one buffered entry per lane (the last one appended): what a waiting warp does instead of spinning
  __device__ __forceinline__ void drain_one() {
    if (cnt > 0) {
      --cnt;
      const float sc = hs[cnt * NL];
      const int j = hc[cnt * NL];
      if (sc < cl.thr && j < nref) cl.replace_max(sc, j, gid);
    }
  }
  // Wait for the accumulators of a step.  Every epilogue warp of the CTA has to release an accumulator buffer before
  // the MMA warp may overwrite it, so the warps wait for each other once per step, and a step in which SOME warp updates
  // its lists is as slow as that warp (ncu: a third of the warp stall samples sit on this wait).  The waiting time is
  // therefore used for the list updates: hits are only appended to the buffers inside a step (cheap), and drained here,
  // one entry per poll, while the barrier is not ready; a drain inside a step happens only when a buffer runs full.
  // -> true when entries were drained (the caller refreshes / publishes its threshold)
  __device__ __forceinline__ bool wait_draining(uint64_t* bar, uint32_t parity) {
    bool drained = false;
    uint32_t spins = 0;
    while (!tc::mbar_try_wait(bar, parity)) {
      if (__any_sync(SCF_FULL, cnt > 0)) {
        drain_one();
        drained = true;
      } else if ((++spins & 0xFFFFu) == 0u) {
        tc::spin_timeout(spins);
      }
    }
    return drained;
  }
};

// Some lane of the warp has a score below its threshold among v[0..32).  What bounds this kernel once the tensor pipe
// is fed is the number of instructions a warp spends per reference tile (ncu: ~10 cycles between two instructions of
// one warp), so the scan is organised to skip as much as it can with warp-uniform branches: chunks of 16 and groups of
// 4 columns are visited only when some lane has a candidate in them (one vote each), a visited group costs four
// predicated appends, and before a group is appended the buffers are drained if one of them could overflow.
// smallest published threshold of the NSHARE other lists of this thread's row (epilogue thread l ^ 128 j, j = 1..NSHARE)
template <int NSHARE>
__device__ __forceinline__ float partner_thr(const volatile float* thr_all, int l) {
  float m = FLT_MAX;
#pragma unroll
  for (int j = 1; j <= NSHARE; ++j) m = fminf(m, thr_all[l ^ (128 * j)]);
  return m;
}

template <int KC, int NSHARE>
__device__ __forceinline__ void hits32(const uint32_t (&v)[32], float m_lo, float m_hi, int jb, Epi<KC>& ep,
                                       volatile float* thr_all, int l) {
  const bool count = (SCF_KNN_DEBUG & 16) && (threadIdx.x & 31) == 0;
  if (count) atomicAdd(&g_dbg[1], 1ull);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!__any_sync(SCF_FULL, (h ? m_hi : m_lo) < ep.thr)) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t* vg = v + 16 * h + 4 * i;
      if (!__any_sync(SCF_FULL, min4(vg) < ep.thr)) continue;
      if (count) atomicAdd(&g_dbg[2], 1ull);
      if (__any_sync(SCF_FULL, ep.cnt > HB - 4)) {
        ep.drain();
        ep.thr = ep.cl.thr;
        if constexpr (NSHARE > 0) {
          thr_all[l] = ep.cl.thr;
          ep.thr = fminf(ep.thr, partner_thr<NSHARE>(thr_all, l));
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float x = __uint_as_float(vg[c]);
        if (x < ep.thr) {
          ep.hs[ep.cnt * NL] = x;
          ep.hc[ep.cnt * NL] = jb + 16 * h + 4 * i + c;
          ++ep.cnt;
        }
      }
    }
  }
}

// The work of one unit (a CTA, or a pair of CTAs) as a sequence of segments (query tile q, reference tiles [t0, t1)).
// Rounds first: in round r unit u sweeps ALL reference tiles for query tile r * units + u -- every unit is at (about)
// the same reference tile at the same time, so a tile is fetched from HBM once and served to the other units from L2.
// (Equal contiguous ranges of the whole (query tile, reference tile) space, the schedule of the first half of round 2,
// put the units at `units` different places of a reference set that does not fit L2 -- 256 MB at 1M cells -- and every
// tile load went to HBM: 94 GB of DRAM reads for the 125k x 1M shard, ncu; 1.1 GB with the rounds.)  Then the query
// tiles that are left (fewer than units) are cut into equal contiguous ranges of tile steps over `units_rem` units, so
// that all finish together; a tile cut by a range boundary keeps one candidate list per part (slot = this unit's index
// minus the index of the unit that holds the tile's first step).  Static grid (the collect pass): one segment.
struct SegWalk {
  int T, units, u, rounds, r, units_rem, q_rem0;
  long long cur, w1, rem_total;
  __device__ __forceinline__ void init(const KnnTcParams& p, int unit, int n_units) {
    T = p.n_ref_tiles, units = n_units, u = unit, r = 0;
    if (!p.balanced) {
      const int tile_begin = (int)blockIdx.y * p.tiles_per_split;
      const int tile_end = min(tile_begin + p.tiles_per_split, T);
      rounds = 0, units_rem = 1, q_rem0 = 0, rem_total = 0;
      cur = (long long)blockIdx.x * T + tile_begin;
      w1 = cur + max(tile_end - tile_begin, 0);
      return;
    }
    rounds = p.rounds, units_rem = p.units_rem, rem_total = p.rem_total, q_rem0 = p.rounds * n_units;
    if (u < units_rem) {
      cur = (long long)u * rem_total / units_rem;
      w1 = (long long)(u + 1) * rem_total / units_rem;
    } else {
      cur = w1 = 0;
    }
  }
  __device__ __forceinline__ bool next(int& q, int& t0, int& t1) {
    if (r < rounds) {
      q = r * units + u, t0 = 0, t1 = T, ++r;
      return true;
    }
    if (cur >= w1) return false;
    q = q_rem0 + (int)(cur / T), t0 = (int)(cur % T);
    t1 = (int)min((long long)T, (long long)t0 + (w1 - cur));
    cur += t1 - t0;
    return true;
  }
  // list slot of this unit's part of query tile q
  __device__ __forceinline__ int slot_of(const KnnTcParams& p, int q) const {
    if (!p.balanced) return (int)blockIdx.y;
    if (q < q_rem0) return 0;
    const long long target = (long long)(q - q_rem0) * T;  // first step of the tile: find c with b(c) <= target < b(c + 1)
    long long c = target * units_rem / rem_total;
    while (c + 1 < (long long)units_rem && (c + 1) * rem_total / units_rem <= target) ++c;
    while (c > 0 && c * rem_total / units_rem > target) --c;
    return u - (int)c;
  }
};

// Two query tiles of 128 rows per CTA (QT = 2) x reference tiles of BN = 128 rows.  (QT = 4 / BN = 64 halves the L2 ->
// shared-memory traffic but an N = 64 MMA is bound by operand bandwidth at 2/3 of the rate: tried and dropped.)
// Warps: 0 TMA; 1 MMA + TMEM owner; 2-17 epilogue; with NI = 2 warp 18 is the second MMA issuer (warp 1 issues for query
// tile 0, warp 18 for tile 1; NI = 1: warp 1 for both, no warp 18).  Epilogue warp w reads TMEM lane quarter (warp id %
// 4) and the 64 accumulator columns [64 g, 64 g + 64), g = w / 4, of the 256-column buffer: query tile g / 2, column
// half g % 2 -- two lists per row, which share their thresholds through shared memory.
template <int KC, int QT, bool COLLECT, int NI>
__global__ void __launch_bounds__(NI == 2 ? NTHREADS : NTHREADS - 32, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                             const __grid_constant__ CUtensorMap tmap_r,
                                                             const KnnTcParams p) {
  constexpr int BN = ACC_COLS / QT;
  static_assert(QT == 2, "two query tiles of 128 rows x reference tiles of 128 rows");
  constexpr int HALVES = 2;
  constexpr int B_STAGE_BYTES = BN * 128;
  int n_fix = 0;
  if constexpr (COLLECT) {  // whole CTA leaves before any barrier / TMEM setup when its query tile is empty
    n_fix = min(*p.fail_count, FIXTC_ROWS);
    if ((int)blockIdx.x * (QT * BM) >= n_fix) return;
  }
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the 128-byte swizzle is a function of the shared-memory address: tiles must sit on 1024-byte boundaries
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                             // [QT][kchunks][128 rows x 128 B]
  unsigned char* sB = sA + (size_t)QT * p.kchunks * A_CHUNK_BYTES;      // [stages][BN rows x 128 B]
  float* hb_score = reinterpret_cast<float*>(sB + (size_t)p.stages * B_STAGE_BYTES);  // hit buffers [HB][NL]
  int* hb_col = reinterpret_cast<int*>(hb_score + (size_t)HB * NL);
  float* thr_x = reinterpret_cast<float*>(hb_col + (size_t)HB * NL);    // published list thresholds [NL]
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr_x + NL);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 1;
  uint64_t* full = bars + 2;               // [stages] TMA data of a stage has landed
  // NI issuing warps = NI independent accumulator pipelines (NI == 2: one per query tile, NI == 1: both tiles together)
  uint64_t* done = full + p.stages;        // [NI][NDONE] every MMA of pipeline i in step t (entry t % NDONE) has
                                           //          completed: its epilogue warps may read the accumulators; the
                                           //          producer refills a stage once ALL pipelines have done the step
  uint64_t* tmem_empty = done + NI * NDONE;  // [NI][2]  the epilogue warps of pipeline i have read its buffer a
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + NI * 2);

  // the warp index through a shuffle: the compiler then knows that the role branches below are warp-uniform and may
  // use the uniform datapath inside them (UTCHMMA takes its descriptors from uniform registers; in a branch it
  // considers divergent every operand of every MMA goes through an R2UR move: ~100 cycles per instruction, measured)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t backoff = (p.flags & 2) ? 32u : 0u;
  SegWalk walk;
  walk.init(p, (int)blockIdx.x, (int)gridDim.x);

  if (threadIdx.x == 0) {
    tc::mbar_init(a_full, 1);
    tc::mbar_init(a_empty, NI);  // one commit per issuing warp
    for (int s = 0; s < p.stages; ++s) tc::mbar_init(full + s, 1);
    for (int d = 0; d < NI * NDONE; ++d) tc::mbar_init(done + d, 1);
    for (int a = 0; a < NI * 2; ++a) tc::mbar_init(tmem_empty + a, NEPI_WARPS / NI);  // one arrive per epilogue warp of the pipeline
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_q);
    tc::tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);  // 2 buffers x 256 columns
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================  (whole warp, one elected lane issues)
    const bool leader = tc::elect_one();
    int seg = 0;
    uint32_t s = 0, ph = 0;  // pipeline stage and its phase, advanced without divisions (stay in uniform registers)
    // A stage is refilled once the step that read its previous content has completed (`done` rings of the MMA warps):
    // one tcgen05.commit per step and query tile instead of one per stage.
    // (rel_step, rel_c): the chunk loaded p.stages chunks ago, i.e. the previous user of the stage about to be filled.
    uint32_t loaded = 0, rel_step = 0, rel_c = 0;
    for (int q, t0, t1; walk.next(q, t0, t1); ++seg) {
      if (seg > 0) tc::mbar_wait(a_empty, (uint32_t)(seg - 1) & 1u, backoff);  // MMAs of the last segment done with sA
      if (leader) {
        tc::mbar_expect_tx(a_full, (uint32_t)(QT * p.kchunks * A_CHUNK_BYTES));
        for (int t = 0; t < QT; ++t)
          for (int c = 0; c < p.kchunks; ++c)
            tc::tma_load_2d(sA + (size_t)(t * p.kchunks + c) * A_CHUNK_BYTES, &tmap_q, a_full, c * KCH,
                            q * (QT * BM) + t * BM);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t)
        for (int c = 0; c < p.kchunks; ++c) {
          if (loaded >= (uint32_t)p.stages) {
#pragma unroll
            for (int i = 0; i < NI; ++i)
              tc::mbar_wait(done + i * NDONE + (rel_step & (NDONE - 1)), (rel_step / NDONE) & 1u, backoff);
            if (++rel_c == (uint32_t)p.kchunks) rel_c = 0, ++rel_step;
          }
          ++loaded;
          if (leader) {
            if (SCF_KNN_DEBUG & 128) {  // timing experiment: no reference traffic, the MMAs run on stale tiles
              tc::mbar_arrive(full + s);
            } else {
              tc::mbar_expect_tx(full + s, B_STAGE_BYTES);
              tc::tma_load_2d(sB + (size_t)s * B_STAGE_BYTES, &tmap_r, full + s, c * KCH, t * BN);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) s = 0, ph ^= 1u;
        }
    }
  } else if (warp == 1 || warp == 2 + NEPI_WARPS) {
    // ===================== MMA issuers: one warp per query tile =====================
    // Two issuing warps because ONE cannot keep the tensor pipe fed: between two MMAs the issuing thread spends ~70
    // cycles on descriptor arithmetic and barrier polls (uniform datapath), as long as an N = 128 instruction runs, so
    // every wait of the issuer (for an accumulator buffer, for a stage) is a bubble in the pipe (ncu: MMA warp 31 % of its
    // time on the accumulator barrier, pipe 52 % active at the C3 shape; microbenchmark tools/ubench_tc.cu: two issuers
    // reach 64.1 cycles per instruction whatever the commit frequency, one issuer 65 - 75).  Each warp owns one query
    // tile: its half of both accumulator buffers, its own `done` ring and buffer barriers -- the two tiles hand their
    // accumulators over to their own eight epilogue warps independently of each other.
    // The whole warp walks the loop with identical (warp-uniform) values and one elected lane issues the tensor
    // instructions: descriptors then live in uniform registers.  Issuing from inside an `if (lane == 0)` region makes
    // the compiler move every descriptor from vector to uniform registers through a waterfall loop per MMA (~110
    // idle cycles between 64-cycle instructions, measured).
    constexpr uint32_t idesc = tc::umma_idesc_f16(BM, BN, false, false);
    const bool leader = tc::elect_one();
    // Everything the instructions take is derived from warp-uniform counters with add / shift / select only, so that
    // the descriptors stay in uniform registers (a division, e.g. `it % stages`, sends them through vector registers
    // and every MMA then waits for a chain of R2UR moves: ~100 cycles per instruction instead of 32 - 64, measured).
    // The tile loops are resolved at compile time (NI): with run-time tile bounds the extra uniform predicates per
    // instruction cost 10 - 40 % of the kernel (measured).
    const uint32_t nk_last = (uint32_t)p.ksteps_last, kchunks = (uint32_t)p.kchunks, stages = (uint32_t)p.stages;
    const uint32_t b_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(sB));
    uint32_t s = 0, ph = 0, lt = 0, seg = 0;
    if constexpr (NI == 2) {
      // this warp issues for query tile tq
      const uint32_t tq = warp == 1 ? 0u : 1u;
      const uint32_t a_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(sA)) + tq * kchunks * (uint32_t)(A_CHUNK_BYTES >> 4);
      uint64_t* my_done = done + tq * NDONE;
      uint64_t* my_empty = tmem_empty + tq * 2;
      for (int q, t0, t1; walk.next(q, t0, t1); ++seg) {
        tc::mbar_wait(a_full, seg & 1u);
        tc::tc_fence_after();
        for (int t = t0; t < t1; ++t, ++lt) {
          const uint32_t acc = lt & 1u;
          tc::mbar_wait(my_empty + acc, ((lt >> 1) & 1u) ^ 1u, backoff);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)ACC_COLS + tq * (uint32_t)BN;
          for (uint32_t c = 0; c < kchunks; ++c) {
            tc::mbar_wait(full + s, ph, backoff);
            tc::tc_fence_after();
            const uint32_t a_lo = a_lo0 + c * (uint32_t)(A_CHUNK_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + s * (uint32_t)(B_STAGE_BYTES >> 4);
            const uint32_t nk = c + 1 == kchunks ? nk_last : (uint32_t)(KCH / KSTEP);
#pragma unroll
            for (uint32_t kk = 0; kk < (uint32_t)(KCH / KSTEP); ++kk)  // K = 16 fp16 = 32 bytes: +2 in 16-byte units
              if (kk < nk && !(SCF_KNN_DEBUG & 64)) {
                if (leader) tc::umma_f16_lo(d_tmem, a_lo + kk * 2u, b_lo + kk * 2u, idesc, (c | kk) != 0u);
              }
            if (++s == stages) s = 0, ph ^= 1u;
          }
          if (leader) tc::umma_commit(my_done + (lt & (NDONE - 1)));  // the tile's accumulator complete, its stage reads done
          __syncwarp();
        }
        if (leader) tc::umma_commit(a_empty);  // the query operand may be replaced once every MMA of the segment has run
        __syncwarp();
      }
    } else if (warp == 1) {
      // one issuer for both query tiles (one K chunk per step, i.e. four instructions per tile: measured faster with
      // one issuer, C2 2.43 ms against 2.60 ms); the second MMA warp has nothing to do
      const uint32_t a_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(sA));
      for (int q, t0, t1; walk.next(q, t0, t1); ++seg) {
        tc::mbar_wait(a_full, seg & 1u);
        tc::tc_fence_after();
        for (int t = t0; t < t1; ++t, ++lt) {
          const uint32_t acc = lt & 1u;
          tc::mbar_wait(tmem_empty + acc, ((lt >> 1) & 1u) ^ 1u, backoff);
          tc::tc_fence_after();
          for (uint32_t c = 0; c < kchunks; ++c) {
            tc::mbar_wait(full + s, ph, backoff);
            tc::tc_fence_after();
            const uint32_t b_lo = b_lo0 + s * (uint32_t)(B_STAGE_BYTES >> 4);
            const uint32_t nk = c + 1 == kchunks ? nk_last : (uint32_t)(KCH / KSTEP);
#pragma unroll
            for (uint32_t tq = 0; tq < (uint32_t)QT; ++tq) {
              const uint32_t a_lo = a_lo0 + (tq * kchunks + c) * (uint32_t)(A_CHUNK_BYTES >> 4);
              const uint32_t d_tmem = tmem_base + acc * (uint32_t)ACC_COLS + tq * (uint32_t)BN;
#pragma unroll
              for (uint32_t kk = 0; kk < (uint32_t)(KCH / KSTEP); ++kk)
                if (kk < nk && !(SCF_KNN_DEBUG & 64)) {
                  if (leader) tc::umma_f16_lo(d_tmem, a_lo + kk * 2u, b_lo + kk * 2u, idesc, (c | kk) != 0u);
                }
            }
            if (++s == stages) s = 0, ph ^= 1u;
          }
          if (leader) tc::umma_commit(done + (lt & (NDONE - 1)));  // accumulators complete, stages of this step reusable
          __syncwarp();
        }
        if (leader) tc::umma_commit(a_empty);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: 64 accumulator columns of one query row per thread =====================
    const int w = warp - 2;
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int g = w >> 2;              // column group
    const int l = w * 32 + lane;       // epilogue thread id
    const int row = quarter * 32 + lane;
    const int qt = g >> 1;             // query tile
    const int half = g & 1;            // column half of the tile's 128 accumulator columns
    uint64_t* my_done = done + (NI == 2 ? qt : 0) * NDONE;
    uint64_t* my_empty = tmem_empty + (NI == 2 ? qt : 0) * 2;
    const int col0 = half * GCOLS;     // first reference of this thread inside a reference tile
    volatile float* thr_all = thr_x;  // the other list of this row (QT = 2): thread l ^ 128, warp w ^ 4
    const int pair_id = 1 + (w & 3) + 4 * (w >> 3);
    int lt = 0;
    for (int q, t0, t1; walk.next(q, t0, t1);) {
      const int qrow = q * (QT * BM) + qt * BM + row;
      if constexpr (COLLECT) {
        // fixed per-row threshold, append-only: every reference whose score is below it goes on the row's list
        const float thr = qrow < n_fix ? p.fix_thr[qrow] : -FLT_MAX;
        for (int t = t0; t < t1; ++t, ++lt) {
          const int acc = lt & 1;
          tc::mbar_wait(my_done + (lt & (NDONE - 1)), (uint32_t)(lt / NDONE) & 1u);
          tc::tc_fence_after();
          const int j0 = t * BN + col0;
          const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + g * GCOLS);
#pragma unroll 1
          for (int cc = 0; cc < GCOLS / 16; ++cc) {
            uint32_t v[16];
            tc::tmem_ld16(t_row + (uint32_t)(cc * 16), v);
            tc::tmem_ld_wait();
            float gm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) gm[i] = min4(v + 4 * i);
            const float m = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
            if (!__any_sync(SCF_FULL, m < thr)) continue;
            if (m < thr) {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (__uint_as_float(v[c]) < thr && j0 + cc * 16 + c < p.nref) {
                  const int pos = atomicAdd(p.fix_cnt + qrow, 1);
                  if (pos < FIXTC_CAP) p.fix_list[(size_t)qrow * FIXTC_CAP + pos] = j0 + cc * 16 + c;
                }
            }
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(my_empty + acc);
        }
      } else {
        const int split = walk.slot_of(p, q);
        const size_t sub = (size_t)qrow * p.nlists + (size_t)(split * HALVES + half);
        Epi<KC> ep;
        ep.cl.init();
        if (qrow >= p.nq) ep.cl.thr = -FLT_MAX;  // pad row: never a candidate
        ep.thr = ep.cl.thr;
        ep.cnt = 0;
        ep.hs = hb_score + l;
        ep.hc = hb_col + l;
        ep.gid = p.cand_idx + sub * KC;
        ep.nref = p.nref;
        if constexpr (HALVES == 2) {
          // both lists of a row publish the threshold of THIS segment before either reads the other's
          thr_all[l] = ep.cl.thr;
          pair_barrier(pair_id);
        }
        for (int t = t0; t < t1; ++t, ++lt) {
          const int acc = lt & 1;
          const bool drained = ep.wait_draining(my_done + (lt & (NDONE - 1)), (uint32_t)(lt / NDONE) & 1u);
          tc::tc_fence_after();
          if (drained) {
            ep.thr = ep.cl.thr;
            if constexpr (HALVES == 2) thr_all[l] = ep.cl.thr;
          }
          if (SCF_KNN_DEBUG & 32) {  // timing experiment: no accumulator read-out at all
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(my_empty + acc);
            continue;
          }
          const int j0 = t * BN + col0;
          const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + g * GCOLS);
          if constexpr (HALVES == 2) ep.thr = fminf(ep.cl.thr, partner_thr<1>(thr_all, l));
          // Two passes of 32 columns through one 32-register buffer (the list itself needs k' registers): minimum of
          // the pass (FMNMX3 tree), one vote; only a pass in which some lane has a candidate is scanned.
          uint32_t v[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            tc::tmem_ld32(t_row + (uint32_t)(32 * hh), v);
            tc::tmem_ld_wait();
            if (hh == 1) {
              // The accumulator buffer is released as soon as its last columns are in registers, BEFORE they are
              // examined: the hand-over (MMA done -> epilogue -> buffer free -> MMA of the step after next) is the
              // loop that paces the kernel with only two buffers in TMEM, and a scan with list updates is its longest
              // part.
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(my_empty + acc);
            }
            if (SCF_KNN_DEBUG & 8) continue;
            const float m_lo = min16(v), m_hi = min16(v + 16);
            if ((SCF_KNN_DEBUG & 16) && lane == 0) atomicAdd(&g_dbg[0], 1ull);
            if (__any_sync(SCF_FULL, fminf(m_lo, m_hi) < ep.thr))
              hits32<KC, HALVES == 2 ? 1 : 0>(v, m_lo, m_hi, j0 + 32 * hh, ep, thr_all, l);
          }
        }
        ep.drain();
        if constexpr (HALVES == 2) thr_all[l] = ep.cl.thr;
        // candidates out: [query, list, KC]; the ids are already in place
#pragma unroll
        for (int e = 0; e < KC; ++e) p.cand_score[sub * KC + e] = ep.cl.r[e];
        p.cand_tau[sub] = ep.cl.thr;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- main, CTA pairs
// The same search issued by PAIRS of CTAs (a 2-CTA cluster on the two SMs of a TPC, tcgen05 cta_group::2): one
// M = 256, N = 256 MMA covers 256 queries (128 from each CTA's shared memory) x 256 references (128 from each CTA's
// shared memory).  An N = 128 single-CTA MMA reads 128 B per clock from shared memory for its operands -- all there
// is -- so the TMA fill of the reference stages takes bandwidth away from the tensor pipe (measured: 15.7 ms of MMA
// become 25 ms with the fill at the C3 shape).  In a pair every CTA reads its 128 query rows and HALF of the reference
// tile per MMA (64 B per clock) and fills half a reference tile per step (32 B per clock): the tensor pipe is the bound.
// Per CTA: 1 query tile (128 rows) in shared memory, stages of 128 references, 2 accumulator buffers of 256 columns;
// epilogue warp w reads lane quarter (warp id % 4), columns [64 g, 64 g + 64), g = w / 4: four lists per query row
// (k' = 16 each: a row fails its guard only when 16 of its k <= 24 neighbours sit in one quarter of a tile), which
// share their thresholds through shared memory.
// Roles and barriers as in knn_tc_kernel, except that TMA completions of BOTH CTAs are counted on the leader's (even
// CTA's) barriers, tcgen05.commit arrives on the barriers of both CTAs (multicast), and the epilogue warps of both
// CTAs release an accumulator buffer on the leader's barrier.
template <int KC>
__global__ void __launch_bounds__(NTHREADS_PAIR, 1) knn_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                  const __grid_constant__ CUtensorMap tmap_r,
                                                                  const KnnTcParams p) {
  constexpr int BN2 = 256;                 // references per step (pair)
  constexpr int B_STAGE_BYTES = BM * 128;  // this CTA's half of a reference tile chunk
  constexpr int LISTS = 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                        // [kchunks][128 rows x 128 B]
  unsigned char* sB = sA + (size_t)p.kchunks * A_CHUNK_BYTES;      // [stages][128 rows x 128 B]
  float* hb_score = reinterpret_cast<float*>(sB + (size_t)p.stages * B_STAGE_BYTES);
  int* hb_col = reinterpret_cast<int*>(hb_score + (size_t)HB * NL);
  float* thr_x = reinterpret_cast<float*>(hb_col + (size_t)HB * NL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr_x + NL);
  uint64_t* a_full = bars;                 // leader's is used
  uint64_t* a_empty = bars + 1;            // both (multicast commit)
  uint64_t* full = bars + 2;               // [stages] leader's is used: bytes of both CTAs
  uint64_t* done = full + p.stages;        // [NDONE]  both (multicast commit)
  uint64_t* tmem_empty = done + NDONE;     // [2]      leader's is used: 2 x NEPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t backoff = (p.flags & 2) ? 32u : 0u;
  const uint32_t cta_rank = tc::cluster_ctarank();  // 0: leader
  const int pid = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  SegWalk walk;  // tiles of 256 references
  walk.init(p, pid, npairs);

  if (threadIdx.x == 0) {
    tc::mbar_init(a_full, 1);
    tc::mbar_init(a_empty, 1);
    for (int s = 0; s < p.stages; ++s) tc::mbar_init(full + s, 1);
    for (int d = 0; d < NDONE; ++d) tc::mbar_init(done + d, 1);
    for (int a = 0; a < 2; ++a) tc::mbar_init(tmem_empty + a, 2 * NEPI_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_q);
    tc::tma_prefetch_desc(&tmap_r);
  }
  __syncthreads();
  tc::cluster_sync();  // the barriers of both CTAs exist before either touches the other's
  if (warp == 1) tc::tmem_alloc_pair<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the leader's copies of the barriers both CTAs signal
  const uint32_t a_full_ld = tc::mapa_u32(tc::smem_u32(a_full), 0);
  const uint32_t full_ld = tc::mapa_u32(tc::smem_u32(full), 0);
  const uint32_t tmem_empty_ld = tc::mapa_u32(tc::smem_u32(tmem_empty), 0);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own query rows, own half of every reference tile) =============
    const bool leader = tc::elect_one();
    int seg = 0;
    uint32_t s = 0, ph = 0;
    uint32_t loaded = 0, rel_step = 0, rel_c = 0;
    for (int q, t0, t1; walk.next(q, t0, t1); ++seg) {
      if (seg > 0) tc::mbar_wait(a_empty, (uint32_t)(seg - 1) & 1u, backoff);
      if (leader) {
        if (cta_rank == 0) tc::mbar_expect_tx(a_full, (uint32_t)(2 * p.kchunks * A_CHUNK_BYTES));
        for (int c = 0; c < p.kchunks; ++c)
          tc::tma_load_2d_pair(sA + (size_t)c * A_CHUNK_BYTES, &tmap_q, a_full_ld, c * KCH,
                               q * (2 * BM) + (int)cta_rank * BM);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t)
        for (int c = 0; c < p.kchunks; ++c) {
          if (loaded >= (uint32_t)p.stages) {
            tc::mbar_wait(done + (rel_step & (NDONE - 1)), (rel_step / NDONE) & 1u, backoff);
            if (++rel_c == (uint32_t)p.kchunks) rel_c = 0, ++rel_step;
          }
          ++loaded;
          if (leader) {
            if (SCF_KNN_DEBUG & 128) {  // timing experiment: no reference traffic, the MMAs run on stale tiles
              if (cta_rank == 0) tc::mbar_arrive(full + s);
            } else {
              if (cta_rank == 0) tc::mbar_expect_tx(full + s, 2 * B_STAGE_BYTES);
              tc::tma_load_2d_pair(sB + (size_t)s * B_STAGE_BYTES, &tmap_r, full_ld + s * 8u, c * KCH,
                                   t * BN2 + (int)cta_rank * BM);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) s = 0, ph ^= 1u;
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_f16(2 * BM, BN2, false, false);
      const bool leader = tc::elect_one();
      const uint32_t a_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(sA));
      const uint32_t b_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(sB));
      const uint32_t nk_last = (uint32_t)p.ksteps_last, kchunks = (uint32_t)p.kchunks, stages = (uint32_t)p.stages;
      uint32_t s = 0, ph = 0, lt = 0, seg = 0;
      for (int q, t0, t1; walk.next(q, t0, t1); ++seg) {
        tc::mbar_wait(a_full, seg & 1u);
        tc::tc_fence_after();
        for (int t = t0; t < t1; ++t, ++lt) {
          const uint32_t acc = lt & 1u;
          tc::mbar_wait(tmem_empty + acc, ((lt >> 1) & 1u) ^ 1u, backoff);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)ACC_COLS;
          for (uint32_t c = 0; c < kchunks; ++c) {
            tc::mbar_wait(full + s, ph, backoff);
            tc::tc_fence_after();
            const uint32_t a_lo = a_lo0 + c * (uint32_t)(A_CHUNK_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + s * (uint32_t)(B_STAGE_BYTES >> 4);
            const uint32_t nk = c + 1 == kchunks ? nk_last : (uint32_t)(KCH / KSTEP);
#pragma unroll
            for (uint32_t kk = 0; kk < (uint32_t)(KCH / KSTEP); ++kk)
              if (kk < nk && !(SCF_KNN_DEBUG & 64)) {
                if (leader) tc::umma_f16_lo_pair(d_tmem, a_lo + kk * 2u, b_lo + kk * 2u, idesc, (c | kk) != 0u);
              }
            if (++s == stages) s = 0, ph ^= 1u;
          }
          if (leader) tc::umma_commit_pair(done + (lt & (NDONE - 1)), 3);
          __syncwarp();
        }
        if (leader) tc::umma_commit_pair(a_empty, 3);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: 64 accumulator columns of one query row per thread =====================
    const int w = warp - 2;
    const int quarter = warp & 3;
    const int g = w >> 2;              // column group: references [64 g, 64 g + 64) of the tile
    const int l = w * 32 + lane;
    const int row = quarter * 32 + lane;
    volatile float* thr_all = thr_x;   // the other three lists of this row: threads l ^ 128, l ^ 256, l ^ 384
    const int quad_id = 1 + (w & 3);
    int lt = 0;
    for (int q, t0, t1; walk.next(q, t0, t1);) {
      const int qrow = q * (2 * BM) + (int)cta_rank * BM + row;
      const int split = walk.slot_of(p, q);
      const size_t sub = (size_t)qrow * p.nlists + (size_t)(split * LISTS + g);
      Epi<KC> ep;
      ep.cl.init();
      if (qrow >= p.nq) ep.cl.thr = -FLT_MAX;
      ep.thr = ep.cl.thr;
      ep.cnt = 0;
      ep.hs = hb_score + l;
      ep.hc = hb_col + l;
      ep.gid = p.cand_idx + sub * KC;
      ep.nref = p.nref;
      thr_all[l] = ep.cl.thr;
      asm volatile("bar.sync %0, 128;" ::"r"(quad_id) : "memory");  // all four lists of a row have published
      for (int t = t0; t < t1; ++t, ++lt) {
        const int acc = lt & 1;
        if (ep.wait_draining(done + (lt & (NDONE - 1)), (uint32_t)(lt / NDONE) & 1u)) thr_all[l] = ep.cl.thr;
        tc::tc_fence_after();
        if (SCF_KNN_DEBUG & 32) {  // timing experiment: no accumulator read-out at all
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(tmem_empty_ld + (uint32_t)acc * 8u);
          continue;
        }
        const int j0 = t * BN2 + g * GCOLS;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + g * GCOLS);
        ep.thr = fminf(ep.cl.thr, partner_thr<3>(thr_all, l));
        uint32_t v[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tc::tmem_ld32(t_row + (uint32_t)(32 * hh), v);
          tc::tmem_ld_wait();
          if (SCF_KNN_DEBUG & 8) continue;
          const float m_lo = min16(v), m_hi = min16(v + 16);
          if (__any_sync(SCF_FULL, fminf(m_lo, m_hi) < ep.thr)) hits32<KC, 3>(v, m_lo, m_hi, j0 + 32 * hh, ep, thr_all, l);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(tmem_empty_ld + (uint32_t)acc * 8u);
      }
      ep.drain();
      thr_all[l] = ep.cl.thr;
#pragma unroll
      for (int e = 0; e < KC; ++e) p.cand_score[sub * KC + e] = ep.cl.r[e];
      p.cand_tau[sub] = ep.cl.thr;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();  // neither CTA leaves (or frees TMEM) while the other may still use its shared memory / barriers
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc_pair<512>(tmem_base);
  }
}

// The oracle's distance: sequential in t, separate multiply and add, in float64.  Rows are read as float4 (the row
// stride is a multiple of 4 floats and the rows are 16-byte aligned when ld % 4 == 0), four loads in flight, so the
// dependent FP64 chain does not wait for one scalar load per term.
__device__ __forceinline__ double oracle_dist(const float* __restrict__ a, const float* __restrict__ b, int dim,
                                              bool vec) {
  double acc = 0.0;
  if (vec) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const int n4 = dim >> 2;
#pragma unroll 4
    for (int t = 0; t < n4; ++t) {
      const float4 av = a4[t], bv = __ldg(b4 + t);
      double df = __dsub_rn((double)av.x, (double)bv.x);
      acc = __dadd_rn(acc, __dmul_rn(df, df));
      df = __dsub_rn((double)av.y, (double)bv.y);
      acc = __dadd_rn(acc, __dmul_rn(df, df));
      df = __dsub_rn((double)av.z, (double)bv.z);
      acc = __dadd_rn(acc, __dmul_rn(df, df));
      df = __dsub_rn((double)av.w, (double)bv.w);
      acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    for (int t = n4 << 2; t < dim; ++t) {
      const double df = __dsub_rn((double)a[t], (double)__ldg(b + t));
      acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
  } else {
    for (int t = 0; t < dim; ++t) {
      const double df = __dsub_rn((double)a[t], (double)__ldg(b + t));
      acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------- re-rank
constexpr int MAXU = 8;  // candidates per lane: nsplit * kc <= 256

__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ q, int64_t nq,
                                                         const float* __restrict__ ref, int64_t nref, int dim,
                                                         int64_t ld, int k, int64_t self_offset, int kc, int nsplit, int kp,
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
  const bool vec = (ld & 3) == 0 && ((((uintptr_t)q) | ((uintptr_t)ref)) & 15) == 0;
  // scaled units: the operands were multiplied by s, so scores / norms / eps are s^2 times the distance scale
  const double an2 = qnorm2[qi], an = sqrt(an2), bm = (double)bmax[0], dbm = (double)bmax[1], s2 = (double)bmax[2];
  const double eps = 2.0 * ((double)qerr[qi] * bm + an * dbm) * (1.0 + 1e-6) +
                     eps_acc(kp) * (2.0 * an * bm + bm * bm) + EPS_SPLIT_ABS;
  // Which candidates need the exact distance at all?  A tensor-core score is within eps of (exact scaled distance -
  // |a|^2).  Let t be the (k + 1)-th smallest score of the row's candidates (k of those k + 1 are not the query itself):
  // a candidate that scores above t + 2 eps is strictly farther than k others and cannot be among the k nearest.  The
  // FP64 distance (500 FP64 instructions per candidate at D = 100, on a part that issues 16 FP64 lanes per clock and
  // SM) is evaluated for the others only -- normally k + a few of the 64 - 128 candidates.
  int cj[MAXU];
  float cs[MAXU];
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int c = lane + 32 * u;
    cj[u] = -1, cs[u] = FLT_MAX;
    if (c < ncand) {
      const int j = cand_idx[(size_t)qi * ncand + c];
      if (j >= 0 && j < nref) cj[u] = j, cs[u] = cand_score[(size_t)qi * ncand + c];
    }
  }
  float cut = FLT_MAX;  // evaluate candidates with score <= cut
  {
    float rest[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) rest[u] = cs[u];
    float t = -FLT_MAX;
    bool full = true;
    for (int r = 0; r <= k; ++r) {  // the k + 1 smallest scores, one per round
      float m = rest[0];
#pragma unroll
      for (int u = 1; u < MAXU; ++u) m = fminf(m, rest[u]);
      float wm = m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wm = fminf(wm, __shfl_xor_sync(SCF_FULL, wm, o));
      if (!(wm < FLT_MAX)) {  // fewer than k + 1 candidates: nothing to prune
        full = false;
        break;
      }
      t = wm;
      const unsigned holders = __ballot_sync(SCF_FULL, m == wm);
      if (lane == __ffs(holders) - 1) {  // remove ONE instance (the first slot of the first lane that holds it)
        bool gone = false;
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
          if (!gone && rest[u] == wm) rest[u] = FLT_MAX, gone = true;
      }
    }
    // margins: the kept scores carry their list slot in the low mantissa bits (<= 31 ulp: 1e-5 relative covers it), and a
    // pruned candidate must stay farther after the distances are rounded to float32 (the order is (float32 distance,
    // index): a tie would be decided by the index), hence 1e-6 of the scaled distance |t| + |a|^2
    if (full)
      cut = __double2float_ru((double)t + 2.0 * eps + 1e-5 * (fabs((double)t) + 2.0 * eps) + 1e-6 * (fabs((double)t) + an2) +
                              1e-30);
  }
  unsigned long long key[MAXU];
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    key[u] = ~0ull;
    const int j = cj[u];
    if (j >= 0 && j != self && cs[u] <= cut) {
      const double acc = oracle_dist(a, ref + (int64_t)j * ld, dim, vec);
      key[u] = ((unsigned long long)__float_as_uint((float)acc) << 32) | (unsigned)j;
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
    const float dk = __uint_as_float((unsigned)(last >> 32));
    if (ok && tau < 1e29f) {  // lists were full: something was rejected, prove it is farther than the k-th kept
      const double lower = ((double)tau - eps + an2) * (1.0 - 1e-12) / s2;  // bound on any rejected exact distance
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
        const double t = (double)dk * s2 * (1.0 + 0x1p-22) - an2 + eps;
        fix_thr[slot] = enough ? __double2float_ru(t + fabs(t) * 1e-9) : FLT_MAX;
      }
    }
  }
}

// rows of the augmented query operand of the failed queries, compacted for the collect pass
__global__ void __launch_bounds__(256) knn_fix_gather_kernel(const __half* __restrict__ qop, int kp,
                                                             const int64_t* __restrict__ fail_ids,
                                                             const int* __restrict__ fail_count,
                                                             __half* __restrict__ qfix) {
  const int n = min(*fail_count, FIXTC_ROWS);
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    const __half* src = qop + fail_ids[r] * kp;
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
  const bool vec = (ld & 3) == 0 && ((((uintptr_t)q) | ((uintptr_t)ref)) & 15) == 0;
      unsigned long long key[FIXTC_CAP / 32];
      int live = 0;
#pragma unroll
      for (int u = 0; u < FIXTC_CAP / 32; ++u) {
        key[u] = ~0ull;
        const int e = lane + 32 * u;
        if (e < c) {
          const int j = fix_list[(size_t)w * FIXTC_CAP + e];
          if (j != self) {
            const double acc = oracle_dist(a, ref + (int64_t)j * ld, dim, vec);
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
  int kp, kchunks, kc, qt, bn, halves, stages, nsplit, nlists, tiles_per_split, n_ref_tiles, grid, pair, rounds, units_rem;
  long long total_work, rem_total;
  int64_t nq_pad, nr_pad;
  size_t smem, psmem;  // dynamic shared memory: single-CTA kernels / pair kernel
  int pstages;
  // workspace offsets (bytes)
  size_t off_qop, off_rop, off_qn, off_qe, off_cs, off_ci, off_tau, off_fail, off_fkey, off_misc, off_fix, off_qfix,
      off_fthr, off_fcnt, off_flist, off_rest, off_rkey, total;
};

// CTA pairs (knn_tc_pair_kernel): measured slower than the single-CTA kernel on every shape of this round (see the
// kernel's comment), so they are opt-in: SCF_KNN_PAIR = 1 selects them (the tests run both kernels on the same shapes).
bool use_pair(int) {
  const char* e = getenv("SCF_KNN_PAIR");  // read per call
  return e && atoi(e) != 0;
}

bool make_plan(int64_t nq, int64_t nref, int dim, int k, Plan& pl) {
  pl.kc = pick_kc(k);
  pl.kp = (dim + 3 + KCH - 1) / KCH * KCH;
  pl.kchunks = pl.kp / KCH;
  if (pl.kc == 0 || pl.kchunks > 3) return false;  // k > 24 or dim > 189: method 0 handles those
  pl.qt = 2;  // see knn_tc_kernel (QT = 4 / BN = 64 halves the L2 traffic but an N = 64 MMA runs at 2/3 of the rate)
  pl.bn = ACC_COLS / pl.qt;
  pl.pair = use_pair(pl.kchunks) ? 1 : 0;
  pl.halves = pl.pair ? 4 : (pl.qt == 2 ? 2 : 1);
  if (pl.pair) pl.kc = 16;  // four lists per row
  const int rows = pl.qt * BM;
  const int ref_step = pl.pair ? 2 * pl.bn : pl.bn;  // references per tile step
  pl.nq_pad = (nq + rows - 1) / rows * rows;
  pl.nr_pad = (nref + ref_step - 1) / ref_step * ref_step;
  pl.n_ref_tiles = (int)(pl.nr_pad / ref_step);
  const int64_t q_tiles = pl.nq_pad / rows;
  // Schedule (SegWalk): `rounds` whole query tiles per unit, all units in step, then the remaining query tiles cut into
  // equal ranges of tile steps over units_rem units.  A remaining tile spans at most nsplit = ceil(T / L) + 1 ranges (L =
  // steps per range) and keeps one candidate list per range (and column group); the re-rank kernel reads
  // nsplit * halves * k' <= 32 * MAXU candidates per row, which bounds units_rem from above.
  const int units_max = pl.pair ? SCF_NUM_SMS / 2 : SCF_NUM_SMS;  // CTAs (pairs of CTAs) that share the work
  const long long work = (long long)q_tiles * pl.n_ref_tiles;
  const int units = (int)std::max<long long>(1, std::min<long long>(units_max, work / (8 * 128 / pl.bn)));
  // rounds (units in step) only when the reference operand is too large to stay in L2 anyway: with a set that fits
  // (C2: 13 MB) every tile load hits L2 wherever the units are, and equal ranges over the whole space keep all lists
  // long (measured at C2: 2.53 ms without rounds, 2.62 ms with)
  bool fits_l2 = (size_t)pl.nr_pad * pl.kp * 2 <= (size_t)48 << 20;
  if (const char* e = getenv("SCF_KNN_ROUNDS")) fits_l2 = atoi(e) == 0;  // developer switch: 1 forces the rounds, 0 forbids them
  pl.rounds = fits_l2 ? 0 : (int)(q_tiles / units);
  const long long rem_tiles = q_tiles - (long long)pl.rounds * units;
  pl.rem_total = rem_tiles * pl.n_ref_tiles;
  pl.units_rem = 1, pl.nsplit = 1;
  if (rem_tiles > 0) {
    long long ur = std::max<long long>(1, std::min<long long>(units, pl.rem_total / (8 * 128 / pl.bn)));
    for (;;) {
      const long long L = pl.rem_total / ur;
      pl.nsplit = (int)((pl.n_ref_tiles + L - 1) / L) + 1;
      if (pl.nsplit * pl.halves * pl.kc <= 32 * MAXU || ur == 1) break;
      --ur;
    }
    if (ur == 1) pl.nsplit = 1;  // one unit takes whole tiles: never cut
    pl.units_rem = (int)ur;
  }
  pl.nlists = pl.nsplit * pl.halves;
  pl.grid = units;
  pl.total_work = work;
  pl.tiles_per_split = pl.n_ref_tiles;
  auto smem_for = [&](int q_tiles_in_smem, int stages) {
    return (size_t)q_tiles_in_smem * pl.kchunks * A_CHUNK_BYTES + (size_t)stages * pl.bn * 128 + (size_t)2 * HB * NL * 4 +
           (size_t)NL * 4 + (size_t)(2 + stages + 2 * NDONE + 4 + 2) * 8 + 64;
  };
  // single-CTA layout (main pass without pairs, and the collect pass): QT query tiles in shared memory
  pl.stages = 8;
  while (pl.stages > 2 && smem_for(pl.qt, pl.stages) + 1024 > 227 * 1024) --pl.stages;
  pl.smem = smem_for(pl.qt, pl.stages) + 1024;  // slack for the 1024-byte alignment of the swizzled tiles
  // a stage is released when the whole step that read it has completed: more stages than chunks per step are needed
  if (pl.smem > 227 * 1024 || pl.stages < pl.kchunks + 1) return false;
  // pair layout: one query tile per CTA
  pl.pstages = 8;
  while (pl.pstages > 2 && smem_for(1, pl.pstages) + 1024 > 227 * 1024) --pl.pstages;
  pl.psmem = smem_for(1, pl.pstages) + 1024;
  if (pl.pair && (pl.psmem > 227 * 1024 || pl.pstages < pl.kchunks + 1)) return false;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  pl.off_qop = o, o = al(o + (size_t)std::max<int64_t>(pl.nq_pad, 512) * pl.kp * 2);
  pl.off_rop = o, o = al(o + (size_t)pl.nr_pad * pl.kp * 2);
  pl.off_qn = o, o = al(o + (size_t)pl.nq_pad * 8);
  pl.off_qe = o, o = al(o + (size_t)pl.nq_pad * 4);
  pl.off_cs = o, o = al(o + (size_t)pl.nq_pad * pl.nlists * pl.kc * 4);
  pl.off_ci = o, o = al(o + (size_t)pl.nq_pad * pl.nlists * pl.kc * 4);
  pl.off_tau = o, o = al(o + (size_t)pl.nq_pad * pl.nlists * 4);
  pl.off_fail = o, o = al(o + (size_t)nq * 8);
  pl.off_fkey = o, o = al(o + (size_t)nq * 8);
  pl.off_misc = o, o = al(o + 256);
  pl.off_fix = o, o = al(o + knn_exact_fix_scratch_bytes(nq, k));
  pl.off_qfix = o, o = al(o + (size_t)FIXTC_ROWS * pl.kp * 2);
  pl.off_fthr = o, o = al(o + (size_t)FIXTC_ROWS * 4);
  pl.off_fcnt = o, o = al(o + (size_t)FIXTC_ROWS * 4);
  pl.off_flist = o, o = al(o + (size_t)FIXTC_ROWS * FIXTC_CAP * 4);
  pl.off_rest = o, o = al(o + (size_t)nq * 8);
  pl.off_rkey = o, o = al(o + (size_t)nq * 8);
  pl.total = o;
  return true;
}

typedef void (*KnnKernel)(const CUtensorMap, const CUtensorMap, const KnnTcParams);
// NI = MMA-issuing warps: two (one per query tile) when a step has two or more K chunks, else one
KnnKernel pick_kernel(int kc, int issuers, bool collect) {
  if (collect) return issuers == 2 ? (KnnKernel)knn_tc_kernel<16, 2, true, 2> : (KnnKernel)knn_tc_kernel<16, 2, true, 1>;
  if (issuers == 2) return kc == 16 ? (KnnKernel)knn_tc_kernel<16, 2, false, 2> : (KnnKernel)knn_tc_kernel<32, 2, false, 2>;
  return kc == 16 ? (KnnKernel)knn_tc_kernel<16, 2, false, 1> : (KnnKernel)knn_tc_kernel<32, 2, false, 1>;
}

}  // namespace

// measurement hook (bench.py): events the next scf_knn_l2(method 1) of this thread records around its tensor kernel
static thread_local void* g_time_start = nullptr;
static thread_local void* g_time_stop = nullptr;
extern "C" int32_t scf_knn_time_next_call(void* event_start, void* event_stop) {
  g_time_start = event_start, g_time_stop = event_stop;
  return 0;
}

int64_t knn_tc_fail_count_offset(int64_t nq, int64_t nref, int dim, int k) {
  Plan pl;
  if (!make_plan(nq, nref, dim, k, pl)) return -1;
  return (int64_t)pl.off_misc;
}

// diagnostics (scf_knn_plan): how the tensor-core path would run a shape; false when it takes method 0 instead
bool knn_tc_plan_describe(int64_t nq, int64_t nref, int dim, int k, int32_t* out) {
  Plan pl;
  if (!make_plan(nq, nref, dim, k, pl)) return false;
  const int v[16] = {pl.kc, pl.kchunks, pl.pair, pl.grid, pl.rounds, pl.units_rem, pl.nsplit, pl.nlists, pl.halves,
                     pl.n_ref_tiles, (int)(pl.nq_pad / (pl.qt * BM)), pl.pair ? pl.pstages : pl.stages,
                     pl.kchunks >= 2 ? 2 : 1, (int)((pl.pair ? pl.psmem : pl.smem) >> 10), 32 * MAXU, 0};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return true;
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
  if (!make_plan(nq, nref, dim, k, pl))  // k > 24 or dim > 189: outside the tensor-core kernel's shapes
    return knn_exact_launch(q, nullptr, nullptr, nq, ref, nref, dim, ld, ld, k, self_offset, out_idx, out_dist,
                            stream);
  if (!workspace || workspace_bytes < (int64_t)pl.total) {
    scf_set_error("scf_knn_l2: workspace too small (%lld < %zu)", (long long)workspace_bytes, pl.total);
    return 1;
  }
  unsigned char* ws = (unsigned char*)workspace;
  __half* qop = (__half*)(ws + pl.off_qop);
  __half* rop = (__half*)(ws + pl.off_rop);
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
  __half* qfix = (__half*)(ws + pl.off_qfix);
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
  float* range = bmax + 4;  // [0] max |b|^2 of the references, [1] max |a_t| of the queries (zeroed with the misc block)
  knn_range_kernel<<<4 * SCF_NUM_SMS, 256, 0, stream>>>(q, nq, ref, nref, dim, ld, range);
  int32_t rc = scf_check_launch("scf_knn_l2(range)");
  if (rc) return rc;
  knn_prep_kernel<<<4 * SCF_NUM_SMS, 256, 0, stream>>>(q, nq, pl.nq_pad, ref, nref, pl.nr_pad, dim, ld, pl.kp, qop, rop,
                                                       qn, qe, range, bmax);
  rc = scf_check_launch("scf_knn_l2(prep)");
  if (rc) return rc;
  CUtensorMap tq, tr;
  rc = scf_make_tmap_2d_f16(&tq, qop, (uint64_t)pl.nq_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BM);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f16(&tr, rop, (uint64_t)pl.nr_pad, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, (uint32_t)pl.bn);
  if (rc) return rc;
  KnnTcParams prm;
  prm.kchunks = pl.kchunks, prm.stages = pl.stages, prm.n_ref_tiles = pl.n_ref_tiles;
  prm.nq = (int)nq, prm.nref = (int)nref;
  prm.ksteps_last = ((dim + 3 + KSTEP - 1) / KSTEP) - (pl.kchunks - 1) * (KCH / KSTEP);
  int issuers = pl.kchunks >= 2 ? 2 : 1;
  {
    const char* e = getenv("SCF_KNN_ISSUERS");  // developer switch
    if (e) issuers = atoi(e) == 1 ? 1 : 2;
  }
  prm.tiles_per_split = pl.tiles_per_split, prm.nlists = pl.nlists, prm.cand_score = cs, prm.cand_idx = ci, prm.cand_tau = ctau;
  prm.balanced = 1, prm.rounds = pl.rounds, prm.units_rem = pl.units_rem, prm.rem_total = pl.rem_total;
  // list slots that no range fills (a query tile that is not cut uses one of its nsplit slots): ids -1, tau "nothing rejected"
  e = cudaMemsetAsync(ci, 0xFF, (size_t)pl.nq_pad * pl.nlists * pl.kc * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctau, 0x7F, (size_t)pl.nq_pad * pl.nlists * 4, stream);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  {
    const char* f = getenv("SCF_KNN_FLAGS");
    prm.flags = f ? atoi(f) : 2;
  }
  prm.fail_count = nullptr, prm.fix_thr = nullptr, prm.fix_cnt = nullptr, prm.fix_list = nullptr;
  KnnKernel kern = pl.pair ? (KnnKernel)knn_tc_pair_kernel<16> : pick_kernel(pl.kc, issuers, false);
  const size_t main_smem = pl.pair ? pl.psmem : pl.smem;
  if (pl.pair) prm.stages = pl.pstages;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  if (g_time_start) cudaEventRecord((cudaEvent_t)g_time_start, stream);
  if (pl.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)pl.grid), cfg.blockDim = dim3(NTHREADS_PAIR);
    cfg.dynamicSmemBytes = main_smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, tq, tr, prm);
    if (e != cudaSuccess) {
      scf_set_error("scf_knn_l2(pair launch): %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
  } else {
    kern<<<dim3((unsigned)pl.grid), issuers == 2 ? NTHREADS : NTHREADS - 32, pl.smem, stream>>>(tq, tr, prm);
  }
  if (g_time_stop) cudaEventRecord((cudaEvent_t)g_time_stop, stream);
  g_time_start = g_time_stop = nullptr;  // one shot
  rc = scf_check_launch("scf_knn_l2(tcgen05)");
  if (rc) return rc;
  if (SCF_KNN_DEBUG & (4 | 8 | 32 | 64 | 128)) return 0;  // timing experiments: candidates are meaningless, stop here
  knn_rerank_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, stream>>>(q, nq, ref, nref, dim, ld, k, self_offset, pl.kc,
                                                                  pl.nlists, pl.kp, cs, ci, ctau, qn, qe, bmax, out_idx, out_dist,
                                                                  fail_ids, fail_keys, fix_thr, fail_count);
  rc = scf_check_launch("scf_knn_l2(rerank)");
  if (rc) return rc;
  if (SCF_KNN_DEBUG & 16) {
    unsigned long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_dbg, sizeof(h));
    fprintf(stderr, "[knn dbg] qt %d kc %d nlists %d | warp-steps %llu, hit halves %llu (%.1f%% of steps), hit groups %llu, "
            "drains %llu, drain iterations %llu, buffered entries %llu, inserts %llu (%.1f / query)\n",
            pl.qt, pl.kc, pl.nlists, h[0], h[1], 100.0 * h[1] / (double)std::max(1ull, h[0]), h[2], h[3], h[4], h[5], h[6],
            (double)h[6] / (double)nq);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(g_dbg, h, sizeof(h));
  }
  // ---- repair of the rows whose guard could not be proven: tensor-core collect pass, then FP64 on the few left ----
  knn_fix_gather_kernel<<<SCF_NUM_SMS, 256, 0, stream>>>(qop, pl.kp, fail_ids, fail_count, qfix);
  rc = scf_check_launch("scf_knn_l2(fix,gather)");
  if (rc) return rc;
  CUtensorMap tf;
  rc = scf_make_tmap_2d_f16(&tf, qfix, (uint64_t)FIXTC_ROWS, (uint64_t)pl.kp, (uint64_t)pl.kp, KCH, BM);
  if (rc) return rc;
  KnnTcParams fp = prm;
  fp.nq = FIXTC_ROWS;
  fp.balanced = 0;  // collect pass: static (query tile, reference split) grid
  fp.stages = pl.stages;
  fp.n_ref_tiles = (int)(pl.nr_pad / pl.bn);  // tiles of 128 references (the pair kernel walks tiles of 256)
  int fsplit = std::max(1, std::min(FIXTC_NSPLIT, fp.n_ref_tiles));
  fp.tiles_per_split = (fp.n_ref_tiles + fsplit - 1) / fsplit;
  fsplit = (fp.n_ref_tiles + fp.tiles_per_split - 1) / fp.tiles_per_split;
  fp.fail_count = fail_count, fp.fix_thr = fix_thr, fp.fix_cnt = fix_cnt, fp.fix_list = fix_list;
  KnnKernel fkern = pick_kernel(16, issuers, true);
  e = cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_knn_l2: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  fkern<<<dim3(FIXTC_ROWS / (pl.qt * BM), (unsigned)fsplit), issuers == 2 ? NTHREADS : NTHREADS - 32, pl.smem, stream>>>(tf, tr,
                                                                                                                     fp);
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
