// K2 on the 5th-generation tensor cores: G += Z^T Z, TF32 x TF32 -> FP32 in TMEM, int64 fixed point in HBM.
//
//   Z is row-major [n_rows, ldz] float32 (cells x features), so BOTH operands are "MN-major": the M / N index
//   (a feature) is contiguous in memory and K (the cell) strides by ldz.  One 3-D TMA box {32 features, KR cells,
//   4 feature blocks} lands in shared memory as 4 blocks of [KR cells][128 B] with the 32-byte-atom 128-byte
//   swizzle, which is the canonical MN-major SWIZZLE_128B_BASE32B operand layout of tcgen05.mma -- the only one
//   it takes for MN-major tf32 (LBO = KR*128 B between feature blocks, 4-cell atoms of 512 B, two per K step).
//
//   work item   = (128 x 256 output tile of the upper triangle, slab of SCF_GRAM_SLAB cells)
//   CTA         = one tile x every nsplit-th slab; warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM),
//                 warps 2-5 = epilogue.  Two 256-column accumulators in TMEM: the epilogue of slab i overlaps
//                 the MMAs of slab i+1.
//   epilogue    = tcgen05.ld -> x 2^SCF_GRAM_SHIFT -> int64 -> shared staging rows -> one bulk reduce-add
//                 (cp.reduce.async.bulk .add.u64) per 256-byte row segment into g_fx.  Integer adds commute, so G
//                 is bit-identical for any slab-to-CTA / slab-to-GPU assignment (slab partials are FP32 sums in a
//                 fixed order).
//   mode 3      = 3xTF32: Z = hi + lo with hi = the 19 bits the tensor core keeps and lo = Z - hi (a second
//                 plane written by scf_csr_norm_scale); G ~ hi'hi + hi'lo + lo'hi, error ~2^-20 per product.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int GM = 128;          // tile rows   (features, TMEM lanes)
constexpr int GN = 256;          // tile columns (features, TMEM columns)
constexpr int STAGES = 3;
constexpr int NTHREADS = 192;    // 6 warps
constexpr int STG_STRIDE = 272;  // bytes per staging row: 32 int64 + 16 B pad (conflict-free 16-byte stores)
constexpr int STG_BYTES = 32 * STG_STRIDE;

template <int MODE>
struct Cfg {
  static constexpr int PLANES = MODE == 3 ? 2 : 1;
  static constexpr int KR = MODE == 3 ? 16 : 32;  // cells per pipeline stage
  static constexpr int A_BYTES = GM * KR * 4;
  static constexpr int B_BYTES = GN * KR * 4;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 4 * 2 * STG_BYTES + (2 * STAGES + 4) * 8 + 64 + 1024;
};

struct GramParams {
  int64_t n_rows;
  int n_cols;
  int n_mtiles;  // ceil(n_cols / 128)
  int n_ntiles;  // ceil(n_cols / 256)
  int n_slabs;
  long long* g_fx;
  int64_t ldg;
  int flags;  // developer switches (SCF_GRAM_FLAGS): 1 = per-element atomics instead of bulk reduce-add
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(m), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_red_add_u64(long long* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(gdst),
               "r"(tc::smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap tmap_hi,
                                                              const __grid_constant__ CUtensorMap tmap_lo,
                                                              const GramParams p) {
  using C = Cfg<MODE>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;                                       // [STAGES][A_hi | B_hi | A_lo | B_lo]
  unsigned char* staging = stages + (size_t)STAGES * C::STAGE_BYTES;  // [4 warps][2][32 rows][272 B]
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + 4 * 2 * STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // upper-triangle tile: row block mi (128 features) x column block nj (256 features), nj >= mi / 2
  int t = blockIdx.x, mi = 0;
  while (t >= p.n_ntiles - (mi >> 1)) {
    t -= p.n_ntiles - (mi >> 1);
    ++mi;
  }
  const int nj = (mi >> 1) + t;
  const int split = blockIdx.y, nsplit = gridDim.y;
  const int my_slabs = split < p.n_slabs ? (p.n_slabs - split + nsplit - 1) / nsplit : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tmem_full + a, 1);
      tc::mbar_init(tmem_empty + a, 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_hi);
    if (MODE == 3) tc::tma_prefetch_desc(&tmap_lo);
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int STAGES_PER_SLAB = (SCF_GRAM_SLAB + C::KR - 1) / C::KR;

  if (warp == 0) {
    // ===================== TMA producer =====================  (whole warp, one elected lane issues)
    {
      const bool leader = tc::elect_one();
      int it = 0;
      for (int i = 0; i < my_slabs; ++i) {
        const int64_t r0 = (int64_t)(split + i * nsplit) * SCF_GRAM_SLAB;
        for (int st = 0; st < STAGES_PER_SLAB; ++st, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          tc::mbar_wait(empty + s, ph ^ 1u, 32);
          unsigned char* base = stages + (size_t)s * C::STAGE_BYTES;
          const int row = (int)(r0 + (int64_t)st * C::KR);
          if (leader) {
            tc::mbar_expect_tx(full + s, C::STAGE_BYTES);
            // one request = 4 feature blocks (128 features) x KR cells; B (256 features) takes two
            tma_load_3d(base, &tmap_hi, full + s, 0, row, mi * 4);
            tma_load_3d(base + C::A_BYTES, &tmap_hi, full + s, 0, row, nj * 8);
            tma_load_3d(base + 2 * C::A_BYTES, &tmap_hi, full + s, 0, row, nj * 8 + 4);
            if (MODE == 3) {
              unsigned char* lo = base + C::A_BYTES + C::B_BYTES;
              tma_load_3d(lo, &tmap_lo, full + s, 0, row, mi * 4);
              tma_load_3d(lo + C::A_BYTES, &tmap_lo, full + s, 0, row, nj * 8);
              tma_load_3d(lo + 2 * C::A_BYTES, &tmap_lo, full + s, 0, row, nj * 8 + 4);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp, warp-uniform values, one elected lane issues: descriptors stay in uniform registers (issuing from an
    // `if (lane == 0)` region costs a vector->uniform waterfall per instruction, see knn_tc.cu)
    {
      constexpr uint32_t idesc = tc::umma_idesc_tf32(GM, GN, true, true);
      constexpr uint32_t LBO = C::KR * 128;  // bytes between 32-feature blocks
      const bool leader = tc::elect_one();
      const uint32_t stages_u = tc::smem_u32(stages);
      int it = 0;
      for (int i = 0; i < my_slabs; ++i) {
        const int acc = i & 1;
        const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
        tc::mbar_wait(tmem_empty + acc, acc_ph ^ 1u, 32);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * GN);
        for (int st = 0; st < STAGES_PER_SLAB; ++st, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          tc::mbar_wait(full + s, ph);
          tc::tc_fence_after();
          const uint32_t base = stages_u + (uint32_t)s * C::STAGE_BYTES;
          const uint64_t da = tc::umma_desc_mn_sw128_32b_u32(base, LBO, 512);
          const uint64_t db = tc::umma_desc_mn_sw128_32b_u32(base + C::A_BYTES, LBO, 512);
          const uint64_t dal = tc::umma_desc_mn_sw128_32b_u32(base + C::A_BYTES + C::B_BYTES, LBO, 512);
          const uint64_t dbl = tc::umma_desc_mn_sw128_32b_u32(base + 2 * C::A_BYTES + C::B_BYTES, LBO, 512);
          // cells of this stage that still belong to the slab (SCF_GRAM_SLAB is a multiple of 8)
          const int rows_left = SCF_GRAM_SLAB - st * C::KR;
          const int nk = (rows_left < C::KR ? rows_left : C::KR) / 8;
#pragma unroll
          for (int kk = 0; kk < C::KR / 8; ++kk) {
            if (kk < nk && leader) {
              const uint64_t adv = (uint64_t)(kk * (1024 >> 4));  // 8 cells = two 512-B atoms per K step
              tc::umma_tf32(d_tmem, da + adv, db + adv, idesc, (st | kk) != 0);
              if (MODE == 3) {
                tc::umma_tf32(d_tmem, da + adv, dbl + adv, idesc, 1u);
                tc::umma_tf32(d_tmem, dal + adv, db + adv, idesc, 1u);
              }
            }
          }
          if (leader) tc::umma_commit(empty + s);
          __syncwarp();
        }
        if (leader) tc::umma_commit(tmem_full + acc);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue =====================
    const int qd = warp & 3;  // TMEM lane quarter this warp may read
    const int m = mi * GM + qd * 32 + lane;
    unsigned char* stg = staging + (size_t)(warp - 2) * 2 * STG_BYTES;
    const float scale = (float)(1ull << SCF_GRAM_SHIFT);
    int nbuf = 0;
    for (int i = 0; i < my_slabs; ++i) {
      const int acc = i & 1;
      const uint32_t acc_ph = (uint32_t)(i >> 1) & 1u;
      tc::mbar_wait(tmem_full + acc, acc_ph);
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * GN);
#pragma unroll 1
      for (int c = 0; c < GN / 32; ++c) {
        const int n0 = nj * GN + c * 32;
        // warp-uniform skip: block entirely below the diagonal or outside the matrix
        if (n0 + 31 < mi * GM + qd * 32 || n0 >= p.n_cols || mi * GM + qd * 32 >= p.n_cols) continue;
        uint32_t v[32];
        tc::tmem_ld32(t_row + (uint32_t)(c * 32), v);
        unsigned char* buf = stg + (size_t)(nbuf & 1) * STG_BYTES + (size_t)lane * STG_STRIDE;
        bulk_wait_read<1>();  // the bulk reduce issued two chunks ago has finished reading this buffer
        tc::tmem_ld_wait();
        long long last_val = 0;
        const int ncols_here = min(32, p.n_cols - n0);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          long long x[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int n = n0 + j + u;
            const bool keep = m <= n && m < p.n_cols && n < p.n_cols;
            x[u] = keep ? __float2ll_rn(__uint_as_float(v[j + u]) * scale) : 0ll;
            if (j + u == ncols_here - 1) last_val = x[u];
          }
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(tc::smem_u32(buf + j * 8)), "l"(x[0]), "l"(x[1])
                       : "memory");
        }
        if (p.flags & 1) {
#pragma unroll 1
          for (int j = 0; j < ncols_here; ++j) {
            const long long xv = *reinterpret_cast<const long long*>(buf + j * 8);
            if (m < p.n_cols && xv != 0) atomicAdd((unsigned long long*)(p.g_fx + (int64_t)m * p.ldg + n0 + j), (unsigned long long)xv);
          }
          ++nbuf;
          continue;
        }
        tc::fence_proxy_async();
        if (m < p.n_cols) {
          long long* dst = p.g_fx + (int64_t)m * p.ldg + n0;
          const uint32_t bytes = (uint32_t)(ncols_here & ~1) * 8u;
          if (bytes) bulk_red_add_u64(dst, buf, bytes);
          if (ncols_here & 1) atomicAdd((unsigned long long*)(dst + ncols_here - 1), (unsigned long long)last_val);
        }
        bulk_commit();
        ++nbuf;
      }
      tc::tc_fence_before();
      tc::mbar_arrive(tmem_empty + acc);
    }
    bulk_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

__global__ void gram_mirror_kernel(long long* __restrict__ g, int n, int64_t ldg) {
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  __shared__ long long tile[32][33];
  // lower(r, c) = upper(c, r): transpose 32x32 blocks of the strictly-upper part into the lower part
  if (blockIdx.x < blockIdx.y) return;  // handle each (block row <= block col) pair once
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i;
    tile[i][threadIdx.x] = (r < n && c < n) ? g[(int64_t)r * ldg + c] : 0ll;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int rr = blockIdx.x * 32 + i, cc = r0 + threadIdx.x;  // (rr, cc) = transpose position
    if (rr < n && cc < n && rr > cc) g[(int64_t)rr * ldg + cc] = tile[threadIdx.x][i];
  }
}

}  // namespace

// 3-D view of a row-major float32 matrix [rows, ld]: {32 floats, rows, ld / 32 blocks}, box {32, box_rows, box_blocks}
int32_t scf_make_tmap_colblocks_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t ld,
                                    uint32_t box_rows, uint32_t box_blocks);

int32_t scf_gram_tc(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows, int32_t n_cols, int64_t* g_fx,
                    int64_t ldg, int32_t mode, cudaStream_t stream) {
  if ((ldz & 31) != 0) {
    scf_set_error("scf_gram_accumulate: tensor-core modes need ldz %% 32 == 0 (got %lld)", (long long)ldz);
    return 1;
  }
  if ((ldg & 1) != 0 || ((uintptr_t)g_fx & 15) != 0 || ((uintptr_t)z & 15) != 0) {
    scf_set_error("scf_gram_accumulate: tensor-core modes need an even ldg and 16-byte aligned z / g_fx");
    return 1;
  }
  if (mode == 3 && !z_lo) {
    scf_set_error("scf_gram_accumulate: mode 3 (3xTF32) needs the z_lo plane written by scf_csr_norm_scale");
    return 1;
  }
  if (n_rows >= 2147483647LL) {
    scf_set_error("scf_gram_accumulate: n_rows must fit int32 (split the rows)");
    return 1;
  }
  GramParams p;
  p.n_rows = n_rows, p.n_cols = n_cols, p.g_fx = (long long*)g_fx, p.ldg = ldg;
  p.n_mtiles = (n_cols + GM - 1) / GM;
  p.n_ntiles = (n_cols + GN - 1) / GN;
  p.n_slabs = (int)((n_rows + SCF_GRAM_SLAB - 1) / SCF_GRAM_SLAB);
  {
    const char* f = getenv("SCF_GRAM_FLAGS");
    p.flags = f ? atoi(f) : 0;
  }
  int n_tiles = 0;
  for (int mi = 0; mi < p.n_mtiles; ++mi) n_tiles += p.n_ntiles - (mi >> 1);
  int nsplit = SCF_NUM_SMS / n_tiles;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > p.n_slabs) nsplit = p.n_slabs;
  CUtensorMap th, tl;
  int32_t rc;
  const int kr = mode == 3 ? Cfg<3>::KR : Cfg<1>::KR;
  // one tensor map per plane with a 4-block (128-feature) box: A is one request per stage, B two
  rc = scf_make_tmap_colblocks_f32(&th, z, (uint64_t)n_rows, (uint64_t)ldz, (uint32_t)kr, 4);
  if (rc) return rc;
  if (mode == 3) {
    rc = scf_make_tmap_colblocks_f32(&tl, z_lo, (uint64_t)n_rows, (uint64_t)ldz, (uint32_t)kr, 4);
    if (rc) return rc;
  } else {
    tl = th;
  }
  dim3 grid((unsigned)n_tiles, (unsigned)nsplit);
  cudaError_t e;
  if (mode == 3) {
    e = cudaFuncSetAttribute(gram_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<3>::SMEM);
    if (e == cudaSuccess) gram_tc_kernel<3><<<grid, NTHREADS, Cfg<3>::SMEM, stream>>>(th, tl, p);
  } else {
    e = cudaFuncSetAttribute(gram_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<1>::SMEM);
    if (e == cudaSuccess) gram_tc_kernel<1><<<grid, NTHREADS, Cfg<1>::SMEM, stream>>>(th, tl, p);
  }
  if (e != cudaSuccess) {
    scf_set_error("scf_gram_accumulate: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  return scf_check_launch("scf_gram_accumulate(tcgen05)");
}

int32_t scf_gram_mirror(int64_t* g_fx, int32_t n_cols, int64_t ldg, cudaStream_t stream) {
  const unsigned nb = (unsigned)((n_cols + 31) / 32);
  gram_mirror_kernel<<<dim3(nb, nb), dim3(32, 8), 0, stream>>>((long long*)g_fx, n_cols, ldg);
  return scf_check_launch("scf_gram_symmetrize");
}
