// K4 on tensor cores: Y[:, :dims] = Z @ V  (AnnStream.reducer, scarf/ann.py:138) as a 3xTF32 tcgen05 GEMM.
//
// The FP32 SIMT kernel (gemm_simt.cu) is bound by FMA issue (2*H*D FLOP per cell at 128 FMA/clk/SM), five to six times
// above the HBM time of reading Z.  Here one CTA owns 128 cells: Z and its TF32 remainder Z_lo (the two planes the
// Gram kernel reads) stream through a TMA pipeline as K-major operands, V^T is split once into TF32 hi / lo planes,
// and every K step issues hi*hi + lo*hi + hi*lo (M128 x N=round_up(dims,16) x K8).  Consecutive K chunks rotate over
// four TMEM accumulators that the epilogue adds in FP32 round-to-nearest: the tensor core's accumulator truncates, and
// four short chains keep that bias at ~1e-5 relative (one long chain: ~5e-5).  HBM-bound: 8 B per Z element.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int PM = 128;     // cells per CTA (TMEM lanes)
constexpr int PKC = 32;     // float32 per 128-byte swizzle row = one K chunk
constexpr int NACC = 4;     // accumulators the K chunks rotate over
constexpr int PTHREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue

struct ProjParams {
  int n_chunks;   // ceil(n_cols / 32)
  int n;          // MMA N = round_up(dims, 16)
  int dims;
  int stages;
  int64_t n_rows, ldy;
  float* y;
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// vt_hi / vt_lo [n][ldz]: transposed loadings split into the part the tensor core reads (top 19 bits) and the rest
__global__ void project_split_kernel(const float* __restrict__ v, int64_t ldv, int n_cols, int dims, int n, int64_t ldz,
                                     float* __restrict__ vt_hi, float* __restrict__ vt_lo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * ldz) return;
  const int d = (int)(t / ldz);
  const int64_t j = t - (int64_t)d * ldz;
  float x = 0.f;
  if (d < dims && j < n_cols) x = v[j * ldv + d];
  const float hi = tf32_hi(x);
  vt_hi[t] = hi;
  vt_lo[t] = x - hi;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(PTHREADS, 1) project_tc_kernel(const __grid_constant__ CUtensorMap tmap_z,
                                                                 const __grid_constant__ CUtensorMap tmap_zlo,
                                                                 const __grid_constant__ CUtensorMap tmap_vhi,
                                                                 const __grid_constant__ CUtensorMap tmap_vlo,
                                                                 const ProjParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = PM * 128, b_bytes = (uint32_t)p.n * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;  // Z hi | Z lo | V^T hi | V^T lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* done = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * PM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_z);
    tc::tma_prefetch_desc(&tmap_zlo);
    tc::tma_prefetch_desc(&tmap_vhi);
    tc::tma_prefetch_desc(&tmap_vlo);
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    const bool leader = tc::elect_one();
    for (int c = 0; c < p.n_chunks; ++c) {
      const int s = c % p.stages;
      const uint32_t ph = (uint32_t)(c / p.stages) & 1u;
      tc::mbar_wait(empty + s, ph ^ 1u, 32);
      if (leader) {
        unsigned char* base = smem + (size_t)s * stage_bytes;
        tc::mbar_expect_tx(full + s, stage_bytes);
        tc::tma_load_2d(base, &tmap_z, full + s, c * PKC, row0);
        tc::tma_load_2d(base + a_bytes, &tmap_zlo, full + s, c * PKC, row0);
        tc::tma_load_2d(base + 2 * a_bytes, &tmap_vhi, full + s, c * PKC, 0);
        tc::tma_load_2d(base + 2 * a_bytes + b_bytes, &tmap_vlo, full + s, c * PKC, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::umma_idesc_tf32(PM, p.n, false, false);
    const uint32_t smem_u = tc::smem_u32(smem);
    for (int c = 0; c < p.n_chunks; ++c) {
      const int s = c % p.stages;
      const uint32_t ph = (uint32_t)(c / p.stages) & 1u;
      tc::mbar_wait(full + s, ph);
      tc::tc_fence_after();
      const uint32_t base = smem_u + (uint32_t)s * stage_bytes;
      const uint64_t da = tc::umma_desc_k_sw128_u32(base);
      const uint64_t dal = tc::umma_desc_k_sw128_u32(base + a_bytes);
      const uint64_t db = tc::umma_desc_k_sw128_u32(base + 2 * a_bytes);
      const uint64_t dbl = tc::umma_desc_k_sw128_u32(base + 2 * a_bytes + b_bytes);
      const uint32_t d_tmem = tmem_base + (uint32_t)((c % NACC) * p.n);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < PKC / 8; ++kk) {  // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
          const uint64_t adv = (uint64_t)(kk * 2);
          tc::umma_tf32(d_tmem, da + adv, db + adv, idesc, (c >= NACC || kk > 0) ? 1u : 0u);
          tc::umma_tf32(d_tmem, dal + adv, db + adv, idesc, 1u);
          tc::umma_tf32(d_tmem, da + adv, dbl + adv, idesc, 1u);
        }
        tc::umma_commit(empty + s);
      }
      __syncwarp();
    }
    if (leader) tc::umma_commit(done);
    __syncwarp();
  } else {
    // ===================== epilogue: one cell per thread =====================
    const int quarter = warp & 3;
    const int64_t row = (int64_t)row0 + quarter * 32 + lane;
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    const int nacc = min(NACC, p.n_chunks);
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* yr = p.y + row * p.ldy;
    for (int c0 = 0; c0 < p.n; c0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      for (int a = 0; a < nacc; ++a) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)(a * p.n + c0), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
      }
      if (row < p.n_rows) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int col = c0 + j;
          if (col < p.ldy) {
            float4 o;
            o.x = col + 0 < p.dims ? acc[j + 0] : 0.f;
            o.y = col + 1 < p.dims ? acc[j + 1] : 0.f;
            o.z = col + 2 < p.dims ? acc[j + 2] : 0.f;
            o.w = col + 3 < p.dims ? acc[j + 3] : 0.f;
            *reinterpret_cast<float4*>(yr + col) = o;
          }
        }
      }
    }
    // pad columns of y beyond the MMA width
    if (row < p.n_rows)
      for (int64_t col = p.n; col < p.ldy; col += 4) *reinterpret_cast<float4*>(yr + col) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

int round_up16(int x) { return (x + 15) / 16 * 16; }

}  // namespace

extern "C" int64_t scf_project_tc_workspace_bytes(int64_t ldz, int32_t dims) {
  return 2 * (int64_t)round_up16(dims) * ldz * 4;
}

extern "C" int32_t scf_project_tc(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows, int32_t n_cols,
                                  const float* v, int64_t ldv, int32_t dims, float* y, int64_t ldy, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  SCF_ARG(z && z_lo && v && y && workspace, "null pointer");
  SCF_ARG(n_rows >= 0 && n_cols > 0 && dims > 0 && dims <= 128 && ldz >= n_cols && ldv >= dims && ldy >= dims, "bad sizes");
  SCF_ARG((ldz & 31) == 0 && (ldy & 3) == 0, "ldz must be a multiple of 32 and ldy of 4");
  SCF_ARG((((uintptr_t)z | (uintptr_t)z_lo | (uintptr_t)y | (uintptr_t)workspace) & 15) == 0, "z, z_lo, y, workspace must be 16-byte aligned");
  SCF_ARG(workspace_bytes >= scf_project_tc_workspace_bytes(ldz, dims), "workspace too small");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = round_up16(dims);
  float* vt_hi = (float*)workspace;
  float* vt_lo = vt_hi + (size_t)n * ldz;
  const int64_t total = (int64_t)n * ldz;
  project_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v, ldv, n_cols, dims, n, ldz, vt_hi, vt_lo);
  int32_t rc = scf_check_launch("scf_project_tc(split)");
  if (rc) return rc;
  CUtensorMap tz, tzl, tvh, tvl;
  rc = scf_make_tmap_2d_f32(&tz, z, (uint64_t)n_rows, (uint64_t)ldz, (uint64_t)ldz, PKC, PM);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f32(&tzl, z_lo, (uint64_t)n_rows, (uint64_t)ldz, (uint64_t)ldz, PKC, PM);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f32(&tvh, vt_hi, (uint64_t)n, (uint64_t)ldz, (uint64_t)ldz, PKC, (uint32_t)n);
  if (rc) return rc;
  rc = scf_make_tmap_2d_f32(&tvl, vt_lo, (uint64_t)n, (uint64_t)ldz, (uint64_t)ldz, PKC, (uint32_t)n);
  if (rc) return rc;
  ProjParams p;
  p.n_chunks = (n_cols + PKC - 1) / PKC;
  p.n = n, p.dims = dims, p.n_rows = n_rows, p.ldy = ldy, p.y = y;
  const size_t stage_bytes = (size_t)2 * PM * 128 + (size_t)2 * n * 128;
  p.stages = 4;
  while (p.stages > 2 && p.stages * stage_bytes + 256 + 1024 > 227 * 1024) --p.stages;
  const size_t smem = p.stages * stage_bytes + (size_t)(2 * p.stages + 1) * 8 + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scf_set_error("scf_project_tc: %s", cudaGetErrorString(e));
    return -(int32_t)e;
  }
  project_tc_kernel<<<(unsigned)((n_rows + PM - 1) / PM), PTHREADS, smem, st>>>(tz, tzl, tvh, tvl, p);
  return scf_check_launch("scf_project_tc");
}
