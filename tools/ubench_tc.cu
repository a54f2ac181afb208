// Developer microbenchmark (B200): what the tcgen05 accumulator read-out and the MMA issue rate really are, alone and
// together.  Settles the "TMEM read-out floor" of round 1 (the 64 B/clk/SM figure of B300_MICROARCH.md was
// beaten by the kernel it was supposed to bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scarf_b200/csrc tools/ubench_tc.cu -o tools/build/ubench_tc
// Every kernel runs one CTA per SM (grid 148); cycles are clock64() around the measured loop, max over the warps of
// CTA 0 .. 147 (reported: median over CTAs).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

namespace {

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
  float m = __uint_as_float(v[0]);
#pragma unroll
  for (int i = 1; i < 32; ++i) m = fminf(m, __uint_as_float(v[i]));
  return m;
}

// mode bit 0: readers on; bit 1: min tree on the loaded values; bit 2: MMA issue on; bit 3: two loads in flight
// W reader warps: warp w reads lane quarter (w & 3), column group (w >> 2) of W / 4 groups over the 512 columns.
template <int W, int N>
__global__ void __launch_bounds__(32 * W + 64, 1) ubench(int mode, int reps_ld, int reps_mma, long long* cyc_ld,
                                                        long long* cyc_mma, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;               // 128 rows x 128 B
  unsigned char* sB = smem + 16384;       // N rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_ld[W];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (mode & 4) {
      constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, false, false);
      const bool leader = tc::elect_one();
      const uint64_t da = tc::umma_desc_k_sw128_u32(tc::smem_u32(sA));
      const uint64_t db = tc::umma_desc_k_sw128_u32(tc::smem_u32(sB));
      const long long t0 = clock64();
      for (int r = 0; r < reps_mma; ++r) {
        const uint32_t d = tmem_base + (uint32_t)((r & 1) * (N == 256 ? 256 : 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) tc::umma_f16(d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, kk != 0);
        __syncwarp();
      }
      if (leader) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (lane == 0) cyc_mma[blockIdx.x] = t1 - t0;
    }
  } else if (warp >= 2) {
    const int w = warp - 2;
    if (mode & 1) {
      const int quarter = warp & 3;  // TMEM lane quarter of a warp = warp id % 4
      constexpr int NG = W / 4;
      const int grp = w >> 2;
      constexpr int SPAN = 512 / NG;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * SPAN);
      float acc = 3e38f;
      const long long t0 = clock64();
      if (mode & 8) {
        uint32_t va[32], vb[32];
        for (int r = 0; r < reps_ld; r += 2) {
          tc::tmem_ld32(t_row + (uint32_t)((r * 32) % SPAN), va);
          tc::tmem_ld32(t_row + (uint32_t)(((r + 1) * 32) % SPAN), vb);
          tc::tmem_ld_wait();
          if (mode & 2) acc = fminf(acc, fminf(min32(va), min32(vb)));
        }
      } else {
        uint32_t v[32];
        for (int r = 0; r < reps_ld; ++r) {
          tc::tmem_ld32(t_row + (uint32_t)((r * 32) % SPAN), v);
          tc::tmem_ld_wait();
          if (mode & 2) acc = fminf(acc, min32(v));
        }
      }
      const long long t1 = clock64();
      if (lane == 0) t_ld[w] = t1 - t0;
      if (acc == 1.2345f) sink[0] = acc;
    } else if (lane == 0) {
      t_ld[w] = 0;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int i = 0; i < W; ++i) m = t_ld[i] > m ? t_ld[i] : m;
    cyc_ld[blockIdx.x] = m;
  }
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// The issue pattern of knn_tc_kernel: descriptors advance with a stage counter (uniform arithmetic), a commit after
// every `per_commit` instructions, nothing else running.  variant 0: 64-bit descriptors built per instruction,
// variant 1: 32-bit low words + constant high word (umma_f16_lo).
template <int N>
__global__ void __launch_bounds__(64, 1) ubench_issue(int variant, int reps, int per_commit, int stages,
                                                      long long* cyc_mma) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, sink_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 6 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&sink_bar, 1 << 20);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, false, false);
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(smem));
    const uint32_t b_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(smem + 16384));
    uint32_t s = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += per_commit) {
      const uint32_t b_lo = b_lo0 + s * (uint32_t)(N * 128 >> 4);
      const uint32_t d = tmem_base + (uint32_t)((r / per_commit) & 1) * 256u;
      for (int kk = 0; kk < per_commit; ++kk) {
        if (variant == 0) {
          const uint64_t da = ((uint64_t)tc::UMMA_DESC_HI_K_SW128 << 32) | (a_lo0 + (uint32_t)(kk & 3) * 2u);
          const uint64_t db = ((uint64_t)tc::UMMA_DESC_HI_K_SW128 << 32) | (b_lo + (uint32_t)(kk & 3) * 2u);
          if (leader) tc::umma_f16(d + (uint32_t)(kk >> 2) * N % 256u, da, db, idesc, (kk & 3) != 0);
        } else {
          if (leader)
            tc::umma_f16_lo(d + (uint32_t)(kk >> 2) * N % 256u, a_lo0 + (uint32_t)(kk & 3) * 2u,
                            b_lo + (uint32_t)(kk & 3) * 2u, idesc, (kk & 3) != 0);
        }
      }
      if (leader) tc::umma_commit(&sink_bar);
      __syncwarp();
      if (++s == (uint32_t)stages) s = 0;
    }
    if (leader) tc::umma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) cyc_mma[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// Two warps issue alternately-sized streams of their own (different accumulator columns), each with its own commits:
// does a commit stall only the warp that issued it (then two issuers hide each other's stalls) or the tensor pipe?
template <int N>
__global__ void __launch_bounds__(96, 1) ubench_issue2(int reps, int per_commit, int issuers, long long* cyc_mma) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2], sink_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end[2];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 6 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_init(&sink_bar, 1 << 20);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async();
  if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const long long t0 = clock64();
  if (warp < issuers) {
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, false, false);
    const bool leader = tc::elect_one();
    const uint32_t a_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(smem));
    const uint32_t b_lo0 = tc::umma_desc_lo_k_sw128(tc::smem_u32(smem + 16384));
    const uint32_t d = tmem_base + (uint32_t)warp * 256u;
    const int mine = reps / issuers;
    for (int r = 0; r < mine; r += per_commit) {
      for (int kk = 0; kk < per_commit; ++kk)
        if (leader)
          tc::umma_f16_lo(d + (uint32_t)(kk >> 2) * N % 256u, a_lo0 + (uint32_t)(kk & 3) * 2u, b_lo0 + (uint32_t)(kk & 3) * 2u,
                          idesc, (kk & 3) != 0);
      if (leader) tc::umma_commit(&sink_bar);
      __syncwarp();
    }
    if (leader) tc::umma_commit(&bar[warp]);
    __syncwarp();
    tc::mbar_wait(&bar[warp], 0);
    if (lane == 0) t_end[warp] = clock64();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) cyc_mma[blockIdx.x] = (issuers == 2 ? max(t_end[0], t_end[1]) : t_end[0]) - t0;
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int N>
void run_issue2(int per_commit, int issuers) {
  const int reps = 8192;
  long long* c_mma;
  CK(cudaMalloc(&c_mma, 148 * 8));
  CK(cudaMemset(c_mma, 0, 148 * 8));
  const int smem = 16384 + 6 * N * 128 + 1024;
  CK(cudaFuncSetAttribute(ubench_issue2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int it = 0; it < 2; ++it) {
    ubench_issue2<N><<<148, 96, smem>>>(reps, per_commit, issuers, c_mma);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), c_mma, 148 * 8, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  printf("issue pattern: N=%3d, %d issuing warp(s), commit every %2d | %.1f cyc per MMA\n", N, issuers, per_commit,
         (double)h[74] / reps);
  CK(cudaFree(c_mma));
}

template <int N>
void run_issue(int variant, int per_commit) {
  const int reps = 8192;
  long long* c_mma;
  CK(cudaMalloc(&c_mma, 148 * 8));
  CK(cudaMemset(c_mma, 0, 148 * 8));
  const int smem = 16384 + 6 * N * 128 + 1024;
  CK(cudaFuncSetAttribute(ubench_issue<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int it = 0; it < 2; ++it) {
    ubench_issue<N><<<148, 64, smem>>>(variant, reps, per_commit, 6, c_mma);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), c_mma, 148 * 8, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  printf("issue pattern: N=%3d variant %d (%s), commit every %2d | %.1f cyc per MMA\n", N, variant,
         variant ? "32-bit descriptor words" : "64-bit descriptors", per_commit, (double)h[74] / reps);
  CK(cudaFree(c_mma));
}

template <int W, int N>
void run(int mode, const char* label) {
  const int reps_ld = 4096, reps_mma = 2048;
  long long *c_ld, *c_mma;
  float* sink;
  CK(cudaMalloc(&c_ld, 148 * 8));
  CK(cudaMalloc(&c_mma, 148 * 8));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(c_ld, 0, 148 * 8));
  CK(cudaMemset(c_mma, 0, 148 * 8));
  const int smem = 16384 + N * 128 + 1024;
  CK(cudaFuncSetAttribute(ubench<W, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) {
    CK(cudaEventRecord(e0));
    ubench<W, N><<<148, 32 * W + 64, smem>>>(mode, reps_ld, reps_mma, c_ld, c_mma, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h_ld(148), h_mma(148);
  CK(cudaMemcpy(h_ld.data(), c_ld, 148 * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_mma.data(), c_mma, 148 * 8, cudaMemcpyDeviceToHost));
  std::sort(h_ld.begin(), h_ld.end());
  std::sort(h_mma.begin(), h_mma.end());
  const double cl = (double)h_ld[74], cm = (double)h_mma[74];
  printf("%-44s W=%2d N=%3d mode=%2d | %.3f ms", label, W, N, mode, ms);
  if (mode & 1)
    printf(" | ld: %.0f cyc, %.1f cyc per x32 load per warp, %.1f B/clk/SM", cl, cl / reps_ld,
           (double)W * reps_ld * 4096.0 / cl);
  if (mode & 4)
    printf(" | mma: %.0f cyc, %.1f cyc per M128xN%dxK16 (%.0f MAC/clk/SM)", cm, cm / (reps_mma * 4.0), N,
           128.0 * N * 16.0 * reps_mma * 4.0 / cm);
  printf("\n");
  CK(cudaFree(c_ld));
  CK(cudaFree(c_mma));
  CK(cudaFree(sink));
}

}  // namespace

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  if (getenv("UBENCH_ISSUE2")) {
    for (int pc : {4, 8, 14, 32})
      for (int is = 1; is <= 2; ++is) run_issue2<128>(pc, is);
    return 0;
  }
  for (int v = 0; v < 2; ++v) {
    run_issue<64>(v, 16);
    run_issue<128>(v, 8);
    run_issue<128>(v, 4);
    run_issue<256>(v, 4);
  }
  run<8, 64>(4, "MMA only, SS operands");
  run<4, 128>(1, "TMEM read only, load+wait");
  run<8, 128>(1, "TMEM read only, load+wait");
  run<16, 128>(1, "TMEM read only, load+wait");
  run<4, 128>(1 | 8, "TMEM read only, 2 loads in flight");
  run<8, 128>(1 | 8, "TMEM read only, 2 loads in flight");
  run<16, 128>(1 | 8, "TMEM read only, 2 loads in flight");
  run<4, 128>(1 | 2, "TMEM read + min tree");
  run<8, 128>(1 | 2, "TMEM read + min tree");
  run<16, 128>(1 | 2, "TMEM read + min tree");
  run<8, 128>(1 | 2 | 8, "TMEM read + min tree, 2 in flight");
  run<16, 128>(1 | 2 | 8, "TMEM read + min tree, 2 in flight");
  run<8, 128>(4, "MMA only, SS operands");
  run<8, 256>(4, "MMA only, SS operands");
  run<8, 128>(1 | 2 | 4, "MMA + TMEM read + min tree");
  run<16, 128>(1 | 2 | 4, "MMA + TMEM read + min tree");
  run<8, 256>(1 | 2 | 4, "MMA + TMEM read + min tree");
  run<16, 256>(1 | 2 | 4, "MMA + TMEM read + min tree");
  return 0;
}
