// Dense row block -> CSR: the adapter between the reference's on-disk count matrix (dense uint32 N x G in Zarr chunks,
// scarf/writers.py:164-204) and the CSR the path computes on (what Assay.to_raw_sparse builds per chunk on the CPU,
// scarf/assay.py:175-199).  Two streaming passes over the block, warp per row, 16-byte loads, four in flight per lane:
// pass 1 counts the stored values of every row, the caller turns the counts into offsets (exclusive scan), pass 2
// writes (column, count) pairs in ascending column order.  HBM-bound: 4 * ld bytes per row and pass.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ uint4 ld_block4(const uint32_t* row, int64_t v, int64_t nvec) {
  if (v >= nvec) return make_uint4(0u, 0u, 0u, 0u);
  const int4 x = ld_stream4(reinterpret_cast<const int4*>(row) + v);
  return make_uint4((uint32_t)x.x, (uint32_t)x.y, (uint32_t)x.z, (uint32_t)x.w);
}

// columns >= n_cols of the last vector are padding of the staging buffer: masked, whatever they hold
__device__ __forceinline__ uint32_t nz_mask(const uint4& x, int64_t v, int32_t n_cols) {
  const int64_t c = 4 * v;
  return (uint32_t)(x.x != 0u && c < n_cols) | ((uint32_t)(x.y != 0u && c + 1 < n_cols) << 1) |
         ((uint32_t)(x.z != 0u && c + 2 < n_cols) << 2) | ((uint32_t)(x.w != 0u && c + 3 < n_cols) << 3);
}

__global__ void __launch_bounds__(kThreads) dense_row_nnz_kernel(const uint32_t* __restrict__ dense, int64_t n_rows,
                                                                 int32_t n_cols, int64_t ld,
                                                                 int64_t* __restrict__ row_nnz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t nvec = (n_cols + 3) / 4;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const uint32_t* row = dense + r * ld;
    int cnt = 0;
    for (int64_t v0 = 0; v0 < nvec; v0 += 32 * kUnroll) {
      uint4 x[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) x[u] = ld_block4(row, v0 + 32 * u + lane, nvec);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) cnt += __popc(nz_mask(x[u], v0 + 32 * u + lane, n_cols));
    }
    cnt = warp_sum(cnt);
    if (lane == 0) row_nnz[r] = cnt;
  }
}

__global__ void __launch_bounds__(kThreads) dense_to_csr_kernel(const uint32_t* __restrict__ dense, int64_t n_rows,
                                                                int32_t n_cols, int64_t ld,
                                                                const int64_t* __restrict__ row_ptr,
                                                                int32_t* __restrict__ indices,
                                                                uint32_t* __restrict__ data) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t nvec = (n_cols + 3) / 4;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const uint32_t* row = dense + r * ld;
    int64_t base = row_ptr[r];
    for (int64_t v0 = 0; v0 < nvec; v0 += 32 * kUnroll) {
      uint4 x[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) x[u] = ld_block4(row, v0 + 32 * u + lane, nvec);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + 32 * u + lane;
        const uint32_t m = nz_mask(x[u], v, n_cols);
        if (__any_sync(SCF_FULL, m != 0u)) {  // most 128-column stretches of a count matrix hold something
          const int c = __popc(m);
          int incl = c;  // inclusive scan of the per-lane counts: lanes keep ascending column order
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(SCF_FULL, incl, o);
            if (lane >= o) incl += t;
          }
          int64_t p = base + incl - c;
          const uint32_t vals[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (m & (1u << j)) {
              indices[p] = (int32_t)(4 * v + j);
              data[p] = vals[j];
              ++p;
            }
          base += __shfl_sync(SCF_FULL, incl, 31);
        }
      }
    }
  }
}

int grid_for(const void* kernel) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  return SCF_NUM_SMS * (per_sm < 1 ? 1 : per_sm);
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" int32_t scf_dense_row_nnz(const uint32_t* dense, int64_t n_rows, int32_t n_cols, int64_t ld,
                                     int64_t* row_nnz, void* stream) {
  SCF_ARG(dense && row_nnz, "null pointer");
  SCF_ARG(n_rows >= 0 && n_cols > 0 && ld >= n_cols && (ld & 3) == 0 && aligned16(dense),
          "bad sizes (ld must be a multiple of 4 and the block 16-byte aligned)");
  if (n_rows == 0) return 0;
  dense_row_nnz_kernel<<<grid_for((const void*)dense_row_nnz_kernel), kThreads, 0, (cudaStream_t)stream>>>(
      dense, n_rows, n_cols, ld, row_nnz);
  return scf_check_launch("scf_dense_row_nnz");
}

extern "C" int32_t scf_dense_to_csr(const uint32_t* dense, int64_t n_rows, int32_t n_cols, int64_t ld,
                                    const int64_t* row_ptr, int32_t* indices, uint32_t* data, void* stream) {
  SCF_ARG(dense && row_ptr && indices && data, "null pointer");
  SCF_ARG(n_rows >= 0 && n_cols > 0 && ld >= n_cols && (ld & 3) == 0 && aligned16(dense),
          "bad sizes (ld must be a multiple of 4 and the block 16-byte aligned)");
  if (n_rows == 0) return 0;
  dense_to_csr_kernel<<<grid_for((const void*)dense_to_csr_kernel), kThreads, 0, (cudaStream_t)stream>>>(
      dense, n_rows, n_cols, ld, row_ptr, indices, data);
  return scf_check_launch("scf_dense_to_csr");
}
