// Error plumbing of the C-ABI (include/scarf_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void scf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t scf_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    scf_set_error("%s: %s", what, cudaGetErrorString(e));
    return -(int32_t)e;
  }
  return 0;
}

extern "C" int32_t scf_version(void) { return SCF_VERSION; }
extern "C" const char* scf_last_error(void) { return g_err; }

// ---- TMA descriptor helper shared by the tcgen05 kernels ----
#include "tc_common.cuh"
static int32_t make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, uint32_t esize,
                            uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows);
int32_t scf_make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_2d(out, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cols, ld, box_cols, box_rows);
}
int32_t scf_make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_2d(out, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rows, cols, ld, box_cols, box_rows);
}
static int32_t make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, uint32_t esize,
                            uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  static scf_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      scf_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
      return e != cudaSuccess ? -(int32_t)e : -999;
    }
    fn = (scf_encode_tiled_fn)p;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    scf_set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return -(int32_t)r - 10000;
  }
  return 0;
}

// 3-D view {32 floats, rows, ld / 32 column blocks} of a row-major float32 matrix [rows, ld]: a box
// {32, box_rows, box_blocks} lands in shared memory as box_blocks x [box_rows][128 B] (128-byte swizzle with
// 32-byte atoms), the MN-major tf32 operand layout of tcgen05.mma (gram_tc.cu).
int32_t scf_make_tmap_colblocks_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t ld,
                                    uint32_t box_rows, uint32_t box_blocks) {
  scf_encode_tiled_fn fn = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      scf_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
      return e != cudaSuccess ? -(int32_t)e : -999;
    }
    fn = (scf_encode_tiled_fn)p;
  }
  cuuint64_t dims[3] = {32, rows, ld / 32};
  cuuint64_t strides[2] = {ld * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, box_blocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    scf_set_error("cuTensorMapEncodeTiled(3d) failed (%d): rows=%llu ld=%llu box=%ux%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)ld, box_rows, box_blocks);
    return -(int32_t)r - 10000;
  }
  return 0;
}
