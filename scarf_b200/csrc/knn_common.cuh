#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// method 0 (knn_exact.cu); q_ids (nullable) = subset of query rows to (re)compute in place
// n_ids_dev (nullable) = device-side length of q_ids (then nq is only the upper bound used to size the grid);
// ld / ldr = row strides of q / ref
int32_t knn_exact_launch(const float* q, const int64_t* q_ids, const int* n_ids_dev, int64_t nq, const float* ref,
                         int64_t nref, int dim, int64_t ld, int64_t ldr, int k, int64_t self_offset, int64_t* out_idx,
                         float* out_dist, cudaStream_t stream);
// fail-list repair used by method 1: rows q_ids[0 .. *n_ids_dev) are recomputed exactly.  q_keys[s] = (float32
// distance bits << 32 | index) of row s's current k-th candidate (~0 = unknown): a threshold scan over the
// references collects every key <= it; up to KNN_FIX_CAP rows are repaired that way, the rest (and overflowing
// rows) by the FP64 tile kernel.
constexpr int KNN_FIX_CAP = 16384;
size_t knn_exact_fix_scratch_bytes(int64_t nq, int k);
int32_t knn_exact_fix_launch(const float* q, const int64_t* q_ids, const unsigned long long* q_keys,
                             const int* n_ids_dev, int64_t nq_max, const float* ref, int64_t nref, int dim, int64_t ld,
                             int k, int64_t self_offset, int64_t* out_idx, float* out_dist, void* scratch,
                             cudaStream_t stream);
// method 1 (knn_tc.cu)
int64_t knn_tc_workspace_bytes(int64_t nq, int64_t nref, int dim, int k);
int64_t knn_tc_fail_count_offset(int64_t nq, int64_t nref, int dim, int k);
bool knn_tc_plan_describe(int64_t nq, int64_t nref, int dim, int k, int32_t* out);
int32_t knn_tc_launch(const float* q, int64_t nq, const float* ref, int64_t nref, int dim, int64_t ld, int k,
                      int64_t self_offset, int64_t* out_idx, float* out_dist, void* workspace,
                      int64_t workspace_bytes, cudaStream_t stream);
