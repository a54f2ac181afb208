// K5 entry points (include/scarf_b200.h).
#include "common.cuh"
#include "knn_common.cuh"

extern "C" int64_t scf_knn_workspace_bytes(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t method) {
  if (method == 1) return knn_tc_workspace_bytes(nq, nref, dim, k);
  return 0;
}

extern "C" int64_t scf_knn_fail_count_offset(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t method) {
  return method == 1 ? knn_tc_fail_count_offset(nq, nref, dim, k) : -1;
}

extern "C" int32_t scf_knn_plan(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t* out16) {
  SCF_ARG(out16, "null pointer");
  SCF_ARG(nq > 0 && nref > 0 && dim > 0 && k > 0, "bad sizes");
  return knn_tc_plan_describe(nq, nref, dim, k, out16) ? 0 : 1;
}

extern "C" int32_t scf_knn_l2(const float* q, int64_t nq, const float* ref, int64_t nref, int32_t dim, int64_t ld,
                              int32_t k, int64_t self_offset, int64_t* out_idx, float* out_dist, int32_t method,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  SCF_ARG(q && ref && out_idx && out_dist, "null pointer");
  SCF_ARG(nq >= 0 && nref > 0 && dim > 0 && ld >= dim && k > 0, "bad sizes");
  SCF_ARG(nref < 2147483647LL, "nref must fit int32");
  SCF_ARG((self_offset >= 0 ? nref - 1 : nref) >= k, "k exceeds the number of eligible references");
  SCF_ARG(self_offset < 0 || self_offset + nq <= nref, "self_offset + nq exceeds nref");
  SCF_ARG(method == 0 || method == 1, "method must be 0 (fp64 simt) or 1 (tcgen05 + re-rank)");
  if (nq == 0) return 0;
  if (method == 0)
    return knn_exact_launch(q, nullptr, nullptr, nq, ref, nref, dim, ld, ld, k, self_offset, out_idx, out_dist,
                            (cudaStream_t)stream);
  return knn_tc_launch(q, nq, ref, nref, dim, ld, k, self_offset, out_idx, out_dist, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
