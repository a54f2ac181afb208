// placeholder until the tcgen05 kernel lands
#include "common.cuh"
int32_t scf_gram_tc(const float*, int64_t, int64_t, int32_t, int64_t*, int64_t, int32_t, cudaStream_t) {
  scf_set_error("scf_gram_accumulate: tcgen05 modes not built");
  return 2;
}
