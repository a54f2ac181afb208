// Device-side robust LOWESS: the same algorithm as scf_host_lowess (host_lowess.cu), run by ONE CTA so that
// mark_hvgs' trend removal (scarf/feat_utils.py:22,38-40 -> statsmodels lowess, frac 0.1, it 100, delta 0) needs no
// device->host round trip.  <= 512 points (the reference bins the genes into 200); the arithmetic
// uses the round-to-nearest intrinsics of the host routine (no FMA contraction); window sums are formed by four lanes
// (see below), so the two fits agree to ~1e-13 relative rather than bit for bit.
#include "lowess_dev.cuh"

namespace {

constexpr int kMaxN = kLowessMaxN;
constexpr int kThreads = kLowessThreads;

__global__ void __launch_bounds__(kThreads) lowess_kernel(const double* __restrict__ endog,
                                                          const double* __restrict__ exog,
                                                          const uint8_t* __restrict__ valid, int n_in, double frac, int it,
                                                          int cache_tri, double* __restrict__ out) {
  extern __shared__ double dyn[];
  lowess_block(endog, exog, valid, n_in, frac, it, cache_tri, out, dyn);
}

}  // namespace

extern "C" int32_t scf_lowess(const double* endog, const double* exog, const uint8_t* valid, int32_t n, double frac,
                              int32_t it, double* out, void* stream) {
  SCF_ARG(endog && exog && out, "null pointer");
  SCF_ARG(n >= 1 && n <= kMaxN && it >= 0, "n must be within [1, 512]");
  // shared-memory cache of the tricube and normalised weights: 2 * n * k doubles with k <= frac * n + 1
  const int64_t kmax = (int64_t)(frac * (double)n + 1e-10) + 1;
  size_t dyn = (size_t)2 * n * (kmax > 0 ? kmax : 1) * sizeof(double);
  int cache = 1;
  if (dyn > 160 * 1024) dyn = 0, cache = 0;
  if (dyn > 20 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lowess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) {
      scf_set_error("scf_lowess: %s", cudaGetErrorString(e));
      return -(int32_t)e;
    }
  }
  lowess_kernel<<<1, kThreads, dyn, (cudaStream_t)stream>>>(endog, exog, valid, n, frac, it, cache, out);
  return scf_check_launch("scf_lowess");
}
