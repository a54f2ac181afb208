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
