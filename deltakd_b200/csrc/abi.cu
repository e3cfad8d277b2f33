// Error plumbing, version and device check of the C ABI (include/deltakd.h).
#include <stdarg.h>

#include "common.cuh"

namespace dkd {
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels enqueued by this library (all threads; statistics only)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DKD_E_LAUNCH;
  }
  return DKD_OK;
}
}  // namespace dkd

extern "C" {

int dkd_version(void) { return 100; }

const char* dkd_last_error(void) { return dkd::g_err; }

unsigned long long dkd_launch_count(void) { return __atomic_load_n(&dkd::g_launches, __ATOMIC_RELAXED); }

int dkd_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    dkd::set_error("no CUDA device: libdeltakd_sm100 has no CPU path");
    return DKD_E_ARCH;
  }
  if (major != 10) {
    dkd::set_error("device compute capability %d.x is not sm_100: libdeltakd_sm100 is B200-only", major);
    return DKD_E_ARCH;
  }
  return DKD_OK;
}

}  // extern "C"
