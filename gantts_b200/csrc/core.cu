// Error reporting, version, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace gantts {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return GANTTS_E_CUDA;
}
}  // namespace gantts

extern "C" int gantts_version(void) { return 101; }

extern "C" const char* gantts_last_error_string(void) { return gantts::g_err; }

extern "C" int gantts_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}
