// Error plumbing and device queries of the C-ABI (include/unipose_b200.h).
#include <string.h>

#include "up_internal.h"

namespace up {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(UP_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

}  // namespace up

extern "C" const char* up_last_error(void) { return up::g_err; }

extern "C" int up_version(void) { return UP_VERSION; }

extern "C" int up_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  int rc = up::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  cudaDeviceProp prop;
  rc = up::check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
  if (rc) return rc;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
