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

DeviceInfo* device_info() {
  static DeviceInfo infos[64];
  int dev = 0;
  if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
  if (dev < 0 || dev >= 64) {
    fail(UP_ERR_UNSUPPORTED, "device index %d out of range", dev);
    return nullptr;
  }
  DeviceInfo& di = infos[dev];
  if (di.sm_count == 0) {
    cudaDeviceProp prop;
    if (check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return nullptr;
    if (prop.major != 10) {
      fail(UP_ERR_UNSUPPORTED, "unipose_b200 needs an sm_100 class GPU (found sm_%d%d)", prop.major, prop.minor);
      return nullptr;
    }
    di.max_smem = prop.sharedMemPerBlockOptin;
    di.sm_count = prop.multiProcessorCount;
  }
  return &di;
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
