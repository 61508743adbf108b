// du_abi.cu — error plumbing and device queries of the C ABI (include/du_b200.h).
#include <cstdarg>
#include <cstdio>

#include "du_common.cuh"

namespace du {
static thread_local char g_err[512] = "";
static thread_local int g_want_device = -1;   // -1: whatever device is current (single-GPU processes never call du_set_device)
int wanted_device() { return g_want_device; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  return set_error(DU_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
}
}  // namespace du

extern "C" const char* du_last_error(void) { return du::g_err; }
extern "C" int du_version(void) { return 1; }
extern "C" int du_num_sms(int device) {
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return du::check_cuda(e, "cudaDeviceGetAttribute");
  return n;
}
extern "C" int du_set_device(int device) {
  if (device < 0) { du::g_want_device = -1; return DU_OK; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return du::check_cuda(e, "cudaGetDeviceCount");
  if (device >= count) return du::set_error(DU_ERR_BAD_ARG, "du_set_device: device %d out of range (%d devices)", device, count);
  du::g_want_device = device;
  return DU_OK;
}
