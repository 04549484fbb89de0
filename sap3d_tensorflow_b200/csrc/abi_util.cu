#include "abi_util.cuh"

#include "../../include/sap3d.h"

namespace sap3d {

static thread_local char g_err[1024] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int require_device() {
  static thread_local int ok = -1;
  if (ok == 1) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  if (major != 10) return set_error("device compute capability %d.x is not sm_100 (B200 required)", major);
  ok = 1;
  return 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace sap3d

extern "C" {
const char* sap3d_last_error(void) { return sap3d::g_err; }
int sap3d_abi_version(void) { return 1; }
int sap3d_device_ok(void) { return sap3d::require_device() == 0 ? 1 : 0; }
}
