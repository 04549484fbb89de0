// Error reporting and device checks shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

namespace sap3d {

// thread-local message returned by sap3d_last_error(); returns 1 so callers can `return set_error(...)`
int set_error(const char* fmt, ...);
// 0 when a usable sm_100 device is current; otherwise sets the error and returns 1 (no CPU fallback)
int require_device();
int check_launch(const char* what);

// SAP3D_PDL=0 switches programmatic dependent launch off (every launch is then fully serialised, as in round 1)
bool pdl_enabled();

// kernel launch with optional thread-block cluster and the programmatic-stream-serialisation attribute (kernels launched
// through this helper call pdl_wait() before they touch anything a predecessor may have written)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace sap3d
