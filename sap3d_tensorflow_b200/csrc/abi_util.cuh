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

}  // namespace sap3d
