#include "abi_util.cuh"
#include "conv_tc.cuh"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace sap3d

namespace {
// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8 on the host: the checksum TensorFlow's tensor-bundle
// checkpoints carry per tensor and per index block
struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xffu];
  }
};
}  // namespace

extern "C" {
uint32_t sap3d_crc32c(uint32_t crc, const void* data, size_t n) {
  static const Crc32cTables T;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu]; --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
        T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu];
  return ~c;
}
const char* sap3d_last_error(void) { return sap3d::g_err; }
int sap3d_abi_version(void) { return 2; }
// developer probe (tools/conv_phase_probe.py): per-CTA phase time stamps of conv_tc_kernel into `buf` ([cta][16][2] uint64,
// device memory); NULL switches it off.  Process-global and NOT part of the re-entrant surface.
int sap3d_debug_conv_timing(void* buf) {
  sap3d::tc_set_debug_buffer(buf);
  return 0;
}
long long sap3d_debug_conv_halo_launches(void) { return sap3d::tc_halo_launches(); }
long long sap3d_debug_conv_swap_launches(void) { return sap3d::tc_swap_launches(); }
int sap3d_device_ok(void) { return sap3d::require_device() == 0 ? 1 : 0; }
}
