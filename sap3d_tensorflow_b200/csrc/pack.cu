// One-launch re-packing of ALL convolution filters after an optimizer step:
// fp32 TF-layout master weights -> bf16 K-major operand matrices of the tensor-core kernels
// (forward B operand [cout_pad][taps*cin] and data-gradient B operand [cin_pad][taps*cout]).
// A device-resident table describes every (source, destination) pair; entries are multiples of 64
// elements, so 8-element chunks never straddle entries.
#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

__global__ void __launch_bounds__(256) pack_multi_kernel(const sap3d_pack_entry* __restrict__ tab, int n, long long total) {
  const long long nchunk = total / 8;
  for (long long ch = blockIdx.x * (long long)blockDim.x + threadIdx.x; ch < nchunk; ch += (long long)gridDim.x * blockDim.x) {
    const long long e0 = ch * 8;
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].start <= e0) lo = mid;
      else hi = mid - 1;
    }
    const sap3d_pack_entry en = tab[lo];
    if (en.cols % 8 != 0) {              // odd inner extents (stem: cin = 3): element-wise indexing
      const long long i = e0 - en.start;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long idx = i + j;
        const int c = (int)(idx % en.cols);
        const long long t = idx / en.cols;
        const int tap = (int)(t % en.taps);
        const int r = (int)(t / en.taps);
        v[j] = r < en.rows ? __ldg(en.src + tap * en.s_tap + r * en.s_r + c * en.s_c) : 0.f;
      }
      Vec8<bf16>::store(reinterpret_cast<bf16*>(en.dst) + i, v);
      continue;
    }
    long long q = (e0 - en.start) / 8;   // chunk within the entry
    const int c8n = en.cols / 8;
    int r, tap, c;
    if (en.s_c == 1) {                   // source contiguous along the destination's inner axis: chunk order = destination order
      c = (int)(q % c8n) * 8; q /= c8n;
      tap = (int)(q % en.taps);
      r = (int)(q / en.taps);
    } else {                             // transposing entry (s_r == 1): consecutive threads walk the SOURCE-contiguous axis r,
      r = (int)(q % en.rows_pad); q /= en.rows_pad;   // so every one of the 8 loads is coalesced across the warp
      c = (int)(q % c8n) * 8;
      tap = (int)(q / c8n);
    }
    float v[8];
    const float* src = en.src + tap * en.s_tap + r * en.s_r + c * en.s_c;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = r < en.rows ? __ldg(src + j * en.s_c) : 0.f;
    Vec8<bf16>::store(reinterpret_cast<bf16*>(en.dst) + ((long long)r * en.taps + tap) * en.cols + c, v);
  }
}

}  // namespace

extern "C" int sap3d_pack_multi(const sap3d_pack_entry* entries_dev, int32_t n, int64_t total, void* stream) {
  if (require_device()) return 1;
  if (n <= 0) return 0;
  if (total % 8 != 0) return set_error("pack_multi: total must be a multiple of 8");
  long long blocks = (total / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_multi_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(entries_dev, n, total);
  return check_launch("pack_multi");
}
