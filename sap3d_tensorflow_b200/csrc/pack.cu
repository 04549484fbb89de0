// One-launch re-packing of ALL convolution filters after an optimizer step:
// fp32 TF-layout master weights -> bf16 K-major operand matrices of the tensor-core kernels
// (forward B operand [cout_pad][taps*cin] and data-gradient B operand [cin_pad][taps*cout]).
// A device-resident table describes every (source, destination) pair; the caller lays the entries out in
// a concatenated index space with every `start` (and the total) a multiple of 4096 elements.
#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

constexpr int PACK_MAX_CACHED = 1024;
constexpr int PACK_SPAN = 4096;   // elements per block iteration; the caller aligns every entry's `start` to it

// one 8-element chunk of an entry (chunk index q within the entry): the r01 form, still used for entries whose inner extent
// is not a multiple of 64 and for the non-transposing entries (already coalesced on both sides)
__device__ __forceinline__ void pack_chunk(const sap3d_pack_entry& en, long long q) {
  if (en.cols % 8 != 0) {              // odd inner extents (stem: cin = 3): element-wise indexing
    const long long i = q * 8;
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
    return;
  }
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
  if (en.s_c == 1 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {   // contiguous source: two 16-byte loads instead of eight scalar ones
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < en.rows) {
      a = __ldg(reinterpret_cast<const float4*>(src));
      b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = r < en.rows ? __ldg(src + j * en.s_c) : 0.f;
  }
  Vec8<bf16>::store(reinterpret_cast<bf16*>(en.dst) + ((long long)r * en.taps + tap) * en.cols + c, v);
}

// One block iteration = one PACK_SPAN of the concatenated index space.  Transposing entries (TF [tap][cin][cout] -> forward
// operand [cout][tap][cin]) whose extents are multiples of 64 go through a 64 x 64 shared-memory tile: 256-byte source rows in,
// whole 128-byte destination lines out (r01/r02a: 16-byte stores to 32 different rows per warp; ncu 1.02 GB moved for 0.68 GB).
__global__ void __launch_bounds__(256) pack_multi_kernel(const sap3d_pack_entry* __restrict__ tab, int n, long long total) {
  __shared__ float tile[64][65];
  __shared__ long long s_start[PACK_MAX_CACHED];   // the entries' starts: the per-span binary search then never leaves the SM
  const int tid = threadIdx.x;
  const bool cached = n <= PACK_MAX_CACHED;
  if (cached) {
    for (int i = tid; i < n; i += 256) s_start[i] = tab[i].start;
    __syncthreads();
  }
  const long long nspan = total / PACK_SPAN;
  for (long long span = blockIdx.x; span < nspan; span += gridDim.x) {
    const long long e0 = span * PACK_SPAN;
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((cached ? s_start[mid] : tab[mid].start) <= e0) lo = mid;
      else hi = mid - 1;
    }
    const sap3d_pack_entry en = tab[lo];
    const long long size = (long long)en.rows_pad * en.taps * en.cols;
    const long long rel = e0 - en.start;
    if (rel >= size) continue;         // alignment gap behind an entry
    if (en.s_r == 1 && en.s_c != 1 && en.cols % 64 == 0 && en.rows_pad % 64 == 0 && en.rows % 4 == 0) {
      const int c_tiles = en.cols / 64, r_tiles = en.rows_pad / 64;
      long long q = rel / PACK_SPAN;
      const int ct = (int)(q % c_tiles); q /= c_tiles;
      const int rt = (int)(q % r_tiles);
      const int tap = (int)(q / r_tiles);
      const int r0 = rt * 64 + (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cl = (tid >> 4) + 16 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 < en.rows) v = __ldg(reinterpret_cast<const float4*>(en.src + tap * en.s_tap + r0 + (long long)(ct * 64 + cl) * en.s_c));
        float* t = &tile[cl][(tid & 15) * 4];
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      }
      __syncthreads();
      const int rl = tid >> 2, cq = (tid & 3) * 16;
      bf16* dst = reinterpret_cast<bf16*>(en.dst) + ((long long)(rt * 64 + rl) * en.taps + tap) * en.cols + ct * 64 + cq;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tile[cq + h * 8 + j][rl];
        Vec8<bf16>::store(dst + h * 8, v);
      }
      __syncthreads();
    } else {
#pragma unroll
      for (int k = 0; k < PACK_SPAN / 8 / 256; ++k) {
        const long long q = rel / 8 + tid + k * 256;
        if (q * 8 < size) pack_chunk(en, q);
      }
    }
  }
}

}  // namespace

extern "C" int sap3d_pack_multi(const sap3d_pack_entry* entries_dev, int32_t n, int64_t total, void* stream) {
  if (require_device()) return 1;
  if (n <= 0) return 0;
  if (total % PACK_SPAN != 0) return set_error("pack_multi: total (and every entry's start) must be a multiple of 4096");
  long long blocks = total / PACK_SPAN;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_multi_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(entries_dev, n, total);
  return check_launch("pack_multi");
}
