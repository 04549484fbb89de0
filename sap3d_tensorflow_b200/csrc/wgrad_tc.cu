// tcgen05 / TMEM filter-gradient kernel for sm_100a.
//
//   dW_tap[m][n] += sum over positions  P(pos + tap offset)[m] * Q(pos)[n]
//
// (conv: P = input activations, Q = dy; transposed conv: P = dy parity view, Q = input.)  The
// reduction dimension is the POSITION axis, so both MMA operands are "MN-major": a TMA box of
// (bn x bd x bh x bw) positions x 64 channels lands in shared memory as [position][64 ch] rows of
// 128 bytes (128-byte swizzle), which is exactly the canonical MN-major SWIZZLE_128B UMMA layout
// (K = position index: 8-row groups 1024 B apart; 64-channel blocks one box apart).  TMA zero-fill
// of out-of-bounds coordinates implements the 'SAME' padding of the shifted operand.
//
// One CTA owns one [128 x BLOCK_N] block of one tap's dW and a contiguous slice of the position
// tiles (split-K); partial results are reduced into the fp32 gradient with red.global.add.f32.
// Replaces cuDNN's conv3d backward-filter kernels behind tf.gradients (train.py:168).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "conv_tc.cuh"

namespace sap3d {

constexpr int WG_MAX_MAPS = 28;
constexpr int WG_MAX_TAPS = 32;
constexpr int WG_THREADS = 192;

struct WgTap {
  int8_t map, dw, dh, dd;
  int32_t pad;
  long long dw_ofs;
};

struct alignas(64) WgParams {
  CUtensorMap pmap[WG_MAX_MAPS];
  CUtensorMap qmap;
  WgTap taps[WG_MAX_TAPS];
  int ntaps, mblocks, nblocks, splits;
  int m_tiles;
  int tiles[4], box[4];
  int box_rows;
  int M, N;        // valid channels of P / Q
  int p_c0;        // first channel (multiple of 64) of P consumed
  long long ldw;
  float* dw;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  constexpr int P_BYTES = 2 * 16384;                 // 128 channels of P: two 64-channel boxes
  constexpr int Q_BYTES = (BLOCK_N / 64) * 16384;
  constexpr int STAGE_BYTES = P_BYTES + Q_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode: (split, (mblock, nblock, tap)), tap fastest ----------------------------------
  // CTAs that are resident together (consecutive blockIdx) then cover ALL taps of the SAME position slice: the dy tile is
  // shared by every tap and the shifted x tiles overlap, so each wave streams its slice of x and dy from DRAM once and the
  // other taps hit L2.  (With the split index fastest every wave walked the whole of x and dy -- 3 waves = 3.0x the
  // algorithmic DRAM bytes in the r01/r02 ncu captures.)
  int w = blockIdx.x;
  const int tap_id = w % p.ntaps;
  w /= p.ntaps;
  const int nb = w % p.nblocks;
  w /= p.nblocks;
  const int mb = w % p.mblocks;
  const int split = w / p.mblocks;
  const WgTap tap = p.taps[tap_id];
  const int per = (p.m_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * per;
  const int t_end = min(p.m_tiles, t_begin + per);
  const int ntiles = max(0, t_end - t_begin);
  const int ksteps = (p.box_rows + 15) / 16;

  // rows >= box_rows of every operand block are never written by TMA: zero them once so the
  // K (= position) padding contributes nothing
  if (p.box_rows < 128) {
    const int row_bytes0 = p.box_rows * 128;
    for (int s = 0; s < STAGES; ++s)
      for (int blk = 0; blk < 2 + BLOCK_N / 64; ++blk) {
        uint8_t* b0 = smem + s * STAGE_BYTES + blk * 16384;
        for (int i = row_bytes0 + threadIdx.x * 16; i < 16384; i += WG_THREADS * 16) *reinterpret_cast<uint4*>(b0 + i) = make_uint4(0, 0, 0, 0);
      }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (STAGES + s) * 8, 1);
    }
    mbar_init(bar_base + 2 * STAGES * 8, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.qmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = p.box_rows * 128 * (2 + BLOCK_N / 64);
      const void* pmap = &p.pmap[tap.map];
      for (int t = t_begin; t < t_end; ++t) {
        int r = t;
        const int tw = r % p.tiles[0]; r /= p.tiles[0];
        const int th = r % p.tiles[1]; r /= p.tiles[1];
        const int td = r % p.tiles[2];
        const int tn = r / p.tiles[2];
        const int w0 = tw * p.box[0], h0 = th * p.box[1], d0 = td * p.box[2], n0 = tn * p.box[3];
        mbar_wait(bar_base + (STAGES + stage) * 8, phase ^ 1u);
        const uint32_t full = bar_base + stage * 8;
        const uint32_t sa = base + stage * STAGE_BYTES;
        mbar_expect_tx(full, tx_bytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d(sa + j * 16384, pmap, full, p.p_c0 + (mb * 2 + j) * 64, w0 + tap.dw, h0 + tap.dh, d0 + tap.dd, n0);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_5d(sa + P_BYTES + j * 16384, &p.qmap, full, nb * BLOCK_N + j * 64, w0, h0, d0, n0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 1, 1);  // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar_base + stage * 8, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa, 16384, 1024);
        const uint64_t bdesc = umma_desc_sw128(sa + P_BYTES, 16384, 1024);
        for (int k = 0; k < ksteps; ++k) {
          // 16 positions = two 8-row groups = 2048 bytes: +128 in the (addr >> 4) field
          tc_mma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (t | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar_base + (STAGES + stage) * 8);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (ntiles > 0) tc_commit(bar_base + 2 * STAGES * 8);
    }
    __syncwarp();
  } else if (ntiles > 0) {
    const int q = warp & 3;
    const int m = mb * 128 + q * 32 + lane;  // channel of P = row of dW block
    mbar_wait(bar_base + 2 * STAGES * 8, 0);
    tc_fence_after();
    float* out = p.dw + tap.dw_ofs + (long long)m * p.ldw + nb * BLOCK_N;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      if (nb * BLOCK_N + c * 32 >= p.N) break;
      uint32_t rr[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
      tmem_ld_wait();
      if (m < p.M) {
        // N is a multiple of 64 on this path, so whole 4-float groups are valid: vector reductions (RED.128)
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (nb * BLOCK_N + c * 32 + j < p.N)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c * 32 + j), "f"(__uint_as_float(rr[j])),
                         "f"(__uint_as_float(rr[j + 1])), "f"(__uint_as_float(rr[j + 2])), "f"(__uint_as_float(rr[j + 3]))
                         : "memory");
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}


// -------------------------------------------------------------------------------------------------------------------
// TWO TAPS PER CTA, swapped roles, for Q (output-gradient) blocks of exactly 128 channels -- the decoder's 128-channel layers.
// A 128 x 128 x 16 tcgen05.mma reads 4 KB + 4 KB of operands from shared memory for 64 cycles of math, so the shared-memory port
// is as busy as the tensor pipe (see conv_tc.cu, conv_tc_swap_kernel).  Here the Q tile is the M side (128 couts = TMEM lanes) and
// the N side is [P(tap a) | P(tap b)], 2 x 128 input channels: one 128 x 256 x 16 instruction per 16 positions computes the
// transposed blocks dW_a^T | dW_b^T, the dy tile is fetched once for both taps (96 KB instead of 128 KB per position tile and
// tap pair), and the epilogue thread owns a cout column of dW: its 256 accumulator columns go out as scalar red.global.add.f32
// that are contiguous across the warp.  An odd last tap runs alone with N = 128.
// MEASURED NEGATIVE (r02, x_1_2 filter gradient, same box): 368 us for both segment launches against 334 us with one tap per
// CTA; 437 us with four 64-position stages.  Halving the number of output blocks (14 tap pairs instead of 27 taps) needs 42
// position splits instead of 16 to fill the SMs, i.e. 2.6x the fp32 reduction traffic, and in the transposed accumulator a thread's
// consecutive columns are different dW rows, so the reductions are scalar instead of 16-byte: the L2 atomic unit, not the tensor
// pipe, sets the time (ncu: tensor pipe 45 %).  Kept opt-in (SAP3D_WGRAD_PAIR=1) with its parity tests; not used by default.
// -------------------------------------------------------------------------------------------------------------------
// ROWS = positions per pipeline stage (the TMA box): 64 gives four 48 KB stages; two 96 KB stages (ROWS = 128) left the tensor pipe
// waiting for TMA (ncu: 45 % active, slower than the one-tap kernel).
template <int STAGES, int ROWS>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_pair_kernel(const __grid_constant__ WgParams p) {
  constexpr int BLK = ROWS * 128;                    // one [ROWS positions][64 channels] box
  constexpr int Q_BYTES = 2 * BLK;                   // 128 channels of Q: two 64-channel boxes
  constexpr int P_BYTES = 4 * BLK;                   // 2 taps x 128 channels of P
  constexpr int STAGE_BYTES = Q_BYTES + P_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode: (split, mblock, tap pair), pair fastest (CTAs resident together share the position slice) ----
  const int npairs = (p.ntaps + 1) / 2;
  int w = blockIdx.x;
  const int pair_id = w % npairs;
  w /= npairs;
  const int mb = w % p.mblocks;
  const int split = w / p.mblocks;
  const WgTap tap_a = p.taps[2 * pair_id];
  const bool has_b = 2 * pair_id + 1 < p.ntaps;
  const WgTap tap_b = p.taps[has_b ? 2 * pair_id + 1 : 2 * pair_id];
  const int per = (p.m_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * per;
  const int t_end = min(p.m_tiles, t_begin + per);
  const int ntiles = max(0, t_end - t_begin);
  const int ksteps = (p.box_rows + 15) / 16;

  if (p.box_rows < ROWS) {    // rows >= box_rows are never written by TMA: zero them once (K padding contributes nothing)
    const int row_bytes0 = p.box_rows * 128;
    for (int s = 0; s < STAGES; ++s)
      for (int blk = 0; blk < 6; ++blk) {
        uint8_t* b0 = smem + s * STAGE_BYTES + blk * BLK;
        for (int i = row_bytes0 + threadIdx.x * 16; i < BLK; i += WG_THREADS * 16) *reinterpret_cast<uint4*>(b0 + i) = make_uint4(0, 0, 0, 0);
      }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (STAGES + s) * 8, 1);
    }
    mbar_init(bar_base + 2 * STAGES * 8, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.qmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = p.box_rows * 128 * (has_b ? 6 : 4);
      const void* pmap_a = &p.pmap[tap_a.map];
      const void* pmap_b = &p.pmap[tap_b.map];
      for (int t = t_begin; t < t_end; ++t) {
        int r = t;
        const int tw = r % p.tiles[0]; r /= p.tiles[0];
        const int th = r % p.tiles[1]; r /= p.tiles[1];
        const int td = r % p.tiles[2];
        const int tn = r / p.tiles[2];
        const int w0 = tw * p.box[0], h0 = th * p.box[1], d0 = td * p.box[2], n0 = tn * p.box[3];
        mbar_wait(bar_base + (STAGES + stage) * 8, phase ^ 1u);
        const uint32_t full = bar_base + stage * 8;
        const uint32_t sa = base + stage * STAGE_BYTES;
        mbar_expect_tx(full, tx_bytes);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_5d(sa + j * BLK, &p.qmap, full, j * 64, w0, h0, d0, n0);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_5d(sa + Q_BYTES + j * BLK, pmap_a, full, p.p_c0 + (mb * 2 + j) * 64, w0 + tap_a.dw, h0 + tap_a.dh, d0 + tap_a.dd, n0);
        if (has_b) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_5d(sa + Q_BYTES + (2 + j) * BLK, pmap_b, full, p.p_c0 + (mb * 2 + j) * 64, w0 + tap_b.dw, h0 + tap_b.dh, d0 + tap_b.dd, n0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = has_b ? umma_idesc_bf16(128, 256, 1, 1) : umma_idesc_bf16(128, 128, 1, 1);  // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar_base + stage * 8, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa, BLK, 1024);               // Q: 128 couts = M
        const uint64_t bdesc = umma_desc_sw128(sa + Q_BYTES, BLK, 1024);     // P(tap a) | P(tap b): 64-channel blocks one box apart
        for (int k = 0; k < ksteps; ++k) tc_mma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (t | k) != 0 ? 1u : 0u);
        tc_commit(bar_base + (STAGES + stage) * 8);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (ntiles > 0) tc_commit(bar_base + 2 * STAGES * 8);
    }
    __syncwarp();
  } else if (ntiles > 0) {
    const int q = warp & 3;
    const int n = q * 32 + lane;          // channel of Q = column of the dW blocks = TMEM lane
    mbar_wait(bar_base + 2 * STAGES * 8, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < (has_b ? 8 : 4); ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
      tmem_ld_wait();
      const WgTap& tap = c < 4 ? tap_a : tap_b;
      const int m0 = mb * 128 + (c & 3) * 32;     // first P channel (row of dW) of this chunk
      float* out = p.dw + tap.dw_ofs + (long long)m0 * p.ldw + n;
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (m0 + j < p.M) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(out + (long long)j * p.ldw), "f"(__uint_as_float(rr[j])) : "memory");
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

static int wg_encode_view(CUtensorMap* m, const TcView& v, const int box[4], char* err, size_t errlen) {
  EncodeTiledFn fn = wg_encode_fn();
  if (!fn) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable");
    return 1;
  }
  cuuint64_t gdim[5] = {(cuuint64_t)v.C, (cuuint64_t)v.dim[0], (cuuint64_t)v.dim[1], (cuuint64_t)v.dim[2], (cuuint64_t)v.dim[3]};
  cuuint64_t gstr[4];
  for (int i = 0; i < 4; ++i) {
    gstr[i] = (cuuint64_t)v.stride[i] * 2;
    if (gstr[i] == 0) gstr[i] = 16;
  }
  cuuint32_t bdim[5] = {64u, (cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2], (cuuint32_t)box[3]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(v.base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(wgrad view) failed: %d (C=%d dims=%d,%d,%d,%d)", (int)r, v.C, v.dim[0], v.dim[1],
             v.dim[2], v.dim[3]);
    return 1;
  }
  return 0;
}

template <int BLOCK_N, int STAGES>
static int wg_launch_t(const WgParams& prm, int grid, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int SMEM = STAGES * (2 * 16384 + (BLOCK_N / 64) * 16384) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(wgrad_tc) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    attr_done = true;
  }
  wgrad_tc_kernel<BLOCK_N, STAGES><<<grid, WG_THREADS, SMEM, stream>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "wgrad_tc launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

static int wg_launch_pair(const WgParams& prm, int grid, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int STAGES = 2, ROWS = 128;   // (4 stages of 64 positions measured slower still: 437 us)
  constexpr int SMEM = STAGES * 6 * ROWS * 128 + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_pair_kernel<STAGES, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(wgrad_tc_pair) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    attr_done = true;
  }
  wgrad_tc_pair_kernel<STAGES, ROWS><<<grid, WG_THREADS, SMEM, stream>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "wgrad_tc_pair launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
// SAP3D_WGRAD_PAIR=1: 128-channel output-gradient blocks take two taps per CTA (off by default: measured slower)
static bool wg_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_WGRAD_PAIR");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

static long long wg_choose_box(const int ext[4], int box[4]) {
  long long best = -1;
  int bb[4] = {1, 1, 1, 1};
  for (int bw = 1; bw <= std::min(ext[0], 128); ++bw)
    for (int bh = 1; bh <= std::min(ext[1], 128 / bw); ++bh)
      for (int bd = 1; bd <= std::min(ext[2], 128 / (bw * bh)); ++bd) {
        int bn = std::min(ext[3], 128 / (bw * bh * bd));
        if (bn < 1) continue;
        long long tiles = (long long)((ext[0] + bw - 1) / bw) * ((ext[1] + bh - 1) / bh) * ((ext[2] + bd - 1) / bd) *
                          ((ext[3] + bn - 1) / bn);
        if (best < 0 || tiles < best || (tiles == best && bw > bb[0])) {
          best = tiles;
          bb[0] = bw; bb[1] = bh; bb[2] = bd; bb[3] = bn;
        }
      }
  for (int i = 0; i < 4; ++i) box[i] = bb[i];
  return best;
}

int tc_wgrad_launch(const TcWgradProblem& pb, cudaStream_t stream, char* err, size_t errlen) {
  if ((int)pb.pviews.size() > WG_MAX_MAPS || (int)pb.taps.size() > WG_MAX_TAPS) {
    snprintf(err, errlen, "tc_wgrad: too many views/taps (%d/%d)", (int)pb.pviews.size(), (int)pb.taps.size());
    return 1;
  }
  if (pb.M % 64 != 0 || pb.N % 64 != 0 || pb.p_c_begin % 64 != 0) {
    snprintf(err, errlen, "tc_wgrad: channel counts must be multiples of 64 (M=%d N=%d)", pb.M, pb.N);
    return 1;
  }
  static thread_local WgParams prm;
  memset(&prm, 0, sizeof(prm));
  int box[4];
  const long long m_tiles = wg_choose_box(pb.ext, box);
  // two taps per CTA (wgrad_tc_pair_kernel; opt-in, see there)
  const bool pair = pb.N == 128 && pb.taps.size() >= 2 && wg_pair_enabled();
  for (size_t v = 0; v < pb.pviews.size(); ++v)
    if (wg_encode_view(&prm.pmap[v], pb.pviews[v], box, err, errlen)) return 1;
  if (wg_encode_view(&prm.qmap, pb.q, box, err, errlen)) return 1;
  for (size_t t = 0; t < pb.taps.size(); ++t) {
    prm.taps[t].map = (int8_t)pb.taps[t].view;
    prm.taps[t].dw = (int8_t)pb.taps[t].off[0];
    prm.taps[t].dh = (int8_t)pb.taps[t].off[1];
    prm.taps[t].dd = (int8_t)pb.taps[t].off[2];
    prm.taps[t].dw_ofs = pb.taps[t].dw_ofs;
  }
  int block_n = pb.N % 256 == 0 ? 256 : (pb.N % 128 == 0 ? 128 : 64);
  prm.ntaps = (int)pb.taps.size();
  prm.mblocks = (pb.M + 127) / 128;
  prm.nblocks = (pb.N + block_n - 1) / block_n;
  prm.m_tiles = (int)m_tiles;
  const long long out_tiles = pair ? (long long)((prm.ntaps + 1) / 2) * prm.mblocks : (long long)prm.ntaps * prm.mblocks * prm.nblocks;
  // split count: the CTAs are one per SM (197 KB of shared memory), so the grid should fill WHOLE waves of 148.  Among the
  // splits that give 2.5 .. 4 waves pick the one wasting the least of its last wave; ties go to fewer splits (less fp32
  // reduction traffic).  (ceil(4*148 / tiles) used to give e.g. 54 tiles x 11 = 594 CTAs = four waves plus two CTAs.)
  long long splits = (4 * 148 + out_tiles - 1) / out_tiles;
  {
    const long long lo = std::max(1ll, (5 * 74) / out_tiles), hi = std::max(lo, (4 * 148) / out_tiles);
    double best_eff = -1.0;
    long long best = splits;
    for (long long sp = lo; sp <= hi; ++sp) {
      const long long g = out_tiles * sp, waves = (g + 147) / 148;
      const double eff = (double)g / (double)(waves * 148);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
    }
    if (best_eff > 0.0) splits = best;
  }
  splits = std::max(1ll, std::min(splits, m_tiles));
  prm.splits = (int)splits;
  for (int i = 0; i < 4; ++i) {
    prm.box[i] = box[i];
    prm.tiles[i] = (pb.ext[i] + box[i] - 1) / box[i];
  }
  prm.box_rows = box[0] * box[1] * box[2] * box[3];
  prm.M = pb.M;
  prm.N = pb.N;
  prm.p_c0 = pb.p_c_begin;
  prm.ldw = pb.ldw;
  prm.dw = pb.dw;
  const long long grid = out_tiles * splits;
  if (grid <= 0 || grid > 0x7fffffffll) {
    snprintf(err, errlen, "tc_wgrad: bad grid %lld", grid);
    return 1;
  }
  if (pair) return wg_launch_pair(prm, (int)grid, stream, err, errlen);
  switch (block_n) {
    case 64: return wg_launch_t<64, 4>(prm, (int)grid, stream, err, errlen);
    case 128: return wg_launch_t<128, 3>(prm, (int)grid, stream, err, errlen);
    case 256: return wg_launch_t<256, 2>(prm, (int)grid, stream, err, errlen);
  }
  return 1;
}

}  // namespace sap3d
