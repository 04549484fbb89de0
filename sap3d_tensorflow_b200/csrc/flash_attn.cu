// Fused self-attention core for the SAGAN-style blocks of utils/network.py:157-193
//   beta = softmax(g f^T) over keys,  o = beta h          (tf.matmul / tf.nn.softmax / tf.matmul, :184-186)
// as flash-style tcgen05 kernels: the [Nq x Nk] score matrix (315 MB fp32 per clip at x_1_3) never leaves the SM.
//
// Forward, one CTA per (128 queries, sample); 192 threads = TMA producer warp, MMA-issuer warp, 4 softmax warps
// (one query row per thread = one TMEM lane).  Because d_k is tiny (C/8, zero-padded to 64) and d_v large, the
// scores are simply computed TWICE instead of rescaling the running output:
//   pass 1:  S = Q K_j^T (TMEM)  -> exact row maximum m
//   pass 2:  S again -> P = exp(S - m) (bf16, written to shared memory as the K-major A operand) -> O += P V_j (TMEM),
//            l += rowsum(P);   epilogue: o = O / l,  lse = m + ln l (kept for the backward kernels).
// V_j tiles are consumed MN-major straight from their [key][d_v] row-major layout (no transpose pass).
// Two CTAs per SM (256 TMEM columns, 112 KB shared memory each) overlap one CTA's exponentials with the other's MMAs.
#include <stdio.h>
#include <string.h>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

namespace sap3d {

constexpr float LOG2E = 1.4426950408889634f;

SAP3D_DEVINL float ex2_approx(float x) {   // MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct alignas(64) FlashParams {
  CUtensorMap qmap, kmap, vmap, domap;
  int Nq, Nk, nkb, nqb;
  void* o;            // [B][Nq][DV] bf16
  float* lse;         // [B][Nq]
  const float* dsum;  // [B][Nq]  rowsum(dO * O)          (backward)
  void* dq;           // [B][Nq][64] bf16                 (backward)
  void* dk;           // [B][Nk][64] bf16
  void* dv;           // [B][Nk][DV] bf16
};

SAP3D_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 32 fp32 values of one row (columns col0 .. col0+31 of a 128-wide K-major SWIZZLE_128B bf16 tile pair) -> shared memory
SAP3D_DEVINL void store_row_chunk_bf16(uint32_t tile_base, int row, int c /* 32-column chunk 0..3 */, const float (&v)[32]) {
  const uint32_t rbase = tile_base + (c >> 1) * 16384 + row * 128;
  const int unit0 = (c & 1) * 4;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t addr = rbase + (static_cast<uint32_t>((unit0 + u) ^ (row & 7)) << 4);
    st_shared_v4(addr, pack_bf16x2(v[u * 8 + 0], v[u * 8 + 1]), pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]),
                 pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
  }
}

// same, from 16 already-packed bf16x2 words
SAP3D_DEVINL void store_row_chunk_packed(uint32_t tile_base, int row, int c, const uint32_t* w) {
  const uint32_t rbase = tile_base + (c >> 1) * 16384 + row * 128;
  const int unit0 = (c & 1) * 4;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    st_shared_v4(rbase + (static_cast<uint32_t>((unit0 + u) ^ (row & 7)) << 4), w[u * 4], w[u * 4 + 1], w[u * 4 + 2], w[u * 4 + 3]);
}

template <int DV>
__global__ void __launch_bounds__(192, (DV == 128 ? 2 : 1)) flash_fwd_kernel(const __grid_constant__ FlashParams p) {
  constexpr int KST = 2;
  constexpr int VCH = DV / 64;
  constexpr uint32_t Q_OFF = 0, K_OFF = 16384, V_OFF = K_OFF + KST * 16384, P_OFF = V_OFF + VCH * 16384, BAR_OFF = P_OFF + 32768;
  constexpr uint32_t TMEM_COLS = DV == 128 ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  // barriers: 0 q_full, 1-2 k_full, 3-4 k_empty, 5 v_full, 6 v_empty, 7 s_full, 8 s_empty, 9 p_full, 10 p_empty, 11 o_full
  const uint32_t bar = base + BAR_OFF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 12 * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y;
  const int nkb = p.nkb;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(bar + i * 8, (i == 8 || i == 9) ? 128 : 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar + 0, 16384);
      tma_load_3d(base + Q_OFF, &p.qmap, bar + 0, 0, qt * 128, b);
      int ks = 0;
      uint32_t kph = 0, vph = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar + (3 + ks) * 8, kph ^ 1u);
          mbar_expect_tx(bar + (1 + ks) * 8, 16384);
          tma_load_3d(base + K_OFF + ks * 16384, &p.kmap, bar + (1 + ks) * 8, 0, kb * 128, b);
          if (++ks == KST) { ks = 0; kph ^= 1u; }
          if (pass == 1) {
            mbar_wait(bar + 6 * 8, vph ^ 1u);
            vph ^= 1u;
            mbar_expect_tx(bar + 5 * 8, VCH * 16384);
#pragma unroll
            for (int j = 0; j < VCH; ++j) tma_load_3d(base + V_OFF + j * 16384, &p.vmap, bar + 5 * 8, j * 64, kb * 128, b);
          }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV, 0, 1);   // B = V tile, MN-major
      mbar_wait(bar + 0, 0);
      tc_fence_after();
      const uint64_t qdesc = umma_desc_sw128(base + Q_OFF, 16, 1024);
      int ks = 0, n_s = 0;
      uint32_t kph = 0, pfph = 0, vfph = 0;
      auto issue_s = [&]() {
        mbar_wait(bar + (1 + ks) * 8, kph);
        mbar_wait(bar + 8 * 8, (static_cast<uint32_t>(n_s) & 1u) ^ 1u);   // previous score tile consumed
        ++n_s;
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(base + K_OFF + ks * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        tc_commit(bar + (3 + ks) * 8);
        tc_commit(bar + 7 * 8);
        if (++ks == KST) { ks = 0; kph ^= 1u; }
      };
      for (int kb = 0; kb < nkb; ++kb) issue_s();   // pass 1
      issue_s();                                     // pass 2, tile 0
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb + 1 < nkb) issue_s();                 // next scores run while the softmax warps work on this tile
        mbar_wait(bar + 9 * 8, pfph); pfph ^= 1u;
        mbar_wait(bar + 5 * 8, vfph); vfph ^= 1u;
        tc_fence_after();
        const uint64_t vdesc = umma_desc_sw128(base + V_OFF, 16384, 1024);
#pragma unroll
        for (int c64 = 0; c64 < 2; ++c64) {
          const uint64_t pdesc = umma_desc_sw128(base + P_OFF + c64 * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_o, pdesc + 2 * k, vdesc + 128 * (c64 * 4 + k), idesc_o, (kb | c64 | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 10 * 8);
        tc_commit(bar + 6 * 8);
      }
      tc_commit(bar + 11 * 8);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t trow = static_cast<uint32_t>(q * 32) << 16;
    uint32_t sfph = 0;
    float m = -INFINITY;
    for (int kb = 0; kb < nkb; ++kb) {               // pass 1: exact row maximum
      mbar_wait(bar + 7 * 8, sfph); sfph ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_s + trow + c * 32, rr);
        tmem_ld_wait();
        const int col0 = kb * 128 + c * 32;
        if (col0 + 32 <= p.Nk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rr[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < p.Nk) m = fmaxf(m, __uint_as_float(rr[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(bar + 8 * 8);
    }
    const float mneg = -m * LOG2E;
    float l = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {               // pass 2: probabilities -> shared memory -> O += P V
      mbar_wait(bar + 7 * 8, sfph); sfph ^= 1u;
      tc_fence_after();
      // The scores are pulled out of TMEM in two halves and the S buffer is released right after the second load, so the
      // next Q K^T runs under this tile's exponentials; the P buffer is only waited for once the probabilities sit packed
      // in registers, so the previous P V runs under them as well.
      uint32_t pp[64];
      const bool last = (kb + 1) * 128 > p.Nk;       // only the last key block needs the column mask
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(tmem_s + trow + (2 * h) * 32, r0);
        tmem_ld_32x32(tmem_s + trow + (2 * h + 1) * 32, r1);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          mbar_arrive(bar + 8 * 8);
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const uint32_t* rr = cc == 0 ? r0 : r1;
          const int c = 2 * h + cc;
          const int col0 = kb * 128 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(rr[j]), LOG2E, mneg));
            float e1 = ex2_approx(fmaf(__uint_as_float(rr[j + 1]), LOG2E, mneg));
            if (last) {
              if (col0 + j >= p.Nk) e0 = 0.f;
              if (col0 + j + 1 >= p.Nk) e1 = 0.f;
            }
            l += e0 + e1;
            pp[c * 16 + j / 2] = pack_bf16x2(e0, e1);
          }
        }
      }
      mbar_wait(bar + 10 * 8, (static_cast<uint32_t>(kb) & 1u) ^ 1u);   // P buffer free (previous PV retired)
#pragma unroll
      for (int c = 0; c < 4; ++c) store_row_chunk_packed(base + P_OFF, row, c, pp + c * 16);
      fence_proxy_async();
      mbar_arrive(bar + 9 * 8);
    }
    mbar_wait(bar + 11 * 8, 0);
    tc_fence_after();
    const int qrow = qt * 128 + row;
    const float inv = 1.f / l;
    bf16* o = reinterpret_cast<bf16*>(p.o) + ((long long)b * p.Nq + qrow) * DV;
#pragma unroll 1
    for (int c = 0; c < DV / 32; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(tmem_o + trow + c * 32, rr);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = __uint_as_float(rr[g * 8 + j]) * inv;
          Vec8<bf16>::store(o + c * 32 + g * 8, w8);
        }
      }
    }
    if (qrow < p.Nq && p.lse) p.lse[(long long)b * p.Nq + qrow] = m + logf(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// =================================================================================================
// Backward.  With P = exp(S - lse), D = rowsum(dO * O):
//   dV = P^T dO,   dP = dO V^T,   dS = P * (dP - D),   dQ = dS K,   dK = dS^T Q
// Kernel A (one CTA per 128 queries) recomputes S and produces dQ; kernel B (one CTA per 128 keys and a slice of the
// query tiles) recomputes S^T = K Q^T directly (so P^T / dS^T come out of TMEM already transposed) and accumulates
// dV, dK in TMEM, reduced across query slices with red.global.add.v4.f32 into fp32 buffers.
// =================================================================================================

// D[row] = sum_c dO[row][c] * O[row][c]    (one warp per row)
__global__ void __launch_bounds__(256) flash_dsum_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long rows, int dv,
                                                         float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = w; r < rows; r += nw) {
    float acc = 0.f;
    for (int c = lane * 8; c < dv; c += 256) {
      float a[8], b[8];
      Vec8<bf16>::load(o + r * dv + c, a);
      Vec8<bf16>::load(d_o + r * dv + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) dsum[r] = acc;
  }
}

template <int DV>
__global__ void __launch_bounds__(192) flash_bwd_dq_kernel(const __grid_constant__ FlashParams p) {
  constexpr int KST = 2;
  constexpr int VCH = DV / 64;
  constexpr uint32_t Q_OFF = 0, DO_OFF = 16384, K_OFF = DO_OFF + VCH * 16384, V_OFF = K_OFF + KST * 16384, DS_OFF = V_OFF + VCH * 16384,
                     BAR_OFF = DS_OFF + 32768;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  // 0 qdo_full, 1-2 k_full, 3-4 k_empty, 5 v_full, 6 v_empty, 7-8 s_full, 9-10 s_empty, 11 dp_full, 12 dp_empty, 13 ds_full,
  // 14 ds_empty, 15 dq_full
  const uint32_t bar = base + BAR_OFF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 16 * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y;
  const int nkb = p.nkb;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(bar + i * 8, (i == 9 || i == 10 || i == 12 || i == 13) ? 128 : 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); tma_prefetch_desc(&p.domap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 256, tmem_dq = tmem_base + 384;   // S double-buffered: +0, +128

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar + 0, 16384 + VCH * 16384);
      tma_load_3d(base + Q_OFF, &p.qmap, bar + 0, 0, qt * 128, b);
#pragma unroll
      for (int j = 0; j < VCH; ++j) tma_load_3d(base + DO_OFF + j * 16384, &p.domap, bar + 0, j * 64, qt * 128, b);
      for (int kb = 0; kb < nkb; ++kb) {
        const int ks = kb & 1;
        mbar_wait(bar + (3 + ks) * 8, ((static_cast<uint32_t>(kb) >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar + (1 + ks) * 8, 16384);
        tma_load_3d(base + K_OFF + ks * 16384, &p.kmap, bar + (1 + ks) * 8, 0, kb * 128, b);
        mbar_wait(bar + 6 * 8, (static_cast<uint32_t>(kb) & 1u) ^ 1u);
        mbar_expect_tx(bar + 5 * 8, VCH * 16384);
#pragma unroll
        for (int j = 0; j < VCH; ++j) tma_load_3d(base + V_OFF + j * 16384, &p.vmap, bar + 5 * 8, j * 64, kb * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);   // B = K tile, MN-major (N = d_k, K = keys)
      mbar_wait(bar + 0, 0);
      tc_fence_after();
      const uint64_t qdesc = umma_desc_sw128(base + Q_OFF, 16, 1024);
      auto issue_s = [&](int kb) {
        const int ks = kb & 1;
        const uint32_t use = static_cast<uint32_t>(kb) >> 1;
        mbar_wait(bar + (1 + ks) * 8, use & 1u);
        mbar_wait(bar + (9 + ks) * 8, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(base + K_OFF + ks * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_s + ks * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        tc_commit(bar + (7 + ks) * 8);
      };
      issue_s(0);
      for (int kb = 0; kb < nkb; ++kb) {
        // dP(kb) = dO V_kb^T
        mbar_wait(bar + 5 * 8, static_cast<uint32_t>(kb) & 1u);
        mbar_wait(bar + 12 * 8, (static_cast<uint32_t>(kb) & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < VCH; ++c) {
          const uint64_t ad = umma_desc_sw128(base + DO_OFF + c * 16384, 16, 1024);
          const uint64_t bd = umma_desc_sw128(base + V_OFF + c * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_dp, ad + 2 * k, bd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 11 * 8);
        tc_commit(bar + 6 * 8);
        if (kb + 1 < nkb) issue_s(kb + 1);
        // dQ += dS(kb) K_kb
        mbar_wait(bar + 13 * 8, static_cast<uint32_t>(kb) & 1u);
        tc_fence_after();
        const uint64_t kd = umma_desc_sw128(base + K_OFF + (kb & 1) * 16384, 16384, 1024);
#pragma unroll
        for (int c64 = 0; c64 < 2; ++c64) {
          const uint64_t ad = umma_desc_sw128(base + DS_OFF + c64 * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_dq, ad + 2 * k, kd + 128 * (c64 * 4 + k), idesc_dq, (kb | c64 | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 14 * 8);
        tc_commit(bar + (3 + (kb & 1)) * 8);
      }
      tc_commit(bar + 15 * 8);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t trow = static_cast<uint32_t>(q * 32) << 16;
    const int qrow = qt * 128 + row;
    const bool valid = qrow < p.Nq;
    const float lneg = valid ? -p.lse[(long long)b * p.Nq + qrow] * LOG2E : -INFINITY;
    const float dsum = valid ? p.dsum[(long long)b * p.Nq + qrow] : 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      const int sb = kb & 1;
      uint32_t pp[64];   // the 128 probabilities of this row, bf16x2
      const bool last = (kb + 1) * 128 > p.Nk;
      mbar_wait(bar + (7 + sb) * 8, (static_cast<uint32_t>(kb) >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_s + sb * 128 + trow + c * 32, rr);
        tmem_ld_wait();
        const int col0 = kb * 128 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float e0 = ex2_approx(fmaf(__uint_as_float(rr[j]), LOG2E, lneg));
          float e1 = ex2_approx(fmaf(__uint_as_float(rr[j + 1]), LOG2E, lneg));
          if (last) {
            if (col0 + j >= p.Nk) e0 = 0.f;
            if (col0 + j + 1 >= p.Nk) e1 = 0.f;
          }
          pp[c * 16 + j / 2] = pack_bf16x2(e0, e1);
        }
      }
      tc_fence_before();
      mbar_arrive(bar + (9 + sb) * 8);
      mbar_wait(bar + 11 * 8, static_cast<uint32_t>(kb) & 1u);
      tc_fence_after();
      mbar_wait(bar + 14 * 8, (static_cast<uint32_t>(kb) & 1u) ^ 1u);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_dp + trow + c * 32, rr);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 pr = unpack_bf16x2(pp[c * 16 + j / 2]);
          ds[j] = pr.x * (__uint_as_float(rr[j]) - dsum);
          ds[j + 1] = pr.y * (__uint_as_float(rr[j + 1]) - dsum);
        }
        store_row_chunk_bf16(base + DS_OFF, row, c, ds);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar + 12 * 8);
      mbar_arrive(bar + 13 * 8);
    }
    mbar_wait(bar + 15 * 8, 0);
    tc_fence_after();
    bf16* dq = reinterpret_cast<bf16*>(p.dq) + ((long long)b * p.Nq + qrow) * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(tmem_dq + trow + c * 32, rr);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = __uint_as_float(rr[g * 8 + j]);
          Vec8<bf16>::store(dq + c * 32 + g * 8, w8);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct FlashDkvExtra {
  int splits, tiles_per_split;
  float* dk32;   // [B][Nk][64]
  float* dv32;   // [B][Nk][DV]
};

template <int DV>
__global__ void __launch_bounds__(192) flash_bwd_dkv_kernel(const __grid_constant__ FlashParams p, const FlashDkvExtra x) {
  constexpr int VCH = DV / 64;
  constexpr uint32_t K_OFF = 0, V_OFF = 16384, Q_OFF = V_OFF + VCH * 16384, DO_OFF = Q_OFF + 2 * 16384, PT_OFF = DO_OFF + 2 * VCH * 16384,
                     DST_OFF = PT_OFF + 32768, STAT_OFF = DST_OFF + 32768, BAR_OFF = STAT_OFF + 2 * 2 * 128 * 4;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  // 0 kv_full, 1-2 q_full, 3-4 q_empty, 5 st_full, 6 st_empty, 7 dpt_full, 8 dpt_empty, 9 pt_full, 10 pt_empty, 11 dst_full,
  // 12 dst_empty, 13 acc_full
  const uint32_t bar = base + BAR_OFF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 14 * 8);
  float* s_stat = reinterpret_cast<float*>(smem + STAT_OFF);   // [2 stages][lse | dsum][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbk = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int t_begin = split * x.tiles_per_split;
  const int t_end = min(p.nqb, t_begin + x.tiles_per_split);
  const int nt = max(0, t_end - t_begin);
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 14; ++i) mbar_init(bar + i * 8, (i == 6 || i == 8 || i == 9 || i == 11) ? 128 : 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); tma_prefetch_desc(&p.domap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 256 + DV;

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      mbar_expect_tx(bar + 0, 16384 + VCH * 16384);
      tma_load_3d(base + K_OFF, &p.kmap, bar + 0, 0, kbk * 128, b);
#pragma unroll
      for (int j = 0; j < VCH; ++j) tma_load_3d(base + V_OFF + j * 16384, &p.vmap, bar + 0, j * 64, kbk * 128, b);
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1;
        mbar_wait(bar + (3 + s) * 8, ((static_cast<uint32_t>(i) >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar + (1 + s) * 8, 16384 + VCH * 16384);
        tma_load_3d(base + Q_OFF + s * 16384, &p.qmap, bar + (1 + s) * 8, 0, (t_begin + i) * 128, b);
#pragma unroll
        for (int j = 0; j < VCH; ++j)
          tma_load_3d(base + DO_OFF + (s * VCH + j) * 16384, &p.domap, bar + (1 + s) * 8, j * 64, (t_begin + i) * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dv = umma_idesc_bf16(128, DV, 0, 1);   // B = dO tile, MN-major (N = d_v, K = queries)
      constexpr uint32_t idesc_dk = umma_idesc_bf16(128, 64, 0, 1);   // B = Q tile,  MN-major (N = d_k, K = queries)
      mbar_wait(bar + 0, 0);
      tc_fence_after();
      const uint64_t kdesc = umma_desc_sw128(base + K_OFF, 16, 1024);
      auto issue_scores = [&](int i) {   // S^T(i) = K Q_i^T and dP^T(i) = V dO_i^T
        const int s = i & 1;
        mbar_wait(bar + (1 + s) * 8, (static_cast<uint32_t>(i) >> 1) & 1u);
        mbar_wait(bar + 6 * 8, (static_cast<uint32_t>(i) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(base + Q_OFF + s * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_st, kdesc + 2 * k, qd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        tc_commit(bar + 5 * 8);
        mbar_wait(bar + 8 * 8, (static_cast<uint32_t>(i) & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < VCH; ++c) {
          const uint64_t ad = umma_desc_sw128(base + V_OFF + c * 16384, 16, 1024);
          const uint64_t bd = umma_desc_sw128(base + DO_OFF + (s * VCH + c) * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_dpt, ad + 2 * k, bd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 7 * 8);
      };
      issue_scores(0);
      for (int i = 0; i < nt; ++i) {
        const int s = i & 1;
        if (i + 1 < nt) issue_scores(i + 1);
        // dV += P^T(i) dO_i
        mbar_wait(bar + 9 * 8, static_cast<uint32_t>(i) & 1u);
        tc_fence_after();
        const uint64_t dod = umma_desc_sw128(base + DO_OFF + s * VCH * 16384, 16384, 1024);
#pragma unroll
        for (int c64 = 0; c64 < 2; ++c64) {
          const uint64_t ad = umma_desc_sw128(base + PT_OFF + c64 * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_dv, ad + 2 * k, dod + 128 * (c64 * 4 + k), idesc_dv, (i | c64 | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 10 * 8);
        // dK += dS^T(i) Q_i
        mbar_wait(bar + 11 * 8, static_cast<uint32_t>(i) & 1u);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(base + Q_OFF + s * 16384, 16384, 1024);
#pragma unroll
        for (int c64 = 0; c64 < 2; ++c64) {
          const uint64_t ad = umma_desc_sw128(base + DST_OFF + c64 * 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_dk, ad + 2 * k, qd + 128 * (c64 * 4 + k), idesc_dk, (i | c64 | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar + 12 * 8);
        tc_commit(bar + (3 + s) * 8);
      }
      tc_commit(bar + 13 * 8);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;           // key row of this thread
    const int et = threadIdx.x - 64;         // 0..127
    const uint32_t trow = static_cast<uint32_t>(q * 32) << 16;
    const int krow = kbk * 128 + row;
    for (int i = 0; i < nt; ++i) {
      const int s = i & 1;
      {   // per-query statistics of this tile -> shared memory (queries are the COLUMNS here)
        const int qg = (t_begin + i) * 128 + et;
        const bool qv = qg < p.Nq;
        s_stat[(s * 2 + 0) * 128 + et] = qv ? -p.lse[(long long)b * p.Nq + qg] * LOG2E : -INFINITY;
        s_stat[(s * 2 + 1) * 128 + et] = qv ? p.dsum[(long long)b * p.Nq + qg] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float* lneg = s_stat + (s * 2 + 0) * 128;
      const float* dsm = s_stat + (s * 2 + 1) * 128;
      uint32_t pp[64];
      mbar_wait(bar + 5 * 8, static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      mbar_wait(bar + 10 * 8, (static_cast<uint32_t>(i) & 1u) ^ 1u);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_st + trow + c * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 ln = *reinterpret_cast<const float4*>(lneg + c * 32 + j);   // broadcast LDS.128
          const float e0 = ex2_approx(fmaf(__uint_as_float(rr[j]), LOG2E, ln.x));
          const float e1 = ex2_approx(fmaf(__uint_as_float(rr[j + 1]), LOG2E, ln.y));
          const float e2 = ex2_approx(fmaf(__uint_as_float(rr[j + 2]), LOG2E, ln.z));
          const float e3 = ex2_approx(fmaf(__uint_as_float(rr[j + 3]), LOG2E, ln.w));
          pp[c * 16 + j / 2] = pack_bf16x2(e0, e1);
          pp[c * 16 + j / 2 + 1] = pack_bf16x2(e2, e3);
        }
        store_row_chunk_packed(base + PT_OFF, row, c, pp + c * 16);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar + 6 * 8);
      mbar_arrive(bar + 9 * 8);
      mbar_wait(bar + 7 * 8, static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      mbar_wait(bar + 12 * 8, (static_cast<uint32_t>(i) & 1u) ^ 1u);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_dpt + trow + c * 32, rr);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 dm = *reinterpret_cast<const float4*>(dsm + c * 32 + j);
          const float2 p0 = unpack_bf16x2(pp[c * 16 + j / 2]), p1 = unpack_bf16x2(pp[c * 16 + j / 2 + 1]);
          ds[j] = p0.x * (__uint_as_float(rr[j]) - dm.x);
          ds[j + 1] = p0.y * (__uint_as_float(rr[j + 1]) - dm.y);
          ds[j + 2] = p1.x * (__uint_as_float(rr[j + 2]) - dm.z);
          ds[j + 3] = p1.y * (__uint_as_float(rr[j + 3]) - dm.w);
        }
        store_row_chunk_bf16(base + DST_OFF, row, c, ds);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar + 8 * 8);
      mbar_arrive(bar + 11 * 8);
    }
    if (nt > 0) {
      mbar_wait(bar + 13 * 8, 0);
      tc_fence_after();
      const bool valid = krow < p.Nk;
      float* dv = x.dv32 + ((long long)b * p.Nk + krow) * DV;
      float* dk = x.dk32 + ((long long)b * p.Nk + krow) * 64;
#pragma unroll 1
      for (int c = 0; c < DV / 32 + 2; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_dv + trow + c * 32, rr);   // dK's 64 columns follow dV's in TMEM
        tmem_ld_wait();
        if (valid) {
          float* out = c < DV / 32 ? dv + c * 32 : dk + (c - DV / 32) * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + j), "f"(__uint_as_float(rr[j])),
                         "f"(__uint_as_float(rr[j + 1])), "f"(__uint_as_float(rr[j + 2])), "f"(__uint_as_float(rr[j + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void __launch_bounds__(256) flash_cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<float>::load(src + i * 8, v);
    Vec8<bf16>::store(dst + i * 8, v);
  }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn fa_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// [B][rows][cols] bf16, dense -> rank-3 map {cols, rows, B}, box {64, 128, 1}, 128B swizzle, zero OOB fill
static int encode_rows(CUtensorMap* m, const void* ptr, int B, int rows, int cols) {
  EncodeTiledFn fn = fa_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  cuuint32_t bdim[3] = {64u, 128u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(flash rows=%d cols=%d B=%d) failed: %d", rows, cols, B, (int)r);
  return 0;
}

template <int DV>
static int launch_fwd(const FlashParams& prm, int B, cudaStream_t st) {
  constexpr int SMEM = 16384 + 2 * 16384 + (DV / 64) * 16384 + 32768 + 12 * 8 + 32;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(flash_fwd): %s", cudaGetErrorString(e));
    done = true;
  }
  flash_fwd_kernel<DV><<<dim3(prm.nqb, B), 192, SMEM, st>>>(prm);
  return check_launch("flash_attn_fwd");
}


template <int DV>
static int launch_bwd(const FlashParams& prm, const FlashDkvExtra& ex, int B, cudaStream_t st) {
  constexpr int SMEM_DQ = 16384 + (DV / 64) * 16384 + 2 * 16384 + (DV / 64) * 16384 + 32768 + 16 * 8 + 32;
  constexpr int SMEM_DKV = 16384 + (DV / 64) * 16384 + 2 * 16384 + 2 * (DV / 64) * 16384 + 32768 + 32768 + 2 * 2 * 128 * 4 + 14 * 8 + 32;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(flash_bwd_dq_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_bwd_dkv_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(flash_bwd): %s", cudaGetErrorString(e));
    done = true;
  }
  flash_bwd_dq_kernel<DV><<<dim3(prm.nqb, B), 192, SMEM_DQ, st>>>(prm);
  if (check_launch("flash_attn_bwd dq")) return 1;
  flash_bwd_dkv_kernel<DV><<<dim3(prm.nkb, B, ex.splits), 192, SMEM_DKV, st>>>(prm, ex);
  return check_launch("flash_attn_bwd dkv");
}

}  // namespace sap3d

using namespace sap3d;

extern "C" int sap3d_flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t Nq, int32_t Nk,
                                    int32_t dk, int32_t dv, void* stream) {
  if (require_device()) return 1;
  if (dk != 64 || (dv != 128 && dv != 256)) return set_error("flash_attn_fwd: needs d_k == 64 (zero-padded) and d_v in {128, 256} (got %d, %d)", dk, dv);
  if (B < 1 || Nq < 1 || Nk < 1) return set_error("flash_attn_fwd: empty problem");
  static thread_local FlashParams prm;
  memset(&prm, 0, sizeof(prm));
  if (encode_rows(&prm.qmap, q, B, Nq, 64) || encode_rows(&prm.kmap, k, B, Nk, 64) || encode_rows(&prm.vmap, v, B, Nk, dv)) return 1;
  prm.Nq = Nq; prm.Nk = Nk; prm.nkb = (Nk + 127) / 128; prm.nqb = (Nq + 127) / 128;
  prm.o = o; prm.lse = lse;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dv == 128 ? launch_fwd<128>(prm, B, st) : launch_fwd<256>(prm, B, st);
}

extern "C" size_t sap3d_flash_attn_bwd_workspace(int32_t B, int32_t Nq, int32_t Nk, int32_t dv) {
  return ((size_t)B * Nq + (size_t)B * Nk * (64 + dv)) * sizeof(float) + 256;
}

extern "C" int sap3d_flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                                    void* dq, void* dk, void* dv_out, int32_t B, int32_t Nq, int32_t Nk, int32_t dk_dim, int32_t dv,
                                    void* workspace, void* stream) {
  if (require_device()) return 1;
  if (dk_dim != 64 || dv != 128) return set_error("flash_attn_bwd: needs d_k == 64 (zero-padded) and d_v == 128 (got %d, %d)", dk_dim, dv);
  if (!workspace || !lse) return set_error("flash_attn_bwd: NULL workspace / lse");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static thread_local FlashParams prm;
  memset(&prm, 0, sizeof(prm));
  if (encode_rows(&prm.qmap, q, B, Nq, 64) || encode_rows(&prm.kmap, k, B, Nk, 64) || encode_rows(&prm.vmap, v, B, Nk, dv) ||
      encode_rows(&prm.domap, d_o, B, Nq, dv))
    return 1;
  float* ws = reinterpret_cast<float*>(workspace);
  float* dsum = ws;
  float* dk32 = ws + (((size_t)B * Nq + 63) / 64) * 64;
  float* dv32 = dk32 + (size_t)B * Nk * 64;
  prm.Nq = Nq; prm.Nk = Nk; prm.nkb = (Nk + 127) / 128; prm.nqb = (Nq + 127) / 128;
  prm.lse = const_cast<float*>(lse); prm.dsum = dsum; prm.dq = dq; prm.dk = dk; prm.dv = dv_out;
  const long long rows = (long long)B * Nq;
  {
    long long blocks = (rows * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    flash_dsum_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), rows, dv, dsum);
    if (check_launch("flash_attn_bwd dsum")) return 1;
  }
  if (cudaMemsetAsync(dk32, 0, (size_t)B * Nk * (64 + dv) * sizeof(float), st) != cudaSuccess) return set_error("flash_attn_bwd: memset failed");
  FlashDkvExtra ex;
  // split the query tiles so that (key blocks x samples x splits) fills the 148 SMs about evenly
  const long long base_ctas = (long long)prm.nkb * B;
  int splits = 1;
  double best = 0.0;
  for (int s = 1; s <= 8 && s <= prm.nqb; ++s) {
    const double waves = (double)(base_ctas * s) / 148.0;
    const double eff = waves / (double)((long long)(waves + 0.999999));
    if (eff > best + 0.02) { best = eff; splits = s; }
  }
  ex.splits = splits;
  ex.tiles_per_split = (prm.nqb + splits - 1) / splits;
  ex.dk32 = dk32; ex.dv32 = dv32;
  if (launch_bwd<128>(prm, ex, B, st)) return 1;
  {
    const long long n8k = (long long)B * Nk * 64 / 8, n8v = (long long)B * Nk * dv / 8;
    flash_cast_kernel<<<(unsigned)((n8k + 255) / 256 > 1184 ? 1184 : (n8k + 255) / 256), 256, 0, st>>>(dk32, reinterpret_cast<bf16*>(dk), n8k);
    flash_cast_kernel<<<(unsigned)((n8v + 255) / 256 > 1184 ? 1184 : (n8v + 255) / 256), 256, 0, st>>>(dv32, reinterpret_cast<bf16*>(dv_out), n8v);
    if (check_launch("flash_attn_bwd cast")) return 1;
  }
  return 0;
}
