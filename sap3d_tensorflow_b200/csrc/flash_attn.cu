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
      mbar_wait(bar + 10 * 8, (static_cast<uint32_t>(kb) & 1u) ^ 1u);   // P buffer free (previous PV retired)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_s + trow + c * 32, rr);
        tmem_ld_wait();
        const int col0 = kb * 128 + c * 32;
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float e = ex2_approx(fmaf(__uint_as_float(rr[j]), LOG2E, mneg));
          if (col0 + j >= p.Nk) e = 0.f;
          pv[j] = e;
          l += e;
        }
        store_row_chunk_bf16(base + P_OFF, row, c, pv);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar + 8 * 8);
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
