// tcgen05 / TMEM implicit-GEMM convolution for sm_100a, operands fed by TMA.
//
// One CTA computes a [128 output positions] x [BLOCK_N output channels] tile.  The 128 positions
// are a (bn x bd x bh x bw) box of the NDHWC output; for every filter tap the matching input box
// (shifted by the tap offset; out-of-bounds = TF 'SAME' zero padding, filled by the TMA unit) is
// fetched with ONE 5-D cp.async.bulk.tensor into a 128-byte-swizzled K-major tile, so no im2col
// buffer ever exists.  Concatenated inputs are extra K segments (extra tensor maps), transposed
// convolutions are output-parity classes (each a stride-1 problem with a strided output view).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + tcgen05.mma issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / affine / ReLU -> per-channel sum & sum-of-squares for
// BatchNorm/GroupNorm -> bf16 128-bit stores).
//
// Replaces the cuDNN conv3d fprop/dgrad kernels TensorFlow dispatches for tf.nn.conv3d /
// tf.layers.conv3d / tf.layers.conv3d_transpose (reference p3d.py:18-27,86,112,125,375-393;
// utils/network.py:100-110).
#include "conv_tc.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>

#include "abi_util.cuh"
#include "common.cuh"

namespace sap3d {

constexpr int TC_MAX_MAPS = 28;   // 27 parity views of dy (transposed conv k3 s4 data gradient) + slack
constexpr int TC_MAX_TAPS = 128;
constexpr int TC_MAX_CLS = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_QD = 4;          // persistent kernels: depth of the in-CTA unit queue (producer -> MMA / epilogue warps)
constexpr int TC_HALO_MAX_ROWS = 160;   // halo-tile kernels: 128 + 2 * halo_inner rows per activation box, halo_inner <= 16

struct TcTap {
  int8_t map, dw, dh, dd;
  int16_t nchunk, c0;  // channel chunks of 64; first chunk
  int32_t kofs;
};
struct TcClass {
  int16_t tap_begin, tap_count;
  int32_t nkb;  // total k-blocks of the class
  long long out_ofs;
};

struct alignas(64) TcConvParams {
  CUtensorMap amap[TC_MAX_MAPS];
  CUtensorMap bmap;
  CUtensorMap bmap_half;     // box {64, BLOCK_N / 2}: the part of a weight tile one CTA of a 2-CTA cluster multicasts
  CUtensorMap bmap_quarter;  // box {64, BLOCK_N / 4}: ... of a 4-CTA cluster
  TcTap taps[TC_MAX_TAPS];
  TcClass cls[TC_MAX_CLS];
  int ncls, m_tiles, n_tiles;
  int tiles[4];  // W,H,D,N
  int box[4];
  int ext[4];
  long long so[4];
  int cout, box_rows;
  void* out;
  void* out2;          // merged two-segment data gradient: columns >= seg_split go to out2 (same row strides)
  int seg_split, accumulate2;
  const float* bias;
  float* stats;
  const float* scale;
  const float* shift;
  int relu, accumulate, out_f32;
  int b_batched;       // the weight-side operand is per sample: third TMA coordinate = the tile's batch index
  int balanced;        // persistent kernels, one class and one column tile: units are m-tile groups, singles at the end
  int n_units;         // persistent kernels: work units handed out through the ticket counter
  int n_pair_units;    // balanced: units [0, n_pair_units) are groups of MT consecutive m-tiles, the rest single m-tiles
  unsigned int* sched; // persistent kernels: {next unit ticket, CTAs finished}; zero at launch, reset by the last CTA
  TcFuseBN fuse;       // conv_tc_kernel: BatchNorm finished inside the launch when fuse.y != NULL (grid barrier on gbar)
  unsigned int* gbar;  // {arrivals, CTAs past the barrier}: a slot of the same zero-initialised pool as sched
  int halo_rows;       // halo-tile kernels: rows of one activation box (128 + 2 * halo_inner); taps are sorted in triples
  int halo_inner;      // ... rows per step along the halo axis (product of the inner box extents; multiple of 8)
  unsigned long long* dbg;   // phase-timing probe (tools/conv_phase_probe.py): [cta][16][2] = (clock64, globaltimer); normally NULL
};

SAP3D_DEVINL void dbg_mark(const TcConvParams& p, int i) {
  if (p.dbg != nullptr) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[(blockIdx.x * 16 + i) * 2] = clock64();
    p.dbg[(blockIdx.x * 16 + i) * 2 + 1] = g;
  }
}

// column sums of a 32(lanes) x 32(values) block: on return lane l holds the sum over all lanes of
// the caller's v[l].  31 shuffles instead of 160.
SAP3D_DEVINL float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      float send = upper ? v[i] : v[i + o];
      float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// MT = number of 128-row M sub-tiles per CTA (each with its own TMEM accumulator) that share every B tile:
// MT = 2 cuts the L2->SM bytes per MAC by 25 % (the kernel is L2-bandwidth bound at 128x128 tiles).
// SPLIT > 1 (cluster of SPLIT CTAs along K, layers with few output tiles): every CTA of the cluster accumulates the
// same [128 x BLOCK_N] tile over 1/SPLIT of the k-blocks, then the partial accumulators are reduce-scattered through
// distributed shared memory (each CTA owns BLOCK_N/SPLIT columns, receives its peers' fp32 slices with
// st.shared::cluster, and runs the normal epilogue for its columns).  One SM's L2->SMEM ingest (~85 GB/s) is what bounds
// the K loop of the 16-CTA backbone layers; the split spreads it over SPLIT SMs with no extra global traffic or launch.
template <int BLOCK_N, int STAGES, int MT, int SPLIT = 1>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ TcConvParams p) {
  static_assert(SPLIT == 1 || (MT == 1 && BLOCK_N == 128 && (SPLIT == 2 || SPLIT == 4)), "split-K clusters: 128-column tiles, 2 or 4 CTAs");
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  constexpr int OWN_CHUNKS = SPLIT > 1 ? 4 / SPLIT : 4;                   // 32-column chunks owned per CTA

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], tmem_full, recv (split-K exchange)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + (2 * STAGES + 2) * 8);
  float* s_stats = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + (2 * STAGES + 2) * 8 + 16);  // [4][2][BLOCK_N]
  float* s_ep = s_stats + 4 * 2 * BLOCK_N;   // [3][BLOCK_N]: bias, scale, shift of this tile's columns (staged under the K loop)
  const uint32_t recv_base = (base + STAGES * STAGE_BYTES + (2 * STAGES + 2) * 8 + 16 + (4 * 2 + 3) * BLOCK_N * 4 + 127u) & ~127u;
  const uint32_t recv_bar = bar_base + (2 * STAGES + 1) * 8;
  constexpr uint32_t XCHG_BYTES = OWN_CHUNKS * 16384;   // fp32 partials one peer sends me: [OWN_CHUNKS][128 rows][32 columns]
  const uint32_t rank = SPLIT > 1 ? cluster_ctarank() : 0u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) dbg_mark(p, 0);

  // ---- tile decode -------------------------------------------------------------------------
  int tile = blockIdx.x / SPLIT;
  const int nt = tile % p.n_tiles;
  tile /= p.n_tiles;
  const int m_groups = (p.m_tiles + MT - 1) / MT;
  const int cls_id = tile / m_groups;
  const int mg = tile - cls_id * m_groups;
  int mts[MT], w0s[MT], h0s[MT], d0s[MT], n0s[MT];
  bool live[MT];
#pragma unroll
  for (int sub = 0; sub < MT; ++sub) {
    int mt = mg * MT + sub;
    live[sub] = mt < p.m_tiles;
    if (!live[sub]) mt = p.m_tiles - 1;  // ragged tail: recompute the last tile, drop its epilogue
    mts[sub] = mt;
    int t = mt;
    const int tw = t % p.tiles[0];
    t /= p.tiles[0];
    const int th = t % p.tiles[1];
    t /= p.tiles[1];
    const int td = t % p.tiles[2];
    const int tn = t / p.tiles[2];
    w0s[sub] = tw * p.box[0]; h0s[sub] = th * p.box[1]; d0s[sub] = td * p.box[2]; n0s[sub] = tn * p.box[3];
  }
  const TcClass cls = p.cls[cls_id];
  const int kb_per = (cls.nkb + SPLIT - 1) / SPLIT;
  const int kb_begin = min(cls.nkb, (int)rank * kb_per);
  const int nkb = min(cls.nkb, kb_begin + kb_per) - kb_begin;   // k-blocks of THIS CTA

  // ---- one-time setup ----------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (STAGES + s) * 8, 1);
    }
    mbar_init(bar_base + 2 * STAGES * 8, 1);
    if (SPLIT > 1) {
      mbar_init(recv_bar, 1);
      fence_mbar_init();
      mbar_expect_tx(recv_bar, (SPLIT - 1) * XCHG_BYTES);   // the one arrival; the phase completes when every peer's slice landed
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.bmap);
    tma_prefetch_desc(&p.amap[0]);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), MT * BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (SPLIT > 1) cluster_arrive();   // #1: my barriers exist (peers wait for it before they copy into this CTA)
  // everything above is independent of the predecessor kernel; from here on its outputs (activations, scale / shift) are read
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) dbg_mark(p, 1);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = MT * p.box_rows * 128 + B_BYTES;
      int kbi = 0;
      for (int ti = 0; ti < cls.tap_count; ++ti) {
        const TcTap tap = p.taps[cls.tap_begin + ti];
        const void* amap = &p.amap[tap.map];
        if (SPLIT > 1 && (kbi + tap.nchunk <= kb_begin || kbi >= kb_begin + nkb)) { kbi += tap.nchunk; continue; }
        for (int ch = 0; ch < tap.nchunk; ++ch, ++kbi) {
          if (SPLIT > 1 && (kbi < kb_begin || kbi >= kb_begin + nkb)) continue;
          mbar_wait(bar_base + (STAGES + stage) * 8, phase ^ 1u);
          const uint32_t full = bar_base + stage * 8;
          const uint32_t sa = base + stage * STAGE_BYTES;
          mbar_expect_tx(full, tx_bytes);
#pragma unroll
          for (int sub = 0; sub < MT; ++sub)
            tma_load_5d(sa + sub * A_BYTES, amap, full, (tap.c0 + ch) * 64, w0s[sub] + tap.dw, h0s[sub] + tap.dh, d0s[sub] + tap.dd,
                        n0s[sub]);
          tma_load_3d(sa + MT * A_BYTES, &p.bmap, full, tap.kofs + ch * 64, nt * BLOCK_N, p.b_batched ? n0s[0] : 0);
          if (kbi == kb_begin) dbg_mark(p, 2);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      dbg_mark(p, 9);
    }
    __syncwarp();
    if (SPLIT > 1) { cluster_wait(); cluster_arrive(); }   // keep this warp in step with the cluster barrier phases (#1, #2)
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar_base + stage * 8, phase);
        tc_fence_after();
        if (kb == 0) dbg_mark(p, 3);
        const uint32_t sa = base + stage * STAGE_BYTES;
        const uint64_t bdesc = umma_desc_sw128(sa + MT * A_BYTES, 16, 1024);
#pragma unroll
        for (int sub = 0; sub < MT; ++sub) {
          const uint64_t adesc = umma_desc_sw128(sa + sub * A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128-byte swizzle row: +2 in the (addr>>4) field
            tc_mma_bf16(tmem_base + sub * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(bar_base + (STAGES + stage) * 8);  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (nkb > 0) tc_commit(bar_base + 2 * STAGES * 8);  // accumulator ready
      dbg_mark(p, 4);
    }
    __syncwarp();
    if (SPLIT > 1) { cluster_wait(); cluster_arrive(); }
  } else if constexpr (SPLIT > 1) {
    // ================= split-K phase A: ship the chunks owned by peer CTAs =================
    // The fp32 partials are staged in THIS CTA's shared memory (the operand ring is free once the accumulator is complete)
    // in the exact layout of the receiver's buffer, then moved by one DSMEM bulk copy per peer that signals the peer's
    // mbarrier -- r01 shipped them with st.shared::cluster from registers, which cost ~3.9 us per launch (phase probe).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    {   // per-column epilogue vectors -> shared memory, under the K loop (their global-load latency is off the critical path)
      const int et0 = threadIdx.x - 64;
      for (int cc = et0; cc < BLOCK_N; cc += 128) {
        const int col = nt * BLOCK_N + cc;
        const bool in = col < p.cout;
        s_ep[cc] = (in && p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
        s_ep[BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.scale + col) : 1.f;
        s_ep[2 * BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.shift + col) : 0.f;
      }
    }
    if (nkb > 0) {
      mbar_wait(bar_base + 2 * STAGES * 8, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) dbg_mark(p, 5);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const uint32_t owner = SPLIT == 4 ? (uint32_t)c : (uint32_t)(c >> 1);
      if (owner == rank) continue;
      uint32_t rr[32];
      if (nkb > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = 0u;
      }
      // send area = start of the operand ring: one XCHG_BYTES block per peer, indexed like the peer's receive slots
      const uint32_t pidx = owner < rank ? owner : owner - 1;                       // the peer's index among my peers
      float* dst = reinterpret_cast<float*>(smem) + ((pidx * OWN_CHUNKS + (SPLIT == 4 ? 0 : (c & 1))) * 128 + row) * 32;
#pragma unroll
      for (int g = 0; g < 8; ++g)   // 16-byte groups swizzled by the row: conflict-free here and at the receiver
        *reinterpret_cast<float4*>(dst + ((g ^ (row & 7)) << 2)) =
            make_float4(__uint_as_float(rr[g * 4]), __uint_as_float(rr[g * 4 + 1]), __uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3]));
    }
    tc_fence_before();
    fence_proxy_async();                                  // my generic-proxy writes -> visible to the bulk-copy (async) proxy
    asm volatile("bar.sync 1, 128;" ::: "memory");        // all four epilogue warps have staged their rows (and s_ep)
    if (threadIdx.x == 64) dbg_mark(p, 11);
    cluster_wait();                                       // #1: every peer's receive barrier is initialised
    if (threadIdx.x == 64) dbg_mark(p, 12);
    if (threadIdx.x == 64) {
#pragma unroll
      for (int pj = 0; pj < SPLIT - 1; ++pj) {
        const uint32_t pi = (uint32_t)pj;
        const uint32_t peer = pi < rank ? pi : pi + 1;
        const uint32_t slot = rank < peer ? rank : rank - 1;                        // my index among that peer's peers
        bulk_copy_to_peer(mapa_shared(recv_base + slot * XCHG_BYTES, peer), base + pi * XCHG_BYTES, XCHG_BYTES, mapa_shared(recv_bar, peer));
      }
    }
    mbar_wait(recv_bar, 0);                               // the peers' partials of MY columns have landed
    cluster_arrive();                                     // #2: nothing more will be copied into / out of ... (see the end)
    if (threadIdx.x == 64) dbg_mark(p, 6);
  }
  if (warp >= 2) {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    if (SPLIT == 1) {   // per-column epilogue vectors -> shared memory once per CTA (overlaps the K loop; the chunk loop reads broadcasts)
      const int et0 = threadIdx.x - 64;
      for (int cc = et0; cc < BLOCK_N; cc += 128) {
        const int col = nt * BLOCK_N + cc;
        const bool in = col < p.cout;
        s_ep[cc] = (in && p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
        s_ep[BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.scale + col) : 1.f;
        s_ep[2 * BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.shift + col) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (SPLIT == 1 && nkb > 0) {
      mbar_wait(bar_base + 2 * STAGES * 8, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) dbg_mark(p, 10);
#pragma unroll 1
    for (int sub = 0; sub < MT; ++sub) {
    if (!live[sub]) break;
    const int w0 = w0s[sub], h0 = h0s[sub], d0 = d0s[sub], n0 = n0s[sub], mt = mts[sub];
    int r = row;
    const int iw = r % p.box[0];
    r /= p.box[0];
    const int ih = r % p.box[1];
    r /= p.box[1];
    const int id = r % p.box[2];
    const int in = r / p.box[2];
    const int ow = w0 + iw, oh = h0 + ih, od = d0 + id, on = n0 + in;
    const bool valid = row < p.box_rows && ow < p.ext[0] && oh < p.ext[1] && od < p.ext[2] && on < p.ext[3];
    const long long row_off = cls.out_ofs + ow * p.so[0] + oh * p.so[1] + od * p.so[2] + on * p.so[3];
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      const int col0 = nt * BLOCK_N + c * 32;
      if (col0 >= p.cout) break;  // warp-uniform
      if (SPLIT > 1 && (SPLIT == 4 ? (uint32_t)c : (uint32_t)(c >> 1)) != rank) continue;   // a peer CTA owns these columns
      uint32_t rr[32];
      if (nkb > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sub * BLOCK_N + c * 32, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = 0u;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
      if (SPLIT > 1) {   // add the partial accumulators the peers left in my receive buffer
#pragma unroll
        for (int sl = 0; sl < SPLIT - 1; ++sl) {
          const float* rb = reinterpret_cast<const float*>(smem + (recv_base - base)) + ((sl * OWN_CHUNKS + (SPLIT == 4 ? 0 : (c & 1))) * 128 + row) * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 f = *reinterpret_cast<const float4*>(rb + ((g ^ (row & 7)) << 2));
            v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
          }
        }
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += s_ep[c * 32 + j];
      }
      if (threadIdx.x == 64) dbg_mark(p, 13);
      if (want_stats) {
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float m = valid ? v[j] : 0.f;
          s1[j] = m;
          s2[j] = m * m;
        }
        const float c1 = warp_transpose_sum32(s1, lane);
        const float c2 = warp_transpose_sum32(s2, lane);
        s_stats[(q * 2 + 0) * BLOCK_N + c * 32 + lane] = c1;
        s_stats[(q * 2 + 1) * BLOCK_N + c * 32 + lane] = c2;
      }
      if (threadIdx.x == 64) dbg_mark(p, 14);
      if (p.scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_ep[BLOCK_N + c * 32 + j], s_ep[2 * BLOCK_N + c * 32 + j]);
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (valid) {
        const bool second = p.out2 != nullptr && col0 >= p.seg_split;
        void* const outp = second ? p.out2 : p.out;
        const int colx = second ? col0 - p.seg_split : col0;
        const int accum = second ? p.accumulate2 : p.accumulate;
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(outp) + row_off + colx;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (col0 + g * 4 < p.cout) {
              float4 f = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
              if (accum) {
                const float4 e = *reinterpret_cast<const float4*>(o + g * 4);
                f.x += e.x; f.y += e.y; f.z += e.z; f.w += e.w;
              }
              *reinterpret_cast<float4*>(o + g * 4) = f;
            }
          }
        } else {
          bf16* o = reinterpret_cast<bf16*>(outp) + row_off + colx;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (col0 + g * 8 < p.cout) {
              float w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = v[g * 8 + j];
              if (accum) {
                float e[8];
                Vec8<bf16>::load(o + g * 8, e);
#pragma unroll
                for (int j = 0; j < 8; ++j) w8[j] += e[j];
              }
              Vec8<bf16>::store(o + g * 8, w8);
            }
          }
        }
      }
    }
    if (threadIdx.x == 64) dbg_mark(p, 15);
    if (want_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = threadIdx.x - 64;  // 0..127
      const long long srow = static_cast<long long>(cls_id) * p.m_tiles + mt;
      for (int cc = et; cc < BLOCK_N; cc += 128) {
        const int col = nt * BLOCK_N + cc;
        const bool mine = SPLIT == 1 || (SPLIT == 4 ? (uint32_t)(cc >> 5) : (uint32_t)(cc >> 6)) == rank;
        if (col < p.cout && mine) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            a += s_stats[(qq * 2 + 0) * BLOCK_N + cc];
            b += s_stats[(qq * 2 + 1) * BLOCK_N + cc];
          }
          p.stats[(srow * 2 + 0) * p.cout + col] = a;
          p.stats[(srow * 2 + 1) * p.cout + col] = b;
        }
      }
    }
    if (want_stats && sub + 1 < MT) asm volatile("bar.sync 1, 128;" ::: "memory");  // s_stats is reused by the next sub-tile
    if (BLOCK_N == 128 && MT == 1 && p.fuse.y != nullptr) {
      // ================= BatchNorm finished in place (TcFuseBN) =================
      // (1) grid barrier: every CTA's statistics rows are in global memory.  All CTAs of the launch are resident (<= one per SM,
      //     checked by the host), so spinning is safe.
      const int et = threadIdx.x - 64;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        atomicAdd(p.gbar, 1u);
        unsigned int seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.gbar) : "memory");
        } while (seen < gridDim.x);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // (2) finalise my columns: fp64 sums over the statistics rows in a fixed order, as sap3d_bn_finalize does
      {
        const int cc = et;
        const int col = nt * BLOCK_N + cc;
        const bool mine = SPLIT == 1 || (SPLIT == 4 ? (uint32_t)(cc >> 5) : (uint32_t)(cc >> 6)) == rank;
        if (col < p.cout && mine) {
          double ta = 0.0, tb = 0.0;
          for (int r0 = 0; r0 < p.m_tiles; r0 += 8) {     // 16 loads in flight, then the sums in row order
            float fa[8], fb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool in = r0 + i < p.m_tiles;
              const long long r = in ? r0 + i : r0;
              fa[i] = __ldcg(p.stats + (r * 2 + 0) * p.cout + col);
              fb[i] = __ldcg(p.stats + (r * 2 + 1) * p.cout + col);
              if (!in) { fa[i] = 0.f; fb[i] = 0.f; }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ta += (double)fa[i];
              tb += (double)fb[i];
            }
          }
          const double m = ta / p.fuse.count;
          double vv = tb / p.fuse.count - m * m;
          if (vv < 0.0) vv = 0.0;
          const float mean = (float)m, var = (float)vv;
          const float rstd = rsqrtf(var + p.fuse.eps);
          const float g = p.fuse.gamma ? __ldg(p.fuse.gamma + col) : 1.f;
          const float bt = p.fuse.beta ? __ldg(p.fuse.beta + col) : 0.f;
          const float sc = g * rstd, sh = bt - mean * g * rstd;
          s_ep[BLOCK_N + cc] = sc;
          s_ep[2 * BLOCK_N + cc] = sh;
          if (mt == 0) {      // one CTA per column publishes the per-channel results
            p.fuse.scale[col] = sc;
            p.fuse.shift[col] = sh;
            if (p.fuse.mean) p.fuse.mean[col] = mean;
            if (p.fuse.rstd) p.fuse.rstd[col] = rstd;
            if (p.fuse.moving_mean) {
              p.fuse.moving_mean[col] = p.fuse.moving_mean[col] * p.fuse.momentum + mean * (1.f - p.fuse.momentum);
              p.fuse.moving_var[col] = p.fuse.moving_var[col] * p.fuse.momentum + var * (1.f - p.fuse.momentum);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // (3) second pass over the accumulator (still in TMEM; the peers' partials still in the receive buffer)
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int col0 = nt * BLOCK_N + c * 32;
        if (col0 >= p.cout) break;
        if (SPLIT > 1 && (SPLIT == 4 ? (uint32_t)c : (uint32_t)(c >> 1)) != rank) continue;
        uint32_t rr[32];
        if (nkb > 0) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rr[j] = 0u;
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
        if (SPLIT > 1) {
#pragma unroll
          for (int sl = 0; sl < SPLIT - 1; ++sl) {
            const float* rb = reinterpret_cast<const float*>(smem + (recv_base - base)) + ((sl * OWN_CHUNKS + (SPLIT == 4 ? 0 : (c & 1))) * 128 + row) * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 f = *reinterpret_cast<const float4*>(rb + ((g ^ (row & 7)) << 2));
              v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
            }
          }
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += s_ep[c * 32 + j];
        }
        // the backward pass normalises the STORED (bf16) raw tensor: use the same rounded value here
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_ep[BLOCK_N + c * 32 + j], s_ep[2 * BLOCK_N + c * 32 + j]);
        if (p.fuse.relu1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (valid) {
          if (p.fuse.residual != nullptr) {
            const bf16* rp = reinterpret_cast<const bf16*>(p.fuse.residual) + row_off + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (col0 + g * 8 < p.cout) {
                float e[8];
                Vec8<bf16>::load(rp + g * 8, e);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] += e[j];
              }
            }
          }
          if (p.fuse.relu_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          bf16* o = reinterpret_cast<bf16*>(p.fuse.y) + row_off + col0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (col0 + g * 8 < p.cout) {
              float w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = v[g * 8 + j];
              Vec8<bf16>::store(o + g * 8, w8);
            }
          }
        }
      }
      // (4) hand the barrier slot back: the last CTA to get here zeroes it for its next user
      if (et == 0) {
        const unsigned int prev = atomicAdd(p.gbar + 1, 1u);
        if (prev == gridDim.x - 1) {
          p.gbar[0] = 0u;
          p.gbar[1] = 0u;
          __threadfence();
        }
      }
    }
    }  // sub
    if (threadIdx.x == 64) dbg_mark(p, 7);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, MT * BLOCK_N);
  }
  // #2: every CTA of the cluster has received all its slices, i.e. every bulk copy that READS this CTA's shared memory has
  // completed -- only now may the CTA exit and its shared memory be handed to another block
  if (SPLIT > 1) cluster_wait();
  if (threadIdx.x == 0) dbg_mark(p, 8);
}


// =================================================================================================
// Persistent form for layers with more work units than SMs (the decoder): one CTA per SM takes units from a ticket
// counter (r01: u = blockIdx.x, blockIdx.x + gridDim.x, ...).  The accumulators are DOUBLE-BUFFERED in TMEM
// (2 x MT x BLOCK_N columns when that fits in 512), so the epilogue of unit i (tcgen05.ld, bias, statistics, stores) runs
// under the MMAs of unit i+1, the TMA ring never drains between units, and the wave-quantisation tail of a 5.3-wave
// launch disappears.
// =================================================================================================
// ticket counter of the persistent kernels: every CTA's last act.  All of a CTA's draws precede its call, so when the last CTA
// arrives no draw is outstanding and the pair can be zeroed for the slot's next user.
SAP3D_DEVINL void sched_finish(unsigned int* sched) {
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(sched + 1, 1u);
    if (prev == gridDim.x - 1) {
      sched[0] = 0u;
      sched[1] = 0u;
      __threadfence();
    }
  }
}

// HALO = 0: every k-block (tap, 64-channel chunk) is one ring stage holding MT activation boxes and one weight tile.
// HALO = NB > 0 ("halo tile"): filter taps come in triples that differ only by -1/0/+1 along the OUTERMOST axis of the
// output box (p.halo_inner rows per step, a multiple of 8).  One TMA box with two extra steps along that axis
// (p.halo_rows = 128 + 2 * halo_inner rows) then contains the activation tile of all three taps as contiguous, swizzle-atom
// aligned 128-row windows: the A operand of tap j is the SAME shared-memory buffer at +j * halo_inner * 128 bytes.  The
// activation bytes pulled over the L2->SM fabric drop from 3 x 16 KB to halo_rows x 128 B (20 KB at halo_inner = 16) per
// sub-tile and tap triple -- the fabric is what bounds the non-halo kernel (ncu: 4.17 GB per launch of the dominant decoder
// conv at ~14 TB/s).  Activations and weights then have separate rings: STAGES halo buffers, HALO weight tiles.
// Units are handed out DYNAMICALLY: the producer thread of every CTA draws the next unit index from a ticket counter in global
// memory (p.sched) and passes it to the MMA and epilogue warps through a small shared-memory queue.  A CTA that gets its SM late
// (the SM was held by a filter-gradient CTA of the side stream, or by an NCCL kernel) simply draws fewer units instead of sitting
// on a fixed share of the work while the others idle; the static form doubled the kernel's duration whenever a few of the 148
// CTAs could not be resident at once.  p.balanced (one class, one column tile): units are groups of MT consecutive m-tiles
// followed by about one single m-tile per SM, so the ragged end of the kernel is one sub-tile long; otherwise units are
// (class, MT-group, column tile) triples.  Sub-tiles that are not live are neither loaded nor multiplied.
template <int BLOCK_N, int STAGES, int MT, int HALO = 0>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_persist_kernel(const __grid_constant__ TcConvParams p) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  constexpr int HALO_BUF_BYTES = TC_HALO_MAX_ROWS * 128;          // one sub-tile's halo box
  constexpr int A_RING_BYTES = MT * HALO_BUF_BYTES;
  constexpr int RING_BYTES = HALO ? STAGES * A_RING_BYTES + HALO * B_BYTES : STAGES * STAGE_BYTES;
  constexpr int NBUF = 2 * MT * BLOCK_N <= 512 ? 2 : 1;
  constexpr int TCOLS = NBUF * MT * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + RING_BYTES;
  // HALO = 0: full[STAGES], empty[STAGES], tfull[2], tempty[2]
  // HALO > 0: afull[STAGES], aempty[STAGES], bfull[HALO], bempty[HALO], tfull[2], tempty[2]
  constexpr int NRING_BAR = HALO ? 2 * STAGES + 2 * HALO : 2 * STAGES;
  constexpr int NBAR = NRING_BAR + 4 + 2 * TC_QD;   // ..., tfull[2], tempty[2], qfull[TC_QD], qempty[TC_QD]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RING_BYTES + NBAR * 8);
  int* s_q = reinterpret_cast<int*>(smem + RING_BYTES + NBAR * 8 + 16);           // unit queue [TC_QD]
  float* s_stats = reinterpret_cast<float*>(smem + RING_BYTES + NBAR * 8 + 32);  // [4][2][BLOCK_N]
  float* s_ep = s_stats + 4 * 2 * BLOCK_N;                                        // [3][BLOCK_N]
  const uint32_t tfull = bar_base + NRING_BAR * 8, tempty = tfull + 16;
  const uint32_t qfull = tempty + 16, qempty = qfull + TC_QD * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_groups = (p.m_tiles + MT - 1) / MT;

  struct Unit { int nt, cls_id, nsub; int mts[MT], w0s[MT], h0s[MT], d0s[MT], n0s[MT]; };
  struct Cursor { int qi; uint32_t qp; };
  // producer = true: draws a ticket and publishes it; false: takes the next published unit.  Returns false when the work is gone.
  auto next_unit = [&](Cursor& cq, Unit& t, const bool producer) -> bool {
    int u;
    if (producer) {
      mbar_wait(qempty + cq.qi * 8, cq.qp ^ 1u);
      u = static_cast<int>(atomicAdd(p.sched, 1u));
      if (u >= p.n_units) u = -1;
      s_q[cq.qi] = u;
      mbar_arrive(qfull + cq.qi * 8);
    } else {
      mbar_wait(qfull + cq.qi * 8, cq.qp);
      u = s_q[cq.qi];
      mbar_arrive(qempty + cq.qi * 8);
    }
    if (++cq.qi == TC_QD) { cq.qi = 0; cq.qp ^= 1u; }
    if (u < 0) return false;
    int mt0;
    if (p.balanced) {
      t.nt = 0;
      t.cls_id = 0;
      if (u < p.n_pair_units) {
        mt0 = u * MT;
        t.nsub = MT;
      } else {
        mt0 = p.n_pair_units * MT + (u - p.n_pair_units);
        t.nsub = 1;
      }
    } else {
      t.nt = u % p.n_tiles;
      u /= p.n_tiles;
      t.cls_id = u / m_groups;
      mt0 = (u - t.cls_id * m_groups) * MT;
      t.nsub = min(MT, p.m_tiles - mt0);
    }
#pragma unroll
    for (int sub = 0; sub < MT; ++sub) {
      int r = min(mt0 + sub, p.m_tiles - 1);
      t.mts[sub] = r;
      const int tw = r % p.tiles[0]; r /= p.tiles[0];
      const int th = r % p.tiles[1]; r /= p.tiles[1];
      const int td = r % p.tiles[2];
      const int tn = r / p.tiles[2];
      t.w0s[sub] = tw * p.box[0]; t.h0s[sub] = th * p.box[1]; t.d0s[sub] = td * p.box[2]; t.n0s[sub] = tn * p.box[3];
    }
    return true;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NRING_BAR; ++s) mbar_init(bar_base + s * 8, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b * 8, 1);
      mbar_init(tempty + b * 8, 128);
    }
    for (int s = 0; s < TC_QD; ++s) {
      mbar_init(qfull + s * 8, 1);
      mbar_init(qempty + s * 8, 129);   // the MMA thread + 128 epilogue threads
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.bmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // set-up above overlaps the predecessor's tail; its outputs are read from here on
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer: one continuous ring over all units =================
    if (lane == 0) {
      Cursor cur = {0, 0u};
      Unit t;
      if (HALO) {
        const uint32_t afull = bar_base, aempty = bar_base + STAGES * 8, bfull = bar_base + 2 * STAGES * 8, bempty = bfull + HALO * 8;
        const uint32_t b_ring = base + STAGES * A_RING_BYTES;
        const uint32_t halo_bytes = p.halo_rows * 128;
        int sa_i = 0, sb_i = 0;
        uint32_t pa = 0, pb = 0;
        while (next_unit(cur, t, true)) {
          const TcClass cls = p.cls[t.cls_id];
          for (int g = 0; g < cls.tap_count; g += 3) {
            const TcTap tap = p.taps[cls.tap_begin + g];       // lowest offset along the halo axis: the box origin
            const int k1 = p.taps[cls.tap_begin + g + 1].kofs, k2 = p.taps[cls.tap_begin + g + 2].kofs;
            const void* amap = &p.amap[tap.map];
            for (int ch = 0; ch < tap.nchunk; ++ch) {
              mbar_wait(aempty + sa_i * 8, pa ^ 1u);
              mbar_expect_tx(afull + sa_i * 8, t.nsub * halo_bytes);
#pragma unroll
              for (int sub = 0; sub < MT; ++sub)
                if (sub < t.nsub)
                  tma_load_5d(base + sa_i * A_RING_BYTES + sub * HALO_BUF_BYTES, amap, afull + sa_i * 8, (tap.c0 + ch) * 64, t.w0s[sub] + tap.dw,
                              t.h0s[sub] + tap.dh, t.d0s[sub] + tap.dd, t.n0s[sub]);
              if (++sa_i == STAGES) { sa_i = 0; pa ^= 1u; }
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int kofs = j == 0 ? tap.kofs : (j == 1 ? k1 : k2);
                mbar_wait(bempty + sb_i * 8, pb ^ 1u);
                mbar_expect_tx(bfull + sb_i * 8, B_BYTES);
                tma_load_3d(b_ring + sb_i * B_BYTES, &p.bmap, bfull + sb_i * 8, kofs + ch * 64, t.nt * BLOCK_N, 0);
                if (++sb_i == HALO) { sb_i = 0; pb ^= 1u; }
              }
            }
          }
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        while (next_unit(cur, t, true)) {
          const uint32_t tx_bytes = t.nsub * p.box_rows * 128 + B_BYTES;
          const TcClass cls = p.cls[t.cls_id];
          for (int ti = 0; ti < cls.tap_count; ++ti) {
            const TcTap tap = p.taps[cls.tap_begin + ti];
            const void* amap = &p.amap[tap.map];
            for (int ch = 0; ch < tap.nchunk; ++ch) {
              mbar_wait(bar_base + (STAGES + stage) * 8, phase ^ 1u);
              const uint32_t full = bar_base + stage * 8;
              const uint32_t sa = base + stage * STAGE_BYTES;
              mbar_expect_tx(full, tx_bytes);
#pragma unroll
              for (int sub = 0; sub < MT; ++sub)
                if (sub < t.nsub)
                  tma_load_5d(sa + sub * A_BYTES, amap, full, (tap.c0 + ch) * 64, t.w0s[sub] + tap.dw, t.h0s[sub] + tap.dh, t.d0s[sub] + tap.dd,
                              t.n0s[sub]);
              tma_load_3d(sa + MT * A_BYTES, &p.bmap, full, tap.kofs + ch * 64, t.nt * BLOCK_N, p.b_batched ? t.n0s[0] : 0);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      Cursor cur = {0, 0u};
      int it = 0;
      Unit t;
      int stage = 0, sb_i = 0;
      uint32_t phase = 0, pb = 0;
      for (; next_unit(cur, t, false); ++it) {
        const int buf = it % NBUF;
        const uint32_t use = static_cast<uint32_t>(it / NBUF);
        const int nkb = p.cls[t.cls_id].nkb;
        mbar_wait(tempty + buf * 8, (use & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * MT * BLOCK_N;
        if (HALO) {
          const uint32_t afull = bar_base, aempty = bar_base + STAGES * 8, bfull = bar_base + 2 * STAGES * 8, bempty = bfull + HALO * 8;
          const uint32_t b_ring = base + STAGES * A_RING_BYTES;
          const uint32_t tap_step = p.halo_inner * 128;          // bytes between the windows of consecutive taps (multiple of 1024)
          for (int kg = 0; kg < nkb; kg += 3) {                  // one halo buffer = three k-blocks
            mbar_wait(afull + stage * 8, phase);
            const uint32_t sa = base + stage * A_RING_BYTES;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              mbar_wait(bfull + sb_i * 8, pb);
              tc_fence_after();
              const uint64_t bdesc = umma_desc_sw128(b_ring + sb_i * B_BYTES, 16, 1024);
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
                if (sub < t.nsub) {
                  const uint64_t adesc = umma_desc_sw128(sa + sub * HALO_BUF_BYTES + j * tap_step, 16, 1024);
#pragma unroll
                  for (int k = 0; k < 4; ++k) tc_mma_bf16(tacc + sub * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kg | j | k) != 0 ? 1u : 0u);
                }
              }
              tc_commit(bempty + sb_i * 8);
              if (++sb_i == HALO) { sb_i = 0; pb ^= 1u; }
            }
            tc_commit(aempty + stage * 8);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        } else {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(bar_base + stage * 8, phase);
            tc_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES;
            const uint64_t bdesc = umma_desc_sw128(sa + MT * A_BYTES, 16, 1024);
#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {
              if (sub < t.nsub) {
                const uint64_t adesc = umma_desc_sw128(sa + sub * A_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_bf16(tacc + sub * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(bar_base + (STAGES + stage) * 8);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (nkb > 0) tc_commit(tfull + buf * 8);
        else mbar_arrive(tfull + buf * 8);               // bias-only class: nothing to wait for
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5), overlapped with the next unit's MMAs =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    const int et = threadIdx.x - 64;
    Cursor cur = {0, 0u};
    int it = 0;
    Unit t;
    for (; next_unit(cur, t, false); ++it) {
      const TcClass cls = p.cls[t.cls_id];
      const int nkb = cls.nkb;
      const int buf = it % NBUF;
      const uint32_t use = static_cast<uint32_t>(it / NBUF);
      for (int cc = et; cc < BLOCK_N; cc += 128) {     // per-column epilogue vectors of this unit
        const int col = t.nt * BLOCK_N + cc;
        const bool in = col < p.cout;
        s_ep[cc] = (in && p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
        s_ep[BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.scale + col) : 1.f;
        s_ep[2 * BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.shift + col) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tfull + buf * 8, use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * MT * BLOCK_N;
#pragma unroll 1
      for (int sub = 0; sub < t.nsub; ++sub) {
        int r = row;
        const int iw = r % p.box[0]; r /= p.box[0];
        const int ih = r % p.box[1]; r /= p.box[1];
        const int id = r % p.box[2];
        const int in = r / p.box[2];
        const int ow = t.w0s[sub] + iw, oh = t.h0s[sub] + ih, od = t.d0s[sub] + id, on = t.n0s[sub] + in;
        const bool valid = row < p.box_rows && ow < p.ext[0] && oh < p.ext[1] && od < p.ext[2] && on < p.ext[3];
        const long long row_off = cls.out_ofs + ow * p.so[0] + oh * p.so[1] + od * p.so[2] + on * p.so[3];
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int col0 = t.nt * BLOCK_N + c * 32;
          if (col0 >= p.cout) break;
          uint32_t rr[32];
          if (nkb > 0) {
            tmem_ld_32x32(tacc + (static_cast<uint32_t>(q * 32) << 16) + sub * BLOCK_N + c * 32, rr);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0u;
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += s_ep[c * 32 + j];
          }
          if (want_stats) {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float m = valid ? v[j] : 0.f;
              s1[j] = m;
              s2[j] = m * m;
            }
            const float c1 = warp_transpose_sum32(s1, lane);
            const float c2 = warp_transpose_sum32(s2, lane);
            s_stats[(q * 2 + 0) * BLOCK_N + c * 32 + lane] = c1;
            s_stats[(q * 2 + 1) * BLOCK_N + c * 32 + lane] = c2;
          }
          if (p.scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_ep[BLOCK_N + c * 32 + j], s_ep[2 * BLOCK_N + c * 32 + j]);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (valid) {
            const bool second = p.out2 != nullptr && col0 >= p.seg_split;
            void* const outp = second ? p.out2 : p.out;
            const int colx = second ? col0 - p.seg_split : col0;
            const int accum = second ? p.accumulate2 : p.accumulate;
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(outp) + row_off + colx;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                if (col0 + g * 4 < p.cout) {
                  float4 f = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                  if (accum) {
                    const float4 e = *reinterpret_cast<const float4*>(o + g * 4);
                    f.x += e.x; f.y += e.y; f.z += e.z; f.w += e.w;
                  }
                  *reinterpret_cast<float4*>(o + g * 4) = f;
                }
              }
            } else {
              bf16* o = reinterpret_cast<bf16*>(outp) + row_off + colx;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (col0 + g * 8 < p.cout) {
                  float w8[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) w8[j] = v[g * 8 + j];
                  if (accum) {
                    float e[8];
                    Vec8<bf16>::load(o + g * 8, e);
#pragma unroll
                    for (int j = 0; j < 8; ++j) w8[j] += e[j];
                  }
                  Vec8<bf16>::store(o + g * 8, w8);
                }
              }
            }
          }
        }
        if (sub + 1 == t.nsub) {
          // last sub-tile of the unit has been read out of TMEM: hand the buffer back before the statistics tail
          tc_fence_before();
          mbar_arrive(tempty + buf * 8);
        }
        if (want_stats) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const long long srow = static_cast<long long>(t.cls_id) * p.m_tiles + t.mts[sub];
          for (int cc = et; cc < BLOCK_N; cc += 128) {
            const int col = t.nt * BLOCK_N + cc;
            if (col < p.cout) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                a += s_stats[(qq * 2 + 0) * BLOCK_N + cc];
                b += s_stats[(qq * 2 + 1) * BLOCK_N + cc];
              }
              p.stats[(srow * 2 + 0) * p.cout + col] = a;
              p.stats[(srow * 2 + 1) * p.cout + col] = b;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // s_stats / s_ep are reused by the next sub-tile / unit
        }
      }
      if (!want_stats) asm volatile("bar.sync 1, 128;" ::: "memory");   // s_ep is rewritten at the top of the next unit
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
  sched_finish(p.sched);
}

// -------------------------------------------------------------------------------------------------------------------
// Persistent halo-tile kernel with SWAPPED operands, for plain convolutions with 64 < cout <= 128:
//     D[cout (128 TMEM lanes)][256 positions (columns)] = W[128 x K] * X[256 x K]^T
// With the positions as the M side every tcgen05.mma is 128 x 128 x 16 and reads 4 KB of activations plus 4 KB of weights
// from shared memory; with the positions as the N side one instruction covers TWO 128-position tiles (128 x 256 x 16: 4 KB of
// weights + 8 KB of activations for twice the math), which is the form the 256-column data-gradient launches of the same
// layers already run at (~1.45 vs ~1.23 PFLOP/s stand-alone).  A unit is a PAIR of 128-position tiles that are neighbours
// along H (tile order is H-fastest here), fetched as ONE box of twice the H extent plus the two halo steps along the outermost
// axis (<= 320 rows = 40 KB), so the window of tap j is 256 contiguous, swizzle-atom aligned rows.  The last units of the ticket
// sequence are single tiles (N = 128, second set of tensor maps: amap[TC_SWAP_SINGLE_MAP + view]) so that the ragged end of the
// kernel is one tile long.
// The accumulator is the TRANSPOSE of the output tile: epilogue thread = channel, so bias / affine / ReLU constants and the
// BatchNorm sums are per-thread scalars (no warp transposes); 32 positions at a time go through a [32][128] bf16 staging tile
// in shared memory and leave as 256-byte rows (16-byte stores).
// -------------------------------------------------------------------------------------------------------------------
constexpr int TC_SWAP_SINGLE_MAP = 4;     // amap[v] = pair box of view v, amap[4 + v] = single-tile box
constexpr int TC_SWAP_MAX_ROWS = 320;     // rows of a pair box incl. halo (256 + 2 * 32)

template <int NA, int NW>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_swap_kernel(const __grid_constant__ TcConvParams p) {
  constexpr int ACT_BYTES = TC_SWAP_MAX_ROWS * 128;     // one activation ring slot (pair box with halo)
  constexpr int W_BYTES = 128 * 128;                    // one weight tile: 128 couts x 64 channels
  constexpr int RING_BYTES = NA * ACT_BYTES + NW * W_BYTES;
  constexpr int STAGE_TILE = 32 * 128 * 2;              // [32 positions][128 channels] bf16
  constexpr int NBAR = 2 * NA + 2 * NW + 4 + 2 * TC_QD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  uint8_t* s_stage = smem + RING_BYTES;                                            // [2][32][128] bf16
  long long* s_off = reinterpret_cast<long long*>(smem + RING_BYTES + 2 * STAGE_TILE);   // [2][256] output element offset, -1 = outside
  const uint32_t bar_base = base + RING_BYTES + 2 * STAGE_TILE + 2 * 256 * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RING_BYTES + 2 * STAGE_TILE + 2 * 256 * 8 + NBAR * 8);
  int* s_q = reinterpret_cast<int*>(smem + RING_BYTES + 2 * STAGE_TILE + 2 * 256 * 8 + NBAR * 8 + 16);   // unit queue [TC_QD]
  const uint32_t afull = bar_base, aempty = afull + NA * 8, wfull = aempty + NA * 8, wempty = wfull + NW * 8;
  const uint32_t tfull = wempty + NW * 8, tempty = tfull + 16;
  const uint32_t qfull = tempty + 16, qempty = qfull + TC_QD * 8;
  const uint32_t w_ring = base + NA * ACT_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  struct Unit { int nsub, mt0, w0, h0, d0, n0; };
  struct Cursor { int qi; uint32_t qp; };
  // units come from the ticket counter (see conv_tc_persist_kernel): [0, n_pair_units) are pairs of H-neighbour tiles (tiles are
  // numbered H-fastest and tiles[1] is even, so tile 2u and 2u + 1 are neighbours), the rest single tiles
  auto next_unit = [&](Cursor& cq, Unit& t, const bool producer) -> bool {
    int u;
    if (producer) {
      mbar_wait(qempty + cq.qi * 8, cq.qp ^ 1u);
      u = static_cast<int>(atomicAdd(p.sched, 1u));
      if (u >= p.n_units) u = -1;
      s_q[cq.qi] = u;
      mbar_arrive(qfull + cq.qi * 8);
    } else {
      mbar_wait(qfull + cq.qi * 8, cq.qp);
      u = s_q[cq.qi];
      mbar_arrive(qempty + cq.qi * 8);
    }
    if (++cq.qi == TC_QD) { cq.qi = 0; cq.qp ^= 1u; }
    if (u < 0) return false;
    if (u < p.n_pair_units) {
      t.mt0 = 2 * u;
      t.nsub = 2;
    } else {
      t.mt0 = 2 * p.n_pair_units + (u - p.n_pair_units);
      t.nsub = 1;
    }
    int r = t.mt0;
    const int th = r % p.tiles[1]; r /= p.tiles[1];
    const int tw = r % p.tiles[0]; r /= p.tiles[0];
    const int td = r % p.tiles[2];
    const int tn = r / p.tiles[2];
    t.w0 = tw * p.box[0]; t.h0 = th * p.box[1]; t.d0 = td * p.box[2]; t.n0 = tn * p.box[3];
    return true;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2 * NA + 2 * NW; ++s) mbar_init(bar_base + s * 8, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b * 8, 1);
      mbar_init(tempty + b * 8, 128);
    }
    for (int s = 0; s < TC_QD; ++s) {
      mbar_init(qfull + s * 8, 1);
      mbar_init(qempty + s * 8, 129);   // the MMA thread + 128 epilogue threads
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.bmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  const TcClass cls = p.cls[0];
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      Cursor cur = {0, 0u};
      int sa_i = 0, sw_i = 0;
      uint32_t pa = 0, pw = 0;
      Unit t;
      while (next_unit(cur, t, true)) {
        const uint32_t act_bytes = (128 + 2 * p.halo_inner) * t.nsub * 128;
        const int map_ofs = t.nsub == 2 ? 0 : TC_SWAP_SINGLE_MAP;
        for (int g = 0; g < cls.tap_count; g += 3) {
          const TcTap tap = p.taps[cls.tap_begin + g];
          const int k1 = p.taps[cls.tap_begin + g + 1].kofs, k2 = p.taps[cls.tap_begin + g + 2].kofs;
          const void* amap = &p.amap[map_ofs + tap.map];
          for (int ch = 0; ch < tap.nchunk; ++ch) {
            mbar_wait(aempty + sa_i * 8, pa ^ 1u);
            mbar_expect_tx(afull + sa_i * 8, act_bytes);
            tma_load_5d(base + sa_i * ACT_BYTES, amap, afull + sa_i * 8, (tap.c0 + ch) * 64, t.w0 + tap.dw, t.h0 + tap.dh, t.d0 + tap.dd, t.n0);
            if (++sa_i == NA) { sa_i = 0; pa ^= 1u; }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int kofs = j == 0 ? tap.kofs : (j == 1 ? k1 : k2);
              mbar_wait(wempty + sw_i * 8, pw ^ 1u);
              mbar_expect_tx(wfull + sw_i * 8, W_BYTES);
              tma_load_3d(w_ring + sw_i * W_BYTES, &p.bmap, wfull + sw_i * 8, kofs + ch * 64, 0, 0);
              if (++sw_i == NW) { sw_i = 0; pw ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 256, 0, 0), idesc1 = umma_idesc_bf16(128, 128, 0, 0);
      Cursor cur = {0, 0u};
      int it = 0, sa_i = 0, sw_i = 0;
      uint32_t pa = 0, pw = 0;
      Unit t;
      for (; next_unit(cur, t, false); ++it) {
        const int buf = it & 1;
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        const uint32_t idesc = t.nsub == 2 ? idesc2 : idesc1;
        const uint32_t tap_step = p.halo_inner * t.nsub * 128;   // bytes between the windows of consecutive taps
        mbar_wait(tempty + buf * 8, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * 256;
        for (int kg = 0; kg < cls.nkb; kg += 3) {
          mbar_wait(afull + sa_i * 8, pa);
          const uint32_t sx = base + sa_i * ACT_BYTES;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            mbar_wait(wfull + sw_i * 8, pw);
            tc_fence_after();
            const uint64_t wdesc = umma_desc_sw128(w_ring + sw_i * W_BYTES, 16, 1024);
            const uint64_t xdesc = umma_desc_sw128(sx + j * tap_step, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_bf16(tacc, wdesc + 2 * k, xdesc + 2 * k, idesc, (kg | j | k) != 0 ? 1u : 0u);
            tc_commit(wempty + sw_i * 8);
            if (++sw_i == NW) { sw_i = 0; pw ^= 1u; }
          }
          tc_commit(aempty + sa_i * 8);
          if (++sa_i == NA) { sa_i = 0; pa ^= 1u; }
        }
        tc_commit(tfull + buf * 8);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5): thread = output channel =================
    const int q = warp & 3;
    const int c = q * 32 + lane;          // TMEM lane = output channel
    const int et = threadIdx.x - 64;      // 0..127
    const bool cvalid = c < p.cout;
    const bool want_stats = p.stats != nullptr;
    const float bias_c = (cvalid && p.bias != nullptr) ? __ldg(p.bias + c) : 0.f;
    const float scale_c = (cvalid && p.scale != nullptr) ? __ldg(p.scale + c) : 1.f;
    const float shift_c = (cvalid && p.scale != nullptr) ? __ldg(p.shift + c) : 0.f;
    const bool affine = p.scale != nullptr;
    bf16* const out = reinterpret_cast<bf16*>(p.out);
    Cursor cur = {0, 0u};
    int it = 0, sb = 0;   // sb: staging tile in use
    Unit t;
    for (; next_unit(cur, t, false); ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      const int npos = 128 * t.nsub;
      long long* const offs = s_off + (it & 1) * 256;
      // output offsets of the unit's positions (box order: w fastest, then h over nsub * box[1] rows, then d)
      for (int i = et; i < npos; i += 128) {
        int r = i;
        const int iw = r % p.box[0]; r /= p.box[0];
        const int bh = p.box[1] * t.nsub;
        const int ih = r % bh; r /= bh;
        const int id = r;
        const int ow = t.w0 + iw, oh = t.h0 + ih, od = t.d0 + id;
        const bool valid = ow < p.ext[0] && oh < p.ext[1] && od < p.ext[2] && id < p.box[2];
        offs[i] = valid ? cls.out_ofs + ow * p.so[0] + oh * p.so[1] + od * p.so[2] + t.n0 * p.so[3] : -1ll;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tfull + buf * 8, use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * 256 + (static_cast<uint32_t>(q * 32) << 16);
      float s1 = 0.f, s2 = 0.f;
      const int nchunk = npos / 32;
#pragma unroll 1
      for (int ck = 0; ck < nchunk; ++ck) {
        uint32_t rr[32];
        tmem_ld_32x32(tacc + ck * 32, rr);
        tmem_ld_wait();
        if (ck + 1 == nchunk) {             // the accumulator has been read out: hand the buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(tempty + buf * 8);
        }
        bf16* const st = reinterpret_cast<bf16*>(s_stage + sb * STAGE_TILE);
        const uint32_t vmask = __ballot_sync(0xffffffffu, offs[ck * 32 + lane] >= 0);   // bit j: position j of the chunk is inside
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(rr[j]) + bias_c;
          if (want_stats) {
            const float m = ((vmask >> j) & 1u) ? v : 0.f;
            s1 += m;
            s2 = fmaf(m, m, s2);
          }
          if (affine) v = fmaf(v, scale_c, shift_c);
          if (p.relu) v = fmaxf(v, 0.f);
          st[j * 128 + c] = __float2bfloat16_rn(v);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 32 positions x 256 B: thread -> (row et / 16 of each group of 8 positions, 16-byte segment et % 16)
        const int seg = et & 15;
        if (seg * 8 < p.cout) {
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int pl = ps * 8 + (et >> 4);
            const long long o = offs[ck * 32 + pl];
            if (o >= 0) {
              uint4 val = *reinterpret_cast<const uint4*>(st + pl * 128 + seg * 8);
              bf16* dst = out + o + seg * 8;
              if (p.accumulate) {
                float e[8], w8[8];
                Vec8<bf16>::load(dst, e);
                Vec8<bf16>::unpack(val, w8);
#pragma unroll
                for (int j = 0; j < 8; ++j) w8[j] += e[j];
                Vec8<bf16>::store(dst, w8);
              } else {
                *reinterpret_cast<uint4*>(dst) = val;
              }
            }
          }
        }
        sb ^= 1;   // the other staging tile is free: its readers passed the bar.sync of the previous chunk
        if (want_stats && (ck & 3) == 3) {   // 128 positions done: one statistics row per 128-position tile
          if (cvalid) {
            const long long srow = t.mt0 + (ck >> 2);
            p.stats[(srow * 2 + 0) * p.cout + c] = s1;
            p.stats[(srow * 2 + 1) * p.cout + c] = s2;
          }
          s1 = 0.f;
          s2 = 0.f;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  sched_finish(p.sched);
}

// -------------------------------------------------------------------------------------------------------------------
// Persistent kernel, 2-CTA cluster with MULTICAST weight tiles (single-class problems = plain convolutions).
// The persistent kernel above is bound by the L2->SM crossbar (ncu: 4.17 GB per launch of the dominant decoder conv, 14.4
// TB/s): per 64-channel k-block a CTA pulls MT x 16 KB of activations and BLOCK_N x 128 B of weights, and every CTA pulls
// the SAME weights.  Here two CTAs of a cluster work on different row groups in lock-step; each issues one TMA multicast for
// HALF of the weight tile, so the weights cross the fabric once per cluster (-17 % bytes at BLOCK_N = 128, MT = 2).
// A stage is refilled only after both CTAs' MMAs have read it: the MMA warp's commit arrives on the `empty` barrier of BOTH
// CTAs (tcgen05.commit ... multicast::cluster), whose count is 2.
// -------------------------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, int MT, int CL>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_persist_mc_kernel(const __grid_constant__ TcConvParams p, const int total_pairs) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  constexpr int NBUF = 2 * MT * BLOCK_N <= 512 ? 2 : 1;
  constexpr int TCOLS = NBUF * MT * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  constexpr int NBAR = 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + NBAR * 8);
  float* s_stats = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + NBAR * 8 + 16);  // [4][2][BLOCK_N]
  float* s_ep = s_stats + 4 * 2 * BLOCK_N;                                                  // [3][BLOCK_N]
  const uint32_t tfull = bar_base + 2 * STAGES * 8, tempty = tfull + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_groups = (p.m_tiles + MT - 1) / MT;

  struct Unit { int nt, cls_id; int mts[MT], w0s[MT], h0s[MT], d0s[MT], n0s[MT]; bool live[MT]; };
  // work is dealt in PAIRS of row groups with the same n-tile: the two CTAs of a cluster walk identical (tap, chunk)
  // schedules on different rows, so every weight tile is fetched once per cluster (each CTA multicasts one half of it)
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = gridDim.x / CL, cluster_id = blockIdx.x / CL;
  constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1u);
  auto decode = [&](int q, Unit& t) {
    t.nt = q % p.n_tiles;
    t.cls_id = 0;
    int mg = (q / p.n_tiles) * CL + static_cast<int>(crank);
    const bool dummy = mg >= m_groups;      // odd group count: the last pair's second CTA keeps the pipeline in step on a
    if (dummy) mg = m_groups - 1;           // duplicate of the last group and stores nothing
#pragma unroll
    for (int sub = 0; sub < MT; ++sub) {
      int mt = mg * MT + sub;
      t.live[sub] = !dummy && mt < p.m_tiles;
      if (!t.live[sub]) mt = p.m_tiles - 1;
      t.mts[sub] = mt;
      int r = mt;
      const int tw = r % p.tiles[0]; r /= p.tiles[0];
      const int th = r % p.tiles[1]; r /= p.tiles[1];
      const int td = r % p.tiles[2];
      const int tn = r / p.tiles[2];
      t.w0s[sub] = tw * p.box[0]; t.h0s[sub] = th * p.box[1]; t.d0s[sub] = td * p.box[2]; t.n0s[sub] = tn * p.box[3];
    }
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (STAGES + s) * 8, CL);   // freed when ALL CTAs' MMAs have read the stage (the peers write parts of B)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b * 8, 1);
      mbar_init(tempty + b * 8, 128);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.bmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();      // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer: one continuous ring over all units =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = MT * p.box_rows * 128 + B_BYTES;
      for (int u = cluster_id; u < total_pairs; u += n_clusters) {
        Unit t;
        decode(u, t);
        const TcClass cls = p.cls[t.cls_id];
        for (int ti = 0; ti < cls.tap_count; ++ti) {
          const TcTap tap = p.taps[cls.tap_begin + ti];
          const void* amap = &p.amap[tap.map];
          for (int ch = 0; ch < tap.nchunk; ++ch) {
            mbar_wait(bar_base + (STAGES + stage) * 8, phase ^ 1u);
            const uint32_t full = bar_base + stage * 8;
            const uint32_t sa = base + stage * STAGE_BYTES;
            mbar_expect_tx(full, tx_bytes);
#pragma unroll
            for (int sub = 0; sub < MT; ++sub)
              tma_load_5d(sa + sub * A_BYTES, amap, full, (tap.c0 + ch) * 64, t.w0s[sub] + tap.dw, t.h0s[sub] + tap.dh, t.d0s[sub] + tap.dd,
                          t.n0s[sub]);
            tma_load_3d_mc(sa + MT * A_BYTES + crank * (B_BYTES / CL), CL == 2 ? &p.bmap_half : &p.bmap_quarter, full, tap.kofs + ch * 64,
                           t.nt * BLOCK_N + static_cast<int>(crank) * (BLOCK_N / CL), 0, CL_MASK);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < total_pairs; u += n_clusters, ++it) {
        const int buf = it % NBUF;
        const uint32_t use = static_cast<uint32_t>(it / NBUF);
        const int nkb = p.cls[0].nkb;
        mbar_wait(tempty + buf * 8, (use & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * MT * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_base + stage * 8, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t bdesc = umma_desc_sw128(sa + MT * A_BYTES, 16, 1024);
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
            const uint64_t adesc = umma_desc_sw128(sa + sub * A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_bf16(tacc + sub * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_mc(bar_base + (STAGES + stage) * 8, CL_MASK);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (nkb > 0) tc_commit(tfull + buf * 8);
        else mbar_arrive(tfull + buf * 8);               // bias-only class: nothing to wait for
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5), overlapped with the next unit's MMAs =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int u = cluster_id; u < total_pairs; u += n_clusters, ++it) {
      Unit t;
      decode(u, t);
      const TcClass cls = p.cls[t.cls_id];
      const int nkb = cls.nkb;
      const int buf = it % NBUF;
      const uint32_t use = static_cast<uint32_t>(it / NBUF);
      for (int cc = et; cc < BLOCK_N; cc += 128) {     // per-column epilogue vectors of this unit
        const int col = t.nt * BLOCK_N + cc;
        const bool in = col < p.cout;
        s_ep[cc] = (in && p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
        s_ep[BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.scale + col) : 1.f;
        s_ep[2 * BLOCK_N + cc] = (in && p.scale != nullptr) ? __ldg(p.shift + col) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tfull + buf * 8, use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * MT * BLOCK_N;
      if (!t.live[0]) {   // duplicate unit of an odd tail: nothing to store, only hand the accumulator back
        tc_fence_before();
        mbar_arrive(tempty + buf * 8);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        continue;
      }
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        if (!t.live[sub]) break;
        int r = row;
        const int iw = r % p.box[0]; r /= p.box[0];
        const int ih = r % p.box[1]; r /= p.box[1];
        const int id = r % p.box[2];
        const int in = r / p.box[2];
        const int ow = t.w0s[sub] + iw, oh = t.h0s[sub] + ih, od = t.d0s[sub] + id, on = t.n0s[sub] + in;
        const bool valid = row < p.box_rows && ow < p.ext[0] && oh < p.ext[1] && od < p.ext[2] && on < p.ext[3];
        const long long row_off = cls.out_ofs + ow * p.so[0] + oh * p.so[1] + od * p.so[2] + on * p.so[3];
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int col0 = t.nt * BLOCK_N + c * 32;
          if (col0 >= p.cout) break;
          uint32_t rr[32];
          if (nkb > 0) {
            tmem_ld_32x32(tacc + (static_cast<uint32_t>(q * 32) << 16) + sub * BLOCK_N + c * 32, rr);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0u;
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += s_ep[c * 32 + j];
          }
          if (want_stats) {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float m = valid ? v[j] : 0.f;
              s1[j] = m;
              s2[j] = m * m;
            }
            const float c1 = warp_transpose_sum32(s1, lane);
            const float c2 = warp_transpose_sum32(s2, lane);
            s_stats[(q * 2 + 0) * BLOCK_N + c * 32 + lane] = c1;
            s_stats[(q * 2 + 1) * BLOCK_N + c * 32 + lane] = c2;
          }
          if (p.scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_ep[BLOCK_N + c * 32 + j], s_ep[2 * BLOCK_N + c * 32 + j]);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (valid) {
            const bool second = p.out2 != nullptr && col0 >= p.seg_split;
            void* const outp = second ? p.out2 : p.out;
            const int colx = second ? col0 - p.seg_split : col0;
            const int accum = second ? p.accumulate2 : p.accumulate;
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(outp) + row_off + colx;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                if (col0 + g * 4 < p.cout) {
                  float4 f = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                  if (accum) {
                    const float4 e = *reinterpret_cast<const float4*>(o + g * 4);
                    f.x += e.x; f.y += e.y; f.z += e.z; f.w += e.w;
                  }
                  *reinterpret_cast<float4*>(o + g * 4) = f;
                }
              }
            } else {
              bf16* o = reinterpret_cast<bf16*>(outp) + row_off + colx;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (col0 + g * 8 < p.cout) {
                  float w8[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) w8[j] = v[g * 8 + j];
                  if (accum) {
                    float e[8];
                    Vec8<bf16>::load(o + g * 8, e);
#pragma unroll
                    for (int j = 0; j < 8; ++j) w8[j] += e[j];
                  }
                  Vec8<bf16>::store(o + g * 8, w8);
                }
              }
            }
          }
        }
        if (sub + 1 == MT || !t.live[sub + 1 < MT ? sub + 1 : sub]) {
          // last sub-tile of the unit has been read out of TMEM: hand the buffer back before the statistics tail
          tc_fence_before();
          mbar_arrive(tempty + buf * 8);
        }
        if (want_stats) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const long long srow = static_cast<long long>(t.cls_id) * p.m_tiles + t.mts[sub];
          for (int cc = et; cc < BLOCK_N; cc += 128) {
            const int col = t.nt * BLOCK_N + cc;
            if (col < p.cout) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                a += s_stats[(qq * 2 + 0) * BLOCK_N + cc];
                b += s_stats[(qq * 2 + 1) * BLOCK_N + cc];
              }
              p.stats[(srow * 2 + 0) * p.cout + col] = a;
              p.stats[(srow * 2 + 1) * p.cout + col] = b;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // s_stats / s_ep are reused by the next sub-tile / unit
        }
      }
      if (!want_stats) asm volatile("bar.sync 1, 128;" ::: "memory");   // s_ep is rewritten at the top of the next unit
    }
    tc_fence_before();
  }
  cluster_sync_all();      // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// bf16 view with channels innermost -> rank-5 map {C, d0, d1, d2, d3}, 128B swizzle, zero OOB fill
static int encode_view(CUtensorMap* m, const void* base, int C, const int dim[4], const long long stride[4],
                       const int box[4], char* err, size_t errlen) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2], (cuuint64_t)dim[3]};
  cuuint64_t gstr[4] = {(cuuint64_t)stride[0] * 2, (cuuint64_t)stride[1] * 2, (cuuint64_t)stride[2] * 2,
                        (cuuint64_t)stride[3] * 2};
  // the driver wants strides of size-1 dims to still be valid multiples of 16
  for (int i = 0; i < 4; ++i)
    if (gstr[i] == 0) gstr[i] = 16;
  cuuint32_t bdim[5] = {64u, (cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2], (cuuint32_t)box[3]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen,
             "cuTensorMapEncodeTiled(view) failed: %d (C=%d dims=%d,%d,%d,%d strides=%lld,%lld,%lld,%lld box=%d,%d,%d,%d "
             "base=%p)",
             (int)r, C, dim[0], dim[1], dim[2], dim[3], stride[0], stride[1], stride[2], stride[3], box[0], box[1],
             box[2], box[3], base);
    return 1;
  }
  return 0;
}

// [batch][rows][Ktot] bf16 (batch = 1 for convolutions) -> rank-3 map {K, rows, batch}, box {64, block_n, 1}
static int encode_b(CUtensorMap* m, const void* base, int Ktot, int rows, int block_n, int batch, long long batch_stride, char* err,
                    size_t errlen) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)Ktot, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t gstr[2] = {(cuuint64_t)Ktot * 2, (cuuint64_t)(batch > 1 ? batch_stride * 2 : (long long)Ktot * 2 * rows)};
  cuuint32_t bdim[3] = {64u, (cuuint32_t)block_n, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(B) failed: %d (K=%d rows=%d box_n=%d)", (int)r, Ktot, rows, block_n);
    return 1;
  }
  return 0;
}

// --- dimension merging + box search ---------------------------------------------------------------
struct Merged {
  int ext[4];
  long long so[4];
  std::vector<TcView> views;
  std::vector<TcClassH> classes;
};

static void merge_dims(const TcProblem& pb, Merged& m) {
  for (int i = 0; i < 4; ++i) {
    m.ext[i] = pb.ext[i];
    m.so[i] = pb.so[i];
  }
  m.views = pb.views;
  m.classes = pb.classes;
  if (pb.b_batch > 1) return;   // the batch stays dim 3: a tile's n coordinate selects its B matrix
  // try to fold dim i+1 into dim i (W<-H, then <-D, then <-N) while geometry allows
  int i = 0;
  int live = 4;
  while (i + 1 < live) {
    bool ok = m.so[i + 1] == m.so[i] * m.ext[i];
    for (const TcView& v : m.views)
      ok = ok && v.dim[i] == m.ext[i] && v.dim[i + 1] == m.ext[i + 1] && v.stride[i + 1] == v.stride[i] * v.dim[i];
    for (const TcClassH& c : m.classes)
      for (const TcTapH& tp : c.taps) ok = ok && tp.off[i] == 0 && tp.off[i + 1] == 0;
    if (ok && (long long)m.ext[i] * m.ext[i + 1] < (1ll << 31)) {
      m.ext[i] *= m.ext[i + 1];
      for (TcView& v : m.views) v.dim[i] *= v.dim[i + 1];
      for (int j = i + 1; j + 1 < 4; ++j) {
        m.ext[j] = m.ext[j + 1];
        m.so[j] = m.so[j + 1];
        for (TcView& v : m.views) {
          v.dim[j] = v.dim[j + 1];
          v.stride[j] = v.stride[j + 1];
        }
        for (TcClassH& c : m.classes)
          for (TcTapH& tp : c.taps) tp.off[j] = tp.off[j + 1];
      }
      m.ext[3] = 1;
      m.so[3] = 0;
      for (TcView& v : m.views) {
        v.dim[3] = 1;
        v.stride[3] = 0;
      }
      for (TcClassH& c : m.classes)
        for (TcTapH& tp : c.taps) tp.off[3] = 0;
      --live;
    } else {
      ++i;
    }
  }
}

static long long choose_box(const int ext[4], int box[4], int max_bn = 128) {
  long long best = -1;
  int bb[4] = {1, 1, 1, 1};
  for (int bw = 1; bw <= std::min(ext[0], 128); ++bw) {
    for (int bh = 1; bh <= std::min(ext[1], 128 / bw); ++bh) {
      for (int bd = 1; bd <= std::min(ext[2], 128 / (bw * bh)); ++bd) {
        int bn = std::min(std::min(ext[3], max_bn), 128 / (bw * bh * bd));
        if (bn < 1) continue;
        long long tiles = (long long)((ext[0] + bw - 1) / bw) * ((ext[1] + bh - 1) / bh) * ((ext[2] + bd - 1) / bd) *
                          ((ext[3] + bn - 1) / bn);
        bool better = best < 0 || tiles < best || (tiles == best && bw > bb[0]);
        if (better) {
          best = tiles;
          bb[0] = bw; bb[1] = bh; bb[2] = bd; bb[3] = bn;
        }
      }
    }
  }
  for (int i = 0; i < 4; ++i) box[i] = bb[i];
  return best;
}

int tc_plan_tiles(const TcProblem& pb) {
  Merged m;
  merge_dims(pb, m);
  int box[4];
  return (int)choose_box(m.ext, box, pb.b_batch > 1 ? 1 : 128);
}

template <int BLOCK_N, int STAGES, int MT>
static int launch_t(const TcConvParams& prm, int grid, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int SMEM = STAGES * (MT * 128 * 128 + BLOCK_N * 128) + (2 * STAGES + 2) * 8 + 16 + (4 * 2 + 3) * BLOCK_N * 4 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(conv_tc) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    attr_done = true;
  }
  cudaError_t e = launch_k(conv_tc_kernel<BLOCK_N, STAGES, MT>, dim3((unsigned)grid), dim3(TC_THREADS), SMEM, stream, 1, prm);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_tc launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

template <int SPLIT>
static int launch_split(const TcConvParams& prm, int grid, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int SMEM = 4 * (128 * 128 + 128 * 128) + (2 * 4 + 2) * 8 + 16 + (4 * 2 + 3) * 128 * 4 + (SPLIT - 1) * (4 / SPLIT) * 16384 + 128 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<128, 4, 1, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(conv_tc split) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    attr_done = true;
  }
  cudaError_t e = launch_k(conv_tc_kernel<128, 4, 1, SPLIT>, dim3((unsigned)grid * SPLIT), dim3(TC_THREADS), SMEM, stream, SPLIT, prm);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_tc split-K cluster launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// how many SPLIT-CTA clusters of the split-K kernel can be resident at once (GPC granularity can leave it below SMs / SPLIT):
// the in-launch BatchNorm needs EVERY CTA of the launch resident, its grid barrier would otherwise never complete
template <int SPLIT>
static int split_max_clusters() {
  static int cached = -1;
  if (cached >= 0) return cached;
  constexpr int SMEM = 4 * (128 * 128 + 128 * 128) + (2 * 4 + 2) * 8 + 16 + (4 * 2 + 3) * 128 * 4 + (SPLIT - 1) * (4 / SPLIT) * 16384 + 128 + 1024;
  cudaFuncSetAttribute(conv_tc_kernel<128, 4, 1, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(SPLIT * 32);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SPLIT;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<128, 4, 1, SPLIT>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cached = n;
  return cached;
}

template <int BLOCK_N, int STAGES, int MT, int HALO = 0>
static int launch_persist(const TcConvParams& prm, int units, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int RING = HALO ? STAGES * MT * TC_HALO_MAX_ROWS * 128 + HALO * BLOCK_N * 128 : STAGES * (MT * 128 * 128 + BLOCK_N * 128);
  constexpr int NBAR = (HALO ? 2 * STAGES + 2 * HALO : 2 * STAGES) + 4 + 2 * TC_QD;
  constexpr int SMEM = RING + NBAR * 8 + 32 + (4 * 2 + 3) * BLOCK_N * 4 + 1024;
  static_assert(SMEM <= 232448, "persistent conv kernel exceeds the 227 KB of shared memory a CTA can opt in to");
  static bool attr_done = false;
  static int sms = 0;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel<BLOCK_N, STAGES, MT, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(conv_tc_persist) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    attr_done = true;
  }
  (void)units;
  cudaError_t e = launch_k(conv_tc_persist_kernel<BLOCK_N, STAGES, MT, HALO>, dim3((unsigned)(prm.n_units < sms ? prm.n_units : sms)), dim3(TC_THREADS),
                           SMEM, stream, 1, prm);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_tc_persist launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

template <int NA, int NW>
static int launch_swap(const TcConvParams& prm, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int SMEM = NA * TC_SWAP_MAX_ROWS * 128 + NW * 128 * 128 + 2 * 32 * 128 * 2 + 2 * 256 * 8 + (2 * NA + 2 * NW + 4 + 2 * TC_QD) * 8 + 32 + 1024;
  static_assert(SMEM <= 232448, "swapped-operand conv kernel exceeds the 227 KB of shared memory a CTA can opt in to");
  static bool attr_done = false;
  static int sms = 0;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_swap_kernel<NA, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(conv_tc_swap) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    attr_done = true;
  }
  const int ctas = prm.n_units < sms ? prm.n_units : sms;
  cudaError_t e = launch_k(conv_tc_swap_kernel<NA, NW>, dim3((unsigned)ctas), dim3(TC_THREADS), SMEM, stream, 1, prm);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_tc_swap launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

template <int BLOCK_N, int STAGES, int MT, int CL>
static int launch_persist_mc(const TcConvParams& prm, int m_groups, cudaStream_t stream, char* err, size_t errlen) {
  constexpr int SMEM = STAGES * (MT * 128 * 128 + BLOCK_N * 128) + (2 * STAGES + 4) * 8 + 16 + (4 * 2 + 3) * BLOCK_N * 4 + 1024;
  static bool attr_done = false;
  static int sms = 0;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_mc_kernel<BLOCK_N, STAGES, MT, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(conv_tc_persist_mc) failed: %s", cudaGetErrorString(e));
      return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 2) sms = 148;
    attr_done = true;
  }
  const int pairs = prm.n_tiles * ((m_groups + CL - 1) / CL);   // groups of CL row groups with the same n-tile
  int clusters = sms / CL;
  if (clusters > pairs) clusters = pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)clusters * CL);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_persist_mc_kernel<BLOCK_N, STAGES, MT, CL>, prm, pairs);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_tc_persist_mc cluster launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

static unsigned long long* g_conv_dbg = nullptr;
void tc_set_debug_buffer(void* buf) { g_conv_dbg = reinterpret_cast<unsigned long long*>(buf); }
static std::atomic<long long> g_halo_launches{0}, g_swap_launches{0};
long long tc_halo_launches() { return g_halo_launches.load(); }
long long tc_swap_launches() { return g_swap_launches.load(); }

// opt-in while it is being measured: SAP3D_CONV_MULTICAST=2 or 4 (cluster size); unset / 0 = off
static int multicast_cluster() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_CONV_MULTICAST");
    v = (e != nullptr && (e[0] == '2' || e[0] == '4')) ? e[0] - '0' : 0;
  }
  return v;
}

// ticket counters of the persistent kernels: a pool of {next, done} pairs, zero-initialised once; every launch takes the next
// slot and its last CTA zeroes it again, so a slot is clean whenever it comes round (4096 persistent launches later)
constexpr int TC_SCHED_SLOTS = 4096;
static unsigned int* sched_slot(cudaStream_t stream, char* err, size_t errlen) {
  static unsigned int* pool[64] = {nullptr};
  static std::atomic<unsigned int> seq{0};   // (host threads may launch concurrently; the pool itself is set up under `mu`)
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) {
    snprintf(err, errlen, "tc_launch: device ordinal %d out of range", dev);
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (pool[dev] == nullptr) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      snprintf(err, errlen, "tc_launch: the first persistent convolution of a process cannot be launched inside a stream capture "
                            "(its ticket counters are allocated then): run the step once eagerly first");
      return nullptr;
    }
    unsigned int* ptr = nullptr;
    if (cudaMalloc(&ptr, TC_SCHED_SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(ptr, 0, TC_SCHED_SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      snprintf(err, errlen, "tc_launch: allocating the ticket counters failed: %s", cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    pool[dev] = ptr;
  }
  return pool[dev] + 2 * (seq.fetch_add(1u) % TC_SCHED_SLOTS);
}
static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
  }
  return sms;
}
// work units of a persistent launch (see conv_tc_persist_kernel): balanced problems = groups of mt tiles, then ~one single tile
// per SM; others = (class, group, column tile) triples
static int plan_units(TcConvParams& prm, int mt, long long grid, cudaStream_t stream, char* err, size_t errlen) {
  if (prm.balanced) {
    const int T = prm.m_tiles;
    int singles = 0;
    if (mt > 1) {
      singles = std::min(T, sm_count());
      while ((T - singles) % mt != 0) ++singles;   // terminates: T - T = 0
    }
    prm.n_pair_units = (T - singles) / mt;
    prm.n_units = prm.n_pair_units + singles;
  } else {
    prm.n_pair_units = 0;
    prm.n_units = (int)grid;
  }
  prm.sched = sched_slot(stream, err, errlen);
  return prm.sched == nullptr ? 1 : 0;
}

// SAP3D_CONV_HALO: 0 = never use the halo-tile kernels, 1 (default) = two-sub-tile units only, 2 = also single-sub-tile units
static int halo_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_CONV_HALO");
    v = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return v;
}
// SAP3D_CONV_SWAP=0: halo-tile launches with 64 < cout <= 128 keep the positions on the M side (no swapped-operand kernel)
static bool swap_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_CONV_SWAP");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
// SAP3D_CONV_NARROW=1: trade 256-column tiles for twice as many 128-column units where the 256-column units fill the persistent
// rounds badly.  Off by default: measured on the level-2 decoder layers (196 units on 148 SMs) it LOSES -- 13 launches went from
// ~62 to ~76 us; the 128-column MMA shape and the extra weight traffic cost more than the third, better filled round gains.
static bool narrow_tiles_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_CONV_NARROW");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}
// SAP3D_CONV_BALANCED=0: persistent kernels take units round-robin even where contiguous equal ranges are possible
static bool balanced_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_CONV_BALANCED");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Halo-tile plan (see conv_tc_persist_kernel): the taps of a one-class problem grouped into triples that differ only by
// consecutive offsets along axis X (1 = H, 2 = D), and an output box of exactly 128 rows whose outermost non-unit axis is X
// with `inner` = rows per X step in {8, 16}.  Accepted only if the box needs exactly as many tiles as the default one.
struct HaloPlan {
  int axis = 0, inner = 0;
  int box[4] = {1, 1, 1, 1};
  long long tiles = 0;
  std::vector<TcTapH> taps;   // reordered: triples, ascending along the axis
};
static bool plan_halo(const Merged& m, long long default_tiles, HaloPlan& best) {
  if (m.classes.size() != 1) return false;
  const std::vector<TcTapH>& taps = m.classes[0].taps;
  if (taps.empty() || taps.size() % 3 != 0) return false;
  bool found = false;
  for (int X = 1; X <= 2; ++X) {
    // group the taps
    std::vector<TcTapH> ordered;
    std::vector<char> used(taps.size(), 0);
    bool ok = true;
    for (size_t i = 0; i < taps.size() && ok; ++i) {
      if (used[i]) continue;
      std::vector<size_t> grp;
      for (size_t j = i; j < taps.size(); ++j) {
        if (used[j]) continue;
        const TcTapH &a = taps[i], &b = taps[j];
        bool same = a.view == b.view && a.c_begin == b.c_begin && a.nch == b.nch;
        for (int d = 0; d < 4; ++d)
          if (d != X) same = same && a.off[d] == b.off[d];
        if (same) grp.push_back(j);
      }
      if (grp.size() != 3) { ok = false; break; }
      std::sort(grp.begin(), grp.end(), [&](size_t a, size_t b) { return taps[a].off[X] < taps[b].off[X]; });
      if (taps[grp[1]].off[X] != taps[grp[0]].off[X] + 1 || taps[grp[2]].off[X] != taps[grp[0]].off[X] + 2) { ok = false; break; }
      for (size_t g : grp) {
        used[g] = 1;
        ordered.push_back(taps[g]);
      }
    }
    if (!ok) continue;
    // boxes of exactly 128 rows, unit extent above X
    for (int bw = 1; bw <= std::min(m.ext[0], 128); ++bw) {
      for (int bh = 1; bh <= std::min(m.ext[1], 128 / bw); ++bh) {
        if (X == 1) {
          if (bw * bh != 128) continue;
        }
        const int bd = X == 2 ? 128 / (bw * bh) : 1;
        if (bw * bh * bd != 128 || bd > m.ext[2]) continue;
        const int bx = X == 2 ? bd : bh;
        const int inner = X == 2 ? bw * bh : bw;
        if (bx < 2 || inner % 8 != 0 || 128 + 2 * inner > TC_HALO_MAX_ROWS) continue;
        const long long tiles = (long long)((m.ext[0] + bw - 1) / bw) * ((m.ext[1] + bh - 1) / bh) * ((m.ext[2] + bd - 1) / bd) * m.ext[3];
        if (tiles != default_tiles) continue;   // sap3d_conv_stats_rows() is planned with the default box: keep the tile count
        const bool better = !found || tiles < best.tiles || (tiles == best.tiles && (inner < best.inner || (inner == best.inner && bw > best.box[0])));
        if (better) {
          found = true;
          best.axis = X;
          best.inner = inner;
          best.box[0] = bw; best.box[1] = bh; best.box[2] = bd; best.box[3] = 1;
          best.tiles = tiles;
          best.taps = ordered;
        }
      }
    }
  }
  return found;
}

int tc_launch(const TcProblem& pb, cudaStream_t stream, char* err, size_t errlen) {
  Merged m;
  merge_dims(pb, m);
  if ((int)m.views.size() > TC_MAX_MAPS) {
    snprintf(err, errlen, "tc_launch: too many views (%d)", (int)m.views.size());
    return 1;
  }
  if ((int)m.classes.size() > TC_MAX_CLS) {
    snprintf(err, errlen, "tc_launch: too many classes (%d)", (int)m.classes.size());
    return 1;
  }
  if (pb.cout % 8 != 0 || pb.Ktot % 64 != 0) {
    snprintf(err, errlen, "tc_launch: cout %% 8 / Ktot %% 64 violated (cout=%d Ktot=%d)", pb.cout, pb.Ktot);
    return 1;
  }
  static thread_local TcConvParams prm;  // ~4 KB, filled per call
  memset(&prm, 0, sizeof(prm));
  int box[4];
  long long m_tiles = choose_box(m.ext, box, pb.b_batch > 1 ? 1 : 128);
  if (pb.b_batch > 1 && (pb.b_batch != m.ext[3] || m.classes.size() != 1)) {
    snprintf(err, errlen, "tc_launch: batched B needs ext[3] == b_batch and one class");
    return 1;
  }
  int block_n = pb.force_block_n;
  if (block_n == 0) {
    if (pb.cout <= 64) block_n = 64;
    else if (pb.cout % 256 == 0 && m_tiles * (long long)m.classes.size() * (pb.cout / 256) >= 148) {
      // 256-column tiles unless they fill the persistent rounds so badly that twice as many 128-column units win despite their
      // less efficient MMA shape (~0.85x): e.g. 196 units on 148 SMs = two rounds at 66 %, 392 half-size units = three at 88 %
      const long long u256 = m_tiles * (long long)m.classes.size() * (pb.cout / 256), u128 = 2 * u256, sms = sm_count();
      const double f256 = (double)u256 / (double)(((u256 + sms - 1) / sms) * sms), f128 = (double)u128 / (double)(((u128 + sms - 1) / sms) * sms);
      block_n = (narrow_tiles_enabled() && f128 * 0.85 > f256) ? 128 : 256;
    } else block_n = 128;
  }
  if (block_n > 64 && pb.rowsB % block_n != 0 && pb.rowsB < block_n) block_n = 64;
  // halo-tile kernel: persistent one-class problems with one column tile whose taps form triples along H or D
  bool use_halo = false, use_swap = false;
  {
    const long long n_tiles0 = (pb.cout + block_n - 1) / block_n;
    const long long ctas0 = (long long)m.classes.size() * m_tiles * n_tiles0;
    const int mt0 = pb.b_batch > 1 ? 1 : (pb.force_mt ? pb.force_mt : (ctas0 >= 4 * 148 ? 2 : 1));
    const long long units0 = (long long)m.classes.size() * ((m_tiles + mt0 - 1) / mt0) * n_tiles0;
    const int hm = halo_mode();
    if (hm > 0 && units0 > 148 && pb.force_split >= 0 && multicast_cluster() == 0 && pb.b_batch <= 1 && n_tiles0 == 1 &&
        (block_n == 128 || block_n == 256) && (mt0 == 2 || hm == 2)) {
      HaloPlan hp;
      if (plan_halo(m, m_tiles, hp)) {
        int hbox[4] = {hp.box[0], hp.box[1], hp.box[2], hp.box[3]};
        hbox[hp.axis] += 2;
        // swapped operands (positions on the N side, 256 per instruction): halo along D, tiles paired along H
        use_swap = swap_enabled() && hp.axis == 2 && block_n == 128 && pb.cout > 64 && pb.cout <= 128 && !pb.out_f32 && pb.out2 == nullptr &&
                   mt0 == 2 && (int)m.views.size() <= TC_SWAP_SINGLE_MAP && ((m.ext[1] + hp.box[1] - 1) / hp.box[1]) % 2 == 0 &&
                   2 * (128 + 2 * hp.inner) <= TC_SWAP_MAX_ROWS && hp.tiles >= 2 * 148;
        bool enc_ok = true;
        for (size_t v = 0; v < m.views.size() && enc_ok && !pb.query_fuse_bn; ++v) {
          if (use_swap) {
            int pbox[4] = {hbox[0], 2 * hbox[1], hbox[2], hbox[3]};
            enc_ok = encode_view(&prm.amap[v], m.views[v].base, m.views[v].C, m.views[v].dim, m.views[v].stride, pbox, err, errlen) == 0 &&
                     encode_view(&prm.amap[TC_SWAP_SINGLE_MAP + v], m.views[v].base, m.views[v].C, m.views[v].dim, m.views[v].stride, hbox, err,
                                 errlen) == 0;
          } else {
            enc_ok = encode_view(&prm.amap[v], m.views[v].base, m.views[v].C, m.views[v].dim, m.views[v].stride, hbox, err, errlen) == 0;
          }
        }
        if (!enc_ok) use_swap = false;
        if (enc_ok) {
          use_halo = true;
          for (int i = 0; i < 4; ++i) box[i] = hp.box[i];
          m_tiles = hp.tiles;
          m.classes[0].taps = hp.taps;
          prm.halo_inner = hp.inner;
          prm.halo_rows = 128 + 2 * hp.inner;
        }
      }
    }
  }
  if (!use_halo && !pb.query_fuse_bn)
    for (size_t v = 0; v < m.views.size(); ++v)
      if (encode_view(&prm.amap[v], m.views[v].base, m.views[v].C, m.views[v].dim, m.views[v].stride, box, err, errlen))
        return 1;
  prm.b_batched = pb.b_batch > 1 ? 1 : 0;
  if (!pb.query_fuse_bn) {
    if (encode_b(&prm.bmap, pb.B, pb.Ktot, pb.rowsB, block_n, pb.b_batch, pb.b_batch_stride, err, errlen)) return 1;
    if (encode_b(&prm.bmap_half, pb.B, pb.Ktot, pb.rowsB, block_n / 2, pb.b_batch, pb.b_batch_stride, err, errlen)) return 1;
    if (encode_b(&prm.bmap_quarter, pb.B, pb.Ktot, pb.rowsB, block_n / 4, pb.b_batch, pb.b_batch_stride, err, errlen)) return 1;
  }
  int ntap = 0;
  for (size_t c = 0; c < m.classes.size(); ++c) {
    TcClass& dc = prm.cls[c];
    dc.tap_begin = (int16_t)ntap;
    dc.tap_count = (int16_t)m.classes[c].taps.size();
    dc.out_ofs = m.classes[c].out_ofs;
    int nkb = 0;
    for (const TcTapH& tp : m.classes[c].taps) {
      if (ntap >= TC_MAX_TAPS) {
        snprintf(err, errlen, "tc_launch: too many taps");
        return 1;
      }
      if (tp.nch % 64 != 0 || tp.c_begin % 64 != 0) {
        snprintf(err, errlen, "tc_launch: tap channels must be multiples of 64 (nch=%d c_begin=%d)", tp.nch, tp.c_begin);
        return 1;
      }
      TcTap& dt = prm.taps[ntap++];
      dt.map = (int8_t)tp.view;
      dt.dw = (int8_t)tp.off[0];
      dt.dh = (int8_t)tp.off[1];
      dt.dd = (int8_t)tp.off[2];
      if (tp.off[3] != 0) {
        snprintf(err, errlen, "tc_launch: batch offset unsupported");
        return 1;
      }
      dt.nchunk = (int16_t)(tp.nch / 64);
      dt.c0 = (int16_t)(tp.c_begin / 64);
      dt.kofs = tp.kofs;
      nkb += tp.nch / 64;
    }
    dc.nkb = nkb;
  }
  prm.ncls = (int)m.classes.size();
  prm.m_tiles = (int)m_tiles;
  prm.n_tiles = (pb.cout + block_n - 1) / block_n;
  for (int i = 0; i < 4; ++i) {
    prm.box[i] = box[i];
    prm.ext[i] = m.ext[i];
    prm.so[i] = m.so[i];
    prm.tiles[i] = (m.ext[i] + box[i] - 1) / box[i];
  }
  prm.cout = pb.cout;
  prm.box_rows = box[0] * box[1] * box[2] * box[3];
  prm.out = pb.out;
  prm.bias = pb.bias;
  prm.stats = pb.stats;
  prm.scale = pb.scale;
  prm.shift = pb.shift;
  prm.relu = pb.relu;
  prm.accumulate = pb.accumulate;
  prm.out2 = pb.out2;
  prm.seg_split = pb.seg_split;
  prm.accumulate2 = pb.accumulate2;
  prm.out_f32 = pb.out_f32;
  prm.dbg = g_conv_dbg;
  // two M sub-tiles per CTA when the problem still fills the GPU several times over
  const long long ctas1 = (long long)prm.ncls * prm.m_tiles * prm.n_tiles;
  int mt = (pb.force_mt ? pb.force_mt : (ctas1 >= 4 * 148 ? 2 : 1));
  if (pb.b_batch > 1) mt = 1;   // the sub-tiles of a unit share one B tile, so they must belong to one sample
  const long long m_groups = (prm.m_tiles + mt - 1) / mt;
  long long grid = (long long)prm.ncls * m_groups * prm.n_tiles;
  if (grid <= 0 || grid > 0x7fffffffll) {
    snprintf(err, errlen, "tc_launch: bad grid %lld", grid);
    return 1;
  }
  // few output tiles and a long K loop (stage-2/3 backbone layers): split K over a cluster of 2 or 4 CTAs
  int split_sel = 1;
  if (block_n == 128 && mt == 1 && pb.force_split >= 0) {
    int min_nkb = 1 << 30;
    for (int c = 0; c < prm.ncls; ++c) min_nkb = std::min(min_nkb, prm.cls[c].nkb);
    int split = pb.force_split;
    if (split == 0) {
      if (grid <= 37 && min_nkb >= 8) split = 4;
      else if (grid <= 74 && min_nkb >= 8) split = 2;
    }
    if (split == 4 && min_nkb >= 4) split_sel = 4;
    else if (split == 2 && min_nkb >= 2) split_sel = 2;
  }
  // BatchNorm finished inside the launch (TcFuseBN): the non-persistent 128-column kernel with every CTA resident at once
  if (pb.fuse_bn != nullptr || pb.query_fuse_bn) {
    const bool fuse_ok = block_n == 128 && mt == 1 && prm.ncls == 1 && !pb.out_f32 && pb.out2 == nullptr && pb.b_batch <= 1 &&
                         pb.scale == nullptr && !pb.relu && !pb.accumulate && grid <= 148 &&
                         (split_sel == 1 ? grid <= sm_count() : grid <= (split_sel == 4 ? split_max_clusters<4>() : split_max_clusters<2>())) &&
                         pb.force_split >= 0 && (pb.query_fuse_bn || pb.stats != nullptr);
    if (pb.query_fuse_bn) return fuse_ok ? 0 : 2;
    if (!fuse_ok) {
      snprintf(err, errlen, "tc_launch: this problem cannot finish its BatchNorm inside the launch (ask sap3d_conv_fwd_bn_supported first)");
      return 1;
    }
    prm.fuse = *pb.fuse_bn;
    prm.gbar = sched_slot(stream, err, errlen);
    if (prm.gbar == nullptr) return 1;
  }
  if (split_sel == 4) return launch_split<4>(prm, (int)grid, stream, err, errlen);
  if (split_sel == 2) return launch_split<2>(prm, (int)grid, stream, err, errlen);
  // more work units than SMs (decoder layers): persistent CTAs with double-buffered accumulators
  if (grid > 148 && pb.force_split >= 0 && prm.ncls == 1 && block_n >= 128 && pb.b_batch <= 1 && multicast_cluster() != 0) {
    const int mg = (int)m_groups;
    if (multicast_cluster() == 2) {
      if (block_n == 128) return mt == 2 ? launch_persist_mc<128, 4, 2, 2>(prm, mg, stream, err, errlen) : launch_persist_mc<128, 4, 1, 2>(prm, mg, stream, err, errlen);
      return mt == 2 ? launch_persist_mc<256, 3, 2, 2>(prm, mg, stream, err, errlen) : launch_persist_mc<256, 4, 1, 2>(prm, mg, stream, err, errlen);
    }
    if (block_n == 128) return mt == 2 ? launch_persist_mc<128, 4, 2, 4>(prm, mg, stream, err, errlen) : launch_persist_mc<128, 4, 1, 4>(prm, mg, stream, err, errlen);
    return mt == 2 ? launch_persist_mc<256, 3, 2, 4>(prm, mg, stream, err, errlen) : launch_persist_mc<256, 4, 1, 4>(prm, mg, stream, err, errlen);
  }
  prm.balanced = (prm.ncls == 1 && prm.n_tiles == 1 && pb.b_batch <= 1 && balanced_enabled()) ? 1 : 0;
  {   // SAP3D_CONV_TRACE=1: one line per launch on stderr (developer aid: which kernel form every layer takes)
    static int trace = -1;
    if (trace < 0) {
      const char* e = getenv("SAP3D_CONV_TRACE");
      trace = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (trace) {
      int ntaps = 0, nkb = 0;
      for (int c = 0; c < prm.ncls; ++c) { ntaps += prm.cls[c].tap_count; nkb += prm.cls[c].nkb; }
      fprintf(stderr, "[conv_trace] ext=%dx%dx%dx%d cout=%d ncls=%d taps=%d nkb=%d views=%d box=%d,%d,%d,%d m_tiles=%d block_n=%d mt=%d grid=%lld split=%d %s\n",
              prm.ext[0], prm.ext[1], prm.ext[2], prm.ext[3], prm.cout, prm.ncls, ntaps, nkb, (int)m.views.size(), prm.box[0], prm.box[1],
              prm.box[2], prm.box[3], prm.m_tiles, block_n, mt, grid, split_sel,
              use_swap ? "swap" : (use_halo ? "halo" : (split_sel > 1 ? "splitk" : (grid > 148 && pb.force_split >= 0 ? "persist" : "plain"))));
    }
  }
  if (use_halo) {
    if (grid <= 148 || prm.box_rows != 128) {
      snprintf(err, errlen, "tc_launch: halo plan inconsistent (grid %lld, box rows %d)", grid, prm.box_rows);
      return 1;
    }
    if (plan_units(prm, use_swap ? 2 : mt, grid, stream, err, errlen)) return 1;
    ++g_halo_launches;
    if (use_swap) {
      ++g_swap_launches;
      return launch_swap<3, 5>(prm, stream, err, errlen);
    }
    if (block_n == 128) return mt == 2 ? launch_persist<128, 3, 2, 6>(prm, (int)grid, stream, err, errlen) : launch_persist<128, 3, 1, 6>(prm, (int)grid, stream, err, errlen);
    return mt == 2 ? launch_persist<256, 2, 2, 3>(prm, (int)grid, stream, err, errlen) : launch_persist<256, 3, 1, 4>(prm, (int)grid, stream, err, errlen);
  }
  if (grid > 148 && pb.force_split >= 0) {
    if (plan_units(prm, mt, grid, stream, err, errlen)) return 1;
    switch (block_n) {
      case 64: return mt == 2 ? launch_persist<64, 4, 2>(prm, (int)grid, stream, err, errlen) : launch_persist<64, 6, 1>(prm, (int)grid, stream, err, errlen);
      case 128: return mt == 2 ? launch_persist<128, 4, 2>(prm, (int)grid, stream, err, errlen) : launch_persist<128, 4, 1>(prm, (int)grid, stream, err, errlen);
      case 256: return mt == 2 ? launch_persist<256, 3, 2>(prm, (int)grid, stream, err, errlen) : launch_persist<256, 4, 1>(prm, (int)grid, stream, err, errlen);
    }
  }
  switch (block_n) {
    case 64: return mt == 2 ? launch_t<64, 4, 2>(prm, (int)grid, stream, err, errlen) : launch_t<64, 6, 1>(prm, (int)grid, stream, err, errlen);
    case 128: return mt == 2 ? launch_t<128, 4, 2>(prm, (int)grid, stream, err, errlen) : launch_t<128, 4, 1>(prm, (int)grid, stream, err, errlen);
    case 256:  // MT = 2 uses all 512 TMEM columns and 3 x 64 KB stages
      return mt == 2 ? launch_t<256, 3, 2>(prm, (int)grid, stream, err, errlen) : launch_t<256, 4, 1>(prm, (int)grid, stream, err, errlen);
  }
  snprintf(err, errlen, "tc_launch: unsupported block_n %d", block_n);
  return 1;
}

}  // namespace sap3d
