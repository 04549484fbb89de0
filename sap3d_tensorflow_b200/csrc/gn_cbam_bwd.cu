// Backward passes of the GroupNorm + CBAM model variant (gn/p3d_gn.py:24-46,175-177; utils/network.py:65-87,198-274),
// i.e. what tf.gradients derives for the transposes / tf.nn.moments / reduce_mean / reduce_max / dense / 7x7x7 conv /
// sigmoid chain of the reference.  All bandwidth-bound: 128-bit NDHWC row accesses, warp-shuffle reductions,
// deterministic two-stage sums (no float atomics).
//
//   GroupNorm (statistics over one sample and one group of C/G channels, M = S*C/G elements):
//     xhat = (x - mean[n,g]) * rstd[n,g],  z = gamma[c]*xhat + beta[c],  g = dL/dz
//     dgamma[c] = sum_{n,s} g*xhat,  dbeta[c] = sum_{n,s} g
//     dx = gamma*rstd*g - rstd*P[n,g] - xhat*rstd*Q[n,g],  P = sum_group(gamma*g)/M,  Q = sum_group(gamma*g*xhat)/M
//   CBAM on r:  cs = sigmoid(mlp(avg_s r) + mlp(max_s r)),  u = r*cs,  att = sigmoid(conv7([mean_c u, max_c u])),  out = u*att
//     tf.reduce_max gradients are split evenly among ties (TF semantics).
#include <string.h>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

// ==================================================================================================
// GroupNorm apply backward:  y = relu_out?( relu1?(GN1(a)) + relu2?(GN2(b) | b) )
// ==================================================================================================
struct GnBwdArgs {
  const void* dy; const void* a; const void* b;
  const float *s1, *t1, *mean1, *rstd1;  // [N][C], [N][C], [N][G], [N][G]
  const float *s2, *t2, *mean2, *rstd2;  // NULL: b is a plain tensor
  long long S;
  int N, C, G, cpg;
  int relu1, relu2, relu_out;
  float* partial; int rows;              // [N][rows][4][C]
  const float* coef;                     // [N][4][C]
  void* da; void* db; int acc_a, acc_b;
};

template <typename T>
SAP3D_DEVINL void gn_bwd_common(const GnBwdArgs& p, long long e, int n, int c, float (&g1)[8], float (&g2)[8], float (&xh1)[8],
                                float (&xh2)[8]) {
  const T* dy = reinterpret_cast<const T*>(p.dy);
  const T* a = reinterpret_cast<const T*>(p.a);
  const T* b = reinterpret_cast<const T*>(p.b);
  float d[8], av[8], z1[8], z2[8];
  Vec8<T>::load(dy + e, d);
  Vec8<T>::load(a + e, av);
  const long long si = (long long)n * p.C + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = n * p.G + (c + j) / p.cpg;
    z1[j] = fmaf(av[j], p.s1[si + j], p.t1[si + j]);
    xh1[j] = (av[j] - p.mean1[gi]) * p.rstd1[gi];
  }
  if (b) {
    float bv[8];
    Vec8<T>::load(b + e, bv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (p.s2) {
        const int gi = n * p.G + (c + j) / p.cpg;
        z2[j] = fmaf(bv[j], p.s2[si + j], p.t2[si + j]);
        xh2[j] = (bv[j] - p.mean2[gi]) * p.rstd2[gi];
      } else {
        z2[j] = bv[j];
        xh2[j] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { z2[j] = 0.f; xh2[j] = 0.f; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float r1 = p.relu1 ? fmaxf(z1[j], 0.f) : z1[j];
    const float r2 = p.relu2 ? fmaxf(z2[j], 0.f) : z2[j];
    float u = d[j];
    if (p.relu_out && !(r1 + r2 > 0.f)) u = 0.f;
    g1[j] = (p.relu1 && !(z1[j] > 0.f)) ? 0.f : u;
    g2[j] = (p.relu2 && !(z2[j] > 0.f)) ? 0.f : u;
  }
}

// fold the 4 position lanes of a warp, then the 8 warps of the block, and write [4][64] partial sums
SAP3D_DEVINL void block_fold_4x64(float (&acc)[4][8], float (*red)[4][64], float* dst, int C, int cbase) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[i][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[i][j] = v;
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][i][lane * 8 + j] = acc[i][j];
  }
  __syncthreads();
  const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) v += red[w][i][ch];
  if (cbase + ch < C) dst[(long long)i * C + cbase + ch] = v;
}

// grid = (rows, C/64, N); 256 threads = 8 channel-vector lanes x 32 position lanes
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const GnBwdArgs p) {
  __shared__ float red[8][4][64];
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.y * 64 + cv * 8;
  const int n = blockIdx.z;
  const long long per = (p.S + p.rows - 1) / p.rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.S ? pbeg + per : p.S;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (c < p.C) {
    for (long long pos = pbeg + pl; pos < pend; pos += 32) {
      float g1[8], g2[8], xh1[8], xh2[8];
      gn_bwd_common<T>(p, ((long long)n * p.S + pos) * p.C + c, n, c, g1, g2, xh1, xh2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += g1[j];
        acc[1][j] += g1[j] * xh1[j];
        acc[2][j] += g2[j];
        acc[3][j] += g2[j] * xh2[j];
      }
    }
  }
  block_fold_4x64(acc, red, p.partial + ((long long)n * p.rows + blockIdx.x) * 4 * p.C, p.C, blockIdx.y * 64);
}

// one block per (group, sample): rows -> per-channel sums (kept in vsum [N][4][C]) -> group sums weighted by gamma ->
// coef [N][4][C].  nsum = 2 (one norm) or 4 (two norms).  dgamma / dbeta are folded over the samples by gn_bwd_param_kernel
// in a fixed order (deterministic).
__global__ void __launch_bounds__(128) gn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int N, int C, int G, double M,
                                                               const float* __restrict__ gamma1, const float* __restrict__ rstd1,
                                                               const float* __restrict__ gamma2, const float* __restrict__ rstd2,
                                                               int nsum, float* __restrict__ coef, float* __restrict__ vsum) {
  __shared__ float v[4][32];
  __shared__ float gs[4];
  const int g = blockIdx.x, n = blockIdx.y, cpg = C / G, t = threadIdx.x;
  const bool mine = t < nsum * cpg;
  const int i = mine ? t / cpg : 0, cl = mine ? t % cpg : 0;
  const int c = g * cpg + cl;
  if (mine) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += partial[(((long long)n * rows + r) * 4 + i) * C + c];
    v[i][cl] = s;
    vsum[((long long)n * 4 + i) * C + c] = s;
  }
  __syncthreads();
  if (t < nsum) {
    const float* gm = t < 2 ? gamma1 : gamma2;
    double s = 0.0;
    for (int k = 0; k < cpg; ++k) s += (double)gm[g * cpg + k] * (double)v[t][k];
    const float rs = (t < 2 ? rstd1 : rstd2)[n * G + g];
    gs[t] = (float)(s / M) * rs;
  }
  __syncthreads();
  if (mine) coef[((long long)n * 4 + i) * C + c] = gs[i];
}
__global__ void __launch_bounds__(256) gn_bwd_param_kernel(const float* __restrict__ vsum, int N, int C, int nsum, float* dgamma1, float* dbeta1,
                                                            float* dgamma2, float* dbeta2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nsum * C) return;
  const int i = idx / C, c = idx % C;
  float* dst = i == 0 ? dbeta1 : (i == 1 ? dgamma1 : (i == 2 ? dbeta2 : dgamma2));
  if (!dst) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += vsum[((long long)n * 4 + i) * C + c];
  dst[c] += s;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const GnBwdArgs p) {
  const long long nvec = (long long)p.N * p.S * p.C / 8;
  T* da = reinterpret_cast<T*>(p.da);
  T* db = reinterpret_cast<T*>(p.db);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long e = v * 8;
    const long long pos = e / p.C;
    const int c = (int)(e - pos * p.C);
    const int n = (int)(pos / p.S);
    float g1[8], g2[8], xh1[8], xh2[8];
    gn_bwd_common<T>(p, e, n, c, g1, g2, xh1, xh2);
    const float* cf = p.coef + (long long)n * 4 * p.C + c;
    const long long si = (long long)n * p.C + c;
    if (da) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p.s1[si + j] * g1[j] - cf[j] - xh1[j] * cf[p.C + j];
      if (p.acc_a) {
        float old[8];
        Vec8<T>::load(da + e, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      Vec8<T>::store(da + e, o);
    }
    if (db) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p.s2 ? p.s2[si + j] * g2[j] - cf[2 * p.C + j] - xh2[j] * cf[3 * p.C + j] : g2[j];
      if (p.acc_b) {
        float old[8];
        Vec8<T>::load(db + e, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      Vec8<T>::store(db + e, o);
    }
  }
}

// ==================================================================================================
// CBAM block tail backward:  y = relu( GN(c3) + cbam(r) )
// ==================================================================================================
struct TailArgs {
  const void* dy; const void* y; const void* c3; const void* r;
  const float *s3, *mean3, *rstd3;   // [N][C], [N][G], [N][G]
  const float *cscale, *sp, *att;    // [N][C], [N][S][2], [N][S]
  const float* chmax;                // [N][.] with row stride `save_ld`
  long long save_ld;
  const float* dsp;                  // [N][S][2]: d mean-map / C, d max-map / ties
  long long S;
  int N, C, G, cpg;
  float* partial; int rows;          // [N][rows][4][C]: sum m, sum m*xhat3, sum du*r, count(r == chmax)
  const float* coef;                 // [N][4][C]: GN c0, GN c1, d avg / S, d max / ties
  void* dc3; void* dr; int acc_c3, acc_r;
};

// one warp per position: d att_pre = (sum_c m*u) * att*(1-att), 1/ties of the channel arg-max
template <typename T>
__global__ void __launch_bounds__(256) cbam_bwd_pos_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ r,
                                                            const float* __restrict__ cscale, const float* __restrict__ sp,
                                                            const float* __restrict__ att, long long S, int C, long long total,
                                                            float* __restrict__ dpre, float* __restrict__ tinv) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pos = warp_id; pos < total; pos += nwarps) {
    const long long n = pos / S;
    const float* cs = cscale + n * C;
    const float spmax = sp[pos * 2 + 1];
    float datt = 0.f, ties = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      float dv[8], yv[8], rv[8];
      Vec8<T>::load(dy + pos * C + c, dv);
      Vec8<T>::load(y + pos * C + c, yv);
      Vec8<T>::load(r + pos * C + c, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = yv[j] > 0.f ? dv[j] : 0.f;
        const float u = __fmul_rn(rv[j], cs[c + j]);
        datt = fmaf(m, u, datt);
        ties += (u == spmax) ? 1.f : 0.f;
      }
    }
    datt = warp_sum(datt);
    ties = warp_sum(ties);
    if (lane == 0) {
      const float a = att[pos];
      dpre[pos] = datt * a * (1.f - a);
      tinv[pos] = 1.f / fmaxf(ties, 1.f);
    }
  }
}

// gradient of the 7x7x7 'SAME' conv w.r.t. its 2-channel input, pre-divided for the mean (1/C) and max (1/ties) paths
__global__ void __launch_bounds__(128) cbam_bwd_spconv_kernel(const float* __restrict__ dpre, const float* __restrict__ tinv,
                                                               const float* __restrict__ w, int N, int D, int H, int W, float inv_c,
                                                               float* __restrict__ dsp) {
  __shared__ float sw[343 * 2];
  for (int i = threadIdx.x; i < 686; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long total = (long long)N * D * H * W;
  for (long long z = blockIdx.x * (long long)blockDim.x + threadIdx.x; z < total; z += (long long)gridDim.x * blockDim.x) {
    long long q = z;
    const int zw = (int)(q % W); q /= W;
    const int zh = (int)(q % H); q /= H;
    const int zd = (int)(q % D);
    const int n = (int)(q / D);
    float a0 = 0.f, a1 = 0.f;
    for (int a = 0; a < 7; ++a) {
      const int od = zd - a + 3;
      if (od < 0 || od >= D) continue;
      for (int b = 0; b < 7; ++b) {
        const int oh = zh - b + 3;
        if (oh < 0 || oh >= H) continue;
        for (int e = 0; e < 7; ++e) {
          const int ow = zw - e + 3;
          if (ow < 0 || ow >= W) continue;
          const float d = dpre[(((long long)n * D + od) * H + oh) * W + ow];
          const float* ww = sw + ((a * 7 + b) * 7 + e) * 2;
          a0 = fmaf(d, ww[0], a0);
          a1 = fmaf(d, ww[1], a1);
        }
      }
    }
    dsp[z * 2 + 0] = a0 * inv_c;
    dsp[z * 2 + 1] = a1 * tinv[z];
  }
}

// filter gradient of the 7x7x7 conv: one block per tap, dw[tap][k] += sum_o dpre[o] * sp[o + tap - 3][k]
__global__ void __launch_bounds__(256) cbam_bwd_spw_kernel(const float* __restrict__ dpre, const float* __restrict__ sp, int N, int D, int H,
                                                            int W, float* dw) {
  const int tap = blockIdx.x;
  const int a = tap / 49 - 3, b = (tap / 7) % 7 - 3, e = tap % 7 - 3;
  const int total = N * D * H * W;   // < 2^31 (checked by the launcher): 32-bit index arithmetic, no 64-bit divisions
  float a0 = 0.f, a1 = 0.f;
  for (int o = threadIdx.x; o < total; o += blockDim.x) {
    int q = o;
    const int ow = q % W; q /= W;
    const int oh = q % H; q /= H;
    const int od = q % D;
    const int n = q / D;
    const int zd = od + a, zh = oh + b, zw = ow + e;
    if (zd < 0 || zd >= D || zh < 0 || zh >= H || zw < 0 || zw >= W) continue;
    const float d = dpre[o];
    const float* s = sp + (((n * D + zd) * H + zh) * W + zw) * 2;
    a0 = fmaf(d, s[0], a0);
    a1 = fmaf(d, s[1], a1);
  }
  __shared__ float r0[8], r1[8];
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = a0; r1[threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int i = 0; i < 8; ++i) { s0 += r0[i]; s1 += r1[i]; }
    dw[tap * 2 + 0] += s0;
    dw[tap * 2 + 1] += s1;
  }
}

template <typename T>
SAP3D_DEVINL void tail_common(const TailArgs& p, long long gp, int n, int c, float (&m)[8], float (&du)[8], float (&xh)[8], float (&rv)[8]) {
  const long long e = gp * p.C + c;
  float dv[8], yv[8], cv[8];
  Vec8<T>::load(reinterpret_cast<const T*>(p.dy) + e, dv);
  Vec8<T>::load(reinterpret_cast<const T*>(p.y) + e, yv);
  Vec8<T>::load(reinterpret_cast<const T*>(p.r) + e, rv);
  Vec8<T>::load(reinterpret_cast<const T*>(p.c3) + e, cv);
  const float at = p.att[gp], d0 = p.dsp[gp * 2], d1 = p.dsp[gp * 2 + 1], spmax = p.sp[gp * 2 + 1];
  const float* cs = p.cscale + (long long)n * p.C + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = n * p.G + (c + j) / p.cpg;
    m[j] = yv[j] > 0.f ? dv[j] : 0.f;
    const float u = __fmul_rn(rv[j], cs[j]);
    du[j] = fmaf(m[j], at, d0) + (u == spmax ? d1 : 0.f);
    xh[j] = (cv[j] - p.mean3[gi]) * p.rstd3[gi];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cbam_bwd_reduce_kernel(const TailArgs p) {
  __shared__ float red[8][4][64];
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.y * 64 + cv * 8;
  const int n = blockIdx.z;
  const long long per = (p.S + p.rows - 1) / p.rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.S ? pbeg + per : p.S;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (c < p.C) {
    const float* cmx = p.chmax + (long long)n * p.save_ld + c;
    for (long long pos = pbeg + pl; pos < pend; pos += 32) {
      float m[8], du[8], xh[8], rv[8];
      tail_common<T>(p, (long long)n * p.S + pos, n, c, m, du, xh, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += m[j];
        acc[1][j] += m[j] * xh[j];
        acc[2][j] += du[j] * rv[j];
        acc[3][j] += (rv[j] == cmx[j]) ? 1.f : 0.f;
      }
    }
  }
  block_fold_4x64(acc, red, p.partial + ((long long)n * p.rows + blockIdx.x) * 4 * p.C, p.C, blockIdx.y * 64);
}

// channel-attention MLP backward, one block per sample.
//   save [N][2C+2h] = avg, max, ha, hm ;  out: coef[n][2][c] = d avg / S, coef[n][3][c] = d max / ties,
//   mlpg [N][C + 2h] = d pre-sigmoid, d ha, d hm (for the weight-gradient kernel)
__global__ void __launch_bounds__(1024) cbam_bwd_mlp_kernel(const float* __restrict__ partial, int rows, int C, int hidden, long long S,
                                                            const float* __restrict__ cscale, const float* __restrict__ save,
                                                            const float* __restrict__ w0, const float* __restrict__ w1,
                                                            float* __restrict__ coef, float* __restrict__ mlpg) {
  extern __shared__ float sm[];  // dp[C], cnt[C], dha[hidden], dhm[hidden]
  float* dp = sm;
  float* cnt = sm + C;
  float* dha = sm + 2 * C;
  float* dhm = dha + hidden;
  const int n = blockIdx.x;
  const float* sv = save + (long long)n * (2 * C + 2 * hidden);
  float* mg = mlpg + (long long)n * (C + 2 * hidden);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, k = 0.f;
    for (int r = 0; r < rows; ++r) {
      a += partial[(((long long)n * rows + r) * 4 + 2) * C + c];
      k += partial[(((long long)n * rows + r) * 4 + 3) * C + c];
    }
    const float cs = cscale[(long long)n * C + c];
    const float d = a * cs * (1.f - cs);
    dp[c] = d;
    cnt[c] = fmaxf(k, 1.f);
    mg[c] = d;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int j = warp; j < hidden; j += nwarp) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w1[(long long)j * C + c], dp[c], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float a = sv[2 * C + j] > 0.f ? s : 0.f, b = sv[2 * C + hidden + j] > 0.f ? s : 0.f;
      dha[j] = a; dhm[j] = b;
      mg[C + j] = a; mg[C + hidden + j] = b;
    }
  }
  __syncthreads();
  for (int c0 = warp * 4; c0 < C; c0 += nwarp * 4) {   // one warp per 4 channels: rows of w0 are read coalesced, 4 rows in flight
    float da[4] = {0.f, 0.f, 0.f, 0.f}, dm[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < hidden; j += 32) {
      const float a = dha[j], b = dhm[j];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = c0 + u < C ? w0[(long long)(c0 + u) * hidden + j] : 0.f;
        da[u] = fmaf(w, a, da[u]);
        dm[u] = fmaf(w, b, dm[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float x = warp_sum(da[u]), y = warp_sum(dm[u]);
      if (lane == 0 && c0 + u < C) {
        coef[((long long)n * 4 + 2) * C + c0 + u] = x / (float)S;
        coef[((long long)n * 4 + 3) * C + c0 + u] = y / cnt[c0 + u];
      }
    }
  }
}

// dW0 [C][h], db0 [h], dW1 [h][C], db1 [C] (+=), summed over samples in a fixed order
__global__ void __launch_bounds__(256) cbam_bwd_mlp_wgrad_kernel(const float* __restrict__ save, const float* __restrict__ mlpg, int N, int C,
                                                                  int hidden, float* dw0, float* db0, float* dw1, float* db1) {
  const long long total = (long long)C * hidden;
  const int sld = 2 * C + 2 * hidden, gld = C + 2 * hidden;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    {
      const int c = (int)(i / hidden), j = (int)(i % hidden);
      float s = 0.f;
      for (int n = 0; n < N; ++n)
        s += save[(long long)n * sld + c] * mlpg[(long long)n * gld + C + j] + save[(long long)n * sld + C + c] * mlpg[(long long)n * gld + C + hidden + j];
      dw0[i] += s;
    }
    {
      const int j = (int)(i / C), c = (int)(i % C);
      float s = 0.f;
      for (int n = 0; n < N; ++n)
        s += (save[(long long)n * sld + 2 * C + j] + save[(long long)n * sld + 2 * C + hidden + j]) * mlpg[(long long)n * gld + c];
      dw1[i] += s;
    }
    if (i < hidden) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += mlpg[(long long)n * gld + C + i] + mlpg[(long long)n * gld + C + hidden + i];
      db0[i] += s;
    }
    if (i < C) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += mlpg[(long long)n * gld + i];
      db1[i] += 2.f * s;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cbam_bwd_apply_kernel(const TailArgs p) {
  const long long nvec = (long long)p.N * p.S * p.C / 8;
  T* dc3 = reinterpret_cast<T*>(p.dc3);
  T* dr = reinterpret_cast<T*>(p.dr);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long e = v * 8;
    const long long gp = e / p.C;
    const int c = (int)(e - gp * p.C);
    const int n = (int)(gp / p.S);
    float m[8], du[8], xh[8], rv[8];
    tail_common<T>(p, gp, n, c, m, du, xh, rv);
    const float* cf = p.coef + (long long)n * 4 * p.C + c;
    const long long si = (long long)n * p.C + c;
    if (dc3) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p.s3[si + j] * m[j] - cf[j] - xh[j] * cf[p.C + j];
      if (p.acc_c3) {
        float old[8];
        Vec8<T>::load(dc3 + e, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      Vec8<T>::store(dc3 + e, o);
    }
    if (dr) {
      const float* cmx = p.chmax + (long long)n * p.save_ld + c;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(du[j], p.cscale[si + j], cf[2 * p.C + j]) + (rv[j] == cmx[j] ? cf[3 * p.C + j] : 0.f);
      if (p.acc_r) {
        float old[8];
        Vec8<T>::load(dr + e, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      Vec8<T>::store(dr + e, o);
    }
  }
}

// dy [P][ca+cb] -> da [P][ca], db [P][cb]
template <typename T>
__global__ void __launch_bounds__(256) split_channels_kernel(const T* __restrict__ dy, T* da, int acc_a, T* db, int acc_b, long long P, int ca,
                                                              int cb) {
  const int cv = (ca + cb) / 8;
  const long long nvec = P * cv;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long pos = v / cv;
    const int c = (int)(v - pos * cv) * 8;
    float d[8];
    Vec8<T>::load(dy + v * 8, d);
    T* dst = c < ca ? (da ? da + pos * ca + c : nullptr) : (db ? db + pos * cb + (c - ca) : nullptr);
    if (!dst) continue;
    if (c < ca ? acc_a : acc_b) {
      float old[8];
      Vec8<T>::load(dst, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += old[j];
    }
    Vec8<T>::store(dst, d);
  }
}

}  // namespace

extern "C" {

size_t sap3d_gn_bwd_workspace(int32_t N, int64_t S, int32_t C) {
  const size_t rows = (size_t)sap3d_sample_stats_rows(S, C, N);
  // partial [N][rows][4][C] + coef [N][4][C] + (CBAM tail) dpre, tinv [N*S], dsp [N*S][2], mlpg [N][C + 2*(C/8)]
  return ((size_t)N * rows * 4 * C + (size_t)N * 8 * C + (size_t)N * S * 4 + (size_t)N * (C + 2 * (C / 8 + 1)) + 64) * sizeof(float);
}

int sap3d_gn_act_bwd(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                     const float* rstd1, const float* gamma1, int32_t relu1, const void* b, const float* s2, const float* t2,
                     const float* mean2, const float* rstd2, const float* gamma2, int32_t relu2, int32_t relu_out, int32_t N,
                     int64_t S, int32_t C, int32_t G, void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1,
                     float* dbeta1, float* dgamma2, float* dbeta2, void* workspace, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0 || C % G != 0 || C / G > 32) return set_error("gn_act_bwd: unsupported C=%d G=%d", C, G);
  if (!workspace || !dy || !a || !s1 || !mean1 || !rstd1 || !gamma1) return set_error("gn_act_bwd: NULL argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnBwdArgs p;
  memset(&p, 0, sizeof(p));
  p.dy = dy; p.a = a; p.b = b;
  p.s1 = s1; p.t1 = t1; p.mean1 = mean1; p.rstd1 = rstd1;
  p.s2 = s2; p.t2 = t2; p.mean2 = mean2; p.rstd2 = rstd2;
  p.S = S; p.N = N; p.C = C; p.G = G; p.cpg = C / G;
  p.relu1 = relu1; p.relu2 = relu2; p.relu_out = relu_out;
  p.rows = sap3d_sample_stats_rows(S, C, N);
  float* ws = reinterpret_cast<float*>(workspace);
  p.partial = ws;
  float* coef = ws + (size_t)N * p.rows * 4 * C;
  p.coef = coef;
  p.da = da; p.db = db; p.acc_a = acc_a; p.acc_b = acc_b;
  dim3 rgrid(p.rows, (C + 63) / 64, N);
  if (dtype == SAP3D_BF16) gn_bwd_reduce_kernel<bf16><<<rgrid, 256, 0, st>>>(p);
  else gn_bwd_reduce_kernel<float><<<rgrid, 256, 0, st>>>(p);
  if (check_launch("gn_act_bwd reduce")) return 1;
  const bool norm2 = b && s2;
  float* vsum = coef + (size_t)N * 4 * C;
  gn_bwd_finalize_kernel<<<dim3(G, N), 128, 0, st>>>(p.partial, p.rows, N, C, G, (double)S * p.cpg, gamma1, rstd1, norm2 ? gamma2 : gamma1,
                                                     norm2 ? rstd2 : rstd1, norm2 ? 4 : 2, coef, vsum);
  gn_bwd_param_kernel<<<((norm2 ? 4 : 2) * C + 255) / 256, 256, 0, st>>>(vsum, N, C, norm2 ? 4 : 2, dgamma1, dbeta1, norm2 ? dgamma2 : nullptr,
                                                                        norm2 ? dbeta2 : nullptr);
  if (check_launch("gn_act_bwd finalize")) return 1;
  if (da || db) {
    const long long nvec = (long long)N * S * C / 8;
    if (dtype == SAP3D_BF16) gn_bwd_apply_kernel<bf16><<<ew_grid(nvec), 256, 0, st>>>(p);
    else gn_bwd_apply_kernel<float><<<ew_grid(nvec), 256, 0, st>>>(p);
    if (check_launch("gn_act_bwd apply")) return 1;
  }
  return 0;
}

int sap3d_cbam_tail_bwd(int32_t dtype, const void* dy, const void* y, const void* c3, const float* s3, const float* mean3,
                        const float* rstd3, const float* gamma3, const void* r, int32_t N, int32_t D, int32_t H, int32_t W,
                        int32_t C, int32_t G, int32_t hidden, const float* w0, const float* w1, const float* w_sp,
                        const float* cscale, const float* sp, const float* att, const float* save, void* dc3, int32_t acc_c3,
                        void* dr, int32_t acc_r, float* dgamma3, float* dbeta3, float* dw0, float* db0, float* dw1, float* db1,
                        float* dw_sp, void* workspace, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0 || C % G != 0 || C / G > 32) return set_error("cbam_tail_bwd: unsupported C=%d G=%d", C, G);
  if (hidden != C / 8) return set_error("cbam_tail_bwd: hidden must be C/8");
  if (!workspace || !save) return set_error("cbam_tail_bwd: NULL workspace / saved state");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long S = (long long)D * H * W, total = (long long)N * S;
  const int rows = sap3d_sample_stats_rows(S, C, N);
  float* ws = reinterpret_cast<float*>(workspace);
  float* partial = ws;               ws += (size_t)N * rows * 4 * C;
  float* coef = ws;                  ws += (size_t)N * 4 * C;
  float* vsum = ws;                  ws += (size_t)N * 4 * C;
  float* dpre = ws;                  ws += total;
  float* tinv = ws;                  ws += total;
  float* dsp = ws;                   ws += 2 * total;
  float* mlpg = ws;
  TailArgs p;
  memset(&p, 0, sizeof(p));
  p.dy = dy; p.y = y; p.c3 = c3; p.r = r;
  p.s3 = s3; p.mean3 = mean3; p.rstd3 = rstd3;
  p.cscale = cscale; p.sp = sp; p.att = att;
  p.save_ld = 2 * C + 2 * hidden;
  p.chmax = save + C;
  p.dsp = dsp;
  p.S = S; p.N = N; p.C = C; p.G = G; p.cpg = C / G;
  p.partial = partial; p.rows = rows; p.coef = coef;
  p.dc3 = dc3; p.dr = dr; p.acc_c3 = acc_c3; p.acc_r = acc_r;
  const bool bf = dtype == SAP3D_BF16;
  {
    const long long need = (total * 32 + 255) / 256;
    const int blocks = (int)(need > 148 * 8 ? 148 * 8 : need);
    if (bf) cbam_bwd_pos_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(y), reinterpret_cast<const bf16*>(r), cscale, sp, att, S, C, total, dpre, tinv);
    else cbam_bwd_pos_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(y), reinterpret_cast<const float*>(r), cscale, sp, att, S, C, total, dpre, tinv);
    if (check_launch("cbam_bwd pos")) return 1;
  }
  {
    const long long need = (total + 127) / 128;
    cbam_bwd_spconv_kernel<<<(int)(need > 148 * 16 ? 148 * 16 : need), 128, 0, st>>>(dpre, tinv, w_sp, N, D, H, W, 1.f / (float)C, dsp);
    if (check_launch("cbam_bwd spconv")) return 1;
    if (dw_sp) {
      cbam_bwd_spw_kernel<<<343, 256, 0, st>>>(dpre, sp, N, D, H, W, dw_sp);
      if (check_launch("cbam_bwd spw")) return 1;
    }
  }
  dim3 rgrid(rows, (C + 63) / 64, N);
  if (bf) cbam_bwd_reduce_kernel<bf16><<<rgrid, 256, 0, st>>>(p);
  else cbam_bwd_reduce_kernel<float><<<rgrid, 256, 0, st>>>(p);
  if (check_launch("cbam_bwd reduce")) return 1;
  gn_bwd_finalize_kernel<<<dim3(G, N), 128, 0, st>>>(partial, rows, N, C, G, (double)S * p.cpg, gamma3, rstd3, gamma3, rstd3, 2, coef, vsum);
  gn_bwd_param_kernel<<<(2 * C + 255) / 256, 256, 0, st>>>(vsum, N, C, 2, dgamma3, dbeta3, nullptr, nullptr);
  if (check_launch("cbam_bwd gn finalize")) return 1;
  cbam_bwd_mlp_kernel<<<N, C >= 512 ? 1024 : 256, (size_t)(2 * C + 2 * hidden) * sizeof(float), st>>>(partial, rows, C, hidden, S, cscale, save, w0, w1, coef, mlpg);
  if (check_launch("cbam_bwd mlp")) return 1;
  if (dw0 && db0 && dw1 && db1) {
    cbam_bwd_mlp_wgrad_kernel<<<ew_grid((long long)C * hidden), 256, 0, st>>>(save, mlpg, N, C, hidden, dw0, db0, dw1, db1);
    if (check_launch("cbam_bwd mlp wgrad")) return 1;
  }
  if (dc3 || dr) {
    const long long nvec = total * C / 8;
    if (bf) cbam_bwd_apply_kernel<bf16><<<ew_grid(nvec), 256, 0, st>>>(p);
    else cbam_bwd_apply_kernel<float><<<ew_grid(nvec), 256, 0, st>>>(p);
    if (check_launch("cbam_bwd apply")) return 1;
  }
  return 0;
}

int sap3d_split_channels(int32_t dtype, const void* dy, void* da, int32_t acc_a, void* db, int32_t acc_b, int64_t P, int32_t ca,
                         int32_t cb, void* stream) {
  if (require_device()) return 1;
  if (ca % 8 != 0 || cb % 8 != 0) return set_error("split_channels: channel counts must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long nvec = P * (ca + cb) / 8;
  if (dtype == SAP3D_BF16)
    split_channels_kernel<bf16><<<ew_grid(nvec), 256, 0, st>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(da), acc_a, reinterpret_cast<bf16*>(db), acc_b, P, ca, cb);
  else
    split_channels_kernel<float><<<ew_grid(nvec), 256, 0, st>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<float*>(da), acc_a, reinterpret_cast<float*>(db), acc_b, P, ca, cb);
  return check_launch("split_channels");
}

}  // extern "C"
