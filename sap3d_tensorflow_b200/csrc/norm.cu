// BatchNorm / GroupNorm statistics finalisation, the fused affine + ReLU + residual "apply" pass and
// its backward, for NDHWC activations (bf16 or f32 storage, fp32/fp64 statistics).
//
// Replaces, per reference layer, the chain TensorFlow runs for tf.layers.batch_normalization on 5-D
// input (moments + ~5 Eigen elementwise kernels), tf.nn.relu and the residual adds
// (p3d.py:56-81,88,97,114,127,133-134; utils/network.py:65-94).  Bandwidth-bound: 128-bit accesses,
// one read per operand and one write.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

// ------------------------------------------------------------------------------------------------
// finalize: partial [rows][2][C] -> mean/var -> scale/shift (+ moving-average update)
// block = 32 channels x 32 row lanes, deterministic fp64 reduction
// ------------------------------------------------------------------------------------------------
// CPB channels x (1024 / CPB) row lanes per block.  CPB = 32 for few-row layers; CPB = 8 for the decoder's 1000+ statistics rows
// (4x more blocks, 4x shorter serial walk per lane; a lane still reads 32 contiguous bytes per row and statistic).
template <int CPB>
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ stats, int rows, int C, double count,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* moving_mean, float* moving_var, int training,
                                                            float momentum, float eps, float* scale, float* shift,
                                                            float* save_mean, float* save_rstd) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  constexpr int LANES = 1024 / CPB;
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int c = blockIdx.x * CPB + cx;
  __shared__ double s1[LANES][CPB + 1], s2[LANES][CPB + 1];
  double a = 0.0, b = 0.0;
  if (training && c < C) {
    for (int r = ry; r < rows; r += LANES) {
      a += (double)stats[((long long)r * 2 + 0) * C + c];
      b += (double)stats[((long long)r * 2 + 1) * C + c];
    }
  }
  s1[ry][cx] = a;
  s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && c < C) {
    float mean, var;
    if (training) {
      double ta = 0.0, tb = 0.0;
      for (int i = 0; i < LANES; ++i) {      // fixed order: deterministic
        ta += s1[i][cx];
        tb += s2[i][cx];
      }
      const double m = ta / count;
      double v = tb / count - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m;
      var = (float)v;
      if (moving_mean) {
        moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
        moving_var[c] = moving_var[c] * momentum + var * (1.f - momentum);
      }
    } else {
      mean = moving_mean[c];
      var = moving_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f;
    const float bt = beta ? beta[c] : 0.f;
    scale[c] = g * rstd;
    shift[c] = bt - mean * g * rstd;
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = rstd;
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: x [N][S][C] -> per (n, group) mean / rstd -> per (n, c) scale / shift
// one block per (n, group)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, long long S, int C, int G,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float* scale, float* shift, float* save_mean,
                                                        float* save_rstd) {
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const T* xb = x + (long long)n * S * C + g * cpg;
  double a = 0.0, b = 0.0;
  const long long total = S * cpg;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const long long s = i / cpg;
    const int c = (int)(i - s * cpg);
    const float v = to_f32<T>(xb[s * C + c]);
    a += v;
    b += (double)v * v;
  }
  __shared__ double sa[8], sb[8];
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = a;
    sb[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < 8; ++i) {
      ta += sa[i];
      tb += sb[i];
    }
    const double m = ta / (double)total;
    double v = tb / (double)total - m * m;
    if (v < 0.0) v = 0.0;
    const float rstd = (float)(1.0 / sqrt(v + (double)eps));
    if (threadIdx.x == 0) {
      save_mean[n * G + g] = (float)m;
      save_rstd[n * G + g] = rstd;
    }
    for (int c = threadIdx.x; c < cpg; c += 32) {
      const int ch = g * cpg + c;
      const float gm = gamma[ch];
      scale[n * C + ch] = gm * rstd;
      shift[n * C + ch] = beta[ch] - (float)m * gm * rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// apply:  z1 = a*s1+t1 ; r1 = relu1 ? max(z1,0) : z1
//         z2 = b ? (s2 ? b*s2+t2 : b) : 0 ; r2 = relu2 ? max(z2,0) : z2
//         y  = relu_out ? max(r1+r2,0) : r1+r2
// scale index = (per-sample ? n*C : 0) + c
// ------------------------------------------------------------------------------------------------
struct ApplyArgs {
  const void* a; const float* s1; const float* t1;
  const void* b; const float* s2; const float* t2;
  void* y;
  long long P;       // positions
  long long psp;     // positions per sample (0 = per-channel scale)
  int C;
  int relu1, relu2, relu_out;
};

template <typename T>
__global__ void __launch_bounds__(256) apply_kernel(const ApplyArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  const long long nvec = p.P * p.C / 8;
  const T* a = reinterpret_cast<const T*>(p.a);
  const T* b = reinterpret_cast<const T*>(p.b);
  T* y = reinterpret_cast<T*>(p.y);
  // APPLY_INFLIGHT vectors per thread and iteration, all loads issued before the first use: at one 16-byte load in flight per
  // thread the kernel streamed at 2.2-3.1 TB/s (r02 ncu: 50 % occupancy x 16 B = 16 KB in flight per SM)
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    typename Vec8<T>::Raw ra[U], rb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        ra[u] = Vec8<T>::load_raw(a + v * 8);
        if (b) rb[u] = Vec8<T>::load_raw(b + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v >= nvec) break;
      const long long e = v * 8;
      const long long pos = e / p.C;
      const int c = (int)(e - pos * p.C);
      const long long sidx = (p.psp ? (pos / p.psp) * p.C : 0) + c;
      float av[8], r[8];
      Vec8<T>::unpack(ra[u], av);
      if (p.s1) {
        const float4 sa = *reinterpret_cast<const float4*>(p.s1 + sidx), sb = *reinterpret_cast<const float4*>(p.s1 + sidx + 4);
        const float4 ta = *reinterpret_cast<const float4*>(p.t1 + sidx), tb = *reinterpret_cast<const float4*>(p.t1 + sidx + 4);
        const float s[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(av[j], s[j], t[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = av[j];
      }
      if (p.relu1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], 0.f);
      }
      if (b) {
        float bv[8];
        Vec8<T>::unpack(rb[u], bv);
        if (p.s2) {
          const float4 sa = *reinterpret_cast<const float4*>(p.s2 + sidx), sb = *reinterpret_cast<const float4*>(p.s2 + sidx + 4);
          const float4 ta = *reinterpret_cast<const float4*>(p.t2 + sidx), tb = *reinterpret_cast<const float4*>(p.t2 + sidx + 4);
          const float s[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
          const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = fmaf(bv[j], s[j], t[j]);
        }
        if (p.relu2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = fmaxf(bv[j], 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += bv[j];
      }
      if (p.relu_out) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], 0.f);
      }
      Vec8<T>::store(y + e, r);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// fused finalize + apply for layers with few statistics rows (the whole backbone): every block owns one 64-channel
// chunk and a slab of positions; its prologue reduces the conv epilogue's per-tile sums for those 64 channels
// (deterministic fp64), block x == 0 of each chunk also publishes scale/shift/mean/rstd (the backward pass reads
// them) and updates the moving averages.  One launch instead of two (three for two-norm ops), no second grid.
// ------------------------------------------------------------------------------------------------
struct FusedNorm {
  const float* stats; int rows; double count;
  const float *gamma, *beta;
  float *mm, *mv;
  int training;
  float *scale, *shift, *save_mean, *save_rstd;
};
struct FusedApplyArgs {
  ApplyArgs ap;
  FusedNorm n[2];
  int has[2];
  int prows;        // position slabs (gridDim.x)
  float momentum, eps;
};

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_fused_kernel(const FusedApplyArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  __shared__ double red[4][2][64];
  __shared__ float s_scale[2][64], s_shift[2][64];
  const int C = p.ap.C;
  const int cbase = blockIdx.y * 64;
  {
    const int ch = threadIdx.x & 63, part = threadIdx.x >> 6;
    const int c = cbase + ch;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!p.has[k]) continue;
      const FusedNorm& nm = p.n[k];
      if (nm.training) {
        double a = 0.0, b = 0.0;
        if (c < C) {
          int r = part;
          for (; r + 12 < nm.rows; r += 16) {   // 4 rows x 2 sums in flight per iteration
            float va[4], vb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              va[u] = nm.stats[((long long)(r + 4 * u) * 2 + 0) * C + c];
              vb[u] = nm.stats[((long long)(r + 4 * u) * 2 + 1) * C + c];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { a += (double)va[u]; b += (double)vb[u]; }
          }
          for (; r < nm.rows; r += 4) {
            a += (double)nm.stats[((long long)r * 2 + 0) * C + c];
            b += (double)nm.stats[((long long)r * 2 + 1) * C + c];
          }
        }
        red[part][0][ch] = a;
        red[part][1][ch] = b;
      }
      __syncthreads();
      if (part == 0 && c < C) {
        float mean, var;
        if (nm.training) {
          const double ta = red[0][0][ch] + red[1][0][ch] + red[2][0][ch] + red[3][0][ch];
          const double tb = red[0][1][ch] + red[1][1][ch] + red[2][1][ch] + red[3][1][ch];
          const double m = ta / nm.count;
          double v = tb / nm.count - m * m;
          if (v < 0.0) v = 0.0;
          mean = (float)m;
          var = (float)v;
          if (blockIdx.x == 0 && nm.mm) {
            nm.mm[c] = nm.mm[c] * p.momentum + mean * (1.f - p.momentum);
            nm.mv[c] = nm.mv[c] * p.momentum + var * (1.f - p.momentum);
          }
        } else {
          mean = nm.mm[c];
          var = nm.mv[c];
        }
        const float rstd = rsqrtf(var + p.eps);
        const float g = nm.gamma ? nm.gamma[c] : 1.f, bt = nm.beta ? nm.beta[c] : 0.f;
        const float sc = g * rstd, sh = bt - mean * g * rstd;
        s_scale[k][ch] = sc;
        s_shift[k][ch] = sh;
        if (blockIdx.x == 0) {
          nm.scale[c] = sc;
          nm.shift[c] = sh;
          if (nm.save_mean) nm.save_mean[c] = mean;
          if (nm.save_rstd) nm.save_rstd[c] = rstd;
        }
      }
      __syncthreads();
    }
  }
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = cbase + cv * 8;
  if (c >= C) return;
  const T* a = reinterpret_cast<const T*>(p.ap.a);
  const T* b = reinterpret_cast<const T*>(p.ap.b);
  T* y = reinterpret_cast<T*>(p.ap.y);
  const long long per = (p.ap.P + p.prows - 1) / p.prows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.ap.P ? pbeg + per : p.ap.P;
  float s1[8], t1[8], s2[8], t2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[j] = p.has[0] ? s_scale[0][cv * 8 + j] : 1.f;
    t1[j] = p.has[0] ? s_shift[0][cv * 8 + j] : 0.f;
    s2[j] = p.has[1] ? s_scale[1][cv * 8 + j] : 1.f;
    t2[j] = p.has[1] ? s_shift[1][cv * 8 + j] : 0.f;
  }
  for (long long pos = pbeg + pl; pos < pend; pos += 32) {
    const long long e = pos * C + c;
    float av[8], r[8];
    Vec8<T>::load(a + e, av);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r[j] = fmaf(av[j], s1[j], t1[j]);
      if (p.ap.relu1) r[j] = fmaxf(r[j], 0.f);
    }
    if (b) {
      float bv[8];
      Vec8<T>::load(b + e, bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = p.has[1] ? fmaf(bv[j], s2[j], t2[j]) : bv[j];
        if (p.ap.relu2) z = fmaxf(z, 0.f);
        r[j] += z;
      }
    }
    if (p.ap.relu_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaxf(r[j], 0.f);
    }
    Vec8<T>::store(y + e, r);
  }
}

// ------------------------------------------------------------------------------------------------
// backward of apply.  Pass 1: per-channel (or per sample-group for GN) reductions
//     S1a = sum g1, S1b = sum g1*xhat1, S2a = sum g2, S2b = sum g2*xhat2
// with g = dy * mask(relu_out) * mask(relu branch), xhat = (raw - mean) * rstd.
// Pass 2: d raw = (gamma*rstd) * (g - S_a/M - xhat * S_b/M)   (batch-statistics norm)
//         d raw = g * scale                                      (frozen statistics)
//         d b   = g2 (plain tensor), optionally accumulated.
// ------------------------------------------------------------------------------------------------
struct ApplyBwdArgs {
  const void* dy;
  const void* a; const float* s1; const float* t1; const float* mean1; const float* rstd1;
  const void* b; const float* s2; const float* t2; const float* mean2; const float* rstd2;
  long long P, psp;
  int C, G;           // G > 0: GroupNorm (statistics per sample and group of C/G channels)
  int relu1, relu2, relu_out;
  // pass 1
  float* partial;     // [rows][4][C]
  double* totals;     // [4][C] (cooperative kernel, two-level reduction)
  int rows;
  // pass 2
  const float* coef;  // [4][C] (BN) or [4][N*C] (GN): S1a/M, S1b/M, S2a/M, S2b/M
  void* da; void* db;
  int batch_stats1, batch_stats2, acc_a, acc_b;
};

// masks and normalised values of 8 consecutive channels of one position from the already-loaded operand values
SAP3D_DEVINL void bwd_compute(const ApplyBwdArgs& p, const float (&d)[8], const float (&av)[8], const float (&bv)[8], long long sidx,
                              long long midx, float (&g1)[8], float (&g2)[8], float (&xh1)[8], float (&xh2)[8]) {
  float z1[8], z2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = p.s1 ? p.s1[sidx + j] : 1.f, t = p.t1 ? p.t1[sidx + j] : 0.f;
    z1[j] = fmaf(av[j], s, t);
    xh1[j] = p.mean1 ? (av[j] - p.mean1[midx + (p.G ? 0 : j)]) * p.rstd1[midx + (p.G ? 0 : j)] : 0.f;
  }
  if (p.b) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = p.s2 ? p.s2[sidx + j] : 1.f, t = p.s2 ? p.t2[sidx + j] : 0.f;
      z2[j] = fmaf(bv[j], s, t);
      xh2[j] = p.mean2 ? (bv[j] - p.mean2[midx + (p.G ? 0 : j)]) * p.rstd2[midx + (p.G ? 0 : j)] : 0.f;
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

template <typename T>
SAP3D_DEVINL void bwd_common(const ApplyBwdArgs& p, long long e, long long sidx, long long midx, float (&g1)[8], float (&g2)[8],
                             float (&xh1)[8], float (&xh2)[8]) {
  float d[8], av[8], bv[8];
  Vec8<T>::load(reinterpret_cast<const T*>(p.dy) + e, d);
  Vec8<T>::load(reinterpret_cast<const T*>(p.a) + e, av);
  if (p.b) {
    Vec8<T>::load(reinterpret_cast<const T*>(p.b) + e, bv);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = 0.f;
  }
  bwd_compute(p, d, av, bv, sidx, midx, g1, g2, xh1, xh2);
}

// grid = (rows, C/64): each block reduces a slab of positions for one 64-channel chunk.
// 256 threads = 8 channel-vector lanes (64 channels, one 128-byte line per position) x 32 position lanes.
template <typename T>
__global__ void __launch_bounds__(256) apply_bwd_reduce_kernel(const ApplyBwdArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  __shared__ float red[8][4][64];  // [warp][sum][channel]
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.y * 64 + cv * 8;
  const long long per = (p.P + p.rows - 1) / p.rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.P ? pbeg + per : p.P;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (c < p.C) {
    for (long long pos = pbeg + pl; pos < pend; pos += 32) {
      float g1[8], g2[8], xh1[8], xh2[8];
      bwd_common<T>(p, pos * p.C + c, c, c, g1, g2, xh1, xh2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += g1[j];
        acc[1][j] += g1[j] * xh1[j];
        acc[2][j] += g2[j];
        acc[3][j] += g2[j] * xh2[j];
      }
    }
  }
  // lanes with equal (lane & 7) hold the same channels: fold the 4 positions of a warp
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
  {
    const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;  // 4 sums x 64 channels = 256 threads
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][i][ch];
    const int cc = blockIdx.y * 64 + ch;
    if (cc < p.C) p.partial[((long long)blockIdx.x * 4 + i) * p.C + cc] = v;
  }
}

// partial [rows][4][C] -> coef [4][C] (divided by M) ; dgamma/dbeta accumulation (fp32, +=)
// block = 32 channels x 32 row lanes, deterministic fp64 reduction
__global__ void __launch_bounds__(1024) apply_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int C, double M,
                                                                   float* coef, float* dgamma1, float* dbeta1, float* dgamma2,
                                                                   float* dbeta2) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  __shared__ double sh[4][32][33];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  double s[4] = {0, 0, 0, 0};
  if (c < C)
    for (int r = ry; r < rows; r += 32)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += (double)partial[((long long)r * 4 + i) * C + c];
#pragma unroll
  for (int i = 0; i < 4; ++i) sh[i][ry][cx] = s[i];
  __syncthreads();
  if (ry < 4 && c < C) {
    double t = 0.0;
    for (int r = 0; r < 32; ++r) t += sh[ry][r][cx];
    coef[ry * C + c] = (float)(t / M);
    float* dst = ry == 0 ? dbeta1 : (ry == 1 ? dgamma1 : (ry == 2 ? dbeta2 : dgamma2));
    if (dst) dst[c] += (float)t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) apply_bwd_kernel(const ApplyBwdArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  const long long nvec = p.P * p.C / 8;
  T* da = reinterpret_cast<T*>(p.da);
  T* db = reinterpret_cast<T*>(p.db);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long e = v * 8;
    const long long pos = e / p.C;
    const int c = (int)(e - pos * p.C);
    float g1[8], g2[8], xh1[8], xh2[8];
    bwd_common<T>(p, e, c, c, g1, g2, xh1, xh2);
    if (da) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = p.s1 ? p.s1[c + j] : 1.f;
        o[j] = p.batch_stats1 ? s * (g1[j] - p.coef[0 * p.C + c + j] - xh1[j] * p.coef[1 * p.C + c + j]) : s * g1[j];
      }
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
      for (int j = 0; j < 8; ++j) {
        const float s = p.s2 ? p.s2[c + j] : 1.f;
        o[j] = p.batch_stats2 ? s * (g2[j] - p.coef[2 * p.C + c + j] - xh2[j] * p.coef[3 * p.C + c + j]) : s * g2[j];
      }
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

// ------------------------------------------------------------------------------------------------
// The same two passes for the common big-tensor case y = relu?(norm(a)) with NO second operand (every decoder conv / deconv
// -> BN -> ReLU, the stem).  The generic apply kernel re-reads ~100 per-channel constants per 8 elements (LSU-bound, 1.4 TB/s
// in ncu) and the generic reduce kernel keeps one position per thread in flight at 127 registers (2.0 TB/s).  Here a thread
// owns FOUR fixed channels, so its constants live in 16 registers, and FOUR positions are in flight per thread; at 80
// registers three 256-thread blocks are resident per SM (48 KB of loads in flight, what 6.5 TB/s x ~1 us latency needs).
//   mask  : z = a*s + t;   g = dy unless (relu_out && !(relu?(z) > 0)) or (relu && !(z > 0))
//   reduce: S_a = sum g, S_b = sum g*xhat,  xhat = a*rstd - mean*rstd
//   apply : d a = s*g - a*A - B,  A = s*c1*rstd,  B = s*(c0 - c1*mean*rstd)   (== s*(g - c0 - xhat*c1));  frozen: s*g
// block = 16 channel lanes (64 channels) x 16 position lanes; grid = (position slabs, C/64).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<bf16> {
  typedef uint2 Raw;
  static SAP3D_DEVINL Raw load(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
  static SAP3D_DEVINL void unpack(const Raw& u, float (&v)[4]) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static SAP3D_DEVINL void store(bf16* p, const float (&v)[4]) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <> struct Vec4<float> {
  typedef float4 Raw;
  static SAP3D_DEVINL Raw load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static SAP3D_DEVINL void unpack(const Raw& u, float (&v)[4]) { v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
  static SAP3D_DEVINL void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

SAP3D_DEVINL void nob_mask(const float (&d)[4], const float (&av)[4], const float (&s)[4], const float (&t)[4], int relu1, int relu_out,
                           float (&g)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float z = fmaf(av[j], s[j], t[j]);
    const float r1 = relu1 ? fmaxf(z, 0.f) : z;
    float u = d[j];
    if (relu_out && !(r1 > 0.f)) u = 0.f;
    g[j] = (relu1 && !(z > 0.f)) ? 0.f : u;
  }
}

constexpr int NOB_PLANES = 16;     // position lanes per block

template <typename T, int NOB_INFLIGHT>   // NOB_INFLIGHT = positions per thread per iteration (loads issued before any is consumed)
__global__ void __launch_bounds__(256, 3) apply_bwd_reduce_nob_kernel(const ApplyBwdArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  __shared__ float red[8][2][2][64];   // [warp][position lane pair within the warp][sum][channel]
  const int cv = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = blockIdx.y * 64 + cv * 4;
  const long long per = (p.P + p.rows - 1) / p.rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.P ? pbeg + per : p.P;
  const T* dy = reinterpret_cast<const T*>(p.dy);
  const T* a = reinterpret_cast<const T*>(p.a);
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (c < p.C) {
    float s[4], t[4], rs[4], mr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = p.s1 ? p.s1[c + j] : 1.f;
      t[j] = p.t1 ? p.t1[c + j] : 0.f;
      rs[j] = p.mean1 ? p.rstd1[c + j] : 0.f;
      mr[j] = p.mean1 ? p.mean1[c + j] * rs[j] : 0.f;
    }
    const long long step = (long long)NOB_PLANES * p.C;
    long long pos = pbeg + pl;
    for (; pos + (NOB_INFLIGHT - 1) * NOB_PLANES < pend; pos += NOB_PLANES * NOB_INFLIGHT) {   // all loads first
      const T* pd = dy + pos * p.C + c;
      const T* pa = a + pos * p.C + c;
      typename Vec4<T>::Raw rd[NOB_INFLIGHT], ra[NOB_INFLIGHT];
#pragma unroll
      for (int u = 0; u < NOB_INFLIGHT; ++u) {
        rd[u] = Vec4<T>::load(pd + u * step);
        ra[u] = Vec4<T>::load(pa + u * step);
      }
#pragma unroll
      for (int u = 0; u < NOB_INFLIGHT; ++u) {
        float d[4], av[4], g[4];
        Vec4<T>::unpack(rd[u], d);
        Vec4<T>::unpack(ra[u], av);
        nob_mask(d, av, s, t, p.relu1, p.relu_out, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] += g[j];
          acc[1][j] += g[j] * fmaf(av[j], rs[j], -mr[j]);
        }
      }
    }
    for (; pos < pend; pos += NOB_PLANES) {   // ragged tail of the slab
      float d[4], av[4], g[4];
      Vec4<T>::unpack(Vec4<T>::load(dy + pos * p.C + c), d);
      Vec4<T>::unpack(Vec4<T>::load(a + pos * p.C + c), av);
      nob_mask(d, av, s, t, p.relu1, p.relu_out, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0][j] += g[j];
        acc[1][j] += g[j] * fmaf(av[j], rs[j], -mr[j]);
      }
    }
  }
  // a warp holds two position lanes (lane >> 4) of the same 64 channels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp][lane >> 4][i][(lane & 15) * 4 + j] = acc[i][j];
  __syncthreads();
  {
    const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;  // 4 sums x 64 channels; sums 2, 3 (second operand) are zero
    float v = 0.f;
    if (i < 2) {
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][0][i][ch] + red[w][1][i][ch];
    }
    const int cc = blockIdx.y * 64 + ch;
    if (cc < p.C) p.partial[((long long)blockIdx.x * 4 + i) * p.C + cc] = v;
  }
}

// grid = (slabs, C/64); the slab count is independent of the reduce pass
template <typename T, bool ACC, int NOB_INFLIGHT>
__global__ void __launch_bounds__(256, 3) apply_bwd_nob_kernel(const ApplyBwdArgs p) {
  pdl_wait();   // (programmatic dependent launch: the predecessor kernel has completed from here on)
  pdl_launch_dependents();
  const int cv = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int c = blockIdx.y * 64 + cv * 4;
  if (c >= p.C) return;
  const long long per = (p.P + gridDim.x - 1) / gridDim.x;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.P ? pbeg + per : p.P;
  const T* dy = reinterpret_cast<const T*>(p.dy);
  const T* a = reinterpret_cast<const T*>(p.a);
  T* da = reinterpret_cast<T*>(p.da);
  float s[4], t[4], A[4], B[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s[j] = p.s1 ? p.s1[c + j] : 1.f;
    t[j] = p.t1 ? p.t1[c + j] : 0.f;
    if (p.batch_stats1) {
      const float rs = p.rstd1[c + j], mu = p.mean1[c + j];
      const float c0 = p.coef[c + j], c1 = p.coef[p.C + c + j];
      A[j] = s[j] * c1 * rs;
      B[j] = s[j] * (c0 - c1 * mu * rs);
    } else {
      A[j] = 0.f;
      B[j] = 0.f;
    }
  }
  const long long step = (long long)NOB_PLANES * p.C;
  long long pos = pbeg + pl;
  for (; pos + (NOB_INFLIGHT - 1) * NOB_PLANES < pend; pos += NOB_PLANES * NOB_INFLIGHT) {   // all loads first
    const long long e = pos * p.C + c;
    typename Vec4<T>::Raw rd[NOB_INFLIGHT], ra[NOB_INFLIGHT], ro[NOB_INFLIGHT];
#pragma unroll
    for (int u = 0; u < NOB_INFLIGHT; ++u) {
      rd[u] = Vec4<T>::load(dy + e + u * step);
      ra[u] = Vec4<T>::load(a + e + u * step);
      if (ACC) ro[u] = Vec4<T>::load(da + e + u * step);
    }
#pragma unroll
    for (int u = 0; u < NOB_INFLIGHT; ++u) {
      float d[4], av[4], g[4], o[4];
      Vec4<T>::unpack(rd[u], d);
      Vec4<T>::unpack(ra[u], av);
      nob_mask(d, av, s, t, p.relu1, p.relu_out, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaf(s[j], g[j], -fmaf(av[j], A[j], B[j]));
      if (ACC) {
        float old[4];
        Vec4<T>::unpack(ro[u], old);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] += old[j];
      }
      Vec4<T>::store(da + e + u * step, o);
    }
  }
  for (; pos < pend; pos += NOB_PLANES) {   // ragged tail of the slab
    const long long e = pos * p.C + c;
    float d[4], av[4], g[4], o[4];
    Vec4<T>::unpack(Vec4<T>::load(dy + e), d);
    Vec4<T>::unpack(Vec4<T>::load(a + e), av);
    nob_mask(d, av, s, t, p.relu1, p.relu_out, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fmaf(s[j], g[j], -fmaf(av[j], A[j], B[j]));
    if (ACC) {
      float old[4];
      Vec4<T>::unpack(Vec4<T>::load(da + e), old);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += old[j];
    }
    Vec4<T>::store(da + e, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-clip batch-statistics BatchNorm (+ReLU, + second operand) in ONE launch, for the inference path that normalises every clip
// of a batch on its own (placeholder(per_sample_statistics=True): what gen_pred.py's one-window-per-sess.run gives).  The generic
// path takes three launches per norm (per-sample partial sums, finalize, apply); for the backbone's small tensors that is pure
// launch latency (165 norms per forward pass).  Here one block owns (clip n, 64 channels): pass 1 sums x and x^2 over the clip's
// S positions (second operand too when it has its own norm), the block derives scale / shift in shared memory, pass 2 re-reads
// the slab (L2-resident: S x 64 channels <= 128 KB, checked by the host) and writes
//     y = relu_out?( relu1?(a * s1 + t1) + relu2?(b * s2 + t2 | b) ).
// block = 16 channel lanes (4 channels each) x 16 position lanes; grid = (N, C / 64).
// ------------------------------------------------------------------------------------------------
struct SampleNormArgs {
  const void* a; const float* g1; const float* b1;
  const void* b; const float* g2; const float* b2;   // b nullable; g2 != NULL: b has its own per-clip norm
  void* y;
  long long S;
  int C;
  float eps;
  int relu1, relu2, relu_out;
};

template <typename T>
__global__ void __launch_bounds__(256) sample_norm_apply_kernel(const SampleNormArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ double red[4][16][64];     // [sum a, sumsq a, sum b, sumsq b][position lane][channel]
  __shared__ float coef[4][64];         // scale1, shift1, scale2, shift2
  const int cv = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const int n = blockIdx.x;
  const int c = blockIdx.y * 64 + cv * 4;
  const bool live = c < p.C;
  const bool norm2 = p.b != nullptr && p.g2 != nullptr;
  const T* a = reinterpret_cast<const T*>(p.a) + (long long)n * p.S * p.C;
  const T* b = p.b ? reinterpret_cast<const T*>(p.b) + (long long)n * p.S * p.C : nullptr;
  T* y = reinterpret_cast<T*>(p.y) + (long long)n * p.S * p.C;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (live) {
    for (long long pos0 = pl; pos0 < p.S; pos0 += 64) {     // four positions per thread in flight
      typename Vec4<T>::Raw ra[4], rb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pos = pos0 + 16 * u;
        if (pos < p.S) {
          ra[u] = Vec4<T>::load(a + pos * p.C + c);
          if (norm2) rb[u] = Vec4<T>::load(b + pos * p.C + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (pos0 + 16 * u >= p.S) break;
        float v[4];
        Vec4<T>::unpack(ra[u], v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[0][j] += v[j]; acc[1][j] = fmaf(v[j], v[j], acc[1][j]); }
        if (norm2) {
          Vec4<T>::unpack(rb[u], v);
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc[2][j] += v[j]; acc[3][j] = fmaf(v[j], v[j], acc[3][j]); }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[i][pl][cv * 4 + j] = (double)acc[i][j];
  __syncthreads();
  if (threadIdx.x < 128) {     // (norm, channel): fixed-order sums over the 16 position lanes
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    const int cc = blockIdx.y * 64 + ch;
    if (cc < p.C && (which == 0 || norm2)) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int l = 0; l < 16; ++l) { s1 += red[which * 2][l][ch]; s2 += red[which * 2 + 1][l][ch]; }
      const double m = s1 / (double)p.S;
      double var = s2 / (double)p.S - m * m;
      if (var < 0.0) var = 0.0;
      const float rstd = rsqrtf((float)var + p.eps);
      const float* g = which == 0 ? p.g1 : p.g2;
      const float* bt = which == 0 ? p.b1 : p.b2;
      const float gg = g ? g[cc] : 1.f, bb = bt ? bt[cc] : 0.f;
      coef[which * 2][ch] = gg * rstd;
      coef[which * 2 + 1][ch] = bb - (float)m * gg * rstd;
    }
  }
  __syncthreads();
  if (!live) return;
  float s1[4], t1[4], s2[4], t2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s1[j] = coef[0][cv * 4 + j]; t1[j] = coef[1][cv * 4 + j];
    s2[j] = norm2 ? coef[2][cv * 4 + j] : 1.f; t2[j] = norm2 ? coef[3][cv * 4 + j] : 0.f;
  }
  for (long long pos0 = pl; pos0 < p.S; pos0 += 64) {
    typename Vec4<T>::Raw ra[4], rb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pos = pos0 + 16 * u;
      if (pos < p.S) {
        ra[u] = Vec4<T>::load(a + pos * p.C + c);
        if (b) rb[u] = Vec4<T>::load(b + pos * p.C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
    const long long pos = pos0 + 16 * u;
    if (pos >= p.S) break;
    float v[4], r[4];
    Vec4<T>::unpack(ra[u], v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[j] = fmaf(v[j], s1[j], t1[j]);
      if (p.relu1) r[j] = fmaxf(r[j], 0.f);
    }
    if (b) {
      float w[4];
      Vec4<T>::unpack(rb[u], w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float q = norm2 ? fmaf(w[j], s2[j], t2[j]) : w[j];
        if (p.relu2) q = fmaxf(q, 0.f);
        r[j] += q;
      }
    }
    if (p.relu_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = fmaxf(r[j], 0.f);
    }
    Vec4<T>::store(y + pos * p.C + c, r);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// single-launch backward (cooperative grid): reduce -> grid.sync -> per-chunk finalize in shared memory -> apply.
// grid = (rows, C/64) co-resident blocks; every block owns one 64-channel chunk and one slab of positions in both
// passes, so the second pass re-reads what the block itself just read (L1/L2 hits for backbone-sized tensors).
// Replaces the reduce / finalize / apply launch triple (3 launches, 2 dependent grid boundaries per norm op).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_coop_kernel(const ApplyBwdArgs p, const double M, float* dgamma1, float* dbeta1,
                                                          float* dgamma2, float* dbeta2) {
  __shared__ float red[8][4][64];
  __shared__ float coef[4][64];
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int cbase = blockIdx.y * 64;
  const int c = cbase + cv * 8;
  const long long per = (p.P + p.rows - 1) / p.rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < p.P ? pbeg + per : p.P;
  {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    if (c < p.C) {
      for (long long pos = pbeg + pl; pos < pend; pos += 32) {
        float g1[8], g2[8], xh1[8], xh2[8];
        bwd_common<T>(p, pos * p.C + c, c, c, g1, g2, xh1, xh2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += g1[j];
          acc[1][j] += g1[j] * xh1[j];
          acc[2][j] += g2[j];
          acc[3][j] += g2[j] * xh2[j];
        }
      }
    }
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
    if (cbase + ch < p.C) p.partial[((long long)blockIdx.x * 4 + i) * p.C + cbase + ch] = v;
  }
  __threadfence();
  cooperative_groups::this_grid().sync();
  const int nb = gridDim.x * gridDim.y;
  const bool two_level = p.rows > 96;   // few channels, many rows (stage 1/2 inner layers)
  if (two_level) {
    // the row reduction itself is spread over the grid: block j sums ALL rows of one (sum, channel) pair with a fixed
    // tree, a second grid.sync publishes the 4*C totals (instead of every thread walking `rows` values serially)
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    __shared__ double sred[8];
    for (int pair = bid; pair < 4 * p.C; pair += nb) {   // block-uniform trip count
      const int i = pair / p.C, cc = pair % p.C;
      double t = 0.0;
      for (int r = threadIdx.x; r < p.rows; r += 256) t += (double)__ldcg(p.partial + ((long long)r * 4 + i) * p.C + cc);
      t = warp_sum(t);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = t;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tt = 0.0;
        for (int w = 0; w < 8; ++w) tt += sred[w];
        p.totals[pair] = tt;
      }
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;
    const int cc = cbase + ch;
    double t = cc < p.C ? __ldcg(p.totals + i * p.C + cc) : 0.0;
    coef[i][ch] = (float)(t / M);
    if (blockIdx.x == 0 && cc < p.C) {
      float* dst = i == 0 ? dbeta1 : (i == 1 ? dgamma1 : (i == 2 ? dbeta2 : dgamma2));
      if (dst) dst[cc] += (float)t;
    }
  } else {
    const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;
    const int cc = cbase + ch;
    double t = 0.0;
    if (cc < p.C) {
      // fixed summation order (deterministic); 8 independent loads in flight per iteration hide the L2 latency
      const float* src = p.partial + (long long)i * p.C + cc;
      const long long rs = 4ll * p.C;
      int r = 0;
      for (; r + 8 <= p.rows; r += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (r + u) * rs);
#pragma unroll
        for (int u = 0; u < 8; ++u) t += (double)v[u];
      }
      for (; r < p.rows; ++r) t += (double)__ldcg(src + r * rs);
    }
    coef[i][ch] = (float)(t / M);
    if (blockIdx.x == 0 && cc < p.C) {
      float* dst = i == 0 ? dbeta1 : (i == 1 ? dgamma1 : (i == 2 ? dbeta2 : dgamma2));
      if (dst) dst[cc] += (float)t;
    }
  }
  __syncthreads();
  if (c >= p.C || (!p.da && !p.db)) return;
  T* da = reinterpret_cast<T*>(p.da);
  T* db = reinterpret_cast<T*>(p.db);
  for (long long pos = pbeg + pl; pos < pend; pos += 32) {
    const long long e = pos * p.C + c;
    float g1[8], g2[8], xh1[8], xh2[8];
    bwd_common<T>(p, e, c, c, g1, g2, xh1, xh2);
    if (da) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sc = p.s1 ? p.s1[c + j] : 1.f;
        o[j] = p.batch_stats1 ? sc * (g1[j] - coef[0][cv * 8 + j] - xh1[j] * coef[1][cv * 8 + j]) : sc * g1[j];
      }
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
      for (int j = 0; j < 8; ++j) {
        const float sc = p.s2 ? p.s2[c + j] : 1.f;
        o[j] = p.batch_stats2 ? sc * (g2[j] - coef[2][cv * 8 + j] - xh2[j] * coef[3][cv * 8 + j]) : sc * g2[j];
      }
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

// ------------------------------------------------------------------------------------------------
// Few-row tensors (stage 3 of the backbone: 784 positions at batch 8): ONE block owns 8 channels for ALL positions, so the two
// per-channel sums, the coefficients and the apply pass stay inside the block -- no grid barrier, no cooperative launch (which
// waits until the whole grid can be co-resident: ~6 us of stall per launch behind the side stream's filter gradients, r02 trace),
// deterministic.  The second pass re-reads the block's 16-channel slab (<= 100 KB) from L1 / L2.
// 256 threads = 2 channel vectors (one 32-byte sector per position) x 128 position lanes.
// ------------------------------------------------------------------------------------------------
constexpr long long SLAB_MAX_P = 1024;

// per-channel constants of the 8 channels a thread owns (registers: they are the same for every position)
struct SlabConsts {
  float s1[8], t1[8], m1[8], r1[8], s2[8], t2[8], m2[8], r2[8];
};

SAP3D_DEVINL void slab_masks(const ApplyBwdArgs& p, const SlabConsts& k, const float (&d)[8], const float (&av)[8], const float (&bv)[8],
                             float (&g1)[8], float (&g2)[8], float (&xh1)[8], float (&xh2)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z1 = fmaf(av[j], k.s1[j], k.t1[j]);
    const float z2 = p.b ? fmaf(bv[j], k.s2[j], k.t2[j]) : 0.f;
    xh1[j] = (av[j] - k.m1[j]) * k.r1[j];
    xh2[j] = (bv[j] - k.m2[j]) * k.r2[j];
    const float r1 = p.relu1 ? fmaxf(z1, 0.f) : z1;
    const float r2 = p.relu2 ? fmaxf(z2, 0.f) : z2;
    float u = d[j];
    if (p.relu_out && !(r1 + r2 > 0.f)) u = 0.f;
    g1[j] = (p.relu1 && !(z1 > 0.f)) ? 0.f : u;
    g2[j] = (p.relu2 && !(z2 > 0.f)) ? 0.f : u;
  }
}

SAP3D_DEVINL void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
SAP3D_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The block's whole slab (dy, a, b: <= 1024 positions x 8 channels each) is fetched ONCE with cp.async -- every load of the
// block in flight together, one memory latency -- into shared memory laid out [tensor][iteration][thread], so both passes are
// short rolled loops over conflict-free 16-byte shared-memory reads.  (A register-resident variant with both passes fully
// unrolled ran 2x SLOWER than the cooperative kernel: 300 KB of straight-line code executed once per block is instruction-fetch
// bound; r02 chain probe.)
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_slab_kernel(const ApplyBwdArgs p, const double M, float* dgamma1, float* dbeta1,
                                                          float* dgamma2, float* dbeta2) {
  extern __shared__ uint4 slab_sm[];
  __shared__ float red[8][4][8];   // [warp][sum][channel]
  __shared__ float coef[4][8];
  constexpr int VB = 8 * (int)sizeof(T);           // bytes of one 8-channel vector
  constexpr int Q = VB / 16;                       // 16-byte pieces per vector
  const int rl = threadIdx.x;                      // 256 position lanes, one 8-channel vector each
  const int c = blockIdx.x * 8;
  const bool live = c < p.C;
  const int nit = (int)((p.P + 255) / 256);
  const bool has_b = p.b != nullptr;
  const uint32_t sm0 = smem_u32(slab_sm);
  auto slot = [&](int t, int it) -> uint32_t { return sm0 + (uint32_t)(((t * nit + it) * 256 + (int)threadIdx.x) * VB); };
  pdl_wait();
  pdl_launch_dependents();
  if (live) {
#pragma unroll 1
    for (int it = 0; it < nit; ++it) {
      const long long pos = rl + it * 256;
      if (pos >= p.P) break;
      const long long e = pos * p.C + c;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        cp_async16(slot(0, it) + q * 16, reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.dy) + e) + q * 16);
        cp_async16(slot(1, it) + q * 16, reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.a) + e) + q * 16);
        if (has_b) cp_async16(slot(2, it) + q * 16, reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.b) + e) + q * 16);
      }
    }
  }
  SlabConsts k;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cj = live ? c + j : 0;
    k.s1[j] = p.s1 ? p.s1[cj] : 1.f;
    k.t1[j] = p.t1 ? p.t1[cj] : 0.f;
    k.m1[j] = p.mean1 ? p.mean1[cj] : 0.f;
    k.r1[j] = p.mean1 ? p.rstd1[cj] : 0.f;     // frozen statistics: x-hat is not formed (0), as in bwd_compute
    k.s2[j] = p.s2 ? p.s2[cj] : 1.f;
    k.t2[j] = p.s2 ? p.t2[cj] : 0.f;
    k.m2[j] = p.mean2 ? p.mean2[cj] : 0.f;
    k.r2[j] = p.mean2 ? p.rstd2[cj] : 0.f;
  }
  cp_async_wait_all();     // each thread reads back only what it fetched itself: no block barrier needed
  auto fetch = [&](int it, float (&dv)[8], float (&av)[8], float (&bv)[8]) {
    const T* s0 = reinterpret_cast<const T*>(reinterpret_cast<const char*>(slab_sm) + (slot(0, it) - sm0));
    const T* s1 = reinterpret_cast<const T*>(reinterpret_cast<const char*>(slab_sm) + (slot(1, it) - sm0));
    Vec8<T>::load(s0, dv);
    Vec8<T>::load(s1, av);
    if (has_b) {
      Vec8<T>::load(reinterpret_cast<const T*>(reinterpret_cast<const char*>(slab_sm) + (slot(2, it) - sm0)), bv);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = 0.f;
    }
  };
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (live) {
#pragma unroll 1
    for (int it = 0; it < nit; ++it) {
      if (rl + it * 256 >= p.P) break;
      float g1[8], g2[8], xh1[8], xh2[8], dv[8], av[8], bv[8];
      fetch(it, dv, av, bv);
      slab_masks(p, k, dv, av, bv, g1, g2, xh1, xh2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += g1[j];
        acc[1][j] += g1[j] * xh1[j];
        acc[2][j] += g2[j];
        acc[3][j] += g2[j] * xh2[j];
      }
    }
  }
  // every lane of a warp holds the same 8 channels: fold the 32 positions of the warp
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[i][j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[i][j] = v;
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][i][j] = acc[i][j];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int i = threadIdx.x >> 3, ch = threadIdx.x & 7;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)red[w][i][ch];   // fixed order: deterministic
    coef[i][ch] = (float)(t / M);
    const int cc = blockIdx.x * 8 + ch;
    if (cc < p.C) {
      float* dst = i == 0 ? dbeta1 : (i == 1 ? dgamma1 : (i == 2 ? dbeta2 : dgamma2));
      if (dst) dst[cc] += (float)t;
    }
  }
  __syncthreads();
  if (!live || (!p.da && !p.db)) return;
  T* da = reinterpret_cast<T*>(p.da);
  T* db = reinterpret_cast<T*>(p.db);
  float c0[8], c1[8], c2[8], c3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    c0[j] = coef[0][j]; c1[j] = coef[1][j];
    c2[j] = coef[2][j]; c3[j] = coef[3][j];
  }
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    const long long pos = rl + it * 256;
    if (pos >= p.P) break;
    const long long e = pos * p.C + c;
    float g1[8], g2[8], xh1[8], xh2[8], dv[8], av[8], bv[8];
    fetch(it, dv, av, bv);
    slab_masks(p, k, dv, av, bv, g1, g2, xh1, xh2);
    if (da) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p.batch_stats1 ? k.s1[j] * (g1[j] - c0[j] - xh1[j] * c1[j]) : k.s1[j] * g1[j];
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
      for (int j = 0; j < 8; ++j) o[j] = p.batch_stats2 ? k.s2[j] * (g2[j] - c2[j] - xh2[j] * c3[j]) : k.s2[j] * g2[j];
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

constexpr long long COOP_MAX_ELEMS = 4ll << 20;   // above this the three-launch form streams better (more blocks per SM)

bool slab_enabled() {   // SAP3D_BN_BWD_SLAB=0: the cooperative form for every backbone tensor (r01 behaviour, for A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAP3D_BN_BWD_SLAB");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int ew_grid(long long nvec) {
  long long b = (nvec + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int sap3d_bn_finalize(const float* stats, int32_t rows, int32_t C, double count, const float* gamma, const float* beta,
                      float* moving_mean, float* moving_var, int32_t training, float momentum, float eps, float* scale,
                      float* shift, float* save_mean, float* save_rstd, void* stream) {
  if (require_device()) return 1;
  if (!scale || !shift) return set_error("bn_finalize: NULL output");
  if (training && !stats) return set_error("bn_finalize: training mode needs statistics");
  if (!training && (!moving_mean || !moving_var)) return set_error("bn_finalize: inference mode needs moving statistics");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (training && rows >= 256)
    launch_k(bn_finalize_kernel<8>, dim3((C + 7) / 8), dim3(8, 128), 0, st, 1, stats, rows, C, count, gamma, beta, moving_mean, moving_var, training,
             momentum, eps, scale, shift, save_mean, save_rstd);
  else
    launch_k(bn_finalize_kernel<32>, dim3((C + 31) / 32), dim3(32, 32), 0, st, 1, stats, rows, C, count, gamma, beta, moving_mean, moving_var, training,
             momentum, eps, scale, shift, save_mean, save_rstd);
  return check_launch("bn_finalize");
}

int sap3d_gn_stats(int32_t dtype, const void* x, int32_t N, int64_t S, int32_t C, int32_t G, const float* gamma,
                   const float* beta, float eps, float* scale, float* shift, float* save_mean, float* save_rstd, void* stream) {
  if (require_device()) return 1;
  if (C % G != 0) return set_error("gn_stats: C %% G != 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    gn_stats_kernel<bf16><<<N * G, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), S, C, G, gamma, beta, eps, scale, shift, save_mean, save_rstd);
  else
    gn_stats_kernel<float><<<N * G, 256, 0, st>>>(reinterpret_cast<const float*>(x), S, C, G, gamma, beta, eps, scale, shift, save_mean, save_rstd);
  return check_launch("gn_stats");
}

int sap3d_affine_act(int32_t dtype, const void* a, const float* s1, const float* t1, int32_t relu1, const void* b,
                     const float* s2, const float* t2, int32_t relu2, int32_t relu_out, void* y, int64_t P, int32_t C,
                     int64_t positions_per_sample, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("affine_act: C must be a multiple of 8 (got %d)", C);
  ApplyArgs p;
  p.a = a; p.s1 = s1; p.t1 = t1; p.b = b; p.s2 = s2; p.t2 = t2; p.y = y;
  p.P = P; p.psp = positions_per_sample; p.C = C;
  p.relu1 = relu1; p.relu2 = relu2; p.relu_out = relu_out;
  const long long nvec = P * C / 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16) launch_k(apply_kernel<bf16>, dim3(ew_grid(nvec)), dim3(256), 0, st, 1, p);
  else launch_k(apply_kernel<float>, dim3(ew_grid(nvec)), dim3(256), 0, st, 1, p);
  return check_launch("affine_act");
}

int sap3d_sample_norm_apply_supported(int32_t dtype, int64_t S, int32_t C) {
  const int esz = dtype == SAP3D_BF16 ? 2 : 4;
  // one block walks its (clip, 64-channel) slab twice: worth it while the slab is small (the backbone's stages 2-3; at S = 6272 the
  // serial walk of one block per slab was slower than the three parallel launches, measured)
  return (C % 4 == 0 && S > 0 && S * 64 * esz <= (128ll << 10)) ? 1 : 0;
}

int sap3d_sample_norm_apply(int32_t dtype, const void* a, const float* gamma1, const float* beta1, int32_t relu1, const void* b,
                            const float* gamma2, const float* beta2, int32_t relu2, int32_t relu_out, void* y, int32_t N, int64_t S,
                            int32_t C, float eps, void* stream) {
  if (require_device()) return 1;
  if (!a || !y) return set_error("sample_norm_apply: NULL tensor");
  if (!sap3d_sample_norm_apply_supported(dtype, S, C)) return set_error("sample_norm_apply: shape not supported (S = %lld, C = %d)", (long long)S, C);
  if ((gamma2 || beta2) && !b) return set_error("sample_norm_apply: a second norm needs a second operand");
  SampleNormArgs p;
  p.a = a; p.g1 = gamma1; p.b1 = beta1; p.b = b; p.g2 = gamma2; p.b2 = beta2; p.y = y;
  p.S = S; p.C = C; p.eps = eps; p.relu1 = relu1; p.relu2 = relu2; p.relu_out = relu_out;
  const dim3 grid((unsigned)N, (unsigned)((C + 63) / 64));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16) launch_k(sample_norm_apply_kernel<bf16>, grid, dim3(256), 0, st, 1, p);
  else launch_k(sample_norm_apply_kernel<float>, grid, dim3(256), 0, st, 1, p);
  return check_launch("sample_norm_apply");
}

int sap3d_bn_apply_fused(int32_t dtype, const void* a, const float* stats1, int32_t rows1, const float* gamma1, const float* beta1,
                         float* mm1, float* mv1, int32_t training1, float* scale1, float* shift1, float* mean1, float* rstd1,
                         int32_t relu1, const void* b, int32_t has_norm2, const float* stats2, int32_t rows2, const float* gamma2,
                         const float* beta2, float* mm2, float* mv2, int32_t training2, float* scale2, float* shift2, float* mean2,
                         float* rstd2, int32_t relu2, int32_t relu_out, void* y, int64_t P, int32_t C, double count, float momentum,
                         float eps, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("bn_apply_fused: C must be a multiple of 8 (got %d)", C);
  if (!a || !y || !scale1 || !shift1) return set_error("bn_apply_fused: NULL argument");
  if (training1 ? !stats1 : (!mm1 || !mv1)) return set_error("bn_apply_fused: norm 1 needs statistics");
  if (has_norm2 && (!b || !scale2 || !shift2 || (training2 ? !stats2 : (!mm2 || !mv2)))) return set_error("bn_apply_fused: norm 2 needs statistics");
  FusedApplyArgs p;
  memset(&p, 0, sizeof(p));
  p.ap.a = a; p.ap.b = b; p.ap.y = y; p.ap.P = P; p.ap.C = C; p.ap.psp = 0;
  p.ap.relu1 = relu1; p.ap.relu2 = relu2; p.ap.relu_out = relu_out;
  p.has[0] = 1; p.has[1] = has_norm2;
  p.n[0] = FusedNorm{stats1, rows1, count, gamma1, beta1, mm1, mv1, training1, scale1, shift1, mean1, rstd1};
  p.n[1] = FusedNorm{stats2, rows2, count, gamma2, beta2, mm2, mv2, training2, scale2, shift2, mean2, rstd2};
  p.momentum = momentum; p.eps = eps;
  const int chunks = (C + 63) / 64;
  long long prows = (4 * 148 + chunks - 1) / chunks;
  const long long max_rows = (P + 31) / 32;
  if (prows > max_rows) prows = max_rows;
  if (prows < 1) prows = 1;
  p.prows = (int)prows;
  dim3 grid((unsigned)prows, (unsigned)chunks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16) launch_k(bn_apply_fused_kernel<bf16>, grid, dim3(256), 0, st, 1, p);
  else launch_k(bn_apply_fused_kernel<float>, grid, dim3(256), 0, st, 1, p);
  return check_launch("bn_apply_fused");
}

size_t sap3d_affine_act_bwd_workspace(int32_t C) { return (size_t)(296 * 4 + 4 + 8) * (size_t)C * sizeof(float) + 64; }

// phase 0: the whole backward.  phase 1: reductions only (coef = LOCAL sums / count at workspace[0 .. 4C)); phase 2: apply
// only, reading coef from the workspace (the caller has summed it over the replicas in between: synchronised BatchNorm).
// SAP3D_NOB_INFLIGHT=4|8: positions per thread kept in flight by the big-tensor BatchNorm-backward kernels (bf16)
static int nob_inflight() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("SAP3D_NOB_INFLIGHT");
    v = (e != nullptr && e[0] == '4') ? 4 : 8;
  }
  return v;
}

static int affine_act_bwd_impl(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                               const float* rstd1, int32_t relu1, const void* b, const float* s2, const float* t2,
                               const float* mean2, const float* rstd2, int32_t relu2, int32_t relu_out, int64_t P, int32_t C,
                               void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1, float* dbeta1, float* dgamma2,
                               float* dbeta2, void* workspace, void* stream, double count, int phase) {
  if (require_device()) return 1;
  if (C % 8 != 0 || C > 2048) return set_error("affine_act_bwd: C must be a multiple of 8 and <= 2048 (got %d)", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ApplyBwdArgs p;
  memset(&p, 0, sizeof(p));
  p.dy = dy; p.a = a; p.s1 = s1; p.t1 = t1; p.mean1 = mean1; p.rstd1 = rstd1;
  p.b = b; p.s2 = s2; p.t2 = t2; p.mean2 = mean2; p.rstd2 = rstd2;
  p.P = P; p.psp = 0; p.C = C; p.G = 0;
  p.relu1 = relu1; p.relu2 = relu2; p.relu_out = relu_out;
  p.batch_stats1 = mean1 != nullptr; p.batch_stats2 = mean2 != nullptr;
  p.da = da; p.db = db; p.acc_a = acc_a; p.acc_b = acc_b;
  float* ws = reinterpret_cast<float*>(workspace);
  const bool need_reduce = p.batch_stats1 || p.batch_stats2 || dgamma1 || dbeta1 || dgamma2 || dbeta2;
  const bool nob = b == nullptr && db == nullptr && !p.batch_stats2;   // y = relu?(norm(a)): the register-resident kernels
  if (phase == 2) {
    if (!workspace) return set_error("affine_act_bwd: workspace required");
    p.coef = ws;
  } else if (need_reduce) {
    if (!workspace) return set_error("affine_act_bwd: workspace required");
    const int chunks = (C + 63) / 64;
    long long rows = (2 * 148) / chunks;   // floor: rows x chunks must fit the 2 x 148 co-resident blocks of the cooperative form
    const long long max_rows = (P + 31) / 32;   // (ceil put C = 1024 at 19 x 16 = 304 blocks: every stage-3 block tail fell back to 3 launches)
    if (rows > max_rows) rows = max_rows;
    if (rows > 296) rows = 296;
    if (rows < 1) rows = 1;
    p.rows = (int)rows;
    p.partial = ws + 4 * C;
    p.totals = reinterpret_cast<double*>(ws + (size_t)(296 * 4 + 4) * C + (((size_t)(296 * 4 + 4) * C) & 1));   // 8-byte aligned tail
    dim3 rgrid((unsigned)rows, (unsigned)chunks);
    // few-row tensors (stage 3): one ordinary launch, every block owns 8 channels for all positions
    if (phase == 0 && P <= SLAB_MAX_P && slab_enabled()) {
      double M = count;
      const size_t esz = dtype == SAP3D_BF16 ? 2 : 4;
      const size_t smem = (size_t)3 * ((P + 255) / 256) * 256 * 8 * esz;
      static bool attr_done[2] = {false, false};
      bool& done = attr_done[dtype == SAP3D_BF16 ? 0 : 1];
      if (!done) {
        const void* fn = dtype == SAP3D_BF16 ? (const void*)bn_bwd_slab_kernel<bf16> : (const void*)bn_bwd_slab_kernel<float>;
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)3 * (SLAB_MAX_P / 256) * 256 * 8 * esz)) != cudaSuccess)
          return set_error("affine_act_bwd slab: cudaFuncSetAttribute failed");
        done = true;
      }
      cudaError_t e = dtype == SAP3D_BF16
                          ? launch_k(bn_bwd_slab_kernel<bf16>, dim3((unsigned)((C + 7) / 8)), dim3(256), smem, st, 1, p, M, dgamma1, dbeta1, dgamma2, dbeta2)
                          : launch_k(bn_bwd_slab_kernel<float>, dim3((unsigned)((C + 7) / 8)), dim3(256), smem, st, 1, p, M, dgamma1, dbeta1, dgamma2, dbeta2);
      if (e != cudaSuccess) return set_error("affine_act_bwd slab launch: %s", cudaGetErrorString(e));
      return 0;
    }
    // backbone-sized tensors: ONE cooperative launch (reduce -> grid.sync -> finalize -> apply) instead of three
    if (phase == 0 && (long long)P * C <= COOP_MAX_ELEMS) {
      static int max_blocks[2] = {0, 0};
      int& mb = max_blocks[dtype == SAP3D_BF16 ? 0 : 1];
      if (mb == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        const void* fn = dtype == SAP3D_BF16 ? (const void*)bn_bwd_coop_kernel<bf16> : (const void*)bn_bwd_coop_kernel<float>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0) != cudaSuccess) per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        mb = per_sm * sms;
        if (mb < 1) mb = 1;
      }
      if ((long long)rows * chunks <= mb) {
        double M = count;
        void* args[] = {(void*)&p, (void*)&M, (void*)&dgamma1, (void*)&dbeta1, (void*)&dgamma2, (void*)&dbeta2};
        const void* fn = dtype == SAP3D_BF16 ? (const void*)bn_bwd_coop_kernel<bf16> : (const void*)bn_bwd_coop_kernel<float>;
        cudaError_t e = cudaLaunchCooperativeKernel(fn, rgrid, dim3(256), args, 0, st);
        if (e != cudaSuccess) return set_error("affine_act_bwd cooperative launch: %s", cudaGetErrorString(e));
        return 0;
      }
    }
    if (nob) {   // no second operand: constants in registers, four positions in flight, one wave of three blocks per SM
      long long r3 = (3 * 148 + chunks - 1) / chunks;
      if (r3 > (P + 63) / 64) r3 = (P + 63) / 64;
      if (r3 > 296) r3 = 296;
      if (r3 < 1) r3 = 1;
      rows = r3;
      p.rows = (int)rows;
      dim3 g3((unsigned)rows, (unsigned)chunks);
      if (dtype == SAP3D_BF16) {
        if (nob_inflight() == 8) launch_k(apply_bwd_reduce_nob_kernel<bf16, 8>, g3, dim3(256), 0, st, 1, p);
        else launch_k(apply_bwd_reduce_nob_kernel<bf16, 4>, g3, dim3(256), 0, st, 1, p);
      } else launch_k(apply_bwd_reduce_nob_kernel<float, 4>, g3, dim3(256), 0, st, 1, p);
    } else if (dtype == SAP3D_BF16) launch_k(apply_bwd_reduce_kernel<bf16>, rgrid, dim3(256), 0, st, 1, p);
    else launch_k(apply_bwd_reduce_kernel<float>, rgrid, dim3(256), 0, st, 1, p);
    if (check_launch("affine_act_bwd reduce")) return 1;
    launch_k(apply_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(32, 32), 0, st, 1, p.partial, (int)rows, C, count, ws, dgamma1, dbeta1, dgamma2, dbeta2);
    if (check_launch("affine_act_bwd finalize")) return 1;
    p.coef = ws;
  }
  if (phase != 1 && (da || db)) {
    const long long nvec = P * C / 8;
    if (nob && da) {
      const int chunks = (C + 63) / 64;
      long long slabs = (6 * 148 + chunks - 1) / chunks;           // two waves of three blocks per SM
      const long long max_slabs = (P + 63) / 64;
      if (slabs > max_slabs) slabs = max_slabs;
      if (slabs < 1) slabs = 1;
      dim3 ag((unsigned)slabs, (unsigned)chunks);
      if (dtype == SAP3D_BF16) {
        if (nob_inflight() == 8) {
          if (acc_a) launch_k(apply_bwd_nob_kernel<bf16, true, 8>, ag, dim3(256), 0, st, 1, p);
          else launch_k(apply_bwd_nob_kernel<bf16, false, 8>, ag, dim3(256), 0, st, 1, p);
        } else {
          if (acc_a) launch_k(apply_bwd_nob_kernel<bf16, true, 4>, ag, dim3(256), 0, st, 1, p);
          else launch_k(apply_bwd_nob_kernel<bf16, false, 4>, ag, dim3(256), 0, st, 1, p);
        }
      } else {
        if (acc_a) launch_k(apply_bwd_nob_kernel<float, true, 4>, ag, dim3(256), 0, st, 1, p);
        else launch_k(apply_bwd_nob_kernel<float, false, 4>, ag, dim3(256), 0, st, 1, p);
      }
    } else if (dtype == SAP3D_BF16) launch_k(apply_bwd_kernel<bf16>, dim3(ew_grid(nvec)), dim3(256), 0, st, 1, p);
    else launch_k(apply_bwd_kernel<float>, dim3(ew_grid(nvec)), dim3(256), 0, st, 1, p);
    if (check_launch("affine_act_bwd apply")) return 1;
  }
  return 0;
}

int sap3d_affine_act_bwd(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                         const float* rstd1, int32_t relu1, const void* b, const float* s2, const float* t2,
                         const float* mean2, const float* rstd2, int32_t relu2, int32_t relu_out, int64_t P, int32_t C,
                         void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1, float* dbeta1, float* dgamma2,
                         float* dbeta2, void* workspace, void* stream) {
  return affine_act_bwd_impl(dtype, dy, a, s1, t1, mean1, rstd1, relu1, b, s2, t2, mean2, rstd2, relu2, relu_out, P, C, da, acc_a,
                             db, acc_b, dgamma1, dbeta1, dgamma2, dbeta2, workspace, stream, (double)P, 0);
}

int sap3d_affine_act_bwd_sync(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                              const float* rstd1, int32_t relu1, const void* b, const float* s2, const float* t2,
                              const float* mean2, const float* rstd2, int32_t relu2, int32_t relu_out, int64_t P, int32_t C,
                              void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1, float* dbeta1, float* dgamma2,
                              float* dbeta2, void* workspace, void* stream, double global_count, int32_t phase) {
  if (phase != 1 && phase != 2) return set_error("affine_act_bwd_sync: phase must be 1 (reduce) or 2 (apply)");
  if (!(global_count >= (double)P)) return set_error("affine_act_bwd_sync: global_count is smaller than the local position count");
  return affine_act_bwd_impl(dtype, dy, a, s1, t1, mean1, rstd1, relu1, b, s2, t2, mean2, rstd2, relu2, relu_out, P, C, da, acc_a,
                             db, acc_b, dgamma1, dbeta1, dgamma2, dbeta2, workspace, stream, global_count, phase);
}

}  // extern "C"
