// Saliency decoder head and training-step tail kernels:
//   * final transposed conv (Cout = 1, k3 s2) fused with sigmoid     (p3d.py:393,397)
//   * sigmoid + smooth-L1 (sum) loss and its gradient                 (utils/network.py:49-62; train.py:159)
//   * inverted dropout with a counter-based mask                      (p3d.py:392)
//   * attention gate  y = o * gamma + x                               (utils/network.py:191-192)
//   * fused flat Adam (TF formula), step counter, dtype casts         (train.py:168)
// All bandwidth-bound; warp-shuffle reductions, vectorised accesses.
#include <string.h>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

// ------------------------------------------------------------------------------------------------
// head: logits[n,od,oh,ow] = bias + sum_{k,c} x[n,(o-k)/2,c] * w[k][c]   (pb = 0 for k3 s2)
// one warp per output voxel, lanes over channels, weights in shared memory
// ------------------------------------------------------------------------------------------------
struct HeadArgs {
  const void* x; const float* w; const float* bias;
  float* logits; float* pred;
  int N, D, H, W, C, kd, kh, kw, s;
};

template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadArgs p) {
  extern __shared__ float sw[];  // [taps][C]
  const int taps = p.kd * p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.C; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Do = p.D * p.s, Ho = p.H * p.s, Wo = p.W * p.s;
  const int pbd = max(p.kd - p.s, 0) / 2, pbh = max(p.kh - p.s, 0) / 2, pbw = max(p.kw - p.s, 0) / 2;
  const long long total = (long long)p.N * Do * Ho * Wo;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const T* x = reinterpret_cast<const T*>(p.x);
  const float b0 = p.bias ? p.bias[0] : 0.f;
  for (long long o = warp_id; o < total; o += nwarps) {
    long long r = o;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho); r /= Ho;
    const int od = (int)(r % Do);
    const int n = (int)(r / Do);
    float acc = 0.f;
    for (int a = 0; a < p.kd; ++a) {
      const int nd = od + pbd - a;
      if (nd < 0 || nd % p.s != 0 || nd / p.s >= p.D) continue;
      for (int b = 0; b < p.kh; ++b) {
        const int nh = oh + pbh - b;
        if (nh < 0 || nh % p.s != 0 || nh / p.s >= p.H) continue;
        for (int e = 0; e < p.kw; ++e) {
          const int nw = ow + pbw - e;
          if (nw < 0 || nw % p.s != 0 || nw / p.s >= p.W) continue;
          const T* xp = x + ((((long long)n * p.D + nd / p.s) * p.H + nh / p.s) * p.W + nw / p.s) * p.C;
          const float* wp = sw + ((a * p.kh + b) * p.kw + e) * p.C;
          for (int c = lane * 4; c < p.C; c += 128) {
            float v0, v1, v2, v3;
            if (sizeof(T) == 2) {
              const uint2 u = *reinterpret_cast<const uint2*>(xp + c);
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y);
              v0 = f0.x; v1 = f0.y; v2 = f1.x; v3 = f1.y;
            } else {
              const float4 f = *reinterpret_cast<const float4*>(xp + c);
              v0 = f.x; v1 = f.y; v2 = f.z; v3 = f.w;
            }
            acc = fmaf(v0, wp[c], acc);
            acc = fmaf(v1, wp[c + 1], acc);
            acc = fmaf(v2, wp[c + 2], acc);
            acc = fmaf(v3, wp[c + 3], acc);
          }
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float z = acc + b0;
      p.logits[o] = z;
      if (p.pred) p.pred[o] = 1.f / (1.f + __expf(-z));
    }
  }
}

// k = 3, s = 2 specialisation: one warp per INPUT anchor voxel j produces the 2x2x2 output cube p = 2j + r.
// Output parity r = 0 uses taps k = 0 (input j) and k = 2 (input j-1), r = 1 uses k = 1 (input j): the cube
// needs the 8 inputs j - {0,1}^3, each loaded ONCE per lane (3.4x fewer loads than the gather form).
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_k3s2_kernel(const HeadArgs p) {
  extern __shared__ float sw[];  // [27][C]
  for (int i = threadIdx.x; i < 27 * p.C; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Do = p.D * 2, Ho = p.H * 2, Wo = p.W * 2;
  const long long total = (long long)p.N * p.D * p.H * p.W;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const T* x = reinterpret_cast<const T*>(p.x);
  const float b0 = p.bias ? p.bias[0] : 0.f;
  for (long long a = warp_id; a < total; a += nwarps) {
    long long r = a;
    const int jw = (int)(r % p.W); r /= p.W;
    const int jh = (int)(r % p.H); r /= p.H;
    const int jd = (int)(r % p.D);
    const int n = (int)(r / p.D);
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    for (int c = lane * 4; c < p.C; c += 128) {
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        if (jd - dd < 0) continue;
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
          if (jh - dh < 0) continue;
#pragma unroll
          for (int dw = 0; dw < 2; ++dw) {
            if (jw - dw < 0) continue;
            const T* xp = x + ((((long long)n * p.D + jd - dd) * p.H + jh - dh) * p.W + jw - dw) * p.C + c;
            float v0, v1, v2, v3;
            if (sizeof(T) == 2) {
              const uint2 u = *reinterpret_cast<const uint2*>(xp);
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y);
              v0 = f0.x; v1 = f0.y; v2 = f1.x; v3 = f1.y;
            } else {
              const float4 f = *reinterpret_cast<const float4*>(xp);
              v0 = f.x; v1 = f.y; v2 = f.z; v3 = f.w;
            }
            // input offset 0 feeds parity 0 (tap 0) and parity 1 (tap 1); offset 1 feeds parity 0 only (tap 2)
#pragma unroll
            for (int rd = 0; rd < 2; ++rd) {
              if (dd == 1 && rd == 1) continue;
              const int kd = dd == 1 ? 2 : rd;
#pragma unroll
              for (int rh = 0; rh < 2; ++rh) {
                if (dh == 1 && rh == 1) continue;
                const int kh = dh == 1 ? 2 : rh;
#pragma unroll
                for (int rw = 0; rw < 2; ++rw) {
                  if (dw == 1 && rw == 1) continue;
                  const int kw = dw == 1 ? 2 : rw;
                  const float* wp = sw + ((kd * 3 + kh) * 3 + kw) * p.C + c;
                  float t = acc[(rd * 2 + rh) * 2 + rw];
                  t = fmaf(v0, wp[0], t);
                  t = fmaf(v1, wp[1], t);
                  t = fmaf(v2, wp[2], t);
                  t = fmaf(v3, wp[3], t);
                  acc[(rd * 2 + rh) * 2 + rw] = t;
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = warp_sum(acc[o]);
    if (lane < 8) {
      float z = acc[0];
#pragma unroll
      for (int o = 1; o < 8; ++o) z = lane == o ? acc[o] : z;
      z += b0;
      const int rd = lane >> 2, rh = (lane >> 1) & 1, rw = lane & 1;
      const long long o = (((long long)n * Do + 2 * jd + rd) * Ho + 2 * jh + rh) * Wo + 2 * jw + rw;
      p.logits[o] = z;
      if (p.pred) p.pred[o] = 1.f / (1.f + __expf(-z));
    }
  }
}


// ------------------------------------------------------------------------------------------------
// tensor-core head (k3 s2, C % 64 == 0, bf16): the Cout = 1 transposed conv is a [positions x C] x [C x 27] GEMM
// followed by a col2im gather; its gradients are GEMMs over the im2col of dlogits (run through sap3d_gemm_nt/_tn).
// ------------------------------------------------------------------------------------------------
// w [27][C] fp32 -> wf [32][C] bf16 (rows >= 27 zero) and wd [C][64] bf16 (wd[c][k] = w[k][c], k >= 27 zero)
__global__ void __launch_bounds__(256) head_pack_w_kernel(const float* __restrict__ w, int C, bf16* __restrict__ wf, bf16* __restrict__ wd) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 32 * C; i += gridDim.x * blockDim.x) {
    const int k = i / C;
    wf[i] = __float2bfloat16_rn(k < 27 ? w[i] : 0.f);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 64 * C; i += gridDim.x * blockDim.x) {
    const int c = i / 64, k = i % 64;
    wd[i] = __float2bfloat16_rn(k < 27 ? w[k * C + c] : 0.f);
  }
}

// (input index, tap) pairs of a k3 s2 'same' transposed conv that land on output index o: 2i + k = o
SAP3D_DEVINL int head_contrib(int o, int (&i)[2], int (&k)[2]) {
  if (o & 1) { i[0] = o >> 1; k[0] = 1; return 1; }
  i[0] = o >> 1; k[0] = 0;
  i[1] = (o >> 1) - 1; k[1] = 2;
  return (o >> 1) >= 1 ? 2 : 1;
}

// logits[n, o] = bias + sum over (i, k) with 2i + k = o of t[i][k]   (per dim: o even -> (o/2, 0), (o/2 - 1, 2); o odd -> ((o-1)/2, 1))
__global__ void __launch_bounds__(256) head_col2im_kernel(const float* __restrict__ t, int N, int D, int H, int W, const float* bias,
                                                           float* __restrict__ logits, float* __restrict__ pred) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * Do * Ho * Wo;
  const float b0 = bias ? bias[0] : 0.f;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    long long r = o;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho); r /= Ho;
    const int od = (int)(r % Do);
    const int n = (int)(r / Do);
    int id[2], kd[2], ih[2], kh[2], iw[2], kw[2];
    const int nd = head_contrib(od, id, kd), nh = head_contrib(oh, ih, kh), nw = head_contrib(ow, iw, kw);
    float acc = b0;
    for (int a = 0; a < nd; ++a)
      for (int b = 0; b < nh; ++b)
        for (int e = 0; e < nw; ++e)
          acc += __ldg(t + ((((long long)n * D + id[a]) * H + ih[b]) * W + iw[e]) * 32 + (kd[a] * 3 + kh[b]) * 3 + kw[e]);
    logits[o] = acc;
    if (pred) pred[o] = 1.f / (1.f + __expf(-acc));
  }
}

// col[pos][k] = dlogits[n, 2i + k] (k < 27; 0 beyond the output extent and for the 37 padding columns), bf16 [positions][64]
__global__ void __launch_bounds__(256) head_im2col_kernel(const float* __restrict__ dlog, int N, int D, int H, int W, bf16* __restrict__ col) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * D * H * W * 8;   // 8 threads per position, 8 columns each
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int part = (int)(i & 7);
    long long r = i >> 3;
    const long long pos = r;
    const int iw = (int)(r % W); r /= W;
    const int ih = (int)(r % H); r /= H;
    const int id = (int)(r % D);
    const int n = (int)(r / D);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = part * 8 + j;
      float x = 0.f;
      if (k < 27) {
        const int od = 2 * id + k / 9, oh = 2 * ih + (k / 3) % 3, ow = 2 * iw + k % 3;
        if (od < Do && oh < Ho && ow < Wo) x = __ldg(dlog + (((long long)n * Do + od) * Ho + oh) * Wo + ow);
      }
      v[j] = x;
    }
    Vec8<bf16>::store(col + pos * 64 + part * 8, v);
  }
}

__global__ void __launch_bounds__(256) head_wgrad_fold_kernel(const float* __restrict__ d64, int C, float* dw) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 27 * C; i += gridDim.x * blockDim.x) dw[i] += d64[i];
}

// dx[n,i,c] (+)= sum_k dlog[n, s*i + k - pb] * w[k][c] ; thread = (voxel, 8 channels)
struct HeadBwdArgs {
  const float* dlog; const void* x; const float* w; void* dx; float* dw;
  int N, D, H, W, C, kd, kh, kw, s, accumulate;
};

template <typename T>
__global__ void __launch_bounds__(256) head_dgrad_kernel(const HeadBwdArgs p) {
  extern __shared__ float sw[];
  const int taps = p.kd * p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.C; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  const int cv = p.C / 8;
  const int Do = p.D * p.s, Ho = p.H * p.s, Wo = p.W * p.s;
  const int pbd = max(p.kd - p.s, 0) / 2, pbh = max(p.kh - p.s, 0) / 2, pbw = max(p.kw - p.s, 0) / 2;
  const long long total = (long long)p.N * p.D * p.H * p.W * cv;
  T* dx = reinterpret_cast<T*>(p.dx);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int iw = (int)(r % p.W); r /= p.W;
    const int ih = (int)(r % p.H); r /= p.H;
    const int id = (int)(r % p.D);
    const int n = (int)(r / p.D);
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    for (int a = 0; a < p.kd; ++a) {
      const int od = id * p.s + a - pbd;
      if (od < 0 || od >= Do) continue;
      for (int b = 0; b < p.kh; ++b) {
        const int oh = ih * p.s + b - pbh;
        if (oh < 0 || oh >= Ho) continue;
        for (int e = 0; e < p.kw; ++e) {
          const int ow = iw * p.s + e - pbw;
          if (ow < 0 || ow >= Wo) continue;
          const float d = __ldg(p.dlog + (((long long)n * Do + od) * Ho + oh) * Wo + ow);
          const float* wp = sw + ((a * p.kh + b) * p.kw + e) * p.C + c;
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = fmaf(d, wp[j], g[j]);
        }
      }
    }
    const long long off = i * 8;
    if (p.accumulate) {
      float old[8];
      Vec8<T>::load(dx + off, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += old[j];
    }
    Vec8<T>::store(dx + off, g);
  }
}

// dw[k][c] += sum_i dlog[s*i + k - pb] * x[i][c]; grid = (voxel slabs, kd); thread = (c-vec, voxel lane)
template <typename T>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const HeadBwdArgs p) {
  extern __shared__ float red[];  // [lanes][kh*kw][C]
  const int a = blockIdx.y;       // tap index along d
  const int cv = p.C / 8;
  const int lanes = blockDim.x / cv;
  const int cl = threadIdx.x % cv, vl = threadIdx.x / cv;
  const int c = cl * 8;
  const int Do = p.D * p.s, Ho = p.H * p.s, Wo = p.W * p.s;
  const int pbd = max(p.kd - p.s, 0) / 2, pbh = max(p.kh - p.s, 0) / 2, pbw = max(p.kw - p.s, 0) / 2;
  const long long V = (long long)p.N * p.D * p.H * p.W;
  const long long per = (V + gridDim.x - 1) / gridDim.x;
  const long long vbeg = blockIdx.x * per, vend = vbeg + per < V ? vbeg + per : V;
  const T* x = reinterpret_cast<const T*>(p.x);
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int khw = p.kh * p.kw;  // <= 9
  if (vl < lanes) {
    for (long long v = vbeg + vl; v < vend; v += lanes) {
      long long r = v;
      const int iw = (int)(r % p.W); r /= p.W;
      const int ih = (int)(r % p.H); r /= p.H;
      const int id = (int)(r % p.D);
      const int n = (int)(r / p.D);
      const int od = id * p.s + a - pbd;
      if (od < 0 || od >= Do) continue;
      float xv[8];
      Vec8<T>::load(x + v * p.C + c, xv);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t < khw) {
          const int b = t / p.kw, e = t % p.kw;
          const int oh = ih * p.s + b - pbh, ow = iw * p.s + e - pbw;
          if (oh >= 0 && oh < Ho && ow >= 0 && ow < Wo) {
            const float d = __ldg(p.dlog + (((long long)n * Do + od) * Ho + oh) * Wo + ow);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(d, xv[j], acc[t][j]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (t < khw)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(vl * khw + t) * p.C + c + j] = acc[t][j];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < khw * p.C; idx += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * khw * p.C + idx];
    atomicAdd(p.dw + (long long)a * khw * p.C + idx, s);
  }
}

// ------------------------------------------------------------------------------------------------
// sigmoid + smooth-L1 (sigma = 1) sum loss; writes pred, dlogits; accumulates loss (double) and sum dlogits
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                    long long n, int apply_sigmoid, float* pred, float* dlogits,
                                                    double* loss_sum, float* dbias, float sigma2, float w_in, float w_out) {
  double ls = 0.0;
  float ds = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float z = logits[i];
    const float pr = apply_sigmoid ? 1.f / (1.f + __expf(-z)) : z;
    if (pred) pred[i] = pr;
    // utils/network.py:49-62: in = w_in * (pred - target); |in| < 1/sigma^2 ? in^2 sigma^2 / 2 : |in| - 0.5/sigma^2; times w_out
    const float d = w_in * (pr - target[i]);
    const float a = fabsf(d);
    const bool quad = a < 1.f / sigma2;
    ls += w_out * (quad ? 0.5f * sigma2 * d * d : a - 0.5f / sigma2);
    float g = w_out * w_in * (quad ? sigma2 * d : (d > 0.f ? 1.f : -1.f));
    if (apply_sigmoid) g *= pr * (1.f - pr);
    if (dlogits) dlogits[i] = g;
    ds += g;
  }
  ls = warp_sum(ls);
  ds = warp_sum(ds);
  __shared__ double sl[8];
  __shared__ float sd[8];
  if ((threadIdx.x & 31) == 0) {
    sl[threadIdx.x >> 5] = ls;
    sd[threadIdx.x >> 5] = ds;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tl = 0.0;
    float td = 0.f;
    for (int i = 0; i < 8; ++i) {
      tl += sl[i];
      td += sd[i];
    }
    atomicAdd(loss_sum, tl);
    if (dbias) atomicAdd(dbias, td);
  }
}

// ------------------------------------------------------------------------------------------------
// dropout: keep = hash(seed, index) >= rate ; y = x * keep / (1 - rate).  seed = base_seed + *step
// (splitmix64; the oracle reproduces it bit-exactly in NumPy)
// ------------------------------------------------------------------------------------------------
SAP3D_DEVINL bool dropout_keep(unsigned long long seed, unsigned long long idx, float rate) {
  unsigned long long h = idx + seed * 0x9E3779B97F4A7C15ull;
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 27; h *= 0x94D049BB133111EBull;
  h ^= h >> 31;
  const float u = (float)(h >> 40) * (1.f / 16777216.f);
  return u >= rate;
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float rate,
                                                       unsigned long long base_seed, const int* __restrict__ step,
                                                       int accumulate) {
  const unsigned long long seed = base_seed + (step ? (unsigned long long)step[0] : 0ull);
  const float sc = 1.f / (1.f - rate);
  const long long nvec = n / 8;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    float a[8];
    Vec8<T>::load(x + v * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = dropout_keep(seed, (unsigned long long)(v * 8 + j), rate) ? a[j] * sc : 0.f;
    if (accumulate) {
      float old[8];
      Vec8<T>::load(y + v * 8, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += old[j];
    }
    Vec8<T>::store(y + v * 8, a);
  }
}

// ------------------------------------------------------------------------------------------------
// attention gate: y = o * gamma + x
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gate_fwd_kernel(const T* __restrict__ o, const T* __restrict__ x, const float* gamma,
                                                        T* __restrict__ y, long long n) {
  const float g = gamma[0];
  const long long nvec = n / 8;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    Vec8<T>::load(o + v * 8, a);
    Vec8<T>::load(x + v * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], g, b[j]);
    Vec8<T>::store(y + v * 8, a);
  }
}

// do = dy * gamma ; dx (+)= dy ; dgamma += sum dy * o
template <typename T>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ o, const float* gamma,
                                                        T* __restrict__ d_o, T* __restrict__ dx, int acc_x, float* dgamma,
                                                        long long n) {
  const float g = gamma[0];
  const long long nvec = n / 8;
  float s = 0.f;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    float d[8], ov[8], r[8];
    Vec8<T>::load(dy + v * 8, d);
    Vec8<T>::load(o + v * 8, ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = fmaf(d[j], ov[j], s);
      r[j] = d[j] * g;
    }
    Vec8<T>::store(d_o + v * 8, r);
    if (dx) {
      if (acc_x) {
        float old[8];
        Vec8<T>::load(dx + v * 8, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += old[j];
      }
      Vec8<T>::store(dx + v * 8, d);
    }
  }
  s = warp_sum(s);
  __shared__ float ss[8];
  if ((threadIdx.x & 31) == 0) ss[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ss[i];
    atomicAdd(dgamma, t);
  }
}

// ------------------------------------------------------------------------------------------------
// Adam (TF-1.x formula): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps)
// ------------------------------------------------------------------------------------------------
template <typename G>
SAP3D_DEVINL void load_grad4(const G* g, long long i, float (&out)[4]);
template <>
SAP3D_DEVINL void load_grad4<float>(const float* g, long long i, float (&out)[4]) {
  const float4 t = reinterpret_cast<const float4*>(g)[i];
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
template <>
SAP3D_DEVINL void load_grad4<bf16>(const bf16* g, long long i, float (&out)[4]) {
  const uint2 t = reinterpret_cast<const uint2*>(g)[i];
  out[0] = __uint_as_float(t.x << 16); out[1] = __uint_as_float(t.x & 0xffff0000u);
  out[2] = __uint_as_float(t.y << 16); out[3] = __uint_as_float(t.y & 0xffff0000u);
}
// G = float: the gradient buffer of the engine; G = bf16: the all-reduced buckets of the data-parallel exchange, consumed
// directly (no pass that widens them back into the fp32 buffer)
template <typename G>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const G* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, const int* __restrict__ step, float lr,
                                                    float b1, float b2, float eps, float gscale) {
  const float t = (float)step[0];
  const float lr_t = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i];
    float gp[4];
    load_grad4<G>(g, i, gp);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* wp = &wv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = gp[j] * gscale;
      mp[j] = b1 * mp[j] + (1.f - b1) * gg;
      vp[j] = b2 * vp[j] + (1.f - b2) * gg * gg;
      wp[j] -= lr_t * mp[j] / (sqrtf(vp[j]) + eps);
    }
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gg = to_f32<G>(g[i]) * gscale;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm;
    v[i] = vv;
    w[i] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}

__global__ void step_inc_kernel(int* step) { step[0] += 1; }

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  // 8 elements per thread (two 16-byte loads, one 16-byte store) where both pointers allow it
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  const long long n8 = vec ? n / 8 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<float>::load(x + i * 8, v);
    Vec8<bf16>::store(y + i * 8, v);
  }
  for (long long i = n8 * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

int egrid(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int sap3d_head_fwd(int32_t dtype, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, const int32_t* ksize,
                   int32_t stride, const float* w, const float* bias, float* logits, float* pred, void* stream) {
  if (require_device()) return 1;
  if (C % 4 != 0) return set_error("head_fwd: C must be a multiple of 4");
  HeadArgs p;
  p.x = x; p.w = w; p.bias = bias; p.logits = logits; p.pred = pred;
  p.N = N; p.D = D; p.H = H; p.W = W; p.C = C; p.kd = ksize[0]; p.kh = ksize[1]; p.kw = ksize[2]; p.s = stride;
  const size_t smem = (size_t)p.kd * p.kh * p.kw * C * sizeof(float);
  if (smem > 48 * 1024) return set_error("head_fwd: filter too large for shared memory");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = 148 * 8;
  if (p.kd == 3 && p.kh == 3 && p.kw == 3 && stride == 2) {
    if (dtype == SAP3D_BF16) head_fwd_k3s2_kernel<bf16><<<blocks, 256, smem, st>>>(p);
    else head_fwd_k3s2_kernel<float><<<blocks, 256, smem, st>>>(p);
  } else if (dtype == SAP3D_BF16) head_fwd_kernel<bf16><<<blocks, 256, smem, st>>>(p);
  else head_fwd_kernel<float><<<blocks, 256, smem, st>>>(p);
  return check_launch("head_fwd");
}

int sap3d_head_bwd(int32_t dtype, const float* dlogits, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                   const int32_t* ksize, int32_t stride, const float* w, void* dx, int32_t accumulate, float* dw, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("head_bwd: C must be a multiple of 8");
  HeadBwdArgs p;
  p.dlog = dlogits; p.x = x; p.w = w; p.dx = dx; p.dw = dw;
  p.N = N; p.D = D; p.H = H; p.W = W; p.C = C; p.kd = ksize[0]; p.kh = ksize[1]; p.kw = ksize[2]; p.s = stride;
  p.accumulate = accumulate;
  if (p.kh * p.kw > 9) return set_error("head_bwd: kh*kw > 9 unsupported");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)p.kd * p.kh * p.kw * C * sizeof(float);
  if (smem > 48 * 1024) return set_error("head_bwd: filter too large for shared memory");
  if (dx) {
    const long long total = (long long)N * D * H * W * (C / 8);
    if (dtype == SAP3D_BF16) head_dgrad_kernel<bf16><<<egrid(total), 256, smem, st>>>(p);
    else head_dgrad_kernel<float><<<egrid(total), 256, smem, st>>>(p);
    if (check_launch("head_dgrad")) return 1;
  }
  if (dw) {
    const int cv = C / 8;
    int lanes = 256 / cv;
    if (lanes < 1) return set_error("head_bwd: C too large");
    size_t smem2 = (size_t)lanes * p.kh * p.kw * C * sizeof(float);
    while (smem2 > 48 * 1024 && lanes > 1) {
      lanes /= 2;
      smem2 = (size_t)lanes * p.kh * p.kw * C * sizeof(float);
    }
    dim3 grid(296, p.kd);
    if (dtype == SAP3D_BF16) head_wgrad_kernel<bf16><<<grid, lanes * cv, smem2, st>>>(p);
    else head_wgrad_kernel<float><<<grid, lanes * cv, smem2, st>>>(p);
    if (check_launch("head_wgrad")) return 1;
  }
  return 0;
}


/* workspace layout of the tensor-core head: wf [32][C] bf16 | wd [C][64] bf16 | d64 [64][C] f32 | t27 [P][32] f32 | col [P][64] bf16 */
static size_t head_tc_offsets(long long P, int C, size_t* o_wd, size_t* o_d64, size_t* o_t27, size_t* o_col) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  take((size_t)32 * C * 2);
  *o_wd = take((size_t)C * 64 * 2);
  *o_d64 = take((size_t)64 * C * 4);
  *o_t27 = take((size_t)P * 32 * 4);
  *o_col = take((size_t)P * 64 * 2);
  return off;
}

size_t sap3d_head_tc_workspace(int32_t N, int32_t D, int32_t H, int32_t W, int32_t C) {
  size_t a, b, c, d;
  return head_tc_offsets((long long)N * D * H * W, C, &a, &b, &c, &d);
}

int sap3d_head_tc_fwd(const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, const float* w, const float* bias,
                      float* logits, float* pred, void* workspace, void* stream) {
  if (require_device()) return 1;
  if (C % 64 != 0 || !workspace) return set_error("head_tc_fwd: needs C %% 64 == 0 and a workspace");
  const long long P = (long long)N * D * H * W;
  size_t o_wd, o_d64, o_t27, o_col;
  head_tc_offsets(P, C, &o_wd, &o_d64, &o_t27, &o_col);
  char* ws = reinterpret_cast<char*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  head_pack_w_kernel<<<32, 256, 0, st>>>(w, C, reinterpret_cast<bf16*>(ws), reinterpret_cast<bf16*>(ws + o_wd));
  if (check_launch("head_pack_w")) return 1;
  if (sap3d_gemm_nt(x, C, ws, C, 32, ws + o_t27, 32, (int32_t)P, 32, C, 1, 0, stream)) return 1;
  head_col2im_kernel<<<egrid(P * 8), 256, 0, st>>>(reinterpret_cast<const float*>(ws + o_t27), N, D, H, W, bias, logits, pred);
  return check_launch("head_col2im");
}

int sap3d_head_tc_bwd(const float* dlogits, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, void* dx,
                      int32_t accumulate, float* dw, void* workspace, void* stream, void* wgrad_stream) {
  if (require_device()) return 1;
  if (C % 64 != 0 || !workspace) return set_error("head_tc_bwd: needs C %% 64 == 0 and a workspace");
  const long long P = (long long)N * D * H * W;
  size_t o_wd, o_d64, o_t27, o_col;
  head_tc_offsets(P, C, &o_wd, &o_d64, &o_t27, &o_col);
  char* ws = reinterpret_cast<char*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  head_im2col_kernel<<<egrid(P * 8), 256, 0, st>>>(dlogits, N, D, H, W, reinterpret_cast<bf16*>(ws + o_col));
  if (check_launch("head_im2col")) return 1;
  if (dx && sap3d_gemm_nt(ws + o_col, 64, ws + o_wd, 64, C, dx, C, (int32_t)P, C, 64, 0, accumulate, stream)) return 1;
  if (dw) {
    (void)wgrad_stream;
    if (cudaMemsetAsync(ws + o_d64, 0, (size_t)64 * C * 4, st) != cudaSuccess) return set_error("head_tc_bwd: memset failed");
    if (sap3d_gemm_tn(ws + o_col, 64, x, C, reinterpret_cast<float*>(ws + o_d64), C, 64, C, (int32_t)P, stream)) return 1;
    head_wgrad_fold_kernel<<<16, 256, 0, st>>>(reinterpret_cast<const float*>(ws + o_d64), C, dw);
    if (check_launch("head_wgrad_fold")) return 1;
  }
  return 0;
}

int sap3d_loss_smooth_l1(const float* logits, const float* target, int64_t n, int32_t apply_sigmoid, float* pred,
                         float* dlogits, double* loss_sum, float* dbias, void* stream) {
  return sap3d_loss_smooth_l1_ex(logits, target, n, apply_sigmoid, pred, dlogits, loss_sum, dbias, 1.f, 1.f, 1.f, stream);
}

int sap3d_loss_smooth_l1_ex(const float* logits, const float* target, int64_t n, int32_t apply_sigmoid, float* pred,
                            float* dlogits, double* loss_sum, float* dbias, float sigma, float inside_weight, float outside_weight,
                            void* stream) {
  if (require_device()) return 1;
  if (!logits || !target || !loss_sum) return set_error("loss_smooth_l1: NULL pointer");
  if (!(sigma > 0.f)) return set_error("loss_smooth_l1: sigma must be positive");
  loss_kernel<<<egrid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, target, n, apply_sigmoid, pred, dlogits,
                                                                            loss_sum, dbias, sigma * sigma, inside_weight, outside_weight);
  return check_launch("loss_smooth_l1");
}

int sap3d_dropout(int32_t dtype, const void* x, void* y, int64_t n, float rate, uint64_t base_seed, const int32_t* step,
                  int32_t accumulate, void* stream) {
  if (require_device()) return 1;
  if (n % 8 != 0) return set_error("dropout: n must be a multiple of 8");
  if (!(rate >= 0.f && rate < 1.f)) return set_error("dropout: rate must be in [0,1)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    dropout_kernel<bf16><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n, rate, base_seed, step, accumulate);
  else
    dropout_kernel<float><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<float*>(y), n, rate, base_seed, step, accumulate);
  return check_launch("dropout");
}

int sap3d_gate_fwd(int32_t dtype, const void* o, const void* x, const float* gamma, void* y, int64_t n, void* stream) {
  if (require_device()) return 1;
  if (n % 8 != 0) return set_error("gate_fwd: n must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    gate_fwd_kernel<bf16><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(x), gamma, reinterpret_cast<bf16*>(y), n);
  else
    gate_fwd_kernel<float><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const float*>(o), reinterpret_cast<const float*>(x), gamma, reinterpret_cast<float*>(y), n);
  return check_launch("gate_fwd");
}

int sap3d_gate_bwd(int32_t dtype, const void* dy, const void* o, const float* gamma, void* d_o, void* dx, int32_t acc_x,
                   float* dgamma, int64_t n, void* stream) {
  if (require_device()) return 1;
  if (n % 8 != 0) return set_error("gate_bwd: n must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    gate_bwd_kernel<bf16><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(o), gamma, reinterpret_cast<bf16*>(d_o), reinterpret_cast<bf16*>(dx), acc_x, dgamma, n);
  else
    gate_bwd_kernel<float><<<egrid(n / 8), 256, 0, st>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(o), gamma, reinterpret_cast<float*>(d_o), reinterpret_cast<float*>(dx), acc_x, dgamma, n);
  return check_launch("gate_bwd");
}

int sap3d_adam_step(float* w, const float* g, float* m, float* v, int64_t n, const int32_t* step, float lr, float b1, float b2,
                    float eps, float grad_scale, void* stream) {
  if (require_device()) return 1;
  if (!w || !g || !m || !v || !step) return set_error("adam_step: NULL pointer");
  adam_kernel<float><<<egrid(n / 4 + 1), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, g, m, v, n, step, lr, b1, b2, eps, grad_scale);
  return check_launch("adam_step");
}

int sap3d_adam_step_g(float* w, const void* g, int32_t g_dtype, float* m, float* v, int64_t n, const int32_t* step, float lr, float b1, float b2,
                      float eps, float grad_scale, void* stream) {
  if (require_device()) return 1;
  if (!w || !g || !m || !v || !step) return set_error("adam_step_g: NULL pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_dtype == SAP3D_F32) {
    if (reinterpret_cast<uintptr_t>(g) & 15u) return set_error("adam_step_g: fp32 gradient pointer must be 16-byte aligned");
    adam_kernel<float><<<egrid(n / 4 + 1), 256, 0, st>>>(w, reinterpret_cast<const float*>(g), m, v, n, step, lr, b1, b2, eps, grad_scale);
  } else if (g_dtype == SAP3D_BF16) {
    if (reinterpret_cast<uintptr_t>(g) & 7u) return set_error("adam_step_g: bf16 gradient pointer must be 8-byte aligned");
    adam_kernel<bf16><<<egrid(n / 4 + 1), 256, 0, st>>>(w, reinterpret_cast<const bf16*>(g), m, v, n, step, lr, b1, b2, eps, grad_scale);
  } else {
    return set_error("adam_step_g: gradient dtype must be SAP3D_F32 or SAP3D_BF16");
  }
  return check_launch("adam_step_g");
}

int sap3d_step_increment(int32_t* step, void* stream) {
  if (require_device()) return 1;
  step_inc_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(step);
  return check_launch("step_increment");
}

int sap3d_cast(int32_t src_dtype, const void* src, void* dst, int64_t n, void* stream) {
  if (require_device()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_dtype == SAP3D_F32)
    cast_f32_bf16_kernel<<<egrid(n), 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst), n);
  else
    cast_bf16_f32_kernel<<<egrid(n), 256, 0, st>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<float*>(dst), n);
  return check_launch("cast");
}

}  // extern "C"

// ---- channel pad / unpad: [P][c] <-> [P][c_pad] (zero padded), used to give the attention f/g
// projections (C/8 = 16 or 32 channels) the 64-channel rows the tensor-core GEMM needs -------------
namespace {
template <typename T>
__global__ void __launch_bounds__(256) pad_channels_kernel(const T* __restrict__ in, T* __restrict__ out, long long P, int c, int cp,
                                                            int unpad, int accumulate) {
  const long long total = P * (unpad ? c : cp);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!unpad) {
      const long long pos = i / cp;
      const int ch = (int)(i - pos * cp);
      out[i] = ch < c ? in[pos * c + ch] : from_f32<T>(0.f);
    } else {
      const long long pos = i / c;
      const int ch = (int)(i - pos * c);
      float v = to_f32<T>(in[pos * cp + ch]);
      if (accumulate) v += to_f32<T>(out[i]);
      out[i] = from_f32<T>(v);
    }
  }
}
}  // namespace

extern "C" int sap3d_pad_channels(int32_t dtype, const void* in, void* out, int64_t P, int32_t c, int32_t c_pad, int32_t unpad,
                                  int32_t accumulate, void* stream) {
  if (require_device()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = P * (unpad ? c : c_pad);
  if (dtype == SAP3D_BF16)
    pad_channels_kernel<bf16><<<egrid(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), P, c, c_pad, unpad, accumulate);
  else
    pad_channels_kernel<float><<<egrid(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), P, c, c_pad, unpad, accumulate);
  return check_launch("pad_channels");
}
