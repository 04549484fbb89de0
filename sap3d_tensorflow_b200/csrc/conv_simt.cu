// CUDA-core implicit-GEMM convolution kernels (any geometry, bf16 or f32 storage, fp32 accumulate).
// Used where the tensor-core path does not apply: the 3-channel 1x7x7 stem (p3d.py:343), channel
// counts that are not multiples of 64, the fp32 parity path, and filter gradients (v1).
//
// All variants are one "gather" form:   g = (o*mul + off0 + k*offk) / div   per spatial dim
//   conv forward / transposed-conv data-gradient : mul = s, off0 = -pb, offk = +1, div = 1
//   transposed-conv forward / conv data-gradient : mul = 1, off0 = +pb, offk = -1, div = s
// (non-divisible or out-of-range g contributes zero).
#include "conv_simt.cuh"

#include <stdio.h>

#include "common.cuh"

namespace sap3d {

__device__ __forceinline__ bool gather_coord(int o, int k, int mul, int off0, int offk, int div, int extent, int& g) {
  int num = o * mul + off0 + k * offk;
  if (num < 0) return false;
  if (div != 1) {
    if (num % div != 0) return false;
    num /= div;
  }
  g = num;
  return num < extent;
}

// ---------------------------------------------------------------------------------------------
// forward-form kernel: out[pos][co] = sum_{tap,ci} G(pos,tap)[ci] * W(tap,ci,co)
// tile 64 positions x 64 output channels, 256 threads, 4x4 per thread, K chunks of 16
// ---------------------------------------------------------------------------------------------
template <typename T, typename OutT>
__global__ void __launch_bounds__(256) conv_simt_fwd_kernel(const SimtGeom g) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tid = threadIdx.x;
  const long long P = (long long)g.N * g.oD * g.oH * g.oW;
  const long long pos0 = (long long)blockIdx.x * 64;
  const int co0 = blockIdx.y * 64;
  const int K = g.kd * g.kh * g.kw * g.cin_total;
  const int tx = tid & 15, ty = tid >> 4;  // tx -> 4 couts, ty -> 4 positions
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A-load mapping: thread loads positions (tid>>4)*4.. wait: 64 pos x 16 k = 1024 elems, 4 per thread
  const int a_k = tid & 15;        // k within chunk
  const int a_p = tid >> 4;        // position group: positions a_p + 16*i
  // decode the 4 positions once
  int pn[4], pd[4], ph[4], pw[4];
  bool pv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long pos = pos0 + a_p + 16 * i;
    pv[i] = pos < P;
    long long r = pv[i] ? pos : 0;
    pw[i] = (int)(r % g.oW); r /= g.oW;
    ph[i] = (int)(r % g.oH); r /= g.oH;
    pd[i] = (int)(r % g.oD); r /= g.oD;
    pn[i] = (int)r;
  }
  const int b_k = tid >> 4;   // 0..15
  const int b_c = (tid & 15) * 4;

  for (int k0 = 0; k0 < K; k0 += 16) {
    // ---- A tile
    {
      const int k = k0 + a_k;
      int tap = 0, ci = 0, kd = 0, kh = 0, kw = 0;
      const bool kvalid = k < K;
      if (kvalid) {
        tap = k / g.cin_total;
        ci = k - tap * g.cin_total;
        kw = tap % g.kw;
        int t2 = tap / g.kw;
        kh = t2 % g.kh;
        kd = t2 / g.kh;
      }
      const T* xb;
      int cl, cs;
      if (ci < g.cseg[0]) { xb = reinterpret_cast<const T*>(g.x[0]); cl = ci; cs = g.cseg[0]; }
      else { xb = reinterpret_cast<const T*>(g.x[1]); cl = ci - g.cseg[0]; cs = g.cseg[1]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int gd, gh, gw;
        if (kvalid && pv[i] && gather_coord(pd[i], kd, g.mul[0], g.off0[0], g.offk[0], g.div[0], g.iD, gd) &&
            gather_coord(ph[i], kh, g.mul[1], g.off0[1], g.offk[1], g.div[1], g.iH, gh) &&
            gather_coord(pw[i], kw, g.mul[2], g.off0[2], g.offk[2], g.div[2], g.iW, gw)) {
          const long long idx = ((((long long)pn[i] * g.iD + gd) * g.iH + gh) * g.iW + gw) * cs + cl;
          v = to_f32<T>(xb[idx]);
        }
        sA[a_k][a_p + 16 * i] = v;
      }
    }
    // ---- B tile
    {
      const int k = k0 + b_k;
      const bool kvalid = k < K;
      int tap = 0, ci = 0;
      if (kvalid) {
        tap = k / g.cin_total;
        ci = k - tap * g.cin_total;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = co0 + b_c + j;
        float v = 0.f;
        if (kvalid && co < g.cout) v = __ldg(g.w + tap * g.ws_tap + ci * g.ws_ci + (long long)(co + g.co_off) * g.ws_co);
        sB[b_k][b_c + j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- store
  OutT* out = reinterpret_cast<OutT*>(g.y);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long pos = pos0 + ty + 16 * i;
    if (pos >= P) continue;
    long long r = pos;
    const int ow = (int)(r % g.oW); r /= g.oW;
    const int oh = (int)(r % g.oH); r /= g.oH;
    const int od = (int)(r % g.oD); r /= g.oD;
    const long long off = r * g.yo[3] + od * g.yo[2] + oh * g.yo[1] + ow * g.yo[0];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < g.cout) {
        float v = acc[i][j];
        if (g.bias) v += __ldg(g.bias + co);
        if (g.accumulate) v += to_f32<OutT>(out[off + co]);
        out[off + co] = from_f32<OutT>(v);
      }
    }
  }
}

int simt_conv_launch(const SimtGeom& g, int dtype, int out_f32, cudaStream_t stream, char* err, size_t errlen) {
  const long long P = (long long)g.N * g.oD * g.oH * g.oW;
  dim3 grid((unsigned)((P + 63) / 64), (unsigned)((g.cout + 63) / 64));
  if (dtype == 0) {
    if (out_f32) conv_simt_fwd_kernel<bf16, float><<<grid, 256, 0, stream>>>(g);
    else conv_simt_fwd_kernel<bf16, bf16><<<grid, 256, 0, stream>>>(g);
  } else {
    conv_simt_fwd_kernel<float, float><<<grid, 256, 0, stream>>>(g);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_simt launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// filter gradient: dW[tap][gch][qch] += sum_pos G(pos,tap)[gch] * Q(pos)[qch]
// GEMM with M = taps*Cg (rows), N = Cq, K = positions; split-K over blockIdx.z with fp32 atomics
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_simt_wgrad_kernel(const SimtWgradGeom g) {
  __shared__ float sA[16][64 + 4];  // [pos][m]
  __shared__ float sB[16][64 + 4];  // [pos][n]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int Mtot = g.kd * g.kh * g.kw * g.cg;
  const long long P = (long long)g.N * g.qD * g.qH * g.qW;
  const long long chunk = (P + gridDim.z - 1) / gridDim.z;
  const long long pbeg = (long long)blockIdx.z * chunk;
  const long long pend = pbeg + chunk < P ? pbeg + chunk : P;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A-load: thread -> (pos = tid>>4, m = (tid&15) + 16*i)
  const int a_pos = tid >> 4, a_m = tid & 15;
  int mtap[4], mch[4], mkd[4], mkh[4], mkw[4];
  bool mv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + a_m + 16 * i;
    mv[i] = m < Mtot;
    const int mm = mv[i] ? m : 0;
    mtap[i] = mm / g.cg;
    mch[i] = mm - mtap[i] * g.cg;
    mkw[i] = mtap[i] % g.kw;
    const int t2 = mtap[i] / g.kw;
    mkh[i] = t2 % g.kh;
    mkd[i] = t2 / g.kh;
  }
  const T* G = reinterpret_cast<const T*>(g.g);
  const T* Q = reinterpret_cast<const T*>(g.q);
  for (long long p0 = pbeg; p0 < pend; p0 += 16) {
    {
      const long long pos = p0 + a_pos;
      const bool pvalid = pos < pend;
      long long r = pvalid ? pos : 0;
      const int qw = (int)(r % g.qW); r /= g.qW;
      const int qh = (int)(r % g.qH); r /= g.qH;
      const int qd = (int)(r % g.qD); r /= g.qD;
      const int qn = (int)r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int gd, gh, gw;
        if (pvalid && mv[i] && gather_coord(qd, mkd[i], g.mul[0], g.off0[0], g.offk[0], g.div[0], g.gD, gd) &&
            gather_coord(qh, mkh[i], g.mul[1], g.off0[1], g.offk[1], g.div[1], g.gH, gh) &&
            gather_coord(qw, mkw[i], g.mul[2], g.off0[2], g.offk[2], g.div[2], g.gW, gw)) {
          v = to_f32<T>(G[((((long long)qn * g.gD + gd) * g.gH + gh) * g.gW + gw) * g.cg + mch[i]]);
        }
        sA[a_pos][a_m + 16 * i] = v;
      }
      // B tile: same position row, 64 n
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + a_m + 16 * j;
        float v = 0.f;
        if (pvalid && n < g.cq) v = to_f32<T>(Q[pos * g.cq + n]);
        sB[a_pos][a_m + 16 * j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
    const int tap = m / g.cg, ch = m - tap * g.cg;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.cq) atomicAdd(g.dw + tap * g.ws_tap + (long long)(ch + g.g_c0) * g.ws_g + n + g.q_c0, acc[i][j]);
    }
  }
}

int simt_wgrad_launch(const SimtWgradGeom& g, int dtype, cudaStream_t stream, char* err, size_t errlen) {
  const int Mtot = g.kd * g.kh * g.kw * g.cg;
  const long long P = (long long)g.N * g.qD * g.qH * g.qW;
  const int gx = (Mtot + 63) / 64, gy = (g.cq + 63) / 64;
  long long want = (4 * 148 + (long long)gx * gy - 1) / ((long long)gx * gy);
  long long maxz = (P + 255) / 256;
  int gz = (int)std::max(1ll, std::min(want, maxz));
  dim3 grid(gx, gy, gz);
  if (dtype == 0) conv_simt_wgrad_kernel<bf16><<<grid, 256, 0, stream>>>(g);
  else conv_simt_wgrad_kernel<float><<<grid, 256, 0, stream>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "conv_simt_wgrad launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per-channel partial statistics of a [P][C] matrix: out[r][0][c] = sum, out[r][1][c] = sum sq
// (rows r = blockIdx.x, each covering a contiguous slab of positions); also bias gradient (sum only,
// atomically accumulated).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) col_stats_kernel(const T* __restrict__ x, long long P, int C, int rows,
                                                         float* __restrict__ out) {
  // block: 256 threads = 8 position lanes x 32 channel lanes; loops over channel groups of 32
  const int r = blockIdx.x;
  const long long per = (P + rows - 1) / rows;
  const long long pbeg = r * per, pend = pbeg + per < P ? pbeg + per : P;
  __shared__ float s1[8][33], s2[8][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + cl;
    float a = 0.f, b = 0.f;
    if (c < C)
      for (long long p = pbeg + pl; p < pend; p += 8) {
        const float v = to_f32<T>(x[p * C + c]);
        a += v;
        b += v * v;
      }
    s1[pl][cl] = a;
    s2[pl][cl] = b;
    __syncthreads();
    if (pl == 0 && c < C) {
      float ta = 0.f, tb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ta += s1[i][cl];
        tb += s2[i][cl];
      }
      out[((long long)r * 2 + 0) * C + c] = ta;
      out[((long long)r * 2 + 1) * C + c] = tb;
    }
    __syncthreads();
  }
}

int col_stats_launch(const void* x, int dtype, long long P, int C, int rows, float* out, cudaStream_t stream) {
  if (dtype == 0) col_stats_kernel<bf16><<<rows, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), P, C, rows, out);
  else col_stats_kernel<float><<<rows, 256, 0, stream>>>(reinterpret_cast<const float*>(x), P, C, rows, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <typename T>
__global__ void __launch_bounds__(256) col_sum_accum_kernel(const T* __restrict__ x, long long P, int C,
                                                             float* __restrict__ out) {
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < P ? pbeg + per : P;
  __shared__ float s1[8][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + cl;
    float a = 0.f;
    if (c < C)
      for (long long p = pbeg + pl; p < pend; p += 8) a += to_f32<T>(x[p * C + c]);
    s1[pl][cl] = a;
    __syncthreads();
    if (pl == 0 && c < C) {
      float ta = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ta += s1[i][cl];
      atomicAdd(out + c, ta);
    }
    __syncthreads();
  }
}

int col_sum_accum_launch(const void* x, int dtype, long long P, int C, float* out, cudaStream_t stream) {
  int blocks = (int)std::max(1ll, std::min(592ll, (P + 255) / 256));
  if (dtype == 0) col_sum_accum_kernel<bf16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), P, C, out);
  else col_sum_accum_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(x), P, C, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// weight packing fp32 TF layout -> bf16 K-major [rows_pad][taps*cols]
//   out[r][tap*cols + c] = w[tap*s_tap + r*s_r + c*s_c]   (zero for r >= rows)
// ---------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int taps, int rows, int rows_pad,
                                    int cols, long long s_tap, long long s_r, long long s_c) {
  const long long total = (long long)rows_pad * taps * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    long long t = i / cols;
    const int tap = (int)(t % taps);
    const int r = (int)(t / taps);
    float v = 0.f;
    if (r < rows) v = __ldg(w + tap * s_tap + r * s_r + c * s_c);
    out[i] = __float2bfloat16_rn(v);
  }
}

int pack_weights_launch(const float* w, void* out, int taps, int rows, int rows_pad, int cols, long long s_tap,
                        long long s_r, long long s_c, cudaStream_t stream) {
  const long long total = (long long)rows_pad * taps * cols;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pack_weights_kernel<<<blocks, 256, 0, stream>>>(w, reinterpret_cast<bf16*>(out), taps, rows, rows_pad, cols, s_tap,
                                                  s_r, s_c);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace sap3d
