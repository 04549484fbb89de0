// GroupNorm statistics and the CBAM (channel + spatial attention) kernels of the GN model variant
// (gn/p3d_gn.py:24-46,175; utils/network.py:65-87,198-274).  All bandwidth-bound: NDHWC rows are read
// with 128-bit vectors, per-channel / per-position reductions use warp shuffles; no transposes (the
// reference transposes to NCDHW and back around tf.nn.moments).
#include <string.h>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

// ---- per-sample per-channel partial sums: x [N][S][C] -> part [N][rows][3][C] (sum, sumsq, max) ----
// grid = (rows, C/64, N); 256 threads = 8 channel-vector lanes x 32 position lanes (one 128 B line per position)
template <typename T>
__global__ void __launch_bounds__(256) sample_channel_partial_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                                      long long S, int C, int rows, float* __restrict__ part) {
  __shared__ float red[8][3][64];
  const int cv = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.y * 64 + cv * 8;
  const int n = blockIdx.z;
  const long long per = (S + rows - 1) / rows;
  const long long pbeg = blockIdx.x * per, pend = pbeg + per < S ? pbeg + per : S;
  float a[8], b[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; m[j] = -INFINITY; }
  if (c < C) {
    const T* xb = x + (long long)n * S * C + c;
    for (long long pos = pbeg + pl; pos < pend; pos += 32) {
      float v[8];
      Vec8<T>::load(xb + pos * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = scale ? v[j] * scale[(long long)n * C + c + j] : v[j];
        a[j] += t; b[j] += t * t; m[j] = fmaxf(m[j], t);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);  a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
    b[j] += __shfl_xor_sync(0xffffffffu, b[j], 8);  b[j] += __shfl_xor_sync(0xffffffffu, b[j], 16);
    m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], 8)); m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], 16));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][0][lane * 8 + j] = a[j]; red[warp][1][lane * 8 + j] = b[j]; red[warp][2][lane * 8 + j] = m[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 192) {
    const int i = threadIdx.x >> 6, ch = threadIdx.x & 63;
    float v = i == 2 ? -INFINITY : 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v = i == 2 ? fmaxf(v, red[w][i][ch]) : v + red[w][i][ch];
    const int cc = blockIdx.y * 64 + ch;
    if (cc < C) part[(((long long)n * rows + blockIdx.x) * 3 + i) * C + cc] = v;
  }
}

// GroupNorm finalize: part -> per (n, c) scale / shift, per (n, g) mean / rstd.  One block per (n, g).
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* __restrict__ part, int rows, int C, int G, long long S,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                           float* scale, float* shift, float* save_mean, float* save_rstd) {
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < rows * cpg; i += blockDim.x) {
    const int r = i / cpg, c = g * cpg + i % cpg;
    a += (double)part[(((long long)n * rows + r) * 3 + 0) * C + c];
    b += (double)part[(((long long)n * rows + r) * 3 + 1) * C + c];
  }
  __shared__ double sa[4], sb[4];
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  const double cnt = (double)S * cpg;
  const double ta = sa[0] + sa[1] + sa[2] + sa[3], tb = sb[0] + sb[1] + sb[2] + sb[3];
  const double mean = ta / cnt;
  double var = tb / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (threadIdx.x == 0 && save_mean) { save_mean[n * G + g] = (float)mean; save_rstd[n * G + g] = rstd; }
  for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
    const int ch = g * cpg + c;
    const float gm = gamma[ch];
    scale[(long long)n * C + ch] = gm * rstd;
    shift[(long long)n * C + ch] = beta[ch] - (float)mean * gm * rstd;
  }
}

// CBAM channel attention: part (sum, -, max over positions) -> shared MLP -> sigmoid -> cscale [N][C]
// one block per sample (utils/network.py:208-249)
__global__ void __launch_bounds__(1024) cbam_channel_mlp_kernel(const float* __restrict__ part, int rows, int C, int hidden, long long S,
                                                                 const float* __restrict__ w0, const float* __restrict__ b0,
                                                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                                                 float* __restrict__ cscale, float* __restrict__ save) {
  extern __shared__ float sm[];  // avg[C], mx[C], h_avg[hidden], h_max[hidden], scratch[2][8][hidden]
  float* avg = sm;
  float* mx = sm + C;
  float* ha = sm + 2 * C;
  float* hm = ha + hidden;
  float* scr = hm + hidden;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, m = -INFINITY;
    for (int r = 0; r < rows; ++r) {
      a += part[(((long long)n * rows + r) * 3 + 0) * C + c];
      m = fmaxf(m, part[(((long long)n * rows + r) * 3 + 2) * C + c]);
    }
    avg[c] = a / (float)S;
    mx[c] = m;
  }
  __syncthreads();
  // hidden layer: the C-long dot products are split over 8 slices of the block (thread = (slice, j)); w0 rows are read
  // coalesced along j; slices are combined in a fixed order
  {
    const int nsl = blockDim.x / hidden >= 8 ? 8 : (blockDim.x / hidden >= 1 ? blockDim.x / hidden : 1);
    const int j = threadIdx.x % hidden, sl = threadIdx.x / hidden;
    if (sl < nsl) {
      float sa = 0.f, sx = 0.f;
      const int per = (C + nsl - 1) / nsl;
      const int c1 = min(C, (sl + 1) * per);
      for (int c = sl * per; c < c1; ++c) {
        const float w = w0[(long long)c * hidden + j];
        sa = fmaf(avg[c], w, sa);
        sx = fmaf(mx[c], w, sx);
      }
      scr[(0 * 8 + sl) * hidden + j] = sa;
      scr[(1 * 8 + sl) * hidden + j] = sx;
    }
    __syncthreads();
    if (threadIdx.x < hidden) {
      float sa = b0[threadIdx.x], sx = b0[threadIdx.x];
      for (int q = 0; q < nsl; ++q) { sa += scr[(0 * 8 + q) * hidden + threadIdx.x]; sx += scr[(1 * 8 + q) * hidden + threadIdx.x]; }
      ha[threadIdx.x] = fmaxf(sa, 0.f);
      hm[threadIdx.x] = fmaxf(sx, 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 2.f * b1[c];
    for (int j = 0; j < hidden; ++j) s = fmaf(ha[j] + hm[j], w1[(long long)j * C + c], s);
    cscale[(long long)n * C + c] = 1.f / (1.f + __expf(-s));
  }
  if (save) {  // [N][2C + 2 hidden]: avg, max, relu(hidden(avg)), relu(hidden(max)) — kept for the backward pass
    float* sv = save + (long long)n * (2 * C + 2 * hidden);
    for (int i = threadIdx.x; i < 2 * C + 2 * hidden; i += blockDim.x) sv[i] = sm[i];
  }
}

// spatial pooling of the channel-scaled tensor: sp[n][pos][0] = mean_c(x*cscale), [1] = max_c  (one warp per position)
template <typename T>
__global__ void __launch_bounds__(256) cbam_spatial_pool_kernel(const T* __restrict__ x, const float* __restrict__ cscale, long long S,
                                                                 int C, long long total, float* __restrict__ sp) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  // a warp reduces U positions at a time (their loads are issued together) and reads the channel scales as 128-bit loads:
  // with one position and 8 scalar scale loads per lane the kernel streamed at 1.56 TB/s (r02 ncu)
  constexpr int U = 4;
  for (long long pos0 = warp_id * U; pos0 < total; pos0 += nwarps * U) {
    float a[U], m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a[u] = 0.f; m[u] = -INFINITY; }
    for (int c = lane * 8; c < C; c += 256) {
      typename Vec8<T>::Raw rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pos0 + u < total) rv[u] = Vec8<T>::load_raw(x + (pos0 + u) * C + c);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pos0 + u >= total) break;
        const float* cs = cscale + ((pos0 + u) / S) * C + c;
        const float4 c0 = *reinterpret_cast<const float4*>(cs), c1 = *reinterpret_cast<const float4*>(cs + 4);
        const float cs8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        float v[8];
        Vec8<T>::unpack(rv[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = __fmul_rn(v[j], cs8[j]);  // rounded product: the backward pass re-derives arg-max from it
          a[u] += t;
          m[u] = fmaxf(m[u], t);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pos0 + u >= total) break;
      const float as = warp_sum(a[u]);
      const float ms = warp_max(m[u]);
      if (lane == 0) {
        sp[(pos0 + u) * 2 + 0] = as / (float)C;
        sp[(pos0 + u) * 2 + 1] = ms;
      }
    }
  }
}

// 7x7x7 'SAME' conv over the 2-channel map (no bias) + sigmoid -> att [N][D][H][W]
__global__ void __launch_bounds__(128) cbam_spatial_conv_kernel(const float* __restrict__ sp, const float* __restrict__ w, int N, int D,
                                                                 int H, int W, float* __restrict__ att) {
  __shared__ float sw[343 * 2];
  for (int i = threadIdx.x; i < 686; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long total = (long long)N * D * H * W;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    long long r = o;
    const int ow = (int)(r % W); r /= W;
    const int oh = (int)(r % H); r /= H;
    const int od = (int)(r % D);
    const int n = (int)(r / D);
    float acc = 0.f;
    for (int a = 0; a < 7; ++a) {
      const int zd = od + a - 3;
      if (zd < 0 || zd >= D) continue;
      for (int b = 0; b < 7; ++b) {
        const int zh = oh + b - 3;
        if (zh < 0 || zh >= H) continue;
        for (int e = 0; e < 7; ++e) {
          const int zw = ow + e - 3;
          if (zw < 0 || zw >= W) continue;
          const float* s = sp + ((((long long)n * D + zd) * H + zh) * W + zw) * 2;
          const float* ww = sw + ((a * 7 + b) * 7 + e) * 2;
          acc = fmaf(s[0], ww[0], acc);
          acc = fmaf(s[1], ww[1], acc);
        }
      }
    }
    att[o] = 1.f / (1.f + __expf(-acc));
  }
}

// block output: y = relu( (a*s1[n][c]+t1[n][c]) + r * cscale[n][c] * att[n][pos] )   (gn/p3d_gn.py:175-177)
template <typename T>
__global__ void __launch_bounds__(256) cbam_merge_kernel(const T* __restrict__ a, const float* __restrict__ s1, const float* __restrict__ t1,
                                                          const T* __restrict__ r, const float* __restrict__ cscale,
                                                          const float* __restrict__ att, T* __restrict__ y, long long S, int C, long long P) {
  const long long nvec = P * C / 8;
  // four vectors per thread in flight, and the per-(sample, channel) constants as 128-bit loads: the r01 form (one vector in
  // flight, 24 scalar constant loads per vector) streamed at 1.8 TB/s (r02 ncu)
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto ld8 = [](const float* q, float (&v)[8]) {
    const float4 x = *reinterpret_cast<const float4*>(q), y = *reinterpret_cast<const float4*>(q + 4);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  };
  for (long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    typename Vec8<T>::Raw ra[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        rr[u] = Vec8<T>::load_raw(r + v * 8);
        if (a != nullptr) ra[u] = Vec8<T>::load_raw(a + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v >= nvec) break;
      const long long e = v * 8;
      const long long pos = e / C;
      const int c = (int)(e - pos * C);
      const long long n = pos / S;
      float av[8], rv[8], o[8], cs[8];
      Vec8<T>::unpack(rr[u], rv);
      ld8(cscale + n * C + c, cs);
      const float at = att[pos];
      if (a == nullptr) {   // stand-alone cbam_block (utils/network.py:198-206): refined feature, no main branch, no ReLU
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rv[j] * cs[j] * at;
      } else {
        float s8[8], t8[8];
        Vec8<T>::unpack(ra[u], av);
        ld8(s1 + n * C + c, s8);
        ld8(t1 + n * C + c, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(av[j], s8[j], t8[j]) + rv[j] * cs[j] * at, 0.f);
      }
      Vec8<T>::store(y + e, o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) concat_channels_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long P,
                                                               int ca, int cb) {
  const int cv = (ca + cb) / 8;
  const long long nvec = P * cv;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long pos = v / cv;
    const int c = (int)(v - pos * cv) * 8;
    const uint4 u = c < ca ? *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a + pos * ca + c))
                           : *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(b + pos * cb + (c - ca)));
    if (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(y + v * 8) = u;
    } else {
      const T* src = c < ca ? a + pos * ca + c : b + pos * cb + (c - ca);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[v * 8 + j] = src[j];
    }
  }
}

int grid_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int sap3d_sample_stats_rows(int64_t S, int32_t C, int32_t N) {
  const long long chunks = (C + 63) / 64;
  long long rows = (2 * 148 + chunks * N - 1) / (chunks * N);
  const long long mx = (S + 31) / 32;
  if (rows > mx) rows = mx;
  if (rows > 64) rows = 64;
  if (rows < 1) rows = 1;
  return (int)rows;
}

/* part [N][rows][3][C] = per-sample per-channel (sum, sum of squares, max) partials of x (optionally x*scale[n][c]) */
int sap3d_sample_channel_partials(int32_t dtype, const void* x, const float* scale, int32_t N, int64_t S, int32_t C, int32_t rows,
                                  float* part, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("sample_channel_partials: C %% 8 != 0");
  dim3 grid(rows, (C + 63) / 64, N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16) sample_channel_partial_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), scale, S, C, rows, part);
  else sample_channel_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), scale, S, C, rows, part);
  return check_launch("sample_channel_partials");
}

int sap3d_gn_finalize(const float* part, int32_t rows, int32_t N, int64_t S, int32_t C, int32_t G, const float* gamma,
                      const float* beta, float eps, float* scale, float* shift, float* save_mean, float* save_rstd, void* stream) {
  if (require_device()) return 1;
  if (C % G != 0) return set_error("gn_finalize: C %% G != 0");
  gn_finalize_kernel<<<N * G, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, rows, C, G, S, gamma, beta, eps, scale, shift, save_mean, save_rstd);
  return check_launch("gn_finalize");
}

/* CBAM forward on r [N][D][H][W][C]: cscale [N][C], sp [N][S][2], att [N][S]; merged block output
 * y = relu(a*s1+t1 + r*cscale*att) with per-sample affine (s1, t1) of the main branch. */
int sap3d_cbam_fwd(int32_t dtype, const void* r, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t hidden,
                   const float* w0, const float* b0, const float* w1, const float* b1, const float* w_sp, float* part,
                   int32_t rows, float* cscale, float* sp, float* att, float* save, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("cbam_fwd: C %% 8 != 0");
  const long long S = (long long)D * H * W;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (w0 != nullptr) {   // channel attention; w0 == NULL: spatial attention alone, the caller keeps cscale == 1
    if (sap3d_sample_channel_partials(dtype, r, nullptr, N, S, C, rows, part, stream)) return 1;
    if (hidden > 1024) return set_error("cbam_fwd: hidden > 1024");
    const size_t sm = (size_t)(2 * C + 2 * hidden + 16 * hidden) * sizeof(float);
    const int threads = hidden * 8 <= 1024 ? (hidden * 8 < 256 ? 256 : hidden * 8) : 1024;
    cbam_channel_mlp_kernel<<<N, threads, sm, st>>>(part, rows, C, hidden, S, w0, b0, w1, b1, cscale, save);
    if (check_launch("cbam_channel_mlp")) return 1;
  }
  if (w_sp == nullptr) return 0;   // channel attention alone: the caller keeps att == 1
  const long long total = (long long)N * S;
  const int blocks = (int)((total * 32 + 255) / 256 > 148 * 8 ? 148 * 8 : (total * 32 + 255) / 256);
  if (dtype == SAP3D_BF16) cbam_spatial_pool_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(r), cscale, S, C, total, sp);
  else cbam_spatial_pool_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(r), cscale, S, C, total, sp);
  if (check_launch("cbam_spatial_pool")) return 1;
  cbam_spatial_conv_kernel<<<(int)((total + 127) / 128 > 148 * 16 ? 148 * 16 : (total + 127) / 128), 128, 0, st>>>(sp, w_sp, N, D, H, W, att);
  return check_launch("cbam_spatial_conv");
}

int sap3d_cbam_merge(int32_t dtype, const void* a, const float* s1, const float* t1, const void* r, const float* cscale,
                     const float* att, void* y, int32_t N, int64_t S, int32_t C, void* stream) {
  if (require_device()) return 1;
  const long long P = (long long)N * S;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    cbam_merge_kernel<bf16><<<grid_for(P * C / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(a), s1, t1, reinterpret_cast<const bf16*>(r), cscale, att, reinterpret_cast<bf16*>(y), S, C, P);
  else
    cbam_merge_kernel<float><<<grid_for(P * C / 8), 256, 0, st>>>(reinterpret_cast<const float*>(a), s1, t1, reinterpret_cast<const float*>(r), cscale, att, reinterpret_cast<float*>(y), S, C, P);
  return check_launch("cbam_merge");
}

int sap3d_concat_channels(int32_t dtype, const void* a, const void* b, void* y, int64_t P, int32_t ca, int32_t cb, void* stream) {
  if (require_device()) return 1;
  if (ca % 8 != 0 || cb % 8 != 0) return set_error("concat_channels: channel counts must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long nvec = P * (ca + cb) / 8;
  if (dtype == SAP3D_BF16)
    concat_channels_kernel<bf16><<<grid_for(nvec), 256, 0, st>>>(reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), reinterpret_cast<bf16*>(y), P, ca, cb);
  else
    concat_channels_kernel<float><<<grid_for(nvec), 256, 0, st>>>(reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), reinterpret_cast<float*>(y), P, ca, cb);
  return check_launch("concat_channels");
}

}  // extern "C"
