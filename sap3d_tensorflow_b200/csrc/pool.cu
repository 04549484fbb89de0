// 3-D max pooling (NDHWC, TF 'SAME' / 'VALID' geometry) forward and backward, 128-bit channel vectors.
// Replaces tf.nn.max_pool3d (p3d.py:347-348,354,360,366: the k(2,1,1) temporal pools and the
// k(2,3,3)/s(2,2,2) stem pool) and tf.layers.max_pooling3d (utils/network.py:6-7).
#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

struct PoolArgs {
  const void* x; void* y; const void* dy; void* dx;
  int N, D, H, W, C;
  int Do, Ho, Wo;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int accumulate;
  unsigned char* amax;        // forward: window-local index of the first maximum, [outputs][C] (nullable)
  const unsigned char* amax_in;  // backward: the same tensor (nullable -> re-scan the windows)
};

template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const PoolArgs p) {
  const int cv = p.C / 8;
  const long long total = (long long)p.N * p.Do * p.Ho * p.Wo * cv;
  const T* x = reinterpret_cast<const T*>(p.x);
  T* y = reinterpret_cast<T*>(p.y);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int ow = (int)(r % p.Wo); r /= p.Wo;
    const int oh = (int)(r % p.Ho); r /= p.Ho;
    const int od = (int)(r % p.Do);
    const int n = (int)(r / p.Do);
    float m[8];
    unsigned am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; am[j] = 255u; }
    for (int a = 0; a < p.kd; ++a) {
      const int id = od * p.sd + a - p.pd;
      if (id < 0 || id >= p.D) continue;
      for (int b = 0; b < p.kh; ++b) {
        const int ih = oh * p.sh + b - p.ph;
        if (ih < 0 || ih >= p.H) continue;
        for (int e = 0; e < p.kw; ++e) {
          const int iw = ow * p.sw + e - p.pw;
          if (iw < 0 || iw >= p.W) continue;
          float v[8];
          Vec8<T>::load(x + ((((long long)n * p.D + id) * p.H + ih) * p.W + iw) * p.C + c, v);
          const unsigned li = (unsigned)((a * p.kh + b) * p.kw + e);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (v[j] > m[j] || am[j] == 255u) { m[j] = v[j]; am[j] = li; }   // strict '>' keeps the FIRST maximum
        }
      }
    }
    const long long yoff = ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.C + c;
    Vec8<T>::store(y + yoff, m);
    if (p.amax) {
      uint2 u;
      u.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      u.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(p.amax + yoff) = u;
    }
  }
}

// gather-form backward: the gradient of a window goes to its FIRST maximal element (scan order d,h,w)
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const PoolArgs p) {
  const int cv = p.C / 8;
  const long long total = (long long)p.N * p.D * p.H * p.W * cv;
  const T* x = reinterpret_cast<const T*>(p.x);
  const T* dy = reinterpret_cast<const T*>(p.dy);
  T* dx = reinterpret_cast<T*>(p.dx);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int iw = (int)(r % p.W); r /= p.W;
    const int ih = (int)(r % p.H); r /= p.H;
    const int id = (int)(r % p.D);
    const int n = (int)(r / p.D);
    const long long xoff = ((((long long)n * p.D + id) * p.H + ih) * p.W + iw) * p.C + c;
    float me[8], g[8];
    Vec8<T>::load(x + xoff, me);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    // windows containing this element: o*s - pad <= i <= o*s - pad + k - 1
    const int od_lo = max(0, (id + p.pd - p.kd + p.sd) / p.sd), od_hi = min(p.Do - 1, (id + p.pd) / p.sd);
    const int oh_lo = max(0, (ih + p.ph - p.kh + p.sh) / p.sh), oh_hi = min(p.Ho - 1, (ih + p.ph) / p.sh);
    const int ow_lo = max(0, (iw + p.pw - p.kw + p.sw) / p.sw), ow_hi = min(p.Wo - 1, (iw + p.pw) / p.sw);
    for (int od = od_lo; od <= od_hi; ++od)
      for (int oh = oh_lo; oh <= oh_hi; ++oh)
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
          // is `me` the first maximum of window (od,oh,ow)?
          bool first[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) first[j] = true;
          bool before = true;  // scanning elements that precede `me`
          for (int a = 0; a < p.kd; ++a) {
            const int jd = od * p.sd + a - p.pd;
            if (jd < 0 || jd >= p.D) continue;
            for (int b = 0; b < p.kh; ++b) {
              const int jh = oh * p.sh + b - p.ph;
              if (jh < 0 || jh >= p.H) continue;
              for (int e = 0; e < p.kw; ++e) {
                const int jw = ow * p.sw + e - p.pw;
                if (jw < 0 || jw >= p.W) continue;
                if (jd == id && jh == ih && jw == iw) {
                  before = false;
                  continue;
                }
                float v[8];
                Vec8<T>::load(x + ((((long long)n * p.D + jd) * p.H + jh) * p.W + jw) * p.C + c, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (before ? (v[j] >= me[j]) : (v[j] > me[j])) first[j] = false;
                }
              }
            }
          }
          float d[8];
          Vec8<T>::load(dy + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.C + c, d);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (first[j]) g[j] += d[j];
        }
    if (p.accumulate) {
      float old[8];
      Vec8<T>::load(dx + xoff, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += old[j];
    }
    Vec8<T>::store(dx + xoff, g);
  }
}

// backward with the arg-max saved by the forward pass: one (dy, index) vector pair per covering window
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_idx_kernel(const PoolArgs p) {
  const int cv = p.C / 8;
  const long long total = (long long)p.N * p.D * p.H * p.W * cv;
  const T* dy = reinterpret_cast<const T*>(p.dy);
  T* dx = reinterpret_cast<T*>(p.dx);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long r = i / cv;
    const int iw = (int)(r % p.W); r /= p.W;
    const int ih = (int)(r % p.H); r /= p.H;
    const int id = (int)(r % p.D);
    const int n = (int)(r / p.D);
    const long long xoff = ((((long long)n * p.D + id) * p.H + ih) * p.W + iw) * p.C + c;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    const int od_lo = max(0, (id + p.pd - p.kd + p.sd) / p.sd), od_hi = min(p.Do - 1, (id + p.pd) / p.sd);
    const int oh_lo = max(0, (ih + p.ph - p.kh + p.sh) / p.sh), oh_hi = min(p.Ho - 1, (ih + p.ph) / p.sh);
    const int ow_lo = max(0, (iw + p.pw - p.kw + p.sw) / p.sw), ow_hi = min(p.Wo - 1, (iw + p.pw) / p.sw);
    for (int od = od_lo; od <= od_hi; ++od)
      for (int oh = oh_lo; oh <= oh_hi; ++oh)
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
          const unsigned li = (unsigned)(((id - (od * p.sd - p.pd)) * p.kh + (ih - (oh * p.sh - p.ph))) * p.kw + (iw - (ow * p.sw - p.pw)));
          const long long yoff = ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.C + c;
          const uint2 u = *reinterpret_cast<const uint2*>(p.amax_in + yoff);
          float d[8];
          Vec8<T>::load(dy + yoff, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned a = ((j < 4 ? u.x : u.y) >> (8 * (j & 3))) & 255u;
            if (a == li) g[j] += d[j];
          }
        }
    if (p.accumulate) {
      float old[8];
      Vec8<T>::load(dx + xoff, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += old[j];
    }
    Vec8<T>::store(dx + xoff, g);
  }
}

int fill(PoolArgs& p, int N, int D, int H, int W, int C, const int32_t* k, const int32_t* s, int same) {
  p.N = N; p.D = D; p.H = H; p.W = W; p.C = C;
  p.kd = k[0]; p.kh = k[1]; p.kw = k[2];
  p.sd = s[0]; p.sh = s[1]; p.sw = s[2];
  const int I[3] = {D, H, W};
  int O[3], pb[3];
  for (int i = 0; i < 3; ++i) {
    if (same) {
      O[i] = (I[i] + s[i] - 1) / s[i];
      int pt = (O[i] - 1) * s[i] + k[i] - I[i];
      if (pt < 0) pt = 0;
      pb[i] = pt / 2;
    } else {
      O[i] = (I[i] - k[i]) / s[i] + 1;
      pb[i] = 0;
    }
    if (O[i] < 1) return 1;
  }
  p.Do = O[0]; p.Ho = O[1]; p.Wo = O[2];
  p.pd = pb[0]; p.ph = pb[1]; p.pw = pb[2];
  return 0;
}

int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int sap3d_maxpool3d_out_dims(int32_t D, int32_t H, int32_t W, const int32_t* ksize, const int32_t* strides, int32_t same,
                             int32_t* out_dhw) {
  PoolArgs p;
  if (fill(p, 1, D, H, W, 8, ksize, strides, same)) return set_error("maxpool3d: empty output");
  out_dhw[0] = p.Do; out_dhw[1] = p.Ho; out_dhw[2] = p.Wo;
  return 0;
}

int sap3d_maxpool3d_fwd(int32_t dtype, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                        const int32_t* ksize, const int32_t* strides, int32_t same, void* y, uint8_t* argmax, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("maxpool3d: C must be a multiple of 8");
  PoolArgs p;
  if (fill(p, N, D, H, W, C, ksize, strides, same)) return set_error("maxpool3d: empty output");
  if (argmax && ksize[0] * ksize[1] * ksize[2] > 254) return set_error("maxpool3d: window too large for the arg-max encoding");
  p.x = x; p.y = y; p.dy = nullptr; p.dx = nullptr; p.accumulate = 0; p.amax = argmax; p.amax_in = nullptr;
  const long long total = (long long)N * p.Do * p.Ho * p.Wo * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16) maxpool_fwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>(p);
  else maxpool_fwd_kernel<float><<<grid_for(total), 256, 0, st>>>(p);
  return check_launch("maxpool3d_fwd");
}

int sap3d_maxpool3d_bwd(int32_t dtype, const void* x, const void* dy, int32_t N, int32_t D, int32_t H, int32_t W,
                        int32_t C, const int32_t* ksize, const int32_t* strides, int32_t same, const uint8_t* argmax, void* dx,
                        int32_t accumulate, void* stream) {
  if (require_device()) return 1;
  if (C % 8 != 0) return set_error("maxpool3d: C must be a multiple of 8");
  PoolArgs p;
  if (fill(p, N, D, H, W, C, ksize, strides, same)) return set_error("maxpool3d: empty output");
  p.x = x; p.y = nullptr; p.dy = dy; p.dx = dx; p.accumulate = accumulate; p.amax = nullptr; p.amax_in = argmax;
  const long long total = (long long)N * D * H * W * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (argmax) {
    if (dtype == SAP3D_BF16) maxpool_bwd_idx_kernel<bf16><<<grid_for(total), 256, 0, st>>>(p);
    else maxpool_bwd_idx_kernel<float><<<grid_for(total), 256, 0, st>>>(p);
  } else if (dtype == SAP3D_BF16) maxpool_bwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>(p);
  else maxpool_bwd_kernel<float><<<grid_for(total), 256, 0, st>>>(p);
  return check_launch("maxpool3d_bwd");
}

}  // extern "C"
