// SAGAN-style self-attention core of utils/network.py:184-186:
//     s = g . f^T  ;  beta = softmax(s, axis=-1)  ;  o = beta . h
// per sample, with g [Nq, dk], f [Nk, dk], h [Nk, dv] (row strides ldq/ldk/ldv so zero-padded channel
// layouts work).  Two implementations share the same beta buffer layout [B][Nq][ldb]:
//   * generic CUDA-core kernels (any shape, bf16 or f32): used for the tiny x_4_0 site (49 keys) and
//     the fp32 parity path;
//   * row softmax (+ backward) and transposes that glue the tcgen05 GEMMs (sap3d_gemm_nt / _tn) for
//     the large sites (3136 and 25088 queries).
// Replaces tf.matmul / tf.nn.softmax and their gradients.
#include <string.h>

#include <algorithm>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

struct AttnArgs {
  const void* q; const void* k; const void* v;  // g, f, h
  void* beta; void* o;
  const void* d_o; void* dq; void* dk; void* dv;
  int B, Nq, Nk, dk_, dv_;
  int ldq, ldk, ldv, ldb, ldo;
};

template <typename T>
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

// block per (b, q): scores, softmax -> beta[b][q][:]
template <typename T>
__global__ void __launch_bounds__(128) attn_scores_kernel(const AttnArgs p) {
  extern __shared__ float sm[];  // [dk] query + [8] scratch
  float* qs = sm;
  float* sh = sm + p.dk_;
  const int b = blockIdx.x / p.Nq, qi = blockIdx.x % p.Nq;
  const T* q = reinterpret_cast<const T*>(p.q) + ((long long)b * p.Nq + qi) * p.ldq;
  const T* k = reinterpret_cast<const T*>(p.k) + (long long)b * p.Nk * p.ldk;
  T* beta = reinterpret_cast<T*>(p.beta) + ((long long)b * p.Nq + qi) * p.ldb;
  for (int d = threadIdx.x; d < p.dk_; d += blockDim.x) qs[d] = to_f32<T>(q[d]);
  __syncthreads();
  float mx = -INFINITY;
  // scores are recomputed in the second pass (cheap) to avoid a [Nk] buffer for huge Nk
  for (int j = threadIdx.x; j < p.Nk; j += blockDim.x) {
    const T* kr = k + (long long)j * p.ldk;
    float s = 0.f;
    for (int d = 0; d < p.dk_; ++d) s = fmaf(qs[d], to_f32<T>(kr[d]), s);
    mx = fmaxf(mx, s);
  }
  mx = block_reduce<T>(mx, true, sh);
  float sum = 0.f;
  for (int j = threadIdx.x; j < p.Nk; j += blockDim.x) {
    const T* kr = k + (long long)j * p.ldk;
    float s = 0.f;
    for (int d = 0; d < p.dk_; ++d) s = fmaf(qs[d], to_f32<T>(kr[d]), s);
    const float e = __expf(s - mx);
    sum += e;
    beta[j] = from_f32<T>(e);  // un-normalised, fixed below
  }
  sum = block_reduce<T>(sum, false, sh);
  const float inv = 1.f / sum;
  for (int j = threadIdx.x; j < p.ldb; j += blockDim.x) beta[j] = j < p.Nk ? from_f32<T>(to_f32<T>(beta[j]) * inv) : from_f32<T>(0.f);
}

// block per (b, q): o[q][c] = sum_k beta[q][k] h[k][c]
template <typename T>
__global__ void __launch_bounds__(256) attn_pv_kernel(const AttnArgs p) {
  const int b = blockIdx.x / p.Nq, qi = blockIdx.x % p.Nq;
  const T* beta = reinterpret_cast<const T*>(p.beta) + ((long long)b * p.Nq + qi) * p.ldb;
  const T* v = reinterpret_cast<const T*>(p.v) + (long long)b * p.Nk * p.ldv;
  T* o = reinterpret_cast<T*>(p.o) + ((long long)b * p.Nq + qi) * p.ldo;
  for (int c = threadIdx.x; c < p.dv_; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < p.Nk; ++j) acc = fmaf(to_f32<T>(beta[j]), to_f32<T>(v[(long long)j * p.ldv + c]), acc);
    o[c] = from_f32<T>(acc);
  }
}

// block per (b, k): dv[k][c] = sum_q beta[q][k] dO[q][c]
template <typename T>
__global__ void __launch_bounds__(256) attn_dv_kernel(const AttnArgs p) {
  const int b = blockIdx.x / p.Nk, kj = blockIdx.x % p.Nk;
  const T* beta = reinterpret_cast<const T*>(p.beta) + (long long)b * p.Nq * p.ldb + kj;
  const T* d_o = reinterpret_cast<const T*>(p.d_o) + (long long)b * p.Nq * p.ldo;
  T* dv = reinterpret_cast<T*>(p.dv) + ((long long)b * p.Nk + kj) * p.ldv;
  for (int c = threadIdx.x; c < p.dv_; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < p.Nq; ++i) acc = fmaf(to_f32<T>(beta[(long long)i * p.ldb]), to_f32<T>(d_o[(long long)i * p.ldo + c]), acc);
    dv[c] = from_f32<T>(acc);
  }
}

// block per (b, q): dbeta = dO . h^T ; dS = beta * (dbeta - sum(dbeta*beta)) written over `ds`; dq = dS . f
template <typename T>
__global__ void __launch_bounds__(128) attn_ds_dq_kernel(const AttnArgs p, T* ds) {
  extern __shared__ float sm[];  // [dv] dO row + [8]
  float* dos = sm;
  float* sh = sm + p.dv_;
  const int b = blockIdx.x / p.Nq, qi = blockIdx.x % p.Nq;
  const T* d_o = reinterpret_cast<const T*>(p.d_o) + ((long long)b * p.Nq + qi) * p.ldo;
  const T* v = reinterpret_cast<const T*>(p.v) + (long long)b * p.Nk * p.ldv;
  const T* k = reinterpret_cast<const T*>(p.k) + (long long)b * p.Nk * p.ldk;
  const T* beta = reinterpret_cast<const T*>(p.beta) + ((long long)b * p.Nq + qi) * p.ldb;
  T* dsr = ds + ((long long)b * p.Nq + qi) * p.ldb;
  for (int c = threadIdx.x; c < p.dv_; c += blockDim.x) dos[c] = to_f32<T>(d_o[c]);
  __syncthreads();
  float dot = 0.f;
  for (int j = threadIdx.x; j < p.Nk; j += blockDim.x) {
    const T* vr = v + (long long)j * p.ldv;
    float db = 0.f;
    for (int c = 0; c < p.dv_; ++c) db = fmaf(dos[c], to_f32<T>(vr[c]), db);
    dot += db * to_f32<T>(beta[j]);
  }
  dot = block_reduce<T>(dot, false, sh);
  for (int j = threadIdx.x; j < p.ldb; j += blockDim.x) {
    float r = 0.f;
    if (j < p.Nk) {
      const T* vr = v + (long long)j * p.ldv;
      float db = 0.f;
      for (int c = 0; c < p.dv_; ++c) db = fmaf(dos[c], to_f32<T>(vr[c]), db);
      r = to_f32<T>(beta[j]) * (db - dot);
    }
    dsr[j] = from_f32<T>(r);
  }
  __syncthreads();
  T* dq = reinterpret_cast<T*>(p.dq) + ((long long)b * p.Nq + qi) * p.ldq;
  for (int d = threadIdx.x; d < p.dk_; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < p.Nk; ++j) acc = fmaf(to_f32<T>(dsr[j]), to_f32<T>(k[(long long)j * p.ldk + d]), acc);
    dq[d] = from_f32<T>(acc);
  }
}

// block per (b, k): dk[k][d] = sum_q dS[q][k] g[q][d]
template <typename T>
__global__ void __launch_bounds__(128) attn_dk_kernel(const AttnArgs p, const T* ds) {
  const int b = blockIdx.x / p.Nk, kj = blockIdx.x % p.Nk;
  const T* dsc = ds + (long long)b * p.Nq * p.ldb + kj;
  const T* q = reinterpret_cast<const T*>(p.q) + (long long)b * p.Nq * p.ldq;
  T* dk = reinterpret_cast<T*>(p.dk) + ((long long)b * p.Nk + kj) * p.ldk;
  for (int d = threadIdx.x; d < p.dk_; d += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < p.Nq; ++i) acc = fmaf(to_f32<T>(dsc[(long long)i * p.ldb]), to_f32<T>(q[(long long)i * p.ldq + d]), acc);
    dk[d] = from_f32<T>(acc);
  }
}

// ---- row softmax over f32 / bf16 logits -> bf16/f32 probabilities (TC path glue) ------------------
// one CTA per row; the row lives in registers (<= 16 values per thread for rows up to 4096 columns), so the
// logits are read from HBM exactly once.  Longer rows fall back to re-reading (L2-resident) data.
constexpr int SM_MAXV = 16;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const TI* __restrict__ s, TO* __restrict__ out, int cols, int ld_in, int ld_out) {
  __shared__ float sh[8];
  const TI* r = s + (long long)blockIdx.x * ld_in;
  TO* o = out + (long long)blockIdx.x * ld_out;
  const bool in_regs = cols <= SM_MAXV * 256;
  float v[SM_MAXV];
  float mx = -INFINITY;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
      const int j = threadIdx.x + i * 256;
      v[i] = j < cols ? to_f32<TI>(r[j]) : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) mx = fmaxf(mx, to_f32<TI>(r[j]));
  }
  mx = block_reduce<TI>(mx, true, sh);
  float sum = 0.f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
      v[i] = __expf(v[i] - mx);  // exp(-inf) = 0 for the padding
      sum += v[i];
    }
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) sum += __expf(to_f32<TI>(r[j]) - mx);
  }
  sum = block_reduce<TI>(sum, false, sh);
  const float inv = 1.f / sum;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
      const int j = threadIdx.x + i * 256;
      if (j < ld_out) o[j] = from_f32<TO>(j < cols ? v[i] * inv : 0.f);
    }
  } else {
    for (int j = threadIdx.x; j < ld_out; j += blockDim.x) o[j] = from_f32<TO>(j < cols ? __expf(to_f32<TI>(r[j]) - mx) * inv : 0.f);
  }
}

// dS = beta * (dbeta - sum_j dbeta_j beta_j), in place over dbeta
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const T* __restrict__ beta, T* __restrict__ dbeta, int cols, int ld) {
  __shared__ float sh[8];
  const T* br = beta + (long long)blockIdx.x * ld;
  T* dr = dbeta + (long long)blockIdx.x * ld;
  const bool in_regs = cols <= SM_MAXV * 256;
  float b[SM_MAXV], d[SM_MAXV];
  float dot = 0.f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
      const int j = threadIdx.x + i * 256;
      b[i] = j < cols ? to_f32<T>(br[j]) : 0.f;
      d[i] = j < cols ? to_f32<T>(dr[j]) : 0.f;
      dot = fmaf(b[i], d[i], dot);
    }
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) dot += to_f32<T>(br[j]) * to_f32<T>(dr[j]);
  }
  dot = block_reduce<T>(dot, false, sh);
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
      const int j = threadIdx.x + i * 256;
      if (j < ld) dr[j] = from_f32<T>(j < cols ? b[i] * (d[i] - dot) : 0.f);
    }
  } else {
    for (int j = threadIdx.x; j < ld; j += blockDim.x) dr[j] = from_f32<T>(j < cols ? to_f32<T>(br[j]) * (to_f32<T>(dr[j]) - dot) : 0.f);
  }
}

// batched transpose [R][C] (row stride ld_in) -> [C][R] (row stride ld_out, zero padded up to ld_out)
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int Cc, int ld_in, int ld_out,
                                 long long bs_in, long long bs_out) {
  __shared__ T tile[32][33];
  const T* ib = in + blockIdx.z * bs_in;
  T* ob = out + blockIdx.z * bs_out;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? ib[(long long)r * ld_in + c] : from_f32<T>(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < ld_out) ob[(long long)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

extern "C" {

int sap3d_attention_fwd(int32_t dtype, const void* g, const void* f, const void* h, void* beta, void* o, int32_t B, int32_t Nq,
                        int32_t Nk, int32_t dk, int32_t dv, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldb, int32_t ldo,
                        void* stream) {
  if (require_device()) return 1;
  AttnArgs p;
  memset(&p, 0, sizeof(p));
  p.q = g; p.k = f; p.v = h; p.beta = beta; p.o = o;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.dk_ = dk; p.dv_ = dv;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldb = ldb; p.ldo = ldo;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t sm = (size_t)(dk + 8) * sizeof(float);
  if (dtype == SAP3D_BF16) {
    attn_scores_kernel<bf16><<<B * Nq, 128, sm, st>>>(p);
    attn_pv_kernel<bf16><<<B * Nq, 256, 0, st>>>(p);
  } else {
    attn_scores_kernel<float><<<B * Nq, 128, sm, st>>>(p);
    attn_pv_kernel<float><<<B * Nq, 256, 0, st>>>(p);
  }
  return check_launch("attention_fwd");
}

/* ds: scratch with the shape of beta */
int sap3d_attention_bwd(int32_t dtype, const void* g, const void* f, const void* h, const void* beta, const void* d_o, void* ds,
                        void* dg, void* df, void* dh, int32_t B, int32_t Nq, int32_t Nk, int32_t dk, int32_t dv, int32_t ldq,
                        int32_t ldk, int32_t ldv, int32_t ldb, int32_t ldo, void* stream) {
  if (require_device()) return 1;
  AttnArgs p;
  memset(&p, 0, sizeof(p));
  p.q = g; p.k = f; p.v = h; p.beta = const_cast<void*>(beta); p.d_o = d_o; p.dq = dg; p.dk = df; p.dv = dh;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.dk_ = dk; p.dv_ = dv;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldb = ldb; p.ldo = ldo;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t sm = (size_t)(dv + 8) * sizeof(float);
  if (dtype == SAP3D_BF16) {
    attn_dv_kernel<bf16><<<B * Nk, 256, 0, st>>>(p);
    attn_ds_dq_kernel<bf16><<<B * Nq, 128, sm, st>>>(p, reinterpret_cast<bf16*>(ds));
    attn_dk_kernel<bf16><<<B * Nk, 128, 0, st>>>(p, reinterpret_cast<const bf16*>(ds));
  } else {
    attn_dv_kernel<float><<<B * Nk, 256, 0, st>>>(p);
    attn_ds_dq_kernel<float><<<B * Nq, 128, sm, st>>>(p, reinterpret_cast<float*>(ds));
    attn_dk_kernel<float><<<B * Nk, 128, 0, st>>>(p, reinterpret_cast<const float*>(ds));
  }
  return check_launch("attention_bwd");
}

int sap3d_softmax_rows(int32_t in_dtype, const void* logits, void* probs_bf16, int64_t rows, int32_t cols, int32_t ld_in,
                       int32_t ld_out, void* stream) {
  if (require_device()) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (in_dtype == SAP3D_F32)
    softmax_rows_kernel<float, bf16><<<(unsigned)rows, 256, 0, st>>>(reinterpret_cast<const float*>(logits), reinterpret_cast<bf16*>(probs_bf16), cols, ld_in, ld_out);
  else
    softmax_rows_kernel<bf16, bf16><<<(unsigned)rows, 256, 0, st>>>(reinterpret_cast<const bf16*>(logits), reinterpret_cast<bf16*>(probs_bf16), cols, ld_in, ld_out);
  return check_launch("softmax_rows");
}

int sap3d_softmax_bwd_rows(const void* beta_bf16, void* dbeta_bf16, int64_t rows, int32_t cols, int32_t ld, void* stream) {
  if (require_device()) return 1;
  softmax_bwd_rows_kernel<bf16><<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(beta_bf16), reinterpret_cast<bf16*>(dbeta_bf16), cols, ld);
  return check_launch("softmax_bwd_rows");
}

int sap3d_transpose(int32_t dtype, const void* in, void* out, int32_t batch, int32_t R, int32_t Cc, int32_t ld_in, int32_t ld_out,
                    int64_t bs_in, int64_t bs_out, void* stream) {
  if (require_device()) return 1;
  dim3 grid((Cc + 31) / 32, (std::max(R, ld_out) + 31) / 32, batch), block(32, 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SAP3D_BF16)
    transpose_kernel<bf16><<<grid, block, 0, st>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), R, Cc, ld_in, ld_out, bs_in, bs_out);
  else
    transpose_kernel<float><<<grid, block, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), R, Cc, ld_in, ld_out, bs_in, bs_out);
  return check_launch("transpose");
}

}  // extern "C"
