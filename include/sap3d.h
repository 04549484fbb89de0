/*
 * sap3d.h — C ABI of the B200-native P3D saliency hot path (libsap3d_b200.so).
 *
 * This is the drop-in boundary for the hot path of A-Nasiri-M/sap3d_tensorflow: every entry point
 * replaces one class of TensorFlow op call-sites of the reference's graph builders (the reference
 * has no native code; its "FFI" is the set of tf.* calls in p3d.py / gn/p3d_gn.py / utils/network.py
 * / utils/metrics.py).  A TensorFlow custom-op shim (tf_ops/sap3d_tf_ops.cc) or the ctypes binding
 * (sap3d_tensorflow_b200/_abi.py) forwards to these functions.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / TF types.  All tensor pointers are DEVICE pointers.
 *   - activations are NDHWC (channels innermost), dtype SAP3D_BF16 or SAP3D_F32 (desc->dtype);
 *     parameters, statistics and gradients of parameters are fp32.
 *   - the caller owns every buffer (TF allocator / torch allocator); no hidden allocation, no hidden
 *     synchronisation; work is enqueued on the cudaStream_t passed as `stream` (void*).
 *   - re-entrant: no global mutable state besides a thread-local error string.
 *   - return value 0 = OK, non-zero = error (message via sap3d_last_error()).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SAP3D_H_
#define SAP3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAP3D_BF16 0
#define SAP3D_F32 1

#define SAP3D_IMPL_AUTO 0
#define SAP3D_IMPL_SIMT 1 /* CUDA-core implicit GEMM (any shape, bf16 or f32 storage) */
#define SAP3D_IMPL_TC 2   /* tcgen05/TMEM implicit GEMM fed by TMA (bf16, Cin % 64 == 0) */

const char* sap3d_last_error(void);
int sap3d_abi_version(void);
/* 1 when a CUDA device of compute capability 10.x is usable */
int sap3d_device_ok(void);
/* developer probe, not a reference call-site: per-CTA phase time stamps of the tcgen05 convolution kernel are written to
 * buf ([cta][16][2] uint64 device memory: clock64, globaltimer); NULL switches the probe off.  Process-global. */
int sap3d_debug_conv_timing(void* buf);
/* developer probe: how many convolution launches of this process took the halo-tile kernel (conv_tc.cu) so far; tests use it
 * to prove that a case meant for that kernel did not silently take another one. */
long long sap3d_debug_conv_halo_launches(void);
/* ... and how many of those took its swapped-operand form (output channels on the M side, 256 positions per instruction) */
long long sap3d_debug_conv_swap_launches(void);

/* ------------------------------------------------------------------------------------------------
 * Convolution family.  Replaces tf.nn.conv3d (p3d.py:19,24,86,112,125,343), tf.nn.bias_add
 * (p3d.py:19,24), tf.layers.conv3d (utils/network.py:101,164-178,188,262),
 * tf.layers.conv3d_transpose (utils/network.py:107; p3d.py:393) and tf.concat feeding a conv
 * (utils/network.py:97; the concat is fused as K-segments and never materialised), plus their
 * gradients (tf.gradients via AdamOptimizer.minimize, train.py:168).
 *
 * Geometry: input [N,D,H,W,cin_total] with cin_total = sum(cin[0..nseg)), filter TF layout
 * DHWIO [kd,kh,kw,cin_total,cout] for a conv, [kd,kh,kw,cout,cin_total] for a transposed conv
 * (tf.layers.conv3d_transpose kernel layout), padding = TF 'SAME' (conv) / 'same' (transpose):
 *   conv      : O = ceil(I/s), pad_before = max((O-1)s+k-I,0)/2 (extra padding at the end)
 *   transpose : O = I*s, y[p] = sum x[i] w[k], p = i*s + k - pb, pb = max(k-s,0)/2
 * ---------------------------------------------------------------------------------------------- */
typedef struct sap3d_conv_desc {
  int32_t dtype;      /* storage type of x / y / dy / dx */
  int32_t impl;       /* SAP3D_IMPL_* */
  int32_t N, D, H, W; /* input extent */
  int32_t nseg;       /* 1 or 2 channel segments (fused concat) */
  int32_t cin[2];
  int32_t cout;
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t transposed;
  int32_t has_bias;
  int32_t out_f32;    /* forward output stored as f32 regardless of dtype (attention logits) */
} sap3d_conv_desc;

/* output extent of the conv described by d (writes 3 ints: Do,Ho,Wo) */
int sap3d_conv_out_dims(const sap3d_conv_desc* d, int32_t* out_dhw);
/* number of rows of the per-tile statistics buffer conv_fwd writes (stats is [rows][2][cout] f32) */
int sap3d_conv_stats_rows(const sap3d_conv_desc* d);
/* element count (bf16) of the operand buffers of the tcgen05 path; which: 0 = forward (packed filter; for small-Cin
 * convolutions such as the Cin = 3 stem also the im2col matrix and a scratch block), 1 = data-gradient */
size_t sap3d_conv_packed_elems(const sap3d_conv_desc* d, int32_t which);

/* w_tf (fp32, TF layout) -> bf16 K-major matrices for the tensor-core path:
 *   w_fwd  [cout_pad][taps*cin_total]              (B operand of forward)
 *   w_dgrad[cin_total_pad][taps*cout]              (B operand of data gradient)
 * either output may be NULL. */
int sap3d_conv_pack_weights(const sap3d_conv_desc* d, const float* w_tf, void* w_fwd, void* w_dgrad, void* stream);

/* y = conv(x0 ++ x1, w) (+ bias).  stats (nullable): per-tile partial sum / sum-of-squares of the
 * fp32 results, [stats_rows][2][cout] f32, reduced by sap3d_bn_finalize. */
int sap3d_conv_fwd(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf,
                   const void* w_fwd_packed, const float* bias, void* y, float* stats, void* stream);
/* inference-mode fusion: y = relu?((conv(x0 ++ x1, w) + bias) * scale[c] + shift[c]) — tf.layers.batch_normalization with
 * training=False (moving statistics; scale/shift from sap3d_bn_finalize(training = 0)) and tf.nn.relu folded into the conv
 * epilogue (utils/network.py:100-110 with training=False, p3d.py:343-345).  Tensor-core paths only: query with
 * sap3d_conv_fwd_on_tensor_cores(). */
int sap3d_conv_fwd_on_tensor_cores(const sap3d_conv_desc* d);
/* 1 when the forward operand buffer is a workspace sap3d_conv_fwd fills itself (the Cin = 3 stem's im2col form: packed filter +
 * scratch + im2col matrix, which sap3d_conv_wgrad reuses through `fwd_operand`), 0 when it is the pre-packed filter that
 * conv_fwd only reads.  A binding that must not write its inputs (TF custom op) allocates that workspace as an output. */
int sap3d_conv_fwd_operand_is_workspace(const sap3d_conv_desc* d);
int sap3d_conv_fwd_affine(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                          const float* bias, const float* scale, const float* shift, int32_t relu, void* y, void* stream);

/* Training graphs: tf.nn.conv3d followed by tf.layers.batch_normalization(training=True) (+ tf.nn.relu, + the bottleneck's residual
 * add and its ReLU; p3d.py:83-136, make_block p3d.py:143-166) as ONE launch.  The convolution writes its raw output and the
 * per-tile statistics rows as sap3d_conv_fwd does (the backward pass reads them); then every CTA of the launch waits at a grid
 * barrier, finalises the batch statistics of its own columns exactly as sap3d_bn_finalize would (scale / shift / mean / rstd
 * and the moving averages are written once per channel) and writes
 *     y = relu_out?( relu1?(raw * scale + shift) + residual )
 * from the accumulators it still holds.  Only for problems whose CTAs are all resident at once (the backbone's stage-2/3 layers):
 * ask sap3d_conv_fwd_bn_supported() first (1 = yes; needs a device).  Measured on B200: one launch less per layer but ~1 us
 * slower per layer than sap3d_conv_fwd + sap3d_bn_apply_fused, so this repo's engine uses it only on request. */
typedef struct sap3d_bn_fuse {
  const float* gamma;      /* [cout] */
  const float* beta;
  float* moving_mean;      /* nullable; updated with `momentum` (1.0 leaves them unchanged) */
  float* moving_var;
  float momentum, eps;
  float* scale;            /* outputs, [cout] each */
  float* shift;
  float* mean;             /* nullable */
  float* rstd;             /* nullable */
  int32_t relu1;
  const void* residual;    /* nullable, bf16, same shape as the output */
  int32_t relu_out;
  void* y;                 /* normalised output, bf16 */
} sap3d_bn_fuse;
int sap3d_conv_fwd_bn_supported(const sap3d_conv_desc* d);
int sap3d_conv_fwd_bn(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                      const float* bias, void* raw, float* stats, const sap3d_bn_fuse* f, void* stream);
/* dx_seg = data gradient w.r.t. segment `seg`; accumulate != 0 adds into dx */
int sap3d_conv_dgrad(const sap3d_conv_desc* d, int32_t seg, const void* dy, const float* w_tf,
                     const void* w_dgrad_packed, void* dx, int32_t accumulate, void* stream);
/* both segment gradients of a fused-concat conv (two equal, 64-aligned segments, tensor-core path) in ONE launch: dy is read
 * once per tap instead of once per segment */
int sap3d_conv_dgrad2_supported(const sap3d_conv_desc* d);
int sap3d_conv_dgrad2(const sap3d_conv_desc* d, const void* dy, const float* w_tf, const void* w_dgrad_packed, void* dx0,
                      int32_t accumulate0, void* dx1, int32_t accumulate1, void* stream);
/* dw (fp32, TF layout) += filter gradient; db (nullable, [cout]) += bias gradient.
 * The caller zeroes dw/db at the start of a step (gradients accumulate across calls). */
/* fwd_operand (nullable): the packed forward operand buffer given to sap3d_conv_fwd; small-Cin convolutions (the stem)
 * keep their im2col matrix there and reuse it for the filter gradient. */
int sap3d_conv_wgrad(const sap3d_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw,
                     float* db, const void* fwd_operand, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation.  Replaces tf.layers.batch_normalization (p3d.py:58-127,344; utils/network.py:91),
 * GroupNorm (utils/network.py:65-87 == gn/p3d_gn.py:24-46), tf.nn.relu and the residual adds
 * (p3d.py:72,81,133-134).
 * ---------------------------------------------------------------------------------------------- */
/* stats [rows][2][C] (from conv_fwd) -> scale/shift so that y = x*scale+shift is the normalised
 * tensor.  training != 0: biased batch variance over `count` positions, moving averages updated in
 * place with `momentum` (TF: 0.99); training == 0: the moving statistics are used (stats ignored).
 * save_mean/save_rstd (nullable) are kept for the backward pass. */
int sap3d_bn_finalize(const float* stats, int32_t rows, int32_t C, double count, const float* gamma, const float* beta,
                      float* moving_mean, float* moving_var, int32_t training, float momentum, float eps, float* scale,
                      float* shift, float* save_mean, float* save_rstd, void* stream);
/* GroupNorm statistics of x [N][S][C] (G groups): per-(sample,channel) scale/shift [N][C] and
 * per-(sample,group) mean/rstd [N][G]. */
int sap3d_gn_stats(int32_t dtype, const void* x, int32_t N, int64_t S, int32_t C, int32_t G, const float* gamma,
                   const float* beta, float eps, float* scale, float* shift, float* save_mean, float* save_rstd, void* stream);
/* y = relu_out?( relu1?(a*s1+t1) + relu2?(b*s2+t2) ); b, s1, s2 nullable (identity).  scale index is
 * c (positions_per_sample == 0) or n*C + c. */
int sap3d_affine_act(int32_t dtype, const void* a, const float* s1, const float* t1, int32_t relu1, const void* b,
                     const float* s2, const float* t2, int32_t relu2, int32_t relu_out, void* y, int64_t P, int32_t C,
                     int64_t positions_per_sample, void* stream);
/* sap3d_bn_finalize (for one or two norms) fused into sap3d_affine_act: y = relu_out?( relu1?(BN1(a)) + relu2?(BN2(b) | b) )
 * in ONE launch, for layers whose statistics buffers have few rows (the backbone).  scale/shift/mean/rstd are still
 * published for the backward pass; moving averages are updated when training. */
int sap3d_bn_apply_fused(int32_t dtype, const void* a, const float* stats1, int32_t rows1, const float* gamma1, const float* beta1,
                         float* mm1, float* mv1, int32_t training1, float* scale1, float* shift1, float* mean1, float* rstd1,
                         int32_t relu1, const void* b, int32_t has_norm2, const float* stats2, int32_t rows2, const float* gamma2,
                         const float* beta2, float* mm2, float* mv2, int32_t training2, float* scale2, float* shift2, float* mean2,
                         float* rstd2, int32_t relu2, int32_t relu_out, void* y, int64_t P, int32_t C, double count, float momentum,
                         float eps, void* stream);
size_t sap3d_affine_act_bwd_workspace(int32_t C);
/* backward of sap3d_affine_act for per-channel statistics.  mean/rstd non-NULL => that branch is a
 * batch-statistics BatchNorm (full BN backward); NULL => frozen scale.  da/db nullable; d{gamma,beta}
 * (nullable) are accumulated (+=). */
int sap3d_affine_act_bwd(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                         const float* rstd1, int32_t relu1, const void* b, const float* s2, const float* t2,
                         const float* mean2, const float* rstd2, int32_t relu2, int32_t relu_out, int64_t P, int32_t C,
                         void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1, float* dbeta1, float* dgamma2,
                         float* dbeta2, void* workspace, void* stream);
/* the same backward in two phases, for BatchNorm statistics that span the replicas of a data-parallel job (the
 * reference normalises over its whole single-device batch, p3d.py:344 / tf.layers.batch_normalization).
 * phase 1: reductions only -- d{gamma,beta} += LOCAL sums, and workspace[0 .. 4*C) floats = LOCAL sums / global_count;
 * the caller sums those 4*C floats over the replicas (the forward pass did the same with the statistics rows, finalised by
 * sap3d_bn_finalize with the global count); phase 2: applies the data gradient from that workspace. */
int sap3d_affine_act_bwd_sync(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                              const float* rstd1, int32_t relu1, const void* b, const float* s2, const float* t2,
                              const float* mean2, const float* rstd2, int32_t relu2, int32_t relu_out, int64_t P, int32_t C,
                              void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1, float* dbeta1, float* dgamma2,
                              float* dbeta2, void* workspace, void* stream, double global_count, int32_t phase);

/* ------------------------------------------------------------------------------------------------
 * Pooling.  Replaces tf.nn.max_pool3d (p3d.py:347-348,354,360,366) and tf.layers.max_pooling3d
 * (utils/network.py:6-7).  same != 0: TF 'SAME' (padding never wins), else 'VALID'.
 * argmax (nullable, uint8 [outputs][C]): the forward pass records the window-local index of the FIRST maximum; the
 * backward pass then gathers (dy, index) pairs instead of re-scanning the windows (the gradient goes to that element).
 * ---------------------------------------------------------------------------------------------- */
int sap3d_maxpool3d_out_dims(int32_t D, int32_t H, int32_t W, const int32_t* ksize, const int32_t* strides, int32_t same,
                             int32_t* out_dhw);
int sap3d_maxpool3d_fwd(int32_t dtype, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                        const int32_t* ksize, const int32_t* strides, int32_t same, void* y, uint8_t* argmax, void* stream);
int sap3d_maxpool3d_bwd(int32_t dtype, const void* x, const void* dy, int32_t N, int32_t D, int32_t H, int32_t W,
                        int32_t C, const int32_t* ksize, const int32_t* strides, int32_t same, const uint8_t* argmax,
                        void* dx, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder head, loss, dropout, attention gate, optimizer.
 * ---------------------------------------------------------------------------------------------- */
/* logits = conv3d_transpose(x, w[kd,kh,kw,1,C], stride, 'same') + bias; pred = sigmoid(logits) (nullable).
 * Replaces p3d.py:393,397 (tf.layers.conv3d_transpose -> 1 channel, tf.sigmoid). */
int sap3d_head_fwd(int32_t dtype, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, const int32_t* ksize,
                   int32_t stride, const float* w, const float* bias, float* logits, float* pred, void* stream);
int sap3d_head_bwd(int32_t dtype, const float* dlogits, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                   const int32_t* ksize, int32_t stride, const float* w, void* dx, int32_t accumulate, float* dw, void* stream);
/* tensor-core form of the k3 s2 head for bf16 activations with C % 64 == 0: [positions x C] x [C x 27] GEMM + col2im gather
 * (forward), GEMMs over the im2col of dlogits (backward).  workspace: sap3d_head_tc_workspace(N,D,H,W,C) bytes, shared by
 * forward and backward of one head. */
size_t sap3d_head_tc_workspace(int32_t N, int32_t D, int32_t H, int32_t W, int32_t C);
int sap3d_head_tc_fwd(const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, const float* w, const float* bias,
                      float* logits, float* pred, void* workspace, void* stream);
int sap3d_head_tc_bwd(const float* dlogits, const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, void* dx,
                      int32_t accumulate, float* dw, void* workspace, void* stream, void* wgrad_stream);
/* smooth_l1_loss(sigma=1, weights 1) summed over all elements (utils/network.py:49-62, train.py:159):
 * loss_sum[0] += loss; dlogits = dLoss/dlogits; dbias[0] += sum(dlogits); pred = sigmoid(logits). */
int sap3d_loss_smooth_l1(const float* logits, const float* target, int64_t n, int32_t apply_sigmoid, float* pred,
                         float* dlogits, double* loss_sum, float* dbias, void* stream);
/* the general form of utils/network.py:49-62: in = inside_weight * (pred - target); per element
 * |in| < 1/sigma^2 ? in^2 sigma^2 / 2 : |in| - 0.5/sigma^2, times outside_weight, summed over all elements. */
int sap3d_loss_smooth_l1_ex(const float* logits, const float* target, int64_t n, int32_t apply_sigmoid, float* pred,
                            float* dlogits, double* loss_sum, float* dbias, float sigma, float inside_weight, float outside_weight,
                            void* stream);
/* tf.layers.dropout (p3d.py:392): y = x*keep/(1-rate), keep = splitmix64(base_seed + *step, index) >= rate.
 * The backward pass is the same call on dy. */
int sap3d_dropout(int32_t dtype, const void* x, void* y, int64_t n, float rate, uint64_t base_seed, const int32_t* step,
                  int32_t accumulate, void* stream);
/* x = o*gamma + x  (utils/network.py:191-192) */
int sap3d_gate_fwd(int32_t dtype, const void* o, const void* x, const float* gamma, void* y, int64_t n, void* stream);
int sap3d_gate_bwd(int32_t dtype, const void* dy, const void* o, const float* gamma, void* d_o, void* dx, int32_t acc_x,
                   float* dgamma, int64_t n, void* stream);
/* ------------------------------------------------------------------------------------------------
 * Self-attention core (utils/network.py:184-186: tf.matmul, tf.nn.softmax, tf.matmul) and the GEMM /
 * softmax / transpose pieces used to run it on the tensor cores.
 * ---------------------------------------------------------------------------------------------- */
/* generic CUDA-core path: beta[b] = softmax(g[b] f[b]^T) ([B][Nq][ldb], zero padded), o[b] = beta[b] h[b] */
int sap3d_attention_fwd(int32_t dtype, const void* g, const void* f, const void* h, void* beta, void* o, int32_t B, int32_t Nq,
                        int32_t Nk, int32_t dk, int32_t dv, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldb, int32_t ldo,
                        void* stream);
int sap3d_attention_bwd(int32_t dtype, const void* g, const void* f, const void* h, const void* beta, const void* d_o, void* ds,
                        void* dg, void* df, void* dh, int32_t B, int32_t Nq, int32_t Nk, int32_t dk, int32_t dv, int32_t ldq,
                        int32_t ldk, int32_t ldv, int32_t ldb, int32_t ldo, void* stream);
/* fused (flash-style) tcgen05 path: o[b] = softmax(q[b] k[b]^T) v[b] without materialising the [Nq][Nk] scores.
 * q [B][Nq][64], k [B][Nk][64] (d_k zero-padded to 64), v [B][Nk][dv], o [B][Nq][dv], all bf16 and dense; dv in {128, 256};
 * lse (nullable) [B][Nq] = log-sum-exp of every score row, kept for the backward kernels. */
int sap3d_flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t Nq, int32_t Nk,
                         int32_t dk, int32_t dv, void* stream);
/* backward of sap3d_flash_attn_fwd (dv == 128): dq [B][Nq][64], dk [B][Nk][64], dv [B][Nk][dv] (bf16) from o, d_o and lse;
 * the probabilities are recomputed on chip.  workspace: sap3d_flash_attn_bwd_workspace(B, Nq, Nk, dv) bytes. */
size_t sap3d_flash_attn_bwd_workspace(int32_t B, int32_t Nq, int32_t Nk, int32_t dv);
int sap3d_flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, void* dq,
                         void* dk, void* dv_out, int32_t B, int32_t Nq, int32_t Nk, int32_t dk_dim, int32_t dv, void* workspace,
                         void* stream);
/* bf16 tensor-core GEMMs: C[M][N] (+)= A[M][K] B[N][K]^T  (K % 64 == 0, ldb == K, N % 8 == 0); B has rows_b <= N
 * rows, the remaining output columns are computed against zeros */
int sap3d_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t rows_b, void* C, int64_t ldc, int32_t M,
                  int32_t N, int32_t K, int32_t out_f32, int32_t accumulate, void* stream);
/* `batch` independent products in ONE launch: sample n reads A + n*stride_a and B + n*stride_b and writes C + n*stride_c (element
 * strides, multiples of 8).  Replaces the per-sample loops over tf.matmul(g, f, transpose_b=True) / tf.matmul(beta, h) of
 * utils/network.py:176-180 for the attention blocks that are not on the fused path. */
int sap3d_gemm_nt_batched(const void* A, int64_t lda, int64_t stride_a, const void* B, int64_t ldb, int64_t stride_b, int32_t rows_b,
                          void* C, int64_t ldc, int64_t stride_c, int32_t M, int32_t N, int32_t K, int32_t batch, int32_t out_f32,
                          int32_t accumulate, void* stream);
/* D[M][N] (fp32) += sum_pos P[pos][m] Q[pos][n]  (M, N % 64 == 0); the caller zeroes D */
int sap3d_gemm_tn(const void* P, int64_t ldp, const void* Q, int64_t ldq, float* D, int64_t ldd, int32_t M, int32_t N,
                  int32_t Kpos, void* stream);
int sap3d_softmax_rows(int32_t in_dtype, const void* logits, void* probs_bf16, int64_t rows, int32_t cols, int32_t ld_in,
                       int32_t ld_out, void* stream);
/* dbeta <- beta * (dbeta - rowsum(dbeta * beta)), bf16 in place */
int sap3d_softmax_bwd_rows(const void* beta_bf16, void* dbeta_bf16, int64_t rows, int32_t cols, int32_t ld, void* stream);
int sap3d_transpose(int32_t dtype, const void* in, void* out, int32_t batch, int32_t R, int32_t C, int32_t ld_in, int32_t ld_out,
                    int64_t bs_in, int64_t bs_out, void* stream);
/* [P][c] -> [P][c_pad] zero padded (unpad == 0) or back (unpad != 0, optionally accumulating) */
int sap3d_pad_channels(int32_t dtype, const void* in, void* out, int64_t P, int32_t c, int32_t c_pad, int32_t unpad,
                       int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Saliency metrics (utils/metrics.py CC :227, SIM :258, NSS :200, KLdiv :338): out[n][4] = CC, SIM, NSS, KLdiv
 * (fp64) for n (prediction, density, fixation) map triples; fixation may be NULL (NSS = NaN).
 * ---------------------------------------------------------------------------------------------- */
int sap3d_saliency_metrics(const float* pred, const float* density, const float* fixation, int32_t n_maps,
                           int64_t elems_per_map, int64_t pred_stride, int64_t density_stride, int64_t fixation_stride,
                           double* out, void* stream);

/* Test-time evaluation (test.py:164-183): cv2.resize(map, (W, H)) with INTER_LINEAR semantics for n float32 maps, and the
 * fixation-based AUC_Judd / AUC_Borji (utils/metrics.py:25-154): out[n][2].  jitter != 0 adds a counter-hash jitter of
 * 1e-7 (the reference draws np.random noise); AUC_Borji's random locations are splitmix64(seed, fixation, repetition) % elems
 * (the reference accepts any rand_sampler).  workspace: sap3d_saliency_auc_workspace(n_maps, n_rep) bytes. */
/* Input preprocessing of dataflow.py:194-209 / gen_pred.py:113-118 for n decoded BGR uint8 frames [n][h][w][3] (device):
 * RGB order, minus mean_rgb (HOST pointer to 3 floats: [90, 102, 98]), cv2.resize to W x H on the float image, / 255;
 * dst [n][H][W][3] in out_dtype (SAP3D_F32 / SAP3D_BF16) = one frame of the NDHWC network input each. */
int sap3d_preprocess_frames(const uint8_t* bgr, int32_t n, int32_t h, int32_t w, const float* mean_rgb_host, int32_t out_dtype,
                            void* dst, int32_t H, int32_t W, void* stream);
/* NaN-filtered column sums and counts of values [n][m] (fp64): the per-metric (sum, count) pairs of test.py:177-181 */
int sap3d_nan_sum_count(const double* values, int32_t n, int32_t m, double* sums, double* counts, void* stream);
int sap3d_resize_bilinear(const float* src, int32_t n, int32_t h, int32_t w, float* dst, int32_t H, int32_t W, void* stream);
size_t sap3d_saliency_auc_workspace(int32_t n_maps, int32_t n_rep);
int sap3d_saliency_auc(const float* sal, const float* fix, int32_t n_maps, int64_t elems, int32_t jitter, int32_t n_rep, double step,
                       uint64_t seed, double* out, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm + CBAM of the GN model variant (gn/p3d_gn.py:24-46,175; utils/network.py:65-87,198-274).
 * ---------------------------------------------------------------------------------------------- */
int sap3d_sample_stats_rows(int64_t S, int32_t C, int32_t N);
/* part [N][rows][3][C] = per-sample per-channel (sum, sum of squares, max) partials of x [N][S][C] (optionally of
 * x*scale[n][c]) */
int sap3d_sample_channel_partials(int32_t dtype, const void* x, const float* scale, int32_t N, int64_t S, int32_t C, int32_t rows,
                                  float* part, void* stream);
/* Per-clip batch-statistics BatchNorm (+ReLU, + second operand with or without its own per-clip norm) in ONE launch:
 *     y = relu_out?( relu1?(bn_clip(a)) + relu2?(bn_clip(b) | b) ),   statistics over the S positions of each of the N clips.
 * tf.layers.batch_normalization(training=True) as gen_pred.py sees it -- one 16-frame window per sess.run, so every window is
 * normalised on its own (gen_pred.py:88-135) -- for a BATCH of windows; replaces the three launches
 * sap3d_sample_channel_partials + sap3d_gn_finalize + sap3d_affine_act for tensors whose (clip, 64-channel) slab is at most
 * 128 KB (sap3d_sample_norm_apply_supported).  gamma2 == NULL: b is added as it is.  Moving averages are not touched. */
int sap3d_sample_norm_apply_supported(int32_t dtype, int64_t S, int32_t C);
int sap3d_sample_norm_apply(int32_t dtype, const void* a, const float* gamma1, const float* beta1, int32_t relu1, const void* b,
                            const float* gamma2, const float* beta2, int32_t relu2, int32_t relu_out, void* y, int32_t N, int64_t S,
                            int32_t C, float eps, void* stream);
/* GroupNorm: partials -> per-(sample,channel) scale/shift [N][C] (+ per-(sample,group) mean/rstd [N][G]) */
int sap3d_gn_finalize(const float* part, int32_t rows, int32_t N, int64_t S, int32_t C, int32_t G, const float* gamma,
                      const float* beta, float eps, float* scale, float* shift, float* save_mean, float* save_rstd, void* stream);
/* CBAM attention maps of r: channel scale cscale [N][C] (mean & max over D,H,W -> shared MLP C->hidden->C -> sigmoid) and
 * spatial map att [N][D*H*W] (mean & max over C of r*cscale -> 7x7x7 conv, no bias -> sigmoid); scratch: part, sp;
 * save (nullable) [N][2C + 2 hidden] keeps (avg, max, relu hidden(avg), relu hidden(max)) for sap3d_cbam_tail_bwd */
int sap3d_cbam_fwd(int32_t dtype, const void* r, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t hidden,
                   const float* w0, const float* b0, const float* w1, const float* b1, const float* w_sp, float* part,
                   int32_t rows, float* cscale, float* sp, float* att, float* save, void* stream);
/* y = relu((a*s1[n][c]+t1[n][c]) + r*cscale[n][c]*att[n][pos])  (out += cbam(residual); relu — gn/p3d_gn.py:175-177) */
int sap3d_cbam_merge(int32_t dtype, const void* a, const float* s1, const float* t1, const void* r, const float* cscale,
                     const float* att, void* y, int32_t N, int64_t S, int32_t C, void* stream);
/* y [P][ca+cb] = concat(a [P][ca], b [P][cb]) (materialised tf.concat for three-way concatenations) */
int sap3d_concat_channels(int32_t dtype, const void* a, const void* b, void* y, int64_t P, int32_t ca, int32_t cb, void* stream);
/* gradient of sap3d_concat_channels: dy [P][ca+cb] -> da, db (nullable; acc_* != 0 accumulates) */
int sap3d_split_channels(int32_t dtype, const void* dy, void* da, int32_t acc_a, void* db, int32_t acc_b, int64_t P, int32_t ca,
                         int32_t cb, void* stream);

/* Backward passes of the GN variant (what tf.gradients derives for gn/p3d_gn.py:24-46 and utils/network.py:198-274;
 * training driver gn/train_p3d_gn_dataset.py:186-199).  One workspace of sap3d_gn_bwd_workspace(N,S,C) bytes serves both. */
size_t sap3d_gn_bwd_workspace(int32_t N, int64_t S, int32_t C);
/* backward of y = relu_out?( relu1?(GN1(a)) + relu2?(GN2(b) | b) ) with per-sample scale/shift [N][C] and group
 * statistics mean/rstd [N][G] from sap3d_gn_finalize.  s2 == NULL: b is a plain tensor (db = masked dy).
 * da/db nullable; d{gamma,beta} are accumulated (+=). */
int sap3d_gn_act_bwd(int32_t dtype, const void* dy, const void* a, const float* s1, const float* t1, const float* mean1,
                     const float* rstd1, const float* gamma1, int32_t relu1, const void* b, const float* s2, const float* t2,
                     const float* mean2, const float* rstd2, const float* gamma2, int32_t relu2, int32_t relu_out, int32_t N,
                     int64_t S, int32_t C, int32_t G, void* da, int32_t acc_a, void* db, int32_t acc_b, float* dgamma1,
                     float* dbeta1, float* dgamma2, float* dbeta2, void* workspace, void* stream);
/* backward of the block tail y = relu(GN(c3) + cbam_block(r)) (gn/p3d_gn.py:175-177): gradients w.r.t. c3, r, the GroupNorm
 * affine, the channel-attention MLP (w0 [C][hidden], b0, w1 [hidden][C], b1) and the 7x7x7 spatial filter (all +=).
 * cscale / sp / att / save are the tensors sap3d_cbam_fwd produced; reduce_max gradients are split evenly among ties. */
int sap3d_cbam_tail_bwd(int32_t dtype, const void* dy, const void* y, const void* c3, const float* s3, const float* mean3,
                        const float* rstd3, const float* gamma3, const void* r, int32_t N, int32_t D, int32_t H, int32_t W,
                        int32_t C, int32_t G, int32_t hidden, const float* w0, const float* w1, const float* w_sp,
                        const float* cscale, const float* sp, const float* att, const float* save, void* dc3, int32_t acc_c3,
                        void* dr, int32_t acc_r, float* dgamma3, float* dbeta3, float* dw0, float* db0, float* dw1, float* db1,
                        float* dw_sp, void* workspace, void* stream);

/* one-launch re-packing of all filters after an optimizer step (table built with sap3d_conv_pack_entries) */
typedef struct sap3d_pack_entry {
  const float* src;   /* fp32 TF-layout filter */
  void* dst;          /* bf16 [rows_pad][taps*cols] */
  int32_t taps, rows, rows_pad, cols;
  int64_t s_tap, s_r, s_c;
  int64_t start;      /* first element of this entry in the concatenated index space (filled by the caller; index space only,
                         not memory: every start and the total passed to sap3d_pack_multi are multiples of 4096) */
} sap3d_pack_entry;
/* fills up to two entries (forward / data-gradient operand; NULL destination skips one); returns the count */
int sap3d_conv_pack_entries(const sap3d_conv_desc* d, const float* w_tf, void* w_fwd, void* w_dgrad, sap3d_pack_entry* out2);
int sap3d_pack_multi(const sap3d_pack_entry* entries_dev, int32_t n, int64_t total, void* stream);

/* tf.train.AdamOptimizer (train.py:168) over flat fp32 arrays; *step (device) is the 1-based iteration. */
int sap3d_adam_step(float* w, const float* g, float* m, float* v, int64_t n, const int32_t* step, float lr, float b1, float b2,
                    float eps, float grad_scale, void* stream);
/* the same update with the gradient given as fp32 (g_dtype SAP3D_F32) or bf16 (SAP3D_BF16: the all-reduced buckets of the
 * data-parallel exchange, consumed without widening them back into the fp32 gradient buffer; new work, BASELINE configs[3]) */
int sap3d_adam_step_g(float* w, const void* g, int32_t g_dtype, float* m, float* v, int64_t n, const int32_t* step, float lr, float b1,
                      float b2, float eps, float grad_scale, void* stream);
int sap3d_step_increment(int32_t* step, void* stream);
/* f32 -> bf16 (src_dtype == SAP3D_F32) or bf16 -> f32 */
int sap3d_cast(int32_t src_dtype, const void* src, void* dst, int64_t n, void* stream);


/* HOST helper (no device needed): CRC-32C of n bytes, continuing from `crc` (0 to start).  The checksum of TensorFlow's
 * tensor-bundle checkpoint files, which tf.train.Saver writes/reads at train.py:184,267 and gen_pred.py:57-64; used by the
 * checkpoint reader/writer of the Python host (sap3d_tensorflow_b200/checkpoint.py). */
uint32_t sap3d_crc32c(uint32_t crc, const void* data, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* SAP3D_H_ */
