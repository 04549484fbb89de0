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
/* element count of the packed bf16 weights used by the tcgen05 path; which: 0 = forward, 1 = data-gradient */
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
/* dx_seg = data gradient w.r.t. segment `seg`; accumulate != 0 adds into dx */
int sap3d_conv_dgrad(const sap3d_conv_desc* d, int32_t seg, const void* dy, const float* w_tf,
                     const void* w_dgrad_packed, void* dx, int32_t accumulate, void* stream);
/* dw (fp32, TF layout) += filter gradient; db (nullable, [cout]) += bias gradient.
 * The caller zeroes dw/db at the start of a step (gradients accumulate across calls). */
int sap3d_conv_wgrad(const sap3d_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw,
                     float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAP3D_H_ */
