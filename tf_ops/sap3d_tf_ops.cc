// TensorFlow-1.x custom-op shim for libsap3d_b200.so: one op per C-ABI entry-point family of include/sap3d.h, so that the
// reference's graph builders (p3d.py, gn/p3d_gn.py, utils/network.py) can swap TensorFlow's stock kernels for the B200
// kernels through tf.load_op_library (BASELINE.json north_star).  Gradients are wired in tf_ops/sap3d_grads.py.
//
// STATUS: TensorFlow (any version) is not installable in the build image (no network, no wheel), so this file is
// SYNTAX- AND TYPE-CHECKED in CI against stand-ins of the four TF headers it includes (tf_ops/tf_stub/, TF-1.15 signatures;
// tests/test_tf_shim_cpu.py runs `g++ -fsyntax-only`), which checks every sap3d_* call against include/sap3d.h and every TF
// API use against the restated signatures -- it has never been linked against a real TensorFlow.  Build, on a machine with
// TF 1.15 (or tf.compat.v1) and CUDA 12.9:
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++14 -shared -fPIC tf_ops/sap3d_tf_ops.cc -o tf_ops/libsap3d_tf_ops.so $TF_CFLAGS $TF_LFLAGS -DGOOGLE_CUDA=1
//       -I include -I /usr/local/cuda/include -L sap3d_tensorflow_b200/lib -lsap3d_b200 -L /usr/local/cuda/lib64 -lcudart
//       -Wl,-rpath,'$ORIGIN/../sap3d_tensorflow_b200/lib'          (one command line)
//
// Conventions honoured: memory is owned by TF's allocator (allocate_output / allocate_temp; the kernels never
// cudaMalloc), work is enqueued on the op's stream and never synchronised, the C ABI is re-entrant, errors surface through
// OP_REQUIRES.  The ABI accumulates parameter gradients (+=); the ops zero their gradient outputs first (cudaMemsetAsync on
// the op's stream), so every op is a pure function of its inputs as TF expects.  Input tensors are never written: buffers the
// ABI updates in place (moving statistics, optimizer slots, the stem's im2col workspace) are op OUTPUTS initialised by a
// device-to-device copy of the corresponding input.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include <cuda_runtime_api.h>
#include <string.h>

#include <vector>

#include "sap3d.h"

namespace tf = tensorflow;
using tf::shape_inference::DimensionHandle;
using tf::shape_inference::InferenceContext;
using tf::shape_inference::ShapeHandle;

namespace {

inline cudaStream_t CudaStreamOf(tf::OpKernelContext* ctx) { return ctx->eigen_device<Eigen::GpuDevice>().stream(); }
inline void* StreamOf(tf::OpKernelContext* ctx) { return reinterpret_cast<void*>(CudaStreamOf(ctx)); }
inline int32_t DtypeOf(const tf::Tensor& t) { return t.dtype() == tf::DT_BFLOAT16 ? SAP3D_BF16 : SAP3D_F32; }
inline const void* P(const tf::Tensor& t) { return t.tensor_data().data(); }
inline void* P(tf::Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }
inline const float* F(const tf::Tensor& t) { return t.flat<float>().data(); }
inline float* F(tf::Tensor* t) { return t->flat<float>().data(); }
inline const float* FOrNull(const tf::Tensor& t, bool use) { return use ? t.flat<float>().data() : nullptr; }
inline size_t Bytes(const tf::Tensor& t) { return t.tensor_data().size(); }

#define SAP3D_OK(ctx, call) OP_REQUIRES(ctx, (call) == 0, tf::errors::Internal(sap3d_last_error()))
#define SAP3D_CUDA_OK(ctx, call) OP_REQUIRES(ctx, (call) == cudaSuccess, tf::errors::Internal("CUDA runtime call failed in a sap3d op"))

// zero-filled output (the ABI accumulates parameter gradients)
inline tf::Status ZeroOutput(tf::OpKernelContext* ctx, int index, const tf::TensorShape& shape, tf::Tensor** out) {
  TF_RETURN_IF_ERROR(ctx->allocate_output(index, shape, out));
  if (Bytes(**out) > 0 && cudaMemsetAsync(P(*out), 0, Bytes(**out), CudaStreamOf(ctx)) != cudaSuccess)
    return tf::errors::Internal("cudaMemsetAsync failed");
  return tf::Status::OK();
}
// output initialised with a copy of an input (buffers the ABI updates in place)
inline tf::Status CopyOutput(tf::OpKernelContext* ctx, int index, const tf::Tensor& src, tf::Tensor** out) {
  TF_RETURN_IF_ERROR(ctx->allocate_output(index, src.shape(), out));
  if (Bytes(src) > 0 && cudaMemcpyAsync(P(*out), P(src), Bytes(src), cudaMemcpyDeviceToDevice, CudaStreamOf(ctx)) != cudaSuccess)
    return tf::errors::Internal("cudaMemcpyAsync failed");
  return tf::Status::OK();
}

// ---- convolution attributes shared by the conv family ---------------------------------------------------------------
#define SAP3D_CONV_ATTRS                                                                                                   \
  .Attr("input_dims: list(int)").Attr("cin: list(int)").Attr("cout: int").Attr("ksize: list(int)").Attr("strides: list(int)") \
      .Attr("transposed: bool = false").Attr("has_bias: bool = false").Attr("out_f32: bool = false")                          \
      .Attr("storage: {'bf16', 'f32'} = 'bf16'")

struct ConvAttrs {
  std::vector<tf::int32> input_dims, cin, ksize, strides;
  tf::int32 cout;
  bool transposed, has_bias, out_f32;
  tf::string storage;
  tf::Status Init(tf::OpKernelConstruction* c) {
    TF_RETURN_IF_ERROR(c->GetAttr("input_dims", &input_dims));
    TF_RETURN_IF_ERROR(c->GetAttr("cin", &cin));
    TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
    TF_RETURN_IF_ERROR(c->GetAttr("ksize", &ksize));
    TF_RETURN_IF_ERROR(c->GetAttr("strides", &strides));
    TF_RETURN_IF_ERROR(c->GetAttr("transposed", &transposed));
    TF_RETURN_IF_ERROR(c->GetAttr("has_bias", &has_bias));
    TF_RETURN_IF_ERROR(c->GetAttr("out_f32", &out_f32));
    TF_RETURN_IF_ERROR(c->GetAttr("storage", &storage));
    if (input_dims.size() != 4 || ksize.size() != 3 || strides.size() != 3 || cin.empty() || cin.size() > 2)
      return tf::errors::InvalidArgument("sap3d conv: input_dims = [N,D,H,W], ksize / strides of length 3, cin of length 1 or 2");
    return tf::Status::OK();
  }
  void Fill(sap3d_conv_desc* d) const {
    memset(d, 0, sizeof(*d));
    d->dtype = storage == "bf16" ? SAP3D_BF16 : SAP3D_F32;
    d->impl = SAP3D_IMPL_AUTO;
    d->N = input_dims[0]; d->D = input_dims[1]; d->H = input_dims[2]; d->W = input_dims[3];
    d->nseg = static_cast<int32_t>(cin.size());
    d->cin[0] = cin[0]; d->cin[1] = cin.size() > 1 ? cin[1] : 0;
    d->cout = cout;
    d->kd = ksize[0]; d->kh = ksize[1]; d->kw = ksize[2];
    d->sd = strides[0]; d->sh = strides[1]; d->sw = strides[2];
    d->transposed = transposed; d->has_bias = has_bias; d->out_f32 = out_f32;
  }
};

// static output shape of a conv from its attributes (TF 'SAME' / conv3d_transpose 'same' geometry comes from the ABI)
tf::Status ConvOutShape(InferenceContext* c, int out_index) {
  std::vector<tf::int32> in, ks, st, cin;
  tf::int32 cout;
  bool transposed;
  TF_RETURN_IF_ERROR(c->GetAttr("input_dims", &in));
  TF_RETURN_IF_ERROR(c->GetAttr("ksize", &ks));
  TF_RETURN_IF_ERROR(c->GetAttr("strides", &st));
  TF_RETURN_IF_ERROR(c->GetAttr("cin", &cin));
  TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
  TF_RETURN_IF_ERROR(c->GetAttr("transposed", &transposed));
  if (in.size() != 4 || ks.size() != 3 || st.size() != 3 || cin.empty()) return tf::errors::InvalidArgument("sap3d conv: bad geometry attributes");
  sap3d_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.N = in[0]; d.D = in[1]; d.H = in[2]; d.W = in[3];
  d.nseg = static_cast<int32_t>(cin.size()); d.cin[0] = cin[0]; d.cin[1] = cin.size() > 1 ? cin[1] : 0; d.cout = cout;
  d.kd = ks[0]; d.kh = ks[1]; d.kw = ks[2]; d.sd = st[0]; d.sh = st[1]; d.sw = st[2]; d.transposed = transposed;
  int32_t o[3];
  if (sap3d_conv_out_dims(&d, o) != 0) return tf::errors::InvalidArgument(sap3d_last_error());
  c->set_output(out_index, c->MakeShape({in[0], o[0], o[1], o[2], cout}));
  return tf::Status::OK();
}

tf::Status SameAsInput0(InferenceContext* c) {
  c->set_output(0, c->input(0));
  return tf::Status::OK();
}

}  // namespace

// =====================================================================================================================
// fwd, dgrad = Sap3dPackFilter(filter): fp32 TF-layout filter -> bf16 K-major tensor-core operands (sap3d_conv_pack_weights).
// Run once per optimizer step per filter (the engine of this repo does all filters in one launch, sap3d_pack_multi).
// =====================================================================================================================
REGISTER_OP("Sap3dPackFilter")
    .Input("filter: float").Output("fwd: bfloat16").Output("dgrad: bfloat16") SAP3D_CONV_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Vector(c->UnknownDim()));
      c->set_output(1, c->Vector(c->UnknownDim()));
      return tf::Status::OK();
    });

class Sap3dPackFilterOp : public tf::OpKernel {
 public:
  explicit Sap3dPackFilterOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, a_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    sap3d_conv_desc d;
    a_.Fill(&d);
    // a conv whose forward operand buffer is a workspace conv_fwd fills itself (the Cin = 3 stem's im2col form) has no
    // pre-packed forward operand
    const bool ws = sap3d_conv_fwd_operand_is_workspace(&d) == 1;
    tf::Tensor *fwd = nullptr, *dg = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({ws ? 0 : static_cast<tf::int64>(sap3d_conv_packed_elems(&d, 0))}), &fwd));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({static_cast<tf::int64>(sap3d_conv_packed_elems(&d, 1))}), &dg));
    SAP3D_OK(ctx, sap3d_conv_pack_weights(&d, F(ctx->input(0)), ws ? nullptr : P(fwd), P(dg), StreamOf(ctx)));
  }

 private:
  ConvAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dPackFilter").Device(tf::DEVICE_GPU), Sap3dPackFilterOp);

// =====================================================================================================================
// y, stats, fwd_operand = Sap3dConv(x0, x1, filter, packed_filter, bias)
//   replaces tf.nn.conv3d + tf.nn.bias_add (p3d.py:18-27,86,112,125,343), tf.layers.conv3d / conv3d_transpose
//   (utils/network.py:101,107) and tf.concat feeding them (utils/network.py:97: x1 = second channel segment, pass x0 twice
//   when cin has one entry).  stats [rows][2][cout] feeds Sap3dBnFinalize.  fwd_operand: what Sap3dConvGradFilter needs of
//   the forward pass (the stem's im2col matrix; empty for every other conv).
// =====================================================================================================================
REGISTER_OP("Sap3dConv")
    .Input("x0: T").Input("x1: T").Input("filter: float").Input("packed_filter: bfloat16").Input("bias: float")
    .Output("y: Tout").Output("stats: float").Output("fwd_operand: bfloat16")
    .Attr("T: {bfloat16, float}").Attr("Tout: {bfloat16, float}") SAP3D_CONV_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      TF_RETURN_IF_ERROR(ConvOutShape(c, 0));
      tf::int32 cout;
      TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
      c->set_output(1, c->MakeShape({c->UnknownDim(), 2, cout}));
      c->set_output(2, c->Vector(c->UnknownDim()));
      return tf::Status::OK();
    });

class Sap3dConvOp : public tf::OpKernel {
 public:
  explicit Sap3dConvOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, a_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    sap3d_conv_desc d;
    a_.Fill(&d);
    int32_t o[3];
    OP_REQUIRES(ctx, sap3d_conv_out_dims(&d, o) == 0, tf::errors::InvalidArgument(sap3d_last_error()));
    tf::Tensor *y = nullptr, *stats = nullptr, *operand = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.N, o[0], o[1], o[2], d.cout}), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({sap3d_conv_stats_rows(&d), 2, d.cout}), &stats));
    const bool ws = sap3d_conv_fwd_operand_is_workspace(&d) == 1;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({ws ? static_cast<tf::int64>(sap3d_conv_packed_elems(&d, 0)) : 0}), &operand));
    SAP3D_OK(ctx, sap3d_conv_fwd(&d, P(ctx->input(0)), d.nseg > 1 ? P(ctx->input(1)) : nullptr, F(ctx->input(2)),
                                 ws ? P(operand) : P(ctx->input(3)), FOrNull(ctx->input(4), a_.has_bias), P(y), F(stats), StreamOf(ctx)));
  }

 private:
  ConvAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConv").Device(tf::DEVICE_GPU), Sap3dConvOp);

// y = Sap3dConvAffine(...): inference-mode tf.layers.batch_normalization (moving statistics) + tf.nn.relu folded into the conv
// epilogue (utils/network.py:100-110 with training=False): y = relu?((conv + bias) * scale + shift)
REGISTER_OP("Sap3dConvAffine")
    .Input("x0: T").Input("x1: T").Input("filter: float").Input("packed_filter: bfloat16").Input("bias: float")
    .Input("scale: float").Input("shift: float").Output("y: T")
    .Attr("T: {bfloat16, float}").Attr("relu: bool = true") SAP3D_CONV_ATTRS
    .SetShapeFn([](InferenceContext* c) { return ConvOutShape(c, 0); });

class Sap3dConvAffineOp : public tf::OpKernel {
 public:
  explicit Sap3dConvAffineOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, a_.Init(c));
    OP_REQUIRES_OK(c, c->GetAttr("relu", &relu_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    sap3d_conv_desc d;
    a_.Fill(&d);
    OP_REQUIRES(ctx, sap3d_conv_fwd_on_tensor_cores(&d) == 1 && sap3d_conv_fwd_operand_is_workspace(&d) == 0,
                tf::errors::InvalidArgument("Sap3dConvAffine: this geometry does not take the fused epilogue; use Sap3dConv + Sap3dNormApply"));
    int32_t o[3];
    OP_REQUIRES(ctx, sap3d_conv_out_dims(&d, o) == 0, tf::errors::InvalidArgument(sap3d_last_error()));
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.N, o[0], o[1], o[2], d.cout}), &y));
    SAP3D_OK(ctx, sap3d_conv_fwd_affine(&d, P(ctx->input(0)), d.nseg > 1 ? P(ctx->input(1)) : nullptr, F(ctx->input(2)), P(ctx->input(3)),
                                        FOrNull(ctx->input(4), a_.has_bias), F(ctx->input(5)), F(ctx->input(6)), relu_, P(y), StreamOf(ctx)));
  }

 private:
  ConvAttrs a_;
  bool relu_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConvAffine").Device(tf::DEVICE_GPU), Sap3dConvAffineOp);

// dx = Sap3dConvGradInput(dy, filter, packed_dgrad): data gradient w.r.t. channel segment `seg` (tf.gradients of the conv)
REGISTER_OP("Sap3dConvGradInput")
    .Input("dy: T").Input("filter: float").Input("packed_dgrad: bfloat16").Output("dx: T")
    .Attr("T: {bfloat16, float}").Attr("seg: int = 0") SAP3D_CONV_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      std::vector<tf::int32> in, cin;
      tf::int32 seg;
      TF_RETURN_IF_ERROR(c->GetAttr("input_dims", &in));
      TF_RETURN_IF_ERROR(c->GetAttr("cin", &cin));
      TF_RETURN_IF_ERROR(c->GetAttr("seg", &seg));
      if (in.size() != 4 || seg < 0 || seg >= static_cast<tf::int32>(cin.size())) return tf::errors::InvalidArgument("Sap3dConvGradInput: bad seg / input_dims");
      c->set_output(0, c->MakeShape({in[0], in[1], in[2], in[3], cin[seg]}));
      return tf::Status::OK();
    });

class Sap3dConvGradInputOp : public tf::OpKernel {
 public:
  explicit Sap3dConvGradInputOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, a_.Init(c));
    OP_REQUIRES_OK(c, c->GetAttr("seg", &seg_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    sap3d_conv_desc d;
    a_.Fill(&d);
    OP_REQUIRES(ctx, seg_ >= 0 && seg_ < d.nseg, tf::errors::InvalidArgument("Sap3dConvGradInput: seg out of range"));
    tf::Tensor* dx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.N, d.D, d.H, d.W, d.cin[seg_]}), &dx));
    SAP3D_OK(ctx, sap3d_conv_dgrad(&d, seg_, P(ctx->input(0)), F(ctx->input(1)), P(ctx->input(2)), P(dx), 0, StreamOf(ctx)));
  }

 private:
  ConvAttrs a_;
  tf::int32 seg_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConvGradInput").Device(tf::DEVICE_GPU), Sap3dConvGradInputOp);

// dw, db = Sap3dConvGradFilter(x0, x1, dy, fwd_operand): filter (TF layout) and bias gradients
REGISTER_OP("Sap3dConvGradFilter")
    .Input("x0: T").Input("x1: T").Input("dy: T").Input("fwd_operand: bfloat16").Output("dw: float").Output("db: float")
    .Attr("T: {bfloat16, float}") SAP3D_CONV_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      std::vector<tf::int32> ks, cin;
      tf::int32 cout;
      bool transposed;
      TF_RETURN_IF_ERROR(c->GetAttr("ksize", &ks));
      TF_RETURN_IF_ERROR(c->GetAttr("cin", &cin));
      TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
      TF_RETURN_IF_ERROR(c->GetAttr("transposed", &transposed));
      if (ks.size() != 3 || cin.empty()) return tf::errors::InvalidArgument("Sap3dConvGradFilter: bad ksize / cin");
      const tf::int64 ci = cin[0] + (cin.size() > 1 ? cin[1] : 0);
      c->set_output(0, transposed ? c->MakeShape({ks[0], ks[1], ks[2], cout, ci}) : c->MakeShape({ks[0], ks[1], ks[2], ci, cout}));
      c->set_output(1, c->Vector(cout));
      return tf::Status::OK();
    });

class Sap3dConvGradFilterOp : public tf::OpKernel {
 public:
  explicit Sap3dConvGradFilterOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, a_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    sap3d_conv_desc d;
    a_.Fill(&d);
    const tf::int64 ci = d.cin[0] + d.cin[1];
    tf::Tensor *dw = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 0, a_.transposed ? tf::TensorShape({d.kd, d.kh, d.kw, d.cout, ci}) : tf::TensorShape({d.kd, d.kh, d.kw, ci, d.cout}), &dw));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 1, tf::TensorShape({d.cout}), &db));
    const tf::Tensor& operand = ctx->input(3);
    SAP3D_OK(ctx, sap3d_conv_wgrad(&d, P(ctx->input(0)), d.nseg > 1 ? P(ctx->input(1)) : nullptr, P(ctx->input(2)), F(dw),
                                   a_.has_bias ? F(db) : nullptr, operand.NumElements() > 0 ? P(operand) : nullptr, StreamOf(ctx)));
  }

 private:
  ConvAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConvGradFilter").Device(tf::DEVICE_GPU), Sap3dConvGradFilterOp);

// =====================================================================================================================
// One op per normalisation LAYER, so that gamma / beta are direct inputs and tf.gradients needs no knowledge of the kernels'
// split into statistics and apply passes:
//
// y, scale1, shift1, mean1, rstd1, new_mm1, new_mv1, scale2, shift2, mean2, rstd2, new_mm2, new_mv2 =
//     Sap3dBatchNormAct(a, stats_a, gamma1, beta1, moving_mean1, moving_variance1, b, stats_b, gamma2, beta2, moving_mean2, moving_variance2)
//   y = relu_out?( relu1?(BN1(a)) + relu2?(BN2(b) | b) ): tf.layers.batch_normalization (p3d.py:58-127,344; utils/network.py:91; eps 1e-3,
//   momentum 0.99, biased batch variance) + tf.nn.relu + the residual / ST_B / ST_C adds (p3d.py:56-81,133-134) in one pass.
//   stats_* are the [rows][2][C] outputs of Sap3dConv.  has_b: b is used; norm_b: b goes through its own BatchNorm.
//   The new moving averages are OUTPUTS: the Python wrapper assigns them inside UPDATE_OPS, where the stock layer registers its
//   updates (train.py:170-172).
// =====================================================================================================================
REGISTER_OP("Sap3dBatchNormAct")
    .Input("a: T").Input("stats_a: float").Input("gamma1: float").Input("beta1: float").Input("moving_mean1: float").Input("moving_variance1: float")
    .Input("b: T").Input("stats_b: float").Input("gamma2: float").Input("beta2: float").Input("moving_mean2: float").Input("moving_variance2: float")
    .Output("y: T").Output("scale1: float").Output("shift1: float").Output("mean1: float").Output("rstd1: float")
    .Output("new_moving_mean1: float").Output("new_moving_variance1: float")
    .Output("scale2: float").Output("shift2: float").Output("mean2: float").Output("rstd2: float")
    .Output("new_moving_mean2: float").Output("new_moving_variance2: float")
    .Attr("T: {bfloat16, float}").Attr("training: bool").Attr("momentum: float = 0.99").Attr("epsilon: float = 0.001")
    .Attr("relu1: bool = true").Attr("relu2: bool = false").Attr("relu_out: bool = false").Attr("has_b: bool = false").Attr("norm_b: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(0));
      for (int i = 1; i < 7; ++i) c->set_output(i, c->input(2));
      for (int i = 7; i < 13; ++i) c->set_output(i, c->input(8));
      return tf::Status::OK();
    });

class Sap3dBatchNormActOp : public tf::OpKernel {
 public:
  explicit Sap3dBatchNormActOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_)); OP_REQUIRES_OK(c, c->GetAttr("momentum", &momentum_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(0);
    const tf::int64 C = a.dim_size(a.dims() - 1), Pn = a.NumElements() / C;
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &y));
    const bool n2 = hb_ && nb_;
    tf::Tensor* o[2][6];
    for (int k = 0; k < 2; ++k) {
      const int in0 = k == 0 ? 2 : 8;   // gamma_k; moving statistics at in0 + 2, in0 + 3
      const tf::TensorShape shp = (k == 0 || n2) ? ctx->input(in0).shape() : tf::TensorShape({0});
      for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(ctx, ctx->allocate_output(1 + 6 * k + i, shp, &o[k][i]));
      OP_REQUIRES_OK(ctx, CopyOutput(ctx, 1 + 6 * k + 4, ctx->input(in0 + 2), &o[k][4]));
      OP_REQUIRES_OK(ctx, CopyOutput(ctx, 1 + 6 * k + 5, ctx->input(in0 + 3), &o[k][5]));
      if (k == 1 && !n2) continue;
      const tf::Tensor& stats = ctx->input(in0 - 1);
      SAP3D_OK(ctx, sap3d_bn_finalize(training_ ? F(stats) : nullptr, training_ ? static_cast<tf::int32>(stats.dim_size(0)) : 0, static_cast<int32_t>(C),
                                      static_cast<double>(Pn), F(ctx->input(in0)), F(ctx->input(in0 + 1)), F(o[k][4]), F(o[k][5]), training_, momentum_,
                                      eps_, F(o[k][0]), F(o[k][1]), F(o[k][2]), F(o[k][3]), StreamOf(ctx)));
    }
    SAP3D_OK(ctx, sap3d_affine_act(DtypeOf(a), P(a), F(o[0][0]), F(o[0][1]), r1_, hb_ ? P(ctx->input(6)) : nullptr, n2 ? F(o[1][0]) : nullptr,
                                   n2 ? F(o[1][1]) : nullptr, r2_, ro_, P(y), Pn, static_cast<int32_t>(C), 0, StreamOf(ctx)));
  }

 private:
  bool training_, r1_, r2_, ro_, hb_, nb_;
  float momentum_, eps_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dBatchNormAct").Device(tf::DEVICE_GPU), Sap3dBatchNormActOp);

// da, db, dgamma1, dbeta1, dgamma2, dbeta2 = Sap3dBatchNormActGrad(dy, a, scale1, shift1, mean1, rstd1, b, scale2, shift2, mean2, rstd2):
// the full BatchNorm backward (training: through the batch statistics) of the op above, masks recomputed from a / b
REGISTER_OP("Sap3dBatchNormActGrad")
    .Input("dy: T").Input("a: T").Input("scale1: float").Input("shift1: float").Input("mean1: float").Input("rstd1: float")
    .Input("b: T").Input("scale2: float").Input("shift2: float").Input("mean2: float").Input("rstd2: float")
    .Output("da: T").Output("db: T").Output("dgamma1: float").Output("dbeta1: float").Output("dgamma2: float").Output("dbeta2: float")
    .Attr("T: {bfloat16, float}").Attr("training: bool")
    .Attr("relu1: bool = true").Attr("relu2: bool = false").Attr("relu_out: bool = false").Attr("has_b: bool = false").Attr("norm_b: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(1));
      c->set_output(1, c->input(1));
      for (int i = 2; i < 6; ++i) c->set_output(i, c->input(2));
      return tf::Status::OK();
    });

class Sap3dBatchNormActGradOp : public tf::OpKernel {
 public:
  explicit Sap3dBatchNormActGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_));
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(1);
    const tf::int64 C = a.dim_size(a.dims() - 1), Pn = a.NumElements() / C;
    tf::Tensor *da = nullptr, *db = nullptr, *g[4];
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &da));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 1, hb_ ? a.shape() : tf::TensorShape({0}), &db));
    for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 2 + i, tf::TensorShape({C}), &g[i]));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({static_cast<tf::int64>(sap3d_affine_act_bwd_workspace(static_cast<int32_t>(C)) / 4 + 16)}), &ws));
    const bool n2 = hb_ && nb_;
    SAP3D_OK(ctx, sap3d_affine_act_bwd(DtypeOf(a), P(ctx->input(0)), P(a), F(ctx->input(2)), F(ctx->input(3)), FOrNull(ctx->input(4), training_),
                                       FOrNull(ctx->input(5), training_), r1_, hb_ ? P(ctx->input(6)) : nullptr, FOrNull(ctx->input(7), n2),
                                       FOrNull(ctx->input(8), n2), FOrNull(ctx->input(9), n2 && training_), FOrNull(ctx->input(10), n2 && training_), r2_,
                                       ro_, Pn, static_cast<int32_t>(C), P(da), 0, hb_ ? P(db) : nullptr, 0, F(g[0]), F(g[1]), n2 ? F(g[2]) : nullptr,
                                       n2 ? F(g[3]) : nullptr, P(&ws), StreamOf(ctx)));
  }

 private:
  bool training_, r1_, r2_, ro_, hb_, nb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dBatchNormActGrad").Device(tf::DEVICE_GPU), Sap3dBatchNormActGradOp);

// =====================================================================================================================
// y, scale1, shift1, mean1, rstd1, scale2, shift2, mean2, rstd2 = Sap3dGroupNormAct(a, gamma1, beta1, b, gamma2, beta2)
//   y = relu_out?( relu1?(GN1(a)) + relu2?(GN2(b) | b) ): GroupNorm of utils/network.py:65-87 == gn/p3d_gn.py:24-46 (G = min(32, C), eps 1e-5,
//   biased variance over (C/G, D, H, W) per sample, per-channel gamma / beta) + ReLU + adds (gn/p3d_gn.py:49-51,130-177)
// =====================================================================================================================
REGISTER_OP("Sap3dGroupNormAct")
    .Input("a: T").Input("gamma1: float").Input("beta1: float").Input("b: T").Input("gamma2: float").Input("beta2: float")
    .Output("y: T").Output("scale1: float").Output("shift1: float").Output("mean1: float").Output("rstd1: float")
    .Output("scale2: float").Output("shift2: float").Output("mean2: float").Output("rstd2: float")
    .Attr("T: {bfloat16, float}").Attr("groups: int = 32").Attr("epsilon: float = 0.00001")
    .Attr("relu1: bool = true").Attr("relu2: bool = false").Attr("relu_out: bool = false").Attr("has_b: bool = false").Attr("norm_b: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(0));
      for (int i = 1; i < 9; ++i) c->set_output(i, c->UnknownShapeOfRank(2));
      return tf::Status::OK();
    });

class Sap3dGroupNormActOp : public tf::OpKernel {
 public:
  explicit Sap3dGroupNormActOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("groups", &groups_)); OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(0);
    const tf::int64 N = a.dim_size(0), C = a.dim_size(4), S = a.NumElements() / (N * C);
    const tf::int32 G = groups_ < C ? groups_ : static_cast<tf::int32>(C);   // G = min(32, C), utils/network.py:73
    const bool n2 = hb_ && nb_;
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &y));
    tf::Tensor* o[2][4];
    for (int k = 0; k < 2; ++k) {
      const bool on = k == 0 || n2;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(1 + 4 * k, on ? tf::TensorShape({N, C}) : tf::TensorShape({0, 0}), &o[k][0]));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(2 + 4 * k, on ? tf::TensorShape({N, C}) : tf::TensorShape({0, 0}), &o[k][1]));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(3 + 4 * k, on ? tf::TensorShape({N, G}) : tf::TensorShape({0, 0}), &o[k][2]));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(4 + 4 * k, on ? tf::TensorShape({N, G}) : tf::TensorShape({0, 0}), &o[k][3]));
      if (!on) continue;
      const tf::Tensor& x = k == 0 ? a : ctx->input(3);
      SAP3D_OK(ctx, sap3d_gn_stats(DtypeOf(x), P(x), static_cast<int32_t>(N), S, static_cast<int32_t>(C), G, F(ctx->input(k == 0 ? 1 : 4)),
                                   F(ctx->input(k == 0 ? 2 : 5)), eps_, F(o[k][0]), F(o[k][1]), F(o[k][2]), F(o[k][3]), StreamOf(ctx)));
    }
    SAP3D_OK(ctx, sap3d_affine_act(DtypeOf(a), P(a), F(o[0][0]), F(o[0][1]), r1_, hb_ ? P(ctx->input(3)) : nullptr, n2 ? F(o[1][0]) : nullptr,
                                   n2 ? F(o[1][1]) : nullptr, r2_, ro_, P(y), N * S, static_cast<int32_t>(C), S, StreamOf(ctx)));
  }

 private:
  tf::int32 groups_;
  float eps_;
  bool r1_, r2_, ro_, hb_, nb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dGroupNormAct").Device(tf::DEVICE_GPU), Sap3dGroupNormActOp);

// =====================================================================================================================
// y = Sap3dClipBatchNormAct(a, gamma1, beta1, b, gamma2, beta2): the backbone's batch-statistics BatchNorm as gen_pred.py sees it
// (one 16-frame window per sess.run, gen_pred.py:88-135 -- so every window is normalised on its own) for a BATCH of windows:
// statistics over (D, H, W) per clip and channel, + ReLU + second operand, in one launch where the clip's slab is small
// (sap3d_sample_norm_apply), else per-clip partial sums + finalize + apply.  Inference only: no gradient, moving averages untouched.
// =====================================================================================================================
REGISTER_OP("Sap3dClipBatchNormAct")
    .Input("a: T").Input("gamma1: float").Input("beta1: float").Input("b: T").Input("gamma2: float").Input("beta2: float")
    .Output("y: T")
    .Attr("T: {bfloat16, float}").Attr("epsilon: float = 0.001")
    .Attr("relu1: bool = true").Attr("relu2: bool = false").Attr("relu_out: bool = false").Attr("has_b: bool = false").Attr("norm_b: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(0));
      return tf::Status::OK();
    });

class Sap3dClipBatchNormActOp : public tf::OpKernel {
 public:
  explicit Sap3dClipBatchNormActOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(0);
    const tf::int64 N = a.dim_size(0), C = a.dim_size(4), S = a.NumElements() / (N * C);
    const bool n2 = hb_ && nb_;
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &y));
    if (sap3d_sample_norm_apply_supported(DtypeOf(a), S, static_cast<int32_t>(C)) == 1) {
      SAP3D_OK(ctx, sap3d_sample_norm_apply(DtypeOf(a), P(a), F(ctx->input(1)), F(ctx->input(2)), r1_, hb_ ? P(ctx->input(3)) : nullptr,
                                            n2 ? F(ctx->input(4)) : nullptr, n2 ? F(ctx->input(5)) : nullptr, r2_, ro_, P(y),
                                            static_cast<int32_t>(N), S, static_cast<int32_t>(C), eps_, StreamOf(ctx)));
      return;
    }
    // large slabs: GroupNorm's kernels with one channel per group
    tf::Tensor coef[2][4];
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !n2) break;
      const tf::Tensor& x = k == 0 ? a : ctx->input(3);
      for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({N, C}), &coef[k][i]));
      SAP3D_OK(ctx, sap3d_gn_stats(DtypeOf(x), P(x), static_cast<int32_t>(N), S, static_cast<int32_t>(C), static_cast<int32_t>(C),
                                   F(ctx->input(k == 0 ? 1 : 4)), F(ctx->input(k == 0 ? 2 : 5)), eps_, F(&coef[k][0]), F(&coef[k][1]),
                                   F(&coef[k][2]), F(&coef[k][3]), StreamOf(ctx)));
    }
    SAP3D_OK(ctx, sap3d_affine_act(DtypeOf(a), P(a), F(&coef[0][0]), F(&coef[0][1]), r1_, hb_ ? P(ctx->input(3)) : nullptr,
                                   n2 ? F(&coef[1][0]) : nullptr, n2 ? F(&coef[1][1]) : nullptr, r2_, ro_, P(y), N * S, static_cast<int32_t>(C), S,
                                   StreamOf(ctx)));
  }

 private:
  float eps_;
  bool r1_, r2_, ro_, hb_, nb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dClipBatchNormAct").Device(tf::DEVICE_GPU), Sap3dClipBatchNormActOp);

// da, db, dgamma1, dbeta1, dgamma2, dbeta2 = Sap3dGroupNormActGrad(...)
REGISTER_OP("Sap3dGroupNormActGrad")
    .Input("dy: T").Input("a: T").Input("scale1: float").Input("shift1: float").Input("mean1: float").Input("rstd1: float").Input("gamma1: float")
    .Input("b: T").Input("scale2: float").Input("shift2: float").Input("mean2: float").Input("rstd2: float").Input("gamma2: float")
    .Output("da: T").Output("db: T").Output("dgamma1: float").Output("dbeta1: float").Output("dgamma2: float").Output("dbeta2: float")
    .Attr("T: {bfloat16, float}").Attr("relu1: bool = true").Attr("relu2: bool = false").Attr("relu_out: bool = false")
    .Attr("has_b: bool = false").Attr("norm_b: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(1));
      c->set_output(1, c->input(1));
      for (int i = 2; i < 6; ++i) c->set_output(i, c->input(6));
      return tf::Status::OK();
    });

class Sap3dGroupNormActGradOp : public tf::OpKernel {
 public:
  explicit Sap3dGroupNormActGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(1);
    const tf::int64 N = a.dim_size(0), C = a.dim_size(4), S = a.NumElements() / (N * C);
    const tf::int32 G = static_cast<tf::int32>(ctx->input(4).dim_size(1));
    tf::Tensor *da = nullptr, *db = nullptr, *g[4];
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &da));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 1, hb_ ? a.shape() : tf::TensorShape({0}), &db));
    for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 2 + i, tf::TensorShape({C}), &g[i]));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({static_cast<tf::int64>(sap3d_gn_bwd_workspace(static_cast<int32_t>(N), S, static_cast<int32_t>(C)) / 4 + 16)}), &ws));
    SAP3D_CUDA_OK(ctx, cudaMemsetAsync(P(&ws), 0, Bytes(ws), CudaStreamOf(ctx)));
    const bool n2 = hb_ && nb_;
    SAP3D_OK(ctx, sap3d_gn_act_bwd(DtypeOf(a), P(ctx->input(0)), P(a), F(ctx->input(2)), F(ctx->input(3)), F(ctx->input(4)), F(ctx->input(5)),
                                   F(ctx->input(6)), r1_, hb_ ? P(ctx->input(7)) : nullptr, FOrNull(ctx->input(8), n2), FOrNull(ctx->input(9), n2),
                                   FOrNull(ctx->input(10), n2), FOrNull(ctx->input(11), n2), FOrNull(ctx->input(12), n2), r2_, ro_,
                                   static_cast<int32_t>(N), S, static_cast<int32_t>(C), G, P(da), 0, hb_ ? P(db) : nullptr, 0, F(g[0]), F(g[1]),
                                   n2 ? F(g[2]) : nullptr, n2 ? F(g[3]) : nullptr, P(&ws), StreamOf(ctx)));
  }

 private:
  bool r1_, r2_, ro_, hb_, nb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dGroupNormActGrad").Device(tf::DEVICE_GPU), Sap3dGroupNormActGradOp);

// =====================================================================================================================
// CBAM block tail of the GN backbone: y = relu(GN(c3) + cbam_block(r))   (gn/p3d_gn.py:175-177; utils/network.py:198-274)
// cscale / sp / att / save and the GroupNorm statistics are kept for the gradient op.
// =====================================================================================================================
REGISTER_OP("Sap3dCbamTail")
    .Input("c3: T").Input("gamma3: float").Input("beta3: float").Input("r: T")
    .Input("w0: float").Input("b0: float").Input("w1: float").Input("b1: float").Input("w_sp: float")
    .Output("y: T").Output("scale3: float").Output("mean3: float").Output("rstd3: float")
    .Output("cscale: float").Output("sp: float").Output("att: float").Output("save: float")
    .Attr("T: {bfloat16, float}").Attr("groups: int = 32").Attr("epsilon: float = 0.00001")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(3));
      for (int i = 1; i < 8; ++i) c->set_output(i, c->UnknownShape());
      return tf::Status::OK();
    });

class Sap3dCbamTailOp : public tf::OpKernel {
 public:
  explicit Sap3dCbamTailOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("groups", &groups_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& r = ctx->input(3);
    const tf::int32 N = static_cast<tf::int32>(r.dim_size(0)), D = static_cast<tf::int32>(r.dim_size(1)), H = static_cast<tf::int32>(r.dim_size(2)),
                    W = static_cast<tf::int32>(r.dim_size(3)), C = static_cast<tf::int32>(r.dim_size(4));
    const tf::int64 S = static_cast<tf::int64>(D) * H * W;
    const tf::int32 hidden = static_cast<tf::int32>(ctx->input(4).dim_size(1)), G = groups_ < C ? groups_ : C;
    const tf::int32 rows = sap3d_sample_stats_rows(S, C, N);
    tf::Tensor *y = nullptr, *scale3 = nullptr, *mean3 = nullptr, *rstd3 = nullptr, *cscale = nullptr, *sp = nullptr, *att = nullptr, *save = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, r.shape(), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({N, C}), &scale3));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({N, G}), &mean3));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({N, G}), &rstd3));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, tf::TensorShape({N, C}), &cscale));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(5, tf::TensorShape({N, S, 2}), &sp));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(6, tf::TensorShape({N, S}), &att));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(7, tf::TensorShape({N, 2 * C + 2 * hidden}), &save));
    tf::Tensor part, shift3;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({N, rows, 3, C}), &part));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({N, C}), &shift3));
    const tf::Tensor& c3 = ctx->input(0);
    SAP3D_OK(ctx, sap3d_gn_stats(DtypeOf(c3), P(c3), N, S, C, G, F(ctx->input(1)), F(ctx->input(2)), eps_, F(scale3), F(&shift3), F(mean3), F(rstd3),
                                 StreamOf(ctx)));
    SAP3D_OK(ctx, sap3d_cbam_fwd(DtypeOf(r), P(r), N, D, H, W, C, hidden, F(ctx->input(4)), F(ctx->input(5)), F(ctx->input(6)), F(ctx->input(7)),
                                 F(ctx->input(8)), F(&part), rows, F(cscale), F(sp), F(att), F(save), StreamOf(ctx)));
    SAP3D_OK(ctx, sap3d_cbam_merge(DtypeOf(r), P(c3), F(scale3), F(&shift3), P(r), F(cscale), F(att), P(y), N, S, C, StreamOf(ctx)));
  }

 private:
  tf::int32 groups_;
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dCbamTail").Device(tf::DEVICE_GPU), Sap3dCbamTailOp);

// dc3, dgamma3, dbeta3, dr, dw0, db0, dw1, db1, dw_sp = Sap3dCbamTailGrad(...): one gradient per input of Sap3dCbamTail, in its order
REGISTER_OP("Sap3dCbamTailGrad")
    .Input("dy: T").Input("y: T").Input("c3: T").Input("scale3: float").Input("mean3: float").Input("rstd3: float").Input("gamma3: float")
    .Input("r: T").Input("w0: float").Input("w1: float").Input("w_sp: float")
    .Input("cscale: float").Input("sp: float").Input("att: float").Input("save: float")
    .Output("dc3: T").Output("dgamma3: float").Output("dbeta3: float").Output("dr: T")
    .Output("dw0: float").Output("db0: float").Output("dw1: float").Output("db1: float").Output("dw_sp: float")
    .Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(2));
      c->set_output(1, c->input(6));
      c->set_output(2, c->input(6));
      c->set_output(3, c->input(7));
      c->set_output(4, c->input(8));
      c->set_output(5, c->UnknownShapeOfRank(1));
      c->set_output(6, c->input(9));
      c->set_output(7, c->input(6));
      c->set_output(8, c->input(10));
      return tf::Status::OK();
    });

class Sap3dCbamTailGradOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& r = ctx->input(7);
    const tf::int32 N = static_cast<tf::int32>(r.dim_size(0)), D = static_cast<tf::int32>(r.dim_size(1)), H = static_cast<tf::int32>(r.dim_size(2)),
                    W = static_cast<tf::int32>(r.dim_size(3)), C = static_cast<tf::int32>(r.dim_size(4));
    const tf::int64 S = static_cast<tf::int64>(D) * H * W;
    const tf::int32 hidden = static_cast<tf::int32>(ctx->input(8).dim_size(1)), G = static_cast<tf::int32>(ctx->input(4).dim_size(1));
    tf::Tensor *dc3 = nullptr, *dr = nullptr, *dg3 = nullptr, *db3 = nullptr, *g[5];
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, r.shape(), &dc3));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 1, tf::TensorShape({C}), &dg3));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 2, tf::TensorShape({C}), &db3));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, r.shape(), &dr));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 4, tf::TensorShape({C, hidden}), &g[0]));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 5, tf::TensorShape({hidden}), &g[1]));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 6, tf::TensorShape({hidden, C}), &g[2]));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 7, tf::TensorShape({C}), &g[3]));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 8, tf::TensorShape({7, 7, 7, 2, 1}), &g[4]));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({static_cast<tf::int64>(sap3d_gn_bwd_workspace(N, S, C) / 4 + 16)}), &ws));
    SAP3D_CUDA_OK(ctx, cudaMemsetAsync(P(&ws), 0, Bytes(ws), CudaStreamOf(ctx)));
    SAP3D_OK(ctx, sap3d_cbam_tail_bwd(DtypeOf(r), P(ctx->input(0)), P(ctx->input(1)), P(ctx->input(2)), F(ctx->input(3)), F(ctx->input(4)),
                                      F(ctx->input(5)), F(ctx->input(6)), P(r), N, D, H, W, C, G, hidden, F(ctx->input(8)), F(ctx->input(9)),
                                      F(ctx->input(10)), F(ctx->input(11)), F(ctx->input(12)), F(ctx->input(13)), F(ctx->input(14)), P(dc3), 0,
                                      P(dr), 0, F(dg3), F(db3), F(g[0]), F(g[1]), F(g[2]), F(g[3]), F(g[4]), P(&ws), StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dCbamTailGrad").Device(tf::DEVICE_GPU), Sap3dCbamTailGradOp);

// =====================================================================================================================
// Pooling: tf.nn.max_pool3d (p3d.py:347-348,354,360,366) and tf.layers.max_pooling3d (utils/network.py:6-7)
// =====================================================================================================================
REGISTER_OP("Sap3dMaxPool3d")
    .Input("x: T").Output("y: T").Output("argmax: uint8")
    .Attr("T: {bfloat16, float}").Attr("ksize: list(int)").Attr("strides: list(int)").Attr("same: bool = true")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->UnknownShapeOfRank(5));
      c->set_output(1, c->UnknownShapeOfRank(5));
      return tf::Status::OK();
    });

struct PoolAttrs {
  std::vector<tf::int32> ksize, strides;
  bool same;
  tf::Status Init(tf::OpKernelConstruction* c) {
    TF_RETURN_IF_ERROR(c->GetAttr("ksize", &ksize));
    TF_RETURN_IF_ERROR(c->GetAttr("strides", &strides));
    TF_RETURN_IF_ERROR(c->GetAttr("same", &same));
    if (ksize.size() != 3 || strides.size() != 3) return tf::errors::InvalidArgument("Sap3dMaxPool3d: ksize / strides of length 3");
    return tf::Status::OK();
  }
};

class Sap3dMaxPool3dOp : public tf::OpKernel {
 public:
  explicit Sap3dMaxPool3dOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, a_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const tf::int32 N = static_cast<tf::int32>(x.dim_size(0)), D = static_cast<tf::int32>(x.dim_size(1)), H = static_cast<tf::int32>(x.dim_size(2)),
                    W = static_cast<tf::int32>(x.dim_size(3)), C = static_cast<tf::int32>(x.dim_size(4));
    int32_t o[3];
    SAP3D_OK(ctx, sap3d_maxpool3d_out_dims(D, H, W, a_.ksize.data(), a_.strides.data(), a_.same, o));
    tf::Tensor *y = nullptr, *am = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({N, o[0], o[1], o[2], C}), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({N, o[0], o[1], o[2], C}), &am));
    SAP3D_OK(ctx, sap3d_maxpool3d_fwd(DtypeOf(x), P(x), N, D, H, W, C, a_.ksize.data(), a_.strides.data(), a_.same, P(y),
                                      am->flat<tf::uint8>().data(), StreamOf(ctx)));
  }

 private:
  PoolAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dMaxPool3d").Device(tf::DEVICE_GPU), Sap3dMaxPool3dOp);

REGISTER_OP("Sap3dMaxPool3dGrad")
    .Input("x: T").Input("dy: T").Input("argmax: uint8").Output("dx: T")
    .Attr("T: {bfloat16, float}").Attr("ksize: list(int)").Attr("strides: list(int)").Attr("same: bool = true")
    .SetShapeFn(SameAsInput0);

class Sap3dMaxPool3dGradOp : public tf::OpKernel {
 public:
  explicit Sap3dMaxPool3dGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, a_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    tf::Tensor* dx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    SAP3D_OK(ctx, sap3d_maxpool3d_bwd(DtypeOf(x), P(x), P(ctx->input(1)), static_cast<tf::int32>(x.dim_size(0)), static_cast<tf::int32>(x.dim_size(1)),
                                      static_cast<tf::int32>(x.dim_size(2)), static_cast<tf::int32>(x.dim_size(3)), static_cast<tf::int32>(x.dim_size(4)),
                                      a_.ksize.data(), a_.strides.data(), a_.same, ctx->input(2).flat<tf::uint8>().data(), P(dx), 0, StreamOf(ctx)));
  }

 private:
  PoolAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dMaxPool3dGrad").Device(tf::DEVICE_GPU), Sap3dMaxPool3dGradOp);

// =====================================================================================================================
// Self-attention core of utils/network.py:184-186 (tf.matmul(g, f^T) -> tf.nn.softmax -> tf.matmul(beta, h)), fused:
// o, lse = Sap3dFlashAttention(q, k, v) with q [B,Nq,64], k [B,Nk,64] (d_k zero-padded to 64), v [B,Nk,dv], bf16
// =====================================================================================================================
REGISTER_OP("Sap3dFlashAttention")
    .Input("q: bfloat16").Input("k: bfloat16").Input("v: bfloat16").Output("o: bfloat16").Output("lse: float")
    .SetShapeFn([](InferenceContext* c) {
      ShapeHandle q, v;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &q));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 3, &v));
      c->set_output(0, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1), c->Dim(v, 2)}));
      c->set_output(1, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1)}));
      return tf::Status::OK();
    });

class Sap3dFlashAttentionOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2);
    const tf::int32 B = static_cast<tf::int32>(q.dim_size(0)), Nq = static_cast<tf::int32>(q.dim_size(1)), Nk = static_cast<tf::int32>(k.dim_size(1)),
                    dk = static_cast<tf::int32>(q.dim_size(2)), dv = static_cast<tf::int32>(v.dim_size(2));
    tf::Tensor *o = nullptr, *lse = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, Nq, dv}), &o));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, Nq}), &lse));
    SAP3D_OK(ctx, sap3d_flash_attn_fwd(P(q), P(k), P(v), P(o), F(lse), B, Nq, Nk, dk, dv, StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dFlashAttention").Device(tf::DEVICE_GPU), Sap3dFlashAttentionOp);

REGISTER_OP("Sap3dFlashAttentionGrad")
    .Input("q: bfloat16").Input("k: bfloat16").Input("v: bfloat16").Input("o: bfloat16").Input("d_o: bfloat16").Input("lse: float")
    .Output("dq: bfloat16").Output("dk: bfloat16").Output("dv: bfloat16")
    .SetShapeFn([](InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->input(i));
      return tf::Status::OK();
    });

class Sap3dFlashAttentionGradOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2);
    const tf::int32 B = static_cast<tf::int32>(q.dim_size(0)), Nq = static_cast<tf::int32>(q.dim_size(1)), Nk = static_cast<tf::int32>(k.dim_size(1)),
                    dk = static_cast<tf::int32>(q.dim_size(2)), dv = static_cast<tf::int32>(v.dim_size(2));
    tf::Tensor *dq = nullptr, *dkk = nullptr, *dvv = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, q.shape(), &dq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, k.shape(), &dkk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, v.shape(), &dvv));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({static_cast<tf::int64>(sap3d_flash_attn_bwd_workspace(B, Nq, Nk, dv) / 4 + 16)}), &ws));
    SAP3D_OK(ctx, sap3d_flash_attn_bwd(P(q), P(k), P(v), P(ctx->input(3)), P(ctx->input(4)), F(ctx->input(5)), P(dq), P(dkk), P(dvv), B, Nq, Nk, dk, dv,
                                       P(&ws), StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dFlashAttentionGrad").Device(tf::DEVICE_GPU), Sap3dFlashAttentionGradOp);

// generic path (any d_k / d_v, f32 or bf16): o, beta = Sap3dAttention(g, f, h)
REGISTER_OP("Sap3dAttention")
    .Input("g: T").Input("f: T").Input("h: T").Output("o: T").Output("beta: T").Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      ShapeHandle g, f, h;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &g));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &f));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 3, &h));
      c->set_output(0, c->MakeShape({c->Dim(g, 0), c->Dim(g, 1), c->Dim(h, 2)}));
      c->set_output(1, c->MakeShape({c->Dim(g, 0), c->Dim(g, 1), c->Dim(f, 1)}));
      return tf::Status::OK();
    });

class Sap3dAttentionOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &g = ctx->input(0), &f = ctx->input(1), &h = ctx->input(2);
    const tf::int32 B = static_cast<tf::int32>(g.dim_size(0)), Nq = static_cast<tf::int32>(g.dim_size(1)), Nk = static_cast<tf::int32>(f.dim_size(1)),
                    dk = static_cast<tf::int32>(g.dim_size(2)), dv = static_cast<tf::int32>(h.dim_size(2));
    tf::Tensor *o = nullptr, *beta = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, Nq, dv}), &o));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, Nq, Nk}), &beta));
    SAP3D_OK(ctx, sap3d_attention_fwd(DtypeOf(g), P(g), P(f), P(h), P(beta), P(o), B, Nq, Nk, dk, dv, dk, dk, dv, Nk, dv, StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dAttention").Device(tf::DEVICE_GPU), Sap3dAttentionOp);

REGISTER_OP("Sap3dAttentionGrad")
    .Input("g: T").Input("f: T").Input("h: T").Input("beta: T").Input("d_o: T").Output("dg: T").Output("df: T").Output("dh: T")
    .Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->input(i));
      return tf::Status::OK();
    });

class Sap3dAttentionGradOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &g = ctx->input(0), &f = ctx->input(1), &h = ctx->input(2), &beta = ctx->input(3);
    const tf::int32 B = static_cast<tf::int32>(g.dim_size(0)), Nq = static_cast<tf::int32>(g.dim_size(1)), Nk = static_cast<tf::int32>(f.dim_size(1)),
                    dk = static_cast<tf::int32>(g.dim_size(2)), dv = static_cast<tf::int32>(h.dim_size(2));
    tf::Tensor *dg = nullptr, *df = nullptr, *dh = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, g.shape(), &dg));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, f.shape(), &df));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, h.shape(), &dh));
    tf::Tensor ds;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(beta.dtype(), beta.shape(), &ds));
    SAP3D_OK(ctx, sap3d_attention_bwd(DtypeOf(g), P(g), P(f), P(h), P(beta), P(ctx->input(4)), P(&ds), P(dg), P(df), P(dh), B, Nq, Nk, dk, dv, dk, dk,
                                      dv, Nk, dv, StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dAttentionGrad").Device(tf::DEVICE_GPU), Sap3dAttentionGradOp);

// y = o * gamma + x  (utils/network.py:191-192)
REGISTER_OP("Sap3dGate").Input("o: T").Input("x: T").Input("gamma: float").Output("y: T").Attr("T: {bfloat16, float}").SetShapeFn(SameAsInput0);
class Sap3dGateOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& o = ctx->input(0);
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, o.shape(), &y));
    SAP3D_OK(ctx, sap3d_gate_fwd(DtypeOf(o), P(o), P(ctx->input(1)), F(ctx->input(2)), P(y), o.NumElements(), StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dGate").Device(tf::DEVICE_GPU), Sap3dGateOp);

REGISTER_OP("Sap3dGateGrad")
    .Input("dy: T").Input("o: T").Input("gamma: float").Output("d_o: T").Output("dx: T").Output("dgamma: float").Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(0));
      c->set_output(2, c->input(2));
      return tf::Status::OK();
    });
class Sap3dGateGradOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& dy = ctx->input(0);
    tf::Tensor *d_o = nullptr, *dx = nullptr, *dgamma = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, dy.shape(), &d_o));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, dy.shape(), &dx));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 2, ctx->input(2).shape(), &dgamma));
    SAP3D_OK(ctx, sap3d_gate_bwd(DtypeOf(dy), P(dy), P(ctx->input(1)), F(ctx->input(2)), P(d_o), P(dx), 0, F(dgamma), dy.NumElements(), StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dGateGrad").Device(tf::DEVICE_GPU), Sap3dGateGradOp);

// =====================================================================================================================
// Decoder head (p3d.py:393,397: tf.layers.conv3d_transpose -> 1 channel, k3 s2, + tf.sigmoid), loss (utils/network.py:49-62), dropout
// =====================================================================================================================
REGISTER_OP("Sap3dHead")
    .Input("x: T").Input("filter: float").Input("bias: float").Output("logits: float").Output("pred: float")
    .Attr("T: {bfloat16, float}").Attr("ksize: list(int)").Attr("stride: int = 2")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->UnknownShapeOfRank(5));
      c->set_output(1, c->UnknownShapeOfRank(5));
      return tf::Status::OK();
    });
class Sap3dHeadOp : public tf::OpKernel {
 public:
  explicit Sap3dHeadOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &ksize_));
    OP_REQUIRES_OK(c, c->GetAttr("stride", &stride_));
    OP_REQUIRES(c, ksize_.size() == 3, tf::errors::InvalidArgument("Sap3dHead: ksize of length 3"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const tf::int32 N = static_cast<tf::int32>(x.dim_size(0)), D = static_cast<tf::int32>(x.dim_size(1)), H = static_cast<tf::int32>(x.dim_size(2)),
                    W = static_cast<tf::int32>(x.dim_size(3)), C = static_cast<tf::int32>(x.dim_size(4));
    tf::Tensor *logits = nullptr, *pred = nullptr;
    const tf::TensorShape shp({N, D * stride_, H * stride_, W * stride_, 1});
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, shp, &logits));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, shp, &pred));
    SAP3D_OK(ctx, sap3d_head_fwd(DtypeOf(x), P(x), N, D, H, W, C, ksize_.data(), stride_, F(ctx->input(1)), F(ctx->input(2)), F(logits), F(pred),
                                 StreamOf(ctx)));
  }

 private:
  std::vector<tf::int32> ksize_;
  tf::int32 stride_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dHead").Device(tf::DEVICE_GPU), Sap3dHeadOp);

REGISTER_OP("Sap3dHeadGrad")
    .Input("dlogits: float").Input("x: T").Input("filter: float").Output("dx: T").Output("dw: float")
    .Attr("T: {bfloat16, float}").Attr("ksize: list(int)").Attr("stride: int = 2")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(1));
      c->set_output(1, c->input(2));
      return tf::Status::OK();
    });
class Sap3dHeadGradOp : public tf::OpKernel {
 public:
  explicit Sap3dHeadGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &ksize_));
    OP_REQUIRES_OK(c, c->GetAttr("stride", &stride_));
    OP_REQUIRES(c, ksize_.size() == 3, tf::errors::InvalidArgument("Sap3dHeadGrad: ksize of length 3"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(1);
    tf::Tensor *dx = nullptr, *dw = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 1, ctx->input(2).shape(), &dw));
    SAP3D_OK(ctx, sap3d_head_bwd(DtypeOf(x), F(ctx->input(0)), P(x), static_cast<tf::int32>(x.dim_size(0)), static_cast<tf::int32>(x.dim_size(1)),
                                 static_cast<tf::int32>(x.dim_size(2)), static_cast<tf::int32>(x.dim_size(3)), static_cast<tf::int32>(x.dim_size(4)),
                                 ksize_.data(), stride_, F(ctx->input(2)), P(dx), 0, F(dw), StreamOf(ctx)));
  }

 private:
  std::vector<tf::int32> ksize_;
  tf::int32 stride_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dHeadGrad").Device(tf::DEVICE_GPU), Sap3dHeadGradOp);

// loss, dlogits, dbias = Sap3dSmoothL1Loss(logits, target): utils/network.py:49-62 summed over all elements (train.py:159),
// with its gradient w.r.t. the logits (through the sigmoid when apply_sigmoid) in the same pass
REGISTER_OP("Sap3dSmoothL1Loss")
    .Input("logits: float").Input("target: float").Output("loss: double").Output("dlogits: float").Output("dbias: float")
    .Attr("apply_sigmoid: bool = true").Attr("sigma: float = 1.0").Attr("inside_weight: float = 1.0").Attr("outside_weight: float = 1.0")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Scalar());
      c->set_output(1, c->input(0));
      c->set_output(2, c->Vector(1));
      return tf::Status::OK();
    });
class Sap3dSmoothL1LossOp : public tf::OpKernel {
 public:
  explicit Sap3dSmoothL1LossOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("apply_sigmoid", &sig_));
    OP_REQUIRES_OK(c, c->GetAttr("sigma", &sigma_));
    OP_REQUIRES_OK(c, c->GetAttr("inside_weight", &wi_));
    OP_REQUIRES_OK(c, c->GetAttr("outside_weight", &wo_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& logits = ctx->input(0);
    tf::Tensor *loss = nullptr, *dl = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 0, tf::TensorShape({}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, logits.shape(), &dl));
    OP_REQUIRES_OK(ctx, ZeroOutput(ctx, 2, tf::TensorShape({1}), &db));
    SAP3D_OK(ctx, sap3d_loss_smooth_l1_ex(F(logits), F(ctx->input(1)), logits.NumElements(), sig_, nullptr, F(dl), loss->flat<double>().data(), F(db),
                                          sigma_, wi_, wo_, StreamOf(ctx)));
  }

 private:
  bool sig_;
  float sigma_, wi_, wo_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dSmoothL1Loss").Device(tf::DEVICE_GPU), Sap3dSmoothL1LossOp);

// y = Sap3dDropout(x, step): tf.layers.dropout (p3d.py:392), counter-hash mask of (seed, step, element); the gradient is the same op on dy
REGISTER_OP("Sap3dDropout").Input("x: T").Input("step: int32").Output("y: T").Attr("T: {bfloat16, float}").Attr("rate: float").Attr("seed: int = 1234")
    .SetShapeFn(SameAsInput0);
class Sap3dDropoutOp : public tf::OpKernel {
 public:
  explicit Sap3dDropoutOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("rate", &rate_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &y));
    SAP3D_OK(ctx, sap3d_dropout(DtypeOf(x), P(x), P(y), x.NumElements(), rate_, static_cast<uint64_t>(seed_), ctx->input(1).flat<tf::int32>().data(), 0,
                                StreamOf(ctx)));
  }

 private:
  float rate_;
  tf::int64 seed_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dDropout").Device(tf::DEVICE_GPU), Sap3dDropoutOp);

// three-way tf.concat (gn/p3d_gn.py:251,527; p3d.py:267): two-way concats are fused into the convs as channel segments
REGISTER_OP("Sap3dConcatChannels").Input("a: T").Input("b: T").Output("y: T").Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->UnknownShapeOfRank(5));
      return tf::Status::OK();
    });
class Sap3dConcatChannelsOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &a = ctx->input(0), &b = ctx->input(1);
    const tf::int64 ca = a.dim_size(4), cb = b.dim_size(4), Pn = a.NumElements() / ca;
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({a.dim_size(0), a.dim_size(1), a.dim_size(2), a.dim_size(3), ca + cb}), &y));
    SAP3D_OK(ctx, sap3d_concat_channels(DtypeOf(a), P(a), P(b), P(y), Pn, static_cast<int32_t>(ca), static_cast<int32_t>(cb), StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConcatChannels").Device(tf::DEVICE_GPU), Sap3dConcatChannelsOp);

REGISTER_OP("Sap3dSplitChannels").Input("dy: T").Output("da: T").Output("db: T").Attr("T: {bfloat16, float}").Attr("ca: int").Attr("cb: int")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->UnknownShapeOfRank(5));
      c->set_output(1, c->UnknownShapeOfRank(5));
      return tf::Status::OK();
    });
class Sap3dSplitChannelsOp : public tf::OpKernel {
 public:
  explicit Sap3dSplitChannelsOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("ca", &ca_));
    OP_REQUIRES_OK(c, c->GetAttr("cb", &cb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& dy = ctx->input(0);
    const tf::int64 Pn = dy.NumElements() / (ca_ + cb_);
    tf::Tensor *da = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({dy.dim_size(0), dy.dim_size(1), dy.dim_size(2), dy.dim_size(3), ca_}), &da));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({dy.dim_size(0), dy.dim_size(1), dy.dim_size(2), dy.dim_size(3), cb_}), &db));
    SAP3D_OK(ctx, sap3d_split_channels(DtypeOf(dy), P(dy), P(da), 0, P(db), 0, Pn, ca_, cb_, StreamOf(ctx)));
  }

 private:
  tf::int32 ca_, cb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dSplitChannels").Device(tf::DEVICE_GPU), Sap3dSplitChannelsOp);

// =====================================================================================================================
// new_var, new_m, new_v = Sap3dAdam(var, grad, m, v, step): tf.train.AdamOptimizer(lr).minimize (train.py:166-172);
// step (int32 scalar on the device) is the 1-based iteration
// =====================================================================================================================
REGISTER_OP("Sap3dAdam")
    .Input("var: float").Input("grad: float").Input("m: float").Input("v: float").Input("step: int32")
    .Output("new_var: float").Output("new_m: float").Output("new_v: float")
    .Attr("learning_rate: float = 0.0001").Attr("beta1: float = 0.9").Attr("beta2: float = 0.999").Attr("epsilon: float = 0.00000001")
    .SetShapeFn([](InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->input(0));
      return tf::Status::OK();
    });
class Sap3dAdamOp : public tf::OpKernel {
 public:
  explicit Sap3dAdamOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("learning_rate", &lr_));
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    tf::Tensor *w = nullptr, *m = nullptr, *v = nullptr;
    OP_REQUIRES_OK(ctx, CopyOutput(ctx, 0, ctx->input(0), &w));
    OP_REQUIRES_OK(ctx, CopyOutput(ctx, 1, ctx->input(2), &m));
    OP_REQUIRES_OK(ctx, CopyOutput(ctx, 2, ctx->input(3), &v));
    SAP3D_OK(ctx, sap3d_adam_step(F(w), F(ctx->input(1)), F(m), F(v), w->NumElements(), ctx->input(4).flat<tf::int32>().data(), lr_, b1_, b2_, eps_, 1.f,
                                  StreamOf(ctx)));
  }

 private:
  float lr_, b1_, b2_, eps_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dAdam").Device(tf::DEVICE_GPU), Sap3dAdamOp);

// =====================================================================================================================
// Evaluation: utils/metrics.py on the device
// =====================================================================================================================
// m = Sap3dSaliencyMetrics(pred, density, fixation) -> [n, 4] float64 (CC :227, SIM :258, NSS :200, KLdiv :338)
REGISTER_OP("Sap3dSaliencyMetrics").Input("pred: float").Input("density: float").Input("fixation: float").Output("m: double")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), 4));
      return tf::Status::OK();
    });
class Sap3dSaliencyMetricsOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& p = ctx->input(0);
    const tf::int64 n = p.dim_size(0), e = p.NumElements() / n;
    tf::Tensor* m = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 4}), &m));
    SAP3D_OK(ctx, sap3d_saliency_metrics(F(p), F(ctx->input(1)), F(ctx->input(2)), static_cast<int32_t>(n), e, e, e, e, m->flat<double>().data(),
                                         StreamOf(ctx)));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dSaliencyMetrics").Device(tf::DEVICE_GPU), Sap3dSaliencyMetricsOp);

// up = Sap3dResizeBilinear(maps [n,h,w]) -> [n,H,W]: cv2.resize(prediction, (960, 1080)) of test.py:168 (INTER_LINEAR semantics)
REGISTER_OP("Sap3dResizeBilinear").Input("maps: float").Output("up: float").Attr("height: int").Attr("width: int")
    .SetShapeFn([](InferenceContext* c) {
      tf::int32 h, w;
      TF_RETURN_IF_ERROR(c->GetAttr("height", &h));
      TF_RETURN_IF_ERROR(c->GetAttr("width", &w));
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), h, w}));
      return tf::Status::OK();
    });
class Sap3dResizeBilinearOp : public tf::OpKernel {
 public:
  explicit Sap3dResizeBilinearOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("height", &h_));
    OP_REQUIRES_OK(c, c->GetAttr("width", &w_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& s = ctx->input(0);
    tf::Tensor* up = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({s.dim_size(0), h_, w_}), &up));
    SAP3D_OK(ctx, sap3d_resize_bilinear(F(s), static_cast<int32_t>(s.dim_size(0)), static_cast<int32_t>(s.dim_size(1)), static_cast<int32_t>(s.dim_size(2)),
                                        F(up), h_, w_, StreamOf(ctx)));
  }

 private:
  tf::int32 h_, w_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dResizeBilinear").Device(tf::DEVICE_GPU), Sap3dResizeBilinearOp);

// auc = Sap3dSaliencyAuc(saliency [n,...], fixation [n,...]) -> [n, 2] float64 (AUC_Judd, AUC_Borji; utils/metrics.py:25-154, test.py:174-175)
REGISTER_OP("Sap3dSaliencyAuc").Input("saliency: float").Input("fixation: float").Output("auc: double")
    .Attr("jitter: bool = true").Attr("n_rep: int = 100").Attr("step_size: float = 0.1").Attr("seed: int = 0")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), 2));
      return tf::Status::OK();
    });
class Sap3dSaliencyAucOp : public tf::OpKernel {
 public:
  explicit Sap3dSaliencyAucOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("jitter", &jitter_));
    OP_REQUIRES_OK(c, c->GetAttr("n_rep", &n_rep_));
    OP_REQUIRES_OK(c, c->GetAttr("step_size", &step_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& s = ctx->input(0);
    const tf::int32 n = static_cast<tf::int32>(s.dim_size(0));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 2}), &out));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({static_cast<tf::int64>(sap3d_saliency_auc_workspace(n, n_rep_) / 4 + 16)}), &ws));
    SAP3D_OK(ctx, sap3d_saliency_auc(F(s), F(ctx->input(1)), n, s.NumElements() / n, jitter_, n_rep_, step_, static_cast<uint64_t>(seed_),
                                     out->flat<double>().data(), P(&ws), StreamOf(ctx)));
  }

 private:
  bool jitter_;
  tf::int32 n_rep_;
  float step_;
  tf::int64 seed_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dSaliencyAuc").Device(tf::DEVICE_GPU), Sap3dSaliencyAucOp);

// frames = Sap3dPreprocessFrames(bgr_u8 [n,h,w,3]) -> [n,H,W,3]: mapf of dataflow.py:194-209 / gen_pred.py:113-118
REGISTER_OP("Sap3dPreprocessFrames").Input("bgr: uint8").Output("frames: T").Attr("T: {bfloat16, float} = DT_FLOAT")
    .Attr("height: int = 112").Attr("width: int = 112").Attr("mean_rgb: list(float) = [90.0, 102.0, 98.0]")
    .SetShapeFn([](InferenceContext* c) {
      tf::int32 h, w;
      TF_RETURN_IF_ERROR(c->GetAttr("height", &h));
      TF_RETURN_IF_ERROR(c->GetAttr("width", &w));
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), h, w, 3}));
      return tf::Status::OK();
    });
class Sap3dPreprocessFramesOp : public tf::OpKernel {
 public:
  explicit Sap3dPreprocessFramesOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("height", &h_));
    OP_REQUIRES_OK(c, c->GetAttr("width", &w_));
    OP_REQUIRES_OK(c, c->GetAttr("mean_rgb", &mean_));
    OP_REQUIRES(c, mean_.size() == 3, tf::errors::InvalidArgument("Sap3dPreprocessFrames: mean_rgb of length 3"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& f = ctx->input(0);
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({f.dim_size(0), h_, w_, 3}), &out));
    SAP3D_OK(ctx, sap3d_preprocess_frames(f.flat<tf::uint8>().data(), static_cast<int32_t>(f.dim_size(0)), static_cast<int32_t>(f.dim_size(1)),
                                          static_cast<int32_t>(f.dim_size(2)), mean_.data(), DtypeOf(*out), P(out), h_, w_, StreamOf(ctx)));
  }

 private:
  tf::int32 h_, w_;
  std::vector<float> mean_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dPreprocessFrames").Device(tf::DEVICE_GPU), Sap3dPreprocessFramesOp);
