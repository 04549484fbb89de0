// TensorFlow-1.x custom-op shim for libsap3d_b200.so  ——  SOURCE ONLY, UNVERIFIED IN THIS IMAGE.
//
// TensorFlow (any version) is not installable here (no network, no wheel), so this file has never been
// compiled; it documents the exact binding a maintainer adds to use the B200 kernels from the reference's
// graph builders through tf.load_op_library (BASELINE.json north_star).  It is pure marshalling: every
// Compute() forwards TF-owned device buffers and the op's CUDA stream to one C-ABI entry point of
// include/sap3d.h.  Build (on a machine with TF 1.15 / tf.compat.v1 and CUDA 12.9):
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++14 -shared -fPIC tf_ops/sap3d_tf_ops.cc -o tf_ops/libsap3d_tf_ops.so \
//       $TF_CFLAGS $TF_LFLAGS -DGOOGLE_CUDA=1 -I include -I /usr/local/cuda/include \
//       -L sap3d_tensorflow_b200/lib -lsap3d_b200 -Wl,-rpath,'$ORIGIN/../sap3d_tensorflow_b200/lib'
//
// Conventions honoured: memory is owned by TF's allocator (allocate_output /
// allocate_temp; the kernels never cudaMalloc), work is enqueued on the op's stream and never
// synchronised, the C ABI is re-entrant, errors surface through OP_REQUIRES.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "sap3d.h"

namespace tf = tensorflow;

namespace {

inline void* StreamOf(tf::OpKernelContext* ctx) {
  return reinterpret_cast<void*>(ctx->eigen_device<Eigen::GpuDevice>().stream());
}
inline int32_t DtypeOf(const tf::Tensor& t) { return t.dtype() == tf::DT_BFLOAT16 ? SAP3D_BF16 : SAP3D_F32; }
inline const void* P(const tf::Tensor& t) { return t.tensor_data().data(); }
inline void* P(tf::Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }

struct ConvAttrs {
  std::vector<tf::int32> ksize, strides;
  bool transposed, has_bias;
  void Fill(const tf::Tensor& x0, int nseg, int cin1, int cout, sap3d_conv_desc* d) const {
    memset(d, 0, sizeof(*d));
    d->dtype = DtypeOf(x0);
    d->impl = SAP3D_IMPL_AUTO;
    d->N = x0.dim_size(0); d->D = x0.dim_size(1); d->H = x0.dim_size(2); d->W = x0.dim_size(3);
    d->nseg = nseg; d->cin[0] = x0.dim_size(4); d->cin[1] = cin1; d->cout = cout;
    d->kd = ksize[0]; d->kh = ksize[1]; d->kw = ksize[2];
    d->sd = strides[0]; d->sh = strides[1]; d->sw = strides[2];
    d->transposed = transposed; d->has_bias = has_bias;
  }
};

}  // namespace

// y, stats = Sap3dConv(x0, x1, filter, packed_filter, bias)
//   replaces tf.nn.conv3d + tf.nn.bias_add (p3d.py:18-27,86,112,125,343), tf.layers.conv3d /
//   conv3d_transpose (utils/network.py:101,107) and tf.concat feeding them (utils/network.py:97).
REGISTER_OP("Sap3dConv")
    .Input("x0: T").Input("x1: T").Input("filter: float").Input("packed_filter: bfloat16").Input("bias: float")
    .Output("y: T").Output("stats: float")
    .Attr("T: {bfloat16, float}").Attr("nseg: int = 1").Attr("cout: int")
    .Attr("ksize: list(int)").Attr("strides: list(int)").Attr("transposed: bool = false").Attr("has_bias: bool = false");

class Sap3dConvOp : public tf::OpKernel {
 public:
  explicit Sap3dConvOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &a_.ksize));
    OP_REQUIRES_OK(c, c->GetAttr("strides", &a_.strides));
    OP_REQUIRES_OK(c, c->GetAttr("transposed", &a_.transposed));
    OP_REQUIRES_OK(c, c->GetAttr("has_bias", &a_.has_bias));
    OP_REQUIRES_OK(c, c->GetAttr("nseg", &nseg_));
    OP_REQUIRES_OK(c, c->GetAttr("cout", &cout_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x0 = ctx->input(0);
    const tf::Tensor& x1 = ctx->input(1);
    sap3d_conv_desc d;
    a_.Fill(x0, nseg_, nseg_ > 1 ? x1.dim_size(4) : 0, cout_, &d);
    int32_t o[3];
    OP_REQUIRES(ctx, sap3d_conv_out_dims(&d, o) == 0, tf::errors::InvalidArgument(sap3d_last_error()));
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.N, o[0], o[1], o[2], cout_}), &y));
    const int rows = sap3d_conv_stats_rows(&d);
    tf::Tensor* stats = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({rows, 2, cout_}), &stats));
    const int rc = sap3d_conv_fwd(&d, P(x0), nseg_ > 1 ? P(x1) : nullptr, ctx->input(2).flat<float>().data(), P(ctx->input(3)),
                                  a_.has_bias ? ctx->input(4).flat<float>().data() : nullptr, P(y),
                                  stats->flat<float>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(sap3d_last_error()));
  }

 private:
  ConvAttrs a_;
  int nseg_, cout_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dConv").Device(tf::DEVICE_GPU), Sap3dConvOp);

// y = Sap3dNormApply(a, scale1, shift1, b, scale2, shift2): fused tf.layers.batch_normalization apply + tf.nn.relu +
// residual add (p3d.py:56-81,133-134).  scale/shift come from Sap3dBnFinalize (sap3d_bn_finalize).
REGISTER_OP("Sap3dNormApply")
    .Input("a: T").Input("scale1: float").Input("shift1: float").Input("b: T").Input("scale2: float").Input("shift2: float")
    .Output("y: T").Attr("T: {bfloat16, float}")
    .Attr("relu1: bool").Attr("relu2: bool").Attr("relu_out: bool").Attr("has_b: bool").Attr("norm_b: bool")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return tf::Status::OK(); });

class Sap3dNormApplyOp : public tf::OpKernel {
 public:
  explicit Sap3dNormApplyOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("relu1", &r1_)); OP_REQUIRES_OK(c, c->GetAttr("relu2", &r2_));
    OP_REQUIRES_OK(c, c->GetAttr("relu_out", &ro_)); OP_REQUIRES_OK(c, c->GetAttr("has_b", &hb_));
    OP_REQUIRES_OK(c, c->GetAttr("norm_b", &nb_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& a = ctx->input(0);
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, a.shape(), &y));
    const int64_t C = a.dim_size(4), P_ = a.NumElements() / C;
    const int rc = sap3d_affine_act(DtypeOf(a), P(a), ctx->input(1).flat<float>().data(), ctx->input(2).flat<float>().data(), r1_,
                                    hb_ ? P(ctx->input(3)) : nullptr, nb_ ? ctx->input(4).flat<float>().data() : nullptr,
                                    nb_ ? ctx->input(5).flat<float>().data() : nullptr, r2_, ro_, P(y), P_, (int32_t)C, 0, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(sap3d_last_error()));
  }

 private:
  bool r1_, r2_, ro_, hb_, nb_;
};
REGISTER_KERNEL_BUILDER(Name("Sap3dNormApply").Device(tf::DEVICE_GPU), Sap3dNormApplyOp);

// metrics = Sap3dSaliencyMetrics(pred, density, fixation) -> [n, 4] float64 (CC, SIM, NSS, KLdiv) — utils/metrics.py
REGISTER_OP("Sap3dSaliencyMetrics").Input("pred: float").Input("density: float").Input("fixation: float").Output("m: double");
class Sap3dSaliencyMetricsOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& p = ctx->input(0);
    const int64_t n = p.dim_size(0), e = p.NumElements() / n;
    tf::Tensor* m = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 4}), &m));
    const int rc = sap3d_saliency_metrics(p.flat<float>().data(), ctx->input(1).flat<float>().data(), ctx->input(2).flat<float>().data(),
                                          (int32_t)n, e, e, e, e, m->flat<double>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(sap3d_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("Sap3dSaliencyMetrics").Device(tf::DEVICE_GPU), Sap3dSaliencyMetricsOp);

// The remaining ops (Sap3dBnFinalize, Sap3dConvGrad{Input,Filter}, Sap3dNormApplyGrad, Sap3dMaxPool3d[Grad], Sap3dAttention[Grad],
// Sap3dHead[Grad], Sap3dCbam, Sap3dAdam) follow the same three-step pattern (shape -> allocate_output -> one sap3d_* call);
// gradients are wired in Python with @tf.RegisterGradient.
