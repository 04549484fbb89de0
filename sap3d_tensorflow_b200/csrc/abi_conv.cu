// C-ABI entry points of the convolution family (see include/sap3d.h).  Lowers a TF-semantics conv /
// transposed-conv descriptor to (a) the tcgen05 implicit-GEMM "form F" problem of conv_tc.cu or
// (b) the CUDA-core gather kernels of conv_simt.cu.
#include <cuda_bf16.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <vector>

#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"

using namespace sap3d;

extern "C" int sap3d_gemm_tn(const void* P, int64_t ldp, const void* Q, int64_t ldq, float* D, int64_t ldd, int32_t M, int32_t N,
                             int32_t Kpos, void* stream);

namespace {

struct DimGeom {
  int I, O, k, s, pb;
};

struct ConvGeom {
  DimGeom d[3];  // D,H,W
  int taps, cin_total, cout;
  int seg_off[2];
};

void make_geom(const sap3d_conv_desc* c, ConvGeom& g) {
  const int I[3] = {c->D, c->H, c->W};
  const int k[3] = {c->kd, c->kh, c->kw};
  const int s[3] = {c->sd, c->sh, c->sw};
  for (int i = 0; i < 3; ++i) {
    DimGeom& d = g.d[i];
    d.I = I[i];
    d.k = k[i];
    d.s = s[i];
    if (!c->transposed) {
      d.O = (I[i] + s[i] - 1) / s[i];
      int pt = (d.O - 1) * s[i] + k[i] - I[i];
      if (pt < 0) pt = 0;
      d.pb = pt / 2;
    } else {
      d.O = I[i] * s[i];
      int pt = k[i] - s[i];
      if (pt < 0) pt = 0;
      d.pb = pt / 2;
    }
  }
  g.taps = c->kd * c->kh * c->kw;
  g.cin_total = c->cin[0] + (c->nseg > 1 ? c->cin[1] : 0);
  g.cout = c->cout;
  g.seg_off[0] = 0;
  g.seg_off[1] = c->cin[0];
}

int check_desc(const sap3d_conv_desc* c) {
  if (!c) return set_error("conv desc is NULL");
  if (c->nseg < 1 || c->nseg > 2) return set_error("conv desc: nseg must be 1 or 2");
  if (c->N < 1 || c->D < 1 || c->H < 1 || c->W < 1 || c->cout < 1 || c->cin[0] < 1) return set_error("conv desc: bad extent");
  if (c->kd < 1 || c->kh < 1 || c->kw < 1 || c->sd < 1 || c->sh < 1 || c->sw < 1) return set_error("conv desc: bad kernel/stride");
  if (c->dtype != SAP3D_BF16 && c->dtype != SAP3D_F32) return set_error("conv desc: bad dtype");
  return 0;
}

// may the tensor-core path serve this descriptor (all of fwd and dgrad)?
bool tc_eligible(const sap3d_conv_desc* c) {
  if (c->impl == SAP3D_IMPL_SIMT) return false;
  if (c->dtype != SAP3D_BF16) return false;
  for (int s = 0; s < c->nseg; ++s)
    if (c->cin[s] % 64 != 0) return false;
  if (c->cout % 8 != 0) return false;
  const int k[3] = {c->kd, c->kh, c->kw};
  const int st[3] = {c->sd, c->sh, c->sw};
  if (!c->transposed) {
    for (int i = 0; i < 3; ++i)
      if (st[i] != 1 && k[i] != 1) return false;  // strided taps (stem) -> SIMT
  } else {
    long long ncls = (long long)c->sd * c->sh * c->sw;
    if (ncls > 64) return false;
  }
  return true;
}
// dgrad on the tensor-core path additionally needs cout % 64 == 0 (cout is the K dimension there)
bool tc_dgrad_eligible(const sap3d_conv_desc* c) {
  if (!tc_eligible(c)) return false;
  if (c->cout % 64 != 0) return false;
  if (c->transposed) {
    // parity views of dy: at most 27 tensor maps (k3 s4: deconv_pool4 of the concat decoders)
    ConvGeom g;
    make_geom(c, g);
    int nv = 1;
    for (int i = 0; i < 3; ++i) {
      std::map<int, int> rs;
      for (int k = 0; k < g.d[i].k; ++k) rs[((k - g.d[i].pb) % g.d[i].s + g.d[i].s) % g.d[i].s] = 1;
      nv *= (int)rs.size();
    }
    if (nv > 27) return false;
  }
  return true;
}

void out_strides(const ConvGeom& g, int C, long long so[4]) {
  so[0] = C;
  so[1] = (long long)g.d[2].O * C;
  so[2] = (long long)g.d[1].O * g.d[2].O * C;
  so[3] = (long long)g.d[0].O * g.d[1].O * g.d[2].O * C;
}
void in_strides(const ConvGeom& g, int C, long long si[4]) {
  si[0] = C;
  si[1] = (long long)g.d[2].I * C;
  si[2] = (long long)g.d[1].I * g.d[2].I * C;
  si[3] = (long long)g.d[0].I * g.d[1].I * g.d[2].I * C;
}

int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// ---- forward -> TcProblem ---------------------------------------------------------------------
void build_fwd_problem(const sap3d_conv_desc* c, const ConvGeom& g, const void* x0, const void* x1, TcProblem& pb) {
  const void* xs[2] = {x0, x1};
  pb.views.clear();
  pb.classes.clear();
  if (!c->transposed) {
    // views sample the input with the conv stride (only k == 1 dims may be strided)
    for (int s = 0; s < c->nseg; ++s) {
      TcView v;
      v.base = xs[s];
      v.C = c->cin[s];
      long long si[4];
      in_strides(g, c->cin[s], si);
      for (int i = 0; i < 3; ++i) {
        const DimGeom& d = g.d[2 - i];  // i=0 -> W
        v.dim[i] = (d.s == 1) ? d.I : d.O;
        v.stride[i] = si[i] * d.s;
      }
      v.dim[3] = c->N;
      v.stride[3] = si[3];
      pb.views.push_back(v);
    }
    TcClassH cls;
    cls.out_ofs = 0;
    for (int kd = 0; kd < c->kd; ++kd)
      for (int kh = 0; kh < c->kh; ++kh)
        for (int kw = 0; kw < c->kw; ++kw) {
          const int tap = (kd * c->kh + kh) * c->kw + kw;
          for (int s = 0; s < c->nseg; ++s) {
            TcTapH t;
            t.view = s;
            t.off[0] = (g.d[2].s == 1) ? kw - g.d[2].pb : 0;
            t.off[1] = (g.d[1].s == 1) ? kh - g.d[1].pb : 0;
            t.off[2] = (g.d[0].s == 1) ? kd - g.d[0].pb : 0;
            t.off[3] = 0;
            t.kofs = tap * g.cin_total + g.seg_off[s];
            t.c_begin = 0;
            t.nch = c->cin[s];
            cls.taps.push_back(t);
          }
        }
    pb.classes.push_back(cls);
    pb.ext[0] = g.d[2].O; pb.ext[1] = g.d[1].O; pb.ext[2] = g.d[0].O; pb.ext[3] = c->N;
    out_strides(g, c->cout, pb.so);
  } else {
    for (int s = 0; s < c->nseg; ++s) {
      TcView v;
      v.base = xs[s];
      v.C = c->cin[s];
      long long si[4];
      in_strides(g, c->cin[s], si);
      for (int i = 0; i < 3; ++i) {
        v.dim[i] = g.d[2 - i].I;
        v.stride[i] = si[i];
      }
      v.dim[3] = c->N;
      v.stride[3] = si[3];
      pb.views.push_back(v);
    }
    long long so[4];
    out_strides(g, c->cout, so);
    for (int rd = 0; rd < c->sd; ++rd)
      for (int rh = 0; rh < c->sh; ++rh)
        for (int rw = 0; rw < c->sw; ++rw) {
          TcClassH cls;
          cls.out_ofs = rd * so[2] + rh * so[1] + rw * so[0];
          for (int kd = 0; kd < c->kd; ++kd) {
            if ((rd + g.d[0].pb - kd) % c->sd != 0) continue;
            for (int kh = 0; kh < c->kh; ++kh) {
              if ((rh + g.d[1].pb - kh) % c->sh != 0) continue;
              for (int kw = 0; kw < c->kw; ++kw) {
                if ((rw + g.d[2].pb - kw) % c->sw != 0) continue;
                const int tap = (kd * c->kh + kh) * c->kw + kw;
                for (int s = 0; s < c->nseg; ++s) {
                  TcTapH t;
                  t.view = s;
                  t.off[0] = (rw + g.d[2].pb - kw) / c->sw;
                  t.off[1] = (rh + g.d[1].pb - kh) / c->sh;
                  t.off[2] = (rd + g.d[0].pb - kd) / c->sd;
                  t.off[3] = 0;
                  t.kofs = tap * g.cin_total + g.seg_off[s];
                  t.c_begin = 0;
                  t.nch = c->cin[s];
                  cls.taps.push_back(t);
                }
              }
            }
          }
          pb.classes.push_back(cls);
        }
    pb.ext[0] = g.d[2].I; pb.ext[1] = g.d[1].I; pb.ext[2] = g.d[0].I; pb.ext[3] = c->N;
    pb.so[0] = so[0] * c->sw; pb.so[1] = so[1] * c->sh; pb.so[2] = so[2] * c->sd; pb.so[3] = so[3];
  }
  pb.Ktot = g.taps * g.cin_total;
  pb.rowsB = (c->cout + 63) / 64 * 64;
  pb.cout = c->cout;
}

int simt_stats_rows(const sap3d_conv_desc* c, const ConvGeom& g) {
  long long P = (long long)c->N * g.d[0].O * g.d[1].O * g.d[2].O;
  long long rows = (P + 127) / 128;
  if (rows > 592) rows = 592;
  if (rows < 1) rows = 1;
  return (int)rows;
}

// ---- small-Cin convolutions (the 1x7x7 stem, Cin = 3) on the tensor cores ------------------------------------------
// im2col into a bf16 [positions][Kp] buffer (K = taps*cin zero-padded to a multiple of 64) -> plain tcgen05 GEMM with the
// usual bias / statistics epilogue; the filter gradient is the position-contracted GEMM col^T dy.  The buffer lives in the
// caller's "packed forward operand" allocation (sap3d_conv_packed_elems sizes it), so the ABI needs no extra workspace.
bool im2col_eligible(const sap3d_conv_desc* c) {
  if (c->impl == SAP3D_IMPL_SIMT || c->dtype != SAP3D_BF16 || c->transposed || c->nseg != 1 || c->out_f32) return false;
  if (c->cin[0] % 64 == 0 || c->cin[0] > 32 || c->cout % 64 != 0) return false;
  return (long long)c->kd * c->kh * c->kw * c->cin[0] <= 1024;
}
struct Im2colLayout {
  int K, Kp;
  long long P;
  size_t off_d, off_col, total;   // in bf16 elements: [wB: co_pad*Kp][D: Kp*cout fp32][col: P*Kp]
};
Im2colLayout im2col_layout(const sap3d_conv_desc* c, const ConvGeom& g) {
  Im2colLayout L;
  L.K = g.taps * g.cin_total;
  L.Kp = (L.K + 63) / 64 * 64;
  L.P = (long long)c->N * g.d[0].O * g.d[1].O * g.d[2].O;
  const size_t co_pad = (size_t)(c->cout + 63) / 64 * 64;
  auto up = [](size_t v) { return (v + 127) / 128 * 128; };
  L.off_d = up(co_pad * L.Kp);
  L.off_col = L.off_d + up((size_t)2 * L.Kp * c->cout);
  L.total = L.off_col + (size_t)L.P * L.Kp;
  return L;
}

struct Im2colArgs {
  const __nv_bfloat16* x; __nv_bfloat16* col;
  int N, D, H, W, C, oD, oH, oW, kd, kh, kw, sd, sh, sw, pd, ph, pw, K, Kp;
  long long P;
};
__global__ void __launch_bounds__(256) im2col_kernel(const Im2colArgs p) {
  // per-k lookup (tap coordinates and element offset inside the input window) built once per block: the inner loop is
  // three bounds checks, one add and one 2-byte load per element instead of six integer divisions
  __shared__ int s_off[1024];
  __shared__ int s_abe[1024];
  for (int k = threadIdx.x; k < p.Kp; k += blockDim.x) {
    if (k < p.K) {
      const int c = k % p.C;
      int t = k / p.C;
      const int e = t % p.kw; t /= p.kw;
      const int b = t % p.kh;
      const int a = t / p.kh;
      s_off[k] = ((a * p.H + b) * p.W + e) * p.C + c;
      s_abe[k] = a | (b << 8) | (e << 16);
    } else {
      s_off[k] = 0;
      s_abe[k] = -1;
    }
  }
  __syncthreads();
  const int kv = p.Kp / 8;
  const long long total = p.P * kv;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % kv) * 8;
    long long r = i / kv;
    const long long pos = r;
    const int ow = (int)(r % p.oW); r /= p.oW;
    const int oh = (int)(r % p.oH); r /= p.oH;
    const int od = (int)(r % p.oD);
    const int n = (int)(r / p.oD);
    const int d0 = od * p.sd - p.pd, h0 = oh * p.sh - p.ph, w0 = ow * p.sw - p.pw;
    const long long base = ((((long long)n * p.D + d0) * p.H + h0) * p.W + w0) * p.C;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int abe = s_abe[k0 + j];
      const int id = d0 + (abe & 255), ih = h0 + ((abe >> 8) & 255), iw = w0 + ((abe >> 16) & 255);
      const bool ok = abe >= 0 && id >= 0 && id < p.D && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
      v[j] = ok ? p.x[base + s_off[k0 + j]] : zero;
    }
    *reinterpret_cast<uint4*>(p.col + pos * p.Kp + k0) = *reinterpret_cast<const uint4*>(v);
  }
}
// wB[n][k] = w_tf[k][n] (k < K), zero padding to [co_pad][Kp]
__global__ void __launch_bounds__(256) im2col_pack_w_kernel(const float* __restrict__ w, int K, int Kp, int co, int co_pad,
                                                             __nv_bfloat16* __restrict__ wb) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < co_pad * Kp; i += gridDim.x * blockDim.x) {
    const int n = i / Kp, k = i % Kp;
    wb[i] = __float2bfloat16_rn((n < co && k < K) ? w[(long long)k * co + n] : 0.f);
  }
}
__global__ void __launch_bounds__(256) im2col_fold_dw_kernel(const float* __restrict__ d, long long n, float* dw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dw[i] += d[i];
}

void fill_simt_common(SimtGeom& sg, const ConvGeom& g, bool gather_is_conv_like) {
  for (int i = 0; i < 3; ++i) {
    if (gather_is_conv_like) {
      sg.mul[i] = g.d[i].s; sg.off0[i] = -g.d[i].pb; sg.offk[i] = 1; sg.div[i] = 1;
    } else {
      sg.mul[i] = 1; sg.off0[i] = g.d[i].pb; sg.offk[i] = -1; sg.div[i] = g.d[i].s;
    }
  }
}

}  // namespace

extern "C" {

int sap3d_conv_out_dims(const sap3d_conv_desc* d, int32_t* out_dhw) {
  if (check_desc(d)) return 1;
  ConvGeom g;
  make_geom(d, g);
  out_dhw[0] = g.d[0].O;
  out_dhw[1] = g.d[1].O;
  out_dhw[2] = g.d[2].O;
  return 0;
}

int sap3d_conv_stats_rows(const sap3d_conv_desc* d) {
  if (check_desc(d)) return -1;
  ConvGeom g;
  make_geom(d, g);
  if (tc_eligible(d)) {
    TcProblem pb;
    build_fwd_problem(d, g, nullptr, nullptr, pb);
    return (int)pb.classes.size() * tc_plan_tiles(pb);
  }
  if (im2col_eligible(d)) return (int)((im2col_layout(d, g).P + 127) / 128);
  return simt_stats_rows(d, g);
}

size_t sap3d_conv_packed_elems(const sap3d_conv_desc* d, int32_t which) {
  if (check_desc(d)) return 0;
  ConvGeom g;
  make_geom(d, g);
  const size_t co_pad = (size_t)(d->cout + 63) / 64 * 64, ci_pad = (size_t)(g.cin_total + 63) / 64 * 64;
  if (which == 0 && im2col_eligible(d)) return im2col_layout(d, g).total;   // packed filter + scratch + im2col buffer
  if (which == 0) return co_pad * (size_t)g.taps * (size_t)g.cin_total;
  return ci_pad * (size_t)g.taps * (size_t)d->cout;
}

int sap3d_conv_pack_weights(const sap3d_conv_desc* d, const float* w_tf, void* w_fwd, void* w_dgrad, void* stream) {
  if (check_desc(d)) return 1;
  ConvGeom g;
  make_geom(d, g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ci = g.cin_total, co = d->cout, taps = g.taps;
  const int co_pad = (co + 63) / 64 * 64, ci_pad = (ci + 63) / 64 * 64;
  int rc = 0;
  if (!d->transposed) {
    // TF DHWIO: w[tap][ci][co]
    if (w_fwd) rc |= pack_weights_launch(w_tf, w_fwd, taps, co, co_pad, ci, (long long)ci * co, 1, co, st);
    if (w_dgrad) rc |= pack_weights_launch(w_tf, w_dgrad, taps, ci, ci_pad, co, (long long)ci * co, co, 1, st);
  } else {
    // tf.layers.conv3d_transpose kernel: w[tap][co][ci]
    if (w_fwd) rc |= pack_weights_launch(w_tf, w_fwd, taps, co, co_pad, ci, (long long)ci * co, ci, 1, st);
    if (w_dgrad) rc |= pack_weights_launch(w_tf, w_dgrad, taps, ci, ci_pad, co, (long long)ci * co, 1, ci, st);
  }
  if (rc) return set_error("pack_weights launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

static int conv_fwd_impl(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                         const float* bias, void* y, float* stats, const float* ep_scale, const float* ep_shift, int ep_relu, void* stream);

int sap3d_conv_fwd(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                   const float* bias, void* y, float* stats, void* stream) {
  return conv_fwd_impl(d, x0, x1, w_tf, w_fwd_packed, bias, y, stats, nullptr, nullptr, 0, stream);
}

/* 1 when the forward operand buffer (sap3d_conv_packed_elems(d, 0) elements) is a WORKSPACE that sap3d_conv_fwd fills itself
 * (the small-Cin im2col form: packed filter + scratch + im2col matrix, reused by sap3d_conv_wgrad); 0 when it is the
 * pre-packed filter from sap3d_conv_pack_weights that conv_fwd only reads */
int sap3d_conv_fwd_operand_is_workspace(const sap3d_conv_desc* d) {
  if (!d) return 0;
  return im2col_eligible(d) ? 1 : 0;
}

/* 1 when conv_fwd runs on the tensor cores for this descriptor (implicit GEMM or the small-Cin im2col form) */
int sap3d_conv_fwd_on_tensor_cores(const sap3d_conv_desc* d) {
  if (check_desc(d)) return 0;
  return (tc_eligible(d) || im2col_eligible(d)) ? 1 : 0;
}

/* y = relu?((conv(x) + bias) * scale[c] + shift[c]): inference-mode BatchNorm (+ ReLU) folded into the conv epilogue
 * (tensor-core paths only; scale/shift from sap3d_bn_finalize with training = 0) */
int sap3d_conv_fwd_affine(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                          const float* bias, const float* scale, const float* shift, int32_t relu, void* y, void* stream) {
  if (check_desc(d)) return 1;
  if (!scale || !shift) return set_error("conv_fwd_affine: NULL scale / shift");
  if (!(tc_eligible(d) || im2col_eligible(d))) return set_error("conv_fwd_affine: descriptor does not run on the tensor-core path");
  return conv_fwd_impl(d, x0, x1, w_tf, w_fwd_packed, bias, y, nullptr, scale, shift, relu, stream);
}

static int conv_fwd_impl(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                         const float* bias, void* y, float* stats, const float* ep_scale, const float* ep_shift, int ep_relu, void* stream) {
  if (check_desc(d)) return 1;
  if (!x0 || !y || (d->nseg > 1 && !x1)) return set_error("conv_fwd: NULL tensor pointer");
  if (require_device()) return 1;
  ConvGeom g;
  make_geom(d, g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char err[512];
  if (tc_eligible(d)) {
    if (!w_fwd_packed) return set_error("conv_fwd: tensor-core path needs packed weights");
    TcProblem pb;
    build_fwd_problem(d, g, x0, x1, pb);
    pb.B = w_fwd_packed;
    pb.out = y;
    pb.bias = d->has_bias ? bias : nullptr;
    pb.stats = stats;
    pb.scale = ep_scale; pb.shift = ep_shift;
    pb.relu = ep_relu;
    pb.accumulate = 0;
    pb.out_f32 = d->out_f32;
    pb.force_block_n = 0;
    if (tc_launch(pb, st, err, sizeof(err))) return set_error("%s", err);
    return 0;
  }
  if (im2col_eligible(d) && w_fwd_packed && w_tf) {
    const Im2colLayout L = im2col_layout(d, g);
    __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(w_fwd_packed));
    const int co_pad = (d->cout + 63) / 64 * 64;
    im2col_pack_w_kernel<<<32, 256, 0, st>>>(w_tf, L.K, L.Kp, d->cout, co_pad, ws);
    Im2colArgs a;
    a.x = reinterpret_cast<const __nv_bfloat16*>(x0); a.col = ws + L.off_col;
    a.N = d->N; a.D = d->D; a.H = d->H; a.W = d->W; a.C = g.cin_total;
    a.oD = g.d[0].O; a.oH = g.d[1].O; a.oW = g.d[2].O;
    a.kd = d->kd; a.kh = d->kh; a.kw = d->kw; a.sd = d->sd; a.sh = d->sh; a.sw = d->sw;
    a.pd = g.d[0].pb; a.ph = g.d[1].pb; a.pw = g.d[2].pb; a.K = L.K; a.Kp = L.Kp; a.P = L.P;
    long long blocks = (L.P * (L.Kp / 8) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    im2col_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    if (cudaGetLastError() != cudaSuccess) return set_error("conv_fwd: im2col launch failed");
    TcProblem pb;
    TcView v;
    v.base = ws + L.off_col;
    v.C = L.Kp;
    v.dim[0] = (int)L.P; v.dim[1] = 1; v.dim[2] = 1; v.dim[3] = 1;
    v.stride[0] = L.Kp; v.stride[1] = 0; v.stride[2] = 0; v.stride[3] = 0;
    pb.views.push_back(v);
    TcClassH cls;
    cls.out_ofs = 0;
    TcTapH t;
    t.view = 0; t.off[0] = t.off[1] = t.off[2] = t.off[3] = 0; t.kofs = 0; t.c_begin = 0; t.nch = L.Kp;
    cls.taps.push_back(t);
    pb.classes.push_back(cls);
    pb.ext[0] = (int)L.P; pb.ext[1] = 1; pb.ext[2] = 1; pb.ext[3] = 1;
    pb.so[0] = d->cout; pb.so[1] = 0; pb.so[2] = 0; pb.so[3] = 0;
    pb.B = ws; pb.Ktot = L.Kp; pb.rowsB = co_pad; pb.cout = d->cout; pb.out = y;
    pb.bias = d->has_bias ? bias : nullptr; pb.stats = stats; pb.scale = ep_scale; pb.shift = ep_shift;
    pb.relu = ep_relu; pb.accumulate = 0; pb.out_f32 = 0; pb.force_block_n = 0;
    if (tc_launch(pb, st, err, sizeof(err))) return set_error("%s", err);
    return 0;
  }
  if (d->impl == SAP3D_IMPL_TC) return set_error("conv_fwd: descriptor not eligible for the tensor-core path");
  if (!w_tf) return set_error("conv_fwd: CUDA-core path needs the fp32 TF-layout weights");
  SimtGeom sg;
  memset(&sg, 0, sizeof(sg));
  sg.x[0] = x0; sg.x[1] = x1 ? x1 : x0;
  sg.cseg[0] = d->cin[0]; sg.cseg[1] = d->nseg > 1 ? d->cin[1] : 0;
  sg.cin_total = g.cin_total;
  sg.N = d->N; sg.iD = d->D; sg.iH = d->H; sg.iW = d->W;
  sg.oD = g.d[0].O; sg.oH = g.d[1].O; sg.oW = g.d[2].O;
  sg.kd = d->kd; sg.kh = d->kh; sg.kw = d->kw;
  fill_simt_common(sg, g, !d->transposed);
  sg.w = w_tf;
  sg.ws_tap = (long long)g.cin_total * d->cout;
  if (!d->transposed) { sg.ws_ci = d->cout; sg.ws_co = 1; }
  else { sg.ws_ci = 1; sg.ws_co = g.cin_total; }
  sg.co_off = 0;
  sg.bias = d->has_bias ? bias : nullptr;
  sg.cout = d->cout;
  sg.y = y;
  out_strides(g, d->cout, sg.yo);
  sg.accumulate = 0;
  if (simt_conv_launch(sg, d->dtype, d->out_f32, st, err, sizeof(err))) return set_error("%s", err);
  if (stats) {
    long long P = (long long)d->N * sg.oD * sg.oH * sg.oW;
    if (col_stats_launch(y, d->out_f32 ? SAP3D_F32 : d->dtype, P, d->cout, simt_stats_rows(d, g), stats, st))
      return set_error("col_stats launch failed");
  }
  return 0;
}

int sap3d_conv_fwd_bn_supported(const sap3d_conv_desc* d) {
  if (check_desc(d)) return 0;
  if (!tc_eligible(d) || d->out_f32 || d->dtype != SAP3D_BF16 || require_device()) return 0;
  ConvGeom g;
  make_geom(d, g);
  TcProblem pb;
  build_fwd_problem(d, g, nullptr, nullptr, pb);
  pb.B = nullptr; pb.out = nullptr; pb.bias = nullptr; pb.stats = nullptr; pb.scale = nullptr; pb.shift = nullptr;
  pb.relu = 0; pb.accumulate = 0; pb.out_f32 = 0; pb.force_block_n = 0;
  pb.query_fuse_bn = 1;
  char err[256];
  return tc_launch(pb, nullptr, err, sizeof(err)) == 0 ? 1 : 0;
}

int sap3d_conv_fwd_bn(const sap3d_conv_desc* d, const void* x0, const void* x1, const float* w_tf, const void* w_fwd_packed,
                      const float* bias, void* raw, float* stats, const sap3d_bn_fuse* f, void* stream) {
  (void)w_tf;
  if (check_desc(d)) return 1;
  if (!x0 || !raw || !stats || !f || (d->nseg > 1 && !x1)) return set_error("conv_fwd_bn: NULL argument");
  if (!f->y || !f->scale || !f->shift) return set_error("conv_fwd_bn: y / scale / shift must be given");
  if (require_device()) return 1;
  if (!tc_eligible(d) || d->out_f32 || d->dtype != SAP3D_BF16) return set_error("conv_fwd_bn: bf16 tensor-core descriptors only");
  if (!w_fwd_packed) return set_error("conv_fwd_bn: needs the packed weights");
  ConvGeom g;
  make_geom(d, g);
  TcProblem pb;
  build_fwd_problem(d, g, x0, x1, pb);
  pb.B = w_fwd_packed;
  pb.out = raw;
  pb.bias = d->has_bias ? bias : nullptr;
  pb.stats = stats;
  pb.scale = nullptr; pb.shift = nullptr;
  pb.relu = 0; pb.accumulate = 0; pb.out_f32 = 0; pb.force_block_n = 0;
  TcFuseBN fb;
  fb.gamma = f->gamma; fb.beta = f->beta; fb.moving_mean = f->moving_mean; fb.moving_var = f->moving_var;
  fb.momentum = f->momentum; fb.eps = f->eps;
  fb.count = (double)d->N * g.d[0].O * g.d[1].O * g.d[2].O;
  fb.scale = f->scale; fb.shift = f->shift; fb.mean = f->mean; fb.rstd = f->rstd;
  fb.relu1 = f->relu1; fb.residual = f->residual; fb.relu_out = f->relu_out; fb.y = f->y;
  pb.fuse_bn = &fb;
  char err[512];
  if (tc_launch(pb, reinterpret_cast<cudaStream_t>(stream), err, sizeof(err))) return set_error("%s", err);
  return 0;
}

static int conv_dgrad_impl(const sap3d_conv_desc* d, int32_t seg, const void* dy, const float* w_tf, const void* w_dgrad_packed,
                           void* dx, int32_t accumulate, void* dx2, int32_t accumulate2, void* stream);

int sap3d_conv_dgrad(const sap3d_conv_desc* d, int32_t seg, const void* dy, const float* w_tf, const void* w_dgrad_packed,
                     void* dx, int32_t accumulate, void* stream) {
  return conv_dgrad_impl(d, seg, dy, w_tf, w_dgrad_packed, dx, accumulate, nullptr, 0, stream);
}

/* 1 when both segment gradients of a fused-concat conv can come out of ONE launch (sap3d_conv_dgrad2) */
int sap3d_conv_dgrad2_supported(const sap3d_conv_desc* d) {
  if (check_desc(d)) return 0;
  return (d->nseg == 2 && d->cin[0] == d->cin[1] && d->cin[0] % 64 == 0 && tc_dgrad_eligible(d)) ? 1 : 0;
}

/* data gradients of BOTH segments of a fused-concat conv in one launch: dy is read once per tap instead of once per
 * segment (N = cin0 + cin1 output columns; columns >= cin0 are written to dx1) */
int sap3d_conv_dgrad2(const sap3d_conv_desc* d, const void* dy, const float* w_tf, const void* w_dgrad_packed, void* dx0,
                      int32_t accumulate0, void* dx1, int32_t accumulate1, void* stream) {
  if (!sap3d_conv_dgrad2_supported(d)) return set_error("conv_dgrad2: needs two equal 64-aligned segments on the tensor-core path");
  if (!dx1) return set_error("conv_dgrad2: NULL tensor pointer");
  return conv_dgrad_impl(d, 0, dy, w_tf, w_dgrad_packed, dx0, accumulate0, dx1, accumulate1, stream);
}

static int conv_dgrad_impl(const sap3d_conv_desc* d, int32_t seg, const void* dy, const float* w_tf, const void* w_dgrad_packed,
                           void* dx, int32_t accumulate, void* dx2, int32_t accumulate2, void* stream) {
  if (check_desc(d)) return 1;
  if (seg < 0 || seg >= d->nseg) return set_error("conv_dgrad: bad segment");
  if (!dy || !dx) return set_error("conv_dgrad: NULL tensor pointer");
  if (require_device()) return 1;
  ConvGeom g;
  make_geom(d, g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char err[512];
  const int cseg = d->cin[seg];
  const size_t esize = d->dtype == SAP3D_BF16 ? 2 : 4;
  long long si[4];
  in_strides(g, cseg, si);
  bool strided_scatter = false;  // conv with stride > 1: only a sub-lattice of dx is written
  if (!d->transposed)
    for (int i = 0; i < 3; ++i)
      if (g.d[i].s != 1) strided_scatter = true;

  if (tc_dgrad_eligible(d)) {
    if (!w_dgrad_packed) return set_error("conv_dgrad: tensor-core path needs packed weights");
    TcProblem pb;
    const int Ktot = g.taps * d->cout;
    if (!d->transposed) {
      if (strided_scatter && !accumulate) {
        if (cudaMemsetAsync(dx, 0, (size_t)si[3] * d->N * esize, st) != cudaSuccess) return set_error("conv_dgrad: memset failed");
      }
      if (strided_scatter && dx2 && !accumulate2) {
        if (cudaMemsetAsync(dx2, 0, (size_t)si[3] * d->N * esize, st) != cudaSuccess) return set_error("conv_dgrad: memset failed");
      }
      TcView v;
      v.base = dy;
      v.C = d->cout;
      long long so[4];
      out_strides(g, d->cout, so);
      for (int i = 0; i < 3; ++i) {
        v.dim[i] = g.d[2 - i].O;
        v.stride[i] = so[i];
      }
      v.dim[3] = d->N;
      v.stride[3] = so[3];
      pb.views.push_back(v);
      TcClassH cls;
      cls.out_ofs = 0;
      for (int kd = 0; kd < d->kd; ++kd)
        for (int kh = 0; kh < d->kh; ++kh)
          for (int kw = 0; kw < d->kw; ++kw) {
            TcTapH t;
            t.view = 0;
            // stride-1 dims: o = i + pb - k ; strided dims have k == 1, pb == 0: o = i / s handled by the output view
            t.off[0] = (g.d[2].s == 1) ? g.d[2].pb - kw : 0;
            t.off[1] = (g.d[1].s == 1) ? g.d[1].pb - kh : 0;
            t.off[2] = (g.d[0].s == 1) ? g.d[0].pb - kd : 0;
            t.off[3] = 0;
            t.kofs = ((kd * d->kh + kh) * d->kw + kw) * d->cout;
            t.c_begin = 0;
            t.nch = d->cout;
            cls.taps.push_back(t);
          }
      pb.classes.push_back(cls);
      // class-local coordinates = dy coordinates; dx written at i = o*s
      pb.ext[0] = g.d[2].O; pb.ext[1] = g.d[1].O; pb.ext[2] = g.d[0].O; pb.ext[3] = d->N;
      pb.so[0] = si[0] * g.d[2].s; pb.so[1] = si[1] * g.d[1].s; pb.so[2] = si[2] * g.d[0].s; pb.so[3] = si[3];
    } else {
      // dx[i] = sum_k dy[s*i + k - pb] * W[k]: parity views of dy
      long long so[4];
      out_strides(g, d->cout, so);
      std::map<int, int> view_of;  // key = (rd*16+rh)*16+rw
      TcClassH cls;
      cls.out_ofs = 0;
      for (int kd = 0; kd < d->kd; ++kd)
        for (int kh = 0; kh < d->kh; ++kh)
          for (int kw = 0; kw < d->kw; ++kw) {
            const int kk[3] = {kd, kh, kw};
            int r[3], q[3];
            for (int i = 0; i < 3; ++i) {
              const int e = kk[i] - g.d[i].pb;
              q[i] = floordiv(e, g.d[i].s);
              r[i] = e - q[i] * g.d[i].s;
            }
            const int key = (r[0] * 16 + r[1]) * 16 + r[2];
            if (view_of.find(key) == view_of.end()) {
              TcView v;
              v.base = reinterpret_cast<const char*>(dy) + (size_t)(r[0] * so[2] + r[1] * so[1] + r[2] * so[0]) * 2;
              v.C = d->cout;
              for (int i = 0; i < 3; ++i) {
                const DimGeom& dg = g.d[2 - i];
                v.dim[i] = (dg.O - r[2 - i] + dg.s - 1) / dg.s;
                v.stride[i] = so[i] * dg.s;
              }
              v.dim[3] = d->N;
              v.stride[3] = so[3];
              view_of[key] = (int)pb.views.size();
              pb.views.push_back(v);
            }
            TcTapH t;
            t.view = view_of[key];
            t.off[0] = q[2]; t.off[1] = q[1]; t.off[2] = q[0]; t.off[3] = 0;
            t.kofs = ((kd * d->kh + kh) * d->kw + kw) * d->cout;
            t.c_begin = 0;
            t.nch = d->cout;
            cls.taps.push_back(t);
          }
      pb.classes.push_back(cls);
      pb.ext[0] = g.d[2].I; pb.ext[1] = g.d[1].I; pb.ext[2] = g.d[0].I; pb.ext[3] = d->N;
      for (int i = 0; i < 4; ++i) pb.so[i] = si[i];
    }
    pb.Ktot = Ktot;
    pb.B = reinterpret_cast<const char*>(w_dgrad_packed) + (size_t)g.seg_off[seg] * Ktot * 2;
    pb.rowsB = (cseg + 63) / 64 * 64;
    pb.cout = cseg;
    pb.out = dx;
    if (dx2) {   // both segments: N = cin0 + cin1 columns of the packed data-gradient filter, split at cin0
      pb.rowsB = (g.cin_total + 63) / 64 * 64;
      pb.cout = g.cin_total;
      pb.out2 = dx2;
      pb.seg_split = d->cin[0];
      pb.accumulate2 = accumulate2;
    }
    pb.bias = nullptr;
    pb.stats = nullptr;
    pb.scale = pb.shift = nullptr;
    pb.relu = 0;
    pb.accumulate = accumulate;
    pb.out_f32 = 0;
    pb.force_block_n = 0;
    if (tc_launch(pb, st, err, sizeof(err))) return set_error("%s", err);
    return 0;
  }
  if (d->impl == SAP3D_IMPL_TC) return set_error("conv_dgrad: descriptor not eligible for the tensor-core path");
  if (!w_tf) return set_error("conv_dgrad: CUDA-core path needs the fp32 TF-layout weights");
  SimtGeom sg;
  memset(&sg, 0, sizeof(sg));
  sg.x[0] = dy; sg.x[1] = dy;
  sg.cseg[0] = d->cout; sg.cseg[1] = 0;
  sg.cin_total = d->cout;
  sg.N = d->N;
  sg.iD = g.d[0].O; sg.iH = g.d[1].O; sg.iW = g.d[2].O;  // gathered tensor = dy
  sg.oD = d->D; sg.oH = d->H; sg.oW = d->W;              // produced tensor = dx
  sg.kd = d->kd; sg.kh = d->kh; sg.kw = d->kw;
  fill_simt_common(sg, g, d->transposed != 0);
  sg.w = w_tf;
  sg.ws_tap = (long long)g.cin_total * d->cout;
  if (!d->transposed) { sg.ws_ci = 1; sg.ws_co = d->cout; }       // w[k][ci][co]: K index = co, out = ci
  else { sg.ws_ci = g.cin_total; sg.ws_co = 1; }                   // w[k][co][ci]
  sg.co_off = g.seg_off[seg];
  sg.bias = nullptr;
  sg.cout = cseg;
  sg.y = dx;
  for (int i = 0; i < 4; ++i) sg.yo[i] = si[i];
  sg.accumulate = accumulate;
  if (simt_conv_launch(sg, d->dtype, 0, st, err, sizeof(err))) return set_error("%s", err);
  return 0;
}

int sap3d_conv_wgrad(const sap3d_conv_desc* d, const void* x0, const void* x1, const void* dy, float* dw, float* db,
                     const void* fwd_operand, void* stream) {
  if (check_desc(d)) return 1;
  if (!x0 || !dy || !dw) return set_error("conv_wgrad: NULL pointer");
  if (require_device()) return 1;
  ConvGeom g;
  make_geom(d, g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char err[512];
  if (im2col_eligible(d) && fwd_operand) {
    // the forward pass left im2col(x) in the operand buffer: dW[k][co] = sum_pos col[pos][k] dy[pos][co]
    const Im2colLayout L = im2col_layout(d, g);
    __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(fwd_operand));
    float* dsc = reinterpret_cast<float*>(ws + L.off_d);
    if (cudaMemsetAsync(dsc, 0, (size_t)L.Kp * d->cout * sizeof(float), st) != cudaSuccess) return set_error("conv_wgrad: memset failed");
    if (sap3d_gemm_tn(ws + L.off_col, L.Kp, dy, d->cout, dsc, d->cout, L.Kp, d->cout, (int32_t)L.P, stream)) return 1;
    im2col_fold_dw_kernel<<<64, 256, 0, st>>>(dsc, (long long)L.K * d->cout, dw);
    if (cudaGetLastError() != cudaSuccess) return set_error("conv_wgrad: fold launch failed");
    if (db) {
      if (col_sum_accum_launch(dy, d->dtype, L.P, d->cout, db, st)) return set_error("bias-grad launch failed");
    }
    return 0;
  }
  const void* xs[2] = {x0, x1};
  const bool use_tc = tc_dgrad_eligible(d);  // same constraints: bf16, channels % 64, <= 10 parity views
  for (int s = 0; s < d->nseg; ++s) {
    if (use_tc) {
      TcWgradProblem pb;
      long long so[4], si[4];
      out_strides(g, d->cout, so);
      in_strides(g, d->cin[s], si);
      if (!d->transposed) {
        TcView pv;
        pv.base = xs[s];
        pv.C = d->cin[s];
        for (int i = 0; i < 3; ++i) {
          const DimGeom& dg = g.d[2 - i];
          pv.dim[i] = (dg.s == 1) ? dg.I : dg.O;
          pv.stride[i] = si[i] * dg.s;
        }
        pv.dim[3] = d->N;
        pv.stride[3] = si[3];
        pb.pviews.push_back(pv);
        pb.q.base = dy;
        pb.q.C = d->cout;
        for (int i = 0; i < 3; ++i) {
          pb.q.dim[i] = g.d[2 - i].O;
          pb.q.stride[i] = so[i];
        }
        pb.q.dim[3] = d->N;
        pb.q.stride[3] = so[3];
        pb.ext[0] = g.d[2].O; pb.ext[1] = g.d[1].O; pb.ext[2] = g.d[0].O; pb.ext[3] = d->N;
        for (int kd = 0; kd < d->kd; ++kd)
          for (int kh = 0; kh < d->kh; ++kh)
            for (int kw = 0; kw < d->kw; ++kw) {
              TcWgradTap t;
              t.view = 0;
              t.off[0] = (g.d[2].s == 1) ? kw - g.d[2].pb : 0;
              t.off[1] = (g.d[1].s == 1) ? kh - g.d[1].pb : 0;
              t.off[2] = (g.d[0].s == 1) ? kd - g.d[0].pb : 0;
              t.off[3] = 0;
              t.dw_ofs = (long long)((kd * d->kh + kh) * d->kw + kw) * g.cin_total * d->cout + (long long)g.seg_off[s] * d->cout;
              pb.taps.push_back(t);
            }
        pb.M = d->cin[s];
        pb.N = d->cout;
        pb.ldw = d->cout;
      } else {
        std::map<int, int> view_of;
        for (int kd = 0; kd < d->kd; ++kd)
          for (int kh = 0; kh < d->kh; ++kh)
            for (int kw = 0; kw < d->kw; ++kw) {
              const int kk[3] = {kd, kh, kw};
              int r[3], q[3];
              for (int i = 0; i < 3; ++i) {
                const int e = kk[i] - g.d[i].pb;
                q[i] = floordiv(e, g.d[i].s);
                r[i] = e - q[i] * g.d[i].s;
              }
              const int key = (r[0] * 16 + r[1]) * 16 + r[2];
              if (view_of.find(key) == view_of.end()) {
                TcView v;
                v.base = reinterpret_cast<const char*>(dy) + (size_t)(r[0] * so[2] + r[1] * so[1] + r[2] * so[0]) * 2;
                v.C = d->cout;
                for (int i = 0; i < 3; ++i) {
                  const DimGeom& dg = g.d[2 - i];
                  v.dim[i] = (dg.O - r[2 - i] + dg.s - 1) / dg.s;
                  v.stride[i] = so[i] * dg.s;
                }
                v.dim[3] = d->N;
                v.stride[3] = so[3];
                view_of[key] = (int)pb.pviews.size();
                pb.pviews.push_back(v);
              }
              TcWgradTap t;
              t.view = view_of[key];
              t.off[0] = q[2]; t.off[1] = q[1]; t.off[2] = q[0]; t.off[3] = 0;
              t.dw_ofs = (long long)((kd * d->kh + kh) * d->kw + kw) * g.cin_total * d->cout + g.seg_off[s];
              pb.taps.push_back(t);
            }
        pb.q.base = xs[s];
        pb.q.C = d->cin[s];
        for (int i = 0; i < 3; ++i) {
          pb.q.dim[i] = g.d[2 - i].I;
          pb.q.stride[i] = si[i];
        }
        pb.q.dim[3] = d->N;
        pb.q.stride[3] = si[3];
        pb.ext[0] = g.d[2].I; pb.ext[1] = g.d[1].I; pb.ext[2] = g.d[0].I; pb.ext[3] = d->N;
        pb.M = d->cout;
        pb.N = d->cin[s];
        pb.ldw = g.cin_total;
      }
      pb.p_c_begin = 0;
      pb.dw = dw;
      if (tc_wgrad_launch(pb, st, err, sizeof(err))) return set_error("%s", err);
      continue;
    }
    SimtWgradGeom wg;
    memset(&wg, 0, sizeof(wg));
    wg.N = d->N;
    wg.kd = d->kd; wg.kh = d->kh; wg.kw = d->kw;
    for (int i = 0; i < 3; ++i) { wg.mul[i] = g.d[i].s; wg.off0[i] = -g.d[i].pb; wg.offk[i] = 1; wg.div[i] = 1; }
    wg.dw = dw;
    wg.ws_tap = (long long)g.cin_total * d->cout;
    if (!d->transposed) {
      wg.g = xs[s]; wg.cg = d->cin[s]; wg.gD = d->D; wg.gH = d->H; wg.gW = d->W;
      wg.q = dy; wg.cq = d->cout; wg.qD = g.d[0].O; wg.qH = g.d[1].O; wg.qW = g.d[2].O;
      wg.ws_g = d->cout; wg.g_c0 = g.seg_off[s]; wg.q_c0 = 0;
    } else {
      wg.g = dy; wg.cg = d->cout; wg.gD = g.d[0].O; wg.gH = g.d[1].O; wg.gW = g.d[2].O;
      wg.q = xs[s]; wg.cq = d->cin[s]; wg.qD = d->D; wg.qH = d->H; wg.qW = d->W;
      wg.ws_g = g.cin_total; wg.g_c0 = 0; wg.q_c0 = g.seg_off[s];
    }
    if (simt_wgrad_launch(wg, d->dtype, st, err, sizeof(err))) return set_error("%s", err);
  }
  if (db) {
    long long P = (long long)d->N * g.d[0].O * g.d[1].O * g.d[2].O;
    if (col_sum_accum_launch(dy, d->dtype, P, d->cout, db, st)) return set_error("bias-grad launch failed");
  }
  return 0;
}

}  // extern "C"

// ---- plain GEMMs on the same tensor-core kernels (attention matmuls, utils/network.py:184,186) ----
extern "C" {

static int gemm_nt_impl(const void* A, int64_t lda, int64_t stride_a, const void* B, int64_t ldb, int64_t stride_b, int32_t rows_b,
                        void* Cout, int64_t ldc, int64_t stride_c, int32_t M, int32_t N, int32_t K, int32_t batch, int32_t out_f32,
                        int32_t accumulate, void* stream) {
  if (require_device()) return 1;
  if (!A || !B || !Cout) return set_error("gemm_nt: NULL pointer");
  if (K % 64 != 0 || ldb != K || N % 8 != 0 || lda % 8 != 0 || ldc % 8 != 0 || rows_b < 1 || rows_b > N)
    return set_error("gemm_nt: need K %% 64 == 0, ldb == K, N %% 8 == 0, lda/ldc %% 8 == 0, rows_b <= N (M=%d N=%d K=%d lda=%lld ldb=%lld)",
                     M, N, K, (long long)lda, (long long)ldb);
  if (batch < 1 || (batch > 1 && (stride_a % 8 != 0 || stride_b % 8 != 0 || stride_c % 8 != 0)))
    return set_error("gemm_nt: batch >= 1 and batch strides %% 8 == 0 required");
  TcProblem pb;
  TcView v;
  v.base = A;
  v.C = K;
  v.dim[0] = M; v.dim[1] = 1; v.dim[2] = 1; v.dim[3] = batch;
  v.stride[0] = lda; v.stride[1] = 0; v.stride[2] = 0; v.stride[3] = batch > 1 ? stride_a : 0;
  pb.views.push_back(v);
  TcClassH cls;
  cls.out_ofs = 0;
  TcTapH t;
  t.view = 0;
  t.off[0] = t.off[1] = t.off[2] = t.off[3] = 0;
  t.kofs = 0;
  t.c_begin = 0;
  t.nch = K;
  cls.taps.push_back(t);
  pb.classes.push_back(cls);
  pb.ext[0] = M; pb.ext[1] = 1; pb.ext[2] = 1; pb.ext[3] = batch;
  pb.so[0] = ldc; pb.so[1] = 0; pb.so[2] = 0; pb.so[3] = batch > 1 ? stride_c : 0;
  pb.B = B;
  pb.Ktot = K;
  pb.rowsB = rows_b;  // rows beyond rows_b are zero-filled by the TMA unit
  pb.cout = N;
  pb.out = Cout;
  pb.bias = nullptr; pb.stats = nullptr; pb.scale = nullptr; pb.shift = nullptr;
  pb.relu = 0; pb.accumulate = accumulate; pb.out_f32 = out_f32; pb.force_block_n = 0;
  pb.b_batch = batch;
  pb.b_batch_stride = stride_b;
  char err[512];
  if (tc_launch(pb, reinterpret_cast<cudaStream_t>(stream), err, sizeof(err))) return set_error("%s", err);
  return 0;
}

int sap3d_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t rows_b, void* Cout, int64_t ldc, int32_t M,
                  int32_t N, int32_t K, int32_t out_f32, int32_t accumulate, void* stream) {
  return gemm_nt_impl(A, lda, 0, B, ldb, 0, rows_b, Cout, ldc, 0, M, N, K, 1, out_f32, accumulate, stream);
}

int sap3d_gemm_nt_batched(const void* A, int64_t lda, int64_t stride_a, const void* B, int64_t ldb, int64_t stride_b, int32_t rows_b,
                          void* Cout, int64_t ldc, int64_t stride_c, int32_t M, int32_t N, int32_t K, int32_t batch, int32_t out_f32,
                          int32_t accumulate, void* stream) {
  return gemm_nt_impl(A, lda, stride_a, B, ldb, stride_b, rows_b, Cout, ldc, stride_c, M, N, K, batch, out_f32, accumulate, stream);
}

int sap3d_gemm_tn(const void* P, int64_t ldp, const void* Q, int64_t ldq, float* D, int64_t ldd, int32_t M, int32_t N,
                  int32_t Kpos, void* stream) {
  if (require_device()) return 1;
  if (!P || !Q || !D) return set_error("gemm_tn: NULL pointer");
  if (M % 64 != 0 || N % 64 != 0 || ldp % 8 != 0 || ldq % 8 != 0) return set_error("gemm_tn: M, N %% 64 and ldp, ldq %% 8 required");
  TcWgradProblem pb;
  TcView pv;
  pv.base = P;
  pv.C = M;
  pv.dim[0] = Kpos; pv.dim[1] = 1; pv.dim[2] = 1; pv.dim[3] = 1;
  pv.stride[0] = ldp; pv.stride[1] = 0; pv.stride[2] = 0; pv.stride[3] = 0;
  pb.pviews.push_back(pv);
  pb.q = pv;
  pb.q.base = Q;
  pb.q.C = N;
  pb.q.stride[0] = ldq;
  TcWgradTap t;
  t.view = 0;
  t.off[0] = t.off[1] = t.off[2] = t.off[3] = 0;
  t.dw_ofs = 0;
  pb.taps.push_back(t);
  pb.ext[0] = Kpos; pb.ext[1] = 1; pb.ext[2] = 1; pb.ext[3] = 1;
  pb.M = M;
  pb.N = N;
  pb.p_c_begin = 0;
  pb.ldw = ldd;
  pb.dw = D;
  char err[512];
  if (tc_wgrad_launch(pb, reinterpret_cast<cudaStream_t>(stream), err, sizeof(err))) return set_error("%s", err);
  return 0;
}

}  // extern "C"

extern "C" int sap3d_conv_pack_entries(const sap3d_conv_desc* d, const float* w_tf, void* w_fwd, void* w_dgrad, sap3d_pack_entry* out2) {
  if (check_desc(d)) return -1;
  ConvGeom g;
  make_geom(d, g);
  const int ci = g.cin_total, co = d->cout, taps = g.taps;
  const int co_pad = (co + 63) / 64 * 64, ci_pad = (ci + 63) / 64 * 64;
  int n = 0;
  auto add = [&](void* dst, int rows, int rows_pad, int cols, long long s_r, long long s_c) {
    sap3d_pack_entry& e = out2[n++];
    e.src = w_tf; e.dst = dst; e.taps = taps; e.rows = rows; e.rows_pad = rows_pad; e.cols = cols;
    e.s_tap = (long long)ci * co; e.s_r = s_r; e.s_c = s_c; e.start = 0;
  };
  if (im2col_eligible(d)) w_fwd = nullptr;   // packed inside sap3d_conv_fwd ([cout][Kp], K zero-padded)
  if (!d->transposed) {
    if (w_fwd) add(w_fwd, co, co_pad, ci, 1, co);
    if (w_dgrad) add(w_dgrad, ci, ci_pad, co, co, 1);
  } else {
    if (w_fwd) add(w_fwd, co, co_pad, ci, ci, 1);
    if (w_dgrad) add(w_dgrad, ci, ci_pad, co, 1, ci);
  }
  return n;
}
