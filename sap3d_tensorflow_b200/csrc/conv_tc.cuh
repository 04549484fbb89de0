// Host-side description of a tensor-core implicit-GEMM problem ("form F"):
//   out[cls-local position o, col] = sum over taps t of class, channels c:
//        view[t.view](o + t.offset, t.c_begin + c) * B[col][t.kofs + c]
// Every conv / transposed conv / data-gradient of the P3D graph is lowered to this form by abi.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace sap3d {

struct TcView {
  const void* base;   // bf16, channels contiguous
  int C;              // channels addressable through this view
  int dim[4];         // extents along W,H,D,N of the view
  long long stride[4];// element strides along W,H,D,N
};
struct TcTapH {
  int view;
  int off[4];         // coordinate offset added to the class-local output position (W,H,D,N)
  int kofs;           // first K index of this tap in B
  int c_begin, nch;   // channel range of the view consumed by the tap (nch % 64 == 0)
};
struct TcClassH {
  std::vector<TcTapH> taps;
  long long out_ofs;  // element offset of the class inside the output tensor
};
// batch-statistics BatchNorm (+ReLU, + residual) finished INSIDE the convolution kernel (conv_tc_kernel, training graphs):
// the raw output and its statistics rows are still written (the backward pass reads them), then the CTAs of the launch meet at
// a grid barrier, every CTA finalises the statistics of its own columns and writes the normalised tensor from the accumulators
// it still holds in TMEM.  Can replace the sap3d_bn_apply_fused launch that follows every backbone convolution.
// MEASURED NEGATIVE (r02): 140 launches per training step disappear, but the step gets SLOWER (18.2 vs 17.8 ms): the fused tail
// is a chain of global round trips (statistics store + fence, barrier arrive, barrier observe, statistics loads, residual
// load) of ~0.7 us each, ~6.5 us per layer, while the separate kernel -- launched programmatically, its set-up overlapping the
// convolution's tail -- adds ~5.5 us.  The engine therefore uses it only on request (SAP3D_CONV_FUSE_BN=1); the entry point and
// its parity tests stay.
struct TcFuseBN {
  const float* gamma;
  const float* beta;
  float* moving_mean;      // updated with `momentum` (1.0 = leave alone), like sap3d_bn_finalize
  float* moving_var;
  float momentum, eps;
  double count;            // positions the statistics cover
  float* scale;            // per-channel outputs, [cout] each (the backward pass and inference read them)
  float* shift;
  float* mean;
  float* rstd;
  int relu1;               // y = relu_out?( relu1?(raw * scale + shift) + residual )
  const void* residual;    // nullable, bf16, same layout as the output
  int relu_out;
  void* y;                 // normalised output, bf16, same layout as the raw output
};

struct TcProblem {
  std::vector<TcView> views;
  std::vector<TcClassH> classes;
  int ext[4];         // class-local output extents W,H,D,N
  long long so[4];    // output element strides for class-local coordinates
  const void* B;      // bf16 [rowsB][Ktot], K contiguous
  int Ktot, rowsB;
  int cout;           // valid output columns (multiple of 8)
  void* out;          // bf16 (or f32 when out_f32)
  void* out2 = nullptr;   // optional second output tensor for columns >= seg_split (merged two-segment data gradient)
  int seg_split = 0, accumulate2 = 0;
  const float* bias;  // nullable, [cout]
  float* stats;       // nullable, [ncls*m_tiles][2][cout]
  const float* scale; // nullable epilogue affine (inference-folded norm)
  const float* shift;
  int relu, accumulate, out_f32;
  int force_block_n;  // 0 = auto
  int force_mt = 0;   // 0 = auto, 1 / 2 = M sub-tiles per CTA
  int force_split = 0;// 0 = auto, 2 / 4 = split-K cluster size, -1 = never
  // batched GEMM (one B matrix per sample): ext[3] is the batch; sample n uses B + n * b_batch_stride.  Row tiles then never
  // span samples and dims are not merged.  1 = one shared B (every convolution).
  int b_batch = 1;
  long long b_batch_stride = 0;   // elements
  const TcFuseBN* fuse_bn = nullptr;   // finish a batch-statistics BatchNorm inside the launch (see TcFuseBN); tc_launch fails if it cannot
  int query_fuse_bn = 0;               // 1: tc_launch only answers whether fuse_bn is possible for this problem: 0 = yes, 2 = no
};

// number of M tiles (per class) the launcher will use for these extents (after dim merging)
int tc_plan_tiles(const TcProblem& pb);
// returns 0 on success; on failure writes a message into err
int tc_launch(const TcProblem& pb, cudaStream_t stream, char* err, size_t errlen);
// phase-timing probe of conv_tc_kernel (developer tool): [cta][16][2] uint64 device buffer, or NULL to switch it off
void tc_set_debug_buffer(void* buf);
// launches that took the halo-tile persistent kernel so far (process-wide)
long long tc_halo_launches();
// ... and, of those, the swapped-operand kernel (positions on the N side)
long long tc_swap_launches();

// wgrad: D[m][n] (+)= sum_pos P(pos + off)[m] * Q(pos)[n], fp32 atomics into dw
struct TcWgradTap {
  int view;           // P view
  int off[4];
  long long dw_ofs;   // element offset of this tap's [M][N] block in dw
};
struct TcWgradProblem {
  std::vector<TcView> pviews;
  TcView q;
  std::vector<TcWgradTap> taps;
  int ext[4];         // positions iterated (W,H,D,N), Q coordinates
  int M, N;           // channels of P consumed (rows of dw block), channels of Q (cols)
  int p_c_begin;      // channel offset inside the P views
  long long ldw;      // row stride of the dw block (elements)
  float* dw;
};
int tc_wgrad_launch(const TcWgradProblem& pb, cudaStream_t stream, char* err, size_t errlen);

}  // namespace sap3d
