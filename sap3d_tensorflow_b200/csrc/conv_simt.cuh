// Geometry structs for the CUDA-core implicit-GEMM kernels (see conv_simt.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace sap3d {

struct SimtGeom {
  const void* x[2];  // gathered tensor segments, [N,iD,iH,iW,cseg[s]]
  int cseg[2];
  int cin_total;     // channels summed over in K (= cseg[0]+cseg[1])
  int N, iD, iH, iW;
  int oD, oH, oW;
  int kd, kh, kw;
  int mul[3], off0[3], offk[3], div[3];  // D,H,W
  const float* w;    // fp32 master weights (TF layout), element (tap,ci,co) at tap*ws_tap+ci*ws_ci+(co+co_off)*ws_co
  long long ws_tap, ws_ci, ws_co;
  int co_off;
  const float* bias;
  int cout;
  void* y;
  long long yo[4];   // output element strides W,H,D,N
  int accumulate;
};
int simt_conv_launch(const SimtGeom& g, int dtype, int out_f32, cudaStream_t stream, char* err, size_t errlen);

struct SimtWgradGeom {
  const void* g;     // gathered tensor [N,gD,gH,gW,cg]
  const void* q;     // position-aligned tensor [N,qD,qH,qW,cq]
  int cg, cq;
  int N, gD, gH, gW, qD, qH, qW;
  int kd, kh, kw;
  int mul[3], off0[3], offk[3], div[3];
  float* dw;         // element (tap, gch, qch) at tap*ws_tap + (gch+g_c0)*ws_g + qch + q_c0
  long long ws_tap, ws_g;
  int g_c0, q_c0;
};
int simt_wgrad_launch(const SimtWgradGeom& g, int dtype, cudaStream_t stream, char* err, size_t errlen);

int col_stats_launch(const void* x, int dtype, long long P, int C, int rows, float* out, cudaStream_t stream);
int col_sum_accum_launch(const void* x, int dtype, long long P, int C, float* out, cudaStream_t stream);
int pack_weights_launch(const float* w, void* out, int taps, int rows, int rows_pad, int cols, long long s_tap,
                        long long s_r, long long s_c, cudaStream_t stream);

}  // namespace sap3d
