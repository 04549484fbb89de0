// Micro-probe: cost of a dependent-kernel boundary inside a CUDA graph on B200, with and without programmatic dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int PDL>
__global__ void __launch_bounds__(256) tiny(const float* __restrict__ in, float* __restrict__ out, int n) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * 1.0001f + 1.f;
}

template <int PDL>
int run(int blocks, int chain, float* a, float* b, int n, cudaStream_t st) {
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
    CK(cudaLaunchKernelEx(&cfg, tiny<PDL>, in, out, n));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e0, st));
  for (int w = 0; w < 10; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("PDL=%d blocks=%4d chain=%d: %.3f us per kernel\n", PDL, blocks, chain, ms * 1e3f / (10.f * chain));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return 0;
}

int main() {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int n = 148 * 8 * 256;
  float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4));
  for (int blocks : {1, 16, 148, 148 * 8}) {
    if (run<0>(blocks, 1000, a, b, n, st)) return 1;
    if (run<1>(blocks, 1000, a, b, n, st)) return 1;
  }
  return 0;
}
