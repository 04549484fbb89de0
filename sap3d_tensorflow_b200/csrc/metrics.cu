// Fused saliency metrics CC / SIM / NSS / KLdiv for a batch of (prediction, density, fixation) maps:
// one CTA per map, three passes over L2-resident data, fp64 accumulation, warp-shuffle reductions.
// Replaces the per-clip NumPy calls of the drivers (train.py:254-260, test.py:167-176) for
// utils/metrics.py CC :227, SIM :258, NSS :200, KLdiv :338 (scipy.misc.imresize's uint8 quantisation included).
#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

constexpr int MT = 512;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < MT / 32; ++i) r += sh[i];
  return r;
}
__device__ __forceinline__ float block_minmax(float v, bool is_max, float* sh) {
  v = is_max ? warp_max(v) : warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < MT / 32; ++i) r = is_max ? fmaxf(r, sh[i]) : fminf(r, sh[i]);
  return r;
}

__global__ void __launch_bounds__(MT) metrics_kernel(const float* __restrict__ pred, const float* __restrict__ dens,
                                                      const float* __restrict__ fix, long long n, long long sp, long long sd,
                                                      long long sf, double* __restrict__ out) {
  __shared__ double shd[MT / 32];
  __shared__ float shf[MT / 32];
  const float* p = pred + blockIdx.x * sp;
  const float* d = dens + blockIdx.x * sd;
  const float* f = fix ? fix + blockIdx.x * sf : nullptr;
  // ---- pass 1: moments, extrema, fixation sums
  double s_p = 0, s_pp = 0, s_d = 0, s_dd = 0, s_pd = 0, s_fp = 0, s_f = 0;
  float mn_p = INFINITY, mx_p = -INFINITY, mn_d = INFINITY, mx_d = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += MT) {
    const float a = p[i], b = d[i];
    s_p += a; s_pp += (double)a * a; s_d += b; s_dd += (double)b * b; s_pd += (double)a * b;
    mn_p = fminf(mn_p, a); mx_p = fmaxf(mx_p, a); mn_d = fminf(mn_d, b); mx_d = fmaxf(mx_d, b);
    if (f && f[i] > 0.5f) { s_fp += a; s_f += 1.0; }
  }
  s_p = block_sum(s_p, shd); s_pp = block_sum(s_pp, shd); s_d = block_sum(s_d, shd); s_dd = block_sum(s_dd, shd);
  s_pd = block_sum(s_pd, shd); s_fp = block_sum(s_fp, shd); s_f = block_sum(s_f, shd);
  mn_p = block_minmax(mn_p, false, shf); mx_p = block_minmax(mx_p, true, shf);
  mn_d = block_minmax(mn_d, false, shf); mx_d = block_minmax(mx_d, true, shf);
  const double N = (double)n;
  const double mu_p = s_p / N, mu_d = s_d / N;
  const double sg_p = sqrt(fmax(s_pp / N - mu_p * mu_p, 0.0)), sg_d = sqrt(fmax(s_dd / N - mu_d * mu_d, 0.0));
  const double cc = (s_pd / N - mu_p * mu_d) / (sg_p * sg_d);
  const double nss = (s_fp / s_f - mu_p) / sg_p;
  // ---- pass 2: SIM and the sum of the byte-scaled prediction
  const double rp = (double)mx_p - (double)mn_p, rd = (double)mx_d - (double)mn_d;
  const double sum_rp = (s_p - N * mn_p) / rp, sum_rd = (s_d - N * mn_d) / rd;
  const float cscale = (mx_p - mn_p) == 0.f ? 1.f : (mx_p - mn_p);
  const float bscale = 255.0f / cscale;
  double sim = 0, s_q = 0;
  for (long long i = threadIdx.x; i < n; i += MT) {
    const float a = p[i], b = d[i];
    const double na = ((double)a - mn_p) / rp / sum_rp, nb = ((double)b - mn_d) / rd / sum_rd;
    sim += fmin(na, nb);
    const float bd = fminf(fmaxf((a - mn_p) * bscale, 0.f), 255.f) + 0.5f;
    s_q += (double)(unsigned char)bd;
  }
  sim = block_sum(sim, shd);
  s_q = block_sum(s_q, shd);
  // ---- pass 3: KL divergence
  const double eps = 2.2204e-16;
  double kl = 0;
  for (long long i = threadIdx.x; i < n; i += MT) {
    const float a = p[i], b = d[i];
    const float bd = fminf(fmaxf((a - mn_p) * bscale, 0.f), 255.f) + 0.5f;
    double m1 = (double)(unsigned char)bd;
    if (s_q != 0.0) m1 /= s_q;
    double m2 = (double)b;
    if (s_d != 0.0) m2 /= s_d;
    kl += m2 * log(eps + m2 / (m1 + eps));
  }
  kl = block_sum(kl, shd);
  if (threadIdx.x == 0) {
    out[blockIdx.x * 4 + 0] = cc;
    out[blockIdx.x * 4 + 1] = sim;
    out[blockIdx.x * 4 + 2] = nss;
    out[blockIdx.x * 4 + 3] = kl;
  }
}

}  // namespace

extern "C" int sap3d_saliency_metrics(const float* pred, const float* density, const float* fixation, int32_t n_maps,
                                      int64_t elems_per_map, int64_t pred_stride, int64_t density_stride,
                                      int64_t fixation_stride, double* out, void* stream) {
  if (require_device()) return 1;
  if (!pred || !density || !out) return set_error("saliency_metrics: NULL pointer");
  if (n_maps <= 0) return 0;
  metrics_kernel<<<n_maps, MT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, density, fixation, elems_per_map, pred_stride,
                                                                           density_stride, fixation_stride, out);
  return check_launch("saliency_metrics");
}
