// Test-time evaluation path of test.py:164-183: cv2.resize(prediction, (960, 1080)) (bilinear, half-pixel centres) and the
// fixation-based AUC metrics of utils/metrics.py (AUC_Judd :25-89, AUC_Borji :92-154).
//   AUC_Judd : thresholds = saliency at the fixations (descending); tp = (k+1)/n_fix, fp = (#{S >= t_k} - k - 1)/(n_pix - n_fix);
//              trapezoid area.  The reference's O(n_fix * n_pix) count is kept, one CTA per (map, fixation); the ordering is
//              obtained by ranking instead of sorting.  Optional jitter is a counter-based hash (the reference draws
//              np.random noise, which is not reproducible).
//   AUC_Borji: range-normalised map, n_rep sets of n_fix random locations, thresholds np.arange(0, max, step)[::-1];
//              the random locations come from a splitmix64 counter hash (the reference accepts any `rand_sampler`).
// All bandwidth/latency-trivial next to the network; fp64 accumulation where the reference uses Python floats.
#include "../../include/sap3d.h"
#include "abi_util.cuh"
#include "common.cuh"

using namespace sap3d;

namespace {

constexpr int AUC_CAP = 16384;   // fixations per map the workspace holds

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// ---- cv2.resize(src, (W, H), interpolation=INTER_LINEAR) for float32 single-channel maps -------------------------
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, int n, int h, int w, float* __restrict__ dst, int H,
                                                               int W) {
  const double sx = (double)w / W, sy = (double)h / H;
  const long long total = (long long)n * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % W);
    long long r = i / W;
    const int dy = (int)(r % H);
    const int m = (int)(r / H);
    const double cx = (dx + 0.5) * sx - 0.5, cy = (dy + 0.5) * sy - 0.5;   // fp64 coordinate, fp32 fraction (what cv2 4.x does)
    int x0 = (int)floor(cx), y0 = (int)floor(cy);
    float fx = (float)(cx - x0), fy = (float)(cy - y0);
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= w - 1) { x0 = w - 1; fx = 0.f; }
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= h - 1) { y0 = h - 1; fy = 0.f; }
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float* s = src + (long long)m * h * w;
    // horizontal pass first, then vertical (the order of OpenCV's HResizeLinear / VResizeLinear)
    const float r0 = __fadd_rn(__fmul_rn(s[y0 * w + x0], 1.f - fx), __fmul_rn(s[y0 * w + x1], fx));
    const float r1 = __fadd_rn(__fmul_rn(s[y1 * w + x0], 1.f - fx), __fmul_rn(s[y1 * w + x1], fx));
    dst[i] = __fadd_rn(__fmul_rn(r0, 1.f - fy), __fmul_rn(r1, fy));
  }
}

struct AucWs {
  int* count;      // [n]
  int* idx;        // [n][CAP] pixel index of every fixation (arbitrary order)
  int* above;      // [n][CAP] #{S >= S_fix[k]} ordered by rank
  float* minmax;   // [n][2]
  double* rep_auc; // [n][n_rep]
};

__device__ __forceinline__ float jittered(const float* s, long long i, int jitter, unsigned long long seed, int map) {
  float v = s[i];
  if (jitter) {
    const unsigned long long h = splitmix64(seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)map << 40) + (unsigned long long)i);
    v += (float)((double)(h >> 11) * (1.0 / 9007199254740992.0) * 1e-7);
  }
  return v;
}

// fixation list (atomic append) + range of the map
__global__ void __launch_bounds__(256) auc_collect_kernel(const float* __restrict__ sal, const float* __restrict__ fix, long long E, AucWs w) {
  const int m = blockIdx.y;
  const float* s = sal + (long long)m * E;
  const float* f = fix + (long long)m * E;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < E; i += (long long)gridDim.x * blockDim.x) {
    const float v = s[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
    if (f[i] > 0.5f) {
      const int k = atomicAdd(w.count + m, 1);
      if (k < AUC_CAP) w.idx[(long long)m * AUC_CAP + k] = (int)i;
    }
  }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<int*>(w.minmax + 2 * m), __float_as_int(mn) >= 0 ? __float_as_int(mn) : (int)(0x80000000u - (unsigned)__float_as_int(mn)));
    atomicMax(reinterpret_cast<int*>(w.minmax + 2 * m + 1), __float_as_int(mx) >= 0 ? __float_as_int(mx) : (int)(0x80000000u - (unsigned)__float_as_int(mx)));
  }
}
__device__ __forceinline__ float ordered_int_to_float(int v) { return __int_as_float(v >= 0 ? v : (int)(0x80000000u - (unsigned)v)); }

// one CTA per (fixation, map): count of pixels >= its saliency, rank among the fixations -> above[rank]
__global__ void __launch_bounds__(256) auc_judd_count_kernel(const float* __restrict__ sal, long long E, int jitter, unsigned long long seed, AucWs w) {
  const int m = blockIdx.y;
  const int nf = min(w.count[m], AUC_CAP);
  const int k = blockIdx.x;
  if (k >= nf) return;
  const float* s = sal + (long long)m * E;
  const int* idx = w.idx + (long long)m * AUC_CAP;
  const float t = jittered(s, idx[k], jitter, seed, m);
  int cnt = 0, rank = 0;
  for (long long i = threadIdx.x; i < E; i += blockDim.x) cnt += jittered(s, i, jitter, seed, m) >= t ? 1 : 0;
  for (int j = threadIdx.x; j < nf; j += blockDim.x) {
    const float v = jittered(s, idx[j], jitter, seed, m);
    rank += (v > t || (v == t && j < k)) ? 1 : 0;   // descending order, ties by list position
  }
  __shared__ int sc[8], sr[8];
  for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); rank += __shfl_xor_sync(0xffffffffu, rank, o); }
  if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = cnt; sr[threadIdx.x >> 5] = rank; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, r = 0;
    for (int i = 0; i < 8; ++i) { c += sc[i]; r += sr[i]; }
    w.above[(long long)m * AUC_CAP + r] = c;
  }
}

// trapezoid over (fp, tp) in rank order, one thread per map (n_fix is small)
__global__ void auc_judd_area_kernel(long long E, AucWs w, double* out, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int cnt = w.count[m];
  if (cnt == 0 || cnt > AUC_CAP) { out[m * 2] = nan(""); return; }
  const int* ab = w.above + (long long)m * AUC_CAP;
  const double nf = (double)cnt, nn = (double)(E - cnt);
  double area = 0.0, tp0 = 0.0, fp0 = 0.0;
  for (int k = 0; k < cnt; ++k) {
    const double tp1 = (k + 1) / nf, fp1 = (double)(ab[k] - k - 1) / nn;
    area += (fp1 - fp0) * (tp1 + tp0) * 0.5;
    tp0 = tp1; fp0 = fp1;
  }
  area += (1.0 - fp0) * (1.0 + tp0) * 0.5;
  out[m * 2] = area;
}

// AUC_Borji: one CTA per (repetition, map)
__global__ void __launch_bounds__(256) auc_borji_kernel(const float* __restrict__ sal, long long E, int n_rep, double step, unsigned long long seed,
                                                         AucWs w) {
  const int m = blockIdx.y, rep = blockIdx.x;
  const int nf = min(w.count[m], AUC_CAP);
  if (nf == 0) return;
  const float* s = sal + (long long)m * E;
  const int* idx = w.idx + (long long)m * AUC_CAP;
  const float mn = ordered_int_to_float(reinterpret_cast<const int*>(w.minmax)[2 * m]);
  const float mx = ordered_int_to_float(reinterpret_cast<const int*>(w.minmax)[2 * m + 1]);
  const float rng = mx - mn;
  __shared__ float smax[8];
  __shared__ int s_tp[64], s_fp[64];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_tp[i] = 0; s_fp[i] = 0; }
  float vmax = -INFINITY;
  for (int i = threadIdx.x; i < nf; i += blockDim.x) {
    const float a = (s[idx[i]] - mn) / rng;
    const unsigned long long r = splitmix64(seed * 0x9E3779B97F4A7C15ull + (unsigned long long)i * (unsigned long long)n_rep + rep) % (unsigned long long)E;
    const float b = (s[r] - mn) / rng;
    vmax = fmaxf(vmax, fmaxf(a, b));
  }
  vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = vmax;
  __syncthreads();
  vmax = smax[0];
  for (int i = 1; i < 8; ++i) vmax = fmaxf(vmax, smax[i]);
  // thresholds np.arange(0, vmax, step): t_j = j*step for j < ceil(vmax/step)
  int T = (int)ceil((double)vmax / step);
  if (T > 64) T = 64;
  for (int i = threadIdx.x; i < nf; i += blockDim.x) {
    const double a = (double)((s[idx[i]] - mn) / rng);
    const unsigned long long r = splitmix64(seed * 0x9E3779B97F4A7C15ull + (unsigned long long)i * (unsigned long long)n_rep + rep) % (unsigned long long)E;
    const double b = (double)((s[r] - mn) / rng);
    for (int j = 0; j < T; ++j) {
      const double th = (double)j * step;
      if (a >= th) atomicAdd(&s_tp[j], 1);
      if (b >= th) atomicAdd(&s_fp[j], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double area = 0.0, tp0 = 0.0, fp0 = 0.0;
    for (int j = T - 1; j >= 0; --j) {   // thresholds descending
      const double tp1 = s_tp[j] / (double)nf, fp1 = s_fp[j] / (double)nf;
      area += (fp1 - fp0) * (tp1 + tp0) * 0.5;
      tp0 = tp1; fp0 = fp1;
    }
    area += (1.0 - fp0) * (1.0 + tp0) * 0.5;
    w.rep_auc[(long long)m * n_rep + rep] = area;
  }
}
__global__ void auc_borji_mean_kernel(int n_rep, AucWs w, double* out, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int cnt = w.count[m];
  if (cnt == 0 || cnt > AUC_CAP) { out[m * 2 + 1] = nan(""); return; }
  double s = 0.0;
  for (int r = 0; r < n_rep; ++r) s += w.rep_auc[(long long)m * n_rep + r];
  out[m * 2 + 1] = s / n_rep;
}

size_t ws_layout(int n, int n_rep, AucWs* w, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  const size_t o_cnt = take((size_t)n * 4 + (size_t)n * 8);   // counts + min/max (zero / sentinel initialised together)
  const size_t o_idx = take((size_t)n * AUC_CAP * 4);
  const size_t o_ab = take((size_t)n * AUC_CAP * 4);
  const size_t o_rep = take((size_t)n * n_rep * 8);
  if (w) {
    w->count = reinterpret_cast<int*>(base + o_cnt);
    w->minmax = reinterpret_cast<float*>(base + o_cnt + (size_t)n * 4);
    w->idx = reinterpret_cast<int*>(base + o_idx);
    w->above = reinterpret_cast<int*>(base + o_ab);
    w->rep_auc = reinterpret_cast<double*>(base + o_rep);
  }
  return off;
}

__global__ void auc_init_kernel(AucWs w, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  w.count[m] = 0;
  reinterpret_cast<int*>(w.minmax)[2 * m] = 0x7fffffff;
  reinterpret_cast<int*>(w.minmax)[2 * m + 1] = (int)0x80000000;
}

}  // namespace

extern "C" {

int sap3d_resize_bilinear(const float* src, int32_t n, int32_t h, int32_t w, float* dst, int32_t H, int32_t W, void* stream) {
  if (require_device()) return 1;
  if (!src || !dst || n < 1 || h < 1 || w < 1 || H < 1 || W < 1) return set_error("resize_bilinear: bad argument");
  const long long total = (long long)n * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  resize_bilinear_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, n, h, w, dst, H, W);
  return check_launch("resize_bilinear");
}

size_t sap3d_saliency_auc_workspace(int32_t n_maps, int32_t n_rep) { return ws_layout(n_maps, n_rep, nullptr, nullptr); }

/* out[n][2] = (AUC_Judd, AUC_Borji); NaN for maps without fixations (or with more than 16384 of them) */
int sap3d_saliency_auc(const float* sal, const float* fix, int32_t n_maps, int64_t elems, int32_t jitter, int32_t n_rep, double step,
                       uint64_t seed, double* out, void* workspace, void* stream) {
  if (require_device()) return 1;
  if (!sal || !fix || !out || !workspace) return set_error("saliency_auc: NULL pointer");
  if (n_rep < 1 || n_rep > 1024 || !(step > 0.0) || elems < 1 || elems > 0x7fffffffll) return set_error("saliency_auc: bad n_rep / step / size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AucWs w;
  ws_layout(n_maps, n_rep, &w, reinterpret_cast<char*>(workspace));
  auc_init_kernel<<<(n_maps + 127) / 128, 128, 0, st>>>(w, n_maps);
  long long cb = (elems + 255) / 256;
  if (cb > 148 * 4) cb = 148 * 4;
  auc_collect_kernel<<<dim3((unsigned)cb, n_maps), 256, 0, st>>>(sal, fix, elems, w);
  if (check_launch("saliency_auc collect")) return 1;
  // the fixation counts live on the device: launch the per-fixation grid at the capacity bound of the map size
  const int kmax = (int)(elems < AUC_CAP ? elems : AUC_CAP);
  auc_judd_count_kernel<<<dim3(kmax, n_maps), 256, 0, st>>>(sal, elems, jitter, seed, w);
  auc_judd_area_kernel<<<(n_maps + 127) / 128, 128, 0, st>>>(elems, w, out, n_maps);
  auc_borji_kernel<<<dim3(n_rep, n_maps), 256, 0, st>>>(sal, elems, n_rep, step, seed, w);
  auc_borji_mean_kernel<<<(n_maps + 127) / 128, 128, 0, st>>>(n_rep, w, out, n_maps);
  return check_launch("saliency_auc");
}

}  // extern "C"

// ---- input preprocessing (dataflow.py:194-209 `mapf`, gen_pred.py:113-118): BGR uint8 frame -> RGB, minus the per-channel
// mean, cv2.resize(frame, (112, 112)) on the float image (INTER_LINEAR), divided by 255 -> NDHWC network input --------
namespace {
template <typename T>
__global__ void __launch_bounds__(256) preprocess_frames_kernel(const unsigned char* __restrict__ bgr, int n, int h, int w, float m0, float m1,
                                                                 float m2, T* __restrict__ dst, int H, int W) {
  const double sx = (double)w / W, sy = (double)h / H;
  const long long total = (long long)n * H * W;
  const float mean[3] = {m0, m1, m2};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % W);
    long long r = i / W;
    const int dy = (int)(r % H);
    const int f = (int)(r / H);
    const double cx = (dx + 0.5) * sx - 0.5, cy = (dy + 0.5) * sy - 0.5;   // fp64 coordinate, fp32 fraction (what cv2 4.x does)
    int x0 = (int)floor(cx), y0 = (int)floor(cy);
    float fx = (float)(cx - x0), fy = (float)(cy - y0);
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= w - 1) { x0 = w - 1; fx = 0.f; }
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= h - 1) { y0 = h - 1; fy = 0.f; }
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const unsigned char* s = bgr + (long long)f * h * w * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // output channel c (RGB) = input channel 2 - c (BGR)
      const int ci = 2 - c;
      const float p00 = (float)s[(y0 * w + x0) * 3 + ci] - mean[c], p01 = (float)s[(y0 * w + x1) * 3 + ci] - mean[c];
      const float p10 = (float)s[(y1 * w + x0) * 3 + ci] - mean[c], p11 = (float)s[(y1 * w + x1) * 3 + ci] - mean[c];
      const float r0 = __fadd_rn(__fmul_rn(p00, 1.f - fx), __fmul_rn(p01, fx));
      const float r1 = __fadd_rn(__fmul_rn(p10, 1.f - fx), __fmul_rn(p11, fx));
      dst[i * 3 + c] = from_f32<T>(__fadd_rn(__fmul_rn(r0, 1.f - fy), __fmul_rn(r1, fy)) / 255.f);
    }
  }
}
}  // namespace

extern "C" int sap3d_preprocess_frames(const uint8_t* bgr, int32_t n, int32_t h, int32_t w, const float* mean_rgb_host, int32_t out_dtype,
                                       void* dst, int32_t H, int32_t W, void* stream) {
  if (require_device()) return 1;
  if (!bgr || !dst || !mean_rgb_host || n < 1 || h < 1 || w < 1 || H < 1 || W < 1) return set_error("preprocess_frames: bad argument");
  const long long total = (long long)n * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == SAP3D_BF16)
    preprocess_frames_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>(bgr, n, h, w, mean_rgb_host[0], mean_rgb_host[1], mean_rgb_host[2], reinterpret_cast<bf16*>(dst), H, W);
  else
    preprocess_frames_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(bgr, n, h, w, mean_rgb_host[0], mean_rgb_host[1], mean_rgb_host[2], reinterpret_cast<float*>(dst), H, W);
  return check_launch("preprocess_frames");
}

// NaN-filtered column sums / counts of a small [n][m] table of per-clip metric values (test.py:177-181 drops NaN scores
// before averaging); the (sum, count) pairs are what a sharded evaluation all-reduces.
namespace {
__global__ void nan_sum_count_kernel(const double* __restrict__ v, int n, int m, double* __restrict__ sums, double* __restrict__ counts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s = 0.0, c = 0.0;
  for (int i = 0; i < n; ++i) {
    const double x = v[(long long)i * m + j];
    if (x == x) { s += x; c += 1.0; }
  }
  sums[j] = s;
  counts[j] = c;
}
}  // namespace
extern "C" int sap3d_nan_sum_count(const double* values, int32_t n, int32_t m, double* sums, double* counts, void* stream) {
  if (require_device()) return 1;
  if (!values || !sums || !counts || n < 0 || m < 1) return set_error("nan_sum_count: bad argument");
  nan_sum_count_kernel<<<(m + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(values, n, m, sums, counts);
  return check_launch("nan_sum_count");
}
