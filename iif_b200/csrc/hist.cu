// (d) Label histograms and the IIF weight vector.
//   - hist_labels: per-class counts with int32 atomics in shared memory, flushed once per CTA with
//     int64 global atomics (replaces the O(N*C) numpy loops of cls/imbalanced_dataset.py:112,127).
//   - hist_images_dedup: img_freq / instance_freq of seg/lvis_files/idf_1204.csv via an exactly-once
//     (class, image) bitmap (atomicOr returns the previous word).
//   - weights_from_counts: the 7 closed forms of cls/custom.py:16-23 in float64, rounded once to fp32.
// Integer results are bit-exact and order-independent (integer atomics commute).
#include "common.cuh"
#include "ndtri.h"

namespace iif {

constexpr int kSmemBins = 12 * 1024;  // 48 KB of int32 bins

__global__ void __launch_bounds__(256) hist_smem_kernel(const int64_t* __restrict__ labels, int64_t n,
                                                        unsigned long long* __restrict__ counts, int C) {
  extern __shared__ int s_bins[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_bins[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n2 = n / 2;  // 128-bit loads: two labels per access (base pointer is 16-byte aligned)
  const longlong2* l2 = reinterpret_cast<const longlong2*>(labels);
  for (int64_t i = tid; i < n2; i += stride) {
    const longlong2 v = __ldg(l2 + i);
    if (v.x >= 0 && v.x < C) atomicAdd(&s_bins[(int)v.x], 1);
    if (v.y >= 0 && v.y < C) atomicAdd(&s_bins[(int)v.y], 1);
  }
  if (tid == 0 && (n & 1)) {
    const int64_t v = labels[n - 1];
    if (v >= 0 && v < C) atomicAdd(&s_bins[(int)v], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const int c = s_bins[i];
    if (c) atomicAdd(counts + i, (unsigned long long)c);
  }
}

// C too large for shared bins (or unaligned labels): straight global atomics
__global__ void __launch_bounds__(256) hist_global_kernel(const int64_t* __restrict__ labels, int64_t n,
                                                          unsigned long long* __restrict__ counts, int64_t C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t v = __ldg(labels + i);
    if (v >= 0 && v < C) atomicAdd(counts + v, 1ull);
  }
}

__global__ void __launch_bounds__(256) hist_dedup_kernel(const int64_t* __restrict__ img, const int64_t* __restrict__ cat,
                                                         int64_t n, int64_t n_img, int64_t C, int64_t words_per_class,
                                                         unsigned long long* __restrict__ img_freq,
                                                         unsigned long long* __restrict__ inst_freq,
                                                         unsigned int* __restrict__ bitmap) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t c = __ldg(cat + i), im = __ldg(img + i);
    if (c < 0 || c >= C || im < 0 || im >= n_img) continue;
    atomicAdd(inst_freq + c, 1ull);
    const unsigned int bit = 1u << (im & 31);
    const unsigned int old = atomicOr(bitmap + c * words_per_class + (im >> 5), bit);
    if (!(old & bit)) atomicAdd(img_freq + c, 1ull);  // exactly one setter sees the bit clear
  }
}

__device__ __forceinline__ double iif_variant(double f, double n, int variant) {
  switch (variant) {
    case IIF_VARIANT_RAW: return log(n / f);
    case IIF_VARIANT_SMOOTH: return log((n + 1.0) / (f + 1.0)) + 1.0;
    case IIF_VARIANT_REL: return log((n - f) / f);
    case IIF_VARIANT_NORMIT: return -iif_ndtri(f / n);
    case IIF_VARIANT_GOMBIT: return -log(-log(1.0 - (f / n)));
    case IIF_VARIANT_BASE2: return log2(n / f);
    default: return log10(n / f);
  }
}

// one CTA: N = sum(counts) (exact int64), weights in float64, one rounding to fp32, optional p-norm
__global__ void __launch_bounds__(1024) weights_kernel(const int64_t* __restrict__ counts, int C, int64_t total,
                                                       int variant, double norm_p, float* __restrict__ out32,
                                                       double* __restrict__ out64) {
  __shared__ long long s_n[32];
  __shared__ double s_d[32];
  __shared__ double s_total;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  long long part = 0;
  if (total <= 0) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) part += counts[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_n[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < nw; ++w) t += s_n[w];
      s_total = (double)t;
    }
  } else if (threadIdx.x == 0) {
    s_total = (double)total;
  }
  __syncthreads();
  const double n = s_total;
  double pacc = 0.0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const double v = iif_variant((double)counts[i], n, variant);
    if (out64) out64[i] = v;
    const float v32 = (float)v;  // torch.tensor([v], dtype=torch.float): one rounding (cls/custom.py:24)
    out32[i] = v32;
    if (norm_p > 0.0) pacc += pow(fabs((double)v32), norm_p);
  }
  if (norm_p > 0.0) {
    pacc = warp_sum_d(pacc);
    if (lane == 0) s_d[warp] = pacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < nw; ++w) t += s_d[w];
      s_total = pow(t, 1.0 / norm_p);
    }
    __syncthreads();
    const float nrm = (float)s_total;  // v / torch.norm(v, p) in fp32 (cls/custom.py:25-26)
    for (int i = threadIdx.x; i < C; i += blockDim.x) out32[i] = out32[i] / nrm;
  }
}

}  // namespace iif

using namespace iif;

extern "C" int iif_hist_labels_i64(const int64_t* labels, int64_t n, int64_t* counts, int64_t C, void* stream) {
  if (n < 0 || C <= 0 || !counts || (n > 0 && !labels)) return IIF_EINVAL;
  if (n == 0) return IIF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  auto* out = reinterpret_cast<unsigned long long*>(counts);
  int64_t blocks = (n / 2 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;  // persistent-ish grid: 4 CTAs per SM
  if (C <= kSmemBins && aligned16(labels)) {
    hist_smem_kernel<<<(unsigned)blocks, 256, (size_t)C * sizeof(int), st>>>(labels, n, out, (int)C);
  } else {
    hist_global_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels, n, out, C);
  }
  return launch_status();
}

extern "C" size_t iif_hist_images_dedup_ws_bytes(int64_t num_images, int64_t num_classes) {
  if (num_images <= 0 || num_classes <= 0) return 0;
  return (size_t)num_classes * (size_t)((num_images + 31) / 32) * sizeof(uint32_t);
}

extern "C" int iif_hist_images_dedup_i64(const int64_t* image_ids, const int64_t* categories, int64_t n,
                                         int64_t num_images, int64_t C, int64_t* img_freq, int64_t* instance_freq,
                                         uint32_t* bitmap_ws, void* stream) {
  if (n < 0 || C <= 0 || num_images <= 0 || !img_freq || !instance_freq || !bitmap_ws) return IIF_EINVAL;
  if (n > 0 && (!image_ids || !categories)) return IIF_EINVAL;
  if (n == 0) return IIF_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  hist_dedup_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      image_ids, categories, n, num_images, C, (num_images + 31) / 32,
      reinterpret_cast<unsigned long long*>(img_freq), reinterpret_cast<unsigned long long*>(instance_freq), bitmap_ws);
  return launch_status();
}

extern "C" int iif_weights_from_counts(const int64_t* counts, int64_t C, int64_t total, int variant, double norm_p,
                                       float* out_f32, double* out_f64, void* stream) {
  if (!counts || !out_f32 || C <= 0 || C > (1 << 30)) return IIF_EINVAL;
  if (variant < IIF_VARIANT_RAW || variant > IIF_VARIANT_BASE10 || norm_p < 0.0) return IIF_EINVAL;
  weights_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, (int)C, total, variant, norm_p, out_f32, out_f64);
  return launch_status();
}
