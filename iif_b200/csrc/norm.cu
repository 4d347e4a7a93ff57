// Normalised classifiers: every variant the reference defines is  z = r_i (x_i . w_c) c_c + b_c
// with r / c functions of a row's 2-norm:
//   mmdet NormedLinear / IIFNormedLinear (seg/mmdet/models/utils/normed_predictor.py:11-76):
//       x_ = T x / (|x|^p + eps),   w_ = w' / (|w'|^p + eps),  w' = iif_c w   (IIF variant)
//   CosNorm_Classifier (cls/resnet_cifar.py:50-78):  ex = s x / (1 + |x|),  ew = w / |w|
// The reference normalises the OPERANDS and then calls F.linear; so does this path (the GEMMs stay the
// head's tensor-core kernels), with three HBM-bound row kernels around them:
//   row_scale_from_norm : per row n = |pre_i x_i|, multiplier a_i = pre_i r(n) and the backward coefficient
//                         c_i = pre_i^3 r'(n) / n
//   row_dot             : d_i = u_i . v_i
//   rows_axpby          : out_i = a_i u_i + (b_i b2_i) v_i      (y = a x forward;  dx = a g + c (x.g) x backward)
// One warp per row, 128-bit loads, fp32 accumulation in a fixed order (deterministic).
#include "common.cuh"

namespace iif {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool VEC>
__global__ void __launch_bounds__(256) row_scale_from_norm_kernel(const float* __restrict__ x, int64_t ld, int64_t rows,
                                                                   int cols, const float* __restrict__ pre, int mode,
                                                                   float T, float p, float eps, float* __restrict__ a_out,
                                                                   float* __restrict__ c_out, float* __restrict__ n_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ld;
  float ss = 0.f;
  if constexpr (VEC) {
    for (int c = lane * 4; c < cols; c += 128) {
      const float4 v = ld4(xr + c);
      ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  } else {
    for (int c = lane; c < cols; c += 32) { const float v = __ldg(xr + c); ss += v * v; }
  }
  ss = warp_sum(ss);
  if (lane != 0) return;
  const float pr = pre ? __ldg(pre + row) : 1.f;
  const float n = fabsf(pr) * sqrtf(ss);                 // |pre x|
  float r, dr;                                            // r(n), r'(n)
  if (mode == IIF_NORM_NORMED) {                          // T / (n^p + eps)
    const float np_ = p == 1.f ? n : powf(n, p);
    const float den = np_ + eps;
    r = T / den;
    dr = n > 0.f ? -T * p * (p == 1.f ? 1.f : powf(n, p - 1.f)) / (den * den) : 0.f;
  } else if (mode == IIF_NORM_COS) {                      // T / (1 + n)
    r = T / (1.f + n);
    dr = -T / ((1.f + n) * (1.f + n));
  } else {                                                // T / max(n, eps)
    r = T / fmaxf(n, eps);
    dr = n > eps ? -T / (n * n) : 0.f;
  }
  a_out[row] = pr * r;
  if (c_out) c_out[row] = n > 0.f ? pr * pr * pr * dr / n : 0.f;
  if (n_out) n_out[row] = n;
}

template <bool VEC>
__global__ void __launch_bounds__(256) row_dot_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ v,
                                                       int64_t ldv, int64_t rows, int cols, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* ur = u + row * ldu;
  const float* vr = v + row * ldv;
  float acc = 0.f;
  if constexpr (VEC) {
    for (int c = lane * 4; c < cols; c += 128) {
      const float4 a = ld4(ur + c), b = ld4(vr + c);
      acc += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
  } else {
    for (int c = lane; c < cols; c += 32) acc += __ldg(ur + c) * __ldg(vr + c);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

template <bool VEC>
__global__ void __launch_bounds__(256) rows_axpby_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ a,
                                                          const float* __restrict__ v, int64_t ldv,
                                                          const float* __restrict__ b, const float* __restrict__ b2,
                                                          int64_t rows, int cols, float* __restrict__ out, int64_t ldo) {
  const int cpr = VEC ? cols / 4 : cols;
  const int64_t total = rows * (int64_t)cpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c = (int)(i - r * cpr) * (VEC ? 4 : 1);
    const float av = a ? __ldg(a + r) : 1.f;
    const float bv = v ? (b ? __ldg(b + r) : 1.f) * (b2 ? __ldg(b2 + r) : 1.f) : 0.f;
    if constexpr (VEC) {
      float4 x = ld4(u + r * ldu + c);
      x.x *= av; x.y *= av; x.z *= av; x.w *= av;
      if (v) {
        const float4 y = ld4(v + r * ldv + c);
        x.x += bv * y.x; x.y += bv * y.y; x.z += bv * y.z; x.w += bv * y.w;
      }
      *reinterpret_cast<float4*>(out + r * ldo + c) = x;
    } else {
      float x = __ldg(u + r * ldu + c) * av;
      if (v) x += bv * __ldg(v + r * ldv + c);
      out[r * ldo + c] = x;
    }
  }
}

}  // namespace iif

using namespace iif;

extern "C" int iif_row_scale_from_norm(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* pre, int mode,
                                       float temperature, float power, float eps, float* a_out, float* c_out,
                                       float* norm_out, void* stream) {
  if (rows < 0 || cols <= 0 || ldx < cols || !a_out || (rows > 0 && !x) || cols > (1 << 30)) return IIF_EINVAL;
  if (mode != IIF_NORM_NORMED && mode != IIF_NORM_COS && mode != IIF_NORM_UNIT) return IIF_EINVAL;
  if (rows == 0) return IIF_OK;
  const bool vec = (cols % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) row_scale_from_norm_kernel<true><<<grid, 256, 0, st>>>(x, ldx, rows, (int)cols, pre, mode, temperature, power, eps, a_out, c_out, norm_out);
  else row_scale_from_norm_kernel<false><<<grid, 256, 0, st>>>(x, ldx, rows, (int)cols, pre, mode, temperature, power, eps, a_out, c_out, norm_out);
  return launch_status();
}

extern "C" int iif_row_dot(const float* u, int64_t ldu, const float* v, int64_t ldv, int64_t rows, int64_t cols, float* out,
                           void* stream) {
  if (rows < 0 || cols <= 0 || ldu < cols || ldv < cols || !out || (rows > 0 && (!u || !v)) || cols > (1 << 30)) return IIF_EINVAL;
  if (rows == 0) return IIF_OK;
  const bool vec = (cols % 4 == 0) && (ldu % 4 == 0) && (ldv % 4 == 0) && aligned16(u) && aligned16(v);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (vec) row_dot_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(u, ldu, v, ldv, rows, (int)cols, out);
  else row_dot_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(u, ldu, v, ldv, rows, (int)cols, out);
  return launch_status();
}

extern "C" int iif_rows_axpby(const float* u, int64_t ldu, const float* a, const float* v, int64_t ldv, const float* b,
                              const float* b2, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream) {
  if (rows < 0 || cols < 0 || ldu < cols || ldo < cols || (v && ldv < cols) || cols > (1 << 30)) return IIF_EINVAL;
  if (rows == 0 || cols == 0) return IIF_OK;
  if (!u || !out) return IIF_EINVAL;
  const bool vec = (cols % 4 == 0) && (ldu % 4 == 0) && (ldo % 4 == 0) && aligned16(u) && aligned16(out) &&
                   (!v || (ldv % 4 == 0 && aligned16(v)));
  const int64_t items = rows * (vec ? cols / 4 : cols);
  const unsigned grid = (unsigned)((items + 255) / 256 < 8 * kNumSMs ? (items + 255) / 256 : 8 * kNumSMs);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) rows_axpby_kernel<true><<<grid, 256, 0, st>>>(u, ldu, a, v, ldv, b, b2, rows, (int)cols, out, ldo);
  else rows_axpby_kernel<false><<<grid, 256, 0, st>>>(u, ldu, a, v, ldv, b, b2, rows, (int)cols, out, ldo);
  return launch_status();
}
