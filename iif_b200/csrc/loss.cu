// Fused row kernels of the IIF head: softmax-CE fwd+bwd with the per-class IIF scale, the
// softmax / scaled activation, sigmoid-BCE fwd+bwd, row scaling / bf16 cast and the bias-gradient
// column sum.  All HBM-bound: each logit is read once (128-bit, streaming), every intermediate
// (scaled logits, probabilities, one-hot targets) lives in registers only.
//
// Reference semantics restated here (never copied): cls/custom.py:28-39,61-73;
// seg/mmdet/models/losses/iif_loss.py:65-78,187-200; cross_entropy_loss.py:53-111;
// losses/utils.py:28-55; losses/accuracy.py:41-50.
#include <math_constants.h>

#include "common.cuh"

namespace iif {

struct RowArgs {
  const float* z; int64_t ldz;
  const float* iif;
  const int64_t* label;
  const float* cw;
  const float* sw;
  int64_t ignore_index;
  float scale;
  int64_t B; int C;
  float* loss_i; float* loss_sum;
  float* dz32; int64_t lddz32;
  uint16_t* dz16; int64_t lddz16;
  float* lse; int32_t* argmax; int32_t* rank; int32_t* acc_counts; int32_t* ticket;
  float* out; int64_t ldo; int softmax; int on_scaled;
};

// ---- reductions over the TPR threads that own one row (TPR is a multiple of 32) -----------------
template <int TPR>
__device__ __forceinline__ void row_reduce_pass1(float& m, float& bv, int& bi, int& cnt, float* s_m,
                                                 float* s_bv, int* s_bi, int* s_cnt) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if constexpr (TPR > 32) {
    constexpr int WPR = TPR / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_m[warp] = m; s_bv[warp] = bv; s_bi[warp] = bi; s_cnt[warp] = cnt; }
    __syncthreads();
    const int w0 = (warp / WPR) * WPR;
    m = s_m[w0]; bv = s_bv[w0]; bi = s_bi[w0]; cnt = s_cnt[w0];
#pragma unroll 1
    for (int w = 1; w < WPR; ++w) {
      m = fmaxf(m, s_m[w0 + w]);
      float ov = s_bv[w0 + w]; int oi = s_bi[w0 + w];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      cnt += s_cnt[w0 + w];
    }
  }
}

template <int TPR>
__device__ __forceinline__ float row_reduce_sum(float v, float* s_v) {
  v = warp_sum(v);
  if constexpr (TPR > 32) {
    constexpr int WPR = TPR / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_v[warp] = v;
    __syncthreads();
    const int w0 = (warp / WPR) * WPR;
    v = s_v[w0];
#pragma unroll 1
    for (int w = 1; w < WPR; ++w) v += s_v[w0 + w];
  }
  return v;
}

// Deterministic tail executed by the last CTA to finish: fixed-order sum of loss_i and the top-k
// hit counts.  `ticket` is self-resetting.
__device__ __forceinline__ void last_block_reduce(const float* loss_i, const int32_t* rank, int64_t B,
                                                  float* loss_sum, int32_t* acc_counts, int32_t* ticket) {
  __shared__ int s_last;
  __shared__ double s_acc[32];
  __shared__ int s_c1[32], s_c5[32];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  int c1 = 0, c5 = 0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    if (loss_sum) acc += (double)__ldcg(loss_i + i);
    if (acc_counts) { int r = __ldcg(rank + i); c1 += (r < 1); c5 += (r < 5); }
  }
  acc = warp_sum_d(acc); c1 = warp_sum_i(c1); c5 = warp_sum_i(c5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { s_acc[warp] = acc; s_c1[warp] = c1; s_c5[warp] = c5; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0; int k1 = 0, k5 = 0;
    for (int w = 0; w < nw; ++w) { a += s_acc[w]; k1 += s_c1[w]; k5 += s_c5[w]; }
    if (loss_sum) *loss_sum = (float)a;
    if (acc_counts) { acc_counts[0] = k1; acc_counts[1] = k5; }
    *ticket = 0;
  }
}

// MODE 0: softmax-CE forward + backward.  MODE 1: activation (softmax(z*iif) or z*iif).
template <int TPR, int NE, bool VEC, int MODE>
__global__ void __launch_bounds__(TPR > 256 ? TPR : 256)
row_softmax_kernel(const RowArgs a) {
  __shared__ float s_m[32], s_bv[32], s_sum[32];
  __shared__ int s_bi[32], s_cnt[32];
  const int t = threadIdx.x % TPR;
  const int rpb = blockDim.x / TPR;
  const int64_t row = (int64_t)blockIdx.x * rpb + threadIdx.x / TPR;
  const bool active = row < a.B;
  const int C = a.C;
  const float* zr = a.z + (active ? row : 0) * a.ldz;

  int64_t y = -1;
  if (active && a.label) y = a.label[row];
  const bool y_in = active && y >= 0 && y < C;
  const bool y_ok = y_in && y != a.ignore_index;
  const float zy = y_in ? zr[y] : 0.f;
  const float sy = (y_in && a.iif) ? a.iif[y] : 1.f;
  const float ref = a.on_scaled ? zy * sy : zy;
  const int yi = y_in ? (int)y : -1;

  float v[NE];
  float m = -CUDART_INF_F, bv = -CUDART_INF_F;
  int bi = 0x7fffffff, cnt = 0;

  auto visit = [&](int e, int col, float z, float s) {
    const float sc = z * s;
    const float cmp = a.on_scaled ? sc : z;
    if (cmp > bv) { bv = cmp; bi = col; }
    cnt += (cmp > ref) || (cmp == ref && col < yi);
    v[e] = sc;
    m = fmaxf(m, sc);
  };

  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < NE / 4; ++q) {
      const int col = (q * TPR + t) * 4;
      if (active && col < C) {
        const float4 z4 = ldg_stream4(zr + col);
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.iif) s4 = __ldg(reinterpret_cast<const float4*>(a.iif + col));
        visit(4 * q + 0, col + 0, z4.x, s4.x);
        visit(4 * q + 1, col + 1, z4.y, s4.y);
        visit(4 * q + 2, col + 2, z4.z, s4.z);
        visit(4 * q + 3, col + 3, z4.w, s4.w);
      } else {
        v[4 * q] = v[4 * q + 1] = v[4 * q + 2] = v[4 * q + 3] = -CUDART_INF_F;
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int col = e * TPR + t;
      if (active && col < C) visit(e, col, __ldg(zr + col), a.iif ? __ldg(a.iif + col) : 1.f);
      else v[e] = -CUDART_INF_F;
    }
  }
  if (yi < 0) cnt = C;  // label outside [0,C): never inside any top-k
  row_reduce_pass1<TPR>(m, bv, bi, cnt, s_m, s_bv, s_bi, s_cnt);
  if (yi < 0) cnt = C;

  if (MODE == 1 && !a.softmax) {
    // out = z * iif  (cls/custom.py:38)
    float* o = a.out + (active ? row : 0) * a.ldo;
    if constexpr (VEC) {
#pragma unroll
      for (int q = 0; q < NE / 4; ++q) {
        const int col = (q * TPR + t) * 4;
        if (active && col < C) stg_stream4(o + col, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
    } else {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int col = e * TPR + t;
        if (active && col < C) o[col] = v[e];
      }
    }
  } else {
    const float mm = (m == -CUDART_INF_F) ? 0.f : m;
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) { v[e] = expf(v[e] - mm); sum += v[e]; }
    sum = row_reduce_sum<TPR>(sum, s_sum);
    const float inv = 1.f / sum;
    if constexpr (MODE == 1) {
      float* o = a.out + (active ? row : 0) * a.ldo;
      if constexpr (VEC) {
#pragma unroll
        for (int q = 0; q < NE / 4; ++q) {
          const int col = (q * TPR + t) * 4;
          if (active && col < C)
            stg_stream4(o + col, make_float4(v[4 * q] * inv, v[4 * q + 1] * inv, v[4 * q + 2] * inv, v[4 * q + 3] * inv));
        }
      } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int col = e * TPR + t;
          if (active && col < C) o[col] = v[e] * inv;
        }
      }
    } else {
      const float lse = mm + logf(sum);
      float g = 0.f;
      if (y_ok) {
        g = a.scale;
        if (a.cw) g *= a.cw[y];
        if (a.sw) g *= a.sw[row];
      }
      if (active && t == 0) {
        if (a.loss_i) a.loss_i[row] = y_ok ? g * (lse - zy * sy) : 0.f;
        if (a.lse) a.lse[row] = lse;
      }
      if (a.dz32 || a.dz16) {
        float* d32 = a.dz32 ? a.dz32 + (active ? row : 0) * a.lddz32 : nullptr;
        uint16_t* d16 = a.dz16 ? a.dz16 + (active ? row : 0) * a.lddz16 : nullptr;
        if constexpr (VEC) {
#pragma unroll
          for (int q = 0; q < NE / 4; ++q) {
            const int col = (q * TPR + t) * 4;
            if (active && col < C) {
              float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
              if (a.iif) s4 = __ldg(reinterpret_cast<const float4*>(a.iif + col));
              float4 d;
              d.x = s4.x * g * (v[4 * q + 0] * inv - (col + 0 == yi ? 1.f : 0.f));
              d.y = s4.y * g * (v[4 * q + 1] * inv - (col + 1 == yi ? 1.f : 0.f));
              d.z = s4.z * g * (v[4 * q + 2] * inv - (col + 2 == yi ? 1.f : 0.f));
              d.w = s4.w * g * (v[4 * q + 3] * inv - (col + 3 == yi ? 1.f : 0.f));
              if (!y_ok) d = make_float4(0.f, 0.f, 0.f, 0.f);  // ignored row: exact zeros even for inf weights
              if (d32) stg_stream4(d32 + col, d);
              if (d16) stg_stream2(d16 + col, pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int col = e * TPR + t;
            if (active && col < C) {
              const float s = a.iif ? __ldg(a.iif + col) : 1.f;
              float d = s * g * (v[e] * inv - (col == yi ? 1.f : 0.f));
              if (!y_ok) d = 0.f;
              if (d32) d32[col] = d;
              if (d16) d16[col] = bf16_bits(d);
            }
          }
        }
      }
    }
  }
  if (active && t == 0) {
    if (a.argmax) a.argmax[row] = bi;
    if (a.rank) a.rank[row] = cnt;
  }
  if constexpr (MODE == 0) {
    if (a.loss_sum || a.acc_counts) last_block_reduce(a.loss_i, a.rank, a.B, a.loss_sum, a.acc_counts, a.ticket);
  }
}

template <int TPR, int NE, int MODE>
static int launch_row(const RowArgs& a, bool vec, cudaStream_t st) {
  // rows per CTA: as many as fit 256 threads, but never fewer CTAs than ~2 per SM
  int rpb = TPR >= 256 ? 1 : 256 / TPR;
  while (rpb > 1 && (a.B + rpb - 1) / rpb < 2 * kNumSMs) rpb >>= 1;
  const unsigned grid = (unsigned)((a.B + rpb - 1) / rpb);
  if (vec) row_softmax_kernel<TPR, NE, true, MODE><<<grid, TPR * rpb, 0, st>>>(a);
  else row_softmax_kernel<TPR, NE, false, MODE><<<grid, TPR * rpb, 0, st>>>(a);
  return launch_status();
}

template <int MODE>
static int dispatch_row(const RowArgs& a, bool vec, cudaStream_t st) {
  const int C = a.C;
  if (C <= 128) return launch_row<32, 4, MODE>(a, vec, st);
  if (C <= 256) return launch_row<32, 8, MODE>(a, vec, st);
  if (C <= 512) return launch_row<32, 16, MODE>(a, vec, st);
  if (C <= 1024) return launch_row<32, 32, MODE>(a, vec, st);
  if (C <= 2048) return launch_row<128, 16, MODE>(a, vec, st);
  if (C <= 4096) return launch_row<128, 32, MODE>(a, vec, st);
  if (C <= 8192) return launch_row<256, 32, MODE>(a, vec, st);
  if (C <= 16384) return launch_row<512, 32, MODE>(a, vec, st);
  if (C <= 32768) return launch_row<1024, 32, MODE>(a, vec, st);
  return IIF_EUNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// sigmoid BCE forward + backward
// ------------------------------------------------------------------------------------------------
struct BceArgs {
  const float* z; int64_t ldz; const int64_t* label;
  const float* pw; const float* colw; const float* sw;
  int64_t ignore_index; float scale; int64_t B; int C;
  float* loss_elem; int64_t ldl; float* loss_i; float* loss_sum;
  float* dz32; int64_t lddz32; uint16_t* dz16; int64_t lddz16; int32_t* ticket;
};

__device__ __forceinline__ void bce_elem(float z, bool t, float pw, float wgt, float& loss, float& d) {
  // F.binary_cross_entropy_with_logits: (1-t) z + lw * (log1p(exp(-|z|)) + max(-z, 0)), lw = 1 + (pw-1) t
  const float lw = t ? pw : 1.f;
  const float e = expf(-fabsf(z));
  const float sp = log1pf(e) + fmaxf(-z, 0.f);
  loss = wgt * ((t ? 0.f : z) + lw * sp);
  const float sig = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  d = wgt * ((t ? 0.f : 1.f) - lw * (1.f - sig));
}

template <bool VEC>
__global__ void __launch_bounds__(256) bce_kernel(const BceArgs a) {
  constexpr int TPR = 128;
  __shared__ float s_sum[32];
  const int t = threadIdx.x % TPR;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / TPR) + threadIdx.x / TPR;
  const bool active = row < a.B;
  const int C = a.C;
  float acc = 0.f;
  if (active) {
    const int64_t y = a.label[row];
    const bool valid = y >= 0 && y != a.ignore_index;
    const int yi = (valid && y < C) ? (int)y : -1;
    const float wrow = valid ? a.scale * (a.sw ? a.sw[row] : 1.f) : 0.f;
    const float* zr = a.z + row * a.ldz;
    float* le = a.loss_elem ? a.loss_elem + row * a.ldl : nullptr;
    float* d32 = a.dz32 ? a.dz32 + row * a.lddz32 : nullptr;
    uint16_t* d16 = a.dz16 ? a.dz16 + row * a.lddz16 : nullptr;
    if constexpr (VEC) {
      for (int col = t * 4; col < C; col += TPR * 4) {
        const float4 z4 = ldg_stream4(zr + col);
        float4 p4 = make_float4(1.f, 1.f, 1.f, 1.f), c4 = p4, l4, d4;
        if (a.pw) p4 = __ldg(reinterpret_cast<const float4*>(a.pw + col));
        if (a.colw) c4 = __ldg(reinterpret_cast<const float4*>(a.colw + col));
        bce_elem(z4.x, col + 0 == yi, p4.x, wrow * c4.x, l4.x, d4.x);
        bce_elem(z4.y, col + 1 == yi, p4.y, wrow * c4.y, l4.y, d4.y);
        bce_elem(z4.z, col + 2 == yi, p4.z, wrow * c4.z, l4.z, d4.z);
        bce_elem(z4.w, col + 3 == yi, p4.w, wrow * c4.w, l4.w, d4.w);
        if (!valid) { l4 = make_float4(0.f, 0.f, 0.f, 0.f); d4 = l4; }
        acc += (l4.x + l4.y) + (l4.z + l4.w);
        if (le) stg_stream4(le + col, l4);
        if (d32) stg_stream4(d32 + col, d4);
        if (d16) stg_stream2(d16 + col, pack_bf16x2(d4.x, d4.y), pack_bf16x2(d4.z, d4.w));
      }
    } else {
      for (int col = t; col < C; col += TPR) {
        float l, d;
        bce_elem(__ldg(zr + col), col == yi, a.pw ? __ldg(a.pw + col) : 1.f,
                 wrow * (a.colw ? __ldg(a.colw + col) : 1.f), l, d);
        if (!valid) { l = 0.f; d = 0.f; }
        acc += l;
        if (le) le[col] = l;
        if (d32) d32[col] = d;
        if (d16) d16[col] = bf16_bits(d);
      }
    }
  }
  acc = row_reduce_sum<TPR>(acc, s_sum);
  if (active && t == 0 && a.loss_i) a.loss_i[row] = acc;
  if (a.loss_sum) last_block_reduce(a.loss_i, nullptr, a.B, a.loss_sum, nullptr, a.ticket);
}

// ------------------------------------------------------------------------------------------------
// out = in * g (per row / scalar), optional bf16 cast
// ------------------------------------------------------------------------------------------------
template <bool VEC, bool BF16>
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ in, int64_t ldi,
                                                         const float* __restrict__ g, int64_t gs, int64_t rows,
                                                         int cols, void* __restrict__ out, int64_t ldo) {
  const int cpr = VEC ? (cols + 3) / 4 : cols;  // work items per row
  const int64_t total = rows * (int64_t)cpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c = (int)(i - r * cpr) * (VEC ? 4 : 1);
    const float gv = g ? __ldg(g + r * gs) : 1.f;
    if constexpr (VEC) {
      float4 v = ldg_stream4(in + r * ldi + c);
      v.x *= gv; v.y *= gv; v.z *= gv; v.w *= gv;
      if constexpr (BF16) stg_stream2(reinterpret_cast<uint16_t*>(out) + r * ldo + c, pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      else stg_stream4(reinterpret_cast<float*>(out) + r * ldo + c, v);
    } else {
      const float v = __ldg(in + r * ldi + c) * gv;
      if constexpr (BF16) reinterpret_cast<uint16_t*>(out)[r * ldo + c] = bf16_bits(v);
      else reinterpret_cast<float*>(out)[r * ldo + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// db[c] = alpha * sum_i dz[i,c]: 32 columns per CTA, 8 row lanes, fixed-order tree
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ dz, int64_t ld, const float* __restrict__ alpha,
                                                     int64_t rows, int cols, float* __restrict__ db) {
  __shared__ float s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < cols) {
    for (int64_t r = ty; r < rows; r += 8) {
      if constexpr (BF16) acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dz)[r * ld + c]);
      else acc += reinterpret_cast<const float*>(dz)[r * ld + c];
    }
  }
  s[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s[k][tx];
    db[c] = v * (alpha ? __ldg(alpha) : 1.f);
  }
}

}  // namespace iif

using namespace iif;

extern "C" int iif_softmax_ce_fwd_bwd(const float* z, int64_t ldz, const float* iifv, const int64_t* label,
                                      const float* class_weight, const float* sample_weight, int64_t ignore_index,
                                      float scale, int64_t B, int64_t C, float* loss_i, float* loss_sum,
                                      float* dz_f32, int64_t lddz_f32, void* dz_bf16, int64_t lddz_bf16, float* lse,
                                      int32_t* argmax, int32_t* rank, int32_t* acc_counts, int32_t* ticket,
                                      void* stream) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !label)) || ldz < C) return IIF_EINVAL;
  if ((dz_f32 && lddz_f32 < C) || (dz_bf16 && lddz_bf16 < C)) return IIF_EINVAL;
  if ((loss_sum && !loss_i) || (acc_counts && !rank) || ((loss_sum || acc_counts) && !ticket)) return IIF_EINVAL;
  if (C > 32768) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  RowArgs a{};
  a.z = z; a.ldz = ldz; a.iif = iifv; a.label = label; a.cw = class_weight; a.sw = sample_weight;
  a.ignore_index = ignore_index; a.scale = scale; a.B = B; a.C = (int)C;
  a.loss_i = loss_i; a.loss_sum = loss_sum; a.dz32 = dz_f32; a.lddz32 = lddz_f32;
  a.dz16 = reinterpret_cast<uint16_t*>(dz_bf16); a.lddz16 = lddz_bf16; a.lse = lse; a.argmax = argmax; a.rank = rank;
  a.acc_counts = acc_counts; a.ticket = ticket; a.on_scaled = 0;
  const bool vec = (C % 4 == 0) && (ldz % 4 == 0) && aligned16(z) && (!iifv || aligned16(iifv)) &&
                   (!dz_f32 || (aligned16(dz_f32) && lddz_f32 % 4 == 0)) &&
                   (!dz_bf16 || ((reinterpret_cast<uintptr_t>(dz_bf16) & 7u) == 0 && lddz_bf16 % 4 == 0));
  return dispatch_row<0>(a, vec, (cudaStream_t)stream);
}

extern "C" int iif_scaled_activation(const float* z, int64_t ldz, const float* iifv, int softmax, int64_t B, int64_t C,
                                     float* out, int64_t ldo, const int64_t* label, int32_t* argmax, int32_t* rank,
                                     void* stream) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !out)) || ldz < C || ldo < C) return IIF_EINVAL;
  if (rank && !label) return IIF_EINVAL;
  if (C > 32768) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  RowArgs a{};
  a.z = z; a.ldz = ldz; a.iif = iifv; a.label = label; a.ignore_index = INT64_MIN; a.B = B; a.C = (int)C;
  a.out = out; a.ldo = ldo; a.softmax = softmax; a.on_scaled = 1; a.argmax = argmax; a.rank = rank;
  const bool vec = (C % 4 == 0) && (ldz % 4 == 0) && (ldo % 4 == 0) && aligned16(z) && aligned16(out) &&
                   (!iifv || aligned16(iifv));
  return dispatch_row<1>(a, vec, (cudaStream_t)stream);
}

extern "C" int iif_sigmoid_bce_fwd_bwd(const float* z, int64_t ldz, const int64_t* label, const float* pos_weight,
                                       const float* col_weight, const float* sample_weight, int64_t ignore_index,
                                       float scale, int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                                       float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                                       int64_t lddz_bf16, int32_t* ticket, void* stream) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !label)) || ldz < C) return IIF_EINVAL;
  if ((dz_f32 && lddz_f32 < C) || (dz_bf16 && lddz_bf16 < C) || (loss_elem && ldl < C)) return IIF_EINVAL;
  if (loss_sum && (!loss_i || !ticket)) return IIF_EINVAL;
  if (C > (1 << 30)) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  BceArgs a{};
  a.z = z; a.ldz = ldz; a.label = label; a.pw = pos_weight; a.colw = col_weight; a.sw = sample_weight;
  a.ignore_index = ignore_index; a.scale = scale; a.B = B; a.C = (int)C; a.loss_elem = loss_elem; a.ldl = ldl;
  a.loss_i = loss_i; a.loss_sum = loss_sum; a.dz32 = dz_f32; a.lddz32 = lddz_f32;
  a.dz16 = reinterpret_cast<uint16_t*>(dz_bf16); a.lddz16 = lddz_bf16; a.ticket = ticket;
  const bool vec = (C % 4 == 0) && (ldz % 4 == 0) && aligned16(z) && (!pos_weight || aligned16(pos_weight)) &&
                   (!col_weight || aligned16(col_weight)) && (!loss_elem || (aligned16(loss_elem) && ldl % 4 == 0)) &&
                   (!dz_f32 || (aligned16(dz_f32) && lddz_f32 % 4 == 0)) &&
                   (!dz_bf16 || ((reinterpret_cast<uintptr_t>(dz_bf16) & 7u) == 0 && lddz_bf16 % 4 == 0));
  const unsigned grid = (unsigned)((B + 1) / 2);
  if (vec) bce_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else bce_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status();
}

extern "C" int iif_scale_rows(const float* in, int64_t ldi, const float* g, int64_t g_stride, int64_t rows, int64_t cols,
                              void* out, int out_dtype, int64_t ldo, void* stream) {
  if (rows < 0 || cols < 0 || ldi < cols || ldo < cols || (g_stride != 0 && g_stride != 1)) return IIF_EINVAL;
  if (out_dtype != IIF_DTYPE_F32 && out_dtype != IIF_DTYPE_BF16) return IIF_EINVAL;
  if (rows == 0 || cols == 0) return IIF_OK;
  if (!in || !out || cols > (1 << 30)) return IIF_EINVAL;
  const bool bf = out_dtype == IIF_DTYPE_BF16;
  const bool vec = (cols % 4 == 0) && (ldi % 4 == 0) && (ldo % 4 == 0) && aligned16(in) &&
                   (bf ? (reinterpret_cast<uintptr_t>(out) & 7u) == 0 : aligned16(out));
  const int64_t items = rows * (vec ? cols / 4 : cols);
  const unsigned grid = (unsigned)((items + 255) / 256 < 8 * kNumSMs ? (items + 255) / 256 : 8 * kNumSMs);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && bf) scale_rows_kernel<true, true><<<grid, 256, 0, st>>>(in, ldi, g, g_stride, rows, (int)cols, out, ldo);
  else if (vec) scale_rows_kernel<true, false><<<grid, 256, 0, st>>>(in, ldi, g, g_stride, rows, (int)cols, out, ldo);
  else if (bf) scale_rows_kernel<false, true><<<grid, 256, 0, st>>>(in, ldi, g, g_stride, rows, (int)cols, out, ldo);
  else scale_rows_kernel<false, false><<<grid, 256, 0, st>>>(in, ldi, g, g_stride, rows, (int)cols, out, ldo);
  return launch_status();
}

extern "C" int iif_colsum(const void* dz, int dz_dtype, int64_t lddz, const float* alpha_dev, int64_t rows, int64_t cols,
                          float* db, void* stream) {
  if (rows < 0 || cols <= 0 || !db || (rows > 0 && !dz) || lddz < cols || cols > (1 << 30)) return IIF_EINVAL;
  if (dz_dtype != IIF_DTYPE_F32 && dz_dtype != IIF_DTYPE_BF16) return IIF_EINVAL;
  const unsigned grid = (unsigned)((cols + 31) / 32);
  if (dz_dtype == IIF_DTYPE_BF16) colsum_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dz, lddz, alpha_dev, rows, (int)cols, db);
  else colsum_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dz, lddz, alpha_dev, rows, (int)cols, db);
  return launch_status();
}
