// Fused row kernels of the IIF head: softmax-CE fwd+bwd with the per-class IIF scale, the
// softmax / scaled activation, sigmoid-BCE fwd+bwd, row scaling / bf16 cast and the bias-gradient
// column sum.  All HBM-bound: each logit is read once (128-bit, streaming), every intermediate
// (scaled logits, probabilities, one-hot targets) lives in registers only.
//
// Reference semantics restated here (never copied): cls/custom.py:28-39,61-73;
// seg/mmdet/models/losses/iif_loss.py:65-78,187-200; cross_entropy_loss.py:53-111;
// losses/utils.py:28-55; losses/accuracy.py:41-50.
#include "loss_row.cuh"

namespace iif {

// MODE 0: softmax-CE forward + backward.  MODE 1: activation (softmax(z*iif) or z*iif).
template <int TPR, int NE, bool VEC, int MODE>
__global__ void __launch_bounds__(TPR > 256 ? TPR : 256, TPR > 256 ? 1 : (NE <= 8 ? 4 : 3))
row_softmax_kernel(const RowArgs a) {
  constexpr int THREADS = TPR > 256 ? TPR : 256;
  __shared__ RowSmem<THREADS> sm;
  __shared__ float s_row_loss[THREADS / TPR];
  __shared__ int s_row_rank[THREADS / TPR];
  ptx::griddep_launch_dependents();   // the successor's prologue may start; its own wait still orders it after us
  ptx::griddep_wait();
  // Row blocks are walked by a bounded grid (launch_row): with one CTA per row block a 64k-row batch sent
  // 32k CTAs through the ticket of the deterministic tail -- same-address atomics serialise in L2 at ~9 ns
  // each, which was the whole kernel time (293 us for 65536 x 1000, 21 % of HBM bandwidth).
  constexpr int RPB = THREADS / TPR;
  const int64_t nblocks = (a.B + RPB - 1) / RPB;
  double part = 0.0;
  int c1 = 0, c5 = 0;
  for (int64_t rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
    float my_loss; int cnt; bool active;
    if (rb + gridDim.x < nblocks) prefetch_row_block<TPR, NE, VEC>(a, rb + gridDim.x);
    softmax_row_body<TPR, NE, VEC, MODE>(a, rb, sm, my_loss, cnt, active);
    if constexpr (MODE == 0) {
      if (a.loss_sum || a.acc_counts) {
        // CTA partial in row order
        const int t = threadIdx.x % TPR, lrow = threadIdx.x / TPR;
        if (t == 0) { s_row_loss[lrow] = active ? my_loss : 0.f; s_row_rank[lrow] = active ? cnt : 0x7fffffff; }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
          for (int r = 0; r < RPB; ++r) {
            part += (double)s_row_loss[r];
            c1 += s_row_rank[r] < 1; c5 += s_row_rank[r] < 5;
          }
        }
      }
    }
    __syncthreads();                      // the reduction scratch is reused by the next row block
  }
  if constexpr (MODE == 0) {
    if (a.loss_sum || a.acc_counts) grid_tail<THREADS>(part, c1, c5, a.loss_sum, a.acc_counts, a.scratch);
  }
}

template <int TPR, int NE, int MODE>
static int launch_row(const RowArgs& a, bool vec, cudaStream_t st) {
  constexpr int THREADS = TPR > 256 ? TPR : 256;
  constexpr int RPB = THREADS / TPR;
  cudaLaunchConfig_t cfg{};
  const int64_t nblocks = (a.B + RPB - 1) / RPB;
  const int64_t cap = (int64_t)kNumSMs * (THREADS >= 1024 ? 2 : (THREADS >= 512 ? 4 : 8));   // one resident wave
  cfg.gridDim = dim3((unsigned)(nblocks < cap ? nblocks : cap));
  cfg.blockDim = dim3(THREADS);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = vec ? cudaLaunchKernelEx(&cfg, row_softmax_kernel<TPR, NE, true, MODE>, a)
                      : cudaLaunchKernelEx(&cfg, row_softmax_kernel<TPR, NE, false, MODE>, a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return e == cudaSuccess ? IIF_OK : (int)e;
}

// rows per CTA of the configuration dispatch_row would pick (for the scratch size)
static int row_config(int64_t B, int C, int* tpr, int* ne) {
  // Small batches (latency): wide rows, one 128-bit load per thread, every SM busy.
  // Big batches (throughput): FEW threads per row, 16-32 logits each -- the per-row fixed cost (two
  // reductions, label bookkeeping, pointer arithmetic: ~250 instructions per thread) is what made the
  // 8-logits-per-thread configuration issue-bound at 35 % of HBM bandwidth; one warp per row also drops both
  // block barriers.
  const bool big = B > 2048;
  if (C <= 128) { *tpr = 32; *ne = 4; }
  else if (C <= 256) { if (big) { *tpr = 32; *ne = 8; } else { *tpr = 64; *ne = 4; } }
  else if (C <= 512) { if (big) { *tpr = 32; *ne = 16; } else { *tpr = 128; *ne = 4; } }
  else if (C <= 1024) { if (big) { *tpr = 32; *ne = 32; } else { *tpr = 256; *ne = 4; } }
  else if (C <= 2048) { if (big) { *tpr = 64; *ne = 32; } else { *tpr = 256; *ne = 8; } }
  else if (C <= 4096) { if (big) { *tpr = 128; *ne = 32; } else { *tpr = 256; *ne = 16; } }
  else if (C <= 8192) { *tpr = 256; *ne = 32; }
  else if (C <= 10240) { *tpr = 256; *ne = 40; }        // the 10000-class sweep: one block per row, no padding waste
  else if (C <= 16384) { *tpr = 512; *ne = 32; }
  else if (C <= 32768) { *tpr = 1024; *ne = 32; }
  else return IIF_EUNSUPPORTED;
  return IIF_OK;
}

template <int MODE>
static int dispatch_row(const RowArgs& a, bool vec, cudaStream_t st) {
  int tpr, ne;
  if (int rc = row_config(a.B, a.C, &tpr, &ne)) return rc;
#define IIF_ROW(T, N) if (tpr == T && ne == N) return launch_row<T, N, MODE>(a, vec, st)
  IIF_ROW(32, 4); IIF_ROW(32, 8); IIF_ROW(64, 4); IIF_ROW(64, 8); IIF_ROW(128, 4); IIF_ROW(128, 8);
  IIF_ROW(256, 4); IIF_ROW(256, 8); IIF_ROW(256, 16); IIF_ROW(256, 32); IIF_ROW(512, 32); IIF_ROW(1024, 32);
  IIF_ROW(32, 16); IIF_ROW(32, 32); IIF_ROW(64, 32); IIF_ROW(128, 32); IIF_ROW(256, 40);
#undef IIF_ROW
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
  float* dz32; int64_t lddz32; uint16_t* dz16; int64_t lddz16; int32_t* scratch;
  float gamma, alpha;                  // focal modulation (gamma > 0) and class balance (alpha > 0), cls/custom.py:74-89
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

// Focal branch of cls/custom.py:74-89 in logit space (the reference goes sigmoid -> BCELoss -> pow in fp32):
//   q = p_t = t p + (1-t)(1-p),  L = -log(q) (1-q)^gamma * alpha_t,
//   t = 1: L = sp(-z) e^{-gamma sp(z)},  dL/dz = -(1-p)^gamma [(1-p) + gamma p sp(-z)]
//   t = 0: L = sp(z) e^{-gamma sp(-z)},  dL/dz =  p^gamma [p + gamma (1-p) sp(z)]        (sp = softplus)
__device__ __forceinline__ void focal_elem(float z, bool t, float gamma, float alpha, float wgt, float& loss, float& d) {
  const float e = expf(-fabsf(z));
  const float l1 = log1pf(e);
  const float sp_pos = l1 + fmaxf(z, 0.f), sp_neg = l1 + fmaxf(-z, 0.f);   // softplus(z), softplus(-z)
  const float p = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  const float at = alpha > 0.f ? (t ? alpha : 1.f - alpha) : 1.f;
  const float w = wgt * at;
  if (t) {
    const float mod = expf(-gamma * sp_pos);        // (1-p)^gamma
    loss = w * sp_neg * mod;
    d = -w * mod * ((1.f - p) + gamma * p * sp_neg);
  } else {
    const float mod = expf(-gamma * sp_neg);        // p^gamma
    loss = w * sp_pos * mod;
    d = w * mod * (p + gamma * (1.f - p) * sp_pos);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) bce_kernel(const BceArgs a) {
  constexpr int TPR = 128;
  __shared__ float s_sum[8];
  ptx::griddep_launch_dependents();   // the successor's prologue may start; its own wait still orders it after us
  ptx::griddep_wait();
  const int t = threadIdx.x % TPR;
  const int C = a.C;
  double part = 0.0;                         // thread 0: this CTA's loss, row pairs in order
  // bounded grid walking the row pairs (see row_softmax_kernel: the tail's ticket must not see 32k CTAs)
  for (int64_t rp = blockIdx.x; rp * 2 < a.B; rp += gridDim.x) {
  const int64_t row = rp * 2 + threadIdx.x / TPR;
  const bool active = row < a.B;
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
        if (a.gamma > 0.f) {
          focal_elem(z4.x, col + 0 == yi, a.gamma, a.alpha, wrow * c4.x, l4.x, d4.x);
          focal_elem(z4.y, col + 1 == yi, a.gamma, a.alpha, wrow * c4.y, l4.y, d4.y);
          focal_elem(z4.z, col + 2 == yi, a.gamma, a.alpha, wrow * c4.z, l4.z, d4.z);
          focal_elem(z4.w, col + 3 == yi, a.gamma, a.alpha, wrow * c4.w, l4.w, d4.w);
        } else {
        bce_elem(z4.x, col + 0 == yi, p4.x, wrow * c4.x, l4.x, d4.x);
        bce_elem(z4.y, col + 1 == yi, p4.y, wrow * c4.y, l4.y, d4.y);
        bce_elem(z4.z, col + 2 == yi, p4.z, wrow * c4.z, l4.z, d4.z);
        bce_elem(z4.w, col + 3 == yi, p4.w, wrow * c4.w, l4.w, d4.w);
        }
        if (!valid) { l4 = make_float4(0.f, 0.f, 0.f, 0.f); d4 = l4; }
        acc += (l4.x + l4.y) + (l4.z + l4.w);
        if (le) stg_stream4(le + col, l4);
        if (d32) stg_stream4(d32 + col, d4);
        if (d16) stg_stream2(d16 + col, pack_bf16x2(d4.x, d4.y), pack_bf16x2(d4.z, d4.w));
      }
    } else {
      for (int col = t; col < C; col += TPR) {
        float l, d;
        if (a.gamma > 0.f)
          focal_elem(__ldg(zr + col), col == yi, a.gamma, a.alpha, wrow * (a.colw ? __ldg(a.colw + col) : 1.f), l, d);
        else
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
  acc = warp_sum(acc);                      // TPR = 128: 4 warps per row, 2 rows per CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_sum[warp] = acc;
  __syncthreads();
  const int w0 = (warp / 4) * 4;
  acc = (s_sum[w0] + s_sum[w0 + 1]) + (s_sum[w0 + 2] + s_sum[w0 + 3]);
  if (active && t == 0 && a.loss_i) a.loss_i[row] = acc;
  if (a.loss_sum && threadIdx.x == 0) {
    part += (double)((s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]));
    if (rp * 2 + 1 < a.B) part += (double)((s_sum[4] + s_sum[5]) + (s_sum[6] + s_sum[7]));
  }
  __syncthreads();                           // s_sum is reused by the next row pair
  }
  if (a.loss_sum) grid_tail<256>(part, 0, 0, a.loss_sum, nullptr, a.scratch);
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
// data *= *g in place over a flat buffer; returns at once when *g == 1 (the upstream gradient of a loss that is
// the autograd root): the one-launch head step forms dX / dW / db in the FORWARD call with g = 1 and the autograd
// backward only has to apply the actual upstream scalar -- almost always exactly 1.
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) scale_inplace_kernel(void* __restrict__ data, int64_t n, const float* __restrict__ g) {
  const float gv = __ldg(g);
  if (gv == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if constexpr (BF16) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(data) + i;
      *p = __float2bfloat16_rn(__bfloat162float(*p) * gv);
    } else {
      reinterpret_cast<float*>(data)[i] *= gv;
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

// one partial per CTA: at most B CTAs (one row each) or the 2 x 148 CTAs of the loss-fused backward launch
extern "C" size_t iif_loss_scratch_bytes(int64_t B) { return scratch_bytes_for(B > 512 ? B : 512); }

extern "C" int iif_softmax_ce_fwd_bwd(const float* z, int64_t ldz, const float* iifv, const int64_t* label,
                                      const float* class_weight, const float* sample_weight, int64_t ignore_index,
                                      float scale, int64_t B, int64_t C, float* loss_i, float* loss_sum,
                                      float* dz_f32, int64_t lddz_f32, void* dz_bf16, int64_t lddz_bf16, float* lse,
                                      int32_t* argmax, int32_t* rank, int32_t* acc_counts, int32_t* scratch,
                                      void* stream) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !label)) || ldz < C) return IIF_EINVAL;
  if ((dz_f32 && lddz_f32 < C) || (dz_bf16 && lddz_bf16 < C)) return IIF_EINVAL;
  if ((acc_counts && !rank) || ((loss_sum || acc_counts) && !scratch)) return IIF_EINVAL;
  if (C > 32768) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  RowArgs a{};
  const bool vec = make_ce_row_args(a, z, ldz, iifv, label, class_weight, sample_weight, ignore_index, scale, B, C, loss_i,
                                    loss_sum, dz_f32, lddz_f32, dz_bf16, lddz_bf16, lse, argmax, rank, acc_counts, scratch);
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

static int sigmoid_loss(const float* z, int64_t ldz, const int64_t* label, const float* pos_weight,
                        const float* col_weight, const float* sample_weight, int64_t ignore_index,
                        float scale, int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                        float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                        int64_t lddz_bf16, int32_t* scratch, void* stream, float gamma, float alpha) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !label)) || ldz < C) return IIF_EINVAL;
  if ((dz_f32 && lddz_f32 < C) || (dz_bf16 && lddz_bf16 < C) || (loss_elem && ldl < C)) return IIF_EINVAL;
  if (loss_sum && !scratch) return IIF_EINVAL;
  if (C > (1 << 30)) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  BceArgs a{};
  a.z = z; a.ldz = ldz; a.label = label; a.pw = pos_weight; a.colw = col_weight; a.sw = sample_weight;
  a.ignore_index = ignore_index; a.scale = scale; a.B = B; a.C = (int)C; a.loss_elem = loss_elem; a.ldl = ldl;
  a.loss_i = loss_i; a.loss_sum = loss_sum; a.dz32 = dz_f32; a.lddz32 = lddz_f32;
  a.dz16 = reinterpret_cast<uint16_t*>(dz_bf16); a.lddz16 = lddz_bf16; a.scratch = scratch;
  a.gamma = gamma; a.alpha = alpha;
  const bool vec = (C % 4 == 0) && (ldz % 4 == 0) && aligned16(z) && (!pos_weight || aligned16(pos_weight)) &&
                   (!col_weight || aligned16(col_weight)) && (!loss_elem || (aligned16(loss_elem) && ldl % 4 == 0)) &&
                   (!dz_f32 || (aligned16(dz_f32) && lddz_f32 % 4 == 0)) &&
                   (!dz_bf16 || ((reinterpret_cast<uintptr_t>(dz_bf16) & 7u) == 0 && lddz_bf16 % 4 == 0));
  const int64_t pairs = (B + 1) / 2;
  const unsigned grid = (unsigned)(pairs < 8 * kNumSMs ? pairs : 8 * kNumSMs);
  if (vec) bce_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else bce_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status();
}

extern "C" int iif_sigmoid_bce_fwd_bwd(const float* z, int64_t ldz, const int64_t* label, const float* pos_weight,
                                       const float* col_weight, const float* sample_weight, int64_t ignore_index,
                                       float scale, int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                                       float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                                       int64_t lddz_bf16, int32_t* scratch, void* stream) {
  return sigmoid_loss(z, ldz, label, pos_weight, col_weight, sample_weight, ignore_index, scale, B, C, loss_elem, ldl,
                      loss_i, loss_sum, dz_f32, lddz_f32, dz_bf16, lddz_bf16, scratch, stream, 0.f, 0.f);
}

extern "C" int iif_sigmoid_focal_fwd_bwd(const float* z, int64_t ldz, const int64_t* label, float gamma, float alpha,
                                         const float* col_weight, const float* sample_weight, int64_t ignore_index,
                                         float scale, int64_t B, int64_t C, float* loss_elem, int64_t ldl, float* loss_i,
                                         float* loss_sum, float* dz_f32, int64_t lddz_f32, void* dz_bf16,
                                         int64_t lddz_bf16, int32_t* scratch, void* stream) {
  if (!(gamma > 0.f) || alpha >= 1.f) return IIF_EINVAL;
  return sigmoid_loss(z, ldz, label, nullptr, col_weight, sample_weight, ignore_index, scale, B, C, loss_elem, ldl,
                      loss_i, loss_sum, dz_f32, lddz_f32, dz_bf16, lddz_bf16, scratch, stream, gamma, alpha > 0.f ? alpha : 0.f);
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

extern "C" int iif_scale_inplace(void* data, int dtype, int64_t n, const float* g_dev, void* stream) {
  if (n < 0 || !g_dev || (dtype != IIF_DTYPE_F32 && dtype != IIF_DTYPE_BF16)) return IIF_EINVAL;
  if (n == 0) return IIF_OK;
  if (!data) return IIF_EINVAL;
  const unsigned grid = (unsigned)((n + 255) / 256 < 8 * kNumSMs ? (n + 255) / 256 : 8 * kNumSMs);
  if (dtype == IIF_DTYPE_BF16) scale_inplace_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(data, n, g_dev);
  else scale_inplace_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(data, n, g_dev);
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

// Mixup: lam * CE(label_a) + (1 - lam) * CE(label_b) and its gradient from ONE pass over the logits
// (cls/custom.py:116-117 calls the criterion twice on the same logits).  128-bit path only:
// IIF_EUNSUPPORTED when C % 4 != 0 or the rows are unaligned (the caller then runs two passes).
extern "C" int iif_softmax_ce_mixup_fwd_bwd(const float* z, int64_t ldz, const float* iifv, const int64_t* label_a,
                                            const int64_t* label_b, float lam, const float* class_weight,
                                            const float* sample_weight, int64_t ignore_index, float scale, int64_t B,
                                            int64_t C, float* loss_i, float* loss_sum, float* dz_f32, int64_t lddz_f32,
                                            void* dz_bf16, int64_t lddz_bf16, int32_t* argmax, int32_t* rank,
                                            int32_t* acc_counts, int32_t* scratch, void* stream) {
  if (B < 0 || C <= 0 || (B > 0 && (!z || !label_a || !label_b)) || ldz < C) return IIF_EINVAL;
  if ((dz_f32 && lddz_f32 < C) || (dz_bf16 && lddz_bf16 < C)) return IIF_EINVAL;
  if ((acc_counts && !rank) || ((loss_sum || acc_counts) && !scratch)) return IIF_EINVAL;
  if (C > 32768) return IIF_EUNSUPPORTED;
  if (B == 0) return IIF_OK;
  RowArgs a{};
  const bool vec = make_ce_row_args(a, z, ldz, iifv, label_a, class_weight, sample_weight, ignore_index, scale, B, C, loss_i,
                                    loss_sum, dz_f32, lddz_f32, dz_bf16, lddz_bf16, nullptr, argmax, rank, acc_counts, scratch);
  if (!vec) return IIF_EUNSUPPORTED;
  a.label_b = label_b;
  a.lam = lam;
  return dispatch_row<0>(a, true, (cudaStream_t)stream);
}
