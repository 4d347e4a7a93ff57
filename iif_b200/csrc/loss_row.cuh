// Row kernels of the fused IIF softmax-CE (shared by loss.cu and the loss-fused backward GEMM launch of
// gemm_tc.cu).  Reference semantics restated (never copied): cls/custom.py:28-39;
// seg/mmdet/models/losses/iif_loss.py:65-78,187-200; losses/utils.py:28-55; losses/accuracy.py:41-50.
#pragma once
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

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
  float* lse; int32_t* argmax; int32_t* rank; int32_t* acc_counts; int32_t* scratch;
  float* out; int64_t ldo; int softmax; int on_scaled;
  // dual-label (Mixup) loss: lam * CE(label) + (1 - lam) * CE(label_b) from ONE softmax pass (128-bit path only)
  const int64_t* label_b; float lam;
};

// Fill the arguments of the softmax-CE rows; returns whether the 128-bit (VEC) path is legal.
inline bool make_ce_row_args(RowArgs& a, const float* z, int64_t ldz, const float* iifv, const int64_t* label,
                             const float* class_weight, const float* sample_weight, int64_t ignore_index, float scale,
                             int64_t B, int64_t C, float* loss_i, float* loss_sum, float* dz_f32, int64_t lddz_f32,
                             void* dz_bf16, int64_t lddz_bf16, float* lse, int32_t* argmax, int32_t* rank,
                             int32_t* acc_counts, int32_t* scratch) {
  a = RowArgs{};
  a.z = z; a.ldz = ldz; a.iif = iifv; a.label = label; a.cw = class_weight; a.sw = sample_weight;
  a.ignore_index = ignore_index; a.scale = scale; a.B = B; a.C = (int)C;
  a.loss_i = loss_i; a.loss_sum = loss_sum; a.dz32 = dz_f32; a.lddz32 = lddz_f32;
  a.dz16 = reinterpret_cast<uint16_t*>(dz_bf16); a.lddz16 = lddz_bf16; a.lse = lse; a.argmax = argmax; a.rank = rank;
  a.acc_counts = acc_counts; a.scratch = scratch; a.on_scaled = 0;
  return (C % 4 == 0) && (ldz % 4 == 0) && aligned16(z) && (!iifv || aligned16(iifv)) &&
         (!dz_f32 || (aligned16(dz_f32) && lddz_f32 % 4 == 0)) &&
         (!dz_bf16 || ((reinterpret_cast<uintptr_t>(dz_bf16) & 7u) == 0 && lddz_bf16 % 4 == 0));
}

// scratch layout: [0] ticket (zero on entry, reset on exit) | 16: double part[grid] | int c1[grid] | int c5[grid]
__host__ __device__ inline size_t scratch_bytes_for(int64_t grid) { return 16 + (size_t)grid * 16; }

template <int THREADS>
__device__ __forceinline__ double block_sum_d(double v, double* s) {   // fixed order: deterministic
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) r += s[w];
  __syncthreads();
  return r;
}
template <int THREADS>
__device__ __forceinline__ int block_sum_i(int v, int* s) {
  v = warp_sum_i(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) r += s[w];
  __syncthreads();
  return r;
}

// Grid-wide tail: every CTA has published one partial (loss sum, top-1 / top-5 hits); the last CTA
// to take a ticket adds them in index order.  Only thread 0 fences -- the dZ stores of the other
// threads are ordered by the kernel boundary, not by this reduction.
// Returns true (CTA-uniform) in the last CTA, after the totals are written.
template <int THREADS>
__device__ __forceinline__ bool grid_tail(double part, int c1, int c5, float* loss_sum, int32_t* acc_counts,
                                          int32_t* scratch) {
  __shared__ int s_last;
  __shared__ double s_d[THREADS / 32];
  __shared__ int s_i[THREADS / 32];
  double* g_part = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(scratch) + 16);
  int* g_c1 = reinterpret_cast<int*>(g_part + gridDim.x);
  int* g_c5 = g_c1 + gridDim.x;
  if (threadIdx.x == 0) {
    __stcg(g_part + blockIdx.x, part);
    __stcg(g_c1 + blockIdx.x, c1);
    __stcg(g_c5 + blockIdx.x, c5);
    // release our partial / acquire everyone else's in one gpu-scope RMW (no MEMBAR.SC on the tail)
    s_last = (ptx::atom_add_acq_rel(scratch, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  double acc = 0.0;
  int k1 = 0, k5 = 0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += THREADS) {
    acc += __ldcg(g_part + i); k1 += __ldcg(g_c1 + i); k5 += __ldcg(g_c5 + i);
  }
  acc = block_sum_d<THREADS>(acc, s_d);
  k1 = block_sum_i<THREADS>(k1, s_i);
  k5 = block_sum_i<THREADS>(k5, s_i);
  if (threadIdx.x == 0) {
    if (loss_sum) *loss_sum = (float)acc;
    if (acc_counts) { acc_counts[0] = k1; acc_counts[1] = k5; }
    *scratch = 0;   // self-resetting ticket
  }
  return true;
}

// MODE 0: softmax-CE forward + backward.  MODE 1: activation (softmax(z*iif) or z*iif).
// TPR threads share one row (NE elements each); a CTA of max(TPR,256) threads holds 256/TPR rows.
// Small batches use wide rows (TPR = C/4: one 128-bit load per thread, short dependency chains, every
// SM busy); large batches use NE = 8 for more bytes in flight per SM.
struct NoHook { __device__ __forceinline__ void operator()() const {} };
// ZLoad: where the vec body gets a float4 of raw logits from.  NoZLoad = the logits tensor a.z (streaming
// 128-bit loads); the fused head step passes a functor that sums the forward GEMM's split-K partial tiles
// instead (head_fused.cu) -- it is called as zload(col, row) for in-range float4 groups only and must
// return -inf for elements at or beyond column C.
struct NoZLoad {};

template <int THREADS>
struct RowSmem {
  float f[5][THREADS / 32];
  int i[THREADS / 32];
};

__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2: 2 ulp, exp2(-inf) = 0, NaN propagates
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Lean 128-bit path of the rows (C % 4 == 0, aligned): ~15 instructions per logit instead of ~110 in the
// first version (whose 65536 x 1000 launch was ISSUE-bound at 76 % issue-slot utilisation and 23 % of HBM
// bandwidth -- profiles/): every value is loaded once into registers as float4 (logits AND IIF weights),
// out-of-row elements are encoded as z = -inf, s = 1 (so no per-element bounds logic survives the loads),
// the label's column is found with one range test per float4, optional outputs (argmax / rank) sit behind
// warp-uniform branches, exp is one FMUL + MUFU.EX2 on (a - max) * log2(e).
// The label-dependent scalars and the IIF weights of one row: everything the row needs that does NOT depend on the
// logits.  Split from the rest of the body so that a caller whose logits arrive late (the one-launch head step waits
// for the forward GEMM's partial tiles) can have these loads in flight while it waits.
template <int TPR, int NE>
struct RowHead {
  static constexpr int NQ = NE / 4;
  static constexpr bool CACHE_S = NE <= 8;   // wider rows re-read the IIF weights (L1) instead of holding them
  int64_t row; bool active, y_ok, dual; int yi, ybi; float g, gb;
  float4 s4[CACHE_S ? NQ : 1];
};

template <int TPR>
__device__ __forceinline__ float4 row_load_s(const RowArgs& a, int q, int t, bool active) {
  const int col = (q * TPR + t) * 4, C = a.C;
  if (!(a.iif && active && col < C)) return make_float4(1.f, 1.f, 1.f, 1.f);
  if (col + 4 <= C) return __ldg(reinterpret_cast<const float4*>(a.iif + col));
  // ragged last group (C % 4 != 0: only reachable through a ZLoad functor): no read past the vector's end
  return make_float4(__ldg(a.iif + col), col + 1 < C ? __ldg(a.iif + col + 1) : 1.f, col + 2 < C ? __ldg(a.iif + col + 2) : 1.f, 1.f);
}

template <int TPR, int NE, int MODE>
__device__ __forceinline__ void row_head(const RowArgs& a, int64_t row_block, RowHead<TPR, NE>& h) {
  constexpr int THREADS = TPR > 256 ? TPR : 256;
  const int t = threadIdx.x % TPR;
  const int lrow = threadIdx.x / TPR;
  h.row = row_block * (THREADS / TPR) + lrow;
  h.active = h.row < a.B;
  const int C = a.C;
  int64_t y = -1;
  if (h.active && a.label) y = __ldg(a.label + h.row);
  const bool y_in = h.active && y >= 0 && y < C;
  h.y_ok = y_in && y != a.ignore_index;
  h.yi = y_in ? (int)y : -1;
  h.g = 0.f;
  if (MODE == 0 && h.y_ok) {
    h.g = a.scale;
    if (a.cw) h.g *= __ldg(a.cw + y);
    if (a.sw) h.g *= __ldg(a.sw + h.row);
  }
  // second label of a Mixup pair (cls/custom.py:116-117): same rules, weight (1 - lam); the first gets lam
  h.dual = MODE == 0 && a.label_b != nullptr;
  h.ybi = -1;
  h.gb = 0.f;
  if (h.dual) {
    const int64_t yb = h.active ? __ldg(a.label_b + h.row) : -1;
    const bool yb_in = h.active && yb >= 0 && yb < C;
    h.ybi = yb_in ? (int)yb : -1;
    if (yb_in && yb != a.ignore_index) {
      h.gb = a.scale * (1.f - a.lam);
      if (a.cw) h.gb *= __ldg(a.cw + yb);
      if (a.sw) h.gb *= __ldg(a.sw + h.row);
    }
    h.g *= a.lam;
  }
  if constexpr (RowHead<TPR, NE>::CACHE_S) {
#pragma unroll
    for (int q = 0; q < NE / 4; ++q) h.s4[q] = row_load_s<TPR>(a, q, t, h.active);
  }
}

template <int TPR, int NE, int MODE, class Hook>
__device__ __forceinline__ void row_tail(const RowArgs& a, const RowHead<TPR, NE>& h, float4 (&z4)[NE / 4],
                                         RowSmem<(TPR > 256 ? TPR : 256)>& sm, float& my_loss_out, int& cnt_out,
                                         bool& active_out, Hook hook);

template <int TPR, int NE, int MODE, class Hook, class ZLoad = NoZLoad>
__device__ __forceinline__ void softmax_row_body_vec(const RowArgs& a, int64_t row_block,
                                                     RowSmem<(TPR > 256 ? TPR : 256)>& sm, float& my_loss_out,
                                                     int& cnt_out, bool& active_out, Hook hook, ZLoad zload = ZLoad()) {
  constexpr int NQ = NE / 4;
  RowHead<TPR, NE> h;
  row_head<TPR, NE, MODE>(a, row_block, h);
  const int t = threadIdx.x % TPR;
  const int C = a.C;
  const float* zr = a.z + (h.active ? h.row : 0) * a.ldz;
  float4 z4[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int col = (q * TPR + t) * 4;
    if constexpr (std::is_same<ZLoad, NoZLoad>::value)
      z4[q] = (h.active && col < C) ? ldg_stream4(zr + col)
                                    : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    else
      z4[q] = (h.active && col < C) ? zload(col, h.row)
                                    : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  }
  row_tail<TPR, NE, MODE>(a, h, z4, sm, my_loss_out, cnt_out, active_out, hook);
}

// Lean 128-bit path of the rows, continued: everything after the loads.
template <int TPR, int NE, int MODE, class Hook>
__device__ __forceinline__ void row_tail(const RowArgs& a, const RowHead<TPR, NE>& h, float4 (&z4)[NE / 4],
                                         RowSmem<(TPR > 256 ? TPR : 256)>& sm, float& my_loss_out, int& cnt_out,
                                         bool& active_out, Hook hook) {
  constexpr int WPR = TPR / 32;
  constexpr int NQ = NE / 4;
  constexpr bool CACHE_S = RowHead<TPR, NE>::CACHE_S;
  constexpr float L2E = 1.4426950408889634f;
  auto& s_f = sm.f;
  auto& s_i = sm.i;
  const int t = threadIdx.x % TPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = (warp / WPR) * WPR;
  const int64_t row = h.row;
  const bool active = h.active;
  const int C = a.C;
  const bool want_rank = a.rank != nullptr, want_arg = a.argmax != nullptr;
  const bool raw_only = (MODE == 1 && !a.softmax);      // activation mode without softmax: out = z * iif
  const bool y_ok = h.y_ok, dual = h.dual;
  const int yi = h.yi, ybi = h.ybi;
  const float g = h.g, gb = h.gb;
  auto load_s = [&](int q) -> float4 { return row_load_s<TPR>(a, q, t, active); };
  const float4* s4 = h.s4;

  // Register diet (the row loop of a big batch lives on occupancy): only the raw logits (later overwritten
  // by their exponentials) and, for narrow rows, the IIF weights stay in registers; z * s is one FMUL
  // wherever it is needed again.
  auto sval = [&](int q) -> float4 { return CACHE_S ? s4[CACHE_S ? q : 0] : load_s(q); };
  auto adj = [&](int q, const float4& sv) -> float4 {
    return make_float4(z4[q].x * sv.x, z4[q].y * sv.y, z4[q].z * sv.z, z4[q].w * sv.w);
  };

  // ---- pass 1: row max of the adjusted logits, label's raw / adjusted logit, optional arg max
  float m = -CUDART_INF_F, zy = -CUDART_INF_F, ay = 0.f, ayb = 0.f;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int col = (q * TPR + t) * 4;
    const float4 av = adj(q, sval(q));
    m = fmaxf(m, fmaxf(fmaxf(av.x, av.y), fmaxf(av.z, av.w)));
    const unsigned d = (unsigned)(yi - col);
    if (d < 4u) {                                        // this float4 holds the label's column
      zy = d == 0 ? z4[q].x : (d == 1 ? z4[q].y : (d == 2 ? z4[q].z : z4[q].w));
      ay = d == 0 ? av.x : (d == 1 ? av.y : (d == 2 ? av.z : av.w));
    }
    const unsigned db = (unsigned)(ybi - col);
    if (db < 4u) ayb = db == 0 ? av.x : (db == 1 ? av.y : (db == 2 ? av.z : av.w));
    if (want_arg) {
      const float4 c4 = a.on_scaled ? av : z4[q];
      if (c4.x > bv) { bv = c4.x; bi = col; }
      if (c4.y > bv) { bv = c4.y; bi = col + 1; }
      if (c4.z > bv) { bv = c4.z; bi = col + 2; }
      if (c4.w > bv) { bv = c4.w; bi = col + 3; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    zy = fmaxf(zy, __shfl_xor_sync(0xffffffffu, zy, o));
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    if (dual) ayb += __shfl_xor_sync(0xffffffffu, ayb, o);
    if (want_arg) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
  }
  if constexpr (WPR > 1) {
    if (lane == 0) {
      s_f[0][warp] = m; s_f[1][warp] = zy; s_f[2][warp] = bv; s_f[3][warp] = ay; s_f[4][warp] = ayb; s_i[warp] = bi;
    }
    __syncthreads();
    hook();
    m = s_f[0][w0]; zy = s_f[1][w0]; bv = s_f[2][w0]; ay = s_f[3][w0]; ayb = s_f[4][w0]; bi = s_i[w0];
#pragma unroll
    for (int w = 1; w < WPR; ++w) {
      m = fmaxf(m, s_f[0][w0 + w]);
      zy = fmaxf(zy, s_f[1][w0 + w]);
      ay += s_f[3][w0 + w];
      ayb += s_f[4][w0 + w];
      const float ov = s_f[2][w0 + w]; const int oi = s_i[w0 + w];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
  } else {
    hook();
  }
  if (yi < 0) { zy = 0.f; ay = 0.f; }
  if (ybi < 0) ayb = 0.f;

  if (raw_only) {
    float* o = a.out + (active ? row : 0) * a.ldo;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int col = (q * TPR + t) * 4;
      if (active && col < C) stg_stream4(o + col, adj(q, sval(q)));
    }
  }

  // ---- pass 2: optional rank of the label (needs the logits), then exp-sum (overwrites them)
  const float mm = (m == -CUDART_INF_F) ? 0.f : m;
  float sum = 0.f;
  int cnt = 0;
  if (want_rank) {
    const bool on_adj = a.on_scaled || raw_only;
    const float ref = on_adj ? ay : zy;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int col = (q * TPR + t) * 4;
      if (active && col < C) {
        const float4 c4 = on_adj ? adj(q, sval(q)) : z4[q];
        cnt += (c4.x > ref) || (c4.x == ref && col < yi);
        cnt += (c4.y > ref) || (c4.y == ref && col + 1 < yi);
        cnt += (c4.z > ref) || (c4.z == ref && col + 2 < yi);
        cnt += (c4.w > ref) || (c4.w == ref && col + 3 < yi);
      }
    }
  }
  if (!raw_only) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 av = adj(q, sval(q));
      z4[q].x = fast_exp2((av.x - mm) * L2E);             // from here on z4 holds exp(a - max)
      z4[q].y = fast_exp2((av.y - mm) * L2E);
      z4[q].z = fast_exp2((av.z - mm) * L2E);
      z4[q].w = fast_exp2((av.w - mm) * L2E);
      sum += (z4[q].x + z4[q].y) + (z4[q].z + z4[q].w);
    }
  }
  sum = warp_sum(sum);
  if (want_rank) cnt = warp_sum_i(cnt);
  if constexpr (WPR > 1) {
    if (lane == 0) { s_f[0][warp] = sum; s_i[warp] = cnt; }
    __syncthreads();
    sum = s_f[0][w0]; cnt = s_i[w0];
#pragma unroll
    for (int w = 1; w < WPR; ++w) { sum += s_f[0][w0 + w]; cnt += s_i[w0 + w]; }
  }
  if (yi < 0) cnt = C;  // label outside [0,C): never inside any top-k

  float my_loss = 0.f;
  if (!raw_only) {
    const float inv = 1.f / sum;
    if constexpr (MODE == 1) {
      float* o = a.out + (active ? row : 0) * a.ldo;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int col = (q * TPR + t) * 4;
        if (active && col < C)
          stg_stream4(o + col, make_float4(z4[q].x * inv, z4[q].y * inv, z4[q].z * inv, z4[q].w * inv));
      }
    } else {
      const float lse = mm + logf(sum);
      my_loss = y_ok ? g * (lse - ay) : 0.f;
      if (dual && gb != 0.f) my_loss += gb * (lse - ayb);
      if (active && t == 0) {
        if (a.loss_i) a.loss_i[row] = my_loss;
        if (a.lse) a.lse[row] = lse;
      }
      if (a.dz32 || a.dz16) {
        float* d32 = a.dz32 ? a.dz32 + (active ? row : 0) * a.lddz32 : nullptr;
        uint16_t* d16 = a.dz16 ? a.dz16 + (active ? row : 0) * a.lddz16 : nullptr;
        const float ga = y_ok ? g : 0.f, gsum = ga + gb;          // dual: dz = s (G p - ga 1[ya] - gb 1[yb])
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int col = (q * TPR + t) * 4;
          if (active && col < C) {
            const float4 sv = sval(q);
            const unsigned dl = (unsigned)(yi - col);
            float4 d;
            if (dual) {
              const unsigned dlb = (unsigned)(ybi - col);
              d.x = sv.x * (gsum * (z4[q].x * inv) - (dl == 0u ? ga : 0.f) - (dlb == 0u ? gb : 0.f));
              d.y = sv.y * (gsum * (z4[q].y * inv) - (dl == 1u ? ga : 0.f) - (dlb == 1u ? gb : 0.f));
              d.z = sv.z * (gsum * (z4[q].z * inv) - (dl == 2u ? ga : 0.f) - (dlb == 2u ? gb : 0.f));
              d.w = sv.w * (gsum * (z4[q].w * inv) - (dl == 3u ? ga : 0.f) - (dlb == 3u ? gb : 0.f));
              if (gsum == 0.f) d = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
            d.x = (sv.x * g) * (z4[q].x * inv - (dl == 0u ? 1.f : 0.f));
            d.y = (sv.y * g) * (z4[q].y * inv - (dl == 1u ? 1.f : 0.f));
            d.z = (sv.z * g) * (z4[q].z * inv - (dl == 2u ? 1.f : 0.f));
            d.w = (sv.w * g) * (z4[q].w * inv - (dl == 3u ? 1.f : 0.f));
            if (!y_ok) d = make_float4(0.f, 0.f, 0.f, 0.f);  // ignored row: exact zeros even for inf weights
            }
            if (d32) stg_stream4(d32 + col, d);
            if (d16) stg_stream2(d16 + col, pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
          }
        }
      }
    }
  }
  if (active && t == 0) {
    if (want_arg) a.argmax[row] = bi;
    if (want_rank) a.rank[row] = cnt;
  }
  my_loss_out = my_loss;
  cnt_out = cnt;
  active_out = active;
}

// Pull the logits of a LATER row block of this CTA from HBM into L2 while the current one is being
// processed (the body's own loads then cost an L2 hit): the row loop has no other overlap between blocks.
template <int TPR, int NE, bool VEC>
__device__ __forceinline__ void prefetch_row_block(const RowArgs& a, int64_t row_block) {
  if constexpr (VEC) {
    constexpr int THREADS = TPR > 256 ? TPR : 256;
    const int t = threadIdx.x % TPR;
    const int64_t row = row_block * (THREADS / TPR) + threadIdx.x / TPR;
    if (row < a.B) {
      const float* zr = a.z + row * a.ldz;
#pragma unroll
      for (int q = 0; q < NE / 4; ++q) {
        const int col = (q * TPR + t) * 4;
        if (col < a.C && (col & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(zr + col));   // one per 128-byte line
      }
    }
  }
}

// The rows `row_block * (THREADS/TPR) ..` of one CTA.  Results for the row of this thread's group come
// back in (my_loss, cnt, active); only the t == 0 thread of a row group needs them.
// `hook()` runs (every thread) once the row's loads have returned -- the loss-fused GEMM launch uses it to
// start its operand TMA stream without queueing the row's own loads behind it.
template <int TPR, int NE, bool VEC, int MODE, class Hook = NoHook>
__device__ __forceinline__ void softmax_row_body(const RowArgs& a, int64_t row_block,
                                                 RowSmem<(TPR > 256 ? TPR : 256)>& sm, float& my_loss_out, int& cnt_out,
                                                 bool& active_out, Hook hook = Hook()) {
  if constexpr (VEC) {
    softmax_row_body_vec<TPR, NE, MODE, Hook>(a, row_block, sm, my_loss_out, cnt_out, active_out, hook);
    return;
  }
  constexpr int THREADS = TPR > 256 ? TPR : 256;
  constexpr int WPR = TPR / 32;            // warps per row
  auto& s_f = sm.f;
  auto& s_i = sm.i;
  const int t = threadIdx.x % TPR;
  const int lrow = threadIdx.x / TPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = (warp / WPR) * WPR;      // first warp of this row
  const int64_t row = row_block * (THREADS / TPR) + lrow;
  const bool active = row < a.B;
  const int C = a.C;
  const float* zr = a.z + (active ? row : 0) * a.ldz;
  const bool want_rank = a.rank != nullptr, want_arg = a.argmax != nullptr;

  int64_t y = -1;
  if (active && a.label) y = __ldg(a.label + row);
  const bool y_in = active && y >= 0 && y < C;
  const bool y_ok = y_in && y != a.ignore_index;
  const int yi = y_in ? (int)y : -1;
  // label-dependent scalars: issued now, consumed after the reductions.  (The label's IIF weight is NOT
  // fetched with a dependent load: the thread that owns the label's column already has it -- see `ay`.)
  float g = 0.f;
  if (MODE == 0 && y_ok) {
    g = a.scale;
    if (a.cw) g *= __ldg(a.cw + y);
    if (a.sw) g *= __ldg(a.sw + row);
  }

  float v[NE];                             // raw logits, later exp(scaled - max)
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < NE / 4; ++q) {
      const int col = (q * TPR + t) * 4;
      float4 z4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
      if (active && col < C) z4 = ldg_stream4(zr + col);
      v[4 * q] = z4.x; v[4 * q + 1] = z4.y; v[4 * q + 2] = z4.z; v[4 * q + 3] = z4.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int col = e * TPR + t;
      v[e] = (active && col < C) ? __ldg(zr + col) : -CUDART_INF_F;
    }
  }
  auto col_of = [&](int e) { return VEC ? ((e >> 2) * TPR + t) * 4 + (e & 3) : e * TPR + t; };
  auto scale_of = [&](int e) -> float {
    const int col = col_of(e);
    return (a.iif && col < C) ? __ldg(a.iif + col) : 1.f;
  };

  // ---- pass 1: row max of the scaled logits, arg max, the label's raw logit
  float m = -CUDART_INF_F, bv = -CUDART_INF_F, zy = -CUDART_INF_F;
  float ay = 0.f;                          // adjusted logit of the label: summed over the row (all other terms +0)
  int bi = 0x7fffffff;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int col = col_of(e);
    const bool in = active && col < C;
    const float z = v[e];
    const float sc = in ? z * scale_of(e) : -CUDART_INF_F;
    m = fmaxf(m, sc);
    if (want_arg) {
      const float cmp = a.on_scaled ? sc : z;
      if (in && cmp > bv) { bv = cmp; bi = col; }
    }
    if (in && col == yi) { zy = z; ay = sc; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    zy = fmaxf(zy, __shfl_xor_sync(0xffffffffu, zy, o));
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    if (want_arg) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
  }
  if constexpr (WPR > 1) {
    if (lane == 0) { s_f[0][warp] = m; s_f[1][warp] = zy; s_f[2][warp] = bv; s_f[3][warp] = ay; s_i[warp] = bi; }
    __syncthreads();
    hook();
    m = s_f[0][w0]; zy = s_f[1][w0]; bv = s_f[2][w0]; ay = s_f[3][w0]; bi = s_i[w0];
#pragma unroll
    for (int w = 1; w < WPR; ++w) {
      m = fmaxf(m, s_f[0][w0 + w]);
      zy = fmaxf(zy, s_f[1][w0 + w]);
      ay += s_f[3][w0 + w];
      const float ov = s_f[2][w0 + w]; const int oi = s_i[w0 + w];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
  }
  if constexpr (WPR == 1) hook();
  if (yi < 0) { zy = 0.f; ay = 0.f; }

  if (MODE == 1 && !a.softmax) {
    // out = z * iif  (cls/custom.py:38)
    float* o = a.out + (active ? row : 0) * a.ldo;
#pragma unroll
    for (int e = 0; e < NE; ++e) v[e] *= scale_of(e);
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
  }

  // ---- pass 2: exp-sum; rank of the label (needs the label's logit, now known)
  const float mm = (m == -CUDART_INF_F) ? 0.f : m;
  const float ref = a.on_scaled ? ay : zy;
  float sum = 0.f;
  int cnt = 0;
  const bool need_exp = !(MODE == 1 && !a.softmax);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int col = col_of(e);
    const bool in = active && col < C;
    const float z = (MODE == 1 && !a.softmax) ? v[e] : v[e];   // activation mode already holds z*iif
    const float sc = (MODE == 1 && !a.softmax) ? z : (in ? z * scale_of(e) : -CUDART_INF_F);
    if (want_rank) {
      const float cmp = (a.on_scaled || (MODE == 1 && !a.softmax)) ? sc : z;
      cnt += in && ((cmp > ref) || (cmp == ref && col < yi));
    }
    if (need_exp) {
      const float ex = in ? expf(sc - mm) : 0.f;
      v[e] = ex;
      sum += ex;
    }
  }
  sum = warp_sum(sum);
  if (want_rank) cnt = warp_sum_i(cnt);
  if constexpr (WPR > 1) {
    if (lane == 0) { s_f[0][warp] = sum; s_i[warp] = cnt; }
    __syncthreads();
    sum = s_f[0][w0]; cnt = s_i[w0];
#pragma unroll
    for (int w = 1; w < WPR; ++w) { sum += s_f[0][w0 + w]; cnt += s_i[w0 + w]; }
  }
  if (yi < 0) cnt = C;  // label outside [0,C): never inside any top-k

  float my_loss = 0.f;
  if (need_exp) {
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
      my_loss = y_ok ? g * (lse - ay) : 0.f;
      if (active && t == 0) {
        if (a.loss_i) a.loss_i[row] = my_loss;
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
              float4 d;
              d.x = scale_of(4 * q + 0) * g * (v[4 * q + 0] * inv - (col + 0 == yi ? 1.f : 0.f));
              d.y = scale_of(4 * q + 1) * g * (v[4 * q + 1] * inv - (col + 1 == yi ? 1.f : 0.f));
              d.z = scale_of(4 * q + 2) * g * (v[4 * q + 2] * inv - (col + 2 == yi ? 1.f : 0.f));
              d.w = scale_of(4 * q + 3) * g * (v[4 * q + 3] * inv - (col + 3 == yi ? 1.f : 0.f));
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
              float d = scale_of(e) * g * (v[e] * inv - (col == yi ? 1.f : 0.f));
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
    if (want_arg) a.argmax[row] = bi;
    if (want_rank) a.rank[row] = cnt;
  }
  my_loss_out = my_loss;
  cnt_out = cnt;
  active_out = active;
}

}  // namespace iif
