// The whole head step in ONE persistent launch (round 2): fc_cls forward -> IIF softmax-CE rows -> dX, dW, db.
//
//   phase F   Z partials:  every CTA runs its share of the (tile, K-split) items of  X W^T  on the tensor cores
//             (tcgen05.mma, accumulator in TMEM, operands by TMA) and PARKS the fp32 partial tile in an L2-resident
//             workspace: TMEM -> registers -> 128B-swizzled staging -> cp.async.bulk.tensor store; one release
//             increment of the m-tile's arrival counter per item.  No rendezvous, no reduce pass.
//   phase L   loss rows:   every CTA takes rows of the batch; a row SUMS the split-K partials of its 128-row tile in
//             split order (deterministic) + bias while it loads them -- that sum is the logit row Z (written out for
//             the caller) -- and runs the IIF softmax-CE row body of loss_row.cuh on it: loss_i, dZ (bf16), argmax,
//             rank.  The only wait is on the counters of the m-tiles the CTA's rows live in.
//   barrier   one counter: all dZ rows are in L2 (dW needs every row).  Only the TMA-producer thread waits; the
//             X / W operand tiles of the CTA's first backward item were requested before the rows ran.
//   phase B   backward items:  dW tiles (unsplit, TMA store, db from a ones-tile MMA) and dX (tile, K-split) items
//             whose partial tiles are parked like phase F's.
//   phase R   every CTA reduces a slice of the dX partial tiles in split order and writes dX (bf16 / fp32).
//   tail      deterministic loss sum / top-k counts by the last CTA (ticket), which also re-arms every counter.
//
// One CTA per SM, grid <= #SMs, launched COOPERATIVELY: the driver guarantees co-residency, so the in-kernel
// counters cannot dead-lock against kernels of other streams (the data-parallel all-reduce).  Items of a phase are
// dealt to CTAs by a fixed function of blockIdx (snake order: longest items first), so every role thread can
// enumerate them without communication.  Inside a phase the TMA producer runs one item ahead of the MMA issuer.
//
// Reference lines this launch replaces: cls/train.py:66-77 (output = model.fc(x); loss = criterion(output, y);
// loss.backward()) and seg/mmdet/models/roi_heads/bbox_heads/bbox_head.py:118,269-274 + autograd.
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "loss_row.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

#ifndef HF_MIN_BLOCKS
#define HF_MIN_BLOCKS 1
#endif
// Register cap of the step kernel: 168 x 256 threads = 42 K of the SM's 64 K registers, so that a 256-thread CTA of the
// data-parallel all-reduce kernel (<= 80 registers per thread) can be co-resident with a step CTA -- without that room
// the all-reduce of step k could only run in the gaps between step launches instead of under them.
#ifndef HF_MAX_REGS
#define HF_MAX_REGS 168
#endif

namespace iif {
namespace hf {

constexpr int TM = 128, TN = 128, TK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * TK * 2, B_BYTES = TN * TK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_BYTES = 8 * 8192;        // 8 warps x (32 rows x 64 fp32 columns)
constexpr int ONES_BYTES = 2048;
constexpr int BAR_BYTES = 128;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + ONES_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 256;                 // 128 accumulator columns + 16 for db (power of two)
constexpr int TILE_F4 = TM * TN / 4;
constexpr int MAX_SPLITS = 8;
// Counters (int32) in the workspace header, bytes [8192, 32768).  Word 0 = parity; two counter SETS follow: launch k
// works on set (parity) and its last CTA to arrive at the grid barrier zeroes the other set -- used by launch k - 1,
// complete by then -- and flips the parity for launch k + 1.  No launch ever resets a counter another CTA may still
// poll, so nothing at the end of the kernel waits for "everybody is done".
constexpr int CTR_BYTE_OFFSET = 8192;
constexpr int CTR_SET0 = 64, CTR_SET_STRIDE = 2560;
constexpr int CTR_ROWS = 0, CTR_READY_F = 8, CTR_DX = 128;
constexpr int MAX_MT = 96, MAX_DX_TILES = 2048;
constexpr int WS_HEADER_BYTES = 32768;
static_assert(CTR_BYTE_OFFSET + (CTR_SET0 + 2 * CTR_SET_STRIDE) * 4 <= WS_HEADER_BYTES, "counter sets must fit the header");
static_assert(CTR_DX + MAX_DX_TILES <= CTR_SET_STRIDE && CTR_READY_F + MAX_MT <= CTR_DX, "counter set layout");

struct Gemm { int M, N, K, tiles_m, tiles_n, kb_total, splits, items; };

struct Args {
  Gemm f, dx, dw;
  int b_dx_first;            // phase B order: dX items before dW items (the longer K loop goes first)
  int row_blocks;
  float4* part;              // partial tiles [item][128 rows][128 cols] fp32 (phase F, then reused by dX)
  int* ctr;
  const float* bias;
  RowArgs loss;
  void* dx_out; int dx_bf16; int64_t lddx;
  float* db;
  long long* dbg;
  int flags;                 // debug switches (env IIF_B200_FUSED_FLAGS): 1 = writer-side proxy fences as in round 1
};

struct Item {
  int kind;                  // 0 = forward partial, 1 = dX partial, 2 = dW tile
  int split, nsplit, cnt;    // k-blocks split, split + nsplit, ...: cnt of them
  int m0, n0;
  int part_row;              // park items: first row of the partial tile in the partial tensor
  int ctr_idx;               // park items: arrival counter
  int do_db;
};

__device__ __forceinline__ void stamp(const Args& g, int slot) {
  if (g.dbg) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.dbg[(int64_t)blockIdx.x * 32 + slot] = t;
  }
}

__device__ __forceinline__ Item decode_f(const Args& g, int i) {
  Item it;
  const int S = g.f.splits, tile = i / S;
  it.kind = 0; it.split = i - tile * S; it.nsplit = S;
  it.cnt = (g.f.kb_total - it.split + S - 1) / S;
  const int mi = tile / g.f.tiles_n;
  it.m0 = mi * TM; it.n0 = (tile - mi * g.f.tiles_n) * TN;
  it.part_row = i * TM; it.ctr_idx = CTR_READY_F + mi; it.do_db = 0;
  return it;
}
__device__ __forceinline__ Item decode_b(const Args& g, int t) {
  Item it;
  const bool is_dx = g.b_dx_first ? t < g.dx.items : t >= g.dw.items;
  const int j = g.b_dx_first ? (is_dx ? t : t - g.dx.items) : (is_dx ? t - g.dw.items : t);
  if (is_dx) {
    const int S = g.dx.splits, tile = j / S;
    it.kind = 1; it.split = j - tile * S; it.nsplit = S;
    it.cnt = (g.dx.kb_total - it.split + S - 1) / S;
    const int mi = tile / g.dx.tiles_n;
    it.m0 = mi * TM; it.n0 = (tile - mi * g.dx.tiles_n) * TN;
    it.part_row = j * TM; it.ctr_idx = CTR_DX + tile; it.do_db = 0;
  } else {
    it.kind = 2; it.split = 0; it.nsplit = 1; it.cnt = g.dw.kb_total;
    const int mi = j / g.dw.tiles_n;
    it.m0 = mi * TM; it.n0 = (j - mi * g.dw.tiles_n) * TN;
    it.part_row = 0; it.ctr_idx = 0; it.do_db = (g.db != nullptr && it.n0 == 0) ? 1 : 0;
  }
  return it;
}
// k-th item of this CTA in a phase with n items: rounds alternate direction (snake), so the CTA that got the
// longest item of one round gets the shortest of the next.  Returns -1 past the end.
__device__ __forceinline__ int snake(int k, int n) {
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const int t = (k & 1) ? (k + 1) * G - 1 - c : k * G + c;
  return t < n ? t : -1;
}
__device__ __forceinline__ int my_items(int n) {   // how many items of an n-item phase this CTA runs
  int k = 0;
  // rounds are full except the last; a CTA's items are those k with snake(k, n) >= 0 (at most one gap-free prefix
  // plus possibly nothing: snake indices grow with k, so the first miss ends the list)
  while (snake(k, n) >= 0) ++k;
  return k;
}

// The raw logits of this thread's share of a row: sum of the forward partial tiles in split order (deterministic)
// + bias; also written out as Z.  All loads of a group of QG float4 columns are issued before the first one is
// consumed (one L2 round trip per group instead of one per float4); columns at or beyond C come back as -inf.
template <int TPR, int NE>
__device__ __forceinline__ void load_bias(const Args& g, bool active, float4 (&b4)[NE / 4]) {
  const int t = threadIdx.x % TPR, C = g.loss.C;
#pragma unroll
  for (int q = 0; q < NE / 4; ++q) {
    const int col = (q * TPR + t) * 4;
    b4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.bias && active && col < C) {
      if (col + 4 <= C) b4[q] = __ldg(reinterpret_cast<const float4*>(g.bias + col));
      else {
        b4[q].x = __ldg(g.bias + col);
        if (col + 1 < C) b4[q].y = __ldg(g.bias + col + 1);
        if (col + 2 < C) b4[q].z = __ldg(g.bias + col + 2);
      }
    }
  }
}
template <int TPR, int NE>
__device__ __forceinline__ void load_logits(const Args& g, int64_t row, bool active, const float4 (&b4)[NE / 4],
                                            float4 (&z4)[NE / 4]) {
  constexpr int NQ = NE / 4;
  constexpr int QG = NQ < 2 ? NQ : 2;                // float4 columns per batch: QG x MAX_SPLITS loads in flight
  constexpr int SMAX = MAX_SPLITS;
  const int t = threadIdx.x % TPR, C = g.loss.C, S = g.f.splits;
  const int mi = (int)(row >> 7), r = (int)(row & 127);
  float* zrow = g.loss.z ? const_cast<float*>(g.loss.z) + row * g.loss.ldz : nullptr;
  const bool zvec = (g.loss.ldz & 3) == 0 && (reinterpret_cast<uintptr_t>(g.loss.z) & 15u) == 0;
#pragma unroll
  for (int q0 = 0; q0 < NQ; q0 += QG) {
    float4 tt[QG][SMAX];
#pragma unroll
    for (int u = 0; u < QG; ++u) {
      const int col = ((q0 + u) * TPR + t) * 4;
      if (active && col < C) {
        const float4* p = g.part + ((int64_t)(mi * g.f.tiles_n + (col >> 7)) * S * TM + r) * (TN / 4) + ((col & 127) >> 2);
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s < S) tt[u][s] = __ldcg(p + (int64_t)s * TILE_F4);
      }
    }
#pragma unroll
    for (int u = 0; u < QG; ++u) {
      const int q = q0 + u, col = (q * TPR + t) * 4;
      if (!(active && col < C)) {
        z4[q] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        continue;
      }
      float4 acc = tt[u][0];
#pragma unroll
      for (int s = 1; s < SMAX; ++s)                 // fixed split order: deterministic
        if (s < S) { acc.x += tt[u][s].x; acc.y += tt[u][s].y; acc.z += tt[u][s].z; acc.w += tt[u][s].w; }
      acc.x += b4[q].x; acc.y += b4[q].y; acc.z += b4[q].z; acc.w += b4[q].w;
      const bool full = col + 4 <= C;
      if (zrow) {
        if (full && zvec) stg_stream4(zrow + col, acc);
        else {
          zrow[col] = acc.x;
          if (col + 1 < C) zrow[col + 1] = acc.y;
          if (col + 2 < C) zrow[col + 2] = acc.z;
          if (col + 3 < C) zrow[col + 3] = acc.w;
        }
      }
      if (!full) {                                   // ragged last group: columns >= C do not exist
        if (col + 1 >= C) acc.y = -CUDART_INF_F;
        if (col + 2 >= C) acc.z = -CUDART_INF_F;
        acc.w = -CUDART_INF_F;
      }
      z4[q] = acc;
    }
  }
}

// dX = sum over splits of the parked partial tiles: U chunks of 256 float4 (8 rows of one tile) per pass.
// The chunks are dealt to the `workers` CTAs that ran FEWER backward items than the busiest ones (rank `wrank` among
// them; everybody when the items divide evenly): the busiest CTAs are the launch's critical path and go straight to
// the exit, the others would idle.  Static and deterministic -- no work-stealing atomics.
template <int U, int SMAX>
__device__ __forceinline__ void reduce_dx(const Args& g, const int* ctrs, int wrank, int workers) {
  const int S = g.dx.splits;
  const int nchunks = g.dx.tiles_m * g.dx.tiles_n * 16;
  const int G = workers;
  for (int c0 = wrank; c0 < nchunks; c0 += G * U) {
    if ((int)threadIdx.x < U) {
      const int c = c0 + (int)threadIdx.x * G;
      if (c < nchunks) ptx::spin_until_ge(ctrs + CTR_DX + (c >> 4), S);
    }
    __syncthreads();
    if (threadIdx.x == 0 && c0 == wrank) stamp(g, 14);
    float4 t[U][SMAX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * G;
      if (c < nchunks) {
        const int tile = c >> 4, sub = c & 15;
        const float4* p = g.part + (int64_t)tile * S * TILE_F4 + sub * 256 + threadIdx.x;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s < S) t[u][s] = __ldcg(p + (int64_t)s * TILE_F4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * G;
      if (c >= nchunks) continue;
      float4 acc = t[u][0];
#pragma unroll
      for (int s = 1; s < SMAX; ++s)
        if (s < S) { acc.x += t[u][s].x; acc.y += t[u][s].y; acc.z += t[u][s].z; acc.w += t[u][s].w; }
      const int tile = c >> 4, sub = c & 15;
      const int mi = tile / g.dx.tiles_n, ni = tile - mi * g.dx.tiles_n;
      const int m = mi * TM + sub * 8 + ((int)threadIdx.x >> 5), n = ni * TN + 4 * ((int)threadIdx.x & 31);
      if (m >= g.dx.M || n >= g.dx.N) continue;
      const bool full = n + 4 <= g.dx.N;
      if (g.dx_bf16) {
        uint16_t* o = reinterpret_cast<uint16_t*>(g.dx_out) + (int64_t)m * g.lddx + n;
        if (full && (g.lddx & 3) == 0 && (reinterpret_cast<uintptr_t>(g.dx_out) & 7u) == 0)
          stg_stream2(o, pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
        else {
          const float r[4] = {acc.x, acc.y, acc.z, acc.w};
          for (int e = 0; e < 4; ++e) if (n + e < g.dx.N) o[e] = bf16_bits(r[e]);
        }
      } else {
        float* o = reinterpret_cast<float*>(g.dx_out) + (int64_t)m * g.lddx + n;
        if (full && (g.lddx & 3) == 0 && (reinterpret_cast<uintptr_t>(g.dx_out) & 15u) == 0) stg_stream4(o, acc);
        else {
          const float r[4] = {acc.x, acc.y, acc.z, acc.w};
          for (int e = 0; e < 4; ++e) if (n + e < g.dx.N) o[e] = r[e];
        }
      }
    }
    __syncthreads();
  }
}

// TPR threads per loss row, NE logits per thread (C <= TPR * NE), 256 / TPR rows per CTA pass.  REGS: register cap
// (HF_MAX_REGS, or 128 for data-parallel runs: IIF_HEAD_LOW_REGS leaves room for two all-reduce lanes on every SM).
template <int TPR, int NE, int REGS>
__global__ void __maxnreg__(REGS)
head_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmP,
                  const __grid_constant__ CUtensorMap tmDW, const __grid_constant__ Args g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;    // SWIZZLE_128B tiles: 1024-byte aligned
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t staging_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t ones_base = staging_base + STAGING_BYTES;
  const uint32_t bar_base = ones_base + ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(g, 0);
  __shared__ volatile int s_last;                     // grid barrier: 0 = not known yet, 1 = not the last arriver, 2 = the last
  if (threadIdx.x == 96) s_last = 0;                  // (ordered before every later use by the prologue's block barrier)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmW); ptx::prefetch_tensormap(&tmDZ);
    ptx::prefetch_tensormap(&tmP); ptx::prefetch_tensormap(&tmDW);
  }
  if (warp == 1 && lane == 0) {
    // a full barrier takes TWO producer arrivals per phase (B part, A part): the B operand of the first backward
    // item is requested before the loss rows run, its A operand (dZ) only after the grid barrier
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar(s), 2); ptx::mbar_init(empty_bar(s), 1); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp == 3) {                                   // constant B operand of the bias-gradient MMA
    uint4* o = reinterpret_cast<uint4*>(smem_gen + (ones_base - smem_base));
    const uint4 one = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = lane; i < ONES_BYTES / 16; i += 32) o[i] = one;
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (threadIdx.x == 0) stamp(g, 1);
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();                               // the previous launch of the stream (same workspace) is complete
  if (threadIdx.x == 0) stamp(g, 2);
  // The single producer thread runs scalar code at a few cycles per instruction: only the first forward item is
  // decoded in front of the first TMA requests, everything else behind them.
  const int nF = my_items(g.f.items);
  const Item itF0 = decode_f(g, nF > 0 ? snake(0, g.f.items) : 0);

  constexpr int NQ = NE / 4;
  constexpr int RPB = 256 / TPR;
  int parity = 0;
  int* ctrs = nullptr;                               // counter set of this launch: set after the first operand requests

  // ---- pipeline state.  Producer (thread 0): ring position of load #0 of the NEXT item to be (fully) issued, and
  // how many A / B parts of that item are already in flight.  MMA issuer (thread 32): its own ring position.
  int p_stage = 0; uint32_t p_phase = 0; int pre_a = 0, pre_b = 0;
  int m_stage = 0; uint32_t m_phase = 0;
  uint32_t acc_parity = 0;

  auto ring_at = [&](int i, int& stage, uint32_t& phase) {      // ring slot of load #i of the producer's current item
    const int s = p_stage + i;
    stage = s % STAGES;
    phase = p_phase ^ (uint32_t)((s / STAGES) & 1);
  };
  auto load_a = [&](const Item& it, int stage, int kb) {
    const uint32_t sa = smem_base + stage * STAGE_BYTES;
    ptx::mbar_arrive_expect_tx(full_bar(stage), A_BYTES);
    const int k0 = kb * TK;
    if (it.kind == 0) {            // X, K-major
      ptx::tma_load_2d(sa, &tmX, full_bar(stage), k0, it.m0);
      ptx::tma_load_2d(sa + 8192, &tmX, full_bar(stage), k0, it.m0 + 64);
    } else if (it.kind == 1) {     // dZ, K-major (K = classes)
      ptx::tma_load_2d(sa, &tmDZ, full_bar(stage), k0, it.m0);
      ptx::tma_load_2d(sa + 8192, &tmDZ, full_bar(stage), k0, it.m0 + 64);
    } else {                       // dZ^T, MN-major (M = classes contiguous, K = rows)
      ptx::tma_load_2d(sa, &tmDZ, full_bar(stage), it.m0, k0);
      ptx::tma_load_2d(sa + 8192, &tmDZ, full_bar(stage), it.m0 + 64, k0);
    }
  };
  auto load_b = [&](const Item& it, int stage, int kb) {
    const uint32_t sb = smem_base + stage * STAGE_BYTES + A_BYTES;
    ptx::mbar_arrive_expect_tx(full_bar(stage), B_BYTES);
    const int k0 = kb * TK;
    if (it.kind == 0) {            // W, K-major
      ptx::tma_load_2d(sb, &tmW, full_bar(stage), k0, it.n0);
      ptx::tma_load_2d(sb + 8192, &tmW, full_bar(stage), k0, it.n0 + 64);
    } else if (it.kind == 1) {     // W, MN-major (N = d contiguous, K = classes)
      ptx::tma_load_2d(sb, &tmW, full_bar(stage), it.n0, k0);
      ptx::tma_load_2d(sb + 8192, &tmW, full_bar(stage), it.n0 + 64, k0);
    } else {                       // X, MN-major (N = d contiguous, K = rows)
      ptx::tma_load_2d(sb, &tmX, full_bar(stage), it.n0, k0);
      ptx::tma_load_2d(sb + 8192, &tmX, full_bar(stage), it.n0 + 64, k0);
    }
  };
  // producer: bring the A / B parts of loads [.., upto) of the current item in flight (each ring slot is waited
  // for once: by whichever part is issued first)
  auto produce = [&](const Item& it, int upto, bool want_a, bool want_b) {
    if (upto > it.cnt) upto = it.cnt;
    const int lo = pre_a < pre_b ? pre_a : pre_b;
    for (int i = lo; i < upto; ++i) {
      int stage; uint32_t phase;
      ring_at(i, stage, phase);
      const bool need_a = want_a && i >= pre_a, need_b = want_b && i >= pre_b;
      if (!need_a && !need_b) continue;
      if (i >= pre_a && i >= pre_b) ptx::mbar_wait(empty_bar(stage), phase ^ 1u);   // first touch of this slot
      const int kb = it.split + i * it.nsplit;
      if (need_b) load_b(it, stage, kb);
      if (need_a) load_a(it, stage, kb);
    }
    if (want_a && upto > pre_a) pre_a = upto;
    if (want_b && upto > pre_b) pre_b = upto;
  };
  auto producer_next_item = [&](const Item& it) {    // the current item is fully issued: move to the next
    const int s = p_stage + it.cnt;
    p_phase ^= (uint32_t)((s / STAGES) & 1);
    p_stage = s % STAGES;
    pre_a = pre_b = 0;
  };
  auto mma_item = [&](const Item& it) {              // thread 32
    const bool a_mn = it.kind == 2, b_mn = it.kind != 0;
    const uint32_t idesc = ptx::make_idesc_bf16(TM, TN, a_mn, b_mn);
    const uint32_t idesc_db = ptx::make_idesc_bf16(TM, 16, a_mn, false);
    const uint32_t a_step = a_mn ? 2048u : 32u, a_lbo = a_mn ? 8192u : 16u;
    const uint32_t b_step = b_mn ? 2048u : 32u, b_lbo = b_mn ? 8192u : 16u;
    for (int i = 0; i < it.cnt; ++i) {
      ptx::mbar_wait(full_bar(m_stage), m_phase);
      ptx::tc_fence_after();
      if (g.dbg && (i == 0 || i == it.cnt - 1)) stamp(g, (it.kind == 0 ? 17 : (it.kind == 1 ? 19 : 21)) + (i == 0 ? 0 : 1));
      const uint32_t sa = smem_base + m_stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
      for (int k = 0; k < TK / 16; ++k) {
        const uint64_t da = ptx::make_smem_desc_sw128(sa + k * a_step, a_lbo, 1024);
        const uint64_t db = ptx::make_smem_desc_sw128(sb + k * b_step, b_lbo, 1024);
        const uint32_t accum = (i > 0 || k > 0) ? 1u : 0u;
        ptx::umma_bf16(tmem_base, da, db, idesc, accum);
        if (it.do_db)
          ptx::umma_bf16(tmem_base + TN, da, ptx::make_smem_desc_sw128(ones_base + k * 32, 16, 1024), idesc_db, accum);
      }
      ptx::umma_commit(empty_bar(m_stage));
      if (++m_stage == STAGES) { m_stage = 0; m_phase ^= 1u; }
    }
    ptx::umma_commit(tmem_full_bar);
  };

  // drain the accumulator of one item: registers -> swizzled staging -> TMA store (partial tensor or dW)
  const int q = warp & 3, h = warp >> 2;
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);
  const uint32_t wbase = staging_base + (uint32_t)warp * 8192u;
  auto drain_item = [&](const Item& it) {
    ptx::mbar_wait(tmem_full_bar, acc_parity);
    acc_parity ^= 1u;
    ptx::tc_fence_after();
    if (threadIdx.x == 0 && it.kind != 0) stamp(g, 15);      // (last backward item: accumulator complete)
    const uint32_t rbase = wbase + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        ptx::sts128(rbase + (uint32_t)c * 4096u + ((((uint32_t)j) ^ sw) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2],
                    r[4 * j + 3]);
    }
    ptx::fence_proxy_async();                        // generic-proxy smem writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
      if (it.kind == 2) {
        const int mr = it.m0 + q * 32, nc = it.n0 + h * 64;
        if (mr < g.dw.M) {
          if (nc < g.dw.N) ptx::tma_store_2d(&tmDW, wbase, nc, mr);
          if (nc + 32 < g.dw.N) ptx::tma_store_2d(&tmDW, wbase + 4096u, nc + 32, mr);
        }
        ptx::bulk_commit();
        ptx::bulk_wait_read0();                      // the staging slab is free again
      } else {
        ptx::tma_store_2d(&tmP, wbase, h * 64, it.part_row + q * 32);
        ptx::tma_store_2d(&tmP, wbase + 4096u, h * 64 + 32, it.part_row + q * 32);
        ptx::bulk_commit();
        // the partial rows are WRITTEN (not merely read out of shared memory) and visible to this thread, whose
        // release (below, after the block barrier) then publishes them: no proxy fence on the writer's side (a
        // generic `fence.proxy.async` is MEMBAR.ALL.GPU + an L1 invalidate in SASS)
        ptx::bulk_wait0();
        if (g.flags & 1) asm volatile("fence.proxy.async;" ::: "memory");
      }
    }
    if (it.do_db && h == 0) {                        // db of this class row: first of the 16 equal columns
      const uint32_t v = ptx::tmem_ld1(tmem_base + ((uint32_t)(q * 32) << 16) + TN);
      ptx::tmem_ld_wait();
      const int m = it.m0 + q * 32 + lane;
      if (m < g.dw.M) g.db[m] = __uint_as_float(v);
    }
    ptx::tc_fence_before();
    __syncthreads();                                 // accumulator + staging reusable; every warp's park is complete
    ptx::tc_fence_after();
    if (it.kind != 2 && threadIdx.x == 0)   // (counter set from the parity word: first use, long after its load was issued)
      ptx::red_release_add(g.ctr + CTR_SET0 + (parity & 1) * CTR_SET_STRIDE + it.ctr_idx, 1);
  };

  // ============================ phase F ============================
  if (threadIdx.x == 0 && nF > 0) {
    // first fill of a ring nobody has touched: no empty-slot waits
    const int n0 = itF0.cnt < STAGES ? itF0.cnt : STAGES;
    for (int i = 0; i < n0; ++i) {
      const int kb = itF0.split + i * itF0.nsplit;
      load_b(itF0, i, kb);
      load_a(itF0, i, kb);
    }
    pre_a = pre_b = n0;
    stamp(g, 16);
  }
  // Counter set of this launch (stable until the last CTA passes the grid barrier).  Loaded AFTER the first operand
  // requests: an in-order warp stalls at the first USE of a load, and that use must not sit in front of the TMA issue.
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(parity) : "l"(g.ctr));
  const int nB = g.dx.items + g.dw.items;
  const int my_b = my_items(nB);
  const Item itB0 = decode_b(g, my_b > 0 ? snake(0, nB) : 0);

  // The label-dependent scalars, IIF weights and bias of this CTA's FIRST loss-row block do not depend on the
  // forward product: their (HBM-cold) loads are issued now -- after the operand requests of the first forward
  // item, which must not queue behind the label -> class-weight dependency -- and land under phase F.
  constexpr bool PRE_BIAS = NE <= 8;                 // (wide rows: NQ float4 of bias held across phase F would spill)
  RowHead<TPR, NE> head0;
  float4 bias0[PRE_BIAS ? NQ : 1];
  auto prefetch_head = [&]() {
    if ((int)blockIdx.x < g.row_blocks) {
      row_head<TPR, NE, 0>(g.loss, blockIdx.x, head0);
      if constexpr (PRE_BIAS) load_bias<TPR, NE>(g, head0.active, bias0);
    }
  };
  // The two ROLE threads must not stall on these loads (an in-order thread waits at the first use of a loaded value,
  // and the label -> class-weight chain is two HBM round trips): they prefetch after their role work of the first
  // forward item is issued; everybody else -- with nothing to do until the accumulator is complete -- prefetches now.
  const bool role_thread = threadIdx.x == 0 || threadIdx.x == 32;
  if (!role_thread || nF == 0) prefetch_head();
  {
    const int n = nF;
    for (int k = 0; k < n; ++k) {
      const Item it = k == 0 ? itF0 : decode_f(g, snake(k, g.f.items));
      if (threadIdx.x == 0) {
        produce(it, it.cnt, true, true);
        producer_next_item(it);
        if (k + 1 < n) produce(decode_f(g, snake(k + 1, g.f.items)), STAGES, true, true);
        if (k == 0) { stamp(g, 3); prefetch_head(); }
      } else if (threadIdx.x == 32) {
        mma_item(it);
        if (k == 0) prefetch_head();
      }
      __syncwarp();
      drain_item(it);
      if (threadIdx.x == 0 && k == n - 1) stamp(g, 4);
    }
  }

  // first backward item of this CTA: its B operand (X / W tiles) does not depend on the loss -- requested from
  // inside the first loss row (or right here when the CTA has no rows)
  bool b_issued = false;
  auto issue_b = [&]() {
    if (threadIdx.x == 0 && !b_issued) {
      stamp(g, 12);                                  // first loss row: loads returned, first reduction done
      if (my_b > 0) produce(itB0, STAGES, false, true);
    }
    b_issued = true;
  };

  // ============================ phase L ============================
  double loss_part = 0.0;
  int loss_c1 = 0, loss_c5 = 0;
  {
    parity &= 1;
    ctrs = g.ctr + CTR_SET0 + parity * CTR_SET_STRIDE;
    __shared__ RowSmem<256> row_sm;
    __shared__ float s_row_loss[256 / TPR];
    __shared__ int s_row_rank[256 / TPR];
    const int target = g.f.tiles_n * g.f.splits;
    for (int rb = blockIdx.x; rb < g.row_blocks; rb += gridDim.x) {
      RowHead<TPR, NE> hd;
      float4 b4[NQ], z4[NQ];
      if (rb == (int)blockIdx.x) {
        hd = head0;
        if constexpr (PRE_BIAS) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) b4[q] = bias0[q];
        } else {
          load_bias<TPR, NE>(g, hd.active, b4);
        }
      } else {
        row_head<TPR, NE, 0>(g.loss, rb, hd);           // (in flight while thread 0 polls the m-tile's counter)
        load_bias<TPR, NE>(g, hd.active, b4);
      }
      if (threadIdx.x == 0) {
        const int mi = (rb * RPB) >> 7;               // the rows of one block share an m-tile (RPB divides 128)
        ptx::spin_until_ge(ctrs + CTR_READY_F + mi, target);
        if (rb == (int)blockIdx.x) stamp(g, 5);
      }
      __syncthreads();
      load_logits<TPR, NE>(g, hd.row, hd.active, b4, z4);
      float my_loss; int cnt; bool active;
      row_tail<TPR, NE, 0>(g.loss, hd, z4, row_sm, my_loss, cnt, active, issue_b);
      if (threadIdx.x == 0 && rb == (int)blockIdx.x) stamp(g, 13);
      const int t = threadIdx.x % TPR, lrow = threadIdx.x / TPR;
      if (t == 0) { s_row_loss[lrow] = active ? my_loss : 0.f; s_row_rank[lrow] = active ? cnt : 0x7fffffff; }
      __syncthreads();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
          loss_part += (double)s_row_loss[r];
          loss_c1 += s_row_rank[r] < 1; loss_c5 += s_row_rank[r] < 5;
        }
      }
    }
    issue_b();                                        // (a CTA without rows)
    // dZ: generic-proxy stores here, async-proxy (TMA) reads in other CTAs.  The block barrier orders every
    // thread's stores before thread 0's release; the READER fences the proxies after its acquire (phase B).
    if (g.flags & 1) asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
  }

  // ============================ grid barrier + phase B ============================
  {
    double* g_part = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(g.loss.scratch) + 16);
    int* g_c1 = reinterpret_cast<int*>(g_part + gridDim.x);
    int* g_c5 = g_c1 + gridDim.x;
    if (threadIdx.x == 0) {
      stamp(g, 6);
      __stcg(g_part + blockIdx.x, loss_part);         // this CTA's share of the loss / top-k counts: published by ...
      __stcg(g_c1 + blockIdx.x, loss_c1);
      __stcg(g_c5 + blockIdx.x, loss_c5);
      // ... the barrier arrival (release: our rows of dZ and these partials; acquire: everybody else's, for the poll)
      const int old = ptx::atom_add_acq_rel(ctrs + CTR_ROWS, 1);
      __threadfence_block();
      s_last = old == (int)gridDim.x - 1 ? 2 : 1;
      // grid barrier, producer thread only: every row of dZ is in L2 (the other threads go on to wait for the
      // accumulator of the first item)
      ptx::spin_until_ge(ctrs + CTR_ROWS, (int)gridDim.x);
      asm volatile("fence.proxy.async;" ::: "memory");
      stamp(g, 7);
      if (my_b > 0) produce(itB0, STAGES, true, true);
    }
    if (warp == 3) {
      // Duties of the LAST CTA to arrive (it knows every CTA's partials are published), done by an otherwise idle warp
      // while the first backward item's operands are in flight: the deterministic loss sum / top-k counts (fixed
      // lane-strided order + shuffle tree), zeroing the counter set of the NEXT launch, flipping the parity.
      int last = 0;
      if (lane == 0) { while ((last = s_last) == 0) { } }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last == 2) {
        __threadfence_block();
        double acc = 0.0;
        int k1 = 0, k5 = 0;
        for (unsigned i = lane; i < gridDim.x; i += 32) {
          acc += __ldcg(g_part + i); k1 += __ldcg(g_c1 + i); k5 += __ldcg(g_c5 + i);
        }
        acc = warp_sum_d(acc); k1 = warp_sum_i(k1); k5 = warp_sum_i(k5);
        if (lane == 0) {
          if (g.loss.loss_sum) *g.loss.loss_sum = (float)acc;
          if (g.loss.acc_counts) { g.loss.acc_counts[0] = k1; g.loss.acc_counts[1] = k5; }
        }
        int* other = g.ctr + CTR_SET0 + (parity ^ 1) * CTR_SET_STRIDE;
        if (lane == 0) other[CTR_ROWS] = 0;
        for (int i = lane; i < g.f.tiles_m; i += 32) other[CTR_READY_F + i] = 0;
        const int ndx = g.dx.items > 0 ? g.dx.tiles_m * g.dx.tiles_n : 0;
        for (int i = lane; i < ndx; i += 32) other[CTR_DX + i] = 0;
        if (lane == 0) g.ctr[0] = parity ^ 1;
      }
    }
    for (int k = 0; k < my_b; ++k) {
      const Item it = k == 0 ? itB0 : decode_b(g, snake(k, nB));
      if (threadIdx.x == 0) {
        produce(it, it.cnt, true, true);
        producer_next_item(it);
        if (k + 1 < my_b) produce(decode_b(g, snake(k + 1, nB)), STAGES, true, true);
        if (k == 0) stamp(g, 8);
      } else if (threadIdx.x == 32) {
        mma_item(it);
      }
      __syncwarp();
      drain_item(it);
      if (threadIdx.x == 0 && k == my_b - 1) stamp(g, 9);
    }
  }
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);   // (drain_item ended with a fenced block barrier)

  // ============================ phase R ============================
  if (g.dx.items > 0) {
    // workers: the CTAs with fewer than ceil(nB / grid) backward items (see reduce_dx)
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    const int max_b = (nB + G - 1) / G, rem = nB - (max_b - 1) * G;      // `rem` CTAs ran max_b items
    int wrank = c, workers = G;
    if (rem < G && 3 * (G - rem) >= G) {               // (a handful of idle CTAs cannot carry the whole reduce)
      workers = G - rem;
      if ((max_b - 1) & 1) wrank = c < G - rem ? c : -1;                  // odd last round: dealt from the top down
      else wrank = c >= rem ? c - rem : -1;
    }
    if (wrank >= 0) {
      if (g.dx.splits <= 2) reduce_dx<8, 2>(g, ctrs, wrank, workers);
      else if (g.dx.splits <= 4) reduce_dx<4, 4>(g, ctrs, wrank, workers);
      else reduce_dx<2, 8>(g, ctrs, wrank, workers);
    }
  }
  if (threadIdx.x == 0) { stamp(g, 10); stamp(g, 11); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Plan {
  Gemm f, dx, dw;
  int b_dx_first, grid, tpr, ne, row_blocks;
  size_t part_bytes;
};

static int snake_makespan(int G, int n_long, int len_long, int n_short, int len_short) {
  // items sorted longest first, dealt in snake order: the busiest CTA's total k-blocks
  int worst = 0;
  const int n = n_long + n_short;
  for (int c = 0; c < G; ++c) {
    int tot = 0;
    for (int k = 0;; ++k) {
      const int t = (k & 1) ? (k + 1) * G - 1 - c : k * G + c;
      if (t >= n) break;
      tot += t < n_long ? len_long : len_short;
    }
    worst = std::max(worst, tot);
  }
  return worst;
}

// Split counts from a small cost model (microseconds): ~0.4 us per 32 KB k-block a CTA streams, ~1.2 us per item for
// drain + park / store, and the partial tiles cost their bytes twice through L2 (~10 TB/s chip-wide).
static bool make_plan(int64_t B, int64_t D, int64_t C, bool need_dx, int sms, Plan* p) {
  if (B < 1 || D < 1 || C < 1 || C > 4096 || B > 2048) return false;   // larger batches: the multi-launch path (phases long enough)
  Plan pl{};
  pl.f.M = (int)B; pl.f.N = (int)C; pl.f.K = (int)D;
  pl.f.tiles_m = (int)((B + TM - 1) / TM); pl.f.tiles_n = (int)((C + TN - 1) / TN); pl.f.kb_total = (int)((D + TK - 1) / TK);
  pl.dw.M = (int)C; pl.dw.N = (int)D; pl.dw.K = (int)B;
  pl.dw.tiles_m = pl.f.tiles_n; pl.dw.tiles_n = (int)((D + TN - 1) / TN); pl.dw.kb_total = (int)((B + TK - 1) / TK);
  pl.dw.splits = 1; pl.dw.items = pl.dw.tiles_m * pl.dw.tiles_n;
  pl.dx = Gemm{};
  if (need_dx) {
    pl.dx.M = (int)B; pl.dx.N = (int)D; pl.dx.K = (int)C;
    pl.dx.tiles_m = pl.f.tiles_m; pl.dx.tiles_n = pl.dw.tiles_n; pl.dx.kb_total = (int)((C + TK - 1) / TK);
  }
  if (pl.f.tiles_m > MAX_MT) return false;
  const int G = sms;
  {  // forward split
    const int tiles = pl.f.tiles_m * pl.f.tiles_n;
    double best = 1e30; int best_s = 1;
    for (int s = 1; s <= MAX_SPLITS && s <= pl.f.kb_total; ++s) {
      const int per = (pl.f.kb_total + s - 1) / s, items = tiles * s, rounds = (items + G - 1) / G;
      const double cost = rounds * (per * 0.4 + 1.2) + 2.0 * items * 65536.0 / 10e6 + 0.05 * s;
      if (cost < best - 1e-9) { best = cost; best_s = s; }
    }
    if (const char* e = getenv("IIF_B200_FUSED_SF")) { const int v = atoi(e); if (v >= 1 && v <= MAX_SPLITS && v <= pl.f.kb_total) best_s = v; }
    pl.f.splits = best_s; pl.f.items = tiles * best_s;
  }
  if (need_dx) {  // dX split, given the dW items it shares the phase with
    const int tiles = pl.dx.tiles_m * pl.dx.tiles_n;
    if (tiles > MAX_DX_TILES) return false;
    double best = 1e30; int best_s = 1;
    for (int s = 1; s <= MAX_SPLITS && s <= pl.dx.kb_total; ++s) {
      const int per = (pl.dx.kb_total + s - 1) / s, items = tiles * s;
      const bool dx_first = per >= pl.dw.kb_total;
      const int span = dx_first ? snake_makespan(G, items, per, pl.dw.items, pl.dw.kb_total)
                                : snake_makespan(G, pl.dw.items, pl.dw.kb_total, items, per);
      const int rounds = (items + pl.dw.items + G - 1) / G;
      const double cost = span * 0.4 + rounds * 1.2 + 2.0 * items * 65536.0 / 10e6 + 0.05 * s;
      if (cost < best - 1e-9) { best = cost; best_s = s; }
    }
    if (const char* e = getenv("IIF_B200_FUSED_SDX")) { const int v = atoi(e); if (v >= 1 && v <= MAX_SPLITS && v <= pl.dx.kb_total) best_s = v; }
    pl.dx.splits = best_s; pl.dx.items = tiles * best_s;
    pl.b_dx_first = (pl.dx.kb_total + best_s - 1) / best_s >= pl.dw.kb_total;
  }
  // loss-row geometry: a pass of the row loop is latency (one L2 round trip, two block reductions), so batches with
  // many rows per CTA take MORE ROWS PER PASS (fewer threads per row, more logits per thread)
  // (measured: more rows per pass -- 64 threads x 32 logits -- is SLOWER for the LVIS shapes: the wide per-thread
  // state spills under the register cap; the three narrow geometries stay)
  if (C <= 1024) { pl.tpr = 128; pl.ne = 8; }
  else if (C <= 2048) { pl.tpr = 256; pl.ne = 8; }
  else { pl.tpr = 256; pl.ne = 16; }
  const int rpb = 256 / pl.tpr;
  pl.row_blocks = (int)((B + rpb - 1) / rpb);
  // Routing between this launch and the multi-launch chain (profiles/r2_route.jsonl, us/step one launch vs chain):
  // 256x2048x1000 23.6 / 29.2, 256x2048x365 19.2 / 29.4, 512x1024x1204 33.6 / 35.8, 512x2048x1000 33.5 / 30.5,
  // 1024x1024x1204 51.2 / 47.7, 1024x2048x1000 52.2 / 40.0, 2048x1024x1204 88.0 / 52.3, 2048x2048x1000 92.5 / 60.5.
  // The single launch saves the fixed costs between launches (~6 us) but runs one CTA per SM -- no neighbour CTA whose
  // MMAs cover a drain -- and each CTA walks its loss rows one pass after another (~3 us a pass), where the chain's
  // loss launch keeps several CTAs per SM.  So it is taken up to B*D*C = 0.8e9 multiply-adds per GEMM and 4 row passes.
  // IIF_B200_FUSED_MAX_WORK (multiply-adds, 0 = no limit) and IIF_B200_FUSED_MAX_ROW_PASSES (0 = no limit) override.
  {
    int max_passes = 4;
    double max_work = 0.8e9;
    if (const char* e = getenv("IIF_B200_FUSED_MAX_ROW_PASSES")) max_passes = atoi(e);
    if (const char* e = getenv("IIF_B200_FUSED_MAX_WORK")) max_work = atof(e);
    if (max_passes > 0 && (pl.row_blocks + G - 1) / G > max_passes) return false;
    if (max_work > 0 && (double)B * (double)D * (double)C > max_work) return false;
  }
  const int most = std::max(std::max(pl.f.items, pl.dx.items + pl.dw.items), pl.row_blocks);
  pl.grid = std::min(G, most);
  pl.part_bytes = (size_t)std::max(pl.f.items, pl.dx.items) * TM * TN * 4;
  if (pl.part_bytes > ((size_t)1 << 30)) return false;
  *p = pl;
  return true;
}

static long long* g_dbg = nullptr;

template <int TPR, int NE, int REGS = HF_MAX_REGS>
static const void* kernel_ptr() { return reinterpret_cast<const void*>(&head_fused_kernel<TPR, NE, REGS>); }

static const void* pick_kernel(const Plan& p, bool low_regs) {
  if (low_regs) {
    if (p.tpr == 128) return kernel_ptr<128, 8, 128>();
    if (p.ne == 8) return kernel_ptr<256, 8, 128>();
    return kernel_ptr<256, 16, 128>();
  }
  if (p.tpr == 128) return kernel_ptr<128, 8>();
  if (p.ne == 8) return kernel_ptr<256, 8>();
  return kernel_ptr<256, 16>();
}

static int device_sms() {
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!sms[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    for (const void* fn : {kernel_ptr<128, 8>(), kernel_ptr<256, 8>(), kernel_ptr<256, 16>(), kernel_ptr<128, 8, 128>(),
                           kernel_ptr<256, 8, 128>(), kernel_ptr<256, 16, 128>()}) {
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return 0;
      cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    sms[dev] = v;
  }
  return sms[dev];
}

}  // namespace hf
}  // namespace iif

extern "C" void iif_debug_timing_fused(long long* buf) { iif::hf::g_dbg = buf; }

// The plan of the one-launch step for a shape on a `sms`-SM device (host only; no device needed):
// out = {grid, fwd splits, fwd items, dX splits, dX items, dW items, dX-first, threads per row, logits per thread,
// row blocks, partial MB (rounded up), 0}.  Returns 0, or IIF_EUNSUPPORTED when the shape does not qualify.
extern "C" int iif_debug_fused_plan(int64_t B, int64_t D, int64_t C, int need_dx, int sms, int* out12) {
  iif::hf::Plan p;
  if (!out12 || sms < 1) return IIF_EINVAL;
  if (!iif::hf::make_plan(B, D, C, need_dx != 0, sms, &p)) return IIF_EUNSUPPORTED;
  const int v[12] = {p.grid, p.f.splits, p.f.items, p.dx.splits, p.dx.items, p.dw.items, p.b_dx_first, p.tpr, p.ne,
                     p.row_blocks, (int)((p.part_bytes + 1048575) >> 20), 0};
  for (int i = 0; i < 12; ++i) out12[i] = v[i];
  return IIF_OK;
}

namespace iif {

// bytes of workspace the fused step needs for this shape (0 = shape not supported by the fused step)
size_t head_fused_ws_bytes(int64_t B, int64_t D, int64_t C) {
  hf::Plan p;
  if (!hf::make_plan(B, D, C, true, kNumSMs, &p)) return 0;
  return (size_t)hf::WS_HEADER_BYTES + p.part_bytes;
}

// The whole step in one launch.  IIF_EUNSUPPORTED (nothing launched) when the shape / alignment does not qualify.
int head_fused_launch(const iif_head_args* h, void* stream, bool dry_run) {
  if (!h || !h->x || !h->w || !h->label || !h->dz_bf16 || !h->dw) return IIF_EINVAL;
  const int64_t B = h->B, D = h->D, C = h->C;
  if (B <= 0 || D <= 0 || C <= 0) return IIF_EUNSUPPORTED;
  if (h->ldx < D || h->ldw < D || h->lddw < D || h->lddz < C || (h->z && h->ldz < C) || (h->dx && h->lddx < D)) return IIF_EINVAL;
  if (h->ldx % 8 || h->ldw % 8 || h->lddz % 8 || (h->lddw * 4) % 16 || !aligned16(h->x) || !aligned16(h->w) ||
      !aligned16(h->dz_bf16) || !aligned16(h->dw) || (h->iif && !aligned16(h->iif)) || (h->bias && !aligned16(h->bias)))
    return IIF_EUNSUPPORTED;
  if (!h->scratch) return (h->loss_sum || h->acc_counts) ? IIF_EINVAL : IIF_EUNSUPPORTED;   // the CTAs' loss partials live there
  if (h->acc_counts && !h->rank) return IIF_EINVAL;
  // the ragged last float4 group of a row writes its dZ padding columns: the row pitch must hold them
  if (h->lddz < (C + 3) / 4 * 4) return IIF_EUNSUPPORTED;
  const int sms = hf::device_sms();
  if (sms <= 0) { cudaGetLastError(); return IIF_EDRIVER; }
  hf::Plan p;
  if (!hf::make_plan(B, D, C, h->dx != nullptr, sms, &p)) return IIF_EUNSUPPORTED;
  const size_t need = (size_t)hf::WS_HEADER_BYTES + p.part_bytes;
  if (!h->ws || h->ws_bytes < need || !aligned16(h->ws)) return IIF_EUNSUPPORTED;
  if (dry_run) return IIF_OK;

  hf::Args g{};
  g.f = p.f; g.dx = p.dx; g.dw = p.dw; g.b_dx_first = p.b_dx_first; g.row_blocks = p.row_blocks;
  uint8_t* ws = reinterpret_cast<uint8_t*>(h->ws);
  g.ctr = reinterpret_cast<int*>(ws + hf::CTR_BYTE_OFFSET);
  g.part = reinterpret_cast<float4*>(ws + hf::WS_HEADER_BYTES);
  g.bias = h->bias;
  make_ce_row_args(g.loss, h->z, h->ldz, h->iif, h->label, h->class_weight, h->sample_weight, h->ignore_index, h->scale,
                   B, C, h->loss_i, h->loss_sum, nullptr, 0, h->dz_bf16, h->lddz, nullptr, h->argmax, h->rank,
                   h->acc_counts, h->scratch);
  g.dx_out = h->dx; g.dx_bf16 = h->dx_dtype == IIF_DTYPE_BF16; g.lddx = h->lddx;
  g.db = h->db;
  g.dbg = hf::g_dbg;
  {
    static const int flags = [] { const char* e = getenv("IIF_B200_FUSED_FLAGS"); return e ? atoi(e) : 0; }();
    g.flags = flags;
  }
  CUtensorMap mx, mw, mdz, mp, mdw;
  int rc;
  if ((rc = make_map(&mx, h->x, false, (uint64_t)D, (uint64_t)B, (uint64_t)h->ldx, 64, 64))) return rc;
  if ((rc = make_map(&mw, h->w, false, (uint64_t)D, (uint64_t)C, (uint64_t)h->ldw, 64, 64))) return rc;
  if ((rc = make_map(&mdz, h->dz_bf16, false, (uint64_t)C, (uint64_t)B, (uint64_t)h->lddz, 64, 64))) return rc;
  const uint64_t prows = (uint64_t)std::max(p.f.items, std::max(p.dx.items, 1)) * hf::TM;
  if ((rc = make_map(&mp, g.part, true, (uint64_t)hf::TN, prows, (uint64_t)hf::TN, 32, 32))) return rc;
  if ((rc = make_map(&mdw, h->dw, true, (uint64_t)D, (uint64_t)C, (uint64_t)h->lddw, 32, 32))) return rc;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)p.grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = hf::SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  void* kargs[6] = {&mx, &mw, &mdz, &mp, &mdw, &g};
  const cudaError_t e = launch_cooperative(cfg, hf::pick_kernel(p, (h->flags & IIF_HEAD_LOW_REGS) != 0), kargs, true);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return e == cudaSuccess ? IIF_OK : (int)e;
}

}  // namespace iif
