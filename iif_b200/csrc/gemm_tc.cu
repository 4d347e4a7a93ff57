// (a)/(c) fc_cls GEMMs on the 5th-generation tensor cores: tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA through an mbarrier ring, one elected thread
// issuing the MMAs, an 8-warp TMEM drain and a TMA-store epilogue.
//
//   OUT[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias[n]),  out2 = OUT * col_scale[n]
//
// bf16 operands, fp32 accumulation.  Either operand may be K-major (K contiguous in HBM) or
// MN-major (M/N contiguous): the three products of the head use the SAME row-major tensors
//   fwd (a1) : A = X[B,D]  K-major      B = W[C,D]  K-major          Z  = X W^T + b
//   dX  (a10): A = dZ[B,C] K-major      B = W[C,D]  MN-major (n = d) dX = dZ W
//   dW  (a10): A = dZ[B,C] MN-major (m = c)  B = X[B,D] MN-major (n = d)  dW = dZ^T X
// so no transposed copy of W, X or dZ is ever written to HBM.
//
// One launch runs a GROUP of up to two problems (dX and dW share a launch: they depend on the same
// dZ and together fill the 148 SMs); blockIdx.x -> (problem, tile, K split).
//
// Tile: 128 x 128 x 64 per CTA, 128-byte swizzle, ring of 3 x 32 KB stages => ~100 KB of shared
// memory and 256 TMEM columns per CTA, TWO CTAs per SM (296 resident CTAs): the drain / store of one
// CTA overlaps the TMA + MMA stream of its neighbour, and a programmatically launched successor
// kernel finds room for its prologue.
//
// Split-K for the small head shapes (ImageNet-LT: 256 x 1000 x 2048 has only 16 output tiles): the
// CTAs of one tile park their fp32 partial tiles in an L2-resident workspace (register -> global,
// in a layout that is coalesced for both the writer and the reader), meet at a per-tile arrival
// counter (release / acquire at gpu scope) and then EVERY CTA reduces 1/splits of the tile's rows
// in split order: parallel, deterministic, no float atomics, no serial tail.  The counters are
// self-resetting; the host only enables split-K when the whole grid is co-resident.
// (Round-1 first version used a thread-block cluster of 8 for this; 8-CTA clusters schedule at most
// ~14 clusters at once on B200's GPCs, so the 16-tile forward ran as two waves -- see profiles/.)
//
// Epilogue, unsplit tiles: each warp moves its 32 x 64 accumulator slab TMEM -> registers -> (alpha,
// bias) -> a 128B-swizzled box in recycled stage memory -> one TMA store per 32 x 32 fp32 box
// (bank-conflict-free, full 128-byte lines, tile tails clipped by the TMA unit).  Outputs the TMA
// unit cannot address (row pitch not a multiple of 16 bytes, e.g. Places-LT C = 365, or the second
// IIF-scaled output) go through a warp-private transposing staging slab and coalesced row stores.
//
// Bias gradient on the tensor cores: in the dW product the CTAs of the first tile column issue one
// extra N=16 MMA per k-step against a constant tile of ones, so db[c] = sum_b dZ[b,c] * 1 falls out
// of the same operand stream into 16 spare TMEM columns -- no column-sum kernel, no extra HBM read.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = ones tile + bias/scale staging; then ALL 8 warps drain (warp w owns TMEM lanes
// 32*(w%4).. and accumulator columns 64*(w/4)..).
//
// Every kernel begins with griddepcontrol.launch_dependents / .wait (programmatic dependent
// launch): the prologue (barrier init, TMEM alloc, tensor-map prefetch) overlaps the tail of the
// previous kernel of the step.
#include <cuda.h>

#include <mutex>
#include <unordered_map>
#include <utility>

#include "common.cuh"
#include "loss_row.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

namespace iif {

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int BN = 128;
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * TILE_K * 2;      // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int MAX_STAGES = 6;          // 1 CTA / SM (grids that fit one wave of 148 CTAs): deeper TMA ring
constexpr int MIN_STAGES = 3;          // 2 CTAs / SM
constexpr int MAX_SPLITS = 8;
constexpr int ONES_BYTES = 2048;      // 16 rows x 128 B of bf16 1.0 (K-major B operand of the db MMA)
constexpr int TMEM_COLS = 256;        // BN accumulator columns + 16 for db (power of two)
constexpr int BAR_BYTES = 256;
__host__ __device__ constexpr int smem_bytes_for(int stages) {
  return stages * STAGE_BYTES + ONES_BYTES + BAR_BYTES + 2 * BN * 4 + 1024;
}
constexpr int SMEM_BYTES = smem_bytes_for(MIN_STAGES);       // the 2-CTAs-per-SM configuration
constexpr int GEN_PITCH = 65;         // generic epilogue: warp-private 32 x 65 float slab
constexpr int WS_HEADER = 32768;      // [0, 4096) split-K counters: 2 problems x 256 tiles; grid barrier at 4096; [8192, 32768) counters of head_fused.cu
constexpr int TILE_F4 = TILE_M * BN / 4;
// Split-K arrival counters never reset: every CTA of a tile adds EPOCH_UNIT / splits, so each launch
// advances the tile's counter by exactly EPOCH_UNIT whatever its split count (840 = lcm(1..8)); the
// value an arriver gets back tells it which multiple to wait for.  64-bit: no wrap in practice.
constexpr unsigned long long EPOCH_UNIT = 840;
static_assert(8 * 32 * GEN_PITCH * 4 <= MIN_STAGES * STAGE_BYTES, "generic staging must fit in the stage ring");
static_assert(TILE_M * BN * 4 <= MIN_STAGES * STAGE_BYTES, "TMA-store staging must fit in the stage ring");

struct TcProblem {
  int M, N, K;
  int tiles_m, tiles_n, kb_total, kb_per_split, splits;
  int a_mn, b_mn;
  const float* alpha; const float* bias; const float* col_scale;
  void* out; int out_bf16; int64_t ldo; int epi_tma;
  int ksplit_add;                                   // > 1: K split over CTAs that ADD their tiles into a zeroed output
  float* out2; int64_t ldo2;
  float4* partial;                                  // [tiles][splits][TILE_F4] fp32x4, layout part_idx()
  unsigned long long* counters;                     // [tiles] arrival epochs (see EPOCH_UNIT)
  float* db_out; float* db_partial;                 // dW only: db[m] (and [tiles_m][splits][TILE_M] partials)
};

struct TcGroup {
  int nprob, stages;
  int cta_begin[3];
  long long* dbg;                        // optional per-CTA phase timestamps (iif_debug_timing)
  TcProblem p[2];
  // loss-fused backward launch: every CTA first runs rows of the IIF softmax-CE (loss_row.cuh) that
  // PRODUCES the A operand (dZ), the grid meets at `grid_bar`, then the GEMMs consume dZ from L2.
  int fuse_loss, loss_ne;
  int* grid_bar;                         // arrival count, zero between launches
  RowArgs loss;
};

__device__ __forceinline__ void stamp(const TcGroup& g, int slot) {
  if (g.dbg) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.dbg[(int64_t)blockIdx.x * 16 + slot] = t;
  }
}

// Partial-tile layout (float4 units): 8-row groups, inside a group 32 column-chunks x 8 rows.
//  - drain (lane = row, fixed chunk j): 8 lanes x 16 B contiguous -> four full 128 B lines per warp store
//  - reduce (consecutive threads = consecutive float4): fully coalesced; a warp then owns 8 rows x 4
//    chunks = 64 contiguous output bytes per row
__device__ __forceinline__ int part_idx(int j, int row) { return (((row >> 3) * 32 + j) << 3) + (row & 7); }

// One float4 of the final output: alpha, bias, store (fp32 / bf16), optional scaled copy.  Bounds-checked.
__device__ __forceinline__ void emit4(const TcProblem& P, int m, int n, float4 v, float alpha, const float* s_bias,
                                      const float* s_scale, int nl) {
  if (m >= P.M || n >= P.N) return;
  float r[4] = {v.x * alpha + s_bias[nl], v.y * alpha + s_bias[nl + 1], v.z * alpha + s_bias[nl + 2],
                v.w * alpha + s_bias[nl + 3]};
  const bool full = n + 4 <= P.N;
  if (P.out) {
    if (P.out_bf16) {
      uint16_t* o = reinterpret_cast<uint16_t*>(P.out) + (int64_t)m * P.ldo + n;
      if (full && (P.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out) & 7u) == 0)
        stg_stream2(o, pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]));
      else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = bf16_bits(r[j]);
    } else {
      float* o = reinterpret_cast<float*>(P.out) + (int64_t)m * P.ldo + n;
      if (full && (P.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out) & 15u) == 0)
        stg_stream4(o, make_float4(r[0], r[1], r[2], r[3]));
      else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = r[j];
    }
  }
  if (P.out2) {
    float* o = P.out2 + (int64_t)m * P.ldo2 + n;
    for (int j = 0; j < 4; ++j) r[j] *= s_scale[nl + j];
    if (full && (P.ldo2 & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out2) & 15u) == 0)
      stg_stream4(o, make_float4(r[0], r[1], r[2], r[3]));
    else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = r[j];
  }
}

// FUSE_NE: 0 = plain GEMM launch; 4 / 8 / 16 = loss-fused backward launch with that many logits per thread
// of a 256-thread row (one instantiation each keeps the straight-line code small: these kernels live for
// ~10 us, instruction-cache misses are visible in their profile).
template <int FUSE_NE>
__global__ void __launch_bounds__(256, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
               const __grid_constant__ TcGroup g) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int STAGES = g.stages;
  const uint32_t ones_base = smem_base + STAGES * STAGE_BYTES;      // 1024-byte aligned
  const uint32_t bar_base = ones_base + ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* s_bias = reinterpret_cast<float*>(smem_gen + (bar_base + BAR_BYTES - smem_base));
  float* s_scale = s_bias + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(g, 0);
  const int pi = (g.nprob > 1 && (int)blockIdx.x >= g.cta_begin[1]) ? 1 : 0;
  const TcProblem& P = g.p[pi];
  const CUtensorMap* tmA = pi ? &tmA1 : &tmA0;
  const CUtensorMap* tmB = pi ? &tmB1 : &tmB0;
  const CUtensorMap* tmO = pi ? &tmO1 : &tmO0;
  const int local = (int)blockIdx.x - g.cta_begin[pi];
  // K splits: either the rendezvous form (P.splits: partial tiles + in-kernel deterministic reduction) or, for
  // long K with too many CTAs to be co-resident, the additive form (P.ksplit_add: TMA reduce-add into a zeroed
  // fp32 output -- no rendezvous, summation order across splits not fixed)
  const int ksp = P.ksplit_add > 1 ? P.ksplit_add : P.splits;
  const int split = local % ksp;
  const int tile = local / ksp;
  const int n0 = (tile % P.tiles_n) * BN, m0 = (tile / P.tiles_n) * TILE_M;
  // K blocks of a split tile are dealt round-robin: at any moment the `splits` CTAs of a tile stream ADJACENT
  // 128-byte chunks of the same operand rows (splits x 128 B contiguous per row) instead of chunks a quarter
  // of a row apart (measured neutral on B200 for the cold-HBM head shapes; kept because it makes every split
  // non-empty for any split count).  Local index i -> block kb_of(i).
  const int kb_begin = 0;
  const int kb_end = (P.kb_total - split + ksp - 1) / ksp;                    // blocks split, split+S, split+2S, ...
  auto kb_of = [&](int i) { return split + i * ksp; };
  const bool do_db = P.db_out != nullptr && n0 == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(tmA);
    ptx::prefetch_tensormap(tmB);
    if (P.epi_tma) ptx::prefetch_tensormap(tmO);
  }
  if (warp == 1 && lane == 0) {
    // full barriers take TWO producer arrivals per phase (B part, A part): the B operand of a loss-fused
    // launch is requested before the loss rows run, the A operand (dZ) only after the grid barrier
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar(s), 2); ptx::mbar_init(empty_bar(s), 1); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && do_db) {                // constant B operand of the bias-gradient MMA
    uint4* o = reinterpret_cast<uint4*>(smem_gen + (ones_base - smem_base));
    const uint4 one = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = lane; i < ONES_BYTES / 16; i += 32) o[i] = one;
    ptx::fence_proxy_async();              // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;

  if (threadIdx.x == 0) stamp(g, 1);
  const int n_first = min(STAGES, kb_end - kb_begin);       // k-blocks that fit the ring without recycling
  auto load_b = [&](int stage, int kb) {
    const uint32_t sb = smem_base + stage * STAGE_BYTES + A_STAGE_BYTES;
    ptx::mbar_arrive_expect_tx(full_bar(stage), B_STAGE_BYTES);
    const int k0 = kb * TILE_K;
    if (P.b_mn) {
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * 8192, tmB, full_bar(stage), n0 + 64 * j, k0);
    } else {
      ptx::tma_load_2d(sb, tmB, full_bar(stage), k0, n0);
    }
  };
  auto load_a = [&](int stage, int kb) {
    const uint32_t sa = smem_base + stage * STAGE_BYTES;
    ptx::mbar_arrive_expect_tx(full_bar(stage), A_STAGE_BYTES);
    const int k0 = kb * TILE_K;
    if (P.a_mn) {
#pragma unroll
      for (int j = 0; j < TILE_M / 64; ++j) ptx::tma_load_2d(sa + j * 8192, tmA, full_bar(stage), m0 + 64 * j, k0);
    } else {
      ptx::tma_load_2d(sa, tmA, full_bar(stage), k0, m0);
    }
  };
  // (Requesting x / w tiles BEFORE the dependency wait was tried and removed: with a predecessor kernel
  // still running, TMA loads issued ahead of griddepcontrol.wait were occasionally never delivered -- see
  // DESIGN.md "what did not work".)
  const int early_b = FUSE_NE ? n_first : 0;
  ptx::griddep_launch_dependents();      // the next kernel may start its own prologue now
  ptx::griddep_wait();                   // ... and ours ends here: the producer kernel's data is visible
  if (threadIdx.x == 0) stamp(g, 2);

  double loss_part = 0.0;
  int loss_c1 = 0, loss_c5 = 0;
  if constexpr (FUSE_NE != 0) {
    // ---- loss-fused launch: every CTA computes its share of the loss rows: Z -> loss_i, dZ (bf16, global).
    // The B operands (X / W tiles: independent of the loss) are requested from inside the first row, as soon
    // as that row's own loads have returned: before the rows they delay the rows' loads (12 MB of TMA
    // traffic ahead of them), after the rows they delay the grid barrier's polls -- both measured.
    __shared__ RowSmem<256> row_sm;
    bool b_issued = false;
    auto issue_b = [&]() {
      if (threadIdx.x == 0 && !b_issued)
        for (int i = 0; i < n_first; ++i) load_b(i, kb_of(i));
      b_issued = true;
    };
    for (int64_t r = blockIdx.x; r < g.loss.B; r += gridDim.x) {
      float my_loss; int cnt; bool active;
      if (r + gridDim.x < g.loss.B) prefetch_row_block<256, (FUSE_NE ? FUSE_NE : 4), true>(g.loss, r + gridDim.x);
      softmax_row_body<256, (FUSE_NE ? FUSE_NE : 4), true, 0>(g.loss, r, row_sm, my_loss, cnt, active, issue_b);
      if (threadIdx.x == 0) { loss_part += (double)my_loss; loss_c1 += cnt < 1; loss_c5 += cnt < 5; }
      __syncthreads();
    }
    issue_b();                                         // (a CTA without rows)
    asm volatile("fence.proxy.async;" ::: "memory");   // our dZ stores (generic proxy) vs. the TMA reads to come
    __syncthreads();
    if (threadIdx.x == 0) {
      stamp(g, 10);
      // release (cumulative over the CTA's stores ordered by the barrier above) / acquire at gpu scope:
      // no full __threadfence (MEMBAR.SC + L1 invalidate) on this path
      ptx::red_release_add(g.grid_bar, 1);
      ptx::spin_until_ge(g.grid_bar, (int)gridDim.x);   // every row of dZ is in L2
      asm volatile("fence.proxy.async;" ::: "memory");
      stamp(g, 11);
    }
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        if (kb - kb_begin >= early_b) load_b(stage, kb_of(kb));
        load_a(stage, kb_of(kb));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      stamp(g, 3);                       // all TMA loads issued
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(TILE_M, BN, P.a_mn != 0, P.b_mn != 0);
      const uint32_t idesc_db = ptx::make_idesc_bf16(TILE_M, 16, P.a_mn != 0, false);
      // K-major: 16 bf16 = 32 bytes along the swizzle row; 8-row groups 1024 B apart (SBO).
      // MN-major: 16 k-rows = 2048 bytes; 8-row groups 1024 B apart (SBO); 64-wide MN atoms 8192 B apart (LBO).
      const uint32_t a_step = P.a_mn ? 2048u : 32u, a_lbo = P.a_mn ? 8192u : 16u;
      const uint32_t b_step = P.b_mn ? 2048u : 32u, b_lbo = P.b_mn ? 8192u : 16u;
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        if (kb == kb_begin) stamp(g, 4);   // first stage landed
        if (kb == kb_end - 1) stamp(g, 5); // last stage landed
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < TILE_K / 16; ++k) {
          const uint64_t da = ptx::make_smem_desc_sw128(sa + k * a_step, a_lbo, 1024);
          const uint64_t db = ptx::make_smem_desc_sw128(sb + k * b_step, b_lbo, 1024);
          const uint32_t accum = (kb > kb_begin || k > 0) ? 1u : 0u;
          ptx::umma_bf16(tmem_base, da, db, idesc, accum);
          if (do_db)   // ones tile: every value equal, so the swizzle is immaterial
            ptx::umma_bf16(tmem_base + BN, da, ptx::make_smem_desc_sw128(ones_base + k * 32, 16, 1024), idesc_db, accum);
        }
        ptx::umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::umma_commit(tmem_full_bar);       // accumulator complete
    }
  } else if (warp == 3) {
    // epilogue vectors of this tile's 128 columns (read after the dependency wait: parameters)
    for (int i = lane; i < BN; i += 32) {
      const int n = n0 + i;
      s_bias[i] = (P.bias && n < P.N && !(P.ksplit_add > 1 && split != 0)) ? __ldg(P.bias + n) : 0.f;
      s_scale[i] = (P.col_scale && n < P.N) ? __ldg(P.col_scale + n) : 0.f;
    }
  }
  const float alpha = P.alpha ? __ldg(P.alpha) : 1.f;
  __syncthreads();                           // roles issued; s_bias / s_scale visible

  // ===================== drain: all 8 warps =====================
  const int q = warp & 3, h = warp >> 2;     // TMEM lane quarter, accumulator column half
  const int row = q * 32 + lane;             // row inside the tile
  ptx::mbar_wait(tmem_full_bar, 0);
  ptx::tc_fence_after();
  if (threadIdx.x == 0) stamp(g, 6);         // accumulator complete: the stage ring is free for staging
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);

  if (P.splits == 1) {
    if (P.epi_tma) {
      // ---- registers -> swizzled box -> TMA store (per warp; no block barrier)
      const uint32_t wbase = smem_base + (uint32_t)warp * (P.out_bf16 ? 4096u : 8192u);
      const uint32_t rbase = wbase + (uint32_t)lane * 128u;
      const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        const float* bs = s_bias + h * 64 + c * 32;
        if (P.out_bf16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {      // 8 values -> one 16-byte chunk
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              pk[e] = pack_bf16x2(__uint_as_float(r[8 * j + 2 * e]) * alpha + bs[8 * j + 2 * e],
                                  __uint_as_float(r[8 * j + 2 * e + 1]) * alpha + bs[8 * j + 2 * e + 1]);
            ptx::sts128(rbase + ((((uint32_t)(c * 4 + j)) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * j);
            ptx::sts128(rbase + (uint32_t)c * 4096u + ((((uint32_t)j) ^ sw) << 4),
                        __float_as_uint(__uint_as_float(r[4 * j]) * alpha + b4.x),
                        __float_as_uint(__uint_as_float(r[4 * j + 1]) * alpha + b4.y),
                        __float_as_uint(__uint_as_float(r[4 * j + 2]) * alpha + b4.z),
                        __float_as_uint(__uint_as_float(r[4 * j + 3]) * alpha + b4.w));
          }
        }
      }
      ptx::fence_proxy_async();              // generic-proxy smem writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        const int mr = m0 + q * 32, nc = n0 + h * 64;
        if (mr < P.M) {
          if (P.ksplit_add > 1) {                       // fp32 only (host guarantees)
            if (nc < P.N) ptx::tma_reduce_add_2d(tmO, wbase, nc, mr);
            if (nc + 32 < P.N) ptx::tma_reduce_add_2d(tmO, wbase + 4096u, nc + 32, mr);
          } else if (P.out_bf16) {
            if (nc < P.N) ptx::tma_store_2d(tmO, wbase, nc, mr);
          } else {
            if (nc < P.N) ptx::tma_store_2d(tmO, wbase, nc, mr);
            if (nc + 32 < P.N) ptx::tma_store_2d(tmO, wbase + 4096u, nc + 32, mr);
          }
        }
        ptx::bulk_commit();
      }
    } else {
      // ---- generic: warp-private transposing slab, then coalesced (bounds-checked) row stores
      float* slab = reinterpret_cast<float*>(smem_gen) + warp * (32 * GEN_PITCH);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) slab[lane * GEN_PITCH + c * 32 + j] = __uint_as_float(r[j]);
      }
      __syncwarp();
      for (int rr = 0; rr < 32; ++rr) {
        const int m = m0 + q * 32 + rr;
        if (m >= P.M) break;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nl = h * 64 + c * 32 + lane, n = n0 + nl;
          if (n < P.N) {
            const float v = slab[rr * GEN_PITCH + c * 32 + lane] * alpha + s_bias[nl];
            if (P.out) {
              if (P.out_bf16) reinterpret_cast<uint16_t*>(P.out)[(int64_t)m * P.ldo + n] = bf16_bits(v);
              else reinterpret_cast<float*>(P.out)[(int64_t)m * P.ldo + n] = v;
            }
            if (P.out2) P.out2[(int64_t)m * P.ldo2 + n] = v * s_scale[nl];
          }
        }
      }
    }
    if (do_db && h == 0) {                   // db of this row: first of the 16 equal columns
      const uint32_t v = ptx::tmem_ld1(tmem_base + ((uint32_t)(q * 32) << 16) + BN);
      ptx::tmem_ld_wait();
      if (m0 + row < P.M) {
        if (P.ksplit_add > 1) atomicAdd(P.db_out + m0 + row, __uint_as_float(v) * alpha);
        else P.db_out[m0 + row] = __uint_as_float(v) * alpha;
      }
    }
  } else {
    // ---- split-K: partial tile -> L2 workspace, meet at the tile's counter, reduce 1/splits of the rows
    float4* base = P.partial + (int64_t)tile * P.splits * TILE_F4;
    float4* mine = base + (int64_t)split * TILE_F4;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(taddr + c * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        __stcg(mine + part_idx(h * 16 + c * 8 + j, row),
               make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                           __uint_as_float(r[4 * j + 3])));
    }
    float* dbp = do_db ? P.db_partial + (int64_t)(tile / P.tiles_n) * P.splits * TILE_M : nullptr;
    if (do_db && h == 0) {
      const uint32_t v = ptx::tmem_ld1(tmem_base + ((uint32_t)(q * 32) << 16) + BN);
      ptx::tmem_ld_wait();
      __stcg(dbp + split * TILE_M + row, __uint_as_float(v));
    }
    unsigned long long* arrive = P.counters + tile;
    __syncthreads();                         // every thread's partial stores are issued ...
    if (threadIdx.x == 0) {
      stamp(g, 7);
      // ... and published by a gpu-scope RELEASE (cumulative over the CTA's stores ordered by the barrier);
      // the matching ACQUIRE is the poll.  No __threadfence (MEMBAR.SC + L1 invalidate) on this path.
      const unsigned long long inc = EPOCH_UNIT / (unsigned)P.splits;
      const unsigned long long old = ptx::atom_add_acq_rel_u64(arrive, inc);
      const unsigned long long target = (old / EPOCH_UNIT + 1) * EPOCH_UNIT;
      if (old + inc < target) ptx::spin_until_ge_u64(arrive, target);
      stamp(g, 8);
    }
    __syncthreads();
    const int gps = (TILE_M / 8 + P.splits - 1) / P.splits;          // 8-row groups per split
    const int g0 = split * gps, g1 = min(TILE_M / 8, g0 + gps);
    for (int grp = g0; grp < g1; grp += 2) {                          // two groups per pass: 2 x splits loads in flight
      const bool two = grp + 1 < g1;
      const int idx = grp * 256 + threadIdx.x;
      float4 t[2][MAX_SPLITS];
#pragma unroll
      for (int s = 0; s < MAX_SPLITS; ++s)
        if (s < P.splits) {
          t[0][s] = __ldcg(base + (int64_t)s * TILE_F4 + idx);
          if (two) t[1][s] = __ldcg(base + (int64_t)s * TILE_F4 + idx + 256);
        }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        float4 acc = t[u][0];
#pragma unroll
        for (int s = 1; s < MAX_SPLITS; ++s)   // fixed split order: deterministic sum
          if (s < P.splits) { acc.x += t[u][s].x; acc.y += t[u][s].y; acc.z += t[u][s].z; acc.w += t[u][s].w; }
        const int j = threadIdx.x >> 3;
        emit4(P, m0 + (grp + u) * 8 + (threadIdx.x & 7), n0 + 4 * j, acc, alpha, s_bias, s_scale, 4 * j);
      }
    }
    if (do_db) {
      const int r0 = g0 * 8, r1 = g1 * 8;
      if ((int)threadIdx.x < r1 - r0 && m0 + r0 + (int)threadIdx.x < P.M) {
        float acc = 0.f;
        for (int s = 0; s < P.splits; ++s) acc += __ldcg(dbp + s * TILE_M + r0 + threadIdx.x);
        P.db_out[m0 + r0 + threadIdx.x] = acc * alpha;
      }
    }
  }
  ptx::tc_fence_before();
  if (P.splits == 1 && P.epi_tma && lane == 0) ptx::bulk_wait_read0();   // staging must outlive the bulk reads
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) stamp(g, 9);
  if constexpr (FUSE_NE != 0) {
    // off the GEMMs' critical path: deterministic loss sum / top-k counts, then re-arm the grid barrier
    // (the ticket of the tail doubles as the "everyone is past the grid barrier" count: its last CTA re-arms it)
    if (grid_tail<256>(loss_part, loss_c1, loss_c5, g.loss.loss_sum, g.loss.acc_counts, g.loss.scratch) && threadIdx.x == 0)
      *g.grid_bar = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Resident-CTA capacity of the device for this kernel (2 per SM on B200): the split-K rendezvous
// needs every CTA of the grid on an SM at the same time.
static const void* kernel_variant(int v) {
  switch (v) {
    case 1: return reinterpret_cast<const void*>(&gemm_tc_kernel<4>);
    case 2: return reinterpret_cast<const void*>(&gemm_tc_kernel<8>);
    case 3: return reinterpret_cast<const void*>(&gemm_tc_kernel<16>);
    default: return reinterpret_cast<const void*>(&gemm_tc_kernel<0>);
  }
}

static int resident_capacity(int* detail = nullptr) {
  static std::mutex mu;
  static int caps[64] = {};              // per device ordinal; 0 = not yet queried
  static int details[64][6] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> g(mu);
  if (caps[dev] == 0) {
    int sms = 0, per_api = 0, smem_sm = 0, regs_sm = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    int worst_regs = 0, worst_static = 0;
    for (int v = 0; v < 4; ++v) {
      const void* fn = kernel_variant(v);
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(MAX_STAGES)) != cudaSuccess)
        return 0;
      // two ~100 KB CTAs per SM need the full shared-memory carve-out (the default sizes it for one)
      cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncAttributes fv{};
      if (cudaFuncGetAttributes(&fv, fn) != cudaSuccess) return 0;
      if (fv.numRegs > worst_regs) worst_regs = fv.numRegs;
      if ((int)fv.sharedSizeBytes > worst_static) worst_static = (int)fv.sharedSizeBytes;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_api, kernel_variant(0), 256, SMEM_BYTES);
    // Own bound from the hardware limits (shared memory incl. 1 KB/CTA reserved, registers in 8-register
    // warp granules, 512 TMEM columns); the runtime's occupancy answer is taken when it is larger.
    cudaFuncAttributes fa{};
    fa.numRegs = worst_regs;
    fa.sharedSizeBytes = (size_t)worst_static;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    const int by_smem = smem_sm / (SMEM_BYTES + 1024 + (int)fa.sharedSizeBytes);
    const int by_regs = regs_sm / (((fa.numRegs + 7) / 8 * 8) * 256);
    int per = by_smem < by_regs ? by_smem : by_regs;
    if (per > 512 / TMEM_COLS) per = 512 / TMEM_COLS;
    if (per_api > per) per = per_api;
    if (per < 1) per = 1;
    caps[dev] = sms * per;
    const int d[6] = {per_api, by_smem, by_regs, fa.numRegs, smem_sm, (int)fa.sharedSizeBytes};
    for (int i = 0; i < 6; ++i) details[dev][i] = d[i];
    cudaGetLastError();
  }
  if (detail) for (int i = 0; i < 6; ++i) detail[i] = details[dev][i];
  return caps[dev];
}

struct Plan {
  int tiles_m, tiles_n, kb_total, kb_per_split, splits;
  int ksplit_add = 1;
  int ctas() const { return tiles_m * tiles_n * (ksplit_add > 1 ? ksplit_add : splits); }
  size_t partial_bytes() const { return splits > 1 ? (size_t)tiles_m * tiles_n * splits * TILE_M * BN * 4 : 0; }
  size_t db_bytes() const { return splits > 1 ? (size_t)tiles_m * splits * TILE_M * 4 : 0; }   // 512-byte multiples
};

// K splits.  Cost model in microseconds: a CTA pays ~2.5 us of latency (first TMA round trip, drain,
// store), ~0.33 us per 32 KB k-block (doubled when two CTAs share an SM's L2 port), and a split tile
// ~1 us + 0.05 us per split for the exchange.  Split-K is only legal while the whole grid is
// co-resident (`cap` CTAs) and the per-problem counters fit the workspace header.
static Plan make_plan(int64_t M, int64_t N, int64_t K, int other_ctas, int cap) {
  Plan p;
  p.tiles_m = (int)((M + TILE_M - 1) / TILE_M);
  p.tiles_n = (int)((N + BN - 1) / BN);
  p.kb_total = (int)((K + TILE_K - 1) / TILE_K);
  if (p.kb_total < 1) p.kb_total = 1;
  const int tiles = p.tiles_m * p.tiles_n;
  double best = 1e30;
  int best_s = 1;
  for (int s = 1; s <= MAX_SPLITS && s <= p.kb_total; ++s) {
    const int per = (p.kb_total + s - 1) / s;
    if ((p.kb_total + per - 1) / per != s) continue;            // s must be reachable exactly
    const int ctas = tiles * s + other_ctas;
    if (s > 1 && (ctas > cap || tiles > 256)) break;
    const int waves = cap > 0 ? (ctas + cap - 1) / cap : 1;
    const double share = ctas > kNumSMs ? (ctas < 2 * kNumSMs ? (double)ctas / kNumSMs : 2.0) : 1.0;
    const double cost = waves * (2.5 + per * 0.33 * share + (s > 1 ? 1.0 + 0.05 * s : 0.0));
    if (cost < best - 1e-9) { best = cost; best_s = s; }
  }
  p.splits = best_s;
  p.kb_per_split = (p.kb_total + best_s - 1) / best_s;
  return p;
}

struct GemmDesc {
  const void* A; int64_t lda; bool a_mn;
  const void* Bm; int64_t ldb; bool b_mn;
  int64_t M, N, K;
  const float* alpha; const float* bias; const float* col_scale;
  void* out; int out_bf16; int64_t ldo;
  float* out2; int64_t ldo2;
  float* db_out;
};

static int fill_problem(const GemmDesc& d, const Plan& p, float* partial, unsigned long long* counters, float* db_partial,
                        TcProblem* P,
                        CUtensorMap* ma, CUtensorMap* mb, CUtensorMap* mo) {
  if (!aligned16(d.A) || !aligned16(d.Bm) || d.lda % 8 || d.ldb % 8) return IIF_EALIGN;
  int rc;
  // K-major: memory [MN rows, K cols]; MN-major: memory [K rows, MN cols]
  rc = d.a_mn ? make_map(ma, d.A, false, (uint64_t)d.M, (uint64_t)d.K, (uint64_t)d.lda, 64, 64)
              : make_map(ma, d.A, false, (uint64_t)d.K, (uint64_t)d.M, (uint64_t)d.lda, 64, TILE_M);
  if (rc) return rc;
  rc = d.b_mn ? make_map(mb, d.Bm, false, (uint64_t)d.N, (uint64_t)d.K, (uint64_t)d.ldb, 64, 64)
              : make_map(mb, d.Bm, false, (uint64_t)d.K, (uint64_t)d.N, (uint64_t)d.ldb, 64, BN);
  if (rc) return rc;
  *P = TcProblem{};
  P->M = (int)d.M; P->N = (int)d.N; P->K = (int)d.K;
  P->tiles_m = p.tiles_m; P->tiles_n = p.tiles_n; P->kb_total = p.kb_total; P->kb_per_split = p.kb_per_split;
  P->splits = p.splits; P->a_mn = d.a_mn; P->b_mn = d.b_mn;
  P->ksplit_add = p.ksplit_add;
  P->alpha = d.alpha; P->bias = d.bias; P->col_scale = d.col_scale;
  P->out = d.out; P->out_bf16 = d.out_bf16; P->ldo = d.ldo;
  P->out2 = d.out2; P->ldo2 = d.ldo2;
  // TMA store: one output, 16-byte aligned base and row pitch
  P->epi_tma = p.splits == 1 && d.out && !d.out2 && aligned16(d.out) && (d.ldo * (d.out_bf16 ? 2 : 4)) % 16 == 0;
  if (p.ksplit_add > 1 && !P->epi_tma) return IIF_EINVAL;   // (the planner only picks it for TMA-addressable fp32 outputs)
  if (P->epi_tma) {
    rc = make_map(mo, d.out, !d.out_bf16, (uint64_t)d.N, (uint64_t)d.M, (uint64_t)d.ldo, d.out_bf16 ? 64 : 32, 32);
    if (rc) return rc;
  } else {
    *mo = *ma;
  }
  P->partial = reinterpret_cast<float4*>(partial);
  P->counters = counters;
  P->db_out = d.db_out; P->db_partial = db_partial;
  return IIF_OK;
}

static long long* g_dbg = nullptr;   // iif_debug_timing
static std::atomic<int> g_reserved_slots{0};   // iif_gemm_reserve_slots
// In-kernel rendezvous (split-K, the loss-fused backward's grid barrier) need every CTA of the grid resident at
// once.  By default they are only used in grids the driver accepts as COOPERATIVE launches (co-residency
// guaranteed whatever else runs on the device; kernels of other streams can delay such a launch, never dead-lock
// it).  iif_gemm_assume_exclusive(1) additionally allows larger grids -- up to this library's own resident-CTA
// bound -- for callers that guarantee nothing else occupies the SMs while the head's launches run.
static std::atomic<int> g_exclusive{0};

// Launch one or two problems in one grid.
static int launch_group(const GemmDesc* d_in, int nprob, void* ws, size_t ws_bytes, cudaStream_t st,
                        const RowArgs* loss = nullptr, bool dry_run = false) {
  GemmDesc d[2];
  for (int i = 0; i < nprob; ++i) d[i] = d_in[i];
  int detail[6];
  int cap = resident_capacity(detail);
  if (cap <= 0) { cudaGetLastError(); return IIF_EDRIVER; }
  const int coop_cap = (detail[0] > 0 ? detail[0] : 1) * kNumSMs;   // what a cooperative launch of this kernel may hold
  const bool exclusive = g_exclusive.load(std::memory_order_relaxed) != 0;
  if (exclusive) {
    // resident-CTA slots promised to kernels that overlap these launches AND block on other GPUs (the
    // all-reduce): the in-kernel rendezvous may only count on the rest
    cap -= g_reserved_slots.load(std::memory_order_relaxed);
  } else if (cap > coop_cap) {
    cap = coop_cap;
  }
  if (cap < 1) cap = 1;
  TcGroup g{};
  CUtensorMap maps[6] = {};
  Plan plans[2];
  int tiles_total = 0;
  for (int i = 0; i < nprob; ++i) {
    if (!aligned16(d[i].A) || !aligned16(d[i].Bm)) return IIF_EALIGN;
    tiles_total += (int)(((d[i].M + TILE_M - 1) / TILE_M) * ((d[i].N + BN - 1) / BN));
  }
  for (int i = 0; i < nprob; ++i) {
    const int mine = (int)(((d[i].M + TILE_M - 1) / TILE_M) * ((d[i].N + BN - 1) / BN));
    plans[i] = make_plan(d[i].M, d[i].N, d[i].K, nprob > 1 ? tiles_total - mine : 0, cap);
  }
  if (nprob == 2) {   // both plans assumed the other unsplit: keep the sum co-resident, else drop the splits
    const int total = plans[0].tiles_m * plans[0].tiles_n * plans[0].splits + plans[1].tiles_m * plans[1].tiles_n * plans[1].splits;
    if ((plans[0].splits > 1 || plans[1].splits > 1) && total > cap) {
      for (int i = 0; i < 2; ++i)
        if (plans[i].splits > 1 && plans[1 - i].splits > 1) {   // shrink the one with the smaller K first
          const int v = plans[0].kb_total < plans[1].kb_total ? 0 : 1;
          plans[v].splits = 1; plans[v].kb_per_split = plans[v].kb_total;
          break;
        }
      const int t2 = plans[0].tiles_m * plans[0].tiles_n * plans[0].splits + plans[1].tiles_m * plans[1].tiles_n * plans[1].splits;
      if (t2 > cap)
        for (int i = 0; i < 2; ++i) { plans[i].splits = 1; plans[i].kb_per_split = plans[i].kb_total; }
    }
  }
  // Long K, few tiles, grid too large for the rendezvous (e.g. dW of a 64k-row batch: 128 tiles x 1024 k-blocks
  // next to 8192 dX tiles): split K over CTAs that reduce-add into the zeroed fp32 output through the TMA unit.
  for (int i = 0; i < nprob && !dry_run && !loss; ++i) {
    const int tiles = plans[i].tiles_m * plans[i].tiles_n;
    const bool tma_ok = d[i].out && !d[i].out_bf16 && !d[i].out2 && aligned16(d[i].out) && (d[i].ldo * 4) % 16 == 0;
    if (plans[i].splits == 1 && tma_ok && plans[i].kb_total >= 64 && tiles < 4 * kNumSMs) {
      int s = (4 * kNumSMs + tiles - 1) / tiles;
      if (s > plans[i].kb_total / 16) s = plans[i].kb_total / 16;
      if (s > 32) s = 32;
      if (s > 1) {
        plans[i].ksplit_add = s;
        cudaError_t e = cudaMemset2DAsync(d[i].out, (size_t)d[i].ldo * 4, 0, (size_t)d[i].N * 4, (size_t)d[i].M, st);
        if (e == cudaSuccess && d[i].db_out) e = cudaMemsetAsync(d[i].db_out, 0, (size_t)d[i].M * 4, st);
        if (e != cudaSuccess) return (int)e;
      }
    }
  }
  // CTAs are dispatched in block order: the problem with the longer K loop per CTA goes first (no long tail)
  if (nprob == 2) {
    auto per_cta = [&](int i) {
      const int ksp = plans[i].ksplit_add > 1 ? plans[i].ksplit_add : plans[i].splits;
      return (plans[i].kb_total + ksp - 1) / ksp;
    };
    if (per_cta(1) > per_cta(0)) { std::swap(d[0], d[1]); std::swap(plans[0], plans[1]); }
  }
  size_t need = 0;
  bool any_split = false;
  for (int i = 0; i < nprob; ++i) {
    need += plans[i].partial_bytes() + (d[i].db_out ? plans[i].db_bytes() : 0);
    any_split |= plans[i].splits > 1;
  }
  if (any_split || loss) need += WS_HEADER;
  if (need && (!ws || ws_bytes < need || !aligned16(ws))) return IIF_EWORKSPACE;
  int cta = 0;
  size_t off = WS_HEADER;
  for (int i = 0; i < nprob; ++i) {
    float* partial = plans[i].partial_bytes() ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + off) : nullptr;
    off += plans[i].partial_bytes();
    float* dbp = nullptr;
    if (d[i].db_out && plans[i].db_bytes()) {
      dbp = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + off);
      off += plans[i].db_bytes();
    }
    unsigned long long* counters =
        any_split ? reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + i * 2048) : nullptr;
    int rc = fill_problem(d[i], plans[i], partial, counters, dbp, &g.p[i], &maps[2 * i], &maps[2 * i + 1], &maps[4 + i]);
    if (rc) return rc;
    g.cta_begin[i] = cta;
    cta += plans[i].ctas();
  }
  g.cta_begin[nprob] = cta;
  g.nprob = nprob;
  g.dbg = g_dbg;
  // one wave of at most one CTA per SM: give each CTA the whole SM's shared memory (deeper TMA ring, more
  // bytes in flight per CTA); larger grids run two ~100 KB CTAs per SM
  int kb_max = 1;
  for (int i = 0; i < nprob; ++i) if (plans[i].kb_per_split > kb_max) kb_max = plans[i].kb_per_split;
  g.stages = MIN_STAGES;
  // (measured: for the 4-k-block CTAs of the head shapes a deeper ring only delays the first stage -- the
  // cold HBM fetch is not limited by bytes in flight per CTA; long unsplit K loops of a single-wave grid do use it)
  if (cta <= cap / 2 && !loss && kb_max >= 2 * MAX_STAGES) g.stages = MAX_STAGES;
  if (loss) {
    // the grid barrier needs every CTA resident; C <= 4096 with 256 threads per row
    if (cta > cap || loss->C > 4096 || (loss->C & 3) || !loss->scratch) return IIF_EUNSUPPORTED;
    g.fuse_loss = 1;
    g.loss_ne = loss->C <= 1024 ? 4 : (loss->C <= 2048 ? 8 : 16);
    g.grid_bar = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + 4096);
    g.loss = *loss;
  }
  if (nprob == 1) { maps[2] = maps[0]; maps[3] = maps[1]; maps[5] = maps[4]; }

  if (dry_run) return IIF_OK;
  const int variant = !g.fuse_loss ? 0 : (g.loss_ne == 4 ? 1 : (g.loss_ne == 8 ? 2 : 3));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cta);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem_bytes_for(g.stages);
  cfg.stream = st;
  void* kargs[7] = {&maps[0], &maps[1], &maps[2], &maps[3], &maps[4], &maps[5], &g};
  const bool rendezvous = any_split || loss != nullptr;
  cudaError_t e;
  if (rendezvous && cta <= coop_cap) {
    e = launch_cooperative(cfg, kernel_variant(variant), kargs, true);
  } else {
    if (rendezvous && !exclusive) return IIF_EUNSUPPORTED;   // (the planner keeps rendezvous grids within coop_cap)
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelExC(&cfg, kernel_variant(variant), kargs);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return (int)e;
  return IIF_OK;
}

static bool bad_dims(int64_t B, int64_t D, int64_t C) {
  return B < 0 || D <= 0 || C <= 0 || B > INT32_MAX || D > INT32_MAX || C > INT32_MAX;
}

static GemmDesc desc_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* cs,
                         float* z, int64_t ldz, float* zs, int64_t ldzs, int64_t B, int64_t D, int64_t C) {
  GemmDesc d{};
  d.A = x; d.lda = ldx; d.a_mn = false; d.Bm = w; d.ldb = ldw; d.b_mn = false; d.M = B; d.N = C; d.K = D;
  d.bias = bias; d.col_scale = cs; d.out = z; d.out_bf16 = 0; d.ldo = ldz; d.out2 = zs; d.ldo2 = ldzs;
  return d;
}
static GemmDesc desc_dx(const void* dz, int64_t lddz, const void* w, int64_t ldw, const float* alpha, void* dx,
                        int dx_bf16, int64_t lddx, int64_t B, int64_t D, int64_t C) {
  GemmDesc d{};
  d.A = dz; d.lda = lddz; d.a_mn = false; d.Bm = w; d.ldb = ldw; d.b_mn = true; d.M = B; d.N = D; d.K = C;
  d.alpha = alpha; d.out = dx; d.out_bf16 = dx_bf16; d.ldo = lddx;
  return d;
}
static GemmDesc desc_dw(const void* dz, int64_t lddz, const void* x, int64_t ldx, const float* alpha, float* dw,
                        int64_t lddw, int64_t B, int64_t D, int64_t C, float* db_out) {
  GemmDesc d{};
  d.A = dz; d.lda = lddz; d.a_mn = true; d.Bm = x; d.ldb = ldx; d.b_mn = true; d.M = C; d.N = D; d.K = B;
  d.alpha = alpha; d.out = dw; d.out_bf16 = 0; d.ldo = lddw;
  d.db_out = db_out;
  return d;
}

}  // namespace iif

using namespace iif;

extern "C" void iif_debug_timing(long long* buf) { g_dbg = buf; }
extern "C" int iif_gemm_reserve_slots(int slots) {
  if (slots < 0) return IIF_EINVAL;
  g_reserved_slots.store(slots, std::memory_order_relaxed);
  return IIF_OK;
}
extern "C" int iif_debug_capacity(int* detail6) { return resident_capacity(detail6); }
extern "C" int iif_gemm_assume_exclusive(int on) {
  g_exclusive.store(on ? 1 : 0, std::memory_order_relaxed);
  return IIF_OK;
}

extern "C" size_t iif_gemm_ws_bytes(int64_t B, int64_t D, int64_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 0;
  // Upper bound over every launch of the head: split-K never runs more than 2 x 148 CTAs, each
  // parking one 64 KB partial tile (+ 512 B of db partials), plus the counter header.
  const size_t multi = (size_t)WS_HEADER + (size_t)(2 * kNumSMs) * (TILE_M * BN * 4 + TILE_M * 4);
  const size_t fused = head_fused_ws_bytes(B, D, C);      // the one-launch step parks more partial tiles for some shapes
  return fused > multi ? fused : multi;
}

extern "C" int iif_linear_fwd_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                                   const float* col_scale, float* z, int64_t ldz, float* zs, int64_t ldzs, int64_t B,
                                   int64_t D, int64_t C, void* ws, size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !w || (B > 0 && !x) || (!z && !zs) || ldx < D || ldw < D) return IIF_EINVAL;
  if ((z && ldz < C) || (zs && (ldzs < C || !col_scale))) return IIF_EINVAL;
  if (B == 0) return IIF_OK;
  const GemmDesc d = desc_fwd(x, ldx, w, ldw, bias, col_scale, z, ldz, zs, ldzs, B, D, C);
  return launch_group(&d, 1, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dx_bf16(const void* dz, int64_t lddz, const void* w, int64_t ldw, const float* alpha_dev,
                                      void* dx, int dx_dtype, int64_t lddx, int64_t B, int64_t D, int64_t C, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !w || !dx || (B > 0 && !dz) || lddz < C || ldw < D || lddx < D) return IIF_EINVAL;
  if (dx_dtype != IIF_DTYPE_F32 && dx_dtype != IIF_DTYPE_BF16) return IIF_EINVAL;
  if (B == 0) return IIF_OK;
  const GemmDesc d = desc_dx(dz, lddz, w, ldw, alpha_dev, dx, dx_dtype == IIF_DTYPE_BF16, lddx, B, D, C);
  return launch_group(&d, 1, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dw_bf16(const void* dz, int64_t lddz, const void* x, int64_t ldx, const float* alpha_dev,
                                      float* dw, int64_t lddw, int64_t B, int64_t D, int64_t C, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !dw || (B > 0 && (!dz || !x)) || lddz < C || ldx < D || lddw < D) return IIF_EINVAL;
  if (B == 0) {  // empty batch: the gradient is exactly zero
    cudaError_t e = cudaMemset2DAsync(dw, lddw * 4, 0, D * 4, C, (cudaStream_t)stream);
    return e == cudaSuccess ? IIF_OK : (int)e;
  }
  const GemmDesc d = desc_dw(dz, lddz, x, ldx, alpha_dev, dw, lddw, B, D, C, nullptr);
  return launch_group(&d, 1, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_bf16(const void* dz, int64_t lddz, const void* x, int64_t ldx, const void* w, int64_t ldw,
                                   const float* alpha_dev, void* dx, int dx_dtype, int64_t lddx, float* dw,
                                   int64_t lddw, float* db, int64_t B, int64_t D, int64_t C, void* ws,
                                   size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !dw || (B > 0 && (!dz || !x)) || lddz < C || ldx < D || lddw < D) return IIF_EINVAL;
  if (dx && (!w || ldw < D || lddx < D || (dx_dtype != IIF_DTYPE_F32 && dx_dtype != IIF_DTYPE_BF16))) return IIF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {  // empty batch: the parameter gradients are exactly zero, dX is empty
    cudaError_t e = cudaMemset2DAsync(dw, lddw * 4, 0, D * 4, C, st);
    if (e == cudaSuccess && db) e = cudaMemsetAsync(db, 0, C * 4, st);
    return e == cudaSuccess ? IIF_OK : (int)e;
  }
  GemmDesc d[2];
  int n = 0;
  if (dx) d[n++] = desc_dx(dz, lddz, w, ldw, alpha_dev, dx, dx_dtype == IIF_DTYPE_BF16, lddx, B, D, C);
  d[n++] = desc_dw(dz, lddz, x, ldx, alpha_dev, dw, lddw, B, D, C, db);
  return launch_group(d, n, ws, ws_bytes, st);
}

static int loss_linear_bwd(const iif_head_args* h, void* stream, bool dry_run) {
  if (!h || !h->x || !h->label || !h->z || !h->dz_bf16 || !h->dw) return IIF_EINVAL;
  const int64_t B = h->B, D = h->D, C = h->C;
  if (bad_dims(B, D, C) || B == 0 || h->lddz % 8 != 0 || h->lddz < C || h->ldz < C || h->ldx < D || h->lddw < D)
    return B == 0 ? IIF_EUNSUPPORTED : IIF_EINVAL;
  if (h->dx && (!h->w || h->ldw < D || h->lddx < D)) return IIF_EINVAL;
  if ((h->loss_sum || h->acc_counts) && !h->scratch) return IIF_EINVAL;
  if (h->acc_counts && !h->rank) return IIF_EINVAL;
  RowArgs a;
  const bool vec = make_ce_row_args(a, h->z, h->ldz, h->iif, h->label, h->class_weight, h->sample_weight, h->ignore_index,
                                    h->scale, B, C, h->loss_i, h->loss_sum, nullptr, 0, h->dz_bf16, h->lddz, nullptr,
                                    h->argmax, h->rank, h->acc_counts, h->scratch);
  if (!vec) return IIF_EUNSUPPORTED;
  GemmDesc d[2];
  int n = 0;
  if (h->dx)
    d[n++] = desc_dx(h->dz_bf16, h->lddz, h->w, h->ldw, nullptr, h->dx, h->dx_dtype == IIF_DTYPE_BF16, h->lddx, B, D, C);
  d[n++] = desc_dw(h->dz_bf16, h->lddz, h->x, h->ldx, nullptr, h->dw, h->lddw, B, D, C, h->db);
  return launch_group(d, n, h->ws, h->ws_bytes, (cudaStream_t)stream, &a, dry_run);
}

extern "C" int iif_loss_linear_bwd_bf16(const iif_head_args* h, void* stream) { return loss_linear_bwd(h, stream, false); }

// Launches iif_head_fwd_bwd_bf16 will make for these arguments: 2 (loss rows fused into the backward
// launch) or 3; negative = argument error.
extern "C" int iif_head_launches(const iif_head_args* h) {
  if (!h) return IIF_EINVAL;
  if (!(h->flags & IIF_HEAD_NO_PERSISTENT)) {
    const int rf = head_fused_launch(h, nullptr, true);
    if (rf == IIF_OK) return 1;
    if (rf != IIF_EUNSUPPORTED) return rf;
  }
  if (h->flags & IIF_HEAD_NO_FUSED_LOSS) return 3;
  const int rc = loss_linear_bwd(h, nullptr, true);
  if (rc == IIF_OK) return 2;
  return rc == IIF_EUNSUPPORTED ? 3 : rc;
}
