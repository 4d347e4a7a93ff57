// (a)/(c) fc_cls GEMMs on the 5th-generation tensor cores: tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA through an mbarrier ring, one elected thread
// issuing the MMAs, a 4-warp TMEM drain and a block-wide coalesced epilogue.
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
// Tile: 128 x 128 x 64 per CTA, 128-byte swizzle, ring of `stages` 32 KB stages.  The head shapes
// are small (ImageNet-LT: 256 x 1000 x 2048), so K is split across the CTAs of a thread-block
// CLUSTER (cluster size = number of K splits, <= 8): every CTA parks its fp32 partial tile in an
// L2-resident workspace, the cluster barrier (release/acquire) orders the exchange, and then EVERY
// CTA of the cluster reduces 1/splits of the tile's rows in split order -- a parallel,
// deterministic reduction with no float atomics, no spin waits and no serial tail.
//
// Epilogue: the 4 drain warps move TMEM -> registers -> a padded staging tile in (recycled) stage
// memory; after one block barrier all 256 threads write 512-byte row segments (one warp = one row of
// the tile), so partials, fp32 / bf16 outputs and the IIF-scaled second output are fully coalesced.
//
// Bias gradient on the tensor cores: in the dW product the CTAs of the first tile column issue one
// extra N=16 MMA per k-step against a constant tile of ones, so db[c] = sum_b dZ[b,c] * 1 falls out
// of the same operand stream into 16 spare TMEM columns -- no column-sum kernel, no extra HBM read.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = ones tile, warps 4..7 = TMEM drain (warp w may only touch TMEM lanes 32*(w%4) .. +31).
//
// Every kernel begins with griddepcontrol.launch_dependents / .wait (programmatic dependent
// launch): the prologue (barrier init, TMEM alloc, tensor-map prefetch) overlaps the tail of the
// previous kernel of the step.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace iif {

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int BN = 128;
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * TILE_K * 2;      // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int LDS = BN + 4;           // staging tile row pitch (floats): conflict-free float4 rows
constexpr int MAX_STAGES = 6;
constexpr int MAX_SPLITS = 8;         // portable cluster size
constexpr int ONES_BYTES = 2048;      // 16 rows x 128 B of bf16 1.0 (K-major B operand of the db MMA)
constexpr int TMEM_COLS = 256;        // BN accumulator columns + 16 for db (power of two)

struct TcProblem {
  int M, N, K;
  int tiles_m, tiles_n, kb_total, kb_per_split, splits;
  int a_mn, b_mn;
  const float* alpha; const float* bias; const float* col_scale; int bias_vec;
  void* out; int out_bf16; int out_vec; int64_t ldo;
  float* out2; int out2_vec; int64_t ldo2;
  float* partial;                                   // [tiles][splits][TILE_M][BN] fp32
  float* db_out; float* db_partial;                 // dW only: db[m] (and [tiles_m][splits][TILE_M] partials)
};

struct TcGroup {
  int nprob, stages, cluster;
  int cta_begin[3];
  long long* dbg;                        // optional per-CTA phase timestamps (iif_debug_timing)
  TcProblem p[2];
};

__device__ __forceinline__ void stamp(const TcGroup& g, int slot) {
  if (g.dbg) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.dbg[(int64_t)blockIdx.x * 16 + slot] = t;
  }
}

__host__ __device__ inline int smem_bytes_for(int stages) { return stages * STAGE_BYTES + ONES_BYTES + 256 + 1024; }

__device__ __forceinline__ void emit4(const TcProblem& P, int m, int n, float4 v, float alpha) {
  // one thread = 4 consecutive columns of one output row
  if (m >= P.M || n >= P.N) return;
  float r[4] = {v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha};
  const bool full = n + 4 <= P.N;
  if (P.bias) {
    if (full && P.bias_vec) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n));
      r[0] += b.x; r[1] += b.y; r[2] += b.z; r[3] += b.w;
    } else {
      for (int j = 0; j < 4; ++j) if (n + j < P.N) r[j] += __ldg(P.bias + n + j);
    }
  }
  if (P.out) {
    if (P.out_bf16) {
      uint16_t* o = reinterpret_cast<uint16_t*>(P.out) + (int64_t)m * P.ldo + n;
      if (full && P.out_vec) stg_stream2(o, pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]));
      else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = bf16_bits(r[j]);
    } else {
      float* o = reinterpret_cast<float*>(P.out) + (int64_t)m * P.ldo + n;
      if (full && P.out_vec) stg_stream4(o, make_float4(r[0], r[1], r[2], r[3]));
      else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = r[j];
    }
  }
  if (P.out2) {
    float* o = P.out2 + (int64_t)m * P.ldo2 + n;
    for (int j = 0; j < 4; ++j) r[j] *= (n + j < P.N) ? __ldg(P.col_scale + n + j) : 0.f;
    if (full && P.out2_vec) stg_stream4(o, make_float4(r[0], r[1], r[2], r[3]));
    else for (int j = 0; j < 4; ++j) if (n + j < P.N) o[j] = r[j];
  }
}

__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ TcGroup g) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int stages = g.stages;
  const uint32_t ones_base = smem_base + stages * STAGE_BYTES;      // 1024-byte aligned
  const uint32_t bar_base = ones_base + ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(g, 0);
  const int pi = (g.nprob > 1 && (int)blockIdx.x >= g.cta_begin[1]) ? 1 : 0;
  const TcProblem& P = g.p[pi];
  const CUtensorMap* tmA = pi ? &tmA1 : &tmA0;
  const CUtensorMap* tmB = pi ? &tmB1 : &tmB0;
  const int local = (int)blockIdx.x - g.cta_begin[pi];
  const int split = local % P.splits;
  const int tile = local / P.splits;
  const bool has_work = tile < P.tiles_m * P.tiles_n;
  if (!has_work) {                       // padding CTA of a cluster: only keep the barrier balanced
    if (P.splits > 1) ptx::cluster_sync();
    return;
  }
  const int n0 = (tile % P.tiles_n) * BN, m0 = (tile / P.tiles_n) * TILE_M;
  const int kb_begin = split * P.kb_per_split;
  const int kb_end = min(P.kb_total, kb_begin + P.kb_per_split);
  const bool do_db = P.db_out != nullptr && n0 == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(tmA);
    ptx::prefetch_tensormap(tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
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
  ptx::griddep_launch_dependents();      // the next kernel may start its own prologue now
  ptx::griddep_wait();                   // ... and ours ends here: the producer kernel's data is visible
  if (threadIdx.x == 0) stamp(g, 2);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
        const int k0 = kb * TILE_K;
        if (P.a_mn) {
#pragma unroll
          for (int j = 0; j < TILE_M / 64; ++j) ptx::tma_load_2d(sa + j * 8192, tmA, full_bar(stage), m0 + 64 * j, k0);
        } else {
          ptx::tma_load_2d(sa, tmA, full_bar(stage), k0, m0);
        }
        if (P.b_mn) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * 8192, tmB, full_bar(stage), n0 + 64 * j, k0);
        } else {
          ptx::tma_load_2d(sb, tmB, full_bar(stage), k0, n0);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
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
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      ptx::umma_commit(tmem_full_bar);       // accumulator complete
    }
  } else if (warp < 4) {
    // idle until the epilogue
  } else {
    // ===================== drain: TMEM -> registers -> staging tile =====================
    const int q = warp & 3;                  // TMEM lane quarter owned by this warp
    const int row = q * 32 + lane;           // row inside the tile
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    if (threadIdx.x == 128) stamp(g, 6);   // accumulator complete
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* srow = reinterpret_cast<float*>(smem_gen) + row * LDS;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(taddr + c0, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(srow + c0 + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    if (do_db) {                             // db partial of this row: first of the 16 equal columns
      const uint32_t v = ptx::tmem_ld1(taddr + BN);
      ptx::tmem_ld_wait();
      srow[BN] = __uint_as_float(v);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();                           // staging tile complete; TMEM no longer needed
  if (threadIdx.x == 0) stamp(g, 7);
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }

  // ===================== block-wide coalesced epilogue =====================
  const float* stg = reinterpret_cast<const float*>(smem_gen);
  const float alpha = P.alpha ? __ldg(P.alpha) : 1.f;
  constexpr int F4 = BN / 4;                 // float4 per tile row: one warp covers one row
  if (P.splits == 1) {
    for (int idx = threadIdx.x; idx < TILE_M * F4; idx += 256) {
      const int row = idx / F4, c4 = idx % F4;
      emit4(P, m0 + row, n0 + c4 * 4, *reinterpret_cast<const float4*>(stg + row * LDS + c4 * 4), alpha);
    }
    if (do_db && threadIdx.x < TILE_M && m0 + (int)threadIdx.x < P.M)
      P.db_out[m0 + threadIdx.x] = stg[threadIdx.x * LDS + BN] * alpha;
  } else {
    float4* base = reinterpret_cast<float4*>(P.partial) + (int64_t)tile * P.splits * (TILE_M * F4);
    float4* mine = base + (int64_t)split * (TILE_M * F4);
    for (int idx = threadIdx.x; idx < TILE_M * F4; idx += 256) {
      const int row = idx / F4, c4 = idx % F4;
      __stcg(mine + idx, *reinterpret_cast<const float4*>(stg + row * LDS + c4 * 4));
    }
    float* dbp = do_db ? P.db_partial + (int64_t)(tile / P.tiles_n) * P.splits * TILE_M : nullptr;
    if (do_db && threadIdx.x < TILE_M) __stcg(dbp + split * TILE_M + threadIdx.x, stg[threadIdx.x * LDS + BN]);
    if (threadIdx.x == 0) stamp(g, 8);
    ptx::cluster_sync();                     // release our partial / acquire the other splits'
    if (threadIdx.x == 0) stamp(g, 9);
    const int rps = (TILE_M + P.splits - 1) / P.splits;
    const int r0 = split * rps, r1 = min(TILE_M, r0 + rps);
    for (int idx = threadIdx.x; idx < (r1 - r0) * F4; idx += 256) {
      const int row = r0 + idx / F4, c4 = idx % F4;
      float4 t[MAX_SPLITS];
#pragma unroll
      for (int s = 0; s < MAX_SPLITS; ++s)
        if (s < P.splits) t[s] = __ldcg(base + (int64_t)s * (TILE_M * F4) + row * F4 + c4);
      float4 acc = t[0];
#pragma unroll
      for (int s = 1; s < MAX_SPLITS; ++s)   // fixed split order: deterministic sum
        if (s < P.splits) { acc.x += t[s].x; acc.y += t[s].y; acc.z += t[s].z; acc.w += t[s].w; }
      emit4(P, m0 + row, n0 + c4 * 4, acc, alpha);
    }
    if (do_db && (int)threadIdx.x < r1 - r0 && m0 + r0 + (int)threadIdx.x < P.M) {
      float acc = 0.f;
      for (int s = 0; s < P.splits; ++s) acc += __ldcg(dbp + s * TILE_M + r0 + threadIdx.x);
      P.db_out[m0 + r0 + threadIdx.x] = acc * alpha;
    }
  }
  if (threadIdx.x == 0) stamp(g, 10);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t inner, outer, ld; uint32_t box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (uint64_t v : {k.inner, k.outer, k.ld, (uint64_t)k.box_outer}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};

// bf16 row-major [outer, inner] with leading dimension ld; box = 64 (inner) x box_outer, 128B swizzle,
// out-of-bounds elements read as zero (tile tails need no host padding).
static int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, inner, outer, ld, box_outer};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return IIF_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return IIF_EDRIVER;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return IIF_EDRIVER;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return IIF_OK;
}

struct Plan {
  int tiles_m, tiles_n, kb_total, kb_per_split, splits; size_t partial_bytes;
  size_t db_bytes() const { return splits > 1 ? (size_t)tiles_m * splits * TILE_M * 4 : 0; }   // 512-byte multiples
};

// K splits: the cluster size.  Cost model in k-block units: every CTA pays ~3 blocks of prologue +
// epilogue, a split tile ~1.5 more for the exchange; CTAs run in waves of 148.
static Plan make_plan(int64_t M, int64_t N, int64_t K, int other_ctas = 0) {
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
    const int waves = (ctas + kNumSMs - 1) / kNumSMs;
    const double cost = waves * (per + 3.0 + (s > 1 ? 1.5 : 0.0));
    if (cost < best - 1e-9) { best = cost; best_s = s; }
  }
  p.splits = best_s;
  p.kb_per_split = (p.kb_total + best_s - 1) / best_s;
  p.partial_bytes = p.splits > 1 ? (size_t)tiles * p.splits * TILE_M * BN * 4 : 0;
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

static int fill_problem(const GemmDesc& d, const Plan& p, float* partial, float* db_partial, TcProblem* P, CUtensorMap* ma,
                        CUtensorMap* mb) {
  if (!aligned16(d.A) || !aligned16(d.Bm) || d.lda % 8 || d.ldb % 8) return IIF_EALIGN;
  int rc;
  // K-major: memory [MN rows, K cols]; MN-major: memory [K rows, MN cols]
  rc = d.a_mn ? make_map(ma, d.A, (uint64_t)d.M, (uint64_t)d.K, (uint64_t)d.lda, 64)
              : make_map(ma, d.A, (uint64_t)d.K, (uint64_t)d.M, (uint64_t)d.lda, TILE_M);
  if (rc) return rc;
  rc = d.b_mn ? make_map(mb, d.Bm, (uint64_t)d.N, (uint64_t)d.K, (uint64_t)d.ldb, 64)
              : make_map(mb, d.Bm, (uint64_t)d.K, (uint64_t)d.N, (uint64_t)d.ldb, BN);
  if (rc) return rc;
  *P = TcProblem{};
  P->M = (int)d.M; P->N = (int)d.N; P->K = (int)d.K;
  P->tiles_m = p.tiles_m; P->tiles_n = p.tiles_n; P->kb_total = p.kb_total; P->kb_per_split = p.kb_per_split;
  P->splits = p.splits; P->a_mn = d.a_mn; P->b_mn = d.b_mn;
  P->alpha = d.alpha; P->bias = d.bias; P->col_scale = d.col_scale; P->bias_vec = d.bias && aligned16(d.bias);
  P->out = d.out; P->out_bf16 = d.out_bf16; P->ldo = d.ldo;
  P->out_vec = d.out && (d.out_bf16 ? ((reinterpret_cast<uintptr_t>(d.out) & 7u) == 0 && d.ldo % 4 == 0)
                                    : (aligned16(d.out) && d.ldo % 4 == 0));
  P->out2 = d.out2; P->ldo2 = d.ldo2; P->out2_vec = d.out2 && aligned16(d.out2) && d.ldo2 % 4 == 0;
  P->partial = partial;
  P->db_out = d.db_out; P->db_partial = db_partial;
  return IIF_OK;
}

static long long* g_dbg = nullptr;   // iif_debug_timing

// Launch one or two problems in one grid.
static int launch_group(const GemmDesc* d, int nprob, void* ws, size_t ws_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes_for(MAX_STAGES));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  TcGroup g{};
  CUtensorMap maps[4] = {};
  Plan plans[2];
  int tiles_total = 0;
  for (int i = 0; i < nprob; ++i) {
    if (!aligned16(d[i].A) || !aligned16(d[i].Bm)) return IIF_EALIGN;
    tiles_total += (int)(((d[i].M + TILE_M - 1) / TILE_M) * ((d[i].N + BN - 1) / BN));
  }
  for (int i = 0; i < nprob; ++i) {
    const int mine = (int)(((d[i].M + TILE_M - 1) / TILE_M) * ((d[i].N + BN - 1) / BN));
    plans[i] = make_plan(d[i].M, d[i].N, d[i].K, nprob > 1 ? tiles_total - mine : 0);
  }
  int cluster = 1;
  for (int i = 0; i < nprob; ++i) if (plans[i].splits > cluster) cluster = plans[i].splits;
  // a problem's splits must divide the cluster size so that clusters never straddle tiles unevenly
  for (int i = 0; i < nprob; ++i)
    while (cluster % plans[i].splits) {       // fall back to the next smaller exact split count
      int s = plans[i].splits - 1;
      for (; s > 1; --s) {
        const int per = (plans[i].kb_total + s - 1) / s;
        if ((plans[i].kb_total + per - 1) / per == s && cluster % s == 0) break;
      }
      if (s < 1) s = 1;
      plans[i].splits = s;
      plans[i].kb_per_split = (plans[i].kb_total + s - 1) / s;
      const int tiles = plans[i].tiles_m * plans[i].tiles_n;
      plans[i].partial_bytes = s > 1 ? (size_t)tiles * s * TILE_M * BN * 4 : 0;
    }
  size_t need = 0;
  for (int i = 0; i < nprob; ++i) need += plans[i].partial_bytes + (d[i].db_out ? plans[i].db_bytes() : 0);
  if (need && (!ws || ws_bytes < need)) return IIF_EWORKSPACE;
  int cta = 0;
  size_t off = 0;
  for (int i = 0; i < nprob; ++i) {
    float* partial = plans[i].partial_bytes ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + off) : nullptr;
    off += plans[i].partial_bytes;
    float* dbp = nullptr;
    if (d[i].db_out && plans[i].db_bytes()) {
      dbp = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + off);
      off += plans[i].db_bytes();
    }
    int rc = fill_problem(d[i], plans[i], partial, dbp, &g.p[i], &maps[2 * i], &maps[2 * i + 1]);
    if (rc) return rc;
    g.cta_begin[i] = cta;
    int n = plans[i].tiles_m * plans[i].tiles_n * plans[i].splits;
    n = (n + cluster - 1) / cluster * cluster;          // pad to whole clusters
    cta += n;
  }
  g.cta_begin[nprob] = cta;
  g.nprob = nprob;
  g.cluster = cluster;
  g.stages = MAX_STAGES;
  g.dbg = g_dbg;
  if (nprob == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cta);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem_bytes_for(g.stages);
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (cluster > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)cluster;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel, maps[0], maps[1], maps[2], maps[3], g);
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

extern "C" size_t iif_gemm_ws_bytes(int64_t B, int64_t D, int64_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 0;
  // upper bound: any problem of the head with the largest split count
  auto cap = [](int64_t M, int64_t N) {
    return (size_t)((M + TILE_M - 1) / TILE_M) * (size_t)((N + BN - 1) / BN) * MAX_SPLITS * TILE_M * BN * 4;
  };
  auto need = [&](int64_t M, int64_t N, int64_t K, int other) {
    const Plan p = make_plan(M, N, K, other);
    return p.splits > 1 ? cap(M, N) / MAX_SPLITS * p.splits + p.db_bytes() : (size_t)0;
  };
  const int t_dx = (int)(((B + TILE_M - 1) / TILE_M) * ((D + BN - 1) / BN));
  const int t_dw = (int)(((C + TILE_M - 1) / TILE_M) * ((D + BN - 1) / BN));
  size_t m = need(B, C, D, 0);
  size_t t = need(B, D, C, 0); if (t > m) m = t;
  t = need(C, D, B, 0); if (t > m) m = t;
  t = need(B, D, C, t_dw) + need(C, D, B, t_dx); if (t > m) m = t;
  return m;
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
