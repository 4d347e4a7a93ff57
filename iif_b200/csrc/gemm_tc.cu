// (a)/(c) fc_cls GEMMs on the 5th-generation tensor cores: tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA through an mbarrier ring, one elected thread
// issuing the MMAs, a 4-warp epilogue reading TMEM with tcgen05.ld.
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
// Tile: 128 x 128 x 64 per CTA, 6-stage ring (32 KB / stage), 128-byte swizzle.  The head shapes
// are small (ImageNet-LT: 256x1000x2048), so K is split across CTAs to fill the 148 SMs; partial
// tiles go through an L2-resident fp32 workspace and the LAST CTA of a tile (ticket) sums them in
// split order -- deterministic, no float atomics.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (warp w may only touch TMEM lanes 32*(w%4) .. +31).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace iif {

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int STAGES = 6;
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;  // 16 KB

struct TcArgs {
  int M, N, K;
  int kb_total, kb_per_split, splits;
  const float* alpha; const float* bias; const float* col_scale;
  void* out; int out_bf16; int64_t ldo; int out_vec;
  float* out2; int64_t ldo2; int out2_vec;
  float* partial; int* tickets;
};

template <int BN>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * TILE_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

template <int BN, bool OUT2>
__device__ __forceinline__ void epilogue_store(const TcArgs& a, const float (&acc)[32], int m, int n_base, float alpha) {
  // one thread = one output row, 32 consecutive columns starting at n_base
  if (m >= a.M) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int n = n_base + j;
    v[j] = acc[j] * alpha + ((a.bias && n < a.N) ? __ldg(a.bias + n) : 0.f);
  }
  const bool full = n_base + 32 <= a.N;
  if (a.out) {
    if (a.out_bf16) {
      uint16_t* o = reinterpret_cast<uint16_t*>(a.out) + (int64_t)m * a.ldo + n_base;
      if (full && a.out_vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 q = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                               pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
          *reinterpret_cast<uint4*>(o + j) = q;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n_base + j < a.N) o[j] = bf16_bits(v[j]);
      }
    } else {
      float* o = reinterpret_cast<float*>(a.out) + (int64_t)m * a.ldo + n_base;
      if (full && a.out_vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n_base + j < a.N) o[j] = v[j];
      }
    }
  }
  if constexpr (OUT2) {
    if (a.out2) {
      float* o = a.out2 + (int64_t)m * a.ldo2 + n_base;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n_base + j;
        v[j] *= (n < a.N) ? __ldg(a.col_scale + n) : 0.f;
      }
      if (full && a.out2_vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n_base + j < a.N) o[j] = v[j];
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  const uint32_t flag_slot = tmem_slot + 4;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  volatile uint32_t* flag_p = reinterpret_cast<volatile uint32_t*>(smem_raw + (flag_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TILE_M, split = blockIdx.z;
  const int kb_begin = split * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + A_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
        const int k0 = kb * TILE_K;
        if constexpr (A_MN) {
#pragma unroll
          for (int j = 0; j < TILE_M / 64; ++j) ptx::tma_load_2d(sa + j * 8192, &tmA, full_bar(stage), m0 + 64 * j, k0);
        } else {
          ptx::tma_load_2d(sa, &tmA, full_bar(stage), k0, m0);
        }
        if constexpr (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * 8192, &tmB, full_bar(stage), n0 + 64 * j, k0);
        } else {
          ptx::tma_load_2d(sb, &tmB, full_bar(stage), k0, n0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(TILE_M, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES, sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < TILE_K / 16; ++k) {
          // K-major: 16 bf16 = 32 bytes along the swizzle row; 8-row groups 1024 B apart.
          // MN-major: 16 k-rows = 2048 bytes; 8-row groups 1024 B apart (SBO); 64-wide MN atoms 8192 B apart (LBO).
          const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                   : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                   : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
          ptx::umma_bf16(tmem_base, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::umma_commit(tmem_full_bar);       // accumulator complete
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    const int q = warp & 3;                  // TMEM lane quarter owned by this warp
    const int row = q * 32 + lane;           // row inside the tile
    const int m = m0 + row;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float alpha = a.alpha ? __ldg(a.alpha) : 1.f;
    if (a.splits == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c0, r);
        ptx::tmem_ld_wait();
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
        if (n0 + c0 < a.N) epilogue_store<BN, true>(a, acc, m, n0 + c0, alpha);
      }
    } else {
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      float* mine = a.partial + ((int64_t)tile * a.splits + split) * (TILE_M * BN) + row * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          __stcg(reinterpret_cast<float4*>(mine + c0 + j),
                 make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 128) *flag_p = (atomicAdd(a.tickets + tile, 1) == a.splits - 1) ? 1u : 0u;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*flag_p) {
        __threadfence();
        const float* base = a.partial + (int64_t)tile * a.splits * (TILE_M * BN) + row * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (n0 + c0 >= a.N) break;
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.f;
          for (int s = 0; s < a.splits; ++s) {   // fixed split order: deterministic sum
            const float* p = base + (int64_t)s * (TILE_M * BN) + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(p + j));
              acc[j] += t.x; acc[j + 1] += t.y; acc[j + 2] += t.z; acc[j + 3] += t.w;
            }
          }
          epilogue_store<BN, true>(a, acc, m, n0 + c0, alpha);
        }
        if (threadIdx.x == 128) a.tickets[tile] = 0;  // self-resetting
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, BN);
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

struct Plan { int tiles_m, tiles_n, kb_total, kb_per_split, splits; size_t ws_bytes; };

static Plan make_plan(int64_t M, int64_t N, int64_t K) {
  constexpr int BN = 128;
  Plan p;
  p.tiles_m = (int)((M + TILE_M - 1) / TILE_M);
  p.tiles_n = (int)((N + BN - 1) / BN);
  p.kb_total = (int)((K + TILE_K - 1) / TILE_K);
  if (p.kb_total < 1) p.kb_total = 1;
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  int want = tiles > 0 ? (int)(kNumSMs / tiles) : 1;   // fill one wave of the 148 SMs
  if (want < 1) want = 1;
  if (want > p.kb_total) want = p.kb_total;
  p.kb_per_split = (p.kb_total + want - 1) / want;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.ws_bytes = p.splits > 1 ? 1024 + ((size_t)tiles * 4 + 255) / 256 * 256 + (size_t)tiles * p.splits * TILE_M * BN * 4 : 0;
  return p;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcArgs& a, const Plan& p, cudaStream_t st) {
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN>::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid(p.tiles_n, p.tiles_m, p.splits);
  kern<<<grid, 256, SmemLayout<BN>::TOTAL, st>>>(ma, mb, a);
  return launch_status();
}

// OUT[M,N] = alpha * A . B^T;  a_mn / b_mn: operand stored with M (N) contiguous.
static int gemm_bf16(const void* A, int64_t lda, bool a_mn, const void* Bm, int64_t ldb, bool b_mn, int64_t M, int64_t N,
                     int64_t K, const float* alpha, const float* bias, const float* col_scale, void* out, int out_bf16,
                     int64_t ldo, float* out2, int64_t ldo2, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int BN = 128;
  if (M <= 0 || N <= 0) return IIF_OK;
  if (!aligned16(A) || !aligned16(Bm) || lda % 8 || ldb % 8) return IIF_EALIGN;
  const Plan p = make_plan(M, N, K);
  if (p.splits > 1 && (!ws || ws_bytes < p.ws_bytes)) return IIF_EWORKSPACE;
  CUtensorMap ma, mb;
  int rc;
  // K-major: memory [MN rows, K cols]; MN-major: memory [K rows, MN cols]
  rc = a_mn ? make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64)
            : make_map(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TILE_M);
  if (rc) return rc;
  rc = b_mn ? make_map(&mb, Bm, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64)
            : make_map(&mb, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BN);
  if (rc) return rc;
  TcArgs a{};
  a.M = (int)M; a.N = (int)N; a.K = (int)K;
  a.kb_total = p.kb_total; a.kb_per_split = p.kb_per_split; a.splits = p.splits;
  a.alpha = alpha; a.bias = bias; a.col_scale = col_scale;
  a.out = out; a.out_bf16 = out_bf16; a.ldo = ldo;
  a.out_vec = out && aligned16(out) && (out_bf16 ? ldo % 8 == 0 : ldo % 4 == 0);
  a.out2 = out2; a.ldo2 = ldo2; a.out2_vec = out2 && aligned16(out2) && ldo2 % 4 == 0;
  if (p.splits > 1) {
    uint8_t* w = reinterpret_cast<uint8_t*>(ws);
    const size_t tiles = (size_t)p.tiles_m * p.tiles_n;
    a.tickets = reinterpret_cast<int*>(w);
    size_t off = (tiles * 4 + 255) / 256 * 256;
    off = (off + 1023) / 1024 * 1024;
    a.partial = reinterpret_cast<float*>(w + off);
  }
  if (!a_mn && !b_mn) return launch_tc<BN, false, false>(ma, mb, a, p, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true>(ma, mb, a, p, st);
  if (a_mn && b_mn) return launch_tc<BN, true, true>(ma, mb, a, p, st);
  return launch_tc<BN, true, false>(ma, mb, a, p, st);
}

static bool bad_dims(int64_t B, int64_t D, int64_t C) {
  return B < 0 || D <= 0 || C <= 0 || B > INT32_MAX || D > INT32_MAX || C > INT32_MAX;
}

}  // namespace iif

using namespace iif;

extern "C" size_t iif_gemm_ws_bytes(int64_t B, int64_t D, int64_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 0;
  size_t m = make_plan(B, C, D).ws_bytes;
  size_t t = make_plan(B, D, C).ws_bytes; if (t > m) m = t;
  t = make_plan(C, D, B).ws_bytes; if (t > m) m = t;
  return m;
}

extern "C" int iif_linear_fwd_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                                   const float* col_scale, float* z, int64_t ldz, float* zs, int64_t ldzs, int64_t B,
                                   int64_t D, int64_t C, void* ws, size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !w || (B > 0 && !x) || (!z && !zs) || ldx < D || ldw < D) return IIF_EINVAL;
  if ((z && ldz < C) || (zs && (ldzs < C || !col_scale))) return IIF_EINVAL;
  return gemm_bf16(x, ldx, false, w, ldw, false, B, C, D, nullptr, bias, col_scale, z, 0, ldz, zs, ldzs, ws, ws_bytes,
                   (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dx_bf16(const void* dz, int64_t lddz, const void* w, int64_t ldw, const float* alpha_dev,
                                      void* dx, int dx_dtype, int64_t lddx, int64_t B, int64_t D, int64_t C, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !w || !dx || (B > 0 && !dz) || lddz < C || ldw < D || lddx < D) return IIF_EINVAL;
  if (dx_dtype != IIF_DTYPE_F32 && dx_dtype != IIF_DTYPE_BF16) return IIF_EINVAL;
  return gemm_bf16(dz, lddz, false, w, ldw, true, B, D, C, alpha_dev, nullptr, nullptr, dx, dx_dtype == IIF_DTYPE_BF16,
                   lddx, nullptr, 0, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dw_bf16(const void* dz, int64_t lddz, const void* x, int64_t ldx, const float* alpha_dev,
                                      float* dw, int64_t lddw, int64_t B, int64_t D, int64_t C, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (bad_dims(B, D, C) || !dw || (B > 0 && (!dz || !x)) || lddz < C || ldx < D || lddw < D) return IIF_EINVAL;
  if (B == 0) {  // empty batch: the gradient is exactly zero
    cudaError_t e = cudaMemset2DAsync(dw, lddw * 4, 0, D * 4, C, (cudaStream_t)stream);
    return e == cudaSuccess ? IIF_OK : (int)e;
  }
  return gemm_bf16(dz, lddz, true, x, ldx, true, C, D, B, alpha_dev, nullptr, nullptr, dw, 0, lddw, nullptr, 0, ws,
                   ws_bytes, (cudaStream_t)stream);
}
