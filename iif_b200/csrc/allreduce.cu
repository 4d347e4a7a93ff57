// (e) Data-parallel exchange of the head: all-reduce(mean) of the flat fp32 gradient buffer [dW | db]
// over NVLink / NVSwitch PEER MEMORY, by our own kernel instead of a NCCL ring.
//
// Every rank maps every rank's buffer (symmetric memory: the host passes a device array of per-rank base
// pointers; torch.distributed._symmetric_memory is only the allocator / rendezvous).  One launch per rank:
//   1. handshake: every CTA tells its peer CTAs "my gradients are final" (release, system scope) and
//      waits for theirs -- the launch is stream-ordered after the rank's own backward;
//   2. rank r owns 1/world of the elements: it reads that slice from ALL ranks (peer loads over NVLink),
//      adds them in rank order (bit-identical on every rank, run to run), scales by 1/world and writes the
//      result into ALL ranks' buffers (peer stores) -- reduce-scatter and all-gather in one pass, no
//      staging buffer;  with a multicast mapping (NVLS) the same slice is one multimem.ld_reduce (the
//      switch adds) and one multimem.st (the switch broadcasts);
//   3. handshake again: nobody leaves before every slice has landed everywhere.
// Traffic per GPU: (world-1)/world of the buffer in and out (8.2 MB head: 7.2 MB each way at world = 8),
// against 2x that for a ring; latency: two NVLink round trips instead of 2 (world-1) ring steps.
//
// Flags: a zero-initialised symmetric uint32 array [cta][peer] of monotonically growing values (see ar_barrier).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace iif {

constexpr int AR_MAX_WORLD = 16;
constexpr int AR_MAX_CTAS = 192;          // up to one (or a little more than one) CTA per SM
constexpr int AR_THREADS = 256;          // larger CTAs need an SM to themselves and have dead-locked against the GEMM grids
constexpr int AR_LANES = 4;             // independent flag sets: up to 4 all-reduces of one rank may be in flight
constexpr size_t AR_LANE_WORDS = (size_t)AR_MAX_CTAS * AR_MAX_WORLD + AR_MAX_CTAS;

// Cross-rank barrier among the CTAs with the same blockIdx.  Flags only ever grow: launch number e (kept
// per CTA in the rank's own flag memory) uses the values 2e (phase 0) and 2e+1 (phase 1); a signal is a
// plain store into the peer's copy (one-way NVLink latency, no remote atomic round trip), a wait polls the
// rank's own copy.
// System-scope fences are the expensive part of a multi-GPU handshake (a MEMBAR.SYS waits for every
// outstanding write to every peer: ~5 us on an 8-GPU NVSwitch box), so there is exactly ONE per CTA per
// launch: thread 0, before it announces "my slice has landed everywhere" (fence + relaxed stores = release
// pattern, cumulative over the CTA's stores ordered by the bar.sync).  The opening handshake needs none: the
// gradients it announces were written by the PREVIOUS kernel of the stream, and nothing read after either
// handshake can be stale -- peer data is read with ld.cv / multimem (never from a cached copy), and the
// reduced result is consumed by later kernels out of this GPU's own L2.
//
// The poll is an INTER-GPU wait: ranks of a data-parallel job skew by far more than a kernel's lifetime (a
// checkpoint or an evaluation on rank 0, a data-loader stall, a time-sliced GPU), so its bound is wall-clock
// (%globaltimer, which keeps its meaning across preemption), minutes by default and configurable / removable:
// iif_allreduce_set_timeout_ms, env IIF_B200_PEER_TIMEOUT_S (0 = wait forever, like NCCL without a watchdog).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ar_barrier(uint32_t* const* flags, size_t lane_off, int rank, int world, uint32_t value,
                                           int release, unsigned long long timeout_ns) {
  __syncthreads();
  const size_t slot = lane_off + (size_t)blockIdx.x * AR_MAX_WORLD;
  if (threadIdx.x == 0) {
    // release == 1: system scope (remote stores must have landed).  release == 2: gpu scope -- enough when the data
    // being published are LOCAL stores that peers will read over NVLink: peer reads are served by this GPU's L2, the
    // point a gpu-scope fence orders the CTA's stores to; the system-scope fence costs ~7 us even then.
    if (release == 1) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else if (release == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    for (int p = 0; p < world; ++p)
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flags[p] + slot + rank), "r"(value) : "memory");
  }
  if ((int)threadIdx.x < world) {
    const uint32_t* mine = flags[rank] + slot + threadIdx.x;
    unsigned long long t0 = 0;
    uint32_t seen, spins = 0;
    for (;;) {
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if ((int32_t)(seen - value) >= 0) break;
      if (++spins > 64) {                  // first ~64 polls back to back (the common case: the peer is microseconds away)
        __nanosleep(100);                  // then politely: the SM is shared with a CTA of the overlapping step
        if (timeout_ns && (spins & 1023u) == 0) {
          const unsigned long long now = global_ns();
          if (!t0) t0 = now;
          else if (now - t0 > timeout_ns) ptx::wait_timed_out(3);
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {   // never trust a cached copy of peer memory
  float4 r;
  asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Latency-bound by NVLink round trips (~2-4 us): what matters is bytes in flight.  Every thread keeps
// 16 (peer path: world x U) or 8 (multicast) independent 16-byte requests outstanding; the default launch is
// 16 CTAs x 256 threads (measured best under overlap with the next step's GEMMs at 8 GPUs).  At <= 128 registers
// x 256 threads a CTA of this kernel fits NEXT TO one ~100 KB GEMM CTA of the overlapping step (measured: a
// 512-thread CTA needs a whole SM and then waits for the backward launch to drain), costing the GEMM
// launches 16 of their 296 resident slots -- which the host RESERVES (iif_gemm_reserve_slots), because a
// GEMM grid spinning on a CTA that cannot become resident while this kernel waits on another GPU is a
// cross-rank deadlock (seen with 20 x 512-thread CTAs: exactly 256 slots left for a 256-CTA backward).
__device__ __forceinline__ void ar_stamp(long long* dbg, int slot) {
  if (dbg && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[(int64_t)blockIdx.x * 8 + slot] = t;
  }
}

template <bool MULTICAST, int U>
__global__ void __launch_bounds__(AR_THREADS, 2)
allreduce_mean_kernel(float* const* bufs, uint32_t* const* flags, float* mc, int rank, int world, int64_t n4,
                      int64_t off4, int lane, long long* dbg, unsigned long long timeout_ns) {
  ar_stamp(dbg, 0);
  const size_t lane_off = (size_t)lane * AR_LANE_WORDS;
  // launch number of this CTA (stream-ordered launches: no race), stored next to the flags
  uint32_t* epoch_p = flags[rank] + lane_off + (size_t)AR_MAX_CTAS * AR_MAX_WORLD + blockIdx.x;
  const uint32_t epoch = *epoch_p + 1;
  ar_barrier(flags, lane_off, rank, world, 4 * epoch, 0, timeout_ns);
  if (threadIdx.x == 0) *epoch_p = epoch;
  ar_stamp(dbg, 1);
  const int64_t per = (n4 + world - 1) / world;                    // float4 per rank slice
  const int64_t begin = rank * per, end = begin + per < n4 ? begin + per : n4;
  const float inv = 1.f / (float)world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (MULTICAST) {
    for (int64_t i0 = first; i0 < end; i0 += stride * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * stride < end) v[u] = mc_ld_reduce4(mc + (off4 + i0 + u * stride) * 4);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * stride < end) {
          v[u].x *= inv; v[u].y *= inv; v[u].z *= inv; v[u].w *= inv;
          mc_st4(mc + (off4 + i0 + u * stride) * 4, v[u]);
        }
    }
  } else {
    constexpr int PB = 16 / U;                                      // ranks per pass: PB x U loads in flight
    for (int64_t i0 = first; i0 < end; i0 += stride * U) {
      float4 acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p0 = 0; p0 < world; p0 += PB) {
        float4 v[PB][U];
#pragma unroll
        for (int p = 0; p < PB; ++p)
          if (p0 + p < world) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (i0 + u * stride < end) v[p][u] = ld_peer4(bufs[p0 + p] + (off4 + i0 + u * stride) * 4);
          }
#pragma unroll
        for (int p = 0; p < PB; ++p)                                // rank order: the same sum everywhere, every run
          if (p0 + p < world) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (i0 + u * stride < end) { acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w; }
          }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc[u].x *= inv; acc[u].y *= inv; acc[u].z *= inv; acc[u].w *= inv; }
      for (int p = 0; p < world; ++p) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i0 + u * stride < end) *reinterpret_cast<float4*>(bufs[p] + (off4 + i0 + u * stride) * 4) = acc[u];
      }
    }
  }
  ar_stamp(dbg, 2);
  ar_barrier(flags, lane_off, rank, world, 4 * epoch + 2, 1, timeout_ns);               // release: our peer stores land before the "done" flag
  ar_stamp(dbg, 3);
}

// ------------------------------------------------------------------------------------------------------------------
// Pull form (round 2): NO remote stores, hence no system-scope fence that waits for remote write acknowledgements --
// the closing cost of the push form above (8 us at 2 GPUs, 17 us at 8).
//   open : cross-rank barrier ("my gradients are final")
//   1    : rank r reduces ITS slice (multimem.ld_reduce, or peer loads summed in rank order), scales it and stores the
//          result into its OWN buffer -- local stores only
//   mid  : barrier with a system-scope release of those local stores ("my slice is reduced")
//   2    : every rank PULLS the other ranks' reduced slices with plain peer loads (flow-controlled by the loads
//          themselves; CTA b reads exactly what CTA b of the owner wrote, so per-CTA flags suffice) and stores them
//          locally
//   close: barrier ("I have finished reading your buffer": it may be overwritten again)
// In-bound NVLink traffic per GPU is the same (world-1)/world of the buffer; out-bound traffic is loads' replies only.
// ------------------------------------------------------------------------------------------------------------------
// Two geometries, both sized so that all-reduce CTAs are CO-RESIDENT with a CTA of the step kernel they overlap:
//   <256 threads, <= 80 registers>  = 20 K registers: one all-reduce next to the 168-register step kernel;
//   <128 threads, <= 128 registers> = 16 K registers: TWO all-reduces (two lanes in flight) next to the 128-register
//                                     variant of the step kernel (IIF_HEAD_LOW_REGS) -- 32 K + 2 x 16 K = the SM's file.
template <bool MULTICAST, int U, int TH, int MINB>
__global__ void __launch_bounds__(TH, MINB)
allreduce_pull_kernel(float* const* bufs, uint32_t* const* flags, float* mc, int rank, int world, int64_t n4,
                      int64_t off4, int lane, long long* dbg, unsigned long long timeout_ns, int mid_fence) {
  ar_stamp(dbg, 0);
  const size_t lane_off = (size_t)lane * AR_LANE_WORDS;
  uint32_t* epoch_p = flags[rank] + lane_off + (size_t)AR_MAX_CTAS * AR_MAX_WORLD + blockIdx.x;
  const uint32_t epoch = *epoch_p + 1;
  ar_barrier(flags, lane_off, rank, world, 4 * epoch, 0, timeout_ns);
  if (threadIdx.x == 0) *epoch_p = epoch;
  ar_stamp(dbg, 1);
  const int64_t per = (n4 + world - 1) / world;                    // float4 per rank slice
  const float inv = 1.f / (float)world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float* mine = bufs[rank];
  {  // ---- phase 1: my slice
    const int64_t begin = rank * per, end = begin + per < n4 ? begin + per : n4;
    for (int64_t i0 = begin + tid; i0 < end; i0 += stride * U) {
      float4 acc[U];
      if constexpr (MULTICAST) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i0 + u * stride < end) acc[u] = mc_ld_reduce4(mc + (off4 + i0 + u * stride) * 4);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {                            // rank order: the same sum everywhere, every run
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (i0 + u * stride < end) v[u] = ld_peer4(bufs[p] + (off4 + i0 + u * stride) * 4);
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (i0 + u * stride < end) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * stride < end) {
          acc[u].x *= inv; acc[u].y *= inv; acc[u].z *= inv; acc[u].w *= inv;
          *reinterpret_cast<float4*>(mine + (off4 + i0 + u * stride) * 4) = acc[u];
        }
    }
  }
  ar_stamp(dbg, 2);
  ar_barrier(flags, lane_off, rank, world, 4 * epoch + 1, mid_fence, timeout_ns);   // release: my LOCAL stores
  ar_stamp(dbg, 3);
  // ---- phase 2: pull the other ranks' slices.  A thread owns the SAME in-slice offsets o = tid + n * stride in every
  // slice (the owner's phase-1 mapping: CTA b reads exactly what CTA b of the owner wrote, which is what makes the
  // per-CTA barriers sufficient -- an earlier version flattened over all foreign elements instead, read other CTAs'
  // elements and failed the 8-GPU stress check); its work list is the (offset, peer) pairs, U of them in flight, so the
  // pull is a couple of NVLink round trips whatever the world size.  Each rank starts with a different peer.
  {
    const int nper = (int)((per + stride - 1) / stride);
    const int total = nper * (world - 1);
    for (int q0 = 0; q0 < total; q0 += U) {
      float4 v[U];
      int64_t idx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u;
        idx[u] = -1;
        if (q < total) {
          const int n = q / (world - 1), k = q - n * (world - 1);
          const int64_t o = tid + (int64_t)n * stride;
          const int p = (rank + 1 + k) % world;
          const int64_t i = p * per + o;
          if (o < per && i < n4) {
            idx[u] = i;
            v[u] = ld_peer4(bufs[p] + (off4 + i) * 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (idx[u] >= 0) *reinterpret_cast<float4*>(mine + (off4 + idx[u]) * 4) = v[u];
    }
  }
  ar_stamp(dbg, 4);
  ar_barrier(flags, lane_off, rank, world, 4 * epoch + 2, 0, timeout_ns);      // nobody still reads my buffer
  ar_stamp(dbg, 5);
}

}  // namespace iif

using namespace iif;

static long long* g_ar_dbg = nullptr;
extern "C" void iif_debug_timing_allreduce(long long* buf) { g_ar_dbg = buf; }

static std::atomic<long long> g_ar_timeout_ms{-1};       // -1: not set (env IIF_B200_PEER_TIMEOUT_S, else 600 s); 0: wait forever
extern "C" int iif_allreduce_set_timeout_ms(int64_t ms) {
  if (ms < 0) return IIF_EINVAL;
  g_ar_timeout_ms.store(ms, std::memory_order_relaxed);
  return IIF_OK;
}
static unsigned long long ar_timeout_ns() {
  long long ms = g_ar_timeout_ms.load(std::memory_order_relaxed);
  if (ms < 0) {
    const char* e = getenv("IIF_B200_PEER_TIMEOUT_S");
    ms = e ? (long long)(atof(e) * 1e3) : 600000ll;
    if (ms < 0) ms = 0;
    g_ar_timeout_ms.store(ms, std::memory_order_relaxed);
  }
  return (unsigned long long)ms * 1000000ull;
}

extern "C" size_t iif_allreduce_flag_bytes(void) {
  return AR_LANES * AR_LANE_WORDS * sizeof(uint32_t);   // per lane: [cta][peer] flags + [cta] launch numbers
}

extern "C" int iif_allreduce_mean_f32(void* const* peer_bufs_dev, void* const* peer_flags_dev, void* multicast_ptr, int rank,
                                      int world, int64_t offset_elems, int64_t n_elems, int num_ctas, int num_threads,
                                      int lane, void* stream) {
  if (!peer_bufs_dev || !peer_flags_dev || world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world) return IIF_EINVAL;
  if (lane < 0 || lane >= AR_LANES) return IIF_EINVAL;
  if (n_elems < 0 || offset_elems < 0 || (n_elems & 3) || (offset_elems & 3)) return IIF_EALIGN;
  if (n_elems == 0) return IIF_OK;
  {
    // default geometry by form (measured: see iif_b200/parallel.py): the pull form wants a CTA per SM, the push form
    // with in-switch reduction few CTAs
    const char* e = getenv("IIF_B200_AR_ALGO");
    const bool want_pull = (e && e[0] == 'p' && e[1] == 'u' && e[2] == 'l') ||
                           (!(e && e[0] == 'p' && e[1] == 'u' && e[2] == 's') && (world <= 2 || !multicast_ptr));
    if (num_ctas <= 0) num_ctas = want_pull ? kNumSMs : (world < 8 ? 48 : 16);
  }
  if (num_ctas > AR_MAX_CTAS) num_ctas = AR_MAX_CTAS;
  if (num_threads <= 0) num_threads = 256;
  if (num_threads > AR_THREADS || (num_threads & 31) || num_threads < 32) return IIF_EINVAL;
  float* const* bufs = reinterpret_cast<float* const*>(peer_bufs_dev);
  uint32_t* const* flags = reinterpret_cast<uint32_t* const*>(peer_flags_dev);
  cudaStream_t st = (cudaStream_t)stream;
  float* mc = reinterpret_cast<float*>(multicast_ptr);
  const int64_t n4 = n_elems / 4, off4 = offset_elems / 4;
  const unsigned long long tmo = ar_timeout_ns();
  // Form: measured (profiles/r2_multigpu.md) the pull form wins at 2 GPUs (28 us vs 36 us: no system-scope fence), the
  // one-pass push form with in-switch reduction + multicast stores from 4 GPUs on (33 us vs 49 us at 8: NVLS broadcasts
  // one slice to seven peers for one store, a pull moves seven slices as point-to-point loads).
  // IIF_B200_AR_ALGO=pull|push overrides.
  static const int algo = [] { const char* e = getenv("IIF_B200_AR_ALGO");
                               return !e ? 0 : ((e[0] == 'p' && e[1] == 'u' && e[2] == 'l') ? 1 : ((e[0] == 'p' && e[1] == 'u' && e[2] == 's') ? 2 : 0)); }();
  const bool pull = algo == 1 || (algo == 0 && (world <= 2 || !mc));
  if (pull) {
    static const int mid_fence = [] { const char* e = getenv("IIF_B200_AR_MIDFENCE"); return (e && e[0] == 's') ? 1 : 2; }();
    // in-switch reduction pays from 4 ranks on: at 2 ranks multimem.ld_reduce moves 340 GB/s where two plain peer
    // loads + an add move 610 GB/s (measured, profiles/r2_multigpu.md)
    if (num_threads <= 128) {
      if (mc && world >= 4) allreduce_pull_kernel<true, 16, 128, 4><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo, mid_fence);
      else allreduce_pull_kernel<false, 8, 128, 4><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo, mid_fence);
    } else {
      if (mc && world >= 4) allreduce_pull_kernel<true, 8, 256, 3><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo, mid_fence);
      else allreduce_pull_kernel<false, 4, 256, 3><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo, mid_fence);
    }
    return launch_status();
  }
  if (mc) allreduce_mean_kernel<true, 8><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else if (world <= 2) allreduce_mean_kernel<false, 8><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else if (world <= 4) allreduce_mean_kernel<false, 4><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else allreduce_mean_kernel<false, 2><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  return launch_status();
}
