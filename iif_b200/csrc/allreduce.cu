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
                                           bool release, unsigned long long timeout_ns) {
  __syncthreads();
  const size_t slot = lane_off + (size_t)blockIdx.x * AR_MAX_WORLD;
  if (threadIdx.x == 0) {
    if (release) asm volatile("fence.acq_rel.sys;" ::: "memory");
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
  ar_barrier(flags, lane_off, rank, world, 2 * epoch, false, timeout_ns);
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
  ar_barrier(flags, lane_off, rank, world, 2 * epoch + 1, true, timeout_ns);            // release: our peer stores land before the "done" flag
  ar_stamp(dbg, 3);
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
  if (num_ctas <= 0) num_ctas = 16;
  if (num_ctas > AR_MAX_CTAS) num_ctas = AR_MAX_CTAS;
  if (num_threads <= 0) num_threads = 256;
  if (num_threads > AR_THREADS || (num_threads & 31) || num_threads < 32) return IIF_EINVAL;
  float* const* bufs = reinterpret_cast<float* const*>(peer_bufs_dev);
  uint32_t* const* flags = reinterpret_cast<uint32_t* const*>(peer_flags_dev);
  cudaStream_t st = (cudaStream_t)stream;
  float* mc = reinterpret_cast<float*>(multicast_ptr);
  const int64_t n4 = n_elems / 4, off4 = offset_elems / 4;
  const unsigned long long tmo = ar_timeout_ns();
  if (mc) allreduce_mean_kernel<true, 8><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else if (world <= 2) allreduce_mean_kernel<false, 8><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else if (world <= 4) allreduce_mean_kernel<false, 4><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  else allreduce_mean_kernel<false, 2><<<num_ctas, num_threads, 0, st>>>(bufs, flags, mc, rank, world, n4, off4, lane, g_ar_dbg, tmo);
  return launch_status();
}
