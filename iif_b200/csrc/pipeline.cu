// Host-batch pipeline of the head step: the call a training loop makes when the step's features and
// labels sit in (pinned) HOST memory.  Three library-owned streams overlap, slot by slot,
//   h2d     : features + labels  host -> device          (copy engine, PCIe)
//   compute : fc_cls -> IIF softmax-CE fwd+bwd -> dX, dW, db   (the 3 launches of iif_head_fwd_bwd_bf16)
//   d2h     : the step's loss    device -> host           (copy engine)
// so step i+1's 1 MB feature copy hides under step i's kernels.  All GEMM launches stay on ONE compute
// stream: the split-K rendezvous of gemm_tc.cu needs a grid to itself (two such grids running
// concurrently could starve each other of resident-CTA slots), and one workspace serves every slot.
#include <string.h>

#include <new>

#include "common.cuh"

struct iif_pipeline {
  struct Slot {
    iif_head_args a;
    cudaEvent_t h2d_done, step_done, loss_done, release;
    bool used, held;
    bool device_only;                  // latest step came from submit_device: no loss was copied to the host
    bool waited;                       // the host has synchronised on the latest step (iif_pipeline_wait): buffers are free
    // host-batch submits: the loss goes straight into the caller's pinned host word when it is device-addressable
    float* direct_host; float* direct_dev; iif_head_args a_direct; bool direct_last;
    int64_t ar_offset;
    // staged mode (iif_pipeline_enable_staged): library-owned pinned host staging + one CUDA graph per slot
    void* host_x; int64_t* host_y; float* host_loss;   // pinned; host_loss is mapped (the kernel stores into it)
    iif_head_args a_staged;                            // = a, with loss_sum pointing at the mapped host_loss
    cudaGraphExec_t exec;
    int launches;                                      // kernel launches inside the slot's graph
  };
  int nslots;
  cudaStream_t s_h2d, s_compute, s_d2h, s_comm[4];
  int ar_lanes, ar_next;
  bool staged; int primed;               // slot whose inputs are already on the device (prefetched), or -1
  cudaGraphExec_t ring_exec; int ring_launches;   // staged mode: ONE graph holding a step of every slot (see enable_staged)
  Slot* slots;
  // optional data-parallel exchange after every step (iif_pipeline_set_allreduce)
  bool ar_on;
  void* const* ar_bufs; void* const* ar_flags; void* ar_mc;
  int ar_rank, ar_world, ar_ctas, ar_threads;
  int64_t ar_n;
};

#define IIF_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" int iif_pipeline_create(iif_pipeline** out, const iif_head_args* slot_args, int nslots) {
  if (!out || !slot_args || nslots < 1 || nslots > 64) return IIF_EINVAL;
  for (int i = 0; i < nslots; ++i) {
    const iif_head_args& a = slot_args[i];
    if (!a.x || !a.w || !a.label || !a.z || !a.dz_bf16 || !a.dw || !a.loss_sum || a.B <= 0 || a.D <= 0 || a.C <= 0 ||
        a.ldx < a.D)
      return IIF_EINVAL;
  }
  iif_pipeline* p = new (std::nothrow) iif_pipeline();
  if (!p) return IIF_EINVAL;
  p->nslots = nslots;
  p->slots = new (std::nothrow) iif_pipeline::Slot[nslots]();
  if (!p->slots) { delete p; return IIF_EINVAL; }
  IIF_CU(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
  IIF_CU(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
  IIF_CU(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
  {  // the all-reduce gates the reuse of gradient buffers: let its CTAs be scheduled ahead of queued GEMM CTAs
    int lo = 0, hi = 0;
    IIF_CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int l = 0; l < 4; ++l) IIF_CU(cudaStreamCreateWithPriority(&p->s_comm[l], cudaStreamNonBlocking, hi));
  }
  p->ar_lanes = 1;
  for (int i = 0; i < nslots; ++i) {
    iif_pipeline::Slot& s = p->slots[i];
    s.a = slot_args[i];
    IIF_CU(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    IIF_CU(cudaEventCreateWithFlags(&s.step_done, cudaEventDisableTiming));
    IIF_CU(cudaEventCreateWithFlags(&s.loss_done, cudaEventDisableTiming));
    IIF_CU(cudaEventCreateWithFlags(&s.release, cudaEventDisableTiming));
  }
  *out = p;
  return IIF_OK;
}

extern "C" int iif_pipeline_set_allreduce(iif_pipeline* p, void* const* peer_bufs_dev, void* const* peer_flags_dev,
                                          void* multicast_ptr, int rank, int world, const int64_t* slot_offsets_elems,
                                          int64_t n_elems, int num_ctas, int num_threads, int num_lanes) {
  if (num_lanes < 1 || num_lanes > 4) return IIF_EINVAL;
  if (!p || !peer_bufs_dev || !peer_flags_dev || !slot_offsets_elems || world < 1 || rank < 0 || rank >= world) return IIF_EINVAL;
  p->ar_bufs = peer_bufs_dev; p->ar_flags = peer_flags_dev; p->ar_mc = multicast_ptr;
  p->ar_rank = rank; p->ar_world = world; p->ar_n = n_elems; p->ar_ctas = num_ctas; p->ar_threads = num_threads;
  for (int i = 0; i < p->nslots; ++i) p->slots[i].ar_offset = slot_offsets_elems[i];
  p->ar_lanes = num_lanes; p->ar_next = 0;
  p->ar_on = world > 1;
  // two (or more) all-reduces in flight need their CTAs AND the step's CTA on every SM at once: that fits with
  // 128-thread all-reduce CTAs next to the 128-register build of the step kernel (see allreduce.cu)
  for (int i = 0; i < p->nslots; ++i) {
    if (p->ar_on && num_lanes > 1 && num_threads > 0 && num_threads <= 128) p->slots[i].a.flags |= IIF_HEAD_LOW_REGS;
    else p->slots[i].a.flags &= ~IIF_HEAD_LOW_REGS;
  }
  return IIF_OK;
}

// the step on the compute stream, then (data-parallel runs) the all-reduce of its gradients on the comm stream
static int run_step(iif_pipeline* p, iif_pipeline::Slot& s, const iif_head_args* args = nullptr) {
  if (s.held) { IIF_CU(cudaStreamWaitEvent(p->s_compute, s.release, 0)); s.held = false; }  // gradients still in flight
  if (int rc = iif_head_fwd_bwd_bf16(args ? args : &s.a, p->s_compute)) return rc;
  IIF_CU(cudaEventRecord(s.step_done, p->s_compute));
  if (p->ar_on) {
    const int lane = p->ar_next;
    p->ar_next = (p->ar_next + 1) % p->ar_lanes;
    IIF_CU(cudaStreamWaitEvent(p->s_comm[lane], s.step_done, 0));
    if (int rc = iif_allreduce_mean_f32(p->ar_bufs, p->ar_flags, p->ar_mc, p->ar_rank, p->ar_world, s.ar_offset, p->ar_n,
                                        p->ar_ctas, p->ar_threads, lane, p->s_comm[lane]))
      return rc;
    IIF_CU(cudaEventRecord(s.release, p->s_comm[lane]));
    s.held = true;
  }
  return IIF_OK;
}

extern "C" int iif_pipeline_submit(iif_pipeline* p, int slot, const void* host_x, const int64_t* host_label,
                                   float* host_loss) {
  if (!p || slot < 0 || slot >= p->nslots || !host_x || !host_label || !host_loss) return IIF_EINVAL;
  iif_pipeline::Slot& s = p->slots[slot];
  const iif_head_args& a = s.a;
  // the slot's device buffers are free once its previous step (and whoever held its gradients) is done -- nothing to
  // enqueue when the host has already waited for that step (every driver call here is ~2 us of a ~25 us step)
  if (s.used && !s.waited) IIF_CU(cudaStreamWaitEvent(p->s_h2d, s.direct_last ? s.step_done : s.loss_done, 0));
  if (a.ldx == a.D)
    IIF_CU(cudaMemcpyAsync(const_cast<void*>(a.x), host_x, (size_t)a.B * a.D * 2, cudaMemcpyHostToDevice, p->s_h2d));
  else
    IIF_CU(cudaMemcpy2DAsync(const_cast<void*>(a.x), (size_t)a.ldx * 2, host_x, (size_t)a.D * 2, (size_t)a.D * 2, a.B,
                             cudaMemcpyHostToDevice, p->s_h2d));
  IIF_CU(cudaMemcpyAsync(const_cast<int64_t*>(a.label), host_label, (size_t)a.B * 8, cudaMemcpyHostToDevice, p->s_h2d));
  IIF_CU(cudaEventRecord(s.h2d_done, p->s_h2d));
  IIF_CU(cudaStreamWaitEvent(p->s_compute, s.h2d_done, 0));
  // Loss read-back: when the caller's host word is pinned (device-addressable under unified addressing) the kernel's
  // own 4-byte store delivers it -- no copy stream, no extra events; otherwise a D2H copy on the third stream.
  if (s.direct_host != host_loss) {
    s.direct_host = host_loss;
    s.direct_dev = nullptr;
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, host_loss, 0) == cudaSuccess && dp) s.direct_dev = static_cast<float*>(dp);
    else cudaGetLastError();
    s.a_direct = s.a;
    s.a_direct.loss_sum = s.direct_dev;
  }
  if (s.direct_dev) {
    if (int rc = run_step(p, s, &s.a_direct)) return rc;        // (records step_done after the launch)
    s.direct_last = true;
  } else {
    if (int rc = run_step(p, s)) return rc;
    IIF_CU(cudaStreamWaitEvent(p->s_d2h, s.step_done, 0));
    IIF_CU(cudaMemcpyAsync(host_loss, a.loss_sum, 4, cudaMemcpyDeviceToHost, p->s_d2h));
    IIF_CU(cudaEventRecord(s.loss_done, p->s_d2h));
    s.direct_last = false;
  }
  s.used = true;
  s.waited = false;
  s.device_only = false;
  return IIF_OK;
}

// Same step with the inputs ALREADY in the slot's device buffers (no copies): the device-resident loop.
extern "C" int iif_pipeline_submit_device(iif_pipeline* p, int slot) {
  if (!p || slot < 0 || slot >= p->nslots) return IIF_EINVAL;
  p->slots[slot].device_only = true;
  p->slots[slot].waited = false;
  return run_step(p, p->slots[slot]);
}

// Make `stream` wait (on the device) for everything the pipeline has enqueued so far: the compute stream, the copy
// streams and EVERY comm lane -- e.g. to record a timing event after the last all-reduce of a timed region.
extern "C" int iif_pipeline_join(iif_pipeline* p, void* stream) {
  if (!p) return IIF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t ev;
  IIF_CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaStream_t all[7] = {p->s_h2d, p->s_compute, p->s_d2h, p->s_comm[0], p->s_comm[1], p->s_comm[2], p->s_comm[3]};
  for (cudaStream_t s : all) {
    if (s == st) continue;
    IIF_CU(cudaEventRecord(ev, s));
    IIF_CU(cudaStreamWaitEvent(st, ev, 0));
  }
  IIF_CU(cudaEventDestroy(ev));
  return IIF_OK;
}

extern "C" int iif_pipeline_get_streams(iif_pipeline* p, void** h2d, void** compute, void** d2h, void** comm) {
  if (!p) return IIF_EINVAL;
  if (h2d) *h2d = p->s_h2d;
  if (compute) *compute = p->s_compute;
  if (d2h) *d2h = p->s_d2h;
  if (comm) *comm = p->s_comm[0];
  return IIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// Staged mode: ONE driver call per step.  Every slot owns pinned host staging buffers (the data loader
// writes the next batch there) and a CUDA graph captured from two streams:
//     branch A (compute): forward launch -> loss + backward launch     of THIS slot
//     branch B (copy)   : H2D of the NEXT slot's features + labels      (prefetch distance 1)
// so the next batch's PCIe copy hides under this step's kernels without any per-step event traffic, and
// the loss reaches the host through the kernel's own 4-byte store into mapped pinned memory.  Graph
// launches are stream-ordered on the compute stream: steps never overlap each other (the split-K /
// grid-barrier rendezvous of the GEMM launches needs that).  The un-graphed submit path costs ~11
// driver calls per step, which on a slow or shared host is longer than the 30 us step itself.
// ---------------------------------------------------------------------------------------------------
static int copy_in(iif_pipeline* p, iif_pipeline::Slot& s, cudaStream_t st) {
  const iif_head_args& a = s.a;
  if (a.ldx == a.D)
    IIF_CU(cudaMemcpyAsync(const_cast<void*>(a.x), s.host_x, (size_t)a.B * a.D * 2, cudaMemcpyHostToDevice, st));
  else
    IIF_CU(cudaMemcpy2DAsync(const_cast<void*>(a.x), (size_t)a.ldx * 2, s.host_x, (size_t)a.D * 2, (size_t)a.D * 2, a.B,
                             cudaMemcpyHostToDevice, st));
  IIF_CU(cudaMemcpyAsync(const_cast<int64_t*>(a.label), s.host_y, (size_t)a.B * 8, cudaMemcpyHostToDevice, st));
  (void)p;
  return IIF_OK;
}

extern "C" int iif_pipeline_enable_staged(iif_pipeline* p) {
  if (!p || p->ar_on) return IIF_EINVAL;               // (data-parallel runs use the event-driven path)
  if (p->staged) return IIF_OK;
  cudaEvent_t fork, join;
  IIF_CU(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  IIF_CU(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
  for (int i = 0; i < p->nslots; ++i) {
    iif_pipeline::Slot& s = p->slots[i];
    IIF_CU(cudaHostAlloc(&s.host_x, (size_t)s.a.B * s.a.D * 2, cudaHostAllocDefault));
    IIF_CU(cudaHostAlloc(reinterpret_cast<void**>(&s.host_y), (size_t)s.a.B * 8, cudaHostAllocDefault));
    IIF_CU(cudaHostAlloc(reinterpret_cast<void**>(&s.host_loss), 64, cudaHostAllocMapped));
    memset(s.host_x, 0, (size_t)s.a.B * s.a.D * 2);
    memset(s.host_y, 0, (size_t)s.a.B * 8);
    *s.host_loss = 0.f;
    float* dev_loss = nullptr;
    IIF_CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev_loss), s.host_loss, 0));
    s.a_staged = s.a;
    s.a_staged.loss_sum = dev_loss;
    s.launches = iif_head_launches(&s.a_staged);
    if (s.launches < 0) return s.launches;
  }
  // warm every slot once outside capture (function attributes, tensor-map cache), then capture
  for (int i = 0; i < p->nslots; ++i) {
    if (int rc = copy_in(p, p->slots[i], p->s_compute)) return rc;
    if (int rc = iif_head_fwd_bwd_bf16(&p->slots[i].a_staged, p->s_compute)) return rc;
  }
  IIF_CU(cudaStreamSynchronize(p->s_compute));
  for (int i = 0; i < p->nslots; ++i) {
    iif_pipeline::Slot& s = p->slots[i];
    iif_pipeline::Slot& nxt = p->slots[(i + 1) % p->nslots];
    cudaGraph_t graph = nullptr;
    IIF_CU(cudaStreamBeginCapture(p->s_compute, cudaStreamCaptureModeRelaxed));
    IIF_CU(cudaEventRecord(fork, p->s_compute));
    IIF_CU(cudaStreamWaitEvent(p->s_h2d, fork, 0));
    int rc = copy_in(p, nxt, p->s_h2d);                               // branch B: prefetch the next slot's batch
    if (!rc) rc = cudaEventRecord(join, p->s_h2d) == cudaSuccess ? IIF_OK : IIF_EDRIVER;
    if (!rc) rc = iif_head_fwd_bwd_bf16(&s.a_staged, p->s_compute);   // branch A: this slot's step
    if (!rc) rc = cudaStreamWaitEvent(p->s_compute, join, 0) == cudaSuccess ? IIF_OK : IIF_EDRIVER;
    cudaError_t e = cudaStreamEndCapture(p->s_compute, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    IIF_CU(e);
    IIF_CU(cudaGraphInstantiate(&s.exec, graph, 0));
    cudaGraphDestroy(graph);
  }
  // The ring graph: one step of EVERY slot, in order, as one graph launch -- slot k's launch with the H2D of slot
  // k + 1's staged batch as a parallel branch, each slot's loss event recorded as an external event node.  One driver
  // call per `nslots` steps: the per-step graph-launch latency and the host's launch path leave the step time.
  {
    cudaGraph_t graph = nullptr;
    IIF_CU(cudaStreamBeginCapture(p->s_compute, cudaStreamCaptureModeRelaxed));
    int rc = copy_in(p, p->slots[0], p->s_compute);
    int launches = 0;
    for (int i = 0; i < p->nslots && !rc; ++i) {
      iif_pipeline::Slot& s = p->slots[i];
      const bool more = i + 1 < p->nslots;
      if (more) {
        if (cudaEventRecord(fork, p->s_compute) != cudaSuccess || cudaStreamWaitEvent(p->s_h2d, fork, 0) != cudaSuccess) rc = IIF_EDRIVER;
        if (!rc) rc = copy_in(p, p->slots[i + 1], p->s_h2d);
        if (!rc && cudaEventRecord(join, p->s_h2d) != cudaSuccess) rc = IIF_EDRIVER;
      }
      if (!rc) rc = iif_head_fwd_bwd_bf16(&s.a_staged, p->s_compute);
      if (!rc && cudaEventRecordWithFlags(s.loss_done, p->s_compute, cudaEventRecordExternal) != cudaSuccess) rc = IIF_EDRIVER;
      if (!rc && more && cudaStreamWaitEvent(p->s_compute, join, 0) != cudaSuccess) rc = IIF_EDRIVER;
      launches += s.launches;
    }
    cudaError_t e = cudaStreamEndCapture(p->s_compute, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    IIF_CU(e);
    IIF_CU(cudaGraphInstantiate(&p->ring_exec, graph, 0));
    cudaGraphDestroy(graph);
    p->ring_launches = launches;
  }
  cudaEventDestroy(fork);
  cudaEventDestroy(join);
  p->staged = true;
  p->primed = -1;
  return IIF_OK;
}

// One step of every slot (0 .. nslots-1, in order) from the slots' staging buffers, as ONE graph launch.  Contract:
// every slot's staging holds its batch when this is called and is not rewritten before iif_pipeline_wait(slot)
// returns for that slot (a data loader `nslots` batches ahead, refilling a slot once its loss has been read).
extern "C" int iif_pipeline_submit_staged_ring(iif_pipeline* p) {
  if (!p || !p->staged || !p->ring_exec) return IIF_EINVAL;
  IIF_CU(cudaGraphLaunch(p->ring_exec, p->s_compute));
  iif::g_launches.fetch_add((uint64_t)p->ring_launches, std::memory_order_relaxed);
  for (int i = 0; i < p->nslots; ++i) {
    p->slots[i].used = true;
    p->slots[i].waited = false;
    p->slots[i].device_only = false;
    p->slots[i].direct_last = false;
  }
  p->primed = -1;
  return IIF_OK;
}

extern "C" int iif_pipeline_staging(iif_pipeline* p, int slot, void** host_x, int64_t** host_label, float** host_loss) {
  if (!p || !p->staged || slot < 0 || slot >= p->nslots) return IIF_EINVAL;
  if (host_x) *host_x = p->slots[slot].host_x;
  if (host_label) *host_label = p->slots[slot].host_y;
  if (host_loss) *host_loss = p->slots[slot].host_loss;
  return IIF_OK;
}

extern "C" int iif_pipeline_submit_staged(iif_pipeline* p, int slot) {
  if (!p || !p->staged || slot < 0 || slot >= p->nslots) return IIF_EINVAL;
  iif_pipeline::Slot& s = p->slots[slot];
  if (p->primed != slot) {                              // first step, or the slots are not walked in order
    if (int rc = copy_in(p, s, p->s_compute)) return rc;
  }
  IIF_CU(cudaGraphLaunch(s.exec, p->s_compute));
  IIF_CU(cudaEventRecord(s.loss_done, p->s_compute));
  iif::g_launches.fetch_add((uint64_t)s.launches, std::memory_order_relaxed);
  p->primed = (slot + 1) % p->nslots;
  s.used = true;
  s.waited = false;
  s.device_only = false;
  s.direct_last = false;
  return IIF_OK;
}

extern "C" int iif_pipeline_wait(iif_pipeline* p, int slot) {
  if (!p || slot < 0 || slot >= p->nslots) return IIF_EINVAL;
  if (p->slots[slot].device_only) return IIF_EINVAL;   // submit_device copies no loss back: nothing to wait for
  if (!p->slots[slot].used) return IIF_OK;
  IIF_CU(cudaEventSynchronize(p->slots[slot].direct_last ? p->slots[slot].step_done : p->slots[slot].loss_done));
  p->slots[slot].waited = true;
  return IIF_OK;
}

extern "C" int iif_pipeline_stream_wait_step(iif_pipeline* p, int slot, void* stream) {
  if (!p || slot < 0 || slot >= p->nslots) return IIF_EINVAL;
  if (p->slots[slot].used) IIF_CU(cudaStreamWaitEvent((cudaStream_t)stream, p->slots[slot].step_done, 0));
  return IIF_OK;
}

extern "C" int iif_pipeline_hold_slot(iif_pipeline* p, int slot, void* stream) {
  if (!p || slot < 0 || slot >= p->nslots || p->ar_on) return IIF_EINVAL;   // with a built-in all-reduce the pipeline holds slots itself
  IIF_CU(cudaEventRecord(p->slots[slot].release, (cudaStream_t)stream));
  p->slots[slot].held = true;
  return IIF_OK;
}

extern "C" int iif_pipeline_sync(iif_pipeline* p) {
  if (!p) return IIF_EINVAL;
  IIF_CU(cudaStreamSynchronize(p->s_h2d));
  IIF_CU(cudaStreamSynchronize(p->s_compute));
  IIF_CU(cudaStreamSynchronize(p->s_d2h));
  for (int l = 0; l < 4; ++l) IIF_CU(cudaStreamSynchronize(p->s_comm[l]));
  return IIF_OK;
}

extern "C" void iif_pipeline_destroy(iif_pipeline* p) {
  if (!p) return;
  iif_pipeline_sync(p);
  for (int i = 0; i < p->nslots; ++i) {
    if (p->staged) {
      if (p->slots[i].exec) cudaGraphExecDestroy(p->slots[i].exec);
      if (i == 0 && p->ring_exec) cudaGraphExecDestroy(p->ring_exec);
      cudaFreeHost(p->slots[i].host_x);
      cudaFreeHost(p->slots[i].host_y);
      cudaFreeHost(p->slots[i].host_loss);
    }
    cudaEventDestroy(p->slots[i].h2d_done);
    cudaEventDestroy(p->slots[i].step_done);
    cudaEventDestroy(p->slots[i].loss_done);
    cudaEventDestroy(p->slots[i].release);
  }
  cudaStreamDestroy(p->s_h2d);
  cudaStreamDestroy(p->s_compute);
  cudaStreamDestroy(p->s_d2h);
  for (int l = 0; l < 4; ++l) cudaStreamDestroy(p->s_comm[l]);
  delete[] p->slots;
  delete p;
}
