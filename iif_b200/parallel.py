"""Data-parallel plumbing of the head: row sharding + one all-reduce of the parameter gradients.

The head shards by sample: every rank owns its own rows of X / labels / sample
weights, W / b / iif are replicated, the loss normaliser is LOCAL (mean over the local batch,
classification/custom.py:32-33; local avg_factor, mmdet bbox_head.py:267) and DDP then AVERAGES
the parameter gradients over ranks (classification/train.py:231-234; mmdet apis/train.py:81-85).
The only exchange is therefore one all-reduce(mean) of dW [C,D] + db [C] -- kept in one flat fp32
buffer (`ops.HeadStep.grad_flat`).  Two transports:
  * `PeerAllReduce`: the library's own one-pass kernel over NVLink peer memory (csrc/allreduce.cu:
    reduce-scatter + all-gather fused, NVLS multimem when a multicast mapping exists); symmetric memory
    from torch.distributed._symmetric_memory is only the allocator / rendezvous;
  * `allreduce_mean_`: torch.distributed (nccl on GPUs -- the baseline the kernel is measured against;
    gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import ctypes as C
import os

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the rows owned by `rank`: contiguous blocks, sizes differing by at most one
    (the first n_rows % world ranks get the extra row), like torch's DistributedSampler without padding."""
    if world <= 0 or not (0 <= rank < world) or n_rows < 0:
        raise ValueError(f"bad shard request: n_rows={n_rows} rank={rank} world={world}")
    base, extra = divmod(n_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_mean_(flat: torch.Tensor, group=None, async_op: bool = False):
    """In-place mean over ranks of the flat gradient buffer (DDP semantics: sum / world).
    NCCL reduces with ReduceOp.AVG in one pass; gloo (CPU tests) sums then divides."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if flat.is_cuda:
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=False)
    flat.div_(dist.get_world_size(group))
    return w


def allreduce_counts_(counts: torch.Tensor, group=None):
    """Histogram built from sharded labels: integer sum over ranks (exact, order-independent)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class GradReducer:
    """Overlaps the all-reduce of a step's flat gradient with the following compute: the collective
    runs on a side stream ordered after the step's last kernel by an event; `wait()` orders the
    current stream after the reduction (call it before the buffer is read or overwritten)."""

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.stream = torch.cuda.Stream(self.device)
        self._done: Optional[torch.cuda.Event] = None

    def start(self, flat: torch.Tensor):
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            allreduce_mean_(flat, self.group)
            self._done = torch.cuda.Event()
            self._done.record(self.stream)
        flat.record_stream(self.stream)

    def wait(self):
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None


class PeerAllReduce:
    """`nbuf` flat fp32 gradient buffers in symmetric (peer-mapped) memory + the library's all-reduce(mean)
    kernel (iif_allreduce_mean_f32).  `buffer(i)` is handed to `ops.HeadStep(grad_flat=...)`; `all_reduce(i,
    stream)` enqueues the collective for buffer i on `stream` (every rank, same order).

    Raises RuntimeError when symmetric memory cannot be set up (no NVLink peer access, single process
    without a group): callers fall back to `allreduce_mean_` explicitly -- never silently."""

    def __init__(self, numel: int, nbuf: int, device, group=None, use_multicast: bool = True, num_ctas: int = 0,
                 num_threads: int = 0, lanes: int = 2):
        from . import _lib
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device(device)
        self.numel = int(numel)
        self.stride = (self.numel + 63) // 64 * 64           # 256-byte aligned slots
        self.nbuf = int(nbuf)
        self._lib = _lib.load()
        self._check = _lib.check
        # measured on one 8 x B200 box (profiles/r2_multigpu.md): 2 GPUs -> the pull form, one CTA per SM (28 us alone,
        # 33 us/step overlapped); 4+ GPUs -> the push form with in-switch reduction, where FEWER CTAs interfere less
        # with the step they overlap (8 GPUs: 16 CTAs 39.8 us/step, 48 CTAs 44.7, 148 CTAs 45.0; 4 GPUs, where a rank's
        # slice is twice as long: 16 CTAs 43.8, 48 CTAs 41.1, 148 CTAs 49.1)
        self.num_threads = int(num_threads) if num_threads else 256
        if self.num_threads > 256 or self.num_threads % 32:
            raise ValueError("PeerAllReduce: num_threads must be a multiple of 32, at most 256 (a larger CTA needs an SM "
                             "to itself and can dead-lock against the GEMM launches it overlaps)")
        self.lanes = max(1, min(int(lanes), 4))     # all-reduces of consecutive steps that may be in flight at once
        try:
            self.mem = symm_mem.empty(self.nbuf * self.stride, dtype=torch.float32, device=self.device)
            self.mem.zero_()
            self.hdl = symm_mem.rendezvous(self.mem, self.group)
            nflag = int(self._lib.iif_allreduce_flag_bytes()) // 4
            self.flags = symm_mem.empty(nflag, dtype=torch.int32, device=self.device)
            self.flags.zero_()
            self.fhdl = symm_mem.rendezvous(self.flags, self.group)
        except Exception as e:  # noqa: BLE001 - surfaced, not swallowed
            raise RuntimeError(f"PeerAllReduce: symmetric memory unavailable ({type(e).__name__}: {e})") from e
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)               # every rank's flags are zero before anyone signals
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast = bool(mc) and use_multicast
        self._mc = C.c_void_p(mc if self.multicast else 0)
        self._bufs = C.c_void_p(int(self.hdl.buffer_ptrs_dev))
        self._flags = C.c_void_p(int(self.fhdl.buffer_ptrs_dev))
        pull = self.form.startswith("pull")
        self.num_ctas = int(num_ctas) if num_ctas else (148 if pull else (48 if self.world < 8 else 16))
        # The all-reduce overlaps the next step's GEMM launches and blocks on other GPUs: keep its footprint
        # out of the resident-CTA budget their in-kernel rendezvous rely on (include/iif_b200.h).
        self._check(self._lib.iif_gemm_reserve_slots(self.lanes * self.num_ctas),
                    "gemm_reserve_slots")

    @property
    def form(self) -> str:
        """Which kernel iif_allreduce_mean_f32 picks (same rule as csrc/allreduce.cu; IIF_B200_AR_ALGO overrides)."""
        env = os.environ.get("IIF_B200_AR_ALGO", "")[:3]
        pull = env == "pul" or (env != "pus" and (self.world <= 2 or not self.multicast))
        if pull:
            how = "in-switch multimem reduce" if (self.multicast and self.world >= 4) else "peer loads"
            return f"pull form: local slice reduced with {how}, gathered with peer loads"
        return ("push form: in-switch multimem.ld_reduce + multimem.st broadcast" if self.multicast
                else "push form: peer loads + remote stores")

    def buffer(self, i: int) -> torch.Tensor:
        return self.mem[i * self.stride: i * self.stride + self.numel]

    def all_reduce(self, i: int, stream: torch.cuda.Stream, lane: int = 0) -> None:
        n = (self.numel + 3) // 4 * 4           # the slot is padded: reduce whole float4s
        rc = self._lib.iif_allreduce_mean_f32(self._bufs, self._flags, self._mc, self.rank, self.world, i * self.stride, n,
                                              self.num_ctas, self.num_threads, lane, C.c_void_p(stream.cuda_stream))
        if rc:
            self._check(rc, "allreduce_mean_f32")
