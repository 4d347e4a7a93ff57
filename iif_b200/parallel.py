"""Data-parallel plumbing of the head: row sharding + one all-reduce of the parameter gradients.

The head shards by sample: every rank owns its own rows of X / labels / sample
weights, W / b / iif are replicated, the loss normaliser is LOCAL (mean over the local batch,
classification/custom.py:32-33; local avg_factor, mmdet bbox_head.py:267) and DDP then AVERAGES
the parameter gradients over ranks (classification/train.py:231-234; mmdet apis/train.py:81-85).
The only exchange is therefore one all-reduce(mean) of dW [C,D] + db [C] -- kept in one flat fp32
buffer (`ops.HeadStep.grad_flat`) so it is a single NCCL message over NVLink / NVSwitch.
`torch.distributed` is the transport (nccl on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

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
