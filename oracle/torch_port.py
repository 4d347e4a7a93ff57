"""fp32 torch-CPU restatement of one reference training step of the head (TEST INFRASTRUCTURE ONLY).

This is the CPU arm that ``bench.py`` times (``cpu_baseline`` / ``--impl reference``): the same ATen
calls the reference makes, in the same order and precision, restated from the cited lines -- the
reference's own Python cannot travel to the GPU box (``/root/reference`` does not exist there).

    z    = F.linear(x, W, b)                              cls/resnet_pytorch.py:219,293 ; seg/.../bbox_head.py:118
    loss = CE(reduction='none', weight)(z * iif, y)       cls/custom.py:10,30 ; seg/.../iif_loss.py:187-192
    loss = loss.mean() | loss.sum() | (loss*w).sum()/avg  cls/custom.py:32-36 ; seg/.../losses/utils.py:42-55
    loss.backward()                                       cls/train.py:77  (dX, dW, db by autograd)

Parity: pinned -- ``tests/test_oracle_golden.py::test_torch_port_*`` checks it against the golden
vectors frozen from the unmodified reference (tests/golden/cls_iif.npz, mmdet_iif.npz).
Nothing under ``iif_b200/`` may import this module.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def head_step(x, w, b, iif, y, *, reduction="mean", class_weight=None, sample_weight=None, avg_factor=None,
              ignore_index=-100, loss_weight=1.0, need_dx=True):
    """One fwd+bwd of the head on CPU tensors (fp32).  Returns dict(loss, z, dx, dw, db)."""
    x = x.detach().clone().requires_grad_(need_dx)
    w = w.detach().clone().requires_grad_(True)
    b = None if b is None else b.detach().clone().requires_grad_(True)
    z = F.linear(x, w, b)
    a = z if iif is None else z * iif.reshape(1, -1)
    li = F.cross_entropy(a, y, weight=class_weight, reduction="none", ignore_index=ignore_index)
    if sample_weight is not None:
        li = li * sample_weight
    if avg_factor is not None:
        loss = li.sum() / avg_factor if reduction == "mean" else li
    elif reduction == "mean":
        loss = li.mean()
    elif reduction == "sum":
        loss = li.sum()
    else:
        loss = li
    loss = loss_weight * loss
    (loss if loss.dim() == 0 else loss.sum()).backward()
    return dict(loss=loss.detach(), z=z.detach(), dx=x.grad, dw=w.grad, db=None if b is None else b.grad)


def time_head_step(B, D, C, *, steps, warmup, threads, seed=0, need_dx=True):
    """Seconds per step of the fp32 CPU head at [B,D]x[C,D] with `threads` torch threads."""
    torch.set_num_threads(max(int(threads), 1))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, generator=g)
    w = (torch.rand(C, D, generator=g) * 2 - 1) / D ** 0.5
    b = torch.full((C,), 0.01)
    iif = torch.rand(C, generator=g) * 6 + 0.5
    y = torch.randint(0, C, (B,), generator=g)
    fc = torch.nn.Linear(D, C)
    with torch.no_grad():
        fc.weight.copy_(w)
        fc.bias.copy_(b)
    xx = x.clone().requires_grad_(need_dx)

    def step():
        fc.zero_grad(set_to_none=True)
        if xx.grad is not None:
            xx.grad = None
        z = fc(xx)
        loss = F.cross_entropy(z * iif.reshape(1, -1), y, reduction="none").mean()
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(steps, 1)
