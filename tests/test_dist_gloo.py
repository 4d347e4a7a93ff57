"""World-size-2 (and 3) gloo tests of the N > 1 path on CPU: row sharding + all-reduce(mean) of the
flat head gradient reproduces DDP semantics (local normaliser, gradients averaged over ranks), and a
histogram built from sharded labels sums exactly.  Per-rank compute uses the float64 oracle (the
CUDA kernels need a GPU); the collective plumbing is the product code (iif_b200/parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import head_oracle as ho
from _common import head_inputs, iif_row, rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, D, C, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iif_b200 import parallel
        x, w, b, counts, y = head_inputs(B, D, C, seed=11)
        iif = iif_row(counts, "smooth")
        lo, hi = parallel.shard_rows(B, rank, world)
        r = ho.head_fwd_bwd(x[lo:hi], w, b, iif, y[lo:hi])            # local mean (custom.py:32-33)
        flat = torch.from_numpy(np.concatenate([r["dw"].ravel(), r["db"]]))
        parallel.allreduce_mean_(flat)
        cnt = torch.from_numpy(ho.label_hist(y[lo:hi], C))
        parallel.allreduce_counts_(cnt)
        if rank == 0:
            np.savez(out, flat=flat.numpy(), cnt=cnt.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 64), (3, 63)])
def test_sharded_head_matches_full_batch(tmp_path, world, B):
    D, C = 48, 17
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), B, D, C, out), nprocs=world, join=True)
    got = np.load(out)
    x, w, b, counts, y = head_inputs(B, D, C, seed=11)
    full = ho.head_fwd_bwd(x, w, b, iif_row(counts, "smooth"), y)     # equal shards: mean of means == global mean
    ref = np.concatenate([full["dw"].ravel(), full["db"]])
    assert rel_err(got["flat"], ref) < 1e-12
    assert np.array_equal(got["cnt"], ho.label_hist(y, C))


def test_shard_rows_partition():
    from iif_b200.parallel import shard_rows
    for n in (0, 1, 7, 256, 1000, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)
