"""GPU: the peer-memory all-reduce kernel (iif_allreduce_mean_f32).  One GPU: world = 1 through the raw C
ABI (handshake with itself, identity mean).  Two or more GPUs on the box: tools/check_allreduce.py under
torchrun against NCCL (skipped on a single-GPU box; the N>1 host logic is covered on CPU by test_dist_gloo)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allreduce_world1_identity():
    from iif_b200 import _lib
    lib = _lib.load()
    dev = "cuda:0"
    n = 1000 * 64 + 1000
    n4 = (n + 3) // 4 * 4
    buf = torch.randn(n4, device=dev)
    ref = buf.clone()
    flags = torch.zeros(int(lib.iif_allreduce_flag_bytes()) // 4, dtype=torch.int32, device=dev)
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    fl = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):   # flags / launch numbers advance monotonically
        _lib.check(lib.iif_allreduce_mean_f32(C.c_void_p(bufs.data_ptr()), C.c_void_p(fl.data_ptr()), None, 0, 1, 0, n4, 0, 0, _ % 2, st))
    torch.cuda.synchronize()
    assert torch.equal(buf, ref)
    assert lib.iif_allreduce_mean_f32(C.c_void_p(bufs.data_ptr()), C.c_void_p(fl.data_ptr()), None, 0, 1, 0, 6, 0, 0, 0, st) == _lib.EALIGN
    assert lib.iif_allreduce_mean_f32(C.c_void_p(bufs.data_ptr()), C.c_void_p(fl.data_ptr()), None, 2, 1, 0, 8, 0, 0, 0, st) == _lib.EINVAL


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with NVLink peer access")
def test_allreduce_two_gpus_vs_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_allreduce.py"), str(257 * 64 + 4)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "ALLREDUCE CHECK OK" in r.stdout, r.stdout[-3000:]
