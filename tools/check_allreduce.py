"""torchrun --nproc-per-node N tools/check_allreduce.py : the peer-memory all-reduce kernel against NCCL
(values and device time) on N GPUs of one box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from iif_b200.parallel import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
numel = int(sys.argv[1]) if len(sys.argv) > 1 else 1000 * 2048 + 1000
ok = True
CFG = [(True, 0, 0), (False, 0, 0)] + ([(True, 32, 256), (True, 64, 256), (True, 148, 256), (True, 148, 128), (True, 148, 64),
                                        (False, 148, 256)] if os.environ.get("AR_SWEEP") else [])
for mc, nct, nth in CFG:
    try:
        par = PeerAllReduce(numel, 4, dev, use_multicast=mc, num_ctas=nct, num_threads=nth)
    except RuntimeError as e:
        if rank == 0:
            print("PeerAllReduce unavailable:", e)
        ok = False
        break
    if mc and not par.multicast:
        if rank == 0:
            print("no multicast mapping on this box: NVLS variant skipped")
        continue
    st = torch.cuda.current_stream(dev)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    for i in range(4):
        src = torch.randn(numel, generator=g).to(dev)
        par.buffer(i).copy_(src)
        ref = src.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        torch.cuda.synchronize(); dist.barrier()
        par.all_reduce(i, st)
        torch.cuda.synchronize(); dist.barrier()
        out = par.buffer(i)
        err = float((out - ref).abs().max() / ref.abs().max())
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out.contiguous())
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            print(f"multicast={par.multicast} buf {i}: max rel err vs NCCL {err:.2e}; identical on all ranks: {same}")
        ok = ok and err < 1e-6 and same
    # stress: many back-to-back calls on fresh data, every one compared with NCCL (a missing fence shows up here)
    bad = 0
    for it in range(60):
        i = it % 4
        src = torch.randn(numel, generator=g).to(dev)
        par.buffer(i).copy_(src)
        ref = src.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        par.all_reduce(i, st)
        par.all_reduce((i + 1) % 4, st)          # a second one right behind it (other buffer: values checked next round)
        out = par.buffer(i)
        bad += int(float((out - ref).abs().max() / ref.abs().max()) > 1e-6)
    tb = torch.tensor([bad], device=dev)
    dist.all_reduce(tb)
    if rank == 0:
        print(f"  stress: {int(tb)} mismatching calls of {60 * world}")
    ok = ok and int(tb) == 0
    # device time, back to back (each call is a cross-rank rendezvous, so this is the collective's latency)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fn, name in ((lambda i: par.all_reduce(i % 4, st), f"peer kernel multicast={par.multicast} ctas={nct} threads={nth}"),
                     (lambda i: dist.all_reduce(par.buffer(i % 4), op=dist.ReduceOp.AVG), "NCCL AVG")):
        for i in range(20):
            fn(i)
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for i in range(200):
            fn(i)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"  {name}: {float(t):.1f} us per all-reduce of {numel * 4 / 1e6:.1f} MB (max over ranks)")
    # phase timeline of one call (globaltimer stamps of every CTA)
    from iif_b200 import _lib
    import numpy as np
    dbg = torch.zeros(192 * 8, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    _lib.load().iif_debug_timing_allreduce(dbg.data_ptr())
    par.all_reduce(0, st)
    torch.cuda.synchronize()
    _lib.load().iif_debug_timing_allreduce(None)
    t = dbg.cpu().numpy().reshape(-1, 8); t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    if rank == 0:
        print("    rank0 timeline (us, median over CTAs): " + "  ".join(
            f"{n} {np.median(t[:, i] - t0) / 1e3:.1f}" for i, n in enumerate(["start", "open", "reduced", "mid", "pulled", "end"]) if (t[:, i] > 0).any()))
    del par
if rank == 0:
    print("ALLREDUCE CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
