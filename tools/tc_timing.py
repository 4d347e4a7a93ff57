"""Per-CTA phase timeline of the tensor-core kernel (iif_debug_timing): where do the microseconds go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iif_b200 import ops, _lib

dev = "cuda:0"
B, D, C = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,1000").split(","))
lib = _lib.load()
import ctypes
_d = (ctypes.c_int * 6)()
torch.zeros(1, device=dev)
print("resident CTA capacity", lib.iif_debug_capacity(_d), "detail [api/SM, by smem, by regs, regs, smem/SM, static smem]", list(_d))
bf = torch.bfloat16
NAMES = ["start", "prologue", "griddep", "tma_issued", "first_full", "last_full", "acc_done", "partial_out",
         "rendezvous", "end", "loss_rows_done", "grid_barrier"]

def show(tag, fn, n_cta_max=4096):
    buf = torch.zeros(n_cta_max * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # cold run: flush L2 by touching a big buffer
    big = torch.empty(512 << 20, dtype=torch.uint8, device=dev); big.zero_(); torch.cuda.synchronize()
    lib.iif_debug_timing(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.iif_debug_timing(None)
    t = buf.cpu().numpy().reshape(-1, 16)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print(f"== {tag}: {len(t)} CTAs, kernel span {(t[:, :12].max() - t0) / 1e3:.2f} us")
    for i, n in enumerate(NAMES):
        col = t[:, i]
        col = col[col > 0]
        if len(col):
            print(f"  {n:12s} median {np.median(col - t0) / 1e3:7.2f}  min {(col.min() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f} us")

x = torch.randn(B, D, device=dev).to(bf); w = (torch.randn(C, D, device=dev) * D ** -0.5).to(bf)
bias = torch.full((C,), 0.01, device=dev)
dz = (torch.randn(B, ops.pad8(C), device=dev) / B).to(bf)[:, :C]
show("fwd", lambda: ops.linear_fwd(x, w, bias))
show("bwd group", lambda: ops.linear_bwd(dz, x, w, dx_bf16=True))
show("dw only", lambda: ops.linear_bwd(dz, x, w, need_dx=False))
y = torch.randint(0, C, (B,), device=dev)
iif = torch.rand(C, device=dev) * 6 + 0.5
hs = ops.HeadStep(B, D, C, dev)
hs.bind(x, w, bias, iif, y)
k = dict(hs.kernels())
hs.launch(); torch.cuda.synchronize()
if "loss_linear_bwd_bf16" in k:
    show("loss+bwd fused", k["loss_linear_bwd_bf16"])
# ---- the whole step in place: forward, then the loss-fused backward, back to back after an L2 flush
def show_step():
    fns = hs.kernels()
    bufs = [torch.zeros(4096 * 16, dtype=torch.int64, device=dev) for _ in fns]
    for _ in range(3):
        hs.launch()
    torch.cuda.synchronize()
    big = torch.empty(512 << 20, dtype=torch.uint8, device=dev); big.zero_(); torch.cuda.synchronize()
    for (name, fn), b in zip(fns, bufs):
        lib.iif_debug_timing(b.data_ptr())
        fn()
    lib.iif_debug_timing(None)
    torch.cuda.synchronize()
    ts = [b.cpu().numpy().reshape(-1, 16) for b in bufs]
    ts = [t[t[:, 0] > 0] for t in ts]
    t0 = ts[0][:, 0].min()
    print(f"== whole step, in place (times relative to the first CTA of the forward launch)")
    for (name, _), t in zip(fns, ts):
        print(f"  -- {name}: {len(t)} CTAs")
        for i, n in enumerate(NAMES):
            col = t[:, i]; col = col[col > 0]
            if len(col):
                print(f"     {n:14s} median {np.median(col - t0) / 1e3:7.2f}  min {(col.min() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f} us")
show_step()
for i in range(3):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); ops.linear_fwd(x, w, bias); t1.record(); torch.cuda.synchronize()
    print("fwd warm event ms", t0.elapsed_time(t1))
