"""Per-CTA phase timeline of the one-launch head step (iif_debug_timing_fused): where do the microseconds go?

    python tools/fused_timing.py [B,D,C]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iif_b200 import ops, _lib

dev = "cuda:0"
B, D, C = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,1000").split(","))
lib = _lib.load()
bf = torch.bfloat16
NAMES = ["start", "prologue", "griddep", "F_issued", "F_parked", "L_flags", "L_rows_done", "barrier", "B_issued",
         "B_drained", "R_done", "end", "L_hook", "L_row_ret", "R_flags", "B_acc_last", "F_first_req", "F_full_first",
         "F_full_last", "dX_full_first", "dX_full_last", "dW_full_first", "dW_full_last"]
plan = (ctypes.c_int * 12)()
rc = lib.iif_debug_fused_plan(B, D, C, 1, 148, plan)
print("plan rc", rc, "grid/fS/fItems/dxS/dxItems/dwItems/dxFirst/tpr/ne/rowBlocks/partMB", list(plan)[:11])
x = torch.randn(B, D, device=dev).to(bf); w = (torch.randn(C, D, device=dev) * D ** -0.5).to(bf)
bias = torch.full((C,), 0.01, device=dev)
y = torch.randint(0, C, (B,), device=dev)
iif = torch.rand(C, device=dev) * 6 + 0.5
hs = ops.HeadStep(B, D, C, dev)
hs.bind(x, w, bias, iif, y)
print("launches per step", hs.launches_per_step)
for _ in range(3):
    hs.launch()
torch.cuda.synchronize()

def show(tag, flush):
    buf = torch.zeros(4096 * 32, dtype=torch.int64, device=dev)
    if flush:
        big = torch.empty(512 << 20, dtype=torch.uint8, device=dev); big.zero_(); torch.cuda.synchronize()
    lib.iif_debug_timing_fused(buf.data_ptr())
    hs.launch()
    torch.cuda.synchronize()
    lib.iif_debug_timing_fused(None)
    t = buf.cpu().numpy().reshape(-1, 32)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print(f"== {tag}: {len(t)} CTAs, kernel span {(t[:, :32].max() - t0) / 1e3:.2f} us")
    if tag.endswith("flushed"):
        np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                             f"fused_stamps_{B}_{D}_{C}.npy"), t - t0)
    for i, n in enumerate(NAMES):
        col = t[:, i]
        col = col[col > 0]
        if len(col):
            print(f"  {n:12s} median {np.median(col - t0) / 1e3:7.2f}  min {(col.min() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f} us  ({len(col)} CTAs)")

show("one-launch step, L2 flushed", True)
show("one-launch step, L2 warm", False)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for n in (1, 50):
    e0.record()
    for _ in range(n):
        hs.launch()
    e1.record(); torch.cuda.synchronize()
    print(f"eager x{n}: {e0.elapsed_time(e1) * 1e3 / n:.2f} us per step (same set: operands L2-warm)")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        hs.launch()
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"graph of 20 steps: {e0.elapsed_time(e1) * 1e3 / 40:.2f} us per step (L2-warm)")
