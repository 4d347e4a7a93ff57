"""fp32 parity modes of fc_cls on the GPU: fp32 FFMA (csrc/gemm_f32.cu), fp32 as 3 x bf16 on the tensor cores
(csrc/split3.cu + csrc/gemm_tc.cu) and the bf16 product, each timed (CUDA events, 50 calls after 5 warm-ups) with
its max relative error against a float64 product; stock torch fp32 (TF32 off) beside them.
Usage: python tools/fp32_modes.py [B,D,C ...]   -> one line per shape and mode on stdout."""
import sys
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from iif_b200 import ops  # noqa: E402

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    shapes = sys.argv[1:] or ["256,2048,1000", "16384,2048,1000", "2048,1024,1204"]
    for s in shapes:
        B, D, C = (int(v) for v in s.split(","))
        g = torch.Generator(device="cpu").manual_seed(0)
        x = torch.randn(B, D, generator=g).to(DEV)
        w = ((torch.rand(C, D, generator=g) * 2 - 1) / D ** 0.5).to(DEV)
        b = torch.full((C,), 0.01, device=DEV)
        ref = (x.double() @ w.double().T + b.double())
        scale = float(ref.abs().max())
        flops = 2.0 * B * D * C
        xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)

        def x3():
            return ops.linear_fwd(ops.split3(x, k_along_rows=False, side_b=False),
                                  ops.split3(w, k_along_rows=False, side_b=True), b)[0]
        x3a, x3b = ops.split3(x, k_along_rows=False, side_b=False), ops.split3(w, k_along_rows=False, side_b=True)
        modes = [
            ("fp32 FFMA (gemm_f32.cu)", lambda: ops.linear_fwd(x, w, b)[0]),
            ("fp32 as 3xbf16, splits + GEMM", x3),
            ("fp32 as 3xbf16, GEMM only (6K)", lambda: ops.linear_fwd(x3a, x3b, b)[0]),
            ("bf16 operands (gemm_tc.cu)", lambda: ops.linear_fwd(xb, wb, b)[0]),
            ("torch fp32 F.linear, TF32 off", lambda: torch.nn.functional.linear(x, w, b)),
        ]
        for name, fn in modes:
            z = fn()
            err = float((z.double() - ref).abs().max()) / scale
            us = timed(fn)
            print(f"{B}x{D}x{C}  {name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s (useful)  max rel err {err:.2e}",
                  flush=True)


if __name__ == "__main__":
    main()
