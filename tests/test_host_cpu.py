"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol that
include/iif_b200.h declares, the ctypes mirror matches the C struct layout, host-side helper
logic (reductions, error conventions, the shared ndtri restatement) -- no compute calls."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "iif_b200.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"IIF_API\s+[\w\s\*]+?\b(iif_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = header_symbols()
    for n in ("iif_hist_labels_i64", "iif_hist_images_dedup_i64", "iif_weights_from_counts", "iif_softmax_ce_fwd_bwd",
              "iif_sigmoid_bce_fwd_bwd", "iif_scaled_activation", "iif_linear_fwd_bf16", "iif_linear_fwd_f32",
              "iif_linear_bwd_dx_bf16", "iif_linear_bwd_dw_bf16", "iif_head_fwd_bwd_bf16"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from iif_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert sorted(_lib.SIGNATURES) == names          # ctypes mirror and header agree one to one
    for n in names:
        assert hasattr(lib, n), n
    assert lib.iif_abi_version() == 1
    assert b"alignment" in lib.iif_error_string(-2)
    assert lib.iif_gemm_ws_bytes(0, 2048, 1000) == 0
    assert lib.iif_gemm_ws_bytes(256, 2048, 1000) > 0
    assert lib.iif_hist_images_dedup_ws_bytes(100, 10) == 10 * 4 * 4
    # no stray exports: only the iif_* C ABI is visible
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(exported) == names


def test_head_args_struct_layout(tmp_path):
    """ctypes HeadArgs == struct iif_head_args as gcc lays it out."""
    from iif_b200 import _lib
    fields = [f[0] for f in _lib.HeadArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(){",
            'printf("%zu\\n", sizeof(iif_head_args));']
    prog += [f'printf("%zu\\n", offsetof(iif_head_args, {f}));' for f in fields]
    prog += ["return 0;}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(c), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(_lib.HeadArgs)
    assert vals[1:] == [getattr(_lib.HeadArgs, f).offset for f in fields]


def test_ndtri_host_matches_scipy(tmp_path):
    """csrc/ndtri.h (the Cephes restatement shared with the CUDA weight kernel) against
    scipy.special.ndtri -- the function the reference calls (classification/custom.py:4,20)."""
    from scipy.special import ndtri
    c = tmp_path / "nd.c"
    c.write_text(f'#include "{os.path.join(ROOT, "iif_b200", "csrc", "ndtri.h")}"\n'
                 "double nd(double y){return iif_ndtri(y);}\n")
    so = tmp_path / "nd.so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(c), "-o", str(so), "-lm"], check=True)
    f = ctypes.CDLL(str(so)).nd
    f.restype, f.argtypes = ctypes.c_double, [ctypes.c_double]
    rng = np.random.default_rng(0)
    ys = np.concatenate([rng.uniform(0, 1, 4000), 10.0 ** rng.uniform(-300, -1, 2000),
                         1 - 10.0 ** rng.uniform(-15, -1, 1000), [0.5, 0.1353352832366127, 0.8646647167633873]])
    got = np.array([f(float(y)) for y in ys])
    ref = ndtri(ys)
    np.testing.assert_allclose(got, ref, rtol=4e-15, atol=1e-300)
    assert f(0.0) == -np.inf and f(1.0) == np.inf and np.isnan(f(1.5))


def test_mmdet_reduction_contract():
    """losses/utils.py:42-55 folded into (scale, reduce?)."""
    from iif_b200.mmdet import _resolve, _empty_result
    assert _resolve("mean", None, 1.0, 8) == (1 / 8, True)
    assert _resolve("sum", None, 2.0, 8) == (2.0, True)
    assert _resolve("none", None, 1.0, 8) == (1.0, False)
    assert _resolve("mean", 4.0, 0.5, 8) == (0.125, True)
    assert _resolve("none", 4.0, 1.0, 8) == (1.0, False)
    with pytest.raises(ValueError, match="avg_factor can not be used"):
        _resolve("sum", 4.0, 1.0, 8)
    e = torch.zeros(0, 5)
    assert _empty_result(e, "none", None).shape == (0,)
    assert torch.isnan(_empty_result(e, "mean", None))             # torch: mean of empty
    assert float(_empty_result(e, "mean", 3.0)) == 0.0 and float(_empty_result(e, "sum", None)) == 0.0


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors -- never route through a CPU path."""
    from iif_b200 import ops
    z = torch.zeros(4, 8)
    y = torch.zeros(4, dtype=torch.int64)
    for call in (lambda: ops.softmax_ce(z, None, y), lambda: ops.sigmoid_bce(z, y),
                 lambda: ops.scaled_activation(z, None, True), lambda: ops.hist_labels(y, 8),
                 lambda: ops.linear_fwd(z, torch.zeros(3, 8)), lambda: ops.colsum(z)):
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            call()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under iif_b200/ may import it."""
    pkg = os.path.join(ROOT, "iif_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_bench_reads_committed_ncu_traffic():
    """bench.py attaches the DRAM traffic of the dominant kernel from the committed ncu capture (profiles/)."""
    import bench
    t, src = bench.ncu_traffic("loss_linear_bwd_bf16", (256, 2048, 1000))
    assert t is not None and 1e6 < t < 1e8 and "profiles/" in src
    assert bench.ncu_traffic("loss_linear_bwd_bf16", (512, 2048, 1000)) == (None, None)
    p = bench.peaks()
    assert p["hbm"] > 1000 and p["tf_burst"] >= p["tf_sust"] > 100
    t2, src2 = bench.ncu_traffic("head_step_fused_bf16", (256, 2048, 1000))
    assert t2 is not None and 4e6 < t2 < 2e7 and "r2_ncu_head_full_summary" in src2


def test_one_launch_step_routing(monkeypatch):
    """Host-only plan of the one-launch step (iif_debug_fused_plan): the head shapes of BASELINE.json qualify, shapes
    with more work per CTA go to the multi-launch chain (IIF_EUNSUPPORTED = -3), the environment overrides lift the
    limits, and the hard limits (B <= 2048, C <= 4096) stay."""
    from iif_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int * 12)()
    plan = lambda B, D, C: lib.iif_debug_fused_plan(B, D, C, 1, 148, out)
    for shape in [(256, 2048, 1000), (256, 2048, 365), (128, 64, 10), (512, 1024, 1204), (1, 8, 1)]:
        assert plan(*shape) == 0, shape
    assert plan(256, 2048, 1000) == 0 and list(out)[:6] == [148, 8, 128, 2, 64, 128]
    for shape in [(1024, 1024, 1204), (2048, 1024, 1204), (512, 2048, 1000), (2048, 2048, 1000)]:
        assert plan(*shape) == -3, shape
    monkeypatch.setenv("IIF_B200_FUSED_MAX_ROW_PASSES", "0")
    monkeypatch.setenv("IIF_B200_FUSED_MAX_WORK", "0")
    for shape in [(1024, 1024, 1204), (2048, 1024, 1204), (512, 512, 4096)]:
        assert plan(*shape) == 0, shape
    assert plan(4096, 1024, 1204) == -3 and plan(256, 1024, 4097) == -3


def test_one_launch_plan_fuzz(monkeypatch):
    """The host-side planner of the one-launch step over random shapes / SM counts (routing limits lifted): it either
    declines (IIF_EUNSUPPORTED) or returns a plan whose grid fits the device, whose split counts are within the
    kernel's bounds and whose loss-row geometry covers the class count."""
    import random
    from iif_b200 import _lib
    monkeypatch.setenv("IIF_B200_FUSED_MAX_ROW_PASSES", "0")
    monkeypatch.setenv("IIF_B200_FUSED_MAX_WORK", "0")
    lib = _lib.load()
    out = (ctypes.c_int * 12)()
    rnd = random.Random(7)
    accepted = 0
    for _ in range(3000):
        B = rnd.choice([1, 2, 7, 128, 129, 256, 777, 1024, 2048, 2049, rnd.randint(1, 3000)])
        D = rnd.choice([1, 8, 63, 64, 65, 520, 1024, 2048, 8192, rnd.randint(1, 20000)])
        C = rnd.choice([1, 10, 128, 129, 365, 1000, 1203, 1204, 4096, 4097, rnd.randint(1, 5000)])
        sms = rnd.choice([148, 132, 64, 16, 1])
        need_dx = rnd.randint(0, 1)
        rc = lib.iif_debug_fused_plan(B, D, C, need_dx, sms, out)
        assert rc in (0, -3), (rc, B, D, C)
        if B > 2048 or C > 4096:
            assert rc == -3
        if rc == 0:
            accepted += 1
            grid, f_splits, f_items, dx_splits, dx_items, dw_items, _, tpr, ne, row_blocks = list(out)[:10]
            assert 1 <= grid <= sms and 1 <= f_splits <= 8 and f_items >= 1 and dw_items >= 1 and row_blocks >= 1
            assert (dx_items == 0) if not need_dx else (1 <= dx_splits <= 8 and dx_items >= 1)
            assert tpr in (128, 256) and ne in (8, 16) and tpr * ne >= C
    assert accepted > 1500
