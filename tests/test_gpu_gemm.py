"""GPU parity: fc_cls GEMMs (forward, dX, dW) through the C ABI.
fp32 (FFMA) mode against the float64 oracle at 1e-5; bf16 tcgen05 mode against the oracle evaluated
on the SAME bf16-rounded operands (tight: only fp32 accumulation order differs) and against the
fp32 operands at the 2e-2 bar of BASELINE.json."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import TOL_F32, TOL_BF16, rel_err, head_inputs, bf16_round

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    (128, 64, 10),      # CIFAR-10-LT R32
    (256, 2048, 1000),  # ImageNet-LT R50
    (256, 2048, 365),   # Places-LT R152
    (1024, 1024, 1204), # LVIS bbox head
    (1, 8, 1), (3, 72, 5), (5, 100, 7), (130, 200, 129), (77, 520, 1203), (512, 512, 1000),
]


def T(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from iif_b200 import ops as o
    return o


@pytest.mark.parametrize("B,D,C", SHAPES)
def test_gemm_f32(ops, B, D, C):
    x, w, b, counts, y = head_inputs(B, D, C, seed=B + C)
    rng = np.random.default_rng(1)
    dz = (rng.standard_normal((B, C)) / B).astype(np.float32)
    s = rng.uniform(0.5, 7.0, C).astype(np.float32)
    z, zs = ops.linear_fwd(T(x), T(w), T(b), T(s), want_raw=True, want_scaled=True)
    zr = ho.linear_fwd(x, w, b)
    assert rel_err(N(z), zr) < TOL_F32
    assert rel_err(N(zs), zr * s[None, :]) < TOL_F32
    dxr, dwr, dbr = ho.linear_bwd(dz, x, w)
    assert rel_err(N(ops.linear_bwd_dx(T(dz), T(w))), dxr) < TOL_F32
    assert rel_err(N(ops.linear_bwd_dw(T(dz), T(x))), dwr) < TOL_F32
    alpha = T(np.float32(0.25))
    assert rel_err(N(ops.linear_bwd_dw(T(dz), T(x), alpha=alpha)), dwr * 0.25) < TOL_F32
    assert rel_err(N(ops.colsum(T(dz))), dbr) < TOL_F32


@pytest.mark.parametrize("B,D,C", SHAPES)
def test_gemm_bf16_tcgen05(ops, B, D, C):
    x, w, b, counts, y = head_inputs(B, D, C, seed=B + C + 1)
    rng = np.random.default_rng(2)
    dz = (rng.standard_normal((B, C)) / B).astype(np.float32)
    s = rng.uniform(0.5, 7.0, C).astype(np.float32)
    xb, wb, dzb = bf16_round(x), bf16_round(w), bf16_round(dz)
    bf = torch.bfloat16
    z, zs = ops.linear_fwd(T(x, bf), T(w, bf), T(b), T(s), want_raw=True, want_scaled=True)
    zr = ho.linear_fwd(xb, wb, b)
    assert rel_err(N(z), zr) < 2e-5                                   # same operands, fp32 accumulate
    assert rel_err(N(zs), zr * s[None, :]) < 2e-5
    assert rel_err(N(z), ho.linear_fwd(x, w, b)) < TOL_BF16           # the BASELINE bf16-GEMM bar
    dxr, dwr, _ = ho.linear_bwd(dzb, xb, wb)
    dzt = ops.scale_rows(T(dz), None, bf16=True, pad_ld=True)
    assert np.array_equal(N(dzt), dzb)
    dx = ops.linear_bwd_dx(dzt, T(w, bf))
    assert rel_err(N(dx), dxr) < 2e-5
    dxh = ops.linear_bwd_dx(dzt, T(w, bf), out_bf16=True)
    assert dxh.dtype == bf and rel_err(N(dxh), dxr) < 8e-3
    dw = ops.linear_bwd_dw(dzt, T(x, bf), alpha=T(np.float32(2.0)))
    assert rel_err(N(dw), 2.0 * dwr) < 2e-5
    dx32, dw32, _ = ho.linear_bwd(dz, x, w)
    assert rel_err(N(dx), dx32) < TOL_BF16 and rel_err(N(dw), 2.0 * dw32) < TOL_BF16


@pytest.mark.parametrize("B,D,C", SHAPES + [(2048, 512, 1000), (64, 2048, 10000)])
def test_linear_bwd_grouped(ops, B, D, C):
    """dX + dW + db in one launch (db on the tensor cores via the ones tile) vs the oracle on the
    same bf16 operands; alpha read from the device; dX optional (frozen backbone)."""
    x, w, b, counts, y = head_inputs(B, D, C, seed=B + C + 2)
    rng = np.random.default_rng(3)
    dz = (rng.standard_normal((B, C)) / B).astype(np.float32)
    xb, wb, dzb = bf16_round(x), bf16_round(w), bf16_round(dz)
    bf = torch.bfloat16
    dzt = ops.scale_rows(T(dz), None, bf16=True, pad_ld=True)
    dxr, dwr, dbr = ho.linear_bwd(dzb, xb, wb)
    dx, dw, db = ops.linear_bwd(dzt, T(x, bf), T(w, bf), alpha=T(np.float32(0.5)))
    assert rel_err(N(dx), 0.5 * dxr) < 2e-5
    assert rel_err(N(dw), 0.5 * dwr) < 2e-5
    assert rel_err(N(db), 0.5 * dbr) < 2e-5
    dx2, dw2, db2 = ops.linear_bwd(dzt, T(x, bf), T(w, bf), need_dx=False, need_db=False)
    assert dx2 is None and db2 is None and rel_err(N(dw2), dwr) < 2e-5
    dx3, dw3, db3 = ops.linear_bwd(dzt, T(x, bf), T(w, bf), dx_bf16=True)
    assert dx3.dtype == bf and rel_err(N(dx3), dxr) < 8e-3 and rel_err(N(dw3), dwr) < 2e-5


def test_gemm_bf16_deterministic_and_ticket_reset(ops):
    """Split-K partial sums are added in split order by the last CTA: run-to-run bit-identical, and
    the self-resetting tickets allow back-to-back launches on one workspace."""
    x, w, b, _, _ = head_inputs(256, 2048, 1000, seed=9)
    bf = torch.bfloat16
    xt, wt, bt = T(x, bf), T(w, bf), T(b)
    z0, _ = ops.linear_fwd(xt, wt, bt)
    for _ in range(5):
        z1, _ = ops.linear_fwd(xt, wt, bt)
        assert torch.equal(z0, z1)


def test_gemm_bf16_linearity_full_size(ops):
    """Sweep-size GEMM (16384 x 2048 x 1000) through a size-independent property: linearity in X
    and agreement of a random 64-row sample with the oracle."""
    torch.manual_seed(0)
    B, D, C = 16384, 2048, 1000
    bf = torch.bfloat16
    x1 = torch.randn(B, D, device=DEV).to(bf)
    w = (torch.rand(C, D, device=DEV) * 2 - 1).mul_(D ** -0.5).to(bf)
    z1, _ = ops.linear_fwd(x1, w)
    z2, _ = ops.linear_fwd(x1 * 2, w)                 # exact in bf16 (power of two)
    assert torch.equal(z2, z1 * 2)
    idx = torch.randint(0, B, (64,), device=DEV)
    ref = ho.linear_fwd(N(x1[idx]), N(w))
    assert rel_err(N(z1[idx]), ref) < 2e-5


def test_linear_bwd_long_k_additive_split(ops):
    """Big batch: dW = dZ^T X has few tiles and a very long K (= B).  Its K is split over CTAs that reduce-add
    into the zeroed output through the TMA unit (no rendezvous: the grid is far larger than the resident
    capacity), next to thousands of dX tiles in the same launch; db through fp32 atomics."""
    B, D, C = 16384, 512, 1000
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, D)).astype(np.float32)
    w = (rng.uniform(-1, 1, (C, D)) / np.sqrt(D)).astype(np.float32)
    dz = (rng.standard_normal((B, C)) / B).astype(np.float32)
    xb, wb, dzb = bf16_round(x), bf16_round(w), bf16_round(dz)
    bf = torch.bfloat16
    dzt = ops.scale_rows(T(dz), None, bf16=True, pad_ld=True)
    dxr, dwr, dbr = ho.linear_bwd(dzb, xb, wb)
    for _ in range(2):                                    # the output is re-zeroed by every call
        dx, dw, db = ops.linear_bwd(dzt, T(x, bf), T(w, bf))
    assert rel_err(N(dx), dxr) < 2e-5
    assert rel_err(N(dw), dwr) < 2e-5
    assert rel_err(N(db), dbr) < 2e-5
    dw2 = ops.linear_bwd_dw(dzt, T(x, bf), alpha=T(np.float32(0.5)))
    assert rel_err(N(dw2), 0.5 * dwr) < 2e-5
