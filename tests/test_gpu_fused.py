"""GPU parity of the ONE-launch persistent head step (csrc/head_fused.cu) through the C ABI
(iif_head_fwd_bwd_bf16 -> ops.HeadStep): against the float64 oracle evaluated on the same bf16-rounded
operands, against the multi-launch chain of the same library, and its own invariants (counters re-arm,
results reproduce bit for bit run to run)."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import TOL_BF16, rel_err, head_inputs, iif_row, bf16_round

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _one_launch_for_every_shape(monkeypatch):
    """The library routes many-rows-per-CTA shapes (LVIS) to the multi-launch chain for speed; parity of the
    one-launch step is tested on them all the same."""
    monkeypatch.setenv("IIF_B200_FUSED_MAX_ROW_PASSES", "0")
    monkeypatch.setenv("IIF_B200_FUSED_MAX_WORK", "0")
DEV = "cuda:0"
BF = torch.bfloat16


def T(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().float().cpu().numpy()


SHAPES = [
    (256, 2048, 1000),    # ImageNet-LT R50 (the bench shape)
    (256, 2048, 365),     # Places-LT R152: C % 4 != 0, ragged class tile
    (1024, 1024, 1204),   # LVIS bbox head
    (2048, 1024, 1204),   # LVIS as the reference runs it
    (128, 64, 10),        # CIFAR-10-LT sized
    (1, 8, 1),            # degenerate
    (77, 520, 1203),      # ragged everything
    (300, 256, 8),
    (512, 512, 4096),     # widest supported class count
    (129, 72, 130),       # one row / a few columns past a tile edge
]


@pytest.mark.parametrize("B,D,C", SHAPES)
def test_fused_step_vs_oracle_and_chain(B, D, C):
    from iif_b200.ops import HeadStep
    x, w, b, counts, y = head_inputs(B, D, C, seed=B + 3 * C)
    iif = iif_row(counts, "smooth")
    y[::7] = -100                                           # ignored rows
    cw = np.random.default_rng(1).uniform(0.5, 2.0, C).astype(np.float32)
    sw = np.random.default_rng(2).uniform(0.0, 2.0, B).astype(np.float32)
    ref = ho.head_fwd_bwd(bf16_round(x), bf16_round(w), b, iif, y, class_weight=cw, sample_weight=sw)
    args = (T(x, BF), T(w, BF), T(b), T(iif).reshape(-1), T(y))
    kw = dict(class_weight=T(cw), sample_weight=T(sw))
    f = HeadStep(B, D, C, DEV, want_acc=True, dx_bf16=False)
    u = HeadStep(B, D, C, DEV, want_acc=True, dx_bf16=False, persistent=False, fused_loss=False)
    f.bind(*args, **kw); u.bind(*args, **kw)
    assert f.launches_per_step == 1, "this shape is expected to qualify for the one-launch step"
    assert u.launches_per_step == 3
    for _ in range(3):                                      # counters re-arm: several steps on one workspace
        lf, lu = f.launch(), u.launch()
    torch.cuda.synchronize()
    # integers: bit-exact on the step's own logits
    assert np.array_equal(f.argmax.cpu().numpy(), ho.argmax_first(N(f.z)))
    assert np.array_equal(f.rank.cpu().numpy(), ho.label_rank(N(f.z), y))
    r = f.rank.cpu().numpy()
    assert f.acc_counts.cpu().tolist() == [int((r < 1).sum()), int((r < 5).sum())]
    # floats: vs the oracle on the same bf16 operands (6e-3: bf16 storage of dZ) and vs the 3-launch chain
    assert float(lf) == pytest.approx(ref["loss"], rel=1e-4)
    assert rel_err(N(f.z), ref["z"]) < 2e-5
    assert rel_err(N(f.loss_i), ref["loss_i"] / B) < 1e-4        # the kernel's loss_i carries the 1/B scale
    assert rel_err(N(f.dw), ref["dw"]) < 6e-3 and rel_err(N(f.dx), ref["dx"]) < 6e-3
    assert rel_err(N(f.db), ref["db"]) < 6e-3
    assert float(lf) == pytest.approx(float(lu), rel=1e-5)
    assert rel_err(N(f.z), N(u.z)) < 1e-5                    # split counts may differ: fp32 summation order
    assert rel_err(N(f.dw), N(u.dw)) < 2e-2 and rel_err(N(f.dx), N(u.dx)) < 2e-2
    if C % 4:                                                # the ragged last float4 group writes exact zeros past C
        pad = f.dz[:, C:(C + 3) // 4 * 4]
        assert torch.equal(pad, torch.zeros_like(pad))


@pytest.mark.parametrize("B,D,C", [(256, 2048, 1000), (200, 512, 365)])
def test_fused_step_options(B, D, C):
    """No dX (frozen backbone), no bias / iif / db, bf16 dX: every optional pointer of iif_head_args."""
    from iif_b200.ops import HeadStep
    x, w, b, counts, y = head_inputs(B, D, C, seed=5)
    ref = ho.head_fwd_bwd(bf16_round(x), bf16_round(w), None, None, y)
    f = HeadStep(B, D, C, DEV, need_dx=False, need_db=False)
    f.bind(T(x, BF), T(w, BF), None, None, T(y))
    assert f.launches_per_step == 1
    l = f.launch()
    torch.cuda.synchronize()
    assert float(l) == pytest.approx(ref["loss"], rel=1e-4)
    assert rel_err(N(f.dw), ref["dw"]) < 6e-3 and rel_err(N(f.z), ref["z"]) < 2e-5
    g = HeadStep(B, D, C, DEV, dx_bf16=True)
    g.bind(T(x, BF), T(w, BF), T(b), T(iif_row(counts, "raw")).reshape(-1), T(y))
    ref2 = ho.head_fwd_bwd(bf16_round(x), bf16_round(w), b, iif_row(counts, "raw"), y)
    g.launch()
    torch.cuda.synchronize()
    assert rel_err(N(g.dx), ref2["dx"]) < 1e-2 and rel_err(N(g.db), ref2["db"]) < 6e-3


def test_fused_step_reproducible_over_rotating_sets():
    """600 eager steps over rotating input sets sharing ONE workspace: each set keeps reproducing its first
    result bit for bit (fixed summation orders everywhere; counters re-armed by the tail of every launch)."""
    from iif_b200.ops import HeadStep
    B, D, C = 256, 2048, 1000
    ws = torch.zeros(int(HeadStep(B, D, C, DEV).ws_bytes), dtype=torch.uint8, device=DEV)
    sets = []
    for seed in range(6):
        x, w, b, counts, y = head_inputs(B, D, C, seed=seed)
        hs = HeadStep(B, D, C, DEV, ws=ws, want_acc=True)
        hs.bind(T(x, BF), T(w, BF), T(b), T(iif_row(counts, "smooth")).reshape(-1), T(y))
        assert hs.launches_per_step == 1
        sets.append(hs)
    first = []
    for hs in sets:
        hs.launch()
        torch.cuda.synchronize()
        first.append((float(hs.loss), hs.dw.clone(), hs.dx.clone(), hs.z.clone(), hs.acc_counts.clone()))
    for i in range(600):
        sets[i % 6].launch()
    torch.cuda.synchronize()
    for hs, (l0, dw0, dx0, z0, a0) in zip(sets, first):
        assert float(hs.loss) == l0 and torch.equal(hs.dw, dw0) and torch.equal(hs.dx, dx0) and torch.equal(hs.z, z0)
        assert torch.equal(hs.acc_counts, a0)


def test_fused_step_in_cuda_graph():
    """The bench path: the one launch captured in a CUDA graph and replayed."""
    from iif_b200.ops import HeadStep
    B, D, C = 256, 2048, 1000
    x, w, b, counts, y = head_inputs(B, D, C, seed=11)
    hs = HeadStep(B, D, C, DEV)
    hs.bind(T(x, BF), T(w, BF), T(b), T(iif_row(counts, "smooth")).reshape(-1), T(y))
    side = torch.cuda.Stream(DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        hs.launch()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref = (float(hs.loss), hs.dw.clone(), hs.dx.clone())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        hs.launch()
        hs.launch()
    hs.dw.zero_(); hs.dx.zero_()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    assert float(hs.loss) == ref[0] and torch.equal(hs.dw, ref[1]) and torch.equal(hs.dx, ref[2])


@pytest.mark.parametrize("B,D,C", [(65536, 2048, 1000), (16384, 512, 10000)])
def test_head_step_full_size_sampled_rows(B, D, C):
    """BASELINE.json's sweep sizes through the one C-ABI call (the multi-launch chain at these sizes): sampled rows of
    Z / dZ / dX and sampled classes of dW / db against the float64 oracle, plus whole-array identities on the step's own
    outputs (db = column sums of dZ, rows of dZ / iif sum to ~0, loss = sum of the per-row losses)."""
    from iif_b200.ops import HeadStep
    g = torch.Generator(device="cpu").manual_seed(B + C)
    counts = np.maximum((1280 * 0.01 ** (np.arange(C) / (C - 1.0))).astype(np.int64), 1)
    iif = iif_row(counts, "smooth")
    x = torch.randn(B, D, generator=g).to(DEV).to(BF)
    w = ((torch.rand(C, D, generator=g) * 2 - 1) / D ** 0.5).to(DEV).to(BF)
    b = torch.full((C,), 0.01, device=DEV)
    y = torch.multinomial(torch.from_numpy(counts / counts.sum()), B, replacement=True, generator=g).to(DEV)
    hs = HeadStep(B, D, C, DEV, dx_bf16=False, want_acc=True)
    hs.bind(x, w, b, T(iif).reshape(-1), y)
    loss = hs.launch()
    torch.cuda.synchronize()
    rows = np.unique(np.concatenate([[0, 127, 128, B - 1], np.random.default_rng(3).integers(0, B, 60)]))
    xr, wn, yn = N(x[rows]).astype(np.float64), N(w).astype(np.float64), y.cpu().numpy()
    ref = ho.head_fwd_bwd(xr, wn, N(b), iif, yn[rows], scale=1.0 / B)
    assert rel_err(N(hs.z[rows]), ref["z"]) < 2e-5
    assert rel_err(N(hs.dz[rows, :C]), ref["dz"]) < 6e-3            # bf16 storage of dZ
    assert rel_err(N(hs.dx[rows]), ref["dz"] @ wn) < 6e-3
    assert rel_err(N(hs.loss_i[rows]), ref["loss_i"] / B) < 1e-4
    assert np.array_equal(hs.argmax[rows].cpu().numpy(), ho.argmax_first(N(hs.z[rows])))
    assert np.array_equal(hs.rank[rows].cpu().numpy(), ho.label_rank(N(hs.z[rows]), yn[rows]))
    # whole-array identities on the step's own outputs
    dz = hs.dz[:, :C].float()
    assert float(loss) == pytest.approx(float(hs.loss_i.double().sum()), rel=1e-5)
    assert rel_err(N(hs.db), N(dz.double().sum(0))) < 2e-3
    unscaled = dz.double() / T(iif).reshape(1, -1).double()         # dZ_c = iif_c (p_c - [c = y]) / B
    assert float(unscaled.sum(1).abs().max()) < 0.05 / B            # rows of p - onehot sum to 0 up to bf16 rounding
    cls = np.unique(np.concatenate([[0, 127, 128, C - 1], np.random.default_rng(4).integers(0, C, 28)]))
    dw_ref = dz[:, cls].double().T @ x.double()                     # the kernel's own bf16 dZ, fp64 accumulation
    assert rel_err(N(hs.dw[cls]), dw_ref.cpu().numpy()) < 2e-3
    r = hs.rank.cpu().numpy()
    assert hs.acc_counts.cpu().tolist() == [int((r < 1).sum()), int((r < 5).sum())]
