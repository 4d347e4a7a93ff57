"""GPU parity: fused softmax-CE / sigmoid-BCE / activation kernels (through the C ABI) against the
float64 oracle and the golden vectors frozen from the reference's own Python."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import TOL_F32, TOL_BF16, rel_err, head_inputs, iif_row, lt_counts, lt_labels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().float().cpu().numpy() if t.dtype == torch.bfloat16 else t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from iif_b200 import ops as o
    return o


# C values cover every dispatch bucket of the row kernel (vectorised and scalar: C % 4 != 0)
@pytest.mark.parametrize("B,C", [(1, 1), (7, 10), (128, 10), (33, 37), (256, 128), (64, 365), (256, 1000),
                                 (512, 1204), (96, 1203), (40, 2048), (24, 4099), (16, 10000), (8, 16384),
                                 (4, 32768), (300, 513)])
def test_softmax_ce_vs_oracle(ops, B, C):
    rng = np.random.default_rng(B * 131 + C)
    z = (rng.standard_normal((B, C)) * 3).astype(np.float32)
    counts = lt_counts(C)
    iif = iif_row(counts, "smooth")
    y = rng.integers(0, C, size=B).astype(np.int64)
    if B > 4:
        y[1] = -100          # ignored row
    cw = rng.uniform(0.5, 1.5, C).astype(np.float32)
    sw = rng.uniform(0.0, 2.0, B).astype(np.float32)
    scale = 1.0 / B
    r = ops.softmax_ce(T(z), T(iif), T(y), class_weight=T(cw), sample_weight=T(sw), scale=scale,
                       want_dz_f32=True, want_dz_bf16=True, want_acc=True, want_lse=True)
    li, dz, lse = ho.softmax_ce(z, iif, y, cw, sw, -100)
    assert rel_err(N(r["loss_i"]), li * scale) < TOL_F32
    assert abs(float(r["loss_sum"]) - li.sum() * scale) <= TOL_F32 * abs(li.sum() * scale) + 1e-30
    assert rel_err(N(r["dz_f32"]), dz * scale) < TOL_F32
    assert rel_err(N(r["dz_bf16"])[:, :C], dz * scale) < 8e-3          # bf16 storage of dZ
    assert rel_err(N(r["lse"]), lse) < TOL_F32
    # integer outputs: bit-exact (argmax on RAW logits, rank of the label)
    assert np.array_equal(N(r["argmax"]), ho.argmax_first(z))
    assert np.array_equal(N(r["rank"]), ho.label_rank(z, y))
    rk = ho.label_rank(z, y)
    assert N(r["acc_counts"]).tolist() == [int((rk < 1).sum()), int((rk < 5).sum())]


def test_softmax_ce_plain_and_strided(ops):
    """No iif / no weights; z given as a column slice of a wider buffer (ld > C)."""
    rng = np.random.default_rng(5)
    B, C = 50, 1000
    zbuf = T((rng.standard_normal((B, C + 24)) * 2).astype(np.float32))
    z = zbuf[:, :C]
    y = rng.integers(0, C, size=B).astype(np.int64)
    r = ops.softmax_ce(z, None, T(y), scale=1.0)
    li, dz, _ = ho.softmax_ce(N(z), None, y)
    assert rel_err(N(r["loss_i"]), li) < TOL_F32
    assert rel_err(N(r["dz_f32"]), dz) < TOL_F32


def test_softmax_ce_extremes(ops):
    """KAT of the reference's own suite (tests/test_metrics/test_losses.py:8-32): 200.0 and 40.0;
    plus ties (first index wins), negative and inf IIF weights."""
    z = T(np.array([[100.0, -100.0]], np.float32))
    y = T(np.array([1], np.int64))
    r = ops.softmax_ce(z, None, y, scale=1.0)
    assert float(r["loss_sum"]) == pytest.approx(200.0, rel=1e-6)
    r = ops.softmax_ce(z, None, y, class_weight=T(np.array([0.8, 0.2], np.float32)), scale=1.0)
    assert float(r["loss_sum"]) == pytest.approx(40.0, rel=1e-6)
    zt = np.zeros((3, 12), np.float32)
    zt[1, 4] = zt[1, 9] = 2.0
    r = ops.softmax_ce(T(zt), None, T(np.array([0, 9, 11], np.int64)), want_acc=True, scale=1.0)
    assert N(r["argmax"]).tolist() == [0, 4, 0]
    assert N(r["rank"]).tolist() == [0, 1, 11]
    # negative weights exist in the reference tables (coco_files/idf_91.csv:3)
    rng = np.random.default_rng(1)
    zz = rng.standard_normal((9, 81)).astype(np.float32)
    s = rng.uniform(-0.5, 3.0, (1, 81)).astype(np.float32)
    yy = rng.integers(0, 81, 9).astype(np.int64)
    r = ops.softmax_ce(T(zz), T(s), T(yy), scale=1.0)
    li, dz, _ = ho.softmax_ce(zz, s, yy)
    assert rel_err(N(r["loss_i"]), li) < TOL_F32 and rel_err(N(r["dz_f32"]), dz) < TOL_F32
    # ignored rows give exact zeros
    r = ops.softmax_ce(T(zz), T(s), T(np.full(9, -100, np.int64)), scale=1.0)
    assert float(r["loss_sum"]) == 0.0 and not N(r["dz_f32"]).any()


@pytest.mark.parametrize("tag", ["raw_mean", "smooth_mean", "rel_mean", "normit_mean", "gombit_mean", "base2_mean",
                                 "base10_mean", "raw_sum", "smooth_none"])
def test_softmax_ce_cls_golden(ops, golden, tag):
    """Against the outputs of the unmodified classification/custom.py:IIFLoss."""
    g = golden("cls_iif")
    v, red = tag.rsplit("_", 1)
    B = g["z"].shape[0]
    scale = 1.0 / B if red == "mean" else 1.0
    r = ops.softmax_ce(T(g["z"]), T(g[f"iif_{v}"]), T(g["y"]), scale=scale)
    if red == "none":
        assert rel_err(N(r["loss_i"]), g[f"loss_{tag}"]) < TOL_F32
    else:
        assert float(r["loss_sum"]) == pytest.approx(float(g[f"loss_{tag}"]), rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), g[f"dz_{tag}"]) < TOL_F32


def test_softmax_ce_mmdet_golden(ops, golden):
    """Against the unmodified mmdet IIFLoss (all 14 CSV columns, avg_factor + label weights)."""
    g = golden("mmdet_iif")
    base = ["smooth", "raw", "prob", "normit", "gombit", "base2", "base10"]
    cols = base + [c + "_obj" for c in base]
    assert len(cols) == 14
    for col in cols:
        r = ops.softmax_ce(T(g["z"]), T(g[f"iif_{col}"]), T(g["y"]), sample_weight=T(g["w"]),
                           scale=1.0 / float(g["avg_factor"]))
        assert float(r["loss_sum"]) == pytest.approx(float(g[f"loss_{col}_avg"]), rel=TOL_F32)
        assert rel_err(N(r["dz_f32"])[:8], g[f"dz_{col}_avg"]) < TOL_F32


@pytest.mark.parametrize("softmax", [True, False])
@pytest.mark.parametrize("B,C", [(32, 1204), (5, 10), (64, 1000), (3, 4099)])
def test_scaled_activation(ops, B, C, softmax):
    rng = np.random.default_rng(C)
    z = (rng.standard_normal((B, C)) * 2).astype(np.float32)
    s = iif_row(lt_counts(C), "raw")
    y = rng.integers(0, C, B).astype(np.int64)
    out, am, rk = ops.scaled_activation(T(z), T(s), softmax=softmax, label=T(y), want_pred=True)
    a32 = (z * s).astype(np.float32)       # the reference multiplies in fp32 (custom.py:38)
    if softmax:
        assert rel_err(N(out), ho.softmax_activation(z, s)) < TOL_F32
    else:
        assert np.array_equal(N(out), a32)  # one fp32 multiply: bit-exact
    assert np.array_equal(N(am), ho.argmax_first(a32))
    assert np.array_equal(N(rk), ho.label_rank(a32, y))


def test_activation_golden(ops, golden):
    g = golden("mmdet_iif")
    out, _, _ = ops.scaled_activation(T(g["z"]), T(g["iif_raw"]), softmax=True)
    assert rel_err(N(out), g["act_raw"]) < TOL_F32
    c = golden("cls_iif")
    out, _, _ = ops.scaled_activation(T(c["z"]), T(c["iif_smooth"]), softmax=False)
    assert np.array_equal(N(out), c["infer_smooth"])


@pytest.mark.parametrize("B,C", [(24, 1203), (64, 37), (256, 1000), (5, 1), (130, 1204), (17, 4100)])
def test_sigmoid_bce_vs_oracle(ops, B, C):
    rng = np.random.default_rng(B + C)
    z = (rng.standard_normal((B, C)) * 4).astype(np.float32)
    y = rng.integers(0, C + 1, size=B).astype(np.int64)     # label == C: background row, all-zero targets
    if B > 4:
        y[2] = 255 if C < 255 else -100
    ign = 255 if C < 255 else -100
    pw = rng.uniform(0.5, 2.0, C).astype(np.float32)
    sw = rng.uniform(0, 2, B).astype(np.float32)
    scale = 1.0 / 7.0
    r = ops.sigmoid_bce(T(z), T(y), pos_weight=T(pw), sample_weight=T(sw), ignore_index=ign, scale=scale,
                        want_elem=True, want_dz_f32=True, want_dz_bf16=True)
    loss, dz = ho.sigmoid_bce_mmdet(z, y, sw, pw, ign)
    assert rel_err(N(r["loss_elem"]), loss * scale) < TOL_F32
    assert rel_err(N(r["loss_i"]), loss.sum(1) * scale) < TOL_F32
    assert float(r["loss_sum"]) == pytest.approx(loss.sum() * scale, rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), dz * scale) < TOL_F32
    assert rel_err(N(r["dz_bf16"])[:, :C], dz * scale) < 8e-3


def test_sigmoid_bce_golden(ops, golden):
    g = golden("mmdet_bce")
    af = float(g["avg_factor"])
    r = ops.sigmoid_bce(T(g["z"]), T(g["y"]), sample_weight=T(g["w"]), scale=1.0 / af)
    assert float(r["loss_sum"]) == pytest.approx(float(g["loss_avg"]), rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), g["dz_avg"]) < TOL_F32
    r = ops.sigmoid_bce(T(g["z"]), T(g["y"]), scale=1.0 / g["z"].size)
    assert float(r["loss_sum"]) == pytest.approx(float(g["loss_plain"]), rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), g["dz_plain"]) < TOL_F32
    c = golden("cls_bce")
    B, C = c["z"].shape
    r = ops.sigmoid_bce(T(c["z"]), T(c["y"]), col_weight=T(c["weights"]), scale=1.0 / (B * C))
    assert float(r["loss_sum"]) == pytest.approx(float(c["loss_w_mean"]), rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), c["dz_w_mean"]) < TOL_F32


def test_scale_rows_and_colsum(ops):
    rng = np.random.default_rng(3)
    for rows, cols in [(256, 1000), (7, 13), (1024, 1204), (1, 5)]:
        x = rng.standard_normal((rows, cols)).astype(np.float32)
        g = rng.standard_normal(rows).astype(np.float32)
        assert np.array_equal(N(ops.scale_rows(T(x), T(g))), x * g[:, None]) or rows == 1
        assert np.array_equal(N(ops.scale_rows(T(x), T(np.float32(0.5)))), x * np.float32(0.5))
        xb = N(ops.scale_rows(T(x), None, bf16=True))
        assert np.array_equal(xb, T(x).to(torch.bfloat16).float().cpu().numpy())
        assert rel_err(N(ops.colsum(T(x))), x.astype(np.float64).sum(0)) < TOL_F32


def test_full_size_properties(ops):
    """BASELINE sizes the oracle would take long on: size-independent properties instead.
    sum_c dz_ic == 0 per row (softmax-CE gradient), loss >= 0, loss_sum == sum(loss_i),
    and rank == 0 <=> argmax == label."""
    torch.manual_seed(0)
    B, C = 65536, 1000
    z = torch.randn(B, C, device=DEV) * 2
    y = torch.randint(0, C, (B,), device=DEV)
    r = ops.softmax_ce(z, None, y, scale=1.0 / B, want_acc=True)
    assert float(r["dz_f32"].sum(1).abs().max()) < 1e-9 + 1e-6 / B * 10
    assert float(r["loss_i"].min()) >= 0
    assert float(r["loss_sum"]) == pytest.approx(float(r["loss_i"].double().sum()), rel=1e-6)
    assert torch.equal(r["rank"] == 0, r["argmax"].long() == y)
    assert torch.equal(r["argmax"].long(), z.argmax(1))


# ------------------------------------------------------------------ Mixup: both labels in ONE pass
@pytest.mark.parametrize("variant", ["raw", "smooth"])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
@pytest.mark.parametrize("use_cw", [False, True])
@pytest.mark.parametrize("lam", [0.3, 1.0])
def test_mixup_fused_golden(ops, golden, variant, reduction, use_cw, lam):
    """iif_softmax_ce_mixup_fwd_bwd against custom.Mixup.mixup_criterion(custom.IIFLoss) of the unmodified
    reference (tests/golden/cls_mixup.npz)."""
    g = golden("cls_mixup")
    iif = ho.to_f32_row(ho.iif_weights_from_counts(g["counts"])[variant])
    B = g["z"].shape[0]
    scale = 1.0 / B if reduction == "mean" else 1.0
    r = ops.softmax_ce(T(g["z"]), T(iif), T(g["y_a"]), label_b=T(g["y_b"]), lam=lam,
                       class_weight=T(g["cw"]) if use_cw else None, scale=scale)
    tag = f"{variant}_{reduction}_{'cw' if use_cw else 'nocw'}_{lam}"
    assert float(r["loss_sum"]) == pytest.approx(float(g[f"loss_{tag}"]), rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), g[f"dz_{tag}"]) < TOL_F32


@pytest.mark.parametrize("B,C", [(256, 1000), (4096, 1204), (300, 8), (64, 10000)])
def test_mixup_fused_vs_oracle_and_module(ops, B, C):
    """Dual-label pass vs the oracle (ignored labels, sample weights), bf16 dZ, and the classification
    mirror: Mixup.mixup_criterion(IIFLoss) takes the fused path and back-propagates like two calls."""
    rng = np.random.default_rng(B + C)
    counts = lt_counts(C)
    z = (rng.standard_normal((B, C)) * 3).astype(np.float32)
    ya, yb = lt_labels(counts, B, rng), lt_labels(counts, B, rng)
    ya[::9] = -100
    yb[5::11] = -100
    sw = rng.uniform(0.5, 2.0, B).astype(np.float32)
    iif = iif_row(counts, "smooth")
    lam = 0.37
    li, dz = ho.mixup_ce(z, iif, ya, yb, lam, sample_weight=sw)
    r = ops.softmax_ce(T(z), T(iif), T(ya), label_b=T(yb), lam=lam, sample_weight=T(sw), scale=1.0 / B,
                       want_dz_bf16=True)
    assert rel_err(N(r["loss_i"]), li / B) < TOL_F32
    assert float(r["loss_sum"]) == pytest.approx(li.sum() / B, rel=TOL_F32)
    assert rel_err(N(r["dz_f32"]), dz / B) < TOL_F32
    assert rel_err(N(r["dz_bf16"])[:, :C], dz / B) < 8e-3
    # C % 4 != 0: the kernel declines, the mirror falls back to two passes
    with pytest.raises(ops.Unsupported):
        ops.softmax_ce(T(z[:, :C - 1].copy()), T(iif[:, :C - 1].copy()), T(np.clip(ya, -100, C - 2)),
                       label_b=T(np.clip(yb, -100, C - 2)), lam=lam)


def test_mixup_module_path():
    from iif_b200.classification import IIFLoss, Mixup

    class DS:
        def get_cls_num_list(self):
            return lt_counts(100).tolist()
    rng = np.random.default_rng(1)
    z = (rng.standard_normal((64, 100)) * 2).astype(np.float32)
    ya, yb = lt_labels(lt_counts(100), 64, rng), lt_labels(lt_counts(100), 64, rng)
    crit = IIFLoss(DS(), variant="smooth", device=DEV)
    mix = Mixup(crit, alpha=0.2)
    z1 = T(z).requires_grad_(True)
    l1 = mix.mixup_criterion(z1, T(ya), T(yb), 0.25)              # fused
    l1.backward()
    z2 = T(z).requires_grad_(True)
    l2 = 0.25 * crit(z2, T(ya)) + 0.75 * crit(z2, T(yb))          # the reference's two calls
    l2.backward()
    assert float(l1) == pytest.approx(float(l2), rel=1e-6)
    assert rel_err(N(z1.grad), N(z2.grad)) < 1e-6
    z3 = T(z[:, :99].copy()).requires_grad_(True)                  # 99 classes: falls back, still correct
    class DS99:
        def get_cls_num_list(self):
            return lt_counts(100).tolist()[:99]
    mix99 = Mixup(IIFLoss(DS99(), variant="raw", device=DEV))
    l3 = mix99.mixup_criterion(z3, T(np.clip(ya, 0, 98)), T(np.clip(yb, 0, 98)), 0.5)
    l3.backward()
    assert np.isfinite(float(l3)) and z3.grad is not None


# ------------------------------------------------------------------ focal loss (gamma > 0)
@pytest.mark.parametrize("gamma,alpha", [(2.0, None), (2.0, 0.25), (0.5, 0.25)])
@pytest.mark.parametrize("B,C", [(48, 40), (300, 1000), (64, 37)])
def test_focal_vs_oracle(ops, gamma, alpha, B, C):
    """iif_sigmoid_focal_fwd_bwd vs the float64 oracle (large logits included: the logit-space form stays
    accurate where sigmoid -> log in fp32 does not), per-class weights, bf16 dZ, both load paths."""
    rng = np.random.default_rng(B + C)
    z = (rng.standard_normal((B, C)) * 4).astype(np.float32)
    y = rng.integers(0, C, B).astype(np.int64)
    w = rng.uniform(0.5, 1.5, C).astype(np.float32)
    loss, dz = ho.focal_cls(z, y, gamma, alpha, w)
    scale = 1.0 / (B * C)
    r = ops.sigmoid_bce(T(z), T(y), col_weight=T(w), scale=scale, gamma=gamma, alpha=alpha, want_elem=True,
                        want_dz_bf16=True)
    assert float(r["loss_sum"]) == pytest.approx(loss.sum() * scale, rel=TOL_F32)
    assert rel_err(N(r["loss_elem"]), loss * scale) < TOL_F32
    assert rel_err(N(r["dz_f32"]), dz * scale) < TOL_F32
    assert rel_err(N(r["dz_bf16"])[:, :C], dz * scale) < 8e-3


@pytest.mark.parametrize("red", ["mean", "sum"])
def test_focal_golden_and_module(ops, golden, red):
    """Against custom.FocalLoss(gamma=2, alpha=0.25) of the unmodified reference, through the classification mirror."""
    from iif_b200.classification import FocalLoss
    g = golden("cls_focal")
    crit = FocalLoss(gamma=2.0, alpha=0.25, reduction=red, device=DEV, weights=T(g["weights"]))
    z = T(g["z"]).requires_grad_(True)
    loss = crit(z, T(g["y"]))
    loss.backward()
    key = f"2.0_0.25_w_{red}"
    assert float(loss.detach()) == pytest.approx(float(g[f"loss_{key}"]), rel=3e-5)
    assert rel_err(N(z.grad), g[f"dz_{key}"]) < 3e-5
