"""Pin the float64 oracle against the reference's own outputs (tests/golden, made by
make_golden.py from the unmodified reference Python), its CSV weight tables and the
known-answer tests of its mmdet suite.  CPU only."""
import numpy as np
import pytest

from oracle import head_oracle as ho

RTOL = 2e-5  # reference outputs are fp32; the oracle is fp64

CSV_TO_VARIANT = {"smooth": "smooth", "raw": "raw", "prob": "rel", "normit": "normit",
                  "gombit": "gombit", "base2": "base2", "base10": "base10"}


def close(a, b, rtol=RTOL, atol=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    atol = rtol * max(np.abs(b).max(), 1e-30) if atol is None else atol
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ------------------------------------------------------------------ weights
def test_cifar_lt_profile():
    # closed form of cls/imbalanced_dataset.py:23-29 with 50k CIFAR images, r=100
    assert ho.cifar_lt_profile(5000, 10, 0.01) == [5000, 2997, 1796, 1077, 645, 387, 232, 139, 83, 50]


@pytest.mark.parametrize("v", ho.VARIANTS)
def test_cls_weights_bitexact(golden, v):
    g = golden("cls_iif")
    w = ho.iif_weights_from_counts(g["counts"])[v]
    assert np.array_equal(ho.to_f32_row(w), g[f"iif_{v}"])          # one f64->f32 rounding
    n2 = ho.iif_normalise_f32(ho.to_f32_row(w), 2)
    close(n2, g[f"iifn2_{v}"], rtol=1e-6)


@pytest.mark.parametrize("table", ["idf_1204", "idf_1231", "idf_91"])
def test_csv_tables_closed_form(golden, table):
    """Every variant column of the reference CSVs is the closed form of its frequency column."""
    g = golden("weight_tables")
    img, inst = g[f"{table}_img_freq"], g[f"{table}_instance_freq"]
    for total, freq, suf in ((int(g[f"{table}_n_img"]), img, ""), (int(inst.sum()), inst, "_obj")):
        got = ho.iif_weights_from_counts(freq, total)
        for col, var in CSV_TO_VARIANT.items():
            ref = g[f"{table}_{col}{suf}"][1:]
            np.testing.assert_allclose(got[var], ref, rtol=0, atol=5e-13 * max(1, np.abs(ref).max()))
    assert (g["idf_91_prob"][1:] < 0).any()      # negative weights exist (coco person)


def test_csv_column_to_weights(golden):
    g, t = golden("mmdet_iif"), golden("weight_tables")
    for col in ("raw", "smooth", "prob_obj", "base10_obj"):
        w = ho.csv_column_to_weights(t[f"idf_1204_{col}"])
        assert w.shape == (1, 1204) and w[0, -1] == 1.0
        assert np.array_equal(w, g[f"iif_{col}"])


# ------------------------------------------------------------------ histogram
def test_label_hist_and_map():
    rng = np.random.default_rng(0)
    y = rng.integers(-2, 12, size=5000)
    h = ho.label_hist(y, 10)
    assert h.tolist() == [int((y == i).sum()) for i in range(10)]
    cmap = ho.lt_class_map(h)
    h2 = ho.label_hist(cmap[y[(y >= 0) & (y < 10)]], 10)
    assert (np.diff(h2) <= 0).all()                                  # descending after remap
    img = rng.integers(0, 50, size=2000)
    cat = rng.integers(0, 7, size=2000)
    imf, inf_ = ho.image_dedup_hist(img, cat, 7)
    assert inf_.tolist() == [int((cat == c).sum()) for c in range(7)]
    assert imf.tolist() == [len(set(img[cat == c])) for c in range(7)]


# ------------------------------------------------------------------ classification IIFLoss
TAGS = [f"{v}_mean" for v in ho.VARIANTS] + ["raw_sum", "raw_none", "smooth_sum", "smooth_none"]


@pytest.mark.parametrize("tag", TAGS)
def test_cls_iif_head(golden, tag):
    g = golden("cls_iif")
    v, red = tag.rsplit("_", 1)
    iif = g[f"iif_{v}"]
    close(ho.linear_fwd(g["x"], g["w"], g["b"]), g["z"])
    scale = 1.0 / g["x"].shape[0] if red == "mean" else 1.0
    r = ho.head_fwd_bwd(g["x"], g["w"], g["b"], iif, g["y"], scale=scale)
    val, _ = ho.reduce_cls(r["loss_i"], red)
    close(val, g[f"loss_{tag}"])
    for k in ("dz", "dx", "dw", "db"):
        close(r[k], g[f"{k}_{tag}"])
    if red == "mean":
        close(ho.linear_fwd(g["x"], g["w"], g["b"]) * iif, g[f"infer_{v}"])


def test_cls_iif_class_weight_and_norm(golden):
    g = golden("cls_iif")
    r = ho.head_fwd_bwd(g["x"], g["w"], g["b"], g["iif_smooth"], g["y"], class_weight=g["cw"])
    close(r["loss_i"].mean(), g["loss_smooth_cw_mean"])                # plain mean even with weights
    close(r["dw"], g["dw_smooth_cw_mean"])
    r = ho.head_fwd_bwd(g["x"], g["w"], g["b"], g["iifn2_raw"], g["y"])
    close(r["loss"], g["loss_raw_n2_mean"])
    close(r["dz"], g["dz_raw_n2_mean"])


def test_cls_bce(golden):
    g = golden("cls_bce")
    B, C = g["z"].shape
    for tag, w in (("now", None), ("w", g["weights"])):
        loss, dz = ho.sigmoid_bce_cls(g["z"], g["y"], w)
        close(loss.mean(), g[f"loss_{tag}_mean"])
        close(dz / (B * C), g[f"dz_{tag}_mean"])
        close(loss.sum() / B, g[f"loss_{tag}_sum"])
        close(dz / B, g[f"dz_{tag}_sum"])


# ------------------------------------------------------------------ mmdet IIFLoss
def _mm(g, iif, tag, *, y=None, weight=True, avg=True, reduction="mean", cw=None, lw=1.0, ign=-100, rows=None):
    y = g["y"] if y is None else y
    li, dz, _ = ho.softmax_ce(g["z"], iif, y, cw, g["w"] if weight else None, ign)
    val, sc = ho.reduce_mmdet(li, reduction, float(g["avg_factor"]) if avg else None, lw)
    close(val, g[f"loss_{tag}"])
    close((dz * sc)[:rows], g[f"dz_{tag}"])


def test_mmdet_iif_all_columns(golden):
    g = golden("mmdet_iif")
    for col in list(CSV_TO_VARIANT) + [c + "_obj" for c in CSV_TO_VARIANT]:
        _mm(g, g[f"iif_{col}"], f"{col}_avg", rows=8)


def test_mmdet_iif_reductions(golden):
    g = golden("mmdet_iif")
    s = g["iif_raw"]
    _mm(g, s, "raw_plain", weight=False, avg=False)
    _mm(g, s, "raw_w_mean", avg=False)
    _mm(g, s, "raw_none", avg=False, reduction="none")
    _mm(g, s, "raw_sum", avg=False, reduction="sum")
    _mm(g, s, "raw_none_avg", reduction="none")
    _mm(g, s, "raw_ignbg", y=np.maximum(g["y"], 0), ign=1203)
    _mm(g, g["iif_smooth"], "smooth_cw_lw", cw=g["class_weight"], lw=0.5)
    _mm(g, g["iif_normit_obj"], "normit_obj_sum", weight=False, avg=False, reduction="sum")
    _mm(g, g["iif_base10_obj"], "fasa_base10_obj_avg")
    with pytest.raises(ValueError):
        ho.reduce_mmdet(np.ones(3), "sum", 2.0)                        # losses/utils.py:53-54


def test_mmdet_activation_accuracy(golden):
    g = golden("mmdet_iif")
    close(ho.softmax_activation(g["z"], g["iif_raw"]), g["act_raw"])
    close(ho.softmax_activation(g["z"], g["iif_base10_obj"]), g["fasa_act"])
    yc = np.maximum(g["y"], 0)
    a1, a5 = ho.topk_accuracy(g["z"], yc, (1, 5))
    assert np.float32(a1) == g["acc_top1"][0] and np.float32(a5) == g["acc_top5"][0]
    assert np.array_equal(ho.argmax_first(g["z"]), g["topk5_idx"][:, 0])
    r = ho.label_rank(g["z"], yc)
    for i in range(len(yc)):                                          # rank < 5  <=>  label in torch's top-5
        assert (r[i] < 5) == (yc[i] in g["topk5_idx"][i])
    # get_accuracy is on RAW scores with the raw labels (-100 rows can never match)
    a = ho.topk_accuracy(g["z"], g["y"], (1,))[0]
    assert np.float32(a) == g["acc_raw"][0]


def test_fasa_sigmoid_and_cums(golden):
    g = golden("mmdet_iif")
    l, dz = ho.sigmoid_bce_mmdet(g["z"], g["y"], g["w"])              # no IIF in sigmoid mode
    val, sc = ho.reduce_mmdet(l, "mean", float(g["avg_factor"]))
    close(val, g["loss_fasa_sigmoid_avg"])
    close(dz * sc, g["dz_fasa_sigmoid_avg"])
    yc = np.maximum(g["y"], 0)
    cl = np.zeros(1204)
    cn = np.zeros(1204)
    rets = []
    for zz in (g["z"], g["z"] * 0.5):                                  # fasa_iif_loss.py:154-160
        li, _, _ = ho.softmax_ce(zz, g["iif_raw"], yc)
        np.add.at(cl, yc, li)
        np.add.at(cn, yc, 1)
        rets.append(li.mean())
    close(cl, g["fasa_cum_losses"])
    assert np.array_equal(cn, g["fasa_cum_labels"])
    close(rets, g["fasa_cum_ret"])


def test_mmdet_bce(golden):
    g = golden("mmdet_bce")
    avg = float(g["avg_factor"])

    def chk(tag, *, weight=True, avgf=None, red="mean", pw=None, lw=1.0, ign=-100):
        l, dz = ho.sigmoid_bce_mmdet(g["z"], g["y"], g["w"] if weight else None, pw, ign)
        val, sc = ho.reduce_mmdet(l, red, avgf, lw)
        close(val, g[f"loss_{tag}"])
        close(dz * sc, g[f"dz_{tag}"])

    chk("plain", weight=False)
    chk("avg", avgf=avg)
    chk("ign255_avg", avgf=avg, ign=255)
    chk("pw_lw_avg", avgf=avg, pw=g["pos_weight"], lw=2.0)
    chk("none", red="none")
    chk("sum", red="sum")
    li, dz, _ = ho.softmax_ce(g["ce_z"], None, g["ce_y"], None, g["w"])
    val, sc = ho.reduce_mmdet(li, "mean", avg)
    close(val, g["ce_loss"])
    close(dz * sc, g["ce_dz"])


# ------------------------------------------------------------------ reference KATs
def test_kat_ce_loss():
    """seg/tests/test_metrics/test_losses.py:8-32: pred [[100,-100]], label 1 -> 200; cw [0.8,0.2] -> 40."""
    z = np.array([[100.0, -100.0]])
    li, _, _ = ho.softmax_ce(z, None, [1])
    assert abs(ho.reduce_mmdet(li)[0] - 200.0) < 1e-9
    li, _, _ = ho.softmax_ce(z, None, [1], class_weight=[0.8, 0.2])
    assert abs(ho.reduce_mmdet(li)[0] - 40.0) < 1e-9


def test_kat_accuracy():
    """seg/tests/test_metrics/test_losses.py:186-240."""
    pred = np.array([[0.2, 0.3, 0.6, 0.5], [0.1, 0.1, 0.2, 0.6], [0.9, 0.0, 0.0, 0.1],
                     [0.4, 0.7, 0.1, 0.1], [0.0, 0.0, 0.99, 0]], np.float32)
    assert ho.topk_accuracy(pred, [2, 3, 0, 1, 2], (1,))[0] == 100
    assert ho.topk_accuracy(pred, [2, 3, 0, 1, 2], (1,), thresh=0.8)[0] == 40
    assert ho.topk_accuracy(pred, [3, 2, 0, 0, 2], (2,))[0] == 100
    assert ho.topk_accuracy(pred, [2, 3, 0, 1, 2], (1, 2)) == [100.0, 100.0]
    assert ho.topk_accuracy(np.zeros((0, 4), np.float32), [], (1,))[0] == 0


def test_ignore_index_equivalence():
    """seg/tests/test_models/test_loss.py:137-165: ignoring a row == dropping it (sum)."""
    rng = np.random.default_rng(1)
    z = rng.standard_normal((10, 5))
    y = rng.integers(0, 5, 10)
    y2 = y.copy()
    y2[[2, 7]] = 255
    keep = np.ones(10, bool)
    keep[[2, 7]] = False
    a, _, _ = ho.softmax_ce(z, None, y2, ignore_index=255)
    b, _, _ = ho.softmax_ce(z[keep], None, y[keep])
    assert abs(a.sum() - b.sum()) < 1e-12
    a, _ = ho.sigmoid_bce_mmdet(z, y2, ignore_index=255)
    b, _ = ho.sigmoid_bce_mmdet(z[keep], y[keep])
    assert abs(a.sum() - b.sum()) < 1e-12


def test_shot_accuracy():
    preds = np.array([0, 0, 1, 2, 2, 1])
    labels = np.array([0, 0, 1, 2, 2, 2])
    assert ho.shot_accuracy(preds, labels, [500, 50, 5]) == (1.0, 1.0, pytest.approx(2 / 3))


# ------------------------------------------------------------------ torch CPU port (the bench's CPU arm)
@pytest.mark.parametrize("tag", ["raw_mean", "smooth_mean", "normit_mean", "raw_sum", "smooth_none"])
def test_torch_port_cls(golden, tag):
    import torch
    from oracle import torch_port as tp
    g = golden("cls_iif")
    v, red = tag.rsplit("_", 1)
    t = lambda k: torch.from_numpy(g[k])
    r = tp.head_step(t("x"), t("w"), t("b"), t(f"iif_{v}"), t("y"), reduction=red)
    close(r["loss"].numpy(), g[f"loss_{tag}"])
    for k in ("dx", "dw", "db"):
        close(r[k].numpy(), g[f"{k}_{tag}"])


def test_torch_port_mmdet(golden):
    import torch
    from oracle import torch_port as tp
    g = golden("mmdet_iif")
    z = torch.from_numpy(g["z"])
    eye = torch.eye(z.shape[1])
    r = tp.head_step(z, eye, None, torch.from_numpy(g["iif_raw"]), torch.from_numpy(g["y"]),
                     sample_weight=torch.from_numpy(g["w"]), avg_factor=float(g["avg_factor"]))
    close(r["loss"].numpy(), g["loss_raw_avg"])
    close(r["dx"].numpy()[:8], g["dz_raw_avg"])                       # W = I: dX is dZ


# ------------------------------------------------------------------ Mixup (dual-label) loss
@pytest.mark.parametrize("variant", ["raw", "smooth"])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
@pytest.mark.parametrize("use_cw", [False, True])
@pytest.mark.parametrize("lam", [0.3, 1.0])
def test_mixup_golden(golden, variant, reduction, use_cw, lam):
    """custom.Mixup.mixup_criterion around custom.IIFLoss (unmodified reference) vs the oracle's mixup_ce."""
    g = golden("cls_mixup")
    iif = ho.to_f32_row(ho.iif_weights_from_counts(g["counts"])[variant])
    cw = g["cw"] if use_cw else None
    loss_i, dz = ho.mixup_ce(g["z"], iif, g["y_a"], g["y_b"], lam, class_weight=cw)
    val, scale = ho.reduce_cls(loss_i, reduction)
    tag = f"{variant}_{reduction}_{'cw' if use_cw else 'nocw'}_{lam}"
    close(val, g[f"loss_{tag}"])
    close(dz * scale, g[f"dz_{tag}"])


# ------------------------------------------------------------------ focal loss (gamma > 0)
@pytest.mark.parametrize("gamma", [2.0, 0.5])
@pytest.mark.parametrize("alpha", [None, 0.25])
@pytest.mark.parametrize("red", ["mean", "sum"])
@pytest.mark.parametrize("wtag", ["now", "w"])
def test_focal_golden(golden, gamma, alpha, red, wtag):
    """custom.FocalLoss(gamma > 0) of the unmodified reference (fp32, sigmoid -> BCELoss -> pow) vs ho.focal_cls."""
    g = golden("cls_focal")
    B, C = g["z"].shape
    loss, dz = ho.focal_cls(g["z"], g["y"], gamma, alpha, g["weights"] if wtag == "w" else None)
    scale = 1.0 / B if red == "sum" else 1.0 / (B * C)
    key = f"{gamma}_{alpha}_{wtag}_{red}"
    close(loss.sum() * scale, g[f"loss_{key}"], rtol=3e-5)
    close(dz * scale, g[f"dz_{key}"], rtol=3e-5)


# ------------------------------------------------------------------ normalised classifiers
@pytest.mark.parametrize("tag,kw", [("normed_p1", dict(temperature=20.0, power=1.0, eps=1e-6)),
                                    ("normed_p2", dict(temperature=10.0, power=2.0, eps=1e-3)),
                                    ("normed_nobias", dict(temperature=20.0, power=1.0, eps=1e-6)),
                                    ("iifnormed", dict(temperature=20.0, power=1.0, eps=1e-6))])
def test_normed_linear_golden(golden, tag, kw):
    """mmdet NormedLinear / IIFNormedLinear (unmodified reference, float64) vs ho.normed_linear."""
    g = golden("normed")
    iif = g["iifnormed_iif"] if tag == "iifnormed" else None
    b = g[f"{tag}_b"] if f"{tag}_b" in g else None
    z, dx, dw, db = ho.normed_linear(g[f"{tag}_x"], g[f"{tag}_w"], b, g[f"{tag}_gz"], iif=iif, **kw)
    close(z, g[f"{tag}_z"], rtol=1e-9)
    close(dx, g[f"{tag}_dx"], rtol=1e-9)
    close(dw, g[f"{tag}_dw"], rtol=1e-9)
    if b is not None:
        close(db, g[f"{tag}_db"], rtol=1e-9)


def test_cosnorm_golden(golden):
    """cls CosNorm_Classifier (unmodified reference, float64) vs ho.cosnorm_classifier."""
    g = golden("normed")
    z, dx, dw = ho.cosnorm_classifier(g["cosnorm_x"], g["cosnorm_w"], g["cosnorm_gz"], scale=16.0)
    close(z, g["cosnorm_z"], rtol=1e-9)
    close(dx, g["cosnorm_dx"], rtol=1e-9)
    close(dw, g["cosnorm_dw"], rtol=1e-9)


# ------------------------------------------------------------------ size-independent properties of the oracle
def test_oracle_properties_widened_rows():
    """Properties the widened rows must satisfy whatever the size (they also guard the analytic backward
    formulas of the oracle against a finite-difference check)."""
    rng = np.random.default_rng(11)
    B, C, D = 12, 20, 16
    z = rng.standard_normal((B, C)) * 2
    iif = rng.uniform(0.5, 6.0, (1, C))
    ya, yb = rng.integers(0, C, B), rng.integers(0, C, B)
    # Mixup: lam = 1 is the single-label loss, the loss is linear in lam, its gradient is the lam-mix
    l1, d1 = ho.mixup_ce(z, iif, ya, yb, 1.0)
    la, da, _ = ho.softmax_ce(z, iif, ya)
    lb, db, _ = ho.softmax_ce(z, iif, yb)
    np.testing.assert_allclose(l1, la, rtol=1e-12)
    lh, dh = ho.mixup_ce(z, iif, ya, yb, 0.25)
    np.testing.assert_allclose(lh, 0.25 * la + 0.75 * lb, rtol=1e-12)
    np.testing.assert_allclose(dh, 0.25 * da + 0.75 * db, rtol=1e-12, atol=1e-15)
    # focal: gamma -> 0 (no alpha) tends to plain BCE; finite differences of the loss match dz
    lf, df = ho.focal_cls(z, ya, 1e-9)
    lbce, dbce = ho.sigmoid_bce_cls(z, ya)
    np.testing.assert_allclose(lf, lbce, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(df, dbce, rtol=1e-6, atol=1e-9)
    eps = 1e-6
    lf2, df2 = ho.focal_cls(z, ya, 2.0, 0.25)
    zp = z.copy(); zp[3, 5] += eps
    zm = z.copy(); zm[3, 5] -= eps
    fd = (ho.focal_cls(zp, ya, 2.0, 0.25)[0].sum() - ho.focal_cls(zm, ya, 2.0, 0.25)[0].sum()) / (2 * eps)
    assert abs(fd - df2[3, 5]) < 1e-6 * max(1.0, abs(fd))
    # normalised classifiers: with eps = 0 and p = 1 the logits do not change when a feature row is rescaled;
    # finite differences of <gz, z> match dx and dw
    x = rng.standard_normal((B, D)); w = rng.standard_normal((C, D)) * 0.1; b = rng.standard_normal(C) * 0.1
    gz = rng.standard_normal((B, C))
    z0, dx, dw, dbb = ho.normed_linear(x, w, b, gz, temperature=20.0, power=1.0, eps=0.0)
    z1 = ho.normed_linear(x * rng.uniform(0.5, 3.0, (B, 1)), w, b, gz, temperature=20.0, power=1.0, eps=0.0)[0]
    np.testing.assert_allclose(z1, z0, rtol=1e-10, atol=1e-12)
    f = lambda xx, ww: (ho.normed_linear(xx, ww, b, gz, 20.0, 1.5, 1e-3, iif=iif.reshape(-1))[0] * gz).sum()
    _, dx2, dw2, _ = ho.normed_linear(x, w, b, gz, 20.0, 1.5, 1e-3, iif=iif.reshape(-1))
    xp = x.copy(); xp[2, 7] += eps; xm = x.copy(); xm[2, 7] -= eps
    assert abs((f(xp, w) - f(xm, w)) / (2 * eps) - dx2[2, 7]) < 1e-5 * max(1.0, abs(dx2[2, 7]))
    wp = w.copy(); wp[4, 1] += eps; wm = w.copy(); wm[4, 1] -= eps
    assert abs((f(x, wp) - f(x, wm)) / (2 * eps) - dw2[4, 1]) < 1e-5 * max(1.0, abs(dw2[4, 1]))
    zc, dxc, dwc = ho.cosnorm_classifier(x, w, gz)
    assert np.abs(zc).max() <= 16.0 + 1e-9          # |scale * cos| <= scale


# ------------------------------------------------------------------ widened rows of round 2 (tests/golden/widen.npz)
def test_widen_cls_normed_and_cosnorm_lr(golden):
    g = golden("widen")
    z, dx, dw = ho.cls_normed_linear(g["cls_normed_x"], g["cls_normed_w"], g["cls_normed_gz"])
    close(z, g["cls_normed_z"]); close(dx, g["cls_normed_dx"]); close(dw, g["cls_normed_dw"])
    z, dx, dw, ds = ho.cosnorm_classifier_lr(g["cosnorm_lr_x"], g["cosnorm_lr_w"], g["cosnorm_lr_gz"], g["cosnorm_lr_scale"])
    close(z, g["cosnorm_lr_z"]); close(dx, g["cosnorm_lr_dx"]); close(dw, g["cosnorm_lr_dw"])
    close([ds], g["cosnorm_lr_dscale"])


def test_widen_shot_acc(golden):
    g = golden("widen")
    close(ho.shot_accuracy(g["shot_preds"], g["shot_labels"], g["shot_counts"]), g["shot_out"])
    close(ho.shot_accuracy(g["shot_preds"], g["shot_labels"], g["shot_counts"], 1000, 2), g["shot_out_thr"])


def test_widen_bce_dense(golden):
    g = golden("widen")
    z, B, C = g["bced_z"], *g["bced_z"].shape

    def chk(tag, t, red, avgf=None, w=None, pw=None):
        l, dz = ho.bce_dense(z, t, w, pw)
        val, sc = ho.reduce_mmdet(l, red, avgf)
        close(val, g[f"bced_loss_{tag}"]); close(dz * sc, g[f"bced_dz_{tag}"])

    chk("mean", g["bced_t"], "mean")
    chk("soft_sum", g["bced_soft"], "sum")
    chk("wel_avg", g["bced_t"], "mean", 5.0, g["bced_wel"])
    chk("wrow_none", g["bced_t"], "none", None, np.repeat(g["bced_wrow"], C, 1))
    chk("pw_mean", g["bced_t"], "mean", None, None, g["bced_pw"])


def test_widen_fasa_cums_and_stats(golden):
    g = golden("widen")
    l, _ = ho.sigmoid_bce_mmdet(g["cum_z"], g["cum_y"])                 # sigmoid mode: [B,C] loss, rows summed per class
    cl, cn = ho.class_accumulate(g["cum_y"], l, 1204)
    close(cl, g["cum_sig_losses"]); assert np.array_equal(cn, g["cum_sig_labels"]); close([l.mean()], [g["cum_sig_ret"]])
    iif = ho.csv_column_to_weights(golden("weight_tables")["idf_1204_raw"])
    li, _, _ = ho.softmax_ce(g["cum_z"], iif, g["cum_neg_y"])
    cl, cn = ho.class_accumulate(g["cum_neg_y"], li, 1204)              # label -100 lands in bin 1204 - 100
    close(cl, g["cum_neg_losses"]); assert np.array_equal(cn, g["cum_neg_labels"]) and cn[1104] == 1
    close([li.mean()], [g["cum_neg_ret"]])
    nb, D = g["fa_mean1"].shape
    m, v, u = ho.class_feature_stats(g["fa_emb1"], g["fa_lab1"], np.zeros((nb, D)), np.zeros((nb, D)), np.zeros(nb), 0.1)
    close(m, g["fa_mean1"]); close(v, g["fa_std1"]); assert np.array_equal(u, g["fa_used1"])
    m, v, u = ho.class_feature_stats(g["fa_emb2"], g["fa_lab2"], m, v, u, float(g["fa_decay"]))
    close(m, g["fa_mean2"]); close(v, g["fa_std2"]); assert np.array_equal(u, g["fa_used2"])
