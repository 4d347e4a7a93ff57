"""GPU parity of the rows widened in round 2 against golden vectors frozen from the UNMODIFIED reference
(tests/golden/widen.npz, generator tests/golden/make_golden.py::gen_widen) and against the float64 oracle:
cls NormedLinear / CosNorm_Classifier(lr_scale=True), shot_acc, dense-label sigmoid BCE, FASA cums and
feature statistics."""
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import TOL_F32, TOL_BF16, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def T(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().double().cpu().numpy()


@pytest.mark.parametrize("compute,tol", [("fp32", 2e-5), ("bf16", TOL_BF16)])
def test_cls_normed_linear(golden, compute, tol):
    from iif_b200.classification import NormedLinear
    g = golden("widen")
    m = NormedLinear(64, 36, compute=compute, device=DEV)
    assert tuple(m.weight.shape) == (64, 36) and tuple(m.bias.shape) == (36,)      # [in, out] like the reference
    with torch.no_grad():
        m.weight.copy_(T(g["cls_normed_w"], torch.float32))
    x = T(g["cls_normed_x"], torch.float32).requires_grad_(True)
    z = m(x)
    z.backward(T(g["cls_normed_gz"], torch.float32))
    assert rel_err(N(z), g["cls_normed_z"]) < tol
    assert rel_err(N(x.grad), g["cls_normed_dx"]) < tol
    assert rel_err(N(m.weight.grad), g["cls_normed_dw"]) < tol
    assert m.bias.grad is None                                                     # unused in forward, as in the reference


@pytest.mark.parametrize("compute,tol", [("fp32", 2e-5), ("bf16", TOL_BF16)])
def test_cosnorm_lr_scale(golden, compute, tol):
    from iif_b200.classification import CosNorm_Classifier
    g = golden("widen")
    m = CosNorm_Classifier(64, 36, lr_scale=True, compute=compute, device=DEV)
    assert isinstance(m.scale, torch.nn.Parameter) and float(m.scale) == 5.0
    with torch.no_grad():
        m.weight.copy_(T(g["cosnorm_lr_w"], torch.float32))
    x = T(g["cosnorm_lr_x"], torch.float32).requires_grad_(True)
    z = m(x)
    z.backward(T(g["cosnorm_lr_gz"], torch.float32))
    assert rel_err(N(z), g["cosnorm_lr_z"]) < tol
    assert rel_err(N(x.grad), g["cosnorm_lr_dx"]) < tol
    assert rel_err(N(m.weight.grad), g["cosnorm_lr_dw"]) < tol
    assert rel_err(N(m.scale.grad), g["cosnorm_lr_dscale"]) < tol


def test_shot_acc(golden):
    from iif_b200.classification import shot_acc
    from iif_b200 import ops
    g = golden("widen")
    train = np.repeat(np.arange(len(g["shot_counts"])), g["shot_counts"])
    preds, labels = T(g["shot_preds"]), T(g["shot_labels"])
    many, med, low, cacc = shot_acc(preds, labels, train, acc_per_cls=True)
    assert np.allclose([many, med, low], g["shot_out"], rtol=1e-12, atol=0)
    assert np.allclose(cacc, g["shot_class_acc"], rtol=1e-12, atol=0) and len(cacc) == len(g["shot_class_acc"])
    assert np.allclose(shot_acc(preds.int(), labels, T(train), 1000, 2), g["shot_out_thr"], rtol=1e-12, atol=0)
    # integer part bit-exact
    out3, test, correct = ops.shot_accuracy(preds, labels, T(g["shot_counts"]))
    assert np.array_equal(test.cpu().numpy(), np.bincount(g["shot_labels"], minlength=len(g["shot_counts"])))
    hit = g["shot_labels"][g["shot_preds"] == g["shot_labels"]]
    assert np.array_equal(correct.cpu().numpy(), np.bincount(hit, minlength=len(g["shot_counts"])))
    # a big random case against the oracle
    rng = np.random.default_rng(0)
    C, n = 1000, 50000
    tc = rng.integers(1, 400, C)
    lab = rng.integers(0, C, n)
    prd = np.where(rng.random(n) < 0.5, lab, rng.integers(0, C, n))
    out3, _, _ = ops.shot_accuracy(T(prd), T(lab), T(tc))
    assert np.allclose(out3.cpu().numpy(), ho.shot_accuracy(prd, lab, tc), rtol=1e-12, atol=0)


def test_bce_dense_labels(golden):
    from iif_b200 import mmdet as M
    g = golden("widen")
    B, C = g["bced_z"].shape

    def run(tag, label, **kw):
        z = T(g["bced_z"]).requires_grad_(True)
        loss = M.binary_cross_entropy(z, T(label), **kw)
        loss.sum().backward()
        assert rel_err(N(loss), g[f"bced_loss_{tag}"]) < TOL_F32, tag
        assert rel_err(N(z.grad), g[f"bced_dz_{tag}"]) < TOL_F32, tag

    run("mean", g["bced_t"])
    run("soft_sum", g["bced_soft"], reduction="sum")
    run("wel_avg", g["bced_t"], weight=T(g["bced_wel"]), avg_factor=5.0)
    run("wrow_none", g["bced_t"], weight=T(np.repeat(g["bced_wrow"], C, 1)), reduction="none")
    run("pw_mean", g["bced_t"], class_weight=T(g["bced_pw"]))
    # the module path (CrossEntropyLoss(use_sigmoid=True) with expanded labels) and the [B,1] weight form
    crit = M.CrossEntropyLoss(use_sigmoid=True)
    z = T(g["bced_z"]).requires_grad_(True)
    l = crit(z, T(g["bced_t"]), weight=T(g["bced_wrow"]), reduction_override="none")
    assert rel_err(N(l), g["bced_loss_wrow_none"]) < TOL_F32


@pytest.fixture(scope="module")
def csv1204(golden, tmp_path_factory):
    """Rebuild idf_1204.csv from the frozen columns (the reference tree is absent on the GPU box)."""
    import pandas as pd
    g = golden("weight_tables")
    cols = {k[len("idf_1204_"):]: g[k] for k in g if k.startswith("idf_1204_") and g[k].shape == (1204,)}
    p = tmp_path_factory.mktemp("csv") / "idf_1204.csv"
    pd.DataFrame(cols).to_csv(p, index=False, float_format="%.17g")
    return str(p)


def test_fasa_cums_kernel(golden, csv1204):
    from iif_b200 import mmdet as M
    g = golden("widen")
    csv = csv1204
    crit = M.FasaIIFLoss(num_classes=1203, path=csv, variant="raw", use_cums=True, use_sigmoid=True)
    r = crit(T(g["cum_z"]), T(g["cum_y"]))
    assert rel_err(N(crit.cum_losses), g["cum_sig_losses"]) < TOL_F32
    assert np.array_equal(N(crit.cum_labels), g["cum_sig_labels"])
    assert float(r) == pytest.approx(float(g["cum_sig_ret"]), rel=TOL_F32)
    crit = M.FasaIIFLoss(num_classes=1203, path=csv, variant="raw", use_cums=True)
    r = crit(T(g["cum_z"]), T(g["cum_neg_y"]))
    assert rel_err(N(crit.cum_losses), g["cum_neg_losses"]) < TOL_F32
    assert np.array_equal(N(crit.cum_labels), g["cum_neg_labels"])       # label -100 binned at 1204 - 100, like the reference
    assert float(r) == pytest.approx(float(g["cum_neg_ret"]), rel=TOL_F32)


def test_class_accumulate_vs_oracle():
    from iif_b200 import ops
    rng = np.random.default_rng(3)
    for B, nb, cols in ((0, 5, 1), (1, 1, 1), (5000, 1204, 1), (700, 81, 9), (9000, 37, 1)):
        y = rng.integers(-3, nb + 2, B)                                       # a few labels outside [-nb, nb): skipped
        loss = rng.random((B, cols)).astype(np.float32)
        cl0, cn0 = rng.random(nb).astype(np.float32), rng.integers(0, 9, nb).astype(np.float32)
        cl, cn = T(cl0.copy()), T(cn0.copy())
        ops.class_accumulate(T(y), T(loss[:, 0] if cols == 1 else loss), cl, cn)
        ok = (y >= -nb) & (y < nb)
        rl, rn = ho.class_accumulate(y[ok], loss[ok], nb, cl0, cn0)
        assert np.array_equal(N(cn), rn) and rel_err(N(cl), rl) < TOL_F32


def test_class_feature_stats(golden):
    from iif_b200 import ops
    g = golden("widen")
    nb, D = g["fa_mean1"].shape
    mean = torch.zeros(nb, D, device=DEV); var = torch.zeros(nb, D, device=DEV); used = torch.zeros(nb, device=DEV)
    ops.class_feature_stats(T(g["fa_emb1"]), T(g["fa_lab1"]), mean, var, used, float(g["fa_decay"]))
    assert rel_err(N(mean), g["fa_mean1"]) < TOL_F32 and rel_err(N(var), g["fa_std1"]) < 2e-5
    assert np.array_equal(N(used), g["fa_used1"])
    ops.class_feature_stats(T(g["fa_emb2"]), T(g["fa_lab2"]), mean, var, used, float(g["fa_decay"]))
    assert rel_err(N(mean), g["fa_mean2"]) < TOL_F32 and rel_err(N(var), g["fa_std2"]) < 2e-5
    assert np.array_equal(N(used), g["fa_used2"])
    # LVIS-sized: 1204 classes x 1024 features, 2048 RoIs, against the oracle
    rng = np.random.default_rng(1)
    nb, D, B = 1204, 1024, 2048
    x = rng.standard_normal((B, D)).astype(np.float32)
    y = np.where(rng.random(B) < 0.75, 1203, rng.integers(0, 1203, B))
    m0, v0 = rng.random((nb, D)).astype(np.float32), rng.random((nb, D)).astype(np.float32)
    u0 = (rng.random(nb) < 0.5).astype(np.float32)
    mean, var, used = T(m0.copy()), T(v0.copy()), T(u0.copy())
    ops.class_feature_stats(T(x), T(y), mean, var, used, 0.1)
    rm, rv, ru = ho.class_feature_stats(x, y, m0, v0, u0, 0.1)
    assert rel_err(N(mean), rm) < TOL_F32 and rel_err(N(var), rv) < 2e-5 and np.array_equal(N(used), ru)


@pytest.mark.parametrize("B,D,C1,C2", [(1024, 1024, 1204, 4812), (77, 520, 9, 36), (512, 256, 81, 320)])
def test_sibling_fc_cls_fc_reg(B, D, C1, C2):
    """fc_cls + fc_reg on the same RoI features (bbox_head.py:118-119): one GEMM per direction, same results as the two
    separate layers / the oracle on bf16-rounded operands."""
    from iif_b200 import mmdet as M
    from _common import bf16_round
    rng = np.random.default_rng(B + C1)
    x = np.maximum(rng.standard_normal((B, D)), 0).astype(np.float32)
    fc_cls, fc_reg = M.Linear(D, C1).to(DEV), M.Linear(D, C2).to(DEV)
    xt = T(x).requires_grad_(True)
    zc, zr = M.sibling_forward(fc_cls, fc_reg, xt)
    gc, gr = rng.standard_normal((B, C1)).astype(np.float32) / B, rng.standard_normal((B, C2)).astype(np.float32) / B
    (zc * T(gc)).sum().backward(retain_graph=True)
    (zr * T(gr)).sum().backward()
    wc, wr = N(fc_cls.weight), N(fc_reg.weight)
    xb = bf16_round(x)
    for z, w, b in ((zc, wc, N(fc_cls.bias)), (zr, wr, N(fc_reg.bias))):
        assert rel_err(N(z), ho.linear_fwd(xb, bf16_round(w.astype(np.float32)), b)) < 2e-5
    dxc, dwc, dbc = ho.linear_bwd(bf16_round(gc), xb, bf16_round(wc.astype(np.float32)))
    dxr, dwr, dbr = ho.linear_bwd(bf16_round(gr), xb, bf16_round(wr.astype(np.float32)))
    assert rel_err(N(xt.grad), dxc + dxr) < 1e-2
    assert rel_err(N(fc_cls.weight.grad), dwc) < 1e-2 and rel_err(N(fc_reg.weight.grad), dwr) < 1e-2
    assert rel_err(N(fc_cls.bias.grad), dbc) < 1e-2 and rel_err(N(fc_reg.bias.grad), dbr) < 1e-2


@pytest.mark.parametrize("B,D,C", [(256, 2048, 1000), (128, 64, 10), (77, 520, 1203), (1024, 1024, 1204)])
def test_fp32_parity_on_tensor_cores(B, D, C):
    """compute='fp32x3': fp32 operands split into three bf16 terms, six partial products on the tensor cores, fp32
    accumulation -- forward and gradients within the 1e-5 bar of the float64 oracle (csrc/split3.cu)."""
    from iif_b200 import mmdet as M
    from _common import head_inputs
    x, w, b, counts, y = head_inputs(B, D, C, seed=7)
    fc = M.Linear(D, C, compute="fp32x3").to(DEV)
    with torch.no_grad():
        fc.weight.copy_(T(w)); fc.bias.copy_(T(b))
    xt = T(x).requires_grad_(True)
    z = fc(xt)
    gz = np.random.default_rng(1).standard_normal((B, C)).astype(np.float32) / B
    z.backward(T(gz))
    assert rel_err(N(z), ho.linear_fwd(x, w, b)) < TOL_F32
    dx, dw, db = ho.linear_bwd(gz, x, w)
    assert rel_err(N(xt.grad), dx) < TOL_F32 and rel_err(N(fc.weight.grad), dw) < TOL_F32
    assert rel_err(N(fc.bias.grad), db) < TOL_F32
