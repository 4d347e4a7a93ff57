"""GPU parity of the drop-in modules (the reference's own API surface) against the golden vectors
produced by the unmodified reference Python and against the float64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import TOL_F32, TOL_BF16, rel_err, head_inputs, iif_row, bf16_round, lt_labels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def T(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().float().cpu().numpy()


class _DS:
    def __init__(self, counts):
        self.counts = [int(c) for c in counts]

    def get_cls_num_list(self):
        return self.counts


# ------------------------------------------------------------------ classification/custom.py surface
@pytest.mark.parametrize("tag", [f"{v}_mean" for v in ho.VARIANTS] + ["raw_sum", "raw_none", "smooth_sum"])
def test_cls_iifloss_module(golden, tag):
    from iif_b200.classification import IIFLoss
    from iif_b200 import functional as F_
    g = golden("cls_iif")
    v, red = tag.rsplit("_", 1)
    crit = IIFLoss(_DS(g["counts"]), variant=v, reduction=red, device=DEV)
    assert hasattr(crit, "iif") and set(crit.iif) == set(ho.VARIANTS) and crit.iif[v].shape == (1, 10)
    x = T(g["x"]).requires_grad_(True)
    w = T(g["w"]).requires_grad_(True)
    b = T(g["b"]).requires_grad_(True)
    out = F_.linear(x, w, b)                      # fp32 parity mode of fc
    assert rel_err(N(out), g["z"]) < TOL_F32
    loss = crit(out, T(g["y"]))
    if red == "none":
        assert rel_err(N(loss), g[f"loss_{tag}"]) < TOL_F32
        loss.sum().backward()
    else:
        assert float(loss) == pytest.approx(float(g[f"loss_{tag}"]), rel=TOL_F32)
        loss.backward()
    assert rel_err(N(x.grad), g[f"dx_{tag}"]) < TOL_F32
    assert rel_err(N(w.grad), g[f"dw_{tag}"]) < TOL_F32
    assert rel_err(N(b.grad), g[f"db_{tag}"]) < TOL_F32
    if red == "mean":
        assert np.array_equal(N(crit(out.detach(), infer=True)), g[f"infer_{v}"])


def test_cls_iifloss_class_weight_norm_and_mixup(golden):
    from iif_b200.classification import IIFLoss, Mixup
    g = golden("cls_iif")
    z = T(g["z"]).requires_grad_(True)
    crit = IIFLoss(_DS(g["counts"]), variant="smooth", device=DEV, weight=T(g["cw"]))
    loss = crit(z, T(g["y"]))
    assert float(loss) == pytest.approx(float(g["loss_smooth_cw_mean"]), rel=TOL_F32)
    crit = IIFLoss(_DS(g["counts"]), variant="raw", iif_norm=2, device=DEV)
    assert rel_err(N(crit.iif["raw"]), g["iifn2_raw"]) < 1e-6
    z2 = T(g["z"]).requires_grad_(True)
    loss = crit(z2, T(g["y"]))
    assert float(loss) == pytest.approx(float(g["loss_raw_n2_mean"]), rel=TOL_F32)
    loss.backward()
    assert rel_err(N(z2.grad), g["dz_raw_n2_mean"]) < TOL_F32
    # Mixup.mixup_criterion: lam * crit(p, ya) + (1 - lam) * crit(p, yb)  (custom.py:116-117)
    mx = Mixup(crit)
    yb = T(np.roll(g["y"], 3))
    z3 = T(g["z"]).requires_grad_(True)
    lm = mx.mixup_criterion(z3, T(g["y"]), yb, 0.3)
    la, dza, _ = ho.softmax_ce(g["z"], g["iifn2_raw"], g["y"])
    lb, dzb, _ = ho.softmax_ce(g["z"], g["iifn2_raw"], np.roll(g["y"], 3))
    assert float(lm) == pytest.approx(0.3 * la.mean() + 0.7 * lb.mean(), rel=TOL_F32)
    lm.backward()
    assert rel_err(N(z3.grad), (0.3 * dza + 0.7 * dzb) / len(la)) < TOL_F32


def test_cls_focal_gamma0_and_accuracy(golden):
    from iif_b200.classification import FocalLoss, accuracy, predictions
    g = golden("cls_bce")
    for tag, w in (("now", None), ("w", g["weights"])):
        for red in ("mean", "sum"):
            z = T(g["z"]).requires_grad_(True)
            crit = FocalLoss(0, reduction=red, weights=None if w is None else T(w))
            loss = crit(z, T(g["y"]))
            assert float(loss) == pytest.approx(float(g[f"loss_{tag}_{red}"]), rel=TOL_F32)
            loss.backward()
            assert rel_err(N(z.grad), g[f"dz_{tag}_{red}"]) < TOL_F32
    with pytest.raises(ValueError):
        FocalLoss(-1.0)
    c = golden("cls_iif")
    a1, a5 = accuracy(T(c["z"]), T(c["y"]), topk=(1, 5))
    e1, e5 = ho.topk_accuracy(c["z"], c["y"], (1, 5))
    assert float(a1) == pytest.approx(e1, abs=1e-4) and float(a5) == pytest.approx(e5, abs=1e-4)
    assert np.array_equal(predictions(T(c["z"])).cpu().numpy(), ho.argmax_first(c["z"]))
    assert np.array_equal(predictions(T(c["z"]), T(c["iif_raw"])).cpu().numpy(),
                          ho.argmax_first((c["z"] * c["iif_raw"]).astype(np.float32)))


def test_cls_rejects_cpu():
    from iif_b200.classification import IIFLoss
    from iif_b200 import ops
    with pytest.raises(RuntimeError):
        IIFLoss(_DS([5, 3]), device="cpu")
    with pytest.raises(RuntimeError):
        ops.softmax_ce(torch.zeros(2, 3), None, torch.zeros(2, dtype=torch.int64))


# ------------------------------------------------------------------ mmdet surface
@pytest.fixture(scope="module")
def csv1204(golden, tmp_path_factory):
    """Rebuild idf_1204.csv from the frozen columns (the reference tree is absent on the GPU box)."""
    import pandas as pd
    g = golden("weight_tables")
    cols = {k[len("idf_1204_"):]: g[k] for k in g if k.startswith("idf_1204_") and g[k].shape == (1204,)}
    p = tmp_path_factory.mktemp("csv") / "idf_1204.csv"
    pd.DataFrame(cols).to_csv(p, index=False, float_format="%.17g")
    return str(p)


def test_mmdet_iifloss(golden, csv1204):
    from iif_b200 import mmdet as M
    g = golden("mmdet_iif")
    af = float(g["avg_factor"])
    crit = M.IIFLoss(num_classes=1203, path=csv1204, variant="raw")
    assert crit.custom_cls_channels and crit.custom_activation and crit.custom_accuracy
    assert crit.get_cls_channels(1203) == 1204
    with pytest.raises(AssertionError):
        crit.get_cls_channels(80)
    with pytest.raises(AssertionError):
        M.IIFLoss(use_sigmoid=True, path=csv1204)
    with pytest.raises(KeyError):
        M.IIFLoss(path=csv1204, variant="log_adj")
    assert np.array_equal(N(crit.iif_weights), g["iif_raw"])

    def run(tag, rows=None, **kw):
        z = T(g["z"]).requires_grad_(True)
        loss = crit(z, T(g["y"]), **kw)
        ref = g[f"loss_{tag}"]
        if ref.ndim:
            assert rel_err(N(loss), ref) < TOL_F32
            loss.sum().backward()
        else:
            assert float(loss) == pytest.approx(float(ref), rel=TOL_F32)
            loss.backward()
        assert rel_err(N(z.grad)[:rows], g[f"dz_{tag}"]) < TOL_F32

    run("raw_avg", rows=8, weight=T(g["w"]), avg_factor=af)
    run("raw_plain")
    run("raw_w_mean", weight=T(g["w"]))
    run("raw_none", weight=T(g["w"]), reduction_override="none")
    run("raw_sum", weight=T(g["w"]), reduction_override="sum")
    run("raw_none_avg", weight=T(g["w"]), avg_factor=af, reduction_override="none")
    with pytest.raises(ValueError):
        crit(T(g["z"]), T(g["y"]), avg_factor=af, reduction_override="sum")      # losses/utils.py:53-54
    with pytest.raises(AssertionError):
        crit(T(g["z"]), T(g["y"]), reduction_override="bogus")                   # iif_loss.py:132
    # activation / accuracy
    assert rel_err(N(crit.get_activation(T(g["z"]))), g["act_raw"]) < TOL_F32
    acc = crit.get_accuracy(T(g["z"]), T(g["y"]))
    assert np.float32(float(acc["acc_classes"])) == pytest.approx(float(g["acc_raw"][0]), abs=1e-4)
    # empty batch (tests/test_models/test_loss.py:88-101)
    e = crit(torch.zeros(0, 1204, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), avg_factor=1.0)
    assert isinstance(e, torch.Tensor) and float(e) == 0.0
    # loss_weight + class_weight variant
    crit2 = M.IIFLoss(num_classes=1203, path=csv1204, variant="smooth", loss_weight=0.5,
                      class_weight=[float(v) for v in g["class_weight"]])
    z = T(g["z"]).requires_grad_(True)
    loss = crit2(z, T(g["y"]), weight=T(g["w"]), avg_factor=af)
    assert float(loss) == pytest.approx(float(g["loss_smooth_cw_lw"]), rel=TOL_F32)


def test_mmdet_fasa_and_ce(golden, csv1204):
    from iif_b200 import mmdet as M
    g = golden("mmdet_iif")
    af = float(g["avg_factor"])
    crit = M.FasaIIFLoss(use_sigmoid=True, num_classes=1203, path=csv1204, variant="base10_obj")
    z = T(g["z"]).requires_grad_(True)
    loss = crit(z, T(g["y"]), weight=T(g["w"]), avg_factor=af)           # sigmoid: NO iif (fasa_iif_loss.py:35-36)
    assert float(loss) == pytest.approx(float(g["loss_fasa_sigmoid_avg"]), rel=TOL_F32)
    loss.backward()
    assert rel_err(N(z.grad), g["dz_fasa_sigmoid_avg"]) < TOL_F32
    assert rel_err(N(crit.get_activation(T(g["z"]))), g["fasa_act"]) < TOL_F32   # ...but activation does
    # cums (fasa_iif_loss.py:60-71,154-160)
    crit = M.FasaIIFLoss(num_classes=1203, path=csv1204, variant="raw", use_cums=True)
    yc = T(np.maximum(g["y"], 0))
    rets = [float(crit(T(zz), yc)) for zz in (g["z"], g["z"] * 0.5)]
    assert rel_err(np.array(rets), g["fasa_cum_ret"]) < TOL_F32
    assert rel_err(N(crit.cum_losses), g["fasa_cum_losses"]) < TOL_F32
    assert np.array_equal(N(crit.cum_labels), g["fasa_cum_labels"])
    crit.close_cums()
    assert crit.reduction == "mean" and not N(crit.cum_labels).any()
    # plain CrossEntropyLoss softmax / sigmoid (cross_entropy_loss.py:165-249)
    b = golden("mmdet_bce")
    afb = float(b["avg_factor"])
    ce = M.CrossEntropyLoss(use_sigmoid=True)
    for tag, kw in (("plain", {}), ("avg", dict(weight=T(b["w"]), avg_factor=afb)),
                    ("ign255_avg", dict(weight=T(b["w"]), avg_factor=afb, ignore_index=255)),
                    ("none", dict(weight=T(b["w"]), reduction_override="none")),
                    ("sum", dict(weight=T(b["w"]), reduction_override="sum"))):
        z = T(b["z"]).requires_grad_(True)
        loss = ce(z, T(b["y"]), **kw)
        if b[f"loss_{tag}"].ndim:
            assert rel_err(N(loss), b[f"loss_{tag}"]) < TOL_F32
            loss.sum().backward()
        else:
            assert float(loss) == pytest.approx(float(b[f"loss_{tag}"]), rel=TOL_F32)
            loss.backward()
        assert rel_err(N(z.grad), b[f"dz_{tag}"]) < TOL_F32
    ce2 = M.CrossEntropyLoss(use_sigmoid=True, class_weight=[float(v) for v in b["pos_weight"]], loss_weight=2.0)
    loss = ce2(T(b["z"]), T(b["y"]), weight=T(b["w"]), avg_factor=afb)
    assert float(loss) == pytest.approx(float(b["loss_pw_lw_avg"]), rel=TOL_F32)
    ce3 = M.CrossEntropyLoss()
    z = T(b["ce_z"]).requires_grad_(True)
    loss = ce3(z, T(b["ce_y"]), weight=T(b["w"]), avg_factor=afb)
    assert float(loss) == pytest.approx(float(b["ce_loss"]), rel=TOL_F32)
    loss.backward()
    assert rel_err(N(z.grad), b["ce_dz"]) < TOL_F32


def test_mmdet_accuracy_kat():
    """seg/tests/test_metrics/test_losses.py:186-240."""
    from iif_b200.mmdet import accuracy, Accuracy
    pred = T(np.array([[0.2, 0.3, 0.6, 0.5], [0.1, 0.1, 0.2, 0.6], [0.9, 0.0, 0.0, 0.1],
                       [0.4, 0.7, 0.1, 0.1], [0.0, 0.0, 0.99, 0]], np.float32))
    t1 = T(np.array([2, 3, 0, 1, 2], np.int64))
    assert float(Accuracy(topk=1)(pred, t1)) == 100
    assert float(Accuracy(topk=1, thresh=0.8)(pred, t1)) == 40
    assert float(Accuracy(topk=2)(pred, T(np.array([3, 2, 0, 0, 2], np.int64)))) == 100
    a = accuracy(pred, t1, topk=(1, 2))
    assert [float(v) for v in a] == [100.0, 100.0]
    assert float(accuracy(torch.zeros(0, 4, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV))) == 0
    with pytest.raises(AssertionError):
        accuracy(pred, t1, topk=5)


def test_mmdet_linear_fc_cls():
    """fc_cls as an nn.Linear subclass: fp32 parity mode at 1e-5, bf16 tcgen05 mode at 2e-2; .grad
    produced by autograd so DDP hooks fire."""
    from iif_b200.mmdet import Linear
    x, w, b, counts, y = head_inputs(512, 1024, 1204, seed=4, relu=True)
    rng = np.random.default_rng(0)
    gz = (rng.standard_normal((512, 1204)) / 512).astype(np.float32)
    zr = ho.linear_fwd(x, w, b)
    dxr, dwr, dbr = ho.linear_bwd(gz, x, w)
    for compute, tol in (("fp32", TOL_F32), ("bf16", TOL_BF16)):
        fc = Linear(1024, 1204, compute=compute).to(DEV)
        assert isinstance(fc, torch.nn.Linear) and set(dict(fc.named_parameters())) == {"weight", "bias"}
        with torch.no_grad():
            fc.weight.copy_(T(w)); fc.bias.copy_(T(b))
        xt = T(x).requires_grad_(True)
        z = fc(xt)
        assert rel_err(N(z), zr) < tol
        z.backward(T(gz))
        assert rel_err(N(xt.grad), dxr) < tol
        assert rel_err(N(fc.weight.grad), dwr) < tol
        assert rel_err(N(fc.bias.grad), dbr) < tol


# ------------------------------------------------------------------ fused head (bf16 GEMM mode)
@pytest.mark.parametrize("B,D,C,relu", [(256, 2048, 1000, False), (256, 2048, 365, False), (1024, 1024, 1204, True),
                                        (128, 64, 10, False)])
def test_fused_head_bf16(B, D, C, relu):
    from iif_b200 import functional as F_
    from iif_b200.ops import HeadStep
    x, w, b, counts, y = head_inputs(B, D, C, seed=C, relu=relu)
    iif = iif_row(counts, "smooth")
    ref32 = ho.head_fwd_bwd(x, w, b, iif, y)
    ref16 = ho.head_fwd_bwd(bf16_round(x), bf16_round(w), b, iif, y)
    xt = T(x).requires_grad_(True)
    wt = T(w).requires_grad_(True)
    bt = T(b).requires_grad_(True)
    loss, z = F_.iif_head_loss(xt, wt, bt, T(iif), T(y))
    loss.backward()
    for ref, tol in ((ref16, 6e-3), (ref32, TOL_BF16)):     # 6e-3: bf16 storage of dZ
        assert float(loss) == pytest.approx(ref["loss"], rel=tol)
        assert rel_err(N(z), ref["z"]) < tol
        assert rel_err(N(xt.grad), ref["dx"]) < tol
        assert rel_err(N(wt.grad), ref["dw"]) < tol
        assert rel_err(N(bt.grad), ref["db"]) < tol
    # one-call C entry point with pre-allocated buffers (the bench path)
    hs = HeadStep(B, D, C, DEV, want_acc=True, dx_bf16=False)
    l2 = hs.run(T(x, torch.bfloat16), T(w, torch.bfloat16), T(b), T(iif).reshape(-1), T(y))
    torch.cuda.synchronize()
    assert float(l2) == pytest.approx(ref16["loss"], rel=1e-4)
    assert rel_err(N(hs.dw), ref16["dw"]) < 6e-3 and rel_err(N(hs.dx), ref16["dx"]) < 6e-3
    assert rel_err(N(hs.db), ref16["db"]) < 6e-3
    assert np.array_equal(hs.argmax.cpu().numpy(), ho.argmax_first(N(hs.z)))      # bit-exact on its own logits
    assert np.array_equal(hs.rank.cpu().numpy(), ho.label_rank(N(hs.z), y))


@pytest.mark.parametrize("B,D,C", [(256, 2048, 1000), (1024, 1024, 1204), (77, 520, 1204), (512, 512, 4096), (300, 256, 8)])
def test_loss_fused_into_backward_launch(B, D, C):
    """The loss rows riding in the backward launch (iif_loss_linear_bwd_bf16: 2 launches per step) give
    bit-identical dZ / loss_i / argmax / rank and the same loss, dX, dW, db as the 3-launch chain; the
    workspace counters re-arm themselves (several steps back to back on one workspace)."""
    from iif_b200.ops import HeadStep
    from iif_b200 import _lib
    x, w, b, counts, y = head_inputs(B, D, C, seed=B + C)
    iif = iif_row(counts, "raw")
    y[::7] = -100                                     # ignored rows
    bf = torch.bfloat16
    args = (T(x, bf), T(w, bf), T(b), T(iif).reshape(-1), T(y))
    a = HeadStep(B, D, C, DEV, want_acc=True, fused_loss=True, persistent=False)
    u = HeadStep(B, D, C, DEV, want_acc=True, fused_loss=False, persistent=False)
    # these grids exceed what the driver admits as a cooperative launch: only legal with the SMs to ourselves
    _lib.load().iif_gemm_assume_exclusive(1)
    try:
        a.bind(*args); u.bind(*args)
        assert u.launches_per_step == 3
        assert a.launches_per_step == 2, "this shape is expected to qualify for the fused launch"
        for _ in range(3):
            la, lu = a.launch(), u.launch()
        torch.cuda.synchronize()
    finally:
        _lib.load().iif_gemm_assume_exclusive(0)
    assert torch.equal(a.z, u.z) and torch.equal(a.dz[:, :C], u.dz[:, :C]) and torch.equal(a.loss_i, u.loss_i)
    assert torch.equal(a.argmax, u.argmax) and torch.equal(a.rank, u.rank) and torch.equal(a.acc_counts, u.acc_counts)
    assert float(la) == pytest.approx(float(lu), rel=1e-6)
    assert torch.equal(a.dw, u.dw) and torch.equal(a.db, u.db)
    assert rel_err(N(a.dx), N(u.dx)) < 1e-6           # the K split of dX may differ between the two grids


def test_head_step_eager_back_to_back_rotating_sets():
    """Many eager steps over rotating input sets sharing ONE workspace (programmatic dependent launch lets
    consecutive kernels overlap on the GPU): every set must keep reproducing its first result."""
    from iif_b200.ops import HeadStep
    B, D, C = 256, 2048, 1000
    bf = torch.bfloat16
    ws = torch.zeros(int(HeadStep(B, D, C, DEV).ws_bytes), dtype=torch.uint8, device=DEV)
    sets = []
    for seed in range(6):
        x, w, b, counts, y = head_inputs(B, D, C, seed=seed)
        hs = HeadStep(B, D, C, DEV, ws=ws)
        hs.bind(T(x, bf), T(w, bf), T(b), T(iif_row(counts, "smooth")).reshape(-1), T(y))
        sets.append(hs)
    first = []
    for hs in sets:
        hs.launch()
        torch.cuda.synchronize()
        first.append((float(hs.loss), hs.dw.clone(), hs.dx.clone(), hs.z.clone()))
    for i in range(600):
        sets[i % 6].launch()
    torch.cuda.synchronize()
    for hs, (l0, dw0, dx0, z0) in zip(sets, first):
        assert float(hs.loss) == l0 and torch.equal(hs.dw, dw0) and torch.equal(hs.dx, dx0) and torch.equal(hs.z, z0)


def test_head_pipeline_host_batches():
    """ops.HeadPipeline (iif_pipeline_*): steps fed from pinned HOST batches through the three-stream pipeline
    return exactly the loss / gradients of the same step run directly on device tensors, slots are reused
    safely (more batches than slots), and host tensors of the wrong kind are rejected."""
    from iif_b200.ops import HeadStep, HeadPipeline
    B, D, C = 256, 512, 1000
    bf = torch.bfloat16
    x0, w, b, counts, y0 = head_inputs(B, D, C, seed=3)
    iif = T(iif_row(counts, "smooth")).reshape(-1)
    wt, bt = T(w, bf), T(b)
    nslots, nbatch = 3, 8
    slots = []
    for _ in range(nslots):
        hs = HeadStep(B, D, C, DEV, want_acc=True)
        hs.bind(torch.zeros(B, D, dtype=bf, device=DEV), wt, bt, iif, torch.zeros(B, dtype=torch.int64, device=DEV))
        slots.append(hs)
    pipe = HeadPipeline(slots)
    rng = np.random.default_rng(0)
    hx = [torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)).to(bf).pin_memory() for _ in range(nbatch)]
    hy = [torch.from_numpy(lt_labels(counts, B, rng)).pin_memory() for _ in range(nbatch)]
    ref = HeadStep(B, D, C, DEV, want_acc=True)
    want = []
    for i in range(nbatch):
        ref.bind(hx[i].to(DEV), wt, bt, iif, hy[i].to(DEV))
        ref.launch()
        torch.cuda.synchronize()
        want.append((float(ref.loss), ref.dw.clone(), ref.dx.clone(), ref.acc_counts.clone()))
    got = []
    for i in range(nbatch):
        k = i % nslots
        if i >= nslots:
            got.append((pipe.wait(k), slots[k].dw.clone(), slots[k].dx.clone(), slots[k].acc_counts.clone()))
            torch.cuda.synchronize()                  # the copies above run on torch's stream, the pipeline on its own
        pipe.submit(k, hx[i], hy[i])
    for i in range(nbatch - nslots, nbatch):
        k = i % nslots
        got.append((pipe.wait(k), slots[k].dw.clone(), slots[k].dx.clone(), slots[k].acc_counts.clone()))
    pipe.sync()
    assert len(got) == nbatch
    for (l, dw, dx, acc), (l0, dw0, dx0, acc0) in zip(got, want):
        assert l == l0 and torch.equal(dw, dw0) and torch.equal(dx, dx0) and torch.equal(acc, acc0)
    with pytest.raises(RuntimeError):
        pipe.submit(0, hx[0].to(DEV), hy[0])          # device tensor where a host batch is expected
    with pytest.raises(ValueError):
        pipe.submit(0, hx[0].float(), hy[0])
    pipe.close()


# ------------------------------------------------------------------ normalised classifiers
def _load(mod, g, tag):
    with torch.no_grad():
        mod.weight.copy_(T(g[f"{tag}_w"]).float())
        if getattr(mod, "bias", None) is not None:
            mod.bias.copy_(T(g[f"{tag}_b"]).float())


def _check_normed(mod, g, tag, tol):
    x = T(g[f"{tag}_x"]).float().requires_grad_(True)
    z = mod(x)
    z.backward(T(g[f"{tag}_gz"]).float())
    assert rel_err(N(z), g[f"{tag}_z"]) < tol
    assert rel_err(N(x.grad), g[f"{tag}_dx"]) < tol
    assert rel_err(N(mod.weight.grad), g[f"{tag}_dw"]) < tol
    if getattr(mod, "bias", None) is not None:
        assert rel_err(N(mod.bias.grad), g[f"{tag}_db"]) < tol


@pytest.mark.parametrize("compute,tol", [("fp32", 2e-5), ("bf16", TOL_BF16)])
def test_normed_linear_modules_golden(golden, tmp_path, compute, tol):
    """NormedLinear / IIFNormedLinear / CosNorm_Classifier against the unmodified reference modules
    (tests/golden/normed.npz, float64): forward, dX, dW, db."""
    from iif_b200.mmdet import NormedLinear, IIFNormedLinear
    from iif_b200.classification import CosNorm_Classifier
    g = golden("normed")
    m = NormedLinear(64, 36, compute=compute).to(DEV)
    assert isinstance(m, torch.nn.Linear) and m.tempearture == 20
    _load(m, g, "normed_p1"); _check_normed(m, g, "normed_p1", tol)
    m = NormedLinear(64, 36, tempearture=10, power=2.0, eps=1e-3, compute=compute).to(DEV)
    _load(m, g, "normed_p2"); _check_normed(m, g, "normed_p2", tol)
    m = NormedLinear(64, 36, bias=False, compute=compute).to(DEV)
    _load(m, g, "normed_nobias"); _check_normed(m, g, "normed_nobias", tol)
    # the CSV the module reads: placeholder row 0, then the per-class column; the module appends 1.0 for background
    import pandas as pd
    col = np.concatenate([[1.0], g["iifnormed_iif"][:-1]])
    path = str(tmp_path / "idf.csv")
    pd.DataFrame({"base2_obj": col}).to_csv(path, index=False)
    m = IIFNormedLinear(32, 1204, path=path, variant="base2_obj", compute=compute, device=DEV).to(DEV)
    assert np.array_equal(N(m.iif_weights).reshape(-1), g["iifnormed_iif"].astype(np.float32))
    _load(m, g, "iifnormed"); _check_normed(m, g, "iifnormed", tol)
    with pytest.raises(KeyError):
        IIFNormedLinear(32, 1204, path=path, variant="log_adj", device=DEV)
    c = CosNorm_Classifier(64, 36, scale=16, compute=compute, device=DEV)
    _load(c, g, "cosnorm"); _check_normed(c, g, "cosnorm", tol)


def test_normed_linear_head_shape_vs_oracle():
    """LVIS-sized NormedLinear head (1024 RoIs x 1024-d x 1204 classes, bf16 GEMMs) vs the float64 oracle, and the
    row kernels on ragged / unaligned rows."""
    from iif_b200.mmdet import NormedLinear
    from iif_b200 import ops, _lib
    rng = np.random.default_rng(0)
    B, D, C = 1024, 1024, 1204
    x = np.maximum(rng.standard_normal((B, D)), 0).astype(np.float32)
    w = (rng.standard_normal((C, D)) * 0.01).astype(np.float32)
    b = (rng.standard_normal(C) * 0.01).astype(np.float32)
    gz = (rng.standard_normal((B, C)) / B).astype(np.float32)
    m = NormedLinear(D, C, compute="bf16").to(DEV)
    with torch.no_grad():
        m.weight.copy_(T(w)); m.bias.copy_(T(b))
    xt = T(x).requires_grad_(True)
    z = m(xt)
    z.backward(T(gz))
    zr, dxr, dwr, dbr = ho.normed_linear(x, w, b, gz)
    assert rel_err(N(z), zr) < TOL_BF16 and rel_err(N(xt.grad), dxr) < TOL_BF16
    assert rel_err(N(m.weight.grad), dwr) < TOL_BF16 and rel_err(N(m.bias.grad), dbr) < TOL_BF16
    # row kernels alone, 7 x 13 (scalar path) with a zero row
    u = rng.standard_normal((7, 13)).astype(np.float32); u[3] = 0
    v = rng.standard_normal((7, 13)).astype(np.float32)
    pre = rng.uniform(-2, 2, 7).astype(np.float32)
    a, c = ops.row_scale_from_norm(T(u), _lib.NORM_NORMED, pre=T(pre), temperature=3.0, power=1.5, eps=1e-4)
    _, ar, cr = ho._norm_rows(u, pre, "normed", 3.0, 1.5, 1e-4)
    assert rel_err(N(a), ar) < 1e-5 and rel_err(N(c), cr) < 1e-5
    assert rel_err(N(ops.row_dot(T(u), T(v))), (u.astype(np.float64) * v).sum(1)) < 1e-5
    out = ops.rows_axpby(T(u), T(pre), T(v), T(pre), T(pre))
    assert rel_err(N(out), pre[:, None] * u + (pre * pre)[:, None] * v) < 1e-6


def test_head_pipeline_staged_mode():
    """Staged mode (one CUDA graph per slot, next slot's batch prefetched inside the graph, loss stored by the
    kernel into mapped pinned memory) reproduces the direct step for every batch, in order and out of order."""
    from iif_b200.ops import HeadStep, HeadPipeline
    B, D, C = 256, 512, 1000
    bf = torch.bfloat16
    x0, w, b, counts, y0 = head_inputs(B, D, C, seed=4)
    iif = T(iif_row(counts, "raw")).reshape(-1)
    wt, bt = T(w, bf), T(b)
    nslots, nbatch = 3, 9
    slots = []
    for _ in range(nslots):
        hs = HeadStep(B, D, C, DEV)
        hs.bind(torch.zeros(B, D, dtype=bf, device=DEV), wt, bt, iif, torch.zeros(B, dtype=torch.int64, device=DEV))
        slots.append(hs)
    pipe = HeadPipeline(slots)
    pipe.enable_staged()
    rng = np.random.default_rng(1)
    batches = [(torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)).to(bf),
                torch.from_numpy(lt_labels(counts, B, rng))) for _ in range(nbatch)]
    ref = HeadStep(B, D, C, DEV)
    want = []
    for xb, yb in batches:
        ref.bind(xb.to(DEV), wt, bt, iif, yb.to(DEV))
        ref.launch()
        torch.cuda.synchronize()
        want.append((float(ref.loss), ref.dw.clone()))

    def stage(k, i):
        sx, sy = pipe.staging(k)
        sx.copy_(batches[i][0]); sy.copy_(batches[i][1])

    stage(0, 0)
    for i in range(nbatch):                      # round-robin, the loader one batch ahead
        k = i % nslots
        if i + 1 < nbatch:
            pipe.wait((i + 1) % nslots)          # the slot about to be restaged is idle
            stage((i + 1) % nslots, i + 1)
        pipe.submit_staged(k)
        loss = pipe.wait(k)
        torch.cuda.synchronize()
        assert loss == want[i][0] and torch.equal(slots[k].dw, want[i][1])
    stage(2, 4)                                  # out of order: slot 2 right after slot 2's turn was skipped
    pipe.submit_staged(2)
    assert pipe.wait(2) == want[4][0]
    pipe.close()
