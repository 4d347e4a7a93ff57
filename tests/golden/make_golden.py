"""Freeze golden vectors by running the UNMODIFIED reference Python (build container only).

    python tests/golden/make_golden.py          # needs /root/reference

Writes ``tests/golden/*.npz``.  Inputs are seeded (torch.manual_seed / numpy default_rng) and
stored next to the reference's outputs, so the fixtures are self-contained on the GPU box where
/root/reference does not exist.  Reference entry points exercised:

* ``classification/custom.py``: IIFLoss (all 7 variants, reductions, iif_norm, class weight,
  infer=True), FocalLoss(gamma=0) (via the cpu shim for ``torch.cuda.FloatTensor``)
* ``mmdet/models/losses/iif_loss.py``: IIFLoss.forward / get_activation / get_accuracy
* ``mmdet/models/losses/fasa_iif_loss.py``: FasaIIFLoss softmax, sigmoid, use_cums
* ``mmdet/models/losses/cross_entropy_loss.py``: CrossEntropyLoss(use_sigmoid=True)
* ``lvis_files/idf_1204.csv``, ``idf_1231.csv``, ``coco_files/idf_91.csv`` weight tables
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

assert ref_loader.available(), "reference tree not found"
custom = ref_loader.load_classification()
mm = ref_loader.load_mmdet_losses()

VARIANTS = ["raw", "smooth", "rel", "normit", "gombit", "base2", "base10"]
CSV_COLS = ["smooth", "raw", "prob", "normit", "gombit", "base2", "base10"]
CSV_COLS = CSV_COLS + [c + "_obj" for c in CSV_COLS]


class _DS:
    def __init__(self, counts):
        self.c = list(counts)

    def get_cls_num_list(self):
        return self.c


def npy(t):
    return t.detach().cpu().numpy().copy()


def lt_labels(counts, n, gen):
    p = torch.tensor(counts, dtype=torch.float64)
    return torch.multinomial(p / p.sum(), n, replacement=True, generator=gen)


# ------------------------------------------------------------------ classification Mixup (custom.py:91-117)
def gen_cls_mixup():
    """custom.Mixup.mixup_criterion around custom.IIFLoss: lam*CE(y_a) + (1-lam)*CE(y_b), loss and d/dlogits,
    on a 4-divisible class count (the fused kernel's path) with a long-tailed profile."""
    g = torch.Generator().manual_seed(7)
    C, B = 100, 96
    counts = [max(int(500 * (0.01) ** (c / (C - 1.0))), 1) for c in range(C)]
    z0 = torch.randn(B, C, generator=g) * 2.0
    y_a = lt_labels(counts, B, g)
    y_b = y_a[torch.randperm(B, generator=g)]
    cw = torch.tensor(counts, dtype=torch.float32)
    cw = cw.sum() / cw
    out = dict(counts=np.array(counts), z=npy(z0), y_a=npy(y_a), y_b=npy(y_b), cw=npy(cw))
    for variant in ("raw", "smooth"):
        for reduction in ("mean", "sum"):
            for use_cw in (False, True):
                for lam in (0.3, 1.0):
                    crit = custom.IIFLoss(_DS(counts), variant=variant, reduction=reduction, device="cpu",
                                          weight=cw if use_cw else None)
                    mix = custom.Mixup(crit, alpha=0.2)
                    z = z0.clone().requires_grad_(True)
                    loss = mix.mixup_criterion(z, y_a, y_b, lam)
                    loss.backward()
                    tag = f"{variant}_{reduction}_{'cw' if use_cw else 'nocw'}_{lam}"
                    out[f"loss_{tag}"] = npy(loss)
                    out[f"dz_{tag}"] = npy(z.grad)
    np.savez_compressed(os.path.join(HERE, "cls_mixup.npz"), **out)
    print("cls_mixup.npz", len(out), "arrays")


# ------------------------------------------------------------------ classification IIFLoss
def gen_cls_iif():
    g = torch.Generator().manual_seed(0)
    counts = [5000, 2997, 1796, 1077, 645, 387, 232, 139, 83, 50]  # imbalanced_dataset.py:23-29, r=100
    B, D, C = 128, 64, 10
    x = torch.randn(B, D, generator=g)
    w = (torch.rand(C, D, generator=g) * 2 - 1) / D ** 0.5
    b = torch.full((C,), 0.01)
    y = lt_labels(counts, B, g)
    cw = torch.tensor(counts, dtype=torch.float32)
    cw = cw.sum() / cw  # initialisers.get_weights (--deffered)
    out = dict(counts=np.array(counts), x=npy(x), w=npy(w), b=npy(b), y=npy(y), cw=npy(cw))

    def run(crit, tag, reduction):
        xx = x.clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        bb = b.clone().requires_grad_(True)
        z = torch.nn.functional.linear(xx, ww, bb)
        z.retain_grad()
        loss = crit(z, y)
        (loss.sum() if reduction == "none" else loss).backward()
        out[f"loss_{tag}"] = npy(loss)
        out[f"dz_{tag}"] = npy(z.grad)
        out[f"dx_{tag}"] = npy(xx.grad)
        out[f"dw_{tag}"] = npy(ww.grad)
        out[f"db_{tag}"] = npy(bb.grad)
        out["z"] = npy(z)

    for v in VARIANTS:
        crit = custom.IIFLoss(_DS(counts), variant=v, device="cpu")
        out[f"iif_{v}"] = npy(crit.iif[v])
        run(crit, f"{v}_mean", "mean")
        out[f"infer_{v}"] = npy(crit(torch.from_numpy(out["z"]), infer=True))
        critn = custom.IIFLoss(_DS(counts), variant=v, iif_norm=2, device="cpu")
        out[f"iifn2_{v}"] = npy(critn.iif[v])
    for r in ("sum", "none"):
        for v in ("raw", "smooth"):
            run(custom.IIFLoss(_DS(counts), variant=v, reduction=r, device="cpu"), f"{v}_{r}", r)
    run(custom.IIFLoss(_DS(counts), variant="smooth", device="cpu", weight=cw), "smooth_cw_mean", "mean")
    run(custom.IIFLoss(_DS(counts), variant="raw", iif_norm=2, device="cpu"), "raw_n2_mean", "mean")
    np.savez_compressed(os.path.join(HERE, "cls_iif.npz"), **out)


# ------------------------------------------------------------------ classification BCE (FocalLoss gamma=0)
def gen_cls_bce():
    g = torch.Generator().manual_seed(1)
    B, C = 64, 37
    z0 = torch.randn(B, C, generator=g) * 3
    y = torch.randint(0, C, (B,), generator=g)
    wts = torch.rand(C, generator=g) + 0.5
    out = dict(z=npy(z0), y=npy(y), weights=npy(wts))
    with ref_loader.cpu_shims():
        for red in ("mean", "sum"):
            for tag, ww in (("now", None), ("w", wts)):
                crit = custom.FocalLoss(gamma=0, reduction=red, device="cpu", weights=ww)
                z = z0.clone().requires_grad_(True)
                loss = crit(z, y)
                loss.backward()
                out[f"loss_{tag}_{red}"] = npy(loss)
                out[f"dz_{tag}_{red}"] = npy(z.grad)
    np.savez_compressed(os.path.join(HERE, "cls_bce.npz"), **out)


def gen_cls_focal():
    """custom.FocalLoss(gamma > 0) (custom.py:74-89).  The reference only runs in fp32 (its one-hot tensor is a
    FloatTensor) and goes sigmoid -> log, so the logits are kept moderate (std 1.5): there its own chain is
    accurate to ~1e-6 and pins the oracle; alpha on / off, per-class weights on / off."""
    g = torch.Generator().manual_seed(2)
    B, C = 48, 40
    z0 = torch.randn(B, C, generator=g) * 1.5
    y = torch.randint(0, C, (B,), generator=g)
    wts = torch.rand(C, generator=g) + 0.5
    out = dict(z=npy(z0), y=npy(y), weights=npy(wts))
    with ref_loader.cpu_shims():
        for gamma in (2.0, 0.5):
            for alpha in (None, 0.25):
                for red in ("mean", "sum"):
                    for tag, ww in (("now", None), ("w", wts)):
                        crit = custom.FocalLoss(gamma=gamma, alpha=alpha, reduction=red, device="cpu", weights=ww)
                        z = z0.clone().requires_grad_(True)
                        loss = crit(z, y)
                        loss.backward()
                        key = f"{gamma}_{alpha}_{tag}_{red}"
                        out[f"loss_{key}"] = npy(loss)
                        out[f"dz_{key}"] = npy(z.grad)
    np.savez_compressed(os.path.join(HERE, "cls_focal.npz"), **out)
    print("cls_focal.npz", len(out), "arrays")


# ------------------------------------------------------------------ normalised classifiers
def gen_normed():
    """mmdet NormedLinear / IIFNormedLinear (utils/normed_predictor.py) and cls CosNorm_Classifier
    (resnet_cifar.py:50-78), unmodified, in float64: forward and the gradients of x, weight, bias for a
    random upstream gradient."""
    g = torch.Generator().manual_seed(5)
    out = {}
    with ref_loader.cpu_shims():
        npred = ref_loader.load_normed_predictors()
        rc = ref_loader.load_resnet_cifar()

        def run(tag, mod, B, D):
            mod = mod.double()
            with torch.no_grad():
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g, dtype=torch.float64) * 0.1)
                if getattr(mod, "bias", None) is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g, dtype=torch.float64) * 0.1)
            x = (torch.randn(B, D, generator=g, dtype=torch.float64) * 1.5).requires_grad_(True)
            z = mod(x)
            gz = torch.randn(z.shape, generator=g, dtype=torch.float64)
            z.backward(gz)
            out[f"{tag}_x"], out[f"{tag}_w"], out[f"{tag}_gz"], out[f"{tag}_z"] = npy(x), npy(mod.weight), npy(gz), npy(z)
            out[f"{tag}_dx"], out[f"{tag}_dw"] = npy(x.grad), npy(mod.weight.grad)
            if getattr(mod, "bias", None) is not None:
                out[f"{tag}_b"], out[f"{tag}_db"] = npy(mod.bias), npy(mod.bias.grad)

        run("normed_p1", npred.NormedLinear(64, 36), 24, 64)
        run("normed_p2", npred.NormedLinear(64, 36, tempearture=10, power=2.0, eps=1e-3), 24, 64)
        run("normed_nobias", npred.NormedLinear(64, 36, bias=False), 24, 64)
        iifm = npred.IIFNormedLinear(32, 1204, path=ref_loader.csv_path("idf_1204.csv"), variant="base2_obj")
        out["iifnormed_iif"] = npy(iifm.iif_weights.reshape(-1))
        run("iifnormed", iifm, 16, 32)
        run("cosnorm", rc.CosNorm_Classifier(64, 36, scale=16), 24, 64)
    np.savez_compressed(os.path.join(HERE, "normed.npz"), **out)
    print("normed.npz", len(out), "arrays")


# ------------------------------------------------------------------ mmdet IIFLoss / FasaIIFLoss
def gen_mmdet():
    g = torch.Generator().manual_seed(2)
    path = ref_loader.csv_path("idf_1204.csv")
    B, C = 32, 1204
    z0 = torch.randn(B, C, generator=g) * 2
    y = torch.randint(0, 1203, (B,), generator=g)
    y[torch.rand(B, generator=g) < 0.6] = 1203  # background
    y[5] = -100
    y[17] = -100
    wrow = (torch.rand(B, generator=g) > 0.15).float() * (0.5 + torch.rand(B, generator=g))
    avg = max(float((wrow > 0).sum()), 1.0)  # bbox_head.py:267
    cwl = (0.5 + torch.rand(C, generator=g)).tolist()
    out = dict(z=npy(z0), y=npy(y), w=npy(wrow), avg_factor=np.float64(avg), class_weight=np.array(cwl, np.float64))

    def run(crit, tag, yy=None, rows=None, **kw):
        z = z0.clone().requires_grad_(True)
        loss = crit(z, y if yy is None else yy, **kw)
        loss.sum().backward()
        out[f"loss_{tag}"] = npy(loss)
        out[f"dz_{tag}"] = npy(z.grad)[:rows]  # the 14-column sweep keeps the first 8 rows only

    with ref_loader.cpu_shims():
        for col in CSV_COLS:
            crit = mm.iif_loss.IIFLoss(path=path, variant=col, num_classes=1203)
            out[f"iif_{col}"] = npy(crit.iif_weights)
            run(crit, f"{col}_avg", rows=8, weight=wrow, avg_factor=avg)
        crit = mm.iif_loss.IIFLoss(path=path, variant="raw", num_classes=1203)
        run(crit, "raw_plain")
        run(crit, "raw_w_mean", weight=wrow)
        run(crit, "raw_none", weight=wrow, reduction_override="none")
        run(crit, "raw_sum", weight=wrow, reduction_override="sum")
        run(crit, "raw_none_avg", weight=wrow, avg_factor=avg, reduction_override="none")
        run(crit, "raw_ignbg", yy=y.clamp(min=0), weight=wrow, avg_factor=avg, ignore_index=1203)
        out["act_raw"] = npy(crit.get_activation(z0))
        out["acc_raw"] = npy(crit.get_accuracy(z0, y)["acc_classes"])
        acc15 = mm.accuracy.accuracy(z0, y.clamp(min=0), topk=(1, 5))
        out["acc_top1"] = npy(acc15[0])
        out["acc_top5"] = npy(acc15[1])
        out["topk5_idx"] = npy(z0.topk(5, dim=1)[1])
        crit = mm.iif_loss.IIFLoss(path=path, variant="smooth", num_classes=1203, class_weight=cwl, loss_weight=0.5)
        run(crit, "smooth_cw_lw", weight=wrow, avg_factor=avg)
        crit = mm.iif_loss.IIFLoss(path=path, variant="normit_obj", num_classes=1203, reduction="sum")
        run(crit, "normit_obj_sum")
        # FasaIIFLoss: softmax (IIF applied), sigmoid (no IIF), cums
        crit = mm.fasa_iif_loss.FasaIIFLoss(path=path, variant="base10_obj", num_classes=1203)
        run(crit, "fasa_base10_obj_avg", weight=wrow, avg_factor=avg)
        out["fasa_act"] = npy(crit.get_activation(z0))
        crit = mm.fasa_iif_loss.FasaIIFLoss(path=path, variant="base10_obj", num_classes=1203, use_sigmoid=True)
        run(crit, "fasa_sigmoid_avg", weight=wrow, avg_factor=avg)
        crit = mm.fasa_iif_loss.FasaIIFLoss(path=path, variant="raw", num_classes=1203, use_cums=True)
        ypos = y.clamp(min=0)
        l1 = crit(z0, ypos)
        l2 = crit(z0 * 0.5, ypos)
        out["fasa_cum_losses"] = npy(crit.cum_losses)
        out["fasa_cum_labels"] = npy(crit.cum_labels)
        out["fasa_cum_ret"] = np.array([float(l1), float(l2)])
    np.savez_compressed(os.path.join(HERE, "mmdet_iif.npz"), **out)


# ------------------------------------------------------------------ mmdet sigmoid BCE
def gen_mmdet_bce():
    g = torch.Generator().manual_seed(3)
    B, C = 24, 1203
    z0 = torch.randn(B, C, generator=g) * 2
    y = torch.randint(0, C, (B,), generator=g)
    y[torch.rand(B, generator=g) < 0.5] = C  # background: no positive column
    y[3] = 255
    y[9] = -100
    wrow = (torch.rand(B, generator=g) > 0.2).float() * (0.5 + torch.rand(B, generator=g))
    avg = max(float((wrow > 0).sum()), 1.0)
    pw = (0.5 + torch.rand(C, generator=g)).tolist()
    out = dict(z=npy(z0), y=npy(y), w=npy(wrow), avg_factor=np.float64(avg), pos_weight=np.array(pw, np.float64))

    def run(crit, tag, **kw):
        z = z0.clone().requires_grad_(True)
        loss = crit(z, y, **kw)
        loss.sum().backward()
        out[f"loss_{tag}"] = npy(loss)
        out[f"dz_{tag}"] = npy(z.grad)

    CE = mm.cross_entropy_loss.CrossEntropyLoss
    run(CE(use_sigmoid=True), "plain")
    run(CE(use_sigmoid=True), "avg", weight=wrow, avg_factor=avg)
    run(CE(use_sigmoid=True, ignore_index=255), "ign255_avg", weight=wrow, avg_factor=avg)
    run(CE(use_sigmoid=True, class_weight=pw, loss_weight=2.0), "pw_lw_avg", weight=wrow, avg_factor=avg)
    run(CE(use_sigmoid=True), "none", weight=wrow, reduction_override="none")
    run(CE(use_sigmoid=True), "sum", weight=wrow, reduction_override="sum")
    # softmax CrossEntropyLoss == IIF with all-ones scale
    zc = z0[:, :40].contiguous()
    yc = torch.randint(0, 40, (B,), generator=g)
    z = zc.clone().requires_grad_(True)
    loss = CE()(z, yc, weight=wrow, avg_factor=avg)
    loss.backward()
    out.update(ce_z=npy(zc), ce_y=npy(yc), ce_loss=npy(loss), ce_dz=npy(z.grad))
    np.savez_compressed(os.path.join(HERE, "mmdet_bce.npz"), **out)


# ------------------------------------------------------------------ widened rows of round 2
def gen_widen():
    """cls NormedLinear / CosNorm_Classifier(lr_scale=True) (resnet_cifar.py:38-78), shot_acc (per_shot_acc.py:62-105),
    mmdet binary_cross_entropy with already-expanded labels (cross_entropy_loss.py:100-106), FasaIIFLoss cums in sigmoid
    mode and with a negative label (fasa_iif_loss.py:154-160), FasaBBoxHead.fa_update (fasa_bbox_head.py:118-148) --
    all from the unmodified reference source."""
    g = torch.Generator().manual_seed(11)
    out = {}
    with ref_loader.cpu_shims():
        rc = ref_loader.load_resnet_cifar()

        def run(tag, mod, B, D, wscale):
            mod = mod.double()
            with torch.no_grad():
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g, dtype=torch.float64) * wscale)
            x = (torch.randn(B, D, generator=g, dtype=torch.float64) * 1.5).requires_grad_(True)
            z = mod(x)
            gz = torch.randn(z.shape, generator=g, dtype=torch.float64)
            z.backward(gz)
            out[f"{tag}_x"], out[f"{tag}_w"], out[f"{tag}_gz"], out[f"{tag}_z"] = npy(x), npy(mod.weight), npy(gz), npy(z)
            out[f"{tag}_dx"], out[f"{tag}_dw"] = npy(x.grad), npy(mod.weight.grad)
            return mod

        run("cls_normed", rc.NormedLinear(64, 36), 24, 64, 0.3)
        m = run("cosnorm_lr", rc.CosNorm_Classifier(64, 36, lr_scale=True), 24, 64, 0.1)
        out["cosnorm_lr_scale"], out["cosnorm_lr_dscale"] = npy(m.scale), npy(m.scale.grad)

    # ---- shot_acc
    shot = ref_loader.load_functions("classification/per_shot_acc.py", ["shot_acc"])["shot_acc"]
    C = 40
    counts = [max(int(400 * (0.005) ** (c / (C - 1.0))), 1) for c in range(C)]
    train = np.repeat(np.arange(C), counts)
    labels = torch.randint(0, C, (600,), generator=g)
    labels[labels == 7] = 8                                  # a class absent from the test labels
    preds = torch.where(torch.rand(600, generator=g) < 0.6, labels, torch.randint(0, C, (600,), generator=g))
    many, med, low, cacc = shot(preds, labels, train, acc_per_cls=True)
    out.update(shot_counts=np.array(counts), shot_labels=npy(labels), shot_preds=npy(preds),
               shot_out=np.array([many, med, low], np.float64), shot_class_acc=np.array(cacc, np.float64))
    m2 = shot(preds, labels, train, many_shot_thr=1000, low_shot_thr=2)
    out["shot_out_thr"] = np.array(m2, np.float64)           # empty many / low groups -> 0

    # ---- dense-label BCE
    B, C = 12, 37
    z0 = torch.randn(B, C, generator=g) * 2
    tgt = (torch.rand(B, C, generator=g) < 0.2).float()
    soft = torch.rand(B, C, generator=g)
    wel = torch.rand(B, C, generator=g)
    wrow = torch.rand(B, 1, generator=g)
    pw = (0.5 + torch.rand(C, generator=g))
    out.update(bced_z=npy(z0), bced_t=npy(tgt), bced_soft=npy(soft), bced_wel=npy(wel), bced_wrow=npy(wrow), bced_pw=npy(pw))
    bce = mm.cross_entropy_loss.binary_cross_entropy

    def run_bce(tag, label, **kw):
        z = z0.clone().requires_grad_(True)
        loss = bce(z, label, **kw)
        loss.sum().backward()
        out[f"bced_loss_{tag}"], out[f"bced_dz_{tag}"] = npy(loss), npy(z.grad)

    run_bce("mean", tgt)
    run_bce("soft_sum", soft, reduction="sum")
    run_bce("wel_avg", tgt, weight=wel, avg_factor=5.0)
    run_bce("wrow_none", tgt, weight=wrow.expand(B, C), reduction="none")
    run_bce("pw_mean", tgt, class_weight=pw)

    # ---- FASA cums: sigmoid mode ([B,C] loss rows are summed) and a negative label (python indexing from the end)
    with ref_loader.cpu_shims():
        path = ref_loader.csv_path("idf_1204.csv")
        Bc, Cc = 16, 1204
        zc = torch.randn(Bc, Cc, generator=g)
        yc = torch.randint(0, 1203, (Bc,), generator=g)
        yc[3] = yc[5]
        yc[9] = 1203
        out.update(cum_z=npy(zc), cum_y=npy(yc))
        crit = mm.fasa_iif_loss.FasaIIFLoss(path=path, variant="raw", num_classes=1203, use_cums=True, use_sigmoid=True)
        r1 = crit(zc, yc)
        out["cum_sig_losses"], out["cum_sig_labels"], out["cum_sig_ret"] = npy(crit.cum_losses), npy(crit.cum_labels), npy(r1)
        crit = mm.fasa_iif_loss.FasaIIFLoss(path=path, variant="raw", num_classes=1203, use_cums=True)
        yn = yc.clone()
        yn[2] = -100                                         # ignored by the loss, binned at cum[-100] by the reference
        r2 = crit(zc, yn)
        out["cum_neg_y"] = npy(yn)
        out["cum_neg_losses"], out["cum_neg_labels"], out["cum_neg_ret"] = npy(crit.cum_losses), npy(crit.cum_labels), npy(r2)

    # ---- FASA feature statistics: the two methods of FasaBBoxHead, run as they stand on a bare object
    fns = ref_loader.load_functions("instance_segmentation/mmdet/models/roi_heads/bbox_heads/fasa_bbox_head.py",
                                    ["fa_update", "fa_update_push"], class_name="ConvFCFASABBoxHead")

    class _Head:
        pass

    hd = _Head()
    nb, D = 21, 48
    hd.decay_ratio = 0.1
    hd.feature_mean = torch.zeros(nb, D)
    hd.feature_std = torch.zeros(nb, D)
    hd.feature_used = torch.zeros(nb)
    hd.fa_update_push = lambda e, l: fns["fa_update_push"](hd, e, l)
    emb1 = torch.randn(64, D, generator=g)
    lab1 = torch.randint(0, 12, (64,), generator=g)
    lab1[0] = 19                                             # a class with a single row: variance 0
    emb2 = torch.randn(40, D, generator=g) + 0.5
    lab2 = torch.randint(4, 20, (40,), generator=g)
    fns["fa_update"](hd, emb1, lab1)
    out.update(fa_emb1=npy(emb1), fa_lab1=npy(lab1), fa_mean1=npy(hd.feature_mean), fa_std1=npy(hd.feature_std),
               fa_used1=npy(hd.feature_used))
    fns["fa_update"](hd, emb2, lab2)
    out.update(fa_emb2=npy(emb2), fa_lab2=npy(lab2), fa_mean2=npy(hd.feature_mean), fa_std2=npy(hd.feature_std),
               fa_used2=npy(hd.feature_used), fa_decay=np.float64(0.1))
    np.savez_compressed(os.path.join(HERE, "widen.npz"), **out)
    print("widen.npz", len(out), "arrays")


# ------------------------------------------------------------------ CSV weight tables
def gen_tables():
    out = {}
    for name, n_img in (("idf_1204.csv", 100170), ("idf_1231.csv", 57263), ("idf_91.csv", 118287)):
        df = pd.read_csv(ref_loader.csv_path(name))
        key = name.split(".")[0]
        out[f"{key}_img_freq"] = df["img_freq"].values[1:].astype(np.int64)
        out[f"{key}_instance_freq"] = df["instance_freq"].values[1:].astype(np.int64)
        out[f"{key}_n_img"] = np.int64(n_img)
        for col in CSV_COLS:
            out[f"{key}_{col}"] = df[col].values.astype(np.float64)  # row 0 = placeholder kept
    np.savez_compressed(os.path.join(HERE, "weight_tables.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    gen_cls_iif()
    gen_cls_mixup()
    gen_cls_bce()
    gen_cls_focal()
    gen_normed()
    gen_mmdet()
    gen_mmdet_bce()
    gen_tables()
    gen_widen()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
