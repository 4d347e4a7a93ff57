"""float64 numpy restatement of the IIF classifier head (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
``/root/reference``; ``cls/`` = ``classification/``, ``seg/`` =
``instance_segmentation/``).  The arithmetic of the reference lives in PyTorch
(``F.linear``, ``F.cross_entropy``, ``F.binary_cross_entropy_with_logits``,
``topk``; pinned only in prose to torch==1.7.1, ``README.md:24``) and in
``scipy.special.ndtri``; their documented semantics are restated here in float64
so this module is the tolerance arbiter for the fp32 / bf16 CUDA paths.

Parity: pinned -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtri  # the reference uses exactly this (cls/custom.py:4,20)

VARIANTS = ("raw", "smooth", "rel", "normit", "gombit", "base2", "base10")


# ----------------------------------------------------------------------------------------
# a4  histograms
# ----------------------------------------------------------------------------------------
def label_hist(labels, num_classes):
    """Per-class counts, cls/imbalanced_dataset.py:112,127 (np.sum(targets == i) for i<C).

    Labels outside [0, C) are counted by nobody, exactly like the reference loop."""
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    ok = (labels >= 0) & (labels < num_classes)
    return np.bincount(labels[ok], minlength=num_classes).astype(np.int64)


def image_dedup_hist(image_ids, categories, num_classes):
    """img_freq / instance_freq of the CSV tables (seg/lvis_files/idf_1204.csv cols 15-16).

    instance_freq[c] = #annotations of class c; img_freq[c] = #distinct images holding
    class c (set de-duplication, cf. seg/mmdet/datasets/dataset_wrappers.py:245-252)."""
    image_ids = np.asarray(image_ids, dtype=np.int64).reshape(-1)
    categories = np.asarray(categories, dtype=np.int64).reshape(-1)
    ok = (categories >= 0) & (categories < num_classes) & (image_ids >= 0)
    inst = np.bincount(categories[ok], minlength=num_classes).astype(np.int64)
    pairs = np.unique(np.stack([image_ids[ok], categories[ok]], 1), axis=0)
    img = np.bincount(pairs[:, 1], minlength=num_classes).astype(np.int64)
    return img, inst


def cifar_lt_profile(img_max, cls_num, imb_factor):
    """cls/imbalanced_dataset.py:23-29 (imb_type='exp')."""
    return [int(img_max * (imb_factor ** (i / (cls_num - 1.0)))) for i in range(cls_num)]


def lt_class_map(counts):
    """Descending-frequency re-index, cls/imbalanced_dataset.py:115-120."""
    order = np.argsort(-np.asarray(counts))
    cmap = np.zeros(len(counts), dtype=np.int64)
    cmap[order] = np.arange(len(counts))
    return cmap


# ----------------------------------------------------------------------------------------
# a2 / a3  IIF weight vectors
# ----------------------------------------------------------------------------------------
def iif_weights_from_counts(freqs, total=None):
    """All seven variants in float64, cls/custom.py:14-23.

    ``total`` defaults to ``freqs.sum()`` (classification); the CSV tables use the number of
    images for the ``*`` columns and ``sum(instance_freq)`` for the ``*_obj`` columns
    (SURVEY.md section 8c).  Division by zero / log of zero give inf/nan as in numpy."""
    f = np.asarray(freqs, dtype=np.float64)
    n = float(f.sum()) if total is None else float(total)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = {
            "raw": np.log(n / f),
            "smooth": np.log((n + 1) / (f + 1)) + 1,
            "rel": np.log((n - f) / f),
            "normit": -ndtri(f / n),
            "gombit": -np.log(-np.log(1 - (f / n))),
            "base2": np.log2(n / f),
            "base10": np.log10(n / f),
        }
    return out


def to_f32_row(v):
    """``torch.tensor([v], dtype=torch.float)``: one rounding f64->f32, shape [1,C] (cls/custom.py:24)."""
    return np.asarray(v, dtype=np.float64).astype(np.float32).reshape(1, -1)


def iif_normalise_f32(v32, p):
    """``v / torch.norm(v, p=iif_norm)`` on the fp32 row (cls/custom.py:25-26)."""
    v = np.asarray(v32, dtype=np.float32)
    nrm = np.float32((np.abs(v.astype(np.float64)) ** p).sum() ** (1.0 / p))
    return (v / nrm).astype(np.float32)


def csv_column_to_weights(column):
    """Drop row 0 (placeholder), append 1.0 for background, round to f32
    (seg/mmdet/models/losses/iif_loss.py:47-50)."""
    vals = list(np.asarray(column, dtype=np.float64))[1:] + [1.0]
    return np.asarray(vals, dtype=np.float64).astype(np.float32).reshape(1, -1)


# ----------------------------------------------------------------------------------------
# a1 / a10  linear forward / backward
# ----------------------------------------------------------------------------------------
def linear_fwd(x, w, b=None):
    """``F.linear``: Z = X W^T + b (cls/resnet_pytorch.py:219,293; bbox_head.py:118)."""
    z = np.asarray(x, np.float64) @ np.asarray(w, np.float64).T
    if b is not None:
        z = z + np.asarray(b, np.float64)[None, :]
    return z


def linear_bwd(dz, x, w):
    """AddmmBackward: dX = dZ W, dW = dZ^T X, db = sum_i dZ_i."""
    dz = np.asarray(dz, np.float64)
    return dz @ np.asarray(w, np.float64), dz.T @ np.asarray(x, np.float64), dz.sum(0)


# ----------------------------------------------------------------------------------------
# a5 / a6 / a7  softmax cross-entropy with the IIF logit scale
# ----------------------------------------------------------------------------------------
def _log_softmax(a):
    m = a.max(axis=1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(a - m)
    lse = np.log(e.sum(axis=1, keepdims=True)) + m
    return a - lse, lse[:, 0]


def softmax_ce(z, iif, label, class_weight=None, sample_weight=None, ignore_index=-100):
    """Per-sample loss l_i and d(sum_i l_i)/dz.

    l_i = -cw[y_i] * log softmax(z_i * iif)[y_i] * w_i, 0 where y_i == ignore_index
    (cls/custom.py:30 with nn.CrossEntropyLoss(reduction='none', weight) :10;
    seg/.../iif_loss.py:187-200; losses/utils.py:42-44).

    Returns (loss_i [B], dz [B,C], lse [B]); the caller applies the reduction scale."""
    z = np.asarray(z, np.float64)
    B, C = z.shape
    s = np.ones((1, C)) if iif is None else np.asarray(iif, np.float64).reshape(1, C)
    y = np.asarray(label, np.int64).reshape(B)
    a = z * s
    logp, lse = _log_softmax(a)
    valid = y != ignore_index
    ys = np.where(valid, y, 0)
    cw = np.ones(C) if class_weight is None else np.asarray(class_weight, np.float64)
    g = np.where(valid, cw[ys], 0.0)
    if sample_weight is not None:
        g = g * np.asarray(sample_weight, np.float64).reshape(B)
    rows = np.arange(B)
    loss_i = -logp[rows, ys] * g
    p = np.exp(logp)
    onehot = np.zeros_like(p)
    onehot[rows, ys] = 1.0
    dz = s * g[:, None] * (p - onehot)
    return loss_i, dz, lse


def mixup_ce(z, iif, label_a, label_b, lam, class_weight=None, sample_weight=None, ignore_index=-100):
    """Mixup.mixup_criterion (cls/custom.py:116-117): lam * criterion(pred, y_a) + (1 - lam) * criterion(pred, y_b),
    i.e. the same linear combination of the per-sample losses and of their gradients.
    Returns (loss_i [B], dz [B,C]); the caller applies the reduction scale."""
    la, da, _ = softmax_ce(z, iif, label_a, class_weight, sample_weight, ignore_index)
    lb, db, _ = softmax_ce(z, iif, label_b, class_weight, sample_weight, ignore_index)
    return lam * la + (1.0 - lam) * lb, lam * da + (1.0 - lam) * db


def reduce_cls(loss_i, reduction):
    """cls/custom.py:32-36: plain mean / sum / per-sample vector. Returns (value, dscale)."""
    B = loss_i.shape[0]
    if reduction == "mean":
        return loss_i.mean() if B else np.float64("nan"), (1.0 / B if B else 0.0)
    if reduction == "sum":
        return loss_i.sum(), 1.0
    return loss_i, 1.0


def reduce_mmdet(loss, reduction="mean", avg_factor=None, loss_weight=1.0):
    """losses/utils.py:28-55 + ``loss_weight *`` (iif_loss.py:142). ``loss`` already holds the
    element weights.  Returns (value, scale multiplying d(sum loss)/dz)."""
    n = loss.size
    if avg_factor is None:
        if reduction == "mean":
            return loss_weight * (loss.mean() if n else np.float64("nan")), (loss_weight / n if n else 0.0)
        if reduction == "sum":
            return loss_weight * loss.sum(), loss_weight
        return loss_weight * loss, loss_weight
    if reduction == "mean":
        return loss_weight * loss.sum() / avg_factor, loss_weight / avg_factor
    if reduction == "none":
        return loss_weight * loss, loss_weight
    raise ValueError('avg_factor can not be used with reduction="sum"')


def softmax_activation(z, iif):
    """``softmax(iif * cls_score, dim=-1)`` (iif_loss.py:76)."""
    z = np.asarray(z, np.float64)
    s = np.ones((1, z.shape[1])) if iif is None else np.asarray(iif, np.float64).reshape(1, -1)
    logp, _ = _log_softmax(z * s)
    return np.exp(logp)


# ----------------------------------------------------------------------------------------
# a8  sigmoid BCE (no IIF scale)
# ----------------------------------------------------------------------------------------
def expand_onehot(label, sample_weight, C, ignore_index=-100):
    """seg/.../cross_entropy_loss.py:53-71."""
    y = np.asarray(label, np.int64).reshape(-1)
    B = y.shape[0]
    valid = (y >= 0) & (y != ignore_index)
    t = np.zeros((B, C))
    pos = valid & (y < C)
    t[np.nonzero(pos)[0], y[pos]] = 1.0
    wrow = valid.astype(np.float64)
    if sample_weight is not None:
        wrow = wrow * np.asarray(sample_weight, np.float64).reshape(B)
    return t, np.repeat(wrow[:, None], C, axis=1)


def bce_with_logits(z, t, pos_weight=None):
    """``F.binary_cross_entropy_with_logits(reduction='none', pos_weight)``:
    (1-t) z + (1 + (pw-1) t) * softplus(-z); gradient wrt z."""
    z = np.asarray(z, np.float64)
    lw = 1.0 if pos_weight is None else 1.0 + (np.asarray(pos_weight, np.float64)[None, :] - 1.0) * t
    sp = np.log1p(np.exp(-np.abs(z))) + np.maximum(-z, 0.0)
    loss = (1.0 - t) * z + lw * sp
    sig = 1.0 / (1.0 + np.exp(-z))
    dz = (1.0 - t) - lw * (1.0 - sig)
    return loss, dz


def sigmoid_bce_mmdet(z, label, sample_weight=None, class_weight=None, ignore_index=-100):
    """Weighted elementwise loss [B,C] and its z-gradient (cross_entropy_loss.py:74-111)."""
    z = np.asarray(z, np.float64)
    t, w = expand_onehot(label, sample_weight, z.shape[1], ignore_index)
    loss, dz = bce_with_logits(z, t, class_weight)
    return loss * w, dz * w


def sigmoid_bce_cls(z, label, weights=None):
    """cls FocalLoss(gamma=0): BCE(z, onehot) * weights_c (cls/custom.py:61-66)."""
    z = np.asarray(z, np.float64)
    B, C = z.shape
    t = np.zeros((B, C))
    t[np.arange(B), np.asarray(label, np.int64)] = 1.0
    loss, dz = bce_with_logits(z, t, None)
    w = 1.0 if weights is None else np.asarray(weights, np.float64)[None, :]
    return loss * w, dz * w


def focal_cls(z, label, gamma, alpha=None, weights=None):
    """cls FocalLoss with gamma > 0 (cls/custom.py:74-84): p = sigmoid(z); BCELoss(p, onehot) * (1 - p_t)^gamma
    * weights_c * alpha_t, with p_t = p t + (1-p)(1-t), alpha_t = alpha t + (1-alpha)(1-t) when alpha is set.
    Exact arithmetic (float64, log-sigmoid via logaddexp); returns the elementwise loss [B,C] and d/dz."""
    z = np.asarray(z, np.float64)
    B, C = z.shape
    t = np.zeros((B, C))
    t[np.arange(B), np.asarray(label, np.int64)] = 1.0
    logp, log1mp = -np.logaddexp(0.0, -z), -np.logaddexp(0.0, z)
    p = np.exp(logp)
    logq = np.where(t > 0, logp, log1mp)                 # log p_t
    q = np.exp(logq)
    mod = (1.0 - q) ** gamma
    loss = -logq * mod
    # d/dq [-log q (1-q)^g] = -(1-q)^g / q + g (1-q)^(g-1) log q ;  dq/dz = +-p(1-p) = +-q(1-q)
    dq = -(1.0 - q) ** (gamma + 1.0) + gamma * q * mod * logq
    dz = np.where(t > 0, dq, -dq)
    w = 1.0 if weights is None else np.asarray(weights, np.float64)[None, :]
    if alpha:
        w = w * (alpha * t + (1.0 - alpha) * (1.0 - t))
    return loss * w, dz * w


# ----------------------------------------------------------------------------------------
# 8f-1  normalised classifiers
# ----------------------------------------------------------------------------------------
def _norm_rows(x, pre, mode, T, p, eps):
    """y_i = pre_i r(n_i) x_i with n_i = |pre_i x_i|; returns (y, a, c) with the backward
    dx_i = a_i g_i + c_i (x_i . g_i) x_i."""
    x = np.asarray(x, np.float64)
    pre = np.ones(x.shape[0]) if pre is None else np.asarray(pre, np.float64).reshape(-1)
    n = np.abs(pre) * np.sqrt((x * x).sum(1))
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == "normed":      # T / (n^p + eps)
            r = T / (n ** p + eps)
            dr = np.where(n > 0, -T * p * n ** (p - 1.0) / (n ** p + eps) ** 2, 0.0)
        elif mode == "cos":       # T / (1 + n)
            r = T / (1.0 + n)
            dr = -T / (1.0 + n) ** 2
        else:                     # T / n
            r = T / np.maximum(n, eps)
            dr = np.where(n > eps, -T / n ** 2, 0.0)
        c = np.where(n > 0, pre ** 3 * dr / n, 0.0)
    a = pre * r
    return a[:, None] * x, a, c


def _norm_rows_bwd(x, g, a, c):
    x = np.asarray(x, np.float64)
    return a[:, None] * g + (c * (x * g).sum(1))[:, None] * x


def normed_linear(x, w, b, gz, temperature=20.0, power=1.0, eps=1e-6, iif=None):
    """mmdet NormedLinear / IIFNormedLinear forward and backward (utils/normed_predictor.py:36-40, 70-76):
    z = F.linear(T x/(|x|^p+eps), w'/(|w'|^p+eps), b), w' = iif_c w_c.  Returns z, dx, dw, db for upstream gz."""
    x_, ax, cx = _norm_rows(x, None, "normed", temperature, power, eps)
    w_, aw, cw = _norm_rows(w, iif, "normed", 1.0, power, eps)
    z = x_ @ w_.T + (0.0 if b is None else np.asarray(b, np.float64)[None, :])
    gz = np.asarray(gz, np.float64)
    dx = _norm_rows_bwd(x, gz @ w_, ax, cx)
    dw = _norm_rows_bwd(w, gz.T @ x_, aw, cw)
    return z, dx, dw, gz.sum(0)


def cosnorm_classifier(x, w, gz, scale=16.0):
    """CosNorm_Classifier (cls/resnet_cifar.py:66-77): z = (scale x/(1+|x|)) (w/|w|)^T.  Returns z, dx, dw."""
    ex, ax, cx = _norm_rows(x, None, "cos", scale, 1.0, 0.0)
    ew, aw, cw = _norm_rows(w, None, "unit", 1.0, 1.0, 0.0)
    z = ex @ ew.T
    gz = np.asarray(gz, np.float64)
    return z, _norm_rows_bwd(x, gz @ ew, ax, cx), _norm_rows_bwd(w, gz.T @ ex, aw, cw)


def cls_normed_linear(x, w_in_out, gz):
    """classification NormedLinear (cls/resnet_cifar.py:38-48): out = F.normalize(x, dim=1) @ F.normalize(W, dim=0),
    W stored [in, out] (F.normalize: v / max(|v|, 1e-12)).  Returns z, dx, dw ([in, out] like the parameter)."""
    wt = np.asarray(w_in_out, np.float64).T                  # [out, in]: columns of W are rows here
    ex, ax, cx = _norm_rows(x, None, "unit", 1.0, 1.0, 1e-12)
    ew, aw, cw = _norm_rows(wt, None, "unit", 1.0, 1.0, 1e-12)
    z = ex @ ew.T
    gz = np.asarray(gz, np.float64)
    return z, _norm_rows_bwd(x, gz @ ew, ax, cx), _norm_rows_bwd(wt, gz.T @ ex, aw, cw).T


def cosnorm_classifier_lr(x, w, gz, scale):
    """CosNorm_Classifier(lr_scale=True) (cls/resnet_cifar.py:56-57,75-76): z = scale^2 (x/(1+|x|)) (w/|w|)^T with a
    learnable scalar.  Returns z, dx, dw, dscale."""
    s2 = float(np.asarray(scale).reshape(-1)[0]) ** 2
    z0, dx0, dw0 = cosnorm_classifier(x, w, np.asarray(gz, np.float64) * s2, scale=1.0)
    dscale = 2.0 * float(np.asarray(scale).reshape(-1)[0]) * float((np.asarray(gz, np.float64) * z0).sum())
    return z0 * s2, dx0, dw0, dscale


# ----------------------------------------------------------------------------------------
# 8f-4  FASA bookkeeping
# ----------------------------------------------------------------------------------------
def class_accumulate(label, loss, num_bins, cum_losses=None, cum_labels=None):
    """fasa_iif_loss.py:154-160: for u in label.unique(): cum_labels[int(u)] += #rows, cum_losses[int(u)] +=
    loss[rows].sum() -- python indexing, so a negative label counts from the end; [B,C] losses sum their rows."""
    y = np.asarray(label, np.int64).reshape(-1)
    l = np.asarray(loss, np.float64).reshape(y.shape[0], -1).sum(1) if y.shape[0] else np.zeros(0)
    cl = np.zeros(num_bins) if cum_losses is None else np.array(cum_losses, np.float64)
    cn = np.zeros(num_bins) if cum_labels is None else np.array(cum_labels, np.float64)
    for u in np.unique(y):
        m = y == u
        cn[int(u)] += m.sum()
        cl[int(u)] += l[m].sum()
    return cl, cn


def class_feature_stats(x, label, mean, var, used, decay):
    """ConvFCFASABBoxHead.fa_update / fa_update_push (fasa_bbox_head.py:118-148); returns the updated copies."""
    x = np.asarray(x, np.float64)
    y = np.asarray(label, np.int64).reshape(-1)
    mean, var, used = np.array(mean, np.float64), np.array(var, np.float64), np.array(used, np.float64)
    for c in np.unique(y):
        e = x[y == c]
        n = e.shape[0]
        m = e.mean(0)
        v = e.var(0)                                         # unbiased=False ...
        if n > 1:
            v = v * n / (n - 1)                              # ... rescaled to the unbiased estimate
        if used[c] > 0:
            mean[c] = decay * m + (1 - decay) * mean[c]
            var[c] = decay * v + (1 - decay) * var[c]
        else:
            mean[c], var[c] = m, v
            used[c] += 1
    return mean, var, used


def bce_dense(z, target, weight=None, pos_weight=None):
    """binary_cross_entropy with already-expanded labels (cross_entropy_loss.py:100-106): elementwise loss * weight
    and its z-gradient."""
    loss, dz = bce_with_logits(z, np.asarray(target, np.float64), pos_weight)
    if weight is not None:
        w = np.asarray(weight, np.float64)
        loss, dz = loss * w, dz * w
    return loss, dz


# ----------------------------------------------------------------------------------------
# a9  accuracy
# ----------------------------------------------------------------------------------------
def label_rank(z, label):
    """rank_i = #{c : z_ic > z_iy} + #{c < y : z_ic == z_iy}; the label is inside the top-k
    iff rank_i < k (lowest-index-first tie order).  Labels outside [0,C) get rank C."""
    z = np.asarray(z)
    B, C = z.shape
    y = np.asarray(label, np.int64).reshape(B)
    ok = (y >= 0) & (y < C)
    ys = np.where(ok, y, 0)
    zy = z[np.arange(B), ys][:, None]
    gt = (z > zy).sum(1)
    eq_before = ((z == zy) & (np.arange(C)[None, :] < ys[:, None])).sum(1)
    return np.where(ok, gt + eq_before, C).astype(np.int64)


def argmax_first(z):
    """First index of the row maximum (torch.argmax / topk(1) on tie-free rows)."""
    return np.asarray(z).argmax(axis=1).astype(np.int64)


def topk_accuracy(z, label, ks=(1,), thresh=None):
    """100 * mean(label in top-k [and its score > thresh]) (cls/utils.py:165-179;
    losses/accuracy.py:41-50); 0 for B=0."""
    z = np.asarray(z)
    B, C = z.shape
    if B == 0:
        return [0.0 for _ in ks]
    r = label_rank(z, label)
    ok = np.ones(B, bool)
    if thresh is not None:
        y = np.clip(np.asarray(label, np.int64).reshape(B), 0, C - 1)
        ok = z[np.arange(B), y] > thresh
    return [100.0 * float(((r < k) & ok).sum()) / B for k in ks]


def shot_accuracy(preds, labels, train_counts, many_shot_thr=100, low_shot_thr=20):
    """many / median / low-shot accuracy (cls/per_shot_acc.py:62-105)."""
    preds = np.asarray(preds)
    labels = np.asarray(labels)
    many, med, low = [], [], []
    for l in np.unique(labels):
        m = labels == l
        acc = (preds[m] == l).sum() / m.sum()
        n = train_counts[int(l)]
        (many if n > many_shot_thr else low if n < low_shot_thr else med).append(acc)
    f = lambda v: float(np.mean(v)) if len(v) else 0.0
    return f(many), f(med), f(low)


# ----------------------------------------------------------------------------------------
# whole head, float64
# ----------------------------------------------------------------------------------------
def head_fwd_bwd(x, w, b, iif, label, *, class_weight=None, sample_weight=None,
                 ignore_index=-100, scale=None):
    """fc_cls -> IIF scale -> softmax-CE -> backward, in float64.

    ``scale`` multiplies d(sum_i l_i); default 1/B (cls 'mean', cls/custom.py:32-33)."""
    z = linear_fwd(x, w, b)
    loss_i, dz, _ = softmax_ce(z, iif, label, class_weight, sample_weight, ignore_index)
    sc = (1.0 / z.shape[0]) if scale is None else scale
    dz = dz * sc
    dx, dw, db = linear_bwd(dz, x, w)
    return dict(z=z, loss_i=loss_i, loss=loss_i.sum() * sc, dz=dz, dx=dx, dw=dw, db=db)
