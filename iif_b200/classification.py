"""Drop-in for classification/custom.py: IIFLoss, FocalLoss, Mixup and the accuracy helper.

Same constructor / forward signatures, attribute names (`.iif`, `.variant`, `.reduction`) and
reductions as the reference; the arithmetic runs in the fused CUDA kernels (no torch ops on the
logits).  Reference lines are cited per method.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_
from . import histogram, ops


class IIFLoss(nn.Module):
    """classification/custom.py:6-39.

    `IIFLoss(dataset, variant='raw', iif_norm=0, reduction='mean', device='cuda', weight=None)`;
    `dataset.get_cls_num_list()` supplies the class counts.  `.iif` is the dict of seven [1,C] fp32
    weight rows (callers test `hasattr(criterion, 'iif')`, train.py:104)."""

    def __init__(self, dataset, variant="raw", iif_norm=0, reduction="mean", device="cuda", weight=None):
        super().__init__()
        self.reduction = reduction
        self.variant = variant
        self.weight = weight  # per-class CE weight (nn.CrossEntropyLoss(weight=...), custom.py:10)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("iif_b200.IIFLoss needs a CUDA device: the weight vector, the loss and its "
                               "gradient are computed by CUDA kernels and there is no CPU fallback")
        counts = torch.as_tensor(list(dataset.get_cls_num_list()), dtype=torch.int64).to(dev)
        self.iif = histogram.iif_weight_dict(counts, iif_norm=float(iif_norm) if iif_norm > 0 else 0.0)

    def forward(self, pred, targets=None, infer=False):
        s = self.iif[self.variant]
        if infer is not False:  # custom.py:37-39: adjusted logits, no softmax
            out, _, _ = ops.scaled_activation(pred.float(), s, softmax=False)
            return out
        B = pred.shape[0]
        if self.reduction == "mean":  # plain mean even with class weights (custom.py:32-33)
            return F_.iif_cross_entropy(pred, s, targets, class_weight=self.weight, scale=1.0 / max(B, 1))
        if self.reduction == "sum":
            return F_.iif_cross_entropy(pred, s, targets, class_weight=self.weight, scale=1.0)
        return F_.iif_cross_entropy(pred, s, targets, class_weight=self.weight, scale=1.0, reduce=False)

    def mixup_forward(self, pred, y_a, y_b, lam):
        """lam * self(pred, y_a) + (1 - lam) * self(pred, y_b)  (custom.py:116-117) from ONE pass over the logits.
        Raises ops.Unsupported when the fused kernel does not cover the shape (C % 4 != 0)."""
        s = self.iif[self.variant]
        B = pred.shape[0]
        kw = dict(class_weight=self.weight, target_b=y_b, lam=float(lam))
        if self.reduction == "mean":
            return F_.iif_cross_entropy(pred, s, y_a, scale=1.0 / max(B, 1), **kw)
        if self.reduction == "sum":
            return F_.iif_cross_entropy(pred, s, y_a, scale=1.0, **kw)
        return F_.iif_cross_entropy(pred, s, y_a, scale=1.0, reduce=False, **kw)


class FocalLoss(nn.Module):
    """classification/custom.py:42-89: gamma == 0 is sigmoid BCE (`--classif bce`), gamma > 0 the focal
    branch with optional alpha balance (:74-89).  The one-hot target tensor of the reference (:61-63) is
    never built; the focal branch is evaluated in logit space (softplus) instead of sigmoid -> log in fp32."""

    def __init__(self, gamma, alpha=None, reduction="mean", device="cuda", weights=None):
        super().__init__()
        if gamma < 0:
            raise ValueError("gamma must be >= 0")
        self.gamma, self.alpha, self.reduction = gamma, alpha, reduction
        self.weights = weights.unsqueeze(0) if weights is not None else 1

    def set_weights(self, weights):
        self.weights = weights.unsqueeze(0)

    def forward(self, pred, targets):
        B, C = pred.shape
        colw = None if isinstance(self.weights, int) else self.weights.reshape(-1)
        # 'sum' -> sum / B ; anything else -> mean over B*C (custom.py:67-70, 84-88)
        scale = 1.0 / max(B, 1) if self.reduction == "sum" else 1.0 / max(B * C, 1)
        return F_.sigmoid_bce(pred, targets, col_weight=colw, scale=scale, gamma=float(self.gamma),
                              alpha=self.alpha if self.alpha else None)


class Mixup(object):
    """classification/custom.py:91-117.  `mixup_criterion` evaluates both terms in ONE fused pass when the
    criterion offers `mixup_forward` (iif_b200 IIFLoss) and the shape qualifies; otherwise the criterion is
    called twice exactly as in the reference."""

    def __init__(self, criterion, alpha=1):
        self.alpha = alpha
        self.criterion = criterion

    def __call__(self, x, y, use_cuda=True):
        import numpy as np
        lam = np.random.beta(self.alpha, self.alpha) if self.alpha > 0 else 1
        index = torch.randperm(x.size()[0], device=x.device if use_cuda else "cpu")
        mixed_x = lam * x + (1 - lam) * x[index, :]
        return mixed_x, y, y[index], lam

    def mixup_criterion(self, pred, y_a, y_b, lam):
        fused = getattr(self.criterion, "mixup_forward", None)
        if fused is not None:
            try:
                return fused(pred, y_a, y_b, lam)
            except ops.Unsupported:
                pass
        return lam * self.criterion(pred, y_a) + (1 - lam) * self.criterion(pred, y_b)


class NormedLinear(nn.Module):
    """classification/resnet_cifar.py:38-48 (`--classif_norm norm`, resnet_pytorch.py:212-219):
    out = F.normalize(x, dim=1) @ F.normalize(weight, dim=0), weight stored [in_features, out_features] and
    initialised uniform(-1,1).renorm_(2,1,1e-5).mul_(1e5); `bias` exists (random) but is NOT used in forward -- both as
    in the reference.  The operand normalisations (F.normalize: v / max(|v|, 1e-12)) and their backward run in the
    library's row kernels, the contraction in the head's GEMMs (`compute` = 'bf16' tensor cores | 'fp32')."""

    def __init__(self, in_features, out_features, compute="bf16", device="cuda"):
        super().__init__()
        self.compute = compute
        self.weight = nn.Parameter(torch.empty(in_features, out_features, device=device))
        self.weight.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)
        self.bias = nn.Parameter(torch.randn(out_features, device=device))

    def forward(self, x):
        from . import _lib
        ex = F_.normalize_rows(x, _lib.NORM_UNIT, temperature=1.0, eps=1e-12)
        # columns of the [in, out] weight = rows of its transpose (the GEMM's [C, D] operand)
        ew = F_.normalize_rows(self.weight.t(), _lib.NORM_UNIT, temperature=1.0, eps=1e-12)
        return F_.linear(ex, ew, None, bf16=(self.compute == "bf16"))


class CosNorm_Classifier(nn.Module):
    """classification/resnet_cifar.py:50-78: z = scale * (x / (1 + |x|)) . (w / |w|)^T  (no bias).  The two
    operand normalisations and their backward run in the library's row kernels, the contraction in the head's
    GEMMs (`compute` = 'bf16' tensor cores | 'fp32').  `lr_scale=True` (`--classif_norm lr_cosine`): the scale is a
    learnable parameter initialised to 5.0 and applied SQUARED (:56-57,75-76)."""

    def __init__(self, in_dims, out_dims, scale=16, margin=0.5, init_std=0.001, lr_scale=False, compute="bf16",
                 device="cuda"):
        super().__init__()
        import math
        self.in_features, self.out_dims, self.lr_scale = in_dims, out_dims, lr_scale
        if lr_scale is True:
            self.scale = nn.Parameter(5.0 * torch.ones(1, device=device))
        else:
            self.scale = scale
        self.margin, self.compute = margin, compute
        self.weight = nn.Parameter(torch.empty(out_dims, in_dims, device=device))
        stdv = 1.0 / math.sqrt(in_dims)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, input, *args):
        from . import _lib
        ew = F_.normalize_rows(self.weight, _lib.NORM_UNIT, temperature=1.0, eps=0.0)
        if self.lr_scale is True:
            # z = scale^2 * (ex . ew): the learnable factor multiplies the product (one scalar, autograd by torch:
            # d/dscale = 2 scale <gz, z0>), the operands stay normalised with T = 1
            ex = F_.normalize_rows(input, _lib.NORM_COS, temperature=1.0)
            return F_.linear(ex, ew, None, bf16=(self.compute == "bf16")) * (self.scale ** 2)
        ex = F_.normalize_rows(input, _lib.NORM_COS, temperature=float(self.scale))
        return F_.linear(ex, ew, None, bf16=(self.compute == "bf16"))


def accuracy(output, target, topk=(1,)):
    """classification/utils.py:165-179: top-k hit rate x 100/B, one fused pass (rank of the label)."""
    with torch.no_grad():
        r = ops.softmax_ce(output.float(), None, target, want_dz_f32=False, want_acc=True, want_sum=False)
        B = target.size(0)
        return [(r["rank"] < k).sum(dtype=torch.float32) * (100.0 / B) for k in topk]


def shot_acc(preds, labels, train_targets, many_shot_thr=100, low_shot_thr=20, acc_per_cls=False):
    """classification/per_shot_acc.py:62-105: mean class accuracy over the many- / median- / low-shot classes (by TRAIN
    count) among the classes present in `labels`.  `preds` / `labels`: CUDA tensors [n]; `train_targets`: the training
    labels (CUDA tensor / array: histogrammed on the device) .  Per-class counts are integer kernels (bit-exact), the
    three means one small reduction; returns python floats like the reference (np.mean), plus the per-class
    accuracies of the present classes (in class order, like np.unique) when `acc_per_cls`."""
    import numpy as np
    dev = preds.device
    tt = torch.as_tensor(np.asarray(train_targets) if not isinstance(train_targets, torch.Tensor) else train_targets)
    tt = tt.to(dev).reshape(-1).long()
    labels = labels.reshape(-1).long()
    n_cls = int(max(int(labels.max()) if labels.numel() else 0, int(tt.max()) if tt.numel() else 0)) + 1
    train_counts = ops.hist_labels(tt, n_cls)
    r = ops.shot_accuracy(preds, labels, train_counts, many_shot_thr, low_shot_thr, want_class_acc=acc_per_cls)
    many, med, low = (float(v) for v in r[0].cpu())
    if acc_per_cls:
        ca = r[3].cpu().numpy()
        return many, med, low, [float(v) for v in ca[ca >= 0]]
    return many, med, low


def predictions(output, iif=None):
    """argmax of the (optionally IIF-adjusted) logits, first index on ties (per_shot_acc.py:130)."""
    _, am, _ = ops.scaled_activation(output.float(), iif, softmax=False, want_pred=True)
    return am.long()
