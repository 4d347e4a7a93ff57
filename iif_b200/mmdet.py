"""Drop-in for the reference's mmdet loss_cls / fc_cls API (instance_segmentation/mmdet).

* `IIFLoss`          <- mmdet/models/losses/iif_loss.py:12-202
* `FasaIIFLoss`      <- mmdet/models/losses/fasa_iif_loss.py:12-208
* `CrossEntropyLoss` <- mmdet/models/losses/cross_entropy_loss.py:165-249 (softmax and sigmoid modes)
* `accuracy`         <- mmdet/models/losses/accuracy.py:6-51
* `Linear`           <- the plain `Linear` entry of LINEAR_LAYERS (mmdet/models/utils/builder.py:14)

Constructor / forward signatures, capability flags, exceptions and reductions follow the reference;
`register_all()` adds the classes to mmdet's LOSSES / LINEAR_LAYERS registries when mmdet is
importable (it is not in this image).
"""
from __future__ import annotations

import pandas as pd
import torch
import torch.nn as nn

from . import functional as F_
from . import ops


def _resolve(reduction, avg_factor, loss_weight, n_elems):
    """losses/utils.py:42-55 folded into one scalar: returns (scale, reduce?).  `n_elems` = numel of the
    element-wise loss (B for softmax, B*C for sigmoid)."""
    if avg_factor is None:
        if reduction == "mean":
            return loss_weight / max(n_elems, 1), True
        if reduction == "sum":
            return loss_weight, True
        return loss_weight, False
    if reduction == "mean":
        return loss_weight / avg_factor, True
    if reduction == "none":
        return loss_weight, False
    raise ValueError('avg_factor can not be used with reduction="sum"')


def _empty_result(pred, reduction, avg_factor):
    """(0,C) input (tests/test_models/test_loss.py:88-101): same values torch gives the reference."""
    z = pred.sum() * 0
    if reduction == "none":
        return pred.new_zeros((0,)) + z
    if reduction == "mean" and avg_factor is None:
        return z + float("nan")
    return z


def accuracy(pred, target, topk=1, thresh=None):
    """mmdet/models/losses/accuracy.py:6-51 (rank of the label instead of a top-k sort)."""
    assert isinstance(topk, (int, tuple))
    if isinstance(topk, int):
        topk, single = (topk,), True
    else:
        single = False
    maxk = max(topk)
    if pred.size(0) == 0:
        accu = [pred.new_tensor(0.) for _ in topk]
        return accu[0] if single else accu
    assert pred.ndim == 2 and target.ndim == 1
    assert pred.size(0) == target.size(0)
    assert maxk <= pred.size(1), f'maxk {maxk} exceeds pred dimension {pred.size(1)}'
    with torch.no_grad():
        r = ops.softmax_ce(pred.float(), None, target, ignore_index=-(2 ** 62), want_dz_f32=False, want_acc=True,
                           want_sum=False)
        rank = r["rank"]
        ok = torch.ones_like(rank, dtype=torch.bool)
        if thresh is not None:
            tgt = target.clamp(0, pred.size(1) - 1)
            ok = pred.gather(1, tgt[:, None])[:, 0] > thresh
        res = [((rank < k) & ok).float().sum(0, keepdim=True).mul_(100.0 / pred.size(0)) for k in topk]
    return res[0] if single else res


class Accuracy(nn.Module):
    def __init__(self, topk=(1,), thresh=None):
        super().__init__()
        self.topk, self.thresh = topk, thresh

    def forward(self, pred, target):
        return accuracy(pred, target, self.topk, self.thresh)


def cross_entropy(pred, label, weight=None, reduction="mean", avg_factor=None, class_weight=None,
                  ignore_index=-100, iif=None, loss_weight=1.0):
    """cross_entropy_loss.py:10-50 / iif_loss.py:157-202 (with `iif`): fused, one kernel."""
    ignore_index = -100 if ignore_index is None else ignore_index
    if pred.size(0) == 0:
        return loss_weight * _empty_result(pred, reduction, avg_factor)
    scale, reduce = _resolve(reduction, avg_factor, loss_weight, pred.size(0))
    return F_.iif_cross_entropy(pred, iif, label, class_weight=class_weight,
                                sample_weight=None if weight is None else weight.float(),
                                ignore_index=ignore_index, scale=scale, reduce=reduce)


def binary_cross_entropy(pred, label, weight=None, reduction="mean", avg_factor=None, class_weight=None,
                         ignore_index=-100, loss_weight=1.0):
    """cross_entropy_loss.py:74-111: class-index labels [B] (pred.dim() != label.dim(): the one-hot expansion of
    :53-71 happens inside the kernel) or already-expanded labels [B,C] (:100-103 skipped, `weight` element-wise)."""
    ignore_index = -100 if ignore_index is None else ignore_index
    if pred.size(0) == 0:
        return loss_weight * _empty_result(pred, reduction, avg_factor)
    scale, reduce = _resolve(reduction, avg_factor, loss_weight, pred.numel())
    if pred.dim() == label.dim():
        return F_.sigmoid_bce_dense(pred, label, pos_weight=class_weight, weight=weight, scale=scale, reduce=reduce)
    return F_.sigmoid_bce(pred, label, pos_weight=class_weight,
                          sample_weight=None if weight is None else weight.float(), ignore_index=ignore_index,
                          scale=scale, reduce=reduce)


class CrossEntropyLoss(nn.Module):
    """mmdet/models/losses/cross_entropy_loss.py:165-249 (use_mask is out of scope: mask head)."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None, ignore_index=None,
                 loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if use_mask:
            raise NotImplementedError("mask_cross_entropy belongs to the mask head, outside the IIF classifier path")
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction, self.loss_weight = reduction, loss_weight
        self.class_weight, self.ignore_index = class_weight, ignore_index
        self.cls_criterion = binary_cross_entropy if use_sigmoid else cross_entropy

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if ignore_index is None:
            ignore_index = self.ignore_index
        class_weight = None
        if self.class_weight is not None:
            class_weight = cls_score.new_tensor(self.class_weight, device=cls_score.device)
        return self.cls_criterion(cls_score, label, weight, class_weight=class_weight, reduction=reduction,
                                  avg_factor=avg_factor, ignore_index=ignore_index, loss_weight=self.loss_weight,
                                  **kwargs)


def _load_iif_csv(path, variant, device="cuda"):
    """iif_loss.py:47-50: CSV column, drop row 0, append 1.0 for background, fp32 [1,C+1].
    An unknown variant raises pandas' KeyError like the reference."""
    vals = pd.read_csv(path)[variant].values.tolist()
    vals = vals[1:] + [1.0]
    return torch.tensor(vals, device=device, dtype=torch.float).unsqueeze(0)


class IIFLoss(nn.Module):
    """mmdet/models/losses/iif_loss.py:12-202."""

    def __init__(self, use_sigmoid=False, reduction='mean', class_weight=None, ignore_index=None, loss_weight=1.0,
                 num_classes=1203, path='./lvis_files/idf_1204.csv', variant='raw'):
        super().__init__()
        assert (use_sigmoid is False)
        self.use_sigmoid = use_sigmoid
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = class_weight
        self.ignore_index = ignore_index
        self.num_classes = num_classes
        self.iif_weights = _load_iif_csv(path, variant)
        self.cls_criterion = self.cross_entropy
        self.custom_cls_channels = True
        self.custom_activation = True
        self.custom_accuracy = True

    def get_activation(self, cls_score):
        """softmax(iif * cls_score) (:65-78), one fused pass."""
        out, _, _ = ops.scaled_activation(cls_score.float(), self.iif_weights, softmax=True)
        return out

    def get_cls_channels(self, num_classes):
        assert num_classes == self.num_classes
        return num_classes + 1

    def get_accuracy(self, cls_score, labels):
        """Accuracy on the RAW scores (:92-107)."""
        return dict(acc_classes=accuracy(cls_score, labels))

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if ignore_index is None:
            ignore_index = self.ignore_index
        class_weight = None
        if self.class_weight is not None:
            class_weight = cls_score.new_tensor(self.class_weight, device=cls_score.device)
        return self.cls_criterion(cls_score, label, weight, class_weight=class_weight, reduction=reduction,
                                  avg_factor=avg_factor, ignore_index=ignore_index, **kwargs)

    def cross_entropy(self, pred, label, weight=None, reduction='mean', avg_factor=None, class_weight=None,
                      ignore_index=-100):
        return cross_entropy(pred, label, weight, reduction, avg_factor, class_weight, ignore_index,
                             iif=self.iif_weights, loss_weight=self.loss_weight)


class FasaIIFLoss(nn.Module):
    """mmdet/models/losses/fasa_iif_loss.py:12-208: softmax (IIF applied) / sigmoid (NO IIF, :35-36)
    and the per-class running sums `cum_losses` / `cum_labels` (:60-71,154-160)."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, loss_weight=1.0,
                 use_cums=False, num_classes=1203, path='./lvis_files/idf_1204.csv', variant='raw'):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if use_mask:
            raise NotImplementedError("mask_cross_entropy belongs to the mask head, outside the IIF classifier path")
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction, self.loss_weight, self.class_weight = reduction, loss_weight, class_weight
        self.cls_criterion = self._sigmoid if use_sigmoid else self.cross_entropy
        self.num_classes = num_classes
        self.use_cums = use_cums
        if self.use_cums:
            self.open_cums()
        self.iif_weights = _load_iif_csv(path, variant)
        self.custom_cls_channels = True
        self.custom_activation = True
        self.custom_accuracy = True

    def open_cums(self):
        self.use_cums = True
        self.reduction_old = self.reduction
        self.reduction = 'none'
        self.cum_losses = torch.zeros(self.num_classes + 1).cuda()
        self.cum_labels = torch.zeros(self.num_classes + 1).cuda()

    def close_cums(self):
        self.use_cums = False
        self.reduction = self.reduction_old
        self.cum_losses = torch.zeros(self.num_classes + 1).cuda()
        self.cum_labels = torch.zeros(self.num_classes + 1).cuda()

    get_activation = IIFLoss.get_activation
    get_cls_channels = IIFLoss.get_cls_channels
    get_accuracy = IIFLoss.get_accuracy

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        class_weight = None
        if self.class_weight is not None:
            class_weight = cls_score.new_tensor(self.class_weight, device=cls_score.device)
        loss_cls = self.cls_criterion(cls_score, label, weight, class_weight=class_weight, reduction=reduction,
                                      avg_factor=avg_factor, **kwargs)
        if self.use_cums:
            # :154-160 without the python loop over label.unique() and its .item() syncs: one segmented-sum kernel
            # (a [B,C] sigmoid loss contributes its row sums, negative labels index from the end like the
            # reference's cum[int(u_l)])
            with torch.no_grad():
                ops.class_accumulate(label, loss_cls.detach().float(), self.cum_losses, self.cum_labels)
            loss_cls = loss_cls.mean()
        return loss_cls

    def cross_entropy(self, pred, label, weight=None, reduction='mean', avg_factor=None, class_weight=None,
                      ignore_index=-100):
        return cross_entropy(pred, label, weight, reduction, avg_factor, class_weight, ignore_index,
                             iif=self.iif_weights, loss_weight=self.loss_weight)

    def _sigmoid(self, pred, label, weight=None, reduction='mean', avg_factor=None, class_weight=None,
                 ignore_index=-100):
        return binary_cross_entropy(pred, label, weight, reduction, avg_factor, class_weight, ignore_index,
                                    loss_weight=self.loss_weight)


class Linear(nn.Linear):
    """fc_cls (`LINEAR_LAYERS['Linear']`, utils/builder.py:14; built at bbox_head.py:66-75): an
    nn.Linear subclass -- same parameter names / init -- whose forward and backward GEMMs are the
    head's own kernels.  compute='bf16' uses the tcgen05 path with a cached bf16 copy of the
    weight (refreshed when the fp32 master weight changes); 'fp32' is the 1e-5 parity mode."""

    def __init__(self, *args, compute="bf16", **kwargs):
        super().__init__(*args, **kwargs)
        assert compute in ("bf16", "fp32", "fp32x3")
        self.compute = compute
        self._w16 = None
        self._w16_key = None

    def weight_bf16(self):
        """bf16 operand copy of the fp32 master weight.  Refreshed on EVERY training forward (one small cast kernel):
        optimizers, EMA hooks and fp16 master-copy hooks write through `.data`, which does not bump `_version`.  In
        eval mode the copy is kept while (version, storage) are unchanged; `invalidate()` drops it explicitly."""
        key = (self.weight._version, self.weight.data_ptr())
        if self.training or self._w16 is None or self._w16_key != key:
            with torch.no_grad():
                self._w16 = ops.scale_rows(self.weight.detach().float(), None, bf16=True)
            self._w16_key = key
        return self._w16

    def invalidate(self):
        self._w16 = None

    def forward(self, x):
        if self.compute == "bf16":
            return F_.linear(x, self.weight, self.bias, bf16=True, weight_bf16=self.weight_bf16())
        return F_.linear(x, self.weight, self.bias, bf16="x3" if self.compute == "fp32x3" else False)


class NormedLinear(Linear):
    """`LINEAR_LAYERS['NormedLinear']` (utils/normed_predictor.py:11-40):
    z = F.linear(T x / (|x|^p + eps), w / (|w|^p + eps), bias) -- the operand normalisations and their
    backward run in the library's row kernels, the contraction in the head's GEMMs.  Same constructor
    (including the reference's spelling `tempearture`), same parameter names / init."""

    def __init__(self, *args, tempearture=20, power=1.0, eps=1e-6, compute="bf16", **kwargs):
        super().__init__(*args, compute=compute, **kwargs)
        self.tempearture, self.power, self.eps = tempearture, power, eps
        self.init_weights()

    def init_weights(self):
        nn.init.normal_(self.weight, mean=0, std=0.01)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    def _class_scale(self):
        return None

    def forward(self, x):
        from . import _lib
        w_ = F_.normalize_rows(self.weight, _lib.NORM_NORMED, pre=self._class_scale(), temperature=1.0,
                               power=self.power, eps=self.eps)
        x_ = F_.normalize_rows(x, _lib.NORM_NORMED, temperature=self.tempearture, power=self.power, eps=self.eps)
        return F_.linear(x_, w_, self.bias, bf16=(self.compute == "bf16"))


class IIFNormedLinear(NormedLinear):
    """`LINEAR_LAYERS['IIFNormedLinear']` (utils/normed_predictor.py:43-76): NormedLinear on the class rows
    pre-multiplied by their IIF weight, w' = iif_c w_c (CSV column `variant`, background entry 1.0)."""

    def __init__(self, *args, tempearture=20, power=1.0, eps=1e-6, path="./lvis_files/idf_1204.csv",
                 variant="base2_obj", compute="bf16", device="cuda", **kwargs):
        super().__init__(*args, tempearture=tempearture, power=power, eps=eps, compute=compute, **kwargs)
        import pandas as pd
        vals = pd.read_csv(path)[variant].values.tolist()        # KeyError on an unknown column, as the reference
        vals = vals[1:] + [1.0]                                    # drop the placeholder row, +1 for background
        self.iif_weights = torch.tensor(vals, device=device, dtype=torch.float).unsqueeze(1)

    def _class_scale(self):
        return self.iif_weights.reshape(-1)


def sibling_forward(fc_cls, fc_reg, x):
    """`cls_score = self.fc_cls(x); bbox_pred = self.fc_reg(x)` of BBoxHead.forward (bbox_head.py:118-119,
    convfc_bbox_head.py:188-189) as ONE GEMM per direction over the shared RoI features.  `fc_cls` / `fc_reg` stay
    two nn.Linear modules (checkpoint names, init_cfg overrides and `selectp` keep working); a normalised fc_cls
    (NormedLinear / IIFNormedLinear) or the fp32 parity mode falls back to the two separate calls."""
    plain = type(fc_cls) in (Linear, nn.Linear) and type(fc_reg) in (Linear, nn.Linear)
    if not plain or getattr(fc_cls, "compute", "bf16") != "bf16" or getattr(fc_reg, "compute", "bf16") != "bf16" \
            or not x.is_cuda or x.numel() == 0:
        return fc_cls(x), fc_reg(x)
    return F_.sibling_linear(x, fc_cls.weight, fc_cls.bias, fc_reg.weight, fc_reg.bias)


def register_all():
    """Register into mmdet's registries (LOSSES: IIFLoss, FasaIIFLoss; LINEAR_LAYERS: Linear, NormedLinear,
    IIFNormedLinear) when mmdet is importable; returns the names registered."""
    done = []
    try:
        from mmdet.models.builder import LOSSES  # type: ignore
        for cls in (IIFLoss, FasaIIFLoss):
            LOSSES.register_module(name=cls.__name__, force=True, module=cls)
            done.append(cls.__name__)
        from mmdet.models.utils.builder import LINEAR_LAYERS  # type: ignore
        for cls in (Linear, NormedLinear, IIFNormedLinear):
            LINEAR_LAYERS.register_module(name=cls.__name__, force=True, module=cls)
            done.append(cls.__name__)
    except ImportError:
        pass
    return done
