"""Tensor-level wrappers over the C ABI: raw device pointers + the current CUDA stream.

torch is plumbing here (device memory, streams); every arithmetic op of the path runs in the
hand-written kernels of libiif_b200.so.  CPU tensors are rejected loudly -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_WS = {}       # (device index, stream) -> GEMM workspace
_SCRATCH = {}  # (device index, stream) -> zeroed loss scratch


def _cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"iif_b200: `{name}` must be a CUDA tensor (no CPU fallback exists); got "
                           f"{getattr(t, 'device', type(t))}")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"iif_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t


def _rows(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """2-D, unit inner stride (any leading dimension)."""
    _cuda(t, name, dtype)
    if t.dim() != 2:
        raise ValueError(f"iif_b200: `{name}` must be 2-D, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    """Leading dimension in elements.  A single row never steps by it: report an aligned value."""
    return int(t.stride(0)) if t.shape[0] > 1 else (max(int(t.shape[1]), 1) + 7) // 8 * 8


def _vec(t: Optional[torch.Tensor], name: str, n: int, dtype=torch.float32):
    if t is None:
        return None
    _cuda(t, name)
    t = t.reshape(-1)
    if t.numel() != n:
        raise ValueError(f"iif_b200: `{name}` must have {n} elements, got {t.numel()}")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def loss_scratch(device, B: int) -> torch.Tensor:
    """Zeroed scratch of the fused loss kernels (ticket + per-CTA partials), one per (device, stream)."""
    n = int(_lib.load().iif_loss_scratch_bytes(int(B)))
    key = (torch.device(device).index or 0, torch.cuda.current_stream(device).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None or t.numel() * 4 < n:
        t = torch.zeros((n + 3) // 4, dtype=torch.int32, device=device)
        _SCRATCH[key] = t
    return t


def gemm_workspace(B: int, D: int, Cc: int, device) -> Optional[torch.Tensor]:
    n = int(_lib.load().iif_gemm_ws_bytes(B, D, Cc))
    if n == 0:
        return None
    key = (torch.device(device).index or 0, torch.cuda.current_stream(device).cuda_stream)
    t = _WS.get(key)
    if t is None or t.numel() < n:
        t = torch.zeros(n, dtype=torch.uint8, device=device)   # split-K counters start (and stay) zero
        _WS[key] = t
    return t


def launch_count() -> int:
    return int(_lib.load().iif_launch_count())


# ------------------------------------------------------------------------------------------------
# histogram + weights
# ------------------------------------------------------------------------------------------------
def hist_labels(labels: torch.Tensor, num_classes: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    labels = _cuda(labels, "labels", torch.int64).reshape(-1).contiguous()
    if out is None:
        out = torch.zeros(num_classes, dtype=torch.int64, device=labels.device)
    _lib.check(_lib.load().iif_hist_labels_i64(_ptr(labels), labels.numel(), _ptr(out), num_classes,
                                               _stream(labels.device)), "hist_labels")
    return out


def hist_images_dedup(image_ids: torch.Tensor, categories: torch.Tensor, num_images: int, num_classes: int):
    image_ids = _cuda(image_ids, "image_ids", torch.int64).reshape(-1).contiguous()
    categories = _cuda(categories, "categories", torch.int64).reshape(-1).contiguous()
    if image_ids.numel() != categories.numel():
        raise ValueError("image_ids and categories must have the same length")
    dev = image_ids.device
    lib = _lib.load()
    ws = torch.zeros(max(int(lib.iif_hist_images_dedup_ws_bytes(num_images, num_classes)) // 4, 1),
                     dtype=torch.int32, device=dev)
    img = torch.zeros(num_classes, dtype=torch.int64, device=dev)
    inst = torch.zeros(num_classes, dtype=torch.int64, device=dev)
    _lib.check(lib.iif_hist_images_dedup_i64(_ptr(image_ids), _ptr(categories), image_ids.numel(), num_images,
                                             num_classes, _ptr(img), _ptr(inst), _ptr(ws), _stream(dev)),
               "hist_images_dedup")
    return img, inst


def weights_from_counts(counts: torch.Tensor, variant: str, total: int = 0, norm_p: float = 0.0,
                        return_f64: bool = False):
    counts = _cuda(counts, "counts", torch.int64).reshape(-1).contiguous()
    if variant not in _lib.VARIANT_IDS:
        raise KeyError(variant)
    n = counts.numel()
    out = torch.empty(n, dtype=torch.float32, device=counts.device)
    o64 = torch.empty(n, dtype=torch.float64, device=counts.device) if return_f64 else None
    _lib.check(_lib.load().iif_weights_from_counts(_ptr(counts), n, int(total), _lib.VARIANT_IDS[variant],
                                                   float(norm_p), _ptr(out), _ptr(o64), _stream(counts.device)),
               "weights_from_counts")
    return (out, o64) if return_f64 else out


# ------------------------------------------------------------------------------------------------
# loss kernels
# ------------------------------------------------------------------------------------------------
def pad64(n: int) -> int:
    """Row pitch (elements) of a bf16 matrix whose rows start on 128-byte boundaries: TMA boxes are 128 bytes wide
    and a row that straddles two 128-byte lines costs two L2 requests per box row."""
    return (int(n) + 63) // 64 * 64


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class Unsupported(RuntimeError):
    """The kernel does not cover this shape / alignment (IIF_EUNSUPPORTED): the caller picks another path."""


def softmax_ce(z, iif, label, *, class_weight=None, sample_weight=None, ignore_index=-100, scale=1.0,
               want_dz_f32=True, want_dz_bf16=False, want_acc=False, want_sum=True, want_lse=False,
               label_b=None, lam=1.0):
    """Returns dict(loss_i, loss_sum, dz_f32, dz_bf16 [B,pad8(C)], argmax, rank, acc_counts, lse).
    With `label_b`: the dual-label Mixup loss lam*CE(label) + (1-lam)*CE(label_b) in one pass
    (raises `Unsupported` when the 128-bit path does not apply)."""
    z = _rows(z, "z", torch.float32)
    B, Cc = z.shape
    dev = z.device
    label = _vec(label, "label", B, torch.int64)
    iif = _vec(iif, "iif", Cc)
    cw = _vec(class_weight, "class_weight", Cc)
    sw = _vec(sample_weight, "sample_weight", B)
    r = dict(loss_i=torch.empty(B, dtype=torch.float32, device=dev))
    r["loss_sum"] = torch.zeros((), dtype=torch.float32, device=dev) if want_sum else None
    r["dz_f32"] = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_dz_f32 else None
    r["dz_bf16"] = torch.empty(B, pad8(Cc), dtype=torch.bfloat16, device=dev) if want_dz_bf16 else None
    r["argmax"] = torch.empty(B, dtype=torch.int32, device=dev) if want_acc else None
    r["rank"] = torch.empty(B, dtype=torch.int32, device=dev) if want_acc else None
    r["acc_counts"] = torch.zeros(2, dtype=torch.int32, device=dev) if want_acc else None
    r["lse"] = torch.empty(B, dtype=torch.float32, device=dev) if want_lse else None
    if B == 0:
        return r
    tk = loss_scratch(dev, B)
    if label_b is not None:
        if want_lse:
            raise ValueError("want_lse is not available with label_b")
        label_b = _vec(label_b, "label_b", B, torch.int64)
        rc = _lib.load().iif_softmax_ce_mixup_fwd_bwd(
            _ptr(z), _ld(z), _ptr(iif), _ptr(label), _ptr(label_b), float(lam), _ptr(cw), _ptr(sw), int(ignore_index),
            float(scale), B, Cc, _ptr(r["loss_i"]), _ptr(r["loss_sum"]), _ptr(r["dz_f32"]), Cc, _ptr(r["dz_bf16"]),
            pad8(Cc), _ptr(r["argmax"]), _ptr(r["rank"]), _ptr(r["acc_counts"]), _ptr(tk), _stream(dev))
        if rc == _lib.EUNSUPPORTED:
            raise Unsupported("softmax_ce_mixup: needs C % 4 == 0 and 16-byte aligned rows")
        _lib.check(rc, "softmax_ce_mixup_fwd_bwd")
        return r
    _lib.check(_lib.load().iif_softmax_ce_fwd_bwd(
        _ptr(z), _ld(z), _ptr(iif), _ptr(label), _ptr(cw), _ptr(sw), int(ignore_index), float(scale), B, Cc,
        _ptr(r["loss_i"]), _ptr(r["loss_sum"]), _ptr(r["dz_f32"]), Cc, _ptr(r["dz_bf16"]), pad8(Cc), _ptr(r["lse"]),
        _ptr(r["argmax"]), _ptr(r["rank"]), _ptr(r["acc_counts"]), _ptr(tk), _stream(dev)), "softmax_ce_fwd_bwd")
    return r


def scaled_activation(z, iif, softmax: bool, label=None, want_pred=False):
    z = _rows(z, "z", torch.float32)
    B, Cc = z.shape
    dev = z.device
    iif = _vec(iif, "iif", Cc)
    out = torch.empty(B, Cc, dtype=torch.float32, device=dev)
    lab = _vec(label, "label", B, torch.int64) if label is not None else None
    am = torch.empty(B, dtype=torch.int32, device=dev) if want_pred else None
    rk = torch.empty(B, dtype=torch.int32, device=dev) if (want_pred and lab is not None) else None
    if B:
        _lib.check(_lib.load().iif_scaled_activation(_ptr(z), _ld(z), _ptr(iif), int(bool(softmax)), B, Cc, _ptr(out),
                                                     Cc, _ptr(lab), _ptr(am), _ptr(rk), _stream(dev)),
                   "scaled_activation")
    return out, am, rk


def sigmoid_bce(z, label, *, pos_weight=None, col_weight=None, sample_weight=None, ignore_index=-100, scale=1.0,
                want_elem=False, want_dz_f32=True, want_dz_bf16=False, want_sum=True, gamma=0.0, alpha=None):
    """Sigmoid BCE forward + backward; gamma > 0 selects the focal form (cls/custom.py:74-89)."""
    z = _rows(z, "z", torch.float32)
    B, Cc = z.shape
    dev = z.device
    label = _vec(label, "label", B, torch.int64)
    pw = _vec(pos_weight, "pos_weight", Cc)
    colw = _vec(col_weight, "col_weight", Cc)
    sw = _vec(sample_weight, "sample_weight", B)
    r = dict(loss_i=torch.empty(B, dtype=torch.float32, device=dev))
    r["loss_sum"] = torch.zeros((), dtype=torch.float32, device=dev) if want_sum else None
    r["loss_elem"] = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_elem else None
    r["dz_f32"] = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_dz_f32 else None
    r["dz_bf16"] = torch.empty(B, pad8(Cc), dtype=torch.bfloat16, device=dev) if want_dz_bf16 else None
    if B == 0:
        return r
    if gamma and gamma > 0:
        if pw is not None:
            raise ValueError("pos_weight is not part of the focal form")
        _lib.check(_lib.load().iif_sigmoid_focal_fwd_bwd(
            _ptr(z), _ld(z), _ptr(label), float(gamma), float(alpha) if alpha else 0.0, _ptr(colw), _ptr(sw),
            int(ignore_index), float(scale), B, Cc, _ptr(r["loss_elem"]), Cc, _ptr(r["loss_i"]), _ptr(r["loss_sum"]),
            _ptr(r["dz_f32"]), Cc, _ptr(r["dz_bf16"]), pad8(Cc), _ptr(loss_scratch(dev, B)), _stream(dev)),
            "sigmoid_focal_fwd_bwd")
        return r
    _lib.check(_lib.load().iif_sigmoid_bce_fwd_bwd(
        _ptr(z), _ld(z), _ptr(label), _ptr(pw), _ptr(colw), _ptr(sw), int(ignore_index), float(scale), B, Cc,
        _ptr(r["loss_elem"]), Cc, _ptr(r["loss_i"]), _ptr(r["loss_sum"]), _ptr(r["dz_f32"]), Cc, _ptr(r["dz_bf16"]),
        pad8(Cc), _ptr(loss_scratch(dev, B)), _stream(dev)), "sigmoid_bce_fwd_bwd")
    return r


def sigmoid_bce_dense(z, target, *, pos_weight=None, weight=None, scale=1.0, want_elem=False, want_dz=True,
                      want_sum=True):
    """Sigmoid BCE with already-expanded (dense / soft) targets [B,C]; `weight`: None, [B], [B,1] or [B,C]."""
    z = _rows(z, "z", torch.float32)
    B, Cc = z.shape
    dev = z.device
    target = _rows(target, "target", torch.float32)
    if tuple(target.shape) != (B, Cc):
        raise ValueError(f"iif_b200: dense BCE target must be [{B},{Cc}], got {tuple(target.shape)}")
    pw = _vec(pos_weight, "pos_weight", Cc)
    ldw = 0
    if weight is not None:
        _cuda(weight, "weight")
        weight = weight.float()
        if weight.numel() == B:
            weight = weight.reshape(-1).contiguous()
        elif tuple(weight.shape) == (B, Cc):
            weight = _rows(weight, "weight", torch.float32)
            ldw = _ld(weight)
        else:
            raise ValueError(f"iif_b200: dense BCE weight must have {B} or {B}x{Cc} elements")
    r = dict(loss_i=torch.empty(B, dtype=torch.float32, device=dev))
    r["loss_sum"] = torch.zeros((), dtype=torch.float32, device=dev) if want_sum else None
    r["loss_elem"] = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_elem else None
    r["dz_f32"] = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_dz else None
    if B == 0:
        return r
    _lib.check(_lib.load().iif_sigmoid_bce_dense_fwd_bwd(
        _ptr(z), _ld(z), _ptr(target), _ld(target), _ptr(pw), _ptr(weight), ldw, float(scale), B, Cc,
        _ptr(r["loss_elem"]), Cc, _ptr(r["loss_i"]), _ptr(r["loss_sum"]), _ptr(r["dz_f32"]), Cc,
        _ptr(loss_scratch(dev, B)), _stream(dev)), "sigmoid_bce_dense_fwd_bwd")
    return r


def class_accumulate(label, loss, cum_losses, cum_labels):
    """cum_labels[c] += #{label == c}; cum_losses[c] += sum of loss rows with label c (fasa_iif_loss.py:154-160).
    In place on the two fp32 accumulators [num_bins]; `loss` is [B] or [B,C] (rows are summed)."""
    label = _vec(label, "label", label.numel(), torch.int64)
    B = label.numel()
    _cuda(loss, "loss", torch.float32)
    if loss.dim() == 1:
        loss2 = loss.contiguous().reshape(B, 1) if B else loss.reshape(0, 1)
    else:
        loss2 = _rows(loss.reshape(B, -1), "loss", torch.float32)
    _cuda(cum_losses, "cum_losses", torch.float32)
    _cuda(cum_labels, "cum_labels", torch.float32)
    nb = cum_losses.numel()
    if cum_labels.numel() != nb or not cum_losses.is_contiguous() or not cum_labels.is_contiguous():
        raise ValueError("iif_b200: cum_losses / cum_labels must be contiguous and of equal length")
    if B:
        _lib.check(_lib.load().iif_class_accumulate(_ptr(label), _ptr(loss2), _ld(loss2) if loss2.shape[1] > 1 else 1,
                                                    int(loss2.shape[1]), B, nb, _ptr(cum_losses), _ptr(cum_labels),
                                                    _stream(label.device)), "class_accumulate")


_STATS_WS = {}


def class_feature_stats(x, label, feature_mean, feature_var, feature_used, decay):
    """FasaBBoxHead.fa_update (fasa_bbox_head.py:118-148) for all classes present in `label`, in place on the running
    statistics feature_mean / feature_var [num_bins, D] and feature_used [num_bins] (fp32)."""
    x = _rows(x, "x", torch.float32)
    B, D = x.shape
    label = _vec(label, "label", B, torch.int64)
    for t, n in ((feature_mean, "feature_mean"), (feature_var, "feature_var")):
        _cuda(t, n, torch.float32)
        if t.dim() != 2 or t.shape[1] != D or t.stride(1) != 1:
            raise ValueError(f"iif_b200: {n} must be [num_bins,{D}] with unit inner stride")
    nb = feature_mean.shape[0]
    _cuda(feature_used, "feature_used", torch.float32)
    if feature_var.shape[0] != nb or feature_used.numel() != nb or feature_mean.stride(0) != feature_var.stride(0):
        raise ValueError("iif_b200: feature_mean / feature_var / feature_used disagree")
    key = (x.device.index, nb)
    ws = _STATS_WS.get(key)
    if ws is None:
        ws = _STATS_WS[key] = torch.zeros(nb, dtype=torch.int32, device=x.device)
    if B:
        _lib.check(_lib.load().iif_class_feature_stats(_ptr(x), _ld(x), _ptr(label), B, D, nb, float(decay),
                                                       _ptr(feature_mean), _ptr(feature_var), int(feature_mean.stride(0)),
                                                       _ptr(feature_used), _ptr(ws), _stream(x.device)),
                   "class_feature_stats")


def shot_accuracy(preds, labels, train_counts, many_shot_thr=100, low_shot_thr=20, want_class_acc=False):
    """many / median / low-shot accuracy (cls/per_shot_acc.py:62-105).  preds int32/int64 [n], labels int64 [n],
    train_counts int64 [C] (per-class TRAIN counts).  Returns (out3 float64 [3] on the device, test_counts,
    correct_counts[, class_acc float64 [C], -1 where the class is absent from `labels`])."""
    _cuda(preds, "preds")
    n = preds.numel()
    preds = preds.reshape(-1).to(torch.int32).contiguous()
    labels = _vec(labels, "labels", n, torch.int64)
    train_counts = _vec(train_counts, "train_counts", train_counts.numel(), torch.int64)
    Cc = train_counts.numel()
    dev = preds.device
    test = torch.empty(Cc, dtype=torch.int64, device=dev)
    correct = torch.empty(Cc, dtype=torch.int64, device=dev)
    out3 = torch.empty(3, dtype=torch.float64, device=dev)
    cacc = torch.empty(Cc, dtype=torch.float64, device=dev) if want_class_acc else None
    _lib.check(_lib.load().iif_shot_accuracy(_ptr(preds), _ptr(labels), n, _ptr(train_counts), Cc, int(many_shot_thr),
                                             int(low_shot_thr), _ptr(test), _ptr(correct), _ptr(out3), _ptr(cacc),
                                             _stream(dev)), "shot_accuracy")
    return (out3, test, correct, cacc) if want_class_acc else (out3, test, correct)


def scale_rows(x, g=None, *, bf16=False, pad_ld=False, out=None):
    """x[rows, cols] fp32 * g (None | 0-dim device scalar | [rows]) -> fp32 or bf16 (optionally ld padded to 8).
    `out`: write into this [rows, cols] view (unit inner stride, any leading dimension) instead of allocating."""
    x = _rows(x, "x", torch.float32)
    rows, cols = x.shape
    dev = x.device
    if out is not None:
        _cuda(out, "out", torch.bfloat16 if bf16 else torch.float32)
        if tuple(out.shape) != (rows, cols) or (cols > 1 and out.stride(1) != 1):
            raise ValueError(f"iif_b200: scale_rows out must be [{rows},{cols}] with unit inner stride")
    gs = 0
    if g is not None:
        g = _cuda(g, "g")
        g = g.to(torch.float32).reshape(-1).contiguous()
        if g.numel() not in (1, rows):
            raise ValueError("g must be a scalar or have one entry per row")
        gs = 0 if g.numel() == 1 and rows != 1 else (1 if g.numel() == rows and rows > 1 else 0)
    if out is not None:
        ldo, ret = _ld(out), out
    else:
        ldo = pad8(cols) if (bf16 and pad_ld) else cols
        out = torch.empty(rows, ldo, dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
        ret = out if ldo == cols else out[:, :cols]
    if rows and cols:
        _lib.check(_lib.load().iif_scale_rows(_ptr(x), _ld(x), _ptr(g), gs, rows, cols, _ptr(out),
                                              _lib.DTYPE_BF16 if bf16 else _lib.DTYPE_F32, ldo, _stream(dev)),
                   "scale_rows")
    return ret


def row_scale_from_norm(x, mode, *, pre=None, temperature=1.0, power=1.0, eps=1e-6, want_c=True):
    """(a, c): per-row operand multiplier a_i = pre_i r(|pre_i x_i|) and backward coefficient c_i (see the header)."""
    x = _rows(x, "x", torch.float32)
    rows, cols = x.shape
    dev = x.device
    pre = _vec(pre, "pre", rows)
    a = torch.empty(rows, dtype=torch.float32, device=dev)
    c = torch.empty(rows, dtype=torch.float32, device=dev) if want_c else None
    if rows:
        _lib.check(_lib.load().iif_row_scale_from_norm(_ptr(x), _ld(x), rows, cols, _ptr(pre), int(mode), float(temperature),
                                                       float(power), float(eps), _ptr(a), _ptr(c), None, _stream(dev)),
                   "row_scale_from_norm")
    return a, c


def row_dot(u, v):
    u = _rows(u, "u", torch.float32)
    v = _rows(v, "v", torch.float32)
    if u.shape != v.shape:
        raise ValueError("row_dot: shape mismatch")
    out = torch.empty(u.shape[0], dtype=torch.float32, device=u.device)
    if u.shape[0]:
        _lib.check(_lib.load().iif_row_dot(_ptr(u), _ld(u), _ptr(v), _ld(v), u.shape[0], u.shape[1], _ptr(out),
                                           _stream(u.device)), "row_dot")
    return out


def rows_axpby(u, a=None, v=None, b=None, b2=None):
    """out_i = a_i u_i + (b_i b2_i) v_i  (fp32)."""
    u = _rows(u, "u", torch.float32)
    rows, cols = u.shape
    dev = u.device
    if v is not None:
        v = _rows(v, "v", torch.float32)
        if v.shape != u.shape:
            raise ValueError("rows_axpby: shape mismatch")
    a, b, b2 = _vec(a, "a", rows), _vec(b, "b", rows), _vec(b2, "b2", rows)
    out = torch.empty(rows, cols, dtype=torch.float32, device=dev)
    if rows and cols:
        _lib.check(_lib.load().iif_rows_axpby(_ptr(u), _ld(u), _ptr(a), _ptr(v), 0 if v is None else _ld(v), _ptr(b),
                                              _ptr(b2), rows, cols, _ptr(out), cols, _stream(dev)), "rows_axpby")
    return out


def scale_inplace_(t, g):
    """t *= g (0-dim / 1-element fp32 device tensor) in place; a no-op launch when g == 1."""
    _cuda(t, "t")
    _cuda(g, "g", torch.float32)
    if t.dtype not in (torch.float32, torch.bfloat16) or not t.is_contiguous():
        raise TypeError("iif_b200: scale_inplace_ takes a contiguous fp32 / bf16 tensor")
    _lib.check(_lib.load().iif_scale_inplace(_ptr(t), _lib.DTYPE_BF16 if t.dtype == torch.bfloat16 else _lib.DTYPE_F32,
                                             t.numel(), _ptr(g.reshape(1)), _stream(t.device)), "scale_inplace")
    return t


def split3(x, *, k_along_rows: bool, side_b: bool):
    """fp32 [rows, cols] -> the six-copy bf16 expansion along the contraction dimension (csrc/split3.cu):
    [rows, 6 * pad8(cols)] (K along columns) or [6 * pad8(rows), pad8-pitched cols] (K along rows)."""
    x = _rows(x, "x", torch.float32)
    rows, cols = x.shape
    if k_along_rows:
        out = torch.empty(6 * pad8(rows), pad8(cols), dtype=torch.bfloat16, device=x.device)[:, :cols]
    else:
        out = torch.empty(rows, 6 * pad8(cols), dtype=torch.bfloat16, device=x.device)
    if rows and cols:
        _lib.check(_lib.load().iif_split3_bf16(_ptr(x), _ld(x), rows, cols, int(bool(k_along_rows)), int(bool(side_b)),
                                               _ptr(out), int(out.stride(0)), _stream(x.device)), "split3_bf16")
    return out


def colsum(dz, alpha=None):
    _cuda(dz, "dz")
    rows, cols = dz.shape
    dt = _lib.DTYPE_BF16 if dz.dtype == torch.bfloat16 else _lib.DTYPE_F32
    if dz.dtype not in (torch.bfloat16, torch.float32) or (cols > 1 and dz.stride(1) != 1):
        raise TypeError("colsum expects fp32/bf16 rows with unit inner stride")
    db = torch.empty(cols, dtype=torch.float32, device=dz.device)
    al = None if alpha is None else _cuda(alpha, "alpha").to(torch.float32).reshape(-1)[:1].contiguous()
    _lib.check(_lib.load().iif_colsum(_ptr(dz), dt, _ld(dz), _ptr(al), rows, cols, _ptr(db), _stream(dz.device)),
               "colsum")
    return db


# ------------------------------------------------------------------------------------------------
# GEMMs
# ------------------------------------------------------------------------------------------------
def _alpha(alpha, dev):
    if alpha is None:
        return None
    return _cuda(alpha, "alpha").to(torch.float32).reshape(-1)[:1].contiguous()


def _bf16_rows(t, name):
    t = _rows(t, name, torch.bfloat16)
    if _ld(t) % 8 or t.data_ptr() % 16:
        c = t.shape[1]
        buf = torch.empty(t.shape[0], pad8(c), dtype=torch.bfloat16, device=t.device)
        buf[:, :c].copy_(t)
        t = buf[:, :c]
    return t


def linear_fwd(x, w, bias=None, col_scale=None, *, want_raw=True, want_scaled=False):
    """Z = X W^T + b (and optionally Z * col_scale).  dtype of x/w selects the kernel: bf16 -> tcgen05, fp32 -> FFMA."""
    bf = x.dtype == torch.bfloat16
    if w.dtype != x.dtype:
        raise TypeError(f"x ({x.dtype}) and w ({w.dtype}) must share a dtype")
    x = _bf16_rows(x, "x") if bf else _rows(x, "x", torch.float32)
    w = _bf16_rows(w, "w") if bf else _rows(w, "w", torch.float32)
    B, D = x.shape
    Cc, D2 = w.shape
    if D != D2:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)} vs w {tuple(w.shape)}")
    dev = x.device
    bias = _vec(bias, "bias", Cc)
    cs = _vec(col_scale, "col_scale", Cc)
    z = torch.empty(B, Cc, dtype=torch.float32, device=dev) if want_raw else None
    zs = torch.empty(B, Cc, dtype=torch.float32, device=dev) if (want_scaled and cs is not None) else None
    if B == 0:
        return z, zs
    lib = _lib.load()
    if bf:
        ws = gemm_workspace(B, D, Cc, dev)
        _lib.check(lib.iif_linear_fwd_bf16(_ptr(x), _ld(x), _ptr(w), _ld(w), _ptr(bias), _ptr(cs), _ptr(z), Cc,
                                           _ptr(zs), Cc, B, D, Cc, _ptr(ws), 0 if ws is None else ws.numel(),
                                           _stream(dev)), "linear_fwd_bf16")
    else:
        _lib.check(lib.iif_linear_fwd_f32(_ptr(x), _ld(x), _ptr(w), _ld(w), _ptr(bias), _ptr(cs), _ptr(z), Cc,
                                          _ptr(zs), Cc, B, D, Cc, _stream(dev)), "linear_fwd_f32")
    return z, zs


def linear_bwd_dx(dz, w, alpha=None, out_bf16=False):
    bf = dz.dtype == torch.bfloat16
    dz = _bf16_rows(dz, "dz") if bf else _rows(dz, "dz", torch.float32)
    w = _bf16_rows(w, "w") if bf else _rows(w, "w", torch.float32)
    B, Cc = dz.shape
    D = w.shape[1]
    dev = dz.device
    al = _alpha(alpha, dev)
    lib = _lib.load()
    if bf:
        dx = torch.empty(B, D, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
        if B:
            ws = gemm_workspace(B, D, Cc, dev)
            _lib.check(lib.iif_linear_bwd_dx_bf16(_ptr(dz), _ld(dz), _ptr(w), _ld(w), _ptr(al), _ptr(dx),
                                                  _lib.DTYPE_BF16 if out_bf16 else _lib.DTYPE_F32, D, B, D, Cc,
                                                  _ptr(ws), 0 if ws is None else ws.numel(), _stream(dev)),
                       "linear_bwd_dx_bf16")
    else:
        dx = torch.empty(B, D, dtype=torch.float32, device=dev)
        if B:
            _lib.check(lib.iif_linear_bwd_dx_f32(_ptr(dz), _ld(dz), _ptr(w), _ld(w), _ptr(al), _ptr(dx), D, B, D, Cc,
                                                 _stream(dev)), "linear_bwd_dx_f32")
    return dx


def linear_bwd_dw(dz, x, alpha=None):
    bf = dz.dtype == torch.bfloat16
    dz = _bf16_rows(dz, "dz") if bf else _rows(dz, "dz", torch.float32)
    x = _bf16_rows(x, "x") if bf else _rows(x, "x", torch.float32)
    B, Cc = dz.shape
    D = x.shape[1]
    dev = dz.device
    al = _alpha(alpha, dev)
    dw = torch.empty(Cc, D, dtype=torch.float32, device=dev)
    lib = _lib.load()
    if bf:
        ws = gemm_workspace(B, D, Cc, dev)
        _lib.check(lib.iif_linear_bwd_dw_bf16(_ptr(dz), _ld(dz), _ptr(x), _ld(x), _ptr(al), _ptr(dw), D, B, D, Cc,
                                              _ptr(ws), 0 if ws is None else ws.numel(), _stream(dev)),
                   "linear_bwd_dw_bf16")
    else:
        _lib.check(lib.iif_linear_bwd_dw_f32(_ptr(dz), _ld(dz), _ptr(x), _ld(x), _ptr(al), _ptr(dw), D, B, D, Cc,
                                             _stream(dev)), "linear_bwd_dw_f32")
    return dw


def linear_bwd(dz, x, w, alpha=None, *, need_dx=True, need_db=True, dx_bf16=False):
    """AddmmBackward of fc_cls in ONE launch (bf16 tcgen05): returns (dX | None, dW, db | None)."""
    dz = _bf16_rows(dz, "dz")
    x = _bf16_rows(x, "x")
    w = _bf16_rows(w, "w")
    B, Cc = dz.shape
    D = x.shape[1]
    if tuple(w.shape) != (Cc, D) or x.shape[0] != B:
        raise ValueError("linear_bwd: shape mismatch")
    dev = dz.device
    al = _alpha(alpha, dev)
    dx = torch.empty(B, D, dtype=torch.bfloat16 if dx_bf16 else torch.float32, device=dev) if need_dx else None
    flat = torch.empty(Cc * D + (Cc if need_db else 0), dtype=torch.float32, device=dev)
    dw = flat[:Cc * D].view(Cc, D)
    db = flat[Cc * D:] if need_db else None
    ws = gemm_workspace(B, D, Cc, dev)
    _lib.check(_lib.load().iif_linear_bwd_bf16(
        _ptr(dz), _ld(dz), _ptr(x), _ld(x), _ptr(w), _ld(w), _ptr(al), _ptr(dx),
        _lib.DTYPE_BF16 if dx_bf16 else _lib.DTYPE_F32, D, _ptr(dw), D, _ptr(db), B, D, Cc, _ptr(ws),
        0 if ws is None else ws.numel(), _stream(dev)), "linear_bwd_bf16")
    return dx, dw, db


class HeadStep:
    """Pre-allocated buffers + one C call (`iif_head_fwd_bwd_bf16`) per head step.

    fc_cls -> IIF softmax-CE fwd+bwd -> db, dX, dW with bf16 GEMM operands.  Buffers are allocated
    once so the step can be captured in a CUDA graph.  dW and db live in ONE flat fp32 buffer
    (`grad_flat`) so the data-parallel all-reduce of the head's parameter gradients is one message."""

    def __init__(self, B, D, Cc, device, *, need_dx=True, dx_bf16=True, need_db=True, want_acc=False, ws=None,
                 fused_loss=True, grad_flat=None, persistent=True, scratch=None):
        dev = torch.device(device)
        self.B, self.D, self.C, self.device = B, D, Cc, dev
        f32, i32 = torch.float32, torch.int32
        self.z = torch.empty(B, Cc, dtype=f32, device=dev)
        self.loss_i = torch.empty(B, dtype=f32, device=dev)
        self.loss = torch.zeros((), dtype=f32, device=dev)
        self.dz = torch.empty(B, pad64(Cc), dtype=torch.bfloat16, device=dev)
        self.dx = torch.empty(B, D, dtype=torch.bfloat16 if dx_bf16 else f32, device=dev) if need_dx else None
        ng = Cc * D + (Cc if need_db else 0)
        if grad_flat is not None:      # caller-owned storage (e.g. a peer-mapped buffer of parallel.PeerAllReduce)
            if grad_flat.dtype != f32 or grad_flat.numel() != ng or not grad_flat.is_contiguous() or not grad_flat.is_cuda:
                raise ValueError(f"grad_flat must be a contiguous fp32 CUDA tensor of {ng} elements")
        self.grad_flat = grad_flat if grad_flat is not None else torch.empty(ng, dtype=f32, device=dev)
        self.dw = self.grad_flat[:Cc * D].view(Cc, D)
        self.db = self.grad_flat[Cc * D:] if need_db else None
        self.argmax = torch.empty(B, dtype=i32, device=dev) if want_acc else None
        self.rank = torch.empty(B, dtype=i32, device=dev) if want_acc else None
        self.acc_counts = torch.zeros(2, dtype=i32, device=dev) if want_acc else None
        self.scratch = scratch if scratch is not None else \
            torch.zeros((int(_lib.load().iif_loss_scratch_bytes(B)) + 3) // 4, dtype=i32, device=dev)
        n = int(_lib.load().iif_gemm_ws_bytes(B, D, Cc))
        if ws is not None and ws.numel() < n:
            raise ValueError("HeadStep: shared workspace too small")
        self.ws = ws if ws is not None else torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
        self.ws_bytes = n
        self.dx_dtype = _lib.DTYPE_BF16 if dx_bf16 else _lib.DTYPE_F32
        self.fused_loss = bool(fused_loss)
        self.persistent = bool(persistent)   # the whole step in ONE persistent launch when the shape qualifies
        self.launches_per_step = 3     # refined by bind(): 1 = persistent step, 2 = loss rows ride in the backward launch
        self._args = None
        self._keep = None

    def bind(self, x, w, bias, iif, label, *, class_weight=None, sample_weight=None, ignore_index=-100,
             scale=None):
        """Fix the input tensors of this step (their storage must stay alive and in place; new values
        are copied INTO them).  Fills the C argument struct once; `launch()` then costs one C call."""
        B, D, Cc = self.B, self.D, self.C
        _cuda(x, "x", torch.bfloat16)
        _cuda(w, "w", torch.bfloat16)
        _cuda(label, "label", torch.int64)
        if tuple(x.shape) != (B, D) or tuple(w.shape) != (Cc, D) or label.numel() != B:
            raise ValueError("HeadStep: shape mismatch")
        if x.stride(1) != 1 or w.stride(1) != 1:
            raise ValueError("HeadStep: x and w need unit inner stride")
        bias = _vec(bias, "bias", Cc)
        iif = _vec(iif, "iif", Cc)
        class_weight = _vec(class_weight, "class_weight", Cc)
        sample_weight = _vec(sample_weight, "sample_weight", B)
        a = _lib.HeadArgs()
        a.x, a.ldx, a.w, a.ldw = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0)
        a.bias = None if bias is None else bias.data_ptr()
        a.iif = None if iif is None else iif.data_ptr()
        a.label = label.data_ptr()
        a.class_weight = None if class_weight is None else class_weight.data_ptr()
        a.sample_weight = None if sample_weight is None else sample_weight.data_ptr()
        a.ignore_index = int(ignore_index)
        a.scale = float(1.0 / B if scale is None else scale)
        a.B, a.D, a.C = B, D, Cc
        a.z, a.ldz = self.z.data_ptr(), Cc
        a.loss_i, a.loss_sum = self.loss_i.data_ptr(), self.loss.data_ptr()
        a.dz_bf16, a.lddz = self.dz.data_ptr(), self.dz.stride(0)
        a.dx = None if self.dx is None else self.dx.data_ptr()
        a.dx_dtype, a.lddx = self.dx_dtype, D
        a.dw, a.lddw = self.dw.data_ptr(), D
        a.db = None if self.db is None else self.db.data_ptr()
        a.argmax = None if self.argmax is None else self.argmax.data_ptr()
        a.rank = None if self.rank is None else self.rank.data_ptr()
        a.acc_counts = None if self.acc_counts is None else self.acc_counts.data_ptr()
        a.scratch = self.scratch.data_ptr()
        a.ws, a.ws_bytes = self.ws.data_ptr(), self.ws_bytes
        a.flags = (0 if self.fused_loss else _lib.HEAD_NO_FUSED_LOSS) | (0 if self.persistent else _lib.HEAD_NO_PERSISTENT)
        n = int(_lib.load().iif_head_launches(C.byref(a)))
        if n < 0:
            _lib.check(n, "head_launches")
        self.launches_per_step = n
        self._args = a
        self._keep = (x, w, bias, iif, label, class_weight, sample_weight)
        self._fn = _lib.load().iif_head_fwd_bwd_bf16
        return self

    def launch(self):
        """Enqueue the bound step on the current stream; returns the 0-dim device loss."""
        rc = self._fn(C.byref(self._args), _stream(self.device))
        if rc:
            _lib.check(rc, "head_fwd_bwd_bf16")
        return self.loss

    def run(self, x, w, bias, iif, label, **kw):
        self.bind(x, w, bias, iif, label, **kw)
        return self.launch()

    def kernels(self):
        """The individual launches of the bound step as (name, callable) pairs -- same buffers, same C
        entry points as `launch()`; used by bench.py to time each kernel on its own."""
        lib, a = _lib.load(), self._args
        st = lambda: _stream(self.device)
        p = C.c_void_p
        if self.launches_per_step == 1:
            return [("head_step_fused_bf16", lambda: lib.iif_head_fwd_bwd_bf16(C.byref(a), st()))]
        out = [("linear_fwd_bf16", lambda: lib.iif_linear_fwd_bf16(p(a.x), a.ldx, p(a.w), a.ldw, p(a.bias), None,
                                                                   p(a.z), a.ldz, None, 0, a.B, a.D, a.C, p(a.ws),
                                                                   a.ws_bytes, st()))]
        if self.launches_per_step == 2:
            out.append(("loss_linear_bwd_bf16", lambda: lib.iif_loss_linear_bwd_bf16(C.byref(a), st())))
            return out
        out.append(("softmax_ce_fwd_bwd", lambda: lib.iif_softmax_ce_fwd_bwd(
            p(a.z), a.ldz, p(a.iif), p(a.label), p(a.class_weight), p(a.sample_weight), a.ignore_index, a.scale,
            a.B, a.C, p(a.loss_i), p(a.loss_sum), None, 0, p(a.dz_bf16), a.lddz, None, p(a.argmax), p(a.rank),
            p(a.acc_counts), p(a.scratch), st())))
        out.append(("linear_bwd_bf16", lambda: lib.iif_linear_bwd_bf16(
            p(a.dz_bf16), a.lddz, p(a.x), a.ldx, p(a.w), a.ldw, None, p(a.dx), a.dx_dtype, a.lddx, p(a.dw), a.lddw,
            p(a.db), a.B, a.D, a.C, p(a.ws), a.ws_bytes, st())))
        return out

class SigmoidHeadStep:
    """The head step in SIGMOID mode (mmdet CrossEntropyLoss(use_sigmoid=True) / FasaIIFLoss(use_sigmoid=True) as
    loss_cls, seg/mmdet/models/losses/cross_entropy_loss.py:74-111; no IIF scale in this mode, fasa_iif_loss.py:35-36):
    fc_cls GEMM -> sigmoid-BCE fwd+bwd (bf16 dZ, no one-hot tensor) -> dX, dW, db in one grouped launch -- three
    launches on pre-allocated buffers, same layout as `HeadStep` (flat [dW | db] gradient buffer)."""

    def __init__(self, B, D, Cc, device, *, need_dx=True, dx_bf16=True, need_db=True, grad_flat=None, ws=None):
        dev = torch.device(device)
        self.B, self.D, self.C, self.device = B, D, Cc, dev
        f32 = torch.float32
        self.z = torch.empty(B, Cc, dtype=f32, device=dev)
        self.loss_i = torch.empty(B, dtype=f32, device=dev)
        self.loss = torch.zeros((), dtype=f32, device=dev)
        self.dz = torch.empty(B, pad64(Cc), dtype=torch.bfloat16, device=dev)
        self.dz[:, Cc:].zero_()
        self.dx = torch.empty(B, D, dtype=torch.bfloat16 if dx_bf16 else f32, device=dev) if need_dx else None
        ng = Cc * D + (Cc if need_db else 0)
        self.grad_flat = grad_flat if grad_flat is not None else torch.empty(ng, dtype=f32, device=dev)
        self.dw = self.grad_flat[:Cc * D].view(Cc, D)
        self.db = self.grad_flat[Cc * D:] if need_db else None
        self.scratch = torch.zeros((int(_lib.load().iif_loss_scratch_bytes(B)) + 3) // 4, dtype=torch.int32, device=dev)
        n = int(_lib.load().iif_gemm_ws_bytes(B, D, Cc))
        self.ws = ws if ws is not None else torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
        self.ws_bytes = n
        self.dx_dtype = _lib.DTYPE_BF16 if dx_bf16 else _lib.DTYPE_F32
        self.launches_per_step = 3

    def bind(self, x, w, bias, label, *, pos_weight=None, sample_weight=None, ignore_index=-100, scale=None):
        B, D, Cc = self.B, self.D, self.C
        _cuda(x, "x", torch.bfloat16)
        _cuda(w, "w", torch.bfloat16)
        _cuda(label, "label", torch.int64)
        if tuple(x.shape) != (B, D) or tuple(w.shape) != (Cc, D) or label.numel() != B:
            raise ValueError("SigmoidHeadStep: shape mismatch")
        self._bias = _vec(bias, "bias", Cc)
        self._pw = _vec(pos_weight, "pos_weight", Cc)
        self._sw = _vec(sample_weight, "sample_weight", B)
        self._x, self._w, self._y = x, w, label
        self._ign = int(ignore_index)
        self._scale = float(1.0 / (B * Cc) if scale is None else scale)      # 'mean' over B*C elements
        return self

    def kernels(self):
        lib, p = _lib.load(), C.c_void_p
        st = lambda: _stream(self.device)
        B, D, Cc = self.B, self.D, self.C
        x, w, y = self._x, self._w, self._y
        return [
            ("linear_fwd_bf16", lambda: lib.iif_linear_fwd_bf16(p(x.data_ptr()), x.stride(0), p(w.data_ptr()), w.stride(0),
                                                                _ptr(self._bias), None, _ptr(self.z), Cc, None, 0, B, D, Cc,
                                                                _ptr(self.ws), self.ws_bytes, st())),
            ("sigmoid_bce_fwd_bwd", lambda: lib.iif_sigmoid_bce_fwd_bwd(
                _ptr(self.z), Cc, _ptr(y), _ptr(self._pw), None, _ptr(self._sw), self._ign, self._scale, B, Cc, None, 0,
                _ptr(self.loss_i), _ptr(self.loss), None, 0, _ptr(self.dz), self.dz.stride(0), _ptr(self.scratch), st())),
            ("linear_bwd_bf16", lambda: lib.iif_linear_bwd_bf16(
                _ptr(self.dz), self.dz.stride(0), p(x.data_ptr()), x.stride(0), p(w.data_ptr()), w.stride(0), None,
                _ptr(self.dx), self.dx_dtype, D, _ptr(self.dw), D, _ptr(self.db), B, D, Cc, _ptr(self.ws), self.ws_bytes,
                st())),
        ]

    def launch(self):
        for name, fn in self.kernels():
            rc = fn()
            if rc:
                _lib.check(rc, name)
        return self.loss


class HeadPipeline:
    """Host-batch pipeline over bound `HeadStep` slots (C: iif_pipeline_*, csrc/pipeline.cu).

    `submit(slot, host_x, host_label)` enqueues H2D copy -> head step -> D2H of the loss on three
    library-owned streams and returns at once; `wait(slot)` blocks until that step's loss is in host
    memory and returns it.  Consecutive slots overlap: the next batch's PCIe copy hides under the
    current step's kernels.  This is the loop of cls/train.py:66-77 (batch from the loader ->
    criterion -> loss.item()) for the head alone."""

    def __init__(self, steps):
        if not steps or any(hs._args is None for hs in steps):
            raise ValueError("HeadPipeline needs bound HeadStep slots")
        self.steps = list(steps)
        n = len(self.steps)
        arr = (_lib.HeadArgs * n)()
        for i, hs in enumerate(self.steps):
            arr[i] = hs._args
        self._h = C.c_void_p()
        self._lib = _lib.load()
        torch.cuda.synchronize(self.steps[0].device)      # the slots' buffers were produced on torch streams
        _lib.check(self._lib.iif_pipeline_create(C.byref(self._h), arr, n), "pipeline_create")
        self.host_loss = torch.zeros(n, dtype=torch.float32).pin_memory()
        self._loss_ptr = self.host_loss.data_ptr()
        self._staged_last = [False] * n        # which mode produced each slot's latest loss

    def submit(self, slot: int, host_x: torch.Tensor, host_label: torch.Tensor) -> None:
        hs = self.steps[slot]
        if host_x.is_cuda or host_label.is_cuda:
            raise RuntimeError("HeadPipeline.submit takes HOST tensors (use HeadStep.launch for device inputs)")
        if host_x.dtype != torch.bfloat16 or tuple(host_x.shape) != (hs.B, hs.D) or not host_x.is_contiguous():
            raise ValueError(f"host_x must be a contiguous bf16 [{hs.B},{hs.D}] tensor")
        if host_label.dtype != torch.int64 or host_label.numel() != hs.B or not host_label.is_contiguous():
            raise ValueError(f"host_label must be a contiguous int64 [{hs.B}] tensor")
        rc = self._lib.iif_pipeline_submit(self._h, slot, host_x.data_ptr(), host_label.data_ptr(),
                                           self._loss_ptr + 4 * slot)
        if rc:
            _lib.check(rc, "pipeline_submit")
        self._staged_last[slot] = False

    def submit_device(self, slot: int) -> None:
        """The step on the slot's current device buffers (no copies)."""
        rc = self._lib.iif_pipeline_submit_device(self._h, slot)
        if rc:
            _lib.check(rc, "pipeline_submit_device")

    def join(self, stream: torch.cuda.Stream) -> None:
        """Make `stream` wait on the device for everything enqueued so far on all of the pipeline's streams."""
        _lib.check(self._lib.iif_pipeline_join(self._h, C.c_void_p(stream.cuda_stream)), "pipeline_join")

    def set_allreduce(self, peer) -> None:
        """Data-parallel runs: all-reduce(mean) every step's gradients with `peer` (parallel.PeerAllReduce whose
        buffer(i) is slot i's grad_flat) on the pipeline's comm stream, overlapping the following steps."""
        n = len(self.steps)
        for i, hs in enumerate(self.steps):
            if hs.grad_flat.data_ptr() != peer.buffer(i).data_ptr():
                raise ValueError("slot gradients must live in the PeerAllReduce buffers (HeadStep(grad_flat=peer.buffer(i)))")
        offs = (C.c_int64 * n)(*[i * peer.stride for i in range(n)])
        self._peer = peer
        _lib.check(self._lib.iif_pipeline_set_allreduce(self._h, peer._bufs, peer._flags, peer._mc, peer.rank, peer.world,
                                                        offs, (peer.numel + 3) // 4 * 4, peer.num_ctas, peer.num_threads,
                                                        peer.lanes),
                   "pipeline_set_allreduce")

    def streams(self):
        """(h2d, compute, d2h, comm) as torch ExternalStreams (to record timing events on them)."""
        ptrs = [C.c_void_p() for _ in range(4)]
        _lib.check(self._lib.iif_pipeline_get_streams(self._h, *[C.byref(p) for p in ptrs]), "pipeline_get_streams")
        dev = self.steps[0].device
        return tuple(torch.cuda.ExternalStream(p.value, device=dev) for p in ptrs)

    def enable_staged(self) -> None:
        """Staged mode: library-owned pinned host staging per slot + one CUDA graph per slot (this slot's
        launches with the H2D of the NEXT slot's staged batch as a parallel branch): one driver call per step.
        Fill `staging(k)` ahead of `submit_staged(k - 1)`; walk the slots round-robin."""
        import numpy as np
        _lib.check(self._lib.iif_pipeline_enable_staged(self._h), "pipeline_enable_staged")
        self._staging = []
        for k, hs in enumerate(self.steps):
            px, py, pl = C.c_void_p(), C.c_void_p(), C.c_void_p()
            _lib.check(self._lib.iif_pipeline_staging(self._h, k, C.byref(px), C.byref(py), C.byref(pl)), "pipeline_staging")
            xb = (C.c_uint16 * (hs.B * hs.D)).from_address(px.value)
            yb = (C.c_int64 * hs.B).from_address(py.value)
            x = torch.from_numpy(np.frombuffer(xb, dtype=np.uint16).reshape(hs.B, hs.D)).view(torch.bfloat16)
            y = torch.from_numpy(np.frombuffer(yb, dtype=np.int64))
            self._staging.append((x, y, C.c_float.from_address(pl.value)))

    def staging(self, slot: int):
        """(x [B,D] bf16, label [B] int64): CPU tensors aliasing the slot's pinned staging buffers."""
        return self._staging[slot][0], self._staging[slot][1]

    def submit_staged(self, slot: int) -> None:
        rc = self._lib.iif_pipeline_submit_staged(self._h, slot)
        if rc:
            _lib.check(rc, "pipeline_submit_staged")
        self._staged_last[slot] = True

    def submit_staged_ring(self) -> None:
        """One step of every slot, in order, from the staging buffers: ONE graph launch for len(steps) steps."""
        rc = self._lib.iif_pipeline_submit_staged_ring(self._h)
        if rc:
            _lib.check(rc, "pipeline_submit_staged_ring")
        for k in range(len(self.steps)):
            self._staged_last[k] = True

    def wait(self, slot: int) -> float:
        rc = self._lib.iif_pipeline_wait(self._h, slot)
        if rc:
            _lib.check(rc, "pipeline_wait")
        return float(self._staging[slot][2].value) if self._staged_last[slot] else float(self.host_loss[slot])

    def stream_wait_step(self, slot: int, stream: torch.cuda.Stream) -> None:
        _lib.check(self._lib.iif_pipeline_stream_wait_step(self._h, slot, C.c_void_p(stream.cuda_stream)),
                   "pipeline_stream_wait_step")

    def hold_slot(self, slot: int, stream: torch.cuda.Stream) -> None:
        _lib.check(self._lib.iif_pipeline_hold_slot(self._h, slot, C.c_void_p(stream.cuda_stream)), "pipeline_hold_slot")

    def sync(self) -> None:
        if self._h:
            _lib.check(self._lib.iif_pipeline_sync(self._h), "pipeline_sync")

    def close(self) -> None:
        if self._h:
            self._lib.iif_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
