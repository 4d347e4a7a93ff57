"""Autograd entry points of the IIF head (host-side mirror of the reference's torch calls).

Each Function replaces a chain of ATen ops of the reference by one fused kernel:
  iif_cross_entropy  <- pred*iif ; F.cross_entropy(reduction='none', weight, ignore_index) ; *weight ; sum
                        (classification/custom.py:30 ; mmdet iif_loss.py:187-200 ; losses/utils.py:42-55)
  sigmoid_bce        <- _expand_onehot_labels ; F.binary_cross_entropy_with_logits ; *weight ; sum
                        (mmdet cross_entropy_loss.py:53-111 ; classification/custom.py:61-73)
  linear             <- F.linear and its AddmmBackward (resnet_pytorch.py:293 ; bbox_head.py:118)
  iif_head_loss      <- the three above chained without leaving the GPU stream (bf16 GEMM mode)
"""
from __future__ import annotations

import torch

from . import ops


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


class _IIFCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, iif, target, class_weight, sample_weight, ignore_index, scale, reduce, target_b=None,
                lam=1.0):
        need = ctx.needs_input_grad[0]
        r = ops.softmax_ce(_f32(pred), iif, target, class_weight=class_weight, sample_weight=sample_weight,
                           ignore_index=ignore_index, scale=scale, want_dz_f32=need, want_sum=reduce,
                           label_b=target_b, lam=lam)
        ctx.in_dtype = pred.dtype
        if need:
            ctx.save_for_backward(r["dz_f32"])
        return r["loss_sum"] if reduce else r["loss_i"]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        out = ops.scale_rows(dz, g) if dz.numel() else dz
        if ctx.in_dtype != torch.float32:
            out = out.to(ctx.in_dtype)
        return out, None, None, None, None, None, None, None, None, None


def iif_cross_entropy(pred, iif, target, *, class_weight=None, sample_weight=None, ignore_index=-100,
                      scale=1.0, reduce=True, target_b=None, lam=1.0):
    """scale * sum_i w_i cw[y_i] CE(pred_i * iif, y_i)  (reduce=True) or the per-sample vector.
    With `target_b`: the Mixup pair lam*CE(target) + (1-lam)*CE(target_b) from one pass over the logits."""
    return _IIFCrossEntropy.apply(pred, iif, target, class_weight, sample_weight, ignore_index, scale, reduce,
                                  target_b, lam)


class _SigmoidBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, pos_weight, col_weight, sample_weight, ignore_index, scale, mode, gamma=0.0,
                alpha=None):
        need = ctx.needs_input_grad[0]
        r = ops.sigmoid_bce(_f32(pred), target, pos_weight=pos_weight, col_weight=col_weight,
                            sample_weight=sample_weight, ignore_index=ignore_index, scale=scale,
                            want_elem=(mode == "elem"), want_dz_f32=need, want_sum=(mode == "sum"), gamma=gamma,
                            alpha=alpha)
        ctx.in_dtype, ctx.mode = pred.dtype, mode
        if need:
            ctx.save_for_backward(r["dz_f32"])
        return r["loss_sum"] if mode == "sum" else r["loss_elem"]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        if ctx.mode == "sum":
            out = ops.scale_rows(dz, g) if dz.numel() else dz
        else:  # elementwise upstream gradient (reduction='none'): not a hot path
            out = dz * g
        if ctx.in_dtype != torch.float32:
            out = out.to(ctx.in_dtype)
        return out, None, None, None, None, None, None, None, None, None


def sigmoid_bce(pred, target, *, pos_weight=None, col_weight=None, sample_weight=None, ignore_index=-100,
                scale=1.0, reduce=True, gamma=0.0, alpha=None):
    """Sigmoid BCE (gamma = 0) or focal loss (gamma > 0, optional alpha balance) with fused backward."""
    return _SigmoidBCE.apply(pred, target, pos_weight, col_weight, sample_weight, ignore_index, scale,
                             "sum" if reduce else "elem", gamma, alpha)


class _SigmoidBCEDense(torch.autograd.Function):
    """Sigmoid BCE against already-expanded (dense / soft) targets [B,C] with element / row weights
    (mmdet cross_entropy_loss.py:100-106, the pred.dim() == label.dim() branch)."""

    @staticmethod
    def forward(ctx, pred, target, pos_weight, weight, scale, mode):
        need = ctx.needs_input_grad[0]
        r = ops.sigmoid_bce_dense(_f32(pred), _f32(target), pos_weight=pos_weight, weight=weight, scale=scale,
                                  want_elem=(mode == "elem"), want_dz=need, want_sum=(mode == "sum"))
        ctx.in_dtype, ctx.mode = pred.dtype, mode
        if need:
            ctx.save_for_backward(r["dz_f32"])
        return r["loss_sum"] if mode == "sum" else r["loss_elem"]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        out = (ops.scale_rows(dz, g) if dz.numel() else dz) if ctx.mode == "sum" else dz * g
        if ctx.in_dtype != torch.float32:
            out = out.to(ctx.in_dtype)
        return out, None, None, None, None, None


def sigmoid_bce_dense(pred, target, *, pos_weight=None, weight=None, scale=1.0, reduce=True):
    return _SigmoidBCEDense.apply(pred, target, pos_weight, weight, scale, "sum" if reduce else "elem")


class _NormalizeRows(torch.autograd.Function):
    """y_i = pre_i r(|pre_i x_i|) x_i -- the operand normalisation of the normalised classifiers
    (mmdet normed_predictor.py:36-40,70-76; cls/resnet_cifar.py:66-71) with its exact backward
    dx_i = a_i g_i + c_i (x_i . g_i) x_i, all in the library's row kernels (csrc/norm.cu)."""

    @staticmethod
    def forward(ctx, x, pre, mode, temperature, power, eps):
        x2 = _f32(x.reshape(-1, x.shape[-1]))
        need = ctx.needs_input_grad[0]
        a, c = ops.row_scale_from_norm(x2, mode, pre=pre, temperature=temperature, power=power, eps=eps, want_c=need)
        y = ops.rows_axpby(x2, a)
        if need:
            ctx.save_for_backward(x2, a, c)
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, g):
        x2, a, c = ctx.saved_tensors
        g2 = _f32(g.reshape(-1, g.shape[-1]))
        dot = ops.row_dot(x2, g2)
        dx = ops.rows_axpby(g2, a, x2, c, dot)
        return dx.to(ctx.x_dtype).reshape(ctx.x_shape), None, None, None, None, None


def normalize_rows(x, mode, *, pre=None, temperature=1.0, power=1.0, eps=1e-6):
    """Rows of `x` times r(|pre x|) (and pre): see include/iif_b200.h IIF_NORM_*."""
    return _NormalizeRows.apply(x, pre, mode, temperature, power, eps)


class _Linear(torch.autograd.Function):
    """Z = X W^T + b with the head's own GEMM kernels; `bf16` selects the tcgen05 path."""

    @staticmethod
    def forward(ctx, x, weight, bias, bf16, weight_bf16):
        x2 = x.reshape(-1, x.shape[-1])
        if bf16:
            xo = x2 if x2.dtype == torch.bfloat16 else ops.scale_rows(_f32(x2), None, bf16=True)
            wo = weight_bf16 if weight_bf16 is not None else ops.scale_rows(_f32(weight), None, bf16=True)
        else:
            xo, wo = _f32(x2), _f32(weight)
        z, _ = ops.linear_fwd(xo, wo, bias)
        ctx.save_for_backward(xo, wo)
        ctx.bf16, ctx.x_shape, ctx.x_dtype, ctx.has_bias = bf16, x.shape, x.dtype, bias is not None
        return z.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gz):
        xo, wo = ctx.saved_tensors
        gz2 = _f32(gz.reshape(-1, gz.shape[-1]))
        dz = ops.scale_rows(gz2, None, bf16=True, pad_ld=True) if ctx.bf16 else gz2
        dx = dw = db = None
        if ctx.bf16 and gz2.shape[0] > 0:
            dx, dw, _ = ops.linear_bwd(dz, xo, wo, need_dx=ctx.needs_input_grad[0], need_db=False,
                                       dx_bf16=(ctx.x_dtype == torch.bfloat16))
            if dx is not None:
                dx = dx.to(ctx.x_dtype).reshape(ctx.x_shape)
            if not ctx.needs_input_grad[1]:
                dw = None
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = ops.colsum(gz2)          # fp32 upstream gradient: exact bias gradient
            return dx, dw, db, None, None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_bwd_dx(dz, wo, out_bf16=(ctx.x_dtype == torch.bfloat16 and ctx.bf16))
            dx = dx.to(ctx.x_dtype).reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = ops.linear_bwd_dw(dz, xo)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(gz2)
        return dx, dw, db, None, None


def linear(x, weight, bias=None, *, bf16=False, weight_bf16=None):
    return _Linear.apply(x, weight, bias, bf16, weight_bf16)


class _IIFHeadLoss(torch.autograd.Function):
    """fc_cls -> IIF softmax-CE in one autograd node (bf16 GEMM operands, fp32 accumulate / logits).

    forward : Z = X W^T + b (tcgen05), fused loss kernel emitting loss and bf16 dZ (never fp32 dZ in HBM)
    backward: dX = g dZ W, dW = g dZ^T X, db = g sum_i dZ_i in ONE grouped tcgen05 launch, the upstream scalar g read on device
    Returns (loss, raw logits Z); Z is non-differentiable (training accuracy is taken on raw logits,
    classification/train.py:81, mmdet iif_loss.py:103)."""

    @staticmethod
    def forward(ctx, x, weight, bias, iif, target, class_weight, sample_weight, ignore_index, scale,
                weight_bf16):
        x2 = x.reshape(-1, x.shape[-1])
        xo = x2 if x2.dtype == torch.bfloat16 else ops.scale_rows(_f32(x2), None, bf16=True)
        wo = weight_bf16 if weight_bf16 is not None else ops.scale_rows(_f32(weight), None, bf16=True)
        z, _ = ops.linear_fwd(xo, wo, bias)
        B = z.shape[0]
        sc = (1.0 / max(B, 1)) if scale is None else scale
        r = ops.softmax_ce(z, iif, target, class_weight=class_weight, sample_weight=sample_weight,
                           ignore_index=ignore_index, scale=sc, want_dz_f32=False, want_dz_bf16=True)
        ctx.save_for_backward(xo, wo, r["dz_bf16"])
        ctx.x_shape, ctx.x_dtype, ctx.has_bias, ctx.C = x.shape, x.dtype, bias is not None, z.shape[1]
        ctx.mark_non_differentiable(z)
        return r["loss_sum"], z

    @staticmethod
    def backward(ctx, g, _gz):
        xo, wo, dzp = ctx.saved_tensors
        dz = dzp[:, :ctx.C]
        if dz.shape[0] == 0:
            return (
                torch.zeros(ctx.x_shape, dtype=ctx.x_dtype, device=dz.device) if ctx.needs_input_grad[0] else None,
                torch.zeros(wo.shape, dtype=torch.float32, device=dz.device) if ctx.needs_input_grad[1] else None,
                torch.zeros(wo.shape[0], dtype=torch.float32, device=dz.device)
                if (ctx.has_bias and ctx.needs_input_grad[2]) else None, None, None, None, None, None, None, None)
        # dX, dW and db in one launch; the upstream scalar g is read on the device (no host sync)
        dx, dw, db = ops.linear_bwd(dz, xo, wo, alpha=g, need_dx=ctx.needs_input_grad[0],
                                    need_db=ctx.has_bias and ctx.needs_input_grad[2],
                                    dx_bf16=(ctx.x_dtype == torch.bfloat16))
        if dx is not None:
            dx = dx.to(ctx.x_dtype).reshape(ctx.x_shape)
        if not ctx.needs_input_grad[1]:
            dw = None
        return dx, dw, db, None, None, None, None, None, None, None


def iif_head_loss(x, weight, bias, iif, target, *, class_weight=None, sample_weight=None, ignore_index=-100,
                  scale=None, weight_bf16=None):
    """(loss, logits) of the fused head; `scale` defaults to 1/B (classification 'mean')."""
    return _IIFHeadLoss.apply(x, weight, bias, iif, target, class_weight, sample_weight, ignore_index, scale,
                              weight_bf16)
