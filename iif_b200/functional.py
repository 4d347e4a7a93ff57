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
    """Z = X W^T + b with the head's own GEMM kernels; `bf16` selects the path: True = bf16 operands on the tensor
    cores, False = fp32 FFMA (csrc/gemm_f32.cu), "x3" = fp32 operands split into three bf16 terms and multiplied on the
    tensor cores as six partial products (csrc/split3.cu) -- the two fp32 modes meet the 1e-5 parity bar."""

    @staticmethod
    def forward(ctx, x, weight, bias, bf16, weight_bf16):
        x2 = x.reshape(-1, x.shape[-1])
        if bf16 == "x3":
            xf, wf = _f32(x2), _f32(weight)
            z, _ = ops.linear_fwd(ops.split3(xf, k_along_rows=False, side_b=False),
                                  ops.split3(wf, k_along_rows=False, side_b=True), bias)
            ctx.save_for_backward(xf, wf)
            ctx.bf16, ctx.x_shape, ctx.x_dtype, ctx.has_bias = bf16, x.shape, x.dtype, bias is not None
            return z.reshape(*x.shape[:-1], weight.shape[0])
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
        dx = dw = db = None
        if ctx.bf16 == "x3":
            if gz2.shape[0] > 0:
                if ctx.needs_input_grad[0]:       # dX = dZ W: K = classes (columns of dZ, rows of W)
                    dx = ops.linear_bwd_dx(ops.split3(gz2, k_along_rows=False, side_b=False),
                                           ops.split3(wo, k_along_rows=True, side_b=True))
                    dx = dx.to(ctx.x_dtype).reshape(ctx.x_shape)
                if ctx.needs_input_grad[1]:       # dW = dZ^T X: K = batch rows of both
                    dw = ops.linear_bwd_dw(ops.split3(gz2, k_along_rows=True, side_b=False),
                                           ops.split3(xo, k_along_rows=True, side_b=True))
            else:
                dx = torch.zeros(ctx.x_shape, dtype=ctx.x_dtype, device=gz.device) if ctx.needs_input_grad[0] else None
                dw = torch.zeros_like(wo) if ctx.needs_input_grad[1] else None
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = ops.colsum(gz2)
            return dx, dw, db, None, None
        dz = ops.scale_rows(gz2, None, bf16=True, pad_ld=True) if ctx.bf16 else gz2
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


class _SiblingLinear(torch.autograd.Function):
    """fc_cls and fc_reg of a bbox head on the SAME RoI features in one GEMM per direction:
    [Z_cls | Z_reg] = X [W_cls ; W_reg]^T + [b_cls | b_reg]  (bbox_head.py:118-119 runs two F.linear calls),
    backward dX = [dZ_cls | dZ_reg] [W_cls ; W_reg], [dW_cls ; dW_reg] = [dZ_cls | dZ_reg]^T X and the bias sums in ONE
    grouped tcgen05 launch -- X is read once per direction instead of twice.  The packed bf16 operand is built by the
    same per-step cast kernels the separate layers run (each weight cast straight into its row block)."""

    @staticmethod
    def forward(ctx, x, w_cls, b_cls, w_reg, b_reg):
        x2 = x.reshape(-1, x.shape[-1])
        xo = x2 if x2.dtype == torch.bfloat16 else ops.scale_rows(_f32(x2), None, bf16=True)
        C1, C2, D = w_cls.shape[0], w_reg.shape[0], w_cls.shape[1]
        wo = torch.empty(C1 + C2, D, dtype=torch.bfloat16, device=x.device)
        ops.scale_rows(_f32(w_cls), None, bf16=True, out=wo[:C1])
        ops.scale_rows(_f32(w_reg), None, bf16=True, out=wo[C1:])
        bias = None
        if b_cls is not None or b_reg is not None:
            bias = torch.zeros(C1 + C2, dtype=torch.float32, device=x.device)
            if b_cls is not None:
                bias[:C1] = b_cls.detach()
            if b_reg is not None:
                bias[C1:] = b_reg.detach()
        z, _ = ops.linear_fwd(xo, wo, bias)
        ctx.save_for_backward(xo, wo)
        ctx.meta = (x.shape, x.dtype, C1, C2, b_cls is not None, b_reg is not None)
        lead = x.shape[:-1]
        return z[:, :C1].reshape(*lead, C1), z[:, C1:].reshape(*lead, C2)

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        xo, wo = ctx.saved_tensors
        x_shape, x_dtype, C1, C2, has_bc, has_br = ctx.meta
        B = xo.shape[0]
        dz = torch.empty(B, ops.pad8(C1 + C2), dtype=torch.bfloat16, device=xo.device)
        for g, lo, n in ((g_cls, 0, C1), (g_reg, C1, C2)):
            if g is None:
                dz[:, lo:lo + n].zero_()
            else:
                ops.scale_rows(_f32(g.reshape(-1, n)), None, bf16=True, out=dz[:, lo:lo + n])
        if B == 0:
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=xo.device)
            return (torch.zeros(x_shape, dtype=x_dtype, device=xo.device), z(C1, wo.shape[1]), z(C1) if has_bc else None,
                    z(C2, wo.shape[1]), z(C2) if has_br else None)
        dx, dw, db = ops.linear_bwd(dz[:, :C1 + C2], xo, wo, need_dx=ctx.needs_input_grad[0], need_db=True,
                                    dx_bf16=(x_dtype == torch.bfloat16))
        if dx is not None:
            dx = dx.to(x_dtype).reshape(x_shape)
        return (dx, dw[:C1] if ctx.needs_input_grad[1] else None, db[:C1] if has_bc and ctx.needs_input_grad[2] else None,
                dw[C1:] if ctx.needs_input_grad[3] else None, db[C1:] if has_br and ctx.needs_input_grad[4] else None)


def sibling_linear(x, w_cls, b_cls, w_reg, b_reg):
    """(cls_score, bbox_pred) of two linear layers on the same features: one GEMM forward, one grouped launch backward."""
    return _SiblingLinear.apply(x, w_cls, b_cls, w_reg, b_reg)


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
        B, D, Cc = xo.shape[0], xo.shape[1], wo.shape[0]
        ctx.fused = False
        if B > 0 and x.dtype in (torch.float32, torch.bfloat16):
            # the whole step -- forward, loss, dX, dW, db -- in ONE persistent launch when the shape qualifies
            # (csrc/head_fused.cu); the gradients are formed here with upstream g = 1 and backward() applies g
            dev = xo.device
            hs = ops.HeadStep(B, D, Cc, dev, need_dx=ctx.needs_input_grad[0], dx_bf16=(x.dtype == torch.bfloat16),
                              need_db=bias is not None, ws=ops.gemm_workspace(B, D, Cc, dev),
                              scratch=ops.loss_scratch(dev, B))
            hs.bind(xo.contiguous(), wo.contiguous(), bias, iif, target, class_weight=class_weight,
                    sample_weight=sample_weight, ignore_index=ignore_index, scale=scale)
            if hs.launches_per_step == 1:
                hs.launch()
                ctx.fused = True
                ctx.hs = hs
                ctx.x_shape, ctx.x_dtype, ctx.has_bias, ctx.C = x.shape, x.dtype, bias is not None, Cc
                ctx.mark_non_differentiable(hs.z)
                return hs.loss, hs.z
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
        if ctx.fused:
            hs = ctx.hs
            g32 = g.detach().float()
            ops.scale_inplace_(hs.grad_flat, g32)            # dW | db in one flat buffer; no-op when g == 1
            dx = None
            if hs.dx is not None and ctx.needs_input_grad[0]:
                dx = ops.scale_inplace_(hs.dx, g32).reshape(ctx.x_shape)
            dw = hs.dw if ctx.needs_input_grad[1] else None
            db = hs.db if (ctx.has_bias and ctx.needs_input_grad[2]) else None
            return dx, dw, db, None, None, None, None, None, None, None
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
