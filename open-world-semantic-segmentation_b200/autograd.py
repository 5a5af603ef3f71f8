"""torch.autograd bindings of the CUDA kernels (distance logits, fused DCE+VL loss)."""
from __future__ import annotations

from typing import Optional

import torch

from . import head as H
from ._lib import check, lib, ptr, require_cuda, stream_ptr


class _DistanceLogits(torch.autograd.Function):
    """z = -||x - mu_k||^2 (anomaly/models/models.py:645-651; network/utils.py:98-111).
    Backward: dx = -2 [ (sum_k g_k) x - sum_k g_k mu_k ]."""

    @staticmethod
    def forward(ctx, x, centers, magnitude):
        out = H.dml_head(x, centers=centers, magnitude=magnitude, want_logits=True, label_dtype=None)
        ctx.save_for_backward(x)
        ctx.centers = centers
        ctx.magnitude = magnitude
        return out.logits

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        gsum = g.sum(dim=1, keepdim=True)
        if ctx.centers is None:
            # mu = m I  =>  sum_k g_k mu_kd = m g_d
            dx = -2.0 * (gsum * x - ctx.magnitude * g)
        else:
            mu = ctx.centers.to(device=x.device, dtype=x.dtype)
            dx = -2.0 * (gsum * x - torch.einsum("bkhw,kd->bdhw", g, mu))
        return dx, None, None


def distance_logits(x: torch.Tensor, centers: Optional[torch.Tensor] = None, magnitude: float = H.DEFAULT_MAGNITUDE):
    """Differentiable distance logits [B,K,H,W] from an NCHW embedding."""
    if centers is not None:
        m = H.scaled_identity_magnitude(centers)
        if m is not None and centers.shape[0] == x.shape[1]:
            centers, magnitude = None, m
    if not x.requires_grad:
        return H.dml_head(x, centers=centers, magnitude=magnitude, want_logits=True, label_dtype=None).logits
    return _DistanceLogits.apply(x, centers, magnitude)


class _DmlLoss(torch.autograd.Function):
    """Fused DCE + alpha*VL + beta*Inter straight from the embedding (kernel (b))."""

    @staticmethod
    def forward(ctx, x, target, centers, magnitude, alpha, beta, ignore_index, is_logits):
        require_cuda(x, "x")
        require_cuda(target, "target")
        # the C ABI takes raw const float* / label pointers: anything else (fp16 / bf16 under autocast, float64,
        # a target of another size) must be rejected or converted HERE, not read out of bounds by the kernel
        if x.dim() != 4:
            raise ValueError("x must be a [B,D,H,W] tensor")
        if x.dtype != torch.float32:
            raise ValueError(f"x must be float32, got {x.dtype} (dml_loss() upcasts fp16 / bf16 / fp64 inputs)")
        x = x.contiguous()
        target = target.contiguous()
        B, D, Hh, Ww = x.shape
        if target.numel() != B * Hh * Ww:
            raise ValueError(f"target must hold B*H*W = {B * Hh * Ww} labels, got {tuple(target.shape)}")
        dev = x.device
        mu = None if centers is None else centers.to(device=dev, dtype=torch.float32).contiguous()
        K = D if mu is None else mu.shape[0]
        if target.dtype not in (torch.uint8, torch.int64):
            target = target.to(torch.int64)
        out5 = torch.empty(5, dtype=torch.float64, device=dev)
        nbytes = lib().dml_loss_workspace_bytes(B, Hh, Ww)
        partials = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        u8 = target.dtype == torch.uint8
        with torch.cuda.device(dev):
            check(lib().dml_loss_forward(ptr(x), 1 if is_logits else 0, ptr(mu), magnitude, ptr(target) if u8 else None,
                                         None if u8 else ptr(target), ignore_index, B, D, K, Hh, Ww, alpha, beta,
                                         ptr(partials), ptr(out5), stream_ptr(dev)), "dml_loss_forward")
        ctx.save_for_backward(x, target, out5)
        ctx.mu, ctx.magnitude, ctx.alpha, ctx.beta, ctx.ignore_index, ctx.K = mu, magnitude, alpha, beta, ignore_index, K
        ctx.is_logits = is_logits
        ctx.mark_non_differentiable(out5)
        return out5[0].to(torch.float32), out5

    @staticmethod
    def backward(ctx, grad_loss, _grad_parts):
        x, target, out5 = ctx.saved_tensors
        B, D, Hh, Ww = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        g = grad_loss.to(device=dev, dtype=torch.float32).contiguous().view(1)
        u8 = target.dtype == torch.uint8
        with torch.cuda.device(dev):
            check(lib().dml_loss_backward(ptr(x), 1 if ctx.is_logits else 0, ptr(ctx.mu), ctx.magnitude, ptr(target) if u8 else None,
                                          None if u8 else ptr(target), ctx.ignore_index, B, D, ctx.K, Hh, Ww, ctx.alpha,
                                          ctx.beta, ptr(out5), ptr(g), ptr(dx), stream_ptr(dev)), "dml_loss_backward")
        return dx, None, None, None, None, None, None, None


def dml_loss(x: torch.Tensor, target: torch.Tensor, *, centers: Optional[torch.Tensor] = None,
             magnitude: float = H.DEFAULT_MAGNITUDE, alpha: float = 0.0, beta: float = 0.0, ignore_index: int = 255,
             return_parts: bool = False, input_is_logits: bool = False):
    """loss = (CE + alpha*VL + beta*Inter)/n from the embedding ``x`` [B,D,H,W] and labels [B,H,W].
    ``return_parts`` also returns the float64 device vector (loss, CE, VL, Inter, n_valid).
    ``input_is_logits``: ``x`` holds the logits z [B,K,H,W] (the reference criterion's calling convention)."""
    if centers is not None:
        m = H.scaled_identity_magnitude(centers)
        if m is not None and centers.shape[0] == x.shape[1]:
            centers, magnitude = None, m
    if x.dtype != torch.float32:
        if not x.is_floating_point():
            raise ValueError(f"x must be a floating-point tensor, got {x.dtype}")
        x = x.float()        # differentiable cast (autocast fp16 / bf16, float64): the kernels compute in fp32
    loss, parts = _DmlLoss.apply(x, target, centers, float(magnitude), float(alpha), float(beta), int(ignore_index),
                                 bool(input_is_logits))
    return (loss, parts) if return_parts else loss
