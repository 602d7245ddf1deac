"""`rwkvfla.modules`: LayerNorm and the loss modules the reference imports."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from rwkvtts_b200 import core, fused

from .l2warp import l2_warp


class LayerNorm(nn.LayerNorm):
    """rwkvfla.modules.LayerNorm(hidden_size, elementwise_affine=True, bias=False, eps=1e-5)
    (used as rwkv_asr_whisper.py:17).  With (x, residual, prenorm=True) it returns
    (norm(x + residual), x + residual) like rwkvfla's fused add+norm."""

    def __init__(self, hidden_size: int, elementwise_affine: bool = True, bias: bool = False, eps: float = 1e-5):
        super().__init__(hidden_size, eps=eps, elementwise_affine=elementwise_affine, bias=bias)

    def forward(self, x, residual=None, prenorm: bool = False):
        if core.FUSED and self.elementwise_affine and fused.ln_usable(x) and (residual is None or residual.shape == x.shape):
            # one kernel for the residual add and the norm (and one for their adjoint): csrc/tmix_fused.cu
            y, total = fused.add_layernorm(x, residual, self.weight, self.bias, self.eps)
            return (y, total) if prenorm else y
        if residual is not None:
            x = x + residual
        y = super().forward(x)
        return (y, x) if prenorm else y


class FusedCrossEntropyLoss(nn.CrossEntropyLoss):
    """Mean cross entropy over labels != ignore_index (spark_llm.py:8,:139-158)."""

    def __init__(self, ignore_index: int = -100, reduction: str = "mean", label_smoothing: float = 0.0,
                 inplace_backward: bool = False, **kwargs):
        super().__init__(ignore_index=ignore_index, reduction=reduction, label_smoothing=label_smoothing)

    def forward(self, input, target):
        return F.cross_entropy(input.float(), target, ignore_index=self.ignore_index, reduction=self.reduction,
                               label_smoothing=self.label_smoothing)


class FusedLinearCrossEntropyLoss(nn.Module):
    """loss(hidden [.., D], labels [..], weight [V, D], bias) = mean CE of hidden @ weight.T without ever
    holding the full [tokens, V] logits: token chunks are recomputed in the backward."""

    def __init__(self, ignore_index: int = -100, label_smoothing: float = 0.0, num_chunks: int = 8,
                 reduction: str = "mean", use_l2warp: bool = False, **kwargs):
        super().__init__()
        self.ignore_index, self.label_smoothing = ignore_index, label_smoothing
        self.num_chunks, self.reduction, self.use_l2warp = num_chunks, reduction, use_l2warp

    def _chunk(self, h, y, weight, bias, l2_factor=0.0):
        logits = F.linear(h, weight, bias).float()
        loss = F.cross_entropy(logits, y, ignore_index=self.ignore_index, reduction="sum",
                               label_smoothing=self.label_smoothing)
        if l2_factor:
            loss = l2_warp(loss, logits.unsqueeze(0), l2_factor)       # [1, n, V]: l2_warp divides by the first two dims
        return loss

    def forward(self, x, target, weight, bias=None):
        h = x.reshape(-1, x.shape[-1])
        y = target.reshape(-1)
        if core.FUSED and not self.use_l2warp and self.reduction in ("mean", "sum") and fused.linear_ce_usable(h, weight, bias):
            # CUDA bf16: logits GEMM -> CE kernel (loss + gradient in place) -> gradient GEMMs, chunk by chunk
            return fused.linear_cross_entropy(h, y, weight, self.ignore_index, self.label_smoothing, self.reduction, bias=bias)
        n = max(1, min(self.num_chunks, h.shape[0]))
        total = h.new_zeros((), dtype=torch.float32)
        for hc, yc in zip(h.chunk(n), y.chunk(n)):
            # l2_warp (rwkvfla.modules.l2warp): gradient 1e-4 / tokens on each row's max logit
            l2f = 1e-4 * hc.shape[0] / max(1, h.shape[0]) if self.use_l2warp else 0.0
            if torch.is_grad_enabled() and (hc.requires_grad or weight.requires_grad):
                total = total + checkpoint(self._chunk, hc, yc, weight, bias, l2f, use_reentrant=False)
            else:
                total = total + self._chunk(hc, yc, weight, bias, l2f)
        if self.reduction == "sum":
            return total
        return total / (y != self.ignore_index).sum().clamp(min=1)


__all__ = ["LayerNorm", "FusedCrossEntropyLoss", "FusedLinearCrossEntropyLoss", "l2_warp"]
