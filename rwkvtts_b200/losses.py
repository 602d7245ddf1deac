"""Cosy-layout loss: same constructor, arguments and value as

    LabelSmoothingLoss        /root/reference/third_party/cosyvoice/transformer/label_smoothing_loss.py:20-96

(`RWKV7CosyLM.criterion_ce`, model/llm/cosy_llm.py:145).  The reference materialises the smoothed one-hot target
[tokens, V], the log-softmax, the element-wise KL tensor and its masked copy -- four [tokens, V] tensors (V = 6562) -- and
reads the number of valid tokens back to the host.  The KL divergence against a smoothed one-hot has a closed form per
row,  c log c + (V-1) e log e - c logp[t] - e (sum_j logp[j] - logp[t]),  c = 1 - smoothing, e = smoothing / (V - 1),
so one log-softmax (kept in fp32), one row sum and one gather are enough, and the normaliser stays on the device."""
from __future__ import annotations

import math

import torch
from torch import nn


class LabelSmoothingLoss(nn.Module):
    def __init__(self, size: int, padding_idx: int, smoothing: float, normalize_length: bool = False):
        super().__init__()
        self.padding_idx = padding_idx
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.size = size
        self.normalize_length = normalize_length

    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """x (batch, seqlen, class) logits; target (batch, seqlen) with `padding_idx` on ignored positions.  Sum of the
        per-token KL terms over valid tokens, divided by their number (normalize_length) or by the batch size."""
        assert x.size(2) == self.size
        batch_size = x.size(0)
        logp = torch.log_softmax(x.reshape(-1, self.size).float(), dim=1)
        target = target.reshape(-1)
        ignore = target == self.padding_idx
        tgt = target.masked_fill(ignore, 0)
        c, e = self.confidence, self.smoothing / (self.size - 1)
        xlogx = lambda p: p * math.log(p) if p > 0 else 0.0           # KLDivLoss takes 0 log 0 = 0
        lt = logp.gather(1, tgt.unsqueeze(1)).squeeze(1)
        row = (xlogx(c) + (self.size - 1) * xlogx(e)) - c * lt - e * (logp.sum(1) - lt)
        total = row.masked_fill(ignore, 0.0).sum()
        denom = (~ignore).sum() if self.normalize_length else batch_size
        return (total / denom).to(x.dtype)


def xy_channel_losses(hidden_states: torch.Tensor, heads, labels: torch.Tensor, label_smoothing: float = 0.0,
                      ignore_index: int = -100, num_chunks: int = 8) -> torch.Tensor:
    """Sum over the codebook channels of the mean cross entropy of `heads[i](hidden_states)` against `labels[:, :, i]`
    -- the loss loop of RWKV7XYLM.forward (/root/reference/model/llm/xy_llm.py:233-240) -- without ever holding a
    channel's logits: channel 0 has 66.7 k classes, [2, 8192, 66.7 k] bf16 = 2.2 GB per GPU at config c5, and the
    reference keeps all eight logit tensors alive for the backward.  Each head goes through the chunked linear +
    cross-entropy of the `rwkvfla` package (token chunks recomputed in the backward).  A channel without a valid label
    gives NaN, as `nn.CrossEntropyLoss` does."""
    from rwkvfla.modules import FusedLinearCrossEntropyLoss
    crit = FusedLinearCrossEntropyLoss(ignore_index=ignore_index, label_smoothing=label_smoothing, num_chunks=num_chunks,
                                       reduction="sum")
    total = None
    for i, head in enumerate(heads):
        y = labels[:, :, i]
        loss = crit(hidden_states, y, head.weight, head.bias) / (y != ignore_index).sum()
        total = loss if total is None else total + loss
    return total
