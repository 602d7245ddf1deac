"""`rwkvfla.layers.rwkv6.LoRA` (imported by /root/reference/model/test/test_performance.py:29-37)."""
from __future__ import annotations

import torch
import torch.nn as nn


class LoRA(nn.Module):
    """x -> up(act(down(x))) [+ bias]; parameters `lora.0.weight` [low,in], `lora.2.weight` [out,low],
    `lora.2.bias` [out] -- the names the reference's optimizer grouping keys on
    (train_spark_rwkv7speech_jsonl.py:161-172)."""

    def __init__(self, input_dim: int, output_dim: int, low_rank_dim: int, bias: bool = True,
                 activation: str | None = "tanh"):
        super().__init__()
        self.input_dim, self.output_dim, self.low_rank_dim, self.bias = input_dim, output_dim, low_rank_dim, bias
        if activation is None:
            act = nn.Identity()
        elif activation == "sigmoid":
            act = nn.Sigmoid()
        elif activation == "tanh":
            act = nn.Tanh()
        elif activation == "relu":
            act = nn.ReLU()
        else:
            raise ValueError(f"Not supported activation `{activation}`.")
        self.activation = activation
        self.lora = nn.Sequential(nn.Linear(input_dim, low_rank_dim, bias=False), act,
                                  nn.Linear(low_rank_dim, output_dim, bias=bias))
        nn.init.zeros_(self.lora[0].weight)
        shape = self.lora[2].weight.shape
        gain = (shape[0] / shape[1]) ** 0.5 if shape[0] > shape[1] else 1.0
        nn.init.orthogonal_(self.lora[2].weight, gain=gain * 0.1)
        if bias:
            nn.init.zeros_(self.lora[2].bias)

    def set_bias_value(self, value):
        with torch.no_grad():
            if isinstance(value, torch.Tensor):
                self.lora[2].bias.copy_(value.to(self.lora[2].bias.dtype))
            else:
                self.lora[2].bias.fill_(float(value))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.lora(x)
