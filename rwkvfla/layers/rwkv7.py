"""`rwkvfla.layers.rwkv7.RWKV7Attention`: RWKV-7 time-mix with the rwkvfla parameter names
(`x_r..x_g, k_k, k_a, r_k, r_proj, k_proj, v_proj, o_proj, {w,a,v,g}_lora.lora.{0,2}, g_norm`;
map to the BlinkDL names: /root/reference/utils/convert_rwkv.py:17-41).

The math is rwkvtts_b200.core.tmix (one restatement for both stacks): rwkvfla's log-decay
w = -0.6065*sigmoid(w_lora(xw)) equals log(exp(-exp(w_pre))) with w_pre = -softplus(-w_lora(xw)) - 0.5,
which is what the CUDA op takes.  The recurrent state kept in the Cache is fp32 [B,H,64,64] VALUE-major
(the layout of the reference's own CUDA ops); `Cache.to_fla_layout()` gives rwkvfla's key-major view.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from rwkvtts_b200 import core, ops
from .rwkv6 import LoRA


class RWKV7Attention(nn.Module):
    def __init__(self, mode: str = "chunk", hidden_size: int = 1024, head_dim: Optional[int] = 64,
                 num_heads: Optional[int] = None, decay_low_rank_dim: int = 64, gate_low_rank_dim: int = 128,
                 a_low_rank_dim: int = 64, v_low_rank_dim: int = 16, elementwise_affine: bool = True,
                 norm_eps: float = 1e-5, layer_idx: int = None, fuse_norm: bool = False, value_dim: int = None,
                 num_hidden_layers: int = None, **kwargs):
        super().__init__()
        assert mode in ("chunk", "fused_recurrent"), f"Not supported mode `{mode}`."
        if head_dim is None and num_heads is None:
            raise ValueError("Either `head_dim` or `num_heads` must be specified.")
        self.mode, self.hidden_size = mode, hidden_size
        self.head_dim = head_dim if head_dim is not None else hidden_size // num_heads
        self.num_heads = hidden_size // self.head_dim
        if self.head_dim != core.HEAD:
            raise ValueError("the CUDA library is compiled for head_dim 64")
        self.key_dim = hidden_size
        self.value_dim = value_dim if value_dim is not None else hidden_size
        if self.value_dim != hidden_size:
            raise ValueError("value_dim != hidden_size is not used by any RWKVTTS configuration")
        self.layer_idx, self.num_hidden_layers, self.fuse_norm = layer_idx, num_hidden_layers, fuse_norm
        C = hidden_size
        for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g"):
            setattr(self, n, nn.Parameter(torch.zeros(1, 1, C)))
        self.k_k = nn.Parameter(torch.zeros(C))
        self.k_a = nn.Parameter(torch.zeros(C))
        self.r_k = nn.Parameter(torch.zeros(self.num_heads, self.head_dim))
        self.r_proj = nn.Linear(C, C, bias=False)
        self.k_proj = nn.Linear(C, C, bias=False)
        self.v_proj = nn.Linear(C, C, bias=False)
        self.o_proj = nn.Linear(C, C, bias=False)
        self.w_lora = LoRA(C, C, low_rank_dim=decay_low_rank_dim, activation="tanh")
        if layer_idx != 0:
            self.v_lora = LoRA(C, C, low_rank_dim=v_low_rank_dim, activation=None)
        self.a_lora = LoRA(C, C, low_rank_dim=a_low_rank_dim, activation=None)
        self.g_lora = LoRA(C, C, low_rank_dim=gate_low_rank_dim, activation="sigmoid", bias=False)
        self.g_norm = nn.GroupNorm(self.num_heads, C, eps=self.head_dim * norm_eps, affine=elementwise_affine)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """BlinkDL's layer-dependent init (rwkv_s2s_single_ffn.py:74-156), in rwkvfla's parameter names."""
        if self.layer_idx is None or self.num_hidden_layers is None:
            return
        C, N, L, i = self.hidden_size, self.head_dim, self.num_hidden_layers, self.layer_idx
        r01, r10 = i / max(L - 1, 1), 1.0 - i / L
        n = torch.arange(C, dtype=torch.float32)
        ddd = (n / C).view(1, 1, C)
        linear = n / (C - 1) - 0.5
        zz = ((n % N) - (N - 1) / 2) / ((N - 1) / 2)
        zigzag = zz * zz.abs()
        www = -6 + 6 * (n / (C - 1)) ** (1 + r01 ** 0.3)
        for name, e in (("x_r", 0.2), ("x_w", 0.9), ("x_k", 0.7), ("x_v", 0.7), ("x_a", 0.9), ("x_g", 0.2)):
            getattr(self, name).copy_(1.0 - torch.pow(ddd, e * r10))
        self.k_a.fill_(1.02)
        self.r_k.fill_(-0.04)
        self.k_k.copy_(0.71 - linear * 0.1)
        self.w_lora.set_bias_value(www + 0.5 + zigzag * 2.5)
        self.a_lora.set_bias_value(-0.19 + zigzag * 0.3 + linear * 0.4)
        if i != 0:
            self.v_lora.set_bias_value(0.73 - linear * 0.4)
        self.g_norm.weight.fill_(((i + 1) / L) ** 0.7)
        nn.init.orthogonal_(self.r_proj.weight)
        nn.init.orthogonal_(self.k_proj.weight, gain=0.1)
        nn.init.orthogonal_(self.v_proj.weight)
        self.o_proj.weight.zero_()

    def params(self) -> core.TmixParams:
        lo = lambda m: (m.lora[0].weight.t(), m.lora[2].weight.t(), m.lora[2].bias)
        w1, w2, w0 = lo(self.w_lora)
        a1, a2, a0 = lo(self.a_lora)
        g1, g2, _ = lo(self.g_lora)
        v1 = v2 = v0 = None
        if self.layer_idx != 0:
            v1, v2, v0 = lo(self.v_lora)
        return core.TmixParams(
            x_r=self.x_r, x_w=self.x_w, x_k=self.x_k, x_v=self.x_v, x_a=self.x_a, x_g=self.x_g,
            w0=w0, w1=w1, w2=w2, a0=a0, a1=a1, a2=a2, v0=v0, v1=v1, v2=v2, g1=g1, g2=g2,
            k_k=self.k_k, k_a=self.k_a, r_k=self.r_k,
            W_r=self.r_proj.weight, W_k=self.k_proj.weight, W_v=self.v_proj.weight, W_o=self.o_proj.weight,
            ln_w=self.g_norm.weight, ln_b=self.g_norm.bias, ln_eps=self.g_norm.eps)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                past_key_values=None, use_cache: Optional[bool] = False, output_attentions: Optional[bool] = False,
                v_first: torch.Tensor = None, cu_seqlens: Optional[torch.LongTensor] = None, **kwargs):
        plan = None
        if cu_seqlens is not None:
            # packed varlen batch [1, T_total, D] (chunk_rwkv7(..., cu_seqlens=) of rwkv-fla): the chunked kernels and the
            # token shift restart at every boundary; RWKV7Model hands down one ops.VarlenPlan for all layers
            plan = cu_seqlens if isinstance(cu_seqlens, ops.VarlenPlan) else ops.VarlenPlan(cu_seqlens, hidden_states.shape[1])
            if hidden_states.shape[0] != 1 or attention_mask is not None or use_cache:
                raise ValueError("cu_seqlens expects one packed sequence [1, T_total, D] without attention_mask / cache")
        B, T, _ = hidden_states.shape
        am = None
        if attention_mask is not None:
            assert attention_mask.dim() == 2, "Expected attention_mask as a 0-1 matrix [batch_size, seq_len]"
            am = core.mask3(attention_mask, T, hidden_states.dtype)
        last = None
        if past_key_values is not None and len(past_key_values) > self.layer_idx:
            last = past_key_values[self.layer_idx]
        shift = last["conv_state"] if last is not None else None
        state = last["recurrent_state"] if last is not None else None
        # rwkvfla masks the input and v only (r,w,k of padded positions see x = 0)
        out, v_first, new_shift, new_state = core.tmix(
            self.params(), self.layer_idx, hidden_states, v_first, mask=am, mask_rwk=False,
            shift_state=shift, wkv_state=state, need_state=bool(use_cache),
            inplace_state=not torch.is_grad_enabled(), plan=plan)
        if past_key_values is not None and use_cache:
            past_key_values.update(recurrent_state=new_state, conv_state=new_shift, layer_idx=self.layer_idx,
                                   offset=T)
        return out, None, past_key_values, v_first
