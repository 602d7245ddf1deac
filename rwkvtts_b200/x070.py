"""The reference's in-repo RWKV-7 stack behind the same class / function names and parameter names,
running on the CUDA library of this repo.

  RWKV_Tmix_x070, RWKV_CMix_x070, Block, RWKV7S2S_SingleFFN, L2Wrap
        /root/reference/model/llm/rwkv_s2s_single_ffn.py:61-330
  RWKV_x070_TMix_one / _seq, RWKV_x070_CMix_one / _seq  (stateful inference, B = 1)
        /root/reference/model/llm/rwkv_s2s_single_ffn.py:482-556

A state_dict of the reference loads unchanged (same parameter names and shapes).  `args` is the same
Namespace the reference takes (n_embd, n_layer, head_size_a, head_size_divisor, vocab_size,
text_vocab_size, audio_vocab_size, dropout, grad_cp, need_init_tmix, need_init_cmix).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from . import core, ops


def _ortho(shape, scale):
    x = torch.zeros(shape)
    gain = math.sqrt(shape[0] / shape[1]) if shape[0] > shape[1] else 1
    nn.init.orthogonal_(x, gain=gain * scale)
    return x


class RWKV_Tmix_x070(nn.Module):
    def __init__(self, args, layer_id):
        super().__init__()
        self.args, self.layer_id = args, layer_id
        self.head_size = args.head_size_a
        assert self.head_size == ops.HEAD_SIZE, "the CUDA library is compiled for head size 64 (:10-12)"
        C = args.n_embd
        assert C % self.head_size == 0
        H = self.n_head = C // self.head_size
        N = self.head_size
        r01 = layer_id / max(args.n_layer - 1, 1)
        r10 = 1.0 - layer_id / args.n_layer
        n = torch.arange(C, dtype=torch.float32)
        ddd = (n / C).view(1, 1, C)
        linear = n / (C - 1) - 0.5
        zz = ((n % N) - (N - 1) / 2) / ((N - 1) / 2)
        zigzag = zz * zz.abs()
        www = -6 + 6 * (n / (C - 1)) ** (1 + r01 ** 0.3)
        for name, e in (("x_r", 0.2), ("x_w", 0.9), ("x_k", 0.7), ("x_v", 0.7), ("x_a", 0.9), ("x_g", 0.2)):
            setattr(self, name, nn.Parameter(1.0 - torch.pow(ddd, e * r10)))
        lora = lambda f: max(32, int(round(f * C ** 0.5 / 32) * 32))
        Dw, Da, Dv, Dg = lora(1.8), lora(1.8), lora(1.3), 128
        self.w1 = nn.Parameter(torch.zeros(C, Dw))
        self.w2 = nn.Parameter(_ortho((Dw, C), 0.1))
        self.w0 = nn.Parameter((www + 0.5 + zigzag * 2.5).view(1, 1, C))
        self.a1 = nn.Parameter(torch.zeros(C, Da))
        self.a2 = nn.Parameter(_ortho((Da, C), 0.1))
        self.a0 = nn.Parameter((-0.19 + zigzag * 0.3 + linear * 0.4).view(1, 1, C))
        if layer_id != 0:
            self.v1 = nn.Parameter(torch.zeros(C, Dv))
            self.v2 = nn.Parameter(_ortho((Dv, C), 0.1))
            self.v0 = nn.Parameter((0.73 - linear * 0.4).view(1, 1, C))
        self.g1 = nn.Parameter(torch.zeros(C, Dg))
        self.g2 = nn.Parameter(_ortho((Dg, C), 0.1))
        self.k_k = nn.Parameter((0.71 - linear * 0.1).view(1, 1, C))
        self.k_a = nn.Parameter(torch.full((1, 1, C), 1.02))
        self.r_k = nn.Parameter(torch.full((H, N), -0.04))
        self.receptance = nn.Linear(C, C, bias=False)
        self.key = nn.Linear(C, C, bias=False)
        self.value = nn.Linear(C, C, bias=False)
        self.output = nn.Linear(C, C, bias=False)
        self.ln_x = nn.GroupNorm(H, C, eps=1e-5 * args.head_size_divisor ** 2)
        if getattr(args, "need_init_tmix", False):
            self._init_params(args)

    def _init_params(self, args):
        C = args.n_embd
        self.receptance.weight.data.uniform_(-0.5 / C ** 0.5, 0.5 / C ** 0.5)
        self.key.weight.data.uniform_(-0.05 / C ** 0.5, 0.05 / C ** 0.5)
        self.value.weight.data.uniform_(-0.5 / C ** 0.5, 0.5 / C ** 0.5)
        self.output.weight.data.zero_()

    def params(self) -> core.TmixParams:
        g = lambda n: getattr(self, n, None)
        return core.TmixParams(
            x_r=self.x_r, x_w=self.x_w, x_k=self.x_k, x_v=self.x_v, x_a=self.x_a, x_g=self.x_g,
            w0=self.w0, w1=self.w1, w2=self.w2, a0=self.a0, a1=self.a1, a2=self.a2,
            v0=g("v0"), v1=g("v1"), v2=g("v2"), g1=self.g1, g2=self.g2,
            k_k=self.k_k, k_a=self.k_a, r_k=self.r_k,
            W_r=self.receptance.weight, W_k=self.key.weight, W_v=self.value.weight, W_o=self.output.weight,
            ln_w=self.ln_x.weight, ln_b=self.ln_x.bias, ln_eps=self.ln_x.eps)

    def forward(self, x, attention_mask=None, v_first=None):
        out, v_first, _, _ = core.tmix(self.params(), self.layer_id, x, v_first, mask=attention_mask)
        return out, v_first

    @torch.inference_mode()
    def forward_batch(self, x, attention_mask=None, v_first=None, x_prev=None, state=None):
        """Stateful inference over a batch (/root/reference/model/llm/rwkv_asr_cuda_whisper.py:181-215): `x_prev` [B,C] is
        the token-shift state, `state` fp32 [B,H,64,64] is advanced IN PLACE (RWKV7_BATCH_OP, :210).  Only x and v are
        masked (:184, :209).  Returns (out, v_first, out[:, -1], state): the third value is the last row of the OUTPUT, as
        in the reference (:215 returns `x[:,-1,:]` after `x` has been reassigned to the projection)."""
        out, v_first, _, new_state = core.tmix(self.params(), self.layer_id, x, v_first, mask=attention_mask, mask_rwk=False,
                                               mask_kk=False, shift_state=x_prev, wkv_state=state, need_state=True,
                                               inplace_state=True)
        if state is not None and new_state is not state:
            state.copy_(new_state)
            new_state = state
        return out, v_first, out[:, -1, :], new_state


class RWKV_CMix_x070(nn.Module):
    def __init__(self, args, layer_id):
        super().__init__()
        self.args, self.layer_id = args, layer_id
        C = args.n_embd
        r10 = 1.0 - layer_id / args.n_layer
        ddd = (torch.arange(C, dtype=torch.float32) / C).view(1, 1, C)
        self.x_k = nn.Parameter(1.0 - torch.pow(ddd, r10 ** 4))
        self.key = nn.Linear(C, C * 4, bias=False)
        self.value = nn.Linear(C * 4, C, bias=False)
        if getattr(args, "need_init_cmix", False):
            self._init_params(args)

    def _init_params(self, args):
        self.key.weight.data.uniform_(-0.5 / args.n_embd ** 0.5, 0.5 / args.n_embd ** 0.5)
        self.value.weight.data.zero_()

    def forward(self, x, attention_mask):
        return core.cmix(self.x_k, self.key.weight, self.value.weight, x, mask=attention_mask)[0]

    @torch.inference_mode()
    def forward_batch(self, x, attention_mask=None, x_prev=None):
        """rwkv_asr_cuda_whisper.py:277-285: returns (out, last masked input row = the next call's x_prev)."""
        return core.cmix(self.x_k, self.key.weight, self.value.weight, x, mask=attention_mask, shift_state=x_prev,
                         need_state=True)


class Block(nn.Module):
    def __init__(self, args, layer_id):
        super().__init__()
        self.args, self.layer_id = args, layer_id
        self.ln1 = nn.LayerNorm(args.n_embd)
        self.ln2 = nn.LayerNorm(args.n_embd)
        if layer_id == 0:
            self.ln0 = nn.LayerNorm(args.n_embd)
        self.att = RWKV_Tmix_x070(args, layer_id)
        self.ffn = RWKV_CMix_x070(args, layer_id)
        if args.dropout > 0:
            self.drop0 = nn.Dropout(p=args.dropout)
            self.drop1 = nn.Dropout(p=args.dropout)

    def forward(self, x, attention_mask, v_first=None):
        if self.layer_id == 0:
            x = self.ln0(x)
        x_attn, v_first = self.att(self.ln1(x), attention_mask, v_first)
        x = x + x_attn
        x = x + self.ffn(self.ln2(x), attention_mask)
        return x, v_first

    @torch.inference_mode()
    def forward_batch(self, x, attention_mask=None, v_first=None, tx_prev=None, state=None, cx_prev=None):
        """rwkv_asr_cuda_whisper.py:318-326."""
        if self.layer_id == 0:
            x = self.ln0(x)
        x_attn, v_first, tx_prev, state = self.att.forward_batch(self.ln1(x), attention_mask, v_first, tx_prev, state)
        x = x + x_attn
        x_ffn, cx_prev = self.ffn.forward_batch(self.ln2(x), attention_mask, cx_prev)
        x = x + x_ffn
        return x, v_first, tx_prev, state, cx_prev


class L2Wrap(torch.autograd.Function):
    """:261-274 -- pulls the largest logit towards 0 in the backward."""

    @staticmethod
    def forward(ctx, loss, y):
        ctx.save_for_backward(y)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        y = ctx.saved_tensors[0]
        factor = 1e-4 / (y.shape[0] * y.shape[1])
        maxx, ids = torch.max(y, -1, keepdim=True)
        gy = torch.zeros_like(y)
        gy.scatter_(-1, ids, maxx * factor)
        return grad_output, gy


class RWKV7S2S_SingleFFN(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        assert args.n_embd % 32 == 0
        self.emb = nn.Embedding(args.vocab_size, args.n_embd)
        self.blocks = nn.ModuleList([Block(args, i) for i in range(args.n_layer)])
        self.ln_out = nn.LayerNorm(args.n_embd)
        self.head = nn.Linear(args.n_embd, args.text_vocab_size, bias=False)
        self.audio_head = nn.Linear(args.n_embd, args.audio_vocab_size, bias=False)
        if args.dropout > 0:
            self.drop0 = nn.Dropout(p=args.dropout)

    def forward(self, idx, attention_mask=None, is_text=True):
        args = self.args
        B, T = idx.size()
        if attention_mask is None:
            attention_mask = torch.ones(B, T, dtype=torch.bool, device=idx.device)
        else:
            assert attention_mask.shape == (B, T), \
                f"attention_mask shape: {attention_mask.shape}, idx shape: {idx.shape}"
        attention_mask = attention_mask.unsqueeze(-1)
        x = self.emb(idx)
        if args.dropout > 0:
            x = self.drop0(x)
        v_first = torch.empty_like(x)
        for block in self.blocks:
            if self.training and getattr(args, "grad_cp", 0) == 1:
                # the reference calls deepspeed.checkpointing.checkpoint (:315-316); same recompute semantics
                x, v_first = checkpoint(block, x, attention_mask, v_first, use_reentrant=False)
            else:
                x, v_first = block(x, attention_mask, v_first)
        x = self.ln_out(x)
        if is_text:
            return self.head(x), None
        return None, self.audio_head(x)


# -----------------------------------------------------------------------------------------------
# stateful inference functions (B = 1), same signatures as the reference (:482-556)
# -----------------------------------------------------------------------------------------------
def _params_from_args(x_r, x_w, x_k, x_v, x_a, x_g, w0, w1, w2, a0, a1, a2, v0, v1, v2, g1, g2, k_k, k_a, r_k,
                      R_, K_, V_, O_, ln_w, ln_b, H, N):
    # R_, K_, V_, O_ are [in, out] in these functions (x @ R_), the core takes nn.Linear layout
    return core.TmixParams(x_r=x_r, x_w=x_w, x_k=x_k, x_v=x_v, x_a=x_a, x_g=x_g, w0=w0, w1=w1, w2=w2,
                           a0=a0, a1=a1, a2=a2, v0=v0, v1=v1, v2=v2, g1=g1, g2=g2, k_k=k_k, k_a=k_a,
                           r_k=r_k.view(H, N), W_r=R_.t(), W_k=K_.t(), W_v=V_.t(), W_o=O_.t(),
                           ln_w=ln_w, ln_b=ln_b, ln_eps=64e-5)


@torch.inference_mode()
def RWKV_x070_TMix_seq(layer_id: int, H: int, N: int, x, x_prev, v_first, state, x_r, x_w, x_k, x_v, x_a, x_g,
                       w0, w1, w2, a0, a1, a2, v0, v1, v2, g1, g2, k_k, k_a, r_k, R_, K_, V_, O_, ln_w, ln_b):
    """x [T,C]; x_prev [C]; state fp32 [H,N,N] value-major, updated IN PLACE like the reference's
    RWKV7_OP (:536).  Returns (out [T,C], x[-1], state, v_first)."""
    p = _params_from_args(x_r, x_w, x_k, x_v, x_a, x_g, w0, w1, w2, a0, a1, a2, v0, v1, v2, g1, g2, k_k, k_a, r_k,
                          R_, K_, V_, O_, ln_w, ln_b, H, N)
    vf = None if layer_id == 0 else v_first.unsqueeze(0)
    out, vf, xl, st = core.tmix(p, layer_id, x.unsqueeze(0), vf, shift_state=x_prev.unsqueeze(0),
                                wkv_state=state.unsqueeze(0), need_state=True)
    state.copy_(st[0])
    return out[0], xl[0], state, vf[0]


@torch.inference_mode()
def RWKV_x070_TMix_one(layer_id: int, H: int, N: int, x, x_prev, v_first, state, *weights):
    """One decode step (:482-506): x [C]."""
    vf = v_first if layer_id == 0 else v_first.unsqueeze(0)
    out, xl, state, vf = RWKV_x070_TMix_seq(layer_id, H, N, x.unsqueeze(0), x_prev, vf, state, *weights)
    return out[0], xl, state, vf[0]


@torch.inference_mode()
def RWKV_x070_CMix_seq(x, x_prev, x_k, K_, V_):
    out, xl = core.cmix(x_k, K_.t(), V_.t(), x.unsqueeze(0), shift_state=x_prev.unsqueeze(0), need_state=True)
    return out[0], xl[0]


@torch.inference_mode()
def RWKV_x070_CMix_one(x, x_prev, x_k, K_, V_):
    out, xl = RWKV_x070_CMix_seq(x.unsqueeze(0), x_prev, x_k, K_, V_)
    return out[0], xl
