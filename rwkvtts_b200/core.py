"""RWKV-7 (x070) time-mix / channel-mix around the WKV-7 op, shared by the two module stacks.

One functional restatement of

  RWKV_Tmix_x070.forward      /root/reference/model/llm/rwkv_s2s_single_ffn.py:158-196   (training)
  RWKV_x070_TMix_one / _seq   /root/reference/model/llm/rwkv_s2s_single_ffn.py:482-540   (stateful inference)
  RWKV_CMix_x070.forward      :223-230,  RWKV_x070_CMix_one / _seq  :545-556
  RWKV7Attention.forward      rwkv-fla (third party): same math with
                              w = -0.6065*sigmoid(lora) == log(exp(-exp(w_pre))), w_pre = -softplus(-lora) - 0.5

driven by a plain namespace of tensors (`TmixParams`), so the BlinkDL-named modules (x070.py) and the
rwkvfla-named modules (rwkvfla/) are two thin parameter adapters over the same code.

The recurrence itself always runs in the CUDA library (ops.py): WindBackstepping for training,
the snapshot-free tcgen05 forward for no-grad prefill, the stateful scan for any-T / decode steps.
On CUDA bf16 activations the elementwise chain between the GEMMs runs in three fused kernels (fused.py:
token-shift + lerps; decay / gates / kk / WKV operands; GroupNorm + bonus + gate), forward and backward; the
GEMMs go through cuBLAS.  The ATen formulation below them is what the CPU oracle tests and fp32 callers use
(`FUSED = False` or env RWKVTTS_FUSED=0 selects it on the GPU too).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

import os

from . import fused, ops

HEAD = ops.HEAD_SIZE
FUSED = os.environ.get("RWKVTTS_FUSED", "1") != "0"
DECODE_BATCHED_GEMM = os.environ.get("RWKVTTS_DECODE_BMM", "1") != "0"
# measurement hook (bench.py R-GPU-model leg): an autograd.Function with WindBackstepping's signature that replaces the
# training op (the bench binds the compiled reference kernels there); None = this library
WKV_TRAIN_OP = None
# "exact" inference: the reference's decode step operation for operation -- the ATen elementwise chain (bf16
# intermediates, as RWKV_Tmix_x070.forward_batch, rwkv_asr_cuda_whisper.py:181-215) around the stateful kernel in the
# reference's summation order (csrc/wkv7_step_exact.cu), prompt included.  Greedy token ids are then bit-identical to
# the reference loop's (north_star); the default path trades that for speed (fused kernels keep fp32 intermediates).
EXACT = False


class exact_mode:
    """with core.exact_mode(): ...   (what generate(..., exact=True) enters)"""

    def __enter__(self):
        global FUSED, EXACT
        from . import _lib
        self.prev = (FUSED, EXACT, _lib.lib().rwkvtts_get_step_mode())
        FUSED, EXACT = False, True
        _lib.lib().rwkvtts_set_step_mode(1)
        return self

    def __exit__(self, *exc):
        global FUSED, EXACT
        from . import _lib
        FUSED, EXACT = self.prev[0], self.prev[1]
        _lib.lib().rwkvtts_set_step_mode(self.prev[2])
        return False


@dataclass
class TmixParams:
    """Time-mix tensors.  Linear weights are [out, in] (nn.Linear); LoRA down/up are [in, low] / [low, out]
    (BlinkDL layout; the rwkvfla adapter passes transposed views).  x_* / k_* / *0 broadcast over [.., C]."""
    x_r: torch.Tensor
    x_w: torch.Tensor
    x_k: torch.Tensor
    x_v: torch.Tensor
    x_a: torch.Tensor
    x_g: torch.Tensor
    w0: torch.Tensor
    w1: torch.Tensor
    w2: torch.Tensor
    a0: torch.Tensor
    a1: torch.Tensor
    a2: torch.Tensor
    v0: Optional[torch.Tensor]
    v1: Optional[torch.Tensor]
    v2: Optional[torch.Tensor]
    g1: torch.Tensor
    g2: torch.Tensor
    k_k: torch.Tensor
    k_a: torch.Tensor
    r_k: torch.Tensor          # [H, 64]
    W_r: torch.Tensor
    W_k: torch.Tensor
    W_v: torch.Tensor
    W_o: torch.Tensor
    ln_w: torch.Tensor
    ln_b: torch.Tensor
    ln_eps: float = 64e-5      # 1e-5 * head_size_divisor**2 (:145) == head_dim * norm_eps (rwkvfla)


_MASK_CACHE: list = [None]     # (weakref to the [B, T_all] mask, key, the [B, T, 1] result)


def mask3(attention_mask: Optional[torch.Tensor], T: int, dtype: torch.dtype) -> Optional[torch.Tensor]:
    """The last T columns of a 0/1 `attention_mask` [B, T_all] as a [B, T, 1] tensor of the activations' dtype.  Every
    layer of a model asks for the same one: the conversion runs once per forward (one entry, keyed on the mask object,
    its version counter and the autograd / inference mode) instead of twice per layer.  Not cached while a CUDA graph is
    being captured (the result would live in the graph's private pool)."""
    if attention_mask is None:
        return None
    convert = lambda: attention_mask.narrow(1, attention_mask.size(1) - T, T).unsqueeze(-1).to(dtype)
    if attention_mask.is_cuda and torch.cuda.is_current_stream_capturing():
        return convert()
    import weakref
    try:
        ver = attention_mask._version
    except RuntimeError:                        # an inference tensor tracks no version: nothing to key the entry on
        return convert()
    key = (ver, T, dtype, torch.is_inference_mode_enabled(), attention_mask.device)
    hit = _MASK_CACHE[0]
    if hit is not None and hit[0]() is attention_mask and hit[1] == key:
        return hit[2]
    am = convert()
    _MASK_CACHE[0] = (weakref.ref(attention_mask), key, am)
    return am


def token_shift(x: torch.Tensor, prev: Optional[torch.Tensor]) -> torch.Tensor:
    """time_shift(x) - x with time_shift = ZeroPad2d((0,0,1,-1)) (:162); `prev` [B,C] is the last
    token of the previous call (:511), zeros if None."""
    if prev is None:
        shifted = F.pad(x, (0, 0, 1, -1))
    else:
        shifted = torch.cat((prev.unsqueeze(1).to(x.dtype), x[:, :-1]), dim=1)
    return shifted - x


def _wkv(r, w, k, v, a, b, state, need_state, inplace_state=False, plan=None):
    """Dispatch of the recurrence.  r,w,k,v,a,b: bf16 [B,T,C] contiguous; state fp32 [B,H,64,64] or None.
    Returns y [B,T,C] and the final state (None unless need_state).  plan (ops.VarlenPlan): the batch is ONE packed
    sequence [1, T_total, C] whose recurrences restart at the plan's cu_seqlens (chunked tcgen05 kernels, any lengths)."""
    B, T, C = r.shape
    H = C // HEAD
    if plan is not None:
        assert B == 1 and state is None and not need_state, "a packed batch carries no recurrent state in or out"
        sh = lambda t: t.view(1, T, H, HEAD)
        return ops.wkv7_varlen(sh(w), sh(r), sh(k), sh(v), sh(a), sh(b), plan).view(1, T, C), None
    grad = torch.is_grad_enabled() and any(t.requires_grad for t in (r, w, k, v, a, b))
    if T % ops.CHUNK_LEN == 0 and not (EXACT and not grad):
        sh = lambda t: t.view(B, T, H, HEAD)
        if state is None and not need_state:
            if WKV_TRAIN_OP is not None:
                return WKV_TRAIN_OP.apply(sh(w), sh(r), sh(k), sh(v), sh(a), sh(b)).view(B, T, C), None
            return ops.RUN_CUDA_RWKV7g(r, w, k, v, a, b), None                     # :191 (training / no_grad)
        y, sT = ops.wkv7_with_state(sh(w), sh(r), sh(k), sh(v), sh(a), sh(b), state)
        return y.view(B, T, C), sT
    if grad:
        # ragged length under autograd: right-pad to a multiple of 16 with k=v=a=b=0 tokens (they leave
        # every real output untouched; the state after them is not used)
        assert not need_state, "final state with T % 16 != 0 is only available without autograd"
        pad = (-T) % ops.CHUNK_LEN
        pz = lambda t: F.pad(t, (0, 0, 0, pad))
        sh = lambda t: pz(t).view(B, T + pad, H, HEAD)
        y, _ = ops.wkv7_with_state(sh(w), sh(r), sh(k), sh(v), sh(a), sh(b), state)
        return y.view(B, T + pad, C)[:, :T], None
    # any T, no autograd: the stateful op (decode step for T == 1), state updated in place (:536)
    if state is None:
        st = torch.zeros(B, H, HEAD, HEAD, dtype=torch.float32, device=r.device)
    else:
        st = state if inplace_state else state.clone()
    y = ops.RWKV7_BATCH_OP(st, r, w, k, v, a, b)
    return y, st


def tmix(p: TmixParams, layer_id: int, x: torch.Tensor, v_first: Optional[torch.Tensor],
         mask: Optional[torch.Tensor] = None, mask_rwk: bool = True,
         shift_state: Optional[torch.Tensor] = None, wkv_state: Optional[torch.Tensor] = None,
         need_state: bool = False, inplace_state: bool = False, mask_kk: bool = True, plan=None):
    """x [B,T,C] (bf16 on the GPU).  mask [B,T,1] of 0/1 or None.  Returns
    (out [B,T,C], v_first, new_shift_state [B,C] | None, new_wkv_state | None).  With `inplace_state` the
    stateful (decode) path advances `wkv_state` in place like the reference's RWKV7_OP (:536).  `mask_kk=False` is the
    masking of the reference's inference-only `forward_batch` (rwkv_asr_cuda_whisper.py:184, :209: x and v only)."""
    B, T, C = x.shape
    H = C // HEAD
    if FUSED and fused.usable(x) and (mask is None or mask_kk):
        return _tmix_fused(p, layer_id, x, v_first, mask, mask_rwk, shift_state, wkv_state, need_state, inplace_state, plan)
    if plan is not None:
        raise NotImplementedError("packed (cu_seqlens) input needs the fused CUDA path; unpack it per sequence otherwise")
    if mask is not None:
        x = x * mask                                                            # :160
    xx = token_shift(x, shift_state)                                            # :162
    # EXACT: two roundings per lerp like the reference's `x + xx * self.x_r` (:164-169); addcmul rounds once
    lerp = (lambda m: x + xx * m) if EXACT else (lambda m: torch.addcmul(x, xx, m))
    xr, xw, xk, xv, xa, xg = (lerp(m) for m in (p.x_r, p.x_w, p.x_k, p.x_v, p.x_a, p.x_g))
    r = F.linear(xr, p.W_r)
    w = -F.softplus(-(p.w0 + torch.tanh(xw @ p.w1) @ p.w2)) - 0.5               # :172
    k = F.linear(xk, p.W_k)
    v = F.linear(xv, p.W_v)
    if mask is not None and mask_rwk:
        r, w, k, v = r * mask, w * mask, k * mask, v * mask                     # :175-178
    if layer_id == 0:
        v_first = v                                                             # :180
    else:
        v = v + (v_first - v) * torch.sigmoid(p.v0 + (xv @ p.v1) @ p.v2)        # :182
    a = torch.sigmoid(p.a0 + (xa @ p.a1) @ p.a2)                                # :183
    g = torch.sigmoid(xg @ p.g1) @ p.g2                                         # :184
    kk = F.normalize((k * p.k_k).view(B, T, H, HEAD), dim=-1, p=2.0).view(B, T, C)   # :186-187
    if mask is not None:
        if mask_kk:
            kk = kk * mask                                                      # :188
        v = v * mask                                                            # :190
    k = k * (1 + (a - 1) * p.k_a)                                               # :189
    c = lambda t: t.to(torch.bfloat16).contiguous()
    y, new_state = _wkv(c(r), c(w), c(k), c(v), c(-kk), c(kk * a), wkv_state, need_state,
                        inplace_state)                                                    # :191
    y = F.group_norm(y.reshape(B * T, C).to(r.dtype), H, p.ln_w, p.ln_b, p.ln_eps).view(B, T, C)   # :192
    y = y + ((r.view(B, T, H, HEAD) * k.view(B, T, H, HEAD) * p.r_k).sum(dim=-1, keepdim=True)
             * v.view(B, T, H, HEAD)).view(B, T, C)                              # :194
    out = F.linear(y * g, p.W_o)                                                 # :195
    return out, v_first, (x[:, -1] if need_state else None), new_state


def _tmix_fused(p: TmixParams, layer_id: int, x, v_first, mask, mask_rwk, shift_state, wkv_state, need_state,
                inplace_state, plan=None):
    """tmix() with the elementwise chain in the fused kernels (same reference lines, same results up to bf16
    rounding of intermediates the fused kernels keep in fp32)."""
    if DECODE_BATCHED_GEMM and not torch.is_grad_enabled() and x.shape[1] == 1 and layer_id != 0 and plan is None:
        return _tmix_decode(p, layer_id, x, v_first, mask, mask_rwk, shift_state, wkv_state, need_state, inplace_state)
    xr, xw, xk, xv, xa, xg = fused.shift_mix(x, (p.x_r, p.x_w, p.x_k, p.x_v, p.x_a, p.x_g), mask, shift_state,
                                             seq_first=None if plan is None else plan.first)                # :160-169
    r = F.linear(xr, p.W_r)
    k = F.linear(xk, p.W_k)
    v = F.linear(xv, p.W_v)
    w_lo = torch.tanh(xw @ p.w1) @ p.w2
    a_lo = (xa @ p.a1) @ p.a2
    g = torch.sigmoid(xg @ p.g1) @ p.g2                                         # :184
    if mask is not None and mask_rwk:
        r = r * mask                                                            # :175
    if layer_id == 0:
        w, k2, v2, a_op, b_op = fused.prep(k, v, w_lo, a_lo, None, None, p.w0, p.a0, None, p.k_k, p.k_a, mask, mask_rwk)
        v_first = v2 if (mask is None or mask_rwk) else v                       # :180 (rwkvfla: the unmasked v)
    else:
        w, k2, v2, a_op, b_op = fused.prep(k, v, w_lo, a_lo, (xv @ p.v1) @ p.v2, v_first, p.w0, p.a0, p.v0, p.k_k,
                                           p.k_a, mask, mask_rwk)               # :172-190
    y, new_state = _wkv(r.contiguous(), w, k2, v2.contiguous(), a_op, b_op, wkv_state, need_state, inplace_state,
                        plan=plan)                                                                          # :191
    o = fused.out(y, r, k2, v2, g, p.r_k, p.ln_w, p.ln_b, p.ln_eps)             # :192-195
    shift_out = None
    if need_state:
        shift_out = x[:, -1] if mask is None else x[:, -1] * mask[:, -1]
    return F.linear(o, p.W_o), v_first, shift_out, new_state


def _tmix_decode(p: TmixParams, layer_id: int, x, v_first, mask, mask_rwk, shift_state, wkv_state, need_state, inplace_state):
    """One decode step (T = 1, no autograd, layer > 0) of _tmix_fused with the 11 skinny GEMMs between the fused
    kernels batched into 3 (+ output projection): at 32 rows every GEMM is a ~5 us launch that reads 0.1-2 MB of weights, so the step is bound
    by their count.  Weights are stacked once and cached until modified in place."""
    B, T, C = x.shape
    # order (r, k, v | w, a, g): two contiguous groups of projection inputs
    inplace_shift = (need_state and inplace_state and shift_state is not None and shift_state.dtype == torch.bfloat16
                     and shift_state.is_contiguous())
    X = fused.shift_mix_stacked(x, (p.x_r, p.x_k, p.x_v, p.x_w, p.x_a, p.x_g), mask, shift_state,
                                update_prev=inplace_shift)                      # [6,B,1,C]
    X = X.view(6, B, C)
    W3 = fused.cached(p.W_r, (p.W_r, p.W_k, p.W_v), "rkv", lambda: torch.stack((p.W_r.t(), p.W_k.t(), p.W_v.t())).contiguous())
    D = max(p.v1.shape[1], p.w1.shape[1], p.a1.shape[1], p.g1.shape[1])
    padc = lambda m: F.pad(m, (0, D - m.shape[1]))
    padr = lambda m: F.pad(m, (0, 0, 0, D - m.shape[0]))
    # the four LoRAs read X[2:6] = (xv, xw, xa, xg): one batched down- and one batched up-projection, ranks zero-padded
    L1 = fused.cached(p.w1, (p.v1, p.w1, p.a1, p.g1), "lora1",
                      lambda: torch.stack((padc(p.v1), padc(p.w1), padc(p.a1), padc(p.g1))).contiguous())
    L2 = fused.cached(p.w2, (p.v2, p.w2, p.a2, p.g2), "lora2",
                      lambda: torch.stack((padr(p.v2), padr(p.w2), padr(p.a2), padr(p.g2))).contiguous())
    rkv = torch.bmm(X[0:3], W3)                                                  # [3,B,C]
    h = torch.bmm(X[2:6], L1)                                                    # [4,B,D]
    h[1].tanh_()
    if p.g1.shape[1] == D:
        h[3].sigmoid_()
    else:                                                                        # sigmoid(0) != 0: keep the padding at zero
        h[3, :, :p.g1.shape[1]].sigmoid_()
    lo = torch.bmm(h, L2)                                                        # [4,B,C]: v_lo, w_lo, a_lo, g
    r, k, v = (rkv[i].view(B, 1, C) for i in range(3))
    v_lo, w_lo, a_lo, g = (lo[i].view(B, 1, C) for i in range(4))
    if mask is not None and mask_rwk:
        r = r * mask
    w, k2, v2, a_op, b_op = fused.prep(k, v, w_lo, a_lo, v_lo, v_first, p.w0, p.a0, p.v0, p.k_k, p.k_a, mask, mask_rwk)
    y, new_state = _wkv(r.contiguous(), w, k2, v2.contiguous(), a_op, b_op, wkv_state, need_state, inplace_state)
    o = fused.out(y, r, k2, v2, g, p.r_k, p.ln_w, p.ln_b, p.ln_eps)
    shift_out = None
    if need_state:
        shift_out = shift_state if inplace_shift else (x[:, -1] if mask is None else x[:, -1] * mask[:, -1])
    return F.linear(o, p.W_o), v_first, shift_out, new_state


def cmix(x_k: torch.Tensor, W_key: torch.Tensor, W_value: torch.Tensor, x: torch.Tensor,
         mask: Optional[torch.Tensor] = None, shift_state: Optional[torch.Tensor] = None,
         need_state: bool = False, inplace_state: bool = False, plan=None):
    """RWKV_CMix_x070.forward (:223-230) / RWKV_x070_CMix_seq (:551-556).  plan: packed batch (see _wkv)."""
    if plan is not None:
        if not (FUSED and fused.usable(x)):
            raise NotImplementedError("packed (cu_seqlens) input needs the fused CUDA path")
        (xk,) = fused.shift_mix(x, (x_k,), mask, None, seq_first=plan.first)
        k = F.linear(xk, W_key)
        k = fused.sqrelu(k) if k.numel() % 8 == 0 else torch.relu(k) ** 2
        return F.linear(k, W_value), None
    if (FUSED and fused.usable(x) and not torch.is_grad_enabled() and x.shape[1] == 1 and need_state and inplace_state
            and shift_state is not None and shift_state.dtype == torch.bfloat16 and shift_state.is_contiguous()):
        # decode step: the kernel also writes the new shift state into the caller's buffer
        xk = fused.shift_mix_stacked(x, (x_k,), mask, shift_state, update_prev=True)[0]
        k = F.linear(xk, W_key)
        k = fused.sqrelu(k) if k.numel() % 8 == 0 else torch.relu(k) ** 2
        return F.linear(k, W_value), shift_state
    if FUSED and fused.usable(x):
        (xk,) = fused.shift_mix(x, (x_k,), mask, shift_state)                   # :224-226
        k = F.linear(xk, W_key)
        k = fused.sqrelu(k) if k.numel() % 8 == 0 else torch.relu(k) ** 2        # :228
        last = None
        if need_state:
            last = x[:, -1] if mask is None else x[:, -1] * mask[:, -1]
        return F.linear(k, W_value), last
    if mask is not None:
        x = x * mask
    xx = token_shift(x, shift_state)
    k = torch.relu(F.linear((x + xx * x_k) if EXACT else torch.addcmul(x, xx, x_k), W_key)) ** 2
    return F.linear(k, W_value), (x[:, -1] if need_state else None)
