"""Whole-model decode step in one persistent kernel (csrc/decode_step.cu; include/rwkvtts_wkv7.h `rwkvtts_decode_*`).

The reference generates token by token through HF `generate` over rwkvfla's RWKV7ForCausalLM
(inference/rwkv7speech_inference.py; model/llm/spark_llm.py:54-102): per token ~25 kernels per layer.  `MegaDecodeStep`
hands the model's weights and the recurrent Cache to the C-ABI once and then advances ALL layers, the final norm and the
head with one cooperative launch per token; greedy sampling (EOS masking, finished rows, padding) can run on the device
as well, so a greedy decode of n tokens is n launches and nothing else.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib

BF16 = torch.bfloat16
# enums of include/rwkvtts_wkv7.h
(DEC_B, DEC_C, DEC_H, DEC_L, DEC_V, DEC_F, DEC_DW, DEC_DA, DEC_DV, DEC_DG, DEC_NDIM) = range(11)
(DEC_EMB, DEC_LN0_W, DEC_LN0_B, DEC_LNF_W, DEC_LNF_B, DEC_HEAD, DEC_NMODEL) = range(7)
LAYER_PTRS = ("LN1_W", "LN1_B", "LN2_W", "LN2_B", "X_R", "X_W", "X_K", "X_V", "X_A", "X_G", "W_R", "W_K", "W_V", "W_O",
              "W1", "W2", "W0", "A1", "A2", "A0", "V1", "V2", "V0", "G1", "G2", "K_K", "K_A", "R_K", "GN_W", "GN_B",
              "FFN_X_K", "FFN_KEY", "FFN_VALUE", "STATE", "ATT_SHIFT", "FFN_SHIFT")
DEC_NPTR = len(LAYER_PTRS)
MAX_BATCH, MAX_EOS = 32, 8


def _ok(t: Optional[torch.Tensor], dtype=BF16) -> bool:
    return t is None or (t.is_cuda and t.dtype == dtype and t.is_contiguous() and t.data_ptr() % 16 == 0)


def _layer_tensors(block) -> dict:
    """The tensors of one rwkvfla RWKV7Block under the pointer names of the C-ABI (no copies: every one of them is
    already in the layout the kernel reads -- nn.Linear weights are [out, in])."""
    at, ff = block.attn, block.ffn
    lo = lambda m: (m.lora[0].weight, m.lora[2].weight, m.lora[2].bias)
    w1, w2, w0 = lo(at.w_lora)
    a1, a2, a0 = lo(at.a_lora)
    g1, g2, _ = lo(at.g_lora)
    v1 = v2 = v0 = None
    if hasattr(at, "v_lora"):
        v1, v2, v0 = lo(at.v_lora)
    return dict(LN1_W=block.attn_norm.weight, LN1_B=block.attn_norm.bias, LN2_W=block.ffn_norm.weight,
                LN2_B=block.ffn_norm.bias, X_R=at.x_r, X_W=at.x_w, X_K=at.x_k, X_V=at.x_v, X_A=at.x_a, X_G=at.x_g,
                W_R=at.r_proj.weight, W_K=at.k_proj.weight, W_V=at.v_proj.weight, W_O=at.o_proj.weight,
                W1=w1, W2=w2, W0=w0, A1=a1, A2=a2, A0=a0, V1=v1, V2=v2, V0=v0, G1=g1, G2=g2, K_K=at.k_k, K_A=at.k_a,
                R_K=at.r_k, GN_W=at.g_norm.weight, GN_B=at.g_norm.bias, FFN_X_K=ff.x_k, FFN_KEY=ff.key.weight,
                FFN_VALUE=ff.value.weight)


def unsupported_reason(model, batch: int) -> Optional[str]:
    """None when `model` (rwkvfla RWKV7ForCausalLM layout) can run on the one-kernel step, else why not."""
    try:
        cfg, layers = model.config, model.model.layers
        emb, head = model.model.embeddings.weight, model.lm_head
    except AttributeError as e:
        return f"not an RWKV7ForCausalLM layout ({e})"
    if not emb.is_cuda:
        return "model is not on a CUDA device"
    if batch < 1 or batch > MAX_BATCH:
        return f"batch {batch} outside 1..{MAX_BATCH}"
    if head.bias is not None:
        return "lm_head has a bias"
    C = cfg.hidden_size
    F = layers[0].ffn.key.weight.shape[0]
    if C % 64 != 0 or C > 2048 or F % C != 0 or F // C > 8:
        return f"hidden size {C} / channel-mix width {F} outside the kernel's range"
    ranks = [layers[-1].attn.w_lora.low_rank_dim, layers[-1].attn.a_lora.low_rank_dim,
             getattr(getattr(layers[-1].attn, "v_lora", None), "low_rank_dim", 32), layers[-1].attn.g_lora.low_rank_dim]
    if any(r < 32 or r % 32 != 0 for r in ranks) or sum(ranks) > 512:
        return f"LoRA ranks {ranks} are not multiples of 32 (sum <= 512)"
    for l, blk in enumerate(layers):
        if l != 0 and hasattr(blk, "pre_norm"):
            return "pre_norm on a layer other than the first"
        for n, t in _layer_tensors(blk).items():
            if not _ok(t):
                return f"layer {l} tensor {n} is not a contiguous CUDA bf16 tensor"
        if blk.attn.g_norm.weight is None:
            return "GroupNorm without affine parameters"
    for t in (emb, head.weight, model.model.norm.weight, model.model.norm.bias):
        if not _ok(t):
            return "embedding / head / final norm is not a contiguous CUDA bf16 tensor"
    return None


class MegaDecodeStep:
    """token ids [B] -> fp32 logits [B, V], every layer's recurrent state advanced in place, in ONE kernel launch.
    `cache` must hold the states of a prefill over the same batch (rwkvfla Cache: recurrent_state fp32 [B,H,64,64],
    conv_state / ffn_state [B,C]); the token-shift states are moved into contiguous bf16 buffers (and put back into the
    cache, so it stays valid for the eager path)."""

    def __init__(self, model, cache, batch: int, device):
        why = unsupported_reason(model, batch)
        if why is not None:
            raise ValueError(f"one-kernel decode step unavailable: {why}")
        self.model, self.cache, self.B = model, cache, batch
        self.device = torch.device(device)
        cfg, layers = model.config, model.model.layers
        C, L = cfg.hidden_size, len(layers)
        H = C // 64
        self.V = model.lm_head.weight.shape[0]
        F = layers[0].ffn.key.weight.shape[0]
        at = layers[-1].attn
        dims = [0] * DEC_NDIM
        dims[DEC_B], dims[DEC_C], dims[DEC_H], dims[DEC_L], dims[DEC_V], dims[DEC_F] = batch, C, H, L, self.V, F
        dims[DEC_DW], dims[DEC_DA], dims[DEC_DG] = at.w_lora.low_rank_dim, at.a_lora.low_rank_dim, at.g_lora.low_rank_dim
        dims[DEC_DV] = at.v_lora.low_rank_dim if hasattr(at, "v_lora") else 32
        self._dims = (ctypes.c_int * DEC_NDIM)(*dims)
        offs = (ctypes.c_size_t * 3)()
        lib = _lib.lib()
        nbytes = lib.rwkvtts_decode_workspace_bytes(self._dims, offs)
        if nbytes == 0:
            raise ValueError(f"one-kernel decode step unavailable: dims {dims} rejected by the library")
        self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._ws_ptr = (self.ws.data_ptr() + 255) // 256 * 256
        base = self._ws_ptr - self.ws.data_ptr()
        view = lambda off, n, dt: self.ws[base + off: base + off + n].view(dt)
        self.logits = view(offs[0], batch * self.V * 4, torch.float32).view(batch, self.V)
        self.next_tok = view(offs[1], MAX_BATCH * 8, torch.int64)[:batch]
        self.done = view(offs[2], MAX_BATCH * 4, torch.int32)[:batch]
        # recurrent state: the WKV state is used in place, the token-shift states as contiguous bf16 buffers
        self._keep: List[torch.Tensor] = []
        lp = (ctypes.c_void_p * (L * DEC_NPTR))()
        for l, blk in enumerate(layers):
            st = cache.states[l]
            rs = st["recurrent_state"]
            if not (_ok(rs, torch.float32) and tuple(rs.shape) == (batch, H, 64, 64)):
                rs = rs.to(torch.float32).contiguous().clone()
                st["recurrent_state"] = rs
            for k in ("conv_state", "ffn_state"):
                s = st[k]
                if not (_ok(s) and tuple(s.shape) == (batch, C)):
                    s = s.to(BF16).reshape(batch, C).contiguous().clone()
                    st[k] = s
            t = _layer_tensors(blk)
            t.update(STATE=rs, ATT_SHIFT=st["conv_state"], FFN_SHIFT=st["ffn_state"])
            self._keep.extend(v for v in t.values() if v is not None)
            for i, n in enumerate(LAYER_PTRS):
                lp[l * DEC_NPTR + i] = None if t[n] is None else t[n].data_ptr()
        pre = getattr(layers[0], "pre_norm", None)
        mp = (ctypes.c_void_p * DEC_NMODEL)()
        ptr = lambda t: None if t is None else t.data_ptr()
        mp[DEC_EMB] = model.model.embeddings.weight.data_ptr()
        mp[DEC_LN0_W], mp[DEC_LN0_B] = (ptr(pre.weight), ptr(pre.bias)) if pre is not None else (None, None)
        mp[DEC_LNF_W], mp[DEC_LNF_B] = ptr(model.model.norm.weight), ptr(model.model.norm.bias)
        mp[DEC_HEAD] = model.lm_head.weight.data_ptr()
        eps = (ctypes.c_float * 2)(float(cfg.norm_eps), float(at.g_norm.eps))
        with torch.cuda.device(self.device):
            rc = lib.rwkvtts_decode_init(self._dims, eps, mp, lp, ctypes.c_void_p(self._ws_ptr), nbytes, self._stream())
        _lib.check(rc, "rwkvtts_decode_init")
        self._no_eos = (ctypes.c_longlong * 1)(0)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if self._ws_ptr:
            _lib.lib().rwkvtts_decode_release(ctypes.c_void_p(self._ws_ptr))
            self._ws_ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, nxt: torch.Tensor) -> torch.Tensor:
        """One step on the given token ids; returns the workspace's logits buffer (overwritten by the next step)."""
        tok = nxt.reshape(-1).to(torch.int64).contiguous()
        assert tok.numel() == self.B and tok.is_cuda
        with torch.cuda.device(self.device):
            rc = _lib.lib().rwkvtts_decode_step(ctypes.c_void_p(self._ws_ptr), ctypes.c_void_p(tok.data_ptr()), None, 0, 0,
                                                self._no_eos, 0, 0, self._stream())
        _lib.check(rc, "rwkvtts_decode_step")
        self.cache._seen_tokens += 1
        return self.logits

    def phase_times(self, nxt: torch.Tensor) -> torch.Tensor:
        """One step with the kernel's time line: float64 [7 L + 1] microseconds spent between consecutive grid barriers
        (per layer: ln1, projections, wkv, output projection, ln2, channel-mix key, channel-mix value; then the final norm)."""
        L = len(self.model.model.layers)
        n = 8 * L + 4
        stamps = torch.zeros(n * 513 + 128, dtype=torch.int64, device=self.device)
        tok = nxt.reshape(-1).to(torch.int64).contiguous()
        with torch.cuda.device(self.device):
            rc = _lib.lib().rwkvtts_decode_step_profile(ctypes.c_void_p(self._ws_ptr), ctypes.c_void_p(tok.data_ptr()),
                                                        ctypes.c_void_p(stamps.data_ptr()), self._stream())
        _lib.check(rc, "rwkvtts_decode_step_profile")
        self.cache._seen_tokens += 1
        self.last_arrive = stamps[n: n + 256 * n].view(n, 256)[1: 7 * L + 2]        # [barrier, CTA] ns (0 = no such CTA)
        self.last_release = stamps[257 * n: 257 * n + 256 * n].view(n, 256)[1: 7 * L + 2]
        self.last_fine = stamps[n * 513:]                  # cycle stamps inside layer 1's phases (csrc/decode_step.cu)
        t = stamps[: 7 * L + 2].double()
        return (t[1:] - t[:-1]) / 1e3

    def greedy(self, first: torch.Tensor, steps: int, eos: Sequence[int] = (), pad: int = 0, min_new_tokens: int = 0,
               step0: int = 0, done: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`steps` greedy tokens after `first` [B] (the token the caller sampled from the prefill logits), sampled on the
        device: returns int64 [steps, B].  Step s (counted from `step0`) masks the EOS ids while s < min_new_tokens; rows
        that are finished (`done`, updated in place when given) emit `pad`."""
        assert len(eos) <= MAX_EOS
        out = torch.empty(steps, self.B, dtype=torch.int64, device=self.device)
        if steps == 0:
            return out
        eos_arr = (ctypes.c_longlong * max(len(eos), 1))(*eos) if eos else self._no_eos
        self.done.copy_(done.to(torch.int32) if done is not None else torch.zeros_like(self.done))
        tok = first.reshape(-1).to(torch.int64).contiguous()
        lib, ws = _lib.lib(), ctypes.c_void_p(self._ws_ptr)
        with torch.cuda.device(self.device):
            st = self._stream()
            for s in range(steps):
                rc = lib.rwkvtts_decode_step(ws, ctypes.c_void_p(tok.data_ptr()) if s == 0 else None,
                                             ctypes.c_void_p(out[s].data_ptr()), 1, int(step0 + s < min_new_tokens), eos_arr,
                                             len(eos), int(pad), st)
                if rc != 0:
                    _lib.check(rc, "rwkvtts_decode_step")
        self.cache._seen_tokens += steps
        if done is not None:
            done.copy_(self.done.to(done.dtype))
        return out
