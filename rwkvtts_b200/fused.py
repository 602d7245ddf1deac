"""Fused elementwise kernels of the RWKV-7 time-mix around the WKV-7 op (autograd functions over the C ABI).

They replace the ATen elementwise chain of RWKV_Tmix_x070.forward between its GEMMs
(/root/reference/model/llm/rwkv_s2s_single_ffn.py:160-195) and the token-shift lerp of RWKV_CMix_x070.forward (:226):

  shift_mix(x, mixes, mask, prev)                                  -> n tensors  x + (shift(x) - x) * mix_i      (:160-169, :226)
  prep(k, v, w_lo, a_lo, v_lo, v_first, w0, a0, v0, k_k, k_a, mask) -> w, k', v', a_op, b_op                      (:172-190)
  out(y, r, k', v', g, r_k, ln_w, ln_b, eps)                        -> (GroupNorm(y) + bonus) * g                 (:192-195)

Kernels: rwkvtts_b200/csrc/tmix_fused.cu.  CUDA bf16 only; there is no CPU fallback (core.py keeps the ATen
formulation for everything these kernels do not cover: CPU tensors in the oracle tests, fp32 activations).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Sequence

import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

BF16 = torch.bfloat16


MAX_C_FORWARD, MAX_C_BACKWARD = 4096, 2048     # the adjoint kernels hold C/8 threads per row in CTAs of 256 threads


def usable(x: torch.Tensor) -> bool:
    """The fused kernels cover CUDA bf16 activations with C a multiple of 64 (one head = 8 lanes x 8 channels), C <= 4096
    without autograd and C <= 2048 with it (beyond that the ATen chain runs: ADVICE round 1 -- the adjoint kernels cannot
    launch a row wider than their 256-thread CTAs)."""
    if not (x.is_cuda and x.dtype == BF16 and x.shape[-1] % 64 == 0):
        return False
    return x.shape[-1] <= (MAX_C_BACKWARD if torch.is_grad_enabled() else MAX_C_FORWARD)


# fp32 copies of per-channel parameters (~14 tiny conversion kernels per layer and call otherwise): cached per parameter OBJECT (weak references: a freed parameter's address can be reused by
# another tensor) and version counter.
_PCACHE: dict = {}       # id(parameter) -> (weakref to it, version key, fp32 copy)
_EPOCH = [0]


def invalidate_param_cache() -> None:
    """Called by whatever rewrites parameter storage behind autograd's back: the Adam kernel writes through raw
    pointers and the engine's all-gather lands in the flat buffer the parameters alias, so no parameter's `_version`
    moves.  Without this a generate() after a training step would run on the cached copies of the old weights."""
    _EPOCH[0] += 1
    _PCACHE.clear()


def _cached(p: torch.Tensor, ver, make):
    k = id(p)
    ver = (_EPOCH[0], ver)
    hit = _PCACHE.get(k)
    if hit is not None and hit[0]() is p and hit[1] == ver and hit[2].device == p.device:
        return hit[2]
    val = make()
    _PCACHE[k] = (weakref.ref(p, lambda _r, k=k: _PCACHE.pop(k, None)), ver, val)
    return val


def _f32(p: Optional[torch.Tensor], C: int) -> Optional[torch.Tensor]:
    if p is None:
        return None
    make = lambda: p.detach().reshape(-1).to(torch.float32).contiguous()
    # the copy is detached in either mode (the kernels' adjoints return the parameter gradients themselves), so it is
    # shared by every call until the parameter changes: in-place updates move `_version`, the engine's raw-pointer Adam
    # calls invalidate_param_cache() -- ~330 tiny conversion kernels per training step otherwise
    return _cached(p, ("f32", p._version), make)


def _stack32(mixes: Sequence[torch.Tensor]) -> torch.Tensor:
    """[n, C] stack of the lerp coefficients (fp32 under no_grad, cached on the first coefficient like _f32)."""
    if torch.is_grad_enabled():
        return torch.stack([p.reshape(-1) for p in mixes])
    ver = ("stack",) + tuple((id(p), p._version) for p in mixes)
    return _cached(mixes[0], ver, lambda: torch.stack([p.detach().reshape(-1) for p in mixes]).to(torch.float32).contiguous())


def _mask2d(mask: Optional[torch.Tensor], B: int, T: int) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    return mask.detach().reshape(B, T).to(BF16).contiguous()


def _scratch(B, T, C, n, dev):
    return torch.empty(_lib.lib().rwkvtts_tmix_scratch_floats(B, T, C, n), dtype=torch.float32, device=dev)


def _param_grads(dparams: torch.Tensor, metas):
    """Rows of the fp32 [n, C] parameter-gradient block as tensors of the parameters' dtypes / shapes; metas[i] = (dtype,
    shape) or None.  One conversion kernel for the whole block when the dtypes agree (the usual case: all bf16), instead
    of one per parameter (~190 tiny launches per training step)."""
    dts = {m[0] for m in metas if m is not None}
    conv = dparams.to(next(iter(dts))) if len(dts) == 1 else None
    return [None if m is None else (conv[i] if conv is not None else dparams[i].to(m[0])).reshape(m[1])
            for i, m in enumerate(metas)]


def _ptr_array(ts: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


class _ShiftMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mixes, mask, prev, first=None):
        # x [B,T,C] bf16; mixes [n,C] (any float dtype); mask [B,T] bf16 | None; prev [B,C] bf16 | None;
        # first [B*T] uint8 | None: 1 on the first token of every sequence of a packed batch
        _need_cuda(x, mixes, mask, prev, first)
        B, T, C = x.shape
        n = mixes.shape[0]
        x = x.contiguous()
        mix32 = mixes.detach().to(torch.float32).contiguous()
        outs = [torch.empty_like(x) for _ in range(n)]
        with torch.cuda.device(x.device):
            rc = _lib.lib().rwkvtts_tmix_shift_mix_forward_varlen(B, T, C, n, _ptr(x), _ptr(mask), _ptr(prev), _ptr(mix32),
                                                                  _ptr_array(outs), None, _ptr(first), _stream())
        _lib.check(rc, "rwkvtts_tmix_shift_mix_forward")
        ctx.save_for_backward(x, mix32, mask, prev, first)
        ctx.mix_dtype = mixes.dtype
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        x, mix32, mask, prev, first = ctx.saved_tensors
        B, T, C = x.shape
        n = mix32.shape[0]
        douts = [torch.zeros_like(x) if d is None else d.contiguous() for d in douts]
        dx = torch.empty_like(x)
        dmix = torch.empty_like(mix32)
        scratch = _scratch(B, T, C, n, x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().rwkvtts_tmix_shift_mix_backward_varlen(B, T, C, n, _ptr(x), _ptr(mask), _ptr(prev), _ptr(mix32),
                                                                   _ptr_array(douts), _ptr(dx), _ptr(dmix), _ptr(scratch),
                                                                   _ptr(first), _stream())
        _lib.check(rc, "rwkvtts_tmix_shift_mix_backward")
        return dx, dmix.to(ctx.mix_dtype), None, None, None


def shift_mix(x: torch.Tensor, mixes: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None,
              prev: Optional[torch.Tensor] = None, seq_first: Optional[torch.Tensor] = None):
    """[x + (shift(x*mask) - x*mask) * m for m in mixes]; mixes broadcastable to [C] (n = 1 or 6).  seq_first (uint8
    [B*T], packed batches): the shift restarts with zeros at every flagged token."""
    B, T, C = x.shape
    m = _stack32(mixes)
    prev_ = None if prev is None else prev.detach().to(BF16).contiguous()
    return _ShiftMix.apply(x, m, _mask2d(mask, B, T), prev_, seq_first)


@torch.no_grad()
def shift_mix_stacked(x: torch.Tensor, mixes: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None,
                      prev: Optional[torch.Tensor] = None, update_prev: bool = False) -> torch.Tensor:
    """shift_mix without autograd, outputs as one [n, B, T, C] buffer (inputs of batched projections on the decode path).
    update_prev (T == 1 only): `prev` [B,C] bf16 contiguous is overwritten with the new shift state by the same kernel."""
    _need_cuda(x, mask, prev)
    B, T, C = x.shape
    n = len(mixes)
    x = x.contiguous()
    mix32 = _stack32(mixes)
    out = torch.empty(n, B, T, C, dtype=x.dtype, device=x.device)
    prev_ = None if prev is None else prev.to(BF16).contiguous()
    if update_prev:
        assert T == 1 and prev_ is prev, "in-place shift state: T == 1 and a contiguous bf16 [B,C] buffer"
    with torch.cuda.device(x.device):
        rc = _lib.lib().rwkvtts_tmix_shift_mix_forward(B, T, C, n, _ptr(x), _ptr(_mask2d(mask, B, T)), _ptr(prev_), _ptr(mix32),
                                                       _ptr_array([out[i] for i in range(n)]),
                                                       _ptr(prev_) if update_prev else None, _stream())
    _lib.check(rc, "rwkvtts_tmix_shift_mix_forward")
    return out


def cached(anchor: torch.Tensor, tensors: Sequence[torch.Tensor], tag: str, make):
    """make() cached on `anchor` until any of `tensors` is modified in place (no_grad paths only)."""
    base = lambda t: t._base if t._base is not None else t        # .t() views are rebuilt per call: key on their parameter
    return _cached(base(anchor), (tag,) + tuple((id(base(t)), t._version) for t in tensors), make)


class _Prep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, k, v, w_lo, a_lo, v_lo, v_first, w0, a0, v0, k_k, k_a, mask, mask_rwk):
        _need_cuda(k, v, w_lo, a_lo, v_lo, v_first)
        B, T, C = k.shape
        k, v, w_lo, a_lo = (t.contiguous() for t in (k, v, w_lo, a_lo))
        has_v = v_lo is not None
        if has_v:
            v_lo, v_first = v_lo.contiguous(), v_first.contiguous()
        p32 = [_f32(p, C) for p in (w0, a0, v0 if has_v else None, k_k, k_a)]
        w, k2, a_op, b_op = (torch.empty_like(k) for _ in range(4))
        need_v2 = has_v or mask is not None
        v2 = torch.empty_like(v) if need_v2 else None
        with torch.cuda.device(k.device):
            rc = _lib.lib().rwkvtts_tmix_prep_forward(B, T, C, _ptr(k), _ptr(v), _ptr(w_lo), _ptr(a_lo), _ptr(v_lo),
                                                      _ptr(v_first), _ptr(mask), *[_ptr(p) for p in p32], int(mask_rwk), _ptr(w),
                                                      _ptr(k2), _ptr(v2), _ptr(a_op), _ptr(b_op), _stream())
        _lib.check(rc, "rwkvtts_tmix_prep_forward")
        ctx.save_for_backward(k, v, w_lo, a_lo, v_lo, v_first, mask, *[p for p in p32 if p is not None])
        ctx.has_v, ctx.need_v2, ctx.mask_rwk = has_v, need_v2, int(mask_rwk)
        ctx.dtypes = [p.dtype for p in (w0, a0, k_k, k_a)] + [v0.dtype if has_v else None]
        ctx.shapes = [p.shape for p in (w0, a0, k_k, k_a)] + [v0.shape if has_v else None]
        return w, k2, (v2 if need_v2 else v), a_op, b_op

    @staticmethod
    def backward(ctx, dw, dk2, dv2, da_op, db_op):
        saved = ctx.saved_tensors
        k, v, w_lo, a_lo, v_lo, v_first, mask = saved[:7]
        ps = list(saved[7:])
        if ctx.has_v:
            w0, a0, v0, k_k, k_a = ps
        else:
            (w0, a0, k_k, k_a), v0 = ps, None
        B, T, C = k.shape
        z = lambda d: torch.zeros_like(k) if d is None else d.contiguous()
        dw, dk2, da_op, db_op = z(dw), z(dk2), z(da_op), z(db_op)
        dv2 = z(dv2)
        if not ctx.need_v2:                       # v' was v itself: its gradient passes straight through
            dk, dw_lo, da_lo = (torch.empty_like(k) for _ in range(3))
            dv_out, dv_lo, dv_first, dv_ptr, dv2_ptr = dv2, None, None, None, None
        else:
            dk, dw_lo, da_lo, dv_out = (torch.empty_like(k) for _ in range(4))
            dv_lo = torch.empty_like(k) if ctx.has_v else None
            dv_first = torch.empty_like(k) if ctx.has_v else None
            dv_ptr, dv2_ptr = _ptr(dv_out), _ptr(dv2)
        dparams = torch.empty(5, C, dtype=torch.float32, device=k.device)
        scratch = _scratch(B, T, C, 5, k.device)
        with torch.cuda.device(k.device):
            rc = _lib.lib().rwkvtts_tmix_prep_backward(
                B, T, C, _ptr(k), _ptr(v), _ptr(w_lo), _ptr(a_lo), _ptr(v_lo), _ptr(v_first), _ptr(mask), _ptr(w0), _ptr(a0),
                _ptr(v0), _ptr(k_k), _ptr(k_a), ctx.mask_rwk, _ptr(dw), _ptr(dk2), dv2_ptr, _ptr(da_op), _ptr(db_op), _ptr(dk), dv_ptr,
                _ptr(dw_lo), _ptr(da_lo), _ptr(dv_lo), _ptr(dv_first), _ptr(dparams), _ptr(scratch), _stream())
        _lib.check(rc, "rwkvtts_tmix_prep_backward")
        meta = lambda j: (ctx.dtypes[j], ctx.shapes[j])
        gw0, ga0, gv0, gkk, gka = _param_grads(dparams, [meta(0), meta(1), meta(4) if ctx.has_v else None, meta(2), meta(3)])
        return (dk, dv_out, dw_lo, da_lo, dv_lo, dv_first, gw0, ga0, gv0, gkk, gka, None, None)


def prep(k, v, w_lo, a_lo, v_lo, v_first, w0, a0, v0, k_k, k_a, mask=None, mask_rwk=True):
    """-> (w, k', v', a_op, b_op): everything between the projections / LoRAs and the WKV-7 op.
    v_lo / v_first / v0 are None on layer 0 (v' = v, masked).  mask_rwk: also mask w, k, v before use (the in-repo
    stack, :175-178); False = rwkvfla's RWKV7Attention, which masks only its input, kk and v'."""
    B, T, C = k.shape
    return _Prep.apply(k, v, w_lo, a_lo, v_lo, v_first, w0, a0, v0, k_k, k_a, _mask2d(mask, B, T), mask_rwk)


class _Out(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, r, k2, v2, g, r_k, ln_w, ln_b, eps):
        _need_cuda(y, r, k2, v2, g)
        B, T, C = y.shape
        y, r, k2, v2, g = (t.contiguous() for t in (y, r, k2, v2, g))
        p32 = [_f32(p, C) for p in (r_k, ln_w, ln_b)]
        o = torch.empty_like(y)
        with torch.cuda.device(y.device):
            rc = _lib.lib().rwkvtts_tmix_out_forward(B, T, C, _ptr(y), _ptr(r), _ptr(k2), _ptr(v2), _ptr(g),
                                                     *[_ptr(p) for p in p32], float(eps), _ptr(o), _stream())
        _lib.check(rc, "rwkvtts_tmix_out_forward")
        ctx.save_for_backward(y, r, k2, v2, g, *p32)
        ctx.eps = float(eps)
        ctx.meta = [(p.dtype, p.shape) for p in (r_k, ln_w, ln_b)]
        return o

    @staticmethod
    def backward(ctx, d_o):
        y, r, k2, v2, g, r_k, ln_w, ln_b = ctx.saved_tensors
        B, T, C = y.shape
        d_o = d_o.contiguous()
        dy, dr, dk2, dv2, dg = (torch.empty_like(y) for _ in range(5))
        dparams = torch.empty(3, C, dtype=torch.float32, device=y.device)
        scratch = _scratch(B, T, C, 3, y.device)
        with torch.cuda.device(y.device):
            rc = _lib.lib().rwkvtts_tmix_out_backward(B, T, C, _ptr(y), _ptr(r), _ptr(k2), _ptr(v2), _ptr(g), _ptr(r_k),
                                                      _ptr(ln_w), _ptr(ln_b), ctx.eps, _ptr(d_o), _ptr(dy), _ptr(dr), _ptr(dk2),
                                                      _ptr(dv2), _ptr(dg), _ptr(dparams), _ptr(scratch), _stream())
        _lib.check(rc, "rwkvtts_tmix_out_backward")
        dp = _param_grads(dparams, ctx.meta)
        return dy, dr, dk2, dv2, dg, dp[0], dp[1], dp[2], None


def out(y, r, k2, v2, g, r_k, ln_w, ln_b, eps):
    """(GroupNorm_H(y) * ln_w + ln_b + (sum_head r*k'*r_k) * v') * g: the input of the output projection."""
    return _Out.apply(y, r, k2, v2, g, r_k, ln_w, ln_b, eps)


class _SqRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = _lib.lib().rwkvtts_sqrelu_forward(x.numel(), _ptr(x), _ptr(y), _stream())
        _lib.check(rc, "rwkvtts_sqrelu_forward")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = _lib.lib().rwkvtts_sqrelu_backward(x.numel(), _ptr(x), _ptr(dy), _ptr(dx), _stream())
        _lib.check(rc, "rwkvtts_sqrelu_backward")
        return dx


def sqrelu(x: torch.Tensor) -> torch.Tensor:
    """relu(x) ** 2, the channel-mix activation (:228), one pass forward and one backward."""
    return _SqRelu.apply(x)


def ln_usable(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == BF16 and x.shape[-1] % 256 == 0
            and x.shape[-1] <= (MAX_C_BACKWARD if torch.is_grad_enabled() else MAX_C_FORWARD))


class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, w, b, eps):
        _need_cuda(x, res, w, b)
        C = x.shape[-1]
        rows = x.numel() // C
        x = x.contiguous()
        res = None if res is None else res.contiguous()
        w32, b32 = _f32(w, C), _f32(b, C)
        y = torch.empty_like(x)
        s = torch.empty_like(x) if res is not None else None
        need_bwd = any(ctx.needs_input_grad[:4])     # (grad mode is off inside forward: ask the context)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device) if need_bwd else None
        with torch.cuda.device(x.device):
            rc = _lib.lib().rwkvtts_add_layernorm_forward(rows, C, _ptr(x), _ptr(res), _ptr(w32), _ptr(b32), float(eps),
                                                          _ptr(y), _ptr(s), _ptr(stats), _stream())
        _lib.check(rc, "rwkvtts_add_layernorm_forward")
        total = s if res is not None else x
        ctx.save_for_backward(total, stats, w32)
        ctx.meta = (w.dtype, w.shape, None if b is None else (b.dtype, b.shape), res is not None)
        return y, total

    @staticmethod
    def backward(ctx, dy, ds):
        total, stats, w32 = ctx.saved_tensors
        C = total.shape[-1]
        rows = total.numel() // C
        dy = torch.zeros_like(total) if dy is None else dy.contiguous()
        ds = None if ds is None else ds.contiguous()
        dx = torch.empty_like(total)
        dparams = torch.empty(2, C, dtype=torch.float32, device=total.device)
        scratch = _scratch(1, rows, C, 2, total.device)
        with torch.cuda.device(total.device):
            rc = _lib.lib().rwkvtts_add_layernorm_backward(rows, C, _ptr(total), _ptr(stats), _ptr(w32), _ptr(dy), _ptr(ds),
                                                           _ptr(dx), _ptr(dparams), _ptr(scratch), _stream())
        _lib.check(rc, "rwkvtts_add_layernorm_backward")
        wdt, wsh, bmeta, has_res = ctx.meta
        dw, db = _param_grads(dparams, [(wdt, wsh), bmeta])
        return dx, (dx if has_res else None), dw, db, None


def add_layernorm(x, residual, weight, bias, eps):
    """(LayerNorm(x + residual) * weight + bias, x + residual); residual may be None (then the second result is x).
    forward stats are kept only when a backward can follow (no_grad: inference / decode)."""
    return _AddLayerNorm.apply(x, residual, weight, bias, eps)


# ---------------------------------------------------------------------------------------------------------------
# caller side of the path: embedding gather / concat of a speech layout, and the fused linear + cross-entropy head
# (kernels: csrc/gather.cu, csrc/linear_ce.cu)
# ---------------------------------------------------------------------------------------------------------------
TABLE_SHIFT = 40                      # row_src = (table << 40) | row, negative = padding


def gather_usable(weights) -> bool:
    w0 = weights[0]
    return (w0.is_cuda and all(w.dtype == BF16 and w.is_contiguous() and w.shape[1] == w0.shape[1] for w in weights)
            and w0.shape[1] % 8 == 0 and len(weights) <= 8)


class _EmbedRows(torch.autograd.Function):
    """out[r] = weights[table(r)][row(r)] for every position of the padded / packed batch, zeros where row_src < 0: one
    kernel instead of a lookup and a scatter per table into a zero-filled buffer.  Backward: what nn.Embedding does,
    per table, on the rows that came from it (aten::embedding_dense_backward, so the numerics of the tables' gradients
    are the reference's)."""

    @staticmethod
    def forward(ctx, row_src, n_tables, *rest):
        weights, ids, dst = rest[:n_tables], rest[n_tables:2 * n_tables], rest[2 * n_tables:]
        rows, D = row_src.numel(), weights[0].shape[1]
        out = torch.empty(rows, D, dtype=BF16, device=row_src.device)
        tabs = _ptr_array(weights)
        with torch.cuda.device(row_src.device):
            rc = _lib.lib().rwkvtts_embed_rows(tabs, n_tables, _ptr(row_src), rows, D, _ptr(out), _stream())
        _lib.check(rc, "rwkvtts_embed_rows")
        ctx.n_tables, ctx.sizes = n_tables, [w.shape[0] for w in weights]
        ctx.save_for_backward(*ids, *dst)
        return out

    @staticmethod
    def backward(ctx, d_out):
        n = ctx.n_tables
        ids, dst = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        grads = []
        for k in range(n):
            if not ctx.needs_input_grad[2 + k] or ids[k].numel() == 0:
                grads.append(None)
                continue
            g = d_out.index_select(0, dst[k])
            grads.append(torch.ops.aten.embedding_dense_backward(g, ids[k], ctx.sizes[k], -1, False))
        return (None, None, *grads, *([None] * (2 * n)))


def embed_rows(weights, row_src, ids, dst):
    """weights: embedding tables [n_k, D] bf16; row_src int64 [rows] on the device; ids[k] / dst[k]: the rows taken from
    table k and where they went (for the tables' gradients).  Returns [rows, D]."""
    _need_cuda(row_src, *weights)
    return _EmbedRows.apply(row_src, len(weights), *weights, *ids, *dst)


def linear_ce_usable(h, weight, bias) -> bool:
    return (h.is_cuda and h.dtype == BF16 and weight.dtype == BF16 and (bias is None or bias.dtype == BF16)
            and h.shape[-1] % 8 == 0 and weight.shape[0] <= (1 << 20))


class _LinearCE(torch.autograd.Function):
    """mean / sum cross-entropy of h @ W^T against labels without the [tokens, V] logits: per token chunk one cuBLAS
    GEMM for the logits, the CE kernel (loss and d logits in one in-place pass), and the two gradient GEMMs -- all in
    the forward; the backward only scales by the incoming gradient.  dW accumulates in fp32 across the chunks."""

    @staticmethod
    def forward(ctx, h, labels, weight, bias, ignore_index, label_smoothing, reduction, chunk_rows):
        N, D = h.shape
        V = weight.shape[0]
        Vp = (V + 7) // 8 * 8
        Wp = weight if Vp == V else torch.nn.functional.pad(weight, (0, 0, 0, Vp - V))       # aligned tensor-core GEMMs
        bp = None if bias is None else (bias if Vp == V else torch.nn.functional.pad(bias, (0, Vp - V)))
        valid = labels != ignore_index
        if reduction == "mean":
            scale = (1.0 / valid.sum().clamp(min=1).to(torch.float32)).reshape(1)
        else:
            scale = torch.ones(1, dtype=torch.float32, device=h.device)
        loss_rows = torch.empty(N, dtype=torch.float32, device=h.device)
        need = torch.is_grad_enabled() or h.requires_grad or weight.requires_grad
        dh = torch.empty_like(h) if need else None
        dW = torch.zeros(Vp, D, dtype=torch.float32, device=h.device) if need else None
        db = torch.zeros(Vp, dtype=torch.float32, device=h.device) if (need and bias is not None) else None
        L = _lib.lib()
        for s in range(0, N, chunk_rows):
            e = min(N, s + chunk_rows)
            hc = h[s:e]
            logits = torch.mm(hc, Wp.t()) if bp is None else torch.addmm(bp, hc, Wp.t())
            with torch.cuda.device(h.device):
                rc = L.rwkvtts_ce_forward_backward(_ptr(logits), e - s, V, Vp, _ptr(labels[s:e]), int(ignore_index),
                                                   float(label_smoothing), _ptr(scale), _ptr(loss_rows[s:e]), _stream())
            _lib.check(rc, "rwkvtts_ce_forward_backward")
            if need:
                torch.mm(logits, Wp, out=dh[s:e])
                dW += torch.mm(logits.t(), hc, out_dtype=torch.float32)
                if db is not None:
                    db += logits.sum(0, dtype=torch.float32)
        loss = loss_rows.sum() * scale[0]
        if need:
            ctx.has_bias = bias is not None
            ctx.save_for_backward(dh, dW[:V].to(BF16), *([db[:V].to(BF16)] if db is not None else []))
        return loss

    @staticmethod
    def backward(ctx, g):
        dh, dW = ctx.saved_tensors[:2]
        g = g.to(torch.float32)
        dbias = (ctx.saved_tensors[2] * g).to(BF16) if ctx.has_bias else None
        return (dh * g).to(dh.dtype), None, (dW * g).to(dW.dtype), dbias, None, None, None, None


def linear_cross_entropy(h, labels, weight, ignore_index=-100, label_smoothing=0.0, reduction="mean", chunk_rows=4096,
                         bias=None):
    """h [N, D] bf16, labels [N] int64, weight [V, D] bf16 (+ bias [V]) -> scalar fp32 loss (see _LinearCE)."""
    _need_cuda(h, labels, weight, bias)
    return _LinearCE.apply(h.contiguous(), labels.contiguous(), weight.contiguous(), bias, ignore_index, label_smoothing,
                           reduction, chunk_rows)
