"""Host-side mirror of the reference's WKV-7 operator interface.

Same names, argument order, asserts and in-place semantics as the reference wrappers, so the
reference models (RWKV_Tmix_x070 & co.) can `from rwkvtts_b200.ops import ...` unchanged:

  WindBackstepping, RUN_CUDA_RWKV7g   /root/reference/model/llm/rwkv_s2s_single_ffn.py:15-40
  WKV_7, RWKV7_OP                     /root/reference/model/llm/rwkv_s2s_single_ffn.py:45-59
  WKV_7_batch, RWKV7_BATCH_OP         /root/reference/model/llm/rwkv_asr_cuda_whisper.py:67-83

and the torch dispatcher ops with the reference's schemas (CUDA key):

  wind_backstepping::forward/backward   model/llm/cuda/wkv7_op.cpp:21-29
  wkv7s::forward                        model/llm/cuda/wkv7s_op.cpp:13-15
  rwkv7_state_fwd_fp16::forward         model/llm/cuda/rwkv7_state_fwd_fp16.cpp:12-14

All of them go through the C ABI (include/rwkvtts_wkv7.h) on the caller's current CUDA stream.
PyTorch is only the owner of device memory and streams here.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

HEAD_SIZE = 64
CHUNK_LEN = 16
DTYPE = torch.bfloat16


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# bench.py sets this to {"fwd": [], "bwd": []}: every launch of the training pair is then bracketed by CUDA events on
# the launching stream, so the kernels' durations are measured inside the timed region of a whole-model step.
TIMING = None


class _timed:
    def __init__(self, kind):
        self.rec = TIMING[kind] if TIMING is not None else None

    def __enter__(self):
        if self.rec is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.rec is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.rec.append((self.e0, e1))
        return False


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.RwkvttsError("rwkvtts_b200 ops need CUDA tensors: there is no CPU fallback")


# ---------------------------------------------------------------------------------------
# raw calls (tensor-level; used by the dispatcher ops, the autograd functions and bench.py)
# ---------------------------------------------------------------------------------------
def wkv7_forward_(w, q, k, v, z, a, y, s, sa, s0=None, sT=None):
    _need_cuda(w, q, k, v, z, a, y, s, sa, s0, sT)
    B, T, H, _ = w.shape
    with torch.cuda.device(w.device), _timed("fwd"):
        rc = _lib.lib().rwkvtts_wkv7_forward_ex(B, T, H, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(z), _ptr(a),
                                                _ptr(y), _ptr(s), _ptr(sa), _ptr(s0), _ptr(sT), _stream())
    _lib.check(rc, "rwkvtts_wkv7_forward")


def wkv7_forward_infer_(w, q, k, v, z, a, y, s0=None, sT=None):
    """Snapshot-free forward (no backward possible): the chunked tcgen05 kernel."""
    _need_cuda(w, q, k, v, z, a, y, s0, sT)
    B, T, H, _ = w.shape
    with torch.cuda.device(w.device):
        rc = _lib.lib().rwkvtts_wkv7_forward_infer(B, T, H, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(z), _ptr(a),
                                                   _ptr(y), _ptr(s0), _ptr(sT), _stream())
    _lib.check(rc, "rwkvtts_wkv7_forward_infer")


def wkv7_backward_(w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da, s0=None, dsT=None, ds0=None, sT=None):
    _need_cuda(w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da, s0, dsT, ds0, sT)
    B, T, H, _ = w.shape
    with torch.cuda.device(w.device), _timed("bwd"):
        rc = _lib.lib().rwkvtts_wkv7_backward_ex(B, T, H, _ptr(w), _ptr(q), _ptr(k), _ptr(v), _ptr(z), _ptr(a),
                                                 _ptr(dy), _ptr(s), _ptr(sa), _ptr(s0), _ptr(sT), _ptr(dsT), _ptr(dw),
                                                 _ptr(dq), _ptr(dk), _ptr(dv), _ptr(dz), _ptr(da), _ptr(ds0),
                                                 _stream())
    _lib.check(rc, "rwkvtts_wkv7_backward")


def wkv7_state_forward_(B, T, C, H, state, r, w, k, v, a, b, y):
    _need_cuda(state, r, w, k, v, a, b, y)
    with torch.cuda.device(r.device):
        rc = _lib.lib().rwkvtts_wkv7_state_forward(int(B), int(T), int(C), int(H), _ptr(state), _ptr(r), _ptr(w),
                                                   _ptr(k), _ptr(v), _ptr(a), _ptr(b), _ptr(y), _stream())
    _lib.check(rc, "rwkvtts_wkv7_state_forward")


# ---------------------------------------------------------------------------------------
# torch dispatcher ops with the reference schemas
# ---------------------------------------------------------------------------------------
_registered = False
_libs = []


def register_torch_ops() -> None:
    """Defines the reference's op schemas and binds their CUDA kernels to the C ABI."""
    global _registered
    if _registered:
        return
    L = torch.library.Library("wind_backstepping", "DEF")
    L.define("forward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, "
             "Tensor(a!) y, Tensor(b!) s, Tensor(c!) sa) -> ()")
    L.define("backward(Tensor w, Tensor q, Tensor k, Tensor v, Tensor z, Tensor a, Tensor dy, Tensor s, "
             "Tensor sa, Tensor(a!) dw, Tensor(b!) dq, Tensor(c!) dk, Tensor(d!) dv, Tensor(e!) dz, "
             "Tensor(f!) da) -> ()")
    L.impl("forward", lambda w, q, k, v, z, a, y, s, sa: wkv7_forward_(w, q, k, v, z, a, y, s, sa), "CUDA")
    L.impl("backward", lambda w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da:
           wkv7_backward_(w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da), "CUDA")
    _libs.append(L)
    for ns in ("wkv7s", "rwkv7_state_fwd_fp16"):
        L = torch.library.Library(ns, "DEF")
        L.define("forward(int B, int T, int C, int H, Tensor(a!) state, Tensor r, Tensor w, Tensor k, Tensor v, "
                 "Tensor a, Tensor b, Tensor(b!) y) -> ()")
        L.impl("forward", wkv7_state_forward_, "CUDA")
        _libs.append(L)
    _registered = True


# ---------------------------------------------------------------------------------------
# the reference's Python wrappers, same names
# ---------------------------------------------------------------------------------------
class WindBackstepping(torch.autograd.Function):
    """rwkv_s2s_single_ffn.py:15-35 (argument order w,q,k,v,z,b)."""

    @staticmethod
    def forward(ctx, w, q, k, v, z, b):
        B, T, H, C = w.shape
        assert T % CHUNK_LEN == 0
        assert all(i.dtype == torch.bfloat16 for i in [w, q, k, v, z, b])
        assert all(i.is_contiguous() for i in [w, q, k, v, z, b])
        y = torch.empty_like(v)
        if not any(ctx.needs_input_grad):      # torch.no_grad() / frozen inputs: no scratch, tcgen05 kernel
            wkv7_forward_infer_(w, q, k, v, z, b, y)
            return y
        s = torch.empty(B, H, T // CHUNK_LEN, C, C, dtype=torch.float32, device=w.device)
        sa = torch.empty(B, T, H, C, dtype=torch.float32, device=w.device)
        torch.ops.wind_backstepping.forward(w, q, k, v, z, b, y, s, sa)
        ctx.save_for_backward(w, q, k, v, z, b, s, sa)
        return y

    @staticmethod
    def backward(ctx, dy):
        assert all(i.dtype == torch.bfloat16 for i in [dy])
        assert all(i.is_contiguous() for i in [dy])
        w, q, k, v, z, b, s, sa = ctx.saved_tensors
        dw, dq, dk, dv, dz, db = [torch.empty_like(x) for x in [w, q, k, v, z, b]]
        torch.ops.wind_backstepping.backward(w, q, k, v, z, b, dy, s, sa, dw, dq, dk, dv, dz, db)
        return dw, dq, dk, dv, dz, db


def RUN_CUDA_RWKV7g(q, w, k, v, a, b):
    """rwkv_s2s_single_ffn.py:37-40."""
    B, T, HC = q.shape
    q, w, k, v, a, b = [i.view(B, T, HC // 64, 64) for i in [q, w, k, v, a, b]]
    return WindBackstepping.apply(w, q, k, v, a, b).view(B, T, HC)


class WKV_7(torch.autograd.Function):
    """rwkv_s2s_single_ffn.py:45-57: stateful forward, B = 1, state [H,64,64] updated in place."""

    @staticmethod
    def forward(ctx, state, r, w, k, v, a, b):
        with torch.no_grad():
            T, C = r.size()
            H = C // HEAD_SIZE
            assert HEAD_SIZE == C // H
            assert all(x.dtype == DTYPE for x in [r, w, k, v, a, b])
            assert all(x.is_contiguous() for x in [r, w, k, v, a, b])
            y = torch.empty((T, C), device=k.device, dtype=DTYPE, requires_grad=False,
                            memory_format=torch.contiguous_format)
            torch.ops.wkv7s.forward(1, T, C, H, state, r, w, k, v, a, b, y)
            return y


def RWKV7_OP(state, r, w, k, v, a, b):
    return WKV_7.apply(state, r, w, k, v, a, b)


class WKV_7_batch(torch.autograd.Function):
    """rwkv_asr_cuda_whisper.py:67-81: batched stateful forward, state [B,H,64,64] in place."""

    @staticmethod
    def forward(ctx, state, r, w, k, v, a, b):
        with torch.no_grad():
            B, T, C = r.size()
            H = C // HEAD_SIZE
            assert HEAD_SIZE == C // H
            assert all(x.dtype == DTYPE for x in [r, w, k, v, a, b])
            assert all(x.is_contiguous() for x in [r, w, k, v, a, b])
            y = torch.empty((B, T, C), device=k.device, dtype=DTYPE, requires_grad=False,
                            memory_format=torch.contiguous_format)
            torch.ops.rwkv7_state_fwd_fp16.forward(B, T, C, H, state, r, w, k, v, a, b, y)
            return y


def RWKV7_BATCH_OP(state, r, w, k, v, a, b):
    return WKV_7_batch.apply(state, r, w, k, v, a, b)


class _Wkv7WithState(torch.autograd.Function):
    """chunk_rwkv7(r,w,k,v,a,b, initial_state, output_final_state) semantics on the
    reference's value-major state layout; `w` is the BlinkDL pre-activation."""

    @staticmethod
    def forward(ctx, w, q, k, v, a, b, s0):
        B, T, H, C = w.shape
        y = torch.empty_like(v)
        sT = torch.empty(B, H, C, C, dtype=torch.float32, device=w.device)
        if not any(ctx.needs_input_grad):
            wkv7_forward_infer_(w, q, k, v, a, b, y, s0=s0, sT=sT)
            return y, sT
        s = torch.empty(B, H, T // CHUNK_LEN, C, C, dtype=torch.float32, device=w.device)
        sa = torch.empty(B, T, H, C, dtype=torch.float32, device=w.device)
        wkv7_forward_(w, q, k, v, a, b, y, s, sa, s0=s0, sT=sT)
        ctx.save_for_backward(w, q, k, v, a, b, s, sa, s0, sT)
        return y, sT

    @staticmethod
    def backward(ctx, dy, dsT):
        w, q, k, v, a, b, s, sa, s0, sT = ctx.saved_tensors
        grads = [torch.empty_like(x) for x in (w, q, k, v, a, b)]
        ds0 = torch.empty_like(sT) if s0 is not None else None
        wkv7_backward_(w, q, k, v, a, b, dy.contiguous(), s, sa, *grads, s0=s0,
                       dsT=dsT.contiguous() if dsT is not None else None, ds0=ds0, sT=sT.detach())
        return (*grads, ds0)


class VarlenPlan:
    """Device-side description of a packed batch, built once per batch (no host sync): cu_seqlens as int32, the first
    chunk slot of every sequence inside the scratch tensors, and the per-token `first` flags the token shift needs."""

    def __init__(self, cu_seqlens: torch.Tensor, total: int):
        cu = cu_seqlens.reshape(-1).to(torch.int32).contiguous()
        self.cu, self.N, self.total = cu, cu.numel() - 1, int(total)
        chunks = (cu[1:] - cu[:-1] + (CHUNK_LEN - 1)) // CHUNK_LEN
        self.cbase = torch.cat([torch.zeros(1, dtype=torch.int32, device=cu.device), chunks.cumsum(0).to(torch.int32)]).contiguous()
        first = torch.zeros(self.total + 1, dtype=torch.uint8, device=cu.device)
        first.index_fill_(0, cu.to(torch.long), 1)
        self.first = first[:self.total].contiguous()


class _Wkv7Varlen(torch.autograd.Function):
    """chunk_rwkv7(..., cu_seqlens=) semantics on one packed [1, T_total, H, 64] tensor: zero state at every boundary."""

    @staticmethod
    def forward(ctx, w, q, k, v, a, b, plan):
        _need_cuda(w, q, k, v, a, b)
        _, T, H, C = w.shape
        y = torch.empty_like(v)
        L = _lib.lib()
        train = any(ctx.needs_input_grad[:6])
        s = sa = None
        if train:
            ns, nsa = ctypes.c_size_t(), ctypes.c_size_t()
            L.rwkvtts_wkv7_varlen_scratch_floats(T, H, plan.N, ctypes.byref(ns), ctypes.byref(nsa))
            s = torch.empty(ns.value, dtype=torch.float32, device=w.device)
            sa = torch.empty(nsa.value, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device), _timed("fwd"):
            rc = L.rwkvtts_wkv7_forward_varlen(T, H, plan.N, _ptr(plan.cu), _ptr(plan.cbase), _ptr(w), _ptr(q), _ptr(k),
                                               _ptr(v), _ptr(a), _ptr(b), _ptr(y), _ptr(s), _ptr(sa), _stream())
        _lib.check(rc, "rwkvtts_wkv7_forward_varlen")
        if train:
            ctx.save_for_backward(w, q, k, v, a, b, s, sa)
            ctx.plan = plan
        return y

    @staticmethod
    def backward(ctx, dy):
        w, q, k, v, a, b, s, sa = ctx.saved_tensors
        plan = ctx.plan
        _, T, H, C = w.shape
        grads = [torch.empty_like(x) for x in (w, q, k, v, a, b)]
        with torch.cuda.device(w.device), _timed("bwd"):
            rc = _lib.lib().rwkvtts_wkv7_backward_varlen(T, H, plan.N, _ptr(plan.cu), _ptr(plan.cbase), _ptr(w), _ptr(q),
                                                         _ptr(k), _ptr(v), _ptr(a), _ptr(b), _ptr(dy.contiguous()), _ptr(s),
                                                         _ptr(sa), *[_ptr(g) for g in grads], _stream())
        _lib.check(rc, "rwkvtts_wkv7_backward_varlen")
        return (*grads, None)


def wkv7_varlen(w, q, k, v, a, b, plan: "VarlenPlan"):
    """y for a packed batch: w,q,k,v,a,b bf16 [1, T_total, H, 64] contiguous (op order), `plan` from VarlenPlan."""
    return _Wkv7Varlen.apply(w, q, k, v, a, b, plan)


def wkv7_with_state(w, q, k, v, a, b, initial_state=None):
    """y, final_state = WKV-7 over [B,T,H,64] starting from `initial_state` ([B,H,64,64] or None)."""
    return _Wkv7WithState.apply(w, q, k, v, a, b, initial_state)
