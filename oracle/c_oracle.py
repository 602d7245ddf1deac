"""ctypes loaders for the oracle's native pieces -- TEST INFRASTRUCTURE ONLY.

* ``liboracle_wkv7.so``: plain-C fp32 port of the reference algorithm (wkv7_oracle.c).
* ``_ref/libref_*.so``: the unmodified reference CUDA kernels compiled for sm_100a by
  ``oracle/Makefile`` (GPU cross-check and "reference CUDA op" speed baseline).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp = ctypes.c_void_p


def build(ref: bool = True) -> None:
    """Compile the C oracle and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference/model/llm/cuda"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing -- run `make -C oracle` (and `make -C oracle ref`)")
    return ctypes.CDLL(path)


_c = None


def clib():
    global _c
    if _c is None:
        _c = _load(os.path.join(_HERE, "liboracle_wkv7.so"))
        _c.oracle_wkv7_num_threads.restype = ctypes.c_int
    return _c


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _chk(*ts):
    for t in ts:
        assert t.device.type == "cpu" and t.is_contiguous()


def num_threads() -> int:
    return int(clib().oracle_wkv7_num_threads())


def c_forward(w, q, k, v, a, b, save=True):
    """fp32 C port of forward_kernel (wkv7_cuda.cu:10-52).  bf16 CPU tensors [B,T,H,64]."""
    _chk(w, q, k, v, a, b)
    B, T, H, C = w.shape
    assert C == 64 and w.dtype == torch.bfloat16
    y = torch.empty_like(v)
    s = torch.empty(B, H, T // 16, C, C, dtype=torch.float32) if save else None
    sa = torch.empty(B, T, H, C, dtype=torch.float32) if save else None
    clib().oracle_wkv7_forward(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(y), _p(s), _p(sa))
    return y, s, sa


def c_backward(w, q, k, v, a, b, dy, s, sa):
    """fp32 C port of backward_kernel (wkv7_cuda.cu:54-130)."""
    _chk(w, q, k, v, a, b, dy, s, sa)
    B, T, H, C = w.shape
    assert T % 16 == 0
    outs = [torch.empty_like(w) for _ in range(6)]
    clib().oracle_wkv7_backward(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(dy), _p(s), _p(sa),
                                *[_p(o) for o in outs])
    return outs


def c_state_forward(state, r, w, k, v, a, b):
    """fp32 C port of rwkv7_state_fwd_fp16.cu:9-57; ``state`` [B,H,64,64] updated in place."""
    _chk(state, r, w, k, v, a, b)
    B, T, HC = r.shape
    H = HC // 64
    y = torch.empty_like(r)
    clib().oracle_wkv7_state_forward(B, T, H, _p(state), _p(r), _p(w), _p(k), _p(v), _p(a), _p(b), _p(y))
    return y


# ----------------------------------------------------------------------------------------
# reference CUDA kernels (oracle/_ref) -- need a GPU
# ----------------------------------------------------------------------------------------
_ref_wind = None
_ref_state = None


def ref_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", n))
               for n in ("libref_wind_backstepping.so", "libref_state_fwd.so"))


def ref_wind():
    global _ref_wind
    if _ref_wind is None:
        _ref_wind = _load(os.path.join(_HERE, "_ref", "libref_wind_backstepping.so"))
    return _ref_wind


def ref_state():
    global _ref_state
    if _ref_state is None:
        _ref_state = _load(os.path.join(_HERE, "_ref", "libref_state_fwd.so"))
    return _ref_state


def ref_forward(w, q, k, v, a, b):
    """Reference forward_kernel on the GPU (legacy default stream, like the reference)."""
    B, T, H, C = w.shape
    y = torch.empty_like(v)
    s = torch.empty(B, H, T // 16, C, C, dtype=torch.float32, device=w.device)
    sa = torch.empty(B, T, H, C, dtype=torch.float32, device=w.device)
    rc = ref_wind().ref_wind_forward(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(y), _p(s), _p(sa))
    assert rc == 0, f"reference forward launch failed: cudaError {rc}"
    return y, s, sa


def ref_backward(w, q, k, v, a, b, dy, s, sa):
    B, T, H, C = w.shape
    outs = [torch.empty_like(w) for _ in range(6)]
    rc = ref_wind().ref_wind_backward(B, T, H, _p(w), _p(q), _p(k), _p(v), _p(a), _p(b), _p(dy), _p(s), _p(sa),
                                      *[_p(o) for o in outs])
    assert rc == 0, f"reference backward launch failed: cudaError {rc}"
    return outs


def ref_state_forward(state, r, w, k, v, a, b):
    B, T, HC = r.shape
    y = torch.empty_like(r)
    rc = ref_state().ref_state_forward(B, T, HC, HC // 64, _p(state), _p(r), _p(w), _p(k), _p(v), _p(a), _p(b),
                                       _p(y))
    assert rc == 0, f"reference state forward launch failed: cudaError {rc}"
    return y
