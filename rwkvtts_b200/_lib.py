"""ctypes binding of librwkvtts_wkv7.so (include/rwkvtts_wkv7.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is
missing or a symbol of the header is absent, importing/calling fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RWKVTTS_LIB selects another build of the same library (the stress variants of rwkvtts_b200/build.py)
LIB_PATH = os.environ.get("RWKVTTS_LIB") or os.path.join(_HERE, "librwkvtts_wkv7.so")

_vp, _i, _fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p

# every symbol include/rwkvtts_wkv7.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "rwkvtts_version": (_i, []),
    "rwkvtts_strerror": (ctypes.c_char_p, [_i]),
    "rwkvtts_last_cuda_error": (_i, []),
    "rwkvtts_kernel_launches": (ctypes.c_longlong, []),
    "rwkvtts_watchdog_report": (_i, [ctypes.c_char_p, ctypes.c_size_t]),
    "rwkvtts_set_impl": (_i, [_i]),
    "rwkvtts_get_impl": (_i, []),
    "rwkvtts_set_step_mode": (_i, [_i]),
    "rwkvtts_get_step_mode": (_i, []),
    "rwkvtts_wkv7_scratch_floats": (ctypes.c_size_t, [_i, _i, _i, ctypes.POINTER(ctypes.c_size_t),
                                                      ctypes.POINTER(ctypes.c_size_t)]),
    "rwkvtts_wkv7_forward": (_i, [_i, _i, _i] + [_vp] * 6 + [_vp, _fp, _fp, _vp]),
    "rwkvtts_wkv7_forward_infer": (_i, [_i, _i, _i] + [_vp] * 6 + [_vp, _fp, _fp, _vp]),
    "rwkvtts_wkv7_backward": (_i, [_i, _i, _i] + [_vp] * 7 + [_fp, _fp] + [_vp] * 6 + [_vp]),
    "rwkvtts_wkv7_forward_ex": (_i, [_i, _i, _i] + [_vp] * 6 + [_vp, _fp, _fp, _fp, _fp, _vp]),
    "rwkvtts_wkv7_backward_ex": (_i, [_i, _i, _i] + [_vp] * 7 + [_fp, _fp, _fp, _fp, _fp] + [_vp] * 6 + [_fp, _vp]),
    "rwkvtts_wkv7_varlen_scratch_floats": (ctypes.c_size_t, [_i, _i, _i, ctypes.POINTER(ctypes.c_size_t),
                                                             ctypes.POINTER(ctypes.c_size_t)]),
    "rwkvtts_wkv7_forward_varlen": (_i, [_i, _i, _i, _vp, _vp] + [_vp] * 6 + [_vp, _fp, _fp, _vp]),
    "rwkvtts_wkv7_backward_varlen": (_i, [_i, _i, _i, _vp, _vp] + [_vp] * 7 + [_fp, _fp] + [_vp] * 6 + [_vp]),
    "rwkvtts_wkv7_state_forward": (_i, [_i, _i, _i, _i, _fp] + [_vp] * 6 + [_vp, _vp]),
    "rwkvtts_tmix_scratch_floats": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "rwkvtts_tmix_shift_mix_forward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _fp, ctypes.POINTER(_vp), _vp, _vp]),
    "rwkvtts_tmix_shift_mix_backward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _fp, ctypes.POINTER(_vp), _vp, _fp, _fp, _vp]),
    "rwkvtts_tmix_shift_mix_forward_varlen": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _fp, ctypes.POINTER(_vp), _vp, _vp, _vp]),
    "rwkvtts_tmix_shift_mix_backward_varlen": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _fp, ctypes.POINTER(_vp), _vp, _fp, _fp, _vp, _vp]),
    "rwkvtts_tmix_prep_forward": (_i, [_i, _i, _i] + [_vp] * 7 + [_fp] * 5 + [_i] + [_vp] * 5 + [_vp]),
    "rwkvtts_tmix_prep_backward": (_i, [_i, _i, _i] + [_vp] * 7 + [_fp] * 5 + [_i] + [_vp] * 5 + [_vp] * 6 + [_fp, _fp, _vp]),
    "rwkvtts_tmix_out_forward": (_i, [_i, _i, _i] + [_vp] * 5 + [_fp] * 3 + [ctypes.c_float, _vp, _vp]),
    "rwkvtts_tmix_out_backward": (_i, [_i, _i, _i] + [_vp] * 5 + [_fp] * 3 + [ctypes.c_float] + [_vp] * 6 + [_fp, _fp, _vp]),
    "rwkvtts_add_layernorm_forward": (_i, [ctypes.c_longlong, _i, _vp, _vp, _fp, _fp, ctypes.c_float, _vp, _vp, _fp, _vp]),
    "rwkvtts_add_layernorm_backward": (_i, [ctypes.c_longlong, _i, _vp, _fp, _fp, _vp, _vp, _vp, _fp, _fp, _vp]),
    "rwkvtts_sqrelu_forward": (_i, [ctypes.c_longlong, _vp, _vp, _vp]),
    "rwkvtts_sqrelu_backward": (_i, [ctypes.c_longlong, _vp, _vp, _vp, _vp]),
    "rwkvtts_adam_shard": (_i, [_fp, _fp, _fp, _vp, _i, _vp, _i, ctypes.c_longlong] + [ctypes.c_float] * 5 + [_i]
                           + [ctypes.c_float] * 3 + [_vp]),
    "rwkvtts_adam_multi": (_i, [_fp, _fp, _fp, _vp, _i, _vp, _i, ctypes.c_longlong, _vp, _vp, _i,
                                ctypes.POINTER(ctypes.c_float), _i] + [ctypes.c_float] * 3 + [_i, _fp, ctypes.c_float, _vp, _vp]),
    "rwkvtts_embed_rows": (_i, [ctypes.POINTER(_vp), _i, _vp, ctypes.c_longlong, _i, _vp, _vp]),
    "rwkvtts_multi_copy": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                ctypes.POINTER(_i), _i, _vp, _i, _vp]),
    "rwkvtts_ce_forward_backward": (_i, [_vp, ctypes.c_longlong, _i, ctypes.c_longlong, _vp, ctypes.c_longlong, ctypes.c_float,
                                         _fp, _fp, _vp]),
    "rwkvtts_adam_p2p": (_i, [_fp, _fp, _fp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _vp, _i, ctypes.c_longlong,
                              ctypes.c_longlong, _vp, _vp, _i, ctypes.POINTER(ctypes.c_float), _i] + [ctypes.c_float] * 3
                         + [_i, _fp, _fp, _vp, _vp]),
    "rwkvtts_grad_stat": (_i, [_vp, _i, ctypes.c_longlong, _fp, _vp]),
    "rwkvtts_decode_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_size_t)]),
    "rwkvtts_decode_init": (_i, [ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                 _vp, ctypes.c_size_t, _vp]),
    "rwkvtts_decode_step": (_i, [_vp, _vp, _vp, _i, _i, ctypes.POINTER(ctypes.c_longlong), _i, ctypes.c_longlong, _vp]),
    "rwkvtts_decode_step_profile": (_i, [_vp, _vp, _vp, _vp]),
    "rwkvtts_decode_release": (_i, [_vp]),
}

_lib = None


class RwkvttsError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RwkvttsError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m rwkvtts_b200.build`). "
                "There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def watchdog_report() -> str:
    """Text of the mbarrier watchdog record of the chunked kernels ('' if none fired); include/rwkvtts_wkv7.h."""
    buf = ctypes.create_string_buffer(4096)
    return buf.value.decode() if lib().rwkvtts_watchdog_report(buf, 4096) else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        msg = L.rwkvtts_strerror(rc).decode()
        extra = f" (cudaError {L.rwkvtts_last_cuda_error()})" if rc == -4 else ""
        raise RwkvttsError(f"{what}: {msg}{extra}")
