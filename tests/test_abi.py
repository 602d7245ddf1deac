"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/rwkvtts_wkv7.h declares; the reference op schemas are registered; nothing falls back to CPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from rwkvtts_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "rwkvtts_wkv7.h")).read()
    declared = re.findall(r"RWKVTTS_API\s+[\w\s\*]+?\b(rwkvtts_\w+)\s*\(", hdr)
    assert len(declared) >= 10
    L = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(L, name), name
    from rwkvtts_b200 import _lib
    assert set(declared) == set(_lib.SYMBOLS), set(declared) ^ set(_lib.SYMBOLS)
    assert _lib.lib().rwkvtts_version() >= 100


def test_argument_validation_without_gpu(built):
    """Shape / pointer checks happen before any CUDA call, so they run on a CPU box."""
    from rwkvtts_b200 import _lib
    L = _lib.lib()
    buf = torch.zeros(1 << 16, dtype=torch.float32)
    p = buf.data_ptr()
    ps = [p] * 6
    assert L.rwkvtts_wkv7_forward(1, 15, 1, *ps, p, p, p, None) == -1          # T % 16 != 0
    assert L.rwkvtts_wkv7_forward(0, 16, 1, *ps, p, p, p, None) == -1
    assert L.rwkvtts_wkv7_forward(1, 16, 1, None, *ps[1:], p, p, p, None) == -2
    assert L.rwkvtts_wkv7_forward(1, 16, 1, p + 2, *ps[1:], p, p, p, None) == -3
    assert L.rwkvtts_wkv7_state_forward(1, 1, 100, 2, p, *ps, p, None) == -1   # C != H*64
    assert b"16" in L.rwkvtts_strerror(-1)
    s = torch.zeros(2, dtype=torch.int64)
    import ctypes
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    tot = L.rwkvtts_wkv7_scratch_floats(8, 4096, 16, ctypes.byref(a), ctypes.byref(b))
    assert a.value == 8 * 16 * 256 * 64 * 64 and b.value == 8 * 4096 * 16 * 64 and tot == a.value + b.value


def test_fused_kernel_argument_validation_without_gpu(built):
    """The fused time-mix / LayerNorm / activation entry points validate shapes and pointers before any CUDA call."""
    import ctypes
    from rwkvtts_b200 import _lib, fused
    L = _lib.lib()
    buf = torch.zeros(1 << 16, dtype=torch.float32)
    p = buf.data_ptr()
    outs = (ctypes.c_void_p * 6)(*[p] * 6)
    assert L.rwkvtts_tmix_shift_mix_forward(1, 4, 100, 6, p, None, None, p, outs, None, None) == -1      # C % 64 != 0
    assert L.rwkvtts_tmix_shift_mix_forward(1, 4, 64, 3, p, None, None, p, outs, None, None) == -1       # n must be 1 or 6
    assert L.rwkvtts_tmix_shift_mix_forward(1, 4, 64, 6, None, None, None, p, outs, None, None) == -2
    assert L.rwkvtts_tmix_shift_mix_forward(1, 4, 64, 6, p, None, p, p, outs, p, None) == -1             # in-place state: T == 1 only
    assert L.rwkvtts_add_layernorm_forward(8, 100, p, None, p, None, 1e-5, p, None, None, None) == -1    # C % 256 != 0
    assert L.rwkvtts_add_layernorm_forward(8, 256, p, None, None, None, 1e-5, p, None, None, None) == -2
    assert L.rwkvtts_sqrelu_forward(12, p, p, None) == -1                                                 # n % 8 != 0
    assert L.rwkvtts_tmix_out_forward(1, 4, 64, p, p, p, p, p, p, p, None, 1e-5, p, None) == -2           # ln_b missing
    assert L.rwkvtts_tmix_scratch_floats(1, 4, 100, 5) == 0
    # host side: CPU tensors never reach the fused kernels (the ATen formulation serves the CPU oracle tests)
    x = torch.zeros(1, 4, 64, dtype=torch.bfloat16)
    assert not fused.usable(x) and not fused.ln_usable(x)
    with pytest.raises(_lib.RwkvttsError):
        fused.sqrelu(x)


def test_reference_schemas_registered_and_no_cpu_fallback(built):
    import rwkvtts_b200 as R
    s = str(torch.ops.wind_backstepping.forward.default._schema)
    assert "Tensor(a!) y" in s and "Tensor(c!) sa" in s
    assert "Tensor(f!) da" in str(torch.ops.wind_backstepping.backward.default._schema)
    for ns in ("wkv7s", "rwkv7_state_fwd_fp16"):
        assert "int B, int T, int C, int H" in str(getattr(torch.ops, ns).forward.default._schema)
    x = torch.zeros(1, 16, 64, dtype=torch.bfloat16)
    with pytest.raises((NotImplementedError, R._lib.RwkvttsError)):
        R.RUN_CUDA_RWKV7g(x, x, x, x, x, x)
    with pytest.raises(R._lib.RwkvttsError):
        R.wkv7_state_forward_(1, 16, 64, 1, torch.zeros(1, 64, 64), x, x, x, x, x, x, x.clone())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rwkvtts_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "liboracle" not in src and "oracle/" not in src, f


def test_decode_plan_entry_points_validate_on_the_host():
    """rwkvtts_decode_* (one-kernel decode step): dims are checked without touching a GPU."""
    import ctypes
    from rwkvtts_b200 import _lib
    from rwkvtts_b200.decode import DEC_NDIM
    L = _lib.lib()
    ok = (ctypes.c_int * DEC_NDIM)(32, 1024, 16, 24, 8193, 4096, 64, 64, 32, 128)
    offs = (ctypes.c_size_t * 3)()
    n = L.rwkvtts_decode_workspace_bytes(ok, offs)
    assert n > 32 * 8193 * 4 and 0 < offs[0] < n and 0 < offs[1] < n and 0 < offs[2] < n and len({offs[0], offs[1], offs[2]}) == 3
    for bad in [(33, 1024, 16, 24, 8193, 4096, 64, 64, 32, 128),      # more than 32 rows
                (32, 1024, 15, 24, 8193, 4096, 64, 64, 32, 128),      # C != H * 64
                (32, 4096, 64, 24, 8193, 16384, 64, 64, 32, 128),     # hidden size beyond the kernel
                (32, 1024, 16, 24, 8193, 4096, 64, 64, 16, 128),      # a LoRA rank that is not a multiple of 32
                (32, 1024, 16, 24, 8193, 5000, 64, 64, 32, 128)]:     # channel-mix width not a multiple of C
        assert L.rwkvtts_decode_workspace_bytes((ctypes.c_int * DEC_NDIM)(*bad), None) == 0, bad
    assert L.rwkvtts_decode_init(None, None, None, None, None, 0, None) == -2          # RWKVTTS_ERR_NULL
    assert L.rwkvtts_decode_step(None, None, None, 1, 0, None, 0, 0, None) == -2
    assert L.rwkvtts_decode_release(None) == 0


def test_chunked_kernels_are_blackwell_native_in_the_shipped_sass(built):
    """What the shipped library's machine code contains (no GPU needed: cuobjdump reads the .so): tcgen05 MMAs with
    tensor-memory loads / stores in both chunked kernels, and the backward's inputs staged by tensor-map copies
    (cp.async.bulk.tensor -> UTMALDG; north_star: "TMA-staged (B,T,H,D) tiles"), not by per-thread copies."""
    import shutil
    import subprocess
    from rwkvtts_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.lib()._name], capture_output=True, text=True, check=True).stdout
    per = {}
    fn = None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            per[fn] = {"UTCHMMA": 0, "LDTM": 0, "UTMALDG": 0, "LDGSTS": 0}
        elif fn is not None:
            for m in per[fn]:
                if m in line:
                    per[fn][m] += 1
    fwd = [v for k, v in per.items() if "wkv7_tc_fwd_kernel" in k]
    bwd = [v for k, v in per.items() if "wkv7_tc_bwd_kernel" in k]
    assert len(fwd) == 4 and len(bwd) == 2
    assert all(v["UTCHMMA"] >= 16 and v["LDTM"] >= 5 for v in fwd + bwd)
    assert all(v["UTMALDG"] == 14 and v["LDGSTS"] == 0 for v in bwd)      # 7 tiles, issued in the prologue and in the loop
