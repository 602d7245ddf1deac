"""First GPU contact of the forward-v2 sketch (proto/wkv7_tc_fwd_v2.cu): build it into its own
shared object, compare with the f64 oracle at growing sizes, then time it at config c2 next to the shipped forward.
The kernel has never run: call this under a watchdog, e.g.  timeout 120 python scripts/check_tc_fwd_v2.py
(a mis-synchronised mbarrier pipeline hangs rather than fails).  --build-only compiles without touching the GPU."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "proto", "_build")
LIB = os.path.join(OUT, "libfwd_v2.so")


def build():
    os.makedirs(OUT, exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "rwkvtts_b200", "csrc"),
           os.path.join(ROOT, "proto", "wkv7_tc_fwd_v2.cu"), os.path.join(ROOT, "proto", "wkv7_tc_fwd_v2_capi.cu"), "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


def main():
    build()
    if "--build-only" in sys.argv:
        print("built", LIB)
        return
    import torch
    from oracle import wkv7_oracle as O
    from rwkvtts_b200 import ops
    lib = ctypes.CDLL(LIB)
    lib.fwd_v2.restype = ctypes.c_int
    lib.fwd_v2.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 12
    ORDER = "wqkvab"

    def fwd(d, y, s0=None, sT=None, ck=None, sa=None):
        B, T, H, _ = d["w"].shape
        rc = lib.fwd_v2(B, T, H, *[d[n].data_ptr() for n in ORDER], y.data_ptr(),
                        None if ck is None else ck.data_ptr(), None if sa is None else sa.data_ptr(),
                        None if s0 is None else s0.data_ptr(), None if sT is None else sT.data_ptr(),
                        torch.cuda.current_stream().cuda_stream)
        assert rc == 0, f"launch failed: cudaError {rc}"

    worst = 0.0
    for (B, T, H, with_s0) in [(1, 16, 1, False), (1, 32, 1, False), (1, 64, 1, False), (1, 80, 2, False), (2, 96, 2, True),
                               (2, 512, 12, False), (1, 1024, 4, False)]:
        x = O.make_inputs(B, T, H, seed=B * 1000 + T)
        d = {n: t.cuda() for n, t in x.items()}
        s0 = torch.randn(B, H, 64, 64) * 0.1 if with_s0 else None
        y = torch.empty_like(d["v"])
        sT = torch.empty(B, H, 64, 64, dtype=torch.float32, device="cuda")
        fwd(d, y, None if s0 is None else s0.cuda(), sT)
        torch.cuda.synchronize()
        y64, sT64 = O.wkv7_forward(*[x[n] for n in ORDER], s0=s0)
        exc, err, floor = O.excess_rel_l2(y.cpu(), y64)
        print(f"v2 B{B} T{T} H{H} s0={with_s0}: y excess {exc:.2e} (err {err:.2e}, bf16 floor {floor:.2e}); "
              f"sT rel {O.rel_l2(sT.cpu(), sT64):.2e}; nan={bool(torch.isnan(y.float()).any())}", flush=True)
        worst = max(worst, exc)
    print("worst excess", worst, "(bar 1e-3)")

    # training variant: its checkpoints and `sa` must drive the SHIPPED backward to the oracle's gradients
    import rwkvtts_b200 as R
    for (B, T, H) in [(1, 64, 1), (2, 208, 2), (1, 1024, 4)]:
        x = O.make_inputs(B, T, H, seed=7 * B + T)
        d = {n: t.cuda() for n, t in x.items()}
        y = torch.empty_like(d["v"])
        ck = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
        sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
        fwd(d, y, ck=ck, sa=sa)
        grads = [torch.empty_like(d["v"]) for _ in range(6)]
        R.wkv7_backward_(*[d[n] for n in ORDER], d["dy"], ck, sa, *grads)
        torch.cuda.synchronize()
        g64 = O.wkv7_backward(*[x[n] for n in ORDER], x["dy"])
        y64 = O.wkv7_forward(*[x[n] for n in ORDER])
        y64 = y64[0] if isinstance(y64, tuple) else y64
        errs = {"y": O.excess_rel_l2(y.cpu(), y64)[0]}
        for i, n in enumerate(ORDER):
            errs["d" + n] = O.excess_rel_l2(grads[i].cpu(), g64[i])[0]
        print(f"v2 training B{B} T{T} H{H}: " + "  ".join(f"{n} {e:.1e}" for n, e in errs.items()), flush=True)

    B, T, H = 8, 4096, 16
    xs = O.make_inputs(1, T, H, seed=3)
    d = {n: t.cuda().repeat(B, 1, 1, 1).contiguous() for n, t in xs.items()}
    y = torch.empty_like(d["v"])

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for name, fn in (("v2 sketch", lambda: fwd(d, y)), ("shipped no-grad forward", lambda: ops.wkv7_forward_infer_(*[d[n] for n in ORDER], y))):
        ms = timed(fn)
        print(f"{name}: {ms:.3f} ms at c2 [8,4096,16,64] -> {B * T * H * 896 / ms / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
