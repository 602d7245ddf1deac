"""GPU check of one forward kernel family against the oracle, plus timing.
usage: python scripts/check_tc_fwd.py [impl] (0 scan, 1 tcgen05)"""
import sys, time
import torch
import rwkvtts_b200 as R
from rwkvtts_b200 import ops
from oracle import wkv7_oracle as O

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = R._lib.lib()
assert lib.rwkvtts_set_impl(impl) == 0
ORDER = "wqkvab"

def run(B, T, H, seed, s0=None):
    x = O.make_inputs(B, T, H, seed=seed)
    d = {n: t.cuda() for n, t in x.items()}
    y = torch.empty_like(d["v"])
    sT = torch.empty(B, H, 64, 64, dtype=torch.float32, device="cuda")
    s0d = None if s0 is None else s0.cuda()
    ops.wkv7_forward_infer_(*[d[n] for n in ORDER], y, s0=s0d, sT=sT)
    torch.cuda.synchronize()
    y64, sT64 = O.wkv7_forward(*[x[n] for n in ORDER], s0=s0)
    exc, err, floor = O.excess_rel_l2(y.cpu(), y64)
    es = O.rel_l2(sT.cpu(), sT64)
    print(f"impl {impl} B{B} T{T} H{H}: y excess {exc:.2e} err {err:.2e} floor {floor:.2e}; sT rel {es:.2e}"
          f"  nan={bool(torch.isnan(y.float()).any())}", flush=True)
    return exc

worst = 0
for (B, T, H) in [(1, 16, 1), (1, 64, 1), (1, 80, 2), (2, 512, 12), (1, 1024, 4)]:
    worst = max(worst, run(B, T, H, seed=B * 1000 + T))
worst = max(worst, run(2, 96, 2, 9, s0=torch.randn(2, 2, 64, 64) * 0.1))
print("worst excess", worst)

# timing at c2
B, T, H = 8, 4096, 16
x = O.make_inputs(B, T, H, seed=1) if False else None
g = torch.Generator(device="cuda").manual_seed(0)
from rwkvtts_b200 import synth
d = synth.make_inputs_cuda(B, T, H, seed=0) if hasattr(synth, "make_inputs_cuda") else None
if d is None:
    xs = O.make_inputs(1, T, H, seed=3)
    d = {n: t.cuda().repeat(B, 1, 1, 1).contiguous() for n, t in xs.items()}
y = torch.empty_like(d["v"])
s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
for _ in range(3):
    ops.wkv7_forward_infer_(*[d[n] for n in ORDER], y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.wkv7_forward_infer_(*[d[n] for n in ORDER], y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"impl {impl} fwd c2 [8,4096,16,64]: {ms:.3f} ms  -> {B*T*H*896/ms/1e6:.0f} GB/s algorithmic "
      f"({B*T*H*896/ms/1e6/6550.4*100:.1f}% of 6550 GB/s)")
