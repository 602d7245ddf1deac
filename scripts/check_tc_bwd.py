"""GPU check of the tcgen05 training pair (forward with checkpoints + chunked backward) against the oracle,
plus timing at config c2."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkvtts_b200 as R
from rwkvtts_b200 import ops
from oracle import wkv7_oracle as O

lib = R._lib.lib()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
assert lib.rwkvtts_set_impl(impl) == 0
ORDER = "wqkvab"

def run(B, T, H, seed, with_state=False):
    x = O.make_inputs(B, T, H, seed=seed)
    d = {n: t.cuda() for n, t in x.items()}
    leaves = [d[n].clone().requires_grad_(True) for n in ORDER]
    s0 = dsT = None
    if with_state:
        s0 = torch.randn(B, H, 64, 64) * 0.1
        dsT = torch.randn(B, H, 64, 64) * 0.1
        s0d = s0.cuda().requires_grad_(True)
        y, sT = R.wkv7_with_state(*leaves, s0d)
        torch.autograd.backward([y, sT], [d["dy"], dsT.cuda()])
    else:
        y = R.WindBackstepping.apply(*leaves)
        y.backward(d["dy"])
    torch.cuda.synchronize()
    y64, _ = O.wkv7_forward(*[x[n] for n in ORDER], s0=s0)
    g64 = O.wkv7_backward(*[x[n] for n in ORDER], x["dy"], s0=s0, dsT=dsT)
    msg = [f"y {O.excess_rel_l2(y.cpu(), y64)[0]:.1e}"]
    worst = 0
    for n, leaf, g in zip(ORDER, leaves, g64):
        e = O.excess_rel_l2(leaf.grad.cpu(), g)[0]
        bad = not bool(torch.isfinite(leaf.grad.float()).all())
        msg.append(f"d{n} {e:.1e}{'(NaN)' if bad else ''}")
        worst = max(worst, e)
    if with_state:
        msg.append(f"ds0 {O.rel_l2(s0d.grad.cpu(), g64[6]):.1e}")
    print(f"impl {impl} B{B} T{T} H{H} state={with_state}: " + "  ".join(msg), flush=True)
    return worst

worst = 0
for (B, T, H) in [(1, 16, 1), (1, 32, 1), (1, 64, 1), (1, 80, 2), (1, 208, 2), (2, 512, 12)]:
    worst = max(worst, run(B, T, H, seed=B * 1000 + T))
worst = max(worst, run(2, 96, 2, 9, with_state=True))
print("worst excess", worst)

B, T, H = 8, 4096, 16
xs = O.make_inputs(1, T, H, seed=3)
d = {n: t.cuda().repeat(B, 1, 1, 1).contiguous() for n, t in xs.items()}
ins = [d[n] for n in ORDER]
y = torch.empty_like(d["v"])
s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
grads = [torch.empty_like(d["v"]) for _ in range(6)]
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f = timed(lambda: ops.wkv7_forward_(*ins, y, s, sa))
b = timed(lambda: ops.wkv7_backward_(*ins, d["dy"], s, sa, *grads))
th = B * T * H
print(f"impl {impl} c2: train fwd {f:.3f} ms ({th*896/f/1e6:.0f} GB/s, {th*896/f/1e6/65.504:.1f}%)  "
      f"bwd {b:.3f} ms ({th*1664/b/1e6:.0f} GB/s, {th*1664/b/1e6/65.504:.1f}%)  step x24 = {(f+b)*24:.1f} ms "
      f"-> {B*T/((f+b)*24e-3):.0f} tokens/s")
