"""Runs the fused time-mix kernels (forward + backward) once at the c2 activation shape: target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkvtts_b200  # noqa
from rwkvtts_b200 import fused as FU
B, T, C, H = 8, 4096, 1024, 16
dev = "cuda"
act = lambda: torch.randn(B, T, C, device=dev).bfloat16().requires_grad_(True)
par = lambda *sh: (0.5 * torch.randn(*sh, device=dev)).bfloat16().requires_grad_(True)
x = act(); mixes = [par(1, 1, C) for _ in range(6)]
k, v, wl, al, vl, vf = (act() for _ in range(6)); pp = [par(1, 1, C) for _ in range(5)]
y, r, g = (act() for _ in range(3)); rk, lw, lb = par(H, 64), par(C), par(C)
do = [torch.randn(B, T, C, device=dev).bfloat16() for _ in range(6)]
lx, lres, lnw, lnb = act(), act(), par(C), par(C)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    ev[0].record(); o1 = FU.shift_mix(x, mixes); ev[1].record(); torch.autograd.backward(o1, do); ev[2].record()
    o2 = FU.prep(k, v, wl, al, vl, vf, *pp); ev[3].record(); torch.autograd.backward(o2, do[:5]); ev[4].record()
    o3 = FU.out(y, r, k, v, g, rk, lw, lb, 64e-5); ev[5].record(); o3.backward(do[0]); ev[6].record()
    o4 = FU.add_layernorm(lx, lres, lnw, lnb, 1e-5); ev[7].record(); torch.autograd.backward(o4, do[:2]); ev[8].record()
    torch.cuda.synchronize()
    print(" ".join(f"{ev[i].elapsed_time(ev[i+1]):.3f}" for i in range(8)),
          "ms: mix fwd/bwd, prep fwd/bwd, out fwd/bwd, add+LayerNorm fwd/bwd")
