"""Runs the tcgen05 training pair and the snapshot-free forward at config c2 a few times (target of the
ncu captures: `ncu ... python scripts/run_pair.py [reps]`).  Prints CUDA-event times when not profiled."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwkvtts_b200 as R
from rwkvtts_b200 import ops
from rwkvtts_b200.synth import make_inputs

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T, H = 8, 4096, 16
x = make_inputs(B, T, H, seed=42)
d = {n: t.cuda() for n, t in x.items()}
ins = [d[n] for n in "wqkvab"]
y = torch.empty_like(d["v"])
s = torch.empty(B, H, T // 16, 64, 64, dtype=torch.float32, device="cuda")
sa = torch.empty(B, T, H, 64, dtype=torch.float32, device="cuda")
grads = [torch.empty_like(d["v"]) for _ in range(6)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    ops.wkv7_forward_(*ins, y, s, sa)
    ev[1].record()
    ops.wkv7_backward_(*ins, d["dy"], s, sa, *grads)
    ev[2].record()
    ops.wkv7_forward_infer_(*ins, y)
    ev[3].record()
torch.cuda.synchronize()
th = B * T * H
f, b, i = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
print(f"train fwd {f:.3f} ms ({th*896/f/1e6:.0f} GB/s)  bwd {b:.3f} ms ({th*1664/b/1e6:.0f} GB/s)  "
      f"infer fwd {i:.3f} ms ({th*896/i/1e6:.0f} GB/s)")
