"""Numerical experiment for DESIGN.md section 7 item 1: can T = (I - N)^-1 of a 16-token chunk be formed on the tensor
cores by doubling, T = (I+N)(I+N^2)(I+N^4)(I+N^8), in tf32 (operands rounded to 10 mantissa bits, fp32 accumulate), instead
of the fp32 column solve of stage B?  Inputs: the synthetic op-level inputs of the bench (rwkvtts_b200.synth), N built as
the kernels build it (window frame, decays accumulated since the window start).  Reports the relative error of
W~ = T A~ against the f64 solve for: fp32 solve + one tf32 rounding of the result (today), tf32 doubling, and
"3xtf32" doubling (operands split hi + lo)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rwkvtts_b200.synth import make_inputs


def tf32(x):
    y = x.float().contiguous().view(torch.int32)
    return ((y + 0x1000) & ~0x1FFF).view(torch.float32).double()


def mm_tf32(a, b):
    return (tf32(a) @ tf32(b)).float().double()          # products exact in fp32 accumulate up to fp32 rounding


def mm_3x(a, b):
    ah, bh = tf32(a), tf32(b)
    al, bl = tf32(a - ah), tf32(b - bh)
    return (ah @ bh + ah @ bl + al @ bh).float().double()


def doubling(N, mm):
    L = N.shape[-1]
    I = torch.eye(L, dtype=torch.float64).expand_as(N)
    T = I + N
    P = N
    for _ in range(L.bit_length() - 2):                 # 3 squarings for L = 16, 5 for L = 64
        P = mm(P, P)
        T = T + mm(T, P)
    return T


x = make_inputs(2, 256, 4, seed=1)
w, a, b = (x[n].double() for n in "wab")
lw = (-torch.exp(w)).clamp(min=-1.35)                    # log decay per step
B, T, H, C = w.shape


def run(L):
    res = {"fp32 solve + tf32 round": [], "tf32 doubling": [], "3xtf32 doubling": []}
    for c0 in range(0, T, L):
        win0 = (c0 // 64) * 64
        G = lw[:, win0:c0 + L].cumsum(1)[:, c0 - win0:]      # [B,L,H,C] decay since the window start
        Gm1 = G - lw[:, c0:c0 + L]
        At = (a[:, c0:c0 + L] * torch.exp(Gm1)).permute(0, 2, 1, 3)      # [B,H,L,C]
        Bt = (b[:, c0:c0 + L] * torch.exp(-G)).permute(0, 2, 1, 3)
        N = torch.tril(tf32(At) @ tf32(Bt).transpose(-1, -2), diagonal=-1)
        I = torch.eye(L, dtype=torch.float64)
        W_ref = torch.linalg.solve_triangular(I - N, tf32(At), upper=False)
        W_now = tf32(torch.linalg.solve_triangular((I - N).float(), tf32(At).float(), upper=False))
        rel = lambda y: float((y - W_ref).norm() / W_ref.norm())
        res["fp32 solve + tf32 round"].append(rel(W_now))
        res["tf32 doubling"].append(rel(tf32(mm_tf32(doubling(N, mm_tf32), At))))
        res["3xtf32 doubling"].append(rel(tf32(mm_tf32(doubling(N, mm_3x), At))))
    for k, v in res.items():
        t = torch.tensor(v)
        print(f"L={L:3d}  {k:28s} rel err of W~: mean {t.mean():.2e}  max {t.max():.2e}")


run(16)      # today's chunk
run(64)      # a whole window as one chunk: every product is a single M = 64, N = 64, K = 64 tensor-core instruction
