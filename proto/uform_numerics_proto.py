"""Numerics of the U-form forward ("forward v2") against today's W~-form, emulating what the tensor
core sees: operands staged in shared memory are rounded to tf32, operands read from tensor memory (the state, the
phase-1 accumulator, U) are TRUNCATED to tf32, accumulation in fp32.  Truth: the f64 oracle.  The W~-form line reproduces
the shipped kernel's measured error (3.9e-4 excess rel-L2 of y), which validates the emulation; the U-form adds one
more truncated operand (the phase-1 accumulator multiplied by T) and costs nothing measurable."""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'proto'))
from oracle import wkv7_oracle as O
from chunk_fwd_proto import rnd_tf32, trunc_tf32

def mmk(A, B, tA=False, tB=False):
    fa = trunc_tf32 if tA else rnd_tf32
    fb = trunc_tf32 if tB else rnd_tf32
    return (fa(A).double() @ fb(B).double()).float()

def chunk_fwd(w, q, k, v, a, b, L=16, form='U', WIN=4):
    dt = torch.float32
    w, q, k, v, a, b = [x.to(dt) for x in (w, q, k, v, a, b)]
    T, C = w.shape
    S = torch.zeros(C, C, dtype=dt)
    ys = []
    tril_s = torch.tril(torch.ones(L, L, dtype=dt), -1); tril_i = torch.tril(torch.ones(L, L, dtype=dt))
    for c0 in range(0, T, L):
        sl = slice(c0, c0 + L)
        lw = -torch.exp(w[sl].double())
        g = torch.cumsum(lw, 0)
        D = torch.exp(g).float(); Dprev = torch.exp(g - lw).float(); iD = torch.exp(-g).float()
        At, Bt, Kt, Qt = a[sl] * Dprev, b[sl] * iD, k[sl] * iD, q[sl] * D
        Aab = mmk(At, Bt.T) * tril_s; Aak = mmk(At, Kt.T) * tril_s
        Aqb = mmk(Qt, Bt.T) * tril_i; Aqk = mmk(Qt, Kt.T) * tril_i
        Tm = torch.linalg.solve_triangular(torch.eye(L, dtype=dt) - Aab, torch.eye(L, dtype=dt), upper=False)
        if form == 'U':      # U = T (A~ S^T + Aak V): R from TMEM (truncated), T rounded
            R = mmk(At, S.T, tB=True) + mmk(Aak, v[sl])
            U = mmk(Tm, R, tB=True)
        else:                # W-form (today's kernel): W~ = T A~, M1 = T Aak in fp32 on CUDA cores, rounded operands
            Wt = Tm @ rnd_tf32(At); M1 = Tm @ Aak
            U = mmk(Wt, S.T, tB=True) + mmk(M1, v[sl])
        Y = mmk(Qt, S.T, tB=True) + mmk(Aqb, U, tB=True) + mmk(Aqk, v[sl])
        S = (S + mmk(U.T, Bt, tA=True) + mmk(v[sl].T, Kt)) * D[-1]
        ys.append(Y)
    return torch.cat(ys), S

x = O.make_inputs(1, 1024, 2, seed=3)
names = 'wqkvab'
y64, s64 = O.wkv7_forward(*[x[n] for n in names])
for form in ('W', 'U'):
    errs = []
    for h in range(2):
        y, S = chunk_fwd(*[x[n][0, :, h] for n in names], form=form)
        errs.append((O.rel_l2(y, y64[0, :, h]), O.excess_rel_l2(y.to(torch.bfloat16), y64[0, :, h])[0], O.rel_l2(S, s64[0, h])))
    print(f"{form}-form  y rel-l2 {max(e[0] for e in errs):.2e}  bf16-excess {max(e[1] for e in errs):.2e}  S_T {max(e[2] for e in errs):.2e}")
