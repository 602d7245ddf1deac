"""CPU prototype of the chunked WKV-7 forward (design aid; not shipped, not the oracle).
Validates the algebra in f64 and emulates tensor-core operand roundings to size the numerics."""
import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import wkv7_oracle as O

def rnd_tf32(x):   # round-to-nearest to 10 explicit mantissa bits (cvt.rna.tf32.f32), fp32 container
    xi = x.float().contiguous().view(torch.int32)
    xi = (xi + 0x1000) & ~0x1FFF
    return xi.view(torch.float32)
def trunc_tf32(x):
    xi = x.float().contiguous().view(torch.int32) & ~0x1FFF
    return xi.view(torch.float32)
def rnd_bf16(x): return x.float().to(torch.bfloat16).float()

class MM:
    def __init__(s, mode): s.mode = mode
    def __call__(s, A, B, exactA=False, exactB=False):
        m = s.mode
        if m == 'f64': return A.double() @ B.double()
        if m == 'f32': return (A.float() @ B.float())
        if m == 'tf32': return (rnd_tf32(A).double() @ rnd_tf32(B).double()).float()
        if m == 'tf32t': return (trunc_tf32(A).double() @ trunc_tf32(B).double()).float()
        if m == 'bf16': return (rnd_bf16(A).double() @ rnd_bf16(B).double()).float()
        if m == 'bf16x3':
            Ah, Bh = rnd_bf16(A), rnd_bf16(B); Al, Bl = rnd_bf16(A.float()-Ah), rnd_bf16(B.float()-Bh)
            return (Ah.double()@Bh.double() + Ah.double()@Bl.double() + Al.double()@Bh.double()).float()
        raise ValueError(m)

def chunk_fwd(w, q, k, v, a, b, L=16, mode='f64', s0=None):
    """U-form. inputs [T,64] for one head (any float dtype). returns y [T,64], S_T [64,64] (value-major)."""
    dt = torch.float64 if mode == 'f64' else torch.float32
    mm = MM(mode)
    w, q, k, v, a, b = [x.to(dt) for x in (w, q, k, v, a, b)]
    T, C = w.shape
    S = torch.zeros(C, C, dtype=dt) if s0 is None else s0.to(dt).clone()
    ys = []
    tril_s = torch.tril(torch.ones(L, L, dtype=dt), -1); tril_i = torch.tril(torch.ones(L, L, dtype=dt))
    for c0 in range(0, T, L):
        sl = slice(c0, c0 + L)
        g = torch.cumsum(-torch.exp(w[sl]), 0)                 # log cumulative decay, [L,C]
        D = torch.exp(g); Dprev = torch.exp(g - (-torch.exp(w[sl])))   # D_{t-1}
        At, Bt, Kt, Qt = a[sl] * Dprev, b[sl] / D, k[sl] / D, q[sl] * D
        Aab = mm(At, Bt.T) * tril_s; Aak = mm(At, Kt.T) * tril_s
        Aqb = mm(Qt, Bt.T) * tril_i; Aqk = mm(Qt, Kt.T) * tril_i
        # T = (I - Aab)^-1 by forward substitution in fp32/64 on CUDA cores
        Tm = torch.linalg.solve_triangular(torch.eye(L, dtype=dt) - Aab, torch.eye(L, dtype=dt), upper=False)
        R = mm(At, S.T) + mm(Aak, v[sl])
        U = mm(Tm, R)
        Y = mm(Qt, S.T) + mm(Aqb, U) + mm(Aqk, v[sl])
        S = (S + mm(U.T, Bt) + mm(v[sl].T, Kt)) * D[-1]
        ys.append(Y)
    return torch.cat(ys), S

if __name__ == '__main__':
    x = O.make_inputs(1, 1024, 2, seed=3)
    names = 'wqkvab'
    y64, s64 = O.wkv7_forward(*[x[n] for n in names])
    for mode in ['f64', 'f32', 'tf32', 'tf32t', 'bf16x3', 'bf16']:
        for L in (16, 32, 64):
            errs = []
            for h in range(2):
                y, S = chunk_fwd(*[x[n][0, :, h] for n in names], L=L, mode=mode)
                errs.append((O.rel_l2(y, y64[0, :, h]), O.excess_rel_l2(y.to(torch.bfloat16), y64[0, :, h])[0], O.rel_l2(S, s64[0, h])))
            print(f"{mode:7s} L={L:3d}  y rel-l2 {max(e[0] for e in errs):.2e}  bf16-excess {max(e[1] for e in errs):.2e}  S_T {max(e[2] for e in errs):.2e}")
