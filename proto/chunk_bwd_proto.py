"""CPU prototype of the chunked WKV-7 backward (design aid; validates the per-chunk adjoint algebra)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import wkv7_oracle as O
from proto.chunk_fwd_proto import MM

def chunk_bwd(w, q, k, v, a, b, dy, L=16, mode='f64', s0=None, dsT=None):
    dt = torch.float64 if mode == 'f64' else torch.float32
    mm = MM(mode)
    w, q, k, v, a, b, dy = [x.to(dt) for x in (w, q, k, v, a, b, dy)]
    T, C = w.shape
    nC = T // L
    eye = torch.eye(L, dtype=dt)
    ts = torch.tril(torch.ones(L, L, dtype=dt), -1); ti = torch.tril(torch.ones(L, L, dtype=dt))
    # forward pass over chunks: keep chunk-start states (the kernel recomputes them from sparse checkpoints)
    S = torch.zeros(C, C, dtype=dt) if s0 is None else s0.to(dt).clone()
    S0s = []
    def prep(c):
        sl = slice(c * L, (c + 1) * L)
        lw = -torch.exp(w[sl]); g = torch.cumsum(lw, 0); D = torch.exp(g); Dp = torch.exp(g - lw)
        At, Bt, Kt, Qt = a[sl] * Dp, b[sl] / D, k[sl] / D, q[sl] * D
        N = mm(At, Bt.T) * ts; Aak = mm(At, Kt.T) * ts; Aqb = mm(Qt, Bt.T) * ti; Aqk = mm(Qt, Kt.T) * ti
        return sl, lw, D, Dp, At, Bt, Kt, Qt, N, Aak, Aqb, Aqk
    for c in range(nC):
        sl, lw, D, Dp, At, Bt, Kt, Qt, N, Aak, Aqb, Aqk = prep(c)
        S0s.append(S.clone())
        R = mm(At, S.T) + mm(Aak, v[sl])
        U = torch.linalg.solve_triangular(eye - N, R, upper=False)
        S = (S + mm(U.T, Bt) + mm(v[sl].T, Kt)) * D[-1]
    dS = torch.zeros(C, C, dtype=dt) if dsT is None else dsT.to(dt).clone()
    outs = {n: torch.empty(T, C, dtype=dt) for n in 'wqkvab'}
    for c in range(nC - 1, -1, -1):
        sl, lw, D, Dp, At, Bt, Kt, Qt, N, Aak, Aqb, Aqk = prep(c)
        S0 = S0s[c]; V = v[sl]; dY = dy[sl]
        R = mm(At, S0.T) + mm(Aak, V)
        U = torch.linalg.solve_triangular(eye - N, R, upper=False)
        Shat = S0 + mm(U.T, Bt) + mm(V.T, Kt)
        dShat = dS * D[-1]
        dgL = (dS * Shat).sum(0) * D[-1]                       # = sum_i dS_L ⊙ S_L
        dU = mm(Aqb.T, dY) + mm(Bt, dShat.T)
        dR = torch.linalg.solve_triangular((eye - N).T, dU, upper=True)
        dN = mm(dR, U.T) * ts; dAak = mm(dR, V.T) * ts; dAqb = mm(dY, U.T) * ti; dAqk = mm(dY, V.T) * ti
        dV = mm(Aqk.T, dY) + mm(Kt, dShat.T) + mm(Aak.T, dR)
        dAt = mm(dR, S0) + mm(dN, Bt) + mm(dAak, Kt)
        dBt = mm(U, dShat) + mm(dN.T, At) + mm(dAqb.T, Qt)
        dKt = mm(V, dShat) + mm(dAak.T, At) + mm(dAqk.T, Qt)
        dQt = mm(dY, S0) + mm(dAqb, Bt) + mm(dAqk, Kt)
        dS = dShat + mm(dY.T, Qt) + mm(dR.T, At)
        outs['a'][sl] = dAt * Dp; outs['b'][sl] = dBt / D; outs['k'][sl] = dKt / D; outs['q'][sl] = dQt * D
        outs['v'][sl] = dV
        dg = -dBt * Bt - dKt * Kt + dQt * Qt
        dg[:-1] += (dAt * At)[1:]
        dg[-1] += dgL
        dlw = torch.flip(torch.cumsum(torch.flip(dg, [0]), 0), [0])
        outs['w'][sl] = dlw * lw
    return outs, dS

if __name__ == '__main__':
    B, T, H = 1, 512, 2
    x = O.make_inputs(B, T, H, seed=5)
    names = 'wqkvab'
    s0 = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    dsT = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    g64 = O.wkv7_backward(*[x[n] for n in names], x['dy'], s0=s0, dsT=dsT)
    for mode in ['f64', 'f32', 'tf32', 'bf16x3']:
        for L in (16, 32):
            worst = {}
            for h in range(H):
                outs, dS0 = chunk_bwd(*[x[n][0, :, h] for n in names], x['dy'][0, :, h], L=L, mode=mode, s0=s0[0, h], dsT=dsT[0, h])
                for i, n in enumerate(names):
                    e = O.excess_rel_l2(outs[n].to(torch.bfloat16), g64[i][0, :, h])[0] if mode != 'f64' else O.rel_l2(outs[n], g64[i][0, :, h])
                    worst[n] = max(worst.get(n, 0), e)
                worst['s0'] = max(worst.get('s0', 0), O.rel_l2(dS0, g64[6][0, h]))
            print(f"{mode:7s} L={L}: " + "  ".join(f"d{n} {e:.1e}" for n, e in worst.items()))
