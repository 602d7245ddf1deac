"""CPU prototype of the chunked WKV-7 backward in the WINDOW frame, written as the exact list of
matrix products the tcgen05 kernel issues (design aid; validates the algebra against the oracle)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import wkv7_oracle as O
from proto.chunk_fwd_proto import MM

L, WIN = 16, 4

def prep(w, q, k, v, a, b, c, dt, mm):
    """operands of chunk c in its window's frame"""
    c0 = (c // WIN) * WIN
    lw_all = torch.clamp(-torch.exp(w[c0 * L:(c + 1) * L]), min=-1.35)
    G = torch.cumsum(lw_all, 0)[(c - c0) * L:]
    sl = slice(c * L, (c + 1) * L)
    lw = lw_all[(c - c0) * L:]
    E, Ep = torch.exp(G), torch.exp(G - lw)
    At, Bt, Kt, Qt = a[sl] * Ep, b[sl] / E, k[sl] / E, q[sl] * E
    ts = torch.tril(torch.ones(L, L, dtype=dt), -1); ti = torch.tril(torch.ones(L, L, dtype=dt))
    N, Aak, Aqb, Aqk = mm(At, Bt.T) * ts, mm(At, Kt.T) * ts, mm(Qt, Bt.T) * ti, mm(Qt, Kt.T) * ti
    Tm = torch.linalg.inv(torch.eye(L, dtype=dt) - N)
    return dict(sl=sl, lw=lw, G=G, E=E, Ep=Ep, At=At, Bt=Bt, Kt=Kt, Qt=Qt, N=N, Aak=Aak, Aqb=Aqb, Aqk=Aqk, Tm=Tm,
                W=mm(Tm, At), M1=mm(Tm, Aak), Bp=mm(Tm.T, Bt), Aqbp=mm(Aqb, Tm), ts=ts, ti=ti)

def fwd_ckpt(w, q, k, v, a, b, mode='f64', s0=None):
    dt = torch.float64 if mode == 'f64' else torch.float32
    mm = MM(mode)
    w, q, k, v, a, b = [x.to(dt) for x in (w, q, k, v, a, b)]
    T, C = w.shape; nC = T // L
    S = torch.zeros(C, C, dtype=dt) if s0 is None else s0.to(dt).clone()
    ck, y = [], torch.empty(T, C, dtype=dt)
    for c in range(nC):
        P = prep(w, q, k, v, a, b, c, dt, mm); V = v[P['sl']]
        ck.append(S.clone())
        U = mm(P['W'], S.T) + mm(P['M1'], V)
        y[P['sl']] = mm(P['Qt'], S.T) + mm(P['Aqb'], U) + mm(P['Aqk'], V)
        S = S + mm(U.T, P['Bt']) + mm(V.T, P['Kt'])
        if c % WIN == WIN - 1 or c == nC - 1:
            S = S * P['E'][-1]
    return y, S, ck

def bwd(w, q, k, v, a, b, dy, ck, sT, mode='f64', dsT=None):
    dt = torch.float64 if mode == 'f64' else torch.float32
    mm = MM(mode)
    w, q, k, v, a, b, dy = [x.to(dt) for x in (w, q, k, v, a, b, dy)]
    T, C = w.shape; nC = T // L
    outs = {n: torch.empty(T, C, dtype=dt) for n in 'wqkvab'}
    dS_next = torch.zeros(C, C, dtype=dt) if dsT is None else dsT.to(dt).clone()   # grad wrt state at START of chunk c+1 (its own frame)
    suffix = torch.zeros(C, dtype=dt); carry_first = torch.zeros(C, dtype=dt)
    for c in range(nC - 1, -1, -1):
        P = prep(w, q, k, v, a, b, c, dt, mm); V, dY = v[P['sl']], dy[P['sl']]
        win_end = (c % WIN == WIN - 1) or (c == nC - 1)
        S0 = ck[c]
        gL = torch.zeros(C, dtype=dt)
        if win_end:
            S_next = sT.to(dt) if c == nC - 1 else ck[c + 1]
            gL = (dS_next * S_next).sum(0)               # boundary term of the rescale
            dSe = dS_next * P['E'][-1]                  # into this window's frame
            suffix = torch.zeros(C, dtype=dt); carry_first = torch.zeros(C, dtype=dt)
        else:
            dSe = dS_next
        U = mm(P['W'], S0.T) + mm(P['M1'], V)                                    # (F)
        Z = mm(P['Aqbp'].T, dY) + mm(P['Bp'], dSe.T)                             # (R1)
        dN, dAak = mm(Z, U.T) * P['ts'], mm(Z, V.T) * P['ts']                    # (G)
        dAqb, dAqk = mm(dY, U.T) * P['ti'], mm(dY, V.T) * P['ti']
        dV = mm(P['Aqk'].T, dY) + mm(P['Kt'], dSe.T) + mm(P['Aak'].T, Z)          # (P3)
        dAt = mm(Z, S0) + mm(dN, P['Bt']) + mm(dAak, P['Kt'])                     # (P1)
        dQt = mm(dY, S0) + mm(dAqb, P['Bt']) + mm(dAqk, P['Kt'])
        dBt = mm(U, dSe) + mm(dN.T, P['At']) + mm(dAqb.T, P['Qt'])                # (P2)
        dKt = mm(V, dSe) + mm(dAak.T, P['At']) + mm(dAqk.T, P['Qt'])
        dS_next = dSe + mm(dY.T, P['Qt']) + mm(Z.T, P['At'])                     # (R2)
        sl = P['sl']
        outs['a'][sl] = dAt * P['Ep']; outs['b'][sl] = dBt / P['E']; outs['k'][sl] = dKt / P['E']
        outs['q'][sl] = dQt * P['E']; outs['v'][sl] = dV
        g = dQt * P['Qt'] - dKt * P['Kt'] - dBt * P['Bt']
        aa = dAt * P['At']
        g[:-1] += aa[1:]
        g[-1] += carry_first + gL
        carry_first = aa[0]
        suf = torch.flip(torch.cumsum(torch.flip(g, [0]), 0), [0]) + suffix
        suffix = suf[0]
        outs['w'][sl] = suf * P['lw']
    return outs, dS_next

if __name__ == '__main__':
    B, T, H = 1, 208, 2
    x = O.make_inputs(B, T, H, seed=5)
    names = 'wqkvab'
    s0 = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    dsT = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    g64 = O.wkv7_backward(*[x[n] for n in names], x['dy'], s0=s0, dsT=dsT)
    y64, sT64 = O.wkv7_forward(*[x[n] for n in names], s0=s0)
    for mode in ['f64', 'tf32']:
        worst = {}
        for h in range(H):
            xs = [x[n][0, :, h] for n in names]
            y, sT, ck = fwd_ckpt(*xs, mode=mode, s0=s0[0, h])
            outs, dS0 = bwd(*xs, x['dy'][0, :, h], ck, sT, mode=mode, dsT=dsT[0, h])
            e = O.rel_l2(y, y64[0, :, h]); worst['y'] = max(worst.get('y', 0), e)
            for i, n in enumerate(names):
                e = O.excess_rel_l2(outs[n].to(torch.bfloat16), g64[i][0, :, h])[0] if mode != 'f64' else O.rel_l2(outs[n], g64[i][0, :, h])
                worst[n] = max(worst.get(n, 0), e)
            worst['s0'] = max(worst.get('s0', 0), O.rel_l2(dS0, g64[6][0, h]))
        print(f"{mode:5s}: " + "  ".join(f"{n} {e:.1e}" for n, e in worst.items()))
