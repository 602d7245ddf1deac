"""Executable blueprint of "backward v2" (DESIGN.md section 7): the product list of proto/tc_bwd_proto.py with the parts
that CHANGE written at operand-tile level -- canonical K-major tiles (tc05.cuh) filled with the offset expressions the
kernel would use, tensor-memory accumulators as [64 rows][columns] arrays, every tcgen05 chain as
D[m][n] (+)= sum_k A[m][k] B[n][k] -- so that the orientation of every new tile is pinned before any CUDA is written:

  B1   [R_B^T | dV^T] = dS [B~ ; K~]^T            one N = 32 chain (today: two N = 16 chains, B' instead of B~)
       R^T += dY^T Aqb                            B operand AqbT[n = t][k = s] = Aqb[s][t]
  B2   Z^T = R^T T                                B operand Tt[n = t][k = s] = T[s][t]   (Z-form: no back substitution)
  G1   [N | Aak ; Aqb | Aqk] = [A~ ; Q~] [B~ ; K~]^T     forward Gram blocks on the tensor core (M = 64, 32 live rows)
  G2   [dAqb | dAqk ; dN | dAak] = [dY ; Z] [U ; V]^T    gradient Gram blocks, operands = the DYZn / UVn tiles of today
  the Gram groups read G1 / G2 with rows on the lanes and write the 16x16 operand tiles of P1 / P2b / P3b.

Everything else (P1, P2, P3, R2, the output scaling and the dw scan) is unchanged and stays at matrix level here.
Checked in f64 against the oracle's gradients."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "proto"))
from oracle import wkv7_oracle as O  # noqa: E402
from fwd_v2_index_emulator import kmajor_off, smem_operand  # noqa: E402
import tc_bwd_proto as P0  # noqa: E402

L, WIN = 16, 4
N32_LBO, N_SBO = 132, 32          # [32 token rows][64 channels]
S16_LBO, S_SBO = 68, 32           # [16][16]   (68: conflict-free column stores, see the forward-v2 emulator)
S32_LBO = 132                     # [32][16]


def token_major(top, bottom):
    """[top ; bottom] (16 + 16 token rows x 64), written one 16-byte piece per (row, k-group) as stage A does"""
    t = np.full(16 * N32_LBO, np.nan)
    for half, X in enumerate((top, bottom)):
        for r in range(16):
            for k4 in range(16):
                o = ((16 * half + r) >> 3) * N_SBO + k4 * N32_LBO + (r & 7) * 4
                t[o:o + 4] = X[r, 4 * k4:4 * k4 + 4]
    return t


def rows32(tile):
    return smem_operand(tile, 32, 64, N32_LBO, N_SBO)


def tile16_rows(X):               # thread = row r writes X[r][:] with four 16-byte stores: tile[n = r][k]
    t = np.full(4 * S16_LBO, np.nan)
    for r in range(16):
        for j in range(4):
            o = kmajor_off(r, 4 * j, S16_LBO, S_SBO)
            t[o:o + 4] = X[r, 4 * j:4 * j + 4]
    return t


def tile16_cols(X):               # thread = row r of X writes column-wise: tile[n = c][k = r] = X[r][c]  (transposed tile)
    t = np.full(4 * S16_LBO, np.nan)
    for r in range(16):
        for c in range(16):
            t[kmajor_off(c, r, S16_LBO, S_SBO)] = X[r, c]
    return t


def op16(tile):
    return smem_operand(tile, 16, 16, S16_LBO, S_SBO)


def chunk_v2(Pc, V, dY, S0, U, dSe):
    """the changed part of one chunk: returns Z and the Gram-derived operand matrices, computed through tiles"""
    At, Bt, Kt, Qt = (Pc[n].numpy() for n in ("At", "Bt", "Kt", "Qt"))
    V, dY, S0, U, dSe = (x.numpy() for x in (V, dY, S0, U, dSe))
    ts, ti = np.tril(np.ones((L, L)), -1), np.tril(np.ones((L, L)))
    AQn, BKn = token_major(At, Qt), token_major(Bt, Kt)
    # G1: forward Gram blocks; rows 0-15 = A~ tokens (lanes 0-15), 16-31 = Q~ tokens (lanes 32-47)
    G1 = rows32(AQn) @ rows32(BKn).T
    N, Aak, Aqb, Aqk = G1[:16, :16] * ts, G1[:16, 16:] * ts, G1[16:, :16] * ti, G1[16:, 16:] * ti
    # Gram group: column solve of T from N^T (as forward v2), operand tiles
    Tm = np.linalg.inv(np.eye(L) - N)
    Tt = tile16_cols(Tm)                      # B operand of B2: Tt[n = t][k = s] = T[s][t]; the thread that solved column c of T
    #                                           holds T[:, c] = row c of this tile -> in the kernel these are 16-byte row stores
    AqbT, AqkT, AakT = tile16_cols(Aqb), tile16_cols(Aqk), tile16_cols(Aak)
    # B1: one N = 32 chain over dS, then the dY^T Aqb term into the R columns
    RV = dSe @ rows32(BKn).T                  # [value][0-15: . B~_t | 16-31: . K~_s]
    R_T = RV[:, :16] + dY.T @ op16(AqbT).T    # A = dYt [value][s] (smem), B = AqbT
    dV_T = RV[:, 16:].copy()                  # P3a
    # B2: Z^T = R^T T
    Z_T = R_T @ op16(Tt).T
    Z = Z_T.T
    # G2: gradient Gram blocks from the tiles that exist today (DYZn = [dY ; Z], UVn = [U ; V])
    G2 = rows32(token_major(dY, Z)) @ rows32(token_major(U, V)).T
    dAqb, dAqk, dN, dAak = G2[:16, :16] * ti, G2[:16, 16:] * ti, G2[16:, :16] * ts, G2[16:, 16:] * ts
    # P3b through the transposed 16x16 tiles
    dV_T += dY.T @ op16(AqkT).T + Z_T @ op16(AakT).T
    return dict(Z=Z, dV=dV_T.T, dN=dN, dAak=dAak, dAqb=dAqb, dAqk=dAqk, Aak=Aak, Aqk=Aqk, Aqb=Aqb, Tm=Tm)


def bwd_v2(w, q, k, v, a, b, dy, ck, sT, dsT=None):
    dt = torch.float64
    mm = P0.MM("f64")
    w, q, k, v, a, b, dy = [x.to(dt) for x in (w, q, k, v, a, b, dy)]
    T, C = w.shape
    nC = T // L
    outs = {n: torch.empty(T, C, dtype=dt) for n in "wqkvab"}
    dS_next = torch.zeros(C, C, dtype=dt) if dsT is None else dsT.to(dt).clone()
    suffix = torch.zeros(C, dtype=dt); carry_first = torch.zeros(C, dtype=dt)
    for c in range(nC - 1, -1, -1):
        Pc = P0.prep(w, q, k, v, a, b, c, dt, mm); V, dY = v[Pc["sl"]], dy[Pc["sl"]]
        win_end = (c % WIN == WIN - 1) or (c == nC - 1)
        S0 = ck[c]
        gL = torch.zeros(C, dtype=dt)
        if win_end:
            S_next = sT.to(dt) if c == nC - 1 else ck[c + 1]
            gL = (dS_next * S_next).sum(0)
            dSe = dS_next * Pc["E"][-1]
            suffix = torch.zeros(C, dtype=dt); carry_first = torch.zeros(C, dtype=dt)
        else:
            dSe = dS_next
        U = mm(Pc["W"], S0.T) + mm(Pc["M1"], V)          # the forward's `sa` (read from HBM by the kernel)
        R = {n: torch.from_numpy(x) for n, x in chunk_v2(Pc, V, dY, S0, U, dSe).items()}
        Z, dV, dN, dAak, dAqb, dAqk = (R[n] for n in ("Z", "dV", "dN", "dAak", "dAqb", "dAqk"))
        dAt = mm(Z, S0) + mm(dN, Pc["Bt"]) + mm(dAak, Pc["Kt"])                     # (P1)
        dQt = mm(dY, S0) + mm(dAqb, Pc["Bt"]) + mm(dAqk, Pc["Kt"])
        dBt = mm(U, dSe) + mm(dN.T, Pc["At"]) + mm(dAqb.T, Pc["Qt"])                # (P2)
        dKt = mm(V, dSe) + mm(dAak.T, Pc["At"]) + mm(dAqk.T, Pc["Qt"])
        dS_next = dSe + mm(dY.T, Pc["Qt"]) + mm(Z.T, Pc["At"])                     # (R2)
        sl = Pc["sl"]
        outs["a"][sl] = dAt * Pc["Ep"]; outs["b"][sl] = dBt / Pc["E"]; outs["k"][sl] = dKt / Pc["E"]
        outs["q"][sl] = dQt * Pc["E"]; outs["v"][sl] = dV
        g = dQt * Pc["Qt"] - dKt * Pc["Kt"] - dBt * Pc["Bt"]
        aa = dAt * Pc["At"]
        g[:-1] += aa[1:]
        g[-1] += carry_first + gL
        carry_first = aa[0]
        suf = torch.flip(torch.cumsum(torch.flip(g, [0]), 0), [0]) + suffix
        suffix = suf[0]
        outs["w"][sl] = suf * Pc["lw"]
    return outs, dS_next


if __name__ == "__main__":
    B, T, H = 1, 208, 2
    x = O.make_inputs(B, T, H, seed=5)
    names = "wqkvab"
    s0 = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    dsT = torch.randn(B, H, 64, 64, dtype=torch.float64) * 0.1
    g64 = O.wkv7_backward(*[x[n] for n in names], x["dy"], s0=s0, dsT=dsT)
    worst = {}
    for h in range(H):
        xs = [x[n][0, :, h] for n in names]
        y, sT, ck = P0.fwd_ckpt(*xs, mode="f64", s0=s0[0, h])
        outs, dS0 = bwd_v2(*xs, x["dy"][0, :, h], ck, sT, dsT=dsT[0, h])
        for i, n in enumerate(names):
            worst[n] = max(worst.get(n, 0), O.rel_l2(outs[n], g64[i][0, :, h]))
        worst["s0"] = max(worst.get("s0", 0), O.rel_l2(dS0, g64[6][0, h]))
    print("backward v2 blueprint vs oracle (f64): " + "  ".join(f"d{n} {e:.1e}" for n, e in worst.items()))
