"""Index-level emulation of the forward-v2 sketch (proto/wkv7_tc_fwd_v2.cu) on the CPU: every shared-memory tile is a
flat float array written and read with the SAME offset expressions as the .cu file, tensor memory is a [64 rows][256
columns] array, and each tcgen05 instruction chain is modelled as D[m][n] (+)= sum_k A[m][k] * B[n][k] with K-major
canonical operands (tc05.cuh).  Products are exact fp32 -> f64 here (no tf32 rounding), so any disagreement with the f64
oracle beyond ~1e-6 is an indexing / orientation / masking / window-frame mistake in the sketch, not numerics.
What it cannot check: barriers, proxy fences, tensor-memory lane quadrants, instruction descriptors."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wkv7_oracle as O  # noqa: E402

L, WIN, kC = 16, 4, 64
WQ_LBO, WQ_SBO = 132, 32
T_SBO, T_LBO = 36, 288
MA_LBO, MA_SBO = 132, 32
QB_LBO, QB_SBO = 68, 32
C_ZY, C_ZY_STRIDE, C_G = 64, 48, 160
kMinLogDecay = -1.35


def kmajor_off(r, k, lbo, sbo):
    return (r >> 3) * sbo + (k >> 2) * lbo + (r & 7) * 4 + (k & 3)


class Slot:
    def __init__(self):
        self.WQ = np.full(16 * WQ_LBO, np.nan); self.BK = np.full(16 * WQ_LBO, np.nan)
        self.Bt = np.full(4 * T_LBO, np.nan); self.Kt = np.full(4 * T_LBO, np.nan); self.Vt = np.full(4 * T_LBO, np.nan)
        self.MA = np.full(4 * MA_LBO, np.nan); self.Aqb = np.full(4 * QB_LBO, np.nan); self.Tt = np.full(4 * QB_LBO, np.nan)


def smem_operand(tile, rows, K, lbo, sbo):
    """what the tensor core reads: X[r][k] for r < rows, k < K"""
    return np.array([[tile[kmajor_off(r, k, lbo, sbo)] for k in range(K)] for r in range(rows)])


WAVES = {"st4": [0, 0], "scalar": [0, 0], "gram group st4": [0, 0], "gram group NT stores": [0, 0], "T tile stores": [0, 0]}
# [warp instructions, wavefronts] per kind of shared-memory store


def _count(kind, word_addrs, width):
    """shared-memory wavefronts of one warp store instruction: `width` consecutive 32-bit words per lane; 32 banks;
    128-bit accesses are served a quarter-warp at a time, 64-bit a half-warp, 32-bit the whole warp"""
    per = {1: 32, 2: 16, 4: 8}[width]
    waves = 0
    for g0 in range(0, len(word_addrs), per):
        banks = {}
        for a0 in word_addrs[g0:g0 + per]:
            for wd in range(width):
                banks.setdefault((a0 + wd) % 32, set()).add(a0 + wd)
        waves += max(len(v) for v in banks.values())
    WAVES[kind][0] += 1
    WAVES[kind][1] += waves


def stage_a(S, w, q, k, v, a, b, gpre):
    """one chunk, thread by thread as in the .cu file: 256 threads, lane = (t & 3) * 8 + (k4 & 7), warp = (t >> 2) * 2 +
    (k4 >> 3); the decay scan with its two warp shuffles and the two-stage cross-warp prefix; every tile store with the
    file's offset expression (and its bank conflicts counted).  Inputs [16][64] float64; gpre [64] = log decay
    accumulated since the window start.  Returns the chunk total and e^{G} of the last token (DLw)."""
    T_ = np.zeros(256, dtype=int); K4 = np.zeros(256, dtype=int)
    gg = np.zeros((256, 4)); lw = np.zeros((256, 4))
    for tp in range(256):
        wp, lane = tp >> 5, tp & 31
        tt, tg = lane >> 3, wp >> 1
        T_[tp], K4[tp] = 4 * tg + tt, 8 * (wp & 1) + (lane & 7)
        lw[tp] = np.maximum(-np.exp(w[T_[tp], 4 * K4[tp]:4 * K4[tp] + 4]), kMinLogDecay)
        gg[tp] = lw[tp]

    def shfl_up(x, delta):
        y = x.copy()
        for tp in range(256):
            if (tp & 31) >= delta:
                y[tp] = x[tp - delta]
        return y
    for delta, need in ((8, 1), (16, 2)):
        x = shfl_up(gg, delta)
        for tp in range(256):
            if ((tp & 31) >> 3) >= need:
                gg[tp] += x[tp]
    wt = np.full((9, kC), np.nan)
    for tp in range(256):
        if ((tp & 31) >> 3) == 3:
            wt[(tp >> 5) >> 1, 4 * K4[tp]:4 * K4[tp] + 4] = gg[tp]
    for ch in range(kC):                                   # threads tp < 64
        run = 0.0
        for ww in range(4):
            x = wt[ww, ch]; wt[ww, ch] = run; run += x
        wt[4, ch] = run
    tot = np.zeros(kC)
    for tp in range(256):
        tg, ch = (tp >> 5) >> 1, slice(4 * K4[tp], 4 * K4[tp] + 4)
        gg[tp] += gpre[ch] + wt[tg, ch]
        tot[ch] = gpre[ch] + wt[4, ch]
    D, Dp, iD = np.exp(gg), np.exp(gg - lw), np.exp(-gg)
    dl = np.zeros(kC)
    for wp in range(8):                                    # one warp instruction at a time, for the conflict count
        tps = range(32 * wp, 32 * wp + 32)
        oa = [(T_[tp] >> 3) * WQ_SBO + K4[tp] * WQ_LBO + (T_[tp] & 7) * 4 for tp in tps]
        oq = [o + 2 * WQ_SBO for o in oa]
        ot = [(K4[tp] >> 1) * T_SBO + (T_[tp] >> 2) * T_LBO + (K4[tp] & 1) * 16 + (T_[tp] & 3) for tp in tps]
        for _ in range(2):
            _count("st4", oq, 4); _count("st4", oa, 4)   # WQ and BK, rows t and 16 + t
        for j in range(4):
            for _ in range(3):
                _count("scalar", [o + 4 * j for o in ot], 1)   # Kt, Vt, Bt
        for i, tp in enumerate(tps):
            t, k4 = T_[tp], K4[tp]
            ch = slice(4 * k4, 4 * k4 + 4)
            S.WQ[oq[i]:oq[i] + 4] = q[t, ch] * D[tp]
            o = k[t, ch] * iD[tp]
            S.BK[oq[i]:oq[i] + 4] = o
            for j in range(4):
                S.Kt[ot[i] + 4 * j] = o[j]
                S.Vt[ot[i] + 4 * j] = v[t, 4 * k4 + j]
            S.WQ[oa[i]:oa[i] + 4] = a[t, ch] * Dp[tp]
            o = b[t, ch] * iD[tp]
            S.BK[oa[i]:oa[i] + 4] = o
            for j in range(4):
                S.Bt[ot[i] + 4 * j] = o[j]
            if t == L - 1:
                dl[ch] = D[tp]
    return tot, dl


def stage_g(S, tm, u):
    NT = np.zeros(16 * 20)
    # bank conflicts of the group's stores: lanes 0-15 active (rows), one warp instruction each
    for j in range(4):
        _count("gram group st4", [kmajor_off(r, 4 * j, MA_LBO, MA_SBO) for r in range(16)], 4)          # Aak
        _count("gram group st4", [kmajor_off(r, 4 * j, QB_LBO, QB_SBO) for r in range(16)], 4)          # Aqb
        _count("gram group st4", [kmajor_off(16 + r, 4 * j, MA_LBO, MA_SBO) for r in range(16)], 4)     # Aqk
    for s_ in range(16):
        _count("gram group NT stores", [s_ * 20 + r for r in range(16)], 1)
    for tt in range(L):
        _count("T tile stores", [kmajor_off(tt, col, QB_LBO, QB_SBO) for col in range(16)], 1)
    for wq in range(2):
        for row in range(16):
            g = tm[16 * wq + row, C_G + 32 * u:C_G + 32 * u + 32]      # lanes 0-15 (wq = 0) / 32-47 (wq = 1) = rows 0-15 / 16-31
            gb, gk = g[:16].copy(), g[16:].copy()
            if wq == 0:
                for s_ in range(16):
                    NT[s_ * 20 + row] = gb[s_] if s_ < row else 0.0
                    gk[s_] = gk[s_] if s_ < row else 0.0
                for j in range(4):
                    o = kmajor_off(row, 4 * j, MA_LBO, MA_SBO)
                    S.MA[o:o + 4] = gk[4 * j:4 * j + 4]
            else:
                for s_ in range(16):
                    gb[s_] = gb[s_] if s_ <= row else 0.0
                    gk[s_] = gk[s_] if s_ <= row else 0.0
                for j in range(4):
                    o = kmajor_off(row, 4 * j, QB_LBO, QB_SBO)
                    S.Aqb[o:o + 4] = gb[4 * j:4 * j + 4]
                    o = kmajor_off(16 + row, 4 * j, MA_LBO, MA_SBO)
                    S.MA[o:o + 4] = gk[4 * j:4 * j + 4]
    for col in range(16):
        acc = [1.0 if tt == col else 0.0 for tt in range(L)]
        for s_ in range(L - 1):
            x = acc[s_]
            for q4 in range((s_ + 1) // 4, 4):
                nn = NT[s_ * 20 + 4 * q4:s_ * 20 + 4 * q4 + 4]
                for e in range(4):
                    if 4 * q4 + e > s_:
                        acc[4 * q4 + e] += nn[e] * x
        for tt in range(L):
            S.Tt[kmajor_off(tt, col, QB_LBO, QB_SBO)] = acc[tt]


def forward(w, q, k, v, a, b, s0=None):
    """one (batch, head): inputs [T][64]; returns y [T][64], S_T [64][64] value-major"""
    T = w.shape[0]
    nC = T // L
    tm = np.zeros((64, 256))
    if s0 is not None:
        tm[:, 0:64] = s0
    slots = [Slot() for _ in range(nC)]          # the ring is not modelled (barriers are out of scope)
    gpre, DL = np.zeros(kC), [None] * nC
    for c in range(nC):
        sl = slice(c * L, (c + 1) * L)
        tot, dl = stage_a(slots[c], w[sl], q[sl], k[sl], v[sl], a[sl], b[sl], gpre)
        win_end = (c % WIN == WIN - 1) or (c == nC - 1)
        gpre = np.zeros(kC) if win_end else tot
        DL[c] = dl
    y = np.zeros((T, kC))
    for c in range(nC):
        S, u = slots[c], c & 1
        uz = C_ZY + C_ZY_STRIDE * u
        uy, uu = uz + 16, uz + 32
        # Gram instruction: A = WQ (M = 64, rows 0-31 live), B = BK (N = 32), K = 64
        A, B = smem_operand(S.WQ, 32, 64, WQ_LBO, WQ_SBO), smem_operand(S.BK, 32, 64, WQ_LBO, WQ_SBO)
        tm[0:32, C_G + 32 * u:C_G + 32 * u + 32] = A @ B.T
        stage_g(S, tm, u)
        # phase 1: [Z^T | Y^T] = S^ [A~ | Q~]^T + V^T [Aak | Aqk]^T
        Sst = tm[:, 0:64].copy()
        tm[:, uz:uz + 32] = Sst @ smem_operand(S.WQ, 32, 64, WQ_LBO, WQ_SBO).T \
            + smem_operand(S.Vt, 64, 16, T_LBO, T_SBO) @ smem_operand(S.MA, 32, 16, MA_LBO, MA_SBO).T
        # phase 1b: U^T = Z^T T^T   (A = tensor-memory columns uz .. uz+15, B = Tt [n = t][k = s])
        tm[:, uu:uu + 16] = tm[:, uz:uz + 16] @ smem_operand(S.Tt, 16, 16, QB_LBO, QB_SBO).T
        # phase 2: S^ += U^T B~ + V^T K~ ;  Y^T += U^T Aqb^T
        Ut = tm[:, uu:uu + 16].copy()
        tm[:, 0:64] += Ut @ smem_operand(S.Bt, 64, 16, T_LBO, T_SBO).T \
            + smem_operand(S.Vt, 64, 16, T_LBO, T_SBO) @ smem_operand(S.Kt, 64, 16, T_LBO, T_SBO).T
        tm[:, uy:uy + 16] += Ut @ smem_operand(S.Aqb, 16, 16, QB_LBO, QB_SBO).T
        # epilogue
        y[c * L:(c + 1) * L] = tm[:, uy:uy + 16].T
        if (c % WIN == WIN - 1) or (c == nC - 1):
            tm[:, 0:64] *= DL[c][None, :]
    return y, tm[:, 0:64].copy()


def check_sa_quad_transpose():
    """training variant: the epilogue's 4x4 quad transposes (two __shfl_xor rounds per block of 4 tokens) followed by
    16-byte stores must produce the backward's `sa` operand tile, element (token, value) at (value/4)*64 + (token/8)*32 +
    (token%8)*4 + value%4 (wkv7_common.cuh), for the 64 value rows of the four epilogue warps"""
    rng = np.random.default_rng(0)
    U = rng.standard_normal((16, 64))                      # [token][value]
    blob = np.full(1024, np.nan)
    for q in range(4):                                     # warp q: lanes 0-15 hold value rows 16q .. 16q+15
        uv = np.array([[U[tok, 16 * q + lane] for tok in range(16)] for lane in range(16)])      # uv[lane][token]
        for blk in range(4):
            for delta, pairs in ((1, ((0, 1), (2, 3))), (2, ((0, 2), (1, 3)))):
                for j0, j1 in pairs:
                    snd = np.array([uv[l, 4 * blk + (j0 if (l & 3) & delta else j1)] for l in range(16)])
                    rcv = np.array([snd[l ^ delta] for l in range(16)])
                    for l in range(16):
                        uv[l, 4 * blk + (j0 if (l & 3) & delta else j1)] = rcv[l]
        for lane in range(16):
            row, vi = 16 * q + lane, lane & 3
            for blk in range(4):
                tok = 4 * blk + vi
                o = (row >> 2) * 64 + (tok >> 3) * 32 + (tok & 7) * 4
                blob[o:o + 4] = uv[lane, 4 * blk:4 * blk + 4]
    want = np.array([U[(o % 64) // 32 * 8 + (o % 32) // 4, (o // 64) * 4 + o % 4] for o in range(1024)])
    assert np.array_equal(blob, want), "sa blob layout"
    return True


if __name__ == "__main__":
    print("sa quad transpose (training epilogue):", "ok" if check_sa_quad_transpose() else "MISMATCH")
    x = O.make_inputs(1, 208, 2, seed=5)          # 13 chunks: full windows and a ragged last one
    names = "wqkvab"
    s0 = torch.randn(1, 2, 64, 64, dtype=torch.float64) * 0.1
    y64, sT64 = O.wkv7_forward(*[x[n] for n in names], s0=s0)
    for h in range(2):
        y, sT = forward(*[x[n][0, :, h].double().numpy() for n in names], s0=s0[0, h].numpy())
        ey = float(np.linalg.norm(y - y64[0, :, h].numpy()) / np.linalg.norm(y64[0, :, h].numpy()))
        es = float(np.linalg.norm(sT - sT64[0, h].numpy()) / np.linalg.norm(sT64[0, h].numpy()))
        print(f"head {h}: y rel-l2 {ey:.2e}   S_T rel-l2 {es:.2e}   nan in y: {bool(np.isnan(y).any())}")
    for kind, (n, wv) in WAVES.items():
        print(f"{kind:24s}: {wv / n:.2f} wavefronts per warp instruction (conflict-free: 1.00 for 4-byte stores, 4.00 / 2.00 for "
              f"16-byte stores of 32 / 16 lanes)")
