// SKETCH, NOT BUILT INTO THE LIBRARY, NEVER RUN ON A GPU (written after the round's GPU budget was spent; it only
// has to compile: nvcc -c -I rwkvtts_b200/csrc proto/wkv7_tc_fwd_v2.cu).  Starting point for round 2.
//
// "Forward v2" of DESIGN.md section 7 (both variants): the shipped forward (csrc/wkv7_tc_fwd.cu) with
//   * the U-form: phase 1 gives Z^T = S^ A~^T + V^T Aak^T (and Y^T), a new phase 1b applies the triangular factor on the
//     tensor core, U^T = Z^T T^T, so W~ = T A~ and M1 = T Aak are never formed on the CUDA cores;
//   * the four Gram blocks as ONE tcgen05 instruction chain per chunk, G[64 x 32] = [A~ ; Q~ ; - ; -] [B~ ; K~]^T
//     (M = 64 with 32 live rows, rows 0-15 -> tensor-memory lanes 0-15, rows 16-31 -> lanes 32-47), issued two chunks
//     ahead of the state chain; its A operand is the slot tile WQ = [A~ ; Q~] that phase 1 uses as B operand, its B operand
//     a new token-major tile BK = [B~ ; K~].  The natural [token][channel] tiles and the mma.sync stage are gone;
//   * a Gram group of two warps per parity (warp % 4 == 0 reads lanes 0-15 = N | Aak, warp % 4 == 1 reads lanes 32-47 =
//     Aqb | Aqk): mask, round, 16-byte operand-tile stores, and the column solve of T only (16 identity columns).
// Expected shared-memory wavefronts per chunk ~650 against 1455 (profiles/r01_tc_pair_smem_lines_v5.txt); numerics of the
// form: proto/uform_numerics_proto.py; tile offsets, operand orientations, masks, tensor-memory columns and window frames
// of THIS file replayed on the CPU against the f64 oracle: proto/fwd_v2_index_emulator.py (1e-15); mbarrier protocol
// (counts, parities, issue order, both variants) model-checked in proto/fwd_v2_sync_model.py.  Not checked by
// anything: proxy / tcgen05 fences, lane quadrants, descriptors, the different-accumulator ordering rule.  Everything not mentioned is the shipped kernel's code.
#include "mma_tf32.cuh"
#include "tc05.cuh"
#include "wkv7_common.cuh"

namespace rwkvtts {
namespace tcfwd2 {
using namespace tc05;

constexpr int L = 16;        // chunk length
constexpr int WIN = 4;       // chunks per window
constexpr int NSLOT = 5;     // operand slots in flight
constexpr float kMinLogDecay = -1.35f;
// tensor-memory columns: S^ 0-63 | per chunk parity u: Z^T 64+48u, Y^T +16, U^T +32 | Gram blocks 160+32u (0-15: x B~, 16-31: x K~)
constexpr uint32_t C_ZY = 64, C_ZY_STRIDE = 48, C_G = 160, kTmemCols = 256;
constexpr uint32_t C_ST = 256, kTmemColsTrain = 512;   // training variant: transposed state S^T in columns 256-319
constexpr float kLog2e = 1.4426950408889634f;   // decays are accumulated as log2 (ex2.approx needs no pre-scale)

// canonical K-major tiles, strides in floats (see tc05.cuh: off = (r/8)*SBO + (k/4)*LBO + (r%8)*4 + k%4)
constexpr int WQ_LBO = 132, WQ_SBO = 32;     // [32 rows: 0-15 A~ tokens, 16-31 Q~ tokens][64 channels]; same strides for BK
constexpr int T_SBO = 36, T_LBO = 288;       // [64 rows: channel / value][16 tokens]   (transposed tiles)
constexpr int MA_LBO = 132, MA_SBO = 32;     // [32 rows: 0-15 Aak, 16-31 Aqk][16]
constexpr int QB_LBO = 68, QB_SBO = 32;      // [16][16]: Aqb and T; 68 = 4 mod 32: the column-per-lane stores of T are conflict-free (64: 4-way)

struct Slot {
    float WQ[16 * WQ_LBO];   // as the M = 64 A operand of the Gram instruction it is read 32 rows past its end (into BK:
    float BK[16 * WQ_LBO];   // those rows land in tensor-memory lanes 64-111 of the Gram columns, which nobody reads)
    float Bt[4 * T_LBO], Kt[4 * T_LBO], Vt[4 * T_LBO];
    float MA[4 * MA_LBO];
    float Aqb[4 * QB_LBO], Tt[4 * QB_LBO];
};
struct Smem {
    Slot slot[NSLOT];
    float NT[2][L * 20];                   // per Gram group: N^T, fp32, for the column solve
    float wtot[9][kC];                     // stage A scan: per-warp totals -> exclusive prefixes, chunk total
    __align__(16) bf16 ybuf[2][L][72];     // epilogue: Y tile [token][value], double buffered
    __align__(16) float Ut[2][4 * T_LBO];  // training: U^T [value][token] operand tile of the transposed-state update
    float DLw[4][kC];                      // e^{G} at the end of a window (ring of 4 windows)
    uint64_t empty[NSLOT], full[NSLOT], a_done[NSLOT], g_ready[2];
    uint64_t p_done, y_ready[2], y_free[2], win_scaled, ut_ready, st_ready, st_free;
    uint32_t tmem_base;
};

struct Params {
    int T, H;
    const bf16 *w, *q, *k, *v, *a, *b;
    bf16 *y;
    float *ckT;          // training: TRANSPOSED state at the start of every chunk (window frame), operand tiles of 4096
                         // floats per chunk (wkv7_common.cuh); null for the snapshot-free forward
    float *sa;           // training: U_t = S_{t-1} a_t (tf32), operand tiles of 1024 floats per chunk
    const float *s0;     // may be null
    float *sT;           // may be null
    long long *dbg;      // phase-cycle counters (profiling builds only), may be null
};

#ifdef RWKVTTS_PROFILE
#define TICK(var) long long var = clock64()
#define ACC(slot, t0, t1) do { if (P_dbg && blockIdx.x == 0) P_dbg[slot] += (t1) - (t0); } while (0)
#else
#define TICK(var)
#define ACC(slot, t0, t1)
#endif

__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

// ---------------------------------------------------------------------------------------------
// stage A: tp in [0,256).  A warp holds 4 tokens x 8 channel groups (the shipped kernel: 2 tokens x 16 groups), lane =
// (t & 3) * 8 + (k4 & 7), warp = (t >> 2) * 2 + (k4 >> 3): a scalar store into a transposed [channel][token] tile then covers
// all four token columns of 8 core-matrix rows = 32 different banks (2 x 16 covers 16 banks twice: 192 instead of 96
// wavefronts per chunk, profiles/r01_tc_pair_smem_lines_v5.txt), and the 16-byte row stores stay one quarter-warp per
// 128-byte row piece.  Thread = token t, channels 4*k4 .. 4*k4+3.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 ldg_nc_v2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void unpack4(const uint2 &u, float *f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
}
__device__ __forceinline__ void load_raw(const Params &P, size_t base, size_t tok_stride, int c, int t, int k4,
                                         uint2 (&raw)[6]) {
    const size_t off = base + (size_t)(c * L + t) * tok_stride + k4 * 4;
    raw[0] = ldg_nc_v2(P.w + off);
    raw[1] = ldg_nc_v2(P.q + off);
    raw[2] = ldg_nc_v2(P.k + off);
    raw[3] = ldg_nc_v2(P.v + off);
    raw[4] = ldg_nc_v2(P.a + off);
    raw[5] = ldg_nc_v2(P.b + off);
}

__device__ void stage_a(const Params &P, Smem &sm, size_t base, size_t tok_stride, int nC, int tp) {
    long long *P_dbg = tp == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int wp = tp >> 5, lane = tp & 31, tt = lane >> 3, tg = wp >> 1;
    const int t = 4 * tg + tt, k4 = 8 * (wp & 1) + (lane & 7);
    uint2 raw[6], nxt[6], nx2[6];
    float gpre[4];
#pragma unroll
    for (int j = 0; j < 4; j++) gpre[j] = 0.f;
    load_raw(P, base, tok_stride, 0, t, k4, raw);
    if (nC > 1) load_raw(P, base, tok_stride, 1, t, k4, nxt);
    for (int c = 0; c < nC; c++) {
        const int si = c % NSLOT;
        Slot &S = sm.slot[si];
        if (c + 2 < nC) load_raw(P, base, tok_stride, c + 2, t, k4, nx2);   // two chunks ahead
        float lw[4], gg[4];
        TICK(ta0);
        {
            float f[4];
            unpack4(raw[0], f);
#pragma unroll
            for (int j = 0; j < 4; j++) {   // log2 of the decay: -e^w log2(e), clamped
                lw[j] = fmaxf(-kLog2e * ex2f(f[j] * kLog2e), kMinLogDecay * kLog2e);
                gg[j] = lw[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {   // inclusive scan over the 4 tokens of this warp (token = lane >> 3)
            float x = __shfl_up_sync(0xffffffffu, gg[j], 8);
            if (tt >= 1) gg[j] += x;
            x = __shfl_up_sync(0xffffffffu, gg[j], 16);
            if (tt >= 2) gg[j] += x;
        }
        // cross-warp prefix in two stages: 4 token-group totals per channel (two warps each cover half the channels) ->
        // 64 threads turn them into exclusive prefixes (+ the chunk total in row 4) -> every thread reads two rows
        float(&wt)[9][kC] = sm.wtot;
        if (tt == 3) st4(&wt[tg][k4 * 4], gg[0], gg[1], gg[2], gg[3]);
        bar_sync(1, 256);
        if (tp < kC) {
            float run = 0.f;
#pragma unroll
            for (int ww = 0; ww < 4; ww++) {
                const float x = wt[ww][tp];
                wt[ww][tp] = run;
                run += x;
            }
            wt[4][tp] = run;
        }
        bar_sync(1, 256);
        float tot[4];
        {
            const float4 pre = *reinterpret_cast<const float4 *>(&wt[tg][k4 * 4]);
            const float4 all = *reinterpret_cast<const float4 *>(&wt[4][k4 * 4]);
            gg[0] += gpre[0] + pre.x; gg[1] += gpre[1] + pre.y; gg[2] += gpre[2] + pre.z; gg[3] += gpre[3] + pre.w;
            tot[0] = gpre[0] + all.x; tot[1] = gpre[1] + all.y; tot[2] = gpre[2] + all.z; tot[3] = gpre[3] + all.w;
        }
        bar_sync(1, 256);       // rows are rewritten by the next chunk
        const bool win_end = (c % WIN == WIN - 1) || (c == nC - 1);
#pragma unroll
        for (int j = 0; j < 4; j++) gpre[j] = win_end ? 0.f : tot[j];

        // wait until the slot of this chunk has been released (phase 2 of chunk c - NSLOT has completed)
        TICK(ta1);
        if (c >= NSLOT) mbar_wait(&sm.empty[si], ((c / NSLOT) - 1) & 1);
        TICK(ta2);
        TICK(ta3);
        {
            float D[4], Dp[4], iD[4], f[4], o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                D[j] = ex2f(gg[j]);
                Dp[j] = ex2f(gg[j] - lw[j]);
                iD[j] = ex2f(-gg[j]);
            }
            // token-major canonical rows: row t (A~, B~) and row 16 + t (Q~, K~) of WQ / BK, channels 4*k4 .. +3
            const int oa = (t >> 3) * WQ_SBO + k4 * WQ_LBO + (t & 7) * 4;
            const int oq = oa + 2 * WQ_SBO;
            // Q~
            unpack4(raw[1], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * D[j]);
            st4(&S.WQ[oq], o[0], o[1], o[2], o[3]);
            // transposed tiles: row = channel 4*k4+j, column = token t
            const int ot = (k4 >> 1) * T_SBO + (t >> 2) * T_LBO + (k4 & 1) * 16 + (t & 3);   // + 4*j
            // K~
            unpack4(raw[2], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * iD[j]);
            st4(&S.BK[oq], o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) S.Kt[ot + 4 * j] = o[j];
            // V (bf16 values are exact in tf32)
            unpack4(raw[3], f);
#pragma unroll
            for (int j = 0; j < 4; j++) S.Vt[ot + 4 * j] = f[j];
            // A~
            unpack4(raw[4], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * Dp[j]);
            st4(&S.WQ[oa], o[0], o[1], o[2], o[3]);
            // B~
            unpack4(raw[5], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * iD[j]);
            st4(&S.BK[oa], o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) S.Bt[ot + 4 * j] = o[j];
            if (win_end && t == L - 1) st4(&sm.DLw[(c / WIN) & 3][k4 * 4], D[0], D[1], D[2], D[3]);
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.a_done[si]);      // consumer: the MMA warp (Gram instruction of this chunk)
#pragma unroll
        for (int i = 0; i < 6; i++) { raw[i] = nxt[i]; nxt[i] = nx2[i]; }
        TICK(ta4); ACC(0, ta0, ta1); ACC(1, ta1, ta2); ACC(2, ta2, ta3); ACC(3, ta3, ta4);
    }
}

// ---------------------------------------------------------------------------------------------
// Gram group: two warps per chunk parity.  wq = 0 (warp % 4 == 0, lanes 0-15 = A~ rows): N | Aak, then the column solve
// of T; wq = 1 (warp % 4 == 1, lanes 32-47 = Q~ rows): Aqb | Aqk.  tp in [0,64).
// ---------------------------------------------------------------------------------------------
__device__ void stage_g(const Params &P, Smem &sm, int nC, int tp, int grp) {
    long long *P_dbg = (grp == 0 && tp == 0) ? P.dbg : nullptr; (void)P_dbg;
    const int wq = tp >> 5, lane = tp & 31, row = lane & 15;
    const bool act = lane < 16;
    float *NT = sm.NT[grp];
    const uint32_t tg = sm.tmem_base + ((uint32_t)(32 * wq) << 16) + C_G + 32 * grp;
    for (int c = grp; c < nC; c += 2) {
        const int si = c % NSLOT;
        Slot &S = sm.slot[si];
        TICK(tb0);
        mbar_wait(&sm.g_ready[grp], (c >> 1) & 1);
        TICK(tb1);
        fence_after_sync();
        float gb[16], gk[16];
        tmem_ld16(tg, gb);          // row . B~_s, s = 0..15
        tmem_ld16(tg + 16, gk);     // row . K~_s
        tmem_wait_ld();
        fence_before_sync();        // the Gram columns may be overwritten once this group has arrived on full[si]
        if (act) {
            if (wq == 0) {          // A~ rows: strictly lower triangle
#pragma unroll
                for (int s_ = 0; s_ < 16; s_++) {
                    NT[s_ * 20 + row] = (s_ < row) ? gb[s_] : 0.f;          // fp32 for the solve, transposed
                    gk[s_] = (s_ < row) ? tf32r(gk[s_]) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                    st4(&S.MA[kmajor_off(row, 4 * j, MA_LBO, MA_SBO)], gk[4 * j], gk[4 * j + 1], gk[4 * j + 2], gk[4 * j + 3]);
            } else {                // Q~ rows: lower triangle with the diagonal
#pragma unroll
                for (int s_ = 0; s_ < 16; s_++) {
                    gb[s_] = (s_ <= row) ? tf32r(gb[s_]) : 0.f;
                    gk[s_] = (s_ <= row) ? tf32r(gk[s_]) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    st4(&S.Aqb[kmajor_off(row, 4 * j, QB_LBO, QB_SBO)], gb[4 * j], gb[4 * j + 1], gb[4 * j + 2], gb[4 * j + 3]);
                    st4(&S.MA[kmajor_off(16 + row, 4 * j, MA_LBO, MA_SBO)], gk[4 * j], gk[4 * j + 1], gk[4 * j + 2], gk[4 * j + 3]);
                }
            }
        }
        bar_sync(2 + grp, 64);
        TICK(tb2);
        // ---- T = (I - N)^-1, one identity column per thread, column-oriented (the shipped kernel solves 80 columns) ---
        if (tp < 16) {
            const int col = tp;
            float acc[L];
#pragma unroll
            for (int tt = 0; tt < L; tt++) acc[tt] = (tt == col) ? 1.f : 0.f;
#pragma unroll
            for (int s_ = 0; s_ < L - 1; s_++) {
                const float x = acc[s_];
#pragma unroll
                for (int q4 = (s_ + 1) / 4; q4 < 4; q4++) {
                    const float4 n4 = *reinterpret_cast<const float4 *>(&NT[s_ * 20 + 4 * q4]);
                    const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (4 * q4 + e > s_) acc[4 * q4 + e] = fmaf(nn[e], x, acc[4 * q4 + e]);
                }
            }
            // B operand of phase 1b: Tt[n = t][k = s] = T[t][s]; this thread holds column s = col
#pragma unroll
            for (int tt = 0; tt < L; tt++) S.Tt[kmajor_off(tt, col, QB_LBO, QB_SBO)] = tf32r(acc[tt]);
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.full[si]);
        TICK(tb3);
        bar_sync(2 + grp, 64);      // NT is reused by the next chunk of this group
        TICK(tb4); ACC(4, tb0, tb1); ACC(5, tb1, tb2); ACC(6, tb2, tb3); ACC(7, tb3, tb4);
    }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer (one warp)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void issue_gram(Smem &sm, uint32_t tb, int c) {   // elected lane only
    constexpr uint32_t I32 = idesc_tf32(64, 32, false, false);
    const Slot &S = sm.slot[c % NSLOT];
    const uint64_t dWQ = smem_desc(smem_u32(S.WQ), WQ_LBO * 4, WQ_SBO * 4);
    const uint64_t dBK = smem_desc(smem_u32(S.BK), WQ_LBO * 4, WQ_SBO * 4);
#pragma unroll
    for (int kk = 0; kk < 8; kk++)
        mma_tf32_ss(tb + C_G + 32 * (c & 1), dWQ + (uint64_t)((kk * 2 * WQ_LBO * 4) >> 4),
                    dBK + (uint64_t)((kk * 2 * WQ_LBO * 4) >> 4), I32, kk > 0);
    mma_commit(&sm.g_ready[c & 1]);
}

template <bool kTrain>
__device__ void mma_warp(const Params &P, Smem &sm, int nC) {
    long long *P_dbg = (threadIdx.x & 31) == 0 ? P.dbg : nullptr; (void)P_dbg;
    const uint32_t tb = sm.tmem_base;
    constexpr uint32_t I16 = idesc_tf32(64, 16, false, false);
    constexpr uint32_t I32 = idesc_tf32(64, 32, false, false);
    constexpr uint32_t I64 = idesc_tf32(64, 64, false, false);
    uint32_t ph = 0;
    // Gram blocks of chunks 0 and 1 (afterwards chunk c + 2 is issued inside iteration c)
    for (int c = 0; c < 2 && c < nC; c++) {
        mbar_wait(&sm.a_done[c % NSLOT], 0);
        fence_after_sync();
        if (elect_one()) issue_gram(sm, tb, c);
        __syncwarp();
    }
    for (int c = 0; c < nC; c++) {
        const int si = c % NSLOT, u = c & 1;
        const Slot &S = sm.slot[si];
        const uint32_t uz = tb + C_ZY + C_ZY_STRIDE * u, uy = uz + 16, uu = uz + 32;
        const uint64_t dWQ = smem_desc(smem_u32(S.WQ), WQ_LBO * 4, WQ_SBO * 4);
        const uint64_t dVt = smem_desc(smem_u32(S.Vt), T_LBO * 4, T_SBO * 4);
        const uint64_t dBt = smem_desc(smem_u32(S.Bt), T_LBO * 4, T_SBO * 4);
        const uint64_t dKt = smem_desc(smem_u32(S.Kt), T_LBO * 4, T_SBO * 4);
        const uint64_t dMA = smem_desc(smem_u32(S.MA), MA_LBO * 4, MA_SBO * 4);
        const uint64_t dQB = smem_desc(smem_u32(S.Aqb), QB_LBO * 4, QB_SBO * 4);
        const uint64_t dTt = smem_desc(smem_u32(S.Tt), QB_LBO * 4, QB_SBO * 4);
        TICK(tm0);
        mbar_wait(&sm.full[si], (c / NSLOT) & 1);     // Aak / Aqk / Aqb / T of chunk c are in the slot; G[u] has been read
        TICK(tm1);
        if (c % WIN == 0) mbar_wait(&sm.win_scaled, (c / WIN) & 1);
        TICK(tm2);
        if (c >= 2) mbar_wait(&sm.y_free[u], ((c >> 1) - 1) & 1);
        if (c + 2 < nC) mbar_wait(&sm.a_done[(c + 2) % NSLOT], ((c + 2) / NSLOT) & 1);   // stage A runs slots ahead
        TICK(tm3);
        fence_after_sync();
        if (elect_one()) {
            // phase 1: [Z^T | Y^T] = S^ [A~ | Q~]^T + V^T [Aak | Aqk]^T
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ts(uz, tb + 8 * kk, dWQ + (uint64_t)((kk * 2 * WQ_LBO * 4) >> 4), I32, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(uz, dVt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), dMA + (uint64_t)((kk * 2 * MA_LBO * 4) >> 4),
                            I32, true);
            mma_commit(&sm.p_done);
            // off the chain, behind phase 1 in the pipe: the Gram blocks of chunk c + 2 into the columns chunk c used
            if (c + 2 < nC) issue_gram(sm, tb, c + 2);
        }
        __syncwarp();
        mbar_wait(&sm.p_done, ph); ph ^= 1;
        fence_after_sync();
        if (elect_one()) {
            // phase 1b: U^T = Z^T T^T
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(uu, uz + 8 * kk, dTt + (uint64_t)((kk * 2 * QB_LBO * 4) >> 4), I16, kk > 0);
            mma_commit(&sm.p_done);
        }
        __syncwarp();
        mbar_wait(&sm.p_done, ph); ph ^= 1;
        TICK(tm4);
        if (kTrain && c > 0) {
            mbar_wait(&sm.ut_ready, (c - 1) & 1);                      // U^T tile of chunk c-1 is in shared memory
            mbar_wait(&sm.st_free, (c - 1) & 1);                       // checkpoint c-1 has been read out of S^T
        }
        fence_after_sync();
        if (elect_one()) {
            // phase 2: S^ += U^T B~ + V^T K~ ;  Y^T += U^T Aqb^T
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(tb, uu + 8 * kk, dBt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb, dVt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), dKt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4),
                            I64, true);
            mma_commit(&sm.p_done);       // the chain only needs S^
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(uy, uu + 8 * kk, dQB + (uint64_t)((kk * 2 * QB_LBO * 4) >> 4), I16, true);
            if (!kTrain) mma_commit(&sm.empty[si]);
            mma_commit(&sm.y_ready[u]);
            if (kTrain && c > 0) {
                // transposed state, one chunk behind: S^T += B~^T U + K~^T V of chunk c-1 (checkpoint of chunk c)
                const Slot &Sp = sm.slot[(c - 1) % NSLOT];
                const uint64_t pBt = smem_desc(smem_u32(Sp.Bt), T_LBO * 4, T_SBO * 4);
                const uint64_t pKt = smem_desc(smem_u32(Sp.Kt), T_LBO * 4, T_SBO * 4);
                const uint64_t pVt = smem_desc(smem_u32(Sp.Vt), T_LBO * 4, T_SBO * 4);
                const uint64_t pUt = smem_desc(smem_u32(sm.Ut[(c - 1) & 1]), T_LBO * 4, T_SBO * 4);
#pragma unroll
                for (int kk = 0; kk < 2; kk++)
                    mma_tf32_ss(tb + C_ST, pBt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), pUt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4),
                                I64, true);
#pragma unroll
                for (int kk = 0; kk < 2; kk++)
                    mma_tf32_ss(tb + C_ST, pKt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), pVt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4),
                                I64, true);
                mma_commit(&sm.st_ready);
                mma_commit(&sm.empty[(c - 1) % NSLOT]);
            }
        }
        __syncwarp();
        mbar_wait(&sm.p_done, ph); ph ^= 1;
        TICK(tm5); ACC(8, tm0, tm1); ACC(9, tm1, tm2); ACC(10, tm2, tm3); ACC(11, tm3, tm4); ACC(12, tm4, tm5);
    }
}

// ---------------------------------------------------------------------------------------------
// epilogue group: warp q in [0,4) owns tensor-memory lanes 32q..32q+15 = value rows 16q..16q+15 of S^, U^T and Y^T.
// Training variant (as shipped): U goes to HBM in the backward's operand layout and to a [value][token] shared tile for
// the transposed-state update.
// ---------------------------------------------------------------------------------------------
template <bool kTrain>
__device__ void epilogue(const Params &P, Smem &sm, size_t base, size_t tok_stride, int bh, int nC, int tid) {
    long long *P_dbg = tid == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int q = tid >> 5, lane = tid & 31;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16);
    {   // initial state -> tensor memory
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
            if (P.s0 != nullptr) {
                const float4 *sp = reinterpret_cast<const float4 *>(P.s0 + (size_t)bh * kC * kC + row * kC + 16 * cb);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 x = sp[i];
                    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
                }
            }
            tmem_st16(tb + 16 * cb, v);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive_warp(&sm.win_scaled);
    }
    for (int c = 0; c < nC; c++) {
        const int u = c & 1;
        const bool last = (c == nC - 1);
        const bool win_end = (c % WIN == WIN - 1) || last;
        const float *dl = sm.DLw[(c / WIN) & 3];
        TICK(te0);
        mbar_wait(&sm.y_ready[u], (c >> 1) & 1);
        TICK(te1);
        fence_after_sync();
        float yv[16], uv[16];
        tmem_ld16(tb + C_ZY + C_ZY_STRIDE * u + 16, yv);
        if (kTrain) tmem_ld16(tb + C_ZY + C_ZY_STRIDE * u + 32, uv);
        tmem_wait_ld();
        if (win_end) {
            // state after this chunk, rescaled: the frame origin moves to the next window
            float *dsT = (last && P.sT != nullptr) ? P.sT + (size_t)bh * kC * kC + row * kC : nullptr;
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                float v[16];
                tmem_ld16(tb + 16 * cb, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] *= dl[16 * cb + i];
                if (!last) tmem_st16(tb + 16 * cb, v);
                if (act && dsT != nullptr) {
                    float4 *dp = reinterpret_cast<float4 *>(dsT + 16 * cb);
#pragma unroll
                    for (int i = 0; i < 4; i++) dp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            tmem_wait_st();
        }
        fence_before_sync();
        mbar_arrive_warp(&sm.y_free[u]);
        if (win_end) mbar_arrive_warp(&sm.win_scaled);
        {   // Y tile: [value lanes][16 tokens] -> shared [token][value] bf16 -> 128-byte rows to HBM
            bf16(&yb)[L][72] = sm.ybuf[u];
            if (act) {
#pragma unroll
                for (int j = 0; j < 16; j++) yb[j][row] = __float2bfloat16_rn(yv[j]);
                if (kTrain) {
#pragma unroll
                    for (int j = 0; j < 16; j++) uv[j] = tf32r(uv[j]);
                    float *ut = sm.Ut[u] + (row >> 3) * T_SBO + (row & 7) * 4;      // [value][token] operand tile
#pragma unroll
                    for (int i = 0; i < 4; i++) st4(ut + i * T_LBO, uv[4 * i], uv[4 * i + 1], uv[4 * i + 2], uv[4 * i + 3]);
                    // `sa` blob ([token][value] operand tile of the backward): element (token, value) at (value/4)*64 +
                    // (token/8)*32 + (token%8)*4 + value%4, i.e. the four values of a lane quad are adjacent for one token.
                    // 4x4 transposes inside the quad (two shuffle rounds per block of 4 tokens) give every lane the four
                    // values of ONE token: 4 x 16-byte stores per lane instead of 16 x 4-byte (256 -> 64 requests per chunk)
                    const int vi = lane & 3;
#pragma unroll
                    for (int blk = 0; blk < 4; blk++) {
                        float *r = &uv[4 * blk];
#pragma unroll
                        for (int j = 0; j < 4; j += 2) {
                            const float snd = (vi & 1) ? r[j] : r[j + 1];
                            const float rcv = __shfl_xor_sync(0x0000ffffu, snd, 1);
                            if (vi & 1) r[j] = rcv; else r[j + 1] = rcv;
                        }
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const float snd = (vi & 2) ? r[j] : r[j + 2];
                            const float rcv = __shfl_xor_sync(0x0000ffffu, snd, 2);
                            if (vi & 2) r[j] = rcv; else r[j + 2] = rcv;
                        }
                    }
                    // now uv[4*blk + j] = U[token 4*blk + vi][value 4*(row/4) + j]
                    float *sap = P.sa + ((size_t)bh * nC + c) * kUFloats + (row >> 2) * kULbo;
#pragma unroll
                    for (int blk = 0; blk < 4; blk++) {
                        const int tok = 4 * blk + vi;
                        st4(sap + (tok >> 3) * 32 + (tok & 7) * 4, uv[4 * blk], uv[4 * blk + 1], uv[4 * blk + 2], uv[4 * blk + 3]);
                    }
                }
            }
            if (kTrain) {
                fence_proxy_async();
                mbar_arrive_warp(&sm.ut_ready);
            }
            bar_sync(4, 128);
            const int tok = tid >> 3, part = tid & 7;
            const uint4 v = *reinterpret_cast<const uint4 *>(&yb[tok][part * 8]);
            *reinterpret_cast<uint4 *>(P.y + base + (size_t)(c * L + tok) * tok_stride + part * 8) = v;
        }
        TICK(te2); ACC(13, te0, te1); ACC(14, te1, te2);
    }
}

// ---------------------------------------------------------------------------------------------
// checkpoint group (training variant only): warp q owns tensor-memory lanes 32q..32q+15 = KEY rows 16q..16q+15 of
// the transposed state S^T.  After the MMA warp has added chunk c, S^T is the chunk-start checkpoint of chunk c+1:
// 16-byte pieces of the backward's K-major operand tile, 256 contiguous bytes per warp store.
// ---------------------------------------------------------------------------------------------
__device__ void ckpt_group(const Params &P, Smem &sm, int bh, int nC, int tid) {
    const int q = (tid >> 5) & 3, lane = tid & 31;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16) + C_ST;
    float *ckg = P.ckT + (size_t)bh * nC * kCkFloats + (row >> 3) * 32 + (row & 7) * 4;   // key = row
    auto store_ck = [&](const float (&v)[16], int cb, int cc) {
        if (act) {
            float *dst = ckg + (size_t)cc * kCkFloats + (4 * cb) * kCkLbo;
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<float4 *>(dst + i * kCkLbo) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    };
#pragma unroll
    for (int cb = 0; cb < 4; cb++) {   // S^T <- s0^T (or 0) = checkpoint 0
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i++)
            v[i] = (P.s0 != nullptr) ? P.s0[(size_t)bh * kC * kC + (16 * cb + i) * kC + row] : 0.f;
        tmem_st16(tb + 16 * cb, v);
        store_ck(v, cb, 0);
    }
    tmem_wait_st();
    fence_before_sync();
    mbar_arrive_warp(&sm.st_free);
    for (int c = 0; c + 1 < nC; c++) {
        const bool win_end = (c % WIN == WIN - 1);
        mbar_wait(&sm.st_ready, c & 1);
        fence_after_sync();
        const float dr = win_end ? sm.DLw[(c / WIN) & 3][row] : 1.f;   // window end: rows move to the next window's frame
        float v[4][16];
#pragma unroll
        for (int cb = 0; cb < 4; cb++) tmem_ld16(tb + 16 * cb, v[cb]);
        tmem_wait_ld();
        if (win_end) {
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[cb][i] *= dr;
                tmem_st16(tb + 16 * cb, v[cb]);
            }
            tmem_wait_st();
        }
        fence_before_sync();
        mbar_arrive_warp(&sm.st_free);
#pragma unroll
        for (int cb = 0; cb < 4; cb++) store_ck(v[cb], cb, c + 1);
    }
}

constexpr int kMmaWarp = 20, kThreads = 32 * (kMmaWarp + 1), kThreadsTrain = kThreads + 128;

template <bool kTrain>
__global__ void __launch_bounds__(kTrain ? kThreadsTrain : kThreads, 1) wkv7_tc_fwd_v2_kernel(const Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int bh = blockIdx.x, bb = bh / P.H, hh = bh % P.H;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nC = P.T / L;
    const size_t tok_stride = (size_t)P.H * kC;
    const size_t base = (size_t)bb * P.T * tok_stride + (size_t)hh * kC;
    constexpr uint32_t kCols = kTrain ? kTmemColsTrain : kTmemCols;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; i++) { mbar_init(&sm.empty[i], 1); mbar_init(&sm.full[i], 2); mbar_init(&sm.a_done[i], 8); }
        mbar_init(&sm.p_done, 1);
        for (int i = 0; i < 2; i++) { mbar_init(&sm.y_ready[i], 1); mbar_init(&sm.y_free[i], 4); mbar_init(&sm.g_ready[i], 1); }
        mbar_init(&sm.win_scaled, 4); mbar_init(&sm.ut_ready, 4); mbar_init(&sm.st_ready, 1); mbar_init(&sm.st_free, 4);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, kCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    // tensor-memory lane quadrants follow warp % 4: the Gram groups need one warp on lanes 0-31 and one on lanes 32-63
    if (warp < 4) epilogue<kTrain>(P, sm, base, tok_stride, bh, nC, tid);
    else if (warp < 12) stage_a(P, sm, base, tok_stride, nC, tid - 128);
    else if (warp == 12 || warp == 13) stage_g(P, sm, nC, tid - 384, 0);
    else if (warp == 16 || warp == 17) stage_g(P, sm, nC, tid - 512, 1);
    else if (warp == kMmaWarp) mma_warp<kTrain>(P, sm, nC);
    else if (kTrain && warp > kMmaWarp) ckpt_group(P, sm, bh, nC, tid);
    // warps 14, 15, 18, 19 only keep the warp numbering aligned (to be dropped with a renumbering)

    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(sm.tmem_base, kCols);
}

}  // namespace tcfwd2

// ckT == nullptr: snapshot-free forward; otherwise the training forward (same scratch contract as launch_tc_fwd)
cudaError_t launch_tc_fwd_v2(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                             const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                             cudaStream_t st) {
    using namespace tcfwd2;
    static_assert(sizeof(Smem) <= 232448, "shared memory budget");
    Params P{T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)a,
             (const bf16 *)b, (bf16 *)y, ckT, sa, s0, sT, nullptr};
    if (ckT != nullptr) {
        cudaError_t e = cudaFuncSetAttribute(wkv7_tc_fwd_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        wkv7_tc_fwd_v2_kernel<true><<<dim3(B * H), dim3(kThreadsTrain), sizeof(Smem), st>>>(P);
    } else {
        cudaError_t e = cudaFuncSetAttribute(wkv7_tc_fwd_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        wkv7_tc_fwd_v2_kernel<false><<<dim3(B * H), dim3(kThreads), sizeof(Smem), st>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace rwkvtts
