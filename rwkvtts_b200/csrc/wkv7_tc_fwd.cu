// WKV-7 training forward for sm_100a: chunked DPLR form on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators and the recurrent state in tensor memory).
//
// Reference operator: model/llm/cuda/wkv7_cuda.cu:17-42 (forward_kernel); same recurrence, evaluated
// 16 tokens (one chunk) at a time.  S is the value-major state [64 values][64 keys].
//
// Frames.  Tokens are grouped in windows of 64 (4 chunks).  Inside a window every vector is scaled
// by the decay accumulated since the window start, G_t = sum_{s<=t} log d_s (G < 0):
//     Q~ = q e^{G_t}    A~ = a e^{G_{t-1}}    K~ = k e^{-G_t}    B~ = b e^{-G_t}
// and the state is kept as S^ = S diag(e^{-G}) so that inside a window it is only ACCUMULATED:
//     per chunk c:  N = stril(A~ B~^T)  Aak = stril(A~ K~^T)  Aqb = tril(Q~ B~^T)  Aqk = tril(Q~ K~^T)
//                   T = (I - N)^-1,  W~ = T A~,  M1 = T Aak                          (CUDA cores / mma.sync)
//         phase 1:  [U^T | Y^T] = S^ [W~ | Q~]^T + V^T [M1 | Aqk]^T                  (tcgen05, N = 32)
//         phase 2:  S^ += U^T B~ + V^T K~ ;   Y^T += U^T Aqb^T                       (tcgen05, N = 64 / 16)
// (U_t = S_{t-1} a_t is the reference's `sa`).  At a window end S^ is multiplied by e^{G_64}
// (tensor memory -> registers -> tensor memory).  The training variant also writes, per chunk, the
// TRANSPOSED state at the chunk start (window frame) and `sa` for the chunked backward (wkv7_tc_bwd.cu), as dense
// tensor-core operand tiles (wkv7_common.cuh) inside the caller's scratch tensors `s` / `sa` (reference sizes); the
// transposed state is kept up to date by the tensor core itself (S^T += B~^T U + K~^T V, one chunk behind).
// |G| <= 64 * 1.35 stays inside the fp32 exponent range; log-decays below -1.35
// per step (decay < 0.26; the model's range is (-0.6065, 0), rwkv_s2s_single_ffn.py:172) are clamped.
//
// tcgen05 facts this kernel relies on (measured with tests/csrc/umma_probe.cu on a B200):
//   * M = 64 accumulators put row 16q+i in lane 32q+i; an M = 64 A operand in tensor memory uses the
//     same lanes, so an accumulator (S^, U^T) is directly the A operand of the next product;
//   * an MMA that reads tensor memory written by an MMA with a DIFFERENT accumulator is not ordered
//     behind it: every phase boundary is a tcgen05.commit + mbarrier wait (~170 cycles);
//   * tf32 operands in shared memory must be K-major (no-swizzle canonical layout, tc05.cuh).
//
// One CTA per (batch, head), 21 warps (25 in the training variant), one mbarrier arrival per warp at every hand-off:
//   warps  0-3   epilogue: Y^T tensor memory -> bf16 -> HBM, window-end rescale of S^, s0 / sT; training: U to HBM in
//                the backward's operand layout and as the [value][token] operand tile of the transposed-state update
//   warps  4-11  stage A: HBM loads, decay scan (log2 domain, two-stage cross-warp prefix), scaled operands in both
//                orientations
//   warps 12-15  stage B (even chunks), warps 16-19 stage B (odd chunks): Gram blocks (mma.sync tf32),
//                triangular solve, W~ / M1 / Aqk / Aqb into the canonical operand slot
//   warp   20    MMA issuer (one elected lane), owns the tensor-memory allocation; training: also S^T += B~^T U + K~^T V
//   warps 21-24  (training) checkpoint group: reads S^T with keys on the lanes -> 16-byte pieces of the backward's
//                K-major S0^T operand tile, 256 contiguous bytes per warp store
// The operand slots form a ring of NSLOT chunks.  Bound by the shared-memory data pipe (~75-80 % of peak, ncu).
#include "mma_tf32.cuh"
#include "tc05.cuh"
#include "wkv7_common.cuh"

namespace rwkvtts {
namespace tcfwd {
using namespace tc05;

constexpr int L = 16;        // chunk length
constexpr int WIN = 4;       // chunks per window
constexpr int NSLOT = 5;     // operand slots in flight
constexpr int NNAT = 3;      // stage A -> stage B hand-off buffers
constexpr int LDN = 68;      // row stride of the natural [token][channel] tiles
constexpr float kMinLogDecay = -1.35f;
constexpr uint32_t C_ST = 128;   // tensor-memory columns of the transposed state (training variant)
constexpr float kLog2e = 1.4426950408889634f;   // decays are accumulated as log2 (ex2.approx needs no pre-scale)

// canonical K-major tiles, strides in floats (see tc05.cuh: off = (r/8)*SBO + (k/4)*LBO + (r%8)*4 + k%4)
constexpr int WQ_LBO = 132, WQ_SBO = 32;     // [32 rows: 0-15 W~ tokens, 16-31 Q~ tokens][64 channels]
constexpr int T_SBO = 36, T_LBO = 288;       // [64 rows: channel / value][16 tokens]   (transposed tiles)
constexpr int MA_LBO = 132, MA_SBO = 32;     // [32 rows: 0-15 M1, 16-31 Aqk][16]; +4: stage B writes columns conflict-free
constexpr int QB_LBO = 64, QB_SBO = 32;      // [16][16]

struct Slot {
    float WQ[16 * WQ_LBO];
    float Bt[4 * T_LBO], Kt[4 * T_LBO], Vt[4 * T_LBO];
    float MA[4 * MA_LBO];
    float Aqb[4 * QB_LBO];
};
struct Nat {
    float At[L * LDN], Bn[L * LDN], Kn[L * LDN], Qn[L * LDN];
};
struct Smem {
    Slot slot[NSLOT];
    Nat nat[NNAT];
    float NT[2][L * 20], Aak[2][L * 20];   // per stage-B group: N^T and Aak, fp32
    float wtot[9][kC];                     // stage A scan: per-warp totals -> exclusive prefixes, chunk total
    __align__(16) bf16 ybuf[2][L][72];     // epilogue: Y tile [token][value], double buffered
    __align__(16) float Ut[2][4 * T_LBO];  // training: U^T [value][token] operand tile of the transposed-state update
    float DLw[4][kC];                      // e^{G} at the end of a window (ring of 4 windows)
    uint64_t empty[NSLOT], full[NSLOT], a_done[NNAT], nat_empty[NNAT];
    uint64_t p_done, y_ready[2], y_free[2], win_scaled, ut_ready, st_ready, st_free;
    uint32_t tmem_base;
};

struct Params {
    int T, H;
    const bf16 *w, *q, *k, *v, *a, *b;
    bf16 *y;
    float *ckT;          // training: TRANSPOSED state at the start of every chunk, in the frame of the chunk's
                         // window, [B*H][T/16] operand tiles of 4096 floats (wkv7_common.cuh); null for the
                         // snapshot-free forward
    float *sa;           // training: U_t = S_{t-1} a_t (tf32), [B*H][T/16] operand tiles of 1024 floats
    const float *s0;     // may be null
    float *sT;           // may be null
    long long *dbg;      // phase-cycle counters (profiling builds only), may be null
    const int *cu;       // packed launch: cu_seqlens [N+1] (device), else null
    const int *cbase;    // packed launch: first chunk slot of every sequence [N+1] (exclusive prefix of ceil(len/16))
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
// stage A: tp in [0,256); token t = tp>>4, channels 4*k4 .. 4*k4+3; warp wp holds tokens 2wp, 2wp+1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 ldg_nc_v2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void unpack4(const uint2 &u, float *f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
}
template <bool kVar>
__device__ __forceinline__ void load_raw(const Params &P, size_t base, size_t tok_stride, int c, int tp, int len,
                                         uint2 (&raw)[6]) {
    const int t = tp >> 4, k4 = tp & 15;
    if (kVar && c * L + t >= len) {          // beyond the end of a packed sequence: a token that changes nothing and is never stored
#pragma unroll
        for (int i = 0; i < 6; i++) raw[i] = make_uint2(0u, 0u);
        return;
    }
    const size_t off = base + (size_t)(c * L + t) * tok_stride + k4 * 4;
    raw[0] = ldg_nc_v2(P.w + off);
    raw[1] = ldg_nc_v2(P.q + off);
    raw[2] = ldg_nc_v2(P.k + off);
    raw[3] = ldg_nc_v2(P.v + off);
    raw[4] = ldg_nc_v2(P.a + off);
    raw[5] = ldg_nc_v2(P.b + off);
}

template <bool kVar>
__device__ void stage_a(const Params &P, Smem &sm, size_t base, size_t tok_stride, int nC, int len, int tp) {
    long long *P_dbg = tp == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int t = tp >> 4, k4 = tp & 15, wp = tp >> 5;
    uint2 raw[6], nxt[6], nx2[6];
    float gpre[4];
#pragma unroll
    for (int j = 0; j < 4; j++) gpre[j] = 0.f;
    load_raw<kVar>(P, base, tok_stride, 0, tp, len, raw);
    if (nC > 1) load_raw<kVar>(P, base, tok_stride, 1, tp, len, nxt);
    for (int c = 0; c < nC; c++) {
        const int si = c % NSLOT, ni = c % NNAT;
        Slot &S = sm.slot[si];
        Nat &N = sm.nat[ni];
        if (c + 2 < nC) load_raw<kVar>(P, base, tok_stride, c + 2, tp, len, nx2);   // two chunks ahead
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
        for (int j = 0; j < 4; j++) {   // inclusive scan over the 2 tokens of this warp (lane = (t&1)*16 + k4)
            const float x = __shfl_up_sync(0xffffffffu, gg[j], 16);
            if (t & 1) gg[j] += x;
        }
        // cross-warp prefix in two stages: 8 per-warp totals per channel -> 64 threads turn them into exclusive
        // prefixes (+ the chunk total in row 8) -> every thread reads two rows (the one-stage version had every thread
        // read all 8 rows: 256 of the kernel's ~1660 shared-memory wavefronts per chunk)
        float(&wt)[9][kC] = sm.wtot;
        if (t & 1) st4(&wt[wp][k4 * 4], gg[0], gg[1], gg[2], gg[3]);
        bar_sync(1, 256);
        if (tp < kC) {
            float run = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ww++) {
                const float x = wt[ww][tp];
                wt[ww][tp] = run;
                run += x;
            }
            wt[8][tp] = run;
        }
        bar_sync(1, 256);
        float tot[4];
        {
            const float4 pre = *reinterpret_cast<const float4 *>(&wt[wp][k4 * 4]);
            const float4 all = *reinterpret_cast<const float4 *>(&wt[8][k4 * 4]);
            gg[0] += gpre[0] + pre.x; gg[1] += gpre[1] + pre.y; gg[2] += gpre[2] + pre.z; gg[3] += gpre[3] + pre.w;
            tot[0] = gpre[0] + all.x; tot[1] = gpre[1] + all.y; tot[2] = gpre[2] + all.z; tot[3] = gpre[3] + all.w;
        }
        bar_sync(1, 256);       // rows are rewritten by the next chunk
        const bool win_end = (c % WIN == WIN - 1) || (c == nC - 1);
#pragma unroll
        for (int j = 0; j < 4; j++) gpre[j] = win_end ? 0.f : tot[j];

        // wait until the slot / hand-off buffer of this chunk have been released
        TICK(ta1);
        if (c >= NSLOT) mbar_wait(&sm.empty[si], ((c / NSLOT) - 1) & 1);
        TICK(ta2);
        if (c >= NNAT) mbar_wait(&sm.nat_empty[ni], ((c / NNAT) - 1) & 1);
        TICK(ta3);
        {
            float D[4], Dp[4], iD[4], f[4], o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                D[j] = ex2f(gg[j]);
                Dp[j] = ex2f(gg[j] - lw[j]);
                iD[j] = ex2f(-gg[j]);
            }
            const int on = t * LDN + k4 * 4;
            // Q~: natural + canonical rows 16..31 of WQ
            unpack4(raw[1], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * D[j]);
            st4(&N.Qn[on], o[0], o[1], o[2], o[3]);
            {
                const int r = 16 + t;
                st4(&S.WQ[(r >> 3) * WQ_SBO + k4 * WQ_LBO + (r & 7) * 4], o[0], o[1], o[2], o[3]);
            }
            // transposed tiles: row = channel 4*k4+j, column = token t
            const int ot = (k4 >> 1) * T_SBO + (t >> 2) * T_LBO + (k4 & 1) * 16 + (t & 3);   // + 4*j
            // K~
            unpack4(raw[2], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * iD[j]);
            st4(&N.Kn[on], o[0], o[1], o[2], o[3]);
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
            st4(&N.At[on], o[0], o[1], o[2], o[3]);
            // B~
            unpack4(raw[5], f);
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = tf32r(f[j] * iD[j]);
            st4(&N.Bn[on], o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) S.Bt[ot + 4 * j] = o[j];
            if (win_end && t == L - 1) st4(&sm.DLw[(c / WIN) & 3][k4 * 4], D[0], D[1], D[2], D[3]);
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.a_done[ni]);
#pragma unroll
        for (int i = 0; i < 6; i++) { raw[i] = nxt[i]; nxt[i] = nx2[i]; }
        TICK(ta4); ACC(0, ta0, ta1); ACC(1, ta1, ta2); ACC(2, ta2, ta3); ACC(3, ta3, ta4);
    }
}

// ---------------------------------------------------------------------------------------------
// stage B: tp in [0,128), group grp (0: even chunks, 1: odd chunks)
// ---------------------------------------------------------------------------------------------
__device__ void stage_b(const Params &P, Smem &sm, int nC, int tp, int grp) {
    long long *P_dbg = (grp == 0 && tp == 0) ? P.dbg : nullptr; (void)P_dbg;
    const int wp = tp >> 5, lane = tp & 31, g = lane >> 2, tq = lane & 3;
    float *NT = sm.NT[grp], *AK = sm.Aak[grp];
    for (int c = grp; c < nC; c += 2) {
        const int si = c % NSLOT, ni = c % NNAT;
        Slot &S = sm.slot[si];
        const Nat &N = sm.nat[ni];
        TICK(tb0);
        mbar_wait(&sm.a_done[ni], (c / NNAT) & 1);
        TICK(tb1);
        // ---- Gram blocks, one 16x16 block per warp ------------------------------------------
        {
            const int rowsel = wp & 1, colsel = wp >> 1;
            const float *Ar = rowsel ? N.Qn : N.At;
            const float *Bc = colsel ? N.Kn : N.Bn;
            float acc[2][4] = {};
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                uint32_t af[4], bfr[2];
                lda(af, Ar, LDN, 1, 0, 8 * kb, g, tq);
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    ldb(bfr, Bc, 1, LDN, 8 * kb, 8 * nt, g, tq);
                    mma_tf32(acc[nt], af, bfr);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = 8 * nt + 2 * tq + e;      // s
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int row = g + 8 * hh;           // t
                        float x = acc[nt][2 * hh + e];
                        if (rowsel) {   // Q~ rows: inclusive lower triangle, tensor-core operand
                            x = (col <= row) ? tf32r(x) : 0.f;
                            if (colsel) S.MA[kmajor_off(16 + row, col, MA_LBO, MA_SBO)] = x;
                            else S.Aqb[kmajor_off(row, col, QB_LBO, QB_SBO)] = x;
                        } else {        // A~ rows: strict lower triangle, fp32 for the solve
                            x = (col < row) ? x : 0.f;
                            if (colsel) AK[row * 20 + col] = x;
                            else NT[col * 20 + row] = x;
                        }
                    }
                }
        }
        bar_sync(2 + grp, 128);
        TICK(tb2);
        // ---- [M1 | W~] = (I - N)^-1 [Aak | A~], one column per thread, column-oriented ----------
        if (tp < 16 + kC) {
            const int col = tp;
            float acc[L];
#pragma unroll
            for (int tt = 0; tt < L; tt++) acc[tt] = (col < 16) ? AK[tt * 20 + col] : N.At[tt * LDN + col - 16];
#pragma unroll
            for (int s = 0; s < L - 1; s++) {
                const float x = acc[s];
#pragma unroll
                for (int q4 = (s + 1) / 4; q4 < 4; q4++) {
                    const float4 n4 = *reinterpret_cast<const float4 *>(&NT[s * 20 + 4 * q4]);
                    const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (4 * q4 + e > s) acc[4 * q4 + e] = fmaf(nn[e], x, acc[4 * q4 + e]);
                }
            }
            if (col < 16) {
#pragma unroll
                for (int tt = 0; tt < L; tt++) S.MA[kmajor_off(tt, col, MA_LBO, MA_SBO)] = tf32r(acc[tt]);
            } else {
#pragma unroll
                for (int tt = 0; tt < L; tt++) S.WQ[kmajor_off(tt, col - 16, WQ_LBO, WQ_SBO)] = tf32r(acc[tt]);
            }
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.full[si]);
        mbar_arrive_warp(&sm.nat_empty[ni]);
        TICK(tb3);
        bar_sync(2 + grp, 128);     // NT / Aak are reused by the next chunk of this group
        TICK(tb4); ACC(4, tb0, tb1); ACC(5, tb1, tb2); ACC(6, tb2, tb3); ACC(7, tb3, tb4);
    }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer (one warp)
// ---------------------------------------------------------------------------------------------
template <bool kTrain>
__device__ void mma_warp(const Params &P, Smem &sm, int nC) {
    long long *P_dbg = (threadIdx.x & 31) == 0 ? P.dbg : nullptr; (void)P_dbg;
    const uint32_t tb = sm.tmem_base;
    constexpr uint32_t I16 = idesc_tf32(64, 16, false, false);
    constexpr uint32_t I32 = idesc_tf32(64, 32, false, false);
    constexpr uint32_t I64 = idesc_tf32(64, 64, false, false);
    uint32_t ph = 0;
    for (int c = 0; c < nC; c++) {
        const int si = c % NSLOT, u = c & 1;
        const Slot &S = sm.slot[si];
        const uint32_t uy = tb + 64 + 32 * u;
        const uint64_t dWQ = smem_desc(smem_u32(S.WQ), WQ_LBO * 4, WQ_SBO * 4);
        const uint64_t dVt = smem_desc(smem_u32(S.Vt), T_LBO * 4, T_SBO * 4);
        const uint64_t dBt = smem_desc(smem_u32(S.Bt), T_LBO * 4, T_SBO * 4);
        const uint64_t dKt = smem_desc(smem_u32(S.Kt), T_LBO * 4, T_SBO * 4);
        const uint64_t dMA = smem_desc(smem_u32(S.MA), MA_LBO * 4, MA_SBO * 4);
        const uint64_t dQB = smem_desc(smem_u32(S.Aqb), QB_LBO * 4, QB_SBO * 4);
        TICK(tm0);
        mbar_wait(&sm.full[si], (c / NSLOT) & 1);
        TICK(tm1);
        if (c % WIN == 0) mbar_wait(&sm.win_scaled, (c / WIN) & 1);
        TICK(tm2);
        if (c >= 2) mbar_wait(&sm.y_free[u], ((c >> 1) - 1) & 1);
        TICK(tm3);
        fence_after_sync();
        if (elect_one()) {
            // phase 1: [U^T | Y^T] = S^ [W~ | Q~]^T + V^T [M1 | Aqk]^T
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ts(uy, tb + 8 * kk, dWQ + (uint64_t)((kk * 2 * WQ_LBO * 4) >> 4), I32, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(uy, dVt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), dMA + (uint64_t)((kk * 2 * MA_LBO * 4) >> 4),
                            I32, true);
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
                mma_tf32_ts(tb, uy + 8 * kk, dBt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb, dVt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4), dKt + (uint64_t)((kk * 2 * T_LBO * 4) >> 4),
                            I64, true);
            mma_commit(&sm.p_done);       // the chain only needs S^
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(uy + 16, uy + 8 * kk, dQB + (uint64_t)((kk * 2 * QB_LBO * 4) >> 4), I16, true);
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
// epilogue group: warp q in [0,4) owns tensor-memory lanes 32q..32q+15 = rows 16q..16q+15 (values for S^, U^T,
// Y^T; keys for S^T).  Training variant, per chunk: U goes to HBM in the backward's operand layout and to a
// [value][token] shared tile from which the MMA warp keeps a TRANSPOSED copy of the state up to date
// (S^T += B~^T U + K~^T V, one chunk behind the main chain); the chunk-start checkpoint is that copy, read
// with keys on the lanes, so it leaves as 16-byte pieces of the backward's K-major operand tile.
// ---------------------------------------------------------------------------------------------
template <bool kTrain, bool kVar>
__device__ void epilogue(const Params &P, Smem &sm, size_t base, size_t tok_stride, int bh, size_t ck0, int nC, int len,
                         int tid) {
    long long *P_dbg = tid == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int q = tid >> 5, lane = tid & 31;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16);
    float *sag = kTrain ? P.sa + ck0 * kUFloats + (row >> 2) * kULbo + (row & 3) : nullptr;     // value = row
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
        tmem_ld16(tb + 64 + 32 * u + 16, yv);
        if (kTrain) tmem_ld16(tb + 64 + 32 * u, uv);
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
        TICK(te1a);
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
                    float *sap = sag + (size_t)c * kUFloats;                         // [token][value] operand tile
#pragma unroll
                    for (int j = 0; j < 16; j++) sap[(j >> 3) * 32 + (j & 7) * 4] = uv[j];
                }
            }
            TICK(te1b);
            // (handing the U^T tile over BEFORE the Y staging and the 16 strided stores of U was measured: the hand-off
            // comes 750 cycles earlier and the kernel takes the same time -- the training variant is slower than the no-grad
            // one in every role by the same 40 %, not on this chain; profiles/r02_tc_pair_roles.txt)
            if (kTrain) {
                fence_proxy_async();
                mbar_arrive_warp(&sm.ut_ready);
            }
            TICK(te1c);
            bar_sync(4, 128);
            TICK(te1d);
            ACC(16, te1, te1a); ACC(17, te1a, te1b); ACC(18, te1b, te1c); ACC(19, te1c, te1d);
            const int tok = tid >> 3, part = tid & 7;
            const uint4 v = *reinterpret_cast<const uint4 *>(&yb[tok][part * 8]);
            if (!kVar || c * L + tok < len) *reinterpret_cast<uint4 *>(P.y + base + (size_t)(c * L + tok) * tok_stride + part * 8) = v;
        }
        TICK(te2); ACC(13, te0, te1); ACC(14, te1, te2);
    }
}

// ---------------------------------------------------------------------------------------------
// checkpoint group (training variant only): warp q owns tensor-memory lanes 32q..32q+15 = KEY rows 16q..16q+15 of
// the transposed state S^T.  After the MMA warp has added chunk c, S^T is the chunk-start checkpoint of chunk c+1:
// 16-byte pieces of the backward's K-major operand tile, 256 contiguous bytes per warp store.
// ---------------------------------------------------------------------------------------------
__device__ void ckpt_group(const Params &P, Smem &sm, int bh, size_t ck0, int nC, int tid) {
    const int q = (tid >> 5) & 3, lane = tid & 31;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16) + C_ST;
    float *ckg = P.ckT + ck0 * kCkFloats + (row >> 3) * 32 + (row & 7) * 4;   // key = row
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

template <bool kTrain, bool kVar>
__global__ void __launch_bounds__(kTrain ? kThreadsTrain : kThreads, 1) wkv7_tc_fwd_kernel(const Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const SeqWork W = seq_work(P.T, P.H, kVar ? P.cu : nullptr, P.cbase);
    const int bh = W.bh, nC = W.nC;
    const int tid = threadIdx.x, warp = tid >> 5;
    const size_t tok_stride = (size_t)P.H * kC;
    const size_t base = W.base;
    if (kVar && nC == 0) return;         // an empty sequence of a packed launch (uniform over the CTA)

    if (tid == 0) {
        for (int i = 0; i < NSLOT; i++) { mbar_init(&sm.empty[i], 1); mbar_init(&sm.full[i], 4); }
        for (int i = 0; i < NNAT; i++) { mbar_init(&sm.a_done[i], 8); mbar_init(&sm.nat_empty[i], 4); }
        mbar_init(&sm.p_done, 1);
        for (int i = 0; i < 2; i++) { mbar_init(&sm.y_ready[i], 1); mbar_init(&sm.y_free[i], 4); }
        mbar_init(&sm.win_scaled, 4); mbar_init(&sm.ut_ready, 4); mbar_init(&sm.st_ready, 1); mbar_init(&sm.st_free, 4);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, 256);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    if (warp < 4) epilogue<kTrain, kVar>(P, sm, base, tok_stride, bh, W.ck0, nC, W.len, tid);
    else if (warp < 12) stage_a<kVar>(P, sm, base, tok_stride, nC, W.len, tid - 128);
    else if (warp < 16) stage_b(P, sm, nC, tid - 384, 0);
    else if (warp < 20) stage_b(P, sm, nC, tid - 512, 1);
    else if (warp == kMmaWarp) mma_warp<kTrain>(P, sm, nC);
    else if (kTrain) ckpt_group(P, sm, bh, W.ck0, nC, tid);

    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(sm.tmem_base, 256);
}

}  // namespace tcfwd

long long *g_tc_dbg = nullptr;   // set by the profiling harness only

const char *tc_fwd_barrier_name(unsigned off) {
    using tcfwd::Smem;
    struct { size_t off; int n; const char *name; } t[] = {
        {offsetof(Smem, empty), tcfwd::NSLOT, "empty[slot]"}, {offsetof(Smem, full), tcfwd::NSLOT, "full[slot]"},
        {offsetof(Smem, a_done), tcfwd::NNAT, "a_done[nat]"}, {offsetof(Smem, nat_empty), tcfwd::NNAT, "nat_empty[nat]"},
        {offsetof(Smem, p_done), 1, "p_done"}, {offsetof(Smem, y_ready), 2, "y_ready[parity]"},
        {offsetof(Smem, y_free), 2, "y_free[parity]"}, {offsetof(Smem, win_scaled), 1, "win_scaled"},
        {offsetof(Smem, ut_ready), 1, "ut_ready"}, {offsetof(Smem, st_ready), 1, "st_ready"}, {offsetof(Smem, st_free), 1, "st_free"}};
    for (auto &e : t)
        if (off >= e.off && off < e.off + 8 * (size_t)e.n) return e.name;
    return "unknown";
}

// ckT == nullptr: snapshot-free forward.  Otherwise the training forward: per-chunk transposed state
// checkpoints (same size as the reference's `s`) and `sa`, consumed by wkv7_tc_bwd.cu.
cudaError_t launch_tc_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                          const int *cu, const int *cbase, cudaStream_t st) {
    using namespace tcfwd;
    static_assert(sizeof(Smem) <= 232448, "shared memory budget");
    Params P{T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)a,
             (const bf16 *)b, (bf16 *)y, ckT, sa, s0, sT, g_tc_dbg, cu, cbase};
    if (watchdog_needs_install(0, st)) {
        cudaError_t e = watchdog_install(watchdog_record(), 1);
        if (e != cudaSuccess) return e;
    }
    count_launch();
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        kern<<<dim3(B * H), dim3(threads), sizeof(Smem), st>>>(P);
        return cudaSuccess;
    };
    cudaError_t e;
    if (ckT != nullptr) e = cu ? go(wkv7_tc_fwd_kernel<true, true>, kThreadsTrain) : go(wkv7_tc_fwd_kernel<true, false>, kThreadsTrain);
    else e = cu ? go(wkv7_tc_fwd_kernel<false, true>, kThreads) : go(wkv7_tc_fwd_kernel<false, false>, kThreads);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace rwkvtts
