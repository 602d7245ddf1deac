// WKV-7 training backward for sm_100a: chunked DPLR adjoint on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32; dS, dS^T and all gradient tiles in tensor memory, S0^T as a shared-memory operand).
//
// Reference operator: the exact adjoint of model/llm/cuda/wkv7_cuda.cu:17-42 (backward_kernel :54-130).
// Unlike the reference it never un-steps the state (no division by the decay): the forward
// (wkv7_tc_fwd.cu, training variant) leaves, per 16-token chunk, the TRANSPOSED state at the chunk start
// (window frame) and U_t = S_{t-1} a_t (`sa`); everything else is recomputed.  Algebra validated in f64
// against the oracle in proto/tc_bwd_proto.py (1e-15) and with tf32 operand rounding (<= 6e-4).
//
// Frame: as in the forward (windows of 64 tokens, G = log-decay accumulated since the window start):
//     Q~ = q e^{G}  A~ = a e^{G_{t-1}}  K~ = k e^{-G}  B~ = b e^{-G},   S^ = S diag(e^{-G}).
// Per chunk (processed last to first), with S0 = state at the chunk start, dS = gradient w.r.t. the state
// at the chunk end, N/Aak/Aqb/Aqk the forward's 16x16 Gram blocks and T = (I-N)^-1:
//     B' = T^T B~,  Aqb' = Aqb T                                          (stage B: back substitution)
//     Z^T  = dY^T Aqb' + dS B'^T                                          (R1)   Z = dL/dU after the solve
//     dN = stril(Z U^T)  dAak = stril(Z V^T)  dAqb = tril(dY U^T)  dAqk = tril(dY V^T)   (mma.sync, 16x16)
//     dQ~^T|dA~^T = S0^T [dY;Z]^T + B~^T [dAqb;dN]^T + K~^T [dAqk;dAak]^T  (P1)
//     dB~^T|dK~^T = dS^T [U;V]^T  + A~^T [dN^T;dAak^T]^T + Q~^T [dAqb^T;dAqk^T]^T   (P2)
//     dV^T        = dS K~^T + dY^T Aqk + Z^T Aak                           (P3)
//     dS  += dY^T Q~ + Z^T A~ ;   dS^T += Q~^T dY + A~^T Z                  (R2)
// then dq = dQ~ e^{G}, da = dA~ e^{G_{t-1}}, db = dB~ e^{-G}, dk = dK~ e^{-G} and, with
//     g_t = dQ~.Q~ - dK~.K~ - dB~.B~ + (dA~.A~)_{t+1}  (+ sum_v dS.S at a window end),
// dw_t = (-e^{w_t}) * sum_{s >= t in the window} g_s.   All M = 64 products put a channel / value index on
// the tensor-memory lanes, so every gradient tile comes out as [channel][16 tokens] and the dw scan is
// thread-local.
//
// One CTA per (batch, head), 25 warps, every hand-off an mbarrier with one arrival per warp:
//   warps  0-3   group C1 (on the chain): window-boundary rescale of dS / dS^T in tensor memory, Z^T -> shared operand
//                tiles, the four 16x16 gradient Gram blocks (mma.sync)
//   warps  4-11  group C2 (off the chain): output epilogue of chunk `it` while chunk `it+1` is in flight -- gradient
//                accumulators double buffered in tensor memory; scaling, dw suffix scan, bf16, staging in slot tiles
//                that are dead by then, 128-byte rows; window-boundary term sum_v dS.S
//   warps 12-19  stage A: the 7 bf16 input tiles of a chunk land by tensor-map copies (cp.async.bulk.tensor, one chunk
//                ahead), the window's decays (one window ahead), decay prefix, ten operand tiles; U (`sa`) arrives by bulk
//                copy straight into its operand tile
//   warps 20-23  stage B: forward Gram blocks (64-bit conflict-free fragment loads), back substitution
//   warp   24    MMA issuer; also brings S0^T (the checkpoint, already in operand layout) in with one bulk copy per chunk
// Three operand slots; scan scratch, stage-B scratch and the output staging alias slot tiles that are written later /
// already consumed.  Bound by shared memory (bytes through the banks + the dependent hand-offs of a chunk: DESIGN.md 4).
#include "mma_tf32.cuh"
#include "tc05.cuh"
#include "tma_map.h"
#include "wkv7_common.cuh"

namespace rwkvtts {
namespace tcbwd {
using namespace tc05;

constexpr int L = 16, WIN = 4, NS = 3;
constexpr float kMinLogDecay = -1.35f;

// canonical K-major tiles (floats): off = (r/8)*SBO + (k/4)*LBO + (r%8)*4 + k%4
constexpr int N32_LBO = 132, N16_LBO = 68, N_SBO = 32;   // [32|16 rows (tokens)][64 channels]
constexpr int T_SBO = 36, T_LBO = 288;                   // [64 rows (channels)][16 tokens]
constexpr int G_LBO = 296;   // same tiles when stage B also reads them as mma.sync fragments (Q~, A~, B~, K~): = 8 mod 32, so the
                             // 64-bit fragment loads of a half-warp hit 32 different banks
// [32|16 rows][16].  The 16-row tiles (stage B's Gram / column stores) and the transposed gradient-Gram tiles are padded by one
// 16-byte piece per K block: with the dense strides (64 / 128 floats = 0 mod 32 banks) the fragment stores were 8-way / 4-way /
// 2-way bank conflicted (ncu: 64 wavefronts per tile and chunk where 8-16 are needed); 68 makes the Gram stores 2-way and the
// column stores conflict-free, 132 the transposed stores conflict-free (the natural ones stay 2-way for any legal stride)
constexpr int S32_LBO = 128, S32T_LBO = 132, S16_LBO = 68, S_SBO = 32;

struct Slot {
    float UVn[16 * N32_LBO];      // rows 0-15 U (sa), 16-31 V          [.][value]
    float DYZn[16 * N32_LBO];     // rows 0-15 dY, 16-31 Z (group C)     [.][value]
    float Kn[16 * N16_LBO];       // K~ [s][key]
    float Bpn[16 * N16_LBO];      // B' [s][key]                         (stage B)
    float dYt[4 * T_LBO], Qt[4 * G_LBO], At[4 * G_LBO], Bt[4 * G_LBO], Kt[4 * G_LBO];   // [channel][token]
    float Gt[4 * T_LBO];          // G (fp32, not an MMA operand), same layout
    float AqbpT[4 * S16_LBO], AqkT[4 * S16_LBO], AakT[4 * S16_LBO];   // [n=s][k=t]          (stage B)
    // (G at the chunk start lives in the padding of Gt: gs_off())
};
// G at the chunk start, one float per channel row, kept in the 16-byte gaps the T_SBO = 36 stride leaves after every core
// matrix of the Gt tile (outside what any tile access touches)
__host__ __device__ constexpr int gs_off(int row) { return (row >> 3) * T_SBO + ((row & 7) >> 1) * T_LBO + 32 + (row & 1); }
constexpr int NRAW = 2;
// the seven raw input tiles of one chunk as the TMA engine lands them: [16 tokens][64 channels] bf16, 128-byte rows;
// x[i][16 * t + k4] is the 8-byte piece (channels 4*k4 .. 4*k4+3 of token t) one stage-A thread expands
struct RawBuf { uint2 x[7][256]; };
struct Smem {
    Slot slot[NS];
    float ZT[4 * T_LBO];          // Z^T [value][t]
    float QB_N[4 * S32_LBO];      // rows 0-15 dAqb [t][s], 16-31 dN [t][s]
    float QK_AK[4 * S32_LBO];     // rows 0-15 dAqk, 16-31 dAak
    float NT_AKT[4 * S32T_LBO];   // rows 0-15 dN^T [s][t], 16-31 dAak^T
    float QBT_QKT[4 * S32T_LBO];  // rows 0-15 dAqb^T, 16-31 dAqk^T
    __align__(128) RawBuf raw[NRAW];   // stage A: landing buffers of the tensor-map copies, one chunk ahead
    __align__(128) float S0c[kCkFloats];   // checkpoint S0^T of the chunk in flight: K-major operand tile [key][value],
                                           // brought in by the MMA warp with one bulk copy per chunk
    // (the scan partials of stages A and B live in tiles of the slot they own that are written later)
    float gst[WIN][kC];           // stage A: G at the start of each chunk of the current window
    float elast[kC];              // group C: e^{G} at a window end
    float glp[2][kC], suff[kC], cfirst[kC], xch[2][kC];   // group C: boundary-term partials, dw suffix, carries
    uint64_t full[NS], empty[NS], a_done[NS], blob_full[NS], s0_full;
    uint64_t resc, glp_done, ok_free[2], bar_z, c_done;
    // one barrier per iteration parity: with a single barrier the MMA warp could commit iteration it+1 before group C2
    // had observed iteration it (nothing on C2's side gates that commit), the 1-bit phase parity would flip twice and
    // C2 would wait for a phase that can only complete after its own ok_free arrival -- a deadlock (found by the
    // round-1 protocol model and VERDICT; reproduced on the GPU by tests/test_stress_gpu.py).  out_ready[p] is committed at iterations of parity
    // p only, and the next commit on it (it+2) waits for ok_free[p] of iteration `it`, which C2 gives after its wait.
    uint64_t out_ready[2];
    uint64_t raw_full[NRAW];
    uint32_t tmem_base;
};

struct Params {
    int T, H;
    const bf16 *w, *q, *k, *v, *a, *b, *dy;
    const float *ckT, *sa;
    const float *sT, *dsT;       // final state / its gradient (both or neither)
    bf16 *dw, *dq, *dk, *dv, *da, *db;
    float *ds0;                  // may be null
    long long *dbg;              // phase-cycle counters (profiling builds only), may be null
    const int *cu, *cbase;       // packed launch (see SeqWork, wkv7_common.cuh), else null
};

#ifdef RWKVTTS_PROFILE
#define TICK(var) long long var = clock64()
#define ACC(slot, t0, t1) do { if (P_dbg && blockIdx.x == 0) P_dbg[slot] += (t1) - (t0); } while (0)
#else
#define TICK(var)
#define ACC(slot, t0, t1)
#endif

// tensor-memory columns
constexpr uint32_t C_DS = 0, C_DST = 64, C_Z = 128, C_OK = 144, C_OV = 208, C_OBUF = 80;   // OK/OV double buffered

__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_nc_f4(const void *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void unpack4(const uint2 &u, float *f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
}

// ---------------------------------------------------------------------------------------------
// stage A: tp in [0,256); thread = (token t, channels 4*k4 .. 4*k4+3).  Warp wp holds the token group 4*(wp>>1) .. +3:
// lane = 8*tt + k4l is token 4*(wp>>1) + tt and channel quad k4 = k4l + 8*((tt ^ wp) & 1) -- the two warps of a token group
// take opposite halves of the channels on odd / even tokens.  With it (bank model + ncu) every access of the stage is
// conflict-free: a half-warp reads 16 different channel quads of the raw tiles, a scalar store into a transposed
// [channel][token] operand tile hits 8 row groups x 4 tokens = 32 banks (two tokens x 16 quads per warp, the first mapping,
// made all six of them 2-way conflicted: 6 % of the kernel's wavefronts), the 16-byte stores into the [token][channel]
// tiles stay 8 consecutive quads per quarter-warp.
// ---------------------------------------------------------------------------------------------
// raw inputs travel HBM -> shared memory as tensor-map boxes (cp.async.bulk.tensor: one 16-token x 64-channel tile of
// this head per tensor and chunk, no register and no LSU instruction on the way), one chunk ahead; one elected thread
// issues, the mbarrier counts the bytes
__device__ __forceinline__ void issue_raw(Smem &sm, const TmaMaps &M, int h, int tok0, int it, int c) {
    RawBuf &rb = sm.raw[it % NRAW];
    uint64_t *bar = &sm.raw_full[it % NRAW];
    mbar_expect_tx(bar, (uint32_t)sizeof(RawBuf));
#pragma unroll
    for (int i = 0; i < 7; i++) tma_load_box3(&rb.x[i][0], &M.m[i], 0, h, tok0 + c * L, bar);
}

template <bool kVar>
__device__ void stage_a(const Params &P, const TmaMaps &M, Smem &sm, size_t base, size_t tok_stride, size_t ck0, int h, int tok0,
                        int nC, int len, int tp) {
    long long *P_dbg = tp == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int wp = tp >> 5, lane = tp & 31, tt = lane >> 3;
    const int t = 4 * (wp >> 1) + tt, k4 = (lane & 7) + 8 * ((tt ^ wp) & 1);
    // w of every chunk of the window whose last chunk is c_last: needed when a window is entered from its end,
    // fetched one window ahead
    auto loadw = [&](int c_last, uint2 (&wr)[WIN]) {
        const int c0 = (c_last / WIN) * WIN;
#pragma unroll
        for (int j = 0; j < WIN; j++)
            if (c0 + j <= c_last)
                wr[j] = (!kVar || (c0 + j) * L + t < len) ? ldg_nc_v2(P.w + base + (size_t)((c0 + j) * L + t) * tok_stride + k4 * 4)
                                                : make_uint2(0u, 0u);
    };
    uint2 wr[WIN];
    if (tp == 0) issue_raw(sm, M, h, tok0, 0, nC - 1);                     // prologue: chunk of iteration 0
    loadw(nC - 1, wr);
    for (int it = 0; it < nC; it++) {
        const int c = nC - 1 - it, si = it % NS;
        Slot &S = sm.slot[si];
        TICK(ta0);
        // one chunk ahead: every stage-A thread took its pieces of that buffer (iteration it - 1) before the scan
        // barrier of iteration it - 1, which this thread has passed
        if (tp == 0 && it + 1 < nC) issue_raw(sm, M, h, tok0, it + 1, c - 1);
        const bool win_last = (c % WIN == WIN - 1) || (c == nC - 1);
        TICK(ta1);
        if (it >= NS) mbar_wait(&sm.empty[si], ((it / NS) - 1) & 1);
        TICK(ta2);
        float(&lwt)[L][kC] = *reinterpret_cast<float(*)[L][kC]>(S.Bpn);         // scratch: stage B writes the tile later
        static_assert(sizeof(float) * L * kC <= sizeof(S.Bpn), "scan scratch fits in the B' tile");
        if (win_last) {
            // entering a window (from its end): G at the start of each of its chunks
            float(&wtw)[WIN][8][kC] = *reinterpret_cast<float(*)[WIN][8][kC]>(S.DYZn);   // scratch: dY rows come below
            const int c0 = (c / WIN) * WIN, nj = c - c0 + 1;
#pragma unroll
            for (int j = 0; j < WIN; j++) {
                if (j >= nj) break;
                float f[4], s4[4];
                unpack4(wr[j], f);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    s4[e] = fmaxf(-__expf(f[e]), kMinLogDecay);
                    s4[e] += __shfl_xor_sync(0xffffffffu, s4[e], 16);      // lanes tt and tt ^ 2: same channel quad
                }
                if (lane & 16) st4(&wtw[j][wp][k4 * 4], s4[0], s4[1], s4[2], s4[3]);   // this warp's two tokens of the quad
            }
            if (c0 > 0) loadw(c0 - 1, wr);        // the window below
            bar_sync(1, 256);
            if (tp < kC) {
                float run = 0.f;
                for (int j = 0; j < nj; j++) {
                    sm.gst[j][tp] = run;
#pragma unroll
                    for (int ww = 0; ww < 8; ww++) run += wtw[j][ww][tp];
                }
            }
            bar_sync(1, 256);
        }
        mbar_wait(&sm.raw_full[it % NRAW], (it / NRAW) & 1);        // this iteration's raw tiles have landed
        const RawBuf &rb = sm.raw[it % NRAW];
        struct { uint2 x[7]; } raw;
        {
            // beyond the end of a packed sequence (the box holds the next sequence's tokens, or zeros past the tensor's
            // end): a token that changes nothing and is never stored
            const bool live = !kVar || c * L + t < len;
#pragma unroll
            for (int i = 0; i < 7; i++) raw.x[i] = live ? rb.x[i][16 * t + k4] : make_uint2(0u, 0u);
        }
        float lw[4], gg[4];
        {
            float f[4];
            unpack4(raw.x[0], f);
#pragma unroll
            for (int j = 0; j < 4; j++) lw[j] = fmaxf(-__expf(f[j]), kMinLogDecay);
        }
        // G_t = G at the chunk start + inclusive sum of the log-decays over the chunk's tokens: every thread parks its four
        // values, 64 threads (one per channel) run the prefix over the 16 tokens (independent loads, sums in registers) and
        // put G back, every thread takes its own
        st4(&lwt[t][k4 * 4], lw[0], lw[1], lw[2], lw[3]);
        bar_sync(1, 256);
        if (tp < kC) {
            float v[L];
#pragma unroll
            for (int i = 0; i < L; i++) v[i] = lwt[i][tp];
            float run = sm.gst[c % WIN][tp];
#pragma unroll
            for (int i = 0; i < L; i++) { run += v[i]; lwt[i][tp] = run; }
        }
        bar_sync(1, 256);
        {
            const float4 g0 = *reinterpret_cast<const float4 *>(&lwt[t][k4 * 4]);
            gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w;
        }
        if (wp == 0) {   // U of this chunk: operand tile rows, HBM -> slot, bulk copies
            const int ln = tp & 31;
            if (ln == 0) mbar_expect_tx(&sm.blob_full[si], kUFloats * 4);
            __syncwarp();
            if (ln < 16)
                bulk_g2s(&S.UVn[ln * N32_LBO], P.sa + (ck0 + c) * kUFloats + ln * kULbo, kULbo * 4,
                         &sm.blob_full[si]);
        }
        {
            float E[4], Ep[4], iE[4], f[4], o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float g2 = gg[j] * 1.4426950408889634f;
                E[j] = ex2f(g2);
                Ep[j] = ex2f(g2 - lw[j] * 1.4426950408889634f);
                iE[j] = ex2f(-g2);
            }
            // transposed tiles: row = channel 4*k4+j, column = token t
            const int ot = (k4 >> 1) * T_SBO + (t >> 2) * T_LBO + (k4 & 1) * 16 + (t & 3);   // + 4*j
            const int og = (k4 >> 1) * T_SBO + (t >> 2) * G_LBO + (k4 & 1) * 16 + (t & 3);   // Q~, A~, B~, K~ tiles
            const int on32 = (t >> 3) * N_SBO + k4 * N32_LBO + (t & 7) * 4;                // 32-row tiles, row t
            const int on16 = (t >> 3) * N_SBO + k4 * N16_LBO + (t & 7) * 4;
            unpack4(raw.x[1], f);                                   // Q~
#pragma unroll
            for (int j = 0; j < 4; j++) S.Qt[og + 4 * j] = tf32r(f[j] * E[j]);
            unpack4(raw.x[2], f);                                   // K~
#pragma unroll
            for (int j = 0; j < 4; j++) { o[j] = tf32r(f[j] * iE[j]); S.Kt[og + 4 * j] = o[j]; }
            st4(&S.Kn[on16], o[0], o[1], o[2], o[3]);
            unpack4(raw.x[3], f);                                   // V (exact in tf32)
            st4(&S.UVn[on32 + 2 * N_SBO], f[0], f[1], f[2], f[3]);
            unpack4(raw.x[4], f);                                   // A~
#pragma unroll
            for (int j = 0; j < 4; j++) S.At[og + 4 * j] = tf32r(f[j] * Ep[j]);
            unpack4(raw.x[5], f);                                   // B~
#pragma unroll
            for (int j = 0; j < 4; j++) S.Bt[og + 4 * j] = tf32r(f[j] * iE[j]);
            unpack4(raw.x[6], f);                                   // dY (exact in tf32)
#pragma unroll
            for (int j = 0; j < 4; j++) S.dYt[ot + 4 * j] = f[j];
            st4(&S.DYZn[on32], f[0], f[1], f[2], f[3]);
#pragma unroll
            for (int j = 0; j < 4; j++) S.Gt[ot + 4 * j] = gg[j];
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < 4; j++) S.Gt[gs_off(4 * k4 + j)] = gg[j] - lw[j];
            }
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.a_done[si]);
        mbar_arrive_warp(&sm.full[si]);
        TICK(ta3); ACC(0, ta0, ta1); ACC(1, ta1, ta2); ACC(2, ta2, ta3);
    }
}

// ---------------------------------------------------------------------------------------------
// stage B: tp in [0,128), group grp handles iterations it = grp, grp+2, ...
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int perm8(int g) { return (g & 1) | ((g & 2) << 1) | ((g & 4) >> 1); }

__device__ void stage_b(const Params &P, Smem &sm, int nC, int tp) {
    constexpr int grp = 0;
    long long *P_dbg = (tp == 0 && grp == 0) ? P.dbg : nullptr; (void)P_dbg;
    const int wp = tp >> 5, lane = tp & 31, g = lane >> 2, tq = lane & 3;
    for (int it = 0; it < nC; it++) {
        const int si = it % NS;
        Slot &S = sm.slot[si];
        // scratch N and Aqb, natural [t][s] fp32, row stride N32_LBO: the Z-row pieces of DYZn (group C1 fills them later)
        float *Nn = S.DYZn + 64, *AQ = S.DYZn + 96;
        TICK(tb0);
        mbar_wait(&sm.a_done[si], (it / NS) & 1);
        TICK(tb1);
        {   // forward Gram blocks from the transposed tiles: (A~|Q~)(B~|K~)^T, one 16x16 block per warp.  Fragment rows
            // (g, g+8) are tokens (2p, 2p+1) and column g of n-tile nt is token 2p+nt, p = perm8(g): every fragment
            // register pair is one 64-bit load (two adjacent tokens of one channel), conflict-free with G_LBO = 8 mod 32
            const int rowsel = wp & 1, colsel = wp >> 1;
            const int pg = perm8(g);
            const int ofrag = (pg >> 1) * G_LBO + 2 * (pg & 1) + tq * 4;
            const float *Ar = (rowsel ? S.Qt : S.At) + ofrag;
            const float *Bc = (colsel ? S.Kt : S.Bt) + ofrag;
            float acc[2][4] = {};
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                const float2 a01 = *reinterpret_cast<const float2 *>(Ar + kb * T_SBO);
                const float2 a23 = *reinterpret_cast<const float2 *>(Ar + kb * T_SBO + 16);
                const float2 b0 = *reinterpret_cast<const float2 *>(Bc + kb * T_SBO);        // k = tq:   n-tiles 0, 1
                const float2 b1 = *reinterpret_cast<const float2 *>(Bc + kb * T_SBO + 16);   // k = tq+4
                const uint32_t af[4] = {__float_as_uint(a01.x), __float_as_uint(a01.y), __float_as_uint(a23.x),
                                        __float_as_uint(a23.y)};
                const uint32_t bf0[2] = {__float_as_uint(b0.x), __float_as_uint(b1.x)};
                const uint32_t bf1[2] = {__float_as_uint(b0.y), __float_as_uint(b1.y)};
                mma_tf32(acc[0], af, bf0);
                mma_tf32(acc[1], af, bf1);
            }
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = 2 * perm8(2 * tq + e) + nt;   // s
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int row = 2 * pg + hh;              // t
                        float x = acc[nt][2 * hh + e];
                        if (rowsel) {
                            x = (col <= row) ? x : 0.f;
                            if (colsel) S.AqkT[kmajor_off(col, row, S16_LBO, S_SBO)] = tf32r(x);
                            else AQ[row * N32_LBO + col] = x;
                        } else {
                            x = (col < row) ? x : 0.f;
                            if (colsel) S.AakT[kmajor_off(col, row, S16_LBO, S_SBO)] = tf32r(x);
                            else Nn[row * N32_LBO + col] = x;
                        }
                    }
                }
        }
        bar_sync(2 + grp, 128);
        // [B' | Aqb'^T] = (I - N)^-T [B~ | Aqb^T]: back substitution, one column per thread
        if (tp < kC + 16) {
            float acc[L];
            if (tp < kC) {
                const float *pr = S.Bt + (tp >> 3) * T_SBO + (tp & 7) * 4;
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {
                    const float4 x = *reinterpret_cast<const float4 *>(pr + q4 * G_LBO);
                    acc[4 * q4] = x.x; acc[4 * q4 + 1] = x.y; acc[4 * q4 + 2] = x.z; acc[4 * q4 + 3] = x.w;
                }
            } else {
#pragma unroll
                for (int s = 0; s < L; s++) acc[s] = AQ[(tp - kC) * N32_LBO + s];
            }
#pragma unroll
            for (int tt = L - 1; tt >= 1; tt--) {
                const float x = acc[tt];
#pragma unroll
                for (int q4 = 0; q4 <= (tt - 1) / 4; q4++) {
                    const float4 n4 = *reinterpret_cast<const float4 *>(&Nn[tt * N32_LBO + 4 * q4]);
                    const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (4 * q4 + e < tt) acc[4 * q4 + e] = fmaf(nn[e], x, acc[4 * q4 + e]);
                }
            }
            if (tp < kC) {
#pragma unroll
                for (int s = 0; s < L; s++) S.Bpn[kmajor_off(s, tp, N16_LBO, N_SBO)] = tf32r(acc[s]);
            } else {
#pragma unroll
                for (int s = 0; s < L; s++) S.AqbpT[kmajor_off(s, tp - kC, S16_LBO, S_SBO)] = tf32r(acc[s]);
            }
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.full[si]);
        bar_sync(2 + grp, 128);
        TICK(tb2); ACC(3, tb0, tb1); ACC(4, tb1, tb2);
    }
}

// ---------------------------------------------------------------------------------------------
// MMA issuer
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t kadv(uint64_t d, int kk, int lbo_f) { return d + (uint64_t)((kk * 2 * lbo_f * 4) >> 4); }

__device__ void mma_warp(const Params &P, Smem &sm, size_t ck0, int nC) {
    long long *P_dbg = (threadIdx.x & 31) == 0 ? P.dbg : nullptr; (void)P_dbg;
    const uint32_t tb = sm.tmem_base;
    constexpr uint32_t I16 = idesc_tf32(64, 16, false, false);
    constexpr uint32_t I32 = idesc_tf32(64, 32, false, false);
    constexpr uint32_t I64 = idesc_tf32(64, 64, false, false);
    const uint64_t dZT = smem_desc(smem_u32(sm.ZT), T_LBO * 4, T_SBO * 4);
    const uint64_t dQBN = smem_desc(smem_u32(sm.QB_N), S32_LBO * 4, S_SBO * 4);
    const uint64_t dQKAK = smem_desc(smem_u32(sm.QK_AK), S32_LBO * 4, S_SBO * 4);
    const uint64_t dNTAKT = smem_desc(smem_u32(sm.NT_AKT), S32T_LBO * 4, S_SBO * 4);
    const uint64_t dQBTQKT = smem_desc(smem_u32(sm.QBT_QKT), S32T_LBO * 4, S_SBO * 4);
    int nw = 0;
    for (int it = 0; it < nC; it++) {
        const int c = nC - 1 - it, si = it % NS;
        const bool win_last = (c % WIN == WIN - 1) || (c == nC - 1);
        const uint32_t ok = tb + C_OK + C_OBUF * (it & 1), ov = tb + C_OV + C_OBUF * (it & 1);
        const Slot &S = sm.slot[si];
        const uint64_t dS0 = smem_desc(smem_u32(sm.S0c), kCkLbo * 4, 32 * 4);
        const uint64_t dUV = smem_desc(smem_u32(S.UVn), N32_LBO * 4, N_SBO * 4);
        const uint64_t dDYZ = smem_desc(smem_u32(S.DYZn), N32_LBO * 4, N_SBO * 4);
        const uint64_t dKn = smem_desc(smem_u32(S.Kn), N16_LBO * 4, N_SBO * 4);
        const uint64_t dBp = smem_desc(smem_u32(S.Bpn), N16_LBO * 4, N_SBO * 4);
        const uint64_t dYt = smem_desc(smem_u32(S.dYt), T_LBO * 4, T_SBO * 4);
        const uint64_t dQt = smem_desc(smem_u32(S.Qt), G_LBO * 4, T_SBO * 4);
        const uint64_t dAt = smem_desc(smem_u32(S.At), G_LBO * 4, T_SBO * 4);
        const uint64_t dBt = smem_desc(smem_u32(S.Bt), G_LBO * 4, T_SBO * 4);
        const uint64_t dKt = smem_desc(smem_u32(S.Kt), G_LBO * 4, T_SBO * 4);
        const uint64_t dAqbpT = smem_desc(smem_u32(S.AqbpT), S16_LBO * 4, S_SBO * 4);
        const uint64_t dAqkT = smem_desc(smem_u32(S.AqkT), S16_LBO * 4, S_SBO * 4);
        const uint64_t dAakT = smem_desc(smem_u32(S.AakT), S16_LBO * 4, S_SBO * 4);
        TICK(tm0);
        mbar_wait(&sm.full[si], (it / NS) & 1);
        mbar_wait(&sm.blob_full[si], (it / NS) & 1);     // U and S0^T operand tiles (bulk copies)
        TICK(tm1);
        // every MMA of the previous chunk has completed (dS, dS^T are final; Z, the Gram tiles and the slot two
        // iterations back are free); at a window boundary group C1 has moved dS / dS^T into this window's frame;
        // group C2 has drained the gradient accumulators of two iterations back
        #ifdef RWKVTTS_BWD_SINGLE_OUT_READY      // round-1 protocol, kept only so that scripts/stress_wkv7.py can show the hazard
        if (it > 0) mbar_wait(&sm.out_ready[0], (it - 1) & 1);
#else
        if (it > 0) mbar_wait(&sm.out_ready[(it - 1) & 1], ((it - 1) >> 1) & 1);
#endif
        if (win_last) { mbar_wait(&sm.resc, nw & 1); nw++; }
        if (it >= 2) mbar_wait(&sm.ok_free[it & 1], ((it >> 1) - 1) & 1);
        TICK(tm2);
        fence_after_sync();
        if (elect_one()) {
            // S0^T of this chunk (the previous chunk's MMAs and, at a window boundary, group C2 are done with the tile)
            mbar_expect_tx(&sm.s0_full, kCkFloats * 4);
            bulk_g2s(sm.S0c, P.ckT + (ck0 + c) * kCkFloats, kCkFloats * 4, &sm.s0_full);
            // R1: Z^T = dY^T Aqb' + dS B'^T
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb + C_Z, kadv(dYt, kk, T_LBO), kadv(dAqbpT, kk, S16_LBO), I16, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ts(tb + C_Z, tb + C_DS + 8 * kk, kadv(dBp, kk, N16_LBO), I16, true);
            mma_commit(&sm.bar_z);
            // P3a: dV^T = dS K~^T ;  P2a: [dB~^T | dK~^T] = dS^T [U;V]^T     (dS, dS^T before this chunk's update)
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ts(ov, tb + C_DS + 8 * kk, kadv(dKn, kk, N16_LBO), I16, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ts(ok + 32, tb + C_DST + 8 * kk, kadv(dUV, kk, N32_LBO), I32, kk > 0);
        }
        __syncwarp();
        mbar_wait(&sm.bar_z, it & 1);
        TICK(tm3);
        fence_after_sync();
        if (elect_one()) {
            // R2 (dS): dS += dY^T Q~ + Z^T A~
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb + C_DS, kadv(dYt, kk, T_LBO), kadv(dQt, kk, G_LBO), I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(tb + C_DS, tb + C_Z + 8 * kk, kadv(dAt, kk, G_LBO), I64, true);
        }
        __syncwarp();
        mbar_wait(&sm.c_done, it & 1);            // group C1: Z tiles and the gradient Gram tiles
        mbar_wait(&sm.s0_full, it & 1);           // S0^T operand tile
        TICK(tm4);
        fence_after_sync();
        if (elect_one()) {
            // R2 (dS^T): dS^T += Q~^T dY + A~^T Z
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb + C_DST, kadv(dQt, kk, G_LBO), kadv(dYt, kk, T_LBO), I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(tb + C_DST, kadv(dAt, kk, G_LBO), kadv(dZT, kk, T_LBO), I64, true);
            // P1: [dQ~^T | dA~^T] = S0^T [dY;Z]^T + B~^T [dAqb;dN]^T + K~^T [dAqk;dAak]^T
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                mma_tf32_ss(ok, kadv(dS0, kk, kCkLbo), kadv(dDYZ, kk, N32_LBO), I32, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(ok, kadv(dBt, kk, G_LBO), kadv(dQBN, kk, S32_LBO), I32, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(ok, kadv(dKt, kk, G_LBO), kadv(dQKAK, kk, S32_LBO), I32, true);
            // P2b: [dB~^T | dK~^T] += A~^T [dN^T;dAak^T]^T + Q~^T [dAqb^T;dAqk^T]^T
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(ok + 32, kadv(dAt, kk, G_LBO), kadv(dNTAKT, kk, S32T_LBO), I32, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(ok + 32, kadv(dQt, kk, G_LBO), kadv(dQBTQKT, kk, S32T_LBO), I32, true);
            // P3b: dV^T += dY^T Aqk + Z^T Aak
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ss(ov, kadv(dYt, kk, T_LBO), kadv(dAqkT, kk, S16_LBO), I16, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_tf32_ts(ov, tb + C_Z + 8 * kk, kadv(dAakT, kk, S16_LBO), I16, true);
#ifdef RWKVTTS_BWD_SINGLE_OUT_READY
            mma_commit(&sm.out_ready[0]);
#else
            mma_commit(&sm.out_ready[it & 1]);
#endif
        }
        __syncwarp();
        TICK(tm5); ACC(5, tm0, tm1); ACC(6, tm1, tm2); ACC(7, tm2, tm3); ACC(8, tm3, tm4); ACC(9, tm4, tm5);
    }
}

// ---------------------------------------------------------------------------------------------
// group C1: 4 warps, on the chain between the MMA batches of one chunk.  Warp q owns tensor-memory lanes
// 32q..32q+15 = rows 16q..16q+15.  Per chunk: (window boundary) move dS / dS^T into the window's frame; Z^T ->
// shared operand tiles; the four 16x16 gradient Gram blocks (mma.sync) -> operand tiles.
// ---------------------------------------------------------------------------------------------
__device__ void group_c1(const Params &P, Smem &sm, int bh, int nC, int tid) {
    long long *P_dbg = tid == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int q = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16);
    const int trow = (row >> 3) * T_SBO + (row & 7) * 4;
    {   // dS, dS^T <- dsT (or 0)
        const bool have = P.dsT != nullptr && P.sT != nullptr;
        const float *gs = have ? P.dsT + (size_t)bh * kC * kC : nullptr;
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {
            float v[16], vt[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                v[i] = have ? gs[row * kC + 16 * cb + i] : 0.f;              // dS[row][.]
                vt[i] = have ? gs[(16 * cb + i) * kC + row] : 0.f;           // dS^T[row][.] = dS[.][row]
            }
            tmem_st16(tb + C_DS + 16 * cb, v);
            tmem_st16(tb + C_DST + 16 * cb, vt);
        }
        tmem_wait_st();
    }
    int nw = 0;
    for (int it = 0; it < nC; it++) {
        const int c = nC - 1 - it, si = it % NS;
        Slot &S = sm.slot[si];
        const bool win_last = (c % WIN == WIN - 1) || (c == nC - 1);
        TICK(tc0);
        if (win_last) {
            // dS, dS^T arrive in the frame of the NEXT window (or true frame at the sequence end): move them into
            // this window's frame, columns (keys) scaled by e^{G_last}.  Group C2 has taken the boundary term from
            // dS^T first (which also means every MMA of the previous chunk has completed).
            mbar_wait(&sm.full[si], (it / NS) & 1);
            if (it > 0) mbar_wait(&sm.glp_done, (nw - 1) & 1);
            fence_after_sync();
            if (act) sm.elast[row] = __expf(S.Gt[trow + 3 * T_LBO + 3]);
            bar_sync(4, 128);
            const float er = sm.elast[row];
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                float v[16];
                tmem_ld16(tb + C_DS + 16 * cb, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] *= sm.elast[16 * cb + i];
                tmem_st16(tb + C_DS + 16 * cb, v);
                tmem_ld16(tb + C_DST + 16 * cb, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] *= er;
                tmem_st16(tb + C_DST + 16 * cb, v);
            }
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive_warp(&sm.resc);
            nw++;
        }
        TICK(tc2);
        // ---- Z^T -> shared tiles ----------------------------------------------------------------
        mbar_wait(&sm.bar_z, it & 1);
        TICK(tc3);
        fence_after_sync();
        {
            float z[16];
            tmem_ld16(tb + C_Z, z);
            tmem_wait_ld();
            if (act) {
#pragma unroll
                for (int i = 0; i < 16; i++) z[i] = tf32r(z[i]);
#pragma unroll
                for (int i = 0; i < 16; i++) S.DYZn[kmajor_off(16 + i, row, N32_LBO, N_SBO)] = z[i];
                float *pz = sm.ZT + trow;
#pragma unroll
                for (int i = 0; i < 4; i++) st4(pz + i * T_LBO, z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
            }
        }
        bar_sync(4, 128);
        mbar_wait(&sm.blob_full[si], (it / NS) & 1);     // U (and S0^T) of this chunk have landed
        {   // gradient Gram blocks over the value index, one 16x16 block per warp: q = 0 dAqb=tril(dY U^T),
            // 1 dN=stril(Z U^T), 2 dAqk=tril(dY V^T), 3 dAak=stril(Z V^T)
            const int ar = (q & 1) * 16, br = (q >> 1) * 16;
            float acc[2][4] = {};
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                uint32_t af[4], bfr[2];
                const float *pa = S.DYZn + (ar >> 3) * N_SBO + 2 * kb * N32_LBO + g * 4 + tq;
                af[0] = __float_as_uint(pa[0]); af[1] = __float_as_uint(pa[N_SBO]);
                af[2] = __float_as_uint(pa[N32_LBO]); af[3] = __float_as_uint(pa[N32_LBO + N_SBO]);
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    const float *pb = S.UVn + ((br >> 3) + nt) * N_SBO + 2 * kb * N32_LBO + g * 4 + tq;
                    bfr[0] = __float_as_uint(pb[0]); bfr[1] = __float_as_uint(pb[N32_LBO]);
                    mma_tf32(acc[nt], af, bfr);
                }
            }
            float *nat = (q < 2) ? sm.QB_N : sm.QK_AK;                  // [n=t][k=s], rows +16 for the Z grams
            float *trn = (q == 0 || q == 2) ? sm.QBT_QKT : sm.NT_AKT;   // [n=s][k=t]
            const int nrow = (q & 1) * 16, trow_ = (q >> 1) * 16;
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = 8 * nt + 2 * tq + e;      // s
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int r = g + 8 * hh;             // t
                        float x = acc[nt][2 * hh + e];
                        x = ((q & 1) ? (col < r) : (col <= r)) ? tf32r(x) : 0.f;
                        nat[kmajor_off(nrow + r, col, S32_LBO, S_SBO)] = x;
                        trn[kmajor_off(trow_ + col, r, S32T_LBO, S_SBO)] = x;
                    }
                }
        }
        fence_proxy_async();
        mbar_arrive_warp(&sm.c_done);
        TICK(tc4); ACC(10, tc0, tc2); ACC(12, tc2, tc3); ACC(13, tc3, tc4);
    }
}

// ---------------------------------------------------------------------------------------------
// group C2: 8 warps, off the chain: the output epilogue of chunk `it` runs while the MMA warp and C1 work on
// chunk it+1 (the gradient accumulators are double buffered in tensor memory).  Warp (q, hf): q = warp & 3 owns
// lanes 32q..32q+15 = channel rows 16q..16q+15, hf = warp >> 2 takes tokens 8hf..8hf+7.
// ---------------------------------------------------------------------------------------------
template <bool kVar>
__device__ void group_c2(const Params &P, Smem &sm, size_t base, size_t tok_stride, int bh, int nC, int len, int tid) {
    long long *P_dbg = tid == 0 ? P.dbg : nullptr; (void)P_dbg;
    const int wq = tid >> 5, q = wq & 3, hf = wq >> 2, lane = tid & 31;
    const bool act = lane < 16;
    const int row = 16 * q + (lane & 15);
    const uint32_t tb = sm.tmem_base + ((uint32_t)(32 * q) << 16);
    const int trow = (row >> 3) * T_SBO + (row & 7) * 4;
    {   // boundary term of the last window from sT . dsT
        const bool have = P.dsT != nullptr && P.sT != nullptr;
        const float *gs = have ? P.dsT + (size_t)bh * kC * kC : nullptr;
        const float *st = have ? P.sT + (size_t)bh * kC * kC : nullptr;
        float glp = 0.f;
        if (have) {
#pragma unroll 4
            for (int i = 0; i < 32; i++) glp = fmaf(gs[(32 * hf + i) * kC + row], st[(32 * hf + i) * kC + row], glp);
        }
        if (act) { sm.glp[hf][row] = glp; if (hf == 0) { sm.suff[row] = 0.f; sm.cfirst[row] = 0.f; } }
        bar_sync(5, 256);
    }
    for (int it = 0; it < nC; it++) {
        const int c = nC - 1 - it, si = it % NS;
        Slot &S = sm.slot[si];
        const bool win_last = (c % WIN == WIN - 1) || (c == nC - 1);
        const uint32_t ok = tb + C_OK + C_OBUF * (it & 1), ov = tb + C_OV + C_OBUF * (it & 1);
        bf16(*obuf)[L][72] = reinterpret_cast<bf16(*)[L][72]>(S.UVn);      // 6 x 16 x 72 bf16 over UVn + DYZn
        static_assert(6 * L * 72 * 2 <= 2 * 16 * N32_LBO * 4, "output staging fits in UVn + DYZn");
        if (win_last && it > 0) {
            if (act && hf == 0) { sm.suff[row] = 0.f; sm.cfirst[row] = 0.f; }
            bar_sync(5, 256);
        }
        TICK(tc4);
        // ---- outputs: this warp's 8 tokens (hf = 1 is later in time and feeds hf = 0) -------------
#ifdef RWKVTTS_BWD_SINGLE_OUT_READY
#ifdef RWKVTTS_BWD_DELAY_C2              // hold C2 back for about a chunk every so often: the window the hazard needs
        if ((it & 63) == 17 && (blockIdx.x & 7) == 3) { const long long t_ = clock64(); while (clock64() - t_ < 20000) {} }
#endif
        mbar_wait(&sm.out_ready[0], it & 1);
#else
#ifdef RWKVTTS_BWD_DELAY_C2
        if ((it & 63) == 17 && (blockIdx.x & 7) == 3) { const long long t_ = clock64(); while (clock64() - t_ < 20000) {} }
#endif
        mbar_wait(&sm.out_ready[it & 1], (it >> 1) & 1);
#endif
        TICK(tc5);
        fence_after_sync();
        {
            auto tile8 = [&](const float *T_, float (&o)[8], int lbo = G_LBO) {
                const float4 x0 = *reinterpret_cast<const float4 *>(T_ + trow + (2 * hf) * lbo);
                const float4 x1 = *reinterpret_cast<const float4 *>(T_ + trow + (2 * hf + 1) * lbo);
                o[0] = x0.x; o[1] = x0.y; o[2] = x0.z; o[3] = x0.w; o[4] = x1.x; o[5] = x1.y; o[6] = x1.z; o[7] = x1.w;
            };
            float G[8], lwv[8], gsum[8], acc_[8], op[8];
            tile8(S.Gt, G, T_LBO);
            {
                const float gprev = (hf == 0) ? S.Gt[gs_off(row)] : S.Gt[trow + T_LBO + 3];   // G of token 7
                lwv[0] = G[0] - gprev;
#pragma unroll
                for (int i = 1; i < 8; i++) lwv[i] = G[i] - G[i - 1];
            }
            const float suffix_old = sm.suff[row];
            // gradients are staged [array][token][channel] in tiles of this slot that are dead once the chunk's MMAs
            // have completed (UVn + DYZn), then leave as 128-byte rows (2-byte global stores straight from the
            // registers were measured 1.6x slower for the whole stage)
            // (idle lanes used to write into the padding columns: for the warps of rows 0-15 those share banks with the
            // live columns -- 20 % more wavefronts on the stage's hottest line, ncu; the stores are predicated instead)
            auto put = [&](int arr, int i, float x) { if (act) obuf[arr][8 * hf + i][row] = __float2bfloat16_rn(x); };
            float E[8], iE[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float g2 = G[i] * 1.4426950408889634f;
                E[i] = ex2f(g2);
                iE[i] = ex2f(-g2);
            }
            // dQ~ -> dq
            tmem_ld8(ok + 8 * hf, acc_); tmem_wait_ld();
            tile8(S.Qt, op);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                gsum[i] = acc_[i] * op[i];
                put(1, i, acc_[i] * E[i]);
            }
            // dK~ -> dk
            tmem_ld8(ok + 48 + 8 * hf, acc_); tmem_wait_ld();
            tile8(S.Kt, op);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                gsum[i] = fmaf(-acc_[i], op[i], gsum[i]);
                put(2, i, acc_[i] * iE[i]);
            }
            // dB~ -> db
            tmem_ld8(ok + 32 + 8 * hf, acc_); tmem_wait_ld();
            tile8(S.Bt, op);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                gsum[i] = fmaf(-acc_[i], op[i], gsum[i]);
                put(5, i, acc_[i] * iE[i]);
            }
            // dA~ -> da (e^{G_{t-1}} is the previous token's e^{G}); (dA~.A~)_{t+1} joins g_t
            tmem_ld8(ok + 16 + 8 * hf, acc_); tmem_wait_ld();
            tile8(S.At, op);
            {
                const float gprev2 = (G[0] - lwv[0]) * 1.4426950408889634f;
                float ep = ex2f(gprev2);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i > 0) gsum[i - 1] = fmaf(acc_[i], op[i], gsum[i - 1]);
                    put(4, i, acc_[i] * ep);
                    ep = E[i];
                }
            }
            const float aa0 = acc_[0] * op[0];
            {   // dV, this warp's 8 tokens
                float dvv[8];
                tmem_ld8(ov + 8 * hf, dvv); tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 8; i++) put(3, i, dvv[i]);
            }
            if (hf == 1) {
                // token 15 also gets (dA~.A~) of the next chunk's first token and, at a window end, sum_v dS.S
                gsum[7] += sm.cfirst[row] + (win_last ? sm.glp[0][row] + sm.glp[1][row] : 0.f);
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) tot += gsum[i];
                if (act) { sm.xch[0][row] = aa0; sm.xch[1][row] = tot; }
            }
            bar_sync(5, 256);
            float suffix = suffix_old;
            if (hf == 0) {
                gsum[7] += sm.xch[0][row];
                suffix += sm.xch[1][row];
            }
#pragma unroll
            for (int i = 7; i >= 0; i--) {
                suffix += gsum[i];
                put(0, i, suffix * lwv[i]);
            }
            if (hf == 0 && act) { sm.suff[row] = suffix; sm.cfirst[row] = aa0; }
        }
        // window boundary below this chunk: boundary term for the previous window's last token, taken from dS^T
        // (after this chunk's update) before group C1 moves it into the next window's frame
        if (c % WIN == 0 && c > 0) {
            float glp = 0.f;
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int cb = 2 * hf + cc;
                float a_[16];
                tmem_ld16(tb + C_DST + 16 * cb, a_);
                tmem_wait_ld();
                const float *sp = sm.S0c + (row >> 3) * 32 + (row & 7) * 4;      // S0^T[key = row][value]
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 x = *reinterpret_cast<const float4 *>(sp + (4 * cb + i) * kCkLbo);
                    glp = fmaf(a_[4 * i], x.x, glp); glp = fmaf(a_[4 * i + 1], x.y, glp);
                    glp = fmaf(a_[4 * i + 2], x.z, glp); glp = fmaf(a_[4 * i + 3], x.w, glp);
                }
            }
            if (act) sm.glp[hf][row] = glp;
            fence_before_sync();
            mbar_arrive_warp(&sm.glp_done);
        }
        if (c == 0 && P.ds0 != nullptr) {
            float *dst = P.ds0 + (size_t)bh * kC * kC + row * kC;
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int cb = 2 * hf + cc;
                float v[16];
                tmem_ld16(tb + C_DS + 16 * cb, v);
                tmem_wait_ld();
                if (act) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        *reinterpret_cast<float4 *>(dst + 16 * cb + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
        fence_before_sync();
        mbar_arrive_warp(&sm.ok_free[it & 1]);
        bar_sync(5, 256);       // staged tiles complete; suff / cfirst / xch / glp in place for the next chunk
        {   // six gradient tiles [token][channel] bf16 -> 128-byte rows
            bf16 *dst[6] = {P.dw, P.dq, P.dk, P.dv, P.da, P.db};
            uint4 v[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int e = tid + 256 * i, arr = e >> 7, tok = (e >> 3) & 15, part = e & 7;
                v[i] = *reinterpret_cast<const uint4 *>(&obuf[arr][tok][part * 8]);
            }
            mbar_arrive_warp(&sm.empty[si]);            // the slot (tiles and staging) has been read
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int e = tid + 256 * i, arr = e >> 7, tok = (e >> 3) & 15, part = e & 7;
                if (!kVar || c * L + tok < len)
                    *reinterpret_cast<uint4 *>(dst[arr] + base + (size_t)(c * L + tok) * tok_stride + part * 8) = v[i];
            }
        }
        TICK(tc6); ACC(14, tc4, tc5); ACC(15, tc5, tc6);
    }
}

constexpr int kMmaWarp = 24, kThreads = 32 * (kMmaWarp + 1);

template <bool kVar>
__global__ void __launch_bounds__(kThreads, 1) wkv7_tc_bwd_kernel(const Params P, const __grid_constant__ TmaMaps M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const SeqWork W = seq_work(P.T, P.H, kVar ? P.cu : nullptr, P.cbase);
    const int bh = W.bh, nC = W.nC;
    const int tid = threadIdx.x, warp = tid >> 5;
    const size_t tok_stride = (size_t)P.H * kC;
    const size_t base = W.base;
    if (kVar && nC == 0) return;         // an empty sequence of a packed launch (uniform over the CTA)

    if (tid == 0) {
        for (int i = 0; i < NS; i++) {
            mbar_init(&sm.full[i], 8 + 4); mbar_init(&sm.empty[i], 8); mbar_init(&sm.a_done[i], 8);
            mbar_init(&sm.blob_full[i], 1);
        }
        mbar_init(&sm.s0_full, 1);
        mbar_init(&sm.resc, 4); mbar_init(&sm.glp_done, 8); mbar_init(&sm.ok_free[0], 8); mbar_init(&sm.ok_free[1], 8);
        mbar_init(&sm.bar_z, 1); mbar_init(&sm.c_done, 4);
        mbar_init(&sm.out_ready[0], 1); mbar_init(&sm.out_ready[1], 1);
        for (int i = 0; i < NRAW; i++) mbar_init(&sm.raw_full[i], 1);
        mbar_fence_init();
#pragma unroll
        for (int i = 0; i < 7; i++) tma_prefetch_desc(&M.m[i]);
    }
    if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    if (warp < 4) group_c1(P, sm, bh, nC, tid);
    else if (warp < 12) group_c2<kVar>(P, sm, base, tok_stride, bh, nC, W.len, tid - 128);
    else if (warp < 20) stage_a<kVar>(P, M, sm, base, tok_stride, W.ck0, W.h, W.tok0, nC, W.len, tid - 384);
    else if (warp < 24) stage_b(P, sm, nC, tid - 640);
    else mma_warp(P, sm, W.ck0, nC);

    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace tcbwd

long long *g_tcb_dbg = nullptr;   // set by the profiling harness only

const char *tc_bwd_barrier_name(unsigned off) {
    using tcbwd::Smem;
    struct { size_t off; int n; const char *name; } t[] = {
        {offsetof(Smem, full), tcbwd::NS, "full[slot]"}, {offsetof(Smem, empty), tcbwd::NS, "empty[slot]"},
        {offsetof(Smem, a_done), tcbwd::NS, "a_done[slot]"}, {offsetof(Smem, blob_full), tcbwd::NS, "blob_full[slot]"},
        {offsetof(Smem, s0_full), 1, "s0_full"}, {offsetof(Smem, resc), 1, "resc"}, {offsetof(Smem, glp_done), 1, "glp_done"},
        {offsetof(Smem, ok_free), 2, "ok_free[parity]"}, {offsetof(Smem, bar_z), 1, "bar_z"}, {offsetof(Smem, c_done), 1, "c_done"},
        {offsetof(Smem, out_ready), 2, "out_ready[parity]"}, {offsetof(Smem, raw_full), tcbwd::NRAW, "raw_full[buffer]"}};
    for (auto &e : t)
        if (off >= e.off && off < e.off + 8 * (size_t)e.n) return e.name;
    return "unknown";
}

cudaError_t launch_tc_bwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, const void *dy, const float *ckT, const float *sa,
                          const float *sT, const float *dsT, void *dw, void *dq, void *dk, void *dv, void *da,
                          void *db, float *ds0, const int *cu, const int *cbase, cudaStream_t st) {
    using namespace tcbwd;
    static_assert(sizeof(Smem) <= 232448, "shared memory budget");
    cudaError_t e = cu ? cudaFuncSetAttribute(wkv7_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))
                       : cudaFuncSetAttribute(wkv7_tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    Params P{T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)a,
             (const bf16 *)b, (const bf16 *)dy, ckT, sa, sT, dsT, (bf16 *)dw, (bf16 *)dq, (bf16 *)dk, (bf16 *)dv,
             (bf16 *)da, (bf16 *)db, ds0, g_tcb_dbg, cu, cbase};
    if (watchdog_needs_install(1, st)) {
        e = watchdog_install(watchdog_record(), 2);
        if (e != cudaSuccess) return e;
    }
    // tensor maps of the seven inputs: [tokens, H, 64] bf16, box = one 16-token chunk of one head (tma_map.h)
    TmaMaps M;
    {
        const long long n_tok = cu ? (long long)T : (long long)B * T;      // packed launch: T is T_total
        const void *src[7] = {w, q, k, v, a, b, dy};
        for (int i = 0; i < 7; i++) {
            e = make_chunk_map(&M.m[i], src[i], n_tok, H);
            if (e != cudaSuccess) return e;
        }
    }
    count_launch();
    if (cu) wkv7_tc_bwd_kernel<true><<<dim3(B * H), dim3(kThreads), sizeof(Smem), st>>>(P, M);
    else wkv7_tc_bwd_kernel<false><<<dim3(B * H), dim3(kThreads), sizeof(Smem), st>>>(P, M);
    return cudaGetLastError();
}

}  // namespace rwkvtts
