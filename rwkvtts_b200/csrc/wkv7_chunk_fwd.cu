// WKV-7 training forward, chunked (L = 16 tokens) on the tensor cores (mma.sync tf32, fp32 accumulate).
//
// Per chunk, with S the value-major state [64 values][64 keys] at the chunk start, g_t the
// cumulative log-decay inside the chunk, D_t = exp(g_t) (SURVEY.md section 7 "chunked DPLR form"):
//     A~ = a*D_{t-1}   B~ = b/D_t   K~ = k/D_t   Q~ = q*D_t                       [16 x 64]
//     N   = strict_tril(A~ B~^T)    Aak = strict_tril(A~ K~^T)
//     Aqb = tril(Q~ B~^T)           Aqk = tril(Q~ K~^T)                            [16 x 16]
//     [M1 | W] = (I - N)^-1 [Aak | A~]          (forward substitution, fp32 CUDA cores)
//     U^T = S W^T + V^T M1^T                    (U_t = S_{t-1} a_t, the reference's "sa")
//     Y^T = S Q~^T + U^T Aqb^T + V^T Aqk^T
//     S'  = (S + U^T B~ + V^T K~) diag(D_16)
// which is algebraically the recurrence of wkv7_cuda.cu:17-42 (validated against the oracle in
// proto/chunk_fwd_proto.py to 3e-16 in f64; with tf32-rounded operands 3.7e-4 relative).
//
// One CTA of 256 threads per (batch, head) chain:
//   warps 0-3 ("state" group): warp v owns value rows [16v,16v+16) of S as mma accumulators
//       (never leaves registers); per chunk 76 mma per warp, everything that depends on S.
//   warps 4-7 ("prep" group): everything that does not depend on S for the NEXT chunk --
//       global loads, decay scan, scaling, the four Gram blocks, the triangular solve -- written
//       to a double-buffered shared-memory stage.  One __syncthreads per chunk hands over.
// The state at every chunk start is written to `s` (the reference's scratch tensor, same size)
// for the backward kernel.
#include "mma_tf32.cuh"
#include "wkv7_common.cuh"

namespace rwkvtts {
namespace chunkfwd {

constexpr int L = 16;
constexpr int LD4 = 68;   // row stride for tiles read with 32-bit fragment loads (== 4 mod 32)
constexpr int LD8 = 72;   // row stride for tiles read with 64-bit fragment loads (== 8 mod 32)
constexpr int LS = 24;    // row stride of the 16x16 tiles (64-bit fragment loads)

struct Stage {
    float At[L * LD4], Bt[L * LD4], Kt[L * LD4], V[L * LD4];
    float Qt[L * LD8], W[L * LD8];
    float N[L * LS], Aak[L * LS], M1[L * LS], Aqb[L * LS], Aqk[L * LS];
    float DL[kC];
};
struct Smem {
    Stage st[2];
    float wtot[4][kC];
    __align__(16) bf16 y[L][72];
};

struct Params {
    int T, H;
    const bf16 *w, *q, *k, *v, *a, *b;
    bf16 *y;
    float *s;            // chunk-start states [B*H][T/16][64][64]
    const float *s0;     // may be null
    float *sT;           // may be null
    long long *dbg;      // phase-cycle counters (profiling builds only), may be null
};

#ifdef RWKVTTS_PROFILE
#define TICK(var) long long var = clock64()
#define ACC(slot, t0, t1) do { if (P.dbg && blockIdx.x == 0 && (threadIdx.x & 127) == 0) P.dbg[slot] += (t1) - (t0); } while (0)
#else
#define TICK(var)
#define ACC(slot, t0, t1)
#endif

__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

// ------------------------------------------------------------------------------------------
// prep group: tp in [0,128)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prep_load(const Params &P, size_t base, size_t tok_stride, int c, int tp,
                                          uint4 (&raw)[6]) {
    const int t = tp >> 3, kg = tp & 7;
    const size_t off = base + (size_t)(c * L + t) * tok_stride + kg * 8;
    raw[0] = ldg_nc_v4(P.w + off);
    raw[1] = ldg_nc_v4(P.q + off);
    raw[2] = ldg_nc_v4(P.k + off);
    raw[3] = ldg_nc_v4(P.v + off);
    raw[4] = ldg_nc_v4(P.a + off);
    raw[5] = ldg_nc_v4(P.b + off);
}

__device__ __forceinline__ void prep_chunk(const Params &P, Smem &sm, Stage &S, size_t base, size_t tok_stride,
                                           int c, int nC, int tp, uint4 (&raw)[6]) {
    const int t = tp >> 3, kg = tp & 7, wp = tp >> 5, lane = tp & 31, g = lane >> 2, tq = lane & 3;
    TICK(tp0);
    // ---- P0: decay scan + scaling ---------------------------------------------------------
    float lw[8], gg[8];
    {
        float f[8];
        unpack8(raw[0], f);
#pragma unroll
        for (int j = 0; j < 8; j++) { lw[j] = -__expf(f[j]); gg[j] = lw[j]; }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {   // inclusive scan over the 4 tokens of this warp (lane = (t&3)*8 + kg)
        float y = __shfl_up_sync(0xffffffffu, gg[j], 8);
        if ((t & 3) >= 1) gg[j] += y;
        y = __shfl_up_sync(0xffffffffu, gg[j], 16);
        if ((t & 3) >= 2) gg[j] += y;
    }
    if ((t & 3) == 3) {
        st4(&sm.wtot[wp][kg * 8], gg[0], gg[1], gg[2], gg[3]);
        st4(&sm.wtot[wp][kg * 8 + 4], gg[4], gg[5], gg[6], gg[7]);
    }
    bar_sync(1, 128);
    for (int ww = 0; ww < wp; ww++) {
        const float4 x0 = *reinterpret_cast<const float4 *>(&sm.wtot[ww][kg * 8]);
        const float4 x1 = *reinterpret_cast<const float4 *>(&sm.wtot[ww][kg * 8 + 4]);
        gg[0] += x0.x; gg[1] += x0.y; gg[2] += x0.z; gg[3] += x0.w;
        gg[4] += x1.x; gg[5] += x1.y; gg[6] += x1.z; gg[7] += x1.w;
    }
    {
        float D[8], Dp[8], iD[8], f[8], o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            D[j] = __expf(gg[j]);
            Dp[j] = __expf(gg[j] - lw[j]);
            iD[j] = __expf(-gg[j]);
        }
        const int o4 = t * LD4 + kg * 8, o8 = t * LD8 + kg * 8;
        unpack8(raw[1], f);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = tf32r(f[j] * D[j]);
        st4(&S.Qt[o8], o[0], o[1], o[2], o[3]); st4(&S.Qt[o8 + 4], o[4], o[5], o[6], o[7]);
        unpack8(raw[2], f);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = tf32r(f[j] * iD[j]);
        st4(&S.Kt[o4], o[0], o[1], o[2], o[3]); st4(&S.Kt[o4 + 4], o[4], o[5], o[6], o[7]);
        unpack8(raw[3], f);   // bf16 values are exact in tf32
        st4(&S.V[o4], f[0], f[1], f[2], f[3]); st4(&S.V[o4 + 4], f[4], f[5], f[6], f[7]);
        unpack8(raw[4], f);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = tf32r(f[j] * Dp[j]);
        st4(&S.At[o4], o[0], o[1], o[2], o[3]); st4(&S.At[o4 + 4], o[4], o[5], o[6], o[7]);
        unpack8(raw[5], f);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = tf32r(f[j] * iD[j]);
        st4(&S.Bt[o4], o[0], o[1], o[2], o[3]); st4(&S.Bt[o4 + 4], o[4], o[5], o[6], o[7]);
        if (t == L - 1) {
            st4(&S.DL[kg * 8], D[0], D[1], D[2], D[3]); st4(&S.DL[kg * 8 + 4], D[4], D[5], D[6], D[7]);
        }
    }
    if (c + 1 < nC) prep_load(P, base, tok_stride, c + 1, tp, raw);   // in flight during P1/P2
    bar_sync(1, 128);
    TICK(tp1); ACC(0, tp0, tp1);
    // ---- P1: Gram blocks, one 16x16 block per warp ------------------------------------------
    {
        const int rowsel = wp & 1, colsel = wp >> 1;
        const float *Ar = rowsel ? S.Qt : S.At;
        const int lda_ = rowsel ? LD8 : LD4;
        const float *Bc = colsel ? S.Kt : S.Bt;
        float acc[2][4] = {};
#pragma unroll
        for (int kb = 0; kb < 8; kb++) {
            uint32_t af[4], bfr[2];
            lda(af, Ar, lda_, 1, 0, 8 * kb, g, tq);
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                ldb(bfr, Bc, 1, LD4, 8 * kb, 8 * nt, g, tq);
                mma_tf32(acc[nt], af, bfr);
            }
        }
        float *dst = rowsel ? (colsel ? S.Aqk : S.Aqb) : (colsel ? S.Aak : S.N);
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = 8 * nt + 2 * tq + e;
                float v0 = acc[nt][e], v1 = acc[nt][2 + e];
                if (rowsel) {   // Q~ rows: inclusive lower triangle, consumed by mma -> tf32
                    v0 = (col <= g) ? tf32r(v0) : 0.f;
                    v1 = (col <= g + 8) ? tf32r(v1) : 0.f;
                } else {        // A~ rows: strict lower triangle, consumed by the fp32 solve
                    v0 = (col < g) ? v0 : 0.f;
                    v1 = (col < g + 8) ? v1 : 0.f;
                }
                dst[g * LS + col] = v0;
                dst[(g + 8) * LS + col] = v1;
            }
    }
    bar_sync(1, 128);
    TICK(tp2); ACC(1, tp1, tp2);
    // ---- P2: [M1 | W] = (I - N)^-1 [Aak | A~], one column per thread --------------------------
    if (tp < 16 + kC) {
        const int col = tp;
        float X[L];
#pragma unroll
        for (int tt = 0; tt < L; tt++) {
            float acc0 = (col < 16) ? S.Aak[tt * LS + col] : S.At[tt * LD4 + col - 16];
            float acc1 = 0.f;
#pragma unroll
            for (int s4 = 0; s4 < (tt + 3) / 4; s4++) {
                const float4 n4 = *reinterpret_cast<const float4 *>(&S.N[tt * LS + 4 * s4]);
                const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int s = 4 * s4 + e;
                    if (s < tt) {
                        if (e & 1) acc1 = fmaf(nn[e], X[s], acc1);
                        else acc0 = fmaf(nn[e], X[s], acc0);
                    }
                }
            }
            X[tt] = acc0 + acc1;
        }
        if (col < 16) {
#pragma unroll
            for (int tt = 0; tt < L; tt++) S.M1[tt * LS + col] = tf32r(X[tt]);
        } else {
#pragma unroll
            for (int tt = 0; tt < L; tt++) S.W[tt * LD8 + col - 16] = tf32r(X[tt]);
        }
    }
    TICK(tp3); ACC(2, tp2, tp3);
}

// ------------------------------------------------------------------------------------------
// state group: tid in [0,128)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void state_chunk(const Params &P, Smem &sm, const Stage &S, float (&Sacc)[8][4],
                                            size_t base, size_t tok_stride, int bh, int c, int nC, int tid) {
    const int wv = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int r0 = 16 * wv + g;
    TICK(ts0);
    {   // state at the start of this chunk -> checkpoint for the backward kernel
        float *ck = P.s + ((size_t)bh * nC + c) * (kC * kC);
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            *reinterpret_cast<float2 *>(ck + r0 * kC + 8 * nt + 2 * tq) = make_float2(Sacc[nt][0], Sacc[nt][1]);
            *reinterpret_cast<float2 *>(ck + (r0 + 8) * kC + 8 * nt + 2 * tq) = make_float2(Sacc[nt][2], Sacc[nt][3]);
        }
    }
    TICK(ts1); ACC(4, ts0, ts1);
    float Uacc[2][4] = {}, Yacc[2][4] = {};
    // U^T = S W^T,  Y^T = S Q~^T        (A = S from the accumulators, k = key, permuted order)
#pragma unroll
    for (int kb = 0; kb < 8; kb++) {
        uint32_t af[4], bfr[2];
        acc_to_a_perm(af, Sacc[kb]);
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            ldb_perm_k1(bfr, S.W, LD8, 8 * kb, 8 * nt, g, tq);
            mma_tf32(Uacc[nt], af, bfr);
            ldb_perm_k1(bfr, S.Qt, LD8, 8 * kb, 8 * nt, g, tq);
            mma_tf32(Yacc[nt], af, bfr);
        }
    }
    TICK(ts2); ACC(5, ts1, ts2);
    // + V^T M1^T, + V^T Aqk^T           (A[m=value][k=token] = V[token][value], permuted k)
    uint32_t Va[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++) lda_perm(Va[j], S.V, 1, LD4, 16 * wv, 8 * j, g, tq);
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            uint32_t bfr[2];
            ldb_perm_k1(bfr, S.M1, LS, 8 * j, 8 * nt, g, tq);     // B[k=s][n=t] = M1[t][s]
            mma_tf32(Uacc[nt], Va[j], bfr);
            ldb_perm_k1(bfr, S.Aqk, LS, 8 * j, 8 * nt, g, tq);
            mma_tf32(Yacc[nt], Va[j], bfr);
        }
    // Y^T += U^T Aqb^T ;  S += U^T B~ + V^T K~   (A = U^T from the accumulators, k = token, permuted)
    uint32_t Ua[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++) acc_to_a_perm(Ua[j], Uacc[j]);
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            uint32_t bfr[2];
            ldb_perm_k1(bfr, S.Aqb, LS, 8 * j, 8 * nt, g, tq);
            mma_tf32(Yacc[nt], Ua[j], bfr);
        }
    TICK(ts3); ACC(6, ts2, ts3);
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            uint32_t bfr[2];
            ldb_perm(bfr, S.Bt, LD4, 1, 8 * j, 8 * nt, g, tq);     // B[k=s][n=key] = B~[s][key]
            mma_tf32(Sacc[nt], Ua[j], bfr);
            ldb_perm(bfr, S.Kt, LD4, 1, 8 * j, 8 * nt, g, tq);
            mma_tf32(Sacc[nt], Va[j], bfr);
        }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const float2 d = *reinterpret_cast<const float2 *>(&S.DL[8 * nt + 2 * tq]);
        Sacc[nt][0] *= d.x; Sacc[nt][1] *= d.y; Sacc[nt][2] *= d.x; Sacc[nt][3] *= d.y;
    }
    TICK(ts4); ACC(7, ts3, ts4);
    // y: transpose through shared memory, then 128-byte rows to HBM
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            sm.y[8 * nt + 2 * tq + e][r0] = __float2bfloat16_rn(Yacc[nt][e]);
            sm.y[8 * nt + 2 * tq + e][r0 + 8] = __float2bfloat16_rn(Yacc[nt][2 + e]);
        }
    bar_sync(2, 128);
    {
        const int tok = tid >> 3, part = tid & 7;
        const uint4 v = *reinterpret_cast<const uint4 *>(&sm.y[tok][part * 8]);
        *reinterpret_cast<uint4 *>(P.y + base + (size_t)(c * L + tok) * tok_stride + part * 8) = v;
    }
    TICK(ts5); ACC(8, ts4, ts5);
}

__global__ void __launch_bounds__(256) wkv7_chunk_fwd_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int bh = blockIdx.x, bb = bh / P.H, hh = bh % P.H;
    const int tid = threadIdx.x;
    const bool is_prep = tid >= 128;
    const int nC = P.T / L;
    const size_t tok_stride = (size_t)P.H * kC;
    const size_t base = (size_t)bb * P.T * tok_stride + (size_t)hh * kC;

    float Sacc[8][4];
    uint4 raw[6];
    if (is_prep) {
        prep_load(P, base, tok_stride, 0, tid - 128, raw);
        prep_chunk(P, sm, sm.st[0], base, tok_stride, 0, nC, tid - 128, raw);
    } else {
        const int wv = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3, r0 = 16 * wv + g;
        if (P.s0 != nullptr) {
            const float *sp = P.s0 + (size_t)bh * kC * kC;
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const float2 x = *reinterpret_cast<const float2 *>(sp + r0 * kC + 8 * nt + 2 * tq);
                const float2 z = *reinterpret_cast<const float2 *>(sp + (r0 + 8) * kC + 8 * nt + 2 * tq);
                Sacc[nt][0] = x.x; Sacc[nt][1] = x.y; Sacc[nt][2] = z.x; Sacc[nt][3] = z.y;
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 8; nt++) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;
        }
    }
    __syncthreads();
    for (int c = 0; c < nC; c++) {
        TICK(tl0);
        if (is_prep) {
            if (c + 1 < nC) prep_chunk(P, sm, sm.st[(c + 1) & 1], base, tok_stride, c + 1, nC, tid - 128, raw);
        } else {
            state_chunk(P, sm, sm.st[c & 1], Sacc, base, tok_stride, bh, c, nC, tid);
        }
        TICK(tl1);
        __syncthreads();
        TICK(tl2); ACC(is_prep ? 10 : 12, tl0, tl1); ACC(is_prep ? 11 : 13, tl1, tl2);
    }
    if (!is_prep && P.sT != nullptr) {
        const int wv = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3, r0 = 16 * wv + g;
        float *sp = P.sT + (size_t)bh * kC * kC;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            *reinterpret_cast<float2 *>(sp + r0 * kC + 8 * nt + 2 * tq) = make_float2(Sacc[nt][0], Sacc[nt][1]);
            *reinterpret_cast<float2 *>(sp + (r0 + 8) * kC + 8 * nt + 2 * tq) = make_float2(Sacc[nt][2], Sacc[nt][3]);
        }
    }
}

}  // namespace chunkfwd

long long *g_dbg = nullptr;   // set by the profiling harness only

cudaError_t launch_chunk_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                             const void *a, const void *b, void *y, float *s, const float *s0, float *sT,
                             cudaStream_t st) {
    using namespace chunkfwd;
    cudaError_t e = cudaFuncSetAttribute(wkv7_chunk_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    Params P{T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)a,
             (const bf16 *)b, (bf16 *)y, s, s0, sT, g_dbg};
    count_launch();
    wkv7_chunk_fwd_kernel<<<dim3(B * H), dim3(256), sizeof(Smem), st>>>(P);
    return cudaGetLastError();
}

}  // namespace rwkvtts
