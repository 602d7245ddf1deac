// Fused elementwise kernels of the RWKV-7 time-mix around the WKV-7 op (sm_100a), forward and backward.
//
// Reference: RWKV_Tmix_x070.forward, model/llm/rwkv_s2s_single_ffn.py:158-196 (and RWKV_CMix_x070.forward :223-230 for
// the single-output token-shift lerp).  Between its GEMMs the reference runs ~30 elementwise ATen kernels over
// [B,T,C] activations per layer; here they are three kernels (each with its adjoint):
//
//   shift_mix   x -> x + (shift(x) - x) * mix_i, i < n          (:160-169; n = 6 time-mix, n = 1 channel-mix :226)
//   prep        k, v, LoRA outputs -> w, k', v', -kk, kk*a       (:172, :175-190: decay transform, gates, v-residual,
//                                                                 per-head l2-normalised kk, k update, WKV operands a, b)
//   out         y (WKV), r, k', v', g -> (GroupNorm(y) + bonus) * g   (:192-195, input of the output projection)
//
// All are HBM-bound streaming kernels: one thread owns 8 adjacent channels (16-byte bf16 loads / stores) of a row
// (= one token), a head (64 channels) is 8 adjacent lanes (reductions by shuffle), a CTA processes kRows rows at a time
// and strides over the rows with a fixed thread <-> channel assignment, so per-channel parameters live in registers and
// per-channel parameter gradients are accumulated in registers, reduced over the CTA's rows in shared memory and
// written as one partial per CTA (summed by reduce_partials: deterministic, no atomics).
// Arithmetic in fp32, one rounding to bf16 at the stores.
#include "tc05.cuh"
#include "wkv7_common.cuh"

#include <cstdlib>

namespace rwkvtts {
namespace tmixf {

constexpr int kVec = 8;          // channels per thread
constexpr int kMaxThreads = 512;
constexpr int kBwdThreads = 256;  // backward kernels hold the parameter-gradient accumulators: 2 CTAs of 256 threads per SM

struct Row8 { float v[kVec]; };

__device__ __forceinline__ Row8 ld8(const bf16 *p) {
    Row8 r;
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    unpack8(u, r.v);
    return r;
}
__device__ __forceinline__ Row8 ld8f(const float *p) {
    Row8 r;
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ Row8 zero8() {
    Row8 r;
#pragma unroll
    for (int i = 0; i < kVec; i++) r.v[i] = 0.f;
    return r;
}
__device__ __forceinline__ void st8(bf16 *p, const Row8 &r, bool pred = true) {
    if (!pred) return;
    uint4 u;
    u.x = pack2(r.v[0], r.v[1]); u.y = pack2(r.v[2], r.v[3]); u.z = pack2(r.v[4], r.v[5]); u.w = pack2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4 *>(p) = u;
}
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// sum over the 8 lanes that hold one head
__device__ __forceinline__ float head_sum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// CTA-level reduction of per-thread partial sums over the CTA's row lanes, written to part[blockIdx.x][slot][C]
template <int N>
__device__ __forceinline__ void write_partials(float (&acc)[N][kVec], float *part, int C, int tpr, int rl, int cl,
                                               float *smem /* [kRows][N*C] */) {
    const int nrl = blockDim.x / tpr;
    float *mine = smem + (size_t)rl * N * C;
#pragma unroll
    for (int s = 0; s < N; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) mine[s * C + cl * kVec + i] = acc[s][i];
    __syncthreads();
    float *dst = part + (size_t)blockIdx.x * N * C;
    for (int e = threadIdx.x; e < N * C; e += blockDim.x) {
        float x = 0.f;
        for (int j = 0; j < nrl; j++) x += smem[(size_t)j * N * C + e];
        dst[e] = x;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// shift_mix
// ------------------------------------------------------------------------------------------------------------------
struct MixParams {
    const bf16 *x;          // [B,T,C]
    const bf16 *mask;       // [B,T] or null
    const bf16 *prev;       // [B,C] or null (last token of the previous call)
    bf16 *prev_out;         // [B,C] or null: receives the (masked) last token; may alias `prev` only when T == 1
    const float *mix;       // [n][C]
    bf16 *out[6];           // n outputs [B,T,C]
    const bf16 *dout[6];    // backward: n output gradients
    bf16 *dx;               // backward: [B,T,C]
    float *part;            // backward: [grid][n][C] partial d mix
    const unsigned char *first;   // packed (cu_seqlens) batches: [B*T], 1 on the first token of every sequence; null = dense
    int B, T, C, n;
};
// sequence boundaries: every T rows (dense) or where the caller's flags say so (packed: the shift must not reach across
// two sequences, utils/multiple_jsonl.py:77-135)
__device__ __forceinline__ bool seq_first(const MixParams &P, long row) {
    return P.first != nullptr ? (P.first[row] != 0) : (row % P.T == 0);
}
__device__ __forceinline__ bool seq_last(const MixParams &P, long row, long rows) {
    return P.first != nullptr ? (row + 1 >= rows || P.first[row + 1] != 0) : ((row + 1) % P.T == 0);
}

// Each row lane (C/8 threads) walks a CONTIGUOUS range of rows, so the neighbour row a token needs (t-1 in the forward,
// t+1 and t-1 in the backward) is the row it handled one step earlier / handles next: every element is loaded once.
__device__ __forceinline__ Row8 masked(const Row8 &x, float m) {
    Row8 r;
#pragma unroll
    for (int i = 0; i < kVec; i++) r.v[i] = x.v[i] * m;
    return r;
}

template <int N>
__global__ void __launch_bounds__(kMaxThreads) shift_mix_fwd_kernel(const MixParams P) {
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    float mix[N][kVec];
#pragma unroll
    for (int s = 0; s < N; s++) {
        const Row8 m = ld8f(P.mix + (size_t)s * P.C + c0);
#pragma unroll
        for (int i = 0; i < kVec; i++) mix[s][i] = m.v[i];
    }
    const long rows = (long)P.B * P.T;
    const long lanes = (long)gridDim.x * nrl, per = (rows + lanes - 1) / lanes;
    const long r0 = ((long)blockIdx.x * nrl + rl) * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    auto mask_of = [&](long row) { return P.mask != nullptr ? __bfloat162float(P.mask[row]) : 1.f; };
    Row8 xp = zero8(), x = zero8(), xn = zero8();
    if (r0 < r1) {
        x = masked(ld8(P.x + r0 * P.C + c0), mask_of(r0));
        if (!seq_first(P, r0)) xp = masked(ld8(P.x + (r0 - 1) * P.C + c0), mask_of(r0 - 1));
        if (r0 + 1 < r1) xn = masked(ld8(P.x + (r0 + 1) * P.C + c0), mask_of(r0 + 1));
    }
    for (long row = r0; row < r1; row++) {
        const int t = (int)(row % P.T);
        if (seq_first(P, row)) xp = (P.prev != nullptr && P.first == nullptr) ? ld8(P.prev + (size_t)(row / P.T) * P.C + c0) : zero8();
        Row8 xn2 = zero8();
        if (row + 2 < r1) xn2 = masked(ld8(P.x + (row + 2) * P.C + c0), mask_of(row + 2));     // two rows ahead, in flight
        Row8 xx;
#pragma unroll
        for (int i = 0; i < kVec; i++) xx.v[i] = rbf(xp.v[i] - x.v[i]);       // the reference rounds shift(x) - x to bf16
#pragma unroll
        for (int s = 0; s < N; s++) {
            Row8 o;
#pragma unroll
            for (int i = 0; i < kVec; i++) o.v[i] = fmaf(xx.v[i], mix[s][i], x.v[i]);
            st8(P.out[s] + row * P.C + c0, o);
        }
        if (P.prev_out != nullptr && t == P.T - 1) st8(P.prev_out + (size_t)(row / P.T) * P.C + c0, x);
        xp = x;
        x = xn;
        xn = xn2;
    }
}

template <int N>
__global__ void __launch_bounds__(kBwdThreads, 2) shift_mix_bwd_kernel(const MixParams P) {
    extern __shared__ float red[];
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    float acc[N][kVec];
#pragma unroll
    for (int s = 0; s < N; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
    const long rows = (long)P.B * P.T;
    const long lanes = (long)gridDim.x * nrl, per = (rows + lanes - 1) / lanes;
    const long r0 = ((long)blockIdx.x * nrl + rl) * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    auto mask_of = [&](long row) { return P.mask != nullptr ? __bfloat162float(P.mask[row]) : 1.f; };
    // descending walk: d[s][row+1] is what this lane loaded one step earlier, x[row-1] is the next step's x[row]
    uint4 dn[N];
#pragma unroll
    for (int s = 0; s < N; s++) dn[s] = make_uint4(0, 0, 0, 0);
    Row8 x = zero8();
    if (r0 < r1) {
        const long last = r1 - 1;
        x = masked(ld8(P.x + last * P.C + c0), mask_of(last));
        if (!seq_last(P, last, rows)) {        // the row after the range belongs to the same sequence
#pragma unroll
            for (int s = 0; s < N; s++) dn[s] = *reinterpret_cast<const uint4 *>(P.dout[s] + (last + 1) * P.C + c0);
        }
    }
    for (long row = r1 - 1; row >= r0; row--) {
        const int t = (int)(row % P.T);
        const float m = mask_of(row);
        Row8 xb = zero8();                    // the row below: this token's shift source and the next step's x
        if (row > 0) xb = masked(ld8(P.x + (row - 1) * P.C + c0), mask_of(row - 1));
        Row8 xp = xb;
        if (seq_first(P, row)) xp = (P.prev != nullptr && P.first == nullptr) ? ld8(P.prev + (size_t)(row / P.T) * P.C + c0) : zero8();
        if (seq_last(P, row, rows)) {
#pragma unroll
            for (int s = 0; s < N; s++) dn[s] = make_uint4(0, 0, 0, 0);
        }
        Row8 dx = zero8();
#pragma unroll
        for (int s = 0; s < N; s++) {
            const uint4 du = *reinterpret_cast<const uint4 *>(P.dout[s] + row * P.C + c0);
            Row8 d, dnx;
            unpack8(du, d.v);
            unpack8(dn[s], dnx.v);
            dn[s] = du;
            const Row8 mix = ld8f(P.mix + (size_t)s * P.C + c0);      // L1-resident; keeps 8N registers free
#pragma unroll
            for (int i = 0; i < kVec; i++) {
                acc[s][i] = fmaf(d.v[i], rbf(xp.v[i] - x.v[i]), acc[s][i]);
                dx.v[i] += d.v[i] * (1.f - mix.v[i]) + dnx.v[i] * mix.v[i];
            }
        }
#pragma unroll
        for (int i = 0; i < kVec; i++) dx.v[i] *= m;
        st8(P.dx + row * P.C + c0, dx);
        x = xb;
    }
    write_partials<N>(acc, P.part, P.C, tpr, rl, cl, red);
}


// ------------------------------------------------------------------------------------------------------------------
// shift_mix adjoint, second form (C % 128 == 0, 512 <= C <= 2048).  ncu on the first form (profiles/
// r02_fused_adjoints_full.txt): 128 registers -- 48 d-mix accumulators + 8-channel rows -- leave 16 warps per SM and the
// compiler cannot keep a row's seven loads in flight together, so every row costs several DRAM round trips (26 % of the
// HBM roofline, long-scoreboard stalls).  Here a thread owns FOUR channels: 24 accumulators, rows as packed 8-byte words,
// the next row's loads issued before the current row is consumed, lerp coefficients in shared memory -> ~3 CTAs of C/4
// threads per SM with 14 loads in flight per thread.  One row lane per CTA: the accumulators are the CTA's partial sums.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack4h(const uint2 &u, float (&f)[4]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
}
template <int N>
__global__ void __launch_bounds__(512) shift_mix_bwd4_kernel(const MixParams P) {
    extern __shared__ float smix[];                       // [N][C]
    const int c0 = threadIdx.x * 4;
    for (int e = threadIdx.x; e < N * P.C; e += blockDim.x) smix[e] = P.mix[e];
    __syncthreads();
    float acc[N][4];
#pragma unroll
    for (int s = 0; s < N; s++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[s][i] = 0.f;
    const long rows = (long)P.B * P.T;
    const long per = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    auto mask_of = [&](long row) { return P.mask != nullptr ? __bfloat162float(P.mask[row]) : 1.f; };
    auto ldx = [&](long row) { return *reinterpret_cast<const uint2 *>(P.x + row * P.C + c0); };
    const uint2 zero2 = make_uint2(0u, 0u);
    uint2 dn[N], dc[N], dp[N];                             // d rows: next (row+1), current, prefetched (row-1)
    uint2 xc = zero2, xb = zero2, xbb = zero2;             // x rows: current, below (row-1), prefetched (row-2)
#pragma unroll
    for (int s = 0; s < N; s++) { dn[s] = zero2; dc[s] = zero2; dp[s] = zero2; }
    if (r0 < r1) {
        const long last = r1 - 1;
        xc = ldx(last);
        if (last > 0) xb = ldx(last - 1);
#pragma unroll
        for (int s = 0; s < N; s++) dc[s] = *reinterpret_cast<const uint2 *>(P.dout[s] + last * P.C + c0);
        if (!seq_last(P, last, rows)) {
#pragma unroll
            for (int s = 0; s < N; s++) dn[s] = *reinterpret_cast<const uint2 *>(P.dout[s] + (last + 1) * P.C + c0);
        }
    }
    for (long row = r1 - 1; row >= r0; row--) {
        // loads of the NEXT iteration first: d[.][row-1] and x[row-2]
        if (row - 1 >= r0) {
#pragma unroll
            for (int s = 0; s < N; s++) dp[s] = *reinterpret_cast<const uint2 *>(P.dout[s] + (row - 1) * P.C + c0);
        }
        xbb = (row >= 2) ? ldx(row - 2) : zero2;
        const float m = mask_of(row), mb = row > 0 ? mask_of(row - 1) : 0.f;
        float x[4], xp[4];
        unpack4h(xc, x);
        unpack4h(xb, xp);
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] *= m; xp[i] *= mb; }
        if (seq_first(P, row)) {
            if (P.prev != nullptr && P.first == nullptr) {
                const uint2 pv = *reinterpret_cast<const uint2 *>(P.prev + (size_t)(row / P.T) * P.C + c0);
                unpack4h(pv, xp);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) xp[i] = 0.f;
            }
        }
        const bool last_of_seq = seq_last(P, row, rows);
        float dx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < N; s++) {
            float d[4], dnx[4];
            unpack4h(dc[s], d);
            unpack4h(last_of_seq ? zero2 : dn[s], dnx);
            const float4 mx = *reinterpret_cast<const float4 *>(smix + s * P.C + c0);
            const float mixv[4] = {mx.x, mx.y, mx.z, mx.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[s][i] = fmaf(d[i], rbf(xp[i] - x[i]), acc[s][i]);
                dx[i] += d[i] * (1.f - mixv[i]) + dnx[i] * mixv[i];
            }
        }
        *reinterpret_cast<uint2 *>(P.dx + row * P.C + c0) = make_uint2(pack2(dx[0] * m, dx[1] * m), pack2(dx[2] * m, dx[3] * m));
#pragma unroll
        for (int s = 0; s < N; s++) { dn[s] = dc[s]; dc[s] = dp[s]; }
        xc = xb; xb = xbb;
    }
    float *dst = P.part + (size_t)blockIdx.x * N * P.C;
#pragma unroll
    for (int s = 0; s < N; s++)
        *reinterpret_cast<float4 *>(dst + s * P.C + c0) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
}

// ------------------------------------------------------------------------------------------------------------------
// shift_mix backward, third form: the register pipeline of the form above keeps ONE row (7 x 8 bytes per thread) in flight
// per warp -- 43 KB per SM against the ~1.5 us of loaded HBM latency is 2 TB/s (Little), which is what it reaches.  Here
// the rows come in through a shared-memory ring filled with cp.async: 6 slots of (n + 1) rows, three rows (42 KB per CTA,
// two CTAs per SM) always in flight, independent of the registers; a thread still owns four channels and reads its 8 bytes
// of each array from the ring.  d[row + 1] and x[row] are carried in registers from the iteration before.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRing = 6;
template <int N>
__global__ void __launch_bounds__(512) shift_mix_bwd_ring_kernel(const MixParams P) {
    extern __shared__ __align__(16) unsigned char ring_dyn[];
    float *smix = reinterpret_cast<float *>(ring_dyn);                       // [N][C]
    bf16 *ring = reinterpret_cast<bf16 *>(smix + (size_t)N * P.C);            // [kRing][N + 1][C]: arrays 0..N-1 = d, N = x
    const int C = P.C, tid = threadIdx.x, nthr = blockDim.x, c0 = tid * 4;
    const int S = C / 8, lane_half = tid >= S ? 1 : 0, seg = tid - lane_half * S;      // nthr == 2 S: copy pieces
    for (int e = tid; e < N * C; e += nthr) smix[e] = P.mix[e];
    const long rows = (long)P.B * P.T;
    const long per = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    if (r0 >= r1) {                                                          // no rows: the partial sums are zeros
        float *dst = P.part + (size_t)blockIdx.x * N * C;
        for (int e = tid; e < N * C; e += nthr) dst[e] = 0.f;
        return;
    }
    auto slot = [&](long row, int a) { return ring + ((size_t)((row + kRing) % kRing) * (N + 1) + a) * C; };
    const long lo = r0 > 0 ? r0 - 1 : 0;                                     // x[r0 - 1] is the lowest row anybody reads
    auto issue = [&](long row) {                                             // one cp.async group per row, empty if not needed
        if (row >= lo && row < rows) {
#pragma unroll
            for (int a = lane_half; a <= N; a += 2) {
                const bf16 *src = (a < N ? P.dout[a] : P.x) + row * C + seg * 8;
                tc05::cp_async16(slot(row, a) + seg * 8, src);
            }
        }
        tc05::cp_async_commit();
    };
    auto mask_of = [&](long row) { return (P.mask != nullptr && row >= 0) ? __bfloat162float(P.mask[row]) : 1.f; };
    long next = r1;
#pragma unroll 1
    for (int i = 0; i < 5; i++) issue(next--);                              // rows r1 (for d[row + 1]) .. r1 - 4
    float acc[N][4], mixv[N][4];
    __syncthreads();                                                         // smix is complete
#pragma unroll
    for (int s = 0; s < N; s++) {
        const float4 mx = *reinterpret_cast<const float4 *>(smix + s * C + c0);
        mixv[s][0] = mx.x; mixv[s][1] = mx.y; mixv[s][2] = mx.z; mixv[s][3] = mx.w;
#pragma unroll
        for (int i = 0; i < 4; i++) acc[s][i] = 0.f;
    }
    const uint2 zero2 = make_uint2(0u, 0u);
    uint2 xc = zero2;
    // dx[t] = sum_s d_s[t] (1 - mix_s) + sum_s d_s[t+1] mix_s = A[t] - M[t] + M[t+1] with A = sum_s d_s, M = sum_s d_s mix_s:
    // only M[t+1] (4 floats) is carried from the iteration before, not the six d rows
    float mnext[4] = {0.f, 0.f, 0.f, 0.f};
    float m = mask_of(r1 - 1), mb = mask_of(r1 - 2);
    bool first_iter = true;
#pragma unroll 1
    for (long row = r1 - 1; row >= r0; row--) {
        const float mb2 = mask_of(row - 2);                                  // used two iterations from now
        tc05::cp_async_wait<2>();                                            // rows down to row - 1 have landed (own pieces)
        __syncthreads();                                                     // ... everybody's; slot of row + 2 is free
        issue(next--);                                                       // row - 4 into it
        if (first_iter) {
            first_iter = false;
            xc = *reinterpret_cast<const uint2 *>(slot(row, N) + c0);
            if (row + 1 < rows && !seq_last(P, row, rows)) {
#pragma unroll
                for (int s = 0; s < N; s++) {
                    float d[4];
                    unpack4h(*reinterpret_cast<const uint2 *>(slot(row + 1, s) + c0), d);
#pragma unroll
                    for (int i = 0; i < 4; i++) mnext[i] = fmaf(d[i], mixv[s][i], mnext[i]);
                }
            }
        }
        const uint2 xb = row > 0 ? *reinterpret_cast<const uint2 *>(slot(row - 1, N) + c0) : zero2;
        float x[4], xp[4];
        unpack4h(xc, x);
        unpack4h(xb, xp);
        const float mbb = row > 0 ? mb : 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] *= m; xp[i] *= mbb; }
        if (seq_first(P, row)) {
            if (P.prev != nullptr && P.first == nullptr) {
                const uint2 pv = *reinterpret_cast<const uint2 *>(P.prev + (size_t)(row / P.T) * C + c0);
                unpack4h(pv, xp);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) xp[i] = 0.f;
            }
        }
        const bool last_of_seq = seq_last(P, row, rows);
        float xx[4], asum[4] = {0.f, 0.f, 0.f, 0.f}, msum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; i++) xx[i] = rbf(xp[i] - x[i]);
#pragma unroll
        for (int s = 0; s < N; s++) {
            float d[4];
            unpack4h(*reinterpret_cast<const uint2 *>(slot(row, s) + c0), d);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[s][i] = fmaf(d[i], xx[i], acc[s][i]);
                asum[i] += d[i];
                msum[i] = fmaf(d[i], mixv[s][i], msum[i]);
            }
        }
        float dx[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dx[i] = (asum[i] - msum[i] + (last_of_seq ? 0.f : mnext[i])) * m;
            mnext[i] = msum[i];
        }
        *reinterpret_cast<uint2 *>(P.dx + row * C + c0) = make_uint2(pack2(dx[0], dx[1]), pack2(dx[2], dx[3]));
        xc = xb;
        m = mb; mb = mb2;
    }
    tc05::cp_async_wait<0>();
    float *dst = P.part + (size_t)blockIdx.x * N * C;
#pragma unroll
    for (int s = 0; s < N; s++)
        *reinterpret_cast<float4 *>(dst + s * C + c0) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
}

// ------------------------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------------------------
struct PrepParams {
    const bf16 *k, *v, *w_lo, *a_lo, *v_lo, *v_first;   // [B,T,C]; v_lo / v_first null on layer 0
    const bf16 *mask;                                     // [B,T] or null
    const float *w0, *a0, *v0, *k_k, *k_a;                // [C]
    bf16 *w, *k2, *v2, *a_op, *b_op;                      // outputs [B,T,C]  (v2 null on layer 0 without mask)
    // backward
    const bf16 *dw, *dk2, *dv2, *da_op, *db_op;
    bf16 *dk, *dv, *dw_lo, *da_lo, *dv_lo, *dv_first;
    float *part;                                          // [grid][5][C]: dw0, da0, dv0, dk_k, dk_a
    int B, T, C;
    int mask_rwk;   // 1: w, k, v are masked before use (in-repo stack, :175-178); 0: only kk and v' (rwkvfla)
};

__device__ __forceinline__ float neg_softplus_neg(float z) {   // -softplus(-z) with torch's threshold 20
    const float y = -z;
    return (y > 20.f) ? -y : -__logf(1.f + __expf(y));     // |error| <= 1e-7 absolute: the result is offset by -0.5
}

__global__ void __launch_bounds__(kMaxThreads) prep_fwd_kernel(const PrepParams P) {
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    const Row8 w0 = ld8f(P.w0 + c0), a0 = ld8f(P.a0 + c0), kk_ = ld8f(P.k_k + c0), ka = ld8f(P.k_a + c0);
    const bool has_v = P.v_lo != nullptr;
    const Row8 v0 = has_v ? ld8f(P.v0 + c0) : zero8();
    const long rows = (long)P.B * P.T;
    for (long base_row = (long)blockIdx.x * nrl; base_row < rows; base_row += (long)gridDim.x * nrl) {
        // warp-uniform trip count (lanes of different rows share a warp when C/8 is not a multiple of 32): rows past
        // the end are computed on the last row and not stored
        long row = base_row + rl;
        const bool valid = row < rows;
        if (!valid) row = rows - 1;
        const size_t off = row * P.C + c0;
        const float m = (P.mask != nullptr) ? __bfloat162float(P.mask[row]) : 1.f;
        const float mk = P.mask_rwk ? m : 1.f;
        // all of the row's loads are issued before the first use (one memory round trip per iteration, not two)
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        const uint4 rk = *reinterpret_cast<const uint4 *>(P.k + off), rwl = *reinterpret_cast<const uint4 *>(P.w_lo + off),
                    ral = *reinterpret_cast<const uint4 *>(P.a_lo + off);
        const uint4 rv = (P.v2 != nullptr) ? *reinterpret_cast<const uint4 *>(P.v + off) : zero4;
        const uint4 rvl = has_v ? *reinterpret_cast<const uint4 *>(P.v_lo + off) : zero4;
        const uint4 rvf = has_v ? *reinterpret_cast<const uint4 *>(P.v_first + off) : zero4;
        Row8 k, wl, al;
        unpack8(rk, k.v); unpack8(rwl, wl.v); unpack8(ral, al.v);
        Row8 w, a, u, o;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            k.v[i] *= mk;
            w.v[i] = (neg_softplus_neg(w0.v[i] + wl.v[i]) - 0.5f) * mk;
            a.v[i] = rbf(sigmoidf_(a0.v[i] + al.v[i]));
            u.v[i] = k.v[i] * kk_.v[i];
            ss = fmaf(u.v[i], u.v[i], ss);
        }
        st8(P.w + off, w, valid);
        const float inv = m / fmaxf(sqrtf(head_sum(ss)), 1e-12f);
#pragma unroll
        for (int i = 0; i < kVec; i++) { u.v[i] = rbf(u.v[i] * inv); o.v[i] = -u.v[i]; }     // kk (masked)
        st8(P.a_op + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = u.v[i] * a.v[i];
        st8(P.b_op + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = k.v[i] * (1.f + (a.v[i] - 1.f) * ka.v[i]);
        st8(P.k2 + off, o, valid);
        if (P.v2 != nullptr) {
            Row8 v;
            unpack8(rv, v.v);
            if (has_v) {
                Row8 vl, vf;
                unpack8(rvl, vl.v); unpack8(rvf, vf.v);
#pragma unroll
                for (int i = 0; i < kVec; i++) {
                    const float vm = v.v[i] * mk;         // the reference masks v before the residual mix (:178) ...
                    v.v[i] = (vm + (vf.v[i] - vm) * sigmoidf_(v0.v[i] + vl.v[i])) * m;     // ... and after it (:190)
                }
            } else {
#pragma unroll
                for (int i = 0; i < kVec; i++) v.v[i] *= m;
            }
            st8(P.v2 + off, v, valid);
        }
    }
}

__global__ void __launch_bounds__(kBwdThreads, 2) prep_bwd_kernel(const PrepParams P) {
    extern __shared__ float red[];
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    const Row8 w0 = ld8f(P.w0 + c0), a0 = ld8f(P.a0 + c0), kk_ = ld8f(P.k_k + c0), ka = ld8f(P.k_a + c0);
    const bool has_v = P.v_lo != nullptr;
    const Row8 v0 = has_v ? ld8f(P.v0 + c0) : zero8();
    float acc[5][kVec];
#pragma unroll
    for (int s = 0; s < 5; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
    const long rows = (long)P.B * P.T;
    for (long base_row = (long)blockIdx.x * nrl; base_row < rows; base_row += (long)gridDim.x * nrl) {
        // warp-uniform trip count (lanes of different rows share a warp when C/8 is not a multiple of 32): rows past
        // the end are computed on the last row and not stored
        long row = base_row + rl;
        const bool valid = row < rows;
        if (!valid) row = rows - 1;
        const size_t off = row * P.C + c0;
        const float m = (P.mask != nullptr) ? __bfloat162float(P.mask[row]) : 1.f;
        const float mk = P.mask_rwk ? m : 1.f;
        Row8 k = ld8(P.k + off);
        const Row8 wl = ld8(P.w_lo + off), al = ld8(P.a_lo + off);
        Row8 dw = zero8(), dk2 = zero8(), da_op = zero8(), db_op = zero8();
        if (valid) { dw = ld8(P.dw + off); dk2 = ld8(P.dk2 + off); da_op = ld8(P.da_op + off); db_op = ld8(P.db_op + off); }
        // the v branch's loads are issued with the others (one memory round trip per iteration)
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        const uint4 rdv2 = (P.dv != nullptr && valid) ? *reinterpret_cast<const uint4 *>(P.dv2 + off) : zero4;
        const uint4 rv = has_v ? *reinterpret_cast<const uint4 *>(P.v + off) : zero4;
        const uint4 rvl = has_v ? *reinterpret_cast<const uint4 *>(P.v_lo + off) : zero4;
        const uint4 rvf = has_v ? *reinterpret_cast<const uint4 *>(P.v_first + off) : zero4;
        Row8 a, u, o;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            k.v[i] *= mk;
            a.v[i] = sigmoidf_(a0.v[i] + al.v[i]);
            u.v[i] = k.v[i] * kk_.v[i];
            ss = fmaf(u.v[i], u.v[i], ss);
        }
        // w = (-softplus(-z) - 0.5) * m
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            o.v[i] = dw.v[i] * mk * sigmoidf_(-(w0.v[i] + wl.v[i]));
            acc[0][i] += o.v[i];
        }
        st8(P.dw_lo + off, o, valid);
        // kk = u / n * m ;  a_op = -kk, b_op = kk * a
        const float n = fmaxf(sqrtf(head_sum(ss)), 1e-12f), inv = 1.f / n;
        Row8 kk, dkk;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            kk.v[i] = u.v[i] * inv;
            dkk.v[i] = (db_op.v[i] * a.v[i] - da_op.v[i]) * m;
            dot = fmaf(kk.v[i], dkk.v[i], dot);
        }
        dot = head_sum(dot);
        // a: b_op = kk*m*a, k2 = k (1 + (a-1) k_a)
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            const float da = db_op.v[i] * kk.v[i] * m + dk2.v[i] * k.v[i] * ka.v[i];
            o.v[i] = da * a.v[i] * (1.f - a.v[i]);
            acc[1][i] += o.v[i];
        }
        st8(P.da_lo + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            const float du = (dkk.v[i] - kk.v[i] * dot) * inv;
            acc[3][i] = fmaf(du, k.v[i], acc[3][i]);
            acc[4][i] = fmaf(dk2.v[i] * k.v[i], a.v[i] - 1.f, acc[4][i]);
            o.v[i] = (du * kk_.v[i] + dk2.v[i] * (1.f + (a.v[i] - 1.f) * ka.v[i])) * mk;
        }
        st8(P.dk + off, o, valid);
        if (P.dv != nullptr) {
            Row8 dv2;
            unpack8(rdv2, dv2.v);
            if (has_v) {
                Row8 v, vl, vf;
                unpack8(rv, v.v); unpack8(rvl, vl.v); unpack8(rvf, vf.v);
                Row8 dvl, dvf;
#pragma unroll
                for (int i = 0; i < kVec; i++) {
                    const float s = sigmoidf_(v0.v[i] + vl.v[i]), g = dv2.v[i] * m, vm = v.v[i] * mk;
                    o.v[i] = g * (1.f - s) * mk;
                    dvf.v[i] = g * s;
                    dvl.v[i] = g * (vf.v[i] - vm) * s * (1.f - s);
                    acc[2][i] += dvl.v[i];
                }
                st8(P.dv + off, o, valid);
                st8(P.dv_lo + off, dvl, valid);
                st8(P.dv_first + off, dvf, valid);
            } else {
#pragma unroll
                for (int i = 0; i < kVec; i++) o.v[i] = dv2.v[i] * m;
                st8(P.dv + off, o, valid);
            }
        }
    }
    write_partials<5>(acc, P.part, P.C, tpr, rl, cl, red);
}

// ------------------------------------------------------------------------------------------------------------------
// prep backward, second form: the 11 input rows of a token come in through a shared-memory ring filled with cp.async
// (3 slots of 11 x C bf16, two rows -- 44 KB per CTA, three CTAs per SM -- in flight independent of the registers; the first
// form holds one row per warp in registers: 58 % of HBM with its 40 accumulators at 128 registers).  A thread copies exactly
// the 16-byte pieces it later reads (its 8 channels of every array), so the ring needs no CTA barrier: cp.async.wait_group
// is all the synchronisation there is.  One row lane per CTA of C/8 threads; rows strided over the grid.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPrepRing = 3, kPrepArrays = 11;
__device__ __forceinline__ Row8 ld8s(const bf16 *p) {
    Row8 r;
    unpack8(*reinterpret_cast<const uint4 *>(p), r.v);
    return r;
}
__global__ void __launch_bounds__(256) prep_bwd_ring_kernel(const PrepParams P) {
    extern __shared__ __align__(16) unsigned char prep_dyn[];
    bf16 *ring = reinterpret_cast<bf16 *>(prep_dyn);                         // [kPrepRing][kPrepArrays][C]
    const int C = P.C, c0 = threadIdx.x * kVec;
    const Row8 w0 = ld8f(P.w0 + c0), a0 = ld8f(P.a0 + c0), kk_ = ld8f(P.k_k + c0), ka = ld8f(P.k_a + c0);
    const bool has_v = P.v_lo != nullptr, want_dv = P.dv != nullptr;
    const Row8 v0 = has_v ? ld8f(P.v0 + c0) : zero8();
    // arrays in ring order: k, w_lo, a_lo, dw, dk2, da_op, db_op, dv2, v, v_lo, v_first
    const bf16 *const src[kPrepArrays] = {P.k, P.w_lo, P.a_lo, P.dw, P.dk2, P.da_op, P.db_op, want_dv ? P.dv2 : nullptr,
                                          has_v ? P.v : nullptr, has_v ? P.v_lo : nullptr, has_v ? P.v_first : nullptr};
    const long rows = (long)P.B * P.T;
    auto issue = [&](long it) {
        const long row = (long)blockIdx.x + it * gridDim.x;
        if (row < rows) {
            bf16 *dst = ring + (size_t)(it % kPrepRing) * kPrepArrays * C + c0;
#pragma unroll
            for (int a = 0; a < kPrepArrays; a++)
                if (src[a] != nullptr) tc05::cp_async16(dst + (size_t)a * C, src[a] + row * C + c0);
        }
        tc05::cp_async_commit();
    };
    float acc[5][kVec];
#pragma unroll
    for (int s = 0; s < 5; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
    issue(0);
    issue(1);
    auto mask_of = [&](long row) { return (P.mask != nullptr && row < rows) ? __bfloat162float(P.mask[row]) : 1.f; };
    float m_next = mask_of(blockIdx.x);
#pragma unroll 1
    for (long it = 0;; it++) {
        const long row = (long)blockIdx.x + it * gridDim.x;
        if (row >= rows) break;
        const float m = m_next, mk = P.mask_rwk ? m : 1.f;
        m_next = mask_of(row + gridDim.x);
        tc05::cp_async_wait<1>();                                            // this row's pieces (own) have landed
        const bf16 *slot = ring + (size_t)(it % kPrepRing) * kPrepArrays * C + c0;
        const size_t off = row * C + c0;
        Row8 k = ld8s(slot);
        const Row8 wl = ld8s(slot + (size_t)1 * C), al = ld8s(slot + (size_t)2 * C), dw = ld8s(slot + (size_t)3 * C),
                   dk2 = ld8s(slot + (size_t)4 * C), da_op = ld8s(slot + (size_t)5 * C), db_op = ld8s(slot + (size_t)6 * C);
        Row8 dv2 = zero8(), v = zero8(), vl = zero8(), vf = zero8();
        if (want_dv) dv2 = ld8s(slot + (size_t)7 * C);
        if (has_v) { v = ld8s(slot + (size_t)8 * C); vl = ld8s(slot + (size_t)9 * C); vf = ld8s(slot + (size_t)10 * C); }
        issue(it + 2);                                                       // into the slot read one iteration ago
        Row8 a, u, o;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            k.v[i] *= mk;
            a.v[i] = sigmoidf_(a0.v[i] + al.v[i]);
            u.v[i] = k.v[i] * kk_.v[i];
            ss = fmaf(u.v[i], u.v[i], ss);
        }
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            o.v[i] = dw.v[i] * mk * sigmoidf_(-(w0.v[i] + wl.v[i]));
            acc[0][i] += o.v[i];
        }
        st8(P.dw_lo + off, o);
        const float n = fmaxf(sqrtf(head_sum(ss)), 1e-12f), inv = 1.f / n;
        Row8 kk, dkk;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            kk.v[i] = u.v[i] * inv;
            dkk.v[i] = (db_op.v[i] * a.v[i] - da_op.v[i]) * m;
            dot = fmaf(kk.v[i], dkk.v[i], dot);
        }
        dot = head_sum(dot);
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            const float da = db_op.v[i] * kk.v[i] * m + dk2.v[i] * k.v[i] * ka.v[i];
            o.v[i] = da * a.v[i] * (1.f - a.v[i]);
            acc[1][i] += o.v[i];
        }
        st8(P.da_lo + off, o);
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            const float du = (dkk.v[i] - kk.v[i] * dot) * inv;
            acc[3][i] = fmaf(du, k.v[i], acc[3][i]);
            acc[4][i] = fmaf(dk2.v[i] * k.v[i], a.v[i] - 1.f, acc[4][i]);
            o.v[i] = (du * kk_.v[i] + dk2.v[i] * (1.f + (a.v[i] - 1.f) * ka.v[i])) * mk;
        }
        st8(P.dk + off, o);
        if (want_dv) {
            if (has_v) {
                Row8 dvl, dvf;
#pragma unroll
                for (int i = 0; i < kVec; i++) {
                    const float s = sigmoidf_(v0.v[i] + vl.v[i]), g = dv2.v[i] * m, vm = v.v[i] * mk;
                    o.v[i] = g * (1.f - s) * mk;
                    dvf.v[i] = g * s;
                    dvl.v[i] = g * (vf.v[i] - vm) * s * (1.f - s);
                    acc[2][i] += dvl.v[i];
                }
                st8(P.dv + off, o);
                st8(P.dv_lo + off, dvl);
                st8(P.dv_first + off, dvf);
            } else {
#pragma unroll
                for (int i = 0; i < kVec; i++) o.v[i] = dv2.v[i] * m;
                st8(P.dv + off, o);
            }
        }
    }
    tc05::cp_async_wait<0>();
    float *dst = P.part + (size_t)blockIdx.x * 5 * C + c0;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        *reinterpret_cast<float4 *>(dst + (size_t)s * C) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
        *reinterpret_cast<float4 *>(dst + (size_t)s * C + 4) = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// out
// ------------------------------------------------------------------------------------------------------------------
struct OutParams {
    const bf16 *y, *r, *k2, *v2, *g;          // [B,T,C]
    const float *r_k, *ln_w, *ln_b;           // [C]
    bf16 *o;                                  // [B,T,C]
    // backward
    const bf16 *d_o;
    bf16 *dy, *dr, *dk2, *dv2, *dg;
    float *part;                              // [grid][3][C]: dr_k, dln_w, dln_b
    float eps;
    int B, T, C;
};

__global__ void __launch_bounds__(kMaxThreads) out_fwd_kernel(const OutParams P) {
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    const Row8 rk = ld8f(P.r_k + c0), lw = ld8f(P.ln_w + c0), lb = ld8f(P.ln_b + c0);
    const long rows = (long)P.B * P.T;
    for (long base_row = (long)blockIdx.x * nrl; base_row < rows; base_row += (long)gridDim.x * nrl) {
        // warp-uniform trip count (lanes of different rows share a warp when C/8 is not a multiple of 32): rows past
        // the end are computed on the last row and not stored
        long row = base_row + rl;
        const bool valid = row < rows;
        if (!valid) row = rows - 1;
        const size_t off = row * P.C + c0;
        const Row8 y = ld8(P.y + off), r = ld8(P.r + off), k = ld8(P.k2 + off), v = ld8(P.v2 + off), g = ld8(P.g + off);
        float s1 = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) { s1 += y.v[i]; sb = fmaf(r.v[i] * k.v[i], rk.v[i], sb); }
        const float mu = head_sum(s1) * (1.f / kC);
        sb = head_sum(sb);
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) { const float d = y.v[i] - mu; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(head_sum(s2) * (1.f / kC) + P.eps);
        Row8 o;
#pragma unroll
        for (int i = 0; i < kVec; i++)
            o.v[i] = (rbf((y.v[i] - mu) * rstd * lw.v[i] + lb.v[i]) + sb * v.v[i]) * g.v[i];
        st8(P.o + off, o, valid);
    }
}

__global__ void __launch_bounds__(kBwdThreads, 2) out_bwd_kernel(const OutParams P) {
    extern __shared__ float red[];
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int c0 = cl * kVec;
    const Row8 rk = ld8f(P.r_k + c0), lw = ld8f(P.ln_w + c0), lb = ld8f(P.ln_b + c0);
    float acc[3][kVec];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
    const long rows = (long)P.B * P.T;
    for (long base_row = (long)blockIdx.x * nrl; base_row < rows; base_row += (long)gridDim.x * nrl) {
        // warp-uniform trip count (lanes of different rows share a warp when C/8 is not a multiple of 32): rows past
        // the end are computed on the last row and not stored
        long row = base_row + rl;
        const bool valid = row < rows;
        if (!valid) row = rows - 1;
        const size_t off = row * P.C + c0;
        const Row8 y = ld8(P.y + off), r = ld8(P.r + off), k = ld8(P.k2 + off), v = ld8(P.v2 + off), g = ld8(P.g + off);
        const Row8 d_o = valid ? ld8(P.d_o + off) : zero8();
        float s1 = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) { s1 += y.v[i]; sb = fmaf(r.v[i] * k.v[i], rk.v[i], sb); }
        const float mu = head_sum(s1) * (1.f / kC);
        sb = head_sum(sb);
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) { const float d = y.v[i] - mu; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(head_sum(s2) * (1.f / kC) + P.eps);
        Row8 yh, dz, o;
        float m1 = 0.f, m2 = 0.f, ds = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            yh.v[i] = (y.v[i] - mu) * rstd;
            const float z = yh.v[i] * lw.v[i] + lb.v[i] + sb * v.v[i];
            o.v[i] = d_o.v[i] * z;                                   // dg
            dz.v[i] = d_o.v[i] * g.v[i];
            acc[1][i] = fmaf(dz.v[i], yh.v[i], acc[1][i]);           // d ln_w
            acc[2][i] += dz.v[i];                                    // d ln_b
            const float dyh = dz.v[i] * lw.v[i];
            m1 += dyh;
            m2 = fmaf(dyh, yh.v[i], m2);
            ds = fmaf(dz.v[i], v.v[i], ds);
        }
        st8(P.dg + off, o, valid);
        m1 = head_sum(m1) * (1.f / kC);
        m2 = head_sum(m2) * (1.f / kC);
        ds = head_sum(ds);
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = rstd * (dz.v[i] * lw.v[i] - m1 - yh.v[i] * m2);
        st8(P.dy + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = dz.v[i] * sb;
        st8(P.dv2 + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            o.v[i] = ds * k.v[i] * rk.v[i];
            acc[0][i] = fmaf(ds * r.v[i], k.v[i], acc[0][i]);        // d r_k
        }
        st8(P.dr + off, o, valid);
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = ds * r.v[i] * rk.v[i];
        st8(P.dk2 + off, o, valid);
    }
    write_partials<3>(acc, P.part, P.C, tpr, rl, cl, red);
}

// part [G][n] -> out [n]: 32 columns x 8 slices of G per CTA (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float *part, float *out, int G, int n) {
    __shared__ float sh[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), gl = threadIdx.x >> 5;
    float x = 0.f;
    if (col < n)
        for (int g = gl; g < G; g += 8) x += part[(size_t)g * n + col];
    sh[gl][threadIdx.x & 31] = x;
    __syncthreads();
    if (gl == 0 && col < n) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) t += sh[j][threadIdx.x & 31];
        out[col] = t;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// channel-mix activation: y = relu(x)^2 (RWKV_CMix_x070.forward :228), dx = 2 relu(x) dy.  Pure streaming, 16 bytes per thread.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqrelu_fwd_kernel(const bf16 *x, bf16 *y, long n8) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        Row8 v = ld8(x + i * kVec);
#pragma unroll
        for (int j = 0; j < kVec; j++) { const float r = fmaxf(v.v[j], 0.f); v.v[j] = rbf(r) * rbf(r); }
        st8(y + i * kVec, v);
    }
}
__global__ void __launch_bounds__(256) sqrelu_bwd_kernel(const bf16 *x, const bf16 *dy, bf16 *dx, long n8) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const Row8 v = ld8(x + i * kVec), g = ld8(dy + i * kVec);
        Row8 o;
#pragma unroll
        for (int j = 0; j < kVec; j++) o.v[j] = 2.f * fmaxf(v.v[j], 0.f) * g.v[j];
        st8(dx + i * kVec, o);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// residual add + LayerNorm (Block.forward, rwkv_s2s_single_ffn.py:251-259: x + att(ln1(x)), x + ffn(ln2(x)); rwkvfla's
// fused add+norm): s = x + res, y = (s - mean) * rstd * w + b.  One row = C/8 threads = whole warps (C % 256 == 0), the
// two row reductions go warp shuffle -> shared memory -> every thread.  stats [rows][2] = (mean, rstd) for the backward.
// ------------------------------------------------------------------------------------------------------------------
struct LnParams {
    const bf16 *x, *res;          // [rows, C]; res may be null
    const float *w, *b;           // [C]; b may be null
    bf16 *y, *s;                  // [rows, C]; s (the sum) may be null when res is null
    float *stats;                 // [rows][2] or null
    // backward
    const bf16 *dy, *ds, *sum;    // ds may be null; sum = the forward's s (or x when there was no residual)
    bf16 *dx;                     // = d res
    float *part;                  // [grid][2][C]: dw, db
    float eps;
    long rows;
    int C;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// sum over the C/8 threads of a row lane: red [nrl][wpr] floats, two barriers (uniform trip counts in the callers)
__device__ __forceinline__ float row_sum(float x, float *red, int rl, int wpr, int wil) {
    x = warp_sum(x);
    if ((threadIdx.x & 31) == 0) red[rl * wpr + wil] = x;
    __syncthreads();
    float t = 0.f;
    for (int j = 0; j < wpr; j++) t += red[rl * wpr + j];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kMaxThreads) add_ln_fwd_kernel(const LnParams P) {
    __shared__ float red[kMaxThreads / 32];
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int wpr = tpr / 32, wil = cl / 32, c0 = cl * kVec;
    const Row8 w = ld8f(P.w + c0), b = P.b != nullptr ? ld8f(P.b + c0) : zero8();
    const float invC = 1.f / P.C;
    for (long base_row = (long)blockIdx.x * nrl; base_row < P.rows; base_row += (long)gridDim.x * nrl) {
        long row = base_row + rl;
        const bool valid = row < P.rows;
        if (!valid) row = P.rows - 1;
        const size_t off = row * P.C + c0;
        Row8 x = ld8(P.x + off);
        if (P.res != nullptr) {
            const Row8 r = ld8(P.res + off);
#pragma unroll
            for (int i = 0; i < kVec; i++) x.v[i] = rbf(x.v[i] + r.v[i]);      // the sum is a bf16 tensor in the reference
            if (P.s != nullptr) st8(P.s + off, x, valid);
        }
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) s1 += x.v[i];
        const float mu = row_sum(s1, red, rl, wpr, wil) * invC;
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) { const float d = x.v[i] - mu; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(row_sum(s2, red, rl, wpr, wil) * invC + P.eps);
        Row8 o;
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = (x.v[i] - mu) * rstd * w.v[i] + b.v[i];
        st8(P.y + off, o, valid);
        if (P.stats != nullptr && valid && cl == 0) { P.stats[2 * row] = mu; P.stats[2 * row + 1] = rstd; }
    }
}

// residual add + LayerNorm forward, second form (C % 256 == 0, C <= 1024): ONE WARP PER ROW.  A lane owns C/256 groups of 8
// channels, both row reductions are shuffles (the first form crosses the CTA through shared memory with two barriers per
// reduction: 55 % of the HBM roofline), and a warp has all 2 * C/256 loads of a row in flight together.
template <int NV>
__global__ void __launch_bounds__(256) add_ln_fwd_warp_kernel(const LnParams P) {
    const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float invC = 1.f / P.C;
    const bool has_res = P.res != nullptr;
    const long gw = (long)blockIdx.x * nwarp + wip, nw = (long)gridDim.x * nwarp;
    for (long row = gw; row < P.rows; row += nw) {
        uint4 ux[NV], ur[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const size_t off = row * P.C + (size_t)(j * 32 + lane) * kVec;
            ux[j] = *reinterpret_cast<const uint4 *>(P.x + off);
            if (has_res) ur[j] = *reinterpret_cast<const uint4 *>(P.res + off);
        }
        float xs[NV][kVec];
        float s1 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            unpack8(ux[j], xs[j]);
            if (has_res) {
                float r[kVec];
                unpack8(ur[j], r);
                Row8 t;
#pragma unroll
                for (int i = 0; i < kVec; i++) { xs[j][i] = rbf(xs[j][i] + r[i]); t.v[i] = xs[j][i]; }   // the sum is a bf16 tensor in the reference
                if (P.s != nullptr) st8(P.s + row * P.C + (size_t)(j * 32 + lane) * kVec, t);
            }
#pragma unroll
            for (int i = 0; i < kVec; i++) s1 += xs[j][i];
        }
        const float mu = warp_sum(s1) * invC;
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++)
#pragma unroll
            for (int i = 0; i < kVec; i++) { const float d = xs[j][i] - mu; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(warp_sum(s2) * invC + P.eps);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c0 = (j * 32 + lane) * kVec;
            const Row8 w = ld8f(P.w + c0), b = P.b != nullptr ? ld8f(P.b + c0) : zero8();
            Row8 o;
#pragma unroll
            for (int i = 0; i < kVec; i++) o.v[i] = (xs[j][i] - mu) * rstd * w.v[i] + b.v[i];
            st8(P.y + row * P.C + c0, o);
        }
        if (P.stats != nullptr && lane == 0) { P.stats[2 * row] = mu; P.stats[2 * row + 1] = rstd; }
    }
}

__global__ void __launch_bounds__(kBwdThreads, 2) add_ln_bwd_kernel(const LnParams P) {
    extern __shared__ float redp[];
    __shared__ float red[2 * kBwdThreads / 32];
    const int tpr = P.C / kVec, rl = threadIdx.x / tpr, cl = threadIdx.x % tpr, nrl = blockDim.x / tpr;
    const int wpr = tpr / 32, wil = cl / 32, c0 = cl * kVec;
    const Row8 w = ld8f(P.w + c0);
    const float invC = 1.f / P.C;
    float acc[2][kVec];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
    for (long base_row = (long)blockIdx.x * nrl; base_row < P.rows; base_row += (long)gridDim.x * nrl) {
        long row = base_row + rl;
        const bool valid = row < P.rows;
        if (!valid) row = P.rows - 1;
        const size_t off = row * P.C + c0;
        const Row8 sm_ = ld8(P.sum + off), dy = valid ? ld8(P.dy + off) : zero8();
        const float mu = P.stats[2 * row], rstd = P.stats[2 * row + 1];
        Row8 sh, g;
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            sh.v[i] = (sm_.v[i] - mu) * rstd;
            g.v[i] = dy.v[i] * w.v[i];
            acc[0][i] = fmaf(dy.v[i], sh.v[i], acc[0][i]);
            acc[1][i] += dy.v[i];
            m1 += g.v[i];
            m2 = fmaf(g.v[i], sh.v[i], m2);
        }
        // both row means in one pass over the shared scratch
        m1 = warp_sum(m1); m2 = warp_sum(m2);
        if ((threadIdx.x & 31) == 0) { red[2 * (rl * wpr + wil)] = m1; red[2 * (rl * wpr + wil) + 1] = m2; }
        __syncthreads();
        m1 = 0.f; m2 = 0.f;
        for (int j = 0; j < wpr; j++) { m1 += red[2 * (rl * wpr + j)]; m2 += red[2 * (rl * wpr + j) + 1]; }
        __syncthreads();
        m1 *= invC; m2 *= invC;
        Row8 o;
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = rstd * (g.v[i] - m1 - sh.v[i] * m2);
        if (P.ds != nullptr && valid) {
            const Row8 d2 = ld8(P.ds + off);
#pragma unroll
            for (int i = 0; i < kVec; i++) o.v[i] += d2.v[i];
        }
        st8(P.dx + off, o, valid);
    }
    write_partials<2>(acc, P.part, P.C, tpr, rl, cl, redp);
}


// residual-add + LayerNorm adjoint, second form (C % 256 == 0, C <= 1024): ONE WARP PER ROW.  The first form reduces each
// row across a CTA (two __syncthreads per row, two rows in flight per CTA: 42 % of the HBM roofline in the train step);
// here a lane owns C/256 groups of 8 channels, both row reductions are shuffles, and a warp has all 3 * C/256 loads of a
// row in flight together.  dw / db partials: per lane in registers, summed over the CTA's warps through shared memory.
template <int NV>
__global__ void __launch_bounds__(NV <= 3 ? 256 : 320, NV <= 3 ? 2 : 1) add_ln_bwd_warp_kernel(const LnParams P) {
    extern __shared__ float redw[];                        // [2][C] per CTA
    const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float invC = 1.f / P.C;
    for (int e = threadIdx.x; e < 2 * P.C; e += blockDim.x) redw[e] = 0.f;
    __syncthreads();
    float acc[2][NV][kVec];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int j = 0; j < NV; j++)
#pragma unroll
            for (int i = 0; i < kVec; i++) acc[s][j][i] = 0.f;
    const long gw = (long)blockIdx.x * nwarp + wip, nw = (long)gridDim.x * nwarp;
    for (long row = gw; row < P.rows; row += nw) {
        uint4 us[NV], ud[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const size_t off = row * P.C + (size_t)(j * 32 + lane) * kVec;
            us[j] = *reinterpret_cast<const uint4 *>(P.sum + off);
            ud[j] = *reinterpret_cast<const uint4 *>(P.dy + off);
        }
        const float mu = P.stats[2 * row], rstd = P.stats[2 * row + 1];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            float sv[kVec], dv[kVec];
            unpack8(us[j], sv);
            unpack8(ud[j], dv);
            const Row8 w = ld8f(P.w + (j * 32 + lane) * kVec);
#pragma unroll
            for (int i = 0; i < kVec; i++) {
                const float sh = (sv[i] - mu) * rstd, g = dv[i] * w.v[i];
                acc[0][j][i] = fmaf(dv[i], sh, acc[0][j][i]);
                acc[1][j][i] += dv[i];
                m1 += g;
                m2 = fmaf(g, sh, m2);
            }
        }
        m1 = warp_sum(m1) * invC;
        m2 = warp_sum(m2) * invC;
        // second pass over the row kept packed in registers (sh and g are recomputed: two multiplies instead of 64 registers)
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const size_t off = row * P.C + (size_t)(j * 32 + lane) * kVec;
            float sv[kVec], dv[kVec], d2[kVec];
            unpack8(us[j], sv);
            unpack8(ud[j], dv);
            if (P.ds != nullptr) unpack8(*reinterpret_cast<const uint4 *>(P.ds + off), d2);
            else {
#pragma unroll
                for (int i = 0; i < kVec; i++) d2[i] = 0.f;
            }
            const Row8 w = ld8f(P.w + (j * 32 + lane) * kVec);
            Row8 o;
#pragma unroll
            for (int i = 0; i < kVec; i++) {
                const float sh = (sv[i] - mu) * rstd, g = dv[i] * w.v[i];
                o.v[i] = rstd * (g - m1 - sh * m2) + d2[i];
            }
            st8(P.dx + off, o);
        }
    }
    // the warps of a CTA add their partials one after the other (fixed order: deterministic)
    for (int w = 0; w < nwarp; w++) {
        if (w == wip) {
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int j = 0; j < NV; j++)
#pragma unroll
                    for (int i = 0; i < kVec; i++) redw[s * P.C + (j * 32 + lane) * kVec + i] += acc[s][j][i];
        }
        __syncthreads();
    }
    float *dst = P.part + (size_t)blockIdx.x * 2 * P.C;
    for (int e = threadIdx.x; e < 2 * P.C; e += blockDim.x) dst[e] = redw[e];
}

// residual-add + LayerNorm adjoint, third form (C % 256 == 0, C <= 2048): the three input rows of a token (sum, dy, ds)
// come in through a shared-memory ring filled with cp.async, like prep_bwd_ring_kernel: four rows in flight per CTA
// independent of the registers (the warp-per-row form holds 64 accumulators per lane -- 182 registers, 10 warps per SM --
// and fetches ds only after the row reduction: 46 % of the HBM roofline in the train step).  One row lane per CTA of C/8
// threads, rows strided over the grid; a thread copies exactly the 16-byte pieces it reads later, so the ring needs no
// barrier; the row reduction is one __syncthreads per row (scratch double buffered by row parity); 16 accumulators per thread.
constexpr int kLnRing = 5;
__global__ void __launch_bounds__(256) add_ln_bwd_ring_kernel(const LnParams P) {
    extern __shared__ __align__(16) unsigned char ln_dyn[];
    const int C = P.C, c0 = threadIdx.x * kVec, lane = threadIdx.x & 31, wip = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    bf16 *ring = reinterpret_cast<bf16 *>(ln_dyn);                                          // [kLnRing][3][C]
    float *red = reinterpret_cast<float *>(ln_dyn + (size_t)kLnRing * 3 * C * sizeof(bf16));  // [2 parities][2][8 warps]
    const Row8 w = ld8f(P.w + c0);
    const bool has_ds = P.ds != nullptr;
    const float invC = 1.f / C;
    auto issue = [&](long it) {
        const long row = (long)blockIdx.x + it * gridDim.x;
        if (row < P.rows) {
            bf16 *dst = ring + (size_t)(it % kLnRing) * 3 * C + c0;
            tc05::cp_async16(dst, P.sum + row * C + c0);
            tc05::cp_async16(dst + C, P.dy + row * C + c0);
            if (has_ds) tc05::cp_async16(dst + 2 * (size_t)C, P.ds + row * C + c0);
        }
        tc05::cp_async_commit();
    };
    float acc[2][kVec];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int i = 0; i < kVec; i++) acc[s][i] = 0.f;
#pragma unroll
    for (int i = 0; i < kLnRing - 1; i++) issue(i);
    auto stats_of = [&](long row) {
        return row < P.rows ? *reinterpret_cast<const float2 *>(P.stats + 2 * row) : make_float2(0.f, 0.f);
    };
    float2 st_next = stats_of(blockIdx.x);
#pragma unroll 1
    for (long it = 0;; it++) {
        const long row = (long)blockIdx.x + it * gridDim.x;
        if (row >= P.rows) break;
        const float mu = st_next.x, rstd = st_next.y;
        st_next = stats_of(row + gridDim.x);
        tc05::cp_async_wait<kLnRing - 2>();                                  // this row's pieces (own) have landed
        const bf16 *slot = ring + (size_t)(it % kLnRing) * 3 * C + c0;
        const Row8 sm_ = ld8s(slot), dy = ld8s(slot + C);
        const Row8 d2 = has_ds ? ld8s(slot + 2 * (size_t)C) : zero8();
        issue(it + kLnRing - 1);                                             // into the slot read one iteration ago
        Row8 sh, g;
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; i++) {
            sh.v[i] = (sm_.v[i] - mu) * rstd;
            g.v[i] = dy.v[i] * w.v[i];
            acc[0][i] = fmaf(dy.v[i], sh.v[i], acc[0][i]);
            acc[1][i] += dy.v[i];
            m1 += g.v[i];
            m2 = fmaf(g.v[i], sh.v[i], m2);
        }
        m1 = warp_sum(m1); m2 = warp_sum(m2);
        float *rp = red + (it & 1) * 16;
        if (lane == 0) { rp[wip] = m1; rp[8 + wip] = m2; }
        __syncthreads();
        m1 = 0.f; m2 = 0.f;
        for (int j = 0; j < nwarp; j++) { m1 += rp[j]; m2 += rp[8 + j]; }
        m1 *= invC; m2 *= invC;
        Row8 o;
#pragma unroll
        for (int i = 0; i < kVec; i++) o.v[i] = rstd * (g.v[i] - m1 - sh.v[i] * m2) + d2.v[i];
        st8(P.dx + row * C + c0, o);
    }
    tc05::cp_async_wait<0>();
    float *dst = P.part + (size_t)blockIdx.x * 2 * C + c0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        *reinterpret_cast<float4 *>(dst + (size_t)s * C) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
        *reinterpret_cast<float4 *>(dst + (size_t)s * C + 4) = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
    }
}

// launch geometry: threads per row = C/8; rows per CTA so that the CTA has <= 512 threads; grid = multiple of the SM count
struct Geo { int threads, rows_per_cta, grid; size_t red_bytes(int n, int C) const { return (size_t)rows_per_cta * n * C * 4; } };
inline Geo geometry(int B, int T, int C, int max_rows, int ctas_per_sm, int max_threads = kMaxThreads) {
    Geo g;
    const int tpr = C / kVec;
    g.rows_per_cta = max_threads / tpr;
    if (g.rows_per_cta > max_rows) g.rows_per_cta = max_rows;
    if (g.rows_per_cta < 1) g.rows_per_cta = 1;
    g.threads = tpr * g.rows_per_cta;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long rows = (long)B * T;
    long need = (rows + g.rows_per_cta - 1) / g.rows_per_cta;
    long cap = (long)sms * ctas_per_sm;
    g.grid = (int)(need < cap ? need : cap);
    return g;
}

}  // namespace tmixf

using namespace tmixf;

static bool shape_ok(int B, int T, int C) { return B > 0 && T > 0 && C > 0 && C % kC == 0 && C / kVec <= kMaxThreads; }

int tmix_grid(int B, int T, int C, int which) {
    if (!shape_ok(B, T, C)) return 0;
    (void)which;
    return geometry(B, T, C, 4, 8, kBwdThreads).grid;       // partial rows the adjoints' scratch holds: up to 8 CTAs per SM
}

cudaError_t launch_add_ln_fwd(long rows, int C, const void *x, const void *res, const float *w, const float *b, float eps,
                              void *y, void *s, float *stats, cudaStream_t st) {
    LnParams P{};
    P.x = (const bf16 *)x; P.res = (const bf16 *)res; P.w = w; P.b = b; P.y = (bf16 *)y; P.s = (bf16 *)s; P.stats = stats;
    P.eps = eps; P.rows = rows; P.C = C;
    const Geo g = geometry(1, (int)rows, C, 4, 4);
    count_launch();
    static const int form = [] { const char *e = getenv("RWKVTTS_LN_FWD"); return e != nullptr ? atoi(e) : 0; }();
    if (form == 0 && C % 256 == 0 && C <= 1024 && rows >= 64) {
        // one warp per row, 8 rows per CTA, the CTAs that are resident together (73 registers at C = 1024: three per SM)
        const long need = (rows + 7) / 8, cap = 148L * (C <= 512 ? 4 : 3);
        const int grid = (int)(need < cap ? need : cap);
        switch (C / 256) {
            case 1: add_ln_fwd_warp_kernel<1><<<grid, 256, 0, st>>>(P); break;
            case 2: add_ln_fwd_warp_kernel<2><<<grid, 256, 0, st>>>(P); break;
            case 3: add_ln_fwd_warp_kernel<3><<<grid, 256, 0, st>>>(P); break;
            default: add_ln_fwd_warp_kernel<4><<<grid, 256, 0, st>>>(P); break;
        }
        return cudaGetLastError();
    }
    add_ln_fwd_kernel<<<g.grid, g.threads, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_add_ln_bwd(long rows, int C, const void *sum, const float *stats, const float *w, const void *dy,
                              const void *ds, void *dx, float *dparams /* [2][C] */, float *part, cudaStream_t st) {
    LnParams P{};
    P.sum = (const bf16 *)sum; P.stats = const_cast<float *>(stats); P.w = w; P.dy = (const bf16 *)dy; P.ds = (const bf16 *)ds;
    P.dx = (bf16 *)dx; P.part = part; P.rows = rows; P.C = C;
    const Geo g = geometry(1, (int)rows, C, 4, 4, kBwdThreads);
    count_launch(2);
    static const int form = [] { const char *e = getenv("RWKVTTS_LN_BWD"); return e != nullptr ? atoi(e) : 0; }();
    if (form == 0 && C % 256 == 0 && C <= 2048 && rows >= 64) {
        // rows through a cp.async ring in shared memory (see add_ln_bwd_ring_kernel); the grid stays within the scratch the
        // caller sized with rwkvtts_tmix_scratch_floats
        const size_t shm = (size_t)kLnRing * 3 * C * sizeof(bf16) + 2 * 16 * sizeof(float);
        int per_sm = (int)((227 * 1024) / (shm + 1024));                 // resident CTAs per SM by shared memory
        if (per_sm > 8) per_sm = 8;
        const int cap = geometry(1, (int)rows, C, 4, 8, kBwdThreads).grid;   // = tmix_grid(): rows of the caller's scratch
        int grid = 148 * per_sm;
        if (grid > cap) grid = cap;
        cudaError_t e2 = cudaFuncSetAttribute(add_ln_bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e2 != cudaSuccess) return e2;
        add_ln_bwd_ring_kernel<<<grid, C / kVec, shm, st>>>(P);
        reduce_partials_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(part, dparams, grid, 2 * C);
        return cudaGetLastError();
    }
    if (form != 2 && C % 256 == 0 && C <= 1024 && rows >= 8) {
        // one warp per row; the grid stays within the scratch the caller sized with rwkvtts_tmix_scratch_floats
        const int nv = C / 256;
        int grid = (int)((rows + 7) / 8);
        if (grid > g.grid) grid = g.grid;
        if (grid > 148 * (nv <= 3 ? 2 : 1)) grid = 148 * (nv <= 3 ? 2 : 1);       // resident CTAs per SM (registers)
        const size_t shw = (size_t)2 * C * sizeof(float);
        switch (nv) {
            case 1: add_ln_bwd_warp_kernel<1><<<grid, 256, shw, st>>>(P); break;
            case 2: add_ln_bwd_warp_kernel<2><<<grid, 256, shw, st>>>(P); break;
            case 3: add_ln_bwd_warp_kernel<3><<<grid, 256, shw, st>>>(P); break;
            default: add_ln_bwd_warp_kernel<4><<<grid, 320, shw, st>>>(P); break;   // 182 registers: 10 warps, 1 CTA per SM
        }
        reduce_partials_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(part, dparams, grid, 2 * C);
        return cudaGetLastError();
    }
    const size_t sh = g.red_bytes(2, C);
    cudaError_t e = cudaFuncSetAttribute(add_ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    add_ln_bwd_kernel<<<g.grid, g.threads, sh, st>>>(P);
    reduce_partials_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(part, dparams, g.grid, 2 * C);
    return cudaGetLastError();
}

cudaError_t launch_sqrelu(const void *x, const void *dy, void *out, long n, cudaStream_t st) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long n8 = n / kVec;
    long need = (n8 + 255) / 256;
    const int grid = (int)(need < (long)sms * 16 ? need : (long)sms * 16);
    count_launch();
    if (dy == nullptr) sqrelu_fwd_kernel<<<grid, 256, 0, st>>>((const bf16 *)x, (bf16 *)out, n8);
    else sqrelu_bwd_kernel<<<grid, 256, 0, st>>>((const bf16 *)x, (const bf16 *)dy, (bf16 *)out, n8);
    return cudaGetLastError();
}

cudaError_t launch_shift_mix_fwd(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                 const float *mix, void *const *out, void *prev_out, const unsigned char *first,
                                 cudaStream_t st) {
    MixParams P{};
    P.first = first;
    P.x = (const bf16 *)x; P.mask = (const bf16 *)mask; P.prev = (const bf16 *)prev; P.mix = mix;
    P.prev_out = (bf16 *)prev_out;
    for (int i = 0; i < n; i++) P.out[i] = (bf16 *)out[i];
    P.B = B; P.T = T; P.C = C; P.n = n;
    const Geo g = geometry(B, T, C, 4, 4);
    count_launch();
    if (n == 6) shift_mix_fwd_kernel<6><<<g.grid, g.threads, 0, st>>>(P);
    else if (n == 1) shift_mix_fwd_kernel<1><<<g.grid, g.threads, 0, st>>>(P);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_shift_mix_bwd(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                 const float *mix, const void *const *dout, void *dx, float *dmix, float *part,
                                 const unsigned char *first, cudaStream_t st) {
    MixParams P{};
    P.first = first;
    P.x = (const bf16 *)x; P.mask = (const bf16 *)mask; P.prev = (const bf16 *)prev; P.mix = mix;
    for (int i = 0; i < n; i++) P.dout[i] = (const bf16 *)dout[i];
    P.dx = (bf16 *)dx; P.part = part;
    P.B = B; P.T = T; P.C = C; P.n = n;
    const Geo g = geometry(B, T, C, 4, 4, kBwdThreads);
    const size_t sh = g.red_bytes(n, C);
    count_launch(2);
    cudaError_t e;
    static const int form = [] { const char *e = getenv("RWKVTTS_MIX_BWD"); return e != nullptr ? atoi(e) : 0; }();
    if (form != 4 && C % 128 == 0 && C >= 512 && C <= 2048 && (long)B * T >= 64 && (n == 6 || n == 1)) {
        // rows through a cp.async ring in shared memory (see shift_mix_bwd_ring_kernel); two CTAs per SM while they fit
        const size_t shm = (size_t)n * C * sizeof(float) + (size_t)kRing * (n + 1) * C * sizeof(bf16);
        int grid = 148 * (shm <= 110 * 1024 ? 2 : 1);
        if (grid > g.grid) grid = g.grid;             // the caller's scratch holds g.grid partial rows
        if (n == 6) {
            e = cudaFuncSetAttribute(shift_mix_bwd_ring_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
            if (e != cudaSuccess) return e;
            shift_mix_bwd_ring_kernel<6><<<grid, C / 4, shm, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(shift_mix_bwd_ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
            if (e != cudaSuccess) return e;
            shift_mix_bwd_ring_kernel<1><<<grid, C / 4, shm, st>>>(P);
        }
        reduce_partials_kernel<<<(n * C + 31) / 32, 256, 0, st>>>(part, dmix, grid, n * C);
        return cudaGetLastError();
    }
    if (C % 128 == 0 && C >= 512 && C <= 2048 && (long)B * T >= 64) {
        // four channels per thread, one row lane per CTA of C/4 threads (see shift_mix_bwd4_kernel)
        int grid = 148 * 3;
        if (grid > g.grid) grid = g.grid;             // the caller's scratch holds g.grid partial rows
        const size_t shm = (size_t)n * C * sizeof(float);
        if (n == 6) {
            e = cudaFuncSetAttribute(shift_mix_bwd4_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
            if (e != cudaSuccess) return e;
            shift_mix_bwd4_kernel<6><<<grid, C / 4, shm, st>>>(P);
        } else if (n == 1) {
            shift_mix_bwd4_kernel<1><<<grid, C / 4, shm, st>>>(P);
        } else return cudaErrorInvalidValue;
        reduce_partials_kernel<<<(n * C + 31) / 32, 256, 0, st>>>(part, dmix, grid, n * C);
        return cudaGetLastError();
    }
    if (n == 6) {
        e = cudaFuncSetAttribute(shift_mix_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return e;
        shift_mix_bwd_kernel<6><<<g.grid, g.threads, sh, st>>>(P);
    } else if (n == 1) {
        e = cudaFuncSetAttribute(shift_mix_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return e;
        shift_mix_bwd_kernel<1><<<g.grid, g.threads, sh, st>>>(P);
    } else return cudaErrorInvalidValue;
    reduce_partials_kernel<<<(n * C + 31) / 32, 256, 0, st>>>(part, dmix, g.grid, n * C);
    return cudaGetLastError();
}

cudaError_t launch_prep_fwd(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                            const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                            const float *v0, const float *k_k, const float *k_a, int mask_rwk, void *w, void *k2, void *v2,
                            void *a_op, void *b_op, cudaStream_t st) {
    PrepParams P{};
    P.k = (const bf16 *)k; P.v = (const bf16 *)v; P.w_lo = (const bf16 *)w_lo; P.a_lo = (const bf16 *)a_lo;
    P.v_lo = (const bf16 *)v_lo; P.v_first = (const bf16 *)v_first; P.mask = (const bf16 *)mask;
    P.w0 = w0; P.a0 = a0; P.v0 = v0; P.k_k = k_k; P.k_a = k_a;
    P.w = (bf16 *)w; P.k2 = (bf16 *)k2; P.v2 = (bf16 *)v2; P.a_op = (bf16 *)a_op; P.b_op = (bf16 *)b_op;
    P.B = B; P.T = T; P.C = C; P.mask_rwk = mask_rwk;
    const Geo g = geometry(B, T, C, 4, 4);
    count_launch();
    prep_fwd_kernel<<<g.grid, g.threads, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_prep_bwd(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                            const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                            const float *v0, const float *k_k, const float *k_a, int mask_rwk, const void *dw,
                            const void *dk2, const void *dv2, const void *da_op, const void *db_op, void *dk, void *dv,
                            void *dw_lo, void *da_lo, void *dv_lo, void *dv_first, float *dparams /* [5][C] */, float *part,
                            cudaStream_t st) {
    PrepParams P{};
    P.k = (const bf16 *)k; P.v = (const bf16 *)v; P.w_lo = (const bf16 *)w_lo; P.a_lo = (const bf16 *)a_lo;
    P.v_lo = (const bf16 *)v_lo; P.v_first = (const bf16 *)v_first; P.mask = (const bf16 *)mask;
    P.w0 = w0; P.a0 = a0; P.v0 = v0; P.k_k = k_k; P.k_a = k_a;
    P.dw = (const bf16 *)dw; P.dk2 = (const bf16 *)dk2; P.dv2 = (const bf16 *)dv2; P.da_op = (const bf16 *)da_op;
    P.db_op = (const bf16 *)db_op;
    P.dk = (bf16 *)dk; P.dv = (bf16 *)dv; P.dw_lo = (bf16 *)dw_lo; P.da_lo = (bf16 *)da_lo; P.dv_lo = (bf16 *)dv_lo;
    P.dv_first = (bf16 *)dv_first; P.part = part;
    P.B = B; P.T = T; P.C = C; P.mask_rwk = mask_rwk;
    const Geo g = geometry(B, T, C, 4, 4, kBwdThreads);
    const size_t sh = g.red_bytes(5, C);
    count_launch(2);
    static const int form = [] { const char *e = getenv("RWKVTTS_PREP_BWD"); return e != nullptr ? atoi(e) : 0; }();
    if (form != 1 && C % 256 == 0 && C <= 2048 && (long)B * T >= 64) {
        // rows through a cp.async ring in shared memory (see prep_bwd_ring_kernel)
        const size_t shm = (size_t)kPrepRing * kPrepArrays * C * sizeof(bf16);
        int grid = 148 * (shm <= 72 * 1024 ? 3 : 1);
        if (grid > g.grid) grid = g.grid;             // the caller's scratch holds g.grid partial rows
        cudaError_t e2 = cudaFuncSetAttribute(prep_bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e2 != cudaSuccess) return e2;
        prep_bwd_ring_kernel<<<grid, C / kVec, shm, st>>>(P);
        reduce_partials_kernel<<<(5 * C + 31) / 32, 256, 0, st>>>(part, dparams, grid, 5 * C);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(prep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    prep_bwd_kernel<<<g.grid, g.threads, sh, st>>>(P);
    reduce_partials_kernel<<<(5 * C + 31) / 32, 256, 0, st>>>(part, dparams, g.grid, 5 * C);
    return cudaGetLastError();
}

cudaError_t launch_out_fwd(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                           const void *g_, const float *r_k, const float *ln_w, const float *ln_b, float eps, void *o,
                           cudaStream_t st) {
    OutParams P{};
    P.y = (const bf16 *)y; P.r = (const bf16 *)r; P.k2 = (const bf16 *)k2; P.v2 = (const bf16 *)v2; P.g = (const bf16 *)g_;
    P.r_k = r_k; P.ln_w = ln_w; P.ln_b = ln_b; P.eps = eps; P.o = (bf16 *)o;
    P.B = B; P.T = T; P.C = C;
    const Geo g = geometry(B, T, C, 4, 4);
    count_launch();
    out_fwd_kernel<<<g.grid, g.threads, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_out_bwd(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                           const void *g_, const float *r_k, const float *ln_w, const float *ln_b, float eps,
                           const void *d_o, void *dy, void *dr, void *dk2, void *dv2, void *dg, float *dparams /* [3][C] */,
                           float *part, cudaStream_t st) {
    OutParams P{};
    P.y = (const bf16 *)y; P.r = (const bf16 *)r; P.k2 = (const bf16 *)k2; P.v2 = (const bf16 *)v2; P.g = (const bf16 *)g_;
    P.r_k = r_k; P.ln_w = ln_w; P.ln_b = ln_b; P.eps = eps;
    P.d_o = (const bf16 *)d_o; P.dy = (bf16 *)dy; P.dr = (bf16 *)dr; P.dk2 = (bf16 *)dk2; P.dv2 = (bf16 *)dv2;
    P.dg = (bf16 *)dg; P.part = part;
    P.B = B; P.T = T; P.C = C;
    const Geo g = geometry(B, T, C, 4, 4, kBwdThreads);
    const size_t sh = g.red_bytes(3, C);
    cudaError_t e = cudaFuncSetAttribute(out_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    count_launch(2);
    out_bwd_kernel<<<g.grid, g.threads, sh, st>>>(P);
    reduce_partials_kernel<<<(3 * C + 31) / 32, 256, 0, st>>>(part, dparams, g.grid, 3 * C);
    return cudaGetLastError();
}

}  // namespace rwkvtts
