// One autoregressive decode step of the whole RWKV-7 model in ONE persistent kernel  [HBM roofline: every weight and
// every recurrent state read once per token; SURVEY.md section 8 row f3].
//
// Reference: the decode loop of inference/rwkv7speech_inference.py / model/llm/spark_llm.py:54-102 runs, per token and
// layer, ~25 small kernels (token shift, six lerps, eleven skinny projections, the stateful WKV op
// rwkv7_state_fwd_fp16.cu with T = 1, GroupNorm, gate, output projection, channel mix: rwkv_asr_cuda_whisper.py:181-215,
// :277-285, :318-326).  At 32 rows each of them moves 0.1-8 MB, so the step is bound by launch count, not by bytes: the
// CUDA-graph step of this repo (rwkvfla/models/rwkv7/modeling_rwkv7.py::_GraphDecodeStep) needs 1.59 ms for ~430 nodes
// against 0.19 ms of HBM time for the 0.6 GB of weights + 0.4 GB of state it moves.
//
// Here: one cooperative launch of one CTA per SM (512 threads).  The step is a fixed sequence of phases separated by a
// grid barrier (7 per layer):
//   ln1    rows      x = residual (+ the previous layer's channel-mix output: its split-K partials are summed here),
//                    h = LayerNorm1(x), token shift and the six lerps -> X[6]            rwkv_s2s_single_ffn.py:160-169
//   gemm   columns   r, k, v projections and the four LoRA down-projections (tanh / sigmoid in the epilogue)   :170-184
//   wkv    (b, head) LoRA up-projections, decay / kk / a / k' / v' (:172-190), the WKV-7 state update and read-out
//                    (wkv7_cuda.cu:24-40 with T = 1), GroupNorm + bonus + gate (:192-195)
//   gemm   columns   output projection
//   ln2    rows      x2 = x + att, LayerNorm2, token shift, one lerp                                             :224-226
//   gemm   columns   channel-mix key projection, relu^2 in the epilogue                                             :228
//   gemm   columns   channel-mix value projection, split-K over 1024-wide chunks -> fp32 partials                   :230
// then the final norm, the vocabulary head, and (greedy) the arg-max with EOS handling on the device.
//
// Skinny GEMM (M = 32 rows): a unit is 8 output columns x the whole K; the 16 warps of a CTA split K, each keeps its slice
// of the activations as mma.sync A fragments in registers (loaded once per phase from L2) and streams the weight rows
// straight from HBM into B fragments -- W is [N, K] with K contiguous, which IS the "col" operand layout of
// mma.m16n8k16, so no shared-memory staging: each lane loads 16 contiguous bytes of one row (8 rows x 64 bytes per warp and
// load) and the k index inside a 32-wide block is permuted identically for A and B.  The 16 partial accumulators are
// reduced through shared memory; the next unit's weight loads are already in flight while that happens.
//
// Rounding points are those of the decode path it replaces (bf16 after every projection, LoRA stage, lerp, norm), so the
// two agree to the summation order of the projections; `generate(exact=True)` remains the bit-exact reference mode.
#include "tc05.cuh"
#include "wkv7_common.cuh"
#include "../../include/rwkvtts_wkv7.h"

#include <cstddef>
#include <type_traits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace rwkvtts {
namespace dec {

using tc05::clock_hi;

constexpr int kThreads = 512, kWarps = 16, kRows = 32, kMaxJobs = 8, kMaxEos = 8, kMaxLora = 512;
constexpr int kRedFloats = kWarps * 512;        // 32 KB: GEMM reduction (16 warps x 2 tiles x 256 partial sums) / wkv scratch
enum { kEpiBf16 = 0, kEpiTanh, kEpiSigmoid, kEpiSqRelu, kEpiF32, kEpiLogits };
enum { kRowLn1 = 0, kRowLn2, kRowFinal };

struct Job {
    const bf16 *A;      // [32, lda] activations (rows >= B are zero)
    const bf16 *W;      // [N, ldw], K contiguous
    void *out;          // [32, ldo] bf16 or fp32
    int lda, ldw, ldo, N, K, epi, tile0, pad_;
};
struct Phase {
    Job job[kMaxJobs];
    int njobs, tiles;
};
struct alignas(16) Layer {
    Phase p2, p4, p6, p7;
    const bf16 *ln1_w, *ln1_b, *mix[6];     // what the ln1 phase reads, contiguous with ...
    bf16 *att_shift;                        // ... its token-shift state [B, C], advanced in place (9 pointers)
    const bf16 *ln2_w, *ln2_b, *ffn_mix;
    bf16 *ffn_shift;                        // (4 pointers: the ln2 phase)
    const bf16 *up[4];                      // LoRA up-projections [C, D]: w, a, v (null on layer 0), g
    const bf16 *w0, *a0, *v0, *k_k, *k_a, *r_k, *gn_w, *gn_b;      // (8 pointers: per-channel vectors of the wkv phase)
    float *state;                           // [B, H, 64, 64] value-major, advanced in place
};
struct alignas(16) Desc {
    int B, C, H, L, V, F, D[4], Dtot, nchunk;
    int inv[4], pad_[2];            // ceil(2^20 / (D[i] / 8)): row of a 16-byte piece of a LoRA matrix without a division
    float ln_eps, gn_eps;
    const bf16 *emb, *ln0_w, *ln0_b, *lnf_w, *lnf_b;
    Phase head;
    unsigned *bar;                  // [0] arrivals of the grid barrier, [1] exits
    long long *tok;                 // [32] next input token (written by the arg-max phase)
    int *done;                      // [32]
    bf16 *x, *x2, *att, *o, *Xf, *hN, *vfirst, *X, *rkv, *hl, *kf;
    float *part, *logits;
    Layer *layers;
};
struct StepArgs {
    const long long *tok_in;        // [B] or null: read Desc::tok
    long long *tok_out;             // [B] or null
    long long eos[kMaxEos];
    long long pad;
    unsigned long long *prof;       // null, or the time line described at rwkvtts_decode_step_profile
    int n_eos, greedy, suppress_eos;
    int debug_skip;                 // tuning experiments only (env RWKVTTS_DECODE_SKIP): 1 skip the row phases, 2 skip wkv, 4 no L2 prefetch
};

// The kernel runs ~170 short phases per token, each through code the previous phases have pushed out of the 32 KB
// instruction cache: straight-line code that runs once per phase costs an L2 fetch per 8 instructions, which is what the
// first form of this kernel spent most of its time on (174 KB of code, 9 us per phase with a 1 us barrier).  So the phases
// are out-of-line functions with rolled loops (`#pragma unroll 1` wherever a trip count is 1-4 anyway), rare paths (tanh,
// sigmoid, arg-max, the watchdog) are kept out of the hot code, and loads are batched by hand where a rolled loop would
// serialise their latencies.

// ---- small device helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float neg_softplus_neg(float z) {
    const float y = -z;
    return (y > 20.f) ? -y : -__logf(1.f + __expf(y));
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// activations written by other CTAs earlier in this launch: L2 only (an L1 line could be stale); parameters: read-only path
__device__ __forceinline__ float2 ld_act2(const bf16 *p) {
    const unsigned u = __ldcg(reinterpret_cast<const unsigned *>(p));
    return make_float2(bf16_lo(u), bf16_hi(u));
}
__device__ __forceinline__ float2 ld_par2(const bf16 *p) {
    const unsigned u = __ldg(reinterpret_cast<const unsigned *>(p));
    return make_float2(bf16_lo(u), bf16_hi(u));
}
__device__ __forceinline__ void st2(bf16 *p, float a, float b) { *reinterpret_cast<uint32_t *>(p) = pack2(a, b); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// tuning aid: cycle stamps of CTA 0 / thread 0 at marked points of the phases (null = off)
#define PROF_POINT(fine, k) do { if ((fine) != nullptr && blockIdx.x == 0 && threadIdx.x == 0) (fine)[k] = clock64(); } while (0)
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
[[noreturn]] static __device__ __noinline__ void grid_die(unsigned epoch) {
    unsigned long long *r = tc05::g_wd_rec;
    if (r != nullptr) {
        r[8 + (epoch % tc05::kWatchdogEntries)] = (1ull << 63) | (3ull << 60) |
                                                  ((unsigned long long)(blockIdx.x & 0xffffffu) << 24) | (epoch & 0xffffffu);
        r[0] = tc05::kWatchdogMagic;
        __threadfence_system();
    }
    const long long t1 = clock64();
    while (clock64() - t1 < 100000000ll) {}
    __trap();
    for (;;) {}
}

// ---- L2 prefetch during the barrier waits ------------------------------------------------------------------------------
// What a later phase will read from HBM (weights, parameters, recurrent state: nothing of it depends on this token) is
// pulled into L2 one to two phases ahead, so that after a barrier only L2 latencies are left on the critical path.  One
// bulk-prefetch instruction per contiguous span, run by the copy engine; issuing one costs the warp ~100 cycles and a
// `prefetch.global.L2` per 128-byte line ~6 cycles of the SM's load pipe each (both measured), so the spans are issued by
// ONE warp while thread 0 polls the grid barrier -- the ~1 us every CTA waits there anyway.
__device__ __forceinline__ void prefetch_bulk(const void *p, unsigned bytes) {        // 16-byte aligned, bytes % 16 == 0
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
struct Ranges { int r[6][2]; };     // this CTA's tile range of: projections (layer 0), projections, out, key, value, head
enum { kRgP2First = 0, kRgP2, kRgP4, kRgP6, kRgP7, kRgHead };
enum { kPfNone = 0, kPfWkv, kPfOut, kPfKey, kPfValue, kPfNextProj, kPfHead };

// the weight rows of this CTA's tiles of a GEMM phase: one span per job segment where rows are contiguous (ldw == K),
// else the first 8 rows only (split-K jobs: every row is its own span)
__device__ __forceinline__ void prefetch_tiles(const Phase &P, int t0, int t1, int lane) {
    if (t0 >= t1) return;
    int j = 0;
    while (j + 1 < P.njobs && t0 >= P.job[j + 1].tile0) j++;
    int t = t0;
#pragma unroll 1
    while (t < t1) {
        const Job &J = P.job[j];
        const int jend = min(t1, J.tile0 + ((J.N + 7) >> 3)), n0 = (t - J.tile0) * 8, rows = min(J.N - n0, (jend - t) * 8);
        if (J.ldw == J.K) {
            if (lane == 0) prefetch_bulk(J.W + (size_t)n0 * J.ldw, (unsigned)(rows * J.K * 2));
        } else if (lane < min(rows, 8)) {
            prefetch_bulk(J.W + (size_t)(n0 + lane) * J.ldw, (unsigned)(J.K * 2));
        }
        t = jend;
        j++;
    }
}
__device__ __noinline__ void prefetch_plan(int plan, const Desc &D, const Layer &Ly, const Layer &Nx, const Ranges &R, bool first,
                                           int lane) {
    if (plan == kPfWkv) {
        // recurrent state of the four units of each round, and the head's LoRA up-projection rows
        const int NQ = D.H * ((D.B + 3) >> 2);
#pragma unroll 1
        for (int q = blockIdx.x; q < NQ; q += gridDim.x) {
            const int h = q % D.H, b0 = 4 * (q / D.H);
            if (lane < 4) {
                if (b0 + lane < D.B) prefetch_bulk(Ly.state + (size_t)((b0 + lane) * D.H + h) * kC * kC, kC * kC * sizeof(float));
            } else if (lane < 8) {
                const int L = lane - 4;
                if (Ly.up[L] != nullptr) prefetch_bulk(Ly.up[L] + (size_t)h * kC * D.D[L], (unsigned)(kC * D.D[L] * sizeof(bf16)));
            }
        }
    } else if (plan == kPfOut) {
        prefetch_tiles(Ly.p4, R.r[kRgP4][0], R.r[kRgP4][1], lane);
        if ((int)blockIdx.x < D.B && lane >= 8 && lane < 12) {               // ln2_w, ln2_b, ffn_mix, this row of ffn_shift
            const bf16 *q = (&Ly.ln2_w)[lane - 8];
            if (q != nullptr) prefetch_bulk(lane == 11 ? q + (size_t)blockIdx.x * D.C : q, (unsigned)(D.C * sizeof(bf16)));
        }
    } else if (plan == kPfKey) {
        prefetch_tiles(Ly.p6, R.r[kRgP6][0], R.r[kRgP6][1], lane);
    } else if (plan == kPfValue) {
        prefetch_tiles(Ly.p7, R.r[kRgP7][0], R.r[kRgP7][1], lane);
    } else if (plan == kPfNextProj) {
        prefetch_tiles(Nx.p2, R.r[kRgP2][0], R.r[kRgP2][1], lane);
        if ((int)blockIdx.x < D.B && lane >= 8 && lane < 17) {               // ln1_w, ln1_b, mix[6], this row of att_shift
            const bf16 *q = (&Nx.ln1_w)[lane - 8];
            if (q != nullptr) prefetch_bulk(lane == 16 ? q + (size_t)blockIdx.x * D.C : q, (unsigned)(D.C * sizeof(bf16)));
        }
    } else if (plan == kPfHead) {
        prefetch_tiles(D.head, R.r[kRgHead][0], R.r[kRgHead][1], lane);
        if ((int)blockIdx.x < D.B && lane >= 8 && lane < 10) {
            const bf16 *q = (&D.lnf_w)[lane - 8];
            if (q != nullptr) prefetch_bulk(q, (unsigned)(D.C * sizeof(bf16)));
        }
    }
    (void)first;
}

// Grid barrier on a monotonic arrival counter (cooperative launch: all CTAs are resident): release-add by one thread
// (cumulative over what the CTA wrote before the bar.sync), acquire-poll, ~1.0 us from the last arrival to the releases
// (measured; a per-CTA flag written by the last arriver takes 2.0 us, a poll with back-off the same 1.0).  Bounded like the
// mbarrier waits of the chunked kernels: a CTA that never arrives turns into a trap with a record, not a hung device.
constexpr int kProfCtas = 256;      // profile layout: [n] CTA 0 after each barrier, then [n][256] arrivals, [n][256] releases
struct SyncCtx {
    const Desc *D;
    const Ranges *R;
    unsigned long long *prof;
    int nprof, no_prefetch;
};
__device__ __noinline__ unsigned grid_sync(const SyncCtx &cx, unsigned epoch, int plan, const Layer *Ly, const Layer *Nx) {
    __syncthreads();
    epoch++;
    unsigned *bar = cx.D->bar;
    if (threadIdx.x == 0) {
        if (cx.prof != nullptr && blockIdx.x < kProfCtas) cx.prof[cx.nprof + epoch * kProfCtas + blockIdx.x] = globaltimer();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
    }
    if (plan != kPfNone && !cx.no_prefetch && (threadIdx.x >> 5) == 1) prefetch_plan(plan, *cx.D, *Ly, *Nx, *cx.R, false, threadIdx.x & 31);
    if (threadIdx.x == 0) {
        const unsigned target = epoch * gridDim.x;
        if (ld_acquire(bar) < target) {
            const uint32_t t0 = clock_hi();
            while (ld_acquire(bar) < target) {
                if (clock_hi() - t0 >= 2u) grid_die(epoch);
            }
        }
        if (cx.prof != nullptr) {
            const unsigned long long t = globaltimer();
            if (blockIdx.x == 0) cx.prof[epoch] = t;
            if (blockIdx.x < kProfCtas) cx.prof[cx.nprof * (1 + kProfCtas) + epoch * kProfCtas + blockIdx.x] = t;
        }
    }
    __syncthreads();
    return epoch;
}
__device__ __forceinline__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; i++) t += red[i];
    __syncthreads();
    return t;
}

// ---- row phases: ln1 (+ embedding / previous layer's channel-mix sum), ln2, final norm -------------------------------------------
// Each thread owns the channel pairs 2 tid (+ 1024): the row never leaves registers, and everything the phase reads that
// does not depend on the row itself (norm weights, lerp coefficients, shift state) is requested before the first reduction.
// PP = channel pairs per thread: 1 for C <= 1024, 2 for C <= 2048
template <int kPP>
__device__ __forceinline__ void ln_regs(float2 (&x)[kPP], int C, const float2 (&w)[kPP], const float2 (&b)[kPP], float eps, float *red) {
    float s1 = 0.f;
#pragma unroll
    for (int pp = 0; pp < kPP; pp++) s1 += x[pp].x + x[pp].y;           // pairs beyond C hold zeros
    const float mu = block_sum(s1, red) / (float)C;
    float s2 = 0.f;
#pragma unroll
    for (int pp = 0; pp < kPP; pp++)
        if (2 * (int)threadIdx.x + pp * 2 * kThreads < C) {
            const float d0 = x[pp].x - mu, d1 = x[pp].y - mu;
            s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
        }
    const float rstd = rsqrtf(block_sum(s2, red) / (float)C + eps);
#pragma unroll
    for (int pp = 0; pp < kPP; pp++) {
        x[pp].x = rbf((x[pp].x - mu) * rstd * w[pp].x + b[pp].x);
        x[pp].y = rbf((x[pp].y - mu) * rstd * w[pp].y + b[pp].y);
    }
}

template <int kPP>
__device__ __noinline__ void phase_rows(const Desc &D, const Layer &Ly, int kind, bool first_layer, const StepArgs &a, float *red,
                                        long long *fine) {
    const int C = D.C;
#pragma unroll 1
    for (int b = blockIdx.x; b < D.B; b += gridDim.x) {
        PROF_POINT(fine, 0);
        const size_t rb = (size_t)b * C;
        const bf16 *nw = kind == kRowLn1 ? Ly.ln1_w : kind == kRowLn2 ? Ly.ln2_w : D.lnf_w;
        const bf16 *nb = kind == kRowLn1 ? Ly.ln1_b : kind == kRowLn2 ? Ly.ln2_b : D.lnf_b;
        bf16 *shift = (kind == kRowLn1 ? Ly.att_shift : Ly.ffn_shift) + rb;
        const bf16 *const *mix = kind == kRowLn1 ? Ly.mix : &Ly.ffn_mix;
        const int n = kind == kRowLn1 ? 6 : kind == kRowLn2 ? 1 : 0;
        const bool emb = kind == kRowLn1 && first_layer;
        const bf16 *e = nullptr;
        if (emb) e = D.emb + (size_t)(a.tok_in != nullptr ? a.tok_in[b] : __ldcg(D.tok + b)) * C;
        float2 x[kPP], w[kPP], bb[kPP], w0[kPP], b0[kPP], sh[kPP], mx[kPP][6];
#pragma unroll
        for (int pp = 0; pp < kPP; pp++) {
            const int c = 2 * threadIdx.x + pp * 2 * kThreads;
            const bool ok = c < C;
            const float2 z = make_float2(0.f, 0.f);
            // the row
            if (!ok) {
                x[pp] = z;
            } else if (emb) {
                x[pp] = ld_par2(e + c);
            } else if (kind == kRowLn2) {
                const float2 u = ld_act2(D.x + rb + c), t = ld_act2(D.att + rb + c);
                x[pp] = make_float2(rbf(u.x + t.x), rbf(u.y + t.y));
            } else {
                // x2 + the channel-mix value projection: its split-K partials are summed here, as the bf16 tensor it returns
                const float2 u = ld_act2(D.x2 + rb + c);
                float2 pv[kMaxJobs];
#pragma unroll
                for (int kc = 0; kc < kMaxJobs; kc++)
                    pv[kc] = kc < D.nchunk ? __ldcg(reinterpret_cast<const float2 *>(D.part + ((size_t)kc * kRows + b) * C + c)) : z;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int kc = 0; kc < kMaxJobs; kc++) { s0 += pv[kc].x; s1 += pv[kc].y; }
                x[pp] = make_float2(rbf(u.x + rbf(s0)), rbf(u.y + rbf(s1)));
            }
            // everything else the phase reads
            w[pp] = ok ? ld_par2(nw + c) : z;
            bb[pp] = ok && nb != nullptr ? ld_par2(nb + c) : z;
            w0[pp] = ok && emb && D.ln0_w != nullptr ? ld_par2(D.ln0_w + c) : z;
            b0[pp] = ok && emb && D.ln0_b != nullptr ? ld_par2(D.ln0_b + c) : z;
            sh[pp] = ok && n > 0 ? ld_act2(shift + c) : z;
#pragma unroll
            for (int s = 0; s < 6; s++) mx[pp][s] = ok && s < n ? ld_par2(mix[s] + c) : z;
        }
        PROF_POINT(fine, 1);
        if (emb && D.ln0_w != nullptr) ln_regs<kPP>(x, C, w0, b0, D.ln_eps, red);
        bf16 *keep = (kind == kRowLn2 ? D.x2 : D.x) + rb;          // the residual stream
#pragma unroll
        for (int pp = 0; pp < kPP; pp++) {
            const int c = 2 * threadIdx.x + pp * 2 * kThreads;
            if (c < C) st2(keep + c, x[pp].x, x[pp].y);
        }
        PROF_POINT(fine, 2);
        ln_regs<kPP>(x, C, w, bb, D.ln_eps, red);
        PROF_POINT(fine, 3);
#pragma unroll
        for (int pp = 0; pp < kPP; pp++) {
            const int c = 2 * threadIdx.x + pp * 2 * kThreads;
            if (c >= C) continue;
            const float h0 = x[pp].x, h1 = x[pp].y;
            if (kind == kRowFinal) {
                st2(D.hN + rb + c, h0, h1);
            } else {
                bf16 *out = (kind == kRowLn1 ? D.X : D.Xf) + rb;
                const float xx0 = rbf(sh[pp].x - h0), xx1 = rbf(sh[pp].y - h1);      // the reference rounds shift(x) - x to bf16
#pragma unroll
                for (int s = 0; s < 6; s++)
                    if (s < n) st2(out + (size_t)s * kRows * C + c, fmaf(xx0, mx[pp][s].x, h0), fmaf(xx1, mx[pp][s].y, h1));
                st2(shift + c, h0, h1);
            }
        }
        PROF_POINT(fine, 4);
    }
}

__device__ __forceinline__ bool is_eos(const StepArgs &a, long long t) {
    bool e = false;
#pragma unroll
    for (int i = 0; i < kMaxEos; i++) e |= (i < a.n_eos && a.eos[i] == t);
    return e;
}

// greedy sampling on the device: arg-max (lowest index on ties), EOS ids masked while `suppress_eos`, finished rows
// padded, the finished flag updated (generate(): logits[:, eos] = -inf; where(done, pad, nxt); done |= isin(nxt, eos))
__device__ __noinline__ void phase_argmax(const Desc &D, const StepArgs &a, float *red) {
    int *redi = reinterpret_cast<int *>(red + kWarps);
    for (int b = blockIdx.x; b < D.B; b += gridDim.x) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int n = threadIdx.x; n < D.V; n += kThreads) {
            const float v = __ldcg(D.logits + (size_t)b * D.V + n);
            if (a.suppress_eos && is_eos(a, n)) continue;
            if (v > best || bi == 0x7fffffff) { best = v; bi = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; redi[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < kWarps; i++)
                if (red[i] > best || (red[i] == best && redi[i] < bi)) { best = red[i]; bi = redi[i]; }
            const int was_done = __ldcg(D.done + b);
            const long long t = was_done ? a.pad : (long long)bi;
            D.tok[b] = t;
            if (a.tok_out != nullptr) a.tok_out[b] = t;
            if (is_eos(a, t)) D.done[b] = 1;
        }
        __syncthreads();
    }
}

// ---- (b, head) phase: LoRA ups, decay / kk / a / k' / v', state update, GroupNorm + bonus + gate ---------------------------
// A CTA takes one head and FOUR batch rows per round (one per quarter of its threads), so the head's LoRA up-projection rows
// are brought into shared memory once for all four, the up-projections of the four rows are one set of mma tiles, and at
// batch 32 the 128 (head, four rows) units are ONE round on 148 CTAs (with two rows per round most CTAs ran two rounds,
// paying load issue, LoRA tiles, per-channel stage and epilogue twice: ~3 us per layer).
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory"); }
__device__ __forceinline__ float pair_sum(float x) { return x + __shfl_xor_sync(0xffffffffu, x, 1); }

// 64 rows x rank of every LoRA up-projection of head h (contiguous in each [C, rank] matrix) -> shared memory, rows padded
// by 16 bytes (conflict-free 16-byte reads by the (row, quarter) lanes), asynchronously: for the first round it is issued
// BEFORE the grid barrier that precedes the phase (the weights do not depend on the token)
__device__ __noinline__ void stage_up(const Desc &D, const Layer &Ly, int h, bf16 *upw, int tid, int nthreads) {
    int off = 0;
#pragma unroll 1
    for (int L = 0; L < 4; L++) {
        const int Dl = D.D[L], S = Dl >> 3, inv = D.inv[L];
        if (Ly.up[L] != nullptr) {
            const bf16 *src = Ly.up[L] + (size_t)h * kC * Dl;
#pragma unroll 1
            for (int idx = tid; idx < kC * S; idx += nthreads) {
                const int row = (int)(((unsigned)idx * (unsigned)inv) >> 20), seg = idx - row * S;
                tc05::cp_async16(upw + off + row * (Dl + 8) + seg * 8, src + row * Dl + seg * 8);
            }
        }
        off += kC * (Dl + 8);
    }
}

__device__ __noinline__ void phase_wkv(const Desc &D, const Layer &Ly, bool first_layer, float *smem, bf16 *upw, long long *fine) {
    const int grp = threadIdx.x >> 7, t = threadIdx.x & 127, i = t >> 1, p = t & 1, lane = t & 31;
    const int warp = threadIdx.x >> 5, g = (threadIdx.x & 31) >> 2, q4 = threadIdx.x & 3;
    const bool lead_warp = (t >> 5) == 0;                  // first warp of the group: the per-channel work
    // scratch: the four rows of LoRA hidden activations as bf16 [4][kMaxLora + 8], then per group los [4][64], vec [6][64], ys [64]
    constexpr int kHS = kMaxLora + 8, kHB = 2 * kHS;       // floats taken by the four bf16 rows
    bf16 *hbf = reinterpret_cast<bf16 *>(smem);
    float *los = smem + kHB + grp * 768, *vec = los + 4 * kC, *ys = vec + 6 * kC;
    const int C = D.C, H = D.H, NQ = H * ((D.B + 3) >> 2);
#pragma unroll 1
    for (int q = blockIdx.x; q < NQ; q += gridDim.x) {
        PROF_POINT(fine, 0);
        const int h = q % H, b = 4 * (q / H) + grp;
        const bool valid = b < D.B, lead = lead_warp && valid;
        const int u = b * H + h;
        // thread (i, p) owns value row i and the keys 8 m + 4 p .. + 3, m < 8
        float4 *srow = reinterpret_cast<float4 *>(Ly.state + ((size_t)u * kC + i) * kC + 4 * p);
        float4 s4[8];
        if (valid) {
#pragma unroll
            for (int m = 0; m < 8; m++) s4[m] = __ldcg(srow + 2 * m);          // the long-latency loads first
        }
        if (t * 8 < D.Dtot) {                                                  // this group's row of hl, 16 bytes per thread
            if (valid) tc05::cp_async16(hbf + grp * kHS + t * 8, D.hl + (size_t)b * D.Dtot + t * 8);
            else *reinterpret_cast<uint4 *>(hbf + grp * kHS + t * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
        // the lead warp's operands (two channels per lane), all in flight together
        const size_t at = (size_t)b * C + h * kC + 2 * lane;
        const int ch = h * kC + 2 * lane;
        float2 r2, k2, v2, vf, w0, a0, kk_, ka, v0, rk, gw, gb;
        if (lead) {
            r2 = ld_act2(D.rkv + at);
            k2 = ld_act2(D.rkv + (size_t)kRows * C + at);
            v2 = ld_act2(D.rkv + (size_t)2 * kRows * C + at);
            vf = first_layer ? make_float2(0.f, 0.f) : ld_act2(D.vfirst + at);
            w0 = ld_par2(Ly.w0 + ch); a0 = ld_par2(Ly.a0 + ch); kk_ = ld_par2(Ly.k_k + ch); ka = ld_par2(Ly.k_a + ch);
            v0 = first_layer ? make_float2(0.f, 0.f) : ld_par2(Ly.v0 + ch);
            rk = ld_par2(Ly.r_k + ch); gw = ld_par2(Ly.gn_w + ch);
            gb = Ly.gn_b != nullptr ? ld_par2(Ly.gn_b + ch) : make_float2(0.f, 0.f);
        }
        PROF_POINT(fine, 1);
        tc05::cp_async_commit();
        tc05::cp_async_wait<0>();                                            // the head's up-projection rows, the hl rows
        __syncthreads();
        PROF_POINT(fine, 2);
        // LoRA up-projections of the four rows on the tensor cores: D[16 x 8] = A[16 x K] B[K x 8] with rows 0..3 of A the
        // groups' hl rows (the other 12 are zero), B = 8 channels of the staged [64, rank] matrix (K contiguous: the
        // "col" operand as it lies).  Warp w: channels 8 (w & 7) .., LoRAs {w, a} (w < 8) or {v, g}.
        {
            const int nt = warp & 7, half = warp >> 3;
            int off_h = 0, off_w = 0;
#pragma unroll 1
            for (int L = 0; L < 4; L++) {
                const int Dl = D.D[L];
                if ((L >> 1) == half && Ly.up[L] != nullptr) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    const bf16 *ap = hbf + g * kHS + off_h + 2 * q4;                       // rows >= 4 are zero
                    const bf16 *bp = upw + off_w + (nt * 8 + g) * (Dl + 8) + 2 * q4;
#pragma unroll 1
                    for (int ks = 0; ks < Dl; ks += 16) {
                        uint32_t af[4];
                        af[0] = g < 4 ? *reinterpret_cast<const uint32_t *>(ap + ks) : 0u;
                        af[1] = 0u;
                        af[2] = g < 4 ? *reinterpret_cast<const uint32_t *>(ap + ks + 8) : 0u;
                        af[3] = 0u;
                        mma_bf16(acc, af, *reinterpret_cast<const uint32_t *>(bp + ks), *reinterpret_cast<const uint32_t *>(bp + ks + 8));
                    }
                    if (g < 4) {
                        float *dst = smem + kHB + g * 768 + L * kC + nt * 8 + 2 * q4;          // los of group g
                        *reinterpret_cast<float2 *>(dst) = make_float2(rbf(acc[0]), rbf(acc[1]));
                    }
                }
                off_h += Dl;
                off_w += kC * (Dl + 8);
            }
        }
        __syncthreads();                                                     // los complete; everybody is done with the staged rows
        PROF_POINT(fine, 3);
        // a next round's rows (more than grid quads: batch > 36 at 16 heads), issued by the 12 warps that now wait for the leads
        if (q + (int)gridDim.x < NQ && !lead_warp) stage_up(D, Ly, (q + gridDim.x) % H, upw, (int)threadIdx.x - 32 * (grp + 1), kThreads - 128);
        tc05::cp_async_commit();
        PROF_POINT(fine, 4);
        float g0 = 0.f, g1 = 0.f;
        if (lead) {
            const int c0 = 2 * lane, c1 = c0 + 1;
            const bool hv = Ly.up[2] != nullptr;
            g0 = los[3 * kC + c0]; g1 = los[3 * kC + c1];
            const float wa = rbf(neg_softplus_neg(w0.x + los[c0]) - 0.5f), wb = rbf(neg_softplus_neg(w0.y + los[c1]) - 0.5f);
            vec[c0] = expf(-expf(wa)); vec[c1] = expf(-expf(wb));                          // decay (wkv7_cuda.cu:24)
            const float aa = rbf(sigmoidf_(a0.x + los[kC + c0])), ab = rbf(sigmoidf_(a0.y + los[kC + c1]));
            const float ua = k2.x * kk_.x, ub = k2.y * kk_.y;
            const float inv = 1.f / fmaxf(sqrtf(warp_sum(fmaf(ua, ua, ub * ub))), 1e-12f);
            const float kka = rbf(ua * inv), kkb = rbf(ub * inv);
            if (first_layer || !hv) {
                if (first_layer) st2(D.vfirst + at, v2.x, v2.y);
            } else {
                v2.x = rbf(v2.x + (vf.x - v2.x) * sigmoidf_(v0.x + los[2 * kC + c0]));
                v2.y = rbf(v2.y + (vf.y - v2.y) * sigmoidf_(v0.y + los[2 * kC + c1]));
            }
            k2.x = rbf(k2.x * (1.f + (aa - 1.f) * ka.x));
            k2.y = rbf(k2.y * (1.f + (ab - 1.f) * ka.y));
            *reinterpret_cast<float2 *>(vec + 1 * kC + c0) = r2;
            *reinterpret_cast<float2 *>(vec + 2 * kC + c0) = k2;
            *reinterpret_cast<float2 *>(vec + 3 * kC + c0) = v2;
            *reinterpret_cast<float2 *>(vec + 4 * kC + c0) = make_float2(-kka, -kkb);
            *reinterpret_cast<float2 *>(vec + 5 * kC + c0) = make_float2(rbf(kka * aa), rbf(kkb * ab));
        }
        group_bar(grp);
        PROF_POINT(fine, 5);
        if (valid) {
            // sa_i = sum_j S_ij a_j ; S_ij <- S_ij d_j + sa_i b_j + k_j v_i ; y_i = sum_j S_ij q_j (wkv7_cuda.cu:27-40, T = 1);
            // the six 64-vectors are read as 16-byte pieces
            float sa = 0.f;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const float4 av = *reinterpret_cast<const float4 *>(vec + 4 * kC + 8 * m + 4 * p);
                sa = fmaf(s4[m].x, av.x, sa); sa = fmaf(s4[m].y, av.y, sa); sa = fmaf(s4[m].z, av.z, sa); sa = fmaf(s4[m].w, av.w, sa);
            }
            sa = pair_sum(sa);
            const float vi = vec[3 * kC + i];
            float yy = 0.f;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int cc = 8 * m + 4 * p;
                const float4 dv = *reinterpret_cast<const float4 *>(vec + cc), qv = *reinterpret_cast<const float4 *>(vec + 1 * kC + cc),
                             kv = *reinterpret_cast<const float4 *>(vec + 2 * kC + cc), bv = *reinterpret_cast<const float4 *>(vec + 5 * kC + cc);
                float4 S = s4[m];
                S.x = fmaf(S.x, dv.x, fmaf(sa, bv.x, kv.x * vi)); yy = fmaf(S.x, qv.x, yy);
                S.y = fmaf(S.y, dv.y, fmaf(sa, bv.y, kv.y * vi)); yy = fmaf(S.y, qv.y, yy);
                S.z = fmaf(S.z, dv.z, fmaf(sa, bv.z, kv.z * vi)); yy = fmaf(S.z, qv.z, yy);
                S.w = fmaf(S.w, dv.w, fmaf(sa, bv.w, kv.w * vi)); yy = fmaf(S.w, qv.w, yy);
                s4[m] = S;
            }
            yy = pair_sum(yy);
            if (p == 0) ys[i] = rbf(yy);
        }
        group_bar(grp);
        PROF_POINT(fine, 6);
        if (valid) {                                                         // the new state leaves behind the barrier
#pragma unroll
            for (int m = 0; m < 8; m++) __stcg(srow + 2 * m, s4[m]);
        }
        if (lead) {
            const float2 y = *reinterpret_cast<const float2 *>(ys + 2 * lane);
            const float mu = warp_sum(y.x + y.y) * (1.f / kC);
            const float d0 = y.x - mu, d1 = y.y - mu;
            const float rstd = rsqrtf(warp_sum(fmaf(d0, d0, d1 * d1)) * (1.f / kC) + D.gn_eps);
            const float sb = warp_sum(fmaf(r2.x * k2.x, rk.x, r2.y * k2.y * rk.y));
            st2(D.o + at, (rbf(d0 * rstd * gw.x + gb.x) + sb * v2.x) * g0, (rbf(d1 * rstd * gw.y + gb.y) + sb * v2.y) * g1);
        }
        PROF_POINT(fine, 7);
        if (fine != nullptr) fine += 8;                                       // next round: next 8 slots
    }
}

// ---- skinny GEMM phase ---------------------------------------------------------------------------------------------------
// weight fragments of up to NT 8-column tiles: wp = this lane's row of the first tile at its k offset; rows_left /
// k_left bound what exists (last tile of an odd vocabulary, K smaller than the 16 warps' slices)
template <int NKB, int NT>
__device__ __forceinline__ void load_w(const bf16 *wp, int ldw, int ntile, int rows_left, int k_left, uint4 (&wv)[NT][NKB]) {
#pragma unroll
    for (int tt = 0; tt < NT; tt++)
#pragma unroll
        for (int kb = 0; kb < NKB; kb++) {
            const bool ok = tt < ntile && tt * 8 < rows_left && kb * 32 < k_left;
            wv[tt][kb] = ok ? ldg_nc_v4(wp + (size_t)tt * 8 * ldw + kb * 32) : make_uint4(0u, 0u, 0u, 0u);
        }
}
// the LoRA down-projections' activations: rare tiles, kept out of the hot code
__device__ __noinline__ float epi_act(float v, int epi) {
    return epi == kEpiTanh ? tanhf(v) : 1.f / (1.f + expf(-v));
}

// NT = 8-column tiles per pass.  Four per pass (one pass instead of two for the 2.8 - 3.5 tiles a CTA has of the projections
// and the channel mix) was measured slower: a pass costs instructions in proportion to its tiles, and the phases are bound
// by instruction issue, not by the number of passes.
template <int NKB, int NT /* = 2 */>
__device__ __noinline__ void phase_gemm(const Phase &P, int t0, int t1, int B, float *red, long long *fine) {
    if (t0 >= t1) return;
    PROF_POINT(fine, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int koff = warp * NKB * 32 + q * 8;              // this lane's first k of its warp's slice
    int j = 0;
    while (j + 1 < P.njobs && t0 >= P.job[j + 1].tile0) j++;
    int t = t0;
#pragma unroll 1
    while (t < t1) {                                        // one job (projection) at a time: its fields live in registers
        const Job J = P.job[j];
        const int jend = min(t1, J.tile0 + ((J.N + 7) >> 3)), k_left = J.K - koff;
        int n0 = (t - J.tile0) * 8;
        const bf16 *wp = J.W + (size_t)(n0 + g) * J.ldw + koff;
        uint4 wv[NT][NKB];
        load_w<NKB, NT>(wp, J.ldw, jend - t, J.N - n0 - g, k_left, wv);
        uint32_t af[NKB][2][2][4];  // [k block][m tile][mma of the block][a0..a3], already in the instruction's register order
#pragma unroll
        for (int kb = 0; kb < NKB; kb++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;              // rows g and g + 8 of the m tile
                if (kb * 32 < k_left) {
                    const bf16 *src = J.A + (size_t)(mt * 16 + g) * J.lda + koff + kb * 32;
                    lo = __ldcg(reinterpret_cast<const uint4 *>(src));
                    hi = __ldcg(reinterpret_cast<const uint4 *>(src + (size_t)8 * J.lda));
                }
                // physical k = 8q + 4j + {0,1} <-> logical k = 2q + {0,1}; 8q + 4j + {2,3} <-> 2q + 8 + {0,1} (mma j = 0, 1)
                af[kb][mt][0][0] = lo.x; af[kb][mt][0][1] = hi.x; af[kb][mt][0][2] = lo.y; af[kb][mt][0][3] = hi.y;
                af[kb][mt][1][0] = lo.z; af[kb][mt][1][1] = hi.z; af[kb][mt][1][2] = lo.w; af[kb][mt][1][3] = hi.w;
            }
        // one pass of mma -> reduce -> store over TT (compile time) tiles: a single tile (the output projection's one
        // tile per CTA, the odd tile at the end of a range) runs half the instructions of a pair
        auto pass = [&](auto tt_const) {
            constexpr int TT = decltype(tt_const)::value;
            PROF_POINT(fine, 1);
            float acc[TT][2][4];
#pragma unroll
            for (int tt = 0; tt < TT; tt++)
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[tt][mt][e] = 0.f;
#pragma unroll
            for (int tt = 0; tt < TT; tt++)
#pragma unroll
                for (int kb = 0; kb < NKB; kb++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        mma_bf16(acc[tt][mt], af[kb][mt][0], wv[tt][kb].x, wv[tt][kb].y);
                        mma_bf16(acc[tt][mt], af[kb][mt][1], wv[tt][kb].z, wv[tt][kb].w);
                    }
            PROF_POINT(fine, 2);
            wp += (size_t)NT * 8 * J.ldw;
            if (t + NT < jend) load_w<NKB, NT>(wp, J.ldw, jend - t - NT, J.N - n0 - NT * 8 - g, k_left, wv);   // in flight during the reduction
#pragma unroll
            for (int tt = 0; tt < TT; tt++)
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
                    *reinterpret_cast<float4 *>(red + ((warp * NT + tt) * 2 + mt) * 128 + lane * 4) =
                        make_float4(acc[tt][mt][0], acc[tt][mt][1], acc[tt][mt][2], acc[tt][mt][3]);
            __syncthreads();
            PROF_POINT(fine, 3);
            // thread -> output (tile tid / 256, element tid % 256) of a pair; a single tile: each half of the CTA sums the
            // partial tiles of 8 warps, the halves meet in shared memory
            const int hf = threadIdx.x >> 8, idx = threadIdx.x & 255;
            float x = 0.f;
            if (TT == 2) {
#pragma unroll
                for (int w = 0; w < kWarps; w++) x += red[(w * NT + hf) * 256 + idx];
            } else {
#pragma unroll
                for (int w = 0; w < kWarps / 2; w++) x += red[((w + 8 * hf) * NT) * 256 + idx];
            }
            __syncthreads();                     // the reduction buffer is free; the stores below have no barrier behind them
            if (TT == 1) {
                if (hf == 1) red[idx] = x;
                __syncthreads();
                if (hf == 0) x += red[idx];
                __syncthreads();
            }
            const int mt = idx >> 7, ln = (idx >> 2) & 31, e = idx & 3;
            const int row = mt * 16 + (ln >> 2) + ((e >> 1) << 3), n = n0 + (TT == 2 ? hf : 0) * 8 + (ln & 3) * 2 + (e & 1);
            if ((TT == 2 || hf == 0) && row < B && n < J.N) {
                const size_t at = (size_t)row * J.ldo + n;
                if (J.epi == kEpiF32) {
                    static_cast<float *>(J.out)[at] = x;
                } else if (J.epi == kEpiLogits) {
                    static_cast<float *>(J.out)[at] = rbf(x);
                } else {
                    float v = rbf(x);
                    if (J.epi == kEpiSqRelu) { v = fmaxf(v, 0.f); v = v * v; }
                    else if (J.epi != kEpiBf16) v = epi_act(v, J.epi);
                    static_cast<bf16 *>(J.out)[at] = __float2bfloat16_rn(v);
                }
            }
            PROF_POINT(fine, 4);
            if (fine != nullptr) fine += 4;
            t += TT;
            n0 += NT * 8;
        };
#pragma unroll 1
        while (t < jend) {
            if (jend - t >= 2) pass(std::integral_constant<int, 2>{});
            else pass(std::integral_constant<int, 1>{});
        }
        j++;
    }
}

__device__ __forceinline__ void copy_async(void *smem_dst, const void *gsrc, int bytes) {      // bytes % 16 == 0
#pragma unroll 1
    for (int i = threadIdx.x * 16; i < bytes; i += kThreads * 16)
        tc05::cp_async16(static_cast<char *>(smem_dst) + i, static_cast<const char *>(gsrc) + i);
}

template <int NKB>
__global__ void __launch_bounds__(kThreads, 1) decode_step_kernel(const Desc *__restrict__ Dp, const StepArgs a) {
    extern __shared__ __align__(16) float smem[];      // [kRedFloats] scratch of the running phase, then the staged LoRA rows
    __shared__ Desc sD;             // the plan lives in shared memory: a descriptor read from HBM at the head of every
    __shared__ Layer sL[2];         // phase would put a DRAM round trip in front of each of the 170 phases
    __shared__ Ranges sR;
    static_assert(sizeof(Desc) % 16 == 0 && sizeof(Layer) % 16 == 0, "cp.async pieces");
    static_assert(offsetof(Layer, att_shift) - offsetof(Layer, ln1_w) == 8 * sizeof(void *), "ln1 pointer group");
    static_assert(offsetof(Layer, ffn_shift) - offsetof(Layer, ln2_w) == 3 * sizeof(void *), "ln2 pointer group");
    static_assert(offsetof(Desc, lnf_b) - offsetof(Desc, lnf_w) == sizeof(void *), "final norm pointer group");
    copy_async(&sD, Dp, sizeof(Desc));
    tc05::cp_async_commit();
    tc05::cp_async_wait<0>();
    __syncthreads();
    const Desc &D = sD;
    copy_async(&sL[0], D.layers, sizeof(Layer));
    if (D.L > 1) copy_async(&sL[1], D.layers + 1, sizeof(Layer));
    tc05::cp_async_commit();
    tc05::cp_async_wait<0>();
    __syncthreads();
    if (threadIdx.x < 6) {          // this CTA's share of the tiles of each kind of GEMM phase (the same in every layer > 0)
        const int k = threadIdx.x;
        const Phase &P = k == kRgP2First ? sL[0].p2 : k == kRgP2 ? sL[D.L > 1 ? 1 : 0].p2 : k == kRgP4 ? sL[0].p4
                       : k == kRgP6 ? sL[0].p6 : k == kRgP7 ? sL[0].p7 : D.head;
        sR.r[k][0] = (int)((unsigned)blockIdx.x * (unsigned)P.tiles / gridDim.x);
        sR.r[k][1] = (int)((unsigned)(blockIdx.x + 1) * (unsigned)P.tiles / gridDim.x);
    }
    __syncthreads();
    float *red_rows = smem + kRedFloats - 64;   // reductions of the row phases
    bf16 *upw = reinterpret_cast<bf16 *>(smem + kRedFloats);
    const int nprof = 8 * D.L + 4;
    const SyncCtx cx{&sD, &sR, a.prof, nprof, a.debug_skip & 4};
    // fine stamps of layer 1 (cycles): [16..] projections, [32..] wkv rounds, [48..] output projection, [64..] key
    long long *fine = a.prof != nullptr ? reinterpret_cast<long long *>(a.prof + (size_t)nprof * (1 + 2 * kProfCtas)) : nullptr;
    unsigned epoch = 0;
    if (a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.prof[0] = globaltimer();
    if ((threadIdx.x >> 5) == 1) {
        prefetch_tiles(sL[0].p2, sR.r[kRgP2First][0], sR.r[kRgP2First][1], threadIdx.x & 31);
    }
    const int NQ = D.H * ((D.B + 3) >> 2);
#pragma unroll 1
    for (int l = 0; l < D.L; l++) {
        const Layer &Ly = sL[l & 1], &Nx = sL[(l + 1) & 1];
        const bool f1 = l == 1 && fine != nullptr, last = l + 1 == D.L;
        const int rp2 = l == 0 ? kRgP2First : kRgP2;
        if (!(a.debug_skip & 1)) phase_rows<NKB / 2>(D, Ly, kRowLn1, l == 0, a, red_rows, f1 ? fine + 80 : nullptr);
        epoch = grid_sync(cx, epoch, kPfWkv, &Ly, &Nx);
        phase_gemm<NKB, 2>(Ly.p2, sR.r[rp2][0], sR.r[rp2][1], D.B, smem, f1 ? fine + 16 : nullptr);
        if (!(a.debug_skip & 2) && (int)blockIdx.x < NQ) stage_up(D, Ly, blockIdx.x % D.H, upw, threadIdx.x, kThreads);      // first round of the wkv phase
        tc05::cp_async_commit();
        epoch = grid_sync(cx, epoch, kPfOut, &Ly, &Nx);
        if (!(a.debug_skip & 2)) phase_wkv(D, Ly, l == 0, smem, upw, f1 ? fine + 32 : nullptr);
        epoch = grid_sync(cx, epoch, kPfKey, &Ly, &Nx);
        phase_gemm<NKB, 2>(Ly.p4, sR.r[kRgP4][0], sR.r[kRgP4][1], D.B, smem, f1 ? fine + 48 : nullptr);
        epoch = grid_sync(cx, epoch, kPfValue, &Ly, &Nx);
        if (!(a.debug_skip & 1)) phase_rows<NKB / 2>(D, Ly, kRowLn2, false, a, red_rows, f1 ? fine + 88 : nullptr);
        tc05::cp_async_wait<0>();                        // own pieces of layer l + 1's descriptor; the barrier publishes them
        epoch = grid_sync(cx, epoch, kPfNone, &Ly, &Nx);
        phase_gemm<NKB, 2>(Ly.p6, sR.r[kRgP6][0], sR.r[kRgP6][1], D.B, smem, f1 ? fine + 64 : nullptr);
        epoch = grid_sync(cx, epoch, last ? kPfHead : kPfNextProj, &Ly, &Nx);
        phase_gemm<NKB, 2>(Ly.p7, sR.r[kRgP7][0], sR.r[kRgP7][1], D.B, smem, nullptr);
        epoch = grid_sync(cx, epoch, kPfNone, &Ly, &Nx);
        // every thread is past its last use of sL[l & 1]: bring in layer l + 2
        if (l + 2 < D.L) copy_async(&sL[l & 1], D.layers + l + 2, sizeof(Layer));
        tc05::cp_async_commit();
    }
    phase_rows<NKB / 2>(D, sL[0], kRowFinal, false, a, red_rows, nullptr);
    epoch = grid_sync(cx, epoch, kPfNone, &sL[0], &sL[0]);
    phase_gemm<NKB, 2>(D.head, sR.r[kRgHead][0], sR.r[kRgHead][1], D.B, smem, nullptr);
    if (a.greedy) {
        epoch = grid_sync(cx, epoch, kPfNone, &sL[0], &sL[0]);
        phase_argmax(D, a, red_rows);
    }
    // the last CTA out resets the counters for the next launch (every CTA has left its last barrier by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(D.bar + 1, 1u) == gridDim.x - 1) {
            D.bar[0] = 0u;
            D.bar[1] = 0u;
            __threadfence();
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
struct HostInfo { int grid, nkb, device; size_t smem; };
static std::mutex g_mu;
static std::unordered_map<void *, HostInfo> g_plans;

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
struct Layout {
    size_t desc, layers, bar, tok, done, x, x2, att, o, Xf, hN, vfirst, X, rkv, hl, kf, part, logits, total;
};
static Layout layout(const int *d) {
    const size_t B32 = kRows, C = d[RWKVTTS_DEC_C], L = d[RWKVTTS_DEC_L], V = d[RWKVTTS_DEC_V], F = d[RWKVTTS_DEC_F];
    const size_t Dtot = (size_t)d[RWKVTTS_DEC_DW] + d[RWKVTTS_DEC_DA] + d[RWKVTTS_DEC_DV] + d[RWKVTTS_DEC_DG];
    Layout o{};
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t r = at; at = align_up(at + bytes); return r; };
    o.desc = take(sizeof(Desc));
    o.layers = take(sizeof(Layer) * L);
    o.bar = take(256);
    o.tok = take(sizeof(long long) * B32);
    o.done = take(sizeof(int) * B32);
    const size_t row = B32 * C * sizeof(bf16);
    o.x = take(row); o.x2 = take(row); o.att = take(row); o.o = take(row); o.Xf = take(row); o.hN = take(row);
    o.vfirst = take(row);
    o.X = take(6 * row);
    o.rkv = take(3 * row);
    o.hl = take(B32 * Dtot * sizeof(bf16));
    o.kf = take(B32 * F * sizeof(bf16));
    o.part = take((F / C) * B32 * C * sizeof(float));
    o.logits = take(B32 * V * sizeof(float));
    o.total = at;
    return o;
}
static bool dims_ok(const int *d) {
    const int B = d[RWKVTTS_DEC_B], C = d[RWKVTTS_DEC_C], H = d[RWKVTTS_DEC_H], L = d[RWKVTTS_DEC_L], V = d[RWKVTTS_DEC_V],
              F = d[RWKVTTS_DEC_F];
    if (B < 1 || B > kRows || H < 1 || C != H * kC || C % 32 != 0 || C > 2048 || L < 1 || V < 1 || F < C || F % C != 0 ||
        F / C > kMaxJobs)
        return false;
    int tot = 0;
    for (int i = RWKVTTS_DEC_DW; i <= RWKVTTS_DEC_DG; i++) {
        if (d[i] < 32 || d[i] % 32 != 0) return false;
        tot += d[i];
    }
    return tot <= kMaxLora;
}
static void add_job(Phase &P, const void *A, int lda, const void *W, int ldw, void *out, int ldo, int N, int K, int epi) {
    Job &J = P.job[P.njobs++];
    J.A = static_cast<const bf16 *>(A); J.W = static_cast<const bf16 *>(W); J.out = out;
    J.lda = lda; J.ldw = ldw; J.ldo = ldo; J.N = N; J.K = K; J.epi = epi; J.tile0 = P.tiles; J.pad_ = 0;
    P.tiles += (N + 7) / 8;
}

cudaError_t decode_init(const int *d, const float *eps, const void *const *mp, const void *const *lp, void *ws, cudaStream_t st) {
    const Layout lo = layout(d);
    const int B = d[RWKVTTS_DEC_B], C = d[RWKVTTS_DEC_C], L = d[RWKVTTS_DEC_L], V = d[RWKVTTS_DEC_V], F = d[RWKVTTS_DEC_F];
    (void)B;
    char *base = static_cast<char *>(ws);
    std::vector<char> host(lo.bar, 0);
    Desc &D = *reinterpret_cast<Desc *>(host.data() + lo.desc);
    Layer *Ls = reinterpret_cast<Layer *>(host.data() + lo.layers);
    D.B = d[RWKVTTS_DEC_B]; D.C = C; D.H = d[RWKVTTS_DEC_H]; D.L = L; D.V = V; D.F = F;
    D.Dtot = 0;
    for (int i = 0; i < 4; i++) { D.D[i] = d[RWKVTTS_DEC_DW + i]; D.Dtot += D.D[i]; }
    D.nchunk = F / C;
    for (int i = 0; i < 4; i++) D.inv[i] = ((1 << 20) + D.D[i] / 8 - 1) / (D.D[i] / 8);
    D.ln_eps = eps[0]; D.gn_eps = eps[1];
    auto bp = [](const void *p) { return static_cast<const bf16 *>(p); };
    D.emb = bp(mp[RWKVTTS_DEC_EMB]); D.ln0_w = bp(mp[RWKVTTS_DEC_LN0_W]); D.ln0_b = bp(mp[RWKVTTS_DEC_LN0_B]);
    D.lnf_w = bp(mp[RWKVTTS_DEC_LNF_W]); D.lnf_b = bp(mp[RWKVTTS_DEC_LNF_B]);
    auto dev = [&](size_t off) { return static_cast<void *>(base + off); };
    D.bar = static_cast<unsigned *>(dev(lo.bar));
    D.tok = static_cast<long long *>(dev(lo.tok));
    D.done = static_cast<int *>(dev(lo.done));
    D.x = (bf16 *)dev(lo.x); D.x2 = (bf16 *)dev(lo.x2); D.att = (bf16 *)dev(lo.att); D.o = (bf16 *)dev(lo.o);
    D.Xf = (bf16 *)dev(lo.Xf); D.hN = (bf16 *)dev(lo.hN); D.vfirst = (bf16 *)dev(lo.vfirst); D.X = (bf16 *)dev(lo.X);
    D.rkv = (bf16 *)dev(lo.rkv); D.hl = (bf16 *)dev(lo.hl); D.kf = (bf16 *)dev(lo.kf);
    D.part = (float *)dev(lo.part); D.logits = (float *)dev(lo.logits);
    D.layers = static_cast<Layer *>(dev(lo.layers));
    const size_t RC = (size_t)kRows * C;
    for (int l = 0; l < L; l++) {
        const void *const *p = lp + (size_t)l * RWKVTTS_DEC_NPTR;
        Layer &Y = Ls[l];
        Y.ln1_w = bp(p[RWKVTTS_DEC_LN1_W]); Y.ln1_b = bp(p[RWKVTTS_DEC_LN1_B]);
        Y.ln2_w = bp(p[RWKVTTS_DEC_LN2_W]); Y.ln2_b = bp(p[RWKVTTS_DEC_LN2_B]);
        for (int s = 0; s < 6; s++) Y.mix[s] = bp(p[RWKVTTS_DEC_X_R + s]);
        Y.ffn_mix = bp(p[RWKVTTS_DEC_FFN_X_K]);
        Y.up[0] = bp(p[RWKVTTS_DEC_W2]); Y.up[1] = bp(p[RWKVTTS_DEC_A2]); Y.up[2] = bp(p[RWKVTTS_DEC_V2]); Y.up[3] = bp(p[RWKVTTS_DEC_G2]);
        Y.w0 = bp(p[RWKVTTS_DEC_W0]); Y.a0 = bp(p[RWKVTTS_DEC_A0]); Y.v0 = bp(p[RWKVTTS_DEC_V0]);
        Y.k_k = bp(p[RWKVTTS_DEC_K_K]); Y.k_a = bp(p[RWKVTTS_DEC_K_A]); Y.r_k = bp(p[RWKVTTS_DEC_R_K]);
        Y.gn_w = bp(p[RWKVTTS_DEC_GN_W]); Y.gn_b = bp(p[RWKVTTS_DEC_GN_B]);
        Y.state = static_cast<float *>(const_cast<void *>(p[RWKVTTS_DEC_STATE]));
        Y.att_shift = static_cast<bf16 *>(const_cast<void *>(p[RWKVTTS_DEC_ATT_SHIFT]));
        Y.ffn_shift = static_cast<bf16 *>(const_cast<void *>(p[RWKVTTS_DEC_FFN_SHIFT]));
        // X slots: 0 r, 1 w, 2 k, 3 v, 4 a, 5 g
        add_job(Y.p2, D.X + 0 * RC, C, p[RWKVTTS_DEC_W_R], C, D.rkv + 0 * RC, C, C, C, kEpiBf16);
        add_job(Y.p2, D.X + 2 * RC, C, p[RWKVTTS_DEC_W_K], C, D.rkv + 1 * RC, C, C, C, kEpiBf16);
        add_job(Y.p2, D.X + 3 * RC, C, p[RWKVTTS_DEC_W_V], C, D.rkv + 2 * RC, C, C, C, kEpiBf16);
        int off = 0;
        add_job(Y.p2, D.X + 1 * RC, C, p[RWKVTTS_DEC_W1], C, D.hl + off, D.Dtot, D.D[0], C, kEpiTanh); off += D.D[0];
        add_job(Y.p2, D.X + 4 * RC, C, p[RWKVTTS_DEC_A1], C, D.hl + off, D.Dtot, D.D[1], C, kEpiBf16); off += D.D[1];
        if (p[RWKVTTS_DEC_V1] != nullptr)
            add_job(Y.p2, D.X + 3 * RC, C, p[RWKVTTS_DEC_V1], C, D.hl + off, D.Dtot, D.D[2], C, kEpiBf16);
        off += D.D[2];
        add_job(Y.p2, D.X + 5 * RC, C, p[RWKVTTS_DEC_G1], C, D.hl + off, D.Dtot, D.D[3], C, kEpiSigmoid);
        add_job(Y.p4, D.o, C, p[RWKVTTS_DEC_W_O], C, D.att, C, C, C, kEpiBf16);
        add_job(Y.p6, D.Xf, C, p[RWKVTTS_DEC_FFN_KEY], C, D.kf, F, F, C, kEpiSqRelu);
        for (int kc = 0; kc < D.nchunk; kc++)
            add_job(Y.p7, D.kf + (size_t)kc * C, F, bp(p[RWKVTTS_DEC_FFN_VALUE]) + (size_t)kc * C, F,
                    D.part + (size_t)kc * RC, C, C, C, kEpiF32);
    }
    add_job(D.head, D.hN, C, mp[RWKVTTS_DEC_HEAD], C, D.logits, V, V, C, kEpiLogits);

    cudaError_t e = cudaMemsetAsync(ws, 0, lo.total, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(ws, host.data(), host.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    HostInfo hi{};
    if ((e = cudaGetDevice(&hi.device)) != cudaSuccess) return e;
    int sms = 0, per_sm = 0, coop = 0;
    if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, hi.device)) != cudaSuccess) return e;
    if (!coop) return cudaErrorNotSupported;           // the grid barrier needs every CTA resident
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, hi.device)) != cudaSuccess) return e;
    hi.nkb = C <= 1024 ? 2 : 4;
    // dynamic shared memory: the phase scratch + the staged LoRA up-projection rows of two (b, head) units
    hi.smem = (size_t)kRedFloats * sizeof(float) + (size_t)kC * (D.Dtot + 4 * 8) * sizeof(bf16);
    const void *fn = hi.nkb == 2 ? (const void *)decode_step_kernel<2> : (const void *)decode_step_kernel<4>;
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hi.smem)) != cudaSuccess) return e;
    e = hi.nkb == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<2>, kThreads, hi.smem)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<4>, kThreads, hi.smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorCooperativeLaunchTooLarge;
    hi.grid = sms;
    std::lock_guard<std::mutex> lk(g_mu);
    g_plans[ws] = hi;
    return cudaSuccess;
}

cudaError_t decode_step(void *ws, const long long *tok_in, long long *tok_out, int greedy, int suppress_eos,
                        const long long *eos, int n_eos, long long pad, unsigned long long *prof, cudaStream_t st) {
    HostInfo hi{};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_plans.find(ws);
        if (it == g_plans.end()) return cudaErrorInvalidValue;
        hi = it->second;
    }
    if (watchdog_needs_install(2, st)) {
        cudaError_t e = tc05::watchdog_install(watchdog_record(), 3);
        if (e != cudaSuccess) return e;
    }
    StepArgs a{};
    a.tok_in = tok_in; a.tok_out = tok_out; a.pad = pad; a.greedy = greedy; a.suppress_eos = suppress_eos; a.prof = prof;
    const char *dbg = getenv("RWKVTTS_DECODE_SKIP");
    a.debug_skip = dbg != nullptr ? atoi(dbg) : 0;
    a.n_eos = n_eos < kMaxEos ? n_eos : kMaxEos;
    for (int i = 0; i < a.n_eos; i++) a.eos[i] = eos[i];
    const Desc *Dp = static_cast<const Desc *>(ws);
    void *args[] = {(void *)&Dp, (void *)&a};
    count_launch();
    return hi.nkb == 2
        ? cudaLaunchCooperativeKernel((const void *)decode_step_kernel<2>, dim3(hi.grid), dim3(kThreads), args, hi.smem, st)
        : cudaLaunchCooperativeKernel((const void *)decode_step_kernel<4>, dim3(hi.grid), dim3(kThreads), args, hi.smem, st);
}

void decode_release(void *ws) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_plans.erase(ws);
}

size_t decode_workspace(const int *d, size_t *offs) {
    if (!dims_ok(d)) return 0;
    const Layout lo = layout(d);
    if (offs != nullptr) { offs[0] = lo.logits; offs[1] = lo.tok; offs[2] = lo.done; }
    return lo.total;
}

}  // namespace dec
}  // namespace rwkvtts
