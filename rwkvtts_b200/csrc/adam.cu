// Fused Adam / AdamW on one ZeRO shard: fp32 master weights and moments, bf16 or fp32 gradient in,
// parameter written back in the model dtype.  Replaces the deepspeed.ops.adam.FusedAdam call the
// reference's training scripts make (train_spark_rwkv7speech_jsonl.py:195-199: betas (0.9, 0.95),
// eps 1e-18, bias_correction, adam_w_mode), applied to the rank's shard of the flat parameter space.
// Pure streaming [HBM roofline]: per element 3 fp32 + 1 gradient read, 3 fp32 + 1 parameter write
// (28 B with bf16 gradients / parameters); every thread moves 4 adjacent elements per iteration
// (16-byte fp32 accesses, 8-byte bf16 accesses).
//
//   adam_shard_kernel   one hyper-parameter set for the whole range (stand-alone FusedAdam.step, round-1 ABI)
//   adam_multi_kernel   the engine's call: the range is a run of segments (one per parameter tensor piece) that
//                       map to optimizer param groups with their own lr / weight decay / bias corrections; the
//                       gradient scale (clipping) and the skip decision (non-finite gradients) are read from a
//                       device-resident {norm^2, non-finite flag} pair, so engine.step() needs no host sync.
#include "wkv7_common.cuh"

namespace rwkvtts {

struct AdamHp { float lr, wd, bc1, bc2_sqrt; };
constexpr int kMaxGroups = 8;
struct AdamGroups { AdamHp g[kMaxGroups]; };

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float *p, float (&x)[4]) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&x)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(x[0], x[1], x[2], x[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16 *p, float (&x)[4]) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
    }
    static __device__ __forceinline__ void store(bf16 *p, const float (&x)[4]) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack2(x[0], x[1]), pack2(x[2], x[3]));
    }
};

__device__ __forceinline__ void adam_elem(float &w, float &m, float &v, float g, const AdamHp &hp, float b1, float b2,
                                          float eps, int adamw) {
    if (!adamw) g = fmaf(hp.wd, w, g);                  // L2 regularisation folded into the gradient
    m = fmaf(b1, m, (1.f - b1) * g);
    v = fmaf(b2, v, (1.f - b2) * g * g);
    const float denom = sqrtf(v) / hp.bc2_sqrt + eps;
    float upd = (m / hp.bc1) / denom;
    if (adamw) upd = fmaf(hp.wd, w, upd);               // decoupled weight decay
    w = fmaf(-hp.lr, upd, w);
}

// n4 = n / 4 vector iterations + scalar tail; all base pointers 16-byte (fp32) / 8-byte (bf16) aligned
template <typename G, typename P>
__global__ void __launch_bounds__(256) adam_shard_kernel(float *__restrict__ master, float *__restrict__ m,
                                                         float *__restrict__ v, const G *__restrict__ grad,
                                                         P *__restrict__ param, long long n, AdamHp hp, float b1,
                                                         float b2, float eps, int adamw, float gscale, int vec_ok) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = vec_ok ? n / 4 : 0;
    for (long long i = tid0; i < n4; i += stride) {
        float w[4], mm[4], vv[4], g[4];
        Vec4<float>::load(master + 4 * i, w); Vec4<float>::load(m + 4 * i, mm); Vec4<float>::load(v + 4 * i, vv);
        Vec4<G>::load(grad + 4 * i, g);
#pragma unroll
        for (int e = 0; e < 4; e++) adam_elem(w[e], mm[e], vv[e], g[e] * gscale, hp, b1, b2, eps, adamw);
        Vec4<float>::store(master + 4 * i, w); Vec4<float>::store(m + 4 * i, mm); Vec4<float>::store(v + 4 * i, vv);
        Vec4<P>::store(param + 4 * i, w);
    }
    for (long long i = 4 * n4 + tid0; i < n; i += stride) {
        float w = master[i], mi = m[i], vi = v[i];
        adam_elem(w, mi, vi, (float)grad[i] * gscale, hp, b1, b2, eps, adamw);
        master[i] = w; m[i] = mi; v[i] = vi; param[i] = (P)w;
    }
}

// Segment table: seg_end[s] = exclusive end (element index inside this range) of segment s, ascending, every end a
// multiple of 4 except possibly the last (the engine pads parameter tensors to 8 elements); seg_group[s] = param group.
// One CTA owns kTile consecutive elements: it finds its first segment by bisection once, then walks forward.
constexpr int kTile = 4096;
template <typename G, typename P>
__global__ void __launch_bounds__(256) adam_multi_kernel(float *__restrict__ master, float *__restrict__ m,
                                                         float *__restrict__ v, const G *__restrict__ grad,
                                                         P *__restrict__ param, long long n,
                                                         const long long *__restrict__ seg_end,
                                                         const int *__restrict__ seg_group, int nseg, AdamGroups hps,
                                                         float b1, float b2, float eps, int adamw,
                                                         const float *__restrict__ stat, float clip,
                                                         unsigned long long *__restrict__ skipped) {
    float gscale = 1.f;
    if (stat != nullptr) {
        const float norm = sqrtf(stat[0]);
        if (stat[1] > 0.f || !isfinite(norm)) {          // same decision in every CTA and on every rank: skip the step
            if (skipped != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped, 1ull);
            return;
        }
        if (clip > 0.f && norm > clip) gscale = clip / (norm + 1e-6f);
    }
    const long long ntiles = (n + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long t0 = tile * kTile, t1 = min(n, t0 + (long long)kTile);
        int lo = 0, hi = nseg - 1;                       // first segment with seg_end > t0
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (seg_end[mid] > t0) hi = mid; else lo = mid + 1;
        }
        int seg = lo;
        for (long long i = t0 + 4 * threadIdx.x; i < t1; i += 4 * 256) {
            while (seg < nseg - 1 && seg_end[seg] <= i) seg++;
            const AdamHp hp = hps.g[seg_group[seg]];
            if (i + 4 <= t1) {
                float w[4], mm[4], vv[4], g[4];
                Vec4<float>::load(master + i, w); Vec4<float>::load(m + i, mm); Vec4<float>::load(v + i, vv);
                Vec4<G>::load(grad + i, g);
#pragma unroll
                for (int e = 0; e < 4; e++) adam_elem(w[e], mm[e], vv[e], g[e] * gscale, hp, b1, b2, eps, adamw);
                Vec4<float>::store(master + i, w); Vec4<float>::store(m + i, mm); Vec4<float>::store(v + i, vv);
                Vec4<P>::store(param + i, w);
            } else {
                for (long long j = i; j < t1; j++) {
                    float w = master[j], mi = m[j], vi = v[j];
                    adam_elem(w, mi, vi, (float)grad[j] * gscale, hp, b1, b2, eps, adamw);
                    master[j] = w; m[j] = mi; v[j] = vi; param[j] = (P)w;
                }
            }
        }
    }
}

// ---- the same update fused with BOTH collectives of the ZeRO step over NVLink / NVSwitch --------------------------------
// reduce-scatter(AVG) of the gradients -> Adam on the rank's slice -> all-gather of the updated parameters, in ONE kernel
// and without staging buffers: every rank's flat gradient / parameter buffer lives in symmetric memory (same layout on
// every GPU, mapped into every peer), so for element i of its slice a rank
//   kMulticast: issues ONE multimem.ld_reduce on the multicast address of the gradient buffers -- the NVSwitch adds the
//               W ranks' values in flight (fp32 accumulation) and returns the sum -- and ONE multimem.st of the new bf16
//               parameter, which the switch delivers to all W copies (NVLS: 1x instead of (W-1)x link traffic each way);
//   otherwise : loads the W peers' gradients through their mapped pointers (16-byte loads, all in flight together) and
//               stores the parameter into each peer's buffer.
// The caller brackets the launches with symmetric-memory barriers (all gradients final before, all parameter writes
// landed after) -- rwkvtts_b200/engine.py.  bf16 gradients and parameters only.  Same segment walk as adam_multi_kernel.
constexpr int kMaxPeers = 8;
struct PeerPtrs { const bf16 *grad[kMaxPeers]; bf16 *param[kMaxPeers]; };

__device__ __forceinline__ uint4 ld_peer_v4(const void *p) {       // never cached: the peer rewrites it every step
    uint4 r;
    asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void mc_ld_reduce_bf16x8(const void *mc, float (&x)[8]) {
    uint32_t r[4];
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(mc) : "memory");
#pragma unroll
    for (int i = 0; i < 4; i++) { x[2 * i] = bf16_lo(r[i]); x[2 * i + 1] = bf16_hi(r[i]); }
}
__device__ __forceinline__ void mc_st_bf16x8(void *mc, const float (&x)[8]) {
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "r"(pack2(x[0], x[1])), "r"(pack2(x[2], x[3])), "r"(pack2(x[4], x[5])), "r"(pack2(x[6], x[7]))
                 : "memory");
}

constexpr int kTileP = 8192;         // elements per CTA tile: 8 per thread and iteration
template <bool kMulticast>
__global__ void __launch_bounds__(256) adam_p2p_kernel(float *__restrict__ master, float *__restrict__ m,
                                                       float *__restrict__ v, PeerPtrs peers, const bf16 *mc_grad,
                                                       bf16 *mc_param, int world, long long flat_off, long long n,
                                                       const long long *__restrict__ seg_end,
                                                       const int *__restrict__ seg_group, int nseg, AdamGroups hps,
                                                       float b1, float b2, float eps, int adamw,
                                                       const float *__restrict__ stat, float *__restrict__ norm_sq,
                                                       unsigned long long *__restrict__ skipped) {
    if (stat != nullptr && stat[1] > 0.f) {              // a non-finite gradient on some rank: skip the step everywhere
        if (skipped != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped, 1ull);
        return;
    }
    const float inv_w = 1.f / (float)world;
    float ss = 0.f;
    const long long ntiles = (n + kTileP - 1) / kTileP;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long t0 = tile * kTileP, t1 = min(n, t0 + (long long)kTileP);
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (seg_end[mid] > t0) hi = mid; else lo = mid + 1;
        }
        int seg = lo;
        for (long long i = t0 + 8 * threadIdx.x; i < t1; i += 8 * 256) {       // n is a multiple of 8 (engine padding)
            while (seg < nseg - 1 && seg_end[seg] <= i) seg++;
            const AdamHp hp = hps.g[seg_group[seg]];
            float g[8];
            if (kMulticast) {
                mc_ld_reduce_bf16x8(mc_grad + flat_off + i, g);
            } else {
                uint4 raw[kMaxPeers];
#pragma unroll
                for (int r = 0; r < kMaxPeers; r++)
                    if (r < world) raw[r] = ld_peer_v4(peers.grad[r] + flat_off + i);
#pragma unroll
                for (int e = 0; e < 8; e++) g[e] = 0.f;
#pragma unroll
                for (int r = 0; r < kMaxPeers; r++)
                    if (r < world) {
                        float f[8];
                        unpack8(raw[r], f);
#pragma unroll
                        for (int e = 0; e < 8; e++) g[e] += f[e];
                    }
            }
            float w[8], mm[8], vv[8];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float a4[4], b4[4], c4[4];
                Vec4<float>::load(master + i + 4 * h, a4); Vec4<float>::load(m + i + 4 * h, b4); Vec4<float>::load(v + i + 4 * h, c4);
#pragma unroll
                for (int e = 0; e < 4; e++) { w[4 * h + e] = a4[e]; mm[4 * h + e] = b4[e]; vv[4 * h + e] = c4[e]; }
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float ge = g[e] * inv_w;
                ss = fmaf(ge, ge, ss);
                adam_elem(w[e], mm[e], vv[e], ge, hp, b1, b2, eps, adamw);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float a4[4] = {w[4 * h], w[4 * h + 1], w[4 * h + 2], w[4 * h + 3]};
                const float b4[4] = {mm[4 * h], mm[4 * h + 1], mm[4 * h + 2], mm[4 * h + 3]};
                const float c4[4] = {vv[4 * h], vv[4 * h + 1], vv[4 * h + 2], vv[4 * h + 3]};
                Vec4<float>::store(master + i + 4 * h, a4); Vec4<float>::store(m + i + 4 * h, b4); Vec4<float>::store(v + i + 4 * h, c4);
            }
            if (kMulticast) {
                mc_st_bf16x8(mc_param + flat_off + i, w);
            } else {
                const uint4 out = make_uint4(pack2(w[0], w[1]), pack2(w[2], w[3]), pack2(w[4], w[5]), pack2(w[6], w[7]));
#pragma unroll
                for (int r = 0; r < kMaxPeers; r++)
                    if (r < world) *reinterpret_cast<uint4 *>(peers.param[r] + flat_off + i) = out;
            }
        }
    }
    if (norm_sq != nullptr) {                            // squared norm of the averaged gradient of this slice
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(norm_sq, ss);
    }
}

// {sum of squares, any non-finite} of a gradient range, accumulated into stat[0..1] (fp32 atomics; the caller zeroes
// stat first).  One pass over the gradients: 2 bytes per element.
template <typename G>
__global__ void __launch_bounds__(256) grad_stat_kernel(const G *__restrict__ grad, long long n, float *__restrict__ stat) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float ss = 0.f;
    bool bad = false;
    const long long n4 = n / 4;
    for (long long i = tid0; i < n4; i += stride) {
        float g[4];
        Vec4<G>::load(grad + 4 * i, g);
#pragma unroll
        for (int e = 0; e < 4; e++) { ss = fmaf(g[e], g[e], ss); bad |= !isfinite(g[e]); }
    }
    for (long long i = 4 * n4 + tid0; i < n; i += stride) {
        const float g = (float)grad[i];
        ss = fmaf(g, g, ss); bad |= !isfinite(g);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const bool anybad = __any_sync(0xffffffffu, bad);
    __shared__ float wsum[8];
    __shared__ int wbad[8];
    if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = ss; wbad[threadIdx.x >> 5] = anybad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f; int b = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { t += wsum[i]; b |= wbad[i]; }
        atomicAdd(&stat[0], t);
        if (b) atomicAdd(&stat[1], 1.f);
    }
}

static unsigned grid_for(long long work_items, int per_block) {
    long long blocks = (work_items + per_block - 1) / per_block;
    if (blocks > 148 * 8) blocks = 148 * 8;             // grid-stride: a multiple of the SM count
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

template <typename G, typename P>
static cudaError_t launch(float *master, float *m, float *v, const void *grad, void *param, long long n, float lr,
                          float b1, float b2, float eps, float wd, int adamw, float bc1, float bc2_sqrt,
                          float gscale, cudaStream_t st) {
    const int vec_ok = ((reinterpret_cast<uintptr_t>(master) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15u) == 0 &&
                       (reinterpret_cast<uintptr_t>(grad) & (4 * sizeof(G) - 1)) == 0 &&
                       (reinterpret_cast<uintptr_t>(param) & (4 * sizeof(P) - 1)) == 0;
    count_launch();
    adam_shard_kernel<G, P><<<grid_for(n / 4 + 1, 256), 256, 0, st>>>(master, m, v, (const G *)grad, (P *)param, n,
                                                                     AdamHp{lr, wd, bc1, bc2_sqrt}, b1, b2, eps, adamw,
                                                                     gscale, vec_ok);
    return cudaGetLastError();
}

cudaError_t launch_adam_shard(float *master, float *m, float *v, const void *grad, int grad_is_bf16, void *param,
                              int param_is_bf16, long long n, float lr, float b1, float b2, float eps, float wd,
                              int adamw, float bc1, float bc2_sqrt, float gscale, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (grad_is_bf16) {
        return param_is_bf16 ? launch<bf16, bf16>(master, m, v, grad, param, n, lr, b1, b2, eps, wd, adamw, bc1, bc2_sqrt, gscale, st)
                             : launch<bf16, float>(master, m, v, grad, param, n, lr, b1, b2, eps, wd, adamw, bc1, bc2_sqrt, gscale, st);
    }
    return param_is_bf16 ? launch<float, bf16>(master, m, v, grad, param, n, lr, b1, b2, eps, wd, adamw, bc1, bc2_sqrt, gscale, st)
                         : launch<float, float>(master, m, v, grad, param, n, lr, b1, b2, eps, wd, adamw, bc1, bc2_sqrt, gscale, st);
}

template <typename G, typename P>
static cudaError_t launch_multi(float *master, float *m, float *v, const void *grad, void *param, long long n,
                                const long long *seg_end, const int *seg_group, int nseg, const AdamGroups &hps, float b1,
                                float b2, float eps, int adamw, const float *stat, float clip,
                                unsigned long long *skipped, cudaStream_t st) {
    count_launch();
    adam_multi_kernel<G, P><<<grid_for(n, kTile), 256, 0, st>>>(master, m, v, (const G *)grad, (P *)param, n, seg_end,
                                                               seg_group, nseg, hps, b1, b2, eps, adamw, stat, clip,
                                                               skipped);
    return cudaGetLastError();
}

// group_hp: host array [ngroups][4] = {lr, weight_decay, bias_correction1, sqrt(bias_correction2)}
cudaError_t launch_adam_multi(float *master, float *m, float *v, const void *grad, int grad_is_bf16, void *param,
                              int param_is_bf16, long long n, const long long *seg_end, const int *seg_group, int nseg,
                              const float *group_hp, int ngroups, float b1, float b2, float eps, int adamw,
                              const float *stat, float clip, unsigned long long *skipped, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    AdamGroups hps{};
    for (int i = 0; i < ngroups && i < kMaxGroups; i++)
        hps.g[i] = AdamHp{group_hp[4 * i], group_hp[4 * i + 1], group_hp[4 * i + 2], group_hp[4 * i + 3]};
    if (grad_is_bf16) {
        return param_is_bf16 ? launch_multi<bf16, bf16>(master, m, v, grad, param, n, seg_end, seg_group, nseg, hps, b1, b2, eps, adamw, stat, clip, skipped, st)
                             : launch_multi<bf16, float>(master, m, v, grad, param, n, seg_end, seg_group, nseg, hps, b1, b2, eps, adamw, stat, clip, skipped, st);
    }
    return param_is_bf16 ? launch_multi<float, bf16>(master, m, v, grad, param, n, seg_end, seg_group, nseg, hps, b1, b2, eps, adamw, stat, clip, skipped, st)
                         : launch_multi<float, float>(master, m, v, grad, param, n, seg_end, seg_group, nseg, hps, b1, b2, eps, adamw, stat, clip, skipped, st);
}

cudaError_t launch_adam_p2p(float *master, float *m, float *v, const void *const *grad_ptrs, void *const *param_ptrs,
                            const void *mc_grad, void *mc_param, int world, long long flat_off, long long n,
                            const long long *seg_end, const int *seg_group, int nseg, const float *group_hp, int ngroups,
                            float b1, float b2, float eps, int adamw, const float *stat, float *norm_sq,
                            unsigned long long *skipped, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    AdamGroups hps{};
    for (int i = 0; i < ngroups && i < kMaxGroups; i++)
        hps.g[i] = AdamHp{group_hp[4 * i], group_hp[4 * i + 1], group_hp[4 * i + 2], group_hp[4 * i + 3]};
    PeerPtrs pp{};
    for (int r = 0; r < world && r < kMaxPeers; r++) {
        pp.grad[r] = static_cast<const bf16 *>(grad_ptrs[r]);
        pp.param[r] = static_cast<bf16 *>(param_ptrs[r]);
    }
    count_launch();
    const unsigned grid = grid_for(n, kTileP);
    if (mc_grad != nullptr && mc_param != nullptr)
        adam_p2p_kernel<true><<<grid, 256, 0, st>>>(master, m, v, pp, (const bf16 *)mc_grad, (bf16 *)mc_param, world, flat_off,
                                                    n, seg_end, seg_group, nseg, hps, b1, b2, eps, adamw, stat, norm_sq, skipped);
    else
        adam_p2p_kernel<false><<<grid, 256, 0, st>>>(master, m, v, pp, nullptr, nullptr, world, flat_off, n, seg_end,
                                                     seg_group, nseg, hps, b1, b2, eps, adamw, stat, norm_sq, skipped);
    return cudaGetLastError();
}

cudaError_t launch_grad_stat(const void *grad, int grad_is_bf16, long long n, float *stat, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_launch();
    if (grad_is_bf16)
        grad_stat_kernel<bf16><<<grid_for(n / 4 + 1, 256 * 8), 256, 0, st>>>((const bf16 *)grad, n, stat);
    else
        grad_stat_kernel<float><<<grid_for(n / 4 + 1, 256 * 8), 256, 0, st>>>((const float *)grad, n, stat);
    return cudaGetLastError();
}

}  // namespace rwkvtts
