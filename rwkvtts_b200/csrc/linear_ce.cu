// Cross-entropy over a chunk of logits, loss AND gradient in one pass, in place  [HBM roofline: one read + one write of
// the chunk].  The middle piece of the fused linear + cross-entropy head (rwkvtts_b200/fused.py::linear_cross_entropy):
//
//   reference: the Spark / XY wrappers call rwkvfla's FusedLinearCrossEntropyLoss(hidden, labels, lm_head.weight) while
//   training (model/llm/spark_llm.py:139-158) so that the [tokens, V] logits (V = 8193 Spark, 66.7 k XY channel 0,
//   SURVEY.md section 8 row a11) are never held; the eager alternative is lm_head -> float() -> CrossEntropyLoss.
//
// Per token chunk the host side runs   logits = h_chunk @ W^T (cuBLAS, bf16)  ->  THIS KERNEL  ->  dh_chunk = g @ W,
// dW += g^T @ h_chunk (cuBLAS), where the kernel turns every logits row into
//     loss_r = logsumexp(z_r) - (1 - eps) z_r[label] - eps mean(z_r)          (eps = label smoothing)
//     g_r    = (softmax(z_r) - (1 - eps) onehot(label) - eps / V) * scale      (scale = 1 / #valid labels, device scalar)
// and rows whose label is `ignore_index` into loss 0 and a zero gradient row.  Columns >= V (the weight is zero-padded
// to a multiple of 8 rows so that the GEMMs stay on the aligned tensor-core path) are excluded and get gradient 0.
// One CTA per row; the row lives in registers when V <= 16384 (Spark: 8193 -> 5 x 8 values per thread), otherwise it is
// re-read from L2.  fp32 math, exp2 with pre-scaled arguments.
#include "wkv7_common.cuh"

namespace rwkvtts {

constexpr int kCeThreads = 256;
constexpr int kCeRegVec = 8;                 // uint4 pieces per thread kept in registers: 8 * 8 * 256 = 16384 columns

__device__ __forceinline__ float block_reduce(float x, bool is_max, float *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float y = __shfl_xor_sync(0xffffffffu, x, o);
        x = is_max ? fmaxf(x, y) : x + y;
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[w] = x;
    __syncthreads();
    float r = sh[0];
#pragma unroll
    for (int i = 1; i < kCeThreads / 32; i++) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
    __syncthreads();
    return r;
}

template <bool kInRegs>
__global__ void __launch_bounds__(kCeThreads) ce_fwd_bwd_kernel(bf16 *__restrict__ logits, long long ld, int V,
                                                                const long long *__restrict__ labels,
                                                                long long ignore_index, float eps,
                                                                const float *__restrict__ scale_dev,
                                                                float *__restrict__ loss_rows) {
    __shared__ float sh[kCeThreads / 32];
    const long long r = blockIdx.x;
    bf16 *row = logits + r * ld;
    const long long label = labels[r];
    const int nvec = (int)((ld) >> 3);           // ld is a multiple of 8; columns >= V are padding
    uint4 *rv = reinterpret_cast<uint4 *>(row);
    if (label == ignore_index || label < 0 || label >= V) {
        for (int i = threadIdx.x; i < nvec; i += kCeThreads) rv[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) loss_rows[r] = 0.f;
        return;
    }
    constexpr float kL2e = 1.4426950408889634f;
    // the row in registers (compile-time indexed: kCeRegVec pieces per thread) or re-read from L2
    uint4 keep[kInRegs ? kCeRegVec : 1];
    const int npass = kInRegs ? kCeRegVec : (nvec + kCeThreads - 1) / kCeThreads;
    float mx = -INFINITY, sum_z = 0.f;
#pragma unroll
    for (int j = 0; j < (kInRegs ? kCeRegVec : 1 << 30); j++) {
        if (!kInRegs && j >= npass) break;
        const int i = threadIdx.x + j * kCeThreads;
        if (i < nvec) {
            const uint4 u = rv[i];
            if (kInRegs) keep[j] = u;
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (8 * i + e < V) { mx = fmaxf(mx, f[e]); sum_z += f[e]; }
        }
    }
    mx = block_reduce(mx, true, sh);
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < (kInRegs ? kCeRegVec : 1 << 30); j++) {
        if (!kInRegs && j >= npass) break;
        const int i = threadIdx.x + j * kCeThreads;
        if (i < nvec) {
            const uint4 u = kInRegs ? keep[j] : rv[i];
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (8 * i + e < V) se += exp2f((f[e] - mx) * kL2e);
        }
    }
    se = block_reduce(se, false, sh);
    if (eps > 0.f) sum_z = block_reduce(sum_z, false, sh);
    const float lse = mx + __logf(se);
    const float z_label = __bfloat162float(row[label]);
    __syncthreads();                                  // every thread has read what it needs of the row before it is rewritten
    if (threadIdx.x == 0) loss_rows[r] = lse - (1.f - eps) * z_label - (eps > 0.f ? eps * sum_z / V : 0.f);
    const float scale = *scale_dev;
    const float inv = 1.f / se, smooth = eps / V;
#pragma unroll
    for (int j = 0; j < (kInRegs ? kCeRegVec : 1 << 30); j++) {
        if (!kInRegs && j >= npass) break;
        const int i = threadIdx.x + j * kCeThreads;
        if (i < nvec) {
            const uint4 u = kInRegs ? keep[j] : rv[i];
            float f[8], g[8];
            unpack8(u, f);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int c = 8 * i + e;
                float p = (c < V) ? exp2f((f[e] - mx) * kL2e) * inv - smooth : 0.f;
                if (c == label) p -= (1.f - eps);
                g[e] = p * scale;
            }
            rv[i] = make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
        }
    }
}

cudaError_t launch_ce_fwd_bwd(void *logits, long long rows, int V, long long ld, const long long *labels,
                              long long ignore_index, float label_smoothing, const float *scale_dev, float *loss_rows,
                              cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    count_launch();
    if (ld <= (long long)kCeRegVec * 8 * kCeThreads)
        ce_fwd_bwd_kernel<true><<<(unsigned)rows, kCeThreads, 0, st>>>((bf16 *)logits, ld, V, labels, ignore_index,
                                                                        label_smoothing, scale_dev, loss_rows);
    else
        ce_fwd_bwd_kernel<false><<<(unsigned)rows, kCeThreads, 0, st>>>((bf16 *)logits, ld, V, labels, ignore_index,
                                                                         label_smoothing, scale_dev, loss_rows);
    return cudaGetLastError();
}

}  // namespace rwkvtts
