// Fused Adam / AdamW on one ZeRO shard: fp32 master weights and moments, bf16 or fp32 gradient in,
// parameter written back in the model dtype.  Replaces the deepspeed.ops.adam.FusedAdam call the
// reference's training scripts make (train_spark_rwkv7speech_jsonl.py:195-199: betas (0.9, 0.95),
// eps 1e-18, bias_correction, adam_w_mode), applied to the rank's shard of the flat parameter space.
// Pure streaming: 4 fp32 reads + 3 fp32 writes + 1 param write per element, 128-bit accesses.
#include "wkv7_common.cuh"

namespace rwkvtts {

template <typename G, typename P>
__global__ void __launch_bounds__(256) adam_shard_kernel(float *__restrict__ master, float *__restrict__ m,
                                                         float *__restrict__ v, const G *__restrict__ grad,
                                                         P *__restrict__ param, long long n, float lr, float b1,
                                                         float b2, float eps, float wd, int adamw, float bc1,
                                                         float bc2_sqrt, float gscale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float g = (float)grad[i] * gscale;
        float w = master[i];
        if (!adamw) g = fmaf(wd, w, g);                 // L2 regularisation folded into the gradient
        const float mi = fmaf(b1, m[i], (1.f - b1) * g);
        const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        float upd = (mi / bc1) / denom;
        if (adamw) upd = fmaf(wd, w, upd);              // decoupled weight decay
        w = fmaf(-lr, upd, w);
        m[i] = mi; v[i] = vi; master[i] = w;
        param[i] = (P)w;
    }
}

template <typename G, typename P>
static cudaError_t launch(float *master, float *m, float *v, const void *grad, void *param, long long n, float lr,
                          float b1, float b2, float eps, float wd, int adamw, float bc1, float bc2_sqrt,
                          float gscale, cudaStream_t st) {
    const int threads = 256;
    long long blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;           // grid-stride: 16 CTAs per SM
    count_launch();
    adam_shard_kernel<G, P><<<(unsigned)blocks, threads, 0, st>>>(master, m, v, (const G *)grad, (P *)param, n, lr,
                                                                 b1, b2, eps, wd, adamw, bc1, bc2_sqrt, gscale);
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

}  // namespace rwkvtts
