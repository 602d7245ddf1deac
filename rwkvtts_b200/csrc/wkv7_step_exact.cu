// Stateful WKV-7 forward in the REFERENCE'S ARITHMETIC ORDER: bit-identical outputs and state to the reference's
// rwkv7_state_fwd_fp16 / wkv7s kernels (model/llm/cuda/rwkv7_state_fwd_fp16.cu:9-57, wkv7s.cu:9-57).
//
// Why it exists: the north_star's only bit-exact criterion is "identical argmax token ids under greedy decode".  Greedy
// decoding amplifies a last-bit difference of one logit into a different sequence, so the criterion can only be met by
// reproducing the reference step's floating-point operations one for one.  The default step kernel (wkv7_scan.cu:
// four lanes per state row, quad reductions) sums in another order; this one keeps one thread per value row and the
// exact operation sequence read off the SASS of the reference kernel as the checker builds it (cuobjdump; compiled
// with the reference's --use_fast_math: every operation flushes denormals, FFMA.FTZ / FMUL.FTZ / MUFU.EX2):
//     d_j   = ex2(-(ex2(w_j * log2e)) * log2e)                                  (__expf(-__expf(w)))
//     sa    = fma(a_63, S_63, ... fma(a_1, S_1, fma(a_0, S_0, 0)))              ascending j
//     S_j   = fma(sa, b_j, fma(S_j, d_j, k_j * v_i))
//     y     = fma(S_63, r_63, ... fma(S_0, r_0, 0))                             ascending j, updated S
// This file is compiled with --use_fast_math (rwkvtts_b200/build.py) so that the same instructions come out.
// Selected with rwkvtts_set_step_mode(1) ("exact" decode: generate(..., exact=True)); tests/test_wkv7_gpu.py compares
// it bit for bit with the reference kernel.  One CTA of 64 threads per (batch, head), any T.
#include "wkv7_common.cuh"

namespace rwkvtts {

__global__ void __launch_bounds__(kC) wkv7_state_exact_kernel(int T, int H, const bf16 *__restrict__ w,
                                                              const bf16 *__restrict__ q, const bf16 *__restrict__ k,
                                                              const bf16 *__restrict__ v, const bf16 *__restrict__ a,
                                                              const bf16 *__restrict__ b, bf16 *__restrict__ y,
                                                              float *__restrict__ state) {
    __shared__ float sq[kC], sd[kC], sk[kC], sa_[kC], sb[kC];
    const int bb = blockIdx.x / H, hh = blockIdx.x % H, i = threadIdx.x;
    float *row = state + ((size_t)blockIdx.x * kC + i) * kC;         // S[b][h][value = i][key]
    float S[kC];
#pragma unroll
    for (int j = 0; j < kC; j++) S[j] = row[j];
    const size_t tok_stride = (size_t)H * kC;
    size_t off = (size_t)bb * T * tok_stride + (size_t)hh * kC + i;
    for (int t = 0; t < T; t++, off += tok_stride) {
        __syncthreads();
        sq[i] = __bfloat162float(q[off]);
        sd[i] = __expf(-__expf(__bfloat162float(w[off])));
        sk[i] = __bfloat162float(k[off]);
        sa_[i] = __bfloat162float(a[off]);
        sb[i] = __bfloat162float(b[off]);
        __syncthreads();
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < kC; j++) dot = __fmaf_rn(sa_[j], S[j], dot);
        const float vi = __bfloat162float(v[off]);
        float out = 0.f;
#pragma unroll
        for (int j = 0; j < kC; j++) {
            S[j] = __fmaf_rn(dot, sb[j], __fmaf_rn(S[j], sd[j], __fmul_rn(sk[j], vi)));
            out = __fmaf_rn(S[j], sq[j], out);
        }
        y[off] = __float2bfloat16_rn(out);
    }
#pragma unroll
    for (int j = 0; j < kC; j++) row[j] = S[j];
}

cudaError_t launch_state_exact(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                               const void *a, const void *b, void *y, float *state, cudaStream_t st) {
    count_launch();
    wkv7_state_exact_kernel<<<dim3(B * H), dim3(kC), 0, st>>>(T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k,
                                                             (const bf16 *)v, (const bf16 *)a, (const bf16 *)b,
                                                             (bf16 *)y, state);
    return cudaGetLastError();
}

}  // namespace rwkvtts
