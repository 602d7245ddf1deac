// C entry point for the forward-v2 sketch (proto/wkv7_tc_fwd_v2.cu), used only by scripts/check_tc_fwd_v2.py.
#include <cuda_runtime.h>
namespace rwkvtts {
cudaError_t launch_tc_fwd_v2(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                             const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                             cudaStream_t st);
}
extern "C" __attribute__((visibility("default"))) int fwd_v2(int B, int T, int H, const void *w, const void *q, const void *k,
                                                              const void *v, const void *a, const void *b, void *y,
                                                              float *ckT, float *sa, const float *s0, float *sT, void *stream) {
    return (int)rwkvtts::launch_tc_fwd_v2(B, T, H, w, q, k, v, a, b, y, ckT, sa, s0, sT, (cudaStream_t)stream);
}
