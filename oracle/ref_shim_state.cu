// extern "C" shim over the reference's stateful launcher
// (/root/reference/model/llm/cuda/rwkv7_state_fwd_fp16.cu:59-63), compiled next to this file
// by oracle/Makefile from where it lies.  Test infrastructure only.
#include "ATen/ATen.h"
#include <cuda_runtime.h>
typedef at::BFloat16 dtype;
void cuda_forward(int B, int T, int C, int H, float* state, dtype* r, dtype* w, dtype* k, dtype* v, dtype* a,
                  dtype* b, dtype* y);
extern "C" int ref_state_forward(int B, int T, int C, int H, float* state, void* r, void* w, void* k, void* v,
                                 void* a, void* b, void* y) {
    cuda_forward(B, T, C, H, state, (dtype*)r, (dtype*)w, (dtype*)k, (dtype*)v, (dtype*)a, (dtype*)b, (dtype*)y);
    return (int)cudaGetLastError();
}
