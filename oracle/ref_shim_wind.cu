// extern "C" shim over the reference's launchers so ctypes can call them.
// Declares (does not copy) cuda_forward / cuda_backward of
// /root/reference/model/llm/cuda/wkv7_cuda.cu:132-138, which is compiled next to this
// file by oracle/Makefile from where it lies.  Test infrastructure only.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
using bf = __nv_bfloat16;
void cuda_forward(int B, int T, int H, bf* w, bf* q, bf* k, bf* v, bf* z, bf* a, bf* y, float* s, float* sa);
void cuda_backward(int B, int T, int H, bf* w, bf* q, bf* k, bf* v, bf* z, bf* a, bf* dy, float* s, float* sa,
                   bf* dw, bf* dq, bf* dk, bf* dv, bf* dz, bf* da);
extern "C" int ref_wind_forward(int B, int T, int H, void* w, void* q, void* k, void* v, void* z, void* a,
                                void* y, float* s, float* sa) {
    cuda_forward(B, T, H, (bf*)w, (bf*)q, (bf*)k, (bf*)v, (bf*)z, (bf*)a, (bf*)y, s, sa);
    return (int)cudaGetLastError();
}
extern "C" int ref_wind_backward(int B, int T, int H, void* w, void* q, void* k, void* v, void* z, void* a,
                                 void* dy, float* s, float* sa, void* dw, void* dq, void* dk, void* dv,
                                 void* dz, void* da) {
    cuda_backward(B, T, H, (bf*)w, (bf*)q, (bf*)k, (bf*)v, (bf*)z, (bf*)a, (bf*)dy, s, sa, (bf*)dw, (bf*)dq,
                  (bf*)dk, (bf*)dv, (bf*)dz, (bf*)da);
    return (int)cudaGetLastError();
}
