// Shared device helpers for the WKV-7 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace rwkvtts {

extern std::atomic<long long> g_kernel_launches;   // defined in capi.cu
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// Pinned, device-visible record the chunked kernels' mbarrier watchdog writes before it traps (tc05.cuh);
// owned by capi.cu, nullptr if it could not be allocated.  8 x u64.
unsigned long long *watchdog_record();
// true once per (translation unit slot, device): the caller then installs the record pointer into its own
// copy of the device symbol.  Never true while `st` is being captured into a CUDA graph.
bool watchdog_needs_install(int slot, cudaStream_t st);

constexpr int kC = 64;       // head size (reference: -D_C_=64)
constexpr int kChunk = 16;   // snapshot spacing of the scan kernels (reference: _CHUNK_LEN_)

using bf16 = __nv_bfloat16;

// Scratch layout of the tcgen05 training pair inside the caller's `s` / `sa` tensors (reference sizes).
// Both are stored per (batch*head, chunk) as dense tensor-core operand tiles (UMMA canonical K-major,
// no swizzle: 8-row x 16-byte core matrices of 128 contiguous bytes), so the backward brings them into shared
// memory with bulk copies and feeds them to tcgen05.mma without touching a register:
//   checkpoint (4096 floats): S0^T, row = key, k = value    off = (value/4)*256 + (key/8)*32 + (key%8)*4 + value%4
//   sa         (1024 floats): U,    row = token, k = value  off = (value/4)*64 + (token/8)*32 + (token%8)*4 + value%4
constexpr int kCkFloats = 4096, kCkLbo = 256;
constexpr int kUFloats = 1024, kULbo = 64;

// What one CTA of the chunked kernels works on: a (batch, head) pair of a dense [B,T,H,64] launch, or a (sequence, head)
// pair of a packed launch ([1,T_total,H,64] with cu_seqlens: state and token-shift restart at every boundary;
// data/utils/spark_dataset.py:111-162, utils/multiple_jsonl.py:77-135 build such batches).  `len` need not be a
// multiple of 16 in the packed case: tokens of the last chunk beyond `len` are loaded as zeros and never stored.
struct SeqWork {
    size_t base;     // element offset of token 0, channel 0 of this head
    size_t ck0;      // index of the sequence's first chunk slot in the scratch tensors (per head)
    int len, nC;     // tokens, 16-token chunks (ceil)
    int bh;          // dense: b*H + h (s0 / sT / ds0 index); packed: unused
    int h, tok0;     // head and index of the sequence's first token in the [tokens, H, 64] view (tensor-map coordinates)
};
__device__ __forceinline__ SeqWork seq_work(int T, int H, const int *cu, const int *cbase) {
    SeqWork w;
    const int hh = blockIdx.x % H, bb = blockIdx.x / H;
    const size_t tok_stride = (size_t)H * kC;
    if (cu != nullptr) {
        const int t0 = cu[bb];
        w.len = cu[bb + 1] - t0;
        w.nC = (w.len + kChunk - 1) / kChunk;
        w.base = (size_t)t0 * tok_stride + (size_t)hh * kC;
        w.ck0 = (size_t)cbase[bb] * H + (size_t)hh * w.nC;
        w.tok0 = t0;
    } else {
        w.len = T;
        w.nC = T / kChunk;
        w.base = (size_t)bb * T * tok_stride + (size_t)hh * kC;
        w.ck0 = (size_t)blockIdx.x * w.nC;
        w.tok0 = bb * T;
    }
    w.bh = blockIdx.x;
    w.h = hh;
    return w;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// 8 bf16 (one uint4) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4 &u, float *f) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
    f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
    f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// sum over the 4 lanes that share a row (lane bits 0..1)
__device__ __forceinline__ float quad_sum(float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    return x;
}

}  // namespace rwkvtts
