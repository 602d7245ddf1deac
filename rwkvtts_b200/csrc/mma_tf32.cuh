// Warp-level tf32 tensor-core helpers (mma.sync.m16n8k8, fp32 accumulate).
//
// Fragment coordinates (PTX ISA, m16n8k8 .tf32), g = lane>>2, tq = lane&3:
//   A (16x8):  a0=(g, tq)  a1=(g+8, tq)  a2=(g, tq+4)  a3=(g+8, tq+4)
//   B (8x8):   b0=(k=tq, n=g)            b1=(k=tq+4, n=g)
//   C (16x8):  c0=(g, 2tq) c1=(g, 2tq+1) c2=(g+8, 2tq) c3=(g+8, 2tq+1)
//
// "perm" variants use a permuted order of the 8 k-indices of a k-step: slot tq <-> k = 2tq,
// slot tq+4 <-> k = 2tq+1.  A C fragment is then directly an A fragment of the next product
// (a0=c0, a1=c2, a2=c1, a3=c3) and a B fragment is two ADJACENT elements (one 64-bit load).
// Any product may use the permuted order as long as A and B agree.
//
// Operands in shared memory are fp32 words already rounded to tf32 (cvt.rna) when written,
// so fragment loads are plain LDS.
#pragma once
#include <stdint.h>

namespace rwkvtts {

// round to nearest (ties away) at bit 13 = cvt.rna.tf32.f32 for finite inputs (the cvt itself expands to ~5
// instructions with its NaN/Inf handling).  The low bits are cleared although the tensor cores ignore them: the
// CUDA cores read the same tiles (dw = sum of products with heavy cancellation) and must see the same values.
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ float tf32r(float x) { return __uint_as_float(f2tf32(x)); }

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// accumulator tile -> A fragment in the permuted k order (rounds to tf32)
__device__ __forceinline__ void acc_to_a_perm(uint32_t (&a)[4], const float (&c)[4]) {
    a[0] = f2tf32(c[0]); a[1] = f2tf32(c[2]); a[2] = f2tf32(c[1]); a[3] = f2tf32(c[3]);
}

// A fragment of the logical matrix A[m][k] = p[m*ms + k*ks], rows m0.., k-step k0.., natural order
__device__ __forceinline__ void lda(uint32_t (&a)[4], const float *p, int ms, int ks, int m0, int k0, int g, int tq) {
    const float *q = p + (m0 + g) * ms + (k0 + tq) * ks;
    a[0] = __float_as_uint(q[0]);
    a[1] = __float_as_uint(q[8 * ms]);
    a[2] = __float_as_uint(q[4 * ks]);
    a[3] = __float_as_uint(q[8 * ms + 4 * ks]);
}
// same, permuted k order
__device__ __forceinline__ void lda_perm(uint32_t (&a)[4], const float *p, int ms, int ks, int m0, int k0, int g,
                                         int tq) {
    const float *q = p + (m0 + g) * ms + (k0 + 2 * tq) * ks;
    a[0] = __float_as_uint(q[0]);
    a[1] = __float_as_uint(q[8 * ms]);
    a[2] = __float_as_uint(q[ks]);
    a[3] = __float_as_uint(q[8 * ms + ks]);
}
// B fragment of the logical matrix B[k][n] = p[k*ks + n*ns], natural order
__device__ __forceinline__ void ldb(uint32_t (&b)[2], const float *p, int ks, int ns, int k0, int n0, int g, int tq) {
    const float *q = p + (k0 + tq) * ks + (n0 + g) * ns;
    b[0] = __float_as_uint(q[0]);
    b[1] = __float_as_uint(q[4 * ks]);
}
// permuted k order, generic strides
__device__ __forceinline__ void ldb_perm(uint32_t (&b)[2], const float *p, int ks, int ns, int k0, int n0, int g,
                                         int tq) {
    const float *q = p + (k0 + 2 * tq) * ks + (n0 + g) * ns;
    b[0] = __float_as_uint(q[0]);
    b[1] = __float_as_uint(q[ks]);
}
// permuted k order with k contiguous in memory (ks == 1, address 8-byte aligned): one LDS.64
__device__ __forceinline__ void ldb_perm_k1(uint32_t (&b)[2], const float *p, int ns, int k0, int n0, int g, int tq) {
    const float2 v = *reinterpret_cast<const float2 *>(p + (n0 + g) * ns + k0 + 2 * tq);
    b[0] = __float_as_uint(v.x);
    b[1] = __float_as_uint(v.y);
}

// named barrier over `n` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

}  // namespace rwkvtts
