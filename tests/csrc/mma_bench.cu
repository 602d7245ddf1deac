// Micro-benchmark: legacy mma.sync throughput on sm_100a (tf32 m16n8k8 vs bf16 m16n8k16), to size the
// tensor budget of the chunked WKV kernels.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int KIND>
__global__ void k(float *out, int iters) {
    float d[8][4] = {};
    uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[u][0]), "+f"(d[u][1]), "+f"(d[u][2]), "+f"(d[u][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[u][0]), "+f"(d[u][1]), "+f"(d[u][2]), "+f"(d[u][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0;
    for (int u = 0; u < 8; u++) s += d[u][0] + d[u][1] + d[u][2] + d[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int kind = 0; kind < 2; kind++)
        for (int warps : {4, 8, 16}) {
            int iters = 20000;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (kind == 0) k<0><<<p.multiProcessorCount, warps * 32>>>(out, iters); else k<1><<<p.multiProcessorCount, warps * 32>>>(out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double mmas = (double)p.multiProcessorCount * warps * iters * 8;
            double macs = mmas * (kind == 0 ? 1024.0 : 2048.0);
            printf("%s warps/SM=%2d: %.1f TFLOP/s, %.1f MAC/clk/SM @%d MHz nominal, %.2f mma/clk/SM\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16",
                   warps, 2 * macs / ms / 1e9, macs / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000,
                   mmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
        }
    return 0;
}
