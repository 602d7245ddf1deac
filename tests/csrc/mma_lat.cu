// mma.sync tf32 m16n8k8: throughput vs number of independent accumulator chains per warp (latency probe).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int CH>
__global__ void k(float *out, int iters) {
    float d[CH][4] = {};
    uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < CH; u++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[u][0]), "+f"(d[u][1]), "+f"(d[u][2]), "+f"(d[u][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int u = 0; u < CH; u++) s += d[u][0] + d[u][1] + d[u][2] + d[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> void run(float *out, int warps, int clk) {
    int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k<CH><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("chains/warp=%d warps/SM=%d: %.1f cycles per mma per warp (=> dependent latency if chains=1)\n", CH, warps, ms * 1e-3 * clk * 1e3 / (iters * (double)CH));
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * 4); cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int w : {1, 4, 8}) { run<1>(out, w, p.clockRate); run<2>(out, w, p.clockRate); run<4>(out, w, p.clockRate); run<8>(out, w, p.clockRate); }
}
