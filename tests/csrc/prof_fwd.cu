// Profiling harness (not shipped): phase-cycle breakdown of the chunked forward for one CTA.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace rwkvtts {
extern long long *g_dbg;
cudaError_t launch_chunk_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                             const void *a, const void *b, void *y, float *s, const float *s0, float *sT, cudaStream_t st);
}
int main() {
    int B = 8, T = 4096, H = 16; size_t n = (size_t)B * T * H * 64;
    std::vector<__nv_bfloat16> h(n);
    void *t[7]; for (int i = 0; i < 7; i++) { cudaMalloc(&t[i], n * 2); for (size_t j = 0; j < n; j++) h[j] = __float2bfloat16(i == 0 ? -1.0f : 0.05f * ((j * 7 + i) % 13 - 6)); cudaMemcpy(t[i], h.data(), n * 2, cudaMemcpyHostToDevice); }
    float *s; cudaMalloc(&s, (size_t)B * H * (T / 16) * 4096 * 4);
    long long *dbg; cudaMalloc(&dbg, 16 * 8); cudaMemset(dbg, 0, 128);
    rwkvtts::g_dbg = dbg;
    for (int r = 0; r < 2; r++) { cudaMemset(dbg, 0, 128); rwkvtts::launch_chunk_fwd(B, T, H, t[0], t[1], t[2], t[3], t[4], t[5], t[6], s, nullptr, nullptr, 0); cudaDeviceSynchronize(); }
    long long hd[16]; cudaMemcpy(hd, dbg, 128, cudaMemcpyDeviceToHost);
    const char *nm[16] = {"prep P0 scan+scale", "prep P1 gram", "prep P2 solve", "", "state ckpt store", "state S*W,S*Q (32 mma)", "state V-terms+Aqb (12 mma)", "state S update (32 mma)", "state decay+y out", "", "prep total", "prep wait@sync", "state total", "state wait@sync", "", ""};
    for (int i = 0; i < 14; i++) if (nm[i][0]) printf("%-28s %8.0f cycles/chunk\n", nm[i], (double)hd[i] / (T / 16));
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
