// Profiling harness (not shipped): per-role cycle breakdown of the tcgen05 forward, CTA 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DRWKVTTS_PROFILE -o prof_tc_fwd prof_tc_fwd.cu ../../rwkvtts_b200/csrc/wkv7_tc_fwd.cu
#include <cstdio>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace rwkvtts {
std::atomic<long long> g_kernel_launches{0};
unsigned long long *watchdog_record() { return nullptr; }
bool watchdog_needs_install(int, cudaStream_t) { return false; }
extern long long *g_tc_dbg;
cudaError_t launch_tc_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                          const int *cu, const int *cbase, cudaStream_t st);
}
int main() {
    int B = 8, T = 4096, H = 16; size_t n = (size_t)B * T * H * 64;
    std::vector<__nv_bfloat16> h(n);
    void *t[7]; for (int i = 0; i < 7; i++) { cudaMalloc(&t[i], n * 2); for (size_t j = 0; j < n; j++) h[j] = __float2bfloat16(i == 0 ? -1.0f : 0.05f * ((j * 7 + i) % 13 - 6)); cudaMemcpy(t[i], h.data(), n * 2, cudaMemcpyHostToDevice); }
    float *s; cudaMalloc(&s, (size_t)B * H * (T / 16) * 4096 * 4);
    long long *dbg; cudaMalloc(&dbg, 32 * 8);
    rwkvtts::g_tc_dbg = dbg;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; r++) { cudaMemset(dbg, 0, 256); cudaEventRecord(e0); rwkvtts::launch_tc_fwd(B, T, H, t[0], t[1], t[2], t[3], t[4], t[5], t[6], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0); cudaEventRecord(e1); cudaDeviceSynchronize(); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hd[32]; cudaMemcpy(hd, dbg, 256, cudaMemcpyDeviceToHost);
    const char *nm[16] = {"A scan+prefix", "A wait slot empty", "A wait nat empty", "A scale+write+load", "B wait a_done", "B gram", "B solve+write",
                          "B end barrier", "M wait full", "M wait win_scaled", "M wait y_free", "M phase 1", "M phase 2", "E wait y_ready", "E work", ""};
    printf("kernel %.3f ms (%s)\n", ms, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 15; i++) printf("%-22s %8.0f cycles/chunk%s\n", nm[i], (double)hd[i] / (T / 16), (i >= 4 && i < 8) ? "  (x2: group handles every other chunk)" : "");
    const char *ne[4] = {"E  tmem loads / rescale", "E  Y + U tiles, U to HBM", "E  proxy fence + arrive", "E  group barrier"};
    for (int i = 0; i < 4; i++) printf("%-26s %8.0f cycles/chunk\n", ne[i], (double)hd[16 + i] / (T / 16));
}
