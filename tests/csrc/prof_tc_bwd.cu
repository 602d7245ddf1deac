// Profiling harness (not shipped): per-role cycle breakdown of the tcgen05 training pair, CTA 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DRWKVTTS_PROFILE -o prof_tc_bwd prof_tc_bwd.cu ../../rwkvtts_b200/csrc/wkv7_tc_fwd.cu ../../rwkvtts_b200/csrc/wkv7_tc_bwd.cu
#include <cstdio>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace rwkvtts {
std::atomic<long long> g_kernel_launches{0};
unsigned long long *watchdog_record() { return nullptr; }
bool watchdog_needs_install(int, cudaStream_t) { return false; }
extern long long *g_tc_dbg, *g_tcb_dbg;
cudaError_t launch_tc_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                          const int *cu, const int *cbase, cudaStream_t st);
cudaError_t launch_tc_bwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, const void *dy, const float *ckT, const float *sa,
                          const float *sT, const float *dsT, void *dw, void *dq, void *dk, void *dv, void *da,
                          void *db, float *ds0, const int *cu, const int *cbase, cudaStream_t st);
}
int main() {
    int B = 8, T = 4096, H = 16; size_t n = (size_t)B * T * H * 64;
    std::vector<__nv_bfloat16> h(n);
    void *t[14]; for (int i = 0; i < 14; i++) { cudaMalloc(&t[i], n * 2); if (i < 8) { for (size_t j = 0; j < n; j++) h[j] = __float2bfloat16(i == 0 ? -1.0f : 0.05f * ((j * 7 + i) % 13 - 6)); cudaMemcpy(t[i], h.data(), n * 2, cudaMemcpyHostToDevice); } }
    float *s, *sa; cudaMalloc(&s, (size_t)B * H * (T / 16) * 4096 * 4); cudaMalloc(&sa, n * 4);
    long long *dbg, *dbg2; cudaMalloc(&dbg, 32 * 8); cudaMalloc(&dbg2, 16 * 8);
    rwkvtts::g_tc_dbg = dbg; rwkvtts::g_tcb_dbg = dbg2;
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int r = 0; r < 3; r++) {
        cudaMemset(dbg, 0, 256); cudaMemset(dbg2, 0, 128); cudaEventRecord(e0);
        rwkvtts::launch_tc_fwd(B, T, H, t[0], t[1], t[2], t[3], t[4], t[5], t[6], s, sa, nullptr, nullptr, nullptr, nullptr, 0);
        cudaEventRecord(e1);
        rwkvtts::launch_tc_bwd(B, T, H, t[0], t[1], t[2], t[3], t[4], t[5], t[7], s, sa, nullptr, nullptr, t[8], t[9], t[10], t[11], t[12], t[13], nullptr, nullptr, nullptr, 0);
        cudaEventRecord(e2); cudaDeviceSynchronize();
    }
    float ms1, ms2; cudaEventElapsedTime(&ms1, e0, e1); cudaEventElapsedTime(&ms2, e1, e2);
    long long hd[32], hb[16]; cudaMemcpy(hd, dbg, 256, cudaMemcpyDeviceToHost); cudaMemcpy(hb, dbg2, 128, cudaMemcpyDeviceToHost);
    printf("train fwd %.3f ms, bwd %.3f ms (%s)\n", ms1, ms2, cudaGetErrorString(cudaGetLastError()));
    const char *nm[16] = {"A scan+prefix", "A wait slot empty", "A wait nat empty", "A scale+write+load", "B wait a_done", "B gram", "B solve+write",
                          "B end barrier", "M wait full", "M wait win_scaled", "M wait y_free", "M phase 1", "M phase 2 (+s_free)", "E wait y_ready", "E work", ""};
    printf("-- forward (training variant), cycles/chunk\n");
    for (int i = 0; i < 15; i++) printf("%-22s %8.0f\n", nm[i], (double)hd[i] / (T / 16));
    const char *ne[4] = {"E  tmem loads / rescale", "E  Y + U tiles, U to HBM", "E  proxy fence + arrive", "E  group barrier"};
    for (int i = 0; i < 4; i++) printf("%-26s %8.0f\n", ne[i], (double)hd[16 + i] / (T / 16));
    const char *nb[16] = {"A issue loads", "A wait slot empty", "A prescan+scan+tiles", "B wait a_done", "B gram+solve", "M wait full+blob", "M wait prev done / resc / ok_free",
                          "M R1 (+P2a,P3a issue) + wait Z", "M R2a + wait C1 grams + S0", "M late products issue", "C1 window rescale", "(unused)", "C1 wait Z", "C1 Z tiles + grams", "C2 wait out_ready", "C2 outputs"};
    printf("-- backward, cycles/chunk\n");
    for (int i = 0; i < 16; i++) printf("%-24s %8.0f\n", nb[i], (double)hb[i] / (T / 16));
}
