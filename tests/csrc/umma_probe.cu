// Probe (not shipped): validates the tcgen05 building blocks the chunked WKV-7 kernels rely on and
// measures their latencies on a B200.
//   T1  SS  A K-major,  B K-major          D = A B^T            (M=64, N=64, K=64, tf32)
//   T2  SS  A MN-major, B MN-major         same product from transposed storage
//   T3  TS  A in tensor memory (M=64 lane layout), B K-major
//   T4  TS->TS chain with no wait in between: D2 = A W^T (N=32), D3 = D2[:, :16] Bm (K=16, N=64)
//   timing: dependent chains, independent MMAs, tcgen05.ld/st round trip
// All inputs are small integers, so every product is exact and compared with ==.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rwkvtts_b200/csrc/tc05.cuh"

using namespace rwkvtts::tc05;

struct Smem {
    float Ak[64 * 64];    // A, K-major canonical
    float Bk[64 * 64];    // B, K-major canonical
    float Amn[64 * 64];   // A, MN-major canonical
    float Bmn[64 * 64];   // B, MN-major canonical
    float Wk[32 * 64];    // W [n=32][k=64], K-major canonical
    float Bm[16 * 64];    // Bm [k=16][n=64], MN-major canonical
    uint64_t bar;
    uint32_t tmem_base;
};

constexpr int LBO_K64 = 64 * 4, SBO_K64 = 32;      // floats: K-major tile with 64 rows
constexpr int SBO_MN64 = 64 * 4, LBO_MN64 = 32;    // floats: MN-major tile with K extent 64
constexpr int LBO_W = 32 * 4, SBO_W = 32;          // K-major, 32 rows
constexpr int LBO_BM = 64 * 4, SBO_BM = 32;        // K-major, 64 rows (n), K extent 16

__device__ void dump(float *out, uint32_t tbase, int col0, int ncol16, int warp, int lane) {
    // full 128 lanes x (16*ncol16) columns -> out[lane][col]
    for (int c = 0; c < ncol16; c++) {
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + col0 + 16 * c, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; i++) out[(warp * 32 + lane) * 64 + 16 * c + i] = v[i];
    }
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, const float *W, const float *Bm,
                                             float *out, long long *cyc) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 64; i += 128) {
        const int r = i >> 6, k = i & 63;
        sm.Ak[kmajor_off(r, k, LBO_K64, SBO_K64)] = A[i];
        sm.Bk[kmajor_off(r, k, LBO_K64, SBO_K64)] = B[i];
        sm.Amn[mnmajor_off(r, k, LBO_MN64, SBO_MN64)] = A[i];
        sm.Bmn[mnmajor_off(r, k, LBO_MN64, SBO_MN64)] = B[i];
    }
    for (int i = tid; i < 32 * 64; i += 128) sm.Wk[kmajor_off(i >> 6, i & 63, LBO_W, SBO_W)] = W[i];
    for (int i = tid; i < 16 * 64; i += 128) sm.Bm[kmajor_off(i & 63, i >> 6, LBO_BM, SBO_BM)] = Bm[i];   // Bm[k][n] -> [n][k] K-major
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
    if (tid == 0) { mbar_init(&sm.bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = sm.tmem_base;
    uint32_t phase = 0;
    const uint32_t I64 = idesc_tf32(64, 64, false, false);
    const uint32_t I64mn = idesc_tf32(64, 64, true, true);
    const uint32_t I32 = idesc_tf32(64, 32, false, false);
    const uint32_t I64bmn = idesc_tf32(64, 64, false, false);

    // ---- T1 ----
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 8; kk++)
            mma_tf32_ss(tb + 0, smem_desc(smem_u32(sm.Ak) + kk * 2 * LBO_K64 * 4, LBO_K64 * 4, SBO_K64 * 4),
                        smem_desc(smem_u32(sm.Bk) + kk * 2 * LBO_K64 * 4, LBO_K64 * 4, SBO_K64 * 4), I64, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    dump(out + 0 * 128 * 64, tb, 0, 4, warp, lane);
    // ---- T2 ----
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 8; kk++)
            mma_tf32_ss(tb + 64, smem_desc(smem_u32(sm.Amn) + kk * LBO_MN64 * 4, LBO_MN64 * 4, SBO_MN64 * 4),
                        smem_desc(smem_u32(sm.Bmn) + kk * LBO_MN64 * 4, LBO_MN64 * 4, SBO_MN64 * 4), I64mn, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    dump(out + 1 * 128 * 64, tb, 64, 4, warp, lane);
    // ---- T3: A -> tensor memory, rows 16w+i in lanes 32w+i (i<16) ----
    fence_before_sync(); __syncthreads(); fence_after_sync();
    for (int c = 0; c < 4; c++) {
        float v[16];
        for (int i = 0; i < 16; i++) v[i] = (lane < 16) ? A[(16 * warp + lane) * 64 + 16 * c + i] : -77.f;
        tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + 128 + 16 * c, v);
    }
    tmem_wait_st();
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 8; kk++)
            mma_tf32_ts(tb + 192, tb + 128 + 8 * kk,
                        smem_desc(smem_u32(sm.Bk) + kk * 2 * LBO_K64 * 4, LBO_K64 * 4, SBO_K64 * 4), I64, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    dump(out + 2 * 128 * 64, tb, 192, 4, warp, lane);
    // ---- T4: D2 = A W^T (N=32) -> cols 256..287 ; D3 = D2[:, 0:16] Bm (K=16, N=64) -> cols 320..383 ----
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 8; kk++)
            mma_tf32_ts(tb + 256, tb + 128 + 8 * kk,
                        smem_desc(smem_u32(sm.Wk) + kk * 2 * LBO_W * 4, LBO_W * 4, SBO_W * 4), I32, kk > 0);
        for (int kk = 0; kk < 2; kk++)
            mma_tf32_ts(tb + 320, tb + 256 + 8 * kk,
                        smem_desc(smem_u32(sm.Bm) + kk * 2 * LBO_BM * 4, LBO_BM * 4, SBO_BM * 4), I64bmn, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    dump(out + 3 * 128 * 64, tb, 256, 2, warp, lane);
    dump(out + 4 * 128 * 64, tb, 320, 4, warp, lane);
    // ---- T4c: same as T4 with a commit + wait between the two groups -> cols 288..319 / 384..447 ----
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 8; kk++)
            mma_tf32_ts(tb + 288, tb + 128 + 8 * kk,
                        smem_desc(smem_u32(sm.Wk) + kk * 2 * LBO_W * 4, LBO_W * 4, SBO_W * 4), I32, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    if (warp == 0 && elect_one()) {
        for (int kk = 0; kk < 2; kk++)
            mma_tf32_ts(tb + 384, tb + 288 + 8 * kk,
                        smem_desc(smem_u32(sm.Bm) + kk * 2 * LBO_BM * 4, LBO_BM * 4, SBO_BM * 4), I64bmn, kk > 0);
        mma_commit(&sm.bar);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    dump(out + 5 * 128 * 64, tb, 384, 4, warp, lane);

    // ---- timing (warp 0, elected lane issues; descriptors are warp-uniform) -----------------------
    fence_before_sync(); __syncthreads(); fence_after_sync();
    constexpr int IT = 256;
    if (warp == 0) {
        const uint64_t dA = smem_desc(smem_u32(sm.Ak), LBO_K64 * 4, SBO_K64 * 4);
        const uint64_t dB = smem_desc(smem_u32(sm.Bk), LBO_K64 * 4, SBO_K64 * 4);
        const uint64_t dW = smem_desc(smem_u32(sm.Wk), LBO_W * 4, SBO_W * 4);
        const uint64_t dBm = smem_desc(smem_u32(sm.Bm), LBO_BM * 4, SBO_BM * 4);
        const uint64_t stepK64 = (2 * LBO_K64 * 4) >> 4, stepW = (2 * LBO_W * 4) >> 4, stepBm = (2 * LBO_BM * 4) >> 4;
        long long t0, t1;
#define TIME_BEGIN() __syncwarp(); t0 = clock64(); if (elect_one()) {
#define TIME_END(slot) mma_commit(&sm.bar); } __syncwarp(); mbar_wait(&sm.bar, phase); phase ^= 1; t1 = clock64(); if (lane == 0) cyc[slot] = (t1 - t0) / IT;
        // (0) dependent chain, no waits: S(cols 128..191) -> D2(256..287) -> S
        TIME_BEGIN()
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
#pragma unroll
            for (int kk = 0; kk < 8; kk++) mma_tf32_ts(tb + 256, tb + 128 + 8 * kk, dW + kk * stepW, I32, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ts(tb + 128, tb + 256 + 8 * kk, dBm + kk * stepBm, I64, true);
        }
        TIME_END(0)
        // (1) three-phase chain as in the planned kernel: N=32 K=64 ; N=16 K=16 ; N=64 K=16 x2 + N=16 K=16
        TIME_BEGIN()
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
#pragma unroll
            for (int kk = 0; kk < 8; kk++) mma_tf32_ts(tb + 256, tb + 128 + 8 * kk, dW + kk * stepW, I32, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ss(tb + 256, dA + kk * stepK64, dW + kk * stepW, I32, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ts(tb + 288, tb + 256 + 8 * kk, dW + kk * stepW, idesc_tf32(64, 16, false, false), kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ts(tb + 128, tb + 288 + 8 * kk, dBm + kk * stepBm, I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ss(tb + 128, dA + kk * stepK64, dBm + kk * stepBm, I64, true);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) mma_tf32_ts(tb + 272, tb + 288 + 8 * kk, dW + kk * stepW, idesc_tf32(64, 16, false, false), true);
        }
        TIME_END(1)
        // (2..6) independent SS MMAs, M=64, N = 16,32,64,128,256 (8 per iteration, two accumulators)
#define INDEP(slot, NN, MM) TIME_BEGIN() \
        _Pragma("unroll 1") for (int it = 0; it < IT; it++) { \
            _Pragma("unroll") for (int kk = 0; kk < 8; kk++) \
                mma_tf32_ss(tb + 256 * (kk & 1), dA + kk * stepK64, dB + kk * stepK64, idesc_tf32(MM, NN, false, false), true); } \
        TIME_END(slot)
        INDEP(2, 16, 64) INDEP(3, 32, 64) INDEP(4, 64, 64) INDEP(5, 128, 64) INDEP(6, 256, 64)
        INDEP(7, 64, 128) INDEP(8, 128, 128) INDEP(9, 256, 128)
        // (10) independent TS N=64
        TIME_BEGIN()
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
#pragma unroll
            for (int kk = 0; kk < 8; kk++) mma_tf32_ts(tb + 384 + 64 * (kk & 1), tb + 128 + 8 * kk, dB + kk * stepK64, I64, true);
        }
        TIME_END(10)
        // (11) one TS N=16 MMA + commit + wait: issue -> completion visible to a thread
        __syncwarp(); t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
            if (elect_one()) { mma_tf32_ts(tb + 384, tb + 128, dW, idesc_tf32(64, 16, false, false), false); mma_commit(&sm.bar); }
            __syncwarp();
            mbar_wait(&sm.bar, phase); phase ^= 1;
        }
        t1 = clock64(); if (lane == 0) cyc[11] = (t1 - t0) / IT;
        // (12) dependent single MMAs back to back (D of one is A of the next), N=16 K=8: pure dependency latency
        TIME_BEGIN()
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
            mma_tf32_ts(tb + 400, tb + 384, dW, idesc_tf32(64, 16, false, false), false);
            mma_tf32_ts(tb + 384, tb + 400, dW, idesc_tf32(64, 16, false, false), false);
        }
        TIME_END(12)
    }
    fence_before_sync(); __syncthreads(); fence_after_sync();
    {   // (13) ld 64 columns -> scale -> st 64 columns, all 4 warps, per-iteration cost
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < IT; it++) {
            float v[4][16];
            for (int c = 0; c < 4; c++) tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + 384 + 16 * c, v[c]);
            tmem_wait_ld();
            for (int c = 0; c < 4; c++) {
                for (int i = 0; i < 16; i++) v[c][i] *= 0.5f;
                tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + 384 + 16 * c, v[c]);
            }
            tmem_wait_st();
        }
        long long t1 = clock64();
        if (tid == 0) cyc[13] = (t1 - t0) / IT;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

static int check(const char *name, const float *out, const std::vector<float> &ref, int ncol) {
    // find, for each logical row m, the lane that holds it
    int bad = 0, found = 0;
    int lane_of[64];
    for (int m = 0; m < 64; m++) {
        lane_of[m] = -1;
        for (int l = 0; l < 128; l++) {
            bool ok = true;
            for (int n = 0; n < ncol && ok; n++) ok = out[l * 64 + n] == ref[m * ncol + n];
            if (ok) { lane_of[m] = l; found++; break; }
        }
        if (lane_of[m] < 0) bad++;
    }
    printf("%-34s rows matched %2d/64 ; row->lane: m0:%d m1:%d m15:%d m16:%d m32:%d m48:%d m63:%d\n", name, found,
           lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[32], lane_of[48], lane_of[63]);
    if (bad) {
        printf("   first rows of output lane 0: ");
        for (int n = 0; n < 8; n++) printf("%g ", out[n]);
        printf(" | expected row 0: ");
        for (int n = 0; n < 8; n++) printf("%g ", ref[n]);
        printf("\n");
    }
    return bad;
}

int main() {
    std::vector<float> A(64 * 64), B(64 * 64), W(32 * 64), Bm(16 * 64);
    srand(1);
    auto rnd = [] { return (float)(rand() % 5 - 2); };
    for (auto &x : A) x = rnd();
    for (auto &x : B) x = rnd();
    for (auto &x : W) x = rnd();
    for (auto &x : Bm) x = rnd();
    std::vector<float> ref1(64 * 64, 0.f), ref4a(64 * 32, 0.f), ref4b(64 * 64, 0.f);
    for (int m = 0; m < 64; m++)
        for (int n = 0; n < 64; n++) {
            float s = 0;
            for (int k = 0; k < 64; k++) s += A[m * 64 + k] * B[n * 64 + k];
            ref1[m * 64 + n] = s;
        }
    for (int m = 0; m < 64; m++)
        for (int n = 0; n < 32; n++) {
            float s = 0;
            for (int k = 0; k < 64; k++) s += A[m * 64 + k] * W[n * 64 + k];
            ref4a[m * 32 + n] = s;
        }
    for (int m = 0; m < 64; m++)
        for (int n = 0; n < 64; n++) {
            float s = 0;
            for (int k = 0; k < 16; k++) s += ref4a[m * 32 + k] * Bm[k * 64 + n];
            ref4b[m * 64 + n] = s;
        }
    float *dA, *dB, *dW, *dBm, *dout; long long *dcyc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dW, W.size() * 4);
    cudaMalloc(&dBm, Bm.size() * 4); cudaMalloc(&dout, 6 * 128 * 64 * 4); cudaMalloc(&dcyc, 16 * 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBm, Bm.data(), Bm.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 6 * 128 * 64 * 4); cudaMemset(dcyc, 0, 128);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    probe<<<1, 128, sizeof(Smem) + 1024>>>(dA, dB, dW, dBm, dout, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> out(6 * 128 * 64);
    long long cyc[16];
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, dcyc, 128, cudaMemcpyDeviceToHost);
    int bad = 0;
    bad += check("T1 SS K-major/K-major", out.data() + 0 * 8192, ref1, 64);
    check("T2 SS MN-major no-swizzle (expected to fail: tf32 MN-major needs SW128_32B)", out.data() + 1 * 8192, ref1, 64);
    bad += check("T3 TS A(tmem) / B K-major", out.data() + 2 * 8192, ref1, 64);
    bad += check("T4a TS N=32", out.data() + 3 * 8192, ref4a, 32);
    bad += check("T4b TS chained (no wait) K=16", out.data() + 4 * 8192, ref4b, 64);
    bad += check("T4c TS chained (commit+wait) K=16", out.data() + 5 * 8192, ref4b, 64);
    const char *nm[14] = {"(0) chain 8xN32 -> 2xN64, no waits", "(1) 3-phase chunk chain (18 MMAs)", "(2) 8 indep SS M64 N16",
                          "(3) 8 indep SS M64 N32", "(4) 8 indep SS M64 N64", "(5) 8 indep SS M64 N128", "(6) 8 indep SS M64 N256",
                          "(7) 8 indep SS M128 N64", "(8) 8 indep SS M128 N128", "(9) 8 indep SS M128 N256",
                          "(10) 8 indep TS M64 N64", "(11) 1 TS N16 + commit + wait", "(12) 2 dependent TS N16 K8",
                          "(13) tmem ld64 + scale + st64"};
    for (int i = 0; i < 14; i++) printf("%-38s %6lld cycles/iter\n", nm[i], cyc[i]);
    printf("%s\n", bad ? "PROBE FAILED" : "PROBE OK");
    return bad ? 2 : 0;
}
