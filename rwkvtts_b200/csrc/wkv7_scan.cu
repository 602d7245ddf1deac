// WKV-7 sequential-scan kernels (CUDA cores, fp32 state in registers).
//
// One CTA of 256 threads per (batch, head).  The 64x64 state is spread over the CTA:
// thread (i = tid>>2, p = tid&3) owns row i and the 16 columns {16m + 4p + c}.  Rows are
// reduced with two warp shuffles, so the inner time loop needs no block barrier; inputs are
// staged 16 tokens at a time through shared memory (register prefetch of the next tile
// overlaps the current tile's arithmetic) and the decay exp(-exp(w)) is evaluated once per
// token instead of once per row.
//
// These kernels serve the stateful / decode entry points (any T) and are the fall-back
// scan for the training ops; the chunked tensor-core kernels live in wkv7_chunk.cu.
//
// Math restated from the reference op definition (SURVEY.md section 8, wkv7_cuda.cu:17-42 forward,
// :62-129 backward); the thread decomposition, staging and snapshot layout are ours.
#include "wkv7_common.cuh"

namespace rwkvtts {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// shared tile: [token 16][vector 6][64] fp32; vector order: decay, q, k, a, b, v
struct FwdSmem {
    float in[kChunk][6][kC];
    float y[kChunk][kC];
    float sa[kChunk][kC];
};

template <bool kSave>
__global__ void __launch_bounds__(kThreads)
wkv7_scan_fwd_kernel(int T, int H, const bf16 *__restrict__ w_, const bf16 *__restrict__ q_,
                     const bf16 *__restrict__ k_, const bf16 *__restrict__ v_,
                     const bf16 *__restrict__ a_, const bf16 *__restrict__ b_, bf16 *__restrict__ y_,
                     float *__restrict__ s_, float *__restrict__ sa_, const float *s0_,
                     float *sT_) {   // s0_/sT_ may alias (in-place state)
    __shared__ __align__(16) FwdSmem sm;
    const int bh = blockIdx.x, bb = bh / H, hh = bh % H;
    const int tid = threadIdx.x, i = tid >> 2, p = tid & 3;
    const size_t tok_stride = (size_t)H * kC;
    const size_t base = (size_t)bb * T * tok_stride + (size_t)hh * kC;
    const bf16 *src[6] = {w_, q_, k_, a_, b_, v_};

    float S[16];
    if (s0_ != nullptr) {
        const float *sp = s0_ + ((size_t)bh * kC + i) * kC;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            float4 x = *reinterpret_cast<const float4 *>(sp + 16 * m + 4 * p);
            S[4 * m] = x.x; S[4 * m + 1] = x.y; S[4 * m + 2] = x.z; S[4 * m + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) S[j] = 0.f;
    }

    // tile = 16 tokens x 6 vectors x 8 uint4; element e = tid + 256 r: vec = e>>7, tok = (e>>3)&15, part = e&7
    uint4 pre[3];
    auto prefetch = [&](int t0) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int e = tid + kThreads * r, vec = e >> 7, tok = (e >> 3) & 15, part = e & 7;
            if (t0 + tok < T)
                pre[r] = ldg_nc_v4(src[vec] + base + (size_t)(t0 + tok) * tok_stride + part * 8);
            else
                pre[r] = make_uint4(0, 0, 0, 0);
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int e = tid + kThreads * r, vec = e >> 7, tok = (e >> 3) & 15, part = e & 7;
            float f[8];
            unpack8(pre[r], f);
            if (vec == 0) {
#pragma unroll
                for (int x = 0; x < 8; x++) f[x] = expf(-expf(f[x]));
            }
            float4 *dst = reinterpret_cast<float4 *>(&sm.in[tok][vec][part * 8]);
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    };

    prefetch(0);
    for (int t0 = 0; t0 < T; t0 += kChunk) {
        __syncthreads();                 // previous tile fully consumed (incl. y/sa staging)
        commit();
        __syncthreads();
        if (t0 + kChunk < T) prefetch(t0 + kChunk);
        const int nt = min(kChunk, T - t0);
        for (int tt = 0; tt < nt; tt++) {
            float sa = 0.f;
            float av[16];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                float4 x = *reinterpret_cast<const float4 *>(&sm.in[tt][3][16 * m + 4 * p]);
                av[4 * m] = x.x; av[4 * m + 1] = x.y; av[4 * m + 2] = x.z; av[4 * m + 3] = x.w;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) sa = fmaf(S[j], av[j], sa);
            sa = quad_sum(sa);
            const float vv = sm.in[tt][5][i];
            float y = 0.f;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const float4 d = *reinterpret_cast<const float4 *>(&sm.in[tt][0][16 * m + 4 * p]);
                const float4 q = *reinterpret_cast<const float4 *>(&sm.in[tt][1][16 * m + 4 * p]);
                const float4 k = *reinterpret_cast<const float4 *>(&sm.in[tt][2][16 * m + 4 * p]);
                const float4 b = *reinterpret_cast<const float4 *>(&sm.in[tt][4][16 * m + 4 * p]);
                float s;
                s = fmaf(S[4 * m + 0], d.x, fmaf(sa, b.x, k.x * vv)); S[4 * m + 0] = s; y = fmaf(s, q.x, y);
                s = fmaf(S[4 * m + 1], d.y, fmaf(sa, b.y, k.y * vv)); S[4 * m + 1] = s; y = fmaf(s, q.y, y);
                s = fmaf(S[4 * m + 2], d.z, fmaf(sa, b.z, k.z * vv)); S[4 * m + 2] = s; y = fmaf(s, q.z, y);
                s = fmaf(S[4 * m + 3], d.w, fmaf(sa, b.w, k.w * vv)); S[4 * m + 3] = s; y = fmaf(s, q.w, y);
            }
            y = quad_sum(y);
            if (p == 0) {
                sm.y[tt][i] = y;
                if (kSave) sm.sa[tt][i] = sa;
            }
        }
        if (kSave) {   // snapshot at the end of every 16-token chunk, value-major like the state
            float *sp = s_ + (((size_t)bh * (T / kChunk) + t0 / kChunk) * kC + i) * kC;
#pragma unroll
            for (int m = 0; m < 4; m++)
                *reinterpret_cast<float4 *>(sp + 16 * m + 4 * p) =
                    make_float4(S[4 * m], S[4 * m + 1], S[4 * m + 2], S[4 * m + 3]);
        }
        __syncthreads();
        // coalesced write-out of the staged outputs: 16 tokens x 64 -> 128 B (bf16) per token
        {
            const int tok = tid >> 4, c4 = (tid & 15) * 4;
            if (tok < nt) {
                const float4 yv = *reinterpret_cast<const float4 *>(&sm.y[tok][c4]);
                uint2 pk = make_uint2(pack2(yv.x, yv.y), pack2(yv.z, yv.w));
                *reinterpret_cast<uint2 *>(y_ + base + (size_t)(t0 + tok) * tok_stride + c4) = pk;
                if (kSave)
                    *reinterpret_cast<float4 *>(sa_ + base + (size_t)(t0 + tok) * tok_stride + c4) =
                        *reinterpret_cast<const float4 *>(&sm.sa[tok][c4]);
            }
        }
    }
    if (sT_ != nullptr) {
        float *sp = sT_ + ((size_t)bh * kC + i) * kC;
#pragma unroll
        for (int m = 0; m < 4; m++)
            *reinterpret_cast<float4 *>(sp + 16 * m + 4 * p) =
                make_float4(S[4 * m], S[4 * m + 1], S[4 * m + 2], S[4 * m + 3]);
    }
}

// ---------------------------------------------------------------------------------------
// backward (reverse-time scan; states are un-stepped from the 16-token snapshots exactly as
// the reference does, wkv7_cuda.cu:76-94, so the two agree step for step)
// ---------------------------------------------------------------------------------------
// vector order in shared memory: decay, q, k, a, b, v, dy, sa, wfac(-exp(w))
struct BwdSmem {
    float in[kChunk][9][kC];
    float out[kChunk][6][kC];   // dw, dq, dk, dv, da, db
    float dSb[2][kC];
};

__global__ void __launch_bounds__(kThreads)
wkv7_scan_bwd_kernel(int T, int H, const bf16 *__restrict__ w_, const bf16 *__restrict__ q_,
                     const bf16 *__restrict__ k_, const bf16 *__restrict__ v_,
                     const bf16 *__restrict__ a_, const bf16 *__restrict__ b_,
                     const bf16 *__restrict__ dy_, const float *__restrict__ s_,
                     const float *__restrict__ sa_, const float *__restrict__ dsT_,
                     bf16 *__restrict__ dw_, bf16 *__restrict__ dq_, bf16 *__restrict__ dk_,
                     bf16 *__restrict__ dv_, bf16 *__restrict__ da_, bf16 *__restrict__ db_,
                     float *__restrict__ ds0_) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    const int bh = blockIdx.x, bb = bh / H, hh = bh % H;
    const int tid = threadIdx.x, i = tid >> 2, p = tid & 3;
    const size_t tok_stride = (size_t)H * kC;
    const size_t base = (size_t)bb * T * tok_stride + (size_t)hh * kC;
    const bf16 *src[7] = {w_, q_, k_, a_, b_, v_, dy_};
    bf16 *dst[6] = {dw_, dq_, dk_, dv_, da_, db_};

    // index x = 4m + c  <->  "other" index o(x) = 16m + 4p + c
    float sT[16];    // S[value o(x)][key i]
    float dsT[16];   // dS[value o(x)][key i]
    float ds[16];    // dS[value i][key o(x)]
#pragma unroll
    for (int x = 0; x < 16; x++) sT[x] = 0.f;
    if (dsT_ != nullptr) {
        const float *g = dsT_ + (size_t)bh * kC * kC;
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int o = 16 * m + 4 * p + c;
                ds[4 * m + c] = g[i * kC + o];
                dsT[4 * m + c] = g[o * kC + i];
            }
    } else {
#pragma unroll
        for (int x = 0; x < 16; x++) { ds[x] = 0.f; dsT[x] = 0.f; }
    }

    for (int t0 = T - kChunk; t0 >= 0; t0 -= kChunk) {
        __syncthreads();
        // stage 16 tokens: 7 bf16 vectors (896 uint4) + sa (fp32, 256 float4)
        for (int e = tid; e < 7 * 128; e += kThreads) {
            const int vec = e >> 7, tok = (e >> 3) & 15, part = e & 7;
            float f[8];
            unpack8(ldg_nc_v4(src[vec] + base + (size_t)(t0 + tok) * tok_stride + part * 8), f);
            int slot = vec;
            if (vec == 0) {
                float g[8];
#pragma unroll
                for (int x = 0; x < 8; x++) { g[x] = -expf(f[x]); f[x] = expf(g[x]); }
                float4 *d2 = reinterpret_cast<float4 *>(&sm.in[tok][8][part * 8]);
                d2[0] = make_float4(g[0], g[1], g[2], g[3]);
                d2[1] = make_float4(g[4], g[5], g[6], g[7]);
            }
            float4 *d = reinterpret_cast<float4 *>(&sm.in[tok][slot][part * 8]);
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        {
            const int tok = tid >> 4, c4 = (tid & 15) * 4;
            *reinterpret_cast<float4 *>(&sm.in[tok][7][c4]) =
                *reinterpret_cast<const float4 *>(sa_ + base + (size_t)(t0 + tok) * tok_stride + c4);
        }
        // snapshot at the end of this chunk: S[value][key]
        {
            const float *sp = s_ + ((size_t)bh * (T / kChunk) + t0 / kChunk) * kC * kC;
#pragma unroll
            for (int m = 0; m < 4; m++)
#pragma unroll
                for (int c = 0; c < 4; c++) sT[4 * m + c] = sp[(16 * m + 4 * p + c) * kC + i];
        }
        __syncthreads();

        for (int tt = kChunk - 1; tt >= 0; tt--) {
            const float(*in)[kC] = sm.in[tt];
            const float w_i = in[0][i], q_i = in[1][i], k_i = in[2][i], a_i = in[3][i], b_i = in[4][i];
            const float dy_i = in[6][i];
            float vv[16], dyv[16], sav[16];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int o = 16 * m + 4 * p;
                float4 x;
                x = *reinterpret_cast<const float4 *>(&in[5][o]);
                vv[4 * m] = x.x; vv[4 * m + 1] = x.y; vv[4 * m + 2] = x.z; vv[4 * m + 3] = x.w;
                x = *reinterpret_cast<const float4 *>(&in[6][o]);
                dyv[4 * m] = x.x; dyv[4 * m + 1] = x.y; dyv[4 * m + 2] = x.z; dyv[4 * m + 3] = x.w;
                x = *reinterpret_cast<const float4 *>(&in[7][o]);
                sav[4 * m] = x.x; sav[4 * m + 1] = x.y; sav[4 * m + 2] = x.z; sav[4 * m + 3] = x.w;
            }
            // dq[key i] = sum_value S_t[value][i] dy[value]
            float dq = 0.f;
#pragma unroll
            for (int x = 0; x < 16; x++) dq = fmaf(sT[x], dyv[x], dq);
            // un-step S_t -> S_{t-1}; add dy q^T into both views of dS
            const float iw = 1.0f / w_i;
            float dw = 0.f, dk = 0.f, db = 0.f, dv = 0.f, dSb = 0.f;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int o = 16 * m + 4 * p;
                const float4 qk = *reinterpret_cast<const float4 *>(&in[1][o]);
                const float4 kk = *reinterpret_cast<const float4 *>(&in[2][o]);
                const float4 bk = *reinterpret_cast<const float4 *>(&in[4][o]);
                const float qa[4] = {qk.x, qk.y, qk.z, qk.w};
                const float ka[4] = {kk.x, kk.y, kk.z, kk.w};
                const float ba[4] = {bk.x, bk.y, bk.z, bk.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int x = 4 * m + c;
                    sT[x] = (sT[x] - k_i * vv[x] - b_i * sav[x]) * iw;
                    ds[x] = fmaf(dy_i, qa[c], ds[x]);
                    dsT[x] = fmaf(q_i, dyv[x], dsT[x]);
                    dw = fmaf(dsT[x], sT[x], dw);
                    dk = fmaf(dsT[x], vv[x], dk);
                    db = fmaf(dsT[x], sav[x], db);
                    dv = fmaf(ds[x], ka[c], dv);
                    dSb = fmaf(ds[x], ba[c], dSb);
                }
            }
            dq = quad_sum(dq); dw = quad_sum(dw); dk = quad_sum(dk);
            db = quad_sum(db); dv = quad_sum(dv); dSb = quad_sum(dSb);
            if (p == 0) {
                sm.out[tt][0][i] = dw * w_i * in[8][i];
                sm.out[tt][1][i] = dq;
                sm.out[tt][2][i] = dk;
                sm.out[tt][3][i] = dv;
                sm.out[tt][5][i] = db;
                sm.dSb[tt & 1][i] = dSb;
            }
            __syncthreads();
            // da[key i] = sum_value S_{t-1}[value][i] dSb[value]; then propagate dS
            float da = 0.f;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int o = 16 * m + 4 * p;
                const float4 g = *reinterpret_cast<const float4 *>(&sm.dSb[tt & 1][o]);
                const float4 wk = *reinterpret_cast<const float4 *>(&in[0][o]);
                const float4 ak = *reinterpret_cast<const float4 *>(&in[3][o]);
                const float ga[4] = {g.x, g.y, g.z, g.w};
                const float wa[4] = {wk.x, wk.y, wk.z, wk.w};
                const float aa[4] = {ak.x, ak.y, ak.z, ak.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int x = 4 * m + c;
                    da = fmaf(sT[x], ga[c], da);
                    ds[x] = fmaf(ds[x], wa[c], dSb * aa[c]);
                    dsT[x] = fmaf(dsT[x], w_i, a_i * ga[c]);
                }
            }
            da = quad_sum(da);
            if (p == 0) sm.out[tt][4][i] = da;
        }
        __syncthreads();
        // write the six gradient tiles: 6 x 16 tokens x 64 -> bf16, 8 B per thread-item
        for (int e = tid; e < 6 * 16 * 16; e += kThreads) {
            const int vec = e >> 8, tok = (e >> 4) & 15, c4 = (e & 15) * 4;
            const float4 g = *reinterpret_cast<const float4 *>(&sm.out[tok][vec][c4]);
            *reinterpret_cast<uint2 *>(dst[vec] + base + (size_t)(t0 + tok) * tok_stride + c4) =
                make_uint2(pack2(g.x, g.y), pack2(g.z, g.w));
        }
    }
    if (ds0_ != nullptr) {
        float *g = ds0_ + ((size_t)bh * kC + i) * kC;
#pragma unroll
        for (int m = 0; m < 4; m++)
            *reinterpret_cast<float4 *>(g + 16 * m + 4 * p) =
                make_float4(ds[4 * m], ds[4 * m + 1], ds[4 * m + 2], ds[4 * m + 3]);
    }
}

// ---------------------------------------------------------------------------------------
// single decode step (T = 1): the per-token op of the AR loop (rwkv7_state_fwd_fp16.cu:9-57 with T = 1;
// RWKV_x070_TMix_one, rwkv_s2s_single_ffn.py:497-502).  33.7 KB per (b,h) -- state read + write -- is all it moves, so
// it is a pure streaming kernel: thread (i, p) owns value row i and keys 16m+4p..16m+4p+3, m < 4 (the scan kernel's
// ownership and summation order, so T single steps are bit-identical to one T-step call); four 16-byte state accesses
// each way, 64 contiguous bytes per quad and access; the six 64-vectors come in through shared memory once per CTA.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wkv7_step_kernel(int H, const bf16 *__restrict__ w, const bf16 *__restrict__ q,
                                                        const bf16 *__restrict__ k, const bf16 *__restrict__ v,
                                                        const bf16 *__restrict__ a, const bf16 *__restrict__ b,
                                                        bf16 *__restrict__ y, float *__restrict__ state) {
    __shared__ float vec[6][kC];     // d, q, k, v, a, b
    const int bh = blockIdx.x, tid = threadIdx.x, i = tid >> 2, p = tid & 3;
    const size_t off = (size_t)bh * kC;                      // [B,1,H,64]: (b*H + h)*64
    float4 *srow = reinterpret_cast<float4 *>(state + (size_t)bh * kC * kC + i * kC + 4 * p);
    float4 s4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) s4[j] = srow[4 * j];         // state first: the long-latency loads
    if (tid < 6 * kC / 2) {                                  // 192 threads x 2 elements
        const int arr = tid / (kC / 2), c = (tid % (kC / 2)) * 2;
        const bf16 *src = arr == 0 ? w : arr == 1 ? q : arr == 2 ? k : arr == 3 ? v : arr == 4 ? a : b;
        const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162 *>(src + off + c);
        float x0 = __bfloat162float(x.x), x1 = __bfloat162float(x.y);
        if (arr == 0) { x0 = expf(-expf(x0)); x1 = expf(-expf(x1)); }              // decay (wkv7_cuda.cu:24)
        vec[arr][c] = x0; vec[arr][c + 1] = x1;
    }
    __syncthreads();
    float S[16];
#pragma unroll
    for (int j = 0; j < 4; j++) { S[4 * j] = s4[j].x; S[4 * j + 1] = s4[j].y; S[4 * j + 2] = s4[j].z; S[4 * j + 3] = s4[j].w; }
    float sa = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) sa = fmaf(S[j], vec[4][16 * (j >> 2) + 4 * p + (j & 3)], sa);
    sa = quad_sum(sa);
    const float vi = vec[3][i];
    float yy = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int c = 16 * (j >> 2) + 4 * p + (j & 3);
        S[j] = fmaf(S[j], vec[0][c], fmaf(sa, vec[5][c], vec[2][c] * vi));
        yy = fmaf(S[j], vec[1][c], yy);
    }
    yy = quad_sum(yy);
#pragma unroll
    for (int j = 0; j < 4; j++) srow[4 * j] = make_float4(S[4 * j], S[4 * j + 1], S[4 * j + 2], S[4 * j + 3]);
    if (p == 0) y[off + i] = __float2bfloat16_rn(yy);
}

cudaError_t launch_step(int B, int H, const void *w, const void *q, const void *k, const void *v, const void *a,
                        const void *b, void *y, float *state, cudaStream_t st) {
    count_launch();
    wkv7_step_kernel<<<dim3(B * H), dim3(256), 0, st>>>(H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k,
                                                         (const bf16 *)v, (const bf16 *)a, (const bf16 *)b, (bf16 *)y, state);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// launchers (called from capi.cu)
// ---------------------------------------------------------------------------------------
cudaError_t launch_scan_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                            const void *a, const void *b, void *y, float *s, float *sa, const float *s0,
                            float *sT, bool save, cudaStream_t st) {
    dim3 grid(B * H), block(kThreads);
    count_launch();
    if (save)
        wkv7_scan_fwd_kernel<true><<<grid, block, 0, st>>>(T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k,
                                                          (const bf16 *)v, (const bf16 *)a, (const bf16 *)b,
                                                          (bf16 *)y, s, sa, s0, sT);
    else
        wkv7_scan_fwd_kernel<false><<<grid, block, 0, st>>>(T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k,
                                                           (const bf16 *)v, (const bf16 *)a, (const bf16 *)b,
                                                           (bf16 *)y, nullptr, nullptr, s0, sT);
    return cudaGetLastError();
}

cudaError_t launch_scan_bwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                            const void *a, const void *b, const void *dy, const float *s, const float *sa,
                            const float *dsT, void *dw, void *dq, void *dk, void *dv, void *da, void *db,
                            float *ds0, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(wkv7_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(BwdSmem));
    if (e != cudaSuccess) return e;
    count_launch();
    wkv7_scan_bwd_kernel<<<dim3(B * H), dim3(kThreads), sizeof(BwdSmem), st>>>(
        T, H, (const bf16 *)w, (const bf16 *)q, (const bf16 *)k, (const bf16 *)v, (const bf16 *)a, (const bf16 *)b,
        (const bf16 *)dy, s, sa, dsT, (bf16 *)dw, (bf16 *)dq, (bf16 *)dk, (bf16 *)dv, (bf16 *)da, (bf16 *)db, ds0);
    return cudaGetLastError();
}

}  // namespace rwkvtts
