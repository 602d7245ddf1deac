/* CPU oracle (plain C, fp32) for the WKV-7 recurrence -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.  The
 * product (rwkvtts_b200/) never links or calls it.
 *
 * It is a scalar restatement of the reference's native algorithm in the same
 * arithmetic (bf16 in, fp32 state, bf16 out) so that it can stand in as the CPU port
 * of the path:
 *   forward   -> /root/reference/model/llm/cuda/wkv7_cuda.cu:10-52
 *   backward  -> /root/reference/model/llm/cuda/wkv7_cuda.cu:54-130  (incl. the
 *                16-step snapshot + un-stepping scheme, :76-94)
 *   stateful  -> /root/reference/model/llm/cuda/rwkv7_state_fwd_fp16.cu:9-57
 * One (b,h) pair is one unit of work; pairs are spread over OpenMP threads.
 * Head size is fixed at 64 like the reference build (-D_C_=64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define C 64
#define CHUNK 16

static inline float bf2f(uint16_t u) {
    uint32_t x = ((uint32_t)u) << 16;
    float f;
    memcpy(&f, &x, 4);
    return f;
}
/* round-to-nearest-even, like __float2bfloat16_rn (wkv7_cuda.cu:6) */
static inline uint16_t f2bf(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40);
    x += 0x7fffu + ((x >> 16) & 1u);
    return (uint16_t)(x >> 16);
}

int oracle_wkv7_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* wkv7_cuda.cu:10-52.  s: [B,H,T/16,C,C] with s[..., j, i] = state[i][j]; sa: [B,T,H,C]. */
void oracle_wkv7_forward(int B, int T, int H, const uint16_t *w_, const uint16_t *q_,
                         const uint16_t *k_, const uint16_t *v_, const uint16_t *a_,
                         const uint16_t *b_, uint16_t *y_, float *s_, float *sa_) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bh = 0; bh < B * H; bh++) {
        int bb = bh / H, hh = bh % H;
        float *state = (float *)calloc(C * C, sizeof(float)); /* state[i*C+j], i=value */
        float q[C], k[C], w[C], a[C], b[C];
        for (int t = 0; t < T; t++) {
            size_t base = ((size_t)bb * T + t) * H * C + (size_t)hh * C;
            for (int j = 0; j < C; j++) {
                q[j] = bf2f(q_[base + j]);
                w[j] = expf(-expf(bf2f(w_[base + j])));
                k[j] = bf2f(k_[base + j]);
                a[j] = bf2f(a_[base + j]);
                b[j] = bf2f(b_[base + j]);
            }
            for (int i = 0; i < C; i++) {
                float *st = state + i * C;
                float sa = 0.f;
                for (int j = 0; j < C; j++) sa += a[j] * st[j];
                if (sa_) sa_[base + i] = sa;
                float v = bf2f(v_[base + i]);
                float y = 0.f;
                for (int j = 0; j < C; j++) {
                    float s = st[j] * w[j] + sa * b[j] + k[j] * v;
                    st[j] = s;
                    y += s * q[j];
                }
                y_[base + i] = f2bf(y);
            }
            if (s_ && (t + 1) % CHUNK == 0) {
                size_t sb = (((size_t)bb * H + hh) * (T / CHUNK) + t / CHUNK) * C * C;
                for (int i = 0; i < C; i++)
                    for (int j = 0; j < C; j++) s_[sb + (size_t)j * C + i] = state[i * C + j];
            }
        }
        free(state);
    }
}

/* wkv7_cuda.cu:54-130. */
void oracle_wkv7_backward(int B, int T, int H, const uint16_t *w_, const uint16_t *q_,
                          const uint16_t *k_, const uint16_t *v_, const uint16_t *a_,
                          const uint16_t *b_, const uint16_t *dy_, const float *s_,
                          const float *sa_, uint16_t *dw_, uint16_t *dq_, uint16_t *dk_,
                          uint16_t *dv_, uint16_t *da_, uint16_t *db_) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bh = 0; bh < B * H; bh++) {
        int bb = bh / H, hh = bh % H;
        /* stateT[i*C+j] = S[value j][key i]; dstate[i*C+j] = dS[value i][key j];
           dstateT[i*C+j] = dS[value j][key i]  (the three per-thread arrays, :58) */
        float *stateT = (float *)calloc(C * C, sizeof(float));
        float *dstate = (float *)calloc(C * C, sizeof(float));
        float *dstateT = (float *)calloc(C * C, sizeof(float));
        float w[C], q[C], k[C], v[C], a[C], b[C], dy[C], sa[C], wfac[C], dSb[C];
        for (int t = T - 1; t >= 0; t--) {
            size_t base = ((size_t)bb * T + t) * H * C + (size_t)hh * C;
            for (int i = 0; i < C; i++) {
                q[i] = bf2f(q_[base + i]);
                wfac[i] = -expf(bf2f(w_[base + i]));
                w[i] = expf(wfac[i]);
                k[i] = bf2f(k_[base + i]);
                a[i] = bf2f(a_[base + i]);
                b[i] = bf2f(b_[base + i]);
                v[i] = bf2f(v_[base + i]);
                dy[i] = bf2f(dy_[base + i]);
                sa[i] = sa_[base + i];
            }
            if ((t + 1) % CHUNK == 0) {
                size_t sb = (((size_t)bb * H + hh) * (T / CHUNK) + t / CHUNK) * C * C;
                memcpy(stateT, s_ + sb, sizeof(float) * C * C);
            }
            for (int i = 0; i < C; i++) {
                float *sT = stateT + i * C, *ds = dstate + i * C, *dsT = dstateT + i * C;
                float dq = 0.f;
                for (int j = 0; j < C; j++) dq += sT[j] * dy[j];
                dq_[base + i] = f2bf(dq);
                float iwi = 1.0f / w[i];
                for (int j = 0; j < C; j++) {
                    sT[j] = (sT[j] - k[i] * v[j] - b[i] * sa[j]) * iwi;
                    ds[j] += dy[i] * q[j];
                    dsT[j] += q[i] * dy[j];
                }
                float dw = 0, dk = 0, dv = 0, db = 0, dsb = 0;
                for (int j = 0; j < C; j++) {
                    dw += dsT[j] * sT[j];
                    dk += dsT[j] * v[j];
                    dv += ds[j] * k[j];
                    dsb += ds[j] * b[j];
                    db += dsT[j] * sa[j];
                }
                dw_[base + i] = f2bf(dw * w[i] * wfac[i]);
                dk_[base + i] = f2bf(dk);
                dv_[base + i] = f2bf(dv);
                db_[base + i] = f2bf(db);
                dSb[i] = dsb;
            }
            for (int i = 0; i < C; i++) {
                float *sT = stateT + i * C, *ds = dstate + i * C, *dsT = dstateT + i * C;
                float da = 0.f;
                for (int j = 0; j < C; j++) da += sT[j] * dSb[j];
                da_[base + i] = f2bf(da);
                for (int j = 0; j < C; j++) {
                    ds[j] = ds[j] * w[j] + dSb[i] * a[j];
                    dsT[j] = dsT[j] * w[i] + a[i] * dSb[j];
                }
            }
        }
        free(stateT);
        free(dstate);
        free(dstateT);
    }
}

/* rwkv7_state_fwd_fp16.cu:9-57.  state [B,H,C,C] value-major, updated in place;
   r,w,k,v,a,b,y are [B,T,H*C] bf16. */
void oracle_wkv7_state_forward(int B, int T, int H, float *state_, const uint16_t *r_,
                               const uint16_t *w_, const uint16_t *k_, const uint16_t *v_,
                               const uint16_t *a_, const uint16_t *b_, uint16_t *y_) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bh = 0; bh < B * H; bh++) {
        int bb = bh / H, hh = bh % H;
        float *state = state_ + (size_t)bh * C * C;
        float r[C], k[C], w[C], a[C], b[C];
        for (int t = 0; t < T; t++) {
            size_t base = ((size_t)bb * T + t) * H * C + (size_t)hh * C;
            for (int j = 0; j < C; j++) {
                r[j] = bf2f(r_[base + j]);
                w[j] = expf(-expf(bf2f(w_[base + j])));
                k[j] = bf2f(k_[base + j]);
                a[j] = bf2f(a_[base + j]);
                b[j] = bf2f(b_[base + j]);
            }
            for (int i = 0; i < C; i++) {
                float *st = state + i * C;
                float sa = 0.f;
                for (int j = 0; j < C; j++) sa += a[j] * st[j];
                float vv = bf2f(v_[base + i]);
                float y = 0.f;
                for (int j = 0; j < C; j++) {
                    float s = st[j] * w[j] + k[j] * vv + sa * b[j];
                    st[j] = s;
                    y += s * r[j];
                }
                y_[base + i] = f2bf(y);
            }
        }
    }
}
