// extern "C" entry points of librwkvtts_wkv7.so -- see include/rwkvtts_wkv7.h.
// Argument validation mirrors the reference wrappers' asserts
// (rwkv_s2s_single_ffn.py:19-21,30-31,52-54; wkv7_cuda.cu:136; wkv7s.cu:61).
#include "../../include/rwkvtts_wkv7.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include <cstdio>
#include <mutex>

namespace rwkvtts {
std::atomic<long long> g_kernel_launches{0};

// ---- watchdog record (see tc05.cuh) -------------------------------------------------------------------------
constexpr int kWdEntries = 32, kWdWords = 8 + kWdEntries;      // tc05.cuh: kWatchdogEntries / kWatchdogWords
static unsigned long long *g_wd_host = nullptr;
static std::once_flag g_wd_once;
unsigned long long *watchdog_record() {
    std::call_once(g_wd_once, [] {
        void *p = nullptr;
        if (cudaHostAlloc(&p, kWdWords * sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            g_wd_host = static_cast<unsigned long long *>(p);
            for (int i = 0; i < kWdWords; i++) g_wd_host[i] = 0;
        } else {
            (void)cudaGetLastError();
        }
    });
    return g_wd_host;
}
static std::atomic<unsigned long long> g_wd_installed[4];      // per slot: bit d = installed on device d
bool watchdog_needs_install(int slot, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || slot < 0 || slot >= 4) return false;
    const unsigned long long bit = 1ull << dev;
    if (g_wd_installed[slot].load(std::memory_order_acquire) & bit) return false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return false;
    }
    return !(g_wd_installed[slot].fetch_or(bit, std::memory_order_acq_rel) & bit);
}
cudaError_t launch_scan_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                            const void *a, const void *b, void *y, float *s, float *sa, const float *s0,
                            float *sT, bool save, cudaStream_t st);
cudaError_t launch_step(int B, int H, const void *w, const void *q, const void *k, const void *v, const void *a,
                        const void *b, void *y, float *state, cudaStream_t st);
cudaError_t launch_state_exact(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                               const void *a, const void *b, void *y, float *state, cudaStream_t st);
cudaError_t launch_tc_fwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, void *y, float *ckT, float *sa, const float *s0, float *sT,
                          const int *cu, const int *cbase, cudaStream_t st);
cudaError_t launch_tc_bwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *a, const void *b, const void *dy, const float *ckT, const float *sa,
                          const float *sT, const float *dsT, void *dw, void *dq, void *dk, void *dv, void *da,
                          void *db, float *ds0, const int *cu, const int *cbase, cudaStream_t st);
cudaError_t launch_scan_bwd(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                            const void *a, const void *b, const void *dy, const float *s, const float *sa,
                            const float *dsT, void *dw, void *dq, void *dk, void *dv, void *da, void *db,
                            float *ds0, cudaStream_t st);
cudaError_t launch_adam_shard(float *master, float *m, float *v, const void *grad, int grad_is_bf16, void *param,
                              int param_is_bf16, long long n, float lr, float b1, float b2, float eps, float wd,
                              int adamw, float bc1, float bc2_sqrt, float gscale, cudaStream_t st);
const char *tc_fwd_barrier_name(unsigned off);
const char *tc_bwd_barrier_name(unsigned off);
inline const char *watchdog_barrier_name(unsigned kernel, unsigned off) {
    return kernel == 1 ? tc_fwd_barrier_name(off) : kernel == 2 ? tc_bwd_barrier_name(off)
         : kernel == 3 ? "grid barrier (number = offset)" : "?";
}
cudaError_t launch_adam_multi(float *master, float *m, float *v, const void *grad, int grad_is_bf16, void *param,
                              int param_is_bf16, long long n, const long long *seg_end, const int *seg_group, int nseg,
                              const float *group_hp, int ngroups, float b1, float b2, float eps, int adamw,
                              const float *stat, float clip, unsigned long long *skipped, cudaStream_t st);
cudaError_t launch_grad_stat(const void *grad, int grad_is_bf16, long long n, float *stat, cudaStream_t st);
cudaError_t launch_adam_p2p(float *master, float *m, float *v, const void *const *grad_ptrs, void *const *param_ptrs,
                            const void *mc_grad, void *mc_param, int world, long long flat_off, long long n,
                            const long long *seg_end, const int *seg_group, int nseg, const float *group_hp, int ngroups,
                            float b1, float b2, float eps, int adamw, const float *stat, float *norm_sq,
                            unsigned long long *skipped, cudaStream_t st);
cudaError_t launch_embed_rows(const void *const *tables, int ntab, const long long *row_src, long long rows, int D,
                              void *out, cudaStream_t st);
cudaError_t launch_multi_copy(const void *const *srcs, const long long *dst_off, const long long *n, const int *accumulate,
                              int count, void *flat, int elem_bytes, cudaStream_t st);
cudaError_t launch_ce_fwd_bwd(void *logits, long long rows, int V, long long ld, const long long *labels,
                              long long ignore_index, float label_smoothing, const float *scale_dev, float *loss_rows,
                              cudaStream_t st);
namespace dec {
cudaError_t decode_init(const int *dims, const float *eps, const void *const *mp, const void *const *lp, void *ws, cudaStream_t st);
cudaError_t decode_step(void *ws, const long long *tok_in, long long *tok_out, int greedy, int suppress_eos,
                        const long long *eos, int n_eos, long long pad, unsigned long long *prof, cudaStream_t st);
void decode_release(void *ws);
size_t decode_workspace(const int *dims, size_t *offs);
}  // namespace dec
int tmix_grid(int B, int T, int C, int which);
cudaError_t launch_sqrelu(const void *x, const void *dy, void *out, long n, cudaStream_t st);
cudaError_t launch_add_ln_fwd(long rows, int C, const void *x, const void *res, const float *w, const float *b, float eps,
                              void *y, void *s, float *stats, cudaStream_t st);
cudaError_t launch_add_ln_bwd(long rows, int C, const void *sum, const float *stats, const float *w, const void *dy,
                              const void *ds, void *dx, float *dparams, float *part, cudaStream_t st);
cudaError_t launch_shift_mix_fwd(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                 const float *mix, void *const *out, void *prev_out, const unsigned char *first,
                                 cudaStream_t st);
cudaError_t launch_shift_mix_bwd(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                 const float *mix, const void *const *dout, void *dx, float *dmix, float *part,
                                 const unsigned char *first, cudaStream_t st);
cudaError_t launch_prep_fwd(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                            const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                            const float *v0, const float *k_k, const float *k_a, int mask_rwk, void *w, void *k2, void *v2,
                            void *a_op, void *b_op, cudaStream_t st);
cudaError_t launch_prep_bwd(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                            const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                            const float *v0, const float *k_k, const float *k_a, int mask_rwk, const void *dw,
                            const void *dk2, const void *dv2, const void *da_op, const void *db_op, void *dk, void *dv,
                            void *dw_lo, void *da_lo, void *dv_lo, void *dv_first, float *dparams, float *part,
                            cudaStream_t st);
cudaError_t launch_out_fwd(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                           const void *g_, const float *r_k, const float *ln_w, const float *ln_b, float eps, void *o,
                           cudaStream_t st);
cudaError_t launch_out_bwd(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                           const void *g_, const float *r_k, const float *ln_w, const float *ln_b, float eps,
                           const void *d_o, void *dy, void *dr, void *dk2, void *dv2, void *dg, float *dparams,
                           float *part, cudaStream_t st);
}  // namespace rwkvtts

namespace {
thread_local int g_last_cuda_error = 0;

// Kernel family: 1 = chunked tcgen05 kernels (default), 0 = sequential scan on the CUDA cores.  The two
// families lay the scratch tensors `s` / `sa` out differently (scan: state at chunk ends, the reference's
// convention, un-stepped by the backward; tcgen05: transposed state at chunk starts in the window frame,
// recomputed from by the backward), so the training forward records the family in `s` ... no: the caller
// keeps forward and backward of one autograd node under the same setting (rwkvtts_set_impl).
int initial_impl() {
    const char *e = getenv("RWKVTTS_WKV7_IMPL");
    if (e != nullptr && e[0] == 's') return 0;
    return 1;
}
std::atomic<int> g_impl{initial_impl()};
int initial_step_mode() {
    const char *e = getenv("RWKVTTS_STEP_MODE");
    return (e != nullptr && e[0] == 'e') ? 1 : 0;       // "exact"
}
std::atomic<int> g_step_mode{initial_step_mode()};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_ptrs(std::initializer_list<const void *> ps) {
    for (const void *p : ps) {
        if (p == nullptr) return RWKVTTS_ERR_NULL;
        if (!aligned16(p)) return RWKVTTS_ERR_ALIGN;
    }
    return RWKVTTS_OK;
}

int check_opt(std::initializer_list<const void *> ps) {
    for (const void *p : ps)
        if (p != nullptr && !aligned16(p)) return RWKVTTS_ERR_ALIGN;
    return RWKVTTS_OK;
}

int finish(cudaError_t e) {
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return RWKVTTS_ERR_CUDA;
    }
    return RWKVTTS_OK;
}
}  // namespace

extern "C" {

int rwkvtts_version(void) { return 100; }

const char *rwkvtts_strerror(int code) {
    switch (code) {
        case RWKVTTS_OK: return "ok";
        case RWKVTTS_ERR_SHAPE: return "bad shape (B,T,H must be > 0, T % 16 == 0 for training ops, C == H*64)";
        case RWKVTTS_ERR_NULL: return "required pointer is NULL";
        case RWKVTTS_ERR_ALIGN: return "tensor pointer is not 16-byte aligned";
        case RWKVTTS_ERR_CUDA: return "CUDA launch failed (see rwkvtts_last_cuda_error)";
        case RWKVTTS_ERR_DEVICE: return "current device is not sm_100";
        default: return "unknown error";
    }
}

int rwkvtts_last_cuda_error(void) { return g_last_cuda_error; }

int rwkvtts_set_impl(int impl) {
    if (impl != 0 && impl != 1) return RWKVTTS_ERR_SHAPE;
    g_impl.store(impl);
    return RWKVTTS_OK;
}
int rwkvtts_get_impl(void) { return g_impl.load(); }

int rwkvtts_set_step_mode(int mode) {
    if (mode != 0 && mode != 1) return RWKVTTS_ERR_SHAPE;
    g_step_mode.store(mode);
    return RWKVTTS_OK;
}
int rwkvtts_get_step_mode(void) { return g_step_mode.load(); }

long long rwkvtts_kernel_launches(void) { return rwkvtts::g_kernel_launches.load(); }

int rwkvtts_watchdog_report(char *buf, size_t n) {
    const unsigned long long *r = rwkvtts::g_wd_host;
    if (r == nullptr || r[0] == 0) {
        if (buf != nullptr && n > 0) buf[0] = 0;
        return 0;
    }
    if (buf != nullptr && n > 0) {
        size_t off = (size_t)snprintf(buf, n, "rwkvtts watchdog: hand-offs that never arrived (one waiter per barrier):");
        for (int i = 0; i < rwkvtts::kWdEntries && off + 1 < n; i++) {
            const unsigned long long e = r[8 + i];
            if ((e >> 63) == 0) continue;
            const unsigned kid = (unsigned)((e >> 60) & 7u), smem_off = (unsigned)(e & 0xffffffu);
            off += (size_t)snprintf(buf + off, n - off, " %s: %s (smem +%u) parity %u, block %u warp %u;",
                                    kid == 1 ? "wkv7_tc_fwd" : kid == 2 ? "wkv7_tc_bwd" : kid == 3 ? "decode_step" : "?",
                                    rwkvtts::watchdog_barrier_name(kid, smem_off), smem_off, (unsigned)((e >> 59) & 1u),
                                    (unsigned)((e >> 24) & 0xffffffu), (unsigned)((e >> 48) & 63u));
            if (off >= n) { off = n - 1; break; }
        }
        buf[off < n ? off : n - 1] = 0;
    }
    return 1;
}

size_t rwkvtts_wkv7_scratch_floats(int B, int T, int H, size_t *s_floats, size_t *sa_floats) {
    const size_t s = (size_t)B * H * (T / RWKVTTS_CHUNK_LEN) * RWKVTTS_HEAD_SIZE * RWKVTTS_HEAD_SIZE;
    const size_t sa = (size_t)B * T * H * RWKVTTS_HEAD_SIZE;
    if (s_floats) *s_floats = s;
    if (sa_floats) *sa_floats = sa;
    return s + sa;
}

int rwkvtts_wkv7_forward_ex(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                            const void *z, const void *a, void *y, float *s, float *sa, const float *s0,
                            float *sT, void *stream) {
    if (B <= 0 || T <= 0 || H <= 0 || T % RWKVTTS_CHUNK_LEN != 0) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({w, q, k, v, z, a, y, s, sa})) return rc;
    if (int rc = check_opt({s0, sT})) return rc;
    if (g_impl.load() == 1)
        return finish(rwkvtts::launch_tc_fwd(B, T, H, w, q, k, v, z, a, y, s, sa, s0, sT, nullptr, nullptr,
                                             (cudaStream_t)stream));
    return finish(rwkvtts::launch_scan_fwd(B, T, H, w, q, k, v, z, a, y, s, sa, s0, sT, true,
                                           (cudaStream_t)stream));
}

int rwkvtts_wkv7_forward_infer(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                               const void *z, const void *a, void *y, const float *s0, float *sT, void *stream) {
    if (B <= 0 || T <= 0 || H <= 0 || T % RWKVTTS_CHUNK_LEN != 0) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({w, q, k, v, z, a, y})) return rc;
    if (int rc = check_opt({s0, sT})) return rc;
    if (g_impl.load() == 1)
        return finish(rwkvtts::launch_tc_fwd(B, T, H, w, q, k, v, z, a, y, nullptr, nullptr, s0, sT, nullptr, nullptr,
                                             (cudaStream_t)stream));
    return finish(rwkvtts::launch_scan_fwd(B, T, H, w, q, k, v, z, a, y, nullptr, nullptr, s0, sT, false,
                                           (cudaStream_t)stream));
}

int rwkvtts_wkv7_forward(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                         const void *z, const void *a, void *y, float *s, float *sa, void *stream) {
    return rwkvtts_wkv7_forward_ex(B, T, H, w, q, k, v, z, a, y, s, sa, nullptr, nullptr, stream);
}

int rwkvtts_wkv7_backward_ex(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                             const void *z, const void *a, const void *dy, const float *s, const float *sa,
                             const float *s0, const float *sT, const float *dsT, void *dw, void *dq, void *dk,
                             void *dv, void *dz, void *da, float *ds0, void *stream) {
    if (B <= 0 || T <= 0 || H <= 0 || T % RWKVTTS_CHUNK_LEN != 0) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da})) return rc;
    if (int rc = check_opt({s0, sT, dsT, ds0})) return rc;
    (void)s0;  // states are rebuilt from the snapshots in `s`; s0 is only part of the signature
    if (g_impl.load() == 1) {
        if (dsT != nullptr && sT == nullptr) return RWKVTTS_ERR_NULL;   // the window-end term needs S_T
        return finish(rwkvtts::launch_tc_bwd(B, T, H, w, q, k, v, z, a, dy, s, sa, sT, dsT, dw, dq, dk, dv, dz, da,
                                             ds0, nullptr, nullptr, (cudaStream_t)stream));
    }
    return finish(rwkvtts::launch_scan_bwd(B, T, H, w, q, k, v, z, a, dy, s, sa, dsT, dw, dq, dk, dv, dz, da,
                                           ds0, (cudaStream_t)stream));
}

int rwkvtts_wkv7_backward(int B, int T, int H, const void *w, const void *q, const void *k, const void *v,
                          const void *z, const void *a, const void *dy, const float *s, const float *sa,
                          void *dw, void *dq, void *dk, void *dv, void *dz, void *da, void *stream) {
    return rwkvtts_wkv7_backward_ex(B, T, H, w, q, k, v, z, a, dy, s, sa, nullptr, nullptr, nullptr, dw, dq, dk,
                                    dv, dz, da, nullptr, stream);
}

// ---- packed (cu_seqlens) launches of the chunked kernels -----------------------------------------------------------
size_t rwkvtts_wkv7_varlen_scratch_floats(int T_total, int H, int N, size_t *s_floats, size_t *sa_floats) {
    // every sequence may end in a partial chunk: at most T_total/16 + N chunk slots per head
    const size_t slots = ((size_t)T_total / RWKVTTS_CHUNK_LEN + (size_t)N) * (size_t)H;
    const size_t s = slots * RWKVTTS_HEAD_SIZE * RWKVTTS_HEAD_SIZE, sa = slots * RWKVTTS_CHUNK_LEN * RWKVTTS_HEAD_SIZE;
    if (s_floats) *s_floats = s;
    if (sa_floats) *sa_floats = sa;
    return s + sa;
}

int rwkvtts_wkv7_forward_varlen(int T_total, int H, int N, const int *cu_seqlens, const int *chunk_base, const void *w,
                                const void *q, const void *k, const void *v, const void *z, const void *a, void *y,
                                float *s, float *sa, void *stream) {
    if (T_total <= 0 || H <= 0 || N <= 0) return RWKVTTS_ERR_SHAPE;
    if (cu_seqlens == nullptr || chunk_base == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({w, q, k, v, z, a, y})) return rc;
    if ((s == nullptr) != (sa == nullptr)) return RWKVTTS_ERR_NULL;          // both (training) or neither (no-grad)
    if (int rc = check_opt({s, sa})) return rc;
    return finish(rwkvtts::launch_tc_fwd(N, T_total, H, w, q, k, v, z, a, y, s, sa, nullptr, nullptr, cu_seqlens, chunk_base,
                                         (cudaStream_t)stream));
}

int rwkvtts_wkv7_backward_varlen(int T_total, int H, int N, const int *cu_seqlens, const int *chunk_base, const void *w,
                                 const void *q, const void *k, const void *v, const void *z, const void *a, const void *dy,
                                 const float *s, const float *sa, void *dw, void *dq, void *dk, void *dv, void *dz,
                                 void *da, void *stream) {
    if (T_total <= 0 || H <= 0 || N <= 0) return RWKVTTS_ERR_SHAPE;
    if (cu_seqlens == nullptr || chunk_base == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({w, q, k, v, z, a, dy, s, sa, dw, dq, dk, dv, dz, da})) return rc;
    return finish(rwkvtts::launch_tc_bwd(N, T_total, H, w, q, k, v, z, a, dy, s, sa, nullptr, nullptr, dw, dq, dk, dv, dz,
                                         da, nullptr, cu_seqlens, chunk_base, (cudaStream_t)stream));
}

int rwkvtts_wkv7_state_forward(int B, int T, int C, int H, float *state, const void *r, const void *w,
                               const void *k, const void *v, const void *a, const void *b, void *y,
                               void *stream) {
    if (B <= 0 || T <= 0 || H <= 0 || C != H * RWKVTTS_HEAD_SIZE) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({state, r, w, k, v, a, b, y})) return rc;
    // op-boundary order of the kernels is (w, q=r, k, v, a, b)
    if (g_step_mode.load() == 1)   // the reference's operation order, bit for bit (wkv7_step_exact.cu)
        return finish(rwkvtts::launch_state_exact(B, T, H, w, r, k, v, a, b, y, state, (cudaStream_t)stream));
    if (T == 1)   // the decode step: dedicated streaming kernel
        return finish(rwkvtts::launch_step(B, H, w, r, k, v, a, b, y, state, (cudaStream_t)stream));
    return finish(rwkvtts::launch_scan_fwd(B, T, H, w, r, k, v, a, b, y, nullptr, nullptr, state, state, false,
                                           (cudaStream_t)stream));
}

int rwkvtts_adam_shard(float *master, float *exp_avg, float *exp_avg_sq, const void *grad, int grad_is_bf16,
                       void *param, int param_is_bf16, long long n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int adamw_mode, float bias_correction1, float bias_correction2_sqrt,
                       float grad_scale, void *stream) {
    if (n < 0) return RWKVTTS_ERR_SHAPE;
    if (n == 0) return RWKVTTS_OK;
    if (master == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || grad == nullptr || param == nullptr)
        return RWKVTTS_ERR_NULL;
    return finish(rwkvtts::launch_adam_shard(master, exp_avg, exp_avg_sq, grad, grad_is_bf16, param, param_is_bf16, n,
                                             lr, beta1, beta2, eps, weight_decay, adamw_mode, bias_correction1,
                                             bias_correction2_sqrt, grad_scale, (cudaStream_t)stream));
}

int rwkvtts_adam_multi(float *master, float *exp_avg, float *exp_avg_sq, const void *grad, int grad_is_bf16,
                       void *param, int param_is_bf16, long long n, const long long *seg_end, const int *seg_group,
                       int nseg, const float *group_hp, int ngroups, float beta1, float beta2, float eps, int adamw_mode,
                       const float *stat, float clip, unsigned long long *skipped, void *stream) {
    if (n < 0 || nseg <= 0 || ngroups <= 0 || ngroups > 8) return RWKVTTS_ERR_SHAPE;
    if (n == 0) return RWKVTTS_OK;
    if (group_hp == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({master, exp_avg, exp_avg_sq, seg_end, seg_group})) return rc;
    if (grad == nullptr || param == nullptr) return RWKVTTS_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(grad) & 7u) || (reinterpret_cast<uintptr_t>(param) & 7u)) return RWKVTTS_ERR_ALIGN;
    return finish(rwkvtts::launch_adam_multi(master, exp_avg, exp_avg_sq, grad, grad_is_bf16, param, param_is_bf16, n,
                                             seg_end, seg_group, nseg, group_hp, ngroups, beta1, beta2, eps, adamw_mode,
                                             stat, clip, skipped, (cudaStream_t)stream));
}

int rwkvtts_adam_p2p(float *master, float *exp_avg, float *exp_avg_sq, const void *const *grad_ptrs, void *const *param_ptrs,
                     const void *mc_grad, void *mc_param, int world, long long flat_off, long long n,
                     const long long *seg_end, const int *seg_group, int nseg, const float *group_hp, int ngroups, float beta1,
                     float beta2, float eps, int adamw_mode, const float *stat, float *norm_sq, unsigned long long *skipped,
                     void *stream) {
    if (n < 0 || n % 8 != 0 || flat_off < 0 || flat_off % 8 != 0 || nseg <= 0 || ngroups <= 0 || ngroups > 8 || world < 1 ||
        world > 8)
        return RWKVTTS_ERR_SHAPE;
    if (n == 0) return RWKVTTS_OK;
    if (grad_ptrs == nullptr || param_ptrs == nullptr || group_hp == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({master, exp_avg, exp_avg_sq, seg_end, seg_group})) return rc;
    for (int r = 0; r < world; r++)
        if (int rc = check_ptrs({grad_ptrs[r], param_ptrs[r]})) return rc;
    if ((mc_grad == nullptr) != (mc_param == nullptr)) return RWKVTTS_ERR_NULL;
    return finish(rwkvtts::launch_adam_p2p(master, exp_avg, exp_avg_sq, grad_ptrs, param_ptrs, mc_grad, mc_param, world,
                                           flat_off, n, seg_end, seg_group, nseg, group_hp, ngroups, beta1, beta2, eps,
                                           adamw_mode, stat, norm_sq, skipped, (cudaStream_t)stream));
}

int rwkvtts_grad_stat(const void *grad, int grad_is_bf16, long long n, float *stat, void *stream) {
    if (n < 0) return RWKVTTS_ERR_SHAPE;
    if (n == 0) return RWKVTTS_OK;
    if (grad == nullptr || stat == nullptr) return RWKVTTS_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(grad) & 7u) return RWKVTTS_ERR_ALIGN;
    return finish(rwkvtts::launch_grad_stat(grad, grad_is_bf16, n, stat, (cudaStream_t)stream));
}

int rwkvtts_embed_rows(const void *const *tables, int ntab, const long long *row_src, long long rows, int D, void *out,
                       void *stream) {
    if (ntab <= 0 || ntab > 8 || rows < 0 || D <= 0 || D % 8 != 0) return RWKVTTS_ERR_SHAPE;
    if (rows == 0) return RWKVTTS_OK;
    if (tables == nullptr) return RWKVTTS_ERR_NULL;
    for (int i = 0; i < ntab; i++)
        if (int rc = check_ptrs({tables[i]})) return rc;
    if (int rc = check_ptrs({out})) return rc;
    if (row_src == nullptr) return RWKVTTS_ERR_NULL;
    return finish(rwkvtts::launch_embed_rows(tables, ntab, row_src, rows, D, out, (cudaStream_t)stream));
}

int rwkvtts_multi_copy(const void *const *srcs, const long long *dst_off, const long long *n, const int *accumulate, int count,
                       void *flat, int elem_bytes, void *stream) {
    if (count < 0 || (elem_bytes != 2 && elem_bytes != 4)) return RWKVTTS_ERR_SHAPE;
    if (count == 0) return RWKVTTS_OK;
    if (srcs == nullptr || dst_off == nullptr || n == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({flat})) return rc;
    for (int i = 0; i < count; i++) {
        if (int rc = check_ptrs({srcs[i]})) return rc;
        if (n[i] <= 0 || n[i] > 0x7fffffffll || dst_off[i] < 0 || (dst_off[i] * elem_bytes) % 16 != 0) return RWKVTTS_ERR_SHAPE;
    }
    return finish(rwkvtts::launch_multi_copy(srcs, dst_off, n, accumulate, count, flat, elem_bytes, (cudaStream_t)stream));
}

int rwkvtts_ce_forward_backward(void *logits, long long rows, int V, long long ld, const long long *labels,
                                long long ignore_index, float label_smoothing, const float *scale_dev, float *loss_rows,
                                void *stream) {
    if (rows < 0 || V <= 0 || ld < V || ld % 8 != 0 || label_smoothing < 0.f || label_smoothing >= 1.f)
        return RWKVTTS_ERR_SHAPE;
    if (rows == 0) return RWKVTTS_OK;
    if (int rc = check_ptrs({logits})) return rc;
    if (labels == nullptr || scale_dev == nullptr || loss_rows == nullptr) return RWKVTTS_ERR_NULL;
    return finish(rwkvtts::launch_ce_fwd_bwd(logits, rows, V, ld, labels, ignore_index, label_smoothing, scale_dev,
                                             loss_rows, (cudaStream_t)stream));
}

// ---- whole-model decode step ------------------------------------------------------------------------------------------
size_t rwkvtts_decode_workspace_bytes(const int *dims, size_t *offsets) {
    if (dims == nullptr) return 0;
    return rwkvtts::dec::decode_workspace(dims, offsets);
}

int rwkvtts_decode_init(const int *dims, const float *eps, const void *const *model_ptrs, const void *const *layer_ptrs,
                        void *workspace, size_t workspace_bytes, void *stream) {
    if (dims == nullptr || eps == nullptr || model_ptrs == nullptr || layer_ptrs == nullptr) return RWKVTTS_ERR_NULL;
    const size_t need = rwkvtts::dec::decode_workspace(dims, nullptr);
    if (need == 0 || workspace_bytes < need) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({workspace})) return rc;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return RWKVTTS_ERR_ALIGN;
    for (int i = 0; i < RWKVTTS_DEC_NMODEL; i++) {
        const bool optional = i == RWKVTTS_DEC_LN0_W || i == RWKVTTS_DEC_LN0_B || i == RWKVTTS_DEC_LNF_B;
        if (int rc = optional ? check_opt({model_ptrs[i]}) : check_ptrs({model_ptrs[i]})) return rc;
    }
    const int L = dims[RWKVTTS_DEC_L];
    for (int l = 0; l < L; l++)
        for (int i = 0; i < RWKVTTS_DEC_NPTR; i++) {
            const void *p = layer_ptrs[(size_t)l * RWKVTTS_DEC_NPTR + i];
            const bool optional = i == RWKVTTS_DEC_LN1_B || i == RWKVTTS_DEC_LN2_B || i == RWKVTTS_DEC_GN_B ||
                                  i == RWKVTTS_DEC_V1 || i == RWKVTTS_DEC_V2 || i == RWKVTTS_DEC_V0;
            if (int rc = optional ? check_opt({p}) : check_ptrs({p})) return rc;
        }
    return finish(rwkvtts::dec::decode_init(dims, eps, model_ptrs, layer_ptrs, workspace, (cudaStream_t)stream));
}

int rwkvtts_decode_step(void *workspace, const long long *tok_in, long long *tok_out, int greedy, int suppress_eos,
                        const long long *eos, int n_eos, long long pad, void *stream) {
    if (workspace == nullptr || (n_eos > 0 && eos == nullptr)) return RWKVTTS_ERR_NULL;
    if (n_eos < 0 || n_eos > 8) return RWKVTTS_ERR_SHAPE;
    return finish(rwkvtts::dec::decode_step(workspace, tok_in, tok_out, greedy, suppress_eos, eos, n_eos, pad, nullptr,
                                            (cudaStream_t)stream));
}

int rwkvtts_decode_step_profile(void *workspace, const long long *tok_in, unsigned long long *stamps, void *stream) {
    if (workspace == nullptr || stamps == nullptr) return RWKVTTS_ERR_NULL;
    return finish(rwkvtts::dec::decode_step(workspace, tok_in, nullptr, 0, 0, nullptr, 0, 0, stamps, (cudaStream_t)stream));
}

int rwkvtts_decode_release(void *workspace) {
    rwkvtts::dec::decode_release(workspace);
    return RWKVTTS_OK;
}

// ---- fused time-mix elementwise kernels ------------------------------------------------------------------------
static bool tmix_shape_ok(int B, int T, int C) { return B > 0 && T > 0 && C > 0 && C % RWKVTTS_HEAD_SIZE == 0 && C <= 4096; }
// the adjoint kernels run C/8 threads per row inside CTAs of 256 threads (launch bounds): rows wider than 2048 channels
// cannot launch -- refuse them here instead of failing in the launch
static bool tmix_bwd_shape_ok(int B, int T, int C) { return tmix_shape_ok(B, T, C) && C <= 2048; }

size_t rwkvtts_tmix_scratch_floats(int B, int T, int C, int n_params) {
    if (!tmix_shape_ok(B, T, C) || n_params <= 0) return 0;
    return (size_t)rwkvtts::tmix_grid(B, T, C, 0) * n_params * C;
}

int rwkvtts_tmix_shift_mix_forward_varlen(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                          const float *mix, void *const *out, void *prev_out,
                                          const unsigned char *seq_first, void *stream) {
    if (!tmix_shape_ok(B, T, C) || (n != 1 && n != 6)) return RWKVTTS_ERR_SHAPE;
    if (out == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({x, mix})) return rc;
    for (int i = 0; i < n; i++)
        if (int rc = check_ptrs({out[i]})) return rc;
    if (int rc = check_opt({prev, prev_out})) return rc;
    if (prev_out != nullptr && prev_out == prev && T != 1) return RWKVTTS_ERR_SHAPE;   // in-place state update: decode only
    if (seq_first != nullptr && (prev != nullptr || prev_out != nullptr)) return RWKVTTS_ERR_SHAPE;   // packed: no carried state
    return finish(rwkvtts::launch_shift_mix_fwd(B, T, C, n, x, mask, prev, mix, out, prev_out, seq_first,
                                                (cudaStream_t)stream));
}

int rwkvtts_tmix_shift_mix_forward(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                   const float *mix, void *const *out, void *prev_out, void *stream) {
    return rwkvtts_tmix_shift_mix_forward_varlen(B, T, C, n, x, mask, prev, mix, out, prev_out, nullptr, stream);
}

int rwkvtts_tmix_shift_mix_backward_varlen(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                           const float *mix, const void *const *dout, void *dx, float *dmix, float *scratch,
                                           const unsigned char *seq_first, void *stream) {
    if (!tmix_bwd_shape_ok(B, T, C) || (n != 1 && n != 6)) return RWKVTTS_ERR_SHAPE;
    if (dout == nullptr) return RWKVTTS_ERR_NULL;
    if (int rc = check_ptrs({x, mix, dx, dmix, scratch})) return rc;
    for (int i = 0; i < n; i++)
        if (int rc = check_ptrs({dout[i]})) return rc;
    if (int rc = check_opt({prev})) return rc;
    if (seq_first != nullptr && prev != nullptr) return RWKVTTS_ERR_SHAPE;
    return finish(rwkvtts::launch_shift_mix_bwd(B, T, C, n, x, mask, prev, mix, dout, dx, dmix, scratch, seq_first,
                                                (cudaStream_t)stream));
}

int rwkvtts_tmix_shift_mix_backward(int B, int T, int C, int n, const void *x, const void *mask, const void *prev,
                                    const float *mix, const void *const *dout, void *dx, float *dmix, float *scratch,
                                    void *stream) {
    return rwkvtts_tmix_shift_mix_backward_varlen(B, T, C, n, x, mask, prev, mix, dout, dx, dmix, scratch, nullptr, stream);
}

int rwkvtts_tmix_prep_forward(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                              const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                              const float *v0, const float *k_k, const float *k_a, int mask_rwk, void *w, void *k2,
                              void *v2, void *a_op, void *b_op, void *stream) {
    if (!tmix_shape_ok(B, T, C)) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({k, w_lo, a_lo, w0, a0, k_k, k_a, w, k2, a_op, b_op})) return rc;
    if ((v_lo == nullptr) != (v_first == nullptr) || (v_lo != nullptr && v0 == nullptr)) return RWKVTTS_ERR_NULL;
    if ((v_lo != nullptr || mask != nullptr) && (v == nullptr || v2 == nullptr)) return RWKVTTS_ERR_NULL;
    if (int rc = check_opt({v, v_lo, v_first, v0, v2})) return rc;
    return finish(rwkvtts::launch_prep_fwd(B, T, C, k, v, w_lo, a_lo, v_lo, v_first, mask, w0, a0, v0, k_k, k_a, mask_rwk, w, k2,
                                           v2, a_op, b_op, (cudaStream_t)stream));
}

int rwkvtts_tmix_prep_backward(int B, int T, int C, const void *k, const void *v, const void *w_lo, const void *a_lo,
                               const void *v_lo, const void *v_first, const void *mask, const float *w0, const float *a0,
                               const float *v0, const float *k_k, const float *k_a, int mask_rwk, const void *dw,
                               const void *dk2, const void *dv2, const void *da_op, const void *db_op, void *dk, void *dv,
                               void *dw_lo, void *da_lo, void *dv_lo, void *dv_first, float *dparams, float *scratch,
                               void *stream) {
    if (!tmix_bwd_shape_ok(B, T, C)) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({k, w_lo, a_lo, w0, a0, k_k, k_a, dw, dk2, da_op, db_op, dk, dw_lo, da_lo, dparams, scratch}))
        return rc;
    if ((v_lo == nullptr) != (v_first == nullptr) || (v_lo != nullptr && v0 == nullptr)) return RWKVTTS_ERR_NULL;
    if (dv != nullptr && (dv2 == nullptr || v == nullptr)) return RWKVTTS_ERR_NULL;
    if (v_lo != nullptr && (dv == nullptr || dv_lo == nullptr || dv_first == nullptr)) return RWKVTTS_ERR_NULL;
    if (int rc = check_opt({v, v_lo, v_first, v0, dv2, dv, dv_lo, dv_first})) return rc;
    return finish(rwkvtts::launch_prep_bwd(B, T, C, k, v, w_lo, a_lo, v_lo, v_first, mask, w0, a0, v0, k_k, k_a, mask_rwk, dw, dk2,
                                           dv2, da_op, db_op, dk, dv, dw_lo, da_lo, dv_lo, dv_first, dparams, scratch,
                                           (cudaStream_t)stream));
}

int rwkvtts_tmix_out_forward(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                             const void *g, const float *r_k, const float *ln_w, const float *ln_b, float eps, void *o,
                             void *stream) {
    if (!tmix_shape_ok(B, T, C)) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({y, r, k2, v2, g, r_k, ln_w, ln_b, o})) return rc;
    return finish(rwkvtts::launch_out_fwd(B, T, C, y, r, k2, v2, g, r_k, ln_w, ln_b, eps, o, (cudaStream_t)stream));
}

int rwkvtts_tmix_out_backward(int B, int T, int C, const void *y, const void *r, const void *k2, const void *v2,
                              const void *g, const float *r_k, const float *ln_w, const float *ln_b, float eps,
                              const void *d_o, void *dy, void *dr, void *dk2, void *dv2, void *dg, float *dparams,
                              float *scratch, void *stream) {
    if (!tmix_bwd_shape_ok(B, T, C)) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({y, r, k2, v2, g, r_k, ln_w, ln_b, d_o, dy, dr, dk2, dv2, dg, dparams, scratch})) return rc;
    return finish(rwkvtts::launch_out_bwd(B, T, C, y, r, k2, v2, g, r_k, ln_w, ln_b, eps, d_o, dy, dr, dk2, dv2, dg, dparams,
                                          scratch, (cudaStream_t)stream));
}

int rwkvtts_add_layernorm_forward(long long rows, int C, const void *x, const void *res, const float *w, const float *b,
                                  float eps, void *y, void *s, float *stats, void *stream) {
    if (rows <= 0 || rows > 0x7fffffffLL || C <= 0 || C % 256 != 0 || C > 4096) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({x, w, y})) return rc;
    if (int rc = check_opt({res, b, s, stats})) return rc;
    return finish(rwkvtts::launch_add_ln_fwd((long)rows, C, x, res, w, b, eps, y, s, stats, (cudaStream_t)stream));
}

int rwkvtts_add_layernorm_backward(long long rows, int C, const void *sum, const float *stats, const float *w,
                                   const void *dy, const void *ds, void *dx, float *dparams, float *scratch, void *stream) {
    if (rows <= 0 || rows > 0x7fffffffLL || C <= 0 || C % 256 != 0 || C > 2048) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({sum, stats, w, dy, dx, dparams, scratch})) return rc;
    if (int rc = check_opt({ds})) return rc;
    return finish(rwkvtts::launch_add_ln_bwd((long)rows, C, sum, stats, w, dy, ds, dx, dparams, scratch, (cudaStream_t)stream));
}

int rwkvtts_sqrelu_forward(long long n, const void *x, void *y, void *stream) {
    if (n <= 0 || n % 8 != 0) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({x, y})) return rc;
    return finish(rwkvtts::launch_sqrelu(x, nullptr, y, (long)n, (cudaStream_t)stream));
}

int rwkvtts_sqrelu_backward(long long n, const void *x, const void *dy, void *dx, void *stream) {
    if (n <= 0 || n % 8 != 0) return RWKVTTS_ERR_SHAPE;
    if (int rc = check_ptrs({x, dy, dx})) return rc;
    return finish(rwkvtts::launch_sqrelu(x, dy, dx, (long)n, (cudaStream_t)stream));
}

}  // extern "C"
