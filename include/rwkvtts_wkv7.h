/* rwkvtts_wkv7.h -- C ABI of the B200-native WKV-7 hot path (librwkvtts_wkv7.so).
 *
 * Drop-in boundary for the native ops of yynil/RWKVTTS (reference paths relative to
 * /root/reference).  Every entry point is what the reference's torch binding for the path
 * would call instead of its own launcher:
 *
 *   rwkvtts_wkv7_forward        replaces cuda_forward   model/llm/cuda/wkv7_op.cpp:5-10
 *                                                       (launcher wkv7_cuda.cu:132-134)
 *   rwkvtts_wkv7_backward       replaces cuda_backward  model/llm/cuda/wkv7_op.cpp:12-19
 *                                                       (launcher wkv7_cuda.cu:135-138)
 *   rwkvtts_wkv7_state_forward  replaces cuda_forward   model/llm/cuda/wkv7s_op.cpp:7-11 and
 *                                                       model/llm/cuda/rwkv7_state_fwd_fp16.cpp:6-10
 *                                                       (launchers wkv7s.cu:59-64,
 *                                                        rwkv7_state_fwd_fp16.cu:59-63)
 *   rwkvtts_wkv7_forward_ex /   the same recurrence with an initial / final state, i.e. the
 *   rwkvtts_wkv7_backward_ex    rwkvfla call chunk_rwkv7(..., initial_state, output_final_state)
 *                               that model/llm/spark_llm.py reaches through RWKV7Attention
 *                               
 *
 * Conventions (all from the reference):
 *   - head size is compile-time 64 (-D_C_=64 / -D_N_=64, rwkv_s2s_single_ffn.py:10-12);
 *   - w,q,k,v,z,a,y,dy and the six gradients are bf16, contiguous [B,T,H,64]
 *     ([B,T,H*64] for the stateful op: same memory); argument order at the op boundary is
 *     (w, q=r, k, v, z=a, a=b) (wkv7_op.cpp:7); `w` is the BlinkDL pre-activation,
 *     decay = exp(-exp(w)).  RANGE CONTRACT of the default (chunked tensor-core) family: the per-step log-decay
 *     -exp(w) is clamped at -1.35, i.e. results equal the reference's for w <= ln(1.35) = 0.30 -- every value the models
 *     produce (w <= -0.5, rwkv_s2s_single_ffn.py:172) and the w = 0 of masked tokens (:176) -- and for larger w the decay
 *     saturates at exp(-1.35) = 0.26 (dw is not masked there).  The sequential family (rwkvtts_set_impl(0)) and the
 *     stateful forward accept any w, like the reference kernels;
 *   - recurrent state is fp32 [B,H,64,64], value-major S[b][h][value][key], updated in place;
 *   - `s` and `sa` are the caller-allocated scratch tensors of WindBackstepping
 *     (rwkv_s2s_single_ffn.py:22-25): s  = B*H*(T/16)*64*64 floats, sa = B*T*H*64 floats.
 *     They are opaque: this library lays its own checkpoints out inside them and only
 *     requires the reference's sizes (query them with rwkvtts_wkv7_scratch_floats).
 *   - no allocation, no host synchronisation, no torch types; kernels are enqueued on
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream, which is what
 *     the reference uses, wkv7_cuda.cu:133).
 *   - return value: 0 on success, a negative RWKVTTS_ERR_* otherwise (the reference aborts
 *     via assert(), wkv7_cuda.cu:136; the Python mirror turns codes into exceptions).
 */
#ifndef RWKVTTS_WKV7_H_
#define RWKVTTS_WKV7_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RWKVTTS_API __attribute__((visibility("default")))
#else
#define RWKVTTS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RWKVTTS_HEAD_SIZE 64
#define RWKVTTS_CHUNK_LEN 16 /* T must be a multiple of this for the training ops */

enum {
    RWKVTTS_OK = 0,
    RWKVTTS_ERR_SHAPE = -1,   /* B,T,H <= 0, T % 16 != 0, C != H*64 */
    RWKVTTS_ERR_NULL = -2,    /* a required pointer is NULL */
    RWKVTTS_ERR_ALIGN = -3,   /* a tensor pointer is not 16-byte aligned */
    RWKVTTS_ERR_CUDA = -4,    /* launch failed; see rwkvtts_last_cuda_error() */
    RWKVTTS_ERR_DEVICE = -5   /* current device is not sm_100 */
};

RWKVTTS_API int rwkvtts_version(void);
RWKVTTS_API const char *rwkvtts_strerror(int code);
/* cudaError_t of the last failed launch on the calling thread (0 if none). */
RWKVTTS_API int rwkvtts_last_cuda_error(void);

/* Kernel family: 1 = chunked tcgen05 tensor-core kernels (default), 0 = sequential scan on the CUDA
 * cores.  The families lay `s`/`sa` out differently, so forward and backward of one autograd node must
 * run under the same setting.  Initial value from env RWKVTTS_WKV7_IMPL ("scan" -> 0). */
RWKVTTS_API int rwkvtts_set_impl(int impl);
RWKVTTS_API int rwkvtts_get_impl(void);

/* Arithmetic of the stateful forward / decode step (rwkvtts_wkv7_state_forward): 0 = fast (default: four lanes per
 * state row, fused-multiply-add forms chosen for speed; outputs agree with the reference kernel to fp32 rounding),
 * 1 = the reference's operation order (rwkv7_state_fwd_fp16.cu:36-52, wkv7s.cu same lines): y and the state are
 * bit-identical to the reference kernels, which is what "identical greedy token ids" needs.  Initial value from env
 * RWKVTTS_STEP_MODE ("exact" -> 1). */
RWKVTTS_API int rwkvtts_set_step_mode(int mode);
RWKVTTS_API int rwkvtts_get_step_mode(void);

/* Process-wide count of CUDA kernels this library has launched (evidence for bench.py's
 * "gpu_launches": every kernel of the path goes through the entry points below). */
RWKVTTS_API long long rwkvtts_kernel_launches(void);

/* Watchdog of the chunked tensor-core kernels.  Every mbarrier hand-off inside them is bounded (4 s; a
 * launch takes < 2 ms): a wait that expires writes {kernel, barrier, parity, block, thread, time} to a
 * pinned host record and traps, so the next CUDA call of the process fails (sticky launch failure) instead
 * of the device spinning forever -- the reference kernels have no such hand-offs (wkv7_cuda.cu uses
 * __syncthreads only).  Formats the record into buf (NUL-terminated, at most n bytes); returns 1 if a
 * record exists, 0 otherwise.  Readable after the CUDA context has been lost. */
RWKVTTS_API int rwkvtts_watchdog_report(char *buf, size_t n);

/* Number of floats the caller must provide for the scratch tensors (reference sizes). */
RWKVTTS_API size_t rwkvtts_wkv7_scratch_floats(int B, int T, int H, size_t *s_floats, size_t *sa_floats);

/* Training forward.  y[b,t,h,:] = S_t q_t, state starts at 0. */
RWKVTTS_API int rwkvtts_wkv7_forward(int B, int T, int H, const void *w, const void *q, const void *k,
                         const void *v, const void *z, const void *a, void *y, float *s,
                         float *sa, void *stream);

/* Forward without the backward's scratch tensors: what the reference runs under torch.no_grad()
 * (evaluation, prefill of prompts through RWKV7Attention with T >= 64).
 * y, optional initial state s0 and final state sT as in rwkvtts_wkv7_forward_ex. */
RWKVTTS_API int rwkvtts_wkv7_forward_infer(int B, int T, int H, const void *w, const void *q, const void *k,
                               const void *v, const void *z, const void *a, void *y, const float *s0,
                               float *sT, void *stream);

/* Training backward: the exact adjoint of rwkvtts_wkv7_forward. */
RWKVTTS_API int rwkvtts_wkv7_backward(int B, int T, int H, const void *w, const void *q, const void *k,
                          const void *v, const void *z, const void *a, const void *dy,
                          const float *s, const float *sa, void *dw, void *dq, void *dk,
                          void *dv, void *dz, void *da, void *stream);

/* Forward / backward with an initial state s0 (may be NULL = zeros) and optional final
 * state sT / incoming final-state gradient dsT / outgoing initial-state gradient ds0
 * (each may be NULL).  fp32 [B,H,64,64] value-major.  The backward takes the forward's final state
 * sT when dsT is given (required by the tcgen05 family, ignored by the scan family). */
RWKVTTS_API int rwkvtts_wkv7_forward_ex(int B, int T, int H, const void *w, const void *q, const void *k,
                            const void *v, const void *z, const void *a, void *y, float *s,
                            float *sa, const float *s0, float *sT, void *stream);
RWKVTTS_API int rwkvtts_wkv7_backward_ex(int B, int T, int H, const void *w, const void *q, const void *k,
                             const void *v, const void *z, const void *a, const void *dy,
                             const float *s, const float *sa, const float *s0, const float *sT,
                             const float *dsT, void *dw, void *dq, void *dk, void *dv,
                             void *dz, void *da, float *ds0, void *stream);

/* Packed variable-length launches: the sequences of a batch back to back as ONE [1, T_total, H, 64] tensor with
 * cu_seqlens (the `_culens` collators the reference's operating points use: data/utils/spark_dataset.py:111-162,
 * utils/multiple_jsonl.py:77-135, :236-311; rwkvfla's chunk_rwkv7(..., cu_seqlens=), SURVEY.md section 8 rows a10 / b).
 * The recurrent state restarts from zero at every boundary; no padding token is computed or moved, sequence lengths
 * need not be multiples of 16.  cu_seqlens: device int32 [N+1] (0 ... T_total); chunk_base: device int32 [N+1], the
 * exclusive prefix sum of ceil(len_i / 16) (where sequence i's checkpoints start inside `s` / `sa`).  Scratch sizes
 * from rwkvtts_wkv7_varlen_scratch_floats (T_total/16 + N chunk slots per head); s = sa = NULL selects the
 * snapshot-free forward.  Chunked tensor-core family only. */
RWKVTTS_API size_t rwkvtts_wkv7_varlen_scratch_floats(int T_total, int H, int N, size_t *s_floats, size_t *sa_floats);
RWKVTTS_API int rwkvtts_wkv7_forward_varlen(int T_total, int H, int N, const int *cu_seqlens, const int *chunk_base,
                                const void *w, const void *q, const void *k, const void *v, const void *z,
                                const void *a, void *y, float *s, float *sa, void *stream);
RWKVTTS_API int rwkvtts_wkv7_backward_varlen(int T_total, int H, int N, const int *cu_seqlens, const int *chunk_base,
                                 const void *w, const void *q, const void *k, const void *v, const void *z,
                                 const void *a, const void *dy, const float *s, const float *sa, void *dw, void *dq,
                                 void *dk, void *dv, void *dz, void *da, void *stream);

/* Stateful forward (prefill for T > 1, the per-token decode op for T == 1).
 * C must equal H*64.  `state` is read, advanced by T steps and written back. */
RWKVTTS_API int rwkvtts_wkv7_state_forward(int B, int T, int C, int H, float *state, const void *r,
                               const void *w, const void *k, const void *v, const void *a,
                               const void *b, void *y, void *stream);

/* Fused Adam / AdamW step on one contiguous shard of the flat parameter space (ZeRO-2 style: every
 * rank updates only the slice it owns).  Replaces the deepspeed.ops.adam.FusedAdam / DeepSpeedCPUAdam
 * step the reference's scripts run through engine.step() (train_spark_rwkv7speech_jsonl.py:195-199,
 * :481-482).  master / exp_avg / exp_avg_sq: fp32 [n]; grad: bf16 or fp32 [n] (multiplied by
 * grad_scale, e.g. the clipping coefficient); param: bf16 or fp32 [n] receives the updated weights.
 * bias_correction1 = 1 - beta1^t, bias_correction2_sqrt = sqrt(1 - beta2^t) (pass 1.0 to disable). */
RWKVTTS_API int rwkvtts_adam_shard(float *master, float *exp_avg, float *exp_avg_sq, const void *grad,
                       int grad_is_bf16, void *param, int param_is_bf16, long long n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int adamw_mode, float bias_correction1,
                       float bias_correction2_sqrt, float grad_scale, void *stream);

/* The engine's form of the same update (rwkvtts_b200/engine.py): the range [0,n) of the rank's shard is a
 * run of `nseg` segments -- seg_end[s] (device, int64, ascending exclusive ends, multiples of 4) and
 * seg_group[s] (device, int32) -- that map to optimizer param groups with their own hyper-parameters,
 * group_hp (HOST, float[ngroups][4] = lr, weight_decay, bias_correction1, sqrt(bias_correction2); the
 * reference's scripts rewrite param_groups[i]['lr'] every step, train_spark_rwkv7speech.py:586-600).
 * `stat` (device, float[2] = {global squared gradient norm, non-finite count}, or NULL) is read on the
 * device: a non-finite gradient skips the update on every rank (and bumps *skipped, device u64, may be NULL),
 * and `clip` > 0 scales the gradient by clip / (norm + 1e-6) when norm > clip (DeepSpeed's
 * gradient_clipping) -- no host synchronisation anywhere in engine.step().  ngroups <= 8. */
RWKVTTS_API int rwkvtts_adam_multi(float *master, float *exp_avg, float *exp_avg_sq, const void *grad,
                       int grad_is_bf16, void *param, int param_is_bf16, long long n, const long long *seg_end,
                       const int *seg_group, int nseg, const float *group_hp, int ngroups, float beta1, float beta2,
                       float eps, int adamw_mode, const float *stat, float clip, unsigned long long *skipped,
                       void *stream);
/* The whole exchange of the ZeRO-2 step fused with the update, over NVLink / NVSwitch peer memory: for the rank's slice
 * [flat_off, flat_off + n) of the flat parameter space, reduce-scatter(AVG) of the W ranks' bf16 gradients -> Adam ->
 * all-gather of the new bf16 parameters, in one kernel and without staging buffers.  grad_ptrs / param_ptrs: HOST
 * arrays of `world` device pointers, every rank's flat gradient / parameter buffer mapped into this process (symmetric
 * memory; the reference gets the same exchange from DeepSpeed ZeRO-2's NCCL reduce_scatter + allgather_partitions,
 * train_spark_rwkv7speech.py:483-516).  mc_grad / mc_param: multicast addresses of the same buffers, or NULL: with them
 * the reduction happens IN the NVSwitch (multimem.ld_reduce) and the parameter store is broadcast by it (multimem.st).
 * master / exp_avg / exp_avg_sq, seg_end / seg_group, group_hp as rwkvtts_adam_multi (indices relative to the slice);
 * stat[1] > 0 skips the step; *norm_sq (device, may be NULL) accumulates the squared norm of the averaged gradient.
 * The caller orders the launch between two cross-rank barriers.  n, flat_off multiples of 8; world <= 8. */
RWKVTTS_API int rwkvtts_adam_p2p(float *master, float *exp_avg, float *exp_avg_sq, const void *const *grad_ptrs,
                     void *const *param_ptrs, const void *mc_grad, void *mc_param, int world, long long flat_off,
                     long long n, const long long *seg_end, const int *seg_group, int nseg, const float *group_hp,
                     int ngroups, float beta1, float beta2, float eps, int adamw_mode, const float *stat, float *norm_sq,
                     unsigned long long *skipped, void *stream);
/* stat[0] += sum of squares of grad[0..n), stat[1] += 1 if any element is non-finite (device float[2],
 * zeroed by the caller; the engine all-reduces it across ranks before rwkvtts_adam_multi). */
RWKVTTS_API int rwkvtts_grad_stat(const void *grad, int grad_is_bf16, long long n, float *stat, void *stream);

/* ---- caller side of the path: batch assembly and the loss head -------------------
 * rwkvtts_embed_rows: the padded [rows, D] bf16 embedding batch of a speech layout in one kernel.  Replaces the
 * per-sample nn.Embedding lookups + torch.cat + pad_sequence of the reference's batch builders
 * (data/utils/spark_dataset.py:163-239, utils/multiple_jsonl.py:4-75, inference/rwkv7speech_inference.py:35-67,
 * model/llm/cosy_llm.py:64-73).  tables: HOST array of ntab (<= 8) device pointers to bf16 [n_i, D] tables;
 * row_src (device, int64 [rows]): (table << 40) | row for every output position, negative = padding (zeros).
 * D % 8 == 0. */
RWKVTTS_API int rwkvtts_embed_rows(const void *const *tables, int ntab, const long long *row_src, long long rows,
                                   int D, void *out, void *stream);
/* rwkvtts_multi_copy: `count` contiguous gradient tensors (HOST arrays: device pointers 16-byte aligned, element offsets
 * into `flat` that are multiples of 16 bytes, element counts) moved into a flat buffer by one launch per 128 tensors;
 * accumulate[i] != 0 adds to what is there (micro-steps after the first).  elem_bytes 2 (bf16) or 4 (fp32).  What
 * DeepSpeed's `contiguous_gradients` does per parameter (train_scripts/train_spark_rwkv7speech.py:483-516), per bucket. */
RWKVTTS_API int rwkvtts_multi_copy(const void *const *srcs, const long long *dst_off, const long long *n, const int *accumulate,
                                   int count, void *flat, int elem_bytes, void *stream);
/* rwkvtts_ce_forward_backward: cross-entropy of a chunk of bf16 logits [rows, ld] (V valid columns, ld % 8 == 0),
 * loss and gradient in one pass, IN PLACE: loss_rows[r] (fp32) and logits[r,:] <- d loss / d logits * *scale_dev.
 * Rows with labels[r] == ignore_index give loss 0 and a zero gradient row.  The middle piece of the fused
 * linear + CE head the reference reaches through rwkvfla's FusedLinearCrossEntropyLoss (model/llm/spark_llm.py:139-158;
 * label_smoothing as third_party/cosyvoice/transformer/label_smoothing_loss.py:68-96 in closed form). */
RWKVTTS_API int rwkvtts_ce_forward_backward(void *logits, long long rows, int V, long long ld, const long long *labels,
                                            long long ignore_index, float label_smoothing, const float *scale_dev,
                                            float *loss_rows, void *stream);

/* ---- whole-model decode step in one persistent kernel -------------------------------------
 * Replaces, per generated token, the ~25 kernels per layer of the reference decode loop
 * (inference/rwkv7speech_inference.py; model/llm/spark_llm.py:54-102; per layer model/llm/rwkv_asr_cuda_whisper.py:181-215,
 * :277-285, :318-326 with the stateful op model/llm/cuda/rwkv7_state_fwd_fp16.cu at T = 1) by ONE cooperative launch that
 * walks all layers, the final norm and the vocabulary head, and optionally samples greedily on the device.
 * All tensors bf16 unless noted; projection weights are nn.Linear layout [out, in] with `in` contiguous, LoRA
 * down-projections [rank, C], LoRA up-projections [C, rank]; per-channel vectors [C].
 *   dims[RWKVTTS_DEC_NDIM]   : B (<= 32), C (= H * 64, <= 2048), H, L, V, F (channel-mix width, multiple of C, <= 8 C),
 *                              ranks of the w, a, v, g LoRAs (multiples of 32, sum <= 512)
 *   eps[2]                   : LayerNorm eps, GroupNorm eps (head_size_divisor^2 * 1e-5 in the reference: 64e-5)
 *   model_ptrs[RWKVTTS_DEC_NMODEL], layer_ptrs[L][RWKVTTS_DEC_NPTR]: device pointers, indices below.  NULL allowed for
 *       LN0_* (no pre-norm), every *_B (norm without bias) and V1 / V2 / V0 (layer 0).  STATE fp32 [B,H,64,64] value-major,
 *       ATT_SHIFT / FFN_SHIFT [B,C]: the recurrent state of the layer, advanced IN PLACE by every step.
 * rwkvtts_decode_workspace_bytes -> bytes of device workspace (0 = unsupported dims); offsets[3] = byte offsets inside it of
 *       the fp32 logits [B, V] the step leaves, the int64 next-token buffer [32] and the int32 finished flags [32].
 * rwkvtts_decode_init   : writes the plan into `workspace` (zeroes it first); call again after any pointer changes.
 * rwkvtts_decode_step   : one token for all B rows.  tok_in (device int64 [B]) or NULL = the token the previous greedy
 *       step chose.  greedy != 0: arg-max on the device (lowest index on ties) with the EOS handling of generate():
 *       ids in eos[n_eos] (host, <= 8) are masked while suppress_eos != 0, rows already finished emit `pad`, a row
 *       finishes when it emits an EOS id; the chosen ids go to the workspace token buffer and tok_out (device, may be NULL).
 * rwkvtts_decode_release: forget the plan of `workspace` (the memory is the caller's). */
enum rwkvtts_decode_dim {
    RWKVTTS_DEC_B = 0, RWKVTTS_DEC_C, RWKVTTS_DEC_H, RWKVTTS_DEC_L, RWKVTTS_DEC_V, RWKVTTS_DEC_F,
    RWKVTTS_DEC_DW, RWKVTTS_DEC_DA, RWKVTTS_DEC_DV, RWKVTTS_DEC_DG, RWKVTTS_DEC_NDIM
};
enum rwkvtts_decode_model_ptr {
    RWKVTTS_DEC_EMB = 0, RWKVTTS_DEC_LN0_W, RWKVTTS_DEC_LN0_B, RWKVTTS_DEC_LNF_W, RWKVTTS_DEC_LNF_B, RWKVTTS_DEC_HEAD,
    RWKVTTS_DEC_NMODEL
};
enum rwkvtts_decode_layer_ptr {
    RWKVTTS_DEC_LN1_W = 0, RWKVTTS_DEC_LN1_B, RWKVTTS_DEC_LN2_W, RWKVTTS_DEC_LN2_B,
    RWKVTTS_DEC_X_R, RWKVTTS_DEC_X_W, RWKVTTS_DEC_X_K, RWKVTTS_DEC_X_V, RWKVTTS_DEC_X_A, RWKVTTS_DEC_X_G,
    RWKVTTS_DEC_W_R, RWKVTTS_DEC_W_K, RWKVTTS_DEC_W_V, RWKVTTS_DEC_W_O,
    RWKVTTS_DEC_W1, RWKVTTS_DEC_W2, RWKVTTS_DEC_W0, RWKVTTS_DEC_A1, RWKVTTS_DEC_A2, RWKVTTS_DEC_A0,
    RWKVTTS_DEC_V1, RWKVTTS_DEC_V2, RWKVTTS_DEC_V0, RWKVTTS_DEC_G1, RWKVTTS_DEC_G2,
    RWKVTTS_DEC_K_K, RWKVTTS_DEC_K_A, RWKVTTS_DEC_R_K, RWKVTTS_DEC_GN_W, RWKVTTS_DEC_GN_B,
    RWKVTTS_DEC_FFN_X_K, RWKVTTS_DEC_FFN_KEY, RWKVTTS_DEC_FFN_VALUE,
    RWKVTTS_DEC_STATE, RWKVTTS_DEC_ATT_SHIFT, RWKVTTS_DEC_FFN_SHIFT, RWKVTTS_DEC_NPTR
};
RWKVTTS_API size_t rwkvtts_decode_workspace_bytes(const int *dims, size_t *offsets);
RWKVTTS_API int rwkvtts_decode_init(const int *dims, const float *eps, const void *const *model_ptrs,
                                    const void *const *layer_ptrs, void *workspace, size_t workspace_bytes, void *stream);
RWKVTTS_API int rwkvtts_decode_step(void *workspace, const long long *tok_in, long long *tok_out, int greedy,
                                    int suppress_eos, const long long *eos, int n_eos, long long pad, void *stream);
/* The same step (logits only) with a time line: stamps (device u64 [n * 513 + 128], n = 8 L + 4; the last 128 words: cycle
 * stamps of CTA 0 at marked points inside layer 1's phases) receives the GPU nanosecond
 * timer of CTA 0 at kernel start ([0]) and after every grid barrier ([1..7 L + 1]: 7 per layer -- ln1, projections, wkv,
 * output projection, ln2, channel-mix key, channel-mix value -- then the final norm); then, per barrier e and CTA c < 256,
 * the arrival time at [n + 256 e + c] and the release time at [257 n + 256 e + c].  A tuning aid: the kernel is bound by
 * latency per phase, and this separates barrier cost from imbalance between CTAs. */
RWKVTTS_API int rwkvtts_decode_step_profile(void *workspace, const long long *tok_in, unsigned long long *stamps, void *stream);
RWKVTTS_API int rwkvtts_decode_release(void *workspace);

/* ---- fused elementwise kernels of the time-mix around the WKV-7 op ---------------------------------------------
 * Replace the ~30 ATen elementwise kernels RWKV_Tmix_x070.forward runs per layer between its GEMMs
 * (model/llm/rwkv_s2s_single_ffn.py:160-195) and the token-shift lerp of RWKV_CMix_x070.forward (:226).
 * Activations bf16 [B,T,C] contiguous, C = H*64; `mask` bf16 [B,T] of 0/1 or NULL (attention_mask, :160,:175-190);
 * per-channel parameters fp32 [C]; parameter gradients fp32, summed deterministically through `scratch`
 * (rwkvtts_tmix_scratch_floats floats, caller-allocated).  Each *_backward is the exact adjoint of its forward.
 * C <= 4096 for the forward entry points, C <= 2048 for the *_backward ones (RWKVTTS_ERR_SHAPE beyond). */

/* scratch floats for a backward with n_params per-channel parameter vectors (6 / 1 shift_mix, 5 prep, 3 out) */
RWKVTTS_API size_t rwkvtts_tmix_scratch_floats(int B, int T, int C, int n_params);

/* out[i] = x + (shift(x) - x) * mix[i], i < n (n = 6: x_r,x_w,x_k,x_v,x_a,x_g :164-169; n = 1: x_k of channel-mix :226).
 * shift(x)[t] = x[t-1], first token from `prev` [B,C] (NULL = zeros, ZeroPad2d :162); x is multiplied by mask first.
 * prev_out [B,C] (or NULL) receives the masked last token = the shift state of the next call (:511); it may be the
 * same buffer as `prev` only for T == 1 (the decode step updates its state in place). */
RWKVTTS_API int rwkvtts_tmix_shift_mix_forward(int B, int T, int C, int n, const void *x, const void *mask,
                                               const void *prev, const float *mix, void *const *out, void *prev_out,
                                               void *stream);
RWKVTTS_API int rwkvtts_tmix_shift_mix_backward(int B, int T, int C, int n, const void *x, const void *mask,
                                                const void *prev, const float *mix, const void *const *dout, void *dx,
                                                float *dmix, float *scratch, void *stream);
/* The same pair for packed (cu_seqlens) batches: seq_first (device, uint8 [B*T], 1 on the first token of every
 * sequence, or NULL = dense) cuts the shift at sequence boundaries -- shift(x)[t] = 0 where seq_first[t] -- so a
 * packed batch gives what the per-sample runs give (no state is carried: prev / prev_out must be NULL with it). */
RWKVTTS_API int rwkvtts_tmix_shift_mix_forward_varlen(int B, int T, int C, int n, const void *x, const void *mask,
                                                      const void *prev, const float *mix, void *const *out,
                                                      void *prev_out, const unsigned char *seq_first, void *stream);
RWKVTTS_API int rwkvtts_tmix_shift_mix_backward_varlen(int B, int T, int C, int n, const void *x, const void *mask,
                                                       const void *prev, const float *mix, const void *const *dout,
                                                       void *dx, float *dmix, float *scratch,
                                                       const unsigned char *seq_first, void *stream);

/* From the projections k, v and the LoRA outputs w_lo = tanh(xw@w1)@w2, a_lo = (xa@a1)@a2, v_lo = (xv@v1)@v2:
 *   w = -softplus(-(w0 + w_lo)) - 0.5 (:172);  a = sigmoid(a0 + a_lo) (:183);
 *   v' = v + (v_first - v) * sigmoid(v0 + v_lo) (:182; v_lo = v_first = NULL on layer 0, then v' = v);
 *   kk = l2-normalise per head (k * k_k) (:186-187);  k' = k * (1 + (a - 1) * k_a) (:189);
 *   WKV operands a_op = -kk, b_op = kk * a (:191);  mask_rwk = 1: w, k, v, kk masked as :175-190 (in-repo stack);
 *   mask_rwk = 0: only kk and v' are masked (rwkvfla's RWKV7Attention masks its input and v).
 * v2 may be NULL when there is neither a mask nor a v residual (v' = v). */
RWKVTTS_API int rwkvtts_tmix_prep_forward(int B, int T, int C, const void *k, const void *v, const void *w_lo,
                                          const void *a_lo, const void *v_lo, const void *v_first, const void *mask,
                                          const float *w0, const float *a0, const float *v0, const float *k_k,
                                          const float *k_a, int mask_rwk, void *w, void *k2, void *v2, void *a_op,
                                          void *b_op, void *stream);
/* dparams: fp32 [5][C] = d w0, d a0, d v0, d k_k, d k_a.  dv / dv_lo / dv_first may be NULL together with dv2. */
RWKVTTS_API int rwkvtts_tmix_prep_backward(int B, int T, int C, const void *k, const void *v, const void *w_lo,
                                           const void *a_lo, const void *v_lo, const void *v_first, const void *mask,
                                           const float *w0, const float *a0, const float *v0, const float *k_k,
                                           const float *k_a, int mask_rwk, const void *dw, const void *dk2, const void *dv2,
                                           const void *da_op, const void *db_op, void *dk, void *dv, void *dw_lo,
                                           void *da_lo, void *dv_lo, void *dv_first, float *dparams, float *scratch,
                                           void *stream);

/* o = (GroupNorm_H(y; ln_w, ln_b, eps) + (sum_head r * k' * r_k) * v') * g  (:192-195, the input of the output
 * projection).  dparams: fp32 [3][C] = d r_k, d ln_w, d ln_b. */
RWKVTTS_API int rwkvtts_tmix_out_forward(int B, int T, int C, const void *y, const void *r, const void *k2,
                                         const void *v2, const void *g, const float *r_k, const float *ln_w,
                                         const float *ln_b, float eps, void *o, void *stream);
RWKVTTS_API int rwkvtts_tmix_out_backward(int B, int T, int C, const void *y, const void *r, const void *k2,
                                          const void *v2, const void *g, const float *r_k, const float *ln_w,
                                          const float *ln_b, float eps, const void *d_o, void *dy, void *dr, void *dk2,
                                          void *dv2, void *dg, float *dparams, float *scratch, void *stream);

/* Residual add + LayerNorm of the block (Block.forward, model/llm/rwkv_s2s_single_ffn.py:251-259; rwkvfla's
 * LayerNorm(x, residual, prenorm=True)): s = x + res (bf16; res / s may be NULL), y = (s - mean) * rstd * w + b (b may be
 * NULL).  rows = B*T, C % 256 == 0.  stats fp32 [rows][2] = (mean, rstd) is what the backward needs besides s (NULL for
 * inference).  Backward: dx (= d res) = LayerNorm adjoint of dy (+ ds, the gradient flowing into the sum directly);
 * dparams fp32 [2][C] = d w, d b; scratch = rwkvtts_tmix_scratch_floats(1, rows, C, 2) floats. */
RWKVTTS_API int rwkvtts_add_layernorm_forward(long long rows, int C, const void *x, const void *res, const float *w,
                                              const float *b, float eps, void *y, void *s, float *stats, void *stream);
RWKVTTS_API int rwkvtts_add_layernorm_backward(long long rows, int C, const void *sum, const float *stats, const float *w,
                                               const void *dy, const void *ds, void *dx, float *dparams, float *scratch,
                                               void *stream);

/* Channel-mix activation y = relu(x)^2 on n bf16 elements (n % 8 == 0) and its adjoint dx = 2 relu(x) dy
 * (RWKV_CMix_x070.forward, model/llm/rwkv_s2s_single_ffn.py:228: `torch.relu(self.key(k)) ** 2`). */
RWKVTTS_API int rwkvtts_sqrelu_forward(long long n, const void *x, void *y, void *stream);
RWKVTTS_API int rwkvtts_sqrelu_backward(long long n, const void *x, const void *dy, void *dx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RWKVTTS_WKV7_H_ */
