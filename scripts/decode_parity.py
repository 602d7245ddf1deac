"""Greedy-id parity of the decode path against the reference's decode step (north_star: "bit-exact for argmax token
IDs under greedy decode"; round-1 VERDICT missing #3).

The reference loop (model/llm/rwkv_asr_cuda_whisper.py:694-717): forward_batch over the prompt, then per token
`emb(next) -> forward_batch(states) -> sample_logits(top_k=1)`, where forward_batch is the eager ATen chain of
RWKV_Tmix_x070 / RWKV_CMix_x070 / Block (:181-215, :285-326) around the reference's own stateful kernel
RWKV7_BATCH_OP (rwkv7_state_fwd_fp16.cu:9-57).  Here that loop is rebuilt from the same pieces: this repo's
restatement of those lines with every fused kernel switched off and the reference's operation order (core.FUSED = False,
core.EXACT = True: plain ATen, bf16 intermediates, pinned bit for bit to the reference classes on CPU) and the WKV op bound to the UNMODIFIED reference kernel compiled in oracle/_ref
(libref_state_fwd.so).  It is driven teacher-forced with the ids the product path produced: at every step the
reference's argmax must be the id the product path chose next -- if that holds for all steps, the free-running
reference loop generates the identical sequence (induction), which is the north_star's criterion.

compare() returns counts and, for any disagreement, how close the reference's own top-2 logits were there.
"""
from __future__ import annotations

import torch


def reference_step_logits(model, prompt_ids, forced_ids, n_steps):
    """Yields the reference loop's last-position logits [B, V] after the prompt and after each forced token."""
    from oracle import c_oracle as CO
    from rwkvfla.models.utils import Cache
    from rwkvtts_b200 import core, ops
    old_fused, old_exact, old_op = core.FUSED, core.EXACT, ops.RWKV7_BATCH_OP
    # FUSED off + EXACT on = the reference's eager chain operation for operation (tests/test_forward_batch_cpu.py pins
    # that chain bit for bit to the reference's own classes); the recurrence itself is the reference kernel
    core.FUSED, core.EXACT = False, True
    ops.RWKV7_BATCH_OP = lambda state, r, w, k, v, a, b: CO.ref_state_forward(state, r, w, k, v, a, b)
    try:
        with torch.no_grad():
            cache = Cache()
            out = model(input_ids=prompt_ids, past_key_values=cache, use_cache=True, logits_to_keep=1)
            cache = out.past_key_values
            yield out.logits[:, -1].float()
            for t in range(n_steps - 1):
                out = model(input_ids=forced_ids[:, t:t + 1], past_key_values=cache, use_cache=True, logits_to_keep=1)
                cache = out.past_key_values
                yield out.logits[:, -1].float()
    finally:
        core.FUSED, core.EXACT, ops.RWKV7_BATCH_OP = old_fused, old_exact, old_op


def compare(model, prompt_ids, our_new_ids, n_steps):
    """our_new_ids [B, n_steps]: what the product path (generate, CUDA-graph step, fused kernels) produced greedily."""
    from oracle import c_oracle as CO
    if not CO.ref_available():
        return {"error": "oracle/_ref/libref_state_fwd.so not built"}
    B = prompt_ids.shape[0]
    mism, first, worst_gap, gaps = 0, None, 0.0, []
    for t, lg in enumerate(reference_step_logits(model, prompt_ids, our_new_ids, n_steps)):
        ref_next = lg.argmax(-1)
        ours = our_new_ids[:, t]
        bad = ref_next != ours
        if bool(bad.any()):
            top2 = lg.topk(2, dim=-1).values
            gap_ref = (top2[:, 0] - lg.gather(1, ours[:, None]).squeeze(1))[bad]     # how far our id is from the reference's max
            n = int(bad.sum())
            mism += n
            if first is None:
                first = t
            worst_gap = max(worst_gap, float(gap_ref.max()))
            gaps += [float(g) for g in gap_ref[:4]]
    return {"identical": mism == 0, "prompts": B, "steps": n_steps, "ids_compared": B * n_steps, "mismatches": mism,
            "first_mismatch_step": first, "max_reference_logit_gap_at_mismatch": worst_gap, "sample_gaps": gaps[:8],
            "how": "teacher-forced on the product path's ids; reference step = ATen chain (FUSED off) + unmodified "
                   "rwkv7_state_fwd_fp16 kernel (oracle/_ref)"}
