"""Token sampling of the Cosy-layout decode loop: same names, arguments and results as

    ras_sampling, nucleus_sampling, random_sampling       /root/reference/third_party/cosyvoice/utils/common.py:106-137

which `RWKV7LM.inference` (model/llm/llm.py:242-264) calls once per generated token through `sampling_ids`.  The
reference's nucleus sampling walks the sorted probabilities element by element in Python -- every `cum_prob < top_p`
on a device tensor is a device-to-host sync, up to `top_k` = 25 of them plus 2 x 25 tiny tensor constructions per token;
here the kept prefix is computed with one cumulative sum and one host read of its length (`exact_stream=False`: none).  The
repetition check of `ras_sampling` keeps its one host decision.  The random-number stream is consumed exactly as in the
reference: on the same generator state the functions return the same token ids and leave the same generator state
(tests/test_sampling.py, CPU)."""
from __future__ import annotations

import torch


def nucleus_sampling(weighted_scores: torch.Tensor, top_p: float = 0.8, top_k: int = 25, exact_stream: bool = True) -> torch.Tensor:
    """weighted_scores [V] (log-probabilities or logits).  Keeps the most probable tokens while the probability mass
    collected BEFORE a token is < top_p and fewer than top_k are kept (:115-123), then draws one of them."""
    sorted_value, sorted_idx = weighted_scores.softmax(dim=0).sort(descending=True, stable=True)
    k = min(int(top_k), sorted_value.numel())
    head = sorted_value[:k]
    # the reference accumulates `cum_prob += sorted_value[i]` from a Python float 0.0: a running fp32 sum
    before = torch.cumsum(head, 0) - head
    keep = before < top_p
    keep[0] = True                      # 0.0 < top_p for any positive top_p: the first token is always kept
    if exact_stream:
        # torch.multinomial draws one random number per ELEMENT of its input, so only an input of the reference's length
        # leaves the generator where the reference leaves it: one host read of the prefix length (the loop reads the
        # sampled id back every token anyway)
        prob = head[:int(keep.sum())]
    else:
        prob = head * keep              # zero weight outside the kept prefix (`keep` is a prefix): same pick, no sync
    pick = prob.multinomial(1, replacement=True)
    return sorted_idx[:k][pick]


def random_sampling(weighted_scores: torch.Tensor, decoded_tokens, sampling) -> torch.Tensor:
    return weighted_scores.softmax(dim=0).multinomial(1, replacement=True)


def ras_sampling(weighted_scores: torch.Tensor, decoded_tokens, sampling, top_p: float = 0.8, top_k: int = 25,
                 win_size: int = 10, tau_r: float = 0.1) -> torch.Tensor:
    """Repetition-aware sampling (VALL-E 2): nucleus sample; if that token already appears at least win_size * tau_r
    times among the last win_size decoded tokens, resample from the full distribution (:107-112)."""
    top_ids = nucleus_sampling(weighted_scores, top_p=top_p, top_k=top_k)
    recent = decoded_tokens[-win_size:]
    if len(recent):
        rep_num = int((torch.tensor(recent, device=weighted_scores.device) == top_ids).sum())
        if rep_num >= win_size * tau_r:
            top_ids = random_sampling(weighted_scores, decoded_tokens, sampling)
    return top_ids
