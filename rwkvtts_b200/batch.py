"""Batch builders of the Spark layout, the caller side of the hot path.

Same signature and result as the reference's

    create_inputs_and_labels(batch, tokenizer, model, eos_token_id, device)        /root/reference/utils/multiple_jsonl.py:4-75

which loops over the samples in Python and, per sample, builds five index tensors on the device, runs six embedding
lookups, two `torch.cat` and two `torch.full` (about 20 kernels and as many small host-to-device copies per sample,
then `pad_sequence` over the list).  Here the layout of the whole batch is computed once on the host as flat index
arrays, shipped to the device in one copy, and the batch is assembled with one gather per embedding table and one
scatter per table into the zero-initialised padded result (8 kernels for any batch size, autograd-friendly: the
embedding tables receive their gradients through `index_copy`).

Layout per sample (`:34-41`): [tag2, text..., tag0, global..., tag1, semantic..., eos]; labels are -100 on the prefix and
the semantic ids (+ eos) after it (`:46-53`), padding value -100 / 0.0, attention_mask 1 on real positions (`:66-69`).
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np
import torch


def create_inputs_and_labels(batch: Dict[str, Any], tokenizer, model, eos_token_id: int, device) -> Dict[str, torch.Tensor]:
    texts = batch["text"]
    glob = batch["global_tokens"]
    sem = batch["semantic_tokens"]
    text_ids = [tokenizer.encode(t, add_special_tokens=False) for t in texts]
    B = len(texts)
    lens = [1 + len(text_ids[i]) + 1 + len(glob[i]) + 1 + len(sem[i]) + 1 for i in range(B)]
    Tmax = max(lens)
    # flat destination rows (b * Tmax + position) and source ids per embedding table
    dst = {k: [] for k in ("tag", "text", "global", "semantic")}
    ids = {k: [] for k in ("tag", "text", "global", "semantic")}
    labels = np.full((B, Tmax), -100, dtype=np.int64)
    mask = np.zeros((B, Tmax), dtype=np.int64)
    for i in range(B):
        base, nt, ng, ns = i * Tmax, len(text_ids[i]), len(glob[i]), len(sem[i]) + 1
        p_text, p_tag0 = 1, 1 + nt
        p_glob, p_tag1 = p_tag0 + 1, p_tag0 + 1 + ng
        p_sem = p_tag1 + 1
        dst["tag"] += [base, base + p_tag0, base + p_tag1]
        ids["tag"] += [2, 0, 1]
        dst["text"] += range(base + p_text, base + p_text + nt)
        ids["text"] += text_ids[i]
        dst["global"] += range(base + p_glob, base + p_glob + ng)
        ids["global"] += list(glob[i])
        sem_ids = list(sem[i]) + [eos_token_id]
        dst["semantic"] += range(base + p_sem, base + p_sem + ns)
        ids["semantic"] += sem_ids
        labels[i, p_sem:p_sem + ns] = sem_ids
        mask[i, :lens[i]] = 1
    # one host -> device copy for every index array
    order = ("tag", "text", "global", "semantic")
    sizes = [len(ids[k]) for k in order]
    packed = np.concatenate([np.asarray(ids[k], dtype=np.int64) for k in order]
                            + [np.asarray(dst[k], dtype=np.int64) for k in order]
                            + [labels.reshape(-1), mask.reshape(-1)])
    packed_t = torch.from_numpy(packed)
    if torch.device(device).type == "cuda":
        packed_t = packed_t.pin_memory().to(device, non_blocking=True)
    else:
        packed_t = packed_t.to(device)
    cuts = np.cumsum([0] + sizes + sizes + [B * Tmax, B * Tmax])
    part = [packed_t[cuts[j]:cuts[j + 1]] for j in range(len(cuts) - 1)]
    tables = {"tag": model.tts_tag_embedder, "text": model.text_embedder, "global": model.global_embedder,
              "semantic": model.model.embeddings}
    first = tables["semantic"](part[3])
    out = torch.zeros(B * Tmax, first.shape[-1], dtype=first.dtype, device=first.device)
    for j, k in enumerate(order):
        if sizes[j] == 0:
            continue
        emb = first if k == "semantic" else tables[k](part[j])
        out = out.index_copy(0, part[4 + j], emb)
    return {"input_embs": out.view(B, Tmax, -1), "labels": part[8].view(B, Tmax), "attention_mask": part[9].view(B, Tmax)}
