"""Batch builders of the Spark layout, the caller side of the hot path.

Same signature and result as the reference's

    create_inputs_and_labels(batch, tokenizer, model, eos_token_id, device)        /root/reference/utils/multiple_jsonl.py:4-75

which loops over the samples in Python and, per sample, builds five index tensors on the device, runs six embedding
lookups, two `torch.cat` and two `torch.full` (about 20 kernels and as many small host-to-device copies per sample,
then `pad_sequence` over the list).  Here the layout of the whole batch is computed once on the host as flat index
arrays, shipped to the device in one copy, and the batch is assembled with one gather per embedding table and one
scatter per table into the zero-initialised padded result (8 kernels for any batch size, autograd-friendly: the
embedding tables receive their gradients through the in-place `index_copy_` into a fresh buffer).

Layout per sample (`:34-41`): [tag2, text..., tag0, global..., tag1, semantic..., eos]; labels are -100 on the prefix and
the semantic ids (+ eos) after it (`:46-53`), padding value -100 / 0.0, attention_mask 1 on real positions (`:66-69`).
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np
import torch


def _gather(tables, order, ids, dst, total, row_src=None):
    """The [total, D] batch from per-table (ids, destination rows).  CUDA bf16 tables: ONE kernel writes every row once
    (csrc/gather.cu through fused.embed_rows; `row_src` = (table << 40) | id per row, -1 = padding, built on the device
    when the caller has no host copy).  Otherwise one lookup and one scatter per table into a zero-filled buffer."""
    from . import core, fused
    weights = [tables[k].weight for k in order]
    used = [j for j, k in enumerate(order) if ids[j].numel()]
    if core.FUSED and fused.gather_usable(weights):
        if row_src is None:
            row_src = torch.full((total,), -1, dtype=torch.long, device=weights[0].device)
            for j in used:
                row_src.index_copy_(0, dst[j], ids[j].to(torch.long) + (j << fused.TABLE_SHIFT))
        return fused.embed_rows(weights, row_src, [ids[j] for j in range(len(order))], [dst[j] for j in range(len(order))])
    out = None
    for j in used:
        emb = tables[order[j]](ids[j])
        if out is None:
            out = torch.zeros(total, emb.shape[-1], dtype=emb.dtype, device=emb.device)
        out.index_copy_(0, dst[j], emb)
    return out


def _row_src(order, ids, dst, total):
    """Host twin of the device construction in _gather: int64 [total], (table << 40) | id, -1 for padding."""
    from . import fused
    rs = np.full(total, -1, dtype=np.int64)
    for j, k in enumerate(order):
        if len(ids[k]):
            rs[np.asarray(dst[k], dtype=np.int64)] = np.asarray(ids[k], dtype=np.int64) + (j << fused.TABLE_SHIFT)
    return rs


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
                            + [labels.reshape(-1), mask.reshape(-1), _row_src(order, ids, dst, B * Tmax)])
    packed_t = torch.from_numpy(packed)
    if torch.device(device).type == "cuda":
        packed_t = packed_t.pin_memory().to(device, non_blocking=True)
    else:
        packed_t = packed_t.to(device)
    cuts = np.cumsum([0] + sizes + sizes + [B * Tmax, B * Tmax, B * Tmax])
    part = [packed_t[cuts[j]:cuts[j + 1]] for j in range(len(cuts) - 1)]
    tables = {"tag": model.tts_tag_embedder, "text": model.text_embedder, "global": model.global_embedder,
              "semantic": model.model.embeddings}
    out = _gather(tables, order, part[0:4], part[4:8], B * Tmax, row_src=part[10])
    return {"input_embs": out.view(B, Tmax, -1), "labels": part[8].view(B, Tmax), "attention_mask": part[9].view(B, Tmax)}


def process_single_batch(batch: Dict[str, torch.Tensor], rwkv7speech_model, eos_token_id: int = 8192) -> Dict[str, torch.Tensor]:
    """Same signature and result as the reference's process_single_batch (/root/reference/data/utils/spark_dataset.py:
    165-239), the collator of train_spark_rwkv7speech.py: the batch holds LEFT-padded id matrices with their masks
    (`input_ids`, `global_tokens_ids`, `semantic_tokens_ids` + `*attention_mask*`); the result is the LEFT-padded
    [tag2, text, tag0, global, tag1, semantic] embedding batch, its attention mask, and labels that are already shifted
    (`labels[i, -S-1:-1] = semantic ids, labels[i, -1] = eos`, :231-232).

    The reference reads six lengths per sample with `.item()` (host syncs) and launches ~25 kernels per sample; here the
    three length vectors come back in one transfer, the layout is computed on the host, and the batch is assembled with
    one id gather, one embedding lookup and one scatter per table."""
    model = rwkv7speech_model
    device = model.device
    ids_t, ids_g, ids_s = batch["input_ids"], batch["global_tokens_ids"], batch["semantic_tokens_ids"]
    B = ids_t.shape[0]
    lens = torch.stack([batch["attention_mask_input_ids"].sum(1), batch["global_tokens_attention_mask"].sum(1),
                        batch["semantic_tokens_attention_mask"].sum(1)]).tolist()          # the only host sync
    tl, gl, sl = ([int(v) for v in row] for row in lens)
    total = [tl[i] + gl[i] + sl[i] + 3 for i in range(B)]
    Tmax = max(total)
    src = {k: [] for k in ("text", "global", "semantic")}       # flat positions inside the padded id matrices
    dst = {k: [] for k in ("tag", "text", "global", "semantic")}
    lab_dst, mask = [], np.zeros((B, Tmax), dtype=np.int64)
    for i in range(B):
        pad = Tmax - total[i]
        base = i * Tmax + pad
        mask[i, pad:] = 1
        p_text, p_tag0 = 1, 1 + tl[i]
        p_glob, p_tag1 = p_tag0 + 1, p_tag0 + 1 + gl[i]
        p_sem = p_tag1 + 1
        dst["tag"] += [base, base + p_tag0, base + p_tag1]
        for k, ids, n, p in (("text", ids_t, tl[i], p_text), ("global", ids_g, gl[i], p_glob), ("semantic", ids_s, sl[i], p_sem)):
            L = ids.shape[1]
            src[k] += range(i * L + L - n, i * L + L)              # the last n entries of row i (left padding)
            dst[k] += range(base + p, base + p + n)
        lab_dst += range(i * Tmax + Tmax - sl[i] - 1, i * Tmax + Tmax - 1)
    order = ("text", "global", "semantic")
    arrays = [np.asarray(src[k], dtype=np.int64) for k in order] + [np.asarray(dst[k], dtype=np.int64) for k in ("tag",) + order] \
        + [np.asarray(lab_dst, dtype=np.int64), np.asarray([i * Tmax + Tmax - 1 for i in range(B)], dtype=np.int64), mask.reshape(-1)]
    cuts = np.cumsum([0] + [len(a) for a in arrays])
    packed = torch.from_numpy(np.concatenate(arrays))
    packed = packed.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else packed.to(device)
    part = [packed[cuts[j]:cuts[j + 1]] for j in range(len(arrays))]
    s_text, s_glob, s_sem, d_tag, d_text, d_glob, d_sem, d_lab, d_eos, m_flat = part
    tok = {"text": ids_t.to(device).reshape(-1)[s_text], "global": ids_g.to(device).reshape(-1)[s_glob],
           "semantic": ids_s.to(device).reshape(-1)[s_sem]}
    tables = {"tag": model.tts_tag_embedder, "text": model.text_embedder, "global": model.global_embedder,
              "semantic": model.model.embeddings}
    tag_ids = torch.tensor([2, 0, 1], dtype=torch.long, device=device).repeat(B)
    out = _gather(tables, ("tag", "text", "global", "semantic"), [tag_ids, tok["text"], tok["global"], tok["semantic"]],
                  [d_tag, d_text, d_glob, d_sem], B * Tmax)
    labels = torch.full((B * Tmax,), -100, dtype=torch.long, device=device)
    labels.index_copy_(0, d_lab, tok["semantic"].to(torch.long))
    labels.index_fill_(0, d_eos, eos_token_id)
    return {"input_embs": out.view(B, Tmax, -1), "attention_mask": m_flat.view(B, Tmax), "labels": labels.view(B, Tmax)}


def create_inputs(texts, global_tokens_ids, semantic_tokens_ids, tokenizer, llm, pad_token_id: int = 0):
    """Same signature and result as the reference's create_inputs (/root/reference/inference/rwkv7speech_inference.py:
    35-67), the prompt builder of the AR decode loop and of collate_fn_for_rwkv7speech: LEFT-padded
    [tag2, text, tag0, global, tag1, semantic] embeddings [B, Tmax, D] and their attention mask [B, Tmax].  One host ->
    device copy of the index arrays, one lookup + one scatter per embedding table (the reference: 6 lookups, 2 cats and 6
    small copies per sample)."""
    assert len(texts) == len(global_tokens_ids) == len(semantic_tokens_ids), \
        f"input lists differ in length: texts({len(texts)}), global_tokens_ids({len(global_tokens_ids)}), " \
        f"semantic_tokens_ids({len(semantic_tokens_ids)})"
    device = llm.device
    B = len(texts)
    text_ids = [tokenizer.encode(t) for t in texts]
    total = [len(text_ids[i]) + len(global_tokens_ids[i]) + len(semantic_tokens_ids[i]) + 3 for i in range(B)]
    Tmax = max(total)
    ids = {k: [] for k in ("tag", "text", "global", "semantic")}
    dst = {k: [] for k in ("tag", "text", "global", "semantic")}
    mask = np.zeros((B, Tmax), dtype=np.int64)
    for i in range(B):
        pad = Tmax - total[i]
        base = i * Tmax + pad
        mask[i, pad:] = 1
        nt, ng, ns = len(text_ids[i]), len(global_tokens_ids[i]), len(semantic_tokens_ids[i])
        p_tag0 = 1 + nt
        p_tag1 = p_tag0 + 1 + ng
        ids["tag"] += [2, 0, 1]
        dst["tag"] += [base, base + p_tag0, base + p_tag1]
        for k, seq, n, p in (("text", text_ids[i], nt, 1), ("global", global_tokens_ids[i], ng, p_tag0 + 1),
                             ("semantic", semantic_tokens_ids[i], ns, p_tag1 + 1)):
            ids[k] += list(seq)
            dst[k] += range(base + p, base + p + n)
    order = ("tag", "text", "global", "semantic")
    arrays = [np.asarray(ids[k], dtype=np.int64) for k in order] + [np.asarray(dst[k], dtype=np.int64) for k in order] \
        + [mask.reshape(-1)]
    cuts = np.cumsum([0] + [len(a) for a in arrays])
    packed = torch.from_numpy(np.concatenate(arrays))
    packed = packed.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else packed.to(device)
    part = [packed[cuts[j]:cuts[j + 1]] for j in range(len(arrays))]
    tables = {"tag": llm.tts_tag_embedder, "text": llm.text_embedder, "global": llm.global_embedder,
              "semantic": llm.model.embeddings}
    out = _gather(tables, order, part[0:4], part[4:8], B * Tmax)
    return out.view(B, Tmax, -1), part[8].view(B, Tmax)


def create_inputs_and_labels_culens(batch: Dict[str, Any], tokenizer, model, eos_token_id: int, device) -> Dict[str, torch.Tensor]:
    """Packed-varlen variant (/root/reference/utils/multiple_jsonl.py:77-135): the samples of
    create_inputs_and_labels back to back as one sequence, `input_embs` [1, total, D], `labels` [1, total] and
    `cu_seqlens` [B+1].  Same host-side layout pass; the destination rows are simply consecutive."""
    texts, glob, sem = batch["text"], batch["global_tokens"], batch["semantic_tokens"]
    text_ids = [tokenizer.encode(t, add_special_tokens=False) for t in texts]
    B = len(texts)
    order = ("tag", "text", "global", "semantic")
    ids = {k: [] for k in order}
    dst = {k: [] for k in order}
    labels, cu = [], [0]
    for i in range(B):
        base, nt, ng = cu[-1], len(text_ids[i]), len(glob[i])
        sem_ids = list(sem[i]) + [eos_token_id]
        p_tag0 = 1 + nt
        p_tag1 = p_tag0 + 1 + ng
        p_sem = p_tag1 + 1
        ids["tag"] += [2, 0, 1]
        dst["tag"] += [base, base + p_tag0, base + p_tag1]
        ids["text"] += text_ids[i]
        dst["text"] += range(base + 1, base + 1 + nt)
        ids["global"] += list(glob[i])
        dst["global"] += range(base + p_tag0 + 1, base + p_tag0 + 1 + ng)
        ids["semantic"] += sem_ids
        dst["semantic"] += range(base + p_sem, base + p_sem + len(sem_ids))
        labels += [-100] * p_sem + sem_ids
        cu.append(base + p_sem + len(sem_ids))
    arrays = [np.asarray(ids[k], dtype=np.int64) for k in order] + [np.asarray(dst[k], dtype=np.int64) for k in order] \
        + [np.asarray(labels, dtype=np.int64), np.asarray(cu, dtype=np.int64)]
    cuts = np.cumsum([0] + [len(a) for a in arrays])
    packed = torch.from_numpy(np.concatenate(arrays))
    packed = packed.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else packed.to(device)
    part = [packed[cuts[j]:cuts[j + 1]] for j in range(len(arrays))]
    tables = {"tag": model.tts_tag_embedder, "text": model.text_embedder, "global": model.global_embedder,
              "semantic": model.model.embeddings}
    out = _gather(tables, order, part[0:4], part[4:8], cu[-1])
    return {"input_embs": out.unsqueeze(0), "labels": part[8].unsqueeze(0), "cu_seqlens": part[9]}


# ---------------------------------------------------------------------------------------------------------------
# Generic row assembler for the remaining layouts: a row is a list of segments (table, ids, predict); the embedding
# of segment ids comes from `table`, and the label of each position is its id when `predict` else -100.
# ---------------------------------------------------------------------------------------------------------------
_ORDER = ("tag", "text", "global", "semantic")


def _tables(model):
    return {"tag": model.tts_tag_embedder, "text": model.text_embedder, "global": model.global_embedder,
            "semantic": model.model.embeddings}


def _to_device(arrays, device):
    arrays = [np.asarray(a, dtype=np.int64).reshape(-1) for a in arrays]
    cuts = np.cumsum([0] + [a.size for a in arrays])
    packed = torch.from_numpy(np.concatenate(arrays))
    packed = packed.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else packed.to(device)
    return [packed[cuts[j]:cuts[j + 1]] for j in range(len(arrays))]


def _assemble_rows(rows, model, device, packed: bool):
    """rows -> (embeddings, labels, mask-or-cu_seqlens).  Padded ([R, Tmax, D], right padding, labels -100 / embeddings
    0.0 outside the row) or packed ([1, total, D] with cu_seqlens [R+1]).  One host pass, one host -> device copy, one
    lookup + one scatter per embedding table."""
    R = len(rows)
    lens = [sum(len(s[1]) for s in row) for row in rows]
    if packed:
        starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        total = int(starts[-1])
    else:
        Tmax = max(lens)
        starts = np.arange(R, dtype=np.int64) * Tmax
        total = R * Tmax
    ids = {k: [] for k in _ORDER}
    dst = {k: [] for k in _ORDER}
    labels = np.full(total, -100, dtype=np.int64)
    for r, row in enumerate(rows):
        p = int(starts[r])
        for table, seg, predict in row:
            n = len(seg)
            ids[table] += list(seg)
            dst[table] += range(p, p + n)
            if predict and n:
                labels[p:p + n] = seg
            p += n
    if packed:
        tail = starts
    else:
        tail = (np.arange(Tmax)[None, :] < np.asarray(lens)[:, None]).astype(np.int64)
    part = _to_device([ids[k] for k in _ORDER] + [dst[k] for k in _ORDER] + [labels, tail, _row_src(_ORDER, ids, dst, total)],
                      device)
    out = _gather(_tables(model), _ORDER, part[0:4], part[4:8], total, row_src=part[10])
    if packed:
        return out.unsqueeze(0), part[8].unsqueeze(0), part[9]
    return out.view(R, Tmax, -1), part[8].view(R, Tmax), part[9].view(R, Tmax)


def _properties_tokens(batch, i, tokenizer):
    from .properties import convert_properties_to_tokens
    s = convert_properties_to_tokens(batch["age"][i], batch["gender"][i], batch["emotion"][i], batch["pitch"][i], batch["speed"][i])
    return tokenizer.encode(s, add_special_tokens=False)


def _property_rows(batch, tokenizer, eos_token_id, plain: bool, predict_semantic: bool):
    """Rows of the controllable-TTS layouts (/root/reference/utils/multiple_jsonl.py:139-233, :313-400): per sample an
    optional plain row [tag2, text, tag0, global, tag1, semantic+eos] predicting the semantic ids, then the row with the
    property tokens (text table) in front, predicting the global ids and, optionally, the semantic ids."""
    rows = []
    for i, text in enumerate(batch["text"]):
        t = tokenizer.encode(text, add_special_tokens=False)
        g = list(batch["global_tokens"][i])
        s = list(batch["semantic_tokens"][i]) + [eos_token_id]
        if plain:
            rows.append([("tag", [2], False), ("text", t, False), ("tag", [0], False), ("global", g, False),
                         ("tag", [1], False), ("semantic", s, True)])
        rows.append([("text", _properties_tokens(batch, i, tokenizer), False), ("tag", [2], False), ("text", t, False),
                     ("tag", [0], False), ("global", g, True), ("tag", [1], False), ("semantic", s, predict_semantic)])
    return rows


def _padded(rows, model, device):
    e, l, m = _assemble_rows(rows, model, device, packed=False)
    return {"input_embs": e, "labels": l, "attention_mask": m}


def _packed(rows, model, device):
    e, l, cu = _assemble_rows(rows, model, device, packed=True)
    return {"input_embs": e, "labels": l, "cu_seqlens": cu}


def create_inputs_and_labels_with_properties(batch, tokenizer, model, eos_token_id, device):
    """/root/reference/utils/multiple_jsonl.py:139-233: 2 rows per sample (plain, then with properties)."""
    return _padded(_property_rows(batch, tokenizer, eos_token_id, plain=True, predict_semantic=True), model, device)


def create_inputs_and_labels_with_properties_culens(batch, tokenizer, model, eos_token_id, device):
    """/root/reference/utils/multiple_jsonl.py:236-311: the same rows packed back to back."""
    return _packed(_property_rows(batch, tokenizer, eos_token_id, plain=True, predict_semantic=True), model, device)


def create_inputs_and_labels_with_properties_global_tokens(batch, tokenizer, model, eos_token_id, device):
    """/root/reference/utils/multiple_jsonl.py:313-400: only the row with properties, only the global ids predicted."""
    return _padded(_property_rows(batch, tokenizer, eos_token_id, plain=False, predict_semantic=False), model, device)


def create_inputs_and_labels_with_properties_global_tokens_culens(batch, tokenizer, model, eos_token_id, device):
    """/root/reference/utils/multiple_jsonl.py:403-476."""
    return _packed(_property_rows(batch, tokenizer, eos_token_id, plain=False, predict_semantic=False), model, device)


def process_single_batch_culens(batch, rwkv7speech_model, eos_token_id: int = 8192, max_cu_seqlens: int = 8192):
    """/root/reference/data/utils/spark_dataset.py:111-162: the packed form of process_single_batch, cut at
    `max_cu_seqlens` tokens.  As in the reference (:153-156) the sample that crosses the limit is still part of
    `input_embs` / `labels` but gets no entry in `cu_seqlens`, and nothing after it is read."""
    model = rwkv7speech_model
    device = model.device
    ids_t, ids_g, ids_s = batch["input_ids"], batch["global_tokens_ids"], batch["semantic_tokens_ids"]
    B = ids_t.shape[0]
    lens = torch.stack([batch["attention_mask_input_ids"].sum(1), batch["global_tokens_attention_mask"].sum(1),
                        batch["semantic_tokens_attention_mask"].sum(1)]).tolist()          # the only host sync
    tl, gl, sl = ([int(v) for v in row] for row in lens)
    src = {k: [] for k in ("text", "global", "semantic")}
    dst = {k: [] for k in _ORDER}
    lab_dst, eos_dst, cu = [], [], [0]
    base = 0
    for i in range(B):
        n_i = tl[i] + gl[i] + sl[i] + 3
        p_tag0 = 1 + tl[i]
        p_tag1 = p_tag0 + 1 + gl[i]
        dst["tag"] += [base, base + p_tag0, base + p_tag1]
        for k, ids, n, p in (("text", ids_t, tl[i], 1), ("global", ids_g, gl[i], p_tag0 + 1), ("semantic", ids_s, sl[i], p_tag1 + 1)):
            L = ids.shape[1]
            src[k] += range(i * L + L - n, i * L + L)
            dst[k] += range(base + p, base + p + n)
        lab_dst += range(base + n_i - sl[i] - 1, base + n_i - 1)
        eos_dst.append(base + n_i - 1)
        base += n_i
        if cu[-1] + n_i > max_cu_seqlens:
            break
        cu.append(cu[-1] + n_i)
    n_rows = len(eos_dst)
    order = ("text", "global", "semantic")
    part = _to_device([src[k] for k in order] + [dst[k] for k in _ORDER] + [lab_dst, eos_dst, cu], device)
    s_text, s_glob, s_sem, d_tag, d_text, d_glob, d_sem, d_lab, d_eos, cu_t = part
    tok = {"text": ids_t.to(device).reshape(-1)[s_text], "global": ids_g.to(device).reshape(-1)[s_glob],
           "semantic": ids_s.to(device).reshape(-1)[s_sem]}
    tag_ids = torch.tensor([2, 0, 1], dtype=torch.long, device=device).repeat(n_rows)
    out = _gather(_tables(model), _ORDER, [tag_ids, tok["text"], tok["global"], tok["semantic"]],
                  [d_tag, d_text, d_glob, d_sem], base)
    labels = torch.full((base,), -100, dtype=torch.long, device=device)
    labels.index_copy_(0, d_lab, tok["semantic"].to(torch.long))
    labels.index_fill_(0, d_eos, eos_token_id)
    return {"input_embs": out.unsqueeze(0), "labels": labels.unsqueeze(0), "cu_seqlens": cu_t}


# ---------------------------------------------------------------------------------------------------------------
# Cosy layout (/root/reference/model/llm/cosy_llm.py:64-73, :86-88; same code in model/llm/llm.py:73-83, :101-103)
# ---------------------------------------------------------------------------------------------------------------
def pad_unpad_sequence(sos_eos_emb, text_token, text_token_len, task_id_emb, speech_token, speech_token_len,
                       padding_value: float = -1):
    """Same arguments and result as the reference's method of that name (without `self`): `text_token` [B, Lt, D] and
    `speech_token` [B, Ls, D] are right-padded EMBEDDINGS with their lengths; the result is the right-padded
    [sos, text_i, task_id, speech_i] batch [B, Tmax, D] (padding value IGNORE_ID = -1, as the reference pads its
    embeddings, :71) and an int32 attention mask [B, Tmax].  The reference unpads to 2B tensors, concatenates per sample
    and pads again; here it is one length read-back, one index array and one gather + scatter per source."""
    device = text_token.device
    B, Lt, D = text_token.shape
    Ls = speech_token.shape[1]
    tl = [int(v) for v in text_token_len.tolist()]
    sl = [int(v) for v in speech_token_len.tolist()]
    n = [2 + tl[i] + sl[i] for i in range(B)]
    Tmax = max(n)
    src_t, dst_t, src_s, dst_s, dst_sos, dst_task = [], [], [], [], [], []
    for i in range(B):
        base = i * Tmax
        dst_sos.append(base)
        src_t += range(i * Lt, i * Lt + tl[i])
        dst_t += range(base + 1, base + 1 + tl[i])
        dst_task.append(base + 1 + tl[i])
        src_s += range(i * Ls, i * Ls + sl[i])
        dst_s += range(base + 2 + tl[i], base + n[i])
    mask = (np.arange(Tmax)[None, :] < np.asarray(n)[:, None]).astype(np.int64)
    s_t, d_t, s_s, d_s, d_sos, d_task, m = _to_device([src_t, dst_t, src_s, dst_s, dst_sos, dst_task, mask], device)
    out = torch.full((B * Tmax, D), padding_value, dtype=text_token.dtype, device=device)
    out.index_copy_(0, d_sos, sos_eos_emb.reshape(1, D).to(out.dtype).expand(B, D))
    out.index_copy_(0, d_task, task_id_emb.reshape(1, D).to(out.dtype).expand(B, D))
    if len(src_t):
        out.index_copy_(0, d_t, text_token.reshape(B * Lt, D).index_select(0, s_t))
    if len(src_s):
        out.index_copy_(0, d_s, speech_token.reshape(B * Ls, D).to(out.dtype).index_select(0, s_s))
    return out.view(B, Tmax, D), m.view(B, Tmax).to(torch.int32)


def cosy_lm_target(text_token_len, speech_token, speech_token_len, speech_token_size: int, ignore_id: int = -1):
    """The reference's `lm_target` (/root/reference/model/llm/cosy_llm.py:86-88): per sample
    [ignore] * (2 + text_len) + speech ids + [speech_token_size], right-padded with `ignore_id`, int64 on the device of
    `speech_token` (the caller takes `[:, 1:]` as labels, :111).  The reference converts every row to a Python list
    (`.tolist()`, one sync per sample); here the lengths come back once and the ids never leave the device."""
    device = speech_token.device
    B, Ls = speech_token.shape
    tl = [int(v) for v in text_token_len.tolist()]
    sl = [int(v) for v in speech_token_len.tolist()]
    n = [2 + tl[i] + sl[i] + 1 for i in range(B)]
    Tmax = max(n)
    src, dst, eos = [], [], []
    for i in range(B):
        src += range(i * Ls, i * Ls + sl[i])
        dst += range(i * Tmax + 2 + tl[i], i * Tmax + 2 + tl[i] + sl[i])
        eos.append(i * Tmax + n[i] - 1)
    s, d, e = _to_device([src, dst, eos], device)
    out = torch.full((B * Tmax,), ignore_id, dtype=torch.long, device=device)
    if len(src):
        out.index_copy_(0, d, speech_token.reshape(-1).to(torch.long).index_select(0, s))
    out.index_fill_(0, e, speech_token_size)
    return out.view(B, Tmax)


# ---------------------------------------------------------------------------------------------------------------
# XY layout (/root/reference/train_scripts/train_xy_llm.py:91-216): 8 codebook channels in a staircase
# ---------------------------------------------------------------------------------------------------------------
def xy_staircase_batch(processed_features, num_channels: int, speech_vocab_size: int, text_pad_token_id: int,
                       ignore_id: int = -100) -> Dict[str, torch.Tensor]:
    """The layout half of the reference's `process_batch` (:121-216).  Each feature holds `text` ids [T1] and `speech`
    codes [num_channels, T2] (channel 0 already shifted into the text vocabulary).  Per sample, T1 + T2 + channels - 1
    steps: channel 0 = text then the text pad id, every other cell the audio pad id (speech_vocab_size - 1), and channel
    c's codes start c steps late (:146-153).  Labels are the inputs shifted by one step (:156), ignored on the text part
    except its last step (:159), ignored wherever they equal either pad id (:162-163), then each channel gets its pad id
    as the end-of-stream target right after its last code (:164-166).  Samples are padded to the longest with pad ids /
    ignore / mask 0.  CPU int64 tensors, as in the reference.

    The reference fills the staircase cell by cell (8 * (T2 + 7) tensor writes per sample, about a second per sample at
    T2 = 8192); here it is one strided copy per channel."""
    audio_pad = speech_vocab_size - 1
    n = [int(f["text"].shape[0]) + int(f["speech"].shape[1]) + num_channels - 1 for f in processed_features]
    B, Tmax = len(n), max(n)
    ids = np.full((B, Tmax, num_channels), audio_pad, dtype=np.int64)
    ids[:, :, 0] = text_pad_token_id
    labels = np.full((B, Tmax, num_channels), ignore_id, dtype=np.int64)
    mask = np.zeros((B, Tmax), dtype=np.int64)
    for i, f in enumerate(processed_features):
        text = np.asarray(f["text"].cpu() if torch.is_tensor(f["text"]) else f["text"], dtype=np.int64)
        speech = np.asarray(f["speech"].cpu() if torch.is_tensor(f["speech"]) else f["speech"], dtype=np.int64)
        T1, T2 = text.shape[0], speech.shape[1]
        ids[i, :T1, 0] = text
        for ch in range(num_channels):
            ids[i, T1 + ch:T1 + ch + T2, ch] = speech[ch]
        lab = labels[i, :n[i]]
        lab[:-1] = ids[i, 1:n[i]]
        lab[:T1 - 1] = ignore_id
        lab[(lab == audio_pad) | (lab == text_pad_token_id)] = ignore_id
        for ch in range(num_channels):
            lab[T1 + T2 - 1 + ch, ch] = text_pad_token_id if ch == 0 else audio_pad
        mask[i, :n[i]] = 1
    return {"input_ids": torch.from_numpy(ids), "labels": torch.from_numpy(labels), "attention_mask": torch.from_numpy(mask)}


def process_batch(features, text_tokenizer, xy_tokenizer, num_channels, text_shift_size, speech_vocab_size, device):
    """Same signature and result as the reference's `process_batch` (/root/reference/train_scripts/train_xy_llm.py:91-216):
    tokenises the text ("[S0]...[CTL0]", :100), runs the caller's audio codec (`xy_tokenizer.encode`, not part of this
    repo), shifts channel 0 by `text_shift_size` (:116) and lays the batch out with `xy_staircase_batch`.  Samples without
    text or audio are skipped; an empty batch gives {} (:120-121)."""
    processed = []
    for feature in features:
        text = f"[S0]{feature.get('json', {}).get('text', '')}[CTL0]"
        audio_np = feature.get("audio", {}).get("array")
        if not text or audio_np is None:
            continue
        text_tokens = text_tokenizer(text, return_tensors="pt").input_ids.squeeze(0)
        with torch.no_grad():
            codes = xy_tokenizer.encode([torch.from_numpy(audio_np).to(device)], device=device)["codes_list"][0].clone()
            codes[0, :] = codes[0, :] + text_shift_size
        processed.append({"text": text_tokens, "speech": codes})
    if not processed:
        return {}
    return xy_staircase_batch(processed, num_channels, speech_vocab_size, text_tokenizer.vocab_size - 1)


def collate_fn_for_rwkv7speech(batch, tokenizer, rwkv7speech_model, max_length=2048, pad_to_max_length=True, vocab_size=8193):
    """Same signature and result as the reference's collator of that name (/root/reference/data/utils/spark_dataset.py:41-52):
    a list of samples {text, global_tokens, semantic_tokens} -> LEFT-padded `input_embs` / `attention_mask` (create_inputs
    with the eos id `vocab_size - 1` appended to every semantic sequence) and `labels` = the semantic ids + eos placed one
    step early at the end of each row (`labels[i, -(n+1):-1]`, :50), -100 elsewhere.  `max_length` / `pad_to_max_length` are
    accepted and, as in the reference, unused.  One host index array and one scatter for the labels instead of a tensor
    construction and a slice assignment per sample."""
    device = rwkv7speech_model.device
    texts = [sample["text"] for sample in batch]
    global_tokens_ids = [sample["global_tokens"] for sample in batch]
    semantic_tokens_ids = [list(sample["semantic_tokens"]) + [vocab_size - 1] for sample in batch]
    input_ids_embs, attention_mask = create_inputs(texts, global_tokens_ids, semantic_tokens_ids, tokenizer, rwkv7speech_model)
    B, Tmax = input_ids_embs.shape[0], input_ids_embs.shape[1]
    dst, ids = [], []
    for i, sem in enumerate(semantic_tokens_ids):
        n = len(sem)
        dst += range(i * Tmax + Tmax - n - 1, i * Tmax + Tmax - 1)
        ids += sem
    d, v = _to_device([dst, ids], device)
    labels = torch.full((B * Tmax,), -100, dtype=torch.long, device=device)
    labels.index_copy_(0, d, v)
    return {"input_embs": input_ids_embs, "attention_mask": attention_mask, "labels": labels.view(B, Tmax)}


# /root/reference/data/utils/collator.py:8-130 is the same collator under another name
xy_data_collator = process_batch
