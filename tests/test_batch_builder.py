"""rwkvtts_b200.batch.create_inputs_and_labels against the reference's own function (utils/multiple_jsonl.py:4-75) when
/root/reference is mounted (build container), and against a committed golden made by that function otherwise."""
import importlib.util
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/utils/multiple_jsonl.py"
GOLD = os.path.join(ROOT, "tests", "golden", "batch_builder.pt")


class Tok:
    """stand-in tokenizer: deterministic ids from the characters (the builder only calls encode())"""
    def encode(self, text, add_special_tokens=False):
        return [(ord(c) * 7 + i) % 500 for i, c in enumerate(text)]


def make_model(D=16, seed=0):
    torch.manual_seed(seed)
    m = types.SimpleNamespace()
    m.text_embedder = torch.nn.Embedding(500, D)
    m.global_embedder = torch.nn.Embedding(64, D)
    m.tts_tag_embedder = torch.nn.Embedding(3, D)
    m.model = types.SimpleNamespace(embeddings=torch.nn.Embedding(130, D))
    return m


def make_batch():
    g = torch.Generator().manual_seed(1)
    texts = ["hello world", "a", "the quick brown fox jumps"]
    return {"text": texts,
            "global_tokens": [torch.randint(0, 64, (n,), generator=g).tolist() for n in (4, 4, 2)],
            "semantic_tokens": [torch.randint(0, 128, (n,), generator=g).tolist() for n in (9, 20, 3)]}


def _reference_fn():
    if not os.path.exists(REF):
        return None
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.dirname(REF)]
    sys.modules.setdefault("utils", pkg)
    spec = importlib.util.spec_from_file_location("utils.multiple_jsonl", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.create_inputs_and_labels


def test_matches_reference_builder_and_golden():
    from rwkvtts_b200.batch import create_inputs_and_labels
    model, batch = make_model(), make_batch()
    got = create_inputs_and_labels(batch, Tok(), model, 128, "cpu")
    ref_fn = _reference_fn()
    if ref_fn is not None:
        ref = ref_fn(batch, Tok(), model, 128, "cpu")
        if not os.path.exists(GOLD):
            torch.save({k: v.detach() for k, v in ref.items()}, GOLD)
    else:
        ref = torch.load(GOLD)
    assert torch.equal(got["input_embs"], ref["input_embs"])
    assert torch.equal(got["labels"], ref["labels"])
    assert torch.equal(got["attention_mask"], ref["attention_mask"].to(got["attention_mask"].dtype))
    gold = torch.load(GOLD)
    assert torch.equal(got["input_embs"].detach(), gold["input_embs"])


def test_embedding_tables_receive_gradients():
    from rwkvtts_b200.batch import create_inputs_and_labels
    model, batch = make_model(seed=3), make_batch()
    out = create_inputs_and_labels(batch, Tok(), model, 128, "cpu")
    w = torch.randn_like(out["input_embs"])
    (out["input_embs"] * w).sum().backward()
    # same gradient as assembling the rows one by one
    ref_grad = torch.zeros_like(model.model.embeddings.weight)
    for i, sem in enumerate(batch["semantic_tokens"]):
        p = 1 + len(Tok().encode(batch["text"][i])) + 1 + len(batch["global_tokens"][i]) + 1
        for j, t in enumerate(sem + [128]):
            ref_grad[t] += w[i, p + j]
    assert torch.allclose(model.model.embeddings.weight.grad, ref_grad, atol=1e-6)
    for emb in (model.text_embedder, model.global_embedder, model.tts_tag_embedder):
        assert emb.weight.grad is not None and float(emb.weight.grad.abs().sum()) > 0


# ---------------------------------------------------------------------------------------------------------------
# process_single_batch (data/utils/spark_dataset.py:165-239): the reference module imports `datasets` and its inference
# package at the top, so the function's own source is extracted with ast and executed on its own (it only needs torch)
# ---------------------------------------------------------------------------------------------------------------
REF2 = "/root/reference/data/utils/spark_dataset.py"
GOLD2 = os.path.join(ROOT, "tests", "golden", "process_single_batch.pt")


def _reference_process_single_batch():
    if not os.path.exists(REF2):
        return None
    import ast
    src = open(REF2).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "process_single_batch"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF2, "exec"), ns)
    return ns["process_single_batch"]


def _padded_batch():
    g = torch.Generator().manual_seed(5)
    B, Lt, Lg, Ls = 4, 12, 6, 20
    tl, gl, sl = [12, 3, 7, 1], [6, 6, 2, 4], [20, 5, 11, 1]
    def left(L, lens, hi):
        ids = torch.randint(0, hi, (B, L), generator=g)
        m = torch.zeros(B, L, dtype=torch.long)
        for i, n in enumerate(lens):
            m[i, L - n:] = 1
        return ids * m, m
    it, mt = left(Lt, tl, 500)
    ig, mg = left(Lg, gl, 64)
    is_, ms = left(Ls, sl, 128)
    return {"input_ids": it, "attention_mask_input_ids": mt, "global_tokens_ids": ig, "global_tokens_attention_mask": mg,
            "semantic_tokens_ids": is_, "semantic_tokens_attention_mask": ms}


def test_process_single_batch_matches_reference_and_golden():
    from rwkvtts_b200.batch import process_single_batch
    model = make_model(seed=7)
    model.device = torch.device("cpu")
    batch = _padded_batch()
    got = process_single_batch(batch, model, eos_token_id=129)
    ref_fn = _reference_process_single_batch()
    if ref_fn is not None:
        ref = ref_fn(batch, model, eos_token_id=129)
        if not os.path.exists(GOLD2):
            torch.save({k: v.detach() for k, v in ref.items()}, GOLD2)
    else:
        ref = torch.load(GOLD2)
    for k in ("input_embs", "attention_mask", "labels"):
        assert torch.equal(got[k], ref[k].to(got[k].dtype)), k
    gold = torch.load(GOLD2)
    assert torch.equal(got["labels"], gold["labels"]) and torch.equal(got["input_embs"].detach(), gold["input_embs"])


REF3 = "/root/reference/inference/rwkv7speech_inference.py"
GOLD3 = os.path.join(ROOT, "tests", "golden", "create_inputs.pt")


def _reference_create_inputs():
    if not os.path.exists(REF3):
        return None
    import ast
    from typing import List
    src = open(REF3).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "create_inputs"][0]
    ns = {"torch": torch, "List": List}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF3, "exec"), ns)
    return ns["create_inputs"]


def test_create_inputs_matches_reference_and_golden():
    """inference/rwkv7speech_inference.py:35-67, including the empty semantic prompt the decode loop starts from (:95)"""
    from rwkvtts_b200.batch import create_inputs
    model = make_model(seed=11)
    model.device = torch.device("cpu")
    b = make_batch()
    sem = [b["semantic_tokens"][0], [], b["semantic_tokens"][2]]
    got_e, got_m = create_inputs(b["text"], b["global_tokens"], sem, Tok(), model)
    ref_fn = _reference_create_inputs()
    if ref_fn is not None:
        ref_e, ref_m = ref_fn(b["text"], b["global_tokens"], sem, Tok(), model)
        if not os.path.exists(GOLD3):
            torch.save({"embs": ref_e.detach(), "mask": ref_m}, GOLD3)
    else:
        g_ = torch.load(GOLD3)
        ref_e, ref_m = g_["embs"], g_["mask"]
    assert torch.equal(got_e, ref_e.to(got_e.dtype)) and torch.equal(got_m, ref_m)
    gold = torch.load(GOLD3)
    assert torch.equal(got_e.detach(), gold["embs"].to(got_e.dtype))


def test_culens_builder_matches_reference_and_feeds_unpack_varlen():
    from rwkvtts_b200.batch import create_inputs_and_labels, create_inputs_and_labels_culens
    model, batch = make_model(seed=13), make_batch()
    got = create_inputs_and_labels_culens(batch, Tok(), model, 128, "cpu")
    if os.path.exists(REF):
        import importlib
        _reference_fn()                                   # loads utils.multiple_jsonl
        ref = sys.modules["utils.multiple_jsonl"] if "utils.multiple_jsonl" in sys.modules else None
        spec = importlib.util.spec_from_file_location("utils.multiple_jsonl", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        r = mod.create_inputs_and_labels_culens(batch, Tok(), model, 128, "cpu")
        for k in ("input_embs", "labels", "cu_seqlens"):
            assert torch.equal(got[k], r[k]), k
    # consistency with the padded builder (which is pinned to the reference and its golden) through the model's own
    # varlen unpacking: sample i of the padded batch == rows cu[i]:cu[i+1] of the packed one
    from rwkvfla.models.rwkv7.modeling_rwkv7 import unpack_varlen
    pad = create_inputs_and_labels(batch, Tok(), model, 128, "cpu")
    padded, _ = unpack_varlen(got["input_embs"], got["cu_seqlens"])
    assert torch.equal(padded, pad["input_embs"])
    cu = got["cu_seqlens"].tolist()
    for i, (a, b) in enumerate(zip(cu[:-1], cu[1:])):
        assert torch.equal(got["labels"][0, a:b], pad["labels"][i, :b - a])
        assert int(pad["attention_mask"][i].sum()) == b - a
