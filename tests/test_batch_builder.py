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
