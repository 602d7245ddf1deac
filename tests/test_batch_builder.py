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


# ---------------------------------------------------------------------------------------------------------------
# controllable-TTS layouts (utils/multiple_jsonl.py:139-476) and the property tokens (utils/properties_util.py)
# ---------------------------------------------------------------------------------------------------------------
REF4 = "/root/reference/utils/properties_util.py"
GOLD4 = os.path.join(ROOT, "tests", "golden", "properties_tokens.json")
GOLD5 = os.path.join(ROOT, "tests", "golden", "batch_builder_properties.pt")


def _property_grid():
    ages = ["child", "teenager", "youth-adult", "middle-aged", "elderly", "Elderly"]
    genders = ["female", "male", "Male"]
    emotions = ["NEUTRAL", "happy", "NO-AGREEMENT", "CONTEMPT"]
    pitches = [0.0, 109.9, 110, 114, 115, 121, 125, 128, 130, 131, 142, 143, 147, 151, 153, 166, 170, 176, 187, 189.99, 190,
               191, 195, 208, 209, 211, 213, 215, 232, 238, 249.99, 250, 270, 289, 290, 400.5]
    speeds = [0.0, 3.5, 3.51, 3.99, 4.0, 4.01, 4.5, 4.51, 5.0, 5.01, 9.0]
    for a in ages:
        for g in genders:
            for e in emotions:
                for p in pitches:
                    for s in (speeds if e == "NEUTRAL" else speeds[3:6]):
                        yield a, g, e, p, s


def _load_reference_properties():
    if not os.path.exists(REF4):
        return None
    spec = importlib.util.spec_from_file_location("_ref_properties_util", REF4)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_property_tokens_match_reference_and_golden():
    import hashlib
    import json
    from rwkvtts_b200 import properties as P
    grid = list(_property_grid())
    got = [P.convert_properties_to_tokens(*x) for x in grid]
    ref = _load_reference_properties()
    if ref is not None:
        assert got == [ref.convert_properties_to_tokens(*x) for x in grid]
        for g in ("female", "male", "unknown", "other"):
            for a in ("child", "teenager", "youth-adult", "middle-aged", "elderly", "n/a"):
                for p in (0, 100, 113.9, 114, 129.9, 130, 150, 179.9, 180, 219.9, 220, 251, 300):
                    assert P.classify_pitch(p, g, a) == ref.classify_pitch(p, g, a), (p, g, a)
        assert P.convert_standard_properties_to_tokens("child", "FEMALE", "sad", "HIGH_PITCH", "Fast") == \
            ref.convert_standard_properties_to_tokens("child", "FEMALE", "sad", "HIGH_PITCH", "Fast")
        for name in ("SPEED_MAP", "PITCH_MAP", "AGE_MAP", "GENDER_MAP", "EMOTION_MAP"):
            assert getattr(P, name) == getattr(ref, name), name
        if not os.path.exists(GOLD4):
            json.dump({"n": len(grid), "sha256": hashlib.sha256("\n".join(got).encode()).hexdigest(),
                       "first": got[:5], "speed_4.0": P.classify_speed(4.0)}, open(GOLD4, "w"), indent=1)
    gold = json.load(open(GOLD4))
    assert gold["n"] == len(got) and gold["sha256"] == hashlib.sha256("\n".join(got).encode()).hexdigest()
    assert got[:5] == gold["first"] and P.classify_speed(4.0) == gold["speed_4.0"] == "very_fast"
    # error behaviour of the reference: unknown categories raise KeyError (the second GENDER_MAP has no "unknown")
    for bad in (("adult", "male", "SAD", 100.0, 4.2), ("child", "unknown", "SAD", 100.0, 4.2), ("child", "male", "BORED", 100.0, 4.2)):
        with pytest.raises(KeyError):
            P.convert_properties_to_tokens(*bad)


def _property_batch():
    b = make_batch()
    b.update({"age": ["child", "elderly", "youth-adult"], "gender": ["female", "male", "male"],
              "emotion": ["HAPPY", "sad", "NEUTRAL"], "pitch": [251.0, 120.5, 131.0], "speed": [4.0, 3.2, 4.7]})
    return b


PROP_FNS = ("create_inputs_and_labels_with_properties", "create_inputs_and_labels_with_properties_culens",
            "create_inputs_and_labels_with_properties_global_tokens", "create_inputs_and_labels_with_properties_global_tokens_culens")


def test_property_builders_match_reference_and_golden(capsys):
    import rwkvtts_b200.batch as mine
    model, batch = make_model(seed=17), _property_batch()
    got = {n: getattr(mine, n)(batch, Tok(), model, 128, "cpu") for n in PROP_FNS}
    if os.path.exists(REF):
        _reference_fn()
        spec = importlib.util.spec_from_file_location("utils.multiple_jsonl", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = {n: getattr(mod, n)(batch, Tok(), model, 128, "cpu") for n in PROP_FNS}
        capsys.readouterr()                          # the reference prints its first property string once
        if not os.path.exists(GOLD5):
            torch.save({n: {k: v.detach() for k, v in d.items()} for n, d in ref.items()}, GOLD5)
    else:
        ref = torch.load(GOLD5)
    gold = torch.load(GOLD5)
    for n in PROP_FNS:
        assert set(got[n]) == set(ref[n]) == set(gold[n])
        for k in got[n]:
            assert torch.equal(got[n][k], ref[n][k].to(got[n][k].dtype)), (n, k)
            assert torch.equal(got[n][k].detach(), gold[n][k].to(got[n][k].dtype)), (n, k)
    # 2 rows per sample (plain, with properties); the global-token variant has 1 and predicts no semantic id
    assert got[PROP_FNS[0]]["input_embs"].shape[0] == 6 and got[PROP_FNS[2]]["input_embs"].shape[0] == 3
    assert got[PROP_FNS[1]]["cu_seqlens"].numel() == 7 and got[PROP_FNS[3]]["cu_seqlens"].numel() == 4
    lab = got[PROP_FNS[2]]["labels"]
    assert int((lab >= 0).sum()) == sum(len(g) for g in batch["global_tokens"])


def _reference_process_single_batch_culens():
    if not os.path.exists(REF2):
        return None
    import ast
    src = open(REF2).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "process_single_batch_culens"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF2, "exec"), ns)
    return ns["process_single_batch_culens"]


@pytest.mark.parametrize("limit", [8192, 60, 30, 5])
def test_process_single_batch_culens_matches_reference_and_padded_collator(limit):
    from rwkvtts_b200.batch import process_single_batch, process_single_batch_culens
    model = make_model(seed=19)
    model.device = torch.device("cpu")
    batch = _padded_batch()
    got = process_single_batch_culens(batch, model, eos_token_id=129, max_cu_seqlens=limit)
    ref_fn = _reference_process_single_batch_culens()
    if ref_fn is not None:
        ref = ref_fn(batch, model, eos_token_id=129, max_cu_seqlens=limit)
        for k in ("input_embs", "labels", "cu_seqlens"):
            assert torch.equal(got[k], ref[k]), k
    # against the padded collator, which is pinned to the reference and to its golden: row i of the left-padded batch
    # (its last n_i positions) is segment i of the packed one, for every sample the limit let through
    pad = process_single_batch(batch, model, eos_token_id=129)
    n = pad["attention_mask"].sum(1).tolist()
    cu = got["cu_seqlens"].tolist()
    assert cu[0] == 0 and all(b - a == n[i] for i, (a, b) in enumerate(zip(cu[:-1], cu[1:])))
    assert len(cu) - 1 == sum(1 for i in range(len(n)) if sum(n[:i + 1]) <= limit and all(sum(n[:j + 1]) <= limit for j in range(i)))
    rows = min(len(cu), len(n))                    # the sample that crossed the limit is still in the buffers
    assert got["input_embs"].shape[1] == sum(n[:rows]) == got["labels"].shape[1]
    off = 0
    for i in range(rows):
        assert torch.equal(got["input_embs"][0, off:off + n[i]], pad["input_embs"][i, -n[i]:])
        assert torch.equal(got["labels"][0, off:off + n[i]], pad["labels"][i, -n[i]:])
        off += n[i]


# ---------------------------------------------------------------------------------------------------------------
# Cosy layout: RWKV7CosyLM.pad_unpad_sequence and the lm_target lines of its forward (model/llm/cosy_llm.py:64-88)
# ---------------------------------------------------------------------------------------------------------------
REF5 = "/root/reference/model/llm/cosy_llm.py"
GOLD6 = os.path.join(ROOT, "tests", "golden", "cosy_pad_unpad.pt")


def _reference_pad_unpad_sequence():
    if not os.path.exists(REF5):
        return None
    import ast
    from torch.nn.utils.rnn import pad_sequence, unpad_sequence
    cls = [n for n in ast.parse(open(REF5).read()).body if isinstance(n, ast.ClassDef) and n.name == "RWKV7CosyLM"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "pad_unpad_sequence"][0]
    ns = {"torch": torch, "pad_sequence": pad_sequence, "unpad_sequence": unpad_sequence, "IGNORE_ID": -1}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF5, "exec"), ns)
    return lambda *a: ns["pad_unpad_sequence"](None, *a)


def test_cosy_pad_unpad_sequence_and_target_match_reference_and_golden():
    from torch.nn.utils.rnn import pad_sequence
    from rwkvtts_b200.batch import cosy_lm_target, pad_unpad_sequence
    g = torch.Generator().manual_seed(23)
    B, Lt, Ls, D, V = 4, 9, 14, 8, 6561
    tl, sl = torch.tensor([9, 1, 4, 7]), torch.tensor([3, 14, 8, 1])
    text_ids = torch.randint(0, 100, (B, Lt), generator=g)
    speech_ids = torch.randint(0, V, (B, Ls), generator=g)
    text_emb = torch.randn(B, Lt, D, generator=g, requires_grad=True)
    speech_emb = torch.randn(B, Ls, D, generator=g, requires_grad=True)
    sos, task = torch.randn(1, 1, D, generator=g, requires_grad=True), torch.randn(1, 1, D, generator=g)
    got_x, got_m = pad_unpad_sequence(sos, text_emb, tl, task, speech_emb, sl)
    got_t = cosy_lm_target(tl, speech_ids, sl, V)
    ref_fn = _reference_pad_unpad_sequence()
    if ref_fn is not None:
        ref_x, ref_m = ref_fn(sos, text_emb, tl, task, speech_emb, sl)
        # cosy_llm.py:86-88
        ref_t = pad_sequence([torch.tensor([-1] * (2 + tl[i]) + speech_ids[i, :sl[i]].tolist() + [V]) for i in range(B)],
                             batch_first=True, padding_value=-1)
        if not os.path.exists(GOLD6):
            torch.save({"x": ref_x.detach(), "mask": ref_m, "target": ref_t}, GOLD6)
    else:
        g_ = torch.load(GOLD6)
        ref_x, ref_m, ref_t = g_["x"], g_["mask"], g_["target"]
    gold = torch.load(GOLD6)
    assert torch.equal(got_x, ref_x) and torch.equal(got_x.detach(), gold["x"])
    assert got_m.dtype == ref_m.dtype == torch.int32 and torch.equal(got_m, ref_m)
    assert got_t.dtype == ref_t.dtype and torch.equal(got_t, ref_t) and torch.equal(got_t, gold["target"])
    # gradients flow to the three sources; padding positions get none
    w = torch.randn(got_x.shape, generator=g)
    (got_x * w).sum().backward()
    assert torch.equal(sos.grad.reshape(-1), w[:, 0].sum(0))
    for i in range(B):
        assert torch.equal(text_emb.grad[i, :tl[i]], w[i, 1:1 + tl[i]]) and float(text_emb.grad[i, tl[i]:].abs().sum()) == 0
        assert torch.equal(speech_emb.grad[i, :sl[i]], w[i, 2 + tl[i]:2 + tl[i] + sl[i]])


# ---------------------------------------------------------------------------------------------------------------
# XY layout: process_batch of train_scripts/train_xy_llm.py:91-216 (staircase over 8 codebook channels)
# ---------------------------------------------------------------------------------------------------------------
REF6 = "/root/reference/train_scripts/train_xy_llm.py"
GOLD7 = os.path.join(ROOT, "tests", "golden", "xy_process_batch.pt")


class _XYTextTok:
    vocab_size = 300                                   # pad id 299

    def __call__(self, text, return_tensors="pt"):
        ids = torch.tensor([[(ord(c) * 11 + i) % 299 for i, c in enumerate(text)]])
        return types.SimpleNamespace(input_ids=ids)


class _XYCodec:
    """stand-in audio codec: 8 x T2 codes derived from the samples; some codes equal the pad ids on purpose
    (speech_vocab_size - 1 in any channel, text pad - shift in channel 0) to hit the reference's label masking"""
    def encode(self, wavs, device="cpu"):
        a = wavs[0]
        T2 = a.numel() // 4
        g = torch.Generator().manual_seed(int(a.numel()))
        codes = torch.randint(0, 40, (8, T2), generator=g)
        codes[3, T2 // 2] = 39
        codes[0, min(2, T2 - 1)] = 299 - 256
        return {"codes_list": [codes]}


def _reference_process_batch():
    if not os.path.exists(REF6):
        return None
    import ast
    import logging
    fn = [n for n in ast.parse(open(REF6).read()).body if isinstance(n, ast.FunctionDef) and n.name == "process_batch"][0]
    ns = {"torch": torch, "logger": logging.getLogger("xy")}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF6, "exec"), ns)
    return ns["process_batch"]


def test_xy_process_batch_matches_reference_and_golden():
    import numpy as np
    from rwkvtts_b200.batch import process_batch
    feats = [{"json": {"text": "hello"}, "audio": {"array": np.zeros(44, dtype=np.float32)}},
             {"json": {"text": "a much longer line of text"}, "audio": {"array": np.zeros(8, dtype=np.float32)}},
             {"json": {"text": "no audio"}, "audio": {}},
             {"json": {}, "audio": {"array": np.zeros(100, dtype=np.float32)}}]
    args = (_XYTextTok(), _XYCodec(), 8, 256, 40, "cpu")
    got = process_batch(feats, *args)
    ref_fn = _reference_process_batch()
    if ref_fn is not None:
        ref = ref_fn(feats, *args)
        if not os.path.exists(GOLD7):
            torch.save(ref, GOLD7)
        assert ref_fn([feats[2]], *args) == {}
    else:
        ref = torch.load(GOLD7)
    gold = torch.load(GOLD7)
    assert process_batch([feats[2]], *args) == {}
    assert got["input_ids"].shape[0] == 3 and got["input_ids"].shape[2] == 8
    for k in ("input_ids", "labels", "attention_mask"):
        assert got[k].dtype == ref[k].dtype and torch.equal(got[k], ref[k]), k
        assert torch.equal(got[k], gold[k]), k
    # the masking quirk is exercised: a real code equal to a pad id is not a target
    assert int((got["labels"] == 39).sum()) == 7 * 3 and int((got["labels"] == 299).sum()) == 3


# ---------------------------------------------------------------------------------------------------------------
# randomised ragged batches (including empty text / global / semantic lists and a single sample) against the
# reference's own functions; build container only
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_random_ragged_batches_match_reference(capsys):
    import random
    import rwkvtts_b200.batch as mine
    _reference_fn()
    spec = importlib.util.spec_from_file_location("utils.multiple_jsonl", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref_psb, ref_psb_cu = _reference_process_single_batch(), _reference_process_single_batch_culens()
    ref_ci = _reference_create_inputs()
    model = make_model(seed=29)
    model.device = torch.device("cpu")
    rng = random.Random(1234)
    ages, genders, emos = ["child", "teenager", "youth-adult", "middle-aged", "elderly"], ["male", "female"], ["HAPPY", "sad", "NEUTRAL", "WHISPER"]
    for trial in range(25):
        B = rng.choice([1, 1, 2, 3, 5])
        lo = 0 if trial % 2 else 1                      # every other trial allows empty pieces
        batch = {"text": ["".join(rng.choice("abc xyz") for _ in range(rng.randint(lo, 9))) for _ in range(B)],
                 "global_tokens": [[rng.randrange(64) for _ in range(rng.randint(lo, 5))] for _ in range(B)],
                 "semantic_tokens": [[rng.randrange(128) for _ in range(rng.randint(lo, 14))] for _ in range(B)],
                 "age": [rng.choice(ages) for _ in range(B)], "gender": [rng.choice(genders) for _ in range(B)],
                 "emotion": [rng.choice(emos) for _ in range(B)], "pitch": [rng.uniform(60, 330) for _ in range(B)],
                 "speed": [rng.choice([3.0, 3.5, 3.9, 4.0, 4.2, 4.5, 4.8, 5.0, 6.0]) for _ in range(B)]}
        for name in ("create_inputs_and_labels", "create_inputs_and_labels_culens") + PROP_FNS:
            got = getattr(mine, name)(batch, Tok(), model, 128, "cpu")
            want = getattr(ref, name)(batch, Tok(), model, 128, "cpu")
            assert set(got) == set(want), name
            for k in got:
                assert got[k].shape == want[k].shape and torch.equal(got[k], want[k].to(got[k].dtype)), (trial, name, k)
        ge, gm = mine.create_inputs(batch["text"], batch["global_tokens"], batch["semantic_tokens"], Tok(), model)
        we, wm = ref_ci(batch["text"], batch["global_tokens"], batch["semantic_tokens"], Tok(), model)
        assert torch.equal(ge, we.to(ge.dtype)) and torch.equal(gm, wm), trial
        # left-padded id matrices (lengths >= 1: with a zero length the reference's `[i, -0:]` slice takes the whole row)
        Lt, Lg, Ls = rng.randint(3, 9), rng.randint(2, 5), rng.randint(4, 14)
        def left(L, hi):
            lens = [rng.randint(1, L) for _ in range(B)]
            ids = torch.tensor([[rng.randrange(hi) for _ in range(L)] for _ in range(B)])
            m = torch.zeros(B, L, dtype=torch.long)
            for i, n in enumerate(lens):
                m[i, L - n:] = 1
            return ids * m, m
        it, mt = left(Lt, 500); ig, mg = left(Lg, 64); is_, ms = left(Ls, 128)
        pb = {"input_ids": it, "attention_mask_input_ids": mt, "global_tokens_ids": ig, "global_tokens_attention_mask": mg,
              "semantic_tokens_ids": is_, "semantic_tokens_attention_mask": ms}
        got, want = mine.process_single_batch(pb, model, eos_token_id=129), ref_psb(pb, model, eos_token_id=129)
        for k in got:
            assert torch.equal(got[k], want[k].to(got[k].dtype)), (trial, "process_single_batch", k)
        limit = rng.choice([8192, 40, 25, 12])
        got = mine.process_single_batch_culens(pb, model, eos_token_id=129, max_cu_seqlens=limit)
        want = ref_psb_cu(pb, model, eos_token_id=129, max_cu_seqlens=limit)
        for k in got:
            assert torch.equal(got[k], want[k]), (trial, "process_single_batch_culens", limit, k)
    capsys.readouterr()


@pytest.mark.skipif(not os.path.exists(REF5), reason="reference tree not mounted")
def test_random_cosy_and_xy_batches_match_reference():
    import random
    import numpy as np
    from torch.nn.utils.rnn import pad_sequence
    from rwkvtts_b200.batch import cosy_lm_target, pad_unpad_sequence, process_batch
    ref_pad, ref_xy = _reference_pad_unpad_sequence(), _reference_process_batch()
    rng = random.Random(99)
    g = torch.Generator().manual_seed(99)
    for trial in range(15):
        B, Lt, Ls, D, V = rng.choice([1, 2, 4]), rng.randint(1, 8), rng.randint(1, 12), 8, 50
        tl = torch.tensor([rng.randint(1, Lt) for _ in range(B)])
        sl = torch.tensor([rng.randint(0 if trial % 3 == 0 else 1, Ls) for _ in range(B)])
        text_emb, speech_emb = torch.randn(B, Lt, D, generator=g), torch.randn(B, Ls, D, generator=g)
        speech_ids = torch.randint(0, V, (B, Ls), generator=g)
        sos, task = torch.randn(1, 1, D, generator=g), torch.randn(1, 1, D, generator=g)
        gx, gm = pad_unpad_sequence(sos, text_emb, tl, task, speech_emb, sl)
        wx, wm = ref_pad(sos, text_emb, tl, task, speech_emb, sl)
        assert torch.equal(gx, wx) and torch.equal(gm, wm) and gm.dtype == wm.dtype, trial
        want_t = pad_sequence([torch.tensor([-1] * (2 + int(tl[i])) + speech_ids[i, :int(sl[i])].tolist() + [V]) for i in range(B)],
                              batch_first=True, padding_value=-1)
        assert torch.equal(cosy_lm_target(tl, speech_ids, sl, V), want_t), trial
    args = (_XYTextTok(), _XYCodec(), 8, 256, 40, "cpu")
    for trial in range(6):
        feats = [{"json": {"text": "".join(rng.choice("ab c") for _ in range(rng.randint(0, 12)))},
                  "audio": {"array": np.zeros(4 * rng.randint(1, 9), dtype=np.float32)}} for _ in range(rng.choice([1, 2, 3]))]
        got, want = process_batch(feats, *args), ref_xy(feats, *args)
        for k in ("input_ids", "labels", "attention_mask"):
            assert torch.equal(got[k], want[k]), (trial, k)


@pytest.mark.skipif(not os.path.exists(REF2), reason="reference tree not mounted")
def test_collate_fn_for_rwkv7speech_matches_reference():
    """data/utils/spark_dataset.py:41-52 (it calls the reference's own create_inputs)"""
    import ast
    import random
    from rwkvtts_b200.batch import collate_fn_for_rwkv7speech
    fn = [n for n in ast.parse(open(REF2).read()).body if isinstance(n, ast.FunctionDef) and n.name == "collate_fn_for_rwkv7speech"][0]
    ns = {"torch": torch, "create_inputs": _reference_create_inputs()}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF2, "exec"), ns)
    model = make_model(seed=31)
    model.device = torch.device("cpu")
    rng = random.Random(3)
    for trial in range(10):
        batch = [{"text": "".join(rng.choice("abc xyz") for _ in range(rng.randint(1, 9))),
                  "global_tokens": [rng.randrange(64) for _ in range(rng.randint(1, 5))],
                  "semantic_tokens": [rng.randrange(128) for _ in range(rng.randint(1, 14))]} for _ in range(rng.choice([1, 2, 4]))]
        got = collate_fn_for_rwkv7speech(batch, Tok(), model, vocab_size=130)
        want = ns["collate_fn_for_rwkv7speech"](batch, Tok(), model, vocab_size=130)
        assert set(got) == set(want)
        for k in got:
            assert torch.equal(got[k], want[k].to(got[k].dtype)), (trial, k)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/utils/collator.py"), reason="reference tree not mounted")
def test_xy_data_collator_alias_matches_reference():
    import ast
    import logging
    import numpy as np
    from rwkvtts_b200.batch import xy_data_collator
    path = "/root/reference/data/utils/collator.py"
    fn = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "xy_data_collator"][0]
    ns = {"torch": torch, "logger": logging.getLogger("xy")}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    feats = [{"json": {"text": "one"}, "audio": {"array": np.zeros(24, dtype=np.float32)}},
             {"json": {"text": "another sample"}, "audio": {"array": np.zeros(12, dtype=np.float32)}}]
    args = (_XYTextTok(), _XYCodec(), 8, 256, 40, "cpu")
    got, want = xy_data_collator(feats, *args), ns["xy_data_collator"](feats, *args)
    for k in ("input_ids", "labels", "attention_mask"):
        assert torch.equal(got[k], want[k]), k
