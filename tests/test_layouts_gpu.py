"""SURVEY.md section 8 row a11 on the GPU: goldens produced by the reference's OWN wrapper classes (Spark / Cosy / XY,
run on CPU in fp32 with the recurrence bound to the f64 oracle: tests/golden/make_wrapper_golden.py) replayed through this
repo's classes on the real kernels in bf16 -- logits and loss of the eval forward, and the training forward (fused
linear + cross-entropy heads, no logits) against the same loss."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu
TOL = 1.5e-2          # bf16 model against an fp32 / f64 reference: relative L2 of the logits


def _rel(a, b):
    return float((a.float().cpu() - b.float()).norm() / b.float().norm())


def _load(cls, cfg_cls, g, **extra):
    m = cls(cfg_cls(**g["config"], **extra))
    missing = m.load_state_dict({k: v.float() for k, v in g["state"].items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.cuda().to(torch.bfloat16)


def test_spark_wrapper_golden():
    import test_batch_builder as tb
    from rwkvtts_b200.batch import create_inputs_and_labels
    from rwkvtts_b200.spark import RWKV7ForSpeech, RWKV7SpeechConfig
    g = torch.load(os.path.join(GOLD, "wrapper_spark.pt"))
    m = _load(RWKV7ForSpeech, RWKV7SpeechConfig, g).eval()
    out = create_inputs_and_labels(g["batch"], tb.Tok(), m, g["eos"], "cuda")        # the gather kernel builds the batch
    assert torch.equal(out["labels"].cpu(), g["labels"]) and torch.equal(out["attention_mask"].cpu(), g["attention_mask"])
    assert _rel(out["input_embs"], g["input_embs"]) < 1e-6                            # bf16-exact tables: pure copies
    with torch.no_grad():
        r = m(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"])
    assert _rel(r.logits, g["logits"]) < TOL, _rel(r.logits, g["logits"])
    assert abs(float(r.loss) - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    # training forward: fused linear + CE (no logits), dropout off to compare
    m.train()
    m.dropout.p = 0.0
    m.config.fuse_linear_cross_entropy = True
    rt = m(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"])
    assert rt.logits is None and abs(float(rt.loss.detach()) - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    rt.loss.backward()
    assert all(torch.isfinite(p.grad.float()).all() for p in m.parameters() if p.grad is not None)


def test_cosy_wrapper_golden():
    from rwkvtts_b200.layouts import RWKV7CosyConfig, RWKV7CosyLM
    g = torch.load(os.path.join(GOLD, "wrapper_cosy.pt"))
    m = _load(RWKV7CosyLM, RWKV7CosyConfig, g).eval()
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    with torch.no_grad():
        r = m(batch=batch)
    # right-padded rows: compare the valid positions (padding positions see the pad value -1 as input, not a parity matter)
    lens = (2 + g["batch"]["text_token_len"] + g["batch"]["speech_token_len"]).tolist()
    for i, l in enumerate(lens):
        assert _rel(r.logits[i, :l], g["logits"][i, :l]) < TOL, (i, _rel(r.logits[i, :l], g["logits"][i, :l]))
    assert abs(float(r.loss) - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))


def test_xy_wrapper_golden():
    from rwkvtts_b200.layouts import RWKV7XYConfig, RWKV7XYLM
    g = torch.load(os.path.join(GOLD, "wrapper_xy.pt"))
    m = _load(RWKV7XYLM, RWKV7XYConfig, g).eval()
    ids, labels, mask = g["input_ids"].cuda(), g["labels"].cuda(), g["attention_mask"].cuda()
    with torch.no_grad():
        r = m(input_ids=ids, attention_mask=mask, labels=labels)
    for i in range(8):
        assert _rel(r.logits[i], g["logits"][i]) < TOL, (i, _rel(r.logits[i], g["logits"][i]))
    assert abs(float(r.loss) - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    m.train()                                   # 8 fused heads, no logits held
    rt = m(input_ids=ids, attention_mask=mask, labels=labels)
    assert rt.logits == [] and abs(float(rt.loss.detach()) - float(g["loss"])) < 1e-2 * abs(float(g["loss"]))
    rt.loss.backward()
    assert all(h.weight.grad is not None and torch.isfinite(h.weight.grad.float()).all() for h in m.heads)
    assert all(e.weight.grad is not None for e in m.embs)
