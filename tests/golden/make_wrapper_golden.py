"""Goldens of the reference's OWN layout wrappers (round-1 VERDICT 4.ii).

Run in the build container (needs /root/reference):   python tests/golden/make_wrapper_golden.py

The unmodified classes RWKV7ForSpeech (model/llm/spark_llm.py), RWKV7CosyLM (cosy_llm.py) and RWKV7XYLM (xy_llm.py) are
imported from the reference tree and run on CPU in fp32 on this repo's `rwkvfla` seam, with the recurrence bound to the
f64 oracle (oracle/wkv7_oracle.py) -- i.e. reference code + oracle, no CUDA kernel of this repo anywhere.  Saved per
layout: the config, the state dict, the batch as the reference's collators hand it over, and loss / logits.  The GPU
tests (tests/test_layouts_gpu.py) load the state dicts into this repo's classes and replay the batches on the real
kernels in bf16.
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")

SMALL = dict(hidden_size=128, num_hidden_layers=2, head_dim=64, decay_low_rank_dim=32, a_low_rank_dim=32, v_low_rank_dim=16,
             gate_low_rank_dim=32)


def _load(name, rel):
    third = os.path.join(REF, "third_party")
    if third not in sys.path:
        sys.path.append(third)
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def bind_oracle():
    """core._wkv -> f64 oracle (forward only; value-major state as the kernels)."""
    from oracle import wkv7_oracle as O
    from rwkvtts_b200 import core

    def wkv(r, w, k, v, a, b, state, need_state, inplace_state=False, plan=None):
        B, T, C = r.shape
        H = C // 64
        sh = lambda t: t.detach().to(torch.bfloat16).view(B, T, H, 64)      # the op's I/O is bf16
        y, sT = O.wkv7_forward(sh(w), sh(r), sh(k), sh(v), sh(a), sh(b), s0=state)
        return y.to(torch.bfloat16).to(r.dtype).view(B, T, C), (sT.float() if need_state else None)

    core._wkv = wkv
    core.FUSED = False


def randomize(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in m.named_parameters():
            if float(p.abs().sum()) == 0:                       # the reference zero-inits several projections
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            # keep every parameter exactly representable in bf16: the GPU replay casts the model to bf16
            p.copy_(p.to(torch.bfloat16).float())


def spark():
    import test_batch_builder as tb
    from rwkvtts_b200.batch import create_inputs_and_labels
    mod = _load("ref_spark_llm", "model/llm/spark_llm.py")
    torch.manual_seed(0)
    kw = dict(vocab_size=131, text_vocab_size=500, audio_global_vocab_size=64, fuse_cross_entropy=False, **SMALL)
    m = mod.RWKV7ForSpeech(mod.RWKV7SpeechConfig(**kw))
    randomize(m, 1)
    m.eval()
    out = create_inputs_and_labels(tb.make_batch(), tb.Tok(), m, 130, "cpu")
    with torch.no_grad():
        r = m(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"], return_dict=True)
    torch.save({"config": kw, "state": {k: v.to(torch.bfloat16) for k, v in m.state_dict().items()}, "batch": tb.make_batch(), "eos": 130,
                "input_embs": out["input_embs"], "attention_mask": out["attention_mask"], "labels": out["labels"],
                "loss": r.loss, "logits": r.logits}, os.path.join(OUT, "wrapper_spark.pt"))
    print("spark loss", float(r.loss), tuple(r.logits.shape))


def cosy():
    mod = _load("ref_cosy_llm", "model/llm/cosy_llm.py")
    torch.manual_seed(0)
    kw = dict(vocab_size=100, speech_token_size=50, lsm_weight=0.1, **SMALL)
    m = mod.RWKV7CosyLM(mod.RWKV7CosyConfig(**kw))
    randomize(m, 2)
    m.eval()
    g = torch.Generator().manual_seed(3)
    batch = {"text_token": torch.randint(0, 100, (3, 7), generator=g), "text_token_len": torch.tensor([7, 2, 5]),
             "speech_token": torch.randint(0, 50, (3, 11), generator=g), "speech_token_len": torch.tensor([4, 11, 9])}
    with torch.no_grad():
        r = m(batch=batch, return_dict=True)
    torch.save({"config": kw, "state": {k: v.to(torch.bfloat16) for k, v in m.state_dict().items()}, "batch": batch, "loss": r.loss, "logits": r.logits},
               os.path.join(OUT, "wrapper_cosy.pt"))
    print("cosy loss", float(r.loss), tuple(r.logits.shape))


def xy():
    mod = _load("ref_xy_llm", "model/llm/xy_llm.py")
    torch.manual_seed(0)
    kw = dict(vocab_size=300, speech_vocab_size=40, num_channels=8, text_shift_size=256, **SMALL)
    m = mod.RWKV7XYLM(mod.RWKV7XYConfig(**kw))
    randomize(m, 4)
    m.eval()
    g = torch.Generator().manual_seed(5)
    B, T = 2, 33
    ids = torch.cat([torch.randint(0, 299, (B, T, 1), generator=g), torch.randint(0, 39, (B, T, 7), generator=g)], dim=2)
    labels = torch.cat([torch.randint(0, 299, (B, T, 1), generator=g), torch.randint(0, 39, (B, T, 7), generator=g)], dim=2)
    labels[0, :4] = -100
    mask = torch.ones(B, T, dtype=torch.long)
    with torch.no_grad():
        r = m(input_ids=ids, attention_mask=mask, labels=labels, return_dict=True)
    torch.save({"config": kw, "state": {k: v.to(torch.bfloat16) for k, v in m.state_dict().items()}, "input_ids": ids, "labels": labels, "attention_mask": mask,
                "loss": r.loss, "logits": list(r.logits)}, os.path.join(OUT, "wrapper_xy.pt"))
    print("xy loss", float(r.loss), [tuple(l.shape) for l in r.logits][:2])


if __name__ == "__main__":
    assert os.path.isdir(os.path.join(REF, "model", "llm")), "needs the reference tree"
    bind_oracle()
    spark()
    cosy()
    xy()
