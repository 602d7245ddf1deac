"""the reference's OWN layout wrappers (model/llm/spark_llm.py, cosy_llm.py, xy_llm.py) run,
unmodified, on this repo's `rwkvfla` package and batch builders.  They are imported from /root/reference, so these tests
only run in the build container (the GPU box has no reference tree); the recurrent blocks are replaced by a causal
stand-in because the real ones need the CUDA library -- what is checked is the seam: construction through
PreTrainedModel / post_init, parameter registration, `self.model(inputs_embeds=..., attention_mask=...)`, the loss
modules the wrappers import from `rwkvfla`, and gradients reaching the wrappers' own embedding tables and heads."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model", "llm")), reason="reference tree not mounted")

sys.path.insert(0, os.path.join(ROOT, "tests"))


def _load(name, rel):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    third = os.path.join(REF, "third_party")
    if third not in sys.path:
        sys.path.append(third)                      # cosy_llm.py imports cosyvoice.* from the reference's third_party
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod                          # transformers looks the defining module up by name
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def stand_in_blocks(monkeypatch):
    from rwkvfla.models.rwkv7 import modeling_rwkv7 as M

    def fake_block(self, hidden_states, attention_mask=None, past_key_values=None, use_cache=False,
                   output_attentions=False, v_first=None, cu_seqlens=None, **kw):
        t = torch.arange(1, hidden_states.shape[1] + 1, dtype=hidden_states.dtype).view(1, -1, 1)
        return hidden_states.cumsum(1) / t + 0.1 * (self.layer_idx + 1), None, past_key_values, v_first

    monkeypatch.setattr(M.RWKV7Block, "forward", fake_block)
    return M


SMALL = dict(hidden_size=64, num_hidden_layers=2, decay_low_rank_dim=32, a_low_rank_dim=32, v_low_rank_dim=32,
             gate_low_rank_dim=32)


def test_reference_spark_wrapper_trains_on_the_seam(stand_in_blocks):
    import test_batch_builder as tb
    from rwkvtts_b200.batch import create_inputs_and_labels
    mod = _load("ref_spark_llm", "model/llm/spark_llm.py")
    torch.manual_seed(0)
    cfg = mod.RWKV7SpeechConfig(vocab_size=131, text_vocab_size=500, audio_global_vocab_size=64, fuse_cross_entropy=True, **SMALL)
    m = mod.RWKV7ForSpeech(cfg)
    assert isinstance(m, stand_in_blocks.RWKV7ForCausalLM) and isinstance(m.model, stand_in_blocks.RWKV7Model)
    names = dict(m.named_parameters())
    for n in ("model.embeddings.weight", "lm_head.weight", "text_embedder.weight", "global_embedder.weight",
              "tts_tag_embedder.weight", "model.layers.1.attn.w_lora.lora.2.bias", "model.layers.0.ffn.key.weight"):
        assert n in names, n
    out = create_inputs_and_labels(tb.make_batch(), tb.Tok(), m, 130, "cpu")          # this repo's builder feeds it
    m.train()
    m.dropout.p = 0.0                                                                   # make train == eval comparable
    r = m(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"], return_dict=True)
    assert r.logits is None                           # training + fuse_cross_entropy: rwkvfla's FusedLinearCrossEntropyLoss
    r.loss.backward()
    for n in ("lm_head.weight", "text_embedder.weight", "global_embedder.weight", "tts_tag_embedder.weight",
              "model.embeddings.weight", "model.norm.weight"):
        assert names[n].grad is not None and float(names[n].grad.abs().sum()) > 0, n
    m.eval()
    with torch.no_grad():
        e = m(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"], return_dict=True)
    assert e.logits.shape == (3, out["labels"].shape[1], 131)
    # the wrapper shifts the labels itself (spark_llm.py:156); both loss modules must agree with plain torch
    lab = torch.cat((out["labels"][:, 1:], torch.full_like(out["labels"][:, :1], -100)), 1)
    want = torch.nn.functional.cross_entropy(e.logits.reshape(-1, 131).float(), lab.reshape(-1), ignore_index=-100)
    assert abs(float(e.loss) - float(want)) < 1e-5 and abs(float(r.loss.detach()) - float(want)) < 1e-5


def test_reference_xy_wrapper_trains_on_the_seam(stand_in_blocks):
    import numpy as np
    import test_batch_builder as tb
    from rwkvtts_b200.batch import process_batch
    mod = _load("ref_xy_llm", "model/llm/xy_llm.py")
    torch.manual_seed(0)
    cfg = mod.RWKV7XYConfig(vocab_size=300, speech_vocab_size=40, num_channels=8, text_shift_size=256, **SMALL)
    m = mod.RWKV7XYLM(cfg)
    feats = [{"json": {"text": "hello"}, "audio": {"array": np.zeros(44, dtype=np.float32)}},
             {"json": {"text": "a longer line"}, "audio": {"array": np.zeros(20, dtype=np.float32)}}]
    b = process_batch(feats, tb._XYTextTok(), tb._XYCodec(), 8, 256, 40, "cpu")       # this repo's staircase builder
    m.train()
    r = m(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"], return_dict=True)
    assert len(r.logits) == 8 and r.logits[0].shape[-1] == 300 and r.logits[1].shape[-1] == 40
    want = sum(torch.nn.functional.cross_entropy(r.logits[i].reshape(-1, r.logits[i].shape[-1]), b["labels"][:, :, i].reshape(-1))
               for i in range(8))
    assert abs(float(r.loss.detach()) - float(want)) < 1e-5
    r.loss.backward()
    assert all(h.weight.grad is not None for h in m.heads) and all(e.weight.grad is not None for e in m.embs)


def test_reference_cosy_wrapper_runs_on_the_seam(stand_in_blocks):
    from rwkvtts_b200.batch import cosy_lm_target, pad_unpad_sequence
    mod = _load("ref_cosy_llm", "model/llm/cosy_llm.py")
    torch.manual_seed(0)
    cfg = mod.RWKV7CosyConfig(vocab_size=51, speech_token_size=50, **SMALL)
    m = mod.RWKV7CosyLM(cfg)
    # the attributes the reference's training script attaches before calling forward(batch=...)
    m.text_embedding = torch.nn.Embedding(100, 64)
    m.llm_embedding = torch.nn.Embedding(2, 64)
    m.speech_embedding = torch.nn.Embedding(51, 64)
    m.sos_eos, m.task_id, m.dropout = 0, 1, None
    m.criterion_ce = mod.LabelSmoothingLoss(size=51, padding_idx=-1, smoothing=0.0, normalize_length=True)
    g = torch.Generator().manual_seed(3)
    batch = {"text_token": torch.randint(0, 100, (3, 7), generator=g), "text_token_len": torch.tensor([7, 2, 5]),
             "speech_token": torch.randint(0, 50, (3, 11), generator=g), "speech_token_len": torch.tensor([4, 11, 9])}
    m.train()
    ref = m(batch=batch, return_dict=True)                 # the wrapper's own pad_unpad_sequence and lm_target lines
    # the same forward fed by this repo's builders
    x, mask = pad_unpad_sequence(m.llm_embedding.weight[0].reshape(1, 1, -1), m.text_embedding(batch["text_token"]),
                                 batch["text_token_len"], m.llm_embedding.weight[1].reshape(1, 1, -1),
                                 m.speech_embedding(batch["speech_token"]), batch["speech_token_len"])
    labels = cosy_lm_target(batch["text_token_len"], batch["speech_token"], batch["speech_token_len"], 50)[:, 1:].contiguous()
    mine = m(inputs_embeds=x, attention_mask=mask, labels=labels, return_dict=True)
    assert torch.equal(mine.logits, ref.logits) and torch.equal(mine.loss, ref.loss) and bool(torch.isfinite(ref.loss))


def _script_functions(rel, names, ns):
    """functions of a reference training script, without running its top-level imports (datasets, wandb, ...)"""
    import ast
    path = os.path.join(REF, rel)
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def test_reference_training_script_functions_run_on_the_engine(stand_in_blocks, tmp_path, capsys):
    """train_scripts/train_spark_rwkv7speech_jsonl.py: its own configure_optimizer (:161-199), update_learning_rate
    (:223-243), train_step (:271-285) and save_checkpoint (:201-221), on its own RWKV7ForSpeech, this repo's `deepspeed`
    shim (engine.py), and this repo's batch builders -- padded and packed (cu_seqlens) steps, with the script's default
    ZeRO-3 + offload config (:365-397), which the engine must accept."""
    import logging
    import types
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import deepspeed
    import test_batch_builder as tb
    from rwkvtts_b200.batch import create_inputs_and_labels, create_inputs_and_labels_culens
    mod = _load("ref_spark_llm", "model/llm/spark_llm.py")
    configure_optimizer, update_learning_rate, train_step, save_checkpoint = _script_functions(
        "train_scripts/train_spark_rwkv7speech_jsonl.py",
        ["configure_optimizer", "update_learning_rate", "train_step", "save_checkpoint"],
        {"torch": torch, "os": os, "deepspeed": deepspeed})
    torch.manual_seed(0)
    cfg = mod.RWKV7SpeechConfig(vocab_size=131, text_vocab_size=500, audio_global_vocab_size=64, fuse_cross_entropy=True, **SMALL)
    model = mod.RWKV7ForSpeech(cfg)
    model.dropout.p = 0.0
    model.train()
    args = types.SimpleNamespace(weight_decay=0.01, ds_optimizer_offload=True, learning_rate=1e-3, learning_rate_final=1e-5)
    optimizer = configure_optimizer(model, args)
    groups = {g["name"]: g for g in optimizer.param_groups}
    assert set(groups) == {"lr_1x", "lr_2x", "lr_decay"} and len(groups["lr_2x"]["params"]) == 2     # one decay bias per layer
    ds_config = {"distributed_backend": "nccl", "train_batch_size": 3, "bf16": {"enabled": False},
                 "zero_optimization": {"stage": 3, "stage3_max_live_parameters": 1e9, "stage3_max_reuse_distance": 1e9,
                                       "stage3_prefetch_bucket_size": 5e6, "memory_efficient_linear": True,
                                       "stage3_param_persistence_threshold": 1e4,
                                       "offload_param": {"device": "cpu", "pin_memory": True, "buffer_count": 4, "buffer_size": 1e8},
                                       "offload_optimizer": {"device": "cpu", "pin_memory": True, "buffer_count": 4},
                                       "allgather_partitions": True, "reduce_scatter": True, "reduce_bucket_size": 5e6,
                                       "overlap_comm": False, "contiguous_gradients": True},
                 "zero_force_ds_cpu_initialization": True, "gradient_checkpointing": False, "dump_state": False}
    engine, opt2, _, _ = deepspeed.initialize(model=model, config=ds_config, model_parameters=model.parameters(), optimizer=optimizer)
    assert opt2 is optimizer and engine.local_rank == 0
    before = {n: p.detach().clone() for n, p in engine.module.named_parameters()}
    batch = tb.make_batch()
    losses = []
    for step, builder in enumerate((create_inputs_and_labels, create_inputs_and_labels_culens, create_inputs_and_labels)):
        update_learning_rate(optimizer, step, 100, 2, args.learning_rate, args.learning_rate_final, args, True)
        assert abs(groups["lr_2x"]["lr"] - 2.0 * groups["lr_1x"]["lr"]) < 1e-12 and groups["lr_decay"]["weight_decay"] == 0.01
        out = train_step(engine, **builder(batch, tb.Tok(), engine, 130, engine.device))     # the script passes the engine as `model`
        assert bool(torch.isfinite(out["loss"]))
        engine.backward(out["loss"])
        engine.step()
        losses.append(float(out["loss"].detach()))
    assert losses[2] < losses[0]                                  # the same batch again after two updates
    changed = [n for n, p in engine.module.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert {"lm_head.weight", "text_embedder.weight", "tts_tag_embedder.weight", "model.embeddings.weight"} <= set(changed)
    save_checkpoint(engine, str(tmp_path / "ckpt"), 0, 3, logging.getLogger("t"))
    capsys.readouterr()
    saved = tmp_path / "ckpt" / "epoch_0_step_3"
    assert (saved / "latest").exists() and any(f.name.startswith("mp_rank_00") for f in saved.rglob("*.pt"))


def test_reference_lr_scheduler_path_on_the_engine(stand_in_blocks):
    """train_scripts/train_spark_rwkv7speech.py: its configure_optimizer (:178-197), get_lr_scheduler (:219-232, a LambdaLR
    handed to deepspeed.initialize(lr_scheduler=...), :566-572) and process_single_batch-style collated batches."""
    import types
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import deepspeed
    import test_batch_builder as tb
    from rwkvtts_b200.batch import process_single_batch
    mod = _load("ref_spark_llm", "model/llm/spark_llm.py")
    configure_optimizer, get_lr_scheduler, train_step = _script_functions(
        "train_scripts/train_spark_rwkv7speech.py", ["configure_optimizer", "get_lr_scheduler", "train_step"],
        {"torch": torch, "os": os, "deepspeed": deepspeed})
    torch.manual_seed(0)
    cfg = mod.RWKV7SpeechConfig(vocab_size=131, text_vocab_size=500, audio_global_vocab_size=64, fuse_cross_entropy=True, **SMALL)
    model = mod.RWKV7ForSpeech(cfg)
    model.dropout.p = 0.0
    model.train()
    args = types.SimpleNamespace(weight_decay=0.01, ds_optimizer_offload=False, learning_rate=1e-3, learning_rate_final=1e-5)
    optimizer = configure_optimizer(model, args)
    sched = get_lr_scheduler(optimizer, 10, 2, args.learning_rate, args.learning_rate_final)
    engine, opt2, _, sched2 = deepspeed.initialize(model=model, config={"train_batch_size": 4, "bf16": {"enabled": False},
                                                                        "zero_optimization": {"stage": 2}},
                                                   model_parameters=model.parameters(), optimizer=optimizer, lr_scheduler=sched)
    assert opt2 is optimizer and sched2 is sched
    lrs, losses = [], []
    for step in range(4):
        batch = process_single_batch(tb._padded_batch(), engine, eos_token_id=130)      # engine passed where a model is expected
        out = train_step(engine, **batch)
        engine.backward(out["loss"])
        engine.step()                                   # steps the scheduler as DeepSpeed does
        lrs.append(optimizer.param_groups[0]["lr"])
        losses.append(float(out["loss"].detach()))
    want = [args.learning_rate * f for f in (0.5, 1.0, 1.0 - (1 / 8) * (1 - 0.01), 1.0 - (2 / 8) * (1 - 0.01))]
    assert all(abs(a - b) < 1e-12 for a, b in zip(lrs, want)), (lrs, want)
    assert losses[-1] < losses[1]                       # the first step runs at lr = 0 (warm-up from zero)


# ---------------------------------------------------------------------------------------------------------------
# this repo's own layout classes (rwkvtts_b200/spark.py, layouts.py: what runs where the reference tree is absent)
# against the reference's classes: same state-dict keys, same loss / logits on the same batch
# ---------------------------------------------------------------------------------------------------------------
def _same(a, b, tol=1e-5):
    return float((a.float() - b.float()).abs().max()) <= tol * max(1.0, float(b.float().abs().max()))


def test_own_layout_classes_are_interchangeable_with_the_reference_classes(stand_in_blocks):
    import test_batch_builder as tb
    from rwkvtts_b200 import layouts, spark
    from rwkvtts_b200.batch import create_inputs_and_labels
    # Spark
    mod = _load("ref_spark_llm2", "model/llm/spark_llm.py")
    torch.manual_seed(0)
    kw = dict(vocab_size=131, text_vocab_size=500, audio_global_vocab_size=64, fuse_cross_entropy=False, **SMALL)
    ref = mod.RWKV7ForSpeech(mod.RWKV7SpeechConfig(**kw)).eval()
    own = spark.RWKV7ForSpeech(spark.RWKV7SpeechConfig(**kw)).eval()
    assert set(own.state_dict()) == set(ref.state_dict())
    own.load_state_dict(ref.state_dict(), strict=True)
    out = create_inputs_and_labels(tb.make_batch(), tb.Tok(), ref, 130, "cpu")
    with torch.no_grad():
        a = ref(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"], return_dict=True)
        b = own(inputs_embeds=out["input_embs"], attention_mask=out["attention_mask"], labels=out["labels"])
    assert _same(b.logits, a.logits) and _same(b.loss, a.loss)
    # Cosy
    mod = _load("ref_cosy_llm2", "model/llm/cosy_llm.py")
    kw = dict(vocab_size=100, speech_token_size=50, lsm_weight=0.1, **SMALL)
    ref = mod.RWKV7CosyLM(mod.RWKV7CosyConfig(**kw)).eval()
    own = layouts.RWKV7CosyLM(layouts.RWKV7CosyConfig(**kw)).eval()
    assert set(own.state_dict()) == set(ref.state_dict())
    own.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(3)
    batch = {"text_token": torch.randint(0, 100, (3, 7), generator=g), "text_token_len": torch.tensor([7, 2, 5]),
             "speech_token": torch.randint(0, 50, (3, 11), generator=g), "speech_token_len": torch.tensor([4, 11, 9])}
    with torch.no_grad():
        a, b = ref(batch=batch, return_dict=True), own(batch=batch)
    assert _same(b.logits, a.logits) and _same(b.loss, a.loss)
    # XY
    mod = _load("ref_xy_llm2", "model/llm/xy_llm.py")
    kw = dict(vocab_size=300, speech_vocab_size=40, num_channels=8, text_shift_size=256, **SMALL)
    ref = mod.RWKV7XYLM(mod.RWKV7XYConfig(**kw)).eval()
    own = layouts.RWKV7XYLM(layouts.RWKV7XYConfig(**kw)).eval()
    assert set(own.state_dict()) == set(ref.state_dict())
    own.load_state_dict(ref.state_dict(), strict=True)
    ids = torch.cat([torch.randint(0, 299, (2, 9, 1), generator=g), torch.randint(0, 39, (2, 9, 7), generator=g)], dim=2)
    lab = torch.cat([torch.randint(0, 299, (2, 9, 1), generator=g), torch.randint(0, 39, (2, 9, 7), generator=g)], dim=2)
    with torch.no_grad():
        a = ref(input_ids=ids, attention_mask=torch.ones(2, 9, dtype=torch.long), labels=lab, return_dict=True)
        b = own(input_ids=ids, attention_mask=torch.ones(2, 9, dtype=torch.long), labels=lab)
    assert _same(b.loss, a.loss) and all(_same(x, y) for x, y in zip(b.logits, a.logits))
    own.train()
    c = own(input_ids=ids, attention_mask=torch.ones(2, 9, dtype=torch.long), labels=lab)       # chunked heads, no logits
    assert c.logits == [] and _same(c.loss.detach(), a.loss, 1e-4)


def test_xy_sample_loop_equals_the_reference_sample(stand_in_blocks):
    """The 8-channel sampling loop (xy_llm.py:39-146) against the reference's `_sample` called directly (HF's generate()
    of this image's transformers can no longer reach it): same seed -> same tokens on every channel, including the
    reference's literal stopping rule (reference_termination=True)."""
    from transformers.generation import GenerationConfig, LogitsProcessorList, StoppingCriteriaList
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper
    from transformers.generation.stopping_criteria import MaxLengthCriteria
    from rwkvtts_b200 import layouts
    mod = _load("ref_xy_llm3", "model/llm/xy_llm.py")
    torch.manual_seed(0)
    kw = dict(vocab_size=300, speech_vocab_size=40, num_channels=8, text_shift_size=256, **SMALL)
    ref = mod.RWKV7XYLM(mod.RWKV7XYConfig(**kw)).eval()
    own = layouts.RWKV7XYLM(layouts.RWKV7XYConfig(**kw)).eval()
    own.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(11)
    ids = torch.cat([torch.randint(0, 255, (2, 5, 1), generator=g), torch.randint(0, 39, (2, 5, 7), generator=g)], dim=2)
    procs = LogitsProcessorList([TemperatureLogitsWarper(0.8), TopKLogitsWarper(10)])
    crit = StoppingCriteriaList([MaxLengthCriteria(max_length=12)])
    gc = GenerationConfig(eos_token_id=299, output_scores=False, return_dict_in_generate=False)
    # the reference's loop re-feeds the whole sequence through prepare_inputs_for_generation; give it the plain one
    ref.prepare_inputs_for_generation = lambda input_ids, **kw_: {"input_ids": input_ids}
    ref._update_model_kwargs_for_generation = lambda outputs, model_kwargs, **kw_: model_kwargs
    torch.manual_seed(123)
    want = ref._sample(ids, procs, crit, gc, False, None)
    torch.manual_seed(123)
    got = own.sample(ids, max_length=12, eos_token_id=299, temperature=0.8, top_k=10, reference_termination=True)
    assert got.shape == want.shape and torch.equal(got, want), (got[:, 5:], want[:, 5:])
    assert got.shape[1] == 6                       # the literal rule: every row stops after its first step
    assert bool(own.is_audio_token(got[:, 5, 0]).all())
    # the intended rule runs to max_length here (channel 0 is constrained to audio tokens, so no flush ever starts)
    torch.manual_seed(123)
    long = own.sample(ids, max_length=12, eos_token_id=299, temperature=0.8, top_k=10)
    assert long.shape[1] == 12 and torch.equal(long[:, :6], want)
