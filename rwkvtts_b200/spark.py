"""Spark-layout speech LM on the `rwkvfla` seam, for runs where the reference tree is not mounted (bench.py, the GPU
tests): the same module and parameter names as the reference's wrapper, so a checkpoint of either loads into the other
(state_dict keys: model.*, lm_head.weight, text_embedder.weight, global_embedder.weight, tts_tag_embedder.weight).

Reference: /root/reference/model/llm/spark_llm.py:13-52 (constructor: three extra embedding tables, dropout 0.02 on the
input embeddings while training, :121-122) and :105-172 (forward: backbone, label shift inside, fused linear + CE while
training -- all of which `RWKV7ForCausalLM.forward` of this repo's rwkvfla already does, so the subclass only adds what
the layout adds).  With the reference tree present the reference's own file runs unmodified on the same seam
(tests/test_reference_wrappers.py); this class is what the bench trains.
"""
from __future__ import annotations

import torch
from torch import nn

from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM

TAG_GLOBAL, TAG_SEMANTIC, TAG_START_TTS = 0, 1, 2        # rows of tts_tag_embedder (spark_llm.py:30)


class RWKV7SpeechConfig(RWKV7Config):
    def __init__(self, text_vocab_size: int = 65536, audio_global_vocab_size: int = 4096, **kwargs):
        super().__init__(**kwargs)
        self.text_vocab_size = text_vocab_size
        self.audio_global_vocab_size = audio_global_vocab_size


class RWKV7ForSpeech(RWKV7ForCausalLM):
    config_class = RWKV7SpeechConfig
    input_dropout = 0.02

    def __init__(self, config: RWKV7SpeechConfig):
        super().__init__(config)
        d = config.hidden_size
        self.text_embedder = nn.Embedding(config.text_vocab_size, d)
        self.global_embedder = nn.Embedding(config.audio_global_vocab_size, d)
        self.tts_tag_embedder = nn.Embedding(3, d)
        self.dropout = nn.Dropout(self.input_dropout)
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, **kwargs):
        if self.training and inputs_embeds is not None:
            inputs_embeds = self.dropout(inputs_embeds)
        # spark_llm.py:137: `fuse_linear_and_cross_entropy = self.config.fuse_cross_entropy and self.training` -- while
        # training the head and the loss are one op and no logits come back (:139-158)
        cfg, prev = self.config, self.config.fuse_linear_cross_entropy
        cfg.fuse_linear_cross_entropy = bool(prev or (cfg.fuse_cross_entropy and self.training))
        try:
            return super().forward(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds, **kwargs)
        finally:
            cfg.fuse_linear_cross_entropy = prev


def spark_0p4b_config(**over) -> RWKV7SpeechConfig:
    """BASELINE.json configs[1..3]: RWKV-7 0.4B (D 1024, 24 layers, 16 heads of 64; LoRA 64/64/32/128, SURVEY.md
    section 8 table), semantic vocabulary 8192 + EOS = 8193 (spark_llm.py:26), text vocabulary 65536, 4096 global tokens."""
    kw = dict(hidden_size=1024, num_hidden_layers=24, head_dim=64, vocab_size=8193, decay_low_rank_dim=64,
              a_low_rank_dim=64, v_low_rank_dim=32, gate_low_rank_dim=128, fuse_cross_entropy=True,
              text_vocab_size=65536, audio_global_vocab_size=4096)
    kw.update(over)
    return RWKV7SpeechConfig(**kw)


def synthetic_spark_batch(B: int, T: int, seed: int = 42, text_len: int = 128, global_len: int = 32,
                          text_vocab: int = 65536, global_vocab: int = 4096, semantic_vocab: int = 8192, pin: bool = True):
    """SURVEY.md section 8(d) "model level, Spark": per sample `text_len` text ids, `global_len` global ids and semantic
    ids filling the row to exactly T positions ([tag2, text, tag0, global, tag1, semantic]: T - text - global - 3
    semantic ids; the EOS lives in the pre-shifted labels only, spark_dataset.py:231-232), as the LEFT-padded id matrices + masks the collator of train_spark_rwkv7speech.py hands to
    process_single_batch (data/utils/spark_dataset.py:165-239).  Host tensors (pinned when a GPU is present)."""
    g = torch.Generator().manual_seed(seed)
    n_sem = T - text_len - global_len - 3
    assert n_sem > 0
    mk = lambda n, hi: torch.randint(0, hi, (B, n), generator=g, dtype=torch.long)
    batch = {"input_ids": mk(text_len, text_vocab), "global_tokens_ids": mk(global_len, global_vocab),
             "semantic_tokens_ids": mk(n_sem, semantic_vocab)}
    batch["attention_mask_input_ids"] = torch.ones_like(batch["input_ids"])
    batch["global_tokens_attention_mask"] = torch.ones_like(batch["global_tokens_ids"])
    batch["semantic_tokens_attention_mask"] = torch.ones_like(batch["semantic_tokens_ids"])
    if pin and torch.cuda.is_available():
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch
