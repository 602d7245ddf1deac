"""Cosy and XY layout wrappers on the `rwkvfla` seam, for runs where the reference tree is not mounted (the GPU tests
replay goldens of the reference's own classes through these; bench.py / the Spark layout: rwkvtts_b200/spark.py).
Module and parameter names are the reference's, so state dicts are interchangeable.

  RWKV7CosyLM   /root/reference/model/llm/cosy_llm.py:24-160: text / speech / llm embeddings, head of
                speech_token_size + 1 classes WITH bias, `forward(batch=...)` builds [sos, text, task, speech] rows
                (pad_unpad_sequence :64-73), target [IGNORE]*(2+text_len) + speech + [eos] shifted by one (:86-88,:113),
                LabelSmoothingLoss normalised by the number of valid tokens (:46-51).
  RWKV7XYLM     /root/reference/model/llm/xy_llm.py:147-262: input = SUM of the 8 channel embeddings (:208-214), 8 heads
                with bias (vocab_size for channel 0, speech_vocab_size for 1..7), loss = sum of the 8 mean cross entropies
                against labels[:, :, i] (labels are pre-shifted by the collator, train_xy_llm.py:158).
What differs from the reference is only HOW: the batch rows are assembled by rwkvtts_b200.batch (index arrays, one
gather), the Cosy loss is the closed-form label-smoothing loss, and while training the XY heads never materialise their
logits (fused linear + cross-entropy per head, csrc/linear_ce.cu; channel 0 alone is 2.2 GB per GPU at config c5).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn
from transformers.modeling_outputs import CausalLMOutputWithPast

from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM

from .batch import cosy_lm_target, pad_unpad_sequence
from .losses import LabelSmoothingLoss, xy_channel_losses

IGNORE_ID = -1          # third_party/cosyvoice/utils/common.py


class RWKV7CosyConfig(RWKV7Config):
    def __init__(self, speech_token_size: int = 6561, length_normalized_loss: bool = True, lsm_weight: float = 0.0,
                 drop_ratio: float = 0.0, **kwargs):
        super().__init__(**kwargs)
        self.llm_input_size = kwargs.get("llm_input_size", self.hidden_size)
        self.llm_output_size = kwargs.get("llm_output_size", self.hidden_size)
        self.speech_token_size = speech_token_size
        self.length_normalized_loss = length_normalized_loss
        self.lsm_weight = lsm_weight
        self.mix_ratio = kwargs.get("mix_ratio", [5, 15])
        self.drop_ratio = drop_ratio


class RWKV7CosyLM(RWKV7ForCausalLM):
    config_class = RWKV7CosyConfig

    def __init__(self, config: RWKV7CosyConfig):
        super().__init__(config)
        self.sos_eos, self.task_id, self.fill_token = 0, 1, 2
        self.llm_embedding = nn.Embedding(2, config.llm_input_size)
        self.text_embedding = nn.Embedding(config.vocab_size, config.llm_input_size)
        self.speech_embedding = nn.Embedding(config.speech_token_size + 1, config.llm_input_size)
        self.lm_head = nn.Linear(config.hidden_size, config.speech_token_size + 1)
        self.criterion_ce = LabelSmoothingLoss(size=config.speech_token_size + 1, padding_idx=IGNORE_ID,
                                               smoothing=config.lsm_weight, normalize_length=config.length_normalized_loss)
        self.dropout = nn.Dropout(config.drop_ratio) if config.drop_ratio > 0 else None
        self.speech_token_size = config.speech_token_size
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, past_key_values=None, labels=None,
                use_cache=None, output_hidden_states=None, return_dict=None, batch=None, **kwargs):
        if batch is not None:
            tt, tl = batch["text_token"], batch["text_token_len"]
            st, sl = batch["speech_token"], batch["speech_token_len"]
            target = cosy_lm_target(tl, st, sl, self.speech_token_size, IGNORE_ID).to(tt.device)
            inputs_embeds, attention_mask = pad_unpad_sequence(
                self.llm_embedding.weight[self.sos_eos].reshape(1, 1, -1), self.text_embedding(tt), tl,
                self.llm_embedding.weight[self.task_id].reshape(1, 1, -1), self.speech_embedding(st), sl)
            if self.dropout is not None:
                inputs_embeds = self.dropout(inputs_embeds)
            labels = target[:, 1:].contiguous()
        out = self.model(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                         past_key_values=past_key_values, use_cache=use_cache, output_hidden_states=output_hidden_states,
                         return_dict=True)
        logits = self.lm_head(out.last_hidden_state)
        loss = self.criterion_ce(logits, labels) if labels is not None else None
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states, attentions=None)


class RWKV7XYConfig(RWKV7Config):
    def __init__(self, speech_vocab_size: int = 1024, num_channels: int = 8, lsm_weight: float = 0.0, drop_ratio: float = 0.0,
                 text_shift_size: int = 65536, **kwargs):
        super().__init__(**kwargs)
        self.llm_input_size = kwargs.get("llm_input_size", self.hidden_size)
        self.speech_vocab_size = speech_vocab_size
        self.length_normalized_loss = kwargs.get("length_normalized_loss", True)
        self.lsm_weight = lsm_weight
        self.num_channels = num_channels
        self.drop_ratio = drop_ratio
        self.speech_pad_token = kwargs.get("speech_pad_token", speech_vocab_size - 1)
        self.text_shift_size = text_shift_size


class RWKV7XYLM(RWKV7ForCausalLM):
    config_class = RWKV7XYConfig

    def __init__(self, config: RWKV7XYConfig):
        super().__init__(config)
        sizes = [config.vocab_size] + [config.speech_vocab_size] * (config.num_channels - 1)
        self.embs = nn.ModuleList(nn.Embedding(v, config.hidden_size, padding_idx=v - 1) for v in sizes)
        self.heads = nn.ModuleList(nn.Linear(config.hidden_size, v) for v in sizes)
        self.criterions = nn.ModuleList(nn.CrossEntropyLoss(label_smoothing=config.lsm_weight) for _ in sizes)
        self.dropout = nn.Dropout(config.drop_ratio) if config.drop_ratio > 0 else None
        self.post_init()

    def zero_embs(self):
        with torch.no_grad():
            for e in self.embs:
                if e.padding_idx is not None:
                    e.weight[e.padding_idx].zero_()

    def embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        """Sum over the channels of embs[i](input_ids[:, :, i]) (xy_llm.py:208-214), accumulated in place."""
        if input_ids.dim() != 3 or input_ids.shape[2] != self.config.num_channels:
            raise ValueError(f"input_ids must have shape (B, T, num_channels), but got {tuple(input_ids.shape)}")
        x = self.embs[0](input_ids[:, :, 0])
        for i in range(1, self.config.num_channels):
            x = x + self.embs[i](input_ids[:, :, i])
        return x

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, past_key_values=None, labels=None,
                use_cache=None, output_hidden_states=None, return_dict=None, **kwargs):
        if inputs_embeds is None and input_ids is not None:
            inputs_embeds = self.embed(input_ids)
        if self.dropout is not None:
            inputs_embeds = self.dropout(inputs_embeds)
        out = self.model(inputs_embeds=inputs_embeds, attention_mask=attention_mask, past_key_values=past_key_values,
                         use_cache=use_cache, output_hidden_states=output_hidden_states, return_dict=True)
        h = out.last_hidden_state
        loss, logits = None, []
        if labels is not None and self.training:
            # no channel's logits are held: 8 x (GEMM -> CE kernel -> gradient GEMMs) per token chunk
            loss = xy_channel_losses(h, self.heads, labels, label_smoothing=self.config.lsm_weight)
        else:
            logits = [head(h) for head in self.heads]
            if labels is not None:
                loss = sum(c(lg.reshape(-1, lg.shape[-1]).float(), labels[:, :, i].reshape(-1))
                           for i, (c, lg) in enumerate(zip(self.criterions, logits)))
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states, attentions=None)

    # ---- multi-channel decode loop (xy_llm.py:39-146) ----------------------------------------------------------------
    def is_audio_token(self, token_id: torch.Tensor) -> torch.Tensor:
        lo = self.config.text_shift_size
        return (token_id >= lo) & (token_id < lo + self.config.speech_vocab_size)

    @torch.no_grad()
    def sample(self, input_ids: torch.Tensor, max_length: int, eos_token_id: Optional[int] = None, temperature: float = 1.0,
               top_k: Optional[int] = None, top_p: Optional[float] = None, generator: Optional[torch.Generator] = None,
               reference_termination: bool = False, streamer=None) -> torch.Tensor:
        """The reference's 8-channel sampling loop (`CustomGenerationMixin._sample`, xy_llm.py:39-146) on the recurrent
        cache: per step one forward over the NEW position only (the reference re-feeds through HF's
        prepare_inputs_for_generation), channel 0 constrained to the audio range [text_shift_size, text_shift_size +
        speech_vocab_size) (:86-88), the same warpers on every channel (:91), one multinomial per channel in channel order
        (:95-99: the generator is consumed exactly as the reference consumes it), the flush countdown (:103-115: a
        non-audio token on channel 0 starts `channels - 1` more steps during which channel 0 emits EOS and channel i is
        padded once its last delayed token is out), finished rows keep emitting EOS / pad (:119-124).

        reference_termination=True reproduces the reference's stopping rule literally (:131-133): a row counts as finished
        whenever its countdown value is -1, which is also its value during NORMAL generation, so every row stops after
        its first step.  The default keeps generating until the flush countdown of a row has run out (what the comments
        at :62 and :103 describe) or max_length is reached."""
        from rwkvfla.models.rwkv7.modeling_rwkv7 import _filter_logits
        from rwkvfla.models.utils import Cache
        cfg = self.config
        B, cur, C = input_ids.shape
        pad_idx, lo, nsp = cfg.speech_pad_token, cfg.text_shift_size, cfg.speech_vocab_size
        unfinished = torch.ones(B, dtype=torch.long, device=input_ids.device)
        countdown = torch.full((B,), -1, dtype=torch.long, device=input_ids.device)
        started = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
        cache, feed = Cache(), input_ids
        while True:
            out = self(input_ids=feed, past_key_values=cache, use_cache=True)
            cache = out.past_key_values
            scores = [lg[:, -1, :].clone().float() for lg in out.logits]
            keep = torch.zeros_like(scores[0], dtype=torch.bool)
            keep[:, lo:lo + nsp] = True
            scores[0].masked_fill_(~keep, float("-inf"))
            nxt = []
            for s in scores:
                if temperature != 1.0:
                    s = s / temperature
                s = _filter_logits(s, top_k, top_p)
                nxt.append(torch.multinomial(torch.softmax(s, dim=-1), 1, generator=generator).squeeze(1))
            nxt = torch.stack(nxt, dim=-1)
            start_flush = (~self.is_audio_token(nxt[:, 0])) & (countdown < 0)
            countdown = torch.where(start_flush, torch.full_like(countdown, C - 1), countdown)
            started |= start_flush
            flushing = countdown >= 0
            if eos_token_id is not None:
                nxt[:, 0] = torch.where(flushing, torch.full_like(nxt[:, 0], eos_token_id), nxt[:, 0])
            for i in range(1, C):
                padc = flushing & (countdown < C - i)
                nxt[:, i] = torch.where(padc, torch.full_like(nxt[:, i], pad_idx), nxt[:, i])
            text_fill = eos_token_id if eos_token_id is not None else 0
            nxt[:, 0] = nxt[:, 0] * unfinished + text_fill * (1 - unfinished)
            nxt[:, 1:] = nxt[:, 1:] * unfinished.unsqueeze(-1) + pad_idx * (1 - unfinished.unsqueeze(-1))
            input_ids = torch.cat([input_ids, nxt[:, None, :]], dim=1)
            feed = nxt[:, None, :]
            if streamer is not None:
                streamer.put(nxt[:, 0].cpu())
            countdown = torch.where(flushing, countdown - 1, countdown)
            too_long = input_ids.shape[1] >= max_length
            done_row = (countdown == -1) if reference_termination else (started & (countdown == -1))
            unfinished = unfinished & (~done_row).long()
            if too_long:
                unfinished = torch.zeros_like(unfinished)
            if int(unfinished.max()) == 0:
                break
        if streamer is not None:
            streamer.end()
        return input_ids
