"""`rwkvfla.models.rwkv7.modeling_rwkv7`: RWKV7Model / RWKV7ForCausalLM / RWKV7Block / RWKV7PreTrainedModel,
plus the re-exports the reference imports from this module (Cache, FusedCrossEntropyLoss,
FusedLinearCrossEntropyLoss; spark_llm.py:7-8).

Parameter names follow rwkv-fla (`model.embeddings`, `model.layers.N.{pre_norm,attn_norm,attn.*,ffn_norm,
ffn.{x_k,key,value}}`, `model.norm`, `lm_head`), so `from_pretrained` of the checkpoints the reference
loads (README.md:140-142) and utils/convert_rwkv.py:17-41 keep working.  `generate` is a native loop
(greedy / temperature / top-k / top-p, eos, left-padded batches) over the recurrent Cache: one
prefill through the chunked kernels, then one stateful step per token.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint
from transformers.generation.utils import GenerationMixin
from transformers.modeling_outputs import BaseModelOutputWithPast, CausalLMOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from rwkvtts_b200 import core, ops
from rwkvtts_b200.decode import MegaDecodeStep, unsupported_reason
from rwkvtts_b200.fused import cached as fused_cached
from rwkvtts_b200.fused import ln_usable as fused_ln_usable
from rwkvtts_b200.fused import usable as fused_usable
from ...layers.rwkv7 import RWKV7Attention
from ...modules import FusedCrossEntropyLoss, FusedLinearCrossEntropyLoss, LayerNorm, l2_warp
from ..utils import Cache
from .configuration_rwkv7 import RWKV7Config


class RWKV7FeedForward(nn.Module):
    """Channel mix: value(relu(key(x + (shift(x) - x) * x_k))**2)  (rwkv_s2s_single_ffn.py:223-230)."""

    def __init__(self, hidden_size: int, hidden_ratio: Optional[int] = None, intermediate_size: Optional[int] = None,
                 hidden_act: str = "sqrelu", layer_idx: int = None, num_hidden_layers: int = None):
        super().__init__()
        assert hidden_act == "sqrelu"
        self.hidden_size = hidden_size
        if hidden_ratio is None:
            hidden_ratio = 4
        if intermediate_size is None:
            intermediate_size = 32 * ((int(hidden_size * hidden_ratio) + 31) // 32)
        self.hidden_ratio, self.intermediate_size = hidden_ratio, intermediate_size
        self.layer_idx, self.num_hidden_layers = layer_idx, num_hidden_layers
        self.x_k = nn.Parameter(torch.zeros(hidden_size))
        self.key = nn.Linear(hidden_size, intermediate_size, bias=False)
        self.value = nn.Linear(intermediate_size, hidden_size, bias=False)
        with torch.no_grad():
            if layer_idx is not None and num_hidden_layers is not None:
                r10 = 1.0 - layer_idx / num_hidden_layers
                ddd = torch.arange(hidden_size, dtype=torch.float32) / hidden_size
                self.x_k.copy_(1.0 - torch.pow(ddd, r10 ** 4))
            nn.init.orthogonal_(self.key.weight)
            self.value.weight.zero_()

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, state: Optional[Cache] = None,
                cu_seqlens=None, use_cache: bool = False, **kwargs):
        am = None
        if attention_mask is not None:
            am = core.mask3(attention_mask, x.shape[1], x.dtype)
        shift = None
        if state is not None and len(state) > self.layer_idx:
            shift = state[self.layer_idx].get("ffn_state")
        out, new_shift = core.cmix(self.x_k, self.key.weight, self.value.weight, x, mask=am, shift_state=shift,
                                   need_state=state is not None and use_cache,
                                   inplace_state=not torch.is_grad_enabled(),
                                   plan=cu_seqlens if isinstance(cu_seqlens, ops.VarlenPlan) else None)
        if state is not None and use_cache:
            state.update(ffn_state=new_shift, layer_idx=self.layer_idx, offset=0)
        return out, state


class RWKV7Block(nn.Module):
    def __init__(self, config: RWKV7Config, layer_idx: int):
        super().__init__()
        self.config, self.layer_idx = config, layer_idx
        norm = lambda: LayerNorm(config.hidden_size, bias=config.norm_bias, eps=config.norm_eps)
        if config.norm_first and layer_idx == 0:
            self.pre_norm = norm()
        self.attn_norm = norm()
        self.attn = RWKV7Attention(
            mode=config.attn_mode, hidden_size=config.hidden_size, head_dim=config.head_dim,
            num_heads=config.num_heads, decay_low_rank_dim=config.decay_low_rank_dim,
            gate_low_rank_dim=config.gate_low_rank_dim, a_low_rank_dim=config.a_low_rank_dim,
            v_low_rank_dim=config.v_low_rank_dim, norm_eps=config.norm_eps, fuse_norm=config.fuse_norm,
            layer_idx=layer_idx, value_dim=config.value_dim[layer_idx], num_hidden_layers=config.num_hidden_layers)
        self.ffn_norm = norm()
        self.ffn = RWKV7FeedForward(hidden_size=config.hidden_size, hidden_ratio=config.hidden_ratio,
                                    intermediate_size=config.intermediate_size, hidden_act=config.hidden_act,
                                    layer_idx=layer_idx, num_hidden_layers=config.num_hidden_layers)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                past_key_values: Optional[Cache] = None, use_cache: Optional[bool] = False,
                output_attentions: Optional[bool] = False, v_first: torch.Tensor = None, cu_seqlens=None,
                residual_in: Optional[torch.Tensor] = None, defer_add: bool = False, **kwargs):
        """`residual_in` / `defer_add` (RWKV7Model's own loop): the block's last operation, `residual + ffn(...)`, is left to
        the NEXT block's attention norm, which adds and normalises in one kernel (`LayerNorm(x, residual, prenorm=True)`: the
        same bf16 sum, one pass over the activations less in the forward and one accumulation kernel less in the backward
        per layer).  With defer_add the block returns the pair (ffn output, residual) instead of their sum."""
        if residual_in is not None:
            hidden_states, residual = self.attn_norm(hidden_states, residual_in, True)
        else:
            residual = self.pre_norm(hidden_states) if hasattr(self, "pre_norm") else hidden_states
            hidden_states = self.attn_norm(residual)
        hidden_states, attentions, past_key_values, v_first = self.attn(
            hidden_states=hidden_states, attention_mask=attention_mask, past_key_values=past_key_values,
            use_cache=use_cache, output_attentions=output_attentions, v_first=v_first, cu_seqlens=cu_seqlens)
        hidden_states, residual = self.ffn_norm(hidden_states, residual, True)
        hidden_states, past_key_values = self.ffn(hidden_states, attention_mask, past_key_values, cu_seqlens,
                                                  use_cache=use_cache)
        if defer_add:
            return (hidden_states, residual), attentions, past_key_values, v_first
        hidden_states = residual + hidden_states
        return hidden_states, attentions, past_key_values, v_first


class RWKV7PreTrainedModel(PreTrainedModel):
    config_class = RWKV7Config
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["RWKV7Block"]
    _supports_cache_class = True
    _skip_keys_device_placement = ["past_key_values"]

    def _init_weights(self, module: nn.Module, rescale_prenorm_residual: bool = True, num_residuals_per_layer: int = 2):
        # time-mix / channel-mix modules initialise themselves (BlinkDL's layer-dependent init); only the
        # embedding and the head are left (rwkv-fla: uniform embedding, orthogonal-ish head)
        if isinstance(module, nn.Embedding):
            nn.init.uniform_(module.weight, a=-1e-4, b=1e-4)
        elif isinstance(module, nn.Linear) and getattr(module, "_is_lm_head", False):
            nn.init.normal_(module.weight, mean=0.0, std=self.config.initializer_range)


def unpack_varlen(x: torch.Tensor, cu_seqlens: torch.Tensor):
    """Packed varlen batch [1, total, D] + cu_seqlens [N+1] (rwkvfla's `cu_seqlens` convention, produced by the reference's
    `*_culens` collators, utils/multiple_jsonl.py:78-135) -> RIGHT-padded [N, Tmax, D] and the flat row index of every packed
    token.  The recurrence, both token shifts and the norms are causal and per sequence, so the padded tail never
    influences a real position and no mask is needed; every sequence starts from a zero state, which is exactly what
    `cu_seqlens` means.  One host read of cu_seqlens (rwkvfla's own chunk index preparation does the same)."""
    if x.shape[0] != 1:
        raise ValueError("cu_seqlens expects a packed batch of size 1")
    cu = [int(v) for v in cu_seqlens.tolist()]
    lens = [b - a for a, b in zip(cu[:-1], cu[1:])]
    if cu[0] != 0 or cu[-1] != x.shape[1] or min(lens) <= 0:
        raise ValueError(f"cu_seqlens {cu} does not partition a packed sequence of length {x.shape[1]}")
    n, tmax = len(lens), max(lens)
    idx = torch.cat([torch.arange(i * tmax, i * tmax + l) for i, l in enumerate(lens)]).to(x.device)
    padded = x.new_zeros(n * tmax, x.shape[-1]).index_copy(0, idx, x[0])
    return padded.view(n, tmax, -1), idx


def repack_varlen(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[N, Tmax, D] -> packed [1, total, D] (inverse of unpack_varlen)."""
    return x.reshape(-1, x.shape[-1]).index_select(0, idx).unsqueeze(0)


class RWKV7Model(RWKV7PreTrainedModel):
    def __init__(self, config: RWKV7Config):
        super().__init__(config)
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        self.embeddings = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([RWKV7Block(config, i) for i in range(config.num_hidden_layers)])
        self.norm = LayerNorm(config.hidden_size, bias=config.norm_bias, eps=config.norm_eps)
        self.gradient_checkpointing = False
        self.post_init()

    def get_input_embeddings(self):
        return self.embeddings

    def set_input_embeddings(self, value):
        self.embeddings = value

    def forward(self, input_ids: Optional[torch.LongTensor] = None, attention_mask: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.FloatTensor] = None, past_key_values: Optional[Cache] = None,
                use_cache: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None,
                cu_seqlens: Optional[torch.LongTensor] = None, **kwargs) -> Union[Tuple, BaseModelOutputWithPast]:
        if output_attentions:
            warnings.warn("`RWKV7Model` does not `output_attentions` now, setting it to `False`.")
        output_attentions = False
        output_hidden_states = output_hidden_states if output_hidden_states is not None else \
            getattr(self.config, "output_hidden_states", False)
        use_cache = use_cache if use_cache is not None else (self.config.use_cache if not self.training else False)
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        if inputs_embeds is None:
            inputs_embeds = self.embeddings(input_ids)
        hidden_states = inputs_embeds
        packed_idx = None
        if cu_seqlens is not None:
            if attention_mask is not None or (past_key_values is not None and len(past_key_values) > 0):
                raise NotImplementedError("cu_seqlens together with attention_mask / a filled cache")
            use_cache = False                  # a packed batch carries no per-sequence state out
            if core.FUSED and fused_usable(hidden_states) and hidden_states.shape[0] == 1:
                # packed all the way down: the chunked kernels and the token shifts restart at every boundary, no padding
                # position is computed or moved (one plan for all layers, built on the device without a host sync)
                cu_seqlens = ops.VarlenPlan(cu_seqlens, hidden_states.shape[1])
            else:
                # CPU / fp32 / ATen chain: the equivalent right-padded batch (see unpack_varlen)
                hidden_states, packed_idx = unpack_varlen(hidden_states, cu_seqlens)
                cu_seqlens = None
        if use_cache and not isinstance(past_key_values, Cache):
            past_key_values = Cache.from_legacy_cache(past_key_values)
        all_hidden_states = () if output_hidden_states else None
        v_first = torch.zeros_like(hidden_states)
        # between blocks the residual add is deferred into the next norm (RWKV7Block.forward) unless somebody wants to see
        # the hidden states of every layer or the fused add+norm kernel does not apply
        defer = (not output_hidden_states and core.FUSED and fused_ln_usable(hidden_states)
                 and not (self.gradient_checkpointing and self.training))
        residual = None
        for li, layer in enumerate(self.layers):
            if output_hidden_states:
                all_hidden_states += (hidden_states,)
            if self.gradient_checkpointing and self.training:
                hidden_states, _, past_key_values, v_first = checkpoint(
                    layer, hidden_states, attention_mask, past_key_values, use_cache, False, v_first, cu_seqlens,
                    use_reentrant=False)
            else:
                hidden_states, _, past_key_values, v_first = layer(
                    hidden_states, attention_mask=attention_mask, past_key_values=past_key_values,
                    use_cache=use_cache, output_attentions=False, v_first=v_first, cu_seqlens=cu_seqlens,
                    residual_in=residual, defer_add=defer)
                if defer:
                    hidden_states, residual = hidden_states
        hidden_states = self.norm(hidden_states, residual) if residual is not None else self.norm(hidden_states)
        if packed_idx is not None:
            hidden_states = repack_varlen(hidden_states, packed_idx)
            if output_hidden_states:
                all_hidden_states = tuple(repack_varlen(h, packed_idx) for h in all_hidden_states)
        if output_hidden_states:
            all_hidden_states += (hidden_states,)
        if not return_dict:
            return tuple(i for i in [hidden_states, past_key_values, all_hidden_states] if i is not None)
        return BaseModelOutputWithPast(last_hidden_state=hidden_states, past_key_values=past_key_values,
                                       hidden_states=all_hidden_states, attentions=None)


def _filter_logits(logits: torch.Tensor, top_k: Optional[int], top_p: Optional[float]) -> torch.Tensor:
    if top_k is not None and 0 < top_k < logits.shape[-1]:
        kth = torch.topk(logits, top_k, dim=-1).values[..., -1, None]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and 0.0 < top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=False, dim=-1)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -1:] = False
        logits = logits.masked_fill(remove.scatter(-1, sorted_idx, remove), float("-inf"))
    return logits


class _GraphDecodeStep:
    """One decode step of RWKV7ForCausalLM (token -> logits, every layer's state advanced in place) captured in a CUDA
    graph: the step is ~25 small kernels per layer, so on the eager path it is bound by launch and Python overhead
    (SURVEY section 8 row a12).  The recurrent state is already advanced in place by the stateful WKV op; the token-shift
    states are copied into static buffers inside the graph so that every replay reads what the previous one wrote."""

    def __init__(self, model: "RWKV7ForCausalLM", cache: Cache, batch: int, device):
        self.model, self.cache = model, cache
        self.tok = torch.zeros(batch, 1, dtype=torch.long, device=device)
        self.logits = None
        keys = ("conv_state", "ffn_state")
        self.static = [{k: st[k].clone() for k in keys if torch.is_tensor(st.get(k))} for st in cache.states]
        for st, sb in zip(cache.states, self.static):
            st.update(sb)
        # warm-up on a side stream (cuBLAS workspaces, lazy module state), then put the states back
        saved = [{k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for st in cache.states]
        seen = cache._seen_tokens
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream(device).wait_stream(side)
        for st, sv in zip(cache.states, saved):
            for k, v in sv.items():
                st[k].copy_(v)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        cache._seen_tokens = seen

    def _step(self):
        out = self.model(input_ids=self.tok, past_key_values=self.cache, use_cache=True, logits_to_keep=1)
        for st, sb in zip(self.cache.states, self.static):
            for k, buf in sb.items():
                if st[k] is not buf:
                    buf.copy_(st[k])
                    st[k] = buf
        lg = out.logits[:, -1].float()
        if self.logits is None:
            self.logits = torch.empty_like(lg)
        self.logits.copy_(lg)

    def __call__(self, nxt: torch.Tensor) -> torch.Tensor:
        self.tok.copy_(nxt.view(-1, 1))
        self.graph.replay()
        self.cache._seen_tokens += 1
        return self.logits


class RWKV7ForCausalLM(RWKV7PreTrainedModel, GenerationMixin):
    _tied_weights_keys = ["lm_head.weight"]

    def __init__(self, config):
        super().__init__(config)
        self.model = RWKV7Model(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.lm_head._is_lm_head = True
        self.criterion = None
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embeddings

    def set_input_embeddings(self, value):
        self.model.embeddings = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        self.lm_head = new_embeddings

    def set_decoder(self, decoder):
        self.model = decoder

    def get_decoder(self):
        return self.model

    def prepare_inputs_for_generation(self, input_ids=None, past_key_values=None, attention_mask=None,
                                      inputs_embeds=None, use_cache=True, logits_to_keep=None, **kwargs):
        if past_key_values is not None and len(past_key_values) > 0:
            input_ids = input_ids[:, -1:]
        if inputs_embeds is not None and (past_key_values is None or len(past_key_values) == 0):
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids.contiguous()}
        model_inputs.update({"past_key_values": past_key_values, "use_cache": use_cache,
                             "attention_mask": attention_mask, "logits_to_keep": logits_to_keep})
        return model_inputs

    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.Tensor] = None, past_key_values: Optional[Cache] = None,
                labels: Optional[torch.LongTensor] = None, shift_labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None,
                logits_to_keep: Optional[int] = 0, **kwargs) -> Union[Tuple, CausalLMOutputWithPast]:
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        outputs = self.model(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                             past_key_values=past_key_values, use_cache=use_cache,
                             output_hidden_states=output_hidden_states, return_dict=True,
                             cu_seqlens=kwargs.get("cu_seqlens"))
        hidden_states = outputs.last_hidden_state
        loss, logits = None, None
        has_labels = labels is not None or shift_labels is not None
        if not (self.config.fuse_linear_cross_entropy and has_labels):
            h = hidden_states if not logits_to_keep else hidden_states[:, -logits_to_keep:]
            W = self.lm_head.weight
            if W.is_cuda and W.dtype == torch.bfloat16 and W.shape[0] % 8 != 0 and self.lm_head.bias is None:
                # an odd vocabulary (Spark: 8193) sends cuBLAS to an unaligned legacy kernel, 5x slower on B200:
                # multiply by the weight zero-padded to a multiple of 8 rows and return the view without the padding
                pad = lambda: F.pad(W, (0, 0, 0, 8 - W.shape[0] % 8))
                Wp = pad() if torch.is_grad_enabled() else fused_cached(W, (W,), "lm_head_pad8", lambda: pad().detach())
                logits = F.linear(h, Wp)[..., :W.shape[0]]
            else:
                logits = self.lm_head(h)
        if has_labels:
            criterion = self.criterion
            if criterion is None:
                if self.config.fuse_linear_cross_entropy:
                    criterion = FusedLinearCrossEntropyLoss(use_l2warp=self.config.use_l2warp)
                elif self.config.fuse_cross_entropy:
                    criterion = FusedCrossEntropyLoss(inplace_backward=True)
                else:
                    criterion = nn.CrossEntropyLoss()
            if shift_labels is None:
                shift_labels = torch.cat((labels[..., 1:], torch.full_like(labels[:, :1], criterion.ignore_index)), 1)
            shift_labels = shift_labels.to(hidden_states.device)
            if self.config.fuse_linear_cross_entropy:
                loss = criterion(hidden_states, shift_labels, self.lm_head.weight, self.lm_head.bias)
            else:
                loss = criterion(logits.view(shift_labels.numel(), -1), shift_labels.view(-1))
                loss = l2_warp(loss, logits) if self.config.use_l2warp else loss
        if not return_dict:
            output = (logits, outputs.past_key_values)
            return (loss,) + output if loss is not None else output
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=outputs.past_key_values,
                                      hidden_states=outputs.hidden_states, attentions=None)

    # ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids: Optional[torch.LongTensor] = None, inputs_embeds: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, max_new_tokens: Optional[int] = None,
                 max_length: Optional[int] = None, min_new_tokens: int = 0, do_sample: bool = False,
                 temperature: float = 1.0, top_k: Optional[int] = None, top_p: Optional[float] = None,
                 eos_token_id: Union[int, List[int], None] = None, pad_token_id: Optional[int] = None,
                 use_cache: bool = True, generator: Optional[torch.Generator] = None,
                 return_dict_in_generate: bool = False, use_cuda_graph: Optional[bool] = None,
                 eos_check_interval: int = 1, exact: bool = False, use_megakernel: Optional[bool] = None, **kwargs):
        """Autoregressive decode over the recurrent Cache (the call inference/rwkv7speech_inference.py and
        spark_llm.py:54-102 make).  Returns [B, prompt + new] token ids when `input_ids` is given and
        [B, new] when only `inputs_embeds` is given (HF convention); finished rows are padded with
        `pad_token_id`.  On CUDA the per-token step runs as one CUDA graph (`use_cuda_graph`, default on when more than 8
        tokens are requested); `eos_check_interval` > 1 polls the all-finished flag (a host sync) only every that many
        steps (finished rows are padded either way, so the result does not change).  `use_megakernel` (default: on
        whenever the model fits it and neither `exact` nor `use_cuda_graph=False` is asked for): the per-token step is ONE
        persistent kernel over all layers (rwkvtts_b200/decode.py, csrc/decode_step.cu) instead of the ~430-node graph, and
        a greedy decode also samples on the device -- n tokens are n launches."""
        if exact:
            # the reference decode step operation for operation (core.exact_mode): greedy ids bit-identical to the
            # reference loop (rwkv_asr_cuda_whisper.py:694-717), at the reference's eager speed plus the CUDA graph
            from rwkvtts_b200 import core as _core
            with _core.exact_mode():
                return self.generate(input_ids=input_ids, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                                     max_new_tokens=max_new_tokens, max_length=max_length, min_new_tokens=min_new_tokens,
                                     do_sample=do_sample, temperature=temperature, top_k=top_k, top_p=top_p,
                                     eos_token_id=eos_token_id, pad_token_id=pad_token_id, use_cache=use_cache,
                                     generator=generator, return_dict_in_generate=return_dict_in_generate,
                                     use_cuda_graph=use_cuda_graph, eos_check_interval=eos_check_interval,
                                     use_megakernel=False, **kwargs)
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("pass exactly one of input_ids / inputs_embeds")
        prompt_len = input_ids.shape[1] if input_ids is not None else inputs_embeds.shape[1]
        if max_new_tokens is None:
            max_new_tokens = (max_length - prompt_len) if max_length is not None else 20
        eos = eos_token_id if eos_token_id is not None else self.config.eos_token_id
        eos = [] if eos is None else ([eos] if isinstance(eos, int) else list(eos))
        pad = pad_token_id if pad_token_id is not None else (self.config.pad_token_id
                                                             if self.config.pad_token_id is not None else (eos[0] if eos else 0))
        if do_sample and top_k is None:
            top_k = 50                                   # HF default
        B = input_ids.shape[0] if input_ids is not None else inputs_embeds.shape[0]
        dev = input_ids.device if input_ids is not None else inputs_embeds.device
        cache = Cache()
        out = self(input_ids=input_ids, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                   past_key_values=cache, use_cache=True, logits_to_keep=1)
        cache = out.past_key_values
        eos_t = torch.tensor(eos, device=dev, dtype=torch.long) if eos else None
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        new_tokens = []
        logits = out.logits[:, -1].float()
        if use_cuda_graph is None:
            use_cuda_graph = dev.type == "cuda" and max_new_tokens > 8
        graph_step, mega = None, None
        auto_mega = use_megakernel is None
        if auto_mega:
            use_megakernel = (dev.type == "cuda" and use_cuda_graph and max_new_tokens > 1 and core.FUSED and not core.EXACT
                              and len(eos) <= 8 and unsupported_reason(self, B) is None)
        if use_megakernel and max_new_tokens > 1:
            try:
                mega = graph_step = MegaDecodeStep(self, cache, B, dev)   # raises if the model does not fit the kernel
            except (ValueError, RuntimeError) as e:
                if not auto_mega:
                    raise
                warnings.warn(f"one-kernel decode step unavailable ({e}); using the CUDA-graph step")
        if mega is None and use_cuda_graph and max_new_tokens > 1:
            try:
                graph_step = _GraphDecodeStep(self, cache, B, dev)
            except RuntimeError as e:                  # e.g. an op that cannot be captured in a user subclass: decode eagerly
                warnings.warn(f"CUDA-graph decode step unavailable ({e}); falling back to the eager step")
                torch.cuda.synchronize(dev)
        for step in range(max_new_tokens):
            if eos_t is not None and step < min_new_tokens:
                logits[:, eos_t] = float("-inf")
            if do_sample:
                lg = logits / max(temperature, 1e-6)
                probs = _filter_logits(lg, top_k, top_p).softmax(dim=-1)
                nxt = torch.multinomial(probs, 1, generator=generator).squeeze(-1)
            else:
                nxt = logits.argmax(dim=-1)
            nxt = torch.where(done, torch.full_like(nxt, pad), nxt)
            new_tokens.append(nxt)
            if eos_t is not None:
                done = done | torch.isin(nxt, eos_t)
                if (step + 1) % max(eos_check_interval, 1) == 0 and bool(done.all()):
                    break
            if mega is not None and not do_sample and step + 1 < max_new_tokens:
                # greedy: the rest of the loop runs on the device (arg-max, EOS masking, finished rows, padding), one
                # launch per token; the host only polls the all-finished flag every 64 tokens
                left, s0, cur = max_new_tokens - 1 - step, step + 1, nxt
                while left > 0:
                    n = min(left, 64) if eos else left
                    chunk = mega.greedy(cur, n, eos, pad, min_new_tokens, step0=s0, done=done)
                    new_tokens.extend(chunk.unbind(0))
                    cur, s0, left = chunk[-1], s0 + n, left - n
                    if eos and bool(done.all()):
                        break
                if eos:       # cut where the host loop would have stopped: the first polled step with every row finished
                    gen = torch.stack(new_tokens, dim=1)
                    fin = (torch.isin(gen, eos_t).cumsum(dim=1) > 0).all(dim=0)
                    k = max(eos_check_interval, 1)
                    polled = (torch.arange(1, gen.shape[1] + 1, device=dev) % k) == 0
                    hit = torch.nonzero(fin & polled)
                    if hit.numel():
                        new_tokens = new_tokens[: int(hit[0]) + 1]
                break
            if step + 1 < max_new_tokens:
                if graph_step is not None:
                    logits = graph_step(nxt).clone()
                else:
                    out = self(input_ids=nxt[:, None], past_key_values=cache, use_cache=True, logits_to_keep=1)
                    cache = out.past_key_values
                    logits = out.logits[:, -1].float()
        gen = torch.stack(new_tokens, dim=1) if new_tokens else torch.empty(B, 0, dtype=torch.long, device=dev)
        seq = torch.cat([input_ids, gen], dim=1) if input_ids is not None else gen
        if return_dict_in_generate:
            return {"sequences": seq, "past_key_values": cache}
        return seq


__all__ = ["RWKV7Config", "RWKV7Model", "RWKV7ForCausalLM", "RWKV7PreTrainedModel", "RWKV7Block", "Cache",
           "FusedLinearCrossEntropyLoss", "FusedCrossEntropyLoss", "RWKV7FeedForward"]
