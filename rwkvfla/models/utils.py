"""`rwkvfla.models.utils.Cache` (third_party/cosyvoice/cli/model.py:26; llm.py:250-254; spark_llm.py:81-84):
per-layer recurrent state of the RWKV-7 stack.

`states[layer]` is a dict with
    recurrent_state  fp32 [B,H,64,64]  VALUE-major S[value][key] (layout of the reference's CUDA ops; rwkvfla
                     itself keeps [B,H,K,V] -- `to_fla_layout()` / `from_fla_layout()` transpose once)
    conv_state       [B,C]  last input of the time-mix token shift
    ffn_state        [B,C]  last input of the channel-mix token shift
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch


class Cache:
    is_compileable = False

    def __init__(self, seen_tokens: int = 0, **kwargs: Any) -> None:
        self.states: List[Dict[str, Any]] = []
        self._seen_tokens = seen_tokens

    # -- container protocol -----------------------------------------------------------------
    def __getitem__(self, layer_idx: int) -> Dict[str, Any]:
        if layer_idx < len(self):
            return self.states[layer_idx]
        raise KeyError(f"Cache only has {len(self)} layers, attempted to access layer with index {layer_idx}")

    def __iter__(self):
        yield from self.states

    def __len__(self):
        return len(self.states)

    @property
    def seen_tokens(self) -> int:
        return self._seen_tokens

    def update(self, recurrent_state: Optional[torch.Tensor] = None, attn_state=None,
               conv_state: Optional[torch.Tensor] = None, ffn_state: Optional[torch.Tensor] = None,
               layer_idx: int = 0, offset: Optional[int] = 1, cache_kwargs: Optional[Dict[str, Any]] = None):
        if len(self.states) <= layer_idx:
            while len(self.states) <= layer_idx:
                self.states.append(dict(recurrent_state=None, attn_state=None, conv_state=None, ffn_state=None))
        st = self.states[layer_idx]
        if recurrent_state is not None:
            st["recurrent_state"] = recurrent_state
            if layer_idx == 0 and offset:
                self._seen_tokens += offset
        if conv_state is not None:
            st["conv_state"] = conv_state
        if ffn_state is not None:
            st["ffn_state"] = ffn_state
        return st

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return self._seen_tokens if len(self.states) > (layer_idx or 0) else 0

    def get_max_length(self) -> Optional[int]:
        return None

    def get_max_cache_shape(self) -> Optional[int]:
        return None

    def reset(self):
        self.states.clear()
        self._seen_tokens = 0

    def to_legacy_cache(self):
        return tuple(self.states)

    @classmethod
    def from_legacy_cache(cls, past_key_values=None, seen_tokens: int = 0) -> "Cache":
        if isinstance(past_key_values, cls):
            return past_key_values
        cache = cls(seen_tokens)
        if past_key_values is not None:
            for st in past_key_values:
                cache.states.append(dict(st))
        return cache

    def batch_select(self, idx: torch.Tensor) -> "Cache":
        """Sub-batch of the cache (used when finished sequences are dropped)."""
        out = Cache(self._seen_tokens)
        for st in self.states:
            out.states.append({k: (v.index_select(0, idx) if torch.is_tensor(v) else v) for k, v in st.items()})
        return out

    # -- layout converters --------------------------------------------------------------------
    def to_fla_layout(self) -> List[Dict[str, Any]]:
        """states with recurrent_state as rwkvfla's key-major [B,H,K,V]."""
        return [{**st, "recurrent_state": None if st["recurrent_state"] is None
                 else st["recurrent_state"].transpose(-1, -2).contiguous()} for st in self.states]

    @classmethod
    def from_fla_layout(cls, states, seen_tokens: int = 0) -> "Cache":
        cache = cls(seen_tokens)
        for st in states:
            st = dict(st)
            if st.get("recurrent_state") is not None:
                st["recurrent_state"] = st["recurrent_state"].transpose(-1, -2).contiguous()
            cache.states.append(st)
        return cache
