from .rwkv7 import RWKV7Config, RWKV7ForCausalLM, RWKV7Model
from .utils import Cache

__all__ = ["RWKV7Config", "RWKV7ForCausalLM", "RWKV7Model", "Cache"]
