from .configuration_rwkv7 import RWKV7Config
from .modeling_rwkv7 import RWKV7Block, RWKV7ForCausalLM, RWKV7Model, RWKV7PreTrainedModel

__all__ = ["RWKV7Config", "RWKV7ForCausalLM", "RWKV7Model", "RWKV7PreTrainedModel", "RWKV7Block"]
