from .rwkv6 import LoRA
from .rwkv7 import RWKV7Attention

__all__ = ["LoRA", "RWKV7Attention"]
