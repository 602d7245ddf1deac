"""rwkvtts_b200: the B200-native WKV-7 hot path of yynil/RWKVTTS behind the reference's own
operator interface.  See DESIGN.md / INTEGRATION.md."""
from . import _lib
from .ops import (CHUNK_LEN, HEAD_SIZE, RUN_CUDA_RWKV7g, RWKV7_BATCH_OP, RWKV7_OP, WKV_7, WKV_7_batch,
                  WindBackstepping, register_torch_ops, wkv7_backward_, wkv7_forward_, wkv7_forward_infer_, wkv7_state_forward_,
                  wkv7_with_state)

register_torch_ops()

__all__ = ["CHUNK_LEN", "HEAD_SIZE", "RUN_CUDA_RWKV7g", "RWKV7_BATCH_OP", "RWKV7_OP", "WKV_7", "WKV_7_batch",
           "WindBackstepping", "register_torch_ops", "wkv7_backward_", "wkv7_forward_", "wkv7_forward_infer_", "wkv7_state_forward_",
           "wkv7_with_state", "_lib"]
