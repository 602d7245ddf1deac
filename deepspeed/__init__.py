"""`deepspeed` surface used by yynil/RWKVTTS's training scripts, served by rwkvtts_b200.engine (ZeRO-2
semantics on NCCL, fused Adam on the rank's shard).  DeepSpeed itself is not a dependency of this
repository; put this directory AFTER a real DeepSpeed on sys.path if you want the original."""
from rwkvtts_b200.engine import Engine as DeepSpeedEngine
from rwkvtts_b200.engine import init_distributed, initialize
from . import checkpointing, ops

__version__ = "0.16.7+rwkvtts_b200"
__all__ = ["initialize", "init_distributed", "DeepSpeedEngine", "checkpointing", "ops"]
