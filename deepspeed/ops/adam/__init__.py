from rwkvtts_b200.engine import DeepSpeedCPUAdam, FusedAdam

__all__ = ["FusedAdam", "DeepSpeedCPUAdam"]
