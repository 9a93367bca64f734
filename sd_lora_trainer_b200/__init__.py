"""B200-native LoRA + textual-inversion training step for SD1.5 / SDXL UNets (sm_100a kernels behind a C ABI)."""
from . import arch  # noqa: F401

__all__ = ["arch"]
