"""Mirror of the reference's trainer/ti_cross_attn_loss.py: the cross-attention score hook.

The reference replaces the processor of every attn2 in down_blocks / up_blocks with DAAMLossAttnProcessor2_0, which
runs SDPA and ALSO an explicit per-head QK^T/sqrt(d) GEMM summed over heads (:201-212).  The B200 executor emits the
same tensor from one full-width GEMM inside the attention forward (unet.Attn.fwd) and folds its gradient into dQ/dK."""
from __future__ import annotations

from typing import List

from .loss import process_and_stack_attention_scores


class DAAMLoss:
    def __init__(self, unet):
        self.unet = unet
        self.layer_names = [f"hooked_attn2_{i}" for i in range(len(unet.hooked))]

    def get_all_cross_attention_scores(self) -> List:
        return [a.scores for a in self.unet.hooked]

    def process_and_stack_attention_scores(self, img_ratio: float):
        return process_and_stack_attention_scores(self.get_all_cross_attention_scores(), img_ratio)


def find_attnprocessor2_0(unet) -> List[str]:
    """ti_cross_attn_loss.py:88-112: down_blocks then up_blocks, never the mid block."""
    return [f"hooked_attn2_{i}" for i in range(len(unet.hooked))]


def init_daam_loss(pipeline):
    """ti_cross_attn_loss.py:336-364: install the hook on every eligible layer, return (pipeline, daam_loss)."""
    print(f"Found: {len(pipeline.unet.hooked)} modules")
    pipeline.unet.set_capture(True)
    return pipeline, DAAMLoss(pipeline.unet)
