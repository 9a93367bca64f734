"""UNet graph descriptions for the two model families the reference trains (trainer/models.py:15-28):
SD1.5 and SDXL-base, as published in their diffusers ``unet/config.json``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple


@dataclass(frozen=True)
class UNetArch:
    name: str = "sdxl"
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_has_attn: Tuple[bool, ...] = (False, True, True)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20)
    cross_attention_dim: int = 2048
    use_linear_projection: bool = True
    addition_embed_type: Optional[str] = "text_time"
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32
    vae_scaling_factor: float = 0.13025

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


def sdxl() -> UNetArch:
    return UNetArch()


def sd15() -> UNetArch:
    return UNetArch(name="sd15", block_out_channels=(320, 640, 1280, 1280), down_has_attn=(True, True, True, False),
                    transformer_layers_per_block=(1, 1, 1, 1), num_attention_heads=(8, 8, 8, 8),
                    cross_attention_dim=768, use_linear_projection=False, addition_embed_type=None,
                    vae_scaling_factor=0.18215)


def tiny_sdxl() -> UNetArch:
    return UNetArch(name="tiny_sdxl", block_out_channels=(64, 128, 256), down_has_attn=(False, True, True),
                    layers_per_block=1, transformer_layers_per_block=(1, 1, 2), num_attention_heads=(1, 2, 4),
                    cross_attention_dim=128, addition_time_embed_dim=32,
                    projection_class_embeddings_input_dim=64 + 6 * 32)


def tiny_sd15() -> UNetArch:
    return UNetArch(name="tiny_sd15", block_out_channels=(64, 128, 128), down_has_attn=(True, True, False),
                    layers_per_block=1, transformer_layers_per_block=(1, 1, 1), num_attention_heads=(2, 2, 2),
                    cross_attention_dim=96, use_linear_projection=False, addition_embed_type=None,
                    vae_scaling_factor=0.18215)


def by_name(name: str) -> UNetArch:
    return {"sdxl": sdxl, "sd15": sd15, "tiny_sdxl": tiny_sdxl, "tiny_sd15": tiny_sd15}[name]()
