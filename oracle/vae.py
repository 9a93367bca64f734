"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Pure-torch restatement of what ``self.vae_encoder.encode(image).latent_dist`` computes in the reference's one-time
dataset prologue (trainer/dataset.py:141-179, ``_process``; the VAE is kept in fp32, main.py:186): diffusers==0.29.2
``AutoencoderKL.encode`` = ``Encoder`` (conv_in -> 4 x DownEncoderBlock2D -> UNetMidBlock2D with one single-head
attention -> GroupNorm/SiLU/conv_out, double_z) followed by the 1x1 ``quant_conv``; the result is the ``parameters``
tensor ``[B, 8, h, w]`` (mean | logvar) that ``DiagonalGaussianDistribution`` wraps.  Both model families use the same
graph (``vae/config.json`` of SD1.5 and SDXL-base: block_out_channels (128, 256, 512, 512), layers_per_block 2,
norm_num_groups 32, latent_channels 4); they differ only in weights and ``scaling_factor``.  Parameter names are
diffusers' so a real ``vae`` state dict loads unchanged (decoder / post_quant_conv keys are ignored).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class VAEConfig:
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(32, 64, 64), layers_per_block=1, norm_num_groups=8)


class ResnetBlock(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps=1e-6, swish, output_scale_factor=1."""

    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample(nn.Module):
    """Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownBlock(nn.Module):
    def __init__(self, cin: int, cout: int, layers: int, groups: int, add_downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(cin if j == 0 else cout, cout, groups) for j in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class MidAttention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, norm_num_groups, eps=1e-6, residual_connection=True, bias=True)."""

    def __init__(self, c: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x.view(B, C, H * W)).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(B, C, H, W)
        return o + x


class MidBlock(nn.Module):
    def __init__(self, c: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(c, c, groups), ResnetBlock(c, c, groups)])
        self.attentions = nn.ModuleList([MidAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        blocks, cout = [], boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            blocks.append(DownBlock(cin, cout, cfg.layers_per_block, g, add_downsample=i < len(boc) - 1))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = MidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoder(nn.Module):
    """The encode half of AutoencoderKL: ``moments = quant_conv(encoder(x))``."""

    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)

    @torch.no_grad()
    def encode_moments(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, 3, H, W] in [-1, 1] -> DiagonalGaussianDistribution.parameters [B, 8, H/8, W/8]."""
        return self.quant_conv(self.encoder(image))


def diagonal_gaussian_sample(parameters: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """DiagonalGaussianDistribution.sample() with the draw injected: mean + exp(0.5 * clamp(logvar, -30, 20)) * eps."""
    mean, logvar = torch.chunk(parameters, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps


def build_vae(cfg: VAEConfig = VAEConfig(), seed: int = 0, dtype=torch.float32) -> AutoencoderKLEncoder:
    """Random-init encoder (torch default initialisers; there are no checkpoints in the build environment)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = AutoencoderKLEncoder(cfg).to(dtype).eval()
    torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def state_dict_of(m: AutoencoderKLEncoder) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
