"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Pure-torch restatement of diffusers==0.29.2 ``UNet2DConditionModel`` for the two
architectures the reference trains (SD1.5 and SDXL-base), with diffusers-
compatible parameter names so a real checkpoint's ``unet`` state dict loads
unchanged.  Third-party semantics restated from SURVEY.md Appendix A/B; call
sites in the reference: trainer/models.py:17-32 (construction),
main.py:329-336 (forward), trainer/ti_cross_attn_loss.py:130-230 (the
cross-attention processor with the head-summed score capture).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Subset of diffusers' unet/config.json that shapes the graph (Appendix A)."""
    name: str = "sdxl"
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    # True = CrossAttn{Down,Up}Block2D, False = {Down,Up}Block2D (down order)
    down_has_attn: Tuple[bool, ...] = (False, True, True)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    # heads per block (diffusers' mis-named ``attention_head_dim``)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20)
    cross_attention_dim: int = 2048
    use_linear_projection: bool = True
    addition_embed_type: Optional[str] = "text_time"
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32
    time_embed_dim_mult: int = 4
    vae_scaling_factor: float = 0.13025

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * self.time_embed_dim_mult

    @staticmethod
    def sdxl() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig(
            name="sd15", block_out_channels=(320, 640, 1280, 1280),
            down_has_attn=(True, True, True, False), layers_per_block=2,
            transformer_layers_per_block=(1, 1, 1, 1), num_attention_heads=(8, 8, 8, 8),
            cross_attention_dim=768, use_linear_projection=False, addition_embed_type=None,
            vae_scaling_factor=0.18215)

    @staticmethod
    def tiny_sdxl() -> "UNetConfig":
        """Reduced-depth/width SDXL-shaped graph for CPU-speed tests (same code paths)."""
        return UNetConfig(
            name="tiny_sdxl", block_out_channels=(64, 128, 256), down_has_attn=(False, True, True),
            layers_per_block=1, transformer_layers_per_block=(1, 1, 2), num_attention_heads=(1, 2, 4),
            cross_attention_dim=128, use_linear_projection=True, addition_embed_type="text_time",
            addition_time_embed_dim=32, projection_class_embeddings_input_dim=64 + 6 * 32)

    @staticmethod
    def tiny_sd15() -> "UNetConfig":
        return UNetConfig(
            name="tiny_sd15", block_out_channels=(64, 128, 128), down_has_attn=(True, True, False),
            layers_per_block=1, transformer_layers_per_block=(1, 1, 1), num_attention_heads=(2, 2, 2),
            cross_attention_dim=96, use_linear_projection=False, addition_embed_type=None,
            vae_scaling_factor=0.18215)


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)``; fp32 out."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers ``Attention`` with AttnProcessor2_0; ``capture`` switches on the
    reference's DAAMLossAttnProcessor2_0 score capture (ti_cross_attn_loss.py:201-212)."""

    def __init__(self, dim: int, heads: int, cross_dim: Optional[int]):
        super().__init__()
        self.heads = heads
        kv = cross_dim if cross_dim is not None else dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv, dim, bias=False)
        self.to_v = nn.Linear(kv, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
        self.capture = False
        self.cross_attention_scores = None

    def forward(self, x, ehs=None):
        b, l, _ = x.shape
        q = self.to_q(x)
        src = x if ehs is None else ehs
        k = self.to_k(src)
        v = self.to_v(src)
        d = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        if self.capture and ehs is not None:
            scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)
            self.cross_attention_scores = scores.sum(dim=1)      # einops Reduce "b h i t -> b i t", sum
        o = o.transpose(1, 2).reshape(b, -1, self.heads * d).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, None)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ehs):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ehs) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, depth: int, cross_dim: int, groups: int, linear_proj: bool):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, heads, cross_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ehs):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x)
        if self.linear_proj:
            y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
            y = self.proj_in(y)
        else:
            y = self.proj_in(y)
            y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, ehs)
        if self.linear_proj:
            y = self.proj_out(y)
            y = y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        else:
            y = y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
            y = self.proj_out(y)
        return y + res


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, i: int, cin: int, cout: int, has_attn: bool, add_down: bool):
        super().__init__()
        n = cfg.layers_per_block
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if j == 0 else cout, cout, cfg.time_embed_dim, cfg.norm_num_groups) for j in range(n)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, cfg.num_attention_heads[i], cfg.transformer_layers_per_block[i],
                                cfg.cross_attention_dim, cfg.norm_num_groups, cfg.use_linear_projection)
             for _ in range(n)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ehs):
        outs = []
        for j, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        c = cfg.block_out_channels[-1]
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, cfg.time_embed_dim, cfg.norm_num_groups) for _ in range(2)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(c, cfg.num_attention_heads[-1], cfg.transformer_layers_per_block[-1],
                                cfg.cross_attention_dim, cfg.norm_num_groups, cfg.use_linear_projection)])

    def forward(self, x, temb, ehs):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, i: int, cin: int, cout: int, cprev: int, has_attn: bool, add_up: bool):
        super().__init__()
        n = cfg.layers_per_block + 1
        ri = len(cfg.block_out_channels) - 1 - i
        rs = []
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, cfg.time_embed_dim, cfg.norm_num_groups))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, cfg.num_attention_heads[ri], cfg.transformer_layers_per_block[ri],
                                cfg.cross_attention_dim, cfg.norm_num_groups, cfg.use_linear_projection)
             for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb, ehs):
        for j, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)
        downs, cout = [], boc[0]
        for i in range(len(boc)):
            cin, cout = cout, boc[i]
            downs.append(DownBlock(cfg, i, cin, cout, cfg.down_has_attn[i], add_down=i < len(boc) - 1))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg)
        rev = list(reversed(boc))
        rev_attn = list(reversed(cfg.down_has_attn))
        ups, cout = [], rev[0]
        for i in range(len(boc)):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            ups.append(UpBlock(cfg, i, cin, cout, cprev, rev_attn[i], add_up=i < len(boc) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timesteps, encoder_hidden_states, timestep_cond=None,
                added_cond_kwargs=None, return_dict=False):
        cfg = self.cfg
        t_emb = timestep_embedding(timesteps, cfg.block_out_channels[0]).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        if cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            te = timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim)
            te = te.reshape(text_embeds.shape[0], -1)
            add = torch.cat([text_embeds, te], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x,)


# --- the reference's cross-attention hook (ti_cross_attn_loss.py:88-112, 336-364) -------------

def hooked_attention_modules(unet: nn.Module) -> List[Tuple[str, Attention]]:
    """attn2 of every transformer block in down_blocks/up_blocks, NOT mid_block
    (find_attnprocessor2_0, ti_cross_attn_loss.py:97), in the reference's
    enumeration order (block type, block, attention, transformer block)."""
    found = []
    for bt in ("down_blocks", "up_blocks"):
        blocks = getattr(unet, bt)
        for bi, blk in enumerate(blocks):
            if getattr(blk, "attentions", None) is None:
                continue
            for ai, tr in enumerate(blk.attentions):
                for ti, tb in enumerate(tr.transformer_blocks):
                    found.append((f"{bt}.{bi}.attentions.{ai}.transformer_blocks.{ti}.attn2", tb.attn2))
    return found


def init_score_capture(unet: nn.Module) -> List[Attention]:
    mods = [m for _, m in hooked_attention_modules(unet)]
    for m in mods:
        m.capture = True
    return mods
