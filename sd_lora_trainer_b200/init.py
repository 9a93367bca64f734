"""Random-init UNet weights in the diffusers naming/shape layout (there are no checkpoints and no network in the
build / bench environment: SURVEY.md 0.5).  Distributions follow torch's nn.Linear / nn.Conv2d defaults
(U(-1/sqrt(fan_in), 1/sqrt(fan_in))), norms start at weight 1 / bias 0."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .arch import UNetArch


def param_shapes(a: UNetArch) -> List[Tuple[str, Tuple[int, ...]]]:
    out: List[Tuple[str, Tuple[int, ...]]] = []
    boc, ted, cd = a.block_out_channels, a.time_embed_dim, a.cross_attention_dim

    def lin(n, i, o, bias=True):
        out.append((f"{n}.weight", (o, i)))
        if bias:
            out.append((f"{n}.bias", (o,)))

    def conv(n, i, o, k):
        out.append((f"{n}.weight", (o, i, k, k)))
        out.append((f"{n}.bias", (o,)))

    def norm(n, c):
        out.append((f"{n}.weight", (c,)))
        out.append((f"{n}.bias", (c,)))

    def resnet(p, cin, cout):
        norm(f"{p}.norm1", cin)
        conv(f"{p}.conv1", cin, cout, 3)
        lin(f"{p}.time_emb_proj", ted, cout)
        norm(f"{p}.norm2", cout)
        conv(f"{p}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{p}.conv_shortcut", cin, cout, 1)

    def transformer(p, c, depth):
        norm(f"{p}.norm", c)
        if a.use_linear_projection:
            lin(f"{p}.proj_in", c, c)
        else:
            conv(f"{p}.proj_in", c, c, 1)
        for j in range(depth):
            b = f"{p}.transformer_blocks.{j}"
            norm(f"{b}.norm1", c)
            for nm, kv in (("attn1", c), ("attn2", cd)):
                if nm == "attn2":
                    norm(f"{b}.norm2", c)
                lin(f"{b}.{nm}.to_q", c, c, bias=False)
                lin(f"{b}.{nm}.to_k", kv, c, bias=False)
                lin(f"{b}.{nm}.to_v", kv, c, bias=False)
                lin(f"{b}.{nm}.to_out.0", c, c)
            norm(f"{b}.norm3", c)
            lin(f"{b}.ff.net.0.proj", c, 8 * c)
            lin(f"{b}.ff.net.2", 4 * c, c)
        if a.use_linear_projection:
            lin(f"{p}.proj_out", c, c)
        else:
            conv(f"{p}.proj_out", c, c, 1)

    conv("conv_in", a.in_channels, boc[0], 3)
    lin("time_embedding.linear_1", boc[0], ted)
    lin("time_embedding.linear_2", ted, ted)
    if a.addition_embed_type == "text_time":
        lin("add_embedding.linear_1", a.projection_class_embeddings_input_dim, ted)
        lin("add_embedding.linear_2", ted, ted)
    cout = boc[0]
    for i in range(len(boc)):
        cin, cout = cout, boc[i]
        for j in range(a.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if a.down_has_attn[i]:
            for j in range(a.layers_per_block):
                transformer(f"down_blocks.{i}.attentions.{j}", cout, a.transformer_layers_per_block[i])
        if i < len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    c = boc[-1]
    resnet("mid_block.resnets.0", c, c)
    resnet("mid_block.resnets.1", c, c)
    transformer("mid_block.attentions.0", c, a.transformer_layers_per_block[-1])
    rev = list(reversed(boc))
    rev_attn = list(reversed(a.down_has_attn))
    cout = rev[0]
    for i in range(len(boc)):
        cprev, cout = cout, rev[i]
        cin = rev[min(i + 1, len(boc) - 1)]
        n = a.layers_per_block + 1
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout)
        if rev_attn[i]:
            ri = len(boc) - 1 - i
            for j in range(n):
                transformer(f"up_blocks.{i}.attentions.{j}", cout, a.transformer_layers_per_block[ri])
        if i < len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    norm("conv_norm_out", boc[0])
    conv("conv_out", boc[0], a.out_channels, 3)
    return out


def random_state_dict(a: UNetArch, seed: int = 0, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    fan: Dict[str, int] = {}
    for name, shape in param_shapes(a):
        base = name.rsplit(".", 1)[0]
        is_norm = ".norm" in name or name.startswith("conv_norm_out") or base.endswith(".norm")
        if is_norm:
            sd[name] = (torch.ones if name.endswith("weight") else torch.zeros)(shape, device=device, dtype=dtype)
            continue
        if name.endswith(".weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            fan[base] = fan_in
        bound = fan[base] ** -0.5
        sd[name] = ((torch.rand(shape, device=device, generator=g) * 2 - 1) * bound).to(dtype)
    return sd


def vae_encoder_param_shapes(block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, in_channels: int = 3,
                             latent_channels: int = 4) -> List[Tuple[str, Tuple[int, ...]]]:
    """diffusers AutoencoderKL keys the VAE-encode prologue reads (``encoder.*`` + ``quant_conv.*``), in vae.py's order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(n, i, o, k):
        out.append((f"{n}.weight", (o, i, k, k)))
        out.append((f"{n}.bias", (o,)))

    def norm(n, c):
        out.append((f"{n}.weight", (c,)))
        out.append((f"{n}.bias", (c,)))

    def resnet(p, cin, cout):
        norm(f"{p}.norm1", cin)
        conv(f"{p}.conv1", cin, cout, 3)
        norm(f"{p}.norm2", cout)
        conv(f"{p}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{p}.conv_shortcut", cin, cout, 1)

    boc = block_out_channels
    conv("encoder.conv_in", in_channels, boc[0], 3)
    cout = boc[0]
    for i, c in enumerate(boc):
        cin, cout = cout, c
        for j in range(layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    c = boc[-1]
    resnet("encoder.mid_block.resnets.0", c, c)
    resnet("encoder.mid_block.resnets.1", c, c)
    a = "encoder.mid_block.attentions.0"
    norm(f"{a}.group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out.append((f"{a}.{n}.weight", (c, c)))
        out.append((f"{a}.{n}.bias", (c,)))
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", c, 2 * latent_channels, 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    return out


def random_vae_encoder_state_dict(seed: int = 0, device="cpu", dtype=torch.float32, **arch) -> Dict[str, torch.Tensor]:
    """Random-init VAE encoder weights (torch default U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norms at 1 / 0)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    fan: Dict[str, int] = {}
    for name, shape in vae_encoder_param_shapes(**arch):
        base = name.rsplit(".", 1)[0]
        if "norm" in base.rsplit(".", 1)[-1]:
            sd[name] = (torch.ones if name.endswith("weight") else torch.zeros)(shape, device=device, dtype=dtype)
            continue
        if name.endswith(".weight"):
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            fan[base] = fan_in
        sd[name] = ((torch.rand(shape, device=device, generator=g) * 2 - 1) * fan[base] ** -0.5).to(dtype)
    return sd
