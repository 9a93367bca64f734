"""Single-file (LDM / "original" layout) checkpoint -> diffusers-layout state dicts, the part of
``StableDiffusion(XL)Pipeline.from_single_file`` (reference trainer/models.py:15-28) the training step needs.

diffusers is not installed here, so its published conversion (diffusers 0.29.2 ``convert_ldm_unet_checkpoint`` /
``convert_ldm_clip_checkpoint`` / ``convert_open_clip_checkpoint`` / ``convert_ldm_vae_checkpoint``) is restated from the
two layouts themselves.  **Parity unpinned**: no real checkpoint is available offline; tests/test_single_file_cpu.py checks
that an LDM-layout dict enumerated independently from the architecture maps onto exactly the key set and shapes of the
diffusers-layout UNet (oracle/unet.py) for SD1.5 and SDXL, and that tensors land where their LDM names say.

LDM UNet layout (``model.diffusion_model.``):
  time_embed.{0,2}; label_emb.0.{0,2} (SDXL); input_blocks.0.0 = conv_in; input_blocks.i.0 = ResBlock or Downsample(op),
  input_blocks.i.1 = SpatialTransformer; middle_block.{0,1,2}; output_blocks.i.{0 ResBlock, 1 SpatialTransformer | Upsample,
  2 Upsample}; out.{0,2}.  ResBlock: in_layers.{0,2}, emb_layers.1, out_layers.{0,3}, skip_connection.  The transformer
  blocks use the same names in both layouts.
"""
from __future__ import annotations

import re
from typing import Dict, Optional, Tuple

import torch

UNET_PREFIX = "model.diffusion_model."
_RES = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj", "out_layers.0": "norm2",
        "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}


def _res(rest: str) -> str:
    for k, v in _RES.items():
        if rest.startswith(k + "."):
            return v + rest[len(k):]
    raise KeyError(f"unknown ResBlock parameter '{rest}'")


def convert_ldm_unet(sd: Dict[str, torch.Tensor], layers_per_block: int = 2) -> Dict[str, torch.Tensor]:
    """``model.diffusion_model.*`` -> diffusers ``UNet2DConditionModel`` names (SD1.5 and SDXL-base)."""
    src = {k[len(UNET_PREFIX):]: v for k, v in sd.items() if k.startswith(UNET_PREFIX)}
    if not src:
        raise KeyError(f"no '{UNET_PREFIX}*' keys: not a single-file (LDM layout) checkpoint")
    lpb = layers_per_block
    # which output_blocks.i.1 are up-samplers (no attention at that level): they hold only conv.{weight,bias}
    out_sub1 = {}
    for k in src:
        m = re.match(r"output_blocks\.(\d+)\.1\.(.+)", k)
        if m:
            out_sub1.setdefault(int(m.group(1)), set()).add(m.group(2))
    out = {}
    for k, v in src.items():
        parts = k.split(".")
        head = parts[0]
        if head == "time_embed":
            out[f"time_embedding.linear_{1 if parts[1] == '0' else 2}.{parts[2]}"] = v
        elif head == "label_emb":
            out[f"add_embedding.linear_{1 if parts[2] == '0' else 2}.{parts[3]}"] = v
        elif head == "out":
            out[f"{'conv_norm_out' if parts[1] == '0' else 'conv_out'}.{parts[2]}"] = v
        elif head == "input_blocks":
            i, sub, rest = int(parts[1]), parts[2], ".".join(parts[3:])
            if i == 0:
                out[f"conv_in.{rest}"] = v
                continue
            blk, lay = (i - 1) // (lpb + 1), (i - 1) % (lpb + 1)
            if sub == "0" and rest.startswith("op."):
                out[f"down_blocks.{blk}.downsamplers.0.conv.{rest[3:]}"] = v
            elif sub == "0":
                out[f"down_blocks.{blk}.resnets.{lay}.{_res(rest)}"] = v
            else:
                out[f"down_blocks.{blk}.attentions.{lay}.{rest}"] = v
        elif head == "middle_block":
            sub, rest = parts[1], ".".join(parts[2:])
            if sub == "1":
                out[f"mid_block.attentions.0.{rest}"] = v
            else:
                out[f"mid_block.resnets.{0 if sub == '0' else 1}.{_res(rest)}"] = v
        elif head == "output_blocks":
            i, sub, rest = int(parts[1]), parts[2], ".".join(parts[3:])
            blk, lay = i // (lpb + 1), i % (lpb + 1)
            if sub == "0":
                out[f"up_blocks.{blk}.resnets.{lay}.{_res(rest)}"] = v
            elif sub == "1" and out_sub1.get(i, set()) <= {"conv.weight", "conv.bias"}:
                out[f"up_blocks.{blk}.upsamplers.0.{rest}"] = v
            elif sub == "1":
                out[f"up_blocks.{blk}.attentions.{lay}.{rest}"] = v
            else:
                out[f"up_blocks.{blk}.upsamplers.0.{rest}"] = v
        else:
            raise KeyError(f"unknown UNet parameter '{k}'")
    return out


def convert_ldm_clip(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    """``cond_stage_model.transformer.`` (SD1.5) / ``conditioner.embedders.0.transformer.`` (SDXL): already transformers names."""
    out = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    out.pop("text_model.embeddings.position_ids", None)
    return out


def convert_open_clip(sd: Dict[str, torch.Tensor], prefix: str = "conditioner.embedders.1.model.") -> Dict[str, torch.Tensor]:
    """OpenCLIP bigG text tower (SDXL's second encoder) -> ``CLIPTextModelWithProjection`` names."""
    src = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    out = {}
    for k, v in src.items():
        if k == "token_embedding.weight":
            out["text_model.embeddings.token_embedding.weight"] = v
        elif k == "positional_embedding":
            out["text_model.embeddings.position_embedding.weight"] = v
        elif k == "text_projection":
            out["text_projection.weight"] = v.t().contiguous()
        elif k.startswith("ln_final."):
            out["text_model.final_layer_norm." + k[len("ln_final."):]] = v
        elif k.startswith("transformer.resblocks."):
            parts = k.split(".")
            n, rest = parts[2], ".".join(parts[3:])
            base = f"text_model.encoder.layers.{n}."
            if rest.startswith("attn.in_proj_"):
                kind = rest[len("attn.in_proj_"):]                       # weight | bias
                d = v.shape[0] // 3
                for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
                    out[f"{base}self_attn.{name}.{kind}"] = v[j * d:(j + 1) * d].contiguous()
            else:
                for a, b in (("ln_1.", "layer_norm1."), ("ln_2.", "layer_norm2."), ("attn.out_proj.", "self_attn.out_proj."),
                             ("mlp.c_fc.", "mlp.fc1."), ("mlp.c_proj.", "mlp.fc2.")):
                    if rest.startswith(a):
                        out[base + b + rest[len(a):]] = v
                        break
                else:
                    raise KeyError(f"unknown OpenCLIP parameter '{k}'")
        elif k in ("logit_scale", "attn_mask"):
            continue
        else:
            raise KeyError(f"unknown OpenCLIP parameter '{k}'")
    return out


def convert_ldm_vae_encoder(sd: Dict[str, torch.Tensor], prefix: str = "first_stage_model.") -> Dict[str, torch.Tensor]:
    """``first_stage_model.{encoder.*, quant_conv.*}`` -> diffusers ``AutoencoderKL`` names (the encode path of
    trainer/dataset.py:141-179; the decoder is not on the training path)."""
    src = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    out = {}
    attn = {"q": "to_q", "k": "to_k", "v": "to_v", "proj_out": "to_out.0", "norm": "group_norm"}
    for k, v in src.items():
        if k.startswith("quant_conv."):
            out[k] = v
            continue
        if not k.startswith("encoder."):
            continue
        rest = k[len("encoder."):]
        parts = rest.split(".")
        if parts[0] in ("conv_in", "conv_out"):
            out[k] = v
        elif parts[0] == "norm_out":
            out["encoder.conv_norm_out." + parts[1]] = v
        elif parts[0] == "down":
            lvl = parts[1]
            if parts[2] == "block":
                name = ".".join(parts[4:]).replace("nin_shortcut", "conv_shortcut")
                out[f"encoder.down_blocks.{lvl}.resnets.{parts[3]}.{name}"] = v
            elif parts[2] == "downsample":
                out[f"encoder.down_blocks.{lvl}.downsamplers.0.conv.{parts[4]}"] = v
        elif parts[0] == "mid":
            if parts[1].startswith("block_"):
                out[f"encoder.mid_block.resnets.{int(parts[1][-1]) - 1}.{'.'.join(parts[2:])}"] = v
            elif parts[1] == "attn_1":
                w = v.reshape(v.shape[0], v.shape[1]) if (v.dim() == 4 and parts[2] != "norm") else v    # 1x1 conv -> linear
                out[f"encoder.mid_block.attentions.0.{attn[parts[2]]}.{parts[3]}"] = w
        else:
            raise KeyError(f"unknown VAE encoder parameter '{k}'")
    return out


def split_single_file(sd: Dict[str, torch.Tensor]) -> Tuple[str, Dict[str, Dict[str, torch.Tensor]]]:
    """Family detection as models.py:15-28 does it (try SDXL, else SD1.5) and the four converted state dicts."""
    family = "sdxl" if any(k.startswith("conditioner.embedders.1.") for k in sd) else "sd15"
    parts = {"unet": convert_ldm_unet(sd), "vae": convert_ldm_vae_encoder(sd)}
    if family == "sdxl":
        parts["text_encoder"] = convert_ldm_clip(sd, "conditioner.embedders.0.transformer.")
        parts["text_encoder_2"] = convert_open_clip(sd)
    else:
        parts["text_encoder"] = convert_ldm_clip(sd, "cond_stage_model.transformer.")
    return family, parts


def is_single_file(sd_keys) -> bool:
    return any(k.startswith(UNET_PREFIX) for k in sd_keys)
