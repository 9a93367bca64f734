"""Mirror of the reference's trainer/checkpoint.py writer: the drop-in FILE surface of a run.

``save_checkpoint`` (checkpoint.py:104-221) leaves, in ``output_dir``:

  * ``{name}_{version}_embeddings.safetensors``  - the trained token rows, keys ``clip_l`` / ``clip_g``
                                                   (embedding_handler.py:401-422)
  * ``special_params.json``                       - the token dictionary (checkpoint.py:163-168)
  * ``adapter_config.json``                       - what ``unet.save_pretrained`` of a PEFT model writes (:175)
  * ``{name}_{version}_lora.safetensors``         - the LoRA factors under kohya / WebUI keys (:84-102, 206-209)
  * is_lora=False (:211-213): ``config.json`` + ``diffusion_pytorch_model.safetensors`` instead of the two LoRA files

and deletes the intermediate ``pytorch_lora_weights.safetensors`` (:215-219), which is therefore never written here.

peft / diffusers are not in this image, so their key conversions are RESTATED (parity unpinned, as for the oracle):
  get_peft_model_state_dict        ``base_model.model.<path>.lora_A.default.weight`` -> adapter name dropped [3P peft 0.10.0]
  convert_state_dict_to_diffusers  -> ``<path>.lora_A.weight`` then ``unet.`` prefix by save_lora_weights [3P diffusers 0.29.2]
  convert_all_state_dict_to_peft / convert_state_dict_to_kohya:
      "unet" -> "lora_unet", lora_A -> lora_down, lora_B -> lora_up, every "." but the last two -> "_",
      plus ``<module>.alpha = tensor(len(lora_down))`` (= the rank)                            [3P diffusers 0.29.2]
  the reference then strips the ``base_model_model_`` infix PEFT's wrapper leaves in the names (checkpoint.py:93-100).
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional

import torch

LORA_TARGETS = ["to_k", "to_q", "to_v", "to_out.0", "conv2"]            # trainer/optimizer.py:84


def remove_delimiter_characters(name: str, max_length: int = 255) -> str:
    """checkpoint.py:58-81."""
    cleaned = re.sub(r"[^\w.-]+", "_", name)
    cleaned = re.sub(r"_+", "_", cleaned)
    cleaned = cleaned.strip("_.").lstrip(".")[:max_length]
    if not cleaned:
        raise ValueError("Malformed name")
    return cleaned


def peft_lora_state_dict(store) -> Dict[str, torch.Tensor]:
    """What ``get_peft_model_state_dict(unet)`` returns for the PEFT-wrapped UNet: adapter name stripped, the
    ``base_model.model.`` wrapper prefix kept; conv factors in conv layout ([r, Cin, 3, 3] / [Cout, r, 1, 1])."""
    out = {}
    for key, t in store.export_peft().items():
        out["base_model.model." + key.replace(".default.", ".")] = t.detach().to("cpu").contiguous()
    return out


def convert_state_dict_to_kohya(unet_peft_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """save_lora_weights (``unet.`` prefix) -> convert_all_state_dict_to_peft -> convert_state_dict_to_kohya -> the
    reference's ``base_model_model_`` strip (checkpoint.py:84-102)."""
    kohya: Dict[str, torch.Tensor] = {}
    for key, weight in unet_peft_sd.items():
        k = ("unet." + key).replace("unet", "lora_unet", 1)
        k = k.replace("lora_A", "lora_down").replace("lora_B", "lora_up")
        k = k.replace(".", "_", k.count(".") - 2)
        k = k.replace("base_model_model_", "")
        kohya[k] = weight
        if "lora_down" in k:
            kohya[f'{k.split(".")[0]}.alpha'] = torch.tensor(len(weight))
    return kohya


def kohya_to_peft_keys(kohya_sd: Dict[str, torch.Tensor], slot_names: List[str]) -> Dict[str, torch.Tensor]:
    """Inverse mapping for our own files: kohya module name -> LoraStore slot name (dots are not recoverable from the
    underscore form in general, so the store's slot names provide the dictionary)."""
    by_kohya = {"lora_unet_" + n.replace(".", "_"): n for n in slot_names}
    out = {}
    for k, t in kohya_sd.items():
        mod, _, rest = k.partition(".")
        if rest == "alpha":
            continue
        name = by_kohya[mod]
        which = "lora_A" if rest.startswith("lora_down") else "lora_B"
        out[f"{name}.{which}.default.weight"] = t
    return out


def adapter_config(rank: int, lora_alpha_multiplier: float, use_dora: bool = False) -> dict:
    """The ``adapter_config.json`` PEFT 0.10.0 writes for ``LoraConfig(r, lora_alpha=r*mult, init_lora_weights="gaussian",
    target_modules=[...], use_dora=...)`` (trainer/optimizer.py:86-92) on a diffusers UNet [3P, restated]."""
    return {
        "alpha_pattern": {}, "auto_mapping": {"base_model_class": "UNet2DConditionModel",
                                              "parent_library": "diffusers.models.unets.unet_2d_condition"},
        "base_model_name_or_path": None, "bias": "none", "fan_in_fan_out": False, "inference_mode": True,
        "init_lora_weights": "gaussian", "layer_replication": None, "layers_pattern": None, "layers_to_transform": None,
        "loftq_config": {}, "lora_alpha": rank * lora_alpha_multiplier, "lora_dropout": 0.0, "megatron_config": None,
        "megatron_core": "megatron.core", "modules_to_save": None, "peft_type": "LORA", "r": rank, "rank_pattern": {},
        "revision": None, "target_modules": list(LORA_TARGETS), "task_type": None, "use_dora": use_dora, "use_rslora": False,
    }


def unet_config_json(arch) -> dict:
    """The architecture fields of diffusers' ``unet/config.json`` that the executor's graph description carries."""
    attn = ["CrossAttnDownBlock2D" if a else "DownBlock2D" for a in arch.down_has_attn]
    up = ["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(arch.down_has_attn)]
    return {"_class_name": "UNet2DConditionModel", "in_channels": arch.in_channels, "out_channels": arch.out_channels,
            "block_out_channels": list(arch.block_out_channels), "down_block_types": attn, "up_block_types": up,
            "layers_per_block": arch.layers_per_block,
            "transformer_layers_per_block": list(arch.transformer_layers_per_block),
            "attention_head_dim": list(arch.num_attention_heads), "cross_attention_dim": arch.cross_attention_dim,
            "use_linear_projection": arch.use_linear_projection, "addition_embed_type": arch.addition_embed_type,
            "addition_time_embed_dim": arch.addition_time_embed_dim if arch.addition_embed_type else None,
            "projection_class_embeddings_input_dim": arch.projection_class_embeddings_input_dim if arch.addition_embed_type else None,
            "norm_num_groups": arch.norm_num_groups}


def save_checkpoint(output_dir: str, global_step: int, unet, embedding_handler, token_dict: dict, is_lora: bool,
                    unet_lora_parameters, pretrained_model_version: str, name: Optional[str] = None,
                    text_encoder_peft_models: Optional[list] = None, lora_alpha_multiplier: float = 1.0) -> None:
    """checkpoint.py:104-221.  ``unet`` is the B200 UNet executor (``unet.store`` holds the LoRA factors)."""
    from safetensors.torch import save_file
    if pretrained_model_version not in ("sdxl", "sd15"):
        raise ValueError(f"Invalid pretrained_model_version: {pretrained_model_version}. Expected one of: 'sdxl' or 'sd15'")
    if text_encoder_peft_models and any(m is not None for m in text_encoder_peft_models):
        raise NotImplementedError("text-encoder LoRA is outside the B200 training step (SURVEY.md 2, row 3)")
    print(f"Saving checkpoint at step.. {global_step}")
    name = remove_delimiter_characters(name)
    os.makedirs(output_dir, exist_ok=True)
    if embedding_handler is not None:
        embedding_handler.save_embeddings(os.path.join(output_dir, f"{name}_{pretrained_model_version}_embeddings.safetensors"))
    with open(os.path.join(output_dir, "special_params.json"), "w") as f:
        json.dump(token_dict, f)
    if not is_lora:
        # checkpoint.py:211-213: ``unet.save_pretrained(output_dir)`` of a plain diffusers UNet = config.json +
        # diffusion_pytorch_model.safetensors under diffusers parameter names
        if getattr(unet, "dense", None) is None:
            raise NotImplementedError("is_lora=False needs an executor built for full fine-tuning (UNetB200(dense=True))")
        save_file(unet.dense.export(), os.path.join(output_dir, "diffusion_pytorch_model.safetensors"))
        with open(os.path.join(output_dir, "config.json"), "w") as f:
            json.dump(unet_config_json(unet.arch), f, indent=2, sort_keys=True)
        return
    assert len(unet_lora_parameters) > 0, "Expected len(unet_lora_parameters) to be greater than zero if is_lora is True"
    with open(os.path.join(output_dir, "adapter_config.json"), "w") as f:
        json.dump(adapter_config(unet.rank, lora_alpha_multiplier), f, indent=2, sort_keys=True)
    kohya = convert_state_dict_to_kohya(peft_lora_state_dict(unet.store))
    save_file({k: v.contiguous() for k, v in kohya.items()},
              os.path.join(output_dir, f"{name}_{pretrained_model_version}_lora.safetensors"))


def load_lora_weights(lora_path: str, unet, trainer=None) -> None:
    """Read a ``*_lora.safetensors`` written by save_checkpoint back into the executor's flat LoRA buffer (the training-side
    counterpart of the reference's inference-only load_checkpoint, checkpoint.py:223-297).  Pass the owning ``trainer``
    (TrainerB200) when training continues afterwards: its optimizer state (Adam moments, Prodigy's p0 / s / d) and
    conditioning cache refer to the parameters they were built on and are reset."""
    from safetensors.torch import load_file
    sd = kohya_to_peft_keys(load_file(lora_path), [s.name for s in unet.store.slots])
    missing = [s.name for s in unet.store.slots if f"{s.name}.lora_A.default.weight" not in sd]
    if missing:
        raise KeyError(f"LoRA file lacks {len(missing)} modules, e.g. {missing[:3]}")
    unet.store.load_peft(sd)
    if trainer is not None:
        trainer.reset_optimizer_state()
