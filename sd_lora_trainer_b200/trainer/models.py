"""Mirror of the reference's trainer/models.py:7-54 for the hot path.

``load_models(pretrained_model, device, weight_dtype)`` returns the same 2-tuple
``((pipe, tokenizer_one, tokenizer_two, noise_scheduler, text_encoder_one, text_encoder_two, vae, unet), family)``.
diffusers is absent in this image, so ``pretrained_model["path"]`` (.safetensors) is read directly: a single-file
checkpoint in the original LDM layout - what the reference's ``from_single_file`` takes - goes through the restated key
conversion of trainer/single_file.py (UNet, both CLIP text encoders, VAE encoder); a diffusers-layout UNet state dict is
taken as is; ``pretrained_model["random_init"]`` builds random weights.  The text encoders are materialised as transformers
CLIP modules from the converted weights; the VAE-encoder weights are handed to vae.VAEEncoderB200 by the caller
(``pipe.vae_state``); tokenizers need a vocabulary file (``pretrained_model["tokenizer_dir"]``, optional)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict

import torch

from .. import arch as _arch
from ..init import random_state_dict
from .loss import alphas_cumprod_table


class NoiseScheduler:
    """The DDPMScheduler surface main.py / loss.py touch: .config, .alphas_cumprod, .add_noise (main.py:321-326)."""

    def __init__(self, device="cpu"):
        self.config = SimpleNamespace(num_train_timesteps=1000, prediction_type="epsilon")
        self.alphas_cumprod = alphas_cumprod_table(device=device)

    def add_noise(self, original_samples, noise, timesteps):
        from .. import ops
        acp = self.alphas_cumprod.to(original_samples.device)
        nz = noise.to(torch.bfloat16).contiguous().clone()
        noisy, _ = ops.noise_prologue(original_samples.float().contiguous(), nz, None, 0.0, acp, timesteps.long())
        return noisy


class B200Pipe:
    """What the reference reads off its diffusers pipeline: .unet .vae .text_encoder(_2) .tokenizer(_2) .scheduler."""

    def __init__(self, family: str, unet_state: Dict[str, torch.Tensor], text_encoders, device):
        self.family, self.unet_state, self.device = family, unet_state, device
        self.text_encoder, self.text_encoder_2 = text_encoders
        self.tokenizer = self.tokenizer_2 = self.vae = None
        self.scheduler = NoiseScheduler(device)
        self.unet = None            # becomes the B200 executor once LoRA is injected (optimizer.get_unet_lora_parameters)


def load_models(pretrained_model: dict, device, weight_dtype=torch.bfloat16, text_encoders=(None, None)):
    if weight_dtype != torch.bfloat16:
        raise ValueError("the B200 step computes in bf16 (trainer/config.py:99 weight_type default); "
                         f"got {weight_dtype}")
    if "random_init" in pretrained_model:
        family = pretrained_model["random_init"]
        sd = random_state_dict(_arch.by_name(family), seed=pretrained_model.get("seed", 0), device=device)
        vae_state = None
    else:
        from safetensors.torch import load_file
        from .single_file import is_single_file, split_single_file
        sd = load_file(pretrained_model["path"], device=str(device))
        vae_state = None
        if is_single_file(sd.keys()):
            family, parts = split_single_file(sd)                                # models.py:15-28: SDXL first, else SD1.5
            sd, vae_state = parts["unet"], parts["vae"]
            if text_encoders[0] is None:
                text_encoders = _build_text_encoders(family, parts, device)
        else:
            sd = {k[len("unet."):] if k.startswith("unet.") else k: v for k, v in sd.items()}
            family = "sdxl" if "add_embedding.linear_1.weight" in sd else "sd15"
    pipe = B200Pipe(family, sd, text_encoders, device)
    pipe.vae_state = vae_state
    tok_dir = pretrained_model.get("tokenizer_dir")
    if tok_dir:                                                                  # vocab.json + merges.txt of the CLIP BPE tokenizer
        from transformers import CLIPTokenizer
        pipe.tokenizer = CLIPTokenizer.from_pretrained(tok_dir)
        pipe.tokenizer_2 = CLIPTokenizer.from_pretrained(tok_dir, pad_token="!") if family.endswith("sdxl") else None
    sd_model_version = "sdxl" if family.endswith("sdxl") else "sd15"
    return (pipe, pipe.tokenizer, pipe.tokenizer_2, pipe.scheduler, pipe.text_encoder, pipe.text_encoder_2, pipe.vae_state,
            pipe.unet), sd_model_version


def _build_text_encoders(family: str, parts: Dict[str, Dict[str, torch.Tensor]], device):
    """transformers CLIP text modules with the checkpoint's converted weights (shapes give the configuration)."""
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection

    def cfg_of(sd, act, proj):
        emb = sd["text_model.embeddings.token_embedding.weight"]
        layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("text_model.encoder.layers."))
        hidden = emb.shape[1]
        return CLIPTextConfig(vocab_size=emb.shape[0], hidden_size=hidden, intermediate_size=4 * hidden, num_hidden_layers=layers,
                              num_attention_heads=hidden // 64, max_position_embeddings=77, hidden_act=act, projection_dim=proj,
                              bos_token_id=49406, eos_token_id=49407, pad_token_id=49407)

    out = []
    sd1 = parts["text_encoder"]
    with torch.device(device):
        te1 = CLIPTextModel(cfg_of(sd1, "quick_gelu", 768)).to(torch.bfloat16)
    te1.load_state_dict(sd1, strict=False)
    out.append(te1)
    if family == "sdxl":
        sd2 = parts["text_encoder_2"]
        with torch.device(device):
            te2 = CLIPTextModelWithProjection(cfg_of(sd2, "gelu", sd2["text_projection.weight"].shape[0])).to(torch.bfloat16)
        te2.load_state_dict(sd2, strict=False)
        out.append(te2)
    else:
        out.append(None)
    return tuple(out)
