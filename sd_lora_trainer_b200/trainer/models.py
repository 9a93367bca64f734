"""Mirror of the reference's trainer/models.py:7-54 for the hot path.

``load_models(pretrained_model, device, weight_dtype)`` returns the same 2-tuple
``((pipe, tokenizer_one, tokenizer_two, noise_scheduler, text_encoder_one, text_encoder_two, vae, unet), family)``.
diffusers is absent in this image, so the checkpoint is read as a plain diffusers-layout UNet state dict
(``pretrained_model["path"]`` -> .safetensors) or random-initialised (``pretrained_model["random_init"]``); VAE and
tokenizers are outside the accelerated path and are returned as None."""
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
    else:
        from safetensors.torch import load_file
        sd = load_file(pretrained_model["path"], device=str(device))
        sd = {k[len("unet."):] if k.startswith("unet.") else k: v for k, v in sd.items()}
        family = "sdxl" if "add_embedding.linear_1.weight" in sd else "sd15"     # models.py:15-28 try/except
    pipe = B200Pipe(family, sd, text_encoders, device)
    sd_model_version = "sdxl" if family.endswith("sdxl") else "sd15"
    return (pipe, None, None, pipe.scheduler, pipe.text_encoder, pipe.text_encoder_2, None, pipe.unet), sd_model_version
