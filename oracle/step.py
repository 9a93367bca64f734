"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

One training step of the reference, restated from main.py:263-382 with every
random draw INJECTED (SURVEY §7 "RNG"): latent sample, noise, offset noise,
timesteps, token ids.  Setup mirrors main.py:92-176 (token init, freeze, LoRA
inject, AdamW optimizers; trainer/optimizer.py:6-39,107-155).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .unet import UNet2DConditionModel, UNetConfig, init_score_capture
from .lora import inject_lora
from .losses import (DDPMSchedulerOracle, DistributionLossOracle, compute_diffusion_loss,
                     compute_token_attention_loss)
from .text import add_time_ids, build_text_encoders, encode_prompt, initialize_new_tokens


@dataclass
class StepConfig:
    """The TrainingConfig fields the step reads (trainer/config.py:38-119 defaults)."""
    family: str = "sdxl"               # "sdxl" | "sd15"
    tiny: bool = False
    resolution: int = 1024
    lora_rank: int = 16
    is_lora: bool = True               # False: full-UNet fine-tune (main.py:143-148)
    lora_alpha_multiplier: float = 1.0
    lora_weight_decay: float = 0.004
    unet_lr: float = 0.0003
    ti_lr: float = 0.001
    ti_weight_decay: float = 0.0
    disable_ti: bool = False
    n_tokens: int = 3
    token_attention_loss_w: float = 3e-7
    l1_penalty: float = 0.03
    noise_offset: float = 0.02
    snr_gamma: Optional[float] = 5.0
    gradient_accumulation_steps: int = 1
    max_train_steps: int = 300
    unet_lr_warmup_steps: Optional[int] = None
    freeze_ti_after_completion_f: float = 0.7
    unet_optimizer_type: str = "adamw"
    ti_optimizer: str = "adamw"
    prodigy_d_coef: float = 1.0
    unet_prodigy_growth_factor: float = 1.05
    weight_dtype: torch.dtype = torch.bfloat16
    seed: int = 0

    def unet_config(self) -> UNetConfig:
        if self.tiny:
            return UNetConfig.tiny_sdxl() if self.family == "sdxl" else UNetConfig.tiny_sd15()
        return UNetConfig.sdxl() if self.family == "sdxl" else UNetConfig.sd15()


def lr_schedule(cfg: StepConfig, global_step: int, completion_f: float):
    """main.py:237-240, 268-291."""
    ti_lr = None
    if not cfg.disable_ti:
        ti_lr = cfg.ti_lr * (1 - completion_f) ** 1.7
        if completion_f > cfg.freeze_ti_after_completion_f:
            ti_lr = 0.0
    base = 2.0e-4 if cfg.disable_ti else 5.0e-5
    if not cfg.is_lora:
        base = 1.0e-5                                                     # main.py:239-240
    warm = cfg.unet_lr_warmup_steps if cfg.unet_lr_warmup_steps is not None else cfg.max_train_steps
    unet_lr = base * (cfg.unet_lr / base) ** (global_step / warm)
    return ti_lr, unet_lr


class OracleTrainer:
    def __init__(self, cfg: StepConfig, device="cpu"):
        self.cfg = cfg
        self.device = torch.device(device)
        dt = cfg.weight_dtype
        torch.manual_seed(cfg.seed)
        self.unet = UNet2DConditionModel(cfg.unet_config())
        self.scheduler = DDPMSchedulerOracle()
        self.text_encoders = build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1)
        self.unet.to(self.device, dt)
        for te in self.text_encoders:
            if te is not None:
                te.to(self.device, dt)
        self.hooked = init_score_capture(self.unet)                       # main.py:50-52
        self.train_ids: List[int] = []
        self.ti_params: List[torch.nn.Parameter] = []
        self.std_regs: Dict[str, DistributionLossOracle] = {}
        if not cfg.disable_ti:
            self.train_ids = initialize_new_tokens(self.text_encoders, cfg.n_tokens, seed=cfg.seed)
        for te in self.text_encoders:                                     # main.py:109-114
            if te is not None:
                te.requires_grad_(False)
        if not cfg.disable_ti:
            for i, te in enumerate(self.text_encoders):
                if te is None:
                    continue
                w = te.text_model.embeddings.token_embedding.weight
                self.std_regs[f"txt_encoder_{i}"] = DistributionLossOracle(w.data)   # loss.py:179-194
                w.requires_grad_(True)                                    # optimizer.py:116-121
                self.ti_params.append(w)
            if cfg.ti_optimizer == "prodigy":                             # optimizer.py:134-144
                from .prodigy import Prodigy
                self.opt_ti = Prodigy([{"params": self.ti_params, "lr": 1.0, "weight_decay": cfg.ti_weight_decay}],
                                      d_coef=1.0, lr=1.0, decouple=True, use_bias_correction=True, safeguard_warmup=True,
                                      weight_decay=cfg.ti_weight_decay, betas=(0.9, 0.99))
            else:
                self.opt_ti = torch.optim.AdamW(
                    [{"params": self.ti_params, "lr": cfg.ti_lr, "weight_decay": cfg.ti_weight_decay}],
                    weight_decay=cfg.ti_weight_decay)                     # optimizer.py:144-148
        else:
            self.opt_ti = None
        if cfg.is_lora:
            self.unet.requires_grad_(False)                               # main.py:108
            self.unet, self.lora_params = inject_lora(self.unet, cfg.lora_rank, cfg.lora_alpha_multiplier,
                                                      seed=cfg.seed + 2)
            self.unet.to(self.device)                                     # adapters were created on the host
            self.lora_params = [p for p in self.unet.parameters() if p.requires_grad]
            trainable = self.lora_params
        else:                                                             # main.py:143-148: full fine-tuning
            self.unet.requires_grad_(True)
            self.lora_params = []                                         # unet_lora_parameters = None: no L1 penalty
            trainable = list(self.unet.parameters())
        if cfg.unet_optimizer_type == "prodigy":                          # optimizer.py:22-34
            from .prodigy import Prodigy
            self.opt_unet = Prodigy([{"params": trainable, "weight_decay": cfg.lora_weight_decay}], d_coef=cfg.prodigy_d_coef,
                                    lr=1.0, decouple=True, use_bias_correction=True, safeguard_warmup=True,
                                    weight_decay=cfg.lora_weight_decay, betas=(0.9, 0.99),
                                    growth_rate=cfg.unet_prodigy_growth_factor)
        else:
            self.opt_unet = torch.optim.AdamW(
                [{"params": trainable, "weight_decay": cfg.lora_weight_decay}],
                lr=1e-4, weight_decay=cfg.lora_weight_decay)              # optimizer.py:16-17
        self.global_step = 0
        self._accum = 0

    # ---- one micro-step: main.py:263-363 -------------------------------------------------
    def forward_loss(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cfg, dt = self.cfg, self.cfg.weight_dtype
        sdxl = cfg.family == "sdxl"
        mask = inputs["mask"].to(self.device)
        prompt_embeds, pooled = encode_prompt(sdxl, self.text_encoders,
                                              [t.to(self.device) for t in inputs["token_ids"]])
        time_ids = add_time_ids(prompt_embeds.shape[0], cfg.resolution, prompt_embeds.dtype, self.device) if sdxl else None
        vae_latent = inputs["vae_latent"].to(self.device).to(dt)          # main.py:311
        noise = inputs["noise"].to(self.device).to(dt).clone()            # randn_like(vae_latent)
        if cfg.noise_offset > 0.0:
            noise += cfg.noise_offset * inputs["offset_noise"].to(self.device)   # fp32 added in place, main.py:314-317
        timesteps = inputs["timesteps"].to(self.device).long()
        noisy = self.scheduler.add_noise(vae_latent, noise, timesteps)
        pred = self.unet(noisy, timesteps, encoder_hidden_states=prompt_embeds, timestep_cond=None,
                         added_cond_kwargs={"text_embeds": pooled, "time_ids": time_ids}, return_dict=False)[0]
        out = {"model_pred": pred, "noisy_latent": noisy, "prompt_embeds": prompt_embeds}
        loss = compute_diffusion_loss(cfg.snr_gamma, pred, noise, mask, self.scheduler, timesteps)
        out["img_loss"] = loss.detach()
        if not cfg.disable_ti:
            scores = [m.cross_attention_scores for m in self.hooked]
            tal = compute_token_attention_loss(scores, mask, inputs["token_indices"], self.train_ids)
            out["token_attention_loss"] = tal.detach()
            out["attention_scores"] = [s.detach() for s in scores]
            loss = loss + cfg.token_attention_loss_w * tal
        if cfg.l1_penalty > 0.0 and self.lora_params:
            l1 = sum(p.abs().sum() for p in self.lora_params) / sum(p.numel() for p in self.lora_params)
            loss += cfg.l1_penalty * l1
        if self.opt_ti is not None and self.opt_ti.param_groups[0]["lr"] > 0.0:
            std_losses = []
            for key, reg in self.std_regs.items():
                i = int(key.rsplit("_", 1)[1])
                w = self.text_encoders[i].text_model.embeddings.token_embedding.weight
                rows = w[torch.tensor(self.train_ids, dtype=torch.long, device=w.device)]
                std_losses.append(reg.compute_std_loss(rows))
            mean_std = torch.stack(std_losses).mean()
            loss += 0.01 * mean_std                                       # loss.py:222-231
            out["token_std_loss"] = mean_std.detach()
        out["tot_loss"] = loss.detach()
        out["loss"] = loss
        return out

    def set_lrs(self, completion_f: float):
        ti_lr, unet_lr = lr_schedule(self.cfg, self.global_step, completion_f)
        if self.opt_ti is not None and self.cfg.ti_optimizer != "prodigy":        # main.py:269
            self.opt_ti.param_groups[0]["lr"] = ti_lr
        self.opt_unet.param_groups[0]["lr"] = unet_lr

    def step(self, inputs, completion_f: float = 0.0, do_optimizer: bool = True):
        """forward + backward (+ optimizer when the accumulation window closes)."""
        self.set_lrs(completion_f)
        out = self.forward_loss(inputs)
        (out["loss"] / self.cfg.gradient_accumulation_steps).backward()
        self._accum += 1
        if do_optimizer and self._accum % self.cfg.gradient_accumulation_steps == 0:
            self.optimizer_step()
        self.global_step += 1
        return out

    def optimizer_step(self):
        if self.opt_ti is not None:
            for w in self.ti_params:                                      # main.py:368-371
                w.grad.data[:-self.cfg.n_tokens, :] *= 0.0
            self.opt_ti.step()
        self.opt_unet.step()
        if self.opt_ti is not None:
            self.opt_ti.zero_grad()
        self.opt_unet.zero_grad()


def make_inputs(cfg: StepConfig, batch: int, seed: int = 1234, latent_hw: Optional[int] = None,
                face_mask: bool = False, vocab: Optional[int] = None, train_ids: Optional[List[int]] = None):
    """Synthetic step inputs in the shape SURVEY §8(d) fixes."""
    g = torch.Generator().manual_seed(seed)
    ucfg = cfg.unet_config()
    hw = latent_hw if latent_hw is not None else cfg.resolution // 8
    lat = torch.randn(batch, 4, hw, hw, generator=g) * ucfg.vae_scaling_factor
    noise = torch.randn(batch, 4, hw, hw, generator=g).to(cfg.weight_dtype)
    off = torch.randn(batch, 4, 1, 1, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    if face_mask:
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, hw), torch.linspace(-1, 1, hw), indexing="ij")
        blob = torch.exp(-(xx ** 2 + yy ** 2) / 0.35).clamp_min(0.05)
        blob = blob / blob.max()
        mask = blob[None, None].repeat(batch, 4, 1, 1)
    else:
        mask = torch.ones(batch, 4, hw, hw)
    if vocab is None:
        vocab = 128 if cfg.tiny else 49408
    bos, eos = (126, 127) if cfg.tiny else (49406, 49407)
    n_enc = 2 if cfg.family == "sdxl" else 1
    ntok = 0 if cfg.disable_ti else cfg.n_tokens
    ids, token_indices = [], []
    base = torch.full((batch, 77), eos, dtype=torch.long)
    for b in range(batch):
        n_words = int(torch.randint(4, 12, (1,), generator=g))
        words = torch.randint(0, bos, (n_words,), generator=g).tolist()
        tids = (train_ids or [vocab + i for i in range(ntok)])[:ntok]
        seq = [bos] + tids + words + [eos]
        base[b, :len(seq)] = torch.tensor(seq)
        token_indices.append(seq)
    for _ in range(n_enc):
        ids.append(base.clone())
    return {"vae_latent": lat, "noise": noise, "offset_noise": off, "timesteps": t, "mask": mask,
            "token_ids": ids, "token_indices": token_indices}
