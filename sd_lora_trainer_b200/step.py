"""One LoRA + textual-inversion training step on a B200 (main.py:263-382 of the reference), re-plumbed:

    prologue kernel (offset noise + add_noise) -> UNetB200.forward -> fused loss + dPred kernel
    -> token-attention regulariser on the captured scores -> UNetB200.backward (dA/dB into one flat fp32 buffer,
       d(prompt_embeds), d(pooled)) -> CLIP backward (stock torch) into the n_tokens embedding rows
    -> [data parallel: ONE all-reduce of the flat gradient buffer] -> ONE fused AdamW over LoRA factors + TI rows.

Every random draw is an INPUT (latent sample, noise, offset noise, timesteps, token ids) so a step can be replayed
against the oracle from identical state.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .arch import UNetArch, by_name
from .text import TIEmbedding, add_time_ids, encode_prompt, init_ti_rows, install_ti_rows
from .trainer.loss import (DistributionLoss, alphas_cumprod_table, compute_diffusion_loss,
                           compute_token_attention_loss)
from .unet import UNetB200

BF16 = torch.bfloat16


@dataclass
class StepConfig:
    """The TrainingConfig fields the step reads, with the reference's defaults (trainer/config.py:38-119)."""
    family: str = "sdxl"               # "sdxl" | "sd15"
    tiny: bool = False
    resolution: int = 1024
    lora_rank: int = 16
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
    seed: int = 0

    def arch(self) -> UNetArch:
        name = self.family if not self.tiny else f"tiny_{self.family}"
        return by_name(name)


def lr_schedule(cfg: StepConfig, global_step: int, completion_f: float):
    """main.py:237-240 (cold-start base lr), 268-291 (TI decay + freeze, exponential UNet warm-up)."""
    ti_lr = None
    if not cfg.disable_ti:
        ti_lr = cfg.ti_lr * (1 - completion_f) ** 1.7
        if completion_f > cfg.freeze_ti_after_completion_f:
            ti_lr = 0.0
    base = 2.0e-4 if cfg.disable_ti else 5.0e-5
    warm = cfg.unet_lr_warmup_steps if cfg.unet_lr_warmup_steps is not None else cfg.max_train_steps
    return ti_lr, base * (cfg.unet_lr / base) ** (global_step / warm)


class TrainerB200:
    """Owns the UNet executor, the flat parameter/gradient/moment buffers and the text encoders."""

    def __init__(self, cfg: StepConfig, unet_state_dict: Dict[str, torch.Tensor], text_encoders: Sequence,
                 device="cuda:0", ti_init: Optional[List[torch.Tensor]] = None, process_group=None):
        self.cfg, self.device = cfg, torch.device(device)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.sdxl = cfg.family == "sdxl"
        self.text_encoders = [te.to(self.device, BF16) if te is not None else None for te in text_encoders]
        for te in self.text_encoders:
            if te is not None:
                te.requires_grad_(False)
        dims = [te.text_model.embeddings.token_embedding.weight.shape[1] for te in self.text_encoders if te is not None]
        ntok = 0 if cfg.disable_ti else cfg.n_tokens
        ti_elems = sum(ntok * d for d in dims)
        self.unet = UNetB200(cfg.arch(), unet_state_dict, cfg.lora_rank, cfg.lora_alpha_multiplier, self.device,
                             ti_elems=ti_elems, lora_seed=cfg.seed + 2)
        self.store = self.unet.store
        self.acp = alphas_cumprod_table(device=self.device)
        self.train_ids: List[int] = []
        self.ti_rows: List[torch.Tensor] = []
        self.std_regs: List[DistributionLoss] = []
        if ntok:
            g = torch.Generator().manual_seed(cfg.seed)
            off = self.store.n_lora
            for i, te in enumerate(t for t in self.text_encoders if t is not None):
                table = te.text_model.embeddings.token_embedding.weight.data
                vocab, dim = table.shape
                rows = self.store.params[off:off + ntok * dim].view(ntok, dim)
                rows.copy_(ti_init[i].to(self.device, BF16) if ti_init is not None else init_ti_rows(table, ntok, g))
                self.std_regs.append(DistributionLoss(torch.cat([table, rows], dim=0)))      # loss.py:179-194
                rows.requires_grad_(True)
                install_ti_rows(te, rows)
                self.ti_rows.append(rows)
                self.train_ids = list(range(vocab, vocab + ntok))
                off += ntok * dim
            self.unet.set_capture(True)                        # init_daam_loss, main.py:50-52
        self.global_step = 0
        self.opt_step = 0
        self._accum = 0
        self.last_lrs = (None, None)

    # ---- one micro-step: forward, losses, backward ------------------------------------------------
    def forward_backward(self, inputs: Dict[str, torch.Tensor], ti_active: bool = True) -> Dict[str, torch.Tensor]:
        cfg, dev = self.cfg, self.device
        ga = cfg.gradient_accumulation_steps * self.world       # DP-N == accumulation-N
        latent = inputs["vae_latent"].to(dev, torch.float32).contiguous()
        noise = inputs["noise"].to(dev, BF16).contiguous().clone()
        mask = inputs["mask"].to(dev, torch.float32).contiguous()
        timesteps = inputs["timesteps"].to(dev).long().contiguous()
        B, Cc, H, W = latent.shape
        offset = inputs["offset_noise"].to(dev, torch.float32).reshape(B, Cc).contiguous() if cfg.noise_offset > 0 else None
        token_ids = [t.to(dev) for t in inputs["token_ids"]]
        need_text_grad = bool(self.ti_rows)
        with torch.set_grad_enabled(need_text_grad):
            prompt_embeds, pooled = encode_prompt(self.sdxl, self.text_encoders, token_ids)
        time_ids = add_time_ids(B, cfg.resolution, BF16, dev) if self.sdxl else None

        noisy, noisy8 = ops.noise_prologue(latent, noise, offset, cfg.noise_offset, self.acp, timesteps)
        pred8, scores = self.unet.forward(noisy8, B, H, W, timesteps, prompt_embeds.detach(),
                                          None if pooled is None else pooled.detach(), time_ids)
        out = {"noisy_latent": noisy, "model_pred8": pred8, "noise": noise}
        img_loss, dpred8 = compute_diffusion_loss(cfg.snr_gamma, pred8, noise, mask, self.acp, timesteps,
                                                  loss_scale=1.0 / ga)
        out["img_loss"] = img_loss
        total = img_loss.clone()
        dscores = None
        if not cfg.disable_ti:
            leaves = [s.detach().requires_grad_(True) for s in scores]
            tal = compute_token_attention_loss(leaves, mask, inputs["token_indices"], self.train_ids)
            (cfg.token_attention_loss_w * tal / ga).backward()
            dscores = [l.grad if l.grad is not None else torch.zeros_like(l) for l in leaves]
            out["token_attention_loss"] = tal.detach()
            out["attention_scores"] = scores
            total = total + cfg.token_attention_loss_w * tal.detach().float()
        if cfg.l1_penalty > 0.0:
            l1 = torch.zeros(1, dtype=torch.float32, device=dev)
            ops.abs_sum(self.store.params[:self.store.n_lora], l1)
            total = total + cfg.l1_penalty * l1 / self.store.numel_logical
        d_ctx, d_text = self.unet.backward(dpred8, dscores)
        out["d_prompt_embeds"] = d_ctx
        if need_text_grad:
            roots, grads = [prompt_embeds], [d_ctx.to(prompt_embeds.dtype)]
            if pooled is not None and d_text is not None:
                roots.append(pooled)
                grads.append(d_text.to(pooled.dtype))
            if ti_active:                                     # token-std regulariser, loss.py:222-231
                std = torch.stack([reg.compute_std_loss(rows) for reg, rows in zip(self.std_regs, self.ti_rows)]).mean()
                out["token_std_loss"] = std.detach()
                total = total + 0.01 * std.detach().float()
                roots.append(0.01 * std / ga)
                grads.append(torch.ones_like(std))
            torch.autograd.backward(roots, grads)
            off = self.store.n_lora
            for rows in self.ti_rows:                         # move the 3-row gradients into the flat fp32 buffer
                n = rows.numel()
                self.store.grads[off:off + n] += rows.grad.float().flatten()
                rows.grad = None
                off += n
        out["tot_loss"] = total
        return out

    # ---- optimizer: ONE kernel over LoRA factors + TI rows -------------------------------------------
    def optimizer_step(self):
        cfg = self.cfg
        ti_lr, unet_lr = self.last_lrs
        if self.pg is not None and self.world > 1:
            torch.distributed.all_reduce(self.store.grads, group=self.pg)     # the step's only collective
        self.opt_step += 1
        l1c = 0.0
        if cfg.l1_penalty > 0.0:
            # what autograd hands to every LoRA element for  loss += l1_penalty * sum|p| / numel  in bf16
            c = torch.tensor(cfg.l1_penalty, dtype=BF16) / self.store.numel_logical
            l1c = float(c)
        ops.adamw(self.store.params, self.store.grads, self.store.m, self.store.v, self.store.n_lora,
                  lr=unet_lr, wd=cfg.lora_weight_decay, l1_coeff=l1c, lr2=ti_lr or 0.0, wd2=cfg.ti_weight_decay,
                  step=self.opt_step, grad_scale=1.0, zero_grad=True)

    def step(self, inputs, completion_f: float = 0.0, do_optimizer: bool = True):
        ti_lr, unet_lr = lr_schedule(self.cfg, self.global_step, completion_f)
        self.last_lrs = (ti_lr, unet_lr)
        out = self.forward_backward(inputs, ti_active=bool(ti_lr and ti_lr > 0.0))
        self._accum += 1
        if do_optimizer and self._accum % self.cfg.gradient_accumulation_steps == 0:
            self.optimizer_step()
        self.global_step += 1
        return out
