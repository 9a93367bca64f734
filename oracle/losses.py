"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Behaviour-exact restatement of the reference-owned loss code:
  * DDPMScheduler.add_noise / alphas_cumprod      diffusers 0.29.2 [3P], SURVEY Appendix B
  * compute_snr                                   trainer/loss.py:83-106
  * compute_diffusion_loss                        trainer/loss.py:127-170 (incl. the
        ``mask.mean(dim=[])`` full-reduction quirk at :165)
  * process_and_stack_attention_scores            trainer/ti_cross_attn_loss.py:239-268
  * compute_token_attention_loss                  trainer/loss.py:10-80
  * DistributionLoss.compute_std_loss             trainer/loss.py:258-297
The tokenizer is absent in this image, so ``token_indices`` (the result of
``pipe.tokenizer.encode(caption)``, loss.py:33) is passed in explicitly.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F


class DDPMSchedulerOracle:
    """scaled_linear betas 0.00085 -> 0.012, 1000 steps, epsilon prediction (Appendix A)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, prediction_type: str = "epsilon"):
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device).to(dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = acp[timesteps] ** 0.5
        sa = sa.flatten()
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
        so = (1 - acp[timesteps]) ** 0.5
        so = so.flatten()
        while so.dim() < original_samples.dim():
            so = so.unsqueeze(-1)
        return sa * original_samples + so * noise


def compute_snr(scheduler: DDPMSchedulerOracle, timesteps):
    acp = scheduler.alphas_cumprod
    sa = (acp ** 0.5).to(device=timesteps.device)[timesteps].float()
    so = ((1.0 - acp) ** 0.5).to(device=timesteps.device)[timesteps].float()
    return (sa / so) ** 2


def compute_diffusion_loss(snr_gamma: Optional[float], model_pred, noise, mask,
                           scheduler: DDPMSchedulerOracle, timesteps):
    assert scheduler.prediction_type == "epsilon"
    target = noise
    loss = (model_pred - target).pow(2) * mask
    if snr_gamma is None or snr_gamma == 0.0:
        mean_mask_values = mask.mean(dim=list(range(1, len(loss.shape))))
        mean_mask_values = mean_mask_values / mean_mask_values.mean()
        loss = loss.mean(dim=list(range(1, len(loss.shape)))) / mean_mask_values
        loss = loss.mean()
    else:
        snr = compute_snr(scheduler, timesteps)
        base_weight = torch.stack([snr, snr_gamma * torch.ones_like(timesteps)], dim=1).min(dim=1)[0] / snr
        mse_loss_weights = base_weight
        mse_loss_weights = mse_loss_weights / mse_loss_weights.mean()
        loss = loss.mean(dim=list(range(1, len(loss.shape)))) * mse_loss_weights
        # loss is 1-D here, so the dim list below is [] -> full reduction (the quirk)
        mean_mask_values = mask.mean(dim=list(range(1, len(loss.shape))))
        mean_mask_values = mean_mask_values / mean_mask_values.mean()
        loss = loss.mean(dim=list(range(1, len(loss.shape)))) / mean_mask_values
        loss = loss.mean()
    return loss


def process_and_stack_attention_scores(scores: Sequence[torch.Tensor], img_ratio: float):
    reshaped, min_px, min_shape = [], math.inf, None
    for score in scores:
        bs, seq_len, ch = score.shape
        width = round(math.sqrt(seq_len * img_ratio))
        height = round(width / img_ratio)
        r = score.reshape(bs, height, width, ch)
        reshaped.append(r)
        if height * width < min_px:
            min_px = height * width
            min_shape = (height, width)
    for i, hm in enumerate(reshaped):
        if hm.shape[1] * hm.shape[2] != min_px:
            hm = F.interpolate(hm.permute(0, 3, 1, 2), size=min_shape, mode="bicubic").permute(0, 2, 3, 1)
            reshaped[i] = hm
    return torch.stack(reshaped, dim=0)


def compute_token_attention_loss(scores: Sequence[torch.Tensor], masks, token_indices: List[List[int]],
                                 train_ids: List[int]):
    masks = masks[:, 0].float()
    img_ratio = masks.shape[-1] / masks.shape[-2]
    att_L2_losses, ti_heatmaps, ti_masks = [], [], []
    att_reg_threshold = 0.0
    attention_maps = process_and_stack_attention_scores(scores, img_ratio)
    n_layers, batch_size, w, h, n_tokens = attention_maps.shape
    masks = F.interpolate(masks.unsqueeze(1), size=(attention_maps.shape[-3], attention_maps.shape[-2])).squeeze(1)
    masks = masks.unsqueeze(0).unsqueeze(-1)
    masks = masks.repeat(n_layers, 1, 1, 1, n_tokens)
    for batch_index, tok in enumerate(token_indices):
        mean_att_per_token = attention_maps[:, batch_index, :, :, 1:len(tok) - 1].mean(dim=[0, 1, 2])
        att_L2_losses.append((torch.relu(mean_att_per_token - att_reg_threshold) ** 2).mean())
        try:
            ti_token_indices = [tok.index(t) for t in train_ids]
        except ValueError:
            continue
        hms, mks = [], []
        for ti in ti_token_indices:
            hms.append(attention_maps[:, batch_index, :, :, ti].mean(dim=0).float())
            mks.append(masks[:, batch_index, :, :, ti].mean(dim=0))
        ti_heatmaps.append(torch.stack(hms))
        ti_masks.append(torch.stack(mks))
    if len(ti_heatmaps) == 0:
        return torch.tensor(0.0).to(masks.dtype)
    ti_heatmaps = torch.stack(ti_heatmaps)
    ti_masks = torch.stack(ti_masks)
    token_means = ti_heatmaps.mean(dim=[2, 3])
    token_attention_scores = token_means.var(dim=1)
    reg_loss_0 = 5.0 * torch.stack(att_L2_losses).mean()
    reg_loss_1 = 1.0 * (torch.relu(ti_heatmaps * ti_masks) ** 2).mean()
    reg_loss_2 = 2.0 * (torch.relu(ti_heatmaps * (1 - ti_masks) + 10) ** 2).mean()
    reg_loss_3 = 1.0 * token_attention_scores.mean()
    return (reg_loss_0 + reg_loss_1 + reg_loss_2 + reg_loss_3).to(masks.dtype)


class DistributionLossOracle:
    """trainer/loss.py:254-297; only the std statistics (the covariance target is
    computed by the reference but its weight is 0.0 by default, config.py:75)."""

    def __init__(self, pretrained_embeddings: torch.Tensor):
        self.target_stds = pretrained_embeddings.std(-1)
        self.target_stds_mean = self.target_stds.mean()
        self.target_stds_var = self.target_stds.std() ** 2 / self.target_stds.mean()

    def compute_std_loss(self, new_embeddings):
        if new_embeddings.size(1) == 1:
            new_embeddings = new_embeddings.unsqueeze(0)
        return ((self.target_stds_mean - new_embeddings.std(-1)) ** 2 / self.target_stds_var).mean()
