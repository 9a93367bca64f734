"""Mirror of the reference's trainer/loss.py for the hot path.

compute_diffusion_loss / compute_snr dispatch into the fused sm_100a kernels; the token-attention regulariser
(loss.py:10-80) and the heat-map stacking (ti_cross_attn_loss.py:239-268) operate on [layers, B, 32, 32, 77]-sized
maps and run as stock torch ops on the device for now (next to be fused).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from .. import ops


def alphas_cumprod_table(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                         device="cpu") -> torch.Tensor:
    """DDPMScheduler(scaled_linear) table, fp32 (models.py:32 -> [3P] DDPMScheduler.from_config)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(device)


def compute_snr(alphas_cumprod: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """trainer/loss.py:83-106 (used by the mirror API; the fused loss uses b200_snr_weights)."""
    sa = (alphas_cumprod ** 0.5)[timesteps].float()
    so = ((1.0 - alphas_cumprod) ** 0.5)[timesteps].float()
    return (sa / so) ** 2


def compute_diffusion_loss(snr_gamma: Optional[float], pred8: torch.Tensor, noise: torch.Tensor, mask: torch.Tensor,
                           alphas_cumprod: torch.Tensor, timesteps: torch.Tensor, loss_scale: float = 1.0,
                           want_grad: bool = True):
    """trainer/loss.py:127-170 for epsilon prediction.  pred8: NHWC [B*HW, 8]; noise NCHW bf16; mask NCHW fp32.
    Returns (loss [1] fp32, dpred8 or None).  With snr_gamma the reference's final mask normalisation divides by
    exactly 1.0 (the ``mask.mean(dim=[])`` quirk); without it each sample is divided by its mask mean
    normalised over the batch."""
    B = noise.shape[0]
    if snr_gamma is None or snr_gamma == 0.0:
        mm = mask.mean(dim=[1, 2, 3])
        weights = (1.0 / (mm / mm.mean())).float().contiguous()
    else:
        weights = ops.snr_weights(alphas_cumprod, timesteps, float(snr_gamma))
    return ops.diffusion_loss(pred8, pred8.stride(0), noise, mask, weights, loss_scale, want_grad)


def process_and_stack_attention_scores(scores: Sequence[torch.Tensor], img_ratio: float) -> torch.Tensor:
    """trainer/ti_cross_attn_loss.py:239-268."""
    reshaped, min_px, min_shape = [], math.inf, None
    for score in scores:
        bs, seq_len, ch = score.shape
        width = round(math.sqrt(seq_len * img_ratio))
        height = round(width / img_ratio)
        reshaped.append(score.reshape(bs, height, width, ch))
        if height * width < min_px:
            min_px, min_shape = height * width, (height, width)
    for i, hm in enumerate(reshaped):
        if hm.shape[1] * hm.shape[2] != min_px:
            reshaped[i] = F.interpolate(hm.permute(0, 3, 1, 2), size=min_shape, mode="bicubic").permute(0, 2, 3, 1)
    return torch.stack(reshaped, dim=0)


def compute_token_attention_loss(scores: Sequence[torch.Tensor], masks: torch.Tensor,
                                 token_indices: List[List[int]], train_ids: List[int]) -> torch.Tensor:
    """trainer/loss.py:10-80; ``token_indices[b]`` is ``pipe.tokenizer.encode(captions[b])``."""
    masks = masks[:, 0].float()
    img_ratio = masks.shape[-1] / masks.shape[-2]
    att_L2_losses, ti_heatmaps, ti_masks = [], [], []
    attention_maps = process_and_stack_attention_scores(scores, img_ratio)
    n_layers, batch_size, w, h, n_tokens = attention_maps.shape
    masks = F.interpolate(masks.unsqueeze(1), size=(attention_maps.shape[-3], attention_maps.shape[-2])).squeeze(1)
    masks = masks.unsqueeze(0).unsqueeze(-1).repeat(n_layers, 1, 1, 1, n_tokens)
    for b, tok in enumerate(token_indices):
        mean_att = attention_maps[:, b, :, :, 1:len(tok) - 1].mean(dim=[0, 1, 2])
        att_L2_losses.append((torch.relu(mean_att - 0.0) ** 2).mean())
        try:
            ti_idx = [tok.index(t) for t in train_ids]
        except ValueError:
            continue
        ti_heatmaps.append(torch.stack([attention_maps[:, b, :, :, i].mean(dim=0).float() for i in ti_idx]))
        ti_masks.append(torch.stack([masks[:, b, :, :, i].mean(dim=0) for i in ti_idx]))
    if len(ti_heatmaps) == 0:
        return torch.tensor(0.0).to(masks.dtype)
    ti_heatmaps, ti_masks = torch.stack(ti_heatmaps), torch.stack(ti_masks)
    token_attention_scores = ti_heatmaps.mean(dim=[2, 3]).var(dim=1)
    reg0 = 5.0 * torch.stack(att_L2_losses).mean()
    reg1 = 1.0 * (torch.relu(ti_heatmaps * ti_masks) ** 2).mean()
    reg2 = 2.0 * (torch.relu(ti_heatmaps * (1 - ti_masks) + 10) ** 2).mean()
    reg3 = 1.0 * token_attention_scores.mean()
    return (reg0 + reg1 + reg2 + reg3).to(masks.dtype)


class DistributionLoss:
    """trainer/loss.py:254-297, std statistics only (covariance weight is 0.0 by default, config.py:75)."""

    def __init__(self, pretrained_embeddings: torch.Tensor):
        self.target_stds = pretrained_embeddings.std(-1)
        self.target_stds_mean = self.target_stds.mean()
        self.target_stds_var = self.target_stds.std() ** 2 / self.target_stds.mean()

    def compute_std_loss(self, new_embeddings: torch.Tensor) -> torch.Tensor:
        if new_embeddings.size(1) == 1:
            new_embeddings = new_embeddings.unsqueeze(0)
        return ((self.target_stds_mean - new_embeddings.std(-1)) ** 2 / self.target_stds_var).mean()
