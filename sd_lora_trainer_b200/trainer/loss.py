"""Mirror of the reference's trainer/loss.py for the hot path.

compute_diffusion_loss / compute_snr dispatch into the fused sm_100a kernels; the token-attention regulariser
(loss.py:10-80) and the heat-map stacking (ti_cross_attn_loss.py:239-268) operate on [layers, B, 32, 32, 77]-sized
maps; the bicubic map resize is an sm_100a kernel pair (ops.bicubic_fwd/bwd), the remaining reductions are stock
torch ops on the device.
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


class _BicubicResize(torch.autograd.Function):
    """F.interpolate(hm.permute(0, 3, 1, 2), size, mode="bicubic").permute(0, 2, 3, 1) on channels-last maps, as one
    sm_100a kernel each way (ti_cross_attn_loss.py:262-266); the stock ATen bicubic kernels cost 26 ms per SDXL step."""

    @staticmethod
    def forward(ctx, hm: torch.Tensor, size):
        ctx.in_hw = (hm.shape[1], hm.shape[2])
        return ops.bicubic_fwd(hm, size[0], size[1])

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        return ops.bicubic_bwd(dy, *ctx.in_hw), None


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
            reshaped[i] = _BicubicResize.apply(hm, min_shape)
    return torch.stack(reshaped, dim=0)


def token_index_tensors(token_indices: List[List[int]], train_ids: List[int], device="cpu"):
    """Host-side part of loss.py:32-43: caption lengths and the positions of the trainable tokens
    (``token_indices.index(token_id)``; -1 marks a caption that lacks one, which the reference skips)."""
    tok_len = torch.tensor([len(t) for t in token_indices], dtype=torch.long)
    pos = torch.full((len(token_indices), max(len(train_ids), 1)), -1, dtype=torch.long)
    for b, tok in enumerate(token_indices):
        try:
            pos[b, :len(train_ids)] = torch.tensor([tok.index(t) for t in train_ids], dtype=torch.long)
        except ValueError:
            pos[b] = -1
    return tok_len.to(device), pos.to(device)


def token_attention_loss_tensors(scores: Sequence[torch.Tensor], masks: torch.Tensor, tok_len: torch.Tensor,
                                 ti_pos: torch.Tensor) -> torch.Tensor:
    """trainer/loss.py:10-80 with the per-caption Python loop replaced by index tensors (same arithmetic, same
    dtypes), so the whole regulariser is shape-static and can sit inside a CUDA graph."""
    img_ratio = masks.shape[-1] / masks.shape[-2]
    maps = process_and_stack_attention_scores(scores, img_ratio)              # [layers, B, h, w, 77]
    return token_attention_loss_from_maps(maps, masks, tok_len, ti_pos)


def token_attention_loss_from_maps(maps: torch.Tensor, masks: torch.Tensor, tok_len: torch.Tensor,
                                   ti_pos: torch.Tensor) -> torch.Tensor:
    """The regulariser on the stacked maps [layers, B, h, w, 77] (everything after ti_cross_attn_loss.py:239-268).  It sees
    the maps only through their mean over layers (and over pixels), so its gradient is the same map for every layer."""
    masks = masks[:, 0].float()
    n_layers, B, h, w, n_text = maps.shape
    masks = F.interpolate(masks.unsqueeze(1), size=(h, w)).squeeze(1)         # [B, h, w] (nearest)
    # (1) mean attention of every real caption token (positions 1 .. len-2)
    per_tok = maps.mean(dim=[0, 2, 3])                                        # [B, 77], maps' dtype
    sq = torch.relu(per_tok - 0.0) ** 2
    t = torch.arange(n_text, device=maps.device)[None, :]
    sel = ((t >= 1) & (t < (tok_len[:, None] - 1))).float()
    att_L2 = ((sq.float() * sel).sum(-1) / sel.sum(-1)).to(maps.dtype)        # [B]
    reg0 = 5.0 * att_L2.mean()
    # (2) heat-maps of the trainable tokens
    valid = (ti_pos >= 0).all(dim=1)                                          # captions holding every TI token
    nv = valid.float().sum()
    layer_mean = maps.mean(dim=0)                                             # [B, h, w, 77]
    idx = ti_pos.clamp(min=0)[:, None, None, :].expand(B, h, w, ti_pos.shape[1])
    hm = torch.gather(layer_mean, 3, idx).permute(0, 3, 1, 2).float()         # [B, n_tok, h, w]
    mk = masks[:, None]                                                       # [B, 1, h, w]
    vw = valid.float()[:, None, None, None]
    nv_safe = nv.clamp(min=1.0)                                               # nv == 0: every masked sum is 0, the
    denom = nv_safe * hm.shape[1] * h * w                                     # result is discarded below, grads stay 0
    reg1 = 1.0 * ((torch.relu(hm * mk) ** 2) * vw).sum() / denom
    reg2 = 2.0 * ((torch.relu(hm * (1 - mk) + 10) ** 2) * vw).sum() / denom
    reg3 = 1.0 * (hm.mean(dim=[2, 3]).var(dim=1) * valid.float()).sum() / nv_safe
    total = (reg0 + reg1 + reg2 + reg3).to(masks.dtype)
    return torch.where(nv > 0, total, torch.zeros_like(total))                # loss.py:55-56


def compute_token_attention_loss(scores: Sequence[torch.Tensor], masks: torch.Tensor,
                                 token_indices: List[List[int]], train_ids: List[int]) -> torch.Tensor:
    """trainer/loss.py:10-80; ``token_indices[b]`` is ``pipe.tokenizer.encode(captions[b])``."""
    tok_len, ti_pos = token_index_tensors(token_indices, train_ids, device=masks.device)
    return token_attention_loss_tensors(scores, masks, tok_len, ti_pos)


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
