"""Mirror of the per-step part of the reference's trainer/dataset.py: ``__getitem__`` draws the
training latent from the cached VAE posterior and scales it (dataset.py:181-193).  The one-time VAE encode that fills the
cache (dataset.py:141-179) is outside the step; this class takes its result - the posterior
``parameters`` tensor diffusers' ``DiagonalGaussianDistribution`` wraps ([1, 8, h, w]: mean | logvar) - as given."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .. import ops


def prepare_captions(captions: Sequence[Optional[str]], substitute_caption_map: Optional[dict] = None):
    """dataset.py:46-52: captions are lower-cased, every key of ``substitute_caption_map`` (``config.token_dict``:
    {"TOK": "<s0><s1><s2>"}) is replaced - key lower-cased too - by its value, missing captions become ''."""
    out = []
    for c in captions:
        if c is None or (isinstance(c, float) and c != c):
            out.append("")
            continue
        c = str(c).lower()
        for key, value in (substitute_caption_map or {}).items():
            c = c.replace(key.lower(), value)
        out.append(c)
    return out


class CachedLatentDataset:
    """``PreprocessedDataset`` with ``do_cache=True`` after its constructor ran: captions, posterior parameters, masks."""

    def __init__(self, captions: Sequence[str], posterior_parameters: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                 vae_scaling_factor: float):
        assert len(captions) == len(posterior_parameters) == len(masks)
        self.captions = list(captions)
        self.params = [p.to(torch.float32) for p in posterior_parameters]
        self.masks = list(masks)
        self.vae_scaling_factor = float(vae_scaling_factor)

    @classmethod
    def from_images(cls, vae_encoder, captions: Sequence[str], images: Sequence[torch.Tensor],
                    masks: Optional[Sequence[Optional[torch.Tensor]]], vae_scaling_factor: float,
                    substitute_caption_map: Optional[dict] = None) -> "CachedLatentDataset":
        """The constructor's caching pass (dataset.py:85-99 -> ``_process`` 141-179): every image ([3, H, W] in [-1, 1],
        what ``prepare_image`` returns) goes through the VAE encoder once and its posterior parameters are kept; a mask
        ([1, H, W] or [H, W] in [0, 1], what ``prepare_mask`` returns) is resized to the latent grid with nearest
        neighbour and repeated over the latent channels, no mask means all ones (dataset.py:160-175).
        ``vae_encoder`` is a ``sd_lora_trainer_b200.vae.VAEEncoderB200``."""
        params, out_masks = [], []
        for i, img in enumerate(images):
            p = vae_encoder.encode_moments(img[None] if img.dim() == 3 else img)
            c, h, w = p.shape[1] // 2, p.shape[2], p.shape[3]
            m = None if masks is None else masks[i]
            if m is None:
                m = torch.ones(c, h, w, dtype=torch.float32, device=p.device)
            else:
                m = m.to(p.device, torch.float32).reshape(1, 1, *m.shape[-2:])
                m = torch.nn.functional.interpolate(m, size=(h, w), mode="nearest").repeat(1, c, 1, 1).squeeze(0)
            params.append(p)
            out_masks.append(m)
        if substitute_caption_map is not None:
            captions = prepare_captions(captions, substitute_caption_map)
        return cls(captions, params, out_masks, vae_scaling_factor)

    def __len__(self) -> int:
        return len(self.captions)

    def sample(self, idx: int, eps: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """``self.vae_latents[idx].sample() * self.vae_scaling_factor`` (dataset.py:186); ``eps`` injects the draw."""
        p = self.params[idx]
        mean, logvar = torch.chunk(p, 2, dim=1)
        if eps is None:
            eps = torch.randn(mean.shape, generator=generator, device=mean.device, dtype=torch.float32)
        return ops.latent_sample(mean.contiguous(), logvar.contiguous(), eps.to(mean.device, torch.float32),
                                 self.vae_scaling_factor)

    def __getitem__(self, idx: int) -> Tuple[str, torch.Tensor, torch.Tensor]:
        return self.captions[idx], self.sample(idx).squeeze().detach(), self.masks[idx].detach()
