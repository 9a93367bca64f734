"""Synthetic step inputs in the shapes the reference's dataloader produces (trainer/dataset.py:181-193 ->
main.py:294-324): fp32 VAE latents scaled by the VAE scaling factor, bf16 noise, offset noise, timesteps, a mask
(all-ones for 'style', a smooth blob floored at 0.05 for 'face' as CLIPSeg + blur produce, preprocess.py:213-215)
and CLIP token ids with the trainable tokens at positions 1..n (no tokenizer vocabulary exists offline)."""
from __future__ import annotations

from typing import Dict, List

import torch

BOS, EOS, VOCAB = 49406, 49407, 49408


def synthetic_inputs(family: str, batch: int, resolution: int, n_tokens: int, seed: int, face_mask: bool,
                     vae_scaling_factor: float, tiny: bool = False, pin: bool = False) -> Dict[str, object]:
    g = torch.Generator().manual_seed(seed)
    hw = resolution // 8
    bos, eos, vocab = (126, 127, 128) if tiny else (BOS, EOS, VOCAB)
    lat = torch.randn(batch, 4, hw, hw, generator=g) * vae_scaling_factor
    noise = torch.randn(batch, 4, hw, hw, generator=g).to(torch.bfloat16)
    off = torch.randn(batch, 4, 1, 1, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    if face_mask:
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, hw), torch.linspace(-1, 1, hw), indexing="ij")
        blob = torch.exp(-(xx ** 2 + yy ** 2) / 0.35).clamp_min(0.05)
        mask = (blob / blob.max())[None, None].repeat(batch, 4, 1, 1).contiguous()
    else:
        mask = torch.ones(batch, 4, hw, hw)
    ids = torch.full((batch, 77), eos, dtype=torch.long)
    token_indices: List[List[int]] = []
    for b in range(batch):
        n_words = int(torch.randint(4, 12, (1,), generator=g))
        words = torch.randint(0, bos, (n_words,), generator=g).tolist()
        seq = [bos] + [vocab + i for i in range(n_tokens)] + words + [eos]
        ids[b, :len(seq)] = torch.tensor(seq)
        token_indices.append(seq)
    out = {"vae_latent": lat, "noise": noise, "offset_noise": off, "timesteps": t, "mask": mask,
           "token_ids": [ids.clone() for _ in range(2 if family == "sdxl" else 1)], "token_indices": token_indices}
    if pin and torch.cuda.is_available():
        for k, v in list(out.items()):
            if isinstance(v, torch.Tensor):
                out[k] = v.pin_memory()
        out["token_ids"] = [v.pin_memory() for v in out["token_ids"]]
    return out
