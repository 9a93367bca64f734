"""Text side of the step.  The CLIP encoders stay on stock torch / transformers (2-4 % of step FLOPs);
what changes is HOW the trainable textual-inversion rows are held: the reference makes the whole 49k-row embedding
table trainable and zeroes every gradient row but the last ``n_tokens`` (main.py:368-371, optimizer.py:116-121).
Here the table stays frozen and the ``n_tokens`` rows are a separate small leaf that lives in the flat optimizer
buffer next to the LoRA factors, so backward and AdamW touch 3 rows instead of 2 x 49 411."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class TIEmbedding(nn.Module):
    """Drop-in for ``text_model.embeddings.token_embedding``: ids >= vocab read the trainable rows."""

    def __init__(self, frozen_weight: torch.Tensor, ti_rows: Optional[torch.Tensor]):
        super().__init__()
        self.register_buffer("weight", frozen_weight, persistent=False)
        self.vocab = frozen_weight.shape[0]
        self.ti_rows = ti_rows                       # leaf view into the flat parameter buffer, or None
        self.embedding_dim = frozen_weight.shape[1]
        self.num_embeddings = self.vocab + (0 if ti_rows is None else ti_rows.shape[0])

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        base = F.embedding(ids.clamp(max=self.vocab - 1), self.weight)
        if self.ti_rows is None:
            return base
        is_ti = ids >= self.vocab
        # one-hot matmul instead of ti_rows[index]: the gather's backward (indexing_backward_kernel) serialises over duplicate
        # indices - 119 us per encoder with every non-TI position clamped onto row 0; this form's backward is one tiny GEMM
        sel = (ids[..., None] == self.vocab + torch.arange(self.ti_rows.shape[0], device=ids.device)).to(self.ti_rows.dtype)
        return torch.where(is_ti[..., None], sel @ self.ti_rows, base)


def install_ti_rows(text_encoder, ti_rows: Optional[torch.Tensor]):
    """Swap the encoder's token embedding for a TIEmbedding over its (frozen) table."""
    emb = text_encoder.text_model.embeddings
    old = emb.token_embedding
    w = old.weight.data if not isinstance(old, TIEmbedding) else old.weight
    emb.token_embedding = TIEmbedding(w, ti_rows)
    return emb.token_embedding


def init_ti_rows(table: torch.Tensor, n_tokens: int, generator: torch.Generator) -> torch.Tensor:
    """initialize_new_tokens (trainer/embedding_handler.py:199-213): randn rescaled to the table's mean row-std."""
    std_token_embedding = table.std(dim=1).mean()
    init = torch.randn(n_tokens, table.shape[1], generator=generator).to(device=table.device).to(dtype=table.dtype)
    return init * 1.0 * std_token_embedding / init.std(dim=1).mean()


def encode_prompt(is_sdxl: bool, text_encoders: Sequence, token_ids: List[torch.Tensor]):
    """diffusers encode_prompt as get_conditioning_signals uses it (trainer/inference.py:131-177), with grad."""
    te1, te2 = text_encoders
    if not is_sdxl:
        return te1(token_ids[0])[0], None
    embeds, pooled = [], None
    for te, ids in zip((te1, te2), token_ids):
        out = te(ids, output_hidden_states=True)
        pooled = out[0]
        embeds.append(out.hidden_states[-2])
    return torch.concat(embeds, dim=-1), pooled


def add_time_ids(batch: int, resolution: int, dtype, device) -> torch.Tensor:
    """[original 1024x1024 (hard-coded, inference.py:159), crop 0,0, target res x res] repeated per sample."""
    ids = torch.tensor([[1024, 1024, 0, 0, resolution, resolution]], dtype=dtype, device=device)
    return ids.repeat(batch, 1)
