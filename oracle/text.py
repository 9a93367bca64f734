"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Text side of the step: CLIP encoders from the installed ``transformers`` (random
init - no weights or tokenizer vocab exist in this image), the reference's
token initialisation (trainer/embedding_handler.py:157-223) and diffusers'
``encode_prompt`` semantics as used by get_conditioning_signals
(trainer/inference.py:131-177): SD1.5 -> final-LN last hidden
state; SDXL -> ``hidden_states[-2]`` of both encoders concatenated on the feature
dim, pooled = projected ``text_embeds`` of encoder 2, add_time_ids =
[1024, 1024, 0, 0, res, res] (original_size hard-coded, inference.py:159).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

BOS, EOS = 49406, 49407


def clip_configs(family: str, tiny: bool = False):
    from transformers import CLIPTextConfig
    if tiny:
        kw = dict(vocab_size=128, max_position_embeddings=77, bos_token_id=126, eos_token_id=127, pad_token_id=127)
        c1 = CLIPTextConfig(hidden_size=64 if family != "sd15" else 96, intermediate_size=128, num_hidden_layers=2,
                            num_attention_heads=2, hidden_act="quick_gelu", **kw)
        c2 = CLIPTextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                            hidden_act="gelu", projection_dim=64, **kw)
    else:
        kw = dict(vocab_size=49408, max_position_embeddings=77, bos_token_id=BOS, eos_token_id=EOS, pad_token_id=EOS)
        c1 = CLIPTextConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                            hidden_act="quick_gelu", projection_dim=768, **kw)
        c2 = CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                            hidden_act="gelu", projection_dim=1280, **kw)
    return (c1, None) if family == "sd15" else (c1, c2)


def build_text_encoders(family: str, tiny: bool = False, seed: int = 0):
    from transformers import CLIPTextModel, CLIPTextModelWithProjection
    torch.manual_seed(seed)
    c1, c2 = clip_configs(family, tiny)
    te1 = CLIPTextModel(c1)
    te2 = CLIPTextModelWithProjection(c2) if c2 is not None else None
    return te1, te2


def initialize_new_tokens(text_encoders, n_tokens: int, seed: int = 0) -> List[int]:
    """embedding_handler.py:157-223 without a tokenizer: append ``n_tokens`` rows, fill them with
    randn rescaled to the table's mean row-std.  (transformers>=4.46 ``resize_token_embeddings``
    draws its own random rows first; they are overwritten here, so only the RNG stream differs.)"""
    torch.manual_seed(seed)
    train_ids = None
    for te in text_encoders:
        if te is None:
            continue
        emb = te.text_model.embeddings.token_embedding
        old = emb.weight.data
        vocab, dim = old.shape
        new = torch.empty(vocab + n_tokens, dim, dtype=old.dtype, device=old.device)
        new[:vocab] = old
        std_token_embedding = old.std(dim=1).mean()
        init = torch.randn(n_tokens, dim).to(device=old.device).to(dtype=old.dtype)
        init = init * 1.0 * std_token_embedding / init.std(dim=1).mean()
        new[vocab:] = init
        emb.weight = torch.nn.Parameter(new, requires_grad=emb.weight.requires_grad)
        emb.num_embeddings = vocab + n_tokens
        te.config.vocab_size = vocab + n_tokens
        train_ids = list(range(vocab, vocab + n_tokens))
    return train_ids


def encode_prompt(family_is_sdxl: bool, text_encoders, token_ids: List[torch.Tensor]):
    """Returns (prompt_embeds, pooled_prompt_embeds or None); runs WITH grad (main.py:306)."""
    te1, te2 = text_encoders
    if not family_is_sdxl:
        return te1(token_ids[0])[0], None
    embeds, pooled = [], None
    for te, ids in zip((te1, te2), token_ids):
        out = te(ids, output_hidden_states=True)
        pooled = out[0]
        embeds.append(out.hidden_states[-2])
    return torch.concat(embeds, dim=-1), pooled


def add_time_ids(batch: int, resolution: int, dtype, device):
    ids = torch.tensor([[1024, 1024, 0, 0, resolution, resolution]], dtype=dtype, device=device)
    return ids.repeat(batch, 1)
