"""Mirror of the reference's trainer/embedding_handler.py for the hot path: token initialisation (157-223), the
trainable rows (25-62) and the safetensors save / load format (401-457; keys ``clip_l`` / ``clip_g``)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ..text import init_ti_rows, install_ti_rows

BF16 = torch.bfloat16


class TokenEmbeddingsHandler:
    def __init__(self, text_encoders, tokenizers=None):
        self.text_encoders = text_encoders
        self.tokenizers = tokenizers or [None] * len(text_encoders)
        self.train_ids: Optional[List[int]] = None
        self.inserting_toks: Optional[List[str]] = None
        self.embeddings_settings: Dict[str, torch.Tensor] = {}
        self.rows: List[torch.Tensor] = []
        self.token_regularizer = None

    def initialize_new_tokens(self, inserting_toks: List[str], starting_toks: Optional[List[str]] = None, seed: int = 0,
                              store=None):
        """Appends len(inserting_toks) rows per encoder, randn rescaled to the table's mean row-std; with ``store``
        (the flat LoRA/TI buffer) the rows are placed behind the LoRA factors so one AdamW kernel updates both."""
        assert isinstance(inserting_toks, list), "inserting_toks should be a list of strings."
        assert all(isinstance(tok, str) for tok in inserting_toks), "All elements in inserting_toks should be strings."
        if starting_toks is not None:
            raise NotImplementedError("starting_toks needs the CLIP tokenizer vocabulary, which is not available offline")
        self.inserting_toks = inserting_toks
        n = len(inserting_toks)
        g = torch.Generator().manual_seed(seed)
        off = store.n_lora if store is not None else 0
        for idx, te in enumerate(self.text_encoders):
            if te is None:
                continue
            table = te.text_model.embeddings.token_embedding.weight.data
            vocab, dim = table.shape
            init = init_ti_rows(table, n, g)
            if store is not None:
                rows = store.params[off:off + n * dim].view(n, dim)
                rows.copy_(init.to(rows.device, BF16))
                off += n * dim
            else:
                rows = init.clone()
            self.embeddings_settings[f"std_token_embedding_{idx}"] = table.std(dim=1).mean()
            self.train_ids = list(range(vocab, vocab + n))
            rows.requires_grad_(True)
            install_ti_rows(te, rows)
            self.rows.append(rows)

    def make_embeddings_trainable(self):
        for rows in self.rows:
            rows.requires_grad_(True)

    def get_trainable_embeddings(self):
        embeddings, tokens = {}, {}
        i = 0
        for idx, te in enumerate(self.text_encoders):
            if te is None:
                continue
            embeddings[f"txt_encoder_{idx}"] = self.rows[i]
            tokens[f"txt_encoder_{idx}"] = self.inserting_toks
            i += 1
        return embeddings, tokens

    def save_embeddings(self, file_path: str, txt_encoder_keys=("clip_l", "clip_g")):
        from safetensors.torch import save_file
        tensors, i = {}, 0
        for idx, te in enumerate(self.text_encoders):
            if te is None:
                continue
            tensors[txt_encoder_keys[idx]] = self.rows[i].detach().clone().contiguous().cpu()
            i += 1
        save_file(tensors, file_path)

    def load_embeddings(self, file_path: str, txt_encoder_keys=("clip_l", "clip_g"), trainer=None):
        """``trainer`` (TrainerB200, optional): reset its optimizer state / conditioning cache after the rows changed."""
        from safetensors.torch import load_file
        tensors, i = load_file(file_path), 0
        for idx, te in enumerate(self.text_encoders):
            if te is None:
                continue
            with torch.no_grad():
                self.rows[i].copy_(tensors[txt_encoder_keys[idx]].to(self.rows[i].device, self.rows[i].dtype))
            i += 1
        if trainer is not None:
            trainer.reset_optimizer_state()
