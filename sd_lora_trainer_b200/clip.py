"""B200-native CLIP text encoders for the step's conditioning path: what
``get_conditioning_signals`` runs with grad (trainer/inference.py:131-177 -> diffusers ``encode_prompt`` ->
transformers ``CLIPTextModel`` / ``CLIPTextModelWithProjection``), as an explicit forward AND backward over the same
sm_100a kernels the UNet uses.  The encoders are frozen; the only trainable inputs are the ``n_tokens`` textual-inversion
rows, so the backward is input-gradient only (no weight gradients) and ends in ONE small GEMM that sums the embedding
gradients of the positions holding a trainable token straight into the flat fp32 gradient buffer (the reference instead
lets autograd fill two full 49 411-row tables and zeroes all but 3 rows, main.py:368-371).

Per layer (pre-LN transformer block, causal self-attention, head_dim 64, 77 tokens):
  * q/k/v projections are ONE GEMM against the stacked ``[3C, C]`` weight; q, k, v are column slices of its output;
  * scores: batched GEMM over (head, sample) whose epilogue adds the causal mask (a ``[77, 80]`` bf16 ``-inf`` upper
    triangle passed as the residual operand), then the softmax kernel, then the batched ``P.V`` GEMM;
  * out-projection and fc2 add the residual stream in their epilogues; the MLP activation (quick_gelu for CLIP-L, erf
    GELU for OpenCLIP bigG) is ``b200_act_fwd/bwd``;
  * backward mirrors it: dQ / dK / dV are written into the column slices of one ``[M, 3C]`` buffer so the input
    gradient of the three projections is ONE GEMM against the stacked weight read MN-major.
SDXL reads ``hidden_states[-2]`` of both encoders (the last layer of encoder 1 is never run) plus the projected pooled
output of encoder 2; SD1.5 reads the final-LayerNorm output of its single encoder.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import BF16, Mat, kmajor, mnmajor
from .unet import LN, Lin, _r8

_ACT = {"gelu": ops.ACT_GELU, "quick_gelu": ops.ACT_QUICK_GELU}


class _Block:
    def __init__(self, sd: Dict[str, torch.Tensor], p: str, heads: int, act: int, device):
        w = lambda n: sd[f"{p}.{n}"].detach().to(device, BF16).contiguous()
        self.ln1, self.ln2 = LN(w("layer_norm1.weight"), w("layer_norm1.bias")), LN(w("layer_norm2.weight"), w("layer_norm2.bias"))
        self.Wqkv = torch.cat([w("self_attn.q_proj.weight"), w("self_attn.k_proj.weight"), w("self_attn.v_proj.weight")], 0).contiguous()
        self.bqkv = torch.cat([w("self_attn.q_proj.bias"), w("self_attn.k_proj.bias"), w("self_attn.v_proj.bias")], 0).contiguous()
        self.out = Lin(w("self_attn.out_proj.weight"), w("self_attn.out_proj.bias"))
        self.fc1, self.fc2 = Lin(w("mlp.fc1.weight"), w("mlp.fc1.bias")), Lin(w("mlp.fc2.weight"), w("mlp.fc2.bias"))
        self.h, self.act = heads, act
        self.C = self.Wqkv.shape[1]
        self.sv = None

    def _heads(self, t: torch.Tensor, L: int, ld: int, mn: bool = False) -> Mat:
        """[B*L, H*d] column block of a row-major buffer with row stride `ld`, batched over (head, sample)."""
        d = self.C // self.h
        return Mat(t, L, d, ld, mn=mn, sb0=d, sb1=L * ld, batched=True)

    def fwd(self, x: torch.Tensor, B: int, L: int, mask: torch.Tensor, need_bwd: bool) -> torch.Tensor:
        C, H = self.C, self.h
        d, M, Lp, dev = C // H, B * L, _r8(L), x.device
        hn = self.ln1.fwd(x)
        qkv = torch.empty(M, 3 * C, dtype=BF16, device=dev)
        ops.gemm(qkv, M, 3 * C, [(kmajor(hn), kmajor(self.Wqkv), C)], bias=self.bqkv, static_b=True)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        sS = (Lp, 1, L * Lp, H * L * Lp)
        S = torch.empty(B, H, L, Lp, dtype=torch.float32, device=dev)
        ops.gemm(S, L, L, [(self._heads(q, L, 3 * C), self._heads(k, L, 3 * C), d)], d_strides=sS, alpha=d ** -0.5,
                 residual=mask, r_strides=(Lp, 1, 0, 0), nb0=H, nb1=B)           # epilogue adds the causal mask
        P = torch.empty(B, H, L, Lp, dtype=BF16, device=dev)
        ops.softmax_fwd(S, P, B * H * L, L, Lp, Lp)
        del S
        O = torch.empty(M, C, dtype=BF16, device=dev)
        ops.gemm(O, L, d, [(Mat(P, L, L, Lp, sb0=L * Lp, sb1=H * L * Lp, batched=True), self._heads(v, L, 3 * C, mn=True), L)],
                 d_strides=(C, 1, d, L * C), nb0=H, nb1=B)
        x1 = self.out.fwd(O, residual=x, save=False)
        f = self.fc1.fwd(self.ln2.fwd(x1), save=False)
        x2 = self.fc2.fwd(ops.act_fwd(f, self.act), residual=x1, save=False)
        if need_bwd:
            self.sv = (qkv, P, f, B, L)
        else:
            self.ln1.sv = self.ln2.sv = None
        return x2

    def bwd(self, dx2: torch.Tensor) -> torch.Tensor:
        qkv, P, f, B, L = self.sv
        self.sv = None
        C, H = self.C, self.h
        d, M, Lp, dev = C // H, B * L, _r8(L), dx2.device
        scale = d ** -0.5
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        df = ops.act_bwd(self.fc2.bwd(dx2), f, self.act)
        dx1 = self.ln2.bwd(self.fc1.bwd(df), dres=dx2)
        dO = self.out.bwd(dx1)
        dqkv = torch.empty(M, 3 * C, dtype=BF16, device=dev)
        dq, dk, dv = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:]
        sS = (Lp, 1, L * Lp, H * L * Lp)
        s3 = (3 * C, 1, d, L * 3 * C)
        pb = dict(sb0=L * Lp, sb1=H * L * Lp, batched=True)
        # dV = P^T dO ; dP = dO V^T ; dS = P * (dP - rowsum(P dP)) (masked entries have P = 0) ; dQ = s dS K ; dK = s dS^T Q
        ops.gemm(dv, L, d, [(Mat(P, L, L, Lp, mn=True, **pb), self._heads(dO, L, C, mn=True), L)], d_strides=s3, nb0=H, nb1=B)
        dP = torch.empty(B, H, L, Lp, dtype=torch.float32, device=dev)
        ops.gemm(dP, L, L, [(self._heads(dO, L, C), self._heads(v, L, 3 * C), d)], d_strides=sS, nb0=H, nb1=B)
        dS = torch.empty(B, H, L, Lp, dtype=BF16, device=dev)
        ops.softmax_bwd(P, dP, dS, B * H * L, L, Lp, Lp)
        del dP, P
        ops.gemm(dq, L, d, [(Mat(dS, L, L, Lp, **pb), self._heads(k, L, 3 * C, mn=True), L)], d_strides=s3, alpha=scale,
                 nb0=H, nb1=B)
        ops.gemm(dk, L, d, [(Mat(dS, L, L, Lp, mn=True, **pb), self._heads(q, L, 3 * C, mn=True), L)], d_strides=s3,
                 alpha=scale, nb0=H, nb1=B)
        dh = torch.empty(M, C, dtype=BF16, device=dev)
        ops.gemm(dh, M, C, [(kmajor(dqkv), mnmajor(self.Wqkv), 3 * C)], static_b=True)
        return self.ln1.bwd(dh, dres=dx1)


class CLIPTextB200:
    """One frozen CLIP text encoder; ``ti_rows`` ([n_tokens, C] bf16 view into the flat parameter buffer) are the rows
    ids >= vocab read.  ``use`` selects what the step consumes: "penultimate" (SDXL encoder 1), "penultimate+pooled"
    (SDXL encoder 2) or "last" (SD1.5: final-LayerNorm output)."""

    def __init__(self, text_encoder, ti_rows: Optional[torch.Tensor], use: str, device):
        assert use in ("penultimate", "penultimate+pooled", "last")
        cfg = text_encoder.config
        sd = {k: v for k, v in text_encoder.state_dict().items()}
        self.device, self.use, self.ti_rows = torch.device(device), use, ti_rows
        self.C, self.heads, self.nl = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
        assert self.C // self.heads * self.heads == self.C
        self.eos_id = cfg.eos_token_id
        emb = text_encoder.text_model.embeddings
        tok = emb.token_embedding
        self.table = (tok.weight.data if hasattr(tok, "weight") else sd["text_model.embeddings.token_embedding.weight"]).to(self.device, BF16)
        self.vocab = self.table.shape[0]
        self.pos = emb.position_embedding.weight.data.to(self.device, BF16).contiguous()
        act = _ACT[cfg.hidden_act]
        n_run = self.nl - 1 if use == "penultimate" else self.nl       # SDXL encoder 1 never needs its last layer
        self.blocks = [_Block(sd, f"text_model.encoder.layers.{i}", self.heads, act, self.device) for i in range(n_run)]
        w = lambda n: sd[n].detach().to(self.device, BF16).contiguous()
        self.final_ln = LN(w("text_model.final_layer_norm.weight"), w("text_model.final_layer_norm.bias")) if use != "penultimate" else None
        self.proj = Lin(w("text_projection.weight"), None) if use == "penultimate+pooled" else None
        self._mask: Dict[int, torch.Tensor] = {}
        self._fw = None

    def _causal_mask(self, L: int) -> torch.Tensor:
        m = self._mask.get(L)
        if m is None:
            Lp = _r8(L)
            m = torch.zeros(L, Lp, dtype=BF16, device=self.device)
            m[:, :L] = torch.full((L, L), float("-inf"), device=self.device).triu(1).to(BF16)
            self._mask[L] = m
        return m

    def prepare(self, L: int):
        """Builds the cached constants outside any CUDA-graph capture."""
        self._causal_mask(L)

    def _embed(self, ids: torch.Tensor) -> torch.Tensor:
        base = self.table[ids.clamp(max=self.vocab - 1)]
        if self.ti_rows is not None:
            rows = self.ti_rows.detach()[(ids - self.vocab).clamp(min=0)]
            base = torch.where((ids >= self.vocab)[..., None], rows, base)
        B, L = ids.shape
        return ops.add(base.reshape(B * L, self.C).contiguous(), self.pos[:L].repeat(B, 1).contiguous())

    def forward(self, ids: torch.Tensor, need_bwd: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """ids [B, L] int64.  Returns (hidden [B, L, C] as `use` says, projected pooled [B, P] or None)."""
        B, L = ids.shape
        mask = self._causal_mask(L)
        x = self._embed(ids)
        pen = None
        for i, blk in enumerate(self.blocks):
            if i == self.nl - 1:
                pen = x                                              # hidden_states[-2]: input of the last layer
            x = blk.fwd(x, B, L, mask, need_bwd)
        hidden, pooled, eos_rows = None, None, None
        if self.use == "penultimate":
            hidden = x
        else:
            last = self.final_ln.fwd(x)
            if not need_bwd:
                self.final_ln.sv = None
            hidden = last if self.use == "last" else pen
            if self.proj is not None:
                if self.eos_id == 2:                                 # transformers' legacy rule (CLIPTextTransformer.forward)
                    eos_pos = ids.to(torch.int).argmax(dim=-1)
                else:
                    eos_pos = (ids.to(torch.int) == self.eos_id).int().argmax(dim=-1)
                eos_rows = torch.arange(B, device=ids.device) * L + eos_pos
                pooled = self.proj.fwd(last.index_select(0, eos_rows).contiguous(), save=False)
        self._fw = (ids, B, L, eos_rows) if need_bwd else None
        return hidden.view(B, L, self.C), pooled

    def backward(self, d_hidden: torch.Tensor, d_pooled: Optional[torch.Tensor], grad_rows: Optional[torch.Tensor]):
        """d_hidden [B, L, C] bf16 (gradient of what forward returned), d_pooled [B, P] or None.  Adds the gradient of
        the trainable rows into ``grad_rows`` ([n_tokens, C] fp32 view of the flat gradient buffer)."""
        ids, B, L, eos_rows = self._fw
        self._fw = None
        M, C = B * L, self.C
        d_pen = d_hidden.reshape(M, C).to(BF16).contiguous()
        if self.use == "penultimate":
            d, first = d_pen, len(self.blocks) - 1
        else:
            if self.use == "last":
                d_ln = d_pen
            else:
                d_ln = torch.zeros(M, C, dtype=BF16, device=d_pen.device)
                if d_pooled is not None:
                    d_ln.index_copy_(0, eos_rows, self.proj.bwd(d_pooled.to(BF16).contiguous()))
            d = self.final_ln.bwd(d_ln)
            d = self.blocks[-1].bwd(d)
            if self.use != "last":
                d = ops.add(d, d_pen)
            first = len(self.blocks) - 2
        for i in range(first, -1, -1):
            d = self.blocks[i].bwd(d)
        if grad_rows is not None and self.ti_rows is not None:
            n = self.ti_rows.shape[0]
            Mp = _r8(M)
            sel = torch.zeros(n, Mp, dtype=BF16, device=d.device)
            sel[:, :M] = (ids.reshape(1, M) == (self.vocab + torch.arange(n, device=ids.device)).reshape(n, 1)).to(BF16)
            # grad_rows[j, :] += sum over positions holding token vocab+j of d(x0)[position, :]
            ops.gemm(grad_rows, n, C, [(Mat(sel, n, M, Mp), mnmajor(d), M)], d_strides=(C, 1, 0, 0), atomic=True)
        return d


class TextStackB200:
    """Both encoders of a family behind ``encode_prompt``'s contract (trainer/inference.py:131-177)."""

    def __init__(self, is_sdxl: bool, text_encoders, ti_rows: List[Optional[torch.Tensor]], device):
        te1, te2 = text_encoders
        rows = list(ti_rows) + [None, None]
        if is_sdxl:
            self.encs = [CLIPTextB200(te1, rows[0], "penultimate", device), CLIPTextB200(te2, rows[1], "penultimate+pooled", device)]
        else:
            self.encs = [CLIPTextB200(te1, rows[0], "last", device)]
        self.is_sdxl = is_sdxl

    def prepare(self, L: int):
        for e in self.encs:
            e.prepare(L)

    def encode_prompt(self, token_ids: List[torch.Tensor], need_bwd: bool = True):
        outs = [e.forward(ids, need_bwd) for e, ids in zip(self.encs, token_ids)]
        if not self.is_sdxl:
            return outs[0][0], None
        return torch.cat([outs[0][0], outs[1][0]], dim=-1), outs[1][1]

    def backward(self, d_prompt_embeds: torch.Tensor, d_pooled: Optional[torch.Tensor], grad_rows: List[Optional[torch.Tensor]]):
        if not self.is_sdxl:
            self.encs[0].backward(d_prompt_embeds, None, grad_rows[0] if grad_rows else None)
            return
        c1 = self.encs[0].C
        self.encs[0].backward(d_prompt_embeds[..., :c1], None, grad_rows[0] if grad_rows else None)
        self.encs[1].backward(d_prompt_embeds[..., c1:], d_pooled, grad_rows[1] if grad_rows else None)
