"""One LoRA + textual-inversion training step on a B200 (main.py:263-382 of the reference), re-plumbed:

    prologue kernel (offset noise + add_noise) -> UNetB200.forward -> fused loss + dPred kernel
    -> token-attention regulariser on the captured scores -> UNetB200.backward (dA/dB into one flat fp32 buffer,
       d(prompt_embeds), d(pooled)) -> CLIP backward (stock torch) into the n_tokens embedding rows
    -> [data parallel: ONE all-reduce of the flat gradient buffer] -> ONE fused AdamW over LoRA factors + TI rows.

Every random draw is an INPUT (latent sample, noise, offset noise, timesteps, token ids) so a step can be replayed
against the oracle from identical state.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .arch import UNetArch, by_name
from .text import add_time_ids, encode_prompt, init_ti_rows, install_ti_rows
from .trainer.loss import (DistributionLoss, alphas_cumprod_table, compute_diffusion_loss,
                           token_attention_loss_tensors)
from .unet import UNetB200

BF16 = torch.bfloat16


@dataclass
class StepConfig:
    """The TrainingConfig fields the step reads, with the reference's defaults (trainer/config.py:38-119)."""
    family: str = "sdxl"               # "sdxl" | "sd15"
    tiny: bool = False
    resolution: int = 1024
    lora_rank: int = 16
    is_lora: bool = True               # False: full-UNet fine-tune (main.py:143-148; BASELINE config 5)
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
    freeze_unet_before_completion_f: float = 0.0
    unet_optimizer_type: str = "adamw"       # "adamw" | "prodigy" | "AdamW8bit" -> AdamW (trainer/optimizer.py:6-39)
    ti_optimizer: str = "adamw"              # "adamw" | "prodigy" (trainer/optimizer.py:107-155)
    prodigy_d_coef: float = 1.0
    unet_prodigy_growth_factor: float = 1.05
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
        if cfg.ti_optimizer == "prodigy":                      # main.py:269: no decay, no freeze; the group lr stays 1.0
            ti_lr = 1.0
    base = 2.0e-4 if cfg.disable_ti else 5.0e-5
    if not cfg.is_lora:
        base = 1.0e-5                                           # main.py:239-240
    warm = cfg.unet_lr_warmup_steps if cfg.unet_lr_warmup_steps is not None else cfg.max_train_steps
    unet_lr = base * (cfg.unet_lr / base) ** (global_step / warm)
    if completion_f < cfg.freeze_unet_before_completion_f:     # main.py:290-291
        unet_lr = 0.0
    if cfg.unet_lr <= 0.0:                                     # main.py:164-176: no UNet optimizer at all (ti_SDXL.json)
        unet_lr = 0.0
    return ti_lr, unet_lr


class _PinnedRing:
    """A small ring of pinned host buffers for the per-step scalar block: the async H2D copy of step n must have READ its
    buffer before the host repacks it, and a host that runs ahead of the stream (graph replay, no per-step read-back) would
    otherwise overwrite it.  next() hands out the oldest slot after waiting for its copy; push() enqueues the copy."""

    def __init__(self, numel: int, slots: int = 8):
        cuda = torch.cuda.is_available()
        self.bufs = [torch.zeros(numel, dtype=torch.float32).pin_memory() if cuda else torch.zeros(numel, dtype=torch.float32)
                     for _ in range(slots)]
        self.events = [None] * slots
        self.i = -1

    def next(self) -> torch.Tensor:
        self.i = (self.i + 1) % len(self.bufs)
        if self.events[self.i] is not None:
            self.events[self.i].synchronize()
        return self.bufs[self.i]

    def push(self, dev: torch.Tensor):
        dev.copy_(self.bufs[self.i], non_blocking=True)
        if dev.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.events[self.i] = ev


class TrainerB200:
    """Owns the UNet executor, the flat parameter/gradient/moment buffers and the text encoders."""

    def __init__(self, cfg: StepConfig, unet_state_dict: Dict[str, torch.Tensor], text_encoders: Sequence,
                 device="cuda:0", ti_init: Optional[List[torch.Tensor]] = None, process_group=None,
                 use_cuda_graph: bool = False, native_text: Optional[bool] = None):
        if cfg.unet_optimizer_type == "AdamW8bit":
            # declared substitution: bitsandbytes' blockwise 8-bit states need its quantisation maps,
            # which are not available offline; the step runs AdamW with bf16 states instead - same update rule, the
            # moments are kept at higher precision than the reference keeps them
            import warnings
            warnings.warn("unet_optimizer_type='AdamW8bit' runs as AdamW with bf16 moments on the B200 path")
        elif cfg.unet_optimizer_type not in ("adamw", "prodigy"):
            raise NotImplementedError(f"Invalid optimizer_name for unet: {cfg.unet_optimizer_type}")
        if cfg.ti_optimizer not in ("adamw", "prodigy"):
            raise NotImplementedError(f"Invalid optimizer_name: '{cfg.ti_optimizer}'")
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
        self.dense_mode = not cfg.is_lora
        self.unet = UNetB200(cfg.arch(), unet_state_dict, 0 if self.dense_mode else cfg.lora_rank,
                             cfg.lora_alpha_multiplier, self.device, ti_elems=ti_elems, lora_seed=cfg.seed + 2,
                             dense=self.dense_mode)
        self.store = self.unet.store
        self.dense = self.unet.dense
        self.acp = alphas_cumprod_table(device=self.device)
        self.train_ids: List[int] = []
        self.ti_rows: List[torch.Tensor] = []
        self.std_regs: List[DistributionLoss] = []
        self._std_mu: List[float] = []
        self._std_var: List[float] = []
        if ntok:
            g = torch.Generator().manual_seed(cfg.seed)
            off = self.store.n_lora
            for i, te in enumerate(t for t in self.text_encoders if t is not None):
                table = te.text_model.embeddings.token_embedding.weight.data
                vocab, dim = table.shape
                rows = self.store.params[off:off + ntok * dim].view(ntok, dim)
                rows.copy_(ti_init[i].to(self.device, BF16) if ti_init is not None else init_ti_rows(table, ntok, g))
                self.std_regs.append(DistributionLoss(torch.cat([table, rows], dim=0)))      # loss.py:179-194
                self._std_mu.append(float(self.std_regs[-1].target_stds_mean))
                self._std_var.append(float(self.std_regs[-1].target_stds_var))
                rows.requires_grad_(True)
                install_ti_rows(te, rows)
                self.ti_rows.append(rows)
                self.train_ids = list(range(vocab, vocab + ntok))
                off += ntok * dim
            self.unet.set_capture(True)                        # init_daam_loss, main.py:50-52
        # disable_ti: the reference still ADDS the tokens (main.py:92-100 runs unconditionally) - their rows are initialised
        # and simply never trained, and captions / caption dropout (main.py:302-305) keep using their ids
        self.frozen_rows: List[torch.Tensor] = []
        if cfg.disable_ti and cfg.n_tokens > 0:
            g = torch.Generator().manual_seed(cfg.seed)
            for i, te in enumerate(t for t in self.text_encoders if t is not None):
                table = te.text_model.embeddings.token_embedding.weight.data
                rows = (ti_init[i].to(self.device, BF16) if ti_init is not None else init_ti_rows(table, cfg.n_tokens, g)).detach()
                install_ti_rows(te, rows)
                self.frozen_rows.append(rows)
        # Text encoders: stock transformers modules under autograd (default), or the explicit fwd/bwd executor over
        # our kernels (clip.py; B200_NATIVE_CLIP=1 / native_text=True).
        if native_text is None:
            native_text = os.environ.get("B200_NATIVE_CLIP", "0") == "1"
        # token-attention backward with one shared gradient map per resolution (measured on the B200: 78.2 vs 78.7 ms/step,
        # same losses / gradients - profiles/r02a_*; B200_SHARED_DSCORES=0 restores the per-layer autograd path)
        self.shared_dscores = os.environ.get("B200_SHARED_DSCORES", "1") == "1"
        # the regulariser itself as kernels (b200_token_attention_loss; forward and gradient map in three launches) instead of
        # ~40 torch ops + autograd on the stacked maps; B200_TAL_KERNEL=0 keeps the torch form of trainer/loss.py
        self.tal_kernel = os.environ.get("B200_TAL_KERNEL", "1") == "1"
        # conditioning cache for the phases without a trainable token row (B200_TEXT_CACHE=0 turns it off)
        self.cache_text = os.environ.get("B200_TEXT_CACHE", "1") != "0"
        self._text_cache: Dict[tuple, tuple] = {}
        self.text = None
        if native_text and self.text_encoders[0] is not None:
            from .clip import TextStackB200
            self.text = TextStackB200(self.sdxl, self.text_encoders, self.ti_rows or self.frozen_rows, self.device)
        self._prodigy = {}                                     # segment -> (lo, hi, scalars, hyper_host, hyper_dev, kwargs)
        segs = []
        if cfg.unet_optimizer_type == "prodigy":
            if self.dense_mode:
                raise NotImplementedError("prodigy drives the LoRA factors here; full fine-tuning runs on AdamW")
            segs.append(("unet", 0, self.store.n_lora, dict(weight_decay=cfg.lora_weight_decay, d_coef=cfg.prodigy_d_coef,
                                                            growth_rate=cfg.unet_prodigy_growth_factor)))
        if cfg.ti_optimizer == "prodigy" and ntok:
            segs.append(("ti", self.store.n_lora, self.store.params.numel(),
                         dict(weight_decay=cfg.ti_weight_decay, d_coef=1.0, growth_rate=float("inf"))))
        if segs:
            # prodigy state beside the Adam moments: s and the start point p0 (trainer/optimizer.py:24-34, 135-144)
            self._prodigy_s = torch.zeros_like(self.store.params)
            self._prodigy_p0 = self.store.params.detach().clone()
            for name, lo, hi, kw in segs:
                host = torch.zeros(12, dtype=torch.float32)
                if torch.cuda.is_available():
                    host = host.pin_memory()
                self._prodigy[name] = (lo, hi, ops.prodigy_init_scalars(1e-6, self.device), host,
                                       torch.zeros(12, dtype=torch.float32, device=self.device), kw)
        self.global_step = 0
        self.opt_step = 0
        self._accum = 0
        self.last_lrs = (None, None)
        self.use_graph = use_cuda_graph
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self._static_key = None
        self._graphs: Dict[object, object] = {}
        self.h2d_bytes_last = 0
        self._tid = None
        self.launches_per_step = 0
        self._hyper_dev = torch.zeros(12, dtype=torch.float32, device=self.device)
        self._hyper_ring = _PinnedRing(12)
        for name, (lo, hi, scal, host, dev, kw) in list(self._prodigy.items()):
            self._prodigy[name] = (lo, hi, scal, _PinnedRing(12), dev, kw)

    def _time_ids(self, B: int) -> torch.Tensor:
        if self._tid is None or self._tid.shape[0] != B:      # built once, outside any graph capture
            self._tid = add_time_ids(B, self.cfg.resolution, BF16, self.device)
        return self._tid

    # ---- inputs: host dict -> device tensors (static buffers when the step is graph-captured) ------
    # ---- conditioning cache: with no trainable token row in play, a caption's embeddings never change ----------
    def _cached_text(self, token_ids: Sequence[torch.Tensor]):
        """(prompt_embeds [B, 77, D], pooled [B, P] | None) from a per-caption cache keyed by the token ids; only the
        captions not seen yet go through the (frozen) text encoders, without autograd.  Used when the token rows are
        absent (disable_ti) or frozen (ti_lr = 0): the reference re-encodes every caption every step (main.py:306)."""
        ids = [t.long().cpu() for t in token_ids]
        B = ids[0].shape[0]
        keys = [tuple(int(v) for t in ids for v in t[b].tolist()) for b in range(B)]
        miss = [b for b in range(B) if keys[b] not in self._text_cache]
        miss = [b for i, b in enumerate(miss) if keys[b] not in [keys[c] for c in miss[:i]]]     # one encode per caption
        if miss:
            sub = [t[miss].to(self.device) for t in ids]
            with torch.no_grad():
                if self.text is not None:
                    self.text.prepare(sub[0].shape[1])
                    pe, pooled = self.text.encode_prompt(sub, need_bwd=False)
                else:
                    pe, pooled = encode_prompt(self.sdxl, self.text_encoders, sub)
            for j, b in enumerate(miss):
                self._text_cache[keys[b]] = (pe[j].detach().clone(), None if pooled is None else pooled[j].detach().clone())
        got = [self._text_cache[k] for k in keys]
        pe = torch.stack([g[0] for g in got])
        pooled = None if got[0][1] is None else torch.stack([g[1] for g in got])
        return pe, pooled

    def _stage_inputs(self, inputs: Dict[str, torch.Tensor], ti_active: bool = True) -> Dict[str, torch.Tensor]:
        """Copies one step's inputs to the device.  Returns the dict the step body reads; under CUDA graphs the
        same tensors are reused (and overwritten) every step so the captured pointers stay valid."""
        cfg, dev = self.cfg, self.device
        B, Cc = inputs["vae_latent"].shape[:2]
        staged = {
            "vae_latent": inputs["vae_latent"].to(torch.float32),
            "noise": inputs["noise"].to(BF16),
            "mask": inputs["mask"].to(torch.float32),
            "timesteps": inputs["timesteps"].long(),
            "offset_noise": inputs["offset_noise"].to(torch.float32).reshape(B, Cc),
        }
        for i, t in enumerate(inputs["token_ids"]):
            staged[f"token_ids_{i}"] = t.long()
        if not cfg.disable_ti:
            from .trainer.loss import token_index_tensors
            staged["tok_len"], staged["ti_pos"] = token_index_tensors(inputs["token_indices"], self.train_ids)
        if bool(self.ti_rows) and ti_active:
            self._text_cache.clear()                            # the rows are about to move: cached embeddings go stale
        elif self.cache_text and self.text_encoders[0] is not None:
            staged["prompt_embeds"], pooled = self._cached_text(inputs["token_ids"])
            if pooled is not None:
                staged["pooled"] = pooled
        key = tuple((k, tuple(v.shape)) for k, v in staged.items())
        if self.sdxl:
            self._time_ids(B)
        if self._static is None or self._static_key != key:
            self._static = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in staged.items()}
            self._static_key = key
            self._graphs = {}
        nbytes = 0
        for k, v in staged.items():
            src = v if (v.is_cuda or v.is_pinned() or not torch.cuda.is_available()) else v.contiguous().pin_memory()
            self._static[k].copy_(src, non_blocking=True)
            if not v.is_cuda:
                nbytes += v.numel() * v.element_size()          # host -> device traffic only
        self.h2d_bytes_last = nbytes
        return self._static

    # ---- one micro-step: forward, losses, backward ------------------------------------------------
    def forward_backward(self, inputs: Dict[str, torch.Tensor], ti_active: bool = True) -> Dict[str, torch.Tensor]:
        return self._body(self._stage_inputs(inputs, ti_active), ti_active)

    def _body(self, st: Dict[str, torch.Tensor], ti_active: bool) -> Dict[str, torch.Tensor]:
        cfg, dev = self.cfg, self.device
        ga = cfg.gradient_accumulation_steps * self.world       # DP-N == accumulation-N
        latent, mask, timesteps = st["vae_latent"], st["mask"], st["timesteps"]
        noise = st["noise"].clone()
        B, Cc, H, W = latent.shape
        offset = st["offset_noise"] if cfg.noise_offset > 0 else None
        token_ids = [st[f"token_ids_{i}"] for i in range(2 if self.sdxl else 1)]
        # Once the TI learning rate is frozen at 0 (main.py:273-274, completion_f > freeze_ti_after_completion_f, and it
        # never rises again) the token rows cannot change: AdamW with lr = 0 is the identity on them.  The reference still
        # back-propagates through both text encoders; here that backward is skipped - same parameters after the step.
        need_text_grad = bool(self.ti_rows) and ti_active
        if "prompt_embeds" in st and not need_text_grad:        # staged from the conditioning cache (or a kernel-only probe)
            prompt_embeds, pooled = st["prompt_embeds"], st.get("pooled")
        elif self.text is not None:
            self.text.prepare(token_ids[0].shape[1])
            prompt_embeds, pooled = self.text.encode_prompt(token_ids, need_bwd=need_text_grad)
        elif self.text_encoders[0] is not None:
            with torch.set_grad_enabled(need_text_grad):
                prompt_embeds, pooled = encode_prompt(self.sdxl, self.text_encoders, token_ids)
        else:                                                   # kernel-only probes: synthetic conditioning
            prompt_embeds, pooled = st["prompt_embeds"], st.get("pooled")
        time_ids = self._time_ids(B) if self.sdxl else None

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
            if self.tal_kernel:
                tal, dscores = self._token_attention_kernel(scores, mask, st["tok_len"], st["ti_pos"], ga)
            elif self.shared_dscores:
                tal, dscores = self._token_attention_shared(scores, mask, st["tok_len"], st["ti_pos"], ga)
            else:
                leaves = [s.detach().requires_grad_(True) for s in scores]
                tal = token_attention_loss_tensors(leaves, mask, st["tok_len"], st["ti_pos"])
                (cfg.token_attention_loss_w * tal / ga).backward()
                dscores = [l.grad if l.grad is not None else torch.zeros_like(l) for l in leaves]
            out["token_attention_loss"] = tal.detach().reshape(())
            out["attention_scores"] = scores
            total = total + cfg.token_attention_loss_w * tal.detach().float().reshape(())
        if cfg.l1_penalty > 0.0 and self.store.n_lora > 0:      # main.py:353: only with LoRA parameters
            l1 = torch.zeros(1, dtype=torch.float32, device=dev)
            ops.abs_sum(self.store.params[:self.store.n_lora], l1)
            total = total + cfg.l1_penalty * l1 / self.store.numel_logical
        d_ctx, d_text = self.unet.backward(dpred8, dscores, need_dctx=need_text_grad or self.text_encoders[0] is None)
        out["d_prompt_embeds"] = d_ctx
        if need_text_grad:
            roots, grads = [], []
            if self.text is not None:                         # explicit CLIP backward: TI-row gradients land in the flat buffer
                off, views = self.store.n_lora, []
                for rows in self.ti_rows:
                    views.append(self.store.grads[off:off + rows.numel()].view(rows.shape))
                    off += rows.numel()
                self.text.backward(d_ctx, d_text if pooled is not None else None, views)
            else:
                roots, grads = [prompt_embeds], [d_ctx.to(prompt_embeds.dtype)]
                if pooled is not None and d_text is not None:
                    roots.append(pooled)
                    grads.append(d_text.to(pooled.dtype))
            if ti_active:                                     # token-std regulariser, loss.py:222-231 (kernel: value + row gradients)
                off, gviews = self.store.n_lora, []
                for rows in self.ti_rows:
                    gviews.append(self.store.grads[off:off + rows.numel()].view(rows.shape))
                    off += rows.numel()
                std = ops.token_std_loss([r.detach() for r in self.ti_rows], gviews, self._std_mu, self._std_var, 0.01 / ga)
                out["token_std_loss"] = std.reshape(())
                total = total + 0.01 * std.reshape(())
            if roots:
                torch.autograd.backward(roots, grads)
            off = self.store.n_lora
            for rows in self.ti_rows:                         # move the 3-row gradients into the flat fp32 buffer
                n = rows.numel()
                if rows.grad is not None:
                    self.store.grads[off:off + n] += rows.grad.float().flatten()
                    rows.grad = None
                off += n
        out["tot_loss"] = total
        return out

    def _token_attention_kernel(self, scores, mask, tok_len, ti_pos, ga):
        """compute_token_attention_loss (trainer/loss.py:10-80) and its gradient as kernels: the larger maps are resized to
        the smallest one (process_and_stack_attention_scores, ti_cross_attn_loss.py:239-268) by the bicubic kernel, the layer
        sum / regularisers / gradient map by b200_token_attention_loss; the gradient is the same map for every layer (one
        bicubic adjoint per larger resolution)."""
        import math
        img_ratio = mask.shape[-1] / mask.shape[-2]
        shapes = []
        for s_ in scores:
            wd = round(math.sqrt(s_.shape[1] * img_ratio))
            shapes.append((round(wd / img_ratio), wd))
        h, w = min(shapes, key=lambda hw_: hw_[0] * hw_[1])
        Bsz, n_text = scores[0].shape[0], scores[0].shape[2]
        maps = []
        for s_, (hi, wi) in zip(scores, shapes):
            if hi * wi != h * w:
                s_ = ops.bicubic_fwd(s_.reshape(Bsz, hi, wi, n_text), h, w).reshape(Bsz, h * w, n_text)
            maps.append(s_)
        tal, G = ops.token_attention_loss(maps, h, w, n_text, mask[:, 0], tok_len, ti_pos, self.cfg.token_attention_loss_w / ga)
        shared, out = {h * w: G}, []
        for s_, (hi, wi) in zip(scores, shapes):
            L = hi * wi
            if L not in shared:                                   # one adjoint resize per larger resolution
                shared[L] = ops.bicubic_bwd(G.reshape(Bsz, h, w, G.shape[2]), hi, wi).reshape(Bsz, L, G.shape[2])
            out.append(shared[L])
        return tal, out

    def _token_attention_shared(self, scores, mask, tok_len, ti_pos, ga):
        """Default path (B200_SHARED_DSCORES=0 turns it off): the regulariser only sees the mean of the stacked maps over layers, so every hooked layer
        of a resolution receives the SAME gradient map.  Differentiate once with the stacked tensor as the leaf, pad that
        one map to the score buffers' 80-column rows and hand the same tensor to every layer (one bicubic adjoint for the
        layers that were resized) - instead of 60 per-layer gradients, 60 zero-fill + copy pairs and 10 bicubic adjoints.
        Values are identical to the per-layer path."""
        import torch.nn.functional as F
        from .trainer.loss import process_and_stack_attention_scores, token_attention_loss_from_maps
        img_ratio = mask.shape[-1] / mask.shape[-2]
        with torch.no_grad():
            maps = process_and_stack_attention_scores([s.detach() for s in scores], img_ratio)
        maps.requires_grad_(True)
        tal = token_attention_loss_from_maps(maps, mask, tok_len, ti_pos)
        (self.cfg.token_attention_loss_w * tal / ga).backward()
        n_text = maps.shape[-1]
        pad = (n_text + 7) // 8 * 8 - n_text
        if maps.grad is None:
            return tal, [torch.zeros_like(s) for s in scores]
        G = maps.grad[0]                                          # [B, h, w, 77]; maps.grad[l] is the same for every l
        Bsz, h, w = G.shape[:3]
        shared = {h * w: F.pad(G.reshape(Bsz, h * w, n_text), (0, pad))}
        out = []
        for s in scores:
            L = s.shape[1]
            if L not in shared:                                   # a layer that was bicubic-resized down to (h, w)
                wi = round((L * img_ratio) ** 0.5)
                hi = round(wi / img_ratio)
                shared[L] = F.pad(ops.bicubic_bwd(G, hi, wi).reshape(Bsz, L, n_text), (0, pad))
            out.append(shared[L])
        return tal, out

    def close(self):
        """Release the captured CUDA graphs and their static buffers.  Under data parallelism the graphs hold the NCCL
        all-reduce kernels they captured; measured on 2 x B200, destroy_process_group() does not return while they are
        alive - call this first (main.train does at its end)."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self._graphs = {}
        self._static, self._static_key = None, None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def reset_optimizer_state(self):
        """After parameters were overwritten from a checkpoint (load_lora_weights / load_embeddings): Adam moments and
        Prodigy's s restart from zero, Prodigy's start point p0 is re-snapshotted, the d estimate restarts, cached
        conditioning is dropped (the reference builds fresh optimizers after loading, main.py:116-176)."""
        for p, g, m, v, _ in self._train_sets():
            g.zero_()
            m.zero_()
            v.zero_()
        if self._prodigy:
            self._prodigy_s.zero_()
            self._prodigy_p0.copy_(self.store.params)
            for name, (lo, hi, scal, ring, dev, kw) in self._prodigy.items():
                scal.copy_(ops.prodigy_init_scalars(1e-6, self.device))
        self._text_cache.clear()
        self.opt_step = 0

    # ---- optimizer: ONE kernel over LoRA factors + TI rows -------------------------------------------
    def _l1_coeff(self) -> float:
        if self.cfg.l1_penalty <= 0.0 or self.store.n_lora == 0:
            return 0.0
        # what autograd hands to every LoRA element for  loss += l1_penalty * sum|p| / numel  in bf16
        return float(torch.tensor(self.cfg.l1_penalty, dtype=BF16) / self.store.numel_logical)

    def _set_hyper(self):
        """Pack this optimizer step's scalars on the host and refresh the device copy (outside any graph)."""
        ti_lr, unet_lr = self.last_lrs
        host = self._hyper_ring.next()
        ops.adamw_pack_hyper(host, lr=unet_lr, wd=self.cfg.lora_weight_decay, l1_coeff=self._l1_coeff(),
                             lr2=ti_lr or 0.0, wd2=self.cfg.ti_weight_decay, step=self.opt_step + 1, grad_scale=1.0)
        self._hyper_ring.push(self._hyper_dev)
        for name, (lo, hi, scal, ring, dev, kw) in self._prodigy.items():
            # k counts this optimizer's steps (the package skips the increment only when every gradient is exactly zero)
            host = ring.next()
            ops.prodigy_pack_hyper(host, lr=unet_lr if name == "unet" else 1.0, k=self.opt_step,
                                   l1_coeff=self._l1_coeff() if name == "unet" else 0.0, **kw)
            ring.push(dev)

    def _train_sets(self):
        """The (params, grads, m, v, n_first) buffer sets the optimizer walks: the flat LoRA + TI buffers; in dense (full
        fine-tune) mode every UNet parameter, plus the TI rows (the LoRA store then holds nothing else, n_first = 0)."""
        sets = []
        if self.dense_mode:
            d = self.dense
            sets.append((d.params, d.grads, d.m, d.v, d.params.numel()))
        if self.store.params.numel() > 0:
            sets.append((self.store.params, self.store.grads, self.store.m, self.store.v, self.store.n_lora))
        return sets

    def _train_tensors(self):
        ts = [t for st in self._train_sets() for t in st[:4]]
        if self._prodigy:
            ts += [self._prodigy_s] + [v[2] for v in self._prodigy.values()]
        return ts

    def _optimizer_body(self, collective: bool = True):
        for p, g, m, v, n_first in self._train_sets():
            if collective and self.pg is not None and self.world > 1:
                torch.distributed.all_reduce(g, group=self.pg)                # the step's only collective (LoRA mode: one)
            if not self._prodigy or p is not self.store.params:
                ops.adamw_dev(p, g, m, v, n_first, self._hyper_dev, zero_grad=True)
                continue
            # mixed / prodigy optimizers: one launch set per segment of the flat buffer (LoRA factors | TI rows)
            for name, lo, hi in (("unet", 0, n_first), ("ti", n_first, p.numel())):
                if hi <= lo:
                    continue
                if name in self._prodigy:
                    _, _, scal, _, dev, _ = self._prodigy[name]
                    ops.prodigy_step(p[lo:hi], g[lo:hi], self._prodigy_s[lo:hi], self._prodigy_p0[lo:hi], m[lo:hi], v[lo:hi],
                                     scal, dev, zero_grad=True)
                else:
                    ops.adamw_dev(p[lo:hi], g[lo:hi], m[lo:hi], v[lo:hi], hi - lo if name == "unet" else 0, self._hyper_dev,
                                  zero_grad=True)

    def optimizer_step(self):
        self._set_hyper()
        self._optimizer_body()
        self.opt_step += 1

    def step(self, inputs, completion_f: float = 0.0, do_optimizer: bool = True, optimizer_now: Optional[bool] = None):
        """One micro-step.  The optimizer runs every ``gradient_accumulation_steps`` calls unless ``optimizer_now``
        states it explicitly (main.py:365-366 also steps on the last batch of an epoch)."""
        ti_lr, unet_lr = lr_schedule(self.cfg, self.global_step, completion_f)
        self.last_lrs = (ti_lr, unet_lr)
        ti_active = bool(ti_lr and ti_lr > 0.0)
        self._accum += 1
        opt_now = do_optimizer and self._accum % self.cfg.gradient_accumulation_steps == 0
        if optimizer_now is not None:
            opt_now = bool(optimizer_now)
        if self.use_graph:
            out = self._graph_step(inputs, ti_active, opt_now)
        else:
            from . import _lib
            l0 = _lib.launch_count() if torch.cuda.is_available() else 0
            out = self.forward_backward(inputs, ti_active=ti_active)
            if opt_now:
                self.optimizer_step()
            if torch.cuda.is_available():
                self.launches_per_step = _lib.launch_count() - l0
        self.global_step += 1
        return out

    # ---- CUDA-graph path: the whole step (text encoders, UNet fwd/bwd, losses, all-reduce, AdamW) is ONE graph ----
    def _graph_step(self, inputs, ti_active: bool, opt_now: bool):
        st = self._stage_inputs(inputs, ti_active)
        if opt_now:
            self._set_hyper()
        key = (ti_active, opt_now)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(st, ti_active, opt_now)
            self._graphs[key] = entry
        graph, out = entry
        graph.replay()
        if opt_now:
            self.opt_step += 1
        return out

    def step_resident(self, completion_f: float = 0.0):
        """One more step on the inputs already staged in HBM (no host->device traffic): bench.py's `value` leg."""
        assert self._static is not None, "stage a batch with step() first"
        ti_lr, unet_lr = lr_schedule(self.cfg, self.global_step, completion_f)
        self.last_lrs = (ti_lr, unet_lr)
        ti_active = bool(ti_lr and ti_lr > 0.0)
        if self.use_graph:
            self._set_hyper()
            graph, out = self._graphs[(ti_active, True)]
            graph.replay()
            self.opt_step += 1
        else:
            out = self._body(self._static, ti_active)
            self.optimizer_step()
        self.global_step += 1
        return out

    def profile_gemms(self, inputs) -> Dict[str, object]:
        """Eager, instrumented pass over one step: every distinct tcgen05 GEMM signature is timed live (CUDA events
        around a graph of 10 launches of the very call, ops._profile_gemm) and weighted by its launch count.
        Training state is restored afterwards."""
        snap = [t.clone() for t in self._train_tensors()]
        st = self._stage_inputs(inputs)
        ops.GEMM_PROFILE = {}
        try:
            self._body(st, True)
            self._set_hyper()
            self._optimizer_body(collective=False)      # rank-local pass (bench.py runs it on rank 0 only)
            torch.cuda.synchronize()
            recs = ops.GEMM_PROFILE
        finally:
            ops.GEMM_PROFILE = None
        for t, s_ in zip(self._train_tensors(), snap):
            t.copy_(s_)
        for rows in self.ti_rows:
            rows.grad = None
        by_shape: Dict[str, Dict[str, float]] = {}
        tot_ms = tot_fl = 0.0
        launches = 0
        for key, r in recs.items():
            ms = r["count"] * r["us"] / 1e3
            tot_ms += ms
            tot_fl += r["flop"]
            launches += r["count"]
            by_shape[str(key)] = {"count": r["count"], "us_per_launch": r["us"], "ms": ms, "flop": r["flop"],
                                  "tflops": r["flop"] / max(ms, 1e-9) / 1e9}
        return {"launches": launches, "ms": tot_ms, "tflop": tot_fl / 1e12, "tflops": tot_fl / max(tot_ms, 1e-9) / 1e9,
                "by_shape": dict(sorted(by_shape.items(), key=lambda kv: -kv[1]["ms"]))}

    def _capture(self, st, ti_active: bool, opt_now: bool):
        """Warm up eagerly on a side stream, then capture.  The warm-up runs must not change training state, so the
        flat buffers are snapshotted and restored around them."""
        snap = [t.clone() for t in self._train_tensors()]

        def run():
            out = self._body(st, ti_active)
            if opt_now:
                self._optimizer_body()
            return out

        from . import _lib
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                l0 = _lib.launch_count()
                run()
                self.launches_per_step = _lib.launch_count() - l0     # our kernels per step (graph replays them)
        torch.cuda.current_stream().wait_stream(side)
        for t, s_ in zip(self._train_tensors(), snap):
            t.copy_(s_)
        for rows in self.ti_rows:
            rows.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = run()
        # the capture itself executed nothing; state is still the snapshot
        return graph, out
