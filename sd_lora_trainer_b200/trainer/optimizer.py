"""Mirror of the reference's trainer/optimizer.py (6-39, 74-155, 237-275) for the AdamW path.

The reference builds torch.optim.AdamW objects over PEFT parameter lists and writes the learning rate into
``optimizer.param_groups[0]['lr']`` every step (main.py:271, 288).  Here both "optimizers" are views of ONE flat
buffer updated by ONE fused kernel; the param_groups surface is kept so the caller's LR writes still work."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .. import arch as _arch
from .. import ops
from ..unet import UNetB200

BF16 = torch.bfloat16


class FlatAdamW:
    """torch.optim.AdamW-shaped handle on a segment of the flat LoRA/TI buffer."""

    def __init__(self, store, segment: str, lr: float, weight_decay: float):
        self.store, self.segment = store, segment
        self.param_groups = [{"lr": lr, "weight_decay": weight_decay, "betas": (0.9, 0.999), "eps": 1e-8}]

    def zero_grad(self, set_to_none: bool = True):
        pass            # the fused kernel zeroes the gradients it consumes


class FlatProdigy(FlatAdamW):
    """prodigyopt.Prodigy-shaped handle (trainer/optimizer.py:22-34, 134-144) on a segment of the flat buffer: same
    ``param_groups[0]['lr']`` surface; the state (s, p0, the d scalars) lives here, exp_avg / exp_avg_sq reuse the
    store's moment buffers."""

    def __init__(self, store, segment: str, lr: float, weight_decay: float, d_coef: float, growth_rate: float):
        super().__init__(store, segment, lr, weight_decay)
        lo, hi = (0, store.n_lora) if segment == "lora" else (store.n_lora, store.params.numel())
        self.lo, self.hi, self.d_coef, self.growth_rate, self.k = lo, hi, d_coef, growth_rate, 0
        self.s = torch.zeros(hi - lo, dtype=BF16, device=store.params.device)
        self.p0 = store.params[lo:hi].detach().clone()
        self.scal = ops.prodigy_init_scalars(1e-6, store.params.device)
        self.hyper_host = torch.zeros(12, dtype=torch.float32)
        self.hyper_dev = torch.zeros(12, dtype=torch.float32, device=store.params.device)

    def step(self, l1_coeff: float = 0.0):
        g, st = self.param_groups[0], self.store
        ops.prodigy_pack_hyper(self.hyper_host, lr=g["lr"], weight_decay=g["weight_decay"], d_coef=self.d_coef,
                               growth_rate=self.growth_rate, k=self.k, l1_coeff=l1_coeff)
        self.hyper_dev.copy_(self.hyper_host)
        sl = slice(self.lo, self.hi)
        ops.prodigy_step(st.params[sl], st.grads[sl], self.s, self.p0, st.m[sl], st.v[sl], self.scal, self.hyper_dev)
        self.k += 1


class OptimizerCollection:
    """optimizer.py:237-275: .optimizers[key], .step(), .zero_grad(), .get_lr(key)."""

    def __init__(self, optimizer_textual_inversion=None, optimizer_text_encoders=None, optimizer_unet=None, debug=False,
                 l1_penalty: float = 0.0):
        if optimizer_text_encoders is not None:
            raise NotImplementedError("text-encoder LoRA is outside the accelerated path (SURVEY.md 2, row 3)")
        self.debug = debug
        self.optimizers = {"textual_inversion": optimizer_textual_inversion, "text_encoders": None, "unet": optimizer_unet}
        self.learning_rate_tracker = {"textual_inversion": [], "text_encoders": [], "unet": []}
        self.l1_penalty = l1_penalty
        self.steps = 0

    def get_lr(self, key):
        opt = self.optimizers[key]
        return None if opt is None else opt.param_groups[0]["lr"]

    def zero_grad(self):
        pass

    def step(self):
        unet_opt, ti_opt = self.optimizers["unet"], self.optimizers["textual_inversion"]
        store = (unet_opt or ti_opt).store
        self.steps += 1
        l1c = float(torch.tensor(self.l1_penalty, dtype=BF16) / store.numel_logical) if self.l1_penalty > 0 else 0.0
        g_u = unet_opt.param_groups[0] if unet_opt is not None else {"lr": 0.0, "weight_decay": 0.0}
        g_t = ti_opt.param_groups[0] if ti_opt is not None else {"lr": 0.0, "weight_decay": 0.0}
        if not isinstance(unet_opt, FlatProdigy) and not isinstance(ti_opt, FlatProdigy):
            ops.adamw(store.params, store.grads, store.m, store.v, store.n_lora, lr=g_u["lr"], wd=g_u["weight_decay"],
                      l1_coeff=l1c, lr2=g_t["lr"], wd2=g_t["weight_decay"], step=self.steps, zero_grad=True)
        else:                                          # per segment: LoRA factors | token rows
            n, nl = store.params.numel(), store.n_lora
            for opt, lo, hi, grp, l1 in ((unet_opt, 0, nl, g_u, l1c), (ti_opt, nl, n, g_t, 0.0)):
                if hi <= lo:
                    continue
                if isinstance(opt, FlatProdigy):
                    opt.step(l1)
                else:
                    sl = slice(lo, hi)
                    ops.adamw(store.params[sl], store.grads[sl], store.m[sl], store.v[sl], hi - lo, lr=grp["lr"],
                              wd=grp["weight_decay"], l1_coeff=l1, lr2=0.0, wd2=0.0, step=self.steps, zero_grad=True)
        if self.debug:
            for key in self.optimizers:
                if self.optimizers[key] is not None:
                    self.learning_rate_tracker[key].append(self.get_lr(key))


def get_unet_lora_parameters(lora_rank, lora_alpha_multiplier: float, lora_weight_decay: float, use_dora: bool, unet,
                             pipe, ti_elems: int = 0, lora_seed: int = 0):
    """optimizer.py:74-105.  Builds the B200 executor with LoRA injected on to_k/to_q/to_v/to_out.0/conv2, sets
    ``pipe.unet`` and returns (unet, param_groups, lora_parameter_list)."""
    if use_dora:
        raise NotImplementedError("DoRA is not part of the accelerated path (no BASELINE config uses it)")
    unet = UNetB200(_arch.by_name(pipe.family), pipe.unet_state, lora_rank, lora_alpha_multiplier, pipe.device,
                    ti_elems=ti_elems, lora_seed=lora_seed)
    pipe.unet = unet
    pipe.unet_state = None
    lora_params = [unet.store.params[:unet.store.n_lora]]
    groups = [{"params": lora_params, "weight_decay": lora_weight_decay}]
    return unet, groups, lora_params


def get_unet_optimizer(prodigy_d_coef: float, prodigy_growth_factor: float, lora_weight_decay: float, use_dora: bool,
                       unet_trainable_params: Iterable, optimizer_name="adamw", unet: Optional[UNetB200] = None):
    """optimizer.py:6-39.  Only 'adamw' runs on the fused kernel; the placeholder lr 1e-4 is overwritten each step."""
    wd = lora_weight_decay if not use_dora else 0.0
    if optimizer_name == "AdamW8bit":
        # declared substitution: bitsandbytes' quantisation maps are not available offline
        import warnings
        warnings.warn("optimizer_name='AdamW8bit' runs as AdamW with bf16 moments on the B200 path")
        opt = FlatAdamW(unet.store, "lora", 1e-4, lora_weight_decay)
    elif optimizer_name == "adamw":
        opt = FlatAdamW(unet.store, "lora", 1e-4, wd)
    elif optimizer_name == "prodigy":
        opt = FlatProdigy(unet.store, "lora", 1.0, wd, prodigy_d_coef, prodigy_growth_factor)
    else:
        raise NotImplementedError(f"Invalid optimizer_name for unet: {optimizer_name}")
    print(f"Created {optimizer_name} optimizer for unet!")
    return opt


def get_textual_inversion_optimizer(text_encoders: list, textual_inversion_lr: float, textual_inversion_weight_decay,
                                    optimizer_name: str, unet: Optional[UNetB200] = None):
    """optimizer.py:107-155.  Returns (optimizer, parameter list); the parameters are the n_tokens rows only."""
    if optimizer_name == "prodigy":
        opt = FlatProdigy(unet.store, "ti", 1.0, textual_inversion_weight_decay, 1.0, float("inf"))
    elif optimizer_name == "adamw":
        opt = FlatAdamW(unet.store, "ti", textual_inversion_lr, textual_inversion_weight_decay)
    else:
        raise NotImplementedError(f"Invalid optimizer_name: '{optimizer_name}'")
    print(f"Created {optimizer_name} optimizer for textual inversion!")
    return opt, [unet.store.params[unet.store.n_lora:]]
