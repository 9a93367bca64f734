"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Restatement of ``prodigyopt.Prodigy`` (prodigyopt==1.0, requirements.txt; absent from this image and from
/root/reference), the optimizer the reference builds for ``unet_optimizer_type="prodigy"`` / ``ti_optimizer="prodigy"``
(trainer/optimizer.py:22-34, 134-144: ``decouple=True, use_bias_correction=True, safeguard_warmup=True,
betas=(0.9, 0.99)``, ``d_coef`` / ``growth_rate`` from the config).  Algorithm: Mishchenko & Defazio, "Prodigy: An
Expeditiously Adaptive Parameter-Free Learner" (2023), Adam variant, as the package implements it:

    bias_correction = sqrt(1 - beta2^(k+1)) / (1 - beta1^(k+1))        (use_bias_correction)
    dlr             = d * lr * bias_correction
    d_numerator    *= beta3                                            (beta3 = sqrt(beta2))
    per tensor (lr > 0):
        d_numerator += (d / d0) * dlr * <grad, p0 - p>
        exp_avg      = beta1 * exp_avg    + d   * (1 - beta1) * grad
        exp_avg_sq   = beta2 * exp_avg_sq + d^2 * (1 - beta2) * grad^2
        s            = beta3 * s + (d / d0) * d * grad                 (safeguard_warmup; else ... * dlr * grad)
        d_denom     += |s|_1
    d_hat = d_coef * d_numerator / d_denom ; if d == d0: d = max(d, d_hat) ; d_max = max(d_max, d_hat)
    d     = min(d_max, d * growth_rate)
    per tensor:  denom = sqrt(exp_avg_sq) + d * eps ; p -= weight_decay * dlr * p (decoupled) ; p -= dlr * exp_avg / denom
    k += 1
State tensors take the parameter's dtype (bf16 here), every in-place op rounds as torch rounds it.
"""
from __future__ import annotations

import math

import torch


class Prodigy(torch.optim.Optimizer):
    def __init__(self, params, lr=1.0, betas=(0.9, 0.999), beta3=None, eps=1e-8, weight_decay=0.0, decouple=True,
                 use_bias_correction=False, safeguard_warmup=False, d0=1e-6, d_coef=1.0, growth_rate=float("inf")):
        defaults = dict(lr=lr, betas=betas, beta3=beta3, eps=eps, weight_decay=weight_decay, d=d0, d0=d0, d_max=d0,
                        d_numerator=0.0, d_coef=d_coef, k=0, growth_rate=growth_rate,
                        use_bias_correction=use_bias_correction, decouple=decouple, safeguard_warmup=safeguard_warmup)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        group = self.param_groups[0]
        beta1, beta2 = group["betas"]
        beta3 = group["beta3"] if group["beta3"] is not None else math.sqrt(beta2)
        k, d, d_max, d_coef = group["k"], group["d"], group["d_max"], group["d_coef"]
        lr = max(g["lr"] for g in self.param_groups)
        bias_correction = ((1 - beta2 ** (k + 1)) ** 0.5) / (1 - beta1 ** (k + 1)) if group["use_bias_correction"] else 1.0
        dlr = d * lr * bias_correction
        growth_rate, decouple = group["growth_rate"], group["decouple"]
        d_numerator = group["d_numerator"] * beta3
        d_denom = 0.0
        for group in self.param_groups:
            decay, d0, group_lr, safeguard = group["weight_decay"], group["d0"], group["lr"], group["safeguard_warmup"]
            if group_lr not in [lr, 0.0]:
                raise RuntimeError("Setting different lr values in different parameter groups is only supported for values of 0")
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.data
                if decay != 0 and not decouple:
                    grad.add_(p.data, alpha=decay)
                state = self.state[p]
                if "step" not in state:
                    state["step"] = 0
                    state["s"] = torch.zeros_like(p.data).detach()
                    state["p0"] = p.detach().clone()
                    state["exp_avg"] = torch.zeros_like(p.data).detach()
                    state["exp_avg_sq"] = torch.zeros_like(p.data).detach()
                exp_avg, exp_avg_sq, s, p0 = state["exp_avg"], state["exp_avg_sq"], state["s"], state["p0"]
                if group_lr > 0.0:
                    d_numerator += (d / d0) * dlr * torch.dot(grad.flatten(), (p0.data - p.data).flatten()).item()
                    exp_avg.mul_(beta1).add_(grad, alpha=d * (1 - beta1))
                    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=d * d * (1 - beta2))
                    if safeguard:
                        s.mul_(beta3).add_(grad, alpha=(d / d0) * d)
                    else:
                        s.mul_(beta3).add_(grad, alpha=(d / d0) * dlr)
                    d_denom += s.abs().sum().item()
        d_hat = d
        if d_denom == 0:
            return None
        if lr > 0.0:
            d_hat = d_coef * d_numerator / d_denom
            if d == group["d0"]:
                d = max(d, d_hat)
            d_max = max(d_max, d_hat)
            d = min(d_max, d * growth_rate)
        for group in self.param_groups:
            group["d_numerator"], group["d_denom"], group["d"], group["d_max"], group["d_hat"] = d_numerator, d_denom, d, d_max, d_hat
            decay, eps = group["weight_decay"], group["eps"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                state["step"] += 1
                denom = state["exp_avg_sq"].sqrt().add_(d * eps)
                if decay != 0 and decouple:
                    p.data.add_(p.data, alpha=-decay * dlr)
                p.data.addcdiv_(state["exp_avg"], denom, value=-dlr)
            group["k"] = k + 1
        return None
