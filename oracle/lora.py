"""ORACLE (test infrastructure; parity unpinned - see oracle/__init__.py).

Restatement of peft==0.10.0 LoRA injection as the reference configures it
(trainer/optimizer.py:74-105): targets ["to_k","to_q","to_v","to_out.0","conv2"]
suffix-matched, r = lora_rank, lora_alpha = r * lora_alpha_multiplier,
init_lora_weights="gaussian" (A ~ N(0, (1/r)^2), B = 0), dropout 0, adapter
weights in the base weight's dtype, forward
``result = base(x); result = result + lora_B(lora_A(x)) * scaling`` with every
op rounding to the storage dtype.
Parameter names follow PEFT: ``<path>.base_layer.weight``,
``<path>.lora_A.default.weight``, ``<path>.lora_B.default.weight``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn

LORA_TARGETS = ["to_k", "to_q", "to_v", "to_out.0", "conv2"]   # trainer/optimizer.py:84


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        nn.init.zeros_(self.lora_B["default"].weight)
        self.lora_A.to(base.weight.dtype)
        self.lora_B.to(base.weight.dtype)

    def forward(self, x):
        result = self.base_layer(x)
        return result + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling


class LoraConv2d(nn.Module):
    def __init__(self, base: nn.Conv2d, r: int, alpha: float):
        super().__init__()
        self.base_layer = base
        self.r = r
        self.scaling = alpha / r
        self.lora_A = nn.ModuleDict({"default": nn.Conv2d(
            base.in_channels, r, base.kernel_size, base.stride, base.padding, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Conv2d(r, base.out_channels, 1, bias=False)})
        nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)
        nn.init.zeros_(self.lora_B["default"].weight)
        self.lora_A.to(base.weight.dtype)
        self.lora_B.to(base.weight.dtype)

    def forward(self, x):
        result = self.base_layer(x)
        return result + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling


def _matches(name: str) -> bool:
    return any(name == t or name.endswith("." + t) for t in LORA_TARGETS)


def lora_target_names(model: nn.Module) -> List[str]:
    return [n for n, m in model.named_modules()
            if _matches(n) and isinstance(m, (nn.Linear, nn.Conv2d))]


def inject_lora(unet: nn.Module, lora_rank: int, lora_alpha_multiplier: float = 1.0,
                seed: int | None = None) -> Tuple[nn.Module, List[nn.Parameter]]:
    """get_unet_lora_parameters (trainer/optimizer.py:74-105): freeze everything,
    wrap the targets, return (unet, lora parameter list in named_parameters order)."""
    if seed is not None:
        torch.manual_seed(seed)
    for p in unet.parameters():
        p.requires_grad_(False)
    alpha = lora_rank * lora_alpha_multiplier
    for name in lora_target_names(unet):
        parent_name, _, attr = name.rpartition(".")
        parent = unet.get_submodule(parent_name) if parent_name else unet
        base = getattr(parent, attr) if not attr.isdigit() else parent[int(attr)]
        wrapped = (LoraLinear if isinstance(base, nn.Linear) else LoraConv2d)(base, lora_rank, alpha)
        if attr.isdigit():
            parent[int(attr)] = wrapped
        else:
            setattr(parent, attr, wrapped)
    params = [p for p in unet.parameters() if p.requires_grad]
    return unet, params
