"""Developer probe: time fwd / bwd / optimizer of one full-size step (random-init weights, synthetic inputs)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sd_lora_trainer_b200 import _lib
from sd_lora_trainer_b200.arch import by_name
from sd_lora_trainer_b200.init import random_state_dict
from sd_lora_trainer_b200.step import StepConfig, TrainerB200

ap = argparse.ArgumentParser()
ap.add_argument("--family", default="sdxl")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--rank", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--no-ti", action="store_true")
args = ap.parse_args()

dev = "cuda:0"
cfg = StepConfig(family=args.family, resolution=args.res, lora_rank=args.rank, disable_ti=True)
t0 = time.time()
sd = random_state_dict(cfg.arch(), seed=0, device=dev)
print("weights", time.time() - t0, flush=True)
tr = TrainerB200(cfg, sd, [None, None], device=dev)
del sd
a = cfg.arch()
B, hw = args.batch, args.res // 8
g = torch.Generator(device=dev).manual_seed(1)
x8 = torch.zeros(B * hw * hw, 8, dtype=torch.bfloat16, device=dev)
x8[:, :4] = torch.randn(B * hw * hw, 4, device=dev, generator=g)
ctx = torch.randn(B, 77, a.cross_attention_dim, device=dev, generator=g).to(torch.bfloat16)
pooled = torch.randn(B, 1280, device=dev, generator=g).to(torch.bfloat16) if args.family == "sdxl" else None
tid = torch.tensor([[1024, 1024, 0, 0, args.res, args.res]] * B, dtype=torch.bfloat16, device=dev) if args.family == "sdxl" else None
t = torch.randint(0, 1000, (B,), device=dev)
d8 = torch.zeros(B * hw * hw, 8, dtype=torch.bfloat16, device=dev)
d8[:, :4] = torch.randn(B * hw * hw, 4, device=dev, generator=g) * 1e-3
tr.unet.set_capture(not args.no_ti)
for it in range(args.steps):
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    w0 = time.time()
    e[0].record()
    pred, scores = tr.unet.forward(x8, B, hw, hw, t, ctx, pooled, tid)
    e[1].record()
    w1 = time.time()
    ds = [torch.zeros_like(s) for s in scores] if scores else None
    tr.unet.backward(d8, ds)
    e[2].record()
    w2 = time.time()
    tr.last_lrs = (None, 1e-4)
    tr.optimizer_step()
    e[3].record()
    torch.cuda.synchronize()
    print(json.dumps({"iter": it, "fwd_ms": e[0].elapsed_time(e[1]), "bwd_ms": e[1].elapsed_time(e[2]),
                      "opt_ms": e[2].elapsed_time(e[3]), "host_fwd_ms": (w1 - w0) * 1e3, "host_bwd_ms": (w2 - w1) * 1e3,
                      "launches": _lib.launch_count() - l0, "pred_absmax": float(pred.float().abs().max()),
                      "finite": bool(torch.isfinite(pred.float()).all()),
                      "mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
