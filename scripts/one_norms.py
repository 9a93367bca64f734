"""Launch the HBM-bound kernels at the step's dominant shapes (for `ncu --set full` captures and quick CUDA-event timing):
LayerNorm fwd/bwd [2048, 1280], GroupNorm(+SiLU) fwd/bwd [2 x 1024, 1280] and [2 x 16384, 320], GEGLU fwd/bwd [2048, 2 x 5120],
AdamW over 25.4 M elements, the batched LoRA weight-gradient launch of one transformer block."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sd_lora_trainer_b200 import ops

BF = torch.bfloat16
dev = "cuda"


def rnd(*s, scale=1.0, dtype=BF):
    return (torch.randn(*s, device=dev) * scale).to(dtype)


def timeit(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # evict L2 (126 MB) between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:46s} {med:8.1f} us  {nbytes / med / 1e3:7.0f} GB/s (algorithmic bytes {nbytes / 1e6:.1f} MB, L2 flushed)", flush=True)


x = rnd(2048, 1280)
g, b = rnd(1280), rnd(1280)
y, st = ops.layernorm_fwd(x, g, b)
dy, dres = rnd(2048, 1280), rnd(2048, 1280)
timeit("layernorm_fwd [2048,1280]", lambda: ops.layernorm_fwd(x, g, b), 2 * x.numel() * 2)
timeit("layernorm_bwd [2048,1280] (+dres)", lambda: ops.layernorm_bwd(dy, x, g, st, dres), 4 * x.numel() * 2)
for (B, hw, C) in ((2, 1024, 1280), (2, 16384, 320), (2, 4096, 640)):
    xg = rnd(B * hw, C)
    gg, bg = rnd(C), rnd(C)
    yg, sg = ops.groupnorm_fwd(xg, gg, bg, B, hw, C, 32, 1e-5, True)
    dyg = rnd(B * hw, C)
    timeit(f"groupnorm_fwd+silu [{B}x{hw},{C}]", lambda: ops.groupnorm_fwd(xg, gg, bg, B, hw, C, 32, 1e-5, True), 3 * xg.numel() * 2)
    timeit(f"groupnorm_bwd+silu [{B}x{hw},{C}]", lambda: ops.groupnorm_bwd(dyg, xg, gg, bg, sg, B, hw, C, 32, True), 5 * xg.numel() * 2)
h = rnd(2048, 10240)
dyh = rnd(2048, 5120)
timeit("geglu_fwd [2048, 2x5120]", lambda: ops.geglu_fwd(h), h.numel() * 2 + dyh.numel() * 2)
timeit("geglu_bwd [2048, 2x5120]", lambda: ops.geglu_bwd(dyh, h), 2 * h.numel() * 2 + dyh.numel() * 2)
n = 25_425_920
p, m, v = rnd(n, scale=0.01), torch.zeros(n, dtype=BF, device=dev), torch.zeros(n, dtype=BF, device=dev)
gr = rnd(n, scale=1e-3, dtype=torch.float32)
timeit("adamw 25.4M elements", lambda: ops.adamw(p, gr, m, v, n, lr=1e-4, wd=0.004, l1_coeff=1e-9, lr2=1e-3, wd2=0.0, step=3, zero_grad=True), 14 * n)
M, C, r = 2048, 1280, 16
probs = []
for i in range(6):
    probs.append((rnd(M, C), rnd(M, r), torch.zeros(C, r, device=dev), M, C, r, r, 1))
    probs.append((rnd(M, C), rnd(M, r), torch.zeros(r, C, device=dev), M, C, r, 1, C))
timeit("lora_wgrad_batch 12 x [2048 x 1280 x 16]", lambda: ops.lora_wgrad_batch(probs), 12 * (M * C * 2 + M * r * 2))
