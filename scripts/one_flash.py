"""Launch a few fused-attention forward / backward kernels (for ncu --set full captures): B=2, H=10, L=Lk=4096."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

B, L = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 10
LK = int(os.environ.get("LK", L))                 # keys per sequence (default: self-attention)
C = H * 64
BF = torch.bfloat16
q, do = (torch.randn(B * L, C, device="cuda").to(BF) for _ in range(2))
k, v = (torch.randn(B * LK, C, device="cuda").to(BF) for _ in range(2))
for _ in range(4):
    o, lse = ops.flash_attn_fwd(q, k, v, B, H, L, LK, 0.125)
    ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, LK, 0.125)
torch.cuda.synchronize()
if os.environ.get("TIME", "0") == "1":              # CUDA-event timing (not under a profiler): forward / backward per call
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 20
    ev[0].record()
    for _ in range(n):
        ops.flash_attn_fwd(q, k, v, B, H, L, LK, 0.125)
    ev[1].record()
    for _ in range(n):
        ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, LK, 0.125)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"flash B={B} H={H} L={L} Lk={LK}: fwd {ev[0].elapsed_time(ev[1]) / n * 1e3:.1f} us  bwd(+delta, convert) "
          f"{ev[1].elapsed_time(ev[2]) / n * 1e3:.1f} us  TAILSPLIT={os.environ.get('B200_FLASH_TAILSPLIT', '1')}")
