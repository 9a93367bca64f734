"""Launch a few fused-attention forward / backward kernels (for ncu --set full captures): B=2, H=10, L=Lk=4096."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

B, L = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 10
C = H * 64
BF = torch.bfloat16
q, k, v, do = (torch.randn(B * L, C, device="cuda").to(BF) for _ in range(4))
for _ in range(4):
    o, lse = ops.flash_attn_fwd(q, k, v, B, H, L, L, 0.125)
    ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, L, 0.125)
torch.cuda.synchronize()
