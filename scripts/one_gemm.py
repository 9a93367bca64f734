"""Launch a handful of identical GEMMs (for ncu --set full captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
r = int(sys.argv[5]) if len(sys.argv) > 5 else 16          # side rank (48 = the fused q|k|v projection)
BF = torch.bfloat16
a = torch.randn(M, K, device="cuda").to(BF)
b = torch.randn(N, K, device="cuda").to(BF)
out = torch.empty(M, N, dtype=BF, device="cuda")
A = torch.randn(r, K, device="cuda").to(BF)
Bm = torch.randn(N, r, device="cuda").to(BF)
T = torch.empty(M, r, dtype=BF, device="cuda")
for _ in range(6):
    if mode == "plain":
        ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)])
    else:
        ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)],
                 side=(ops.Mat(A, r, K, K), ops.Mat(Bm, N, r, r), r, 1.0, T), pair_mode=1)
torch.cuda.synchronize()
