"""Launch a handful of identical GEMMs (for ncu --set full captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
BF = torch.bfloat16
a = torch.randn(M, K, device="cuda").to(BF)
b = torch.randn(N, K, device="cuda").to(BF)
out = torch.empty(M, N, dtype=BF, device="cuda")
A = torch.randn(16, K, device="cuda").to(BF)
Bm = torch.randn(N, 16, device="cuda").to(BF)
T = torch.empty(M, 16, dtype=BF, device="cuda")
for _ in range(6):
    if mode == "plain":
        ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)])
    else:
        ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)],
                 side=(ops.Mat(A, 16, K, K), ops.Mat(Bm, N, 16, 16), 16, 1.0, T))
torch.cuda.synchronize()
