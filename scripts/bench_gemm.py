"""Developer probe: tcgen05 GEMM throughput on the step's real shapes vs torch.matmul (cuBLAS) on the same box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda:0"


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


shapes = [(2048, 1280, 1280), (8192, 640, 640), (2048, 10240, 1280), (2048, 1280, 5120), (8192, 5120, 640),
          (8192, 640, 2560), (32768, 320, 2880), (8192, 8192, 8192), (2048, 16, 1280), (154, 1280, 2048)]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev).to(BF)
    b = torch.randn(N, K, device=dev).to(BF)
    out = torch.empty(M, N, dtype=BF, device=dev)
    t_ours = timeit(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)]))
    t_cublas = timeit(lambda: torch.matmul(a, b.t()))
    fl = 2.0 * M * N * K
    print(json.dumps({"M": M, "N": N, "K": K, "ours_us": t_ours * 1e3, "cublas_us": t_cublas * 1e3,
                      "ours_tflops": fl / t_ours / 1e9, "cublas_tflops": fl / t_cublas / 1e9}), flush=True)
# fused LoRA forward vs the reference's 3 GEMMs + add
for M, N, K, r in [(2048, 1280, 1280, 16), (8192, 640, 640, 16)]:
    x = torch.randn(M, K, device=dev).to(BF)
    w = torch.randn(N, K, device=dev).to(BF)
    A = torch.randn(r, K, device=dev).to(BF)
    Bm = torch.randn(N, r, device=dev).to(BF)
    T = torch.empty(M, r, dtype=BF, device=dev)
    y = torch.empty(M, N, dtype=BF, device=dev)

    def ours():
        ops.gemm(T, M, r, [(ops.kmajor(x), ops.kmajor(A), K)])
        ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K), (ops.kmajor(T), ops.kmajor(Bm), r)])

    def ref():
        return x @ w.t() + (x @ A.t()) @ Bm.t()
    print(json.dumps({"lora_fwd": [M, N, K, r], "ours_us": timeit(ours) * 1e3, "torch_us": timeit(ref) * 1e3}), flush=True)
