import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16
lib = _lib.load()
lib.b200_debug_gemm_stamps.argtypes = [ctypes.c_void_p]
lib.b200_debug_gemm_stamps.restype = None
names = ["entry", "setup done", "1st TMA issued", "1st full", "acc committed", "epi start", "epi done", "exit", "chunk0 ld done", "chunk0 staged", "chunk0 stored"]
for M, N, K in [(2048, 1280, 64)]:
    a = torch.randn(M, K, device="cuda").to(BF)
    b = torch.randn(N, K, device="cuda").to(BF)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    for _ in range(3):
        ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)])
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib.b200_debug_gemm_stamps(buf.data_ptr())
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)])
    torch.cuda.synchronize()
    lib.b200_debug_gemm_stamps(None)
    t = buf.view(148, 16).cpu()
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    rel = (t - t0).float() / 1e3
    print(f"{M}x{N}x{K}: {t.shape[0]} CTAs; us since first CTA entry  (min / mean / max)")
    for i, n in enumerate(names):
        print(f"   {n:16s} {float(rel[:, i].min()):7.2f} {float(rel[:, i].mean()):7.2f} {float(rel[:, i].max()):7.2f}")

# fused side-path variant
M, N, K, r = 2048, 1280, 1280, 16
x = torch.randn(M, K, device="cuda").to(BF)
w = torch.randn(N, K, device="cuda").to(BF)
A = torch.randn(r, K, device="cuda").to(BF)
Bm = torch.randn(N, r, device="cuda").to(BF)
T = torch.empty(M, r, dtype=BF, device="cuda")
y = torch.empty(M, N, dtype=BF, device="cuda")
run = lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], side=(ops.Mat(A, r, K, K), ops.Mat(Bm, N, r, r), r, 1.0, T))
for _ in range(3):
    run()
buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib.b200_debug_gemm_stamps(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.b200_debug_gemm_stamps(None)
t = buf.view(148, 16).cpu()
t = t[t[:, 0] > 0]
rel = (t - int(t[:, 0].min())).float() / 1e3
print(f"fused side {M}x{N}x{K} r={r}: {t.shape[0]} CTAs")
for i, n in enumerate(names):
    print(f"   {n:16s} {float(rel[:, i].min()):7.2f} {float(rel[:, i].mean()):7.2f} {float(rel[:, i].max()):7.2f}")
