"""Developer probe: globaltimer stamps inside the CTA-pair GEMM (entry, ready, loads issued, first/last k-block, T-phase,
epilogue start/end, exit) for the step's dominant shapes."""
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
names = ["entry", "ready (sync+pdl)", "tile0 loads issued", "1st k-block", "last k-block", "epi start", "epi done", "pre-exit",
         "T-phase start", "T ready (MMA)"]


def probe(label, run):
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
    t0 = int(t[:, 0].min())
    print(f"{label}: {t.shape[0]} CTAs; us since first CTA entry (min / mean / max over CTAs that stamped)")
    for i, n in enumerate(names):
        col = t[:, i]
        col = col[col > 0]
        if col.numel() == 0:
            continue
        rel = (col - t0).float() / 1e3
        print(f"   {n:20s} {float(rel.min()):7.2f} {float(rel.mean()):7.2f} {float(rel.max()):7.2f}")


M, N, K, r = 2048, 1280, 1280, 16
x, w = torch.randn(M, K, device="cuda").to(BF), torch.randn(N, K, device="cuda").to(BF)
A, Bm = torch.randn(r, K, device="cuda").to(BF), torch.randn(N, r, device="cuda").to(BF)
T, y = torch.empty(M, r, dtype=BF, device="cuda"), torch.empty(M, N, dtype=BF, device="cuda")
for bn in (160, 256):
    probe(f"pair plain {M}x{N}x{K} bn={bn}", lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], pair_mode=1, block_n=bn))
probe(f"pair plain MN-major B bn=256", lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.mnmajor(w), K)], pair_mode=1, block_n=256))
for bn in (192, 256):
    probe(f"pair fused side fwd bn={bn}", lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], pair_mode=1, block_n=bn,
                                                          side=(ops.Mat(A, r, K, K), ops.Mat(Bm, N, r, r), r, 1.0, T)))
probe("pair fused side dgrad bn=256", lambda: ops.gemm(y, M, K, [(ops.kmajor(x), ops.mnmajor(w), N)], pair_mode=1, block_n=256,
                                                      side=(ops.Mat(Bm, N, r, r, mn=True), ops.Mat(A, r, K, K, mn=True), r, 1.0, T)))
