"""Tile-width sweep of the CTA-pair GEMM on the step's attention-projection shapes (fused q|k|v with its rank-48 side path,
single projections with rank 16): each (shape, block_n) is timed as a CUDA graph of 20 launches between CUDA events and
compared with what pick_pair_bn chooses (block_n = 0).  Output: gpurun_out/sweep_bn.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sd_lora_trainer_b200 import ops

BF = torch.bfloat16


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).to(BF)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)


def fwd_case(M, N, K, r):
    x, w, A, B2 = rnd(M, K), rnd(N, K, scale=0.05), rnd(r, K, scale=0.1), rnd(N, r, scale=0.1)
    T, y = torch.empty(M, r, dtype=BF, device="cuda"), torch.empty(M, N, dtype=BF, device="cuda")
    return lambda bn: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], side=(ops.Mat(A, r, K, K), ops.Mat(B2, N, r, r), r, 1.0, T),
                               block_n=bn, pair_mode=1, static_b=True)


def dgrad_case(M, N, K, r):          # dX[M, K] = dY[M, N].W[N, K] + (dY.Bt^T).A
    dy, w, A, Bt = rnd(M, N), rnd(N, K, scale=0.05), rnd(r, K, scale=0.1), rnd(r, N, scale=0.1)
    U, dx = torch.empty(M, r, dtype=BF, device="cuda"), torch.empty(M, K, dtype=BF, device="cuda")
    return lambda bn: ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)],
                               side=(ops.Mat(Bt, r, N, N), ops.Mat(A, r, K, K, mn=True), r, 1.0, U), block_n=bn, pair_mode=1, static_b=True)


def plain_case(M, N, K, mn=False):
    x = rnd(M, K)
    w = rnd(K, N, scale=0.05) if mn else rnd(N, K, scale=0.05)
    y = torch.empty(M, N, dtype=BF, device="cuda")
    return lambda bn: ops.gemm(y, M, N, [(ops.kmajor(x), ops.mnmajor(w) if mn else ops.kmajor(w), K)], block_n=bn, pair_mode=1, static_b=True)


cases = [("fwd qkv 2048x3840x1280 r48", fwd_case(2048, 3840, 1280, 48), 2048 * 3840 * 1280, False),
         ("dgrad qkv 2048x1280x3840 r48", dgrad_case(2048, 3840, 1280, 48), 2048 * 3840 * 1280, True),
         ("fwd proj 2048x1280x1280 r16", fwd_case(2048, 1280, 1280, 16), 2048 * 1280 * 1280, False),
         ("dgrad proj 2048x1280x1280 r16", dgrad_case(2048, 1280, 1280, 16), 2048 * 1280 * 1280, True),
         ("fwd qkv 8192x1920x640 r48", fwd_case(8192, 1920, 640, 48), 8192 * 1920 * 640, False),
         ("dgrad qkv 8192x640x1920 r48", dgrad_case(8192, 1920, 640, 48), 8192 * 1920 * 640, True),
         ("fwd proj 8192x640x640 r16", fwd_case(8192, 640, 640, 16), 8192 * 640 * 640, False),
         ("dgrad proj 8192x640x640 r16", dgrad_case(8192, 640, 640, 16), 8192 * 640 * 640, True),
         ("ff2 fwd 2048x1280x5120", plain_case(2048, 1280, 5120), 2048 * 1280 * 5120, False),
         ("ff2 dgrad 2048x5120x1280 (W MN-major)", plain_case(2048, 5120, 1280, mn=True), 2048 * 1280 * 5120, True)]
lines = []
for name, fn, mac, mn in cases:
    row = [f"{name:40s}"]
    t0 = timed(lambda: fn(0))
    row.append(f"auto {t0:6.1f} us ({2 * mac / t0 / 1e6:5.0f} TF/s) |")
    for bn in ([128, 256] if mn else [96, 128, 160, 192, 224, 256]):
        try:
            t = timed(lambda: fn(bn))
            row.append(f"bn{bn} {t:6.1f}")
        except Exception as e:  # noqa: BLE001
            row.append(f"bn{bn} n/a")
    lines.append(" ".join(row))
    print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/sweep_bn.txt", "w").write("\n".join(lines) + "\n")
