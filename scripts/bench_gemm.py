"""Developer probe: true per-launch GPU time of the tcgen05 GEMM on the step's shapes (CUDA-graph replay, so no host
launch overhead), next to torch.matmul (cuBLAS) on the same box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda:0"


def graph_time(fn, n_in_graph=20, replays=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n_in_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * n_in_graph) * 1e3     # us


def r(*shape):
    return torch.randn(*shape, device=dev).to(BF)


rows = []
for M, N, K in [(2048, 1280, 1280), (8192, 640, 640), (2048, 10240, 1280), (2048, 1280, 5120), (8192, 5120, 640),
                (8192, 640, 2560), (8192, 8192, 8192), (2048, 16, 1280), (154, 1280, 2048)]:
    a, b, out = r(M, K), r(N, K), torch.empty(M, N, dtype=BF, device=dev)
    t = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)]))
    tc = graph_time(lambda: torch.matmul(a, b.t(), out=out))
    fl = 2.0 * M * N * K
    rows.append({"kind": "plain", "M": M, "N": N, "K": K, "ours_us": t, "cublas_us": tc, "ours_tflops": fl / t / 1e6,
                 "cublas_tflops": fl / tc / 1e6})
    print(json.dumps(rows[-1]), flush=True)
for M, N, K, rr in [(2048, 1280, 1280, 16), (8192, 640, 640, 16)]:
    x, w, A, Bm = r(M, K), r(N, K), r(rr, K), r(N, rr)
    T, y = torch.empty(M, rr, dtype=BF, device=dev), torch.empty(M, N, dtype=BF, device=dev)
    t_fused = graph_time(lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)],
                                          side=(ops.Mat(A, rr, K, K), ops.Mat(Bm, N, rr, rr), rr, 1.0, T)))

    def two():
        ops.gemm(T, M, rr, [(ops.kmajor(x), ops.kmajor(A), K)])
        ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K), (ops.kmajor(T), ops.kmajor(Bm), rr)])
    t_two = graph_time(two)
    t_torch = graph_time(lambda: x @ w.t() + (x @ A.t()) @ Bm.t())
    dy, dA = r(M, N), torch.zeros(rr, K, dtype=torch.float32, device=dev)
    U = r(M, rr)
    t_wg = graph_time(lambda: ops.gemm(dA, K, rr, [(ops.mnmajor(x), ops.mnmajor(U), M)], d_strides=(1, K, 0, 0),
                                       splits=14, atomic=True))
    rows.append({"kind": "lora_fwd", "M": M, "N": N, "K": K, "r": rr, "fused_us": t_fused, "two_launch_us": t_two,
                 "torch_3gemm_us": t_torch, "wgrad_us": t_wg})
    print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_gemm.json", "w"), indent=1)
