import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops  # noqa: E402

BF = torch.bfloat16


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
    return e0.elapsed_time(e1) / (replays * n_in_graph) * 1e3


def main():
    res = {"stages": os.environ.get("B200_GEMM_STAGES", "max")}
    for M, N, K, bn in [(2048, 16, 1280, 0), (2048, 1280, 1280, 0), (2048, 1280, 1280, 256), (2048, 1280, 64, 0), (2048, 1280, 128, 0),
                        (8192, 8192, 2048, 0), (128, 128, 1280, 0), (128, 128, 64, 0)]:
        a = torch.randn(M, K, device="cuda").to(BF)
        b = torch.randn(N, K, device="cuda").to(BF)
        out = torch.empty(M, N, dtype=BF, device="cuda")
        res[f"{M}x{N}x{K}/bn{bn}"] = round(graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], block_n=bn)), 2)
    # an empty-ish kernel for the launch floor
    x = torch.zeros(1024, dtype=BF, device="cuda")
    res["silu_1k (launch floor)"] = round(graph_time(lambda: ops.silu_fwd(x)), 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
