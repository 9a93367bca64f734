"""Developer probe: true per-launch GPU time of the tcgen05 GEMMs on the step's shapes (CUDA-graph replay, so no host
launch overhead): single-CTA kernel vs CTA-pair kernel vs torch.matmul (cuBLAS) on the same box."""
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


def main():
    rows = []
    pair_ok = os.environ.get("B200_GEMM2", "1") != "0"
    for M, N, K in [(2048, 1280, 1280), (8192, 640, 640), (2048, 10240, 1280), (2048, 1280, 5120), (8192, 5120, 640),
                    (8192, 640, 2560), (8192, 8192, 8192), (2048, 16, 1280), (154, 1280, 2048)]:
        a, b, out = r(M, K), r(N, K), torch.empty(M, N, dtype=BF, device=dev)
        t1 = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], pair_mode=-1))
        t2 = None
        if pair_ok and M >= 256 and N >= 64:
            t2 = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], pair_mode=1))
        tc = graph_time(lambda: torch.matmul(a, b.t(), out=out))
        fl = 2.0 * M * N * K
        rows.append({"kind": "plain", "M": M, "N": N, "K": K, "single_us": round(t1, 2), "pair_us": t2 and round(t2, 2),
                     "cublas_us": round(tc, 2), "single_tflops": round(fl / t1 / 1e6, 1),
                     "pair_tflops": t2 and round(fl / t2 / 1e6, 1), "cublas_tflops": round(fl / tc / 1e6, 1)})
        print(json.dumps(rows[-1]), flush=True)
    if pair_ok:
        M, N, K = 2048, 1280, 1280
        a, b, out = r(M, K), r(N, K), torch.empty(M, N, dtype=BF, device=dev)
        for bn in (128, 160, 192, 224, 256):
            t = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], pair_mode=1, block_n=bn))
            rows.append({"kind": "pair_bn_sweep", "M": M, "N": N, "K": K, "bn": bn, "us": round(t, 2)})
            print(json.dumps(rows[-1]), flush=True)
    if pair_ok:
        # where does the MN-major (dgrad) form lose time?  plain K-major vs plain MN-major B at the same tile width
        M, N, K = 2048, 1280, 1280
        a, w, out = r(M, K), r(N, K), torch.empty(M, N, dtype=BF, device=dev)
        for bn in (128, 256):
            tk = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(w), K)], pair_mode=1, block_n=bn))
            tm = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.mnmajor(w), K)], pair_mode=1, block_n=bn))
            rows.append({"kind": "pair_major", "bn": bn, "b_kmajor_us": round(tk, 2), "b_mnmajor_us": round(tm, 2)})
            print(json.dumps(rows[-1]), flush=True)
        rr = 16
        A, Bm, T = r(rr, K), r(N, rr), torch.empty(M, rr, dtype=BF, device=dev)
        for bn in (128, 192, 224, 256):
            t = graph_time(lambda: ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(w), K)], pair_mode=1, block_n=bn,
                                            side=(ops.Mat(A, rr, K, K), ops.Mat(Bm, N, rr, rr), rr, 1.0, T)))
            rows.append({"kind": "pair_side_fwd_bn", "bn": bn, "us": round(t, 2)})
            print(json.dumps(rows[-1]), flush=True)
        for bn in (128, 256):
            t = graph_time(lambda: ops.gemm(out, M, K, [(ops.kmajor(a), ops.mnmajor(w), N)], pair_mode=1, block_n=bn,
                                            side=(ops.Mat(Bm, N, rr, rr, mn=True), ops.Mat(A, rr, K, K, mn=True), rr, 1.0, T)))
            rows.append({"kind": "pair_side_dgrad_bn", "bn": bn, "us": round(t, 2)})
            print(json.dumps(rows[-1]), flush=True)
    for M, N, K, rr in [(2048, 1280, 1280, 16), (8192, 640, 640, 16)]:
        x, w, A, Bm = r(M, K), r(N, K), r(rr, K), r(N, rr)
        T, y = torch.empty(M, rr, dtype=BF, device=dev), torch.empty(M, N, dtype=BF, device=dev)
        side = (ops.Mat(A, rr, K, K), ops.Mat(Bm, N, rr, rr), rr, 1.0, T)
        t_f1 = graph_time(lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], side=side, pair_mode=-1))
        t_f2 = graph_time(lambda: ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], side=side, pair_mode=1)) if pair_ok else None
        dy, dx, U = r(M, N), torch.empty(M, K, dtype=BF, device=dev), torch.empty(M, rr, dtype=BF, device=dev)
        bside = (ops.Mat(Bm, N, rr, rr, mn=True), ops.Mat(A, rr, K, K, mn=True), rr, 1.0, U)
        t_b1 = graph_time(lambda: ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)], side=bside, pair_mode=-1))
        t_b2 = graph_time(lambda: ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)], side=bside, pair_mode=1)) if pair_ok else None
        t_torch = graph_time(lambda: x @ w.t() + (x @ A.t()) @ Bm.t())
        dA = torch.zeros(rr, K, dtype=torch.float32, device=dev)
        t_wg = graph_time(lambda: ops.gemm(dA, K, rr, [(ops.mnmajor(x), ops.mnmajor(U), M)], d_strides=(1, K, 0, 0),
                                           splits=14, atomic=True))
        rows.append({"kind": "lora", "M": M, "N": N, "K": K, "r": rr, "fwd_single_us": round(t_f1, 2),
                     "fwd_pair_us": t_f2 and round(t_f2, 2), "dgrad_single_us": round(t_b1, 2),
                     "dgrad_pair_us": t_b2 and round(t_b2, 2), "torch_3gemm_us": round(t_torch, 2), "wgrad_us": round(t_wg, 2)})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
