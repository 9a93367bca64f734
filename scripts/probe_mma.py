import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sd_lora_trainer_b200 import ops
from scripts.probe_one import graph_time
BF = torch.bfloat16
res = {}
K = 10240
for N, bn in [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]:
    a = torch.randn(128, K, device="cuda").to(BF)
    b = torch.randn(N, K, device="cuda").to(BF)
    out = torch.empty(128, N, dtype=BF, device="cuda")
    t = graph_time(lambda: ops.gemm(out, 128, N, [(ops.kmajor(a), ops.kmajor(b), K)], block_n=bn), n_in_graph=5, replays=5)
    res[f"1cta_bn{bn}"] = {"us": round(t, 1), "ns_per_kblock": round(t * 1e3 / (K / 64), 1), "rows_per_block": 128 + bn}
print(json.dumps(res))
