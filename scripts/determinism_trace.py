"""Find the first kernel whose output differs between two identical forward passes of the tiny net: every ops.* call is
wrapped, its outputs are hashed (bitwise, via int64 sums of the raw words) after a device sync, and the two hash
sequences are compared.   B200_STREAMK=0 python scripts/determinism_trace.py [sdxl|sd15]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sd_lora_trainer_b200 import ops
from tests.test_unet_gpu import _build, _product

fam = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 8
TRACE = []


def _hash(t):
    if not torch.is_tensor(t) or not t.is_cuda:
        return None
    t = t.contiguous()
    raw = t.view(torch.uint8).to(torch.int64)
    w = torch.arange(1, raw.numel() + 1, device=raw.device, dtype=torch.int64) % 1000003
    return int((raw.flatten() * w).sum())


def wrap(name, fn):
    def inner(*a, **k):
        out = fn(*a, **k)
        torch.cuda.synchronize()
        outs = out if isinstance(out, (tuple, list)) else (out,)
        hs = [_hash(o) for o in outs]
        if name == "gemm":                      # gemm(out, M, N, segs, ...) writes its first argument
            hs = [_hash(a[0])]
            desc = f"gemm M={a[1]} N={a[2]} K={[s[2] for s in a[3]]} side={k.get('side') is not None} nb={k.get('nb0', 1)}x{k.get('nb1', 1)} splits={k.get('splits', 1)} atomic={k.get('atomic', False)}"
        else:
            desc = name
        TRACE.append((desc, hs))
        return out
    return inner


for name in ["gemm", "flash_attn_fwd", "flash_attn_bwd", "softmax_fwd", "groupnorm_fwd", "groupnorm_bwd", "layernorm_fwd", "layernorm_bwd",
             "geglu_fwd", "geglu_bwd", "silu_fwd", "silu_bwd", "add", "upsample2x_fwd", "im2col3x3", "head_pad", "timestep_embedding",
             "noise_prologue", "diffusion_loss", "shift_sum9", "colsum", "bicubic_fwd"]:
    setattr(ops, name, wrap(name, getattr(ops, name)))

cfg, orc, inputs = _build(fam, rank=rank, batch=2, **({"disable_ti": True} if rank == 64 else {}))
tr = _product(cfg, orc)
runs = []
for i in range(3):
    TRACE.clear()
    tr.store.grads.zero_()
    out = tr.step(inputs, completion_f=0.0, do_optimizer=False)
    torch.cuda.synchronize()
    runs.append((list(TRACE), float(out["img_loss"])))
print("B200_STREAMK =", os.environ.get("B200_STREAMK", "1 (default)"), fam, "losses", [r[1] for r in runs])
for j in (1, 2):
    a, b = runs[0][0], runs[j][0]
    assert len(a) == len(b)
    first = next((i for i in range(len(a)) if a[i][1] != b[i][1]), None)
    ndiff = sum(1 for i in range(len(a)) if a[i][1] != b[i][1])
    print(f"run{j} vs run0: {ndiff} of {len(a)} ops differ; first differing op:", None if first is None else (first, a[first][0]))
    if first is not None:
        kinds = {}
        for i in range(len(a)):
            if a[i][1] != b[i][1]:
                kinds[a[i][0].split(" ")[0]] = kinds.get(a[i][0].split(" ")[0], 0) + 1
        print("   first 6 differing:", [(i, a[i][0]) for i in range(len(a)) if a[i][1] != b[i][1]][:6])
