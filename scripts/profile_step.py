"""In-situ kernel time table of the benched step: torch.profiler (CUPTI activity records, no replay, no serialisation,
caches as warm as in the real step) over a few CUDA-graph replays of `TrainerB200.step_resident`.  Complements the ncu
launch list (cold-cache, serialised).   python scripts/profile_step.py [--family sdxl --rank 16 --batch 2 --res 1024]
Writes gpurun_out/step_kernels_<tag>.txt (+ .json)."""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0]] + sys.argv[1:]
ap = argparse.ArgumentParser()
ap.add_argument("--family", default="sdxl")
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--rank", type=int, default=16)
ap.add_argument("--full-ft", action="store_true")
ap.add_argument("--native-clip", action="store_true")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--tag", default="sdxl_r16")
args = ap.parse_args()

import bench  # noqa: E402  (build_text_encoders)
from sd_lora_trainer_b200.data import synthetic_inputs  # noqa: E402
from sd_lora_trainer_b200.init import random_state_dict  # noqa: E402
from sd_lora_trainer_b200.step import StepConfig, TrainerB200  # noqa: E402

dev = "cuda:0"
cfg = StepConfig(family=args.family, resolution=args.res, lora_rank=args.rank, disable_ti=args.full_ft,
                 is_lora=not args.full_ft)
sd = random_state_dict(cfg.arch(), seed=0, device=dev)
tes = bench.build_text_encoders(args.family, dev)
tr = TrainerB200(cfg, sd, tes, device=dev, use_cuda_graph=not args.full_ft, native_text=True if args.native_clip else None)
tr.cache_text = False
del sd
batch = synthetic_inputs(args.family, args.batch, args.res, 0 if args.full_ft else cfg.n_tokens, seed=1000, face_mask=True,
                         vae_scaling_factor=cfg.arch().vae_scaling_factor, pin=True)
for _ in range(3):
    out = tr.step(batch, completion_f=0.0)
float(out["tot_loss"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    tr.step_resident(completion_f=0.0)
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / args.steps
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        tr.step_resident(completion_f=0.0)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    name = ev.name.split("(")[0]
    agg[name][0] += 1
    agg[name][1] += dur
tot = sum(v[1] for v in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/step_kernels_{args.tag}"
with open(path + ".txt", "w") as f:
    f.write(f"# in-situ kernel times (CUPTI activity records over {args.steps} graph replays; per step)\n")
    f.write(f"# step (CUDA events, no profiler): {ms_plain:.2f} ms; sum of kernel durations: {tot / args.steps / 1e3:.2f} ms\n")
    for name, (n, us) in rows:
        f.write(f"{name[:90]:90s} n={n // args.steps:5d} ms={us / args.steps / 1e3:8.3f} avg_us={us / max(n, 1):8.1f} "
                f"{100 * us / tot:5.1f}%\n")
json.dump({"ms_per_step": ms_plain, "kernel_ms_per_step": tot / args.steps / 1e3,
           "kernels": {k: {"n": v[0] // args.steps, "ms": v[1] / args.steps / 1e3} for k, v in rows}}, open(path + ".json", "w"), indent=1)
print(open(path + ".txt").read()[:6000])
