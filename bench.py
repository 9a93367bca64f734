#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native LoRA + textual-inversion training step.

Workload (BASELINE.json metric): SDXL LoRA rank-16 'face' mode + 3 textual-inversion tokens, 1024x1024 (latent
128x128), bf16, batch 2 per GPU, random-init weights of the full SDXL-base architecture (UNet 2.57 B, CLIP-L,
OpenCLIP-bigG), synthetic inputs.  One "step" = text encoders fwd -> noise prologue -> UNet fwd (LoRA fused) ->
losses -> UNet bwd (dA/dB only) -> CLIP bwd to the TI rows -> [all-reduce] -> fused AdamW.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (CPU oracle of the same step on the host cores)
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_TFLOP_PER_IMAGE = {("sdxl", 16): 14.505, ("sdxl", 32): 14.676, ("sd15", 16): 1.761, ("sd15", 4): 1.742,
                         ("sdxl", "ft"): 20.284}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_read"] + d["dram_bytes_write"]


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_text_encoders(family: str, device):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    kw = dict(vocab_size=49408, max_position_embeddings=77, bos_token_id=49406, eos_token_id=49407, pad_token_id=49407)
    c1 = CLIPTextConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                        hidden_act="quick_gelu", projection_dim=768, **kw)
    torch.manual_seed(1)
    with torch.device(device):
        te1 = CLIPTextModel(c1).to(torch.bfloat16)
        te2 = None
        if family == "sdxl":
            c2 = CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                                hidden_act="gelu", projection_dim=1280, **kw)
            te2 = CLIPTextModelWithProjection(c2).to(torch.bfloat16)
    return te1, te2


def run_ours(args):
    from sd_lora_trainer_b200 import _lib, ops
    from sd_lora_trainer_b200.data import synthetic_inputs
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.step import StepConfig, TrainerB200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # a rank that dies leaves its peers spinning in NCCL: bound the whole run (stacks go to stderr, exit code 1)
    faulthandler.dump_traceback_later(args.watchdog, exit=True)
    # stdout carries the ONE JSON line and nothing else: NCCL's version / debug chatter goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
        pg = torch.distributed.group.WORLD
    cfg = StepConfig(family=args.family, resolution=args.res, lora_rank=args.rank, disable_ti=args.full_ft,
                     is_lora=not args.full_ft, max_train_steps=max(args.steps + args.warmup, 300))
    sd = random_state_dict(cfg.arch(), seed=0, device=dev)
    tes = build_text_encoders(args.family, dev)
    tr = TrainerB200(cfg, sd, tes, device=dev, process_group=pg, use_cuda_graph=not args.no_graph,
                     native_text=True if args.native_clip else None)
    tr.cache_text = False      # timed steps encode every caption every step, as the reference does (no cached outputs)
    del sd
    B = args.batch
    host_batches = [synthetic_inputs(args.family, B, args.res, 0 if args.full_ft else cfg.n_tokens,
                                     seed=1000 + rank * 97 + i, face_mask=True,
                                     vae_scaling_factor=cfg.arch().vae_scaling_factor, pin=True) for i in range(4)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (captures the graph) ------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        out = tr.step(host_batches[i % 4], completion_f=0.0)
    float(out["tot_loss"])
    sync_all()

    if args.profile_step:
        # ncu window (`ncu --profile-from-start off`): exactly one step between cudaProfilerStart / Stop, then leave
        rt = torch.cuda.cudart()
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        tr.step_resident(completion_f=0.0)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        if rank == 0:
            print(json.dumps({"profiled_step": True, "gpu_launches_per_step": tr.launches_per_step}), flush=True)
        return

    # ---- timed region 1: inputs resident in HBM (the staged static buffers), K steps -----------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        out = tr.step_resident(completion_f=0.0)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches_timed = _lib.launch_count() - l0
    # ---- timed region 2: end to end through the public API, host (pinned) inputs, loss read back -------
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e2.record()
    for i in range(args.steps):
        out = tr.step(host_batches[i % 4], completion_f=0.0)
        loss_val = float(out["tot_loss"])                     # device -> host read of the step's result
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.finish()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    launches_per_step = tr.launches_per_step

    result = None
    if rank == 0:
        pk, pk_src = peaks()
        imgs = B * world * args.steps
        value = imgs / (ms / 1e3)
        e2e_value = imgs / (ms_e2e / 1e3)
        tf_img = TRAIN_TFLOP_PER_IMAGE.get((args.family, "ft" if args.full_ft else args.rank))
        result = {
            "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"{args.family.upper()} full-UNet fine-tune (is_lora=false, disable_ti) " if args.full_ft else
                                    f"{args.family.upper()} LoRA r={args.rank} face+TI(3 tokens) ") + f"{args.res}x{args.res} "
                                   f"bf16 batch {B}/GPU, fwd+bwd+AdamW, random-init full-size weights",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": "working set (5 GB weights + ~25 GB activations per step) exceeds the 126 MB L2",
                       "cuda_graph": not args.no_graph, "text_encoders": "native" if tr.text is not None else "transformers",
                       "shared_dscores": bool(tr.shared_dscores)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": tr.h2d_bytes_last, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps if not launches_timed else launches_timed,
            "gpu_launches_per_step": launches_per_step,
            "loss": loss_val,
        }
        if tf_img:
            result["step_roofline"] = {"bound": "tensor", "achieved": tf_img * value / world, "peak": pk["bf16_tflops_sustained"],
                                       "unit": "TFLOP/s", "frac": tf_img * value / world / pk["bf16_tflops_sustained"],
                                       "peak_source": pk_src + " (sustained)", "train_tflop_per_image": tf_img}
    # ---- roofline of the dominant kernel (the tcgen05 GEMM), measured live with CUDA events, eager pass ------
    if rank == 0 and not args.skip_roofline:
        prof = tr.profile_gemms(host_batches[0])
        pk, pk_src = peaks()
        result["roofline"] = {"bound": "tensor", "kernel": "tcgen05 GEMM family (gemm2_kernel CTA pairs + gemm_tcgen05_kernel)",
                              "achieved": prof["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                              "frac": prof["tflops"] / pk["bf16_tflops"],
                              # dram__bytes_read.sum + dram__bytes_write.sum of the family's most frequent launch, per launch,
                              # from the committed `ncu --set full` capture (profiles/ncu_traffic.json names it); null if absent
                              "traffic": ncu_traffic(),
                              "peak_source": pk_src + " (burst: every GEMM signature is timed alone, as a CUDA graph of "
                                                      "10 launches of the step's own call, between CUDA events)",
                              "launches_per_step": prof["launches"], "gemm_ms_per_step": prof["ms"],
                              "gemm_share_of_step": prof["ms"] / (ms / args.steps),
                              "algorithmic_tflop_per_step": prof["tflop"]}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(prof["by_shape"], open(os.path.join(ROOT, "gpurun_out", "gemm_by_shape.json"), "w"), indent=1)
    if rank == 0 and world == 1 and not args.skip_cpu:
        # CPU baseline + full-size step-loss delta in a CHILD process with a timeout: a fault or a hang in that extra leg
        # (it runs GPU shapes - batch 1 - that the timed workload does not) cannot cost the measured line.  The parent is
        # idle meanwhile, so the oracle has the host cores to itself.
        child, child_err = None, None
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--delta-only", "--family", args.family,
                                 "--rank", str(args.rank), "--cpu-res", str(args.cpu_res), "--cpu-batch", str(args.cpu_batch),
                                 "--cpu-dtype", args.cpu_dtype] + (["--full-ft"] if args.full_ft else []),
                                capture_output=True, text=True, timeout=args.delta_timeout)
            lines = [ln for ln in cp.stdout.splitlines() if ln.startswith("{")]
            child = json.loads(lines[-1]) if lines else None
            child_err = None if child else f"exit {cp.returncode}: {cp.stderr[-300:]}"
        except Exception as e:                              # noqa: BLE001  (timeout, launch failure, bad output)
            child_err = f"{type(e).__name__}: {e}"[:300]
        if child is not None:
            result["cpu_baseline"], result["step_loss_delta"] = child["cpu_baseline"], child["step_loss_delta"]
        else:
            result["cpu_baseline"] = cpu_baseline(args, steps=1, warmup=0)
            result["step_loss_delta"] = {"error": child_err}
    if rank == 0 and world == 1 and not args.skip_gpu_baseline:
        # stock-torch-on-the-same-B200 comparator, in a child process (its own CUDA context and memory; bounded by a timeout)
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--gpu-baseline-only", "--family", args.family, "--rank",
                   str(args.rank), "--res", str(args.res), "--batch", str(args.batch)] + (["--full-ft"] if args.full_ft else [])
            cp = subprocess.run(cmd, capture_output=True, text=True, timeout=args.delta_timeout)
            lines = [ln for ln in cp.stdout.splitlines() if ln.startswith("{")]
            result["gpu_baseline"] = json.loads(lines[-1])["gpu_baseline"] if lines else \
                {"error": f"exit {cp.returncode}: {cp.stderr[-300:]}"}
        except Exception as e:                                  # noqa: BLE001
            result["gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        gb = result["gpu_baseline"]
        if "value" in gb:
            gb["ours_over_stock_torch"] = result["value"] / gb["value"]
    if rank == 0:
        print(json.dumps(result), flush=True)
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        # Tear-down: destroy_process_group() blocks while CUDA graphs that captured NCCL kernels are alive (measured on
        # 2 x B200 in round 1), so the graphs are released first (TrainerB200.close) and the group destroyed from a helper
        # thread with a deadline; if it still does not return, leave without the NCCL / graph destructors (the line is out).
        # Non-zero ranks wait (rendezvous store, not NCCL) until rank 0 has printed.
        try:
            import datetime
            store = torch.distributed.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("b200_bench_done", "1")
            else:
                store.wait(["b200_bench_done"], datetime.timedelta(seconds=900))
        except Exception:                                    # best effort: never turn a finished run into a failure
            pass
        sys.stdout.flush()
        clean = False
        try:
            tr.close()
            th = threading.Thread(target=torch.distributed.destroy_process_group, daemon=True)
            th.start()
            th.join(timeout=20)
            clean = not th.is_alive()
        except Exception:                                    # noqa: BLE001
            clean = False
        print(f"[bench] rank {rank}: process group {'destroyed cleanly' if clean else 'NOT destroyed within 20 s - leaving via os._exit'}",
              file=sys.stderr, flush=True)
        sys.stderr.flush()
        if not clean:
            os._exit(0)


def cpu_baseline(args, steps: int, warmup: int, keep=None):
    """The oracle (a port of the reference step, oracle/) on the host cores: SAME model and config, batch 1 so that a
    step stays bounded.  fp32 on the CPU - the reference's bf16 path is several times slower there (no AMX needed)."""
    from oracle.step import OracleTrainer, StepConfig as OCfg, make_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    dtype = torch.float32 if args.cpu_dtype == "fp32" else torch.bfloat16
    cfg = OCfg(family=args.family, resolution=args.cpu_res, lora_rank=args.rank, weight_dtype=dtype,
               is_lora=not args.full_ft, disable_ti=args.full_ft)
    t0 = time.time()
    orc = OracleTrainer(cfg, device="cpu")
    if keep is not None:
        # the oracle computes in fp32 on the values bf16 weights hold (what the GPU step loads), so that the two losses
        # differ by arithmetic only; the trainable tensors are snapshotted because the timed step updates them
        with torch.no_grad():
            for mod in [orc.unet] + [te for te in orc.text_encoders if te is not None]:
                for p_ in mod.parameters():
                    p_.copy_(p_.to(torch.bfloat16).to(p_.dtype))
        keep["pre"] = {n: p_.detach().clone() for n, p_ in orc.unet.named_parameters() if p_.requires_grad}
        keep["ti"] = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
                      for te in orc.text_encoders if te is not None] if not cfg.disable_ti else None
    build_s = time.time() - t0
    inp = make_inputs(cfg, batch=args.cpu_batch, face_mask=True, train_ids=orc.train_ids or None)
    if keep is not None:                                       # inputs at the values the bf16 step sees (main.py:311-312)
        for k in ("vae_latent", "noise"):
            inp[k] = inp[k].to(torch.bfloat16).to(inp[k].dtype)
    for _ in range(warmup):
        orc.step(inp)
    t0 = time.time()
    for _ in range(steps):
        out = orc.step(inp)
        float(out["tot_loss"])
    dt = time.time() - t0
    if keep is not None:
        keep.update(orc=orc, cfg=cfg, inputs=inp, loss=float(out["tot_loss"]), img_loss=float(out["img_loss"]))
    return {"value": args.cpu_batch * steps / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} step(s) of the same {args.family.upper()} " +
                      ("full-UNet fine-tune" if args.full_ft else f"r={args.rank} face+TI") + f" step, batch {args.cpu_batch}, "
                      f"{args.cpu_res}x{args.cpu_res}, {args.cpu_dtype} oracle on the host cores (fwd+bwd+AdamW), "
                      f"build {build_s:.0f}s not timed",
            "seconds": dt}


def step_loss_delta(keep, dev):
    """BASELINE.json metric, second half ("step-loss delta vs ref") at FULL size: the GPU step on the oracle's own
    weights (bf16 values), token rows and inputs - the state the oracle's timed step started from - against the loss that
    step reported.  Oracle: fp32 arithmetic on the host; GPU: bf16 kernels with fp32 accumulation."""
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig, TrainerB200
    orc, ocfg = keep["orc"], keep["cfg"]
    sd = dict(orc.unet.state_dict())
    sd.update(keep["pre"])                                     # trainable tensors as they were before the oracle's update
    pcfg = StepConfig(**{k: getattr(ocfg, k) for k in StepConfig.__dataclass_fields__ if hasattr(ocfg, k)})
    tes = build_text_encoders(ocfg.family, ocfg.tiny, seed=ocfg.seed + 1)       # the seed OracleTrainer builds them with
    tr2 = TrainerB200(pcfg, sd, tes, device=dev, ti_init=keep["ti"])
    tr2.cache_text = False
    out = tr2.step(keep["inputs"], completion_f=0.0, do_optimizer=False)
    a, b = float(out["tot_loss"]), keep["loss"]
    ai, bi = float(out["img_loss"]), keep["img_loss"]
    del tr2
    torch.cuda.empty_cache()
    return {"ours": a, "oracle_fp32_cpu": b, "rel": abs(a - b) / abs(b), "img_loss_ours": ai, "img_loss_oracle": bi,
            "img_loss_rel": abs(ai - bi) / abs(bi),
            "config": f"same {ocfg.family} " + (f"r={ocfg.lora_rank}" if ocfg.is_lora else "full-UNet fine-tune") + " step, batch "
            f"{keep['inputs']['vae_latent'].shape[0]}, {ocfg.resolution}x{ocfg.resolution}, identical weights / rows / inputs",
            "north_star_bound": 1e-3}


def gpu_baseline(args, steps: int = 5, warmup: int = 2):
    """SURVEY 2.3 / BASELINE.md section 3: the bar is the reference's own code path on stock torch (cuBLAS / cuDNN SDPA / ATen
    autograd / torch.optim.AdamW) ON THE SAME B200.  diffusers / peft cannot be installed, so this is the oracle port of that
    step (oracle/) in bf16 on cuda:0 - same model, same batch, same synthetic inputs as the timed workload - between CUDA
    events.  A comparator leg like cpu_baseline: nothing of it is on the product path."""
    from oracle.step import OracleTrainer, StepConfig as OCfg, make_inputs
    cfg = OCfg(family=args.family, resolution=args.res, lora_rank=args.rank, weight_dtype=torch.bfloat16,
               is_lora=not args.full_ft, disable_ti=args.full_ft)
    orc = OracleTrainer(cfg, device="cuda:0")
    inp = make_inputs(cfg, batch=args.batch, face_mask=True, train_ids=orc.train_ids or None)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    inp["token_ids"] = [t.cuda() for t in inp["token_ids"]]
    for _ in range(warmup):
        float(orc.step(inp)["tot_loss"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        out = orc.step(inp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": args.batch / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "kind": "port", "steps": steps,
            "warmup": warmup, "loss": float(out["tot_loss"]), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "sample": f"oracle port of the reference step on stock torch {torch.__version__} (cuBLAS, SDPA, ATen autograd, "
                      f"torch.optim.AdamW incl. the full embedding tables) in bf16 on the same B200, {args.family.upper()} "
                      f"r={args.rank} {args.res}x{args.res} batch {args.batch}, inputs resident, eager"}


def run_gpu_baseline_only(args):
    try:
        gb = gpu_baseline(args)
    except Exception as e:                                      # noqa: BLE001
        gb = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps({"gpu_baseline": gb}), flush=True)


def run_delta_only(args):
    """Child process of the default run: the bounded CPU-oracle step, then the GPU step on the oracle's own state."""
    keep = {}
    cb = cpu_baseline(args, steps=1, warmup=0, keep=keep)
    try:
        delta = step_loss_delta(keep, "cuda:0")
    except Exception as e:                                      # noqa: BLE001
        delta = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps({"cpu_baseline": cb, "step_loss_delta": delta}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, args.ref_max_steps), min(args.warmup, 1)
    cb = cpu_baseline(args, steps=steps, warmup=warmup)
    line = {"impl": "reference", "metric": "training images/sec", "value": cb["value"], "unit": "images/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
            "ms_per_step": cb["seconds"] / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.cpu_dtype, "data": "synthetic",
            "config": {"workload": f"{args.family.upper()} LoRA r={args.rank} face+TI {args.cpu_res}x{args.cpu_res} "
                                   f"batch {args.cpu_batch}, CPU oracle (port of the reference step; diffusers/peft absent)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--family", default="sdxl")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--rank", type=int, default=16)
    ap.add_argument("--full-ft", action="store_true",
                    help="BASELINE config 5: full-UNet fine-tune (dense backward, AdamW over every parameter), disable_ti")
    ap.add_argument("--native-clip", action="store_true",
                    help="text encoders on the native CLIP executor (clip.py) instead of the stock transformers modules")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--skip-roofline", action="store_true")
    ap.add_argument("--cpu-res", type=int, default=None, help="resolution of the CPU-oracle legs (default: --res)")
    ap.add_argument("--cpu-batch", type=int, default=1)
    ap.add_argument("--cpu-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--ref-max-steps", type=int, default=10)
    ap.add_argument("--delta-only", action="store_true", help="internal: CPU baseline + step-loss delta, printed as JSON")
    ap.add_argument("--gpu-baseline-only", action="store_true", help="internal: the stock-torch-on-B200 comparator leg")
    ap.add_argument("--skip-gpu-baseline", action="store_true")
    ap.add_argument("--delta-timeout", type=int, default=600, help="seconds granted to the --delta-only child process")
    ap.add_argument("--watchdog", type=int, default=1500, help="seconds after which a stuck run dumps its stacks and exits 1")
    args = ap.parse_args()
    if args.cpu_res is None:
        args.cpu_res = args.res
    if args.delta_only:
        run_delta_only(args)
    elif args.gpu_baseline_only:
        run_gpu_baseline_only(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
