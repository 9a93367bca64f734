"""Data-parallel semantics on CPU (gloo, world_size 2): each rank steps on its own micro-batch, ONE all-reduce sums the
flat fp32 gradient buffer, both ranks apply the same fused AdamW.  Must equal a single process doing gradient
accumulation over the same two micro-batches (main.py:362-366), and both ranks must stay in lock-step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

BF = torch.bfloat16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _patch_ops():
    import sys
    from tests import cpu_mock_ops
    import sd_lora_trainer_b200.step as step_mod
    import sd_lora_trainer_b200.unet as unet_mod
    import sd_lora_trainer_b200.trainer.loss as loss_mod
    for mod in (step_mod, unet_mod, loss_mod):
        mod.ops = cpu_mock_ops


def _make(world_ga, pg=None, ga=1):
    from oracle.step import OracleTrainer, StepConfig as OCfg
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.step import StepConfig, TrainerB200
    cfg = StepConfig(family="sd15", tiny=True, resolution=64, lora_rank=8, gradient_accumulation_steps=ga)
    sd = random_state_dict(cfg.arch(), seed=0, device="cpu")
    tes = build_text_encoders("sd15", True, seed=1)
    tr = TrainerB200(cfg, sd, tes, device="cpu", process_group=pg)
    g = torch.Generator().manual_seed(3)
    for s in tr.store.slots:                               # non-zero B so every gradient path is live
        s.B()[:, :s.r].copy_((torch.randn(s.fan_out, s.r, generator=g) * 0.05).to(BF))
    return cfg, tr


def _inputs(cfg, seed):
    from sd_lora_trainer_b200.data import synthetic_inputs
    return synthetic_inputs("sd15", 1, cfg.resolution, cfg.n_tokens, seed=seed, face_mask=True,
                            vae_scaling_factor=0.18215, tiny=True)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_ops()
    cfg, tr = _make(world, pg=dist.group.WORLD)
    out = tr.step(_inputs(cfg, 100 + rank))
    torch.save({"params": tr.store.params.clone(), "loss": float(out["tot_loss"])}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_equals_gradient_accumulation(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["params"], r1["params"]), "ranks diverged"
    # single process, gradient_accumulation_steps = 2 over the same two micro-batches
    _patch_ops()
    cfg, tr = _make(1, ga=2)
    p0 = tr.store.params.clone()
    tr.step(_inputs(cfg, 100))
    assert torch.equal(tr.store.params, p0), "no update before the accumulation window closes"
    tr.step(_inputs(cfg, 101))
    moved = (tr.store.params != p0)
    assert int(moved.sum()) > 0
    # same fp32 gradient sum up to summation order -> identical bf16 update except for rare rounding ties
    diff = (tr.store.params.float() - r0["params"].float()).abs()
    frac_diff = float((diff > 0).float().mean())
    assert frac_diff < 5e-3, frac_diff
    assert float(diff.max()) <= 2.5 * float((tr.store.params.float() - p0.float()).abs().max())


def _train_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_ops()
    import sd_lora_trainer_b200.trainer.dataset as ds_mod
    from tests import cpu_mock_ops
    ds_mod.ops = cpu_mock_ops
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200 import main as main_mod
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from tests.test_train_loop_cpu import _dataset, _tokenize_factory
    seen = []
    real_step = main_mod.TrainerB200.step

    holder = {}

    def spy(self, inputs, completion_f=0.0, do_optimizer=True, optimizer_now=None):
        seen.append(inputs["vae_latent"].clone())
        holder["trainer"] = self
        return real_step(self, inputs, completion_f, do_optimizer, optimizer_now)

    main_mod.TrainerB200.step = spy
    cfg = main_mod.TrainingConfig(lora_training_urls="dp", sd_model_version="sd15", seed=5, resolution=64, train_batch_size=1,
                                  max_train_steps=3, lora_rank=4, output_dir=os.path.join(out_dir, "run"), device="cpu",
                                  caption_dropout=0.0)
    gen = main_mod.train(cfg, _dataset(5, 8), build_text_encoders("sd15", tiny=True, seed=2),
                         random_state_dict(by_name("tiny_sd15"), seed=0, device="cpu"), _tokenize_factory("sd15"), tiny=True,
                         process_group=dist.group.WORLD)
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        out_cfg, out_path = stop.value
    torch.save({"latents": torch.stack(seen), "out_path": out_path, "done": gen.gi_frame is None,
                "params": holder["trainer"].store.params.clone()}, os.path.join(out_dir, f"t{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_train_generator_shards_batches_across_ranks(tmp_path):
    """train() under data parallelism: the ranks share the shuffle but take different images of every global batch, stay
    in lock-step through the in-step all-reduce, and only rank 0 writes the checkpoint."""
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    t0, t1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    assert t0["done"] and t1["done"] and t0["out_path"] == t1["out_path"]
    assert torch.equal(t0["params"], t1["params"]), "replicas diverged"
    assert t0["latents"].shape == t1["latents"].shape and t0["latents"].shape[0] >= 3
    for a, b in zip(t0["latents"], t1["latents"]):
        assert not torch.equal(a, b)                       # different images (and different posterior draws) per rank
    files = sorted(os.listdir(t0["out_path"]))
    assert "dp_sd15_lora.safetensors" in files and "training_args.json" in files
    assert sorted(os.listdir(os.path.join(tmp_path, "run", "checkpoints"))) == [os.path.basename(t0["out_path"])]
