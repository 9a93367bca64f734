"""Mirror of the reference's ``main.train(config)`` generator (main.py:34-551) for the accelerated path - the trainer
contract ``predict.py:153-163`` / ``node.py:103-111`` rely on: it yields ``float`` progress while
training and returns ``(config, output_save_dir)``; checkpoints carry the reference's file set (trainer/checkpoint.py).

What stays: the epoch / dataloader structure (shuffled, last batch short), ``completion_f`` and both learning-rate
schedules (main.py:264-291), caption dropout (302-305), the per-step random draws in the reference's order (noise,
offset noise, timesteps: main.py:311-324), gradient accumulation with the ``last_batch`` rule (365-366), the checkpoint
cadence and the final-save rule (403-404, 466-469), the progress arithmetic (458-461).
What changes: the body of the step is ONE call into ``TrainerB200`` (kernels behind the C ABI).
What is not here (SURVEY.md 8 marks them out of scope): preprocessing / captioning, validation renders, debug plots,
text-encoder LoRA, AdamW8bit.  The dataset arrives already
cached (``CachedLatentDataset``, built from the VAE-encode prologue) and captions are turned into token ids by a
caller-supplied ``tokenize`` (no CLIP vocabulary exists offline)."""
from __future__ import annotations

import json
import math
import os
import random
import shutil
import time
from dataclasses import asdict, dataclass, field
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .step import StepConfig, TrainerB200
from .trainer.checkpoint import remove_delimiter_characters, save_checkpoint
from .trainer.embedding_handler import TokenEmbeddingsHandler


@dataclass
class TrainingConfig:
    """The ``trainer/config.py:38-119`` fields the training loop reads, same names and defaults."""
    lora_training_urls: str = ""
    concept_mode: str = "style"
    caption_dropout: float = 0.1
    sd_model_version: Optional[str] = None
    pretrained_model: Optional[dict] = None
    seed: Optional[int] = None
    resolution: int = 512
    train_batch_size: int = 4
    max_train_steps: int = 300
    num_train_epochs: Optional[int] = None
    checkpointing_steps: int = 10000
    gradient_accumulation_steps: int = 1
    is_lora: bool = True
    unet_optimizer_type: str = "adamw"
    unet_lr_warmup_steps: Optional[int] = None
    unet_lr: float = 0.0003
    prodigy_d_coef: float = 1.0
    unet_prodigy_growth_factor: float = 1.05
    lora_weight_decay: float = 0.004
    ti_lr: float = 0.001
    ti_weight_decay: float = 0.0
    ti_optimizer: str = "adamw"
    freeze_ti_after_completion_f: float = 0.7
    freeze_unet_before_completion_f: float = 0.0
    token_attention_loss_w: float = 3e-7
    cond_reg_w: float = 0.0
    tok_cond_reg_w: float = 0.0
    tok_cov_reg_w: float = 0.0
    token_warmup_steps: int = 0
    aspect_ratio_bucketing: bool = False
    l1_penalty: float = 0.03
    noise_offset: float = 0.02
    snr_gamma: float = 5.0
    lora_alpha_multiplier: float = 1.0
    lora_rank: int = 16
    use_dora: bool = False
    name: Optional[str] = None
    output_dir: str = "eden_lora_training_runs"
    debug: bool = False
    disable_ti: bool = False
    weight_type: str = "bf16"
    n_tokens: int = 3
    inserting_list_tokens: List[str] = field(default_factory=lambda: ["<s0>", "<s1>", "<s2>"])
    token_dict: Dict[str, str] = field(default_factory=lambda: {"TOK": "<s0><s1><s2>"})
    device: str = "cuda:0"
    training_attributes: dict = field(default_factory=dict)
    text_encoder_lora_optimizer: Optional[str] = None
    start_time: float = 0.0
    job_time: float = 0.0

    def __post_init__(self):
        """config.py:121-166 (derived fields), minus the timestamped directory name and the GPU picker."""
        if not self.name:
            self.name = os.path.basename(self.lora_training_urls)[:40] or "unnamed"
        self.name = remove_delimiter_characters(self.name)
        if self.seed is None:
            self.seed = int(time.time())
        if self.unet_lr_warmup_steps is None:
            self.unet_lr_warmup_steps = self.max_train_steps
        if self.checkpointing_steps < 1:
            self.checkpointing_steps = self.max_train_steps
        if self.use_dora:
            raise NotImplementedError("DoRA is not part of the accelerated path (no BASELINE config uses it)")
        self.inserting_list_tokens = [f"<s{i}>" for i in range(self.n_tokens)]
        self.token_dict = {"TOK": "".join(self.inserting_list_tokens)}
        if not self.start_time:
            self.start_time = time.time()

    @classmethod
    def from_json(cls, file_path: str) -> "TrainingConfig":
        with open(file_path, "r") as f:
            data = json.load(f)
        known = {k: v for k, v in data.items() if k in cls.__dataclass_fields__}      # other keys drive out-of-scope stages
        return cls(**known)

    def dict(self) -> dict:
        return asdict(self)

    def save_as_json(self, file_path: str) -> None:
        with open(file_path, "w") as f:
            json.dump(self.dict(), f, indent=4)

    def step_config(self, family: str, tiny: bool = False) -> StepConfig:
        return StepConfig(family=family, tiny=tiny, resolution=self.resolution, lora_rank=self.lora_rank, is_lora=self.is_lora,
                          lora_alpha_multiplier=self.lora_alpha_multiplier, lora_weight_decay=self.lora_weight_decay,
                          unet_lr=self.unet_lr, ti_lr=self.ti_lr, ti_weight_decay=self.ti_weight_decay,
                          disable_ti=self.disable_ti, n_tokens=self.n_tokens,
                          token_attention_loss_w=self.token_attention_loss_w, l1_penalty=self.l1_penalty,
                          noise_offset=self.noise_offset, snr_gamma=self.snr_gamma,
                          gradient_accumulation_steps=self.gradient_accumulation_steps,
                          max_train_steps=self.max_train_steps, unet_lr_warmup_steps=self.unet_lr_warmup_steps,
                          freeze_ti_after_completion_f=self.freeze_ti_after_completion_f,
                          freeze_unet_before_completion_f=self.freeze_unet_before_completion_f,
                          unet_optimizer_type=self.unet_optimizer_type, ti_optimizer=self.ti_optimizer,
                          prodigy_d_coef=self.prodigy_d_coef, unet_prodigy_growth_factor=self.unet_prodigy_growth_factor,
                          seed=self.seed)


def seed_everything(seed: int):
    """trainer/utils/utils.py:49-53."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _check_supported(config: TrainingConfig):
    if config.unet_optimizer_type not in ("adamw", "prodigy", "AdamW8bit"):      # AdamW8bit: declared substitution, step.py
        raise NotImplementedError(f"Invalid optimizer_name for unet: {config.unet_optimizer_type}")
    if config.ti_optimizer not in ("adamw", "prodigy"):
        raise NotImplementedError(f"Invalid optimizer_name: '{config.ti_optimizer}'")
    # experimental switches of the reference that are off in its defaults and in every shipped train_config: refuse them
    # loudly rather than train something else (trainer/loss.py:203-225, embedding_handler.py:321-333, dataset.py bucketing)
    for name in ("cond_reg_w", "tok_cond_reg_w", "tok_cov_reg_w", "token_warmup_steps"):
        if getattr(config, name) > 0:
            raise NotImplementedError(f"{name} > 0 is outside the accelerated path (off by default in the reference)")
    if config.aspect_ratio_bucketing:
        raise NotImplementedError("aspect_ratio_bucketing is outside the accelerated path (dataset-side feature)")
    if config.text_encoder_lora_optimizer is not None:
        raise NotImplementedError("text-encoder LoRA is outside the accelerated path (SURVEY.md 2, row 3)")
    if config.weight_type != "bf16":
        raise ValueError(f"the B200 step computes in bf16; got weight_type={config.weight_type}")


def _save(config: TrainingConfig, trainer: TrainerB200, handler: Optional[TokenEmbeddingsHandler], out_dir: str,
          global_step: int, family: str):
    os.makedirs(out_dir, exist_ok=True)
    config.save_as_json(os.path.join(out_dir, "training_args.json"))
    save_checkpoint(output_dir=out_dir, global_step=global_step, unet=trainer.unet, embedding_handler=handler,
                    token_dict=config.token_dict, is_lora=config.is_lora,
                    unet_lora_parameters=[trainer.store.params[:trainer.store.n_lora]] if config.is_lora else None,
                    name=config.name,
                    pretrained_model_version=family, lora_alpha_multiplier=config.lora_alpha_multiplier)


def train(config: TrainingConfig, dataset, text_encoders: Sequence, unet_state_dict: Dict[str, torch.Tensor],
          tokenize: Callable[[List[str]], Tuple[List[torch.Tensor], List[List[int]]]], tiny: bool = False,
          use_cuda_graph: bool = False, process_group=None) -> Iterator[float]:
    """Generator: yields progress in (0, 1], returns ``(config, output_save_dir)`` (main.py:34, 460, 551).

    dataset: ``len()`` / ``[i] -> (caption, vae_latent [4, h, w] fp32 already scaled, mask [4, h, w])``
             (``trainer.dataset.CachedLatentDataset``).
    tokenize: captions -> ([B, 77] int64 ids per text encoder, the unpadded id list of every caption as
              ``pipe.tokenizer.encode`` returns it - trainer/loss.py:33).
    text_encoders: (CLIPTextModel, CLIPTextModelWithProjection | None) with their ORIGINAL embedding tables: ids at or beyond
             the table's row count address the ``n_tokens`` trainable rows (the tokenizer has the new tokens, the encoders are
             not resized)."""
    _check_supported(config)
    # data parallel: one process per GPU, every rank runs this generator; a step's global batch is
    # train_batch_size x world images, the shuffle is shared (seeded by config.seed), the random draws are per rank
    world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
    seed_everything(config.seed + rank)
    perm_gen = torch.Generator().manual_seed(config.seed)
    family = config.sd_model_version or ("sdxl" if "add_embedding.linear_1.weight" in unet_state_dict else "sd15")
    config.sd_model_version = family
    config.pretrained_model = dict(config.pretrained_model or {}, version=family)
    trainer = TrainerB200(config.step_config(family, tiny), unet_state_dict, text_encoders, device=config.device,
                          process_group=process_group, use_cuda_graph=use_cuda_graph)
    # The reference adds the tokens and builds the handler unconditionally (main.py:92-100) and every checkpoint carries
    # {name}_{version}_embeddings.safetensors - its own load_checkpoint indexes that file ([0] over *embeddings.safetensors).
    # With disable_ti the rows are the initialised-but-never-trained ones the UNet was conditioned on (TOK captions).
    handler = TokenEmbeddingsHandler(list(text_encoders))
    handler.inserting_toks = config.inserting_list_tokens
    if config.disable_ti:
        te0 = next(te for te in trainer.text_encoders if te is not None) if any(te is not None for te in trainer.text_encoders) else None
        vocab = te0.text_model.embeddings.token_embedding.weight.shape[0] if te0 is not None else 0
        handler.train_ids, handler.rows = list(range(vocab, vocab + config.n_tokens)), trainer.frozen_rows
    else:
        handler.train_ids, handler.rows = trainer.train_ids, trainer.ti_rows
    dev = trainer.device
    n = len(dataset)
    bs = config.train_batch_size
    gbs = bs * world
    n_batches = int(math.ceil(n / gbs))                        # DataLoader(drop_last=False)
    config.num_train_epochs = int(math.ceil(config.max_train_steps / n_batches))
    checkpoint_dir = os.path.join(str(config.output_dir), "checkpoints")
    if rank == 0:
        if os.path.exists(checkpoint_dir):
            shutil.rmtree(checkpoint_dir)
        os.makedirs(checkpoint_dir)
    global_step, last_save_step = 0, 0
    losses: Dict[str, List[float]] = {"img_loss": [], "tot_loss": [], "token_std_loss": [], "token_attention_loss": []}
    config.training_attributes = dict(config.training_attributes, losses=losses)
    progress_every = max(config.max_train_steps // 100, 1)     # the reference divides by zero below 100 steps
    start_time, images_done = time.time(), 0

    for epoch in range(config.num_train_epochs):
        order = torch.randperm(n, generator=perm_gen).tolist()  # DataLoader(shuffle=True)
        for step in range(n_batches):
            if world == 1:
                idx = order[step * bs:(step + 1) * bs]          # the last batch of an epoch may be short
            else:                                               # every rank needs a full micro-batch: wrap around
                idx = [order[(step * gbs + rank * bs + j) % n] for j in range(bs)]
            completion_f = (epoch + step / n_batches) / config.num_train_epochs
            items = [dataset[i] for i in idx]
            captions = [it[0] for it in items]
            vae_latent = torch.stack([it[1] for it in items]).to(dev, torch.float32)
            mask = torch.stack([it[2] for it in items]).to(dev, torch.float32)
            if config.caption_dropout > 0.0:
                for i in range(len(captions)):
                    if np.random.rand() < config.caption_dropout:
                        captions[i] = config.token_dict["TOK"]
            token_ids, token_indices = tokenize(captions)
            # the step's random draws, in the reference's order and dtype (main.py:311-324)
            noise = torch.randn(vae_latent.shape, device=dev, dtype=torch.bfloat16)
            offset = torch.randn((noise.shape[0], noise.shape[1], 1, 1), device=dev)
            timesteps = torch.randint(0, 1000, (vae_latent.shape[0],), device=dev).long()
            inputs = {"vae_latent": vae_latent, "noise": noise, "offset_noise": offset, "timesteps": timesteps,
                      "mask": mask, "token_ids": [t.to(dev) for t in token_ids], "token_indices": token_indices}
            last_batch = step + 1 == n_batches
            opt_now = (step + 1) % config.gradient_accumulation_steps == 0 or last_batch
            out = trainer.step(inputs, completion_f=completion_f, optimizer_now=opt_now)
            for k in losses:
                if k in out:
                    losses[k].append(float(out[k]))

            if (global_step % config.checkpointing_steps == 0) and (global_step < (config.max_train_steps - 25)):
                output_save_dir = f"{checkpoint_dir}/checkpoint-{global_step}"
                if rank == 0:                                   # replicas are identical: one writer
                    print(f"\n---- avg training fps: {images_done / (time.time() - start_time):.2f}", end="\r", flush=True)
                    _save(config, trainer, handler, output_save_dir, global_step, family)
                last_save_step = global_step
            images_done += gbs
            global_step += 1
            if global_step % progress_every == 0:
                yield float(min(global_step / config.max_train_steps + 0.05, 1.0))
            if global_step > config.max_train_steps:
                print("Reached max steps, stopping training!", flush=True)
                break

    if (global_step - last_save_step) > 26:
        output_save_dir = f"{checkpoint_dir}/checkpoint-{global_step}"
    else:
        output_save_dir = f"{checkpoint_dir}/checkpoint-{last_save_step}"
    if rank == 0:                                              # replicas are identical: one writer
        if not os.path.exists(output_save_dir):
            _save(config, trainer, handler, output_save_dir, global_step, family)
        else:
            print(f"Skipping final save, {output_save_dir} already exists")
    config.job_time = time.time() - config.start_time
    if rank == 0:
        config.save_as_json(os.path.join(output_save_dir, "training_args.json"))
    print("Training job complete, saving outputs...", flush=True)
    trainer.close()                     # captured graphs hold the NCCL kernels of the in-step all-reduce: release them so the
    return config, output_save_dir      # caller can destroy_process_group() (it blocks while such graphs are alive)
