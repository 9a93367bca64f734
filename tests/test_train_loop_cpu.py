"""CPU test of the ``train(config)`` mirror (sd_lora_trainer_b200/main.py <- reference main.py:34-551, the trainer
contract of SURVEY.md 8b): generator protocol, progress values, checkpoint cadence and file set, final-save rule,
gradient accumulation with the last-batch rule - driven through tests/cpu_mock_ops.py on tiny nets."""
import json
import os

import pytest
import torch

from tests import cpu_mock_ops


def _tokenize_factory(family, vocab=128, bos=126, eos=127, n_tokens=3):
    def tokenize(captions):
        ids = torch.full((len(captions), 77), eos, dtype=torch.long)
        lists = []
        for b, cap in enumerate(captions):
            words = [hash(w) % bos for w in cap.replace("<s0><s1><s2>", "").split()]
            seq = [bos] + ([vocab + i for i in range(n_tokens)] if "<s0><s1><s2>" in cap else []) + words + [eos]
            ids[b, :len(seq)] = torch.tensor(seq)
            lists.append(seq)
        return [ids.clone() for _ in range(2 if family == "sdxl" else 1)], lists
    return tokenize


def _dataset(n, hw, tok="<s0><s1><s2>"):
    from sd_lora_trainer_b200.trainer.dataset import CachedLatentDataset
    g = torch.Generator().manual_seed(0)
    params = [torch.randn(1, 8, hw, hw, generator=g) for _ in range(n)]
    masks = [torch.ones(4, hw, hw) for _ in range(n)]
    caps = [f"a photo of {tok} number {i}" for i in range(n)]
    return CachedLatentDataset(caps, params, masks, 0.13025)


@pytest.mark.parametrize("family", ["sdxl", "sd15"])
def test_train_generator_contract(monkeypatch, tmp_path, family):
    cpu_mock_ops.install(monkeypatch)
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.main import TrainingConfig, train
    cfg = TrainingConfig(lora_training_urls="unit/test concept", concept_mode="face", sd_model_version=family, seed=1,
                         resolution=64, train_batch_size=2, max_train_steps=6, checkpointing_steps=2,
                         gradient_accumulation_steps=2, lora_rank=4, output_dir=str(tmp_path), device="cpu",
                         caption_dropout=0.5)
    assert cfg.name == "test_concept" and cfg.unet_lr_warmup_steps == 6 and cfg.token_dict == {"TOK": "<s0><s1><s2>"}
    sd = random_state_dict(by_name(f"tiny_{family}"), seed=0, device="cpu")
    tes = build_text_encoders(family, tiny=True, seed=2)
    gen = train(cfg, _dataset(3, 8), tes, sd, _tokenize_factory(family), tiny=True)
    progress = []
    while True:
        try:
            progress.append(next(gen))
        except StopIteration as stop:
            out_cfg, out_dir = stop.value
            break
    # 3 images / batch 2 -> 2 batches per epoch, 3 epochs for 6 steps; one progress value per step (6 // 100 -> every step)
    assert out_cfg.num_train_epochs == 3 and len(progress) == 6
    assert all(isinstance(p, float) for p in progress) and progress == sorted(progress) and progress[-1] == 1.0
    assert abs(progress[0] - (1 / 6 + 0.05)) < 1e-9
    # checkpoint cadence: global_step % 2 == 0 and global_step < 6 - 25 is never true -> only the final save, whose
    # directory is checkpoint-0 (global_step - last_save_step = 6 <= 26, main.py:466-469)
    ck = os.path.join(str(tmp_path), "checkpoints")
    assert sorted(os.listdir(ck)) == ["checkpoint-0"] and out_dir == f"{ck}/checkpoint-0"
    files = sorted(os.listdir(out_dir))
    name = "test_concept"
    assert files == sorted(["adapter_config.json", "special_params.json", "training_args.json",
                            f"{name}_{family}_embeddings.safetensors", f"{name}_{family}_lora.safetensors"])
    args = json.load(open(os.path.join(out_dir, "training_args.json")))
    assert args["sd_model_version"] == family and args["job_time"] > 0 and args["pretrained_model"]["version"] == family
    hist = args["training_attributes"]["losses"]
    assert len(hist["tot_loss"]) == 6 and all(v == v and v > 0 for v in hist["tot_loss"])
    assert json.load(open(os.path.join(out_dir, "special_params.json"))) == {"TOK": "<s0><s1><s2>"}
    from safetensors.torch import load_file
    emb = load_file(os.path.join(out_dir, f"{name}_{family}_embeddings.safetensors"))
    assert set(emb) == ({"clip_l", "clip_g"} if family == "sdxl" else {"clip_l"}) and emb["clip_l"].shape[0] == 3


def test_train_checkpoint_cadence_and_accumulation(monkeypatch, tmp_path):
    """40 steps, checkpoint every 5: saves at 0, 5, 10 (global_step < 40 - 25), the loop leaves once global_step > 40, i.e. after 41 steps
    (main.py:463-465), so the final save is checkpoint-41 (41 - 10 > 26).
    Gradient accumulation 2 over 3 batches per epoch: the optimizer runs on batches 2 and 3 (last-batch rule)."""
    cpu_mock_ops.install(monkeypatch)
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200 import main as main_mod
    calls = []
    real_step = main_mod.TrainerB200.step

    def spy(self, inputs, completion_f=0.0, do_optimizer=True, optimizer_now=None):
        calls.append((round(completion_f, 6), optimizer_now, inputs["vae_latent"].shape[0]))
        # keep the unit test fast: run the real step only for the first epoch
        if len(calls) <= 3:
            return real_step(self, inputs, completion_f, do_optimizer, optimizer_now)
        self.global_step += 1
        return {"tot_loss": torch.tensor(1.0), "img_loss": torch.tensor(1.0)}

    monkeypatch.setattr(main_mod.TrainerB200, "step", spy)
    cfg = main_mod.TrainingConfig(lora_training_urls="cadence", sd_model_version="sd15", seed=3, resolution=64,
                                  train_batch_size=2, max_train_steps=40, checkpointing_steps=5,
                                  gradient_accumulation_steps=2, lora_rank=4, output_dir=str(tmp_path), device="cpu",
                                  caption_dropout=0.0, disable_ti=True)
    sd = random_state_dict(by_name("tiny_sd15"), seed=0, device="cpu")
    gen = main_mod.train(cfg, _dataset(5, 8, tok="something"), build_text_encoders("sd15", tiny=True, seed=2), sd, _tokenize_factory("sd15"),
                         tiny=True)
    out = None
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        out = stop.value
    assert out[1].endswith("checkpoint-41") and len(calls) == 41
    assert sorted(os.listdir(os.path.join(str(tmp_path), "checkpoints")), key=lambda s: int(s.split("-")[1])) == \
        ["checkpoint-0", "checkpoint-5", "checkpoint-10", "checkpoint-41"]
    # 5 images / batch 2 -> 3 batches (2, 2, 1); 14 epochs x 3 = 42 >= 40 steps; the loop leaves at global_step 41 > 40
    assert [c[2] for c in calls[:3]] == [2, 2, 1]
    assert [c[1] for c in calls[:6]] == [False, True, True, False, True, True]
    assert calls[0][0] == 0.0 and abs(calls[4][0] - (1 + 1 / 3) / 14) < 1e-6
    # disable_ti still writes the (never-trained) token rows, as the reference does (main.py:92-100; its load_checkpoint
    # indexes *embeddings.safetensors unconditionally)
    from safetensors.torch import load_file
    emb = load_file(os.path.join(out[1], "cadence_sd15_embeddings.safetensors"))
    assert set(emb) == {"clip_l"} and emb["clip_l"].shape[0] == cfg.n_tokens


def test_train_rejects_unsupported_modes(tmp_path):
    from sd_lora_trainer_b200.main import TrainingConfig, train
    for kw, exc in ((dict(unet_optimizer_type="lion"), NotImplementedError), (dict(ti_optimizer="adagrad"), NotImplementedError),
                    (dict(tok_cov_reg_w=1e-3), NotImplementedError), (dict(aspect_ratio_bucketing=True), NotImplementedError), (dict(weight_type="fp16"), ValueError),
                    (dict(text_encoder_lora_optimizer="adamw"), NotImplementedError)):
        cfg = TrainingConfig(lora_training_urls="x", output_dir=str(tmp_path), device="cpu", **kw)
        with pytest.raises(exc):
            next(train(cfg, [], (None, None), {}, lambda c: ([], [])))
    with pytest.raises(NotImplementedError):
        TrainingConfig(lora_training_urls="x", use_dora=True)


def test_training_config_json_roundtrip(tmp_path):
    from sd_lora_trainer_b200.main import TrainingConfig
    p = tmp_path / "args.json"
    # a reference config file carries keys of out-of-scope stages too (train_configs/*.json): they are ignored
    json.dump({"lora_training_urls": "https://x/y.zip", "concept_mode": "face", "sd_model_version": "sdxl", "max_train_steps": 300,
               "lora_rank": 16, "n_tokens": 2, "caption_model": "florence", "n_sample_imgs": 4, "checkpointing_steps": 0}, open(p, "w"))
    cfg = TrainingConfig.from_json(str(p))
    assert cfg.inserting_list_tokens == ["<s0>", "<s1>"] and cfg.token_dict == {"TOK": "<s0><s1>"}
    assert cfg.checkpointing_steps == 300 and cfg.name == "y.zip"
    cfg.save_as_json(str(p))
    assert json.load(open(p))["lora_rank"] == 16


def test_train_full_finetune_like_the_reference_example(monkeypatch, tmp_path):
    """train_configs/full_finetuning_example.json in miniature: is_lora=false, disable_ti=true, AdamW8bit (a declared
    substitution: AdamW with bf16 moments, announced by a warning); the checkpoint is a diffusers UNet folder."""
    cpu_mock_ops.install(monkeypatch)
    from safetensors.torch import load_file
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.main import TrainingConfig, train
    cfg = TrainingConfig(lora_training_urls="ft/stitchly", concept_mode="style", sd_model_version="sd15", seed=1, resolution=64,
                         train_batch_size=2, max_train_steps=3, is_lora=False, disable_ti=True, unet_optimizer_type="AdamW8bit",
                         unet_lr=1e-4, lora_rank=4, output_dir=str(tmp_path), device="cpu", caption_dropout=0.2)
    sd = random_state_dict(by_name("tiny_sd15"), seed=0, device="cpu")
    gen = train(cfg, _dataset(4, 8, tok="stitchly"), build_text_encoders("sd15", tiny=True, seed=2), sd, _tokenize_factory("sd15"),
                tiny=True)
    with pytest.warns(UserWarning, match="AdamW8bit"):
        try:
            while True:
                next(gen)
        except StopIteration as stop:
            out_cfg, out_dir = stop.value
    assert sorted(os.listdir(out_dir)) == ["config.json", "diffusion_pytorch_model.safetensors", "special_params.json",
                                           "stitchly_sd15_embeddings.safetensors", "training_args.json"]
    trained = load_file(os.path.join(out_dir, "diffusion_pytorch_model.safetensors"))
    assert set(trained) == set(sd) and all(trained[k].shape == sd[k].shape for k in sd)
    changed = sum(int(not torch.equal(trained[k], sd[k])) for k in sd)
    # most tensors moved; bf16 norm weights sitting at 1.0 (ulp 0.0078) swallow steps of 1e-4, in the reference too
    assert changed > 0.7 * len(sd)
    hist = json.load(open(os.path.join(out_dir, "training_args.json")))["training_attributes"]["losses"]
    assert len(hist["tot_loss"]) == 4 and hist["token_attention_loss"] == [] and hist["tot_loss"] == hist["img_loss"]
