"""GPU run of the ``train(config)`` generator mirror (sd_lora_trainer_b200/main.py <- reference main.py:34-551) on tiny nets:
real kernels, CUDA-graph replay with a short last batch (re-capture), checkpoints in the reference's file set, and the
LoRA file loading back into a fresh executor.  (Sorts late: written after the round's GPU budget was spent.)"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_latent_sample_matches_diagonal_gaussian():
    """dataset.py:181-193: DiagonalGaussianDistribution.sample() * scaling_factor with the draw injected (fp32)."""
    from sd_lora_trainer_b200.trainer.dataset import CachedLatentDataset
    g = torch.Generator(device="cuda").manual_seed(3)
    params = [torch.randn(1, 8, 16, 16, device="cuda", generator=g) * 3 for _ in range(2)]
    params[1][:, 4:] = params[1][:, 4:] * 20                     # exercises the logvar clamp to [-30, 20]
    ds = CachedLatentDataset(["a", "b"], params, [torch.ones(4, 16, 16, device="cuda")] * 2, 0.13025)
    for i in range(2):
        eps = torch.randn(1, 4, 16, 16, device="cuda", generator=g)
        mean, logvar = torch.chunk(params[i], 2, dim=1)
        ref = (mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps) * 0.13025
        out = ds.sample(i, eps=eps)
        assert out.dtype == torch.float32 and out.shape == ref.shape
        # fp32, same op order; allow a couple of ulps of the LARGER operand (the sum can cancel) for exp() differences
        scale = (mean.abs() + (torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps).abs()) * 0.13025
        assert ((out - ref).abs() <= 1e-6 * scale + 1e-30).all()
    cap, lat, mask = ds[0]
    assert cap == "a" and lat.shape == (4, 16, 16) and mask.shape == (4, 16, 16)


@pytest.mark.parametrize("family,graph", [("sdxl", True), ("sd15", False)])
def test_train_generator_on_gpu(tmp_path, family, graph):
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.main import TrainingConfig, train
    from sd_lora_trainer_b200.trainer.checkpoint import load_lora_weights
    from sd_lora_trainer_b200.trainer.dataset import CachedLatentDataset
    from sd_lora_trainer_b200.unet import UNetB200
    from tests.test_train_loop_cpu import _tokenize_factory
    g = torch.Generator(device="cuda").manual_seed(0)
    n, hw = 5, 16
    ds = CachedLatentDataset([f"a photo of <s0><s1><s2> number {i}" for i in range(n)],
                             [torch.randn(1, 8, hw, hw, device="cuda", generator=g) for _ in range(n)],
                             [torch.ones(4, hw, hw, device="cuda") for _ in range(n)], 0.13025)
    cfg = TrainingConfig(lora_training_urls="gpu/run", concept_mode="face", sd_model_version=family, seed=1, resolution=hw * 8,
                         train_batch_size=2, max_train_steps=8, checkpointing_steps=4, lora_rank=8, output_dir=str(tmp_path),
                         device="cuda:0", caption_dropout=0.3)
    arch = by_name(f"tiny_{family}")
    sd = random_state_dict(arch, seed=0, device="cuda")
    tes = build_text_encoders(family, tiny=True, seed=2)
    gen = train(cfg, ds, tes, sd, _tokenize_factory(family), tiny=True, use_cuda_graph=graph)
    progress = []
    while True:
        try:
            progress.append(next(gen))
        except StopIteration as stop:
            out_cfg, out_dir = stop.value
            break
    torch.cuda.synchronize()
    assert len(progress) >= 8 and progress == sorted(progress) and progress[-1] == 1.0
    hist = out_cfg.training_attributes["losses"]
    assert len(hist["tot_loss"]) == len(progress) and all(v == v and 0 < v < 10 for v in hist["tot_loss"])
    assert len(hist["token_attention_loss"]) == len(progress)
    lora = os.path.join(out_dir, f"run_{family}_lora.safetensors")
    assert os.path.exists(lora) and os.path.exists(os.path.join(out_dir, f"run_{family}_embeddings.safetensors"))
    fresh = UNetB200(arch, sd, 8, device="cuda:0")
    before = fresh.store.params.clone()
    load_lora_weights(lora, fresh)
    assert not torch.equal(fresh.store.params, before)          # trained factors differ from a fresh init
    assert float(fresh.store.params[:fresh.store.n_lora].float().abs().max()) > 0
