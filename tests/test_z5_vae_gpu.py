"""GPU parity of the VAE-encode prologue (trainer/dataset.py:141-179): the bf16 kernel path against
the fp32 oracle restatement of AutoencoderKL.encode on the same weights and images.  (File name sorts last on purpose:
this widening was written after the round's GPU budget was spent - first executed by the round-end GPU run.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _pair(cfg, seed):
    from oracle.vae import build_vae, state_dict_of
    from sd_lora_trainer_b200.vae import VAEEncoderB200
    orc = build_vae(cfg, seed=seed).cuda()
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    for n, p in orc.named_parameters():
        if "norm" in n:
            p.data.add_(torch.randn(p.shape, generator=g, device="cuda") * 0.2)
    enc = VAEEncoderB200(state_dict_of(orc), device="cuda:0", block_out_channels=cfg.block_out_channels,
                         layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
    return orc, enc, g


def _check(orc, enc, img):
    ref = orc.encode_moments(img)
    out = enc.encode_moments(img)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == torch.float32 and torch.isfinite(out).all()
    with torch.autocast("cuda", dtype=BF):
        noise = rel(orc.encode_moments(img), ref)                # what torch's own bf16 path loses on this graph
    err = rel(out, ref)
    assert err < max(3e-2, 2.0 * noise), (err, noise)
    return err


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 32, 96), (3, 16, 16)])
def test_vae_tiny_matches_oracle(B, H, W):
    from oracle.vae import VAEConfig
    orc, enc, g = _pair(VAEConfig.tiny(), 3)
    _check(orc, enc, torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1)


@pytest.mark.parametrize("B,S", [(2, 128), (1, 256), (1, 512)])
def test_vae_full_graph_matches_oracle(B, S):
    """Published SD / SDXL VAE graph (128-256-512-512): implicit convs (S=128), im2col for wider maps (256, 512)."""
    from oracle.vae import VAEConfig
    orc, enc, g = _pair(VAEConfig(), 5)
    _check(orc, enc, torch.rand(B, 3, S, S, device="cuda", generator=g) * 2 - 1)


def test_vae_full_size_1024_and_dataset_cache():
    """BASELINE's 1024x1024: 1 Mpixel GEMM rows, a 16384-token single-head attention, then the cached-posterior
    dataset (dataset.py:157-158, 181-193) built from the encoder's output."""
    from oracle.vae import VAEConfig, diagonal_gaussian_sample
    from sd_lora_trainer_b200.trainer.dataset import CachedLatentDataset
    orc, enc, g = _pair(VAEConfig(), 7)
    img = torch.rand(1, 3, 1024, 1024, device="cuda", generator=g) * 2 - 1
    _check(orc, enc, img)
    ds = CachedLatentDataset.from_images(enc, ["a photo of <s0><s1><s2>"], [img[0]], None, 0.13025)
    eps = torch.randn(1, 4, 128, 128, device="cuda", generator=g)
    x0 = ds.sample(0, eps=eps)
    want = diagonal_gaussian_sample(ds.params[0], eps) * 0.13025
    assert torch.allclose(x0, want, rtol=1e-5, atol=1e-6)
    cap, lat, mask = ds[0]
    assert lat.shape == (4, 128, 128) and mask.shape == (4, 128, 128) and float(mask.min()) == 1.0
