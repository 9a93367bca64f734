"""CPU tests of the VAE-encode prologue: the oracle restatement of AutoencoderKL.encode, and the
product executor's HOST logic (mirrored coordinates, flipped taps, folded quant_conv, operand descriptors of the
single-head attention) driven through tests/cpu_mock_ops.py against that oracle."""
import pytest
import torch

from tests import cpu_mock_ops

BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _install(monkeypatch):
    import sys
    import sd_lora_trainer_b200.unet as unet_mod
    import sd_lora_trainer_b200.vae as vae_mod
    me = sys.modules[cpu_mock_ops.__name__]
    for mod in (unet_mod, vae_mod):
        monkeypatch.setattr(mod, "ops", me)


def test_oracle_vae_graph_counts():
    """Parameter count of the published SD / SDXL VAE encoder + quant_conv (34 163 592 + 72)."""
    from oracle.vae import AutoencoderKLEncoder, VAEConfig
    with torch.device("meta"):
        m = AutoencoderKLEncoder(VAEConfig())
    n_enc = sum(p.numel() for p in m.encoder.parameters())
    n_q = sum(p.numel() for p in m.quant_conv.parameters())
    assert n_enc == 34_163_592 and n_q == 72
    names = set(m.state_dict())
    for k in ("encoder.conv_in.weight", "encoder.down_blocks.1.resnets.0.conv_shortcut.weight",
              "encoder.down_blocks.2.downsamplers.0.conv.bias", "encoder.mid_block.attentions.0.group_norm.weight",
              "encoder.mid_block.attentions.0.to_out.0.bias", "encoder.conv_norm_out.bias", "quant_conv.weight"):
        assert k in names, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in names
    assert "encoder.down_blocks.0.resnets.0.conv_shortcut.weight" not in names


def test_oracle_downsample_is_bottom_right_padded():
    from oracle.vae import Downsample
    d = Downsample(1)
    d.conv.weight.data.fill_(1.0)
    d.conv.bias.data.zero_()
    x = torch.arange(16.0).view(1, 1, 4, 4)
    y = d(x)
    assert y.shape == (1, 1, 2, 2)
    assert float(y[0, 0, 0, 0]) == float(x[0, 0, 0:3, 0:3].sum())            # no top/left padding
    assert float(y[0, 0, 1, 1]) == float(x[0, 0, 2:4, 2:4].sum())            # bottom/right zero padding


@pytest.mark.parametrize("B,H,W,full", [(2, 32, 32, False), (1, 16, 48, False), (1, 32, 32, True), (1, 256, 256, True)])
def test_vae_encoder_host_logic_matches_oracle(monkeypatch, B, H, W, full):
    _install(monkeypatch)
    from oracle.vae import VAEConfig, build_vae, state_dict_of
    from sd_lora_trainer_b200.vae import VAEEncoderB200
    cfg = VAEConfig() if full else VAEConfig.tiny()             # full = the published SD / SDXL VAE graph
    orc = build_vae(cfg, seed=3)
    g = torch.Generator().manual_seed(11)
    for n, p in orc.named_parameters():                      # non-trivial norms / biases
        if "norm" in n:
            p.data.add_(torch.randn(p.shape, generator=g) * 0.2)
    img = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    ref = orc.encode_moments(img)
    enc = VAEEncoderB200(state_dict_of(orc), device="cpu", block_out_channels=cfg.block_out_channels,
                         layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
    out = enc.encode_moments(img)
    f = 1 << (len(cfg.block_out_channels) - 1)
    assert out.shape == ref.shape == (B, 8, H // f, W // f) and out.dtype == torch.float32
    # bf16 activations vs the fp32 oracle; the bf16 oracle itself sits at the same distance
    bf = build_vae(cfg, seed=3, dtype=BF)
    bf.load_state_dict(orc.state_dict())
    noise = rel(bf.to(BF).encode_moments(img.to(BF)), ref)
    err = rel(out, ref)
    assert err < max(3e-2, 2.0 * noise), (err, noise)
    # a wrong flip / padding side would be an O(1) error: check the un-mirrored version is far off
    assert rel(out.flip(2, 3), ref) > 10 * err


def test_vae_attention_chunks_over_images(monkeypatch):
    """max_score_bytes forces one image per score buffer; results must not change."""
    _install(monkeypatch)
    from oracle.vae import VAEConfig, build_vae, state_dict_of
    from sd_lora_trainer_b200.vae import VAEEncoderB200
    cfg = VAEConfig.tiny()
    orc = build_vae(cfg, seed=5)
    img = torch.rand(3, 3, 16, 16, generator=torch.Generator().manual_seed(2)) * 2 - 1
    kw = dict(device="cpu", block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
              norm_num_groups=cfg.norm_num_groups)
    a = VAEEncoderB200(state_dict_of(orc), **kw).encode_moments(img)
    b = VAEEncoderB200(state_dict_of(orc), max_score_bytes=1, **kw).encode_moments(img)
    assert torch.equal(a, b)


def test_vae_encode_returns_scaled_sample(monkeypatch):
    _install(monkeypatch)
    from oracle.vae import VAEConfig, build_vae, diagonal_gaussian_sample, state_dict_of
    from sd_lora_trainer_b200.vae import VAEEncoderB200
    cfg = VAEConfig.tiny()
    orc = build_vae(cfg, seed=9)
    img = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(4)) * 2 - 1
    enc = VAEEncoderB200(state_dict_of(orc), device="cpu", block_out_channels=cfg.block_out_channels,
                         layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
    eps = torch.randn(1, 4, 4, 4, generator=torch.Generator().manual_seed(5))
    params, x0 = enc.encode(img, eps=eps, scaling_factor=0.13025)
    want = diagonal_gaussian_sample(params, eps) * 0.13025
    assert torch.allclose(x0, want, rtol=1e-6, atol=1e-7)


def test_cached_dataset_from_images(monkeypatch):
    """dataset.py:141-179: posterior cached per image, mask nearest-resized to the latent grid and repeated over channels."""
    _install(monkeypatch)
    import sd_lora_trainer_b200.trainer.dataset as ds_mod
    monkeypatch.setattr(ds_mod, "ops", cpu_mock_ops)
    from oracle.vae import VAEConfig, build_vae, state_dict_of
    from sd_lora_trainer_b200.vae import VAEEncoderB200
    cfg = VAEConfig.tiny()
    orc = build_vae(cfg, seed=1)
    enc = VAEEncoderB200(state_dict_of(orc), device="cpu", block_out_channels=cfg.block_out_channels,
                         layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
    g = torch.Generator().manual_seed(8)
    imgs = [torch.rand(3, 16, 16, generator=g) * 2 - 1 for _ in range(2)]
    mask = (torch.rand(1, 16, 16, generator=g) > 0.5).float()
    ds = ds_mod.CachedLatentDataset.from_images(enc, ["a", "b"], imgs, [mask, None], 0.18215)
    assert len(ds) == 2 and ds.params[0].shape == (1, 8, 4, 4)
    want = torch.nn.functional.interpolate(mask[None], size=(4, 4), mode="nearest").repeat(1, 4, 1, 1)[0]
    assert torch.equal(ds.masks[0], want) and float(ds.masks[1].min()) == 1.0 and ds.masks[1].shape == (4, 4, 4)
    cap, lat, m = ds[1]
    assert cap == "b" and lat.shape == (4, 4, 4) and m.shape == (4, 4, 4)


@pytest.mark.parametrize("tiny", [False, True])
def test_vae_param_shapes_match_oracle_state_dict(tiny):
    """The product's own enumeration of the VAE keys (random init for benches) names and shapes every oracle parameter."""
    from oracle.vae import AutoencoderKLEncoder, VAEConfig
    from sd_lora_trainer_b200.init import random_vae_encoder_state_dict, vae_encoder_param_shapes
    cfg = VAEConfig.tiny() if tiny else VAEConfig()
    with torch.device("meta"):
        ref = AutoencoderKLEncoder(cfg)
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    kw = dict(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block)
    assert dict(vae_encoder_param_shapes(**kw)) == want
    if tiny:
        sd = random_vae_encoder_state_dict(seed=1, **kw)
        assert float(sd["encoder.conv_norm_out.weight"].min()) == 1.0 and float(sd["encoder.conv_in.weight"].abs().max()) <= 27 ** -0.5


def test_prepare_captions_like_the_reference_dataset():
    """dataset.py:46-52: lower-case, TOK substitution (key lower-cased), missing caption -> ''."""
    from sd_lora_trainer_b200.trainer.dataset import prepare_captions
    got = prepare_captions(["A photo of TOK, smiling", "tok and TOK", None, float("nan"), "No Token Here"], {"TOK": "<s0><s1><s2>"})
    # the replacement is a plain substring replace of the lower-cased key, so "token" is hit too - as in the reference
    assert got == ["a photo of <s0><s1><s2>, smiling", "<s0><s1><s2> and <s0><s1><s2>", "", "", "no <s0><s1><s2>en here"]
    assert prepare_captions(["Keep CASE?"], None) == ["keep case?"]
