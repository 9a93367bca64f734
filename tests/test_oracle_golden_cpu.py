"""Pins the ORACLE: (a) against the committed golden step fixtures (tests/golden/make_golden.py), (b) against
known-answer facts of the third-party algorithms it restates, (c) against the
reference-owned quirks the survey documents (loss.py:165 full-reduction)."""
import math
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


@pytest.mark.parametrize("name", ["step_sdxl_r8_b2.pt", "step_sd15_r4_b1.pt"])
def test_oracle_reproduces_golden_step(name):
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    fix = torch.load(os.path.join(GOLD, name))
    cfg = StepConfig(family=fix["family"], tiny=True, resolution=fix["hw"] * 8, lora_rank=fix["rank"])
    orc = OracleTrainer(cfg, device="cpu")
    g = torch.Generator().manual_seed(7)
    for n, p in orc.unet.named_parameters():
        if "lora_B" in n:
            p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(torch.bfloat16))
    inputs = make_inputs(cfg, batch=fix["batch"], latent_hw=fix["hw"], face_mask=True, train_ids=orc.train_ids)
    out = orc.step(inputs, do_optimizer=False)
    assert torch.equal(out["noisy_latent"].float(), fix["noisy_latent"])          # elementwise bf16: bit-exact
    for k in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        assert abs(float(out[k]) - fix[k]) / abs(fix[k]) < 2e-3, k               # bf16 GEMM summation order
    assert rel(out["model_pred"], fix["model_pred"]) < 2e-2
    assert rel(out["attention_scores"][0], fix["score0"]) < 2e-2
    grads = {n: p.grad.float() for n, p in orc.unet.named_parameters() if p.grad is not None}
    for n, v in fix["grad_norms"].items():
        assert abs(float(grads[n].norm()) - v) / v < 0.1, n


def test_scheduler_known_answers():
    """SD's scaled_linear schedule: published alphas_cumprod end points (0.99915 ... 0.004660)."""
    from oracle.losses import DDPMSchedulerOracle, compute_snr
    s = DDPMSchedulerOracle()
    assert abs(float(s.alphas_cumprod[0]) - 0.99915) < 1e-5
    assert abs(float(s.alphas_cumprod[999]) - 0.004660) < 1e-5
    t = torch.tensor([0, 500, 999])
    snr = compute_snr(s, t)
    a = s.alphas_cumprod[t]
    assert torch.allclose(snr, a / (1 - a), rtol=1e-5)
    x0, eps = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    noisy = s.add_noise(x0, eps, t)
    ref = a.sqrt().view(3, 1, 1, 1) * x0 + (1 - a).sqrt().view(3, 1, 1, 1) * eps
    assert torch.allclose(noisy, ref, atol=1e-6)


def test_architecture_census_matches_survey():
    """Parameter / target / hook counts computed in SURVEY.md Appendix C from the published configs."""
    import torch.nn as nn
    from oracle.lora import lora_target_names
    from oracle.unet import UNet2DConditionModel, UNetConfig, hooked_attention_modules
    want = {"sd15": (859_520_964, 128, 22, 15, 6_414_336), "sdxl": (2_567_463_684, 560, 17, 60, 25_425_920)}
    for name, (params, n_lin, n_conv, n_hook, lora16) in want.items():
        with torch.device("meta"):
            m = UNet2DConditionModel(getattr(UNetConfig, name)())
        assert sum(p.numel() for p in m.parameters()) == params
        names = lora_target_names(m)
        mods = [m.get_submodule(n) for n in names]
        assert sum(isinstance(x, nn.Linear) for x in mods) == n_lin
        assert sum(isinstance(x, nn.Conv2d) for x in mods) == n_conv
        assert len(hooked_attention_modules(m)) == n_hook
        tot = sum(16 * (x.in_features + x.out_features) if isinstance(x, nn.Linear)
                  else 16 * x.in_channels * 9 + x.out_channels * 16 for x in mods)
        assert tot == lora16


def test_loss_mask_quirk_and_weights():
    """loss.py:162-167: with snr_gamma set, the mask multiplies the squared error but the final 'divide by the mask
    mean' is a division by exactly 1.0 (mask.mean(dim=[]) reduces everything)."""
    from oracle.losses import DDPMSchedulerOracle, compute_diffusion_loss
    s = DDPMSchedulerOracle()
    g = torch.Generator().manual_seed(0)
    pred, noise = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    mask = torch.rand(2, 4, 8, 8, generator=g)
    t = torch.tensor([100, 900])
    loss = compute_diffusion_loss(5.0, pred, noise, mask, s, t)
    a = s.alphas_cumprod[t]
    snr = a / (1 - a)
    w = torch.minimum(snr, torch.tensor(5.0)) / snr
    w = w / w.mean()
    want = ((((pred - noise) ** 2) * mask).mean(dim=[1, 2, 3]) * w).mean()
    assert torch.allclose(loss, want, rtol=1e-5)
    half = compute_diffusion_loss(5.0, pred, noise, mask * 0.5, s, t)
    assert torch.allclose(half, want * 0.5, rtol=1e-5)                 # no renormalisation by the mask mean


def test_lora_wrapper_semantics():
    """peft 0.10.0: result = base(x) + lora_B(lora_A(x)) * (alpha / r); gaussian init => B == 0 => identity at step 0."""
    import torch.nn as nn
    from oracle.lora import LoraConv2d, LoraLinear
    torch.manual_seed(0)
    lin = nn.Linear(16, 12)
    lo = LoraLinear(lin, 4, 8.0)
    x = torch.randn(5, 16)
    assert torch.equal(lo(x), lin(x))
    lo.lora_B["default"].weight.data.normal_()
    want = lin(x) + (x @ lo.lora_A["default"].weight.T @ lo.lora_B["default"].weight.T) * 2.0
    assert torch.allclose(lo(x), want, atol=1e-5)
    assert abs(float(lo.lora_A["default"].weight.std()) - 0.25) < 0.06      # std = 1/r
    conv = nn.Conv2d(8, 6, 3, padding=1)
    lc = LoraConv2d(conv, 4, 4.0)
    assert lc.lora_A["default"].weight.shape == (4, 8, 3, 3) and lc.lora_B["default"].weight.shape == (6, 4, 1, 1)


def test_timestep_embedding_layout():
    from oracle.unet import timestep_embedding
    e = timestep_embedding(torch.tensor([0.0, 1.0]), 8)
    assert torch.allclose(e[0], torch.tensor([1.0, 1, 1, 1, 0, 0, 0, 0]))          # [cos | sin] (flip_sin_to_cos)
    assert abs(float(e[1, 0]) - math.cos(1.0)) < 1e-6 and abs(float(e[1, 4]) - math.sin(1.0)) < 1e-6


def test_oracle_reproduces_golden_dense_step():
    """Full-UNet fine-tune branch of the oracle (main.py:143-148) against its committed fixture."""
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    fix = torch.load(os.path.join(GOLD, "dense_step_sd15_b1.pt"))
    cfg = StepConfig(family=fix["family"], tiny=True, resolution=fix["hw"] * 8, is_lora=False, disable_ti=True)
    orc = OracleTrainer(cfg, device="cpu")
    out = orc.step(make_inputs(cfg, batch=fix["batch"], latent_hw=fix["hw"], face_mask=True), do_optimizer=False)
    assert abs(float(out["img_loss"]) - fix["img_loss"]) / fix["img_loss"] < 2e-3
    assert float(out["tot_loss"]) == float(out["img_loss"])                      # no L1 / TI terms without LoRA and TI
    assert rel(out["model_pred"], fix["model_pred"]) < 2e-2
    grads = {n: p.grad.float() for n, p in orc.unet.named_parameters()}
    assert len(grads) == fix["n_params"] and all(p.requires_grad for p in orc.unet.parameters())
    for n, v in fix["grad_norms"].items():
        assert abs(float(grads[n].norm()) - v) / max(v, 1e-12) < 0.1, n


def test_oracle_vae_reproduces_golden_moments():
    from oracle.vae import VAEConfig, build_vae
    fix = torch.load(os.path.join(GOLD, "vae_tiny_moments.pt"))
    out = build_vae(VAEConfig.tiny(), seed=3).encode_moments(fix["image"])
    assert out.shape == fix["moments"].shape == (2, 8, 8, 8)
    assert rel(out, fix["moments"]) < 1e-5                                       # fp32; summation order only
