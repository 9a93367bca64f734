"""CPU tests of the dense (full-UNet fine-tune, BASELINE config 5) backward's HOST logic: every UNet parameter's
gradient - linear / conv weights and biases, GroupNorm / LayerNorm affine parameters, the separately trained conv1 /
time_emb_proj biases, the timestep MLP - driven through tests/cpu_mock_ops.py and compared with the oracle's autograd
(main.py:143-148: ``unet.requires_grad_(True)``, AdamW over ``unet.parameters()``)."""
import pytest
import torch

from tests import cpu_mock_ops

BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _setup(family, batch, hw, dt=BF, disable_ti=True):
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    from oracle.text import build_text_encoders
    cfg = StepConfig(family=family, tiny=True, resolution=hw * 8, is_lora=False, disable_ti=disable_ti, weight_dtype=dt)
    orc = OracleTrainer(cfg, device="cpu")
    g = torch.Generator().manual_seed(3)
    for n, p in orc.unet.named_parameters():                  # non-trivial norm parameters
        if "norm" in n:
            p.data.add_((torch.randn(p.shape, generator=g) * 0.1).to(p.dtype))
    inputs = make_inputs(cfg, batch=batch, latent_hw=hw, face_mask=True, train_ids=orc.train_ids or None)
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    tes = build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1)
    ti_init = None
    if not disable_ti:
        ti_init = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
                   for te in orc.text_encoders if te is not None]
    tr = TrainerB200(pcfg, {k: v.to(BF) for k, v in orc.unet.state_dict().items()}, tes, device="cpu", ti_init=ti_init)
    return cfg, orc, tr, inputs


@pytest.mark.parametrize("family,batch,hw", [("sdxl", 2, 8), ("sd15", 1, 8)])
def test_dense_step_matches_oracle(monkeypatch, family, batch, hw):
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup(family, batch, hw)
    assert tr.dense_mode and tr.store.n_lora == 0
    # the flat buffer holds every parameter of the oracle UNet, under diffusers names and shapes
    want = {k: tuple(v.shape) for k, v in orc.unet.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in tr.dense.export().items()}
    assert got == want
    assert tr.dense.numel_logical == sum(p.numel() for p in orc.unet.parameters())
    before = tr.dense.export()
    for k, v in orc.unet.state_dict().items():
        assert torch.equal(before[k], v.to(BF)), k
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    for key in ("img_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 2e-3, f"{key}: ours {a} vs oracle {b}"
    assert "token_attention_loss" not in out_p
    grads = tr.dense.export(grads=True)
    worst, n = (0.0, ""), 0
    for name, p in orc.unet.named_parameters():
        assert p.grad is not None, name
        g = grads[name]
        assert g.shape == p.grad.shape, name
        e = rel(g, p.grad)
        worst = max(worst, (e, name))
        n += 1
        # bf16-vs-bf16 backward noise (both sides round every op); the fp32-referenced bound is in the test below
        assert e < 0.3 or float(p.grad.float().abs().max()) < 1e-6, (name, e)
    assert n == len(want)
    orc.optimizer_step()
    tr.optimizer_step()
    after = tr.dense.export()
    moved = 0
    for name, p in orc.unet.named_parameters():
        d_o = p.detach().float() - before[name].float()
        d_p = after[name].float() - before[name].float()
        moved += int((d_o != 0).sum())
        assert float((d_o - d_p).abs().max()) <= 2.5 * float(d_o.abs().max() + 1e-12), name
    assert moved > 0 and float(tr.dense.grads.abs().max()) == 0.0


def test_dense_gradients_against_fp32_autograd(monkeypatch):
    """Same weights evaluated by the oracle in fp32: the executor's dense gradients must be as close to exact arithmetic as
    the oracle's own bf16 backward is (x3 + floor), parameter by parameter."""
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup("sdxl", 2, 8)
    import dataclasses
    from oracle.step import OracleTrainer
    o32 = OracleTrainer(dataclasses.replace(cfg, weight_dtype=torch.float32), device="cpu")
    o32.unet.load_state_dict({k: v.float() for k, v in orc.unet.state_dict().items()})
    for t32, t16 in zip(o32.text_encoders, orc.text_encoders):
        if t32 is not None:
            t32.load_state_dict({k: v.float() for k, v in t16.state_dict().items()})
    o32.step(inputs, do_optimizer=False)
    orc.step(inputs, do_optimizer=False)
    tr.step(inputs, do_optimizer=False)
    grads = tr.dense.export(grads=True)
    g32 = {n: p.grad for n, p in o32.unet.named_parameters()}
    bad = []
    for n, p in orc.unet.named_parameters():
        e_ref, e_ours = rel(p.grad, g32[n]), rel(grads[n], g32[n])
        if e_ours > 3.0 * e_ref + 2e-2:
            bad.append((n, e_ours, e_ref))
    assert not bad, bad[:8]


def test_dense_step_with_textual_inversion(monkeypatch):
    """is_lora=False with the reference's default disable_ti=False: the UNet trains densely AND the token rows train
    (score hook, token-attention and token-std regularisers, CLIP backward); two AdamW launches share one hyper block."""
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup("sdxl", 2, 8, disable_ti=False)
    assert tr.dense_mode and tr.store.n_lora == 0 and tr.store.params.numel() == 3 * (64 + 64)
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    for key in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 2e-3, f"{key}: ours {a} vs oracle {b}"
    grads = tr.dense.export(grads=True)
    for name, p in orc.unet.named_parameters():
        assert rel(grads[name], p.grad) < 0.3 or float(p.grad.float().abs().max()) < 1e-6, name
    off = 0
    for te, rows in zip([t for t in orc.text_encoders if t is not None], tr.ti_rows):
        gref = te.text_model.embeddings.token_embedding.weight.grad[-cfg.n_tokens:]
        assert rel(tr.store.grads[off:off + rows.numel()].view_as(rows), gref) < 0.25
        off += rows.numel()
    rows_before = [r.detach().clone() for r in tr.ti_rows]
    w_before = tr.dense.params.clone()
    orc.optimizer_step()
    tr.optimizer_step()
    assert not torch.equal(tr.dense.params, w_before) and all(not torch.equal(a, b) for a, b in zip(rows_before, tr.ti_rows))
    for te, rows in zip([t for t in orc.text_encoders if t is not None], tr.ti_rows):
        want = te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:]
        # same AdamW step from nearly equal gradients: the updated rows agree to a fraction of the step size (lr = 1e-3)
        assert float((rows.detach().float() - want.float()).abs().max()) <= 2.5e-3
    assert float(tr.dense.grads.abs().max()) == 0.0 and float(tr.store.grads.abs().max()) == 0.0


def test_dense_mode_lr_schedule():
    from sd_lora_trainer_b200.step import StepConfig, lr_schedule
    assert lr_schedule(StepConfig(is_lora=False, disable_ti=True, unet_lr=1e-5), 0, 0.0)[1] == pytest.approx(1e-5)


def test_dense_full_size_sd15_descriptors(monkeypatch):
    """The published SD1.5 graph (859 520 964 parameters) through the dense forward + backward at 128x128: every GEMM /
    norm / im2col descriptor of the full-size layers passes the C wrappers' argument rules (cpu_mock_ops enforces them),
    every parameter receives a gradient.  (SDXL, 2 567 463 684 parameters, passes the same run; it needs ~40 GB of host
    memory, so it is not part of the routine suite.)"""
    cpu_mock_ops.install(monkeypatch)
    from sd_lora_trainer_b200.data import synthetic_inputs
    from sd_lora_trainer_b200.init import random_state_dict
    from sd_lora_trainer_b200.step import StepConfig, TrainerB200
    cfg = StepConfig(family="sd15", resolution=128, is_lora=False, disable_ti=True)
    tr = TrainerB200(cfg, random_state_dict(cfg.arch(), seed=0, device="cpu"), (None, None), device="cpu")
    assert tr.dense.numel_logical == 859_520_964
    inp = synthetic_inputs("sd15", 1, 128, 0, seed=1, face_mask=True, vae_scaling_factor=cfg.arch().vae_scaling_factor)
    st = dict(tr._stage_inputs(inp))
    st["prompt_embeds"] = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(0)).to(BF)
    out = tr._body(st, False)
    assert 0.05 < float(out["tot_loss"]) < 5.0
    for key, (pv, gv, kind, meta) in tr.dense.views.items():
        assert float(gv.abs().max()) > 0.0, key
