"""CPU test of the reference-facing mirrors in call order (main.py:36-52, 134-176, 242-262): load_models ->
init_daam_loss -> TokenEmbeddingsHandler -> get_unet_lora_parameters -> get_*_optimizer -> OptimizerCollection, for the
AdamW and the Prodigy handles, through tests/cpu_mock_ops.py."""
import pytest
import torch

from tests import cpu_mock_ops

BF = torch.bfloat16


@pytest.mark.parametrize("unet_opt,ti_opt", [("adamw", "adamw"), ("prodigy", "adamw"), ("prodigy", "prodigy")])
def test_reference_call_sequence(monkeypatch, unet_opt, ti_opt):
    cpu_mock_ops.install(monkeypatch)
    import sd_lora_trainer_b200.trainer.optimizer as opt_mod
    monkeypatch.setattr(opt_mod, "ops", cpu_mock_ops)
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.trainer.embedding_handler import TokenEmbeddingsHandler
    from sd_lora_trainer_b200.trainer.models import load_models
    from sd_lora_trainer_b200.trainer.optimizer import (FlatProdigy, OptimizerCollection, get_textual_inversion_optimizer,
                                                        get_unet_lora_parameters, get_unet_optimizer)
    from sd_lora_trainer_b200.trainer.ti_cross_attn_loss import init_daam_loss
    tes = build_text_encoders("sdxl", tiny=True, seed=1)
    (pipe, tok1, tok2, sched, te1, te2, vae, unet), version = load_models({"random_init": "tiny_sdxl"}, "cpu", BF, text_encoders=tes)
    assert version == "sdxl" and unet is None and vae is None and sched.config.num_train_timesteps == 1000
    n_tok, dims = 3, [64, 64]
    unet, groups, lora_params = get_unet_lora_parameters(4, 1.0, 0.004, False, unet, pipe, ti_elems=n_tok * sum(dims))
    assert pipe.unet is unet and groups[0]["weight_decay"] == 0.004 and lora_params[0].numel() == unet.store.n_lora
    pipe, daam = init_daam_loss(pipe)
    assert all(a.capture for a in unet.hooked) and len(daam.layer_names) == len(unet.hooked)
    handler = TokenEmbeddingsHandler([te1, te2])
    handler.initialize_new_tokens(["<s0>", "<s1>", "<s2>"], seed=0, store=unet.store)
    assert handler.train_ids == [128, 129, 130] and len(handler.rows) == 2
    o_unet = get_unet_optimizer(1.0, 1.05, 0.004, False, lora_params, optimizer_name=unet_opt, unet=unet)
    o_ti, ti_params = get_textual_inversion_optimizer([te1, te2], 1e-3, 0.0, ti_opt, unet=unet)
    assert ti_params[0].numel() == n_tok * sum(dims)
    assert isinstance(o_unet, FlatProdigy) == (unet_opt == "prodigy") and isinstance(o_ti, FlatProdigy) == (ti_opt == "prodigy")
    coll = OptimizerCollection(optimizer_textual_inversion=o_ti, optimizer_unet=o_unet, l1_penalty=0.03)
    coll.optimizers["unet"].param_groups[0]["lr"] = 3e-4 if unet_opt == "adamw" else 1.0            # main.py:288
    if ti_opt != "prodigy":
        coll.optimizers["textual_inversion"].param_groups[0]["lr"] = 1e-3                          # main.py:271
    g = torch.Generator().manual_seed(0)
    before = unet.store.params.clone()
    for _ in range(2):
        unet.store.grads.copy_(torch.randn(unet.store.grads.shape, generator=g) * 1e-3)
        coll.step()
        coll.zero_grad()
        assert float(unet.store.grads.abs().max()) == 0.0                                          # consumed by the kernel
    nl = unet.store.n_lora
    assert not torch.equal(unet.store.params[:nl], before[:nl])
    if ti_opt == "adamw":
        assert not torch.equal(unet.store.params[nl:], before[nl:])
    assert coll.get_lr("unet") == (3e-4 if unet_opt == "adamw" else 1.0)
    with pytest.raises(NotImplementedError):
        get_unet_optimizer(1.0, 1.05, 0.004, False, lora_params, optimizer_name="lion", unet=unet)
    with pytest.warns(UserWarning, match="AdamW8bit"):                                              # declared substitution
        assert not isinstance(get_unet_optimizer(1.0, 1.05, 0.004, False, lora_params, optimizer_name="AdamW8bit", unet=unet),
                              FlatProdigy)
    with pytest.raises(NotImplementedError):
        get_unet_lora_parameters(4, 1.0, 0.004, True, None, pipe)                                   # DoRA


def test_token_attention_loss_without_ti_tokens_has_finite_zero_gradients():
    """trainer/loss.py:55-56: when no caption of the batch holds every trainable token the reference returns a grad-less
    0.0; the tensorised form must give loss 0 and exactly-zero (not NaN) gradients."""
    import torch
    from sd_lora_trainer_b200.trainer.loss import token_attention_loss_from_maps, token_index_tensors
    maps = torch.randn(3, 2, 4, 4, 77, dtype=torch.bfloat16).requires_grad_(True)
    masks = torch.rand(2, 4, 8, 8)
    tok_len, ti_pos = token_index_tensors([[1, 5, 6, 2], [1, 7, 2]], train_ids=[100, 101, 102])
    assert int((ti_pos >= 0).sum()) == 0
    loss = token_attention_loss_from_maps(maps, masks, tok_len, ti_pos)
    loss.backward()
    assert float(loss) == 0.0 and torch.isfinite(maps.grad).all() and float(maps.grad.abs().max()) == 0.0
    # one caption with the tokens, one without: finite, non-zero
    maps2 = maps.detach().clone().requires_grad_(True)
    tok_len, ti_pos = token_index_tensors([[1, 100, 101, 102, 6, 2], [1, 7, 2]], train_ids=[100, 101, 102])
    loss2 = token_attention_loss_from_maps(maps2, masks, tok_len, ti_pos)
    loss2.backward()
    assert float(loss2) > 0.0 and torch.isfinite(maps2.grad).all()
