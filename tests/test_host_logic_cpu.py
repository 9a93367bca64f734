"""CPU tests of the executor's HOST logic: the product's forward/backward graph (operand descriptors, strides, skip and
gradient plumbing, LoRA slot layout, optimizer wiring) driven through tests/cpu_mock_ops.py and compared with the
oracle.  The CUDA kernels themselves are covered by the `-m gpu` tests."""
import pytest
import torch

from tests import cpu_mock_ops

BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _setup(family, rank, batch, hw, disable_ti=False):
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    cfg = StepConfig(family=family, tiny=True, resolution=hw * 8, lora_rank=rank, disable_ti=disable_ti)
    orc = OracleTrainer(cfg, device="cpu")
    g = torch.Generator().manual_seed(7)
    for n, p in orc.unet.named_parameters():
        if "lora_B" in n:
            p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(BF))
    inputs = make_inputs(cfg, batch=batch, latent_hw=hw, face_mask=True, train_ids=orc.train_ids)
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    tes = build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1)
    ti_init = None
    if not disable_ti:
        ti_init = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
                   for te in orc.text_encoders if te is not None]
    tr = TrainerB200(pcfg, orc.unet.state_dict(), tes, device="cpu", ti_init=ti_init)
    return cfg, orc, tr, inputs


# hw = 12: a latent width that does not divide 128 (768x768 trains 96-wide maps, training_args_face_sd15.json): every 3x3
# convolution and conv-LoRA takes the im2col fallback
# ("sdxl", 16, 2, 32): 512 tokens at the first attention level - the fused q|k|v projection (rank-48 side path, LinQKV) runs
# there, the separate-projection fallback below it
@pytest.mark.parametrize("family,rank,batch,hw", [("sdxl", 8, 2, 8), ("sd15", 4, 1, 8), ("sd15", 8, 1, 12), ("sdxl", 16, 2, 32)])
def test_step_host_logic_matches_oracle(monkeypatch, family, rank, batch, hw):
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup(family, rank, batch, hw)
    p_before = tr.store.export_peft()
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    assert torch.equal(out_p["noisy_latent"], out_o["noisy_latent"])
    for key in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 2e-3, f"{key}: ours {a} vs oracle {b}"
    pred = out_p["model_pred8"][:, :4].reshape(batch, hw, hw, 4).permute(0, 3, 1, 2)
    assert rel(pred, out_o["model_pred"]) < 2e-2
    for s, so in zip(out_p["attention_scores"], out_o["attention_scores"]):
        assert rel(s, so) < 8e-2          # bf16 noise floor of the per-head-rounded oracle scores; fp32-referenced bound is in the GPU test
    ours = tr.store.export_peft(grads=True)
    n_checked = 0
    for n, p in orc.unet.named_parameters():
        if p.grad is not None:
            assert rel(ours[n].reshape(p.grad.shape), p.grad) < 0.25, n   # bf16-vs-bf16 backward noise; fp32-referenced bound: GPU test
            n_checked += 1
    assert n_checked == 2 * len(tr.store.slots)
    off = tr.store.n_lora
    for te, rows in zip([t for t in orc.text_encoders if t is not None], tr.ti_rows):
        gref = te.text_model.embeddings.token_embedding.weight.grad[-cfg.n_tokens:]
        assert rel(tr.store.grads[off:off + rows.numel()].view_as(rows), gref) < 0.25
        off += rows.numel()
    orc.optimizer_step()
    tr.optimizer_step()
    after = tr.store.export_peft()
    changed = 0
    for n, p in orc.unet.named_parameters():
        if "lora_" in n:
            d_o = p.detach().float() - p_before[n].reshape(p.shape).float()
            d_p = after[n].reshape(p.shape).float() - p_before[n].reshape(p.shape).float()
            changed += int((d_o != 0).sum())
            assert float((d_o - d_p).abs().max()) <= 2.5 * float(d_o.abs().max() + 1e-12), n
    assert changed > 0
    assert float(tr.store.grads.abs().max()) == 0.0


def test_step_host_logic_rank32(monkeypatch):
    """BASELINE config 4's rank: the rank-32 side path / conv-LoRA segments pass the wrappers' argument rules; with
    32 random rank-1 terms per layer the two bf16 paths drift further apart than at rank 8 (looser loss bound)."""
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup("sdxl", 32, 1, 8)
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    for key in ("img_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 5e-3, f"{key}: ours {a} vs oracle {b}"
    assert all(s.r == 32 and s.rs == 32 for s in tr.store.slots)


def test_lora_store_roundtrip_and_layout():
    from sd_lora_trainer_b200.unet import LoraStore
    st = LoraStore("cpu", 1.0)
    a = st.add("x.to_q", "linear", 4, 64, 32)
    c = st.add("y.conv2", "conv", 4, 16, 24)
    st.finalize(extra=10)
    assert a.offA % 8 == 0 and a.offB % 8 == 0 and c.offA % 8 == 0 and c.offB % 8 == 0
    assert st.params.numel() == st.n_lora + 10
    assert st.numel_logical == 4 * 64 + 32 * 4 + 9 * 4 * 16 + 24 * 4
    sd = {"x.to_q.lora_A.default.weight": torch.randn(4, 64), "x.to_q.lora_B.default.weight": torch.randn(32, 4),
          "y.conv2.lora_A.default.weight": torch.randn(4, 16, 3, 3), "y.conv2.lora_B.default.weight": torch.randn(24, 4, 1, 1)}
    st.load_peft(sd)
    back = st.export_peft()
    for k, v in sd.items():
        assert torch.equal(back[k], v.to(torch.bfloat16)), k
    assert float(a.B()[:, 4:].abs().max()) == 0.0          # rank padding stays zero


def test_product_never_imports_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sd_lora_trainer_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("name", ["tiny_sdxl", "tiny_sd15", "sdxl", "sd15"])
def test_param_shapes_match_oracle_state_dict(name):
    """The product's own parameter enumeration (used for random init) names and shapes every diffusers parameter."""
    from oracle.unet import UNet2DConditionModel, UNetConfig
    from sd_lora_trainer_b200.arch import by_name
    from sd_lora_trainer_b200.init import param_shapes
    with torch.device("meta"):
        ref = UNet2DConditionModel(getattr(UNetConfig, name)())
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    got = dict(param_shapes(by_name(name)))
    assert got == want


def test_frozen_ti_phase_skips_text_backward_without_changing_the_step(monkeypatch):
    """completion_f > freeze_ti_after_completion_f: ti_lr = 0 (main.py:273-274).  The product skips the CLIP backward;
    losses, LoRA gradients and the post-step parameters (LoRA factors AND token rows) still match the oracle, which
    back-propagates into the embedding tables and applies AdamW with lr = 0."""
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup("sdxl", 8, 2, 8)
    rows_before = [r.detach().clone() for r in tr.ti_rows]
    tabs_before = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
                   for te in orc.text_encoders if te is not None]
    p_before = tr.store.export_peft()
    out_o = orc.step(inputs, completion_f=0.9, do_optimizer=False)
    out_p = tr.step(inputs, completion_f=0.9, do_optimizer=False)
    assert "token_std_loss" not in out_o and "token_std_loss" not in out_p           # regulariser is off with ti_lr = 0
    for key in ("img_loss", "token_attention_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 2e-3, f"{key}: ours {a} vs oracle {b}"
    assert float(tr.store.grads[tr.store.n_lora:].abs().max()) == 0.0               # no TI gradient was produced
    ours = tr.store.export_peft(grads=True)
    for n, p in orc.unet.named_parameters():
        if p.grad is not None:
            assert rel(ours[n].reshape(p.grad.shape), p.grad) < 0.25, n
    orc.optimizer_step()
    tr.optimizer_step()
    for r0, r1, t0, te in zip(rows_before, tr.ti_rows, tabs_before, [t for t in orc.text_encoders if t is not None]):
        assert torch.equal(r0, r1.detach())                                          # ours: untouched
        assert torch.equal(t0, te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:])   # oracle: lr = 0
    after = tr.store.export_peft()
    assert any(not torch.equal(after[k], p_before[k]) for k in after)               # the LoRA factors did move


def test_conditioning_cache_encodes_each_caption_once(monkeypatch):
    """disable_ti (or frozen TI): a caption's embeddings cannot change, so the step caches them per token-id row and only
    unseen captions reach the text encoders; the step's results equal the uncached path bit for bit."""
    cpu_mock_ops.install(monkeypatch)
    import sd_lora_trainer_b200.step as step_mod
    calls = []
    real = step_mod.encode_prompt

    def spy(is_sdxl, tes, ids):
        calls.append(int(ids[0].shape[0]))
        return real(is_sdxl, tes, ids)

    monkeypatch.setattr(step_mod, "encode_prompt", spy)
    cfg, orc, tr, inputs = _setup("sdxl", 8, 2, 8, disable_ti=True)
    assert tr.cache_text and not tr.ti_rows
    out1 = tr.step(inputs, do_optimizer=False)
    assert calls == [2]                                         # both captions of the batch were new
    g1 = tr.store.grads.clone()
    tr.store.grads.zero_()
    out2 = tr.step(inputs, do_optimizer=False)
    assert calls == [2]                                         # second visit: served from the cache
    assert float(out1["tot_loss"]) == float(out2["tot_loss"]) and torch.equal(g1, tr.store.grads)
    swapped = dict(inputs, token_ids=[t.flip(0) for t in inputs["token_ids"]])
    tr.step(swapped, do_optimizer=False)
    assert calls == [2]                                         # same captions in another order: still no encoder call
    monkeypatch.setenv("B200_TEXT_CACHE", "0")
    cfg, orc, tr2, _ = _setup("sdxl", 8, 2, 8, disable_ti=True)
    tr2.store.params.copy_(tr.store.params)
    tr2.store.grads.zero_()
    out3 = tr2.step(inputs, do_optimizer=False)
    assert not tr2.cache_text and calls == [2, 2]                # the uncached path encodes inside the step
    assert float(out3["tot_loss"]) == float(out1["tot_loss"]) and torch.equal(tr2.store.grads, g1)


def test_step_host_logic_rank64_falls_back_to_two_launch_lora(monkeypatch):
    """train_configs/training_args_style_sd15_noti.json trains rank 64, beyond the in-kernel side path (rank <= 32): the
    linear layers compute T / U by their own GEMM and add the low-rank product as a second K segment."""
    cpu_mock_ops.install(monkeypatch)
    cfg, orc, tr, inputs = _setup("sd15", 64, 1, 8, disable_ti=True)
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    a, b = float(out_p["tot_loss"]), float(out_o["tot_loss"])
    assert abs(a - b) / abs(b) <= 1e-2, (a, b)                   # 64 random rank-1 terms per layer: bf16 paths drift more
    ours = tr.store.export_peft(grads=True)
    n = 0
    for name, p in orc.unet.named_parameters():
        if p.grad is not None:
            assert rel(ours[name].reshape(p.grad.shape), p.grad) < 0.3, name
            n += 1
    assert n == 2 * len(tr.store.slots) and all(s.r == 64 for s in tr.store.slots)


def test_unet_lr_zero_means_no_unet_update():
    """ti_SDXL.json: unet_lr = 0 -> the reference builds no UNet optimizer (main.py:164-176)."""
    from sd_lora_trainer_b200.step import StepConfig, lr_schedule
    cfg = StepConfig(unet_lr=0.0)
    assert lr_schedule(cfg, 0, 0.0)[1] == 0.0 and lr_schedule(cfg, 10, 0.1)[1] == 0.0


@pytest.mark.parametrize("family,rank,n_tokens", [("sd15", 6, 4), ("sdxl", 20, 2)])
def test_odd_rank_and_token_count(monkeypatch, family, rank, n_tokens):
    """predict.py lets lora_rank be any integer and n_tokens 1..4: ranks that are not multiples of 8 (padded slots) and
    other token counts through one full step incl. the optimizer."""
    cpu_mock_ops.install(monkeypatch)
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    cfg = StepConfig(family=family, tiny=True, resolution=64, lora_rank=rank, n_tokens=n_tokens)
    orc = OracleTrainer(cfg, device="cpu")
    g = torch.Generator().manual_seed(7)
    for n, p in orc.unet.named_parameters():
        if "lora_B" in n:
            p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(BF))
    inputs = make_inputs(cfg, batch=2, latent_hw=8, face_mask=True, train_ids=orc.train_ids)
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    ti_init = [te.text_model.embeddings.token_embedding.weight.data[-n_tokens:].clone() for te in orc.text_encoders if te is not None]
    tr = TrainerB200(pcfg, orc.unet.state_dict(), build_text_encoders(family, True, seed=cfg.seed + 1), device="cpu", ti_init=ti_init)
    assert all(r.shape[0] == n_tokens for r in tr.ti_rows) and all(s.rs == (rank + 7) // 8 * 8 for s in tr.store.slots)
    out_o, out_p = orc.step(inputs), tr.step(inputs)
    a, b = float(out_p["tot_loss"]), float(out_o["tot_loss"])
    assert abs(a - b) / abs(b) <= 3e-3, (a, b)
    after = tr.store.export_peft()
    for n, p in orc.unet.named_parameters():
        if "lora_" in n:
            assert rel(after[n].reshape(p.shape), p) < 2e-3, n          # same AdamW step from nearly equal gradients
    for s in tr.store.slots:
        assert float(s.B()[:, s.r:].abs().max()) == 0.0 if s.rs > s.r else True      # rank padding never leaves zero


@pytest.mark.parametrize("family,hw", [("sdxl", 8), ("sd15", 12)])
def test_shared_dscores_path_is_value_identical(monkeypatch, family, hw):
    """B200_SHARED_DSCORES=1: one gradient map per resolution handed to every hooked layer (the regulariser only sees the
    layer mean) must reproduce the per-layer autograd path bit for bit: losses, LoRA gradients, token-row gradients."""
    cpu_mock_ops.install(monkeypatch)
    monkeypatch.setenv("B200_SHARED_DSCORES", "0")
    cfg, orc, tr_a, inputs = _setup(family, 8, 2, hw)
    out_a = tr_a.step(inputs, do_optimizer=False)
    monkeypatch.setenv("B200_SHARED_DSCORES", "1")
    cfg, orc, tr_b, inputs = _setup(family, 8, 2, hw)
    assert tr_b.shared_dscores and not tr_a.shared_dscores
    sizes = {s.shape[1] for s in out_a["attention_scores"]}
    assert len(sizes) >= 2                                       # at least one resolution goes through the bicubic adjoint
    out_b = tr_b.step(inputs, do_optimizer=False)
    for k in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        assert float(out_a[k]) == float(out_b[k]), k
    assert torch.equal(tr_a.store.grads, tr_b.store.grads)


def test_reset_optimizer_state_and_close(monkeypatch):
    """After a checkpoint is loaded into a live trainer the optimizer state must restart from the loaded parameters
    (Adam moments zero, Prodigy's start point = the new parameters, d back at d0, conditioning cache empty); close() drops
    the captured graphs / static buffers (what lets destroy_process_group() return under data parallelism)."""
    cpu_mock_ops.install(monkeypatch)
    from oracle.step import OracleTrainer, StepConfig, make_inputs
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    cfg = StepConfig(family="sd15", tiny=True, resolution=64, lora_rank=4, unet_optimizer_type="prodigy", ti_optimizer="prodigy")
    orc = OracleTrainer(cfg, device="cpu")
    inputs = make_inputs(cfg, batch=1, latent_hw=8, face_mask=True, train_ids=orc.train_ids)
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    tr = TrainerB200(pcfg, orc.unet.state_dict(), build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1), device="cpu")
    for _ in range(2):
        tr.step(inputs)
    assert float(tr.store.m.float().abs().max()) > 0 and tr.opt_step == 2
    tr.store.params[:tr.store.n_lora].mul_(0.5)                  # "a checkpoint was loaded"
    tr._text_cache[("x",)] = (None, None)
    tr.reset_optimizer_state()
    assert float(tr.store.m.float().abs().max()) == 0 and float(tr.store.v.float().abs().max()) == 0
    assert float(tr.store.grads.abs().max()) == 0 and float(tr._prodigy_s.float().abs().max()) == 0
    assert torch.equal(tr._prodigy_p0, tr.store.params) and tr.opt_step == 0 and not tr._text_cache
    for name, (lo, hi, scal, ring, dev, kw) in tr._prodigy.items():
        assert float(scal[0]) == float(scal[1]) and abs(float(scal[0]) - 1e-6) < 1e-12 and float(scal[2]) == 0.0
    tr._graphs[("k",)] = object()
    tr._static = {"a": torch.zeros(1)}
    tr.close()
    assert tr._graphs == {} and tr._static is None
    tr.step(inputs)                                              # and the trainer keeps working afterwards


def test_ti_embedding_matches_a_table_with_appended_rows():
    """text.TIEmbedding (frozen table + a small trainable leaf, looked up through a one-hot product) against the reference's
    form: ONE trainable [vocab + n, dim] table indexed directly (main.py:92-100, 368-371) - same output, and the gradient of
    the appended rows equals the gradient of the leaf, duplicates and absent tokens included."""
    import torch
    import torch.nn.functional as F
    from sd_lora_trainer_b200.text import TIEmbedding
    torch.manual_seed(0)
    vocab, dim, n = 50, 16, 3
    table = torch.randn(vocab, dim)
    rows = torch.randn(n, dim, requires_grad=True)
    emb = TIEmbedding(table, rows)
    ids = torch.tensor([[1, 7, vocab + 0, vocab + 2, 7, vocab + 0, 49, 0], [vocab + 2, 3, 3, 3, 3, 3, 3, 3]])
    out = emb(ids)
    full = torch.cat([table, rows.detach()], 0).requires_grad_(True)
    ref = F.embedding(ids, full)
    assert torch.equal(out, ref)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert torch.allclose(rows.grad, full.grad[vocab:], atol=1e-6)
    assert float(rows.grad[1].abs().max()) == 0.0                  # token 1 does not occur
    assert emb.num_embeddings == vocab + n and TIEmbedding(table, None)(ids.clamp(max=vocab - 1)).shape == (2, 8, dim)
