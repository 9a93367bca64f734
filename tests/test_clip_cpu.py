"""CPU tests of the native CLIP text-encoder executor's HOST logic (sd_lora_trainer_b200/clip.py; SURVEY.md 8a row a2,
8f row 3) through tests/cpu_mock_ops.py, against the installed transformers CLIP modules under autograd - the very
modules the reference's get_conditioning_signals path runs (trainer/inference.py:131-177)."""
import pytest
import torch

from tests import cpu_mock_ops

BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _encoders(family, seed=3):
    from oracle.text import build_text_encoders, initialize_new_tokens
    tes = build_text_encoders(family, tiny=True, seed=seed)
    train_ids = initialize_new_tokens(tes, 3, seed=seed)
    return [te.to(BF) if te is not None else None for te in tes], train_ids


def _ids(B, L, vocab, train_ids, eos, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (B, L), generator=g)
    ids[:, 0] = vocab - 2
    ids[:, 1:1 + len(train_ids)] = torch.tensor(train_ids)
    for b in range(B):
        ids[b, 9 + 3 * b:] = eos                                 # first EOS at a different position per sample
    return ids


@pytest.mark.parametrize("family", ["sdxl", "sd15"])
def test_native_clip_matches_transformers_autograd(monkeypatch, family):
    cpu_mock_ops.install(monkeypatch)
    from oracle.text import encode_prompt
    from sd_lora_trainer_b200.clip import TextStackB200
    tes, train_ids = _encoders(family)
    sdxl = family == "sdxl"
    B, L = 2, 77
    vocab = tes[0].config.vocab_size - 3
    ids = [_ids(B, L, vocab, train_ids, tes[0].config.eos_token_id, seed=i) for i in range(2 if sdxl else 1)]
    # reference: full tables trainable, autograd
    for te in tes:
        if te is not None:
            te.requires_grad_(False)
            te.text_model.embeddings.token_embedding.weight.requires_grad_(True)
    pe, pooled = encode_prompt(sdxl, tes, ids)
    g = torch.Generator().manual_seed(5)
    d_pe = (torch.randn(pe.shape, generator=g) * 0.1).to(BF)
    d_pool = (torch.randn(pooled.shape, generator=g) * 0.1).to(BF) if pooled is not None else None
    roots, grads = [pe], [d_pe]
    if pooled is not None:
        roots.append(pooled)
        grads.append(d_pool)
    torch.autograd.backward(roots, grads)
    # product: frozen tables + 3 separate rows, explicit backward into fp32 row-gradient buffers
    live = [te for te in tes if te is not None]
    rows = [te.text_model.embeddings.token_embedding.weight.data[-3:].clone() for te in live]
    import copy
    frozen = []
    for te in live:
        t2 = copy.deepcopy(te)
        emb = t2.text_model.embeddings.token_embedding
        emb.weight = torch.nn.Parameter(emb.weight.data[:-3].clone(), requires_grad=False)
        frozen.append(t2)
    stack = TextStackB200(sdxl, (frozen + [None])[:2], rows, "cpu")
    out_pe, out_pool = stack.encode_prompt(ids, need_bwd=True)
    assert out_pe.shape == pe.shape and rel(out_pe, pe) < 2e-2, rel(out_pe, pe)
    if pooled is not None:
        assert rel(out_pool, pooled) < 2e-2, rel(out_pool, pooled)
    gbuf = [torch.zeros(3, r.shape[1]) for r in rows]
    stack.backward(d_pe, d_pool, gbuf)
    for te, gb in zip(live, gbuf):
        gref = te.text_model.embeddings.token_embedding.weight.grad[-3:]
        assert float(gref.abs().max()) > 0
        assert rel(gb, gref) < 6e-2, rel(gb, gref)
    # forward-only mode keeps no state
    stack.encode_prompt(ids, need_bwd=False)
    for e in stack.encs:
        assert e._fw is None and all(b.sv is None and b.ln1.sv is None for b in e.blocks)


def test_native_clip_row_gradient_sums_repeated_tokens(monkeypatch):
    """A trainable token that appears at several positions (and in several samples) receives the SUM of their gradients."""
    cpu_mock_ops.install(monkeypatch)
    from sd_lora_trainer_b200.clip import CLIPTextB200
    tes, train_ids = _encoders("sd15", seed=4)
    te = tes[0]
    vocab = te.config.vocab_size - 3
    ids = _ids(2, 16, vocab, train_ids, te.config.eos_token_id)
    ids[1, 5] = train_ids[0]                                     # token 0 now occurs three times in total
    te.requires_grad_(False)
    te.text_model.embeddings.token_embedding.weight.requires_grad_(True)
    out = te(ids)[0]
    d = (torch.randn(out.shape, generator=torch.Generator().manual_seed(1)) * 0.1).to(BF)
    out.backward(d)
    gref = te.text_model.embeddings.token_embedding.weight.grad[-3:]
    import copy
    t2 = copy.deepcopy(te)
    emb = t2.text_model.embeddings.token_embedding
    rows = emb.weight.data[-3:].clone()
    emb.weight = torch.nn.Parameter(emb.weight.data[:-3].clone(), requires_grad=False)
    enc = CLIPTextB200(t2, rows, "last", "cpu")
    h, _ = enc.forward(ids)
    assert rel(h, out) < 2e-2
    gb = torch.zeros(3, rows.shape[1])
    enc.backward(d, None, gb)
    assert rel(gb, gref) < 6e-2


@pytest.mark.parametrize("family,rank,batch,hw", [("sdxl", 8, 2, 8), ("sd15", 4, 1, 8)])
def test_step_with_native_text_matches_oracle(monkeypatch, family, rank, batch, hw):
    """The whole step with the native text stack switched on: losses and TI-row gradients against the oracle."""
    cpu_mock_ops.install(monkeypatch)
    from tests.test_host_logic_cpu import _setup
    cfg, orc, tr0, inputs = _setup(family, rank, batch, hw)
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    tes = build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1)
    ti_init = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
               for te in orc.text_encoders if te is not None]
    tr = TrainerB200(pcfg, orc.unet.state_dict(), tes, device="cpu", ti_init=ti_init, native_text=True)
    assert tr.text is not None
    out_o = orc.step(inputs, do_optimizer=False)
    out_p = tr.step(inputs, do_optimizer=False)
    for key in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        a, b = float(out_p[key]), float(out_o[key])
        assert abs(a - b) / abs(b) <= 2e-3, f"{key}: ours {a} vs oracle {b}"
    off = tr.store.n_lora
    for te, rows in zip([t for t in orc.text_encoders if t is not None], tr.ti_rows):
        gref = te.text_model.embeddings.token_embedding.weight.grad[-cfg.n_tokens:]
        assert rel(tr.store.grads[off:off + rows.numel()].view_as(rows), gref) < 0.25
        off += rows.numel()


def test_native_clip_full_size_sdxl(monkeypatch):
    """CLIP-L (12 x 768, quick_gelu) + OpenCLIP bigG (32 x 1280, gelu, projection) at full size: descriptors pass the C
    wrappers' rules (cpu_mock_ops enforces them) and the result tracks transformers' bf16 path."""
    cpu_mock_ops.install(monkeypatch)
    import copy
    from oracle.text import build_text_encoders, encode_prompt, initialize_new_tokens
    from sd_lora_trainer_b200.clip import TextStackB200
    tes = build_text_encoders("sdxl", tiny=False, seed=1)
    initialize_new_tokens(tes, 3, seed=1)
    tes = [t.to(BF) for t in tes]
    vocab = 49408
    ids = torch.full((1, 77), 49407, dtype=torch.long)
    seq = [49406, vocab, vocab + 1, vocab + 2, 320, 1125, 539, 49407]
    ids[0, :len(seq)] = torch.tensor(seq)
    ids = [ids, ids.clone()]
    rows, frozen = [], []
    for te in tes:
        te.requires_grad_(False)
        te.text_model.embeddings.token_embedding.weight.requires_grad_(True)
        rows.append(te.text_model.embeddings.token_embedding.weight.data[-3:].clone())
        t2 = copy.deepcopy(te)
        emb = t2.text_model.embeddings.token_embedding
        emb.weight = torch.nn.Parameter(emb.weight.data[:-3].clone(), requires_grad=False)
        frozen.append(t2)
    pe, pooled = encode_prompt(True, tes, ids)
    g = torch.Generator().manual_seed(5)
    d_pe, d_pool = (torch.randn(pe.shape, generator=g) * 0.1).to(BF), (torch.randn(pooled.shape, generator=g) * 0.1).to(BF)
    torch.autograd.backward([pe, pooled], [d_pe, d_pool])
    stack = TextStackB200(True, frozen, rows, "cpu")
    assert len(stack.encs[0].blocks) == 11 and len(stack.encs[1].blocks) == 32       # encoder 1 skips its last layer
    out_pe, out_pool = stack.encode_prompt(ids)
    assert out_pe.shape == (1, 77, 2048) and out_pool.shape == (1, 1280)
    assert rel(out_pe, pe) < 4e-2 and rel(out_pool, pooled) < 5e-2                   # bf16 vs bf16 over 32 layers
    gb = [torch.zeros(3, r.shape[1]) for r in rows]
    stack.backward(d_pe, d_pool, gb)
    for te, g_ in zip(tes, gb):
        assert rel(g_, te.text_model.embeddings.token_embedding.weight.grad[-3:]) < 6e-2
