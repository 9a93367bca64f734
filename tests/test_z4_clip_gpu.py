"""GPU parity of the native CLIP text-encoder executor (sd_lora_trainer_b200/clip.py)
and of its activation kernels, against the installed transformers CLIP modules under autograd (bf16 and fp32).  (File
name sorts late on purpose: written after the round's GPU budget was spent - first executed by the round-end GPU run;
the step keeps the stock-transformers text path as its default until this file has been seen green.)"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("kind", ["gelu", "quick_gelu"])
@pytest.mark.parametrize("n", [8, 1000, 154 * 3072 + 3])
def test_activation_kernels_match_torch(kind, n):
    from sd_lora_trainer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    x = (torch.randn(n, device="cuda", generator=g) * 3).to(BF)
    dy = torch.randn(n, device="cuda", generator=g).to(BF)
    k = ops.ACT_GELU if kind == "gelu" else ops.ACT_QUICK_GELU
    xf = x.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(xf) if kind == "gelu" else xf * torch.sigmoid(1.702 * xf)
    ref.backward(dy.float())
    y, dx = ops.act_fwd(x, k), ops.act_bwd(dy, x, k)
    torch.cuda.synchronize()
    # one bf16 rounding of an fp32 result: half an ulp (<= 2^-8 relative) plus fp32 evaluation noise
    assert ((y.float() - ref).abs() <= 1.01 * 2.0 ** -8 * ref.abs() + 2e-6).all()
    assert ((dx.float() - xf.grad).abs() <= 1.01 * 2.0 ** -8 * xf.grad.abs() + 2e-6).all()
    # unaligned views take the scalar path
    if n > 8:
        y2 = ops.act_fwd(x[1:].clone()[1:].contiguous(), k)
        assert torch.equal(y2, y[2:])


def _pair(tes, sdxl, n_tok=3):
    """(reference encoders with trainable full tables, product stack over frozen tables + separate rows)."""
    from sd_lora_trainer_b200.clip import TextStackB200
    live = [te for te in tes if te is not None]
    rows, frozen = [], []
    for te in live:
        te.requires_grad_(False)
        te.text_model.embeddings.token_embedding.weight.requires_grad_(True)
        rows.append(te.text_model.embeddings.token_embedding.weight.data[-n_tok:].clone())
        t2 = copy.deepcopy(te)
        emb = t2.text_model.embeddings.token_embedding
        emb.weight = torch.nn.Parameter(emb.weight.data[:-n_tok].clone(), requires_grad=False)
        frozen.append(t2)
    return rows, TextStackB200(sdxl, (frozen + [None])[:2], rows, "cuda:0")


def _ids(B, L, vocab, n_tok, eos, bos, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, min(vocab, 40000) - 2, (B, L), generator=g)
    ids[:, 0] = bos
    ids[:, 1:1 + n_tok] = torch.arange(vocab, vocab + n_tok)
    for b in range(B):
        ids[b, 9 + 3 * b:] = eos
    return ids.cuda()


def _check(tes, sdxl, B, tol_fwd=3.0, tol_bwd=3.0):
    from oracle.text import encode_prompt
    cfg = tes[0].config
    vocab = cfg.vocab_size - 3
    ids = [_ids(B, 77, vocab, 3, cfg.eos_token_id, cfg.bos_token_id, seed=i) for i in range(2 if sdxl else 1)]
    tes32 = [copy.deepcopy(t).float() if t is not None else None for t in tes]
    rows, stack = _pair(tes, sdxl)
    for t in tes32:
        if t is not None:
            t.requires_grad_(False)
            t.text_model.embeddings.token_embedding.weight.requires_grad_(True)
    pe, pooled = encode_prompt(sdxl, tes, ids)
    pe32, pooled32 = encode_prompt(sdxl, tes32, ids)
    g = torch.Generator(device="cuda").manual_seed(5)
    d_pe = (torch.randn(pe.shape, device="cuda", generator=g) * 0.1).to(BF)
    d_pool = (torch.randn(pooled.shape, device="cuda", generator=g) * 0.1).to(BF) if pooled is not None else None
    torch.autograd.backward([pe] + ([pooled] if sdxl else []), [d_pe] + ([d_pool] if sdxl else []))
    torch.autograd.backward([pe32] + ([pooled32] if sdxl else []), [d_pe.float()] + ([d_pool.float()] if sdxl else []))
    out_pe, out_pool = stack.encode_prompt(ids, need_bwd=True)
    gbuf = [torch.zeros(3, r.shape[1], device="cuda") for r in rows]
    stack.backward(d_pe, d_pool, gbuf)
    torch.cuda.synchronize()
    # as close to exact (fp32) arithmetic as transformers' own bf16 path is (x3 + floor)
    assert rel(out_pe, pe32) <= tol_fwd * rel(pe, pe32) + 2e-3, (rel(out_pe, pe32), rel(pe, pe32))
    if sdxl:
        assert rel(out_pool, pooled32) <= tol_fwd * rel(pooled, pooled32) + 2e-3, (rel(out_pool, pooled32), rel(pooled, pooled32))
    live, live32 = [t for t in tes if t is not None], [t for t in tes32 if t is not None]
    for te, te32, gb in zip(live, live32, gbuf):
        g32 = te32.text_model.embeddings.token_embedding.weight.grad[-3:]
        g16 = te.text_model.embeddings.token_embedding.weight.grad[-3:]
        assert float(g32.abs().max()) > 0
        assert rel(gb, g32) <= tol_bwd * rel(g16, g32) + 5e-3, (rel(gb, g32), rel(g16, g32))


@pytest.mark.parametrize("family,B", [("sdxl", 2), ("sd15", 1), ("sd15", 4)])
def test_native_clip_tiny_matches_transformers(family, B):
    from oracle.text import build_text_encoders, initialize_new_tokens
    tes = build_text_encoders(family, tiny=True, seed=3)
    initialize_new_tokens(tes, 3, seed=3)
    tes = [t.cuda().to(BF) if t is not None else None for t in tes]
    _check(tes, family == "sdxl", B)


def test_native_clip_full_size_sdxl_matches_transformers():
    """CLIP-L (12 x 768, quick_gelu) + OpenCLIP bigG (32 x 1280, gelu, projected pooled output), random weights."""
    from oracle.text import build_text_encoders, initialize_new_tokens
    tes = build_text_encoders("sdxl", tiny=False, seed=1)
    initialize_new_tokens(tes, 3, seed=1)
    tes = [t.cuda().to(BF) for t in tes]
    _check(tes, True, 2)


@pytest.mark.parametrize("family,rank,batch", [("sdxl", 8, 2), ("sd15", 4, 1)])
def test_training_step_with_native_text_matches_oracle(family, rank, batch):
    from tests.test_unet_gpu import _build, _fp32_twin
    from oracle.text import build_text_encoders
    from sd_lora_trainer_b200.step import StepConfig as PCfg, TrainerB200
    cfg, orc, inputs = _build(family, rank=rank, batch=batch)
    pcfg = PCfg(**{k: getattr(cfg, k) for k in PCfg.__dataclass_fields__ if hasattr(cfg, k)})
    tes = build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1)
    ti_init = [te.text_model.embeddings.token_embedding.weight.data[-cfg.n_tokens:].clone()
               for te in orc.text_encoders if te is not None]
    tr = TrainerB200(pcfg, orc.unet.state_dict(), tes, device="cuda", ti_init=ti_init, native_text=True)
    orc32 = _fp32_twin(cfg, orc)
    out_32 = orc32.step(inputs, completion_f=0.0, do_optimizer=False)
    out_o = orc.step(inputs, completion_f=0.0, do_optimizer=False)
    out_p = tr.step(inputs, completion_f=0.0, do_optimizer=False)
    torch.cuda.synchronize()
    for key in ("img_loss", "token_attention_loss", "token_std_loss", "tot_loss"):
        a, b, c = float(out_p[key]), float(out_o[key]), float(out_32[key])
        assert abs(a - c) / abs(c) <= max(1e-3, 1.5 * abs(b - c) / abs(c)), f"{key}: ours {a} bf16-oracle {b} fp32-oracle {c}"
    off = tr.store.n_lora
    for te, rows in zip([t for t in orc.text_encoders if t is not None], tr.ti_rows):
        gref = te.text_model.embeddings.token_embedding.weight.grad[-cfg.n_tokens:]
        assert rel(tr.store.grads[off:off + rows.numel()].view_as(rows), gref) < 0.25
        off += rows.numel()
    # the step under a CUDA graph with the native text stack: replays reproduce the eager loss
    tr2 = TrainerB200(pcfg, orc.unet.state_dict(), build_text_encoders(cfg.family, cfg.tiny, seed=cfg.seed + 1), device="cuda",
                      ti_init=ti_init, native_text=True, use_cuda_graph=True)
    l_graph = float(tr2.step(inputs, completion_f=0.0)["tot_loss"])
    assert abs(l_graph - float(out_p["tot_loss"])) <= 2e-3 * abs(float(out_p["tot_loss"]))
