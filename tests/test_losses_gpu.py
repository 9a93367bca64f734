"""GPU parity of the loss kernels against the torch restatement of the reference formulas under autograd:
b200_token_attention_loss (trainer/loss.py:10-80: value and the per-layer gradient map) and b200_token_std_loss
(loss.py:222-231, 291-297: value and the gradient of the trainable embedding rows)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _case(n_layers, B, h, w, n_tok, captions, seed=0, Hm=64, Wm=64):
    g = torch.Generator(device="cuda").manual_seed(seed)
    bufs = [(torch.randn(B, h * w, 80, device="cuda", generator=g) * 3.0).to(BF) for _ in range(n_layers)]
    maps = [b[:, :, :77] for b in bufs]
    mask = torch.rand(B, 4, Hm, Wm, device="cuda", generator=g)
    train_ids = list(range(1000, 1000 + n_tok))
    from sd_lora_trainer_b200.trainer.loss import token_index_tensors
    tok_len, ti_pos = token_index_tensors(captions, train_ids, device="cuda")
    return maps, mask, tok_len, ti_pos


def _reference(maps, mask, tok_len, ti_pos, h, w, scale):
    from sd_lora_trainer_b200.trainer.loss import token_attention_loss_from_maps
    B = maps[0].shape[0]
    stacked = torch.stack([m.reshape(B, h, w, 77) for m in maps]).detach().clone().requires_grad_(True)
    loss = token_attention_loss_from_maps(stacked, mask, tok_len, ti_pos)
    if loss.requires_grad:
        (loss * scale).backward()
    return loss.detach().float(), stacked.grad


@pytest.mark.parametrize("n_layers,B,h,w,n_tok", [(60, 2, 32, 32, 3), (15, 4, 8, 8, 3), (7, 3, 16, 24, 2), (1, 1, 4, 4, 4)])
def test_token_attention_loss_kernel_matches_the_torch_formula(n_layers, B, h, w, n_tok):
    from sd_lora_trainer_b200 import ops
    tids = list(range(1000, 1000 + n_tok))
    captions = [[1] + tids + [5, 6, 7, 8][: 2 + b] + [2] for b in range(B)]
    if B >= 3:
        captions[1] = [1, 9, 10, 11, 2]                                   # a caption without the trainable tokens
    maps, mask, tok_len, ti_pos = _case(n_layers, B, h, w, n_tok, captions, seed=n_layers)
    scale = 3e-3
    loss, G = ops.token_attention_loss(maps, h, w, 77, mask[:, 0], tok_len, ti_pos, scale)
    torch.cuda.synchronize()
    ref, gref = _reference(maps, mask, tok_len, ti_pos, h, w, scale)
    assert abs(float(loss) - float(ref)) <= 2e-3 * abs(float(ref)), (float(loss), float(ref))
    assert float(G[:, :, 77:].abs().max()) == 0.0
    # every layer receives the same gradient map; autograd (bf16 graph) vs the kernel (fp32, rounded once)
    g0 = gref[0].reshape(B, h * w, 77)
    assert rel(gref[-1].reshape(B, h * w, 77), g0) < 1e-6
    assert rel(G[:, :, :77], g0) < 3e-2, rel(G[:, :, :77], g0)
    # the trainable tokens' columns carry the mask-dependent terms: check them on their own
    for b in range(B):
        if int(ti_pos[b, 0]) >= 0:
            cols = ti_pos[b].tolist()
            assert rel(G[b][:, cols], g0[b][:, cols]) < 3e-2


def test_token_attention_loss_kernel_without_trainable_tokens_is_zero():
    from sd_lora_trainer_b200 import ops
    maps, mask, tok_len, ti_pos = _case(5, 2, 8, 8, 3, [[1, 4, 5, 2], [1, 6, 2]])
    loss, G = ops.token_attention_loss(maps, 8, 8, 77, mask[:, 0], tok_len, ti_pos, 1.0)
    torch.cuda.synchronize()
    assert float(loss) == 0.0 and float(G.abs().max()) == 0.0 and torch.isfinite(G).all()


@pytest.mark.parametrize("dims", [(768,), (768, 1280), (96, 160)])
def test_token_std_loss_kernel(dims):
    from sd_lora_trainer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(len(dims))
    rows = [(torch.randn(3, d, device="cuda", generator=g) * 0.02).to(BF) for d in dims]
    grads = [torch.randn(3, d, device="cuda", generator=g) for d in dims]
    before = [x.clone() for x in grads]
    mu_t, var_t = [0.0151, 0.0149][: len(dims)], [3.1e-4, 2.7e-4][: len(dims)]
    coeff = 0.01 / 2
    loss = ops.token_std_loss(rows, grads, mu_t, var_t, coeff)
    torch.cuda.synchronize()
    tot = 0.0
    for e, r in enumerate(rows):
        x = r.detach().float().requires_grad_(True)
        le = ((mu_t[e] - x.std(-1)) ** 2 / var_t[e]).mean()
        (le * coeff / len(rows)).backward()
        tot += float(le) / len(rows)
        assert rel(grads[e] - before[e], x.grad) < 2e-2, rel(grads[e] - before[e], x.grad)
    assert abs(float(loss) - tot) <= 2e-2 * abs(tot), (float(loss), tot)
