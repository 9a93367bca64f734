"""GPU parity of the fused tcgen05 attention (forward + backward, head_dim 64) against torch fp32 attention + autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("B,H,L,Lk", [(2, 4, 256, 256), (2, 3, 384, 77), (1, 2, 1024, 1024), (2, 2, 200, 150),
                                      (1, 5, 128, 640), (2, 10, 4096, 4096), (1, 2, 130, 64), (1, 1, 128, 65), (2, 2, 256, 192),
                                      (1, 1, 64, 3), (2, 20, 512, 512), (2, 20, 1024, 1024),
                                      # one key block (cross-attention) at SDXL's tile counts: 320 / 640 CTAs on 296 slots
                                      (2, 20, 1024, 77), (2, 20, 1000, 77), (2, 10, 4096, 77)])
def test_flash_attention_fwd_bwd(B, H, L, Lk):
    from sd_lora_trainer_b200 import ops
    C = H * 64
    g = torch.Generator(device="cuda").manual_seed(L + Lk)
    q = torch.randn(B * L, C, device="cuda", generator=g).to(BF)
    k = torch.randn(B * Lk, C, device="cuda", generator=g).to(BF)
    v = torch.randn(B * Lk, C, device="cuda", generator=g).to(BF)
    do = torch.randn(B * L, C, device="cuda", generator=g).to(BF)
    scale = 64 ** -0.5
    o, lse = ops.flash_attn_fwd(q, k, v, B, H, L, Lk, scale)
    dq, dk, dv = ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, Lk, scale)
    # (2, 20, 512, 512) / (2, 20, 1024, 1024) / (2, 10, 4096, 4096): 160 / 320 / 640 work items on 148 SMs - the items of the
    # last, partial round are cut into query ranges (fp32 atomics + last-arriver rounding); a second launch must find the
    # workspace zero again
    dq2, dk2, dv2 = ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, Lk, scale)
    torch.cuda.synchronize()
    assert rel(dk2, dk) < 1e-3 and rel(dv2, dv) < 1e-3 and rel(dq2, dq) < 1e-3
    if L * Lk <= 1024 * 1024:
        qr = q.float().view(B, L, H, 64).transpose(1, 2).requires_grad_(True)
        kr = k.float().view(B, Lk, H, 64).transpose(1, 2).requires_grad_(True)
        vr = v.float().view(B, Lk, H, 64).transpose(1, 2).requires_grad_(True)
        s = qr @ kr.transpose(-1, -2) * scale
        ref = torch.softmax(s, -1) @ vr
        ref.backward(do.float().view(B, L, H, 64).transpose(1, 2))
        unh = lambda t, n: t.transpose(1, 2).reshape(B * n, C)
        assert rel(o, unh(ref, L)) < 1e-2, ("o", rel(o, unh(ref, L)))
        assert rel(lse, torch.logsumexp(s, -1)) < 1e-3
        assert rel(dv, unh(vr.grad, Lk)) < 2e-2, ("dv", rel(dv, unh(vr.grad, Lk)))
        assert rel(dq, unh(qr.grad, L)) < 2e-2, ("dq", rel(dq, unh(qr.grad, L)))
        assert rel(dk, unh(kr.grad, Lk)) < 2e-2, ("dk", rel(dk, unh(kr.grad, Lk)))
    else:
        # full SDXL 64x64-latent level: compare with torch's own fused SDPA in bf16 (memory-safe) and time ours
        qr = q.view(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
        kr = k.view(B, Lk, H, 64).transpose(1, 2).detach().requires_grad_(True)
        vr = v.view(B, Lk, H, 64).transpose(1, 2).detach().requires_grad_(True)
        ref = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)
        ref.backward(do.view(B, L, H, 64).transpose(1, 2))
        unh = lambda t, n: t.transpose(1, 2).reshape(B * n, C)
        assert rel(o, unh(ref, L)) < 2e-2
        assert rel(dq, unh(qr.grad, L)) < 4e-2 and rel(dk, unh(kr.grad, Lk)) < 4e-2 and rel(dv, unh(vr.grad, Lk)) < 4e-2
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(3):
            ops.flash_attn_fwd(q, k, v, B, H, L, Lk, scale)
        ev[0].record()
        for _ in range(5):
            ops.flash_attn_fwd(q, k, v, B, H, L, Lk, scale)
        ev[1].record()
        for _ in range(5):
            ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, Lk, scale)
        ev[2].record()
        torch.cuda.synchronize()
        fl = 4.0 * B * H * L * Lk * 64
        tf, tb = ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5
        print(f"\nflash fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s)  bwd {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TFLOP/s)")


@pytest.mark.parametrize("B,H,L,Lk", [(1, 2, 256, 512), (2, 3, 200, 700)])
def test_flash_forward_rescales_when_later_keys_dominate(B, H, L, Lk):
    """Key blocks whose scores grow by far more than 2^8 from block to block: the running reference must move and O (in
    tensor memory) must be rescaled on every block - a path random normal data never takes after the first block.  Both
    key-half warp groups of a row have to agree on the new reference (the forward keeps one O accumulator)."""
    from sd_lora_trainer_b200 import ops
    C = H * 64
    g = torch.Generator(device="cuda").manual_seed(7)
    q = (4.0 * torch.randn(B * L, C, device="cuda", generator=g)).to(BF)
    k = torch.randn(B, Lk, C, device="cuda", generator=g)
    grow = 1.0 + 3.0 * (torch.arange(Lk, device="cuda") // 128).float()          # x1, x4, x7, ... per 128-key block
    grow[Lk // 2:Lk // 2 + 64] *= 0.05                                            # and one weak half block in between
    k = (k * grow[None, :, None]).reshape(B * Lk, C).to(BF)
    v = torch.randn(B * Lk, C, device="cuda", generator=g).to(BF)
    do = torch.randn(B * L, C, device="cuda", generator=g).to(BF)
    scale = 64 ** -0.5
    o, lse = ops.flash_attn_fwd(q, k, v, B, H, L, Lk, scale)
    dq, dk, dv = ops.flash_attn_bwd(q, k, v, o, do, lse, B, H, L, Lk, scale)
    torch.cuda.synchronize()
    qr = q.float().view(B, L, H, 64).transpose(1, 2).requires_grad_(True)
    kr = k.float().view(B, Lk, H, 64).transpose(1, 2).requires_grad_(True)
    vr = v.float().view(B, Lk, H, 64).transpose(1, 2).requires_grad_(True)
    s = qr @ kr.transpose(-1, -2) * scale
    blockmax = torch.stack([s[..., j:j + 128].amax(-1) for j in range(0, Lk, 128)], -1)
    assert float((blockmax[..., 1:] - blockmax[..., :-1]).amax() * 1.4427) > 8.0       # the data does what the docstring says
    ref = torch.softmax(s, -1) @ vr
    ref.backward(do.float().view(B, L, H, 64).transpose(1, 2))
    unh = lambda t, n: t.transpose(1, 2).reshape(B * n, C)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert rel(o, unh(ref, L)) < 1e-2, ("o", rel(o, unh(ref, L)))
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-3
    assert rel(dv, unh(vr.grad, Lk)) < 2e-2 and rel(dq, unh(qr.grad, L)) < 2e-2 and rel(dk, unh(kr.grad, Lk)) < 2e-2


@pytest.mark.parametrize("rows,H,d_src,d_dst", [(300, 8, 40, 64), (300, 8, 64, 40), (77, 2, 32, 64), (513, 5, 64, 64)])
def test_head_pad_is_an_exact_repitch(rows, H, d_src, d_dst):
    from sd_lora_trainer_b200 import ops
    x = torch.randn(rows, H * d_src, device="cuda").to(BF)
    y = ops.head_pad(x, H, d_src, d_dst)
    torch.cuda.synchronize()
    n = min(d_src, d_dst)
    ref = torch.zeros(rows, H, d_dst, device="cuda", dtype=BF)
    ref[:, :, :n] = x.view(rows, H, d_src)[:, :, :n]
    assert torch.equal(y, ref.view(rows, H * d_dst))


@pytest.mark.parametrize("B,H,L,Lk,d", [(2, 8, 1024, 1024, 40), (2, 8, 512, 77, 40), (1, 2, 256, 256, 32), (4, 8, 4096, 4096, 40)])
def test_narrow_heads_through_the_fused_kernel(B, H, L, Lk, d):
    """SD1.5's head dims below 64 (d = 40 at the 64x64-latent level, L = 4096) run on the 64-wide fused kernel after a zero
    re-pitch of q / k / v (unet.Attn): same softmax(q k^T / sqrt(d)) v and gradients as torch attention on the d-wide heads."""
    from sd_lora_trainer_b200 import ops
    C = H * d
    g = torch.Generator(device="cuda").manual_seed(L + Lk + d)
    q, k, v, do = (torch.randn(B * n, C, device="cuda", generator=g).to(BF) for n in (L, Lk, Lk, L))
    scale = d ** -0.5
    qf, kf, vf, dof = (ops.head_pad(t, H, d, 64) for t in (q, k, v, do))
    of, lse = ops.flash_attn_fwd(qf, kf, vf, B, H, L, Lk, scale)
    dqf, dkf, dvf = ops.flash_attn_bwd(qf, kf, vf, of, dof, lse, B, H, L, Lk, scale)
    o, dq, dk, dv = (ops.head_pad(t, H, 64, d) for t in (of, dqf, dkf, dvf))
    torch.cuda.synchronize()
    # the padded channels of every output are exactly zero
    for t in (of, dqf, dkf, dvf):
        assert float(t.view(-1, H, 64)[:, :, d:].abs().max()) == 0.0
    qr, kr, vr = (t.view(B, n, H, d).transpose(1, 2).detach().requires_grad_(True) for t, n in ((q, L), (k, Lk), (v, Lk)))
    if L * Lk <= 1024 * 1024:
        qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (qr, kr, vr))
        ref = torch.softmax(qr @ kr.transpose(-1, -2) * scale, -1) @ vr
        ref.backward(do.float().view(B, L, H, d).transpose(1, 2))
        tol = (1e-2, 2e-2)
    else:
        ref = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)       # torch's own fused bf16 kernel (memory-safe)
        ref.backward(do.view(B, L, H, d).transpose(1, 2))
        tol = (2e-2, 4e-2)
    unh = lambda t, n: t.transpose(1, 2).reshape(B * n, C)
    assert rel(o, unh(ref, L)) < tol[0], ("o", rel(o, unh(ref, L)))
    for name, a, b_, n in (("dq", dq, qr.grad, L), ("dk", dk, kr.grad, Lk), ("dv", dv, vr.grad, Lk)):
        assert rel(a, unh(b_, n)) < tol[1], (name, rel(a, unh(b_, n)))


@pytest.mark.parametrize("B,H,L,Lk,d", [(2, 20, 1024, 77, 64), (2, 10, 4096, 77, 64), (1, 3, 300, 77, 64), (2, 8, 256, 77, 40), (1, 2, 130, 64, 64)])
def test_score_hook_gradient_folded_into_the_backward(B, H, L, Lk, d):
    """The DAAM hook's head-summed pre-softmax score  sc[b, l, t] = sum_h scale * q_h . k_h  has the gradient dsc (one map,
    shared by all heads); passed to the fused backward it must add  scale * dsc . K_h  to dQ_h and  scale * dsc^T . Q_h
    to dK_h - what the two extra GEMMs of the unfused form computed."""
    from sd_lora_trainer_b200 import ops
    C = H * d
    g = torch.Generator(device="cuda").manual_seed(L + Lk + d)
    q, k, v, do = (torch.randn(B * n, C, device="cuda", generator=g).to(BF) for n in (L, Lk, Lk, L))
    Lp = (Lk + 7) // 8 * 8
    dsc = torch.zeros(B, L, Lp, device="cuda", dtype=BF)
    dsc[:, :, :Lk] = (torch.randn(B, L, Lk, device="cuda", generator=g) * 0.3).to(BF)
    scale = d ** -0.5
    if d != 64:
        qf, kf, vf, dof = (ops.head_pad(t, H, d, 64) for t in (q, k, v, do))
    else:
        qf, kf, vf, dof = q, k, v, do
    of, lse = ops.flash_attn_fwd(qf, kf, vf, B, H, L, Lk, scale)
    dq0, dk0, dv0 = ops.flash_attn_bwd(qf, kf, vf, of, dof, lse, B, H, L, Lk, scale)
    dq1, dk1, dv1 = ops.flash_attn_bwd(qf, kf, vf, of, dof, lse, B, H, L, Lk, scale, dsc=dsc)
    torch.cuda.synchronize()
    assert rel(dv1, dv0) < 1e-3                # dV does not see the hook (query-split launches sum dV with fp32 atomics: order noise)
    qh = qf.float().view(B, L, H, 64).transpose(1, 2)
    kh = kf.float().view(B, Lk, H, 64).transpose(1, 2)
    gs = dsc[:, :, :Lk].float()[:, None] * scale                                      # [B, 1, L, Lk]
    add_q = (gs @ kh).transpose(1, 2).reshape(B * L, H * 64)
    add_k = (gs.transpose(-1, -2) @ qh).transpose(1, 2).reshape(B * Lk, H * 64)
    assert rel(dq1.float() - dq0.float(), add_q) < 3e-2, rel(dq1.float() - dq0.float(), add_q)
    assert rel(dk1.float() - dk0.float(), add_k) < 3e-2, rel(dk1.float() - dk0.float(), add_k)
