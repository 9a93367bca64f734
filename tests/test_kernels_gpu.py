"""GPU parity of the HBM-bound kernels (through the C ABI) against plain torch references of the same ops."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rand(*shape, scale=1.0, seed=0, dtype=BF):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def _close(out, ref, tol=2e-2, what=""):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"{what}: max abs err {err:.4g} vs max ref {den:.4g}"


@pytest.mark.parametrize("B,HW,C,silu", [(2, 256, 64, True), (2, 1024, 320, True), (1, 4096, 640, False),
                                         (2, 64, 1280, True), (2, 256, 960, True), (2, 100, 2560, True)])
def test_groupnorm_fwd_bwd(B, HW, C, silu):
    from sd_lora_trainer_b200 import ops
    x = _rand(B * HW, C, seed=1) + 0.5
    gamma, beta = (_rand(C, seed=2) * 0.2 + 1).to(BF), _rand(C, seed=3, scale=0.2)
    dy = _rand(B * HW, C, seed=4)
    y, stats = ops.groupnorm_fwd(x, gamma, beta, B, HW, C, 32, 1e-5, silu)
    xr = x.float().view(B, HW, C).permute(0, 2, 1).reshape(B, C, HW, 1).requires_grad_(True)
    ref = F.group_norm(xr, 32, gamma.float(), beta.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    _close(y, ref.reshape(B, C, HW).permute(0, 2, 1).reshape(B * HW, C), what="gn fwd")
    ref.backward(dy.float().view(B, HW, C).permute(0, 2, 1).reshape(B, C, HW, 1))
    dx = ops.groupnorm_bwd(dy, x, gamma, beta, stats, B, HW, C, 32, silu)
    _close(dx, xr.grad.reshape(B, C, HW).permute(0, 2, 1).reshape(B * HW, C), tol=3e-2, what="gn bwd")


@pytest.mark.parametrize("rows,C", [(300, 64), (2048, 1280), (513, 640), (64, 320), (40, 2560)])
def test_layernorm_fwd_bwd(rows, C):
    from sd_lora_trainer_b200 import ops
    x = _rand(rows, C, seed=1) * 2 + 0.3
    gamma, beta = (_rand(C, seed=2) * 0.2 + 1).to(BF), _rand(C, seed=3, scale=0.2)
    dy = _rand(rows, C, seed=4)
    y, stats = ops.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gamma.float(), beta.float(), 1e-5)
    _close(y, ref, what="ln fwd")
    ref.backward(dy.float())
    dx = ops.layernorm_bwd(dy, x, gamma, stats)
    _close(dx, xr.grad, tol=3e-2, what="ln bwd")


def test_geglu_silu_add():
    from sd_lora_trainer_b200 import ops
    h = _rand(333, 2 * 640, seed=1)
    dy = _rand(333, 640, seed=2)
    hr = h.float().requires_grad_(True)
    a, g = hr.chunk(2, -1)
    ref = a * F.gelu(g)
    _close(ops.geglu_fwd(h), ref, what="geglu fwd")
    ref.backward(dy.float())
    _close(ops.geglu_bwd(dy, h), hr.grad, what="geglu bwd")
    x = _rand(1000, seed=3)
    xr = x.float().requires_grad_(True)
    s = F.silu(xr)
    _close(ops.silu_fwd(x), s, what="silu")
    d = _rand(1000, seed=4)
    s.backward(d.float())
    _close(ops.silu_bwd(d, x), xr.grad, what="silu bwd")
    p, q, r = _rand(1237, seed=5), _rand(1237, seed=6), _rand(1237, seed=7)
    _close(ops.add(p, q), p.float() + q.float(), what="add2")
    _close(ops.add(p, q, r), p.float() + q.float() + r.float(), what="add3")


def test_layout_helpers():
    from sd_lora_trainer_b200 import ops
    N, H, W, C = 2, 6, 8, 16
    x = _rand(N * H * W, C, seed=1)
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    up = ops.upsample2x_fwd(x, N, H, W, C)
    ref = F.interpolate(xn, scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).reshape(-1, C)
    assert torch.equal(up.float(), ref)
    dy = _rand(N * 4 * H * W, C, seed=2)
    dref = dy.float().view(N, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(-1, C)
    _close(ops.upsample2x_bwd(dy, N, H, W, C), dref, tol=1e-2, what="upsample bwd")
    for stride in (1, 2):
        col = ops.im2col3x3(x, N, H, W, C, stride)
        unf = F.unfold(xn, 3, padding=1, stride=stride)                      # [N, C*9, L] with (c, kh, kw) ordering
        L = unf.shape[-1]
        ref = unf.view(N, C, 9, L).permute(0, 3, 2, 1).reshape(N * L, 9 * C)
        assert torch.equal(col.float(), ref), f"im2col stride {stride}"
        dcol = _rand(*col.shape, seed=3)
        dref = F.fold(dcol.float().view(N, L, 9, C).permute(0, 3, 2, 1).reshape(N, C * 9, L), (H, W), 3, padding=1,
                      stride=stride).permute(0, 2, 3, 1).reshape(-1, C)
        _close(ops.col2im3x3(dcol, N, H, W, C, stride), dref, tol=1e-2, what=f"col2im stride {stride}")
    r = 8
    U = _rand(N * H * W, r, seed=4)
    U9 = ops.shift_stack9(U, N, H, W, r).float().view(N, H, W, 9, r)
    Un = U.float().view(N, H, W, r)
    for tap in range(9):
        dh, dw = tap // 3 - 1, tap % 3 - 1
        ref = torch.zeros_like(Un)
        hs = slice(max(dh, 0), H + min(dh, 0))
        ws = slice(max(dw, 0), W + min(dw, 0))
        hsrc = slice(max(-dh, 0), H + min(-dh, 0))
        wsrc = slice(max(-dw, 0), W + min(-dw, 0))
        ref[:, hs, ws] = Un[:, hsrc, wsrc]
        assert torch.equal(U9[:, :, :, tap], ref), f"shift tap {tap}"
    t = torch.tensor([0.0, 1.0, 500.0, 999.0], device="cuda")
    emb = ops.timestep_embedding(t, 320)
    half = 160
    ex = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half
    e = t[:, None] * torch.exp(ex)[None]
    ref = torch.cat([torch.cos(e), torch.sin(e)], -1)
    _close(emb, ref, tol=1e-2, what="timestep embedding")


def test_prologue_loss():
    from sd_lora_trainer_b200 import ops
    B, C, H, W = 3, 4, 16, 16
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(B, C, H, W, device="cuda", generator=g) * 0.13
    noise = torch.randn(B, C, H, W, device="cuda", generator=g).to(BF)
    off = torch.randn(B, C, 1, 1, device="cuda", generator=g)
    t = torch.tensor([5, 500, 990], device="cuda")
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    acp = torch.cumprod(1 - betas, 0).cuda()
    # torch restatement of main.py:311-326 (op-by-op bf16)
    n_ref = noise.clone()
    n_ref += 0.02 * off
    a = acp.to(BF)
    sa = (a[t] ** 0.5).view(B, 1, 1, 1)
    so = ((1 - a[t]) ** 0.5).view(B, 1, 1, 1)
    noisy_ref = sa * lat.to(BF) + so * n_ref
    nz = noise.clone()
    noisy, nhwc8 = ops.noise_prologue(lat, nz, off.view(B, C).contiguous(), 0.02, acp, t)
    assert torch.equal(nz, n_ref), "offset noise"
    assert torch.equal(noisy, noisy_ref), "add_noise bit-exact"
    assert torch.equal(nhwc8[:, :4].view(B, H, W, 4).permute(0, 3, 1, 2), noisy_ref)
    # loss
    pred = torch.randn(B * H * W, 8, device="cuda", generator=g).to(BF)
    mask = torch.rand(B, C, H, W, device="cuda", generator=g)
    w = ops.snr_weights(acp, t, 5.0)
    snr = (acp[t] ** 0.5 / (1 - acp[t]) ** 0.5) ** 2
    bw = torch.minimum(snr, torch.full_like(snr, 5.0)) / snr
    _close(w, bw / bw.mean(), tol=1e-5, what="snr weights")
    pr = pred[:, :4].reshape(B, H, W, 4).permute(0, 3, 1, 2).clone().requires_grad_(True)
    l = ((pr - nz).pow(2) * mask).mean(dim=[1, 2, 3]) * (bw / bw.mean())
    l = l.mean()
    l.backward()
    loss, dpred = ops.diffusion_loss(pred, 8, nz, mask, w, 1.0)
    _close(loss, l.detach().view(1), tol=1e-3, what="loss")
    _close(dpred[:, :4].view(B, H, W, 4).permute(0, 3, 1, 2), pr.grad, tol=2e-2, what="dpred")
    out = torch.zeros(1, device="cuda")
    ops.abs_sum(pred, out)
    _close(out, pred.float().abs().sum().view(1), tol=1e-4, what="abs_sum")


@pytest.mark.parametrize("wd,l1", [(0.004, 0.0), (0.0, 0.0), (0.004, 2.0 ** -10)])
def test_adamw_bit_exact_vs_torch(wd, l1):
    """The fused AdamW reproduces torch.optim.AdamW on bf16 tensors bit for bit (equal bf16 grads)."""
    from sd_lora_trainer_b200 import ops
    n = 100_003
    g = torch.Generator(device="cuda").manual_seed(1)
    p0 = (torch.randn(n, device="cuda", generator=g) * 0.06).to(BF)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=3e-4, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        grad = (torch.randn(n, device="cuda", generator=g) * 1e-3).to(BF)
        gref = grad.clone()
        if l1:
            gref = gref + l1 * torch.sign(p_ref.detach())
        p_ref.grad = gref
        opt.step()
        g32 = grad.float()
        ops.adamw(p, g32, m, v, n, lr=3e-4, wd=wd, l1_coeff=l1, lr2=0.0, wd2=0.0, step=step)
        assert float(g32.abs().max()) == 0.0
        diff = (p.float() - p_ref.detach().float()).abs()
        assert torch.equal(p, p_ref.detach()), f"step {step}: {int((diff > 0).sum())} of {n} differ, max {float(diff.max())}"


@pytest.mark.parametrize("B,hw,C", [(2, 1024, 1280), (2, 4096, 640), (2, 16384, 320), (4, 64, 1280), (1, 100, 8)])
def test_colsum(B, hw, C):
    from sd_lora_trainer_b200 import ops
    x = _rand(B * hw, C, seed=5)
    out = ops.colsum(x, B, hw, C)
    ref = x.float().view(B, hw, C).sum(dim=1)
    _close(out, ref, tol=1e-2, what="colsum")


@pytest.mark.parametrize("B,Hi,Wi,Ho,Wo,C,ld", [(2, 64, 64, 32, 32, 77, 80), (4, 64, 64, 8, 8, 77, 80),
                                                (2, 32, 32, 8, 8, 77, 77), (1, 16, 24, 8, 12, 5, 8),
                                                (2, 8, 8, 16, 16, 7, 7)])
def test_bicubic_resize_matches_interpolate(B, Hi, Wi, Ho, Wo, C, ld):
    """ops.bicubic_fwd/bwd against F.interpolate(mode="bicubic") in fp32 (ti_cross_attn_loss.py:262-266), on the
    strided channels-last views the score hook produces ([B, L, 80] sliced to 77 text tokens)."""
    from sd_lora_trainer_b200 import ops
    buf = _rand(B, Hi * Wi, ld, seed=6)
    x = buf[:, :, :C].reshape(B, Hi, Wi, C)
    y = ops.bicubic_fwd(x, Ho, Wo)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.interpolate(xr, size=(Ho, Wo), mode="bicubic")
    refl = ref.permute(0, 2, 3, 1)
    # one bf16 rounding of an fp32 result: within half an ulp of the fp32 reference (2^-9 relative) + summation slack
    assert ((y.float() - refl).abs() <= 2.0 ** -8 * refl.abs() + 1e-6).all()
    dy = _rand(B, Ho, Wo, C, seed=7)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dx = ops.bicubic_bwd(dy, Hi, Wi)
    gref = xr.grad.permute(0, 2, 3, 1)
    assert ((dx.float() - gref).abs() <= 2.0 ** -8 * gref.abs() + 1e-5).all()
    # adjoint identity <J x, dy> == <x, J^T dy> in fp64 on the fp32 references
    lhs = (refl.double() * dy.double()).sum()
    rhs = (x.double() * gref.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * abs(float(lhs)) + 1e-6
