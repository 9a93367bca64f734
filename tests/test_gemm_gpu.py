"""GPU parity of the tcgen05 GEMM (through the C ABI) against a plain torch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(BF)


def _close(out, ref, tol=2e-2, what=""):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"{what}: max abs err {err:.4g} vs max ref {den:.4g}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (200, 72, 88), (2048, 1280, 1280), (77, 640, 2048),
                                   (4096, 10240, 1280), (2048, 16, 1280), (130, 264, 8)])
def test_plain_kmajor(M, N, K):
    from sd_lora_trainer_b200 import ops
    a, b = _rand(M, K), _rand(N, K, seed=1)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)])
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().T, what=f"plain {M}x{N}x{K}")


@pytest.mark.parametrize("bn", [16, 32, 48, 64, 80, 128, 144, 208, 256])
def test_block_n_sweep(bn):
    from sd_lora_trainer_b200 import ops
    M, N, K = 384, 528, 192
    a, b = _rand(M, K), _rand(N, K, seed=1)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], block_n=bn)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().T, what=f"bn={bn}")


def test_epilogue_alpha_bias_residual_fp32():
    from sd_lora_trainer_b200 import ops
    M, N, K = 300, 200, 136
    a, b, bias, res = _rand(M, K), _rand(N, K, seed=1), _rand(N, seed=2), _rand(M, N, seed=3)
    ref = 0.5 * (a.float() @ b.float().T) + bias.float() + res.float()
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], alpha=0.5, bias=bias, residual=res)
    _close(out, ref, what="bf16 epilogue")
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(out32, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], alpha=0.5, bias=bias, residual=res)
    _close(out32, ref, tol=5e-3, what="fp32 epilogue")
    # per-image bias (conv1 + time embedding): rows grouped by 100
    bias2 = _rand(3, N, seed=4)
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], bias=bias2, bias_rows=100, bias_sb=N)
    ref2 = a.float() @ b.float().T + bias2.float().repeat_interleave(100, 0)
    _close(out, ref2, what="row-batched bias")


@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (2048, 1280, 1280), (200, 328, 72), (128, 64, 16)])
def test_b_mn_major_dgrad(M, N, K):
    """dX[M, N] = dY[M, K] . W[K, N]  (W row-major [K, N] read as the MN-major B operand)."""
    from sd_lora_trainer_b200 import ops
    dy, w = _rand(M, K), _rand(K, N, seed=1)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(dy), ops.mnmajor(w), K)])
    _close(out, dy.float() @ w.float(), what=f"B mn-major {M}x{N}x{K}")


@pytest.mark.parametrize("Mred,Kout,r,splits", [(256, 128, 16, 1), (2048, 1280, 16, 8), (1000, 320, 8, 3), (4096, 640, 32, 4)])
def test_wgrad_both_mn_major_splitk(Mred, Kout, r, splits):
    """dA^T[Kout, r] = X^T . U  (both operands MN-major, split-K fp32 atomics, transposed store into dA[r, Kout])."""
    from sd_lora_trainer_b200 import ops
    x, u = _rand(Mred, Kout), _rand(Mred, r, seed=1)
    dA = torch.zeros(r, Kout, dtype=torch.float32, device="cuda")
    ops.gemm(dA, Kout, r, [(ops.mnmajor(x), ops.mnmajor(u), Mred)], d_strides=(1, Kout, 0, 0), splits=splits, atomic=True)
    _close(dA, u.float().T @ x.float(), tol=5e-3, what="wgrad")
    # accumulate a second time: grads += ...
    ops.gemm(dA, Kout, r, [(ops.mnmajor(x), ops.mnmajor(u), Mred)], d_strides=(1, Kout, 0, 0), splits=splits, atomic=True)
    _close(dA, 2 * (u.float().T @ x.float()), tol=5e-3, what="wgrad accumulate")


@pytest.mark.parametrize("M,N,K,r", [(256, 256, 128, 16), (2048, 1280, 1280, 16), (300, 640, 320, 8), (2048, 640, 2048, 32)])
def test_two_segment_lora_fwd(M, N, K, r):
    """Y = X.W^T + T.B^T in ONE kernel (T = s.X.A^T computed by a skinny call of the same kernel)."""
    from sd_lora_trainer_b200 import ops
    x, w = _rand(M, K), _rand(N, K, seed=1, scale=0.05)
    A, Bm = _rand(r, K, seed=2, scale=0.1), _rand(N, r, seed=3, scale=0.1)
    T = torch.empty(M, r, dtype=BF, device="cuda")
    ops.gemm(T, M, r, [(ops.kmajor(x), ops.kmajor(A), K)])
    _close(T, x.float() @ A.float().T, what="T")
    y = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K), (ops.kmajor(T), ops.kmajor(Bm), r)])
    ref = x.float() @ w.float().T + T.float() @ Bm.float().T
    _close(y, ref, what="fused lora fwd")


def test_two_segment_lora_dgrad():
    """dX = dY.W + U.A  with W [N, K] and A [r, K] both read MN-major."""
    from sd_lora_trainer_b200 import ops
    M, N, K, r = 512, 640, 320, 16
    dy, w = _rand(M, N), _rand(N, K, seed=1, scale=0.05)
    A, Bm = _rand(r, K, seed=2, scale=0.1), _rand(N, r, seed=3, scale=0.1)
    U = torch.empty(M, r, dtype=BF, device="cuda")
    ops.gemm(U, M, r, [(ops.kmajor(dy), ops.mnmajor(Bm), N)])
    _close(U, dy.float() @ Bm.float(), what="U")
    dx = torch.empty(M, K, dtype=BF, device="cuda")
    ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N), (ops.kmajor(U), ops.mnmajor(A), r)])
    _close(dx, dy.float() @ w.float() + U.float() @ A.float(), what="fused lora dgrad")


@pytest.mark.parametrize("B,H,L,Lk,d", [(2, 4, 256, 256, 64), (2, 3, 200, 77, 40), (1, 2, 1024, 1024, 64), (2, 2, 64, 77, 160)])
def test_batched_attention_gemms(B, H, L, Lk, d):
    from sd_lora_trainer_b200 import ops
    Cc = H * d
    q, k, v = _rand(B * L, Cc), _rand(B * Lk, Cc, seed=1), _rand(B * Lk, Cc, seed=2)
    Lkp = (Lk + 7) // 8 * 8
    S = torch.zeros(B, H, L, Lkp, dtype=torch.float32, device="cuda")
    qm = ops.Mat(q, L, d, Cc, mn=False, sb0=d, sb1=L * Cc, batched=True)
    km = ops.Mat(k, Lk, d, Cc, mn=False, sb0=d, sb1=Lk * Cc, batched=True)
    ops.gemm(S, L, Lk, [(qm, km, d)], d_strides=(Lkp, 1, L * Lkp, H * L * Lkp), alpha=d ** -0.5, nb0=H, nb1=B)
    q4 = q.float().view(B, L, H, d).transpose(1, 2)
    k4 = k.float().view(B, Lk, H, d).transpose(1, 2)
    v4 = v.float().view(B, Lk, H, d).transpose(1, 2)
    Sref = q4 @ k4.transpose(-1, -2) * d ** -0.5
    _close(S[..., :Lk], Sref, tol=5e-3, what="S = QK^T")
    P = torch.zeros(B, H, L, Lkp, dtype=BF, device="cuda")
    ops.softmax_fwd(S, P, B * H * L, Lk, Lkp, Lkp)
    Pref = torch.softmax(Sref, -1)
    _close(P[..., :Lk], Pref, what="softmax")
    assert float(P[..., Lk:].abs().max()) == 0.0 if Lkp > Lk else True
    # O[b, l, h, :] = P . V  with V read MN-major straight out of the [B*Lk, C] projection output
    O = torch.empty(B * L, Cc, dtype=BF, device="cuda")
    pm = ops.Mat(P, L, Lk, Lkp, mn=False, sb0=L * Lkp, sb1=H * L * Lkp, batched=True)
    vm = ops.Mat(v, Lk, d, Cc, mn=True, sb0=d, sb1=Lk * Cc, batched=True)
    ops.gemm(O, L, d, [(pm, vm, Lk)], d_strides=(Cc, 1, d, L * Cc), nb0=H, nb1=B)
    Oref = (P[..., :Lk].float() @ v4).transpose(1, 2).reshape(B * L, Cc)
    _close(O, Oref, what="O = PV")
    # backward pieces: dV = P^T dO (A mn-major, B mn-major), dP = dO V^T
    dO = _rand(B * L, Cc, seed=5)
    dV = torch.empty(B * Lk, Cc, dtype=BF, device="cuda")
    pTm = ops.Mat(P, L, Lk, Lkp, mn=True, sb0=L * Lkp, sb1=H * L * Lkp, batched=True)
    dOm = ops.Mat(dO, L, d, Cc, mn=True, sb0=d, sb1=L * Cc, batched=True)
    ops.gemm(dV, Lk, d, [(pTm, dOm, L)], d_strides=(Cc, 1, d, Lk * Cc), nb0=H, nb1=B)
    dO4 = dO.float().view(B, L, H, d).transpose(1, 2)
    dVref = (P[..., :Lk].float().transpose(-1, -2) @ dO4).transpose(1, 2).reshape(B * Lk, Cc)
    _close(dV, dVref, what="dV = P^T dO")
    dP = torch.zeros(B, H, L, Lkp, dtype=torch.float32, device="cuda")
    dOk = ops.Mat(dO, L, d, Cc, mn=False, sb0=d, sb1=L * Cc, batched=True)
    vk = ops.Mat(v, Lk, d, Cc, mn=False, sb0=d, sb1=Lk * Cc, batched=True)
    ops.gemm(dP, L, Lk, [(dOk, vk, d)], d_strides=(Lkp, 1, L * Lkp, H * L * Lkp), nb0=H, nb1=B)
    dPref = dO4 @ v4.transpose(-1, -2)
    _close(dP[..., :Lk], dPref, tol=5e-3, what="dP")
    dS = torch.empty_like(P)
    ops.softmax_bwd(P, dP, dS, B * H * L, Lk, Lkp, Lkp)
    Pf = P[..., :Lk].float()
    dSref = Pf * (dPref - (Pf * dPref).sum(-1, keepdim=True))
    _close(dS[..., :Lk], dSref, what="softmax bwd")


@pytest.mark.parametrize("N,H,W,C,Cout", [(2, 16, 16, 64, 128), (2, 8, 8, 128, 64), (1, 32, 32, 192, 96), (2, 128, 128, 64, 32),
                                          (3, 4, 4, 64, 64), (2, 64, 64, 8, 320), (2, 16, 16, 320, 4)])
def test_implicit_conv3x3(N, H, W, C, Cout):
    from sd_lora_trainer_b200 import ops
    x = _rand(N * H * W, C)
    w = _rand(Cout, C, 3, 3, seed=1, scale=0.05)
    bias = _rand(Cout, seed=2)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()          # [Cout, (kh, kw, c)]
    out = torch.empty(N * H * W, Cout, dtype=BF, device="cuda")
    ops.gemm(out, N * H * W, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C)], bias=bias)
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xn, w.float(), bias.float(), padding=1).permute(0, 2, 3, 1).reshape(N * H * W, Cout)
    _close(out, ref, what=f"conv {N}x{H}x{W}x{C}->{Cout}")


def test_conv_lora_fused_and_tapped_b():
    """conv2(x) + B.(A * x): segment 0 implicit conv, segment 1 the rank-r side path; T = conv_A(x) uses b_tap_n."""
    from sd_lora_trainer_b200 import ops
    N, H, W, C, Cout, r = 2, 16, 16, 128, 128, 16
    x = _rand(N * H * W, C)
    w = _rand(Cout, C, 3, 3, seed=1, scale=0.05)
    A = _rand(r, C, 3, 3, seed=2, scale=0.1)
    Bm = _rand(Cout, r, seed=3, scale=0.1)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()
    Ak = A.permute(2, 3, 0, 1).reshape(9 * r, C).contiguous()            # [(kh, kw, r), C]  master layout
    T = torch.empty(N * H * W, r, dtype=BF, device="cuda")
    ops.gemm(T, N * H * W, r, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=0, b_tap_n=r), ops.kmajor(Ak), 9 * C)])
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    Tref = F.conv2d(xn, A.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, r)
    _close(T, Tref, what="T = conv_A(x)")
    y = torch.empty(N * H * W, Cout, dtype=BF, device="cuda")
    ops.gemm(y, N * H * W, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C),
                                  (ops.kmajor(T), ops.kmajor(Bm), r)])
    ref = F.conv2d(xn, w.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout) + T.float() @ Bm.float().T
    _close(y, ref, what="conv + lora")


@pytest.mark.parametrize("M,N,K,r", [(256, 128, 128, 16), (2048, 1280, 1280, 16), (300, 640, 320, 8), (2048, 640, 2048, 32),
                                     (154, 1280, 2048, 16), (1000, 200, 136, 4), (8192, 640, 640, 16)])
def test_fused_side_path_forward(M, N, K, r):
    """Y = X.W^T + (s.X.A^T).B^T as ONE launch (rank-r product accumulated in TMEM, staged through smem)."""
    from sd_lora_trainer_b200 import ops
    rs = (r + 7) // 8 * 8
    x, w = _rand(M, K), _rand(N, K, seed=1, scale=0.05)
    A = _rand(r, K, seed=2, scale=0.1)
    Bp = torch.zeros(N, rs, dtype=BF, device="cuda")
    Bp[:, :r] = _rand(N, r, seed=3, scale=0.1)
    bias, res = _rand(N, seed=4), _rand(M, N, seed=5)
    T = torch.full((M, rs), 7.0, dtype=BF, device="cuda")
    y = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], bias=bias, residual=res,
             side=(ops.Mat(A, r, K, K), ops.Mat(Bp, N, r, rs), r, 0.5, T))
    Tref = (0.5 * (x.float() @ A.float().T)).to(BF)
    _close(T[:, :r], Tref, what="T out")
    ref = x.float() @ w.float().T + Tref.float() @ Bp[:, :r].float().T + bias.float() + res.float()
    _close(y, ref, what="fused side fwd")


@pytest.mark.parametrize("M,N,K,r", [(512, 640, 320, 16), (2048, 1280, 1280, 16), (154, 2048, 1280, 16), (300, 264, 200, 8),
                                     (2048, 1280, 1280, 32), (333, 128, 96, 4)])
def test_fused_side_path_dgrad(M, N, K, r):
    """dX[M, K] = dY.W + (s.dY.B).A as ONE launch; every operand read MN-major from its forward layout."""
    from sd_lora_trainer_b200 import ops
    rs = (r + 7) // 8 * 8
    dy, w = _rand(M, N), _rand(N, K, seed=1, scale=0.05)
    A = _rand(r, K, seed=2, scale=0.1)
    Bp = torch.zeros(N, rs, dtype=BF, device="cuda")
    Bp[:, :r] = _rand(N, r, seed=3, scale=0.1)
    acc = _rand(M, K, seed=6)
    U = torch.empty(M, rs, dtype=BF, device="cuda")
    dx = acc.clone()
    ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)], residual=dx,
             side=(ops.Mat(Bp, N, r, rs, mn=True), ops.Mat(A, r, K, K, mn=True), r, 2.0, U))
    Uref = (2.0 * (dy.float() @ Bp[:, :r].float())).to(BF)
    _close(U[:, :r], Uref, what="U out")
    _close(dx, dy.float() @ w.float() + Uref.float() @ A.float() + acc.float(), what="fused side dgrad")


@pytest.mark.parametrize("Mred,C,r", [(2048, 1280, 16), (8192, 640, 16), (1000, 320, 8), (2048, 1280, 32), (300, 200, 4)])
def test_grouped_weight_gradients(Mred, C, r):
    """dB[C, r] += dY^T.T and dA[r, C] += U^T.X of one square LoRA layer as ONE grouped launch (split-K, fp32 atomics)."""
    from sd_lora_trainer_b200 import ops
    rs = (r + 7) // 8 * 8
    dy, x = _rand(Mred, C), _rand(Mred, C, seed=1)
    T, U = _rand(Mred, rs, seed=2, scale=0.3), _rand(Mred, rs, seed=3, scale=0.3)
    gB = torch.zeros(C, rs, dtype=torch.float32, device="cuda")
    gA = torch.zeros(r, C, dtype=torch.float32, device="cuda")
    gB += 1.0                                                       # accumulation, not overwrite
    tiles = (C + 127) // 128
    splits = max(1, min(148 // (2 * tiles), (Mred + 63) // 64, 32))
    ops.gemm(gB, C, r, [(ops.Mat(dy, Mred, C, C, mn=True), ops.Mat(T, Mred, r, rs, mn=True), Mred),
                        (ops.Mat(x, Mred, C, C, mn=True), ops.Mat(U, Mred, r, rs, mn=True), Mred)],
             d_strides=(rs, 1, 0, 0), atomic=True, splits=splits, group_out=(gA, (1, C)))
    torch.cuda.synchronize()
    _close(gB[:, :r], 1.0 + dy.float().T @ T[:, :r].float(), tol=5e-3, what="grouped dB")
    _close(gA, (x.float().T @ U[:, :r].float()).T, tol=5e-3, what="grouped dA")
