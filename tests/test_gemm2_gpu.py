"""GPU parity of the CTA-pair (cta_group::2, 256-row tiles) GEMM against a plain torch fp32 reference, and against the
single-CTA kernel on the same inputs.  Every case forces pair_mode=1 so the pair kernel is the one that runs."""
import pytest
import torch

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


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (256, 256, 128), (300, 200, 136), (2048, 1280, 1280), (2048, 10240, 1280),
                                   (8192, 640, 640), (4096, 5120, 640), (257, 72, 1000), (1000, 1000, 72)])
def test_pair_plain_kmajor(M, N, K):
    from sd_lora_trainer_b200 import ops
    a, b = _rand(M, K), _rand(N, K, seed=1)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], pair_mode=1)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().T, what=f"pair plain {M}x{N}x{K}")
    # bit-level agreement with the single-CTA kernel is not required (different summation grouping), closeness is
    ref1 = torch.empty_like(out)
    ops.gemm(ref1, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], pair_mode=-1)
    _close(out, ref1, tol=1e-2, what="pair vs single-CTA")


@pytest.mark.parametrize("bn", [32, 64, 96, 128, 160, 192, 224, 256])
def test_pair_block_n_sweep(bn):
    from sd_lora_trainer_b200 import ops
    M, N, K = 640, 528, 192
    a, b = _rand(M, K), _rand(N, K, seed=1)
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], block_n=bn, pair_mode=1)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().T, what=f"pair bn={bn}")


def test_pair_epilogue_alpha_bias_residual_fp32():
    from sd_lora_trainer_b200 import ops
    M, N, K = 300, 200, 136
    a, b, bias, res = _rand(M, K), _rand(N, K, seed=1), _rand(N, seed=2), _rand(M, N, seed=3)
    ref = 0.5 * (a.float() @ b.float().T) + bias.float() + res.float()
    out = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], alpha=0.5, bias=bias, residual=res, pair_mode=1)
    _close(out, ref, what="pair bf16 epilogue")
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(out32, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], alpha=0.5, bias=bias, residual=res, pair_mode=1)
    _close(out32, ref, tol=5e-3, what="pair fp32 epilogue")
    bias2 = _rand(3, N, seed=4)
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.kmajor(b), K)], bias=bias2, bias_rows=100, bias_sb=N, pair_mode=1)
    ref2 = a.float() @ b.float().T + bias2.float().repeat_interleave(100, 0)
    _close(out, ref2, what="pair row-batched bias")


@pytest.mark.parametrize("M,N,K", [(512, 640, 320), (2048, 1280, 1280), (2048, 1280, 10240), (300, 264, 200), (8192, 640, 5120)])
def test_pair_b_mn_major_dgrad(M, N, K):
    """dX[M, K_out] = dY[M, N] . W[N, K_out]: the frozen weight read MN-major (forward layout)."""
    from sd_lora_trainer_b200 import ops
    dy, w = _rand(M, N), _rand(N, K, seed=1, scale=0.05)
    dx = torch.empty(M, K, dtype=BF, device="cuda")
    ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)], pair_mode=1)
    _close(dx, dy.float() @ w.float(), what=f"pair dgrad {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K,r", [(256, 128, 128, 16), (2048, 1280, 1280, 16), (300, 640, 320, 8), (2048, 640, 2048, 32),
                                     (1000, 200, 136, 4), (8192, 640, 640, 16), (4096, 2560, 640, 16),
                                     (2048, 3840, 1280, 48), (8192, 1920, 640, 48), (512, 384, 128, 24), (300, 264, 200, 64),
                                     (2048, 3840, 1280, 40)])
def test_pair_fused_side_path_forward(M, N, K, r):
    """Y = X.W^T + (s.X.A^T).B^T as ONE launch on CTA pairs (incl. the single-accumulator BN = 256 configuration)."""
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
             side=(ops.Mat(A, r, K, K), ops.Mat(Bp, N, r, rs), r, 0.5, T), pair_mode=1, static_b=(r != 8))
    Tref = (0.5 * (x.float() @ A.float().T)).to(BF)
    _close(T[:, :r], Tref, what="pair T out")
    ref = x.float() @ w.float().T + Tref.float() @ Bp[:, :r].float().T + bias.float() + res.float()
    _close(y, ref, what="pair fused side fwd")


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_pair_fused_side_block_n(bn):
    from sd_lora_trainer_b200 import ops
    M, N, K, r = 1024, 768, 256, 16
    x, w = _rand(M, K), _rand(N, K, seed=1, scale=0.05)
    A, Bm = _rand(r, K, seed=2, scale=0.1), _rand(N, r, seed=3, scale=0.1)
    T = torch.empty(M, r, dtype=BF, device="cuda")
    y = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], side=(ops.Mat(A, r, K, K), ops.Mat(Bm, N, r, r), r, 1.0, T),
             block_n=bn, pair_mode=1)
    Tref = (x.float() @ A.float().T).to(BF)
    _close(T, Tref, what=f"pair T bn={bn}")
    _close(y, x.float() @ w.float().T + Tref.float() @ Bm.float().T, what=f"pair side bn={bn}")


@pytest.mark.parametrize("M,N,K,r", [(2048, 3840, 1280, 48), (8192, 1920, 640, 48), (512, 384, 128, 24), (300, 264, 200, 64)])
def test_pair_wide_side_path_dgrad_kmajor_s(M, N, K, r):
    """The fused q|k|v input gradient: dX[M, K] = dY[M, N=3C].W + (s.dY.Bt^T).A with the side operand S = Bt [r, N] K-major
    (the packed transpose of the block-diagonal LoRA-B) and rank up to 64."""
    from sd_lora_trainer_b200 import ops
    dy, w = _rand(M, N), _rand(N, K, seed=1, scale=0.05)
    A = _rand(r, K, seed=2, scale=0.1)
    Bt = _rand(r, N, seed=3, scale=0.1)
    U = torch.empty(M, r, dtype=BF, device="cuda")
    dx = torch.empty(M, K, dtype=BF, device="cuda")
    ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)], side=(ops.Mat(Bt, r, N, N), ops.Mat(A, r, K, K, mn=True), r, 2.0, U),
             pair_mode=1, static_b=True)
    Uref = (2.0 * (dy.float() @ Bt.float().T)).to(BF)
    _close(U, Uref, what="pair wide U out")
    _close(dx, dy.float() @ w.float() + Uref.float() @ A.float(), what="pair wide side dgrad")


@pytest.mark.parametrize("M,N,K,r", [(512, 640, 320, 16), (2048, 1280, 1280, 16), (300, 264, 200, 8), (2048, 1280, 1280, 32),
                                     (333, 128, 96, 4), (8192, 640, 640, 16)])
def test_pair_fused_side_path_dgrad(M, N, K, r):
    """dX[M, K] = dY.W + (s.dY.B).A as ONE launch on CTA pairs; every operand read MN-major from its forward layout."""
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
             side=(ops.Mat(Bp, N, r, rs, mn=True), ops.Mat(A, r, K, K, mn=True), r, 2.0, U), pair_mode=1, static_b=(r == 16))
    Uref = (2.0 * (dy.float() @ Bp[:, :r].float())).to(BF)
    _close(U[:, :r], Uref, what="pair U out")
    _close(dx, dy.float() @ w.float() + Uref.float() @ A.float() + acc.float(), what="pair fused side dgrad")


@pytest.mark.parametrize("N,K,M", [(1280, 1280, 1024), (10240, 1280, 1024), (640, 2560, 4096), (304, 264, 200), (1280, 11520, 2048),
                                   (256, 128, 64), (320, 2880, 16384)])
def test_pair_dense_weight_gradient_mn_major_a_accumulates(N, K, M):
    """dW[N, K] += dY[M, N]^T . X[M, K] (full fine-tune): both operands MN-major from their forward layouts, fp32 output
    accumulated in place by the tile's one owner."""
    from sd_lora_trainer_b200 import ops
    dy, x = _rand(M, N, scale=0.5), _rand(M, K, seed=1, scale=0.5)
    g = torch.Generator(device="cuda").manual_seed(N + K)
    dw = torch.randn(N, K, device="cuda", generator=g)
    before = dw.clone()
    ops.gemm(dw, N, K, [(ops.Mat(dy, M, N, N, mn=True), ops.Mat(x, M, K, K, mn=True), M)], d_strides=(K, 1, 0, 0), atomic=True, pair_mode=1)
    torch.cuda.synchronize()
    ref = dy.float().T @ x.float()
    _close(dw - before, ref, tol=5e-3, what=f"pair dense wgrad {N}x{K}x{M}")
    ops.gemm(dw, N, K, [(ops.Mat(dy, M, N, N, mn=True), ops.Mat(x, M, K, K, mn=True), M)], d_strides=(K, 1, 0, 0), atomic=True, pair_mode=1)
    _close(dw - before, 2 * ref, tol=5e-3, what="second accumulation")


@pytest.mark.parametrize("M,N,K", [(2048, 5120, 1280), (8192, 2560, 640), (512, 512, 128), (300, 160, 200), (256, 64, 64)])
def test_pair_fused_geglu_backward_epilogue(M, N, K):
    """dh = geglu_bwd(dY . W, h) from the epilogue of the input-gradient GEMM of FeedForward.net.2 (W [K, N] read MN-major):
    same result as the GEMM followed by the stand-alone GEGLU backward kernel."""
    from sd_lora_trainer_b200 import ops
    dy, w = _rand(M, K, scale=0.5), _rand(K, N, seed=1, scale=0.05)
    h = _rand(M, 2 * N, seed=2)
    dh = torch.full((M, 2 * N), 7.0, dtype=BF, device="cuda")
    ops.gemm(dh, M, N, [(ops.kmajor(dy), ops.mnmajor(w), K)], geglu_h=h, pair_mode=1, static_b=True)
    mid = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(mid, M, N, [(ops.kmajor(dy), ops.mnmajor(w), K)], pair_mode=1)
    ref = ops.geglu_bwd(mid, h)
    torch.cuda.synchronize()
    _close(dh, ref, tol=1e-2, what=f"fused geglu bwd {M}x{N}x{K}")
    # and against plain torch on the fp32 product
    d32 = (dy.float() @ w.float()).to(BF).float()
    val, gate = h[:, :N].float(), h[:, N:].float()
    gl = torch.nn.functional.gelu(gate)
    cdf = 0.5 * (1 + torch.erf(gate / 2 ** 0.5))
    pdf = torch.exp(-0.5 * gate * gate) / (2 * 3.141592653589793) ** 0.5
    _close(dh[:, :N], d32 * gl, tol=2e-2, what="value half")
    _close(dh[:, N:], d32 * val * (cdf + gate * pdf), tol=2e-2, what="gate half")


def _interleave(t, il=128):
    """[value | gate] along dim 0 (rows of a weight / entries of a bias) -> blocks of il value rows alternating with il gate rows."""
    inner = t.shape[0] // 2
    rest = t.shape[1:]
    return torch.stack([t[:inner].reshape(inner // il, il, *rest), t[inner:].reshape(inner // il, il, *rest)], dim=1).reshape(t.shape).contiguous()


@pytest.mark.parametrize("M,N,K,bias", [(2048, 10240, 1280, True), (8192, 5120, 640, True), (512, 512, 128, False), (300, 256, 200, True)])
def test_pair_fused_geglu_forward_epilogue(M, N, K, bias):
    """FeedForward.net.0.proj with rows interleaved in blocks of 128 + GEGLU in the epilogue: h (interleaved layout) equals the
    plain projection re-ordered, y equals diffusers' GEGLU  value * gelu(gate)  of the rounded projection."""
    from sd_lora_trainer_b200 import ops
    x, w = _rand(M, K, scale=0.5), _rand(N, K, seed=1, scale=0.05)
    b = _rand(N, seed=2) if bias else None
    wi, bi = _interleave(w), (_interleave(b) if bias else None)
    h = torch.full((M, N), 7.0, dtype=BF, device="cuda")
    y = torch.full((M, N // 2), 7.0, dtype=BF, device="cuda")
    ops.gemm(h, M, N, [(ops.kmajor(x), ops.kmajor(wi), K)], bias=bi, geglu_out=y, pair_mode=1, static_b=True)
    plain = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(plain, M, N, [(ops.kmajor(x), ops.kmajor(w), K)], bias=b, pair_mode=1)
    torch.cuda.synchronize()
    inner = N // 2
    hv = h.view(M, inner // 128, 2, 128)
    _close(hv[:, :, 0].reshape(M, inner), plain[:, :inner], tol=1e-2, what="value half of h")
    _close(hv[:, :, 1].reshape(M, inner), plain[:, inner:], tol=1e-2, what="gate half of h")
    ref = hv[:, :, 0].reshape(M, inner).float() * torch.nn.functional.gelu(hv[:, :, 1].reshape(M, inner).float()).to(BF).float()
    _close(y, ref, tol=1e-2, what="fused GEGLU output")
    # the stand-alone kernels read the interleaved layout too
    _close(ops.geglu_fwd(h, 128), y, tol=1e-2, what="stand-alone geglu_fwd on the interleaved h")
    dy = _rand(M, inner, seed=5)
    dh_i = ops.geglu_bwd(dy, h, 128)
    h_plain = torch.cat([hv[:, :, 0].reshape(M, inner), hv[:, :, 1].reshape(M, inner)], dim=1).contiguous()
    dh_p = ops.geglu_bwd(dy, h_plain, 0)
    dv = dh_i.view(M, inner // 128, 2, 128)
    assert torch.equal(dv[:, :, 0].reshape(M, inner), dh_p[:, :inner]) and torch.equal(dv[:, :, 1].reshape(M, inner), dh_p[:, inner:])


def test_pair_back_to_back_launches_are_ordered():
    """Programmatic dependent launch + clusters: a chain of dependent GEMMs (each reads the previous output)."""
    from sd_lora_trainer_b200 import ops
    M, C = 2048, 1280
    x = _rand(M, C, scale=0.5)
    ws = [_rand(C, C, seed=10 + i, scale=0.03) for i in range(6)]
    cur = x
    ref = x.float()
    for w in ws:
        nxt = torch.empty(M, C, dtype=BF, device="cuda")
        ops.gemm(nxt, M, C, [(ops.kmajor(cur), ops.kmajor(w), C)], pair_mode=1, static_b=True)   # weights prefetched before the PDL wait
        ref = (ref @ w.float().T).to(BF).float()
        cur = nxt
    torch.cuda.synchronize()
    _close(cur, ref, tol=3e-2, what="pair chain")


@pytest.mark.parametrize("N,H,W,C,Cout", [(2, 16, 16, 64, 128), (2, 32, 32, 320, 320), (2, 128, 128, 64, 64), (2, 64, 64, 8, 320),
                                          (1, 32, 32, 192, 96), (4, 8, 8, 128, 640), (2, 32, 32, 1280, 1280)])
def test_pair_implicit_conv3x3(N, H, W, C, Cout):
    import torch.nn.functional as F
    from sd_lora_trainer_b200 import ops
    x = _rand(N * H * W, C)
    w = _rand(Cout, C, 3, 3, seed=1, scale=0.05)
    bias = _rand(Cout, seed=2)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()          # [Cout, (kh, kw, c)]
    out = torch.empty(N * H * W, Cout, dtype=BF, device="cuda")
    ops.gemm(out, N * H * W, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C)], bias=bias, pair_mode=1,
             static_b=(C % 128 == 0))
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xn, w.float(), bias.float(), padding=1).permute(0, 2, 3, 1).reshape(N * H * W, Cout)
    _close(out, ref, what=f"pair conv {N}x{H}x{W}x{C}->{Cout}")


def test_pair_conv_plus_lora_segment_and_dgrad_extra():
    """conv2(x) + T.B^T (segment 1 K-major) and the dgrad form  conv_T(dy) + U9.A  (segment 1 with an MN-major B)."""
    import torch.nn.functional as F
    from sd_lora_trainer_b200 import ops
    N, H, W, C, Cout, r = 2, 32, 32, 128, 256, 16
    M = N * H * W
    x = _rand(M, C)
    w = _rand(Cout, C, 3, 3, seed=1, scale=0.05)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()
    T, Bm = _rand(M, r, seed=2, scale=0.3), _rand(Cout, r, seed=3, scale=0.1)
    res = _rand(M, Cout, seed=4)
    y = torch.empty(M, Cout, dtype=BF, device="cuda")
    ops.gemm(y, M, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C), (ops.kmajor(T), ops.kmajor(Bm), r)],
             residual=res, pair_mode=1)
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xn, w.float(), padding=1).permute(0, 2, 3, 1).reshape(M, Cout) + T.float() @ Bm.float().T + res.float()
    _close(y, ref, what="pair conv + lora segment")
    # dgrad-like: segment 1 = U9 [M, 9r] (K-major) against A [9r, C] read MN-major
    U9, A = _rand(M, 9 * r, seed=5, scale=0.2), _rand(9 * r, Cout, seed=6, scale=0.1)
    z = torch.empty(M, Cout, dtype=BF, device="cuda")
    ops.gemm(z, M, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C),
                          (ops.kmajor(U9), ops.Mat(A, 9 * r, Cout, Cout, mn=True), 9 * r)], pair_mode=1)
    ref2 = F.conv2d(xn, w.float(), padding=1).permute(0, 2, 3, 1).reshape(M, Cout) + U9.float() @ A.float()
    _close(z, ref2, what="pair conv + MN-major segment")


def test_pair_two_plain_segments():
    from sd_lora_trainer_b200 import ops
    M, N, K, r = 1024, 640, 320, 16
    x, w = _rand(M, K), _rand(N, K, seed=1, scale=0.05)
    T, Bm = _rand(M, r, seed=2, scale=0.3), _rand(N, r, seed=3, scale=0.1)
    y = torch.empty(M, N, dtype=BF, device="cuda")
    ops.gemm(y, M, N, [(ops.kmajor(x), ops.kmajor(w), K), (ops.kmajor(T), ops.kmajor(Bm), r)], pair_mode=1)
    _close(y, x.float() @ w.float().T + T.float() @ Bm.float().T, what="pair two segments")


@pytest.mark.parametrize("M,N,K,r,pair", [(2048, 1280, 1280, 16, 1), (512, 640, 320, 16, 1), (300, 264, 200, 8, 1),
                                          (2048, 1280, 1280, 32, 1), (154, 2048, 1280, 16, -1), (2048, 1280, 1280, 16, -1)])
def test_fused_side_dgrad_with_kmajor_b_copy(M, N, K, r, pair):
    """The step's input-gradient form: W read MN-major, the side operand a K-major COPY of LoRA-B (ops.lora_transpose_b),
    LoRA-A read MN-major."""
    from sd_lora_trainer_b200 import ops
    rs = (r + 7) // 8 * 8
    dy, w = _rand(M, N), _rand(N, K, seed=1, scale=0.05)
    A = _rand(r, K, seed=2, scale=0.1)
    flat = torch.zeros(N * rs + 64, dtype=BF, device="cuda")
    Bp = flat[32:32 + N * rs].view(N, rs)
    Bp[:, :r] = _rand(N, r, seed=3, scale=0.1)
    bt_flat = torch.full((N * rs + 16,), 3.0, dtype=BF, device="cuda")
    table = torch.tensor([[32, 8, N, rs]], dtype=torch.int64, device="cuda")
    ops.lora_transpose_b(flat, bt_flat, table)
    Bt = bt_flat[8:8 + N * rs].view(rs, N)
    assert torch.equal(Bt, Bp.t())
    U = torch.empty(M, rs, dtype=BF, device="cuda")
    dx = torch.empty(M, K, dtype=BF, device="cuda")
    ops.gemm(dx, M, K, [(ops.kmajor(dy), ops.mnmajor(w), N)],
             side=(ops.Mat(Bt, r, N, N), ops.Mat(A, r, K, K, mn=True), r, 2.0, U), pair_mode=pair, static_b=True)
    Uref = (2.0 * (dy.float() @ Bp[:, :r].float())).to(BF)
    _close(U[:, :r], Uref, what="U out (K-major B copy)")
    _close(dx, dy.float() @ w.float() + Uref.float() @ A.float(), what="fused side dgrad (K-major B copy)")


@pytest.mark.parametrize("M,N,K,mn,res,bias", [(2048, 1280, 10240, True, False, False), (2048, 1280, 5120, False, True, True),
                                               (2048, 1280, 5120, False, "inplace", False), (1024, 512, 4096, False, False, True),
                                               (8192, 640, 5120, True, True, False), (512, 256, 2048, False, True, True)])
def test_pair_stream_k(M, N, K, mn, res, bias):
    """Few tiles + long K -> stream-K ranges with TMA reduce-add partial tiles (D pre-set to residual / zero)."""
    from sd_lora_trainer_b200 import ops
    a = _rand(M, K, scale=0.5)
    w = _rand(K, N, seed=1, scale=0.05) if mn else _rand(N, K, seed=1, scale=0.05)
    bvec = _rand(N, seed=2) if bias else None
    ref = a.float() @ (w.float() if mn else w.float().T)
    if bias:
        ref = ref + bvec.float()
    out = torch.full((M, N), 5.0, dtype=BF, device="cuda")
    R = None
    if res == "inplace":
        out = _rand(M, N, seed=3)
        R = out
        ref = ref + out.float()
    elif res:
        R = _rand(M, N, seed=3)
        ref = ref + R.float()
    ops.gemm(out, M, N, [(ops.kmajor(a), ops.mnmajor(w) if mn else ops.kmajor(w), K)], bias=bvec, residual=R, pair_mode=1)
    torch.cuda.synchronize()
    _close(out, ref, what=f"stream-K {M}x{N}x{K}")


def test_pair_stream_k_conv():
    import torch.nn.functional as F
    from sd_lora_trainer_b200 import ops
    N, H, W, C, Cout = 2, 32, 32, 1280, 1280
    x = _rand(N * H * W, C, scale=0.5)
    w = _rand(Cout, C, 3, 3, seed=1, scale=0.02)
    bias, res = _rand(Cout, seed=2), _rand(N * H * W, Cout, seed=3)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()
    out = torch.empty(N * H * W, Cout, dtype=BF, device="cuda")
    ops.gemm(out, N * H * W, Cout, [(ops.Conv3x3(x, N, H, W, C, b_tap_k=C), ops.kmajor(wk), 9 * C)], bias=bias, residual=res,
             pair_mode=1)
    xn = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xn, w.float(), bias.float(), padding=1).permute(0, 2, 3, 1).reshape(N * H * W, Cout) + res.float()
    _close(out, ref, what="stream-K conv")
