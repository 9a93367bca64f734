"""TEST INFRASTRUCTURE: a torch/CPU stand-in for sd_lora_trainer_b200.ops with the SAME call signatures, used only by
the `not gpu` tests to check the HOST logic of the executor (operand descriptors, strides, gradient plumbing, skip
bookkeeping) against the oracle without a GPU.  It interprets the exact Mat / Conv3x3 descriptors the product code
builds, so a wrong stride or transposition fails here before any GPU time is spent.  Never imported by the product."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from sd_lora_trainer_b200 import ops as real

BF16 = torch.bfloat16
Mat, Conv3x3, kmajor, mnmajor, conv_supported = real.Mat, real.Conv3x3, real.kmajor, real.mnmajor, real.conv_supported


def _bf(x):
    return x.to(BF16).contiguous()


def _view(t: torch.Tensor, shape, strides, extra_off=0):
    need = extra_off + sum((s - 1) * st for s, st in zip(shape, strides)) + 1
    flat = t.as_strided((t.untyped_storage().nbytes() // t.element_size(),), (1,), 0)
    assert t.storage_offset() + need <= flat.numel(), "descriptor reaches outside the tensor's storage"
    return flat.as_strided(shape, strides, t.storage_offset() + extra_off)


def _operand(m: Mat, b0: int, b1: int, mn_extent: int, k_extent: int) -> torch.Tensor:
    """Returns the logical [MN, K] fp32 matrix with TMA-style zero fill outside (rows, inner)."""
    off = (b0 * m.sb0 + b1 * m.sb1) if m.batched else 0
    v = _view(m.t, (m.rows, m.inner), (m.row_stride, 1), off).float()
    if m.mn:
        v = v.t()                                   # stored [K, MN]
    out = torch.zeros(mn_extent, k_extent)
    r, c = min(mn_extent, v.shape[0]), min(k_extent, v.shape[1])
    out[:r, :c] = v[:r, :c]
    return out


def _chk_tma(m: Mat, what: str):
    """The constraints gemm_host.cu's tensor-map encoder enforces (encode_map_ex): bf16, 16-byte aligned base, every
    stride a multiple of 16 bytes.  CPU allocations are 64-byte aligned, so the view's offset decides alignment here as
    it does on the device."""
    assert m.t.dtype == BF16, f"{what}: operand must be bf16"
    assert (m.t.storage_offset() * 2) % 16 == 0, f"{what}: base not 16-byte aligned (offset {m.t.storage_offset()})"
    for name, st in (("row", m.row_stride),) + ((("sb0", m.sb0), ("sb1", m.sb1)) if m.batched else ()):
        st = st if st > 0 else 8
        assert (st * 2) % 16 == 0, f"{what}: {name} stride {st} elements is not a multiple of 16 bytes"


def _chk_gemm(out, M, N, segs, nb0, nb1, splits, atomic, side, group_out, pair_mode=0):
    """Argument rules of b200_gemm (gemm_host.cu) that are not implied by the arithmetic."""
    assert M >= 1 and N >= 1 and len(segs) in (1, 2) and nb0 >= 1 and nb1 >= 1 and splits >= 1
    assert out.dtype in (torch.float32, BF16)
    assert not atomic or out.dtype == torch.float32, "atomic accumulation needs an fp32 output"
    conv = isinstance(segs[0][0], Conv3x3)
    assert splits == 1 or (atomic and (len(segs) == 1 or group_out is not None) and not conv), "split-K rules"
    for i, (a, b, k) in enumerate(segs):
        assert k >= 1
        if isinstance(a, Conv3x3):
            assert i == 0 and conv_supported(a.H, a.W) and a.C % 8 == 0 and a.N * a.H * a.W == M, "conv geometry"
            assert (a.t.storage_offset() * 2) % 16 == 0
        else:
            assert (a.rows >= k) if a.mn else (a.inner >= k), f"A[{i}] smaller than K"
            _chk_tma(a, f"A[{i}]")
            assert not a.batched or ((nb0 == 1 or a.sb0 > 0) and (nb1 == 1 or a.sb1 > 0)), "batched operand strides"
        _chk_tma(b, f"B[{i}]")
        assert not b.batched or ((nb0 == 1 or b.sb0 > 0) and (nb1 == 1 or b.sb1 > 0)), "batched operand strides"
    if pair_mode > 0:                               # forced CTA-pair kernel: gemm_host.cu pair_eligible
        k0 = 9 * ((segs[0][0].C + 63) // 64) * 64 if conv else segs[0][2]
        assert M >= 256 and N >= 64 and k0 >= 64 and nb0 == 1 and nb1 == 1 and splits == 1 and group_out is None, "pair kernel rules"
        assert not atomic or out.dtype == torch.float32
        if not conv and segs[0][0].mn:
            assert side is None and len(segs) == 1 and segs[0][0].inner % 8 == 0, "MN-major A: one plain segment"
    kblocks0 = 9 * ((segs[0][0].C + 63) // 64) if conv else (segs[0][2] + 63) // 64
    assert splits <= kblocks0, f"more splits ({splits}) than K blocks ({kblocks0})"
    if side is not None:
        s_mat, b2_mat, r, _, t_out = side
        assert len(segs) == 1 and not conv and nb0 == 1 and nb1 == 1 and splits == 1 and 1 <= r <= 64, "side path rules"
        # ranks above 32 exist only in the CTA-pair kernel (gemm_host.cu: pair_eligible + the size gate)
        assert r <= 32 or (M >= 256 and N >= 64 and (pair_mode > 0 or M * N >= 256 * 512) and segs[0][2] >= 64
                           and not segs[0][0].mn and out.dtype == BF16), "side rank > 32 needs a pair-kernel problem"
        _chk_tma(s_mat, "S")
        _chk_tma(b2_mat, "B2")
    if group_out is not None:
        assert len(segs) == 2 and not conv and side is None and nb0 == 1 and nb1 == 1 and segs[0][2] == segs[1][2]
        assert out.dtype == torch.float32 and atomic and group_out[0].dtype == torch.float32


def gemm(out, M, N, segs, *, d_strides=None, alpha=1.0, bias=None, bias_rows=0, bias_sb=0, residual=None,
         r_strides=None, nb0=1, nb1=1, splits=1, atomic=False, block_n=0, side=None, pair_mode=0, group_out=None, static_b=False,
         geglu_h=None, geglu_out=None):
    _chk_gemm(out, M, N, segs, nb0, nb1, splits, atomic, side, group_out, pair_mode)
    if geglu_out is not None:
        # fused GEGLU forward (CTA-pair kernel, 256-wide tiles): out [M, N] is the projection in the 128-interleaved layout,
        # geglu_out [M, N/2] <- value * gelu(gate)
        assert M >= 256 and N % 256 == 0 and segs[0][2] >= 64 and len(segs) == 1 and side is None and residual is None
        assert not atomic and nb0 == 1 and nb1 == 1 and out.dtype == BF16 and geglu_out.dtype == BF16 and geglu_out.shape == (M, N // 2)
        assert pair_mode > 0 and block_n in (0, 256) and geglu_out.stride(1) == 1 and geglu_out.stride(0) % 8 == 0
        gemm(out, M, N, segs, alpha=alpha, bias=bias)
        geglu_out.copy_(geglu_fwd(out, 128))
        return out
    if geglu_h is not None:
        # fused GEGLU backward (CTA-pair kernel only): the product is dy [M, N]; out [M, 2N] <- geglu_bwd(bf16(dy), h)
        assert M >= 256 and N >= 64 and N % 32 == 0 and segs[0][2] >= 64 and len(segs) == 1 and side is None and bias is None
        assert residual is None and not atomic and nb0 == 1 and nb1 == 1 and out.dtype == BF16 and out.shape == (M, 2 * N)
        assert geglu_h.dtype == BF16 and geglu_h.shape == (M, 2 * N) and geglu_h.stride(0) % 8 == 0 and (geglu_h.storage_offset() * 2) % 16 == 0
        dy = torch.empty(M, N, dtype=BF16)
        gemm(dy, M, N, segs, alpha=alpha)
        out.copy_(geglu_bwd(dy, geglu_h))
        return out
    assert residual is None or residual.dtype == BF16
    if group_out is not None:                       # two independent problems sharing one launch
        out2, (sm2, sn2) = group_out
        gemm(out, M, N, [segs[0]], d_strides=d_strides, alpha=alpha, atomic=atomic)
        gemm(out2, M, N, [segs[1]], d_strides=(sm2, sn2, 0, 0), alpha=alpha, atomic=atomic)
        return out
    if d_strides is None:
        d_strides = (out.stride(0), 1, 0, 0)
    if residual is not None and r_strides is None:
        r_strides = (residual.stride(0), 1, 0, 0)
    for b1 in range(nb1):
        for b0 in range(nb0):
            acc = torch.zeros(M, N)
            for a, b, k in segs:
                if isinstance(a, Conv3x3):
                    x = _view(a.t, (a.N, a.H, a.W, a.C), (a.H * a.W * a.C, a.W * a.C, a.C, 1)).float()
                    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
                    bv = _view(b.t, (b.rows, b.inner), (b.row_stride, 1)).float()
                    for tap in range(9):
                        kh, kw = tap // 3, tap % 3
                        xs = xp[:, kh:kh + a.H, kw:kw + a.W, :].reshape(M, a.C)
                        rows = torch.zeros(N, a.C)
                        r0, c0 = tap * a.b_tap_n, tap * a.b_tap_k
                        rr = max(0, min(N, bv.shape[0] - r0))
                        cc = max(0, min(a.C, bv.shape[1] - c0))
                        rows[:rr, :cc] = bv[r0:r0 + rr, c0:c0 + cc]
                        acc += xs @ rows.t()
                else:
                    A = _operand(a, b0, b1, M, k)
                    Bm = _operand(b, b0, b1, N, k)
                    acc += A @ Bm.t()
            acc = acc * alpha
            if side is not None:
                s_mat, b2_mat, r, s_alpha, t_out = side
                a0, _, k0 = segs[0]
                T = (_operand(a0, 0, 0, M, k0) @ _operand(s_mat, 0, 0, r, k0).t() * s_alpha).to(BF16)
                if t_out is not None:
                    _view(t_out, (M, r), (t_out.stride(0), 1)).copy_(T)
                acc += T.float() @ _operand(b2_mat, 0, 0, N, r).t()
            if bias is not None:
                if bias_rows:
                    idx = torch.arange(M) // bias_rows
                    bv = _view(bias, (int(idx.max()) + 1, N), (bias_sb, 1)).float()
                    acc += bv[idx]
                else:
                    acc += bias.float()[:N]
            if residual is not None:
                acc += _view(residual, (M, N), (r_strides[0], r_strides[1]), b0 * r_strides[2] + b1 * r_strides[3]).float()
            dv = _view(out, (M, N), (d_strides[0], d_strides[1]), b0 * d_strides[2] + b1 * d_strides[3])
            if atomic:
                dv += acc.to(out.dtype)
            else:
                dv.copy_(acc.to(out.dtype))
    return out


def lora_wgrad_batch(problems):
    for X, Y, out, M, Nout, r, sn, sj in problems:
        assert X.dtype == BF16 and Y.dtype == BF16 and out.dtype == torch.float32 and r <= 32
        assert X.stride(0) % 8 == 0 and Y.stride(0) % 8 == 0 and Nout % 8 == 0 and X.stride(-1) == 1 and Y.stride(-1) == 1
        assert (X.storage_offset() * 2) % 16 == 0 and (Y.storage_offset() * 2) % 16 == 0
        assert X.shape[0] >= M and X.shape[1] >= Nout and Y.shape[0] >= M and Y.shape[1] >= r
        upd = X[:M, :Nout].float().t() @ Y[:M, :r].float()              # [Nout, r]
        view = _view(out, (Nout, r), (sn, sj))
        view += upd


def flash_attn_fwd(q, k, v, B, H, L, Lk, scale):
    C_ = H * 64
    assert q.stride(0) % 8 == 0 and k.stride(0) == q.stride(0) and v.stride(0) == q.stride(0)
    q4 = q.float().reshape(B, L, H, 64).transpose(1, 2)
    k4 = k.float().reshape(B, Lk, H, 64).transpose(1, 2)
    v4 = v.float().reshape(B, Lk, H, 64).transpose(1, 2)
    s = q4 @ k4.transpose(-1, -2) * scale
    o = torch.softmax(s, -1).to(BF16).float() @ v4
    return _bf(o.transpose(1, 2).reshape(B * L, C_)), torch.logsumexp(s, -1)


def flash_attn_bwd(q, k, v, o, d_o, lse, B, H, L, Lk, scale, dk=None, dv=None, dq=None, dsc=None):
    C_ = H * 64
    assert k.stride(0) == q.stride(0) and v.stride(0) == q.stride(0) and o.stride(0) == d_o.stride(0)
    for t in (dq, dk, dv):
        assert t is None or (t.stride(1) == 1 and t.stride(0) % 8 == 0 and (t.storage_offset() * 2) % 16 == 0)
    q4 = q.float().reshape(B, L, H, 64).transpose(1, 2).detach().requires_grad_(True)
    k4 = k.float().reshape(B, Lk, H, 64).transpose(1, 2).detach().requires_grad_(True)
    v4 = v.float().reshape(B, Lk, H, 64).transpose(1, 2).detach().requires_grad_(True)
    out = torch.softmax(q4 @ k4.transpose(-1, -2) * scale, -1) @ v4
    out.backward(d_o.float().reshape(B, L, H, 64).transpose(1, 2))
    if dsc is not None:                              # hook: score = sum_h scale * q_h.k_h  ->  its gradient joins dS for every head
        assert dsc.dtype == BF16 and dsc.shape[:2] == (B, L) and dsc.shape[2] % 8 == 0 and dsc.shape[2] >= Lk and Lk <= 128
        assert (dsc.storage_offset() * 2) % 16 == 0 and dsc.stride(1) % 8 == 0
        g_ = dsc[:, :, :Lk].float()[:, None] * scale                      # [B, 1, L, Lk]
        q4.grad += g_ @ k4.detach()
        k4.grad += g_.transpose(-1, -2) @ q4.detach()
    unh = lambda t, n: _bf(t.transpose(1, 2).reshape(B * n, C_))
    gk, gv = unh(k4.grad, Lk), unh(v4.grad, Lk)
    if dk is not None:
        dk.copy_(gk)
        gk = dk
    if dv is not None:
        dv.copy_(gv)
        gv = dv
    gq = unh(q4.grad, L)
    if dq is not None:
        dq.copy_(gq)
        gq = dq
    return gq, gk, gv


def softmax_fwd(S, P, rows, cols, ld_s, ld_p):
    assert cols >= 1 and ld_p >= cols and ld_s >= cols and S.dtype == torch.float32 and P.dtype == BF16
    s = _view(S, (rows, cols), (ld_s, 1))
    p = _view(P, (rows, ld_p), (ld_p, 1))
    p.zero_()
    p[:, :cols] = torch.softmax(s.float(), -1).to(BF16)


def softmax_bwd(P, dP, dS, rows, cols, ld_p, ld_dp):
    p = _view(P, (rows, cols), (ld_p, 1)).float()
    dp = _view(dP, (rows, cols), (ld_dp, 1)).float()
    ds = _view(dS, (rows, ld_p), (ld_p, 1))
    ds.zero_()
    ds[:, :cols] = (p * (dp - (p * dp).sum(-1, keepdim=True))).to(BF16)


def _chk_vec(*ts):
    """Kernels that move 8 bf16 per thread (load8 / store8 / uint4) need 16-byte aligned, contiguous tensors."""
    for t in ts:
        if t is not None:
            assert t.is_contiguous() and (t.storage_offset() * t.element_size()) % 16 == 0, "16-byte vector access"


def groupnorm_fwd(x, gamma, beta, batch, hw, C_, groups, eps, silu):
    assert C_ % 8 == 0 and C_ % groups == 0 and groups <= 64 and C_ // 8 <= 1024 and x.shape == (batch * hw, C_)
    _chk_vec(x, gamma, beta)
    xr = x.float().view(batch, hw, C_).permute(0, 2, 1)
    y = F.group_norm(xr, groups, gamma.float(), beta.float(), eps)
    if silu:
        y = F.silu(_bf(y).float())
    return _bf(y.permute(0, 2, 1).reshape(batch * hw, C_)), (eps,)


def groupnorm_bwd(dy, x, gamma, beta, stats, batch, hw, C_, groups, silu, dres=None):
    xr = x.float().view(batch, hw, C_).permute(0, 2, 1).detach().requires_grad_(True)
    y = F.group_norm(xr, groups, gamma.float(), beta.float(), stats[0])
    if silu:
        y = F.silu(y)
    y.backward(dy.float().view(batch, hw, C_).permute(0, 2, 1))
    dx = xr.grad.permute(0, 2, 1).reshape(batch * hw, C_)
    if dres is not None:
        dx = _bf(dx).float() + dres.float()
    return _bf(dx)


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    assert x.shape[1] % 8 == 0 and x.shape[1] <= 2560
    _chk_vec(x, gamma, beta)
    return _bf(F.layer_norm(x.float(), (x.shape[1],), gamma.float(), beta.float(), eps)), None


def layernorm_bwd(dy, x, gamma, stats, dres=None):
    xr = x.float().detach().requires_grad_(True)
    F.layer_norm(xr, (x.shape[1],), gamma.float(), torch.zeros_like(gamma).float(), 1e-5).backward(dy.float())
    dx = xr.grad
    if dres is not None:
        dx = _bf(dx).float() + dres.float()
    return _bf(dx)


def norm_param_grad(dy, x, gamma, beta, stats, dgamma, dbeta, hw=0, groups=0, silu=False):
    rows, C_ = x.shape
    assert C_ % 8 == 0 and dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32
    assert not groups or (C_ % groups == 0 and hw >= 1 and rows % hw == 0)
    assert not silu or (groups and gamma is not None and beta is not None)
    _chk_vec(dy, x, gamma, beta)
    g = (gamma.float() if gamma is not None else torch.ones(C_)).detach().clone().requires_grad_(True)
    b = (beta.float() if beta is not None else torch.zeros(C_)).detach().clone().requires_grad_(True)
    if groups:
        batch = rows // hw
        y = F.group_norm(x.float().view(batch, hw, C_).permute(0, 2, 1), groups, g, b, stats[0])
        if silu:
            y = F.silu(y)
        y.backward(dy.float().view(batch, hw, C_).permute(0, 2, 1))
    else:
        F.layer_norm(x.float(), (C_,), g, b, 1e-5).backward(dy.float())
    dgamma += g.grad
    dbeta += b.grad


def _geglu_split(h, interleave):
    """(value, gate) halves of h for the plain ([value | gate]) or interleaved (blocks of `interleave` columns) layout."""
    if not interleave:
        return h.chunk(2, -1)
    rows, two_inner = h.shape
    assert interleave % 8 == 0 and (two_inner // 2) % interleave == 0
    hb = h.reshape(rows, two_inner // (2 * interleave), 2, interleave)
    return hb[:, :, 0].reshape(rows, -1), hb[:, :, 1].reshape(rows, -1)


def _geglu_join(a, g, interleave):
    if not interleave:
        return torch.cat([a, g], -1)
    rows, inner = a.shape
    return torch.stack([a.reshape(rows, inner // interleave, interleave), g.reshape(rows, inner // interleave, interleave)],
                       dim=2).reshape(rows, 2 * inner)


def geglu_fwd(h, interleave=0):
    assert (h.shape[1] // 2) % 8 == 0
    _chk_vec(h)
    a, g = _geglu_split(h.float(), interleave)
    return _bf(a * _bf(F.gelu(g)).float())


def geglu_bwd(dy, h, interleave=0):
    a, g = _geglu_split(h.float(), interleave)
    a, g = a.detach().requires_grad_(True), g.detach().requires_grad_(True)
    (a * F.gelu(g)).backward(dy.float())
    return _bf(_geglu_join(a.grad, g.grad, interleave))


def silu_fwd(x):
    return _bf(F.silu(x.float()))


def silu_bwd(dy, x):
    xr = x.float().detach().requires_grad_(True)
    F.silu(xr).backward(dy.float())
    return _bf(xr.grad)


ACT_GELU, ACT_QUICK_GELU = real.ACT_GELU, real.ACT_QUICK_GELU


def _act(x, kind):
    return F.gelu(x) if kind == ACT_GELU else x * torch.sigmoid(1.702 * x)


def act_fwd(x, kind):
    assert x.is_contiguous() and x.dtype == BF16
    return _bf(_act(x.float(), kind))


def act_bwd(dy, x, kind):
    assert x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape
    xf = x.float().detach().requires_grad_(True)
    _act(xf, kind).backward(dy.float())
    return _bf(xf.grad)


def add(a, b, c=None, out=None):
    y = a.float() + b.float()
    if c is not None:
        y = _bf(y).float() + c.float()
    y = _bf(y)
    if out is not None:
        out.copy_(y)
        return out
    return y


def head_pad(x, heads, d_src, d_dst, out=None):
    assert x.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == heads * d_src
    assert d_src % 8 == 0 and d_dst % 8 == 0 and x.stride(0) % 8 == 0 and (x.storage_offset() * 2) % 16 == 0
    rows = x.shape[0]
    y = torch.empty(rows, heads * d_dst, dtype=BF16) if out is None else out
    assert y.shape == (rows, heads * d_dst) and y.stride(1) == 1 and y.stride(0) % 8 == 0 and (y.storage_offset() * 2) % 16 == 0
    n = min(d_src, d_dst)
    full = torch.zeros(rows, heads, d_dst, dtype=BF16)
    full[:, :, :n] = x.reshape(rows, heads, d_src)[:, :, :n]
    y.copy_(full.reshape(rows, heads * d_dst))
    return y


def upsample2x_fwd(x, N, H, W, C_):
    assert C_ % 8 == 0
    _chk_vec(x)
    xn = x.float().view(N, H, W, C_).permute(0, 3, 1, 2)
    return _bf(F.interpolate(xn, scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).reshape(-1, C_))


def upsample2x_bwd(dy, N, H, W, C_):
    return _bf(dy.float().view(N, H, 2, W, 2, C_).sum(dim=(2, 4)).reshape(-1, C_))


def im2col3x3(x, N, H, W, C_, stride):
    assert C_ % 8 == 0 and stride in (1, 2) and x.numel() == N * H * W * C_
    _chk_vec(x)
    xn = x.float().view(N, H, W, C_).permute(0, 3, 1, 2)
    unf = F.unfold(xn, 3, padding=1, stride=stride)
    L = unf.shape[-1]
    return _bf(unf.view(N, C_, 9, L).permute(0, 3, 2, 1).reshape(N * L, 9 * C_))


def col2im3x3(col, N, H, W, C_, stride):
    L = col.shape[0] // N
    f = F.fold(col.float().view(N, L, 9, C_).permute(0, 3, 2, 1).reshape(N, C_ * 9, L), (H, W), 3, padding=1, stride=stride)
    return _bf(f.permute(0, 2, 3, 1).reshape(-1, C_))


def shift_stack9(U, N, H, W, r):
    ld_out = (9 * r + 7) // 8 * 8
    Un = _view(U, (N, H, W, r), (H * W * U.stride(0), W * U.stride(0), U.stride(0), 1)).float()
    out = torch.zeros(N, H, W, ld_out)
    for tap in range(9):
        dh, dw = tap // 3 - 1, tap % 3 - 1
        hs, ws = slice(max(dh, 0), H + min(dh, 0)), slice(max(dw, 0), W + min(dw, 0))
        hsrc, wsrc = slice(max(-dh, 0), H + min(-dh, 0)), slice(max(-dw, 0), W + min(-dw, 0))
        out[:, hs, ws, tap * r:(tap + 1) * r] = Un[:, hsrc, wsrc]
    return _bf(out.reshape(N * H * W, ld_out))


def shift_sum9(Z, N, H, W, r, alpha, ld_t):
    assert Z.dtype == torch.float32 and Z.shape[0] == N * H * W and Z.shape[1] >= 9 * r and ld_t >= r
    z = Z[:, :9 * r].reshape(N, H, W, 9, r)
    zp = F.pad(z, (0, 0, 0, 0, 1, 1, 1, 1))                      # pad W and H by one pixel
    acc = torch.zeros(N, H, W, r)
    for tap in range(9):
        kh, kw = tap // 3, tap % 3
        acc += zp[:, kh:kh + H, kw:kw + W, tap]
    T = torch.zeros(N * H * W, ld_t, dtype=BF16)
    T[:, :r] = (acc * torch.tensor(alpha, dtype=torch.float32)).reshape(N * H * W, r).to(BF16)
    return T


def colsum(x, batch, hw, C_):
    assert C_ % 8 == 0 and x.numel() == batch * hw * C_
    _chk_vec(x)
    return _bf(x.float().view(batch, hw, C_).sum(1))


def lora_transpose_b(params, bt, table):
    for off_b, off_bt, n, rs in table.tolist():
        bt[off_bt:off_bt + n * rs].view(rs, n).copy_(params[off_b:off_b + n * rs].view(n, rs).t())


def lora_pack(params, dst, table):
    for off_b, off_d, N, rs, dst_ld, tr in table.tolist():
        src = params[off_b:off_b + N * rs].view(N, rs)
        if tr:
            _view(dst, (rs, N), (dst_ld, 1), off_d).copy_(src.t())
        else:
            _view(dst, (N, rs), (dst_ld, 1), off_d).copy_(src)


def bicubic_fwd(x, Ho, Wo):
    return _bf(F.interpolate(x.float().permute(0, 3, 1, 2), size=(Ho, Wo), mode="bicubic").permute(0, 2, 3, 1).contiguous())


def bicubic_bwd(dy, Hi, Wi):
    B, Ho, Wo, C_ = dy.shape
    with torch.enable_grad():                       # called from inside an autograd backward
        x = torch.zeros(B, C_, Hi, Wi, requires_grad=True)
        y = F.interpolate(x, size=(Ho, Wo), mode="bicubic")
        (g,) = torch.autograd.grad(y, x, dy.float().permute(0, 3, 1, 2))
    return _bf(g.permute(0, 2, 3, 1).contiguous())


def timestep_embedding(t, dim):
    half = dim // 2
    ex = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    e = t.float()[:, None] * torch.exp(ex)[None]
    return _bf(torch.cat([torch.cos(e), torch.sin(e)], -1))


def latent_sample(mean, logvar, eps, scaling_factor):
    return (mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps) * scaling_factor


def noise_prologue(latent, noise, offset, offset_scale, acp, timesteps):
    B, Cc, H, W = latent.shape
    if offset is not None:
        noise += offset_scale * offset.view(B, Cc, 1, 1)
    a = acp.to(BF16)
    sa = (a[timesteps] ** 0.5).view(B, 1, 1, 1)
    so = ((1 - a[timesteps]) ** 0.5).view(B, 1, 1, 1)
    noisy = sa * latent.to(BF16) + so * noise
    n8 = torch.zeros(B * H * W, 8, dtype=BF16)
    n8[:, :Cc] = noisy.permute(0, 2, 3, 1).reshape(-1, Cc)
    return noisy, n8


def snr_weights(acp, timesteps, snr_gamma):
    snr = ((acp ** 0.5)[timesteps] / ((1 - acp) ** 0.5)[timesteps]) ** 2
    w = torch.minimum(snr, torch.full_like(snr, snr_gamma)) / snr
    return (w / w.mean()).float()


def diffusion_loss(pred, ld_pred, noise, mask, weights, loss_scale, want_grad=True):
    B, Cc, H, W = noise.shape
    pr = pred[:, :Cc].reshape(B, H, W, Cc).permute(0, 3, 1, 2).clone().requires_grad_(True)
    l = (((pr - noise).pow(2) * mask).mean(dim=[1, 2, 3]) * weights).mean()
    (l * loss_scale).backward()
    d8 = torch.zeros(B * H * W, 8, dtype=BF16)
    d8[:, :Cc] = pr.grad.permute(0, 2, 3, 1).reshape(-1, Cc).to(BF16)
    return l.detach().float().view(1), (d8 if want_grad else None)


def token_attention_loss(maps, h, w, n_text, mask3, tok_len, ti_pos, grad_scale, want_grad=True):
    """Reference arithmetic of the kernel: the tensorised regulariser of trainer/loss.py under torch autograd."""
    from sd_lora_trainer_b200.trainer.loss import token_attention_loss_from_maps
    B, hw = maps[0].shape[0], h * w
    assert len(maps) <= 64 and n_text <= 80 and ti_pos.shape[1] <= 8 and mask3.dtype == torch.float32
    for m in maps:
        assert m.dtype == BF16 and m.shape[:2] == (B, hw) and m.shape[2] >= n_text and m.stride(2) == 1
    stacked = torch.stack([m[:, :, :n_text].reshape(B, h, w, n_text) for m in maps]).detach().clone().requires_grad_(want_grad)
    loss = token_attention_loss_from_maps(stacked, mask3[:, None], tok_len, ti_pos)
    G = None
    if want_grad:
        ld_g = (n_text + 7) // 8 * 8
        G = torch.zeros(B, hw, ld_g, dtype=BF16)
        if loss.requires_grad:
            loss.backward()
            if stacked.grad is not None:
                G[:, :, :n_text] = (stacked.grad[0].float() * grad_scale).reshape(B, hw, n_text).to(BF16)
    return loss.detach().float().reshape(1), G


def token_std_loss(rows, grads, mu_t, var_t, coeff):
    total = torch.zeros(1)
    for e, r in enumerate(rows):
        x = r.detach().clone().requires_grad_(True)
        mu, var = torch.tensor(mu_t[e], dtype=BF16), torch.tensor(var_t[e], dtype=BF16)
        le = ((mu - x.std(-1)) ** 2 / var).mean()
        (le * coeff / len(rows)).backward()
        if grads[e] is not None:
            grads[e] += x.grad.float()
        total += le.detach().float() / len(rows)
    return total


def abs_sum(p, out):
    out += p.float().abs().sum()
    return out


def adamw(p, grad, m, v, n_first, *, lr, wd, l1_coeff, lr2, wd2, beta1=0.9, beta2=0.999, eps=1e-8, step,
          grad_scale=1.0, zero_grad=True):
    """Same op-by-op bf16 arithmetic as the CUDA kernel (== torch.optim.AdamW on bf16 tensors)."""
    def seg(sl, lr_, wd_, l1_):
        pv = p[sl]
        g = (grad[sl] * grad_scale).to(BF16)
        if l1_:
            g = (g.float() + l1_ * torch.sign(pv.float())).to(BF16)
        if wd_:
            pv.mul_(1 - lr_ * wd_)
        m[sl].lerp_(g, 1 - beta1)
        v[sl].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = (v[sl].sqrt() / (bc2 ** 0.5)).add_(eps)
        pv.addcdiv_(m[sl], denom, value=-(lr_ / bc1))
    seg(slice(0, n_first), lr, wd, l1_coeff)
    if p.numel() > n_first:
        seg(slice(n_first, p.numel()), lr2, wd2, 0.0)
    if zero_grad:
        grad.zero_()


adamw_pack_hyper = real.adamw_pack_hyper          # pure host code inside the library: packs the 12 floats


def adamw_dev(p, grad, m, v, n_first, hyper_dev, zero_grad=True):
    """Interprets the packed hyper-parameters exactly as the CUDA kernel does (op-by-op bf16 arithmetic)."""
    h = hyper_dev.tolist()
    s0, s1 = h[0:3], h[3:6]
    one_minus_b1, beta2, one_minus_b2, eps, bc2_sqrt, grad_scale = h[6:12]
    f32 = torch.float32

    def seg(sl, decay, neg_step, l1):
        pv = p[sl].float()
        g = (grad[sl] * grad_scale).to(BF16).float()
        if l1 != 0.0:
            g = (g + l1 * torch.sign(pv)).to(BF16).float()
        if decay != 1.0:
            pv = (pv * torch.tensor(decay, dtype=f32)).to(BF16).float()
        mv = m[sl].float()
        mv = (mv + torch.tensor(one_minus_b1, dtype=f32) * (g - mv)).to(BF16).float()
        vv = (v[sl].float() * torch.tensor(beta2, dtype=f32)).to(BF16).float()
        vv = (vv + torch.tensor(one_minus_b2, dtype=f32) * (g * g)).to(BF16).float()
        d = vv.sqrt().to(BF16).float()
        d = (d / torch.tensor(bc2_sqrt, dtype=f32)).to(BF16).float()
        d = (d + torch.tensor(eps, dtype=f32)).to(BF16).float()
        pv = (pv + torch.tensor(neg_step, dtype=f32) * (mv / d)).to(BF16)
        p[sl] = pv
        m[sl] = mv.to(BF16)
        v[sl] = vv.to(BF16)
    seg(slice(0, n_first), *s0)
    if p.numel() > n_first:
        seg(slice(n_first, p.numel()), *s1)
    if zero_grad:
        grad.zero_()


prodigy_init_scalars, prodigy_pack_hyper = real.prodigy_init_scalars, real.prodigy_pack_hyper     # host-only helpers


def prodigy_step(p, grad, s, p0, exp_avg, exp_avg_sq, scal, hyper_dev, zero_grad=True):
    """The three Prodigy kernels (optim.cu) in torch: per-op bf16 rounding, global sums in fp32 / double."""
    f32 = torch.float32
    lr, beta1, beta2, beta3, eps, decay, d_coef, growth, d0, bc, l1, gscale = [float(x) for x in hyper_dev.tolist()]
    t32 = lambda x: float(torch.tensor(x, dtype=f32))
    d = float(scal[0])
    pv = p.float()
    g = (grad * gscale).to(BF16).float()
    if l1 != 0.0:
        g = (g + l1 * torch.sign(pv)).to(BF16).float()
    if zero_grad:
        grad.zero_()
    if lr > 0.0:
        dot = float((g * (p0.float() - pv).to(BF16).float()).sum())
        mv = (exp_avg.float() * beta1).to(BF16).float()
        mv = (mv + t32(d * (1.0 - beta1)) * g).to(BF16)
        vv = (exp_avg_sq.float() * beta2).to(BF16).float()
        vv = (vv + t32(d * d * (1.0 - beta2)) * (g * g)).to(BF16)
        sv = (s.float() * beta3).to(BF16).float()
        sv = (sv + t32((d / d0) * d) * g).to(BF16)
        exp_avg.copy_(mv)
        exp_avg_sq.copy_(vv)
        s.copy_(sv)
        scal[4] += dot
        scal[3] += float(sv.float().abs().sum())
    den = float(scal[3])
    dlr = d * lr * bc
    num = float(scal[2]) * beta3
    if lr > 0.0:
        num += (d / d0) * dlr * float(scal[4])
    scal[5] = dlr
    scal[3] = 0.0
    scal[4] = 0.0
    if den == 0.0:
        scal[6] = 1.0
        return
    d_max = float(scal[1])
    d_hat = d
    if lr > 0.0:
        d_hat = d_coef * num / den
        if d == d0:
            d = max(d, d_hat)
        d_max = max(d_max, d_hat)
        d = min(d_max, d * growth)
    scal[0], scal[1], scal[2], scal[6], scal[7] = d, d_max, num, 0.0, d_hat
    denom = exp_avg_sq.float().sqrt().to(BF16).float()
    denom = (denom + t32(d * eps)).to(BF16).float()
    pv = p.float()
    if decay != 0.0:
        pv = (pv + t32(-decay * dlr) * pv).to(BF16).float()
    pv = (pv - t32(dlr) * (exp_avg.float() / denom)).to(BF16)
    p.copy_(pv)


def install(monkeypatch):
    """Patch every product module that holds a reference to ops."""
    import sd_lora_trainer_b200.step as step_mod
    import sd_lora_trainer_b200.unet as unet_mod
    import sd_lora_trainer_b200.trainer.loss as loss_mod
    import sd_lora_trainer_b200.clip as clip_mod
    import sd_lora_trainer_b200.vae as vae_mod
    import sd_lora_trainer_b200.trainer.dataset as dataset_mod
    import sys
    me = sys.modules[__name__]
    for mod in (step_mod, unet_mod, loss_mod, clip_mod, vae_mod, dataset_mod):
        monkeypatch.setattr(mod, "ops", me)
