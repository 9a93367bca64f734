"""B200-native VAE-encode prologue: ``vae.encode(image).latent_dist`` of the reference's one-time dataset pass
(trainer/dataset.py:141-179) as a forward-only executor over the same sm_100a kernels the UNet uses
(ops.py): NHWC bf16 activations, every 3x3 convolution a tcgen05 GEMM (implicit for maps up to 128 wide, im2col above),
GroupNorm+SiLU fused in one pass, residual adds in the GEMM epilogue, the single-head mid-block attention as two batched
GEMMs around the softmax kernel.  Graph = diffusers 0.29.2 ``AutoencoderKL.encode`` (Encoder + quant_conv) with
diffusers parameter names, so a real ``vae`` state dict loads unchanged.

Two load-time rewrites keep the run-time graph to kernels that already exist:

* **mirrored coordinates.**  ``Downsample2D(padding=0)`` pads bottom/right only (``F.pad(x, (0, 1, 0, 1))``) before its
  stride-2 convolution; the stride-2 im2col kernel pads top/left.  A convolution commutes with a spatial flip if its
  taps are flipped too, and GroupNorm / SiLU / 1x1 convolutions / attention do not care about pixel order, so the whole
  encoder runs on the image flipped in H and W with every 3x3 kernel flipped once at load time: bottom/right padding
  becomes top/left padding (even H and W).  The [B, 8, h, w] result is flipped back (a few KB).
* **quant_conv folded into conv_out.**  ``quant_conv(conv_out(x)) = (Wq.Wc) * x + (Wq.bc + bq)``: composed in fp32 at
  load time, rounded to bf16 once.

The reference keeps the VAE in fp32 (main.py:186); this path computes in bf16 with fp32 accumulation, so the posterior
parameters carry bf16 rounding (tests state the tolerance).  No backward: the VAE is frozen and sits before the step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import BF16, Mat
from .unet import GN, Conv3, Lin, _r8

VAE_BLOCK_OUT_CHANNELS = (128, 256, 512, 512)       # vae/config.json of SD1.5 and SDXL-base
VAE_LAYERS_PER_BLOCK = 2
VAE_NORM_GROUPS = 32


class _Res:
    """ResnetBlock2D without a time embedding: conv2(silu(gn2(conv1(silu(gn1(x)))))) + shortcut(x)."""

    def __init__(self, n1: GN, c1: Conv3, n2: GN, c2: Conv3, sc: Optional[Lin]):
        self.n1, self.c1, self.n2, self.c2, self.sc = n1, c1, n2, c2, sc

    def fwd(self, x: torch.Tensor, N: int, H: int, W: int) -> torch.Tensor:
        h = self.c1.fwd(_gn(self.n1, x, N, H * W), N, H, W)
        sc = self.sc.fwd(x, save=False) if self.sc is not None else x
        return self.c2.fwd(_gn(self.n2, h, N, H * W), N, H, W, residual=sc)


def _gn(gn: GN, x: torch.Tensor, N: int, hw: int) -> torch.Tensor:
    y = gn.fwd(x, N, hw)
    gn.sv = None                                    # forward only: nothing is kept for a backward
    return y


class VAEEncoderB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0",
                 block_out_channels: Tuple[int, ...] = VAE_BLOCK_OUT_CHANNELS,
                 layers_per_block: int = VAE_LAYERS_PER_BLOCK, norm_num_groups: int = VAE_NORM_GROUPS,
                 max_score_bytes: int = 8 << 30):
        self.device = torch.device(device)
        self.groups = norm_num_groups
        self.max_score_bytes = max_score_bytes
        self._sd = {k: v for k, v in state_dict.items() if k.startswith("encoder.") or k.startswith("quant_conv.")}
        boc = block_out_channels
        self.conv_in = self._conv("encoder.conv_in")
        self.down: List[Tuple[List[_Res], Optional[Conv3]]] = []
        cout = boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            p = f"encoder.down_blocks.{i}"
            rs = [self._res(f"{p}.resnets.{j}") for j in range(layers_per_block)]
            ds = self._conv(f"{p}.downsamplers.0.conv", stride=2) if i < len(boc) - 1 else None
            self.down.append((rs, ds))
        self.mid_res = [self._res("encoder.mid_block.resnets.0"), self._res("encoder.mid_block.resnets.1")]
        a = "encoder.mid_block.attentions.0"
        self.attn_norm = GN(self._w(f"{a}.group_norm.weight"), self._w(f"{a}.group_norm.bias"), self.groups, 1e-6, False)
        self.to_q, self.to_k, self.to_v, self.to_out = (self._lin(f"{a}.{n}") for n in ("to_q", "to_k", "to_v", "to_out.0"))
        self.norm_out = GN(self._w("encoder.conv_norm_out.weight"), self._w("encoder.conv_norm_out.bias"), self.groups,
                           1e-6, True)
        # quant_conv . conv_out as ONE 3x3 convolution (composed in fp32)
        wc = self._sd["encoder.conv_out.weight"].to(self.device, torch.float32)
        bc = self._sd["encoder.conv_out.bias"].to(self.device, torch.float32)
        wq = self._sd["quant_conv.weight"].to(self.device, torch.float32).flatten(1)            # [8, 8]
        bq = self._sd["quant_conv.bias"].to(self.device, torch.float32)
        w = torch.einsum("om,mikl->oikl", wq, wc)
        self.conv_out = Conv3(w.flip(2, 3).to(BF16), (wq @ bc + bq).to(BF16), need_dgrad=False)
        self.moment_channels = w.shape[0]
        self._sd = None

    # ---- construction ------------------------------------------------------------------------------
    def _w(self, name: str) -> torch.Tensor:
        return self._sd[name].detach().to(self.device, BF16).contiguous()

    def _conv(self, name: str, stride: int = 1) -> Conv3:
        # taps flipped once: the executor runs in mirrored image coordinates (module docstring)
        return Conv3(self._w(f"{name}.weight").flip(2, 3).contiguous(), self._w(f"{name}.bias"), stride=stride,
                     need_dgrad=False)

    def _lin(self, name: str) -> Lin:
        W = self._w(f"{name}.weight")
        if W.dim() == 4:                                           # 1x1 conv_shortcut; legacy checkpoints: conv q/k/v
            W = W.reshape(W.shape[0], W.shape[1]).contiguous()
        return Lin(W, self._w(f"{name}.bias"))

    def _res(self, p: str) -> _Res:
        g = self.groups
        sc = self._lin(f"{p}.conv_shortcut") if f"{p}.conv_shortcut.weight" in self._sd else None
        return _Res(GN(self._w(f"{p}.norm1.weight"), self._w(f"{p}.norm1.bias"), g, 1e-6, True), self._conv(f"{p}.conv1"),
                    GN(self._w(f"{p}.norm2.weight"), self._w(f"{p}.norm2.bias"), g, 1e-6, True), self._conv(f"{p}.conv2"), sc)

    # ---- the single-head mid-block attention (head_dim = C): S and P live in HBM, one image group at a time ----------
    def _attention(self, x: torch.Tensor, N: int, L: int) -> torch.Tensor:
        C = x.shape[1]
        hn = _gn(self.attn_norm, x, N, L)
        q, k, v = (p.fwd(hn, save=False) for p in (self.to_q, self.to_k, self.to_v))
        Lp = _r8(L)
        scale = C ** -0.5
        O = torch.empty(N * L, C, dtype=BF16, device=x.device)
        per_image = L * Lp * 6                                     # fp32 scores + bf16 probabilities
        nb = max(1, min(N, self.max_score_bytes // max(per_image, 1)))
        for n0 in range(0, N, nb):
            n = min(nb, N - n0)
            qs, ks, vs, os_ = (t[n0 * L:(n0 + n) * L] for t in (q, k, v, O))
            S = torch.empty(n, L, Lp, dtype=torch.float32, device=x.device)
            P = torch.empty(n, L, Lp, dtype=BF16, device=x.device)
            ops.gemm(S, L, L, [(Mat(qs, L, C, C, sb1=L * C, batched=True), Mat(ks, L, C, C, sb1=L * C, batched=True), C)],
                     d_strides=(Lp, 1, 0, L * Lp), alpha=scale, nb0=1, nb1=n)
            ops.softmax_fwd(S, P, n * L, L, Lp, Lp)
            del S
            ops.gemm(os_, L, C, [(Mat(P, L, L, Lp, sb1=L * Lp, batched=True),
                                  Mat(vs, L, C, C, mn=True, sb1=L * C, batched=True), L)],
                     d_strides=(C, 1, 0, L * C), nb0=1, nb1=n)
            del P
        return self.to_out.fwd(O, residual=x, save=False)

    # ---- forward -------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_moments(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, 3, H, W] in [-1, 1] (H, W multiples of 2**(levels-1)) -> posterior parameters [B, 8, H/8, W/8] fp32
        (mean | logvar), i.e. ``vae.encode(image).latent_dist.parameters``."""
        B, Ci, H, W = image.shape
        levels = len(self.down)
        assert H % (1 << (levels - 1)) == 0 and W % (1 << (levels - 1)) == 0, "image sides must divide by the VAE stride"
        # NHWC, mirrored in H and W, channels padded to 8 (16-byte pixels for TMA)
        x = torch.zeros(B, H, W, self.conv_in.cin_p, dtype=BF16, device=self.device)
        x[..., :Ci] = image.to(self.device).flip(2, 3).permute(0, 2, 3, 1)
        x = self.conv_in.fwd(x.view(B * H * W, -1), B, H, W)
        h, w = H, W
        for rs, ds in self.down:
            for r in rs:
                x = r.fwd(x, B, h, w)
            if ds is not None:
                x = ds.fwd(x, B, h, w)
                h, w = h // 2, w // 2
        x = self.mid_res[0].fwd(x, B, h, w)
        x = self._attention(x, B, h * w)
        x = self.mid_res[1].fwd(x, B, h, w)
        y = self.conv_out.fwd(_gn(self.norm_out, x, B, h * w), B, h, w)
        m = self.moment_channels
        return y[:, :m].float().view(B, h, w, m).flip(1, 2).permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def encode(self, image: torch.Tensor, eps: Optional[torch.Tensor] = None, scaling_factor: float = 1.0,
               generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(parameters, sample * scaling_factor): the cached posterior and one draw from it (dataset.py:157-158, 186)."""
        params = self.encode_moments(image)
        mean, logvar = torch.chunk(params, 2, dim=1)
        if eps is None:
            eps = torch.randn(mean.shape, generator=generator, device=mean.device, dtype=torch.float32)
        return params, ops.latent_sample(mean.contiguous(), logvar.contiguous(), eps.to(mean.device, torch.float32),
                                         scaling_factor)
