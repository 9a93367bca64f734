"""B200-native LoRA UNet executor: explicit forward AND backward over the sm_100a kernels (ops.py), no autograd
inside.  Stands in for ``unet(...)`` + ``loss.backward()`` of the reference step (main.py:329-336, 363) with the
PEFT-wrapped diffusers UNet2DConditionModel behind it:

  * activations are NHWC ``[B*H*W, C]`` bf16 so every conv / linear is a K-major tcgen05 GEMM operand;
  * every LoRA target (to_q/to_k/to_v/to_out.0/conv2, trainer/optimizer.py:84) runs ``W.x + s.B.(A.x)`` as ONE
    dual-segment GEMM; its backward emits dX (frozen W read MN-major - no transposed copies, no dW) and
    accumulates dA/dB straight into one flat fp32 gradient buffer (split-K atomics);
  * frozen affine/bias parameters never get gradients;
  * attention keeps P = softmax(QK^T/sqrt(d)) resident in HBM for the backward (180 GB makes that the cheap
    choice) and exposes the head-summed pre-softmax cross-attention scores the reference's
    DAAMLossAttnProcessor2_0 captures (trainer/ti_cross_attn_loss.py:201-212) as ONE full-width GEMM, since
    sum_h q_h.k_h == q.k over all channels.
Parameter names follow diffusers / PEFT so reference checkpoints load unchanged.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .arch import UNetArch
from .ops import BF16, Conv3x3, Mat, kmajor, mnmajor

LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0", "conv2")      # trainer/optimizer.py:84
MAX_SIDE_RANK = 32                                                # widest in-kernel LoRA side path (b200_gemm_t.side_r)
_SMS = 148


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


class _WgradLane:
    """Optional (B200_WGRAD_STREAM=1): run the LoRA weight-gradient GEMMs (dA, dB) on a second stream - inside the
    step's CUDA graph a fork per layer and ONE join at the end of the backward.  Measured on the B200 it does NOT help
    (91.7 vs 91.1 ms/step, profiles/r01c_*): every GEMM CTA holds > 200 KB of shared memory, so the forked kernels cannot
    co-reside with the main chain and only add dependency edges.  Off by default; kept for the A/B.  Operands are kept
    alive until the join (the caching allocator must not hand their memory to the main stream meanwhile)."""

    def __init__(self):
        self.stream: Optional[torch.cuda.Stream] = None
        self.keep: List[torch.Tensor] = []

    def begin(self, enabled: bool):
        self.keep = []
        self.stream = torch.cuda.Stream() if (enabled and torch.cuda.is_available()) else None

    def run(self, fn, *operands):
        """fn() launches the weight-gradient GEMMs; everything they read was produced on the current stream."""
        if self.stream is None:
            fn()
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(t for t in operands if t is not None)

    def join(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep = []
        self.stream = None


WGRAD = _WgradLane()
GROUP_WGRAD = os.environ.get("B200_GROUP_WGRAD", "1") != "0"


class _WgradQueue:
    """The dA / dB weight-gradient problems of the LoRA linears are queued during a transformer block's backward and leave
    as ONE table-driven launch at its end (ops.lora_wgrad_batch; 12 problems per SDXL block) instead of one tcgen05 GEMM
    launch per layer - 84 MFLOP problems whose cost was launch latency (B200_WGRAD_BATCH=0 restores the per-layer form).
    The queued operands stay referenced until the flush, so the caching allocator cannot recycle them."""

    def __init__(self):
        self.enabled = os.environ.get("B200_WGRAD_BATCH", "1") != "0"
        self.active = False
        self.items: List[tuple] = []

    def accepts(self, *mats_rows) -> bool:
        return self.active and all(t.stride(0) % 8 == 0 and t.stride(-1) == 1 for t in mats_rows)

    def flush(self):
        if self.items:
            ops.lora_wgrad_batch(self.items)
            self.items = []


WQ = _WgradQueue()


def _wgrad_splits(out_rows: int, reduce_len: int) -> int:
    tiles = (out_rows + 127) // 128
    kblocks = (reduce_len + 63) // 64
    return max(1, min(_SMS // max(tiles, 1), kblocks, 32))


# =================================================================================================
# LoRA parameter store: ONE flat bf16 parameter buffer, ONE flat fp32 gradient buffer, bf16 Adam moments
# =================================================================================================
class LoraSlot:
    __slots__ = ("name", "kind", "r", "rs", "fan_in", "fan_out", "offA", "offB", "offBt", "store")

    def __init__(self, name, kind, r, fan_in, fan_out):
        self.name, self.kind, self.r, self.rs = name, kind, r, _r8(r)
        self.fan_in, self.fan_out = fan_in, fan_out
        self.offA = self.offB = self.offBt = -1
        self.store = None

    # A: linear [r, K]; conv [9*r, Cin] (tap-major master layout).  B: [N, rs] with columns >= r kept at zero.
    @property
    def a_rows(self):
        return self.r if self.kind == "linear" else 9 * self.r

    @property
    def numel_logical(self):
        return self.a_rows * self.fan_in + self.fan_out * self.r

    def A(self):
        return self.store.params[self.offA:self.offA + self.a_rows * self.fan_in].view(self.a_rows, self.fan_in)

    def B(self):
        return self.store.params[self.offB:self.offB + self.fan_out * self.rs].view(self.fan_out, self.rs)

    def Bt(self):
        """K-major copy [rs, N] of B (linear layers), refreshed once per step by LoraStore.refresh_bt()."""
        return self.store.params_bt[self.offBt:self.offBt + self.fan_out * self.rs].view(self.rs, self.fan_out)

    def gA(self):
        return self.store.grads[self.offA:self.offA + self.a_rows * self.fan_in].view(self.a_rows, self.fan_in)

    def gB(self):
        return self.store.grads[self.offB:self.offB + self.fan_out * self.rs].view(self.fan_out, self.rs)


class LoraStore:
    def __init__(self, device, scaling: float):
        self.device = device
        self.scaling = scaling
        self.slots: List[LoraSlot] = []
        self.by_name: Dict[str, LoraSlot] = {}
        self.total = 0
        self.params = self.grads = self.m = self.v = None
        self.pack_rows: List[List[int]] = []
        self.pack_total = 0
        self.packed = self.pack_table = None

    def add(self, name: str, kind: str, r: int, fan_in: int, fan_out: int) -> LoraSlot:
        if name in self.by_name:                       # pre-allocated (contiguous batched groups)
            return self.by_name[name]
        s = LoraSlot(name, kind, r, fan_in, fan_out)
        s.offA = self.total
        self.total += _r8(s.a_rows * fan_in)
        s.offB = self.total
        self.total += _r8(fan_out * s.rs)
        s.store = self
        self.slots.append(s)
        self.by_name[name] = s
        return s

    def add_fused(self, names: List[str], r: int, fan_in: int, fan_outs: List[int]) -> List[LoraSlot]:
        """Linear slots whose A factors sit back to back ([len(names)*r, fan_in] is one contiguous matrix): the q|k|v
        projections of a self-attention block, run as ONE GEMM with a rank-3r side path (LinQKV)."""
        slots = [LoraSlot(n, "linear", r, fan_in, fo) for n, fo in zip(names, fan_outs)]
        for s in slots:
            assert s.name not in self.by_name and (s.a_rows * fan_in) % 8 == 0
            s.offA = self.total
            self.total += s.a_rows * fan_in
        for s in slots:
            s.offB = self.total
            self.total += _r8(s.fan_out * s.rs)
            s.store = self
            self.slots.append(s)
            self.by_name[s.name] = s
        return slots

    def add_pack(self, rows: List[List[int]], numel: int) -> int:
        """Reserve `numel` elements of the derived-operand buffer (zero except for the listed blocks) and register the
        lora_pack table rows [off_B, off_dst (relative), N, rs, dst_ld, transpose]; returns the base offset."""
        base = self.pack_total
        for r_ in rows:
            self.pack_rows.append([r_[0], base + r_[1], r_[2], r_[3], r_[4], r_[5]])
        self.pack_total += _r8(numel)
        return base

    def finalize(self, extra: int = 0):
        """`extra` trailing elements hold the trainable textual-inversion rows (same optimizer kernel)."""
        n = self.total + extra
        self.n_lora = self.total
        self.params = torch.zeros(n, dtype=BF16, device=self.device)
        self.packed = torch.zeros(max(self.pack_total, 8), dtype=BF16, device=self.device)
        self.pack_table = torch.tensor(self.pack_rows, dtype=torch.int64, device=self.device) if self.pack_rows else None
        # K-major copies of the linear layers' B factors (side operand of the fused input-gradient GEMM)
        rows, off = [], 0
        for s in self.slots:
            if s.kind == "linear":
                s.offBt = off
                rows.append([s.offB, off, s.fan_out, s.rs])
                off += _r8(s.fan_out * s.rs)
        self.params_bt = torch.zeros(max(off, 8), dtype=BF16, device=self.device)
        self.bt_table = torch.tensor(rows, dtype=torch.int64, device=self.device) if rows else None
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.m = torch.zeros(n, dtype=BF16, device=self.device)
        self.v = torch.zeros(n, dtype=BF16, device=self.device)

    def refresh_bt(self):
        if self.bt_table is not None:
            ops.lora_transpose_b(self.params, self.params_bt, self.bt_table)

    def refresh_packed(self):
        """Block-diagonal LoRA-B operands of the fused q|k|v projections follow the trained factors (once per step)."""
        if self.pack_table is not None:
            ops.lora_pack(self.params, self.packed, self.pack_table)

    @property
    def numel_logical(self) -> int:
        return sum(s.numel_logical for s in self.slots)

    def init_gaussian(self, seed: int):
        """PEFT init_lora_weights='gaussian': A ~ N(0, (1/r)^2), B = 0."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for s in self.slots:
            a = torch.randn(s.a_rows, s.fan_in, generator=g) * (1.0 / s.r)
            s.A().copy_(a.to(BF16))
            s.B().zero_()

    # ---- PEFT-layout import / export ----------------------------------------------------------
    def load_peft(self, sd: Dict[str, torch.Tensor]):
        for s in self.slots:
            a = sd[f"{s.name}.lora_A.default.weight"].to(self.device, BF16)
            b = sd[f"{s.name}.lora_B.default.weight"].to(self.device, BF16)
            if s.kind == "conv":
                a = a.permute(2, 3, 0, 1).reshape(9 * s.r, s.fan_in)       # [r,Cin,3,3] -> [(kh,kw,r), Cin]
                b = b.reshape(s.fan_out, s.r)
            s.A().copy_(a)
            s.B().zero_()
            s.B()[:, :s.r].copy_(b)

    def export_peft(self, grads: bool = False) -> Dict[str, torch.Tensor]:
        out = {}
        for s in self.slots:
            a = (s.gA() if grads else s.A()).clone()
            b = (s.gB() if grads else s.B())[:, :s.r].clone()
            if s.kind == "conv":
                a = a.view(3, 3, s.r, s.fan_in).permute(2, 3, 0, 1).contiguous()
                b = b.reshape(s.fan_out, s.r, 1, 1)
            out[f"{s.name}.lora_A.default.weight"] = a
            out[f"{s.name}.lora_B.default.weight"] = b.contiguous()
        return out


# =================================================================================================
# Dense parameter store (full-UNet fine-tune, BASELINE config 5 / main.py:143-148): EVERY UNet parameter in one flat
# bf16 buffer (in the executor's native layouts), one flat fp32 gradient buffer, bf16 Adam moments
# =================================================================================================
class DenseStore:
    def __init__(self, device):
        self.device = device
        self.entries: List[Tuple[object, str, str, str, str, Tuple[int, ...]]] = []
        self.total = 0
        self.params = self.grads = self.m = self.v = None
        self.views: Dict[str, Tuple[torch.Tensor, torch.Tensor, str, Tuple[int, ...]]] = {}

    def add(self, owner, attr: str, gattr: str, key: str, kind: str, meta: Tuple[int, ...] = ()):
        """owner.attr holds the parameter now; finalize() re-points it (and owner.gattr, its fp32 gradient) at views of
        the flat buffers.  kind: 'mat' ([N, K] <-> diffusers [N, K] or [N, K, 1, 1]), 'vec', 'conv' ([Cout, (kh,kw,c_p)]
        <-> [Cout, Cin, 3, 3], meta = (cout, cin, cin_p))."""
        self.entries.append((owner, attr, gattr, key, kind, meta))

    def finalize(self):
        offs = []
        for owner, attr, *_ in self.entries:
            offs.append(self.total)
            self.total += _r8(getattr(owner, attr).numel())
        n = max(self.total, 8)
        self.params = torch.zeros(n, dtype=BF16, device=self.device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.m = torch.zeros(n, dtype=BF16, device=self.device)
        self.v = torch.zeros(n, dtype=BF16, device=self.device)
        for off, (owner, attr, gattr, key, kind, meta) in zip(offs, self.entries):
            t = getattr(owner, attr)
            pv = self.params[off:off + t.numel()].view(t.shape)
            pv.copy_(t)
            gv = self.grads[off:off + t.numel()].view(t.shape)
            setattr(owner, attr, pv)
            setattr(owner, gattr, gv)
            self.views[key] = (pv, gv, kind, meta)

    @property
    def numel_logical(self) -> int:
        n = 0
        for pv, _, kind, meta in self.views.values():
            n += meta[0] * meta[1] * 9 if kind == "conv" else pv.numel()
        return n

    def export(self, grads: bool = False) -> Dict[str, torch.Tensor]:
        """diffusers-layout state dict (or gradients) of the trained UNet."""
        out = {}
        for key, (pv, gv, kind, meta) in self.views.items():
            t = gv if grads else pv
            if kind == "conv":
                cout, cin, cin_p = meta
                t = t.view(cout, 3, 3, cin_p)[..., :cin].permute(0, 3, 1, 2)
            elif kind == "mat" and meta:
                t = t.reshape(meta)
            out[key] = t.contiguous().clone()
        return out


PAIR_WGRAD = os.environ.get("B200_PAIR_WGRAD", "1") != "0"


def _pair_wgrad_ok(out_rows: int, out_cols: int, reduce_len: int) -> bool:
    """Dense weight gradient dW[N, K] += dY^T . X on the CTA-pair kernel (MN-major A): what gemm_host.cu's pair_eligible takes,
    and enough output tiles to be worth it (otherwise the split-K single-CTA form)."""
    return PAIR_WGRAD and out_rows >= 256 and out_cols >= 128 and reduce_len >= 64 and out_rows % 8 == 0 and out_cols % 8 == 0


def _dense_splits(out_rows: int, out_cols: int, reduce_len: int) -> int:
    tiles = ((out_rows + 127) // 128) * ((out_cols + 255) // 256)
    kblocks = (reduce_len + 63) // 64
    return max(1, min(_SMS // max(tiles, 1), kblocks, 32))


# =================================================================================================
# layers
# =================================================================================================
class Lin:
    """Frozen linear y = x.W^T + b, optionally with the fused low-rank side path."""

    def __init__(self, W: torch.Tensor, b: Optional[torch.Tensor], lora: Optional[LoraSlot] = None):
        self.W, self.b, self.lora = W.contiguous(), b, lora
        self.N, self.K = W.shape
        self.x = self.T = None
        self.gW = self.gb = None                      # fp32 gradient views (DenseStore) when the layer itself trains

    def fwd(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, save: bool = True, bias=None,
            bias_rows: int = 0) -> torch.Tensor:
        M = x.shape[0]
        y = torch.empty(M, self.N, dtype=BF16, device=x.device)
        T, side = None, None
        segs = [(kmajor(x), kmajor(self.W), self.K)]
        if self.lora is not None:
            lo = self.lora
            T = torch.empty(M, lo.rs, dtype=BF16, device=x.device)
            if lo.r <= MAX_SIDE_RANK:
                # ONE launch: the rank-r product T = s.x.A^T accumulates in TMEM beside the main tile, is rounded to
                # bf16 in-kernel and multiplied with B by a final MMA; T also leaves for the dB weight-gradient GEMM.
                side = (Mat(lo.A(), lo.r, self.K, self.K), Mat(lo.B(), self.N, lo.r, lo.rs), lo.r, lo.store.scaling, T)
            else:
                # ranks beyond the in-kernel side path (e.g. train_configs/training_args_style_sd15_noti.json: 64): T by
                # its own GEMM, then T.B^T as a second K segment of the main launch (the conv-LoRA form)
                ops.gemm(T, M, lo.r, [(kmajor(x), Mat(lo.A(), lo.r, self.K, self.K), self.K)], d_strides=(lo.rs, 1, 0, 0),
                         alpha=lo.store.scaling)
                segs.append((Mat(T, M, lo.r, lo.rs), Mat(lo.B(), self.N, lo.r, lo.rs), lo.r))
        ops.gemm(y, M, self.N, segs, bias=self.b if bias is None else bias,
                 bias_rows=bias_rows, bias_sb=self.N if bias_rows else 0, residual=residual, side=side, static_b=True)
        if save:
            self.x, self.T = x, T
        return y

    def bwd(self, dy: torch.Tensor, need_dx: bool = True, accum: Optional[torch.Tensor] = None):
        M = dy.shape[0]
        x, lo, T = self.x, self.lora, self.T
        self.x = self.T = None
        dx, U, side = None, None, None
        segs = [(kmajor(dy), mnmajor(self.W), self.N)]
        if lo is not None:
            r, rs = lo.r, lo.rs
            U = torch.empty(M, rs, dtype=BF16, device=dy.device)
            if need_dx and r > MAX_SIDE_RANK:
                ops.gemm(U, M, r, [(kmajor(dy), Mat(lo.B(), self.N, r, rs, mn=True), self.N)], d_strides=(rs, 1, 0, 0),
                         alpha=lo.store.scaling)
                segs.append((Mat(U, M, r, rs), Mat(lo.A(), r, self.K, self.K, mn=True), r))
            elif need_dx:
                # dX = dY.W + (s.dY.B).A in ONE launch; U = s.dY.B leaves for the dA weight-gradient GEMM
                # (S = the K-major copy of B: read in place it is a tiny MN-major box every CTA re-fetches per k-block)
                side = (Mat(lo.Bt(), r, self.N, self.N), Mat(lo.A(), r, self.K, self.K, mn=True), r, lo.store.scaling, U)
            else:
                ops.gemm(U, M, r, [(kmajor(dy), Mat(lo.B(), self.N, r, rs, mn=True), self.N)], d_strides=(rs, 1, 0, 0),
                         alpha=lo.store.scaling)
        if need_dx:
            dx = accum if accum is not None else torch.empty(M, self.K, dtype=BF16, device=dy.device)
            ops.gemm(dx, M, self.K, segs, residual=accum, side=side, static_b=True)
        if lo is not None and lo.r <= 32 and self.N % 8 == 0 and self.K % 8 == 0 and WQ.accepts(dy, T, x, U):
            # dB[N, r] += dY^T . T      dA[r, K] += U^T . X : queued, one batched launch per transformer block
            WQ.items.append((dy, T, lo.gB(), M, self.N, r, rs, 1))
            WQ.items.append((x, U, lo.gA(), M, self.K, r, 1, self.K))
        elif lo is not None:
            # dB[N, r] += dY^T . T      dA[r, K] += U^T . X   (both operands MN-major, split-K fp32 atomics)
            def wgrad():
                seg_b = (Mat(dy, M, self.N, dy.stride(0), mn=True), Mat(T, M, r, rs, mn=True), M)
                seg_a = (Mat(x, M, self.K, x.stride(0), mn=True), Mat(U, M, r, rs, mn=True), M)
                if self.N == self.K and GROUP_WGRAD:
                    # square projection: both weight gradients have the same tiling -> ONE grouped launch
                    ops.gemm(lo.gB(), self.N, r, [seg_b, seg_a], d_strides=(rs, 1, 0, 0), atomic=True,
                             splits=_wgrad_splits(2 * self.N, M), group_out=(lo.gA(), (1, self.K)))
                else:
                    ops.gemm(lo.gB(), self.N, r, [seg_b], d_strides=(rs, 1, 0, 0), splits=_wgrad_splits(self.N, M), atomic=True)
                    ops.gemm(lo.gA(), self.K, r, [seg_a], d_strides=(1, self.K, 0, 0), splits=_wgrad_splits(self.K, M), atomic=True)
            WGRAD.run(wgrad, dy, T, x, U)
        if self.gW is not None:
            # dense fine-tune: dW[N, K] += dY^T . X (both operands MN-major, split-K fp32 atomics), db += colsum(dY)
            seg = (Mat(dy, M, self.N, dy.stride(0), mn=True), Mat(x, M, self.K, x.stride(0), mn=True), M)
            if _pair_wgrad_ok(self.N, self.K, M):
                # CTA-pair kernel: both operands MN-major, every [256 x BN] tile of dW read-modify-written by its one owner
                ops.gemm(self.gW, self.N, self.K, [seg], d_strides=(self.K, 1, 0, 0), atomic=True, pair_mode=1)
            else:
                ops.gemm(self.gW, self.N, self.K, [seg], d_strides=(self.K, 1, 0, 0), splits=_dense_splits(self.N, self.K, M), atomic=True)
            if self.gb is not None:
                self.gb += ops.colsum(dy if dy.is_contiguous() else dy.contiguous(), 1, M, self.N)[0].float()
        return dx


def _interleave_geglu_rows(lin: "Lin", il: int = 128):
    """FeedForward.net.0.proj for the fused GEGLU epilogue: rows (and bias) re-ordered so that blocks of `il` value rows
    alternate with the matching `il` gate rows - every 256-wide output tile then holds both halves of 128 GEGLU outputs.
    The layer is frozen (LoRA mode), so the permutation is invisible outside: its output h is only ever read by the GEGLU
    kernels (told the layout) and its input gradient sums over all rows."""
    inner = lin.N // 2
    if inner % il:
        return
    W = lin.W
    lin.W = torch.stack([W[:inner].view(inner // il, il, lin.K), W[inner:].view(inner // il, il, lin.K)], dim=1) \
        .reshape(lin.N, lin.K).contiguous()
    if lin.b is not None:
        b = lin.b
        lin.b = torch.stack([b[:inner].view(inner // il, il), b[inner:].view(inner // il, il)], dim=1).reshape(lin.N).contiguous()
    lin.geglu_il = il


def _lin_fwd_geglu(self, x: torch.Tensor):
    """(h, y): the FF up-projection h = x.W^T + b and y = GEGLU(h); one launch when the rows are interleaved and the problem
    fits the CTA-pair kernel (the GEGLU arithmetic rides in the epilogue, under the next tile's mainloop)."""
    M, il = x.shape[0], getattr(self, "geglu_il", 0)
    if il and FUSE_GEGLU_FWD and self.lora is None and M >= 256 and self.N % 256 == 0 and self.K >= 64:
        h = torch.empty(M, self.N, dtype=BF16, device=x.device)
        y = torch.empty(M, self.N // 2, dtype=BF16, device=x.device)
        ops.gemm(h, M, self.N, [(kmajor(x), kmajor(self.W), self.K)], bias=self.b, geglu_out=y, static_b=True, pair_mode=1)
        self.x, self.T = x, None
        return h, y
    h = self.fwd(x)
    return h, ops.geglu_fwd(h, il)


def _lin_bwd_geglu(self, dy: torch.Tensor, h: torch.Tensor, il: int = 0) -> torch.Tensor:
    """Input gradient of FeedForward.net.2 with the GEGLU backward fused into the GEMM's epilogue: returns
    dh [M, 2K] = [dY.W * gelu(gate) | dY.W * value * gelu'(gate)] for h = [value | gate]; dY.W itself is never stored.
    (Only with B200_FUSE_GEGLU=1 - see FUSE_GEGLU below; otherwise, and for small problems, the two-kernel form.)"""
    M = dy.shape[0]
    if not (FUSE_GEGLU and il == 0 and self.lora is None and M >= 256 and self.K >= 64 and self.K % 32 == 0 and self.N >= 64
            and h.shape == (M, 2 * self.K) and h.stride(1) == 1):
        return ops.geglu_bwd(self.bwd(dy), h, il)
    x = self.x
    self.x = self.T = None
    dh = torch.empty(M, 2 * self.K, dtype=BF16, device=dy.device)
    ops.gemm(dh, M, self.K, [(kmajor(dy), mnmajor(self.W), self.N)], geglu_h=h, static_b=True, pair_mode=1)
    if self.gW is not None:                            # dense fine-tune: this layer's own weight / bias gradients
        seg = (Mat(dy, M, self.N, dy.stride(0), mn=True), Mat(x, M, self.K, x.stride(0), mn=True), M)
        if _pair_wgrad_ok(self.N, self.K, M):
            ops.gemm(self.gW, self.N, self.K, [seg], d_strides=(self.K, 1, 0, 0), atomic=True, pair_mode=1)
        else:
            ops.gemm(self.gW, self.N, self.K, [seg], d_strides=(self.K, 1, 0, 0), splits=_dense_splits(self.N, self.K, M), atomic=True)
        if self.gb is not None:
            self.gb += ops.colsum(dy if dy.is_contiguous() else dy.contiguous(), 1, M, self.N)[0].float()
    return dh


# Measured on the B200 (profiles/r02h_*): correct (test_gemm2_gpu.py::test_pair_fused_geglu_backward_epilogue) but SLOWER than
# the two-kernel form, 70.0 vs 68.9 ms/step - the epilogue warps now wait on two dependent global loads of h per chunk and
# run the erf arithmetic, so the epilogue (not the 5.4 us mainloop) bounds each tile.  Off by default; kept for the A/B.
FUSE_GEGLU = os.environ.get("B200_FUSE_GEGLU", "0") == "1"
# The forward fusion needs no extra loads (both halves are in the tile's accumulator): on by default in LoRA mode
FUSE_GEGLU_FWD = os.environ.get("B200_FUSE_GEGLU_FWD", "1") != "0"


class LinQKV:
    """to_q | to_k | to_v of a self-attention block as ONE projection: y[M, 3C] = x.W_qkv^T + (s.x.A_qkv^T).B_bd^T
    with the three frozen weights stacked, the three LoRA-A factors contiguous ([3r, K]) and B_bd the block-diagonal
    [3C, 3r] arrangement of the three LoRA-B factors (a derived buffer, LoraStore.refresh_packed); the input gradient
    dx = dqkv.W_qkv + (s.dqkv.B_bd).A_qkv is ONE launch as well (side operand: the packed transpose of B_bd).  q, k, v and
    their gradients are column slices of one [M, 3C] buffer.  Needs the CTA-pair kernel (side rank 3r <= 64, M >= 256);
    smaller problems run the three projections separately."""

    def __init__(self, lq: Lin, lk: Lin, lv: Lin):
        self.lins = (lq, lk, lv)
        self.C, self.K = lq.N, lq.K
        self.W_all = torch.cat([lq.W, lk.W, lv.W], dim=0).contiguous()
        for i, l in enumerate(self.lins):                           # views: the stacked copy is the only copy
            l.W = self.W_all[i * self.C:(i + 1) * self.C]
        self.bias = None
        if any(l.b is not None for l in self.lins):
            self.bias = torch.cat([l.b if l.b is not None else torch.zeros(self.C, dtype=BF16, device=lq.W.device) for l in self.lins])
        lo = lq.lora
        self.r, self.rs, self.store = lo.r, lo.rs, lo.store
        assert self.r == self.rs and all(l.lora.offA == lo.offA + i * self.r * self.K for i, l in enumerate(self.lins))
        C, rs = self.C, self.rs
        rows = []
        for i, l in enumerate(self.lins):
            rows.append([l.lora.offB, i * C * 3 * rs + i * rs, C, rs, 3 * rs, 0])                      # B_bd [3C, 3rs]
            rows.append([l.lora.offB, 3 * C * 3 * rs + i * rs * 3 * C + i * C, C, rs, 3 * C, 1])        # Bt_bd [3rs, 3C]
        self.pack_off = self.store.add_pack(rows, 2 * 3 * C * 3 * rs)
        self.x = self.T = None

    def usable(self, M: int) -> bool:
        return M >= 256 and self.K >= 64 and self.C >= 64      # what the CTA-pair kernel takes (gemm_host.cu: pair_eligible)

    def _a_all(self):
        lo = self.lins[0].lora
        return self.store.params[lo.offA:lo.offA + 3 * self.r * self.K].view(3 * self.r, self.K)

    def _b_bd(self):
        n = 3 * self.C * 3 * self.rs
        return self.store.packed[self.pack_off:self.pack_off + n].view(3 * self.C, 3 * self.rs)

    def _bt_bd(self):
        n = 3 * self.C * 3 * self.rs
        return self.store.packed[self.pack_off + n:self.pack_off + 2 * n].view(3 * self.rs, 3 * self.C)

    def fwd(self, x: torch.Tensor) -> torch.Tensor:
        M, C, K, r3 = x.shape[0], self.C, self.K, 3 * self.r
        y = torch.empty(M, 3 * C, dtype=BF16, device=x.device)
        T = torch.empty(M, 3 * self.rs, dtype=BF16, device=x.device)
        side = (Mat(self._a_all(), r3, K, K), Mat(self._b_bd(), 3 * C, r3, 3 * self.rs), r3, self.store.scaling, T)
        ops.gemm(y, M, 3 * C, [(kmajor(x), kmajor(self.W_all), K)], bias=self.bias, side=side, static_b=True, pair_mode=1)
        self.x, self.T = x, T
        return y

    def bwd(self, dqkv: torch.Tensor) -> torch.Tensor:
        M, C, K, r, rs = dqkv.shape[0], self.C, self.K, self.r, self.rs
        x, T = self.x, self.T
        self.x = self.T = None
        U = torch.empty(M, 3 * rs, dtype=BF16, device=dqkv.device)
        dx = torch.empty(M, K, dtype=BF16, device=dqkv.device)
        side = (Mat(self._bt_bd(), 3 * r, 3 * C, 3 * C), Mat(self._a_all(), 3 * r, K, K, mn=True), 3 * r, self.store.scaling, U)
        ops.gemm(dx, M, K, [(kmajor(dqkv), mnmajor(self.W_all), 3 * C)], side=side, static_b=True, pair_mode=1)
        # six weight-gradient problems: dB_i = dqkv_i^T . T_i, dA_i = U_i^T . x  (column slices of the fused buffers)
        items = []
        for i, l in enumerate(self.lins):
            lo = l.lora
            items.append((dqkv[:, i * C:(i + 1) * C], T[:, i * rs:(i + 1) * rs], lo.gB(), M, C, r, rs, 1))
            items.append((x, U[:, i * rs:(i + 1) * rs], lo.gA(), M, K, r, 1, K))
        if WQ.active:
            WQ.items.extend(items)
        else:
            ops.lora_wgrad_batch(items)
        return dx


Lin.bwd_geglu = _lin_bwd_geglu
Lin.fwd_geglu = _lin_fwd_geglu


class Conv3:
    """Frozen 3x3 / pad 1 convolution on NHWC activations as an implicit GEMM (stride 1) or im2col GEMM (stride 2),
    optionally with the PEFT conv-LoRA side path (3x3 A: C->r, 1x1 B: r->C)."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1, lora: Optional[LoraSlot] = None,
                 need_dgrad: bool = True):
        cout, cin = w.shape[0], w.shape[1]
        self.cin, self.cout, self.stride, self.lora, self.b = cin, cout, stride, lora, b
        self.cin_p = _r8(cin)                                    # TMA needs 16-byte pixel strides
        wp = torch.zeros(cout, 3, 3, self.cin_p, dtype=BF16, device=w.device)
        wp[..., :cin] = w.permute(0, 2, 3, 1)
        self.wk = wp.reshape(cout, 9 * self.cin_p).contiguous()   # [Cout, (kh, kw, c)]
        self.wd = None
        if need_dgrad and stride == 1:
            self.cout_p = _r8(cout)
            wd = torch.zeros(cin, 3, 3, self.cout_p, dtype=BF16, device=w.device)
            wd[..., :cout] = w.flip(2, 3).permute(1, 2, 3, 0)    # dgrad = conv with flipped, transposed taps
            self.wd = wd.reshape(cin, 9 * self.cout_p).contiguous()
        self.sv = None
        self.gwk = self.gb = None                     # fp32 gradient views (DenseStore) when the layer itself trains

    def refresh_dgrad(self):
        """Dense fine-tune: the input-gradient copy of the taps follows the trained forward copy (once per step)."""
        if self.wd is not None:
            wk4 = self.wk.view(self.cout, 3, 3, self.cin_p)[..., :self.cin]
            self.wd.view(self.cin, 3, 3, self.cout_p)[..., :self.cout].copy_(wk4.flip(1, 2).permute(3, 1, 2, 0))

    def fwd(self, x: torch.Tensor, N: int, H: int, W: int, bias=None, bias_rows: int = 0,
            residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: [N*H*W, cin_p].  Returns [N*Ho*Wo, cout]."""
        assert x.shape[1] == self.cin_p, (x.shape, self.cin_p)
        s = self.stride
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        Mo = N * Ho * Wo
        bias = self.b if bias is None else bias
        lo, T, col = self.lora, None, None
        if s == 1 and ops.conv_supported(H, W):
            a0 = Conv3x3(x, N, H, W, self.cin_p, b_tap_k=self.cin_p)
        else:
            col = ops.im2col3x3(x, N, H, W, self.cin_p, s)
            a0 = kmajor(col)
        segs = [(a0, kmajor(self.wk), 9 * self.cin_p)]
        if lo is not None:
            assert s == 1
            if col is None:
                # the rank-r 3x3 convolution as ONE plain GEMM over all nine taps, Z[p, (tap, j)] = X[p, :] . A[tap, j, :]
                # (N = 9r, fp32 out), and a 9-point shift-sum: as a 16-wide implicit convolution it walked 180 k-blocks on
                # 16 CTAs (98 us per 1280-wide layer, 8 TFLOP/s - profiles/r01j_gemm_roofline_table.md)
                Z = torch.empty(Mo, 9 * lo.r, dtype=torch.float32, device=x.device)
                ops.gemm(Z, Mo, 9 * lo.r, [(Mat(x, Mo, self.cin, x.stride(0)), Mat(lo.A(), 9 * lo.r, self.cin, self.cin), self.cin)],
                         static_b=True)
                T = ops.shift_sum9(Z, N, H, W, lo.r, lo.store.scaling, lo.rs)
            else:   # unsupported image geometry: T = col . A_flat^T needs A as [r, (tap, c)]
                T = torch.empty(Mo, lo.rs, dtype=BF16, device=x.device)
                af = lo.A().view(9, lo.r, self.cin).permute(1, 0, 2).reshape(lo.r, 9 * self.cin).contiguous()
                ops.gemm(T, Mo, lo.r, [(kmajor(col), kmajor(af), 9 * self.cin)], d_strides=(lo.rs, 1, 0, 0),
                         alpha=lo.store.scaling)
            segs.append((Mat(T, Mo, lo.r, lo.rs), Mat(lo.B(), self.cout, lo.r, lo.rs), lo.r))
        ld_out = self.cout if self.cout % 8 == 0 else _r8(self.cout)
        y = (torch.empty if ld_out == self.cout else torch.zeros)(Mo, ld_out, dtype=BF16, device=x.device)
        ops.gemm(y, Mo, self.cout, segs, d_strides=(ld_out, 1, 0, 0), bias=bias, bias_rows=bias_rows,
                 bias_sb=self.cout if bias_rows else 0, residual=residual, static_b=True)
        self.sv = (x if (lo is not None or self.gwk is not None) else None, T, N, H, W)
        return y

    def bwd(self, dy: torch.Tensor, need_dx: bool = True) -> Optional[torch.Tensor]:
        """dy: [N*Ho*Wo, cout_p].  Returns dx [N*H*W, cin]."""
        x, T, N, H, W = self.sv
        self.sv = None
        s, lo = self.stride, self.lora
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        Mo, Mi = N * Ho * Wo, N * H * W
        extra = []
        if lo is not None:
            r, rs = lo.r, lo.rs
            U = torch.empty(Mo, rs, dtype=BF16, device=dy.device)
            ops.gemm(U, Mo, r, [(kmajor(dy), Mat(lo.B(), self.cout, r, rs, mn=True), self.cout)],
                     d_strides=(rs, 1, 0, 0), alpha=lo.store.scaling)
            ops.gemm(lo.gB(), self.cout, r, [(Mat(dy, Mo, self.cout, dy.stride(0), mn=True), Mat(T, Mo, r, rs, mn=True), Mo)],
                     d_strides=(rs, 1, 0, 0), splits=_wgrad_splits(self.cout, Mo), atomic=True)
            # U9[p, (tap, j)] = U[p - off(tap), j] turns both conv-LoRA backward pieces into plain GEMMs:
            #   dA[(tap, j), c] += U9^T . X          dX += U9 . A
            U9 = ops.shift_stack9(U, N, H, W, r)
            ops.gemm(lo.gA(), self.cin, 9 * r, [(Mat(x, Mi, self.cin, x.stride(0), mn=True),
                                                 Mat(U9, Mi, 9 * r, U9.stride(0), mn=True), Mi)],
                     d_strides=(1, self.cin, 0, 0), splits=_wgrad_splits(self.cin, Mi), atomic=True)
            extra = [(Mat(U9, Mi, 9 * r, U9.stride(0)), Mat(lo.A(), 9 * r, self.cin, self.cin, mn=True), 9 * r)]
        if self.gwk is not None:
            # dense fine-tune: dW[Cout, (kh,kw,c)] += dY^T . im2col(X), db += colsum(dY)
            K9 = 9 * self.cin_p
            col = ops.im2col3x3(x, N, H, W, self.cin_p, s)
            seg = (Mat(dy, Mo, self.cout, dy.stride(0), mn=True), Mat(col, Mo, K9, K9, mn=True), Mo)
            if _pair_wgrad_ok(self.cout, K9, Mo):
                ops.gemm(self.gwk, self.cout, K9, [seg], d_strides=(K9, 1, 0, 0), atomic=True, pair_mode=1)
            else:
                ops.gemm(self.gwk, self.cout, K9, [seg], d_strides=(K9, 1, 0, 0), splits=_dense_splits(self.cout, K9, Mo), atomic=True)
            del col
            if self.gb is not None:
                dyc = dy if dy.is_contiguous() else dy.contiguous()
                self.gb += ops.colsum(dyc, 1, Mo, dyc.shape[1])[0, :self.cout].float()
        if not need_dx:
            return None
        dx = torch.empty(Mi, self.cin, dtype=BF16, device=dy.device)
        if s == 1 and ops.conv_supported(H, W):
            assert dy.shape[1] == self.cout_p
            ops.gemm(dx, Mi, self.cin, [(Conv3x3(dy, N, H, W, self.cout_p, b_tap_k=self.cout_p), kmajor(self.wd),
                                         9 * self.cout_p)] + extra, static_b=True)
        else:
            dcol = torch.empty(Mo, 9 * self.cin_p, dtype=BF16, device=dy.device)
            ops.gemm(dcol, Mo, 9 * self.cin_p, [(Mat(dy, Mo, self.cout, dy.stride(0)), mnmajor(self.wk), self.cout)])
            dxc = ops.col2im3x3(dcol, N, H, W, self.cin_p, s)
            if extra:
                ops.gemm(dx, Mi, self.cin, extra, residual=dxc[:, :self.cin] if self.cin_p != self.cin else dxc)
            else:
                dx = dxc if self.cin_p == self.cin else dxc[:, :self.cin].contiguous()
        return dx


class GN:
    def __init__(self, gamma, beta, groups: int, eps: float, silu: bool):
        self.g, self.b, self.groups, self.eps, self.silu = gamma, beta, groups, eps, silu
        self.C = gamma.numel()
        self.sv = None
        self.gg = self.gbt = None                     # fp32 gradient views (DenseStore) when the layer itself trains

    def fwd(self, x, batch: int, hw: int):
        y, stats = ops.groupnorm_fwd(x, self.g, self.b, batch, hw, self.C, self.groups, self.eps, self.silu)
        self.sv = (x, stats, batch, hw)
        return y

    def bwd(self, dy, dres=None):
        x, stats, batch, hw = self.sv
        self.sv = None
        if self.gg is not None:
            ops.norm_param_grad(dy, x, self.g, self.b, stats, self.gg, self.gbt, hw=hw, groups=self.groups, silu=self.silu)
        return ops.groupnorm_bwd(dy, x, self.g, self.b, stats, batch, hw, self.C, self.groups, self.silu, dres)


class LN:
    def __init__(self, gamma, beta):
        self.g, self.b = gamma, beta
        self.sv = None
        self.gg = self.gbt = None

    def fwd(self, x):
        y, stats = ops.layernorm_fwd(x, self.g, self.b, 1e-5)
        self.sv = (x, stats)
        return y

    def bwd(self, dy, dres=None):
        x, stats = self.sv
        self.sv = None
        if self.gg is not None:
            ops.norm_param_grad(dy, x, None, None, stats, self.gg, self.gbt)
        return ops.layernorm_bwd(dy, x, self.g, stats, dres)


class Attn:
    """Multi-head attention over [B*L, C] projections; P stays in HBM for the backward."""

    def __init__(self, heads: int, to_q: Lin, to_k: Lin, to_v: Lin, to_out: Lin, cross: bool):
        self.h, self.to_q, self.to_k, self.to_v, self.to_out, self.cross = heads, to_q, to_k, to_v, to_out, cross
        self.capture = False
        self.scores = None
        self.sv = None
        self.kv_batch: Optional["CrossKVBatch"] = None     # set for cross-attention layers whose K/V are batched
        self.kv_index = -1
        self.qkv: Optional[LinQKV] = None                  # set for self-attention layers whose q|k|v run as one projection

    def fwd(self, x, ctx, B: int, L: int, Lk: int, residual):
        C = self.to_q.N
        H, d = self.h, C // self.h
        src = ctx if self.cross else x
        fused = self.qkv is not None and self.qkv.usable(x.shape[0])
        if fused:
            qkv = self.qkv.fwd(x)
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        else:
            q = self.to_q.fwd(x)
            if self.kv_batch is not None:
                k, v = self.kv_batch.kv(self.kv_index)
            else:
                k, v = self.to_k.fwd(src), self.to_v.fwd(src)
        scale = d ** -0.5
        Lp = _r8(Lk)
        dev = x.device
        P = lse = None
        fa = None
        if d == 64 or (d < 64 and d % 8 == 0):
            # fused tcgen05 attention: S and O live in TMEM, P only ever exists as a swizzled smem tile.  The kernel takes
            # 64-wide heads; narrower ones (SD1.5's d = 40 on the 64x64 latent: L = 4096, where a materialised [B,H,L,L]
            # would be gigabytes) are re-pitched to 64 with zero channels - exact for QK^T, PV and every gradient
            qf, kf, vf = (q, k, v) if d == 64 else tuple(ops.head_pad(t, H, d, 64) for t in (q, k, v))
            Of, lse = ops.flash_attn_fwd(qf, kf, vf, B, H, L, Lk, scale)
            O = Of if d == 64 else ops.head_pad(Of, H, 64, d)
            fa = (qf, kf, vf, Of)
        else:
            S = torch.empty(B, H, L, Lp, dtype=torch.float32, device=dev)
            P = torch.empty(B, H, L, Lp, dtype=BF16, device=dev)
            sS = (Lp, 1, L * Lp, H * L * Lp)
            qm = Mat(q, L, d, C, sb0=d, sb1=L * C, batched=True)
            km = Mat(k, Lk, d, C, sb0=d, sb1=Lk * C, batched=True)
            ops.gemm(S, L, Lk, [(qm, km, d)], d_strides=sS, alpha=scale, nb0=H, nb1=B)
            ops.softmax_fwd(S, P, B * H * L, Lk, Lp, Lp)
            del S
            O = torch.empty(B * L, C, dtype=BF16, device=dev)
            pm = Mat(P, L, Lk, Lp, sb0=L * Lp, sb1=H * L * Lp, batched=True)
            vm = Mat(v, Lk, d, C, mn=True, sb0=d, sb1=Lk * C, batched=True)
            ops.gemm(O, L, d, [(pm, vm, Lk)], d_strides=(C, 1, d, L * C), nb0=H, nb1=B)
        if self.capture and self.cross:
            # sum over heads of q_h.k_h / sqrt(d)  ==  (q.k over all C channels) / sqrt(d): one batched GEMM
            sc = torch.empty(B, L, Lp, dtype=BF16, device=dev)
            ops.gemm(sc, L, Lk, [(Mat(q, L, C, C, sb1=L * C, batched=True), Mat(k, Lk, C, C, sb1=Lk * C, batched=True), C)],
                     d_strides=(Lp, 1, 0, L * Lp), alpha=scale, nb0=1, nb1=B)
            self.scores = sc[:, :, :Lk]
        y = self.to_out.fwd(O, residual=residual)
        self.sv = (q, k, v, P, B, L, Lk, fa, lse, fused)
        return y

    def bwd(self, dy, d_ctx_accum, dscores: Optional[torch.Tensor]):
        q, k, v, P, B, L, Lk, fa, lse, fused = self.sv
        self.sv = None
        C = self.to_q.N
        H, d = self.h, C // self.h
        scale = d ** -0.5
        Lp = _r8(Lk)
        dev = dy.device
        dO = self.to_out.bwd(dy)
        dQ_out = dK_out = dV_out = dqkv = None
        if self.kv_batch is not None:
            dK_out, dV_out = self.kv_batch.dkv(self.kv_index)
        if fused:                                      # the three gradients land in column slices of one [M, 3C] buffer
            dqkv = torch.empty(B * L, 3 * C, dtype=BF16, device=dev)
            dQ_out, dK_out, dV_out = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:]
        dsc = None
        if dscores is not None:
            if dscores.shape[-1] == Lp and dscores.is_contiguous():
                dsc = dscores                              # already in the padded row layout (shared across layers; read-only)
            else:
                dsc = torch.zeros(B, L, Lp, dtype=BF16, device=dev)
                dsc[:, :, :Lk] = dscores
        fold = dsc is not None and fa is not None and Lk <= 128      # the hook's gradient joins dS inside the fused kernel
        if fa is not None and d == 64:
            qf, kf, vf, Of = fa
            dQ, dK, dV = ops.flash_attn_bwd(qf, kf, vf, Of, dO, lse, B, H, L, Lk, scale, dk=dK_out, dv=dV_out, dq=dQ_out,
                                            dsc=dsc if fold else None)
        elif fa is not None:
            qf, kf, vf, Of = fa
            dQf, dKf, dVf = ops.flash_attn_bwd(qf, kf, vf, Of, ops.head_pad(dO, H, d, 64), lse, B, H, L, Lk, scale,
                                               dsc=dsc if fold else None)
            dQ = ops.head_pad(dQf, H, 64, d, out=dQ_out)
            dK = ops.head_pad(dKf, H, 64, d, out=dK_out)
            dV = ops.head_pad(dVf, H, 64, d, out=dV_out)
        else:
            sS = (Lp, 1, L * Lp, H * L * Lp)
            # dV = P^T dO
            dV = torch.empty(B * Lk, C, dtype=BF16, device=dev) if dV_out is None else dV_out
            ops.gemm(dV, Lk, d, [(Mat(P, L, Lk, Lp, mn=True, sb0=L * Lp, sb1=H * L * Lp, batched=True),
                                  Mat(dO, L, d, C, mn=True, sb0=d, sb1=L * C, batched=True), L)],
                     d_strides=(C, 1, d, Lk * C), nb0=H, nb1=B)
            # dP = dO V^T ; dS = P * (dP - rowsum(P dP))
            dP = torch.empty(B, H, L, Lp, dtype=torch.float32, device=dev)
            ops.gemm(dP, L, Lk, [(Mat(dO, L, d, C, sb0=d, sb1=L * C, batched=True),
                                  Mat(v, Lk, d, C, sb0=d, sb1=Lk * C, batched=True), d)], d_strides=sS, nb0=H, nb1=B)
            dS = torch.empty(B, H, L, Lp, dtype=BF16, device=dev)
            ops.softmax_bwd(P, dP, dS, B * H * L, Lk, Lp, Lp)
            del dP, P
            # dQ = scale dS K ; dK = scale dS^T Q
            dQ = torch.empty(B * L, C, dtype=BF16, device=dev)
            ops.gemm(dQ, L, d, [(Mat(dS, L, Lk, Lp, sb0=L * Lp, sb1=H * L * Lp, batched=True),
                                 Mat(k, Lk, d, C, mn=True, sb0=d, sb1=Lk * C, batched=True), Lk)],
                     d_strides=(C, 1, d, L * C), alpha=scale, nb0=H, nb1=B)
            dK = torch.empty(B * Lk, C, dtype=BF16, device=dev) if dK_out is None else dK_out
            ops.gemm(dK, Lk, d, [(Mat(dS, L, Lk, Lp, mn=True, sb0=L * Lp, sb1=H * L * Lp, batched=True),
                                  Mat(q, L, d, C, mn=True, sb0=d, sb1=L * C, batched=True), L)],
                     d_strides=(C, 1, d, Lk * C), alpha=scale, nb0=H, nb1=B)
        if dsc is not None and not fold:
            ops.gemm(dQ, L, C, [(Mat(dsc, L, Lk, Lp, sb1=L * Lp, batched=True),
                                 Mat(k, Lk, C, C, mn=True, sb1=Lk * C, batched=True), Lk)],
                     d_strides=(C, 1, 0, L * C), alpha=scale, residual=dQ, r_strides=(C, 1, 0, L * C), nb0=1, nb1=B)
            ops.gemm(dK, Lk, C, [(Mat(dsc, L, Lk, Lp, mn=True, sb1=L * Lp, batched=True),
                                  Mat(q, L, C, C, mn=True, sb1=L * C, batched=True), L)],
                     d_strides=(C, 1, 0, Lk * C), alpha=scale, residual=dK, r_strides=(C, 1, 0, Lk * C), nb0=1, nb1=B)
        self.scores = None
        if fused:
            return self.qkv.bwd(dqkv)
        dx = self.to_q.bwd(dQ)
        if self.kv_batch is not None:
            pass                                       # dK / dV already sit in the group's buffer; handled at the end
        elif self.cross:                               # d_ctx_accum is None: nobody needs d(prompt embedding)
            self.to_k.bwd(dK, need_dx=d_ctx_accum is not None, accum=d_ctx_accum)
            self.to_v.bwd(dV, need_dx=d_ctx_accum is not None, accum=d_ctx_accum)
        else:
            self.to_k.bwd(dK, accum=dx)
            self.to_v.bwd(dV, accum=dx)
        return dx


class CrossKVBatch:
    """All cross-attention to_k / to_v projections of one channel width share the same input (the prompt embedding),
    so they run as ONE batched GEMM pair up front instead of 2 launches per layer, and their backward
    (dX, dA, dB of every layer) as four batched GEMMs at the end.  Frozen weights are stacked once; the LoRA slots of
    the group are allocated contiguously so A / B / their gradients are strided batches of the flat buffers."""

    def __init__(self, store: LoraStore, names: List[str], W_all: torch.Tensor, rank: int):
        self.store, self.names, self.W_all, self.r = store, names, W_all, rank
        self.n, self.C, self.Kc = W_all.shape
        self.slots = [store.by_name[n] for n in names] if rank > 0 else []
        if self.slots:
            self.rs = self.slots[0].rs
            self.stride = self.slots[1].offA - self.slots[0].offA if len(self.slots) > 1 else 0
            for a, b in zip(self.slots[:-1], self.slots[1:]):
                assert b.offA - a.offA == self.stride and b.offB - a.offB == self.stride
        self.sv = None

    def _ab(self, buf):
        s0 = self.slots[0]
        return buf[s0.offA:], buf[s0.offB:]

    def fwd(self, ctx2: torch.Tensor):
        M, n, C, Kc = ctx2.shape[0], self.n, self.C, self.Kc
        dev = ctx2.device
        KV = torch.empty(n, M, C, dtype=BF16, device=dev)
        segs = [(kmajor(ctx2), Mat(self.W_all, C, Kc, Kc, sb0=C * Kc, batched=True), Kc)]
        T = None
        if self.slots:
            r, rs = self.r, self.rs
            A_all, B_all = self._ab(self.store.params)
            T = torch.empty(n, M, rs, dtype=BF16, device=dev)
            ops.gemm(T, M, r, [(kmajor(ctx2), Mat(A_all, r, Kc, Kc, sb0=self.stride, batched=True), Kc)],
                     d_strides=(rs, 1, M * rs, 0), alpha=self.store.scaling, nb0=n)
            segs.append((Mat(T, M, r, rs, sb0=M * rs, batched=True), Mat(B_all, C, r, rs, sb0=self.stride, batched=True), r))
        ops.gemm(KV, M, C, segs, d_strides=(C, 1, M * C, 0), nb0=n)
        self.sv = (ctx2, T, KV, torch.empty(n, M, C, dtype=BF16, device=dev))
        return KV

    def kv(self, i: int):
        return self.sv[2][2 * i], self.sv[2][2 * i + 1]

    def dkv(self, i: int):
        return self.sv[3][2 * i], self.sv[3][2 * i + 1]

    def bwd(self, d_ctx: Optional[torch.Tensor]):
        """d_ctx None: only the LoRA weight gradients are produced (the prompt embedding needs no gradient)."""
        ctx2, T, _, dKV = self.sv
        self.sv = None
        M, n, C, Kc = ctx2.shape[0], self.n, self.C, self.Kc
        dev = ctx2.device
        dkv_k = Mat(dKV, M, C, C, sb0=M * C, batched=True)
        segs = [(dkv_k, Mat(self.W_all, C, Kc, Kc, mn=True, sb0=C * Kc, batched=True), C)]
        if self.slots:
            r, rs = self.r, self.rs
            A_all, B_all = self._ab(self.store.params)
            gA_all, gB_all = self._ab(self.store.grads)
            U = torch.empty(n, M, rs, dtype=BF16, device=dev)
            ops.gemm(U, M, r, [(dkv_k, Mat(B_all, C, r, rs, mn=True, sb0=self.stride, batched=True), C)],
                     d_strides=(rs, 1, M * rs, 0), alpha=self.store.scaling, nb0=n)
            ops.gemm(gB_all, C, r, [(Mat(dKV, M, C, C, mn=True, sb0=M * C, batched=True),
                                     Mat(T, M, r, rs, mn=True, sb0=M * rs, batched=True), M)],
                     d_strides=(rs, 1, self.stride, 0), nb0=n, atomic=True)
            ops.gemm(gA_all, Kc, r, [(Mat(ctx2, M, Kc, Kc, mn=True), Mat(U, M, r, rs, mn=True, sb0=M * rs, batched=True), M)],
                     d_strides=(1, Kc, self.stride, 0), nb0=n, atomic=True)
            segs.append((Mat(U, M, r, rs, sb0=M * rs, batched=True), Mat(A_all, r, Kc, Kc, mn=True, sb0=self.stride, batched=True), r))
        if d_ctx is None:
            return
        dctx_all = torch.empty(n, M, Kc, dtype=BF16, device=dev)
        ops.gemm(dctx_all, M, Kc, segs, d_strides=(Kc, 1, M * Kc, 0), nb0=n)
        d_ctx.add_(dctx_all.sum(dim=0))               # plumbing: one reduction over the layer axis


class TBlock:
    def __init__(self, ln1, attn1, ln2, attn2, ln3, ff1: Lin, ff2: Lin):
        self.ln1, self.attn1, self.ln2, self.attn2, self.ln3, self.ff1, self.ff2 = ln1, attn1, ln2, attn2, ln3, ff1, ff2
        self.h = None

    def fwd(self, x, ctx, B, L, Lctx):
        x1 = self.attn1.fwd(self.ln1.fwd(x), None, B, L, L, residual=x)
        x2 = self.attn2.fwd(self.ln2.fwd(x1), ctx, B, L, Lctx, residual=x1)
        h, y = self.ff1.fwd_geglu(self.ln3.fwd(x2))
        self.h = h
        return self.ff2.fwd(y, residual=x2)

    def bwd(self, dx3, d_ctx, dscores):
        dh = self.ff2.bwd_geglu(dx3, self.h, getattr(self.ff1, "geglu_il", 0))
        self.h = None
        dx2 = self.ln3.bwd(self.ff1.bwd(dh), dres=dx3)
        dx1 = self.ln2.bwd(self.attn2.bwd(dx2, d_ctx, dscores), dres=dx2)
        dx0 = self.attn1.bwd(dx1, None, None)
        WQ.flush()                                   # the block's LoRA weight gradients: one launch
        return self.ln1.bwd(dx0, dres=dx1)


class Transformer2D:
    def __init__(self, norm: GN, proj_in: Lin, blocks: List[TBlock], proj_out: Lin):
        self.norm, self.proj_in, self.blocks, self.proj_out = norm, proj_in, blocks, proj_out

    def fwd(self, x, ctx, B, HW, Lctx):
        y = self.proj_in.fwd(self.norm.fwd(x, B, HW))
        for blk in self.blocks:
            y = blk.fwd(y, ctx, B, HW, Lctx)
        return self.proj_out.fwd(y, residual=x)

    def bwd(self, dy, d_ctx, dscore_list):
        d = self.proj_out.bwd(dy)
        for blk in reversed(self.blocks):
            d = blk.bwd(d, d_ctx, dscore_list.pop() if dscore_list is not None and blk.attn2.capture else None)
        return self.norm.bwd(self.proj_in.bwd(d), dres=dy)


class Resnet:
    def __init__(self, norm1: GN, conv1: Conv3, tproj: Lin, norm2: GN, conv2: Conv3, shortcut: Optional[Lin]):
        self.norm1, self.conv1, self.tproj, self.norm2, self.conv2, self.shortcut = norm1, conv1, tproj, norm2, conv2, shortcut
        # dense fine-tune: conv1.bias and time_emb_proj.bias are two parameters that enter the graph only as their sum
        self.b_time = self.b_conv = self.g_time = self.g_conv = None

    def fwd(self, x, temb_act, N, H, W):
        hw = H * W
        if self.b_time is not None:
            self.tproj.b = ops.add(self.b_time, self.b_conv)
        tb = self.tproj.fwd(temb_act)                             # [N, Cout] = time_emb_proj(silu(emb)) + both biases
        h = self.conv1.fwd(self.norm1.fwd(x, N, hw), N, H, W, bias=tb, bias_rows=hw)
        sc = self.shortcut.fwd(x) if self.shortcut is not None else x
        return self.conv2.fwd(self.norm2.fwd(h, N, hw), N, H, W, residual=sc)

    def bwd(self, dout, d_temb_act, N, H, W):
        dh1 = self.norm2.bwd(self.conv2.bwd(dout))
        self.tproj.bwd(ops.colsum(dh1, N, H * W, dh1.shape[1]), accum=d_temb_act)
        if self.g_time is not None:                    # both biases see the gradient of their sum
            self.g_time += self.tproj.gb
            self.g_conv += self.tproj.gb
            self.tproj.gb.zero_()
        dsc = self.shortcut.bwd(dout) if self.shortcut is not None else dout
        return self.norm1.bwd(self.conv1.bwd(dh1), dres=dsc)


# =================================================================================================
# the UNet
# =================================================================================================
class UNetB200:
    def __init__(self, arch: UNetArch, state_dict: Dict[str, torch.Tensor], lora_rank: int,
                 lora_alpha_multiplier: float = 1.0, device="cuda:0", ti_elems: int = 0, lora_seed: int = 0,
                 batch_cross_kv: bool = True, dense: bool = False):
        """dense=True: full-UNet fine-tune (main.py:143-148): every parameter trains, gradients land in ``self.dense``."""
        self.arch, self.device = arch, torch.device(device)
        sd = {k.replace("base_model.model.", "").replace(".base_layer.", "."): v for k, v in state_dict.items()}
        self._sd = sd
        self.store = LoraStore(self.device, scaling=(lora_rank * lora_alpha_multiplier) / max(lora_rank, 1))
        self.rank = lora_rank
        self.dense: Optional[DenseStore] = DenseStore(self.device) if dense else None
        self._convs: List[Conv3] = []
        if dense:
            batch_cross_kv = False            # every to_k / to_v keeps its own weight (and weight gradient)
        self.hooked: List[Attn] = []
        self._cross: Dict[str, Attn] = {}
        self.kv_groups: List[CrossKVBatch] = []
        # cross-attention K/V projections all read the prompt embedding: group them by width and give each group
        # contiguous LoRA slots so they can run as strided batches (CrossKVBatch)
        groups: Dict[int, List[str]] = {}
        if batch_cross_kv:
            for key in sd:
                if key.endswith(".attn2.to_k.weight"):
                    groups.setdefault(sd[key].shape[0], []).append(key[:-len(".to_k.weight")])
            if lora_rank > 0:
                for width, paths in groups.items():
                    for pth in paths:
                        for proj in ("to_k", "to_v"):
                            w = sd[f"{pth}.{proj}.weight"]
                            self.store.add(f"{pth}.{proj}", "linear", lora_rank, w.shape[1], w.shape[0])
        # self-attention q|k|v as one projection: the three LoRA-A factors of a block are allocated back to back
        self.fuse_qkv = (os.environ.get("B200_FUSE_QKV", "1") != "0" and not dense and lora_rank > 0 and lora_rank % 8 == 0
                         and 3 * lora_rank <= 64)
        if self.fuse_qkv:
            for key in sd:
                if key.endswith(".attn1.to_q.weight"):
                    pth = key[:-len(".to_q.weight")]
                    w = sd[key]
                    self.store.add_fused([f"{pth}.to_q", f"{pth}.to_k", f"{pth}.to_v"], lora_rank, w.shape[1], [w.shape[0]] * 3)
        a, boc, g = arch, arch.block_out_channels, arch.norm_num_groups
        ted = a.time_embed_dim
        self.time1, self.time2 = self._lin("time_embedding.linear_1"), self._lin("time_embedding.linear_2")
        if a.addition_embed_type == "text_time":
            self.add1, self.add2 = self._lin("add_embedding.linear_1"), self._lin("add_embedding.linear_2")
        self.conv_in = self._conv("conv_in", need_dgrad=False)
        self.down: List[Tuple[List[Resnet], Optional[List[Transformer2D]], Optional[Conv3]]] = []
        for i in range(len(boc)):
            p = f"down_blocks.{i}"
            rs = [self._resnet(f"{p}.resnets.{j}") for j in range(a.layers_per_block)]
            at = [self._transformer(f"{p}.attentions.{j}", boc[i], a.num_attention_heads[i],
                                    a.transformer_layers_per_block[i], True) for j in range(a.layers_per_block)] \
                if a.down_has_attn[i] else None
            ds = self._conv(f"{p}.downsamplers.0.conv", stride=2) if i < len(boc) - 1 else None
            self.down.append((rs, at, ds))
        self.mid_res = [self._resnet("mid_block.resnets.0"), self._resnet("mid_block.resnets.1")]
        self.mid_attn = self._transformer("mid_block.attentions.0", boc[-1], a.num_attention_heads[-1],
                                          a.transformer_layers_per_block[-1], False)
        self.up: List[Tuple[List[Resnet], Optional[List[Transformer2D]], Optional[Conv3]]] = []
        rev_attn = list(reversed(a.down_has_attn))
        for i in range(len(boc)):
            p = f"up_blocks.{i}"
            ri = len(boc) - 1 - i
            n = a.layers_per_block + 1
            rs = [self._resnet(f"{p}.resnets.{j}") for j in range(n)]
            at = [self._transformer(f"{p}.attentions.{j}", boc[ri], a.num_attention_heads[ri],
                                    a.transformer_layers_per_block[ri], True) for j in range(n)] if rev_attn[i] else None
            us = self._conv(f"{p}.upsamplers.0.conv") if i < len(boc) - 1 else None
            self.up.append((rs, at, us))
        self.norm_out = self._gn("conv_norm_out", 1e-5, True)
        self.conv_out = self._conv("conv_out")
        for width, paths in groups.items():
            names, ws = [], []
            for i, pth in enumerate(paths):
                at = self._cross[pth]
                names += [f"{pth}.to_k", f"{pth}.to_v"]
                ws += [at.to_k.W, at.to_v.W]
            W_all = torch.stack(ws, dim=0).contiguous()
            grp = CrossKVBatch(self.store, names, W_all, lora_rank)
            for i, pth in enumerate(paths):
                at = self._cross[pth]
                at.to_k.W, at.to_v.W = W_all[2 * i], W_all[2 * i + 1]      # views: the stacked copy is the only copy
                at.kv_batch, at.kv_index = grp, i
            self.kv_groups.append(grp)
        # the reference enumerates hooked processors down_blocks first, then up_blocks (ti_cross_attn_loss.py:95-110)
        self.store.finalize(extra=ti_elems)
        if self.dense is not None:
            self.dense.finalize()
        lora_keys = [k for k in sd if ".lora_A." in k]
        if lora_keys:
            self.store.load_peft(sd)
        else:
            self.store.init_gaussian(lora_seed)
        self._sd = None
        self._fw = None

    # ---- construction helpers -------------------------------------------------------------------
    def _w(self, name: str) -> torch.Tensor:
        return self._sd[name].detach().to(self.device, BF16).contiguous()

    def _is_target(self, name: str) -> bool:
        return self.rank > 0 and any(name == t or name.endswith("." + t) for t in LORA_TARGETS)

    def _lin(self, name: str, extra_bias: Optional[torch.Tensor] = None) -> Lin:
        W = self._w(f"{name}.weight")
        if W.dim() == 4:                                          # SD1.5 1x1 conv projections / shortcuts
            W = W.reshape(W.shape[0], W.shape[1]).contiguous()
        b = self._w(f"{name}.bias") if f"{name}.bias" in self._sd else None
        if extra_bias is not None:
            b = (b.float() + extra_bias.float()).to(BF16) if b is not None else extra_bias
        lora = self.store.add(name, "linear", self.rank, W.shape[1], W.shape[0]) if self._is_target(name) else None
        lin = Lin(W, b, lora)
        if self.dense is not None:
            self.dense.add(lin, "W", "gW", f"{name}.weight", "mat", tuple(self._sd[f"{name}.weight"].shape))
            if extra_bias is not None:                 # resnet time_emb_proj: its bias is registered by _resnet
                lin.gb = torch.zeros(W.shape[0], dtype=torch.float32, device=self.device)
            elif b is not None:
                self.dense.add(lin, "b", "gb", f"{name}.bias", "vec")
        return lin

    def _gn(self, name: str, eps: float, silu: bool) -> GN:
        gn = GN(self._w(f"{name}.weight"), self._w(f"{name}.bias"), self.arch.norm_num_groups, eps, silu)
        if self.dense is not None:
            self.dense.add(gn, "g", "gg", f"{name}.weight", "vec")
            self.dense.add(gn, "b", "gbt", f"{name}.bias", "vec")
        return gn

    def _ln(self, name: str) -> LN:
        ln = LN(self._w(f"{name}.weight"), self._w(f"{name}.bias"))
        if self.dense is not None:
            self.dense.add(ln, "g", "gg", f"{name}.weight", "vec")
            self.dense.add(ln, "b", "gbt", f"{name}.bias", "vec")
        return ln

    def _conv(self, name: str, stride: int = 1, need_dgrad: bool = True, use_bias: bool = True) -> Conv3:
        w = self._w(f"{name}.weight")
        b = self._w(f"{name}.bias") if use_bias else None
        lora = self.store.add(name, "conv", self.rank, w.shape[1], w.shape[0]) if self._is_target(name) else None
        conv = Conv3(w, b, stride=stride, lora=lora, need_dgrad=need_dgrad)
        if self.dense is not None:
            self.dense.add(conv, "wk", "gwk", f"{name}.weight", "conv", (conv.cout, conv.cin, conv.cin_p))
            if b is not None:
                self.dense.add(conv, "b", "gb", f"{name}.bias", "vec")
            self._convs.append(conv)
        return conv

    def _resnet(self, p: str) -> Resnet:
        n1 = self._gn(f"{p}.norm1", 1e-5, True)
        n2 = self._gn(f"{p}.norm2", 1e-5, True)
        conv1 = self._conv(f"{p}.conv1", use_bias=False)
        tproj = self._lin(f"{p}.time_emb_proj", extra_bias=self._w(f"{p}.conv1.bias"))   # conv1 bias rides along
        conv2 = self._conv(f"{p}.conv2")
        sc = self._lin(f"{p}.conv_shortcut") if f"{p}.conv_shortcut.weight" in self._sd else None
        res = Resnet(n1, conv1, tproj, n2, conv2, sc)
        if self.dense is not None:
            res.b_time, res.b_conv = self._w(f"{p}.time_emb_proj.bias"), self._w(f"{p}.conv1.bias")
            self.dense.add(res, "b_time", "g_time", f"{p}.time_emb_proj.bias", "vec")
            self.dense.add(res, "b_conv", "g_conv", f"{p}.conv1.bias", "vec")
        return res

    def _attn(self, p: str, heads: int, cross: bool, hook: bool) -> Attn:
        at = Attn(heads, self._lin(f"{p}.to_q"), self._lin(f"{p}.to_k"), self._lin(f"{p}.to_v"),
                  self._lin(f"{p}.to_out.0"), cross)
        if cross:
            self._cross[p] = at
        if cross and hook:
            self.hooked.append(at)
        d = at.to_q.N // heads
        if not cross and self.fuse_qkv and (d == 64 or (d < 64 and d % 8 == 0)) and at.to_q.K == at.to_k.K == at.to_v.K:
            at.qkv = LinQKV(at.to_q, at.to_k, at.to_v)
        return at

    def _transformer(self, p: str, dim: int, heads: int, depth: int, hook: bool) -> Transformer2D:
        norm = self._gn(f"{p}.norm", 1e-6, False)
        blocks = []
        for j in range(depth):
            b = f"{p}.transformer_blocks.{j}"
            ff1 = self._lin(f"{b}.ff.net.0.proj")
            if self.dense is None and FUSE_GEGLU_FWD:
                _interleave_geglu_rows(ff1)
            blocks.append(TBlock(self._ln(f"{b}.norm1"),
                                 self._attn(f"{b}.attn1", heads, False, False),
                                 self._ln(f"{b}.norm2"),
                                 self._attn(f"{b}.attn2", heads, True, hook),
                                 self._ln(f"{b}.norm3"),
                                 ff1, self._lin(f"{b}.ff.net.2")))
        return Transformer2D(norm, self._lin(f"{p}.proj_in"), blocks, self._lin(f"{p}.proj_out"))

    def set_capture(self, on: bool):
        """Install / remove the reference's cross-attention score hook (init_daam_loss, main.py:50-52)."""
        for at in self.hooked:
            at.capture = on

    # ---- forward --------------------------------------------------------------------------------
    def forward(self, x8: torch.Tensor, B: int, H: int, W: int, timesteps: torch.Tensor, ctx: torch.Tensor,
                text_embeds: Optional[torch.Tensor] = None, time_ids: Optional[torch.Tensor] = None):
        """x8: noisy latents NHWC padded to 8 channels [B*H*W, 8]; ctx: [B, Lctx, Dc] bf16.
        Returns (pred [B*H*W, 8] with channels 0..3 valid, [head-summed scores [B, HW_l, Lctx]] per hooked layer)."""
        a = self.arch
        Lctx = ctx.shape[1]
        self.store.refresh_packed()                    # derived LoRA-B operands follow the optimizer's last update
        ctx2 = ctx.reshape(B * Lctx, ctx.shape[2]).contiguous()
        t_emb = ops.timestep_embedding(timesteps, a.block_out_channels[0])
        e1 = self.time1.fwd(t_emb)
        emb = self.time2.fwd(ops.silu_fwd(e1))
        a1 = None
        if a.addition_embed_type == "text_time":
            tid = ops.timestep_embedding(time_ids.flatten(), a.addition_time_embed_dim).view(B, -1)
            add_in = torch.cat([text_embeds.to(BF16), tid], dim=1).contiguous()
            a1 = self.add1.fwd(add_in)
            emb = self.add2.fwd(ops.silu_fwd(a1), residual=emb)
        temb_act = ops.silu_fwd(emb)
        for grp in self.kv_groups:
            grp.fwd(ctx2)
        x = self.conv_in.fwd(x8, B, H, W)
        skips = [x]
        dims = [(H, W)]
        h, w = H, W
        for rs, at, ds in self.down:
            for j, r in enumerate(rs):
                x = r.fwd(x, temb_act, B, h, w)
                if at is not None:
                    x = at[j].fwd(x, ctx2, B, h * w, Lctx)
                skips.append(x)
            if ds is not None:
                x = ds.fwd(x, B, h, w)
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                skips.append(x)
        x = self.mid_res[0].fwd(x, temb_act, B, h, w)
        x = self.mid_attn.fwd(x, ctx2, B, h * w, Lctx)
        x = self.mid_res[1].fwd(x, temb_act, B, h, w)
        cat_splits = []
        for rs, at, us in self.up:
            for j, r in enumerate(rs):
                sk = skips.pop()
                cat_splits.append(x.shape[1])
                x = r.fwd(torch.cat([x, sk], dim=1), temb_act, B, h, w)
                if at is not None:
                    x = at[j].fwd(x, ctx2, B, h * w, Lctx)
            if us is not None:
                x = us.fwd(ops.upsample2x_fwd(x, B, h, w, x.shape[1]), B, 2 * h, 2 * w)
                h, w = 2 * h, 2 * w
        pred = self.conv_out.fwd(self.norm_out.fwd(x, B, h * w), B, h, w)
        scores = [at_.scores for at_ in self.hooked if at_.capture]
        self._fw = (B, H, W, Lctx, ctx.shape[2], e1, emb, a1, cat_splits)
        return pred, scores

    # ---- backward -------------------------------------------------------------------------------
    def backward(self, dpred8: torch.Tensor, dscores: Optional[List[torch.Tensor]] = None, need_dctx: bool = True):
        """dpred8: [B*H*W, 8] bf16 (channels >= 4 zero).  LoRA gradients ACCUMULATE into store.grads.
        Returns (d_ctx [B, Lctx, Dc], d_text_embeds [B, P] or None); need_dctx=False (frozen text side) skips the
        input gradients of the cross-attention K/V projections and returns d_ctx = None."""
        a = self.arch
        B, H, W, Lctx, Dc, e1, emb, a1, cat_splits = self._fw
        self._fw = None
        dev = dpred8.device
        d_ctx = torch.zeros(B * Lctx, Dc, dtype=BF16, device=dev) if need_dctx else None
        d_temb_act = torch.zeros(B, a.time_embed_dim, dtype=BF16, device=dev)
        WGRAD.begin(os.environ.get("B200_WGRAD_STREAM", "0") == "1" and dev.type == "cuda")
        WQ.active, WQ.items = WQ.enabled and self.dense is None, []
        self.store.refresh_bt()                        # LoRA-B does not change between here and the optimizer
        for conv in self._convs:                       # dense fine-tune: input-gradient tap copies follow the trained taps
            conv.refresh_dgrad()
        # hooked layers were enumerated down_blocks..., up_blocks...; backward visits up (reversed) then down (reversed)
        n_down_hooks = sum(len(t.blocks) for rs, at, ds in self.down if at is not None for t in at)
        ds_down = list(dscores[:n_down_hooks]) if dscores is not None else None
        ds_up = list(dscores[n_down_hooks:]) if dscores is not None else None
        nlev = len(a.block_out_channels)
        h, w = H, W
        d = self.norm_out.bwd(self.conv_out.bwd(dpred8))
        dskips: List[torch.Tensor] = []
        for rs, at, us in reversed(self.up):
            if us is not None:
                d = ops.upsample2x_bwd(us.bwd(d), B, h // 2, w // 2, us.cin)
                h, w = h // 2, w // 2
            for j in reversed(range(len(rs))):
                if at is not None:
                    d = at[j].bwd(d, d_ctx, ds_up)
                dcat = rs[j].bwd(d, d_temb_act, B, h, w)
                c1 = cat_splits.pop()
                d = dcat[:, :c1].contiguous()
                dskips.append(dcat[:, c1:].contiguous())
        d = self.mid_res[1].bwd(d, d_temb_act, B, h, w)
        d = self.mid_attn.bwd(d, d_ctx, None)
        d = self.mid_res[0].bwd(d, d_temb_act, B, h, w)
        # the up path consumed skips last-pushed-first, so walking it backwards filled dskips in PUSH order:
        # dskips[-1] belongs to the deepest (last pushed) skip, dskips[0] to conv_in's output
        for li in reversed(range(nlev)):
            rs, at, ds = self.down[li]
            if ds is not None:
                d = ops.add(d, dskips.pop())
                d = ds.bwd(d)
                h, w = h * 2, w * 2
            for j in reversed(range(len(rs))):
                d = ops.add(d, dskips.pop())
                if at is not None:
                    d = at[j].bwd(d, d_ctx, ds_down)
                d = rs[j].bwd(d, d_temb_act, B, h, w)
        d = ops.add(d, dskips.pop())
        assert not dskips
        self.conv_in.bwd(d, need_dx=False)          # the input latents need no gradient
        for grp in self.kv_groups:
            grp.bwd(d_ctx)
        # time / added-condition embedding path (only the pooled text embedding needs a gradient)
        d_emb = ops.silu_bwd(d_temb_act, emb)
        d_text = None
        if a.addition_embed_type == "text_time":
            da1 = ops.silu_bwd(self.add2.bwd(d_emb), a1)
            d_add_in = self.add1.bwd(da1)
            d_text = d_add_in[:, :a.projection_class_embeddings_input_dim - 6 * a.addition_time_embed_dim].contiguous()
        if self.dense is not None:                   # the timestep MLP trains too: dW / db of linear_2 and linear_1
            self.time1.bwd(ops.silu_bwd(self.time2.bwd(d_emb), e1), need_dx=False)
        self.time1.x = self.time2.x = None
        WQ.flush()
        WQ.active = False
        WGRAD.join()                                 # every dA / dB has landed in store.grads before the optimizer
        return (d_ctx.view(B, Lctx, Dc) if d_ctx is not None else None), d_text
