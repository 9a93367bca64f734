"""GPU parity of the three Prodigy kernels (b200_prodigy_step) against the torch restatement of prodigyopt.Prodigy
(oracle/prodigy.py) configured as trainer/optimizer.py:22-34 does: a LoRA-like flat problem (A factors random, B factors
zero) on a quadratic bowl, each side following its own trajectory.  (Sorts late: written after the round's GPU budget
was spent - first executed by the round-end GPU run.)"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.mark.parametrize("wd,growth,n_extra", [(0.004, 1.05, 0), (0.0, float("inf"), 0), (0.004, 1.05, 1 << 20)])
def test_prodigy_kernels_track_the_package_restatement(wd, growth, n_extra):
    from oracle.prodigy import Prodigy
    from sd_lora_trainer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(16, 64), (64, 16), (9 * 8, 32), (128,)] + ([(n_extra,)] if n_extra else [])
    params = [torch.nn.Parameter((torch.randn(s, device="cuda", generator=g) * 0.05).to(BF) * float(i % 2 == 0))
              for i, s in enumerate(shapes)]
    opt = Prodigy(params, d_coef=1.0, lr=1.0, decouple=True, use_bias_correction=True, safeguard_warmup=True, weight_decay=wd,
                  betas=(0.9, 0.99), growth_rate=growth)
    flat = torch.cat([p.detach().flatten() for p in params]).clone()
    n = flat.numel()
    grads = torch.zeros(n, device="cuda")
    s, m, v = (torch.zeros(n, dtype=BF, device="cuda") for _ in range(3))
    p0 = flat.clone()
    scal = ops.prodigy_init_scalars(1e-6, "cuda")
    host, dev = torch.zeros(12).pin_memory(), torch.zeros(12, device="cuda")
    targets = [p.detach().float() + 1.0 + 0.1 * torch.randn(p.shape, device="cuda", generator=g) for p in params]
    tflat = torch.cat([t.flatten() for t in targets])
    ds = []
    for k in range(60):
        for p, t in zip(params, targets):
            p.grad = (0.05 * (p.detach().float() - t)).to(BF)
        opt.step()
        grads.copy_((0.05 * (flat.float() - tflat)).to(BF).float())
        ops.prodigy_pack_hyper(host, lr=1.0, weight_decay=wd, d_coef=1.0, growth_rate=growth, k=k)
        dev.copy_(host)
        ops.prodigy_step(flat, grads, s, p0, m, v, scal, dev, zero_grad=True)
        torch.cuda.synchronize()
        ds.append((opt.param_groups[0]["d"], float(scal[0])))
    assert float(grads.abs().max()) == 0.0
    for k, (d_ref, d_ours) in enumerate(ds):
        assert abs(d_ours - d_ref) <= 3e-2 * d_ref, (k, d_ref, d_ours)      # per-tensor bf16 .item() sums vs fp32 global sums
    if not n_extra:
        # d leaves d0 only while zero-initialised tensors dominate |s|_1; with a 1 M-element non-zero bf16 tensor the
        # package's own d stays at d0 = 1e-6 - the tracking check above is the parity statement there
        assert ds[-1][0] > 5 * ds[0][0]
    ref =torch.cat([p.detach().flatten() for p in params]).float()
    step_sz = float((ref - p0.float()).abs().max())
    assert step_sz > 0 and float((ref - flat.float()).abs().max()) <= 0.15 * step_sz
    exp_avg = torch.cat([st["exp_avg"].flatten() for st in opt.state.values()]).float()
    assert float((exp_avg - m.float()).norm() / exp_avg.norm()) < 3e-2


def test_prodigy_zero_gradient_step_is_a_no_op():
    from sd_lora_trainer_b200 import ops
    flat = torch.randn(4096, device="cuda").to(BF)
    before = flat.clone()
    grads = torch.zeros(4096, device="cuda")
    s, m, v = (torch.zeros(4096, dtype=BF, device="cuda") for _ in range(3))
    scal = ops.prodigy_init_scalars(1e-6, "cuda")
    host, dev = torch.zeros(12).pin_memory(), torch.zeros(12, device="cuda")
    ops.prodigy_pack_hyper(host, lr=1.0, weight_decay=0.01, d_coef=1.0, growth_rate=1.05, k=0)
    dev.copy_(host)
    ops.prodigy_step(flat, grads, s, flat.clone(), m, v, scal, dev)
    torch.cuda.synchronize()
    assert torch.equal(flat, before) and float(scal[6]) == 1.0 and float(scal[0]) == float(scal[1])
