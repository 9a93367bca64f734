"""GPU parity of the table-driven batched LoRA weight-gradient kernel (b200_lora_wgrad_batch) against fp32 torch matmuls:
mixed shapes in one launch, ragged M / Nout tails, odd ranks, strided operands, both output orientations (dB [N, r] and
dA [r, K]), accumulation into non-zero gradients, more problems than one launch holds."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _problem(g, M, Nout, r, ld_x=None, ld_y=None, transposed_out=False):
    ld_x = ld_x or Nout
    rs = (r + 7) // 8 * 8
    ld_y = ld_y or rs
    X = torch.randn(M, ld_x, device="cuda", generator=g).to(BF)[:, :Nout]
    Yfull = torch.randn(M, ld_y, device="cuda", generator=g).to(BF)
    Yfull[:, r:] = float("nan")                       # columns past the rank are never read into a written output
    Y = Yfull[:, :rs]
    if transposed_out:                                # dA layout: out[j, n]
        out = torch.randn(r, Nout, device="cuda", generator=g)
        sn, sj = 1, Nout
    else:                                             # dB layout: out[n, rs] (columns >= r stay untouched)
        out = torch.randn(Nout, rs, device="cuda", generator=g)
        sn, sj = rs, 1
    return X, Y, out, M, Nout, r, sn, sj


def _check(problems):
    from sd_lora_trainer_b200 import ops
    before = [p[2].clone() for p in problems]
    ops.lora_wgrad_batch(problems)
    torch.cuda.synchronize()
    for (X, Y, out, M, Nout, r, sn, sj), b in zip(problems, before):
        ref = X.float().t() @ Y[:, :r].float()        # [Nout, r]
        got = (out - b)
        got = got.t() if sn == 1 else got[:, :r]
        err = float((got - ref).abs().max()) / (float(ref.abs().max()) + 1e-6)
        assert err < 2e-3, (M, Nout, r, sn, err)      # fp32 accumulation of exact bf16 products: only summation order differs
        if sn != 1 and out.shape[1] > r:
            assert torch.equal(out[:, r:], b[:, r:])
        assert torch.isfinite(out).all()


def test_mixed_problem_table():
    g = torch.Generator(device="cuda").manual_seed(0)
    problems = [_problem(g, 2048, 1280, 16), _problem(g, 2048, 1280, 16, transposed_out=True),
                _problem(g, 8192, 640, 16), _problem(g, 8192, 640, 16, transposed_out=True),
                _problem(g, 154, 2048, 16, transposed_out=True), _problem(g, 2048, 1280, 32),
                _problem(g, 1000, 72, 5), _problem(g, 77, 8, 3, transposed_out=True),
                _problem(g, 300, 200, 8, ld_x=256, ld_y=24), _problem(g, 64, 64, 32, transposed_out=True)]
    _check(problems)


def test_more_problems_than_one_launch():
    g = torch.Generator(device="cuda").manual_seed(1)
    problems = [_problem(g, 256 + 8 * i, 128, 16, transposed_out=bool(i % 2)) for i in range(40)]
    _check(problems)


@pytest.mark.parametrize("M", [1, 63, 64, 65, 4096])
def test_reduction_length_edges(M):
    g = torch.Generator(device="cuda").manual_seed(M)
    _check([_problem(g, M, 320, 16), _problem(g, M, 320, 4, transposed_out=True)])
