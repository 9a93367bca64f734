"""Run-to-run reproducibility probe of the step on a tiny SDXL-shaped net: three identical fwd+bwd passes of one trainer;
prints the spread of the losses and of the flat LoRA gradient buffer.  Run with B200_STREAMK=0 and =1: stream-K adds its
partial tiles with bf16 TMA reduce-adds (order = arrival order), split-K weight gradients and the flash dQ path use fp32
atomics / reduce-adds, so with stream-K on the forward itself is order-dependent at the bf16 level."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.test_unet_gpu import _build, _product, rel

fam = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
cfg, orc, inputs = _build(fam, rank=8, batch=2)
tr = _product(cfg, orc)
losses, grads = [], []
for i in range(3):
    tr.store.grads.zero_()
    out = tr.step(inputs, completion_f=0.0, do_optimizer=False)
    torch.cuda.synchronize()
    losses.append((float(out["img_loss"]), float(out["tot_loss"])))
    grads.append(tr.store.grads.clone())
print("B200_STREAMK =", os.environ.get("B200_STREAMK", "1 (default)"), fam)
print("img/tot losses:", losses)
print("loss spread (rel):", max(abs(a[0] - losses[0][0]) for a in losses) / abs(losses[0][0]))
print("grad rel diff run1 vs run0:", rel(grads[1], grads[0]), " run2 vs run0:", rel(grads[2], grads[0]))
