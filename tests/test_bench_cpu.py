"""CPU test of bench.py's host-side legs: the bounded oracle baseline and the full-size step-loss delta (BASELINE.json
metric: "step-loss delta vs ref") - the GPU step on the oracle's own weights / token rows / inputs, here through
tests/cpu_mock_ops.py on the published SD1.5 graph at 64x64."""
import argparse
import os
import sys

import torch

from tests import cpu_mock_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_loss_delta_leg(monkeypatch):
    cpu_mock_ops.install(monkeypatch)
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(family="sd15", cpu_res=64, rank=4, cpu_dtype="fp32", cpu_batch=1, full_ft=False)
    keep = {}
    cb = bench.cpu_baseline(args, steps=1, warmup=0, keep=keep)
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1 and "batch 1" in cb["sample"]
    assert set(keep) >= {"orc", "cfg", "inputs", "loss", "pre", "ti"} and len(keep["pre"]) == 2 * (128 + 22)
    w = next(iter(keep["orc"].unet.parameters()))
    assert torch.equal(w.data, w.data.to(torch.bfloat16).float())           # the oracle ran on bf16-representable weights
    d = bench.step_loss_delta(keep, "cpu")
    assert d["oracle_fp32_cpu"] == keep["loss"] and d["ours"] > 0
    # bf16 kernels (here: their torch mock) vs fp32 arithmetic on identical state: the north-star bound is 1e-3
    assert d["rel"] <= 2e-3 and d["img_loss_rel"] <= 2e-3, d
